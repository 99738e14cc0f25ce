// Host-side objects behind the opaque C handles of include/openfly_b200.h.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/openfly_b200.h"
#include "device_common.cuh"
#include "setfl.hpp"

namespace ofly {

  struct Error : std::runtime_error {
    int code;
    Error(int c, std::string const &m) : std::runtime_error(m), code(c) {}
  };

  void set_last_error(std::string const &m);

#define OFLY_CUDA(expr)                                                                                        \
  do {                                                                                                         \
    cudaError_t e_ = (expr);                                                                                   \
    if (e_ != cudaSuccess) {                                                                                   \
      throw ::ofly::Error(OFLY_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));                  \
    }                                                                                                          \
  } while (0)

  // Device buffer that grows but never shrinks.
  template <typename T>
  struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    void reserve(size_t count) {
      if (count <= n) return;
      if (p) cudaFree(p);
      p = nullptr, n = 0;
      OFLY_CUDA(cudaMalloc(reinterpret_cast<void **>(&p), count * sizeof(T)));
      n = count;
    }
    void release() {
      if (p) cudaFree(p);
      p = nullptr, n = 0;
    }
  };

}  // namespace ofly

struct ofly_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  long launches = 0;
  long ticks = 0;  // search ticks executed (search.cuh)
  // dense eigenvalue solves: a few independent cuSOLVER lanes (handle + stream + workspace), created on first use
  struct SolverLane {
    void *handle = nullptr;  // cusolverDnHandle_t
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    ofly::DevBuf<double> work;
    ofly::DevBuf<int> info;
  };
  std::vector<SolverLane> solver;
  cudaEvent_t solver_fork = nullptr;
  int sm_count = 0;
  int max_smem_optin = 0;
};

struct ofly_tables {
  ofly_ctx *ctx = nullptr;
  ofly::HostTables host;
  ofly::DevBuf<double4> F, phi, V;
  ofly::TablesDev dev{};
  std::atomic<int> refs{1};
};

// The instantiated tick graph of the last search call (search_api.cu): reused while every kernel argument is unchanged.
struct ofly_search_graph {
  std::vector<char> key;  // bytes of everything the captured launches depend on
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaGraphConditionalHandle cond_while{}, cond_if[8] = {};  // one IF handle per tick of the WHILE body
  bool nested = false;
  void reset() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    exec = nullptr, graph = nullptr, key.clear();
  }
};

struct ofly_batch {
  ofly_ctx *ctx = nullptr;
  ofly_tables *tables = nullptr;
  long N = 0, R = 0;
  double extents[3] = {0, 0, 0};
  int periodic[3] = {1, 1, 1};
  double r_list = -1;  // radius the current grid was made for (-1: none yet)
  bool built = false;
  bool tiles_off = false;  // the tile path overflowed its shared-memory budget once: row kernels from then on
  int tile_wcap = 0;       // window budget reached so far (kept across grid changes)
  ofly::BatchDev d{};
  std::vector<uint8_t> type_host;

  ofly::DevBuf<uint8_t> type, frozen;
  ofly::DevBuf<double4> pos4, img_pos;
  ofly::DevBuf<uint32_t> nbr, cell_items;
  ofly::DevBuf<uint16_t> nbr16;  // tile path (tile.cuh)
  ofly::DevBuf<int> cnt, cell_of, cell_count, cell_start, fault;
  ofly::DevBuf<int> sel_all;  // [R + 1]: 0..R-1 then R -- the "every replica" selection of the grid kernels
  ofly::DevBuf<double> stage_a, stage_b, stage_c, scratch;  // [R][3N] staging for the host-pointer entry points
  ofly::DevBuf<char> search_ws;                              // workspace of the search engine
  ofly::DevBuf<double> hess_H, hess_aux, hess_mass, hess_work, hess_eig;  // dense Hessian path, kept between calls
  ofly::DevBuf<int> hess_info, hess_sel;
  std::vector<uint8_t> frozen_host;
  std::vector<ofly::DevBuf<char>> io_pool;                   // device staging of the host-pointer search calls
  std::vector<int> fault_host;
  ofly_search_graph search_graph[3];  // one per stage of the device-side loop (shrinking launch grids)
  ofly_batch *sibling = nullptr;  // same system with one more frozen atom (mech_api.cu: the relaxations of do_adj_min)
  long sibling_atom = -1;
};
