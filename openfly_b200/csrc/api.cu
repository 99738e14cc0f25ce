// C ABI implementation (include/openfly_b200.h): contexts, EAM tables, batches, neigh::List and
// potential::Generic entry points. The search entry points live in search_api.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>
#include <new>
#include <sstream>

#include "eam.cuh"
#include "eigen.cuh"
#include "hessian.cuh"
#include "internal.hpp"
#include "neigh.cuh"
#include "tile.cuh"

namespace ofly {

  static thread_local std::string g_last_error;
  void set_last_error(std::string const &m) { g_last_error = m; }

  template <typename F>
  static int guarded(F &&f) {
    try {
      f();
      return OFLY_OK;
    } catch (Error const &e) {
      set_last_error(e.what());
      return e.code;
    } catch (std::bad_alloc const &) {
      set_last_error("host allocation failed");
      return OFLY_ERR_INVALID;
    } catch (std::exception const &e) {
      set_last_error(e.what());
      return OFLY_ERR_INVALID;
    }
  }

  static void require(bool ok, char const *msg) {
    if (!ok) throw Error(OFLY_ERR_INVALID, msg);
  }

  static void use(ofly_ctx *ctx) { OFLY_CUDA(cudaSetDevice(ctx->device)); }

  static int env_int(char const *name, int dflt) {
    char const *v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
  }

  // Lanes per atom for the pair kernels: few atoms -> wide groups (more threads in flight), many atoms -> narrow
  // groups (fewer idle lanes on the ragged row tail). The sums do not depend on it (eam.cuh, VirtSum).
  int pick_group(ofly_ctx *ctx, long atoms) {
    int forced = env_int("OFLY_GROUP", 0);
    if (forced == 4 || forced == 8 || forced == 16) return forced;
    long slots = (long)ctx->sm_count * 2048;
    if (atoms * 4 >= 2 * slots) return 4;
    if (atoms * 8 >= 2 * slots) return 8;
    return 16;
  }

  template <typename K, typename... A>
  static void launch(ofly_ctx *ctx, K kernel, size_t threads_total, int block, A... args) {
    if (threads_total == 0) return;
    unsigned grid = (unsigned)((threads_total + block - 1) / block);
    kernel<<<grid, block, 0, ctx->stream>>>(args...);
    ctx->launches++;
    OFLY_CUDA(cudaGetLastError());
  }

  // Grid of a grid-stride kernel: enough blocks to fill the chip `per_sm` deep, never more than the work.
  static unsigned stride_grid(ofly_ctx *ctx, size_t blocks_of_work, int per_sm) {
    return (unsigned)std::max<size_t>(1, std::min<size_t>(blocks_of_work, (size_t)ctx->sm_count * per_sm));
  }
  static Sel all_replicas(ofly_batch *b) { return Sel{b->sel_all.p, b->sel_all.p + b->R}; }

  // ---- grid set-up: HyperGrid ctor (hypergrid.hpp:43-56) --------------------------------------------
  static void make_grid(ofly_batch *b, double r_list) {
    GridDev g{};
    require(r_list > 0, "r_cut is negative");
    for (int i = 0; i < 3; i++) {
      require(b->extents[i] > r_list, "r_cut is too big for the box");
      g.shape[i] = 2 + (int)(b->extents[i] / r_list);
      g.cell[i] = b->extents[i] / std::floor(b->extents[i] / r_list);
      g.inv_cell[i] = 1.0 / g.cell[i];
    }
    g.prod[0] = 1, g.prod[1] = g.shape[0], g.prod[2] = g.shape[0] * g.shape[1];
    g.ncells = g.shape[0] * g.shape[1] * g.shape[2];
    g.r_list = r_list;
    g.r_list_sq = r_list * r_list;
    b->d.grid = g;
    b->r_list = r_list;

    size_t R = (size_t)b->R;
    b->cell_count.reserve(R * g.ncells);
    b->cell_start.reserve(R * (g.ncells + 1));
    b->d.cell_count = b->cell_count.p, b->d.cell_start = b->cell_start.p;

    // Expected images per atom for a uniform density, with head-room; regrown on overflow.
    double ratio = 1;
    for (int i = 0; i < 3; i++) {
      if (b->periodic[i]) ratio *= 1 + std::min(2.0, 2 * r_list / b->extents[i]);
    }
    long cap = std::min<long>(27 * b->N, (long)(1.3 * ratio * (double)b->N) + 64);
    if (cap > b->d.img_cap) {
      b->cell_items.reserve(R * cap), b->img_pos.reserve(R * cap);
      b->d.img_cap = cap;
    }
    b->d.cell_items = b->cell_items.p, b->d.img_pos = b->img_pos.p;

    // Expected neighbours per atom, with head-room; regrown on overflow.
    double vol = b->extents[0] * b->extents[1] * b->extents[2];
    double expect = (double)b->N / vol * (4.0 / 3.0) * M_PI * r_list * r_list * r_list;
    int K = (int)(1.4 * expect) + 16;
    K = (K + 15) / 16 * 16;  // rows are consumed in rounds of 16 entries (eam.cuh, VirtSum)
    if (K > b->d.K) {
      b->nbr.reserve(R * (size_t)b->N * K);
      b->d.K = K;
    }
    b->d.nbr = b->nbr.p;

    // Cell-tiled fast path (tile.cuh): fully periodic boxes (every real atom then sits in a real cell, every image in the
    // ghost layer) whose 4 x 4 x 4-cell windows fit the shared-memory budget.
    TileDev &t = b->d.tile;
    t = TileDev{};
    const bool periodic = b->periodic[0] && b->periodic[1] && b->periodic[2];
    if (periodic && env_int("OFLY_TILES", 0) && !b->tiles_off) {  // opt-in: measured slower than the row kernels (profiles/)
      const double images = ratio * (double)b->N;
      const double per_cell = images / (double)g.ncells;
      long want = std::min<long>((long)(64 * per_cell * 1.35) + 64, (long)images + 64);
      want = std::max<long>(want, b->tile_wcap);
      if (want <= kTileMaxWindow) {
        for (int i = 0; i < 3; i++) t.nt[i] = (g.shape[i] - 2 + 1) / 2;
        t.n_tiles = t.nt[0] * t.nt[1] * t.nt[2];
        t.wcap = (int)((want + 63) / 64 * 64);
        b->tile_wcap = t.wcap;
        b->nbr16.reserve(R * (size_t)b->N * b->d.K);
        t.nbr16 = b->nbr16.p;
      }
    }
  }

  static void grow_after_fault(ofly_batch *b) {
    size_t R = (size_t)b->R;
    // Which bound was hit is not recorded separately; grow both (rare path).
    long cap = std::min<long>(27 * b->N, b->d.img_cap * 2);
    b->cell_items.release(), b->img_pos.release();
    b->cell_items.reserve(R * cap), b->img_pos.reserve(R * cap);
    b->d.img_cap = cap, b->d.cell_items = b->cell_items.p, b->d.img_pos = b->img_pos.p;
    int K = b->d.K * 2;
    b->nbr.release();
    b->nbr.reserve(R * (size_t)b->N * K);
    b->d.K = K, b->d.nbr = b->nbr.p;
    if (b->d.tile.n_tiles > 0) {  // the tile windows may have been what overflowed: more room, or back to the row kernels
      if (b->d.tile.wcap >= kTileMaxWindow) {
        b->tiles_off = true, b->d.tile = TileDev{};
        b->nbr16.release();
      } else {
        b->d.tile.wcap = std::min(kTileMaxWindow, (b->d.tile.wcap * 3 / 2 + 63) / 64 * 64);
        b->tile_wcap = b->d.tile.wcap;
        b->nbr16.release();
        b->nbr16.reserve(R * (size_t)b->N * K);
        b->d.tile.nbr16 = b->nbr16.p;
      }
    }
  }

  static unsigned tile_grid(ofly_batch *b) {
    return stride_grid(b->ctx, (size_t)b->R * (size_t)b->d.tile.n_tiles, env_int("OFLY_TILE_CTAS", 12));
  }

  void enqueue_rebuild(ofly_batch *b, const double *pos_dev) {
    ofly_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    size_t atoms = (size_t)b->R * b->N;
    Sel all = all_replicas(b);
    launch(ctx, k_nb_zero, std::min<size_t>((size_t)b->R * b->d.grid.ncells, 1 << 20), 256, b->d);
    k_nb_canon_count<<<stride_grid(ctx, (atoms + 255) / 256, 8), 256, 0, st>>>(b->d, all, pos_dev, (size_t)3 * b->N);
    k_nb_scan<<<stride_grid(ctx, (size_t)b->R, 2), 1024, 0, st>>>(b->d, all);
    k_nb_scatter<<<stride_grid(ctx, (atoms + 255) / 256, 8), 256, 0, st>>>(b->d, all);
    k_nb_sort_cells<<<stride_grid(ctx, ((size_t)b->R * b->d.grid.ncells + 127) / 128, 16), 128, 0, st>>>(b->d, all);
    if (b->d.tile.n_tiles > 0) {
      OFLY_CUDA(tile_prepare());
      tile_launch_build(b->d, all, tile_grid(b), st);
      launch_order_rows(b->d, all, stride_grid(ctx, (atoms + 7) / 8, 8), st);
      ctx->launches += 6;
      OFLY_CUDA(cudaGetLastError());
      return;
    }
    int G = env_int("OFLY_GROUP", 0);  // tuning override; the rows do not depend on the width
    if (!(G == 4 || G == 8 || G == 16 || G == 32)) G = build_group_for((long)atoms, (long)ctx->sm_count * 2048);
    const unsigned gg = stride_grid(ctx, (size_t)b->R * ((b->N + 256 / G - 1) / (256 / G)), 8);
    switch (G) {
      case 32:
        k_nb_build<32><<<gg, 256, 0, st>>>(b->d, all);
        break;
      case 16:
        k_nb_build<16><<<gg, 256, 0, st>>>(b->d, all);
        break;
      case 8:
        k_nb_build<8><<<gg, 256, 0, st>>>(b->d, all);
        break;
      default:
        k_nb_build<4><<<gg, 256, 0, st>>>(b->d, all);
        break;
    }
    launch_order_rows(b->d, all, stride_grid(ctx, (atoms + 7) / 8, 8), st);
    ctx->launches += 6;
    OFLY_CUDA(cudaGetLastError());
  }

  // Synchronise and fold the per-replica fault words; returns the OR of all of them.
  int collect_faults(ofly_batch *b) {
    b->fault_host.resize((size_t)b->R);
    OFLY_CUDA(cudaMemcpyAsync(b->fault_host.data(), b->fault.p, sizeof(int) * (size_t)b->R, cudaMemcpyDeviceToHost, b->ctx->stream));
    OFLY_CUDA(cudaStreamSynchronize(b->ctx->stream));
    int all = 0;
    for (int f : b->fault_host) all |= f;
    return all;
  }
  static void clear_faults(ofly_batch *b) { OFLY_CUDA(cudaMemsetAsync(b->fault.p, 0, sizeof(int) * (size_t)b->R, b->ctx->stream)); }

  void rebuild_checked(ofly_batch *b, double r_list, const double *pos_dev) {
    use(b->ctx);
    if (!(b->r_list == r_list)) make_grid(b, r_list);
    for (int attempt = 0;; attempt++) {
      enqueue_rebuild(b, pos_dev);
      int f = collect_faults(b);
      if (!(f & kFaultCapacity)) break;
      if (attempt == 3) throw Error(OFLY_ERR_CAPACITY, "neighbour buffers overflowed after regrowing");
      grow_after_fault(b);
    }
    b->built = true;
  }

  void enqueue_gradient(ofly_batch *b, double *grad_dev) {
    ofly_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    size_t atoms = (size_t)b->R * b->N;
    TablesDev const &T = b->tables->dev;
    Sel all = all_replicas(b);
    const size_t stride = (size_t)3 * b->N;
    if (b->d.tile.n_tiles > 0) {
      OFLY_CUDA(tile_prepare());
      tile_launch_gradient(b->d, T, all, tile_grid(b), st, grad_dev, stride, nullptr, 0);
      ctx->launches += 2;
      OFLY_CUDA(cudaGetLastError());
      return;
    }
    const int G = pick_group(ctx, (long)atoms);
    const unsigned gg = stride_grid(ctx, (size_t)b->R * ((b->N + 256 / G - 1) / (256 / G)), env_int("OFLY_PAIR_CTAS", 32));
    switch (G) {
      case 16:
        k_eam_rho<16><<<gg, 256, 0, st>>>(b->d, T, all);
        k_eam_force<16><<<gg, 256, 0, st>>>(b->d, T, all, grad_dev, stride, nullptr, 0);
        break;
      case 8:
        k_eam_rho<8><<<gg, 256, 0, st>>>(b->d, T, all);
        k_eam_force<8><<<gg, 256, 0, st>>>(b->d, T, all, grad_dev, stride, nullptr, 0);
        break;
      default:
        k_eam_rho<4><<<gg, 256, 0, st>>>(b->d, T, all);
        if (env_int("OFLY_FORCE_MINB", 3) == 4) {  // tuning: registers capped at 64 for four CTAs per SM
          k_eam_force<4, 4><<<gg, 256, 0, st>>>(b->d, T, all, grad_dev, stride, nullptr, 0);
        } else {
          k_eam_force<4><<<gg, 256, 0, st>>>(b->d, T, all, grad_dev, stride, nullptr, 0);
        }
        break;
    }
    ctx->launches += 2;
    OFLY_CUDA(cudaGetLastError());
  }

  // used by search_api.cu
  void make_grid_for(ofly_batch *b, double r_list) { make_grid(b, r_list); }
  void grow_buffers(ofly_batch *b) { grow_after_fault(b); }

  static void throw_on_fault(int f) {
    if (f & kFaultOutOfBounds) throw Error(OFLY_ERR_OUT_OF_BOUNDS, "x is outside tabulated region (Spline::OutOfBounds)");
    if (f & kFaultCapacity) throw Error(OFLY_ERR_CAPACITY, "neighbour capacity fault");
  }

  static size_t vec_count(ofly_batch *b) { return (size_t)b->R * 3 * (size_t)b->N; }

}  // namespace ofly

using namespace ofly;

extern "C" {

const char *ofly_last_error(void) { return g_last_error.c_str(); }
const char *ofly_version(void) { return "openfly_b200 0.1 (sm_100a)"; }

int ofly_ctx_create(int device, ofly_ctx **out) {
  return guarded([&] {
    require(out != nullptr, "null output pointer");
    int count = 0;
    OFLY_CUDA(cudaGetDeviceCount(&count));
    require(device >= 0 && device < count, "no such CUDA device");
    auto *c = new ofly_ctx;
    c->device = device;
    OFLY_CUDA(cudaSetDevice(device));
    OFLY_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    OFLY_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    OFLY_CUDA(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    *out = c;
  });
}
namespace ofly {
  void destroy_solver_lanes(ofly_ctx *ctx);  // defined with the Hessian entry points
}
void ofly_ctx_destroy(ofly_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ofly::destroy_solver_lanes(ctx);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}
void *ofly_ctx_stream(ofly_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int ofly_ctx_synchronize(ofly_ctx *ctx) {
  return guarded([&] {
    use(ctx);
    OFLY_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}
long ofly_ctx_launch_count(ofly_ctx *ctx) { return ctx ? ctx->launches : 0; }
long ofly_ctx_tick_count(ofly_ctx *ctx) { return ctx ? ctx->ticks : 0; }

// ---- tables ------------------------------------------------------------------------------------
namespace ofly {
  // Upload host tables and fill the device descriptor.
  static ofly_tables *upload_tables(ofly_ctx *ctx, HostTables &&host) {
    auto *t = new ofly_tables;
    try {
      t->ctx = ctx;
      t->host = std::move(host);
      HostTables const &h = t->host;
      static_assert(sizeof(Knot) == sizeof(double4), "knot layout");
      t->F.reserve(h.F.size()), t->phi.reserve(h.phi.size()), t->V.reserve(h.V.size());
      OFLY_CUDA(cudaMemcpy(t->F.p, h.F.data(), h.F.size() * sizeof(Knot), cudaMemcpyHostToDevice));
      OFLY_CUDA(cudaMemcpy(t->phi.p, h.phi.data(), h.phi.size() * sizeof(Knot), cudaMemcpyHostToDevice));
      OFLY_CUDA(cudaMemcpy(t->V.p, h.V.data(), h.V.size() * sizeof(Knot), cudaMemcpyHostToDevice));
      TablesDev &d = t->dev;
      d.F = t->F.p, d.phi = t->phi.p, d.V = t->V.p;
      d.n_types = h.n_types, d.kF = (int)h.knots_rho, d.kR = (int)h.knots_r;
      d.dF = h.d_rho, d.inv_dF = 1 / h.d_rho, d.dR = h.d_r, d.inv_dR = 1 / h.d_r;  // spline.cpp:27 (m_inv_dx = 1/dx)
      d.r_cut = h.r_cut, d.r_cut_sq = h.r_cut * h.r_cut;
    } catch (...) {
      delete t;
      throw;
    }
    return t;
  }
}  // namespace ofly

int ofly_tables_load_setfl(ofly_ctx *ctx, const char *path, ofly_tables **out) {
  return guarded([&] {
    require(ctx && path && out, "null argument");
    use(ctx);
    *out = upload_tables(ctx, read_setfl(std::string(path)));
  });
}
int ofly_tables_load_setfl_mem(ofly_ctx *ctx, const char *text, size_t len, ofly_tables **out) {
  return guarded([&] {
    require(ctx && text && out, "null argument");
    use(ctx);
    std::istringstream in(std::string(text, len));
    *out = upload_tables(ctx, read_setfl(in));
  });
}
int ofly_tables_create_from_knots(ofly_ctx *ctx, int n_types, const char *const *names, const double *mass, long n_knots_F, double d_rho,
                                  long n_knots_r, double d_r, double r_cut, const double *F_abcd, const double *phi_abcd,
                                  const double *V_abcd, ofly_tables **out) {
  return guarded([&] {
    require(ctx && names && mass && F_abcd && phi_abcd && V_abcd && out, "null argument");
    require(n_types >= 1 && n_types <= kMaxTypes, "unsupported number of types");
    require(n_knots_F >= 1 && n_knots_r >= 1 && d_rho > 0 && d_r > 0 && r_cut > 0, "bad tabulation");
    use(ctx);
    HostTables h;
    const size_t n = (size_t)n_types;
    h.n_types = n_types, h.knots_rho = n_knots_F, h.knots_r = n_knots_r;
    h.n_rho = n_knots_F - 4, h.n_r = n_knots_r - 4;
    h.d_rho = d_rho, h.d_r = d_r, h.r_cut = r_cut;
    for (size_t i = 0; i < n; i++) {
      require(names[i] != nullptr, "null type name");
      h.names.emplace_back(names[i]), h.mass.push_back(mass[i]);
    }
    static_assert(sizeof(Knot) == 4 * sizeof(double), "knot layout");
    h.F.resize(n * (size_t)n_knots_F), h.phi.resize(n * n * (size_t)n_knots_r), h.V.resize(n * n * (size_t)n_knots_r);
    std::memcpy(h.F.data(), F_abcd, h.F.size() * sizeof(Knot));
    std::memcpy(h.phi.data(), phi_abcd, h.phi.size() * sizeof(Knot));
    std::memcpy(h.V.data(), V_abcd, h.V.size() * sizeof(Knot));
    *out = upload_tables(ctx, std::move(h));
  });
}
const char *ofly_tables_type_name(const ofly_tables *t, int i) {
  if (!t || i < 0 || i >= t->host.n_types) return nullptr;
  return t->host.names[(size_t)i].c_str();
}
int ofly_eam_check_type_map(const ofly_tables *t, int n_types, const char *const *names) {
  return guarded([&] {
    require(t && names, "null argument");
    const int need = t->host.n_types;
    if (n_types < need) {
      throw Error(OFLY_ERR_INVALID, "Not enough types in eam data file: " + std::to_string(n_types) + " < " + std::to_string(need));
    }
    for (int i = 0; i < need; i++) {
      require(names[i] != nullptr, "null type name");
      if (t->host.names[(size_t)i] != names[i]) {
        throw Error(OFLY_ERR_INVALID, "TypeID mismatch: expecting type " + t->host.names[(size_t)i] + ", got " + names[i] + " at id=" + std::to_string(i));
      }
    }
  });
}
void ofly_tables_retain(ofly_tables *t) {
  if (t) t->refs++;
}
void ofly_tables_release(ofly_tables *t) {
  if (t && --t->refs == 0) {
    cudaSetDevice(t->ctx->device);
    t->F.release(), t->phi.release(), t->V.release();
    delete t;
  }
}
int ofly_tables_info(const ofly_tables *t, long info[3], double *dinfo) {
  return guarded([&] {
    require(t && info && dinfo, "null argument");
    info[0] = t->host.n_types, info[1] = t->host.knots_rho, info[2] = t->host.knots_r;
    dinfo[0] = t->host.d_rho, dinfo[1] = t->host.d_r, dinfo[2] = t->host.r_cut;
    for (int i = 0; i < t->host.n_types; i++) dinfo[3 + i] = t->host.mass[(size_t)i];
  });
}
int ofly_tables_knots(const ofly_tables *t, int kind, int i, int j, double *out_host) {
  return guarded([&] {
    require(t && out_host, "null argument");
    int n = t->host.n_types;
    require(i >= 0 && i < n && j >= 0 && j < n && kind >= 0 && kind <= 2, "bad table selector");
    cudaSetDevice(t->ctx->device);
    if (kind == 0) {
      OFLY_CUDA(cudaMemcpy(out_host, t->F.p + (size_t)i * t->host.knots_rho, sizeof(Knot) * t->host.knots_rho, cudaMemcpyDeviceToHost));
    } else if (kind == 1) {
      OFLY_CUDA(cudaMemcpy(out_host, t->phi.p + (size_t)(i + n * j) * t->host.knots_r, sizeof(Knot) * t->host.knots_r, cudaMemcpyDeviceToHost));
    } else {
      size_t s = (size_t)(std::max(i, j) + n * std::min(i, j));
      OFLY_CUDA(cudaMemcpy(out_host, t->V.p + s * t->host.knots_r, sizeof(Knot) * t->host.knots_r, cudaMemcpyDeviceToHost));
    }
  });
}
int ofly_setfl_info_host(const char *path, long info[3], double *dinfo) {
  return guarded([&] {
    require(path && info && dinfo, "null argument");
    HostTables h = read_setfl(path);
    info[0] = h.n_types, info[1] = h.knots_rho, info[2] = h.knots_r;
    dinfo[0] = h.d_rho, dinfo[1] = h.d_r, dinfo[2] = h.r_cut;
    for (int i = 0; i < h.n_types; i++) dinfo[3 + i] = h.mass[(size_t)i];
  });
}
int ofly_setfl_knots_host(const char *path, int kind, int i, int j, double *out_host) {
  return guarded([&] {
    require(path && out_host, "null argument");
    HostTables h = read_setfl(path);
    int n = h.n_types;
    require(i >= 0 && i < n && j >= 0 && j < n && kind >= 0 && kind <= 2, "bad table selector");
    if (kind == 0) {
      std::memcpy(out_host, h.F.data() + (size_t)i * h.knots_rho, sizeof(Knot) * h.knots_rho);
    } else if (kind == 1) {
      std::memcpy(out_host, h.phi.data() + (size_t)(i + n * j) * h.knots_r, sizeof(Knot) * h.knots_r);
    } else {
      std::memcpy(out_host, h.V.data() + (size_t)(std::max(i, j) + n * std::min(i, j)) * h.knots_r, sizeof(Knot) * h.knots_r);
    }
  });
}
double ofly_eam_r_cut(const ofly_tables *t) { return t ? t->host.r_cut : 0.0; }

// ---- batch -------------------------------------------------------------------------------------
int ofly_batch_create(ofly_ctx *ctx, ofly_tables *tables, const double extents[3], const int periodic[3], long n_atoms,
                      long n_replicas, const uint32_t *type_host, const uint8_t *frozen_host, ofly_batch **out) {
  return guarded([&] {
    require(ctx && tables && extents && periodic && type_host && frozen_host && out, "null argument");
    require(n_atoms > 0 && n_atoms < (1l << 24), "atom count must be in [1, 2^24)");
    require(n_replicas > 0 && n_replicas * n_atoms < (1l << 31), "replicas x atoms must fit in 31 bits");
    use(ctx);
    auto *b = new ofly_batch;
    try {
      b->ctx = ctx, b->tables = tables;
      ofly_tables_retain(tables);
      b->N = n_atoms, b->R = n_replicas;
      b->type_host.resize((size_t)n_atoms);
      std::vector<uint8_t> fz((size_t)n_atoms);
      for (long i = 0; i < n_atoms; i++) {
        // EAM ctor checks the type map against the table (eam.hpp:50-63); here: ids must index the table.
        require(type_host[i] < (uint32_t)tables->host.n_types, "TypeID outside the EAM table");
        b->type_host[(size_t)i] = (uint8_t)type_host[i];
        fz[(size_t)i] = frozen_host[i] ? 1 : 0;
      }
      for (int i = 0; i < 3; i++) {
        require(extents[i] > 0, "Orthorhombic extents are negative");
        b->extents[i] = extents[i], b->periodic[i] = periodic[i] ? 1 : 0;
        b->d.box.L[i] = extents[i], b->d.box.invL[i] = 1.0 / extents[i], b->d.box.periodic[i] = b->periodic[i];
      }
      size_t atoms = (size_t)b->R * b->N;
      b->type.reserve((size_t)n_atoms), b->frozen.reserve((size_t)n_atoms);
      OFLY_CUDA(cudaMemcpy(b->type.p, b->type_host.data(), (size_t)n_atoms, cudaMemcpyHostToDevice));
      OFLY_CUDA(cudaMemcpy(b->frozen.p, fz.data(), (size_t)n_atoms, cudaMemcpyHostToDevice));
      b->frozen_host = fz;
      b->pos4.reserve(atoms), b->cnt.reserve(atoms), b->cell_of.reserve(atoms), b->fault.reserve((size_t)b->R);
      {
        std::vector<int> iota((size_t)b->R + 1);
        for (long r = 0; r <= b->R; r++) iota[(size_t)r] = (int)r;  // last entry doubles as the count
        b->sel_all.reserve(iota.size());
        OFLY_CUDA(cudaMemcpy(b->sel_all.p, iota.data(), sizeof(int) * iota.size(), cudaMemcpyHostToDevice));
      }
      OFLY_CUDA(cudaMemset(b->fault.p, 0, sizeof(int) * (size_t)b->R));
      OFLY_CUDA(cudaMemset(b->cnt.p, 0, sizeof(int) * atoms));
      b->d.N = (int)b->N, b->d.R = (int)b->R, b->d.K = 0, b->d.img_cap = 0;
      b->d.type = b->type.p, b->d.frozen = b->frozen.p, b->d.pos4 = b->pos4.p, b->d.cnt = b->cnt.p, b->d.cell_of = b->cell_of.p;
      b->d.fault = b->fault.p;
    } catch (...) {
      ofly_batch_destroy(b);
      throw;
    }
    *out = b;
  });
}
void ofly_batch_destroy(ofly_batch *b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  if (b->sibling) ofly_batch_destroy(b->sibling), b->sibling = nullptr;
  for (auto &g : b->search_graph) g.reset();
  b->type.release(), b->frozen.release(), b->pos4.release(), b->img_pos.release(), b->nbr.release(), b->nbr16.release(), b->cell_items.release();
  b->cnt.release(), b->cell_of.release(), b->cell_count.release(), b->cell_start.release(), b->fault.release(), b->sel_all.release();
  b->stage_a.release(), b->stage_b.release(), b->stage_c.release(), b->scratch.release(), b->search_ws.release();
  for (auto &buf : b->io_pool) buf.release();
  b->hess_H.release(), b->hess_aux.release(), b->hess_mass.release(), b->hess_work.release(), b->hess_eig.release(), b->hess_info.release(), b->hess_sel.release();
  ofly_tables_release(b->tables);
  delete b;
}
long ofly_batch_atoms(const ofly_batch *b) { return b ? b->N : 0; }
long ofly_batch_replicas(const ofly_batch *b) { return b ? b->R : 0; }

// ---- neigh::List ---------------------------------------------------------------------------------
int ofly_neigh_rebuild_dev(ofly_batch *b, double r_list, const double *pos_dev) {
  return guarded([&] {
    require(b && pos_dev, "null argument");
    rebuild_checked(b, r_list, pos_dev);
  });
}
int ofly_neigh_rebuild(ofly_batch *b, double r_list, const double *pos_host) {
  return guarded([&] {
    require(b && pos_host, "null argument");
    use(b->ctx);
    b->stage_a.reserve(vec_count(b));
    OFLY_CUDA(cudaMemcpyAsync(b->stage_a.p, pos_host, sizeof(double) * vec_count(b), cudaMemcpyHostToDevice, b->ctx->stream));
    rebuild_checked(b, r_list, b->stage_a.p);
  });
}
int ofly_neigh_update_dev(ofly_batch *b, const double *delta_dev) {
  return guarded([&] {
    require(b && delta_dev, "null argument");
    require(b->built, "update() before rebuild()");
    use(b->ctx);
    launch(b->ctx, k_nb_update, (size_t)b->R * b->N, 256, b->d, delta_dev);
  });
}
int ofly_neigh_update(ofly_batch *b, const double *delta_host) {
  return guarded([&] {
    require(b && delta_host, "null argument");
    require(b->built, "update() before rebuild()");
    use(b->ctx);
    b->stage_a.reserve(vec_count(b));
    OFLY_CUDA(cudaMemcpyAsync(b->stage_a.p, delta_host, sizeof(double) * vec_count(b), cudaMemcpyHostToDevice, b->ctx->stream));
    launch(b->ctx, k_nb_update, (size_t)b->R * b->N, 256, b->d, (const double *)b->stage_a.p);
    OFLY_CUDA(cudaStreamSynchronize(b->ctx->stream));
  });
}
int ofly_neigh_export(ofly_batch *b, long replica, int *counts_host, int *idx_host, int *shift_host, long cap, long *total) {
  return guarded([&] {
    require(b && counts_host && idx_host && shift_host && total, "null argument");
    require(b->built, "export before rebuild()");
    require(replica >= 0 && replica < b->R, "replica out of range");
    use(b->ctx);
    OFLY_CUDA(cudaStreamSynchronize(b->ctx->stream));
    std::vector<int> cnt((size_t)b->N);
    std::vector<uint32_t> rows((size_t)b->N * b->d.K);
    OFLY_CUDA(cudaMemcpy(cnt.data(), b->cnt.p + (size_t)replica * b->N, sizeof(int) * (size_t)b->N, cudaMemcpyDeviceToHost));
    OFLY_CUDA(cudaMemcpy(rows.data(), b->nbr.p + (size_t)replica * b->N * b->d.K, sizeof(uint32_t) * rows.size(), cudaMemcpyDeviceToHost));
    long w = 0;
    for (long i = 0; i < b->N; i++) {
      counts_host[i] = cnt[(size_t)i];
      for (int k = 0; k < cnt[(size_t)i]; k++, w++) {
        if (w < cap) {
          uint32_t e = rows[(size_t)i * b->d.K + row_slot(k)];
          idx_host[w] = (int)(e & kIdxMask), shift_host[w] = (int)((e >> 24) & 31u);
        }
      }
    }
    *total = w;
  });
}
int ofly_neigh_cells(ofly_batch *b, long replica, int *cell_host) {
  return guarded([&] {
    require(b && cell_host, "null argument");
    require(b->built, "cells before rebuild()");
    require(replica >= 0 && replica < b->R, "replica out of range");
    use(b->ctx);
    OFLY_CUDA(cudaStreamSynchronize(b->ctx->stream));
    OFLY_CUDA(cudaMemcpy(cell_host, b->cell_of.p + (size_t)replica * b->N, sizeof(int) * (size_t)b->N, cudaMemcpyDeviceToHost));
  });
}
int ofly_neigh_positions(ofly_batch *b, long replica, double *pos_host) {
  return guarded([&] {
    require(b && pos_host, "null argument");
    require(b->built, "positions before rebuild()");
    require(replica >= 0 && replica < b->R, "replica out of range");
    use(b->ctx);
    OFLY_CUDA(cudaStreamSynchronize(b->ctx->stream));
    std::vector<double4> p((size_t)b->N);
    OFLY_CUDA(cudaMemcpy(p.data(), b->pos4.p + (size_t)replica * b->N, sizeof(double4) * (size_t)b->N, cudaMemcpyDeviceToHost));
    for (long i = 0; i < b->N; i++) pos_host[3 * i] = p[(size_t)i].x, pos_host[3 * i + 1] = p[(size_t)i].y, pos_host[3 * i + 2] = p[(size_t)i].z;
  });
}
int ofly_neigh_mean_count(ofly_batch *b, double *mean) {
  return guarded([&] {
    require(b && mean, "null argument");
    require(b->built, "mean_count before rebuild()");
    use(b->ctx);
    OFLY_CUDA(cudaStreamSynchronize(b->ctx->stream));
    std::vector<int> cnt((size_t)b->R * b->N);
    OFLY_CUDA(cudaMemcpy(cnt.data(), b->cnt.p, sizeof(int) * cnt.size(), cudaMemcpyDeviceToHost));
    double s = 0;
    for (int c : cnt) s += c;
    *mean = s / (double)cnt.size();
  });
}

// ---- potential::Generic --------------------------------------------------------------------------
int ofly_eam_gradient_dev(ofly_batch *b, double *grad_dev) {
  return guarded([&] {
    require(b && grad_dev, "null argument");
    require(b->built, "gradient() before rebuild()");
    require(b->tables->host.r_cut <= b->r_list, "r_cut is bigger than the internal r_cut");  // list.hpp:141-143
    use(b->ctx);
    enqueue_gradient(b, grad_dev);
  });
}
int ofly_eam_gradient(ofly_batch *b, double *grad_host) {
  return guarded([&] {
    require(b && grad_host, "null argument");
    require(b->built, "gradient() before rebuild()");
    require(b->tables->host.r_cut <= b->r_list, "r_cut is bigger than the internal r_cut");
    use(b->ctx);
    b->stage_b.reserve(vec_count(b));
    clear_faults(b);
    enqueue_gradient(b, b->stage_b.p);
    OFLY_CUDA(cudaMemcpyAsync(grad_host, b->stage_b.p, sizeof(double) * vec_count(b), cudaMemcpyDeviceToHost, b->ctx->stream));
    throw_on_fault(collect_faults(b));
  });
}
namespace ofly {
  // energy (eam.cpp:33-51) of every replica over the current list -> energy_dev[R]; faults are left in b->fault.
  void energies_dev(ofly_batch *b, double *energy_dev) {
    require(b->built, "energy() before rebuild()");
    require(b->tables->host.r_cut <= b->r_list, "r_cut is bigger than the internal r_cut");
    use(b->ctx);
    size_t atoms = (size_t)b->R * b->N;
    b->scratch.reserve(atoms + (size_t)b->R);
    clear_faults(b);
    TablesDev const &T = b->tables->dev;
    switch (pick_group(b->ctx, (long)atoms)) {
      case 16:
        launch(b->ctx, k_eam_energy_atoms<16>, atoms * 16, 256, b->d, T, b->scratch.p);
        break;
      case 8:
        launch(b->ctx, k_eam_energy_atoms<8>, atoms * 8, 256, b->d, T, b->scratch.p);
        break;
      default:
        launch(b->ctx, k_eam_energy_atoms<4>, atoms * 4, 256, b->d, T, b->scratch.p);
        break;
    }
    k_reduce_rows<<<(unsigned)b->R, 1024, 0, b->ctx->stream>>>(b->scratch.p, (int)b->N, energy_dev);
    b->ctx->launches++;
    OFLY_CUDA(cudaGetLastError());
  }
}  // namespace ofly

int ofly_eam_energy(ofly_batch *b, double *energy_host) {
  return guarded([&] {
    require(b && energy_host, "null argument");
    size_t atoms = (size_t)b->R * b->N;
    b->scratch.reserve(atoms + (size_t)b->R);
    energies_dev(b, b->scratch.p + atoms);
    OFLY_CUDA(cudaMemcpyAsync(energy_host, b->scratch.p + atoms, sizeof(double) * (size_t)b->R, cudaMemcpyDeviceToHost, b->ctx->stream));
    throw_on_fault(collect_faults(b));
  });
}
namespace ofly {
  // Dense Hessians of `count` replicas of the batch (ids in replicas_host) assembled into b->hess_H: matrix s at
  // hess_H + s (3N)^2, column-major, zero-initialised here. One launch per pass for the whole set.
  void assemble_hessians(ofly_batch *b, const int *replicas_host, int count, int mass_weight) {
    require(b->built, "hessian() before rebuild()");
    require(count >= 1 && count <= b->R, "replica count out of range");
    for (int s = 0; s < count; s++) require(replicas_host[s] >= 0 && replicas_host[s] < b->R, "replica out of range");
    require(b->tables->host.r_cut <= b->r_list, "r_cut is bigger than the internal r_cut");
    use(b->ctx);
    const size_t n3 = 3 * (size_t)b->N, hs = n3 * n3;
    b->hess_H.reserve(hs * (size_t)count), b->hess_aux.reserve(4 * (size_t)b->N * (size_t)count), b->hess_mass.reserve(kMaxTypes);
    b->hess_sel.reserve((size_t)b->R);
    std::vector<double> m(kMaxTypes, 1.0);
    for (int t = 0; t < b->tables->host.n_types; t++) m[(size_t)t] = b->tables->host.mass[(size_t)t];
    cudaStream_t st = b->ctx->stream;
    OFLY_CUDA(cudaMemcpyAsync(b->hess_mass.p, m.data(), sizeof(double) * kMaxTypes, cudaMemcpyHostToDevice, st));
    OFLY_CUDA(cudaMemcpyAsync(b->hess_sel.p, replicas_host, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, st));
    OFLY_CUDA(cudaStreamSynchronize(st));  // both sources are the caller's / locals
    OFLY_CUDA(cudaMemsetAsync(b->hess_H.p, 0, sizeof(double) * hs * (size_t)count, st));
    clear_faults(b);
    const size_t rows = (size_t)b->N * (size_t)count;
    launch(b->ctx, k_hess_rho_mu<8>, rows * 8, 256, b->d, b->tables->dev, (const int *)b->hess_sel.p, count, b->hess_aux.p);
    launch(b->ctx, k_hess_columns<32>, rows * 32, 256, b->d, b->tables->dev, (const int *)b->hess_sel.p, count, (const double *)b->hess_aux.p,
           b->hess_H.p, hs);
    if (mass_weight) {
      launch(b->ctx, k_hess_mass_weight, n3 * n3, 256, b->hess_H.p, hs, count, (int)b->N, (const uint8_t *)b->type.p, (const double *)b->hess_mass.p);
    }
  }

  // cuSOLVER is loaded on first use (dlopen), so the library itself carries no load-time dependency on it.
  struct Cusolver {
    void *lib = nullptr;
    int (*create)(void **) = nullptr;
    int (*destroy)(void *) = nullptr;
    int (*set_stream)(void *, cudaStream_t) = nullptr;
    int (*syevd_size)(void *, int, int, int, const double *, int, const double *, int *) = nullptr;
    int (*syevd)(void *, int, int, int, double *, int, double *, double *, int, int *) = nullptr;
    bool ok() const { return lib && create && destroy && set_stream && syevd_size && syevd; }
  };
  static Cusolver &cusolver() {
    static Cusolver c = [] {
      Cusolver x;
      for (char const *name : {"libcusolver.so.11", "libcusolver.so", "libcusolver.so.12"}) {
        x.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (x.lib) break;
      }
      if (x.lib) {
        x.create = reinterpret_cast<decltype(x.create)>(dlsym(x.lib, "cusolverDnCreate"));
        x.destroy = reinterpret_cast<decltype(x.destroy)>(dlsym(x.lib, "cusolverDnDestroy"));
        x.set_stream = reinterpret_cast<decltype(x.set_stream)>(dlsym(x.lib, "cusolverDnSetStream"));
        x.syevd_size = reinterpret_cast<decltype(x.syevd_size)>(dlsym(x.lib, "cusolverDnDsyevd_bufferSize"));
        x.syevd = reinterpret_cast<decltype(x.syevd)>(dlsym(x.lib, "cusolverDnDsyevd"));
      }
      return x;
    }();
    return c;
  }
  void destroy_solver_lanes(ofly_ctx *ctx) {
    for (auto &ln : ctx->solver) {
      if (ln.handle && cusolver().ok()) cusolver().destroy(ln.handle);
      if (ln.stream) cudaStreamDestroy(ln.stream);
      if (ln.done) cudaEventDestroy(ln.done);
      ln.work.release(), ln.info.release();
    }
    ctx->solver.clear();
    if (ctx->solver_fork) cudaEventDestroy(ctx->solver_fork);
    ctx->solver_fork = nullptr;
  }

  // Hessian::eigenvalues() (hessian.hpp:86-93) of `count` dense symmetric n3 x n3 matrices on the device (lower triangle
  // read, matrices destroyed), ascending, into eig_dev[count][n3].
  // Batches go to the own kernel (eigen.cuh: one CTA per matrix, all matrices in flight, HBM-bound); a few large matrices
  // alone would leave most of the chip idle there and go to cuSOLVER's Dsyevd instead (a library call, loaded on first
  // use; spread over a few solver lanes). OFLY_EIGEN=own / cusolver forces either.
  void eigenvalues_dev(ofly_ctx *ctx, double *H_dev, int n3, int count, double *eig_dev) {
    char const *mode = std::getenv("OFLY_EIGEN");
    const bool force_own = mode && std::string(mode) == "own", force_lib = mode && std::string(mode) == "cusolver";
    if (n3 <= kEigMaxN && !force_lib && (force_own || count >= env_int("OFLY_EIGEN_MIN_BATCH", 6))) {
      static unsigned long long allowed = 0;
      const unsigned long long bit = 1ull << (ctx->device & 63);
      if (!(allowed & bit)) {
        OFLY_CUDA(cudaFuncSetAttribute(k_sym_eigenvalues, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eig_smem_bytes(kEigMaxN)));
        allowed |= bit;
      }
      DevBuf<double> de;
      de.reserve((size_t)count * 2 * (size_t)n3);
      k_sym_eigenvalues<<<(unsigned)count, kEigThreads, eig_smem_bytes(n3), ctx->stream>>>(H_dev, (size_t)n3 * (size_t)n3, n3, eig_dev, de.p);
      ctx->launches++;
      cudaError_t e = cudaGetLastError();
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      de.release();
      OFLY_CUDA(e);
      if (std::getenv("OFLY_VERBOSE")) std::fprintf(stderr, "openfly_b200: eigenvalues of %d matrices (n = %d) by the own kernel\n", count, n3);
      return;
    }
    Cusolver &cs = cusolver();
    if (!cs.ok()) throw Error(OFLY_ERR_CUDA, "libcusolver could not be loaded (needed for the dense symmetric eigenvalue solve)");
    if (std::getenv("OFLY_VERBOSE")) std::fprintf(stderr, "openfly_b200: eigenvalues of %d matrices (n = %d) by cuSOLVER\n", count, n3);
    const int lanes = std::max(1, std::min(std::min(count, env_int("OFLY_SOLVER_LANES", 4)), 16));
    if (!ctx->solver_fork) OFLY_CUDA(cudaEventCreateWithFlags(&ctx->solver_fork, cudaEventDisableTiming));
    while ((int)ctx->solver.size() < lanes) {
      ofly_ctx::SolverLane ln;
      if (cs.create(&ln.handle) != 0) throw Error(OFLY_ERR_CUDA, "cusolverDnCreate failed");
      ctx->solver.push_back(ln);
      auto &l = ctx->solver.back();
      OFLY_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
      OFLY_CUDA(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
      if (cs.set_stream(l.handle, l.stream) != 0) throw Error(OFLY_ERR_CUDA, "cusolverDnSetStream failed");
    }
    const size_t hs = (size_t)n3 * (size_t)n3;
    int lwork = 0;
    // jobz = CUSOLVER_EIG_MODE_NOVECTOR (0), uplo = CUBLAS_FILL_MODE_LOWER (0): the lower triangle is what system::Hessian
    // fills and what Eigen's SelfAdjointEigenSolver reads (hessian.hpp:86-93)
    if (cs.syevd_size(ctx->solver[0].handle, 0, 0, n3, H_dev, n3, eig_dev, &lwork) != 0) throw Error(OFLY_ERR_CUDA, "cusolverDnDsyevd_bufferSize failed");
    OFLY_CUDA(cudaEventRecord(ctx->solver_fork, ctx->stream));
    for (int l = 0; l < lanes; l++) {
      auto &ln = ctx->solver[(size_t)l];
      ln.work.reserve((size_t)std::max(lwork, 1)), ln.info.reserve((size_t)count);
      OFLY_CUDA(cudaStreamWaitEvent(ln.stream, ctx->solver_fork, 0));
    }
    for (int s = 0; s < count; s++) {
      auto &ln = ctx->solver[(size_t)(s % lanes)];
      if (cs.syevd(ln.handle, 0, 0, n3, H_dev + (size_t)s * hs, n3, eig_dev + (size_t)s * n3, ln.work.p, lwork, ln.info.p + s) != 0)
        throw Error(OFLY_ERR_CUDA, "cusolverDnDsyevd failed");
    }
    std::vector<int> info((size_t)count, 0);
    for (int l = 0; l < lanes; l++) {
      auto &ln = ctx->solver[(size_t)l];
      OFLY_CUDA(cudaEventRecord(ln.done, ln.stream));
      OFLY_CUDA(cudaStreamWaitEvent(ctx->stream, ln.done, 0));
    }
    OFLY_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int l = 0; l < lanes; l++) {
      auto &ln = ctx->solver[(size_t)l];
      for (int s = l; s < count; s += lanes) OFLY_CUDA(cudaMemcpy(&info[(size_t)s], ln.info.p + s, sizeof(int), cudaMemcpyDeviceToHost));
    }
    for (int s = 0; s < count; s++) {
      if (info[(size_t)s] != 0) throw Error(OFLY_ERR_CUDA, "cusolverDnDsyevd did not converge (info = " + std::to_string(info[(size_t)s]) + ")");
    }
    ctx->launches += count;
  }
}  // namespace ofly

int ofly_eam_hessian(ofly_batch *b, long replica, int mass_weight, double *hess_host) {
  return guarded([&] {
    require(b && hess_host, "null argument");
    const int r = (int)replica;
    require(replica >= 0 && replica < b->R, "replica out of range");
    assemble_hessians(b, &r, 1, mass_weight);
    size_t n3 = 3 * (size_t)b->N;
    OFLY_CUDA(cudaMemcpyAsync(hess_host, b->hess_H.p, sizeof(double) * n3 * n3, cudaMemcpyDeviceToHost, b->ctx->stream));
    throw_on_fault(collect_faults(b));
  });
}

int ofly_eam_hessian_batch(ofly_batch *b, int mass_weight, double *hess_host) {
  return guarded([&] {
    require(b && hess_host, "null argument");
    std::vector<int> all((size_t)b->R);
    for (long r = 0; r < b->R; r++) all[(size_t)r] = (int)r;
    assemble_hessians(b, all.data(), (int)b->R, mass_weight);
    size_t n3 = 3 * (size_t)b->N;
    OFLY_CUDA(cudaMemcpyAsync(hess_host, b->hess_H.p, sizeof(double) * n3 * n3 * (size_t)b->R, cudaMemcpyDeviceToHost, b->ctx->stream));
    throw_on_fault(collect_faults(b));
  });
}

int ofly_eam_hessian_eigenvalues(ofly_batch *b, long replica, int mass_weight, double *eig_host) {
  return guarded([&] {
    require(b && eig_host, "null argument");
    const int r = (int)replica;
    require(replica >= 0 && replica < b->R, "replica out of range");
    assemble_hessians(b, &r, 1, mass_weight);
    throw_on_fault(collect_faults(b));
    const int n3 = 3 * (int)b->N;
    b->hess_eig.reserve((size_t)n3);
    eigenvalues_dev(b->ctx, b->hess_H.p, n3, 1, b->hess_eig.p);
    OFLY_CUDA(cudaMemcpyAsync(eig_host, b->hess_eig.p, sizeof(double) * (size_t)n3, cudaMemcpyDeviceToHost, b->ctx->stream));
    OFLY_CUDA(cudaStreamSynchronize(b->ctx->stream));
  });
}

// ---- measurement aid: achievable FP64 FMA rate (the co-limiter of the pair kernels) ---------
namespace ofly {
  // 8 independent FMA chains per thread, registers only: DFMA issue rate of the whole chip.
  static __global__ void __launch_bounds__(256) k_fp64_fma(int iters, double seed, double *sink) {
    double a[8];
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = seed + (double)(threadIdx.x + q);
    const double m = 1.0 - 1e-9, c = 1e-9;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] = fma(a[q], m, c);
      }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) s += a[q];
    if (s == 123.456) sink[0] = s;  // never true: keeps the chains alive
  }
}  // namespace ofly

int ofly_bench_fp64_fma(ofly_ctx *ctx, int iters, double *fma_per_s) {
  return guarded([&] {
    require(ctx && fma_per_s && iters > 0, "bad argument");
    use(ctx);
    DevBuf<double> sink;
    sink.reserve(1);
    const unsigned grid = (unsigned)ctx->sm_count * 8;
    cudaEvent_t e0, e1;
    OFLY_CUDA(cudaEventCreate(&e0));
    OFLY_CUDA(cudaEventCreate(&e1));
    k_fp64_fma<<<grid, 256, 0, ctx->stream>>>(iters / 8 + 1, 1.0, sink.p);  // warm-up
    OFLY_CUDA(cudaEventRecord(e0, ctx->stream));
    k_fp64_fma<<<grid, 256, 0, ctx->stream>>>(iters, 1.0, sink.p);
    OFLY_CUDA(cudaEventRecord(e1, ctx->stream));
    OFLY_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    OFLY_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    sink.release();
    ctx->launches += 2;
    *fma_per_s = (double)grid * 256.0 * 32.0 * (double)iters / ((double)ms * 1e-3);
  });
}

// ---- raw device memory helpers -------------------------------------------------------------------
int ofly_dev_alloc(ofly_ctx *ctx, size_t bytes, void **out_dev) {
  return guarded([&] {
    require(ctx && out_dev, "null argument");
    use(ctx);
    OFLY_CUDA(cudaMalloc(out_dev, bytes));
  });
}
int ofly_dev_free(ofly_ctx *ctx, void *dev) {
  return guarded([&] {
    require(ctx, "null argument");
    use(ctx);
    OFLY_CUDA(cudaFree(dev));
  });
}
int ofly_dev_upload(ofly_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes) {
  return guarded([&] {
    require(ctx && dst_dev && src_host, "null argument");
    use(ctx);
    OFLY_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    OFLY_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}
int ofly_dev_download(ofly_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes) {
  return guarded([&] {
    require(ctx && dst_host && src_dev, "null argument");
    use(ctx);
    OFLY_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    OFLY_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

}  // extern "C"
