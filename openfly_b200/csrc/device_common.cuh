// Device-side data descriptors and small helpers shared by every kernel of the path.
//
// HBM layout of a batch (R replicas x N atoms, see DESIGN.md "Data layout"):
//   pos4   [R][N]    double4  list-space position {x,y,z} + w = F'(rho) written by EAM pass 1,
//                             32 B aligned so one 256-bit load fetches a neighbour's whole record
//   nbr    [R][N][K] uint32   neighbour entries, atom-major rows: bits 0-23 real index j,
//                             bits 24-28 image shift code (sx+1)+3(sy+1)+9(sz+1), bits 29-31 type of j.
//                             A row is ordered by (j, shift) and stored in rounds of 16 entries, each round
//                             transposed as a 4 x 4 block: the k-th neighbour lives at row[row_slot(k)].
//                             Ghost atoms are never materialised after the build: the image position
//                             is pos4[j] + shift*L, which is exactly how the reference created it.
//   cnt    [R][N]    int32    stored neighbours per atom
//   tables F/phi/V   double4  spline knots {a,b,c,d}, one 32 B sector per knot, L2 resident
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ofly {

  constexpr int kShiftCentre = 13;     // (0,0,0) image
  constexpr uint32_t kIdxMask = 0x00FFFFFFu;
  constexpr int kMaxTypes = 8;

  // Storage slot of the k-th neighbour of a row (and, the map being an involution, the rank of the entry in slot k).
  // Rounds of 16 entries are stored as transposed 4 x 4 blocks: the four lanes of a sub-warp that vector-load slots
  // 4t .. 4t+3 of a round work, in their m-th pair, on the CONSECUTIVE ranks 4m .. 4m+3. Rows are ordered by neighbour
  // index, so the 32 records one warp-wide load instruction gathers (8 consecutive atoms x 4 consecutive neighbours) sit
  // in a few 128 B lines instead of one line per lane: the L1 data pipe serves a load per distinct line.
  __host__ __device__ constexpr int row_slot(int k) { return (k & ~15) | ((k & 3) << 2) | ((k >> 2) & 3); }
  static_assert(row_slot(0) == 0 && row_slot(1) == 4 && row_slot(4) == 1 && row_slot(7) == 13 && row_slot(15) == 15, "4 x 4 transpose of a round");
  static_assert(row_slot(16 + 6) == 16 + 9 && row_slot(row_slot(37)) == 37 && row_slot(row_slot(100)) == 100, "rounds stay in place, the map is an involution");

  // Fault bits recorded per replica.
  constexpr int kFaultOutOfBounds = 1;  // Spline::OutOfBounds
  constexpr int kFaultCapacity = 2;     // neighbour row or image buffer overflow

  struct BoxDev {
    double L[3];
    double invL[3];
    int periodic[3];
  };

  struct GridDev {
    int shape[3];
    int prod[3];
    double cell[3];
    double inv_cell[3];
    double r_list;
    double r_list_sq;  // r_list * r_list rounded once, as the reference's m_r_cut * m_r_cut
    int ncells;
  };

  struct TablesDev {
    const double4 *F;    // [type][kF]
    const double4 *phi;  // [src + n*dst][kR]
    const double4 *V;    // [max + n*min][kR]
    int n_types;
    int kF, kR;  // knots per spline
    double dF, inv_dF, dR, inv_dR;
    double r_cut, r_cut_sq;
  };

  // Cell-tiled fast path (tile.cuh): tiles of 2 x 2 x 2 real cells; n_tiles == 0 switches it off.
  struct TileDev {
    int nt[3];        // tiles per axis
    int n_tiles;
    int wcap;         // window slots a CTA may stage (shared-memory budget)
    uint16_t *nbr16;  // [R][N][K] neighbour entries as window slots, same rows as nbr
  };

  struct BatchDev {
    int N, R, K;
    long img_cap;  // image slots per replica
    BoxDev box;
    GridDev grid;
    const uint8_t *type;    // [N]
    const uint8_t *frozen;  // [N]
    double4 *pos4;          // [R][N]
    uint32_t *nbr;          // [R][N][K]
    int *cnt;               // [R][N]
    int *cell_of;           // [R][N]   HyperGrid::cell_idx of each real atom
    int *cell_count;        // [R][ncells]      scratch: histogram / fill cursor
    int *cell_start;        // [R][ncells + 1]  exclusive scan
    uint32_t *cell_items;   // [R][img_cap]     (j | code << 24) sorted inside each cell
    double4 *img_pos;       // [R][img_cap]     image position, w = bit pattern of the entry
    int *fault;             // [R]
    TileDev tile;
  };

  // ---- 256-bit global loads (sm_100: LDG.E.256) ----------------------------------------------
  __device__ __forceinline__ double4 ld256(const double4 *p) {
    double4 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
  }
  // read-only (non-coherent) path for the immutable tables
  __device__ __forceinline__ double4 ld256_nc(const double4 *p) {
    double4 v;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
  }
  __device__ __forceinline__ void st256(double4 *p, double4 v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
  }

  // 1/sqrt(x) for positive normal x: hardware seed (MUFU.RSQ64H) + one third-order correction,
  // no special-case branch. Relative error ~1e-16.
  __device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(x, -(y * y), 1.0);
    double p = fma(e, 0.375, 0.5);
    return fma(p, y * e, y);
  }

  // Image shift along one axis from its base-3 digit: 0 -> -L, 1 -> 0, 2 -> +L.
  __device__ __forceinline__ double shift_of(int digit, double L) { return digit == 1 ? 0.0 : (digit == 0 ? -L : L); }

  // HyperGrid::cell_idx (hypergrid.hpp:75-99): trunc((x + cell) * inv_cell) clamped, x fastest.
  // Explicit round-to-nearest intrinsics so no FMA contraction can change a boundary decision.
  __device__ __forceinline__ int cell_index(const GridDev &g, double x, double y, double z) {
    int kx = __double2int_rz(__dmul_rn(__dadd_rn(x, g.cell[0]), g.inv_cell[0]));
    int ky = __double2int_rz(__dmul_rn(__dadd_rn(y, g.cell[1]), g.inv_cell[1]));
    int kz = __double2int_rz(__dmul_rn(__dadd_rn(z, g.cell[2]), g.inv_cell[2]));
    kx = min(max(kx, 0), g.shape[0] - 1);
    ky = min(max(ky, 0), g.shape[1] - 1);
    kz = min(max(kz, 0), g.shape[2] - 1);
    return kx * g.prod[0] + ky * g.prod[1] + kz * g.prod[2];
  }

  // gnorm_sq of a 3-vector in the reference's order, (x*x + y*y) + z*z, without contraction.
  __device__ __forceinline__ double norm_sq_rn(double dx, double dy, double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  }

  // Lanes of the calling thread's G-wide (power of two) sub-warp.
  template <int G>
  __device__ __forceinline__ unsigned group_mask() {
    if constexpr (G == 32) {
      return 0xffffffffu;
    } else {
      return ((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G);
    }
  }

  // Sum over the G lanes of a sub-warp in a fixed butterfly order; every lane gets the result.
  // Only the sub-warp's own lanes are named in the mask, so sub-warps may diverge or exit freely.
  template <int G>
  __device__ __forceinline__ double group_sum(double v) {
    const unsigned m = group_mask<G>();
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      v += __shfl_xor_sync(m, v, o);
    }
    return v;
  }

}  // namespace ofly
