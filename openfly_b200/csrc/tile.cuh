// Cell-tiled neighbour build and EAM passes: the fast path for fully periodic boxes.
//
// The row kernels (neigh.cuh / eam.cuh) gather every neighbour's 32 B record from global memory: 58 gathers per atom and
// pass, one L1 wavefront each -- the L1 data pipe, not HBM or the FP64 pipe, is what bounds them (profiles/). Here one CTA
// owns a *tile* of 2 x 2 x 2 grid cells (~128 atoms). Everything those atoms can interact with lies in the 4 x 4 x 4-cell
// *window* around the tile (~1000 images: cells are at least one list radius wide), and in the cell-sorted image list of
// the counting sort (neigh.cuh) every row of window cells along x is one contiguous run of slots. The CTA stages the window
// once -- 16 runs of items, each item's record gathered from pos4 and shifted to its image: 8 gathers per atom instead of
// 58 -- into shared memory as {x,y} / {z,F'} / {index|shift|type} arrays, and every pair term is then served from shared
// memory. Neighbour rows for this path are 16-bit window slots (half the index stream of the 32-bit rows, which are still
// written for the Hessian / export / energy consumers).
//   k_tile_build  List::rebuild's build_neigh_list (list.cpp:134-165): candidates of an atom = the 3 x 3 rows of x-runs
//                 around its cell, tested from shared memory in the reference's operation order; same rows as k_nb_build
//   k_tile_rho    EAM pass 1 (eam.cpp:62-73)      } the pair terms and the order of every sum are those of the row
//   k_tile_force  EAM pass 2 (eam.cpp:90-107)     } kernels (eam.cuh: rho_term / force_term / VirtSum): same bits
// A tile's window may not exceed the shared-memory budget (wcap slots); a window that does raises the replica's capacity
// fault and the host regrows the budget or falls back to the row kernels.
#pragma once

#include "eam.cuh"

namespace ofly {

  constexpr int kTileThreads = 256;
  constexpr int kTileMaxWindow = 1792;                           // slots: 63 KB of shared memory, three CTAs per SM
  constexpr int kTileSlotBytes = 2 * (int)sizeof(double2) + 4;   // {x,y}, {z,w}, meta

  struct TileWindow {  // static shared memory of a tile kernel
    int gstart[16][5];    // global slot (cell_start) of window cell (wx, row) and of the row's end
    int cellbase[16][5];  // the same as window slots
    int rowlen[16];
    int central[4][2];    // window-slot ranges of the tile's own atoms (its <= 2 x 2 rows, x cells 1..cx_hi)
    int cprefix[5];       // atoms before each central range
    int nx, ny, nz, total;
  };

  struct TileSmem {
    double2 *xy, *zw;
    uint32_t *meta;
    __device__ TileSmem(unsigned char *base, int wcap) {
      xy = reinterpret_cast<double2 *>(base);
      zw = xy + wcap;
      meta = reinterpret_cast<uint32_t *>(zw + wcap);
    }
  };

  // Window geometry of tile t of replica r into W. Every thread of the CTA calls; ends with a barrier.
  __device__ __forceinline__ void tile_window(const BatchDev &B, int r, int t, TileWindow &W) {
    const int tx = t % B.tile.nt[0], ty = (t / B.tile.nt[0]) % B.tile.nt[1], tz = t / (B.tile.nt[0] * B.tile.nt[1]);
    const int gx0 = 2 * tx, gy0 = 2 * ty, gz0 = 2 * tz;  // window origin in grid cells (cell 0 is the ghost layer)
    const int nx = min(4, B.grid.shape[0] - gx0), ny = min(4, B.grid.shape[1] - gy0), nz = min(4, B.grid.shape[2] - gz0);
    const int *start = B.cell_start + (size_t)r * (B.grid.ncells + 1);
    if (threadIdx.x < 16) {
      const int rw = threadIdx.x, wy = rw & 3, wz = rw >> 2;
      int len = 0;
      if (wy < ny && wz < nz) {
        const int c0 = gx0 + (gy0 + wy) * B.grid.prod[1] + (gz0 + wz) * B.grid.prod[2];
#pragma unroll
        for (int wx = 0; wx <= 4; wx++) W.gstart[rw][wx] = (int)min((long)start[c0 + min(wx, nx)], B.img_cap);  // (overflowed sort: fault is set)
        len = W.gstart[rw][4] - W.gstart[rw][0];
      } else {
#pragma unroll
        for (int wx = 0; wx <= 4; wx++) W.gstart[rw][wx] = 0;
      }
      W.rowlen[rw] = len;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int run = 0;
      for (int rw = 0; rw < 16; rw++) {
        for (int wx = 0; wx <= 4; wx++) W.cellbase[rw][wx] = run + (W.gstart[rw][wx] - W.gstart[rw][0]);
        run += W.rowlen[rw];
      }
      W.total = run, W.nx = nx, W.ny = ny, W.nz = nz;
      // the tile's own cells: window coordinates 1..2 per axis, without the far ghost layer
      const int cx_hi = min(2, B.grid.shape[0] - 2 - gx0), cy_hi = min(2, B.grid.shape[1] - 2 - gy0), cz_hi = min(2, B.grid.shape[2] - 2 - gz0);
      int q = 0, atoms = 0;
      for (int wz = 1; wz <= 2; wz++) {
        for (int wy = 1; wy <= 2; wy++, q++) {
          const int rw = wz * 4 + wy;
          const bool in = wz <= cz_hi && wy <= cy_hi && cx_hi >= 1;
          W.central[q][0] = in ? W.cellbase[rw][1] : 0;
          W.central[q][1] = in ? W.cellbase[rw][cx_hi + 1] : 0;
          W.cprefix[q] = atoms;
          atoms += W.central[q][1] - W.central[q][0];
        }
      }
      W.cprefix[4] = atoms;
    }
    __syncthreads();
  }

  // Stage the window: warp w takes rows w and w + 8; every slot's record is gathered from pos4 and moved to its image.
  // EXACT: images by round-to-nearest add/sub as the reference's gen_image creates its ghosts (bit-exact build).
  __device__ __forceinline__ void tile_stage(const BatchDev &B, int r, const TileWindow &W, TileSmem &S) {
    const uint32_t *items = B.cell_items + (size_t)r * B.img_cap;
    const double4 *pos = B.pos4 + (size_t)r * B.N;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int rw = wid; rw < 16; rw += kTileThreads / 32) {
      const int len = W.rowlen[rw], g0 = W.gstart[rw][0], w0 = W.cellbase[rw][0];
      for (int off = lane; off < len; off += 32) {
        const uint32_t it = items[g0 + off];
        const int j = (int)(it & kIdxMask), code = (int)(it >> 24);
        double4 p = ld256(pos + j);
        if (code != kShiftCentre) {
          const int sx = code % 3, sy = (code / 3) % 3, sz = code / 9;
          p.x = sx == 1 ? p.x : (sx == 0 ? __dsub_rn(p.x, B.box.L[0]) : __dadd_rn(p.x, B.box.L[0]));
          p.y = sy == 1 ? p.y : (sy == 0 ? __dsub_rn(p.y, B.box.L[1]) : __dadd_rn(p.y, B.box.L[1]));
          p.z = sz == 1 ? p.z : (sz == 0 ? __dsub_rn(p.z, B.box.L[2]) : __dadd_rn(p.z, B.box.L[2]));
        }
        S.xy[w0 + off] = make_double2(p.x, p.y);
        S.zw[w0 + off] = make_double2(p.z, p.w);
        S.meta[w0 + off] = it | ((uint32_t)B.type[j] << 29);
      }
    }
    __syncthreads();
  }

  // Window slot of the a-th atom of the tile.
  __device__ __forceinline__ int tile_atom_slot(const TileWindow &W, int a, int &range) {
    int q = 0;
    while (q < 3 && a >= W.cprefix[q + 1]) q++;
    range = q;
    return W.central[q][0] + (a - W.cprefix[q]);
  }

  // ---- build -----------------------------------------------------------------------------------------------------
  template <int G>
  static __global__ void __launch_bounds__(kTileThreads) k_tile_build(BatchDev B, Sel Sl) {
    extern __shared__ __align__(16) unsigned char tile_smem[];
    __shared__ TileWindow W;
    TileSmem S(tile_smem, B.tile.wcap);
    const int lane = threadIdx.x % G;
    const int sub = ((threadIdx.x & 31) / G) * G;
    const unsigned gmask = group_mask<G>();
    const long items = (long)(*Sl.count) * B.tile.n_tiles;
    const double rl2 = B.grid.r_list_sq;
    for (long it = blockIdx.x; it < items; it += gridDim.x) {
      const int r = Sl.list[it / B.tile.n_tiles], t = (int)(it % B.tile.n_tiles);
      tile_window(B, r, t, W);
      if (W.total > B.tile.wcap) {  // uniform
        if (threadIdx.x == 0 && W.cprefix[4] > 0) atomicOr(B.fault + r, kFaultCapacity);
        __syncthreads();
        continue;
      }
      tile_stage(B, r, W, S);
      const int atoms = W.cprefix[4];
      for (int a = threadIdx.x / G; a < atoms; a += kTileThreads / G) {
        int q;
        const int wc = tile_atom_slot(W, a, q);
        const int ay = 1 + (q & 1), az = 1 + (q >> 1), rwc = az * 4 + ay;
        const int ax = wc >= W.cellbase[rwc][2] ? 2 : 1;
        const double2 c1 = S.xy[wc], c2 = S.zw[wc];
        const uint32_t self = S.meta[wc] & 0x1FFFFFFFu;
        const int i = (int)(self & kIdxMask);
        uint32_t *row = B.nbr + ((size_t)r * B.N + i) * B.K;
        uint16_t *row16 = B.tile.nbr16 + ((size_t)r * B.N + i) * B.K;
        const int xl = ax - 1, xh = min(ax + 1, W.nx - 1) + 1;
        int count = 0;
        for (int oz = az - 1; oz <= min(az + 1, W.nz - 1); oz++) {
          for (int oy = ay - 1; oy <= min(ay + 1, W.ny - 1); oy++) {
            const int rr = oz * 4 + oy;
            const int lo = W.cellbase[rr][xl], hi = W.cellbase[rr][xh];
            for (int a0 = lo; a0 < hi; a0 += G) {
              const int w = a0 + lane;
              bool hit = false;
              uint32_t entry = 0;
              if (w < hi) {
                const double2 q1 = S.xy[w];
                const double qz = S.zw[w].x;
                entry = S.meta[w];
                const double dx = __dsub_rn(c1.x, q1.x), dy = __dsub_rn(c1.y, q1.y), dz = __dsub_rn(c2.x, qz);
                hit = (entry & 0x1FFFFFFFu) != self && norm_sq_rn(dx, dy, dz) < rl2;
              }
              const unsigned bal = __ballot_sync(gmask, hit) >> sub;
              if (hit) {
                const int k = count + __popc(bal & ((1u << lane) - 1u));
                if (k < B.K) row[k] = entry, row16[k] = (uint16_t)w;
              }
              count += __popc(bal);
            }
          }
        }
        if (lane == 0) {
          B.cnt[(size_t)r * B.N + i] = min(count, B.K);
          if (count > B.K) atomicOr(B.fault + r, kFaultCapacity);
        }
      }
      __syncthreads();  // the window is overwritten by the next item
    }
  }

  // J consecutive 16-bit entries with one read-only load.
  template <int J>
  __device__ __forceinline__ void load_entries16(const uint16_t *p, uint32_t (&e)[J]) {
    if constexpr (J == 4) {
      uint32_t a, b;
      asm("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
      e[0] = a & 0xffffu, e[1] = a >> 16, e[2] = b & 0xffffu, e[3] = b >> 16;
    } else if constexpr (J == 2) {
      uint32_t a;
      asm("ld.global.nc.u32 %0, [%1];" : "=r"(a) : "l"(p));
      e[0] = a & 0xffffu, e[1] = a >> 16;
    } else {
      unsigned short a;
      asm("ld.global.nc.u16 %0, [%1];" : "=h"(a) : "l"(p));
      e[0] = a;
    }
  }

  __device__ __forceinline__ PairGeom tile_pair(const TileSmem &S, int w, const double4 &pi) {
    PairGeom g;
    const double2 q1 = S.xy[w], q2 = S.zw[w];
    const uint32_t m = S.meta[w];
    g.j = (int)(m & kIdxMask), g.tj = (int)(m >> 29), g.fprime_j = q2.y;
    pair_geometry(g, q1.x, q1.y, q2.x, pi);
    return g;
  }

  // ---- EAM passes ------------------------------------------------------------------------------------------------------
  // FORCE = false: pass 1, F'(rho_i) -> pos4[i].w. FORCE = true: pass 2, gradient of replica r -> out + r * stride +
  // offset[r * offset_stride] (as k_eam_force).
  template <int G, bool FORCE>
  static __global__ void __launch_bounds__(kTileThreads, 3) k_tile_eam(BatchDev B, TablesDev T, Sel Sl, double *__restrict__ out, size_t stride,
                                                                       const unsigned long long *__restrict__ offset, size_t offset_stride) {
    extern __shared__ __align__(16) unsigned char tile_smem[];
    __shared__ TileWindow W;
    TileSmem S(tile_smem, B.tile.wcap);
    constexpr int J = kVirt / G;
    const int lane = threadIdx.x % G;
    const long items = (long)(*Sl.count) * B.tile.n_tiles;
    for (long it = blockIdx.x; it < items; it += gridDim.x) {
      const int r = Sl.list[it / B.tile.n_tiles], t = (int)(it % B.tile.n_tiles);
      tile_window(B, r, t, W);
      if (W.total > B.tile.wcap) {  // uniform; the build of this list has raised the fault already
        __syncthreads();
        continue;
      }
      tile_stage(B, r, W, S);
      const int atoms = W.cprefix[4];
      for (int a = threadIdx.x / G; a < atoms; a += kTileThreads / G) {
        int q;
        const int wc = tile_atom_slot(W, a, q);
        const double2 c1 = S.xy[wc], c2 = S.zw[wc];
        const double4 pi = make_double4(c1.x, c1.y, c2.x, c2.y);
        const uint32_t mc = S.meta[wc];
        const int i = (int)(mc & kIdxMask), ti = (int)(mc >> 29);
        const int n = B.cnt[(size_t)r * B.N + i];
        const uint16_t *row = B.tile.nbr16 + ((size_t)r * B.N + i) * B.K;
        bool oob = false;
        if constexpr (!FORCE) {
          VirtSum<G, 2> acc;
          for (int kb = 0; kb < n; kb += kVirt) {
            uint32_t e[J];
            load_entries16<J>(row + kb + lane * J, e);
#pragma unroll
            for (int j = 0; j < J; j++) {
              if (kb + row_slot(lane * J + j) < n) rho_term<false>(T, ti, tile_pair(S, (int)e[j], pi), acc.a[j][0], acc.a[j][1], oob);
            }
          }
          double tot[2];
          acc.total(tot);
          if (lane == 0) reinterpret_cast<double *>(B.pos4 + (size_t)r * B.N + i)[3] = embed(T, ti, tot[0], true, oob);
        } else {
          double tot[3] = {0, 0, 0};
          if (!B.frozen[i]) {
            const double Fi = pi.w;
            VirtSum<G, 3> acc;
            for (int kb = 0; kb < n; kb += kVirt) {
              uint32_t e[J];
              load_entries16<J>(row + kb + lane * J, e);
#pragma unroll
              for (int j = 0; j < J; j++) {
                if (kb + row_slot(lane * J + j) < n) force_term(T, ti, Fi, tile_pair(S, (int)e[j], pi), acc.a[j][0], acc.a[j][1], acc.a[j][2], oob);
              }
            }
            acc.total(tot);
          }
          if (lane == 0) {
            double *o = out + (size_t)r * stride + (offset ? (size_t)offset[(size_t)r * offset_stride] : 0) + (size_t)i * 3;
            o[0] = tot[0], o[1] = tot[1], o[2] = tot[2];
          }
        }
        if (oob) atomicOr(B.fault + r, kFaultOutOfBounds);
      }
      __syncthreads();  // the window is overwritten by the next item
    }
  }

  // ---- host side ---------------------------------------------------------------------------------------------------
  constexpr int kTileBuildG = 16, kTilePairG = 4;

  static inline size_t tile_smem_bytes(const BatchDev &B) { return (size_t)B.tile.wcap * kTileSlotBytes; }

  // Dynamic shared memory above 48 KB is an opt-in per kernel -- and per translation unit: the kernels are static, so this
  // function and its flag must be too (an `inline` function's static would be shared and the second unit left out).
  static cudaError_t tile_prepare() {
    static unsigned long long done_mask = 0;  // per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    const bool done = (done_mask & bit) != 0;
    if (!done) {
      const int bytes = kTileMaxWindow * kTileSlotBytes;
      e = cudaFuncSetAttribute(k_tile_build<kTileBuildG>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tile_eam<kTilePairG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k_tile_eam<kTilePairG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e == cudaSuccess) done_mask |= bit;
    }
    return e;
  }

  static inline void tile_launch_build(const BatchDev &B, Sel sel, unsigned blocks, cudaStream_t st) {
    k_tile_build<kTileBuildG><<<blocks, kTileThreads, tile_smem_bytes(B), st>>>(B, sel);
  }
  static inline void tile_launch_gradient(const BatchDev &B, const TablesDev &T, Sel sel, unsigned blocks, cudaStream_t st, double *out, size_t stride,
                                   const unsigned long long *offset, size_t offset_stride) {
    k_tile_eam<kTilePairG, false><<<blocks, kTileThreads, tile_smem_bytes(B), st>>>(B, T, sel, nullptr, 0, nullptr, 0);
    k_tile_eam<kTilePairG, true><<<blocks, kTileThreads, tile_smem_bytes(B), st>>>(B, T, sel, out, stride, offset, offset_stride);
  }

}  // namespace ofly
