// Cell-list neighbour build / update on the device -- the B200 counterpart of neigh::List
// (include/libfly/neigh/list.hpp:71-210, src/neigh/list.cpp:37-165).
//
// The reference materialises periodic ghost atoms serially (make_ghosts, list.cpp:100-132), links
// real + ghost atoms into per-cell linked lists and then scans the <=27 adjacent cells of each real
// atom. Here the same *set* is produced in parallel:
//   1. canon_count   canonical image of every atom (orthorhombic.hpp:132-139); every valid image
//                    (a ghost exists per axis iff x < r_list -> x+L, or x > L-r_list -> x-L,
//                    orthorhombic.hpp:54-74; axes combine independently, which is what ghosts-of-
//                    ghosts produce) is binned with HyperGrid::cell_idx into a histogram.
//   2. scan          exclusive scan of the histogram per replica.
//   3. scatter       each image is written to a slot of its cell as (j | shift code << 24).
//   4. sort_cells    the image positions x_j + s*L are written in cell order, one 32 B record each (the slot
//                    order left by the atomics is arbitrary: step 6 makes the neighbour order -- and so every
//                    later sum -- deterministic).
//   5. build         a sub-warp per atom walks the 3x3 rows of adjacent cells (contiguous slots),
//                    tests (pos_i - pos_n)^2 < r_list^2 in the reference's operation order with
//                    round-to-nearest intrinsics (bit-exact neighbour sets), and compacts the hits
//                    into the atom's row with warp ballots.
//   6. order_rows    a warp per row sorts it by (neighbour index, shift) and stores it in the transposed
//                    round layout of row_slot (device_common.cuh): one gather instruction of the EAM passes
//                    then asks for consecutive neighbours of consecutive atoms, often the same record twice.
// All functions are written against one replica so that both the grid-wide kernels below and the
// persistent search kernel (search.cuh) can call them.
#pragma once

#include <cooperative_groups.h>

#include <type_traits>

#include "device_common.cuh"

namespace ofly {

  // ---- per-atom image enumeration --------------------------------------------------------------
  template <typename F>
  __device__ __forceinline__ void for_each_image(const BatchDev &B, double x, double y, double z, F &&f) {
    const double r = B.grid.r_list;
    bool okx[3], oky[3], okz[3];
    okx[0] = B.box.periodic[0] && x > __dsub_rn(B.box.L[0], r), okx[1] = true, okx[2] = B.box.periodic[0] && x < r;
    oky[0] = B.box.periodic[1] && y > __dsub_rn(B.box.L[1], r), oky[1] = true, oky[2] = B.box.periodic[1] && y < r;
    okz[0] = B.box.periodic[2] && z > __dsub_rn(B.box.L[2], r), okz[1] = true, okz[2] = B.box.periodic[2] && z < r;
#pragma unroll
    for (int sz = 0; sz < 3; sz++) {
      if (!okz[sz]) continue;
      double iz = sz == 1 ? z : (sz == 0 ? __dsub_rn(z, B.box.L[2]) : __dadd_rn(z, B.box.L[2]));
#pragma unroll
      for (int sy = 0; sy < 3; sy++) {
        if (!oky[sy]) continue;
        double iy = sy == 1 ? y : (sy == 0 ? __dsub_rn(y, B.box.L[1]) : __dadd_rn(y, B.box.L[1]));
#pragma unroll
        for (int sx = 0; sx < 3; sx++) {
          if (!okx[sx]) continue;
          double ix = sx == 1 ? x : (sx == 0 ? __dsub_rn(x, B.box.L[0]) : __dadd_rn(x, B.box.L[0]));
          f(sx + 3 * sy + 9 * sz, ix, iy, iz);
        }
      }
    }
  }

  // Step 1 for atom i of replica r. `user_pos` is the caller's [3N] array of this replica.
  __device__ __forceinline__ void nb_canon_count(const BatchDev &B, const double *__restrict__ user_pos, int r, int i) {
    double x = user_pos[3 * i], y = user_pos[3 * i + 1], z = user_pos[3 * i + 2];
    // x - L * floor(x * (1/L))  (orthorhombic.hpp:138)
    x = __dsub_rn(x, __dmul_rn(B.box.L[0], floor(__dmul_rn(x, B.box.invL[0]))));
    y = __dsub_rn(y, __dmul_rn(B.box.L[1], floor(__dmul_rn(y, B.box.invL[1]))));
    z = __dsub_rn(z, __dmul_rn(B.box.L[2], floor(__dmul_rn(z, B.box.invL[2]))));
    double4 *p = B.pos4 + (size_t)r * B.N + i;
    st256(p, make_double4(x, y, z, 0.0));
    int *hist = B.cell_count + (size_t)r * B.grid.ncells;
    B.cell_of[(size_t)r * B.N + i] = cell_index(B.grid, x, y, z);
    for_each_image(B, x, y, z, [&](int, double ix, double iy, double iz) { atomicAdd(hist + cell_index(B.grid, ix, iy, iz), 1); });
  }

  // Step 3 for atom i: the histogram is consumed as a fill cursor (counts go back to zero).
  __device__ __forceinline__ void nb_scatter(const BatchDev &B, int r, int i) {
    double4 p = ld256(B.pos4 + (size_t)r * B.N + i);
    int *hist = B.cell_count + (size_t)r * B.grid.ncells;
    const int *start = B.cell_start + (size_t)r * (B.grid.ncells + 1);
    uint32_t *items = B.cell_items + (size_t)r * B.img_cap;
    for_each_image(B, p.x, p.y, p.z, [&](int code, double ix, double iy, double iz) {
      int c = cell_index(B.grid, ix, iy, iz);
      long slot = start[c] + (atomicSub(hist + c, 1) - 1);
      if (slot < B.img_cap) items[slot] = (uint32_t)i | ((uint32_t)code << 24);
    });
  }

  constexpr int kRowSortCap = 512;  // entries one warp sorts (step 6); a longer row (no EAM list has one) keeps the build order

  // Step 4 for cell c: image positions in slot order. The slot order left by the atomics of step 3 is arbitrary; it only
  // decides the order in which step 5 finds an atom's neighbours, and step 6 sorts every row it can hold, so the cell's
  // items are sorted here (insertion sort) only when rows may be too long for step 6.
  __device__ __forceinline__ void nb_sort_cell(const BatchDev &B, int r, int c) {
    const int *start = B.cell_start + (size_t)r * (B.grid.ncells + 1);
    long s = start[c], e = start[c + 1];
    if (e > B.img_cap) e = B.img_cap;
    uint32_t *items = B.cell_items + (size_t)r * B.img_cap;
    if (B.K > kRowSortCap) {
      for (long a = s + 1; a < e; a++) {
        uint32_t v = items[a];
        long b = a - 1;
        while (b >= s && items[b] > v) {
          items[b + 1] = items[b];
          b--;
        }
        items[b + 1] = v;
      }
    }
    double4 *img = B.img_pos + (size_t)r * B.img_cap;
    const double4 *pos = B.pos4 + (size_t)r * B.N;
    for (long a = s; a < e; a++) {
      uint32_t v = items[a];
      int code = (int)(v >> 24);
      double4 p = ld256(pos + (v & kIdxMask));
      int sx = code % 3, sy = (code / 3) % 3, sz = code / 9;
      double ix = sx == 1 ? p.x : (sx == 0 ? __dsub_rn(p.x, B.box.L[0]) : __dadd_rn(p.x, B.box.L[0]));
      double iy = sy == 1 ? p.y : (sy == 0 ? __dsub_rn(p.y, B.box.L[1]) : __dadd_rn(p.y, B.box.L[1]));
      double iz = sz == 1 ? p.z : (sz == 0 ? __dsub_rn(p.z, B.box.L[2]) : __dadd_rn(p.z, B.box.L[2]));
      // w = the row entry this image becomes: (index | shift code << 24 | type << 29)
      st256(img + a, make_double4(ix, iy, iz, __longlong_as_double((long long)(v | ((uint32_t)B.type[v & kIdxMask] << 29)))));
    }
  }

  // Step 5 for atom i, executed by the G lanes of one sub-warp (all G lanes must call). The kernel is bound by instruction
  // issue (profiles/), so the candidate loop is branch-free: 32-bit slot indices, the record of a lane beyond the run's
  // end is a clamped (discarded) load, the entry with its type bits comes ready-made out of the image record.
  template <int G>
  __device__ __forceinline__ void nb_build_atom(const BatchDev &B, int r, int i, int lane) {
    const int sub = ((threadIdx.x & 31) / G) * G;  // first lane of this sub-warp inside the warp
    const unsigned gmask = group_mask<G>();
    const unsigned below = (1u << lane) - 1u;

    const double4 pi = ld256(B.pos4 + (size_t)r * B.N + i);
    const int *start = B.cell_start + (size_t)r * (B.grid.ncells + 1);
    const double4 *img = B.img_pos + (size_t)r * B.img_cap;
    uint32_t *row = B.nbr + ((size_t)r * B.N + i) * B.K;
    const uint32_t self = (uint32_t)i | ((uint32_t)kShiftCentre << 24);
    const double rl2 = B.grid.r_list_sq;
    const int K = B.K;
    const int cap = (int)(B.img_cap < 0x7fffffffL ? B.img_cap : 0x7fffffffL);

    const int c = B.cell_of[(size_t)r * B.N + i];
    const int cx = c % B.grid.shape[0], cy = (c / B.grid.shape[0]) % B.grid.shape[1], cz = c / (B.grid.shape[0] * B.grid.shape[1]);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, B.grid.shape[0] - 1);

    int count = 0;
    for (int oz = max(cz - 1, 0); oz <= min(cz + 1, B.grid.shape[2] - 1); oz++) {
      for (int oy = max(cy - 1, 0); oy <= min(cy + 1, B.grid.shape[1] - 1); oy++) {
        const int base = oy * B.grid.prod[1] + oz * B.grid.prod[2];
        const int s = start[base + x0], e = min(start[base + x1 + 1], cap);
        for (int a0 = s; a0 < e; a0 += G) {
          const int a = a0 + lane;
          const bool in = a < e;
          const double4 q = ld256(img + (in ? a : e - 1));
          const uint32_t entry = (uint32_t)__double_as_longlong(q.w);
          const double dx = __dsub_rn(pi.x, q.x), dy = __dsub_rn(pi.y, q.y), dz = __dsub_rn(pi.z, q.z);
          const bool hit = in && (entry & 0x1FFFFFFFu) != self && norm_sq_rn(dx, dy, dz) < rl2;
          const unsigned b = __ballot_sync(gmask, hit) >> sub;  // sub-warps loop independently: ballot over own lanes only
          const int k = count + __popc(b & below);
          if (hit && k < K) row[k] = entry;
          count += __popc(b);
        }
      }
    }
    if (lane == 0) {
      B.cnt[(size_t)r * B.N + i] = min(count, K);
      if (count > K) atomicOr(B.fault + r, kFaultCapacity);
    }
  }

  // update(delta): r <- r - delta (list.cpp:37-51). Images follow implicitly (pos4[j] + s*L).
  __device__ __forceinline__ void nb_update_atom(const BatchDev &B, const double *__restrict__ delta, int r, int i) {
    double4 *p = B.pos4 + (size_t)r * B.N + i;
    double4 v = ld256(p);
    v.x -= delta[3 * i], v.y -= delta[3 * i + 1], v.z -= delta[3 * i + 2];
    st256(p, v);
  }

  // ---- grid kernels ------------------------------------------------------------------------------
  // Every kernel works on a *selection* of replicas read from device memory (list + count), with a
  // grid-stride loop over (selected replica, item). The stand-alone List::rebuild call selects all
  // replicas; the lock-step search (search.cuh) selects the replicas whose skin was used up this
  // tick -- the launch geometry is fixed (graph-capturable) and a launch with an empty selection
  // returns at once.
  struct Sel {
    const int *list;   // replica ids
    const int *count;  // number of entries
  };

  static __global__ void k_nb_zero(BatchDev B) {
    size_t n = (size_t)B.R * B.grid.ncells;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) B.cell_count[t] = 0;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < (size_t)B.R; t += (size_t)gridDim.x * blockDim.x) B.fault[t] = 0;
  }

  // pos: base of the callers' position arrays; replica r's [3N] array starts at pos + r * stride.
  static __global__ void __launch_bounds__(256) k_nb_canon_count(BatchDev B, Sel S, const double *__restrict__ pos, size_t stride) {
    const size_t total = (size_t)(*S.count) * B.N;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
      int r = S.list[t / B.N], i = (int)(t % B.N);
      nb_canon_count(B, pos + (size_t)r * stride, r, i);
    }
  }

  // Step 2 by one CTA: exclusive scan of the replica's histogram into cell_start[0..ncells].
  // `sm` is 33 ints of shared memory; every thread of the CTA must call.
  __device__ __forceinline__ void nb_scan_block(const BatchDev &B, int r, int *sm) {
    const int *hist = B.cell_count + (size_t)r * B.grid.ncells;
    int *start = B.cell_start + (size_t)r * (B.grid.ncells + 1);
    int n = B.grid.ncells;
    int per = (n + blockDim.x - 1) / blockDim.x;
    int lo = min((int)threadIdx.x * per, n), hi = min(lo + per, n);
    // The histogram is produced by atomics (performed in L2); read it there: a CTA that runs several rebuilds in one
    // kernel (the resident engine, search.cuh) may still hold an older copy of these lines in its L1.
    int sum = 0;
    for (int c = lo; c < hi; c++) sum += __ldcg(hist + c);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
      int v = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0;
      int vi = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, vi, o);
        if (lane >= o) vi += u;
      }
      sm[lane] = vi - v;
      if (lane == 31) sm[32] = vi;
    }
    __syncthreads();
    int run = sm[w] + inc - sum;
    for (int c = lo; c < hi; c++) {
      start[c] = run;
      run += __ldcg(hist + c);
    }
    if (threadIdx.x == 0) {
      start[n] = sm[32];
      if (sm[32] > B.img_cap) atomicOr(B.fault + r, kFaultCapacity);
    }
    __syncthreads();
  }

  static __global__ void __launch_bounds__(1024) k_nb_scan(BatchDev B, Sel S) {
    __shared__ int sm[33];
    const int n = *S.count;
    for (int a = blockIdx.x; a < n; a += gridDim.x) nb_scan_block(B, S.list[a], sm);
  }

  static __global__ void __launch_bounds__(256) k_nb_scatter(BatchDev B, Sel S) {
    const size_t total = (size_t)(*S.count) * B.N;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
      nb_scatter(B, S.list[t / B.N], (int)(t % B.N));
    }
  }

  static __global__ void __launch_bounds__(128) k_nb_sort_cells(BatchDev B, Sel S) {
    const size_t total = (size_t)(*S.count) * B.grid.ncells;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
      nb_sort_cell(B, S.list[t / B.grid.ncells], (int)(t % B.grid.ncells));
    }
  }

  // One work item = one tile of 256/G consecutive atoms of one selected replica.
  template <int G>
  __device__ __forceinline__ void nb_build_tiles(const BatchDev &B, const Sel &S) {
    constexpr int kTile = 256 / G;
    const int tiles = (B.N + kTile - 1) / kTile;
    const long items = (long)(*S.count) * tiles;
    const int lane = threadIdx.x % G;
    for (long w = blockIdx.x; w < items; w += gridDim.x) {
      int r = S.list[w / tiles];
      int i = (int)(w % tiles) * kTile + threadIdx.x / G;
      if (i < B.N) nb_build_atom<G>(B, r, i, lane);  // whole sub-warps are in or out together
    }
  }
  template <int G>
  static __global__ void __launch_bounds__(256) k_nb_build(BatchDev B, Sel S) {
    nb_build_tiles<G>(B, S);
  }
  // Sub-warp width of the build (the stored row is a compaction in candidate order, so the list does not depend on it).
  // The kernel is bound by instruction issue, and wide groups walk the candidate runs in fewer, less divergent steps:
  // measured on 64 x 16 000 atoms, 2.01 / 1.69 / 1.67 / 2.27 ms per rebuild for G = 4 / 8 / 16 / 32; 32 wins while the
  // atoms x 32 lanes still fit a few waves of resident threads.
  __host__ __device__ inline int build_group_for(long atoms, long slots) { return atoms * 32 <= 4 * slots ? 32 : 16; }

  // Inside a search the width is picked on the device from the number of replicas that asked for a rebuild this tick
  // (usually a few): 32 / 16 / 8 lanes while atoms x lanes fit `budget` threads, 16 beyond that. Measured on C2 against
  // 32-or-16 only: 409 against 414 ms per step.
  static __global__ void __launch_bounds__(256) k_nb_build_dyn(BatchDev B, Sel S, long budget) {
    const long atoms = (long)(*S.count) * B.N;
    if (atoms * 32 <= budget) {
      nb_build_tiles<32>(B, S);
    } else if (atoms * 16 <= budget) {
      nb_build_tiles<16>(B, S);
    } else if (atoms * 8 <= budget) {
      nb_build_tiles<8>(B, S);
    } else {
      nb_build_tiles<16>(B, S);
    }
  }

  // Step 6: every row of the selected replicas ordered by (neighbour index, shift code) and stored in the transposed
  // round layout (row_slot). One warp per row: bitonic sort of the row, padded to a power of two, in shared memory.
  // WITH16: the tile path's rows of 16-bit window slots (tile.cuh) are carried along as values.
  // Rows of up to 32 E entries in registers: element p = 32 e + lane, bitonic network with shuffles below 32 and
  // register pairs above. Returns with v[e] = the key of rank 32 e + lane.
  template <int E>
  __device__ __forceinline__ void warp_bitonic(uint32_t (&v)[E], int lane) {
#pragma unroll
    for (int k2 = 2; k2 <= 32 * E; k2 <<= 1) {
#pragma unroll
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        if (j >= 32) {  // partner in the same thread; the direction depends on e alone
#pragma unroll
          for (int e = 0; e < E; e++) {
            const int f = e | (j >> 5);
            if (f != e) {
              const bool up = ((32 * e) & k2) == 0;
              const uint32_t lo = min(v[e], v[f]), hi = max(v[e], v[f]);
              v[e] = up ? lo : hi, v[f] = up ? hi : lo;
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < E; e++) {
            const uint32_t o = __shfl_xor_sync(0xffffffffu, v[e], j);
            const bool up = (((32 * e) | lane) & k2) == 0, lower = (lane & j) == 0;
            v[e] = (up == lower) ? min(v[e], o) : max(v[e], o);
          }
        }
      }
    }
  }
  template <int E>
  __device__ __forceinline__ void order_row_regs(uint32_t *row, int n, int padded, int lane) {
    uint32_t v[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
      const int p = 32 * e + lane;
      v[e] = 0xffffffffu;
      if (p < n) {
        const uint32_t x = row[p];
        v[e] = (x << 8) | (x >> 24);  // index in the high 24 bits, then type and shift code
      }
    }
    warp_bitonic<E>(v, lane);
#pragma unroll
    for (int e = 0; e < E; e++) {
      const int k = 32 * e + lane;
      if (k < padded) row[row_slot(k)] = (v[e] >> 8) | (v[e] << 24);
    }
  }

  template <bool WITH16>
  using RowSortKey = typename std::conditional<WITH16, unsigned long long, uint32_t>::type;

  // buf: kRowSortCap keys per warp of the (256-thread) CTA
  template <bool WITH16>
  __device__ __forceinline__ void nb_order_rows(const BatchDev &B, const Sel &S, RowSortKey<WITH16> *buf) {
    using Key = RowSortKey<WITH16>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Key *key = buf + (size_t)warp * kRowSortCap;
    const long rows = (long)(*S.count) * B.N;
    for (long w = blockIdx.x * 8L + warp; w < rows; w += gridDim.x * 8L) {
      const int r = S.list[w / B.N], i = (int)(w % B.N);
      const size_t at = (size_t)r * B.N + i;
      uint32_t *row = B.nbr + at * B.K;
      uint16_t *row16 = WITH16 ? B.tile.nbr16 + at * B.K : nullptr;
      const int n = B.cnt[at];
      const int padded = (n + 15) & ~15;  // K is a multiple of 16
      if constexpr (!WITH16) {
        if (n <= 128) {  // warp-uniform: the usual case, sorted in registers
          if (n <= 32) {
            order_row_regs<1>(row, n, padded, lane);
          } else if (n <= 64) {
            order_row_regs<2>(row, n, padded, lane);
          } else {
            order_row_regs<4>(row, n, padded, lane);
          }
          continue;
        }
      }
      if (n > kRowSortCap) {              // layout only, two rounds at a time
        for (int base = 0; base < padded; base += 32) {
          const int k = base + lane;
          uint32_t e = 0;
          uint16_t e16 = 0;
          if (k < padded) {
            e = row[k];
            if (WITH16) e16 = row16[k];
          }
          __syncwarp();
          if (k < padded) {
            row[row_slot(k)] = e;
            if (WITH16) row16[row_slot(k)] = e16;
          }
        }
        continue;
      }
      int P = 16;
      while (P < n) P <<= 1;
      for (int p = lane; p < P; p += 32) {
        Key v = ~Key(0);
        if (p < n) {
          const uint32_t e = row[p];
          const uint32_t rot = (e << 8) | (e >> 24);  // index in the high 24 bits, then type and shift code
          if constexpr (WITH16) {
            v = ((Key)rot << 16) | (Key)row16[p];
          } else {
            v = rot;
          }
        }
        key[p] = v;
      }
      __syncwarp();
      for (int k2 = 2; k2 <= P; k2 <<= 1) {
        for (int j = k2 >> 1; j > 0; j >>= 1) {
          for (int t = lane; t < P / 2; t += 32) {
            const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1)), b = a | j;
            const Key x = key[a], y = key[b];
            const bool up = (a & k2) == 0;
            if ((x > y) == up) key[a] = y, key[b] = x;
          }
          __syncwarp();
        }
      }
      for (int k = lane; k < padded; k += 32) {
        const Key v = key[k];
        uint32_t rot;
        if constexpr (WITH16) {
          rot = (uint32_t)(v >> 16);
          row16[row_slot(k)] = (uint16_t)v;
        } else {
          rot = v;
        }
        row[row_slot(k)] = (rot >> 8) | (rot << 24);
      }
      __syncwarp();  // the buffer is refilled by the next row
    }
  }
  template <bool WITH16>
  static __global__ void __launch_bounds__(256) k_nb_order_rows(BatchDev B, Sel S) {
    __shared__ RowSortKey<WITH16> buf[8 * kRowSortCap];
    nb_order_rows<WITH16>(B, S, buf);
  }
  // blocks: a grid-stride launch over rows_of_work rows, eight rows per CTA
  static inline void launch_order_rows(const BatchDev &B, Sel sel, unsigned blocks, cudaStream_t st) {
    if (B.tile.n_tiles > 0) {
      k_nb_order_rows<true><<<blocks, 256, 0, st>>>(B, sel);
    } else {
      k_nb_order_rows<false><<<blocks, 256, 0, st>>>(B, sel);
    }
  }

  // Steps 1-6 for the selected replicas in ONE cooperative launch, grid-wide barriers between the steps. Inside a search
  // most ticks have nobody to rebuild: this kernel then costs one empty launch, where six separate kernels behind an IF
  // node cost the evaluation of the conditional node (~10 us of a ~70 us tick late in a batch, profiles/). Same rows as
  // the separate kernels. The grid must be co-resident (cooperative launch); every loop is a grid-stride loop.
  static __global__ void __launch_bounds__(256) k_nb_rebuild_fused(BatchDev B, Sel S, const double *__restrict__ pos, size_t stride, long budget) {
    const int nsel = *S.count;
    if (nsel == 0) return;  // the whole grid leaves
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ uint32_t buf[8 * kRowSortCap];
    const size_t gtid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, gthreads = (size_t)gridDim.x * blockDim.x;
    const size_t atoms = (size_t)nsel * B.N;
    for (size_t t = gtid; t < atoms; t += gthreads) {
      const int r = S.list[t / B.N];
      nb_canon_count(B, pos + (size_t)r * stride, r, (int)(t % B.N));
    }
    grid.sync();
    for (int a = blockIdx.x; a < nsel; a += gridDim.x) nb_scan_block(B, S.list[a], reinterpret_cast<int *>(buf));
    grid.sync();
    for (size_t t = gtid; t < atoms; t += gthreads) nb_scatter(B, S.list[t / B.N], (int)(t % B.N));
    grid.sync();
    const size_t cells = (size_t)nsel * B.grid.ncells;
    for (size_t t = gtid; t < cells; t += gthreads) nb_sort_cell(B, S.list[t / B.grid.ncells], (int)(t % B.grid.ncells));
    grid.sync();
    if ((long)atoms * 32 <= budget) {  // widths as k_nb_build_dyn
      nb_build_tiles<32>(B, S);
    } else if ((long)atoms * 16 <= budget) {
      nb_build_tiles<16>(B, S);
    } else if ((long)atoms * 8 <= budget) {
      nb_build_tiles<8>(B, S);
    } else {
      nb_build_tiles<16>(B, S);
    }
    grid.sync();
    nb_order_rows<false>(B, S, buf);
  }

  static __global__ void __launch_bounds__(256) k_nb_update(BatchDev B, const double *__restrict__ delta) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)B.R * B.N) return;
    int r = (int)(t / B.N), i = (int)(t % B.N);
    nb_update_atom(B, delta + (size_t)r * 3 * B.N, r, i);
  }

}  // namespace ofly
