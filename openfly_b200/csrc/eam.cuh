// EAM energy / gradient over a device-resident neighbour list -- the B200 counterpart of
// potential::EAM (src/potential/EAM/eam.cpp:33-120) iterating with neigh::List::for_neighbours
// (include/libfly/neigh/list.hpp:138-154) and evaluating Spline::f/fp (spline.hpp:81-154).
//
// Two-pass, scatter-free gather. A sub-warp of G lanes owns one atom; each lane takes 16 / G consecutive
// storage slots of every 16-slot round of the atom's neighbour row (one vector load), fetches each neighbour's
// {x,y,z,F'} record with one 256-bit load, the spline knot {a,b,c,d} with one 256-bit read-only load, and the
// partial sums are combined in a fixed, width-independent order (VirtSum). No atomics, no stores to other atoms.
// Rows are ordered by neighbour index and their rounds stored transposed (row_slot, device_common.cuh): the records one
// warp-wide gather asks for are then 8 consecutive atoms x 4 consecutive neighbours, which is what the L1 data pipe --
// the bound of these kernels -- serves cheapest.
// These "row kernels" gather from global memory and work for every box; the tile kernels (tile.cuh, opt-in: measured
// slower) stage cell windows in shared memory for fully periodic boxes and produce the same bits.
//   pass 1: rho_b = sum_a phi_{a->b}(r);  F'_b = F_b'(rho_b) is stored in pos4[b].w   (eam.cpp:62-73)
//   pass 2: grad_g = - sum_a [V'(r) + F'_g phi'_{a->g}(r) + F'_a phi'_{g->a}(r)] / r * dr (eam.cpp:90-107)
// The FP64 pipe is the co-limiter of this path, so r and 1/r come from one rsqrt seed + one
// correction (fast_rsqrt) instead of an IEEE sqrt and an IEEE division.
#pragma once

#include "neigh.cuh"

namespace ofly {

  // One neighbour as the pair terms see it: image position minus centre, F'(rho) of the neighbour, its type.
  struct PairGeom {
    double dx, dy, dz;  // r_n - r_i, neighbour image minus centre (list.hpp:148)
    double r2;
    double fprime_j;  // pos4[j].w
    int j, tj;
  };

  // Geometry of a pair from the neighbour's image position (already shifted) and the centre.
  // r2 in gnorm_sq's own order, (x x + y y) + z z without contraction (core.hpp:655-680): the second cut-off filter of
  // for_neighbours (list.hpp:148-150) then admits exactly the reference's pairs, not a set that may differ in the last bit.
  __device__ __forceinline__ void pair_geometry(PairGeom &g, double qx, double qy, double qz, const double4 &pi) {
    g.dx = qx - pi.x, g.dy = qy - pi.y, g.dz = qz - pi.z;
    g.r2 = norm_sq_rn(g.dx, g.dy, g.dz);
  }

  // Decode one neighbour entry and fetch its record from global memory. The image position is pos4[j] + shift, the
  // arithmetic that created the reference's ghost atom; (x + 0.0) for the home image is x itself.
  __device__ __forceinline__ PairGeom load_pair(const BatchDev &B, const double4 *__restrict__ pos, const double4 &pi, uint32_t e) {
    PairGeom g;
    g.j = (int)(e & kIdxMask);
    g.tj = (int)(e >> 29);
    int code = (int)((e >> 24) & 31u);
    double4 pj = ld256(pos + g.j);
    double sx = 0, sy = 0, sz = 0;
    if (code != kShiftCentre) {
      int dxg = code % 3, dyg = (code / 3) % 3, dzg = code / 9;
      sx = shift_of(dxg, B.box.L[0]), sy = shift_of(dyg, B.box.L[1]), sz = shift_of(dzg, B.box.L[2]);
    }
    pair_geometry(g, pj.x + sx, pj.y + sy, pj.z + sz, pi);
    g.fprime_j = pj.w;
    return g;
  }

  // Spline::fetch (spline.hpp:135-154): interval index and offset. Flags OutOfBounds, clamps the read.
  // The reference converts x * inv_dx with static_cast<size_t>: -1 < s < 0 truncates to interval 0 (no throw), s <= -1
  // and NaN become a huge index and throw, s >= n_knots throws. The range test is done on the double so that neither a
  // negative nor a NaN argument can index outside the table.
  __device__ __forceinline__ int knot_index(double x, double inv_dx, double dx, int n_knots, double &x0, bool &oob) {
    const double s = x * inv_dx;
    int i = 0;
    if (s > -1.0 && s < (double)n_knots) {
      i = __double2int_rz(s);
    } else {
      oob = true;
      i = s > 0.0 ? n_knots - 1 : 0;  // NaN lands on 0
    }
    x0 = fma(-(double)i, dx, x);
    return i;
  }
  __device__ __forceinline__ double spline_f(const double4 &k, double x0) { return fma(x0, fma(x0, fma(x0, k.w, k.z), k.y), k.x); }
  __device__ __forceinline__ double spline_fp(const double4 &k, double x0) { return fma(x0, fma(x0 * 3.0, k.w, 2.0 * k.z), k.y); }
  __device__ __forceinline__ double spline_fpp(const double4 &k, double x0) { return fma(6.0 * x0, k.w, 2.0 * k.z); }

  // ---- the pair terms, shared by the row kernels below and the tile kernels (tile.cuh): same bits on both paths ------
  // pass 1 (eam.cpp:62-73, and :41-46 for the energy): rho += phi_{tj -> ti}(r) [, v += V(r)]
  template <bool WITH_V>
  __device__ __forceinline__ void rho_term(const TablesDev &T, int ti, const PairGeom &g, double &rho, double &v, bool &oob) {
    if (g.r2 < T.r_cut_sq) {
      double rr = g.r2 * fast_rsqrt(g.r2);
      double x0;
      int idx = knot_index(rr, T.inv_dR, T.dR, T.kR, x0, oob);
      rho += spline_f(ld256_nc(T.phi + (size_t)(g.tj + T.n_types * ti) * T.kR + idx), x0);
      if (WITH_V) {
        int hi = max(ti, g.tj), lo = min(ti, g.tj);
        v += spline_f(ld256_nc(T.V + (size_t)(hi + T.n_types * lo) * T.kR + idx), x0);
      }
    }
  }
  // pass 2 (eam.cpp:90-107): grad -= [V'(r) + F'_i phi'_{j->i}(r) + F'_j phi'_{i->j}(r)] / r * dr
  __device__ __forceinline__ void force_term(const TablesDev &T, int ti, double Fi, const PairGeom &g, double &ax, double &ay, double &az, bool &oob) {
    if (g.r2 < T.r_cut_sq) {
      double rinv = fast_rsqrt(g.r2);
      double rr = g.r2 * rinv;
      double x0;
      int idx = knot_index(rr, T.inv_dR, T.dR, T.kR, x0, oob);
      int hi = max(ti, g.tj), lo = min(ti, g.tj);
      double mag = spline_fp(ld256_nc(T.V + (size_t)(hi + T.n_types * lo) * T.kR + idx), x0);
      double p_ji = spline_fp(ld256_nc(T.phi + (size_t)(g.tj + T.n_types * ti) * T.kR + idx), x0);  // phi(a -> g)
      double p_ij = p_ji;
      if (g.tj != ti) p_ij = spline_fp(ld256_nc(T.phi + (size_t)(ti + T.n_types * g.tj) * T.kR + idx), x0);  // phi(g -> a)
      mag = fma(Fi, p_ji, mag);
      mag = fma(g.fprime_j, p_ij, mag);
      double s = mag * rinv;
      ax = fma(-s, g.dx, ax), ay = fma(-s, g.dy, ay), az = fma(-s, g.dz, az);
    }
  }

  // ---- per-atom sums whose bits do not depend on the sub-warp width ---------------------------------
  // A row is consumed in rounds of kVirt = 16 storage slots; slot v of a round belongs to *virtual lane* v (the entry it
  // holds is the round's neighbour of rank row_slot(v), device_common.cuh), a virtual lane adds its terms round after
  // round, and the 16 partial sums are combined by the butterfly (v, v^1), (v, v^2), (v, v^4), (v, v^8). A physical
  // sub-warp of G lanes carries J = 16 / G consecutive virtual lanes per thread (lane l owns v = l J .. l J + J - 1: its
  // entries of a round are contiguous, ONE vector load), folds the low butterfly levels in registers and does the
  // remaining ones with shuffles. Every G therefore produces the SAME bits, and so does the tile
  // path (tile.cuh), so a replica's energies, forces and trajectory do not depend on how many replicas share its batch
  // (the tick kernels pick G from the active count), on the code path, or on how a batch is sharded over GPUs.
  constexpr int kVirt = 16;

  template <int G, int NS>
  struct VirtSum {
    static_assert(G == 4 || G == 8 || G == 16, "sub-warp width must divide the virtual width");
    static constexpr int J = kVirt / G;
    double a[J][NS];
    __device__ __forceinline__ VirtSum() {
#pragma unroll
      for (int j = 0; j < J; j++) {
#pragma unroll
        for (int q = 0; q < NS; q++) a[j][q] = 0.0;
      }
    }
    // Totals on every lane of the sub-warp.
    __device__ __forceinline__ void total(double (&out)[NS]) {
#pragma unroll
      for (int h = 1; h < J; h <<= 1) {  // butterfly levels 1, 2 (those that stay inside a thread)
#pragma unroll
        for (int j = 0; j < J; j += 2 * h) {
#pragma unroll
          for (int q = 0; q < NS; q++) a[j][q] = a[j][q] + a[j + h][q];
        }
      }
      const unsigned m = group_mask<G>();
#pragma unroll
      for (int q = 0; q < NS; q++) {
        double v = a[0][q];
#pragma unroll
        for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(m, v, o);  // levels J, 2 J, ... 8
        out[q] = v;
      }
    }
  };

  // J consecutive row entries with one read-only vector load (rows are 64 B aligned, k is a multiple of J).
  template <int J>
  __device__ __forceinline__ void load_entries(const uint32_t *p, uint32_t (&e)[J]) {
    if constexpr (J == 4) {
      asm("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]) : "l"(p));
    } else if constexpr (J == 2) {
      asm("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(e[0]), "=r"(e[1]) : "l"(p));
    } else {
      asm("ld.global.nc.u32 %0, [%1];" : "=r"(e[0]) : "l"(p));
    }
  }

  // Pass 1 for atom i by a sub-warp of G lanes: returns rho_i on every lane. If WITH_V, also accumulates
  // sum_a V(r) (energy, eam.cpp:41-46) into *v_sum.
  template <int G, bool WITH_V>
  __device__ __forceinline__ double eam_rho_atom(const BatchDev &B, const TablesDev &T, int r, int i, int lane, double *v_sum, bool &oob) {
    constexpr int J = kVirt / G;
    const double4 *pos = B.pos4 + (size_t)r * B.N;
    const uint32_t *row = B.nbr + ((size_t)r * B.N + i) * B.K;
    const int n = B.cnt[(size_t)r * B.N + i];
    const int ti = B.type[i];
    const double4 pi = ld256(pos + i);
    VirtSum<G, 2> acc;
    // the entries of the next round are requested before this round's pairs are worked on: the entry load heads the
    // dependent chain entry -> record -> knot and is the first stall of the kernel (long_scoreboard, profiles/)
    uint32_t e[J], en[J];
    const int p0 = lane * J;  // this lane's storage slots inside every round
    if (n > 0) load_entries<J>(row + p0, e);
    for (int kb = 0; kb < n; kb += kVirt) {
      if (kb + kVirt < n) load_entries<J>(row + kb + kVirt + p0, en);
#pragma unroll
      for (int j = 0; j < J; j++) {
        if (kb + row_slot(p0 + j) < n) {  // slot -> rank: pair j of the sub-warp's lanes are consecutive neighbours
          PairGeom g = load_pair(B, pos, pi, e[j]);
          rho_term<WITH_V>(T, ti, g, acc.a[j][0], acc.a[j][1], oob);
        }
      }
#pragma unroll
      for (int j = 0; j < J; j++) e[j] = en[j];
    }
    double tot[2];
    acc.total(tot);
    if (WITH_V) *v_sum = tot[1];
    return tot[0];
  }

  // Pass 2 for atom i: gradient on every lane.
  template <int G>
  __device__ __forceinline__ void eam_force_atom(const BatchDev &B, const TablesDev &T, int r, int i, int lane, double &gx, double &gy,
                                                 double &gz, bool &oob) {
    constexpr int J = kVirt / G;
    const double4 *pos = B.pos4 + (size_t)r * B.N;
    const uint32_t *row = B.nbr + ((size_t)r * B.N + i) * B.K;
    const int n = B.cnt[(size_t)r * B.N + i];
    const int ti = B.type[i];
    const double4 pi = ld256(pos + i);
    const double Fi = pi.w;
    VirtSum<G, 3> acc;
    uint32_t e[J], en[J];  // entries one round ahead (see eam_rho_atom)
    const int p0 = lane * J;
    if (n > 0) load_entries<J>(row + p0, e);
    for (int kb = 0; kb < n; kb += kVirt) {
      if (kb + kVirt < n) load_entries<J>(row + kb + kVirt + p0, en);
#pragma unroll
      for (int j = 0; j < J; j++) {
        if (kb + row_slot(p0 + j) < n) {
          PairGeom g = load_pair(B, pos, pi, e[j]);
          force_term(T, ti, Fi, g, acc.a[j][0], acc.a[j][1], acc.a[j][2], oob);
        }
      }
#pragma unroll
      for (int j = 0; j < J; j++) e[j] = en[j];
    }
    double tot[3];
    acc.total(tot);
    gx = tot[0], gy = tot[1], gz = tot[2];
  }

  // F_i'(rho) / F_i(rho) with the OutOfBounds check of Spline::fetch.
  __device__ __forceinline__ double embed(const TablesDev &T, int ti, double rho, bool deriv, bool &oob) {
    double x0;
    int idx = knot_index(rho, T.inv_dF, T.dF, T.kF, x0, oob);
    double4 k = ld256_nc(T.F + (size_t)ti * T.kF + idx);
    return deriv ? spline_fp(k, x0) : spline_f(k, x0);
  }

  // ---- grid kernels ------------------------------------------------------------------------------
  // Same selection scheme as the neighbour kernels (neigh.cuh, struct Sel): one work item is a tile
  // of 256/G consecutive atoms of one selected replica; blocks stride over the items, so the launch
  // geometry does not depend on how many replicas are still searching.

  // Pass 1: rho and F'(rho) -> pos4[].w
  template <int G>
  __device__ __forceinline__ void eam_rho_tiles(const BatchDev &B, const TablesDev &T, const Sel &S) {
    constexpr int kTile = 256 / G;
    const int tiles = (B.N + kTile - 1) / kTile;
    const long items = (long)(*S.count) * tiles;
    const int lane = threadIdx.x % G;
    for (long w = blockIdx.x; w < items; w += gridDim.x) {
      const int r = S.list[w / tiles];
      const int i = (int)(w % tiles) * kTile + threadIdx.x / G;
      if (i >= B.N) continue;  // whole sub-warps leave together
      bool oob = false;
      double rho = eam_rho_atom<G, false>(B, T, r, i, lane, nullptr, oob);
      if (lane == 0) {
        double fp = embed(T, B.type[i], rho, true, oob);
        reinterpret_cast<double *>(B.pos4 + (size_t)r * B.N + i)[3] = fp;
      }
      if (oob) atomicOr(B.fault + r, kFaultOutOfBounds);
    }
  }
  // Pass 1 is capped for five CTAs per SM (48 registers, 32 B spilled outside the pair loop): measured 0.709 ms per launch
  // pair of the 64 x 16 000 batch against 0.719 uncapped (60 registers, four CTAs) and 0.766 capped for six (40 registers).
#ifndef OFLY_RHO_MINB
#define OFLY_RHO_MINB 5
#endif
  template <int G>
  static __global__ void __launch_bounds__(256, OFLY_RHO_MINB) k_eam_rho(BatchDev B, TablesDev T, Sel S) {
    eam_rho_tiles<G>(B, T, S);
  }

  // Pass 2: gradient of replica r -> out + r * stride + offset[r * offset_stride] (offset may be null).
  // The search keeps, per replica, the slot its next gradient belongs in (g0, g1, g1' of the rotor).
  template <int G>
  __device__ __forceinline__ void eam_force_tiles(const BatchDev &B, const TablesDev &T, const Sel &S, double *__restrict__ out, size_t stride,
                                                  const unsigned long long *__restrict__ offset, size_t offset_stride) {
    constexpr int kTile = 256 / G;
    const int tiles = (B.N + kTile - 1) / kTile;
    const long items = (long)(*S.count) * tiles;
    const int lane = threadIdx.x % G;
    for (long w = blockIdx.x; w < items; w += gridDim.x) {
      const int r = S.list[w / tiles];
      const int i = (int)(w % tiles) * kTile + threadIdx.x / G;
      if (i >= B.N) continue;
      bool oob = false;
      double gx = 0, gy = 0, gz = 0;
      if (!B.frozen[i]) eam_force_atom<G>(B, T, r, i, lane, gx, gy, gz, oob);
      if (lane == 0) {
        double *o = out + (size_t)r * stride + (offset ? (size_t)offset[(size_t)r * offset_stride] : 0) + (size_t)i * 3;
        o[0] = gx, o[1] = gy, o[2] = gz;
      }
      if (oob) atomicOr(B.fault + r, kFaultOutOfBounds);
    }
  }
  // MINB = CTAs per SM the register allocation is capped for. 3 (80 registers, nothing spilled) is the default: capping at
  // 64 registers for a fourth CTA spills ~100 bytes inside the pair loop (local-memory traffic on the L1 pipe that
  // already bounds the kernel): measured 0.857 ms against 0.819 per launch pair, 620 against 569 ms per 64-search step.
  template <int G, int MINB = 3>
  static __global__ void __launch_bounds__(256, MINB) k_eam_force(BatchDev B, TablesDev T, Sel S, double *__restrict__ out, size_t stride,
                                                            const unsigned long long *__restrict__ offset, size_t offset_stride) {
    eam_force_tiles<G>(B, T, S, out, stride, offset, offset_stride);
  }

  // Sub-warp width for `atoms` atoms of work on a chip with `slots` resident threads: few atoms -> wide groups
  // (more threads, shorter per-lane chains: latency), many atoms -> narrow groups (fewer idle lanes on the ragged
  // row tail: throughput). Measured on a lone 16 000-atom replica: the widest group that still fits ONE wave of
  // resident threads (the pair kernels hold ~1024 threads per SM, half of `slots`) is the fastest.
  // The width only decides how the work is spread: the sums are G-invariant (VirtSum), widths are 4, 8 or 16.
  __host__ __device__ inline int group_for(long atoms, long slots) {
    const long wave = slots / 2;
    if (atoms * 16 <= wave) return 16;
    if (atoms * 8 <= wave) return 8;
    return 4;
  }

  // The search's tick kernels pick G on the device from the number of replicas still active: for most ticks of
  // a batch only a few replicas are left and the evaluation is a latency chain, not a throughput problem.
  // trace (tuning aid, may be null): [0] = count, then (mark, globaltimer) pairs -- block 0 stamps its start and end
  __device__ __forceinline__ void kernel_stamp(unsigned long long *trace, int cap, int mark) {
    if (trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      unsigned long long slot = atomicAdd(trace, 1ull);
      if (slot < (unsigned long long)cap) trace[1 + 2 * slot] = (unsigned long long)mark, trace[2 + 2 * slot] = t;
    }
  }

  static __global__ void __launch_bounds__(256, OFLY_RHO_MINB) k_eam_rho_dyn(BatchDev B, TablesDev T, Sel S, long slots, unsigned long long *trace, int trace_cap) {
    kernel_stamp(trace, trace_cap, 91);
    switch (group_for((long)(*S.count) * B.N, slots)) {
      case 4: eam_rho_tiles<4>(B, T, S); break;
      case 8: eam_rho_tiles<8>(B, T, S); break;
      default: eam_rho_tiles<16>(B, T, S); break;
    }
    kernel_stamp(trace, trace_cap, 92);
  }
#ifndef OFLY_DYN_MINB
#define OFLY_DYN_MINB 3
#endif
  static __global__ void __launch_bounds__(256, OFLY_DYN_MINB) k_eam_force_dyn(BatchDev B, TablesDev T, Sel S, double *__restrict__ out, size_t stride,
                                                                const unsigned long long *__restrict__ offset, size_t offset_stride, long slots,
                                                                unsigned long long *trace, int trace_cap) {
    kernel_stamp(trace, trace_cap, 93);
    switch (group_for((long)(*S.count) * B.N, slots)) {
      case 4: eam_force_tiles<4>(B, T, S, out, stride, offset, offset_stride); break;
      case 8: eam_force_tiles<8>(B, T, S, out, stride, offset, offset_stride); break;
      default: eam_force_tiles<16>(B, T, S, out, stride, offset, offset_stride); break;
    }
    kernel_stamp(trace, trace_cap, 94);
  }

  // Per-atom energy e_i = 0.5 sum V + F(rho) into scratch[R][N] (all replicas).
  template <int G>
  static __global__ void __launch_bounds__(256) k_eam_energy_atoms(BatchDev B, TablesDev T, double *__restrict__ scratch) {
    size_t g = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / G;
    int lane = threadIdx.x % G;
    if (g >= (size_t)B.R * B.N) return;
    int r = (int)(g / B.N), i = (int)(g % B.N);
    bool oob = false;
    double vs = 0;
    double rho = eam_rho_atom<G, true>(B, T, r, i, lane, &vs, oob);
    if (lane == 0) scratch[g] = fma(0.5, vs, embed(T, B.type[i], rho, false, oob));
    if (oob) atomicOr(B.fault + r, kFaultOutOfBounds);
  }

  // One CTA per replica: fixed-order tree sum of scratch[r][0..N) -> out[r].
  static __global__ void __launch_bounds__(1024) k_reduce_rows(const double *__restrict__ scratch, int N, double *__restrict__ out) {
    __shared__ double part[32];
    const double *row = scratch + (size_t)blockIdx.x * N;
    double s = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s += row[i];
    s = group_sum<32>(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      double v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
      v = group_sum<32>(v);
      if (threadIdx.x == 0) out[blockIdx.x] = v;
    }
  }

}  // namespace ofly
