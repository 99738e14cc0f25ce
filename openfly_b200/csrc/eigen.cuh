// Batched eigenvalues of dense symmetric matrices -- system::Hessian::eigenvalues() (include/libfly/system/hessian.hpp:86-93,
// Eigen's SelfAdjointEigenSolver in EigenvaluesOnly mode) for the mass-weighted Hessians of all saddle points of an
// environment at once (consumers: src/saddle/find.cpp:846-871, 1107-1135).
//
// One CTA per matrix, every matrix of the batch in flight at the same time (a C3 environment yields ~150 saddles: one per
// SM). The work is HBM-bound streaming, not a contraction, so there is nothing for the tensor cores:
//   1. Householder tridiagonalisation, unblocked, with the rank-2 update of step k-1 applied lazily inside the pass that
//      forms the matrix-vector product of step k: ONE read and ONE write of the trailing block per step (16 n^3 / 3 bytes per
//      matrix instead of 24 n^3 / 3). The matrix is held full (the lower triangle is mirrored first), so a thread owns whole
//      rows: p_i = sum_j a_ij v_j needs no transposed contributions and the only reductions are three scalars per step.
//      Loads are coalesced along columns (consecutive rows on consecutive threads); v, w of the previous step and the
//      current v sit in shared memory and are read as broadcasts.
//   2. Eigenvalues of the tridiagonal by bisection on the Sturm count, one eigenvalue per thread (n independent chains
//      of ~60 x n divisions): absolute accuracy ~ulp(|T|), ascending by construction.
// Reference-free arithmetic (any correct symmetric solver is an acceptable oracle: SURVEY 8c); tested against numpy on the
// oracle's Hessians to 1e-9 and through the prefactors to 1e-4 relative.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace ofly {

  constexpr int kEigThreads = 1024;
  constexpr int kEigRowLanes = 512;                         // threads along a column (consecutive rows: coalesced)
  constexpr int kEigColSlices = kEigThreads / kEigRowLanes;  // column slices whose partial products are summed in shared memory
  constexpr int kEigMaxRows = 3;                             // rows per thread: n <= 3 * 512 = 1536 (3N of a 512-atom cell) within 64 registers
  constexpr int kEigMaxN = kEigMaxRows * kEigRowLanes;

  __host__ __device__ inline size_t eig_smem_bytes(int n) { return (size_t)(4 + kEigColSlices) * n * sizeof(double) + 64 * sizeof(double); }

  // Sum over the CTA in a fixed order; every thread gets the result. red: 33 doubles of shared memory.
  __device__ __forceinline__ double eig_block_sum(double x, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      double y = threadIdx.x < (kEigThreads >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) y += __shfl_xor_sync(0xffffffffu, y, o);
      if (threadIdx.x == 0) red[32] = y;
    }
    __syncthreads();
    const double r = red[32];
    __syncthreads();
    return r;
  }

  // Matrix b of the batch: A = H_all + b * h_stride (n x n column-major, lower triangle valid on entry, destroyed);
  // eigenvalues ascending -> eig_all + b * n; de_all + b * 2 n is scratch for the tridiagonal (d, e).
  static __global__ void __launch_bounds__(kEigThreads, 1) k_sym_eigenvalues(double *__restrict__ H_all, size_t h_stride, int n,
                                                                            double *__restrict__ eig_all, double *__restrict__ de_all) {
    extern __shared__ __align__(16) double eig_smem[];
    double *v = eig_smem, *w = v + n, *vp = w + n, *wp = vp + n, *pp = wp + n;  // pp: [kEigColSlices][n]
    double *red = pp + (size_t)kEigColSlices * n;                                // 64 doubles
    double *A = H_all + (size_t)blockIdx.x * h_stride;
    double *d = de_all + (size_t)blockIdx.x * 2 * n, *e = d + n;
    const int tid = threadIdx.x;
    const int rl = tid % kEigRowLanes, cs = tid / kEigRowLanes;

    // mirror the lower triangle: A(j, i) = A(i, j) for i > j (writes coalesced along j)
    for (int i = cs; i < n; i += kEigColSlices) {
      for (int j = rl; j < i; j += kEigRowLanes) A[j + (size_t)i * n] = A[i + (size_t)j * n];
    }
    for (int t = tid; t < n; t += kEigThreads) vp[t] = 0.0, wp[t] = 0.0;
    __syncthreads();

    for (int k = 0; k + 1 < n; k++) {
      // (1) column k of the trailing block with the pending rank-2 update; x -> v
      const double vpk = vp[k], wpk = wp[k];
      double part = 0.0;
      for (int i = k + tid; i < n; i += kEigThreads) {
        const double a = A[i + (size_t)k * n] - (vp[i] * wpk + wp[i] * vpk);
        v[i] = a;
        if (i > k + 1) part = fma(a, a, part);
      }
      const double sigma = eig_block_sum(part, red);  // (barrier: v is complete)
      const double x0 = v[k + 1];
      if (tid == 0) d[k] = v[k];
      double beta = 0.0;
      if (sigma == 0.0) {  // nothing to annihilate below the sub-diagonal (also: frozen rows / columns, the last step)
        if (tid == 0) e[k] = x0;
      } else {
        const double mu = sqrt(fma(x0, x0, sigma));
        const double alpha = x0 <= 0.0 ? mu : -mu;
        const double v0 = x0 - alpha;
        beta = 2.0 / fma(v0, v0, sigma);
        if (tid == 0) e[k] = alpha;
        __syncthreads();
        if (tid == 0) v[k + 1] = v0;
      }
      __syncthreads();

      // (2) trailing block rows / columns k+1..n-1: apply the pending update, store, accumulate A v
      const int lo = k + 1;
      double acc[kEigMaxRows], vpi[kEigMaxRows], wpi[kEigMaxRows];
      int nrow = 0;
#pragma unroll
      for (int q = 0; q < kEigMaxRows; q++) {
        const int i = lo + rl + q * kEigRowLanes;
        acc[q] = 0.0;
        vpi[q] = i < n ? vp[i] : 0.0, wpi[q] = i < n ? wp[i] : 0.0;
        if (i < n) nrow = q + 1;
      }
      if (beta != 0.0) {
        // Two columns per iteration, all loads issued before the first store: the stores go to the array the loads come
        // from, so the compiler may not hoist the next column's loads over them by itself, and with one column in flight a
        // thread's memory-level parallelism (a few 8 B loads) caps an SM at ~30 GB/s.
        int j = lo + cs;
        for (; j + kEigColSlices < n; j += 2 * kEigColSlices) {
          const double vj0 = v[j], vpj0 = vp[j], wpj0 = wp[j];
          const double vj1 = v[j + kEigColSlices], vpj1 = vp[j + kEigColSlices], wpj1 = wp[j + kEigColSlices];
          double *c0 = A + (size_t)j * n + lo + rl, *c1 = c0 + (size_t)kEigColSlices * n;
          double a0[kEigMaxRows], a1[kEigMaxRows];
#pragma unroll
          for (int q = 0; q < kEigMaxRows; q++) {
            if (q < nrow) a0[q] = __ldcg(c0 + q * kEigRowLanes), a1[q] = __ldcg(c1 + q * kEigRowLanes);
          }
#pragma unroll
          for (int q = 0; q < kEigMaxRows; q++) {
            if (q < nrow) {
              const double x0n = a0[q] - fma(vpi[q], wpj0, wpi[q] * vpj0), x1n = a1[q] - fma(vpi[q], wpj1, wpi[q] * vpj1);
              __stcg(c0 + q * kEigRowLanes, x0n), __stcg(c1 + q * kEigRowLanes, x1n);
              acc[q] = fma(x1n, vj1, fma(x0n, vj0, acc[q]));
            }
          }
        }
        for (; j < n; j += kEigColSlices) {
          const double vj = v[j], vpj = vp[j], wpj = wp[j];
          double *col = A + (size_t)j * n + lo + rl;
#pragma unroll
          for (int q = 0; q < kEigMaxRows; q++) {
            if (q < nrow) {
              double a = __ldcg(col + q * kEigRowLanes);
              a -= fma(vpi[q], wpj, wpi[q] * vpj);
              __stcg(col + q * kEigRowLanes, a);
              acc[q] = fma(a, vj, acc[q]);
            }
          }
        }
      } else {  // no reflection this step: only the pending update has to reach memory
        for (int j = lo + cs; j < n; j += kEigColSlices) {
          const double vpj = vp[j], wpj = wp[j];
          double *col = A + (size_t)j * n + lo + rl;
#pragma unroll
          for (int q = 0; q < kEigMaxRows; q++) {
            if (q < nrow) col[q * kEigRowLanes] -= fma(vpi[q], wpj, wpi[q] * vpj);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < kEigMaxRows; q++) {
        if (q < nrow) pp[(size_t)cs * n + lo + rl + q * kEigRowLanes] = acc[q];
      }
      __syncthreads();

      // (3) p = beta A v, w = p - (beta p.v / 2) v; becomes the pending update
      double pv = 0.0;
      for (int i = lo + tid; i < n; i += kEigThreads) {
        double p = 0.0;
#pragma unroll
        for (int c = 0; c < kEigColSlices; c++) p += pp[(size_t)c * n + i];
        p *= beta;
        w[i] = p;
        pv = fma(p, v[i], pv);
      }
      const double s = eig_block_sum(pv, red);
      const double half = 0.5 * beta * s;
      for (int i = tid; i < n; i += kEigThreads) {
        const bool in = i >= lo && beta != 0.0;
        vp[i] = in ? v[i] : 0.0;
        wp[i] = in ? fma(-half, v[i], w[i]) : 0.0;
      }
      __syncthreads();
    }
    if (tid == 0) d[n - 1] = A[(size_t)(n - 1) * n + (n - 1)] - 2.0 * vp[n - 1] * wp[n - 1];
    __syncthreads();

    // ---- eigenvalues of the tridiagonal (d, e) by bisection on the Sturm count ----
    double *ds = v, *e2 = w;  // shared copies: diagonal, squared off-diagonal
    double bound = 0.0;
    for (int i = tid; i < n; i += kEigThreads) {
      const double di = d[i], el = i > 0 ? e[i - 1] : 0.0, er = i + 1 < n ? e[i] : 0.0;
      ds[i] = di, e2[i] = el * el;  // e2[i] couples i-1 and i
      bound = fmax(bound, fabs(di) + fabs(el) + fabs(er));
    }
    // Gershgorin: every eigenvalue lies in [-bound, bound]
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bound = fmax(bound, __shfl_xor_sync(0xffffffffu, bound, o));
    if ((tid & 31) == 0) red[tid >> 5] = bound;
    __syncthreads();
    for (int q = 0; q < (kEigThreads >> 5); q++) bound = fmax(bound, red[q]);
    bound = bound * (1.0 + 1e-12) + 1e-300;
    const double pivmin = 2.2250738585072014e-308 * fmax(1.0, bound * bound);
    double *eig = eig_all + (size_t)blockIdx.x * n;
    for (int kth = tid; kth < n; kth += kEigThreads) {
      double lo_x = -bound, hi_x = bound;
      for (int it = 0; it < 80; it++) {
        const double mid = 0.5 * (lo_x + hi_x);
        if (mid <= lo_x || mid >= hi_x) break;
        int below = 0;  // eigenvalues < mid = negative pivots of T - mid I
        double q = ds[0] - mid;
        if (fabs(q) < pivmin) q = -pivmin;
        below += q < 0.0;
        for (int i = 1; i < n; i++) {
          q = ds[i] - mid - e2[i] / q;
          if (fabs(q) < pivmin) q = -pivmin;
          below += q < 0.0;
        }
        if (below > kth) {
          hi_x = mid;
        } else {
          lo_x = mid;
        }
      }
      eig[kth] = 0.5 * (lo_x + hi_x);
    }
  }

}  // namespace ofly
