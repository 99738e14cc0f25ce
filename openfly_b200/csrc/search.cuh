// Device-resident minimiser and dimer saddle search: one persistent team (a thread-block cluster of
// C CTAs) owns one replica for the WHOLE search, so there is no per-step host round trip and replicas
// with different iteration counts, rebuild decisions and exits never wait for each other. Teams pull
// replicas from a device queue until the batch is drained.
//
// Restates, per replica:
//   minimise::LBFGS::minimise        src/minimise/LBFGS/lbfgs.cpp:29-115
//   minimise::StepLBFGS::newton_step src/minimise/LBFGS/lbfgs_core.cpp:29-87
//   saddle::Rotor::eff_gradient      src/saddle/rotor.cpp:33-164
//   saddle::Dimer::find_sp           src/saddle/dimer.cpp:28-171
// calling the same neighbour (neigh.cuh) and EAM (eam.cuh) device functions as the grid kernels.
//
// Vector algebra: every 3N-vector op is one strided pass by the whole team; dot products are fused
// into the pass that produces their operand (the L-BFGS two-loop costs one pass + one reduction per
// history pair) and reduced warp-shuffle -> shared memory -> (cluster: distributed shared memory) in a
// fixed order, so all threads of the team obtain bit-identical scalars and take the same branches
// without any broadcast. Per-replica scalars (trust, acc, curvature, counters) live in registers.
#pragma once

#include <cooperative_groups.h>

#include "eam.cuh"
#include "neigh.cuh"

namespace ofly {

  namespace cg = cooperative_groups;

  constexpr int kSearchThreads = 512;
  constexpr int kSearchGroup = 8;  // lanes per atom inside the search kernel
  constexpr int kMaxHist = 16;     // upper bound on L-BFGS memory
  constexpr int kMaxRed = 8;       // values per fused reduction

  // Exit / status codes beyond the reference's enums.
  constexpr int kStatusOutOfBounds = -5;
  constexpr int kStatusCapacity = -6;

  struct SearchParams {
    int mode;  // 0: LBFGS::minimise, 1: Dimer::find_sp
    // translation / minimiser options
    int n_hist, iter_max;
    double f2norm, proj_tol, max_trust, min_trust, grow_trust, shrink_trust;
    double skin;  // list radius = base radius + skin; rebuild when acc > skin / 2
    // dimer
    int hist_check_freq, use_history;
    double convex_max;
    int rot_n, iter_max_rot, relax_in_convex;
    double delta_r, rot_theta_tol;
    double cos_theta_tol;
    int count_frozen;
    long n_sp;
    // io (device pointers)
    const double *pos_in, *axis_in, *in_min, *hist_sp;
    double *pos_out, *axis_out, *grad_out;
    int *status;
    long *counters;  // [R][4] or null
    double *ws;      // per-team workspace
    size_t ws_stride;
    int *queue;
  };

  // ---- team = cluster of C CTAs ------------------------------------------------------------------
  template <int C>
  struct Team {
    int tid, nthreads, rank;
    double *red;  // shared: [2][C][kMaxRed] cluster slots, then [2][32][kMaxRed] warp partials
    int parity = 0;

    __device__ Team(double *smem) : red(smem) {
      if constexpr (C > 1) {
        rank = (int)cg::this_cluster().block_rank();
      } else {
        rank = 0;
      }
      tid = rank * blockDim.x + threadIdx.x;
      nthreads = C * blockDim.x;
    }

    __device__ __forceinline__ void sync() {
      if constexpr (C > 1) {
        cg::this_cluster().sync();
      } else {
        __syncthreads();
      }
    }

    // Sum NV values over the team; result (identical bits on every thread) overwrites v.
    template <int NV>
    __device__ __forceinline__ void reduce(double (&v)[NV]) {
      static_assert(NV <= kMaxRed, "too many fused values");
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
      double *warp_part = red + 2 * C * kMaxRed + (size_t)parity * 32 * kMaxRed;
      double *slots = red + (size_t)parity * C * kMaxRed;
#pragma unroll
      for (int q = 0; q < NV; q++) {
        double x = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) warp_part[w * kMaxRed + q] = x;
      }
      __syncthreads();
      if constexpr (C == 1) {
#pragma unroll
        for (int q = 0; q < NV; q++) {
          double x = 0;
          for (int k = 0; k < nw; k++) x += warp_part[k * kMaxRed + q];
          v[q] = x;
        }
      } else {
        if (threadIdx.x < NV) {
          double x = 0;
          for (int k = 0; k < nw; k++) x += warp_part[k * kMaxRed + threadIdx.x];
          cg::cluster_group cl = cg::this_cluster();
          for (int c = 0; c < C; c++) {
            double *remote = cl.map_shared_rank(slots, c);
            remote[rank * kMaxRed + threadIdx.x] = x;
          }
        }
        cg::this_cluster().sync();
#pragma unroll
        for (int q = 0; q < NV; q++) {
          double x = 0;
#pragma unroll
          for (int c = 0; c < C; c++) x += slots[c * kMaxRed + q];
          v[q] = x;
        }
      }
      parity ^= 1;
    }
  };

  __host__ __device__ constexpr size_t team_smem_doubles(int C) { return 2 * (size_t)C * kMaxRed + 2 * 32 * kMaxRed; }

  // ---- per-replica building blocks ------------------------------------------------------------------

  // neigh::List::rebuild(x) by the team (same five steps as the grid kernels in neigh.cuh).
  template <int C>
  __device__ __noinline__ void team_rebuild(Team<C> &tm, const BatchDev &B, int r, const double *x, int *scan_smem) {
    int *hist = B.cell_count + (size_t)r * B.grid.ncells;
    for (int c = tm.tid; c < B.grid.ncells; c += tm.nthreads) hist[c] = 0;
    tm.sync();
    for (int i = tm.tid; i < B.N; i += tm.nthreads) nb_canon_count(B, x, r, i);
    tm.sync();
    if (tm.rank == 0) nb_scan_block(B, r, scan_smem);
    tm.sync();
    for (int i = tm.tid; i < B.N; i += tm.nthreads) nb_scatter(B, r, i);
    tm.sync();
    for (int c = tm.tid; c < B.grid.ncells; c += tm.nthreads) nb_sort_cell(B, r, c);
    tm.sync();
    constexpr int G = kSearchGroup;
    const int lane = tm.tid % G, ngroups = tm.nthreads / G;
    for (int i = tm.tid / G; i < B.N; i += ngroups) nb_build_atom<G>(B, r, i, lane);
    tm.sync();
  }

  // potential::EAM::gradient by the team; returns false if a spline table was overrun.
  template <int C>
  __device__ __noinline__ bool team_gradient(Team<C> &tm, const BatchDev &B, const TablesDev &T, int r, double *out) {
    constexpr int G = kSearchGroup;
    const int lane = tm.tid % G, ngroups = tm.nthreads / G;
    bool oob = false;
    for (int i = tm.tid / G; i < B.N; i += ngroups) {
      double rho = eam_rho_atom<G, false>(B, T, r, i, lane, nullptr, oob);
      if (lane == 0) reinterpret_cast<double *>(B.pos4 + (size_t)r * B.N + i)[3] = embed(T, B.type[i], rho, true, oob);
    }
    if (oob) atomicOr(B.fault + r, kFaultOutOfBounds);
    tm.sync();
    for (int i = tm.tid / G; i < B.N; i += ngroups) {
      double gx = 0, gy = 0, gz = 0;
      if (!B.frozen[i]) eam_force_atom<G>(B, T, r, i, lane, gx, gy, gz, oob);
      if (lane == 0) out[3 * i] = gx, out[3 * i + 1] = gy, out[3 * i + 2] = gz;
    }
    if (oob) atomicOr(B.fault + r, kFaultOutOfBounds);
    tm.sync();
    return (__ldcg(B.fault + r) & kFaultOutOfBounds) == 0;
  }

  // L-BFGS history of one optimiser (StepLBFGS, lbfgs_core.hpp:101-118).
  struct Core {
    double *prev_x, *prev_g, *step, *s, *y;  // s, y: [n][n3p]
    int n, k;
    size_t n3p;
    double rho[kMaxHist], alpha[kMaxHist];
  };

  // StepLBFGS::newton_step: core.step <- H * g (lbfgs_core.cpp:29-87). Also returns step.step and
  // step.w (w may be null) from the final pass.
  template <int C>
  __device__ __noinline__ void newton_step(Team<C> &tm, Core &c, const double *x, const double *g, int n3, const double *w, double &out_nrm2,
                              double &out_dotw) {
    const int n = c.n;
    const int prev = c.k > 0 ? (c.k - 1) % n : 0;
    const int incur = c.k <= n ? 0 : c.k - n;
    const int bound = c.k <= n ? c.k : n;
    double *q = c.step;

    if (c.k == 0) {  // identity Hessian: step = g
      double a[2] = {0, 0};
      for (int t = tm.tid; t < n3; t += tm.nthreads) {
        double gv = g[t];
        c.prev_x[t] = x[t], c.prev_g[t] = gv, q[t] = gv;
        a[0] = fma(gv, gv, a[0]);
        if (w) a[1] = fma(gv, w[t], a[1]);
      }
      tm.reduce<2>(a);
      out_nrm2 = a[0], out_dotw = a[1];
      c.k = 1;
      return;
    }

    double yy, d;
    {  // newest pair (s, y), prev <- current, q <- g, first dot of loop 1
      double *s = c.s + (size_t)prev * c.n3p, *y = c.y + (size_t)prev * c.n3p;
      double a[3] = {0, 0, 0};
      for (int t = tm.tid; t < n3; t += tm.nthreads) {
        double xv = x[t], gv = g[t];
        double sv = xv - c.prev_x[t], yv = gv - c.prev_g[t];
        s[t] = sv, y[t] = yv, c.prev_x[t] = xv, c.prev_g[t] = gv, q[t] = gv;
        a[0] = fma(sv, yv, a[0]), a[1] = fma(yv, yv, a[1]), a[2] = fma(sv, gv, a[2]);
      }
      tm.reduce<3>(a);
      c.rho[prev] = 1.0 / fabs(a[0]);  // abs: never an ascent direction (lbfgs_core.cpp:50)
      yy = a[1], d = a[2];
    }
    const double h0 = 1.0 / c.rho[prev] / yy;  // initial Hessian scaling (lbfgs_core.cpp:70)

    // loop 1, newest -> oldest: alpha_j = rho_j s_j.q ; q -= alpha_j y_j   (dot for the next pair fused)
    for (int i = bound - 1; i >= 0; --i) {
      const int j = (i + incur) % n;
      const double alpha = c.rho[j] * d;
      c.alpha[j] = alpha;
      const double *y = c.y + (size_t)j * c.n3p;
      double a[1] = {0};
      if (i > 0) {
        const double *sn = c.s + (size_t)((i - 1 + incur) % n) * c.n3p;
        for (int t = tm.tid; t < n3; t += tm.nthreads) {
          double qv = fma(-alpha, y[t], q[t]);
          q[t] = qv;
          a[0] = fma(sn[t], qv, a[0]);
        }
      } else {  // last pair: apply H0 and start loop 2 with the same y
        for (int t = tm.tid; t < n3; t += tm.nthreads) {
          double rv = fma(-alpha, y[t], q[t]) * h0;
          q[t] = rv;
          a[0] = fma(y[t], rv, a[0]);
        }
      }
      tm.reduce<1>(a);
      d = a[0];
    }
    // loop 2, oldest -> newest: beta = rho_j y_j.r ; r += (alpha_j - beta) s_j
    for (int i = 0; i < bound; ++i) {
      const int j = (i + incur) % n;
      const double coef = c.alpha[j] - c.rho[j] * d;
      const double *s = c.s + (size_t)j * c.n3p;
      if (i + 1 < bound) {
        const double *yn = c.y + (size_t)((i + 1 + incur) % n) * c.n3p;
        double a[1] = {0};
        for (int t = tm.tid; t < n3; t += tm.nthreads) {
          double rv = fma(coef, s[t], q[t]);
          q[t] = rv;
          a[0] = fma(yn[t], rv, a[0]);
        }
        tm.reduce<1>(a);
        d = a[0];
      } else {
        double a[2] = {0, 0};
        for (int t = tm.tid; t < n3; t += tm.nthreads) {
          double rv = fma(coef, s[t], q[t]);
          q[t] = rv;
          a[0] = fma(rv, rv, a[0]);
          if (w) a[1] = fma(rv, w[t], a[1]);
        }
        tm.reduce<2>(a);
        out_nrm2 = a[0], out_dotw = a[1];
      }
    }
    c.k++;
  }

  struct Work {  // per-team vectors (each n3p doubles)
    double *x, *g;          // positions; gradient (minimise) or effective gradient (dimer)
    double *axis, *g0, *g1, *g1p, *dprev, *axisp, *dg;
    Core trans, rot;
  };

  struct Tally {
    long grad_evals = 0, rebuilds = 0, updates = 0, steps = 0;
  };

  // Rotor::eff_gradient (rotor.cpp:49-164). Leaves the effective gradient in W.g, the rotated axis in
  // W.axis and the list positions back where they were. If `step` is non-null the final pass also
  // returns eff.eff and eff.step. Returns false on Spline::OutOfBounds.
  template <int C>
  __device__ __noinline__ bool eff_gradient(Team<C> &tm, const BatchDev &B, const TablesDev &T, const SearchParams &P, Work &W, int r, int n3,
                               const double *step, double &curv, double &eff2, double &eff_step, Tally &tally) {
    const double dr = P.delta_r;
    double4 *pos = B.pos4 + (size_t)r * B.N;
    W.rot.k = 0;  // m_core.clear()

    if (!team_gradient(tm, B, T, r, W.g0)) return false;
    tally.grad_evals++;
    // delta = -dr * axis ; nl.update(delta) ; delta_prev = delta     (rotor.cpp:77-83)
    for (int i = tm.tid; i < B.N; i += tm.nthreads) {
      double4 p = ld256(pos + i);
      double d0 = -dr * W.axis[3 * i], d1 = -dr * W.axis[3 * i + 1], d2 = -dr * W.axis[3 * i + 2];
      W.dprev[3 * i] = d0, W.dprev[3 * i + 1] = d1, W.dprev[3 * i + 2] = d2;
      p.x -= d0, p.y -= d1, p.z -= d2;
      st256(pos + i, p);
    }
    tally.updates++;
    tm.sync();
    if (!team_gradient(tm, B, T, r, W.g1)) return false;
    tally.grad_evals++;

    const double inv_n = 1.0 / (double)B.N;
    for (int it = 0;; it++) {
      // torque: dg = (g1 - g0) - ((g1 - g0).axis) axis, minus its mean if nothing is frozen (rotor.cpp:89-94)
      double a7[7] = {0, 0, 0, 0, 0, 0, 0};
      for (int t = tm.tid; t < n3; t += tm.nthreads) {
        double dv = W.g1[t] - W.g0[t], av = W.axis[t];
        a7[0] = fma(dv, av, a7[0]);
        int c = t % 3;
        a7[1 + c] += dv, a7[4 + c] += av;
      }
      tm.reduce<7>(a7);
      const double pax = a7[0];
      double mean[3] = {0, 0, 0};
      if (P.count_frozen == 0) {
        for (int c = 0; c < 3; c++) mean[c] = (a7[1 + c] - pax * a7[4 + c]) * inv_n;
      }
      for (int t = tm.tid; t < n3; t += tm.nthreads) {
        W.dg[t] = (W.g1[t] - W.g0[t]) - pax * W.axis[t] - mean[t % 3];
      }
      tm.sync();

      double nrm2_unused, th_ax;
      newton_step(tm, W.rot, W.axis, W.dg, n3, W.axis, nrm2_unused, th_ax);
      double *theta = W.rot.step;
      // theta -= (theta.axis) axis ; normalise ; b_1 = (g1-g0).theta / dr   (rotor.cpp:99-103)
      double a2[2] = {0, 0};
      for (int t = tm.tid; t < n3; t += tm.nthreads) {
        double tv = fma(-th_ax, W.axis[t], theta[t]);
        theta[t] = tv;
        a2[0] = fma(tv, tv, a2[0]);
        a2[1] = fma(W.g1[t] - W.g0[t], tv, a2[1]);
      }
      tm.reduce<2>(a2);
      const double inv_th = 1 / sqrt(a2[0]);
      const double b_1 = a2[1] * inv_th / dr;
      const double c_x0 = pax / dr;
      const double theta_1 = -0.5 * atan(b_1 / fabs(c_x0));

      if (fabs(theta_1) < P.rot_theta_tol || (it >= P.iter_max_rot && c_x0 < 0) || it > 10 * P.iter_max_rot) {
        curv = c_x0;
        break;
      }
      // trial axis, move the list end-point from the old offset to the new one (rotor.cpp:111-116)
      const double c1 = cos(theta_1), s1 = sin(theta_1);
      for (int i = tm.tid; i < B.N; i += tm.nthreads) {
        double4 p = ld256(pos + i);
        double up[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          int t = 3 * i + c;
          double tn = theta[t] * inv_th;
          theta[t] = tn;
          double ap = W.axis[t] * c1 + tn * s1;
          W.axisp[t] = ap;
          double nw = -dr * ap;
          up[c] = nw - W.dprev[t];
          W.dprev[t] = nw;
        }
        p.x -= up[0], p.y -= up[1], p.z -= up[2];
        st256(pos + i, p);
      }
      tally.updates++;
      tm.sync();
      if (!team_gradient(tm, B, T, r, W.g1p)) return false;
      tally.grad_evals++;

      double a1[1] = {0};
      for (int t = tm.tid; t < n3; t += tm.nthreads) a1[0] = fma(W.g1p[t] - W.g0[t], W.axisp[t], a1[0]);
      tm.reduce<1>(a1);
      const double c_x1 = a1[0] / dr;
      const double a_1 = (c_x0 - c_x1 + b_1 * sin(2 * theta_1)) / (1 - cos(2 * theta_1));
      double theta_min = 0.5 * atan(b_1 / a_1);
      if (a_1 * cos(2 * theta_min) - a_1 + b_1 * sin(2 * theta_min) > 0) theta_min += M_PI / 2;  // extremum was a maximum

      const double cm = cos(theta_min), sm = sin(theta_min);
      const double k1 = sin(theta_1 - theta_min) / s1, k2 = sm / s1, k3 = 1 - cm - sm * tan(0.5 * theta_1);
      double ac[1] = {0};
      for (int t = tm.tid; t < n3; t += tm.nthreads) {
        double an = W.axis[t] * cm + theta[t] * sm;
        W.axis[t] = an;
        double g0v = W.g0[t];
        double gn = k1 * W.g1[t] + k2 * W.g1p[t] + k3 * g0v;  // interpolated end-point gradient (rotor.cpp:132-134)
        W.g1[t] = gn;
        ac[0] = fma(gn - g0v, an, ac[0]);
      }
      if (fabs(theta_min) < P.rot_theta_tol) {
        tm.reduce<1>(ac);
        curv = ac[0] / dr;
        break;
      }
      tm.sync();
    }

    // undo the end-point offset (rotor.cpp:150-151) and form the effective gradient (:153-161)
    double a7[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int t = tm.tid; t < n3; t += tm.nthreads) {
      double gv = W.g0[t], av = W.axis[t];
      a7[0] = fma(gv, av, a7[0]);
      int c = t % 3;
      a7[1 + c] += gv, a7[4 + c] += av;
    }
    tm.reduce<7>(a7);
    const bool parallel_only = !P.relax_in_convex && curv > 0;
    const double p0 = a7[0];
    const double ca = parallel_only ? 0.0 : 1.0, cb = parallel_only ? -p0 : -2 * p0;  // eff = ca g0 + cb axis
    double mean[3] = {0, 0, 0};
    if (P.count_frozen == 0) {
      for (int c = 0; c < 3; c++) mean[c] = (ca * a7[1 + c] + cb * a7[4 + c]) * inv_n;
    }
    double a2[2] = {0, 0};
    for (int i = tm.tid; i < B.N; i += tm.nthreads) {
      double4 p = ld256(pos + i);
      p.x += W.dprev[3 * i], p.y += W.dprev[3 * i + 1], p.z += W.dprev[3 * i + 2];
      st256(pos + i, p);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        int t = 3 * i + c;
        double e = ca * W.g0[t] + cb * W.axis[t] - mean[c];
        W.g[t] = e;
        a2[0] = fma(e, e, a2[0]);
        if (step) a2[1] = fma(e, step[t], a2[1]);
      }
    }
    tally.updates++;
    tm.reduce<2>(a2);
    eff2 = a2[0], eff_step = a2[1];
    return true;
  }

  // Trust-limited step shared by the minimiser and the dimer (lbfgs.cpp:83-100, dimer.cpp:133-151):
  // Hg *= min(1, trust/|Hg|); acc += |Hg|; x -= Hg; then rebuild if acc > skin/2 else update(Hg).
  template <int C>
  __device__ void take_step(Team<C> &tm, const BatchDev &B, const SearchParams &P, Work &W, int r, int n3, double nrm2, double trust,
                            double &acc, int *scan_smem, Tally &tally) {
    const double nrm = sqrt(nrm2);
    const double scale = fmin(1.0, trust / nrm);
    acc += scale * nrm;
    const bool rebuild = acc > 0.5 * P.skin;
    double *step = W.trans.step;
    double4 *pos = B.pos4 + (size_t)r * B.N;
    for (int i = tm.tid; i < B.N; i += tm.nthreads) {
      double h[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        int t = 3 * i + c;
        h[c] = step[t] * scale;
        step[t] = h[c];
        W.x[t] -= h[c];
      }
      if (!rebuild) {
        double4 p = ld256(pos + i);
        p.x -= h[0], p.y -= h[1], p.z -= h[2];
        st256(pos + i, p);
      }
    }
    tm.sync();
    if (rebuild) {
      team_rebuild(tm, B, r, W.x, scan_smem);
      acc = 0;
      tally.rebuilds++;
    } else {
      tally.updates++;
    }
  }

  __device__ __forceinline__ double next_trust(const SearchParams &P, double trust, double proj) {
    if (proj < -P.proj_tol) return fmax(P.min_trust, P.shrink_trust * trust);
    if (proj > P.proj_tol) return fmin(P.max_trust, P.grow_trust * trust);
    return trust;
  }

  template <int C>
  __device__ int run_minimise(Team<C> &tm, const BatchDev &B, const TablesDev &T, const SearchParams &P, Work &W, int r, int n3,
                              int *scan_smem, Tally &tally) {
    W.trans.k = 0;
    team_rebuild(tm, B, r, W.x, scan_smem);
    tally.rebuilds++;
    if (__ldcg(B.fault + r) & kFaultCapacity) return kStatusCapacity;
    if (!team_gradient(tm, B, T, r, W.g)) return kStatusOutOfBounds;
    tally.grad_evals++;
    double a2[2] = {0, 0};
    for (int t = tm.tid; t < n3; t += tm.nthreads) a2[0] = fma(W.g[t], W.g[t], a2[0]);
    tm.reduce<2>(a2);
    double mag_g = a2[0];
    double trust = P.min_trust, acc = 0;

    for (int i = 0; i < P.iter_max; ++i) {
      if (mag_g < P.f2norm * P.f2norm) return 0;
      double nrm2, unused;
      newton_step(tm, W.trans, W.x, W.g, n3, (const double *)nullptr, nrm2, unused);
      take_step(tm, B, P, W, r, n3, nrm2, trust, acc, scan_smem, tally);
      if (__ldcg(B.fault + r) & kFaultCapacity) return kStatusCapacity;
      if (!team_gradient(tm, B, T, r, W.g)) return kStatusOutOfBounds;
      tally.grad_evals++, tally.steps++;
      double b2[2] = {0, 0};
      const double *step = W.trans.step;
      for (int t = tm.tid; t < n3; t += tm.nthreads) {
        double gv = W.g[t];
        b2[0] = fma(gv, gv, b2[0]), b2[1] = fma(gv, step[t], b2[1]);
      }
      tm.reduce<2>(b2);
      mag_g = b2[0];
      trust = next_trust(P, trust, b2[1]);
    }
    return 1;  // iter_max
  }

  template <int C>
  __device__ int run_dimer(Team<C> &tm, const BatchDev &B, const TablesDev &T, const SearchParams &P, Work &W, int r, int n3,
                           int *scan_smem, Tally &tally) {
    W.trans.k = 0;
    team_rebuild(tm, B, r, W.x, scan_smem);
    tally.rebuilds++;
    if (__ldcg(B.fault + r) & kFaultCapacity) return kStatusCapacity;
    double curv, mag_g, proj;
    if (!eff_gradient(tm, B, T, P, W, r, n3, (const double *)nullptr, curv, mag_g, proj, tally)) return kStatusOutOfBounds;

    double trust = P.min_trust, acc = 0, convex_count = 0;

    for (int i = 0; i < P.iter_max; ++i) {
      if (mag_g < P.f2norm * P.f2norm) return 0;           // success
      if (convex_count >= P.convex_max) return -1;         // convex
      if (P.use_history && i % P.hist_check_freq == 0) {   // dimer.cpp:102-131
        double ctheta = -1;
        for (long k = 0; k < P.n_sp; k++) {
          const double *sp = P.hist_sp + (size_t)k * n3;
          double a3[3] = {0, 0, 0};
          for (int t = tm.tid; t < n3; t += tm.nthreads) {
            double m = P.in_min[t];
            double a = W.x[t] - m, b = sp[t] - m;
            a3[0] = fma(a, b, a3[0]), a3[1] = fma(b, b, a3[1]), a3[2] = fma(a, a, a3[2]);
          }
          tm.reduce<3>(a3);
          ctheta = fmax(ctheta, a3[0] / (sqrt(a3[1]) * sqrt(a3[2])));
          if (ctheta > P.cos_theta_tol) return -3;  // collision
        }
      }
      double nrm2, unused;
      newton_step(tm, W.trans, W.x, W.g, n3, (const double *)nullptr, nrm2, unused);
      take_step(tm, B, P, W, r, n3, nrm2, trust, acc, scan_smem, tally);
      if (__ldcg(B.fault + r) & kFaultCapacity) return kStatusCapacity;
      if (!eff_gradient(tm, B, T, P, W, r, n3, W.trans.step, curv, mag_g, proj, tally)) return kStatusOutOfBounds;
      tally.steps++;
      convex_count = curv > 0 ? convex_count + 1 : 0;
      trust = next_trust(P, trust, proj);
    }
    return -2;  // iter_max
  }

  __device__ __forceinline__ void carve_core(Core &c, double *&p, int n, size_t n3p) {
    c.n = n, c.k = 0, c.n3p = n3p;
    c.prev_x = p, p += n3p;
    c.prev_g = p, p += n3p;
    c.step = p, p += n3p;
    c.s = p, p += (size_t)n * n3p;
    c.y = p, p += (size_t)n * n3p;
  }

  __host__ __device__ inline size_t search_ws_doubles(int mode, int n_hist, int rot_n, size_t n3p) {
    size_t v = 2 + (3 + 2 * (size_t)n_hist);
    if (mode == 1) v += 7 + (3 + 2 * (size_t)rot_n);
    return v * n3p;
  }

  template <int C>
  __global__ void __launch_bounds__(kSearchThreads, 1) k_search(BatchDev B, TablesDev T, SearchParams P) {
    __shared__ double red_smem[team_smem_doubles(C)];
    __shared__ int scan_smem[33];
    __shared__ int next_replica;

    Team<C> tm(red_smem);
    const int team_id = blockIdx.x / C;
    const int n3 = 3 * B.N;
    const size_t n3p = ((size_t)n3 + 3) & ~(size_t)3;

    Work W;
    {
      double *p = P.ws + (size_t)team_id * P.ws_stride;
      W.x = p, p += n3p;
      W.g = p, p += n3p;
      carve_core(W.trans, p, P.n_hist, n3p);
      if (P.mode == 1) {
        W.axis = p, p += n3p;
        W.g0 = p, p += n3p;
        W.g1 = p, p += n3p;
        W.g1p = p, p += n3p;
        W.dprev = p, p += n3p;
        W.axisp = p, p += n3p;
        W.dg = p, p += n3p;
        carve_core(W.rot, p, P.rot_n, n3p);
      }
    }

    for (;;) {
      // the team leader claims the next replica; every CTA of the cluster learns it through its own smem
      if (tm.tid == 0) {
        int r = atomicAdd(P.queue, 1);
        if constexpr (C > 1) {
          cg::cluster_group cl = cg::this_cluster();
          for (int c = 0; c < C; c++) *cl.map_shared_rank(&next_replica, c) = r;
        } else {
          next_replica = r;
        }
      }
      tm.sync();
      const int r = next_replica;
      tm.sync();
      if (r >= B.R) break;

      for (int t = tm.tid; t < n3; t += tm.nthreads) {
        W.x[t] = P.pos_in[(size_t)r * n3 + t];
        if (P.mode == 1) W.axis[t] = P.axis_in[(size_t)r * n3 + t];
      }
      if (tm.tid == 0) B.fault[r] = 0;
      tm.sync();

      Tally tally;
      int status = P.mode == 0 ? run_minimise(tm, B, T, P, W, r, n3, scan_smem, tally) : run_dimer(tm, B, T, P, W, r, n3, scan_smem, tally);

      tm.sync();
      for (int t = tm.tid; t < n3; t += tm.nthreads) {
        P.pos_out[(size_t)r * n3 + t] = W.x[t];
        if (P.mode == 1) P.axis_out[(size_t)r * n3 + t] = W.axis[t];
        if (P.mode == 0 && P.grad_out) P.grad_out[(size_t)r * n3 + t] = W.g[t];
      }
      if (tm.tid == 0) {
        P.status[r] = status;
        if (P.counters) {
          long *c = P.counters + 4 * (size_t)r;
          c[0] = tally.grad_evals, c[1] = tally.rebuilds, c[2] = tally.updates, c[3] = tally.steps;
        }
      }
      tm.sync();
    }
  }

}  // namespace ofly
