// Device-resident minimiser and dimer saddle search, batched in lock-step over replicas.
//
// Restates, per replica:
//   minimise::LBFGS::minimise        src/minimise/LBFGS/lbfgs.cpp:29-115
//   minimise::StepLBFGS::newton_step src/minimise/LBFGS/lbfgs_core.cpp:29-87
//   saddle::Rotor::eff_gradient      src/saddle/rotor.cpp:33-164
//   saddle::Dimer::find_sp           src/saddle/dimer.cpp:28-171
//
// Every algorithm above is a chain of 3N-vector passes punctuated by EAM gradient evaluations (and,
// when the skin is used up, a neighbour rebuild). The gradient is the expensive, embarrassingly
// parallel part, so the batch advances in *ticks*: in one tick every still-searching replica runs its
// own vector algebra up to its next gradient request (k_segment: one team = one CTA or one cluster of
// 2-16 CTAs per entry of the list of active replicas; replicas may be at completely different places
// of the algorithm), then ONE set of grid kernels (neigh.cuh / eam.cuh) rebuilds the lists of the
// replicas that asked for it and evaluates the gradient of all active replicas with the whole chip.
// A tick is a fixed sequence of kernel nodes (segment, six rebuild kernels behind an IF node that is
// only armed when somebody asked, two pair kernels); the host enqueues it as a CUDA graph whose body
// repeats on the device (WHILE conditional node) until no replica is active -- no per-step host round
// trip, no host-side branch. Finished replicas drop out of the list, so the chip stays as full as the
// batch allows; for most ticks of a batch only a few replicas are left (search lengths vary 50x) and a
// tick is a latency chain, which is what the choices below are made for.
//
// Vector algebra: every pass is strided over the team; dot products are fused into the pass that
// produces their operand and reduced warp-shuffle -> shared memory -> (cluster: st.async posts into
// every peer's shared memory, counted in by mbarriers) in a fixed order, so all threads of a team
// obtain bit-identical scalars and take the same branches without a broadcast. Every pass is indexed
// by vector component, so a thread only re-reads data it wrote itself and no pass needs a memory fence
// or a cluster barrier. The L-BFGS step takes two passes and two reductions whatever the history
// length (newton_step_gram); transcendental scalars are computed by one warp and shared.
// Per-replica scalars (phase, trust, acc, curvature, L-BFGS rho_j and Gram entries, counters) persist
// in RepState. OFLY_TRACE=<replica> records time stamps inside the kernels (trace_mark).
#pragma once

#include <cooperative_groups.h>

#include "eam.cuh"
#include "neigh.cuh"

namespace ofly {

  namespace cg = cooperative_groups;

  constexpr int kMaxHist = 16;  // upper bound on L-BFGS memory
  constexpr int kMaxRed = 8;    // values per fused reduction
  constexpr int kGramVals = 5;  // dot products per history pair of the single-pass L-BFGS step
  constexpr int kGramMax = kGramVals * kMaxHist;

  // Exit / status codes beyond the reference's enums.
  constexpr int kStatusOutOfBounds = -5;
  constexpr int kStatusCapacity = -6;
  constexpr int kStatusUnfinished = -7;  // tick budget exhausted (should not happen)

  // Where a replica resumes when its requested gradient has arrived.
  enum Phase : int { kInit = 0, kMinGrad = 1, kRotG0 = 2, kRotG1 = 3, kRotG1p = 4, kDone = 5 };

  struct CoreState {  // StepLBFGS (lbfgs_core.hpp:101-118): history fill + rho_j
    int k, pad;
    double rho[kMaxHist];
    // Gram entries of the stored pairs, [j * n + k]: sy = s_j . y_k, yy = y_j . y_k (newton_step_gram)
    double sy[kMaxHist * kMaxHist], yy[kMaxHist * kMaxHist];
  };

  struct RepState {  // one per resident slot
    int phase, status;
    int iter;     // translation / minimiser step index i
    int rot_it;   // rotor iteration
    int first;    // no step taken yet (first effective gradient / first gradient)
    int need_rebuild;
    int qid, pad;  // index of the search this slot is running, in the caller's queue of start states
    unsigned long long gdst;  // offset (doubles) inside the replica's workspace for the next gradient
    double trust, acc, convex_count, curv, mag_g;
    double c_x0, b_1, theta_1;  // rotor scalars carried over the trial-gradient request
    CoreState trans, rot;
    long tally[4];  // grad_evals, rebuilds, updates, steps
  };

  struct Control {
    int n_active, n_rebuild;  // entries in the selection lists for the grid kernels of this tick
    int ticket;               // teams that have finished their segment in this tick
    int tick;
    long long kernels;  // kernels executed so far (an IF-guarded rebuild only counts when it ran)
    int queue_head;     // next search of the queue that no slot has claimed yet
    int fault_or;       // OR of the fault bits that ended a search (slots are re-armed, their own fault words reset)
    int if_state[3][8]; // value last given to the rebuild IF handle of tick u of the body of stage L's graph (0 at launch)
  };

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
    long n_queue;  // searches in the caller's queue (>= 1); the io arrays below are indexed by search, not by slot
    // io (device pointers)
    const double *pos_in, *axis_in, *in_min, *hist_sp;
    double *pos_out, *axis_out, *grad_out;
    int *status;
    long *counters;  // [n_queue][4] or null
    // engine
    double *ws;  // [R][ws_stride] vectors
    size_t ws_stride;
    RepState *state;  // [R]
    Control *ctl;
    int *active, *rebuild;  // [R] selection lists
    int *flags;             // [R] per entry of `active`: bit 0 still active, bit 1 wants a rebuild (written by its team)
    int max_ticks;
    int exit_below;  // the device-side loop of this graph ends once at most this many replicas are active (0: runs to the end)
    int use_cond;
    unsigned long long *trace;  // tuning aid (OFLY_TRACE): [0] = count, then (mark, globaltimer ns) pairs of replica trace_replica
    int trace_replica, trace_cap;
    int gram;  // 1: single-pass L-BFGS step (newton_step_gram), 0: pass-per-pair two-loop (newton_step)
    cudaGraphConditionalHandle cond;
    int stage, body_slot;  // which graph (stage of the device-side loop) and which tick of its body this launch is
    int fused_rebuild;     // the rebuild of a tick is ONE cooperative kernel (k_nb_rebuild_fused) instead of six behind an IF node
    int use_cond_rebuild;  // the rebuild kernels of a tick sit in an IF node that is skipped when nobody asked
    cudaGraphConditionalHandle cond_rebuild;
  };

  // Time stamp `mark` for the traced replica (thread 0 of its team only; off unless OFLY_TRACE is set).
  __device__ __forceinline__ void trace_mark(const SearchParams &P, int r, int tid, int mark) {
    if (P.trace != nullptr && tid == 0 && r == P.trace_replica) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      unsigned long long slot = atomicAdd(P.trace, 1ull);
      if (slot < (unsigned long long)P.trace_cap) P.trace[1 + 2 * slot] = (unsigned long long)mark, P.trace[2 + 2 * slot] = t;
    }
  }

  // ---- team = one CTA or a cluster of C CTAs ------------------------------------------------------
  // Cluster-wide sums go through distributed shared memory without a cluster barrier: every CTA posts
  // its partial sums into all peers' slots with st.async, which also counts the bytes in on the
  // receiver's mbarrier; a CTA only waits on its *own* mbarrier (CTA-scope acquire). A cluster-scope
  // barrier would invalidate the L1 on every reduction (and there are 20-40 per segment).
  __device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
  __device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, int cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
  }
  __device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  }
  __device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  }
  __device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
  }
  // 8-byte store into a peer CTA's shared memory that completes 8 bytes on that CTA's mbarrier.
  __device__ __forceinline__ void dsmem_post(uint32_t remote_addr, double v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
                 "l"(__double_as_longlong(v)), "r"(remote_bar)
                 : "memory");
  }

  // Shared-memory block of the single-pass L-BFGS step, one per CTA.
  template <int C>
  struct GramShared {
    double sy[kMaxHist * kMaxHist], yy[kMaxHist * kMaxHist];  // working copies of CoreState::sy / yy
    double rho[kMaxHist];
    double part[16][kGramMax];   // per-warp partial sums (a CTA has at most 16 warps)
    double slots[C][kGramMax];   // per-CTA sums posted by every CTA of the cluster
    double out[kGramMax];        // team-wide sums, order index i (oldest pair first): {s_p.y_i, y_p.s_i, y_p.y_i, g.s_i, g.y_i}
    double alpha[kMaxHist], w[kMaxHist], h0;
    unsigned long long off[kMaxHist];  // element offset of pair i's row inside Core::s / Core::y
    int slot[kMaxHist];                // ring-buffer slot of pair i
  };

  template <int C>
  struct Team {
    int tid, nthreads, rank;
    double *red;    // shared: [2][C][kMaxRed] cluster slots, then [2][32][kMaxRed] warp partials
    uint64_t *bar;  // shared: [3] mbarriers: one per slot buffer, one for reduce_wide (clusters only)
    GramShared<C> *gram = nullptr;
    double *scal = nullptr;  // shared: 8 doubles, scalars computed by warp 0 for the whole CTA (share_scalars)
    const SearchParams *trace_params = nullptr;  // set per replica by the segment (tuning aid)
    int trace_r = -1;
    __device__ __forceinline__ void mark(int id) const {
      if (trace_params) trace_mark(*trace_params, trace_r, tid, id);
    }
    int parity = 0;
    uint32_t phase = 0;  // bit p: parity the next wait on bar[p] looks for

    __device__ Team(double *smem, uint64_t *bars) : red(smem), bar(bars) {
      if constexpr (C > 1) {
        rank = (int)cg::this_cluster().block_rank();
      } else {
        rank = 0;
      }
      tid = rank * blockDim.x + threadIdx.x;
      nthreads = C * blockDim.x;
    }

    // Once per kernel, by every CTA of the cluster, before the first reduce.
    __device__ __forceinline__ void init() {
      if constexpr (C > 1) {
        if (threadIdx.x == 0) {
          mbar_init(bar, 1), mbar_init(bar + 1, 1), mbar_init(bar + 2, 1);
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cg::this_cluster().sync();
      }
    }
    // Before the CTAs of a cluster part ways: nobody's shared memory may go away while a peer's
    // posts to it could still be in flight.
    __device__ __forceinline__ void fini() {
      if constexpr (C > 1) cg::this_cluster().sync();
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
        uint64_t *mb = bar + parity;
        if (threadIdx.x < NV) {
          double x = 0;
          for (int k = 0; k < nw; k++) x += warp_part[k * kMaxRed + threadIdx.x];
          const uint32_t slot = smem_u32(slots + rank * kMaxRed + threadIdx.x), mbl = smem_u32(mb);
#pragma unroll
          for (int c = 0; c < C; c++) dsmem_post(map_to_cta(slot, c), x, map_to_cta(mbl, c));
        }
        if (threadIdx.x == 0) mbar_expect_tx(mb, C * NV * 8);
        mbar_wait(mb, (phase >> parity) & 1u);
        phase ^= 1u << parity;
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

    // Transcendental scalar algebra is done by ONE warp per CTA and broadcast through shared memory: sixteen warps
    // evaluating the same fp64 sin/cos/atan chains would queue on the SM's FP64 pipe for several microseconds.
    // f(out) fills out[0..n); every CTA of a team computes the same bits from the same inputs.
    template <int N, typename F>
    __device__ __forceinline__ void share_scalars(double (&out)[N], F &&f) {
      static_assert(N <= 8, "scal holds 8 doubles");
      if (threadIdx.x < 32) {
        f(out);
        if (threadIdx.x == 0) {
#pragma unroll
          for (int q = 0; q < N; q++) scal[q] = out[q];
        }
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < N; q++) out[q] = scal[q];
      __syncthreads();  // scal may be rewritten right away
    }

    // Sum nv <= kGramMax values over the team. In: gram->part[warp][v] written by lane 0 of every warp;
    // out: gram->out[v], readable by every thread on return. Single-buffered: at least one reduce<>() lies
    // between two calls, which is what keeps a fast CTA from posting into slots a peer still reads.
    __device__ __forceinline__ void reduce_wide(int nv) {
      GramShared<C> &G = *gram;
      const int nw = blockDim.x >> 5;
      __syncthreads();
      if constexpr (C == 1) {
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
          double x = 0;
          for (int k = 0; k < nw; k++) x += G.part[k][v];
          G.out[v] = x;
        }
      } else {
        uint64_t *mb = bar + 2;
        const uint32_t mbl = smem_u32(mb);
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
          double x = 0;
          for (int k = 0; k < nw; k++) x += G.part[k][v];
          const uint32_t slot = smem_u32(&G.slots[rank][v]);
#pragma unroll
          for (int c = 0; c < C; c++) dsmem_post(map_to_cta(slot, c), x, map_to_cta(mbl, c));
        }
        if (threadIdx.x == 0) mbar_expect_tx(mb, C * nv * 8);
        mbar_wait(mb, (phase >> 2) & 1u);
        phase ^= 4u;
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
          double x = 0;
#pragma unroll
          for (int c = 0; c < C; c++) x += G.slots[c][v];
          G.out[v] = x;
        }
      }
      __syncthreads();
    }
  };

  __host__ __device__ constexpr size_t team_smem_doubles(int C) { return 2 * (size_t)C * kMaxRed + 2 * 32 * kMaxRed; }

  // One strided pass over n elements by the whole team. The passes are latency-bound (a few dozen
  // elements per thread, each a round trip to L2/HBM), so all inputs of kPassUnroll elements are fetched
  // before the first is used. body(t, v) sees v[k] = src[k][t]; a thread visits its elements in
  // ascending t, so fused dot products accumulate in the same order whatever the unroll.
#ifndef OFLY_PASS_UNROLL
#define OFLY_PASS_UNROLL 4
#endif
  constexpr int kPassUnroll = OFLY_PASS_UNROLL;
  template <int NIN, typename F>
  __device__ __forceinline__ void team_pass(int tid, int nthreads, int n, const double *const (&src)[NIN], F &&body) {
    for (int t0 = tid; t0 < n; t0 += kPassUnroll * nthreads) {
      double v[kPassUnroll][NIN];
#pragma unroll
      for (int u = 0; u < kPassUnroll; u++) {
        const int t = t0 + u * nthreads;
        if (t < n) {
#pragma unroll
          for (int k = 0; k < NIN; k++) v[u][k] = src[k][t];
        }
      }
#pragma unroll
      for (int u = 0; u < kPassUnroll; u++) {
        const int t = t0 + u * nthreads;
        if (t < n) body(t, v[u]);
      }
    }
  }

  // L-BFGS history of one optimiser: vectors in the replica's workspace, scalars in RepState.
  struct Core {
    double *prev_x, *prev_g, *step, *s, *y;  // s, y: [n][n3p]
    CoreState *st;
    int n, k;
    size_t n3p;
  };

  // StepLBFGS::newton_step: core.step <- H * g (lbfgs_core.cpp:29-87). Also returns step.step and
  // step.w (w may be null) from the final pass. `k` is advanced in the local copy; the caller stores it.
  template <int C>
  __device__ __forceinline__ void newton_step(Team<C> &tm, Core &c, const double *x, const double *g, int n3, const double *w, double &out_nrm2,
                                           double &out_dotw) {
    const int n = c.n;
    const int prev = c.k > 0 ? (c.k - 1) % n : 0;
    const int incur = c.k <= n ? 0 : c.k - n;
    const int bound = c.k <= n ? c.k : n;
    double *q = c.step;

    if (c.k == 0) {  // identity Hessian: step = g
      double a[2] = {0, 0};
      const double *const src[3] = {g, x, w ? w : g};
      team_pass<3>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[3]) {
        c.prev_x[t] = v[1], c.prev_g[t] = v[0], q[t] = v[0];
        a[0] = fma(v[0], v[0], a[0]);
        if (w) a[1] = fma(v[0], v[2], a[1]);
      });
      tm.reduce<2>(a);
      out_nrm2 = a[0], out_dotw = a[1];
      c.k = 1;
      return;
    }

    double yy, d, rho_prev;
    {  // newest pair (s, y), prev <- current, q <- g, first dot of loop 1
      double *s = c.s + (size_t)prev * c.n3p, *y = c.y + (size_t)prev * c.n3p;
      double a[3] = {0, 0, 0};
      const double *const src[4] = {x, g, c.prev_x, c.prev_g};
      team_pass<4>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[4]) {
        double xv = v[0], gv = v[1];
        double sv = xv - v[2], yv = gv - v[3];
        s[t] = sv, y[t] = yv, c.prev_x[t] = xv, c.prev_g[t] = gv, q[t] = gv;
        a[0] = fma(sv, yv, a[0]), a[1] = fma(yv, yv, a[1]), a[2] = fma(sv, gv, a[2]);
      });
      tm.reduce<3>(a);
      rho_prev = 1.0 / fabs(a[0]);  // abs: never an ascent direction (lbfgs_core.cpp:50)
      if (tm.tid == 0) c.st->rho[prev] = rho_prev;
      yy = a[1], d = a[2];
    }
    const double h0 = 1.0 / rho_prev / yy;  // initial Hessian scaling (lbfgs_core.cpp:70)

    double alpha_of[kMaxHist];
    // loop 1, newest -> oldest: alpha_j = rho_j s_j.q ; q -= alpha_j y_j   (dot for the next pair fused)
    for (int i = bound - 1; i >= 0; --i) {
      const int j = (i + incur) % n;
      const double rho_j = j == prev ? rho_prev : c.st->rho[j];
      const double alpha = rho_j * d;
      alpha_of[j] = alpha;
      const double *y = c.y + (size_t)j * c.n3p;
      double a[1] = {0};
      if (i > 0) {
        const double *sn = c.s + (size_t)((i - 1 + incur) % n) * c.n3p;
        const double *const src[3] = {y, q, sn};
        team_pass<3>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[3]) {
          double qv = fma(-alpha, v[0], v[1]);
          q[t] = qv;
          a[0] = fma(v[2], qv, a[0]);
        });
      } else {  // last pair: apply H0 and start loop 2 with the same y
        const double *const src[2] = {y, q};
        team_pass<2>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[2]) {
          double rv = fma(-alpha, v[0], v[1]) * h0;
          q[t] = rv;
          a[0] = fma(v[0], rv, a[0]);
        });
      }
      tm.reduce<1>(a);
      d = a[0];
    }
    // loop 2, oldest -> newest: beta = rho_j y_j.r ; r += (alpha_j - beta) s_j
    for (int i = 0; i < bound; ++i) {
      const int j = (i + incur) % n;
      const double rho_j = j == prev ? rho_prev : c.st->rho[j];
      const double coef = alpha_of[j] - rho_j * d;
      const double *s = c.s + (size_t)j * c.n3p;
      if (i + 1 < bound) {
        const double *yn = c.y + (size_t)((i + 1 + incur) % n) * c.n3p;
        double a[1] = {0};
        const double *const src[3] = {s, q, yn};
        team_pass<3>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[3]) {
          double rv = fma(coef, v[0], v[1]);
          q[t] = rv;
          a[0] = fma(v[2], rv, a[0]);
        });
        tm.reduce<1>(a);
        d = a[0];
      } else {
        double a[2] = {0, 0};
        const double *const src[3] = {s, q, w ? w : q};
        team_pass<3>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[3]) {
          double rv = fma(coef, v[0], v[1]);
          q[t] = rv;
          a[0] = fma(rv, rv, a[0]);
          if (w) a[1] = fma(rv, v[2], a[1]);
        });
        tm.reduce<2>(a);
        out_nrm2 = a[0], out_dotw = a[1];
      }
    }
    c.k++;
  }

  // The rotor's torque as an on-the-fly gradient for newton_step_gram: g = (g1 - g0) - pax x - mean[component].
  struct TorqueGen {
    const double *g1, *g0;
    double pax, mean[3];
  };

  // Pass B of newton_step_gram for up to MAXB stored pairs. MAXB = 8 (the rotor's n = 6): the y and s rows of an
  // element are all requested together, one round trip per element; MAXB = 16: y rows, then s rows (registers).
  template <int C, int MAXB>
  __device__ __forceinline__ void gram_combine(const GramShared<C> &G, const Core &c, const double *g, const double *w, int n3, int tid, int nth,
                                               int bound, double h0, double (&a)[2]) {
    double *out = c.step;
    for (int t = tid; t < n3; t += nth) {
      double h[MAXB], hs[MAXB <= 8 ? MAXB : 1];
#pragma unroll
      for (int i = 0; i < MAXB; i++) {
        if (i < bound) {
          h[i] = c.y[G.off[i] + t];
          if constexpr (MAXB <= 8) hs[i] = c.s[G.off[i] + t];
        }
      }
      double q = g[t];
      const double wv = w ? w[t] : 0.0;
#pragma unroll
      for (int i = MAXB - 1; i >= 0; --i) {
        if (i < bound) q = fma(-G.alpha[i], h[i], q);
      }
      q *= h0;
      if constexpr (MAXB > 8) {
#pragma unroll
        for (int i = 0; i < MAXB; i++) {
          if (i < bound) h[i] = c.s[G.off[i] + t];
        }
      }
#pragma unroll
      for (int i = 0; i < MAXB; i++) {
        if (i < bound) {
          if constexpr (MAXB <= 8) {
            q = fma(G.w[i], hs[i], q);
          } else {
            q = fma(G.w[i], h[i], q);
          }
        }
      }
      out[t] = q;
      a[0] = fma(q, q, a[0]);
      if (w) a[1] = fma(q, wv, a[1]);
    }
  }

  // StepLBFGS::newton_step in TWO streaming passes instead of one pass per history pair.
  //
  // The two-loop recursion only ever needs the running vector q through its dot products with the stored
  // pairs, and q stays in span{g, y_k, s_k}. With the Gram entries s_j.y_k and y_j.y_k of the stored pairs
  // (kept in CoreState, one row/column refreshed per step) plus s_k.g and y_k.g, every alpha_j / beta_j of
  // lbfgs_core.cpp:57-82 is scalar arithmetic:
  //   s_j.q = s_j.g - sum_{k newer than j} alpha_k s_j.y_k
  //   y_j.r = h0 (y_j.g - sum_k alpha_k y_j.y_k) + sum_{k older than j} w_k s_k.y_j ,  w_k = alpha_k - beta_k
  // Pass A streams the pairs once and accumulates the 5 dot products per pair that involve the newest pair
  // or g; ONE team-wide reduction follows; thread 0 of each CTA runs the O(n^2) scalar recursion; pass B
  // forms step = h0 (g - sum alpha_k y_k) + sum w_k s_k element by element in the two-loop's own operation
  // order. The dependent chain per step is 2 passes + 2 reductions instead of 2n+2 of each -- the search is
  // bound by exactly that chain once few replicas are left -- for the same bytes streamed.
  // Differs from the two-loop only in how alpha/beta are rounded (scripts/lbfgs_gram_study.py).
  template <int C>
  __device__ __forceinline__ void newton_step_gram(Team<C> &tm, Core &c, const double *x, double *g, int n3, const double *w,
                                                double &out_nrm2, double &out_dotw, const TorqueGen *tg = nullptr) {
    if (c.k == 0) return newton_step(tm, c, x, g, n3, w, out_nrm2, out_dotw);  // identity Hessian: one pass

    constexpr int GS = 3;  // pairs per sweep of pass A (5 accumulators each)
    constexpr int CH = 2;  // elements whose loads are issued back to back
    GramShared<C> &G = *tm.gram;
    const int n = c.n;
    const int prev = (c.k - 1) % n;
    const int incur = c.k <= n ? 0 : c.k - n;
    const int bound = c.k <= n ? c.k : n;
    const int tid = tm.tid, nth = tm.nthreads;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *sp = c.s + (size_t)prev * c.n3p, *yp = c.y + (size_t)prev * c.n3p;

    // working copies of the persistent scalars (consumed after the reduction's barriers)
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) G.sy[t] = c.st->sy[t], G.yy[t] = c.st->yy[t];
    if (threadIdx.x < n) G.rho[threadIdx.x] = c.st->rho[threadIdx.x];

    if (tg == nullptr) {  // newest pair (s, y); prev <- current   (lbfgs_core.cpp:37-52)
      const double *const src[4] = {x, g, c.prev_x, c.prev_g};
      team_pass<4>(tid, nth, n3, src, [&](int t, const double(&v)[4]) {
        sp[t] = v[0] - v[2], yp[t] = v[1] - v[3], c.prev_x[t] = v[0], c.prev_g[t] = v[1];
      });
    } else {  // the same with g = the rotor's torque formed on the fly from g1, g0 and x = axis (rotor.cpp:89-94)
      const double pax = tg->pax, m0 = tg->mean[0], m1 = tg->mean[1], m2 = tg->mean[2];
      const double *const src[5] = {x, tg->g1, tg->g0, c.prev_x, c.prev_g};
      team_pass<5>(tid, nth, n3, src, [&](int t, const double(&v)[5]) {
        const int cmp = t % 3;
        const double gv = (v[1] - v[2]) - pax * v[0] - (cmp == 0 ? m0 : (cmp == 1 ? m1 : m2));
        g[t] = gv;
        sp[t] = v[0] - v[3], yp[t] = gv - v[4], c.prev_x[t] = v[0], c.prev_g[t] = gv;
      });
    }

    tm.mark(20);
    // pass A: every thread re-reads only its own elements of the newest pair
    for (int i0 = 0; i0 < bound; i0 += GS) {
      const double *sk[GS], *yk[GS];
#pragma unroll
      for (int q = 0; q < GS; q++) {
        const int i = min(i0 + q, bound - 1);  // a short last sweep repeats its last pair; those sums are dropped
        const size_t row = (size_t)((i + incur) % n) * c.n3p;
        sk[q] = c.s + row, yk[q] = c.y + row;
      }
      double acc[GS][kGramVals];
#pragma unroll
      for (int q = 0; q < GS; q++) {
#pragma unroll
        for (int d = 0; d < kGramVals; d++) acc[q][d] = 0;
      }
      for (int t0 = tid; t0 < n3; t0 += CH * nth) {
        double vs[CH], vy[CH], vg[CH], hs[CH][GS], hy[CH][GS];
#pragma unroll
        for (int u = 0; u < CH; u++) {
          const int t = t0 + u * nth;
          if (t < n3) {
            vs[u] = sp[t], vy[u] = yp[t], vg[u] = g[t];
#pragma unroll
            for (int q = 0; q < GS; q++) hs[u][q] = sk[q][t], hy[u][q] = yk[q][t];
          }
        }
#pragma unroll
        for (int u = 0; u < CH; u++) {
          const int t = t0 + u * nth;
          if (t < n3) {
#pragma unroll
            for (int q = 0; q < GS; q++) {
              acc[q][0] = fma(vs[u], hy[u][q], acc[q][0]);
              acc[q][1] = fma(vy[u], hs[u][q], acc[q][1]);
              acc[q][2] = fma(vy[u], hy[u][q], acc[q][2]);
              acc[q][3] = fma(vg[u], hs[u][q], acc[q][3]);
              acc[q][4] = fma(vg[u], hy[u][q], acc[q][4]);
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < GS; q++) {
#pragma unroll
        for (int d = 0; d < kGramVals; d++) {
          double v = acc[q][d];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0 && i0 + q < bound) G.part[wid][kGramVals * (i0 + q) + d] = v;
        }
      }
    }
    tm.mark(21);
    tm.reduce_wide(kGramVals * bound);
    tm.mark(22);

    // Scalar recursion by warp 0 (identical inputs, identical arithmetic on every CTA of the team). Lane i owns
    // pair i (order index, oldest first) with its Gram rows in registers; both triangular sweeps broadcast one
    // coefficient per step, so the dependent chain is ~n shuffles + FMAs instead of ~n^2 shared-memory round trips.
    // Each lane accumulates in the order the serial recursion would.
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      const int p = prev;
      const bool mine = lane < bound;
      const int j = mine ? (lane + incur) % n : 0;
      const double *v = G.out + kGramVals * (mine ? lane : 0);
      if (mine) {  // refresh row / column p with this step's sums (entry (p, p) is written twice with the same bits)
        G.sy[p * n + j] = v[0], G.sy[j * n + p] = v[1], G.yy[p * n + j] = v[2], G.yy[j * n + p] = v[2];
        G.off[lane] = (unsigned long long)j * c.n3p;
        G.slot[lane] = j;
      }
      __syncwarp();
      const double rho_p = 1.0 / fabs(G.sy[p * n + p]);  // abs: never an ascent direction (lbfgs_core.cpp:50)
      const double h0 = 1.0 / rho_p / G.yy[p * n + p];   // initial Hessian scaling (lbfgs_core.cpp:70)
      const double rho_i = !mine ? 0.0 : (j == p ? rho_p : G.rho[j]);
      const double *sy_row = G.sy + j * n, *yy_row = G.yy + j * n;  // s_i.y_k and y_i.y_k at column slot(k)
      // loop 1, newest -> oldest: alpha_i = rho_i (s_i.g - sum_{k newer} alpha_k s_i.y_k)
      double sq = v[3], alpha_i = 0.0;
      double yq = v[4];  // y_i.g - sum_k alpha_k y_i.y_k, k newest -> oldest (enters loop 2)
#pragma unroll 1
      for (int k = bound - 1; k >= 0; --k) {
        const int jk = G.slot[k];
        const double s_iyk = sy_row[jk], y_iyk = yy_row[jk];
        const double alpha_k = __shfl_sync(0xffffffffu, rho_i * sq, k);
        if (lane == k) alpha_i = alpha_k;
        if (lane < k) sq = fma(-alpha_k, s_iyk, sq);
        yq = fma(-alpha_k, y_iyk, yq);
      }
      // loop 2, oldest -> newest: w_i = alpha_i - rho_i (h0 yq_i + sum_{k older} w_k s_k.y_i)
      double yr = yq * h0, w_i = 0.0;
#pragma unroll 1
      for (int k = 0; k < bound; k++) {
        const double s_kyi = G.sy[G.slot[k] * n + j];
        const double w_k = __shfl_sync(0xffffffffu, alpha_i - rho_i * yr, k);
        if (lane == k) w_i = w_k;
        if (lane > k) yr = fma(w_k, s_kyi, yr);
      }
      if (mine) G.alpha[lane] = alpha_i, G.w[lane] = w_i;
      if (lane == 0) G.h0 = h0;
      if (tm.rank == 0) {  // persist the refreshed row / column
        if (lane == 0) c.st->rho[p] = rho_p;
        if (mine) {
          c.st->sy[p * n + j] = v[0], c.st->sy[j * n + p] = v[1];
          c.st->yy[p * n + j] = v[2], c.st->yy[j * n + p] = v[2];
        }
      }
    }
    __syncthreads();
    tm.mark(23);

    // pass B: step = (g - sum alpha_k y_k) h0 + sum w_k s_k, accumulated newest->oldest then oldest->newest
    const double h0 = G.h0;
    double a[2] = {0, 0};
    if (bound <= 8) {
      gram_combine<C, 8>(G, c, g, w, n3, tid, nth, bound, h0, a);
    } else {
      gram_combine<C, kMaxHist>(G, c, g, w, n3, tid, nth, bound, h0, a);
    }
    tm.mark(24);
    tm.reduce<2>(a);
    tm.mark(25);
    out_nrm2 = a[0], out_dotw = a[1];
    c.k++;
  }

  struct Work {  // per-replica vectors (each n3p doubles)
    double *base;
    double *x, *g;  // positions; gradient (minimise) or effective gradient (dimer)
    double *axis, *g0, *g1, *g1p, *dprev, *axisp, *dg;
    Core trans, rot;
  };

  __device__ __forceinline__ void carve_core(Core &c, double *&p, int n, size_t n3p, CoreState *st) {
    c.n = n, c.k = st->k, c.n3p = n3p, c.st = st;
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

  __device__ __forceinline__ double next_trust(const SearchParams &P, double trust, double proj) {
    if (proj < -P.proj_tol) return fmax(P.min_trust, P.shrink_trust * trust);
    if (proj > P.proj_tol) return fmin(P.max_trust, P.grow_trust * trust);
    return trust;
  }

  // One replica's segment: everything between two gradient requests. All threads of the team run the
  // same control flow on identical scalars; thread 0 stores the state that must survive the tick.
  template <int C>
  struct Segment {
    Team<C> &tm;
    const BatchDev &B;
    const SearchParams &P;
    const int r, n3;
    int entry = 0;  // position of this replica in the active list (where its fate is reported)
    size_t q;       // search of the queue this slot is running
    bool done = false;  // the search finished in this segment (the slot may take the next one)
    RepState *S;    // global
    Work W;
    // working copies of the persistent scalars
    int iter, rot_it, first, need_rebuild;
    double trust, acc, convex_count, curv, mag_g, c_x0, b_1, theta_1;
    long tally[4];

    __device__ Segment(Team<C> &tm_, const BatchDev &B_, const SearchParams &P_, int r_) : tm(tm_), B(B_), P(P_), r(r_), n3(3 * B_.N) {
      S = P.state + r;
      q = (size_t)S->qid;
      const size_t n3p = ((size_t)n3 + 3) & ~(size_t)3;
      double *p = P.ws + (size_t)r * P.ws_stride;
      W.base = p;
      W.x = p, p += n3p;
      W.g = p, p += n3p;
      carve_core(W.trans, p, P.n_hist, n3p, &S->trans);
      if (P.mode == 1) {
        W.axis = p, p += n3p;
        W.g0 = p, p += n3p;
        W.g1 = p, p += n3p;
        W.g1p = p, p += n3p;
        W.dprev = p, p += n3p;
        W.axisp = p, p += n3p;
        W.dg = p, p += n3p;
        carve_core(W.rot, p, P.rot_n, n3p, &S->rot);
      }
      iter = S->iter, rot_it = S->rot_it, first = S->first, need_rebuild = 0;
      trust = S->trust, acc = S->acc, convex_count = S->convex_count, curv = S->curv, mag_g = S->mag_g;
      c_x0 = S->c_x0, b_1 = S->b_1, theta_1 = S->theta_1;
      for (int k = 0; k < 4; k++) tally[k] = S->tally[k];
    }

    __device__ double4 *pos() const { return B.pos4 + (size_t)r * B.N; }
    // Component t = 3 i + c of a [3N] vector lives at double index 4 i + c of the replica's pos4 records.
    // Every pass, including the list-position updates, is indexed by t, so a thread only ever re-reads
    // what it wrote itself: inside a segment no global data crosses threads.
    static __device__ __forceinline__ int pos_slot(int t) { return t + t / 3; }

    __device__ __forceinline__ void lbfgs_step(Core &c, const double *x, double *g, const double *w, double &nrm2, double &dotw,
                                               const TorqueGen *tg = nullptr) {
      if (P.gram) return newton_step_gram<C>(tm, c, x, g, n3, w, nrm2, dotw, tg);
      newton_step(tm, c, x, g, n3, w, nrm2, dotw);
    }

    // Suspend until the gradient of the current list positions is in `dst`; resume at `phase`.
    __device__ __forceinline__ void yield(int phase, double *dst) {
      if (tm.tid == 0) {
        S->phase = phase, S->iter = iter, S->rot_it = rot_it, S->first = first, S->need_rebuild = need_rebuild;
        P.flags[entry] = 1 | (need_rebuild ? 2 : 0);
        S->gdst = (unsigned long long)(dst - W.base);
        S->trust = trust, S->acc = acc, S->convex_count = convex_count, S->curv = curv, S->mag_g = mag_g;
        S->c_x0 = c_x0, S->b_1 = b_1, S->theta_1 = theta_1;
        S->trans.k = W.trans.k;
        if (P.mode == 1) S->rot.k = W.rot.k;
        for (int k = 0; k < 4; k++) S->tally[k] = tally[k];
      }
    }

    // Write the results of a finished search.
    __device__ __forceinline__ void finish(int status) {
      __syncthreads();
      for (int t = tm.tid; t < n3; t += tm.nthreads) {
        P.pos_out[q * n3 + t] = W.x[t];
        if (P.mode == 1) P.axis_out[q * n3 + t] = W.axis[t];
        if (P.mode == 0 && P.grad_out) P.grad_out[q * n3 + t] = W.g[t];
      }
      if (tm.tid == 0) {
        P.status[q] = status;
        if (P.counters) {
          long *c = P.counters + 4 * q;
          for (int k = 0; k < 4; k++) c[k] = tally[k];
        }
        S->phase = kDone, S->status = status, S->need_rebuild = 0;
        P.flags[entry] = 0;
        if (status == kStatusCapacity) atomicOr(&P.ctl->fault_or, kFaultCapacity);
      }
      done = true;
    }

    // Continuous refill (saddle::Master::find_n's `while (tot < max_searches)` loop, src/saddle/find.cpp:374-422, without
    // its batch boundary): a slot whose search has finished claims the next start state of the queue and begins it in
    // the same segment, so the chip stays full until the queue is empty. Which slot runs which search depends on
    // timing; what a search computes does not (every sum is independent of the batch composition), and results are
    // stored by search index. Returns false when the queue is exhausted (the slot then leaves the active list).
    __device__ __forceinline__ bool take_next_search() {
      double v[1] = {0.0};
      if (tm.tid == 0) v[0] = (double)atomicAdd(&P.ctl->queue_head, 1);
      tm.reduce<1>(v);  // every thread of the team learns the claimed index
      const long next = (long)v[0];
      if (next >= P.n_queue) return false;
      q = (size_t)next;
      done = false;
      if (tm.tid == 0) S->qid = (int)next, B.fault[r] = 0;
      return true;
    }

    // neigh::List::rebuild(x) request: the grid kernels of this tick do the work; the histogram of
    // the replica must be zero before they start.
    __device__ __forceinline__ void request_rebuild() {
      int *hist = B.cell_count + (size_t)r * B.grid.ncells;
      for (int c = tm.tid; c < B.grid.ncells; c += tm.nthreads) hist[c] = 0;
      need_rebuild = 1;
      tally[1]++;
    }

    // Trust-limited step shared by the minimiser and the dimer (lbfgs.cpp:83-100, dimer.cpp:133-151):
    // Hg *= min(1, trust/|Hg|); acc += |Hg|; x -= Hg; then rebuild if acc > skin/2 else update(Hg).
    __device__ __forceinline__ void take_step(double nrm2) {
      const double nrm = sqrt(nrm2);
      const double scale = fmin(1.0, trust / nrm);
      acc += scale * nrm;
      const bool rebuild = acc > 0.5 * P.skin;
      double *step = W.trans.step;
      double *p4 = reinterpret_cast<double *>(pos());
      {
        const double *const src[2] = {step, W.x};
        team_pass<2>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[2]) {
          const double h = v[0] * scale;
          step[t] = h;
          W.x[t] = v[1] - h;
          if (!rebuild) p4[pos_slot(t)] -= h;
        });
      }
      if (rebuild) {
        request_rebuild();
        acc = 0;
      } else {
        tally[2]++;
      }
    }

    // ---- minimise::LBFGS::minimise ----------------------------------------------------------------
    __device__ __forceinline__ void start_minimise() {
      for (int t = tm.tid; t < n3; t += tm.nthreads) W.x[t] = P.pos_in[q * n3 + t];
      W.trans.k = 0;
      iter = 0, first = 1, trust = P.min_trust, acc = 0;
      for (int k = 0; k < 4; k++) tally[k] = 0;
      request_rebuild();
      yield(kMinGrad, W.g);
    }
    __device__ __forceinline__ void run_minimise(int phase) {
      if (phase == kInit) return start_minimise();
      // phase == kMinGrad: the gradient at x is in W.g
      tally[0]++;
      double b2[2] = {0, 0};
      const double *step = W.trans.step;
      {
        const double *const src[2] = {W.g, step};
        const bool with_step = !first;
        team_pass<2>(tm.tid, tm.nthreads, n3, src, [&](int, const double(&v)[2]) {
          b2[0] = fma(v[0], v[0], b2[0]);
          if (with_step) b2[1] = fma(v[0], v[1], b2[1]);
        });
      }
      tm.reduce<2>(b2);
      mag_g = b2[0];
      if (!first) {
        tally[3]++;
        trust = next_trust(P, trust, b2[1]);
        iter++;
      }
      first = 0;
      if (iter >= P.iter_max) return finish(1);                 // iter_max
      if (mag_g < P.f2norm * P.f2norm) return finish(0);        // converged
      double nrm2, unused;
      lbfgs_step(W.trans, W.x, W.g, (const double *)nullptr, nrm2, unused);
      take_step(nrm2);
      yield(kMinGrad, W.g);
    }

    // ---- saddle::Rotor::eff_gradient + saddle::Dimer::find_sp ---------------------------------------
    // Returns true if the replica finished or suspended; false never.
    __device__ __forceinline__ void start_dimer() {
      for (int t = tm.tid; t < n3; t += tm.nthreads) {
        W.x[t] = P.pos_in[q * n3 + t];
        W.axis[t] = P.axis_in[q * n3 + t];
      }
      W.trans.k = 0, W.rot.k = 0;
      iter = 0, rot_it = 0, first = 1, trust = P.min_trust, acc = 0, convex_count = 0;
      for (int k = 0; k < 4; k++) tally[k] = 0;
      request_rebuild();
      yield(kRotG0, W.g0);  // eff_gradient starts with pot.gradient(g0) (rotor.cpp:75)
    }
    __device__ __forceinline__ void run_dimer(int phase) {
      const double dr = P.delta_r;
      const double inv_n = 1.0 / (double)B.N;
      double *p4 = reinterpret_cast<double *>(pos());

      if (phase == kInit) return start_dimer();

      bool rot_head = false, have_torque = false;
      double a7[7] = {0, 0, 0, 0, 0, 0, 0};  // (g1-g0).axis, sum(g1-g0)[3], sum(axis)[3]
      if (phase == kRotG0) {
        // delta = -dr * axis ; nl.update(delta) ; delta_prev = delta ; gradient(g1)   (rotor.cpp:77-85)
        tally[0]++;
        W.rot.k = 0;  // m_core.clear()
        for (int t = tm.tid; t < n3; t += tm.nthreads) {
          const double d = -dr * W.axis[t];
          W.dprev[t] = d;
          p4[pos_slot(t)] -= d;
        }
        tally[2]++;
        return yield(kRotG1, W.g1);
      }
      if (phase == kRotG1) {
        tally[0]++;
        rot_it = 0;
        rot_head = true;
      }
      if (phase == kRotG1p) {
        // the trial gradient g1p at axis' arrived (rotor.cpp:118-134)
        tally[0]++;
        const double *theta = W.rot.step;
        double a1[1] = {0};
        {
          const double *const src[3] = {W.g1p, W.g0, W.axisp};
          team_pass<3>(tm.tid, tm.nthreads, n3, src, [&](int, const double(&v)[3]) { a1[0] = fma(v[0] - v[1], v[2], a1[0]); });
        }
        tm.reduce<1>(a1);
        double sc[6];  // theta_min, cos, sin, k1, k2, k3
        {
          const double th1 = theta_1, cx0 = c_x0, b1 = b_1, a1v = a1[0];
          tm.share_scalars(sc, [&](double(&o)[6]) {
            const double s1 = sin(th1);
            const double c_x1 = a1v / dr;
            const double a_1 = (cx0 - c_x1 + b1 * sin(2 * th1)) / (1 - cos(2 * th1));
            double tmin = 0.5 * atan(b1 / a_1);
            if (a_1 * cos(2 * tmin) - a_1 + b1 * sin(2 * tmin) > 0) tmin += M_PI / 2;  // extremum was a maximum
            const double cmv = cos(tmin), smv = sin(tmin);
            o[0] = tmin, o[1] = cmv, o[2] = smv;
            o[3] = sin(th1 - tmin) / s1, o[4] = smv / s1, o[5] = 1 - cmv - smv * tan(0.5 * th1);
          });
        }
        const double theta_min = sc[0], cm = sc[1], sm = sc[2], k1 = sc[3], k2 = sc[4], k3 = sc[5];
        // The same pass accumulates what the next rotation iteration's torque needs ((g1-g0).axis and the component
        // sums of g1-g0 and axis), element by element exactly as the separate pass below would.
        {
          const double *const src[5] = {W.axis, theta, W.g0, W.g1, W.g1p};
          team_pass<5>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[5]) {
            double an = v[0] * cm + v[1] * sm;
            W.axis[t] = an;
            double g0v = v[2];
            double gn = k1 * v[3] + k2 * v[4] + k3 * g0v;  // interpolated end-point gradient (rotor.cpp:132-134)
            W.g1[t] = gn;
            double dv = gn - g0v;
            a7[0] = fma(dv, an, a7[0]);
            int c = t % 3;
            a7[1 + c] += dv, a7[4 + c] += an;
          });
        }
        tm.reduce<7>(a7);
        if (fabs(theta_min) < P.rot_theta_tol) {
          curv = a7[0] / dr;
        } else {
          rot_it++;
          rot_head = true;
          have_torque = true;
        }
      }

      tm.mark(10);
      if (rot_head) {
        // torque: dg = (g1 - g0) - ((g1 - g0).axis) axis, minus its mean if nothing is frozen (rotor.cpp:89-94)
        if (!have_torque) {
          const double *const src[3] = {W.g1, W.g0, W.axis};
          team_pass<3>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[3]) {
            double dv = v[0] - v[1], av = v[2];
            a7[0] = fma(dv, av, a7[0]);
            int c = t % 3;
            a7[1 + c] += dv, a7[4 + c] += av;
          });
          tm.reduce<7>(a7);
        }
        const double pax = a7[0];
        TorqueGen tg{W.g1, W.g0, pax, {0, 0, 0}};
        if (P.count_frozen == 0) {
          for (int c = 0; c < 3; c++) tg.mean[c] = (a7[1 + c] - pax * a7[4 + c]) * inv_n;
        }
        // With stored pairs the single-pass step forms dg inside its first pass; otherwise it is written here.
        const bool fused = P.gram && W.rot.k > 0;
        if (!fused) {
          const double *const src[3] = {W.g1, W.g0, W.axis};
          team_pass<3>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[3]) { W.dg[t] = (v[0] - v[1]) - pax * v[2] - tg.mean[t % 3]; });
          __syncthreads();
        }

        tm.mark(11);
        double nrm2_unused, th_ax;
        lbfgs_step(W.rot, W.axis, W.dg, W.axis, nrm2_unused, th_ax, fused ? &tg : nullptr);
        tm.mark(12);
        double *theta = W.rot.step;
        // theta -= (theta.axis) axis ; normalise ; b_1 = (g1-g0).theta / dr   (rotor.cpp:99-103)
        double a2[2] = {0, 0};
        {
          const double *const src[4] = {W.axis, theta, W.g1, W.g0};
          team_pass<4>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[4]) {
            double tv = fma(-th_ax, v[0], v[1]);
            theta[t] = tv;
            a2[0] = fma(tv, tv, a2[0]);
            a2[1] = fma(v[2] - v[3], tv, a2[1]);
          });
        }
        tm.reduce<2>(a2);
        const double inv_th = 1 / sqrt(a2[0]);
        b_1 = a2[1] * inv_th / dr;
        c_x0 = pax / dr;
        double st[3];  // theta_1, cos, sin
        {
          const double b1 = b_1, cx0 = c_x0;
          tm.share_scalars(st, [&](double(&o)[3]) {
            o[0] = -0.5 * atan(b1 / fabs(cx0));
            o[1] = cos(o[0]), o[2] = sin(o[0]);
          });
        }
        theta_1 = st[0];

        if (fabs(theta_1) < P.rot_theta_tol || (rot_it >= P.iter_max_rot && c_x0 < 0) || rot_it > 10 * P.iter_max_rot) {
          curv = c_x0;
        } else {
          // trial axis, move the list end-point from the old offset to the new one (rotor.cpp:111-116)
          const double c1 = st[1], s1 = st[2];
          {
            const double *const src[3] = {theta, W.axis, W.dprev};
            team_pass<3>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[3]) {
              const double tn = v[0] * inv_th;
              theta[t] = tn;
              const double ap = v[1] * c1 + tn * s1;
              W.axisp[t] = ap;
              const double nw = -dr * ap;
              W.dprev[t] = nw;
              p4[pos_slot(t)] -= nw - v[2];
            });
          }
          tally[2]++;
          tm.mark(13);
          return yield(kRotG1p, W.g1p);
        }
      }

      tm.mark(14);
      // ---- rotation finished: undo the end-point offset (rotor.cpp:150-151), effective gradient (:153-161)
      {
        double a7[7] = {0, 0, 0, 0, 0, 0, 0};
        {
          const double *const src[2] = {W.g0, W.axis};
          team_pass<2>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[2]) {
            a7[0] = fma(v[0], v[1], a7[0]);
            int c = t % 3;
            a7[1 + c] += v[0], a7[4 + c] += v[1];
          });
        }
        tm.reduce<7>(a7);
        const bool parallel_only = !P.relax_in_convex && curv > 0;
        const double p0 = a7[0];
        const double ca = parallel_only ? 0.0 : 1.0, cb = parallel_only ? -p0 : -2 * p0;  // eff = ca g0 + cb axis
        double mean[3] = {0, 0, 0};
        if (P.count_frozen == 0) {
          for (int c = 0; c < 3; c++) mean[c] = (ca * a7[1 + c] + cb * a7[4 + c]) * inv_n;
        }
        const double *step = W.trans.step;
        double a2[2] = {0, 0};
        {
          const double *const src[4] = {W.g0, W.axis, W.dprev, step};
          const bool with_step = !first;
          team_pass<4>(tm.tid, tm.nthreads, n3, src, [&](int t, const double(&v)[4]) {
            p4[pos_slot(t)] += v[2];
            const double e = ca * v[0] + cb * v[1] - mean[t % 3];
            W.g[t] = e;
            a2[0] = fma(e, e, a2[0]);
            if (with_step) a2[1] = fma(e, v[3], a2[1]);
          });
        }
        tally[2]++;
        tm.reduce<2>(a2);
        mag_g = a2[0];
        if (!first) {  // tail of the translation loop body (dimer.cpp:153-167)
          tally[3]++;
          convex_count = curv > 0 ? convex_count + 1 : 0;
          trust = next_trust(P, trust, a2[1]);
          iter++;
        }
        first = 0;
      }

      // ---- head of the translation loop (dimer.cpp:73-151)
      if (iter >= P.iter_max) return finish(-2);                   // iter_max
      if (mag_g < P.f2norm * P.f2norm) return finish(0);           // success
      if (convex_count >= P.convex_max) return finish(-1);         // convex
      if (P.use_history && iter % P.hist_check_freq == 0) {        // dimer.cpp:102-131
        double ctheta = -1;
        for (long k = 0; k < P.n_sp; k++) {
          const double *sp = P.hist_sp + (size_t)k * n3;
          double a3[3] = {0, 0, 0};
          const double *const src[3] = {P.in_min, W.x, sp};
          team_pass<3>(tm.tid, tm.nthreads, n3, src, [&](int, const double(&v)[3]) {
            double a = v[1] - v[0], b = v[2] - v[0];
            a3[0] = fma(a, b, a3[0]), a3[1] = fma(b, b, a3[1]), a3[2] = fma(a, a, a3[2]);
          });
          tm.reduce<3>(a3);
          ctheta = fmax(ctheta, a3[0] / (sqrt(a3[1]) * sqrt(a3[2])));
          if (ctheta > P.cos_theta_tol) return finish(-3);  // collision
        }
      }
      tm.mark(15);
      double nrm2, unused;
      lbfgs_step(W.trans, W.x, W.g, (const double *)nullptr, nrm2, unused);
      tm.mark(16);
      take_step(nrm2);
      tm.mark(17);
      yield(kRotG0, W.g0);
    }
  };

  // One tick's vector algebra: gridDim.x = R * C, team t owns replica t.
  // MAXT = 512: large systems, up to 128 registers so the batched loads of team_pass stay in registers;
  // MAXT = 128: small systems with thousands of replicas, 8 teams resident per SM.
  template <int C, int MAXT>
  __global__ void __launch_bounds__(MAXT, MAXT == 128 ? 4 : 1) k_segment(BatchDev B, SearchParams P) {
    __shared__ double red_smem[team_smem_doubles(C)];
    __shared__ uint64_t red_bar[3];
    __shared__ GramShared<C> gram_smem;
    __shared__ double scal_smem[8];
    __shared__ int last_team;

    Team<C> tm(red_smem, red_bar);
    tm.gram = &gram_smem;
    tm.scal = scal_smem;
    // Team t owns entry t of the list of replicas that were active after the previous tick (tick 0: all); the
    // teams beyond the list leave at once. (A loop over entries inside the kernel would let the compiler hoist
    // every loop-invariant address out of it and spill them all.)
    const int n_listed = P.ctl->n_active;
    const int a = blockIdx.x / C;
    if (a >= n_listed) return;  // whole clusters leave together
    {
      // One thread reads the slot's header and shares it: the segment overwrites phase / iter before it ends, and a warp
      // that is late to this point must not see the new values (every CTA of a cluster reads before the cluster barrier
      // of Team::init, so no peer can be ahead either).
      __shared__ int head[4];
      if (threadIdx.x == 0) {
        const int r0 = P.active[a];
        head[0] = r0, head[1] = P.state[r0].phase, head[2] = P.state[r0].iter, head[3] = __ldcg(B.fault + r0);
      }
      __syncthreads();
      const int r = head[0];
      const int phase = head[1];
      if (phase != kDone) {
        tm.init();
        const int fault = head[3];
        tm.trace_params = &P, tm.trace_r = r;
        tm.mark(1);
        Segment<C> seg(tm, B, P, r);
        seg.entry = a;
        tm.mark(2);
        if (phase != kInit && (fault & kFaultCapacity)) {
          seg.finish(kStatusCapacity);
        } else if (phase != kInit && (fault & kFaultOutOfBounds)) {
          seg.finish(kStatusOutOfBounds);  // the reference throws Spline::OutOfBounds out of find_sp / minimise
        } else if (head[2] < 0) {
          seg.finish(kStatusUnfinished);
        } else {
          if (phase == kInit && tm.tid == 0) B.fault[r] = 0;
          if (P.mode == 0) {
            seg.run_minimise(phase);
          } else {
            seg.run_dimer(phase);
          }
        }
        if (seg.done && head[2] >= 0 && seg.take_next_search()) {  // (out of ticks: nothing new is started)
          if (P.mode == 0) {
            seg.start_minimise();
          } else {
            seg.start_dimer();
          }
        }
        tm.mark(90);
        tm.fini();
      } else if (tm.tid == 0) {
        P.flags[a] = 0;
      }
    }

    // The last team to arrive compiles the selection lists of this tick (in replica order, so the
    // grid kernels see a deterministic work order) and decides whether the graph body repeats.
    if (tm.rank == 0) {
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        last_team = atomicAdd(&P.ctl->ticket, 1) == n_listed - 1;
      }
      __syncthreads();
      if (last_team && threadIdx.x < 32) {
        __threadfence();
        const int lane = threadIdx.x;
        const int tick = P.ctl->tick + 1;
        const bool out_of_ticks = tick >= P.max_ticks;
        // Filter the old list in place (order preserved) with the flags its teams left; four independent loads
        // per lane and sweep keep this serial tail short while thousands of replicas are listed.
        int na = 0, nr = 0;
        for (int base = 0; base < n_listed; base += 128) {
          int q[4], f[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int a = base + 32 * u + lane;
            q[u] = a < n_listed ? P.active[a] : 0;
            f[u] = a < n_listed ? __ldcg(P.flags + a) : 0;
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const bool act = f[u] & 1, reb = (f[u] & 2) != 0;
            if (act && out_of_ticks) P.state[q[u]].iter = -1;  // flagged: finishes as kStatusUnfinished next tick
            unsigned ma = __ballot_sync(0xffffffffu, act), mr = __ballot_sync(0xffffffffu, reb);
            if (act) P.active[na + __popc(ma & ((1u << lane) - 1u))] = q[u];
            if (reb) P.rebuild[nr + __popc(mr & ((1u << lane) - 1u))] = q[u];
            na += __popc(ma), nr += __popc(mr);
          }
        }
        if (lane == 0) {
          P.ctl->n_active = na, P.ctl->n_rebuild = nr, P.ctl->ticket = 0, P.ctl->tick = tick;
          P.ctl->kernels += 3 + (P.fused_rebuild ? 1 : ((!P.use_cond_rebuild || nr > 0) ? 6 : 0));
          __threadfence();
          // The device-side call that sets a conditional handle is only made when the value changes (it is a good part of
          // the ~10 us this bookkeeping takes in a late tick): the WHILE handle is 1 at launch and only ever goes to 0, the IF
          // handle of this tick of the body is 0 at launch and its last value is kept in the control block.
          if (P.use_cond && !(na > P.exit_below && tick <= P.max_ticks)) cudaGraphSetConditional(P.cond, 0u);
          if (P.use_cond_rebuild) {
            const int want = nr > 0 ? 1 : 0;
            int *have = &P.ctl->if_state[P.stage][P.body_slot];
            if (want || *have) cudaGraphSetConditional(P.cond_rebuild, (unsigned)want), *have = want;  // 1 is always (re)stated
          }
          if (P.trace != nullptr) trace_mark(P, P.trace_replica, 0, 95);  // end of the tick's bookkeeping
        }
      }
    }
  }

  // Start of a search call: the first n0 = min(slots, queue) searches go to slots 0..n0-1, the queue head stands behind them,
  // every search is marked unfinished until a slot reports it.
  static __global__ void k_search_init(SearchParams P, int n0) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nt = gridDim.x * blockDim.x;
    for (long s = t; s < P.n_queue; s += nt) P.status[s] = kStatusUnfinished;
    for (int a = t; a < n0; a += nt) P.active[a] = a, P.state[a].qid = a;
    if (t == 0) P.ctl->n_active = n0, P.ctl->queue_head = n0;
  }

}  // namespace ofly
