// C ABI: the numerics of saddle::Master that follow a successful dimer search, batched over the saddle points of one
// environment:
//   Master::do_adj_min        src/saddle/find.cpp:1014-1090  +-nudge along the axis, two minimisations, swap, drift correction
//   Master::saddle_2_mech     src/saddle/find.cpp:757-871    pathway checks, three energies, mass-weighted Hessian at the
//                                                            saddle, eigenvalue rules, sum of log(lambda)
//   Master::calc_minima_hess  src/saddle/find.cpp:1094-1144  the same rules for the initial minimum
//   prefactor                 src/saddle/find.cpp:284-290
// The reference runs these per saddle inside an OpenMP task; here S saddles go through each stage together: one kernel
// builds the 2S nudged start states, ONE queue call relaxes all of them (search.cuh), one kernel per saddle orders /
// drift-corrects the pathway and measures it, the energies and the dense Hessians are evaluated R replicas at a time,
// the Hessians of a chunk are assembled by one launch (hessian.cuh) and solved concurrently.
// Not covered: the reconstruction test (recon_relax, find.cpp:873-940) -- it needs env::Geometry (nauty), out of scope.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "internal.hpp"
#include "neigh.cuh"

namespace ofly {
  // api.cu (the first three are defined inside its extern "C" block)
  extern "C" {
  void assemble_hessians(ofly_batch *b, const int *replicas_host, int count, int mass_weight);
  void eigenvalues_dev(ofly_ctx *ctx, double *H_dev, int n3, int count, double *eig_dev);
  void energies_dev(ofly_batch *b, double *energy_dev);
  }
  void rebuild_checked(ofly_batch *b, double r_list, const double *pos_dev);
  int collect_faults(ofly_batch *b);

  namespace {

    template <typename F>
    int guarded(F &&f) {
      try {
        f();
        return OFLY_OK;
      } catch (Error const &e) {
        set_last_error(e.what());
        return e.code;
      } catch (std::exception const &e) {
        set_last_error(e.what());
        return OFLY_ERR_INVALID;
      }
    }
    void require(bool ok, char const *msg) {
      if (!ok) throw Error(OFLY_ERR_INVALID, msg);
    }

    constexpr int kMechThreads = 256;

    // Fixed-order block sum of NV values (every thread gets the totals).
    template <int NV>
    __device__ __forceinline__ void block_sum(double (&v)[NV], double *sm /* [NV][8] */) {
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
      for (int q = 0; q < NV; q++) {
        double x = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sm[q * 8 + w] = x;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < NV; q++) {
        double x = 0;
        for (int k = 0; k < kMechThreads / 32; k++) x += sm[q * 8 + k];
        v[q] = x;
      }
      __syncthreads();
    }

    // do_adj_min, find.cpp:1019-1024 and :1043: start[2s] = sp + axis * disp * nudge, start[2s+1] = sp - axis * disp * nudge,
    // disp = |sp - in|.
    __global__ void __launch_bounds__(kMechThreads) k_mech_nudge(int n3, const double *__restrict__ in_min, const double *__restrict__ sp,
                                                                 const double *__restrict__ axis, double nudge_frac, double *__restrict__ start) {
      __shared__ double sm[8];
      const size_t s = blockIdx.x;
      const double *x = sp + s * n3, *ax = axis + s * n3;
      double a[1] = {0};
      for (int t = threadIdx.x; t < n3; t += kMechThreads) {
        const double d = x[t] - in_min[t];
        a[0] = fma(d, d, a[0]);
      }
      block_sum<1>(a, sm);
      const double disp = sqrt(a[0]);
      double *f = start + (2 * s) * n3, *r = start + (2 * s + 1) * n3;
      for (int t = threadIdx.x; t < n3; t += kMechThreads) {
        const double h = ax[t] * disp * nudge_frac;
        f[t] = x[t] + h, r[t] = x[t] - h;
      }
    }

    struct PathRec {
      double i2r, r2f, r2w, w2f;
      int max_atom, swapped;
    };

    // do_adj_min, find.cpp:1052-1082, and the four pathway lengths of saddle_2_mech (:772-775) for one saddle per CTA.
    __global__ void __launch_bounds__(kMechThreads) k_mech_pathway(int N, const double *__restrict__ in_min, const double *__restrict__ sp,
                                                                   const double *__restrict__ relaxed, double *__restrict__ rev_out,
                                                                   double *__restrict__ sp_out, double *__restrict__ fwd_out, PathRec *__restrict__ rec) {
      __shared__ double sm[8 * 8];
      __shared__ double best_v[kMechThreads / 32];
      __shared__ int best_i[kMechThreads / 32];
      const int n3 = 3 * N;
      const size_t s = blockIdx.x;
      const double *fwd = relaxed + (2 * s) * n3, *rev = relaxed + (2 * s + 1) * n3, *x = sp + s * n3;
      // swap if forward went to final: gnorm(rev - in) > gnorm(fwd - in)
      double d2[2] = {0, 0};
      for (int t = threadIdx.x; t < n3; t += kMechThreads) {
        const double a = rev[t] - in_min[t], b = fwd[t] - in_min[t];
        d2[0] = fma(a, a, d2[0]), d2[1] = fma(b, b, d2[1]);
      }
      block_sum<2>(d2, sm);
      const bool swapped = sqrt(d2[0]) > sqrt(d2[1]);
      if (swapped) {
        const double *tmp = fwd;
        fwd = rev, rev = tmp;
      }
      // centre-of-mass drift picked up during the search: centroid(rev) - centroid(in)
      double c[6] = {0, 0, 0, 0, 0, 0};
      for (int i = threadIdx.x; i < N; i += kMechThreads) {
#pragma unroll
        for (int k = 0; k < 3; k++) c[k] += rev[3 * i + k], c[3 + k] += in_min[3 * i + k];
      }
      block_sum<6>(c, sm);
      double drift[3];
#pragma unroll
      for (int k = 0; k < 3; k++) drift[k] = c[k] / (double)N - c[3 + k] / (double)N;
      double *ro = rev_out + s * n3, *so = sp_out + s * n3, *fo = fwd_out + s * n3;
      double l[4] = {0, 0, 0, 0};
      double bv = -1.0;
      int bi = 0;
      for (int i = threadIdx.x; i < N; i += kMechThreads) {
        double rr[3], ff[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const int t = 3 * i + k;
          const double r = rev[t] - drift[k], f = fwd[t] - drift[k], w = x[t] - drift[k];
          ro[t] = r, fo[t] = f, so[t] = w;
          rr[k] = r, ff[k] = f;
          const double a = r - in_min[t], b = f - r, cc = r - w, d = f - w;
          l[0] = fma(a, a, l[0]), l[1] = fma(b, b, l[1]), l[2] = fma(cc, cc, l[2]), l[3] = fma(d, d, l[3]);
        }
        // min_max (find.cpp:148-174): atom with the largest |rev_i - fwd_i|^2, first one on ties
        const double dx = rr[0] - ff[0], dy = rr[1] - ff[1], dz = rr[2] - ff[2];
        const double v = (dx * dx + dy * dy) + dz * dz;
        if (v > bv) bv = v, bi = i;
      }
      block_sum<4>(l, sm);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
      }
      if ((threadIdx.x & 31) == 0) best_v[threadIdx.x >> 5] = bv, best_i[threadIdx.x >> 5] = bi;
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int k = 1; k < kMechThreads / 32; k++) {
          if (best_v[k] > bv || (best_v[k] == bv && best_i[k] < bi)) bv = best_v[k], bi = best_i[k];
        }
        PathRec out;
        out.i2r = sqrt(l[0]), out.r2f = sqrt(l[1]), out.r2w = sqrt(l[2]), out.w2f = sqrt(l[3]);
        out.max_atom = bi, out.swapped = swapped ? 1 : 0;
        rec[s] = out;
      }
    }

    // m_sep_list (find.cpp:223-244): for every atom the index of the atom furthest away under the minimum image.
    __global__ void __launch_bounds__(256) k_sep_list(int N, BoxDev box, const double *__restrict__ pos, int *__restrict__ out) {
      const int i = blockIdx.x;
      __shared__ double bvs[8];
      __shared__ int bis[8];
      const double xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
      double bv = 0;
      int bi = 0;
      for (int j = threadIdx.x; j < N; j += 256) {
        double d[3] = {pos[3 * j] - xi, pos[3 * j + 1] - yi, pos[3 * j + 2] - zi};
#pragma unroll
        for (int k = 0; k < 3; k++) {
          if (box.periodic[k]) d[k] -= box.L[k] * rint(d[k] * box.invL[k]);
        }
        const double v = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
        if (v > bv) bv = v, bi = j;  // strict: the first of equally distant atoms wins (ascending j per thread)
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) bv = ov, bi = oi;
      }
      if ((threadIdx.x & 31) == 0) bvs[threadIdx.x >> 5] = bv, bis[threadIdx.x >> 5] = bi;
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int k = 1; k < 8; k++) {
          if (bvs[k] > bv || (bvs[k] == bv && bis[k] < bi)) bv = bvs[k], bi = bis[k];
        }
        out[i] = bi;
      }
    }

    // List::for_neighbours (list.hpp:138-154) of ONE atom, served on the device: entries of atom i's row within r_cut as
    // (real index, r, dr) triples, compacted in row order. One warp.
    __global__ void __launch_bounds__(32) k_neigh_for_atom(BatchDev B, int r, int i, double r_cut_sq, int cap, int *__restrict__ idx,
                                                           double *__restrict__ dist, double *__restrict__ dr, int *__restrict__ count) {
      const double4 *pos = B.pos4 + (size_t)r * B.N;
      const uint32_t *row = B.nbr + ((size_t)r * B.N + i) * B.K;
      const int n = B.cnt[(size_t)r * B.N + i];
      const double4 pi = pos[i];
      int found = 0;
      for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + (int)threadIdx.x;
        bool hit = false;
        int j = 0;
        double dx = 0, dy = 0, dz = 0, r2 = 0;
        if (k < n) {
          const uint32_t e = row[row_slot(k)];
          j = (int)(e & kIdxMask);
          const int code = (int)((e >> 24) & 31u);
          const double4 pj = pos[j];
          dx = (pj.x + shift_of(code % 3, B.box.L[0])) - pi.x;
          dy = (pj.y + shift_of((code / 3) % 3, B.box.L[1])) - pi.y;
          dz = (pj.z + shift_of(code / 9, B.box.L[2])) - pi.z;
          r2 = (dx * dx + dy * dy) + dz * dz;  // gnorm_sq's order (core.hpp:655-680)
          hit = r2 < r_cut_sq;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (hit) {
          const int w = found + __popc(bal & ((1u << threadIdx.x) - 1u));
          if (w < cap) idx[w] = j, dist[w] = sqrt(r2), dr[3 * w] = dx, dr[3 * w + 1] = dy, dr[3 * w + 2] = dz;
        }
        found += __popc(bal);
      }
      if (threadIdx.x == 0) *count = found;
    }

    // for_neighbours of EVERY atom of a replica within r_cut (the loop of Catalogue::rebuild_impl / rebuild_geo_from_nl,
    // src/env/catalogue.cpp:80-96, and of DetectVacancies, src/utility/lattice.cpp:69-84): a sub-warp of 8 lanes per atom.
    // FILL = false: counts[i] = hits; FILL = true: (index, r, dr) written from offsets[i], in stored order.
    template <bool FILL>
    __global__ void __launch_bounds__(256) k_neigh_within_all(BatchDev B, int r, double r_cut_sq, int *__restrict__ counts,
                                                             const long *__restrict__ offsets, int *__restrict__ idx, double *__restrict__ dist,
                                                             double *__restrict__ dr) {
      constexpr int G = 8;
      const int i = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) / G), lane = threadIdx.x % G;
      if (i >= B.N) return;
      const unsigned gmask = group_mask<G>();
      const int sub = ((threadIdx.x & 31) / G) * G;
      const double4 *pos = B.pos4 + (size_t)r * B.N;
      const uint32_t *row = B.nbr + ((size_t)r * B.N + i) * B.K;
      const int n = B.cnt[(size_t)r * B.N + i];
      const double4 pi = pos[i];
      long found = 0;
      for (int k0 = 0; k0 < n; k0 += G) {
        const int k = k0 + lane;
        bool hit = false;
        int j = 0;
        double dx = 0, dy = 0, dz = 0, r2 = 0;
        if (k < n) {
          const uint32_t e = row[row_slot(k)];
          j = (int)(e & kIdxMask);
          const int code = (int)((e >> 24) & 31u);
          const double4 pj = pos[j];
          dx = (pj.x + shift_of(code % 3, B.box.L[0])) - pi.x;
          dy = (pj.y + shift_of((code / 3) % 3, B.box.L[1])) - pi.y;
          dz = (pj.z + shift_of(code / 9, B.box.L[2])) - pi.z;
          r2 = (dx * dx + dy * dy) + dz * dz;
          hit = r2 < r_cut_sq;
        }
        const unsigned bal = __ballot_sync(gmask, hit) >> sub;
        if (FILL && hit) {
          const long w = offsets[i] + found + __popc(bal & ((1u << lane) - 1u));
          idx[w] = j, dist[w] = sqrt(r2), dr[3 * w] = dx, dr[3 * w + 1] = dy, dr[3 * w + 2] = dz;
        }
        found += __popc(bal & ((1u << G) - 1u));
      }
      if (!FILL && lane == 0) counts[i] = (int)found;
    }

    // Master::perturb (find.cpp:1147-1214) for one start state per CTA. z: this state's standard normal draws, [N][6]
    // (three for the position, three for the axis of every atom; only those of atoms inside r_pert are used).
    __global__ void __launch_bounds__(kMechThreads) k_master_perturb(BatchDev B, int centre, double r_pert, const double *__restrict__ in_min,
                                                                     const double *__restrict__ stddev, const double *__restrict__ z_all,
                                                                     double *__restrict__ pos_out, double *__restrict__ axis_out) {
      __shared__ double sm[8];
      const size_t s = blockIdx.x;
      const int n3 = 3 * B.N;
      double *out = pos_out + s * n3, *ax = axis_out + s * n3;
      const double *z = z_all + s * (size_t)B.N * 6;
      const double sd = stddev[s];
      for (int t = threadIdx.x; t < n3; t += kMechThreads) out[t] = in_min[t], ax[t] = 0.0;
      __syncthreads();
      // neighbours of the centre within r_pert, from replica 0's list (built from in_min at a radius >= r_pert)
      const double4 *pos = B.pos4;
      const uint32_t *row = B.nbr + (size_t)centre * B.K;
      const int n = B.cnt[centre];
      const double4 pc = pos[centre];
      for (int k = threadIdx.x; k < n; k += kMechThreads) {
        const uint32_t e = row[row_slot(k)];
        const int j = (int)(e & kIdxMask), code = (int)((e >> 24) & 31u);
        const double4 pj = pos[j];
        const double dx = (pj.x + shift_of(code % 3, B.box.L[0])) - pc.x, dy = (pj.y + shift_of((code / 3) % 3, B.box.L[1])) - pc.y,
                     dz = (pj.z + shift_of(code / 9, B.box.L[2])) - pc.z;
        const double r2 = (dx * dx + dy * dy) + dz * dz;
        if (r2 < r_pert * r_pert && !B.frozen[j]) {
          const double env = 1. - sqrt(r2) / r_pert;
          for (int c = 0; c < 3; c++) {  // atomics: the same atom may be reached through two images in a small box
            atomicAdd(out + 3 * j + c, sd * z[6 * j + c] * env);
            atomicAdd(ax + 3 * j + c, z[6 * j + 3 + c]);
          }
        }
      }
      __syncthreads();
      if (threadIdx.x < 3) {
        out[3 * centre + threadIdx.x] += sd * z[6 * centre + threadIdx.x];
        ax[3 * centre + threadIdx.x] += z[6 * centre + 3 + threadIdx.x];
      }
      __syncthreads();
      double a[1] = {0};
      for (int t = threadIdx.x; t < n3; t += kMechThreads) a[0] = fma(ax[t], ax[t], a[0]);
      block_sum<1>(a, sm);
      const double nrm = sqrt(a[0]);
      for (int t = threadIdx.x; t < n3; t += kMechThreads) ax[t] /= nrm;
    }

    // centroid_align(x, with) (lattice.hpp:214-226) followed by gnorm(x - with): one CTA.
    __global__ void __launch_bounds__(kMechThreads) k_centroid_align(int N, double *__restrict__ x, const double *__restrict__ with, double *__restrict__ norm_out) {
      __shared__ double sm[6 * 8];
      double c[6] = {0, 0, 0, 0, 0, 0};
      for (int i = threadIdx.x; i < N; i += kMechThreads) {
#pragma unroll
        for (int k = 0; k < 3; k++) c[k] += with[3 * i + k], c[3 + k] += x[3 * i + k];
      }
      block_sum<6>(c, sm);
      double delta[3];
#pragma unroll
      for (int k = 0; k < 3; k++) delta[k] = c[k] / (double)N - c[3 + k] / (double)N;
      double a[1] = {0};
      for (int i = threadIdx.x; i < N; i += kMechThreads) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double v = x[3 * i + k] + delta[k];
          x[3 * i + k] = v;
          const double d = v - with[3 * i + k];
          a[0] = fma(d, d, a[0]);
        }
      }
      block_sum<1>(a, sm);
      if (threadIdx.x == 0) *norm_out = sqrt(a[0]);
    }

    // Eigenvalue rules shared by the minimum and the saddles (find.cpp:849-864, 1118-1128): modes [first, n) are summed if
    // > tol, counted as zero if within tol; a mode < -tol stops the scan. Returns the index of that mode or -1.
    int scan_modes(const double *freq, int n, int first, double tol, double &log_sum, int &zeros) {
      log_sum = 0, zeros = 0;
      for (int i = first; i < n; i++) {
        if (freq[i] < -tol) return i;
        if (freq[i] > tol) {
          log_sum += std::log(freq[i]);
        } else {
          ++zeros;
        }
      }
      return -1;
    }

    // The batch with `atom` frozen on top of b's own frozen flags (do_adj_min freezes the atom furthest from the centre
    // to pin the translational modes, find.cpp:1026-1034); kept with b and rebuilt when the atom changes.
    ofly_batch *relax_sibling(ofly_batch *b, long atom) {
      if (atom < 0) return b;
      if (b->sibling && b->sibling_atom == atom) return b->sibling;
      if (b->sibling) ofly_batch_destroy(b->sibling), b->sibling = nullptr;
      std::vector<uint32_t> type((size_t)b->N);
      for (long i = 0; i < b->N; i++) type[(size_t)i] = b->type_host[(size_t)i];
      std::vector<uint8_t> fz = b->frozen_host;
      fz[(size_t)atom] = 1;
      ofly_batch *sib = nullptr;
      int rc = ofly_batch_create(b->ctx, b->tables, b->extents, b->periodic, b->N, b->R, type.data(), fz.data(), &sib);
      if (rc != OFLY_OK) throw Error(rc, ofly_last_error());
      b->sibling = sib, b->sibling_atom = atom;
      return sib;
    }

    // Energies (and faults) of `count` <= R states given as device pointers, through b's R slots at radius r_cut.
    void energies_of(ofly_batch *b, std::vector<const double *> const &states, std::vector<double> &energy, std::vector<int> &fault) {
      const size_t n3 = 3 * (size_t)b->N, R = (size_t)b->R;
      require(!states.empty() && states.size() <= R, "internal: chunk larger than the batch");
      cudaStream_t st = b->ctx->stream;
      b->stage_a.reserve(R * n3), b->scratch.reserve(R * (size_t)b->N + R);
      for (size_t r = 0; r < R; r++) {
        const double *src = states[std::min(r, states.size() - 1)];  // pad with the last state
        OFLY_CUDA(cudaMemcpyAsync(b->stage_a.p + r * n3, src, n3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
      }
      rebuild_checked(b, b->tables->host.r_cut, b->stage_a.p);
      double *e_dev = b->scratch.p + R * (size_t)b->N;
      energies_dev(b, e_dev);
      std::vector<double> e(R);
      OFLY_CUDA(cudaMemcpyAsync(e.data(), e_dev, R * sizeof(double), cudaMemcpyDeviceToHost, st));
      collect_faults(b);  // synchronises
      energy.assign(e.begin(), e.begin() + (long)states.size());
      fault.assign(b->fault_host.begin(), b->fault_host.begin() + (long)states.size());
    }

  }  // namespace
}  // namespace ofly

using namespace ofly;

extern "C" {

void ofly_mech_options_default(ofly_mech_options *o) {
  if (!o) return;
  // saddle::Master::Options, include/libfly/saddle/find.hpp:112-184
  o->nudge_frac = 0.01, o->basin_tol = 0.05, o->stationary_tol = 0.05 / 2, o->hessian_eigen_zero_tol = 1e-6;
  o->centre = -1, o->freeze_atom = -1, o->num_zero_modes = 3, o->skip_pathway = 0;
}

double ofly_mech_prefactor(double log_prod_eigen_min, double log_prod_eigen_sp) {
  return std::sqrt(std::exp(log_prod_eigen_min - log_prod_eigen_sp) / (2 * M_PI * 1.6605390666050e-27));  // find.cpp:287
}

int ofly_mech_sep_list(ofly_batch *b, const double *pos_host, int *sep_host) {
  return guarded([&] {
    require(b && pos_host && sep_host, "null argument");
    OFLY_CUDA(cudaSetDevice(b->ctx->device));
    const size_t n3 = 3 * (size_t)b->N;
    cudaStream_t st = b->ctx->stream;
    b->stage_b.reserve(n3);
    DevBuf<int> out;
    out.reserve((size_t)b->N);
    OFLY_CUDA(cudaMemcpyAsync(b->stage_b.p, pos_host, n3 * sizeof(double), cudaMemcpyHostToDevice, st));
    k_sep_list<<<(unsigned)b->N, 256, 0, st>>>((int)b->N, b->d.box, b->stage_b.p, out.p);
    b->ctx->launches++;
    OFLY_CUDA(cudaMemcpyAsync(sep_host, out.p, sizeof(int) * (size_t)b->N, cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaStreamSynchronize(st));
    out.release();
  });
}

int ofly_neigh_for_atom(ofly_batch *b, long replica, long atom, double r_cut, int *idx_host, double *r_host, double *dr_host, long cap,
                        long *count) {
  return guarded([&] {
    require(b && idx_host && r_host && dr_host && count && cap >= 0, "null argument");
    require(b->built, "for_neighbours before rebuild()");
    require(replica >= 0 && replica < b->R && atom >= 0 && atom < b->N, "Atom is not in neighbour list");
    require(r_cut <= b->r_list, "r_cut is bigger then than the internal r_cut");  // list.hpp:141-143
    OFLY_CUDA(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    const size_t K = (size_t)b->d.K;
    DevBuf<int> d_idx;
    DevBuf<double> d_val;
    d_idx.reserve(K + 1), d_val.reserve(4 * K);
    k_neigh_for_atom<<<1, 32, 0, st>>>(b->d, (int)replica, (int)atom, r_cut * r_cut, (int)K, d_idx.p + 1, d_val.p, d_val.p + K, d_idx.p);
    b->ctx->launches++;
    std::vector<int> hi(K + 1);
    std::vector<double> hv(4 * K);
    OFLY_CUDA(cudaMemcpyAsync(hi.data(), d_idx.p, (K + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaMemcpyAsync(hv.data(), d_val.p, 4 * K * sizeof(double), cudaMemcpyDeviceToHost, st));
    cudaError_t e = cudaStreamSynchronize(st);
    d_idx.release(), d_val.release();
    OFLY_CUDA(e);
    const long n = hi[0];
    *count = n;
    for (long k = 0; k < std::min(n, cap); k++) {
      idx_host[k] = hi[(size_t)k + 1], r_host[k] = hv[(size_t)k];
      for (int c = 0; c < 3; c++) dr_host[3 * k + c] = hv[K + 3 * (size_t)k + c];
    }
  });
}

int ofly_neigh_within_all(ofly_batch *b, long replica, double r_cut, long *offsets_host, int *idx_host, double *r_host, double *dr_host,
                          long cap, long *total) {
  return guarded([&] {
    require(b && offsets_host && total, "null argument");
    require(b->built, "for_neighbours before rebuild()");
    require(replica >= 0 && replica < b->R, "replica out of range");
    require(r_cut <= b->r_list, "r_cut is bigger then than the internal r_cut");  // list.hpp:141-143
    OFLY_CUDA(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    const size_t N = (size_t)b->N;
    DevBuf<int> d_cnt, d_idx;
    DevBuf<long> d_off;
    DevBuf<double> d_val;
    struct Rel {
      DevBuf<int> *a, *b2;
      DevBuf<long> *c;
      DevBuf<double> *d;
      ~Rel() { a->release(), b2->release(), c->release(), d->release(); }
    } rel{&d_cnt, &d_idx, &d_off, &d_val};
    d_cnt.reserve(N), d_off.reserve(N + 1);
    const unsigned grid = (unsigned)((N * 8 + 255) / 256);
    k_neigh_within_all<false><<<grid, 256, 0, st>>>(b->d, (int)replica, r_cut * r_cut, d_cnt.p, nullptr, nullptr, nullptr, nullptr);
    std::vector<int> cnt(N);
    OFLY_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt.p, N * sizeof(int), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaStreamSynchronize(st));
    offsets_host[0] = 0;
    for (size_t i = 0; i < N; i++) offsets_host[i + 1] = offsets_host[i] + cnt[i];
    const long tot = offsets_host[N];
    *total = tot;
    b->ctx->launches++;
    if (tot == 0 || !idx_host || !r_host || !dr_host || cap < tot) return;  // sizes only (or the caller's buffers are too small)
    d_idx.reserve((size_t)tot), d_val.reserve(4 * (size_t)tot);
    OFLY_CUDA(cudaMemcpyAsync(d_off.p, offsets_host, (N + 1) * sizeof(long), cudaMemcpyHostToDevice, st));
    k_neigh_within_all<true><<<grid, 256, 0, st>>>(b->d, (int)replica, r_cut * r_cut, nullptr, d_off.p, d_idx.p, d_val.p, d_val.p + tot);
    OFLY_CUDA(cudaGetLastError());
    b->ctx->launches++;
    OFLY_CUDA(cudaMemcpyAsync(idx_host, d_idx.p, (size_t)tot * sizeof(int), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaMemcpyAsync(r_host, d_val.p, (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaMemcpyAsync(dr_host, d_val.p + tot, 3 * (size_t)tot * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaStreamSynchronize(st));
  });
}

int ofly_master_perturb(ofly_batch *b, long n_states, const double *in_min_host, long centre, double r_pert, const double *stddev_host,
                        const double *normals_host, double *pos_out_host, double *axis_out_host) {
  return guarded([&] {
    require(b && in_min_host && stddev_host && normals_host && pos_out_host && axis_out_host, "null argument");
    require(n_states >= 1 && centre >= 0 && centre < b->N, "bad argument");
    require(!b->frozen_host[(size_t)centre], "perturbation centred on a frozen atom");  // find.cpp:1205
    OFLY_CUDA(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    const size_t n3 = 3 * (size_t)b->N, S = (size_t)n_states;
    // nl_pert (find.cpp:217-219): replica 0's list rebuilt from the minimum at r_pert
    b->stage_a.reserve((size_t)b->R * n3);
    for (long r = 0; r < b->R; r++) OFLY_CUDA(cudaMemcpyAsync(b->stage_a.p + (size_t)r * n3, in_min_host, n3 * sizeof(double), cudaMemcpyHostToDevice, st));
    rebuild_checked(b, r_pert, b->stage_a.p);
    DevBuf<double> d_sd, d_z, d_pos, d_ax;
    struct Rel {
      DevBuf<double> *v[4];
      ~Rel() {
        for (auto *x : v) x->release();
      }
    } rel{{&d_sd, &d_z, &d_pos, &d_ax}};
    d_sd.reserve(S), d_z.reserve(S * 6 * (size_t)b->N), d_pos.reserve(S * n3), d_ax.reserve(S * n3);
    OFLY_CUDA(cudaMemcpyAsync(d_sd.p, stddev_host, S * sizeof(double), cudaMemcpyHostToDevice, st));
    OFLY_CUDA(cudaMemcpyAsync(d_z.p, normals_host, S * 6 * (size_t)b->N * sizeof(double), cudaMemcpyHostToDevice, st));
    k_master_perturb<<<(unsigned)S, kMechThreads, 0, st>>>(b->d, (int)centre, r_pert, b->stage_a.p, d_sd.p, d_z.p, d_pos.p, d_ax.p);
    OFLY_CUDA(cudaGetLastError());
    b->ctx->launches++;
    OFLY_CUDA(cudaMemcpyAsync(pos_out_host, d_pos.p, S * n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaMemcpyAsync(axis_out_host, d_ax.p, S * n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaStreamSynchronize(st));
  });
}

int ofly_skmc_relax_step(ofly_batch *b, const ofly_lbfgs_options *min_opt, const double *cell_host, const double *raw_recon_host,
                         double *rel_recon_host, double *E0, double *Ef, double *dR, int *min_failed) {
  return guarded([&] {
    require(b && min_opt && cell_host && raw_recon_host && E0 && Ef && dR && min_failed, "null argument");
    OFLY_CUDA(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    const size_t n3 = 3 * (size_t)b->N;
    DevBuf<double> d_cell, d_raw, d_rel, d_norm;
    DevBuf<int> d_status;
    struct Rel {
      DevBuf<double> *v[4];
      DevBuf<int> *i;
      ~Rel() {
        for (auto *x : v) x->release();
        i->release();
      }
    } rel{{&d_cell, &d_raw, &d_rel, &d_norm}, &d_status};
    d_cell.reserve(n3), d_raw.reserve(n3), d_rel.reserve(n3), d_norm.reserve(1), d_status.reserve(1);
    OFLY_CUDA(cudaMemcpyAsync(d_cell.p, cell_host, n3 * sizeof(double), cudaMemcpyHostToDevice, st));
    OFLY_CUDA(cudaMemcpyAsync(d_raw.p, raw_recon_host, n3 * sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<double> e;
    std::vector<int> fault;
    energies_of(b, {d_cell.p}, e, fault);  // E0 = energy(cell), skmc.hpp:297
    if (fault[0] & kFaultOutOfBounds) throw Error(OFLY_ERR_OUT_OF_BOUNDS, "x is outside tabulated region (Spline::OutOfBounds)");
    *E0 = e[0];
    int rc = ofly_lbfgs_minimise_stream_dev(b, min_opt, 1, d_raw.p, d_rel.p, nullptr, d_status.p, nullptr);  // skmc.hpp:301-303
    if (rc != OFLY_OK) throw Error(rc, ofly_last_error());
    int status = 0;
    OFLY_CUDA(cudaMemcpyAsync(&status, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    k_centroid_align<<<1, kMechThreads, 0, st>>>((int)b->N, d_rel.p, d_raw.p, d_norm.p);  // skmc.hpp:305, 309
    OFLY_CUDA(cudaGetLastError());
    b->ctx->launches++;
    OFLY_CUDA(cudaMemcpyAsync(dR, d_norm.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    if (rel_recon_host) OFLY_CUDA(cudaMemcpyAsync(rel_recon_host, d_rel.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaStreamSynchronize(st));
    if (status == -5) throw Error(OFLY_ERR_OUT_OF_BOUNDS, "x is outside tabulated region (Spline::OutOfBounds)");
    *min_failed = status != 0;
    energies_of(b, {d_rel.p}, e, fault);  // Ef = energy(rel_recon), skmc.hpp:307
    if (fault[0] & kFaultOutOfBounds) throw Error(OFLY_ERR_OUT_OF_BOUNDS, "x is outside tabulated region (Spline::OutOfBounds)");
    *Ef = e[0];
  });
}

int ofly_mech_minimum(ofly_batch *b, const double *in_min_host, double zero_tol, double *log_prod_eigen, int *num_zero_modes,
                      int *count_frozen, double *eig_host) {
  return guarded([&] {
    require(b && in_min_host && log_prod_eigen && num_zero_modes, "null argument");
    OFLY_CUDA(cudaSetDevice(b->ctx->device));
    const int n3 = 3 * (int)b->N;
    int frozen = 0;
    for (uint8_t f : b->frozen_host) frozen += f ? 1 : 0;
    if (count_frozen) *count_frozen = frozen;
    const int expect = frozen > 0 ? 3 * frozen : 3;  // find.cpp:1105
    cudaStream_t st = b->ctx->stream;
    b->stage_b.reserve((size_t)n3);
    OFLY_CUDA(cudaMemcpyAsync(b->stage_b.p, in_min_host, sizeof(double) * (size_t)n3, cudaMemcpyHostToDevice, st));
    std::vector<double> e;
    std::vector<int> fault;
    energies_of(b, {b->stage_b.p}, e, fault);
    if (fault[0] & kFaultOutOfBounds) throw Error(OFLY_ERR_OUT_OF_BOUNDS, "x is outside tabulated region (Spline::OutOfBounds)");
    const int r0 = 0;
    assemble_hessians(b, &r0, 1, 1);
    if (collect_faults(b) & kFaultOutOfBounds) throw Error(OFLY_ERR_OUT_OF_BOUNDS, "x is outside tabulated region (Spline::OutOfBounds)");
    b->hess_eig.reserve((size_t)n3);
    eigenvalues_dev(b->ctx, b->hess_H.p, n3, 1, b->hess_eig.p);
    std::vector<double> freq((size_t)n3);
    OFLY_CUDA(cudaMemcpy(freq.data(), b->hess_eig.p, sizeof(double) * (size_t)n3, cudaMemcpyDeviceToHost));
    if (eig_host) std::memcpy(eig_host, freq.data(), sizeof(double) * (size_t)n3);
    double log_sum;
    int zeros;
    const int bad = scan_modes(freq.data(), n3, 0, zero_tol, log_sum, zeros);
    if (bad >= 0) throw Error(OFLY_ERR_INVALID, "Mode " + std::to_string(bad) + " has eigenvalue = " + std::to_string(freq[(size_t)bad]));
    *log_prod_eigen = log_sum, *num_zero_modes = zeros;
    if (zeros != expect) throw Error(OFLY_ERR_INVALID, "Expecting " + std::to_string(expect) + " zero modes but fount " + std::to_string(zeros));
  });
}

int ofly_mech_from_saddles(ofly_batch *b, const ofly_mech_options *opt, const ofly_lbfgs_options *min_opt, long n_saddles,
                           const double *sp_host, const double *axis_host, const double *in_min_host, double *rev_host,
                           double *sp_out_host, double *fwd_host, ofly_mech_record *records) {
  return guarded([&] {
    require(b && opt && sp_host && in_min_host && records, "null argument");
    require(n_saddles >= 1, "number of saddles must be positive");
    require(opt->skip_pathway || (axis_host && min_opt), "the pathway needs the dimer axes and minimiser options");
    require(opt->freeze_atom < b->N && opt->centre < b->N, "atom index out of range");
    OFLY_CUDA(cudaSetDevice(b->ctx->device));
    ofly_ctx *ctx = b->ctx;
    cudaStream_t st = ctx->stream;
    const size_t S = (size_t)n_saddles, N = (size_t)b->N, n3 = 3 * N, R = (size_t)b->R;
    const double tol = opt->hessian_eigen_zero_tol;

    DevBuf<double> d_sp, d_ax, d_min, d_start, d_relaxed, d_rev, d_spo, d_fwd;
    DevBuf<int> d_status;
    DevBuf<PathRec> d_rec;
    struct Release {
      std::vector<DevBuf<double> *> v;
      DevBuf<int> *i;
      DevBuf<PathRec> *p;
      ~Release() {
        for (auto *x : v) x->release();
        i->release(), p->release();
      }
    } rel{{&d_sp, &d_ax, &d_min, &d_start, &d_relaxed, &d_rev, &d_spo, &d_fwd}, &d_status, &d_rec};

    d_sp.reserve(S * n3), d_min.reserve(n3);
    OFLY_CUDA(cudaMemcpyAsync(d_sp.p, sp_host, S * n3 * sizeof(double), cudaMemcpyHostToDevice, st));
    OFLY_CUDA(cudaMemcpyAsync(d_min.p, in_min_host, n3 * sizeof(double), cudaMemcpyHostToDevice, st));
    for (size_t s = 0; s < S; s++) records[s] = ofly_mech_record{};

    std::vector<PathRec> rec(S);
    std::vector<int> min_status(2 * S, 0);
    const double *rev_p = nullptr, *spo_p = d_sp.p, *fwd_p = nullptr;
    if (!opt->skip_pathway) {
      d_ax.reserve(S * n3), d_start.reserve(2 * S * n3), d_relaxed.reserve(2 * S * n3), d_status.reserve(2 * S);
      d_rev.reserve(S * n3), d_spo.reserve(S * n3), d_fwd.reserve(S * n3), d_rec.reserve(S);
      OFLY_CUDA(cudaMemcpyAsync(d_ax.p, axis_host, S * n3 * sizeof(double), cudaMemcpyHostToDevice, st));
      k_mech_nudge<<<(unsigned)S, kMechThreads, 0, st>>>((int)n3, d_min.p, d_sp.p, d_ax.p, opt->nudge_frac, d_start.p);
      OFLY_CUDA(cudaGetLastError());
      int frozen = 0;
      for (uint8_t f : b->frozen_host) frozen += f ? 1 : 0;
      ofly_batch *relax = relax_sibling(b, frozen == 0 ? opt->freeze_atom : -1);
      int rc = ofly_lbfgs_minimise_stream_dev(relax, min_opt, (long)(2 * S), d_start.p, d_relaxed.p, nullptr, d_status.p, nullptr);
      if (rc != OFLY_OK) throw Error(rc, ofly_last_error());
      k_mech_pathway<<<(unsigned)S, kMechThreads, 0, st>>>((int)N, d_min.p, d_sp.p, d_relaxed.p, d_rev.p, d_spo.p, d_fwd.p, d_rec.p);
      OFLY_CUDA(cudaGetLastError());
      ctx->launches += 2;
      OFLY_CUDA(cudaMemcpyAsync(rec.data(), d_rec.p, S * sizeof(PathRec), cudaMemcpyDeviceToHost, st));
      OFLY_CUDA(cudaMemcpyAsync(min_status.data(), d_status.p, 2 * S * sizeof(int), cudaMemcpyDeviceToHost, st));
      OFLY_CUDA(cudaStreamSynchronize(st));
      rev_p = d_rev.p, spo_p = d_spo.p, fwd_p = d_fwd.p;
      if (rev_host) OFLY_CUDA(cudaMemcpyAsync(rev_host, d_rev.p, S * n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (sp_out_host) OFLY_CUDA(cudaMemcpyAsync(sp_out_host, d_spo.p, S * n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
      if (fwd_host) OFLY_CUDA(cudaMemcpyAsync(fwd_host, d_fwd.p, S * n3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    } else if (sp_out_host) {
      std::memcpy(sp_out_host, sp_host, S * n3 * sizeof(double));
    }

    // pathway checks, in the reference's order
    std::vector<size_t> alive;
    for (size_t s = 0; s < S; s++) {
      ofly_mech_record &m = records[s];
      if (!opt->skip_pathway) {
        const PathRec &p = rec[s];
        m.i2r = p.i2r, m.r2f = p.r2f, m.r2w = p.r2w, m.w2f = p.w2f;
        const int sa = min_status[2 * s], sb = min_status[2 * s + 1];
        if (sa == -5 || sb == -5) {
          m.status = OFLY_MECH_OUT_OF_BOUNDS;
        } else if (sa != 0 || sb != 0) {
          m.status = OFLY_MECH_MINIMISATION_FAILED;  // find.cpp:1038-1050
        } else if (opt->centre >= 0 && p.max_atom != opt->centre) {
          m.status = OFLY_MECH_NOT_CENTRED;  // find.cpp:1084
        } else if (p.i2r > opt->basin_tol) {
          m.status = OFLY_MECH_UNLINKED;  // find.cpp:779
        } else if (p.r2f < opt->basin_tol) {
          m.status = OFLY_MECH_BACK_TO_INITIAL;  // find.cpp:783
        } else if (p.r2w < opt->stationary_tol) {
          m.status = OFLY_MECH_MIN_TO_SP_TOO_SHORT;  // find.cpp:787
        } else if (p.w2f < opt->stationary_tol) {
          m.status = OFLY_MECH_SP_TO_MIN_TOO_SHORT;  // find.cpp:791
        }
      }
      if (m.status == OFLY_MECH_OK) alive.push_back(s);
    }

    // energies and Hessians of the survivors, R at a time
    const size_t hs = n3 * n3;
    const size_t h_chunk = std::max<size_t>(1, std::min<size_t>(R, (size_t)12e9 / (hs * sizeof(double))));
    require(hs * sizeof(double) <= (size_t)40e9, "system too large for a dense Hessian");
    std::vector<double> e;
    std::vector<int> fault;
    double e_in = 0;
    if (opt->skip_pathway && !alive.empty()) {  // E0 of the given minimum
      energies_of(b, {d_min.p}, e, fault);
      if (fault[0] & kFaultOutOfBounds) throw Error(OFLY_ERR_OUT_OF_BOUNDS, "x is outside tabulated region (Spline::OutOfBounds)");
      e_in = e[0];
    }
    for (size_t c0 = 0; c0 < alive.size(); c0 += R) {
      const size_t m = std::min(R, alive.size() - c0);
      std::vector<const double *> st_rev(m), st_sp(m), st_fwd(m);
      for (size_t k = 0; k < m; k++) {
        const size_t s = alive[c0 + k];
        st_sp[k] = spo_p + s * n3;
        if (!opt->skip_pathway) st_rev[k] = rev_p + s * n3, st_fwd[k] = fwd_p + s * n3;
      }
      auto mark_oob = [&](std::vector<int> const &f) {
        for (size_t k = 0; k < m; k++) {
          if (f[k] & kFaultOutOfBounds) records[alive[c0 + k]].status = OFLY_MECH_OUT_OF_BOUNDS;
        }
      };
      if (!opt->skip_pathway) {
        energies_of(b, st_rev, e, fault);
        mark_oob(fault);
        for (size_t k = 0; k < m; k++) records[alive[c0 + k]].E0 = e[k];
        energies_of(b, st_fwd, e, fault);
        mark_oob(fault);
        for (size_t k = 0; k < m; k++) records[alive[c0 + k]].Ef = e[k];
      }
      energies_of(b, st_sp, e, fault);  // leaves the saddles' lists (at r_cut) in slots 0..m-1 for the Hessians
      mark_oob(fault);
      std::vector<int> want;  // slots whose Hessian is needed
      for (size_t k = 0; k < m; k++) {
        ofly_mech_record &r = records[alive[c0 + k]];
        r.Esp = e[k];
        if (opt->skip_pathway) r.E0 = e_in, r.Ef = e_in;
        r.barrier = r.Esp - r.E0, r.delta = r.Ef - r.E0;
        if (r.status == OFLY_MECH_OK && !(r.Esp > r.E0)) r.status = OFLY_MECH_ERR_SP_BELOW_MIN;  // verify, find.cpp:807
        if (r.status == OFLY_MECH_OK) want.push_back((int)k);
      }
      for (size_t w0 = 0; w0 < want.size(); w0 += h_chunk) {
        const int cnt = (int)std::min(h_chunk, want.size() - w0);
        const auto t0 = std::chrono::steady_clock::now();
        assemble_hessians(b, want.data() + w0, cnt, 1);
        collect_faults(b);
        const auto t1 = std::chrono::steady_clock::now();
        b->hess_eig.reserve((size_t)cnt * n3);
        eigenvalues_dev(ctx, b->hess_H.p, (int)n3, cnt, b->hess_eig.p);
        if (std::getenv("OFLY_VERBOSE")) {
          const auto t2 = std::chrono::steady_clock::now();
          std::fprintf(stderr, "openfly_b200: %d Hessians assembled in %.1f ms, eigenvalues in %.1f ms\n", cnt,
                       std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
        }
        std::vector<double> freq((size_t)cnt * n3);
        OFLY_CUDA(cudaMemcpy(freq.data(), b->hess_eig.p, freq.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int q = 0; q < cnt; q++) {
          const int slot = want[w0 + (size_t)q];
          ofly_mech_record &r = records[alive[c0 + (size_t)slot]];
          if (b->fault_host[(size_t)slot] & kFaultOutOfBounds) {
            r.status = OFLY_MECH_OUT_OF_BOUNDS;
            continue;
          }
          const double *f = freq.data() + (size_t)q * n3;
          r.min_mode = f[0];
          if (!(f[0] < -tol)) {  // verify, find.cpp:846
            r.status = OFLY_MECH_ERR_NO_NEGATIVE_MODE;
            continue;
          }
          double log_sum;
          int zeros;
          const int bad = scan_modes(f, (int)n3, 1, tol, log_sum, zeros);
          r.count_zeros = zeros;
          if (bad >= 0) {
            r.status = OFLY_MECH_HIGHER_ORDER_SADDLE;  // find.cpp:852-855
            continue;
          }
          r.log_prod_eigen = log_sum;
          if (zeros != opt->num_zero_modes) r.status = OFLY_MECH_ERR_ZERO_MODES;  // throw, find.cpp:862-867
        }
      }
    }
    OFLY_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
