// C ABI: minimise::LBFGS::minimise and saddle::Dimer::find_sp over a whole batch (search.cuh).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "internal.hpp"
#include "search.cuh"

namespace ofly {
  // from api.cu
  int collect_faults(ofly_batch *b);
  void make_grid_for(ofly_batch *b, double r_list);
  void grow_buffers(ofly_batch *b);

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

    // Cluster width: as many SMs per replica as the batch leaves free, but never more CTAs than the
    // replica has rounds of atoms for.
    int pick_cluster(ofly_ctx *ctx, long R, long N) {
      if (char const *v = std::getenv("OFLY_CLUSTER")) {
        int c = std::atoi(v);
        if (c == 1 || c == 2 || c == 4 || c == 8) return c;
      }
      int c = 1;
      while (c < 8 && R * (c * 2) <= ctx->sm_count && (long)(c * 2) * kSearchThreads / kSearchGroup <= N) c *= 2;
      return c;
    }

    template <int C>
    void launch_search(ofly_ctx *ctx, int teams, BatchDev const &B, TablesDev const &T, SearchParams const &P) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(teams * C));
      cfg.blockDim = dim3(kSearchThreads);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = ctx->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = C, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = C > 1 ? 1 : 0;
      OFLY_CUDA(cudaLaunchKernelEx(&cfg, k_search<C>, B, T, P));
      ctx->launches++;
    }

    struct DevPtrs {
      const double *pos_in, *axis_in, *in_min, *hist_sp;
      double *pos_out, *axis_out, *grad_out;
      int *status;
      long *counters;
    };

    // Shared driver: grid set-up, workspace, launch, capacity-fault retry. All pointers are device memory.
    void run_search(ofly_batch *b, SearchParams P, double base_radius, double skin_frac, DevPtrs io) {
      ofly_ctx *ctx = b->ctx;
      OFLY_CUDA(cudaSetDevice(ctx->device));
      require(P.n_hist >= 1 && P.n_hist <= kMaxHist, "L-BFGS history size must be in [1, 16]");
      require(P.mode == 0 || (P.rot_n >= 1 && P.rot_n <= kMaxHist), "rotor L-BFGS history size must be in [1, 16]");

      // skin = max(skin_frac^(1/3) - 1, 0) * r   (lbfgs.cpp:39, dimer.cpp:47)
      double skin = std::max(std::pow(skin_frac, 1. / 3.) - 1, 0.0) * base_radius;
      double r_list = base_radius + skin;
      P.skin = skin;
      if (!(b->r_list == r_list)) make_grid_for(b, r_list);

      const size_t n3p = ((size_t)3 * b->N + 3) & ~(size_t)3;
      const int C = pick_cluster(ctx, b->R, b->N);
      const int teams = (int)std::min<long>(b->R, std::max(1, ctx->sm_count / C));
      P.ws_stride = search_ws_doubles(P.mode, P.n_hist, P.rot_n, n3p);
      b->search_ws.reserve((P.ws_stride * (size_t)teams) * sizeof(double) + 256);
      P.ws = reinterpret_cast<double *>(b->search_ws.p + 256);
      P.queue = reinterpret_cast<int *>(b->search_ws.p);
      P.pos_in = io.pos_in, P.axis_in = io.axis_in, P.in_min = io.in_min, P.hist_sp = io.hist_sp;
      P.pos_out = io.pos_out, P.axis_out = io.axis_out, P.grad_out = io.grad_out, P.status = io.status, P.counters = io.counters;

      for (int attempt = 0;; attempt++) {
        OFLY_CUDA(cudaMemsetAsync(P.queue, 0, sizeof(int), ctx->stream));
        switch (C) {
          case 8:
            launch_search<8>(ctx, teams, b->d, b->tables->dev, P);
            break;
          case 4:
            launch_search<4>(ctx, teams, b->d, b->tables->dev, P);
            break;
          case 2:
            launch_search<2>(ctx, teams, b->d, b->tables->dev, P);
            break;
          default:
            launch_search<1>(ctx, teams, b->d, b->tables->dev, P);
            break;
        }
        b->built = true;
        // The fault words double as the capacity report; reading them is the call's one synchronisation.
        int f = collect_faults(b);
        if (!(f & kFaultCapacity)) break;
        if (attempt == 2) throw Error(OFLY_ERR_CAPACITY, "neighbour buffers overflowed after regrowing");
        grow_buffers(b);
      }
    }

    SearchParams lbfgs_params(const ofly_lbfgs_options &o) {
      SearchParams P{};
      P.mode = 0;
      P.n_hist = o.n, P.iter_max = o.iter_max;
      P.f2norm = o.f2norm, P.proj_tol = o.proj_tol, P.max_trust = o.max_trust, P.min_trust = o.min_trust;
      P.grow_trust = o.grow_trust, P.shrink_trust = o.shrink_trust;
      return P;
    }

    SearchParams dimer_params(const ofly_dimer_options &o, const ofly_rotor_options &ro, double theta_tol, int count_frozen, long n_hist) {
      SearchParams P{};
      P.mode = 1;
      P.n_hist = o.n, P.iter_max = o.max_steps;
      P.f2norm = o.f2norm, P.proj_tol = o.proj_tol, P.max_trust = o.max_trust, P.min_trust = o.min_trust;
      P.grow_trust = o.grow_trust, P.shrink_trust = o.shrink_trust;
      P.hist_check_freq = std::max(1, o.hist_check_freq), P.use_history = o.use_history, P.convex_max = o.convex_max;
      P.rot_n = ro.n, P.iter_max_rot = ro.iter_max_rot, P.relax_in_convex = ro.relax_in_convex;
      P.delta_r = ro.delta_r, P.rot_theta_tol = ro.theta_tol;
      P.cos_theta_tol = std::cos(theta_tol);
      P.count_frozen = count_frozen;
      P.n_sp = n_hist;
      return P;
    }

    // Host-pointer staging: one device arena per call.
    struct Arena {
      ofly_ctx *ctx;
      std::vector<void *> ptrs;
      explicit Arena(ofly_ctx *c) : ctx(c) {}
      ~Arena() {
        for (void *p : ptrs) cudaFree(p);
      }
      template <typename T>
      T *alloc(size_t n) {
        void *p = nullptr;
        OFLY_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
        ptrs.push_back(p);
        return static_cast<T *>(p);
      }
      template <typename T>
      T *upload(const T *host, size_t n) {
        T *d = alloc<T>(n);
        if (n) OFLY_CUDA(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        return d;
      }
    };

  }  // namespace
}  // namespace ofly

using namespace ofly;

extern "C" {

void ofly_lbfgs_options_default(ofly_lbfgs_options *o) {
  if (!o) return;
  o->n = 10, o->iter_max = 2000, o->f2norm = 1e-6, o->skin_frac = 1.1, o->proj_tol = 0, o->max_trust = 0.5, o->min_trust = 0.01;
  o->grow_trust = 1.5, o->shrink_trust = 0.5;
}
void ofly_dimer_options_default(ofly_dimer_options *o) {
  if (!o) return;
  o->n = 10, o->max_steps = 1000, o->hist_check_freq = 5, o->use_history = 1, o->f2norm = 1e-5, o->skin_frac = 1.1, o->proj_tol = 0;
  o->max_trust = 0.5, o->min_trust = 0.05, o->grow_trust = 1.5, o->shrink_trust = 0.5, o->convex_max = 3;
}
void ofly_rotor_options_default(ofly_rotor_options *o) {
  if (!o) return;
  o->n = 6, o->iter_max_rot = 20, o->relax_in_convex = 1, o->delta_r = 0.001, o->theta_tol = 1 * M_PI / 360.;
}

int ofly_lbfgs_minimise_dev(ofly_batch *b, const ofly_lbfgs_options *opt, const double *pos_in_dev, double *pos_out_dev,
                            double *grad_out_dev, int *status_dev, long *counters_dev) {
  return guarded([&] {
    require(b && opt && pos_in_dev && pos_out_dev && status_dev, "null argument");
    DevPtrs io{pos_in_dev, nullptr, nullptr, nullptr, pos_out_dev, nullptr, grad_out_dev, status_dev, counters_dev};
    run_search(b, lbfgs_params(*opt), b->tables->host.r_cut, opt->skin_frac, io);
  });
}

int ofly_lbfgs_minimise(ofly_batch *b, const ofly_lbfgs_options *opt, const double *pos_in_host, double *pos_out_host,
                        double *grad_out_host, int *status_host, long *counters_host) {
  return guarded([&] {
    require(b && opt && pos_in_host && pos_out_host && status_host, "null argument");
    OFLY_CUDA(cudaSetDevice(b->ctx->device));
    size_t nv = (size_t)b->R * 3 * (size_t)b->N, R = (size_t)b->R;
    Arena A(b->ctx);
    DevPtrs io{};
    io.pos_in = A.upload(pos_in_host, nv);
    io.pos_out = A.alloc<double>(nv);
    io.grad_out = A.alloc<double>(nv);
    io.status = A.alloc<int>(R);
    io.counters = A.alloc<long>(4 * R);
    run_search(b, lbfgs_params(*opt), b->tables->host.r_cut, opt->skin_frac, io);
    cudaStream_t st = b->ctx->stream;
    OFLY_CUDA(cudaMemcpyAsync(pos_out_host, io.pos_out, nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (grad_out_host) OFLY_CUDA(cudaMemcpyAsync(grad_out_host, io.grad_out, nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaMemcpyAsync(status_host, io.status, R * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (counters_host) OFLY_CUDA(cudaMemcpyAsync(counters_host, io.counters, 4 * R * sizeof(long), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaStreamSynchronize(st));
  });
}

int ofly_dimer_find_sp_dev(ofly_batch *b, const ofly_dimer_options *opt, const ofly_rotor_options *ropt, const double *pos_in_dev,
                           const double *axis_in_dev, const double *in_min_dev, const double *hist_sp_dev, long n_hist,
                           double theta_tol, int count_frozen, double *pos_out_dev, double *axis_out_dev, int *exit_code_dev,
                           long *counters_dev) {
  return guarded([&] {
    require(b && opt && ropt && pos_in_dev && axis_in_dev && pos_out_dev && axis_out_dev && exit_code_dev, "null argument");
    require(n_hist == 0 || (in_min_dev && hist_sp_dev), "history given without in_min / hist_sp");
    DevPtrs io{pos_in_dev, axis_in_dev, in_min_dev, hist_sp_dev, pos_out_dev, axis_out_dev, nullptr, exit_code_dev, counters_dev};
    // list radius: rotor.r_cut(pot) = pot.r_cut + 2 delta_r (rotor.hpp:123) plus the skin (dimer.cpp:47-53)
    run_search(b, dimer_params(*opt, *ropt, theta_tol, count_frozen, n_hist), b->tables->host.r_cut + 2 * ropt->delta_r, opt->skin_frac, io);
  });
}

int ofly_dimer_find_sp(ofly_batch *b, const ofly_dimer_options *opt, const ofly_rotor_options *ropt, const double *pos_in_host,
                       const double *axis_in_host, const double *in_min_host, const double *hist_sp_host, long n_hist,
                       double theta_tol, int count_frozen, double *pos_out_host, double *axis_out_host, int *exit_code_host,
                       long *counters_host) {
  return guarded([&] {
    require(b && opt && ropt && pos_in_host && axis_in_host && pos_out_host && axis_out_host && exit_code_host, "null argument");
    require(n_hist == 0 || (in_min_host && hist_sp_host), "history given without in_min / hist_sp");
    OFLY_CUDA(cudaSetDevice(b->ctx->device));
    size_t n3 = 3 * (size_t)b->N, nv = (size_t)b->R * n3, R = (size_t)b->R;
    Arena A(b->ctx);
    DevPtrs io{};
    io.pos_in = A.upload(pos_in_host, nv);
    io.axis_in = A.upload(axis_in_host, nv);
    io.in_min = in_min_host ? A.upload(in_min_host, n3) : nullptr;
    io.hist_sp = n_hist ? A.upload(hist_sp_host, (size_t)n_hist * n3) : nullptr;
    io.pos_out = A.alloc<double>(nv);
    io.axis_out = A.alloc<double>(nv);
    io.status = A.alloc<int>(R);
    io.counters = A.alloc<long>(4 * R);
    run_search(b, dimer_params(*opt, *ropt, theta_tol, count_frozen, n_hist), b->tables->host.r_cut + 2 * ropt->delta_r, opt->skin_frac, io);
    cudaStream_t st = b->ctx->stream;
    OFLY_CUDA(cudaMemcpyAsync(pos_out_host, io.pos_out, nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaMemcpyAsync(axis_out_host, io.axis_out, nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaMemcpyAsync(exit_code_host, io.status, R * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (counters_host) OFLY_CUDA(cudaMemcpyAsync(counters_host, io.counters, 4 * R * sizeof(long), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaStreamSynchronize(st));
  });
}

}  // extern "C"
