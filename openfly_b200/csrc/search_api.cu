// C ABI: minimise::LBFGS::minimise and saddle::Dimer::find_sp over a whole batch (search.cuh).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>
#include <string>

#include "internal.hpp"
#include "search.cuh"
#include "tile.cuh"

namespace ofly {
  // from api.cu
  int collect_faults(ofly_batch *b);
  void make_grid_for(ofly_batch *b, double r_list);
  void grow_buffers(ofly_batch *b);
  int pick_group(ofly_ctx *ctx, long atoms);

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

    constexpr int kSegMaxThreads = 512;
    constexpr int kMaxTickUnroll = 8;   // ticks per pass of the device-side loop's body (tick_unroll)
    constexpr int kHostLoopTicks = 8;   // ticks between host checks when the device-side loop is off

    int env_int(char const *name, int dflt) {
      char const *v = std::getenv(name);
      return v ? std::atoi(v) : dflt;
    }

    struct TeamShape {
      int threads, cluster;
    };

    // Team of one replica's vector algebra. The passes are latency-bound streams (a thread's elements are
    // sequential round trips to L2/HBM) and for most ticks of a batch only a few replicas are still active, so
    // the team is made as wide as the vector allows: >= 2 elements per thread up to a 512-thread CTA, then a
    // cluster while every thread keeps >= 8 elements. The segment kernel strides over the list of active
    // replicas, so wide teams cost extra sweeps only while many replicas are active (the first few % of ticks).
    // Clusters of C 512-thread CTAs of the segment kernel the device can co-schedule (0: the shape cannot launch at
    // all -- MIG slices, parts with fewer SMs per GPC, ...). Asked once per device and shape.
    template <int C>
    int active_clusters(ofly_ctx *ctx) {
      static std::map<int, int> known;
      static std::mutex known_mutex;  // one host thread per device may be in here at the same time
      std::lock_guard<std::mutex> lock(known_mutex);
      auto it = known.find(ctx->device);
      if (it != known.end()) return it->second;
      int n = 0;
      if (C > 8 && cudaFuncSetAttribute(k_segment<C, kSegMaxThreads>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
      } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(C), cfg.blockDim = dim3(kSegMaxThreads);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = C, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr, cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, k_segment<C, kSegMaxThreads>, &cfg) != cudaSuccess) cudaGetLastError(), n = 0;
      }
      known[ctx->device] = n;
      return n;
    }
    int active_clusters_of(ofly_ctx *ctx, int C) {
      switch (C) {
        case 16: return active_clusters<16>(ctx);
        case 8: return active_clusters<8>(ctx);
        case 4: return active_clusters<4>(ctx);
        case 2: return active_clusters<2>(ctx);
        default: return 1;
      }
    }

    TeamShape pick_team(ofly_ctx *ctx, long, long n3) {
      TeamShape t{128, 1};
      while (t.threads < kSegMaxThreads && n3 > (long)t.threads * 2) t.threads *= 2;
      if (t.threads == kSegMaxThreads) {
        while (t.cluster < 8 && n3 > (long)t.cluster * t.threads * 8) t.cluster *= 2;
        // 16 CTAs (one GPC's worth; non-portable cluster size) once 8 would leave > 8 elements per thread
        if (t.cluster == 8 && n3 > 8L * t.threads * 8 && env_int("OFLY_CLUSTER16", 1)) t.cluster = 16;
      }
      int ft = env_int("OFLY_TEAM_THREADS", 0);  // tuning overrides
      if (ft == 128 || ft == 256 || ft == 512) t.threads = ft, t.cluster = 1;
      int st = env_int("OFLY_SEG_THREADS", 0);  // narrower CTAs of the same cluster: two CTAs per SM
      if ((st == 128 || st == 256) && st < t.threads) t.threads = st;
      int forced = env_int("OFLY_CLUSTER", 0);
      if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16) t.cluster = forced;
      // step down while the device cannot co-schedule one cluster of this size (the team width is a property of the
      // system size and the device, never of the batch, so results stay independent of the batch composition)
      while (t.cluster > 1 && active_clusters_of(ctx, t.cluster) < 1) t.cluster /= 2;
      return t;
    }

    // One team per replica slot; the teams beyond the list of active replicas leave at once.
    template <int C, int MAXT>
    void launch_segment(ofly_ctx *ctx, long cap, int threads, BatchDev const &B, SearchParams const &P) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)(cap * C));
      cfg.blockDim = dim3((unsigned)threads);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = ctx->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = C, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = C > 1 ? 1 : 0;
      if (C > 8) {  // non-portable cluster size: allowed per function AND per device
        static std::atomic<unsigned long long> allowed{0};
        const unsigned long long bit = 1ull << (ctx->device & 63);
        if (!(allowed.load() & bit)) {
          OFLY_CUDA(cudaFuncSetAttribute(k_segment<C, MAXT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
          allowed.fetch_or(bit);
        }
      }
      OFLY_CUDA(cudaLaunchKernelEx(&cfg, k_segment<C, MAXT>, B, P));
    }

    ofly_ctx *ctx_of(ofly_batch *b) { return b->ctx; }
    unsigned stride_grid(ofly_ctx *ctx, size_t items, int per_sm) {
      return (unsigned)std::max<size_t>(1, std::min<size_t>(items, (size_t)ctx->sm_count * per_sm));
    }

    // One tick = segment, rebuild of the flagged replicas, gradient of the active ones (search.cuh), enqueued in
    // three parts so that the device-side loop can put the rebuild kernels behind an IF node.
    // `cap`: an upper bound on the replicas still active -- every launch grid of a tick is sized for it.
    void enqueue_segment(ofly_batch *b, SearchParams const &P, TeamShape shape, long cap) {
      ofly_ctx *ctx = b->ctx;
      BatchDev const &B = b->d;
      switch (shape.cluster) {
        case 16:
          launch_segment<16, kSegMaxThreads>(ctx, cap, shape.threads, B, P);
          break;
        case 8:
          launch_segment<8, kSegMaxThreads>(ctx, cap, shape.threads, B, P);
          break;
        case 4:
          launch_segment<4, kSegMaxThreads>(ctx, cap, shape.threads, B, P);
          break;
        case 2:
          launch_segment<2, kSegMaxThreads>(ctx, cap, shape.threads, B, P);
          break;
        default:
          if (shape.threads <= 128) {
            launch_segment<1, 128>(ctx, cap, shape.threads, B, P);
          } else {
            launch_segment<1, kSegMaxThreads>(ctx, cap, shape.threads, B, P);
          }
          break;
      }
    }

    // Sub-warp width of the pair kernels: chosen on the device, every tick, from the number of replicas selected.
    // OFLY_GROUP pins it for tuning runs.
    struct PairGrid {
      int forced;       // 0: dynamic
      unsigned blocks;  // grid-stride launch: enough blocks for the widest groups
      long slots;       // resident threads of the chip
    };
    PairGrid pair_grid(ofly_batch *b, long cap) {
      int forced = env_int("OFLY_GROUP", 0);
      if (!(forced == 4 || forced == 8 || forced == 16)) forced = 0;
      const int G = forced ? forced : 16;
      const size_t tiles = (size_t)cap * ((b->N + 256 / G - 1) / (256 / G));
      return {forced, stride_grid(b->ctx, tiles, env_int("OFLY_PAIR_CTAS", 32)), (long)b->ctx->sm_count * 2048};
    }

    void enqueue_rebuild(ofly_batch *b, SearchParams const &P, long cap) {
      ofly_ctx *ctx = b->ctx;
      cudaStream_t st = ctx->stream;
      BatchDev const &B = b->d;
      const size_t atoms = (size_t)cap * b->N;
      Sel reb{P.rebuild, &P.ctl->n_rebuild};
      k_nb_canon_count<<<stride_grid(ctx, (atoms + 255) / 256, 8), 256, 0, st>>>(B, reb, P.ws, P.ws_stride);
      k_nb_scan<<<stride_grid(ctx, (size_t)cap, 2), 1024, 0, st>>>(B, reb);
      k_nb_scatter<<<stride_grid(ctx, (atoms + 255) / 256, 8), 256, 0, st>>>(B, reb);
      k_nb_sort_cells<<<stride_grid(ctx, ((size_t)cap * B.grid.ncells + 127) / 128, 16), 128, 0, st>>>(B, reb);
      if (B.tile.n_tiles > 0) {  // cell-tiled build (tile.cuh): same rows, candidates tested from shared memory
        OFLY_CUDA(tile_prepare());
        tile_launch_build(B, reb, stride_grid(ctx, (size_t)cap * (size_t)B.tile.n_tiles, env_int("OFLY_TILE_CTAS", 12)), st);
        launch_order_rows(B, reb, stride_grid(ctx, (atoms + 7) / 8, 8), st);
        OFLY_CUDA(cudaGetLastError());
        return;
      }
      const PairGrid pg = pair_grid(b, cap);
      switch (pg.forced) {
        case 16:
          k_nb_build<16><<<pg.blocks, 256, 0, st>>>(B, reb);
          break;
        case 8:
          k_nb_build<8><<<pg.blocks, 256, 0, st>>>(B, reb);
          break;
        case 4:
          k_nb_build<4><<<pg.blocks, 256, 0, st>>>(B, reb);
          break;
        default:
          k_nb_build_dyn<<<pg.blocks, 256, 0, st>>>(B, reb, 4 * pg.slots);  // budget: four waves of resident threads
          break;
      }
      launch_order_rows(B, reb, stride_grid(ctx, (atoms + 7) / 8, 8), st);
      OFLY_CUDA(cudaGetLastError());
    }

    // Co-resident CTAs of the fused rebuild kernel on this device (0: cooperative launch unavailable).
    int fused_rebuild_blocks(ofly_ctx *ctx) {
      static std::map<int, int> known;
      static std::mutex known_mutex;
      std::lock_guard<std::mutex> lock(known_mutex);
      auto it = known.find(ctx->device);
      if (it != known.end()) return it->second;
      int coop = 0, per_sm = 0;
      if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device) != cudaSuccess) cudaGetLastError(), coop = 0;
      if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_nb_rebuild_fused, 256, 0) != cudaSuccess) cudaGetLastError(), per_sm = 0;
      return known[ctx->device] = coop ? per_sm * ctx->sm_count : 0;
    }

    // The rebuild of a tick as one cooperative kernel (neigh.cuh, k_nb_rebuild_fused): row kernels only.
    void enqueue_rebuild_fused(ofly_batch *b, SearchParams const &P, long cap) {
      ofly_ctx *ctx = b->ctx;
      BatchDev const &B = b->d;
      Sel reb{P.rebuild, &P.ctl->n_rebuild};
      const long slots = (long)ctx->sm_count * 2048;
      const size_t want = ((size_t)cap * b->N * 32 + 255) / 256;  // a 32-lane group per atom
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)std::max<size_t>(1, std::min<size_t>(want, (size_t)fused_rebuild_blocks(ctx))));
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = ctx->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeCooperative;
      attr[0].val.cooperative = 1;
      cfg.attrs = attr, cfg.numAttrs = 1;
      const double *pos = P.ws;
      OFLY_CUDA(cudaLaunchKernelEx(&cfg, k_nb_rebuild_fused, B, reb, pos, P.ws_stride, 4 * slots));
    }

    void enqueue_gradient(ofly_batch *b, SearchParams const &P, long cap) {
      cudaStream_t st = b->ctx->stream;
      BatchDev const &B = b->d;
      TablesDev const &T = b->tables->dev;
      Sel act{P.active, &P.ctl->n_active};
      const unsigned long long *goff = &P.state->gdst;
      const size_t goff_stride = sizeof(RepState) / sizeof(unsigned long long);
      if (B.tile.n_tiles > 0) {  // cell-tiled passes (tile.cuh): same bits as the row kernels
        OFLY_CUDA(tile_prepare());
        tile_launch_gradient(B, T, act, stride_grid(ctx_of(b), (size_t)cap * (size_t)B.tile.n_tiles, env_int("OFLY_TILE_CTAS", 12)), st, P.ws,
                             P.ws_stride, goff, goff_stride);
        OFLY_CUDA(cudaGetLastError());
        return;
      }
      const PairGrid pg = pair_grid(b, cap);
      switch (pg.forced) {
        case 16:
          k_eam_rho<16><<<pg.blocks, 256, 0, st>>>(B, T, act);
          k_eam_force<16><<<pg.blocks, 256, 0, st>>>(B, T, act, P.ws, P.ws_stride, goff, goff_stride);
          break;
        case 8:
          k_eam_rho<8><<<pg.blocks, 256, 0, st>>>(B, T, act);
          k_eam_force<8><<<pg.blocks, 256, 0, st>>>(B, T, act, P.ws, P.ws_stride, goff, goff_stride);
          break;
        case 4:
          k_eam_rho<4><<<pg.blocks, 256, 0, st>>>(B, T, act);
          k_eam_force<4><<<pg.blocks, 256, 0, st>>>(B, T, act, P.ws, P.ws_stride, goff, goff_stride);
          break;
        default:
          k_eam_rho_dyn<<<pg.blocks, 256, 0, st>>>(B, T, act, pg.slots, P.trace, P.trace_cap);
          k_eam_force_dyn<<<pg.blocks, 256, 0, st>>>(B, T, act, P.ws, P.ws_stride, goff, goff_stride, pg.slots, P.trace, P.trace_cap);
          break;
      }
      OFLY_CUDA(cudaGetLastError());
    }

    void enqueue_tick(ofly_batch *b, SearchParams const &P, TeamShape shape, long cap) {
      enqueue_segment(b, P, shape, cap);
      enqueue_rebuild(b, P, cap);
      enqueue_gradient(b, P, cap);
    }

    // Ends a capture that an exception interrupted, then rethrows.
    struct CaptureScope {
      cudaStream_t st;
      bool open = true;
      explicit CaptureScope(cudaStream_t s) : st(s) {}
      void end() {
        open = false;
        OFLY_CUDA(cudaStreamEndCapture(st, nullptr));
      }
      ~CaptureScope() {
        if (open) {
          cudaGraph_t dead = nullptr;
          cudaStreamEndCapture(st, &dead);
        }
      }
    };

    // Ticks per pass of the WHILE body. The loop-back of a conditional node costs about as much as two kernel boundaries
    // (9 us of a 72 us tick with one small replica left, profiles/r3h_tick_anatomy_*): a body of several ticks pays it once.
    // The ticks after the one that ends the loop find an empty list and leave at once.
    int tick_unroll() { return std::max(1, std::min(kMaxTickUnroll, env_int("OFLY_TICK_UNROLL", 4))); }

    // Everything a captured tick depends on: the kernels' by-value arguments, the launch shapes and the graph's form.
    std::vector<char> graph_key(ofly_batch *b, SearchParams const &P, TeamShape shape, bool nested, bool fused, long cap) {
      const PairGrid pg = pair_grid(b, cap);
      std::vector<char> k;
      auto put = [&k](const void *p, size_t n) { k.insert(k.end(), (const char *)p, (const char *)p + n); };
      put(&P, sizeof P), put(&b->d, sizeof b->d), put(&b->tables->dev, sizeof b->tables->dev);
      put(&shape, sizeof shape), put(&pg, sizeof pg), put(&nested, sizeof nested), put(&cap, sizeof cap);
      const int unroll = tick_unroll();
      put(&unroll, sizeof unroll), put(&fused, sizeof fused);
      return k;
    }

    // The whole search as ONE graph launch: a WHILE conditional node whose body is one tick; the segment
    // kernel clears the condition when no replica is active. With `nested` the six rebuild kernels sit in
    // an IF node inside the body that the segment kernel arms only when some replica asked for a rebuild
    // (late in a batch most ticks have none, and five empty launches are a sixth of such a tick).
    // The instantiated graph is kept in the batch and reused while the key above is unchanged (a Master calls with the
    // same buffers batch after batch); the conditional handles are part of the kernels' arguments, so they are kept too.
    // Returns false if the graph could not be constructed (nothing was launched).
    bool try_device_loop(ofly_batch *b, SearchParams P, TeamShape shape, bool nested, bool fused, long grid_cap, int level) {
      ofly_ctx *ctx = b->ctx;
      ofly_search_graph &g = b->search_graph[level];
      P.use_cond = 0, P.use_cond_rebuild = 0, P.cond = cudaGraphConditionalHandle{}, P.cond_rebuild = cudaGraphConditionalHandle{};
      P.fused_rebuild = fused ? 1 : 0;
      P.stage = level, P.body_slot = 0;
      std::vector<char> key = graph_key(b, P, shape, nested, fused, grid_cap);
      const bool reuse = g.exec != nullptr && g.key == key && !std::getenv("OFLY_NO_GRAPH_CACHE");
      if (!reuse) {
        g.reset();
        try {
          OFLY_CUDA(cudaGraphCreate(&g.graph, 0));
          OFLY_CUDA(cudaGraphConditionalHandleCreate(&g.cond_while, g.graph, 1, cudaGraphCondAssignDefault));
          cudaGraphNodeParams np{};
          np.type = cudaGraphNodeTypeConditional;
          np.conditional.handle = g.cond_while;
          np.conditional.type = cudaGraphCondTypeWhile;
          np.conditional.size = 1;
          cudaGraphNode_t node;
          OFLY_CUDA(cudaGraphAddNode(&node, g.graph, nullptr, 0, &np));
          cudaGraph_t body = np.conditional.phGraph_out[0];
          P.use_cond = 1, P.cond = g.cond_while;
          const int unroll = tick_unroll();
          cudaGraph_t if_body[kMaxTickUnroll] = {};
          SearchParams Pu[kMaxTickUnroll];
          for (int u = 0; u < unroll; u++) {
            Pu[u] = P;
            Pu[u].body_slot = u;
            if (nested) {  // a handle belongs to one conditional node: one per tick of the body
              OFLY_CUDA(cudaGraphConditionalHandleCreate(&g.cond_if[u], g.graph, 0, cudaGraphCondAssignDefault));
              Pu[u].use_cond_rebuild = 1, Pu[u].cond_rebuild = g.cond_if[u];
            }
          }
          {
            OFLY_CUDA(cudaStreamBeginCaptureToGraph(ctx->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
            CaptureScope cap(ctx->stream);
            for (int u = 0; u < unroll; u++) {
              enqueue_segment(b, Pu[u], shape, grid_cap);
              if (nested) {
                cudaStreamCaptureStatus status;
                const cudaGraphNode_t *deps = nullptr;
                size_t n_deps = 0;
                OFLY_CUDA(cudaStreamGetCaptureInfo(ctx->stream, &status, nullptr, nullptr, &deps, &n_deps));
                cudaGraphNodeParams ip{};
                ip.type = cudaGraphNodeTypeConditional;
                ip.conditional.handle = g.cond_if[u];
                ip.conditional.type = cudaGraphCondTypeIf;
                ip.conditional.size = 1;
                cudaGraphNode_t if_node;
                OFLY_CUDA(cudaGraphAddNode(&if_node, body, deps, n_deps, &ip));
                if_body[u] = ip.conditional.phGraph_out[0];
                OFLY_CUDA(cudaStreamUpdateCaptureDependencies(ctx->stream, &if_node, 1, cudaStreamSetCaptureDependencies));
              } else if (fused) {
                enqueue_rebuild_fused(b, Pu[u], grid_cap);
              } else {
                enqueue_rebuild(b, Pu[u], grid_cap);
              }
              enqueue_gradient(b, Pu[u], grid_cap);
            }
            cap.end();
          }
          for (int u = 0; nested && u < unroll; u++) {
            OFLY_CUDA(cudaStreamBeginCaptureToGraph(ctx->stream, if_body[u], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
            CaptureScope cap(ctx->stream);
            enqueue_rebuild(b, Pu[u], grid_cap);
            cap.end();
          }
          OFLY_CUDA(cudaGraphInstantiate(&g.exec, g.graph, 0));
          g.key = std::move(key), g.nested = nested;
        } catch (Error const &e) {
          cudaGetLastError();
          g.reset();
          if (!nested && !fused) throw;
          if (std::getenv("OFLY_VERBOSE")) std::fprintf(stderr, "openfly_b200: %s unavailable (%s); next form of the tick graph\n", fused ? "fused cooperative rebuild" : "nested rebuild IF node", e.what());
          return false;
        }
        if (std::getenv("OFLY_VERBOSE")) std::fprintf(stderr, "openfly_b200: search graph with %s rebuild\n", fused ? "fused cooperative" : (nested ? "IF-guarded" : "unconditional"));
      }
      OFLY_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
      OFLY_CUDA(cudaStreamSynchronize(ctx->stream));
      return true;
    }

    // The device-side loop in up to three stages with shrinking launch grids. A tick's kernels are launched for an upper bound
    // on the active replicas (teams and blocks beyond the work leave at once, but they still have to be launched: with 1024
    // slots, 35 of the 72 us of a late tick with one 432-atom replica left were spent between the kernels, profiles/). A graph
    // cannot change its grids, so a stage's loop ends once the active count has fallen to the next stage's bound and the host
    // launches the next, smaller graph (a few tens of microseconds, twice per call). Results do not depend on the staging.
    void run_device_loop(ofly_batch *b, SearchParams P, TeamShape shape, long n0) {
      ofly_ctx *ctx = b->ctx;
      char const *v = std::getenv("OFLY_REBUILD_IF");
      bool nested = !(v && std::string(v) == "0");
      // one cooperative rebuild kernel instead of the IF node: opt-in (OFLY_REBUILD_FUSED=1, row kernels only). Measured: a late
      // tick of a 432-atom replica 71.1 -> 69.6 us, but a tick with 32 of 64 16 000-atom replicas busy 814 -> 882 us (the
      // cooperative grid holds 4 CTAs per SM where the separate build kernel runs at its own occupancy).
      bool fused = env_int("OFLY_REBUILD_FUSED", 0) && b->d.tile.n_tiles == 0 && fused_rebuild_blocks(ctx) > 0;
      long caps[3] = {n0, 0, 0};
      int levels = 1;
      if (env_int("OFLY_STAGES", 1)) {
        for (; levels < 3 && caps[levels - 1] >= 32; levels++) caps[levels] = (caps[levels - 1] + 7) / 8;
      }
      long active = n0;
      for (int L = 0; L < levels && active > 0; L++) {
        if (L + 1 < levels && active <= caps[L + 1]) continue;  // already small enough for the next stage
        P.exit_below = L + 1 < levels ? (int)caps[L + 1] : 0;
        if (fused && !try_device_loop(b, P, shape, false, true, caps[L], L)) fused = false;
        if (!fused && !(nested && try_device_loop(b, P, shape, true, false, caps[L], L))) {
          nested = false;
          try_device_loop(b, P, shape, false, false, caps[L], L);
        }
        Control c{};
        OFLY_CUDA(cudaMemcpyAsync(&c, P.ctl, sizeof(Control), cudaMemcpyDeviceToHost, ctx->stream));
        OFLY_CUDA(cudaStreamSynchronize(ctx->stream));
        active = c.n_active;
        if (c.tick > P.max_ticks) break;
      }
    }

    // Fallback / profiling mode (OFLY_SEARCH_LOOP=host): plain launches, the host looks at the active
    // count every kHostLoopTicks ticks.
    void run_host_loop(ofly_batch *b, SearchParams P, TeamShape shape) {
      ofly_ctx *ctx = b->ctx;
      P.use_cond = 0;
      Control c{};
      for (;;) {
        for (int t = 0; t < kHostLoopTicks; t++) enqueue_tick(b, P, shape, b->R);
        OFLY_CUDA(cudaMemcpyAsync(&c, P.ctl, sizeof(Control), cudaMemcpyDeviceToHost, ctx->stream));
        OFLY_CUDA(cudaStreamSynchronize(ctx->stream));
        if (c.n_active == 0 || c.tick > P.max_ticks + 1) break;
      }
    }

    struct DevPtrs {
      const double *pos_in, *axis_in, *in_min, *hist_sp;
      double *pos_out, *axis_out, *grad_out;
      int *status;
      long *counters;
    };

    // Shared driver: grid set-up, workspace, launch, capacity-fault retry. All pointers are device memory.
    // `n_queue` searches run through the batch's R resident slots (n_queue == R: the reference's batch call).
    void run_search(ofly_batch *b, SearchParams P, double base_radius, double skin_frac, DevPtrs io, long n_queue) {
      ofly_ctx *ctx = b->ctx;
      OFLY_CUDA(cudaSetDevice(ctx->device));
      require(P.n_hist >= 1 && P.n_hist <= kMaxHist, "L-BFGS history size must be in [1, 16]");
      require(P.mode == 0 || (P.rot_n >= 1 && P.rot_n <= kMaxHist), "rotor L-BFGS history size must be in [1, 16]");
      require(P.iter_max >= 0, "iteration limit is negative");
      require(n_queue >= 1 && n_queue < (1L << 30), "number of searches must be in [1, 2^30)");
      P.n_queue = n_queue;

      // skin = max(skin_frac^(1/3) - 1, 0) * r   (lbfgs.cpp:39, dimer.cpp:47)
      double skin = std::max(std::pow(skin_frac, 1. / 3.) - 1, 0.0) * base_radius;
      double r_list = base_radius + skin;
      P.skin = skin;
      if (!(b->r_list == r_list)) make_grid_for(b, r_list);
      if (b->d.tile.n_tiles > 0) OFLY_CUDA(tile_prepare());  // before any capture starts

      const size_t R = (size_t)b->R;
      const size_t n3p = ((size_t)3 * b->N + 3) & ~(size_t)3;
      const TeamShape shape = pick_team(ctx, b->R, 3 * b->N);
      P.ws_stride = search_ws_doubles(P.mode, P.n_hist, P.rot_n, n3p);

      // arena: control block | replica states | two selection lists + entry flags | vectors
      static_assert(sizeof(RepState) % 8 == 0 && sizeof(Control) % 8 == 0, "layout");
      const size_t head = 256 + ((R * sizeof(RepState) + 3 * R * sizeof(int) + 255) & ~(size_t)255);
      b->search_ws.reserve(head + P.ws_stride * R * sizeof(double));
      char *base = b->search_ws.p;
      P.ctl = reinterpret_cast<Control *>(base);
      P.state = reinterpret_cast<RepState *>(base + 256);
      P.active = reinterpret_cast<int *>(base + 256 + R * sizeof(RepState));
      P.rebuild = P.active + R;
      P.flags = P.rebuild + R;
      P.ws = reinterpret_cast<double *>(base + head);
      P.pos_in = io.pos_in, P.axis_in = io.axis_in, P.in_min = io.in_min, P.hist_sp = io.hist_sp;
      P.pos_out = io.pos_out, P.axis_out = io.axis_out, P.grad_out = io.grad_out, P.status = io.status, P.counters = io.counters;
      // gradient requests per replica are bounded by steps x (2 + rotations); the budget only guards
      // against a device-side loop that would otherwise never end.
      const long per_step = P.mode == 0 ? 1 : 2 + 10L * std::max(1, P.iter_max_rot) + 2;
      const long n0 = std::min<long>(b->R, n_queue);
      const long rounds = (n_queue + n0 - 1) / n0;  // searches a slot may run one after the other
      P.max_ticks = (int)std::min<long>(((long)(P.iter_max + 1) * per_step + 8) * rounds, 1L << 30);

      // The L-BFGS step is the literal two-loop recursion of lbfgs_core.cpp:29-87 (one pass per history pair). OFLY_LBFGS=gram
      // selects the two-pass Gram form (newton_step_gram): 5-10 % faster per step, but its alpha / beta come out of differences
      // of large dot products and are rounded ~1e-12 apart from the recursion's -- on 256 searches against the CPU oracle
      // 5 trajectories branch to another saddle with it (the earliest after 317 evaluations) against 2 (after 1336) with the
      // literal order (scripts/lbfgs_branch_study.py), so the reference's arithmetic is the default.
      P.gram = [] {
        char const *v = std::getenv("OFLY_LBFGS");
        return (v && std::string(v) == "gram") ? 1 : 0;
      }();
      const bool device_loop = [] {
        char const *v = std::getenv("OFLY_SEARCH_LOOP");
        return !(v && std::string(v) == "host");
      }();

      // OFLY_TRACE=<replica>: time stamps inside the segment kernel for that replica, summarised on stderr.
      DevBuf<unsigned long long> trace;
      const int trace_cap = 1 << 19;
      if (char const *tv = std::getenv("OFLY_TRACE")) {
        trace.reserve(1 + 2 * (size_t)trace_cap);
        OFLY_CUDA(cudaMemsetAsync(trace.p, 0, 8, ctx->stream));
        P.trace = trace.p, P.trace_replica = std::atoi(tv), P.trace_cap = trace_cap;
      }
      struct TraceDump {
        DevBuf<unsigned long long> &t;
        int cap;
        ~TraceDump() {
          if (!t.p) return;
          std::vector<unsigned long long> h(1 + 2 * (size_t)cap);
          if (cudaMemcpy(h.data(), t.p, h.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
            size_t n = std::min<size_t>(h[0], (size_t)cap);
            std::map<std::pair<int, int>, std::pair<double, long>> acc;
            for (size_t k = 1; k < n; k++) {
              int a = (int)h[1 + 2 * (k - 1)], b = (int)h[1 + 2 * k];
              double dt = (double)(h[2 + 2 * k] - h[2 + 2 * (k - 1)]);
              auto &e = acc[{a, b}];
              e.first += dt, e.second++;
            }
            std::fprintf(stderr, "openfly_b200 trace: %zu marks\n", n);
            for (auto const &kv : acc)
              std::fprintf(stderr, "  %2d -> %2d : n=%7ld mean %9.1f ns total %9.3f ms\n", kv.first.first, kv.first.second, kv.second.second,
                           kv.second.first / kv.second.second, kv.second.first * 1e-6);
          }
          t.release();
        }
      } trace_dump{trace, trace_cap};

      for (int attempt = 0;; attempt++) {
        OFLY_CUDA(cudaMemsetAsync(base, 0, head, ctx->stream));
        // tick 0: the first n0 searches of the queue are active, one per slot
        k_search_init<<<(unsigned)std::min<long>((n_queue + 255) / 256, 1024), 256, 0, ctx->stream>>>(P, (int)n0);
        OFLY_CUDA(cudaGetLastError());
        if (device_loop) {
          run_device_loop(b, P, shape, n0);
        } else {
          run_host_loop(b, P, shape);
        }
        Control c{};
        OFLY_CUDA(cudaMemcpyAsync(&c, P.ctl, sizeof(Control), cudaMemcpyDeviceToHost, ctx->stream));
        b->built = true;
        // The fault words double as the capacity report; reading them is the call's synchronisation.
        int f = collect_faults(b);
        ctx->launches += (long)c.kernels;  // kernels executed (by the device-side loop or the host loop)
        ctx->ticks += c.tick;
        if (c.n_active != 0) throw Error(OFLY_ERR_CUDA, "search loop ended with active replicas");
        f |= c.fault_or;
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

    // Host-pointer staging: the k-th buffer a call asks for is the k-th buffer of the batch's pool (grown on demand,
    // kept until the batch goes), so a search call through host pointers does no cudaMalloc / cudaFree of its own.
    struct Arena {
      ofly_batch *b;
      size_t next = 0;
      explicit Arena(ofly_batch *batch) : b(batch) {}
      template <typename T>
      T *alloc(size_t n) {
        if (next == b->io_pool.size()) b->io_pool.emplace_back();
        DevBuf<char> &buf = b->io_pool[next++];
        buf.reserve(std::max<size_t>(n, 1) * sizeof(T));
        return reinterpret_cast<T *>(buf.p);
      }
      template <typename T>
      T *upload(const T *host, size_t n) {
        T *d = alloc<T>(n);
        if (n) OFLY_CUDA(cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, b->ctx->stream));
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

int ofly_lbfgs_minimise_stream_dev(ofly_batch *b, const ofly_lbfgs_options *opt, long n_searches, const double *pos_in_dev, double *pos_out_dev,
                                   double *grad_out_dev, int *status_dev, long *counters_dev) {
  return guarded([&] {
    require(b && opt && pos_in_dev && pos_out_dev && status_dev, "null argument");
    DevPtrs io{pos_in_dev, nullptr, nullptr, nullptr, pos_out_dev, nullptr, grad_out_dev, status_dev, counters_dev};
    run_search(b, lbfgs_params(*opt), b->tables->host.r_cut, opt->skin_frac, io, n_searches);
  });
}

int ofly_lbfgs_minimise_stream(ofly_batch *b, const ofly_lbfgs_options *opt, long n_searches, const double *pos_in_host, double *pos_out_host,
                               double *grad_out_host, int *status_host, long *counters_host) {
  return guarded([&] {
    require(b && opt && pos_in_host && pos_out_host && status_host, "null argument");
    require(n_searches >= 1, "number of searches must be positive");
    OFLY_CUDA(cudaSetDevice(b->ctx->device));
    size_t nv = (size_t)n_searches * 3 * (size_t)b->N, Q = (size_t)n_searches;
    Arena A(b);
    DevPtrs io{};
    io.pos_in = A.upload(pos_in_host, nv);
    io.pos_out = A.alloc<double>(nv);
    io.grad_out = A.alloc<double>(nv);
    io.status = A.alloc<int>(Q);
    io.counters = A.alloc<long>(4 * Q);
    run_search(b, lbfgs_params(*opt), b->tables->host.r_cut, opt->skin_frac, io, n_searches);
    cudaStream_t st = b->ctx->stream;
    OFLY_CUDA(cudaMemcpyAsync(pos_out_host, io.pos_out, nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (grad_out_host) OFLY_CUDA(cudaMemcpyAsync(grad_out_host, io.grad_out, nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaMemcpyAsync(status_host, io.status, Q * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (counters_host) OFLY_CUDA(cudaMemcpyAsync(counters_host, io.counters, 4 * Q * sizeof(long), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaStreamSynchronize(st));
  });
}

int ofly_lbfgs_minimise_dev(ofly_batch *b, const ofly_lbfgs_options *opt, const double *pos_in_dev, double *pos_out_dev,
                            double *grad_out_dev, int *status_dev, long *counters_dev) {
  return ofly_lbfgs_minimise_stream_dev(b, opt, b ? b->R : 0, pos_in_dev, pos_out_dev, grad_out_dev, status_dev, counters_dev);
}

int ofly_lbfgs_minimise(ofly_batch *b, const ofly_lbfgs_options *opt, const double *pos_in_host, double *pos_out_host,
                        double *grad_out_host, int *status_host, long *counters_host) {
  return ofly_lbfgs_minimise_stream(b, opt, b ? b->R : 0, pos_in_host, pos_out_host, grad_out_host, status_host, counters_host);
}

int ofly_dimer_find_sp_stream_dev(ofly_batch *b, const ofly_dimer_options *opt, const ofly_rotor_options *ropt, long n_searches,
                                  const double *pos_in_dev, const double *axis_in_dev, const double *in_min_dev, const double *hist_sp_dev,
                                  long n_hist, double theta_tol, int count_frozen, double *pos_out_dev, double *axis_out_dev,
                                  int *exit_code_dev, long *counters_dev) {
  return guarded([&] {
    require(b && opt && ropt && pos_in_dev && axis_in_dev && pos_out_dev && axis_out_dev && exit_code_dev, "null argument");
    require(n_hist == 0 || (in_min_dev && hist_sp_dev), "history given without in_min / hist_sp");
    DevPtrs io{pos_in_dev, axis_in_dev, in_min_dev, hist_sp_dev, pos_out_dev, axis_out_dev, nullptr, exit_code_dev, counters_dev};
    // list radius: rotor.r_cut(pot) = pot.r_cut + 2 delta_r (rotor.hpp:123) plus the skin (dimer.cpp:47-53)
    run_search(b, dimer_params(*opt, *ropt, theta_tol, count_frozen, n_hist), b->tables->host.r_cut + 2 * ropt->delta_r, opt->skin_frac, io,
               n_searches);
  });
}

int ofly_dimer_find_sp_stream(ofly_batch *b, const ofly_dimer_options *opt, const ofly_rotor_options *ropt, long n_searches,
                              const double *pos_in_host, const double *axis_in_host, const double *in_min_host, const double *hist_sp_host,
                              long n_hist, double theta_tol, int count_frozen, double *pos_out_host, double *axis_out_host,
                              int *exit_code_host, long *counters_host) {
  return guarded([&] {
    require(b && opt && ropt && pos_in_host && axis_in_host && pos_out_host && axis_out_host && exit_code_host, "null argument");
    require(n_hist == 0 || (in_min_host && hist_sp_host), "history given without in_min / hist_sp");
    require(n_searches >= 1, "number of searches must be positive");
    OFLY_CUDA(cudaSetDevice(b->ctx->device));
    size_t n3 = 3 * (size_t)b->N, nv = (size_t)n_searches * n3, Q = (size_t)n_searches;
    Arena A(b);
    DevPtrs io{};
    io.pos_in = A.upload(pos_in_host, nv);
    io.axis_in = A.upload(axis_in_host, nv);
    io.in_min = in_min_host ? A.upload(in_min_host, n3) : nullptr;
    io.hist_sp = n_hist ? A.upload(hist_sp_host, (size_t)n_hist * n3) : nullptr;
    io.pos_out = A.alloc<double>(nv);
    io.axis_out = A.alloc<double>(nv);
    io.status = A.alloc<int>(Q);
    io.counters = A.alloc<long>(4 * Q);
    run_search(b, dimer_params(*opt, *ropt, theta_tol, count_frozen, n_hist), b->tables->host.r_cut + 2 * ropt->delta_r, opt->skin_frac, io,
               n_searches);
    cudaStream_t st = b->ctx->stream;
    OFLY_CUDA(cudaMemcpyAsync(pos_out_host, io.pos_out, nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaMemcpyAsync(axis_out_host, io.axis_out, nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaMemcpyAsync(exit_code_host, io.status, Q * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (counters_host) OFLY_CUDA(cudaMemcpyAsync(counters_host, io.counters, 4 * Q * sizeof(long), cudaMemcpyDeviceToHost, st));
    OFLY_CUDA(cudaStreamSynchronize(st));
  });
}

int ofly_dimer_find_sp_dev(ofly_batch *b, const ofly_dimer_options *opt, const ofly_rotor_options *ropt, const double *pos_in_dev,
                           const double *axis_in_dev, const double *in_min_dev, const double *hist_sp_dev, long n_hist,
                           double theta_tol, int count_frozen, double *pos_out_dev, double *axis_out_dev, int *exit_code_dev,
                           long *counters_dev) {
  return ofly_dimer_find_sp_stream_dev(b, opt, ropt, b ? b->R : 0, pos_in_dev, axis_in_dev, in_min_dev, hist_sp_dev, n_hist, theta_tol,
                                       count_frozen, pos_out_dev, axis_out_dev, exit_code_dev, counters_dev);
}

int ofly_dimer_find_sp(ofly_batch *b, const ofly_dimer_options *opt, const ofly_rotor_options *ropt, const double *pos_in_host,
                       const double *axis_in_host, const double *in_min_host, const double *hist_sp_host, long n_hist,
                       double theta_tol, int count_frozen, double *pos_out_host, double *axis_out_host, int *exit_code_host,
                       long *counters_host) {
  return ofly_dimer_find_sp_stream(b, opt, ropt, b ? b->R : 0, pos_in_host, axis_in_host, in_min_host, hist_sp_host, n_hist, theta_tol,
                                   count_frozen, pos_out_host, axis_out_host, exit_code_host, counters_host);
}

}  // extern "C"
