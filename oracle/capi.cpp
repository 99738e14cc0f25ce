// SPDX-License-Identifier: GPL-3.0-or-later
// A restatement of algorithms of openFLY (ConorWilliams/openFLY, GPL-3.0-or-later); see LICENSE.
// TEST INFRASTRUCTURE ONLY -- C entry points over fly_oracle.hpp so that tests/ and bench.py's
// cpu_baseline / --impl reference legs can drive the CPU oracle through ctypes.
// Nothing under openfly_b200/ may load this library.
#include <omp.h>

#include <chrono>
#include <cstring>
#include <string>

#include "fly_oracle.hpp"

using namespace oracle;

namespace {
  thread_local std::string g_err;

  struct System {
    std::shared_ptr<DataEAM const> data;
    std::unique_ptr<EAM> pot;
    Box box;
    std::unique_ptr<List> nl;
  };

  template <typename F>
  int guard(F &&f) {
    try {
      f();
      return 0;
    } catch (OutOfBounds const &e) {
      g_err = e.what();
      return 2;
    } catch (std::exception const &e) {
      g_err = e.what();
      return 1;
    }
  }
}  // namespace

extern "C" {

char const *orc_last_error() { return g_err.c_str(); }

// ---- tables ---------------------------------------------------------------------------------
void *orc_tables_load(char const *path) {
  try {
    return new std::shared_ptr<DataEAM const>(std::make_shared<DataEAM const>(path));
  } catch (std::exception const &e) {
    g_err = e.what();
    return nullptr;
  }
}
void orc_tables_free(void *h) { delete static_cast<std::shared_ptr<DataEAM const> *>(h); }

static Spline const &pick(DataEAM const &d, int kind, int i, int j) {
  switch (kind) {
    case 0:
      return d.f(i);
    case 1:
      return d.phi(i, j);
    default:
      return d.v(i, j);
  }
}

// info: [n_types, numP, numR] ; dinfo: [delP, delR, rcut, mass...]
void orc_tables_info(void *h, long *info, double *dinfo) {
  auto const &d = **static_cast<std::shared_ptr<DataEAM const> *>(h);
  info[0] = static_cast<long>(d.num_types()), info[1] = static_cast<long>(d.numP), info[2] = static_cast<long>(d.numR);
  dinfo[0] = d.delP, dinfo[1] = d.delR, dinfo[2] = d.rcut();
  for (std::size_t i = 0; i < d.num_types(); i++) dinfo[3 + i] = d.mass[i];
}
long orc_tables_num_knots(void *h, int kind, int i, int j) {
  return static_cast<long>(pick(**static_cast<std::shared_ptr<DataEAM const> *>(h), kind, i, j).knots().size());
}
void orc_tables_knots(void *h, int kind, int i, int j, double *out) {
  auto const &k = pick(**static_cast<std::shared_ptr<DataEAM const> *>(h), kind, i, j).knots();
  std::memcpy(out, k.data(), k.size() * sizeof(Spline::Knot));
}
// which: 0 f, 1 fp, 2 fpp
int orc_tables_eval(void *h, int kind, int i, int j, int which, double x, double *out) {
  return guard([&] {
    auto const &s = pick(**static_cast<std::shared_ptr<DataEAM const> *>(h), kind, i, j);
    *out = which == 0 ? s.f(x) : (which == 1 ? s.fp(x) : s.fpp(x));
  });
}

// Stand-alone spline (tests/utility/spline.cpp goldens)
int orc_spline_eval(double const *y, long n, double dx, int which, double const *x, long nx, double *out) {
  return guard([&] {
    Spline s(std::vector<double>(y, y + n), dx);
    for (long i = 0; i < nx; i++) out[i] = which == 0 ? s.f(x[i]) : (which == 1 ? s.fp(x[i]) : s.fpp(x[i]));
  });
}

// ---- geometry -------------------------------------------------------------------------------
int orc_grid_info(double const *extents, double r_cut, int *shape, double *cell) {
  return guard([&] {
    Grid g({extents[0], extents[1], extents[2]}, r_cut);
    for (int i = 0; i < 3; i++) shape[i] = g.shape[i], cell[i] = g.cell[i];
  });
}
int orc_cell_idx(double const *extents, double r_cut, double const *x, long n, int *out) {
  return guard([&] {
    Grid g({extents[0], extents[1], extents[2]}, r_cut);
    for (long i = 0; i < n; i++) out[i] = g.cell_idx({x[3 * i], x[3 * i + 1], x[3 * i + 2]});
  });
}
void orc_canon_image(double const *extents, double const *x, long n, double *out) {
  Box b({extents[0], extents[1], extents[2]}, {true, true, true});
  for (long i = 0; i < n; i++) {
    Vec3 c = b.canon_image({x[3 * i], x[3 * i + 1], x[3 * i + 2]});
    out[3 * i] = c.x, out[3 * i + 1] = c.y, out[3 * i + 2] = c.z;
  }
}
// returns 1 and writes the image if one is generated, else 0
int orc_gen_image(double const *extents, double r_cut, double const *x, int ax, int sign, double *out) {
  Grid g({extents[0], extents[1], extents[2]}, r_cut);
  auto im = g.gen_image({x[0], x[1], x[2]}, ax, sign);
  if (im) out[0] = im->x, out[1] = im->y, out[2] = im->z;
  return im ? 1 : 0;
}
int orc_adjacent(int const *shape, int cell, int *out) {
  AdjacentCells adj({shape[0], shape[1], shape[2]});
  auto const &a = adj.cells[static_cast<std::size_t>(cell)];
  for (std::size_t k = 0; k < a.count; k++) out[k] = a.ids[k];
  return static_cast<int>(a.count);
}

// ---- system: tables + box + neighbour list ----------------------------------------------------
void *orc_system_create(void *tables, double const *extents, int const *periodic) {
  auto *s = new System;
  s->data = *static_cast<std::shared_ptr<DataEAM const> *>(tables);
  s->pot = std::make_unique<EAM>(s->data);
  s->box = Box({extents[0], extents[1], extents[2]}, {periodic[0] != 0, periodic[1] != 0, periodic[2] != 0});
  return s;
}
void orc_system_free(void *h) { delete static_cast<System *>(h); }

int orc_list_rebuild(void *h, double r_cut, double const *pos, long n, int threads) {
  auto *s = static_cast<System *>(h);
  return guard([&] {
    if (!s->nl || s->nl->r_cut() != r_cut) s->nl = std::make_unique<List>(s->box, r_cut);
    s->nl->rebuild(pos, n, threads);
  });
}
int orc_list_update(void *h, double const *delta, long n) {
  auto *s = static_cast<System *>(h);
  return guard([&] { s->nl->update(delta, n); });
}
long orc_list_num_plus_ghosts(void *h) { return static_cast<System *>(h)->nl->num_plus_ghosts(); }

// Raw stored lists (before the r_cut filter of for_neighbours): per atom count, then flattened
// (real index, shift code) pairs where shift code = (sx+1) + 3(sy+1) + 9(sz+1), s = round((ghost - real)/L).
long orc_list_export(void *h, int *counts, int *real_idx, int *shift, long cap) {
  auto *s = static_cast<System *>(h);
  List const &nl = *s->nl;
  long w = 0;
  for (long i = 0; i < nl.size(); i++) {
    auto const &l = nl.raw_list(i);
    counts[i] = static_cast<int>(l.size());
    for (int n : l) {
      if (w < cap) {
        long j = nl.image_to_real(n);
        int code = 0, mul = 1;
        for (int c = 0; c < 3; c++) {
          double d = (nl.position(n)[c] - nl.position(j)[c]) / s->box.extents[c];
          code += (static_cast<int>(std::lround(d)) + 1) * mul;
          mul *= 3;
        }
        real_idx[w] = static_cast<int>(j), shift[w] = code;
      }
      w++;
    }
  }
  return w;
}
// Cell index of every real atom's canonical image in the list's grid.
void orc_list_cells(void *h, int *out) {
  List const &nl = *static_cast<System *>(h)->nl;
  for (long i = 0; i < nl.size(); i++) out[i] = nl.grid().cell_idx(nl.position(i));
}
// for_neighbours(i, r_cut): flattened (real index, r, dr) ; returns count
long orc_list_for_neighbours(void *h, long i, double r_cut, int *idx, double *r, double *dr, long cap) {
  long w = 0;
  static_cast<System *>(h)->nl->for_neighbours(i, r_cut, [&](long n, double rr, Vec3 const &d) {
    if (w < cap) idx[w] = static_cast<int>(n), r[w] = rr, dr[3 * w] = d.x, dr[3 * w + 1] = d.y, dr[3 * w + 2] = d.z;
    w++;
  });
  return w;
}

// ---- EAM --------------------------------------------------------------------------------------
int orc_eam_energy(void *h, std::uint32_t const *id, long n, int threads, double *out) {
  auto *s = static_cast<System *>(h);
  return guard([&] { *out = s->pot->energy(id, n, *s->nl, threads); });
}
int orc_eam_gradient(void *h, std::uint32_t const *id, std::uint8_t const *frozen, long n, int threads, double *out) {
  auto *s = static_cast<System *>(h);
  return guard([&] { s->pot->gradient(out, id, frozen, n, *s->nl, threads); });
}
// out: dense column-major (3n)^2, lower triangle meaningful
int orc_eam_hessian(void *h, std::uint32_t const *id, std::uint8_t const *frozen, long n, int threads, double *out, int mass_weight) {
  auto *s = static_cast<System *>(h);
  return guard([&] {
    Hessian H;
    s->pot->hessian(H, id, frozen, n, *s->nl, threads);
    if (mass_weight) H.mass_weight(s->data->mass, id, n);
    std::memcpy(out, H.h.data(), H.h.size() * sizeof(double));
  });
}

// ---- minimiser / dimer --------------------------------------------------------------------------
// lopt: [n, iter_max] ; dopt: [f2norm, skin_frac, proj_tol, max_trust, min_trust, grow_trust, shrink_trust]
// counters: [grad_evals, rebuilds, updates, steps]. *ret = 0 converged, 1 iter_max (reference bool).
int orc_lbfgs_minimise(void *h, int const *lopt, double const *dopt, double const *in_r, std::uint32_t const *id,
                       std::uint8_t const *frozen, long n, int threads, double *out_r, double *out_g, int *ret, long *counters) {
  auto *s = static_cast<System *>(h);
  return guard([&] {
    LBFGSOptions o;
    if (lopt) o.n = lopt[0], o.iter_max = lopt[1];
    if (dopt)
      o.f2norm = dopt[0], o.skin_frac = dopt[1], o.proj_tol = dopt[2], o.max_trust = dopt[3], o.min_trust = dopt[4],
      o.grow_trust = dopt[5], o.shrink_trust = dopt[6];
    LBFGS m(o, s->box);
    *ret = m.minimise(out_r, out_g, in_r, id, frozen, n, *s->pot, threads) ? 1 : 0;
    if (counters) counters[0] = m.counters.grad_evals, counters[1] = m.counters.rebuilds, counters[2] = m.counters.updates, counters[3] = m.counters.steps;
  });
}

// iopt: [n, max_steps, hist_check_freq, use_history, rot_n, iter_max_rot, relax_in_convex]
// dopt: [f2norm, skin_frac, proj_tol, max_trust, min_trust, grow_trust, shrink_trust, convex_max, delta_r, rot_theta_tol]
int orc_dimer_find_sp(void *h, int const *iopt, double const *dopt, double const *in_r, double const *in_ax, std::uint32_t const *id,
                      std::uint8_t const *frozen, long n, double const *in_min, double const *hist_sp, long n_hist, double theta_tol,
                      int count_frozen, int threads, double *out_r, double *out_ax, int *exit_code, long *counters) {
  auto *s = static_cast<System *>(h);
  return guard([&] {
    DimerOptions o;
    RotorOptions ro;
    if (iopt)
      o.n = iopt[0], o.max_steps = iopt[1], o.hist_check_freq = iopt[2], o.use_history = iopt[3] != 0, ro.n = iopt[4],
      ro.iter_max_rot = iopt[5], ro.relax_in_convex = iopt[6] != 0;
    if (dopt)
      o.f2norm = dopt[0], o.skin_frac = dopt[1], o.proj_tol = dopt[2], o.max_trust = dopt[3], o.min_trust = dopt[4],
      o.grow_trust = dopt[5], o.shrink_trust = dopt[6], o.convex_max = dopt[7], ro.delta_r = dopt[8], ro.theta_tol = dopt[9];
    Dimer d(o, ro, s->box);
    std::vector<Arr> hist;
    for (long k = 0; k < n_hist; k++) hist.emplace_back(hist_sp + 3 * n * k, hist_sp + 3 * n * (k + 1));
    *exit_code = d.find_sp(out_r, out_ax, in_r, in_ax, id, frozen, n, in_min, *s->pot, hist, theta_tol, count_frozen, threads);
    if (counters) counters[0] = d.counters.grad_evals, counters[1] = d.counters.rebuilds, counters[2] = d.counters.updates, counters[3] = d.counters.steps;
  });
}

// Many independent searches, one single-threaded search per OpenMP thread -- mirrors how
// saddle::Master runs its batch (src/saddle/find.cpp:698,990). Arrays are [n_rep][3n].
int orc_dimer_find_sp_batch(void *h, int const *iopt, double const *dopt, double const *in_r, double const *in_ax,
                            std::uint32_t const *id, std::uint8_t const *frozen, long n, long n_rep, double const *in_min,
                            double const *hist_sp, long n_hist, double theta_tol, int count_frozen, int threads, double *out_r,
                            double *out_ax, int *exit_code, long *counters) {
  int bad = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic)
  for (long k = 0; k < n_rep; k++) {
    System local;
    auto *s = static_cast<System *>(h);
    local.data = s->data, local.pot = std::make_unique<EAM>(s->data), local.box = s->box;
    int rc = orc_dimer_find_sp(&local, iopt, dopt, in_r + 3 * n * k, in_ax + 3 * n * k, id, frozen, n, in_min, hist_sp, n_hist,
                               theta_tol, count_frozen, 1, out_r + 3 * n * k, out_ax + 3 * n * k, exit_code + k,
                               counters ? counters + 4 * k : nullptr);
    if (rc == 2) {
      exit_code[k] = -5;  // OutOfBounds (Master re-perturbs, find.cpp:718-722)
    } else if (rc != 0) {
#pragma omp atomic write
      bad = 1;
    }
  }
  return bad;
}
int orc_lbfgs_minimise_batch(void *h, int const *lopt, double const *dopt, double const *in_r, std::uint32_t const *id,
                             std::uint8_t const *frozen, long n, long n_rep, int threads, double *out_r, double *out_g, int *ret,
                             long *counters) {
  int bad = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic)
  for (long k = 0; k < n_rep; k++) {
    System local;
    auto *s = static_cast<System *>(h);
    local.data = s->data, local.pot = std::make_unique<EAM>(s->data), local.box = s->box;
    int rc = orc_lbfgs_minimise(&local, lopt, dopt, in_r + 3 * n * k, id, frozen, n, 1, out_r + 3 * n * k, out_g + 3 * n * k, ret + k,
                                counters ? counters + 4 * k : nullptr);
    if (rc == 2) {
      ret[k] = -5;
    } else if (rc != 0) {
#pragma omp atomic write
      bad = 1;
    }
  }
  return bad;
}

// ---- input builders -----------------------------------------------------------------------------
void orc_bcc_lattice(double a, int nx, int ny, int nz, double *out) {
  Arr r = bcc_lattice(a, nx, ny, nz);
  std::memcpy(out, r.data(), r.size() * sizeof(double));
}
// Gaussian jitter with Xoshiro + std::normal_distribution (SURVEY 8d synthetic inputs)
void orc_jitter(double *r, long n3, std::uint64_t const *seed, double stddev) {
  Xoshiro g({seed[0], seed[1], seed[2], seed[3]});
  std::normal_distribution<double> d(0, stddev);
  for (long i = 0; i < n3; i++) r[i] += d(g);
}
// Choose k distinct indices in [0, n) with Xoshiro (partial Fisher-Yates); sorted ascending on output.
void orc_choose(long n, long k, std::uint64_t const *seed, long *out) {
  Xoshiro g({seed[0], seed[1], seed[2], seed[3]});
  std::vector<long> idx(static_cast<std::size_t>(n));
  for (long i = 0; i < n; i++) idx[i] = i;
  for (long i = 0; i < k; i++) {
    std::uniform_int_distribution<long> d(i, n - 1);
    std::swap(idx[i], idx[d(g)]);
  }
  std::sort(idx.begin(), idx.begin() + k);
  std::copy(idx.begin(), idx.begin() + k, out);
}
// Master::perturb for replica `rep`: seed chain = Xoshiro(seed) then long_jump() x rep.
// The perturbation neighbour list (r_pert) is built on in_r with the system's box.
int orc_perturb(void *h, double const *in_r, std::uint8_t const *frozen, long n, long centre, std::uint64_t const *seed, long rep,
                double r_pert, double stddev, double *out_r, double *out_ax) {
  auto *s = static_cast<System *>(h);
  return guard([&] {
    List nl(s->box, r_pert);
    nl.rebuild(in_r, n, 1);
    Xoshiro g({seed[0], seed[1], seed[2], seed[3]});
    for (long k = 0; k < rep; k++) g.long_jump();
    perturb(out_r, out_ax, in_r, frozen, n, centre, nl, g, r_pert, stddev);
  });
}

// ---- CPU baseline timing loops (bench.py cpu_baseline / --impl reference) -----------------------
// Repeats [update(delta) ; gradient] `iters` times with `threads` OpenMP threads, exactly the
// per-evaluation work of the rotor inner loop (rotor.cpp:77-85). Returns seconds.
double orc_time_gradient(void *h, std::uint32_t const *id, std::uint8_t const *frozen, long n, double const *delta, int threads,
                         int iters, double *out) {
  auto *s = static_cast<System *>(h);
  std::vector<double> neg(delta, delta + 3 * n);
  for (auto &x : neg) x = -x;
  auto t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < iters; it++) {
    s->nl->update(it % 2 == 0 ? delta : neg.data(), n);
    s->pot->gradient(out, id, frozen, n, *s->nl, threads);
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
double orc_time_rebuild(void *h, double r_cut, double const *pos, long n, int threads, int iters) {
  auto *s = static_cast<System *>(h);
  if (!s->nl || s->nl->r_cut() != r_cut) s->nl = std::make_unique<List>(s->box, r_cut);
  auto t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < iters; it++) s->nl->rebuild(pos, n, threads);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
int orc_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
