// SPDX-License-Identifier: GPL-3.0-or-later
// A restatement of algorithms of openFLY (ConorWilliams/openFLY, GPL-3.0-or-later); see LICENSE.
// TEST INFRASTRUCTURE ONLY -- CPU oracle for the openFLY force-evaluation hot path.
//
// This header is an Eigen-free C++17 restatement of the reference's algorithms for the path
// named in BASELINE.json (setfl tables, natural cubic splines, orthorhombic cell grid, ghost-atom
// linked-cell neighbour list, EAM energy/gradient/Hessian, L-BFGS, dimer rotor/translator).
// It exists to CHECK the CUDA product path and to be timed as the CPU baseline. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
// The product (openfly_b200/) never includes, links or calls anything in oracle/.
//
// Parity status: pinned against the reference's own known-answer tests (see tests/test_oracle_*.py):
//   tests/potential/EAM/eam.cpp:119,125,131-141,167,177-183; tests/potential/EAM/data.cpp:42-43;
//   tests/neigh/adjacent.cpp:27-93; tests/system/boxes/hypergrid.cpp:28-44;
//   tests/system/boxes/orthorhombic.cpp:55-99; tests/utility/spline.cpp:29-35; tests/neigh/list.cpp:42-126.
// LBFGS / Rotor / Dimer have NO reference tests: for those the oracle is "parity unpinned"
// (followed statement by statement from the reference sources cited at each function).
//
// Every function cites the reference file:line it follows (paths relative to the openFLY tree).
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <fstream>
#include <functional>
#include <limits>
#include <memory>
#include <optional>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace oracle {

  struct Vec3 {
    double x = 0, y = 0, z = 0;
    double &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  };

  inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
  inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
  inline Vec3 operator*(double s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }

  // include/libfly/utility/core.hpp:655-658 (gnorm_sq). Eigen reduces a fixed 3-array with one
  // 2-wide packet followed by the scalar tail, i.e. (x*x + y*y) + z*z.
  inline double norm_sq(Vec3 a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }

  struct RuntimeError : std::runtime_error {
    using std::runtime_error::runtime_error;
  };

  // include/libfly/utility/spline.hpp:46-51
  struct OutOfBounds : RuntimeError {
    using RuntimeError::RuntimeError;
  };

  // ------------------------------------------------------------------------------------------
  // Natural cubic spline on a uniform grid.
  // ctor: src/utility/spline.cpp:27-103 ; evaluation: include/libfly/utility/spline.hpp:81-154
  // ------------------------------------------------------------------------------------------
  class Spline {
  public:
    struct Knot {
      double a, b, c, d;
    };

    Spline() = default;

    Spline(std::vector<double> y, double dx) : m_dx(dx), m_inv_dx(1 / dx) {
      double last = y.back();  // spline.cpp:29-33 soft padding with 5 copies of the last value
      for (int i = 0; i < 5; i++) {
        y.push_back(last);
      }
      std::size_t n = y.size() - 1;

      std::vector<double> &a = y;
      std::vector<double> b(n), d(n), alpha(n), c(n + 1), l(n + 1), mu(n + 1), z(n + 1);

      for (std::size_t i = 1; i <= n - 1; ++i) {  // spline.cpp:59-61
        alpha[i] = 3 * (a[i + 1] - 2 * a[i] + a[i - 1]) / dx;
      }
      l[0] = 1, mu[0] = 0, z[0] = 0;
      for (std::size_t i = 1; i <= n - 1; ++i) {  // spline.cpp:69-73
        l[i] = dx * (4 - mu[i - 1]);
        mu[i] = dx / l[i];
        z[i] = (alpha[i] - dx * z[i - 1]) / l[i];
      }
      l[n] = 1, z[n] = 0, c[n] = 0;
      for (std::size_t i = 1; i <= n; ++i) {  // spline.cpp:81-87
        std::size_t j = n - i;
        c[j] = z[j] - mu[j] * c[j + 1];
        b[j] = (a[j + 1] - a[j]) / dx - dx * (c[j + 1] + 2 * c[j]) / 3;
        d[j] = (c[j + 1] - c[j]) / (3 * dx);
      }
      m_knots.reserve(n);
      for (std::size_t i = 0; i <= n - 1; ++i) {
        m_knots.push_back({a[i], b[i], c[i], d[i]});
      }
    }

    double f(double x) const {  // spline.hpp:81-84
      auto [x0, k] = fetch(x);
      return k.a + x0 * (k.b + x0 * (k.c + x0 * k.d));
    }
    double fp(double x) const {  // spline.hpp:99-102
      auto [x0, k] = fetch(x);
      return k.b + x0 * (2.0 * k.c + x0 * 3.0 * k.d);
    }
    double fpp(double x) const {  // spline.hpp:120-123
      auto [x0, k] = fetch(x);
      return 2 * k.c + 6.0 * x0 * k.d;
    }

    std::vector<Knot> const &knots() const { return m_knots; }
    double dx() const { return m_dx; }
    double inv_dx() const { return m_inv_dx; }

  private:
    std::vector<Knot> m_knots;
    double m_dx = 0, m_inv_dx = 0;

    std::pair<double, Knot const &> fetch(double x) const {  // spline.hpp:135-154
      auto i = static_cast<std::size_t>(x * m_inv_dx);
      if (i >= m_knots.size()) {
        throw OutOfBounds("x=" + std::to_string(x) + " is outside tabulated region");
      }
      return {x - static_cast<double>(i) * m_dx, m_knots[i]};
    }
  };

  // ------------------------------------------------------------------------------------------
  // LAMMPS setfl eam/fs reader.  src/potential/EAM/data.cpp:67-153 ;
  // index helpers include/libfly/potential/EAM/data.hpp:109-116
  // ------------------------------------------------------------------------------------------
  class DataEAM {
  public:
    explicit DataEAM(std::string const &path) {
      std::ifstream in(path);
      if (!in.good()) {
        throw RuntimeError("Could not open eam file " + path);
      }
      auto getline = [&in]() {
        std::string line;
        if (!std::getline(in, line)) {
          throw std::runtime_error("File terminated too soon");
        }
        return std::istringstream{line};
      };
      auto read_chunked = [&](std::size_t n, std::size_t k) {  // data.cpp:48-65
        std::vector<double> raw;
        std::istringstream line;
        for (std::size_t i = 0; i < n; ++i) {
          if (i % k == 0) {
            line = getline();
          }
          double tmp;
          line >> tmp;
          raw.push_back(tmp);
        }
        return raw;
      };

      for (int i = 0; i < 3; ++i) {
        getline();
      }
      {
        int n;
        auto line = getline();
        line >> n;
        m_n = static_cast<std::size_t>(n);
        names.resize(m_n);
        for (auto &nm : names) {
          line >> nm;
        }
      }
      mass.resize(m_n);
      m_f.resize(m_n);
      m_phi.resize(m_n * m_n);
      m_v.resize(m_n * m_n);

      getline() >> numP >> delP >> numR >> delR >> m_rcut;

      for (std::size_t i = 0; i < m_n; ++i) {
        std::size_t atomic;
        getline() >> atomic >> mass[i];
        m_f[i] = Spline{read_chunked(numP, 5), delP};
        for (std::size_t j = 0; j < m_n; ++j) {
          m_phi[index(i, j)] = Spline{read_chunked(numR, 5), delR};
        }
      }
      for (std::size_t i = 0; i < m_n; ++i) {  // data.cpp:134-152: tabulated as r*V, symmetric
        for (std::size_t j = 0; j <= i; ++j) {
          std::vector<double> raw = read_chunked(numR, 5);
          for (std::size_t k = 0; k < raw.size(); k++) {
            raw[k] /= delR * static_cast<double>(k);
          }
          raw[0] = raw[1];
          m_v[sym_index(i, j)] = Spline{std::move(raw), delR};
        }
      }
    }

    double rcut() const { return m_rcut; }
    std::size_t num_types() const { return m_n; }
    Spline const &f(std::size_t i) const { return m_f[i]; }
    Spline const &phi(std::size_t i, std::size_t j) const { return m_phi[index(i, j)]; }
    Spline const &v(std::size_t i, std::size_t j) const { return m_v[sym_index(i, j)]; }

    std::vector<std::string> names;
    std::vector<double> mass;
    std::size_t numP = 0, numR = 0;
    double delP = 0, delR = 0;

  private:
    std::size_t m_n = 0;
    double m_rcut = 0;
    std::vector<Spline> m_f, m_phi, m_v;

    std::size_t index(std::size_t i, std::size_t j) const { return i + m_n * j; }
    std::size_t sym_index(std::size_t i, std::size_t j) const { return std::max(i, j) + m_n * std::min(i, j); }
  };

  // ------------------------------------------------------------------------------------------
  // Orthorhombic box + cell grid.
  //   Orthorhombic::canon_image  include/libfly/system/boxes/orthorhombic.hpp:132-139
  //   OrthoGrid::gen_image       include/libfly/system/boxes/orthorhombic.hpp:54-74
  //   HyperGrid ctor / cell_idx  include/libfly/system/boxes/hypergrid.hpp:43-99
  //   product_scan               include/libfly/utility/core.hpp:570-578
  // The triclinic alternative of system::Box is out of scope (no BASELINE config uses it).
  // ------------------------------------------------------------------------------------------
  struct Box {
    Vec3 extents;
    std::array<bool, 3> periodic{true, true, true};
    Vec3 inv_extents;

    Box() = default;
    Box(Vec3 ex, std::array<bool, 3> pd) : extents(ex), periodic(pd) {
      for (int i = 0; i < 3; i++) {
        if (!(ex[i] > 0)) {
          throw RuntimeError("Orthorhombic extents are negative");
        }
        inv_extents[i] = 1.0 / ex[i];
      }
    }

    Vec3 canon_image(Vec3 x) const {
      Vec3 out;
      for (int i = 0; i < 3; i++) {
        out[i] = x[i] - extents[i] * std::floor(x[i] * inv_extents[i]);
      }
      return out;
    }
  };

  struct Grid {
    double r_cut = 0;
    std::array<int, 3> shape{}, prod{};
    Vec3 cell, inv_cell, extents;

    Grid() = default;
    Grid(Vec3 ex, double rc) : r_cut(rc), extents(ex) {
      if (!(rc > 0)) {
        throw RuntimeError("r_cut is negative");
      }
      for (int i = 0; i < 3; i++) {
        if (!(ex[i] > rc)) {
          throw RuntimeError("r_cut is too big for box");
        }
        shape[i] = 2 + static_cast<int>(ex[i] / rc);
        cell[i] = ex[i] / std::floor(ex[i] / rc);
        inv_cell[i] = 1.0 / cell[i];
      }
      int p = 1;
      for (int i = 0; i < 3; i++) {
        prod[i] = p;
        p *= shape[i];
      }
    }

    int num_cells() const { return shape[0] * shape[1] * shape[2]; }

    int cell_idx(Vec3 x) const {  // hypergrid.hpp:75-99
      int idx = 0;
      for (int i = 0; i < 3; i++) {
        int k = static_cast<int>((x[i] + cell[i]) * inv_cell[i]);
        k = std::min(std::max(k, 0), shape[i] - 1);
        idx += k * prod[i];
      }
      return idx;
    }

    // sign = +1: image at x + L if x < r_cut ; sign = -1: image at x - L if x > L - r_cut
    std::optional<Vec3> gen_image(Vec3 x, int ax, int sign) const {
      if (sign > 0) {
        if (x[ax] < r_cut) {
          x[ax] += extents[ax];
          return x;
        }
      } else {
        if (x[ax] > extents[ax] - r_cut) {
          x[ax] -= extents[ax];
          return x;
        }
      }
      return std::nullopt;
    }
  };

  // include/libfly/neigh/adjacent.hpp:40-70 ; loop order from template_for core.hpp:451-467
  // (last axis outermost, axis 0 innermost).
  struct AdjacentCells {
    struct Adj {
      std::array<int, 26> ids{};
      std::size_t count = 0;
    };
    std::vector<Adj> cells;

    AdjacentCells() = default;
    explicit AdjacentCells(std::array<int, 3> const &shape) {
      cells.resize(static_cast<std::size_t>(shape[0]) * shape[1] * shape[2]);
      std::array<int, 3> prod{1, shape[0], shape[0] * shape[1]};
      for (int cz = 0; cz < shape[2]; cz++)
        for (int cy = 0; cy < shape[1]; cy++)
          for (int cx = 0; cx < shape[0]; cx++) {
            int n = cx * prod[0] + cy * prod[1] + cz * prod[2];
            std::size_t slot = 0;
            for (int oz = std::max(cz - 1, 0); oz < std::min(cz + 2, shape[2]); oz++)
              for (int oy = std::max(cy - 1, 0); oy < std::min(cy + 2, shape[1]); oy++)
                for (int ox = std::max(cx - 1, 0); ox < std::min(cx + 2, shape[0]); ox++) {
                  int m = ox * prod[0] + oy * prod[1] + oz * prod[2];
                  if (m != n) {
                    cells[n].ids[slot++] = m;
                  }
                }
            cells[n].count = slot;
          }
    }
  };

  // ------------------------------------------------------------------------------------------
  // neigh::List  include/libfly/neigh/list.hpp:71-210 ; src/neigh/list.cpp:37-165
  // Positions are interleaved xyz double[3N] exactly as the reference's SoA columns.
  // ------------------------------------------------------------------------------------------
  class List {
  public:
    static constexpr long MAX_GHOST_RATIO = 26;
    static constexpr int NONE = std::numeric_limits<int>::max();

    List() = default;
    List(Box const &box, double r_cut) : m_box(box), m_grid(box.extents, r_cut), m_cells(m_grid.shape), m_r_cut(r_cut) {}

    long size() const { return static_cast<long>(m_neigh.size()); }
    long num_plus_ghosts() const { return m_num_plus_ghosts; }
    double r_cut() const { return m_r_cut; }
    Grid const &grid() const { return m_grid; }
    AdjacentCells const &adjacent() const { return m_cells; }

    void rebuild(double const *positions, long n, int num_threads = 1) {  // list.cpp:53-95
      if (num_threads <= 0) {
        throw RuntimeError("invalid number of threads for rebuild()");
      }
      if (size() != n) {
        m_pos.assign(static_cast<std::size_t>(n * (1 + MAX_GHOST_RATIO)), Vec3{});
        m_index.assign(m_pos.size(), 0);
        m_next.assign(m_pos.size(), 0);
        m_neigh.resize(static_cast<std::size_t>(n));
        for (long i = 0; i < n; ++i) {
          m_index[i] = i;
        }
      }
      for (long i = 0; i < n; ++i) {
        m_pos[i] = m_box.canon_image({positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]});
      }
      make_ghosts();

      m_head.assign(static_cast<std::size_t>(m_grid.num_cells()), NONE);
      for (long i = 0; i < m_num_plus_ghosts; i++) {
        int c = m_grid.cell_idx(m_pos[i]);
        m_next[i] = std::exchange(m_head[c], static_cast<int>(i));
      }
#pragma omp parallel for num_threads(num_threads) schedule(static)
      for (long i = 0; i < n; i++) {
        build_neigh_list(i);
      }
    }

    void update(double const *delta, long n) {  // list.cpp:37-51  (r <- r - x)
      if (n != size()) {
        throw RuntimeError("Input to update has wrong number of atoms");
      }
      for (long i = 0; i < n; i++) {
        m_pos[i].x -= delta[3 * i], m_pos[i].y -= delta[3 * i + 1], m_pos[i].z -= delta[3 * i + 2];
      }
      for (long i = n; i < m_num_plus_ghosts; i++) {
        long j = m_index[i];
        m_pos[i].x -= delta[3 * j], m_pos[i].y -= delta[3 * j + 1], m_pos[i].z -= delta[3 * j + 2];
      }
    }

    template <typename F>
    void for_neighbours(long i, double r_cut, F &&f) const {  // list.hpp:138-154
      if (r_cut > m_r_cut) {
        throw RuntimeError("r_cut is bigger than the internal r_cut");
      }
      for (int n : m_neigh[i]) {
        Vec3 dr = m_pos[n] - m_pos[i];
        double r_sq = norm_sq(dr);
        if (r_sq < r_cut * r_cut) {
          f(m_index[n], std::sqrt(r_sq), dr);
        }
      }
    }

    // Introspection for parity tests (not part of the reference API).
    std::vector<int> const &raw_list(long i) const { return m_neigh[i]; }
    long image_to_real(long i) const { return m_index[i]; }
    Vec3 const &position(long i) const { return m_pos[i]; }

  private:
    Box m_box;
    Grid m_grid;
    AdjacentCells m_cells;
    double m_r_cut = 0;
    std::vector<Vec3> m_pos;
    std::vector<long> m_index;
    std::vector<int> m_next;
    long m_num_plus_ghosts = 0;
    std::vector<int> m_head;
    std::vector<std::vector<int>> m_neigh;

    void make_ghosts() {  // list.cpp:100-132
      long next_slot = size();
      for (int ax = 0; ax < 3; ++ax) {
        if (m_box.periodic[ax]) {
          long const end = next_slot;
          for (int sign : {+1, -1}) {
            for (long j = 0; j < end; ++j) {
              if (auto im = m_grid.gen_image(m_pos[j], ax, sign)) {
                m_pos[next_slot] = *im;
                m_index[next_slot] = m_index[j];
                next_slot++;
              }
            }
          }
        }
      }
      m_num_plus_ghosts = next_slot;
    }

    void build_neigh_list(long i) {  // list.cpp:134-165
      auto &out = m_neigh[i];
      out.clear();
      int i_cell = m_grid.cell_idx(m_pos[i]);
      for (int n = m_head[i_cell]; n != NONE; n = m_next[n]) {
        if (n != i && norm_sq(m_pos[i] - m_pos[n]) < m_r_cut * m_r_cut) {
          out.push_back(n);
        }
      }
      auto const &adj = m_cells.cells[i_cell];
      for (std::size_t k = 0; k < adj.count; k++) {
        for (int n = m_head[adj.ids[k]]; n != NONE; n = m_next[n]) {
          if (norm_sq(m_pos[i] - m_pos[n]) < m_r_cut * m_r_cut) {
            out.push_back(n);
          }
        }
      }
    }
  };

  // ------------------------------------------------------------------------------------------
  // system::Hessian  include/libfly/system/hessian.hpp:42-148.  Column-major dense (3N)x(3N),
  // like Eigen::MatrixXd; only the lower triangle is meaningful (SelfAdjointEigenSolver reads it).
  // ------------------------------------------------------------------------------------------
  struct Hessian {
    long n3 = 0;
    std::vector<double> h;

    void zero_for(long n) {
      n3 = 3 * n;
      h.assign(static_cast<std::size_t>(n3 * n3), 0.0);
    }
    double &at(long row, long col) { return h[static_cast<std::size_t>(col * n3 + row)]; }

    void mass_weight(std::vector<double> const &mass, std::uint32_t const *id, long n) {  // hessian.hpp:136-143
      for (long i = 0; i < n; ++i) {
        double mi = mass[id[i]];
        for (long j = 0; j < n; ++j) {
          double s = 1 / std::sqrt(mi * mass[id[j]]);
          for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) {
              at(3 * j + a, 3 * i + b) *= s;
            }
        }
      }
    }
  };

  // ------------------------------------------------------------------------------------------
  // potential::EAM  src/potential/EAM/eam.cpp:33-233
  // ------------------------------------------------------------------------------------------
  class EAM {
  public:
    explicit EAM(std::shared_ptr<DataEAM const> data) : m_data(std::move(data)) {}

    double r_cut() const { return m_data->rcut(); }
    DataEAM const &data() const { return *m_data; }

    double energy(std::uint32_t const *id, long n, List const &nl, int num_threads = 1) {  // eam.cpp:33-51
      double v_sum = 0, f_sum = 0;
#pragma omp parallel for reduction(+ : v_sum, f_sum) num_threads(num_threads) schedule(static)
      for (long a = 0; a < n; a++) {
        double rho = 0;
        nl.for_neighbours(a, r_cut(), [&](long b, double r, Vec3 const &) {
          v_sum += m_data->v(id[a], id[b]).f(r);
          rho += m_data->phi(id[b], id[a]).f(r);
        });
        f_sum += m_data->f(id[a]).f(rho);
      }
      return (0.5 * v_sum) + f_sum;
    }

    // eam.cpp:53-120. Serial branch lets OutOfBounds propagate (eam.cpp:75-79); in the OpenMP
    // branch we trap the exception per thread and rethrow after the loop (the reference would
    // terminate instead -- callers only ever catch it with num_threads == 1).
    void gradient(double *out, std::uint32_t const *id, std::uint8_t const *frozen, long n, List const &nl, int num_threads = 1) {
      m_fprime.resize(static_cast<std::size_t>(n));
      bool oob = false;

      auto body1 = [&](long b) {
        double rho = 0;
        nl.for_neighbours(b, r_cut(), [&](long a, double r, Vec3 const &) { rho += m_data->phi(id[a], id[b]).f(r); });
        m_fprime[b] = m_data->f(id[b]).fp(rho);
      };
      auto body2 = [&](long g) {
        Vec3 grad{};
        if (!frozen[g]) {
          nl.for_neighbours(g, r_cut(), [&](long a, double r, Vec3 const &dr) {
            double mag = m_data->v(id[a], id[g]).fp(r) + m_fprime[g] * m_data->phi(id[a], id[g]).fp(r)
                         + m_fprime[a] * m_data->phi(id[g], id[a]).fp(r);
            double s = mag / r;
            grad.x -= s * dr.x, grad.y -= s * dr.y, grad.z -= s * dr.z;
          });
        }
        out[3 * g] = grad.x, out[3 * g + 1] = grad.y, out[3 * g + 2] = grad.z;
      };

      if (num_threads == 1) {
        for (long b = 0; b < n; b++) body1(b);
        for (long g = 0; g < n; g++) body2(g);
      } else {
#pragma omp parallel for num_threads(num_threads) schedule(static)
        for (long b = 0; b < n; b++) {
          try {
            body1(b);
          } catch (OutOfBounds const &) {
            oob = true;
          }
        }
        if (oob) throw OutOfBounds("rho outside tabulated region");
#pragma omp parallel for num_threads(num_threads) schedule(static)
        for (long g = 0; g < n; g++) {
          try {
            body2(g);
          } catch (OutOfBounds const &) {
            oob = true;
          }
        }
        if (oob) throw OutOfBounds("r outside tabulated region");
      }
    }

    // eam.cpp:122-233. 3x3 blocks out(i,j) live at rows 3i.., cols 3j.. of the dense matrix.
    void hessian(Hessian &out, std::uint32_t const *id, std::uint8_t const *frozen, long n, List const &nl, int num_threads = 1) {
      m_rho.resize(static_cast<std::size_t>(n));
      m_mu.resize(static_cast<std::size_t>(n));
      out.zero_for(n);

#pragma omp parallel for num_threads(num_threads) schedule(static)
      for (long b = 0; b < n; ++b) {
        double rho = 0;
        Vec3 mu{};
        nl.for_neighbours(b, r_cut(), [&](long a, double r, Vec3 const &dr) {
          rho += m_data->phi(id[a], id[b]).f(r);
          double p = m_data->phi(id[a], id[b]).fp(r);
          mu.x -= p * dr.x / r, mu.y -= p * dr.y / r, mu.z -= p * dr.z / r;
        });
        m_rho[b] = rho;
        m_mu[b] = mu;
      }

      auto add_outer = [&out](long bi, long bj, double s, Vec3 const &u, Vec3 const &v) {
        for (int c = 0; c < 3; c++)
          for (int r = 0; r < 3; r++) {
            out.at(3 * bi + r, 3 * bj + c) += s * u[r] * v[c];
          }
      };
      auto add_ident = [&out](long bi, long bj, double s) {
        for (int r = 0; r < 3; r++) {
          out.at(3 * bi + r, 3 * bj + r) += s;
        }
      };

#pragma omp parallel for num_threads(num_threads) schedule(dynamic)
      for (long z = 0; z < n; ++z) {
        if (frozen[z]) {
          continue;
        }
        Vec3 v = m_mu[z];
        add_outer(z, z, m_data->f(id[z]).fpp(m_rho[z]), v, v);

        nl.for_neighbours(z, r_cut(), [&](long a, double r, Vec3 const &dr) {
          double Fp_z = m_data->f(id[z]).fp(m_rho[z]);
          double Fp_a = m_data->f(id[a]).fp(m_rho[a]);
          double Fpp_z = m_data->f(id[z]).fpp(m_rho[z]);
          double Fpp_a = m_data->f(id[a]).fpp(m_rho[a]);
          Spline const &V = m_data->v(id[a], id[z]);
          Spline const &phi_az = m_data->phi(id[a], id[z]);
          Spline const &phi_za = m_data->phi(id[z], id[a]);

          double A = (V.fp(r) + Fp_z * phi_az.fp(r) + Fp_a * phi_za.fp(r)) / r;
          double B = V.fpp(r) + Fp_z * phi_az.fpp(r) + Fp_a * phi_za.fpp(r);
          double pz = phi_za.fp(r);
          double ABFpp = (A - B - Fpp_a * (pz * pz)) / (r * r);

          add_ident(z, z, A);
          add_outer(z, z, -ABFpp, dr, dr);

          if (a > z && !frozen[a]) {
            double BArr = (B - A) / (r * r);
            double ur = Fpp_z * phi_az.fp(r) / r;
            double ru = Fpp_a * phi_za.fp(r) / r;
            add_outer(a, z, -BArr, dr, dr);
            add_ident(a, z, -A);
            add_outer(a, z, ur, dr, m_mu[z]);
            add_outer(a, z, -ru, m_mu[a], dr);
          }

          double ddFg = Fpp_a * phi_za.fp(r);

          nl.for_neighbours(a, r_cut(), [&](long g, double r_ag, Vec3 const &dr_ag) {
            if ((g > z && !frozen[g]) || (g == z && std::sqrt(norm_sq(dr + dr_ag)) > 1e-6)) {
              double mag = ddFg / (r * r_ag) * m_data->phi(id[g], id[a]).fp(r_ag);
              add_outer(g, z, -mag, dr_ag, dr);
            }
          });
        });
      }
    }

  private:
    std::shared_ptr<DataEAM const> m_data;
    std::vector<double> m_fprime, m_rho;
    std::vector<Vec3> m_mu;
  };

  // ------------------------------------------------------------------------------------------
  // Small dense-vector helpers (stand-ins for the Eigen array expressions gdot / gnorm on
  // dynamic 3N arrays, include/libfly/utility/core.hpp:620-680). Plain left-to-right sums.
  // ------------------------------------------------------------------------------------------
  using Arr = std::vector<double>;

  inline double dot(Arr const &a, Arr const &b) {
    double s = 0;
    for (std::size_t i = 0; i < a.size(); i++) s += a[i] * b[i];
    return s;
  }
  inline double dot_diff(Arr const &a, Arr const &b, Arr const &c) {  // (a - b) . c
    double s = 0;
    for (std::size_t i = 0; i < a.size(); i++) s += (a[i] - b[i]) * c[i];
    return s;
  }

  // ------------------------------------------------------------------------------------------
  // minimise::StepLBFGS  src/minimise/LBFGS/lbfgs_core.cpp:29-87 ; lbfgs_core.hpp:37-119
  // ------------------------------------------------------------------------------------------
  class StepLBFGS {
  public:
    explicit StepLBFGS(int n) : m_n(n), m_hist(static_cast<std::size_t>(n)) {}
    void clear() { m_k = 0; }

    Arr &newton_step(Arr const &r, Arr const &g) {
#ifdef FLY_ORACLE_LBFGS_GRAM
      return newton_step_gram(r, g);
#endif
      int prev = (m_k - 1) % m_n;
      m_r.resize(r.size());
      if (m_k > 0) {
        Elem &e = m_hist[prev];
        e.s.resize(r.size()), e.y.resize(r.size());
        for (std::size_t i = 0; i < r.size(); i++) {
          e.s[i] = r[i] - m_prev_x[i];
          e.y[i] = g[i] - m_prev_g[i];
        }
        e.rho = 1.0 / std::abs(dot(e.s, e.y));
      }
      m_prev_x = r;
      m_prev_g = g;
      m_q = g;

      int incur = m_k <= m_n ? 0 : m_k - m_n;
      int bound = m_k <= m_n ? m_k : m_n;

      for (int i = bound - 1; i >= 0; --i) {
        Elem &e = m_hist[(i + incur) % m_n];
        e.alpha = e.rho * dot(e.s, m_q);
        for (std::size_t t = 0; t < m_q.size(); t++) m_q[t] -= e.alpha * e.y[t];
      }
      if (m_k > 0) {
        double s = 1.0 / m_hist[prev].rho / dot(m_hist[prev].y, m_hist[prev].y);
        for (std::size_t t = 0; t < m_q.size(); t++) m_r[t] = m_q[t] * s;
      } else {
        m_r = m_q;
      }
      for (int i = 0; i <= bound - 1; ++i) {
        Elem &e = m_hist[(i + incur) % m_n];
        double beta = e.rho * dot(e.y, m_r);
        double w = e.alpha - beta;
        for (std::size_t t = 0; t < m_r.size(); t++) m_r[t] += w * e.s[t];
      }
      ++m_k;
      return m_r;
    }

#ifdef FLY_ORACLE_LBFGS_GRAM
    // STUDY VARIANT, never part of the oracle builds (scripts/lbfgs_gram_study.py defines the macro):
    // the same recursion with every dot product s_j.q / y_j.r expanded over the Gram entries
    // s_j.y_k, y_j.y_k, s_j.g, y_j.g (what the device computes in ONE pass, search.cuh newton_step),
    // followed by the element-wise combination in the two-loop's own operation order. Used to check
    // that the re-association only perturbs trajectories at rounding level.
    Arr &newton_step_gram(Arr const &r, Arr const &g) {
      int n = m_n;
      int prev = m_k > 0 ? (m_k - 1) % n : 0;
      m_r.resize(r.size());
      if (m_sy.empty()) m_sy.assign(n * n, 0.0), m_yy.assign(n * n, 0.0);
      if (m_k > 0) {
        Elem &e = m_hist[prev];
        e.s.resize(r.size()), e.y.resize(r.size());
        for (std::size_t i = 0; i < r.size(); i++) e.s[i] = r[i] - m_prev_x[i], e.y[i] = g[i] - m_prev_g[i];
      }
      m_prev_x = r, m_prev_g = g;
      int incur = m_k <= n ? 0 : m_k - n;
      int bound = m_k <= n ? m_k : n;
      if (m_k == 0) {
        m_r = g;
        ++m_k;
        return m_r;
      }
      std::vector<double> sg(n), yg(n);
      for (int i = 0; i < bound; i++) {
        int k = (i + incur) % n;
        m_sy[prev * n + k] = dot(m_hist[prev].s, m_hist[k].y);
        m_sy[k * n + prev] = dot(m_hist[k].s, m_hist[prev].y);
        m_yy[prev * n + k] = m_yy[k * n + prev] = dot(m_hist[prev].y, m_hist[k].y);
        sg[k] = dot(m_hist[k].s, g), yg[k] = dot(m_hist[k].y, g);
      }
      m_hist[prev].rho = 1.0 / std::abs(m_sy[prev * n + prev]);
      double h0 = 1.0 / m_hist[prev].rho / m_yy[prev * n + prev];
      std::vector<double> w(n);
      for (int i = bound - 1; i >= 0; --i) {
        int j = (i + incur) % n;
        double sq = sg[j];
        for (int i2 = bound - 1; i2 > i; --i2) {
          int k = (i2 + incur) % n;
          sq -= m_hist[k].alpha * m_sy[j * n + k];
        }
        m_hist[j].alpha = m_hist[j].rho * sq;
      }
      for (int i = 0; i < bound; ++i) {
        int j = (i + incur) % n;
        double yq = yg[j];
        for (int i2 = bound - 1; i2 >= 0; --i2) {
          int k = (i2 + incur) % n;
          yq -= m_hist[k].alpha * m_yy[j * n + k];
        }
        double yr = yq * h0;
        for (int i2 = 0; i2 < i; ++i2) {
          int k = (i2 + incur) % n;
          yr += w[k] * m_sy[k * n + j];
        }
        w[j] = m_hist[j].alpha - m_hist[j].rho * yr;
      }
      for (std::size_t t = 0; t < m_r.size(); t++) {
        double q = g[t];
        for (int i = bound - 1; i >= 0; --i) {
          int j = (i + incur) % n;
          q -= m_hist[j].alpha * m_hist[j].y[t];
        }
        q *= h0;
        for (int i = 0; i < bound; ++i) {
          int j = (i + incur) % n;
          q += w[j] * m_hist[j].s[t];
        }
        m_r[t] = q;
      }
      ++m_k;
      return m_r;
    }
    std::vector<double> m_sy, m_yy;
#endif

  private:
    struct Elem {
      Arr s, y;
      double rho = 0, alpha = 0;
    };
    int m_n, m_k = 0;
    std::vector<Elem> m_hist;
    Arr m_prev_x, m_prev_g, m_q, m_r;
  };

  // Counters so the benchmark can convert searches into atom-gradient evaluations.
  struct Counters {
    long grad_evals = 0, rebuilds = 0, updates = 0, steps = 0;
  };

  // ------------------------------------------------------------------------------------------
  // minimise::LBFGS  src/minimise/LBFGS/lbfgs.cpp:29-115 ; Options lbfgs.hpp:53-88
  // ------------------------------------------------------------------------------------------
  struct LBFGSOptions {
    int n = 10;
    int iter_max = 2000;
    double f2norm = 1e-6;
    double skin_frac = 1.1;
    double proj_tol = 0;
    double max_trust = 0.50;
    double min_trust = 0.01;
    double grow_trust = 1.5;
    double shrink_trust = 0.5;
  };

  class LBFGS {
  public:
    LBFGS(LBFGSOptions const &opt, Box const &box) : m_opt(opt), m_core(opt.n), m_box(box) {}

    Counters counters;

    // returns false on convergence, true on iter_max (sic, lbfgs.hpp:108-118)
    bool minimise(double *out_r, double *out_g, double const *in_r, std::uint32_t const *id, std::uint8_t const *frozen, long n,
                  EAM &pot, int num_threads = 1) {
      m_core.clear();
      double skin = std::max(std::pow(m_opt.skin_frac, 1. / 3.) - 1, 0.0) * pot.r_cut();
      if (double r_cut = pot.r_cut() + skin; !(std::exchange(m_r_cut, r_cut) == r_cut) || !m_nl) {
        m_nl = List(m_box, r_cut);
      }
      std::size_t n3 = static_cast<std::size_t>(3 * n);
      Arr r(in_r, in_r + n3), g(n3);

      m_nl->rebuild(r.data(), n, num_threads), counters.rebuilds++;
      pot.gradient(g.data(), id, frozen, n, *m_nl, num_threads), counters.grad_evals++;

      double trust = m_opt.min_trust;
      double acc = 0;
      bool ret = true;

      for (int i = 0; i < m_opt.iter_max; ++i) {
        double mag_g = dot(g, g);
        if (mag_g < m_opt.f2norm * m_opt.f2norm) {
          ret = false;
          break;
        }
        Arr &Hg = m_core.newton_step(r, g);
        double nrm = std::sqrt(dot(Hg, Hg));
        double scale = std::min(1.0, trust / nrm);
        for (auto &x : Hg) x *= scale;
        // (Hg*Hg).colwise().sum().maxCoeff() on a (3N x 1) array == whole-vector squared norm (lbfgs.cpp:91)
        acc += std::sqrt(dot(Hg, Hg));
        for (std::size_t t = 0; t < n3; t++) r[t] -= Hg[t];

        if (acc > 0.5 * skin) {
          m_nl->rebuild(r.data(), n, num_threads), counters.rebuilds++;
          acc = 0;
        } else {
          m_nl->update(Hg.data(), n), counters.updates++;
        }
        pot.gradient(g.data(), id, frozen, n, *m_nl, num_threads), counters.grad_evals++;
        counters.steps++;

        double proj = dot(g, Hg);
        if (proj < -m_opt.proj_tol) {
          trust = std::max(m_opt.min_trust, m_opt.shrink_trust * trust);
        } else if (proj > m_opt.proj_tol) {
          trust = std::min(m_opt.max_trust, m_opt.grow_trust * trust);
        }
      }
      std::copy(r.begin(), r.end(), out_r);
      std::copy(g.begin(), g.end(), out_g);
      return ret;
    }

  private:
    LBFGSOptions m_opt;
    StepLBFGS m_core;
    Box m_box;
    std::optional<List> m_nl;
    double m_r_cut = -1;
  };

  // ------------------------------------------------------------------------------------------
  // saddle::Rotor  src/saddle/rotor.cpp:33-164 ; Options include/libfly/saddle/rotor.hpp:47-60
  // ------------------------------------------------------------------------------------------
  struct RotorOptions {
    int n = 6;
    int iter_max_rot = 20;
    double delta_r = 0.001;
    double theta_tol = 1 * M_PI / 360.;
    bool relax_in_convex = true;
  };

  inline void trans_proj(Arr &x) {  // rotor.cpp:33-47
    long n = static_cast<long>(x.size() / 3);
    Vec3 p{};
    for (long i = 0; i < n; i++) {
      p.x += x[3 * i], p.y += x[3 * i + 1], p.z += x[3 * i + 2];
    }
    double s = 1.0 / static_cast<double>(n);
    p.x *= s, p.y *= s, p.z *= s;
    for (long i = 0; i < n; i++) {
      x[3 * i] -= p.x, x[3 * i + 1] -= p.y, x[3 * i + 2] -= p.z;
    }
  }

  class Rotor {
  public:
    explicit Rotor(RotorOptions const &opt) : m_opt(opt), m_core(opt.n) {}

    double r_cut(EAM const &pot) const { return pot.r_cut() + m_opt.delta_r * 2; }  // rotor.hpp:123

    Counters counters;

    double eff_gradient(Arr &out, Arr &axis, std::uint32_t const *id, std::uint8_t const *frozen, long n, EAM &pot, List &nl,
                        int count_frozen, int num_threads = 1) {
      std::size_t n3 = static_cast<std::size_t>(3 * n);
      m_delta.resize(n3), m_delta_prev.resize(n3), m_axisp.resize(n3);
      m_g0.resize(n3), m_g1.resize(n3), m_g1p.resize(n3), m_delta_g.resize(n3);
      out.resize(n3);

      m_core.clear();

      pot.gradient(m_g0.data(), id, frozen, n, nl, num_threads), counters.grad_evals++;
      for (std::size_t t = 0; t < n3; t++) m_delta[t] = -m_opt.delta_r * axis[t];
      nl.update(m_delta.data(), n), counters.updates++;
      std::swap(m_delta, m_delta_prev);
      pot.gradient(m_g1.data(), id, frozen, n, nl, num_threads), counters.grad_evals++;

      double curv = 0;
      for (int i = 0;; i++) {
        for (std::size_t t = 0; t < n3; t++) m_delta_g[t] = m_g1[t] - m_g0[t];
        {
          double p = dot(m_delta_g, axis);
          for (std::size_t t = 0; t < n3; t++) m_delta_g[t] -= p * axis[t];
        }
        if (count_frozen == 0) {
          trans_proj(m_delta_g);
        }
        Arr &theta = m_core.newton_step(axis, m_delta_g);
        {
          double p = dot(theta, axis);
          for (std::size_t t = 0; t < n3; t++) theta[t] -= p * axis[t];
          double inv = 1 / std::sqrt(dot(theta, theta));
          for (auto &x : theta) x *= inv;
        }
        double b_1 = dot_diff(m_g1, m_g0, theta) / m_opt.delta_r;
        double c_x0 = dot_diff(m_g1, m_g0, axis) / m_opt.delta_r;
        double theta_1 = -0.5 * std::atan(b_1 / std::abs(c_x0));

        if (std::abs(theta_1) < m_opt.theta_tol || (i >= m_opt.iter_max_rot && c_x0 < 0) || i > 10 * m_opt.iter_max_rot) {
          curv = c_x0;
          break;
        }
        for (std::size_t t = 0; t < n3; t++) m_axisp[t] = axis[t] * std::cos(theta_1) + theta[t] * std::sin(theta_1);

        for (std::size_t t = 0; t < n3; t++) m_delta[t] = -m_opt.delta_r * m_axisp[t];
        std::swap(m_delta, m_delta_prev);
        for (std::size_t t = 0; t < n3; t++) m_delta[t] = m_delta_prev[t] - m_delta[t];
        nl.update(m_delta.data(), n), counters.updates++;

        pot.gradient(m_g1p.data(), id, frozen, n, nl, num_threads), counters.grad_evals++;

        double c_x1 = dot_diff(m_g1p, m_g0, m_axisp) / m_opt.delta_r;
        double a_1 = (c_x0 - c_x1 + b_1 * std::sin(2 * theta_1)) / (1 - std::cos(2 * theta_1));
        double theta_min = 0.5 * std::atan(b_1 / a_1);

        if (a_1 * std::cos(2 * theta_min) - a_1 + b_1 * std::sin(2 * theta_min) > 0) {
          theta_min += M_PI / 2;
        }
        for (std::size_t t = 0; t < n3; t++) axis[t] = axis[t] * std::cos(theta_min) + theta[t] * std::sin(theta_min);

        double k1 = std::sin(theta_1 - theta_min) / std::sin(theta_1);
        double k2 = std::sin(theta_min) / std::sin(theta_1);
        double k3 = 1 - std::cos(theta_min) - std::sin(theta_min) * std::tan(0.5 * theta_1);
        for (std::size_t t = 0; t < n3; t++) m_g1[t] = k1 * m_g1[t] + k2 * m_g1p[t] + k3 * m_g0[t];

        if (std::abs(theta_min) < m_opt.theta_tol) {
          curv = dot_diff(m_g1, m_g0, axis) / m_opt.delta_r;
          break;
        }
      }

      for (std::size_t t = 0; t < n3; t++) m_delta[t] = -m_delta_prev[t];
      nl.update(m_delta.data(), n), counters.updates++;

      double p = dot(m_g0, axis);
      if (!m_opt.relax_in_convex && curv > 0) {
        for (std::size_t t = 0; t < n3; t++) out[t] = -p * axis[t];
      } else {
        for (std::size_t t = 0; t < n3; t++) out[t] = m_g0[t] - 2 * p * axis[t];
      }
      if (count_frozen == 0) {
        trans_proj(out);
      }
      return curv;
    }

  private:
    RotorOptions m_opt;
    StepLBFGS m_core;
    Arr m_delta, m_delta_prev, m_axisp, m_g0, m_g1, m_g1p, m_delta_g;
  };

  // ------------------------------------------------------------------------------------------
  // saddle::Dimer  src/saddle/dimer.cpp:28-171 ; Options include/libfly/saddle/dimer.hpp:56-98
  // ------------------------------------------------------------------------------------------
  struct DimerOptions {
    int n = 10;
    int max_steps = 1000;
    int hist_check_freq = 5;
    double f2norm = 1e-5;
    double skin_frac = 1.1;
    double proj_tol = 0;
    double max_trust = 0.5;
    double min_trust = 0.05;
    double grow_trust = 1.5;
    double shrink_trust = 0.5;
    double convex_max = 3;
    bool use_history = true;
  };

  enum Exit : int { success = 0, convex = -1, iter_max = -2, collision = -3, uninit = -4 };

  class Dimer {
  public:
    Dimer(DimerOptions const &opt, RotorOptions const &r_opt, Box const &box) : m_opt(opt), m_core(opt.n), m_rotor(r_opt), m_box(box) {}

    Counters counters;

    Exit find_sp(double *out_r, double *out_ax, double const *in_r, double const *in_ax, std::uint32_t const *id,
                 std::uint8_t const *frozen, long n, double const *in_min, EAM &pot, std::vector<Arr> const &hist_sp, double theta_tol,
                 int count_frozen, int num_threads = 1) {
      std::size_t n3 = static_cast<std::size_t>(3 * n);
      m_core.clear();
      m_rotor.counters = {};

      double skin = std::max(std::pow(m_opt.skin_frac, 1. / 3.) - 1, 0.0) * m_rotor.r_cut(pot);
      if (double r_cut = m_rotor.r_cut(pot) + skin; !(std::exchange(m_r_cut, r_cut) == r_cut) || !m_nl) {
        m_nl = List(m_box, r_cut);
      }
      Arr r(in_r, in_r + n3), axis(in_ax, in_ax + n3);
      Arr rmin(in_min, in_min + n3);

      auto finish = [&](Exit e) {
        std::copy(r.begin(), r.end(), out_r);
        std::copy(axis.begin(), axis.end(), out_ax);
        counters.grad_evals += m_rotor.counters.grad_evals;
        counters.updates += m_rotor.counters.updates;
        return e;
      };

      m_nl->rebuild(r.data(), n, num_threads), counters.rebuilds++;
      double curv = m_rotor.eff_gradient(m_eff, axis, id, frozen, n, pot, *m_nl, count_frozen, num_threads);

      double trust = m_opt.min_trust;
      double acc = 0;
      double convex_count = 0;

      for (int i = 0; i < m_opt.max_steps; ++i) {
        double mag_g = dot(m_eff, m_eff);
        if (mag_g < m_opt.f2norm * m_opt.f2norm) {
          return finish(success);
        } else if (convex_count >= m_opt.convex_max) {
          return finish(convex);
        }
        if (m_opt.use_history && i % m_opt.hist_check_freq == 0) {  // dimer.cpp:102-131
          double ctheta = -1;
          for (auto const &sp : hist_sp) {
            double dt = 0, sn = 0, on = 0;
            for (std::size_t t = 0; t < n3; t++) {
              double a = r[t] - rmin[t], b = sp[t] - rmin[t];
              dt += a * b, sn += b * b, on += a * a;
            }
            ctheta = std::max(ctheta, dt / (std::sqrt(sn) * std::sqrt(on)));
            if (ctheta > std::cos(theta_tol)) {
              return finish(collision);
            }
          }
        }
        Arr &Hg = m_core.newton_step(r, m_eff);
        double nrm = std::sqrt(dot(Hg, Hg));
        double scale = std::min(1.0, trust / nrm);
        for (auto &x : Hg) x *= scale;
        acc += std::sqrt(dot(Hg, Hg));  // whole-vector norm, dimer.cpp:141
        for (std::size_t t = 0; t < n3; t++) r[t] -= Hg[t];

        if (acc > 0.5 * skin) {
          m_nl->rebuild(r.data(), n, num_threads), counters.rebuilds++;
          acc = 0;
        } else {
          m_nl->update(Hg.data(), n), counters.updates++;
        }
        curv = m_rotor.eff_gradient(m_eff, axis, id, frozen, n, pot, *m_nl, count_frozen, num_threads);
        counters.steps++;

        if (curv > 0) {
          convex_count += 1;
        } else {
          convex_count = 0;
        }
        double proj = dot(m_eff, Hg);
        if (proj < -m_opt.proj_tol) {
          trust = std::max(m_opt.min_trust, m_opt.shrink_trust * trust);
        } else if (proj > m_opt.proj_tol) {
          trust = std::min(m_opt.max_trust, m_opt.grow_trust * trust);
        }
      }
      return finish(iter_max);
    }

  private:
    DimerOptions m_opt;
    StepLBFGS m_core;
    Rotor m_rotor;
    Box m_box;
    std::optional<List> m_nl;
    double m_r_cut = -1;
    Arr m_eff;
  };

  // ------------------------------------------------------------------------------------------
  // Input builders.
  //   Xoshiro 256**            include/libfly/utility/random.hpp:42-162
  //   motif_to_lattice (bcc)   include/libfly/utility/lattice.hpp:50-109 (x fastest, +pi shift :106)
  //   Master::perturb          src/saddle/find.cpp:1147-1214
  // ------------------------------------------------------------------------------------------
  class Xoshiro {
  public:
    using result_type = std::uint64_t;
    explicit Xoshiro(std::array<result_type, 4> const &seed) : s(seed) {}
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return std::numeric_limits<result_type>::max(); }

    result_type operator()() {
      result_type const result = rotl(s[1] * 5, 7) * 9;
      result_type const t = s[1] << 17;
      s[2] ^= s[0], s[3] ^= s[1], s[1] ^= s[2], s[0] ^= s[3];
      s[2] ^= t;
      s[3] = rotl(s[3], 45);
      return result;
    }
    void long_jump() { jump_impl({0x76e15d3efefdcbbf, 0xc5004e441c522fb3, 0x77710069854ee241, 0x39109bb02acbe635}); }
    void jump() { jump_impl({0x180ec6d33cfd0aba, 0xd5a61266f0c9392c, 0xa9582618e03fc9aa, 0x39abdc4529b1661c}); }

  private:
    std::array<result_type, 4> s;
    static result_type rotl(result_type x, int k) { return (x << k) | (x >> (64 - k)); }
    void jump_impl(std::array<result_type, 4> const &J) {
      std::array<result_type, 4> t{0, 0, 0, 0};
      for (result_type j : J) {
        for (int b = 0; b < 64; ++b) {
          if (j & result_type{1} << b) {
            for (int q = 0; q < 4; q++) t[q] ^= s[q];
          }
          (*this)();
        }
      }
      s = t;
    }
  };

  // bcc lattice of nx*ny*nz cells (2 atoms per cell), x fastest, all coordinates + pi.
  inline Arr bcc_lattice(double a, int nx, int ny, int nz) {
    Arr r;
    r.reserve(static_cast<std::size_t>(6) * nx * ny * nz);
    for (int k = 0; k < nz; k++)
      for (int j = 0; j < ny; j++)
        for (int i = 0; i < nx; i++)
          for (double off : {0.0, 0.5}) {
            r.push_back((a * off + a * i) + M_PI);
            r.push_back((a * off + a * j) + M_PI);
            r.push_back((a * off + a * k) + M_PI);
          }
    return r;
  }

  // Perturb positions around `centre` and draw a random unit axis (find.cpp:1147-1214).
  inline double perturb(double *out_r, double *out_ax, double const *in_r, std::uint8_t const *frozen, long n, long centre,
                        List const &nl, Xoshiro &prng, double r_pert = 4.0, double stddev_opt = 0.6) {
    std::normal_distribution<double> normal(0., 1.);
    std::normal_distribution<double> prime(stddev_opt, stddev_opt / 3.0);
    double stddev = -1;
    while (stddev < 0) {
      stddev = prime(prng);
    }
    std::normal_distribution<double> gauss(0, stddev);

    std::copy(in_r, in_r + 3 * n, out_r);
    std::fill(out_ax, out_ax + 3 * n, 0.0);

    nl.for_neighbours(centre, r_pert, [&](long m, double r, Vec3 const &) {
      if (!frozen[m]) {
        double w = (1. - r / r_pert);
        for (int c = 0; c < 3; c++) out_r[3 * m + c] += gauss(prng) * w;
        for (int c = 0; c < 3; c++) out_ax[3 * m + c] += normal(prng);
      }
    });
    for (int c = 0; c < 3; c++) out_r[3 * centre + c] += gauss(prng);
    for (int c = 0; c < 3; c++) out_ax[3 * centre + c] += normal(prng);

    double s = 0;
    for (long t = 0; t < 3 * n; t++) s += out_ax[t] * out_ax[t];
    s = std::sqrt(s);
    for (long t = 0; t < 3 * n; t++) out_ax[t] /= s;
    return stddev;
  }

}  // namespace oracle
