// C++20 host shim over the C ABI (include/openfly_b200.h): the reference's class surface for the
// force-evaluation path, same names, argument meaning and error behaviour, so a libfly translation unit
// can forward to the B200 path with a one-line body (INTEGRATION.md shows the forwarding stubs).
//
//   fly_b200::system::Box          system::Box                   include/libfly/system/box.hpp:84-197 (orthorhombic only)
//   fly_b200::potential::DataEAM   potential::DataEAM            include/libfly/potential/EAM/data.hpp:59-116
//   fly_b200::potential::Generic   potential::Generic (EAM)      include/libfly/potential/generic.hpp:79-136
//   fly_b200::neigh::List          neigh::List                   include/libfly/neigh/list.hpp:86-161
//   fly_b200::minimise::LBFGS      minimise::LBFGS               include/libfly/minimise/LBFGS/lbfgs.hpp:53-118
//   fly_b200::saddle::Rotor        saddle::Rotor (options only)  include/libfly/saddle/rotor.hpp:47-60
//   fly_b200::saddle::Dimer        saddle::Dimer                 include/libfly/saddle/dimer.hpp:56-155
//
// Properties cross the boundary as the reference lays them out (include/libfly/system/adaptor.hpp:70-78):
// Position / Axis / Delta / PotentialGradient are contiguous double[3N], xyz interleaved; TypeID is
// uint32[N]; Frozen one byte per atom -- `soa[r_].data()` of a libfly SoA is the pointer handed in here,
// no copy on the host side. The one extension is the batch: a `System` owns R replicas of one N-atom
// system on the device, and every call also exists in a [R][3N] form (what saddle::Master::find_batch
// launches, src/saddle/find.cpp:684-755).
//
// Header-only; link with libopenfly_b200.so. No CUDA headers needed.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../openfly_b200.h"

namespace fly_b200 {

  // fly::RuntimeError (include/libfly/utility/core.hpp:89-138)
  struct RuntimeError : std::runtime_error {
    using std::runtime_error::runtime_error;
  };
  // Spline::OutOfBounds (include/libfly/utility/spline.hpp:46-51): callers catch it to re-perturb (find.cpp:94,120,718-722)
  struct OutOfBounds : RuntimeError {
    using RuntimeError::RuntimeError;
  };
  struct CudaError : RuntimeError {
    using RuntimeError::RuntimeError;
  };

  namespace detail {
    inline void check(int rc) {
      if (rc == OFLY_OK) return;
      std::string msg = ofly_last_error();
      if (rc == OFLY_ERR_OUT_OF_BOUNDS) throw OutOfBounds(msg);
      if (rc == OFLY_ERR_CUDA) throw CudaError(msg);
      throw RuntimeError(msg);
    }
    inline void expect(bool ok, char const *msg) {
      if (!ok) throw RuntimeError(msg);
    }
  }  // namespace detail

  // One device + stream; one Context per (host thread, device). No global state.
  class Context {
  public:
    explicit Context(int device = 0) {
      ofly_ctx *c = nullptr;
      detail::check(ofly_ctx_create(device, &c));
      m_h.reset(c, ofly_ctx_destroy);
    }
    ofly_ctx *handle() const noexcept { return m_h.get(); }
    void synchronize() const { detail::check(ofly_ctx_synchronize(m_h.get())); }
    long launch_count() const noexcept { return ofly_ctx_launch_count(m_h.get()); }

  private:
    std::shared_ptr<ofly_ctx> m_h;
  };

  namespace system {
    // Orthorhombic system::Box: extents = diagonal of the basis. box.hpp:63-76 builds a Triclinic grid when the
    // off-diagonal norm reaches 1e-4; that cell type is outside the B200 path and is refused here.
    struct Box {
      std::array<double, 3> extents{};
      std::array<bool, 3> periodic{true, true, true};

      Box() = default;
      Box(std::array<double, 3> ext, std::array<bool, 3> pd = {true, true, true}) : extents(ext), periodic(pd) {}
      // basis: row-major 3x3, as Mat in box.hpp:84
      static Box from_basis(std::span<double const, 9> basis, std::array<bool, 3> pd) {
        double skew = 0;
        for (int i = 0; i < 3; i++) {
          for (int j = 0; j < 3; j++) {
            if (i != j) skew += basis[3 * i + j] * basis[3 * i + j];
          }
        }
        detail::expect(std::sqrt(skew) < 1e-4, "triclinic boxes are outside the B200 path (orthorhombic only)");
        return Box({basis[0], basis[4], basis[8]}, pd);
      }
    };
  }  // namespace system

  namespace potential {
    // Parsed + splined setfl tables resident on the device: shared, immutable, ref-counted like the
    // std::shared_ptr<DataEAM const> every potential::EAM copy holds (eam.hpp:113).
    class DataEAM {
    public:
      DataEAM(Context const &ctx, std::string const &setfl_path) : m_ctx(ctx) {
        ofly_tables *t = nullptr;
        detail::check(ofly_tables_load_setfl(ctx.handle(), setfl_path.c_str(), &t));
        m_h.reset(t, ofly_tables_release);
        long info[3];
        double dinfo[3 + 8];
        detail::check(ofly_tables_info(t, info, dinfo));
        m_types = static_cast<int>(info[0]);
        m_rcut = dinfo[2];
        m_mass.assign(dinfo + 3, dinfo + 3 + m_types);
      }
      ofly_tables *handle() const noexcept { return m_h.get(); }
      Context const &context() const noexcept { return m_ctx; }
      int num_types() const noexcept { return m_types; }
      double r_cut() const noexcept { return m_rcut; }
      double mass(int type) const { return m_mass.at(static_cast<std::size_t>(type)); }

    private:
      Context m_ctx;
      std::shared_ptr<ofly_tables> m_h;
      int m_types = 0;
      double m_rcut = 0;
      std::vector<double> m_mass;
    };
  }  // namespace potential

  // R replicas of one N-atom system on the device: positions in list space, neighbour lists, workspaces.
  // It stands for the {TypeID, Frozen} columns + Box that the reference passes to every call.
  class System {
  public:
    System(potential::DataEAM const &data, system::Box const &box, std::span<std::uint32_t const> type_id,
           std::span<std::uint8_t const> frozen = {}, long replicas = 1)
        : m_data(data), m_box(box), m_n(static_cast<long>(type_id.size())), m_r(replicas) {
      detail::expect(frozen.empty() || frozen.size() == type_id.size(), "Frozen and TypeID columns differ in length");
      int pd[3] = {box.periodic[0], box.periodic[1], box.periodic[2]};
      std::vector<std::uint8_t> none(frozen.empty() ? type_id.size() : 0, 0);
      ofly_batch *b = nullptr;
      detail::check(ofly_batch_create(data.context().handle(), data.handle(), box.extents.data(), pd, m_n, replicas, type_id.data(),
                                      frozen.empty() ? none.data() : frozen.data(), &b));
      m_h.reset(b, ofly_batch_destroy);
    }
    ofly_batch *handle() const noexcept { return m_h.get(); }
    long size() const noexcept { return m_n; }
    long replicas() const noexcept { return m_r; }
    system::Box const &box() const noexcept { return m_box; }
    potential::DataEAM const &data() const noexcept { return m_data; }
    // [R][3N] buffers
    void expect_vec(std::size_t got, char const *what) const {
      if (got != static_cast<std::size_t>(3 * m_n * m_r)) throw RuntimeError(std::string(what) + ": expected R x 3N values");
    }

  private:
    potential::DataEAM m_data;
    system::Box m_box;
    long m_n, m_r;
    std::shared_ptr<ofly_batch> m_h;
  };

  namespace neigh {
    // NOTE: every List, LBFGS and Dimer made on one System shares that System's single device list: a List is valid until the
    // next rebuild / minimise / find_sp on the same System, by whomever. (The reference's objects each own their list; so do
    // the classes of include/openfly_b200/libfly.hpp, which is the surface to use for reference-style code.)
    class List {
    public:
      static constexpr long MAX_GHOST_RATIO = 26;  // list.hpp:78

      List(System const &sys, double r_cut) : m_sys(sys), m_r_cut(r_cut) {}

      // rebuild(positions): canonicalise, ghosts, cells, lists (list.cpp:53-95); [R][3N]
      void rebuild(std::span<double const> positions, int num_threads = 1) {
        detail::expect(num_threads > 0, "not a valid number of threads for rebuild()");
        m_sys.expect_vec(positions.size(), "rebuild");
        detail::check(ofly_neigh_rebuild(m_sys.handle(), m_r_cut, positions.data()));
      }
      // update(x): r <- r - x for real atoms and their images (list.cpp:37-51)
      void update(std::span<double const> delta) {
        m_sys.expect_vec(delta.size(), "update");
        detail::check(ofly_neigh_update(m_sys.handle(), delta.data()));
      }
      // for_neighbours(i, r_cut, f(n, r, dr)) (list.hpp:138-154); dr = r_n - r_i with the neighbour's image shift applied
      template <typename F>
      void for_neighbours(long i, double r_cut, F &&f, long replica = 0) const {
        detail::expect(r_cut <= m_r_cut, "r_cut is bigger then than the internal r_cut");
        detail::expect(i >= 0 && i < size(), "atom is not in the neighbour list");
        fetch(replica);
        long off = 0;
        for (long a = 0; a < i; a++) off += m_counts[static_cast<std::size_t>(a)];
        auto const &L = m_sys.box().extents;
        for (long k = off; k < off + m_counts[static_cast<std::size_t>(i)]; k++) {
          int n = m_idx[static_cast<std::size_t>(k)], code = m_shift[static_cast<std::size_t>(k)];
          int s[3] = {code % 3 - 1, (code / 3) % 3 - 1, code / 9 - 1};
          std::array<double, 3> dr{};
          for (int c = 0; c < 3; c++) dr[c] = (m_pos[3 * n + c] + s[c] * L[c]) - m_pos[3 * i + c];
          double r_sq = (dr[0] * dr[0] + dr[1] * dr[1]) + dr[2] * dr[2];
          if (r_sq < r_cut * r_cut) f(static_cast<long>(n), std::sqrt(r_sq), dr);
        }
      }
      long size() const noexcept { return m_sys.size(); }
      double r_cut() const noexcept { return m_r_cut; }
      System const &system() const noexcept { return m_sys; }

    private:
      void fetch(long replica) const {
        std::size_t n = static_cast<std::size_t>(size());
        m_counts.resize(n), m_pos.resize(3 * n);
        long cap = static_cast<long>(n) * 512, total = 0;
        m_idx.resize(static_cast<std::size_t>(cap)), m_shift.resize(static_cast<std::size_t>(cap));
        detail::check(ofly_neigh_export(m_sys.handle(), replica, m_counts.data(), m_idx.data(), m_shift.data(), cap, &total));
        detail::check(ofly_neigh_positions(m_sys.handle(), replica, m_pos.data()));
      }
      System m_sys;
      double m_r_cut;
      mutable std::vector<int> m_counts, m_idx, m_shift;
      mutable std::vector<double> m_pos;
    };
  }  // namespace neigh

  namespace potential {
    // potential::Generic holding an EAM (generic.hpp:79-136). `num_threads` is accepted and ignored.
    class Generic {
    public:
      explicit Generic(DataEAM data) : m_data(std::move(data)) {}
      double r_cut() const noexcept { return m_data.r_cut(); }
      // energy of every replica (eam.cpp:33-51); returns replica 0, fills `all` if given
      double energy(neigh::List const &nl, int = 1, std::span<double> all = {}) const {
        std::vector<double> e(static_cast<std::size_t>(nl.system().replicas()));
        detail::check(ofly_eam_energy(nl.system().handle(), e.data()));
        for (std::size_t k = 0; k < all.size() && k < e.size(); k++) all[k] = e[k];
        return e[0];
      }
      // gradient (eam.cpp:53-120): out [R][3N], zero on frozen atoms
      void gradient(std::span<double> out, neigh::List const &nl, int = 1) const {
        nl.system().expect_vec(out.size(), "gradient");
        detail::check(ofly_eam_gradient(nl.system().handle(), out.data()));
      }
      // hessian (eam.cpp:122-233): dense column-major (3N)^2, lower triangle (system::Hessian, hessian.hpp:42-148)
      void hessian(std::span<double> out, neigh::List const &nl, int = 1, long replica = 0, bool mass_weight = false) const {
        std::size_t n3 = static_cast<std::size_t>(3 * nl.size());
        detail::expect(out.size() == n3 * n3, "hessian: expected (3N)^2 values");
        detail::check(ofly_eam_hessian(nl.system().handle(), replica, mass_weight ? 1 : 0, out.data()));
      }

      // system::Hessian::eigenvalues() (hessian.hpp:86-93) of that matrix, ascending, solved on the device: out [3N].
      // What saddle::Master feeds its Vineyard prefactors with (find.cpp:846-871, 1107-1135).
      void hessian_eigenvalues(std::span<double> out, neigh::List const &nl, long replica = 0, bool mass_weight = true) const {
        detail::expect(out.size() == static_cast<std::size_t>(3 * nl.size()), "hessian_eigenvalues: expected 3N values");
        detail::check(ofly_eam_hessian_eigenvalues(nl.system().handle(), replica, mass_weight ? 1 : 0, out.data()));
      }

    private:
      DataEAM m_data;
    };
  }  // namespace potential

  namespace minimise {
    class LBFGS {
    public:
      struct Options : ofly_lbfgs_options {  // lbfgs.hpp:53-88 defaults
        Options() { ofly_lbfgs_options_default(this); }
      };
      LBFGS(Options const &opt, system::Box const &) : m_opt(opt) {}

      // minimise(out{Position, PotentialGradient}, in{Position}) -> false on convergence, true on iter_max
      // (lbfgs.hpp:108-118), replica 0; `failed` (optional) receives every replica's flag. A replica that ran off
      // a spline table throws OutOfBounds like the reference does.
      bool minimise(std::span<double> out_pos, std::span<double> out_grad, std::span<double const> in_pos, System const &sys,
                    potential::Generic &, int = 1, std::vector<bool> *failed = nullptr) {
        sys.expect_vec(out_pos.size(), "minimise out"), sys.expect_vec(in_pos.size(), "minimise in");
        std::vector<int> status(static_cast<std::size_t>(sys.replicas()));
        detail::check(ofly_lbfgs_minimise(sys.handle(), &m_opt, in_pos.data(), out_pos.data(), out_grad.empty() ? nullptr : out_grad.data(),
                                          status.data(), nullptr));
        for (int s : status) {
          if (s == -5) throw OutOfBounds("x is outside tabulated region");
          if (s == -7) throw RuntimeError("minimisation ended without a verdict (tick budget exhausted)");
          if (s < 0) throw RuntimeError("neighbour capacity fault");
        }
        if (failed) failed->assign(status.begin(), status.end());
        return status[0] != 0;
      }

    private:
      Options m_opt;
    };
  }  // namespace minimise

  namespace saddle {
    struct Rotor {
      struct Options : ofly_rotor_options {  // rotor.hpp:47-60 defaults
        Options() { ofly_rotor_options_default(this); }
      };
    };

    class Dimer {
    public:
      struct Options : ofly_dimer_options {  // dimer.hpp:56-98 defaults
        Options() { ofly_dimer_options_default(this); }
      };
      enum Exit : int { success = 0, convex = -1, iter_max = -2, collision = -3, uninit = -4 };  // dimer.hpp:123-129

      Dimer(Options const &opt, Rotor::Options const &r_opt, system::Box const &) : m_opt(opt), m_ropt(r_opt) {}

      // find_sp for every replica of `sys` (dimer.hpp:148-155): out/in [R][3N], in_min [3N], hist_sp n x [3N].
      // Returns the exits; a replica that hit Spline::OutOfBounds throws (Master catches and re-perturbs).
      std::vector<Exit> find_sp_batch(std::span<double> out_pos, std::span<double> out_axis, std::span<double const> in_pos,
                                      std::span<double const> in_axis, std::span<double const> in_min, System const &sys,
                                      potential::Generic &, std::span<double const> hist_sp, double theta_tol, int count_frozen,
                                      int = 1, bool throw_oob = true) {
        sys.expect_vec(out_pos.size(), "find_sp out"), sys.expect_vec(in_pos.size(), "find_sp in");
        sys.expect_vec(out_axis.size(), "find_sp out axis"), sys.expect_vec(in_axis.size(), "find_sp in axis");
        std::size_t n3 = static_cast<std::size_t>(3 * sys.size());
        detail::expect(in_min.size() == n3 && hist_sp.size() % n3 == 0, "find_sp: in_min / hist_sp are not [3N]");
        std::vector<int> code(static_cast<std::size_t>(sys.replicas()));
        detail::check(ofly_dimer_find_sp(sys.handle(), &m_opt, &m_ropt, in_pos.data(), in_axis.data(), in_min.data(),
                                         hist_sp.empty() ? nullptr : hist_sp.data(), static_cast<long>(hist_sp.size() / n3), theta_tol,
                                         count_frozen, out_pos.data(), out_axis.data(), code.data(), nullptr));
        std::vector<Exit> out;
        for (int c : code) {
          if (c == -5 && throw_oob) throw OutOfBounds("x is outside tabulated region");
          if (c == -6) throw RuntimeError("neighbour capacity fault");
          if (c == -7) throw RuntimeError("search ended without a verdict (tick budget exhausted)");
          out.push_back(static_cast<Exit>(c));
        }
        return out;
      }
      // The reference's single-system call (R = 1).
      Exit find_sp(std::span<double> out_pos, std::span<double> out_axis, std::span<double const> in_pos, std::span<double const> in_axis,
                   std::span<double const> in_min, System const &sys, potential::Generic &pot, std::span<double const> hist_sp,
                   double theta_tol, int count_frozen, int num_threads = 1) {
        detail::expect(sys.replicas() == 1, "find_sp: use find_sp_batch for R > 1");
        return find_sp_batch(out_pos, out_axis, in_pos, in_axis, in_min, sys, pot, hist_sp, theta_tol, count_frozen, num_threads)[0];
      }

    private:
      Options m_opt;
      Rotor::Options m_ropt;
    };
  }  // namespace saddle

}  // namespace fly_b200
