// The reference's class surface for the force-evaluation path, in namespace fly, with the reference's own signatures,
// forwarding to the B200 path through the C ABI (include/openfly_b200.h). A libfly translation unit that includes this
// header instead of the four libfly headers below compiles unchanged; inside libfly, the member-function bodies in this
// file ARE the forwarding stubs a maintainer adds (INTEGRATION.md quotes them).
//
//   fly::potential::Generic   include/libfly/potential/generic.hpp:79-136
//       r_cut() const noexcept -> double
//       energy(system::SoA<TypeID const&, Frozen const&> in, neigh::List const& nl, int threads = 1) -> double
//       gradient(system::SoA<PotentialGradient&> out, system::SoA<TypeID const&, Frozen const&> in, neigh::List const& nl, int = 1)
//       hessian(system::Hessian& out, system::SoA<TypeID const&, Frozen const&> in, neigh::List const& nl, int = 1)
//   fly::neigh::List          include/libfly/neigh/list.hpp:86-161
//       List(system::Box const& box, double r_cut); rebuild(system::SoA<Position const&>, int = 1); update(system::SoA<Delta const&>)
//       for_neighbours(Index i, double r_cut, F&& f) const; size() const noexcept
//   fly::minimise::LBFGS      include/libfly/minimise/LBFGS/lbfgs.hpp:53-118
//       minimise(system::SoA<Position&, PotentialGradient&> out, system::SoA<Position const&, TypeID const&, Frozen const&> in,
//                potential::Generic& pot, int num_threads = 1) -> bool
//   fly::saddle::Dimer        include/libfly/saddle/dimer.hpp:56-155
//       find_sp(system::SoA<Position&, Axis&> out, system::SoA<Position const&, Axis const&, TypeID const&, Frozen const&> in,
//               system::SoA<Position const&> in_min, potential::Generic& pot, std::vector<system::SoA<Position>> const& hist_sp,
//               double theta_tol, int count_frozen, int num_threads = 1) -> Exit
//
// What stands in for the rest of libfly here is the minimum those signatures need: property tags (property.hpp:75,157-215), a
// pointer + size `system::SoA` whose columns are single contiguous arrays exactly like the reference's Eigen-backed ones
// (adaptor.hpp:70-78,200,266: `soa[r_].data()` is the double[3N] the C ABI takes, so there is no host-side copy or re-layout),
// an orthorhombic `system::Box`, `system::TypeMap`, a dense `system::Hessian`, `fly::RuntimeError` and `Spline::OutOfBounds`.
// Inside libfly these are libfly's own types and only the member bodies move over.
//
// Every neigh::List / LBFGS / Dimer object owns its device state (an ofly_batch of one replica), as the reference's objects own
// their lists and L-BFGS histories: a List is valid until ITS next rebuild, whatever other objects do in between. One device
// context per host thread (fly::b200::context()). `threads` arguments are accepted and ignored.
// Batched forms (many replicas of one system per call -- where the speed comes from) are the extra members *_batch.
//
// Header-only, C++20; link with libopenfly_b200.so. No CUDA headers needed.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <istream>
#include <iterator>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "../openfly_b200.h"

namespace fly {

  using Index_t = long;  // Eigen::Index
  inline constexpr int spatial_dims = 3;
  using Vec = std::array<double, 3>;

  // fly::RuntimeError (include/libfly/utility/core.hpp:89-138)
  struct RuntimeError : std::runtime_error {
    using std::runtime_error::runtime_error;
  };
  inline RuntimeError error(std::string const &msg) { return RuntimeError(msg); }

  // Spline::OutOfBounds (include/libfly/utility/spline.hpp:46-51): callers catch it to re-perturb (find.cpp:94,120,718-722)
  struct Spline {
    struct OutOfBounds : RuntimeError {
      using RuntimeError::RuntimeError;
    };
  };

  // ---- properties (property.hpp) ---------------------------------------------------------------------------------------
  template <typename S, int N>
  struct Property {
    using scalar_t = S;
    static constexpr int size() { return N; }
  };
  struct Position : Property<double, 3> {};
  struct Axis : Property<double, 3> {};
  struct Delta : Property<double, 3> {};
  struct PotentialGradient : Property<double, 3> {};
  struct TypeID : Property<std::uint32_t, 1> {};
  struct Frozen : Property<bool, 1> {};
  inline constexpr Position r_{};
  inline constexpr Axis ax_{};
  inline constexpr Delta del_{};
  inline constexpr PotentialGradient g_{};
  inline constexpr TypeID id_{};
  inline constexpr Frozen fzn_{};

  namespace b200 {
    inline void check(int rc) {
      if (rc == OFLY_OK) return;
      std::string msg = ofly_last_error();
      if (rc == OFLY_ERR_OUT_OF_BOUNDS) throw Spline::OutOfBounds(msg);
      throw RuntimeError(msg);
    }
    // One device context per host thread (device: OFLY_DEVICE, default 0).
    inline ofly_ctx *context() {
      struct Holder {
        ofly_ctx *c = nullptr;
        Holder() {
          char const *d = std::getenv("OFLY_DEVICE");
          check(ofly_ctx_create(d ? std::atoi(d) : 0, &c));
        }
        ~Holder() { ofly_ctx_destroy(c); }
      };
      thread_local Holder h;
      return h.c;
    }
  }  // namespace b200

  namespace system {

    // ---- SoA: owning (T) or viewing (T&, T const&) columns, one contiguous array per property (adaptor.hpp) -------------
    namespace detail {
      template <typename M>
      struct Column {  // owning
        using tag = M;
        using scalar = typename M::scalar_t;
        static constexpr bool owns = true, is_const = false;
        std::vector<std::conditional_t<std::is_same_v<scalar, bool>, std::uint8_t, scalar>> v;
        void resize(Index_t n) { v.assign(static_cast<std::size_t>(n * M::size()), {}); }
        auto *ptr() { return v.data(); }
        auto const *ptr() const { return v.data(); }
      };
      template <typename M>
      struct Column<M &> {  // mutable view
        using tag = M;
        using scalar = typename M::scalar_t;
        static constexpr bool owns = false, is_const = false;
        std::conditional_t<std::is_same_v<scalar, bool>, std::uint8_t, scalar> *p = nullptr;
        auto *ptr() const { return p; }
      };
      template <typename M>
      struct Column<M const &> {  // const view
        using tag = M;
        using scalar = typename M::scalar_t;
        static constexpr bool owns = false, is_const = true;
        std::conditional_t<std::is_same_v<scalar, bool>, std::uint8_t, scalar> const *p = nullptr;
        auto const *ptr() const { return p; }
      };
      template <typename P>
      struct Span {
        P *p;
        Index_t n;
        P *data() const { return p; }
        Index_t size() const { return n; }
        P &operator[](Index_t i) const { return p[i]; }
      };
    }  // namespace detail

    template <typename... Ms>
    class SoA {
      template <typename Tag>
      static constexpr bool has = (std::is_same_v<Tag, std::remove_cvref_t<Ms>> || ...);

    public:
      static constexpr bool owns_all = (detail::Column<Ms>::owns && ...);

      SoA() = default;
      explicit SoA(Index_t n) : m_n(n) { (resize_if_owned<Ms>(n), ...); }

      // From any SoA that holds (at least) the same properties: owned columns are copied, views are bound (adaptor.hpp:200,266).
      template <typename... Us>
        requires(!std::is_same_v<SoA<Us...>, SoA<Ms...>>)
      SoA(SoA<Us...> &other) : m_n(other.size()) {
        (bind<Ms>(other), ...);
      }
      template <typename... Us>
        requires(!std::is_same_v<SoA<Us...>, SoA<Ms...>>)
      SoA(SoA<Us...> const &other) : m_n(other.size()) {
        (bind<Ms>(other), ...);
      }

      Index_t size() const noexcept { return m_n; }

      // soa[r_]: the whole column; .data() is the contiguous array (double[3N], uint32[N], one byte per atom for Frozen)
      template <typename Tag>
      auto operator[](Tag) {
        auto *p = col<Tag>().ptr();
        return detail::Span<std::remove_pointer_t<decltype(p)>>{p, m_n * Tag::size()};
      }
      template <typename Tag>
      auto operator[](Tag) const {
        auto *p = col<Tag>().ptr();
        return detail::Span<std::remove_pointer_t<decltype(p)>>{p, m_n * Tag::size()};
      }
      // soa(r_, i): atom i's entry (pointer to its Tag::size() scalars; a reference for scalar properties)
      template <typename Tag>
      decltype(auto) operator()(Tag, Index_t i) {
        auto *p = col<Tag>().ptr() + i * Tag::size();
        if constexpr (Tag::size() == 1) {
          return *p;
        } else {
          return p;
        }
      }
      template <typename Tag>
      decltype(auto) operator()(Tag, Index_t i) const {
        auto *p = col<Tag>().ptr() + i * Tag::size();
        if constexpr (Tag::size() == 1) {
          return *p;
        } else {
          return p;
        }
      }
      // rebind a view column to another SoA's column (adaptor.hpp rebind)
      template <typename Tag, typename... Us>
      void rebind(Tag, SoA<Us...> &other) {
        col<Tag>().p = other[Tag{}].data();
        m_n = other.size();
      }
      template <typename Tag, typename... Us>
      void rebind(Tag, SoA<Us...> const &other) {
        col<Tag>().p = other[Tag{}].data();
        m_n = other.size();
      }

    private:
      template <typename M>
      void resize_if_owned(Index_t n) {
        if constexpr (detail::Column<M>::owns) std::get<detail::Column<M>>(m_cols).resize(n);
      }
      template <typename M, typename Other>
      void bind(Other &other) {
        using Tag = std::remove_cvref_t<M>;
        auto &c = std::get<detail::Column<M>>(m_cols);
        if constexpr (detail::Column<M>::owns) {
          c.resize(m_n);
          auto src = other[Tag{}];
          for (Index_t k = 0; k < src.size(); k++) c.v[static_cast<std::size_t>(k)] = src[k];
        } else {
          c.p = other[Tag{}].data();
        }
      }
      template <typename Tag>
      static constexpr std::size_t index_of() {
        static_assert(has<Tag>, "this SoA has no such property");
        constexpr bool match[] = {std::is_same_v<Tag, std::remove_cvref_t<Ms>>...};
        std::size_t k = 0;
        while (!match[k]) k++;
        return k;
      }
      template <typename Tag>
      auto &col() {
        return std::get<index_of<Tag>()>(m_cols);
      }
      template <typename Tag>
      auto const &col() const {
        return std::get<index_of<Tag>()>(m_cols);
      }

      Index_t m_n = 0;
      std::tuple<detail::Column<Ms>...> m_cols;
    };

    // Orthorhombic system::Box (box.hpp:84-197). box.hpp:63-76 selects a Triclinic cell when the off-diagonal norm of the basis
    // reaches 1e-4; that cell type is outside the B200 path and is refused.
    class Box {
    public:
      // basis: row-major 3x3 with the basis vectors as columns (Mat in box.hpp:84)
      Box(std::array<double, 9> const &basis, std::array<bool, 3> const &periodic) : m_periodic(periodic) {
        double skew = 0;
        for (int i = 0; i < 3; i++) {
          for (int j = 0; j < 3; j++) {
            if (i != j) skew += basis[static_cast<std::size_t>(3 * i + j)] * basis[static_cast<std::size_t>(3 * i + j)];
          }
        }
        if (!(std::sqrt(skew) < 1e-4)) throw error("triclinic boxes are outside the B200 path (orthorhombic only)");
        m_extents = {basis[0], basis[4], basis[8]};
      }
      std::array<double, 3> const &extents() const noexcept { return m_extents; }
      bool periodic(Index_t i) const { return m_periodic[static_cast<std::size_t>(i)]; }

    private:
      std::array<double, 3> m_extents{};
      std::array<bool, 3> m_periodic{};
    };

    // system::TypeMap<> (typemap.hpp): type names by TypeID.
    class TypeMap {
    public:
      explicit TypeMap(std::vector<std::string> names) : m_names(std::move(names)) {}
      Index_t num_types() const noexcept { return static_cast<Index_t>(m_names.size()); }
      std::string const &get(std::uint32_t id) const { return m_names.at(id); }
      std::vector<std::string> const &names() const noexcept { return m_names; }

    private:
      std::vector<std::string> m_names;
    };

    // system::Hessian (hessian.hpp:42-148): dense column-major (3N)^2, lower triangle meaningful.
    class Hessian {
    public:
      auto zero_for(Index_t num_active) -> void {
        if (num_active < 0) throw error("not a valid number of active atoms");
        m_n = spatial_dims * num_active;
        m_hess.assign(static_cast<std::size_t>(m_n * m_n), 0.0);
      }
      double &operator()(Index_t row, Index_t col) { return m_hess[static_cast<std::size_t>(col * m_n + row)]; }
      double operator()(Index_t row, Index_t col) const { return m_hess[static_cast<std::size_t>(col * m_n + row)]; }
      Index_t rows() const noexcept { return m_n; }
      std::vector<double> &get() noexcept { return m_hess; }
      std::vector<double> const &get() const noexcept { return m_hess; }
      // mass_weight (hessian.hpp:136-143): every element scaled by 1 / sqrt(m_row m_col)
      template <typename SoaWithId>
      auto mass_weight(std::vector<double> const &mass_of_type, SoaWithId const &in) -> void {
        for (Index_t i = 0; i < in.size(); ++i) {
          double mi = mass_of_type.at(in(id_, i));
          for (Index_t j = 0; j < in.size(); ++j) {
            double f = 1 / std::sqrt(mi * mass_of_type.at(in(id_, j)));
            for (int a = 0; a < 3; a++) {
              for (int b = 0; b < 3; b++) (*this)(3 * j + a, 3 * i + b) *= f;
            }
          }
        }
      }

    private:
      Index_t m_n = 0;
      std::vector<double> m_hess;
    };

  }  // namespace system

  namespace b200 {
    // Device state of one N-atom system (R replicas), shared by value like the reference's objects are copied per thread.
    struct Batch {
      std::shared_ptr<ofly_batch> h;
      std::vector<std::uint32_t> type;
      std::vector<std::uint8_t> frozen;
      long replicas = 0;

      template <typename In>
      bool matches(In const &in, long R) const {
        if (!h || replicas != R || static_cast<Index_t>(type.size()) != in.size()) return false;
        for (Index_t i = 0; i < in.size(); i++) {
          if (type[static_cast<std::size_t>(i)] != in(id_, i) || (frozen[static_cast<std::size_t>(i)] != 0) != (in(fzn_, i) != 0)) return false;
        }
        return true;
      }
      template <typename In>
      void create(ofly_tables *tables, system::Box const &box, In const &in, long R) {
        type.resize(static_cast<std::size_t>(in.size())), frozen.resize(type.size());
        for (Index_t i = 0; i < in.size(); i++) {
          type[static_cast<std::size_t>(i)] = in(id_, i), frozen[static_cast<std::size_t>(i)] = in(fzn_, i) ? 1 : 0;
        }
        int pd[3] = {box.periodic(0), box.periodic(1), box.periodic(2)};
        ofly_batch *b = nullptr;
        check(ofly_batch_create(context(), tables, box.extents().data(), pd, in.size(), R, type.data(), frozen.data(), &b));
        h.reset(b, ofly_batch_destroy);
        replicas = R;
      }
    };
  }  // namespace b200

  namespace potential {

    // potential::DataEAM (data.hpp:59-116): parsed + splined tables, resident on the device, shared and immutable.
    class DataEAM {
    public:
      struct Options {
        bool debug = false;
        bool symmetric = false;
      };
      explicit DataEAM(std::string const &setfl_path) {
        ofly_tables *t = nullptr;
        b200::check(ofly_tables_load_setfl(b200::context(), setfl_path.c_str(), &t));
        init(t);
      }
      // DataEAM(Options const&, std::istream) (data.cpp:67)
      DataEAM(Options const &, std::istream &in) {
        std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        ofly_tables *t = nullptr;
        b200::check(ofly_tables_load_setfl_mem(b200::context(), text.data(), text.size(), &t));
        init(t);
      }
      ofly_tables *handle() const noexcept { return m_h.get(); }
      double r_cut() const noexcept { return m_rcut; }
      system::TypeMap const &type_map() const noexcept { return *m_map; }
      std::vector<double> const &masses() const noexcept { return m_mass; }

    private:
      void init(ofly_tables *t) {
        m_h.reset(t, ofly_tables_release);
        long info[3];
        double dinfo[3 + 8];
        b200::check(ofly_tables_info(t, info, dinfo));
        m_rcut = dinfo[2];
        std::vector<std::string> names;
        for (int i = 0; i < static_cast<int>(info[0]); i++) names.emplace_back(ofly_tables_type_name(t, i)), m_mass.push_back(dinfo[3 + i]);
        m_map = std::make_shared<system::TypeMap>(std::move(names));
      }
      std::shared_ptr<ofly_tables> m_h;
      std::shared_ptr<system::TypeMap> m_map;
      std::vector<double> m_mass;
      double m_rcut = 0;
    };

    // potential::EAM (eam.hpp:40-117); its constructor checks the caller's TypeMap against the table's (eam.hpp:50-63).
    class EAM {
    public:
      EAM(system::TypeMap const &map, std::shared_ptr<DataEAM const> data) : m_data(std::move(data)) {
        std::vector<char const *> names;
        for (auto const &n : map.names()) names.push_back(n.c_str());
        b200::check(ofly_eam_check_type_map(m_data->handle(), static_cast<int>(names.size()), names.data()));
      }
      auto r_cut() const noexcept -> double { return m_data->r_cut(); }
      DataEAM const &data() const noexcept { return *m_data; }

    private:
      std::shared_ptr<DataEAM const> m_data;
    };

  }  // namespace potential

  namespace neigh {

    class List {
    public:
      static constexpr Index_t MAX_GHOST_RATIO = 26;  // list.hpp:78

      List(system::Box const &box, double r_cut) : m_box(box), m_r_cut(r_cut) {}

      // list.cpp:53-95
      auto rebuild(system::SoA<Position const &> positions, int num_threads = 1) -> void {
        if (num_threads <= 0) throw error("not a valid number of threads for rebuild()");
        m_pos.assign(positions[r_].data(), positions[r_].data() + positions[r_].size());
        m_n = positions.size();
        m_fresh = false;
        if (m_dev.h && static_cast<Index_t>(m_dev.type.size()) == m_n) issue_rebuild();  // else: at the first device_batch(in)
        m_lists_valid = m_pos_valid = false;
      }
      // list.cpp:37-51: r <- r - x
      auto update(system::SoA<Delta const &> x) -> void {
        if (x.size() != m_n) throw error("update size mismatch");
        for (std::size_t k = 0; k < m_pos.size(); k++) m_pos[k] -= x[del_][static_cast<Index_t>(k)];
        if (m_fresh) b200::check(ofly_neigh_update(m_dev.h.get(), x[del_].data()));
        m_pos_valid = false;
      }
      // list.hpp:138-154. The stored lists and the list-space positions are fetched from the device once per rebuild /
      // update and kept with a prefix-sum offset table, so iterating over all atoms is O(total pairs).
      template <typename F>
      auto for_neighbours(Index_t i, double r_cut, F &&f) const -> void {
        if (r_cut > m_r_cut) throw error("r_cut is bigger then than the internal r_cut");
        if (!(i >= 0 && i < size())) throw error("Atom is not in neighbour list");
        fetch();
        auto const &L = m_box.extents();
        for (long k = m_off[static_cast<std::size_t>(i)]; k < m_off[static_cast<std::size_t>(i) + 1]; k++) {
          int n = m_idx[static_cast<std::size_t>(k)], code = m_shift[static_cast<std::size_t>(k)];
          int s[3] = {code % 3 - 1, (code / 3) % 3 - 1, code / 9 - 1};
          Vec dr{};
          for (std::size_t c = 0; c < 3; c++) dr[c] = (m_lpos[3 * static_cast<std::size_t>(n) + c] + s[c] * L[c]) - m_lpos[3 * static_cast<std::size_t>(i) + c];
          double r_sq = (dr[0] * dr[0] + dr[1] * dr[1]) + dr[2] * dr[2];
          if (r_sq < r_cut * r_cut) f(static_cast<Index_t>(n), std::sqrt(r_sq), dr);
        }
      }
      auto size() const noexcept -> Index_t { return m_n; }

      // ---- forwarding plumbing (not in the reference) ----
      // The device batch holding this list, created (or re-created) for the caller's TypeID / Frozen columns: the list rows
      // carry the neighbour's type, so a change of types re-issues the last rebuild.
      ofly_batch *device_batch(system::SoA<TypeID const &, Frozen const &> in, ofly_tables *tables) const {
        if (in.size() != m_n) throw error("size mismatch between the list and the per-atom data");
        if (!m_dev.matches(in, 1) || m_tables != tables) {
          m_dev.create(tables, m_box, in, 1);
          m_tables = tables;
          m_fresh = false;
        }
        if (!m_fresh) issue_rebuild();
        return m_dev.h.get();
      }

    private:
      void issue_rebuild() const {
        b200::check(ofly_neigh_rebuild(m_dev.h.get(), m_r_cut, m_pos.data()));
        m_fresh = true;
      }
      void fetch() const {
        if (!m_dev.h || !m_fresh) {  // for_neighbours before any potential call: types do not matter for the geometry
          system::SoA<TypeID, Frozen> none(m_n);
          if (!m_tables) throw error("for_neighbours needs a device list: evaluate a potential once or use device_batch()");
          device_batch(none, m_tables);
        }
        std::size_t n = static_cast<std::size_t>(m_n);
        if (!m_lists_valid) {
          std::vector<int> counts(n);
          long cap = static_cast<long>(n) * 256, total = 0;
          m_idx.resize(static_cast<std::size_t>(cap)), m_shift.resize(static_cast<std::size_t>(cap));
          b200::check(ofly_neigh_export(m_dev.h.get(), 0, counts.data(), m_idx.data(), m_shift.data(), cap, &total));
          if (total > cap) throw error("neighbour export overflow");
          m_off.assign(n + 1, 0);
          for (std::size_t a = 0; a < n; a++) m_off[a + 1] = m_off[a] + counts[a];
          m_lists_valid = true;
        }
        if (!m_pos_valid) {
          m_lpos.resize(3 * n);
          b200::check(ofly_neigh_positions(m_dev.h.get(), 0, m_lpos.data()));
          m_pos_valid = true;
        }
      }
      system::Box m_box;
      double m_r_cut;
      Index_t m_n = 0;
      std::vector<double> m_pos;  // user-space positions of the last rebuild minus the updates since
      mutable b200::Batch m_dev;
      mutable ofly_tables *m_tables = nullptr;
      mutable bool m_fresh = false, m_lists_valid = false, m_pos_valid = false;
      mutable std::vector<int> m_idx, m_shift;
      mutable std::vector<long> m_off;
      mutable std::vector<double> m_lpos;
    };

  }  // namespace neigh

  namespace potential {

    // potential::Generic (generic.hpp:79-136) holding an EAM.
    class Generic {
    public:
      explicit Generic(EAM pot) : m_pot(std::move(pot)) {}

      auto r_cut() const noexcept -> double { return m_pot.r_cut(); }

      auto energy(system::SoA<TypeID const &, Frozen const &> in, neigh::List const &nl, int /*threads*/ = 1) -> double {
        double e = 0;
        b200::check(ofly_eam_energy(nl.device_batch(in, tables()), &e));  // eam.cpp:33-51
        return e;
      }

      auto gradient(system::SoA<PotentialGradient &> out, system::SoA<TypeID const &, Frozen const &> in, neigh::List const &nl,
                    int /*threads*/ = 1) -> void {
        if (in.size() != out.size()) throw error("size mismatch!");             // eam.cpp:55
        b200::check(ofly_eam_gradient(nl.device_batch(in, tables()), out[g_].data()));  // eam.cpp:53-120; OutOfBounds re-thrown
      }

      auto hessian(system::Hessian &out, system::SoA<TypeID const &, Frozen const &> in, neigh::List const &nl, int /*threads*/ = 1) -> void {
        out.zero_for(in.size());  // hessian.hpp:59-62
        b200::check(ofly_eam_hessian(nl.device_batch(in, tables()), 0, 0, out.get().data()));  // eam.cpp:122-233
      }

      ofly_tables *tables() const noexcept { return m_pot.data().handle(); }
      EAM const &eam() const noexcept { return m_pot; }

    private:
      EAM m_pot;
    };

  }  // namespace potential

  namespace minimise {

    class LBFGS {
    public:
      struct Options : ofly_lbfgs_options {  // lbfgs.hpp:53-88 defaults
        Options() { ofly_lbfgs_options_default(this); }
      };

      LBFGS(Options const &opt, system::Box const &box) : m_opt{opt}, m_box{box} {}

      // lbfgs.cpp:29-115. Returns false on convergence, true on iter_max (lbfgs.hpp:108-118).
      auto minimise(system::SoA<Position &, PotentialGradient &> out, system::SoA<Position const &, TypeID const &, Frozen const &> in,
                    potential::Generic &pot, int /*num_threads*/ = 1) -> bool {
        if (!m_dev.matches(in, 1) || m_tables != pot.tables()) m_dev.create(pot.tables(), m_box, in, 1), m_tables = pot.tables();
        int status = 0;
        b200::check(ofly_lbfgs_minimise(m_dev.h.get(), &m_opt, in[r_].data(), out[r_].data(), out[g_].data(), &status, nullptr));
        if (status == -5) throw Spline::OutOfBounds("x is outside tabulated region");
        if (status == -6) throw error("neighbour capacity fault");
        if (status == -7) throw error("minimisation ended without a verdict (tick budget)");
        return status != 0;
      }

    private:
      Options m_opt;
      system::Box m_box;
      b200::Batch m_dev;
      ofly_tables *m_tables = nullptr;
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

      Dimer(Options const &opt, Rotor::Options const &r_opt, system::Box const &box) : m_opt(opt), m_ropt(r_opt), m_box(box) {}

      // dimer.cpp:28-171
      auto find_sp(system::SoA<Position &, Axis &> out, system::SoA<Position const &, Axis const &, TypeID const &, Frozen const &> in,
                   system::SoA<Position const &> in_min, potential::Generic &pot, std::vector<system::SoA<Position>> const &hist_sp,
                   double theta_tol, int count_frozen, int /*num_threads*/ = 1) -> Exit {
        if (!m_dev.matches(in, 1) || m_tables != pot.tables()) m_dev.create(pot.tables(), m_box, in, 1), m_tables = pot.tables();
        const std::size_t n3 = static_cast<std::size_t>(3 * in.size());
        m_hist.resize(hist_sp.size() * n3);  // the SP cache as one contiguous [n][3N] array
        for (std::size_t k = 0; k < hist_sp.size(); k++) {
          auto col = hist_sp[k][r_];
          std::copy(col.data(), col.data() + n3, m_hist.begin() + static_cast<std::ptrdiff_t>(k * n3));
        }
        int code = uninit;
        b200::check(ofly_dimer_find_sp(m_dev.h.get(), &m_opt, &m_ropt, in[r_].data(), in[ax_].data(), in_min[r_].data(),
                                       m_hist.empty() ? nullptr : m_hist.data(), static_cast<long>(hist_sp.size()), theta_tol, count_frozen,
                                       out[r_].data(), out[ax_].data(), &code, nullptr));
        if (code == -5) throw Spline::OutOfBounds("x is outside tabulated region");  // Master re-perturbs (find.cpp:718-722)
        if (code == -6) throw error("neighbour capacity fault");
        if (code == -7) throw error("search ended without a verdict (tick budget)");
        return static_cast<Exit>(code);
      }

      // The batch Master::find_batch launches (find.cpp:684-755), as ONE call: `n` start states [n][3N] (positions, axes)
      // stream through `slots` resident replicas; exits[n]; a search that ran off a table reports -5 (re-perturb it).
      auto find_sp_batch(double *out_pos, double *out_axis, double const *in_pos, double const *in_axis, long n, long slots,
                         system::SoA<TypeID const &, Frozen const &> in, system::SoA<Position const &> in_min, potential::Generic &pot,
                         std::vector<system::SoA<Position>> const &hist_sp, double theta_tol, int count_frozen) -> std::vector<int> {
        if (!m_many.matches(in, slots) || m_tables_many != pot.tables()) m_many.create(pot.tables(), m_box, in, slots), m_tables_many = pot.tables();
        const std::size_t n3 = static_cast<std::size_t>(3 * in.size());
        m_hist.resize(hist_sp.size() * n3);
        for (std::size_t k = 0; k < hist_sp.size(); k++) {
          auto col = hist_sp[k][r_];
          std::copy(col.data(), col.data() + n3, m_hist.begin() + static_cast<std::ptrdiff_t>(k * n3));
        }
        std::vector<int> code(static_cast<std::size_t>(n), uninit);
        b200::check(ofly_dimer_find_sp_stream(m_many.h.get(), &m_opt, &m_ropt, n, in_pos, in_axis, in_min[r_].data(),
                                              m_hist.empty() ? nullptr : m_hist.data(), static_cast<long>(hist_sp.size()), theta_tol, count_frozen,
                                              out_pos, out_axis, code.data(), nullptr));
        return code;
      }

    private:
      Options m_opt;
      Rotor::Options m_ropt;
      system::Box m_box;
      b200::Batch m_dev, m_many;
      ofly_tables *m_tables = nullptr, *m_tables_many = nullptr;
      std::vector<double> m_hist;
    };

  }  // namespace saddle

}  // namespace fly
