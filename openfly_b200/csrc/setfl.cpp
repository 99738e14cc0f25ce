#include "setfl.hpp"

#include <algorithm>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace ofly {

  std::vector<Knot> natural_spline(std::vector<double> y, double dx) {
    // Pad the tail so that evaluation slightly past the table is smooth (spline.cpp:29-33).
    y.insert(y.end(), 5, y.back());

    std::size_t const n = y.size() - 1;

    // Tridiagonal system for the second-derivative coefficients c[] of a natural spline with
    // constant spacing; forward elimination keeps (mu, z), back-substitution yields c, b, d.
    std::vector<double> mu(n + 1, 0.0), z(n + 1, 0.0), c(n + 1, 0.0);

    for (std::size_t i = 1; i + 1 <= n; ++i) {
      double rhs = 3 * (y[i + 1] - 2 * y[i] + y[i - 1]) / dx;  // spline.cpp:60
      double diag = dx * (4 - mu[i - 1]);                       // spline.cpp:70
      mu[i] = dx / diag;
      z[i] = (rhs - dx * z[i - 1]) / diag;
    }

    std::vector<Knot> out(n);

    for (std::size_t j = n; j-- > 0;) {
      c[j] = z[j] - mu[j] * c[j + 1];                                            // spline.cpp:84
      out[j].a = y[j];
      out[j].b = (y[j + 1] - y[j]) / dx - dx * (c[j + 1] + 2 * c[j]) / 3;         // spline.cpp:85
      out[j].d = (c[j + 1] - c[j]) / (3 * dx);                                   // spline.cpp:86
      out[j].c = c[j];
    }
    return out;
  }

  namespace {

    struct Reader {
      std::istream &in;
      explicit Reader(std::istream &stream) : in(stream) {}
      std::istringstream line() {
        std::string s;
        if (!std::getline(in, s)) {
          throw std::runtime_error("setfl file terminated too soon");
        }
        return std::istringstream{s};
      }
      // `count` numbers laid out `per_line` to a line (data.cpp:48-65).
      std::vector<double> block(long count, long per_line) {
        std::vector<double> v;
        v.reserve(static_cast<std::size_t>(count));
        std::istringstream ls;
        for (long i = 0; i < count; ++i) {
          if (i % per_line == 0) {
            ls = line();
          }
          double x;
          ls >> x;
          v.push_back(x);
        }
        return v;
      }
    };

    void place(std::vector<Knot> &dst, std::size_t slot, std::vector<Knot> const &k) {
      std::copy(k.begin(), k.end(), dst.begin() + static_cast<std::ptrdiff_t>(slot * k.size()));
    }

  }  // namespace

  HostTables read_setfl(std::string const &path) {
    std::ifstream in(path);
    if (!in.good()) {
      throw std::runtime_error("Could not open eam file '" + path + "'");
    }
    return read_setfl(in);
  }

  HostTables read_setfl(std::istream &in) {
    Reader rd(in);
    HostTables t;

    for (int i = 0; i < 3; ++i) {
      rd.line();  // three comment lines
    }
    {
      auto ls = rd.line();
      ls >> t.n_types;
      if (t.n_types <= 0 || t.n_types > 8) {
        throw std::runtime_error("setfl: unsupported number of types");
      }
      t.names.resize(static_cast<std::size_t>(t.n_types));
      for (auto &nm : t.names) {
        ls >> nm;
      }
    }
    rd.line() >> t.n_rho >> t.d_rho >> t.n_r >> t.d_r >> t.r_cut;
    if (t.n_rho < 2 || t.n_r < 2 || !(t.d_rho > 0) || !(t.d_r > 0)) {
      throw std::runtime_error("setfl: bad tabulation header");
    }
    t.knots_rho = t.n_rho + 4;
    t.knots_r = t.n_r + 4;

    std::size_t const n = static_cast<std::size_t>(t.n_types);
    t.mass.resize(n);
    t.F.assign(n * static_cast<std::size_t>(t.knots_rho), Knot{0, 0, 0, 0});
    t.phi.assign(n * n * static_cast<std::size_t>(t.knots_r), Knot{0, 0, 0, 0});
    t.V.assign(n * n * static_cast<std::size_t>(t.knots_r), Knot{0, 0, 0, 0});

    for (std::size_t i = 0; i < n; ++i) {
      long atomic_number;
      rd.line() >> atomic_number >> t.mass[i];
      place(t.F, i, natural_spline(rd.block(t.n_rho, 5), t.d_rho));
      for (std::size_t j = 0; j < n; ++j) {
        place(t.phi, i + n * j, natural_spline(rd.block(t.n_r, 5), t.d_r));  // data.hpp:109-112
      }
    }
    // Pair tables hold r*V(r), lower triangle only (data.cpp:134-152).
    for (std::size_t i = 0; i < n; ++i) {
      for (std::size_t j = 0; j <= i; ++j) {
        std::vector<double> rv = rd.block(t.n_r, 5);
        for (std::size_t k = 0; k < rv.size(); ++k) {
          rv[k] /= t.d_r * static_cast<double>(k);
        }
        rv[0] = rv[1];
        place(t.V, std::max(i, j) + n * std::min(i, j), natural_spline(std::move(rv), t.d_r));  // data.hpp:114-116
      }
    }
    return t;
  }

}  // namespace ofly
