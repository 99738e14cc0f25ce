// Host-side EAM table construction for the B200 path: LAMMPS setfl (eam/fs) reader and the
// natural-cubic-spline coefficient build. Arithmetic follows the reference so the knots are
// bit-identical to what fly::potential::DataEAM holds:
//   parser  src/potential/EAM/data.cpp:67-153 (type order, "for i: F_i, phi_ij" blocks, r*V tables)
//   spline  src/utility/spline.cpp:27-103 (5 padded knots, natural ends, tridiagonal sweep)
// Compiled with -ffp-contract=off so no FMA changes the coefficient bits.
#pragma once

#include <istream>
#include <string>
#include <vector>

namespace ofly {

  struct Knot {
    double a, b, c, d;
  };

  struct HostTables {
    int n_types = 0;
    long n_rho = 0, n_r = 0;      // tabulated points in the file (numP, numR)
    long knots_rho = 0, knots_r = 0;  // spline intervals = points + 4
    double d_rho = 0, d_r = 0, r_cut = 0;
    std::vector<std::string> names;
    std::vector<double> mass;
    // F: [type][knots_rho]; phi: [src + n*dst][knots_r]; V: [max + n*min][knots_r] (other slots zero)
    std::vector<Knot> F, phi, V;
  };

  // Throws std::runtime_error on malformed input.
  HostTables read_setfl(std::string const &path);
  // DataEAM(std::istream) (data.cpp:67): the same parse from any stream.
  HostTables read_setfl(std::istream &in);

  // Coefficients of the natural cubic spline through `y` (uniform spacing dx), y.size()+4 knots.
  std::vector<Knot> natural_spline(std::vector<double> y, double dx);

}  // namespace ofly
