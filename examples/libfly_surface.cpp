// A libfly-style translation unit written against include/openfly_b200/libfly.hpp: namespace fly, the reference's own
// signatures (generic.hpp:79-136, list.hpp:86-161, lbfgs.hpp:96-118, dimer.hpp:148-155), system::SoA columns handed over
// without copies. Prints one JSON object that tests/test_cpp_shim.py compares with the CPU oracle.
//
//   libfly_surface --compile-only                               (no GPU: the surface instantiates and links)
//   libfly_surface <setfl> <cells> <replicas> <start_states.bin> (start states: [R][2][3N] doubles, positions then axes)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numbers>
#include <vector>

#include "openfly_b200/libfly.hpp"

using namespace fly;

int main(int argc, char **argv) {
  try {
    if (argc >= 2 && std::strcmp(argv[1], "--compile-only") == 0) {
      minimise::LBFGS::Options lo;
      saddle::Dimer::Options d;
      saddle::Rotor::Options r;
      system::SoA<Position, TypeID, Frozen> s(4);
      system::SoA<Position const &> v = s;  // owner -> view, no copy
      s(r_, 2)[1] = 7.0;
      std::printf("{\"lbfgs_n\": %d, \"dimer_n\": %d, \"rotor_n\": %d, \"view_sees_owner\": %d}\n", lo.n, d.n, r.n, v(r_, 2)[1] == 7.0);
      return 0;
    }
    if (argc < 5) return 2;
    const int cells = std::atoi(argv[2]);
    const long R = std::atol(argv[3]);
    const double a = 2.855, L = cells * a;

    // bcc motif tiled x-fastest, shifted by +pi (lattice.hpp:50-109); remove atom 1, add one H (apps/hydrogen.cpp:139)
    std::vector<double> pos;
    long site = 0;
    for (int z = 0; z < cells; z++)
      for (int y = 0; y < cells; y++)
        for (int x = 0; x < cells; x++)
          for (int m = 0; m < 2; m++, site++) {
            if (site == 1) continue;
            double f = m ? 0.5 : 0.0;
            pos.insert(pos.end(), {(x + f) * a + std::numbers::pi, (y + f) * a + std::numbers::pi, (z + f) * a + std::numbers::pi});
          }
    pos.insert(pos.end(), {2.857 / 2 + 3.14, 2.857 / 2 + 3.14, 2.857 / 4 + 3.14});
    const Index_t N = static_cast<Index_t>(pos.size() / 3);
    const std::size_t n3 = pos.size();

    system::SoA<Position, Axis, PotentialGradient, TypeID, Frozen> cell(N);
    for (std::size_t k = 0; k < n3; k++) cell[r_][static_cast<Index_t>(k)] = pos[k];
    for (Index_t i = 0; i < N; i++) cell(id_, i) = (i == N - 1) ? 1u : 0u, cell(fzn_, i) = 0;

    system::Box box({L, 0, 0, 0, L, 0, 0, 0, L}, {true, true, true});
    auto data = std::make_shared<potential::DataEAM const>(argv[1]);
    system::TypeMap map({"Fe", "H"});
    potential::Generic pot{potential::EAM{map, data}};

    neigh::List nl(box, pot.r_cut());
    nl.rebuild(cell);                       // SoA<...> -> SoA<Position const&>
    double E = pot.energy(cell, nl);        // SoA<...> -> SoA<TypeID const&, Frozen const&>
    pot.gradient(cell, cell, nl);
    double gmax = 0;
    for (std::size_t k = 0; k < n3; k++) gmax = std::max(gmax, std::fabs(cell[g_][static_cast<Index_t>(k)]));
    long neigh0 = 0, pairs = 0;
    nl.for_neighbours(0, pot.r_cut(), [&](Index_t, double, Vec const &) { neigh0++; });
    for (Index_t i = 0; i < N; i++) nl.for_neighbours(i, pot.r_cut(), [&](Index_t, double, Vec const &) { pairs++; });
    system::Hessian H;
    pot.hessian(H, cell, nl);
    double h00 = H(0, 0), h30 = H(3, 0);

    // a second List, other radius, other positions: `nl` keeps its own state (each List owns its device lists)
    system::SoA<Position> moved(cell);
    for (std::size_t k = 0; k < n3; k++) moved[r_][static_cast<Index_t>(k)] += 0.01 * std::sin(static_cast<double>(k));
    neigh::List nl2(box, pot.r_cut() + 0.3);
    nl2.rebuild(moved);
    double E2 = pot.energy(cell, nl2);
    double E_again = pot.energy(cell, nl);

    // update(): r <- r - x, then back
    system::SoA<Delta> dx(N);
    for (std::size_t k = 0; k < n3; k++) dx[del_][static_cast<Index_t>(k)] = 1e-3 * std::cos(static_cast<double>(k));
    nl.update(dx);
    double E_upd = pot.energy(cell, nl);
    for (std::size_t k = 0; k < n3; k++) dx[del_][static_cast<Index_t>(k)] *= -1;
    nl.update(dx);
    double E_back = pot.energy(cell, nl);

    minimise::LBFGS min(minimise::LBFGS::Options{}, box);
    system::SoA<Position, PotentialGradient> relaxed(N);
    bool failed = min.minimise(relaxed, cell, pot);
    nl.rebuild(relaxed);
    double E_min = pot.energy(cell, nl);

    // dimer searches from the given start states: the first through the single-system signature, all through the batch
    std::vector<double> starts(static_cast<std::size_t>(R) * 2 * n3);
    std::ifstream f(argv[4], std::ios::binary);
    f.read(reinterpret_cast<char *>(starts.data()), static_cast<std::streamsize>(starts.size() * sizeof(double)));
    if (!f) throw error("could not read the start states");
    saddle::Dimer dimer(saddle::Dimer::Options{}, saddle::Rotor::Options{}, box);
    system::SoA<Position, Axis, TypeID, Frozen> dim(cell);
    for (std::size_t k = 0; k < n3; k++) dim[r_][static_cast<Index_t>(k)] = starts[k], dim[ax_][static_cast<Index_t>(k)] = starts[n3 + k];
    std::vector<system::SoA<Position>> hist;
    auto exit0 = dimer.find_sp(dim, dim, relaxed, pot, hist, std::numbers::pi / 6, 0);

    std::vector<double> P(static_cast<std::size_t>(R) * n3), A(P.size()), oP(P.size()), oA(P.size());
    for (long r = 0; r < R; r++)
      for (std::size_t k = 0; k < n3; k++) P[static_cast<std::size_t>(r) * n3 + k] = starts[static_cast<std::size_t>(r) * 2 * n3 + k], A[static_cast<std::size_t>(r) * n3 + k] = starts[static_cast<std::size_t>(r) * 2 * n3 + n3 + k];
    auto exits = dimer.find_sp_batch(oP.data(), oA.data(), P.data(), A.data(), R, std::max(1L, R / 2), cell, relaxed, pot, hist, std::numbers::pi / 6, 0);
    double sp_diff = 0;
    for (std::size_t k = 0; k < n3; k++) sp_diff = std::max(sp_diff, std::fabs(oP[k] - dim[r_][static_cast<Index_t>(k)]));

    std::ofstream o(std::string(argv[4]) + ".fly", std::ios::binary);
    o.write(reinterpret_cast<char const *>(relaxed[r_].data()), static_cast<std::streamsize>(n3 * sizeof(double)));
    o.write(reinterpret_cast<char const *>(oP.data()), static_cast<std::streamsize>(oP.size() * sizeof(double)));

    std::printf("{\"atoms\": %ld, \"energy\": %.17g, \"grad_max\": %.17g, \"neigh0\": %ld, \"pairs\": %ld, \"h00\": %.17g, \"h30\": %.17g, "
                "\"energy_other_list\": %.17g, \"energy_again\": %.17g, \"energy_updated\": %.17g, \"energy_back\": %.17g, "
                "\"min_failed\": %d, \"energy_min\": %.17g, \"exit0\": %d, \"single_vs_batch_sp\": %.3g, \"exits\": [",
                static_cast<long>(N), E, gmax, neigh0, pairs, h00, h30, E2, E_again, E_upd, E_back, failed ? 1 : 0, E_min, static_cast<int>(exit0), sp_diff);
    for (std::size_t k = 0; k < exits.size(); k++) std::printf("%s%d", k ? ", " : "", exits[k]);
    std::printf("]}\n");
    return 0;
  } catch (Spline::OutOfBounds const &e) {
    std::fprintf(stderr, "out of bounds: %s\n", e.what());
    return 3;
  } catch (std::exception const &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
