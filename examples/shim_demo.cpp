// Host program written against the C++20 shim (include/openfly_b200/fly.hpp) the way a libfly translation
// unit would use the reference classes: build a bcc Fe cell with a vacancy and an H atom, rebuild the list,
// evaluate energy / gradient, relax, and run a small batch of dimer searches. Prints one JSON object that
// tests/test_cpp_shim.py compares with the CPU oracle.
//
//   shim_demo --host-only <setfl>       no GPU: option defaults + host-side setfl parse through the C ABI
//   shim_demo <setfl> <cells> <replicas> <start_states.bin>   (start states: [R][2][3N] doubles, positions then axes)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <numbers>
#include <vector>

#include "openfly_b200/fly.hpp"

namespace fb = fly_b200;

int main(int argc, char **argv) {
  try {
    if (argc >= 3 && std::strcmp(argv[1], "--host-only") == 0) {
      fb::minimise::LBFGS::Options lo;
      fb::saddle::Dimer::Options d;
      fb::saddle::Rotor::Options r;
      long info[3];
      double dinfo[11];
      fb::detail::check(ofly_setfl_info_host(argv[2], info, dinfo));
      std::printf("{\"version\": \"%s\", \"lbfgs_n\": %d, \"dimer_n\": %d, \"rotor_n\": %d, \"types\": %ld, \"knots_r\": %ld, \"r_cut\": %.17g}\n",
                  ofly_version(), lo.n, d.n, r.n, info[0], info[2], dinfo[2]);
      return 0;
    }
    if (argc < 5) return 2;
    const int cells = std::atoi(argv[2]);
    const long R = std::atol(argv[3]);
    const double a = 2.855;

    // bcc motif tiled x-fastest, shifted by +pi (lattice.hpp:50-109); remove atom 1, add one H (apps/hydrogen.cpp:139)
    std::vector<double> pos;
    std::vector<std::uint32_t> id;
    long site = 0;
    for (int z = 0; z < cells; z++)
      for (int y = 0; y < cells; y++)
        for (int x = 0; x < cells; x++)
          for (int m = 0; m < 2; m++, site++) {
            if (site == 1) continue;
            double f = m ? 0.5 : 0.0;
            pos.insert(pos.end(), {(x + f) * a + std::numbers::pi, (y + f) * a + std::numbers::pi, (z + f) * a + std::numbers::pi});
            id.push_back(0);
          }
    pos.insert(pos.end(), {2.857 / 2 + 3.14, 2.857 / 2 + 3.14, 2.857 / 4 + 3.14});
    id.push_back(1);
    const std::size_t n3 = pos.size();

    fb::Context ctx(0);
    fb::potential::DataEAM data(ctx, argv[1]);
    fb::potential::Generic pot(data);
    fb::system::Box box({cells * a, cells * a, cells * a});

    fb::System one(data, box, id);
    fb::neigh::List nl(one, pot.r_cut());
    nl.rebuild(pos);
    double e0 = pot.energy(nl);
    std::vector<double> g(n3);
    pot.gradient(g, nl);
    double gmax = 0;
    for (double v : g) gmax = std::max(gmax, std::abs(v));
    long nn = 0;
    nl.for_neighbours(0, pot.r_cut(), [&](long, double, std::array<double, 3> const &) { nn++; });

    fb::minimise::LBFGS min(fb::minimise::LBFGS::Options{}, box);
    std::vector<double> rmin(n3), gmin(n3);
    bool failed = min.minimise(rmin, gmin, pos, one, pot);
    nl.rebuild(rmin);
    double e1 = pot.energy(nl);

    std::vector<double> start(static_cast<std::size_t>(R) * 2 * n3);
    std::ifstream in(argv[4], std::ios::binary);
    in.read(reinterpret_cast<char *>(start.data()), static_cast<std::streamsize>(start.size() * sizeof(double)));
    if (!in) throw fb::RuntimeError("start-state file is too short");
    std::vector<double> P(R * n3), A(R * n3), outP(R * n3), outA(R * n3);
    for (long r = 0; r < R; r++) {
      std::copy_n(start.begin() + (2 * r) * n3, n3, P.begin() + r * n3);
      std::copy_n(start.begin() + (2 * r + 1) * n3, n3, A.begin() + r * n3);
    }
    fb::System batch(data, box, id, {}, R);
    fb::saddle::Dimer dimer(fb::saddle::Dimer::Options{}, fb::saddle::Rotor::Options{}, box);
    auto exits = dimer.find_sp_batch(outP, outA, P, A, rmin, batch, pot, {}, std::numbers::pi / 6, 0, 1, false);

    std::printf("{\"atoms\": %zu, \"energy\": %.17g, \"grad_max\": %.17g, \"neigh0\": %ld, \"min_failed\": %d, \"energy_min\": %.17g, \"exits\": [",
                id.size(), e0, gmax, nn, failed ? 1 : 0, e1);
    for (std::size_t k = 0; k < exits.size(); k++) std::printf("%s%d", k ? ", " : "", static_cast<int>(exits[k]));
    std::printf("], \"launches\": %ld}\n", ctx.launch_count());
    std::ofstream out(std::string(argv[4]) + ".out", std::ios::binary);
    out.write(reinterpret_cast<char const *>(rmin.data()), static_cast<std::streamsize>(n3 * sizeof(double)));
    out.write(reinterpret_cast<char const *>(outP.data()), static_cast<std::streamsize>(outP.size() * sizeof(double)));
    return 0;
  } catch (fb::OutOfBounds const &e) {
    std::fprintf(stderr, "OutOfBounds: %s\n", e.what());
    return 3;
  } catch (std::exception const &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
