// Replica sharding from a C++ host: one host thread + one fly_b200::Context per
// device, global replica r runs on device r mod G, every device holds the shared immutable inputs (tables, types,
// the minimum), only exits and saddles come back and are merged in replica order -- no collective anywhere.
//
//   multi_gpu_batch <setfl> <cells> <replicas> <start_states.bin> [devices]
// start states: [R][2][3N] doubles (positions then axis per replica), as examples/shim_demo takes them.
// Prints one JSON object; writes the merged saddles to <start_states.bin>.mgpu.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <fstream>
#include <numbers>
#include <thread>
#include <vector>

#include "openfly_b200/fly.hpp"

namespace fb = fly_b200;

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  try {
    const int cells = std::atoi(argv[2]);
    const long R = std::atol(argv[3]);
    const int G = argc > 5 ? std::atoi(argv[5]) : 1;
    const double a = 2.855;
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
    fb::system::Box box({cells * a, cells * a, cells * a});

    std::vector<double> start(static_cast<std::size_t>(R) * 2 * n3);
    std::ifstream in(argv[4], std::ios::binary);
    in.read(reinterpret_cast<char *>(start.data()), static_cast<std::streamsize>(start.size() * sizeof(double)));
    if (!in) throw fb::RuntimeError("start-state file is too short");

    // the minimum every search starts from: relaxed once, on device 0
    std::vector<double> rmin(n3), gmin(n3);
    {
      fb::Context ctx(0);
      fb::potential::DataEAM data(ctx, argv[1]);
      fb::potential::Generic pot(data);
      fb::System one(data, box, id);
      fb::minimise::LBFGS min(fb::minimise::LBFGS::Options{}, box);
      if (min.minimise(rmin, gmin, pos, one, pot)) throw fb::RuntimeError("relaxation hit iter_max");
    }

    std::vector<int> exits(static_cast<std::size_t>(R), fb::saddle::Dimer::uninit);
    std::vector<double> saddles(static_cast<std::size_t>(R) * n3);
    std::vector<std::exception_ptr> errors(static_cast<std::size_t>(G));
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> workers;
    for (int d = 0; d < G; d++) {
      workers.emplace_back([&, d] {
        try {
          std::vector<long> mine;
          for (long r = d; r < R; r += G) mine.push_back(r);  // r mod G == d
          if (mine.empty()) return;
          const std::size_t m = mine.size();
          std::vector<double> P(m * n3), A(m * n3), outP(m * n3), outA(m * n3);
          for (std::size_t k = 0; k < m; k++) {
            std::copy_n(start.begin() + (2 * mine[k]) * n3, n3, P.begin() + k * n3);
            std::copy_n(start.begin() + (2 * mine[k] + 1) * n3, n3, A.begin() + k * n3);
          }
          fb::Context ctx(d);
          fb::potential::DataEAM data(ctx, argv[1]);
          fb::potential::Generic pot(data);
          fb::System batch(data, box, id, {}, static_cast<long>(m));
          fb::saddle::Dimer dimer(fb::saddle::Dimer::Options{}, fb::saddle::Rotor::Options{}, box);
          auto ex = dimer.find_sp_batch(outP, outA, P, A, rmin, batch, pot, {}, std::numbers::pi / 6, 0, 1, false);
          for (std::size_t k = 0; k < m; k++) {  // merge in replica order (disjoint slots: no lock needed)
            exits[static_cast<std::size_t>(mine[k])] = ex[k];
            std::copy_n(outP.begin() + k * n3, n3, saddles.begin() + mine[k] * n3);
          }
        } catch (...) {
          errors[static_cast<std::size_t>(d)] = std::current_exception();
        }
      });
    }
    for (auto &w : workers) w.join();
    for (auto &e : errors)
      if (e) std::rethrow_exception(e);
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    std::printf("{\"atoms\": %zu, \"devices\": %d, \"replicas\": %ld, \"seconds\": %.6f, \"exits\": [", id.size(), G, R, dt);
    for (long r = 0; r < R; r++) std::printf("%s%d", r ? ", " : "", exits[static_cast<std::size_t>(r)]);
    std::printf("]}\n");
    std::ofstream out(std::string(argv[4]) + ".mgpu", std::ios::binary);
    out.write(reinterpret_cast<char const *>(saddles.data()), static_cast<std::streamsize>(saddles.size() * sizeof(double)));
    return 0;
  } catch (std::exception const &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
