// main.cc — `jams-b200`: run a JAMS configuration file on the B200 llg-heun path.
//   jams-b200 [--name NAME] [--output DIR] config.cfg ['patch string or file' ...]
// Mirrors the command line of the reference (core/args.cc, core/jams++.cc:231-377): every positional argument is a
// config file or a config string, merged left to right.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "jams_host.h"

int main(int argc, char **argv) {
  std::vector<std::string> configs;
  std::string name, output = ".";
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if ((a == "--name" || a == "-n") && i + 1 < argc) name = argv[++i];
    else if ((a == "--output" || a == "-o") && i + 1 < argc) output = argv[++i];
    else if (a == "--help" || a == "-h") { std::printf("usage: jams-b200 [--name NAME] [--output DIR] config.cfg [patch ...]\n"); return 0; }
    else configs.push_back(a);
  }
  if (configs.empty()) { std::fprintf(stderr, "jams-b200: no config file given\n"); return 1; }
  if (name.empty()) {   // simulation name = first config file without directory and extension (core/args.cc)
    name = configs[0];
    const size_t slash = name.find_last_of('/');
    if (slash != std::string::npos) name = name.substr(slash + 1);
    const size_t dot = name.find_last_of('.');
    if (dot != std::string::npos) name = name.substr(0, dot);
  }
  try {
    jams_b200::Simulation sim(configs, name, output);
    std::cout << "solver  " << sim.solver().name() << "\nspins   " << sim.lattice().num_spins << "\nsteps   " << sim.solver().max_steps() << std::endl;
    const auto t0 = std::chrono::steady_clock::now();
    sim.run();
    jb_synchronize(sim.solver().ctx());
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "runtime " << secs << " s (" << (double)sim.lattice().num_spins * sim.solver().iteration() / secs << " spin-updates/s)" << std::endl;
  } catch (const std::exception &e) {
    std::fprintf(stderr, "ERROR: %s\n", e.what());   // jams::die (helpers/error.h:11-23)
    return 1;
  }
  return 0;
}
