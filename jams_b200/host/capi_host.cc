// capi_host.cc — plain-C entry points of the C++ host layer (libjams_b200_host.so) so that the Python tests can
// drive it through ctypes: config parsing/merging, lattice + interaction-template construction (no GPU needed) and a
// complete run of a JAMS configuration on the GPU.
#include <cstring>
#include <string>
#include <vector>

#include "jams_host.h"

using namespace jams_b200;

namespace {
thread_local std::string g_error;
int fail(const std::exception &e) { g_error = e.what(); return 1; }
int copy_out(const std::string &s, char *out, long long capacity) {
  if ((long long)s.size() + 1 > capacity) { g_error = "output buffer too small (" + std::to_string(s.size() + 1) + " bytes needed)"; return 2; }
  std::memcpy(out, s.c_str(), s.size() + 1);
  return 0;
}
std::vector<std::string> split_args(const char *const *args, int n) { return std::vector<std::string>(args, args + n); }
}  // namespace

extern "C" {

__attribute__((visibility("default"))) const char *jbh_last_error() { return g_error.c_str(); }

// merge `n` config arguments (file names or config strings, core/jams++.cc:48-84) and render the result as JSON
__attribute__((visibility("default"))) int jbh_config_to_json(const char *const *args, int n, char *out, long long capacity) {
  try {
    return copy_out(parse_config_strings(split_args(args, n))->to_json(), out, capacity);
  } catch (const std::exception &e) { return fail(e); }
}

// lattice summary + per-site arrays sizes: {"num_spins":..,"M":..,"dims":[..]}
__attribute__((visibility("default"))) int jbh_lattice_info(const char *const *args, int n, int *num_spins, int *num_motif, int *dims3, int *periodic3) {
  try {
    auto cfg = parse_config_strings(split_args(args, n));
    Lattice lat(*cfg);
    *num_spins = lat.num_spins; *num_motif = lat.M;
    for (int k = 0; k < 3; ++k) { dims3[k] = lat.dims[k]; periodic3[k] = lat.periodic[k] ? 1 : 0; }
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}

// per-site arrays of the lattice a config describes (each may be NULL): mus, gyro, alpha (N), spins, positions (N x 3)
__attribute__((visibility("default"))) int jbh_lattice_arrays(const char *const *args, int n, double *mus, double *gyro, double *alpha, double *spins, double *positions) {
  try {
    auto cfg = parse_config_strings(split_args(args, n));
    Lattice lat(*cfg);
    auto put = [](const std::vector<double> &v, double *dst) { if (dst) std::memcpy(dst, v.data(), v.size() * sizeof(double)); };
    put(lat.mus(), mus); put(lat.gyro(), gyro); put(lat.alpha(), alpha); put(lat.initial_spins(0), spins); put(lat.positions(), positions);
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}

// processed exchange template of hamiltonians[ham_index] (must be module "exchange"); arrays sized by `capacity` entries
__attribute__((visibility("default"))) int jbh_exchange_template(const char *const *args, int n, int ham_index, int capacity, int *n_entries,
                                                                 int32_t *mi, int32_t *mj, int32_t *T3, double *J9, long long *n_pairs) {
  try {
    auto cfg = parse_config_strings(split_args(args, n));
    Lattice lat(*cfg);
    std::unique_ptr<Hamiltonian> ham(Hamiltonian::create((*cfg)["hamiltonians"][ham_index], lat));   // exchange or exchange-functional
    ExchangeHamiltonian *hp = dynamic_cast<ExchangeHamiltonian *>(ham.get());
    if (!hp) { g_error = "hamiltonian " + std::to_string(ham_index) + " is not an exchange hamiltonian"; return 1; }
    ExchangeHamiltonian &h = *hp;
    const InteractionTemplate &t = h.interaction_template();
    *n_entries = t.size();
    if (t.size() > capacity) { g_error = "template capacity too small"; return 2; }
    std::memcpy(mi, t.mi.data(), t.mi.size() * sizeof(int32_t)); std::memcpy(mj, t.mj.data(), t.mj.size() * sizeof(int32_t));
    std::memcpy(T3, t.T3.data(), t.T3.size() * sizeof(int32_t)); std::memcpy(J9, t.J9.data(), t.J9.size() * sizeof(double));
    if (n_pairs) *n_pairs = (long long)h.neighbour_list().i.size();   // "computed interactions: N" (hamiltonian/exchange.cc:152)
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}

// run a configuration to completion on the GPU (writes <output_dir>/<name>_mag.tsv etc.); returns the final spins (N x 3)
// if `spins_out` is not NULL; `max_steps_override` >= 0 stops after that many steps (negative: run to t_max)
__attribute__((visibility("default"))) int jbh_run(const char *const *args, int n, const char *name, const char *output_dir, int max_steps_override,
                                                   double *spins_out, int *steps_done) {
  try {
    Simulation sim(split_args(args, n), name, output_dir);
    B200HeunLLGSolver &s = sim.solver();
    int steps = 0;
    while (s.is_running() && (max_steps_override < 0 || steps < max_steps_override)) {   // core/jams++.cc:333-341
      s.update_physics_module();
      s.notify_monitors();
      s.run();
      ++steps;
    }
    s.post_process_monitors();   // core/jams++.cc:360-362
    if (steps_done) *steps_done = steps;
    if (spins_out) { const std::vector<double> sp = s.spins(); std::memcpy(spins_out, sp.data(), sp.size() * sizeof(double)); }
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}

}  // extern "C"
