// b200_llg_heun.cc — see b200_llg_heun.h.  Compiles only inside the JAMS source tree (needs libconfig++ and
// the `globals` object graph); every citation is relative to src/jams/.
#include "jams/solvers/b200_llg_heun.h"

#if HAS_CUDA

#include <stdexcept>
#include <vector>

#include "jams/common.h"
#include "jams/core/globals.h"
#include "jams/core/lattice.h"
#include "jams/core/physics.h"
#include "jams/core/thermostat.h"
#include "jams/hamiltonian/applied_field.h"
#include "jams/hamiltonian/cuda_biquadratic_exchange.h"   // + `friend class B200HeunLLGSolver;`
#include "jams/hamiltonian/exchange.h"
#include "jams/hamiltonian/uniaxial_anisotropy.h"   // + `friend class B200HeunLLGSolver;`
#include "jams/hamiltonian/zeeman.h"                // + `friend class B200HeunLLGSolver;`
#include "jams/helpers/defaults.h"
#include "jams/interface/config.h"

namespace {
// the library draws its own Philox noise; this thermostat only carries T so that Solver::update_thermostat()
// (core/solver.cc:94-97) and monitors that ask thermostat()->temperature() keep working
class PassThroughThermostat : public Thermostat {
 public:
  PassThroughThermostat(double timestep, int num_spins) : Thermostat(0.0, 0.0, timestep, num_spins) {}
  void update() override {}
};
}  // namespace

B200HeunLLGSolver::~B200HeunLLGSolver() { jb_destroy(ctx_); }

void B200HeunLLGSolver::check(int status) const {
  if (status != JB_OK) throw std::runtime_error(std::string("jams_b200: ") + jb_last_error(ctx_));
}

void B200HeunLLGSolver::initialize(const libconfig::Setting &settings) {
  // same keys and conversions as CUDAHeunLLGSolver::initialize (solvers/cuda_llg_heun.cu:21-37)
  step_size_ = jams::config_required<double>(settings, "t_step") / 1e-12;
  auto t_max = jams::config_required<double>(settings, "t_max") / 1e-12;
  auto t_min = jams::config_optional<double>(settings, "t_min", 0.0) / 1e-12;
  max_steps_ = static_cast<int>(t_max / step_size_);
  min_steps_ = static_cast<int>(t_min / step_size_);
  gilbert_prefactor_ = jams::config_optional<bool>(settings, "gilbert_prefactor", false);  // core/lattice.cc:696-697
  rk4_ = lowercase(jams::config_required<std::string>(settings, "module")).find("rk4") != std::string::npos;   // "llg-rk4-b200-gpu"
  seed_ = static_cast<std::uint64_t>(jams::config_optional<int>(globals::config->lookup("sim"), "seed", 0));
  register_thermostat(new PassThroughThermostat(step_size_, globals::num_spins));

  jb_lattice_desc &d = desc_;
  d = jb_lattice_desc{};
  for (int n = 0; n < 3; ++n) {
    d.dims[n] = globals::lattice->size(n);
    d.periodic[n] = globals::lattice->is_periodic(n);
  }
  d.num_motif = globals::lattice->num_basis_sites();
  d.x_begin = 0; d.nx_local = d.dims[0]; d.rank = 0; d.n_ranks = 1; d.device = -1;
  if (jb_create(&ctx_, &d) != JB_OK) throw std::runtime_error(std::string("jams_b200: ") + jb_last_error(nullptr));
}

void B200HeunLLGSolver::build() {
  check(jb_set_materials(ctx_, globals::mus.data(), globals::gyro.data(), globals::alpha.data()));
  int ham_index = -1;   // hamiltonians_ are registered in config order (core/jams++.cc:284-288)
  // exchange, biquadratic, uniaxial, zeeman, applied field: the library holds one term of each kind, except uniaxial terms, which
  // go into up to three slots (K1 + K2 + K3 as separate modules: jb_set_uniaxial_term)
  int seen[5] = {0, 0, 0, 0, 0};
  auto once = [&](int kind, const std::string &hname) {
    if (seen[kind]++ >= (kind == 2 ? 3 : 1)) throw std::runtime_error("llg-heun-b200-gpu: one hamiltonian too many of the kind of '" + hname + "'; the fused solver holds one of each kind and three uniaxial terms (merge them, or use llg-heun-gpu)");
  };
  for (auto &h : hamiltonians_) {
    ++ham_index;
    const int kind = dynamic_cast<ExchangeHamiltonian *>(h.get()) ? 0 : dynamic_cast<CudaBiquadraticExchangeHamiltonian *>(h.get()) ? 1 :
                     dynamic_cast<UniaxialAnisotropyHamiltonian *>(h.get()) ? 2 : dynamic_cast<ZeemanHamiltonian *>(h.get()) ? 3 :
                     dynamic_cast<AppliedFieldHamiltonian *>(h.get()) ? 4 : -1;
    if (kind >= 0) once(kind, h->name());
    if (auto *ex = dynamic_cast<ExchangeHamiltonian *>(h.get())) {
      // ExchangeHamiltonian::neighbour_list() (hamiltonian/exchange.h:13): sorted {i,j} pairs + unique tensors.
      // The library recognises a translation-invariant list and switches to its template kernel.
      const auto &nbr = ex->neighbour_list();
      std::vector<int32_t> pi(nbr.size()), pj(nbr.size()), vid(nbr.size());
      std::vector<double> J9;
      std::vector<Mat3> uniq;
      for (int n = 0; n < nbr.size(); ++n) {
        const auto pr = nbr[n];
        pi[n] = pr.first[0]; pj[n] = pr.first[1];
        int v = 0;
        for (; v < static_cast<int>(uniq.size()); ++v) if (uniq[v] == pr.second) break;
        if (v == static_cast<int>(uniq.size())) {
          uniq.push_back(pr.second);
          for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) J9.push_back(pr.second[a][b]);
        }
        vid[n] = v;
      }
      check(jb_set_exchange_pairs(ctx_, nbr.size(), pi.data(), pj.data(), vid.data(), static_cast<int32_t>(uniq.size()), J9.data()));
    } else if (auto *bq = dynamic_cast<CudaBiquadraticExchangeHamiltonian *>(h.get())) {
      // CudaBiquadraticExchangeHamiltonian keeps its neighbour list (hamiltonian/cuda_biquadratic_exchange.h:40) and inserts
      // B_ij = unit * J[0][0] for every pair whose value exceeds the energy cutoff (cuda_biquadratic_exchange.cu:127-134).  The
      // library takes this term as a translation-invariant template: jb_detect_exchange_template turns the list into one.
      const auto &nbr = bq->neighbour_list_;
      std::vector<int32_t> pi, pj, vid;
      std::vector<double> J9;
      for (int n = 0; n < nbr.size(); ++n) {
        const auto pr = nbr[n];
        const double value = bq->input_energy_unit_conversion_ * pr.second[0][0];
        if (!(value > bq->energy_cutoff_ * bq->input_energy_unit_conversion_)) continue;
        int v = 0;
        for (; v < static_cast<int>(J9.size() / 9); ++v) if (J9[9 * v] == value) break;
        if (v == static_cast<int>(J9.size() / 9)) { J9.insert(J9.end(), 9, 0.0); J9[9 * v] = J9[9 * v + 4] = J9[9 * v + 8] = value; }
        pi.push_back(pr.first[0]); pj.push_back(pr.first[1]); vid.push_back(v);
      }
      const int cap = 4096;
      std::vector<int32_t> mi(cap), mj(cap), T3(3 * cap);
      std::vector<double> J9t(9 * static_cast<size_t>(cap));
      int32_t nt = -1;
      check(jb_detect_exchange_template(&desc_, static_cast<int64_t>(pi.size()), pi.data(), pj.data(), vid.data(), static_cast<int32_t>(J9.size() / 9),
                                        J9.data(), cap, &nt, mi.data(), mj.data(), T3.data(), J9t.data()));
      if (nt < 0) throw std::runtime_error("llg-heun-b200-gpu: the biquadratic-exchange list is not translation invariant; use llg-heun-gpu");
      std::vector<double> B(nt);
      for (int k = 0; k < nt; ++k) B[k] = J9t[9 * static_cast<size_t>(k)];
      check(jb_set_biquadratic_template(ctx_, nt, mi.data(), mj.data(), T3.data(), B.data()));
    } else if (auto *un = dynamic_cast<UniaxialAnisotropyHamiltonian *>(h.get())) {
      check(jb_set_uniaxial_term(ctx_, seen[2] - 1, un->power_, un->magnitude_.data(), un->axis_.data()));   // slot = how many came before
    } else if (auto *ze = dynamic_cast<ZeemanHamiltonian *>(h.get())) {
      check(jb_set_zeeman(ctx_, ze->dc_local_field_.data(),
                          ze->has_ac_local_field_ ? ze->ac_local_field_.data() : nullptr,
                          ze->has_ac_local_field_ ? ze->ac_local_frequency_.data() : nullptr));
    } else if (dynamic_cast<AppliedFieldHamiltonian *>(h.get())) {
      // the TimeDependentField member is protected and opaque (hamiltonian/applied_field.h:20-43): re-read the Hamiltonian's own
      // settings group with the conversions of applied_field.cc:13,35-38,63-68 (seconds -> ps, Hz -> THz)
      const libconfig::Setting &hs = globals::config->lookup("hamiltonians")[ham_index];
      const auto type = lowercase(jams::config_optional<std::string>(hs, "type", "static"));
      const Vec3 B = jams::config_required<Vec3>(hs, "field");
      const double Bv[3] = {B[0], B[1], B[2]};
      if (type == "static") {
        check(jb_set_applied_field(ctx_, Bv, 1));
      } else if (type == "sinc" || type == "sinc-cos") {
        check(jb_set_applied_field_pulse(ctx_, Bv, type == "sinc" ? JB_FIELD_SINC : JB_FIELD_SINC_COS,
                                         jams::config_required<double>(hs, "time_center") / 1e-12,
                                         jams::config_required<double>(hs, "freq_bandwidth") / 1e12,
                                         type == "sinc-cos" ? jams::config_required<double>(hs, "freq_center") / 1e12 : 0.0));
      } else {
        throw std::runtime_error("Unknown field pulse type " + type);
      }
    } else {
      // exchange-functional / exchange-neartree keep their matrix private (hamiltonian/sparse_interaction.h:45-49): INTEGRATION.md
      throw std::runtime_error("llg-heun-b200-gpu: hamiltonian '" + h->name() + "' is not fused; use llg-heun-gpu");
    }
  }
  // The first Solver::update_physics_module() of the main loop (core/jams++.cc:334) has already run: globals::s carries the
  // first pinning rotation.  From here on the state lives in the library and the pinning is applied to it there.
  import_spins();
  if (lowercase(jams::config_optional<std::string>(globals::config->lookup("physics"), "module", "empty")) == "pinned_boundaries")
    setup_pinned_boundaries();
  built_ = true;
}

// PinnedBoundariesPhysics::update (physics/pinned_boundaries.cc:34-46) rotates globals::s in place between steps, and
// Solver::update_physics_module() is not virtual (core/solver.h:57), so the adapter cannot redirect it.  Re-importing globals::s
// before every step would throw away the library's progress on every iteration without a monitor export, and
// exporting after every step would add 96 B per spin to a 120 B step.  Instead the adapter registers the same edge regions with
// the library and applies the same operation -- sum of mu_i s_i over the region, rotation_matrix_between_vectors, s_i <- R s_i --
// to the library's state at the end of run(), i.e. at the point of the main loop where the physics module acts.  The module's own
// rotation of globals::s then only touches a copy that the next export overwrites.
void B200HeunLLGSolver::setup_pinned_boundaries() {
  const libconfig::Setting &ps = globals::config->lookup("physics");
  static const char *names[6] = {"left", "right", "front", "back", "bottom", "top"};   // pinned_boundaries.h:91-107: dimension, upper
  const Vec3i size = globals::lattice->size();
  for (int k = 0; k < 6; ++k) {
    const std::string name = names[k];
    if (!ps.exists(name + "_pinned_magnetisation")) continue;
    const Vec3 target = jams::config_required<Vec3>(ps, name + "_pinned_magnetisation");
    const int cells = jams::config_optional<int>(ps, name + "_pinned_cells", 1);
    const int dim = k / 2;
    const bool upper = (k % 2) == 1;
    std::vector<int32_t> sites;
    for (int i = 0; i < globals::num_spins; ++i) {
      const Vec3i cell = globals::lattice->cell_offset(i);
      if (upper ? cell[dim] >= size[dim] - cells : cell[dim] < cells) sites.push_back(i);
    }
    const int region = static_cast<int>(pinned_.size());
    check(jb_set_region(ctx_, region, static_cast<int32_t>(sites.size()), sites.data()));
    pinned_.push_back({region, {target[0], target[1], target[2]}});
  }
}

void B200HeunLLGSolver::apply_pinned_boundaries() {
  for (const auto &b : pinned_) {
    double M4[4];
    check(jb_region_moment(ctx_, b.region, M4));
    const Mat3 R = rotation_matrix_between_vectors(Vec3{M4[0], M4[1], M4[2]}, Vec3{b.magnetisation[0], b.magnetisation[1], b.magnetisation[2]});
    const double R9[9] = {R[0][0], R[0][1], R[0][2], R[1][0], R[1][1], R[1][2], R[2][0], R[2][1], R[2][2]};
    check(jb_rotate_region(ctx_, b.region, R9));
  }
}

void B200HeunLLGSolver::import_spins() {
  // const device pointer: does not dirty the host copy (containers/synced_memory.h:472-478)
  check(jb_import_spins(ctx_, const_cast<const jams::MultiArray<double, 2> &>(globals::s).device_data(), /*on_device=*/1));
  spins_exported_ = true;
}

void B200HeunLLGSolver::export_spins() {
  if (spins_exported_) return;
  check(jb_export_spins(ctx_, globals::s.device_data(), /*on_device=*/1));  // non-const: host copy becomes stale
  check(jb_synchronize(ctx_));
  spins_exported_ = true;
}

void B200HeunLLGSolver::run() {
  if (!built_) build();
  update_thermostat();                               // T is re-read every step (core/solver.cc:94-97)
  check((rk4_ ? jb_step_rk4 : jb_step)(ctx_, 1, step_size_, time_, thermostat_->temperature(), seed_,
                                       static_cast<std::uint64_t>(iteration_), gilbert_prefactor_ ? 1 : 0));
  spins_exported_ = false;
  iteration_++;
  time_ = iteration_ * step_size_;                   // solvers/cuda_llg_heun.cu:120-121
  apply_pinned_boundaries();                         // what update_physics_module() does next in the main loop (core/jams++.cc:334)
}

void B200HeunLLGSolver::notify_monitors() {
  if (!built_) build();
  bool any = false;
  for (auto &m : monitors_) any = any || m->is_updating(iteration_);
  if (!any) return;
  export_spins();                                    // 48 B/spin, only on output steps
  Solver::notify_monitors();                         // core/solver.cc:110-116
}

void B200HeunLLGSolver::compute_fields() {
  if (!built_) build();
  check(jb_fields(ctx_, JB_TERM_TOTAL, time_, globals::h.device_data(), /*on_device=*/1));
}

#endif  // HAS_CUDA
