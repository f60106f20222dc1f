// jams_host.h — C++ host layer above the C ABI (include/jams_b200.h): the JAMS plugin surface for the llg-heun
// hot path, driven by JAMS configuration files, without the JAMS source tree.
//
// Same names, settings keys, units and error behaviour as the reference classes, so that a JAMS user (and the
// adapter in integration/) finds the familiar surface:
//
//   reference (src/jams/...)                                  here (namespace jams_b200)
//   --------------------------------------------------------  ------------------------------------------------
//   Material            containers/material.h:16-60            Material
//   Lattice             core/lattice.cc:196-207,337-470,        Lattice: materials, unit cell, motif, supercell,
//                       577-756,987-1153                        site numbering, boundary wrap, per-site arrays
//   interaction files   core/interactions.cc:24-124,292-395     Lattice::expand_interactions / neighbour_list
//   Hamiltonian         core/hamiltonian.h:15-78                Hamiltonian (+ create())
//   ExchangeHamiltonian hamiltonian/exchange.cc:12-172          ExchangeHamiltonian
//   UniaxialAnisotropy… hamiltonian/uniaxial_anisotropy.cc      UniaxialAnisotropyHamiltonian
//   ZeemanHamiltonian   hamiltonian/zeeman.cc:12-132            ZeemanHamiltonian
//   AppliedField…       hamiltonian/applied_field.cc:84-148     AppliedFieldHamiltonian (static field)
//   Physics             core/physics.h:14-40                    Physics (temperature, applied_field)
//   Solver              core/solver.h:15-90                     Solver
//   CUDAHeunLLGSolver   solvers/cuda_llg_heun.cu:21-122         B200HeunLLGSolver ("llg-heun-b200-gpu", alias of
//                                                               "llg-heun-gpu" / "llg-heun-cpu" here)
//   Monitor             core/monitor.h:20-95                    Monitor
//   MagnetisationMonitor monitors/magnetisation.cc:21-141       MagnetisationMonitor (mag.tsv, same columns/format)
//   EnergyMonitor       monitors/energy.cc:16-46                EnergyMonitor (eng.tsv)
//   initializer         initializer/init_bloch_domain_wall.cc   Simulation::run_initializer (bloch_domain_wall)
//   main loop           core/jams++.cc:231-377                  Simulation::{initialize,run}
//
// All numerics happen in libjams_b200.so; nothing here computes fields or integrates spins.
#ifndef JAMS_B200_HOST_JAMS_HOST_H
#define JAMS_B200_HOST_JAMS_HOST_H

#include <array>
#include <cstdint>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "config.h"
#include "jams_b200.h"

namespace jams_b200 {

// helpers/consts.h:8-44 (internal units: ps, T, meV)
constexpr double kBohrMagnetonIU = 0.0578838181;
constexpr double kGyromagneticRatioIU = 0.17608596;
constexpr double kJoule2meV = 6.24150907e21;
constexpr double kLatticeTolerance = 1e-4;   // helpers/defaults.h:44

using Vec3 = std::array<double, 3>;
using Mat3 = std::array<std::array<double, 3>, 3>;   // row-major

double energy_unit_conversion(const std::string &name);   // core/units.h:15-26 via core/hamiltonian.cc:117-147

struct Material {   // containers/material.h:16-60
  std::string name;
  double moment = 0.0;   // meV/T
  double gyro = kGyromagneticRatioIU;
  double alpha = 0.01;
  Vec3 spin{{0.0, 0.0, 1.0}};
  bool randomize = false;
};

struct InteractionTemplate {   // processed template: post_process_interactions (core/interactions.cc:292-347)
  std::vector<int32_t> mi, mj, T3;
  std::vector<double> J9;      // meV, row-major 3x3 per entry
  int size() const { return static_cast<int>(mi.size()); }
};

struct NeighbourList {   // jams::InteractionList<Mat3,2> contents (containers/interaction_list.h:28-92)
  std::vector<int32_t> i, j, value_id;
  std::vector<double> values9;
};

struct InteractionInput {   // one line of `interactions = (...)` / an exc_file (core/interactions.cc:128-251)
  bool by_motif = false;   // KKR format: 1-based motif positions instead of material names
  std::string type_i, type_j;
  int motif_i = 0, motif_j = 0;
  Vec3 r{{0, 0, 0}};
  std::array<double, 9> J9{};
};

class Lattice {
 public:
  explicit Lattice(const Setting &config);   // reads `materials`, `unitcell`, `lattice`, `solver.gilbert_prefactor`

  int dims[3] = {1, 1, 1};
  bool periodic[3] = {true, true, true};
  int M = 0;
  int num_spins = 0;
  bool gilbert_prefactor = false;
  double lattice_parameter = 0.0;
  std::vector<Material> materials;
  Mat3 cell{}, cell_inv{};                 // columns are the lattice vectors a, b, c (core/lattice.cc:356-367)
  std::vector<int> motif_material;
  std::vector<Vec3> motif_frac;
  std::vector<Mat3> rotations;             // rotations of the crystal's zero-translation symmetry operations, fractional basis (find_point_operations; spglib in JAMS)

  int material_index(const std::string &name) const;   // throws if unknown
  bool material_exists(const std::string &name) const;
  int site_index(int i, int j, int k, int m) const { return ((i * dims[1] + j) * dims[2] + k) * M + m; }
  double max_interaction_radius() const;   // Lattice::max_interaction_radius (core/lattice.cc:1009-1011,1155-1200), lattice parameters

  // per-site arrays in reference site order (globals::mus/gyro/alpha/s/positions, core/lattice.cc:688-756)
  std::vector<double> mus() const, gyro() const, alpha() const;
  std::vector<double> initial_spins(uint64_t seed) const;   // N x 3; "random" materials use a seeded uniform-on-sphere draw;
                                                            // lattice.spins = "file" replaces them (core/lattice.cc:738-748)
  bool has_impurities = false;                              // lattice.impurities (core/lattice.cc:424-427): random substitution of materials
  uint64_t impurities_seed = 0;
  int material_of_site(int i) const { return site_materials_.empty() ? motif_material[i % M] : site_materials_[i]; }
  std::string spins_file;                                   // lattice.spins
  std::vector<int32_t> site_materials_;                     // per site, only with impurities
  mutable long long snapshot_iteration = 0;                 // iteration recorded in the header of a loaded snapshot (0 = none)
  std::vector<double> positions() const;                    // N x 3, lattice constants
  std::vector<int32_t> site_material() const, site_motif() const;

  InteractionTemplate expand_interactions(const std::vector<InteractionInput> &interactions, double unit, bool fractional,
                                          bool use_symops, double energy_cutoff, double radius_cutoff,
                                          double distance_tolerance, double prefactor) const;
  NeighbourList neighbour_list(const InteractionTemplate &t) const;   // neighbour_list_from_interactions (core/interactions.cc:349-395)

 private:
  std::vector<Mat3> point_group_of_motif(int m) const;   // core/lattice.cc:1127-1153
};

class B200HeunLLGSolver;

class Hamiltonian {   // core/hamiltonian.h:15-78
 public:
  Hamiltonian(const Setting &settings, const Lattice &lattice);
  virtual ~Hamiltonian() = default;
  static Hamiltonian *create(const Setting &settings, const Lattice &lattice);   // core/hamiltonian.cc:80-115

  const std::string &name() const { return name_; }
  virtual int term() const = 0;
  virtual void attach(jb_ctx *ctx) = 0;   // hand the parameters to the library

  std::vector<double> calculate_fields(double time);          // N x 3 meV
  std::vector<double> calculate_energies(double time);        // N
  double calculate_total_energy(double time);

  B200HeunLLGSolver *solver = nullptr;

 protected:
  const Lattice &lattice_;
  std::string name_;
  std::string input_energy_unit_name_;
  double input_energy_unit_conversion_ = 1.0;
};

class ExchangeHamiltonian : public Hamiltonian {   // hamiltonian/exchange.cc:12-172
 public:
  ExchangeHamiltonian(const Setting &settings, const Lattice &lattice);
  int term() const override { return JB_TERM_EXCHANGE; }
  void attach(jb_ctx *ctx) override;
  const InteractionTemplate &interaction_template() const { return template_; }
  NeighbourList neighbour_list() const { return lattice_.neighbour_list(template_); }
 protected:
  struct NoParse {};
  ExchangeHamiltonian(const Setting &settings, const Lattice &lattice, NoParse) : Hamiltonian(settings, lattice) {
    check_symmetry_ = settings.get("check_sparse_matrix_symmetry", true);
  }
  void parse_interactions(const Setting &settings, const Lattice &lattice, double prefactor);   // 'interactions' / 'exc_file' -> template_
  InteractionTemplate template_;
  bool check_symmetry_ = true;
};

// hamiltonian/cuda_biquadratic_exchange.{h,cu}: E = -1/2 sum_ij B_ij (s_i . s_j)^2, field h_i = sum_j 2 B_ij s_j (s_i . s_j); the
// exchange grammar with the scalar B_ij = J[0][0]
class BiquadraticExchangeHamiltonian : public ExchangeHamiltonian {
 public:
  BiquadraticExchangeHamiltonian(const Setting &settings, const Lattice &lattice);
  int term() const override { return JB_TERM_BIQUADRATIC; }
  void attach(jb_ctx *ctx) override;
 private:
  std::vector<double> B_;   // meV per template entry
};

// hamiltonian/exchange_functional.{h,cc}: isotropic J(r_ij) from a closed form inside a cutoff radius per ordered material pair
// (step, exponential, gaussian, gaussian_multi, kaneyoshi, rkky, c3z).  The reference fills the same scalar CSR matrix as
// `exchange` through a near-tree walk over the supercell; on a lattice without impurities that list is translation invariant,
// so it is generated as a template for the same kernels.
class ExchangeFunctionalHamiltonian : public ExchangeHamiltonian {
 public:
  ExchangeFunctionalHamiltonian(const Setting &settings, const Lattice &lattice);
};

// hamiltonian/exchange_neartree.{h,cc}: isotropic exchange by distance shells, interactions = ((A, B, radius, J), ...): every A site
// couples with J to the B sites at distance radius -+ shell_width / 2 (mirrored entry for A != B), |J| <= energy_cutoff dropped
class ExchangeNeartreeHamiltonian : public ExchangeHamiltonian {
 public:
  ExchangeNeartreeHamiltonian(const Setting &settings, const Lattice &lattice);
};

class UniaxialAnisotropyHamiltonian : public Hamiltonian {   // hamiltonian/uniaxial_anisotropy.cc:79-172
 public:
  UniaxialAnisotropyHamiltonian(const Setting &settings, const Lattice &lattice);
  int term() const override { return slot_ == 0 ? JB_TERM_UNIAXIAL : (slot_ == 1 ? JB_TERM_UNIAXIAL_2 : JB_TERM_UNIAXIAL_3); }
  void attach(jb_ctx *ctx) override;
  void set_slot(int slot) { slot_ = slot; }   // the n-th uniaxial Hamiltonian of the configuration (jb_set_uniaxial_term)
 private:
  int slot_ = 0;
  int power_ = 2;
  std::vector<double> magnitude_, axis_;
};

class ZeemanHamiltonian : public Hamiltonian {   // hamiltonian/zeeman.cc:12-132
 public:
  ZeemanHamiltonian(const Setting &settings, const Lattice &lattice);
  int term() const override { return JB_TERM_ZEEMAN; }
  void attach(jb_ctx *ctx) override;
 private:
  std::vector<double> dc_local_field_, ac_local_field_, ac_local_frequency_;
  bool has_ac_local_field_ = false;
};

class AppliedFieldHamiltonian : public Hamiltonian {   // hamiltonian/applied_field.cc:10-148 (types static, sinc, sinc-cos)
 public:
  AppliedFieldHamiltonian(const Setting &settings, const Lattice &lattice);
  int term() const override { return JB_TERM_APPLIED; }
  void attach(jb_ctx *ctx) override;
 private:
  Vec3 field_{{0, 0, 0}};
  int type_ = JB_FIELD_STATIC;
  double time_center_ = 0.0, freq_bandwidth_ = 0.0, freq_center_ = 0.0;   // ps, THz (applied_field.cc:37-38,66-68)
};

class B200HeunLLGSolver;

// core/physics.h:14-40 + physics/empty.h: constant temperature and applied field from `physics`;
// module "pinned_boundaries" (physics/pinned_boundaries.{h,cc}): every iteration the spins of each edge region are rotated
// on the device so that the region's moment points along the pinned direction (jb_region_moment / jb_rotate_region);
// modules "field-cool" (physics/field_cool.cc:9-78) and "two-temperature-model" (physics/two_temperature_model.cc:14-96):
// temperature ramps, host arithmetic only -- the solver hands physics()->temperature() to jb_step every iteration
// (core/solver.cc:94-97)
class Physics {
 public:
  explicit Physics(const Setting *settings, const Setting *sim = nullptr, const std::string &output_prefix = "");
  double temperature() const { return temperature_; }
  double applied_field(int i) const { return applied_field_[i]; }
  void update(B200HeunLLGSolver &solver);   // Physics::update, once per iteration before the monitors (core/jams++.cc:334)
 private:
  struct PinnedBoundary { int dim; bool upper; int cells; Vec3 magnetisation; };
  double temperature_ = 0.0;
  Vec3 applied_field_{{0, 0, 0}};
  std::vector<PinnedBoundary> boundaries_;
  bool regions_set_ = false;
  // field-cool
  bool field_cool_ = false, step_toggle_ = false;
  double init_temp_ = 0, final_temp_ = 0, cool_time_ = 0, integration_time_step_ = 0, t_eq_ = 0, delta_T_ = 0, t_plateau_ = 0;
  int t_steps_ = 0;
  Vec3 init_field_{{0, 0, 0}}, final_field_{{0, 0, 0}};
  // two-temperature model
  bool ttm_ = false;
  std::vector<double> pulse_width_, pulse_fluence_, pulse_start_;
  double electron_temp_ = 0, phonon_temp_ = 0, sink_temp_ = 0, Ce_ = 7.0e2, Cl_ = 3.0e6, G_ = 17.0e17, Gsink_ = 17.0e14;
  Vec3 reversing_field_{{0, 0, 0}};
  int output_step_freq_ = 100;
  std::ofstream ttm_file_;
};

class Solver;

class Monitor {   // core/monitor.h:20-95
 public:
  explicit Monitor(const Setting &settings);
  virtual ~Monitor() = default;
  static Monitor *create(const Setting &settings, const Lattice &lattice, const std::string &output_prefix);
  bool is_updating(int iteration) const { return iteration % output_step_freq_ == 0; }
  virtual void update(B200HeunLLGSolver &solver) = 0;
  virtual void post_process() {}
 protected:
  int output_step_freq_ = 100;   // helpers/defaults.h:28
};

class MagnetisationMonitor : public Monitor {   // monitors/magnetisation.cc:21-141
 public:
  MagnetisationMonitor(const Setting &settings, const Lattice &lattice, const std::string &filename);
  void update(B200HeunLLGSolver &solver) override;
  std::string tsv_header() const;
 private:
  const Lattice &lattice_;
  enum class Grouping { NONE, MATERIALS, POSITIONS } grouping_ = Grouping::MATERIALS;
  bool normalize_ = true;
  int n_groups_ = 1;
  std::vector<int32_t> group_of_spin_;
  const jb_ctx *registered_ctx_ = nullptr;   // context that holds the device copy of group_of_spin_
  std::ofstream tsv_file_;
};

// Spin snapshots for checkpoint / restart.  The reference's `hdf5` monitor writes <name>_NNNNNNN.h5 every output_steps and
// <name>_final.h5 in post_process with the dataset /spins (N x 3 f64; monitors/hdf5.cc:28-52,64-78,94-150) and restarts through
// lattice.spins = "file" (core/lattice.cc:738-748).  HDF5 is not available to this build, so module "hdf5" (alias
// "spins-tsv") writes the same snapshots as whitespace-separated text, <name>_NNNNNNN.tsv / <name>_final.tsv -- the format the
// reference's loader accepts for any extension other than .h5 (helpers/load.h:21-61): one spin per line, 17 significant digits
// (round-trips exactly), '#' comment header.
class SpinsTsvMonitor : public Monitor {
 public:
  SpinsTsvMonitor(const Setting &settings, const std::string &prefix) : Monitor(settings), prefix_(prefix) {}
  void update(B200HeunLLGSolver &solver) override;
  void post_process() override;
  static void write(const std::string &filename, const std::vector<double> &s_aos, int iteration, double time);
 private:
  std::string prefix_;
  std::vector<double> last_;
  int last_iteration_ = 0;
  double last_time_ = 0.0;
  B200HeunLLGSolver *final_from_ = nullptr;   // the final snapshot is taken from the solver that was last seen
};

// monitors/magnetisation_layers.{h,cc}: the moment of every layer of spins along `layer_normal` (Bohr magnetons), per group, every
// output_steps -- what examples/bloch_domain_wall uses for the wall profile.  The reference writes the groups' layer tables and
// the time series into monitors.h5; HDF5 is not available to this build, so the same data go to <name>_mag_layers.tsv:
//   "# group <name> layer <k> position_nm <z> saturation_moment_muB <m> spins <n>"   once per layer, then per update
//   "<iteration> <time_ps> <group> <layer> <mx> <my> <mz>"
class MagnetisationLayersMonitor : public Monitor {
 public:
  MagnetisationLayersMonitor(const Setting &settings, const Lattice &lattice, const std::string &filename);
  void update(B200HeunLLGSolver &solver) override;
  // layer tables of group g (tests read them back through jbh_* or the file)
  const std::vector<double> &layer_positions(size_t g) const { return layer_positions_[g]; }
 private:
  const Lattice &lattice_;
  std::vector<std::string> group_names_;
  std::vector<std::vector<std::vector<int>>> group_layer_spins_;   // [group][layer] -> site ids
  std::vector<std::vector<double>> layer_positions_;
  std::vector<double> mus_;
  std::ofstream tsv_file_;
};

class EnergyMonitor : public Monitor {   // monitors/energy.cc:16-46
 public:
  EnergyMonitor(const Setting &settings, const std::string &filename);
  void update(B200HeunLLGSolver &solver) override;
 private:
  std::string filename_;
  std::ofstream tsv_file_;
  bool header_written_ = false;
};

class B200HeunLLGSolver {   // core/solver.h:15-90 + solvers/cuda_llg_heun.cu:21-122
 public:
  B200HeunLLGSolver(const Setting &settings, const Lattice &lattice, uint64_t seed);
  ~B200HeunLLGSolver();

  std::string name() const { return rk4_ ? "llg-rk4-b200-gpu" : "llg-heun-b200-gpu"; }
  bool is_cuda_solver() const { return true; }
  bool is_running() const { return iteration_ < max_steps_; }
  int iteration() const { return iteration_; }
  double time() const { return time_; }
  double time_step() const { return step_size_; }
  int max_steps() const { return max_steps_; }
  const Physics *physics() const { return physics_.get(); }

  void register_physics_module(Physics *p) { physics_.reset(p); }
  void update_physics_module() { physics_->update(*this); }   // core/solver.cc:99-108
  void register_hamiltonian(Hamiltonian *h);
  void register_monitor(Monitor *m) { monitors_.emplace_back(m); }
  std::vector<std::unique_ptr<Hamiltonian>> &hamiltonians() { return hamiltonians_; }

  void set_spins(const std::vector<double> &s_aos);   // globals::s = ...
  std::vector<double> spins();                         // globals::s
  void run();                                          // one Heun (or RK4) step (core/jams++.cc:341)
  void run_steps(int n);                               // n steps without returning to the host in between
  void notify_monitors();                              // core/solver.cc:110-116
  void post_process_monitors();                        // core/jams++.cc:360-362
  std::vector<double> compute_fields();                // globals::h = sum_k field_k

  jb_ctx *ctx() { build(); return ctx_; }
  const Lattice &lattice() const { return lattice_; }
  void check(int status) const;                        // non-zero status -> std::runtime_error(jb_last_error)

 private:
  void build();

  const Lattice &lattice_;
  jb_ctx *ctx_ = nullptr;
  bool built_ = false;
  bool rk4_ = false;      // module "llg-rk4-b200-gpu" / "llg-rk4-gpu": CudaRK4BaseSolver::run (solvers/cuda_rk4_base.cu:50-108)
  int iteration_ = 0, max_steps_ = 0, min_steps_ = 0;
  uint64_t noise_step_offset_ = 0;   // added to the step index of the noise stream (resumed runs)
  double time_ = 0.0, step_size_ = 1.0;
  uint64_t seed_ = 0;
  std::vector<double> spins0_;
  std::unique_ptr<Physics> physics_;
  std::vector<std::unique_ptr<Hamiltonian>> hamiltonians_;
  std::vector<std::unique_ptr<Monitor>> monitors_;
};

// core/jams++.cc:231-377: build everything from a merged config and run the main loop
class Simulation {
 public:
  Simulation(const std::vector<std::string> &config_args, const std::string &name, const std::string &output_dir);
  void run();                         // while (solver->is_running()) { notify_monitors(); run(); } + post_process
  B200HeunLLGSolver &solver() { return *solver_; }
  const Lattice &lattice() const { return *lattice_; }
  const Setting &config() const { return *config_; }

 private:
  void run_initializer(const Setting &settings);   // initializer/init_bloch_domain_wall.cc:10-32

  std::unique_ptr<Setting> config_;
  std::unique_ptr<Lattice> lattice_;
  std::unique_ptr<B200HeunLLGSolver> solver_;
  std::string prefix_;
};

}  // namespace jams_b200

#endif  // JAMS_B200_HOST_JAMS_HOST_H
