// b200_llg_heun.h — JAMS-side adapter: registers libjams_b200.so as solver module "llg-heun-b200-gpu".
// Lives in the JAMS tree as src/jams/solvers/b200_llg_heun.{h,cc}; it is the only
// JAMS-aware code of the drop-in and contains no numerics: every call forwards to include/jams_b200.h.
#ifndef JAMS_SOLVER_B200_HEUNLLG_H
#define JAMS_SOLVER_B200_HEUNLLG_H

#if HAS_CUDA

#include <cstdint>
#include <string>
#include <vector>

#include "jams/cuda/cuda_solver.h"
#include "jams_b200.h"

class B200HeunLLGSolver : public CudaSolver {
 public:
  B200HeunLLGSolver() = default;
  ~B200HeunLLGSolver() override;

  inline explicit B200HeunLLGSolver(const libconfig::Setting &settings) { initialize(settings); }

  void initialize(const libconfig::Setting &settings) override;   // core/solver.h:20
  void run() override;                                            // core/solver.h:21, one Heun step
  void notify_monitors() override;                                // core/solver.h:63
  void compute_fields() override;                                 // core/solver.h:65, globals::h = sum_k field_k

  std::string name() const override { return "llg-heun-b200-gpu"; }

 private:
  void build();                 // lazy: Hamiltonians are registered after the solver is constructed (core/jams++.cc:274-288)
  void check(int status) const; // jb_status -> std::runtime_error, like CHECK_CUDA_STATUS (cuda/cuda_common.h:43-77)
  void import_spins();          // globals::s (device AoS) -> library SoA
  void export_spins();          // library SoA -> globals::s (device AoS); marks the host copy stale
  void setup_pinned_boundaries();   // physics/pinned_boundaries.cc:12-31 -> jb_set_region
  void apply_pinned_boundaries();   // physics/pinned_boundaries.cc:34-46 on the library's own state

  jb_ctx *ctx_ = nullptr;
  jb_lattice_desc desc_{};   // kept for the host-only helpers (jb_detect_exchange_template)
  bool built_ = false;
  bool spins_exported_ = true;  // globals::s currently equals the library's state
  bool gilbert_prefactor_ = false;
  bool rk4_ = false;            // registered as "llg-rk4-b200-gpu": jb_step_rk4 instead of jb_step
  struct PinnedRegion { int region; double magnetisation[3]; };
  std::vector<PinnedRegion> pinned_;   // physics.module = "pinned_boundaries": the edge regions, held by the library
  std::uint64_t seed_ = 0;
};

#endif  // HAS_CUDA
#endif  // JAMS_SOLVER_B200_HEUNLLG_H
