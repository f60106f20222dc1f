/* jams_b200.h — C ABI of the B200-native llg-heun + exchange hot path for stonerlab/jams.
 *
 * This is the drop-in boundary ("Face 2"): a JAMS `Solver` subclass registered as
 * module "llg-heun-b200-gpu"  calls these entry points instead of the
 * cuSPARSE/cuBLAS/cuRAND path of `CUDAHeunLLGSolver`.  Every entry point names the reference
 * interface it replaces (file:line relative to /root/reference/src/jams/).
 *
 * Conventions
 *  - extern "C", opaque handle, plain pointers and sizes, int status (0 = JB_OK); no exceptions cross.
 *  - units are JAMS internal units: time ps, field T, energy meV, moments meV/T (README.md:92-105).
 *  - host arrays are caller-owned and in the REFERENCE's layout: per-site arrays of length N in
 *    site order ((i*Ny + j)*Nz + k)*M + m (core/lattice.cc:622-657), vector fields N x 3 row-major
 *    (globals::s, core/lattice.cc:688-694).  The library owns all device memory and keeps spins
 *    in its own SoA/ghosted layout.
 *  - `on_device != 0` means the pointer is a device pointer on the context's GPU (what
 *    MultiArray::device_data() returns, containers/multiarray.h:239-253); otherwise host memory.
 *  - one context per (process, GPU); calls on one context must come from one host thread at a
 *    time; calls return after enqueueing work on the context's stream except those that hand
 *    data back to the host, which synchronise.
 *  - there is no CPU fallback: every compute entry point needs a CUDA device and fails with
 *    JB_ERR_CUDA otherwise.
 */
#ifndef JAMS_B200_H
#define JAMS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JB_ABI_VERSION 1

#if defined(__GNUC__)
#define JB_API __attribute__((visibility("default")))
#else
#define JB_API
#endif

typedef struct jb_ctx jb_ctx;

typedef enum jb_status {
  JB_OK = 0,
  JB_ERR_INVALID = 1,   /* bad argument / call order (jams::SanityException analogue) */
  JB_ERR_CUDA = 2,      /* CUDA runtime/driver error (CHECK_CUDA_STATUS, cuda/cuda_common.h:43-77) */
  JB_ERR_UNSUPPORTED = 3,
  JB_ERR_PEER = 4       /* halo peer signalling timed out / peer mapping failed */
} jb_status;

/* Hamiltonian selector for jb_fields / jb_energies (Hamiltonian::create names, core/hamiltonian.cc:80-115) */
typedef enum jb_term {
  JB_TERM_EXCHANGE = 0,  /* "exchange"      hamiltonian/exchange.cc */
  JB_TERM_UNIAXIAL = 1,  /* "uniaxial"      hamiltonian/uniaxial_anisotropy.cc */
  JB_TERM_ZEEMAN = 2,    /* "zeeman"        hamiltonian/zeeman.cc */
  JB_TERM_APPLIED = 3,   /* "applied-field" hamiltonian/applied_field.cc */
  JB_TERM_TOTAL = 4,     /* sum of the registered terms = globals::h after Solver::compute_fields */
  JB_TERM_BIQUADRATIC = 5, /* "biquadratic-exchange" hamiltonian/cuda_biquadratic_exchange.cu */
  JB_TERM_UNIAXIAL_2 = 6, /* a second and a third "uniaxial" Hamiltonian of the configuration (jb_set_uniaxial_term slots 1, 2) */
  JB_TERM_UNIAXIAL_3 = 7
} jb_term;

/* Lattice + slab description.  Replaces what the solver reads from globals::lattice
 * (Lattice::size, is_periodic, num_basis_sites; core/lattice.h:54-153).
 * The supercell has dims[0] x dims[1] x dims[2] unit cells of M motif sites.  This context owns
 * the x-slab [x_begin, x_begin + nx_local) of it; single-GPU: x_begin = 0,
 * nx_local = dims[0], n_ranks = 1.  N (local) = nx_local*dims[1]*dims[2]*M. */
typedef struct jb_lattice_desc {
  int32_t dims[3];
  int32_t num_motif;      /* M */
  int32_t periodic[3];    /* lattice.periodic (core/lattice.cc:413) */
  int32_t x_begin;
  int32_t nx_local;
  int32_t rank;           /* position in the ring of slabs */
  int32_t n_ranks;
  int32_t device;         /* CUDA device ordinal, -1 = current device */
} jb_lattice_desc;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Solver construction (core/jams++.cc:274; CUDAHeunLLGSolver::initialize, solvers/cuda_llg_heun.cu:21-62) */
JB_API int jb_create(jb_ctx **ctx, const jb_lattice_desc *desc);
JB_API void jb_destroy(jb_ctx *ctx);
/* text of the last error on this context (or of a failed jb_create when ctx == NULL);
 * the adapter rethrows it as std::runtime_error like cuda/cuda_common.h:43-77 */
JB_API const char *jb_last_error(const jb_ctx *ctx);
JB_API int jb_abi_version(void);

/* ---- parameters ---------------------------------------------------------------------------- */
/* globals::mus / gyro / alpha, N each (core/lattice.cc:703-713).  gyro already carries the
 * Gilbert prefactor if the caller applied it (jams::gilbert_gyro_prefactor, core/lattice.cc:95-97). */
JB_API int jb_set_materials(jb_ctx *ctx, const double *mus, const double *gyro, const double *alpha);

/* Exchange, translation-invariant form: the processed interaction template of
 * post_process_interactions (core/interactions.cc:292-347): entry n couples motif site motif_i[n]
 * of cell (i,j,k) to motif site motif_j[n] of cell (i,j,k)+T3[3n..3n+2] with tensor J9[9n..] (row-major,
 * meV, already multiplied by interaction_prefactor and the unit conversion, hamiltonian/exchange.cc:165).
 * Boundary handling is Lattice::apply_boundary_conditions (core/lattice.cc:987-1007).
 * Equivalent to the CSR that ExchangeHamiltonian builds (hamiltonian/exchange.cc:162-171) on a lattice
 * without impurities. */
JB_API int jb_set_exchange_template(jb_ctx *ctx, int32_t n, const int32_t *motif_i, const int32_t *motif_j,
                             const int32_t *T3, const double *J9);

/* Biquadratic exchange (CudaBiquadraticExchangeHamiltonian, hamiltonian/cuda_biquadratic_exchange.cu:9-156): field
 * h_i = sum_j 2 B_ij s_j (s_i . s_j) (cuda_biquadratic_exchange_kernel.cuh:5-30), per-spin energy -1/2 s_i . h_i (:235-240), total
 * energy 1/2 sum_i -s_i . (h_i / 2) (:185-201).  Translation-invariant form like jb_set_exchange_template with one scalar B (meV,
 * already converted; the reference keeps only values above energy_cutoff, :127-134 -- the caller filters) per entry.  n = 0 removes
 * the term.  A step with this term runs on the direct-gather stage kernel. */
JB_API int jb_set_biquadratic_template(jb_ctx *ctx, int32_t n, const int32_t *motif_i, const int32_t *motif_j,
                                       const int32_t *T3, const double *B);

/* Exchange, general form: the neighbour list itself, as
 * ExchangeHamiltonian::neighbour_list() exposes it (hamiltonian/exchange.h:13,
 * containers/interaction_list.h:45-47): pairs (i,j) in GLOBAL site ids with an index into the table
 * of unique tensors (row-major 3x3, meV, already scaled).  Only pairs whose i lies in this
 * context's slab are used.  A translation-invariant list (jb_detect_exchange_template) is turned into the
 * template form and runs on the TMA kernels; anything else (impurities, vacancies) is kept as an ELL table (explicit int32
 * indices).  On a slab-decomposed lattice every rank passes the SAME global list: neighbours across a slab face are addressed
 * through the x ghost planes, whose depth is the largest x distance of any pair of the list. */
JB_API int jb_set_exchange_pairs(jb_ctx *ctx, int64_t n_pairs, const int32_t *i, const int32_t *j,
                          const int32_t *value_id, int32_t n_values, const double *J9);

/* Host-only helper (needs no GPU): recognise a translation-invariant neighbour list.  Pairs are in GLOBAL site
 * ids as for jb_set_exchange_pairs; only pairs whose i lies in the slab of `desc` are looked at.  On success
 * *n_template is the number of template entries written to motif_i / motif_j / T3 (3 per entry) / J9_out (9 per
 * entry), in the form jb_set_exchange_template takes; *n_template = -1 means the list is not translation
 * invariant (impurities or vacancies, a periodic dimension shorter than 2*range+1, or more than `capacity`
 * distinct entries) and the general ELL path has to be used.  jb_set_exchange_pairs calls this itself (option
 * "detect_template", default 1), so an adapter can hand over ExchangeHamiltonian::neighbour_list() as it is
 * and still get the template kernel. */
JB_API int jb_detect_exchange_template(const jb_lattice_desc *desc, int64_t n_pairs, const int32_t *i, const int32_t *j,
                                const int32_t *value_id, int32_t n_values, const double *J9, int32_t capacity,
                                int32_t *n_template, int32_t *motif_i, int32_t *motif_j, int32_t *T3, double *J9_out);

/* UniaxialAnisotropyHamiltonian: power_ (2,4,6), magnitude_ (N, meV), axis_ (N x 3, unit vectors)
 * (hamiltonian/uniaxial_anisotropy.h:30-32, .cc:89-114). */
JB_API int jb_set_uniaxial(jb_ctx *ctx, int32_t power, const double *magnitude, const double *axis);
/* The reference sums any number of Hamiltonians (core/solver.cc:43-57, cuda/cuda_solver.cc:11-26), and K1 + K2 anisotropies are
 * written as two "uniaxial" modules (one power_ each, uniaxial_anisotropy.cc:89-114).  slot 0 = jb_set_uniaxial (term
 * JB_TERM_UNIAXIAL); slots 1 and 2 hold a second and a third module (JB_TERM_UNIAXIAL_2 / _3).  A context with a slot > 0 in use
 * runs its steps on the direct-gather stage kernels (the TMA kernels keep one uniaxial term in their parameter block).
 * power = 0 or magnitude = NULL clears the slot. */
JB_API int jb_set_uniaxial_term(jb_ctx *ctx, int32_t slot, int32_t power, const double *magnitude, const double *axis);

/* ZeemanHamiltonian: dc_local_field_ (N x 3, already multiplied by mu_i, meV), optional
 * ac_local_field_ (N x 3, meV) and ac_local_frequency_ (N, rad/ps = 2*pi*f) or NULL
 * (hamiltonian/zeeman.cc:26-71). */
JB_API int jb_set_zeeman(jb_ctx *ctx, const double *dc_local_field, const double *ac_local_field,
                  const double *ac_local_frequency);

/* AppliedFieldHamiltonian: homogeneous B(t) in Tesla; the field on site i is mu_i * B
 * (hamiltonian/applied_field.cc:146-148).  Call again whenever B(t) changes. enable = 0 removes it. */
JB_API int jb_set_applied_field(jb_ctx *ctx, const double B[3], int32_t enable);
/* The same Hamiltonian with a time-dependent amplitude, B(t) = B g(t) (TimeDependentField subclasses,
 * hamiltonian/applied_field.cc:10-82): JB_FIELD_STATIC g = 1; JB_FIELD_SINC g = sinc(pi f_bw (t - t0));
 * JB_FIELD_SINC_COS g = sinc(pi f_bw (t - t0)) cos(2 pi f_c (t - t0)); t0 in ps, f_bw and f_c in THz (the reference converts
 * its config values the same way, :37-38,66-68).  The stage kernels see g at the stage's time (predictor t, corrector t + dt,
 * cpu_llg_heun.cc:46,103-104), jb_fields / jb_energies at the time they are given. */
#define JB_FIELD_STATIC 0
#define JB_FIELD_SINC 1
#define JB_FIELD_SINC_COS 2
JB_API int jb_set_applied_field_pulse(jb_ctx *ctx, const double B[3], int32_t type, double time_center_ps,
                                      double freq_bandwidth_THz, double freq_center_THz);

/* ---- state --------------------------------------------------------------------------------- */
/* globals::s <-> device SoA.  s_aos is N(local) x 3 row-major (core/lattice.cc:688). */
JB_API int jb_import_spins(jb_ctx *ctx, const double *s_aos, int32_t on_device);
JB_API int jb_export_spins(jb_ctx *ctx, double *s_aos, int32_t on_device);

/* ---- the hot path --------------------------------------------------------------------------- */
/* nsteps calls of Solver::run (HeunLLGSolver::run, solvers/cpu_llg_heun.cc:45-148 /
 * CUDAHeunLLGSolver::run, solvers/cuda_llg_heun.cu:64-122) fused into two kernels per step:
 *   dt_ps           step_size_ (ps)
 *   time_ps         solver time at the start of the first step (only AC Zeeman depends on it)
 *   temperature_K   physics_module_->temperature() (core/solver.cc:94-97); 0 disables the thermostat
 *   seed, first_step_index   key / counter of the Philox4x32-10 Langevin noise: the N(0,1) draw for
 *                   (global site, component) at step index first_step_index + n depends on nothing
 *                   else, so slab decompositions give identical trajectories
 *   gilbert_prefactor  only enters sigma_i = sqrt(2 kB alpha_i / (mu_i gyro_i dt [1+alpha_i^2]))
 *                   (solvers/cpu_llg_heun.cc:35-42, thermostats/cuda_thermostat_classical.cc:34-43) */
JB_API int jb_step(jb_ctx *ctx, int32_t nsteps, double dt_ps, double time_ps, double temperature_K,
            uint64_t seed, uint64_t first_step_index, int32_t gilbert_prefactor);

/* nsteps calls of CudaRK4BaseSolver::run with CUDALLGRK4Solver::function_kernel (module "llg-rk4-gpu":
 * solvers/cuda_rk4_base.cu:50-108, solvers/cuda_llg_rk4.cu:17-34, cuda_llg_rk4_kernel.cuh:11-58,
 * cuda_rk4_base_kernel.cuh:1-19, cuda/cuda_spin_ops.cu:4-17): classical RK4 on the LLG right hand side with one
 * white-noise draw per step, unnormalised intermediate states and a normalisation after the combination.  Four
 * launches per step, each fusing the field evaluation, k_i, the next stage input and the running sum of the k's.
 * Arguments as jb_step; slab-decomposed runs exchange halos once per stage like jb_step.  Needs a translation-invariant
 * exchange template. */
JB_API int jb_step_rk4(jb_ctx *ctx, int32_t nsteps, double dt_ps, double time_ps, double temperature_K,
                uint64_t seed, uint64_t first_step_index, int32_t gilbert_prefactor);

/* Thermostat::device_data() equivalent (core/thermostat.h:22-34): the white-noise field
 * xi_ij = sigma_i sqrt(T) n_ij in Tesla that jb_step uses at step `step_index`, N x 3.
 * With normals_only != 0 the raw N(0,1) draws n_ij are returned instead. */
JB_API int jb_noise(jb_ctx *ctx, double dt_ps, double temperature_K, uint64_t seed, uint64_t step_index,
             int32_t gilbert_prefactor, int32_t normals_only, double *xi_aos, int32_t on_device);

/* ---- Hamiltonian / Monitor surface ------------------------------------------------------------ */
/* Hamiltonian::calculate_fields(time) (core/hamiltonian.h:27-42): field of one term (meV, NOT divided
 * by mu) or JB_TERM_TOTAL (= globals::h after Solver::compute_fields, core/solver.cc:43-57), N x 3. */
JB_API int jb_fields(jb_ctx *ctx, int32_t term, double time_ps, double *h_aos, int32_t on_device);

/* Hamiltonian::calculate_energies / calculate_total_energy (core/hamiltonian.h:27-42):
 * per-spin energies e (N, may be NULL; exchange: -s_i.(A s)_i, sparse_interaction.cc:79-84) and the
 * total (exchange carries the factor 1/2, sparse_interaction.cc:86-100; this rank's slab only). */
JB_API int jb_energies(jb_ctx *ctx, int32_t term, double time_ps, double *e, int32_t on_device, double *total);

/* MagnetisationMonitor::update reduction (monitors/magnetisation.cc:88-100, helpers/spinops.cc:55-67):
 * M4[4g..4g+3] = { sum_i mu_i s_i (x,y,z), sum_i mu_i } over the spins of group g in this slab.
 * group_of_spin (N, values in [0,n_groups)) or NULL for a single group. */
JB_API int jb_magnetisation(jb_ctx *ctx, int32_t n_groups, const int32_t *group_of_spin, double *M4);
/* The monitor builds its groups once, in its constructor (monitors/magnetisation.cc:21-60): register them here and call
 * jb_magnetisation with group_of_spin = NULL and the same n_groups afterwards -- no N-long array crosses PCIe per update. */
JB_API int jb_set_magnetisation_groups(jb_ctx *ctx, int32_t n_groups, const int32_t *group_of_spin);

/* ---- physics hooks that rewrite spins on the device ------------------------------------------ */
/* PinnedBoundariesPhysics (physics/pinned_boundaries.cc:12-46) keeps the magnetisation direction of an edge region by
 * rotating its spins every iteration.  A region is a list of local site ids (reference order, what
 * PinnedBoundary::indices holds); up to JB_MAX_REGIONS of them.
 *   jb_region_moment  M4 = {sum mu_i s_i (x,y,z), sum mu_i} over the region
 *                     (jams::vector_field_indexed_scale_and_reduce_cuda / jams::sum_spins_moments, helpers/spinops.cc:55-67)
 *   jb_rotate_region  s_i <- R s_i, R row-major 3x3 (jams::rotate_spins_cuda, cuda/cuda_spin_ops.cu:29-60), ghost images
 *                     refreshed.  The adapter computes R = rotation_matrix_between_vectors(M, pinned_magnetisation)
 *                     (containers/mat3.h:334-366) between the two calls; with several slabs it all-reduces M first. */
#define JB_MAX_REGIONS 8
JB_API int jb_set_region(jb_ctx *ctx, int32_t region, int32_t n_sites, const int32_t *site_index);
JB_API int jb_region_moment(jb_ctx *ctx, int32_t region, double *M4);
JB_API int jb_rotate_region(jb_ctx *ctx, int32_t region, const double *R9);

/* ---- multi-GPU halo plumbing (no reference counterpart) ------------------------- */
/* Each rank exports one opaque handle blob (JB_HALO_HANDLE_BYTES) describing its device buffers;
 * the host layer all-gathers the blobs (torch.distributed / MPI / files) and hands every rank the
 * blobs of its ring neighbours.  After jb_halo_connect the stage kernels store boundary planes
 * straight into the neighbours' ghost planes over NVLink (P2P stores) and signal with flags. */
#define JB_HALO_HANDLE_BYTES 256
JB_API int jb_halo_export_handle(jb_ctx *ctx, void *blob);
JB_API int jb_halo_connect(jb_ctx *ctx, const void *blob_lo_neighbour, const void *blob_hi_neighbour);

/* ---- introspection for benches and tests -------------------------------------------------------- */
/* number of kernels this context has launched so far (bench.py "gpu_launches") */
JB_API int64_t jb_kernel_launches(const jb_ctx *ctx);
/* device time in ms of the stage kernels since the last call (option "time_kernels" > 0), measured with CUDA events on the
 * context's stream; out2[0] = stage A total, out2[1] = stage B total.  Synchronises. */
JB_API int jb_last_step_kernel_ms(jb_ctx *ctx, double *out2);
/* how many steps contributed to those totals: with option "time_kernels" = N the launches of every N-th step of a jb_step call are
 * bracketed by events (N = 1: all of them; an event record between two launches costs about 2 us of stream time, so a bench keeps
 * N > 1 inside its timed region).  Reset by jb_last_step_kernel_ms. */
JB_API int64_t jb_timed_steps(const jb_ctx *ctx);
/* which stage kernel the most recent jb_step ran: JB_KERNEL_DIRECT (one thread per spin, gathers through L1 / L2), JB_KERNEL_PAIR
 * (TMA plane ring, a z pair of sites per thread: jb_stage_pair.cu), JB_KERNEL_ROWS (TMA plane ring, four y rows per thread with
 * register reuse, deep isotropic templates: jb_stage_rows.cu), JB_KERNEL_ELL (general neighbour list); -1 before the first step */
#define JB_KERNEL_DIRECT 0
#define JB_KERNEL_PAIR 2
#define JB_KERNEL_ROWS 4
#define JB_KERNEL_ELL 5
JB_API int jb_stage_kernel(const jb_ctx *ctx);
/* block until all enqueued work of this context has finished */
JB_API int jb_synchronize(jb_ctx *ctx);
/* the cudaStream_t the context launches on (so callers can record events on it) */
JB_API void *jb_stream(jb_ctx *ctx);
/* tuning knobs (tile shape etc.); unknown keys return JB_ERR_INVALID */
JB_API int jb_set_option(jb_ctx *ctx, const char *key, int64_t value);
/* Host-only helper (needs no GPU): the work-item plan the stage kernel's queue uses for a slab of nx_local planes with x ghost
 * depth ghost_x, n_columns yz-column tiles and n_ctas resident CTAs: chunks (x0[k], xc[k]) in queue order (face chunks first,
 * long chunks, a taper of short ones); capacity >= 160.  Exposed so that tests can check coverage and ordering on the CPU. */
JB_API int jb_plan_work_items(int32_t nx_local, int32_t ghost_x, int32_t n_columns, int32_t n_ctas, int32_t capacity,
                              int32_t *n_chunks, int32_t *x0, int32_t *xc);
/* with option "trace" = 1: per resident CTA of the most recent stage-kernel launch 32 x uint64: {SM id, first / last device
 * clock in ns, work items taken, then per item (item id << 40 | start in ns after the CTA's first clock)}, at most
 * `capacity` CTAs -- the load-balance evidence in profiles/.  Synchronises. */
#define JB_TRACE_WORDS_PER_CTA 32
JB_API int jb_last_stage_trace(jb_ctx *ctx, uint64_t *out, int32_t capacity, int32_t *n_ctas);

#ifdef __cplusplus
}
#endif
#endif /* JAMS_B200_H */
