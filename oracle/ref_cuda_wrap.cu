// oracle/ref_cuda_wrap.cu — TEST INFRASTRUCTURE ONLY (the GPU half of oracle/_ref).
//
// The reference's OWN CUDA code for the hot path, compiled by nvcc from the sources where they lie under /root/reference/src
// (oracle/Makefile, target `refcuda` -> oracle/_ref/libjams_ref_cuda.so; nothing of it is copied into this repository):
//   solvers/cuda_llg_heun_kernel.cuh        cuda_heun_llg_kernelA / B (+ the zero-safe pair)          rows a1 / a3
//   solvers/cuda_llg_rk4_kernel.cuh, solvers/cuda_rk4_base_kernel.cuh, cuda/cuda_spin_ops.cu         row f3
//   containers/sparse_matrix.h              SparseMatrix<double>::Builder + multiply_gpu (cuSPARSE)  rows a5 - a8
//   hamiltonian/cuda_uniaxial_anisotropy_kernel.cuh, hamiltonian/cuda_zeeman_kernel.cuh              rows a12 / a13
//   hamiltonian/cuda_biquadratic_exchange_kernel.cuh                                                 row f4
//   cuda/cuda_array_reduction.cu, cuda/cuda_spin_ops.cu, containers/mat3.h                           rows a19 / f4 (pinned boundaries)
//   cuda/cuda_array_kernels.cu, containers/multiarray.h + synced_memory.h                            rows a14 / a18
// What is written here is only the glue the reference keeps in classes that need its whole runtime (globals::, jams::instance(),
// libconfig): the order of calls in CUDAHeunLLGSolver::run (solvers/cuda_llg_heun.cu:66-127), CudaRK4BaseSolver::run
// (solvers/cuda_rk4_base.cu:50-108), CUDALLGRK4Solver::function_kernel / post_step (solvers/cuda_llg_rk4.cu:17-37),
// CudaSolver::compute_fields (cuda/cuda_solver.cc:11-26), the Hamiltonians' calculate_fields launch shapes and
// PinnedBoundariesPhysics::update (physics/pinned_boundaries.cc:34-46) — each cited where it is restated.
//
// Users: tests/test_gpu_reference_cuda.py (pins oracle/jams_oracle.cpp's rk4_run, biquadratic term and pin_region to the reference's
// kernels, and runs the replaced CUDA path beside the product) and bench.py's `reference_cuda` record.  The product never loads it.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusparse.h>
#include <curand.h>

#include "jams/cuda/cuda_common.h"
#include "jams/containers/vec3.h"
#include "jams/containers/mat3.h"
#include "jams/containers/multiarray.h"
#include "jams/containers/sparse_matrix.h"
#include "jams/containers/sparse_matrix_builder.h"

#include "jams/solvers/cuda_llg_heun_kernel.cuh"
#include "jams/solvers/cuda_llg_rk4_kernel.cuh"
#include "jams/solvers/cuda_rk4_base_kernel.cuh"
#include "jams/hamiltonian/cuda_uniaxial_anisotropy_kernel.cuh"
#include "jams/hamiltonian/cuda_zeeman_kernel.cuh"
#include "jams/hamiltonian/cuda_biquadratic_exchange_kernel.cuh"
#include "jams/cuda/cuda_spin_ops.cu"
#include "jams/cuda/cuda_array_reduction.cu"
#include "jams/cuda/cuda_array_kernels.cu"

// declared (not defined) by cuda/cuda_common.h for its CHECK_* macros
const char *cusparseGetStatusString(cusparseStatus_t status) { return cusparseGetErrorString(status); }
const char *cublasGetStatusString(cublasStatus_t status) { return status == CUBLAS_STATUS_SUCCESS ? "success" : "cublas error"; }
const char *curandGetStatusString(curandStatus_t status) { return status == CURAND_STATUS_SUCCESS ? "success" : "curand error"; }

namespace {

thread_local std::string g_error;

using Field = jams::MultiArray<double, 2>;
using Scalar = jams::MultiArray<double, 1>;

struct Term {
  enum Kind { EXCHANGE, UNIAXIAL, ZEEMAN } kind;
  Field field;                               // Hamiltonian::field_
  jams::SparseMatrix<double> matrix;         // SparseInteractionHamiltonian::interaction_matrix_
  int power = 0;                             // CudaUniaxialAnisotropyHamiltonian
  Scalar magnitude;
  Field axis;
  Field dc;                                  // CudaZeemanHamiltonian
  Field ac;
  Scalar omega;
  bool has_ac = false;
};

struct Sim {
  int n = 0;
  Field s, h, ds_dt, s_old, noise, sigma, k1, k2, k3, k4;
  Scalar mus, gyro, alpha;
  std::vector<std::unique_ptr<Term>> terms;
  cusparseHandle_t cusparse = nullptr;
  cublasHandle_t cublas = nullptr;
  curandGenerator_t curand = nullptr;
  double time = 0.0, dt = 0.0, temperature = 0.0;
  long iteration = 0;
  bool zero_safe = false;
  const double *host_normals = nullptr;      // one step's normals (N x 3) handed in by the test, or null -> curand
};

void fill(Field &a, const double *src) { std::memcpy(a.data(), src, sizeof(double) * a.elements()); }
void fill(Scalar &a, const double *src) { std::memcpy(a.data(), src, sizeof(double) * a.elements()); }

// CudaSolver::compute_fields (cuda/cuda_solver.cc:11-26) with the Hamiltonians' calculate_fields bodies
void compute_fields(Sim &sim) {
  if (sim.terms.empty()) {
    cudaMemset(sim.h.device_data(), 0, sizeof(double) * 3 * sim.n);
    return;
  }
  for (auto &tp : sim.terms) {
    Term &t = *tp;
    switch (t.kind) {
      case Term::EXCHANGE:    // hamiltonian/sparse_interaction.cc:36-45
        t.matrix.multiply_gpu(sim.s, t.field, sim.cusparse, nullptr);
        break;
      case Term::UNIAXIAL: {  // hamiltonian/cuda_uniaxial_anisotropy.cu:28-32, dev_blocksize_ = 64 (cuda_uniaxial_anisotropy.h:22)
        const unsigned bs = 64;
        cuda_uniaxial_field_kernel<<<(sim.n + bs - 1) / bs, bs>>>(sim.n, t.power, t.magnitude.device_data(), t.axis.device_data(),
                                                                  sim.s.device_data(), t.field.device_data());
        break;
      }
      case Term::ZEEMAN: {    // hamiltonian/cuda_zeeman.cu:18-41
        dim3 block_size; block_size.x = 32; block_size.y = 4;
        dim3 grid_size; grid_size.x = (sim.n + block_size.x - 1) / block_size.x; grid_size.y = (3 + block_size.y - 1) / block_size.y;
        cudaMemcpy(t.field.device_data(), t.dc.device_data(), sizeof(double) * 3 * sim.n, cudaMemcpyDeviceToDevice);
        if (t.has_ac)
          cuda_zeeman_ac_field_kernel<<<grid_size, block_size>>>(sim.n, sim.time, t.ac.device_data(), t.omega.device_data(),
                                                                sim.s.device_data(), t.field.device_data());
        break;
      }
    }
  }
  cudaMemcpy(sim.h.device_data(), sim.terms[0]->field.device_data(), sizeof(double) * 3 * sim.n, cudaMemcpyDeviceToDevice);
  const double one = 1.0;
  for (size_t i = 1; i < sim.terms.size(); ++i)
    CHECK_CUBLAS_STATUS(cublasDaxpy(sim.cublas, 3 * sim.n, &one, sim.terms[i]->field.device_data(), 1, sim.h.device_data(), 1));
}

// CudaThermostatClassical::update (thermostats/cuda_thermostat_classical.cc:47-56).  With normals handed in by the test the
// curand call is replaced by an upload of those normals; the scaling kernel is the reference's.
void update_thermostat(Sim &sim) {
  if (sim.temperature == 0) {
    CHECK_CUDA_STATUS(cudaMemset(sim.noise.device_data(), 0, sim.noise.elements() * sizeof(double)));
    return;
  }
  const int n3 = 3 * sim.n;
  if (sim.host_normals)
    CHECK_CUDA_STATUS(cudaMemcpy(sim.noise.device_data(), sim.host_normals, sizeof(double) * n3, cudaMemcpyHostToDevice))
  else
    CHECK_CURAND_STATUS(curandGenerateNormalDouble(sim.curand, sim.noise.device_data(), (n3 + (n3 % 2)), 0.0, 1.0));
  cuda_array_elementwise_scale(sim.n, 3, sim.sigma.device_data(), sqrt(sim.temperature), sim.noise.device_data(), 1,
                               sim.noise.device_data(), 1, nullptr);
}

// CUDAHeunLLGSolver::run (solvers/cuda_llg_heun.cu:66-127)
void heun_step(Sim &sim) {
  const double t0 = sim.time;
  const dim3 block_size = {84, 3, 1};
  auto grid_size = cuda_grid_size(block_size, {static_cast<unsigned int>(sim.n), 3, 1});
  cudaMemcpyAsync(sim.s_old.device_data(), sim.s.device_data(), sizeof(double) * 3 * sim.n, cudaMemcpyDeviceToDevice, nullptr);
  update_thermostat(sim);
  compute_fields(sim);
  if (sim.zero_safe)
    cuda_zero_safe_heun_llg_kernelA<<<grid_size, block_size>>>(sim.s.device_data(), sim.ds_dt.device_data(), sim.s_old.device_data(),
        sim.h.device_data(), sim.noise.device_data(), sim.gyro.device_data(), sim.mus.device_data(), sim.alpha.device_data(), sim.dt, sim.n);
  else
    cuda_heun_llg_kernelA<<<grid_size, block_size>>>(sim.s.device_data(), sim.ds_dt.device_data(), sim.s_old.device_data(),
        sim.h.device_data(), sim.noise.device_data(), sim.gyro.device_data(), sim.mus.device_data(), sim.alpha.device_data(), sim.dt, sim.n);
  sim.time = t0 + sim.dt;
  compute_fields(sim);
  if (sim.zero_safe)
    cuda_zero_safe_heun_llg_kernelB<<<grid_size, block_size>>>(sim.s.device_data(), sim.ds_dt.device_data(), sim.s_old.device_data(),
        sim.h.device_data(), sim.noise.device_data(), sim.gyro.device_data(), sim.mus.device_data(), sim.alpha.device_data(), sim.dt, sim.n);
  else
    cuda_heun_llg_kernelB<<<grid_size, block_size>>>(sim.s.device_data(), sim.ds_dt.device_data(), sim.s_old.device_data(),
        sim.h.device_data(), sim.noise.device_data(), sim.gyro.device_data(), sim.mus.device_data(), sim.alpha.device_data(), sim.dt, sim.n);
  sim.iteration++;
  sim.time = sim.iteration * sim.dt;
}

// CUDALLGRK4Solver::function_kernel (solvers/cuda_llg_rk4.cu:17-32)
void rk4_function(Sim &sim, Field &k) {
  compute_fields(sim);
  const dim3 block_size = {64, 1, 1};
  auto grid_size = cuda_grid_size(block_size, {static_cast<unsigned int>(sim.n), 1, 1});
  cuda_llg_rk4_kernel<<<grid_size, block_size>>>(sim.s.device_data(), k.device_data(), sim.h.device_data(), sim.noise.device_data(),
                                                 sim.gyro.device_data(), sim.mus.device_data(), sim.alpha.device_data(), sim.n);
}

// CudaRK4BaseSolver::run (solvers/cuda_rk4_base.cu:50-108); post_step = CUDALLGRK4Solver::post_step (cuda_llg_rk4.cu:35-37)
void rk4_step(Sim &sim) {
  const double t0 = sim.time;
  const int n3 = 3 * sim.n;
  cudaMemcpyAsync(sim.s_old.device_data(), sim.s.device_data(), sizeof(double) * n3, cudaMemcpyDeviceToDevice, nullptr);
  update_thermostat(sim);
  rk4_function(sim, sim.k1);
  double mid_time_step = 0.5 * sim.dt;
  sim.time = t0 + mid_time_step;
  CHECK_CUBLAS_STATUS(cublasDcopy(sim.cublas, n3, sim.s_old.device_data(), 1, sim.s.device_data(), 1));
  CHECK_CUBLAS_STATUS(cublasDaxpy(sim.cublas, n3, &mid_time_step, sim.k1.device_data(), 1, sim.s.device_data(), 1));
  rk4_function(sim, sim.k2);
  mid_time_step = 0.5 * sim.dt;
  sim.time = t0 + mid_time_step;
  CHECK_CUBLAS_STATUS(cublasDcopy(sim.cublas, n3, sim.s_old.device_data(), 1, sim.s.device_data(), 1));
  CHECK_CUBLAS_STATUS(cublasDaxpy(sim.cublas, n3, &mid_time_step, sim.k2.device_data(), 1, sim.s.device_data(), 1));
  rk4_function(sim, sim.k3);
  mid_time_step = sim.dt;
  sim.time = t0 + mid_time_step;
  CHECK_CUBLAS_STATUS(cublasDcopy(sim.cublas, n3, sim.s_old.device_data(), 1, sim.s.device_data(), 1));
  CHECK_CUBLAS_STATUS(cublasDaxpy(sim.cublas, n3, &mid_time_step, sim.k3.device_data(), 1, sim.s.device_data(), 1));
  rk4_function(sim, sim.k4);
  const dim3 block_size = {64, 1, 1};
  auto grid_size = cuda_grid_size(block_size, {static_cast<unsigned int>(n3), 1, 1});
  cuda_rk4_combination_kernel<<<grid_size, block_size>>>(sim.s.device_data(), sim.s_old.device_data(), sim.k1.device_data(),
      sim.k2.device_data(), sim.k3.device_data(), sim.k4.device_data(), sim.dt, n3);
  jams::normalise_spins_cuda(sim.s);
  cudaDeviceSynchronize();
  sim.iteration++;
  sim.time = sim.iteration * sim.dt;
}

template <class F>
int guarded(F &&f) {
  try {
    f();
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { g_error = cudaGetErrorString(e); return 1; }
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return 1;
  }
}

}  // namespace

extern "C" {

const char *jrc_last_error() { return g_error.c_str(); }

int jrc_device_count() {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

void *jrc_sim_create(int n, const double *mus, const double *gyro, const double *alpha) {
  Sim *sim = nullptr;
  int rc = guarded([&] {
    sim = new Sim;
    sim->n = n;
    for (Field *f : {&sim->s, &sim->h, &sim->ds_dt, &sim->s_old, &sim->noise, &sim->sigma, &sim->k1, &sim->k2, &sim->k3, &sim->k4}) {
      f->resize(n, 3);
      f->zero();
    }
    // curandGenerateNormalDouble writes an even count (cuda_thermostat_classical.cc:54); Thermostat allocates noise_ accordingly
    sim->noise.resize(n + 1, 3);
    sim->noise.zero();
    sim->mus.resize(n); sim->gyro.resize(n); sim->alpha.resize(n);
    fill(sim->mus, mus); fill(sim->gyro, gyro); fill(sim->alpha, alpha);
    if (cusparseCreate(&sim->cusparse) != CUSPARSE_STATUS_SUCCESS) throw std::runtime_error("cusparseCreate failed");
    if (cublasCreate(&sim->cublas) != CUBLAS_STATUS_SUCCESS) throw std::runtime_error("cublasCreate failed");
    // jams::instance(): CURAND_RNG_PSEUDO_DEFAULT generator (core/jams++ instance set-up)
    if (curandCreateGenerator(&sim->curand, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS) throw std::runtime_error("curandCreateGenerator failed");
    curandSetPseudoRandomGeneratorSeed(sim->curand, 12345ULL);
  });
  if (rc) { delete sim; return nullptr; }
  return sim;
}

void jrc_sim_destroy(void *hdl) {
  Sim *sim = static_cast<Sim *>(hdl);
  if (!sim) return;
  cudaDeviceSynchronize();
  if (sim->cusparse) cusparseDestroy(sim->cusparse);
  if (sim->cublas) cublasDestroy(sim->cublas);
  if (sim->curand) curandDestroyGenerator(sim->curand);
  delete sim;
}

// ExchangeHamiltonian ctor + SparseInteractionHamiltonian::insert_interaction_tensor / finalize
// (hamiltonian/exchange.cc:80-110, sparse_interaction.cc:20-34,102-130): 3x3 blocks, zero elements skipped, CSR through the Builder
int jrc_sim_add_exchange(void *hdl, long npairs, const int *i, const int *j, const double *J9, int check_symmetric) {
  Sim &sim = *static_cast<Sim *>(hdl);
  return guarded([&] {
    auto t = std::make_unique<Term>();
    t->kind = Term::EXCHANGE;
    t->field.resize(sim.n, 3);
    t->field.zero();
    jams::SparseMatrix<double>::Builder builder(3 * sim.n, 3 * sim.n);
    for (long q = 0; q < npairs; ++q)
      for (int m = 0; m < 3; ++m)
        for (int n = 0; n < 3; ++n) {
          const double value = J9[9 * q + 3 * m + n];
          if (value != 0.0) builder.insert(3 * i[q] + m, 3 * j[q] + n, value);
        }
    if (check_symmetric && !builder.is_symmetric()) throw std::runtime_error("sparse matrix for exchange is not symmetric");
    t->matrix = builder.set_format(jams::SparseMatrixFormat::CSR).build();
    sim.terms.push_back(std::move(t));
  });
}

int jrc_sim_add_uniaxial(void *hdl, int power, const double *magnitude, const double *axis) {
  Sim &sim = *static_cast<Sim *>(hdl);
  return guarded([&] {
    auto t = std::make_unique<Term>();
    t->kind = Term::UNIAXIAL;
    t->power = power;
    t->field.resize(sim.n, 3); t->field.zero();
    t->magnitude.resize(sim.n); fill(t->magnitude, magnitude);
    t->axis.resize(sim.n, 3); fill(t->axis, axis);
    sim.terms.push_back(std::move(t));
  });
}

int jrc_sim_add_zeeman(void *hdl, const double *dc, const double *ac, const double *omega) {
  Sim &sim = *static_cast<Sim *>(hdl);
  return guarded([&] {
    auto t = std::make_unique<Term>();
    t->kind = Term::ZEEMAN;
    t->field.resize(sim.n, 3); t->field.zero();
    t->dc.resize(sim.n, 3); fill(t->dc, dc);
    t->has_ac = ac != nullptr;
    if (t->has_ac) {
      t->ac.resize(sim.n, 3); fill(t->ac, ac);
      t->omega.resize(sim.n); fill(t->omega, omega);
    }
    sim.terms.push_back(std::move(t));
  });
}

long jrc_sim_exchange_nnz(void *hdl, int term) {
  Sim &sim = *static_cast<Sim *>(hdl);
  return static_cast<long>(sim.terms[term]->matrix.num_non_zero());
}

int jrc_sim_set_spins(void *hdl, const double *s) {
  Sim &sim = *static_cast<Sim *>(hdl);
  return guarded([&] {
    fill(sim.s, s);
    // CUDAHeunLLGSolver::initialize (solvers/cuda_llg_heun.cu:42-54): zero-safe kernels when any spin has zero length
    sim.zero_safe = false;
    for (int q = 0; q < sim.n; ++q)
      if (approximately_zero(Vec3{s[3 * q], s[3 * q + 1], s[3 * q + 2]}, DBL_EPSILON)) { sim.zero_safe = true; break; }
  });
}

int jrc_sim_get_spins(void *hdl, double *s) {
  Sim &sim = *static_cast<Sim *>(hdl);
  return guarded([&] { std::memcpy(s, static_cast<const Field &>(sim.s).data(), sizeof(double) * 3 * sim.n); });
}

int jrc_sim_get_h(void *hdl, double *h) {
  Sim &sim = *static_cast<Sim *>(hdl);
  return guarded([&] {
    compute_fields(sim);
    std::memcpy(h, static_cast<const Field &>(sim.h).data(), sizeof(double) * 3 * sim.n);
  });
}

// sigma_(i, j) of CudaThermostatClassical's constructor (cuda_thermostat_classical.cc:34-44) is computed by the caller
// (the oracle's sigma) and handed in as the N x 3 array the reference keeps
int jrc_sim_init_solver(void *hdl, double dt, const double *sigma_n3, double temperature, unsigned long long seed) {
  Sim &sim = *static_cast<Sim *>(hdl);
  return guarded([&] {
    sim.dt = dt;
    sim.time = 0.0;
    sim.iteration = 0;
    sim.temperature = temperature;
    if (sigma_n3) fill(sim.sigma, sigma_n3);
    curandSetPseudoRandomGeneratorSeed(sim.curand, seed);
  });
}

// normals: nsteps x N x 3 standard normals (host) or null (curand, as the reference draws them)
int jrc_sim_run_heun(void *hdl, int nsteps, const double *normals) {
  Sim &sim = *static_cast<Sim *>(hdl);
  return guarded([&] {
    for (int k = 0; k < nsteps; ++k) {
      sim.host_normals = normals ? normals + static_cast<size_t>(k) * 3 * sim.n : nullptr;
      heun_step(sim);
    }
    sim.host_normals = nullptr;
  });
}

int jrc_sim_run_rk4(void *hdl, int nsteps, const double *normals) {
  Sim &sim = *static_cast<Sim *>(hdl);
  return guarded([&] {
    for (int k = 0; k < nsteps; ++k) {
      sim.host_normals = normals ? normals + static_cast<size_t>(k) * 3 * sim.n : nullptr;
      rk4_step(sim);
    }
    sim.host_normals = nullptr;
  });
}

// `steps` steps timed with CUDA events after `warmup` untimed ones; returns milliseconds per step (< 0 on error).
// The timed region is CUDAHeunLLGSolver::run (CudaRK4BaseSolver::run for rk4 != 0) as the reference executes it: curand normals, the
// scaling kernel, two (four) cuSPARSE SpMVs (+ the other terms' kernels and the daxpy sum) and the solver's kernels.
static double time_steps(void *hdl, int steps, int warmup, int rk4) {
  Sim &sim = *static_cast<Sim *>(hdl);
  double ms_per_step = -1.0;
  int rc = guarded([&] {
    sim.host_normals = nullptr;
    for (int k = 0; k < warmup; ++k) rk4 ? rk4_step(sim) : heun_step(sim);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, nullptr);
    for (int k = 0; k < steps; ++k) rk4 ? rk4_step(sim) : heun_step(sim);
    cudaEventRecord(e1, nullptr);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    ms_per_step = ms / steps;
  });
  return rc ? -1.0 : ms_per_step;
}
double jrc_sim_time_heun(void *hdl, int steps, int warmup) { return time_steps(hdl, steps, warmup, 0); }
double jrc_sim_time_rk4(void *hdl, int steps, int warmup) { return time_steps(hdl, steps, warmup, 1); }

// CudaBiquadraticExchangeHamiltonian: scalar N x N matrix through the Builder (hamiltonian/cuda_biquadratic_exchange.cu:100-157)
// and calculate_fields (:159-171)
int jrc_biquadratic_field(int n, long npairs, const int *i, const int *j, const double *B, const double *s, double *h) {
  return guarded([&] {
    jams::SparseMatrix<double>::Builder builder(n, n);
    for (long q = 0; q < npairs; ++q) builder.insert(i[q], j[q], B[q]);
    jams::SparseMatrix<double> matrix = builder.set_format(jams::SparseMatrixFormat::CSR).build();
    Field spins(n, 3), field(n, 3);
    fill(spins, s);
    field.zero();
    const dim3 block_size = {128, 1, 1};
    auto grid_size = cuda_grid_size(block_size, {static_cast<unsigned int>(n), 1, 1});
    cuda_biquadratic_exchange_field_kernel<<<grid_size, block_size>>>(n, spins.device_data(), matrix.row_device_data(),
        matrix.col_device_data(), matrix.val_device_data(), field.device_data());
    cudaDeviceSynchronize();
    std::memcpy(h, static_cast<const Field &>(field).data(), sizeof(double) * 3 * n);
  });
}

// PinnedBoundariesPhysics::update, CUDA branch (physics/pinned_boundaries.cc:36-40)
int jrc_pin_region(int n, double *s, const double *mus, int count, const int *indices, const double *target, double *mag_out) {
  return guarded([&] {
    Field spins(n, 3);
    fill(spins, s);
    Scalar moments(n);
    fill(moments, mus);
    jams::MultiArray<int, 1> idx(indices, indices + count);
    Vec3 mag = jams::vector_field_indexed_scale_and_reduce_cuda(spins, moments, idx);
    auto rotation_matrix = rotation_matrix_between_vectors(mag, Vec3{target[0], target[1], target[2]});
    jams::rotate_spins_cuda(spins, rotation_matrix, idx);
    cudaDeviceSynchronize();
    std::memcpy(s, static_cast<const Field &>(spins).data(), sizeof(double) * 3 * n);
    if (mag_out) for (int c = 0; c < 3; ++c) mag_out[c] = mag[c];
  });
}

// The adapter's data path through the reference's real container (integration/jams/solvers/b200_llg_heun.cc): globals::s is a
// jams::MultiArray<double, 2>; the adapter imports from its CONST device_data() (host copy stays valid, synced_memory.h const_device_data),
// steps, and exports into its NON-const device_data(), which marks the host copy stale so that the monitors' next data() call
// downloads it (synced_memory.h mutable_device_data / const_host_data).  `import_fn` / `export_fn` are jb_import_spins / jb_export_spins
// of the product (handed over as pointers: this library does not link it), `between` runs the steps.
int jrc_multiarray_contract(int n, const double *s_in, double *s_out, double *host_copy_before_export,
                            int (*import_fn)(void *, const double *, int), int (*export_fn)(void *, double *, int), void *ctx,
                            void (*between)(void *)) {
  return guarded([&] {
    Field spins(n, 3);
    fill(spins, s_in);
    if (import_fn(ctx, static_cast<const Field &>(spins).device_data(), 1) != 0) throw std::runtime_error("import failed");
    between(ctx);
    // the host copy has not been touched by the import
    std::memcpy(host_copy_before_export, static_cast<const Field &>(spins).data(), sizeof(double) * 3 * n);
    if (export_fn(ctx, spins.device_data(), 1) != 0) throw std::runtime_error("export failed");
    cudaDeviceSynchronize();
    std::memcpy(s_out, static_cast<const Field &>(spins).data(), sizeof(double) * 3 * n);   // what a monitor reads: synced from the device
  });
}

// the reductions behind the magnetisation monitor on the CUDA path (cuda/cuda_array_reduction.cu): kind 0 = sum s, 1 = sum mus s,
// 2 = indexed sum s, 3 = indexed sum mus s
int jrc_reduce(int kind, int n, const double *s, const double *mus, int count, const int *indices, double *out3) {
  return guarded([&] {
    Field spins(n, 3);
    fill(spins, s);
    Scalar moments(n);
    if (mus) fill(moments, mus);
    Vec3 r{0, 0, 0};
    if (kind == 0) r = jams::vector_field_reduce_cuda(spins);
    else if (kind == 1) r = jams::vector_field_scale_and_reduce_cuda(spins, moments);
    else {
      jams::MultiArray<int, 1> idx(indices, indices + count);
      r = kind == 2 ? jams::vector_field_indexed_reduce_cuda(spins, idx) : jams::vector_field_indexed_scale_and_reduce_cuda(spins, moments, idx);
    }
    for (int c = 0; c < 3; ++c) out3[c] = r[c];
  });
}

}  // extern "C"
