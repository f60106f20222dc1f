// oracle/ref_wrap.cpp — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or
// called from the product path (jams_b200/).  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library built from it.
//
// What this is: a thin extern "C" shell around the REFERENCE'S OWN header-only code,
// compiled from the sources where they lie under /root/reference/src (nothing is copied
// into this repository).  The build recipe is oracle/Makefile (target `ref`), output goes
// to oracle/_ref/libjams_ref.so.  It exists to (1) validate the self-contained restatement
// in oracle/jams_oracle.cpp bit-for-bit and (2) serve as the CPU baseline in bench.py.
//
// Reference code that is executed *as is* through this file:
//   jams::SparseMatrix<double>::Builder::{insert,sort,merge,build_csr,is_symmetric}
//                                         (containers/sparse_matrix_builder.h:98-248,320-362)
//   jams::SparseMatrix<double>::multiply -> jams::Xcsrmv_general
//                                         (containers/sparse_matrix.h:278-326, interface/sparse_blas.h:13-81)
//   jams::InteractionList<Mat3,2>         (containers/interaction_list.h:15-92)
//   Vec3 / Mat3 operators, cross, unit_vector, norm, rotation_matrix_between_vectors
//                                         (containers/vec3.h, containers/mat3.h:334-366)
//   approximately_equal / approximately_zero (helpers/maths.h:16-28)
//   jams::MultiArray                      (containers/multiarray.h)
//
// What cannot be compiled here and is therefore restated (loop for loop, with the
// reference's own types and operators) because the .cc files need libconfig++/spglib/HDF5
// and the `globals` object graph: HeunLLGSolver::run (solvers/cpu_llg_heun.cc:45-148),
// Solver::compute_fields (core/solver.cc:43-57), the per-spin loops of the uniaxial and
// Zeeman Hamiltonians (hamiltonian/uniaxial_anisotropy.cc:155-172, zeeman.cc:121-132) and
// SparseInteractionHamiltonian::insert_interaction_tensor (sparse_interaction.cc:25-34).

#include <cstdint>
#include <cstring>
#include <cmath>
#include <random>
#include <string>
#include <vector>
#include <memory>

#include "jams/helpers/utils.h"
#include "jams/helpers/maths.h"
#include "jams/helpers/consts.h"
#include "jams/containers/vec3.h"
#include "jams/containers/mat3.h"
#include "jams/containers/multiarray.h"
#include "jams/containers/sparse_matrix.h"
#include "jams/containers/sparse_matrix_builder.h"
#include "jams/containers/interaction_list.h"
#include "jams/interface/openmp.h"
#if HAS_OMP
#include <omp.h>
#endif

#if JREF_HAVE_PCG
#include "arrow/vendored/pcg/pcg_random.hpp"
#endif

// ---------------------------------------------------------------------------------------
// Symbols the reference headers expect from BLAS/LAPACK (see oracle/ref_shim/*.h).
// ---------------------------------------------------------------------------------------
extern "C" void cblas_daxpy(const int n, const double alpha, const double *x, const int incx,
                            double *y, const int incy) {
  for (int i = 0; i < n; ++i) y[i * incy] += alpha * x[i * incx];
}

namespace {
template <typename T>
void getrf_generic(int n, T *a, int lda, int *ipiv, int *info) {
  *info = 0;
  for (int k = 0; k < n; ++k) {
    int p = k;
    T best = std::abs(a[k + k * lda]);
    for (int i = k + 1; i < n; ++i) {
      if (std::abs(a[i + k * lda]) > best) { best = std::abs(a[i + k * lda]); p = i; }
    }
    ipiv[k] = p + 1;
    if (best == T(0)) { *info = k + 1; continue; }
    if (p != k) for (int j = 0; j < n; ++j) std::swap(a[k + j * lda], a[p + j * lda]);
    for (int i = k + 1; i < n; ++i) {
      a[i + k * lda] /= a[k + k * lda];
      for (int j = k + 1; j < n; ++j) a[i + j * lda] -= a[i + k * lda] * a[k + j * lda];
    }
  }
}

template <typename T>
void getri_generic(int n, T *a, int lda, const int *ipiv, int *info) {
  // inverse from the packed LU: solve A X = I column by column
  *info = 0;
  std::vector<T> inv(n * n, T(0));
  for (int c = 0; c < n; ++c) {
    std::vector<T> b(n, T(0));
    b[c] = T(1);
    for (int k = 0; k < n; ++k) std::swap(b[k], b[ipiv[k] - 1]);
    for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) b[i] -= a[i + j * lda] * b[j];
    for (int i = n - 1; i >= 0; --i) {
      for (int j = i + 1; j < n; ++j) b[i] -= a[i + j * lda] * b[j];
      b[i] /= a[i + i * lda];
    }
    for (int i = 0; i < n; ++i) inv[i + c * n] = b[i];
  }
  for (int c = 0; c < n; ++c) for (int i = 0; i < n; ++i) a[i + c * lda] = inv[i + c * n];
}
}  // namespace

extern "C" void dgetrf_(int *m, int *n, double *a, int *lda, int *ipiv, int *info) { (void)m; getrf_generic(*n, a, *lda, ipiv, info); }
extern "C" void sgetrf_(int *m, int *n, float *a, int *lda, int *ipiv, int *info) { (void)m; getrf_generic(*n, a, *lda, ipiv, info); }
extern "C" void dgetri_(int *n, double *a, int *lda, int *ipiv, double *, int *, int *info) { getri_generic(*n, a, *lda, ipiv, info); }
extern "C" void sgetri_(int *n, float *a, int *lda, int *ipiv, float *, int *, int *info) { getri_generic(*n, a, *lda, ipiv, info); }

// ---------------------------------------------------------------------------------------
// The reference-backed simulation object
// ---------------------------------------------------------------------------------------
namespace {

thread_local std::string g_error;

struct RefSim {
  int num_spins = 0;
  int num_spins3 = 0;

  // globals::* (core/globals.h:23-40)
  jams::MultiArray<double, 2> s, h, ds_dt;
  jams::MultiArray<double, 1> mus, gyro, alpha;

  // HeunLLGSolver members (solvers/cpu_llg_heun.h:31-33)
  jams::MultiArray<double, 2> s_old_, w_;
  jams::MultiArray<double, 1> sigma_;
  double step_size_ = 0.0;
  double time_ = 0.0;
  int iteration_ = 0;
  double temperature_ = 0.0;

  // Hamiltonians in registration order: each owns field_ (core/hamiltonian.h)
  enum Kind { EXCHANGE, UNIAXIAL, ZEEMAN };
  struct Term {
    Kind kind;
    jams::MultiArray<double, 2> field_;
    // exchange
    jams::SparseMatrix<double> interaction_matrix_;
    // uniaxial (uniaxial_anisotropy.h:30-32)
    int power_ = 2;
    jams::MultiArray<double, 1> magnitude_;
    jams::MultiArray<double, 2> axis_;
    // zeeman (zeeman.h:28-32)
    jams::MultiArray<double, 2> dc_local_field_, ac_local_field_;
    jams::MultiArray<double, 1> ac_local_frequency_;
    bool has_ac_local_field_ = false;
  };
  std::vector<std::unique_ptr<Term>> hamiltonians_;

  std::mt19937_64 fallback_rng_{12345};
#if JREF_HAVE_PCG
  arrow_vendored::pcg32_k1024 random_generator_{42u};
#endif
};

void calculate_fields(RefSim &sim, RefSim::Term &t, double time) {
  switch (t.kind) {
    case RefSim::EXCHANGE:
      // SparseInteractionHamiltonian::calculate_fields (sparse_interaction.cc:36-45)
      t.interaction_matrix_.multiply(sim.s, t.field_);
      break;
    case RefSim::UNIAXIAL:
      // UniaxialAnisotropyHamiltonian::calculate_fields / calculate_field (uniaxial_anisotropy.cc:155-172)
      for (auto i = 0; i < sim.num_spins; ++i) {
        auto dot = (t.axis_(i, 0) * sim.s(i, 0) + t.axis_(i, 1) * sim.s(i, 1) + t.axis_(i, 2) * sim.s(i, 2));
        for (auto j = 0; j < 3; ++j) {
          t.field_(i, j) = t.magnitude_(i) * t.power_ * pow(dot, t.power_ - 1) * t.axis_(i, j);
        }
      }
      break;
    case RefSim::ZEEMAN:
      // ZeemanHamiltonian::calculate_fields (zeeman.cc:121-132)
      for (int i = 0; i < sim.num_spins; ++i) {
        for (int j = 0; j < 3; ++j) {
          t.field_(i, j) = t.dc_local_field_(i, j);
        }
        if (t.has_ac_local_field_) {
          for (int j = 0; j < 3; ++j) {
            t.field_(i, j) += t.ac_local_field_(i, j) * cos(t.ac_local_frequency_(i) * time);
          }
        }
      }
      break;
  }
}

// Solver::compute_fields (core/solver.cc:43-57)
void compute_fields(RefSim &sim) {
  if (sim.hamiltonians_.empty()) return;
  for (auto &hh : sim.hamiltonians_) {
    calculate_fields(sim, *hh, sim.time_);
  }
  std::copy(sim.hamiltonians_[0]->field_.data(), sim.hamiltonians_[0]->field_.data() + sim.num_spins3, sim.h.data());
  if (sim.hamiltonians_.size() == 1) return;
  for (std::size_t i = 1; i < sim.hamiltonians_.size(); ++i) {
    cblas_daxpy(sim.num_spins3, 1.0, sim.hamiltonians_[i]->field_.data(), 1, sim.h.data(), 1);
  }
}

// HeunLLGSolver::run (solvers/cpu_llg_heun.cc:45-148).  `normals` (3N standard normals) replaces
// the reference's pcg32_k1024 + std::normal_distribution stream when given, so that a GPU run
// and this run can consume identical noise; with normals == nullptr the generator is used
// exactly as the reference does (serial std::generate).
void heun_run(RefSim &sim, const double *normals) {
  double t0 = sim.time_;
  const int num_spins = sim.num_spins;
  std::normal_distribution<> normal_distribution;

  sim.s_old_ = sim.s;

  if (sim.temperature_ > 0.0) {
    if (normals) {
      std::copy(normals, normals + sim.num_spins3, sim.w_.begin());
    } else {
#if JREF_HAVE_PCG
      std::generate(sim.w_.begin(), sim.w_.end(), [&]() { return normal_distribution(sim.random_generator_); });
#else
      std::generate(sim.w_.begin(), sim.w_.end(), [&]() { return normal_distribution(sim.fallback_rng_); });
#endif
    }
    const auto sqrt_temperature = sqrt(sim.temperature_);
    OMP_PARALLEL_FOR
    for (auto i = 0; i < num_spins; ++i) {
      for (auto j = 0; j < 3; ++j) {
        sim.w_(i, j) = sim.w_(i, j) * sim.sigma_(i) * sqrt_temperature;
      }
    }
  }

  compute_fields(sim);

  if (sim.temperature_ > 0.0) {
    OMP_PARALLEL_FOR
    for (auto i = 0; i < num_spins; ++i) {
      for (auto j = 0; j < 3; ++j) {
        sim.h(i, j) = (sim.w_(i, j) + sim.h(i, j) / sim.mus(i));
      }
    }
  } else {
    OMP_PARALLEL_FOR
    for (auto i = 0; i < num_spins; ++i) {
      for (auto j = 0; j < 3; ++j) {
        sim.h(i, j) = sim.h(i, j) / sim.mus(i);
      }
    }
  }

  OMP_PARALLEL_FOR
  for (auto i = 0; i < num_spins; ++i) {
    Vec3 spin = {sim.s(i, 0), sim.s(i, 1), sim.s(i, 2)};
    Vec3 field = {sim.h(i, 0), sim.h(i, 1), sim.h(i, 2)};
    Vec3 rhs = -sim.gyro(i) * (cross(spin, field) + sim.alpha(i) * cross(spin, (cross(spin, field))));
    for (auto j = 0; j < 3; ++j) {
      sim.ds_dt(i, j) = 0.5 * rhs[j];
    }
    spin = unit_vector(spin + sim.step_size_ * rhs);
    for (auto j = 0; j < 3; ++j) {
      sim.s(i, j) = spin[j];
    }
  }

  double mid_time_step = sim.step_size_;
  sim.time_ = t0 + mid_time_step;

  compute_fields(sim);

  if (sim.temperature_ > 0.0) {
    OMP_PARALLEL_FOR
    for (auto i = 0; i < num_spins; ++i) {
      for (auto j = 0; j < 3; ++j) {
        sim.h(i, j) = (sim.w_(i, j) + sim.h(i, j) / sim.mus(i));
      }
    }
  } else {
    OMP_PARALLEL_FOR
    for (auto i = 0; i < num_spins; ++i) {
      for (auto j = 0; j < 3; ++j) {
        sim.h(i, j) = sim.h(i, j) / sim.mus(i);
      }
    }
  }

  OMP_PARALLEL_FOR
  for (auto i = 0; i < num_spins; ++i) {
    Vec3 spin = {sim.s(i, 0), sim.s(i, 1), sim.s(i, 2)};
    Vec3 spin_old = {sim.s_old_(i, 0), sim.s_old_(i, 1), sim.s_old_(i, 2)};
    Vec3 field = {sim.h(i, 0), sim.h(i, 1), sim.h(i, 2)};
    Vec3 rhs = -sim.gyro(i) * (cross(spin, field) + sim.alpha(i) * cross(spin, (cross(spin, field))));
    for (auto j = 0; j < 3; ++j) {
      sim.ds_dt(i, j) = sim.ds_dt(i, j) + 0.5 * rhs[j];
    }
    Vec3 ds = {sim.ds_dt(i, 0), sim.ds_dt(i, 1), sim.ds_dt(i, 2)};
    spin = unit_vector(spin_old + sim.step_size_ * ds);
    for (auto j = 0; j < 3; ++j) {
      sim.s(i, j) = spin[j];
    }
  }

  sim.iteration_++;
  sim.time_ = sim.iteration_ * sim.step_size_;
}

template <typename F>
int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const std::exception &e) {
    g_error = e.what();
    return 1;
  } catch (...) {
    g_error = "unknown exception";
    return 1;
  }
}

}  // namespace

extern "C" {

const char *jref_last_error() { return g_error.c_str(); }
int jref_have_pcg() { return JREF_HAVE_PCG; }
int jref_omp_threads() {
#if HAS_OMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- scalar helpers straight from the reference headers -------------------------------
int jref_approximately_equal(double a, double b, double eps) { return approximately_equal(a, b, eps) ? 1 : 0; }
int jref_approximately_zero(double a, double eps) { return approximately_zero(a, eps) ? 1 : 0; }
void jref_unit_vector(const double *a, double *out) {
  Vec3 r = unit_vector(Vec3{a[0], a[1], a[2]});
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
void jref_llg_rhs(const double *s, const double *hf, double gyro, double alpha, double *out) {
  Vec3 spin = {s[0], s[1], s[2]};
  Vec3 field = {hf[0], hf[1], hf[2]};
  Vec3 rhs = -gyro * (cross(spin, field) + alpha * cross(spin, (cross(spin, field))));
  out[0] = rhs[0]; out[1] = rhs[1]; out[2] = rhs[2];
}
void jref_rotation_matrix_between_vectors(const double *a, const double *b, double *R9) {
  Mat3 R = rotation_matrix_between_vectors(Vec3{a[0], a[1], a[2]}, Vec3{b[0], b[1], b[2]});
  for (int m = 0; m < 3; ++m) for (int n = 0; n < 3; ++n) R9[3 * m + n] = R[m][n];
}
void jref_mat3_inverse(const double *A9, double *out9) {
  Mat3 A; for (int m = 0; m < 3; ++m) for (int n = 0; n < 3; ++n) A[m][n] = A9[3 * m + n];
  Mat3 R = inverse(A);
  for (int m = 0; m < 3; ++m) for (int n = 0; n < 3; ++n) out9[3 * m + n] = R[m][n];
}

// ---- jams::InteractionList<Mat3,2>: insert in caller order, read back in stored order ---
// Mirrors the use in neighbour_list_from_interactions (core/interactions.cc:373-388): a pair
// that is already present is an error.  Returns 0 ok, 1 error (duplicate), sizes via out params.
int jref_interaction_list(int64_t n, const int *i, const int *j, const double *J9,
                          int *out_i, int *out_j, int *out_value_id,
                          int *out_n_values, double *out_values9 /* capacity n*9 */) {
  return guarded([&]() {
    jams::InteractionList<Mat3, 2> list;
    std::vector<Mat3> table;  // independent replay of the value table order
    for (int64_t p = 0; p < n; ++p) {
      if (list.contains({i[p], j[p]})) {
        throw std::runtime_error("Multiple interactions for sites " + std::to_string(i[p]) + " and " + std::to_string(j[p]));
      }
      Mat3 J; for (int m = 0; m < 3; ++m) for (int q = 0; q < 3; ++q) J[m][q] = J9[9 * p + 3 * m + q];
      list.insert({i[p], j[p]}, J);
      if (std::find(table.begin(), table.end(), J) == table.end()) table.push_back(J);
    }
    for (int p = 0; p < list.size(); ++p) {
      auto item = list[p];
      out_i[p] = item.first[0];
      out_j[p] = item.first[1];
      auto it = std::find(table.begin(), table.end(), item.second);
      out_value_id[p] = int(it - table.begin());
    }
    *out_n_values = int(table.size());
    for (std::size_t v = 0; v < table.size(); ++v)
      for (int m = 0; m < 3; ++m) for (int q = 0; q < 3; ++q) out_values9[9 * v + 3 * m + q] = table[v][m][q];
  });
}

// ---- simulation object ------------------------------------------------------------------
void *jref_sim_create(int num_spins, const double *mus, const double *gyro, const double *alpha) {
  auto *sim = new RefSim;
  sim->num_spins = num_spins;
  sim->num_spins3 = 3 * num_spins;
  sim->s.resize(num_spins, 3); sim->h.resize(num_spins, 3); sim->ds_dt.resize(num_spins, 3);
  sim->s.zero(); sim->h.zero(); sim->ds_dt.zero();
  sim->mus.resize(num_spins); sim->gyro.resize(num_spins); sim->alpha.resize(num_spins);
  for (int i = 0; i < num_spins; ++i) { sim->mus(i) = mus[i]; sim->gyro(i) = gyro[i]; sim->alpha(i) = alpha[i]; }
  return sim;
}
void jref_sim_destroy(void *p) { delete static_cast<RefSim *>(p); }

// ExchangeHamiltonian ctor tail (hamiltonian/exchange.cc:162-171) on an already scaled pair list:
// insert_interaction_tensor (sparse_interaction.cc:25-34) for each pair, then finalize() with the
// Symmetric check (sparse_interaction.cc:102-138).  J9 is per pair, row-major, in meV.
int jref_sim_add_exchange(void *p, int64_t n_pairs, const int *i, const int *j, const double *J9, int check_symmetric) {
  auto *sim = static_cast<RefSim *>(p);
  return guarded([&]() {
    jams::SparseMatrix<double>::Builder sparse_matrix_builder_(3 * sim->num_spins, 3 * sim->num_spins);
    for (int64_t q = 0; q < n_pairs; ++q) {
      for (auto m = 0; m < 3; ++m) {
        for (auto n = 0; n < 3; ++n) {
          const double value = J9[9 * q + 3 * m + n];
          if (value != 0.0) {
            sparse_matrix_builder_.insert(3 * i[q] + m, 3 * j[q] + n, value);
          }
        }
      }
    }
    if (check_symmetric && !sparse_matrix_builder_.is_symmetric()) {
      throw std::runtime_error("sparse matrix for exchange is not symmetric");
    }
    auto t = std::make_unique<RefSim::Term>();
    t->kind = RefSim::EXCHANGE;
    t->field_.resize(sim->num_spins, 3); t->field_.zero();
    t->interaction_matrix_ = sparse_matrix_builder_.set_format(jams::SparseMatrixFormat::CSR).build();
    sim->hamiltonians_.push_back(std::move(t));
  });
}

int jref_sim_add_uniaxial(void *p, int power, const double *magnitude, const double *axis) {
  auto *sim = static_cast<RefSim *>(p);
  return guarded([&]() {
    auto t = std::make_unique<RefSim::Term>();
    t->kind = RefSim::UNIAXIAL;
    t->power_ = power;
    t->field_.resize(sim->num_spins, 3); t->field_.zero();
    t->magnitude_.resize(sim->num_spins); t->axis_.resize(sim->num_spins, 3);
    for (int i = 0; i < sim->num_spins; ++i) {
      t->magnitude_(i) = magnitude[i];
      for (int j = 0; j < 3; ++j) t->axis_(i, j) = axis[3 * i + j];
    }
    sim->hamiltonians_.push_back(std::move(t));
  });
}

int jref_sim_add_zeeman(void *p, const double *dc_local_field, const double *ac_local_field, const double *ac_local_frequency) {
  auto *sim = static_cast<RefSim *>(p);
  return guarded([&]() {
    auto t = std::make_unique<RefSim::Term>();
    t->kind = RefSim::ZEEMAN;
    t->field_.resize(sim->num_spins, 3); t->field_.zero();
    t->dc_local_field_.resize(sim->num_spins, 3);
    t->ac_local_field_.resize(sim->num_spins, 3); t->ac_local_field_.zero();
    t->ac_local_frequency_.resize(sim->num_spins); t->ac_local_frequency_.zero();
    for (int i = 0; i < sim->num_spins; ++i) for (int j = 0; j < 3; ++j) t->dc_local_field_(i, j) = dc_local_field[3 * i + j];
    if (ac_local_field && ac_local_frequency) {
      t->has_ac_local_field_ = true;
      for (int i = 0; i < sim->num_spins; ++i) {
        for (int j = 0; j < 3; ++j) t->ac_local_field_(i, j) = ac_local_field[3 * i + j];
        t->ac_local_frequency_(i) = ac_local_frequency[i];
      }
    }
    sim->hamiltonians_.push_back(std::move(t));
  });
}

// CSR of exchange term `term` (index into the registration order)
int64_t jref_sim_exchange_nnz(void *p, int term) {
  auto *sim = static_cast<RefSim *>(p);
  return sim->hamiltonians_.at(term)->interaction_matrix_.num_non_zero();
}
void jref_sim_exchange_csr(void *p, int term, int *row, int *col, double *val) {
  auto *sim = static_cast<RefSim *>(p);
  auto &A = sim->hamiltonians_.at(term)->interaction_matrix_;
  std::copy(A.row_data(), A.row_data() + A.num_rows() + 1, row);
  std::copy(A.col_data(), A.col_data() + A.num_non_zero(), col);
  std::copy(A.val_data(), A.val_data() + A.num_non_zero(), val);
}

void jref_sim_set_spins(void *p, const double *s_aos) {
  auto *sim = static_cast<RefSim *>(p);
  std::copy(s_aos, s_aos + sim->num_spins3, sim->s.data());
}
void jref_sim_get_spins(void *p, double *s_aos) {
  auto *sim = static_cast<RefSim *>(p);
  std::copy(sim->s.data(), sim->s.data() + sim->num_spins3, s_aos);
}
void jref_sim_get_h(void *p, double *h_aos) {
  auto *sim = static_cast<RefSim *>(p);
  std::copy(sim->h.data(), sim->h.data() + sim->num_spins3, h_aos);
}

// HeunLLGSolver::initialize (solvers/cpu_llg_heun.cc:15-43); dt in ps.
void jref_sim_init_solver(void *p, double step_size_ps, int use_gilbert_prefactor, uint64_t seed) {
  auto *sim = static_cast<RefSim *>(p);
  sim->step_size_ = step_size_ps;
  sim->time_ = 0.0;
  sim->iteration_ = 0;
  sim->s_old_.resize(sim->num_spins, 3);
  sim->sigma_.resize(sim->num_spins);
  sim->w_.resize(sim->num_spins, 3);
  sim->w_.zero();
  for (int i = 0; i < sim->num_spins; ++i) {
    double denominator = 1.0;
    if (use_gilbert_prefactor) {
      denominator = 1.0 + pow2(sim->alpha(i));
    }
    sim->sigma_(i) = sqrt((2.0 * kBoltzmannIU * sim->alpha(i)) /
                          (sim->mus(i) * sim->gyro(i) * sim->step_size_ * denominator));
  }
  sim->fallback_rng_.seed(seed);
#if JREF_HAVE_PCG
  sim->random_generator_ = arrow_vendored::pcg32_k1024(seed);
#endif
}
void jref_sim_get_sigma(void *p, double *sigma) {
  auto *sim = static_cast<RefSim *>(p);
  std::copy(sim->sigma_.data(), sim->sigma_.data() + sim->num_spins, sigma);
}
void jref_sim_set_temperature(void *p, double T) { static_cast<RefSim *>(p)->temperature_ = T; }
double jref_sim_time(void *p) { return static_cast<RefSim *>(p)->time_; }

// nsteps calls of HeunLLGSolver::run; normals is nullptr or nsteps*3N standard normals
void jref_sim_run(void *p, int nsteps, const double *normals) {
  auto *sim = static_cast<RefSim *>(p);
  for (int n = 0; n < nsteps; ++n) {
    heun_run(*sim, normals ? normals + std::size_t(n) * sim->num_spins3 : nullptr);
  }
}

// Hamiltonian::calculate_fields for one term at `time`, result = its field_ (meV)
void jref_sim_term_fields(void *p, int term, double time, double *field_aos) {
  auto *sim = static_cast<RefSim *>(p);
  auto &t = *sim->hamiltonians_.at(term);
  calculate_fields(*sim, t, time);
  std::copy(t.field_.data(), t.field_.data() + sim->num_spins3, field_aos);
}

// calculate_total_energy per term: exchange (sparse_interaction.cc:86-100), uniaxial
// (uniaxial_anisotropy.cc:118-133), Zeeman (zeeman.cc:74-87)
double jref_sim_term_total_energy(void *p, int term, double time) {
  auto *sim = static_cast<RefSim *>(p);
  auto &t = *sim->hamiltonians_.at(term);
  double e_total = 0.0;
  switch (t.kind) {
    case RefSim::EXCHANGE: {
      calculate_fields(*sim, t, time);
      double total_energy = 0.0;
      for (auto i = 0; i < sim->num_spins; ++i) {
        Vec3 s_i = {sim->s(i, 0), sim->s(i, 1), sim->s(i, 2)};
        Vec3 h_i = {t.field_(i, 0), t.field_(i, 1), t.field_(i, 2)};
        total_energy += -dot(s_i, h_i);
      }
      return 0.5 * total_energy;
    }
    case RefSim::UNIAXIAL:
      for (int i = 0; i < sim->num_spins; ++i) {
        auto dot = (t.axis_(i, 0) * sim->s(i, 0) + t.axis_(i, 1) * sim->s(i, 1) + t.axis_(i, 2) * sim->s(i, 2));
        e_total += (-t.magnitude_(i) * pow(dot, t.power_));
      }
      return e_total;
    case RefSim::ZEEMAN:
      for (int i = 0; i < sim->num_spins; ++i) {
        Vec3 s_i = {sim->s(i, 0), sim->s(i, 1), sim->s(i, 2)};
        Vec3 field = {t.dc_local_field_(i, 0), t.dc_local_field_(i, 1), t.dc_local_field_(i, 2)};
        if (t.has_ac_local_field_) {
          for (int j = 0; j < 3; ++j) field[j] += t.ac_local_field_(i, j) * cos(t.ac_local_frequency_(i) * time);
        }
        e_total += -dot(s_i, field);
      }
      return e_total;
  }
  return 0.0;
}

}  // extern "C"
