// oracle/jams_oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// A self-contained CPU restatement (no reference headers, no third-party code) of the
// stonerlab/jams llg-heun-cpu hot path and of the lattice / neighbour-list construction
// that feeds it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` legs may load the library built from this file; the product path
// (jams_b200/) must never import, link or call it.
//
// Parity status: PINNED.  tests/test_oracle_cpu.py checks this file bit-for-bit against
//   (1) oracle/_ref/libjams_ref.so, i.e. the reference's own SparseMatrix / InteractionList /
//       Vec3 / Mat3 code compiled from /root/reference/src (when that library is present), and
//   (2) the golden vectors in tests/golden/ that were generated from (1) by
//       tests/golden/make_golden.py, plus the reference's known answers
//       (sc 8^3 NN -> 3072 interactions, src/jams/test/interactions.h:241-252).
// Three restatements follow code the reference has only as CUDA kernels and ships no test or vector for: rk4_run
//   (solvers/cuda_rk4_base.cu), the biquadratic-exchange term (hamiltonian/cuda_biquadratic_exchange_kernel.cuh) and, in
//   oracle/__init__.py, pin_region (physics/pinned_boundaries.cc).  They are PINNED on the GPU box against those kernels
//   themselves -- oracle/_ref/libjams_ref_cuda.so, the reference's CUDA sources compiled for sm_100a by oracle/ref_cuda_wrap.cu:
//   tests/test_gpu_reference_cuda.py (<= 1e-12 RK4 trajectories, field and rotation to rounding) -- and, on the CPU, to the
//   mathematics: fourth-order convergence / fixed point / analytic precession (RK4), h = -dE/ds by finite differences
//   (biquadratic), the golden rotation_matrix_between_vectors vectors (pin_region): tests/test_oracle_cpu.py.
//
// All citations are file:line under /root/reference/src/jams/.
// Arithmetic is written out in the same operand order as the reference so that, compiled with
// the reference's flags (no FMA contraction on baseline x86-64), results are bit-identical.

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <stdexcept>
#include <limits>
#include <string>
#include <vector>
#if defined(_OPENMP)
#include <omp.h>
#endif

namespace {

thread_local std::string g_error;

using V3 = std::array<double, 3>;
using M3 = std::array<double, 9>;  // row-major

// ---- helpers/maths.h:16-34 ---------------------------------------------------------------
inline bool approximately_equal(double a, double b, double epsilon) {
  if (std::abs(a - b) <= epsilon) return true;
  return std::abs(a - b) <= (std::max(std::abs(a), std::abs(b)) * epsilon);
}
inline bool approximately_zero(double a, double epsilon) { return std::abs(a) <= epsilon; }
inline bool definately_greater_than(double a, double b, double epsilon) {
  return (a - b) > (std::max(std::abs(a), std::abs(b)) * epsilon);
}
inline bool definately_less_than(double a, double b, double epsilon) {
  return (b - a) > (std::max(std::abs(a), std::abs(b)) * epsilon);
}
// containers/vec3.h:264-271
inline bool approximately_equal(const V3 &a, const V3 &b, double epsilon) {
  for (int n = 0; n < 3; ++n) if (!approximately_equal(a[n], b[n], epsilon)) return false;
  return true;
}

// ---- containers/vec3.h:135-137,165-167,184-188,276-283 ------------------------------------
inline double dot(const V3 &a, const V3 &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double norm(const V3 &a) { return std::sqrt(dot(a, a)); }
inline V3 cross(const V3 &a, const V3 &b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}
inline V3 unit_vector(const V3 &a) {
  const double length = norm(a);
  if (approximately_zero(length, DBL_EPSILON)) return a;
  return {a[0] / length, a[1] / length, a[2] / length};
}

// ---- containers/mat3.h:27-33,73-83 ---------------------------------------------------------
inline V3 matvec(const M3 &A, const V3 &x) {
  return {A[0] * x[0] + A[1] * x[1] + A[2] * x[2],
          A[3] * x[0] + A[4] * x[1] + A[5] * x[2],
          A[6] * x[0] + A[7] * x[1] + A[8] * x[2]};
}
inline M3 matmul(const M3 &A, const M3 &B) {
  M3 R = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) R[3 * i + j] += A[3 * i + k] * B[3 * k + j];
  return R;
}
// The reference inverts the 3x3 cell with LAPACK dgetrf/dgetri (containers/mat3.h:233-262).
// Results are only ever used after snapping with 1e-4 tolerances, so the cofactor
// form is used here.
inline M3 inverse(const M3 &A) {
  const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
  M3 R;
  R[0] = (A[4] * A[8] - A[5] * A[7]) / det; R[1] = (A[2] * A[7] - A[1] * A[8]) / det; R[2] = (A[1] * A[5] - A[2] * A[4]) / det;
  R[3] = (A[5] * A[6] - A[3] * A[8]) / det; R[4] = (A[0] * A[8] - A[2] * A[6]) / det; R[5] = (A[2] * A[3] - A[0] * A[5]) / det;
  R[6] = (A[3] * A[7] - A[4] * A[6]) / det; R[7] = (A[1] * A[6] - A[0] * A[7]) / det; R[8] = (A[0] * A[4] - A[1] * A[3]) / det;
  return R;
}
inline double max_abs(const double *J9) {  // containers/mat3.h:298-308
  double m = 0.0;
  for (int i = 0; i < 9; ++i) if (std::abs(J9[i]) > m) m = std::abs(J9[i]);
  return m;
}

// helpers/consts.h:29-34
constexpr double kBoltzmannIU = 0.0861733326;
constexpr double kPi = 3.14159265358979323846264338327950288;

template <typename F>
int guarded(F &&f) {
  try { f(); return 0; }
  catch (const std::exception &e) { g_error = e.what(); return 1; }
  catch (...) { g_error = "unknown exception"; return 1; }
}

// ---- core/lattice.cc:48-64 -----------------------------------------------------------------
V3 normalise_fractional_coordinate(V3 r_frac, double eps = 1e-4) {
  for (int n = 0; n < 3; ++n) {
    if (r_frac[n] < 0.0) r_frac[n] = r_frac[n] + 1.0;
    if (approximately_equal(r_frac[n], 1.0, eps)) r_frac[n] = 0.0;
  }
  return r_frac;
}

// ---- core/interactions.cc:58-76 --------------------------------------------------------------
V3 lattice_translation_vector(const V3 &r_frac, double tolerance) {
  V3 T;
  for (int n = 0; n < 3; ++n) {
    double nearest_integer = std::nearbyint(r_frac[n]);
    double floored_value = std::floor(r_frac[n]);
    if (approximately_zero(r_frac[n] - nearest_integer, tolerance)) T[n] = nearest_integer;
    else T[n] = floored_value;
  }
  return T;
}

struct TemplateEntry {
  int basis_site_i, basis_site_j;
  int type_i, type_j;
  V3 r_cart;
  int T[3];
  std::array<double, 9> J;
};

}  // namespace

extern "C" {

const char *jo_last_error() { return g_error.c_str(); }
int jo_omp_threads() {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Lattice::generate_supercell numbering (core/lattice.cc:622-657): the loop nest i,j,k,m with a
// running counter, i.e. site = ((i*Ny + j)*Nz + k)*M + m.
int64_t jo_site_index(const int *dims, int M, int i, int j, int k, int m) {
  return ((int64_t(i) * dims[1] + j) * dims[2] + k) * M + m;
}

// Lattice::apply_boundary_conditions (core/lattice.cc:987-1007). Returns 0 if rejected.
int jo_apply_boundary_conditions(const int *dims, const int *periodic, int *abc) {
  for (int l = 0; l < 3; ++l) {
    if (!periodic[l] && (abc[l] < 0 || abc[l] >= dims[l])) return 0;
    abc[l] = (abc[l] + dims[l]) % dims[l];
  }
  return 1;
}

// post_process_interactions (core/interactions.cc:292-347) for settings-style input
// (interactions_from_settings, core/interactions.cc:253-289).
//   cell9        unit cell matrix, row-major, columns are a,b,c (core/lattice.cc:356-367)
//   motif_frac   M x 3 normalised fractional motif positions; motif_type: material id per motif site
//   format       0 = JAMS (type_i/type_j are material ids; motif sites are completed by
//                    complete_interaction_unitcell_positions, core/interactions.cc:98-124)
//                1 = KKR  (type_i/type_j are 0-based motif indices; material ids are filled from the motif)
//   r            n_in x 3 interaction vectors, cartesian unless frac_coords != 0
//   J9           n_in x 9 tensors in the INPUT energy unit (the cut-off compares in that unit)
//   sym_rot9/sym_trans  n_ops space-group operations in fractional basis (what spglib hands
//                the reference, core/lattice.cc:942-967); used only if use_symops
// Outputs (capacity cap entries): basis sites, integer cell translations, tensors, vectors.
// Returns the number of template entries, or -1 on error / -2 if cap is too small.
int64_t jo_expand_template(const double *cell9, int M, const double *motif_frac, const int *motif_type,
                           int format, int frac_coords, int64_t n_in, const int *type_i, const int *type_j,
                           const double *r, const double *J9,
                           int use_symops, int n_ops, const double *sym_rot9, const double *sym_trans,
                           double energy_cutoff, double radius_cutoff, double distance_tolerance,
                           int64_t cap, int *out_mi, int *out_mj, int *out_T, double *out_J9, double *out_r) {
  int64_t result = -1;
  guarded([&]() {
    const double lattice_tolerance = 1e-4;  // helpers/defaults.h:44
    M3 A; std::copy(cell9, cell9 + 9, A.begin());
    const M3 Ainv = inverse(A);
    auto frac_to_cart = [&](const V3 &f) { return matvec(A, f); };     // containers/cell.h:57
    auto cart_to_frac = [&](const V3 &c) { return matvec(Ainv, c); };  // containers/cell.h:56
    auto motif_pos = [&](int k) { return V3{motif_frac[3 * k], motif_frac[3 * k + 1], motif_frac[3 * k + 2]}; };

    std::vector<TemplateEntry> interactions;
    for (int64_t n = 0; n < n_in; ++n) {
      TemplateEntry J{};
      J.basis_site_i = -1; J.basis_site_j = -1; J.type_i = -1; J.type_j = -1;
      if (format == 1) { J.basis_site_i = type_i[n]; J.basis_site_j = type_j[n]; }
      else { J.type_i = type_i[n]; J.type_j = type_j[n]; }
      J.r_cart = {r[3 * n], r[3 * n + 1], r[3 * n + 2]};
      std::copy(J9 + 9 * n, J9 + 9 * n + 9, J.J.begin());
      interactions.push_back(J);
    }

    if (frac_coords) for (auto &J : interactions) J.r_cart = frac_to_cart(J.r_cart);  // :293-298

    if (format == 0) {
      // complete_interaction_unitcell_positions (:98-124) with find_unitcell_partner (:78-87)
      // and find_basis_site_index (:44-54)
      std::vector<TemplateEntry> new_data;
      for (const auto &J : interactions) {
        for (int i = 0; i < M; ++i) {
          auto new_J = J;
          if (motif_type[i] != J.type_i) continue;
          new_J.basis_site_i = i;
          V3 p_i_frac = motif_pos(i);
          V3 r_ij_frac = cart_to_frac(J.r_cart);
          V3 q_ij = {r_ij_frac[0] + p_i_frac[0], r_ij_frac[1] + p_i_frac[1], r_ij_frac[2] + p_i_frac[2]};
          V3 T = lattice_translation_vector(q_ij, distance_tolerance);
          V3 offset = {q_ij[0] - T[0], q_ij[1] - T[1], q_ij[2] - T[2]};
          int partner = -1;
          for (int k = 0; k < M; ++k) {
            if (approximately_equal(motif_pos(k), offset, distance_tolerance)) { partner = k; break; }
          }
          if (partner < 0) continue;
          if (motif_type[partner] != J.type_j) continue;
          new_J.basis_site_j = partner;
          new_data.push_back(new_J);
        }
      }
      interactions.swap(new_data);
    } else {
      // complete_interaction_typenames_names (:89-96)
      for (auto &J : interactions) { J.type_i = motif_type[J.basis_site_i]; J.type_j = motif_type[J.basis_site_j]; }
    }

    if (use_symops) {
      // Lattice::lattice_site_point_group_symops (core/lattice.cc:1127-1153)
      std::vector<std::vector<M3>> point_group(M);
      for (int m = 0; m < M; ++m) {
        V3 motif_position = motif_pos(m);
        for (int n = 0; n < n_ops; ++n) {
          V3 tr = {sym_trans[3 * n], sym_trans[3 * n + 1], sym_trans[3 * n + 2]};
          bool zero = true;  // approximately_zero(Vec3) containers/vec3.h:251-258
          for (int c = 0; c < 3; ++c) if (!approximately_zero(tr[c], lattice_tolerance)) zero = false;
          if (!zero) continue;
          M3 rotation; std::copy(sym_rot9 + 9 * n, sym_rot9 + 9 * n + 9, rotation.begin());
          V3 new_position = normalise_fractional_coordinate(matvec(rotation, motif_position));
          if (approximately_equal(motif_position, new_position, lattice_tolerance)) point_group[m].push_back(rotation);
        }
      }
      // apply_symops (core/interactions.cc:24-37) + Lattice::generate_symmetric_points (core/lattice.cc:1015-1035)
      std::vector<TemplateEntry> symops_interaction_data;
      for (const auto &J : interactions) {
        auto new_J = J;
        const V3 r_frac = cart_to_frac(J.r_cart);
        std::vector<V3> symmetric_points;
        symmetric_points.push_back(J.r_cart);
        for (const auto &rotation_matrix : point_group[J.basis_site_i]) {
          const V3 r_sym = frac_to_cart(matvec(rotation_matrix, r_frac));
          bool exists = false;
          for (const auto &v2 : symmetric_points) if (approximately_equal(r_sym, v2, lattice_tolerance)) { exists = true; break; }
          if (!exists) symmetric_points.push_back(r_sym);
        }
        for (const auto &p : symmetric_points) { new_J.r_cart = p; symops_interaction_data.push_back(new_J); }
      }
      interactions.swap(symops_interaction_data);
    }

    // predicates (:322-330); apply_predicate removes entries for which the predicate is true
    if (energy_cutoff > 0.0) {
      interactions.erase(std::remove_if(interactions.begin(), interactions.end(), [&](const TemplateEntry &J) {
        return definately_less_than(max_abs(J.J.data()), energy_cutoff, DBL_EPSILON); }), interactions.end());
    }
    if (radius_cutoff > 0.0) {
      interactions.erase(std::remove_if(interactions.begin(), interactions.end(), [&](const TemplateEntry &J) {
        return definately_greater_than(norm(J.r_cart), radius_cutoff, lattice_tolerance); }), interactions.end());
    }

    // lattice translation vectors (:333-345)
    for (auto &J : interactions) {
      V3 p_i_frac = motif_pos(J.basis_site_i);
      V3 p_j_frac = motif_pos(J.basis_site_j);
      V3 r_ij_frac = cart_to_frac(J.r_cart);
      V3 q = {r_ij_frac[0] + p_i_frac[0] - p_j_frac[0], r_ij_frac[1] + p_i_frac[1] - p_j_frac[1], r_ij_frac[2] + p_i_frac[2] - p_j_frac[2]};
      V3 T = lattice_translation_vector(q, distance_tolerance);
      J.T[0] = int(T[0]); J.T[1] = int(T[1]); J.T[2] = int(T[2]);
    }

    if (int64_t(interactions.size()) > cap) { result = -2; g_error = "template capacity too small"; return; }
    for (std::size_t n = 0; n < interactions.size(); ++n) {
      const auto &J = interactions[n];
      out_mi[n] = J.basis_site_i; out_mj[n] = J.basis_site_j;
      for (int c = 0; c < 3; ++c) { out_T[3 * n + c] = J.T[c]; out_r[3 * n + c] = J.r_cart[c]; }
      std::copy(J.J.begin(), J.J.end(), out_J9 + 9 * n);
    }
    result = int64_t(interactions.size());
  });
  return result;
}

// neighbour_list_from_interactions (core/interactions.cc:349-395) followed by the storage order of
// jams::InteractionList (containers/interaction_list.h:28-39: pairs kept sorted by {i,j};
// values de-duplicated in first-insertion order, containers/unordered_vector_set.h:38-45).
//   site_type   N material ids (Lattice::lattice_site_material_id), type_of_entry_{i,j} per template entry
// Returns number of pairs, -1 on error (duplicate pair => "Multiple interactions ..."), -2 if cap too small.
int64_t jo_neighbour_list(const int *dims, const int *periodic, int M, const int *site_type,
                          int64_t n_template, const int *mi, const int *mj, const int *T,
                          const int *entry_type_i, const int *entry_type_j, const double *J9,
                          int64_t cap_pairs, int *out_i, int *out_j, int *out_value_id,
                          int cap_values, int *out_n_values, double *out_values9) {
  int64_t result = -1;
  guarded([&]() {
    struct Pair { int i, j, v; };
    std::vector<Pair> pairs;
    std::vector<std::array<double, 9>> table;
    for (int i = 0; i < dims[0]; ++i) {
      for (int j = 0; j < dims[1]; ++j) {
        for (int k = 0; k < dims[2]; ++k) {
          for (int64_t n = 0; n < n_template; ++n) {
            const int m = mi[n];
            int local_site = int(jo_site_index(dims, M, i, j, k, m));
            int d[3] = {i + T[3 * n], j + T[3 * n + 1], k + T[3 * n + 2]};
            if (!jo_apply_boundary_conditions(dims, periodic, d)) continue;
            int nbr_site = int(jo_site_index(dims, M, d[0], d[1], d[2], mj[n]));
            // the duplicate test happens before the material test in the reference (:373-386);
            // it is applied after sorting below, which is equivalent because only inserted pairs
            // are ever compared against... except that a pair skipped by the material test is
            // never inserted, so it can not collide either.
            if (site_type[local_site] != entry_type_i[n] || site_type[nbr_site] != entry_type_j[n]) {
              // still need the reference's order of checks: a collision with an *inserted* pair
              // throws even if this entry would then have been skipped for its material.
              pairs.push_back({local_site, nbr_site, -1});
              continue;
            }
            std::array<double, 9> val; std::copy(J9 + 9 * n, J9 + 9 * n + 9, val.begin());
            int v = -1;
            for (std::size_t q = 0; q < table.size(); ++q) if (table[q] == val) { v = int(q); break; }
            if (v < 0) { table.push_back(val); v = int(table.size()) - 1; }
            pairs.push_back({local_site, nbr_site, v});
          }
        }
      }
    }
    // Reproduce "contains() -> throw" (:373-381): an entry (inserted or material-skipped) that
    // arrives when an inserted entry with the same {i,j} already exists is an error.  With a
    // stable sort by {i,j}, generation order is preserved inside equal runs.
    std::stable_sort(pairs.begin(), pairs.end(), [](const Pair &a, const Pair &b) {
      return a.i != b.i ? a.i < b.i : a.j < b.j; });
    std::vector<Pair> kept;
    kept.reserve(pairs.size());
    for (std::size_t p = 0; p < pairs.size();) {
      std::size_t q = p;
      bool have_inserted = false;
      while (q < pairs.size() && pairs[q].i == pairs[p].i && pairs[q].j == pairs[p].j) {
        if (have_inserted) {
          throw std::runtime_error("Multiple interactions for sites " + std::to_string(pairs[p].i) + " and " + std::to_string(pairs[p].j));
        }
        if (pairs[q].v >= 0) { have_inserted = true; kept.push_back(pairs[q]); }
        ++q;
      }
      p = q;
    }
    if (int64_t(kept.size()) > cap_pairs || int(table.size()) > cap_values) { result = -2; g_error = "pair/value capacity too small"; return; }
    for (std::size_t p = 0; p < kept.size(); ++p) { out_i[p] = kept[p].i; out_j[p] = kept[p].j; out_value_id[p] = kept[p].v; }
    *out_n_values = int(table.size());
    for (std::size_t v = 0; v < table.size(); ++v) std::copy(table[v].begin(), table[v].end(), out_values9 + 9 * v);
    result = int64_t(kept.size());
  });
  return result;
}

// rotation_matrix_between_vectors (containers/mat3.h:334-366) incl. the axis-angle branch (:323-331)
void jo_rotation_matrix_between_vectors(const double *a, const double *b, double *R9) {
  const M3 I = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  auto ssc = [](const V3 &v) { return M3{0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0}; };
  const V3 ua = unit_vector(V3{a[0], a[1], a[2]});
  const V3 ub = unit_vector(V3{b[0], b[1], b[2]});
  const double c = dot(ua, ub);
  M3 R;
  if (approximately_equal(c, 1.0, 1e-12)) {
    R = I;
  } else if (approximately_equal(c, -1.0, 1e-12)) {
    V3 ortho = std::abs(ua[0]) < 0.9 ? V3{1, 0, 0} : V3{0, 1, 0};
    V3 axis = unit_vector(cross(ua, ortho));
    const V3 u = unit_vector(axis);
    const double cc = std::cos(kPi), ss = std::sin(kPi);
    const M3 vx = ssc(u);
    const M3 vx2 = matmul(vx, vx);
    // kIdentityMat3 + s * vx + (1.0 - c) * vx * vx   evaluated left to right: ((1-c)*vx)*vx
    M3 t; for (int n = 0; n < 9; ++n) t[n] = (1.0 - cc) * vx[n];
    const M3 t2 = matmul(t, vx);
    (void)vx2;
    for (int n = 0; n < 9; ++n) R[n] = (I[n] + ss * vx[n]) + t2[n];
  } else {
    V3 v = cross(ua, ub);
    const double s = norm(v);
    M3 vx = ssc(v);
    const double k = (1.0 - c) / (s * s);
    M3 t; for (int n = 0; n < 9; ++n) t[n] = k * vx[n];
    const M3 t2 = matmul(t, vx);  // k * vx * vx == (k*vx)*vx
    for (int n = 0; n < 9; ++n) R[n] = (I[n] + vx[n]) + t2[n];
  }
  std::copy(R.begin(), R.end(), R9);
}

// InitBlochDomainWall::execute (initializer/init_bloch_domain_wall.cc:10-32); positions are cartesian
// in lattice constants (Lattice::lattice_site_position_cart); spins are rotated in place.
void jo_init_bloch_domain_wall(int64_t N, const double *positions, double width, double center,
                               const double *normal_in, const double *domain_in, double *s_aos) {
  V3 normal = {normal_in[0], normal_in[1], normal_in[2]};
  V3 domain = {domain_in[0], domain_in[1], domain_in[2]};
  double ln = norm(normal), ld = norm(domain);  // normalize(): a / norm(a) (containers/vec3.h:244-246)
  normal = {normal[0] / ln, normal[1] / ln, normal[2] / ln};
  domain = {domain[0] / ld, domain[1] / ld, domain[2] / ld};
  for (int64_t i = 0; i < N; ++i) {
    V3 r = {positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]};
    double x = dot(r, normal) - center;
    V3 m = {0, 1.0 / std::cosh(kPi * x / width), std::tanh(kPi * x / width)};
    V3 spin = {s_aos[3 * i], s_aos[3 * i + 1], s_aos[3 * i + 2]};
    double R9[9];
    jo_rotation_matrix_between_vectors(domain.data(), m.data(), R9);
    M3 R; std::copy(R9, R9 + 9, R.begin());
    spin = matvec(R, spin);
    for (int n = 0; n < 3; ++n) s_aos[3 * i + n] = spin[n];
  }
}

}  // extern "C"

// =============================================================================================
// Simulation object: CSR exchange + uniaxial + Zeeman, Heun stepping, monitors
// =============================================================================================
namespace {

struct Term {
  enum Kind { EXCHANGE, UNIAXIAL, ZEEMAN, APPLIED, BIQUADRATIC } kind;   // BIQUADRATIC: row / col / val are the N x N scalar CSR
  std::vector<double> field;  // N x 3 (Hamiltonian::field_, core/hamiltonian.h)
  // exchange CSR (containers/sparse_matrix.h:270-275)
  std::vector<int> row, col;
  std::vector<double> val;
  // uniaxial
  int power = 2;
  std::vector<double> magnitude, axis;
  // zeeman
  std::vector<double> dc, ac, omega;
  bool has_ac = false;
  // applied field B(t) = B g(t) (hamiltonian/applied_field.cc:10-82): 0 static, 1 sinc, 2 sinc-cos; t0 in ps, frequencies in THz
  double B[3] = {0, 0, 0};
  int pulse_type = 0;
  double t0 = 0, fbw = 0, fc = 0;
};

// TimeDependentField::field(time) (applied_field.cc:18-19,41-43,70-73); sinc: helpers/maths.h:441-446
inline V3 applied_b(const Term &t, double time) {
  const double kPi = 3.14159265358979323846, kTwoPi = 2.0 * kPi;
  if (t.pulse_type == 0) return {t.B[0], t.B[1], t.B[2]};
  const double x = kPi * t.fbw * (time - t.t0);
  const double sinc = (x == 0.0) ? 1.0 : sin(x) / x;
  if (t.pulse_type == 1) return {t.B[0] * sinc, t.B[1] * sinc, t.B[2] * sinc};
  const double c = cos(kTwoPi * t.fc * (time - t.t0));
  return {t.B[0] * sinc * c, t.B[1] * sinc * c, t.B[2] * sinc * c};
}

struct Sim {
  int N = 0;
  std::vector<double> s, h, ds_dt, mus, gyro, alpha, s_old, w, sigma;
  std::vector<Term> terms;
  double dt = 0, time = 0, temperature = 0;
  int iteration = 0;
  std::mt19937_64 rng{12345};
};

// jams::Xcsrmv_general with alpha=1, beta=0 (interface/sparse_blas.h:57-73) and
// Xcsrmv_general_row (:13-27): ascending-j accumulation `sum += x[col[j]] * val[j]`.
void csrmv(const Term &t, int m, const double *x, double *y) {
#pragma omp parallel for
  for (int i = 0; i < m; ++i) {
    double sum = 0.0;
    for (int j = t.row[i]; j < t.row[i + 1]; ++j) sum += x[t.col[j]] * t.val[j];
    y[i] = sum;
  }
}

void calculate_fields(Sim &sim, Term &t, double time) {
  const int N = sim.N;
  switch (t.kind) {
    case Term::EXCHANGE:  // hamiltonian/sparse_interaction.cc:36-45
      csrmv(t, 3 * N, sim.s.data(), t.field.data());
      break;
    case Term::UNIAXIAL:  // hamiltonian/uniaxial_anisotropy.cc:155-172 (serial, std::pow)
      for (int i = 0; i < N; ++i) {
        double d = (t.axis[3 * i] * sim.s[3 * i] + t.axis[3 * i + 1] * sim.s[3 * i + 1] + t.axis[3 * i + 2] * sim.s[3 * i + 2]);
        for (int j = 0; j < 3; ++j) t.field[3 * i + j] = t.magnitude[i] * t.power * pow(d, t.power - 1) * t.axis[3 * i + j];
      }
      break;
    case Term::ZEEMAN:  // hamiltonian/zeeman.cc:121-132 (serial)
      for (int i = 0; i < N; ++i) {
        for (int j = 0; j < 3; ++j) t.field[3 * i + j] = t.dc[3 * i + j];
        if (t.has_ac) for (int j = 0; j < 3; ++j) t.field[3 * i + j] += t.ac[3 * i + j] * cos(t.omega[i] * time);
      }
      break;
    case Term::APPLIED: {  // hamiltonian/applied_field.cc:137-148: field_(i, j) = mus(i) * B(t)[j]
      const V3 b = applied_b(t, time);
      for (int i = 0; i < N; ++i) for (int j = 0; j < 3; ++j) t.field[3 * i + j] = sim.mus[i] * b[j];
      break;
    }
    case Term::BIQUADRATIC:  // hamiltonian/cuda_biquadratic_exchange_kernel.cuh:5-30 (the only field implementation of this term)
      for (int idx = 0; idx < N; ++idx) {
        double h_i[3] = {0.0, 0.0, 0.0};
        double s_i[3] = {sim.s[3 * idx + 0], sim.s[3 * idx + 1], sim.s[3 * idx + 2]};
        for (auto m = t.row[idx]; m < t.row[idx + 1]; ++m) {
          auto j = t.col[m];
          double s_j[3] = {sim.s[3 * j + 0], sim.s[3 * j + 1], sim.s[3 * j + 2]};
          double B_ij = t.val[m];
          double s_i_dot_s_j = s_i[0] * s_j[0] + s_i[1] * s_j[1] + s_i[2] * s_j[2];
          for (auto n = 0; n < 3; ++n) h_i[n] += 2.0 * B_ij * s_j[n] * s_i_dot_s_j;
        }
        for (auto n = 0; n < 3; ++n) t.field[3 * idx + n] = h_i[n];
      }
      break;
  }
}

// Solver::compute_fields (core/solver.cc:43-57): h = field_0; h += field_k (daxpy, alpha = 1)
void compute_fields(Sim &sim) {
  if (sim.terms.empty()) return;
  for (auto &t : sim.terms) calculate_fields(sim, t, sim.time);
  std::copy(sim.terms[0].field.begin(), sim.terms[0].field.end(), sim.h.begin());
  for (std::size_t k = 1; k < sim.terms.size(); ++k) {
    const double *x = sim.terms[k].field.data();
    for (int n = 0; n < 3 * sim.N; ++n) sim.h[n] += 1.0 * x[n];
  }
}

inline V3 llg_rhs(const V3 &spin, const V3 &field, double gyro, double alpha) {
  // Vec3 rhs = -gyro * (cross(spin, field) + alpha * cross(spin, cross(spin, field)))  (cpu_llg_heun.cc:89,130)
  const V3 sxh = cross(spin, field);
  const V3 sxsxh = cross(spin, sxh);
  const double mg = -gyro;
  return {mg * (sxh[0] + alpha * sxsxh[0]), mg * (sxh[1] + alpha * sxsxh[1]), mg * (sxh[2] + alpha * sxsxh[2])};
}

// HeunLLGSolver::run (solvers/cpu_llg_heun.cc:45-148)
void heun_run(Sim &sim, const double *normals) {
  const int N = sim.N;
  const double t0 = sim.time;
  sim.s_old = sim.s;  // :51
  const bool thermal = sim.temperature > 0.0;
  if (thermal) {  // :53-64
    if (normals) std::copy(normals, normals + 3 * N, sim.w.begin());
    else { std::normal_distribution<> nd; for (auto &x : sim.w) x = nd(sim.rng); }
    const double sqrt_temperature = sqrt(sim.temperature);
#pragma omp parallel for
    for (int i = 0; i < N; ++i) for (int j = 0; j < 3; ++j) sim.w[3 * i + j] = sim.w[3 * i + j] * sim.sigma[i] * sqrt_temperature;
  }
  for (int stage = 0; stage < 2; ++stage) {
    compute_fields(sim);  // :66 / :106
    if (thermal) {  // :68-74 / :108-114
#pragma omp parallel for
      for (int i = 0; i < N; ++i) for (int j = 0; j < 3; ++j) sim.h[3 * i + j] = (sim.w[3 * i + j] + sim.h[3 * i + j] / sim.mus[i]);
    } else {  // :75-82 / :115-122
#pragma omp parallel for
      for (int i = 0; i < N; ++i) for (int j = 0; j < 3; ++j) sim.h[3 * i + j] = sim.h[3 * i + j] / sim.mus[i];
    }
    if (stage == 0) {  // :84-101
#pragma omp parallel for
      for (int i = 0; i < N; ++i) {
        V3 spin = {sim.s[3 * i], sim.s[3 * i + 1], sim.s[3 * i + 2]};
        V3 field = {sim.h[3 * i], sim.h[3 * i + 1], sim.h[3 * i + 2]};
        V3 rhs = llg_rhs(spin, field, sim.gyro[i], sim.alpha[i]);
        for (int j = 0; j < 3; ++j) sim.ds_dt[3 * i + j] = 0.5 * rhs[j];
        spin = unit_vector(V3{spin[0] + sim.dt * rhs[0], spin[1] + sim.dt * rhs[1], spin[2] + sim.dt * rhs[2]});
        for (int j = 0; j < 3; ++j) sim.s[3 * i + j] = spin[j];
      }
      sim.time = t0 + sim.dt;  // :103-104
    } else {  // :124-144
#pragma omp parallel for
      for (int i = 0; i < N; ++i) {
        V3 spin = {sim.s[3 * i], sim.s[3 * i + 1], sim.s[3 * i + 2]};
        V3 spin_old = {sim.s_old[3 * i], sim.s_old[3 * i + 1], sim.s_old[3 * i + 2]};
        V3 field = {sim.h[3 * i], sim.h[3 * i + 1], sim.h[3 * i + 2]};
        V3 rhs = llg_rhs(spin, field, sim.gyro[i], sim.alpha[i]);
        for (int j = 0; j < 3; ++j) sim.ds_dt[3 * i + j] = sim.ds_dt[3 * i + j] + 0.5 * rhs[j];
        V3 ds = {sim.ds_dt[3 * i], sim.ds_dt[3 * i + 1], sim.ds_dt[3 * i + 2]};
        spin = unit_vector(V3{spin_old[0] + sim.dt * ds[0], spin_old[1] + sim.dt * ds[1], spin_old[2] + sim.dt * ds[2]});
        for (int j = 0; j < 3; ++j) sim.s[3 * i + j] = spin[j];
      }
    }
  }
  sim.iteration++;  // :146-147
  sim.time = sim.iteration * sim.dt;
}

// CudaRK4BaseSolver::run (solvers/cuda_rk4_base.cu:50-108) with CUDALLGRK4Solver::function_kernel
// (solvers/cuda_llg_rk4.cu:17-29, cuda_llg_rk4_kernel.cuh:11-58) and post_step = normalise_spins_cuda
// (cuda/cuda_spin_ops.cu:4-17).  The reference has no CPU RK4 solver: this is the CUDA solver's arithmetic on the
// host.  One noise draw per step (update_thermostat, :65) serves all four stages; the intermediate states
// s_old + a k are NOT normalised; the fields of k2/k3 are evaluated at t0 + dt/2, those of k4 at t0 + dt.
// Deviation: cuda_normalise_spins_kernel has no zero-length guard (a vacancy would become NaN); here, as in
// Vec3 unit_vector (containers/vec3.h:276-283), vectors of length <= DBL_EPSILON are left unchanged.
void rk4_function(Sim &sim, std::vector<double> &k) {
  const int N = sim.N;
  compute_fields(sim);   // cuda_llg_rk4.cu:18
#pragma omp parallel for
  for (int i = 0; i < N; ++i) {   // cuda_llg_rk4_kernel.cuh:25-57
    V3 h, s;
    for (int n = 0; n < 3; ++n) h[n] = ((sim.h[3 * i + n] / sim.mus[i]) + sim.w[3 * i + n]);
    for (int n = 0; n < 3; ++n) s[n] = sim.s[3 * i + n];
    const V3 sxh = {(s[1] * h[2] - s[2] * h[1]), (s[2] * h[0] - s[0] * h[2]), (s[0] * h[1] - s[1] * h[0])};
    const V3 sxsxh = {(s[1] * sxh[2] - s[2] * sxh[1]), (s[2] * sxh[0] - s[0] * sxh[2]), (s[0] * sxh[1] - s[1] * sxh[0])};
    for (int n = 0; n < 3; ++n) k[3 * i + n] = -sim.gyro[i] * (sxh[n] + sim.alpha[i] * sxsxh[n]);
  }
}

void rk4_run(Sim &sim, const double *normals) {
  const int N = sim.N;
  const double t0 = sim.time;
  sim.s_old = sim.s;   // :54-58
  // update_thermostat (:65): CudaThermostatClassical::update (thermostats/cuda_thermostat_classical.cc:47-56)
  if (sim.temperature > 0.0) {
    if (normals) std::copy(normals, normals + 3 * N, sim.w.begin());
    else { std::normal_distribution<> nd; for (auto &x : sim.w) x = nd(sim.rng); }
    const double sqrt_temperature = sqrt(sim.temperature);
    for (int i = 0; i < N; ++i) for (int j = 0; j < 3; ++j) sim.w[3 * i + j] = sim.w[3 * i + j] * sim.sigma[i] * sqrt_temperature;
  } else {
    std::fill(sim.w.begin(), sim.w.end(), 0.0);
  }
  std::vector<double> k1(3 * N), k2(3 * N), k3(3 * N), k4(3 * N);
  auto axpy_from_old = [&](double a, const std::vector<double> &k) {   // cublasDcopy + cublasDaxpy (:73-74, :81-82, :89-90)
    for (int n = 0; n < 3 * N; ++n) sim.s[n] = sim.s_old[n] + a * k[n];
  };
  rk4_function(sim, k1);                         // :68
  sim.time = t0 + 0.5 * sim.dt;                  // :70-71
  axpy_from_old(0.5 * sim.dt, k1);
  rk4_function(sim, k2);                         // :76
  sim.time = t0 + 0.5 * sim.dt;                  // :78-79
  axpy_from_old(0.5 * sim.dt, k2);
  rk4_function(sim, k3);                         // :84
  sim.time = t0 + sim.dt;                        // :86-87
  axpy_from_old(sim.dt, k3);
  rk4_function(sim, k4);                         // :92
  for (int i = 0; i < N; ++i) {
    V3 v;
    for (int n = 0; n < 3; ++n) {                // cuda_rk4_base_kernel.cuh:16: s_old + dt * (k1 + 2 k2 + 2 k3 + k4) / 6.0
      const int q = 3 * i + n;
      v[n] = sim.s_old[q] + sim.dt * (k1[q] + 2 * k2[q] + 2 * k3[q] + k4[q]) / 6.0;
    }
    const double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];   // cuda_spin_ops.cu:9-15: s * rsqrt(s.s)
    const double r = (n2 > DBL_EPSILON * DBL_EPSILON) ? 1.0 / sqrt(n2) : 1.0;
    for (int n = 0; n < 3; ++n) sim.s[3 * i + n] = v[n] * r;
  }
  sim.iteration++;                               // :105-106
  sim.time = sim.iteration * sim.dt;
}

}  // namespace

extern "C" {

void *jo_sim_create(int num_spins, const double *mus, const double *gyro, const double *alpha) {
  auto *sim = new Sim;
  sim->N = num_spins;
  sim->s.assign(3 * num_spins, 0.0); sim->h.assign(3 * num_spins, 0.0); sim->ds_dt.assign(3 * num_spins, 0.0);
  sim->mus.assign(mus, mus + num_spins); sim->gyro.assign(gyro, gyro + num_spins); sim->alpha.assign(alpha, alpha + num_spins);
  return sim;
}
void jo_sim_destroy(void *p) { delete static_cast<Sim *>(p); }

// SparseInteractionHamiltonian::insert_interaction_tensor (sparse_interaction.cc:25-34) for every pair,
// then SparseMatrix::Builder sort (sparse_matrix_builder.h:108-143), merge (:145-164, including its
// skip-after-erase behaviour), optional is_symmetric (:320-362) and build_csr (:186-248).
int jo_sim_add_exchange(void *p, int64_t n_pairs, const int *i, const int *j, const double *J9, int check_symmetric) {
  auto *sim = static_cast<Sim *>(p);
  return guarded([&]() {
    const int num_rows = 3 * sim->N;
    std::vector<int> row_, col_;
    std::vector<double> val_;
    for (int64_t q = 0; q < n_pairs; ++q) {
      for (int m = 0; m < 3; ++m) for (int n = 0; n < 3; ++n) {
        const double value = J9[9 * q + 3 * m + n];
        if (value != 0.0) {
          const int r = 3 * i[q] + m, c = 3 * j[q] + n;
          if (r >= num_rows || r < 0 || c >= num_rows || c < 0) throw std::runtime_error("Invalid index for sparse matrix");
          if (val_.size() >= std::size_t(std::numeric_limits<int>::max() - 1))
            throw std::runtime_error("Number of non zero elements is too large for the sparse matrix index_type");
          row_.push_back(r); col_.push_back(c); val_.push_back(value);
        }
      }
    }
    std::vector<std::size_t> permutation(col_.size());
    std::iota(permutation.begin(), permutation.end(), 0);
    std::sort(permutation.begin(), permutation.end(), [&](std::size_t a, std::size_t b) {
      if (row_[a] < row_[b]) return true;
      if (row_[a] == row_[b]) return col_[a] < col_[b];
      return false; });
    auto apply = [&](auto &v) { auto tmp = v; for (std::size_t n = 0; n < permutation.size(); ++n) tmp[n] = v[permutation[n]]; v.swap(tmp); };
    apply(row_); apply(col_); apply(val_);
    for (std::size_t m = 1; m < row_.size(); ++m) {
      if (row_[m] == row_[m - 1] && col_[m] == col_[m - 1]) {
        val_[m - 1] += val_[m];
        val_.erase(val_.begin() + m); row_.erase(row_.begin() + m); col_.erase(col_.begin() + m);
      }
    }
    if (check_symmetric) {
      for (std::size_t n = 0; n < row_.size(); ++n) {
        const int ii = row_[n], jj = col_[n];
        auto lo = std::lower_bound(row_.cbegin(), row_.cend(), jj);
        if (lo == row_.cend() || *lo != jj) throw std::runtime_error("sparse matrix for exchange is not symmetric");
        auto hi = std::upper_bound(lo, row_.cend(), jj);
        auto cb = col_.cbegin() + (lo - row_.cbegin()), ce = col_.cbegin() + (hi - row_.cbegin());
        auto pos = std::lower_bound(cb, ce, ii);
        if (pos == ce || *pos != ii) throw std::runtime_error("sparse matrix for exchange is not symmetric");
        if (val_[pos - col_.cbegin()] != val_[n]) throw std::runtime_error("sparse matrix for exchange is not symmetric");
      }
    }
    Term t; t.kind = Term::EXCHANGE;
    t.field.assign(3 * sim->N, 0.0);
    const std::size_t nnz = val_.size();
    t.row.assign(num_rows + 1, 0);
    int previous_row = 0;
    for (std::size_t m = 1; m < row_.size(); ++m) {
      const int current_row = row_[m];
      if (current_row == previous_row) continue;
      for (int r = previous_row + 1; r < current_row + 1; ++r) t.row[r] = int(m);
      previous_row = current_row;
    }
    for (int r = previous_row + 1; r < num_rows + 1; ++r) t.row[r] = int(nnz);
    t.col = col_; t.val = val_;
    sim->terms.push_back(std::move(t));
  });
}

// CudaBiquadraticExchangeHamiltonian ctor (hamiltonian/cuda_biquadratic_exchange.cu:127-156): insert(i, j, value) for every pair of the
// neighbour list (the caller applies the value > energy_cutoff filter of :131), Builder sort / merge / is_symmetric / build_csr as for
// the exchange matrix, here N x N with one scalar per pair
int jo_sim_add_biquadratic(void *p, int64_t n_pairs, const int *i, const int *j, const double *B, int check_symmetric) {
  auto *sim = static_cast<Sim *>(p);
  return guarded([&]() {
    const int num_rows = sim->N;
    std::vector<int> row_(i, i + n_pairs), col_(j, j + n_pairs);
    std::vector<double> val_(B, B + n_pairs);
    for (int64_t q = 0; q < n_pairs; ++q)
      if (row_[q] >= num_rows || row_[q] < 0 || col_[q] >= num_rows || col_[q] < 0) throw std::runtime_error("Invalid index for sparse matrix");
    std::vector<std::size_t> permutation(col_.size());
    std::iota(permutation.begin(), permutation.end(), 0);
    std::sort(permutation.begin(), permutation.end(), [&](std::size_t a, std::size_t b) {
      if (row_[a] < row_[b]) return true;
      if (row_[a] == row_[b]) return col_[a] < col_[b];
      return false; });
    auto apply = [&](auto &v) { auto tmp = v; for (std::size_t n = 0; n < permutation.size(); ++n) tmp[n] = v[permutation[n]]; v.swap(tmp); };
    apply(row_); apply(col_); apply(val_);
    for (std::size_t m = 1; m < row_.size(); ++m) {
      if (row_[m] == row_[m - 1] && col_[m] == col_[m - 1]) {
        val_[m - 1] += val_[m];
        val_.erase(val_.begin() + m); row_.erase(row_.begin() + m); col_.erase(col_.begin() + m);
      }
    }
    if (check_symmetric) {
      for (std::size_t n = 0; n < row_.size(); ++n) {
        const int ii = row_[n], jj = col_[n];
        auto lo = std::lower_bound(row_.cbegin(), row_.cend(), jj);
        if (lo == row_.cend() || *lo != jj) throw std::runtime_error("sparse matrix for biquadratic-exchange is not symmetric");
        auto hi = std::upper_bound(lo, row_.cend(), jj);
        auto cb = col_.cbegin() + (lo - row_.cbegin()), ce = col_.cbegin() + (hi - row_.cbegin());
        auto pos = std::lower_bound(cb, ce, ii);
        if (pos == ce || *pos != ii || val_[pos - col_.cbegin()] != val_[n]) throw std::runtime_error("sparse matrix for biquadratic-exchange is not symmetric");
      }
    }
    Term t; t.kind = Term::BIQUADRATIC;
    t.field.assign(3 * sim->N, 0.0);
    t.row.assign(num_rows + 1, 0);
    for (std::size_t m = 0; m < row_.size(); ++m) t.row[row_[m] + 1]++;
    for (int r = 0; r < num_rows; ++r) t.row[r + 1] += t.row[r];
    t.col = col_; t.val = val_;
    sim->terms.push_back(std::move(t));
  });
}

int jo_sim_add_uniaxial(void *p, int power, const double *magnitude, const double *axis) {
  auto *sim = static_cast<Sim *>(p);
  Term t; t.kind = Term::UNIAXIAL; t.power = power;
  t.field.assign(3 * sim->N, 0.0);
  t.magnitude.assign(magnitude, magnitude + sim->N);
  t.axis.assign(axis, axis + 3 * sim->N);
  sim->terms.push_back(std::move(t));
  return 0;
}

int jo_sim_add_zeeman(void *p, const double *dc, const double *ac, const double *omega) {
  auto *sim = static_cast<Sim *>(p);
  Term t; t.kind = Term::ZEEMAN;
  t.field.assign(3 * sim->N, 0.0);
  t.dc.assign(dc, dc + 3 * sim->N);
  if (ac && omega) { t.has_ac = true; t.ac.assign(ac, ac + 3 * sim->N); t.omega.assign(omega, omega + sim->N); }
  sim->terms.push_back(std::move(t));
  return 0;
}

int jo_sim_add_applied_field(void *p, const double *B, int type, double t0_ps, double fbw_THz, double fc_THz) {
  auto *sim = static_cast<Sim *>(p);
  if (type < 0 || type > 2) return 1;
  Term t; t.kind = Term::APPLIED;
  t.field.assign(3 * sim->N, 0.0);
  for (int j = 0; j < 3; ++j) t.B[j] = B[j];
  t.pulse_type = type; t.t0 = t0_ps; t.fbw = fbw_THz; t.fc = fc_THz;
  sim->terms.push_back(std::move(t));
  return 0;
}

int64_t jo_sim_exchange_nnz(void *p, int term) { return int64_t(static_cast<Sim *>(p)->terms.at(term).val.size()); }
void jo_sim_exchange_csr(void *p, int term, int *row, int *col, double *val) {
  auto &t = static_cast<Sim *>(p)->terms.at(term);
  std::copy(t.row.begin(), t.row.end(), row);
  std::copy(t.col.begin(), t.col.end(), col);
  std::copy(t.val.begin(), t.val.end(), val);
}

void jo_sim_set_spins(void *p, const double *s_aos) { auto *sim = static_cast<Sim *>(p); std::copy(s_aos, s_aos + 3 * sim->N, sim->s.begin()); }
void jo_sim_get_spins(void *p, double *s_aos) { auto *sim = static_cast<Sim *>(p); std::copy(sim->s.begin(), sim->s.end(), s_aos); }
void jo_sim_get_h(void *p, double *h_aos) { auto *sim = static_cast<Sim *>(p); std::copy(sim->h.begin(), sim->h.end(), h_aos); }

// HeunLLGSolver::initialize (solvers/cpu_llg_heun.cc:15-43)
void jo_sim_init_solver(void *p, double step_size_ps, int use_gilbert_prefactor, uint64_t seed) {
  auto *sim = static_cast<Sim *>(p);
  sim->dt = step_size_ps; sim->time = 0.0; sim->iteration = 0;
  sim->s_old.assign(3 * sim->N, 0.0); sim->w.assign(3 * sim->N, 0.0); sim->sigma.assign(sim->N, 0.0);
  for (int i = 0; i < sim->N; ++i) {
    double denominator = 1.0;
    if (use_gilbert_prefactor) denominator = 1.0 + sim->alpha[i] * sim->alpha[i];
    sim->sigma[i] = sqrt((2.0 * kBoltzmannIU * sim->alpha[i]) / (sim->mus[i] * sim->gyro[i] * sim->dt * denominator));
  }
  sim->rng.seed(seed);
}
void jo_sim_get_sigma(void *p, double *sigma) { auto *sim = static_cast<Sim *>(p); std::copy(sim->sigma.begin(), sim->sigma.end(), sigma); }
void jo_sim_set_temperature(void *p, double T) { static_cast<Sim *>(p)->temperature = T; }
double jo_sim_time(void *p) { return static_cast<Sim *>(p)->time; }

void jo_sim_run_rk4(void *p, int nsteps, const double *normals) {
  auto *sim = static_cast<Sim *>(p);
  for (int n = 0; n < nsteps; ++n) rk4_run(*sim, normals ? normals + std::size_t(n) * 3 * sim->N : nullptr);
}

void jo_sim_run(void *p, int nsteps, const double *normals) {
  auto *sim = static_cast<Sim *>(p);
  for (int n = 0; n < nsteps; ++n) heun_run(*sim, normals ? normals + std::size_t(n) * 3 * sim->N : nullptr);
}

void jo_sim_term_fields(void *p, int term, double time, double *field_aos) {
  auto *sim = static_cast<Sim *>(p);
  auto &t = sim->terms.at(term);
  calculate_fields(*sim, t, time);
  std::copy(t.field.begin(), t.field.end(), field_aos);
}

// calculate_total_energy: exchange sparse_interaction.cc:86-100; uniaxial uniaxial_anisotropy.cc:118-133;
// Zeeman zeeman.cc:74-87,103-119
double jo_sim_term_total_energy(void *p, int term, double time) {
  auto *sim = static_cast<Sim *>(p);
  auto &t = sim->terms.at(term);
  const int N = sim->N;
  double e_total = 0.0;
  switch (t.kind) {
    case Term::EXCHANGE: {
      calculate_fields(*sim, t, time);
      double total_energy = 0.0;
      for (int i = 0; i < N; ++i) {
        V3 s_i = {sim->s[3 * i], sim->s[3 * i + 1], sim->s[3 * i + 2]};
        V3 h_i = {t.field[3 * i], t.field[3 * i + 1], t.field[3 * i + 2]};
        total_energy += -dot(s_i, h_i);
      }
      return 0.5 * total_energy;
    }
    case Term::UNIAXIAL:
      for (int i = 0; i < N; ++i) {
        double d = (t.axis[3 * i] * sim->s[3 * i] + t.axis[3 * i + 1] * sim->s[3 * i + 1] + t.axis[3 * i + 2] * sim->s[3 * i + 2]);
        e_total += (-t.magnitude[i] * pow(d, t.power));
      }
      return e_total;
    case Term::ZEEMAN:
      for (int i = 0; i < N; ++i) {
        V3 s_i = {sim->s[3 * i], sim->s[3 * i + 1], sim->s[3 * i + 2]};
        V3 field = {t.dc[3 * i], t.dc[3 * i + 1], t.dc[3 * i + 2]};
        if (t.has_ac) for (int j = 0; j < 3; ++j) field[j] += t.ac[3 * i + j] * cos(t.omega[i] * time);
        e_total += -dot(s_i, field);
      }
      return e_total;
    case Term::APPLIED: {  // applied_field.cc:121-128,150-155
      const V3 b = applied_b(t, time);
      for (int i = 0; i < N; ++i) {
        const V3 field = {sim->mus[i] * b[0], sim->mus[i] * b[1], sim->mus[i] * b[2]};
        e_total += -(sim->s[3 * i] * field[0] + sim->s[3 * i + 1] * field[1] + sim->s[3 * i + 2] * field[2]);
      }
      return e_total;
    }
    case Term::BIQUADRATIC: {  // cuda_biquadratic_exchange.cu:185-201: total += -dot(s_i, 0.5 * h_i); return 0.5 * total
      calculate_fields(*sim, t, time);
      double total_energy = 0.0;
      for (int i = 0; i < N; ++i) {
        V3 s_i = {sim->s[3 * i], sim->s[3 * i + 1], sim->s[3 * i + 2]};
        V3 h_i = {0.5 * t.field[3 * i], 0.5 * t.field[3 * i + 1], 0.5 * t.field[3 * i + 2]};
        total_energy += -dot(s_i, h_i);
      }
      return 0.5 * total_energy;
    }
  }
  return 0.0;
}

// per-spin energies (Hamiltonian::calculate_energies): exchange sparse_interaction.cc:60-84
// (e_i = -s_i . (A s)_i, no factor 1/2), uniaxial :126-133, Zeeman :82-87
void jo_sim_term_energies(void *p, int term, double time, double *e) {
  auto *sim = static_cast<Sim *>(p);
  auto &t = sim->terms.at(term);
  const int N = sim->N;
  if (t.kind == Term::EXCHANGE || t.kind == Term::BIQUADRATIC) calculate_fields(*sim, t, time);
  for (int i = 0; i < N; ++i) {
    V3 s_i = {sim->s[3 * i], sim->s[3 * i + 1], sim->s[3 * i + 2]};
    switch (t.kind) {
      case Term::BIQUADRATIC:   // calculate_energy (cuda_biquadratic_exchange.cu:235-240): -0.5 * dot(s_i, field)
        e[i] = -0.5 * dot(s_i, V3{t.field[3 * i], t.field[3 * i + 1], t.field[3 * i + 2]});
        break;
      case Term::EXCHANGE: e[i] = -dot(s_i, V3{t.field[3 * i], t.field[3 * i + 1], t.field[3 * i + 2]}); break;
      case Term::UNIAXIAL: {
        double d = (t.axis[3 * i] * s_i[0] + t.axis[3 * i + 1] * s_i[1] + t.axis[3 * i + 2] * s_i[2]);
        e[i] = 0.0 + (-t.magnitude[i] * pow(d, t.power));
        break;
      }
      case Term::ZEEMAN: {
        V3 field = {t.dc[3 * i], t.dc[3 * i + 1], t.dc[3 * i + 2]};
        if (t.has_ac) for (int j = 0; j < 3; ++j) field[j] += t.ac[3 * i + j] * cos(t.omega[i] * time);
        e[i] = -dot(s_i, field);
        break;
      }
      case Term::APPLIED: {
        const V3 b = applied_b(t, time);
        e[i] = -(s_i[0] * (sim->mus[i] * b[0]) + s_i[1] * (sim->mus[i] * b[1]) + s_i[2] * (sim->mus[i] * b[2]));
        break;
      }
    }
  }
}

// MagnetisationMonitor::update core (monitors/magnetisation.cc:88-100): per group
//   mag = sum_spins_moments (helpers/spinops.cc:55-67, plain sum in index order)
//   mu_total = scalar_field_indexed_reduce (helpers/array_ops.h:54-68, Kahan)
// out4[g] = {mag_x, mag_y, mag_z, mu_total}; group_of_spin[i] in [0, n_groups)
void jo_magnetisation(int64_t N, const double *s_aos, const double *mus, int n_groups, const int *group_of_spin, double *out4) {
  for (int g = 0; g < n_groups; ++g) {
    V3 sum = {0.0, 0.0, 0.0};
    double ksum = 0.0, c = 0.0;
    bool first = true;
    for (int64_t i = 0; i < N; ++i) {
      if (group_of_spin[i] != g) continue;
      for (int j = 0; j < 3; ++j) sum[j] += mus[i] * s_aos[3 * i + j];
      if (first) { ksum = mus[i]; first = false; }
      else { double y = mus[i] - c; double t = ksum + y; c = (t - ksum) - y; ksum = t; }
    }
    out4[4 * g] = sum[0]; out4[4 * g + 1] = sum[1]; out4[4 * g + 2] = sum[2]; out4[4 * g + 3] = ksum;
  }
}

// SpinTemperatureMonitor::update (monitors/spin_temperature.cc:22-37) on caller-supplied s and h
double jo_spin_temperature(int64_t N, const double *s_aos, const double *h_aos) {
  double sum_s_dot_h = 0.0, sum_s_cross_h = 0.0;
  for (int64_t i = 0; i < N; ++i) {
    V3 spin = {s_aos[3 * i], s_aos[3 * i + 1], s_aos[3 * i + 2]};
    V3 field = {h_aos[3 * i], h_aos[3 * i + 1], h_aos[3 * i + 2]};
    V3 c = cross(spin, field);
    sum_s_cross_h += dot(c, c);
    sum_s_dot_h += dot(spin, field);
  }
  return sum_s_cross_h / (2.0 * kBoltzmannIU * sum_s_dot_h);
}

}  // extern "C"
