// jams_host.cc — see jams_host.h.  Setup-time logic only (integers, small float geometry, file formats); the
// per-step numerics are in libjams_b200.so.  Citations are relative to /root/reference/src/jams/.
#include "jams_host.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <functional>
#include <iomanip>
#include <map>
#include <iostream>
#include <random>
#include <sstream>

namespace jams_b200 {

namespace {

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kTwoPi = 2.0 * kPi;
constexpr double kmRyd2meV = 13.605693123;
constexpr double kBoltzmannIU = 0.0861733326;
constexpr double kHBarIU = 0.6582119569;
constexpr double kElectronGFactor = 2.0023193043625;
constexpr double kGyroIU = kElectronGFactor * kBohrMagnetonIU / kHBarIU;   // helpers/consts.h:31

std::string lowercase(std::string s) { for (auto &c : s) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c))); return s; }

// ---- tolerant comparisons (helpers/maths.h:16-46) ----------------------------------------------------------
bool approximately_equal(double a, double b, double eps) {
  if (std::abs(a - b) <= eps) return true;
  return std::abs(a - b) <= std::max(std::abs(a), std::abs(b)) * eps;
}
bool approximately_zero(double a, double eps) { return std::abs(a) <= eps; }
bool definately_greater_than(double a, double b, double eps) { return (a - b) > std::max(std::abs(a), std::abs(b)) * eps; }
bool definately_less_than(double a, double b, double eps) { return (b - a) > std::max(std::abs(a), std::abs(b)) * eps; }
bool vec_approximately_equal(const Vec3 &a, const Vec3 &b, double eps) {
  return approximately_equal(a[0], b[0], eps) && approximately_equal(a[1], b[1], eps) && approximately_equal(a[2], b[2], eps);
}

Vec3 matvec(const Mat3 &A, const Vec3 &v) {
  return {{A[0][0] * v[0] + A[0][1] * v[1] + A[0][2] * v[2], A[1][0] * v[0] + A[1][1] * v[1] + A[1][2] * v[2],
           A[2][0] * v[0] + A[2][1] * v[1] + A[2][2] * v[2]}};
}
double determinant(const Mat3 &A) {
  return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) + A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
}
Mat3 matmul(const Mat3 &A, const Mat3 &B) {
  Mat3 C{};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i][j] = A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j];
  return C;
}
double dot(const Vec3 &a, const Vec3 &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
Vec3 cross(const Vec3 &a, const Vec3 &b) { return {{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]}}; }
double norm(const Vec3 &a) { return std::sqrt(dot(a, a)); }
Vec3 unit_vector(const Vec3 &a) {   // containers/vec3.h:276-283
  const double n = norm(a);
  if (n <= DBL_EPSILON) return a;
  return {{a[0] / n, a[1] / n, a[2] / n}};
}
Mat3 identity() { return {{{{1, 0, 0}}, {{0, 1, 0}}, {{0, 0, 1}}}}; }

Mat3 inverse(const Mat3 &m) {   // 3x3 cofactor inverse (the reference uses LAPACK dgetri, containers/mat3.h:233-262; results are
                                // only used through 1e-4 tolerance snaps)
  const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                     m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  if (det == 0.0) throw std::runtime_error("unit cell matrix is singular");
  Mat3 r{};
  r[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) / det; r[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) / det; r[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / det;
  r[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) / det; r[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) / det; r[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) / det;
  r[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) / det; r[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) / det; r[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) / det;
  return r;
}

Vec3 read_vec3(const Setting &s) {
  if (s.length() != 3) throw ConfigError("setting '" + s.name() + "' must have 3 components");
  return {{s[0].as_double(), s[1].as_double(), s[2].as_double()}};
}

Vec3 normalise_fractional_coordinate(Vec3 r) {   // core/lattice.cc:48-64
  for (int n = 0; n < 3; ++n) {
    if (r[n] < 0.0) r[n] = r[n] + 1.0;
    if (approximately_equal(r[n], 1.0, kLatticeTolerance)) r[n] = 0.0;
  }
  return r;
}

Vec3 lattice_translation_vector(const Vec3 &q, double tolerance) {   // core/interactions.cc:58-76
  Vec3 T;
  for (int n = 0; n < 3; ++n) {
    const double nearest = std::nearbyint(q[n]);
    T[n] = approximately_zero(q[n] - nearest, tolerance) ? nearest : std::floor(q[n]);
  }
  return T;
}

Mat3 rotation_matrix_between_vectors(const Vec3 &a, const Vec3 &b) {   // containers/mat3.h:334-366
  auto ssc = [](const Vec3 &v) -> Mat3 { return {{{{0, -v[2], v[1]}}, {{v[2], 0, -v[0]}}, {{-v[1], v[0], 0}}}}; };
  const Vec3 ua = unit_vector(a), ub = unit_vector(b);
  const double c = dot(ua, ub);
  Mat3 R = identity();
  if (approximately_equal(c, 1.0, 1e-12)) return R;
  if (approximately_equal(c, -1.0, 1e-12)) {
    const Vec3 ortho = std::abs(ua[0]) < 0.9 ? Vec3{{1, 0, 0}} : Vec3{{0, 1, 0}};
    const Mat3 vx = ssc(unit_vector(unit_vector(cross(ua, ortho))));
    Mat3 k1 = vx;
    for (auto &row : k1) for (auto &x : row) x *= (1.0 - std::cos(kPi));
    const Mat3 vx2 = matmul(k1, vx);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i][j] += std::sin(kPi) * vx[i][j] + vx2[i][j];
    return R;
  }
  const Vec3 v = cross(ua, ub);
  const double s = norm(v);
  const Mat3 vx = ssc(v);
  Mat3 kvx = vx;
  const double k = (1.0 - c) / (s * s);
  for (auto &row : kvx) for (auto &x : row) x *= k;
  const Mat3 vx2 = matmul(kvx, vx);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[i][j] += vx[i][j] + vx2[i][j];
  return R;
}

// The rotations R of the crystal's symmetry operations {R|t} with t = 0, in the basis of the given cell: what the reference takes
// from spglib's dataset (core/lattice.cc:783-822) and filters in lattice_site_point_group_symops (:1127-1153).  R runs over the
// integer matrices that preserve the cell's metric (candidate columns in [-2, 2]^3, the lattice holohedry: at most 48); R is kept
// if the motif maps onto itself, type by type, without any translation (cartesian tolerance symprec in lattice parameters).
// Same enumeration order as jams_b200.lattice.find_space_group_operations.
std::vector<Mat3> find_point_operations(const Mat3 &cell, const std::vector<Vec3> &motif_frac, const std::vector<int> &types, double symprec) {
  double G[3][3], lens[3], gmax = 0.0;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    G[i][j] = 0.0;
    for (int k = 0; k < 3; ++k) G[i][j] += cell[k][i] * cell[k][j];
    gmax = std::max(gmax, G[i][j]);
  }
  for (int k = 0; k < 3; ++k) lens[k] = std::sqrt(G[k][k]);
  const double tol = symprec * std::max(1.0, std::sqrt(gmax));
  const double lmax = std::max(lens[0], std::max(lens[1], lens[2]));
  std::vector<Vec3> cols[3];
  for (int a = -2; a <= 2; ++a) for (int b = -2; b <= 2; ++b) for (int c = -2; c <= 2; ++c) {
    const Vec3 v{{double(a), double(b), double(c)}};
    const double len = norm(matvec(cell, v));
    for (int k = 0; k < 3; ++k) if (std::abs(len - lens[k]) <= tol) cols[k].push_back(v);
  }
  auto dotc = [&](const Vec3 &u, const Vec3 &v) { const Vec3 a = matvec(cell, u), b = matvec(cell, v); return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  std::vector<Mat3> out;
  const size_t n = motif_frac.size();
  for (const Vec3 &c0 : cols[0]) for (const Vec3 &c1 : cols[1]) {
    if (std::abs(dotc(c0, c1) - G[0][1]) > tol * lmax) continue;
    for (const Vec3 &c2 : cols[2]) {
      if (std::abs(dotc(c0, c2) - G[0][2]) > tol * lmax || std::abs(dotc(c1, c2) - G[1][2]) > tol * lmax) continue;
      Mat3 R;
      for (int r = 0; r < 3; ++r) { R[r][0] = c0[r]; R[r][1] = c1[r]; R[r][2] = c2[r]; }
      const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                         R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
      if (std::abs(std::abs(det) - 1.0) > 1e-9) continue;
      bool ok = true;
      for (size_t a = 0; a < n && ok; ++a) {
        const Vec3 q = matvec(R, motif_frac[a]);
        bool found = false;
        for (size_t b = 0; b < n && !found; ++b) {
          if (types[b] != types[a]) continue;
          Vec3 d;
          for (int k = 0; k < 3; ++k) { d[k] = q[k] - motif_frac[b][k]; d[k] -= std::nearbyint(d[k]); }
          if (norm(matvec(cell, d)) <= symprec) found = true;
        }
        ok = found;
      }
      if (ok) out.push_back(R);
    }
  }
  return out;
}

// jams::fmt::sci / decimal (helpers/output.h:32-42)
std::ostream &fmt_sci(std::ostream &os) { return os << std::setprecision(8) << std::setw(16) << std::scientific << std::right; }
std::ostream &fmt_decimal(std::ostream &os) { return os << std::setprecision(6) << std::setw(16) << std::fixed << std::right; }

}  // namespace

double energy_unit_conversion(const std::string &name) {
  static const std::map<std::string, double> table = {
      {"joules", kJoule2meV}, {"J", kJoule2meV}, {"milli_electron_volts", 1.0}, {"meV", 1.0}, {"milli_rydbergs", kmRyd2meV},
      {"mRyd", kmRyd2meV}, {"rydbergs", kmRyd2meV * 1e3}, {"Ryd", kmRyd2meV * 1e3}, {"Kelvin", kBoltzmannIU}, {"K", kBoltzmannIU}};
  auto it = table.find(name);
  if (it == table.end()) throw std::runtime_error("energy units: " + name + " is not known");
  return it->second;
}

// =====================================================================================================
// Lattice
// =====================================================================================================
Lattice::Lattice(const Setting &config) {
  // materials (core/lattice.cc:337-352, containers/material.h:28-60)
  const Setting &mats = config.required("materials");
  for (int i = 0; i < mats.length(); ++i) {
    const Setting &cfg = mats[i];
    Material m;
    m.name = cfg.required("name").as_string();
    m.moment = cfg.required("moment").as_double() * kBohrMagnetonIU;
    m.gyro = cfg.get("gyro", 1.0) * kGyroIU;
    m.alpha = cfg.get("alpha", 0.01);
    if (const Setting *sp = cfg.find("spin")) {
      if (sp->is_array()) {
        if (sp->length() == 3) m.spin = read_vec3(*sp);
        else if (sp->length() == 2) {
          const double theta = (*sp)[0].as_double() * kPi / 180.0, phi = (*sp)[1].as_double() * kPi / 180.0;
          m.spin = {{std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta)}};
        } else throw std::runtime_error("spin setting array is not length 2 or 3");
      } else if (sp->is_string() && lowercase(sp->as_string()) == "random") {
        m.randomize = true;
      }
    }
    if (material_exists(m.name)) throw std::runtime_error("the material " + m.name + " is specified twice in the configuration");
    materials.push_back(m);
  }

  // unit cell (core/lattice.cc:354-410)
  const Setting &uc = config.required("unitcell");
  const Setting &basis = uc.required("basis");
  if (basis.length() != 3) throw ConfigError("unitcell.basis must be a 3x3 matrix");
  for (int r = 0; r < 3; ++r) { const Vec3 row = read_vec3(basis[r]); for (int c = 0; c < 3; ++c) cell[r][c] = row[c]; }
  lattice_parameter = uc.required("parameter").as_double();
  if (lattice_parameter < 0.0) throw std::runtime_error("lattice parameter cannot be negative");
  if (lattice_parameter == 0.0) throw std::runtime_error("lattice parameter cannot be zero");
  cell_inv = inverse(cell);

  // lattice (core/lattice.cc:411-427)
  const Setting &lat = config.required("lattice");
  const Setting &size = lat.required("size");
  if (size.length() != 3) throw ConfigError("lattice.size must have 3 components");
  for (int n = 0; n < 3; ++n) dims[n] = static_cast<int>(size[n].as_int());
  if (const Setting *p = lat.find("periodic")) for (int n = 0; n < 3; ++n) periodic[n] = (*p)[n].as_bool();
  if (const Setting *sf = lat.find("spins")) spins_file = sf->as_string();
  const Setting *impurity_settings = lat.find("impurities");
  impurities_seed = static_cast<uint64_t>(lat.get("impurities_seed", 0));   // the reference draws a seed from its global generator if absent
  {   // "Rotating the system" (core/lattice.cc:434-454,515-575): orientation first, then global_rotation; a_k <- R a_k
    auto rotate_cell = [&](const Mat3 &R) {
      const double before = std::abs(determinant(cell));
      cell = matmul(R, cell);   // columns are a, b, c (containers/cell.cc:62-68)
      if (std::abs(std::abs(determinant(cell)) - before) > std::max(before, 1.0) * kLatticeTolerance * kLatticeTolerance * kLatticeTolerance)
        throw std::runtime_error("unitcell volume has changed after rotation");
      cell_inv = inverse(cell);
    };
    if (const Setting *axis = lat.find("orientation_axis")) {
      const Setting *lv = lat.find("orientation_lattice_vector"), *cv = lat.find("orientation_cartesian_vector");
      if (lv && cv) throw std::runtime_error("Only one of 'orientation_lattice_vector' or 'orientation_cartesian_vector' can be defined");
      if (lv || cv) {
        const Vec3 vec = lv ? read_vec3(*lv) : matvec(cell_inv, read_vec3(*cv));
        const Vec3 cart = unit_vector(matvec(cell, vec));
        rotate_cell(rotation_matrix_between_vectors(cart, read_vec3(*axis)));
      }
    }
    if (const Setting *gr = lat.find("global_rotation")) {
      Mat3 R;
      for (int r = 0; r < 3; ++r) { const Vec3 row = read_vec3((*gr)[r]); for (int c2 = 0; c2 < 3; ++c2) R[r][c2] = row[c2]; }
      rotate_cell(R);
    }
  }

  // motif (core/lattice.cc:286-310, 429-470)
  const std::string fmt_name = uc.get("coordinate_format", "FRACTIONAL");
  bool cartesian = false;
  { std::string up = fmt_name; for (auto &c : up) c = static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
    if (up == "CARTESIAN") cartesian = true; else if (up != "FRACTIONAL") throw std::runtime_error("Unknown coordinate format for atom positions in unit cell"); }
  const Setting &pos = uc.required("positions");
  if (!pos.is_list()) throw std::runtime_error("unitcell.positions must be a list (position files are not supported here)");
  for (int i = 0; i < pos.length(); ++i) {
    const std::string mat = pos[i][0].as_string();
    Vec3 p = read_vec3(pos[i][1]);
    if (cartesian) p = matvec(cell_inv, p);
    if (!material_exists(mat)) throw std::runtime_error("material " + mat + " in the motif is not defined in the configuration");
    motif_material.push_back(material_index(mat));
    motif_frac.push_back(normalise_fractional_coordinate(p));
  }
  M = static_cast<int>(motif_frac.size());
  if (M < 1) throw std::runtime_error("unit cell has no motif positions");
  const long long n = 1LL * dims[0] * dims[1] * dims[2] * M;
  if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1 || n >= (1LL << 31)) throw std::runtime_error("invalid lattice size");
  num_spins = static_cast<int>(n);

  if (impurity_settings) {   // read_impurities_from_config (core/lattice.cc:1077-1109) + the substitution loop of generate_supercell (:614-640)
    std::map<int, std::pair<int, double>> impurity_map;
    for (int q = 0; q < impurity_settings->length(); ++q) {
      const Setting &e = (*impurity_settings)[q];
      const std::string a = e[0].as_string(), b = e[1].as_string();
      if (!material_exists(a)) throw std::runtime_error("impurity " + std::to_string(q) + " materialA (" + a + ") does not exist");
      if (!material_exists(b)) throw std::runtime_error("impurity " + std::to_string(q) + " materialB (" + b + ") does not exist");
      const double fraction = e[2].as_double();
      if (fraction < 0.0 || fraction >= 1.0) throw std::runtime_error("impurity " + std::to_string(q) + " fraction must be 0 =< x < 1");
      if (!impurity_map.emplace(material_index(a), std::make_pair(material_index(b), fraction)).second)
        throw std::runtime_error("impurity " + std::to_string(q) + " redefines a previous impurity");
    }
    // pcg32 (setseq_xsh_rr_64_32, default stream) and libstdc++'s generate_canonical<double, 53> over a 32-bit generator, restated from
    // their published definitions (the reference fetches pcg at configure time): one draw per candidate site, in site order
    const uint64_t mult = 6364136223846793005ULL, inc = 1442695040888963407ULL;
    uint64_t state = (impurities_seed + inc) * mult + inc;
    auto next32 = [&]() -> uint32_t {
      const uint64_t old = state;
      state = old * mult + inc;
      const uint32_t x = static_cast<uint32_t>(((old >> 18u) ^ old) >> 27u), rot = static_cast<uint32_t>(old >> 59u);
      return (x >> rot) | (x << ((0u - rot) & 31u));
    };
    auto uniform = [&]() -> double {
      double sum = 0.0, tmp = 1.0;
      for (int k = 0; k < 2; ++k) { sum += static_cast<double>(next32()) * tmp; tmp *= 4294967296.0; }
      const double r = sum / tmp;
      return r < 1.0 ? r : std::nextafter(1.0, 0.0);
    };
    site_materials_.resize(num_spins);
    for (int i = 0; i < num_spins; ++i) {
      int material = motif_material[i % M];
      const auto it = impurity_map.find(material);
      if (it != impurity_map.end() && uniform() < it->second.second) material = it->second.first;
      site_materials_[i] = material;
    }
    has_impurities = !impurity_map.empty();
  }

  if (const Setting *solver = config.find("solver")) gilbert_prefactor = solver->get("gilbert_prefactor", false);   // core/lattice.cc:696-697
  {   // the reference asks spglib here (core/lattice.cc:783-822)
    std::vector<int> types(motif_material.begin(), motif_material.end());
    rotations = find_point_operations(cell, motif_frac, types, kLatticeTolerance);
  }
}

int Lattice::material_index(const std::string &name) const {
  for (size_t i = 0; i < materials.size(); ++i) if (materials[i].name == name) return static_cast<int>(i);
  throw std::runtime_error("material " + name + " does not exist");
}
bool Lattice::material_exists(const std::string &name) const {
  for (const auto &m : materials) if (m.name == name) return true;
  return false;
}

std::vector<int32_t> Lattice::site_material() const {
  std::vector<int32_t> v(num_spins);
  for (int i = 0; i < num_spins; ++i) v[i] = material_of_site(i);
  return v;
}
std::vector<int32_t> Lattice::site_motif() const {
  std::vector<int32_t> v(num_spins);
  for (int i = 0; i < num_spins; ++i) v[i] = i % M;
  return v;
}
std::vector<double> Lattice::mus() const {
  std::vector<double> v(num_spins);
  for (int i = 0; i < num_spins; ++i) v[i] = materials[material_of_site(i)].moment;
  return v;
}
std::vector<double> Lattice::alpha() const {
  std::vector<double> v(num_spins);
  for (int i = 0; i < num_spins; ++i) v[i] = materials[material_of_site(i)].alpha;
  return v;
}
std::vector<double> Lattice::gyro() const {   // core/lattice.cc:91-97,709-713
  std::vector<double> v(num_spins);
  for (int i = 0; i < num_spins; ++i) {
    const Material &m = materials[material_of_site(i)];
    v[i] = gilbert_prefactor ? m.gyro / (1.0 + m.alpha * m.alpha) : m.gyro;
  }
  return v;
}
std::vector<double> Lattice::positions() const {   // core/lattice.cc:622-657,751-756
  std::vector<double> p(3 * static_cast<size_t>(num_spins));
  size_t s = 0;
  for (int i = 0; i < dims[0]; ++i) for (int j = 0; j < dims[1]; ++j) for (int k = 0; k < dims[2]; ++k) for (int m = 0; m < M; ++m) {
    const Vec3 f = {{motif_frac[m][0] + i, motif_frac[m][1] + j, motif_frac[m][2] + k}};
    const Vec3 r = matvec(cell, f);
    p[3 * s] = r[0]; p[3 * s + 1] = r[1]; p[3 * s + 2] = r[2];
    ++s;
  }
  return p;
}
// load_array_from_tsv_file (helpers/load.h:21-47): whitespace-separated numbers, empty lines and lines starting with '#' or '//'
// skipped (helpers/utils.h:115-127), the element count must match
// A snapshot written by the spin-snapshot monitor starts with "# spins N x 3   iteration I   time_ps T": I is handed back so that a
// resumed thermal run continues the noise stream (the Philox counter is the step index) instead of replaying the first segment's.
static std::vector<double> load_spins_tsv(const std::string &file_name, size_t expected, long long *snapshot_iteration = nullptr) {
  if (file_name.size() > 3 && file_name.substr(file_name.size() - 3) == ".h5")
    throw std::runtime_error("lattice.spins: HDF5 is not available in this build; give the whitespace-separated text form (helpers/load.h:21-61)");
  std::ifstream f(file_name);
  if (!f.is_open()) throw std::runtime_error("failed to open file: " + file_name);
  std::vector<double> out;
  out.reserve(expected);
  for (std::string line; std::getline(f, line);) {
    const size_t a = line.find_first_not_of(" \t\r");
    if (a != std::string::npos && line[a] == '#' && snapshot_iteration) {
      const size_t k = line.find("iteration ");
      if (k != std::string::npos) *snapshot_iteration = std::atoll(line.c_str() + k + 10);
    }
    if (a == std::string::npos || line[a] == '#' || (line[a] == '/' && a + 1 < line.size() && line[a + 1] == '/')) continue;
    std::stringstream is(line);
    for (double v; is >> v;) out.push_back(v);
  }
  if (out.size() != expected)
    throw std::runtime_error("loading array from file: '" + file_name + "' expected size: " + std::to_string(expected) + " actual size: " + std::to_string(out.size()));
  return out;
}

std::vector<double> Lattice::initial_spins(uint64_t seed) const {   // core/lattice.cc:703-748
  if (!spins_file.empty()) return load_spins_tsv(spins_file, 3 * static_cast<size_t>(num_spins), &snapshot_iteration);
  std::vector<double> s(3 * static_cast<size_t>(num_spins));
  std::mt19937_64 rng(seed);   // the reference seeds pcg32 from std::random_device here: "random" spins are unpinned by design
  std::normal_distribution<double> nd;
  for (int i = 0; i < num_spins; ++i) {
    const Material &m = materials[material_of_site(i)];
    Vec3 spin = m.spin;
    if (m.randomize) spin = {{nd(rng), nd(rng), nd(rng)}};
    if (m.moment == 0.0) spin = {{0, 0, 0}};   // vacancies
    spin = unit_vector(spin);
    for (int n = 0; n < 3; ++n) s[3 * static_cast<size_t>(i) + n] = spin[n];
  }
  return s;
}

std::vector<Mat3> Lattice::point_group_of_motif(int m) const {   // core/lattice.cc:1127-1153
  std::vector<Mat3> out;
  for (const Mat3 &R : rotations) {
    const Vec3 np = normalise_fractional_coordinate(matvec(R, motif_frac[m]));
    if (vec_approximately_equal(motif_frac[m], np, kLatticeTolerance)) out.push_back(R);
  }
  return out;
}

InteractionTemplate Lattice::expand_interactions(const std::vector<InteractionInput> &interactions, double unit, bool fractional,
                                                 bool use_symops, double energy_cutoff, double radius_cutoff,
                                                 double distance_tolerance, double prefactor) const {
  struct Entry { int mi, mj; Vec3 r; std::array<double, 9> J9; };
  std::vector<Entry> entries;
  for (const auto &in : interactions) {
    Vec3 r = in.r;
    if (fractional) r = matvec(cell, r);
    if (in.by_motif) {
      if (in.motif_i < 0 || in.motif_i >= M || in.motif_j < 0 || in.motif_j >= M) throw std::runtime_error("interaction motif position is invalid");
      entries.push_back({in.motif_i, in.motif_j, r, in.J9});
    } else {   // complete_interaction_unitcell_positions (core/interactions.cc:98-124)
      const int ti = material_index(in.type_i), tj = material_index(in.type_j);
      for (int i = 0; i < M; ++i) {
        if (motif_material[i] != ti) continue;
        const Vec3 qf = matvec(cell_inv, r);
        const Vec3 q = {{qf[0] + motif_frac[i][0], qf[1] + motif_frac[i][1], qf[2] + motif_frac[i][2]}};
        const Vec3 T = lattice_translation_vector(q, distance_tolerance);
        const Vec3 offset = {{q[0] - T[0], q[1] - T[1], q[2] - T[2]}};
        int partner = -1;
        for (int k = 0; k < M; ++k) if (vec_approximately_equal(motif_frac[k], offset, distance_tolerance)) { partner = k; break; }
        if (partner < 0 || motif_material[partner] != tj) continue;
        entries.push_back({i, partner, r, in.J9});
      }
    }
  }
  if (use_symops) {   // apply_symops (core/interactions.cc:24-37) + generate_symmetric_points (core/lattice.cc:1015-1035)
    std::vector<std::vector<Mat3>> groups(M);
    for (int m = 0; m < M; ++m) groups[m] = point_group_of_motif(m);
    std::vector<Entry> expanded;
    for (const Entry &e : entries) {
      const Vec3 rf = matvec(cell_inv, e.r);
      std::vector<Vec3> pts{e.r};
      for (const Mat3 &R : groups[e.mi]) {
        const Vec3 rs = matvec(cell, matvec(R, rf));
        bool seen = false;
        for (const Vec3 &p : pts) if (vec_approximately_equal(rs, p, kLatticeTolerance)) { seen = true; break; }
        if (!seen) pts.push_back(rs);
      }
      for (const Vec3 &p : pts) expanded.push_back({e.mi, e.mj, p, e.J9});
    }
    entries.swap(expanded);
  }
  auto max_abs = [](const std::array<double, 9> &J) { double m = 0; for (double x : J) m = std::max(m, std::abs(x)); return m; };
  if (energy_cutoff > 0.0) {   // core/interactions.cc:322-326
    std::vector<Entry> kept;
    for (const Entry &e : entries) if (!definately_less_than(max_abs(e.J9), energy_cutoff, DBL_EPSILON)) kept.push_back(e);
    entries.swap(kept);
  }
  if (radius_cutoff > 0.0) {   // :327-330
    std::vector<Entry> kept;
    for (const Entry &e : entries) if (!definately_greater_than(norm(e.r), radius_cutoff, kLatticeTolerance)) kept.push_back(e);
    entries.swap(kept);
  }
  InteractionTemplate t;
  for (const Entry &e : entries) {   // :333-345 and hamiltonian/exchange.cc:162-169
    const Vec3 qf = matvec(cell_inv, e.r);
    const Vec3 q = {{qf[0] + motif_frac[e.mi][0] - motif_frac[e.mj][0], qf[1] + motif_frac[e.mi][1] - motif_frac[e.mj][1],
                     qf[2] + motif_frac[e.mi][2] - motif_frac[e.mj][2]}};
    const Vec3 T = lattice_translation_vector(q, distance_tolerance);
    std::array<double, 9> Jij;
    for (int k = 0; k < 9; ++k) Jij[k] = prefactor * unit * e.J9[k];
    if (!(max_abs(Jij) > energy_cutoff * unit)) continue;
    t.mi.push_back(e.mi); t.mj.push_back(e.mj);
    for (int k = 0; k < 3; ++k) t.T3.push_back(static_cast<int32_t>(T[k]));
    t.J9.insert(t.J9.end(), Jij.begin(), Jij.end());
  }
  return t;
}

NeighbourList Lattice::neighbour_list(const InteractionTemplate &t) const {   // core/interactions.cc:349-395
  struct Pair { int32_t i, j, entry; long long order; };
  std::vector<Pair> pairs;
  const int nt = t.size();
  pairs.reserve(static_cast<size_t>(dims[0]) * dims[1] * dims[2] * nt);
  long long order = 0;
  for (int i = 0; i < dims[0]; ++i) for (int j = 0; j < dims[1]; ++j) for (int k = 0; k < dims[2]; ++k) {
    for (int n = 0; n < nt; ++n, ++order) {
      int c[3] = {i + t.T3[3 * n], j + t.T3[3 * n + 1], k + t.T3[3 * n + 2]};
      bool ok = true;
      for (int l = 0; l < 3; ++l) {   // Lattice::apply_boundary_conditions (core/lattice.cc:987-1007)
        if (!periodic[l] && (c[l] < 0 || c[l] >= dims[l])) { ok = false; break; }
        c[l] = (c[l] + dims[l]) % dims[l];
      }
      if (!ok) continue;
      pairs.push_back({site_index(i, j, k, t.mi[n]), site_index(c[0], c[1], c[2], t.mj[n]), n, order});
    }
  }
  if (has_impurities) {
    // the reference looks for duplicate pairs first, then skips pairs whose site materials differ from the entry's types --
    // those of its motif positions -- "presumably an impurity site" (core/interactions.cc:373-385)
    std::vector<Pair> sorted(pairs);
    std::sort(sorted.begin(), sorted.end(), [](const Pair &a, const Pair &b) { return a.i != b.i ? a.i < b.i : a.j < b.j; });
    for (size_t p = 1; p < sorted.size(); ++p)
      if (sorted[p].i == sorted[p - 1].i && sorted[p].j == sorted[p - 1].j)
        throw std::runtime_error("Multiple interactions for sites " + std::to_string(sorted[p].i) + " and " + std::to_string(sorted[p].j));
    std::vector<Pair> kept;
    kept.reserve(pairs.size());
    for (const Pair &p : pairs)
      if (material_of_site(p.i) == motif_material[t.mi[p.entry]] && material_of_site(p.j) == motif_material[t.mj[p.entry]]) kept.push_back(p);
    pairs.swap(kept);
  }
  // unique values in first-insertion order (containers/unordered_vector_set.h:38-45)
  NeighbourList nl;
  std::vector<int> value_of_entry(nt, -1);
  {
    std::vector<char> used(nt, 0);
    std::vector<long long> first(nt, -1);
    for (const Pair &p : pairs) if (first[p.entry] < 0) first[p.entry] = p.order;
    std::vector<int> by_first;
    for (int n = 0; n < nt; ++n) if (first[n] >= 0) by_first.push_back(n);
    std::sort(by_first.begin(), by_first.end(), [&](int a, int b) { return first[a] < first[b]; });
    for (int n : by_first) {
      int found = -1;
      for (size_t v = 0; v < nl.values9.size() / 9; ++v) if (std::memcmp(&nl.values9[9 * v], &t.J9[9 * static_cast<size_t>(n)], 9 * sizeof(double)) == 0) { found = static_cast<int>(v); break; }
      if (found < 0) { found = static_cast<int>(nl.values9.size() / 9); nl.values9.insert(nl.values9.end(), t.J9.begin() + 9 * n, t.J9.begin() + 9 * n + 9); }
      value_of_entry[n] = found;
    }
  }
  std::sort(pairs.begin(), pairs.end(), [](const Pair &a, const Pair &b) { return a.i != b.i ? a.i < b.i : a.j < b.j; });   // jams::VectorSet keeps {i,j} sorted
  for (size_t p = 1; p < pairs.size(); ++p)
    if (pairs[p].i == pairs[p - 1].i && pairs[p].j == pairs[p - 1].j)
      throw std::runtime_error("Multiple interactions for sites " + std::to_string(pairs[p].i) + " and " + std::to_string(pairs[p].j));   // :373-381
  nl.i.reserve(pairs.size()); nl.j.reserve(pairs.size()); nl.value_id.reserve(pairs.size());
  for (const Pair &p : pairs) { nl.i.push_back(p.i); nl.j.push_back(p.j); nl.value_id.push_back(value_of_entry[p.entry]); }
  return nl;
}

// =====================================================================================================
// Hamiltonians
// =====================================================================================================
Hamiltonian::Hamiltonian(const Setting &settings, const Lattice &lattice) : lattice_(lattice) {   // core/hamiltonian.cc:117-147
  name_ = lowercase(settings.required("module").as_string());
  input_energy_unit_name_ = settings.get("energy_units", "joules");   // helpers/defaults.h:25
  if (settings.exists("unit_name")) input_energy_unit_name_ = settings.get("unit_name", "joules");
  input_energy_unit_conversion_ = energy_unit_conversion(input_energy_unit_name_);
}

Hamiltonian *Hamiltonian::create(const Setting &settings, const Lattice &lattice) {   // core/hamiltonian.cc:80-115
  const std::string module = lowercase(settings.required("module").as_string());
  if (module == "exchange") return new ExchangeHamiltonian(settings, lattice);
  if (module == "exchange-functional") return new ExchangeFunctionalHamiltonian(settings, lattice);
  if (module == "exchange-neartree") return new ExchangeNeartreeHamiltonian(settings, lattice);
  if (module == "uniaxial") return new UniaxialAnisotropyHamiltonian(settings, lattice);
  if (module == "zeeman") return new ZeemanHamiltonian(settings, lattice);
  if (module == "applied-field") return new AppliedFieldHamiltonian(settings, lattice);
  if (module == "biquadratic-exchange") return new BiquadraticExchangeHamiltonian(settings, lattice);
  throw std::runtime_error("unknown hamiltonian " + module + " (not on the llg-heun-b200-gpu path)");
}

// exc_file: discover_interaction_file_format + interactions_from_file (core/interactions.cc:126-171,205-252).  Empty lines and
// lines starting with '#' or '//' are skipped (helpers/utils.h:115-126); the first data line fixes the format: 6 columns =
// scalar J, 14 = 3x3 tensor; first two columns material names (JAMS) or 1-based motif indices (KKR)
static std::vector<InteractionInput> read_interaction_file(const std::string &path) {
  std::ifstream file(path);
  if (file.fail()) throw std::runtime_error(path + ": failed to open file");
  auto is_comment = [](const std::string &line) {
    std::stringstream ss(line);
    char a = 0, b = 0;
    ss >> a; ss >> b;
    return (!ss) || a == '#' || (a == '/' && b == '/');
  };
  auto is_int = [](const std::string &t) { return t.find_first_not_of("0123456789") == std::string::npos; };
  std::vector<std::string> lines;
  for (std::string line; std::getline(file, line);) lines.push_back(line);
  int ncols = 0, kkr = -1;
  for (const std::string &line : lines) {
    if (is_comment(line)) continue;
    std::stringstream is(line);
    std::vector<std::string> tok;
    for (std::string t; is >> t;) tok.push_back(t);
    if (tok.size() != 6 && tok.size() != 14) throw std::runtime_error("interaction file has an incorrect number of columns");
    ncols = static_cast<int>(tok.size());
    if (is_int(tok[0]) != is_int(tok[1])) break;
    kkr = is_int(tok[0]) ? 1 : 0;
    break;
  }
  if (kkr < 0) throw std::runtime_error("failed to discover interaction file format");
  std::vector<InteractionInput> out;
  int line_number = 0;
  for (const std::string &line : lines) {
    if (is_comment(line)) continue;
    std::stringstream is(line);
    InteractionInput in;
    in.by_motif = kkr == 1;
    if (kkr) { is >> in.motif_i >> in.motif_j; in.motif_i--; in.motif_j--; }
    else is >> in.type_i >> in.type_j;
    is >> in.r[0] >> in.r[1] >> in.r[2];
    in.J9.fill(0.0);
    if (ncols == 6) { double J = 0; is >> J; in.J9[0] = in.J9[4] = in.J9[8] = J; }
    else for (int k = 0; k < 9; ++k) is >> in.J9[k];
    if (is.bad() || is.fail()) throw std::runtime_error("failed to read line " + std::to_string(line_number) + " of interaction file");
    out.push_back(in);
  }
  return out;
}

ExchangeHamiltonian::ExchangeHamiltonian(const Setting &s, const Lattice &lattice) : Hamiltonian(s, lattice) {
  check_symmetry_ = s.get("check_sparse_matrix_symmetry", true);
  parse_interactions(s, lattice, s.get("interaction_prefactor", 1.0));
}

void ExchangeHamiltonian::parse_interactions(const Setting &s, const Lattice &lattice, double prefactor) {
  const bool use_symops = s.get("symops", true);
  const double energy_cutoff = s.get("energy_cutoff", 0.0), radius_cutoff = s.get("radius_cutoff", 100.0);
  const double distance_tolerance = s.get("distance_tolerance", kLatticeTolerance);
  const std::string coord = lowercase(s.get("coordinate_format", "cartesian"));
  if (coord != "cartesian" && coord != "fractional") throw std::runtime_error("Unknown coordinate format for exchange interactions");
  std::vector<InteractionInput> inputs;
  if (s.exists("exc_file")) {   // kept by the reference for backwards compatibility; wins over 'interactions' (hamiltonian/exchange.cc:118-145)
    inputs = read_interaction_file(s["exc_file"].as_string());
  } else if (s.exists("interactions")) {
    const Setting &list = s["interactions"];
    if (!list.is_list() || list.length() < 1) throw std::runtime_error("exchange settings must be a list");
    // discover_interaction_setting_format (core/interactions.cc:172-203)
    if (list[0][0].is_number() != list[0][1].is_number()) throw std::runtime_error("interaction type format is incorrect");
    const bool kkr = list[0][0].is_number();
    if (!list[0][2].is_array()) throw std::runtime_error("interaction vector format is incorrect");
    const bool scalar = list[0][3].is_number();
    if (!scalar && !(list[0][3].is_array() && list[0][3].length() == 9)) throw std::runtime_error("interaction energy format is incorrect");
    for (int i = 0; i < list.length(); ++i) {   // interactions_from_settings (:254-289)
      InteractionInput in;
      in.by_motif = kkr;
      if (kkr) { in.motif_i = static_cast<int>(list[i][0].as_int()) - 1; in.motif_j = static_cast<int>(list[i][1].as_int()) - 1; }
      else { in.type_i = list[i][0].as_string(); in.type_j = list[i][1].as_string(); }
      in.r = read_vec3(list[i][2]);
      in.J9.fill(0.0);
      if (scalar) { const double J = list[i][3].as_double(); in.J9[0] = in.J9[4] = in.J9[8] = J; }
      else for (int k = 0; k < 9; ++k) in.J9[k] = list[i][3][k].as_double();
      inputs.push_back(in);
    }
  } else {
    throw std::runtime_error("'exc_file' or 'interactions' settings are required");
  }
  template_ = lattice.expand_interactions(inputs, input_energy_unit_conversion_, coord == "fractional", use_symops, energy_cutoff,
                                          radius_cutoff, distance_tolerance, prefactor);
}

// ---- exchange-functional (hamiltonian/exchange_functional.cc) -------------------------------------------------------
double Lattice::max_interaction_radius() const {
  Vec3 a[3];
  for (int k = 0; k < 3; ++k) a[k] = {{cell[0][k] * dims[k], cell[1][k] * dims[k], cell[2][k] * dims[k]}};
  auto add = [](const Vec3 &u, const Vec3 &v, double sv) { return Vec3{{u[0] + sv * v[0], u[1] + sv * v[1], u[2] + sv * v[2]}}; };
  auto height = [](const Vec3 &u, const Vec3 &v, const Vec3 &w) { return std::abs(dot(cross(u, v), w)) / norm(cross(u, v)); };   // parallelepiped_height
  auto pheight = [](const Vec3 &u, const Vec3 &v) { return norm(cross(u, v)) / norm(u); };                                       // parallelogram_height
  const int np = (periodic[0] ? 1 : 0) + (periodic[1] ? 1 : 0) + (periodic[2] ? 1 : 0);
  if (np == 3) return 0.5 * std::min({height(a[0], a[1], a[2]), height(a[2], a[0], a[1]), height(a[1], a[2], a[0])});
  if (np == 2) {
    const int k0 = periodic[0] ? 0 : 1, k1 = periodic[2] ? 2 : 1;
    return 0.5 * std::min(pheight(a[k0], a[k1]), pheight(a[k1], a[k0]));
  }
  if (np == 1) return 0.5 * norm(a[periodic[0] ? 0 : (periodic[1] ? 1 : 2)]);
  const Vec3 s01 = add(a[0], a[1], 1.0);
  return std::max({norm(add(s01, a[2], 1.0)), norm(add(add(a[1], a[0], -1.0), a[2], 1.0)), norm(add(add(a[0], a[1], -1.0), a[2], 1.0)), norm(add(s01, a[2], -1.0))});
}

ExchangeFunctionalHamiltonian::ExchangeFunctionalHamiltonian(const Setting &s, const Lattice &lattice) : ExchangeHamiltonian(s, lattice, NoParse{}) {
  if (lattice.has_impurities) throw std::runtime_error("exchange-functional is not supported on a lattice with impurities by the llg-heun-b200-gpu host layer");   // the template form assumes translation invariance
  const double tol = kLatticeTolerance, E = input_energy_unit_conversion_;
  const std::string dunit = s.get("distance_units", "lattice_constants");   // core/hamiltonian.cc:140-155
  double D = 1.0;
  if (dunit == "nanometers") D = 1e-9 / lattice.lattice_parameter;
  else if (dunit == "angstroms") D = 1e-10 / lattice.lattice_parameter;
  else if (dunit != "lattice_constants") throw std::runtime_error("distance units: " + dunit + " is not known");
  if (!s.exists("interactions")) throw std::runtime_error("no 'interactions' setting in ExchangeFunctional hamiltonian");
  using Fn = std::function<double(const Vec3 &)>;
  std::map<std::pair<std::string, std::string>, std::pair<double, Fn>> functionals;
  double rmax = 0.0;
  const Setting &list = s["interactions"];
  for (int n = 0; n < list.length(); ++n) {   // exchange_functional.cc:118-188
    const Setting &e = list[n];
    if (e.length() < 4) throw std::runtime_error("interaction requires at least 4 elements");
    const std::string ti = e[0].as_string(), tj = e[1].as_string(), name = e[2].as_string();
    const double rc = D * e[3].as_double();
    for (const std::string &t : {ti, tj}) if (!lattice.material_exists(t)) throw std::runtime_error("material " + t + " does not exist in config");
    if ((0.0 - rc) > std::abs(rc) * tol) throw std::runtime_error("cutoff radius cannot be negative");
    if (functionals.count({ti, tj})) throw std::runtime_error("Interaction between types \"" + ti + "\" and \"" + tj + "\" is defined more than once.");
    if (rc > lattice.max_interaction_radius())
      throw std::runtime_error("cutoff radius " + std::to_string(rc) + " is larger than the maximum cutoff radius " + std::to_string(lattice.max_interaction_radius()));
    rmax = std::max(rmax, rc);
    std::vector<double> p;
    for (int k = 4; k < e.length(); ++k) {
      if (e[k].is_aggregate()) { for (int l = 0; l < e[k].length(); ++l) { if (!e[k][l].is_number()) throw std::runtime_error("functional parameter must be numeric"); p.push_back(e[k][l].as_double()); } }
      else { if (!e[k].is_number()) throw std::runtime_error("functional parameter must be numeric"); p.push_back(e[k].as_double()); }
    }
    // validate_functional_params + functional_from_params (:13-88,358-416)
    const std::map<std::string, int> npar = {{"rkky", 3}, {"exponential", 3}, {"gaussian", 3}, {"gaussian_multi", 9}, {"kaneyoshi", 3}, {"c3z", 14}, {"step", 2}};
    if (!npar.count(name)) throw std::runtime_error("unknown exchange functional: " + name);
    if (static_cast<int>(p.size()) != npar.at(name))
      throw std::runtime_error("exchange functional '" + name + "' expects " + std::to_string(npar.at(name)) + " parameters, got " + std::to_string(p.size()));
    auto non_zero = [&](size_t i, const char *pn) { if (std::abs(p[i]) <= tol) throw std::runtime_error("exchange functional '" + name + "' requires non-zero parameter '" + pn + "'"); };
    auto positive = [&](size_t i, const char *pn) { if (!((p[i] - 0.0) > std::abs(p[i]) * tol)) throw std::runtime_error("exchange functional '" + name + "' requires positive parameter '" + pn + "'"); };
    auto gauss = [](double r, double J0, double r0, double sg) { return J0 * std::exp(-(r - r0) * (r - r0) / (2 * sg * sg)); };
    Fn fn;
    if (name == "step") { const double J0 = E * p[0], rcut = D * p[1]; fn = [=](const Vec3 &r) { const double x = norm(r); return (x - rcut) < std::max(std::abs(x), std::abs(rcut)) * tol ? J0 : 0.0; }; }
    else if (name == "exponential") { non_zero(2, "sigma"); const double J0 = E * p[0], r0 = D * p[1], sg = D * p[2]; fn = [=](const Vec3 &r) { return J0 * std::exp(-(norm(r) - r0) / sg); }; }
    else if (name == "gaussian") { non_zero(2, "sigma"); const double J0 = E * p[0], r0 = D * p[1], sg = D * p[2]; fn = [=](const Vec3 &r) { return gauss(norm(r), J0, r0, sg); }; }
    else if (name == "gaussian_multi") {
      non_zero(2, "sigma0"); non_zero(5, "sigma1"); non_zero(8, "sigma2");
      const std::vector<double> q = p;
      fn = [=](const Vec3 &r) { double sum = 0; for (int k = 0; k < 3; ++k) sum += gauss(norm(r), E * q[3 * k], D * q[3 * k + 1], D * q[3 * k + 2]); return sum; };
    }
    else if (name == "kaneyoshi") { non_zero(2, "sigma"); const double J0 = E * p[0], r0 = D * p[1], sg = D * p[2];
      fn = [=](const Vec3 &r) { const double x = norm(r) - r0; return J0 * x * x * std::exp(-x * x / (2 * sg * sg)); }; }
    else if (name == "rkky") { non_zero(2, "k_F"); const double J0 = E * p[0], r0 = D * p[1], kF = p[2];
      fn = [=](const Vec3 &r) { const double kr = 2 * kF * (norm(r) - r0);
        if (std::abs(kr) <= tol) throw std::runtime_error("exchange functional rkky is singular for k_F*(r-r0) = 0");
        return -J0 * (kr * std::cos(kr) - std::sin(kr)) / (kr * kr * kr * kr); }; }
    else {   // c3z (:310-356)
      positive(10, "l0"); positive(11, "l1s"); positive(12, "l1c");
      const Vec3 qs1{{p[0] / D, p[1] / D, p[2] / D}}, qc1{{p[3] / D, p[4] / D, p[5] / D}};
      const double J0 = E * p[6], J1s = E * p[7], J1c = E * p[8], d0 = D * p[9], l0 = D * p[10], l1s = D * p[11], l1c = D * p[12], rstar = D * p[13];
      fn = [=](const Vec3 &rij) {
        const double r = norm(rij);
        const Vec3 rpar{{rij[0], rij[1], 0.0}};
        double ssum = 0, csum = 0;
        for (int k = 0; k < 3; ++k) {
          const double t = 2 * kPi * k / 3.0, c = std::cos(t), sn = std::sin(t);
          const Vec3 qs{{c * qs1[0] - sn * qs1[1], sn * qs1[0] + c * qs1[1], qs1[2]}}, qc{{c * qc1[0] - sn * qc1[1], sn * qc1[0] + c * qc1[1], qc1[2]}};
          ssum += std::sin(dot(qs, rpar)); csum += std::cos(dot(qc, rpar));
        }
        return J0 * std::exp(-std::abs(r - d0) / l0) + J1s * std::exp(-std::abs(r - rstar) / l1s) * ssum + J1c * std::exp(-std::abs(r - rstar) / l1c) * csum;
      };
    }
    functionals[{ti, tj}] = {rc, fn};
  }
  // the near-tree walk of :206-243 as a translation-invariant template
  const Vec3 c0{{lattice.cell[0][0], lattice.cell[1][0], lattice.cell[2][0]}}, c1{{lattice.cell[0][1], lattice.cell[1][1], lattice.cell[2][1]}},
             c2{{lattice.cell[0][2], lattice.cell[1][2], lattice.cell[2][2]}};
  const Vec3 cols[3] = {c0, c1, c2};
  const double vol = std::abs(dot(cross(c0, c1), c2));
  int nmax[3];
  for (int k = 0; k < 3; ++k) nmax[k] = static_cast<int>(std::ceil(rmax * (1 + tol) / (vol / norm(cross(cols[(k + 1) % 3], cols[(k + 2) % 3]))))) + 1;
  for (int mi = 0; mi < lattice.M; ++mi) {
    const Vec3 ri = matvec(lattice.cell, lattice.motif_frac[mi]);
    for (int mj = 0; mj < lattice.M; ++mj) {
      const auto it = functionals.find({lattice.materials[lattice.motif_material[mi]].name, lattice.materials[lattice.motif_material[mj]].name});
      if (it == functionals.end()) continue;
      const double rc = it->second.first;
      for (int tx = -nmax[0]; tx <= nmax[0]; ++tx) for (int ty = -nmax[1]; ty <= nmax[1]; ++ty) for (int tz = -nmax[2]; tz <= nmax[2]; ++tz) {
        if (mi == mj && tx == 0 && ty == 0 && tz == 0) continue;   // no self interaction (:213-215)
        const Vec3 f{{lattice.motif_frac[mj][0] + tx, lattice.motif_frac[mj][1] + ty, lattice.motif_frac[mj][2] + tz}};
        const Vec3 rj = matvec(lattice.cell, f);
        const Vec3 rij{{rj[0] - ri[0], rj[1] - ri[1], rj[2] - ri[2]}};
        const double r = norm(rij);
        if (!((r - rc) < std::max(std::abs(r), std::abs(rc)) * tol)) continue;   // less_than_approx_equal (helpers/maths.h:37-40)
        const double J = it->second.second(rij);
        template_.mi.push_back(mi); template_.mj.push_back(mj);
        template_.T3.push_back(tx); template_.T3.push_back(ty); template_.T3.push_back(tz);
        for (int k = 0; k < 9; ++k) template_.J9.push_back((k % 4 == 0) ? J : 0.0);
      }
    }
  }
}

// ---- exchange-neartree (hamiltonian/exchange_neartree.cc:14-156) ----------------------------------------------------
ExchangeNeartreeHamiltonian::ExchangeNeartreeHamiltonian(const Setting &s, const Lattice &lattice) : ExchangeHamiltonian(s, lattice, NoParse{}) {
  if (lattice.has_impurities) throw std::runtime_error("exchange-neartree is not supported on a lattice with impurities by the llg-heun-b200-gpu host layer");   // the template form assumes translation invariance
  const double E = input_energy_unit_conversion_;
  const std::string dunit = s.get("distance_units", "lattice_constants");
  double D = 1.0;
  if (dunit == "nanometers") D = 1e-9 / lattice.lattice_parameter;
  else if (dunit == "angstroms") D = 1e-10 / lattice.lattice_parameter;
  else if (dunit != "lattice_constants") throw std::runtime_error("distance units: " + dunit + " is not known");
  const double energy_cutoff = s.get("energy_cutoff", 1e-26) * E, shell_width = s.get("shell_width", 1e-3) * D;
  for (int i = 0; i < lattice.M; ++i) for (int j = i + 1; j < lattice.M; ++j) {   // :43-56
    const Vec3 d{{lattice.motif_frac[i][0] - lattice.motif_frac[j][0], lattice.motif_frac[i][1] - lattice.motif_frac[j][1], lattice.motif_frac[i][2] - lattice.motif_frac[j][2]}};
    if (norm(d) < shell_width)
      throw std::runtime_error("Atoms " + std::to_string(i) + " and " + std::to_string(j) + " in the unit cell are close together than the shell_width");
  }
  if (!s.exists("interactions")) throw std::runtime_error("no 'interactions' setting in ExchangeNeartree hamiltonian");
  struct Shell { int A, B; double radius, J; };
  std::vector<Shell> shells;
  double max_radius = 0.0;
  const Setting &list = s["interactions"];
  for (int n = 0; n < list.length(); ++n) {   // :64-89
    const std::string ta = list[n][0].as_string(), tb = list[n][1].as_string();
    for (const std::string &t : {ta, tb})
      if (!lattice.material_exists(t)) throw std::runtime_error("exchange neartree interaction " + std::to_string(n) + ": material " + t + " does not exist in the config");
    const double radius = list[n][2].as_double() * D, J = list[n][3].as_double() * E;
    max_radius = std::max(max_radius, radius);
    const int A = lattice.material_index(ta), B = lattice.material_index(tb);
    shells.push_back({A, B, radius, J});
    if (A != B) shells.push_back({B, A, radius, J});
  }
  if (shells.empty()) return;
  // InteractionNearTree::shell -> NearTree::in_annulus (containers/neartree.h:359-376) with epsilon = shell_width / 10, as a template
  const double eps = shell_width / 10.0;
  auto gt = [&](double a, double b) { return (a - b) > std::max(std::abs(a), std::abs(b)) * eps; };   // definately_greater_than
  const Vec3 c0{{lattice.cell[0][0], lattice.cell[1][0], lattice.cell[2][0]}}, c1{{lattice.cell[0][1], lattice.cell[1][1], lattice.cell[2][1]}},
             c2{{lattice.cell[0][2], lattice.cell[1][2], lattice.cell[2][2]}};
  const Vec3 cols[3] = {c0, c1, c2};
  const double vol = std::abs(dot(cross(c0, c1), c2)), rmax = (max_radius + shell_width) * (1 + eps);
  int nmax[3];
  for (int k = 0; k < 3; ++k) nmax[k] = static_cast<int>(std::ceil(rmax / (vol / norm(cross(cols[(k + 1) % 3], cols[(k + 2) % 3]))))) + 1;
  for (int mi = 0; mi < lattice.M; ++mi) {
    const Vec3 ri = matvec(lattice.cell, lattice.motif_frac[mi]);
    for (int mj = 0; mj < lattice.M; ++mj)
      for (int tx = -nmax[0]; tx <= nmax[0]; ++tx) for (int ty = -nmax[1]; ty <= nmax[1]; ++ty) for (int tz = -nmax[2]; tz <= nmax[2]; ++tz) {
        if (mi == mj && tx == 0 && ty == 0 && tz == 0) continue;   // :119-121
        const Vec3 f{{lattice.motif_frac[mj][0] + tx, lattice.motif_frac[mj][1] + ty, lattice.motif_frac[mj][2] + tz}};
        const Vec3 rj = matvec(lattice.cell, f);
        const double r = norm(Vec3{{rj[0] - ri[0], rj[1] - ri[1], rj[2] - ri[2]}});
        bool seen = false;
        for (const Shell &sh : shells) {
          if (lattice.motif_material[mi] != sh.A || lattice.motif_material[mj] != sh.B) continue;
          const double inner = sh.radius - 0.5 * shell_width, outer = sh.radius + 0.5 * shell_width;
          if (gt(r, outer) || !gt(r, inner)) continue;
          if (seen) throw std::runtime_error("multiple interactions between spins of motif positions " + std::to_string(mi) + " and " + std::to_string(mj));
          seen = true;
          if (std::abs(sh.J) > energy_cutoff) {
            template_.mi.push_back(mi); template_.mj.push_back(mj);
            template_.T3.push_back(tx); template_.T3.push_back(ty); template_.T3.push_back(tz);
            for (int k = 0; k < 9; ++k) template_.J9.push_back((k % 4 == 0) ? sh.J : 0.0);
          }
        }
      }
  }
}

UniaxialAnisotropyHamiltonian::UniaxialAnisotropyHamiltonian(const Setting &s, const Lattice &lattice) : Hamiltonian(s, lattice) {
  for (const char *old : {"d2z", "d4z", "d6z", "K1", "K2", "K3"})
    if (s.exists(old)) throw std::runtime_error("UniaxialHamiltonian: anisotropy should only be specified for a single K1, K2 or K3.");
  const std::string order = s.required("order").as_string();
  if (order == "K1") power_ = 2; else if (order == "K2") power_ = 4; else if (order == "K3") power_ = 6;
  else throw std::runtime_error("Unsupported anisotropy: " + order);
  const int N = lattice.num_spins;
  magnitude_.assign(N, 0.0); axis_.assign(3 * static_cast<size_t>(N), 0.0);
  const Setting &list = s.required("anisotropies");
  for (int a = 0; a < list.length(); ++a) {   // hamiltonian/uniaxial_anisotropy.cc:40-70,89-114
    const Setting &e = list[a];
    int motif_position = -1, material = -1;
    if (e[0].is_number()) {
      motif_position = static_cast<int>(e[0].as_int()) - 1;
      if (motif_position < 0 || motif_position >= lattice.M) throw std::runtime_error("uniaxial anisotropy motif position is invalid");
    } else {
      if (!lattice.material_exists(e[0].as_string())) throw std::runtime_error("uniaxial anisotropy material is invalid");
      material = lattice.material_index(e[0].as_string());
    }
    Vec3 axis = read_vec3(e[1]);
    const double n = norm(axis);
    axis = {{axis[0] / n, axis[1] / n, axis[2] / n}};   // normalize()
    const double energy = e[2].as_double();
    for (int i = 0; i < N; ++i) {
      if ((motif_position >= 0 && i % lattice.M == motif_position) || (material >= 0 && lattice.material_of_site(i) == material)) {
        magnitude_[i] = energy * input_energy_unit_conversion_;
        for (int k = 0; k < 3; ++k) axis_[3 * static_cast<size_t>(i) + k] = axis[k];
      }
    }
  }
}

ZeemanHamiltonian::ZeemanHamiltonian(const Setting &s, const Lattice &lattice) : Hamiltonian(s, lattice) {   // hamiltonian/zeeman.cc:12-72
  const int N = lattice.num_spins, nmat = static_cast<int>(lattice.materials.size());
  const std::vector<double> mus = lattice.mus();
  dc_local_field_.assign(3 * static_cast<size_t>(N), 0.0);
  if (const Setting *dc = s.find("dc_local_field")) {
    if (dc->length() != nmat) throw std::runtime_error("dc_local_field: field must be specified for every material");
    for (int i = 0; i < N; ++i) for (int k = 0; k < 3; ++k)
      dc_local_field_[3 * static_cast<size_t>(i) + k] = (*dc)[lattice.material_of_site(i)][k].as_double() * mus[i];
  }
  if (s.exists("ac_local_field") || s.exists("ac_local_frequency")) {
    if (!(s.exists("ac_local_field") && s.exists("ac_local_frequency"))) throw std::runtime_error("ac_local_field: must have a field and a frequency");
    const Setting &f = s["ac_local_field"], &w = s["ac_local_frequency"];
    if (w.length() != nmat || f.length() != nmat) throw std::runtime_error("ac_local_frequency: must be specified for every material");
    has_ac_local_field_ = true;
    ac_local_field_.assign(3 * static_cast<size_t>(N), 0.0); ac_local_frequency_.assign(N, 0.0);
    for (int i = 0; i < N; ++i) {
      const int mat = lattice.material_of_site(i);
      for (int k = 0; k < 3; ++k) ac_local_field_[3 * static_cast<size_t>(i) + k] = f[mat][k].as_double() * mus[i];
      ac_local_frequency_[i] = kTwoPi * w[mat].as_double();
    }
  }
}

AppliedFieldHamiltonian::AppliedFieldHamiltonian(const Setting &s, const Lattice &lattice) : Hamiltonian(s, lattice) {
  const std::string type = lowercase(s.get("type", "static"));
  if (type == "static") type_ = JB_FIELD_STATIC;
  else if (type == "sinc") type_ = JB_FIELD_SINC;
  else if (type == "sinc-cos") type_ = JB_FIELD_SINC_COS;
  else throw std::runtime_error("Unknown field pulse type " + type);
  field_ = read_vec3(s.required("field"));
  if (type_ != JB_FIELD_STATIC) {
    time_center_ = s.required("time_center").as_double() / 1e-12;
    freq_bandwidth_ = s.required("freq_bandwidth").as_double() / 1e12;
    if (type_ == JB_FIELD_SINC_COS) freq_center_ = s.required("freq_center").as_double() / 1e12;
  }
  name_ += "-" + type;
}

// =====================================================================================================
// Physics, Monitors
// =====================================================================================================
Physics::Physics(const Setting *s, const Setting *sim, const std::string &prefix) {   // core/physics.cc: temperature / applied_field of the `physics` group
  if (!s) return;
  const std::string module = lowercase(s->get("module", "empty"));
  if (module != "empty" && module != "pinned_boundaries" && module != "field-cool" && module != "two-temperature-model")
    throw std::runtime_error("physics module '" + module + "' is not supported by the llg-heun-b200-gpu host layer");
  temperature_ = s->get("temperature", 0.0);
  if (const Setting *f = s->find("applied_field")) applied_field_ = read_vec3(*f);
  output_step_freq_ = s->get("output_steps", 100);
  if (module == "field-cool") {   // physics/field_cool.cc:9-54
    field_cool_ = true;
    init_temp_ = s->required("InitialTemperature").as_double();
    final_temp_ = s->required("FinalTemperature").as_double();
    integration_time_step_ = sim ? sim->get("t_step", 0.0) : 0.0;   // globals::config->lookup("sim.t_step"), as written in the file
    init_field_ = read_vec3(s->required("InitialField"));
    final_field_ = read_vec3(s->required("FinalField"));
    cool_time_ = s->required("CoolTime").as_double();
    for (int i = 0; i < 3; ++i) applied_field_[i] += init_field_[i];
    if (s->exists("TSteps")) {
      t_steps_ = s->get("TSteps", 1);
      delta_T_ = (init_temp_ - final_temp_) / t_steps_;
      t_plateau_ = cool_time_ / t_steps_;
      t_eq_ = sim ? sim->get("t_eq", 0.0) : 0.0;
      step_toggle_ = true;
    }
    temperature_ = init_temp_;
  }
  if (module == "two-temperature-model") {   // physics/two_temperature_model.cc:14-60
    ttm_ = true;
    phonon_temp_ = s->required("InitialTemperature").as_double();
    electron_temp_ = sink_temp_ = phonon_temp_;
    if (const Setting *pulses = s->find("laserPulses")) {
      for (int i = 0; i < pulses->length(); ++i) {
        const Setting &q = (*pulses)[i];
        pulse_width_.push_back(q.required("width").as_double());
        pulse_fluence_.push_back(1.152e20 * q.required("fluence").as_double());   // pumpPower (two_temperature_model.h:23)
        pulse_start_.push_back(q.required("t_start").as_double());
      }
    }
    Ce_ = s->get("Ce", Ce_); Cl_ = s->get("Cl", Cl_); G_ = s->get("Gep", G_); Gsink_ = s->get("Gps", Gsink_);
    reversing_field_ = read_vec3(s->required("ReversingField"));
    if (!s->exists("temperature")) temperature_ = electron_temp_;
    if (!prefix.empty()) {
      ttm_file_.open(prefix + "ttm.tsv");
      ttm_file_ << std::setprecision(8) << "# t [s]\tT_el [K]\tT_ph [K]\tLaser [arb/]\n";
    }
  }
  if (module == "pinned_boundaries") {   // physics/pinned_boundaries.cc:12-31, pinned_boundaries.h:86-107
    const char *names[6] = {"left", "right", "front", "back", "bottom", "top"};
    for (int k = 0; k < 6; ++k) {
      const std::string name = names[k];
      const Setting *m = s->find(name + "_pinned_magnetisation");
      if (!m) continue;
      boundaries_.push_back({k / 2, (k % 2) == 1, s->get(name + "_pinned_cells", 1), read_vec3(*m)});
    }
  }
}

void Physics::update(B200HeunLLGSolver &solver) {   // physics/pinned_boundaries.cc:34-46, field_cool.cc:56-78, two_temperature_model.cc:66-96
  const double time = solver.time();
  if (field_cool_ && time > t_eq_) {
    if (step_toggle_) {
      const int count = static_cast<int>((time - t_eq_) / t_plateau_);
      if (count < t_steps_ + 1) temperature_ = init_temp_ - count * delta_T_;
    } else if (time < cool_time_) {
      for (int i = 0; i < 3; ++i) applied_field_[i] += ((final_field_[i] - init_field_[i]) * integration_time_step_) / cool_time_;
      temperature_ += ((final_temp_ - init_temp_) * integration_time_step_) / cool_time_;
    }
  }
  if (ttm_) {
    const double dt = solver.time_step();
    applied_field_ = reversing_field_;
    double pump = 0.0;
    for (size_t i = 0; i < pulse_fluence_.size(); ++i) {
      const double rel = time - pulse_start_[i];
      if (rel > 0.0 && rel <= 10 * pulse_width_[i]) {
        const double a = (rel - 3 * pulse_width_[i]) / pulse_width_[i];
        pump += pulse_fluence_[i] * std::exp(-a * a);
      }
    }
    electron_temp_ = electron_temp_ + ((-G_ * (electron_temp_ - phonon_temp_) + pump) * dt) / (Ce_ * electron_temp_);
    phonon_temp_ = phonon_temp_ + ((G_ * (electron_temp_ - phonon_temp_) - Gsink_ * (phonon_temp_ - sink_temp_)) * dt) / Cl_;
    temperature_ = electron_temp_;
    if (ttm_file_.is_open() && solver.iteration() % output_step_freq_ == 0)
      ttm_file_ << time << "\t" << electron_temp_ << "\t" << phonon_temp_ << "\t" << pump << "\n";
  }
  if (boundaries_.empty()) return;
  const Lattice &lat = solver.lattice();
  jb_ctx *ctx = solver.ctx();
  if (!regions_set_) {
    for (size_t r = 0; r < boundaries_.size(); ++r) {
      const PinnedBoundary &b = boundaries_[r];
      std::vector<int32_t> sites;
      for (int x = 0; x < lat.dims[0]; ++x) for (int y = 0; y < lat.dims[1]; ++y) for (int z = 0; z < lat.dims[2]; ++z) {
        const int cell[3] = {x, y, z};
        const bool in = b.upper ? cell[b.dim] >= lat.dims[b.dim] - b.cells : cell[b.dim] < b.cells;
        if (!in) continue;
        for (int m = 0; m < lat.M; ++m) sites.push_back(((x * lat.dims[1] + y) * lat.dims[2] + z) * lat.M + m);
      }
      solver.check(jb_set_region(ctx, static_cast<int32_t>(r), static_cast<int32_t>(sites.size()), sites.data()));
    }
    regions_set_ = true;
  }
  for (size_t r = 0; r < boundaries_.size(); ++r) {
    double M4[4];
    solver.check(jb_region_moment(ctx, static_cast<int32_t>(r), M4));
    const Mat3 R = rotation_matrix_between_vectors(Vec3{{M4[0], M4[1], M4[2]}}, boundaries_[r].magnetisation);
    double R9[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R9[3 * i + j] = R[i][j];
    solver.check(jb_rotate_region(ctx, static_cast<int32_t>(r), R9));
  }
}

Monitor::Monitor(const Setting &settings) { output_step_freq_ = settings.get("output_steps", 100); }

Monitor *Monitor::create(const Setting &settings, const Lattice &lattice, const std::string &prefix) {   // core/monitor.cc
  const std::string module = lowercase(settings.required("module").as_string());
  if (module == "magnetisation") return new MagnetisationMonitor(settings, lattice, prefix + "mag.tsv");
  if (module == "energy") return new EnergyMonitor(settings, prefix + "eng.tsv");
  if (module == "magnetisation-layers") return new MagnetisationLayersMonitor(settings, lattice, prefix + "mag_layers.tsv");
  if (module == "hdf5" || module == "spins-tsv") {
    if (module == "hdf5") std::fprintf(stderr, "jams-b200: monitor 'hdf5': HDF5 is not available in this build, spin snapshots are written as text (%sNNNNNNN.tsv, %sfinal.tsv)\n", prefix.c_str(), prefix.c_str());
    return new SpinsTsvMonitor(settings, prefix);
  }
  throw std::runtime_error("unknown monitor " + module + " (not supported by the llg-heun-b200-gpu host layer)");
}

MagnetisationMonitor::MagnetisationMonitor(const Setting &settings, const Lattice &lattice, const std::string &filename)
    : Monitor(settings), lattice_(lattice), tsv_file_(filename) {
  const std::string g = lowercase(settings.get("grouping", "materials"));
  if (g == "none") { grouping_ = Grouping::NONE; n_groups_ = 1; }
  else if (g == "materials") { grouping_ = Grouping::MATERIALS; n_groups_ = static_cast<int>(lattice.materials.size()); group_of_spin_ = lattice.site_material(); }
  else if (g == "positions") { grouping_ = Grouping::POSITIONS; n_groups_ = lattice.M; group_of_spin_ = lattice.site_motif(); }
  else throw std::runtime_error("unknown magnetisation grouping: " + g);
  normalize_ = settings.get("normalize", true);
  if (!tsv_file_) throw std::runtime_error("cannot open " + filename);
  tsv_file_.setf(std::ios::right);
  tsv_file_ << tsv_header();
}

std::string MagnetisationMonitor::tsv_header() const {   // monitors/magnetisation.cc:106-141
  std::stringstream ss;
  ss.width(12);
  for (const char *h : {"time", "T", "hx", "hy", "hz"}) ss << fmt_sci << h;
  if (grouping_ == Grouping::NONE) {
    for (const char *n : {"mx", "my", "mz", "m"}) ss << fmt_decimal << n;
  } else if (grouping_ == Grouping::MATERIALS) {
    for (const auto &m : lattice_.materials) for (const char *suf : {"_mx", "_my", "_mz", "_m"}) ss << fmt_decimal << m.name + suf;
  } else {
    for (int i = 0; i < lattice_.M; ++i)
      for (const char *suf : {"_mx", "_my", "_mz", "_m"}) ss << fmt_sci << std::to_string(i + 1) + "_" + lattice_.materials[lattice_.motif_material[i]].name + suf;
  }
  ss << std::endl;
  return ss.str();
}

void MagnetisationMonitor::update(B200HeunLLGSolver &solver) {   // monitors/magnetisation.cc:77-104
  std::vector<double> M4(4 * static_cast<size_t>(n_groups_));
  if (registered_ctx_ != solver.ctx()) {   // the groups are fixed at construction: one upload, not one per update
    solver.check(jb_set_magnetisation_groups(solver.ctx(), n_groups_, group_of_spin_.empty() ? nullptr : group_of_spin_.data()));
    registered_ctx_ = solver.ctx();
  }
  solver.check(jb_magnetisation(solver.ctx(), n_groups_, nullptr, M4.data()));
  tsv_file_.width(12);
  tsv_file_ << fmt_sci << solver.time();
  tsv_file_ << fmt_sci << solver.physics()->temperature();
  for (int i = 0; i < 3; ++i) tsv_file_ << fmt_sci << solver.physics()->applied_field(i);
  for (int n = 0; n < n_groups_; ++n) {
    const double *m = &M4[4 * static_cast<size_t>(n)];
    const double factor = normalize_ ? 1.0 / m[3] : 1.0 / kBohrMagnetonIU;
    tsv_file_ << fmt_sci << m[0] * factor << fmt_sci << m[1] * factor << fmt_sci << m[2] * factor
              << fmt_sci << std::sqrt(m[0] * m[0] + m[1] * m[1] + m[2] * m[2]) * factor;
  }
  tsv_file_ << std::endl;
}

MagnetisationLayersMonitor::MagnetisationLayersMonitor(const Setting &settings, const Lattice &lattice, const std::string &filename)
    : Monitor(settings), lattice_(lattice), tsv_file_(filename) {   // monitors/magnetisation_layers.cc:15-206
  if (!tsv_file_) throw std::runtime_error("cannot open " + filename);
  const Setting &ns = settings.required("layer_normal");
  const Vec3 layer_normal{{ns[0].as_double(), ns[1].as_double(), ns[2].as_double()}};
  const double layer_thickness = settings.get("layer_thickness", 0.0);
  const double distance_tolerance = settings.get("distance_tolerance", 1e-4);   // jams::defaults::lattice_tolerance
  const std::string grouping = lowercase(settings.get("grouping", "materials"));
  const int N = lattice.num_spins;
  std::vector<std::vector<int>> groups;
  if (grouping == "none") {
    groups.emplace_back(N);
    for (int i = 0; i < N; ++i) groups[0][i] = i;
    group_names_.push_back("total");
  } else if (grouping == "materials") {
    groups.resize(lattice.materials.size());
    for (int i = 0; i < N; ++i) groups[lattice.material_of_site(i)].push_back(i);
    for (const auto &m : lattice.materials) group_names_.push_back(m.name);
  } else if (grouping == "positions") {
    groups.resize(lattice.M);
    for (int i = 0; i < N; ++i) groups[i % lattice.M].push_back(i);
    for (int n = 0; n < lattice.M; ++n) group_names_.push_back(std::to_string(n));
  } else {
    throw std::runtime_error("unknown magnetisation grouping: " + grouping);
  }
  mus_ = lattice.mus();
  const std::vector<double> pos = lattice.positions();
  const Mat3 R = rotation_matrix_between_vectors(layer_normal, Vec3{{0.0, 0.0, 1.0}});
  const double to_nm = lattice.lattice_parameter * 1e9;   // kMeterToNanometer
  group_layer_spins_.resize(groups.size());
  layer_positions_.resize(groups.size());
  tsv_file_ << "# magnetisation-layers: layer_normal " << layer_normal[0] << " " << layer_normal[1] << " " << layer_normal[2]
            << " layer_thickness_nm " << layer_thickness << "\n";
  for (size_t g = 0; g < groups.size(); ++g) {
    std::vector<double> z(N, 0.0);
    double z_min = std::numeric_limits<double>::max();
    for (int i : groups[g]) {
      z[i] = (R[2][0] * pos[3 * i] + R[2][1] * pos[3 * i + 1] + R[2][2] * pos[3 * i + 2]) * to_nm;
      z_min = std::min(z_min, z[i]);
    }
    auto comp_less = [&](double a, double b) -> bool {
      if (layer_thickness == 0.0) return definately_less_than(a, b, distance_tolerance);
      return definately_less_than(std::floor((a - z_min) / layer_thickness), std::floor((b - z_min) / layer_thickness), distance_tolerance);
    };
    std::map<double, std::vector<int>, decltype(comp_less)> unique_positions(comp_less);
    for (int i : groups[g]) unique_positions[z[i]].push_back(i);
    int counter = 0;
    for (const auto &kv : unique_positions) {
      double z_layer = kv.first;
      if (layer_thickness != 0.0) {
        const double bin_index = std::floor((kv.first - (z_min - 0.5 * layer_thickness)) / layer_thickness);
        z_layer = z_min + (bin_index + 0.5) * layer_thickness;
      }
      double sat = 0.0;
      for (int i : kv.second) sat += mus_[i] / kBohrMagnetonIU;
      layer_positions_[g].push_back(z_layer);
      group_layer_spins_[g].push_back(kv.second);
      tsv_file_ << "# group " << group_names_[g] << " layer " << counter << " position_nm " << std::setprecision(12) << z_layer
                << " saturation_moment_muB " << sat << " spins " << kv.second.size() << "\n";
      ++counter;
    }
  }
  tsv_file_ << "# iteration time_ps group layer mx my mz (Bohr magnetons)" << std::endl;
}

void MagnetisationLayersMonitor::update(B200HeunLLGSolver &solver) {   // monitors/magnetisation_layers.cc:208-242
  const std::vector<double> s = solver.spins();
  for (size_t g = 0; g < group_layer_spins_.size(); ++g) {
    for (size_t l = 0; l < group_layer_spins_[g].size(); ++l) {
      double m[3] = {0.0, 0.0, 0.0};   // jams::sum_spins_moments (helpers/spinops.cc:55-67)
      for (int i : group_layer_spins_[g][l]) for (int n = 0; n < 3; ++n) m[n] += mus_[i] * s[3 * static_cast<size_t>(i) + n];
      tsv_file_ << solver.iteration() << " " << std::scientific << std::setprecision(9) << solver.time() << " " << group_names_[g] << " " << l
                << std::setprecision(15) << " " << m[0] / kBohrMagnetonIU << " " << m[1] / kBohrMagnetonIU << " " << m[2] / kBohrMagnetonIU << "\n";
    }
  }
  tsv_file_.flush();
}

EnergyMonitor::EnergyMonitor(const Setting &settings, const std::string &filename) : Monitor(settings), filename_(filename), tsv_file_(filename) {
  if (!tsv_file_) throw std::runtime_error("cannot open " + filename);
  tsv_file_.setf(std::ios::right);
}

void EnergyMonitor::update(B200HeunLLGSolver &solver) {   // monitors/energy.cc:22-46
  if (!header_written_) {   // the reference writes the header in the constructor from globals::solver->hamiltonians()
    std::stringstream ss;
    ss.width(12);
    ss << "time\t";
    for (auto &h : solver.hamiltonians()) ss << h->name() << "_E_meV\t";
    ss << std::endl;
    tsv_file_ << ss.str();
    header_written_ = true;
  }
  tsv_file_.width(12);
  tsv_file_ << std::scientific << solver.time() << "\t";
  for (auto &h : solver.hamiltonians()) tsv_file_ << std::scientific << std::setprecision(15) << h->calculate_total_energy(solver.time()) << "\t";
  tsv_file_ << std::endl;
}

// =====================================================================================================
// Solver
// =====================================================================================================
B200HeunLLGSolver::B200HeunLLGSolver(const Setting &settings, const Lattice &lattice, uint64_t seed) : lattice_(lattice), seed_(seed) {
  // solvers/cuda_llg_heun.cu:21-37 / solvers/cuda_rk4_base.cu:10-29 (same keys)
  rk4_ = lowercase(settings.required("module").as_string()).find("rk4") != std::string::npos;
  step_size_ = settings.required("t_step").as_double() / 1e-12;
  const double t_max = settings.required("t_max").as_double() / 1e-12;
  const double t_min = settings.get("t_min", 0.0) / 1e-12;
  max_steps_ = static_cast<int>(t_max / step_size_);
  min_steps_ = static_cast<int>(t_min / step_size_);
  jb_lattice_desc d{};
  for (int n = 0; n < 3; ++n) { d.dims[n] = lattice.dims[n]; d.periodic[n] = lattice.periodic[n] ? 1 : 0; }
  d.num_motif = lattice.M; d.x_begin = 0; d.nx_local = lattice.dims[0]; d.rank = 0; d.n_ranks = 1;
  d.device = settings.get("device", -1);
  if (jb_create(&ctx_, &d) != JB_OK) throw std::runtime_error(std::string("jams_b200: ") + jb_last_error(nullptr));
  spins0_ = lattice.initial_spins(seed);
  noise_step_offset_ = static_cast<uint64_t>(std::max<long long>(0, lattice.snapshot_iteration));   // resumed run: continue the noise stream
  physics_.reset(new Physics(nullptr));
}

B200HeunLLGSolver::~B200HeunLLGSolver() { jb_destroy(ctx_); }

void B200HeunLLGSolver::check(int status) const {
  if (status != JB_OK) throw std::runtime_error(std::string("jams_b200: ") + jb_last_error(ctx_));
}

void B200HeunLLGSolver::register_hamiltonian(Hamiltonian *h) {
  // The reference sums any number of Hamiltonians (core/solver.cc:43-57).  The fused kernels hold one bilinear exchange list, one
  // biquadratic list, one Zeeman field and one applied field: a second one of those is refused, not dropped.  Uniaxial terms go
  // into up to three slots (K1 + K2 + K3 as separate modules, jb_set_uniaxial_term).
  if (auto *uni = dynamic_cast<UniaxialAnisotropyHamiltonian *>(h)) {
    int n = 0;
    for (const auto &other : hamiltonians_) n += dynamic_cast<UniaxialAnisotropyHamiltonian *>(other.get()) != nullptr;
    if (n >= 3) { delete h; throw std::runtime_error(name() + ": more than three uniaxial hamiltonians (use the reference's solver)"); }
    uni->set_slot(n);
  } else {
    for (const auto &other : hamiltonians_)
      if (other->term() == h->term()) {
        const std::string msg = name() + ": hamiltonians '" + other->name() + "' and '" + h->name() +
                                "' are the same kind of term; the fused solver holds one of each kind (merge them, or use the reference's solver)";
        delete h;
        throw std::runtime_error(msg);
      }
  }
  h->solver = this;
  hamiltonians_.emplace_back(h);
}

void B200HeunLLGSolver::build() {   // Hamiltonians are registered after the solver exists (core/jams++.cc:274-288)
  if (built_) return;
  const std::vector<double> mus = lattice_.mus(), gyro = lattice_.gyro(), alpha = lattice_.alpha();
  check(jb_set_materials(ctx_, mus.data(), gyro.data(), alpha.data()));
  for (auto &h : hamiltonians_) h->attach(ctx_);
  check(jb_import_spins(ctx_, spins0_.data(), 0));
  built_ = true;
}

void B200HeunLLGSolver::set_spins(const std::vector<double> &s) {
  if (s.size() != 3 * static_cast<size_t>(lattice_.num_spins)) throw std::runtime_error("set_spins: wrong array size");
  spins0_ = s;
  if (built_) check(jb_import_spins(ctx_, spins0_.data(), 0));
}

std::vector<double> B200HeunLLGSolver::spins() {
  build();
  std::vector<double> s(3 * static_cast<size_t>(lattice_.num_spins));
  check(jb_export_spins(ctx_, s.data(), 0));
  return s;
}

void B200HeunLLGSolver::run_steps(int n) {
  build();
  check((rk4_ ? jb_step_rk4 : jb_step)(ctx_, n, step_size_, time_, physics_->temperature(), seed_, noise_step_offset_ + static_cast<uint64_t>(iteration_),
                                       lattice_.gilbert_prefactor ? 1 : 0));
  iteration_ += n;
  time_ = iteration_ * step_size_;   // solvers/cuda_llg_heun.cu:120-121
}

void B200HeunLLGSolver::run() { run_steps(1); }

void B200HeunLLGSolver::notify_monitors() {
  build();
  for (auto &m : monitors_) if (m->is_updating(iteration_)) m->update(*this);
}

void B200HeunLLGSolver::post_process_monitors() {
  for (auto &m : monitors_) m->post_process();
}

void SpinsTsvMonitor::write(const std::string &filename, const std::vector<double> &s, int iteration, double time) {
  std::ofstream f(filename);
  if (!f) throw std::runtime_error("cannot open " + filename);
  f << "# spins " << s.size() / 3 << " x 3   iteration " << iteration << "   time_ps " << std::setprecision(17) << time << "\n";
  f << std::setprecision(17);
  for (size_t i = 0; i + 2 < s.size(); i += 3) f << s[i] << ' ' << s[i + 1] << ' ' << s[i + 2] << '\n';
}

void SpinsTsvMonitor::update(B200HeunLLGSolver &solver) {   // monitors/hdf5.cc:64-78: <name>_NNNNNNN
  last_ = solver.spins(); last_iteration_ = solver.iteration(); last_time_ = solver.time();
  char num[16];
  std::snprintf(num, sizeof(num), "%07d", last_iteration_);
  write(prefix_ + num + ".tsv", last_, last_iteration_, last_time_);
  final_from_ = &solver;
}

void SpinsTsvMonitor::post_process() {   // monitors/hdf5.cc:52: <name>_final
  if (!final_from_) return;
  write(prefix_ + "final.tsv", final_from_->spins(), final_from_->iteration(), final_from_->time());
}

std::vector<double> B200HeunLLGSolver::compute_fields() {
  build();
  std::vector<double> h(3 * static_cast<size_t>(lattice_.num_spins));
  check(jb_fields(ctx_, JB_TERM_TOTAL, time_, h.data(), 0));
  return h;
}

std::vector<double> Hamiltonian::calculate_fields(double time) {
  std::vector<double> h(3 * static_cast<size_t>(lattice_.num_spins));
  solver->check(jb_fields(solver->ctx(), term(), time, h.data(), 0));
  return h;
}
std::vector<double> Hamiltonian::calculate_energies(double time) {
  std::vector<double> e(lattice_.num_spins);
  double total = 0.0;
  solver->check(jb_energies(solver->ctx(), term(), time, e.data(), 0, &total));
  return e;
}
double Hamiltonian::calculate_total_energy(double time) {
  double total = 0.0;
  solver->check(jb_energies(solver->ctx(), term(), time, nullptr, 0, &total));
  return total;
}

void ExchangeHamiltonian::attach(jb_ctx *ctx) {
  // check_sparse_matrix_symmetry = false switches the symmetry check off (hamiltonian/exchange.cc:104-110); default: checked
  solver->check(jb_set_option(ctx, "check_symmetry", check_symmetry_ ? 1 : 0));
  if (lattice_.has_impurities) {   // not translation invariant: the neighbour list itself (general kernel)
    const NeighbourList nl = lattice_.neighbour_list(template_);
    solver->check(jb_set_exchange_pairs(ctx, static_cast<int64_t>(nl.i.size()), nl.i.data(), nl.j.data(), nl.value_id.data(),
                                        static_cast<int32_t>(nl.values9.size() / 9), nl.values9.data()));
    return;
  }
  solver->check(jb_set_exchange_template(ctx, template_.size(), template_.mi.data(), template_.mj.data(), template_.T3.data(), template_.J9.data()));
}
// hamiltonian/cuda_biquadratic_exchange.cu:9-156: the exchange grammar, scalar B = J[0][0] * unit (no interaction_prefactor), only
// values above energy_cutoff are inserted (:131)
BiquadraticExchangeHamiltonian::BiquadraticExchangeHamiltonian(const Setting &s, const Lattice &lattice) : ExchangeHamiltonian(s, lattice, NoParse{}) {
  if (lattice.has_impurities) throw std::runtime_error("biquadratic-exchange is not supported on a lattice with impurities by the llg-heun-b200-gpu host layer");   // template form only
  parse_interactions(s, lattice, 1.0);
  const double cutoff = s.get("energy_cutoff", 0.0) * input_energy_unit_conversion_;
  InteractionTemplate kept;
  for (int k = 0; k < template_.size(); ++k) {
    const double B = template_.J9[9 * static_cast<size_t>(k)];
    if (!(B > cutoff)) continue;
    kept.mi.push_back(template_.mi[k]); kept.mj.push_back(template_.mj[k]);
    for (int d = 0; d < 3; ++d) kept.T3.push_back(template_.T3[3 * static_cast<size_t>(k) + d]);
    B_.push_back(B);
  }
  kept.J9.clear();
  template_ = kept;
}
void BiquadraticExchangeHamiltonian::attach(jb_ctx *ctx) {
  solver->check(jb_set_option(ctx, "check_symmetry", check_symmetry_ ? 1 : 0));
  solver->check(jb_set_biquadratic_template(ctx, static_cast<int32_t>(B_.size()), template_.mi.data(), template_.mj.data(), template_.T3.data(), B_.data()));
}
void UniaxialAnisotropyHamiltonian::attach(jb_ctx *ctx) { solver->check(jb_set_uniaxial_term(ctx, slot_, power_, magnitude_.data(), axis_.data())); }
void ZeemanHamiltonian::attach(jb_ctx *ctx) {
  solver->check(jb_set_zeeman(ctx, dc_local_field_.data(), has_ac_local_field_ ? ac_local_field_.data() : nullptr,
                              has_ac_local_field_ ? ac_local_frequency_.data() : nullptr));
}
void AppliedFieldHamiltonian::attach(jb_ctx *ctx) {
  if (type_ == JB_FIELD_STATIC) solver->check(jb_set_applied_field(ctx, field_.data(), 1));
  else solver->check(jb_set_applied_field_pulse(ctx, field_.data(), type_, time_center_, freq_bandwidth_, freq_center_));
}

// =====================================================================================================
// Simulation (core/jams++.cc:231-377)
// =====================================================================================================
Simulation::Simulation(const std::vector<std::string> &config_args, const std::string &name, const std::string &output_dir) {
  config_ = parse_config_strings(config_args);
  prefix_ = output_dir;
  if (!prefix_.empty() && prefix_.back() != '/') prefix_ += '/';
  prefix_ += name + "_";   // jams::output::full_path_filename (helpers/output.cc:38-41)
  uint64_t seed = 0;
  if (const Setting *sim = config_->find("sim")) seed = static_cast<uint64_t>(sim->get("seed", 0LL));
  lattice_.reset(new Lattice(*config_));
  const Setting &solver_settings = config_->required("solver");
  const std::string module = lowercase(solver_settings.required("module").as_string());
  if (module != "llg-heun-b200-gpu" && module != "llg-heun-gpu" && module != "llg-heun-cpu" &&
      module != "llg-rk4-b200-gpu" && module != "llg-rk4-gpu")   // Solver::create (core/solver.cc:60-77)
    throw std::runtime_error("unknown solver " + solver_settings["module"].as_string() + " (this host layer provides the llg-heun and llg-rk4 paths only)");
  solver_.reset(new B200HeunLLGSolver(solver_settings, *lattice_, seed));
  solver_->register_physics_module(new Physics(config_->find("physics"), config_->find("sim"), prefix_));   // Physics::create (core/physics.cc:79-126): empty, pinned_boundaries, field-cool, two-temperature-model
  if (!config_->exists("hamiltonians")) throw std::runtime_error("No hamiltonians group in config");
  const Setting &hams = (*config_)["hamiltonians"];
  for (int i = 0; i < hams.length(); ++i) solver_->register_hamiltonian(Hamiltonian::create(hams[i], *lattice_));
  if (const Setting *mons = config_->find("monitors"))
    for (int i = 0; i < mons->length(); ++i) solver_->register_monitor(Monitor::create((*mons)[i], *lattice_, prefix_));
  if (const Setting *init = config_->find("initializer")) run_initializer(*init);
}

void Simulation::run_initializer(const Setting &s) {   // initializer/init_dispatcher + init_bloch_domain_wall.cc:10-32
  const std::string module = lowercase(s.required("module").as_string());
  if (module != "bloch_domain_wall") throw std::runtime_error("unknown initializer " + module);
  const double width = s.required("width").as_double(), center = s.required("center").as_double();
  Vec3 normal{{1, 0, 0}}, domain{{0, 0, 1}};
  if (const Setting *n = s.find("normal")) normal = read_vec3(*n);
  if (const Setting *d = s.find("domain")) domain = read_vec3(*d);
  { const double n = norm(normal); for (auto &x : normal) x /= n; }
  { const double n = norm(domain); for (auto &x : domain) x /= n; }
  std::vector<double> spins = lattice_->initial_spins(0);
  const std::vector<double> pos = lattice_->positions();
  for (int i = 0; i < lattice_->num_spins; ++i) {
    const Vec3 r = {{pos[3 * static_cast<size_t>(i)], pos[3 * static_cast<size_t>(i) + 1], pos[3 * static_cast<size_t>(i) + 2]}};
    const double x = dot(r, normal) - center;
    const Vec3 m = {{0.0, 1.0 / std::cosh(kPi * x / width), std::tanh(kPi * x / width)}};
    const Vec3 spin = {{spins[3 * static_cast<size_t>(i)], spins[3 * static_cast<size_t>(i) + 1], spins[3 * static_cast<size_t>(i) + 2]}};
    const Vec3 out = matvec(rotation_matrix_between_vectors(domain, m), spin);
    for (int n = 0; n < 3; ++n) spins[3 * static_cast<size_t>(i) + n] = out[n];
  }
  solver_->set_spins(spins);
}

void Simulation::run() {   // run_simulation (core/jams++.cc:326-377)
  while (solver_->is_running()) {
    solver_->update_physics_module();
    solver_->notify_monitors();
    solver_->run();
  }
  solver_->post_process_monitors();
}

}  // namespace jams_b200
