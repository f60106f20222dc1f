// jb_capi.cu — host side of the C ABI declared in include/jams_b200.h.
// Owns the context (device buffers, tables, TMA descriptors, halo peers) and sequences the kernels of
// jb_kernels.cu.  No CPU compute path exists here: without a CUDA device every entry point fails.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <map>
#include <unordered_map>
#include <string>
#include <tuple>
#include <vector>

#include "jb_internal.h"

namespace {

thread_local std::string g_create_error;

constexpr double kBoltzmannIU = 0.0861733326;  // meV/K, reference helpers/consts.h:32

#define JB_FAIL(ctx, code, msg)          \
  do {                                   \
    (ctx)->err = (msg);                  \
    return (code);                       \
  } while (0)

#define JB_CUDA(ctx, call)                                                                       \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                     \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorName(e_) + " - " + cudaGetErrorString(e_); \
      return JB_ERR_CUDA;                                                                        \
    }                                                                                            \
  } while (0)

struct Blob {  // contents of a JB_HALO_HANDLE_BYTES halo handle
  uint32_t magic;
  int32_t pid;
  int32_t device;
  int32_t rank;
  int32_t nx, PY, PZ, M, gx;
  uint64_t base_ptr;             // slab base in the exporting process
  uint64_t off_S0[3], off_S1[3], off_V[3]; // byte offsets inside the slab
  uint64_t off_flags;
  cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(Blob) <= JB_HALO_HANDLE_BYTES, "halo blob too large");

void free_dev(void *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

void release_state(jb_ctx *c) {
  if (c->slab) cudaFree(c->slab);
  c->slab = nullptr;
  for (int k = 0; k < 3; ++k) {
    c->S0[k] = c->S1[k] = nullptr;
    if (c->U[k]) cudaFree(c->U[k]);
    c->U[k] = nullptr;
    c->V[k] = nullptr;   // part of the slab
  }
  c->flags = nullptr;
  c->state_allocated = false;
  c->tmap_valid = false;
}

int ensure_scratch(jb_ctx *c, size_t bytes) {
  if (c->d_scratch_bytes >= bytes) return JB_OK;
  if (c->d_scratch) cudaFree(c->d_scratch);
  c->d_scratch = nullptr; c->d_scratch_bytes = 0;
  JB_CUDA(c, cudaMalloc(&c->d_scratch, bytes));
  c->d_scratch_bytes = bytes;
  return JB_OK;
}

int ensure_aos(jb_ctx *c) {
  const size_t bytes = (size_t)c->N * 3 * sizeof(double);
  if (c->d_aos_bytes >= bytes) return JB_OK;
  if (c->d_aos) cudaFree(c->d_aos);
  c->d_aos = nullptr; c->d_aos_bytes = 0;
  JB_CUDA(c, cudaMalloc(&c->d_aos, bytes));
  c->d_aos_bytes = bytes;
  return JB_OK;
}

int ensure_copy_stream(jb_ctx *c) {
  if (c->copy_stream) return JB_OK;
  JB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k <= JB_COPY_CHUNKS; ++k) JB_CUDA(c, cudaEventCreateWithFlags(&c->copy_ev[k], cudaEventDisableTiming));
  return JB_OK;
}

// work-queue counters + face counters of the stage kernel, and the optional per-CTA trace buffer
int ensure_queue(jb_ctx *c) {
  if (!c->d_queue) {
    JB_CUDA(c, cudaMalloc(&c->d_queue, 8 * sizeof(unsigned int)));
    JB_CUDA(c, cudaMemsetAsync(c->d_queue, 0, 8 * sizeof(unsigned int), c->stream));
    c->stage_launches = 0;
  }
  if (c->opt_trace && !c->d_trace) {
    JB_CUDA(c, cudaMalloc(&c->d_trace, 4096 * JB_TRACE_WORDS * sizeof(unsigned long long)));
    JB_CUDA(c, cudaMemsetAsync(c->d_trace, 0, 4096 * JB_TRACE_WORDS * sizeof(unsigned long long), c->stream));
  }
  return JB_OK;
}

// ---- geometry --------------------------------------------------------------------------------------
void compute_geometry(jb_ctx *c, int gx, int gy, int gz) {
  JbGeom &g = c->g;
  g.nx = c->d.nx_local; g.Ny = c->d.dims[1]; g.Nz = c->d.dims[2]; g.M = c->d.num_motif;
  g.gx = gx; g.gy = gy; g.gz = gz;
  g.PX = g.nx + 2 * gx; g.PY = g.Ny + 2 * gy;
  // row layout: [oz pad/ghost columns][Nz interior][gz ghost + pad up to a multiple of `oz`].  oz = 16 starts every interior run on a
  // 128-byte line; smaller values (8: 64-byte DRAM atoms, 4: 32-byte sectors) shorten the unused gap between the runs
  // of consecutive rows, which is what the DRAM efficiency of the stage kernels turned out to depend on (profiles/README.md)
  int oz = (c->opt_oz == 4 || c->opt_oz == 8 || c->opt_oz == 16) ? c->opt_oz : 16;
  const int need = std::max(4, ((2 * gz + 3) / 4) * 4);   // TMA boxes start up to 2 * ceil_even(gz) columns left of the interior
  while (oz < need) oz *= 2;
  g.oz = oz;
  g.PZ = (g.oz + g.Nz + gz + oz - 1) / oz * oz;
  g.sY = (long long)g.M * g.PZ;
  g.sX = (long long)g.PY * g.sY;
  g.elems = (long long)g.PX * g.sX;
  for (int k = 0; k < 3; ++k) g.per[k] = c->d.periodic[k];
  g.x_begin = c->d.x_begin; g.Nx_global = c->d.dims[0];
  g.n_ranks = c->d.n_ranks; g.rank = c->d.rank;
}

int allocate_state(jb_ctx *c) {
  release_state(c);
  const JbGeom &g = c->g;
  if (g.elems >= (1ll << 31)) JB_FAIL(c, JB_ERR_UNSUPPORTED, "slab too large for 32-bit in-box offsets; use more ranks");
  const size_t comp = ((size_t)g.elems * sizeof(double) + 255) / 256 * 256;
  const size_t total = 9 * comp + 256;   // S0, S1, V (every box a neighbour slab stores into) + the halo flags: one IPC handle
  JB_CUDA(c, cudaMalloc(&c->slab, total));
  c->slab_bytes = total;
  JB_CUDA(c, cudaMemsetAsync(c->slab, 0, total, c->stream));
  char *base = static_cast<char *>(c->slab);
  for (int k = 0; k < 3; ++k) {
    c->S0[k] = reinterpret_cast<double *>(base + (size_t)k * comp);
    c->S1[k] = reinterpret_cast<double *>(base + (size_t)(3 + k) * comp);
    c->V[k] = reinterpret_cast<double *>(base + (size_t)(6 + k) * comp);
    JB_CUDA(c, cudaMalloc(&c->U[k], comp));
    JB_CUDA(c, cudaMemsetAsync(c->U[k], 0, comp, c->stream));
  }
  c->flags = reinterpret_cast<unsigned long long *>(base + 9 * comp);
  c->state_allocated = true;
  c->tmap_valid = false;
  c->halo_connected = false;
  c->epoch = 0;
  return JB_OK;
}

// ---- exchange template -> device tables -------------------------------------------------------------
int build_template_tables(jb_ctx *c) {
  const JbGeom &g = c->g;
  const int n = (int)c->t_mi.size();
  const int M = g.M;
  // unique tensors
  std::vector<std::array<double, 9>> uniq;
  std::vector<int> jidx(n);
  bool iso = true;
  for (int k = 0; k < n; ++k) {
    std::array<double, 9> J;
    std::copy(c->t_J9.begin() + 9 * k, c->t_J9.begin() + 9 * k + 9, J.begin());
    if (!(J[1] == 0 && J[2] == 0 && J[3] == 0 && J[5] == 0 && J[6] == 0 && J[7] == 0 && J[0] == J[4] && J[4] == J[8])) iso = false;
    auto it = std::find(uniq.begin(), uniq.end(), J);
    if (it == uniq.end()) { uniq.push_back(J); jidx[k] = (int)uniq.size() - 1; }
    else jidx[k] = (int)(it - uniq.begin());
  }
  // per motif, ordered like the reference's CSR columns for an interior site: ascending neighbour
  // site id = lexicographic (Tx, Ty, Tz, mj)  (interface/sparse_blas.h:22-25 sums in that order)
  std::vector<int> order(n);
  for (int k = 0; k < n; ++k) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (c->t_mi[a] != c->t_mi[b]) return c->t_mi[a] < c->t_mi[b];
    for (int d = 0; d < 3; ++d) if (c->t_T[3 * a + d] != c->t_T[3 * b + d]) return c->t_T[3 * a + d] < c->t_T[3 * b + d];
    return c->t_mj[a] < c->t_mj[b];
  });
  std::vector<JbNbr> glob(n);
  c->tile_order.assign(order.begin(), order.end());
  c->tile_jidx = jidx;
  for (int m = 0; m <= M; ++m) c->nbr_begin[m] = 0;
  for (int pos = 0; pos < n; ++pos) {
    const int k = order[pos];
    const int mi = c->t_mi[k], mj = c->t_mj[k];
    const int Tx = c->t_T[3 * k], Ty = c->t_T[3 * k + 1], Tz = c->t_T[3 * k + 2];
    if (pos > 0) {
      const int q = order[pos - 1];
      if (c->t_mi[q] == mi && c->t_mj[q] == mj && c->t_T[3 * q] == Tx && c->t_T[3 * q + 1] == Ty && c->t_T[3 * q + 2] == Tz)
        JB_FAIL(c, JB_ERR_INVALID, "Multiple interactions for the same motif pair and translation in the exchange template");
    }
    JbNbr e{};
    e.dx = Tx; e.jidx = jidx[k]; e.J = c->t_J9[9 * k];
    e.delta = (Ty * M + (mj - mi)) * g.PZ + Tz;
    glob[pos] = e;
    c->nbr_begin[mi + 1]++;
  }
  for (int m = 0; m < M; ++m) c->nbr_begin[m + 1] += c->nbr_begin[m];
  c->iso = iso;
  c->n_unique_J = (int)uniq.size();
  if (c->d_nbr_global) cudaFree(c->d_nbr_global);
  if (c->d_Jtab) cudaFree(c->d_Jtab);
  c->d_nbr_global = nullptr; c->d_Jtab = nullptr;
  if (n > 0) {
    JB_CUDA(c, cudaMalloc(&c->d_nbr_global, n * sizeof(JbNbr)));
    JB_CUDA(c, cudaMalloc(&c->d_Jtab, uniq.size() * 9 * sizeof(double)));
    JB_CUDA(c, cudaMemcpy(c->d_nbr_global, glob.data(), n * sizeof(JbNbr), cudaMemcpyHostToDevice));
    JB_CUDA(c, cudaMemcpy(c->d_Jtab, uniq.data(), uniq.size() * 9 * sizeof(double), cudaMemcpyHostToDevice));
  }
  c->tables_built = true;
  c->tiling_valid = false;
  return JB_OK;
}

// ---- per-site parameters -> classes ------------------------------------------------------------------
int build_classes(jb_ctx *c) {
  if (!c->classes_dirty) return JB_OK;
  const JbGeom &g = c->g;
  const int N = c->N;
  if ((int)c->h_mus.size() != N) JB_FAIL(c, JB_ERR_INVALID, "jb_set_materials has not been called");
  std::map<std::array<double, 24>, int> seen;
  c->h_classes.clear(); c->h_class_dc.clear(); c->h_class_ac.clear(); c->h_class_omega.clear(); c->h_class_gyro.clear();
  c->h_uni_extra.clear();
  std::vector<unsigned char> cls(N);
  for (int i = 0; i < N; ++i) {
    std::array<double, 24> key{};
    key[0] = c->h_mus[i]; key[1] = c->h_gyro[i]; key[2] = c->h_alpha[i];
    if (c->uni_power) { key[3] = c->h_K[i]; key[4] = c->h_axis[3 * i]; key[5] = c->h_axis[3 * i + 1]; key[6] = c->h_axis[3 * i + 2]; }
    if (c->has_zeeman) { key[7] = c->h_dc[3 * i]; key[8] = c->h_dc[3 * i + 1]; key[9] = c->h_dc[3 * i + 2]; }
    if (c->has_ac) { key[10] = c->h_ac[3 * i]; key[11] = c->h_ac[3 * i + 1]; key[12] = c->h_ac[3 * i + 2]; key[13] = c->h_omega[i]; }
    for (int q = 0; q < JB_MAX_UNIAXIAL - 1; ++q)
      if (c->uni_powerx[q]) { key[14 + 4 * q] = c->h_Kx[q][i]; for (int d = 0; d < 3; ++d) key[15 + 4 * q + d] = c->h_axisx[q][3 * i + d]; }
    auto it = seen.find(key);
    int id;
    if (it == seen.end()) {
      id = (int)c->h_classes.size();
      if (id >= JB_MAX_CLASSES) JB_FAIL(c, JB_ERR_UNSUPPORTED, "more than JB_MAX_CLASSES distinct per-site parameter sets");
      seen.emplace(key, id);
      JbClass k{};
      k.mu = key[0]; k.inv_mu = (key[0] != 0.0) ? 1.0 / key[0] : 0.0;
      k.alpha = key[2];
      c->h_class_gyro.push_back(key[1]);
      k.K = key[3]; k.Kp = key[3] * c->uni_power; k.ax = key[4]; k.ay = key[5]; k.az = key[6];
      k.power = (c->uni_power && key[3] != 0.0) ? c->uni_power : 0;
      c->h_classes.push_back(k);
      for (int q = 0; q < JB_MAX_UNIAXIAL - 1; ++q) {
        JbUniExtra u{};
        u.K = key[14 + 4 * q]; u.Kp = u.K * c->uni_powerx[q]; u.KpT = u.Kp * k.inv_mu;
        u.ax = key[15 + 4 * q]; u.ay = key[16 + 4 * q]; u.az = key[17 + 4 * q];
        u.power = (c->uni_powerx[q] && u.K != 0.0) ? c->uni_powerx[q] : 0;
        c->h_uni_extra.push_back(u);
      }
      for (int d = 0; d < 3; ++d) { c->h_class_dc.push_back(key[7 + d]); c->h_class_ac.push_back(key[10 + d]); }
      c->h_class_omega.push_back(key[13]);
    } else {
      id = it->second;
    }
    cls[i] = (unsigned char)id;
  }
  // motif-uniform?
  c->motif_uniform = true;
  for (int m = 0; m < g.M && m < JB_MAX_MOTIF; ++m) c->class_of_motif[m] = cls[m];
  for (int i = 0; i < N && c->motif_uniform; ++i) if (cls[i] != cls[i % g.M]) c->motif_uniform = false;
  if (c->d_site_class) cudaFree(c->d_site_class);
  c->d_site_class = nullptr;
  if (!c->motif_uniform) {
    // reorder to interior layout order [x][y][m][z]
    std::vector<unsigned char> lay(N);
    long long q = 0;
    for (int x = 0; x < g.nx; ++x) for (int y = 0; y < g.Ny; ++y) for (int m = 0; m < g.M; ++m) for (int z = 0; z < g.Nz; ++z)
      lay[q++] = cls[(((long long)x * g.Ny + y) * g.Nz + z) * g.M + m];
    JB_CUDA(c, cudaMalloc(&c->d_site_class, N));
    JB_CUDA(c, cudaMemcpy(c->d_site_class, lay.data(), N, cudaMemcpyHostToDevice));
  }
  if (c->d_uni_extra) cudaFree(c->d_uni_extra);
  c->d_uni_extra = nullptr;
  if (c->has_uni_extra()) {
    JB_CUDA(c, cudaMalloc(&c->d_uni_extra, c->h_uni_extra.size() * sizeof(JbUniExtra)));
    JB_CUDA(c, cudaMemcpy(c->d_uni_extra, c->h_uni_extra.data(), c->h_uni_extra.size() * sizeof(JbUniExtra), cudaMemcpyHostToDevice));
  }
  c->classes_dirty = false;
  c->class_sig.clear();
  return JB_OK;
}

// does the constant field of a stage depend on the stage's time? (Zeeman ac term, applied-field pulse)
static bool time_dependent(const jb_ctx *c) { return c->has_ac || (c->has_applied && c->applied_type != JB_FIELD_STATIC); }

// g(t) of the applied field (hamiltonian/applied_field.cc:18-19,41-43,70-73; sinc: helpers/maths.h:441-446)
static double applied_amplitude(const jb_ctx *c, double t) {
  if (c->applied_type == JB_FIELD_STATIC) return 1.0;
  const double kPi = 3.14159265358979323846;
  const double x = kPi * c->applied_fbw * (t - c->applied_t0);
  const double sinc = x == 0.0 ? 1.0 : sin(x) / x;
  if (c->applied_type == JB_FIELD_SINC) return sinc;
  return sinc * cos(2.0 * kPi * c->applied_fc * (t - c->applied_t0));
}

// fill sigma and the constant field of `count` consecutive stage tables and upload them.
//   which_f: JB_TERM_TOTAL (zeeman + applied), JB_TERM_ZEEMAN, JB_TERM_APPLIED, or -1 (none)
int upload_classes(jb_ctx *c, const std::vector<double> &times, double dt, double T, int gilbert, int which_f) {
  const int nc = (int)c->h_classes.size();
  const size_t count = times.size();
  // skip the upload (and its host/device synchronisation) when the table on the device is current
  std::vector<double> sig = {dt, T, (double)gilbert, (double)which_f, c->applied_B[0], c->applied_B[1], c->applied_B[2],
                             c->has_applied ? 1.0 : 0.0, (double)nc, (double)c->applied_type, c->applied_t0, c->applied_fbw, c->applied_fc};
  sig.insert(sig.end(), times.begin(), times.end());
  if (c->d_classes && sig == c->class_sig) return JB_OK;
  std::vector<JbClass> tab(nc * count);
  for (size_t s = 0; s < count; ++s) {
    for (int k = 0; k < nc; ++k) {
      JbClass cl = c->h_classes[k];
      const double gyro = c->h_class_gyro[k];
      if (T > 0.0 && dt > 0.0 && cl.mu != 0.0 && gyro != 0.0) {
        double denominator = 1.0;
        if (gilbert) denominator = 1.0 + cl.alpha * cl.alpha;
        // solvers/cpu_llg_heun.cc:35-42 times sqrt(T) (:57-63)
        cl.sigma = sqrt((2.0 * kBoltzmannIU * cl.alpha) / (cl.mu * gyro * dt * denominator)) * sqrt(T);
      } else {
        cl.sigma = 0.0;
      }
      double f[3] = {0, 0, 0};
      if (which_f == JB_TERM_TOTAL || which_f == JB_TERM_ZEEMAN) {
        for (int d = 0; d < 3; ++d) {
          f[d] += c->has_zeeman ? c->h_class_dc[3 * k + d] : 0.0;
          if (c->has_ac) f[d] += c->h_class_ac[3 * k + d] * cos(c->h_class_omega[k] * times[s]);  // zeeman.cc:126-130
        }
      }
      if ((which_f == JB_TERM_TOTAL || which_f == JB_TERM_APPLIED) && c->has_applied) {
        const double amp = applied_amplitude(c, times[s]);
        for (int d = 0; d < 3; ++d) f[d] += cl.mu * (c->applied_B[d] * amp);  // applied_field.cc:146-148
      }
      cl.fx = f[0]; cl.fy = f[1]; cl.fz = f[2];
      cl.fTx = f[0] * cl.inv_mu; cl.fTy = f[1] * cl.inv_mu; cl.fTz = f[2] * cl.inv_mu;
      cl.KpT = cl.Kp * cl.inv_mu;
      cl.c_full = -gyro * dt; cl.c_half = -gyro * (0.5 * dt);
      cl.gyro = gyro;
      tab[s * nc + k] = cl;
    }
  }
  c->h_class_tab = tab;
  const size_t bytes = tab.size() * sizeof(JbClass);
  static_assert(sizeof(JbClass) % 8 == 0, "class size");
  if (c->h_pinned_bytes < bytes) {
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    c->h_pinned = nullptr; c->h_pinned_bytes = 0;
    JB_CUDA(c, cudaMallocHost(&c->h_pinned, bytes));
    c->h_pinned_bytes = bytes;
    if (c->d_classes) cudaFree(c->d_classes);
    c->d_classes = nullptr;
    JB_CUDA(c, cudaMalloc(&c->d_classes, bytes));
  }
  // the pinned buffer may still be in flight from a previous upload on this stream
  JB_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(c->h_pinned, tab.data(), bytes);
  JB_CUDA(c, cudaMemcpyAsync(c->d_classes, c->h_pinned, bytes, cudaMemcpyHostToDevice, c->stream));
  c->class_sig = sig;
  return JB_OK;
}

// biquadratic template -> device table in ghosted-box terms, per motif site in the reference's CSR column order
int build_biquadratic_tables(jb_ctx *c) {
  const JbGeom &g = c->g;
  const int n = (int)c->bq_mi.size(), M = g.M;
  std::vector<int> order(n);
  for (int k = 0; k < n; ++k) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    if (c->bq_mi[a] != c->bq_mi[b]) return c->bq_mi[a] < c->bq_mi[b];
    for (int d = 0; d < 3; ++d) if (c->bq_T[3 * a + d] != c->bq_T[3 * b + d]) return c->bq_T[3 * a + d] < c->bq_T[3 * b + d];
    return c->bq_mj[a] < c->bq_mj[b];
  });
  std::vector<JbNbr> glob(n);
  for (int m = 0; m <= M; ++m) c->bq_begin[m] = 0;
  for (int pos = 0; pos < n; ++pos) {
    const int k = order[pos];
    JbNbr e{};
    e.dx = c->bq_T[3 * k]; e.J = c->bq_B[k];
    e.delta = (c->bq_T[3 * k + 1] * M + (c->bq_mj[k] - c->bq_mi[k])) * g.PZ + c->bq_T[3 * k + 2];
    glob[pos] = e;
    c->bq_begin[c->bq_mi[k] + 1]++;
  }
  for (int m = 0; m < M; ++m) c->bq_begin[m + 1] += c->bq_begin[m];
  if (c->d_bq_global) cudaFree(c->d_bq_global);
  c->d_bq_global = nullptr;
  if (n > 0) {
    JB_CUDA(c, cudaMalloc(&c->d_bq_global, n * sizeof(JbNbr)));
    JB_CUDA(c, cudaMemcpy(c->d_bq_global, glob.data(), n * sizeof(JbNbr), cudaMemcpyHostToDevice));
  }
  c->bq_built = true;
  return JB_OK;
}

void fill_tables(jb_ctx *c, JbTables &t, int class_table_index) {
  t.nbr_global = c->d_nbr_global; t.Jtab = c->d_Jtab;
  t.classes = c->d_classes + (size_t)class_table_index * c->h_classes.size();
  t.site_class = c->d_site_class;
  for (int m = 0; m <= JB_MAX_MOTIF; ++m) t.nbr_begin[m] = (m <= c->g.M && c->has_template) ? c->nbr_begin[m] : 0;
  for (int m = 0; m < JB_MAX_MOTIF; ++m) t.class_of_motif[m] = c->class_of_motif[m];
  t.n_classes = (int)c->h_classes.size();
  t.iso = c->iso ? 1 : 0;
  t.bq_global = c->has_bq ? c->d_bq_global : nullptr;
  for (int m = 0; m <= JB_MAX_MOTIF; ++m) t.bq_begin[m] = (m <= c->g.M && c->has_bq) ? c->bq_begin[m] : 0;
  t.uni_extra = c->has_uni_extra() ? c->d_uni_extra : nullptr;
}

// ---- tiling of the persistent TMA tile kernel ----------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// shape of the tiles; grid size and x-chunking are decided per kernel variant in tile_launch_shape()
static void choose_pair_tiling(jb_ctx *c) {
  const JbGeom &g = c->g;
  jb_ctx::Tiling t;
  int n_nbr = (int)c->t_mi.size();
  if (c->opt_kernel < 1 || c->opt_kernel == 4 || !c->has_template || !c->motif_uniform || g.gx > JB_TILE_MAX_GX || g.M > JB_TILE_MAX_MOTIF || n_nbr > JB_TILE_MAX_NBR)
    return;
  {
    // a thread owns the sites (z, z + 1) of one y row and of the motif sites ms, ms + msplit, ...; consumer threads =
    // ceil(TZ / 2) x TY x msplit <= 256 so that two CTAs share an SM.  Multi-site motifs get short tiles (the ring slots hold M
    // rows per y), so the motif index is spread over threads as far as the 256 allow: bcc 4 x 64 runs 256 consumer threads with
    // msplit = 2 instead of 128 looping over both sites (8 -> 16 consumer warps per SM: the kernel was latency-bound at 15 %
    // active warps, profiles/r02j_c2_T300_ncu_full.txt).  Tile choice: among z extents 128 / 64 / 32 (long contiguous runs along z serve DRAM best:
    // 4 x 128 beats 8 x 64 beats 16 x 32 on C3, profiles/README.md) take, for each, the tallest tile whose rings fit the
    // shared-memory budget, and keep the candidate with the most sites per plane (a shorter z extent only if it brings
    // 1.5 x the sites).  Motifs with several sites multiply the slot size, so they end up with shorter tiles (bcc: 4 x 64)
    // instead of degenerate one-row tiles.
    const size_t budget = c->opt_ctas_per_sm == 1 ? 220 * 1024 : 113 * 1024;   // per CTA
    const int rmin = 2 * g.gx + 2;
    auto shape = [&](int TY, int TZ, jb_ctx::Tiling &q) -> bool {   // fills q; true if the tile fits
      const int HZ = (TZ + 1) / 2;
      q.TY = TY; q.TZ = TZ;
      q.gzb = (g.gz + 1) & ~1;
      q.BY = TY + 2 * g.gy; q.BZ = ((TZ + 1) & ~1) + 2 * q.gzb;
      q.UZ = (TZ + 1) & ~1;
      q.slotS = (q.BY * g.M * q.BZ + 15) / 16 * 16;
      q.slotU = (q.TY * g.M * q.UZ + 15) / 16 * 16;
      q.msplit = 1;
      if (!c->opt_msplit) { for (int d = g.M; d >= 1; --d) if (g.M % d == 0 && HZ * TY * d <= 256) { q.msplit = d; break; } }
      else if (c->opt_msplit > 0 && g.M % c->opt_msplit == 0) q.msplit = c->opt_msplit;
      q.threads = HZ * TY * q.msplit;
      q.RU = c->opt_RU ? c->opt_RU : 2;
      const size_t slot_bytes = (size_t)3 * q.slotS * 8 + (size_t)n_nbr * sizeof(JbTileNbr);   // ring slot + its phase of the entry table
      const size_t u_bytes = (size_t)q.RU * 3 * q.slotU * 8;
      for (int st = 0; st < 2; ++st) {
        const size_t fixed = 512 + (st == 1 ? u_bytes : 0);   // barriers + item ring
        int R = c->opt_R ? c->opt_R : (budget > fixed ? (int)((budget - fixed) / slot_bytes) : 0);
        R = std::min(R, JB_PAIR_MAX_RING);
        // one plane in flight per CTA is the measured optimum: with the stores in the mix, more outstanding plane loads
        // lower the DRAM efficiency (ring 5 / 6: -8 % / -15 %, profiles/README.md r01f)
        if (!c->opt_R) R = std::min(R, rmin);
        q.Rs[st] = R;
        q.smem[st] = fixed + (size_t)R * slot_bytes;
      }
      q.R = q.Rs[1];
      return q.Rs[0] >= rmin && q.Rs[1] >= rmin && q.threads <= 256 && q.smem[0] <= 220 * 1024 && q.smem[1] <= 220 * 1024 &&
             q.BY * g.M <= 256 && q.BZ <= 256 && q.UZ <= 256 && q.TY * g.M <= 256 && q.RU >= 2 && q.RU <= JB_PAIR_MAX_RING;
    };
    bool found = false, shrunk = false;
    long long best_sites = -1;
    const int zc[3] = {128, 64, 32};
    int prev_TZ = -1;
    for (int k = 0; k < 3; ++k) {
      int TZ = c->opt_TZ ? c->opt_TZ : std::min(zc[k], g.Nz);
      TZ = std::max(1, std::min(TZ, g.Nz));
      if (TZ == prev_TZ) continue;   // the lattice is narrower than this candidate: already tried
      prev_TZ = TZ;
      if (TZ < g.Nz && (TZ & 1)) TZ++;   // several z tiles: their first column must stay 16-byte aligned for TMA
      const int HZ = (TZ + 1) / 2;
      int TY = c->opt_TY ? c->opt_TY : std::max(1, 256 / HZ);
      TY = std::max(1, std::min(TY, std::min(g.Ny, 64)));
      jb_ctx::Tiling q = t;
      bool ok = false;
      const int ty_max = TY;
      for (; TY >= 1; TY = c->opt_TY ? 0 : TY - 1) { if (shape(TY, TZ, q)) { ok = true; break; } }
      if (ok) {
        const long long sites = (long long)q.TY * q.TZ * g.M;
        if (2 * sites > 3 * best_sites) { best_sites = sites; shrunk = q.TY < ty_max; t = q; found = true; }   // a shorter z extent must bring 1.5 x the sites per plane
      }
      if (c->opt_TZ) break;   // fixed by the caller
    }
    if (!found) return;
    // a tile that is mostly halo (deep templates: C4 has ghost depth 3) or tiny is slower than the direct kernel (measured
    // on C4: 0.99 ms against 0.62 ms): leave those to the direct gathers unless the caller insists on a tile
    if (!c->opt_TY && !c->opt_TZ && shrunk) {   // only tiles the shared-memory budget cut down; small lattices keep their tile
      const double interior = (double)t.TY * t.TZ / ((double)t.BY * t.BZ);
      if (interior < 0.4) return;
    }
    // Large lattices: half the z extent -- 128 consumer threads per CTA and four CTAs per SM instead of 256 and two.  Twice as many
    // planes in flight per SM and finer work items outweigh the shorter rows once every CTA has enough planes of work (measured,
    // profiles/README.md r02ai: C3 256^3 4 x 128 -> 4 x 64: stage A 0.176 -> 0.169 ms, B 0.212 -> 0.210; C2 128^3 4 x 64 -> 4 x 32:
    // +2 %; on sc 128^3, 14 planes per CTA, the smaller tile loses 2 %).
    if (!c->opt_TY && !c->opt_TZ && !c->opt_ctas_per_sm && t.threads == 256 && t.TZ >= 64 && !(t.TZ & 3) && t.TZ < 2 * g.Nz) {
      jb_ctx::Tiling h = t;
      if (shape(t.TY, t.TZ / 2, h) && h.threads == 128 && std::max(h.smem[0], h.smem[1]) <= 56 * 1024) {
        const int sms = c->num_sms > 0 ? c->num_sms : 148;
        const long long cols = (long long)((g.Ny + h.TY - 1) / h.TY) * ((g.Nz + h.TZ - 1) / h.TZ);
        if ((double)g.nx * cols / (4.0 * sms) >= 24.0) t = h;
      }
    }
  }
  t.n_yt = (g.Ny + t.TY - 1) / t.TY; t.n_zt = (g.Nz + t.TZ - 1) / t.TZ;
  t.n_cols = t.n_yt * t.n_zt;
  t.ok = true;
  c->tiling = t;

  // tile-relative neighbour table, per motif site in the reference's CSR column order; couplings in Tesla
  c->tile_nbr_begin.assign(g.M + 1, 0);
  std::vector<double> J9T(9 * (size_t)std::max(1, n_nbr));
  // pair kernel: within a motif site the entries with an even z offset come first (their neighbour pair is 16-byte
  // aligned in shared memory), then the odd ones; both groups keep the CSR column order
  std::vector<int> order(c->tile_order.begin(), c->tile_order.end());
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (c->t_mi[a] != c->t_mi[b]) return c->t_mi[a] < c->t_mi[b];
    return (c->t_T[3 * a + 2] & 1) < (c->t_T[3 * b + 2] & 1);
  });
  c->tile_nbr_odd.assign(g.M + 1, 0);
  c->tile_zself.assign(2 * (size_t)g.M, 0.0);
  c->tile_nbr.clear();
  for (int pos = 0; pos < n_nbr; ++pos) {
    const int k = order[pos];
    const int mi = c->t_mi[k], mj = c->t_mj[k];
    const int Tx = c->t_T[3 * k], Ty = c->t_T[3 * k + 1], Tz = c->t_T[3 * k + 2];
    const double inv_mu = c->h_classes[c->class_of_motif[mi]].inv_mu;
    if (c->iso && Tx == 0 && Ty == 0 && mi == mj && (Tz == 1 || Tz == -1)) {   // the other site of the pair / its z neighbour: no table entry
      c->tile_zself[2 * mi + (Tz > 0 ? 1 : 0)] = c->t_J9[9 * k] * inv_mu;
      continue;
    }
    if (!(Tz & 1)) c->tile_nbr_odd[mi]++;   // number of even entries for now
    JbTileNbr e{};
    e.delta = (Ty * g.M + (mj - mi)) * t.BZ + Tz;
    e.d = Tx + g.gx;
    e.J = c->t_J9[9 * k] * inv_mu;
    const size_t at = c->tile_nbr.size();
    for (int q = 0; q < 9; ++q) J9T[9 * at + q] = c->t_J9[9 * (size_t)k + q] * inv_mu;
    c->tile_nbr.push_back(e);
    c->tile_nbr_begin[mi + 1]++;
  }
  n_nbr = (int)c->tile_nbr.size();
  for (int q = 0; q < g.M; ++q) c->tile_nbr_begin[q + 1] += c->tile_nbr_begin[q];
  for (int q = 0; q < g.M; ++q) c->tile_nbr_odd[q] += c->tile_nbr_begin[q];   // -> first odd entry
  if (c->d_tile_nbr) cudaFree(c->d_tile_nbr);
  if (c->d_tile_J9T) cudaFree(c->d_tile_J9T);
  c->d_tile_nbr = nullptr; c->d_tile_J9T = nullptr;
  if (cudaMalloc(&c->d_tile_nbr, std::max(1, n_nbr) * sizeof(JbTileNbr)) != cudaSuccess ||
      cudaMemcpy(c->d_tile_nbr, c->tile_nbr.data(), n_nbr * sizeof(JbTileNbr), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMalloc(&c->d_tile_J9T, J9T.size() * sizeof(double)) != cudaSuccess ||
      cudaMemcpy(c->d_tile_J9T, J9T.data(), J9T.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    c->tiling.ok = false;  // fall back to the direct kernel
  }
}

// ---- rows kernel (jb_stage_rows.cu): deep isotropic templates ---------------------------------------------------------
// Segments: the entries of a motif site that share (dx, dz, mj) and differ only in dy, cut into runs of at most five consecutive
// offsets; sorted by motif site and length.  delta needs the row pitch BZ of the tile.
static void build_row_segments(jb_ctx *c, int BZ) {
  const JbGeom &g = c->g;
  const int n_nbr = (int)c->t_mi.size();
  struct Key { int mi, dx, dz, mj; bool operator<(const Key &o) const { return std::tie(mi, dx, dz, mj) < std::tie(o.mi, o.dx, o.dz, o.mj); } };
  std::map<Key, std::map<int, double>> rows;   // -> dy -> J (meV); duplicate entries add up like the CSR builder merges them
  for (int k = 0; k < n_nbr; ++k)
    rows[Key{c->t_mi[k], c->t_T[3 * k], c->t_T[3 * k + 2], c->t_mj[k]}][c->t_T[3 * k + 1]] += c->t_J9[9 * (size_t)k];
  std::vector<std::pair<int, JbRowSeg>> segs;   // (mi, segment)
  for (const auto &kv : rows) {
    const Key &key = kv.first;
    const double inv_mu = c->h_classes[c->class_of_motif[key.mi]].inv_mu;
    auto it = kv.second.begin();
    while (it != kv.second.end()) {
      const int dy0 = it->first;
      JbRowSeg sg{};
      sg.d = key.dx + g.gx;
      sg.delta = ((dy0 * g.M + (key.mj - key.mi)) * BZ + key.dz) * (int)sizeof(double);
      int last = dy0;
      for (; it != kv.second.end() && it->first < dy0 + JB_ROWS_MAX_L; ++it) { sg.c[it->first - dy0] = it->second * inv_mu; last = it->first; }
      sg.L = last - dy0 + 1;
      segs.push_back({key.mi, sg});
    }
  }
  std::stable_sort(segs.begin(), segs.end(), [](const std::pair<int, JbRowSeg> &a, const std::pair<int, JbRowSeg> &b) {
    if (a.first != b.first) return a.first < b.first;
    if (a.second.L != b.second.L) return a.second.L < b.second.L;
    return a.second.d < b.second.d;
  });
  c->row_segs.clear();
  for (int m = 0; m < JB_TILE_MAX_MOTIF; ++m) for (int l = 0; l <= JB_ROWS_MAX_L; ++l) c->row_begin[m][l] = 0;
  for (const auto &ms : segs) c->row_segs.push_back(ms.second);
  size_t at = 0;
  for (int m = 0; m < g.M; ++m) {
    for (int l = 1; l <= JB_ROWS_MAX_L; ++l) {
      c->row_begin[m][l - 1] = (int)at;
      while (at < segs.size() && segs[at].first == m && segs[at].second.L == l) ++at;
      c->row_begin[m][l] = (int)at;
    }
  }
}

static void choose_rows_tiling(jb_ctx *c) {
  const JbGeom &g = c->g;
  if (c->opt_kernel < 1 || !c->has_template || !c->motif_uniform || !c->iso || g.gx > JB_TILE_MAX_GX || g.M > JB_TILE_MAX_MOTIF) return;
  jb_ctx::Tiling t;
  t.rows = true;
  t.TZ = std::min(32, g.Nz);
  if (t.TZ < g.Nz && (t.TZ & 1)) return;
  t.gzb = (g.gz + 1) & ~1;
  t.BZ = ((t.TZ + 1) & ~1) + 2 * t.gzb;
  t.UZ = (t.TZ + 1) & ~1;
  build_row_segments(c, t.BZ);
  const int n_rows = (int)c->row_segs.size();
  const int R = 2 * g.gx + 2;
  const size_t limit = 227 * 1024;
  int best_threads = 0;
  for (int TY = 16; TY >= JB_ROWS_Q; TY -= JB_ROWS_Q) {
    if (c->opt_TY && TY != c->opt_TY) continue;
    if (TY - JB_ROWS_Q >= g.Ny && !c->opt_TY) continue;   // taller than the lattice
    const int BY = TY + 2 * g.gy;
    const int slotS = (BY * g.M * t.BZ + 15) / 16 * 16;
    const size_t smem = (size_t)R * 3 * slotS * 8 + 512 + (size_t)n_rows * sizeof(JbRowSeg);   // rings + barriers / item ring / face counters + segments
    if (smem > limit || BY * g.M > 256 || t.BZ > 256) continue;
    int ms = 0;
    const int max_warps = c->opt_rows_warps > 0 ? std::min(c->opt_rows_warps, JB_ROWS_MAX_WARPS) : JB_ROWS_MAX_WARPS;
    for (int d = g.M; d >= 1; --d) if (g.M % d == 0 && (TY / JB_ROWS_Q) * d <= max_warps) { ms = d; break; }
    if (c->opt_msplit > 0 && g.M % c->opt_msplit == 0 && (TY / JB_ROWS_Q) * c->opt_msplit <= max_warps) ms = c->opt_msplit;
    if (!ms) continue;
    const int threads = 32 * (TY / JB_ROWS_Q) * ms;
    if (threads > best_threads) {
      best_threads = threads;
      t.TY = TY; t.BY = BY; t.slotS = slotS; t.slotU = 16; t.R = R; t.Rs[0] = t.Rs[1] = R; t.RU = 2; t.msplit = ms; t.threads = threads;
      t.smem[0] = t.smem[1] = smem;
    }
  }
  if (!best_threads) return;
  t.n_yt = (g.Ny + t.TY - 1) / t.TY; t.n_zt = (g.Nz + t.TZ - 1) / t.TZ;
  t.n_cols = t.n_yt * t.n_zt;
  if (c->d_rows) cudaFree(c->d_rows);
  c->d_rows = nullptr;
  if (cudaMalloc(&c->d_rows, std::max(1, n_rows) * sizeof(JbRowSeg)) != cudaSuccess ||
      cudaMemcpy(c->d_rows, c->row_segs.data(), n_rows * sizeof(JbRowSeg), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  t.ok = true;
  c->tiling = t;
}

// kernel choice for a translation-invariant template (option kernel: 0 = direct gathers, 2 = default, 4 = rows kernel where it
// applies): isotropic templates that reach three or more cells along an axis, or two with 40 or more neighbours per site, go to the rows kernel (the pair kernel's tiles
// would be mostly halo, and with that many neighbours the gather is bound by shared-memory bandwidth, which the rows kernel
// halves); everything else to the pair kernel; the direct kernel takes what neither can tile
void choose_tiling(jb_ctx *c) {
  if (c->tiling_valid) return;
  c->tiling = jb_ctx::Tiling();
  c->tiling_valid = true;
  c->tmap_valid = false;
  if (c->has_bq) return;   // the biquadratic field needs s_i inside the neighbour loop: direct kernel
  if (c->has_uni_extra()) return;   // a second / third uniaxial term lives in a table only the direct kernels read (JbUniExtra)
  const int reach = std::max(c->g.gx, std::max(c->g.gy, c->g.gz));
  const bool deep = c->iso && c->has_template && (reach >= 3 || (reach >= 2 && c->t_mi.size() >= (size_t)40 * c->g.M));
  if (c->opt_kernel == 4 || deep) choose_rows_tiling(c);
  if (!c->tiling.ok) choose_pair_tiling(c);
  if (!c->tiling.ok && !deep) choose_rows_tiling(c);
}

void fill_tile_params(jb_ctx *c, JbTileParams &p) {
  const jb_ctx::Tiling &t = c->tiling;
  p.g = c->g;
  p.J9T = c->d_tile_J9T;
  p.TY = t.TY; p.TZ = t.TZ; p.UZ = t.UZ; p.BY = t.BY; p.BZ = t.BZ; p.gzb = t.gzb; p.slotS = t.slotS; p.slotU = t.slotU; p.R = t.R; p.RU = t.RU; p.msplit = t.msplit;
  p.n_yt = t.n_yt; p.n_zt = t.n_zt; p.n_cols = t.n_cols;
  if (t.rows) {
    p.rows = c->d_rows; p.n_rows = (int)c->row_segs.size();
    for (int m = 0; m < JB_TILE_MAX_MOTIF; ++m) for (int l = 0; l <= JB_ROWS_MAX_L; ++l) p.row_begin[m][l] = c->row_begin[m][l];
    p.nbr = nullptr; p.n_nbr = 0;
    return;
  }
  for (size_t q = 0; q < c->tile_nbr_begin.size(); ++q) p.nbr_begin[q] = c->tile_nbr_begin[q];
  for (int q = 0; q < c->g.M; ++q) {
    p.nbr_odd[q] = c->tile_nbr_odd[q];
    p.zself[q][0] = c->tile_zself[2 * q]; p.zself[q][1] = c->tile_zself[2 * q + 1];
    p.has_zself[q] = (p.zself[q][0] != 0.0 || p.zself[q][1] != 0.0) ? 1 : 0;
  }
  p.nbr = c->d_tile_nbr;
  p.n_nbr = (int)c->tile_nbr.size();
}

// The x-chunk plan of one launch: `n` chunks (x0, xc) in QUEUE order.  Work items are (chunk, column) pairs handed out by an
// atomic counter, chunk-major, so CTAs that run at the same time work on neighbouring columns of the same x-range and share
// tile halos through the 126 MB L2.  Order: the two face chunks of the slab first (in a slab-decomposed run their ghost-plane
// stores and the epoch flags travel while the interior is computed), then the long chunks, then a taper of ever shorter chunks
// that evens out the finishing times of the resident CTAs.
//
// Measured on C3 (profiles/README.md r02a): an item boundary costs a CTA next to nothing (the co-resident CTA of the SM uses
// the bandwidth meanwhile: 2048 instead of 1152 items changed the mean CTA busy time by 0.2 %), but a launch lasts as long as
// its slowest CTA: the plan is everything.  So the plan is chosen by SIMULATING the queue: list scheduling of the items over G
// CTAs whose speeds differ by a few per cent (as the per-CTA traces show), cost = planes + a small per-item overhead, over a
// family of (long length, taper share, shortest length) candidates; the candidate with the shortest makespan wins.
struct ChunkPlanCandidate { std::vector<std::pair<int, int>> xs; std::vector<int> taper; };

static ChunkPlanCandidate make_candidate(int nx, int gx, int L, int S, int tail_pct) {
  ChunkPlanCandidate cd;
  const int lmin = std::max(gx, 1);
  L = std::max(lmin, std::min(L, nx));
  S = std::max(lmin, std::min(S, L));
  int tail = (int)((long long)nx * tail_pct / 100);
  if (nx - tail < std::max(L, 2 * lmin)) tail = 0;
  // taper: lengths L/2, L/4, ... >= S, every level the same share of the tail planes
  std::vector<int> tl;
  for (int l = L / 2; l >= S && l >= lmin; l /= 2) tl.push_back(l);
  if (tl.empty() && tail > 0) tl.push_back(S);
  std::vector<int> taper_len;
  int used = 0;
  for (size_t k = 0; k < tl.size() && tail > 0; ++k) {
    const int share = (int)((long long)tail * (k + 1) / tl.size()) - used;
    const int cnt = share / tl[k];
    for (int q = 0; q < cnt; ++q) taper_len.push_back(tl[k]);
    used += cnt * tl[k];
  }
  const int body = nx - used;
  const int nl = std::max(1, (int)std::lround((double)body / L));
  // x layout: long chunks, with the taper just before the last long chunk (so both faces of the slab are long chunks)
  int x = 0;
  for (int k = 0; k < nl; ++k) {
    const int len = (int)((long long)(k + 1) * body / nl) - (int)((long long)k * body / nl);
    if (k == nl - 1 && nl > 1) for (int l : taper_len) { cd.xs.push_back({x, l}); cd.taper.push_back(1); x += l; }
    cd.xs.push_back({x, len}); cd.taper.push_back(0); x += len;
    if (k == nl - 1 && nl == 1) for (int l : taper_len) { cd.xs.push_back({x, l}); cd.taper.push_back(1); x += l; }
  }
  return cd;
}

// queue order of a candidate: face chunks, long chunks, taper (longest first)
static std::vector<int> queue_order(const ChunkPlanCandidate &cd, int nx, int gx) {
  const int n = (int)cd.xs.size();
  std::vector<int> order;
  auto face = [&](int k) { return cd.xs[k].first < gx || cd.xs[k].first + cd.xs[k].second > nx - gx; };
  for (int k = 0; k < n; ++k) if (face(k)) order.push_back(k);
  for (int k = 0; k < n; ++k) if (!face(k) && !cd.taper[k]) order.push_back(k);
  std::vector<int> tp;
  for (int k = 0; k < n; ++k) if (!face(k) && cd.taper[k]) tp.push_back(k);
  std::stable_sort(tp.begin(), tp.end(), [&](int a, int b) { return cd.xs[a].second > cd.xs[b].second; });
  order.insert(order.end(), tp.begin(), tp.end());
  return order;
}

static double simulate_queue(const ChunkPlanCandidate &cd, const std::vector<int> &order, int n_cols, int G, double overhead) {
  // CTA g runs at speed 1 + 4 % * (a fixed pseudo-random number in [-1, 1]); a binary heap of finishing times
  std::vector<std::pair<double, int>> heap(G);
  for (int g = 0; g < G; ++g) heap[g] = {0.0, g};
  auto cmp = [](const std::pair<double, int> &a, const std::pair<double, int> &b) { return a.first > b.first; };
  std::make_heap(heap.begin(), heap.end(), cmp);
  auto slow = [](int g) { uint32_t h = (uint32_t)g * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; return 1.0 + 0.04 * ((double)(h & 0xffff) / 32767.5 - 1.0); };
  for (int q : order) {
    const double cost = cd.xs[q].second + overhead;
    for (int col = 0; col < n_cols; ++col) {
      std::pop_heap(heap.begin(), heap.end(), cmp);
      std::pair<double, int> &w = heap.back();
      w.first += cost * slow(w.second);
      std::push_heap(heap.begin(), heap.end(), cmp);
    }
  }
  double end = 0.0;
  for (const auto &w : heap) end = std::max(end, w.first);
  return end;
}

struct ChunkPlanOptions { int chunks = 0, chunk_long = 0, chunk_short = 0, tail_pct = -1, face_after = 0; };

// result in queue order; returns the number of chunks
static int plan_chunks_core(int nx, int gx, int n_cols, int G, const ChunkPlanOptions &o, int *x0_out, int *xc_out) {
  const ChunkPlanOptions *c = &o;
  ChunkPlanCandidate best;
  if (o.chunks > 0) {
    const int nc = std::max(1, std::min(std::min(o.chunks, nx), JB_TILE_MAX_CHUNKS));
    for (int k = 0; k < nc; ++k) {
      const int x0 = (int)((long long)k * nx / nc), x1 = (int)((long long)(k + 1) * nx / nc);
      best.xs.push_back({x0, x1 - x0}); best.taper.push_back(0);
    }
  } else if (c->chunk_long > 0) {
    best = make_candidate(nx, gx, c->chunk_long, c->chunk_short > 0 ? c->chunk_short : std::max(gx, 2), c->tail_pct >= 0 ? c->tail_pct : 25);
  } else if ((double)nx * n_cols / std::max(1, G) < 24.0) {
    // small lattices (less than three 8-plane items per resident CTA; BASELINE config 2 at 64^3: 3.5 planes per CTA): the
    // chunks must be short enough to give every CTA an item -- equal chunks of 1 ... 8 planes, the makespan of the simulated
    // queue decides; an item pays for its 2 gx halo planes (loads only, from L2 at these sizes) and a fixed overhead.
    // Measured on C2 64^3: 8-plane chunks (160 of 296 CTAs busy) 7.9 G spin-updates/s, 4-plane chunks 11.8 G.
    double best_t = 1e300;
    for (int L : {1, 2, 3, 4, 5, 6, 8, 12, 16}) {
      if (L > nx || L < std::max(gx, 1)) continue;
      const int nc = std::max(1, std::min(std::min((nx + L - 1) / L, nx / std::max(gx, 1)), JB_TILE_MAX_CHUNKS));   // no chunk shorter than the ghost depth
      ChunkPlanCandidate cd;
      for (int k = 0; k < nc; ++k) {
        const int x0 = (int)((long long)k * nx / nc), x1 = (int)((long long)(k + 1) * nx / nc);
        cd.xs.push_back({x0, x1 - x0}); cd.taper.push_back(0);
      }
      const double t = simulate_queue(cd, queue_order(cd, nx, gx), n_cols, G, 0.35 + 0.5 * 2 * gx);
      if (t < best_t) { best_t = t; best = cd; }
    }
    if (best.xs.empty()) best = make_candidate(nx, gx, nx, nx, 0);
  } else {
    const double per_cta = (double)nx * n_cols / std::max(1, G);   // planes of work per resident CTA
    double best_t = 1e300;
    const int Ls[] = {8, 10, 12, 14, 16, 20, 24, 28, 32, 40, 48, 64, 96, 128};
    const int tails[] = {0, 10, 20, 30, 40, 50};
    const int Ss[] = {2, 4, 8};
    for (int L : Ls) {
      if (L > nx || (3 * L > per_cta && L != Ls[0])) continue;   // a CTA takes at least three long items: robust against speed differences the model does not know
      for (int tp : tails) for (int S : Ss) {
        if (tp == 0 && S != Ss[0]) continue;
        if (c->tail_pct >= 0 && tp != tails[0]) continue;
        ChunkPlanCandidate cd = make_candidate(nx, gx, L, std::max(S, gx), c->tail_pct >= 0 ? c->tail_pct : tp);
        if ((int)cd.xs.size() > JB_TILE_MAX_CHUNKS || (long long)cd.xs.size() * n_cols > 200000) continue;
        const double t = simulate_queue(cd, queue_order(cd, nx, gx), n_cols, G, 0.35) * (1.0 + 1e-4 * cd.xs.size());
        if (t < best_t) { best_t = t; best = cd; }
      }
    }
    if (best.xs.empty()) best = make_candidate(nx, gx, nx, nx, 0);
  }
  if ((int)best.xs.size() > JB_TILE_MAX_CHUNKS) best = make_candidate(nx, gx, (nx + JB_TILE_MAX_CHUNKS - 1) / JB_TILE_MAX_CHUNKS, nx, 0);
  std::vector<int> order = queue_order(best, nx, gx);
  const int n = (int)best.xs.size();
  if (o.face_after > 0) {
    // slab-decomposed runs: the face chunks are queued after `face_after` interior chunks, so that the CTAs reach their face
    // planes -- 12 KB of P2P stores each -- spread over the finishing times of the first wave instead of all in the same
    // microsecond (a burst of 1.5 MB per direction that stalls every SM's store path at once); late enough to spread, early
    // enough for the epoch flag to reach the neighbour long before its next stage starts
    auto face = [&](int k) { return best.xs[k].first < gx || best.xs[k].first + best.xs[k].second > nx - gx; };
    std::vector<int> f, rest;
    for (int k : order) (face(k) ? f : rest).push_back(k);
    const size_t at = std::min<size_t>(o.face_after, rest.size() / 3);   // never in the last two thirds of the queue: the flag must travel early
    rest.insert(rest.begin() + at, f.begin(), f.end());
    order = rest;
  }
  for (int q = 0; q < n; ++q) { x0_out[q] = best.xs[order[q]].first; xc_out[q] = best.xs[order[q]].second; }
  return n;
}

void plan_chunks(jb_ctx *c, int G, int n_cols, jb_ctx::Tiling::Shape &sh) {
  const int nx = c->g.nx, gx = c->g.gx;
  ChunkPlanOptions o;
  o.chunks = c->opt_chunks; o.chunk_long = c->opt_chunk_long; o.chunk_short = c->opt_chunk_short; o.tail_pct = c->opt_tail_pct; o.face_after = c->opt_face_after >= 0 ? c->opt_face_after : (c->d.n_ranks > 1 ? (G + n_cols - 1) / std::max(1, n_cols) : 0);   // the face items follow the first wave of interior items (measured on 2 GPUs, profiles/README.md r02r, r02ak: one chunk too early or too late costs 0.5-2 %)
  sh.n_chunks = plan_chunks_core(nx, gx, n_cols, G, o, sh.x0, sh.xc);
  sh.face_items[0] = sh.face_items[1] = 0;
  for (int q = 0; q < sh.n_chunks; ++q) {
    if (sh.x0[q] < gx) sh.face_items[0] += n_cols;
    if (sh.x0[q] + sh.xc[q] > nx - gx) sh.face_items[1] += n_cols;
  }
}

// grid size (resident CTAs) and the x-chunk plan for one kernel variant
// rk4: the stage belongs to the RK4 solver (stage 0..3; the shared-memory layout is that of Heun stage 0 for stage 0, else stage 1)
int tile_launch_shape(jb_ctx *c, const JbTileParams &p, int stage, int thermal, int recu, bool rk4 = false) {
  jb_ctx::Tiling &t = c->tiling;
  jb_ctx::Tiling::Shape &sh = rk4 ? t.shape_rk4[stage][thermal] : t.shape[stage][thermal][recu];
  if (sh.grid > 0) return JB_OK;
  if (c->num_sms == 0) JB_CUDA(c, cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device));
  int per_sm = 0;
  if (rk4) JB_CUDA(c, jbk_rk4_stage_pair_occupancy(p, stage, thermal, t.threads, t.smem[stage > 0 ? 1 : 0], &per_sm));
  else if (t.rows) JB_CUDA(c, jbk_stage_rows_occupancy(stage, thermal, t.threads, t.smem[stage], &per_sm));
  else JB_CUDA(c, jbk_stage_pair_occupancy(p, stage, thermal, c->iso ? 1 : 0, recu, t.threads, t.smem[stage], &per_sm));
  if (per_sm < 1) JB_FAIL(c, JB_ERR_CUDA, "the stage kernel does not fit on an SM with this tiling");
  if (c->opt_ctas_per_sm > 0) per_sm = std::min(per_sm, c->opt_ctas_per_sm);
  int G = per_sm * c->num_sms;
  if (c->opt_grid > 0) G = std::min(G, c->opt_grid);   // experiments: fewer resident CTAs than the occupancy calculation allows
  plan_chunks(c, G, t.n_cols, sh);
  sh.grid = (int)std::min<long long>(G, (long long)sh.n_chunks * t.n_cols);
  if (c->opt_verbose) {
    fprintf(stderr, "jams_b200: %s kernel stage %d thermal %d recover_u %d: tile %dx%d (y,z), %d consumer threads (motif split %d), ring %d/%d, smem %zu B, "
                    "%d CTAs/SM -> grid %d, %d x-chunks x %d columns:", t.rows ? "rows" : "pair", stage, thermal, recu, t.TY, t.TZ, t.threads, t.msplit, t.Rs[rk4 ? (stage > 0 ? 1 : 0) : stage], t.RU,
            t.smem[rk4 ? (stage > 0 ? 1 : 0) : stage], per_sm, sh.grid, sh.n_chunks, t.n_cols);
    for (int q = 0; q < sh.n_chunks; ++q) fprintf(stderr, " %d+%d", sh.x0[q], sh.xc[q]);
    fprintf(stderr, "\n");
  }
  return JB_OK;
}

int build_tmaps(jb_ctx *c) {
  if (c->tmap_valid) return JB_OK;
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    JB_CUDA(c, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) JB_FAIL(c, JB_ERR_CUDA, "cuTensorMapEncodeTiled not available in this driver");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const JbGeom &g = c->g;
  const jb_ctx::Tiling &t = c->tiling;
  const cuuint64_t dims[3] = {(cuuint64_t)g.PZ, (cuuint64_t)g.PY * g.M, (cuuint64_t)g.PX};
  const cuuint64_t strides[2] = {(cuuint64_t)g.PZ * 8, (cuuint64_t)g.sX * 8};
  const cuuint32_t estr[3] = {1, 1, 1};
  for (int a = 0; a < 6; ++a) {
    const bool halo_box = a < 2 || a == 5;
    const cuuint32_t box[3] = {(cuuint32_t)(halo_box ? t.BZ : t.UZ), (cuuint32_t)((halo_box ? t.BY : t.TY) * g.M), 1};
    for (int k = 0; k < 3; ++k) {
      double *base = a == 0 || a == 3 ? c->S0[k] : (a == 1 || a == 4 ? c->S1[k] : (a == 5 ? c->V[k] : c->U[k]));
      CUresult r = encode(&c->tmap[a][k], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) JB_FAIL(c, JB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    }
  }
  c->tmap_valid = true;
  return JB_OK;
}

// make geometry, state, tables and classes consistent with the parameters set so far
int ensure_ready(jb_ctx *c) {
  JB_CUDA(c, cudaSetDevice(c->device));
  int gx = 0, gy = 0, gz = 0;
  if (c->has_template) {
    for (size_t k = 0; k < c->t_mi.size(); ++k) {
      gx = std::max(gx, std::abs(c->t_T[3 * k])); gy = std::max(gy, std::abs(c->t_T[3 * k + 1])); gz = std::max(gz, std::abs(c->t_T[3 * k + 2]));
    }
  }
  gx = std::max(gx, c->pairs_reach_x);   // general neighbour list on several ranks (jb_set_exchange_pairs)
  if (c->has_bq) {
    for (size_t k = 0; k < c->bq_mi.size(); ++k) {
      gx = std::max(gx, std::abs(c->bq_T[3 * k])); gy = std::max(gy, std::abs(c->bq_T[3 * k + 1])); gz = std::max(gz, std::abs(c->bq_T[3 * k + 2]));
    }
  }
  const JbGeom &g = c->g;
  const int reach[3] = {gx, gy, gz};
  const bool geom_changed = !c->state_allocated || gx != g.gx || gy != g.gy || gz != g.gz || c->state_relayout;
  c->state_relayout = false;
  if (geom_changed) {
    // the reference throws "Multiple interactions" when a periodic dimension is so short that two
    // template entries reach the same site (core/interactions.cc:373-381); the ghost scheme needs
    // the same condition
    const int ext[3] = {c->d.dims[0], c->d.dims[1], c->d.dims[2]};
    const int gg[3] = {reach[0], reach[1], reach[2]};
    for (int d = 0; d < 3; ++d) {
      if (gg[d] > 0 && c->d.periodic[d] && ext[d] < 2 * gg[d] + 1)
        JB_FAIL(c, JB_ERR_INVALID, "periodic dimension shorter than 2*range+1 of the exchange template (the reference reports 'Multiple interactions' here)");
    }
    if (reach[0] > c->d.nx_local || (c->d.n_ranks == 1 && c->d.periodic[0] && reach[0] > 0 && c->d.nx_local < 2 * reach[0]))
      JB_FAIL(c, JB_ERR_INVALID, "slab thinner than the x range of the exchange template");
    std::vector<double> keep;
    const bool had_state = c->state_allocated;
    if (had_state) {  // re-layout: carry the spins over
      int rc = ensure_aos(c); if (rc) return rc;
      const double *src[3] = {c->S0[0], c->S0[1], c->S0[2]};
      JB_CUDA(c, jbk_export(c->g, src, c->d_aos, c->stream)); c->launches++;
      JB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    compute_geometry(c, gx, gy, gz);
    int rc = allocate_state(c); if (rc) return rc;
    if (had_state) {
      double *dst[3] = {c->S0[0], c->S0[1], c->S0[2]};
      JB_CUDA(c, jbk_import(c->g, c->d_aos, dst, c->d.n_ranks == 1, c->stream)); c->launches++;
    }
    c->tables_built = false;
    c->bq_built = false;
    c->tiling_valid = false;
    c->classes_dirty = true;
  }
  for (int d = 0; d < 3; ++d) c->reach[d] = reach[d];
  const bool classes_were_dirty = c->classes_dirty;
  int rc = build_classes(c); if (rc) return rc;
  if (classes_were_dirty) c->tiling_valid = false;
  if (c->has_template && !c->tables_built) { rc = build_template_tables(c); if (rc) return rc; }
  if (c->has_bq && !c->bq_built) { rc = build_biquadratic_tables(c); if (rc) return rc; }
  return JB_OK;
}

void record_event(jb_ctx *c, int kind) {
  if (!c->opt_time_kernels || !c->time_this_step) return;
  if (c->ev_used >= c->ev.size()) {
    if (c->ev.size() >= 16384) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    c->ev.push_back(e); c->ev_kind.push_back(0);
  }
  c->ev_kind[c->ev_used] = kind;
  cudaEventRecord(c->ev[c->ev_used++], c->stream);
}

}  // namespace

// =====================================================================================================
extern "C" {

int jb_abi_version(void) { return JB_ABI_VERSION; }

const char *jb_last_error(const jb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int jb_create(jb_ctx **out, const jb_lattice_desc *desc) {
  if (!out || !desc) { g_create_error = "null argument"; return JB_ERR_INVALID; }
  *out = nullptr;
  if (desc->dims[0] < 1 || desc->dims[1] < 1 || desc->dims[2] < 1 || desc->num_motif < 1 || desc->num_motif > JB_MAX_MOTIF) {
    g_create_error = "invalid lattice dimensions or motif size"; return JB_ERR_INVALID;
  }
  if (desc->n_ranks < 1 || desc->rank < 0 || desc->rank >= desc->n_ranks || desc->nx_local < 1 || desc->x_begin < 0 ||
      desc->x_begin + desc->nx_local > desc->dims[0]) {
    g_create_error = "invalid slab description"; return JB_ERR_INVALID;
  }
  if (desc->n_ranks == 1 && (desc->x_begin != 0 || desc->nx_local != desc->dims[0])) {
    g_create_error = "single-rank context must own the whole lattice"; return JB_ERR_INVALID;
  }
  const long long N = (long long)desc->nx_local * desc->dims[1] * desc->dims[2] * desc->num_motif;
  if (N >= (1ll << 31)) { g_create_error = "more than 2^31 spins per context"; return JB_ERR_UNSUPPORTED; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (jams_b200 has no CPU fallback)";
    return JB_ERR_CUDA;
  }
  int dev = desc->device;
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
  if (dev >= ndev) { g_create_error = "device ordinal out of range"; return JB_ERR_INVALID; }
  jb_ctx *c = new jb_ctx;
  c->d = *desc; c->device = dev; c->N = (int)N;
  if ((e = cudaSetDevice(dev)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = std::string("CUDA init failed: ") + cudaGetErrorString(e);
    delete c; return JB_ERR_CUDA;
  }
  compute_geometry(c, 0, 0, 0);
  *out = c;
  return JB_OK;
}

void jb_destroy(jb_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->peer_lo_base) cudaIpcCloseMemHandle(c->peer_lo_base);
  if (c->peer_hi_base && !c->same_peer) cudaIpcCloseMemHandle(c->peer_hi_base);
  release_state(c);
  void *p;
  p = c->d_aos; free_dev(p); p = c->d_scratch; free_dev(p);
  p = c->d_nbr_global; free_dev(p); p = c->d_Jtab; free_dev(p); p = c->d_tile_nbr; free_dev(p); p = c->d_tile_J9T; free_dev(p); p = c->d_rows; free_dev(p); p = c->d_bq_global; free_dev(p);
  p = c->d_classes; free_dev(p); p = c->d_site_class; free_dev(p); p = c->d_uni_extra; free_dev(p);
  p = c->d_ell_idx; free_dev(p); p = c->d_ell_val; free_dev(p); p = c->d_pair_J; free_dev(p);
  p = c->d_queue; free_dev(p); p = c->d_trace; free_dev(p); p = c->d_groups; free_dev(p);
  for (int r = 0; r < JB_MAX_REGIONS; ++r) { p = c->d_region[r]; free_dev(p); }
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  for (auto ev : c->ev) cudaEventDestroy(ev);
  for (int k = 0; k <= JB_COPY_CHUNKS; ++k) if (c->copy_ev[k]) cudaEventDestroy(c->copy_ev[k]);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int jb_set_materials(jb_ctx *c, const double *mus, const double *gyro, const double *alpha) {
  if (!c || !mus || !gyro || !alpha) return JB_ERR_INVALID;
  c->h_mus.assign(mus, mus + c->N); c->h_gyro.assign(gyro, gyro + c->N); c->h_alpha.assign(alpha, alpha + c->N);
  c->classes_dirty = true;
  return JB_OK;
}

int jb_set_exchange_template(jb_ctx *c, int32_t n, const int32_t *mi, const int32_t *mj, const int32_t *T3, const double *J9) {
  if (c) c->pairs_reach_x = 0;
  if (!c || n < 0 || (n > 0 && (!mi || !mj || !T3 || !J9))) return JB_ERR_INVALID;
  for (int k = 0; k < n; ++k) {
    if (mi[k] < 0 || mi[k] >= c->d.num_motif || mj[k] < 0 || mj[k] >= c->d.num_motif) JB_FAIL(c, JB_ERR_INVALID, "motif index out of range in exchange template");
  }
  if (c->opt_check_symmetry) {
    // SparseInteractionHamiltonian::finalize (hamiltonian/sparse_interaction.cc:114-118) refuses a 3N x 3N matrix that is not
    // symmetric (SparseMatrix::Builder::is_symmetric, containers/sparse_matrix_builder.h:320-362: element (3i+a, 3j+b) must
    // equal (3j+b, 3i+a) exactly).  In template terms: entry (mi, mj, T, J) needs the entry (mj, mi, -T, J^T).
    std::map<std::array<int, 5>, int> where;
    for (int k = 0; k < n; ++k) where[{mi[k], mj[k], T3[3 * k], T3[3 * k + 1], T3[3 * k + 2]}] = k;
    for (int k = 0; k < n; ++k) {
      auto it = where.find({mj[k], mi[k], -T3[3 * k], -T3[3 * k + 1], -T3[3 * k + 2]});
      bool ok = it != where.end();
      for (int a = 0; a < 3 && ok; ++a) for (int b = 0; b < 3; ++b) if (J9[9 * (size_t)k + 3 * a + b] != J9[9 * (size_t)it->second + 3 * b + a]) ok = false;
      if (!ok) JB_FAIL(c, JB_ERR_INVALID, "sparse matrix for exchange is not symmetric");
    }
  }
  c->t_mi.assign(mi, mi + n); c->t_mj.assign(mj, mj + n); c->t_T.assign(T3, T3 + 3 * n); c->t_J9.assign(J9, J9 + 9 * (size_t)n);
  c->has_template = n > 0;
  c->has_pairs = false;
  c->tables_built = false;
  c->tiling_valid = false;
  return JB_OK;
}

int jb_set_biquadratic_template(jb_ctx *c, int32_t n, const int32_t *mi, const int32_t *mj, const int32_t *T3, const double *B) {
  if (!c || n < 0 || (n > 0 && (!mi || !mj || !T3 || !B))) return JB_ERR_INVALID;
  for (int k = 0; k < n; ++k)
    if (mi[k] < 0 || mi[k] >= c->d.num_motif || mj[k] < 0 || mj[k] >= c->d.num_motif) JB_FAIL(c, JB_ERR_INVALID, "biquadratic template: motif index out of range");
  if (c->opt_check_symmetry) {   // sparse_matrix_builder_.is_symmetric(), cuda_biquadratic_exchange.cu:136-148
    std::map<std::array<int, 5>, double> seen;
    for (int k = 0; k < n; ++k) {
      const std::array<int, 5> key{mi[k], mj[k], T3[3 * k], T3[3 * k + 1], T3[3 * k + 2]};
      if (seen.count(key)) JB_FAIL(c, JB_ERR_INVALID, "Multiple interactions for the same motif pair and translation in the biquadratic template");
      seen[key] = B[k];
    }
    for (const auto &kv : seen) {
      const std::array<int, 5> rev{kv.first[1], kv.first[0], -kv.first[2], -kv.first[3], -kv.first[4]};
      auto it = seen.find(rev);
      if (it == seen.end() || it->second != kv.second) JB_FAIL(c, JB_ERR_INVALID, "sparse matrix for biquadratic-exchange is not symmetric");
    }
  }
  c->bq_mi.assign(mi, mi + n); c->bq_mj.assign(mj, mj + n); c->bq_T.assign(T3, T3 + 3 * (size_t)n); c->bq_B.assign(B, B + n);
  c->has_bq = n > 0;
  c->bq_built = false;
  c->tiling_valid = false;
  return JB_OK;
}

int jb_detect_exchange_template(const jb_lattice_desc *d, int64_t n_pairs, const int32_t *pi, const int32_t *pj,
                                const int32_t *vid, int32_t n_values, const double *J9, int32_t capacity,
                                int32_t *n_template, int32_t *motif_i, int32_t *motif_j, int32_t *T3, double *J9_out) {
  if (!d || !n_template || n_pairs < 0 || capacity < 0 || (n_pairs > 0 && (!pi || !pj || !vid || !J9))) return JB_ERR_INVALID;
  if (capacity > 0 && (!motif_i || !motif_j || !T3 || !J9_out)) return JB_ERR_INVALID;
  *n_template = -1;
  const int L[3] = {d->dims[0], d->dims[1], d->dims[2]};
  const int M = d->num_motif;
  if (L[0] < 1 || L[1] < 1 || L[2] < 1 || M < 1 || M > 255) return JB_ERR_INVALID;
  const long long Ntot = (long long)L[0] * L[1] * L[2] * M;
  struct Entry { int mi, mj, T[3], vid; long long count; int order; };
  std::unordered_map<uint64_t, Entry> seen;
  auto decode = [&](int s, int cell[3], int &m) {
    m = s % M; int r = s / M; cell[2] = r % L[2]; r /= L[2]; cell[1] = r % L[1]; cell[0] = r / L[1];
  };
  for (int64_t p = 0; p < n_pairs; ++p) {
    if (pi[p] < 0 || pi[p] >= Ntot || pj[p] < 0 || pj[p] >= Ntot || vid[p] < 0 || vid[p] >= n_values) return JB_ERR_INVALID;
    int ci[3], cj[3], mi, mj;
    decode(pi[p], ci, mi);
    if (ci[0] < d->x_begin || ci[0] >= d->x_begin + d->nx_local) continue;
    decode(pj[p], cj, mj);
    int T[3];
    for (int k = 0; k < 3; ++k) {
      int t = cj[k] - ci[k];
      if (d->periodic[k]) {  // minimum image, t in (-L/2, L/2]
        if (2 * t > L[k]) t -= L[k];
        else if (2 * t <= -L[k]) t += L[k];
      }
      if (t < -127 || t > 127) return JB_OK;  // not a short-range template
      T[k] = t;
    }
    const uint64_t key = ((uint64_t)mi << 32) | ((uint64_t)mj << 24) | ((uint64_t)(T[0] + 128) << 16) | ((uint64_t)(T[1] + 128) << 8) | (uint64_t)(T[2] + 128);
    auto it = seen.find(key);
    if (it == seen.end()) {
      if ((int)seen.size() >= capacity) return JB_OK;
      Entry e{mi, mj, {T[0], T[1], T[2]}, vid[p], 1, (int)seen.size()};
      seen.emplace(key, e);
    } else {
      if (it->second.vid != vid[p] && memcmp(J9 + 9 * (size_t)it->second.vid, J9 + 9 * (size_t)vid[p], 9 * sizeof(double)) != 0) return JB_OK;
      it->second.count++;
    }
  }
  std::vector<Entry> entries(seen.size());
  for (auto &kv : seen) entries[kv.second.order] = kv.second;
  for (const Entry &e : entries) {
    long long expect = 1;
    for (int k = 0; k < 3; ++k) {
      const int lo = k == 0 ? d->x_begin : 0, n = k == 0 ? d->nx_local : L[k];
      if (d->periodic[k]) {
        if (e.T[k] != 0 && L[k] < 2 * std::abs(e.T[k]) + 1) return JB_OK;  // images alias: the reference would have thrown
        expect *= n;
      } else {
        long long cnt = 0;
        for (int x = lo; x < lo + n; ++x) cnt += (x + e.T[k] >= 0 && x + e.T[k] < L[k]) ? 1 : 0;
        expect *= cnt;
      }
    }
    if (e.count != expect) return JB_OK;  // some cell lacks (or repeats) this neighbour: impurity / vacancy
  }
  for (size_t n = 0; n < entries.size(); ++n) {
    motif_i[n] = entries[n].mi; motif_j[n] = entries[n].mj;
    for (int k = 0; k < 3; ++k) T3[3 * n + k] = entries[n].T[k];
    memcpy(J9_out + 9 * n, J9 + 9 * (size_t)entries[n].vid, 9 * sizeof(double));
  }
  *n_template = (int32_t)entries.size();
  return JB_OK;
}

int jb_set_exchange_pairs(jb_ctx *c, int64_t n_pairs, const int32_t *pi, const int32_t *pj, const int32_t *vid, int32_t n_values, const double *J9) {
  if (!c || n_pairs < 0 || n_values < 0 || (n_pairs > 0 && (!pi || !pj || !vid || !J9))) return JB_ERR_INVALID;
  if (c->opt_detect_template && n_pairs > 0) {
    const int cap = JB_TILE_MAX_NBR;
    std::vector<int32_t> mi(cap), mj(cap), T3(3 * cap);
    std::vector<double> J9t(9 * (size_t)cap);
    int32_t nt = -1;
    int rc = jb_detect_exchange_template(&c->d, n_pairs, pi, pj, vid, n_values, J9, cap, &nt, mi.data(), mj.data(), T3.data(), J9t.data());
    if (rc == JB_ERR_INVALID) JB_FAIL(c, JB_ERR_INVALID, "pair index out of range");
    if (nt > 0) return jb_set_exchange_template(c, nt, mi.data(), mj.data(), T3.data(), J9t.data());
  }
  const long long n_global = (long long)c->d.dims[0] * c->d.dims[1] * c->d.dims[2] * c->d.num_motif;
  if (c->opt_check_symmetry) {
    // the same symmetry requirement on the explicit list: every (i, j, J) needs (j, i, J^T).  Sort-based, no hash of the list
    std::vector<int64_t> order(n_pairs);
    for (int64_t p = 0; p < n_pairs; ++p) order[p] = p;
    auto key = [&](int64_t p) { return ((uint64_t)(uint32_t)pi[p] << 32) | (uint32_t)pj[p]; };
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key(a) < key(b); });
    for (int64_t p = 0; p < n_pairs; ++p) {
      if (vid[p] < 0 || vid[p] >= n_values) JB_FAIL(c, JB_ERR_INVALID, "pair index out of range");
      const uint64_t want = ((uint64_t)(uint32_t)pj[p] << 32) | (uint32_t)pi[p];
      auto it = std::lower_bound(order.begin(), order.end(), want, [&](int64_t a, uint64_t k) { return key(a) < k; });
      bool ok = it != order.end() && key(*it) == want;
      if (ok) {
        if (vid[*it] < 0 || vid[*it] >= n_values) JB_FAIL(c, JB_ERR_INVALID, "pair index out of range");
        const double *A = J9 + 9 * (size_t)vid[p], *B = J9 + 9 * (size_t)vid[*it];
        for (int a = 0; a < 3 && ok; ++a) for (int b = 0; b < 3; ++b) if (A[3 * a + b] != B[3 * b + a]) ok = false;
      }
      if (!ok) JB_FAIL(c, JB_ERR_INVALID, "sparse matrix for exchange is not symmetric");
    }
  }
  JB_CUDA(c, cudaSetDevice(c->device));
  // Geometry: on one rank the list needs no ghost cells (every neighbour is addressed directly, periodic wrap included).  On a
  // slab-decomposed lattice a neighbour across a slab face is addressed through the x ghost planes, which the neighbouring rank's
  // stage kernel fills (P2P stores) like for a template; the ghost depth is the largest x distance (minimum image) of ANY pair of
  // the list, so that every rank lays its box out alike (they store into each other's boxes with their own geometry).
  const int Nx = c->d.dims[0], Ny = c->d.dims[1], Nz = c->d.dims[2], M = c->d.num_motif;
  auto decode = [&](int ref, int &x, int &y, int &z, int &m) { m = ref % M; int r = ref / M; z = r % Nz; r /= Nz; y = r % Ny; x = r / Ny; };
  auto x_distance = [&](int xi, int xj) {   // signed, minimum image along a periodic x
    int d = xj - xi;
    if (c->d.periodic[0]) { if (2 * d > Nx) d -= Nx; else if (2 * d < -Nx) d += Nx; }
    return d;
  };
  int reach_x = 0;
  for (int64_t p = 0; p < n_pairs; ++p) {
    if (pi[p] < 0 || pi[p] >= n_global || pj[p] < 0 || pj[p] >= n_global || vid[p] < 0 || vid[p] >= n_values) JB_FAIL(c, JB_ERR_INVALID, "pair index out of range");
    if (c->d.n_ranks > 1) {
      int xi, yi, zi, mi, xj, yj, zj, mj;
      decode(pi[p], xi, yi, zi, mi); decode(pj[p], xj, yj, zj, mj);
      reach_x = std::max(reach_x, std::abs(x_distance(xi, xj)));
    }
  }
  if (c->d.n_ranks > 1 && reach_x > c->d.nx_local) JB_FAIL(c, JB_ERR_INVALID, "slab thinner than the x range of the neighbour list");
  c->has_template = false; c->t_mi.clear(); c->t_mj.clear(); c->t_T.clear(); c->t_J9.clear();
  c->has_pairs = false;
  c->pairs_reach_x = reach_x;
  int rc = ensure_ready(c);  // geometry: no y / z ghosts, x ghosts only on several ranks
  if (rc) return rc;
  const JbGeom &g = c->g;
  const int N = c->N;
  const int x0 = c->d.x_begin, nx = c->d.nx_local;
  std::vector<int> count(N, 0);
  auto local_ref = [&](int ref) -> long long {   // global reference id -> local reference id, or -1 outside this slab
    int x, y, z, m; decode(ref, x, y, z, m);
    if (x < x0 || x >= x0 + nx) return -1;
    return ((((long long)(x - x0) * Ny + y) * Nz + z) * M + m);
  };
  for (int64_t p = 0; p < n_pairs; ++p) { const long long li = local_ref(pi[p]); if (li >= 0) count[li]++; }
  int width = 0;
  for (int i = 0; i < N; ++i) width = std::max(width, count[i]);
  std::vector<int> idx((size_t)width * N, -1), val((size_t)width * N, 0), fill(N, 0);
  for (int64_t p = 0; p < n_pairs; ++p) {  // pairs arrive sorted by {i,j}: ascending-j order is kept per row
    const long long li = local_ref(pi[p]);
    if (li < 0) continue;
    int xi, yi, zi, mi, xj, yj, zj, mj;
    decode(pi[p], xi, yi, zi, mi); decode(pj[p], xj, yj, zj, mj);
    const long long q = (((long long)(xi - x0) * g.Ny + yi) * g.M + mi) * g.Nz + zi;   // interior layout order of the row
    // neighbour: inside the slab by its own cell, across a face by the ghost plane at its x distance from the row's cell
    int xl = xj - x0;
    if (c->d.n_ranks > 1 && (xj < x0 || xj >= x0 + nx)) xl = (xi - x0) + x_distance(xi, xj);
    if (xl < -g.gx || xl >= nx + g.gx) JB_FAIL(c, JB_ERR_INVALID, "neighbour list reaches beyond the x ghost planes");
    const long long gj = ((long long)(xl + g.gx) * g.PY + (yj + g.gy)) * g.sY + (long long)mj * g.PZ + (zj + g.oz);
    const int e = fill[li]++;
    idx[(size_t)e * N + q] = (int)gj;
    val[(size_t)e * N + q] = vid[p];
  }
  bool iso = true;
  for (int v = 0; v < n_values; ++v) {
    const double *J = J9 + 9 * v;
    if (!(J[1] == 0 && J[2] == 0 && J[3] == 0 && J[5] == 0 && J[6] == 0 && J[7] == 0 && J[0] == J[4] && J[4] == J[8])) iso = false;
  }
  void *p;
  p = c->d_ell_idx; free_dev(p); p = c->d_ell_val; free_dev(p); p = c->d_pair_J; free_dev(p);
  c->d_ell_idx = c->d_ell_val = nullptr; c->d_pair_J = nullptr;
  if (width > 0) {
    JB_CUDA(c, cudaMalloc(&c->d_ell_idx, idx.size() * sizeof(int)));
    JB_CUDA(c, cudaMalloc(&c->d_ell_val, val.size() * sizeof(int)));
    JB_CUDA(c, cudaMalloc(&c->d_pair_J, (size_t)std::max(1, n_values) * 9 * sizeof(double)));
    JB_CUDA(c, cudaMemcpy(c->d_ell_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    JB_CUDA(c, cudaMemcpy(c->d_ell_val, val.data(), val.size() * sizeof(int), cudaMemcpyHostToDevice));
    JB_CUDA(c, cudaMemcpy(c->d_pair_J, J9, (size_t)n_values * 9 * sizeof(double), cudaMemcpyHostToDevice));
  }
  c->ell_width = width; c->n_pair_values = n_values; c->pairs_iso = iso; c->has_pairs = width > 0;
  return JB_OK;
}

int jb_set_uniaxial(jb_ctx *c, int32_t power, const double *magnitude, const double *axis) {
  if (!c) return JB_ERR_INVALID;
  if (power == 0 || !magnitude) { c->uni_power = 0; c->h_K.clear(); c->h_axis.clear(); c->classes_dirty = true; return JB_OK; }
  if (!(power == 2 || power == 4 || power == 6) || !axis) JB_FAIL(c, JB_ERR_INVALID, "Unsupported anisotropy power (K1,K2,K3 = 2,4,6)");
  c->uni_power = power; c->h_K.assign(magnitude, magnitude + c->N); c->h_axis.assign(axis, axis + 3 * (size_t)c->N);
  c->classes_dirty = true;
  return JB_OK;
}

int jb_set_uniaxial_term(jb_ctx *c, int32_t slot, int32_t power, const double *magnitude, const double *axis) {
  if (!c) return JB_ERR_INVALID;
  if (slot == 0) return jb_set_uniaxial(c, power, magnitude, axis);
  if (slot < 0 || slot >= JB_MAX_UNIAXIAL) JB_FAIL(c, JB_ERR_UNSUPPORTED, "more than three uniaxial Hamiltonians");
  const int q = slot - 1;
  if (power == 0 || !magnitude) { c->uni_powerx[q] = 0; c->h_Kx[q].clear(); c->h_axisx[q].clear(); }
  else {
    if (!(power == 2 || power == 4 || power == 6) || !axis) JB_FAIL(c, JB_ERR_INVALID, "Unsupported anisotropy power (K1,K2,K3 = 2,4,6)");
    c->uni_powerx[q] = power; c->h_Kx[q].assign(magnitude, magnitude + c->N); c->h_axisx[q].assign(axis, axis + 3 * (size_t)c->N);
  }
  c->classes_dirty = true;
  c->tiling_valid = false;   // a context with extra uniaxial terms steps on the direct kernels
  return JB_OK;
}

int jb_set_zeeman(jb_ctx *c, const double *dc, const double *ac, const double *omega) {
  if (!c) return JB_ERR_INVALID;
  if ((ac == nullptr) != (omega == nullptr)) JB_FAIL(c, JB_ERR_INVALID, "must have a field and a frequency");
  c->has_zeeman = dc != nullptr;
  if (dc) c->h_dc.assign(dc, dc + 3 * (size_t)c->N); else c->h_dc.clear();
  c->has_ac = ac != nullptr;
  if (ac) {
    c->h_ac.assign(ac, ac + 3 * (size_t)c->N); c->h_omega.assign(omega, omega + c->N);
    if (!dc) { c->h_dc.assign(3 * (size_t)c->N, 0.0); c->has_zeeman = true; }
  } else { c->h_ac.clear(); c->h_omega.clear(); }
  c->classes_dirty = true;
  return JB_OK;
}

int jb_set_applied_field(jb_ctx *c, const double B[3], int32_t enable) {
  if (!c || (enable && !B)) return JB_ERR_INVALID;
  c->has_applied = enable != 0;
  c->applied_type = JB_FIELD_STATIC;
  for (int d = 0; d < 3; ++d) c->applied_B[d] = enable ? B[d] : 0.0;
  return JB_OK;
}

int jb_set_applied_field_pulse(jb_ctx *c, const double B[3], int32_t type, double t0, double fbw, double fc) {
  if (!c || !B) return JB_ERR_INVALID;
  if (type != JB_FIELD_STATIC && type != JB_FIELD_SINC && type != JB_FIELD_SINC_COS) JB_FAIL(c, JB_ERR_INVALID, "unknown field pulse type");
  c->has_applied = true;
  c->applied_type = type; c->applied_t0 = t0; c->applied_fbw = fbw; c->applied_fc = fc;
  for (int d = 0; d < 3; ++d) c->applied_B[d] = B[d];
  return JB_OK;
}

int jb_import_spins(jb_ctx *c, const double *s_aos, int32_t on_device) {
  if (!c || !s_aos) return JB_ERR_INVALID;
  int rc = ensure_ready(c); if (rc) return rc;
  double *dst[3] = {c->S0[0], c->S0[1], c->S0[2]};
  if (!on_device) {
    // host memory: x-chunks travel on the copy stream (cudaMemcpyAsync: the copy engine, full PCIe rate from pinned memory)
    // while the layout kernel of the previous chunk runs on the context's stream; the x ghost planes (periodic images of the
    // far end of the slab) come last.  globals::s of the reference is exactly such an AoS array (core/lattice.cc:688).
    rc = ensure_aos(c); if (rc) return rc;
    rc = ensure_copy_stream(c); if (rc) return rc;
    const JbGeom &g = c->g;
    const size_t per_plane = (size_t)g.Ny * g.Nz * g.M * 3;
    const int nchunk = std::min(JB_COPY_CHUNKS, g.nx);
    JB_CUDA(c, cudaEventRecord(c->copy_ev[JB_COPY_CHUNKS], c->stream));          // earlier kernels may still read the staging buffer
    JB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->copy_ev[JB_COPY_CHUNKS], 0));
    for (int k = 0; k < nchunk; ++k) {
      const int x0 = (int)((long long)k * g.nx / nchunk), x1 = (int)((long long)(k + 1) * g.nx / nchunk);
      JB_CUDA(c, cudaMemcpyAsync(c->d_aos + per_plane * x0, s_aos + per_plane * x0, per_plane * (x1 - x0) * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
      JB_CUDA(c, cudaEventRecord(c->copy_ev[k], c->copy_stream));
      JB_CUDA(c, cudaStreamWaitEvent(c->stream, c->copy_ev[k], 0));
      JB_CUDA(c, jbk_import_planes(g, c->d_aos, dst, x0, x1, c->stream)); c->launches++;
    }
    JB_CUDA(c, jbk_fill_ghosts(g, dst, c->d.n_ranks == 1, c->stream)); c->launches++;
  } else {
    JB_CUDA(c, jbk_import(c->g, s_aos, dst, c->d.n_ranks == 1, c->stream)); c->launches += 2;
  }
  if (c->d.n_ranks > 1 && c->g.gx > 0) {
    if (!c->halo_connected) JB_FAIL(c, JB_ERR_INVALID, "multi-rank context: call jb_halo_connect before jb_import_spins");
    // neighbours may still be reading my previous ghosts: exchange happens inside an epoch handshake
    const double *s0[3] = {c->S0[0], c->S0[1], c->S0[2]};
    double *lo[3] = {c->peer_lo_S0[0], c->peer_lo_S0[1], c->peer_lo_S0[2]};
    double *hi[3] = {c->peer_hi_S0[0], c->peer_hi_S0[1], c->peer_hi_S0[2]};
    JB_CUDA(c, jbk_push_x_ghosts(c->g, s0, lo, hi, c->stream)); c->launches++;
    c->epoch++;
    JB_CUDA(c, jbk_signal(c->peer_lo_flags ? c->peer_lo_flags + 1 : nullptr, c->peer_hi_flags ? c->peer_hi_flags + 0 : nullptr, c->epoch, c->stream)); c->launches++;
  }
  if (!on_device) JB_CUDA(c, cudaStreamSynchronize(c->stream));
  return JB_OK;
}

int jb_export_spins(jb_ctx *c, double *s_aos, int32_t on_device) {
  if (!c || !s_aos) return JB_ERR_INVALID;
  if (!c->state_allocated) JB_FAIL(c, JB_ERR_INVALID, "no spins have been imported");
  JB_CUDA(c, cudaSetDevice(c->device));
  const double *src[3] = {c->S0[0], c->S0[1], c->S0[2]};
  if (on_device) {
    JB_CUDA(c, jbk_export(c->g, src, s_aos, c->stream)); c->launches++;
    return JB_OK;
  }
  int rc = ensure_aos(c); if (rc) return rc;
  rc = ensure_copy_stream(c); if (rc) return rc;
  // x-chunks: the layout kernel of chunk k + 1 runs while chunk k crosses PCIe on the copy stream
  const JbGeom &g = c->g;
  const size_t per_plane = (size_t)g.Ny * g.Nz * g.M * 3;
  const int nchunk = std::min(JB_COPY_CHUNKS, g.nx);
  for (int k = 0; k < nchunk; ++k) {
    const int x0 = (int)((long long)k * g.nx / nchunk), x1 = (int)((long long)(k + 1) * g.nx / nchunk);
    JB_CUDA(c, jbk_export_planes(g, src, c->d_aos, x0, x1, c->stream)); c->launches++;
    JB_CUDA(c, cudaEventRecord(c->copy_ev[k], c->stream));
    JB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->copy_ev[k], 0));
    JB_CUDA(c, cudaMemcpyAsync(s_aos + per_plane * x0, c->d_aos + per_plane * x0, per_plane * (x1 - x0) * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
  }
  JB_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  JB_CUDA(c, cudaStreamSynchronize(c->stream));
  return JB_OK;
}

// one launch of the persistent TMA kernels (pair / rows / the RK4 stages on the pair kernel): output boxes, class constants of
// this stage, launch shape and work-item plan, queue counters, the in-kernel halo handshake
static int launch_tile_stage(jb_ctx *c, JbTileParams &tp, const JbStageParams &p, int stage, int th, bool recu, bool rk4, int class_table_index, bool fold) {
  int rc;
  for (int k = 0; k < 3; ++k) { tp.out[k] = p.out[k]; tp.out_lo[k] = p.out_lo[k]; tp.out_hi[k] = p.out_hi[k]; tp.u[k] = p.u[k]; }
  tp.step = p.step;
  tp.dt = p.dt;
  const JbClass *cls = c->h_class_tab.data() + (size_t)class_table_index * c->h_classes.size();
  for (int m = 0; m < c->g.M; ++m) tp.cls[m] = cls[c->class_of_motif[m]];
  const int lay = rk4 ? (stage > 0 ? 1 : 0) : stage;   // shared-memory layout: with or without the second ring
  tp.R = c->tiling.Rs[lay];
  rc = tile_launch_shape(c, tp, stage, th, recu ? 1 : 0, rk4); if (rc) return rc;
  const jb_ctx::Tiling::Shape &sh = rk4 ? c->tiling.shape_rk4[stage][th] : c->tiling.shape[stage][th][recu ? 1 : 0];
  tp.n_chunks = sh.n_chunks;
  tp.n_items = sh.n_chunks * tp.n_cols;
  for (int q = 0; q < sh.n_chunks; ++q) { tp.chunk_x0[q] = sh.x0[q]; tp.chunk_xc[q] = sh.xc[q]; }
  tp.queue = c->d_queue + (c->stage_launches & 1ull);          // this launch's item counter (zero: see queue_next)
  tp.queue_next = c->d_queue + ((c->stage_launches + 1) & 1ull);   // ... and the one it zeroes for the next launch
  c->stage_launches++;
  tp.trace = c->opt_trace ? c->d_trace : nullptr;
  tp.halo = JbHalo{};
  if (fold) {
    JbHalo &h = tp.halo;
    h.enabled = (c->peer_lo_flags ? 1 : 0) | (c->peer_hi_flags ? 2 : 0);
    h.flags = c->flags;
    h.sig_lo = c->peer_lo_flags ? c->peer_lo_flags + 1 : nullptr;
    h.sig_hi = c->peer_hi_flags ? c->peer_hi_flags + 0 : nullptr;
    h.wait_epoch = c->epoch; h.signal_epoch = c->epoch + 1;
    h.face_count = c->d_queue + 2;
    h.face_target[0] = (unsigned int)sh.face_items[0];   // one CTA-level arrival per face item (halo_face_done)
    h.face_target[1] = (unsigned int)sh.face_items[1];
  }
  if (rk4) {
    // S ring: the stage input (S0, S1, V, S1); second ring: the tile's own s_old = S0 (tensor maps [3])
    const int in_map = stage == 0 ? 0 : (stage == 2 ? 5 : 1);
    const CUtensorMap tm[6] = {c->tmap[in_map][0], c->tmap[in_map][1], c->tmap[in_map][2], c->tmap[3][0], c->tmap[3][1], c->tmap[3][2]};
    JB_CUDA(c, jbk_rk4_stage_pair(tp, tm, stage, th, c->tiling.threads, sh.grid, c->tiling.smem[lay], c->stream));
  } else {
    const int ua = recu ? 3 : 2;   // recover_u: the corrector's second ring carries the tile's own s_n (S0) instead of u
    const CUtensorMap tm[6] = {c->tmap[stage][0], c->tmap[stage][1], c->tmap[stage][2], c->tmap[ua][0], c->tmap[ua][1], c->tmap[ua][2]};
    if (c->tiling.rows) JB_CUDA(c, jbk_stage_rows(tp, tm, stage, th, c->tiling.threads, sh.grid, c->tiling.smem[stage], c->stream));
    else JB_CUDA(c, jbk_stage_pair(tp, tm, stage, th, c->iso ? 1 : 0, recu ? 1 : 0, c->tiling.threads, sh.grid, c->tiling.smem[stage], c->stream));
  }
  c->trace_ctas = sh.grid;
  return JB_OK;
}

int jb_step(jb_ctx *c, int32_t nsteps, double dt, double time_ps, double T, uint64_t seed, uint64_t first_step, int32_t gilbert) {
  if (!c || nsteps < 0 || !(dt > 0.0) || T < 0.0) return JB_ERR_INVALID;
  if (!c->state_allocated) JB_FAIL(c, JB_ERR_INVALID, "no spins have been imported");
  int rc = ensure_ready(c); if (rc) return rc;
  if (c->d.n_ranks > 1 && c->g.gx > 0 && !c->halo_connected) JB_FAIL(c, JB_ERR_INVALID, "multi-rank context: jb_halo_connect has not been called");
  const bool multi = c->d.n_ranks > 1 && c->g.gx > 0;

  choose_tiling(c);
  const bool use_tile = c->tiling.ok && !c->has_pairs;
  JbTileParams tp{};
  if (use_tile) {
    rc = build_tmaps(c); if (rc) return rc;
    rc = ensure_queue(c); if (rc) return rc;
    fill_tile_params(c, tp);
    for (int r = 0; r < 10; ++r) {   // Philox4x32-10 key schedule (Salmon et al.)
      tp.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
      tp.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
  }
  c->last_stage_kernel = c->has_pairs ? JB_KERNEL_ELL : (use_tile ? (c->tiling.rows ? JB_KERNEL_ROWS : JB_KERNEL_PAIR) : JB_KERNEL_DIRECT);
  const int thermal = T > 0.0 ? 1 : 0;
  // data flow of the TMA kernel (option recover_u): 1 = the corrector rebuilds the Heun intermediate from s_n and s* (120 B per
  // update; at T > 0 it then draws the site's noise a second time), 0 = the predictor stores it with the noise part of the
  // corrector folded in (144 B), 2 = 1 at T = 0 and 0 at T > 0
  const bool recu = use_tile && (c->tiling.rows || c->opt_recover_u == 1 || (c->opt_recover_u == 2 && !thermal));   // the rows kernel has no other data flow
  // the epoch handshake of a slab-decomposed run happens inside the TMA kernel; the other kernels bracket each launch with
  // wait / signal launches
  // (a neighbour slab that lives on THIS device shares its SMs with me: a resident kernel that polls for its flags could keep
  // it from ever running, so that set-up -- only tests use it -- keeps the separate launches unless fold_halo = 2 insists)
  const bool fold = multi && use_tile && (c->opt_fold_halo == 2 || (c->opt_fold_halo == 1 && !c->peer_on_my_device));

  const int max_chunk = time_dependent(c) ? 2048 : nsteps;
  for (int done = 0; done < nsteps;) {
    const int chunk = std::min(nsteps - done, std::max(1, max_chunk));
    std::vector<double> times;
    if (time_dependent(c)) {
      for (int n = 0; n < chunk; ++n) { const double t0 = time_ps + (done + n) * dt; times.push_back(t0); times.push_back(t0 + dt); }  // cpu_llg_heun.cc:46,103-104
    } else {
      times.push_back(time_ps);
    }
    rc = upload_classes(c, times, dt, T, gilbert, JB_TERM_TOTAL); if (rc) return rc;

    for (int n = 0; n < chunk; ++n) {
      // option time_kernels = N: the launches of every N-th step are bracketed by events (N = 1: every step; a record costs ~2 us)
      c->time_this_step = c->opt_time_kernels > 0 && (done + n) % c->opt_time_kernels == 0;
      if (c->time_this_step) c->timed_steps++;
      for (int stage = 0; stage < 2; ++stage) {
        JbStageParams p{};
        p.g = c->g;
        fill_tables(c, p.t, time_dependent(c) ? 2 * n + stage : 0);
        for (int k = 0; k < 3; ++k) {
          p.in[k] = stage == 0 ? c->S0[k] : c->S1[k];
          p.out[k] = stage == 0 ? c->S1[k] : c->S0[k];
          p.u[k] = c->U[k];
          if (multi) {
            p.out_lo[k] = stage == 0 ? c->peer_lo_S1[k] : c->peer_lo_S0[k];
            p.out_hi[k] = stage == 0 ? c->peer_hi_S1[k] : c->peer_hi_S0[k];
          } else if (c->g.per[0] && c->g.gx > 0) {
            p.out_lo[k] = p.out[k]; p.out_hi[k] = p.out[k];
          } else {
            p.out_lo[k] = nullptr; p.out_hi[k] = nullptr;
          }
        }
        p.seed = seed; p.step = first_step + (uint64_t)(done + n);
        // stored-u data flow: the corrector needs no noise, the predictor folds the noise part of its right-hand side into u
        // (jb_device.cuh); recover_u: the corrector draws the noise itself
        const int th = (stage == 0 || recu) ? thermal : 0;
        p.thermal = th;
        if (multi && !fold) {
          // ghosts I read were written by the neighbours' previous stage; the boxes I write into were
          // last read by the neighbours' previous stage: both are covered by their last signal
          JB_CUDA(c, jbk_wait(c->flags, c->peer_lo_flags != nullptr, c->peer_hi_flags != nullptr, c->epoch, c->stream)); c->launches++;
        }
        record_event(c, 2 * stage);
        if (c->has_pairs) {
          JB_CUDA(c, jbk_stage_pairs(p, c->d_ell_idx, c->d_ell_val, c->ell_width, c->d_pair_J, c->pairs_iso ? 1 : 0, stage, c->stream));
        } else if (use_tile) {
          rc = launch_tile_stage(c, tp, p, stage, th, recu, false, time_dependent(c) ? 2 * n + stage : 0, fold); if (rc) return rc;
        } else {
          JB_CUDA(c, jbk_stage_direct(p, stage, c->stream));
        }
        c->launches++;
        record_event(c, 2 * stage + 1);
        if (multi) {
          c->epoch++;
          if (!fold) { JB_CUDA(c, jbk_signal(c->peer_lo_flags ? c->peer_lo_flags + 1 : nullptr, c->peer_hi_flags ? c->peer_hi_flags + 0 : nullptr, c->epoch, c->stream)); c->launches++; }
        }
      }
    }
    done += chunk;
  }
  return JB_OK;
}


int jb_step_rk4(jb_ctx *c, int32_t nsteps, double dt, double time_ps, double T, uint64_t seed, uint64_t first_step, int32_t gilbert) {
  if (!c || nsteps < 0 || !(dt > 0.0) || T < 0.0) return JB_ERR_INVALID;
  if (!c->state_allocated) JB_FAIL(c, JB_ERR_INVALID, "no spins have been imported");
  int rc = ensure_ready(c); if (rc) return rc;
  const bool multi = c->d.n_ranks > 1 && c->g.gx > 0;
  if (multi && !c->halo_connected) JB_FAIL(c, JB_ERR_INVALID, "multi-rank context: jb_halo_connect has not been called");
  if (c->has_pairs) JB_FAIL(c, JB_ERR_UNSUPPORTED, "jb_step_rk4 needs a translation-invariant exchange template (jb_set_exchange_template, or jb_set_exchange_pairs with template detection)");
  const bool periodic_x = c->g.per[0] && c->g.gx > 0;
  // the four stages on the persistent TMA kernel (isotropic templates the pair kernel can tile; option kernel = 0: direct gathers)
  choose_tiling(c);
  const bool use_tile = c->tiling.ok && !c->tiling.rows && c->iso && !c->has_bq && !c->has_uni_extra();
  JbTileParams tp{};
  if (use_tile) {
    rc = build_tmaps(c); if (rc) return rc;
    rc = ensure_queue(c); if (rc) return rc;
    fill_tile_params(c, tp);
    for (int r = 0; r < 10; ++r) {   // Philox4x32-10 key schedule
      tp.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
      tp.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
  }
  const bool fold = multi && use_tile && (c->opt_fold_halo == 2 || (c->opt_fold_halo == 1 && !c->peer_on_my_device));
  c->last_stage_kernel = use_tile ? JB_KERNEL_PAIR : JB_KERNEL_DIRECT;
  for (int done = 0; done < nsteps;) {
    const int chunk = time_dependent(c) ? std::min(nsteps - done, 1024) : nsteps - done;
    std::vector<double> times;
    if (time_dependent(c)) {   // fields at t0, t0 + dt/2 (k2 and k3), t0 + dt (cuda_rk4_base.cu:70-71,78-79,86-87)
      for (int n = 0; n < chunk; ++n) { const double t0 = time_ps + (done + n) * dt; times.push_back(t0); times.push_back(t0 + 0.5 * dt); times.push_back(t0 + dt); }
    } else {
      times.push_back(time_ps);
    }
    rc = upload_classes(c, times, dt, T, gilbert, JB_TERM_TOTAL); if (rc) return rc;
    for (int n = 0; n < chunk; ++n) {
      c->time_this_step = c->opt_time_kernels > 0 && (done + n) % c->opt_time_kernels == 0;
      if (c->time_this_step) c->timed_steps++;
      for (int stage = 0; stage < 4; ++stage) {
        JbStageParams p{};
        p.g = c->g;
        const int tsel = stage == 0 ? 0 : (stage == 3 ? 2 : 1);
        fill_tables(c, p.t, time_dependent(c) ? 3 * n + tsel : 0);
        double *const *in = stage == 0 ? c->S0 : (stage == 2 ? c->V : c->S1);          // S0 -> S1 -> V -> S1 -> S0
        double *const *out = stage == 0 ? c->S1 : (stage == 1 ? c->V : (stage == 2 ? c->S1 : c->S0));
        double *const *plo = stage == 0 ? c->peer_lo_S1 : (stage == 1 ? c->peer_lo_V : (stage == 2 ? c->peer_lo_S1 : c->peer_lo_S0));
        double *const *phi = stage == 0 ? c->peer_hi_S1 : (stage == 1 ? c->peer_hi_V : (stage == 2 ? c->peer_hi_S1 : c->peer_hi_S0));
        for (int k = 0; k < 3; ++k) {
          p.in[k] = in[k]; p.out[k] = out[k]; p.u[k] = c->U[k]; p.s_old[k] = c->S0[k];
          if (multi) { p.out_lo[k] = plo[k]; p.out_hi[k] = phi[k]; }
          else { p.out_lo[k] = periodic_x ? out[k] : nullptr; p.out_hi[k] = periodic_x ? out[k] : nullptr; }
        }
        p.seed = seed; p.step = first_step + (uint64_t)(done + n);
        p.thermal = T > 0.0 ? 1 : 0;
        p.dt = dt;
        if (multi && !fold) {   // as in jb_step: the neighbours' previous stage wrote the ghosts this stage reads and read the boxes it writes
          JB_CUDA(c, jbk_wait(c->flags, c->peer_lo_flags != nullptr, c->peer_hi_flags != nullptr, c->epoch, c->stream)); c->launches++;
        }
        record_event(c, stage < 2 ? 0 : 2);   // jb_last_step_kernel_ms: out2[0] = stages 1 + 2, out2[1] = stages 3 + 4
        if (use_tile) { rc = launch_tile_stage(c, tp, p, stage, p.thermal, false, true, time_dependent(c) ? 3 * n + tsel : 0, fold); if (rc) return rc; }
        else JB_CUDA(c, jbk_rk4_stage_direct(p, stage, c->stream));
        c->launches++;
        record_event(c, stage < 2 ? 1 : 3);
        if (multi) {
          c->epoch++;
          if (!fold) { JB_CUDA(c, jbk_signal(c->peer_lo_flags ? c->peer_lo_flags + 1 : nullptr, c->peer_hi_flags ? c->peer_hi_flags + 0 : nullptr, c->epoch, c->stream)); c->launches++; }
        }
      }
    }
    done += chunk;
  }
  return JB_OK;
}

int jb_noise(jb_ctx *c, double dt, double T, uint64_t seed, uint64_t step, int32_t gilbert, int32_t normals_only, double *xi, int32_t on_device) {
  if (!c || !xi) return JB_ERR_INVALID;
  int rc = ensure_ready(c); if (rc) return rc;
  std::vector<double> times{0.0};
  rc = upload_classes(c, times, dt, T, gilbert, -1); if (rc) return rc;
  JbTables t; fill_tables(c, t, 0);
  double *dst = xi;
  if (!on_device) { rc = ensure_aos(c); if (rc) return rc; dst = c->d_aos; }
  JB_CUDA(c, jbk_noise(c->g, t, seed, step, normals_only, dst, c->stream)); c->launches++;
  if (!on_device) {
    JB_CUDA(c, cudaMemcpyAsync(xi, dst, (size_t)c->N * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    JB_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return JB_OK;
}

int jb_fields(jb_ctx *c, int32_t term, double time_ps, double *h_aos, int32_t on_device) {
  if (!c || !h_aos || term < 0 || term > JB_TERM_UNIAXIAL_3) return JB_ERR_INVALID;
  if (!c->state_allocated) JB_FAIL(c, JB_ERR_INVALID, "no spins have been imported");
  int rc = ensure_ready(c); if (rc) return rc;
  std::vector<double> times{time_ps};
  rc = upload_classes(c, times, 0.0, 0.0, 0, term); if (rc) return rc;
  JbTables t; fill_tables(c, t, 0);
  if (!c->has_template) t.nbr_global = nullptr;
  double *dst = h_aos;
  if (!on_device) { rc = ensure_aos(c); if (rc) return rc; dst = c->d_aos; }
  const double *s[3] = {c->S0[0], c->S0[1], c->S0[2]};
  JB_CUDA(c, jbk_field(c->g, t, s, term, c->has_pairs ? c->d_ell_idx : nullptr, c->d_ell_val, c->ell_width, c->d_pair_J, c->pairs_iso ? 1 : 0, dst, c->stream));
  c->launches++;
  if (!on_device) {
    JB_CUDA(c, cudaMemcpyAsync(h_aos, dst, (size_t)c->N * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    JB_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return JB_OK;
}

int jb_energies(jb_ctx *c, int32_t term, double time_ps, double *e, int32_t on_device, double *total) {
  if (!c || term < 0 || term == JB_TERM_TOTAL || term > JB_TERM_UNIAXIAL_3) return JB_ERR_INVALID;
  if (!c->state_allocated) JB_FAIL(c, JB_ERR_INVALID, "no spins have been imported");
  int rc = ensure_ready(c); if (rc) return rc;
  std::vector<double> times{time_ps};
  rc = upload_classes(c, times, 0.0, 0.0, 0, term); if (rc) return rc;
  JbTables t; fill_tables(c, t, 0);
  if (!c->has_template) t.nbr_global = nullptr;
  rc = ensure_scratch(c, (size_t)(c->N + 4096) * sizeof(double)); if (rc) return rc;
  double *partial = c->d_scratch, *tot = c->d_scratch + 2048, *e_dev = nullptr;
  if (e) e_dev = on_device ? e : c->d_scratch + 4096;
  const double *s[3] = {c->S0[0], c->S0[1], c->S0[2]};
  JB_CUDA(c, jbk_energy(c->g, t, s, term, c->has_pairs ? c->d_ell_idx : nullptr, c->d_ell_val, c->ell_width, c->d_pair_J,
                        c->pairs_iso ? 1 : 0, e_dev, partial, tot, c->stream));
  c->launches += 2;
  if (e && !on_device) JB_CUDA(c, cudaMemcpyAsync(e, e_dev, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  double host_total = 0.0;
  JB_CUDA(c, cudaMemcpyAsync(&host_total, tot, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  JB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (total) *total = host_total;
  return JB_OK;
}

int jb_set_magnetisation_groups(jb_ctx *c, int32_t n_groups, const int32_t *group_of_spin) {
  if (!c || n_groups < 1 || (n_groups > 1 && !group_of_spin)) return JB_ERR_INVALID;
  JB_CUDA(c, cudaSetDevice(c->device));
  JB_CUDA(c, cudaStreamSynchronize(c->stream));
  void *p = c->d_groups; free_dev(p); c->d_groups = nullptr; c->n_groups_set = 0;
  if (n_groups == 1) return JB_OK;
  for (int i = 0; i < c->N; ++i) if (group_of_spin[i] < 0 || group_of_spin[i] >= n_groups) JB_FAIL(c, JB_ERR_INVALID, "jb_set_magnetisation_groups: group index out of range");
  JB_CUDA(c, cudaMalloc(&c->d_groups, (size_t)c->N * sizeof(int)));
  JB_CUDA(c, cudaMemcpy(c->d_groups, group_of_spin, (size_t)c->N * sizeof(int), cudaMemcpyHostToDevice));
  c->n_groups_set = n_groups;
  return JB_OK;
}

int jb_magnetisation(jb_ctx *c, int32_t n_groups, const int32_t *group_of_spin, double *M4) {
  if (!c || !M4 || n_groups < 1) return JB_ERR_INVALID;
  if (!c->state_allocated) JB_FAIL(c, JB_ERR_INVALID, "no spins have been imported");
  int rc = ensure_ready(c); if (rc) return rc;
  // mu comes from the class table: make sure it is the current one (upload_classes skips the copy when nothing changed)
  if (c->d_classes == nullptr || c->class_sig.empty()) { std::vector<double> times{0.0}; rc = upload_classes(c, times, 0.0, 0.0, 0, -1); if (rc) return rc; }
  JbTables t; fill_tables(c, t, 0);
  const size_t need = (size_t)(4096 + 4 * n_groups) * sizeof(double) + (group_of_spin ? (size_t)c->N * sizeof(int) : 0);
  rc = ensure_scratch(c, need + 64); if (rc) return rc;
  double *partial = c->d_scratch, *out4 = c->d_scratch + 4096;
  int *d_groups = nullptr;
  if (!group_of_spin && n_groups > 1) {   // the groups the monitor registered once (jb_set_magnetisation_groups)
    if (n_groups != c->n_groups_set) JB_FAIL(c, JB_ERR_INVALID, "jb_magnetisation: no group array given and jb_set_magnetisation_groups was not called for this number of groups");
    d_groups = c->d_groups;
  }
  if (group_of_spin) {
    d_groups = reinterpret_cast<int *>(c->d_scratch + 4096 + 4 * n_groups + 2);
    JB_CUDA(c, cudaMemcpyAsync(d_groups, group_of_spin, (size_t)c->N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  }
  const double *s[3] = {c->S0[0], c->S0[1], c->S0[2]};
  JB_CUDA(c, jbk_magnetisation(c->g, t, s, n_groups, d_groups, partial, out4, c->stream));
  c->launches += 2 * n_groups;
  JB_CUDA(c, cudaMemcpyAsync(M4, out4, (size_t)4 * n_groups * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  JB_CUDA(c, cudaStreamSynchronize(c->stream));
  return JB_OK;
}


// ---- physics hooks: regions of spins (PinnedBoundariesPhysics, physics/pinned_boundaries.cc:12-46) --------------
int jb_set_region(jb_ctx *c, int32_t region, int32_t n, const int32_t *sites) {
  if (!c || region < 0 || region >= JB_MAX_REGIONS || n < 0 || (n > 0 && !sites)) return JB_ERR_INVALID;
  for (int k = 0; k < n; ++k) if (sites[k] < 0 || sites[k] >= c->N) JB_FAIL(c, JB_ERR_INVALID, "jb_set_region: site index out of range");
  JB_CUDA(c, cudaSetDevice(c->device));
  if (c->d_region[region]) { JB_CUDA(c, cudaStreamSynchronize(c->stream)); cudaFree(c->d_region[region]); c->d_region[region] = nullptr; }
  c->region_n[region] = n;
  if (n > 0) {
    JB_CUDA(c, cudaMalloc(&c->d_region[region], (size_t)n * sizeof(int)));
    JB_CUDA(c, cudaMemcpy(c->d_region[region], sites, (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
  }
  return JB_OK;
}

int jb_region_moment(jb_ctx *c, int32_t region, double *M4) {
  if (!c || !M4 || region < 0 || region >= JB_MAX_REGIONS) return JB_ERR_INVALID;
  if (!c->state_allocated) JB_FAIL(c, JB_ERR_INVALID, "no spins have been imported");
  int rc = ensure_ready(c); if (rc) return rc;
  if (c->d_classes == nullptr || c->class_sig.empty()) { std::vector<double> times{0.0}; rc = upload_classes(c, times, 0.0, 0.0, 0, -1); if (rc) return rc; }
  if (c->region_n[region] == 0) { M4[0] = M4[1] = M4[2] = M4[3] = 0.0; return JB_OK; }
  JbTables t; fill_tables(c, t, 0);
  rc = ensure_scratch(c, (size_t)(4096 + 8) * sizeof(double)); if (rc) return rc;
  double *partial = c->d_scratch, *out4 = c->d_scratch + 4096;
  const double *s[3] = {c->S0[0], c->S0[1], c->S0[2]};
  JB_CUDA(c, jbk_region_moment(c->g, t, s, c->d_region[region], c->region_n[region], partial, out4, c->stream));
  c->launches += 2;
  JB_CUDA(c, cudaMemcpyAsync(M4, out4, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  JB_CUDA(c, cudaStreamSynchronize(c->stream));
  return JB_OK;
}

int jb_rotate_region(jb_ctx *c, int32_t region, const double *R9) {
  if (!c || !R9 || region < 0 || region >= JB_MAX_REGIONS) return JB_ERR_INVALID;
  if (!c->state_allocated) JB_FAIL(c, JB_ERR_INVALID, "no spins have been imported");
  int rc = ensure_ready(c); if (rc) return rc;
  const bool multi = c->d.n_ranks > 1 && c->g.gx > 0;
  if (multi && !c->halo_connected) JB_FAIL(c, JB_ERR_INVALID, "multi-rank context: jb_halo_connect has not been called");
  double *lo[3], *hi[3];
  for (int k = 0; k < 3; ++k) {
    if (multi) { lo[k] = c->peer_lo_S0[k]; hi[k] = c->peer_hi_S0[k]; }
    else if (c->g.per[0] && c->g.gx > 0) { lo[k] = c->S0[k]; hi[k] = c->S0[k]; }
    else { lo[k] = nullptr; hi[k] = nullptr; }
  }
  // slab-decomposed runs: the rotated ghost images go into the neighbours' boxes, so the call is one more epoch of the halo
  // handshake -- wait for the neighbours' last stage, rotate, publish.  Every rank calls it for every region (an empty local
  // region still takes part), exactly as every rank runs every stage.
  if (multi) { JB_CUDA(c, jbk_wait(c->flags, c->peer_lo_flags != nullptr, c->peer_hi_flags != nullptr, c->epoch, c->stream)); c->launches++; }
  if (c->region_n[region] > 0) {
    JB_CUDA(c, jbk_region_rotate(c->g, c->S0, lo, hi, c->d_region[region], c->region_n[region], R9, c->stream));
    c->launches++;
  }
  if (multi) {
    c->epoch++;
    JB_CUDA(c, jbk_signal(c->peer_lo_flags ? c->peer_lo_flags + 1 : nullptr, c->peer_hi_flags ? c->peer_hi_flags + 0 : nullptr, c->epoch, c->stream)); c->launches++;
  }
  return JB_OK;
}

// ---- halo plumbing ------------------------------------------------------------------------------------
int jb_halo_export_handle(jb_ctx *c, void *blob_out) {
  if (!c || !blob_out) return JB_ERR_INVALID;
  int rc = ensure_ready(c); if (rc) return rc;
  Blob b{};
  b.magic = 0x4a42484cu;  // "JBHL"
  b.pid = (int32_t)getpid(); b.device = c->device; b.rank = c->d.rank;
  b.nx = c->g.nx; b.PY = c->g.PY; b.PZ = c->g.PZ; b.M = c->g.M; b.gx = c->g.gx;
  b.base_ptr = (uint64_t)(uintptr_t)c->slab;
  for (int k = 0; k < 3; ++k) {
    b.off_S0[k] = (uint64_t)((char *)c->S0[k] - (char *)c->slab);
    b.off_S1[k] = (uint64_t)((char *)c->S1[k] - (char *)c->slab);
    b.off_V[k] = (uint64_t)((char *)c->V[k] - (char *)c->slab);
  }
  b.off_flags = (uint64_t)((char *)c->flags - (char *)c->slab);
  JB_CUDA(c, cudaIpcGetMemHandle(&b.ipc, c->slab));
  memset(blob_out, 0, JB_HALO_HANDLE_BYTES);
  memcpy(blob_out, &b, sizeof(b));
  return JB_OK;
}

static int map_peer(jb_ctx *c, const Blob &b, void **base_out) {
  if (b.magic != 0x4a42484cu) JB_FAIL(c, JB_ERR_PEER, "bad halo handle");
  if (b.nx != c->g.nx || b.PY != c->g.PY || b.PZ != c->g.PZ || b.M != c->g.M || b.gx != c->g.gx)
    JB_FAIL(c, JB_ERR_PEER, "neighbour slab has a different shape (equal slabs are required)");
  if (b.pid == (int32_t)getpid()) {
    // same process (several contexts driven by one host thread): plain pointers, peer access if needed
    if (b.device == c->device) c->peer_on_my_device = true;
    if (b.device != c->device) {
      int can = 0;
      JB_CUDA(c, cudaDeviceCanAccessPeer(&can, c->device, b.device));
      if (!can) JB_FAIL(c, JB_ERR_PEER, "no peer access between the two devices");
      cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) JB_CUDA(c, e);
      cudaGetLastError();
    }
    *base_out = (void *)(uintptr_t)b.base_ptr;
    return JB_OK;
  }
  void *p = nullptr;
  JB_CUDA(c, cudaIpcOpenMemHandle(&p, b.ipc, cudaIpcMemLazyEnablePeerAccess));
  *base_out = p;
  return JB_OK;
}

int jb_halo_connect(jb_ctx *c, const void *blob_lo, const void *blob_hi) {
  if (!c) return JB_ERR_INVALID;
  int rc = ensure_ready(c); if (rc) return rc;
  JB_CUDA(c, cudaSetDevice(c->device));
  Blob lo{}, hi{};
  void *lo_base = nullptr, *hi_base = nullptr;
  bool lo_ipc = false, hi_ipc = false;
  if (blob_lo) { memcpy(&lo, blob_lo, sizeof(Blob)); rc = map_peer(c, lo, &lo_base); if (rc) return rc; lo_ipc = lo.pid != (int32_t)getpid(); }
  if (blob_hi) {
    memcpy(&hi, blob_hi, sizeof(Blob));
    if (blob_lo && lo.pid == hi.pid && lo.base_ptr == hi.base_ptr && lo.rank == hi.rank) { hi_base = lo_base; c->same_peer = true; }
    else { rc = map_peer(c, hi, &hi_base); if (rc) return rc; hi_ipc = hi.pid != (int32_t)getpid(); }
  }
  for (int k = 0; k < 3; ++k) {
    c->peer_lo_S0[k] = lo_base ? (double *)((char *)lo_base + lo.off_S0[k]) : nullptr;
    c->peer_lo_S1[k] = lo_base ? (double *)((char *)lo_base + lo.off_S1[k]) : nullptr;
    c->peer_hi_S0[k] = hi_base ? (double *)((char *)hi_base + hi.off_S0[k]) : nullptr;
    c->peer_hi_S1[k] = hi_base ? (double *)((char *)hi_base + hi.off_S1[k]) : nullptr;
    c->peer_lo_V[k] = lo_base ? (double *)((char *)lo_base + lo.off_V[k]) : nullptr;
    c->peer_hi_V[k] = hi_base ? (double *)((char *)hi_base + hi.off_V[k]) : nullptr;
  }
  c->peer_lo_flags = lo_base ? (unsigned long long *)((char *)lo_base + lo.off_flags) : nullptr;
  c->peer_hi_flags = hi_base ? (unsigned long long *)((char *)hi_base + hi.off_flags) : nullptr;
  c->peer_lo_base = lo_ipc ? lo_base : nullptr;
  c->peer_hi_base = (hi_ipc && !c->same_peer) ? hi_base : nullptr;
  c->halo_connected = true;
  return JB_OK;
}

// ---- introspection ---------------------------------------------------------------------------------------
int64_t jb_kernel_launches(const jb_ctx *c) { return c ? c->launches : 0; }
int jb_stage_kernel(const jb_ctx *c) { return c ? c->last_stage_kernel : -1; }

int jb_synchronize(jb_ctx *c) {
  if (!c) return JB_ERR_INVALID;
  JB_CUDA(c, cudaSetDevice(c->device));
  JB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->flags && c->d.n_ranks > 1) {
    unsigned long long err = 0;
    JB_CUDA(c, cudaMemcpy(&err, c->flags + 2, sizeof(err), cudaMemcpyDeviceToHost));
    if (err) JB_FAIL(c, JB_ERR_PEER, "timed out waiting for a halo signal from a neighbour rank");
  }
  return JB_OK;
}

void *jb_stream(jb_ctx *c) { return c ? (void *)c->stream : nullptr; }

int jb_last_step_kernel_ms(jb_ctx *c, double *out2) {
  if (!c || !out2) return JB_ERR_INVALID;
  JB_CUDA(c, cudaStreamSynchronize(c->stream));
  out2[0] = out2[1] = 0.0;
  for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
    float ms = 0.f;
    JB_CUDA(c, cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]));
    out2[c->ev_kind[i] / 2] += ms;
  }
  c->ev_used = 0;
  c->timed_steps = 0;
  return JB_OK;
}

int64_t jb_timed_steps(const jb_ctx *c) { return c ? c->timed_steps : 0; }

int jb_set_option(jb_ctx *c, const char *key, int64_t value) {
  if (!c || !key) return JB_ERR_INVALID;
  const std::string k(key);
  if (k == "kernel") c->opt_kernel = (int)value;
  else if (k == "recover_u") c->opt_recover_u = (int)value;
  else if (k == "tile_y") c->opt_TY = (int)value;
  else if (k == "tile_z") c->opt_TZ = (int)value;
  else if (k == "motif_split") c->opt_msplit = (int)value;
  else if (k == "rows_warps") c->opt_rows_warps = (int)value;
  else if (k == "ring") c->opt_R = (int)value;
  else if (k == "ring_u") c->opt_RU = (int)value;
  else if (k == "chunks") c->opt_chunks = (int)value;
  else if (k == "chunk_long") c->opt_chunk_long = (int)value;
  else if (k == "chunk_short") c->opt_chunk_short = (int)value;
  else if (k == "tail_pct") c->opt_tail_pct = (int)value;
  else if (k == "face_after") c->opt_face_after = (int)value;
  else if (k == "ctas_per_sm") c->opt_ctas_per_sm = (int)value;
  else if (k == "grid") c->opt_grid = (int)value;
  else if (k == "row_offset") { c->opt_oz = (int)value; c->state_relayout = true; }
  else if (k == "fold_halo") { c->opt_fold_halo = (int)value; return JB_OK; }
  else if (k == "check_symmetry") { c->opt_check_symmetry = (int)value; return JB_OK; }   // check_sparse_matrix_symmetry (hamiltonian/exchange.cc:104-110)
  else if (k == "trace") { c->opt_trace = (int)value; return JB_OK; }
  else if (k == "verbose") { c->opt_verbose = (int)value; return JB_OK; }
  else if (k == "detect_template") { c->opt_detect_template = (int)value; return JB_OK; }
  else if (k == "time_kernels") { c->opt_time_kernels = (int)value; c->ev_used = 0; c->timed_steps = 0; return JB_OK; }  // no re-tiling
  else JB_FAIL(c, JB_ERR_INVALID, "unknown option " + k);
  c->tiling_valid = false;
  c->tmap_valid = false;
  return JB_OK;
}

int jb_plan_work_items(int32_t nx_local, int32_t ghost_x, int32_t n_columns, int32_t n_ctas, int32_t capacity, int32_t *n_chunks,
                       int32_t *x0, int32_t *xc) {
  if (nx_local < 1 || ghost_x < 0 || n_columns < 1 || n_ctas < 1 || !n_chunks || !x0 || !xc || capacity < JB_TILE_MAX_CHUNKS) return JB_ERR_INVALID;
  int bx0[JB_TILE_MAX_CHUNKS], bxc[JB_TILE_MAX_CHUNKS];
  const int n = plan_chunks_core(nx_local, ghost_x, n_columns, n_ctas, ChunkPlanOptions{}, bx0, bxc);
  for (int q = 0; q < n; ++q) { x0[q] = bx0[q]; xc[q] = bxc[q]; }
  *n_chunks = n;
  return JB_OK;
}

int jb_last_stage_trace(jb_ctx *c, uint64_t *out, int32_t capacity, int32_t *n_ctas) {
  uint64_t *out4 = out;
  if (!c || !out4 || !n_ctas || capacity < 0) return JB_ERR_INVALID;
  *n_ctas = 0;
  if (!c->d_trace) return JB_OK;
  JB_CUDA(c, cudaStreamSynchronize(c->stream));
  const int n = std::min(std::min((int)capacity, c->trace_ctas), 4096);
  JB_CUDA(c, cudaMemcpy(out4, c->d_trace, (size_t)n * JB_TRACE_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  *n_ctas = n;
  return JB_OK;
}

}  // extern "C"
