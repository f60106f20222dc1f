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
  std::map<std::array<double, 18>, int> seen;
  c->h_classes.clear(); c->h_class_dc.clear(); c->h_class_ac.clear(); c->h_class_omega.clear(); c->h_class_gyro.clear();
  std::vector<unsigned char> cls(N);
  for (int i = 0; i < N; ++i) {
    std::array<double, 18> key{};
    key[0] = c->h_mus[i]; key[1] = c->h_gyro[i]; key[2] = c->h_alpha[i];
    if (c->uni_power) { key[3] = c->h_K[i]; key[4] = c->h_axis[3 * i]; key[5] = c->h_axis[3 * i + 1]; key[6] = c->h_axis[3 * i + 2]; }
    if (c->has_zeeman) { key[7] = c->h_dc[3 * i]; key[8] = c->h_dc[3 * i + 1]; key[9] = c->h_dc[3 * i + 2]; }
    if (c->has_ac) { key[10] = c->h_ac[3 * i]; key[11] = c->h_ac[3 * i + 1]; key[12] = c->h_ac[3 * i + 2]; key[13] = c->h_omega[i]; }
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

void fill_tables(jb_ctx *c, JbTables &t, int class_table_index) {
  t.nbr_global = c->d_nbr_global; t.Jtab = c->d_Jtab;
  t.classes = c->d_classes + (size_t)class_table_index * c->h_classes.size();
  t.site_class = c->d_site_class;
  for (int m = 0; m <= JB_MAX_MOTIF; ++m) t.nbr_begin[m] = (m <= c->g.M && c->has_template) ? c->nbr_begin[m] : 0;
  for (int m = 0; m < JB_MAX_MOTIF; ++m) t.class_of_motif[m] = c->class_of_motif[m];
  t.n_classes = (int)c->h_classes.size();
  t.iso = c->iso ? 1 : 0;
}

// ---- tiling of the persistent TMA tile kernel ----------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// shape of the tiles; grid size and x-chunking are decided per kernel variant in tile_launch_shape()
void choose_tiling(jb_ctx *c) {
  if (c->tiling_valid) return;
  const JbGeom &g = c->g;
  jb_ctx::Tiling t;
  c->tiling = t;
  c->tiling_valid = true;
  c->tmap_valid = false;
  const int n_nbr = (int)c->t_mi.size();
  if (!c->has_template || !c->motif_uniform || g.gx > JB_TILE_MAX_GX || g.M > JB_TILE_MAX_MOTIF || n_nbr > JB_TILE_MAX_NBR)
    return;
  const bool fused = c->opt_kernel == 3 && c->fused_geometry && g.gx == 2 * c->reach[0] && c->reach[0] == 1;
  const bool pair = c->opt_kernel >= 2;
  t.pair = pair ? 1 : 0;
  if (fused) {
    // fused step kernel: one CTA per SM; the s_n ring (TMA) and four s* planes share the SM's shared memory.  Tile
    // candidates in order of preference (less redundant predictor work first); a thread owns the pair (z, z + 1) of
    // every motif site of one (y) row, consumer threads <= 256.
    const int rx = c->reach[0], ry = c->reach[1], rz = c->reach[2];
    const int cand[][2] = {{8, 64}, {4, 64}, {8, 32}, {4, 32}, {8, 16}, {4, 16}, {2, 16}, {2, 8}, {1, 8}, {1, 4}, {1, 2}};
    const size_t budget = 220 * 1024;
    bool found = false;
    for (const auto &cd : cand) {
      int TY = c->opt_TY ? c->opt_TY : cd[0], TZ = c->opt_TZ ? c->opt_TZ : cd[1];
      TY = std::max(1, std::min(TY, g.Ny));
      TZ = std::max(1, std::min(TZ, g.Nz));
      if (TZ < g.Nz && (TZ & 1)) TZ++;
      const int HZ = (TZ + 1) / 2;
      t.TY = TY; t.TZ = TZ; t.SPT = 1;
      t.e1z = rz > 0 ? 2 : 0; t.e2z = rz > 0 ? 4 : 0;
      t.gzb = t.e2z;
      t.BY = TY + 4 * ry; t.BZ = 2 * HZ + 2 * t.e2z;
      t.UZ = 2 * HZ;
      t.slotS = (t.BY * g.M * t.BZ + 15) / 16 * 16;
      t.slotU = 16;
      t.threads = HZ * TY;
      t.u_tma = 0; t.RU = 2;
      const int n_ct = (t.threads + 31) / 32 * 32;
      const int n_halo = 2 * ry * g.M * (HZ + t.e1z) + TY * g.M * t.e1z;
      const size_t slot_bytes = (size_t)3 * t.slotS * 8 + (size_t)n_nbr * sizeof(JbTileNbr);
      const size_t fixed = 512 + 4 * slot_bytes;   // four s* planes + their entry-table phases + barriers
      int R = c->opt_R ? c->opt_R : (budget > fixed ? (int)((budget - fixed) / slot_bytes) : 0);
      R = std::min(R, JB_PAIR_MAX_RING);
      if (!c->opt_R) R = std::min(R, 2 * rx + 1 + 3);
      t.R = t.Rs[0] = t.Rs[1] = R;
      t.smem[0] = t.smem[1] = fixed + (size_t)R * slot_bytes;
      t.halo_warps = (n_halo + 31) / 32;
      const bool fits = R >= 2 * rx + 2 && t.smem[0] <= budget && n_ct + 32 * t.halo_warps + 32 <= 384 &&
                        t.BY * g.M <= 256 && t.BZ <= 256 && g.oz >= t.e2z;
      if (fits) { found = true; break; }
      if (c->opt_TY && c->opt_TZ) break;
    }
    if (found) t.fused = 1;
    for (const JbClass &cl : c->h_classes) if (cl.power != 0) t.uni = 1;
  }
  if (t.fused) {
    // shape chosen above
  } else if (pair) {
    // pair kernel: a thread owns the sites (z, z + 1) of every motif site of its y rows; consumer threads =
    // ceil(TZ / 2) x ceil(TY / SPT) <= 256 so that two CTAs share an SM.  Tile choice: among z extents 128 / 64 / 32 (long
    // contiguous runs along z serve DRAM best: 4 x 128 beats 8 x 64 beats 16 x 32 on C3, profiles/README.md) take, for
    // each, the tallest tile whose rings fit the shared-memory budget, and keep the candidate with the most sites per
    // plane (a shorter z extent only if it brings 1.5 x the sites).  Motifs with several sites multiply the slot size, so they end up with shorter tiles
    // (bcc: 4 x 64) instead of degenerate one-row tiles.
    int SPT = c->opt_SPT ? c->opt_SPT : 1;
    if (!(SPT == 1 || SPT == 2)) return;
    const size_t budget = c->opt_ctas_per_sm == 1 ? 220 * 1024 : 113 * 1024;   // per CTA
    const int rmin = 2 * g.gx + 2;
    auto shape = [&](int TY, int TZ, jb_ctx::Tiling &q) -> bool {   // fills q; true if the tile fits
      const int HZ = (TZ + 1) / 2;
      q.TY = TY; q.TZ = TZ; q.SPT = (SPT > TY) ? 1 : SPT;
      q.gzb = (g.gz + 1) & ~1;
      q.BY = TY + 2 * g.gy; q.BZ = ((TZ + 1) & ~1) + 2 * q.gzb;
      q.UZ = (TZ + 1) & ~1;
      q.slotS = (q.BY * g.M * q.BZ + 15) / 16 * 16;
      q.slotU = (q.TY * g.M * q.UZ + 15) / 16 * 16;
      q.threads = HZ * ((TY + q.SPT - 1) / q.SPT);
      q.u_tma = 1;
      q.RU = c->opt_RU ? c->opt_RU : 2;
      const size_t slot_bytes = (size_t)3 * q.slotS * 8 + (size_t)n_nbr * sizeof(JbTileNbr);   // ring slot + its phase of the entry table
      const size_t u_bytes = (size_t)q.RU * 3 * q.slotU * 8;
      // noise ring of the noise warp (one-site motifs, SPT 1): two slots of three fp32 per site of the tile
      const size_t n_bytes = (g.M == 1 && q.SPT == 1 && c->opt_noise_warp) ? (size_t)2 * 3 * q.slotU * 4 : 0;
      for (int st = 0; st < 2; ++st) {
        const size_t fixed = 512 + (st == 1 ? u_bytes : 0) + n_bytes;
        int R = c->opt_R ? c->opt_R : (budget > fixed ? (int)((budget - fixed) / slot_bytes) : 0);
        R = std::min(R, JB_PAIR_MAX_RING);
        // one plane in flight per CTA is the measured optimum: with the stores in the mix, more outstanding plane loads
        // lower the DRAM efficiency (ring 5 / 6: -8 % / -15 %, profiles/README.md r01f)
        if (!c->opt_R) R = std::min(R, rmin);
        q.Rs[st] = R;
        q.smem[st] = fixed + (size_t)R * slot_bytes + (size_t)c->opt_smem_pad * 1024;
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
      int TY = c->opt_TY ? c->opt_TY : std::max(SPT, (256 * SPT) / HZ);
      TY = std::max(1, std::min(TY, std::min(g.Ny, 64)));
      jb_ctx::Tiling q = t;
      bool ok = false;
      const int ty_max = TY;
      for (; TY >= 1; TY = c->opt_TY ? 0 : TY - 1) { if (shape(TY, TZ, q)) { ok = true; break; } }
      if (ok) {
        const long long sites = (long long)q.TY * q.TZ * g.M;
        if (2 * sites > 3 * best_sites) { best_sites = sites; shrunk = q.TY < ty_max;   // a shorter z extent must bring 1.5 x the sites per plane
          const int pair_flag = t.pair; t = q; t.pair = pair_flag; found = true; }
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
  } else {
  int TZ = c->opt_TZ ? c->opt_TZ : (g.Nz >= 64 ? 64 : (g.Nz >= 32 ? 32 : g.Nz));
  TZ = std::max(1, std::min(TZ, g.Nz));
  if (TZ < g.Nz && (TZ & 1)) TZ++;   // several z tiles: their first column must stay 16-byte aligned for TMA
  int SPT = c->opt_SPT ? c->opt_SPT : 1;
  // consumer threads per CTA: <= 480 (+ the producer warp = 512 threads at 64 registers, two CTAs per SM) for SPT 1, 256 for SPT 2
  int TY = c->opt_TY ? c->opt_TY : std::max(SPT, ((SPT == 1 ? 480 : 256) * SPT) / TZ);
  TY = std::max(1, std::min(TY, g.Ny));
  if (TY > 64) TY = 64;
  while (SPT > 1 && (SPT > TY)) SPT /= 2;
  if (!(SPT == 1 || SPT == 2)) return;
  t.R = c->opt_R ? c->opt_R : 2 * g.gx + 2;
  t.RU = c->opt_RU ? c->opt_RU : 2;
  if (t.R < 2 * g.gx + 2 || t.R > 8 || t.RU < 2 || t.RU > 8) return;
  for (;;) {
    t.TY = TY; t.TZ = TZ; t.SPT = SPT;
    t.gzb = (g.gz + 1) & ~1;
    t.BY = TY + 2 * g.gy; t.BZ = TZ + 2 * t.gzb; if (t.BZ & 1) t.BZ++;
    t.UZ = (TZ + 1) & ~1;
    t.slotS = (t.BY * g.M * t.BZ + 15) / 16 * 16;
    t.slotU = (t.TY * g.M * t.UZ + 15) / 16 * 16;
    t.threads = TZ * ((TY + SPT - 1) / SPT);
    t.smem[0] = (size_t)t.R * 3 * t.slotS * 8 + 512 + (size_t)t.R * n_nbr * sizeof(JbTileNbr);
    t.u_tma = c->opt_u_tma ? 1 : 0;
    t.smem[1] = t.smem[0] + (t.u_tma ? (size_t)t.RU * 3 * t.slotU * 8 : 0);
    // wanted: two CTAs per SM in stage B
    if ((t.smem[1] <= 110 * 1024 && t.threads <= (SPT == 1 ? 480 : 256)) || c->opt_TY || TY <= SPT) break;
    TY = std::max(SPT, TY / 2);
  }
  if (t.threads > (t.SPT == 1 ? 480 : 256) || t.BY * g.M > 256 || t.BZ > 256 || t.UZ > 256 || t.TY * g.M > 256) return;
  if (t.smem[1] > 220 * 1024) return;
  t.Rs[0] = t.Rs[1] = t.R;
  }
  t.n_yt = (g.Ny + t.TY - 1) / t.TY; t.n_zt = (g.Nz + t.TZ - 1) / t.TZ;
  t.n_cols = t.n_yt * t.n_zt;
  t.ok = true;
  c->tiling = t;

  // tile-relative neighbour table, per motif site in the reference's CSR column order; couplings in Tesla
  c->tile_nbr.assign(n_nbr, JbTileNbr{});
  c->tile_nbr_begin.assign(g.M + 1, 0);
  std::vector<double> J9T(9 * (size_t)std::max(1, n_nbr));
  // pair kernel: within a motif site the entries with an even z offset come first (their neighbour pair is 16-byte
  // aligned in shared memory), then the odd ones; both groups keep the CSR column order
  std::vector<int> order(c->tile_order.begin(), c->tile_order.end());
  if (t.pair)
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (c->t_mi[a] != c->t_mi[b]) return c->t_mi[a] < c->t_mi[b];
      return (c->t_T[3 * a + 2] & 1) < (c->t_T[3 * b + 2] & 1);
    });
  c->tile_nbr_odd.assign(g.M + 1, 0);
  for (int pos = 0; pos < n_nbr; ++pos) {
    const int k = order[pos];
    const int mi = c->t_mi[k], mj = c->t_mj[k];
    const int Tx = c->t_T[3 * k], Ty = c->t_T[3 * k + 1], Tz = c->t_T[3 * k + 2];
    const double inv_mu = c->h_classes[c->class_of_motif[mi]].inv_mu;
    if (!(Tz & 1)) c->tile_nbr_odd[mi]++;   // number of even entries for now
    JbTileNbr e{};
    e.delta = (Ty * g.M + (mj - mi)) * t.BZ + Tz;
    e.d = Tx + (t.fused ? c->reach[0] : g.gx);
    e.J = c->t_J9[9 * k] * inv_mu;
    for (int q = 0; q < 9; ++q) J9T[9 * (size_t)pos + q] = c->t_J9[9 * (size_t)k + q] * inv_mu;
    c->tile_nbr[pos] = e;
    c->tile_nbr_begin[mi + 1]++;
  }
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

void fill_tile_params(jb_ctx *c, JbTileParams &p) {
  const jb_ctx::Tiling &t = c->tiling;
  p.g = c->g;
  p.J9T = c->d_tile_J9T;
  p.TY = t.TY; p.TZ = t.TZ; p.UZ = t.UZ; p.BY = t.BY; p.BZ = t.BZ; p.gzb = t.gzb; p.slotS = t.slotS; p.slotU = t.slotU; p.R = t.R; p.RU = t.RU;
  p.n_yt = t.n_yt; p.n_zt = t.n_zt; p.n_cols = t.n_cols; p.u_tma = t.u_tma;
  for (size_t q = 0; q < c->tile_nbr_begin.size(); ++q) p.nbr_begin[q] = c->tile_nbr_begin[q];
  for (int q = 0; q < c->g.M; ++q) p.nbr_odd[q] = t.pair ? c->tile_nbr_odd[q] : c->tile_nbr_begin[q + 1];
  p.nbr = c->d_tile_nbr;
  p.n_nbr = (int)c->tile_nbr.size();
  p.rx = c->reach[0]; p.ry = c->reach[1]; p.rz = c->reach[2]; p.e1z = t.e1z; p.e2z = t.e2z;
  p.producer_sleep_ns = c->opt_producer_sleep;
  p.split_wait = c->opt_split_wait;
  p.debug_skip = c->opt_debug_skip;
  p.early_release = c->opt_early_release;
  p.store_hint = c->opt_store_hint;
  p.load_hint = c->opt_load_hint;
  for (int m = 0; m < c->g.M; ++m) {
    int n = c->tile_nbr_begin[m];
    while (n < c->tile_nbr_begin[m + 1] && c->tile_nbr[n].d < 2 * c->g.gx) ++n;
    p.nbr_split[m] = n;
  }
}

// grid size (resident CTAs) and number of x-chunks for one kernel variant
int tile_launch_shape(jb_ctx *c, const JbTileParams &p, int stage, int thermal) {
  jb_ctx::Tiling &t = c->tiling;
  if (t.grid[stage][thermal] > 0) return JB_OK;
  if (c->num_sms == 0) JB_CUDA(c, cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device));
  int per_sm = 0;
  if (t.fused) JB_CUDA(c, jbk_step_fused_occupancy(p, thermal, c->iso ? 1 : 0, t.uni, t.threads, t.halo_warps, t.smem[stage], &per_sm));
  else if (t.pair) JB_CUDA(c, jbk_stage_pair_occupancy(p, stage, thermal, c->iso ? 1 : 0, t.SPT, t.threads, t.smem[stage], &per_sm));
  else JB_CUDA(c, jbk_stage_tile_occupancy(p, stage, thermal, c->iso ? 1 : 0, t.SPT, t.threads, t.smem[stage], &per_sm));
  if (per_sm < 1) JB_FAIL(c, JB_ERR_CUDA, "the tile kernel does not fit on an SM with this tiling");
  if (c->opt_ctas_per_sm > 0) per_sm = std::min(per_sm, c->opt_ctas_per_sm);
  int G = per_sm * c->num_sms;
  if (c->opt_grid > 0) G = std::min(G, c->opt_grid);   // experiments: fewer resident CTAs than the occupancy calculation allows
  const JbGeom &g = c->g;
  int best_c = 1;
  if (c->opt_chunks > 0) {
    best_c = std::min(c->opt_chunks, g.nx);
  } else {
    // cost model: time ~ (items per CTA, rounded up) x (planes marched per item + load-only halo planes)
    double best = 1e300;
    for (int nc = 1; nc <= g.nx; ++nc) {
      const long long items = (long long)nc * t.n_cols;
      const long long per_cta = (items + G - 1) / G;
      const int xc = (g.nx + nc - 1) / nc;
      // planes marched per item + what an item costs on top: halo planes that are only loaded (two-launch kernels) or
      // the extra predictor planes of the fused kernel
      const double cost = (double)per_cta * (xc + (t.fused ? 2.0 * c->reach[0] + 0.6 : 0.7 * 2 * g.gx + 0.3));
      if (cost < best * 0.999) { best = cost; best_c = nc; }
    }
  }
  t.n_chunks[stage][thermal] = best_c;
  t.grid[stage][thermal] = (int)std::min<long long>(G, (long long)best_c * t.n_cols);
  if (c->opt_verbose)
    fprintf(stderr, "jams_b200: %s kernel stage %d thermal %d: tile %dx%d (y,z) spt %d, %d consumer threads, ring %d/%d, smem %zu B, "
                    "%d CTAs/SM -> grid %d, %d x-chunks x %d columns\n", t.fused ? "fused step" : (t.pair ? "pair" : "tile"), stage, thermal, t.TY, t.TZ, t.SPT, t.threads, t.Rs[stage], t.RU,
            t.smem[stage], per_sm, t.grid[stage][thermal], best_c, t.n_cols);
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
  for (int a = 0; a < 5; ++a) {
    const cuuint32_t box[3] = {(cuuint32_t)(a < 2 ? t.BZ : t.UZ), (cuuint32_t)((a < 2 ? t.BY : t.TY) * g.M), 1};
    for (int k = 0; k < 3; ++k) {
      double *base = a == 0 || a == 3 ? c->S0[k] : (a == 1 || a == 4 ? c->S1[k] : c->U[k]);
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
  const JbGeom &g = c->g;
  // the fused step kernel (jb_step_fused.cu) needs ghost zones twice as deep as the template reaches; it handles reach 1
  // along x, reach <= 1 along y and z and motifs of one or two sites.  A periodic axis must then be at least 2 x ghost
  // depth long (no site may sit on both faces) and the slab at least a ghost depth thick.
  const int reach[3] = {gx, gy, gz};
  bool fused = c->opt_kernel == 3 && c->has_template && !c->has_pairs && gx == 1 && gy <= 1 && gz <= 1 && c->d.num_motif <= 2;
  {
    const int ext[3] = {c->d.dims[0], c->d.dims[1], c->d.dims[2]};
    for (int d = 0; d < 3; ++d) if (reach[d] > 0 && c->d.periodic[d] && ext[d] < 4 * reach[d]) fused = false;
    if (c->d.nx_local < 2 * gx || (c->d.n_ranks == 1 && c->d.periodic[0] && c->d.nx_local < 4 * gx)) fused = false;
  }
  if (fused) { gx *= 2; gy *= 2; gz *= 2; }
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
    c->fused_geometry = fused;
    int rc = allocate_state(c); if (rc) return rc;
    if (had_state) {
      double *dst[3] = {c->S0[0], c->S0[1], c->S0[2]};
      JB_CUDA(c, jbk_import(c->g, c->d_aos, dst, c->d.n_ranks == 1, c->stream)); c->launches++;
    }
    c->tables_built = false;
    c->tiling_valid = false;
    c->classes_dirty = true;
  }
  for (int d = 0; d < 3; ++d) c->reach[d] = reach[d];
  const bool classes_were_dirty = c->classes_dirty;
  int rc = build_classes(c); if (rc) return rc;
  if (classes_were_dirty) c->tiling_valid = false;
  if (c->has_template && !c->tables_built) { rc = build_template_tables(c); if (rc) return rc; }
  return JB_OK;
}

void record_event(jb_ctx *c, int kind) {
  if (!c->opt_time_kernels) return;
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
  p = c->d_nbr_global; free_dev(p); p = c->d_Jtab; free_dev(p); p = c->d_tile_nbr; free_dev(p); p = c->d_tile_J9T; free_dev(p);
  p = c->d_classes; free_dev(p); p = c->d_site_class; free_dev(p);
  p = c->d_ell_idx; free_dev(p); p = c->d_ell_val; free_dev(p); p = c->d_pair_J; free_dev(p);
  for (int r = 0; r < JB_MAX_REGIONS; ++r) { p = c->d_region[r]; free_dev(p); }
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  for (auto ev : c->ev) cudaEventDestroy(ev);
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
  if (!c || n < 0 || (n > 0 && (!mi || !mj || !T3 || !J9))) return JB_ERR_INVALID;
  for (int k = 0; k < n; ++k) {
    if (mi[k] < 0 || mi[k] >= c->d.num_motif || mj[k] < 0 || mj[k] >= c->d.num_motif) JB_FAIL(c, JB_ERR_INVALID, "motif index out of range in exchange template");
  }
  c->t_mi.assign(mi, mi + n); c->t_mj.assign(mj, mj + n); c->t_T.assign(T3, T3 + 3 * n); c->t_J9.assign(J9, J9 + 9 * (size_t)n);
  c->has_template = n > 0;
  c->has_pairs = false;
  c->tables_built = false;
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
  if (c->d.n_ranks != 1) JB_FAIL(c, JB_ERR_UNSUPPORTED, "jb_set_exchange_pairs is single-rank only in this version");
  JB_CUDA(c, cudaSetDevice(c->device));
  c->has_template = false; c->t_mi.clear(); c->t_mj.clear(); c->t_T.clear(); c->t_J9.clear();
  c->has_pairs = false;
  int rc = ensure_ready(c);  // geometry without ghosts
  if (rc) return rc;
  const JbGeom &g = c->g;
  const int N = c->N;
  std::vector<int> count(N, 0);
  for (int64_t p = 0; p < n_pairs; ++p) {
    if (pi[p] < 0 || pi[p] >= N || pj[p] < 0 || pj[p] >= N || vid[p] < 0 || vid[p] >= n_values) JB_FAIL(c, JB_ERR_INVALID, "pair index out of range");
    count[pi[p]]++;
  }
  int width = 0;
  for (int i = 0; i < N; ++i) width = std::max(width, count[i]);
  auto layout_q = [&](int ref) {  // reference site id -> interior layout order q and ghosted index
    const int m = ref % g.M; int r = ref / g.M; const int z = r % g.Nz; r /= g.Nz; const int y = r % g.Ny; const int x = r / g.Ny;
    const long long q = (((long long)x * g.Ny + y) * g.M + m) * g.Nz + z;
    const long long gi = ((long long)(x + g.gx) * g.PY + (y + g.gy)) * g.sY + (long long)m * g.PZ + (z + g.oz);
    return std::make_pair(q, gi);
  };
  std::vector<int> idx((size_t)width * N, -1), val((size_t)width * N, 0), fill(N, 0);
  for (int64_t p = 0; p < n_pairs; ++p) {  // pairs arrive sorted by {i,j}: ascending-j order is kept per row
    const auto qi = layout_q(pi[p]);
    const auto qj = layout_q(pj[p]);
    const int e = fill[pi[p]]++;
    idx[(size_t)e * N + qi.first] = (int)qj.second;
    val[(size_t)e * N + qi.first] = vid[p];
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
  const double *src = s_aos;
  if (!on_device) {
    rc = ensure_aos(c); if (rc) return rc;
    JB_CUDA(c, cudaMemcpyAsync(c->d_aos, s_aos, (size_t)c->N * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    src = c->d_aos;
  }
  double *dst[3] = {c->S0[0], c->S0[1], c->S0[2]};
  JB_CUDA(c, jbk_import(c->g, src, dst, c->d.n_ranks == 1, c->stream)); c->launches++;
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
  JB_CUDA(c, jbk_export(c->g, src, c->d_aos, c->stream)); c->launches++;
  JB_CUDA(c, cudaMemcpyAsync(s_aos, c->d_aos, (size_t)c->N * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  JB_CUDA(c, cudaStreamSynchronize(c->stream));
  return JB_OK;
}

int jb_step(jb_ctx *c, int32_t nsteps, double dt, double time_ps, double T, uint64_t seed, uint64_t first_step, int32_t gilbert) {
  if (!c || nsteps < 0 || !(dt > 0.0) || T < 0.0) return JB_ERR_INVALID;
  if (!c->state_allocated) JB_FAIL(c, JB_ERR_INVALID, "no spins have been imported");
  int rc = ensure_ready(c); if (rc) return rc;
  if (c->d.n_ranks > 1 && c->g.gx > 0 && !c->halo_connected) JB_FAIL(c, JB_ERR_INVALID, "multi-rank context: jb_halo_connect has not been called");
  const bool multi = c->d.n_ranks > 1 && c->g.gx > 0;

  choose_tiling(c);
  const bool use_tile = c->opt_kernel >= 1 && c->tiling.ok && !c->has_pairs;
  JbTileParams tp{};
  if (use_tile) {
    rc = build_tmaps(c); if (rc) return rc;
    fill_tile_params(c, tp);
    for (int r = 0; r < 10; ++r) {   // Philox4x32-10 key schedule (Salmon et al.)
      tp.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
      tp.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
  }
  const int thermal = T > 0.0 ? 1 : 0;

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

    for (int n = 0; n < chunk && use_tile && c->tiling.fused; ++n) {
      // ---- fused step kernel: one launch per Heun step, reads S0 (s_n) and writes S1 (s_{n+1}); then the roles swap ----
      for (int k = 0; k < 3; ++k) {
        tp.out[k] = c->S1[k]; tp.u[k] = nullptr;
        if (multi) { tp.out_lo[k] = c->peer_lo_S1[k]; tp.out_hi[k] = c->peer_hi_S1[k]; }
        else if (c->g.per[0] && c->g.gx > 0) { tp.out_lo[k] = c->S1[k]; tp.out_hi[k] = c->S1[k]; }
        else { tp.out_lo[k] = nullptr; tp.out_hi[k] = nullptr; }
      }
      tp.step = first_step + (uint64_t)(done + n);
      const size_t nc = c->h_classes.size();
      const JbClass *cls0 = c->h_class_tab.data() + (size_t)(time_dependent(c) ? 2 * n : 0) * nc;       // fields at t      (cpu_llg_heun.cc:66)
      const JbClass *cls1 = c->h_class_tab.data() + (size_t)(time_dependent(c) ? 2 * n + 1 : 0) * nc;   // fields at t + dt (:103-106)
      for (int m = 0; m < c->g.M; ++m) {
        tp.cls[m] = cls0[c->class_of_motif[m]];
        const JbClass &b = cls1[c->class_of_motif[m]];
        tp.fT1[m][0] = b.fTx; tp.fT1[m][1] = b.fTy; tp.fT1[m][2] = b.fTz;
      }
      tp.R = c->tiling.Rs[0];
      rc = tile_launch_shape(c, tp, 0, thermal); if (rc) return rc;
      tp.n_chunks = c->tiling.n_chunks[0][thermal];
      tp.n_items = tp.n_chunks * tp.n_cols;
      if (multi) { JB_CUDA(c, jbk_wait(c->flags, c->peer_lo_flags != nullptr, c->peer_hi_flags != nullptr, c->epoch, c->stream)); c->launches++; }
      record_event(c, 0);
      const CUtensorMap tm[3] = {c->tmap[0][0], c->tmap[0][1], c->tmap[0][2]};
      JB_CUDA(c, jbk_step_fused(tp, tm, thermal, c->iso ? 1 : 0, c->tiling.uni, c->tiling.threads, c->tiling.halo_warps, c->tiling.grid[0][thermal],
                                c->tiling.smem[0], c->stream));
      c->launches++;
      record_event(c, 1);
      if (multi) {
        c->epoch++;
        JB_CUDA(c, jbk_signal(c->peer_lo_flags ? c->peer_lo_flags + 1 : nullptr, c->peer_hi_flags ? c->peer_hi_flags + 0 : nullptr, c->epoch, c->stream)); c->launches++;
      }
      for (int k = 0; k < 3; ++k) {   // s_{n+1} becomes the current state
        std::swap(c->S0[k], c->S1[k]);
        std::swap(c->tmap[0][k], c->tmap[1][k]);
        std::swap(c->tmap[3][k], c->tmap[4][k]);
        std::swap(c->peer_lo_S0[k], c->peer_lo_S1[k]);
        std::swap(c->peer_hi_S0[k], c->peer_hi_S1[k]);
      }
    }
    for (int n = 0; n < chunk && !(use_tile && c->tiling.fused); ++n) {
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
        // the corrector needs no noise: the predictor folds the noise part of its right-hand side into u (jb_device.cuh)
        // (with recover_u the pair kernel stores no u, and its corrector draws the noise itself)
        // (option recover_u: 2 = where it is measured faster, i.e. at T = 0 -- at T > 0 the second noise draw costs what the 24 B save)
        const bool recu = use_tile && c->tiling.pair && !c->has_pairs && (c->opt_recover_u == 1 || (c->opt_recover_u == 2 && !thermal));
        const int th = (stage == 0 || recu) ? thermal : 0;
        p.thermal = th;
        if (multi) {
          // ghosts I read were written by the neighbours' previous stage; the boxes I write into were
          // last read by the neighbours' previous stage: both are covered by their last signal
          JB_CUDA(c, jbk_wait(c->flags, c->peer_lo_flags != nullptr, c->peer_hi_flags != nullptr, c->epoch, c->stream)); c->launches++;
        }
        record_event(c, 2 * stage);
        if (c->has_pairs) {
          JB_CUDA(c, jbk_stage_pairs(p, c->d_ell_idx, c->d_ell_val, c->ell_width, c->d_pair_J, c->pairs_iso ? 1 : 0, stage, c->stream));
        } else if (use_tile) {
          for (int k = 0; k < 3; ++k) { tp.out[k] = p.out[k]; tp.out_lo[k] = p.out_lo[k]; tp.out_hi[k] = p.out_hi[k]; tp.u[k] = p.u[k]; }
          tp.step = p.step;
          tp.recover_u = recu ? 1 : 0;
          tp.noise_warp = (c->tiling.pair && c->g.M == 1 && c->tiling.SPT == 1) ? c->opt_noise_warp : 0;
          const JbClass *cls = c->h_class_tab.data() + (size_t)(time_dependent(c) ? 2 * n + stage : 0) * c->h_classes.size();
          for (int m = 0; m < c->g.M; ++m) tp.cls[m] = cls[c->class_of_motif[m]];
          tp.R = c->tiling.Rs[stage];
          rc = tile_launch_shape(c, tp, stage, th); if (rc) return rc;
          tp.n_chunks = c->tiling.n_chunks[stage][th];
          tp.n_items = tp.n_chunks * tp.n_cols;
          const int ua = recu ? 3 : 2;   // recover_u: the corrector's second ring carries the tile's own s_n (S0) instead of u
          const CUtensorMap tm[6] = {c->tmap[stage][0], c->tmap[stage][1], c->tmap[stage][2], c->tmap[ua][0], c->tmap[ua][1], c->tmap[ua][2]};
          tp.reverse_items = (c->tiling.pair && stage == 1 && c->opt_reverse_b) ? 1 : 0;
          if (c->tiling.pair)
            JB_CUDA(c, jbk_stage_pair(tp, tm, stage, th, c->iso ? 1 : 0, c->tiling.SPT, c->tiling.threads,
                                      c->tiling.grid[stage][th], c->tiling.smem[stage], c->stream));
          else
            JB_CUDA(c, jbk_stage_tile(tp, tm, stage, th, c->iso ? 1 : 0, c->tiling.SPT, c->tiling.threads,
                                      c->tiling.grid[stage][th], c->tiling.smem[stage], c->stream));
        } else {
          JB_CUDA(c, jbk_stage_direct(p, stage, c->stream));
        }
        c->launches++;
        record_event(c, 2 * stage + 1);
        if (multi) {
          c->epoch++;
          JB_CUDA(c, jbk_signal(c->peer_lo_flags ? c->peer_lo_flags + 1 : nullptr, c->peer_hi_flags ? c->peer_hi_flags + 0 : nullptr, c->epoch, c->stream)); c->launches++;
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
        if (multi) {   // as in jb_step: the neighbours' previous stage wrote the ghosts this stage reads and read the boxes it writes
          JB_CUDA(c, jbk_wait(c->flags, c->peer_lo_flags != nullptr, c->peer_hi_flags != nullptr, c->epoch, c->stream)); c->launches++;
        }
        record_event(c, 0);
        JB_CUDA(c, jbk_rk4_stage_direct(p, stage, c->stream));
        c->launches++;
        record_event(c, 1);
        if (multi) {
          c->epoch++;
          JB_CUDA(c, jbk_signal(c->peer_lo_flags ? c->peer_lo_flags + 1 : nullptr, c->peer_hi_flags ? c->peer_hi_flags + 0 : nullptr, c->epoch, c->stream)); c->launches++;
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
  if (!c || !h_aos || term < 0 || term > JB_TERM_TOTAL) return JB_ERR_INVALID;
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
  if (!c || term < 0 || term >= JB_TERM_TOTAL) return JB_ERR_INVALID;
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

int jb_magnetisation(jb_ctx *c, int32_t n_groups, const int32_t *group_of_spin, double *M4) {
  if (!c || !M4 || n_groups < 1) return JB_ERR_INVALID;
  if (!c->state_allocated) JB_FAIL(c, JB_ERR_INVALID, "no spins have been imported");
  int rc = ensure_ready(c); if (rc) return rc;
  if (c->d_classes == nullptr) { std::vector<double> times{0.0}; rc = upload_classes(c, times, 0.0, 0.0, 0, -1); if (rc) return rc; }
  JbTables t; fill_tables(c, t, 0);
  const size_t need = (size_t)(4096 + 4 * n_groups) * sizeof(double) + (group_of_spin ? (size_t)c->N * sizeof(int) : 0);
  rc = ensure_scratch(c, need + 64); if (rc) return rc;
  double *partial = c->d_scratch, *out4 = c->d_scratch + 4096;
  int *d_groups = nullptr;
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
  if (c->d_classes == nullptr) { std::vector<double> times{0.0}; rc = upload_classes(c, times, 0.0, 0.0, 0, -1); if (rc) return rc; }
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
  JB_CUDA(c, jbk_region_rotate(c->g, c->S0, lo, hi, c->d_region[region], c->region_n[region], R9, c->stream));
  c->launches++;
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
  return JB_OK;
}

int jb_set_option(jb_ctx *c, const char *key, int64_t value) {
  if (!c || !key) return JB_ERR_INVALID;
  const std::string k(key);
  if (k == "kernel") c->opt_kernel = (int)value;
  else if (k == "recover_u") c->opt_recover_u = (int)value;
  else if (k == "noise_warp") c->opt_noise_warp = (int)value;
  else if (k == "tile_y") c->opt_TY = (int)value;
  else if (k == "tile_z") c->opt_TZ = (int)value;
  else if (k == "spt") c->opt_SPT = (int)value;
  else if (k == "ring") c->opt_R = (int)value;
  else if (k == "ring_u") c->opt_RU = (int)value;
  else if (k == "chunks") c->opt_chunks = (int)value;
  else if (k == "ctas_per_sm") c->opt_ctas_per_sm = (int)value;
  else if (k == "grid") c->opt_grid = (int)value;
  else if (k == "u_tma") c->opt_u_tma = (int)value;
  else if (k == "smem_pad") c->opt_smem_pad = (int)value;
  else if (k == "row_offset") { c->opt_oz = (int)value; c->state_relayout = true; }   // experiments: unused dynamic shared memory (KB) per CTA
  else if (k == "producer_sleep") { c->opt_producer_sleep = (int)value; return JB_OK; }
  else if (k == "split_wait") { c->opt_split_wait = (int)value; return JB_OK; }
  else if (k == "verbose") { c->opt_verbose = (int)value; return JB_OK; }
  else if (k == "debug_skip") { c->opt_debug_skip = (int)value; return JB_OK; }
  else if (k == "early_release") { c->opt_early_release = (int)value; return JB_OK; }
  else if (k == "store_hint") { c->opt_store_hint = (int)value; return JB_OK; }
  else if (k == "load_hint") { c->opt_load_hint = (int)value; return JB_OK; }
  else if (k == "reverse_b") { c->opt_reverse_b = (int)value; return JB_OK; }
  else if (k == "detect_template") { c->opt_detect_template = (int)value; return JB_OK; }
  else if (k == "time_kernels") { c->opt_time_kernels = (int)value; c->ev_used = 0; return JB_OK; }  // no re-tiling
  else JB_FAIL(c, JB_ERR_INVALID, "unknown option " + k);
  c->tiling_valid = false;
  c->tmap_valid = false;
  return JB_OK;
}

}  // extern "C"
