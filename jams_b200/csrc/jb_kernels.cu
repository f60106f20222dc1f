// jb_kernels.cu — hand-written sm_100a kernels of the llg-heun + exchange hot path.
//
// Replaces, fused into two launches per Heun step (SURVEY.md 2 "CUDA kernel / library-call inventory"):
//   cusparseSpMV (containers/sparse_matrix.h:366-379), cuda_uniaxial_field_kernel
//   (hamiltonian/cuda_uniaxial_anisotropy_kernel.cuh:14-26), the Zeeman D2D copy + ac kernel
//   (hamiltonian/cuda_zeeman.cu:27-41), cudaMemcpy + cublasDaxpy field summation (cuda/cuda_solver.cc:11-26),
//   curandGenerateNormalDouble + scale (thermostats/cuda_thermostat_classical.cc:47-56), the s -> s_old
//   snapshot (solvers/cuda_llg_heun.cu:71-75) and cuda_heun_llg_kernelA/B (solvers/cuda_llg_heun_kernel.cuh:8-104).
// The arithmetic follows the CPU solver (solvers/cpu_llg_heun.cc:45-148), see DESIGN.md.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

#include "jb_device.cuh"

namespace {

using namespace jbdev;

// own cell + ghost images (see jbdev::store_images)
__device__ __forceinline__ void store_with_images(const JbStageParams &p, int x, int y, int m, int z,
                                                  double vx, double vy, double vz) {
  const JbGeom &g = p.g;
  const long long i0 = gidx(g, x + g.gx, y + g.gy, m, z + g.oz);
  p.out[0][i0] = vx; p.out[1][i0] = vy; p.out[2][i0] = vz;
  if (!(x_image_needed(g, x) | yz_image_needed(g, y, z))) return;
  JbOutBoxes o;
#pragma unroll
  for (int c = 0; c < 3; ++c) { o.out[c] = p.out[c]; o.out_lo[c] = p.out_lo[c]; o.out_hi[c] = p.out_hi[c]; }
  store_images(g, o, x, y, m, z, vx, vy, vz);
}

// interior flat index q (layout order [x][y][m][z]) -> coordinates
__device__ __forceinline__ void decode_site(const JbGeom &g, long long q64, int &x, int &y, int &m, int &z) {
  unsigned q = (unsigned)q64;  // N < 2^31 per context (jb_create): 32-bit divisions
  const unsigned uz = (unsigned)g.Nz, um = (unsigned)g.M, uy = (unsigned)g.Ny;
  unsigned t = q / uz; z = (int)(q - t * uz); q = t;
  t = q / um; m = (int)(q - t * um); q = t;
  t = q / uy; y = (int)(q - t * uy);
  x = (int)t;
}

__device__ __forceinline__ long long ref_site_local(const JbGeom &g, int x, int y, int m, int z) {
  return (((long long)x * g.Ny + y) * g.Nz + z) * g.M + m;
}

// a second / third uniaxial Hamiltonian (jb_set_uniaxial_term, JbUniExtra): H = K p (s.a)^(p-1) a of the slots in `slots` (bit
// q = slot q + 1) added to (hx, hy, hz), in Tesla (K p / mu) or meV (uniaxial_anisotropy.cc:155-163)
__device__ __forceinline__ void uniaxial_extra_site(const JbTables &t, int ci, int slots, bool tesla, double sx, double sy, double sz,
                                                    double &hx, double &hy, double &hz) {
  if (!t.uni_extra) return;
  for (int q = 0; q < JB_MAX_UNIAXIAL - 1; ++q) {
    if (!((slots >> q) & 1)) continue;
    const JbUniExtra u = t.uni_extra[ci * (JB_MAX_UNIAXIAL - 1) + q];
    if (u.power == 0) continue;
    const double d = u.ax * sx + u.ay * sy + u.az * sz;
    double pw = d;
    if (u.power >= 4) pw = d * d * d;
    if (u.power >= 6) pw = pw * d * d;
    const double f = (tesla ? u.KpT : u.Kp) * pw;
    hx = fma(f, u.ax, hx); hy = fma(f, u.ay, hy); hz = fma(f, u.az, hz);
  }
}

// =================================================================================================
// import / export between the reference's AoS site order and the ghosted SoA box
// =================================================================================================
// One block per (x, y, z-chunk of ZC cells): the (z, m) block of a lattice column is contiguous in the reference's AoS order
// (3 M ZC doubles) and M x 3 contiguous runs of ZC doubles in the SoA box, so both sides of the transposition are coalesced and
// the permutation happens in shared memory.  Planes [x_begin, x_end) only: the host path pipelines x-chunks with the PCIe copies.
__global__ void __launch_bounds__(256) import_rows_kernel(const JbGeom g, const double *__restrict__ aos, double *__restrict__ dx,
                                                          double *__restrict__ dy, double *__restrict__ dz, int x_begin, int n_zc, int ZC) {
  extern __shared__ double tile[];   // [zz][m][c]
  const int zc = blockIdx.x % n_zc;
  const int y = (blockIdx.x / n_zc) % g.Ny;
  const int x = x_begin + blockIdx.x / (n_zc * g.Ny);
  const int z0 = zc * ZC, nz = min(ZC, g.Nz - z0);
  const int n = 3 * g.M * nz;
  const double *src = aos + 3 * ((((long long)x * g.Ny + y) * g.Nz + z0) * g.M);
  for (int i = threadIdx.x; i < n; i += blockDim.x) tile[i] = src[i];
  __syncthreads();
  double *const d3[3] = {dx, dy, dz};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i / (g.M * nz), r = i - c * (g.M * nz);
    const int m = r / nz, zz = r - m * nz;
    d3[c][gidx(g, x + g.gx, y + g.gy, m, z0 + zz + g.oz)] = tile[(zz * g.M + m) * 3 + c];
  }
}

__global__ void __launch_bounds__(256) export_rows_kernel(const JbGeom g, const double *__restrict__ sx, const double *__restrict__ sy,
                                                          const double *__restrict__ sz, double *__restrict__ aos, int x_begin, int n_zc, int ZC) {
  extern __shared__ double tile[];   // [zz][m][c]
  const int zc = blockIdx.x % n_zc;
  const int y = (blockIdx.x / n_zc) % g.Ny;
  const int x = x_begin + blockIdx.x / (n_zc * g.Ny);
  const int z0 = zc * ZC, nz = min(ZC, g.Nz - z0);
  const int n = 3 * g.M * nz;
  const double *const s3[3] = {sx, sy, sz};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i / (g.M * nz), r = i - c * (g.M * nz);
    const int m = r / nz, zz = r - m * nz;
    tile[(zz * g.M + m) * 3 + c] = s3[c][gidx(g, x + g.gx, y + g.gy, m, z0 + zz + g.oz)];
  }
  __syncthreads();
  double *dst = aos + 3 * ((((long long)x * g.Ny + y) * g.Nz + z0) * g.M);
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = tile[i];
}

// ghost cells of a freshly imported box: every ghost cell takes the value of the interior cell it is the periodic image of (zero
// across an open boundary: Lattice::apply_boundary_conditions, core/lattice.cc:987-1007); padding columns are zeroed.  One warp
// per (xp, yp, m) row.  x ghost planes only with fill_x (on several ranks they belong to the neighbours, who push into them).
__global__ void __launch_bounds__(256) fill_ghosts_kernel(const JbGeom g, double *__restrict__ dx, double *__restrict__ dy,
                                                          double *__restrict__ dz, int fill_x) {
  const long long n_rows = (long long)g.PX * g.PY * g.M;
  const int lane = threadIdx.x & 31;
  for (long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; row < n_rows; row += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int m = (int)(row % g.M);
    const int yp = (int)((row / g.M) % g.PY);
    const int xp = (int)(row / ((long long)g.M * g.PY));
    int x = xp - g.gx, y = yp - g.gy;
    const bool xg = x < 0 || x >= g.nx, yg = y < 0 || y >= g.Ny;
    if (xg && !fill_x) continue;
    bool ok = true;
    if (xg) { if (g.per[0]) x = (x + g.nx) % g.nx; else ok = false; }
    if (yg) { if (g.per[1]) y = (y + g.Ny) % g.Ny; else ok = false; }
    const long long dst0 = gidx(g, xp, yp, m, 0);
    const long long src0 = ok ? gidx(g, x + g.gx, y + g.gy, m, g.oz) : 0;   // z = 0 of the source row
    const bool ghost_row = xg || yg;
    for (int zp = lane; zp < g.PZ; zp += 32) {
      int z = zp - g.oz;
      const bool zin = z >= 0 && z < g.Nz;
      if (zin && !ghost_row) continue;   // interior cell: imported
      bool okz = ok;
      if (!zin) {
        if (g.per[2] && z >= -g.gz && z < g.Nz + g.gz) z = (z + g.Nz) % g.Nz; else okz = false;
      }
      double vx = 0.0, vy = 0.0, vz = 0.0;
      if (okz) { vx = dx[src0 + z]; vy = dy[src0 + z]; vz = dz[src0 + z]; }
      dx[dst0 + zp] = vx; dy[dst0 + zp] = vy; dz[dst0 + zp] = vz;
    }
  }
}

// copy my gx lowest / highest interior planes (complete planes incl. their y/z ghosts) into the
// neighbours' x ghost planes
__global__ void push_x_ghosts_kernel(const JbGeom g, const double *__restrict__ s0, const double *__restrict__ s1,
                                     const double *__restrict__ s2, double *lo0, double *lo1, double *lo2,
                                     double *hi0, double *hi1, double *hi2) {
  const long long plane = g.sX;
  const long long total = plane * g.gx;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long k = t / plane, r = t % plane;
    if (lo0) {  // my planes xp = gx + k  ->  lo neighbour's xp = gx + nx + k
      const long long src = (g.gx + k) * plane + r, dst = (g.gx + g.nx + k) * plane + r;
      lo0[dst] = s0[src]; lo1[dst] = s1[src]; lo2[dst] = s2[src];
    }
    if (hi0) {  // my planes xp = nx + k  ->  hi neighbour's xp = k
      const long long src = (g.nx + k) * plane + r, dst = k * plane + r;
      hi0[dst] = s0[src]; hi1[dst] = s1[src]; hi2[dst] = s2[src];
    }
  }
}

// biquadratic exchange, field of one site in meV added to (hx, hy, hz): sum_j 2 B_ij s_j (s_i . s_j), accumulated in the
// reference's order ((2 B) s_j[n]) (s_i . s_j) over ascending neighbour ids (cuda_biquadratic_exchange_kernel.cuh:14-24)
__device__ __forceinline__ void biquadratic_field_site(const JbGeom &g, const JbTables &t, const double *__restrict__ inx,
                                                       const double *__restrict__ iny, const double *__restrict__ inz, long long ic, int m,
                                                       double sx, double sy, double sz, double &hx, double &hy, double &hz) {
  if (!t.bq_global) return;
  double bx = 0.0, by = 0.0, bz = 0.0;
  for (int n = t.bq_begin[m]; n < t.bq_begin[m + 1]; ++n) {
    const JbNbr e = t.bq_global[n];
    const long long j = ic + (long long)e.dx * g.sX + e.delta;
    const double jx = inx[j], jy = iny[j], jz = inz[j];
    const double d = sx * jx + sy * jy + sz * jz;
    const double b2 = 2.0 * e.J;
    bx += b2 * jx * d; by += b2 * jy * d; bz += b2 * jz * d;
  }
  hx += bx; hy += by; hz += bz;
}

// =================================================================================================
// stage kernel, variant 0: direct gathers from the ghosted box through L1/L2 (one thread per spin)
// =================================================================================================
template <int STAGE, bool THERMAL, bool ISO>
__global__ void __launch_bounds__(256) stage_direct_kernel(const __grid_constant__ JbStageParams p) {
  const JbGeom &g = p.g;
  const long long total = (long long)g.nx * g.Ny * g.Nz * g.M;
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= total) return;
  int x, y, m, z;
  decode_site(g, q, x, y, m, z);
  const long long ic = gidx(g, x + g.gx, y + g.gy, m, z + g.oz);
  const double *__restrict__ inx = p.in[0];
  const double *__restrict__ iny = p.in[1];
  const double *__restrict__ inz = p.in[2];
  const double sx = inx[ic], sy = iny[ic], sz = inz[ic];

  double hx = 0.0, hy = 0.0, hz = 0.0;
  const int nb = p.t.nbr_begin[m], ne = p.t.nbr_begin[m + 1];
  for (int n = nb; n < ne; ++n) {
    const JbNbr e = p.t.nbr_global[n];
    const long long j = ic + (long long)e.dx * g.sX + e.delta;
    const double jx = inx[j], jy = iny[j], jz = inz[j];
    if (ISO) {
      hx = fma(e.J, jx, hx); hy = fma(e.J, jy, hy); hz = fma(e.J, jz, hz);
    } else {
      const double *__restrict__ J = p.t.Jtab + 9 * e.jidx;
      hx += J[0] * jx + J[1] * jy + J[2] * jz;
      hy += J[3] * jx + J[4] * jy + J[5] * jz;
      hz += J[6] * jx + J[7] * jy + J[8] * jz;
    }
  }
  biquadratic_field_site(g, p.t, inx, iny, inz, ic, m, sx, sy, sz, hx, hy, hz);
  const int ci = p.t.site_class ? (int)p.t.site_class[q] : p.t.class_of_motif[m];
  const JbClass &c = p.t.classes[ci];
  double n0 = 0, n1 = 0, n2 = 0;
  if (THERMAL) site_normals_at(g, p.seed, p.step, x, y, m, z, n0, n1, n2);
  double ux = 0, uy = 0, uz = 0;
  if (STAGE == 1) { ux = p.u[0][ic]; uy = p.u[1][ic]; uz = p.u[2][ic]; }
  double ox, oy, oz, vx, vy, vz;
  hx = fma(hx, c.inv_mu, c.fTx); hy = fma(hy, c.inv_mu, c.fTy); hz = fma(hz, c.inv_mu, c.fTz);   // Tesla
  uniaxial_extra_site(p.t, ci, 3, true, sx, sy, sz, hx, hy, hz);
  llg_site<STAGE, THERMAL>(c, sx, sy, sz, hx, hy, hz, n0, n1, n2, ux, uy, uz, ox, oy, oz, vx, vy, vz);
  if (STAGE == 0) { p.u[0][ic] = vx; p.u[1][ic] = vy; p.u[2][ic] = vz; }
  store_with_images(p, x, y, m, z, ox, oy, oz);
}


// =================================================================================================
// RK4-LLG (llg-rk4-gpu: solvers/cuda_rk4_base.cu:50-108, cuda_llg_rk4_kernel.cuh:11-58), one launch per stage.
// The reference runs per stage: field kernels + field sum, cuda_llg_rk4_kernel (k_i), cublasDcopy + cublasDaxpy
// (next stage input) and keeps k1..k4 as four N x 3 arrays; here a stage launch evaluates the fields, k_i, the next
// stage input y = s_old + a_i dt k_i (not normalised, as in the reference) and the running sum k1 + 2 k2 + 2 k3 in one
// pass; the last stage applies cuda_rk4_combination_kernel (cuda_rk4_base_kernel.cuh:16) and the normalisation
// (cuda/cuda_spin_ops.cu:4-17, here with the zero-length guard of Vec3 unit_vector).
//   HBM bytes per spin: 72 + 120 + 120 + 96 = 408 per step (reference: > 2.5 kB).
// =================================================================================================
template <int STAGE, bool THERMAL, bool ISO>
__global__ void __launch_bounds__(256) rk4_direct_kernel(const __grid_constant__ JbStageParams p) {
  const JbGeom &g = p.g;
  const long long total = (long long)g.nx * g.Ny * g.Nz * g.M;
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= total) return;
  int x, y, m, z;
  decode_site(g, q, x, y, m, z);
  const long long ic = gidx(g, x + g.gx, y + g.gy, m, z + g.oz);
  const double *__restrict__ inx = p.in[0];
  const double *__restrict__ iny = p.in[1];
  const double *__restrict__ inz = p.in[2];
  const double sx = inx[ic], sy = iny[ic], sz = inz[ic];

  double hx = 0.0, hy = 0.0, hz = 0.0;
  const int nb = p.t.nbr_begin[m], ne = p.t.nbr_begin[m + 1];
  for (int n = nb; n < ne; ++n) {
    const JbNbr e = p.t.nbr_global[n];
    const long long j = ic + (long long)e.dx * g.sX + e.delta;
    const double jx = inx[j], jy = iny[j], jz = inz[j];
    if (ISO) {
      hx = fma(e.J, jx, hx); hy = fma(e.J, jy, hy); hz = fma(e.J, jz, hz);
    } else {
      const double *__restrict__ J = p.t.Jtab + 9 * e.jidx;
      hx += J[0] * jx + J[1] * jy + J[2] * jz;
      hy += J[3] * jx + J[4] * jy + J[5] * jz;
      hz += J[6] * jx + J[7] * jy + J[8] * jz;
    }
  }
  biquadratic_field_site(g, p.t, inx, iny, inz, ic, m, sx, sy, sz, hx, hy, hz);
  const int ci = p.t.site_class ? (int)p.t.site_class[q] : p.t.class_of_motif[m];
  const JbClass &c = p.t.classes[ci];
  hx = fma(hx, c.inv_mu, c.fTx); hy = fma(hy, c.inv_mu, c.fTy); hz = fma(hz, c.inv_mu, c.fTz);   // Tesla
  if (c.power != 0) {  // uniaxial (uniaxial_anisotropy.cc:155-163), here / mu
    const double d = c.ax * sx + c.ay * sy + c.az * sz;
    double pw = d;
    if (c.power >= 4) pw = d * d * d;
    if (c.power >= 6) pw = pw * d * d;
    const double f = c.KpT * pw;
    hx = fma(f, c.ax, hx); hy = fma(f, c.ay, hy); hz = fma(f, c.az, hz);
  }
  uniaxial_extra_site(p.t, ci, 3, true, sx, sy, sz, hx, hy, hz);
  if (THERMAL) {   // one draw per step, all four stages (cuda_rk4_base.cu:65)
    double n0, n1, n2;
    site_normals_at(g, p.seed, p.step, x, y, m, z, n0, n1, n2);
    hx = fma(c.sigma, n0, hx); hy = fma(c.sigma, n1, hy); hz = fma(c.sigma, n2, hz);
  }
  // cuda_llg_rk4_kernel.cuh:36-56
  const double ax_ = sy * hz - sz * hy, ay_ = sz * hx - sx * hz, az_ = sx * hy - sy * hx;
  const double bx_ = sy * az_ - sz * ay_, by_ = sz * ax_ - sx * az_, bz_ = sx * ay_ - sy * ax_;
  const double mg = -c.gyro;
  const double kx = mg * (ax_ + c.alpha * bx_), ky = mg * (ay_ + c.alpha * by_), kz = mg * (az_ + c.alpha * bz_);

  double ox, oy, oz;
  if (STAGE == 0) {          // y1 = s_old + dt/2 k1 ; sum = k1
    const double a = 0.5 * p.dt;
    ox = sx + a * kx; oy = sy + a * ky; oz = sz + a * kz;
    p.u[0][ic] = kx; p.u[1][ic] = ky; p.u[2][ic] = kz;
  } else {
    const double s0x = p.s_old[0][ic], s0y = p.s_old[1][ic], s0z = p.s_old[2][ic];
    const double ux = p.u[0][ic], uy = p.u[1][ic], uz = p.u[2][ic];
    if (STAGE == 1 || STAGE == 2) {   // y = s_old + a dt k ; sum += 2 k
      const double a = (STAGE == 1) ? 0.5 * p.dt : p.dt;
      ox = s0x + a * kx; oy = s0y + a * ky; oz = s0z + a * kz;
      p.u[0][ic] = ux + 2 * kx; p.u[1][ic] = uy + 2 * ky; p.u[2][ic] = uz + 2 * kz;
    } else {                          // s = unit(s_old + dt (k1 + 2 k2 + 2 k3 + k4) / 6)
      const double vx = s0x + p.dt * (ux + kx) / 6.0, vy = s0y + p.dt * (uy + ky) / 6.0, vz = s0z + p.dt * (uz + kz) / 6.0;
      const double n2_ = vx * vx + vy * vy + vz * vz;
      const double r = rsqrt_nobranch(n2_);
      const double inv = (n2_ > 4.930380657631324e-32) ? r : 1.0;
      ox = vx * inv; oy = vy * inv; oz = vz * inv;
    }
  }
  store_with_images(p, x, y, m, z, ox, oy, oz);
}

// =================================================================================================
// stage kernel, variant 2: general neighbour list (ELL, explicit int32 indices)
// =================================================================================================
template <int STAGE, bool THERMAL, bool ISO>
__global__ void __launch_bounds__(256) stage_pairs_kernel(const __grid_constant__ JbStageParams p,
                                                          const int *__restrict__ ell_idx, const int *__restrict__ ell_val,
                                                          int width, const double *__restrict__ pairJ) {
  const JbGeom &g = p.g;
  const long long total = (long long)g.nx * g.Ny * g.Nz * g.M;
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= total) return;
  int x, y, m, z;
  decode_site(g, q, x, y, m, z);
  const long long ic = gidx(g, x + g.gx, y + g.gy, m, z + g.oz);
  const double *__restrict__ inx = p.in[0];
  const double *__restrict__ iny = p.in[1];
  const double *__restrict__ inz = p.in[2];
  const double sx = inx[ic], sy = iny[ic], sz = inz[ic];
  double hx = 0.0, hy = 0.0, hz = 0.0;
  for (int e = 0; e < width; ++e) {
    const int j = ell_idx[(long long)e * total + q];
    if (j < 0) continue;
    const double *__restrict__ J = pairJ + 9 * ell_val[(long long)e * total + q];
    const double jx = inx[j], jy = iny[j], jz = inz[j];
    if (ISO) {
      hx = fma(J[0], jx, hx); hy = fma(J[0], jy, hy); hz = fma(J[0], jz, hz);
    } else {
      hx += J[0] * jx + J[1] * jy + J[2] * jz;
      hy += J[3] * jx + J[4] * jy + J[5] * jz;
      hz += J[6] * jx + J[7] * jy + J[8] * jz;
    }
  }
  biquadratic_field_site(g, p.t, inx, iny, inz, ic, m, sx, sy, sz, hx, hy, hz);
  const int ci = p.t.site_class ? (int)p.t.site_class[q] : p.t.class_of_motif[m];
  const JbClass &c = p.t.classes[ci];
  double n0 = 0, n1 = 0, n2 = 0;
  if (THERMAL) site_normals_at(g, p.seed, p.step, x, y, m, z, n0, n1, n2);
  double ux = 0, uy = 0, uz = 0;
  if (STAGE == 1) { ux = p.u[0][ic]; uy = p.u[1][ic]; uz = p.u[2][ic]; }
  double ox, oy, oz, vx, vy, vz;
  hx = fma(hx, c.inv_mu, c.fTx); hy = fma(hy, c.inv_mu, c.fTy); hz = fma(hz, c.inv_mu, c.fTz);   // Tesla
  uniaxial_extra_site(p.t, ci, 3, true, sx, sy, sz, hx, hy, hz);
  llg_site<STAGE, THERMAL>(c, sx, sy, sz, hx, hy, hz, n0, n1, n2, ux, uy, uz, ox, oy, oz, vx, vy, vz);
  if (STAGE == 0) { p.u[0][ic] = vx; p.u[1][ic] = vy; p.u[2][ic] = vz; }
  // one rank: no ghost cells at all (gx = gy = gz = 0: every neighbour is addressed directly); several ranks: the x images of the
  // face planes go into the neighbours' boxes
  store_with_images(p, x, y, m, z, ox, oy, oz);
}

// =================================================================================================
// Hamiltonian / Monitor surface: fields, energies, magnetisation, noise
// =================================================================================================
__device__ __forceinline__ void exchange_field_site(const JbGeom &g, const JbTables &t, const double *__restrict__ inx,
                                                    const double *__restrict__ iny, const double *__restrict__ inz,
                                                    long long q, long long ic, int m, const int *__restrict__ ell_idx,
                                                    const int *__restrict__ ell_val, int width,
                                                    const double *__restrict__ pairJ, int pairs_iso, long long total,
                                                    double &hx, double &hy, double &hz) {
  hx = hy = hz = 0.0;
  if (ell_idx) {
    for (int e = 0; e < width; ++e) {
      const int j = ell_idx[(long long)e * total + q];
      if (j < 0) continue;
      const double *__restrict__ J = pairJ + 9 * ell_val[(long long)e * total + q];
      const double jx = inx[j], jy = iny[j], jz = inz[j];
      if (pairs_iso) {
        hx = fma(J[0], jx, hx); hy = fma(J[0], jy, hy); hz = fma(J[0], jz, hz);
      } else {
        hx += J[0] * jx + J[1] * jy + J[2] * jz;
        hy += J[3] * jx + J[4] * jy + J[5] * jz;
        hz += J[6] * jx + J[7] * jy + J[8] * jz;
      }
    }
    return;
  }
  if (!t.nbr_global) return;
  for (int n = t.nbr_begin[m]; n < t.nbr_begin[m + 1]; ++n) {
    const JbNbr e = t.nbr_global[n];
    const long long j = ic + (long long)e.dx * g.sX + e.delta;
    const double jx = inx[j], jy = iny[j], jz = inz[j];
    if (t.iso) {
      hx = fma(e.J, jx, hx); hy = fma(e.J, jy, hy); hz = fma(e.J, jz, hz);
    } else {
      const double *__restrict__ J = t.Jtab + 9 * e.jidx;
      hx += J[0] * jx + J[1] * jy + J[2] * jz;
      hy += J[3] * jx + J[4] * jy + J[5] * jz;
      hz += J[6] * jx + J[7] * jy + J[8] * jz;
    }
  }
}

__device__ __forceinline__ double ipow_even(double d, int power) {  // d^power for power in {2,4,6}
  double d2 = d * d, r = d2;
  if (power >= 4) r *= d2;
  if (power >= 6) r *= d2;
  return r;
}

__global__ void field_kernel(const JbGeom g, const JbTables t, const double *__restrict__ inx,
                             const double *__restrict__ iny, const double *__restrict__ inz, int term,
                             const int *__restrict__ ell_idx, const int *__restrict__ ell_val, int width,
                             const double *__restrict__ pairJ, int pairs_iso, double *__restrict__ h_aos) {
  const long long total = (long long)g.nx * g.Ny * g.Nz * g.M;
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= total) return;
  int x, y, m, z;
  decode_site(g, q, x, y, m, z);
  const long long ic = gidx(g, x + g.gx, y + g.gy, m, z + g.oz);
  double hx = 0, hy = 0, hz = 0;
  if (term == JB_TERM_EXCHANGE || term == JB_TERM_TOTAL)
    exchange_field_site(g, t, inx, iny, inz, q, ic, m, ell_idx, ell_val, width, pairJ, pairs_iso, total, hx, hy, hz);
  if (term == JB_TERM_BIQUADRATIC || term == JB_TERM_TOTAL)
    biquadratic_field_site(g, t, inx, iny, inz, ic, m, inx[ic], iny[ic], inz[ic], hx, hy, hz);
  const int ci = t.site_class ? (int)t.site_class[q] : t.class_of_motif[m];
  const JbClass c = t.classes[ci];
  if ((term == JB_TERM_UNIAXIAL || term == JB_TERM_TOTAL) && c.power != 0) {
    const double sx = inx[ic], sy = iny[ic], sz = inz[ic];
    const double d = c.ax * sx + c.ay * sy + c.az * sz;
    double pw = d;
    if (c.power >= 4) pw = d * d * d;
    if (c.power >= 6) pw = pw * d * d;
    const double f = c.Kp * pw;
    hx = fma(f, c.ax, hx); hy = fma(f, c.ay, hy); hz = fma(f, c.az, hz);
  }
  if (term == JB_TERM_UNIAXIAL_2 || term == JB_TERM_UNIAXIAL_3 || term == JB_TERM_TOTAL)
    uniaxial_extra_site(t, ci, term == JB_TERM_TOTAL ? 3 : (term == JB_TERM_UNIAXIAL_2 ? 1 : 2), false, inx[ic], iny[ic], inz[ic], hx, hy, hz);
  if (term == JB_TERM_ZEEMAN || term == JB_TERM_APPLIED || term == JB_TERM_TOTAL) { hx += c.fx; hy += c.fy; hz += c.fz; }
  const long long s = ref_site_local(g, x, y, m, z);
  h_aos[3 * s] = hx; h_aos[3 * s + 1] = hy; h_aos[3 * s + 2] = hz;
}

// block-level sum with warp shuffles; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double *sm) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  if (lane == 0) sm[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

// per-spin energy of one term; per-block partial sums to partial[blockIdx.x]
__global__ void __launch_bounds__(256) energy_kernel(const JbGeom g, const JbTables t, const double *__restrict__ inx,
                                                    const double *__restrict__ iny, const double *__restrict__ inz, int term,
                                                    const int *__restrict__ ell_idx, const int *__restrict__ ell_val, int width,
                                                    const double *__restrict__ pairJ, int pairs_iso,
                                                    double *__restrict__ e_out, double *__restrict__ partial) {
  __shared__ double sm[32];
  const long long total = (long long)g.nx * g.Ny * g.Nz * g.M;
  double acc = 0.0;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    int x, y, m, z;
    decode_site(g, q, x, y, m, z);
    const long long ic = gidx(g, x + g.gx, y + g.gy, m, z + g.oz);
    const double sx = inx[ic], sy = iny[ic], sz = inz[ic];
    const int ci = t.site_class ? (int)t.site_class[q] : t.class_of_motif[m];
    const JbClass c = t.classes[ci];
    double e = 0.0;
    if (term == JB_TERM_EXCHANGE) {
      double hx, hy, hz;
      exchange_field_site(g, t, inx, iny, inz, q, ic, m, ell_idx, ell_val, width, pairJ, pairs_iso, total, hx, hy, hz);
      e = -(sx * hx + sy * hy + sz * hz);  // sparse_interaction.cc:79-84
    } else if (term == JB_TERM_BIQUADRATIC) {
      double hx = 0.0, hy = 0.0, hz = 0.0;
      biquadratic_field_site(g, t, inx, iny, inz, ic, m, sx, sy, sz, hx, hy, hz);
      e = -0.5 * (sx * hx + sy * hy + sz * hz);  // cuda_biquadratic_exchange.cu:235-240
    } else if (term == JB_TERM_UNIAXIAL) {
      if (c.power != 0) {
        const double d = c.ax * sx + c.ay * sy + c.az * sz;
        e = -c.K * ipow_even(d, c.power);  // uniaxial_anisotropy.cc:126-133
      }
    } else if (term == JB_TERM_UNIAXIAL_2 || term == JB_TERM_UNIAXIAL_3) {
      if (t.uni_extra) {
        const JbUniExtra u = t.uni_extra[ci * (JB_MAX_UNIAXIAL - 1) + (term == JB_TERM_UNIAXIAL_2 ? 0 : 1)];
        if (u.power != 0) e = -u.K * ipow_even(u.ax * sx + u.ay * sy + u.az * sz, u.power);
      }
    } else {  // ZEEMAN / APPLIED: -s . f   (zeeman.cc:82-87, applied_field.cc:150+)
      e = -(sx * c.fx + sy * c.fy + sz * c.fz);
    }
    if (e_out) e_out[ref_site_local(g, x, y, m, z)] = e;
    acc += e;
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// deterministic final reduction of n partials (one block), scaled
__global__ void __launch_bounds__(256) final_sum_kernel(const double *__restrict__ partial, int n, double scale, double *__restrict__ out) {
  __shared__ double sm[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) out[0] = acc * scale;
}

// sum_i mu_i s_i (x,y,z) and sum_i mu_i over the spins of one group; partial[(blockIdx.x*4 + c)]
__global__ void __launch_bounds__(256) magnetisation_kernel(const JbGeom g, const JbTables t, const double *__restrict__ inx,
                                                           const double *__restrict__ iny, const double *__restrict__ inz,
                                                           const int *__restrict__ group_of_spin, int group,
                                                           double *__restrict__ partial) {
  __shared__ double sm[32];
  const long long total = (long long)g.nx * g.Ny * g.Nz * g.M;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    int x, y, m, z;
    decode_site(g, q, x, y, m, z);
    if (group_of_spin && group_of_spin[ref_site_local(g, x, y, m, z)] != group) continue;
    const long long ic = gidx(g, x + g.gx, y + g.gy, m, z + g.oz);
    const int ci = t.site_class ? (int)t.site_class[q] : t.class_of_motif[m];
    const double mu = t.classes[ci].mu;
    a0 = fma(mu, inx[ic], a0); a1 = fma(mu, iny[ic], a1); a2 = fma(mu, inz[ic], a2); a3 += mu;
  }
  a0 = block_sum(a0, sm); a1 = block_sum(a1, sm); a2 = block_sum(a2, sm); a3 = block_sum(a3, sm);
  if (threadIdx.x == 0) {
    partial[4 * blockIdx.x] = a0; partial[4 * blockIdx.x + 1] = a1; partial[4 * blockIdx.x + 2] = a2; partial[4 * blockIdx.x + 3] = a3;
  }
}

__global__ void __launch_bounds__(256) final_sum4_kernel(const double *__restrict__ partial, int n, double *__restrict__ out4) {
  __shared__ double sm[32];
  for (int c = 0; c < 4; ++c) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[4 * i + c];
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) out4[c] = acc;
  }
}


// =================================================================================================
// region operations for PinnedBoundariesPhysics::update (physics/pinned_boundaries.cc:34-46): the moment of a set of
// spins, sum mu_i s_i (jams::vector_field_indexed_scale_and_reduce_cuda, cuda/cuda_array_reduction.cu:432-470), and the
// rotation of that set, s_i <- R s_i (cuda_rotate_spins_kernel, cuda/cuda_spin_ops.cu:29-43).  `sites` holds local site
// ids in the reference order ((x Ny + y) Nz + z) M + m.
// =================================================================================================
__device__ __forceinline__ void decode_ref_site(const JbGeom &g, int site, int &x, int &y, int &m, int &z) {
  unsigned q = (unsigned)site;
  const unsigned um = (unsigned)g.M, uz = (unsigned)g.Nz, uy = (unsigned)g.Ny;
  unsigned t = q / um; m = (int)(q - t * um); q = t;
  t = q / uz; z = (int)(q - t * uz); q = t;
  t = q / uy; y = (int)(q - t * uy);
  x = (int)t;
}

__global__ void __launch_bounds__(256) region_moment_kernel(const JbGeom g, const JbTables t, const double *__restrict__ inx,
                                                           const double *__restrict__ iny, const double *__restrict__ inz,
                                                           const int *__restrict__ sites, int n, double *__restrict__ partial) {
  __shared__ double sm[32];
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    int x, y, m, z;
    decode_ref_site(g, sites[k], x, y, m, z);
    const long long ic = gidx(g, x + g.gx, y + g.gy, m, z + g.oz);
    const long long q = (((long long)x * g.Ny + y) * g.M + m) * g.Nz + z;   // interior layout order [x][y][m][z]
    const int ci = t.site_class ? (int)t.site_class[q] : t.class_of_motif[m];
    const double mu = t.classes[ci].mu;
    a0 = fma(mu, inx[ic], a0); a1 = fma(mu, iny[ic], a1); a2 = fma(mu, inz[ic], a2); a3 += mu;
  }
  a0 = block_sum(a0, sm); a1 = block_sum(a1, sm); a2 = block_sum(a2, sm); a3 = block_sum(a3, sm);
  if (threadIdx.x == 0) {
    partial[4 * blockIdx.x] = a0; partial[4 * blockIdx.x + 1] = a1; partial[4 * blockIdx.x + 2] = a2; partial[4 * blockIdx.x + 3] = a3;
  }
}

struct JbRot { double r[9]; };

__global__ void __launch_bounds__(128) region_rotate_kernel(const JbGeom g, const JbOutBoxes o, const int *__restrict__ sites, int n, const JbRot R) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int x, y, m, z;
  decode_ref_site(g, sites[k], x, y, m, z);
  const long long ic = gidx(g, x + g.gx, y + g.gy, m, z + g.oz);
  const double s0 = o.out[0][ic], s1 = o.out[1][ic], s2 = o.out[2][ic];
  const double vx = R.r[0] * s0 + R.r[1] * s1 + R.r[2] * s2;   // cuda_spin_ops.cu:38-40
  const double vy = R.r[3] * s0 + R.r[4] * s1 + R.r[5] * s2;
  const double vz = R.r[6] * s0 + R.r[7] * s1 + R.r[8] * s2;
  o.out[0][ic] = vx; o.out[1][ic] = vy; o.out[2][ic] = vz;
  if (x_image_needed(g, x) | yz_image_needed(g, y, z)) store_images(g, o, x, y, m, z, vx, vy, vz);
}

__global__ void noise_kernel(const JbGeom g, const JbTables t, unsigned long long seed, unsigned long long step,
                             int normals_only, double *__restrict__ xi) {
  const long long total = (long long)g.nx * g.Ny * g.Nz * g.M;
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= total) return;
  int x, y, m, z;
  decode_site(g, q, x, y, m, z);
  double n0, n1, n2;
  site_normals_at(g, seed, step, x, y, m, z, n0, n1, n2);
  double sc = 1.0;
  if (!normals_only) {
    const int ci = t.site_class ? (int)t.site_class[q] : t.class_of_motif[m];
    sc = t.classes[ci].sigma;
  }
  const long long s = ref_site_local(g, x, y, m, z);
  xi[3 * s] = sc * n0; xi[3 * s + 1] = sc * n1; xi[3 * s + 2] = sc * n2;
}

// ---- halo signalling: epoch flags in (peer) device memory -----------------------------------------
__global__ void signal_kernel(unsigned long long *lo, unsigned long long *hi, unsigned long long epoch) {
  __threadfence_system();
  if (lo) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(lo), "l"(epoch) : "memory"); }
  if (hi) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(hi), "l"(epoch) : "memory"); }
}

__global__ void wait_kernel(unsigned long long *flags, int wait_lo, int wait_hi, unsigned long long epoch) {
  const long long t0 = clock64();
  const long long limit = 20000000000LL;  // ~10 s at 2 GHz: a dead peer must not hang the GPU
  for (int w = 0; w < 2; ++w) {
    if (!(w == 0 ? wait_lo : wait_hi)) continue;
    unsigned long long v = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + w) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > limit) { flags[2] = 1ull; return; }
      __nanosleep(200);
    }
  }
}

template <typename K, typename... Args>
cudaError_t launch1d(K kernel, long long total, int threads, cudaStream_t stream, Args... args) {
  long long blocks = (total + threads - 1) / threads;
  if (blocks < 1) blocks = 1;
  kernel<<<(unsigned)blocks, threads, 0, stream>>>(args...);
  return cudaGetLastError();
}

}  // namespace

// =================================================================================================
// launchers
// =================================================================================================
static int rows_zc(const JbGeom &g) { return std::max(1, std::min(g.Nz, 2048 / std::max(1, g.M))); }   // 3 M ZC doubles <= 48 KB of shared memory

cudaError_t jbk_import(const JbGeom &g, const double *aos, double *const dst[3], bool fill_x_ghosts, cudaStream_t stream) {
  cudaError_t e = jbk_import_planes(g, aos, dst, 0, g.nx, stream);
  if (e != cudaSuccess) return e;
  return jbk_fill_ghosts(g, dst, fill_x_ghosts, stream);
}

cudaError_t jbk_import_planes(const JbGeom &g, const double *aos, double *const dst[3], int x_begin, int x_end, cudaStream_t stream) {
  if (x_end <= x_begin) return cudaSuccess;
  const int ZC = rows_zc(g), n_zc = (g.Nz + ZC - 1) / ZC;
  const long long blocks = (long long)(x_end - x_begin) * g.Ny * n_zc;
  import_rows_kernel<<<(unsigned)blocks, 256, (size_t)3 * g.M * ZC * sizeof(double), stream>>>(g, aos, dst[0], dst[1], dst[2], x_begin, n_zc, ZC);
  return cudaGetLastError();
}

cudaError_t jbk_fill_ghosts(const JbGeom &g, double *const dst[3], bool fill_x_ghosts, cudaStream_t stream) {
  const long long n_rows = (long long)g.PX * g.PY * g.M;
  long long blocks = (n_rows + 7) / 8;
  if (blocks > 148 * 32) blocks = 148 * 32;
  fill_ghosts_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g, dst[0], dst[1], dst[2], fill_x_ghosts ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t jbk_export(const JbGeom &g, const double *const src[3], double *aos, cudaStream_t stream) {
  return jbk_export_planes(g, src, aos, 0, g.nx, stream);
}

cudaError_t jbk_export_planes(const JbGeom &g, const double *const src[3], double *aos, int x_begin, int x_end, cudaStream_t stream) {
  if (x_end <= x_begin) return cudaSuccess;
  const int ZC = rows_zc(g), n_zc = (g.Nz + ZC - 1) / ZC;
  const long long blocks = (long long)(x_end - x_begin) * g.Ny * n_zc;
  export_rows_kernel<<<(unsigned)blocks, 256, (size_t)3 * g.M * ZC * sizeof(double), stream>>>(g, src[0], src[1], src[2], aos, x_begin, n_zc, ZC);
  return cudaGetLastError();
}

cudaError_t jbk_push_x_ghosts(const JbGeom &g, const double *const src[3], double *const lo[3], double *const hi[3], cudaStream_t stream) {
  if (g.gx == 0) return cudaSuccess;
  const long long total = g.sX * g.gx;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  push_x_ghosts_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g, src[0], src[1], src[2], lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
  return cudaGetLastError();
}

#define JB_DISPATCH_STAGE(KERNEL, LAUNCH)                                            \
  do {                                                                               \
    const bool th = p.thermal != 0, iso = p.t.iso != 0;                              \
    if (stage == 0) {                                                                \
      if (th) { if (iso) { auto k = KERNEL<0, true, true>; LAUNCH; } else { auto k = KERNEL<0, true, false>; LAUNCH; } } \
      else    { if (iso) { auto k = KERNEL<0, false, true>; LAUNCH; } else { auto k = KERNEL<0, false, false>; LAUNCH; } } \
    } else {                                                                         \
      if (th) { if (iso) { auto k = KERNEL<1, true, true>; LAUNCH; } else { auto k = KERNEL<1, true, false>; LAUNCH; } } \
      else    { if (iso) { auto k = KERNEL<1, false, true>; LAUNCH; } else { auto k = KERNEL<1, false, false>; LAUNCH; } } \
    }                                                                                \
  } while (0)

cudaError_t jbk_stage_direct(const JbStageParams &p, int stage, cudaStream_t stream) {
  const long long total = (long long)p.g.nx * p.g.Ny * p.g.Nz * p.g.M;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  JB_DISPATCH_STAGE(stage_direct_kernel, (k<<<blocks, 256, 0, stream>>>(p)));
  return cudaGetLastError();
}


cudaError_t jbk_rk4_stage_direct(const JbStageParams &p, int stage, cudaStream_t stream) {
  const long long total = (long long)p.g.nx * p.g.Ny * p.g.Nz * p.g.M;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  const bool th = p.thermal != 0, iso = p.t.iso != 0;
#define JB_RK4_CASE(ST) \
  if (stage == ST) { \
    if (th) { if (iso) rk4_direct_kernel<ST, true, true><<<blocks, 256, 0, stream>>>(p); else rk4_direct_kernel<ST, true, false><<<blocks, 256, 0, stream>>>(p); } \
    else    { if (iso) rk4_direct_kernel<ST, false, true><<<blocks, 256, 0, stream>>>(p); else rk4_direct_kernel<ST, false, false><<<blocks, 256, 0, stream>>>(p); } \
  }
  JB_RK4_CASE(0) JB_RK4_CASE(1) JB_RK4_CASE(2) JB_RK4_CASE(3)
#undef JB_RK4_CASE
  return cudaGetLastError();
}

cudaError_t jbk_stage_pairs(const JbStageParams &p, const int *ell_idx, const int *ell_val, int width, const double *pairJ,
                            int iso_pairs, int stage, cudaStream_t stream) {
  const long long total = (long long)p.g.nx * p.g.Ny * p.g.Nz * p.g.M;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  JbStageParams q = p;
  q.t.iso = iso_pairs;
  {
    const JbStageParams &p = q;
    JB_DISPATCH_STAGE(stage_pairs_kernel, (k<<<blocks, 256, 0, stream>>>(p, ell_idx, ell_val, width, pairJ)));
  }
  return cudaGetLastError();
}

cudaError_t jbk_field(const JbGeom &g, const JbTables &t, const double *const s[3], int term, const int *ell_idx,
                      const int *ell_val, int width, const double *pairJ, int pairs_iso, double *h_aos, cudaStream_t stream) {
  const long long total = (long long)g.nx * g.Ny * g.Nz * g.M;
  return launch1d(field_kernel, total, 256, stream, g, t, s[0], s[1], s[2], term, ell_idx, ell_val, width, pairJ, pairs_iso, h_aos);
}

cudaError_t jbk_energy(const JbGeom &g, const JbTables &t, const double *const s[3], int term, const int *ell_idx,
                       const int *ell_val, int width, const double *pairJ, int pairs_iso, double *e_out, double *scratch,
                       double *total_out, cudaStream_t stream) {
  const long long total = (long long)g.nx * g.Ny * g.Nz * g.M;
  long long blocks = (total + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  energy_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g, t, s[0], s[1], s[2], term, ell_idx, ell_val, width, pairJ, pairs_iso, e_out, scratch);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  final_sum_kernel<<<1, 256, 0, stream>>>(scratch, (int)blocks, (term == JB_TERM_EXCHANGE || term == JB_TERM_BIQUADRATIC) ? 0.5 : 1.0, total_out);   // biquadratic: 1/2 sum_i -s_i . (h_i / 2), cuda_biquadratic_exchange.cu:185-201
  return cudaGetLastError();
}

cudaError_t jbk_magnetisation(const JbGeom &g, const JbTables &t, const double *const s[3], int n_groups,
                              const int *group_of_spin, double *scratch, double *out4, cudaStream_t stream) {
  const long long total = (long long)g.nx * g.Ny * g.Nz * g.M;
  long long blocks = (total + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  for (int grp = 0; grp < n_groups; ++grp) {
    magnetisation_kernel<<<(unsigned)blocks, 256, 0, stream>>>(g, t, s[0], s[1], s[2], group_of_spin, grp, scratch);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return err;
    final_sum4_kernel<<<1, 256, 0, stream>>>(scratch, (int)blocks, out4 + 4 * grp);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
  }
  return cudaSuccess;
}


cudaError_t jbk_region_moment(const JbGeom &g, const JbTables &t, const double *const s[3], const int *sites, int n,
                              double *scratch, double *out4, cudaStream_t stream) {
  int blocks = (n + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  region_moment_kernel<<<blocks, 256, 0, stream>>>(g, t, s[0], s[1], s[2], sites, n, scratch);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  final_sum4_kernel<<<1, 256, 0, stream>>>(scratch, blocks, out4);
  return cudaGetLastError();
}

cudaError_t jbk_region_rotate(const JbGeom &g, double *const s[3], double *const lo[3], double *const hi[3], const int *sites, int n,
                              const double R9[9], cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  JbOutBoxes o;
  for (int c = 0; c < 3; ++c) { o.out[c] = s[c]; o.out_lo[c] = lo[c]; o.out_hi[c] = hi[c]; }
  JbRot R;
  for (int k = 0; k < 9; ++k) R.r[k] = R9[k];
  region_rotate_kernel<<<(n + 127) / 128, 128, 0, stream>>>(g, o, sites, n, R);
  return cudaGetLastError();
}

cudaError_t jbk_noise(const JbGeom &g, const JbTables &t, unsigned long long seed, unsigned long long step,
                      int normals_only, double *xi_aos, cudaStream_t stream) {
  const long long total = (long long)g.nx * g.Ny * g.Nz * g.M;
  return launch1d(noise_kernel, total, 256, stream, g, t, seed, step, normals_only, xi_aos);
}

cudaError_t jbk_signal(unsigned long long *peer_lo_flag, unsigned long long *peer_hi_flag, unsigned long long epoch, cudaStream_t stream) {
  signal_kernel<<<1, 1, 0, stream>>>(peer_lo_flag, peer_hi_flag, epoch);
  return cudaGetLastError();
}

cudaError_t jbk_wait(unsigned long long *flags, int wait_lo, int wait_hi, unsigned long long epoch, cudaStream_t stream) {
  wait_kernel<<<1, 1, 0, stream>>>(flags, wait_lo, wait_hi, epoch);
  return cudaGetLastError();
}
