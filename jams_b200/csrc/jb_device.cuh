// jb_device.cuh — device-side building blocks shared by the stage kernels (jb_kernels.cu, jb_stage_pair.cu):
// ghosted-box indexing, the Philox4x32-10 Langevin noise, the per-spin LLG-Heun stage update and the
// ghost-image stores.  Reference formulas are cited next to each piece (paths relative to
// /root/reference/src/jams/).
#ifndef JB_DEVICE_CUH
#define JB_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "jb_internal.h"

namespace jbdev {

__device__ __forceinline__ long long gidx(const JbGeom &g, int xp, int yp, int m, int zp) {
  return (long long)xp * g.sX + (long long)yp * g.sY + (long long)m * g.PZ + zp;
}

// ---- Langevin white noise: Philox4x32-10 (Salmon et al., SC'11) in registers + Box-Muller on the SFU ----------
// One draw per step and site, used by both Heun stages (solvers/cuda_llg_heun.cu:79, cpu_llg_heun.cc:53-64), as a function of
// (seed, step, global site id) only -- any slab decomposition, tiling or kernel gives the same numbers.
//
// Stream definition (round 2): ONE Philox call per z-PAIR of sites.  counter = (global id of the pair's even-z site, step),
// key = seed; the 128 output bits w0..w3 feed three Box-Muller transforms k = 0, 1, 2:
//     radius  u_k = 2 - float(1.mantissa = w_k >> 9)   in (0, 1], 23 bits      r_k = sqrt(-2 ln u_k)
//     angle   b_0 = w3 & 0xffff, b_1 = w3 >> 16, b_2 = (w0 & 0xff) | (w1 & 0xff) << 8      theta_k = 2 pi (b_k + 1/2) / 65536
//     even-z site: ( r0 cos th0, r0 sin th0, r1 cos th1 )      odd-z site: ( r1 sin th1, r2 cos th2, r2 sin th2 )
// No output bit is used twice.  Six normals cost one Philox call (was: two calls for six normals and two discarded ones), which
// removes about 50 of the 145 noise instructions a pair of sites cost in round 1.
// Quality notes: |n| <= sqrt(2 * 23 ln 2) = 5.65 (the tail beyond has probability 1e-7 per draw and is lumped at the edge);
// the 65536 equidistant angles make every mixed moment E[n_a^p n_b^q] with p + q < 65536 exactly that of the continuous
// transform (the trapezoidal rule is exact for trigonometric polynomials below the number of nodes); lg2 / sqrt / sin / cos are
// the SFU approximations (absolute error ~1e-6), far below the statistical resolution of any observable -- the reference's two
// backends, pcg + std::normal_distribution on the CPU and cuRAND XORWOW on the GPU, already differ stream for stream.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// the same generator with the 20 round keys precomputed (rk[2r] = k0 + r W0, rk[2r+1] = k1 + r W1): when rk lives
// in the kernel-parameter bank the keys are constant-bank operands of the LOP3s and cost no instructions
__device__ __forceinline__ void philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 const unsigned int *__restrict__ rk, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
    const uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Box-Muller pieces in fp32 on the SFU (MUFU.LG2 / MUFU.SQRT / MUFU.SIN / MUFU.COS); results are exact fp32 values that the
// callers widen to double
__device__ __forceinline__ float bm_radius(uint32_t w) {   // sqrt(-2 ln u), u = 2 - 1.(w >> 9) in (0, 1]
  const float u = 2.0f - __uint_as_float(0x3f800000u | (w >> 9));
  float l2, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l2));   // ln u = ln 2 * lg2 u
  return r;
}
__device__ __forceinline__ float bm_angle(uint32_t b16) {   // 2 pi (b + 1/2) / 65536
  return __fmaf_rn((float)b16, 9.587379924285257e-05f, 4.7936899621426287e-05f);
}
struct PairNormals { float e0, e1, e2, o0, o1, o2; };   // the three draws of the even-z and of the odd-z site of a pair
__device__ __forceinline__ void pair_normals_from_words(const uint32_t w[4], PairNormals &n) {
  const float r0 = bm_radius(w[0]), r1 = bm_radius(w[1]), r2 = bm_radius(w[2]);
  const float t0 = bm_angle(w[3] & 0xffffu), t1 = bm_angle(w[3] >> 16), t2 = bm_angle(__byte_perm(w[0], w[1], 0x7740) & 0xffffu);
  n.e0 = r0 * __cosf(t0); n.e1 = r0 * __sinf(t0);
  n.e2 = r1 * __cosf(t1); n.o0 = r1 * __sinf(t1);
  n.o1 = r2 * __cosf(t2); n.o2 = r2 * __sinf(t2);
}
// all six draws of the pair whose even-z site has global id `gpair` (pair kernel: a thread owns both sites)
__device__ __forceinline__ void pair_normals_rk(const unsigned int *__restrict__ rk, unsigned long long step,
                                                unsigned long long gpair, PairNormals &n) {
  uint32_t w[4];
  philox4x32_10_rk((uint32_t)gpair, (uint32_t)(gpair >> 32), (uint32_t)step, (uint32_t)(step >> 32), rk, w);
  pair_normals_from_words(w, n);
}
// the three draws of ONE site (kernels with one site per thread, jb_noise): the same Philox call, only this site's half of
// the transforms.  gpair = global id of the pair's even-z site, odd = this is the pair's odd-z site
__device__ __forceinline__ void site_normals(unsigned long long seed, unsigned long long step, unsigned long long gpair, bool odd,
                                             double &n0, double &n1, double &n2) {
  uint32_t w[4];
  philox4x32_10((uint32_t)gpair, (uint32_t)(gpair >> 32), (uint32_t)step, (uint32_t)(step >> 32),
                (uint32_t)seed, (uint32_t)(seed >> 32), w);
  const float r1 = bm_radius(w[1]), t1 = bm_angle(w[3] >> 16);
  float a, b, c;
  if (!odd) {
    const float r0 = bm_radius(w[0]), t0 = bm_angle(w[3] & 0xffffu);
    a = r0 * __cosf(t0); b = r0 * __sinf(t0); c = r1 * __cosf(t1);
  } else {
    const float r2 = bm_radius(w[2]), t2 = bm_angle(__byte_perm(w[0], w[1], 0x7740) & 0xffffu);
    a = r1 * __sinf(t1); b = r2 * __cosf(t2); c = r2 * __sinf(t2);
  }
  n0 = (double)a; n1 = (double)b; n2 = (double)c;
}

// the same with the round keys precomputed (stage kernels: rk in the kernel-parameter bank)
__device__ __forceinline__ void site_normals_rk(const unsigned int *__restrict__ rk, unsigned long long step, unsigned long long gpair, bool odd,
                                                double &n0, double &n1, double &n2) {
  uint32_t w[4];
  philox4x32_10_rk((uint32_t)gpair, (uint32_t)(gpair >> 32), (uint32_t)step, (uint32_t)(step >> 32), rk, w);
  const float r1 = bm_radius(w[1]), t1 = bm_angle(w[3] >> 16);
  const uint32_t wr = odd ? w[2] : w[0];
  const uint32_t wa = odd ? (__byte_perm(w[0], w[1], 0x7740) & 0xffffu) : (w[3] & 0xffffu);
  const float r = bm_radius(wr), t = bm_angle(wa);
  const float c1 = __cosf(t1), s1 = __sinf(t1), cc = r * __cosf(t), ss = r * __sinf(t);
  // even-z site: (r0 cos t0, r0 sin t0, r1 cos t1); odd-z site: (r1 sin t1, r2 cos t2, r2 sin t2)
  n0 = (double)(odd ? r1 * s1 : cc); n1 = (double)(odd ? cc : ss); n2 = (double)(odd ? ss : r1 * c1);
}

__device__ __forceinline__ unsigned long long global_site(const JbGeom &g, int x, int y, int m, int z) {
  return (((unsigned long long)(g.x_begin + x) * g.Ny + y) * g.Nz + z) * g.M + m;
}
// the site's three draws through the z-pair it belongs to
__device__ __forceinline__ void site_normals_at(const JbGeom &g, unsigned long long seed, unsigned long long step, int x, int y, int m, int z,
                                                double &n0, double &n1, double &n2) {
  site_normals(seed, step, global_site(g, x, y, m, z & ~1), (z & 1) != 0, n0, n1, n2);
}

// ---- the per-spin physics ---------------------------------------------------------------------------
// Input: the spin s, the exchange + constant field h in TESLA (sum_j J_ij s_j / mu_i + f_i / mu_i), three
// N(0,1) draws, and in the corrector the Heun intermediate u.  Adds the uniaxial field and the noise,
// evaluates the LLG right hand side  rhs = -gyro ( s x h + alpha s x (s x h) )  (cpu_llg_heun.cc:89,130) and
// performs the stage update with the step folded into the class constants c_full = -gyro dt, c_half = -gyro dt/2:
//   STAGE 0 (predictor, :84-101): u = s + dt/2 rhs ; s* = unit(s + dt rhs)   [+ the noise part of rhs*, see corrector_noise_part]
//   STAGE 1 (corrector, :124-144): s' = unit(u + dt/2 rhs*)       [ = unit(s_old + dt (rhs/2 + rhs*/2)) ]
//   (stored-u data flow: stage 1 runs with THERMAL = false, the predictor has already put the noise part of rhs* into u;
//    recover_u data flow: stage 1 adds the noise itself)
// unit() keeps vectors of length <= DBL_EPSILON unchanged (containers/vec3.h:276-283): vacancies stay 0.
// The LLG right-hand side is linear in the field: rhs(s*, H* + xi) = rhs(s*, H*) + rhs(s*, xi), and rhs(s*, xi) needs only
// the site's own predictor spin s* and its own noise draw -- both at hand at the end of the predictor.  So the predictor
// folds dt/2 rhs(s*, xi) into the Heun intermediate it writes anyway (u' = s + dt/2 k1 + dt/2 rhs(s*, xi)) and the
// corrector, s' = unit(u' + dt/2 rhs(s*, H*)), needs no noise at all: the second Philox / Box-Muller evaluation per site
// and step (about a quarter of the corrector's instructions at T > 0) disappears.  Same draw for both stages as in the
// reference (solvers/cuda_llg_heun.cu:79, cpu_llg_heun.cc:53-64,108-114); the result differs by rounding only.
__device__ __forceinline__ void corrector_noise_part(const JbClass &c, double sx, double sy, double sz, double n0, double n1, double n2,
                                                     double &vx, double &vy, double &vz) {
  const double xx = c.sigma * n0, xy = c.sigma * n1, xz = c.sigma * n2;
  const double ax_ = sy * xz - sz * xy, ay_ = sz * xx - sx * xz, az_ = sx * xy - sy * xx;        // s* x xi
  const double bx_ = sy * az_ - sz * ay_, by_ = sz * ax_ - sx * az_, bz_ = sx * ay_ - sy * ax_;  // s* x (s* x xi)
  vx = fma(c.c_half, fma(c.alpha, bx_, ax_), vx);
  vy = fma(c.c_half, fma(c.alpha, by_, ay_), vy);
  vz = fma(c.c_half, fma(c.alpha, bz_, az_), vz);
}

// reciprocal square root without the special-case branch of CUDA's rsqrt(): hardware seed (MUFU.RSQ64H, ~2^-23) refined by
// one third-order step y (1 + e/2 + 3 e^2/8), e = 1 - x y^2 (<= 1 ulp).  x = p.p is either > DBL_EPSILON^2 and far from
// overflow (|p| ~ 1) or the result is discarded by the caller's select.
__device__ __forceinline__ double rsqrt_nobranch(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);
  const double q = fma(e, 0.375, 0.5);
  return fma(q, e * y, y);
}
// Heun intermediate from the predictor spin (pair kernel, option recover_u): k1 is perpendicular to s, so
// s + dt k1 = lambda s* with lambda = (s.s) / (s*.s) and u = s + dt/2 k1 = (s + lambda s*) / 2.  In: (ux,uy,uz) = s_n,
// (px,py,pz) = s*; out: (ux,uy,uz) = u.  A vacancy (s = s* = 0) keeps u = 0.  1 / (s*.s) = rsqrt^2 (s*.s > 0 always: it is
// (s.s) / lambda), <= 3 ulp, i.e. the same 1e-16 absolute error a stored u carries.
__device__ __forceinline__ void recover_u(double px, double py, double pz, double &ux, double &uy, double &uz) {
  const double d = fma(px, ux, fma(py, uy, pz * uz));
  const double nn = fma(ux, ux, fma(uy, uy, uz * uz));
  const double r = rsqrt_nobranch(d);
  const double hl = (d > 4.930380657631324e-32) ? 0.5 * nn * r * r : 0.0;   // lambda / 2
  ux = fma(hl, px, 0.5 * ux); uy = fma(hl, py, 0.5 * uy); uz = fma(hl, pz, 0.5 * uz);
}
// FOLD: the predictor folds the noise part of the corrector's right-hand side into v (stored-u data flow); without it v is
// the plain Heun intermediate (and dead code when the caller does not store it: option recover_u)
template <int STAGE, bool THERMAL, bool FOLD = true>
__device__ __forceinline__ void llg_site(const JbClass &c, double sx, double sy, double sz,
                                         double hx, double hy, double hz,
                                         double n0, double n1, double n2,
                                         double ux, double uy, double uz,
                                         double &ox, double &oy, double &oz, double &vx, double &vy, double &vz) {
  if (c.power != 0) {  // uniaxial: H = K p (s.a)^(p-1) a   (uniaxial_anisotropy.cc:155-163), here / mu
    const double d = c.ax * sx + c.ay * sy + c.az * sz;
    double pw = d;
    if (c.power >= 4) pw = d * d * d;
    if (c.power >= 6) pw = pw * d * d;
    const double f = c.KpT * pw;
    hx = fma(f, c.ax, hx); hy = fma(f, c.ay, hy); hz = fma(f, c.az, hz);
  }
  if (THERMAL) { hx = fma(c.sigma, n0, hx); hy = fma(c.sigma, n1, hy); hz = fma(c.sigma, n2, hz); }  // cpu_llg_heun.cc:68-82

  const double ax_ = sy * hz - sz * hy, ay_ = sz * hx - sx * hz, az_ = sx * hy - sy * hx;        // s x h
  const double bx_ = sy * az_ - sz * ay_, by_ = sz * ax_ - sx * az_, bz_ = sx * ay_ - sy * ax_;  // s x (s x h)
  const double tx = fma(c.alpha, bx_, ax_), ty = fma(c.alpha, by_, ay_), tz = fma(c.alpha, bz_, az_);  // rhs / -gyro

  double px, py, pz;
  if (STAGE == 0) {
    vx = fma(c.c_half, tx, sx); vy = fma(c.c_half, ty, sy); vz = fma(c.c_half, tz, sz);
    px = fma(c.c_full, tx, sx); py = fma(c.c_full, ty, sy); pz = fma(c.c_full, tz, sz);
  } else {
    px = fma(c.c_half, tx, ux); py = fma(c.c_half, ty, uy); pz = fma(c.c_half, tz, uz);
  }
  const double n2_ = px * px + py * py + pz * pz;
  // |p| <= DBL_EPSILON  <=>  p.p <= DBL_EPSILON^2 : leave unchanged (select, no branch; no fp64 divide or sqrt).
  const double r = rsqrt_nobranch(n2_);
  const double inv = (n2_ > 4.930380657631324e-32) ? r : 1.0;
  ox = px * inv; oy = py * inv; oz = pz * inv;
  if (STAGE == 0 && THERMAL && FOLD) corrector_noise_part(c, ox, oy, oz, n0, n1, n2, vx, vy, vz);
}

// One stage of the RK4-LLG solver for one site (cuda_llg_rk4_kernel.cuh:36-56, cuda_rk4_base.cu:66-97, cuda_rk4_base_kernel.cuh:16,
// cuda/cuda_spin_ops.cu:4-17 with the zero-length guard of Vec3 unit_vector).  In: the stage input s (s_old, y1, y2, y3), the
// field h in Tesla (exchange + constant), the draw, s_old (stages 1-3).  Out: the next stage input -- not normalised, as in the
// reference -- or the new spin (stage 3).
// The reference keeps k1..k4 and combines them at the end; the stage inputs carry the same information (y1 = s + dt/2 k1,
// y2 = s + dt/2 k2, y3 = s + dt k3), so  s + dt/6 (k1 + 2 k2 + 2 k3 + k4) = (y1 + 2 y2 + y3 - s) / 3 + dt/6 k4  and no sum of
// k's has to be stored: stage 2 forms c = y1 + 2 y2 from the site's own y1 (aux in) and its input y2 (aux out), stage 3 uses c
// (aux in).  336 instead of 408 B of HBM traffic per spin-update; the difference to the k-sum form is rounding (~1e-16).
template <int STAGE, bool THERMAL>
__device__ __forceinline__ void rk4_site(const JbClass &c, double dt, double sx, double sy, double sz, double hx, double hy, double hz,
                                         double n0, double n1, double n2, double s0x, double s0y, double s0z,
                                         double &ax, double &ay, double &az, double &ox, double &oy, double &oz) {
  if (c.power != 0) {  // uniaxial (uniaxial_anisotropy.cc:155-163), here / mu
    const double d = c.ax * sx + c.ay * sy + c.az * sz;
    double pw = d;
    if (c.power >= 4) pw = d * d * d;
    if (c.power >= 6) pw = pw * d * d;
    const double f = c.KpT * pw;
    hx = fma(f, c.ax, hx); hy = fma(f, c.ay, hy); hz = fma(f, c.az, hz);
  }
  if (THERMAL) { hx = fma(c.sigma, n0, hx); hy = fma(c.sigma, n1, hy); hz = fma(c.sigma, n2, hz); }   // one draw per step, all four stages (cuda_rk4_base.cu:65)
  const double ax_ = sy * hz - sz * hy, ay_ = sz * hx - sx * hz, az_ = sx * hy - sy * hx;
  const double bx_ = sy * az_ - sz * ay_, by_ = sz * ax_ - sx * az_, bz_ = sx * ay_ - sy * ax_;
  const double mg = -c.gyro;
  const double kx = mg * (ax_ + c.alpha * bx_), ky = mg * (ay_ + c.alpha * by_), kz = mg * (az_ + c.alpha * bz_);
  if (STAGE == 0) {          // y1 = s_old + dt/2 k1
    const double a = 0.5 * dt;
    ox = sx + a * kx; oy = sy + a * ky; oz = sz + a * kz;
  } else if (STAGE == 1) {   // y2 = s_old + dt/2 k2
    const double a = 0.5 * dt;
    ox = s0x + a * kx; oy = s0y + a * ky; oz = s0z + a * kz;
  } else if (STAGE == 2) {   // y3 = s_old + dt k3 ; c = y1 + 2 y2
    ox = s0x + dt * kx; oy = s0y + dt * ky; oz = s0z + dt * kz;
    ax = fma(2.0, sx, ax); ay = fma(2.0, sy, ay); az = fma(2.0, sz, az);
  } else {                   // s = unit((c + y3 - s_old) / 3 + dt k4 / 6)
    const double third = 1.0 / 3.0, w = dt / 6.0;
    const double vx = fma(w, kx, (ax + sx - s0x) * third), vy = fma(w, ky, (ay + sy - s0y) * third), vz = fma(w, kz, (az + sz - s0z) * third);
    const double n2_ = vx * vx + vy * vy + vz * vz;
    const double r = rsqrt_nobranch(n2_);
    const double inv = (n2_ > 4.930380657631324e-32) ? r : 1.0;
    ox = vx * inv; oy = vy * inv; oz = vz * inv;
  }
}

// store the ghost images of a freshly computed spin (the value for its own cell has been stored by the
// caller): periodic images in y/z inside this box; x images into the lo/hi boxes, which are this box
// itself on one GPU and the neighbours' boxes -- peer memory over NVLink -- on several.
// Only called for sites within a ghost depth of a face.
struct JbOutBoxes {
  double *out[3];
  double *out_lo[3];
  double *out_hi[3];
};
__device__ __forceinline__ void store_images_inline(const JbGeom &g, const JbOutBoxes &o, int x, int y, int m, int z,
                                                    double vx, double vy, double vz) {
  const bool yb = g.per[1] && ((y < g.gy) | (y >= g.Ny - g.gy));
  const bool zb = g.per[2] && ((z < g.gz) | (z >= g.Nz - g.gz));
  int yps[2], zps[2], ny = 1, nz = 1;
  yps[0] = y + g.gy; zps[0] = z + g.oz;
  if (yb) yps[ny++] = (y < g.gy) ? y + g.gy + g.Ny : y + g.gy - g.Ny;
  if (zb) zps[nz++] = (z < g.gz) ? z + g.oz + g.Nz : z + g.oz - g.Nz;
  // x targets: 0 = own, 1 = lo box, 2 = hi box
  for (int xt = 0; xt < 3; ++xt) {
    double *const *arr;
    int xp;
    if (xt == 0) { arr = o.out; xp = x + g.gx; }
    else if (xt == 1) { if (!(x < g.gx) || o.out_lo[0] == nullptr) continue; arr = o.out_lo; xp = x + g.gx + g.nx; }
    else { if (!(x >= g.nx - g.gx) || o.out_hi[0] == nullptr) continue; arr = o.out_hi; xp = x + g.gx - g.nx; }
    for (int a = 0; a < ny; ++a) {
      for (int b = 0; b < nz; ++b) {
        if (xt == 0 && a == 0 && b == 0) continue;
        const long long i = gidx(g, xp, yps[a], m, zps[b]);
        arr[0][i] = vx; arr[1][i] = vy; arr[2][i] = vz;
      }
    }
  }
}

static __device__ __noinline__ void store_images(const JbGeom &g, const JbOutBoxes &o, int x, int y, int m, int z,
                                                  double vx, double vy, double vz) {
  store_images_inline(g, o, x, y, m, z, vx, vy, vz);
}

// does a site need store_images()?
__device__ __forceinline__ bool yz_image_needed(const JbGeom &g, int y, int z) {
  return (g.per[1] && ((y < g.gy) | (y >= g.Ny - g.gy))) | (g.per[2] && ((z < g.gz) | (z >= g.Nz - g.gz)));
}
__device__ __forceinline__ bool x_image_needed(const JbGeom &g, int x) { return (x < g.gx) | (x >= g.nx - g.gx); }

// ---- warp / block reductions (monitors) ------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace jbdev

#endif  // JB_DEVICE_CUH
