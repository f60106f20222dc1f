// jb_step_fused.cu — one whole LLG-Heun step (predictor AND corrector) in ONE launch for short-range exchange
// templates: the predictor spins s* never leave the SM.
//
// The two-launch scheme (jb_stage_pair.cu) moves 144 B per spin-update through HBM: A reads s_n and writes s* and the
// Heun intermediate u, B reads s* and u and writes s_{n+1}.  Profiles (profiles/README.md, r01f) showed both launches
// limited by the memory system at 4.8-5.4 TB/s of real DRAM traffic, not by instruction issue (39 % issue-active).
// The only way to go faster is to move fewer bytes.  Here a CTA marches along x through a yz-column tile and keeps a
// sliding window of s* planes in shared memory:
//     step k :  phase A  s*(x_a)   on the tile EXTENDED by the template reach r, from the s_n planes x_a-r .. x_a+r
//               (TMA ring, as before); results go to a ring of 4 shared-memory planes, u and the noise of the thread's
//               own pair stay in registers
//               ---- named barrier over the consumer warps ----
//               phase B  s_{n+1}(x_b), x_b = x_a - r, from the s* planes x_b-r .. x_b+r in shared memory, u and the
//               noise kept from step k - r; results to global memory (STG.128) + ghost images
// HBM traffic per spin-update: 24 B read (s_n) + 24 B written (s_{n+1}) = 48 B instead of 144 B; the price is the
// redundant predictor work on the tile's halo ring ((TY+2r)(TZ+4)/(TY TZ) = 1.33 for 8 x 64) and ghost zones of depth
// 2 r (the global boxes are allocated with g.gx = 2 rx, ... when this kernel is selected).  Sites outside an open
// boundary hold s = 0, for which the LLG update returns 0: the predictor values of ghost cells are right without any
// branch; across periodic boundaries the ghost cells hold images and the noise is keyed by the wrapped global site id,
// so every CTA (and every GPU) computes identical halo values.
//
// Reference semantics: solvers/cpu_llg_heun.cc:45-148 (one draw of noise per step, reused by both stages; the fields
// of stage B are evaluated at t + dt: :103-106); per-site formulas in jb_device.cuh.  Restrictions (else the two-launch
// kernels run): reach 1 along x, motif size 1 or 2, motif-uniform parameters.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "jb_tma.cuh"

namespace {

using namespace jbdev;

__device__ __forceinline__ void consumer_barrier(int n_threads) {
  asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

__device__ __forceinline__ int wrap_coord(int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); }

#define JB_FUSED_BARS JB_PAIR_MAX_RING
#define JB_FUSED_PSLOTS 4     // s* planes in shared memory: 2 rx + 1 read by phase B + the one phase A writes next

// exchange field of a pair: adds sum_j J_ij s_j (Tesla) over the template entries [nb, ne) of its motif site; entries
// [nb, no) have an even z offset (one LDS.128 per component), [no, ne) an odd one (two LDS.64).  `base` = shared
// address of the pair in slot 0 of the ring, `tab` = the entry table of the current ring phase (byte offsets include
// the slot), cs8 = component stride in bytes.
template <bool ISO>
__device__ __forceinline__ void gather_pair(uint32_t base, uint32_t tab, int nb, int no, int ne, uint32_t cs8,
                                            const double *__restrict__ J9T, double2 &hx, double2 &hy, double2 &hz) {
#pragma unroll 2
  for (int n = nb; n < no; ++n) {
    const int4 raw = lds_entry(tab + (uint32_t)n * 16u);   // {byte offset, d, J}
    const uint32_t q = base + (uint32_t)raw.x;
    const double2 a = lds128(q), b = lds128(q + cs8), d = lds128(q + 2 * cs8);
    if (ISO) {
      const double J = __hiloint2double(raw.w, raw.z);
      hx.x = fma(J, a.x, hx.x); hx.y = fma(J, a.y, hx.y);
      hy.x = fma(J, b.x, hy.x); hy.y = fma(J, b.y, hy.y);
      hz.x = fma(J, d.x, hz.x); hz.y = fma(J, d.y, hz.y);
    } else {
      const double *__restrict__ Jt = J9T + 9 * n;
      const double J0 = Jt[0], J1 = Jt[1], J2 = Jt[2], J3 = Jt[3], J4 = Jt[4], J5 = Jt[5], J6 = Jt[6], J7 = Jt[7], J8 = Jt[8];
      hx.x += J0 * a.x + J1 * b.x + J2 * d.x; hx.y += J0 * a.y + J1 * b.y + J2 * d.y;
      hy.x += J3 * a.x + J4 * b.x + J5 * d.x; hy.y += J3 * a.y + J4 * b.y + J5 * d.y;
      hz.x += J6 * a.x + J7 * b.x + J8 * d.x; hz.y += J6 * a.y + J7 * b.y + J8 * d.y;
    }
  }
#pragma unroll 2
  for (int n = no; n < ne; ++n) {
    const int4 raw = lds_entry(tab + (uint32_t)n * 16u);
    const uint32_t q = base + (uint32_t)raw.x;
    const double a0 = lds64(q), a1 = lds64(q + 8), b0 = lds64(q + cs8), b1 = lds64(q + cs8 + 8);
    const double d0 = lds64(q + 2 * cs8), d1 = lds64(q + 2 * cs8 + 8);
    if (ISO) {
      const double J = __hiloint2double(raw.w, raw.z);
      hx.x = fma(J, a0, hx.x); hx.y = fma(J, a1, hx.y);
      hy.x = fma(J, b0, hy.x); hy.y = fma(J, b1, hy.y);
      hz.x = fma(J, d0, hz.x); hz.y = fma(J, d1, hz.y);
    } else {
      const double *__restrict__ Jt = J9T + 9 * n;
      const double J0 = Jt[0], J1 = Jt[1], J2 = Jt[2], J3 = Jt[3], J4 = Jt[4], J5 = Jt[5], J6 = Jt[6], J7 = Jt[7], J8 = Jt[8];
      hx.x += J0 * a0 + J1 * b0 + J2 * d0; hx.y += J0 * a1 + J1 * b1 + J2 * d1;
      hy.x += J3 * a0 + J4 * b0 + J5 * d0; hy.y += J3 * a1 + J4 * b1 + J5 * d1;
      hz.x += J6 * a0 + J7 * b0 + J8 * d0; hz.y += J6 * a1 + J7 * b1 + J8 * d1;
    }
  }
}

struct PairNoise { float a0, a1, a2, b0, b1, b2; };   // three N(0,1) draws for each of the two sites of a pair (exact fp32 values)

// predictor of one pair located `off` bytes into a slot component: reads s_n (centre slot `cen`, neighbours through
// `tab`), writes s* into the s* slot at `dst`; returns u = s + dt/2 k1 of both sites and the noise it used
template <bool THERMAL, bool ISO, bool UNI>
__device__ __forceinline__ void predictor_pair(const JbTileParams &p, const JbClass &c, int mm, uint32_t ringN, uint32_t cen, uint32_t tab,
                                               uint32_t dst, uint32_t off, uint32_t cs8, unsigned long long site0, unsigned long long site1,
                                               double2 &ux, double2 &uy, double2 &uz, PairNoise &nz) {
  const uint32_t a = cen + off;
  const double2 sx = lds128(a), sy = lds128(a + cs8), sz = lds128(a + 2 * cs8);
  double2 hx = make_double2(c.fTx, c.fTx), hy = make_double2(c.fTy, c.fTy), hz = make_double2(c.fTz, c.fTz);   // constant field at time t
  gather_pair<ISO>(ringN + off, tab, p.nbr_begin[mm], p.nbr_odd[mm], p.nbr_begin[mm + 1], cs8, p.J9T, hx, hy, hz);
  nz = PairNoise{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (THERMAL) {
    site_normals_rk_f(p.rk, p.step, site0, nz.a0, nz.a1, nz.a2);
    site_normals_rk_f(p.rk, p.step, site1, nz.b0, nz.b1, nz.b2);
  }
  double2 ox, oy, oz;
  const double zero = 0.0;
  llg_site_nb<0, THERMAL, UNI>(c, sx.x, sy.x, sz.x, hx.x, hy.x, hz.x, (double)nz.a0, (double)nz.a1, (double)nz.a2, zero, zero, zero,
                               ox.x, oy.x, oz.x, ux.x, uy.x, uz.x);
  llg_site_nb<0, THERMAL, UNI>(c, sx.y, sy.y, sz.y, hx.y, hy.y, hz.y, (double)nz.b0, (double)nz.b1, (double)nz.b2, zero, zero, zero,
                               ox.y, oy.y, oz.y, ux.y, uy.y, uz.y);
  sts128(dst + off, ox.x, ox.y); sts128(dst + off + cs8, oy.x, oy.y); sts128(dst + off + 2 * cs8, oz.x, oz.y);
}

// corrector of the thread's own pair: reads s* (centre slot `cen`, neighbours through `tab`), u and the noise of the
// predictor two blocks ago; returns s_{n+1} of both sites
template <bool THERMAL, bool ISO, bool UNI>
__device__ __forceinline__ void corrector_pair(const JbTileParams &p, const JbClass &c, int mm, uint32_t ringP, uint32_t cen, uint32_t tab,
                                               uint32_t off, uint32_t cs8, const double2 &ux, const double2 &uy, const double2 &uz,
                                               const PairNoise &nz, double2 &ox, double2 &oy, double2 &oz) {
  const uint32_t a = cen + off;
  const double2 sx = lds128(a), sy = lds128(a + cs8), sz = lds128(a + 2 * cs8);
  double2 hx = make_double2(p.fT1[mm][0], p.fT1[mm][0]), hy = make_double2(p.fT1[mm][1], p.fT1[mm][1]), hz = make_double2(p.fT1[mm][2], p.fT1[mm][2]);   // constant field at t + dt
  gather_pair<ISO>(ringP + off, tab, p.nbr_begin[mm], p.nbr_odd[mm], p.nbr_begin[mm + 1], cs8, p.J9T, hx, hy, hz);
  double2 vx, vy, vz;
  (void)nz;   // the predictor has folded the noise part of this right-hand side into u (jb_device.cuh: corrector_noise_part)
  llg_site_nb<1, false, UNI>(c, sx.x, sy.x, sz.x, hx.x, hy.x, hz.x, 0.0, 0.0, 0.0, ux.x, uy.x, uz.x, ox.x, oy.x, oz.x, vx.x, vy.x, vz.x);
  llg_site_nb<1, false, UNI>(c, sx.y, sy.y, sz.y, hx.y, hy.y, hz.y, 0.0, 0.0, 0.0, ux.y, uy.y, uz.y, ox.y, oy.y, oz.y, vx.y, vy.y, vz.y);
}

// corrector of plane x_b and predictor of plane x_b + 2 rx + 1 ... of the SAME thread in one basic block: the two are
// independent (different rings), so the template entries are walked once for both and the four site updates interleave
// in the instruction stream -- twice the instruction-level parallelism of doing them one after the other.
template <bool THERMAL, bool ISO, bool UNI>
__device__ __forceinline__ void fused_ab_pair(const JbTileParams &p, const JbClass &c, int mm,
                                              uint32_t ringN, uint32_t cenN, uint32_t tabA, uint32_t dstP,
                                              uint32_t ringP, uint32_t cenP, uint32_t tabB, uint32_t off, uint32_t cs8,
                                              unsigned long long site0, unsigned long long site1,
                                              const double2 &ux_o, const double2 &uy_o, const double2 &uz_o, const PairNoise &nz_o,
                                              double2 &ox, double2 &oy, double2 &oz,
                                              double2 &ux_n, double2 &uy_n, double2 &uz_n, PairNoise &nz_n) {
  const uint32_t aA = cenN + off, aB = cenP + off;
  const double2 sxA = lds128(aA), syA = lds128(aA + cs8), szA = lds128(aA + 2 * cs8);
  const double2 sxB = lds128(aB), syB = lds128(aB + cs8), szB = lds128(aB + 2 * cs8);
  double2 hxA = make_double2(c.fTx, c.fTx), hyA = make_double2(c.fTy, c.fTy), hzA = make_double2(c.fTz, c.fTz);
  double2 hxB = make_double2(p.fT1[mm][0], p.fT1[mm][0]), hyB = make_double2(p.fT1[mm][1], p.fT1[mm][1]), hzB = make_double2(p.fT1[mm][2], p.fT1[mm][2]);
  const int nb = p.nbr_begin[mm], no = p.nbr_odd[mm], ne = p.nbr_begin[mm + 1];
  const uint32_t bA = ringN + off, bB = ringP + off;
  if (ISO) {
#pragma unroll 2
    for (int n = nb; n < no; ++n) {
      const int4 rA = lds_entry(tabA + (uint32_t)n * 16u), rB = lds_entry(tabB + (uint32_t)n * 16u);
      const double J = __hiloint2double(rA.w, rA.z);
      const uint32_t qA = bA + (uint32_t)rA.x, qB = bB + (uint32_t)rB.x;
      const double2 a = lds128(qA), b = lds128(qA + cs8), d = lds128(qA + 2 * cs8);
      const double2 e = lds128(qB), f = lds128(qB + cs8), g2 = lds128(qB + 2 * cs8);
      hxA.x = fma(J, a.x, hxA.x); hxA.y = fma(J, a.y, hxA.y); hyA.x = fma(J, b.x, hyA.x); hyA.y = fma(J, b.y, hyA.y);
      hzA.x = fma(J, d.x, hzA.x); hzA.y = fma(J, d.y, hzA.y);
      hxB.x = fma(J, e.x, hxB.x); hxB.y = fma(J, e.y, hxB.y); hyB.x = fma(J, f.x, hyB.x); hyB.y = fma(J, f.y, hyB.y);
      hzB.x = fma(J, g2.x, hzB.x); hzB.y = fma(J, g2.y, hzB.y);
    }
#pragma unroll 2
    for (int n = no; n < ne; ++n) {
      const int4 rA = lds_entry(tabA + (uint32_t)n * 16u), rB = lds_entry(tabB + (uint32_t)n * 16u);
      const double J = __hiloint2double(rA.w, rA.z);
      const uint32_t qA = bA + (uint32_t)rA.x, qB = bB + (uint32_t)rB.x;
      const double a0 = lds64(qA), a1 = lds64(qA + 8), b0 = lds64(qA + cs8), b1 = lds64(qA + cs8 + 8), d0 = lds64(qA + 2 * cs8), d1 = lds64(qA + 2 * cs8 + 8);
      const double e0 = lds64(qB), e1 = lds64(qB + 8), f0 = lds64(qB + cs8), f1 = lds64(qB + cs8 + 8), g0 = lds64(qB + 2 * cs8), g1 = lds64(qB + 2 * cs8 + 8);
      hxA.x = fma(J, a0, hxA.x); hxA.y = fma(J, a1, hxA.y); hyA.x = fma(J, b0, hyA.x); hyA.y = fma(J, b1, hyA.y);
      hzA.x = fma(J, d0, hzA.x); hzA.y = fma(J, d1, hzA.y);
      hxB.x = fma(J, e0, hxB.x); hxB.y = fma(J, e1, hxB.y); hyB.x = fma(J, f0, hyB.x); hyB.y = fma(J, f1, hyB.y);
      hzB.x = fma(J, g0, hzB.x); hzB.y = fma(J, g1, hzB.y);
    }
  } else {
    gather_pair<false>(bA, tabA, nb, no, ne, cs8, p.J9T, hxA, hyA, hzA);
    gather_pair<false>(bB, tabB, nb, no, ne, cs8, p.J9T, hxB, hyB, hzB);
  }
  nz_n = PairNoise{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (THERMAL) {
    site_normals_rk_f(p.rk, p.step, site0, nz_n.a0, nz_n.a1, nz_n.a2);
    site_normals_rk_f(p.rk, p.step, site1, nz_n.b0, nz_n.b1, nz_n.b2);
  }
  double2 sA_x, sA_y, sA_z, vx, vy, vz;
  const double zero = 0.0;
  llg_site_nb<0, THERMAL, UNI>(c, sxA.x, syA.x, szA.x, hxA.x, hyA.x, hzA.x, (double)nz_n.a0, (double)nz_n.a1, (double)nz_n.a2, zero, zero, zero,
                               sA_x.x, sA_y.x, sA_z.x, ux_n.x, uy_n.x, uz_n.x);
  (void)nz_o;   // the noise part of the corrector's right-hand side is already in u (jb_device.cuh: corrector_noise_part)
  llg_site_nb<1, false, UNI>(c, sxB.x, syB.x, szB.x, hxB.x, hyB.x, hzB.x, 0.0, 0.0, 0.0, ux_o.x, uy_o.x, uz_o.x,
                             ox.x, oy.x, oz.x, vx.x, vy.x, vz.x);
  llg_site_nb<0, THERMAL, UNI>(c, sxA.y, syA.y, szA.y, hxA.y, hyA.y, hzA.y, (double)nz_n.b0, (double)nz_n.b1, (double)nz_n.b2, zero, zero, zero,
                               sA_x.y, sA_y.y, sA_z.y, ux_n.y, uy_n.y, uz_n.y);
  llg_site_nb<1, false, UNI>(c, sxB.y, syB.y, szB.y, hxB.y, hyB.y, hzB.y, 0.0, 0.0, 0.0, ux_o.y, uy_o.y, uz_o.y,
                             ox.y, oy.y, oz.y, vx.y, vy.y, vz.y);
  sts128(dstP + off, sA_x.x, sA_x.y); sts128(dstP + off + cs8, sA_y.x, sA_y.y); sts128(dstP + off + 2 * cs8, sA_z.x, sA_z.y);
}

// Warp roles: [0, n_iw) interior warps (a thread owns the pair (z, z+1) of every motif site of one y row: predictor and
// corrector), [n_iw, n_iw + n_hw) halo warps (a thread owns one pair of the halo ring of the s* extent: predictor only),
// the last warp is the TMA producer.  MM: motif size (1 or 2).
template <bool THERMAL, bool ISO, bool UNI, int MM>
__global__ void __launch_bounds__(384, 1) step_fused_kernel(const __grid_constant__ CUtensorMap tS0,
                                                            const __grid_constant__ CUtensorMap tS1,
                                                            const __grid_constant__ CUtensorMap tS2,
                                                            const __grid_constant__ JbTileParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const JbGeom &g = p.g;
  constexpr int M = MM;
  const int rx = p.rx, ry = p.ry;
  const int R = p.R;
  const int slotS = p.slotS;
  double *ringN = reinterpret_cast<double *>(smem_raw);                    // s_n planes (TMA)
  double *ringP = ringN + (size_t)R * 3 * slotS;                           // s* planes (written by the predictors)
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(ringP + (size_t)JB_FUSED_PSLOTS * 3 * slotS);
  unsigned long long *fullN = bars, *emptyN = bars + JB_FUSED_BARS;
  JbTileNbr *s_nbrA = reinterpret_cast<JbTileNbr *>(bars + 2 * JB_FUSED_BARS);
  JbTileNbr *s_nbrB = s_nbrA + (size_t)R * p.n_nbr;

  const int tid = threadIdx.x;
  const int n_cw = (blockDim.x >> 5) - 1;          // consumer warps (interior + halo); warp n_cw is the producer
  const int n_ct = n_cw * 32;
  const int G = gridDim.x, bid = blockIdx.x;

  if (tid == 0) {
    for (int s = 0; s < JB_FUSED_BARS; ++s) { mbar_init(smem_u32(&fullN[s]), 1); mbar_init(smem_u32(&emptyN[s]), n_cw); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // entry tables per ring phase c = slot of the oldest plane the phase reads; byte offsets relative to the pair in slot 0
  for (int idx = tid; idx < (R + JB_FUSED_PSLOTS) * p.n_nbr; idx += blockDim.x) {
    const bool isB = idx >= R * p.n_nbr;
    const int i2 = isB ? idx - R * p.n_nbr : idx;
    const int depth = isB ? JB_FUSED_PSLOTS : R;
    const int c = i2 / p.n_nbr, n = i2 - c * p.n_nbr;
    const JbTileNbr e = p.nbr[n];
    int t = c + e.d;
    if (t >= depth) t -= depth;
    JbTileNbr o;
    o.delta = (t * 3 * slotS + e.delta) * (int)sizeof(double);
    o.d = e.d;
    o.J = e.J;
    (isB ? s_nbrB : s_nbrA)[i2] = o;
  }
  __syncthreads();

  // =========================== producer warp: the stream of s_n planes ===========================
  const int warp_idx = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (warp_idx == n_cw) {
    uint32_t elected = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(elected));
    if (!elected) return;
    const uint32_t bytesS = (uint32_t)(p.BY * M * p.BZ * sizeof(double));
    int slot = 0;
    uint32_t pe = 0xffffffffu;
    for (int item = bid; item < p.n_items; item += G) {
      const ItemGeom it = item_geom(p, item);
      const int np = it.xc + 4 * rx;
      const int zs = it.z0 + g.oz - p.e2z;   // even: 16-byte aligned box start
      for (int j = 0; j < np; ++j) {
        mbar_wait(smem_u32(&emptyN[slot]), (pe >> slot) & 1u);
        pe ^= 1u << slot;
        const uint32_t bar = smem_u32(&fullN[slot]);
        double *dst = ringN + (size_t)slot * 3 * slotS;
        mbar_expect_tx(bar, 3 * bytesS);
        tma_load_3d(smem_u32(dst), &tS0, zs, it.y0 * M, it.x0 + j, bar);
        tma_load_3d(smem_u32(dst + slotS), &tS1, zs, it.y0 * M, it.x0 + j, bar);
        tma_load_3d(smem_u32(dst + 2 * slotS), &tS2, zs, it.y0 * M, it.x0 + j, bar);
        slot = (slot + 1 == R) ? 0 : slot + 1;
      }
    }
    return;
  }

  // =========================== consumers ===========================
  const int HZ = (p.TZ + 1) >> 1;                        // interior pairs per tile row
  const int n_iw = (HZ * p.TY + 31) >> 5;                // interior warps
  const bool halo_role = warp_idx >= n_iw;
  const int zp = tid % HZ, tyg = tid / HZ;
  const bool active = !halo_role && tyg < p.TY;          // interior thread with a pair of its own
  const int ty0 = active ? tyg : 0;
  const uint32_t cs8 = (uint32_t)slotS * 8u;             // component stride inside a slot, bytes
  const uint32_t slot8 = 3u * cs8;                       // slot stride, bytes
  const uint32_t rowB = (uint32_t)p.BZ * 8u;             // row stride, bytes
  const uint32_t own = (uint32_t)(((ty0 + 2 * ry) * M) * p.BZ + p.e2z + 2 * zp) * 8u;   // own pair, motif row 0, inside a slot component
  const uint32_t ringN_a = smem_u32(ringN), ringP_a = smem_u32(ringP);
  const uint32_t tabA0 = smem_u32(s_nbrA), tabB0 = smem_u32(s_nbrB);
  const uint32_t tabPhase = (uint32_t)p.n_nbr * 16u;
  const unsigned long long planeSites = (unsigned long long)g.Ny * g.Nz * M;
  const bool lane0 = (tid & 31) == 0;

  // the halo ring of the s* extent: (TY + 2 ry) M rows x (HZ + e1z) pairs minus the interior; halo thread h owns pair h
  const int hz1 = p.e1z >> 1, W = HZ + 2 * hz1;
  const int n_top = ry * M * W, n_mid = p.TY * M * 2 * hz1, n_halo = 2 * n_top + n_mid;
  uint32_t hoff = 0xffffffffu; int hm = 0, hy = 0, hzr = 0;
  if (halo_role) {
    const int h = tid - n_iw * 32;
    int rE = 0, zq = 0;
    if (h < n_top) { rE = h / W; zq = h - rE * W; }
    else if (h < 2 * n_top) { const int h2 = h - n_top; const int r2 = h2 / W; rE = (ry + p.TY) * M + r2; zq = h2 - r2 * W; }
    else { const int h2 = h - 2 * n_top; const int r2 = h2 / (2 * hz1 > 0 ? 2 * hz1 : 1); const int qq = h2 - r2 * 2 * hz1; rE = ry * M + r2; zq = qq < hz1 ? qq : HZ + qq; }
    if (h < n_halo) hoff = (uint32_t)((rE + ry * M) * p.BZ + p.e2z - p.e1z + 2 * zq) * 8u;
    hm = rE % M;
    hy = rE / M - ry;            // lattice y relative to the tile's y0
    hzr = 2 * zq - p.e1z;        // lattice z (first site of the pair) relative to the tile's z0
  }

  int cslotN = 0;                // slot of the oldest resident s_n plane
  uint32_t phN = 0;
  unsigned int pA = 0;           // running count of predictor planes: the next one goes to s* slot pA % 4
  auto wrapN = [&](int a) { return a >= R ? a - R : a; };

  // the own pair's u and noise travel from its predictor to its corrector two blocks later (rx == 1): a two-deep queue
  double2 ux1[MM], uy1[MM], uz1[MM], ux2[MM], uy2[MM], uz2[MM];
  PairNoise nz1[MM], nz2[MM];
#pragma unroll
  for (int m = 0; m < MM; ++m) {
    ux1[m] = uy1[m] = uz1[m] = ux2[m] = uy2[m] = uz2[m] = make_double2(0, 0);
    nz1[m] = nz2[m] = PairNoise{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  }

  for (int item = bid; item < p.n_items; item += G) {
    const ItemGeom it = item_geom(p, item);
    const int z = it.z0 + 2 * zp;
    const int y = it.y0 + ty0;
    const bool row_ok = active && (y < g.Ny);
    const bool ok0 = row_ok && z < g.Nz, ok1 = row_ok && z + 1 < g.Nz;
    const bool ygen = g.per[1] && ((y < g.gy) | (y >= g.Ny - g.gy));
    int zsh0 = 0, zsh1 = 0;     // periodic z image of each site: shift inside the row, 0 = none (jb_capi.cu guarantees Nz >= 2 gz)
    if (g.per[2]) {
      zsh0 = (z < g.gz) ? g.Nz : ((z >= g.Nz - g.gz) ? -g.Nz : 0);
      zsh1 = (z + 1 < g.gz) ? g.Nz : ((z + 1 >= g.Nz - g.gz) ? -g.Nz : 0);
    }
    int ic = (int)gidx(g, it.x0 + g.gx, y + g.gy, 0, z + g.oz);   // corrector output index of plane x0
    // global site ids without the x part (wrapped: ghost cells carry the noise of the site they image)
    unsigned int rid0[MM], rid1[MM];
    {
      const int yy = halo_role ? it.y0 + hy : y, zz = halo_role ? it.z0 + hzr : z;
      const int yw = wrap_coord(yy, g.Ny), zw0 = wrap_coord(zz, g.Nz), zw1 = wrap_coord(zz + 1, g.Nz);
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        const int mq = halo_role ? hm : m;
        rid0[m] = ((unsigned int)yw * g.Nz + zw0) * M + mq;
        rid1[m] = ((unsigned int)yw * g.Nz + zw1) * M + mq;
      }
    }

    for (int j = 0; j < 2 * rx; ++j) {
      const int s = wrapN(cslotN + j);
      mbar_wait(smem_u32(&fullN[s]), (phN >> s) & 1u);
      phN ^= 1u << s;
    }

    // block j: corrector of plane x0 + j - 2 rx (if j >= 2 rx) together with predictor of plane x0 - rx + j + 1 (if there is one)
    const int last = it.xc + 2 * rx - 1;     // predictor planes are numbered 0 .. last
    for (int j = -1; j <= last; ++j) {
      const bool doA = j + 1 <= last, doB = j >= 2 * rx;
      if (doA) {
        const int s = wrapN(cslotN + 2 * rx);
        mbar_wait(smem_u32(&fullN[s]), (phN >> s) & 1u);
        phN ^= 1u << s;
      }
      const int pslot = (int)(pA & 3u);
      const uint32_t cenN = ringN_a + (uint32_t)wrapN(cslotN + rx) * slot8;
      const uint32_t tabA = tabA0 + (uint32_t)cslotN * tabPhase;
      const uint32_t dstP = ringP_a + (uint32_t)pslot * slot8;
      const int cP = (int)((pA - 1u - 2u * (unsigned)rx) & 3u);            // slot of s*(x_b - rx)
      const uint32_t cenP = ringP_a + (uint32_t)((cP + rx) & 3) * slot8;
      const uint32_t tabB = tabB0 + (uint32_t)cP * tabPhase;
      unsigned long long xbase = 0;
      if (THERMAL) xbase = (unsigned long long)wrap_coord(g.x_begin + it.x0 - rx + j + 1, g.Nx_global) * planeSites;

      if (halo_role) {
        if (doA && hoff != 0xffffffffu) {
          double2 dx, dy, dz; PairNoise dn;
          predictor_pair<THERMAL, ISO, UNI>(p, p.cls[MM == 1 ? 0 : hm], MM == 1 ? 0 : hm, ringN_a, cenN, tabA, dstP, hoff, cs8,
                                            xbase + rid0[0], xbase + rid1[0], dx, dy, dz, dn);
        }
      } else if (active) {
        const int x = it.x0 + j - 2 * rx;
        const bool xb = doB && x_image_needed(g, x);
#pragma unroll
        for (int m = 0; m < MM; ++m) {
          const uint32_t off = own + m * rowB;
          double2 ox, oy, oz, ux_n = ux1[m], uy_n = uy1[m], uz_n = uz1[m];
          PairNoise nz_n = nz1[m];
          if (doA && doB) {
            fused_ab_pair<THERMAL, ISO, UNI>(p, p.cls[m], m, ringN_a, cenN, tabA, dstP, ringP_a, cenP, tabB, off, cs8, xbase + rid0[m], xbase + rid1[m],
                                             ux2[m], uy2[m], uz2[m], nz2[m], ox, oy, oz, ux_n, uy_n, uz_n, nz_n);
          } else if (doA) {
            predictor_pair<THERMAL, ISO, UNI>(p, p.cls[m], m, ringN_a, cenN, tabA, dstP, off, cs8, xbase + rid0[m], xbase + rid1[m], ux_n, uy_n, uz_n, nz_n);
          } else {
            corrector_pair<THERMAL, ISO, UNI>(p, p.cls[m], m, ringP_a, cenP, tabB, off, cs8, ux2[m], uy2[m], uz2[m], nz2[m], ox, oy, oz);
          }
          ux2[m] = ux1[m]; uy2[m] = uy1[m]; uz2[m] = uz1[m]; nz2[m] = nz1[m];
          ux1[m] = ux_n; uy1[m] = uy_n; uz1[m] = uz_n; nz1[m] = nz_n;
          if (doB) {
            const int idx = ic + m * g.PZ;
            if (ok1) { stg128(&p.out[0][idx], ox.x, ox.y); stg128(&p.out[1][idx], oy.x, oy.y); stg128(&p.out[2][idx], oz.x, oz.y); }
            else if (ok0) { p.out[0][idx] = ox.x; p.out[1][idx] = oy.x; p.out[2][idx] = oz.x; }
            if (!(xb | ygen)) {
              if (zsh0 != 0 && ok0) { p.out[0][idx + zsh0] = ox.x; p.out[1][idx + zsh0] = oy.x; p.out[2][idx + zsh0] = oz.x; }
              if (zsh1 != 0 && ok1) { p.out[0][idx + 1 + zsh1] = ox.y; p.out[1][idx + 1 + zsh1] = oy.y; p.out[2][idx + 1 + zsh1] = oz.y; }
            } else {
              if (ok0) tile_store_images(p, x, y, m, z, ox.x, oy.x, oz.x);
              if (ok1) tile_store_images(p, x, y, m, z + 1, ox.y, oy.y, oz.y);
            }
          }
        }
        if (doB) ic += (int)g.sX;
      }
      if (doA) {
        // the oldest s_n plane (after the last predictor of an item: every resident plane) is not read again by this warp
        __syncwarp();
        if (lane0) {
          mbar_arrive(smem_u32(&emptyN[cslotN]));
          if (j + 1 == last) for (int q = 1; q <= 2 * rx; ++q) mbar_arrive(smem_u32(&emptyN[wrapN(cslotN + q)]));
        }
        cslotN = wrapN(cslotN + 1);
        if (j + 1 == last) cslotN = wrapN(cslotN + 2 * rx);
        ++pA;
      }
      // s* of this block complete before the next block's correctors read it; the next block's predictors overwrite the
      // s* plane this block's correctors read as their oldest only after every warp has passed here
      consumer_barrier(n_ct);
    }
  }
}

template <typename F>
cudaError_t with_kernel(int thermal, int iso, int uni, int mm, F &&f) {
#define JB_FUSED_CASE(TH, IS, UN, MMV) \
  if (thermal == TH && iso == IS && uni == UN && mm == MMV) return f(step_fused_kernel<(TH != 0), (IS != 0), (UN != 0), MMV>);
#define JB_FUSED_CASES(TH, IS) JB_FUSED_CASE(TH, IS, 0, 1) JB_FUSED_CASE(TH, IS, 1, 1) JB_FUSED_CASE(TH, IS, 0, 2) JB_FUSED_CASE(TH, IS, 1, 2)
  JB_FUSED_CASES(0, 0) JB_FUSED_CASES(0, 1) JB_FUSED_CASES(1, 0) JB_FUSED_CASES(1, 1)
#undef JB_FUSED_CASES
#undef JB_FUSED_CASE
  return cudaErrorInvalidValue;
}

}  // namespace

// `threads` = interior threads (pairs per tile row x rows), `halo_warps` = warps that compute the halo ring of s*;
// the launch adds the producer warp
cudaError_t jbk_step_fused_occupancy(const JbTileParams &p, int thermal, int iso, int uni, int threads, int halo_warps, size_t smem_bytes, int *blocks_per_sm) {
  return with_kernel(thermal, iso, uni, p.g.M, [&](auto k) -> cudaError_t {
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, ((threads + 31) & ~31) + 32 * halo_warps + 32, smem_bytes);
  });
}

cudaError_t jbk_step_fused(const JbTileParams &p, const CUtensorMap *tm3, int thermal, int iso, int uni, int threads, int halo_warps, int grid,
                           size_t smem_bytes, cudaStream_t stream) {
  return with_kernel(thermal, iso, uni, p.g.M, [&](auto k) -> cudaError_t {
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (err != cudaSuccess) return err;
    k<<<grid, ((threads + 31) & ~31) + 32 * halo_warps + 32, smem_bytes, stream>>>(tm3[0], tm3[1], tm3[2], p);
    return cudaGetLastError();
  });
}
