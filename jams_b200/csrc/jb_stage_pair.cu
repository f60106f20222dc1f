// jb_stage_pair.cu — the hot kernel, fourth generation: one fused LLG-Heun stage (exchange gather + uniaxial +
// Zeeman + Langevin noise + LLG right-hand side + Heun update + renormalisation) as a persistent, TMA-fed,
// warp-specialised kernel for sm_100a in which every consumer thread owns a PAIR of z-adjacent sites.
//
// Replaces per stage: cusparseSpMV (containers/sparse_matrix.h:366-379), the per-Hamiltonian
// field kernels/copies (hamiltonian/cuda_uniaxial_anisotropy_kernel.cuh:14-26, cuda_zeeman.cu:27-41), the
// cudaMemcpy + cublasDaxpy field summation (cuda/cuda_solver.cc:11-26), curandGenerateNormalDouble + scale
// (thermostats/cuda_thermostat_classical.cc:47-56), the s -> s_old snapshot (solvers/cuda_llg_heun.cu:71-75) and
// cuda_heun_llg_kernelA/B (solvers/cuda_llg_heun_kernel.cuh:8-104).  Arithmetic follows the CPU solver
// (solvers/cpu_llg_heun.cc:45-148).
//
// Why pairs (profiles/README.md, r01d/r01e): the one-site-per-thread kernel moved the ideal number of DRAM bytes
// but executed ~410 warp-instructions per 32 sites, which put the issue-limited time next to the HBM time with
// little overlap.  With a pair per thread
//   * every shared-memory gather of an even z offset is one LDS.128 for two sites (odd offsets: two LDS.64),
//   * the template entry (LDS.128: byte offset + coupling), the ring/barrier bookkeeping and all address
//     arithmetic are paid once per two (SPT = 2: four) sites,
//   * results leave as STG.128, u arrives as LDS.128,
//   * two independent LLG evaluations per thread double the instruction-level parallelism.
// The periodic z images of boundary sites are written inline (three predicated stores) instead of through the
// out-of-line general routine, which half of all warps would otherwise enter on a 4-tile z split.
//
// Pipeline (unchanged in spirit): work items = (x-chunk, yz-column tile); resident CTAs march along x; the last
// warp is the TMA producer feeding a ring of R plane slots (full/empty mbarriers per slot), consumers never meet
// at a CTA-wide barrier.  The ring can be deeper than before (up to 12 slots): 2 gx + 1 planes are resident, the
// rest are loads in flight — the bytes in flight per SM are what keeps HBM busy.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "jb_tma.cuh"

namespace {

using namespace jbdev;

// barrier block = {fullS, emptyS, fullU, emptyU} x JB_PAIR_BARS
#define JB_PAIR_BARS JB_PAIR_MAX_RING

// SPT: y sites per thread (x 2 z sites).  MOTIF1: one motif site, class constants through the constant bank.
// 288 threads x 2 CTAs = 18 warps per SM = 5 per scheduler: 16384 / 5 -> at most 96 registers per thread (ptxas picks that
// from the launch bounds; 104-112 registers silently drop the kernel to one CTA per SM: measured 0.32 instead of 0.24 ms)
//
// RECU ("recover u", option `recover_u`): the Heun intermediate is not stored at all.  k1 is perpendicular
// to s_n, so s_n + dt k1 = lambda s* with lambda = (s_n.s_n) / (s*.s_n), hence u = s_n + dt/2 k1 = (s_n + lambda s*) / 2:
// the predictor writes only s* (48 instead of 72 B per site), the corrector reads the site's own s_n through the ring that
// otherwise carries u (same tile box, tensor map over the S box), rebuilds u in registers and writes s_{n+1} in place
// (72 B) -- 120 instead of 144 B of HBM traffic per spin-update.  The corrector then draws the site's noise itself
// (THERMAL instantiation), as the reference does (solvers/cuda_llg_heun.cu:79: one draw per step, used by both stages).
//
// Noise warp (option `noise_warp`, THERMAL && MOTIF1 && SPT == 1): the Philox / Box-Muller evaluation -- 30 % of a consumer
// warp's instructions at T > 0, all of them on its critical path -- moves to one more specialised warp that runs ahead of the
// consumers and leaves the draws of a plane (3 fp32 per site: they are exact fp32 values) in a two-slot shared-memory ring
// with its own full / empty mbarriers.  1 = all draws; 2 = the odd-z site of every pair (the consumer draws the even one).
template <int STAGE, bool THERMAL, bool ISO, int SPT, bool MOTIF1, bool RECU>
__global__ void __launch_bounds__(320, 2) stage_pair_kernel(const __grid_constant__ CUtensorMap tS0,
                                                            const __grid_constant__ CUtensorMap tS1,
                                                            const __grid_constant__ CUtensorMap tS2,
                                                            const __grid_constant__ CUtensorMap tU0,
                                                            const __grid_constant__ CUtensorMap tU1,
                                                            const __grid_constant__ CUtensorMap tU2,
                                                            const __grid_constant__ JbTileParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const JbGeom &g = p.g;
  const int M = MOTIF1 ? 1 : g.M, gx = g.gx;
  const int R = p.R, RU = p.RU;
  const int slotS = p.slotS, slotU = p.slotU;
  double *ringS = reinterpret_cast<double *>(smem_raw);
  double *ringU = ringS + (size_t)R * 3 * slotS;
  constexpr bool NOISEW = THERMAL && MOTIF1 && SPT == 1;
  const int nw = NOISEW ? p.noise_warp : 0;        // 0 = consumers draw their own noise, 1 / 2 = a noise warp draws all / half of it
  float *ringN = reinterpret_cast<float *>(ringU + (STAGE == 1 ? (size_t)RU * 3 * slotU : 0));   // 2 slots x 3 components x slotU floats
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(ringN + (nw ? 2 * 3 * slotU : 0));
  unsigned long long *fullS = bars, *emptyS = bars + JB_PAIR_BARS, *fullU = bars + 2 * JB_PAIR_BARS, *emptyU = bars + 3 * JB_PAIR_BARS;
  unsigned long long *fullN = bars + 4 * JB_PAIR_BARS, *emptyN = fullN + 2;
  JbTileNbr *s_nbr = reinterpret_cast<JbTileNbr *>(bars + 4 * JB_PAIR_BARS + 4);

  const int tid = threadIdx.x;
  const int n_cw = (blockDim.x >> 5) - 1 - (nw ? 1 : 0);   // consumer warps; warp n_cw is the producer, warp n_cw + 1 the noise warp
  const int G = gridDim.x, bid = blockIdx.x;

  if (tid == 0) {
    for (int s = 0; s < JB_PAIR_BARS; ++s) {
      mbar_init(smem_u32(&fullS[s]), 1); mbar_init(smem_u32(&emptyS[s]), n_cw);
      mbar_init(smem_u32(&fullU[s]), 1); mbar_init(smem_u32(&emptyU[s]), n_cw);
    }
    // noise ring: every thread that wrote (read) a slot arrives itself, so each access is ordered by its own release
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&fullN[s]), 32); mbar_init(smem_u32(&emptyN[s]), 32 * n_cw); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // per ring phase c (= slot of the oldest resident plane) and template entry n: byte offset of the neighbour
  // relative to the thread's own pair in slot 0, and the coupling -> one LDS.128 and one add per entry
  for (int idx = tid; idx < R * p.n_nbr; idx += blockDim.x) {
    const int c = idx / p.n_nbr, n = idx - c * p.n_nbr;
    const JbTileNbr e = p.nbr[n];
    int t = c + e.d;
    if (t >= R) t -= R;
    JbTileNbr o;
    o.delta = (t * 3 * slotS + e.delta) * (int)sizeof(double);
    o.d = e.d;
    o.J = e.J;
    s_nbr[idx] = o;
  }
  __syncthreads();

  // =========================== producer warp: the stream of S planes and u planes ===========================
  const int warp_idx = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction
  if (warp_idx == n_cw) {
    uint32_t elected = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(elected));
    if (!elected) return;
    const uint32_t bytesS = (uint32_t)(p.BY * M * p.BZ * sizeof(double));
    const uint32_t bytesU = (uint32_t)(p.TY * M * p.UZ * sizeof(double));
    int slot = 0, uslot = 0;
    uint32_t pe = 0xffffffffu, pue = 0xffffffffu;   // parity to wait for on each empty barrier (first pass: passes at once)
    const int lh = p.load_hint;    // experiments: 0 = none, 1 = evict_first on u, 2 = evict_first on u and evict_last on S, 3 = evict_first on both
    const unsigned long long polF = make_policy(0), polL = make_policy(1);
    for (int item0 = bid; item0 < p.n_items; item0 += G) {
      const int item = p.reverse_items ? p.n_items - 1 - item0 : item0;
      const ItemGeom it = item_geom(p, item);
      const int np = it.xc + 2 * gx;
      const int zs = it.z0 + g.oz - p.gzb;   // first column of the spin box: even, i.e. 16-byte aligned (TMA requirement)
      for (int j = 0; j < np; ++j) {
        {
          mbar_wait(smem_u32(&emptyS[slot]), (pe >> slot) & 1u);
          pe ^= 1u << slot;
          const uint32_t bar = smem_u32(&fullS[slot]);
          double *dst = ringS + (size_t)slot * 3 * slotS;
          mbar_expect_tx(bar, 3 * bytesS);
          if (lh >= 2) {
            const unsigned long long pol = lh == 2 ? polL : polF;
            tma_load_3d_hint(smem_u32(dst), &tS0, zs, it.y0 * M, it.x0 + j, bar, pol);
            tma_load_3d_hint(smem_u32(dst + slotS), &tS1, zs, it.y0 * M, it.x0 + j, bar, pol);
            tma_load_3d_hint(smem_u32(dst + 2 * slotS), &tS2, zs, it.y0 * M, it.x0 + j, bar, pol);
          } else {
            tma_load_3d(smem_u32(dst), &tS0, zs, it.y0 * M, it.x0 + j, bar);
            tma_load_3d(smem_u32(dst + slotS), &tS1, zs, it.y0 * M, it.x0 + j, bar);
            tma_load_3d(smem_u32(dst + 2 * slotS), &tS2, zs, it.y0 * M, it.x0 + j, bar);
          }
          slot = (slot + 1 == R) ? 0 : slot + 1;
        }
        if (STAGE == 1 && j >= 2 * gx) {   // the u plane of step i = j - 2 gx is needed together with S plane j
          mbar_wait(smem_u32(&emptyU[uslot]), (pue >> uslot) & 1u);
          pue ^= 1u << uslot;
          const uint32_t bar = smem_u32(&fullU[uslot]);
          double *dst = ringU + (size_t)uslot * 3 * slotU;
          mbar_expect_tx(bar, 3 * bytesU);
          const int c0 = it.z0 + g.oz, c1 = (it.y0 + g.gy) * M, c2 = it.x0 + (j - 2 * gx) + gx;   // oz, z0 even: aligned
          if (lh >= 1) {   // u is read exactly once
            tma_load_3d_hint(smem_u32(dst), &tU0, c0, c1, c2, bar, polF);
            tma_load_3d_hint(smem_u32(dst + slotU), &tU1, c0, c1, c2, bar, polF);
            tma_load_3d_hint(smem_u32(dst + 2 * slotU), &tU2, c0, c1, c2, bar, polF);
          } else {
            tma_load_3d(smem_u32(dst), &tU0, c0, c1, c2, bar);
            tma_load_3d(smem_u32(dst + slotU), &tU1, c0, c1, c2, bar);
            tma_load_3d(smem_u32(dst + 2 * slotU), &tU2, c0, c1, c2, bar);
          }
          uslot = (uslot + 1 == RU) ? 0 : uslot + 1;
        }
      }
    }
    return;
  }

  // =========================== noise warp: the draws of every plane, one plane ahead of the consumers ===========================
  if (NOISEW && nw && warp_idx == n_cw + 1) {
    const int lane = tid & 31;
    const int nq = nw == 2 ? p.TY * (p.UZ >> 1) : p.TY * p.UZ;   // work units per plane: odd-z sites, or all sites
    int nslot = 0;
    uint32_t pne = 0xffffffffu;
    for (int item0 = bid; item0 < p.n_items; item0 += G) {
      const int item = p.reverse_items ? p.n_items - 1 - item0 : item0;
      const ItemGeom it = item_geom(p, item);
      unsigned long long gs0 = global_site(g, it.x0, it.y0, 0, it.z0);   // M == 1
      for (int i = 0; i < it.xc; ++i) {
        mbar_wait(smem_u32(&emptyN[nslot]), (pne >> nslot) & 1u);
        pne ^= 1u << nslot;
        float *dst = ringN + (size_t)nslot * 3 * slotU;
#pragma unroll 2
        for (int q = lane; q < nq; q += 32) {
          const int sidx = nw == 2 ? 2 * q + 1 : q;          // site of the tile: row ty, column zt
          const int ty = sidx / p.UZ, zt = sidx - ty * p.UZ;
          if (it.y0 + ty < g.Ny && it.z0 + zt < g.Nz) {
            float a, b, c;
            site_normals_rk_f(p.rk, p.step, gs0 + (unsigned long long)ty * g.Nz + zt, a, b, c);
            dst[sidx] = a; dst[slotU + sidx] = b; dst[2 * slotU + sidx] = c;
          }
        }
        mbar_arrive(smem_u32(&fullN[nslot]));
        nslot ^= 1;
        gs0 += (unsigned long long)g.Ny * g.Nz;
      }
    }
    return;
  }

  // =========================== consumers: SPT y rows x M motif sites x one z pair each ===========================
  const int HZ = (p.TZ + 1) >> 1;                        // pairs per tile row
  const int zp = tid % HZ, tyg = tid / HZ;
  const bool padding = tyg * SPT >= p.TY;               // threads that only fill up the last consumer warp
  const int ty0 = padding ? 0 : tyg * SPT;
  const uint32_t cs8 = (uint32_t)slotS * 8u;             // component stride inside a slot, bytes
  const uint32_t slot8 = 3u * cs8;                       // slot stride, bytes
  const uint32_t cu8 = (uint32_t)slotU * 8u;
  const uint32_t kS8 = (uint32_t)(M * p.BZ) * 8u;        // strides between the thread's consecutive y rows, bytes
  const uint32_t kU8 = (uint32_t)(M * p.UZ) * 8u;
  const int kG = M * g.PZ;
  // own pair (k = 0, m = 0, component x) in slot 0 of the S ring / the u ring
  const uint32_t own = smem_u32(ringS) + (uint32_t)(((ty0 + g.gy) * M) * p.BZ + 2 * zp + p.gzb) * 8u;
  const uint32_t uown = smem_u32(ringU) + (uint32_t)((ty0 * M) * p.UZ + 2 * zp) * 8u;
  const uint32_t tab0 = smem_u32(s_nbr);
  const uint32_t tabPhase = (uint32_t)p.n_nbr * 16u;
  const unsigned int kSite = (unsigned int)g.Nz * M;
  const unsigned long long planeSites = (unsigned long long)g.Ny * g.Nz * M;
  const bool lane0 = (tid & 31) == 0;

  int cslotS = 0, cslotU = 0, cslotN = 0;
  uint32_t phS = 0, phU = 0, phN = 0;
  const uint32_t nown = smem_u32(ringN) + (uint32_t)(ty0 * p.UZ + 2 * zp) * 4u;   // own pair's draws in slot 0 of the noise ring
  auto wrapS = [&](int a) { return a >= R ? a - R : a; };
  const int sh = p.store_hint;
  const unsigned long long spol = make_policy(sh == 4 ? 1 : 0);

  for (int item0 = bid; item0 < p.n_items; item0 += G) {
    const int item = p.reverse_items ? p.n_items - 1 - item0 : item0;   // see jb_capi.cu: stage B walks the lattice backwards
    const ItemGeom it = item_geom(p, item);
    const int z = it.z0 + 2 * zp;                         // first site of the pair; the second is z + 1
    unsigned ok0 = 0, ok1 = 0, ygen = 0;                  // per y row k: site z valid, site z + 1 valid, y-face row
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int y = it.y0 + ty0 + k;
      const bool row = !padding && (ty0 + k < p.TY) && (y < g.Ny) && (2 * zp < p.TZ);
      if (row && z < g.Nz) ok0 |= 1u << k;
      if (row && z + 1 < g.Nz) ok1 |= 1u << k;
      if (g.per[1] && ((y < g.gy) | (y >= g.Ny - g.gy))) ygen |= 1u << k;
    }
    // periodic z image of each site of the pair: index shift inside the row, 0 = none (ensure_ready guarantees
    // Nz >= 2 gz + 1 for periodic z, so a site is never on both faces)
    int zsh0 = 0, zsh1 = 0;
    if (g.per[2]) {
      zsh0 = (z < g.gz) ? g.Nz : ((z >= g.Nz - g.gz) ? -g.Nz : 0);
      zsh1 = (z + 1 < g.gz) ? g.Nz : ((z + 1 >= g.Nz - g.gz) ? -g.Nz : 0);
    }
    int ic = (int)gidx(g, it.x0 + gx, it.y0 + ty0 + g.gy, 0, z + g.oz);   // g.elems < 2^31 (jb_capi.cu allocate_state)
    unsigned long long gs = global_site(g, it.x0, it.y0 + ty0, 0, z);

    for (int j = 0; j < 2 * gx; ++j) {
      const int s = wrapS(cslotS + j);
      mbar_wait(smem_u32(&fullS[s]), (phS >> s) & 1u);
      phS ^= 1u << s;
    }

    for (int i = 0; i < it.xc; ++i) {
      // the noise of this plane's sites depends on nothing but (site, step): evaluate it BEFORE waiting for the plane, so
      // the Philox / Box-Muller instructions fill the time the warp would otherwise spend parked at the full barrier
      float fa0 = 0.f, fa1 = 0.f, fa2 = 0.f, fb0 = 0.f, fb1 = 0.f, fb2 = 0.f;
      if (NOISEW && nw) {
        mbar_wait(smem_u32(&fullN[cslotN]), (phN >> cslotN) & 1u);
        phN ^= 1u << cslotN;
        const uint32_t na = nown + (uint32_t)cslotN * 3u * (uint32_t)slotU * 4u;
        const float2 v0 = lds_f2(na), v1 = lds_f2(na + (uint32_t)slotU * 4u), v2 = lds_f2(na + 2u * (uint32_t)slotU * 4u);
        mbar_arrive(smem_u32(&emptyN[cslotN]));
        cslotN ^= 1;
        fb0 = v0.y; fb1 = v1.y; fb2 = v2.y;
        if (nw == 2) site_normals_rk_f(p.rk, p.step, gs, fa0, fa1, fa2);
        else { fa0 = v0.x; fa1 = v1.x; fa2 = v2.x; }
      } else if (THERMAL && MOTIF1 && SPT == 1) {
        site_normals_rk_f(p.rk, p.step, gs, fa0, fa1, fa2);
        site_normals_rk_f(p.rk, p.step, gs + 1, fb0, fb1, fb2);   // M == 1: the site at z + 1 is the next id
      }
      {
        const int s = wrapS(cslotS + 2 * gx);
        mbar_wait(smem_u32(&fullS[s]), (phS >> s) & 1u);
        phS ^= 1u << s;
      }
      if (STAGE == 1) {
        mbar_wait(smem_u32(&fullU[cslotU]), (phU >> cslotU) & 1u);
        phU ^= 1u << cslotU;
      }
      const int x = it.x0 + i;
      const bool xb = x_image_needed(g, x);
      const uint32_t cen = own + (uint32_t)wrapS(cslotS + gx) * slot8;
      const uint32_t tab = tab0 + (uint32_t)cslotS * tabPhase;
      const uint32_t uplane = uown + (uint32_t)cslotU * 3u * cu8;

      if (!(p.debug_skip & 1))
#pragma unroll 1
      for (int m = 0; m < M; ++m) {
        const JbClass &c = p.cls[MOTIF1 ? 0 : m];
        const uint32_t mo = (uint32_t)(m * p.BZ) * 8u;
        double2 sx[SPT], sy[SPT], sz[SPT], hx[SPT], hy[SPT], hz[SPT];
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          const uint32_t a = cen + mo + k * kS8;
          sx[k] = lds128(a); sy[k] = lds128(a + cs8); sz[k] = lds128(a + 2 * cs8);
          hx[k] = make_double2(c.fTx, c.fTx); hy[k] = make_double2(c.fTy, c.fTy); hz[k] = make_double2(c.fTz, c.fTz);   // constant field (Zeeman dc + ac cos wt + applied), Tesla
        }
        // exchange field in Tesla.  Entries of a motif site: first those with an even z offset (the neighbour pair
        // is 16-byte aligned: LDS.128), then the odd ones (two LDS.64); within each group in the reference's CSR
        // column order (interface/sparse_blas.h:22-25)
        int nb = p.nbr_begin[MOTIF1 ? 0 : m], no = p.nbr_odd[MOTIF1 ? 0 : m], ne = p.nbr_begin[(MOTIF1 ? 0 : m) + 1];
        if (p.debug_skip & 4) no = ne = nb;   // timing experiments: no neighbour gathers
        const uint32_t base = own + mo;
#pragma unroll 2
        for (int n = nb; n < no; ++n) {
          const int4 raw = lds_entry(tab + (uint32_t)n * 16u);   // {byte offset, d, J}
          const double J = __hiloint2double(raw.w, raw.z);
          const uint32_t q = base + (uint32_t)raw.x;
          if (ISO) {
#pragma unroll
            for (int k = 0; k < SPT; ++k) {
              const double2 a = lds128(q + k * kS8), b = lds128(q + cs8 + k * kS8), d = lds128(q + 2 * cs8 + k * kS8);
              hx[k].x = fma(J, a.x, hx[k].x); hx[k].y = fma(J, a.y, hx[k].y);
              hy[k].x = fma(J, b.x, hy[k].x); hy[k].y = fma(J, b.y, hy[k].y);
              hz[k].x = fma(J, d.x, hz[k].x); hz[k].y = fma(J, d.y, hz[k].y);
            }
          } else {
            const double *__restrict__ Jt = p.J9T + 9 * n;
            const double J0 = Jt[0], J1 = Jt[1], J2 = Jt[2], J3 = Jt[3], J4 = Jt[4], J5 = Jt[5], J6 = Jt[6], J7 = Jt[7], J8 = Jt[8];
#pragma unroll
            for (int k = 0; k < SPT; ++k) {
              const double2 a = lds128(q + k * kS8), b = lds128(q + cs8 + k * kS8), d = lds128(q + 2 * cs8 + k * kS8);
              hx[k].x += J0 * a.x + J1 * b.x + J2 * d.x; hx[k].y += J0 * a.y + J1 * b.y + J2 * d.y;
              hy[k].x += J3 * a.x + J4 * b.x + J5 * d.x; hy[k].y += J3 * a.y + J4 * b.y + J5 * d.y;
              hz[k].x += J6 * a.x + J7 * b.x + J8 * d.x; hz[k].y += J6 * a.y + J7 * b.y + J8 * d.y;
            }
          }
        }
#pragma unroll 2
        for (int n = no; n < ne; ++n) {
          const int4 raw = lds_entry(tab + (uint32_t)n * 16u);
          const double J = __hiloint2double(raw.w, raw.z);
          const uint32_t q = base + (uint32_t)raw.x;
          if (ISO) {
#pragma unroll
            for (int k = 0; k < SPT; ++k) {
              const uint32_t qq = q + k * kS8;
              const double a0 = lds64(qq), a1 = lds64(qq + 8), b0 = lds64(qq + cs8), b1 = lds64(qq + cs8 + 8);
              const double d0 = lds64(qq + 2 * cs8), d1 = lds64(qq + 2 * cs8 + 8);
              hx[k].x = fma(J, a0, hx[k].x); hx[k].y = fma(J, a1, hx[k].y);
              hy[k].x = fma(J, b0, hy[k].x); hy[k].y = fma(J, b1, hy[k].y);
              hz[k].x = fma(J, d0, hz[k].x); hz[k].y = fma(J, d1, hz[k].y);
            }
          } else {
            const double *__restrict__ Jt = p.J9T + 9 * n;
            const double J0 = Jt[0], J1 = Jt[1], J2 = Jt[2], J3 = Jt[3], J4 = Jt[4], J5 = Jt[5], J6 = Jt[6], J7 = Jt[7], J8 = Jt[8];
#pragma unroll
            for (int k = 0; k < SPT; ++k) {
              const uint32_t qq = q + k * kS8;
              const double a0 = lds64(qq), a1 = lds64(qq + 8), b0 = lds64(qq + cs8), b1 = lds64(qq + cs8 + 8);
              const double d0 = lds64(qq + 2 * cs8), d1 = lds64(qq + 2 * cs8 + 8);
              hx[k].x += J0 * a0 + J1 * b0 + J2 * d0; hx[k].y += J0 * a1 + J1 * b1 + J2 * d1;
              hy[k].x += J3 * a0 + J4 * b0 + J5 * d0; hy[k].y += J3 * a1 + J4 * b1 + J5 * d1;
              hz[k].x += J6 * a0 + J7 * b0 + J8 * d0; hz[k].y += J6 * a1 + J7 * b1 + J8 * d1;
            }
          }
        }
        // early release: the oldest S plane (at the end of an item: all resident planes) is only read by the gathers
        // above, so its slot can go back to the producer while this warp still does the per-site physics
        if (m == M - 1) {
          __syncwarp();
          if (lane0) {
            mbar_arrive(smem_u32(&emptyS[cslotS]));
            if (i == it.xc - 1) for (int j = 1; j <= 2 * gx; ++j) mbar_arrive(smem_u32(&emptyS[wrapS(cslotS + j)]));
          }
        }
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          double2 ux = make_double2(0, 0), uy = ux, uz = ux;
          if (STAGE == 1) {
            const uint32_t ua = uplane + (uint32_t)(m * p.UZ) * 8u + k * kU8;
            ux = lds128(ua); uy = lds128(ua + cu8); uz = lds128(ua + 2 * cu8);
            if (RECU) {   // the ring delivered s_n: rebuild u = (s_n + lambda s*) / 2
              recover_u(sx[k].x, sy[k].x, sz[k].x, ux.x, uy.x, uz.x);
              recover_u(sx[k].y, sy[k].y, sz[k].y, ux.y, uy.y, uz.y);
            }
          }
          double na0 = 0, na1 = 0, na2 = 0, nb0 = 0, nb1 = 0, nb2 = 0;
          if (THERMAL) {
            if (MOTIF1 && SPT == 1) {   // drawn before the barrier wait
              na0 = (double)fa0; na1 = (double)fa1; na2 = (double)fa2; nb0 = (double)fb0; nb1 = (double)fb1; nb2 = (double)fb2;
            } else {
              const unsigned long long site = gs + k * kSite + m;
              site_normals_rk(p.rk, p.step, site, na0, na1, na2);
              site_normals_rk(p.rk, p.step, site + M, nb0, nb1, nb2);   // z + 1: the next site id but M - 1
            }
          }
          double2 ox, oy, oz, vx, vy, vz;
          if (p.debug_skip & 8) {   // timing experiments: no per-site physics, the pipeline only moves data
            ox = make_double2(sx[k].x + ux.x, sx[k].y + ux.y); oy = make_double2(sy[k].x + uy.x, sy[k].y + uy.y); oz = make_double2(sz[k].x + uz.x, sz[k].y + uz.y);
            vx = hx[k]; vy = hy[k]; vz = hz[k];
          } else {
          llg_site<STAGE, THERMAL>(c, sx[k].x, sy[k].x, sz[k].x, hx[k].x, hy[k].x, hz[k].x, na0, na1, na2, ux.x, uy.x, uz.x,
                                   ox.x, oy.x, oz.x, vx.x, vy.x, vz.x);
          llg_site<STAGE, THERMAL>(c, sx[k].y, sy[k].y, sz[k].y, hx[k].y, hy[k].y, hz[k].y, nb0, nb1, nb2, ux.y, uy.y, uz.y,
                                   ox.y, oy.y, oz.y, vx.y, vy.y, vz.y);
          }
          if (p.debug_skip & 2) {   // timing experiments: no stores
            if (ox.x + oy.x + oz.x + vx.x + vy.x + vz.x + ox.y + oy.y + oz.y + vx.y + vy.y + vz.y == 1.2345e300) p.out[0][0] = ox.x;
            continue;
          }
          const int idx = ic + m * g.PZ + k * kG;
          if ((ok1 >> k) & 1u) {          // both sites: 16-byte stores
            if (sh == 0) {
              if (STAGE == 0 && !RECU) { stg128(&p.u[0][idx], vx.x, vx.y); stg128(&p.u[1][idx], vy.x, vy.y); stg128(&p.u[2][idx], vz.x, vz.y); }
              stg128(&p.out[0][idx], ox.x, ox.y); stg128(&p.out[1][idx], oy.x, oy.y); stg128(&p.out[2][idx], oz.x, oz.y);
            } else {
              if (STAGE == 0 && !RECU) { stg128_hint(&p.u[0][idx], vx.x, vx.y, sh, spol); stg128_hint(&p.u[1][idx], vy.x, vy.y, sh, spol); stg128_hint(&p.u[2][idx], vz.x, vz.y, sh, spol); }
              stg128_hint(&p.out[0][idx], ox.x, ox.y, sh, spol); stg128_hint(&p.out[1][idx], oy.x, oy.y, sh, spol); stg128_hint(&p.out[2][idx], oz.x, oz.y, sh, spol);
            }
          } else if ((ok0 >> k) & 1u) {   // odd Nz: the last pair of a row holds one site
            if (STAGE == 0 && !RECU) { p.u[0][idx] = vx.x; p.u[1][idx] = vy.x; p.u[2][idx] = vz.x; }
            p.out[0][idx] = ox.x; p.out[1][idx] = oy.x; p.out[2][idx] = oz.x;
          }
          if (!(xb | ((ygen >> k) & 1u))) {
            // only a z face: its periodic image sits in the same row
            if (zsh0 != 0 && ((ok0 >> k) & 1u)) { p.out[0][idx + zsh0] = ox.x; p.out[1][idx + zsh0] = oy.x; p.out[2][idx + zsh0] = oz.x; }
            if (zsh1 != 0 && ((ok1 >> k) & 1u)) { p.out[0][idx + 1 + zsh1] = ox.y; p.out[1][idx + 1 + zsh1] = oy.y; p.out[2][idx + 1 + zsh1] = oz.y; }
          } else {
            if ((ok0 >> k) & 1u) tile_store_images(p, x, it.y0 + ty0 + k, m, z, ox.x, oy.x, oz.x);
            if ((ok1 >> k) & 1u) tile_store_images(p, x, it.y0 + ty0 + k, m, z + 1, ox.y, oy.y, oz.y);
          }
        }
      }
      // this warp is done with the u plane (and, in the timing experiment, with the oldest S plane)
      __syncwarp();
      if (lane0) {
        if (p.debug_skip & 1) {
          mbar_arrive(smem_u32(&emptyS[cslotS]));
          if (i == it.xc - 1) for (int j = 1; j <= 2 * gx; ++j) mbar_arrive(smem_u32(&emptyS[wrapS(cslotS + j)]));
        }
        if (STAGE == 1) mbar_arrive(smem_u32(&emptyU[cslotU]));
      }
      cslotS = wrapS(cslotS + 1);
      if (i == it.xc - 1) cslotS = wrapS(cslotS + 2 * gx);
      if (STAGE == 1) cslotU = (cslotU + 1 == RU) ? 0 : cslotU + 1;
      ic += (int)g.sX;
      gs += planeSites;
    }
  }
}

template <typename F>
cudaError_t with_kernel(int stage, int thermal, int iso, int spt, int motif1, int recu, F &&f) {
#define JB_PAIR_CASE(ST, TH, IS, SP, M1, RU_) \
  if (stage == ST && thermal == TH && iso == IS && spt == SP && motif1 == M1 && recu == RU_) \
    return f(stage_pair_kernel<ST, (TH != 0), (IS != 0), SP, (M1 != 0), (RU_ != 0)>);
#define JB_PAIR_CASES_SPT(ST, TH, IS) \
  JB_PAIR_CASE(ST, TH, IS, 1, 0, 0) JB_PAIR_CASE(ST, TH, IS, 2, 0, 0) JB_PAIR_CASE(ST, TH, IS, 1, 1, 0) JB_PAIR_CASE(ST, TH, IS, 2, 1, 0) \
  JB_PAIR_CASE(ST, TH, IS, 1, 0, 1) JB_PAIR_CASE(ST, TH, IS, 2, 0, 1) JB_PAIR_CASE(ST, TH, IS, 1, 1, 1) JB_PAIR_CASE(ST, TH, IS, 2, 1, 1)
  JB_PAIR_CASES_SPT(0, 0, 0) JB_PAIR_CASES_SPT(0, 0, 1) JB_PAIR_CASES_SPT(0, 1, 0) JB_PAIR_CASES_SPT(0, 1, 1)
  JB_PAIR_CASES_SPT(1, 0, 0) JB_PAIR_CASES_SPT(1, 0, 1) JB_PAIR_CASES_SPT(1, 1, 0) JB_PAIR_CASES_SPT(1, 1, 1)
#undef JB_PAIR_CASES_SPT
#undef JB_PAIR_CASE
  return cudaErrorInvalidValue;
}

}  // namespace

// the kernel's NOISEW condition, host side
static bool noise_warp_on(const JbTileParams &p, int thermal, int spt) { return p.noise_warp != 0 && thermal != 0 && p.g.M == 1 && spt == 1; }

cudaError_t jbk_stage_pair_occupancy(const JbTileParams &p, int stage, int thermal, int iso, int spt, int threads,
                                     size_t smem_bytes, int *blocks_per_sm) {
  return with_kernel(stage, thermal, iso, spt, p.g.M == 1 ? 1 : 0, p.recover_u ? 1 : 0, [&](auto k) -> cudaError_t {
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, ((threads + 31) & ~31) + 32 + (noise_warp_on(p, thermal, spt) ? 32 : 0), smem_bytes);
  });
}

cudaError_t jbk_stage_pair(const JbTileParams &p, const CUtensorMap *tm, int stage, int thermal, int iso, int spt,
                           int threads, int grid, size_t smem_bytes, cudaStream_t stream) {
  return with_kernel(stage, thermal, iso, spt, p.g.M == 1 ? 1 : 0, p.recover_u ? 1 : 0, [&](auto k) -> cudaError_t {
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (err != cudaSuccess) return err;
    k<<<grid, ((threads + 31) & ~31) + 32 + (noise_warp_on(p, thermal, spt) ? 32 : 0), smem_bytes, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
    return cudaGetLastError();
  });
}
