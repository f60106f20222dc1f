// jb_stage_pair.cu — the hot kernel: one fused LLG-Heun stage (exchange gather + uniaxial + Zeeman + Langevin noise +
// LLG right-hand side + Heun update + renormalisation) as a persistent, TMA-fed, warp-specialised kernel for sm_100a in
// which every consumer thread owns a PAIR of z-adjacent sites.
//
// Replaces per stage: cusparseSpMV (containers/sparse_matrix.h:366-379), the per-Hamiltonian
// field kernels/copies (hamiltonian/cuda_uniaxial_anisotropy_kernel.cuh:14-26, cuda_zeeman.cu:27-41), the
// cudaMemcpy + cublasDaxpy field summation (cuda/cuda_solver.cc:11-26), curandGenerateNormalDouble + scale
// (thermostats/cuda_thermostat_classical.cc:47-56), the s -> s_old snapshot (solvers/cuda_llg_heun.cu:71-75) and
// cuda_heun_llg_kernelA/B (solvers/cuda_llg_heun_kernel.cuh:8-104).  Arithmetic follows the CPU solver
// (solvers/cpu_llg_heun.cc:45-148).
//
// Structure (fifth generation; history in profiles/README.md):
//   * work items = (x-chunk, yz-column tile of TY x TZ cells), handed out by an ATOMIC WORK QUEUE in the order of a plan
//     (JbTileParams::chunk_x0 / chunk_xc): the slab's two face chunks first, then long chunks, short ones last.  Round 1 walked a
//     static list (bid, bid + G, ...) and left 15 % of the SM-time of a launch idle behind straggling CTAs.
//   * the last warp of a CTA is the producer: one elected thread draws item ids from the queue (one item ahead, so the fetch
//     latency hides behind the last planes of the current item), publishes them to the consumers through a small ring in
//     shared memory and streams the item's planes-with-halo through 3-D TMA boxes into a ring of R slots (full / empty
//     mbarriers per slot).  Consumers never meet at a CTA-wide barrier.
//   * a consumer thread owns the sites (z, z + 1) of one (y, m) row: every gather of an even z offset is one LDS.128 for both
//     sites, results leave as STG.128, two independent LLG evaluations per thread (ILP 2).  Multi-site motifs: the motif index is
//     spread over the threads (JbTileParams::msplit), so that a bcc tile of 4 x 64 cells still runs 256 consumer threads.
//   * RECU ("recover u"): the Heun intermediate is not stored.  k1 is perpendicular to s_n, so
//     s_n + dt k1 = lambda s* with lambda = (s_n.s_n) / (s*.s_n) and u = (s_n + lambda s*) / 2: the predictor writes only s*
//     (48 instead of 72 B per site), the corrector reads the site's own s_n through the second ring, rebuilds u in registers and
//     writes s_{n+1} in place -- 120 instead of 144 B of HBM traffic per spin-update.  The corrector then draws the site's noise
//     itself (the same Philox draw: one draw per step used by both stages, solvers/cuda_llg_heun.cu:79).
//   * slab-decomposed runs: ghost images of x-face sites are stored straight into the neighbour's box (peer
//     memory over NVLink), and the epoch handshake that orders those stores lives in this kernel too: the producer of an item
//     that reads ghost planes polls this rank's flag (ld.acquire.sys) before the first ghost plane, and the last consumer warp
//     to finish the face items of a side publishes the stage's epoch in the neighbour's flag (st.release.sys).  Face items come
//     first in the queue, so the flags travel while the interior of the slab is still being computed.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "jb_stage_common.cuh"

namespace {

using namespace jbdev;

// store a freshly computed pair (sites z, z + 1 of one row) at element index `idx` of a box, together with the periodic z
// images of its sites (zsh0 / zsh1: index shift of the image inside the row, 0 = none)
__device__ __forceinline__ void put_pair(double *const *box, long long idx, bool ok0, bool ok1, int zsh0, int zsh1,
                                         const double2 &ox, const double2 &oy, const double2 &oz) {
  if (ok1) {          // both sites: 16-byte stores
    stg128(&box[0][idx], ox.x, ox.y); stg128(&box[1][idx], oy.x, oy.y); stg128(&box[2][idx], oz.x, oz.y);
  } else if (ok0) {   // odd Nz: the last pair of a row holds one site
    box[0][idx] = ox.x; box[1][idx] = oy.x; box[2][idx] = oz.x;
  }
  if (zsh0 != 0) { box[0][idx + zsh0] = ox.x; box[1][idx + zsh0] = oy.x; box[2][idx + zsh0] = oz.x; }
  if (zsh1 != 0) { box[0][idx + 1 + zsh1] = ox.y; box[1][idx + 1 + zsh1] = oy.y; box[2][idx + 1 + zsh1] = oz.y; }
}

// x-face planes (two per slab and stage): the pair and all its y / z images once more, into the box that holds the x image --
// this slab's own box on one GPU, the neighbour's box (peer memory over NVLink) in a slab-decomposed run
static __device__ __noinline__ void put_pair_x_images(const JbTileParams &p, int x, long long idx, long long ysh, bool ok0, bool ok1,
                                                      int zsh0, int zsh1, double2 ox, double2 oy, double2 oz) {
  const JbGeom &g = p.g;
  const long long span = (long long)g.nx * g.sX;
  if (x < g.gx && p.out_lo[0] != nullptr) {
    put_pair(p.out_lo, idx + span, ok0, ok1, zsh0, zsh1, ox, oy, oz);
    if (ysh != 0) put_pair(p.out_lo, idx + span + ysh, ok0, ok1, zsh0, zsh1, ox, oy, oz);
  }
  if (x >= g.nx - g.gx && p.out_hi[0] != nullptr) {
    put_pair(p.out_hi, idx - span, ok0, ok1, zsh0, zsh1, ox, oy, oz);
    if (ysh != 0) put_pair(p.out_hi, idx - span + ysh, ok0, ok1, zsh0, zsh1, ox, oy, oz);
  }
}

// RK4 = true: one of the four stages of the RK4-LLG solver (llg-rk4-gpu: solvers/cuda_rk4_base.cu:50-108, cuda_llg_rk4_kernel.cuh:11-58)
// on the same machinery -- STAGE 0..3, the S ring streams the stage input (s_old, y1, y2, y3) with halos, the second ring the
// tile's own s_old (stages 1-3); no sum of k's is kept (rk4_site, jb_device.cuh): 48 + 72 + 120 + 96 = 336 B of HBM traffic
// per spin-update; DESIGN.md 3.2b.
template <int STAGE, bool THERMAL, bool ISO, bool MOTIF1, bool RECU, bool RK4 = false>
__global__ void __launch_bounds__(288, 2) stage_pair_kernel(const __grid_constant__ CUtensorMap tS0,
                                                            const __grid_constant__ CUtensorMap tS1,
                                                            const __grid_constant__ CUtensorMap tS2,
                                                            const __grid_constant__ CUtensorMap tU0,
                                                            const __grid_constant__ CUtensorMap tU1,
                                                            const __grid_constant__ CUtensorMap tU2,
                                                            const __grid_constant__ JbTileParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr bool HAS_U = RK4 ? (STAGE >= 1) : (STAGE == 1);   // a second ring with the tile's own plane of another tensor
  const JbGeom &g = p.g;
  const int M = MOTIF1 ? 1 : g.M, gx = g.gx;
  const int R = p.R, RU = p.RU;
  const int slotS = p.slotS, slotU = p.slotU;
  double *ringS = reinterpret_cast<double *>(smem_raw);
  double *ringU = ringS + (size_t)R * 3 * slotS;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(ringU + (HAS_U ? (size_t)RU * 3 * slotU : 0));
  unsigned long long *fullS = bars, *emptyS = bars + JB_PAIR_BARS, *fullU = bars + 2 * JB_PAIR_BARS, *emptyU = bars + 3 * JB_PAIR_BARS;
  volatile int *items = reinterpret_cast<volatile int *>(bars + 4 * JB_PAIR_BARS);
  unsigned int *face_arrivals = reinterpret_cast<unsigned int *>(bars + 4 * JB_PAIR_BARS + JB_ITEM_RING / 2);   // [0] lo, [1] hi
  JbTileNbr *s_nbr = reinterpret_cast<JbTileNbr *>(bars + JB_STAGE_TAIL_WORDS);

  const int tid = threadIdx.x;
  const int n_cw = (blockDim.x >> 5) - 1;   // consumer warps; warp n_cw is the producer

  if (tid == 0) {
    for (int s = 0; s < JB_PAIR_BARS; ++s) {
      mbar_init(smem_u32(&fullS[s]), 1); mbar_init(smem_u32(&emptyS[s]), n_cw);
      mbar_init(smem_u32(&fullU[s]), 1); mbar_init(smem_u32(&emptyU[s]), n_cw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    face_arrivals[0] = face_arrivals[1] = 0u;
    if (blockIdx.x == 0) *p.queue_next = 0u;   // the counter of the NEXT launch on this stream
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

  // =========================== producer warp: item ids, the stream of S planes and u planes ===========================
  const int warp_idx = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction
  if (warp_idx == n_cw) {
    stage_producer<HAS_U>(&tS0, &tS1, &tS2, &tU0, &tU1, &tU2, p, M, ringS, ringU, fullS, emptyS, fullU, emptyU, items);
    return;
  }

  // =========================== consumers: M motif sites x one z pair of one y row each ===========================
  const int HZ = (p.TZ + 1) >> 1;                        // pairs per tile row
  const int MS = MOTIF1 ? 1 : p.msplit;                  // threads per (y row, z pair): this one owns motif sites m0, m0 + MS, ...
  const int zp = tid % HZ, trow = tid / HZ;
  const int tyr = MOTIF1 ? trow : trow / MS, m0 = MOTIF1 ? 0 : trow - tyr * MS;
  const bool padding = tyr >= p.TY;                      // threads that only fill up the last consumer warp
  const int ty = padding ? 0 : tyr;
  const uint32_t cs8 = (uint32_t)slotS * 8u;             // component stride inside a slot, bytes
  const uint32_t slot8 = 3u * cs8;                       // slot stride, bytes
  const uint32_t cu8 = (uint32_t)slotU * 8u;
  const uint32_t uslot8 = 3u * cu8;
  // own pair (m = 0, component x) in slot 0 of the S ring / the u ring
  const uint32_t own = smem_u32(ringS) + (uint32_t)(((ty + g.gy) * M) * p.BZ + 2 * zp + p.gzb) * 8u;
  const uint32_t uown = smem_u32(ringU) + (uint32_t)((ty * M) * p.UZ + 2 * zp) * 8u;
  const uint32_t tab0 = smem_u32(s_nbr);
  const uint32_t tabPhase = (uint32_t)p.n_nbr * 16u;
  const uint32_t fullS0 = smem_u32(fullS), emptyS0 = smem_u32(emptyS), fullU0 = smem_u32(fullU), emptyU0 = smem_u32(emptyU);
  const unsigned long long planeSites = (unsigned long long)g.Ny * g.Nz * M;
  const bool lane0 = (tid & 31) == 0;
  const int sX = (int)g.sX;

  int wslot = 0, uslot = 0, qi = 0;     // next S plane / u plane to wait for, next item-ring entry
  uint32_t wpar = 0u, upar = 0u;
  unsigned long long t_first = 0;
  int n_done = 0;
  if (p.trace && tid == 0) t_first = global_timer_ns();

  for (;;) {
    // the first plane of the next item -- or the producer's "no more work"
    const int oslot0 = wslot;
    mbar_wait(fullS0 + 8u * wslot, wpar);
    if (++wslot == R) { wslot = 0; wpar ^= 1u; }
    const int item = items[qi];
    qi = (qi + 1) & (JB_ITEM_RING - 1);
    if (item < 0) break;
    if (p.trace && tid == 0 && n_done < JB_TRACE_ITEMS)   // item id and start time relative to the CTA's first clock
      p.trace[(unsigned long long)JB_TRACE_WORDS * blockIdx.x + 4 + n_done] = ((unsigned long long)item << 40) | ((global_timer_ns() - t_first) & 0xffffffffffull);
    ++n_done;
    const ItemGeom it = item_geom(p, item);
    const int z = it.z0 + 2 * zp;                         // first site of the pair; the second is z + 1
    const int y = it.y0 + ty;
    const bool row = !padding && y < g.Ny;
    const bool ok0 = row && z < g.Nz, ok1 = row && z + 1 < g.Nz;
    // periodic y image of the row: index shift, 0 = none (ensure_ready guarantees Ny >= 2 gy + 1 for periodic y)
    int ysh = 0;
    if (g.per[1] && row) ysh = (y < g.gy) ? g.Ny * (int)g.sY : ((y >= g.Ny - g.gy) ? -g.Ny * (int)g.sY : 0);
    // periodic z image of each site of the pair: index shift inside the row, 0 = none (ensure_ready guarantees
    // Nz >= 2 gz + 1 for periodic z, so a site is never on both faces)
    int zsh0 = 0, zsh1 = 0;
    if (g.per[2]) {
      zsh0 = (z < g.gz) ? g.Nz : ((z >= g.Nz - g.gz) ? -g.Nz : 0);
      zsh1 = (z + 1 < g.gz) ? g.Nz : ((z + 1 >= g.Nz - g.gz) ? -g.Nz : 0);
    }
    if (!ok0) zsh0 = 0;
    if (!ok1) zsh1 = 0;
    int ic = (int)gidx(g, it.x0 + gx, y + g.gy, 0, z + g.oz);   // g.elems < 2^31 (jb_capi.cu allocate_state)
    unsigned long long gs = global_site(g, it.x0, y, 0, z);      // the pair's even-z site, m = 0: the noise key
    const bool face_lo = it.x0 < gx, face_hi = it.x0 + it.xc > g.nx - gx;

    for (int j = 1; j < 2 * gx; ++j) {
      mbar_wait(fullS0 + 8u * wslot, wpar);
      if (++wslot == R) { wslot = 0; wpar ^= 1u; }
    }
    int oslot = oslot0;                                    // slot of the oldest resident plane (x - gx)
    int cslot = oslot0 + gx; if (cslot >= R) cslot -= R;   // slot of the centre plane

    for (int i = 0; i < it.xc; ++i) {
      // the noise of this plane's sites depends on nothing but (site, step): evaluate it BEFORE waiting for the plane, so
      // the Philox / Box-Muller instructions fill the time the warp would otherwise spend parked at the full barrier
      PairNormals nz;
      if (THERMAL && MOTIF1) pair_normals_rk(p.rk, p.step, gs, nz);
      if (gx > 0 || i > 0) {
        mbar_wait(fullS0 + 8u * wslot, wpar);
        if (++wslot == R) { wslot = 0; wpar ^= 1u; }
      }
      if (HAS_U) mbar_wait(fullU0 + 8u * uslot, upar);
      const int x = it.x0 + i;
      const bool xb = x_image_needed(g, x);
      const uint32_t cen = own + (uint32_t)cslot * slot8;
      const uint32_t tab = tab0 + (uint32_t)oslot * tabPhase;
      const uint32_t uplane = uown + (uint32_t)uslot * uslot8;

#pragma unroll 1
      for (int m = m0; m < M; m += MS) {
        const JbClass &c = p.cls[MOTIF1 ? 0 : m];
        const uint32_t mo = (uint32_t)(m * p.BZ) * 8u;
        const uint32_t a = cen + mo;
        const double2 sx = lds128(a), sy = lds128(a + cs8), sz = lds128(a + 2 * cs8);
        // RK4: stage 2 needs the site's own y1 (it sits in the box this stage overwrites with y3), stage 3 the combination
        // c = y1 + 2 y2 that stage 2 left in the U box: straight from global memory, in flight during the gather
        double2 kx = make_double2(0, 0), ky = kx, kz = kx;
        if (RK4 && STAGE >= 2) {
          const int idx = ic + m * g.PZ;
          double *const *src = STAGE == 2 ? p.out : p.u;
          if (ok1) { kx = *reinterpret_cast<const double2 *>(&src[0][idx]); ky = *reinterpret_cast<const double2 *>(&src[1][idx]); kz = *reinterpret_cast<const double2 *>(&src[2][idx]); }
          else if (ok0) { kx.x = src[0][idx]; ky.x = src[1][idx]; kz.x = src[2][idx]; }
        }
        double2 hx = make_double2(c.fTx, c.fTx), hy = make_double2(c.fTy, c.fTy), hz = make_double2(c.fTz, c.fTz);   // constant field (Zeeman dc + ac cos wt + applied), Tesla
        // exchange field in Tesla.  Entries of a motif site: first those with an even z offset (the neighbour pair
        // is 16-byte aligned: LDS.128), then the odd ones (two LDS.64); within each group in the reference's CSR
        // column order (interface/sparse_blas.h:22-25)
        const int nb = p.nbr_begin[MOTIF1 ? 0 : m], no = p.nbr_odd[MOTIF1 ? 0 : m], ne = p.nbr_begin[(MOTIF1 ? 0 : m) + 1];
        const uint32_t base = own + mo;
#pragma unroll 4
        for (int n = nb; n < no; ++n) {
          const int4 raw = lds_entry(tab + (uint32_t)n * 16u);   // {byte offset, d, J}
          const uint32_t q = base + (uint32_t)raw.x;
          const double2 va = lds128(q), vb = lds128(q + cs8), vd = lds128(q + 2 * cs8);
          if (ISO) {
            const double J = __hiloint2double(raw.w, raw.z);
            hx.x = fma(J, va.x, hx.x); hx.y = fma(J, va.y, hx.y);
            hy.x = fma(J, vb.x, hy.x); hy.y = fma(J, vb.y, hy.y);
            hz.x = fma(J, vd.x, hz.x); hz.y = fma(J, vd.y, hz.y);
          } else {
            const double *__restrict__ Jt = p.J9T + 9 * n;
            const double J0 = Jt[0], J1 = Jt[1], J2 = Jt[2], J3 = Jt[3], J4 = Jt[4], J5 = Jt[5], J6 = Jt[6], J7 = Jt[7], J8 = Jt[8];
            hx.x += J0 * va.x + J1 * vb.x + J2 * vd.x; hx.y += J0 * va.y + J1 * vb.y + J2 * vd.y;
            hy.x += J3 * va.x + J4 * vb.x + J5 * vd.x; hy.y += J3 * va.y + J4 * vb.y + J5 * vd.y;
            hz.x += J6 * va.x + J7 * vb.x + J8 * vd.x; hz.y += J6 * va.y + J7 * vb.y + J8 * vd.y;
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
            const double *__restrict__ Jt = p.J9T + 9 * n;
            const double J0 = Jt[0], J1 = Jt[1], J2 = Jt[2], J3 = Jt[3], J4 = Jt[4], J5 = Jt[5], J6 = Jt[6], J7 = Jt[7], J8 = Jt[8];
            hx.x += J0 * a0 + J1 * b0 + J2 * d0; hx.y += J0 * a1 + J1 * b1 + J2 * d1;
            hy.x += J3 * a0 + J4 * b0 + J5 * d0; hy.y += J3 * a1 + J4 * b1 + J5 * d1;
            hz.x += J6 * a0 + J7 * b0 + J8 * d0; hz.y += J6 * a1 + J7 * b1 + J8 * d1;
          }
        }
        if (ISO && p.has_zself[MOTIF1 ? 0 : m]) {
          // the z neighbours inside the row: site z + 1 is the pair's other site (registers), likewise z for site z + 1; only
          // z - 1 and z + 2 come from shared memory (one LDS.64 per component each instead of two)
          const double Jm = p.zself[MOTIF1 ? 0 : m][0], Jp = p.zself[MOTIF1 ? 0 : m][1];
          const double lx = lds64(a - 8u), ly = lds64(a + cs8 - 8u), lz = lds64(a + 2 * cs8 - 8u);
          const double rx = lds64(a + 16u), ry = lds64(a + cs8 + 16u), rz = lds64(a + 2 * cs8 + 16u);
          hx.x = fma(Jm, lx, hx.x); hx.x = fma(Jp, sx.y, hx.x); hx.y = fma(Jm, sx.x, hx.y); hx.y = fma(Jp, rx, hx.y);
          hy.x = fma(Jm, ly, hy.x); hy.x = fma(Jp, sy.y, hy.x); hy.y = fma(Jm, sy.x, hy.y); hy.y = fma(Jp, ry, hy.y);
          hz.x = fma(Jm, lz, hz.x); hz.x = fma(Jp, sz.y, hz.x); hz.y = fma(Jm, sz.x, hz.y); hz.y = fma(Jp, rz, hz.y);
        }
        // early release: the oldest S plane (at the end of an item: all resident planes) is only read by the gathers
        // above, so its slot can go back to the producer while this warp still does the per-site physics
        if (m + MS >= M) {   // this thread's last motif site of the plane (msplit divides M: the same iteration for every lane)
          __syncwarp();
          if (lane0) {
            mbar_arrive(emptyS0 + 8u * oslot);
            if (i == it.xc - 1) {
              int s = oslot;
              for (int j = 1; j <= 2 * gx; ++j) { if (++s == R) s = 0; mbar_arrive(emptyS0 + 8u * s); }
            }
          }
        }
        double2 ux = make_double2(0, 0), uy = ux, uz = ux;
        if (HAS_U) {
          const uint32_t ua = uplane + (uint32_t)(m * p.UZ) * 8u;
          ux = lds128(ua); uy = lds128(ua + cu8); uz = lds128(ua + 2 * cu8);
          if (!RK4 && RECU) {   // the ring delivered s_n: rebuild u = (s_n + lambda s*) / 2
            recover_u(sx.x, sy.x, sz.x, ux.x, uy.x, uz.x);
            recover_u(sx.y, sy.y, sz.y, ux.y, uy.y, uz.y);
          }
        }
        if (THERMAL && !MOTIF1) pair_normals_rk(p.rk, p.step, gs + m, nz);
        double2 ox, oy, oz, vx = make_double2(0, 0), vy = vx, vz = vx;
        if constexpr (RK4) {
          rk4_site<STAGE, THERMAL>(c, p.dt, sx.x, sy.x, sz.x, hx.x, hy.x, hz.x, THERMAL ? (double)nz.e0 : 0.0, THERMAL ? (double)nz.e1 : 0.0,
                                   THERMAL ? (double)nz.e2 : 0.0, ux.x, uy.x, uz.x, kx.x, ky.x, kz.x, ox.x, oy.x, oz.x);
          rk4_site<STAGE, THERMAL>(c, p.dt, sx.y, sy.y, sz.y, hx.y, hy.y, hz.y, THERMAL ? (double)nz.o0 : 0.0, THERMAL ? (double)nz.o1 : 0.0,
                                   THERMAL ? (double)nz.o2 : 0.0, ux.y, uy.y, uz.y, kx.y, ky.y, kz.y, ox.y, oy.y, oz.y);
          if (STAGE == 2) {   // c = y1 + 2 y2 for the last stage: interior only, no images
            const int idx = ic + m * g.PZ;
            if (ok1) { stg128(&p.u[0][idx], kx.x, kx.y); stg128(&p.u[1][idx], ky.x, ky.y); stg128(&p.u[2][idx], kz.x, kz.y); }
            else if (ok0) { p.u[0][idx] = kx.x; p.u[1][idx] = ky.x; p.u[2][idx] = kz.x; }
          }
        } else {
        llg_site<STAGE, THERMAL, !RECU>(c, sx.x, sy.x, sz.x, hx.x, hy.x, hz.x, THERMAL ? (double)nz.e0 : 0.0, THERMAL ? (double)nz.e1 : 0.0,
                                        THERMAL ? (double)nz.e2 : 0.0, ux.x, uy.x, uz.x, ox.x, oy.x, oz.x, vx.x, vy.x, vz.x);
        llg_site<STAGE, THERMAL, !RECU>(c, sx.y, sy.y, sz.y, hx.y, hy.y, hz.y, THERMAL ? (double)nz.o0 : 0.0, THERMAL ? (double)nz.o1 : 0.0,
                                        THERMAL ? (double)nz.o2 : 0.0, ux.y, uy.y, uz.y, ox.y, oy.y, oz.y, vx.y, vy.y, vz.y);
        }
        const int idx = ic + m * g.PZ;
        if (!RK4 && STAGE == 0 && !RECU) {   // the Heun intermediate: interior only, no images
          if (ok1) { stg128(&p.u[0][idx], vx.x, vx.y); stg128(&p.u[1][idx], vy.x, vy.y); stg128(&p.u[2][idx], vz.x, vz.y); }
          else if (ok0) { p.u[0][idx] = vx.x; p.u[1][idx] = vy.x; p.u[2][idx] = vz.x; }
        }
        // the new spins and their ghost images: z images sit in the same row, the y image of a face row one lattice height away,
        // x images (two planes per slab) in the lo / hi box
        put_pair(p.out, idx, ok0, ok1, zsh0, zsh1, ox, oy, oz);
        if (ysh != 0) put_pair(p.out, idx + ysh, ok0, ok1, zsh0, zsh1, ox, oy, oz);
        if (xb) put_pair_x_images(p, x, idx, ysh, ok0, ok1, zsh0, zsh1, ox, oy, oz);
      }
      if (HAS_U) {   // this warp is done with the u plane
        __syncwarp();
        if (lane0) mbar_arrive(emptyU0 + 8u * uslot);
        if (++uslot == RU) { uslot = 0; upar ^= 1u; }
      }
      if (++oslot == R) oslot = 0;
      if (++cslot == R) cslot = 0;
      ic += sX;
      gs += planeSites;
    }
    if (p.halo.enabled && (face_lo | face_hi)) {   // this warp's stores into the neighbours' boxes are on their way
      __syncwarp();
      if (lane0) {
        if (face_lo) halo_face_done(p.halo, 0, smem_u32(face_arrivals), (unsigned int)n_cw);
        if (face_hi) halo_face_done(p.halo, 1, smem_u32(face_arrivals), (unsigned int)n_cw);
      }
    }
  }
  if (p.trace && tid == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    unsigned long long *t = p.trace + (unsigned long long)JB_TRACE_WORDS * blockIdx.x;
    t[0] = smid; t[1] = t_first; t[2] = global_timer_ns(); t[3] = (unsigned long long)n_done;
  }
}

template <typename F>
cudaError_t with_rk4_kernel(int stage, int thermal, int motif1, F &&f) {   // isotropic couplings only (tensor couplings: rk4_direct_kernel)
#define JB_RK4_CASE(ST, TH, M1) \
  if (stage == ST && thermal == TH && motif1 == M1) return f(stage_pair_kernel<ST, (TH != 0), true, (M1 != 0), false, true>);
#define JB_RK4_CASES(ST) JB_RK4_CASE(ST, 0, 0) JB_RK4_CASE(ST, 0, 1) JB_RK4_CASE(ST, 1, 0) JB_RK4_CASE(ST, 1, 1)
  JB_RK4_CASES(0) JB_RK4_CASES(1) JB_RK4_CASES(2) JB_RK4_CASES(3)
#undef JB_RK4_CASES
#undef JB_RK4_CASE
  return cudaErrorInvalidValue;
}

template <typename F>
cudaError_t with_kernel(int stage, int thermal, int iso, int motif1, int recu, F &&f) {
#define JB_PAIR_CASE(ST, TH, IS, M1, RU_) \
  if (stage == ST && thermal == TH && iso == IS && motif1 == M1 && recu == RU_) \
    return f(stage_pair_kernel<ST, (TH != 0), (IS != 0), (M1 != 0), (RU_ != 0)>);
#define JB_PAIR_CASES(ST, TH, RU_) JB_PAIR_CASE(ST, TH, 0, 0, RU_) JB_PAIR_CASE(ST, TH, 0, 1, RU_) JB_PAIR_CASE(ST, TH, 1, 0, RU_) JB_PAIR_CASE(ST, TH, 1, 1, RU_)
  // predictor: stores u (RECU 0) or not (RECU 1), with or without noise; corrector: stored u never needs noise (the predictor
  // folded it into u), a recovered u does at T > 0
  JB_PAIR_CASES(0, 0, 0) JB_PAIR_CASES(0, 1, 0) JB_PAIR_CASES(0, 0, 1) JB_PAIR_CASES(0, 1, 1)
  JB_PAIR_CASES(1, 0, 0) JB_PAIR_CASES(1, 0, 1) JB_PAIR_CASES(1, 1, 1)
#undef JB_PAIR_CASES
#undef JB_PAIR_CASE
  return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t jbk_stage_pair_occupancy(const JbTileParams &p, int stage, int thermal, int iso, int recu, int threads,
                                     size_t smem_bytes, int *blocks_per_sm) {
  return with_kernel(stage, thermal, iso, p.g.M == 1 ? 1 : 0, recu, [&](auto k) -> cudaError_t {
    cudaError_t err = jb_ensure_dynamic_smem(k, smem_bytes);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, ((threads + 31) & ~31) + 32, smem_bytes);
  });
}

cudaError_t jbk_stage_pair(const JbTileParams &p, const CUtensorMap *tm, int stage, int thermal, int iso, int recu,
                           int threads, int grid, size_t smem_bytes, cudaStream_t stream) {
  return with_kernel(stage, thermal, iso, p.g.M == 1 ? 1 : 0, recu, [&](auto k) -> cudaError_t {
    cudaError_t err = jb_ensure_dynamic_smem(k, smem_bytes);
    if (err != cudaSuccess) return err;
    k<<<grid, ((threads + 31) & ~31) + 32, smem_bytes, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
    return cudaGetLastError();
  });
}

cudaError_t jbk_rk4_stage_pair_occupancy(const JbTileParams &p, int stage, int thermal, int threads, size_t smem_bytes, int *blocks_per_sm) {
  return with_rk4_kernel(stage, thermal, p.g.M == 1 ? 1 : 0, [&](auto k) -> cudaError_t {
    cudaError_t err = jb_ensure_dynamic_smem(k, smem_bytes);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, ((threads + 31) & ~31) + 32, smem_bytes);
  });
}

cudaError_t jbk_rk4_stage_pair(const JbTileParams &p, const CUtensorMap *tm, int stage, int thermal, int threads, int grid,
                               size_t smem_bytes, cudaStream_t stream) {
  return with_rk4_kernel(stage, thermal, p.g.M == 1 ? 1 : 0, [&](auto k) -> cudaError_t {
    cudaError_t err = jb_ensure_dynamic_smem(k, smem_bytes);
    if (err != cudaSuccess) return err;
    k<<<grid, ((threads + 31) & ~31) + 32, smem_bytes, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
    return cudaGetLastError();
  });
}
