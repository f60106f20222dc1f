// jb_stage_tile.cu — the hot kernel: one fused LLG-Heun stage (exchange gather + uniaxial + Zeeman + Langevin
// noise + LLG right-hand side + Heun update + renormalisation) as a PERSISTENT, TMA-fed kernel for sm_100a.
//
// Replaces per stage: cusparseSpMV (containers/sparse_matrix.h:366-379), the per-Hamiltonian
// field kernels/copies (hamiltonian/cuda_uniaxial_anisotropy_kernel.cuh:14-26, cuda_zeeman.cu:27-41), the
// cudaMemcpy + cublasDaxpy field summation (cuda/cuda_solver.cc:11-26), curandGenerateNormalDouble + scale
// (thermostats/cuda_thermostat_classical.cc:47-56), the s -> s_old snapshot (solvers/cuda_llg_heun.cu:71-75) and
// cuda_heun_llg_kernelA/B (solvers/cuda_llg_heun_kernel.cuh:8-104).  Arithmetic follows the CPU solver
// (solvers/cpu_llg_heun.cc:45-148).
//
// Structure (DESIGN.md "Kernels"):
//  * The lattice is cut into work items = (x-chunk, yz-column tile of TY x TZ cells).  gridDim.x CTAs (a
//    multiple of the SM count) stay resident and take items bid, bid + G, ... : concurrently running CTAs work
//    on neighbouring columns of the same x-chunk, so tile halos are shared through L2.
//  * A CTA marches along x.  Every plane-with-halo of the input spins is one 3-D TMA box per component
//    (cp.async.bulk.tensor) landing in a ring of R shared-memory slots, completion signalled on an mbarrier per
//    slot.  The plane stream continues across item boundaries, so the pipeline never drains.  In stage B the
//    Heun intermediate u of the centre plane arrives the same way in a second, shallower ring.
//  * Warp specialisation: the last warp of the CTA is the TMA producer; the other warps are consumers and never
//    meet at a CTA-wide barrier.  A "full" mbarrier per slot carries the TMA transaction count, an "empty"
//    mbarrier per slot collects one arrival per consumer warp when the slot's plane is no longer needed.
//  * Each consumer thread owns SPT y-sites x M motif sites of one z column.  The exchange template (16 B per
//    entry, coupling already divided by mu) sits in shared memory, the class constants and Philox round keys
//    in the kernel-parameter bank: no per-site global loads besides the spins themselves.
//  * Results go straight from registers to global memory (256 B per warp and component); sites within a ghost
//    depth of a face also store their periodic / neighbour-slab images (peer memory over NVLink).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "jb_tma.cuh"

namespace {

using namespace jbdev;

#define JB_TILE_BARS 8   // max ring depth; barrier block = {fullS, emptyS, fullU, emptyU} x JB_TILE_BARS

// MOTIF1: the lattice has one motif site (M == 1): no motif loop, class constants through the uniform datapath
template <int STAGE, bool THERMAL, bool ISO, int SPT, bool MOTIF1>
__global__ void __launch_bounds__(SPT == 1 ? 512 : 288, SPT == 1 ? 2 : 3) stage_tile_kernel(const __grid_constant__ CUtensorMap tS0,
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
  const bool use_u = (STAGE == 1) && p.u_tma;
  double *ringS = reinterpret_cast<double *>(smem_raw);
  double *ringU = ringS + (size_t)R * 3 * slotS;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(ringU + (use_u ? (size_t)RU * 3 * slotU : 0));
  unsigned long long *fullS = bars, *emptyS = bars + JB_TILE_BARS, *fullU = bars + 2 * JB_TILE_BARS, *emptyU = bars + 3 * JB_TILE_BARS;
  JbTileNbr *s_nbr = reinterpret_cast<JbTileNbr *>(bars + 4 * JB_TILE_BARS);

  const int tid = threadIdx.x;
  const int n_cw = (blockDim.x >> 5) - 1;          // consumer warps; warp n_cw is the producer
  const int G = gridDim.x, bid = blockIdx.x;

  if (tid == 0) {
    for (int s = 0; s < JB_TILE_BARS; ++s) {
      mbar_init(smem_u32(&fullS[s]), 1); mbar_init(smem_u32(&emptyS[s]), n_cw);
      mbar_init(smem_u32(&fullU[s]), 1); mbar_init(smem_u32(&emptyU[s]), n_cw);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // per ring phase c (= slot of the oldest resident plane) and template entry n: byte offset of the neighbour
  // relative to the thread's own site in slot 0, and the coupling -> one LDS.128 and one add per neighbour
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
  // warp index through a shuffle: the compiler then knows it is warp-uniform and keeps the producer's TMA operands in
  // uniform registers (no per-lane election loops around UTMALDG)
  const int warp_idx = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (warp_idx == n_cw) {
    uint32_t elected = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(elected));
    if (!elected) return;
    const uint32_t bytesS = (uint32_t)(p.BY * M * p.BZ * sizeof(double));
    const uint32_t bytesU = (uint32_t)(p.TY * M * p.UZ * sizeof(double));
    int slot = 0, uslot = 0;
    uint32_t pe = 0xffffffffu, pue = 0xffffffffu;   // parity to wait for on each empty barrier (first pass: passes at once)
    for (int item = bid; item < p.n_items; item += G) {
      const ItemGeom it = item_geom(p, item);
      const int np = it.xc + 2 * gx;
      const int zs = it.z0 + g.oz - p.gzb;   // first column of the spin box: even, i.e. 16-byte aligned (TMA requirement)
      for (int j = 0; j < np; ++j) {
        {
          mbar_wait_backoff(smem_u32(&emptyS[slot]), (pe >> slot) & 1u, p.producer_sleep_ns);
          pe ^= 1u << slot;
          const uint32_t bar = smem_u32(&fullS[slot]);
          double *dst = ringS + (size_t)slot * 3 * slotS;
          mbar_expect_tx(bar, 3 * bytesS);
          tma_load_3d(smem_u32(dst), &tS0, zs, it.y0 * M, it.x0 + j, bar);
          tma_load_3d(smem_u32(dst + slotS), &tS1, zs, it.y0 * M, it.x0 + j, bar);
          tma_load_3d(smem_u32(dst + 2 * slotS), &tS2, zs, it.y0 * M, it.x0 + j, bar);
          slot = (slot + 1 == R) ? 0 : slot + 1;
        }
        if (use_u && j >= 2 * gx) {   // the u plane of step i = j - 2 gx is needed together with S plane j
          mbar_wait_backoff(smem_u32(&emptyU[uslot]), (pue >> uslot) & 1u, p.producer_sleep_ns);
          pue ^= 1u << uslot;
          const uint32_t bar = smem_u32(&fullU[uslot]);
          double *dst = ringU + (size_t)uslot * 3 * slotU;
          mbar_expect_tx(bar, 3 * bytesU);
          const int c0 = it.z0 + g.oz, c1 = (it.y0 + g.gy) * M, c2 = it.x0 + (j - 2 * gx) + gx;   // oz, z0 even: aligned
          tma_load_3d(smem_u32(dst), &tU0, c0, c1, c2, bar);
          tma_load_3d(smem_u32(dst + slotU), &tU1, c0, c1, c2, bar);
          tma_load_3d(smem_u32(dst + 2 * slotU), &tU2, c0, c1, c2, bar);
          uslot = (uslot + 1 == RU) ? 0 : uslot + 1;
        }
      }
    }
    return;
  }

  // =========================== consumers: SPT y-sites x M motif sites of one z column each ===========================
  const int tz = tid % p.TZ, tyg = tid / p.TZ;
  const bool padding = tyg * SPT >= p.TY;            // threads that only fill up the last consumer warp
  const int ty0 = padding ? 0 : tyg * SPT;
  const int soff = ((ty0 + g.gy) * M) * p.BZ + tz + p.gzb;   // centre of (k = 0, m = 0) inside a slot component
  const int uoff = (ty0 * M) * p.UZ + tz;
  const int kS = M * p.BZ, kU = M * p.UZ, kG = M * g.PZ;    // strides between the thread's consecutive y sites
  const unsigned int kSite = (unsigned int)g.Nz * M;
  const unsigned long long planeSites = (unsigned long long)g.Ny * g.Nz * M;
  const bool lane0 = (tid & 31) == 0;
  const int slot3 = 3 * slotS;

  int cslotS = 0, cslotU = 0;
  uint32_t phS = 0, phU = 0;
  auto wrapS = [&](int a) { return a >= R ? a - R : a; };

  for (int item = bid; item < p.n_items; item += G) {
    const ItemGeom it = item_geom(p, item);
    const int z = it.z0 + tz;
    unsigned okm = 0, imgm = 0;
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int y = it.y0 + ty0 + k;
      if (!padding && (ty0 + k < p.TY) && (y < g.Ny) && (z < g.Nz)) okm |= 1u << k;
      if (yz_image_needed(g, y, z)) imgm |= 1u << k;
    }
    int ic = (int)gidx(g, it.x0 + gx, it.y0 + ty0 + g.gy, 0, z + g.oz);   // g.elems < 2^31 (jb_capi.cu allocate_state)
    unsigned long long gs = global_site(g, it.x0, it.y0 + ty0, 0, z);

    for (int j = 0; j < 2 * gx; ++j) {
      const int s = wrapS(cslotS + j);
      mbar_wait_backoff(smem_u32(&fullS[s]), (phS >> s) & 1u, 40);
      phS ^= 1u << s;
    }

    for (int i = 0; i < it.xc; ++i) {
      bool newest_ready = false;   // has this thread waited for S plane i + 2 gx yet?
      auto wait_newest = [&]() {
        const int s = wrapS(cslotS + 2 * gx);
        mbar_wait_backoff(smem_u32(&fullS[s]), (phS >> s) & 1u, 20);
        phS ^= 1u << s;
        newest_ready = true;
      };
      if (!p.split_wait) wait_newest();
      if (use_u) {
        mbar_wait_backoff(smem_u32(&fullU[cslotU]), (phU >> cslotU) & 1u, 40);
        phU ^= 1u << cslotU;
      }
      const int x = it.x0 + i;
      const bool xb = x_image_needed(g, x);
      const double *uplane = ringU + (size_t)cslotU * 3 * slotU + uoff;

      if (p.debug_skip & 1) { if (!newest_ready) wait_newest(); }   // timing experiments: data movement only
      else
#pragma unroll 1
      for (int m = 0; m < M; ++m) {
        const JbClass &c = p.cls[MOTIF1 ? 0 : m];
        const double *base = ringS + soff + m * p.BZ;
        const double *cplane = base + wrapS(cslotS + gx) * slot3;
        double sx[SPT], sy[SPT], sz[SPT], hx[SPT], hy[SPT], hz[SPT];
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          const double *cp = cplane + k * kS;
          sx[k] = cp[0]; sy[k] = cp[slotS]; sz[k] = cp[2 * slotS];
          hx[k] = c.fTx; hy[k] = c.fTy; hz[k] = c.fTz;   // constant field (Zeeman dc + ac cos wt + applied), Tesla
        }
        // exchange field in Tesla; entries in the reference's CSR column order (interface/sparse_blas.h:22-25)
        const int nb = p.nbr_begin[MOTIF1 ? 0 : m], ne = p.nbr_begin[(MOTIF1 ? 0 : m) + 1];
        const int nsplit = newest_ready ? nb : p.nbr_split[MOTIF1 ? 0 : m];   // entries [nsplit, ne) read the newest plane
        const JbTileNbr *tab = s_nbr + cslotS * p.n_nbr;
        auto gather = [&](int n0, int n1) {
#pragma unroll 2
          for (int n = n0; n < n1; ++n) {
            const int4 raw = *reinterpret_cast<const int4 *>(&tab[n]);   // one LDS.128: {byte offset, d, J}
            struct { int off, d; double J; } e = {raw.x, raw.y, __hiloint2double(raw.w, raw.z)};
            const double *q = reinterpret_cast<const double *>(reinterpret_cast<const char *>(base) + e.off);
            if (ISO) {
#pragma unroll
              for (int k = 0; k < SPT; ++k) {
                hx[k] = fma(e.J, q[k * kS], hx[k]);
                hy[k] = fma(e.J, q[slotS + k * kS], hy[k]);
                hz[k] = fma(e.J, q[2 * slotS + k * kS], hz[k]);
              }
            } else {
              const double *__restrict__ J = p.J9T + 9 * n;
              const double J0 = J[0], J1 = J[1], J2 = J[2], J3 = J[3], J4 = J[4], J5 = J[5], J6 = J[6], J7 = J[7], J8 = J[8];
#pragma unroll
              for (int k = 0; k < SPT; ++k) {
                const double jx = q[k * kS], jy = q[slotS + k * kS], jz = q[2 * slotS + k * kS];
                hx[k] += J0 * jx + J1 * jy + J2 * jz;
                hy[k] += J3 * jx + J4 * jy + J5 * jz;
                hz[k] += J6 * jx + J7 * jy + J8 * jz;
              }
            }
          }
        };
        gather(nb, nsplit);
        if (!newest_ready) wait_newest();
        gather(nsplit, ne);
        // early release: the oldest S plane (at the end of an item: all resident planes) is only read by the gathers
        // above, so its slot can go back to the producer while this warp still does the per-site physics
        if (p.early_release && m == M - 1) {
          __syncwarp();
          if (lane0) {
            mbar_arrive(smem_u32(&emptyS[cslotS]));
            if (i == it.xc - 1) for (int j = 1; j <= 2 * gx; ++j) mbar_arrive(smem_u32(&emptyS[wrapS(cslotS + j)]));
          }
        }
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          if (!((okm >> k) & 1u)) continue;
          double n0 = 0, n1 = 0, n2 = 0;
          if (THERMAL) site_normals_rk(p.rk, p.step, gs + k * kSite + m, n0, n1, n2);
          const int idx = ic + m * g.PZ + k * kG;
          double ux = 0, uy = 0, uz = 0;
          if (STAGE == 1) {
            if (p.u_tma) {
              const double *up = uplane + m * p.UZ + k * kU;
              ux = up[0]; uy = up[slotU]; uz = up[2 * slotU];
            } else {
              ux = p.u[0][idx]; uy = p.u[1][idx]; uz = p.u[2][idx];
            }
          }
          double ox, oy, oz, vx, vy, vz;
          llg_site<STAGE, THERMAL>(c, sx[k], sy[k], sz[k], hx[k], hy[k], hz[k], n0, n1, n2, ux, uy, uz, ox, oy, oz, vx, vy, vz);
          if (p.debug_skip & 2) { if (ox + oy + oz + vx + vy + vz == 1.2345e300) p.out[0][idx] = ox; continue; }   // timing experiments: no stores
          if (p.store_hint == 1) {        // streaming (evict-first) stores: the results are not read again by this launch
            if (STAGE == 0) { __stcs(&p.u[0][idx], vx); __stcs(&p.u[1][idx], vy); __stcs(&p.u[2][idx], vz); }
            __stcs(&p.out[0][idx], ox); __stcs(&p.out[1][idx], oy); __stcs(&p.out[2][idx], oz);
          } else if (p.store_hint == 2) { // write-through
            if (STAGE == 0) { __stwt(&p.u[0][idx], vx); __stwt(&p.u[1][idx], vy); __stwt(&p.u[2][idx], vz); }
            __stwt(&p.out[0][idx], ox); __stwt(&p.out[1][idx], oy); __stwt(&p.out[2][idx], oz);
          } else {
            if (STAGE == 0) { p.u[0][idx] = vx; p.u[1][idx] = vy; p.u[2][idx] = vz; }
            p.out[0][idx] = ox; p.out[1][idx] = oy; p.out[2][idx] = oz;
          }
          if (((imgm >> k) & 1u) | xb) tile_store_images(p, x, it.y0 + ty0 + k, m, z, ox, oy, oz);
        }
      }
      // this warp is done with the oldest S plane (and the u plane): one arrival per warp on their empty barriers
      __syncwarp();
      if (lane0) {
        if (!p.early_release || (p.debug_skip & 1)) {
          mbar_arrive(smem_u32(&emptyS[cslotS]));
          if (i == it.xc - 1) for (int j = 1; j <= 2 * gx; ++j) mbar_arrive(smem_u32(&emptyS[wrapS(cslotS + j)]));
        }
        if (use_u) mbar_arrive(smem_u32(&emptyU[cslotU]));
      }
      cslotS = wrapS(cslotS + 1);
      if (i == it.xc - 1) cslotS = wrapS(cslotS + 2 * gx);
      if (use_u) cslotU = (cslotU + 1 == RU) ? 0 : cslotU + 1;
      ic += (int)g.sX;
      gs += planeSites;
    }
  }
}

template <typename F>
cudaError_t with_kernel(int stage, int thermal, int iso, int spt, int motif1, F &&f) {
#define JB_TILE_CASE(ST, TH, IS, SP, M1) \
  if (stage == ST && thermal == TH && iso == IS && spt == SP && motif1 == M1) \
    return f(stage_tile_kernel<ST, (TH != 0), (IS != 0), SP, (M1 != 0)>);
#define JB_TILE_CASES_SPT(ST, TH, IS) \
  JB_TILE_CASE(ST, TH, IS, 1, 0) JB_TILE_CASE(ST, TH, IS, 2, 0) JB_TILE_CASE(ST, TH, IS, 1, 1) JB_TILE_CASE(ST, TH, IS, 2, 1)
  JB_TILE_CASES_SPT(0, 0, 0) JB_TILE_CASES_SPT(0, 0, 1) JB_TILE_CASES_SPT(0, 1, 0) JB_TILE_CASES_SPT(0, 1, 1)
  JB_TILE_CASES_SPT(1, 0, 0) JB_TILE_CASES_SPT(1, 0, 1) JB_TILE_CASES_SPT(1, 1, 0) JB_TILE_CASES_SPT(1, 1, 1)
#undef JB_TILE_CASES_SPT
#undef JB_TILE_CASE
  return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t jbk_stage_tile_occupancy(const JbTileParams &p, int stage, int thermal, int iso, int spt, int threads,
                                     size_t smem_bytes, int *blocks_per_sm) {
  return with_kernel(stage, thermal, iso, spt, p.g.M == 1 ? 1 : 0, [&](auto k) -> cudaError_t {
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, ((threads + 31) & ~31) + 32, smem_bytes);
  });
}

cudaError_t jbk_stage_tile(const JbTileParams &p, const CUtensorMap *tm, int stage, int thermal, int iso, int spt,
                           int threads, int grid, size_t smem_bytes, cudaStream_t stream) {
  return with_kernel(stage, thermal, iso, spt, p.g.M == 1 ? 1 : 0, [&](auto k) -> cudaError_t {
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (err != cudaSuccess) return err;
    k<<<grid, ((threads + 31) & ~31) + 32, smem_bytes, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
    return cudaGetLastError();
  });
}
