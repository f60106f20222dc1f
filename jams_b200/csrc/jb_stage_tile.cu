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
//  * Each thread owns SPT y-sites x M motif sites of one z column; the exchange template, coupling constants
//    and material classes are read from the kernel-parameter bank with uniform indices (no per-thread loads).
//  * Results go straight from registers to global memory (256 B per warp and component); sites within a ghost
//    depth of a face also store their periodic / neighbour-slab images (peer memory over NVLink).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "jb_device.cuh"

namespace {

using namespace jbdev;

__device__ __forceinline__ uint32_t smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

struct ItemGeom { int y0, z0, x0, xc; };

__device__ __forceinline__ ItemGeom item_geom(const JbTileParams &p, int item) {
  ItemGeom it;
  const int chunk = item / p.n_cols, col = item - chunk * p.n_cols;
  const int yt = col / p.n_zt, zt = col - yt * p.n_zt;
  it.y0 = yt * p.TY; it.z0 = zt * p.TZ;
  it.x0 = (int)(((long long)chunk * p.g.nx) / p.n_chunks);
  it.xc = (int)(((long long)(chunk + 1) * p.g.nx) / p.n_chunks) - it.x0;
  return it;
}

#define JB_TILE_BARS 8

template <int STAGE, bool THERMAL, bool ISO, int SPT>
__global__ void __launch_bounds__(512) stage_tile_kernel(const __grid_constant__ CUtensorMap tS0,
                                                         const __grid_constant__ CUtensorMap tS1,
                                                         const __grid_constant__ CUtensorMap tS2,
                                                         const __grid_constant__ CUtensorMap tU0,
                                                         const __grid_constant__ CUtensorMap tU1,
                                                         const __grid_constant__ CUtensorMap tU2,
                                                         const __grid_constant__ JbTileParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const JbGeom &g = p.g;
  const int M = g.M, gx = g.gx, nd = 2 * g.gx + 1;
  const int R = p.R, RU = p.RU;
  const int slotS = p.slotS, slotU = p.slotU;
  double *ringS = reinterpret_cast<double *>(smem_raw);
  double *ringU = ringS + (size_t)R * 3 * slotS;
  unsigned long long *barS = reinterpret_cast<unsigned long long *>(ringU + ((STAGE == 1 && p.u_tma) ? (size_t)RU * 3 * slotU : 0));
  unsigned long long *barU = barS + JB_TILE_BARS;
  JbTileNbr *s_nbr = reinterpret_cast<JbTileNbr *>(barU + JB_TILE_BARS);

  const int tid = threadIdx.x;
  const int G = gridDim.x, bid = blockIdx.x;

  if (tid == 0) {
    for (int s = 0; s < R; ++s) mbar_init(smem_u32(&barS[s]), 1);
    if (STAGE == 1) for (int s = 0; s < JB_TILE_BARS; ++s) mbar_init(smem_u32(&barU[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int n = tid; n < p.n_nbr; n += blockDim.x) s_nbr[n] = p.nbr[n];
  __syncthreads();

  // ---- producer (thread 0): the stream of S planes (and U planes) of this CTA's items ------------------
  const uint32_t bytesS = (uint32_t)(p.BY * M * p.BZ * sizeof(double));
  const uint32_t bytesU = (uint32_t)(p.TY * M * p.UZ * sizeof(double));   // UZ = BZ: see jb_capi.cu choose_tiling
  int ps_item = bid, ps_j = 0, ps_slot = 0, issuedS = 0;
  int pu_item = bid, pu_i = 0, pu_slot = 0, issuedU = 0;
  ItemGeom ps_g = item_geom(p, bid < p.n_items ? bid : 0), pu_g = ps_g;
  int freedS = 0, freedU = 0;

  auto produce = [&]() {
    while (issuedS < freedS + R && ps_item < p.n_items) {
      const uint32_t bar = smem_u32(&barS[ps_slot]);
      double *dst = ringS + (size_t)ps_slot * 3 * slotS;
      mbar_expect_tx(bar, 3 * bytesS);
      tma_load_3d(smem_u32(dst), &tS0, ps_g.z0, ps_g.y0 * M, ps_g.x0 + ps_j, bar);
      tma_load_3d(smem_u32(dst + slotS), &tS1, ps_g.z0, ps_g.y0 * M, ps_g.x0 + ps_j, bar);
      tma_load_3d(smem_u32(dst + 2 * slotS), &tS2, ps_g.z0, ps_g.y0 * M, ps_g.x0 + ps_j, bar);
      ps_slot = (ps_slot + 1 == R) ? 0 : ps_slot + 1;
      ++issuedS;
      if (++ps_j == ps_g.xc + 2 * gx) {
        ps_j = 0; ps_item += G;
        if (ps_item < p.n_items) ps_g = item_geom(p, ps_item);
      }
    }
    if (STAGE == 1 && p.u_tma) {
      while (issuedU < freedU + RU && pu_item < p.n_items) {
        const uint32_t bar = smem_u32(&barU[pu_slot]);
        double *dst = ringU + (size_t)pu_slot * 3 * slotU;
        mbar_expect_tx(bar, 3 * bytesU);
        const int c0 = pu_g.z0, c1 = (pu_g.y0 + g.gy) * M, c2 = pu_g.x0 + pu_i + gx;  // inner start kept 16-byte aligned
        tma_load_3d(smem_u32(dst), &tU0, c0, c1, c2, bar);
        tma_load_3d(smem_u32(dst + slotU), &tU1, c0, c1, c2, bar);
        tma_load_3d(smem_u32(dst + 2 * slotU), &tU2, c0, c1, c2, bar);
        pu_slot = (pu_slot + 1 == RU) ? 0 : pu_slot + 1;
        ++issuedU;
        if (++pu_i == pu_g.xc) {
          pu_i = 0; pu_item += G;
          if (pu_item < p.n_items) pu_g = item_geom(p, pu_item);
        }
      }
    }
  };
  if (tid == 0) produce();

  // ---- consumer: every thread owns SPT y-sites x M motif sites of one z column of the tile ---------------
  const int tz = tid % p.TZ, tyg = tid / p.TZ;
  const int ty0 = tyg * SPT;
  const int soff = ((ty0 + g.gy) * M) * p.BZ + tz + g.gz;   // centre of (k = 0, m = 0) inside a slot component
  const int uoff = (ty0 * M) * p.UZ + tz + g.gz;
  const int kS = M * p.BZ, kU = M * p.UZ, kG = M * g.PZ;    // strides between the thread's consecutive y sites
  const unsigned long long kSite = (unsigned long long)g.Nz * M;
  const unsigned long long planeSites = (unsigned long long)g.Ny * g.Nz * M;

  JbOutBoxes boxes;
#pragma unroll
  for (int c = 0; c < 3; ++c) { boxes.out[c] = p.out[c]; boxes.out_lo[c] = p.out_lo[c]; boxes.out_hi[c] = p.out_hi[c]; }

  int cslotS = 0, cslotU = 0;
  uint32_t phS = 0, phU = 0;
  auto wrapS = [&](int a) { return a >= R ? a - R : a; };

  for (int item = bid; item < p.n_items; item += G) {
    const ItemGeom it = item_geom(p, item);
    const int z = it.z0 + tz;
    bool ok[SPT], img[SPT];
#pragma unroll
    for (int k = 0; k < SPT; ++k) {
      const int y = it.y0 + ty0 + k;
      ok[k] = (ty0 + k < p.TY) && (y < g.Ny) && (z < g.Nz);
      img[k] = yz_image_needed(g, y, z);
    }
    int ic = (int)gidx(g, it.x0 + gx, it.y0 + ty0 + g.gy, 0, z + g.gz);   // g.elems < 2^31 (jb_capi.cu allocate_state)
    unsigned long long gs = global_site(g, it.x0, it.y0 + ty0, 0, z);

    for (int j = 0; j < 2 * gx; ++j) {
      const int s = wrapS(cslotS + j);
      mbar_wait(smem_u32(&barS[s]), (phS >> s) & 1u);
      phS ^= 1u << s;
    }

    for (int i = 0; i < it.xc; ++i) {
      {
        const int s = wrapS(cslotS + 2 * gx);
        mbar_wait(smem_u32(&barS[s]), (phS >> s) & 1u);
        phS ^= 1u << s;
      }
      if (STAGE == 1 && p.u_tma) {
        mbar_wait(smem_u32(&barU[cslotU]), (phU >> cslotU) & 1u);
        phU ^= 1u << cslotU;
      }
      const int x = it.x0 + i;
      const bool xb = x_image_needed(g, x);
      const double *cplane = ringS + (size_t)wrapS(cslotS + gx) * 3 * slotS + soff;
      const double *uplane = ringU + (size_t)cslotU * 3 * slotU + uoff;

      for (int m = 0; m < M; ++m) {
        double sx[SPT], sy[SPT], sz[SPT], hx[SPT], hy[SPT], hz[SPT];
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          const double *cp = cplane + m * p.BZ + k * kS;
          sx[k] = cp[0]; sy[k] = cp[slotS]; sz[k] = cp[2 * slotS];
          hx[k] = 0.0; hy[k] = 0.0; hz[k] = 0.0;
        }
        // exchange field: entries are grouped by dx, inside a group in the reference's CSR column order
        // (ascending neighbour site id, interface/sparse_blas.h:22-25)
        for (int d = 0; d < nd; ++d) {
          const double *pl = ringS + (size_t)wrapS(cslotS + d) * 3 * slotS + soff + m * p.BZ;
          const int nb = p.nbr_begin[m * nd + d], ne = p.nbr_begin[m * nd + d + 1];
          for (int n = nb; n < ne; ++n) {
            const JbTileNbr e = s_nbr[n];
            const double *q = pl + e.delta;
            if (ISO) {
              const double J = e.J;
#pragma unroll
              for (int k = 0; k < SPT; ++k) {
                hx[k] = fma(J, q[k * kS], hx[k]);
                hy[k] = fma(J, q[slotS + k * kS], hy[k]);
                hz[k] = fma(J, q[2 * slotS + k * kS], hz[k]);
              }
            } else {
              const double *__restrict__ J = p.Jtab + 9 * e.jidx;
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
        }
        const JbClass &c = p.cls[p.class_of_motif[m]];
#pragma unroll
        for (int k = 0; k < SPT; ++k) {
          if (!ok[k]) continue;
          double n0 = 0, n1 = 0, n2 = 0;
          if (THERMAL) site_normals(p.seed, p.step, gs + k * kSite + m, n0, n1, n2);
          double ux = 0, uy = 0, uz = 0;
          const int idx = ic + m * g.PZ + k * kG;
          if (STAGE == 1) {
            if (p.u_tma) {
              const double *up = uplane + m * p.UZ + k * kU;
              ux = up[0]; uy = up[slotU]; uz = up[2 * slotU];
            } else {
              ux = p.u[0][idx]; uy = p.u[1][idx]; uz = p.u[2][idx];
            }
          }
          double ox, oy, oz, vx, vy, vz;
          llg_site<STAGE, THERMAL>(c, sx[k], sy[k], sz[k], hx[k], hy[k], hz[k], n0, n1, n2, p.dt, p.half_dt, ux, uy, uz,
                                   ox, oy, oz, vx, vy, vz);
          if (STAGE == 0) { p.u[0][idx] = vx; p.u[1][idx] = vy; p.u[2][idx] = vz; }
          p.out[0][idx] = ox; p.out[1][idx] = oy; p.out[2][idx] = oz;
          if (img[k] | xb) store_images(g, boxes, x, it.y0 + ty0 + k, m, z, ox, oy, oz);
        }
      }
      __syncthreads();  // every thread is done with the oldest S plane and the U plane: their slots can be refilled
      cslotS = wrapS(cslotS + 1);
      ++freedS;
      if (i == it.xc - 1) { cslotS = wrapS(cslotS + 2 * gx); freedS += 2 * gx; }
      if (STAGE == 1) { cslotU = (cslotU + 1 == RU) ? 0 : cslotU + 1; ++freedU; }
      ic += (int)g.sX;
      gs += planeSites;
      if (tid == 0) produce();
    }
  }
}

template <typename F>
cudaError_t with_kernel(int stage, int thermal, int iso, int spt, F &&f) {
#define JB_TILE_CASE(ST, TH, IS, SP) \
  if (stage == ST && thermal == TH && iso == IS && spt == SP) return f(stage_tile_kernel<ST, (TH != 0), (IS != 0), SP>);
#define JB_TILE_CASES_SPT(ST, TH, IS) JB_TILE_CASE(ST, TH, IS, 1) JB_TILE_CASE(ST, TH, IS, 2) JB_TILE_CASE(ST, TH, IS, 4)
  JB_TILE_CASES_SPT(0, 0, 0) JB_TILE_CASES_SPT(0, 0, 1) JB_TILE_CASES_SPT(0, 1, 0) JB_TILE_CASES_SPT(0, 1, 1)
  JB_TILE_CASES_SPT(1, 0, 0) JB_TILE_CASES_SPT(1, 0, 1) JB_TILE_CASES_SPT(1, 1, 0) JB_TILE_CASES_SPT(1, 1, 1)
#undef JB_TILE_CASES_SPT
#undef JB_TILE_CASE
  return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t jbk_stage_tile_smem_bytes(const JbTileParams &p, int stage, size_t *bytes) {
  *bytes = ((size_t)p.R * 3 * p.slotS + ((stage == 1 && p.u_tma) ? (size_t)p.RU * 3 * p.slotU : 0)) * sizeof(double) +
           2 * JB_TILE_BARS * sizeof(unsigned long long) + (size_t)p.n_nbr * sizeof(JbTileNbr) + 128;
  return cudaSuccess;
}

cudaError_t jbk_stage_tile_occupancy(const JbTileParams &p, int stage, int thermal, int iso, int spt, int threads,
                                     size_t smem_bytes, int *blocks_per_sm) {
  (void)p;
  return with_kernel(stage, thermal, iso, spt, [&](auto k) -> cudaError_t {
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, threads, smem_bytes);
  });
}

cudaError_t jbk_stage_tile(const JbTileParams &p, const CUtensorMap *tm, int stage, int thermal, int iso, int spt,
                           int threads, int grid, size_t smem_bytes, cudaStream_t stream) {
  return with_kernel(stage, thermal, iso, spt, [&](auto k) -> cudaError_t {
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (err != cudaSuccess) return err;
    k<<<grid, threads, smem_bytes, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
    return cudaGetLastError();
  });
}
