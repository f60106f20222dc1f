// jb_stage_rows.cu — the stage kernel for DEEP exchange templates (BASELINE config 4: bcc Fe with eight neighbour shells, 112
// neighbours per spin, ghost depth 3): one fused LLG-Heun stage, persistent, TMA-fed and warp-specialised like
// jb_stage_pair.cu, but with the field gather organised around REGISTER REUSE, because with ~100 neighbours the stage is bound
// by shared-memory bandwidth (3 x 8 B per neighbour and site = 2.7 kB per site and stage), not by HBM.
//
// Replaces, per stage, the same reference pieces as jb_stage_pair.cu (cusparseSpMV over the 3N x 3N CSR matrix,
// containers/sparse_matrix.h:366-379, the field summation, cuda_heun_llg_kernelA/B); arithmetic follows
// solvers/cpu_llg_heun.cc:45-148.  The summation order of the exchange field differs from the CSR column order
// (interface/sparse_blas.h:22-25) -- rounding only, covered by the 1e-10 trajectory bar.
//
// Structure:
//   * work items, the atomic work queue, the producer warp, the plane ring and the halo handshake are those of the pair kernel
//     (jb_stage_common.cuh).  One CTA per SM with (nearly) all of its shared memory: a ring of 2 gx + 2 planes-with-halo.
//   * a consumer thread owns FOUR sites: the same z and motif site on four consecutive y rows; the lanes of a warp run along z
//     (every shared-memory access of a warp is 256 contiguous bytes: two wavefronts, no bank conflicts).
//   * the host cuts the exchange template of a motif site into SEGMENTS: entries that differ only in their y offset,
//     dy0 ... dy0 + L - 1, L <= 5 (JbRowSeg).  For one segment a thread loads the L + 3 neighbour spins
//     y + dy0 ... y + dy0 + L + 2 of that column once and uses each of them for up to four of its sites:
//         h[s] += sum_t c[t] * v[s + t],   s = 0..3, t = 0..L-1
//     i.e. (L + 3) loads for 4 L neighbour terms instead of 4 L loads.  On the bcc eight-shell template (112 entries per site:
//     40 segments, 224 loads per component for four sites instead of 448) this halves the shared-memory traffic, the bound.
//     Segments are sorted by length and every length has its own fully unrolled loop (no predicates, registers named at compile
//     time).
//   * data flow: recover_u -- the predictor writes only s*, the corrector reads the site's own s_n straight
//     from global memory (coalesced, issued before the gather so that the latency hides behind it), rebuilds the Heun
//     intermediate and writes s_{n+1} in place: 120 B of HBM traffic per spin-update.
//   * isotropic (scalar) couplings only; tensor couplings and non-uniform sites stay with the direct kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "jb_stage_common.cuh"

namespace {

using namespace jbdev;

// the neighbour spins of one segment (three components, L + 3 rows) and its couplings
template <int L>
struct SegRegs {
  double v[L + 3][3];
  double c[L];
};

template <int L>
__device__ __forceinline__ void seg_load(SegRegs<L> &r, uint32_t tab, uint32_t own, int oslot, int R, uint32_t slot8, uint32_t cs8, uint32_t ys8) {
  const int4 hd = lds_entry(tab);   // {delta, d, L, pad}
  int t = oslot + hd.y;
  if (t >= R) t -= R;
  const uint32_t q = own + (uint32_t)t * slot8 + (uint32_t)hd.x;
  if (L >= 1) { const double2 c01 = lds128(tab + 16u); r.c[0] = c01.x; if (L >= 2) r.c[1] = c01.y; }
  if (L >= 3) { const double2 c23 = lds128(tab + 32u); r.c[2] = c23.x; if (L >= 4) r.c[3] = c23.y; }
  if (L >= 5) r.c[4] = lds64(tab + 48u);
#pragma unroll
  for (int j = 0; j < L + 3; ++j) {
    r.v[j][0] = lds64(q + (uint32_t)j * ys8);
    r.v[j][1] = lds64(q + (uint32_t)j * ys8 + cs8);
    r.v[j][2] = lds64(q + (uint32_t)j * ys8 + 2u * cs8);
  }
}

template <int L>
__device__ __forceinline__ void seg_fma(const SegRegs<L> &r, double (&h)[JB_ROWS_Q][3]) {
#pragma unroll
  for (int t = 0; t < L; ++t) {
#pragma unroll
    for (int s = 0; s < JB_ROWS_Q; ++s) {
      h[s][0] = fma(r.c[t], r.v[s + t][0], h[s][0]);
      h[s][1] = fma(r.c[t], r.v[s + t][1], h[s][1]);
      h[s][2] = fma(r.c[t], r.v[s + t][2], h[s][2]);
    }
  }
}

// all segments [b, e) of one length.  Short segments hold few registers and are unrolled further, so that more loads are in
// flight per warp (a CTA has at most eight consumer warps; the second warp of a scheduler fills the remaining gaps).  A variant
// with explicitly double-buffered segments for CTAs of four consumer warps (255 registers) was measured slower (it spilled):
// profiles/README.md r02n.
template <int L>
__device__ __forceinline__ void seg_class(int b, int e, uint32_t tab0, uint32_t own, int oslot, int R, uint32_t slot8, uint32_t cs8, uint32_t ys8,
                                          double (&h)[JB_ROWS_Q][3]) {
  constexpr int U = L <= 2 ? 4 : (L == 3 ? 3 : 2);
#pragma unroll U
  for (int n = b; n < e; ++n) {
    SegRegs<L> A;
    seg_load<L>(A, tab0 + (uint32_t)n * 64u, own, oslot, R, slot8, cs8, ys8);
    seg_fma<L>(A, h);
  }
}

template <int STAGE, bool THERMAL>
__global__ void __launch_bounds__(288, 1) stage_rows_kernel(const __grid_constant__ CUtensorMap tS0,
                                                            const __grid_constant__ CUtensorMap tS1,
                                                            const __grid_constant__ CUtensorMap tS2,
                                                            const __grid_constant__ JbTileParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const JbGeom &g = p.g;
  const int M = g.M, gx = g.gx;
  const int R = p.R;
  const int slotS = p.slotS;
  double *ringS = reinterpret_cast<double *>(smem_raw);
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(ringS + (size_t)R * 3 * slotS);
  unsigned long long *fullS = bars, *emptyS = bars + JB_PAIR_BARS, *fullU = bars + 2 * JB_PAIR_BARS, *emptyU = bars + 3 * JB_PAIR_BARS;
  volatile int *items = reinterpret_cast<volatile int *>(bars + 4 * JB_PAIR_BARS);
  unsigned int *face_arrivals = reinterpret_cast<unsigned int *>(bars + 4 * JB_PAIR_BARS + JB_ITEM_RING / 2);   // [0] lo, [1] hi
  JbRowSeg *s_rows = reinterpret_cast<JbRowSeg *>(bars + JB_STAGE_TAIL_WORDS);

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
  {
    const int4 *src = reinterpret_cast<const int4 *>(p.rows);
    int4 *dst = reinterpret_cast<int4 *>(s_rows);
    for (int idx = tid; idx < p.n_rows * 4; idx += blockDim.x) dst[idx] = src[idx];
  }
  __syncthreads();

  const int warp_idx = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform by construction
  if (warp_idx == n_cw) {
    stage_producer<false>(&tS0, &tS1, &tS2, &tS0, &tS1, &tS2, p, M, ringS, ringS, fullS, emptyS, fullU, emptyU, items);
    return;
  }

  // =========================== consumers: warp = (y quad, motif sub-index), lane = z ===========================
  const int MS = p.msplit;                    // warps per y quad: this one owns the motif sites m0, m0 + MS, ...
  const int yq = warp_idx / MS, m0 = warp_idx - yq * MS;
  const int zl = tid & 31;
  const int ty0 = JB_ROWS_Q * yq;
  const uint32_t cs8 = (uint32_t)slotS * 8u;             // component stride inside a slot, bytes
  const uint32_t slot8 = 3u * cs8;                       // slot stride, bytes
  const uint32_t ys8 = (uint32_t)(M * p.BZ) * 8u;        // y stride inside a slot, bytes
  // the thread's first site (row ty0, m = 0, component x) in slot 0 of the ring
  // (lanes beyond a narrow tile, TZ < 32, compute on column 0 and store nothing: their loads stay inside the slot)
  const uint32_t own = smem_u32(ringS) + (uint32_t)(((ty0 + g.gy) * M) * p.BZ + (zl < p.TZ ? zl : 0) + p.gzb) * 8u;
  const uint32_t tab0 = smem_u32(s_rows);
  const uint32_t fullS0 = smem_u32(fullS), emptyS0 = smem_u32(emptyS);
  const unsigned long long planeSites = (unsigned long long)g.Ny * g.Nz * M;
  const bool lane0 = zl == 0;
  const int sX = (int)g.sX, sY = (int)g.sY;

  int wslot = 0, qi = 0;
  uint32_t wpar = 0u;
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
    if (p.trace && tid == 0 && n_done < JB_TRACE_ITEMS)
      p.trace[(unsigned long long)JB_TRACE_WORDS * blockIdx.x + 4 + n_done] = ((unsigned long long)item << 40) | ((global_timer_ns() - t_first) & 0xffffffffffull);
    ++n_done;
    const ItemGeom it = item_geom(p, item);
    const int z = it.z0 + zl;
    const int y = it.y0 + ty0;                             // the thread's first row
    const bool col = zl < p.TZ && z < g.Nz;
    bool ok[JB_ROWS_Q];
#pragma unroll
    for (int s = 0; s < JB_ROWS_Q; ++s) ok[s] = col && ty0 + s < p.TY && y + s < g.Ny;
    const bool zface = yz_image_needed(g, g.gy, z);        // (a y value that is on no face: tests z only)
    int ic = (int)gidx(g, it.x0 + gx, y + g.gy, 0, z + g.oz);   // g.elems < 2^31 (jb_capi.cu allocate_state)
    unsigned long long gs = global_site(g, it.x0, y, 0, z & ~1);   // the noise key of the first row's z pair, m = 0
    const bool face_lo = it.x0 < gx, face_hi = it.x0 + it.xc > g.nx - gx;

    for (int j = 1; j < 2 * gx; ++j) {
      mbar_wait(fullS0 + 8u * wslot, wpar);
      if (++wslot == R) { wslot = 0; wpar ^= 1u; }
    }
    int oslot = oslot0;                                    // slot of the oldest resident plane (x - gx)
    int cslot = oslot0 + gx; if (cslot >= R) cslot -= R;   // slot of the centre plane

    for (int i = 0; i < it.xc; ++i) {
      const int x = it.x0 + i;
      const bool xb = x_image_needed(g, x);
#pragma unroll 1
      for (int m = m0; m < M; m += MS) {
        const JbClass &c = p.cls[m];
        const int idx = ic + m * g.PZ;
        // corrector: the site's own s_n, straight from global memory (the only read of S0 in this stage; the same thread
        // overwrites it below).  Issued before the gather: the latency hides behind ~1000 instructions.
        double sn[JB_ROWS_Q][3];
        if (STAGE == 1) {
#pragma unroll
          for (int s = 0; s < JB_ROWS_Q; ++s) {
            sn[s][0] = sn[s][1] = sn[s][2] = 0.0;
            if (ok[s]) { sn[s][0] = p.out[0][idx + s * sY]; sn[s][1] = p.out[1][idx + s * sY]; sn[s][2] = p.out[2][idx + s * sY]; }
          }
        }
        if (m == m0 && (gx > 0 || i > 0)) {
          mbar_wait(fullS0 + 8u * wslot, wpar);
          if (++wslot == R) { wslot = 0; wpar ^= 1u; }
        }
        double h[JB_ROWS_Q][3];
#pragma unroll
        for (int s = 0; s < JB_ROWS_Q; ++s) { h[s][0] = c.fTx; h[s][1] = c.fTy; h[s][2] = c.fTz; }   // constant field (Zeeman dc + ac cos wt + applied), Tesla
        const uint32_t ownm = own + (uint32_t)(m * p.BZ) * 8u;
        const int *rb = p.row_begin[m];
        seg_class<5>(rb[4], rb[5], tab0, ownm, oslot, R, slot8, cs8, ys8, h);
        seg_class<4>(rb[3], rb[4], tab0, ownm, oslot, R, slot8, cs8, ys8, h);
        seg_class<3>(rb[2], rb[3], tab0, ownm, oslot, R, slot8, cs8, ys8, h);
        seg_class<2>(rb[1], rb[2], tab0, ownm, oslot, R, slot8, cs8, ys8, h);
        seg_class<1>(rb[0], rb[1], tab0, ownm, oslot, R, slot8, cs8, ys8, h);
        // the thread's own spins (centre plane)
        double sc[JB_ROWS_Q][3];
        {
          const uint32_t a = ownm + (uint32_t)cslot * slot8;
#pragma unroll
          for (int s = 0; s < JB_ROWS_Q; ++s) {
            sc[s][0] = lds64(a + (uint32_t)s * ys8); sc[s][1] = lds64(a + (uint32_t)s * ys8 + cs8); sc[s][2] = lds64(a + (uint32_t)s * ys8 + 2u * cs8);
          }
        }
        // release: the oldest S plane (at the end of an item: all resident planes) goes back to the producer while this warp
        // does the per-site physics
        if (m + MS >= M) {
          __syncwarp();
          if (lane0) {
            mbar_arrive(emptyS0 + 8u * oslot);
            if (i == it.xc - 1) {
              int s = oslot;
              for (int j = 1; j <= 2 * gx; ++j) { if (++s == R) s = 0; mbar_arrive(emptyS0 + 8u * s); }
            }
          }
        }
#pragma unroll
        for (int s = 0; s < JB_ROWS_Q; ++s) {
          double ux = 0.0, uy = 0.0, uz = 0.0;
          if (STAGE == 1) {   // rebuild u = (s_n + lambda s*) / 2
            ux = sn[s][0]; uy = sn[s][1]; uz = sn[s][2];
            recover_u(sc[s][0], sc[s][1], sc[s][2], ux, uy, uz);
          }
          double n0 = 0.0, n1 = 0.0, n2 = 0.0;
          if (THERMAL) site_normals_rk(p.rk, p.step, gs + (unsigned long long)s * g.Nz * M + m, (z & 1) != 0, n0, n1, n2);
          double ox, oy, oz, vx, vy, vz;
          llg_site<STAGE, THERMAL, false>(c, sc[s][0], sc[s][1], sc[s][2], h[s][0], h[s][1], h[s][2], n0, n1, n2, ux, uy, uz, ox, oy, oz, vx, vy, vz);
          if (ok[s]) {
            const int o = idx + s * sY;
            p.out[0][o] = ox; p.out[1][o] = oy; p.out[2][o] = oz;
            // ghost images: rows and columns within a ghost depth of a face (a few per cent of the sites of a deep template's
            // lattice) take the general routine
            const int ys = y + s;
            if (xb | zface | (g.per[1] && ((ys < g.gy) | (ys >= g.Ny - g.gy)))) tile_store_images(p, x, ys, m, z, ox, oy, oz);
          }
        }
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
cudaError_t with_kernel(int stage, int thermal, F &&f) {
  if (stage == 0 && !thermal) return f(stage_rows_kernel<0, false>);
  if (stage == 0 && thermal) return f(stage_rows_kernel<0, true>);
  if (stage == 1 && !thermal) return f(stage_rows_kernel<1, false>);
  if (stage == 1 && thermal) return f(stage_rows_kernel<1, true>);
  return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t jbk_stage_rows_occupancy(int stage, int thermal, int threads, size_t smem_bytes, int *blocks_per_sm) {
  return with_kernel(stage, thermal, [&](auto k) -> cudaError_t {
    cudaError_t err = jb_ensure_dynamic_smem(k, smem_bytes);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, threads + 32, smem_bytes);
  });
}

cudaError_t jbk_stage_rows(const JbTileParams &p, const CUtensorMap *tm, int stage, int thermal, int threads, int grid,
                           size_t smem_bytes, cudaStream_t stream) {
  return with_kernel(stage, thermal, [&](auto k) -> cudaError_t {
    cudaError_t err = jb_ensure_dynamic_smem(k, smem_bytes);
    if (err != cudaSuccess) return err;
    k<<<grid, threads + 32, smem_bytes, stream>>>(tm[0], tm[1], tm[2], p);
    return cudaGetLastError();
  });
}
