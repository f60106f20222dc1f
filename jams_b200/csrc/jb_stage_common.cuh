// jb_stage_common.cuh — what the persistent TMA stage kernels (jb_stage_pair.cu, jb_stage_rows.cu) share: the producer warp
// (atomic work queue -> item ring -> stream of planes-with-halo through 3-D TMA boxes into the shared-memory ring, full / empty
// mbarriers per slot) and the in-kernel epoch handshake of slab-decomposed runs.
#ifndef JB_STAGE_COMMON_CUH
#define JB_STAGE_COMMON_CUH

#include "jb_tma.cuh"

#define JB_PAIR_BARS JB_PAIR_MAX_RING
// shared memory after the rings: 4 x JB_PAIR_BARS mbarriers, the item ring (JB_ITEM_RING ints), two face-arrival counters + pad
#define JB_STAGE_TAIL_WORDS (4 * JB_PAIR_BARS + JB_ITEM_RING / 2 + 2)   // in 8-byte words; the tables (16-byte aligned) follow

#include <mutex>
#include <unordered_map>

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a driver call of several microseconds: once per kernel, device and size
// instead of once per launch (the stage launches of small lattices are host-bound: BASELINE config 1 runs 11 us kernels)
template <typename K>
inline cudaError_t jb_ensure_dynamic_smem(K kernel, size_t bytes) {
  static std::mutex mu;
  static std::unordered_map<unsigned long long, size_t> done;   // (function, device) -> largest size set so far
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long key = (unsigned long long)reinterpret_cast<uintptr_t>(reinterpret_cast<const void *>(kernel)) * 64ull + (unsigned long long)dev;
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return cudaSuccess;
  const cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (err == cudaSuccess) done[key] = bytes;
  return err;
}

namespace jbdev {

static __device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// wait until a neighbour has published `epoch` in this rank's flag (a dead peer must not hang the GPU: after 10 s the
// error flag is raised and the kernel carries on; jb_synchronize reports it)
static __device__ __noinline__ void halo_poll(unsigned long long *flags, int side, unsigned long long epoch) {
  const unsigned long long t0 = global_timer_ns();
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + side) : "memory");
    if (v >= epoch) break;
    if (global_timer_ns() - t0 > 10000000000ull) { flags[2] = 1ull; break; }
    __nanosleep(64);
  }
  asm volatile("fence.proxy.async;" ::: "memory");   // the TMA engine (async proxy) reads what the neighbour's generic stores wrote
}

// a consumer warp has finished a face item of `side`.  The warps of a CTA first meet at a counter in shared memory (acq_rel at
// CTA scope: whoever completes a multiple of n_cw arrivals has observed the stores of all earlier arrivals), and only that
// warp pays for the system-scope fence and the global counter; the last CTA-level arrival of the launch tells the neighbour.
// (Round 2 first did fence + global atomic per WARP: 8 x 256 system fences per stage, face items 22 % slower per plane than
// interior items, profiles/r02q_mgpu_trace_2gpu.log.)  Warps of one CTA may be on different face items at the same time; the
// n_cw-th, 2 n_cw-th, ... arrival each stand for one finished item, and the CTA's final arrival is ordered after all of them.
static __device__ __noinline__ void halo_face_done(const JbHalo &h, int side, uint32_t cta_counter, unsigned int n_cw) {
  unsigned int seen;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(seen) : "r"(cta_counter + 4u * side) : "memory");
  if ((seen + 1u) % n_cw != 0u) return;
  __threadfence_system();
  const unsigned int old = atomicAdd(h.face_count + side, 1u);
  if (old + 1u == h.face_target[side]) {
    h.face_count[side] = 0u;   // every other CTA-level arrival has been counted: ready for the next launch (stream order)
    __threadfence_system();
    unsigned long long *dst = side == 0 ? h.sig_lo : h.sig_hi;
    if (dst) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(h.signal_epoch) : "memory");
  }
}

// The producer warp of a CTA: one elected thread draws item ids from the queue (one item ahead, so the fetch latency hides behind
// the last planes of the current item), publishes them to the consumers through the item ring and streams the planes of the
// item, x0 - gx ... x0 + xc + gx - 1, into the S ring; HAS_U: also the tile's own plane of the second tensor (the Heun
// intermediate, or s_n with recover_u) into the U ring.  Returns when the queue is empty.
template <bool HAS_U>
__device__ __forceinline__ void stage_producer(const CUtensorMap *tS0, const CUtensorMap *tS1, const CUtensorMap *tS2,
                                               const CUtensorMap *tU0, const CUtensorMap *tU1, const CUtensorMap *tU2,
                                               const JbTileParams &p, const int M, double *ringS, double *ringU,
                                               unsigned long long *fullS, unsigned long long *emptyS, unsigned long long *fullU,
                                               unsigned long long *emptyU, volatile int *items) {
  const JbGeom &g = p.g;
  const int gx = g.gx, R = p.R, RU = p.RU, slotS = p.slotS, slotU = p.slotU;
  uint32_t elected = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(elected));
  if (!elected) return;
  const uint32_t bytesS = (uint32_t)(p.BY * M * p.BZ * sizeof(double));
  const uint32_t bytesU = (uint32_t)(p.TY * M * p.UZ * sizeof(double));
  int slot = 0, uslot = 0, qi = 0;
  uint32_t epar = 1u, upar = 1u;   // parity to wait for on the empty barriers (a fresh barrier passes parity 1 at once)
  bool polled_lo = false, polled_hi = false;
  int next = (int)atomicAdd(p.queue, 1u);
  for (;;) {
    const int item = next;
    items[qi] = item < p.n_items ? item : -1;
    qi = (qi + 1) & (JB_ITEM_RING - 1);
    if (item >= p.n_items) {
      // no more work: complete the phase the consumers wait on for "the first plane of the next item" without a transfer
      mbar_wait(smem_u32(&emptyS[slot]), epar);
      mbar_arrive(smem_u32(&fullS[slot]));
      break;
    }
    const ItemGeom it = item_geom(p, item);
    const int np = it.xc + 2 * gx;
    const int zs = it.z0 + g.oz - p.gzb;   // first column of the spin box: even, i.e. 16-byte aligned (TMA requirement)
    const int fetch_at = np > 4 ? np - 4 : 0;
    for (int j = 0; j < np; ++j) {
      if (j == fetch_at) next = (int)atomicAdd(p.queue, 1u);
      if (p.halo.enabled) {   // ghost planes are written by the neighbours' previous stage
        const int xl = it.x0 - gx + j;
        if (xl < 0 && (p.halo.enabled & 1) && !polled_lo) { halo_poll(p.halo.flags, 0, p.halo.wait_epoch); polled_lo = true; }
        if (xl >= g.nx && (p.halo.enabled & 2) && !polled_hi) { halo_poll(p.halo.flags, 1, p.halo.wait_epoch); polled_hi = true; }
      }
      {
        mbar_wait(smem_u32(&emptyS[slot]), epar);
        const uint32_t bar = smem_u32(&fullS[slot]);
        double *dst = ringS + (size_t)slot * 3 * slotS;
        mbar_expect_tx(bar, 3 * bytesS);
        tma_load_3d(smem_u32(dst), tS0, zs, it.y0 * M, it.x0 + j, bar);
        tma_load_3d(smem_u32(dst + slotS), tS1, zs, it.y0 * M, it.x0 + j, bar);
        tma_load_3d(smem_u32(dst + 2 * slotS), tS2, zs, it.y0 * M, it.x0 + j, bar);
        if (++slot == R) { slot = 0; epar ^= 1u; }
      }
      if (HAS_U && j >= 2 * gx) {   // the u (RECU: s_n) plane of step i = j - 2 gx is needed together with S plane j
        mbar_wait(smem_u32(&emptyU[uslot]), upar);
        const uint32_t bar = smem_u32(&fullU[uslot]);
        double *dst = ringU + (size_t)uslot * 3 * slotU;
        mbar_expect_tx(bar, 3 * bytesU);
        const int c0 = it.z0 + g.oz, c1 = (it.y0 + g.gy) * M, c2 = it.x0 + (j - 2 * gx) + gx;   // oz, z0 even: aligned
        tma_load_3d(smem_u32(dst), tU0, c0, c1, c2, bar);
        tma_load_3d(smem_u32(dst + slotU), tU1, c0, c1, c2, bar);
        tma_load_3d(smem_u32(dst + 2 * slotU), tU2, c0, c1, c2, bar);
        if (++uslot == RU) { uslot = 0; upar ^= 1u; }
      }
    }
  }
}

}  // namespace jbdev

#endif  // JB_STAGE_COMMON_CUH
