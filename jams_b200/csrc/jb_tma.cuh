// jb_tma.cuh — sm_100a building blocks of the persistent stage kernel (jb_stage_pair.cu): mbarrier full/empty handshakes, 3-D TMA loads (cp.async.bulk.tensor), shared-memory accesses by
// 32-bit shared address (always LDS / STS, never a generic LD), 16-byte global stores and the work-item geometry.
#ifndef JB_TMA_CUH
#define JB_TMA_CUH

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "jb_device.cuh"

namespace jbdev {

__device__ __forceinline__ uint32_t smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// wait with a suspend-time hint: the hardware parks the warp until the phase completes (or the hint expires) instead of
// returning after a short default time-out.  A spinning try_wait + branch pair was 18 % of all issued instructions in the
// T = 100 K profile of the pair kernel (profiles/README.md r01g).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "JB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra JB_DONE;\n\t"
      "bra JB_WAIT;\n\t"
      "JB_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ double2 lds128(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ int4 lds_entry(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void stg128(double *ptr, double a, double b) {
  asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(ptr), "d"(a), "d"(b) : "memory");
}
// ghost images of a boundary site, general case (x / y faces and their edges; rare, out of line).  The parameter block
// is __grid_constant__, so its address can be handed over without a local copy: no stack frame in the kernels.
static __device__ __noinline__ void tile_store_images(const JbTileParams &p, int x, int y, int m, int z, double vx, double vy, double vz) {
  JbOutBoxes boxes;
#pragma unroll
  for (int c = 0; c < 3; ++c) { boxes.out[c] = p.out[c]; boxes.out_lo[c] = p.out_lo[c]; boxes.out_hi[c] = p.out_hi[c]; }
  store_images_inline(p.g, boxes, x, y, m, z, vx, vy, vz);
}

// work item = (x-chunk, yz-column tile); the chunk plan lives in the parameter bank (jb_capi.cu plan_chunks)
struct ItemGeom { int y0, z0, x0, xc; };

__device__ __forceinline__ ItemGeom item_geom(const JbTileParams &p, int item) {
  ItemGeom it;
  const int chunk = item / p.n_cols, col = item - chunk * p.n_cols;
  const int yt = col / p.n_zt, zt = col - yt * p.n_zt;
  it.y0 = yt * p.TY; it.z0 = zt * p.TZ;
  it.x0 = p.chunk_x0[chunk];
  it.xc = p.chunk_xc[chunk];
  return it;
}

}  // namespace jbdev

#endif  // JB_TMA_CUH
