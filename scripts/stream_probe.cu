// stream_probe.cu — development probe (not part of the product): what HBM bandwidth does a plain streaming kernel reach
// on this GPU for the read:write mixes of the two Heun stages (A: 3 arrays read, 6 written; B: 6 read, 3 written)
// and for a 1:1 copy, with 16-byte accesses over arrays of the C3 size?  Gives the practical ceiling the stage kernels
// are compared with in profiles/README.md next to the driver's MEASURED_PEAKS.json copy number.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/stream_probe.cu -o scripts/stream_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int NR, int NW>
__global__ void stream_kernel(const double2 *__restrict__ const *in, double2 *const *out, size_t n2) {
  const double2 *ip[NR > 0 ? NR : 1]; double2 *op[NW > 0 ? NW : 1];
#pragma unroll
  for (int k = 0; k < NR; ++k) ip[k] = in[k];
#pragma unroll
  for (int k = 0; k < NW; ++k) op[k] = out[k];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 acc = make_double2(0, 0);
#pragma unroll
    for (int k = 0; k < NR; ++k) { const double2 v = ip[k][i]; acc.x += v.x; acc.y += v.y; }
#pragma unroll
    for (int k = 0; k < NW; ++k) { op[k][i] = make_double2(acc.x + k, acc.y - k); }
    if (NW == 0 && acc.x == 1.2345e300) out[0][i] = acc;   // keep the loads alive in the read-only case
  }
}

template <int NR, int NW>
void run(const char *label, double2 **d_in, double2 **d_out, size_t n2, int grid, int block) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) stream_kernel<NR, NW><<<grid, block>>>(d_in, d_out, n2);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int r = 0; r < reps; ++r) stream_kernel<NR, NW><<<grid, block>>>(d_in, d_out, n2);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  const double bytes = (double)(NR + NW) * n2 * 16.0;
  printf("%-28s grid %5d x %4d: %.3f ms  %.0f GB/s  (%s)\n", label, grid, block, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const size_t n = 258ull * 258 * 288;   // one ghosted component of sc 256^3
  const size_t n2 = n / 2;
  double2 *h_ptr[9]; double2 **d_in, **d_out;
  for (int k = 0; k < 9; ++k) { cudaMalloc(&h_ptr[k], n * 8); cudaMemset(h_ptr[k], 0, n * 8); }
  // inputs: arrays 0..5, outputs: arrays 3..8 (A: read 0-2 write 3-8; B: read 0-5 write 6-8)
  cudaMalloc(&d_in, 9 * sizeof(void *)); cudaMalloc(&d_out, 9 * sizeof(void *));
  cudaMemcpy(d_in, h_ptr, 9 * sizeof(void *), cudaMemcpyHostToDevice);
  for (int grid : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
    for (int block : {256, 512}) {
      cudaMemcpy(d_out, h_ptr + 3, 6 * sizeof(void *), cudaMemcpyHostToDevice);
      run<3, 6>("stage A mix (3R 6W)", d_in, d_out, n2, grid, block);
      cudaMemcpy(d_out, h_ptr + 6, 3 * sizeof(void *), cudaMemcpyHostToDevice);
      run<6, 3>("stage B mix (6R 3W)", d_in, d_out, n2, grid, block);
      cudaMemcpy(d_out, h_ptr + 3, 3 * sizeof(void *), cudaMemcpyHostToDevice);
      run<3, 3>("copy mix (3R 3W)", d_in, d_out, n2, grid, block);
      run<6, 0>("read only (6R)", d_in, d_out, n2, grid, block);
      cudaMemcpy(d_out, h_ptr + 3, 6 * sizeof(void *), cudaMemcpyHostToDevice);
      run<0, 6>("write only (6W)", d_in, d_out, n2, grid, block);
    }
  }
  return 0;
}
