// Microbenchmark: tcgen05.ld bandwidth per SM (bytes/clk) vs number of warps and vector width.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../ullsam_b200/csrc tmem_ld_bw.cu -o tmem_ld_bw
#include <cstdio>
#include "common.cuh"
using namespace ullsam;

template <int WIDTH>
__global__ void k(int iters, long long* cycles, unsigned* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if constexpr (WIDTH == 32) {
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < 4; ++c) { tmem_ld32(base + ((warp >> 2) & 3) * 128 + c * 32, v); tmem_wait_ld(); acc += v[0] + v[31]; }
    } else {
      uint32_t v[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(base + ((warp >> 2) & 3) * 128 + c * 32, v[c]);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 4; ++c) acc += v[c][0] + v[c][31];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(tptr, 512);
}

int main() {
  long long* cyc; unsigned* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 16}) {
      if (mode == 0) k<32><<<148, warps * 32>>>(iters, cyc, sink); else k<128><<<148, warps * 32>>>(iters, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = double(iters) * 4 * 32 * 32 * 4 * warps;  // per SM
      printf("%s warps=%2d: %s  %.1f bytes/clk/SM (cycles %lld)\n", mode ? "4 loads then wait" : "load+wait each  ", warps,
             cudaGetErrorString(e), bytes / h[0], h[0]);
    }
  return 0;
}
