// Microbenchmark: latency / throughput of packed fp32 (FFMA2, FADD2) vs scalar FFMA, FMNMX3 vs FMNMX, and MUFU.EX2,
// per SM sub-partition.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_rate f32x2_rate.cu && ./f32x2_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 256;  // loop trips
constexpr int CH = 8;      // independent chains per thread

template <int OP>
__device__ __forceinline__ void step(uint64_t (&a)[CH], float (&f)[2 * CH], uint64_t b, uint64_t c) {
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
    if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(c));
    if (OP == 2) {
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[2 * i]) : "f"(1.0001f), "f"(0.5f));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[2 * i + 1]) : "f"(1.0001f), "f"(0.5f));
    }
    if (OP == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[2 * i]) : "f"(f[2 * i + 1]), "f"(0.25f));
    if (OP == 4) {
      asm volatile("max.f32 %0, %0, %1;" : "+f"(f[2 * i]) : "f"(f[2 * i + 1]));
      asm volatile("max.f32 %0, %0, %1;" : "+f"(f[2 * i]) : "f"(0.25f));
    }
    if (OP == 5) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[2 * i]));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[2 * i + 1]));
    }
    if (OP == 8) {  // MUFU on normal values: x in [0,1) -> 2^x in [1,2) -> -1 (FADD on the FMA pipe)
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[2 * i]));
      asm volatile("add.f32 %0, %0, %1;" : "+f"(f[2 * i]) : "f"(-1.0f));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[2 * i + 1]));
      asm volatile("add.f32 %0, %0, %1;" : "+f"(f[2 * i + 1]) : "f"(-1.0f));
    }
    if (OP == 9) {  // MUFU.EX2 on a constant-ish normal input, no dependent chain on the result
      float y0, y1;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(f[2 * i]));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(f[2 * i + 1]));
      asm volatile("" :: "f"(y0), "f"(y1));
    }
    if (OP == 6) {  // the softmax mix per 2 elements, packed: FADD2, 2x MUFU, FADD2, F2FP
      uint64_t x;
      asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(x) : "l"(a[i]), "l"(c));
      float x0, x1;
      asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
      asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(x));
      uint32_t p;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(x1), "f"(x0));
      f[2 * i] = __uint_as_float(p);
    }
    if (OP == 7) {  // same mix, scalar
      float x0 = f[2 * i] + 0.5f, x1 = f[2 * i + 1] + 0.5f;
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
      float s0, s1;
      asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(a[i]));
      s0 += x0; s1 += x1;
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(s0), "f"(s1));
      uint32_t p;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(x1), "f"(x0));
      f[2 * i] = __uint_as_float(p) * 1e-30f;
    }
  }
}

template <int OP>
__global__ void bench(long long* out, float* sink, int chains_used) {
  uint64_t a[CH];
  float f[2 * CH];
  for (int i = 0; i < CH; ++i) {
    float lo = threadIdx.x * 1e-4f + i * 0.01f, hi = lo + 0.5f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(lo), "f"(hi));
    f[2 * i] = lo; f[2 * i + 1] = hi;
  }
  uint64_t b, c;
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(1.0001f), "f"(0.9999f));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(0.5f), "f"(0.25f));
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) step<OP>(a, f, b, c);
  long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < CH; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
    acc += lo + hi + f[2 * i] + f[2 * i + 1];
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter_ops) {
  long long* d; float* s;
  cudaMalloc(&d, 8); cudaMalloc(&s, 4 * 1024 * 148);
  for (int threads : {32, 128, 256, 512, 1024}) {
    bench<OP><<<148, threads>>>(d, s, CH);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    int warps_per_smsp = (threads / 32 + 3) / 4;
    double per_instr = double(h) / (ITER * CH * per_iter_ops) / warps_per_smsp;
    printf("%-28s threads %4d: %8lld clk, %.2f clk per warp-instruction per SMSP (%d warps/SMSP)\n", name, threads, h, per_instr, warps_per_smsp);
  }
  cudaFree(d); cudaFree(s);
}

int main() {
  run<0>("FFMA2", 1);
  run<1>("FADD2", 1);
  run<2>("FFMA (2 per step)", 2);
  run<3>("FMNMX3", 1);
  run<4>("FMNMX (2 per step)", 2);
  run<5>("MUFU.EX2 (2 per step)", 2);
  run<8>("MUFU.EX2 normal + FADD (x2)", 2);
  run<9>("MUFU.EX2 normal indep (x2)", 2);
  run<6>("softmax mix packed (per pair)", 1);
  run<7>("softmax mix scalar (per pair)", 1);
  return 0;
}
