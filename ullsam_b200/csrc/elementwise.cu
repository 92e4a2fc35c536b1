// ullsam_b200 — memory-bound glue kernels of the encoder: LayerNorm (+ window partition), PatchEmbed
// im2col, dtype casts.  All vectorised (16-byte accesses) and coalesced; one warp per token row.
#include "elementwise.h"

#include "common.cuh"
#include "host_util.h"

namespace ullsam {

// ---------------------------------------------------------------------------------------------
// LayerNorm over the channel dim of fp32 tokens -> bf16, optionally written in window-partitioned
// order with zero rows for the padding (reference: Block.forward norm1 + window_partition,
// image_encoder.py:168-172, 243-264; the padding is inserted AFTER the norm, so pad rows are 0).
// ---------------------------------------------------------------------------------------------
template <int VPL>  // float4 per lane; dim = VPL * 128
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int grid, int window,
                                                        int rows_out, __nv_bfloat16* __restrict__ y) {
  constexpr int DIM = VPL * 128;
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows_out) return;
  long src = row;
  bool valid = true;
  if (window > 0) {
    const int nwin = (grid + window - 1) / window;
    const int wtok = window * window;
    const int per_img = nwin * nwin * wtok;
    const int b = row / per_img, r = row - b * per_img;
    const int w = r / wtok, t = r - w * wtok;
    const int yy = (w / nwin) * window + t / window;
    const int xx = (w % nwin) * window + t % window;
    valid = (yy < grid) && (xx < grid);
    src = (static_cast<long>(b) * grid + yy) * grid + xx;
  }
  uint2* dst = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * DIM);
  if (!valid) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) dst[i * 32 + lane] = make_uint2(0u, 0u);
    return;
  }
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(src) * DIM);
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / DIM);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    sq += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / DIM) + eps);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    const float o0 = (v[i].x - mean) * rstd * g.x + b.x;
    const float o1 = (v[i].y - mean) * rstd * g.y + b.y;
    const float o2 = (v[i].z - mean) * rstd * g.z + b.z;
    const float o3 = (v[i].w - mean) * rstd * g.w + b.w;
    dst[i * 32 + lane] = make_uint2(pack_bf16(o0, o1), pack_bf16(o2, o3));
  }
}

const char* launch_layernorm(const float* x, const float* gamma, const float* beta, float eps, int B, int grid, int dim,
                             int window, __nv_bfloat16* y, cudaStream_t stream) {
  int rows_out;
  if (window > 0) {
    const int nwin = (grid + window - 1) / window;
    rows_out = B * nwin * nwin * window * window;
  } else {
    rows_out = B * grid * grid;
  }
  if (rows_out == 0) return nullptr;
  const int wpb = 8;
  const int blocks = (rows_out + wpb - 1) / wpb;
  cudaError_t le = cudaSuccess;
  switch (dim) {
    case 256: le = launch_pdl(layernorm_kernel<2>, dim3(blocks), dim3(wpb * 32), 0, stream, 1, x, gamma, beta, eps, grid, window, rows_out, y); break;
    case 768: le = launch_pdl(layernorm_kernel<6>, dim3(blocks), dim3(wpb * 32), 0, stream, 1, x, gamma, beta, eps, grid, window, rows_out, y); break;
    case 1024: le = launch_pdl(layernorm_kernel<8>, dim3(blocks), dim3(wpb * 32), 0, stream, 1, x, gamma, beta, eps, grid, window, rows_out, y); break;
    case 1280: le = launch_pdl(layernorm_kernel<10>, dim3(blocks), dim3(wpb * 32), 0, stream, 1, x, gamma, beta, eps, grid, window, rows_out, y); break;
    case 2048: le = launch_pdl(layernorm_kernel<16>, dim3(blocks), dim3(wpb * 32), 0, stream, 1, x, gamma, beta, eps, grid, window, rows_out, y); break;
    default: return "layernorm: dim must be 256, 768, 1024, 1280 or 2048";
  }
  return le == cudaSuccess ? nullptr : cudaGetErrorString(le);
}

// ---------------------------------------------------------------------------------------------
// InternVL pixel shuffle between the SAM embedding and the LLM tokens (modeling_internvl_sam.py:211-270,
// downsample_ratio 0.5): tokens[b][(y/2)*32 + x/2][(y%2)*512 + (x%2)*256 + c] <-> emb[b][c][y][x] for ps_version 'v2';
// 'v1' leaves height and width swapped in the token order ((x/2)*32 + y/2).  to_tokens = 1: NCHW -> tokens (the
// gather of extract_feature), 0: tokens -> NCHW (the scatter at the end of text_aware_dense_feature).
// One thread per element of the DESTINATION, fastest index contiguous there; 4 MB per image, L2-served.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pixel_shuffle_kernel(const float* __restrict__ in, int B, int v1, int to_tokens,
                                                            float* __restrict__ out) {
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t total = static_cast<size_t>(B) * 256 * 4096;
  if (idx >= total) return;
  const int b = static_cast<int>(idx >> 20);
  const int r = static_cast<int>(idx & ((1u << 20) - 1));
  if (to_tokens) {
    const int tok = r >> 10, f = r & 1023;
    const int a = tok >> 5, bb = tok & 31;            // v2: a = y/2, bb = x/2; v1: a = x/2, bb = y/2
    const int yh = v1 ? bb : a, xh = v1 ? a : bb;
    const int y = 2 * yh + (f >> 9), x = 2 * xh + ((f >> 8) & 1), c = f & 255;
    out[idx] = in[(static_cast<size_t>(b) * 256 + c) * 4096 + y * 64 + x];
  } else {
    const int c = r >> 12, y = (r >> 6) & 63, x = r & 63;
    const int yh = y >> 1, xh = x >> 1;
    const int tok = v1 ? (xh * 32 + yh) : (yh * 32 + xh);
    const int f = (y & 1) * 512 + (x & 1) * 256 + c;
    out[idx] = in[(static_cast<size_t>(b) * 1024 + tok) * 1024 + f];
  }
}

const char* launch_pixel_shuffle(const float* in, int B, int v1, int to_tokens, float* out, cudaStream_t stream) {
  if (B <= 0) return nullptr;
  const size_t total = static_cast<size_t>(B) * 256 * 4096;
  pixel_shuffle_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(in, B, v1, to_tokens, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// ---------------------------------------------------------------------------------------------
// PatchEmbed im2col: NCHW image -> bf16 [B*gh*gw, C*P*P], column = c*P*P + ky*P + kx, which is the
// flattening of Conv2d's weight [embed, C, P, P] (image_encoder.py:387-395).  P is fixed at 16.
// Each thread moves 8 horizontally adjacent pixels (fully coalesced reads, 16-byte writes).
// ---------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256) im2col16_kernel(const TIn* __restrict__ img, int B, int C, int S,
                                                       __nv_bfloat16* __restrict__ out) {
  const int x8_per_row = S / 8;
  const size_t total = static_cast<size_t>(B) * C * S * x8_per_row;
  const int g = S / 16;  // patches per side
  const int Kdim = C * 256;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x8 = static_cast<int>(i % x8_per_row);
    size_t r = i / x8_per_row;
    const int yy = static_cast<int>(r % S);
    r /= S;
    const int c = static_cast<int>(r % C);
    const int b = static_cast<int>(r / C);
    const TIn* src = img + ((static_cast<size_t>(b) * C + c) * S + yy) * S + x8 * 8;
    float f[8];
    if constexpr (sizeof(TIn) == 4) {
      const float4 a = *reinterpret_cast<const float4*>(src);
      const float4 d = *reinterpret_cast<const float4*>(src + 4);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = d.x; f[5] = d.y; f[6] = d.z; f[7] = d.w;
    } else {
      const uint4 a = *reinterpret_cast<const uint4*>(src);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __bfloat1622float2(h[j]);
        f[2 * j] = t.x; f[2 * j + 1] = t.y;
      }
    }
    const int px = x8 >> 1, kx0 = (x8 & 1) * 8;
    const int py = yy >> 4, ky = yy & 15;
    __nv_bfloat16* dst = out + (static_cast<size_t>(b) * g * g + py * g + px) * Kdim + c * 256 + ky * 16 + kx0;
    *reinterpret_cast<uint4*>(dst) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

const char* launch_im2col16(const void* images, int is_bf16, int B, int C, int S, __nv_bfloat16* out, cudaStream_t stream) {
  if (B == 0) return nullptr;
  if (S % 16 != 0) return "im2col: image size must be a multiple of 16";
  const size_t total = static_cast<size_t>(B) * C * S * (S / 8);
  const int blocks = static_cast<int>((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (is_bf16)
    im2col16_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(images), B, C, S, out);
  else
    im2col16_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(images), B, C, S, out);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// ---------------------------------------------------------------------------------------------
// casts
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, size_t n, __nv_bfloat16* __restrict__ y) {
  const size_t n4 = n / 4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(y)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const size_t i = n4 * 4 + threadIdx.x;
    y[i] = __float2bfloat16(x[i]);
  }
}

__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, size_t n, float* __restrict__ y) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}

__global__ void __launch_bounds__(256) cast_f16_f32_kernel(const __half* __restrict__ x, size_t n, float* __restrict__ y) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    y[i] = __half2float(x[i]);
}

static int grid_for(size_t n) {
  size_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  if (b == 0) b = 1;
  return static_cast<int>(b);
}

const char* launch_cast_f32_bf16(const float* x, size_t n, __nv_bfloat16* y, cudaStream_t stream) {
  if (n == 0) return nullptr;
  cast_f32_bf16_kernel<<<grid_for(n / 4 + 1), 256, 0, stream>>>(x, n, y);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* launch_cast_to_f32(const void* x, int dtype, size_t n, float* y, cudaStream_t stream) {
  if (n == 0) return nullptr;
  if (dtype == 1)
    cast_bf16_f32_kernel<<<grid_for(n), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), n, y);
  else if (dtype == 2)
    cast_f16_f32_kernel<<<grid_for(n), 256, 0, stream>>>(static_cast<const __half*>(x), n, y);
  else
    return "cast: unsupported dtype";
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace ullsam
