// ullsam_b200 — PromptEncoder on the device: random-Fourier positional encodings of points, boxes
// and of the dense grid, and the LLM-conditioned dense prompt.  All tiny / memory-bound; fp32 throughout.
#include "prompt.h"

#include <cuda_bf16.h>
#include <math.h>

namespace ullsam {
namespace {

inline const char* last_err() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

constexpr float kTwoPi = 6.283185307179586f;

// [sin, cos](2*pi * ((2c - 1) @ G)) for one coordinate pair in [0,1]^2 (prompt_encoder.py:220-228)
__device__ __forceinline__ void pe_feature(const float* __restrict__ gauss, int F, int f, float cx, float cy, float& s, float& c) {
  const float a = kTwoPi * ((2.f * cx - 1.f) * gauss[f] + (2.f * cy - 1.f) * gauss[F + f]);
  sincosf(a, &s, &c);
}

__global__ void pe_grid_kernel(const float* __restrict__ gauss, int F, int g, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // over F * g * g, pixel fastest
  if (idx >= F * g * g) return;
  const int f = idx / (g * g), pix = idx - f * g * g, y = pix / g, x = pix - y * g;
  float s, c;
  pe_feature(gauss, F, f, (x + 0.5f) / g, (y + 0.5f) / g, s, c);
  out[idx] = s;
  out[static_cast<size_t>(F) * g * g + idx] = c;
}

// one block per (prompt, token); threads over the F frequencies
__global__ void sparse_kernel(const float* __restrict__ gauss, int F, const float* __restrict__ point_emb,
                              const float* __restrict__ not_a_point, const float* __restrict__ coords,
                              const float* __restrict__ labels, int Np, int pad, const float* __restrict__ boxes, float img_w,
                              float img_h, int Ns, float* __restrict__ out) {
  const int p = blockIdx.y, t = blockIdx.x;
  const int npts = coords ? Np + pad : 0;
  float cx, cy;
  const float* add = nullptr;  // learned embedding added to the positional encoding
  bool zero_pe = false;
  if (t < npts) {
    float label = -1.f;
    cx = cy = 0.f;  // the padding point (prompt_encoder.py:84-87): label -1, its encoding is discarded anyway
    if (t < Np) {
      cx = (coords[(static_cast<size_t>(p) * Np + t) * 2] + 0.5f) / img_w;       // shift to the pixel centre (:83)
      cy = (coords[(static_cast<size_t>(p) * Np + t) * 2 + 1] + 0.5f) / img_h;
      label = labels[static_cast<size_t>(p) * Np + t];
    }
    if (label == -1.f) { zero_pe = true; add = not_a_point; }
    else if (label == 0.f) add = point_emb;
    else if (label == 1.f) add = point_emb + 2 * F;
  } else {
    const int corner = t - npts;  // 0: top-left, 1: bottom-right (:96-103)
    cx = (boxes[static_cast<size_t>(p) * 4 + corner * 2] + 0.5f) / img_w;
    cy = (boxes[static_cast<size_t>(p) * 4 + corner * 2 + 1] + 0.5f) / img_h;
    add = point_emb + (2 + corner) * 2 * F;
  }
  float* o = out + (static_cast<size_t>(p) * Ns + t) * 2 * F;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float s = 0.f, c = 0.f;
    if (!zero_pe) pe_feature(gauss, F, f, cx, cy, s, c);
    o[f] = s + (add ? add[f] : 0.f);
    o[F + f] = c + (add ? add[F + f] : 0.f);
  }
}

// LayerNorm over the channel dim of an NCHW map, per pixel.  Block = 32 pixels x 8 channel groups.
template <typename T>
__device__ __forceinline__ float ld(const T* p);
template <>
__device__ __forceinline__ float ld<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void __launch_bounds__(256) llm_dense_kernel(const T* __restrict__ x, int C, int HW, const float* __restrict__ scale,
                                                        const float* __restrict__ bias, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int p = blockIdx.y;
  const int px = blockIdx.x * 32 + (threadIdx.x & 31), grp = threadIdx.x >> 5;
  const bool ok = px < HW;
  const T* xp = x + static_cast<size_t>(p) * C * HW + px;
  float s = 0.f;
  if (ok)
    for (int c = grp; c < C; c += 8) s += ld(xp + static_cast<size_t>(c) * HW);
  red[grp][threadIdx.x & 31] = s;
  __syncthreads();
  float mean = 0.f;
#pragma unroll
  for (int gI = 0; gI < 8; ++gI) mean += red[gI][threadIdx.x & 31];
  mean /= C;
  __syncthreads();
  float q = 0.f;
  if (ok)
    for (int c = grp; c < C; c += 8) {
      const float d = ld(xp + static_cast<size_t>(c) * HW) - mean;
      q = fmaf(d, d, q);
    }
  red[grp][threadIdx.x & 31] = q;
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int gI = 0; gI < 8; ++gI) var += red[gI][threadIdx.x & 31];
  const float rstd = rsqrtf(var / C + 1e-5f);
  const float sc = scale[0], bi = bias[0];
  if (ok) {
    float* op = out + static_cast<size_t>(p) * C * HW + px;
    for (int c = grp; c < C; c += 8)
      op[static_cast<size_t>(c) * HW] = (ld(xp + static_cast<size_t>(c) * HW) - mean) * rstd * sc + bi;
  }
}

}  // namespace

const char* launch_prompt_pe_grid(const float* gauss, int F, int g, float* out, cudaStream_t stream) {
  if (F <= 0 || g <= 0) return "prompt: bad positional-encoding size";
  const int n = F * g * g;
  pe_grid_kernel<<<(n + 255) / 256, 256, 0, stream>>>(gauss, F, g, out);
  return last_err();
}

const char* launch_prompt_sparse(const float* gauss, int F, const float* point_emb, const float* not_a_point, const float* coords,
                                 const float* labels, int P, int Np, int pad, const float* boxes, float img_w, float img_h,
                                 float* out, cudaStream_t stream) {
  if (P <= 0) return nullptr;
  if ((coords == nullptr) != (labels == nullptr)) return "prompt: point coordinates and labels come together";
  const int Ns = (coords ? Np + pad : 0) + (boxes ? 2 : 0);
  if (Ns == 0) return nullptr;
  sparse_kernel<<<dim3(Ns, P), 128, 0, stream>>>(gauss, F, point_emb, not_a_point, coords, labels, Np, pad, boxes, img_w, img_h,
                                                Ns, out);
  return last_err();
}

const char* launch_prompt_llm_dense(const void* x, int is_bf16, int P, int C, int HW, const float* scale, const float* bias,
                                    float* out, cudaStream_t stream) {
  if (P <= 0) return nullptr;
  const dim3 grid((HW + 31) / 32, P);
  if (is_bf16)
    llm_dense_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), C, HW, scale, bias, out);
  else
    llm_dense_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), C, HW, scale, bias, out);
  return last_err();
}

}  // namespace ullsam
