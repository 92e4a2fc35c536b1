// ullsam_b200 — mask post-processing kernels: bilinear resize fused with binarisation, and the
// IoU metric of binarised masks.  Memory-bound; one thread per output pixel, x fastest (coalesced stores).
#include "postprocess.h"

namespace ullsam {
namespace {

inline const char* last_err() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// PyTorch's area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false)
__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& lam) {
  float s = scale * (dst + 0.5f) - 0.5f;
  s = s < 0.f ? 0.f : s;
  i0 = static_cast<int>(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  lam = s - static_cast<float>(i0);
}

__global__ void __launch_bounds__(256) resize_kernel(const float* __restrict__ x, int in_h, int in_w, int crop_h, int crop_w,
                                                     int out_h, int out_w, float scale_h, float scale_w, float threshold,
                                                     float* __restrict__ out_f32, uint8_t* __restrict__ out_u8) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  const int plane = blockIdx.z;
  if (ox >= out_w) return;
  int y0, y1, x0, x1;
  float ly, lx;
  src_index(scale_h, oy, crop_h, y0, y1, ly);
  src_index(scale_w, ox, crop_w, x0, x1, lx);
  const float* p = x + static_cast<size_t>(plane) * in_h * in_w;
  const float v00 = __ldg(p + y0 * in_w + x0), v01 = __ldg(p + y0 * in_w + x1);
  const float v10 = __ldg(p + y1 * in_w + x0), v11 = __ldg(p + y1 * in_w + x1);
  // same association as ATen's upsample_bilinear2d: h0lambda * (w0lambda * v00 + w1lambda * v01) + h1lambda * (...)
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float v = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
  const size_t o = (static_cast<size_t>(plane) * out_h + oy) * out_w + ox;
  if (out_f32) out_f32[o] = v;
  if (out_u8) out_u8[o] = v > threshold ? 1 : 0;
}

__global__ void __launch_bounds__(256) mask_iou_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                       size_t hw, float threshold, unsigned long long* __restrict__ counts) {
  const int plane = blockIdx.y;
  const float* a = logits + static_cast<size_t>(plane) * hw;
  const float* b = target + static_cast<size_t>(plane) * hw;
  unsigned int inter = 0, pa = 0, pb = 0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < hw; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const bool p = a[i] > threshold, t = b[i] > 0.5f;
    inter += (p && t);
    pa += p;
    pb += t;
  }
  for (int o = 16; o > 0; o >>= 1) {
    inter += __shfl_xor_sync(0xffffffffu, inter, o);
    pa += __shfl_xor_sync(0xffffffffu, pa, o);
    pb += __shfl_xor_sync(0xffffffffu, pb, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&counts[plane * 3 + 0], static_cast<unsigned long long>(inter));
    atomicAdd(&counts[plane * 3 + 1], static_cast<unsigned long long>(pa));
    atomicAdd(&counts[plane * 3 + 2], static_cast<unsigned long long>(pb));
  }
}

__global__ void mask_iou_finish_kernel(const unsigned long long* __restrict__ counts, int n, float smooth, float* __restrict__ iou) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float inter = static_cast<float>(counts[i * 3]);
  const float uni = static_cast<float>(counts[i * 3 + 1]) + static_cast<float>(counts[i * 3 + 2]) - inter;
  iou[i] = (inter + smooth) / (uni + smooth);
}

}  // namespace

const char* launch_resize_masks(const float* x, int n, int in_h, int in_w, int crop_h, int crop_w, int out_h, int out_w,
                                float threshold, float* out_f32, uint8_t* out_u8, cudaStream_t stream) {
  if (n <= 0) return nullptr;
  if (in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0) return "resize_masks: bad size";
  if (crop_h <= 0 || crop_h > in_h || crop_w <= 0 || crop_w > in_w) return "resize_masks: crop window outside the input";
  if (!out_f32 && !out_u8) return "resize_masks: no output requested";
  if (n > 65535 || out_h > 65535) return "resize_masks: more than 65535 planes / rows per call";
  const float scale_h = static_cast<float>(crop_h) / out_h, scale_w = static_cast<float>(crop_w) / out_w;
  resize_kernel<<<dim3((out_w + 255) / 256, out_h, n), 256, 0, stream>>>(x, in_h, in_w, crop_h, crop_w, out_h, out_w, scale_h,
                                                                        scale_w, threshold, out_f32, out_u8);
  return last_err();
}

const char* launch_mask_iou(const float* logits, const float* target, int n, size_t hw, float threshold, float smooth, float* iou,
                            cudaStream_t stream) {
  if (n <= 0) return nullptr;
  if (n > 65535) return "mask_iou: more than 65535 planes per call";
  // the counters live at the end of the output buffer's stream-ordered scratch: allocate per call (tiny)
  unsigned long long* counts = nullptr;
  cudaError_t e = cudaMallocAsync(&counts, static_cast<size_t>(n) * 3 * sizeof(unsigned long long), stream);
  if (e != cudaSuccess) return cudaGetErrorString(e);
  cudaMemsetAsync(counts, 0, static_cast<size_t>(n) * 3 * sizeof(unsigned long long), stream);
  int bx = static_cast<int>((hw + 256 * 8 - 1) / (256 * 8));
  if (bx < 1) bx = 1;
  if (bx > 256) bx = 256;
  mask_iou_kernel<<<dim3(bx, n), 256, 0, stream>>>(logits, target, hw, threshold, counts);
  mask_iou_finish_kernel<<<(n + 127) / 128, 128, 0, stream>>>(counts, n, smooth, iou);
  const char* err = last_err();
  cudaFreeAsync(counts, stream);
  return err;
}

}  // namespace ullsam
