// ullsam_b200 — C ABI, operator-level entry points (include/ullsam.h, "operator-level" section).
#include <cuda_runtime.h>

#include <string>

#include "../../include/ullsam.h"
#include "api_internal.h"
#include "attention.h"
#include "elementwise.h"
#include "postprocess.h"
#include "prompt.h"
#include "gemm.h"

namespace ullsam {

thread_local std::string g_thread_error;

int fail(ullsam_handle* h, int code, const std::string& msg);  // model.cu

int op_fail(int code, const char* msg) {
  g_thread_error = msg ? msg : "unknown error";
  return code;
}

int device_sm_count() {
  static int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
    if (prop.major != 10) return -1;
    cached = prop.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace ullsam

using namespace ullsam;

extern "C" {

int ullsam_abi_version(void) { return ULLSAM_ABI_VERSION; }

int ullsam_op_gemm(const ullsam_gemm_desc* d, void* stream) {
  if (!d) return op_fail(ULLSAM_ERR_INVALID, "op_gemm: null descriptor");
  const int sms = device_sm_count();
  if (sms <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_gemm: no sm_100 device (no fallback path exists)");
  GemmParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.a_conv3 = d->a_conv3; p.conv_c = d->conv_c;
  p.bias = d->bias; p.out = d->out; p.ldc = d->ldc; p.act = d->act;
  p.residual = d->residual; p.pos = d->pos; p.pos_rows = d->pos_rows;
  p.q = static_cast<__nv_bfloat16*>(d->q); p.k = static_cast<__nv_bfloat16*>(d->k); p.v = static_cast<__nv_bfloat16*>(d->v);
  p.T = d->T; p.heads = d->heads;
  p.ln_w = d->ln_w; p.ln_b = d->ln_b; p.ln_eps = d->ln_eps; p.hw = d->hw;
  p.out2 = d->out2; p.ln_partition = d->ln_partition;
  p.row_stats = d->row_stats; p.fold_s = d->fold_s; p.pad_bias = d->pad_bias; p.kmax_sq = d->kmax_sq;
  const char* e = launch_gemm(d->epi, d->A, d->W, p, sms, static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

int ullsam_op_layernorm(const float* x, const float* gamma, const float* beta, float eps, int B, int grid, int dim,
                        int window, void* y_bf16, void* stream) {
  if (device_sm_count() <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_layernorm: no sm_100 device");
  const char* e = launch_layernorm(x, gamma, beta, eps, B, grid, dim, window, static_cast<__nv_bfloat16*>(y_bf16),
                                   static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

int ullsam_op_attention(const void* q, const void* k, const void* v, const void* rel_h, const void* rel_w, int G,
                        int heads, int grid, void* out_bf16, void* stream) {
  if (device_sm_count() <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_attention: no sm_100 device");
  const char* e = launch_attention(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                                   static_cast<const __nv_bfloat16*>(v), static_cast<const __nv_bfloat16*>(rel_h),
                                   static_cast<const __nv_bfloat16*>(rel_w), G, heads, grid,
                                   static_cast<__nv_bfloat16*>(out_bf16), static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

int ullsam_op_im2col(const void* images, int dtype, int B, void* patches_bf16, void* stream) {
  if (device_sm_count() <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_im2col: no sm_100 device");
  if (dtype != ULLSAM_F32 && dtype != ULLSAM_BF16) return op_fail(ULLSAM_ERR_INVALID, "op_im2col: dtype must be f32 or bf16");
  const char* e = launch_im2col16(images, dtype == ULLSAM_BF16, B, 3, 1024, static_cast<__nv_bfloat16*>(patches_bf16),
                                  static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

int ullsam_op_prompt_dense_pe(const float* gauss, int F, int g, float* out, void* stream) {
  if (device_sm_count() <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_prompt_dense_pe: no sm_100 device");
  const char* e = launch_prompt_pe_grid(gauss, F, g, out, static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

int ullsam_op_prompt_sparse(const float* gauss, int F, const float* point_emb, const float* not_a_point, const float* coords,
                            const float* labels, int P, int Np, int pad, const float* boxes, float img_w, float img_h,
                            float* out, void* stream) {
  if (device_sm_count() <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_prompt_sparse: no sm_100 device");
  const char* e = launch_prompt_sparse(gauss, F, point_emb, not_a_point, coords, labels, P, Np, pad, boxes, img_w, img_h, out,
                                       static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

int ullsam_op_prompt_llm_dense(const void* x, int dtype, int P, int C, int HW, const float* scale, const float* bias, float* out,
                               void* stream) {
  if (device_sm_count() <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_prompt_llm_dense: no sm_100 device");
  if (dtype != ULLSAM_F32 && dtype != ULLSAM_BF16) return op_fail(ULLSAM_ERR_INVALID, "op_prompt_llm_dense: dtype must be f32 or bf16");
  const char* e = launch_prompt_llm_dense(x, dtype == ULLSAM_BF16, P, C, HW, scale, bias, out, static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

int ullsam_op_prompt_mask_embed(const float* masks, int P, int G, int c1, int c2, int E, const float* const* params, float* out,
                                void* stream) {
  if (device_sm_count() <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_prompt_mask_embed: no sm_100 device");
  if (!masks || !params || !out) return op_fail(ULLSAM_ERR_INVALID, "op_prompt_mask_embed: null argument");
  for (int i = 0; i < 10; ++i)
    if (!params[i]) return op_fail(ULLSAM_ERR_INVALID, "op_prompt_mask_embed: null parameter tensor");
  const char* e = launch_prompt_mask_downscale(masks, P, G, c1, c2, E, params[0], params[1], params[2], params[3], params[4],
                                               params[5], params[6], params[7], params[8], params[9], out,
                                               static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

int ullsam_op_pixel_shuffle(const float* in, int B, int v1, int to_tokens, float* out, void* stream) {
  if (device_sm_count() <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_pixel_shuffle: no sm_100 device");
  const char* e = launch_pixel_shuffle(in, B, v1, to_tokens, out, static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

int ullsam_op_resize_masks(const float* x, int n, int in_h, int in_w, int crop_h, int crop_w, int out_h, int out_w,
                           float threshold, float* out_f32, unsigned char* out_u8, void* stream) {
  if (device_sm_count() <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_resize_masks: no sm_100 device");
  const char* e = launch_resize_masks(x, n, in_h, in_w, crop_h, crop_w, out_h, out_w, threshold, out_f32, out_u8,
                                      static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

int ullsam_op_mask_iou(const float* logits, const float* target, int n, size_t hw, float threshold, float smooth, float* iou,
                       void* stream) {
  if (device_sm_count() <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_mask_iou: no sm_100 device");
  const char* e = launch_mask_iou(logits, target, n, hw, threshold, smooth, iou, static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

int ullsam_op_cast_bf16(const float* x, size_t count, void* y_bf16, void* stream) {
  if (device_sm_count() <= 0) return op_fail(ULLSAM_ERR_NO_DEVICE, "op_cast_bf16: no sm_100 device");
  const char* e = launch_cast_f32_bf16(x, count, static_cast<__nv_bfloat16*>(y_bf16), static_cast<cudaStream_t>(stream));
  return e ? op_fail(ULLSAM_ERR_CUDA, e) : ULLSAM_OK;
}

}  // extern "C"
