// ullsam_b200 — launchers of the memory-bound glue kernels (elementwise.cu).
// Every launcher returns nullptr on success or a static error string.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace ullsam {

const char* launch_layernorm(const float* x, const float* gamma, const float* beta, float eps, int B, int grid, int dim,
                             int window, __nv_bfloat16* y, cudaStream_t stream);
// InternVL pixel shuffle (modeling_internvl_sam.py:211-270): emb fp32 [B,256,64,64] <-> tokens fp32 [B,1024,1024]
const char* launch_pixel_shuffle(const float* in, int B, int v1, int to_tokens, float* out, cudaStream_t stream);
const char* launch_im2col16(const void* images, int is_bf16, int B, int C, int S, __nv_bfloat16* out, cudaStream_t stream);
const char* launch_cast_f32_bf16(const float* x, size_t n, __nv_bfloat16* y, cudaStream_t stream);
const char* launch_cast_to_f32(const void* x, int dtype, size_t n, float* y, cudaStream_t stream);

}  // namespace ullsam
