// ullsam_b200 — PromptEncoder kernels (prompt.cu): the step immediately upstream of the mask decoder
// (reference modeling/prompt_encoder.py).  Every launcher returns nullptr or a static error string.
#pragma once

#include <cuda_runtime.h>

namespace ullsam {

// PositionEmbeddingRandom.forward (prompt_encoder.py:230-241): out[2F][g][g] fp32.
const char* launch_prompt_pe_grid(const float* gauss /*[2][F]*/, int F, int g, float* out, cudaStream_t stream);

// PromptEncoder._embed_points + _embed_boxes (prompt_encoder.py:76-103) -> out [P][Ns][2F] fp32,
// Ns = (Np + pad) + (boxes ? 2 : 0); labels as fp32 (-1 not-a-point, 0 negative, 1 positive).
const char* launch_prompt_sparse(const float* gauss, int F, const float* point_emb /*[4][2F]*/, const float* not_a_point /*[2F]*/,
                                 const float* coords /*[P][Np][2] or null*/, const float* labels /*[P][Np] or null*/, int P,
                                 int Np, int pad, const float* boxes /*[P][4] or null*/, float img_w, float img_h, float* out,
                                 cudaStream_t stream);

// PromptEncoder._normalize_llm_hidden_states (prompt_encoder.py:131-151): x [P][C][HW] (fp32 or bf16) ->
// out fp32 = LayerNorm_C(x) (no affine, eps 1e-5, biased variance) * scale[0] + bias[0].
const char* launch_prompt_llm_dense(const void* x, int is_bf16, int P, int C, int HW, const float* scale, const float* bias,
                                    float* out, cudaStream_t stream);

// PromptEncoder._embed_masks == mask_downscaling (prompt_encoder.py:53-61, 105-108): masks [P][1][4G][4G] fp32 ->
// out [P][E][G][G] fp32.  w0 [c1][1][2][2], w3 [c2][c1][2][2], w6 [E][c2][1][1]; g*/be* are the LayerNorm2d weight / bias.
const char* launch_prompt_mask_downscale(const float* masks, int P, int G, int c1, int c2, int E, const float* w0, const float* b0,
                                         const float* g1, const float* be1, const float* w3, const float* b3, const float* g4,
                                         const float* be4, const float* w6, const float* b6, float* out, cudaStream_t stream);

}  // namespace ullsam
