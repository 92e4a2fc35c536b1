// ullsam_b200 — fused attention with decomposed relative-position bias (attention.cu).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace ullsam {

// q,k,v: bf16 [G][heads][T][64] with T = grid*grid; rel_h, rel_w: bf16 [NREL][64], the module's
// rel_pos_h / rel_pos_w ([2*grid-1][64]) zero-padded to NREL = 128 (grid 64) or 32 (grid 14) rows.
// out: bf16 [B][64][64][heads*64] token-major image layout.  grid == 64: G == B (global attention);
// grid == 14: G == 25*B windows of the 5x5 partition of the zero-padded 70x70 grid, padded query
// rows are dropped on the store (image_encoder.py:224-289, 325-361).
const char* launch_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v,
                             const __nv_bfloat16* rel_h, const __nv_bfloat16* rel_w, int G, int heads, int grid, __nv_bfloat16* out, cudaStream_t stream,
                             float* kmax_sq = nullptr /* grid 64: G * heads floats (else allocated stream-ordered) */,
                             bool kmax_valid = false /* kmax_sq already holds max_t |k[g][head][t]|^2 (written by the QKV GEMM's
                                                        epilogue, GemmParams::kmax_sq): the pre-pass over K is skipped */);

// second windowed design (attention_win.cu): one CTA per (window, head), both query tiles, 7 key tiles of 28
const char* launch_attention_windowed_v2(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v,
                                         const __nv_bfloat16* rel_h, const __nv_bfloat16* rel_w, int G, int heads,
                                         __nv_bfloat16* out, cudaStream_t stream);

}  // namespace ullsam
