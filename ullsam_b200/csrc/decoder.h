// ullsam_b200 — launchers of the prompt-token side of the mask decoder (decoder.cu).
// Every launcher returns nullptr on success or a static error string.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace ullsam {

// Y[g][m][n] = act((X[g][m][:] (+ X2[g][m][:])) . W[g][n][:] + b[g][n]) (+ res[g][m][n]);  fp32, small M.
struct TokLinear {
  const float* x;    // [G][M] rows of K, row stride ldx, group stride x_gs
  const float* x2;   // optional addend with x's layout (the "+ query_pe" of transformer.py:160,166,178)
  const float* w;    // [G][N][K], group stride w_gs
  const float* b;    // [G][N] or nullptr, group stride b_gs
  const float* res;  // optional residual with y's layout
  float* y;          // row stride ldy, group stride y_gs
  int M, N, K, G;
  int ldx, ldy;
  size_t x_gs, w_gs, b_gs, y_gs;
  int act;  // 0 none, 2 ReLU
};

// ---- fused token side (decoder_fused.cu): one cluster-of-four kernel per TwoWayAttentionBlock -------------------------------
constexpr int kDecFusedMaxTokens = 24;   // prompt tokens per prompt (1 + 4 + Ns) the fused kernels hold in shared memory
struct DecTokLayer {   // token-side weights of one block, fp32, nn.Linear layout [out][in]  (transformer.py:125-151)
  const float *sq_w, *sq_b, *sk_w, *sk_b, *sv_w, *sv_b, *so_w, *so_b;   // self_attn q/k/v/out_proj, 256 x 256
  const float *n1_w, *n1_b, *n2_w, *n2_b, *n3_w, *n3_b;
  const float *tq_w, *tq_b, *to_w, *to_b;     // cross_attn_token_to_image q_proj [128][256], out_proj [256][128]
  const float *m1_w, *m1_b, *m2_w, *m2_b;     // mlp lin1 [2048][256], lin2 [256][2048]
  const float *ik_w, *ik_b, *iv_w, *iv_b;     // cross_attn_image_to_token k_proj / v_proj [128][256]
  int mlp_dim;
  int skip_pe;                                // first block: skip_first_layer_pe (transformer.py:157-158)
};
struct DecTokFinal {   // final_attn_token_to_image + norm_final_attn (transformer.py:98-106), hypernets + IoU head (mask_decoder.py:59-69)
  const float *tq_w, *tq_b, *to_w, *to_b, *nf_w, *nf_b;
  const float *hyp_w0, *hyp_b0, *hyp_w1, *hyp_b1, *hyp_w2, *hyp_b2;   // [4][256][256], [4][256], [4][256][256], [4][256], [4][32][256], [4][32]
  const float* iou_w[4];
  const float* iou_b[4];
  int iou_depth, iou_hidden;
};
bool dec_tokens_fused_supported(int Nt, int mlp_dim, int iou_hidden);
// queries fp32 [R][Nt][256] in/out; tokens = point embedding [R][Nt][256]; proj bf16 [R*4096][ldkv] (K cols 0..127, V 128..255);
// k128 / v128 fp32 [R][Nt][128] out (the tokens' keys / values for the image->token attention)
const char* launch_dec_token_layer(const DecTokLayer& L, float* queries, const float* tokens, const __nv_bfloat16* proj, int ldkv,
                                   int R, int Nt, float* k128, float* v128, cudaStream_t stream);
// hyper fp32 [R][4][32] and iou4 fp32 [R][4] may be nullptr (TwoWayTransformer alone)
const char* launch_dec_token_final(const DecTokFinal& F, float* queries, const float* tokens, const __nv_bfloat16* proj, int ldkv,
                                   int R, int Nt, float* hyper, float* iou4, cudaStream_t stream);

const char* launch_tok_linear(const TokLinear& a, cudaStream_t stream);
const char* launch_tok_self_attn(const float* q, const float* k, const float* v, int R, int Nt, float* out,
                                 cudaStream_t stream);
const char* launch_tok_ln(const float* x, const float* gamma, const float* beta, int rows, float eps, float* y,
                          cudaStream_t stream);
// q fp32 [R][Nt][128]; kv bf16 [R*4096][ldkv] (K = cols 0..127, V = cols 128..255); S scratch fp32 [R][8][Nt][4096]
const char* launch_t2i_attention(const float* q, const __nv_bfloat16* kv, int ldkv, int R, int Nt, float* o,
                                 cudaStream_t stream);
// proj bf16 [R*4096][ldp] with the image-side queries at column q_col; k, v fp32 [R][Nt][128]; out bf16 [R*4096][128]
const char* launch_i2t_attention(const __nv_bfloat16* proj, int ldp, int q_col, const float* k, const float* v, int R,
                                 int Nt, __nv_bfloat16* out, cudaStream_t stream);
// keys[r][pix][c] = emb[r / P][c][pix] + dense (dense_mode 0: [dense_batch][256][4096]; 1: [dense_batch][256]; null: none)
const char* launch_keys_init(const float* emb, int R, int P, const float* dense, int dense_mode, int dense_batch,
                             float* keys_f32, __nv_bfloat16* keys_bf16, cudaStream_t stream);
const char* launch_tok_init(const float* iou_token, const float* mask_tokens, int nm, const float* sparse, int Ns, int R,
                            float* tokens, cudaStream_t stream);
const char* launch_slice_copy(const float* in, int R, int C, int m0, int cnt, size_t inner, float* out, cudaStream_t stream);

}  // namespace ullsam
