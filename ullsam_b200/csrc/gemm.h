// ullsam_b200 — host-visible declarations of the tcgen05 GEMM (see gemm_sm100.cuh for the kernel).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ullsam {

enum EpiMode : int {
  EPI_BF16 = 0,     // out bf16 [M,ldc] = act(acc + bias)
  EPI_RES_F32 = 1,  // out fp32 [M,ldc] = acc + bias + residual[M,ldc]      (image_encoder.py:179-180)
  EPI_QKV = 2,      // q/k/v bf16 [G][heads][T][64] <- acc + bias            (image_encoder.py:227-229)
  EPI_PATCH = 3,    // out fp32 = acc + bias + pos[row % pos_rows][n]       (image_encoder.py:107-109)
  EPI_LN_BF16 = 4,  // LayerNorm2d over the N == BN channels -> bf16 NHWC    (common.py:31-43)
  EPI_LN_NCHW = 5,  // LayerNorm2d -> fp32 [B][N][hw]                       (image_encoder.py:114)
  EPI_F32 = 6,      // out fp32 [M,ldc] = act(acc + bias)
  EPI_LN_RES = 7,   // x = acc + bias + residual; LayerNorm(256) -> out fp32 (may alias residual) and out2 bf16
                    //   (TwoWayAttentionBlock: keys = norm4(keys + attn_out), transformer.py:178-182)
  EPI_UP1 = 8,      // ConvT(256->64,k2s2) + LayerNorm2d(64) + GELU -> bf16 NHWC [R,128,128,64]  (mask_decoder.py:54-56)
  EPI_MASK = 9,     // ConvT(64->32,k2s2) + GELU, contracted with hyper_in -> mask logits         (mask_decoder.py:57-58,144)
  EPI_GELU_BF16 = 10,  // EPI_BF16 with the exact-erf GELU resolved at compile time (MLPBlock.lin1, common.py:21-26);
                       // selected internally when EPI_BF16 is launched with act == ACT_GELU
  EPI_RES_LN = 11,  // EPI_RES_F32 + the LayerNorm that FOLLOWS the residual add (Block: norm2 after the attention
                    // branch, the next block's norm1 after the MLP branch, image_encoder.py:166-182) fused in: the CTA pair
                    // walks the N tiles of one 256-row block back to back, accumulates the row statistics on the way,
                    // re-reads the rows it has just written (L2) and stores out2 = bf16 LN(x) (optionally in
                    // window-partitioned row order, image_encoder.py:243-264) — saves the HBM round trip of a separate
                    // LayerNorm kernel
  // ---- LayerNorm folded into the GEMMs on either side of it (Block.norm1 / norm2, image_encoder.py:166-182) ----
  //   LN(x) W^T + b  =  rstd * (x (gamma*W)^T  -  mean * s)  +  c,     s_n = sum_k gamma_k W_nk,   c_n = sum_k beta_k W_nk + b_n
  // so the GEMM that PRODUCES the residual stream also emits a bf16 copy of it and per-row partial sums, and the GEMM that
  // CONSUMES the normalised rows runs on the un-normalised copy with the weight gamma*W and finishes the LayerNorm in
  // its epilogue.  No LayerNorm kernel, no second pass over the fp32 stream.
  EPI_RES_STATS = 12,    // EPI_RES_F32, plus out2 = bf16(out) and row_stats[row][2 * n_blk + half] = (sum, sum of squares)
                         // of the row's 128-column slice (fp32 values, before the bf16 rounding)
  EPI_PATCH_STATS = 13,  // EPI_PATCH, same two extra outputs
  EPI_QKV_FOLD = 14,     // EPI_QKV on the un-normalised rows: acc -> rstd * (acc - mean * fold_s[n]) + bias[n]; mean / rstd
                         // of a row from the K / 128 partial sums in row_stats; pad_bias = the projection's own bias
                         // (k / v rows of the window-padding tokens)
  EPI_GELU_FOLD = 15,    // EPI_GELU_BF16, same
  EPI_RES_RED = 16,      // EPI_RES_F32 with out == residual, the add done by the memory system: the epilogue stages acc + bias
                         // in shared memory (32 x 32 fp32 boxes, 128-byte swizzle) and issues TMA reduce-add stores
                         // (cp.reduce.async.bulk.tensor ... .add) — the residual stream is never read into the SM, so the
                         // epilogue does not wait for DRAM at all.  Bit-identical to EPI_RES_F32 (same two roundings).
};
// EPI_BF16 / EPI_F32 also add pos[row % pos_rows][n] when p.pos != nullptr.

enum ActMode : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2 };

struct GemmParams {
  int M, N, K;        // K = reduction length (conv3x3: 9 * C)
  int a_conv3;        // 0: A is a plain [M,K] matrix; 1: implicit 3x3/pad-1 conv over NHWC [B,64,64,C]
  int conv_c;         // channels of the conv input
  const float* bias;  // [N] or nullptr
  void* out;
  int ldc;
  int act;
  const float* residual;
  const float* pos;
  int pos_rows;
  __nv_bfloat16* q;
  __nv_bfloat16* k;
  __nv_bfloat16* v;
  int T;      // rows per attention group (196 windowed, 4096 global)
  int heads;  // attention heads; model dim = heads*64
  const float* ln_w;
  const float* ln_b;
  float ln_eps;
  int hw;  // tokens per image for the NCHW store
  void* out2;          // EPI_LN_RES / EPI_RES_LN: bf16 copy of the normalised rows
  int ln_partition;    // EPI_RES_LN: 0 = out2 rows in token order, 1 = 14x14-window-partitioned order (25 windows of 196
                       // rows per image; the zero pad rows are never written here)
                       // EPI_QKV: 1 = the rows of A are the image tokens [B,64,64]; the epilogue scatters them into the
                       // 14x14 windows of q/k/v [B*25][heads][196][64] (T must be 196) and writes k = v = bias for the
                       // padding tokens (their q rows are not written)
  const float* hyper;  // EPI_MASK: hyper_in [R][4][32]
  int mask_m0;         // EPI_MASK: first mask channel written
  int mask_cnt;        // EPI_MASK: number of mask channels written (out is [R][mask_cnt][256][256] fp32)
  float* row_stats;       // EPI_*_STATS (written) / EPI_*_FOLD (read): [M][kStatSlots] pairs (sum, sum of squares)
  const float* fold_s;    // EPI_*_FOLD: [N] column sums of the bf16 weight gamma*W
  const float* pad_bias;  // EPI_QKV_FOLD with ln_partition: [N] bias of the projection itself
  float* kmax_sq;         // EPI_QKV / EPI_QKV_FOLD, optional: [M / T][heads] running maximum of |k row|^2 per (group, head),
                          // the key-norm bound the global attention uses (attention.h) — folded with atomicMax on the float
                          // bits (non-negative values), so the caller zeroes it first; T must be a multiple of 32
};

constexpr int kStatSlots = 8;  // 128-column slices per row the statistics buffer has room for (LayerNorm dim <= 1024)

// All return nullptr on success or a static error string.
const char* make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                              uint32_t box_cols);
// fp32 [rows, cols] row-major, 128-byte swizzle (box_cols * 4 <= 128)
const char* make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                             uint32_t box_cols);
const char* make_tmap_nhwc_bf16(CUtensorMap* out, const void* base, int B, int H, int W, int C, uint32_t box_c,
                                uint32_t box_w, uint32_t box_h);
// A: bf16 [M,K] row-major (or NHWC [B,64,64,C] when p.a_conv3); W: bf16 [N,K] row-major.
const char* launch_gemm(int epi, const void* A, const void* W, const GemmParams& p, int num_sms, cudaStream_t stream);

}  // namespace ullsam
