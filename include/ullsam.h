/* ullsam.h — C ABI of libullsam_b200.so: the B200-native (sm_100a) SAM segmentation branch of uLLSAM.
 *
 * This is the drop-in boundary for the one hot path this repository accelerates:
 *   ImageEncoderViT.forward   (reference modeling/image_encoder.py:106-116)
 *   MaskDecoder.forward       (reference modeling/mask_decoder.py:71-149)
 *   TwoWayTransformer.forward (reference modeling/transformer.py:62-108)
 * The reference has no FFI layer of its own (it is pure Python/PyTorch); the Python shims in
 * ullsam_b200/modeling/ keep the reference's nn.Module API and bind these entry points with ctypes
 * (see INTEGRATION.md for the stub a maintainer would add to the reference).
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  Every function returns 0 or a negative
 *     ullsam_status; nothing throws across the ABI.  ullsam_last_error() gives the text.
 *   - unless a name ends in _host, every data pointer is a DEVICE pointer owned by the caller, the
 *     call is asynchronous on `stream` (a cudaStream_t passed as void*), and nothing synchronises.
 *   - one handle per (device, model).  A handle is not thread-safe; distinct handles are independent.
 *   - there is no CPU fallback anywhere: without an sm_100 device every compute call fails.
 */
#ifndef ULLSAM_H_
#define ULLSAM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ULLSAM_ABI_VERSION 1

typedef enum {
  ULLSAM_OK = 0,
  ULLSAM_ERR_INVALID = -1,    /* bad argument / shape / unknown key */
  ULLSAM_ERR_CUDA = -2,       /* a CUDA runtime or driver call failed */
  ULLSAM_ERR_NO_DEVICE = -3,  /* no sm_100 device (there is no fallback path) */
  ULLSAM_ERR_STATE = -4,      /* weights missing / handle not ready */
  ULLSAM_ERR_NOMEM = -5
} ullsam_status;

typedef enum { ULLSAM_F32 = 0, ULLSAM_BF16 = 1, ULLSAM_F16 = 2 } ullsam_dtype;

/* Hyper-parameters: the arguments of ImageEncoderViT.__init__ (image_encoder.py:18-36),
 * TwoWayTransformer.__init__ (transformer.py:17-25) and MaskDecoder.__init__ (mask_decoder.py:17-26)
 * as fixed by build_sam._build_sam (build_sam.py:55-98). */
typedef struct {
  int32_t img_size;      /* 1024 */
  int32_t patch_size;    /* 16   */
  int32_t in_chans;      /* 3    */
  int32_t embed_dim;     /* 768 (vit_b) / 1024 (vit_l); vit_h (1280, 80-wide heads) is rejected by ullsam_create */
  int32_t depth;         /* 12 / 24 (0: decoder-only handle) */
  int32_t num_heads;     /* embed_dim / 64 */
  int32_t mlp_dim;       /* embed_dim * mlp_ratio */
  int32_t out_chans;     /* 256 */
  int32_t window_size;   /* 14 */
  int32_t num_global;    /* number of valid entries in global_attn_indexes */
  int32_t global_attn_indexes[8];
  float enc_ln_eps;      /* 1e-6 (build_sam.py:72) */
  /* decoder */
  int32_t transformer_dim; /* 256 */
  int32_t dec_depth;       /* 2 */
  int32_t dec_heads;       /* 8 */
  int32_t dec_mlp_dim;     /* 2048 */
  int32_t num_mask_tokens; /* num_multimask_outputs + 1 = 4 */
  int32_t iou_head_hidden; /* 256 */
  int32_t iou_head_depth;  /* 3 */
} ullsam_config;

typedef struct ullsam_handle ullsam_handle;

int ullsam_abi_version(void);

/* Text of the most recent failure on this handle (or, with h == NULL, of the last failed
 * ullsam_create on this thread).  Never NULL. */
const char* ullsam_last_error(const ullsam_handle* h);

/* Replaces the construction of the three modules in build_sam._build_sam (build_sam.py:66-98). */
int ullsam_create(const ullsam_config* cfg, int device, ullsam_handle** out);
void ullsam_destroy(ullsam_handle* h);

/* Replaces nn.Module.load_state_dict for the path (build_sam.py:103-106; key names SURVEY App. B,
 * e.g. "image_encoder.blocks.3.attn.qkv.weight", "mask_decoder.iou_token.weight").
 * `data` may be a host or device pointer; it is copied and re-packed into the kernels' layouts
 * (bf16 GEMM operands, fp32 vectors) before the call returns — the caller keeps ownership. */
int ullsam_set_weight(ullsam_handle* h, const char* state_dict_key, const void* data, int dtype,
                      const int64_t* shape, int ndim);
/* Same, stream-ordered and without any host synchronisation: `data` must be a DEVICE pointer (or pinned host memory)
 * that stays valid until the work queued on `stream` so far has run — e.g. a torch tensor allocated on that stream.
 * The copy and the re-pack kernels are queued on `stream`; the next forward call on a different stream waits for them
 * on the device.  Loading a 316-tensor state_dict this way costs 316 enqueues instead of 316 device synchronisations. */
int ullsam_set_weight_async(ullsam_handle* h, const char* state_dict_key, const void* data, int dtype,
                            const int64_t* shape, int ndim, void* stream);
/* Number of state_dict tensors still missing (0 when both modules are fully loaded). */
int ullsam_missing_weights(const ullsam_handle* h, int which /*0 encoder, 1 decoder*/);

/* Workspace the handle will hold for an encoder batch of B images / a decoder batch of R
 * (image, prompt) pairs.  Allocated lazily with cudaMalloc and kept until destroy. */
size_t ullsam_workspace_bytes(const ullsam_handle* h, int B, int R);

/* ImageEncoderViT.forward (image_encoder.py:106-116).
 * images : [B, 3, 1024, 1024] NCHW, dtype in {F32, BF16}
 * out_emb: [B, 256, 64, 64]   NCHW, out_dtype in {F32, BF16} */
int ullsam_encoder_forward(ullsam_handle* h, const void* images, int dtype, int B, void* out_emb,
                           int out_dtype, void* stream);

/* MaskDecoder.forward (mask_decoder.py:71-109) for n_img images x P prompts each; pair r = i*P + j
 * uses image i (the reference's repeat_interleave at mask_decoder.py:126; n_img == 1 is the
 * reference's own calling convention, n_img > 1 with P == 1 is the concatenation of per-image calls
 * made by Sam.forward, sam.py:101-117).
 * emb      : [n_img, 256, 64, 64] fp32 NCHW          image_pe: [256, 64, 64] fp32
 * sparse   : [n_img*P, Ns, 256] fp32
 * dense    : dense_mode 0: [dense_batch, 256, 64, 64] fp32, dense_batch in {1, n_img*P}
 *            dense_mode 1: [dense_batch, 256] fp32 — one vector per pair broadcast over the 64x64
 *                          grid (PromptEncoder's no_mask_embed expand, prompt_encoder.py:196-199)
 * masks    : [n_img*P, multimask ? 3 : 1, 256, 256] fp32      iou: [n_img*P, multimask ? 3 : 1] fp32 */
int ullsam_decoder_forward(ullsam_handle* h, const void* emb, int n_img, const void* image_pe,
                           const void* sparse, int P, int Ns, const void* dense, int dense_mode,
                           int dense_batch, int multimask, void* masks, void* iou, void* stream);

/* Same contract as MaskDecoder.predict_masks (mask_decoder.py:112-149): all 4 masks / IoUs. */
int ullsam_decoder_predict_masks(ullsam_handle* h, const void* emb, int n_img, const void* image_pe,
                                 const void* sparse, int P, int Ns, const void* dense, int dense_mode,
                                 int dense_batch, void* masks4, void* iou4, void* stream);

/* TwoWayTransformer.forward (transformer.py:62-108).
 * image_embedding, image_pe: [R, 256, 64, 64] / [256, 64, 64] fp32; point_embedding: [R, Nt, 256] fp32
 * queries: [R, Nt, 256] fp32; keys: [R, 4096, 256] fp32 */
int ullsam_transformer_forward(ullsam_handle* h, const void* image_embedding, int R, const void* image_pe,
                               const void* point_embedding, int Nt, void* queries, void* keys, void* stream);

/* End-to-end with HOST buffers (pinned recommended; pageable works but its copies are staged and serial): H2D of the
 * images and prompts, encoder, decoder (one prompt set per image), D2H of masks / iou / optionally embeddings;
 * synchronises before returning.  Internally the batch is cut into growing sub-batches (4, 12, 16 images of every 32) on two private streams so
 * the H2D copy of sub-batch i+1 and the D2H copy of sub-batch i-1 overlap the kernels of sub-batch i.
 * bench.py reports this call as `e2e.c_abi_host` (one synchronous call per step); its headline `e2e` figure is measured
 * one level up, through the Python Engine API with pinned host tensors, where copies also overlap ACROSS steps. */
int ullsam_segment_host(ullsam_handle* h, const float* images_host, int B, const float* image_pe_host,
                        const float* sparse_host, int Ns, const float* dense_vec_host, int multimask,
                        float* masks_host, float* iou_host, float* emb_host /* may be NULL */);

/* ---- instrumentation -------------------------------------------------------------------------- */

/* Number of CUDA kernels this handle has launched since creation (bench.py's gpu_launches). */
unsigned long long ullsam_launch_count(const ullsam_handle* h);

/* Per-category device timing with CUDA events recorded on the caller's stream between launches.
 * begin() arms it; the first marked launch after begin() is the time origin (not attributed); end()
 * synchronises on the last event and returns milliseconds / marked-launch counts per category. */
int ullsam_profile_begin(ullsam_handle* h);
int ullsam_profile_end(ullsam_handle* h, float* ms_per_category, int* launches_per_category, int n);
int ullsam_profile_categories(void);
const char* ullsam_profile_category_name(int i);

/* ---- operator-level entry points (the unit the parity tests exercise) ------------------------ */

/* tcgen05 GEMM: C = A[M,K] x W[N,K]^T with a fused epilogue.  Mirrors ullsam::GemmParams. */
typedef struct {
  int32_t epi;  /* 0 bf16 | 1 +residual fp32 | 2 qkv split | 3 +pos fp32 | 4 LN2d bf16 | 5 LN2d NCHW | 6 fp32 |
                 * 11 +residual fp32 AND the following nn.LayerNorm -> out2 bf16 (Block.norm2 / next Block.norm1,
                 *    image_encoder.py:166-182), optionally window-partitioned (image_encoder.py:243-264)
                 * LayerNorm folded into the GEMMs around it (the default encoder path; Block.norm1 / norm2, image_encoder.py:166-182):
                 *    LN(x) W^T + b = rstd * (x (gamma*W)^T - mean * s) + c,  s_n = sum_k (gamma*W)_nk,  c_n = sum_k beta_k W_nk + b_n
                 * 12 epi 1, plus out2 = bf16(out) and row_stats[row][2 * (n / 256) + (n % 256) / 128] = (sum, sum of squares)
                 *    of the row's 128-column slice | 13 epi 3 with the same two extra outputs
                 * 14 epi 2 on the un-normalised rows: W = bf16(gamma*W), bias = c, fold_s = s, mean / rstd from row_stats
                 *    (K / 128 slices), pad_bias = b | 15 epi 0 + GELU(erf), same
                 * 16 epi 1 with out == residual (in place), the add performed by the memory system (TMA reduce-add stores) */
  int32_t M, N, K;
  int32_t a_conv3, conv_c;
  const void* A;     /* bf16 */
  const void* W;     /* bf16 [N,K] */
  const float* bias; /* may be NULL */
  void* out;
  int32_t ldc;
  int32_t act;       /* 0 none | 1 GELU(erf) | 2 ReLU */
  const float* residual;
  const float* pos;
  int32_t pos_rows;
  void* q;
  void* k;
  void* v;
  int32_t T, heads;
  const float* ln_w;
  const float* ln_b;
  float ln_eps;
  int32_t hw;
  void* out2;           /* epi 11: bf16 LayerNorm(out) [M, N] (token order) or [M/4096*25*196, N] (window order) */
  int32_t ln_partition; /* epi 11: 1 = window order (pad rows are not written)
                         * epi 2: 1 = A's rows are the image tokens [B,64,64]; the store performs window_partition
                         *        (image_encoder.py:243-264) into q/k/v [B*25][heads][196][64] and also writes the k / v rows
                         *        of the 804 padding tokens per image (= bias: they are zero after norm1, :168-172, and
                         *        are attended to); their q rows are left untouched (cropped by window_unpartition) */
  float* row_stats;      /* epi 12, 13 (written), 14, 15 (read): fp32 [M][8][2] */
  const float* fold_s;   /* epi 14, 15: [N] */
  const float* pad_bias; /* epi 14 with ln_partition: [N] */
  float* kmax_sq;        /* epi 2, 14 (optional, rows in group order, T % 32 == 0): [M / T][heads], zeroed by the caller; the epilogue
                          * folds max |k row|^2 per (group, head) into it (atomicMax on the float bits) — the key-norm bound of
                          * the global attention, which otherwise takes a pre-pass over K */
} ullsam_gemm_desc;
int ullsam_op_gemm(const ullsam_gemm_desc* d, void* stream);

/* nn.LayerNorm over the last dim (image_encoder.py:168,180) of fp32 x[rows, dim] -> bf16.
 * window > 0 additionally performs window_partition (image_encoder.py:243-264): x is [B, grid, grid, dim]
 * and y is [B * nW, window*window, dim] with zero rows where the padded window leaves the image. */
int ullsam_op_layernorm(const float* x, const float* gamma, const float* beta, float eps, int B, int grid,
                        int dim, int window, void* y_bf16, void* stream);

/* Attention with decomposed relative position (image_encoder.py:224-240, 325-361).
 * q,k,v: bf16 [G][heads][T][64], T = grid*grid (grid = 14 windowed, 64 global); rel_h/rel_w: bf16
 * [NREL, 64] = rel_pos_h / rel_pos_w ([2*grid-1, 64]) zero-padded to NREL = 128 (grid 64) / 32 (grid 14) rows; out: bf16 [B, 64, 64, heads*64] — for grid == 14 the kernel also undoes the window
 * partition (window_unpartition, image_encoder.py:267-289) and drops the padded query rows. */
int ullsam_op_attention(const void* q, const void* k, const void* v, const void* rel_h, const void* rel_w,
                        int G, int heads, int grid, void* out_bf16, void* stream);

/* PatchEmbed's im2col (image_encoder.py:387-395): images [B,3,1024,1024] (dtype) -> bf16 [B*4096, 768]. */
int ullsam_op_im2col(const void* images, int dtype, int B, void* patches_bf16, void* stream);

/* ---- MLLM bridge (InternVL pixel shuffle either side of the LLM) ------------------- */

/* InternVLSAMModel.pixel_shuffle with downsample_ratio 0.5 as used by extract_feature (modeling_internvl_sam.py:211-251)
 * and its inverse at the end of text_aware_dense_feature (:253-270):
 *   to_tokens = 1: emb fp32 [B,256,64,64] -> tokens fp32 [B,1024,1024];  to_tokens = 0: the inverse.
 * v1 = 1 reproduces ps_version 'v1' (height / width left swapped in the token order). */
int ullsam_op_pixel_shuffle(const float* in, int B, int v1, int to_tokens, float* out, void* stream);

/* ---- PromptEncoder on the device (the step upstream of the decoder) ---------------- */

/* PositionEmbeddingRandom.forward == PromptEncoder.get_dense_pe without the batch dim (prompt_encoder.py:65-74, 230-241):
 * gauss = pe_layer.positional_encoding_gaussian_matrix [2, F] fp32 -> out fp32 [2F, g, g]. */
int ullsam_op_prompt_dense_pe(const float* gauss, int F, int g, float* out, void* stream);

/* PromptEncoder._embed_points / _embed_boxes (prompt_encoder.py:76-103) -> sparse embeddings fp32 [P, Ns, 2F] with
 * Ns = (coords ? Np + pad : 0) + (boxes ? 2 : 0).  coords [P, Np, 2] pixel XY, labels [P, Np] fp32 (-1 / 0 / 1),
 * boxes [P, 4] XYXY; point_emb = point_embeddings.{0..3}.weight stacked [4, 2F]; not_a_point [2F].
 * pad = 1 appends the padding point exactly as forward() does when no box is given (:171). */
int ullsam_op_prompt_sparse(const float* gauss, int F, const float* point_emb, const float* not_a_point, const float* coords,
                            const float* labels, int P, int Np, int pad, const float* boxes, float img_w, float img_h,
                            float* out, void* stream);

/* PromptEncoder._normalize_llm_hidden_states (prompt_encoder.py:131-151): x [P, C, HW] (dtype ULLSAM_F32 / ULLSAM_BF16)
 * -> out fp32 [P, C, HW] = layer_norm over C (no affine, eps 1e-5) * llm_scale_factor[0] + llm_bias[0]. */
int ullsam_op_prompt_llm_dense(const void* x, int dtype, int P, int C, int HW, const float* scale, const float* bias,
                               float* out, void* stream);

/* PromptEncoder._embed_masks == mask_downscaling (prompt_encoder.py:53-61, 105-108, 176-177): the mask-prompt branch of
 * stock SAM (Sam.forward feeds low-res mask inputs here, sam.py:106-110; uLLSAM itself passes masks=None).
 * masks fp32 [P, 1, 4G, 4G] -> out fp32 [P, E, G, G].  params = the 10 fp32 device tensors of mask_downscaling in
 * state_dict order: 0.weight [c1,1,2,2], 0.bias, 1.weight, 1.bias (LayerNorm2d c1), 3.weight [c2,c1,2,2], 3.bias,
 * 4.weight, 4.bias (LayerNorm2d c2), 6.weight [E,c2,1,1], 6.bias.  c1 <= 8, c2 <= 32 (SAM: 4, 16, E = 256, G = 64). */
int ullsam_op_prompt_mask_embed(const float* masks, int P, int G, int c1, int c2, int E, const float* const* params,
                                float* out, void* stream);

/* ---- mask post-processing (the step downstream of the decoder) ---------------------- */

/* F.interpolate(x[..., :crop_h, :crop_w], (out_h, out_w), mode="bilinear", align_corners=False) on n = B*C fp32 planes of
 * in_h x in_w, fused with the binarisation `> threshold` (Sam.mask_threshold = 0.0 on logits, sam.py:19,123 ==
 * sigmoid > 0.5, app.py:644).  One call is uLLSAM's upsampling to the model size (train_joint_v2.py:1073-1079,
 * app.py:635-641); two calls (-> img_size, then crop = input_size -> original_size) are Sam.postprocess_masks
 * (sam.py:133-162).  out_f32 (fp32 logits) and out_u8 (0/1 bytes) are optional, at least one must be given. */
int ullsam_op_resize_masks(const float* x, int n, int in_h, int in_w, int crop_h, int crop_w, int out_h, int out_w,
                           float threshold, float* out_f32, unsigned char* out_u8, void* stream);

/* CalcIoU (train_joint_v2.py:666-696): iou[i] = (|P & T| + smooth) / (|P | T| + smooth) with P = logits > threshold,
 * T = target > 0.5, over n planes of hw elements (fp32). */
int ullsam_op_mask_iou(const float* logits, const float* target, int n, size_t hw, float threshold, float smooth, float* iou,
                       void* stream);

/* fp32 -> bf16 (count elements). */
int ullsam_op_cast_bf16(const float* x, size_t count, void* y_bf16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ULLSAM_H_ */
