// ullsam_b200 — C ABI, model level: handle, weight repacking, encoder / decoder orchestration.
//
// The forward passes below are straight-line launch sequences on the caller's stream; every tensor-core
// step is the tcgen05 GEMM (gemm_sm100.cuh) or the fused attention (attention.cu), the rest are the
// coalesced glue kernels of elementwise.cu / decoder.cu.  Reference call stacks: SURVEY.md §3 (d), (e).
#include <cuda_runtime.h>

#include <functional>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ullsam.h"
#include "api_internal.h"
#include "attention.h"
#include "decoder.h"
#include "elementwise.h"
#include "gemm.h"
#include "host_util.h"

using namespace ullsam;
using bf16 = __nv_bfloat16;

namespace {

// dst (contiguous over dims d0..d3, bf16 or fp32) <- src[sum idx_i * stride_i] (fp32)
template <typename TOut>
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, TOut* __restrict__ dst, size_t total, int d1,
                                                   int d2, int d3, long s0, long s1, long s2, long s3) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    size_t r = i;
    const int i3 = static_cast<int>(r % d3); r /= d3;
    const int i2 = static_cast<int>(r % d2); r /= d2;
    const int i1 = static_cast<int>(r % d1); r /= d1;
    const float v = src[static_cast<long>(r) * s0 + i1 * s1 + i2 * s2 + i3 * s3];
    if constexpr (sizeof(TOut) == 2) dst[i] = __float2bfloat16(v);
    else dst[i] = v;
  }
}

struct Slot {
  std::vector<int64_t> shape;
  int which = 0;  // 0 encoder, 1 decoder
  bool loaded = false;
  std::function<const char*(const float*)> pack;  // src: fp32 device copy of the tensor
};

struct EncBlock {
  float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *qkv_b, *proj_b, *lin1_b, *lin2_b;
  bf16 *qkv_w, *proj_w, *lin1_w, *lin2_w, *relh, *relw;
  bool global;
  // LayerNorm folded into the projection behind it (gemm.h, EPI_*_FOLD): fp32 copies of the two weights, and what
  // fold_encoder_weights() derives from them and the norm's affine parameters
  float *qkv_w32, *lin1_w32;
  bf16 *qkv_wf, *lin1_wf;                   // bf16(gamma * W)
  float *qkv_s, *qkv_c, *lin1_s, *lin1_c;   // s_n = sum_k Wf_nk,  c_n = sum_k beta_k W_nk + b_n
};

// Wf = bf16(gamma * W), s = row sums of Wf (as rounded: the epilogue subtracts mean * s from an accumulator that was built
// from exactly these values), c = W beta + b (fp32 W).  One warp per output row.
__global__ void __launch_bounds__(256) fold_ln_kernel(const float* __restrict__ w32, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ bias, int N, int K,
                                                      bf16* __restrict__ wf, float* __restrict__ s_out, float* __restrict__ c_out) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float ss = 0.f, cc = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = w32[static_cast<size_t>(n) * K + k];
    const bf16 r = __float2bfloat16(gamma[k] * w);
    wf[static_cast<size_t>(n) * K + k] = r;
    ss += __bfloat162float(r);
    cc = fmaf(beta[k], w, cc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    cc += __shfl_xor_sync(0xffffffffu, cc, o);
  }
  if (lane == 0) {
    s_out[n] = ss;
    c_out[n] = cc + bias[n];
  }
}

struct DecAttnTok {  // token-side projections kept in fp32
  float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *o_w, *o_b;
};

struct DecLayer {
  DecAttnTok self;
  float *n1_w, *n1_b, *n2_w, *n2_b, *n3_w, *n3_b, *n4_w, *n4_b;
  float *t2i_q_w, *t2i_q_b, *t2i_o_w, *t2i_o_b;
  float *mlp1_w, *mlp1_b, *mlp2_w, *mlp2_b;
  float *i2t_k_w, *i2t_k_b, *i2t_v_w, *i2t_v_b, *i2t_o_b;
  bf16* i2t_o_w;  // [256,128]
  bf16 *cat_w, *catpe_w;  // [384,256]: rows 0..127 t2i.k, 128..255 t2i.v, 256..383 i2t.q ; catpe has V rows zero
  float* cat_b;           // [384]
};

}  // namespace

struct ullsam_handle {
  ullsam_config cfg{};
  int device = 0;
  int sms = 0;
  std::string err;
  bool alloc_failed = false;
  std::vector<void*> allocs;
  std::unordered_map<std::string, Slot> slots;
  // weight upload: copies and re-pack kernels run on the legacy default stream without a host sync per tensor; the first
  // forward call after an upload makes its stream wait for `weights_ev`
  bool weights_dirty = false;
  bool enc_fold_dirty = true;   // an encoder tensor changed since fold_encoder_weights() last ran
  int ln_fold = 1;              // LayerNorm folded into the GEMMs around it: bit 0 = Block.norm1 (mlp.lin2 / patch embedding ->
                                // attn.qkv), bit 1 = Block.norm2 (attn.proj -> mlp.lin1).  ULLSAM_LN_FOLD overrides; 0 = separate
                                // LayerNorm kernels.  Default 1: measured on B200 (tools/time_gemm_fold.py, ViT-B, batch 32) norm1's
                                // fold costs lin2 +15 us and qkv +23 us for a 90 us LayerNorm kernel, norm2's costs proj +75 us
                                // (HBM-bound: the extra bf16 copy is exposed) and lin1 +47 us (GELU epilogue, instruction-bound)
  cudaStream_t weights_stream = 0;
  cudaEvent_t weights_ev = nullptr;
  float* staging = nullptr;
  size_t staging_bytes = 0;
  void* staging_raw = nullptr;
  size_t staging_raw_bytes = 0;

  // encoder weights
  bf16* patch_w = nullptr;
  float *patch_b = nullptr, *pos = nullptr;
  std::vector<EncBlock> blocks;
  bf16 *neck0_w = nullptr, *neck2_w = nullptr;
  float *neck1_w = nullptr, *neck1_b = nullptr, *neck3_w = nullptr, *neck3_b = nullptr;

  // decoder weights
  std::vector<DecLayer> layers;
  float *fin_q_w = nullptr, *fin_q_b = nullptr, *fin_o_w = nullptr, *fin_o_b = nullptr, *nf_w = nullptr, *nf_b = nullptr;
  bf16 *fin_cat_w = nullptr, *fin_catpe_w = nullptr;  // [256,256]
  float* fin_cat_b = nullptr;
  float *iou_token = nullptr, *mask_tokens = nullptr;
  bf16 *up1_w = nullptr, *up2_w = nullptr;
  float *up1_b = nullptr, *up2_b = nullptr, *upln_w = nullptr, *upln_b = nullptr;
  float *hyp_w0 = nullptr, *hyp_b0 = nullptr, *hyp_w1 = nullptr, *hyp_b1 = nullptr, *hyp_w2 = nullptr, *hyp_b2 = nullptr;
  std::vector<float*> iou_w, iou_b;

  // workspaces
  int enc_cap = 0;
  float* x = nullptr;
  bf16 *q = nullptr, *k = nullptr, *v = nullptr, *ao = nullptr, *hid = nullptr;
  float* kmax = nullptr;  // [kEncChunk * heads] max key norm per (image, head): scratch of the global attention
  bf16* xn_t = nullptr;  // LayerNorm outputs in token order; with the folded LayerNorm: the bf16 copy of the residual stream
  float* row_stats = nullptr;  // [tok][kStatSlots][2] partial (sum, sum of squares) per row and 128-column slice
  float* emb_f32 = nullptr;  // [cap,256,4096] when the caller wants bf16 output / host staging
  int dec_cap_r = 0, dec_cap_nt = 0;
  float *tokens = nullptr, *queries = nullptr, *tq = nullptr, *tk = nullptr, *tv = nullptr, *ta = nullptr, *tmp = nullptr,
        *thid = nullptr, *q128 = nullptr, *k128 = nullptr, *v128 = nullptr, *o128 = nullptr;
  float *keys = nullptr, *hyp1 = nullptr, *hyp2 = nullptr, *hyper = nullptr, *iou1 = nullptr, *iou2 = nullptr,
        *iou4 = nullptr;
  bf16 *keys_bf = nullptr, *proj = nullptr, *ao128 = nullptr, *pe_bf = nullptr, *up_h = nullptr;
  std::vector<float*> peproj;  // per layer [4096,384], final [4096,256]
  // launch accounting + optional per-category CUDA-event profiling (ullsam_profile_*)
  unsigned long long launches = 0;
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_cat;
  size_t prof_n = 0;
  // host-call staging
  float *d_images = nullptr, *d_pe = nullptr, *d_sparse = nullptr, *d_dense = nullptr, *d_masks = nullptr, *d_iou = nullptr;
  int host_cap_b = 0, host_cap_ns = 0;
  cudaStream_t hs_copy = nullptr, hs_comp = nullptr;   // internal streams of ullsam_segment_host
  std::vector<cudaEvent_t> hs_ev;
};

namespace ullsam {

int fail(ullsam_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  g_thread_error = msg;
  return code;
}

}  // namespace ullsam

namespace {

enum ProfCat : int {
  PC_PATCH = 0, PC_LN, PC_GEMM_QKV, PC_ATTN_WIN, PC_ATTN_GLOBAL, PC_GEMM_PROJ, PC_GEMM_LIN1, PC_GEMM_LIN2, PC_NECK,
  PC_DEC_TOKENS, PC_DEC_GEMM, PC_DEC_ATTN, PC_DEC_UPSCALE, PC_COUNT
};
const char* kProfNames[PC_COUNT] = {"patch_embed", "layernorm", "gemm_qkv", "attn_windowed", "attn_global", "gemm_proj",
                                    "gemm_mlp_lin1", "gemm_mlp_lin2", "neck", "dec_token_kernels", "dec_image_gemms",
                                    "dec_cross_attention", "dec_upscale_mask"};

// called after every launch: counts kernels and, when profiling, drops a CUDA event on the stream
void mark(ullsam_handle* h, int cat, cudaStream_t st, int nkernels = 1) {
  h->launches += nkernels;
  if (!h->prof_on) return;
  if (h->prof_n == h->prof_ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    h->prof_ev.push_back(e);
  }
  cudaEventRecord(h->prof_ev[h->prof_n++], st);
  h->prof_cat.push_back(cat);
}

#define CKS(expr)                                                        \
  do {                                                                   \
    if (const char* _e = (expr)) return fail(h, ULLSAM_ERR_CUDA, _e);    \
  } while (0)
#define CKC(expr)                                                                                 \
  do {                                                                                            \
    cudaError_t _c = (expr);                                                                      \
    if (_c != cudaSuccess) return fail(h, ULLSAM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_c)); \
  } while (0)

// make the handle's device current for the rest of the scope, restoring the caller's device on exit
#define DEVICE_GUARD(h)                                                                        \
  DeviceGuard _dev_guard((h)->device);                                                         \
  if (_dev_guard.status != cudaSuccess)                                                        \
  return fail(h, ULLSAM_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(_dev_guard.status))

// order the caller's stream after any weight upload still in flight on the legacy default stream
int weights_fence(ullsam_handle* h, cudaStream_t st) {
  if (!h->weights_dirty) return ULLSAM_OK;
  if (!h->weights_ev) CKC(cudaEventCreateWithFlags(&h->weights_ev, cudaEventDisableTiming));
  if (st != h->weights_stream) {
    CKC(cudaEventRecord(h->weights_ev, h->weights_stream));
    CKC(cudaStreamWaitEvent(st, h->weights_ev, 0));
  }
  h->weights_dirty = false;
  return ULLSAM_OK;
}

// "is this environment switch set to 0?"
bool getenv_off(const char* name) {
  const char* v = getenv(name);
  return v != nullptr && atoi(v) == 0;
}

template <typename T>
T* dalloc(ullsam_handle* h, size_t count, bool zero = false) {
  void* p = nullptr;
  if (cudaMalloc(&p, count * sizeof(T) + 256) != cudaSuccess) {
    h->alloc_failed = true;
    return nullptr;
  }
  if (zero) cudaMemset(p, 0, count * sizeof(T) + 256);
  h->allocs.push_back(p);
  return static_cast<T*>(p);
}

void dfree(ullsam_handle* h, void* p) {
  if (!p) return;
  for (auto& a : h->allocs)
    if (a == p) { a = h->allocs.back(); h->allocs.pop_back(); break; }
  cudaFree(p);
}

// stream the re-pack kernels of the weight upload in progress run on (set by set_weight_impl around Slot::pack)
thread_local cudaStream_t g_pack_stream = 0;

template <typename TOut>
const char* pack(const float* src, TOut* dst, int d0, int d1, int d2, int d3, long s0, long s1, long s2, long s3) {
  const size_t total = static_cast<size_t>(d0) * d1 * d2 * d3;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_kernel<TOut><<<static_cast<int>(blocks), 256, 0, g_pack_stream>>>(src, dst, total, d1, d2, d3, s0, s1, s2, s3);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// register a tensor that is copied 1:1 (optionally to bf16) into a fresh device buffer (or a sub-range of one)
template <typename TOut>
void reg_plain(ullsam_handle* h, const std::string& key, std::vector<int64_t> shape, int which, TOut** dst_field,
               TOut* into = nullptr) {
  size_t n = 1;
  for (auto s : shape) n *= static_cast<size_t>(s);
  TOut* dst = into ? into : dalloc<TOut>(h, n);
  if (dst_field) *dst_field = dst;
  Slot s;
  s.shape = std::move(shape);
  s.which = which;
  s.pack = [dst, n](const float* src) { return pack<TOut>(src, dst, 1, 1, 1, static_cast<int>(n), 0, 0, 0, 1); };
  h->slots[key] = std::move(s);
}

void reg_custom(ullsam_handle* h, const std::string& key, std::vector<int64_t> shape, int which,
                std::function<const char*(const float*)> fn) {
  Slot s;
  s.shape = std::move(shape);
  s.which = which;
  s.pack = std::move(fn);
  h->slots[key] = std::move(s);
}

// a projection weight that is kept twice: bf16 (the GEMM operand of the unfolded path) and fp32 (source of the LayerNorm fold)
void reg_w_with_f32(ullsam_handle* h, const std::string& key, int64_t N, int64_t K, bf16** w_bf16, float** w_f32) {
  const size_t n = static_cast<size_t>(N) * K;
  bf16* d0 = dalloc<bf16>(h, n);
  float* d1 = dalloc<float>(h, n);
  *w_bf16 = d0;
  *w_f32 = d1;
  reg_custom(h, key, {N, K}, 0, [d0, d1, n](const float* src) -> const char* {
    if (const char* e = pack<bf16>(src, d0, 1, 1, 1, static_cast<int>(n), 0, 0, 0, 1)) return e;
    return pack<float>(src, d1, 1, 1, 1, static_cast<int>(n), 0, 0, 0, 1);
  });
}

bool is_global(const ullsam_config& c, int i) {
  for (int j = 0; j < c.num_global; ++j)
    if (c.global_attn_indexes[j] == i) return true;
  return false;
}

void register_encoder(ullsam_handle* h) {
  const auto& c = h->cfg;
  const int64_t D = c.embed_dim, M = c.mlp_dim, g = c.img_size / c.patch_size, pk = c.in_chans * c.patch_size * c.patch_size;
  const std::string p = "image_encoder.";
  reg_plain<bf16>(h, p + "patch_embed.proj.weight", {D, c.in_chans, c.patch_size, c.patch_size}, 0, &h->patch_w);
  reg_plain<float>(h, p + "patch_embed.proj.bias", {D}, 0, &h->patch_b);
  reg_plain<float>(h, p + "pos_embed", {1, g, g, D}, 0, &h->pos);
  (void)pk;
  h->blocks.resize(c.depth);
  for (int i = 0; i < c.depth; ++i) {
    EncBlock& b = h->blocks[i];
    b.global = is_global(c, i);
    const std::string bp = p + "blocks." + std::to_string(i) + ".";
    reg_plain<float>(h, bp + "norm1.weight", {D}, 0, &b.ln1_w);
    reg_plain<float>(h, bp + "norm1.bias", {D}, 0, &b.ln1_b);
    reg_plain<float>(h, bp + "norm2.weight", {D}, 0, &b.ln2_w);
    reg_plain<float>(h, bp + "norm2.bias", {D}, 0, &b.ln2_b);
    reg_w_with_f32(h, bp + "attn.qkv.weight", 3 * D, D, &b.qkv_w, &b.qkv_w32);
    b.qkv_wf = dalloc<bf16>(h, 3 * D * D);
    b.qkv_s = dalloc<float>(h, 3 * D);
    b.qkv_c = dalloc<float>(h, 3 * D);
    reg_plain<float>(h, bp + "attn.qkv.bias", {3 * D}, 0, &b.qkv_b);
    reg_plain<bf16>(h, bp + "attn.proj.weight", {D, D}, 0, &b.proj_w);
    reg_plain<float>(h, bp + "attn.proj.bias", {D}, 0, &b.proj_b);
    reg_w_with_f32(h, bp + "mlp.lin1.weight", M, D, &b.lin1_w, &b.lin1_w32);
    b.lin1_wf = dalloc<bf16>(h, M * D);
    b.lin1_s = dalloc<float>(h, M);
    b.lin1_c = dalloc<float>(h, M);
    reg_plain<float>(h, bp + "mlp.lin1.bias", {M}, 0, &b.lin1_b);
    reg_plain<bf16>(h, bp + "mlp.lin2.weight", {D, M}, 0, &b.lin2_w);
    reg_plain<float>(h, bp + "mlp.lin2.bias", {D}, 0, &b.lin2_b);
    const int64_t size = b.global ? g : c.window_size;
    const int nrel = b.global ? 128 : 32;
    b.relh = dalloc<bf16>(h, nrel * 64, true);
    b.relw = dalloc<bf16>(h, nrel * 64, true);
    reg_plain<bf16>(h, bp + "attn.rel_pos_h", {2 * size - 1, 64}, 0, nullptr, b.relh);
    reg_plain<bf16>(h, bp + "attn.rel_pos_w", {2 * size - 1, 64}, 0, nullptr, b.relw);
  }
  const int64_t C = c.out_chans;
  reg_plain<bf16>(h, p + "neck.0.weight", {C, D, 1, 1}, 0, &h->neck0_w);
  reg_plain<float>(h, p + "neck.1.weight", {C}, 0, &h->neck1_w);
  reg_plain<float>(h, p + "neck.1.bias", {C}, 0, &h->neck1_b);
  h->neck2_w = dalloc<bf16>(h, C * C * 9);
  {
    bf16* dst = h->neck2_w;
    const int Ci = static_cast<int>(C);
    // [o][i][ky][kx] -> [o][ky][kx][i]  (K index of the implicit GEMM = tap * C + channel)
    reg_custom(h, p + "neck.2.weight", {C, C, 3, 3}, 0, [dst, Ci](const float* src) {
      return pack<bf16>(src, dst, Ci, 3, 3, Ci, static_cast<long>(Ci) * 9, 3, 1, 9);
    });
  }
  reg_plain<float>(h, p + "neck.3.weight", {C}, 0, &h->neck3_w);
  reg_plain<float>(h, p + "neck.3.bias", {C}, 0, &h->neck3_b);
}

void reg_attn_tok(ullsam_handle* h, const std::string& ap, int internal, DecAttnTok* a) {
  const int64_t E = 256;
  reg_plain<float>(h, ap + "q_proj.weight", {internal, E}, 1, &a->q_w);
  reg_plain<float>(h, ap + "q_proj.bias", {internal}, 1, &a->q_b);
  reg_plain<float>(h, ap + "k_proj.weight", {internal, E}, 1, &a->k_w);
  reg_plain<float>(h, ap + "k_proj.bias", {internal}, 1, &a->k_b);
  reg_plain<float>(h, ap + "v_proj.weight", {internal, E}, 1, &a->v_w);
  reg_plain<float>(h, ap + "v_proj.bias", {internal}, 1, &a->v_b);
  reg_plain<float>(h, ap + "out_proj.weight", {E, internal}, 1, &a->o_w);
  reg_plain<float>(h, ap + "out_proj.bias", {E}, 1, &a->o_b);
}

// one state_dict tensor feeding two packed destinations (cat_w and catpe_w rows)
void reg_dual_bf16(ullsam_handle* h, const std::string& key, std::vector<int64_t> shape, bf16* d0, bf16* d1) {
  size_t n = 1;
  for (auto s : shape) n *= static_cast<size_t>(s);
  reg_custom(h, key, std::move(shape), 1, [d0, d1, n](const float* src) -> const char* {
    if (const char* e = pack<bf16>(src, d0, 1, 1, 1, static_cast<int>(n), 0, 0, 0, 1)) return e;
    if (d1) return pack<bf16>(src, d1, 1, 1, 1, static_cast<int>(n), 0, 0, 0, 1);
    return nullptr;
  });
}

void register_decoder(ullsam_handle* h) {
  const auto& c = h->cfg;
  const int64_t E = 256, I = 128, MD = c.dec_mlp_dim;
  const std::string d = "mask_decoder.", t = "mask_decoder.transformer.";
  h->layers.resize(c.dec_depth);
  for (int i = 0; i < c.dec_depth; ++i) {
    DecLayer& L = h->layers[i];
    const std::string lp = t + "layers." + std::to_string(i) + ".";
    reg_attn_tok(h, lp + "self_attn.", 256, &L.self);
    reg_plain<float>(h, lp + "norm1.weight", {E}, 1, &L.n1_w);
    reg_plain<float>(h, lp + "norm1.bias", {E}, 1, &L.n1_b);
    reg_plain<float>(h, lp + "norm2.weight", {E}, 1, &L.n2_w);
    reg_plain<float>(h, lp + "norm2.bias", {E}, 1, &L.n2_b);
    reg_plain<float>(h, lp + "norm3.weight", {E}, 1, &L.n3_w);
    reg_plain<float>(h, lp + "norm3.bias", {E}, 1, &L.n3_b);
    reg_plain<float>(h, lp + "norm4.weight", {E}, 1, &L.n4_w);
    reg_plain<float>(h, lp + "norm4.bias", {E}, 1, &L.n4_b);
    L.cat_w = dalloc<bf16>(h, 384 * 256, true);
    L.catpe_w = dalloc<bf16>(h, 384 * 256, true);
    L.cat_b = dalloc<float>(h, 384, true);
    const std::string a = lp + "cross_attn_token_to_image.";
    reg_plain<float>(h, a + "q_proj.weight", {I, E}, 1, &L.t2i_q_w);
    reg_plain<float>(h, a + "q_proj.bias", {I}, 1, &L.t2i_q_b);
    reg_dual_bf16(h, a + "k_proj.weight", {I, E}, L.cat_w, L.catpe_w);
    reg_plain<float>(h, a + "k_proj.bias", {I}, 1, nullptr, L.cat_b);
    reg_dual_bf16(h, a + "v_proj.weight", {I, E}, L.cat_w + 128 * 256, nullptr);
    reg_plain<float>(h, a + "v_proj.bias", {I}, 1, nullptr, L.cat_b + 128);
    reg_plain<float>(h, a + "out_proj.weight", {E, I}, 1, &L.t2i_o_w);
    reg_plain<float>(h, a + "out_proj.bias", {E}, 1, &L.t2i_o_b);
    reg_plain<float>(h, lp + "mlp.lin1.weight", {MD, E}, 1, &L.mlp1_w);
    reg_plain<float>(h, lp + "mlp.lin1.bias", {MD}, 1, &L.mlp1_b);
    reg_plain<float>(h, lp + "mlp.lin2.weight", {E, MD}, 1, &L.mlp2_w);
    reg_plain<float>(h, lp + "mlp.lin2.bias", {E}, 1, &L.mlp2_b);
    const std::string b = lp + "cross_attn_image_to_token.";
    reg_dual_bf16(h, b + "q_proj.weight", {I, E}, L.cat_w + 256 * 256, L.catpe_w + 256 * 256);
    reg_plain<float>(h, b + "q_proj.bias", {I}, 1, nullptr, L.cat_b + 256);
    reg_plain<float>(h, b + "k_proj.weight", {I, E}, 1, &L.i2t_k_w);
    reg_plain<float>(h, b + "k_proj.bias", {I}, 1, &L.i2t_k_b);
    reg_plain<float>(h, b + "v_proj.weight", {I, E}, 1, &L.i2t_v_w);
    reg_plain<float>(h, b + "v_proj.bias", {I}, 1, &L.i2t_v_b);
    reg_plain<bf16>(h, b + "out_proj.weight", {E, I}, 1, &L.i2t_o_w);
    reg_plain<float>(h, b + "out_proj.bias", {E}, 1, &L.i2t_o_b);
  }
  {
    const std::string a = t + "final_attn_token_to_image.";
    h->fin_cat_w = dalloc<bf16>(h, 256 * 256, true);
    h->fin_catpe_w = dalloc<bf16>(h, 256 * 256, true);
    h->fin_cat_b = dalloc<float>(h, 256, true);
    reg_plain<float>(h, a + "q_proj.weight", {I, E}, 1, &h->fin_q_w);
    reg_plain<float>(h, a + "q_proj.bias", {I}, 1, &h->fin_q_b);
    reg_dual_bf16(h, a + "k_proj.weight", {I, E}, h->fin_cat_w, h->fin_catpe_w);
    reg_plain<float>(h, a + "k_proj.bias", {I}, 1, nullptr, h->fin_cat_b);
    reg_dual_bf16(h, a + "v_proj.weight", {I, E}, h->fin_cat_w + 128 * 256, nullptr);
    reg_plain<float>(h, a + "v_proj.bias", {I}, 1, nullptr, h->fin_cat_b + 128);
    reg_plain<float>(h, a + "out_proj.weight", {E, I}, 1, &h->fin_o_w);
    reg_plain<float>(h, a + "out_proj.bias", {E}, 1, &h->fin_o_b);
    reg_plain<float>(h, t + "norm_final_attn.weight", {E}, 1, &h->nf_w);
    reg_plain<float>(h, t + "norm_final_attn.bias", {E}, 1, &h->nf_b);
  }
  const int64_t nm = c.num_mask_tokens;
  reg_plain<float>(h, d + "iou_token.weight", {1, E}, 1, &h->iou_token);
  reg_plain<float>(h, d + "mask_tokens.weight", {nm, E}, 1, &h->mask_tokens);
  // ConvTranspose2d(256->64,k2,s2): W[ci][co][ky][kx] -> GEMM weight [(ky,kx,co)][ci]; bias tiled over (ky,kx)
  h->up1_w = dalloc<bf16>(h, 256 * 256);
  h->up1_b = dalloc<float>(h, 256);
  h->up2_w = dalloc<bf16>(h, 128 * 64);
  h->up2_b = dalloc<float>(h, 128);
  {
    bf16* w1 = h->up1_w; float* b1 = h->up1_b; bf16* w2 = h->up2_w; float* b2 = h->up2_b;
    reg_custom(h, d + "output_upscaling.0.weight", {256, 64, 2, 2}, 1,
               [w1](const float* s) { return pack<bf16>(s, w1, 2, 2, 64, 256, 2, 1, 4, 256); });
    reg_custom(h, d + "output_upscaling.0.bias", {64}, 1, [b1](const float* s) { return pack<float>(s, b1, 1, 1, 4, 64, 0, 0, 0, 1); });
    reg_custom(h, d + "output_upscaling.3.weight", {64, 32, 2, 2}, 1,
               [w2](const float* s) { return pack<bf16>(s, w2, 2, 2, 32, 64, 2, 1, 4, 128); });
    reg_custom(h, d + "output_upscaling.3.bias", {32}, 1, [b2](const float* s) { return pack<float>(s, b2, 1, 1, 4, 32, 0, 0, 0, 1); });
  }
  reg_plain<float>(h, d + "output_upscaling.1.weight", {64}, 1, &h->upln_w);
  reg_plain<float>(h, d + "output_upscaling.1.bias", {64}, 1, &h->upln_b);
  // hypernetwork MLPs: the 4 MLPs' tensors are stored back to back so one grouped launch covers them
  h->hyp_w0 = dalloc<float>(h, nm * 256 * 256); h->hyp_b0 = dalloc<float>(h, nm * 256);
  h->hyp_w1 = dalloc<float>(h, nm * 256 * 256); h->hyp_b1 = dalloc<float>(h, nm * 256);
  h->hyp_w2 = dalloc<float>(h, nm * 32 * 256);  h->hyp_b2 = dalloc<float>(h, nm * 32);
  for (int i = 0; i < nm; ++i) {
    const std::string hp = d + "output_hypernetworks_mlps." + std::to_string(i) + ".layers.";
    reg_plain<float>(h, hp + "0.weight", {E, E}, 1, nullptr, h->hyp_w0 + static_cast<size_t>(i) * 256 * 256);
    reg_plain<float>(h, hp + "0.bias", {E}, 1, nullptr, h->hyp_b0 + i * 256);
    reg_plain<float>(h, hp + "1.weight", {E, E}, 1, nullptr, h->hyp_w1 + static_cast<size_t>(i) * 256 * 256);
    reg_plain<float>(h, hp + "1.bias", {E}, 1, nullptr, h->hyp_b1 + i * 256);
    reg_plain<float>(h, hp + "2.weight", {32, E}, 1, nullptr, h->hyp_w2 + static_cast<size_t>(i) * 32 * 256);
    reg_plain<float>(h, hp + "2.bias", {32}, 1, nullptr, h->hyp_b2 + i * 32);
  }
  h->iou_w.resize(c.iou_head_depth);
  h->iou_b.resize(c.iou_head_depth);
  for (int i = 0; i < c.iou_head_depth; ++i) {
    const int64_t in = (i == 0) ? E : c.iou_head_hidden;
    const int64_t out = (i == c.iou_head_depth - 1) ? nm : c.iou_head_hidden;
    const std::string ip = d + "iou_prediction_head.layers." + std::to_string(i) + ".";
    reg_plain<float>(h, ip + "weight", {out, in}, 1, &h->iou_w[i]);
    reg_plain<float>(h, ip + "bias", {out}, 1, &h->iou_b[i]);
  }
  // class 1 = TwoWayTransformer tensors, class 2 = the rest of MaskDecoder (a stand-alone transformer needs only 1)
  for (auto& kv : h->slots)
    if (kv.second.which == 1 && kv.first.rfind("mask_decoder.transformer.", 0) != 0) kv.second.which = 2;
}

int check_ready(ullsam_handle* h, int which) {
  int missing = 0;
  std::string first;
  for (auto& kv : h->slots)
    if (kv.second.which == which && !kv.second.loaded) {
      if (!missing) first = kv.first;
      ++missing;
    }
  if (missing)
    return fail(h, ULLSAM_ERR_STATE, std::to_string(missing) + " state_dict tensors not loaded (e.g. " + first + ")");
  return ULLSAM_OK;
}

int ensure_encoder_ws(ullsam_handle* h, int B) {
  if (B <= h->enc_cap) return ULLSAM_OK;
  for (void* p : {(void*)h->x, (void*)h->xn_t, (void*)h->kmax, (void*)h->q, (void*)h->k, (void*)h->v, (void*)h->ao, (void*)h->hid,
                  (void*)h->emb_f32, (void*)h->row_stats})
    dfree(h, p);
  const size_t D = h->cfg.embed_dim, M = h->cfg.mlp_dim;
  const size_t tokw = static_cast<size_t>(B) * 4900, tok = static_cast<size_t>(B) * 4096;
  h->x = dalloc<float>(h, tok * D);
  h->xn_t = dalloc<bf16>(h, tok * D);
  h->kmax = dalloc<float>(h, static_cast<size_t>(B) * h->cfg.num_heads);
  h->q = dalloc<bf16>(h, tokw * D);
  h->k = dalloc<bf16>(h, tokw * D);
  h->v = dalloc<bf16>(h, tokw * D);
  h->ao = dalloc<bf16>(h, tok * D);
  h->hid = dalloc<bf16>(h, tok * (M > 768 ? M : 768));
  h->emb_f32 = dalloc<float>(h, tok * 256);
  h->row_stats = dalloc<float>(h, tok * kStatSlots * 2);
  if (!h->row_stats || !h->x || !h->xn_t || !h->kmax || !h->q || !h->k || !h->v || !h->ao || !h->hid || !h->emb_f32) {
    h->enc_cap = 0;
    return fail(h, ULLSAM_ERR_NOMEM, "cudaMalloc failed for the encoder workspace");
  }
  // q rows of the window-padding tokens are never written (their outputs are cropped, image_encoder.py:267-289, and
  // attention rows are independent): give them defined contents once
  if (cudaMemset(h->q, 0, tokw * D * sizeof(bf16)) != cudaSuccess) {
    h->enc_cap = 0;
    return fail(h, ULLSAM_ERR_CUDA, "cudaMemset failed for the encoder workspace");
  }
  h->enc_cap = B;
  return ULLSAM_OK;
}

constexpr int kEncChunk = 32;  // images per pass through the encoder (workspace ~75 MB per image for ViT-B)

int ensure_decoder_ws(ullsam_handle* h, int R, int Nt) {
  if (R <= h->dec_cap_r && Nt <= h->dec_cap_nt) return ULLSAM_OK;
  if (R < h->dec_cap_r) R = h->dec_cap_r;
  if (Nt < h->dec_cap_nt) Nt = h->dec_cap_nt;
  for (void* p : {(void*)h->tokens, (void*)h->queries, (void*)h->tq, (void*)h->tk, (void*)h->tv, (void*)h->ta, (void*)h->tmp,
                  (void*)h->thid, (void*)h->q128, (void*)h->k128, (void*)h->v128, (void*)h->o128, (void*)h->keys,
                  (void*)h->hyp1, (void*)h->hyp2, (void*)h->hyper, (void*)h->iou1, (void*)h->iou2, (void*)h->iou4,
                  (void*)h->keys_bf, (void*)h->proj, (void*)h->ao128, (void*)h->pe_bf, (void*)h->up_h})
    dfree(h, p);
  for (float* p : h->peproj) dfree(h, p);
  h->peproj.clear();
  const size_t rt = static_cast<size_t>(R) * Nt, rk = static_cast<size_t>(R) * 4096;
  h->tokens = dalloc<float>(h, rt * 256); h->queries = dalloc<float>(h, rt * 256);
  h->tq = dalloc<float>(h, rt * 256); h->tk = dalloc<float>(h, rt * 256); h->tv = dalloc<float>(h, rt * 256);
  h->ta = dalloc<float>(h, rt * 256); h->tmp = dalloc<float>(h, rt * 256);
  h->thid = dalloc<float>(h, rt * h->cfg.dec_mlp_dim);
  h->q128 = dalloc<float>(h, rt * 128); h->k128 = dalloc<float>(h, rt * 128); h->v128 = dalloc<float>(h, rt * 128);
  h->o128 = dalloc<float>(h, rt * 128);
  h->keys = dalloc<float>(h, rk * 256);
  h->hyp1 = dalloc<float>(h, static_cast<size_t>(R) * 4 * 256); h->hyp2 = dalloc<float>(h, static_cast<size_t>(R) * 4 * 256);
  h->hyper = dalloc<float>(h, static_cast<size_t>(R) * 4 * 32);
  h->iou1 = dalloc<float>(h, static_cast<size_t>(R) * 256); h->iou2 = dalloc<float>(h, static_cast<size_t>(R) * 256);
  h->iou4 = dalloc<float>(h, static_cast<size_t>(R) * 4);
  h->keys_bf = dalloc<bf16>(h, rk * 256);
  h->proj = dalloc<bf16>(h, rk * 384);
  h->ao128 = dalloc<bf16>(h, rk * 128);
  h->pe_bf = dalloc<bf16>(h, 4096 * 256);
  h->up_h = dalloc<bf16>(h, static_cast<size_t>(R) * 16384 * 64);
  bool ok = h->tokens && h->queries && h->tq && h->tk && h->tv && h->ta && h->tmp && h->thid && h->q128 && h->k128 &&
            h->v128 && h->o128 && h->keys && h->hyp1 && h->hyp2 && h->hyper && h->iou1 && h->iou2 && h->iou4 &&
            h->keys_bf && h->proj && h->ao128 && h->pe_bf && h->up_h;
  for (int i = 0; i <= h->cfg.dec_depth; ++i) {
    h->peproj.push_back(dalloc<float>(h, 4096 * 384));
    ok = ok && h->peproj.back();
  }
  if (!ok) {
    h->dec_cap_r = h->dec_cap_nt = 0;
    return fail(h, ULLSAM_ERR_NOMEM, "cudaMalloc failed for the decoder workspace");
  }
  h->dec_cap_r = R;
  h->dec_cap_nt = Nt;
  return ULLSAM_OK;
}

GemmParams gp(int M, int N, int K) {
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.ldc = N;
  return p;
}

// Block.norm1 / norm2 folded into attn.qkv / mlp.lin1 (gemm.h): re-derived whenever an encoder tensor has been uploaded.
// Runs on the stream of the forward call that finds the flag set and is drained before that call goes on, so forward
// calls on other streams see the folded weights too (uploads are rare; one host sync per upload batch).
int fold_encoder_weights(ullsam_handle* h, cudaStream_t st) {
  if (!h->enc_fold_dirty || !h->ln_fold) return ULLSAM_OK;
  const int D = h->cfg.embed_dim, MD = h->cfg.mlp_dim;
  for (EncBlock& b : h->blocks) {
    fold_ln_kernel<<<(3 * D + 7) / 8, 256, 0, st>>>(b.qkv_w32, b.ln1_w, b.ln1_b, b.qkv_b, 3 * D, D, b.qkv_wf, b.qkv_s, b.qkv_c);
    fold_ln_kernel<<<(MD + 7) / 8, 256, 0, st>>>(b.lin1_w32, b.ln2_w, b.ln2_b, b.lin1_b, MD, D, b.lin1_wf, b.lin1_s, b.lin1_c);
  }
  CKC(cudaGetLastError());
  CKC(cudaStreamSynchronize(st));
  h->enc_fold_dirty = false;
  return ULLSAM_OK;
}

// ---------------------------------------------------------------------------------------------------
// ImageEncoderViT.forward for Bc <= kEncChunk images (image_encoder.py:106-116)
// ---------------------------------------------------------------------------------------------------
int encoder_chunk(ullsam_handle* h, const void* images, int dtype, int Bc, float* out_f32, cudaStream_t st) {
  const auto& c = h->cfg;
  const int D = c.embed_dim, MD = c.mlp_dim, heads = c.num_heads, sms = h->sms;
  const int tok = Bc * 4096;
  // LayerNorm folded into the GEMMs on either side of it (default): every GEMM that writes the residual stream (patch
  // embedding, attn.proj, mlp.lin2) also writes its bf16 copy and per-row partial sums, and attn.qkv / mlp.lin1 run on that
  // copy with gamma-scaled weights and finish the normalisation in their epilogue — no LayerNorm kernel, no cast before
  // the neck.  (D % 256 == 0 and D <= 1024: ViT-B and ViT-L.)
  // ULLSAM_KMAX_EPI=0: the global attention takes its key-norm bound from a pre-pass over K instead of the QKV epilogue
  static const bool kmax_epi = !getenv_off("ULLSAM_KMAX_EPI");
  static const bool proj_red = !getenv_off("ULLSAM_PROJ_RED");
  const bool foldable = D % 256 == 0 && D <= 128 * kStatSlots;
  const bool fold1 = foldable && (h->ln_fold & 1), fold2 = foldable && (h->ln_fold & 2);
  if (fold1 || fold2)
    if (int rc = fold_encoder_weights(h, st)) return rc;
  // PatchEmbed (+ pos_embed)
  CKS(launch_im2col16(images, dtype == ULLSAM_BF16, Bc, 3, 1024, h->hid, st));
  {
    GemmParams p = gp(tok, D, 768);
    p.bias = h->patch_b; p.out = h->x; p.pos = h->pos; p.pos_rows = 4096;
    if (fold1) { p.out2 = h->xn_t; p.row_stats = h->row_stats; }
    CKS(launch_gemm(fold1 ? EPI_PATCH_STATS : EPI_PATCH, h->hid, h->patch_w, p, sms, st));
    mark(h, PC_PATCH, st, 2);
  }
  // The LayerNorms that follow a residual add (norm2 after the attention branch, the next block's norm1 after the MLP
  // branch) CAN ride on the epilogue of the GEMM that performs the add (EPI_RES_LN).  Measured on B200 (ViT-B, batch
  // 32) it does not pay: the epilogue has to re-read the rows through L2 once their statistics are complete, those
  // 403 MB per GEMM arrive in synchronized bursts that stall the tensor pipe, and lin2 / proj lose more (+155 / +135 us
  // per launch) than the LayerNorm kernel they replace costs (97 us at 96 % of the HBM roofline).  So the default is
  // one LayerNorm kernel per norm; ULLSAM_LN_FUSION=1 (norm1 on lin2) / =3 (also norm2 on proj) select the fused path.
  static const int fuse_env = getenv("ULLSAM_LN_FUSION") ? atoi(getenv("ULLSAM_LN_FUSION")) : 0;
  // (the fused epilogue is built for 256-wide N tiles of at most 1024 columns: ViT-B and ViT-L)
  const bool fusable = D % 256 == 0 && D <= 1024;
  const bool fuse_ln1 = !fold1 && fusable && (fuse_env & 1), fuse_ln2 = !fold2 && fusable && (fuse_env & 2);
  for (int i = 0; i < c.depth; ++i) {
    const EncBlock& b = h->blocks[i];
    const int win = b.global ? 0 : c.window_size;
    const int T = b.global ? 4096 : 196;
    // norm1 output stays in token order for windowed blocks too: window_partition (image_encoder.py:168-172, 243-264) is
    // done by the QKV epilogue on the store, so the GEMM runs on the 4096 image tokens instead of the 4900 padded ones
    // (-16 % of its rows); the constant k / v rows of the 804 padding tokens (= qkv.bias) ride on the same epilogue.
    if (!fold1 && (i == 0 || !fuse_ln1)) {
      CKS(launch_layernorm(h->x, b.ln1_w, b.ln1_b, c.enc_ln_eps, Bc, 64, D, 0, h->xn_t, st));
      mark(h, PC_LN, st);
    }
    {  // (norm1 +) qkv + head split (+ window_partition)
      GemmParams p = gp(tok, 3 * D, D);
      p.bias = b.qkv_b; p.q = h->q; p.k = h->k; p.v = h->v; p.T = T; p.heads = heads; p.ln_partition = win ? 1 : 0;
      if (fold1) {
        p.bias = b.qkv_c; p.fold_s = b.qkv_s; p.pad_bias = b.qkv_b; p.row_stats = h->row_stats; p.ln_eps = c.enc_ln_eps;
      }
      if (b.global && kmax_epi) {   // max |k|^2 per (image, head) for the global attention's score bound rides on this epilogue
        CKC(cudaMemsetAsync(h->kmax, 0, static_cast<size_t>(Bc) * heads * sizeof(float), st));
        p.kmax_sq = h->kmax;
      }
      CKS(launch_gemm(fold1 ? EPI_QKV_FOLD : EPI_QKV, h->xn_t, fold1 ? b.qkv_wf : b.qkv_w, p, sms, st));
      mark(h, PC_GEMM_QKV, st);
    }
    // attention (+ rel-pos bias, + window_unpartition)
    CKS(launch_attention(h->q, h->k, h->v, b.relh, b.relw, b.global ? Bc : Bc * 25, heads, b.global ? 64 : 14, h->ao, st, h->kmax,
                         b.global && kmax_epi));
    mark(h, b.global ? PC_ATTN_GLOBAL : PC_ATTN_WIN, st, (b.global && !kmax_epi) ? 2 : 1);  // (+ the key-norm pre-pass)
    {  // proj + residual (+ norm2)
      GemmParams p = gp(tok, D, D);
      p.bias = b.proj_b; p.out = h->x; p.residual = h->x;
      if (fuse_ln2) {
        p.ln_w = b.ln2_w; p.ln_b = b.ln2_b; p.ln_eps = c.enc_ln_eps; p.out2 = h->xn_t; p.ln_partition = 0;
      }
      if (fold2) { p.out2 = h->xn_t; p.row_stats = h->row_stats; }
      // (plain residual add: done by the memory system through TMA reduce-add stores, EPI_RES_RED; ULLSAM_PROJ_RED=0: in the SM.
      //  ViT-B only: at D = 1024 the variant's four pipeline stages cost more than its epilogue saves — proj 350 vs 315 us)
      const int proj_epi = fold2 ? EPI_RES_STATS : (fuse_ln2 ? EPI_RES_LN : ((proj_red && D % 256 == 0 && D <= 768) ? EPI_RES_RED : EPI_RES_F32));
      CKS(launch_gemm(proj_epi, h->ao, b.proj_w, p, sms, st));
      mark(h, PC_GEMM_PROJ, st);
    }
    if (!fold2 && !fuse_ln2) {
      CKS(launch_layernorm(h->x, b.ln2_w, b.ln2_b, c.enc_ln_eps, Bc, 64, D, 0, h->xn_t, st));
      mark(h, PC_LN, st);
    }
    {  // MLP
      GemmParams p = gp(tok, MD, D);
      p.bias = b.lin1_b; p.out = h->hid; p.act = ACT_GELU;
      if (fold2) { p.bias = b.lin1_c; p.fold_s = b.lin1_s; p.row_stats = h->row_stats; p.ln_eps = c.enc_ln_eps; }
      CKS(launch_gemm(fold2 ? EPI_GELU_FOLD : EPI_BF16, h->xn_t, fold2 ? b.lin1_wf : b.lin1_w, p, sms, st));
      mark(h, PC_GEMM_LIN1, st);
    }
    {  // lin2 + residual (+ the next block's norm1 and window_partition)
      GemmParams p = gp(tok, D, MD);
      p.bias = b.lin2_b; p.out = h->x; p.residual = h->x;
      const bool fuse_next = fuse_ln1 && i + 1 < c.depth;
      if (fuse_next) {
        const EncBlock& nb = h->blocks[i + 1];
        p.ln_w = nb.ln1_w; p.ln_b = nb.ln1_b; p.ln_eps = c.enc_ln_eps;
        p.out2 = h->xn_t; p.ln_partition = 0;
      }
      if (fold1) { p.out2 = h->xn_t; p.row_stats = h->row_stats; }   // (the last block's copy feeds the neck)
      CKS(launch_gemm(fold1 ? EPI_RES_STATS : (fuse_next ? EPI_RES_LN : EPI_RES_F32), h->hid, b.lin2_w, p, sms, st));
      mark(h, PC_GEMM_LIN2, st);
    }
  }
  // neck: 1x1 conv -> LayerNorm2d -> 3x3 conv -> LayerNorm2d, NCHW out
  if (!fold1) CKS(launch_cast_f32_bf16(h->x, static_cast<size_t>(tok) * D, h->xn_t, st));
  {
    GemmParams p = gp(tok, 256, D);
    p.out = h->q; p.ln_w = h->neck1_w; p.ln_b = h->neck1_b; p.ln_eps = 1e-6f;
    CKS(launch_gemm(EPI_LN_BF16, h->xn_t, h->neck0_w, p, sms, st));
  }
  {
    GemmParams p = gp(tok, 256, 9 * 256);
    p.a_conv3 = 1; p.conv_c = 256;
    p.out = out_f32; p.ln_w = h->neck3_w; p.ln_b = h->neck3_b; p.ln_eps = 1e-6f; p.hw = 4096;
    CKS(launch_gemm(EPI_LN_NCHW, h->q, h->neck2_w, p, sms, st));
    mark(h, PC_NECK, st, fold1 ? 2 : 3);
  }
  return ULLSAM_OK;
}

TokLinear lin(const float* x, const float* x2, const float* w, const float* b, const float* res, float* y, int M, int N,
              int K, int act = 0) {
  TokLinear a{};
  a.x = x; a.x2 = x2; a.w = w; a.b = b; a.res = res; a.y = y;
  a.M = M; a.N = N; a.K = K; a.G = 1; a.ldx = K; a.ldy = N; a.act = act;
  return a;
}

// token -> image attention + residual + LayerNorm (transformer.py:165-170 / :101-106)
int t2i_block(ullsam_handle* h, int R, int Nt, const float* q_w, const float* q_b, const float* o_w, const float* o_b,
              const float* n_w, const float* n_b, int ldkv, cudaStream_t st) {
  const int M = R * Nt;
  CKS(launch_tok_linear(lin(h->queries, h->tokens, q_w, q_b, nullptr, h->q128, M, 128, 256), st));
  mark(h, PC_DEC_TOKENS, st, 1);
  CKS(launch_t2i_attention(h->q128, h->proj, ldkv, R, Nt, h->o128, st));
  mark(h, PC_DEC_ATTN, st, 1);
  CKS(launch_tok_linear(lin(h->o128, nullptr, o_w, o_b, h->queries, h->tmp, M, 256, 128), st));
  mark(h, PC_DEC_TOKENS, st, 1);
  CKS(launch_tok_ln(h->tmp, n_w, n_b, M, 1e-5f, h->queries, st));
  mark(h, PC_DEC_TOKENS, st, 1);
  return ULLSAM_OK;
}

// TwoWayTransformer.forward (transformer.py:62-108).  On entry h->tokens = point_embedding, h->keys(_bf) = image tokens,
// h->pe_bf = image_pe tokens.  On exit h->queries / h->keys(_bf) hold the outputs.
// with_heads: also run the hypernetwork MLPs and the IoU head (h->hyper, h->iou4), mask_decoder.py:136-148.
bool dec_fused_enabled(const ullsam_handle* h, int Nt) {
  static const bool off = getenv("ULLSAM_DEC_FUSED") != nullptr && atoi(getenv("ULLSAM_DEC_FUSED")) == 0;
  return !off && h->cfg.iou_head_depth <= 4 && dec_tokens_fused_supported(Nt, h->cfg.dec_mlp_dim, h->cfg.iou_head_hidden);
}

int heads_run(ullsam_handle* h, int R, int Nt, cudaStream_t st);

int transformer_run(ullsam_handle* h, int R, int Nt, cudaStream_t st, bool with_heads) {
  const auto& c = h->cfg;
  const int sms = h->sms, M = R * Nt, rk = R * 4096;
  CKC(cudaMemcpyAsync(h->queries, h->tokens, static_cast<size_t>(M) * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const bool fused = dec_fused_enabled(h, Nt);
  // (image_pe) x [Wk ; 0 ; Wq]^T for every layer: the "+ key_pe" of transformer.py:167,179 folded into the projections
  for (int i = 0; i <= c.dec_depth; ++i) {
    const bool fin = (i == c.dec_depth);
    GemmParams p = gp(4096, fin ? 256 : 384, 256);
    p.out = h->peproj[i];
    CKS(launch_gemm(EPI_F32, h->pe_bf, fin ? h->fin_catpe_w : h->layers[i].catpe_w, p, sms, st));
    mark(h, PC_DEC_GEMM, st, 1);
  }
  if (fused) {
    // token side of each block as one cluster kernel (decoder_fused.cu); image side as before
    for (int i = 0; i < c.dec_depth; ++i) {
      const DecLayer& L = h->layers[i];
      {
        GemmParams p = gp(rk, 384, 256);
        p.bias = L.cat_b; p.out = h->proj; p.pos = h->peproj[i]; p.pos_rows = 4096;
        CKS(launch_gemm(EPI_BF16, h->keys_bf, L.cat_w, p, sms, st));
        mark(h, PC_DEC_GEMM, st, 1);
      }
      DecTokLayer w{};
      w.sq_w = L.self.q_w; w.sq_b = L.self.q_b; w.sk_w = L.self.k_w; w.sk_b = L.self.k_b;
      w.sv_w = L.self.v_w; w.sv_b = L.self.v_b; w.so_w = L.self.o_w; w.so_b = L.self.o_b;
      w.n1_w = L.n1_w; w.n1_b = L.n1_b; w.n2_w = L.n2_w; w.n2_b = L.n2_b; w.n3_w = L.n3_w; w.n3_b = L.n3_b;
      w.tq_w = L.t2i_q_w; w.tq_b = L.t2i_q_b; w.to_w = L.t2i_o_w; w.to_b = L.t2i_o_b;
      w.m1_w = L.mlp1_w; w.m1_b = L.mlp1_b; w.m2_w = L.mlp2_w; w.m2_b = L.mlp2_b;
      w.ik_w = L.i2t_k_w; w.ik_b = L.i2t_k_b; w.iv_w = L.i2t_v_w; w.iv_b = L.i2t_v_b;
      w.mlp_dim = c.dec_mlp_dim; w.skip_pe = (i == 0);
      CKS(launch_dec_token_layer(w, h->queries, h->tokens, h->proj, 384, R, Nt, h->k128, h->v128, st));
      mark(h, PC_DEC_TOKENS, st, 1);
      CKS(launch_i2t_attention(h->proj, 384, 256, h->k128, h->v128, R, Nt, h->ao128, st));
      mark(h, PC_DEC_ATTN, st, 1);
      {
        GemmParams p = gp(rk, 256, 128);
        p.bias = L.i2t_o_b; p.residual = h->keys; p.out = h->keys; p.out2 = h->keys_bf;
        p.ln_w = L.n4_w; p.ln_b = L.n4_b; p.ln_eps = 1e-5f;
        CKS(launch_gemm(EPI_LN_RES, h->ao128, L.i2t_o_w, p, sms, st));
        mark(h, PC_DEC_GEMM, st, 1);
      }
    }
    {
      GemmParams p = gp(rk, 256, 256);
      p.bias = h->fin_cat_b; p.out = h->proj; p.pos = h->peproj[c.dec_depth]; p.pos_rows = 4096;
      CKS(launch_gemm(EPI_BF16, h->keys_bf, h->fin_cat_w, p, sms, st));
      mark(h, PC_DEC_GEMM, st, 1);
    }
    DecTokFinal f{};
    f.tq_w = h->fin_q_w; f.tq_b = h->fin_q_b; f.to_w = h->fin_o_w; f.to_b = h->fin_o_b; f.nf_w = h->nf_w; f.nf_b = h->nf_b;
    f.hyp_w0 = h->hyp_w0; f.hyp_b0 = h->hyp_b0; f.hyp_w1 = h->hyp_w1; f.hyp_b1 = h->hyp_b1; f.hyp_w2 = h->hyp_w2; f.hyp_b2 = h->hyp_b2;
    for (int i = 0; i < c.iou_head_depth; ++i) { f.iou_w[i] = h->iou_w[i]; f.iou_b[i] = h->iou_b[i]; }
    f.iou_depth = c.iou_head_depth; f.iou_hidden = c.iou_head_hidden;
    CKS(launch_dec_token_final(f, h->queries, h->tokens, h->proj, 256, R, Nt, with_heads ? h->hyper : nullptr,
                               with_heads ? h->iou4 : nullptr, st));
    mark(h, PC_DEC_TOKENS, st, 1);
    return ULLSAM_OK;
  }
  for (int i = 0; i < c.dec_depth; ++i) {
    const DecLayer& L = h->layers[i];
    const float* pe = (i == 0) ? nullptr : h->tokens;  // skip_first_layer_pe (transformer.py:157-158)
    // ---- self attention ----
    CKS(launch_tok_linear(lin(h->queries, pe, L.self.q_w, L.self.q_b, nullptr, h->tq, M, 256, 256), st));
    mark(h, PC_DEC_TOKENS, st, 1);
    CKS(launch_tok_linear(lin(h->queries, pe, L.self.k_w, L.self.k_b, nullptr, h->tk, M, 256, 256), st));
    mark(h, PC_DEC_TOKENS, st, 1);
    CKS(launch_tok_linear(lin(h->queries, nullptr, L.self.v_w, L.self.v_b, nullptr, h->tv, M, 256, 256), st));
    mark(h, PC_DEC_TOKENS, st, 1);
    CKS(launch_tok_self_attn(h->tq, h->tk, h->tv, R, Nt, h->ta, st));
    mark(h, PC_DEC_TOKENS, st, 1);
    CKS(launch_tok_linear(lin(h->ta, nullptr, L.self.o_w, L.self.o_b, (i == 0) ? nullptr : h->queries, h->tmp, M, 256, 256), st));
    mark(h, PC_DEC_TOKENS, st, 1);
    CKS(launch_tok_ln(h->tmp, L.n1_w, L.n1_b, M, 1e-5f, h->queries, st));
    mark(h, PC_DEC_TOKENS, st, 1);
    // ---- image-side projections for this layer: [K_t2i | V_t2i | Q_i2t] in one GEMM ----
    {
      GemmParams p = gp(rk, 384, 256);
      p.bias = L.cat_b; p.out = h->proj; p.pos = h->peproj[i]; p.pos_rows = 4096;
      CKS(launch_gemm(EPI_BF16, h->keys_bf, L.cat_w, p, sms, st));
      mark(h, PC_DEC_GEMM, st, 1);
    }
    // ---- tokens attend to the image ----
    if (int rc = t2i_block(h, R, Nt, L.t2i_q_w, L.t2i_q_b, L.t2i_o_w, L.t2i_o_b, L.n2_w, L.n2_b, 384, st)) return rc;
    // ---- MLP ----
    CKS(launch_tok_linear(lin(h->queries, nullptr, L.mlp1_w, L.mlp1_b, nullptr, h->thid, M, c.dec_mlp_dim, 256, 2), st));
    mark(h, PC_DEC_TOKENS, st, 1);
    CKS(launch_tok_linear(lin(h->thid, nullptr, L.mlp2_w, L.mlp2_b, h->queries, h->tmp, M, 256, c.dec_mlp_dim), st));
    mark(h, PC_DEC_TOKENS, st, 1);
    CKS(launch_tok_ln(h->tmp, L.n3_w, L.n3_b, M, 1e-5f, h->queries, st));
    mark(h, PC_DEC_TOKENS, st, 1);
    // ---- image attends to the tokens; keys = norm4(keys + out_proj(attn)) ----
    CKS(launch_tok_linear(lin(h->queries, h->tokens, L.i2t_k_w, L.i2t_k_b, nullptr, h->k128, M, 128, 256), st));
    mark(h, PC_DEC_TOKENS, st, 1);
    CKS(launch_tok_linear(lin(h->queries, nullptr, L.i2t_v_w, L.i2t_v_b, nullptr, h->v128, M, 128, 256), st));
    mark(h, PC_DEC_TOKENS, st, 1);
    CKS(launch_i2t_attention(h->proj, 384, 256, h->k128, h->v128, R, Nt, h->ao128, st));
    mark(h, PC_DEC_ATTN, st, 1);
    {
      GemmParams p = gp(rk, 256, 128);
      p.bias = L.i2t_o_b; p.residual = h->keys; p.out = h->keys; p.out2 = h->keys_bf;
      p.ln_w = L.n4_w; p.ln_b = L.n4_b; p.ln_eps = 1e-5f;
      CKS(launch_gemm(EPI_LN_RES, h->ao128, L.i2t_o_w, p, sms, st));
      mark(h, PC_DEC_GEMM, st, 1);
    }
  }
  {  // final token -> image attention
    GemmParams p = gp(rk, 256, 256);
    p.bias = h->fin_cat_b; p.out = h->proj; p.pos = h->peproj[c.dec_depth]; p.pos_rows = 4096;
    CKS(launch_gemm(EPI_BF16, h->keys_bf, h->fin_cat_w, p, sms, st));
    mark(h, PC_DEC_GEMM, st, 1);
    if (int rc = t2i_block(h, R, Nt, h->fin_q_w, h->fin_q_b, h->fin_o_w, h->fin_o_b, h->nf_w, h->nf_b, 256, st)) return rc;
  }
  return with_heads ? heads_run(h, R, Nt, st) : ULLSAM_OK;
}

// hypernetwork MLPs on the 4 mask tokens (grouped) and the IoU head on token 0, launch-per-op path (mask_decoder.py:136-148)
int heads_run(ullsam_handle* h, int R, int Nt, cudaStream_t st) {
  const auto& c = h->cfg;
  const int nm = c.num_mask_tokens;
  {
    TokLinear a = lin(h->queries + 256, nullptr, h->hyp_w0, h->hyp_b0, nullptr, h->hyp1, R, 256, 256, 2);
    a.G = nm; a.ldx = Nt * 256; a.x_gs = 256; a.w_gs = 256 * 256; a.b_gs = 256; a.y_gs = static_cast<size_t>(R) * 256;
    CKS(launch_tok_linear(a, st));
    mark(h, PC_DEC_TOKENS, st, 1);
    TokLinear b2 = lin(h->hyp1, nullptr, h->hyp_w1, h->hyp_b1, nullptr, h->hyp2, R, 256, 256, 2);
    b2.G = nm; b2.x_gs = static_cast<size_t>(R) * 256; b2.w_gs = 256 * 256; b2.b_gs = 256; b2.y_gs = static_cast<size_t>(R) * 256;
    CKS(launch_tok_linear(b2, st));
    mark(h, PC_DEC_TOKENS, st, 1);
    TokLinear c3 = lin(h->hyp2, nullptr, h->hyp_w2, h->hyp_b2, nullptr, h->hyper, R, 32, 256, 0);
    c3.G = nm; c3.x_gs = static_cast<size_t>(R) * 256; c3.w_gs = 32 * 256; c3.b_gs = 32; c3.ldy = nm * 32; c3.y_gs = 32;
    CKS(launch_tok_linear(c3, st));
    mark(h, PC_DEC_TOKENS, st, 1);
  }
  {
    const float* x = h->queries;
    int ldx = Nt * 256;
    float* bufs[2] = {h->iou1, h->iou2};
    for (int i = 0; i < c.iou_head_depth; ++i) {
      const bool last = (i == c.iou_head_depth - 1);
      const int K = (i == 0) ? 256 : c.iou_head_hidden, N = last ? nm : c.iou_head_hidden;
      float* y = last ? h->iou4 : bufs[i & 1];
      TokLinear a = lin(x, nullptr, h->iou_w[i], h->iou_b[i], nullptr, y, R, N, K, last ? 0 : 2);
      a.ldx = ldx;
      CKS(launch_tok_linear(a, st));
      mark(h, PC_DEC_TOKENS, st, 1);
      x = y;
      ldx = N;
    }
  }
  return ULLSAM_OK;
}

// MaskDecoder.predict_masks (mask_decoder.py:112-149) writing mask channels [m0, m0+cnt)
int decoder_run(ullsam_handle* h, const float* emb, int n_img, const float* image_pe, const float* sparse, int P, int Ns,
                const float* dense, int dense_mode, int dense_batch, int m0, int cnt, float* masks, float* iou,
                cudaStream_t st) {
  const auto& c = h->cfg;
  if (int rc = check_ready(h, 1)) return rc;
  if (int rc = check_ready(h, 2)) return rc;
  if (n_img <= 0 || P <= 0 || Ns < 0) return fail(h, ULLSAM_ERR_INVALID, "decoder: n_img, P must be positive");
  const int R = n_img * P, nm = c.num_mask_tokens, Nt = 1 + nm + Ns, sms = h->sms;
  if (Nt > 64) return fail(h, ULLSAM_ERR_INVALID, "decoder: more than 64 prompt tokens per prompt is not supported");
  if (dense_batch != 1 && dense_batch != R) return fail(h, ULLSAM_ERR_INVALID, "decoder: dense batch must be 1 or n_img*P");
  if (nm != 4) return fail(h, ULLSAM_ERR_INVALID, "decoder: num_mask_tokens must be 4");
  if (int rc = ensure_decoder_ws(h, R, Nt)) return rc;
  CKS(launch_tok_init(h->iou_token, h->mask_tokens, nm, sparse, Ns, R, h->tokens, st));
  mark(h, PC_DEC_TOKENS, st, 1);
  CKS(launch_keys_init(emb, R, P, dense, dense_mode, dense_batch, h->keys, h->keys_bf, st));
  mark(h, PC_DEC_TOKENS, st, 1);
  CKS(launch_keys_init(image_pe, 1, 1, nullptr, 0, 1, nullptr, h->pe_bf, st));
  mark(h, PC_DEC_TOKENS, st, 1);
  if (int rc = transformer_run(h, R, Nt, st, true)) return rc;
  {
    CKS(launch_slice_copy(h->iou4, R, nm, m0, cnt, 1, iou, st));
    mark(h, PC_DEC_TOKENS, st, 1);
  }
  {  // output_upscaling stage 1: ConvT + LayerNorm2d + GELU
    GemmParams p = gp(R * 4096, 256, 256);
    p.bias = h->up1_b; p.out = h->up_h; p.ln_w = h->upln_w; p.ln_b = h->upln_b; p.ln_eps = 1e-6f;
    CKS(launch_gemm(EPI_UP1, h->keys_bf, h->up1_w, p, sms, st));
    mark(h, PC_DEC_UPSCALE, st, 1);
  }
  {  // stage 2: ConvT + GELU, contracted with hyper_in -> mask logits
    GemmParams p = gp(R * 16384, 128, 64);
    p.bias = h->up2_b; p.out = masks; p.hyper = h->hyper; p.mask_m0 = m0; p.mask_cnt = cnt;
    CKS(launch_gemm(EPI_MASK, h->up_h, h->up2_w, p, sms, st));
    mark(h, PC_DEC_UPSCALE, st, 1);
  }
  return ULLSAM_OK;
}

}  // namespace

extern "C" {

const char* ullsam_last_error(const ullsam_handle* h) { return h ? h->err.c_str() : g_thread_error.c_str(); }

int ullsam_create(const ullsam_config* cfg, int device, ullsam_handle** out) {
  if (!cfg || !out) return fail(nullptr, ULLSAM_ERR_INVALID, "create: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev)
    return fail(nullptr, ULLSAM_ERR_NO_DEVICE, "create: no CUDA device (this path has no CPU fallback)");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return fail(nullptr, ULLSAM_ERR_NO_DEVICE, "create: device is not sm_100 (B200); the kernels are sm_100a-only");
  if (cfg->img_size != 1024 || cfg->patch_size != 16 || cfg->in_chans != 3 || cfg->window_size != 14 || cfg->out_chans != 256)
    return fail(nullptr, ULLSAM_ERR_INVALID, "create: kernels are specialised for img 1024 / patch 16 / window 14 / out 256");
  if (cfg->depth > 0 && (cfg->embed_dim != cfg->num_heads * 64 || (cfg->embed_dim != 768 && cfg->embed_dim != 1024)))
    return fail(nullptr, ULLSAM_ERR_INVALID, "create: embed_dim must be 768 (vit_b) or 1024 (vit_l) with 64-wide heads; vit_h (80-wide heads) is not implemented");
  if (cfg->mlp_dim % 256 != 0 || cfg->depth < 0 || cfg->num_global < 0 || cfg->num_global > 8)
    return fail(nullptr, ULLSAM_ERR_INVALID, "create: bad mlp_dim / global_attn_indexes");
  if (cfg->transformer_dim != 256 || cfg->dec_heads != 8 || cfg->num_mask_tokens != 4 || cfg->dec_mlp_dim % 4 != 0 ||
      cfg->dec_mlp_dim > 2048 || cfg->iou_head_hidden != 256 || cfg->iou_head_depth < 2 || cfg->dec_depth < 1)
    return fail(nullptr, ULLSAM_ERR_INVALID, "create: decoder must be SAM's (dim 256, 8 heads, 4 mask tokens, mlp <= 2048)");
  DeviceGuard guard(device);
  if (guard.status != cudaSuccess) return fail(nullptr, ULLSAM_ERR_CUDA, "create: cudaSetDevice failed");
  ullsam_handle* h = new ullsam_handle();
  h->cfg = *cfg;
  h->device = device;
  h->sms = prop.multiProcessorCount;
  if (getenv("ULLSAM_LN_FOLD") != nullptr) h->ln_fold = atoi(getenv("ULLSAM_LN_FOLD")) & 3;
  if (cfg->depth > 0) register_encoder(h);  // depth == 0: decoder-only handle (MaskDecoder / TwoWayTransformer shims)
  register_decoder(h);
  if (h->alloc_failed) {
    ullsam_destroy(h);
    return fail(nullptr, ULLSAM_ERR_NOMEM, "create: cudaMalloc failed for the weights");
  }
  *out = h;
  return ULLSAM_OK;
}

void ullsam_destroy(ullsam_handle* h) {
  if (!h) return;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  for (void* p : h->allocs) cudaFree(p);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  if (h->weights_ev) cudaEventDestroy(h->weights_ev);
  for (cudaEvent_t e : h->hs_ev) cudaEventDestroy(e);
  if (h->hs_copy) cudaStreamDestroy(h->hs_copy);
  if (h->hs_comp) cudaStreamDestroy(h->hs_comp);
  if (h->staging) cudaFree(h->staging);
  if (h->staging_raw) cudaFree(h->staging_raw);
  delete h;
}

}  // extern "C"

namespace {
// copies + re-packs one state_dict tensor, stream-ordered on `st`, no host synchronisation
int set_weight_impl(ullsam_handle* h, const char* key, const void* data, int dtype, const int64_t* shape, int ndim,
                    cudaStream_t st) {
  if (!h || !key || !data || (!shape && ndim > 0)) return fail(h, ULLSAM_ERR_INVALID, "set_weight: null argument");
  auto it = h->slots.find(key);
  if (it == h->slots.end()) return fail(h, ULLSAM_ERR_INVALID, std::string("set_weight: unknown state_dict key ") + key);
  Slot& s = it->second;
  bool same = static_cast<size_t>(ndim) == s.shape.size();
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    if (same && shape[i] != s.shape[i]) same = false;
    n *= static_cast<size_t>(shape[i]);
  }
  if (!same) return fail(h, ULLSAM_ERR_INVALID, std::string("set_weight: shape mismatch for ") + key);
  DEVICE_GUARD(h);
  if (h->staging_bytes < n * 4) {
    if (h->staging) cudaFree(h->staging);
    h->staging = nullptr;
    h->staging_bytes = 0;
    CKC(cudaMalloc(reinterpret_cast<void**>(&h->staging), n * 4));
    h->staging_bytes = n * 4;
  }
  if (dtype == ULLSAM_F32) {
    CKC(cudaMemcpyAsync(h->staging, data, n * 4, cudaMemcpyDefault, st));
  } else if (dtype == ULLSAM_BF16 || dtype == ULLSAM_F16) {
    if (h->staging_raw_bytes < n * 2) {
      if (h->staging_raw) cudaFree(h->staging_raw);
      h->staging_raw = nullptr;
      h->staging_raw_bytes = 0;
      CKC(cudaMalloc(&h->staging_raw, n * 2));
      h->staging_raw_bytes = n * 2;
    }
    CKC(cudaMemcpyAsync(h->staging_raw, data, n * 2, cudaMemcpyDefault, st));
    CKS(launch_cast_to_f32(h->staging_raw, dtype, n, h->staging, st));
  } else {
    return fail(h, ULLSAM_ERR_INVALID, "set_weight: dtype must be f32, bf16 or f16");
  }
  // Everything above is ordered on `st` (the staging buffers are reused tensor after tensor in stream order; cudaFree
  // of an outgrown buffer synchronises by itself), so there is no host sync per tensor: the next forward call fences
  // its stream against this one once (weights_fence).
  g_pack_stream = st;
  const char* pe = s.pack(h->staging);
  g_pack_stream = 0;
  if (pe) return fail(h, ULLSAM_ERR_CUDA, pe);
  h->weights_dirty = true;
  if (s.which == 0) h->enc_fold_dirty = true;
  s.loaded = true;
  return ULLSAM_OK;
}
}  // namespace

extern "C" {

int ullsam_set_weight(ullsam_handle* h, const char* key, const void* data, int dtype, const int64_t* shape, int ndim) {
  // synchronous flavour: `data` (host or device) has been consumed when the call returns
  if (h && h->weights_dirty && h->weights_stream != 0) {  // keep one upload stream at a time
    DeviceGuard g(h->device);
    cudaStreamSynchronize(h->weights_stream);
  }
  if (h) h->weights_stream = 0;
  if (int rc = set_weight_impl(h, key, data, dtype, shape, ndim, 0)) return rc;
  DeviceGuard g(h->device);
  cudaError_t ce = cudaStreamSynchronize(0);
  if (ce != cudaSuccess) return fail(h, ULLSAM_ERR_CUDA, std::string("set_weight: ") + cudaGetErrorString(ce));
  return ULLSAM_OK;
}

int ullsam_set_weight_async(ullsam_handle* h, const char* key, const void* data, int dtype, const int64_t* shape, int ndim,
                            void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h && h->weights_dirty && h->weights_stream != st) {  // a different upload stream than last time: drain the old one
    DeviceGuard g(h->device);
    cudaStreamSynchronize(h->weights_stream);
  }
  if (h) h->weights_stream = st;
  return set_weight_impl(h, key, data, dtype, shape, ndim, st);
}

int ullsam_missing_weights(const ullsam_handle* h, int which) {
  if (!h) return -1;
  int missing = 0;
  for (auto& kv : h->slots) {
    const int w = kv.second.which == 2 ? 1 : kv.second.which;  // public classes: 0 encoder, 1 decoder
    if (w == which && !kv.second.loaded) ++missing;
  }
  return missing;
}

size_t ullsam_workspace_bytes(const ullsam_handle* h, int B, int R) {
  if (!h) return 0;
  const size_t D = h->cfg.embed_dim, M = h->cfg.mlp_dim;
  const size_t Bc = static_cast<size_t>(B < kEncChunk ? B : kEncChunk);
  size_t enc = Bc * (4096 * D * 4 + 4 * 4900 * D * 2 + 4096 * D * 2 + 4096 * M * 2 + 4096 * 256 * 4 + 4096 * kStatSlots * 8);
  const size_t Nt = 16;
  size_t dec = static_cast<size_t>(R) * (4096 * (256 * 4 + 256 * 2 + 384 * 2 + 128 * 2) + 16384 * 64 * 2 + Nt * 8 * 4096 * 4);
  return enc + dec;
}

int ullsam_encoder_forward(ullsam_handle* h, const void* images, int dtype, int B, void* out_emb, int out_dtype, void* stream) {
  if (!h || !images || !out_emb) return fail(h, ULLSAM_ERR_INVALID, "encoder_forward: null argument");
  if (dtype != ULLSAM_F32 && dtype != ULLSAM_BF16) return fail(h, ULLSAM_ERR_INVALID, "encoder_forward: images must be f32 or bf16");
  if (out_dtype != ULLSAM_F32 && out_dtype != ULLSAM_BF16) return fail(h, ULLSAM_ERR_INVALID, "encoder_forward: output must be f32 or bf16");
  if (B < 0) return fail(h, ULLSAM_ERR_INVALID, "encoder_forward: negative batch");
  if (h->cfg.depth <= 0) return fail(h, ULLSAM_ERR_STATE, "encoder_forward: this handle was created without an encoder (depth 0)");
  if (int rc = check_ready(h, 0)) return rc;
  DEVICE_GUARD(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = weights_fence(h, st)) return rc;
  if (int rc = ensure_encoder_ws(h, B < kEncChunk ? B : kEncChunk)) return rc;
  const size_t img_elems = static_cast<size_t>(3) * 1024 * 1024, emb_elems = static_cast<size_t>(256) * 4096;
  for (int b0 = 0; b0 < B; b0 += kEncChunk) {
    const int Bc = (B - b0 < kEncChunk) ? (B - b0) : kEncChunk;
    const void* img = static_cast<const char*>(images) + static_cast<size_t>(b0) * img_elems * (dtype == ULLSAM_F32 ? 4 : 2);
    if (out_dtype == ULLSAM_F32) {
      if (int rc = encoder_chunk(h, img, dtype, Bc, static_cast<float*>(out_emb) + static_cast<size_t>(b0) * emb_elems, st)) return rc;
    } else {
      if (int rc = encoder_chunk(h, img, dtype, Bc, h->emb_f32, st)) return rc;
      CKS(launch_cast_f32_bf16(h->emb_f32, static_cast<size_t>(Bc) * emb_elems,
                               static_cast<bf16*>(out_emb) + static_cast<size_t>(b0) * emb_elems, st));
    }
  }
  return ULLSAM_OK;
}

int ullsam_decoder_predict_masks(ullsam_handle* h, const void* emb, int n_img, const void* image_pe, const void* sparse, int P,
                                 int Ns, const void* dense, int dense_mode, int dense_batch, void* masks4, void* iou4,
                                 void* stream) {
  if (!h || !emb || !image_pe || !masks4 || !iou4 || (Ns > 0 && !sparse))
    return fail(h, ULLSAM_ERR_INVALID, "decoder: null argument");
  DEVICE_GUARD(h);
  if (int rc = weights_fence(h, static_cast<cudaStream_t>(stream))) return rc;
  return decoder_run(h, static_cast<const float*>(emb), n_img, static_cast<const float*>(image_pe),
                     static_cast<const float*>(sparse), P, Ns, static_cast<const float*>(dense), dense_mode, dense_batch, 0,
                     h->cfg.num_mask_tokens, static_cast<float*>(masks4), static_cast<float*>(iou4),
                     static_cast<cudaStream_t>(stream));
}

int ullsam_decoder_forward(ullsam_handle* h, const void* emb, int n_img, const void* image_pe, const void* sparse, int P, int Ns,
                           const void* dense, int dense_mode, int dense_batch, int multimask, void* masks, void* iou,
                           void* stream) {
  if (!h || !emb || !image_pe || !masks || !iou || (Ns > 0 && !sparse))
    return fail(h, ULLSAM_ERR_INVALID, "decoder: null argument");
  DEVICE_GUARD(h);
  if (int rc = weights_fence(h, static_cast<cudaStream_t>(stream))) return rc;
  const int m0 = multimask ? 1 : 0, cnt = multimask ? h->cfg.num_mask_tokens - 1 : 1;  // mask_decoder.py:102-107
  return decoder_run(h, static_cast<const float*>(emb), n_img, static_cast<const float*>(image_pe),
                     static_cast<const float*>(sparse), P, Ns, static_cast<const float*>(dense), dense_mode, dense_batch, m0,
                     cnt, static_cast<float*>(masks), static_cast<float*>(iou), static_cast<cudaStream_t>(stream));
}

int ullsam_transformer_forward(ullsam_handle* h, const void* image_embedding, int R, const void* image_pe,
                               const void* point_embedding, int Nt, void* queries, void* keys, void* stream) {
  if (!h || !image_embedding || !image_pe || !point_embedding || !queries || !keys)
    return fail(h, ULLSAM_ERR_INVALID, "transformer: null argument");
  if (R <= 0 || Nt <= 0 || Nt > 64) return fail(h, ULLSAM_ERR_INVALID, "transformer: need 1 <= Nt <= 64 and R >= 1");
  if (int rc = check_ready(h, 1)) return rc;
  DEVICE_GUARD(h);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = weights_fence(h, st)) return rc;
  if (int rc = ensure_decoder_ws(h, R, Nt)) return rc;
  CKC(cudaMemcpyAsync(h->tokens, point_embedding, static_cast<size_t>(R) * Nt * 256 * 4, cudaMemcpyDeviceToDevice, st));
  CKS(launch_keys_init(static_cast<const float*>(image_embedding), R, 1, nullptr, 0, 1, h->keys, h->keys_bf, st));
  mark(h, PC_DEC_TOKENS, st, 1);
  CKS(launch_keys_init(static_cast<const float*>(image_pe), 1, 1, nullptr, 0, 1, nullptr, h->pe_bf, st));
  mark(h, PC_DEC_TOKENS, st, 1);
  if (int rc = transformer_run(h, R, Nt, st, false)) return rc;
  CKC(cudaMemcpyAsync(queries, h->queries, static_cast<size_t>(R) * Nt * 256 * 4, cudaMemcpyDeviceToDevice, st));
  CKC(cudaMemcpyAsync(keys, h->keys, static_cast<size_t>(R) * 4096 * 256 * 4, cudaMemcpyDeviceToDevice, st));
  return ULLSAM_OK;
}

unsigned long long ullsam_launch_count(const ullsam_handle* h) { return h ? h->launches : 0; }

int ullsam_profile_begin(ullsam_handle* h) {
  if (!h) return ULLSAM_ERR_INVALID;
  h->prof_on = true;
  h->prof_n = 0;
  h->prof_cat.clear();
  return ULLSAM_OK;
}

int ullsam_profile_categories(void) { return PC_COUNT; }
const char* ullsam_profile_category_name(int i) { return (i >= 0 && i < PC_COUNT) ? kProfNames[i] : ""; }

int ullsam_profile_end(ullsam_handle* h, float* ms_per_category, int* launches_per_category, int n) {
  if (!h || !ms_per_category || !launches_per_category || n < PC_COUNT) return fail(h, ULLSAM_ERR_INVALID, "profile_end: bad argument");
  h->prof_on = false;
  for (int i = 0; i < n; ++i) { ms_per_category[i] = 0.f; launches_per_category[i] = 0; }
  if (h->prof_n == 0) return ULLSAM_OK;
  CKC(cudaEventSynchronize(h->prof_ev[h->prof_n - 1]));
  for (size_t i = 1; i < h->prof_n; ++i) {
    float ms = 0.f;
    CKC(cudaEventElapsedTime(&ms, h->prof_ev[i - 1], h->prof_ev[i]));
    ms_per_category[h->prof_cat[i]] += ms;
    launches_per_category[h->prof_cat[i]] += 1;
  }
  return ULLSAM_OK;
}

int ullsam_segment_host(ullsam_handle* h, const float* images_host, int B, const float* image_pe_host, const float* sparse_host,
                        int Ns, const float* dense_vec_host, int multimask, float* masks_host, float* iou_host, float* emb_host) {
  if (!h || !images_host || !image_pe_host || !sparse_host || !dense_vec_host || !masks_host || !iou_host || B <= 0 || Ns <= 0)
    return fail(h, ULLSAM_ERR_INVALID, "segment_host: null argument");
  if (h->cfg.depth <= 0) return fail(h, ULLSAM_ERR_STATE, "segment_host: this handle was created without an encoder (depth 0)");
  if (int rc = check_ready(h, 0)) return rc;
  if (int rc = check_ready(h, 1)) return rc;
  if (int rc = check_ready(h, 2)) return rc;
  DEVICE_GUARD(h);
  const int cnt = multimask ? 3 : 1;
  const size_t img = static_cast<size_t>(3) * 1024 * 1024, embn = static_cast<size_t>(256) * 4096;
  if (B > h->host_cap_b || Ns > h->host_cap_ns) {
    for (void* p : {(void*)h->d_images, (void*)h->d_pe, (void*)h->d_sparse, (void*)h->d_dense, (void*)h->d_masks, (void*)h->d_iou})
      dfree(h, p);
    const int cb = B > h->host_cap_b ? B : h->host_cap_b, cn = Ns > h->host_cap_ns ? Ns : h->host_cap_ns;
    h->d_images = dalloc<float>(h, cb * img);
    h->d_pe = dalloc<float>(h, 256 * 4096);
    h->d_sparse = dalloc<float>(h, static_cast<size_t>(cb) * cn * 256);
    h->d_dense = dalloc<float>(h, 256);
    h->d_masks = dalloc<float>(h, static_cast<size_t>(cb) * 3 * 65536);
    h->d_iou = dalloc<float>(h, static_cast<size_t>(cb) * 4);
    if (!h->d_images || !h->d_pe || !h->d_sparse || !h->d_dense || !h->d_masks || !h->d_iou) {
      h->host_cap_b = h->host_cap_ns = 0;
      return fail(h, ULLSAM_ERR_NOMEM, "segment_host: cudaMalloc failed");
    }
    h->host_cap_b = cb;
    h->host_cap_ns = cn;
  }
  if (int rc = ensure_encoder_ws(h, B < kEncChunk ? B : kEncChunk)) return rc;
  // Two internal streams: the copy stream moves sub-batch i+1 of the images to the device (and results back) while the
  // compute stream runs encoder + decoder on sub-batch i, so only the first sub-batch's H2D and the last one's D2H are
  // exposed (pinned host buffers; pageable memory degrades to staged, serial copies).  The sub-batches grow — 4, 12, 16
  // images of a 32-image chunk: a small first one so that compute starts after 50 MB of copies instead of 100, large later
  // ones because the GEMMs run below their best on 8-image batches.
  constexpr int kMaxSub = 4;
  static const int kSubSizes[kMaxSub] = {4, 12, 16, kEncChunk};
  if (!h->hs_copy) {
    CKC(cudaStreamCreateWithFlags(&h->hs_copy, cudaStreamNonBlocking));
    CKC(cudaStreamCreateWithFlags(&h->hs_comp, cudaStreamNonBlocking));
    for (int i = 0; i < 2 * kMaxSub; ++i) {
      cudaEvent_t e;
      CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->hs_ev.push_back(e);
    }
  }
  cudaStream_t cs = h->hs_copy, ks = h->hs_comp;
  if (int rc = weights_fence(h, ks)) return rc;
  CKC(cudaMemcpyAsync(h->d_pe, image_pe_host, 256 * 4096 * 4, cudaMemcpyHostToDevice, cs));
  CKC(cudaMemcpyAsync(h->d_sparse, sparse_host, static_cast<size_t>(B) * Ns * 256 * 4, cudaMemcpyHostToDevice, cs));
  CKC(cudaMemcpyAsync(h->d_dense, dense_vec_host, 256 * 4, cudaMemcpyHostToDevice, cs));
  for (int c0 = 0; c0 < B; c0 += kEncChunk) {          // outer chunk: what the embedding staging buffer holds
    const int Bo = (B - c0 < kEncChunk) ? (B - c0) : kEncChunk;
    int nsub = 0, sub_off[kMaxSub + 1] = {0};
    while (sub_off[nsub] < Bo && nsub < kMaxSub) {
      const int left = Bo - sub_off[nsub];
      sub_off[nsub + 1] = sub_off[nsub] + (kSubSizes[nsub] < left ? kSubSizes[nsub] : left);
      ++nsub;
    }
    for (int i = 0; i < nsub; ++i) {
      const int b0 = c0 + sub_off[i], Bc = sub_off[i + 1] - sub_off[i];
      CKC(cudaMemcpyAsync(h->d_images + b0 * img, images_host + b0 * img, Bc * img * 4, cudaMemcpyHostToDevice, cs));
      CKC(cudaEventRecord(h->hs_ev[i], cs));
    }
    for (int i = 0; i < nsub; ++i) {
      const int b0 = c0 + sub_off[i], Bc = sub_off[i + 1] - sub_off[i];
      float* emb = h->emb_f32 + static_cast<size_t>(sub_off[i]) * embn;
      CKC(cudaStreamWaitEvent(ks, h->hs_ev[i], 0));
      if (int rc = encoder_chunk(h, h->d_images + b0 * img, ULLSAM_F32, Bc, emb, ks)) return rc;
      if (int rc = decoder_run(h, emb, Bc, h->d_pe, h->d_sparse + static_cast<size_t>(b0) * Ns * 256, 1, Ns, h->d_dense, 1, 1,
                               multimask ? 1 : 0, cnt, h->d_masks + static_cast<size_t>(b0) * cnt * 65536,
                               h->d_iou + static_cast<size_t>(b0) * cnt, ks))
        return rc;
      CKC(cudaEventRecord(h->hs_ev[kMaxSub + i], ks));
    }
    for (int i = 0; i < nsub; ++i) {
      const int b0 = c0 + sub_off[i], Bc = sub_off[i + 1] - sub_off[i];
      CKC(cudaStreamWaitEvent(cs, h->hs_ev[kMaxSub + i], 0));
      CKC(cudaMemcpyAsync(masks_host + static_cast<size_t>(b0) * cnt * 65536, h->d_masks + static_cast<size_t>(b0) * cnt * 65536,
                          static_cast<size_t>(Bc) * cnt * 65536 * 4, cudaMemcpyDeviceToHost, cs));
      CKC(cudaMemcpyAsync(iou_host + static_cast<size_t>(b0) * cnt, h->d_iou + static_cast<size_t>(b0) * cnt,
                          static_cast<size_t>(Bc) * cnt * 4, cudaMemcpyDeviceToHost, cs));
      if (emb_host)
        CKC(cudaMemcpyAsync(emb_host + static_cast<size_t>(b0) * embn, h->emb_f32 + static_cast<size_t>(sub_off[i]) * embn,
                            static_cast<size_t>(Bc) * embn * 4, cudaMemcpyDeviceToHost, cs));
    }
    // (the next outer chunk's H2D copies queue behind these D2H copies on the copy stream, and its compute waits for
    //  them through hs_ev[i], so the staging buffers are never overwritten while they are still being read)
  }
  CKC(cudaStreamSynchronize(cs));
  CKC(cudaStreamSynchronize(ks));
  return ULLSAM_OK;
}

}  // extern "C"
