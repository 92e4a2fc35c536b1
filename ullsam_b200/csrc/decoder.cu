// ullsam_b200 — token-side kernels of the mask decoder (TwoWayTransformer / MaskDecoder).
//
// The image-token side of the decoder (K/V/Q projections over 4096 tokens per prompt, norm4, the two
// ConvTranspose stages and the mask contraction) runs on the tcgen05 GEMM with fused epilogues
// (gemm_sm100.cuh: EPI_BF16+pos, EPI_LN_RES, EPI_UP1, EPI_MASK).  What is left is the prompt-token side:
// a handful of <= 64 tokens per prompt.  These are latency-bound, fp32, SIMT kernels.
//
// Reference: modeling/transformer.py:62-242, modeling/mask_decoder.py:112-176.
#include "decoder.h"
#include "host_util.h"

#include "common.cuh"

namespace ullsam {

namespace {

// ---------------------------------------------------------------------------------------------
// Y[g][m][n] = act( (X[g][m][:] (+ X2[g][m][:])) . W[g][n][:] + b[g][n] ) (+ R[g][m][n])
// One warp per output column, 4 rows per block; fp32 throughout.
// ---------------------------------------------------------------------------------------------
constexpr int LIN_RB = 4;
constexpr int LIN_CPW = 4;  // columns per warp

__global__ void __launch_bounds__(256) tok_linear_kernel(TokLinear a) {
  extern __shared__ float xs[];  // [LIN_RB][K]
  const int g = blockIdx.z;
  const int m0 = blockIdx.y * LIN_RB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* X = a.x + static_cast<size_t>(g) * a.x_gs;
  const float* X2 = a.x2 ? a.x2 + static_cast<size_t>(g) * a.x_gs : nullptr;
  for (int idx = threadIdx.x; idx < LIN_RB * a.K; idx += blockDim.x) {
    const int r = idx / a.K, k = idx - r * a.K;
    float v = 0.f;
    if (m0 + r < a.M) {
      v = X[static_cast<size_t>(m0 + r) * a.ldx + k];
      if (X2) v += X2[static_cast<size_t>(m0 + r) * a.ldx + k];
    }
    xs[idx] = v;
  }
  __syncthreads();
  const float* W = a.w + static_cast<size_t>(g) * a.w_gs;
  const float* B = a.b ? a.b + static_cast<size_t>(g) * a.b_gs : nullptr;
#pragma unroll 1
  for (int cc = 0; cc < LIN_CPW; ++cc) {
    const int n = (blockIdx.x * 8 + warp) * LIN_CPW + cc;
    if (n >= a.N) break;
    const float4* wr = reinterpret_cast<const float4*>(W + static_cast<size_t>(n) * a.K);
    float acc[LIN_RB] = {0.f, 0.f, 0.f, 0.f};
    for (int k4 = lane; k4 < a.K / 4; k4 += 32) {
      const float4 w4 = __ldg(wr + k4);
#pragma unroll
      for (int r = 0; r < LIN_RB; ++r) {
        const float4 x4 = reinterpret_cast<const float4*>(xs + r * a.K)[k4];
        acc[r] += w4.x * x4.x + w4.y * x4.y + w4.z * x4.z + w4.w * x4.w;
      }
    }
#pragma unroll
    for (int r = 0; r < LIN_RB; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
      const float bias = B ? B[n] : 0.f;
#pragma unroll
      for (int r = 0; r < LIN_RB; ++r) {
        if (m0 + r >= a.M) break;
        float v = acc[r] + bias;
        if (a.act == 2) v = fmaxf(v, 0.f);
        const size_t o = static_cast<size_t>(g) * a.y_gs + static_cast<size_t>(m0 + r) * a.ldy + n;
        if (a.res) v += a.res[o];
        a.y[o] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Token self-attention core (transformer.py:220-242 after the projections): q,k,v [R][Nt][256],
// 8 heads x 32.  One CTA per prompt, one warp per head.  Nt <= 64.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tok_self_attn_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                            const float* __restrict__ v, int Nt, int heads, int dim,
                                                            float* __restrict__ out) {
  const int r = blockIdx.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (h >= heads) return;
  const int hd = dim / heads;  // 32
  const float scale = rsqrtf(static_cast<float>(hd));
  const float* qb = q + static_cast<size_t>(r) * Nt * dim + h * hd;
  const float* kb = k + static_cast<size_t>(r) * Nt * dim + h * hd;
  const float* vb = v + static_cast<size_t>(r) * Nt * dim + h * hd;
  for (int i = 0; i < Nt; ++i) {
    float s0 = -INFINITY, s1 = -INFINITY;
    if (lane < Nt) {
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a += qb[i * dim + d] * kb[lane * dim + d];
      s0 = a * scale;
    }
    if (lane + 32 < Nt) {
      float a = 0.f;
      for (int d = 0; d < hd; ++d) a += qb[i * dim + d] * kb[(lane + 32) * dim + d];
      s1 = a * scale;
    }
    const float mx = warp_max(fmaxf(s0, s1));
    const float p0 = (lane < Nt) ? expf(s0 - mx) : 0.f;
    const float p1 = (lane + 32 < Nt) ? expf(s1 - mx) : 0.f;
    const float inv = 1.0f / warp_sum(p0 + p1);
    float acc = 0.f;  // lane == output dim (hd == 32)
    for (int j = 0; j < Nt; ++j) {
      const float pj = __shfl_sync(0xffffffffu, (j < 32) ? p0 : p1, j & 31);
      if (lane < hd) acc += pj * vb[j * dim + lane];
    }
    if (lane < hd) out[(static_cast<size_t>(r) * Nt + i) * dim + h * hd + lane] = acc * inv;
  }
}

// ---------------------------------------------------------------------------------------------
// y = LayerNorm(x) over 256 channels, eps 1e-5 (nn.LayerNorm default); one warp per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tok_ln_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, int rows, float eps,
                                                     float* __restrict__ y) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * 256);
  const float4 a = xr[lane], b = xr[32 + lane];
  const float mean = warp_sum((a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w)) * (1.0f / 256);
  float sq = 0.f;
  sq += (a.x - mean) * (a.x - mean) + (a.y - mean) * (a.y - mean) + (a.z - mean) * (a.z - mean) + (a.w - mean) * (a.w - mean);
  sq += (b.x - mean) * (b.x - mean) + (b.y - mean) * (b.y - mean) + (b.z - mean) * (b.z - mean) + (b.w - mean) * (b.w - mean);
  const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / 256) + eps);
  const float4 g0 = reinterpret_cast<const float4*>(gamma)[lane], g1 = reinterpret_cast<const float4*>(gamma)[32 + lane];
  const float4 b0 = reinterpret_cast<const float4*>(beta)[lane], b1 = reinterpret_cast<const float4*>(beta)[32 + lane];
  float4* yr = reinterpret_cast<float4*>(y + static_cast<size_t>(row) * 256);
  yr[lane] = make_float4((a.x - mean) * rstd * g0.x + b0.x, (a.y - mean) * rstd * g0.y + b0.y,
                         (a.z - mean) * rstd * g0.z + b0.z, (a.w - mean) * rstd * g0.w + b0.w);
  yr[32 + lane] = make_float4((b.x - mean) * rstd * g1.x + b1.x, (b.y - mean) * rstd * g1.y + b1.y,
                              (b.z - mean) * rstd * g1.z + b1.z, (b.w - mean) * rstd * g1.w + b1.w);
}

// ---------------------------------------------------------------------------------------------
// token -> image cross attention (transformer.py:166-170, 102-105), internal dim 128 = 8 heads x 16:
//   o[r][i][h,:] = softmax_key( q[r][i][h,:] . K[r][key][h,:] / 4 ) . V[r][key][h,:]
// K and V are columns [0,128) and [128,256) of the bf16 projection buffer kv[R*4096][ldkv].
// One CTA per (head, prompt): the 4096 keys are spread over the 256 threads, every thread keeps an online softmax
// (max, sum, 16-wide accumulator) for up to TG prompt tokens at once, so K and V are read ONCE per group of TG tokens
// and the [R][8][Nt][4096] score tensor never exists; the partial results are merged through shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int T2I_TG = 4;

__global__ void __launch_bounds__(256) t2i_attention_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ kv,
                                                            int ldkv, int Nt, float* __restrict__ o) {
  constexpr int TG = T2I_TG;
  __shared__ float qs[TG][16];
  __shared__ float red[8][TG][18];  // per warp and token: max, sum, acc[16]
  const int h = blockIdx.x, r = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i0 = 0; i0 < Nt; i0 += TG) {
    const int ng = (Nt - i0 < TG) ? (Nt - i0) : TG;
    __syncthreads();
    if (threadIdx.x < TG * 16) {
      const int i = threadIdx.x >> 4, d = threadIdx.x & 15;
      qs[i][d] = (i < ng) ? q[(static_cast<size_t>(r) * Nt + i0 + i) * 128 + h * 16 + d] * 0.25f : 0.f;  // 1/sqrt(16)
    }
    __syncthreads();
    float m[TG], l[TG], acc[TG][16];
#pragma unroll
    for (int i = 0; i < TG; ++i) {
      m[i] = -INFINITY;
      l[i] = 0.f;
#pragma unroll
      for (int d = 0; d < 16; ++d) acc[i][d] = 0.f;
    }
#pragma unroll 1
    for (int jj = 0; jj < 16; ++jj) {
      const int key = jj * 256 + threadIdx.x;
      const uint4* kp = reinterpret_cast<const uint4*>(kv + (static_cast<size_t>(r) * 4096 + key) * ldkv + h * 16);
      const uint4* vp = kp + 16;  // + 128 bf16
      const uint4 k0 = kp[0], k1 = kp[1], v0 = vp[0], v1 = vp[1];
      float kf[16], vf[16];
      const __nv_bfloat162* a0 = reinterpret_cast<const __nv_bfloat162*>(&k0);
      const __nv_bfloat162* a1 = reinterpret_cast<const __nv_bfloat162*>(&k1);
      const __nv_bfloat162* c0 = reinterpret_cast<const __nv_bfloat162*>(&v0);
      const __nv_bfloat162* c1 = reinterpret_cast<const __nv_bfloat162*>(&v1);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 t0 = __bfloat1622float2(a0[u]), t1 = __bfloat1622float2(a1[u]);
        kf[2 * u] = t0.x; kf[2 * u + 1] = t0.y; kf[8 + 2 * u] = t1.x; kf[8 + 2 * u + 1] = t1.y;
        const float2 w0 = __bfloat1622float2(c0[u]), w1 = __bfloat1622float2(c1[u]);
        vf[2 * u] = w0.x; vf[2 * u + 1] = w0.y; vf[8 + 2 * u] = w1.x; vf[8 + 2 * u + 1] = w1.y;
      }
#pragma unroll
      for (int i = 0; i < TG; ++i) {
        float sc = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) sc = fmaf(qs[i][d], kf[d], sc);
        const float mn = fmaxf(m[i], sc);
        const float corr = __expf(m[i] - mn), p = __expf(sc - mn);
        l[i] = l[i] * corr + p;
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[i][d] = fmaf(acc[i][d], corr, p * vf[d]);
        m[i] = mn;
      }
    }
    // merge: warp first (common max, rescale, sum), then the 8 warps through shared memory
#pragma unroll
    for (int i = 0; i < TG; ++i) {
      const float wm = warp_max(m[i]);
      const float sc = __expf(m[i] - wm);
      const float ls = warp_sum(l[i] * sc);
      if (lane == 0) { red[warp][i][0] = wm; red[warp][i][1] = ls; }
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float a = warp_sum(acc[i][d] * sc);
        if (lane == 0) red[warp][i][2 + d] = a;
      }
    }
    __syncthreads();
    if (threadIdx.x < TG * 16) {
      const int i = threadIdx.x >> 4, d = threadIdx.x & 15;
      if (i < ng) {
        float gm = red[0][i][0];
#pragma unroll
        for (int w = 1; w < 8; ++w) gm = fmaxf(gm, red[w][i][0]);
        float L = 0.f, A = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float sc = __expf(red[w][i][0] - gm);
          L = fmaf(red[w][i][1], sc, L);
          A = fmaf(red[w][i][2 + d], sc, A);
        }
        o[(static_cast<size_t>(r) * Nt + i0 + i) * 128 + h * 16 + d] = A / L;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// image -> token cross attention core (transformer.py:178-181): every image token attends to the Nt
// prompt tokens.  q = columns [q_col, q_col+128) of the bf16 projection buffer; k, v fp32 [R][Nt][128].
// One thread per (image token, head); softmax over <= 64 tokens kept online.  out: bf16 [R*4096][128].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) i2t_attention_kernel(const __nv_bfloat16* __restrict__ proj, int ldp, int q_col,
                                                            const float* __restrict__ k, const float* __restrict__ v,
                                                            int Nt, __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  // k, v in shared memory as [Nt][8 heads][20]: the 16 floats of a head padded to 20, so that the eight heads of a
  // token (the 8 threads of one 128-bit load phase) hit eight different 16-byte bank groups
  extern __shared__ __align__(16) float kvs[];
  constexpr int HS = 20, TS = 8 * HS;
  const int r = blockIdx.y;
  float* ks = kvs;
  float* vs = kvs + Nt * TS;
  for (int idx = threadIdx.x; idx < Nt * 128; idx += 256) {
    const int j = idx >> 7, c = idx & 127;
    const int o = j * TS + (c >> 4) * HS + (c & 15);
    ks[o] = k[static_cast<size_t>(r) * Nt * 128 + idx] * 0.25f;   // 1/sqrt(16) folded into k
    vs[o] = v[static_cast<size_t>(r) * Nt * 128 + idx];
  }
  __syncthreads();
  const int tok = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int h = threadIdx.x & 7;
  const size_t row = static_cast<size_t>(r) * 4096 + tok;
  const uint4* qp = reinterpret_cast<const uint4*>(proj + row * ldp + q_col + h * 16);
  const uint4 q0 = qp[0], q1 = qp[1];
  float qf[16];
  const __nv_bfloat162* b0 = reinterpret_cast<const __nv_bfloat162*>(&q0);
  const __nv_bfloat162* b1 = reinterpret_cast<const __nv_bfloat162*>(&q1);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t0 = __bfloat1622float2(b0[j]), t1 = __bfloat1622float2(b1[j]);
    qf[2 * j] = t0.x; qf[2 * j + 1] = t0.y;
    qf[8 + 2 * j] = t1.x; qf[8 + 2 * j + 1] = t1.y;
  }
  float m = -INFINITY, l = 0.f, acc[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) acc[d] = 0.f;
  for (int j = 0; j < Nt; ++j) {
    const float4* kj = reinterpret_cast<const float4*>(ks + j * TS + h * HS);
    float s = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
      const float4 kk = kj[d4];
      s += qf[4 * d4] * kk.x + qf[4 * d4 + 1] * kk.y + qf[4 * d4 + 2] * kk.z + qf[4 * d4 + 3] * kk.w;
    }
    const float mn = fmaxf(m, s);
    const float corr = expf(m - mn), p = expf(s - mn);
    l = l * corr + p;
    const float4* vj = reinterpret_cast<const float4*>(vs + j * TS + h * HS);
#pragma unroll
    for (int d4 = 0; d4 < 4; ++d4) {
      const float4 vv = vj[d4];
      acc[4 * d4] = acc[4 * d4] * corr + p * vv.x;
      acc[4 * d4 + 1] = acc[4 * d4 + 1] * corr + p * vv.y;
      acc[4 * d4 + 2] = acc[4 * d4 + 2] * corr + p * vv.z;
      acc[4 * d4 + 3] = acc[4 * d4 + 3] * corr + p * vv.w;
    }
    m = mn;
  }
  const float inv = 1.0f / l;
  uint4* dst = reinterpret_cast<uint4*>(out + row * 128 + h * 16);
  dst[0] = make_uint4(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv),
                      pack_bf16(acc[4] * inv, acc[5] * inv), pack_bf16(acc[6] * inv, acc[7] * inv));
  dst[1] = make_uint4(pack_bf16(acc[8] * inv, acc[9] * inv), pack_bf16(acc[10] * inv, acc[11] * inv),
                      pack_bf16(acc[12] * inv, acc[13] * inv), pack_bf16(acc[14] * inv, acc[15] * inv));
}

// ---------------------------------------------------------------------------------------------
// keys[r][pix][c] = emb[img(r)][c][pix] + dense   (mask_decoder.py:126-127 + transformer.py:82-84)
// 32x32 shared-memory transpose; writes fp32 and/or bf16.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) keys_init_kernel(const float* __restrict__ emb, int P, const float* __restrict__ dense,
                                                        int dense_mode, int dense_batch, float* __restrict__ keys_f32,
                                                        __nv_bfloat16* __restrict__ keys_bf16) {
  pdl_wait();
  pdl_trigger();
  __shared__ float tile[32][33];
  const int r = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = emb + static_cast<size_t>(r / P) * 256 * 4096;
  const int rd = (dense_batch == 1) ? 0 : r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i;
    float v = src[static_cast<size_t>(c) * 4096 + p0 + tx];
    if (dense != nullptr) {
      if (dense_mode == 0) v += dense[(static_cast<size_t>(rd) * 256 + c) * 4096 + p0 + tx];
      else v += dense[static_cast<size_t>(rd) * 256 + c];
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pix = p0 + ty + 8 * i;
    const float v = tile[tx][ty + 8 * i];
    const size_t o = (static_cast<size_t>(r) * 4096 + pix) * 256 + c0 + tx;
    if (keys_f32) keys_f32[o] = v;
    if (keys_bf16) keys_bf16[o] = __float2bfloat16(v);
  }
}

// tokens[r] = cat(iou_token, mask_tokens, sparse[r])  (mask_decoder.py:121-123)
__global__ void __launch_bounds__(256) tok_init_kernel(const float* __restrict__ iou_token, const float* __restrict__ mask_tokens,
                                                       int nm, const float* __restrict__ sparse, int Ns,
                                                       float* __restrict__ tokens) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x, Nt = 1 + nm + Ns;
  for (int idx = threadIdx.x; idx < Nt * 256; idx += 256) {
    const int t = idx >> 8, c = idx & 255;
    float v;
    if (t == 0) v = iou_token[c];
    else if (t <= nm) v = mask_tokens[(t - 1) * 256 + c];
    else v = sparse[(static_cast<size_t>(r) * Ns + (t - 1 - nm)) * 256 + c];
    tokens[(static_cast<size_t>(r) * Nt) * 256 + idx] = v;
  }
}

// out[r][j][...] = in[r][m0 + j][...]: channel slice of [R][4][inner] tensors (mask_decoder.py:102-107)
__global__ void __launch_bounds__(256) slice_copy_kernel(const float* __restrict__ in, int C, int m0, int cnt, size_t inner,
                                                         float* __restrict__ out, size_t total) {
  pdl_wait();
  pdl_trigger();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t in_i = i % inner;
    const size_t rc = i / inner;
    const int j = static_cast<int>(rc % cnt);
    const size_t r = rc / cnt;
    out[i] = in[(r * C + m0 + j) * inner + in_i];
  }
}

const char* last_err() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

const char* launch_tok_linear(const TokLinear& a, cudaStream_t stream) {
  if (a.M <= 0) return nullptr;
  if (a.K % 4 != 0 || a.K > 2048) return "tok_linear: K must be a multiple of 4 and <= 2048";
  dim3 grid((a.N + 8 * LIN_CPW - 1) / (8 * LIN_CPW), (a.M + LIN_RB - 1) / LIN_RB, a.G > 0 ? a.G : 1);
  tok_linear_kernel<<<grid, 256, LIN_RB * a.K * sizeof(float), stream>>>(a);
  return last_err();
}

const char* launch_tok_self_attn(const float* q, const float* k, const float* v, int R, int Nt, float* out,
                                 cudaStream_t stream) {
  if (Nt > 64) return "decoder: more than 64 prompt tokens is not supported";
  tok_self_attn_kernel<<<R, 256, 0, stream>>>(q, k, v, Nt, 8, 256, out);
  return last_err();
}

const char* launch_tok_ln(const float* x, const float* gamma, const float* beta, int rows, float eps, float* y,
                          cudaStream_t stream) {
  tok_ln_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(x, gamma, beta, rows, eps, y);
  return last_err();
}

const char* launch_t2i_attention(const float* q, const __nv_bfloat16* kv, int ldkv, int R, int Nt, float* o,
                                 cudaStream_t stream) {
  t2i_attention_kernel<<<dim3(8, R), 256, 0, stream>>>(q, kv, ldkv, Nt, o);
  return last_err();
}

const char* launch_i2t_attention(const __nv_bfloat16* proj, int ldp, int q_col, const float* k, const float* v, int R,
                                 int Nt, __nv_bfloat16* out, cudaStream_t stream) {
  if (Nt > 32) {
    static PerDeviceOnce configured;
    if (const char* ce = set_max_smem_once(configured, i2t_attention_kernel, 64 * 320 * 4)) return ce;
  }
  launch_pdl(i2t_attention_kernel, dim3(128, R), dim3(256), Nt * 320 * sizeof(float), stream, 1, proj, ldp, q_col, k, v, Nt, out);
  return last_err();
}

const char* launch_keys_init(const float* emb, int R, int P, const float* dense, int dense_mode, int dense_batch,
                             float* keys_f32, __nv_bfloat16* keys_bf16, cudaStream_t stream) {
  launch_pdl(keys_init_kernel, dim3(128, 8, R), dim3(256), 0, stream, 1, emb, P, dense, dense_mode, dense_batch, keys_f32, keys_bf16);
  return last_err();
}

const char* launch_tok_init(const float* iou_token, const float* mask_tokens, int nm, const float* sparse, int Ns, int R,
                            float* tokens, cudaStream_t stream) {
  launch_pdl(tok_init_kernel, dim3(R), dim3(256), 0, stream, 1, iou_token, mask_tokens, nm, sparse, Ns, tokens);
  return last_err();
}

const char* launch_slice_copy(const float* in, int R, int C, int m0, int cnt, size_t inner, float* out, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(R) * cnt * inner;
  if (total == 0) return nullptr;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl(slice_copy_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, 1, in, C, m0, cnt, inner, out, total);
  return last_err();
}

}  // namespace ullsam
