// ullsam_b200 — the prompt-token side of the mask decoder as ONE kernel per TwoWayAttentionBlock (sm_100a).
//
// Reference: modeling/transformer.py:109-196 (TwoWayAttentionBlock.forward), :62-108 (final attention), and
// modeling/mask_decoder.py:136-148 (hypernetwork MLPs, IoU head).  Everything that touches only the Nt <= 24 prompt tokens of
// one prompt — 9 small linears, the token self-attention, the token->image attention over the 4096 image tokens, the MLP and
// three LayerNorms per block — used to be 15 dependent launches of 3-45 us each (41 launches, 0.77 ms per 32 prompts).
//
// Layout of the work: a thread-block CLUSTER of four CTAs per prompt.  Every CTA keeps the full token state (queries, the
// point embedding, the row being normalised) in its own shared memory; each linear is split by OUTPUT COLUMNS over the four
// CTAs (so each weight matrix is read once per prompt, a quarter per SM) and the results are written straight into the
// shared memory of all four CTAs (st.shared::cluster), followed by one cluster barrier.  The column split follows the
// heads: CTA c owns self-attention heads 2c, 2c+1 (64 of 256 columns of q, k, v) and token->image heads 2c, 2c+1 (32 of 128
// columns), so the attention inputs never leave the CTA.  The MLP is split along the HIDDEN dimension instead (each CTA
// holds a quarter of the 2048-wide activation and produces a partial 256-wide result; the partials are summed by
// reading the peers' shared memory).  fp32 throughout, like the launch-per-op path it replaces (decoder.cu).
#include "decoder.h"
#include "host_util.h"

#include "common.cuh"

#ifndef ULLSAM_DEC_TIMELINE
#define ULLSAM_DEC_TIMELINE 0
#endif
#if ULLSAM_DEC_TIMELINE
#define DT_STAMP(name) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.y == 3) { const long long t_ = clock64(); printf("  rank %d %-18s %8lld\n", rank, name, t_ - dt_last); dt_last = t_; } } while (0)
#define DT_BEGIN long long dt_last = clock64()
#else
#define DT_STAMP(name) do { } while (0)
#define DT_BEGIN do { } while (0)
#endif

namespace ullsam {

namespace {

constexpr int FT = 512;        // threads per CTA
constexpr int FW = FT / 32;    // warps
constexpr int CL = 4;          // CTAs per prompt
constexpr int E = 256;         // transformer width
constexpr int T2I_TG = 4;      // tokens per pass of the token->image attention
// floats per token row in shared memory: Bq, Bpe, Bqp, Ba (aliased by the MLP partial), By: 5 x 256; Lq/Lk/Lv: 3 x 64; hid: 512
constexpr int ROW_FLOATS = 5 * E + 3 * 64 + 512;
constexpr int FIXED_FLOATS = 2 * T2I_TG * 16 + 2 * 8 * T2I_TG * 18;   // t2i: q of the pass, per-warp partial results

ULLSAM_DEVICE void st_cluster(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
ULLSAM_DEVICE float ld_cluster(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// 8 partial sums per lane -> the lanes with (lane & 3) == 0 hold the warp total of row 4*bit4 + 2*bit3 + bit2 in a[0]
// (halving exchange: 4 + 2 + 1 + 1 + 1 shuffles instead of 8 x 5)
ULLSAM_DEVICE void reduce8(float (&a)[8], int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 16;
    const float send = up ? a[i] : a[i + 4], keep = up ? a[i + 4] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 8;
    const float send = up ? a[i] : a[i + 2], keep = up ? a[i + 2] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const bool up = lane & 4;
    const float send = up ? a[0] : a[1], keep = up ? a[1] : a[0];
    a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2);
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
}

// y[r][c] = sum_k xs[r][k] * W[c][k] for the 8 rows at xs (row stride ldx floats) and columns c in [0, ncols) of the
// row-major matrix W (row stride ldw floats, K a multiple of 128); emit(r, c, value) is called by ONE lane per (r, c).
// A warp takes two columns at a time: the lanes split K in float4 steps (coalesced weight reads, conflict-free x reads).
template <class Emit>
ULLSAM_DEVICE void gemv8(const float* xs, int ldx, int K, const float* __restrict__ W, int ldw, int ncols, int warp, int lane,
                         Emit emit) {
  for (int c = 2 * warp; c < ncols; c += 2 * FW) {
    const int c1 = (c + 1 < ncols) ? c + 1 : c;
    const float4* w0 = reinterpret_cast<const float4*>(W + static_cast<size_t>(c) * ldw);
    const float4* w1 = reinterpret_cast<const float4*>(W + static_cast<size_t>(c1) * ldw);
    float a0[8], a1[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) a0[r] = a1[r] = 0.f;
    for (int k4 = lane; k4 < K / 4; k4 += 32) {
      const float4 u = __ldg(w0 + k4), v = __ldg(w1 + k4);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(xs + r * ldx + 4 * k4);
        a0[r] = fmaf(x.x, u.x, fmaf(x.y, u.y, fmaf(x.z, u.z, fmaf(x.w, u.w, a0[r]))));
        a1[r] = fmaf(x.x, v.x, fmaf(x.y, v.y, fmaf(x.z, v.z, fmaf(x.w, v.w, a1[r]))));
      }
    }
    reduce8(a0, lane);
    reduce8(a1, lane);
    if ((lane & 3) == 0) {
      const int r = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
      emit(r, c, a0[0]);
      if (c1 != c) emit(r, c1, a1[0]);
    }
  }
}

// one output row: y[c] = act(x . W[c] + b[c]), x in shared memory (K a multiple of 128); warp per column
template <class Emit>
ULLSAM_DEVICE void gemv1(const float* xs, int K, const float* __restrict__ W, int ncols, int warp, int lane, Emit emit) {
  for (int c = warp; c < ncols; c += FW) {
    const float4* w = reinterpret_cast<const float4*>(W + static_cast<size_t>(c) * K);
    float a = 0.f;
    for (int k4 = lane; k4 < K / 4; k4 += 32) {
      const float4 u = __ldg(w + k4);
      const float4 x = *reinterpret_cast<const float4*>(xs + 4 * k4);
      a = fmaf(x.x, u.x, fmaf(x.y, u.y, fmaf(x.z, u.z, fmaf(x.w, u.w, a))));
    }
    a = warp_sum(a);
    if (lane == 0) emit(c, a);
  }
}

struct Smem {
  float *Bq, *Bpe, *Bqp, *Ba, *By, *Lq, *Lk, *Lv, *hid, *part, *qs, *red;
};

// value -> the same location of all four CTAs of the cluster
ULLSAM_DEVICE void bcast(const Smem& s, float* p, float v) {
  const uint32_t a = smem_u32(p);
#pragma unroll
  for (int k = 0; k < CL; ++k) st_cluster(mapa_shared(a, k), v);
}

// Bq[row][:] = LayerNorm(By[row][:]) (eps 1e-5), Bqp = Bq + Bpe; warp per row; rows >= nt are left alone (they stay zero)
ULLSAM_DEVICE void layer_norm_rows(const Smem& s, const float* __restrict__ gamma, const float* __restrict__ beta, int nt,
                                   int warp, int lane) {
  for (int row = warp; row < nt; row += FW) {
    const float4* y = reinterpret_cast<const float4*>(s.By + row * E);
    const float4 a = y[lane], b = y[32 + lane];
    const float mean = warp_sum((a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w)) * (1.0f / E);
    float sq = (a.x - mean) * (a.x - mean) + (a.y - mean) * (a.y - mean) + (a.z - mean) * (a.z - mean) + (a.w - mean) * (a.w - mean);
    sq += (b.x - mean) * (b.x - mean) + (b.y - mean) * (b.y - mean) + (b.z - mean) * (b.z - mean) + (b.w - mean) * (b.w - mean);
    const float rstd = 1.0f / sqrtf(warp_sum(sq) * (1.0f / E) + 1e-5f);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + lane), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 32 + lane);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + lane), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 32 + lane);
    const float4 o0 = make_float4((a.x - mean) * rstd * g0.x + b0.x, (a.y - mean) * rstd * g0.y + b0.y,
                                  (a.z - mean) * rstd * g0.z + b0.z, (a.w - mean) * rstd * g0.w + b0.w);
    const float4 o1 = make_float4((b.x - mean) * rstd * g1.x + b1.x, (b.y - mean) * rstd * g1.y + b1.y,
                                  (b.z - mean) * rstd * g1.z + b1.z, (b.w - mean) * rstd * g1.w + b1.w);
    float4* q = reinterpret_cast<float4*>(s.Bq + row * E);
    q[lane] = o0;
    q[32 + lane] = o1;
    const float4* pe = reinterpret_cast<const float4*>(s.Bpe + row * E);
    const float4 p0 = pe[lane], p1 = pe[32 + lane];
    float4* qp = reinterpret_cast<float4*>(s.Bqp + row * E);
    qp[lane] = make_float4(o0.x + p0.x, o0.y + p0.y, o0.z + p0.z, o0.w + p0.w);
    qp[32 + lane] = make_float4(o1.x + p1.x, o1.y + p1.y, o1.z + p1.z, o1.w + p1.w);
  }
}

// tokens -> image attention (transformer.py:165-170 / :101-106), internal width 128 = 8 heads x 16, followed by the
// output projection, the residual add and the LayerNorm.  CTA `rank` owns heads 2 rank, 2 rank + 1.
//   q = (queries + pe) Wq^T + bq;  o = softmax(q K^T / 4) V;  queries = LN(queries + o Wo^T + bo)
// K, V: bf16 columns [0,128) / [128,256) of kv[4096][ldkv] (this prompt's rows).  Thread halves 0..255 / 256..511 take one
// head each: the 4096 keys are spread over the 256 threads, each keeps an online softmax for up to TG tokens per pass.
ULLSAM_DEVICE void t2i_block(const Smem& s, int rank, int nt, int nr, const float* __restrict__ q_w, const float* __restrict__ q_b,
                             const float* __restrict__ o_w, const float* __restrict__ o_b, const float* __restrict__ n_w,
                             const float* __restrict__ n_b, const __nv_bfloat16* __restrict__ kv, int ldkv, int warp, int lane) {
  constexpr int TG = T2I_TG;
  // q slice: columns [32 rank, 32 rank + 32), scaled by 1/sqrt(16)
  for (int rb = 0; rb < nr; rb += 8)
    gemv8(s.Bqp + rb * E, E, E, q_w + static_cast<size_t>(32 * rank) * E, E, 32, warp, lane,
          [&](int r, int c, float v) { s.Lq[(rb + r) * 64 + c] = (v + __ldg(q_b + 32 * rank + c)) * 0.25f; });
  __syncthreads();
  const int hh = threadIdx.x >> 8, t = threadIdx.x & 255;     // local head, thread inside the half
  const int head = 2 * rank + hh, hw = warp & 7;
  float* qs = s.qs + hh * TG * 16;
  float* red = s.red + hh * 8 * TG * 18;
  for (int i0 = 0; i0 < nt; i0 += TG) {
    const int ng = (nt - i0 < TG) ? (nt - i0) : TG;
    __syncthreads();
    if (t < TG * 16) {
      const int i = t >> 4, d = t & 15;
      qs[i * 16 + d] = (i < ng) ? s.Lq[(i0 + i) * 64 + hh * 16 + d] : 0.f;
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
      const int key = jj * 256 + t;
      const uint4* kp = reinterpret_cast<const uint4*>(kv + static_cast<size_t>(key) * ldkv + head * 16);
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
        for (int d = 0; d < 16; ++d) sc = fmaf(qs[i * 16 + d], kf[d], sc);
        const float mn = fmaxf(m[i], sc);
        const float corr = __expf(m[i] - mn), p = __expf(sc - mn);
        l[i] = l[i] * corr + p;
#pragma unroll
        for (int d = 0; d < 16; ++d) acc[i][d] = fmaf(acc[i][d], corr, p * vf[d]);
        m[i] = mn;
      }
    }
    // merge: inside the warp first (common max, rescale, sum), then the 8 warps of the half through shared memory
#pragma unroll
    for (int i = 0; i < TG; ++i) {
      const float wm = warp_max(m[i]);
      const float sc = __expf(m[i] - wm);
      const float ls = warp_sum(l[i] * sc);
      if (lane == 0) { red[(hw * TG + i) * 18] = wm; red[(hw * TG + i) * 18 + 1] = ls; }
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const float a = warp_sum(acc[i][d] * sc);
        if (lane == 0) red[(hw * TG + i) * 18 + 2 + d] = a;
      }
    }
    __syncthreads();
    if (t < TG * 16) {
      const int i = t >> 4, d = t & 15;
      if (i < ng) {
        float gm = red[i * 18];
#pragma unroll
        for (int w = 1; w < 8; ++w) gm = fmaxf(gm, red[(w * TG + i) * 18]);
        float L = 0.f, A = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float sc = __expf(red[(w * TG + i) * 18] - gm);
          L = fmaf(red[(w * TG + i) * 18 + 1], sc, L);
          A = fmaf(red[(w * TG + i) * 18 + 2 + d], sc, A);
        }
        bcast(s, s.Ba + (i0 + i) * E + head * 16 + d, A / L);   // attention output, full 128 columns in every CTA
      }
    }
  }
  cluster_sync();
  // out_proj (K = 128) + residual: columns [64 rank, 64 rank + 64) -> By of every CTA
  for (int rb = 0; rb < nr; rb += 8)
    gemv8(s.Ba + rb * E, E, 128, o_w + static_cast<size_t>(64 * rank) * 128, 128, 64, warp, lane, [&](int r, int c, float v) {
      const int row = rb + r, n = 64 * rank + c;
      bcast(s, s.By + row * E + n, v + __ldg(o_b + n) + s.Bq[row * E + n]);
    });
  cluster_sync();
  layer_norm_rows(s, n_w, n_b, nt, warp, lane);
  __syncthreads();
}

ULLSAM_DEVICE Smem carve(float* base, int nr) {
  Smem s;
  s.Bq = base;
  s.Bpe = s.Bq + nr * E;
  s.Bqp = s.Bpe + nr * E;
  s.Ba = s.Bqp + nr * E;
  s.part = s.Ba;                 // the MLP partial reuses the attention-output rows (dead by then)
  s.By = s.Ba + nr * E;
  s.Lq = s.By + nr * E;
  s.Lk = s.Lq + nr * 64;
  s.Lv = s.Lk + nr * 64;
  s.hid = s.Lv + nr * 64;
  s.qs = s.hid + nr * 512;
  s.red = s.qs + 2 * T2I_TG * 16;
  return s;
}

// queries / point embedding of prompt r -> shared memory (rows >= nt zero), Bqp = Bq (+ Bpe); the exchange buffers start
// out zero (their rows >= nt are read by the 8-row GEMV blocks); ends with the cluster barrier that must precede the first
// write into a peer's shared memory (every CTA of the cluster is running and has finished clearing)
ULLSAM_DEVICE void load_tokens(const Smem& s, const float* __restrict__ queries, const float* __restrict__ tokens, int r, int nt,
                               int nr, bool add_pe) {
  for (int idx = threadIdx.x; idx < 2 * nr * (E / 4); idx += FT) reinterpret_cast<float4*>(s.Ba)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int idx = threadIdx.x; idx < nr * (512 / 4); idx += FT) reinterpret_cast<float4*>(s.hid)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int idx = threadIdx.x; idx < nr * (E / 4); idx += FT) {
    const int row = idx / (E / 4);
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), p = q;
    if (row < nt) {
      q = __ldg(reinterpret_cast<const float4*>(queries + static_cast<size_t>(r) * nt * E) + idx);
      p = __ldg(reinterpret_cast<const float4*>(tokens + static_cast<size_t>(r) * nt * E) + idx);
    }
    reinterpret_cast<float4*>(s.Bq)[idx] = q;
    reinterpret_cast<float4*>(s.Bpe)[idx] = p;
    reinterpret_cast<float4*>(s.Bqp)[idx] = add_pe ? make_float4(q.x + p.x, q.y + p.y, q.z + p.z, q.w + p.w) : q;
  }
  cluster_sync();
}

ULLSAM_DEVICE void store_queries(const Smem& s, float* __restrict__ queries, int r, int nt) {
  for (int idx = threadIdx.x; idx < nt * (E / 4); idx += FT)
    reinterpret_cast<float4*>(queries + static_cast<size_t>(r) * nt * E)[idx] = reinterpret_cast<const float4*>(s.Bq)[idx];
}

// ---------------------------------------------------------------------------------------------
// One TwoWayAttentionBlock, token side (transformer.py:155-196 except the image->token attention core and its
// out_proj + norm4, which run on the image side): self attention, norm1, tokens->image attention, norm2, MLP, norm3,
// and the k / v projections of the tokens for the image->token attention.
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(FT, 1)
dec_token_layer_kernel(DecTokLayer L, float* __restrict__ queries, const float* __restrict__ tokens,
                       const __nv_bfloat16* __restrict__ proj, int ldkv, int nt, int nr, float* __restrict__ k128,
                       float* __restrict__ v128) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem s = carve(smem_f, nr);
  const int rank = static_cast<int>(cluster_ctarank()), r = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DT_BEGIN;
  load_tokens(s, queries, tokens, r, nt, nr, !L.skip_pe);
  DT_STAMP("load");

  // ---- self attention (transformer.py:156-163): heads 2 rank, 2 rank + 1 = columns [64 rank, 64 rank + 64) ----
  const int n0 = 64 * rank;
  for (int rb = 0; rb < nr; rb += 8) {
    gemv8(s.Bqp + rb * E, E, E, L.sq_w + static_cast<size_t>(n0) * E, E, 64, warp, lane,
          [&](int rr, int c, float v) { s.Lq[(rb + rr) * 64 + c] = (v + __ldg(L.sq_b + n0 + c)) * 0.17677669529663687f; });  // 1/sqrt(32)
    gemv8(s.Bqp + rb * E, E, E, L.sk_w + static_cast<size_t>(n0) * E, E, 64, warp, lane,
          [&](int rr, int c, float v) { s.Lk[(rb + rr) * 64 + c] = v + __ldg(L.sk_b + n0 + c); });
    gemv8(s.Bq + rb * E, E, E, L.sv_w + static_cast<size_t>(n0) * E, E, 64, warp, lane,
          [&](int rr, int c, float v) { s.Lv[(rb + rr) * 64 + c] = v + __ldg(L.sv_b + n0 + c); });
  }
  __syncthreads();
  DT_STAMP("self qkv");
  // one warp per (token, local head): lane = key for the scores, lane = output dim for P V  (nt <= 32)
  for (int w = warp; w < nt * 2; w += FW) {
    const int i = w >> 1, hl = w & 1;
    float sc = -INFINITY;
    if (lane < nt) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) a = fmaf(s.Lq[i * 64 + hl * 32 + d], s.Lk[lane * 64 + hl * 32 + d], a);
      sc = a;
    }
    const float mx = warp_max(sc);
    const float p = (lane < nt) ? expf(sc - mx) : 0.f;
    const float inv = 1.0f / warp_sum(p);
    float acc = 0.f;
    for (int j = 0; j < nt; ++j) acc = fmaf(__shfl_sync(0xffffffffu, p, j), s.Lv[j * 64 + hl * 32 + lane], acc);
    bcast(s, s.Ba + i * E + n0 + hl * 32 + lane, acc * inv);
  }
  cluster_sync();
  DT_STAMP("self attn");
  // out_proj (+ residual unless this is the first layer: there the attention output REPLACES the queries, :157-158)
  for (int rb = 0; rb < nr; rb += 8)
    gemv8(s.Ba + rb * E, E, E, L.so_w + static_cast<size_t>(n0) * E, E, 64, warp, lane, [&](int rr, int c, float v) {
      const int row = rb + rr, n = n0 + c;
      bcast(s, s.By + row * E + n, v + __ldg(L.so_b + n) + (L.skip_pe ? 0.f : s.Bq[row * E + n]));
    });
  cluster_sync();
  DT_STAMP("self out_proj");
  layer_norm_rows(s, L.n1_w, L.n1_b, nt, warp, lane);
  __syncthreads();
  DT_STAMP("norm1");

  // ---- tokens attend to the image, norm2 (:165-170) ----
  t2i_block(s, rank, nt, nr, L.tq_w, L.tq_b, L.to_w, L.to_b, L.n2_w, L.n2_b, proj + static_cast<size_t>(r) * 4096 * ldkv, ldkv, warp,
            lane);

  DT_STAMP("t2i block");
  // ---- MLP (:172-175): this CTA's quarter of the hidden units, then its partial of lin2 over that quarter ----
  const int hq = L.mlp_dim / CL;   // 512
  for (int rb = 0; rb < nr; rb += 8)
    gemv8(s.Bq + rb * E, E, E, L.m1_w + static_cast<size_t>(hq * rank) * E, E, hq, warp, lane,
          [&](int rr, int c, float v) { s.hid[(rb + rr) * 512 + c] = fmaxf(v + __ldg(L.m1_b + hq * rank + c), 0.f); });
  __syncthreads();
  DT_STAMP("mlp lin1");
  for (int rb = 0; rb < nr; rb += 8)
    gemv8(s.hid + rb * 512, 512, hq, L.m2_w + hq * rank, L.mlp_dim, E, warp, lane,
          [&](int rr, int c, float v) { s.part[(rb + rr) * E + c] = v; });
  cluster_sync();
  DT_STAMP("mlp lin2");
  for (int idx = threadIdx.x; idx < nr * 64; idx += FT) {
    const int row = idx >> 6, n = n0 + (idx & 63);
    const uint32_t a = smem_u32(s.part + row * E + n);
    float v = __ldg(L.m2_b + n) + s.Bq[row * E + n];
#pragma unroll
    for (int k = 0; k < CL; ++k) v += ld_cluster(mapa_shared(a, k));
    bcast(s, s.By + row * E + n, v);   // (By was last read by norm2, before the barrier above; the partials live in Ba)
  }
  cluster_sync();
  DT_STAMP("mlp reduce");
  layer_norm_rows(s, L.n3_w, L.n3_b, nt, warp, lane);
  __syncthreads();
  DT_STAMP("norm3");

  // ---- k / v of the tokens for the image->token attention (:177-180): 128 columns, 32 per CTA ----
  for (int rb = 0; rb < nr; rb += 8) {
    gemv8(s.Bqp + rb * E, E, E, L.ik_w + static_cast<size_t>(32 * rank) * E, E, 32, warp, lane, [&](int rr, int c, float v) {
      if (rb + rr < nt) k128[(static_cast<size_t>(r) * nt + rb + rr) * 128 + 32 * rank + c] = v + __ldg(L.ik_b + 32 * rank + c);
    });
    gemv8(s.Bq + rb * E, E, E, L.iv_w + static_cast<size_t>(32 * rank) * E, E, 32, warp, lane, [&](int rr, int c, float v) {
      if (rb + rr < nt) v128[(static_cast<size_t>(r) * nt + rb + rr) * 128 + 32 * rank + c] = v + __ldg(L.iv_b + 32 * rank + c);
    });
  }
  DT_STAMP("i2t k/v");
  if (rank == 0) store_queries(s, queries, r, nt);
}

// ---------------------------------------------------------------------------------------------
// After the last block: final tokens->image attention + norm_final_attn (transformer.py:98-106), then — when hyper / iou
// are given — the four hypernetwork MLPs on the mask tokens (CTA c takes mask token c) and the IoU head on token 0
// (mask_decoder.py:136-148).
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(FT, 1)
dec_token_final_kernel(DecTokFinal F, float* __restrict__ queries, const float* __restrict__ tokens,
                       const __nv_bfloat16* __restrict__ proj, int ldkv, int nt, int nr, float* __restrict__ hyper,
                       float* __restrict__ iou4) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem s = carve(smem_f, nr);
  const int rank = static_cast<int>(cluster_ctarank()), r = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  load_tokens(s, queries, tokens, r, nt, nr, true);
  t2i_block(s, rank, nt, nr, F.tq_w, F.tq_b, F.to_w, F.to_b, F.nf_w, F.nf_b, proj + static_cast<size_t>(r) * 4096 * ldkv, ldkv, warp,
            lane);
  if (rank == 0) store_queries(s, queries, r, nt);
  if (hyper == nullptr) return;
  // hypernetwork MLP of mask token `rank` (row 1 + rank): 256 -> 256 -> 256 -> 32, ReLU between (mask_decoder.py:139-142)
  float* h1 = s.hid;
  float* h2 = s.hid + E;
  {
    const float* x = s.Bq + (1 + rank) * E;
    gemv1(x, E, F.hyp_w0 + static_cast<size_t>(rank) * E * E, E, warp, lane,
          [&](int c, float v) { h1[c] = fmaxf(v + __ldg(F.hyp_b0 + rank * E + c), 0.f); });
    __syncthreads();
    gemv1(h1, E, F.hyp_w1 + static_cast<size_t>(rank) * E * E, E, warp, lane,
          [&](int c, float v) { h2[c] = fmaxf(v + __ldg(F.hyp_b1 + rank * E + c), 0.f); });
    __syncthreads();
    gemv1(h2, E, F.hyp_w2 + static_cast<size_t>(rank) * 32 * E, 32, warp, lane,
          [&](int c, float v) { hyper[(static_cast<size_t>(r) * 4 + rank) * 32 + c] = v + __ldg(F.hyp_b2 + rank * 32 + c); });
  }
  // IoU head on token 0 (mask_decoder.py:145): depth-layer MLP, ReLU between, CTA 3 (the others are done)
  if (rank == CL - 1 && iou4 != nullptr) {
    __syncthreads();
    const float* x = s.Bq;
    float* bufs[2] = {s.hid + 2 * E, s.hid + 2 * E + 512};
    int K = E;
    for (int i = 0; i < F.iou_depth; ++i) {
      const bool last = (i == F.iou_depth - 1);
      const int N = last ? 4 : F.iou_hidden;
      float* y = bufs[i & 1];
      const float* b = F.iou_b[i];
      if (last) {
        gemv1(x, K, F.iou_w[i], N, warp, lane, [&](int c, float v) { iou4[static_cast<size_t>(r) * 4 + c] = v + __ldg(b + c); });
      } else {
        gemv1(x, K, F.iou_w[i], N, warp, lane, [&](int c, float v) { y[c] = fmaxf(v + __ldg(b + c), 0.f); });
      }
      __syncthreads();
      x = y;
      K = N;
    }
  }
}

size_t fused_smem_bytes(int nr) { return (static_cast<size_t>(nr) * ROW_FLOATS + FIXED_FLOATS) * sizeof(float); }

const char* check_launch() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

bool dec_tokens_fused_supported(int Nt, int mlp_dim, int iou_hidden) {
  return Nt >= 5 && Nt <= kDecFusedMaxTokens && mlp_dim == 2048 && iou_hidden <= 512 && iou_hidden % 128 == 0;
}

const char* launch_dec_token_layer(const DecTokLayer& L, float* queries, const float* tokens, const __nv_bfloat16* proj, int ldkv,
                                   int R, int Nt, float* k128, float* v128, cudaStream_t stream) {
  if (R == 0) return nullptr;
  const int nr = (Nt + 7) / 8 * 8;
  const size_t smem = fused_smem_bytes(nr);
  static PerDeviceOnce configured;
  if (const char* e = set_max_smem_once(configured, dec_token_layer_kernel, static_cast<int>(fused_smem_bytes(kDecFusedMaxTokens)))) return e;
  dec_token_layer_kernel<<<dim3(CL, R), FT, smem, stream>>>(L, queries, tokens, proj, ldkv, Nt, nr, k128, v128);
  return check_launch();
}

const char* launch_dec_token_final(const DecTokFinal& F, float* queries, const float* tokens, const __nv_bfloat16* proj, int ldkv,
                                   int R, int Nt, float* hyper, float* iou4, cudaStream_t stream) {
  if (R == 0) return nullptr;
  const int nr = (Nt + 7) / 8 * 8;
  const size_t smem = fused_smem_bytes(nr);
  static PerDeviceOnce configured;
  if (const char* e = set_max_smem_once(configured, dec_token_final_kernel, static_cast<int>(fused_smem_bytes(kDecFusedMaxTokens)))) return e;
  dec_token_final_kernel<<<dim3(CL, R), FT, smem, stream>>>(F, queries, tokens, proj, ldkv, Nt, nr, hyper, iou4);
  return check_launch();
}

}  // namespace ullsam
