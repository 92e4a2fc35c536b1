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
#if ULLSAM_DEC_TIMELINE   // debug builds (tools/build_variant.sh): per-phase cycle counts of one CTA, printed once at the end
#define DT_STAMP(name) do { __syncthreads(); if (dt_n < 16) { const long long t_ = clock64(); dt_name[dt_n] = name; dt_t[dt_n++] = t_ - dt_last; dt_last = clock64(); } } while (0)
#define DT_BEGIN long long dt_last = clock64(); long long dt_t[16]; const char* dt_name[16]; int dt_n = 0
#define DT_END do { if (threadIdx.x == 0 && blockIdx.y == 3 && rank == 1) for (int i_ = 0; i_ < dt_n; ++i_) printf("  %-16s %8lld\n", dt_name[i_], dt_t[i_]); } while (0)
#else
#define DT_STAMP(name) do { } while (0)
#define DT_BEGIN do { } while (0)
#define DT_END do { } while (0)
#endif

namespace ullsam {

namespace {

constexpr int FT = 512;        // threads per CTA
constexpr int FW = FT / 32;    // warps
constexpr int CL = 4;          // CTAs per prompt
constexpr int E = 256;         // transformer width
constexpr int T2I_KEYS = 128;  // keys per staged K / V block of the token->image attention
constexpr int T2I_STAGE_FLOATS = 2 * 2 * T2I_KEYS * 16 / 2;   // K and V of this CTA's two heads, bf16: 16 KB = 4096 floats
// floats per token row in shared memory: Bq, Bpe, Bqp, Ba (aliased by the MLP partial), By: 5 x 256; Lq/Lk/Lv: 3 x 64; hid: 512
constexpr int ROW_FLOATS = 5 * E + 3 * 64 + 512;
constexpr int FIXED_FLOATS = 2 * 8 * 8 * 18;   // t2i: per-warp partial results [head][warp][token lane][max, sum, acc[16]]

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
// (K is a template parameter: the weight loads of a column pair — K / 64 float4 per lane — are all issued before the first
//  FMA, and the next pair's loads are issued before this pair is reduced: the loop is a chain of L2 round trips otherwise.)
template <int K, class Emit>
ULLSAM_DEVICE void gemv8(const float* xs, int ldx, const float* __restrict__ W, int ldw, int ncols, int warp, int lane,
                         Emit emit) {
  constexpr int KI = K / 128;
  float4 u[KI], v[KI];
  auto load_pair = [&](int c) {
    const int c1 = (c + 1 < ncols) ? c + 1 : c;
    const float4* w0 = reinterpret_cast<const float4*>(W + static_cast<size_t>(c) * ldw);
    const float4* w1 = reinterpret_cast<const float4*>(W + static_cast<size_t>(c1) * ldw);
#pragma unroll
    for (int i = 0; i < KI; ++i) {
      u[i] = __ldg(w0 + lane + 32 * i);
      v[i] = __ldg(w1 + lane + 32 * i);
    }
  };
  if (2 * warp < ncols) load_pair(2 * warp);
  for (int c = 2 * warp; c < ncols; c += 2 * FW) {
    const int c1 = (c + 1 < ncols) ? c + 1 : c;
    float a0[8], a1[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) a0[r] = a1[r] = 0.f;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(xs + r * ldx + 4 * (lane + 32 * i));
        a0[r] = fmaf(x.x, u[i].x, fmaf(x.y, u[i].y, fmaf(x.z, u[i].z, fmaf(x.w, u[i].w, a0[r]))));
        a1[r] = fmaf(x.x, v[i].x, fmaf(x.y, v[i].y, fmaf(x.z, v[i].z, fmaf(x.w, v[i].w, a1[r]))));
      }
    }
    if (c + 2 * FW < ncols) load_pair(c + 2 * FW);   // in flight during the reduction and the emit
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
  float *Bq, *Bpe, *Bqp, *Ba, *By, *Lq, *Lk, *Lv, *hid, *part, *red, *stage;
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
// K, V: bf16 columns [0,128) / [128,256) of kv[4096][ldkv] (this prompt's rows); this CTA needs 64 + 64 bytes of each row.
// They are STAGED: blocks of 128 keys (16 KB) travel global -> shared memory with cp.async through a ring of `ns` buffers, so
// ~100 KB per SM are in flight and the loop runs from shared memory (the first version loaded K / V into registers one key
// per thread and iteration and spent 110 us per call waiting for them; a second, SIMT one from shared memory still needed
// ~55 us: 8 token lanes repeated the bf16 -> fp32 conversion of every key).  ONE pass over the keys serves all tokens.
// NG = number of 16-token tiles (1 for up to 16 tokens, 2 for up to 24).
template <int NG>
ULLSAM_DEVICE void t2i_attention(const Smem& s, int rank, int nt, int nr, const __nv_bfloat16* __restrict__ kv, int ldkv, int ns,
                                 int warp, int lane) {
  constexpr int NBLK = 4096 / T2I_KEYS;
  const int t = threadIdx.x;
  // ---- producer side: every thread copies two 16-byte chunks per block ----
  auto issue = [&](int blk) {
    uint8_t* stage = reinterpret_cast<uint8_t*>(s.stage + (blk % ns) * T2I_STAGE_FLOATS);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = t + FT * i;                 // 1024 chunks: [K | V][128 keys][2 heads][2 halves of 16 bytes]
      const int isv = c >> 9, idx = c & 511, key = idx >> 2, hl = (idx >> 1) & 1, part = idx & 1;
      const __nv_bfloat16* src = kv + static_cast<size_t>(blk * T2I_KEYS + key) * ldkv + isv * 128 + (2 * rank + hl) * 16 + part * 8;
      const uint32_t dst = smem_u32(stage + isv * 8192 + hl * 4096 + key * 32 + part * 16);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int b = 0; b < ns - 1; ++b) issue(b);
  // ---- consumer side: warp-level tensor-core MMAs (mma.sync m16n8k16, bf16 operands, fp32 accumulate) ----
  // Warps 0..7 serve head 2 rank, warps 8..15 head 2 rank + 1; warp w of a head takes keys [16 w, 16 w + 16) of every staged
  // block.  Per 16-key chunk and 16-token tile:  S = Q K^T (A = Q split into bf16 hi + lo parts so the fp32 queries lose
  // nothing, B = K^T straight from the staged rows), online softmax on the accumulator fragment (a token row lives in the
  // four lanes of a quad), O += P V (A = the P fragment re-packed as bf16, B = V through ldmatrix.trans).
  const int hh = warp >> 3, hw = warp & 7, g = lane >> 2, tq = lane & 3;
  uint32_t qhi[NG][4], qlo[NG][4];
  float m[NG][2], l[NG][2], o[NG][2][4];   // rows g and g + 8 of tile mt; o[mt][dim tile][c0..c3]
#pragma unroll
  for (int mt = 0; mt < NG; ++mt) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = mt * 16 + g + (i & 1) * 8, d = 2 * tq + (i >> 1) * 8;   // a0: (g, 2tq) a1: (g+8, 2tq) a2: (g, 2tq+8) a3: (g+8, 2tq+8)
      const bool live = row < nr;   // (a 16-row tile may reach past the rows held in shared memory: padding tokens)
      const float x0 = live ? s.Lq[row * 64 + hh * 16 + d] : 0.f, x1 = live ? s.Lq[row * 64 + hh * 16 + d + 1] : 0.f;
      const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
      qhi[mt][i] = pack_bf16(__bfloat162float(h0), __bfloat162float(h1));
      qlo[mt][i] = pack_bf16(x0 - __bfloat162float(h0), x1 - __bfloat162float(h1));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      m[mt][h] = -INFINITY;
      l[mt][h] = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) o[mt][h][i] = 0.f;
    }
  }
  auto mma = [](float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  };
#pragma unroll 1
  for (int blk = 0; blk < NBLK; ++blk) {
    if (blk + ns - 1 < NBLK) issue(blk + ns - 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");   // (keeps the group count uniform)
    // all but the newest ns - 1 groups have landed, i.e. block blk; ns is a run-time value, the wait count an immediate
    if (ns >= 6) asm volatile("cp.async.wait_group 5;" ::: "memory");
    else if (ns >= 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
    else asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const uint8_t* stage = reinterpret_cast<const uint8_t*>(s.stage + (blk % ns) * T2I_STAGE_FLOATS);
    const uint8_t* kb = stage + hh * 4096 + (hw * 16) * 32;          // this warp's 16 keys: K rows of 32 bytes
    const uint8_t* vb = stage + 8192 + hh * 4096 + (hw * 16) * 32;   // V rows
    // B fragments of K^T for the two 8-key tiles: b0 = K[key g][dims 2tq, 2tq+1], b1 = dims + 8
    uint32_t kf[2][2];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      kf[n][0] = *reinterpret_cast<const uint32_t*>(kb + (n * 8 + g) * 32 + 4 * tq);
      kf[n][1] = *reinterpret_cast<const uint32_t*>(kb + (n * 8 + g) * 32 + 16 + 4 * tq);
    }
    // B fragments of V for the two 8-dim tiles (k = the 16 keys): ldmatrix.x2.trans on rows = keys, 16 bytes = 8 dims
    uint32_t vf[2][2];
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const uint32_t addr = smem_u32(vb + (lane & 15) * 32 + n * 16);
      asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(vf[n][0]), "=r"(vf[n][1]) : "r"(addr));
    }
#pragma unroll
    for (int mt = 0; mt < NG; ++mt) {
      float sc[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
        mma(sc[n], qhi[mt], kf[n][0], kf[n][1]);
        mma(sc[n], qlo[mt], kf[n][0], kf[n][1]);
      }
      uint32_t pa[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {   // row g (h = 0: c0, c1) and row g + 8 (h = 1: c2, c3)
        float mx = fmaxf(fmaxf(sc[0][2 * h], sc[0][2 * h + 1]), fmaxf(sc[1][2 * h], sc[1][2 * h + 1]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float mn = fmaxf(m[mt][h], mx);
        const float corr = __expf(m[mt][h] - mn);
        m[mt][h] = mn;
        const float p00 = __expf(sc[0][2 * h] - mn), p01 = __expf(sc[0][2 * h + 1] - mn);
        const float p10 = __expf(sc[1][2 * h] - mn), p11 = __expf(sc[1][2 * h + 1] - mn);
        l[mt][h] = l[mt][h] * corr + ((p00 + p01) + (p10 + p11));   // this lane's share of the row sum
        o[mt][0][2 * h] *= corr; o[mt][0][2 * h + 1] *= corr;
        o[mt][1][2 * h] *= corr; o[mt][1][2 * h + 1] *= corr;
        pa[h] = pack_bf16(p00, p01);        // a0 / a1: keys 2tq, 2tq+1 of the first key tile
        pa[2 + h] = pack_bf16(p10, p11);    // a2 / a3: keys 8 + 2tq, ...
      }
      mma(o[mt][0], pa, vf[0][0], vf[0][1]);
      mma(o[mt][1], pa, vf[1][0], vf[1][1]);
    }
    __syncthreads();   // the buffer is refilled by the next iteration's issue()
  }
  // ---- merge the 8 warps of a head through shared memory: per token row (max, sum, acc[16]) ----
  float* red = s.red + hh * 8 * 8 * 18;   // [warp][row 0..7 of the current half-tile][18]
#pragma unroll
  for (int mt = 0; mt < NG; ++mt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float ls = l[mt][h];
      ls += __shfl_xor_sync(0xffffffffu, ls, 1);
      ls += __shfl_xor_sync(0xffffffffu, ls, 2);
      if (mt + h > 0) __syncthreads();   // the previous rows' partials have been consumed
      float* rw = red + (hw * 8 + g) * 18;
      if (tq == 0) { rw[0] = m[mt][h]; rw[1] = ls; }
      rw[2 + 2 * tq] = o[mt][0][2 * h];          // dims 2tq, 2tq+1 of the first dim tile
      rw[2 + 2 * tq + 1] = o[mt][0][2 * h + 1];
      rw[2 + 8 + 2 * tq] = o[mt][1][2 * h];
      rw[2 + 8 + 2 * tq + 1] = o[mt][1][2 * h + 1];
      __syncthreads();
      const int tl = threadIdx.x & 255;
      if (tl < 128) {
        const int i = tl >> 4, d = tl & 15, row = mt * 16 + h * 8 + i;
        if (row < nt) {
          float gm = red[i * 18];
#pragma unroll
          for (int w = 1; w < 8; ++w) gm = fmaxf(gm, red[(w * 8 + i) * 18]);
          float L = 0.f, A = 0.f;
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const float e = __expf(red[(w * 8 + i) * 18] - gm);
            L = fmaf(red[(w * 8 + i) * 18 + 1], e, L);
            A = fmaf(red[(w * 8 + i) * 18 + 2 + d], e, A);
          }
          bcast(s, s.Ba + row * E + (2 * rank + hh) * 16 + d, A / L);   // attention output, full 128 columns in every CTA
        }
      }
    }
  }
}

ULLSAM_DEVICE void t2i_block(const Smem& s, int rank, int nt, int nr, const float* __restrict__ q_w, const float* __restrict__ q_b,
                             const float* __restrict__ o_w, const float* __restrict__ o_b, const float* __restrict__ n_w,
                             const float* __restrict__ n_b, const __nv_bfloat16* __restrict__ kv, int ldkv, int ns, int warp,
                             int lane) {
  // q slice: columns [32 rank, 32 rank + 32), scaled by 1/sqrt(16)
  for (int rb = 0; rb < nr; rb += 8)
    gemv8<E>(s.Bqp + rb * E, E, q_w + static_cast<size_t>(32 * rank) * E, E, 32, warp, lane,
          [&](int r, int c, float v) { s.Lq[(rb + r) * 64 + c] = (v + __ldg(q_b + 32 * rank + c)) * 0.25f; });
  __syncthreads();
  if (nr <= 16) t2i_attention<1>(s, rank, nt, nr, kv, ldkv, ns, warp, lane);   // one or two 16-token tiles
  else t2i_attention<2>(s, rank, nt, nr, kv, ldkv, ns, warp, lane);
  cluster_sync();
  // out_proj (K = 128) + residual: columns [64 rank, 64 rank + 64) -> By of every CTA
  for (int rb = 0; rb < nr; rb += 8)
    gemv8<128>(s.Ba + rb * E, E, o_w + static_cast<size_t>(64 * rank) * 128, 128, 64, warp, lane, [&](int r, int c, float v) {
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
  s.red = s.hid + nr * 512;
  s.stage = s.red + FIXED_FLOATS;   // ns buffers of T2I_STAGE_FLOATS (16-byte aligned: every term above is a multiple of 4 floats)
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
                       const __nv_bfloat16* __restrict__ proj, int ldkv, int nt, int nr, int ns, float* __restrict__ k128,
                       float* __restrict__ v128) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem s = carve(smem_f, nr);
  const int rank = static_cast<int>(cluster_ctarank()), r = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DT_BEGIN;
  pdl_wait();
  pdl_trigger();
  load_tokens(s, queries, tokens, r, nt, nr, !L.skip_pe);
  DT_STAMP("load");

  // ---- self attention (transformer.py:156-163): heads 2 rank, 2 rank + 1 = columns [64 rank, 64 rank + 64) ----
  const int n0 = 64 * rank;
  for (int rb = 0; rb < nr; rb += 8) {
    gemv8<E>(s.Bqp + rb * E, E, L.sq_w + static_cast<size_t>(n0) * E, E, 64, warp, lane,
          [&](int rr, int c, float v) { s.Lq[(rb + rr) * 64 + c] = (v + __ldg(L.sq_b + n0 + c)) * 0.17677669529663687f; });  // 1/sqrt(32)
    gemv8<E>(s.Bqp + rb * E, E, L.sk_w + static_cast<size_t>(n0) * E, E, 64, warp, lane,
          [&](int rr, int c, float v) { s.Lk[(rb + rr) * 64 + c] = v + __ldg(L.sk_b + n0 + c); });
    gemv8<E>(s.Bq + rb * E, E, L.sv_w + static_cast<size_t>(n0) * E, E, 64, warp, lane,
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
    gemv8<E>(s.Ba + rb * E, E, L.so_w + static_cast<size_t>(n0) * E, E, 64, warp, lane, [&](int rr, int c, float v) {
      const int row = rb + rr, n = n0 + c;
      bcast(s, s.By + row * E + n, v + __ldg(L.so_b + n) + (L.skip_pe ? 0.f : s.Bq[row * E + n]));
    });
  cluster_sync();
  DT_STAMP("self out_proj");
  layer_norm_rows(s, L.n1_w, L.n1_b, nt, warp, lane);
  __syncthreads();
  DT_STAMP("norm1");

  // ---- tokens attend to the image, norm2 (:165-170) ----
  t2i_block(s, rank, nt, nr, L.tq_w, L.tq_b, L.to_w, L.to_b, L.n2_w, L.n2_b, proj + static_cast<size_t>(r) * 4096 * ldkv, ldkv, ns, warp,
            lane);

  DT_STAMP("t2i block");
  // ---- MLP (:172-175): this CTA's quarter of the hidden units, then its partial of lin2 over that quarter ----
  const int hq = L.mlp_dim / CL;   // 512
  for (int rb = 0; rb < nr; rb += 8)
    gemv8<E>(s.Bq + rb * E, E, L.m1_w + static_cast<size_t>(hq * rank) * E, E, hq, warp, lane,
          [&](int rr, int c, float v) { s.hid[(rb + rr) * 512 + c] = fmaxf(v + __ldg(L.m1_b + hq * rank + c), 0.f); });
  __syncthreads();
  DT_STAMP("mlp lin1");
  for (int rb = 0; rb < nr; rb += 8)
    gemv8<512>(s.hid + rb * 512, 512, L.m2_w + hq * rank, L.mlp_dim, E, warp, lane,
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
    gemv8<E>(s.Bqp + rb * E, E, L.ik_w + static_cast<size_t>(32 * rank) * E, E, 32, warp, lane, [&](int rr, int c, float v) {
      if (rb + rr < nt) k128[(static_cast<size_t>(r) * nt + rb + rr) * 128 + 32 * rank + c] = v + __ldg(L.ik_b + 32 * rank + c);
    });
    gemv8<E>(s.Bq + rb * E, E, L.iv_w + static_cast<size_t>(32 * rank) * E, E, 32, warp, lane, [&](int rr, int c, float v) {
      if (rb + rr < nt) v128[(static_cast<size_t>(r) * nt + rb + rr) * 128 + 32 * rank + c] = v + __ldg(L.iv_b + 32 * rank + c);
    });
  }
  DT_STAMP("i2t k/v");
  DT_END;
  if (rank == 0) store_queries(s, queries, r, nt);
}

// ---------------------------------------------------------------------------------------------
// After the last block: final tokens->image attention + norm_final_attn (transformer.py:98-106), then — when hyper / iou
// are given — the four hypernetwork MLPs on the mask tokens (CTA c takes mask token c) and the IoU head on token 0
// (mask_decoder.py:136-148).
// ---------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(FT, 1)
dec_token_final_kernel(DecTokFinal F, float* __restrict__ queries, const float* __restrict__ tokens,
                       const __nv_bfloat16* __restrict__ proj, int ldkv, int nt, int nr, int ns, float* __restrict__ hyper,
                       float* __restrict__ iou4) {
  extern __shared__ __align__(16) float smem_f[];
  const Smem s = carve(smem_f, nr);
  const int rank = static_cast<int>(cluster_ctarank()), r = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  pdl_trigger();
  load_tokens(s, queries, tokens, r, nt, nr, true);
  t2i_block(s, rank, nt, nr, F.tq_w, F.tq_b, F.to_w, F.to_b, F.nf_w, F.nf_b, proj + static_cast<size_t>(r) * 4096 * ldkv, ldkv, ns, warp,
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

int fused_stages(int nr) { return nr <= 8 ? 6 : (nr <= 16 ? 4 : 2); }   // cp.async ring depth of the token->image attention
size_t fused_smem_bytes(int nr) {
  return (static_cast<size_t>(nr) * ROW_FLOATS + FIXED_FLOATS + fused_stages(nr) * T2I_STAGE_FLOATS) * sizeof(float);
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
  cudaError_t e = launch_pdl(dec_token_layer_kernel, dim3(CL, R), dim3(FT), smem, stream, 0, L, queries, tokens, proj, ldkv, Nt, nr,
                             fused_stages(nr), k128, v128);
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* launch_dec_token_final(const DecTokFinal& F, float* queries, const float* tokens, const __nv_bfloat16* proj, int ldkv,
                                   int R, int Nt, float* hyper, float* iou4, cudaStream_t stream) {
  if (R == 0) return nullptr;
  const int nr = (Nt + 7) / 8 * 8;
  const size_t smem = fused_smem_bytes(nr);
  static PerDeviceOnce configured;
  if (const char* e = set_max_smem_once(configured, dec_token_final_kernel, static_cast<int>(fused_smem_bytes(kDecFusedMaxTokens)))) return e;
  cudaError_t e = launch_pdl(dec_token_final_kernel, dim3(CL, R), dim3(FT), smem, stream, 0, F, queries, tokens, proj, ldkv, Nt, nr,
                             fused_stages(nr), hyper, iou4);
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace ullsam
