// ullsam_b200 — windowed (14 x 14) attention with decomposed relative-position bias, second design (sm_100a, tcgen05).
//
// Reference semantics: image_encoder.py:224-240, 325-361 on the 25 windows of the zero-padded 70 x 70 token grid
// (:243-264); T = 196 keys per window, all of them real keys (pad rows carry k = v = qkv.bias, :168-172).
//
// Why a second design: the first one (attention.cu, GRID 14) spends 256 tensor-memory columns on one 128-query tile, so
// only two query tiles fit on an SM and each is a ~13 K-cycle latency chain (TMA -> bias MMAs -> gather -> 2 x softmax
// -> PV -> store) that keeps the MUFU and tensor pipes ~25 % / ~13 % busy.  Here a CTA owns BOTH query tiles of a
// (window, head) at 128 columns each, so FOUR query tiles are in flight per SM and K / V / the bias tables are loaded
// once per window instead of twice:
//   * 4 key tiles of 49 keys (3.5 key-grid rows; MMA N = 64, 15 columns masked) through ONE 64-column S slot per query
//     tile; P is written back over it as packed bf16 (32 columns); O accumulates in the other 64 columns (which hold
//     the Q.R^T bias tables until the gather has consumed them);
//   * one softmax thread per query row (thread == TMEM lane, no cross-thread exchange): pass A folds the raw row
//     maximum into an upper bound of the biased scores, pass B re-reads the scores in 16-column pieces, adds the bias,
//     exponentiates, and stores P (same scheme as the global kernel, attention.cu);
//   * everything is resident in shared memory (Q 32 KB, K 32 KB, V 32 KB), every mbarrier is used exactly once (no
//     rings, no phase bookkeeping): 34 barriers.
// CTA = 10 warps: 8 softmax (2 query tiles x 4 lane quadrants), 1 TMA producer, 1 MMA issuer; 2 CTAs / SM.
#include "attention.h"

#include "common.cuh"
#include "gemm.h"

namespace ullsam {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

struct W {
  static constexpr int GRID = 14, T = 196;
  static constexpr int TK = 49;                        // keys per tile (4 tiles)
  static constexpr int NREL = 32;                      // padded rows of the rel-pos tables
  static constexpr int THREADS = 320;
  static constexpr int OFF_Q = 0;                      // 2 x [128 x 64] bf16
  static constexpr int OFF_K = 32768;                  // 4 x [64 x 64] bf16
  static constexpr int OFF_V = 65536;                  // 4 x [64 x 64] bf16
  static constexpr int OFF_TABW = OFF_V + 3 * 8192;    // the two 4 KB tables ride in V tile 3's slot until the bias MMAs retire
  static constexpr int OFF_TABH = OFF_TABW + 4096;
  static constexpr int OFF_SCR = 98304;                // fp32 [14][256]: rel_h of every row (and the rel_w scatter before it)
  static constexpr int OFF_BAR = OFF_SCR + 14 * 256 * 4;
  static constexpr int NBAR = 2 + 4 + 4 + 8 + 8 + 8;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 256;                // per query tile: S/P [0,64), O (tables first) [64,128)
};

ULLSAM_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(W::THREADS, 2)
windowed_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmRh,
                          const __grid_constant__ CUtensorMap tmRw, int heads, __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* pro_done = bars + 1;
  uint64_t* k_full = bars + 2;        // [4]
  uint64_t* v_full = k_full + 4;      // [4]
  uint64_t* s_full = v_full + 4;      // [qt][4]
  uint64_t* p_ready = s_full + 8;     // [qt][4]
  uint64_t* o_done = p_ready + 8;     // [qt][4]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gh = blockIdx.x;                       // window * heads + head
  const int row_base = gh * W::T;

  if (warp == 8) {
    if (elect_one_sync()) {
      mbar_init(q_full, 1);
      mbar_init(pro_done, 1);
      for (int i = 0; i < 4; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); }
      for (int i = 0; i < 8; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_done[i], 1); }
      fence_barrier_init();
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      mbar_arrive_expect_tx(q_full, 2 * 16384 + 2 * 4096);
      tma_load_2d(smem + W::OFF_Q, &tmQ, q_full, 0, row_base);
      tma_load_2d(smem + W::OFF_Q + 16384, &tmQ, q_full, 0, row_base + 128);
      tma_load_2d(smem + W::OFF_TABW, &tmRw, q_full, 0, 0);
      tma_load_2d(smem + W::OFF_TABH, &tmRh, q_full, 0, 0);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        mbar_arrive_expect_tx(&k_full[t], 8192);
        tma_load_2d(smem + W::OFF_K + t * 8192, &tmK, &k_full[t], 0, row_base + t * W::TK);
      }
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        mbar_arrive_expect_tx(&v_full[t], 8192);
        tma_load_2d(smem + W::OFF_V + t * 8192, &tmV, &v_full[t], 0, row_base + t * W::TK);
      }
    }
    __syncwarp();
  }
  if (warp == 9) {
    tmem_alloc(tmem_ptr_smem, W::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 8) {
    // ================================ TMA producer (rest): V tile 3 once its slot no longer holds the tables ==========
    if (elect_one_sync()) {
      mbar_wait(pro_done, 0);
      mbar_arrive_expect_tx(&v_full[3], 8192);
      tma_load_2d(smem + W::OFF_V + 3 * 8192, &tmV, &v_full[3], 0, row_base + 3 * W::TK);
    }
  } else if (warp == 9) {
    // ================================ MMA issuer ======================================================================
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_rel = make_idesc_bf16(128, W::NREL, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);   // A = P (TMEM), B = V (smem, MN-major)
    const uint32_t sq = smem_u32(smem + W::OFF_Q), sk = smem_u32(smem + W::OFF_K), sv = smem_u32(smem + W::OFF_V);
    mbar_wait(q_full, 0);
    tc_fence_after();
    if (leader) {
      const uint64_t drw = make_smem_desc_sw128(smem_u32(smem + W::OFF_TABW), 16, 1024);
      const uint64_t drh = make_smem_desc_sw128(smem_u32(smem + W::OFF_TABH), 16, 1024);
#pragma unroll
      for (int qt = 0; qt < 2; ++qt) {
        const uint64_t dq = make_smem_desc_sw128(sq + qt * 16384, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_base + qt * 128 + 64, dq + 2 * k, drw + 2 * k, idesc_rel, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_base + qt * 128 + 96, dq + 2 * k, drh + 2 * k, idesc_rel, k != 0);
      }
      umma_commit(pro_done);
    }
    __syncwarp();
    auto issue_qk = [&](int t, int qt) {   // S[qt] = Q[qt] K(t)^T
      if (qt == 0) {
        mbar_wait(&k_full[t], 0);
        tc_fence_after();
      }
      const uint64_t dq = make_smem_desc_sw128(sq + qt * 16384, 16, 1024);
      const uint64_t dk = make_smem_desc_sw128(sk + t * 8192, 16, 1024);
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_base + qt * 128, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[qt * 4 + t]);
      }
      __syncwarp();
    };
    issue_qk(0, 0);
    issue_qk(0, 1);
#pragma unroll 1
    for (int t = 0; t < 4; ++t) {
#pragma unroll 1
      for (int qt = 0; qt < 2; ++qt) {
        if (qt == 0) mbar_wait(&v_full[t], 0);
        const uint64_t dv = make_smem_desc_sw128(sv + t * 8192, 16, 1024);
        mbar_wait(&p_ready[qt * 4 + t], 0);   // P(t) of this query tile is in its slot (and, for t = 0, the tables are consumed)
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int s2 = 0; s2 < 4; ++s2)
            umma_ts(tmem_base + qt * 128 + 64, tmem_base + qt * 128 + s2 * 8, dv + s2 * 128, idesc_pv, (t | s2) != 0);
          umma_commit(&o_done[qt * 4 + t]);
        }
        __syncwarp();
        if (t + 1 < 4) issue_qk(t + 1, qt);   // tcgen05 ops of one thread execute in order: S(t+1) lands after PV(t) has read P(t)
      }
    }
  } else {
    // ================================ softmax warps: one thread per query row ========================================
    const int qt = warp >> 2, quad = warp & 3;
    const int r = quad * 32 + lane;            // row inside the query tile == TMEM lane
    const int tq = qt * 128 + r;               // query index inside the window (>= 196: padding rows of tile 1)
    const int qh = tq / W::GRID, qw = tq % W::GRID;
    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + qt * 128;
    float* scr = reinterpret_cast<float*>(smem + W::OFF_SCR);   // [14][256]
    const int col = qt * 128 + r;

    // ---- bias rows: w[kw] = Tw[row][qw - kw + 13] (registers), h[kh] = Th[row][qh - kh + 13] (shared memory) ----
    mbar_wait(pro_done, 0);
    tc_fence_after();
    float w[14];
    float wmax = -INFINITY;
    {
      uint32_t v[32];
      tmem_ld32(tbase + 64, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 27; ++i) {
        const int kk = qw + 13 - i;
        if (kk >= 0 && kk < 14) scr[kk * 256 + col] = __uint_as_float(v[i]) * kLog2e;
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 14; ++k) {
        w[k] = scr[k * 256 + col];
        wmax = fmaxf(wmax, w[k]);
      }
      __syncwarp();
      tmem_ld32(tbase + 96, v);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 27; ++i) {
        const int kk = qh + 13 - i;
        if (kk >= 0 && kk < 14) scr[kk * 256 + col] = __uint_as_float(v[i]) * kLog2e;
      }
      // rows >= 196 (qh >= 14) have no valid window position: give them finite zeros
      if (qh >= 14) {
#pragma unroll
        for (int k = 0; k < 14; ++k) scr[k * 256 + col] = 0.f;
      }
      __syncwarp();
    }
    tc_fence_before();

    const float cs = 0.125f * kLog2e;
    float m_used = -INFINITY, l0 = 0.f, l1 = 0.f;
    const uint32_t o_addr = tbase + 64;

#pragma unroll
    for (int t = 0; t < 4; ++t) {
      mbar_wait(&s_full[qt * 4 + t], 0);
      tc_fence_after();
      // key c of this tile is key 49 t + c of the window: key-grid row (49 t + c) / 14, column (49 t + c) % 14
      constexpr int kFirstRow[4] = {0, 3, 7, 10};      // (49 t) / 14
      constexpr int kNumRows[4] = {4, 4, 4, 4};        // key-grid rows touched by the tile: 0-3, 3-6, 7-10, 10-13
      float hrow[4];
      float hmax = -INFINITY;
#pragma unroll
      for (int a = 0; a < kNumRows[t]; ++a) {
        hrow[a] = scr[(kFirstRow[t] + a) * 256 + col];
        hmax = fmaxf(hmax, hrow[a]);
      }
      // ---- pass A: raw maximum over the 49 valid columns ----
      float mx0 = -INFINITY, mx1 = -INFINITY;
      {
        uint32_t a0[32], a1[16];
        tmem_ld32(tbase, a0);
        tmem_ld16(tbase + 32, a1);
        const uint32_t a48 = tmem_ld1(tbase + 48);      // column 48 is the last valid one
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx0 = fmax3(mx0, __uint_as_float(a0[i]), __uint_as_float(a0[i + 1]));
#pragma unroll
        for (int i = 0; i < 16; i += 2) mx1 = fmax3(mx1, __uint_as_float(a1[i]), __uint_as_float(a1[i + 1]));
        mx1 = fmaxf(mx1, __uint_as_float(a48));
      }
      const float mx = fmaf(fmaxf(mx0, mx1), cs, wmax + hmax);
      const bool need = mx > m_used + 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2(m_used - mx) : 1.0f;
        l0 *= alpha;
        l1 *= alpha;
        if (t > 0) {
          mbar_wait(&o_done[qt * 4 + t - 1], 0);       // PV(t-1) retired: O is quiescent
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t o[32];
            tmem_ld32(o_addr + h * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(o_addr + h * 32, o);
          }
        }
        if (need) m_used = mx;
      }
      // ---- pass B: P = 2^(s * scale + rel_h + rel_w - m) for the 49 valid keys, 0 for the 15 padding columns ----
#pragma unroll
      for (int a = 0; a < kNumRows[t]; ++a) hrow[a] -= m_used;
#pragma unroll
      for (int pc = 0; pc < 4; ++pc) {
        uint32_t b[16];
        tmem_ld16(tbase + pc * 16, b);
        tmem_wait_ld();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float p[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int c = pc * 16 + i + u;
            if (c < W::TK) {
              const int key = W::TK * t + c;
              const int kh = key / W::GRID - kFirstRow[t], kw = key % W::GRID;
              p[u] = ex2(fmaf(__uint_as_float(b[i + u]), cs, hrow[kh]) + w[kw]);
            } else {
              p[u] = 0.f;
            }
          }
          l0 += p[0];
          l1 += p[1];
          pk[i >> 1] = pack_bf16(p[0], p[1]);
        }
        tmem_st8(tbase + pc * 8, pk);    // overwrites score columns [8 pc, 8 pc + 8): read by piece pc / 2 <= pc
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[qt * 4 + t]);
    }

    // ---- epilogue: O / l -> out, window_unpartition + crop (image_encoder.py:267-289) ----
    mbar_wait(&o_done[qt * 4 + 3], 0);
    tc_fence_after();
    const float inv_l = 1.0f / (l0 + l1);
    long out_row = -1;
    if (tq < W::T) {
      const int g = gh / heads;
      const int b = g / 25, wi = g % 25;
      const int y = (wi / 5) * 14 + qh, x = (wi % 5) * 14 + qw;
      if (y < 64 && x < 64) out_row = static_cast<long>(b) * 4096 + y * 64 + x;
    }
    const int head = gh % heads;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t o[32];
      tmem_ld32(o_addr + h * 32, o);
      tmem_wait_ld();
      if (out_row >= 0) {
        uint4* dst = reinterpret_cast<uint4*>(out + (out_row * heads + head) * 64 + h * 32);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          dst[u] = make_uint4(pack_bf16(__uint_as_float(o[8 * u]) * inv_l, __uint_as_float(o[8 * u + 1]) * inv_l),
                              pack_bf16(__uint_as_float(o[8 * u + 2]) * inv_l, __uint_as_float(o[8 * u + 3]) * inv_l),
                              pack_bf16(__uint_as_float(o[8 * u + 4]) * inv_l, __uint_as_float(o[8 * u + 5]) * inv_l),
                              pack_bf16(__uint_as_float(o[8 * u + 6]) * inv_l, __uint_as_float(o[8 * u + 7]) * inv_l));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, W::TMEM_COLS);
  }
}

}  // namespace

const char* launch_attention_windowed_v2(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v,
                                         const __nv_bfloat16* relh, const __nv_bfloat16* relw, int G, int heads,
                                         __nv_bfloat16* out, cudaStream_t stream) {
  const uint64_t rows = static_cast<uint64_t>(G) * heads * W::T;
  CUtensorMap tq, tk, tv, trh, trw;
  const char* e;
  if ((e = make_tmap_2d_bf16(&tq, q, rows, 64, 128, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tk, k, rows, 64, 64, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tv, v, rows, 64, 64, 64))) return e;
  if ((e = make_tmap_2d_bf16(&trh, relh, W::NREL, 64, W::NREL, 64))) return e;
  if ((e = make_tmap_2d_bf16(&trw, relw, W::NREL, 64, W::NREL, 64))) return e;
  static bool configured = false;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(windowed_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W::TOTAL);
    if (ce != cudaSuccess) return cudaGetErrorString(ce);
    configured = true;
  }
  windowed_attention_kernel<<<G * heads, W::THREADS, W::TOTAL, stream>>>(tq, tk, tv, trh, trw, heads, out);
  cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? nullptr : cudaGetErrorString(ce);
}

}  // namespace ullsam
