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
//   * 7 key tiles of 28 keys = two key-grid rows each (MMA N = 32, 4 columns masked) through TWO 32-column S slots per
//     query tile, so QK(t+1) is computed while the softmax of tile t runs; P is written back over its slot as packed
//     bf16 (16 columns); O accumulates in the upper 64 columns (which hold the Q.R^T bias tables until the gather has
//     consumed them);
//   * one softmax thread per query row (thread == TMEM lane, no cross-thread exchange); the 32 scores of a tile are
//     read once and stay in registers for the bias add, the exact row maximum and the exponentials;
//   * everything is resident in shared memory (Q 32 KB, K 28 KB, V 28 KB), every mbarrier is used exactly once (no
//     rings, no phase bookkeeping).
// CTA = 10 warps: 8 softmax (2 query tiles x 4 lane quadrants) and 2 control warps, each the MMA issuer of one query tile
// (so the two tiles' QK -> softmax -> PV chains do not wait for each other; one of them also issues the TMA loads);
// 2 CTAs / SM.
#include "attention.h"
#include "host_util.h"

#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.h"

#ifndef ULLSAM_WIN_DEFER
#define ULLSAM_WIN_DEFER 1
#endif
#ifndef ULLSAM_WIN_TIMELINE
#define ULLSAM_WIN_TIMELINE 0
#endif
#if ULLSAM_WIN_TIMELINE
#define WTL(i) do { const long long t_ = clock64(); wtl[i] += t_ - wtl_last; wtl_last = t_; } while (0)
#else
#define WTL(i) do { } while (0)
#endif

namespace ullsam {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

struct W {
  static constexpr int GRID = 14, T = 196;
  static constexpr int NT = 7, TK = 28, TN = 32;       // key tiles, keys per tile, MMA N (tile rows in shared memory)
  static constexpr int TILE_BYTES = TN * 128;          // 4 KB
  static constexpr int NREL = 32;                      // padded rows of the rel-pos tables
  static constexpr int THREADS = 320;                  // 8 softmax warps, 2 control warps (TMA + one MMA issuer each)
  static constexpr int Q1_ROWS = 72;                   // query tile 1 has 68 valid rows: only 9 swizzle atoms are loaded; the
                                                       // MMA's remaining A rows read on into the K tiles (finite, never stored)
  static constexpr int OFF_Q = 0;                      // [128 x 64] + [72 x 64] bf16
  static constexpr int OFF_K = 16384 + Q1_ROWS * 128;  // 7 x [32 x 64] bf16
  static constexpr int OFF_V = OFF_K + NT * TILE_BYTES;
  static constexpr int OFF_TABW = OFF_V + 0 * TILE_BYTES;   // the two 4 KB tables ride in V tiles 0 and 1 until the bias MMAs retire
  static constexpr int OFF_TABH = OFF_V + 1 * TILE_BYTES;
  static constexpr int OFF_SCR = OFF_V + NT * TILE_BYTES;   // fp32 [14][256]: rel_h of every row
  static constexpr int OFF_SCRW = OFF_SCR + 14 * 256 * 4;   // fp32 [14][256]: rel_w of every row
  static constexpr int OFF_BAR = OFF_SCRW + 14 * 256 * 4;
  static constexpr int NBAR = 2 + 2 * NT + 3 * 2 * NT + 2 + 2;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 256;                // per query tile: S/P slots [0,32) [32,64), O (tables first) [64,128)
};
static_assert(W::OFF_K % 1024 == 0 && W::OFF_V % 1024 == 0 && W::OFF_SCR % 16 == 0, "swizzle atoms need 1024-byte alignment");

ULLSAM_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(W::THREADS, 2)
windowed_attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQ1,
                          const __grid_constant__ CUtensorMap tmK,
                          const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmRh,
                          const __grid_constant__ CUtensorMap tmRw, int heads, int nitems, __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* pro_done = bars + 1;
  constexpr int NT = W::NT;
  uint64_t* k_full = bars + 2;         // [NT]
  uint64_t* v_full = k_full + NT;      // [NT]
  uint64_t* s_full = v_full + NT;      // [qt][NT]
  uint64_t* p_ready = s_full + 2 * NT; // [qt][NT]
  uint64_t* o_done = p_ready + 2 * NT; // [qt][NT]
  uint64_t* epi_done = o_done + 2 * NT; // [qt]: the query tile's softmax warps have drained O (the next item's tables may land)
  uint64_t* tab_done = epi_done + 2;    // [qt]: this query tile's bias-table MMAs have retired (its softmax warps may gather)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tab_done + 2);

  pdl_wait();      // (the loads issued before the CTA-wide sync already read q / k / v)
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Persistent CTA: items (window * heads + head) blockIdx.x, blockIdx.x + gridDim.x, ...  Every mbarrier completes exactly
  // one phase per item, so the wait parity of item number `it` is simply it & 1 and nothing is ever re-initialised.
  // Q, the bias tables, all K tiles and V tiles 2..4 of one item (one elected lane).  These regions are dead once the LAST
  // QK MMA of the previous item has retired (Q and K feed only the QK MMAs; V(t) is dead after PV(t); the table slots V0 / V1
  // after PV(1)), i.e. two key tiles before that item ends — so the loads of the next item are issued there and their latency
  // (~2 K cycles, which used to sit between two items with nothing else to do) is hidden under the last two softmax rounds.
  // V0, V1 (table slots) and V5, V6 (in use until the previous item's last PV) follow once the bias MMAs have retired.
  auto issue_loads = [&](int gh) {
    const int row_base = gh * W::T;
    mbar_arrive_expect_tx(q_full, 16384 + W::Q1_ROWS * 128 + 2 * 4096);
    tma_load_2d(smem + W::OFF_Q, &tmQ, q_full, 0, row_base);
    tma_load_2d(smem + W::OFF_Q + 16384, &tmQ1, q_full, 0, row_base + 128);
    tma_load_2d(smem + W::OFF_TABW, &tmRw, q_full, 0, 0);
    tma_load_2d(smem + W::OFF_TABH, &tmRh, q_full, 0, 0);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      mbar_arrive_expect_tx(&k_full[t], W::TILE_BYTES);
      tma_load_2d(smem + W::OFF_K + t * W::TILE_BYTES, &tmK, &k_full[t], 0, row_base + t * W::TK);
    }
#pragma unroll
    for (int t = 2; t < 5; ++t) {
      mbar_arrive_expect_tx(&v_full[t], W::TILE_BYTES);
      tma_load_2d(smem + W::OFF_V + t * W::TILE_BYTES, &tmV, &v_full[t], 0, row_base + t * W::TK);
    }
  };

  if (warp == 8) {
    if (elect_one_sync()) {
      mbar_init(q_full, 1);
      mbar_init(pro_done, 2);   // one commit per MMA warp
      for (int i = 0; i < NT; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); }
      for (int i = 0; i < 2 * NT; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 4); mbar_init(&o_done[i], 1); }
      mbar_init(&epi_done[0], 4);
      mbar_init(&epi_done[1], 4);
      mbar_init(&tab_done[0], 1);
      mbar_init(&tab_done[1], 1);
      fence_barrier_init();
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmQ1); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      issue_loads(blockIdx.x);
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

  if (warp == 8 || warp == 9) {
    // ================================ MMA issuers: one warp per query tile (independent chains) ========================
    const bool leader = elect_one_sync();
    const int qt = 9 - warp;   // warp 9 (TMEM owner): query tile 0; warp 8 (TMA producer): query tile 1
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, W::TN, 0, 0);
    constexpr uint32_t idesc_rel = make_idesc_bf16(128, W::NREL, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);   // A = P (TMEM), B = V (smem, MN-major)
    const uint32_t sq = smem_u32(smem + W::OFF_Q) + qt * 16384, sk = smem_u32(smem + W::OFF_K), sv = smem_u32(smem + W::OFF_V);
    const uint32_t tq0 = tmem_base + qt * 128;
    const uint64_t dq = make_smem_desc_sw128(sq, 16, 1024);
    const uint64_t drw = make_smem_desc_sw128(smem_u32(smem + W::OFF_TABW), 16, 1024);
    const uint64_t drh = make_smem_desc_sw128(smem_u32(smem + W::OFF_TABH), 16, 1024);
#pragma unroll 1
    for (int it = 0, gh = blockIdx.x; gh < nitems; ++it, gh += gridDim.x) {
      const uint32_t ph = it & 1;
      mbar_wait(q_full, ph);
      if (it > 0) mbar_wait(&epi_done[qt], ph ^ 1);   // the previous item's O has been read out: the tables may overwrite it
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tq0 + 64, dq + 2 * k, drw + 2 * k, idesc_rel, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tq0 + 96, dq + 2 * k, drh + 2 * k, idesc_rel, k != 0);
        umma_commit(&tab_done[qt]);   // to this query tile's softmax warps (the two tiles' chains stay independent)
        umma_commit(pro_done);        // to the TMA warp: both tiles' table reads have retired, V0 / V1 may land
      }
      __syncwarp();
      if (warp == 8) {
        // V tiles 0, 1 share their slots with the tables, V tiles 5, 6 were in use until the previous item's last PV: load
        // them once BOTH warps' bias MMAs have retired (each warp issued them after its epi_done, i.e. after its last PV)
        mbar_wait(pro_done, ph);
        if (leader) {
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            if (t >= 2 && t < 5) continue;
            mbar_arrive_expect_tx(&v_full[t], W::TILE_BYTES);
            tma_load_2d(smem + W::OFF_V + t * W::TILE_BYTES, &tmV, &v_full[t], 0, gh * W::T + t * W::TK);
          }
        }
        __syncwarp();
      }
      auto issue_qk = [&](int t) {   // S[t & 1] = Q K(t)^T
        mbar_wait(&k_full[t], ph);
        tc_fence_after();
        const uint64_t dk = make_smem_desc_sw128(sk + t * W::TILE_BYTES, 16, 1024);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tq0 + (t & 1) * 32, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[qt * NT + t]);
        }
        __syncwarp();
      };
      issue_qk(0);
      issue_qk(1);
#pragma unroll 1
      for (int t = 0; t < NT; ++t) {
        mbar_wait(&v_full[t], ph);
        const uint64_t dv = make_smem_desc_sw128(sv + t * W::TILE_BYTES, 16, 1024);
        mbar_wait(&p_ready[qt * NT + t], ph);   // P(t) is in its slot (and, for t = 0, the tables are consumed)
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
            umma_ts(tq0 + 64, tq0 + (t & 1) * 32 + s2 * 8, dv + s2 * 128, idesc_pv, (t | s2) != 0);
          umma_commit(&o_done[qt * NT + t]);
        }
        __syncwarp();
        // the slot of tile t is free once PV(t) has read P(t); tcgen05 ops of one thread execute in order
        if (t + 2 < NT) issue_qk(t + 2);
        if (t + 2 == NT - 1 && warp == 8 && gh + static_cast<int>(gridDim.x) < nitems) {
          // the last QK of both query tiles has retired (and with it, in issue order, every earlier QK and PV(0..4)): Q, K,
          // V2..4 and the table slots are free — the next item's loads start now
          mbar_wait(&s_full[0 * NT + NT - 1], ph);
          mbar_wait(&s_full[1 * NT + NT - 1], ph);
          if (leader) issue_loads(gh + gridDim.x);
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ softmax warps: one thread per query row ========================================
    const int qt = warp >> 2, quad = warp & 3;
    const int r = quad * 32 + lane;            // row inside the query tile == TMEM lane
    const int tq = qt * 128 + r;               // query index inside the window (>= 196: padding rows of tile 1)
    const int qh = tq / W::GRID, qw = tq % W::GRID;
    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + qt * 128;
    float* scr = reinterpret_cast<float*>(smem + W::OFF_SCR);    // [14][256] rel_h
    float* scrw = reinterpret_cast<float*>(smem + W::OFF_SCRW);  // [14][256] rel_w
    const int col = qt * 128 + r;

#if ULLSAM_WIN_TIMELINE
    long long wtl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long wtl_last = clock64();
    const long long wtl_begin = wtl_last;
#endif
#pragma unroll 1
    for (int it = 0, gh = blockIdx.x; gh < nitems; ++it, gh += gridDim.x) {
    const uint32_t ph = it & 1;
    // ---- bias rows: w[kw] = Tw[row][qw - kw + 13] (registers), h[kh] = Th[row][qh - kh + 13] (shared memory) ----
    mbar_wait(&tab_done[qt], ph);
    tc_fence_after();
    WTL(0);
    {
      // (both bias rows stay in shared memory and are re-read tile by tile: with 96 registers per thread a register copy
      //  of the 14 rel_w values pushed other state into local memory)
      uint32_t v[32], v2[32];
      tmem_ld32(tbase + 64, v);
      tmem_ld32(tbase + 96, v2);
      tmem_wait_ld();
      // (a branch-free variant — unconditional stores whose ADDRESS is selected, misses parked in a dummy word — was
      //  measured slower here: 0.426 vs 0.411 ms per launch; only 28 of the 54 entries are stored this way)
#pragma unroll
      for (int i = 0; i < 27; ++i) {
        const int kk = qw + 13 - i;
        if (kk >= 0 && kk < 14) scrw[kk * 256 + col] = __uint_as_float(v[i]) * kLog2e;
      }
#pragma unroll
      for (int i = 0; i < 27; ++i) {
        const int kk = qh + 13 - i;
        if (kk >= 0 && kk < 14) scr[kk * 256 + col] = __uint_as_float(v2[i]) * kLog2e;
      }
      // rows >= 196 (qh >= 14) have no valid window position: give them finite zeros
      if (qh >= 14) {
#pragma unroll
        for (int k = 0; k < 14; ++k) scr[k * 256 + col] = 0.f;
      }
      __syncwarp();
    }
    tc_fence_before();
    WTL(1);

    const float cs = 0.125f * kLog2e;
    float m_used = -INFINITY, l0 = 0.f, l1 = 0.f;
    const uint32_t o_addr = tbase + 64;

#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const float h0 = scr[(2 * t) * 256 + col], h1 = scr[(2 * t + 1) * 256 + col];   // key-grid rows 2t, 2t+1
      float w[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) w[k] = scrw[k * 256 + col];
      mbar_wait(&s_full[qt * NT + t], ph);
      tc_fence_after();
      WTL(2);
      const uint32_t s_addr = tbase + (t & 1) * 32;
      uint32_t v[32];
      tmem_ld32(s_addr, v);
      if (ULLSAM_WIN_DEFER && t > 0) {
        // hand P(t-1) over only now: the completion of its tcgen05.st overlaps the tensor-memory read of this tile (the
        // tensor core is a slot ahead, S(t) has long been there, so P(t-1) is not urgent; the round trip of the store is)
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[qt * NT + t - 1]);
      }
      tmem_wait_ld();
      // ---- biased scores (log2 units) of the 28 keys of the tile and their exact maximum ----
      float x[28];
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 14; c += 2) {
        x[c] = fmaf(__uint_as_float(v[c]), cs, h0) + w[c];
        x[c + 1] = fmaf(__uint_as_float(v[c + 1]), cs, h0) + w[c + 1];
        x[14 + c] = fmaf(__uint_as_float(v[14 + c]), cs, h1) + w[c];
        x[15 + c] = fmaf(__uint_as_float(v[15 + c]), cs, h1) + w[c + 1];
        mx0 = fmax3(mx0, x[c], x[c + 1]);
        mx1 = fmax3(mx1, x[14 + c], x[15 + c]);
      }
      const float mx = fmaxf(mx0, mx1);
      WTL(3);
      const bool need = mx > m_used + 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2(m_used - mx) : 1.0f;
        l0 *= alpha;
        l1 *= alpha;
        if (t > 0) {
          mbar_wait(&o_done[qt * NT + t - 1], ph);      // PV(t-1) retired: O is quiescent
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            uint32_t o[16];
            tmem_ld16(o_addr + h * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(o_addr + h * 16, o);
          }
        }
        if (need) m_used = mx;
      }
      WTL(4);
      // ---- P = 2^(x - m) for the 28 keys, 0 for the 4 padding columns; packed bf16 back over the slot ----
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < 28; c += 2) {
        const float p0 = ex2(x[c] - m_used), p1 = ex2(x[c + 1] - m_used);
        l0 += p0;
        l1 += p1;
        pk[c >> 1] = pack_bf16(p0, p1);
      }
      pk[14] = 0u;
      pk[15] = 0u;
      tmem_st16(s_addr, pk);
      WTL(5);
      if (!ULLSAM_WIN_DEFER || t == NT - 1) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[qt * NT + t]);
      }
      WTL(6);
    }

    // ---- epilogue: O / l -> out, window_unpartition + crop (image_encoder.py:267-289) ----
    mbar_wait(&o_done[qt * NT + NT - 1], ph);
    tc_fence_after();
    WTL(7);
    const float inv_l = 1.0f / (l0 + l1);
    long out_row = -1;
    if (tq < W::T) {
      const int g = gh / heads;
      const int b = g / 25, wi = g % 25;
      const int y = (wi / 5) * 14 + qh, x = (wi % 5) * 14 + qw;
      if (y < 64 && x < 64) out_row = static_cast<long>(b) * 4096 + y * 64 + x;
    }
    const int head = gh % heads;
    uint32_t o0[32], o1[32];
    tmem_ld32(o_addr, o0);
    tmem_ld32(o_addr + 32, o1);
    tmem_wait_ld();
    // O is in registers: the next item's bias tables may be written over it
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&epi_done[qt]);
    if (out_row >= 0) {
      uint4* dst = reinterpret_cast<uint4*>(out + (out_row * heads + head) * 64);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dst[u] = make_uint4(pack_bf16(__uint_as_float(o0[8 * u]) * inv_l, __uint_as_float(o0[8 * u + 1]) * inv_l),
                            pack_bf16(__uint_as_float(o0[8 * u + 2]) * inv_l, __uint_as_float(o0[8 * u + 3]) * inv_l),
                            pack_bf16(__uint_as_float(o0[8 * u + 4]) * inv_l, __uint_as_float(o0[8 * u + 5]) * inv_l),
                            pack_bf16(__uint_as_float(o0[8 * u + 6]) * inv_l, __uint_as_float(o0[8 * u + 7]) * inv_l));
        dst[4 + u] = make_uint4(pack_bf16(__uint_as_float(o1[8 * u]) * inv_l, __uint_as_float(o1[8 * u + 1]) * inv_l),
                                pack_bf16(__uint_as_float(o1[8 * u + 2]) * inv_l, __uint_as_float(o1[8 * u + 3]) * inv_l),
                                pack_bf16(__uint_as_float(o1[8 * u + 4]) * inv_l, __uint_as_float(o1[8 * u + 5]) * inv_l),
                                pack_bf16(__uint_as_float(o1[8 * u + 6]) * inv_l, __uint_as_float(o1[8 * u + 7]) * inv_l));
      }
    }
    }  // item loop
#if ULLSAM_WIN_TIMELINE
    if (blockIdx.x == 100 && lane == 0 && quad == 0) {
      const int nit = (nitems - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
      printf("win qt %d, %d items, per item: wait tables %lld | gather %lld | per tile: wait S %lld, ld+bias+max %lld, rescale %lld, exp %lld, st+arrive %lld | wait last PV + epilogue %lld | total per item %lld\n",
             qt, nit, wtl[0] / nit, wtl[1] / nit, wtl[2] / nit / NT, wtl[3] / nit / NT, wtl[4] / nit / NT, wtl[5] / nit / NT,
             wtl[6] / nit / NT, wtl[7] / nit, (clock64() - wtl_begin) / nit);
    }
#endif
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
  CUtensorMap tq, tq1, tk, tv, trh, trw;
  const char* e;
  if ((e = make_tmap_2d_bf16(&tq, q, rows, 64, 128, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tq1, q, rows, 64, W::Q1_ROWS, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tk, k, rows, 64, W::TN, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tv, v, rows, 64, W::TN, 64))) return e;
  if ((e = make_tmap_2d_bf16(&trh, relh, W::NREL, 64, W::NREL, 64))) return e;
  if ((e = make_tmap_2d_bf16(&trw, relw, W::NREL, 64, W::NREL, 64))) return e;
  static PerDeviceOnce configured;
  // two CTAs of ~100 KB need the largest shared-memory carve-out
  if ((e = set_max_smem_once(configured, windowed_attention_kernel, W::TOTAL, 100))) return e;
  const int nitems = G * heads;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = nitems < 2 * sms ? nitems : 2 * sms;   // two resident CTAs per SM, each walks items grid apart
  cudaError_t ce = launch_pdl(windowed_attention_kernel, dim3(grid), dim3(W::THREADS), W::TOTAL, stream, 1, tq, tq1, tk, tv, trh, trw,
                              heads, nitems, out);
  return ce == cudaSuccess ? nullptr : cudaGetErrorString(ce);
}

}  // namespace ullsam
