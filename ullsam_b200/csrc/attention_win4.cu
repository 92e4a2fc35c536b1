// ullsam_b200 — windowed (14 x 14) attention, third design: the WHOLE window in one shot per query tile.
//
// STATUS: EXPERIMENTAL — compiled into the library, selected only by ULLSAM_WIN_ATT=3, and NOT YET RUN ON HARDWARE (the
// round-1 GPU budget was spent before it was written; DESIGN.md section 5, plan item 1).  The default windowed path is
// attention_win.cu.  The parity tests of tests/test_ops_attention_gpu.py apply unchanged once the switch is set.
//
// Reference semantics: image_encoder.py:224-240, 325-361 on the 25 windows of the zero-padded 70 x 70 token grid
// (:243-264); T = 196 keys per window.
//
// Why a third design: attention_win.cu walks seven 28-key rounds per query tile, each a latency chain (QK -> commit ->
// tcgen05.ld -> softmax -> tcgen05.st -> arrive -> PV) of ~2.4 K cycles, plus ~5 K cycles of hand-over between items: one
// (window, head) item per ~12 K cycles per SM against a MUFU floor of ~3.1 K.  Here one CTA owns the SM and all 512
// tensor-memory columns; a query tile is a "lane" of 256 columns:
//     [0,224)   S = Q K^T for all 7 x 32 key slots at once (one N = 224 MMA per k-step); after the softmax,
//     [0,112)   P as packed bf16 (written tile by tile over score columns already consumed),
//     [128,192) O (the score columns there are dead by the time PV starts),
//     [192,256) the two Q.R^T bias tables (one N = 64 MMA) until the gather has copied them out, BEFORE QK is issued.
//   * exact row maximum (pass 1) and one exponentiation pass (pass 2), both straight from tensor memory: no online
//     rescale, no rounds, no slot ring;
//   * the 14 + 14 bias values of a row live in registers for the whole item;
//   * Q / K / V of the NEXT item arrive by TMA into the second shared-memory stage while this one is computed; the
//     rel-pos tables are loaded once per CTA;
//   * the two lanes run independently (one issuing warp each), a third control warp is the TMA producer.
// Barrier discipline: every per-item barrier completes exactly one phase per item (wait parity = item & 1) and is only
// waited on by warps that cannot fall two phases behind; the two shared-memory stages use a full / free ring (free =
// one tcgen05.commit per lane after its PV), so the producer can never observe a barrier two phases ahead.
#include "attention.h"

#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm.h"

namespace ullsam {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

struct W4 {
  static constexpr int GRID = 14, T = 196;
  static constexpr int NT = 7, TK = 28, TN = 32;        // key tiles, keys per tile, key slots per tile (4 masked)
  static constexpr int NS = NT * TN;                     // 224 score columns
  static constexpr int TILE_BYTES = TN * 128;            // 4 KB
  static constexpr int NREL = 32;                        // padded rows of each rel-pos table
  static constexpr int THREADS = 352;                    // 8 softmax warps, 2 issuing warps, 1 TMA producer warp
  static constexpr int Q1_ROWS = 72;                     // query tile 1: 68 valid rows, 9 swizzle atoms loaded; the MMA's
                                                         // remaining A rows read on into the K tiles (finite, never stored)
  static constexpr int Q_BYTES = 16384 + Q1_ROWS * 128;  // 25600
  static constexpr int OFF_K = Q_BYTES;                  // inside a stage
  static constexpr int OFF_V = OFF_K + NT * TILE_BYTES;
  static constexpr int STAGE_BYTES = OFF_V + NT * TILE_BYTES;   // 82944
  static constexpr int OFF_TAB = 2 * STAGE_BYTES;        // [Rw (32 rows) ; Rh (32 rows)] x 64 bf16 = one N = 64 operand
  static constexpr int OFF_SCRH = OFF_TAB + 2 * NREL * 128;     // fp32 [14][256]: rel_h of every row (gather scratch)
  static constexpr int OFF_SCRW = OFF_SCRH + 14 * 256 * 4;      // fp32 [14][256]: rel_w of every row
  static constexpr int OFF_BAR = OFF_SCRW + 14 * 256 * 4;
  static constexpr int NBAR = 2 + 1 + 2 + 6 * 2;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int COL_O = 128, COL_TAB = 192;
};
static_assert(W4::STAGE_BYTES % 1024 == 0 && W4::OFF_K % 1024 == 0 && W4::OFF_V % 1024 == 0 && W4::OFF_TAB % 1024 == 0,
              "swizzle atoms need 1024-byte alignment");
static_assert(W4::TOTAL <= 232448, "more than 227 KB of shared memory");

ULLSAM_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(W4::THREADS, 1)
windowed_attention_oneshot_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQ1,
                                  const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                                  const __grid_constant__ CUtensorMap tmRh, const __grid_constant__ CUtensorMap tmRw,
                                  int heads, int nitems, __nv_bfloat16* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W4::OFF_BAR);
  uint64_t* ld_full = bars + 0;        // [stage]: Q, K, V of one item have landed
  uint64_t* tab_full = bars + 2;       // the rel-pos tables have landed (once per CTA)
  uint64_t* stage_free = bars + 3;     // [stage]: both lanes' MMAs that read the stage have retired
  uint64_t* pro_done = bars + 5;       // [lane]: bias-table MMA retired
  uint64_t* gather_done = bars + 7;    // [lane]: the four softmax warps hold their bias rows in registers
  uint64_t* s_full = bars + 9;         // [lane]: S = Q K^T retired
  uint64_t* p_ready = bars + 11;       // [lane]: P is in tensor memory
  uint64_t* o_done = bars + 13;        // [lane]: O = P V retired
  uint64_t* epi_done = bars + 15;      // [lane]: O has been read out (the lane's columns are free)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + W4::NBAR);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NT = W4::NT;

  if (warp == 10) {
    if (elect_one_sync()) {
      mbar_init(&ld_full[0], 1);
      mbar_init(&ld_full[1], 1);
      mbar_init(tab_full, 1);
      mbar_init(&stage_free[0], 2);   // one tcgen05.commit per lane
      mbar_init(&stage_free[1], 2);
      for (int l = 0; l < 2; ++l) {
        mbar_init(&pro_done[l], 1);
        mbar_init(&gather_done[l], 4);
        mbar_init(&s_full[l], 1);
        mbar_init(&p_ready[l], 4);
        mbar_init(&o_done[l], 1);
        mbar_init(&epi_done[l], 4);
      }
      fence_barrier_init();
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmQ1); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    }
    __syncwarp();
  }
  if (warp == 9) {
    tmem_alloc(tmem_ptr_smem, W4::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 10) {
    // ================================ TMA producer ===================================================================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(tab_full, 2 * W4::NREL * 128);
      tma_load_2d(smem + W4::OFF_TAB, &tmRw, tab_full, 0, 0);
      tma_load_2d(smem + W4::OFF_TAB + W4::NREL * 128, &tmRh, tab_full, 0, 0);
      for (int j = 0, gh = blockIdx.x; gh < nitems; ++j, gh += gridDim.x) {
        const int st = j & 1, use = j >> 1;
        if (use > 0) mbar_wait(&stage_free[st], (use - 1) & 1);
        uint8_t* sb = smem + st * W4::STAGE_BYTES;
        const int row_base = gh * W4::T;
        mbar_arrive_expect_tx(&ld_full[st], W4::STAGE_BYTES);
        tma_load_2d(sb, &tmQ, &ld_full[st], 0, row_base);
        tma_load_2d(sb + 16384, &tmQ1, &ld_full[st], 0, row_base + 128);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          tma_load_2d(sb + W4::OFF_K + t * W4::TILE_BYTES, &tmK, &ld_full[st], 0, row_base + t * W4::TK);
          tma_load_2d(sb + W4::OFF_V + t * W4::TILE_BYTES, &tmV, &ld_full[st], 0, row_base + t * W4::TK);
        }
      }
    }
  } else if (warp == 8 || warp == 9) {
    // ================================ MMA issuers: one warp per lane (query tile) =====================================
    const bool leader = elect_one_sync();
    const int qt = 9 - warp;   // warp 9 (TMEM owner): query tile 0; warp 8: query tile 1
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, W4::NS, 0, 0);
    constexpr uint32_t idesc_tab = make_idesc_bf16(128, 2 * W4::NREL, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);   // A = P (TMEM), B = V (smem, MN-major)
    const uint32_t tq0 = tmem_base + qt * 256;
    const uint64_t dtab = make_smem_desc_sw128(smem_u32(smem + W4::OFF_TAB), 16, 1024);
#pragma unroll 1
    for (int it = 0, gh = blockIdx.x; gh < nitems; ++it, gh += gridDim.x) {
      const uint32_t ph = it & 1;
      const int st = it & 1;
      const uint32_t sb = smem_u32(smem + st * W4::STAGE_BYTES);
      const uint64_t dq = make_smem_desc_sw128(sb + qt * 16384, 16, 1024);
      const uint64_t dk = make_smem_desc_sw128(sb + W4::OFF_K, 16, 1024);   // 28 atoms of 8 key slots, 1024 bytes apart
      mbar_wait(&ld_full[st], (it >> 1) & 1);
      if (it == 0) mbar_wait(tab_full, 0);
      if (it > 0) mbar_wait(&epi_done[qt], ph ^ 1);   // the previous item's O has been read out: the lane is free
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tq0 + W4::COL_TAB, dq + 2 * k, dtab + 2 * k, idesc_tab, k != 0);
        umma_commit(&pro_done[qt]);
      }
      __syncwarp();
      mbar_wait(&gather_done[qt], ph);   // the tables have been copied out: S may overwrite columns [192,224)
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tq0, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[qt]);
      }
      __syncwarp();
      mbar_wait(&p_ready[qt], ph);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const uint64_t dv = make_smem_desc_sw128(sb + W4::OFF_V + t * W4::TILE_BYTES, 16, 1024);
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2)
            umma_ts(tq0 + W4::COL_O, tq0 + t * 16 + s2 * 8, dv + s2 * 128, idesc_pv, (t | s2) != 0);
        }
        umma_commit(&o_done[qt]);
        umma_commit(&stage_free[st]);   // this lane no longer reads the stage
      }
      __syncwarp();
    }
  } else {
    // ================================ softmax warps: one thread per query row ========================================
    const int qt = warp >> 2, quad = warp & 3;
    const int r = quad * 32 + lane;            // row inside the query tile == TMEM lane
    const int tq = qt * 128 + r;               // query index inside the window (>= 196: padding rows of tile 1)
    const int qh = tq / W4::GRID, qw = tq % W4::GRID;
    const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + qt * 256;
    float* scrh = reinterpret_cast<float*>(smem + W4::OFF_SCRH);
    float* scrw = reinterpret_cast<float*>(smem + W4::OFF_SCRW);
    const int col = qt * 128 + r;
    const float cs = 0.125f * kLog2e;

#pragma unroll 1
    for (int it = 0, gh = blockIdx.x; gh < nitems; ++it, gh += gridDim.x) {
      const uint32_t ph = it & 1;
      // ---- bias rows: w[kw] = Tw[row][qw - kw + 13], h[kh] = Th[row][qh - kh + 13], in log2 units ----
      // (the dynamic index goes through this thread's own column of the shared scratch; nobody else touches it)
      mbar_wait(&pro_done[qt], ph);
      tc_fence_after();
      float w[14], h[14];
      {
        uint32_t v[32], v2[32];
        tmem_ld32(tbase + W4::COL_TAB, v);        // Q.Rw^T
        tmem_ld32(tbase + W4::COL_TAB + 32, v2);  // Q.Rh^T
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 27; ++i) {
          const int kk = qw + 13 - i;
          if (kk >= 0 && kk < 14) scrw[kk * 256 + col] = __uint_as_float(v[i]) * kLog2e;
        }
#pragma unroll
        for (int i = 0; i < 27; ++i) {
          const int kk = qh + 13 - i;
          if (kk >= 0 && kk < 14) scrh[kk * 256 + col] = __uint_as_float(v2[i]) * kLog2e;
        }
        // rows >= 196 (qh >= 14) have no valid window position: give them finite zeros
        if (qh >= 14) {
#pragma unroll
          for (int k = 0; k < 14; ++k) scrh[k * 256 + col] = 0.f;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&gather_done[qt]);
#pragma unroll
      for (int k = 0; k < 14; ++k) {
        w[k] = scrw[k * 256 + col];
        h[k] = scrh[k * 256 + col];
      }

      mbar_wait(&s_full[qt], ph);
      tc_fence_after();
      // ---- pass 1: exact maximum of the 196 biased scores of the row (log2 units) ----
      uint32_t vb[2][32];
      float mx0 = -INFINITY, mx1 = -INFINITY;
      tmem_ld32(tbase, vb[0]);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        tmem_wait_ld();
        if (t + 1 < NT) tmem_ld32(tbase + (t + 1) * 32, vb[(t + 1) & 1]);
        const uint32_t (&v)[32] = vb[t & 1];
        const float h0 = h[2 * t], h1 = h[2 * t + 1];   // key-grid rows 2t, 2t+1
#pragma unroll
        for (int c = 0; c < 14; c += 2) {
          const float x0 = fmaf(__uint_as_float(v[c]), cs, h0) + w[c];
          const float x1 = fmaf(__uint_as_float(v[c + 1]), cs, h0) + w[c + 1];
          const float y0 = fmaf(__uint_as_float(v[14 + c]), cs, h1) + w[c];
          const float y1 = fmaf(__uint_as_float(v[15 + c]), cs, h1) + w[c + 1];
          mx0 = fmax3(mx0, x0, x1);
          mx1 = fmax3(mx1, y0, y1);
        }
      }
      const float mx = fmaxf(mx0, mx1);
      // ---- pass 2: P = 2^(x - max) for the 28 keys of each tile, 0 for its 4 padding slots; packed bf16 written over
      //      score columns that have already been read (P(t) occupies columns [16t, 16t+16), inside S(t / 2)) ----
      float l0 = 0.f, l1 = 0.f;
      tmem_ld32(tbase, vb[0]);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        tmem_wait_ld();
        if (t + 1 < NT) tmem_ld32(tbase + (t + 1) * 32, vb[(t + 1) & 1]);
        const uint32_t (&v)[32] = vb[t & 1];
        const float h0 = h[2 * t] - mx, h1 = h[2 * t + 1] - mx;
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 14; c += 2) {
          const float p0 = ex2(fmaf(__uint_as_float(v[c]), cs, h0) + w[c]);
          const float p1 = ex2(fmaf(__uint_as_float(v[c + 1]), cs, h0) + w[c + 1]);
          const float p2 = ex2(fmaf(__uint_as_float(v[14 + c]), cs, h1) + w[c]);
          const float p3 = ex2(fmaf(__uint_as_float(v[15 + c]), cs, h1) + w[c + 1]);
          l0 += p0 + p2;
          l1 += p1 + p3;
          pk[c >> 1] = pack_bf16(p0, p1);
          pk[7 + (c >> 1)] = pack_bf16(p2, p3);
        }
        pk[14] = 0u;
        pk[15] = 0u;
        tmem_st16(tbase + t * 16, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[qt]);

      // ---- epilogue: O / l -> out, window_unpartition + crop (image_encoder.py:267-289) ----
      const float inv_l = 1.0f / (l0 + l1);
      long out_row = -1;
      if (tq < W4::T) {
        const int g = gh / heads;
        const int b = g / 25, wi = g % 25;
        const int y = (wi / 5) * 14 + qh, x = (wi % 5) * 14 + qw;
        if (y < 64 && x < 64) out_row = static_cast<long>(b) * 4096 + y * 64 + x;
      }
      const int head = gh % heads;
      mbar_wait(&o_done[qt], ph);
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld32(tbase + W4::COL_O, o0);
      tmem_ld32(tbase + W4::COL_O + 32, o1);
      tmem_wait_ld();
      // O is in registers: the lane's columns may be overwritten by the next item
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, W4::TMEM_COLS);
  }
}

}  // namespace

const char* launch_attention_windowed_v3(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v,
                                         const __nv_bfloat16* relh, const __nv_bfloat16* relw, int G, int heads,
                                         __nv_bfloat16* out, cudaStream_t stream) {
  const uint64_t rows = static_cast<uint64_t>(G) * heads * W4::T;
  CUtensorMap tq, tq1, tk, tv, trh, trw;
  const char* e;
  if ((e = make_tmap_2d_bf16(&tq, q, rows, 64, 128, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tq1, q, rows, 64, W4::Q1_ROWS, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tk, k, rows, 64, W4::TN, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tv, v, rows, 64, W4::TN, 64))) return e;
  if ((e = make_tmap_2d_bf16(&trh, relh, W4::NREL, 64, W4::NREL, 64))) return e;
  if ((e = make_tmap_2d_bf16(&trw, relw, W4::NREL, 64, W4::NREL, 64))) return e;
  static bool configured = false;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(windowed_attention_oneshot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W4::TOTAL);
    if (ce != cudaSuccess) return cudaGetErrorString(ce);
    configured = true;
  }
  const int nitems = G * heads;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = nitems < sms ? nitems : sms;   // one persistent CTA per SM (it owns all 512 tensor-memory columns)
  windowed_attention_oneshot_kernel<<<grid, W4::THREADS, W4::TOTAL, stream>>>(tq, tq1, tk, tv, trh, trw, heads, nitems, out);
  cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? nullptr : cudaGetErrorString(ce);
}

}  // namespace ullsam
