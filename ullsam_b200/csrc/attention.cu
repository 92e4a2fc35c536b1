// ullsam_b200 — fused attention with in-kernel decomposed relative-position bias (sm_100a, tcgen05).
//
// Reference semantics (modeling/image_encoder.py:224-240, 325-361):
//   attn = softmax( (q*scale) k^T + rel_h[q, kh] + rel_w[q, kw] ) v        rel_* use the UNSCALED q
//   rel_h[q, kh] = q . rel_pos_h[qh - kh + GRID-1],   rel_w[q, kw] = q . rel_pos_w[qw - kw + GRID-1]
// for the global 64x64 grid (T = 4096 keys) and for 14x14 windows (T = 196 keys, all of them real keys:
// the zero-padded rows of edge windows carry k = v = qkv.bias and are attended to, :168-172, 243-264).
//
// CTA = QT query tiles of 128 rows of one (group, head):
//   global   (GRID 64): QT = 2, 384 threads, one CTA per SM, 32 key tiles of 128 keys (2 grid rows)
//   windowed (GRID 14): QT = 1, 192 threads, two CTAs per SM, 2 key tiles of 112 keys (8 grid rows)
// Warp roles: QT softmax warpgroups (thread == query row == TMEM lane), one TMA producer warp, one
// tcgen05.mma issuer warp (single thread) that also owns the TMEM allocation.
//
// Flash-style loop, everything between the two MMAs stays in tensor memory:
//   S = Q K^T            SS MMA into a ring of TMEM slots (3 global / 2 windowed): the tensor core runs ahead
//   softmax warpgroup    loads its whole S row (tcgen05.ld), adds the bias, takes the exact row max,
//                        exponentiates, and stores P as packed bf16 pairs over the first half of the SAME
//                        slot (tcgen05.st) — no shared-memory round trip
//   O += P V             TS MMA: A = P from TMEM, B = V from shared memory (MN-major), O accumulates in TMEM
// The running max is only moved when it grows by more than 2^8 (exact: softmax is invariant to the reference
// point; O and the row sum are rescaled together), so touching O between tiles is rare and the softmax of
// tile j+1 never waits for the PV MMA of tile j.
//
// Bias: the prologue computes Tw = Q rel_pos_w^T and Th = Q rel_pos_h^T with two extra MMAs per query tile
// (tables zero-padded to NREL rows) and every thread scatters its own 2*GRID values into shared memory
// (global: rel_w fp32 in [kw/4][row][4] order for conflict-free 16-byte reads, rel_h as halves) or keeps them
// in registers (windowed).
#include "attention.h"

#include "common.cuh"
#include "gemm.h"

namespace ullsam {

namespace {

constexpr float kLog2e = 1.4426950408889634f;

template <int GRID>
struct AttCfg {
  static constexpr bool GLOBAL = (GRID == 64);
  static constexpr int QT = GLOBAL ? 2 : 1;               // query tiles per CTA
  static constexpr int THREADS = GLOBAL ? 384 : 192;
  static constexpr int CTRL_WARP0 = QT * 4;               // first control warp (TMA); +1 = MMA
  static constexpr int T = GRID * GRID;
  static constexpr int RPT = GLOBAL ? 2 : 8;              // key-grid rows per tile
  static constexpr int BN = RPT * GRID;                   // keys per tile: 128 / 112
  static constexpr int NT = (T + BN - 1) / BN;            // 32 / 2
  static constexpr int NITEMS = NT * QT;                  // (tile, q-tile) work items in issue order
  static constexpr int NSLOT = GLOBAL ? 3 : 2;            // S/P ring in TMEM, 128 columns per slot
  static constexpr int NREL = GLOBAL ? 128 : 32;          // padded rows of the rel-pos tables
  static constexpr int STAGES = GLOBAL ? 3 : 2;
  static constexpr int QBYTES = 128 * 128;                // one 128 x 64 bf16 tile
  static constexpr int KVSLOT = BN * 128;                 // one K (or V) tile: 16 KB / 14 KB
  static constexpr int NROWS = QT * 128;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_KV = QT * QBYTES;              // STAGES x (K slot, V slot)
  static constexpr int OFF_RELW = OFF_KV + STAGES * 2 * KVSLOT;
  //   global  : rel_w fp32 [16][NROWS][4] (64 KB) then rel_h half [64][NROWS] (32 KB)
  //   windowed: scatter scratch fp32 [14][NROWS] (7 KB, padded to 8 KB) then the two rel-pos tables (2 x 4 KB)
  static constexpr int OFF_RELH = OFF_RELW + (GLOBAL ? 16 * NROWS * 16 : 8192);
  static constexpr int OFF_BAR = OFF_RELH + (GLOBAL ? 64 * NROWS * 2 : 8192);
  static constexpr int NBAR = 3 + 2 * STAGES + NSLOT + 2 * QT;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = GLOBAL ? 512 : 256;
  // TMEM columns.  global: slots 0/128/256, O 384 + qt*64.  windowed: slots 0/128, O at 64 (S(0) is dead by then,
  // P(0) only covers columns 0..55).
  static constexpr int COL_O = GLOBAL ? 384 : 64;
  // where the rel-pos tables ride until the prologue MMAs retire
  static constexpr int OFF_RELTAB_W = GLOBAL ? OFF_KV + (STAGES - 1) * 2 * KVSLOT : OFF_RELH;
  static constexpr int OFF_RELTAB_H = GLOBAL ? OFF_RELTAB_W + KVSLOT : OFF_RELTAB_W + 4096;
};

ULLSAM_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int GRID>
__global__ void __launch_bounds__(AttCfg<GRID>::THREADS, AttCfg<GRID>::GLOBAL ? 1 : 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmRh,
                 const __grid_constant__ CUtensorMap tmRw, int heads, __nv_bfloat16* __restrict__ out) {
  using A = AttCfg<GRID>;
  constexpr int S = A::STAGES, BN = A::BN, NT = A::NT, QT = A::QT, NSLOT = A::NSLOT, NITEMS = A::NITEMS;
  constexpr int NROWS = A::NROWS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* pro_mma_done = bars + 1;
  uint64_t* pro_consumed = bars + 2;
  uint64_t* kv_full = bars + 3;
  uint64_t* kv_empty = kv_full + S;
  uint64_t* s_full = kv_empty + S;      // [NSLOT]
  uint64_t* p_ready = s_full + NSLOT;   // [QT]
  uint64_t* o_done = p_ready + QT;      // [QT]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + QT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gh = blockIdx.y;                       // group * heads + head
  const int qblock = blockIdx.x;                   // block of QT*128 queries inside the group
  const int row_base = gh * A::T;                  // first row of this (group, head) in q/k/v
  const int q_row0 = row_base + qblock * NROWS;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(pro_mma_done, 1);
    mbar_init(pro_consumed, QT * 4);
    for (int i = 0; i < S; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < NSLOT; ++i) mbar_init(&s_full[i], 1);
    for (int i = 0; i < QT; ++i) {
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == A::CTRL_WARP0 + 1) {
    tmem_alloc(tmem_ptr_smem, A::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == A::CTRL_WARP0) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      uint8_t* kv = smem + A::OFF_KV;
      mbar_arrive_expect_tx(q_full, QT * A::QBYTES + 2 * A::NREL * 128);
#pragma unroll
      for (int qt = 0; qt < QT; ++qt) tma_load_2d(smem + A::OFF_Q + qt * A::QBYTES, &tmQ, q_full, 0, q_row0 + qt * 128);
      tma_load_2d(smem + A::OFF_RELTAB_W, &tmRw, q_full, 0, 0);
      tma_load_2d(smem + A::OFF_RELTAB_H, &tmRh, q_full, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < NT; ++j) {
        if (A::GLOBAL && j == S - 1) mbar_wait(pro_mma_done, 0);  // that stage still holds the rel-pos tables
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], 2 * A::KVSLOT);
        tma_load_2d(kv + stage * 2 * A::KVSLOT, &tmK, &kv_full[stage], 0, row_base + j * BN);
        tma_load_2d(kv + stage * 2 * A::KVSLOT + A::KVSLOT, &tmV, &kv_full[stage], 0, row_base + j * BN);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == A::CTRL_WARP0 + 1) {
    // =========================================== MMA issuer =============================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, BN, 0, 0);
      constexpr uint32_t idesc_rel = make_idesc_bf16(128, A::NREL, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // A = P (TMEM, K-major), B = V (smem, MN-major)
      const uint32_t sq = smem_u32(smem + A::OFF_Q);
      const uint32_t skv = smem_u32(smem + A::OFF_KV);

      auto issue_qk = [&](int n) {  // S[slot n % NSLOT] = Q[qt] K(j)^T
        const int j = n / QT, qt = n % QT, stage = j % S;
        if (qt == 0) {
          mbar_wait(&kv_full[stage], (j / S) & 1);
          tc_fence_after();
        }
        const uint64_t dq = make_smem_desc_sw128(sq + qt * A::QBYTES, 16, 1024);
        const uint64_t dk = make_smem_desc_sw128(skv + stage * 2 * A::KVSLOT, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_base + (n % NSLOT) * 128, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[n % NSLOT]);
      };
      auto issue_pv = [&](int n) {  // O[qt] (+)= P(n) V(j), P read from the first BN/2 columns of slot n % NSLOT
        const int j = n / QT, qt = n % QT, stage = j % S;
        const uint64_t dv = make_smem_desc_sw128(skv + stage * 2 * A::KVSLOT + A::KVSLOT, 16, 1024);
#pragma unroll
        for (int s = 0; s < BN / 16; ++s) {
          // 16 keys per step: 8 TMEM columns of packed bf16 pairs; 16 V rows x 128 B = 2048 B of shared memory
          umma_ts(tmem_base + A::COL_O + qt * 64, tmem_base + (n % NSLOT) * 128 + s * 8, dv + s * 128, idesc_pv,
                  (j | s) != 0);
        }
        umma_commit(&o_done[qt]);
        if (qt == QT - 1) umma_commit(&kv_empty[stage]);
      };

      // ---- prologue: Tw[qt] = Q[qt] Rw^T, Th[qt] = Q[qt] Rh^T ----
      mbar_wait(q_full, 0);
      tc_fence_after();
      {
        const uint64_t drw = make_smem_desc_sw128(smem_u32(smem + A::OFF_RELTAB_W), 16, 1024);
        const uint64_t drh = make_smem_desc_sw128(smem_u32(smem + A::OFF_RELTAB_H), 16, 1024);
#pragma unroll
        for (int qt = 0; qt < QT; ++qt) {
          const uint64_t dq = make_smem_desc_sw128(sq + qt * A::QBYTES, 16, 1024);
          const uint32_t col_w = A::GLOBAL ? qt * 128 : 128;
          const uint32_t col_h = A::GLOBAL ? 256 + qt * 128 : 160;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + col_w, dq + 2 * k, drw + 2 * k, idesc_rel, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + col_h, dq + 2 * k, drh + 2 * k, idesc_rel, k != 0);
        }
      }
      umma_commit(pro_mma_done);
      // NOTE: issuing QK(0) here, before the softmax warps have drained the prologue tables, was measured to be
      // racy on B200 (run-to-run differences) even though the TMEM column ranges do not overlap; keep it after.
      mbar_wait(pro_consumed, 0);
      tc_fence_after();
      for (int n = 0; n < NSLOT && n < NITEMS; ++n) issue_qk(n);
      for (int n = 0; n < NITEMS; ++n) {
        mbar_wait(&p_ready[n % QT], (n / QT) & 1);  // P(n) is in its slot
        tc_fence_after();
        issue_pv(n);
        // tcgen05 ops of one thread execute in issue order: the next S may overwrite this slot once PV(n) has read it
        if (n + NSLOT < NITEMS) issue_qk(n + NSLOT);
      }
    }
  } else if (warp < A::CTRL_WARP0) {
    // =========================================== softmax warpgroups =====================================
    const int qt = warp >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;           // row inside the query tile == TMEM lane
    const int tidq = qt * 128 + r;            // row inside the CTA
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int tq = qblock * NROWS + tidq;     // query index inside the group
    const int qh = tq / GRID, qw = tq % GRID;
    float* relw_s = reinterpret_cast<float*>(smem + A::OFF_RELW);
    __half* relh_s = reinterpret_cast<__half*>(smem + A::OFF_RELH);

    float relw_reg[A::GLOBAL ? 1 : 14];
    float relh_reg[A::GLOBAL ? 1 : 14];

    // ------------------------------------ prologue: gather the bias rows ------------------------------
    mbar_wait(pro_mma_done, 0);
    tc_fence_after();
    {
      const uint32_t col_w = A::GLOBAL ? qt * 128 : 128;
      const uint32_t col_h = A::GLOBAL ? 256 + qt * 128 : 160;
      // rel_w: Tw[row][i] with i = qw - kw + GRID-1
#pragma unroll
      for (int c = 0; c < A::NREL / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + col_w + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kw = qw + (GRID - 1) - (c * 32 + i);
          if (kw >= 0 && kw < GRID) {
            if constexpr (A::GLOBAL) relw_s[((kw >> 2) * NROWS + tidq) * 4 + (kw & 3)] = __uint_as_float(v[i]) * kLog2e;
            else relw_s[kw * NROWS + tidq] = __uint_as_float(v[i]) * kLog2e;
          }
        }
      }
      if constexpr (!A::GLOBAL) {
#pragma unroll
        for (int kw = 0; kw < 14; ++kw) relw_reg[kw] = relw_s[kw * NROWS + tidq];
      } else {
        relw_reg[0] = 0.f;
      }
      // rel_h: Th[row][i] with i = qh - kh + GRID-1   (windowed: the same thread re-uses its own scratch column)
#pragma unroll
      for (int c = 0; c < A::NREL / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + col_h + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kh = qh + (GRID - 1) - (c * 32 + i);
          if (kh >= 0 && kh < GRID) {
            if constexpr (A::GLOBAL) relh_s[kh * NROWS + tidq] = __float2half_rn(__uint_as_float(v[i]) * kLog2e);
            else relw_s[kh * NROWS + tidq] = __uint_as_float(v[i]) * kLog2e;
          }
        }
      }
      if constexpr (!A::GLOBAL) {
#pragma unroll
        for (int kh = 0; kh < 14; ++kh) relh_reg[kh] = relw_s[kh * NROWS + tidq];
      } else {
        relh_reg[0] = 0.f;
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(pro_consumed);  // the prologue's TMEM columns may now be overwritten

    // ------------------------------------ main loop ---------------------------------------------------
    const float cs = 0.125f * kLog2e;  // head_dim 64: scale = 1/8, folded with log2(e)
    float m_used = -INFINITY;          // reference point of the exponentials (log2 units)
    float l = 0.f;
    const uint32_t o_addr = lane_addr + A::COL_O + qt * 64;
    constexpr int NCH = (BN + 31) / 32;  // 32-column chunks per tile (the last one is partial for BN = 112)

#pragma unroll 1
    for (int j = 0; j < NT; ++j) {
      const int n = j * QT + qt;
      const uint32_t s_addr = lane_addr + (n % NSLOT) * 128;
      mbar_wait(&s_full[n % NSLOT], (n / NSLOT) & 1);
      tc_fence_after();
      uint32_t v[NCH][32];
#pragma unroll
      for (int c = 0; c < NCH; ++c) tmem_ld32(s_addr + c * 32, v[c]);
      tmem_wait_ld();

      // ---- biased scores (log2 units, without the per-key-row term) and the exact row max ----
      float mx;
      float rh0 = 0.f, rh1 = 0.f;
      float rht[8];
      if constexpr (A::GLOBAL) {
        rh0 = __half2float(relh_s[(2 * j) * NROWS + tidq]);
        rh1 = __half2float(relh_s[(2 * j + 1) * NROWS + tidq]);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
          const float4 w = *reinterpret_cast<const float4*>(relw_s + (k4 * NROWS + tidq) * 4);
          const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int col = k4 * 4 + u;  // kw
            const float t0 = fmaf(__uint_as_float(v[col >> 5][col & 31]), cs, wv[u]);
            const float t1 = fmaf(__uint_as_float(v[(col + 64) >> 5][col & 31]), cs, wv[u]);
            v[col >> 5][col & 31] = __float_as_uint(t0);
            v[(col + 64) >> 5][col & 31] = __float_as_uint(t1);
            mx0 = fmaxf(mx0, t0);
            mx1 = fmaxf(mx1, t1);
          }
        }
        mx = fmaxf(mx0 + rh0, mx1 + rh1);
        rht[0] = 0.f;
      } else {
#pragma unroll
        for (int a = 0; a < 8; ++a) rht[a] = (j == 0) ? relh_reg[a] : ((8 + a < 14) ? relh_reg[(8 + a) % 14] : 0.f);
        const int valid_cols = (j == NT - 1) ? (A::T - (NT - 1) * BN) : BN;
        mx = -INFINITY;
#pragma unroll
        for (int col = 0; col < BN; ++col) {
          float t = fmaf(__uint_as_float(v[col >> 5][col & 31]), cs, relw_reg[col % 14] + rht[col / 14]);
          t = (col < valid_cols) ? t : -INFINITY;
          v[col >> 5][col & 31] = __float_as_uint(t);
          mx = fmaxf(mx, t);
        }
      }

      // ---- lazy move of the reference point (whole warp: tcgen05.ld/st are collective) ----
      // o_done is observed exactly once per tile, in phase order (an mbarrier parity wait may never lag two
      // phases behind): here when O has to be touched, otherwise at the end of the tile where it never stalls.
      const bool need = mx > m_used + 8.0f;
      bool o_seen = false;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2(m_used - mx) : 1.0f;  // first tile: 2^-inf = 0, and l, O are empty
        l *= alpha;
        if (j > 0) {
          mbar_wait(&o_done[qt], (j - 1) & 1);  // PV(j-1) and everything before it retired: O is quiescent
          o_seen = true;
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(o_addr + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(o_addr + c * 32, o);
          }
        }
        if (need) m_used = mx;
      }

      // ---- P = 2^(t - m), row sum, packed bf16 pairs back into the first half of the slot ----
      float ls0 = 0.f, ls1 = 0.f;
      if constexpr (A::GLOBAL) {
        const float c0 = rh0 - m_used, c1 = rh1 - m_used;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const float cc = (c < 2) ? c0 : c1;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2(__uint_as_float(v[c][i]) + cc);
            const float p1 = ex2(__uint_as_float(v[c][i + 1]) + cc);
            ls0 += p0;
            ls1 += p1;
            v[c >> 1][(c & 1) * 16 + (i >> 1)] = pack_bf16(p0, p1);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int col = c * 32 + i;
            if (col < BN) {
              const float p0 = ex2(__uint_as_float(v[c][i]) - m_used);
              const float p1 = ex2(__uint_as_float(v[c][i + 1]) - m_used);
              ls0 += p0;
              ls1 += p1;
              v[c >> 1][(c & 1) * 16 + (i >> 1)] = pack_bf16(p0, p1);
            }
          }
        }
      }
      l += ls0 + ls1;
      tmem_st32(s_addr, v[0]);
      if constexpr (A::GLOBAL) {
        tmem_st32(s_addr + 32, v[1]);
      } else {
        // BN = 112: packed columns 32..55
        tmem_st16(s_addr + 32, *reinterpret_cast<uint32_t(*)[16]>(&v[1][0]));
        tmem_st8(s_addr + 48, *reinterpret_cast<uint32_t(*)[8]>(&v[1][16]));
      }
      tmem_wait_st();
      if (j > 0 && !o_seen) mbar_wait(&o_done[qt], (j - 1) & 1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[qt]);
    }

    // ------------------------------------ epilogue: O / l -> out ---------------------------------------
    mbar_wait(&o_done[qt], (NT - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    long out_row = -1;
    if constexpr (A::GLOBAL) {
      out_row = static_cast<long>(gh / heads) * A::T + tq;
    } else {
      if (tq < A::T) {
        const int g = gh / heads;
        const int b = g / 25, w = g % 25;
        const int y = (w / 5) * 14 + qh, x = (w % 5) * 14 + qw;
        if (y < 64 && x < 64) out_row = static_cast<long>(b) * 4096 + y * 64 + x;
      }
    }
    const int head = gh % heads;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(o_addr + c * 32, o);
      tmem_wait_ld();
      if (out_row >= 0) {
        uint4* dst = reinterpret_cast<uint4*>(out + (out_row * heads + head) * 64 + c * 32);
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
  if (warp == A::CTRL_WARP0 + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, A::TMEM_COLS);
  }
}

template <int GRID>
const char* launch_t(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const __nv_bfloat16* relh,
                     const __nv_bfloat16* relw, int G, int heads, __nv_bfloat16* out, cudaStream_t stream) {
  using A = AttCfg<GRID>;
  const uint64_t rows = static_cast<uint64_t>(G) * heads * A::T;
  CUtensorMap tq, tk, tv, trh, trw;
  const char* e;
  if ((e = make_tmap_2d_bf16(&tq, q, rows, 64, 128, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tk, k, rows, 64, A::BN, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tv, v, rows, 64, A::BN, 64))) return e;
  if ((e = make_tmap_2d_bf16(&trh, relh, A::NREL, 64, A::NREL, 64))) return e;
  if ((e = make_tmap_2d_bf16(&trw, relw, A::NREL, 64, A::NREL, 64))) return e;
  static bool configured = false;
  auto kern = attention_kernel<GRID>;
  if (!configured) {
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A::TOTAL);
    if (ce != cudaSuccess) return cudaGetErrorString(ce);
    configured = true;
  }
  dim3 grid((A::T + A::NROWS - 1) / A::NROWS, G * heads);
  kern<<<grid, A::THREADS, A::TOTAL, stream>>>(tq, tk, tv, trh, trw, heads, out);
  cudaError_t ce = cudaGetLastError();
  return ce == cudaSuccess ? nullptr : cudaGetErrorString(ce);
}

}  // namespace

const char* launch_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v,
                             const __nv_bfloat16* rel_h, const __nv_bfloat16* rel_w, int G, int heads, int grid,
                             __nv_bfloat16* out, cudaStream_t stream) {
  if (G <= 0) return nullptr;
  if (grid == 64) return launch_t<64>(q, k, v, rel_h, rel_w, G, heads, out, stream);
  if (grid == 14) {
    if (G % 25 != 0) return "attention: windowed mode needs 25 windows per image";
    return launch_t<14>(q, k, v, rel_h, rel_w, G, heads, out, stream);
  }
  return "attention: grid must be 64 (global) or 14 (windowed)";
}

}  // namespace ullsam
