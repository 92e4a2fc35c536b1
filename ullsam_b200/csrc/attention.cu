// ullsam_b200 — fused attention with in-kernel decomposed relative-position bias (sm_100a, tcgen05).
//
// Reference semantics (modeling/image_encoder.py:224-240, 325-361):
//   attn = softmax( (q*scale) k^T + rel_h[q, kh] + rel_w[q, kw] ) v        rel_* use the UNSCALED q
//   rel_h[q, kh] = q . rel_pos_h[qh - kh + GRID-1],   rel_w[q, kw] = q . rel_pos_w[qw - kw + GRID-1]
// for the global 64x64 grid (T = 4096 keys) and for 14x14 windows (T = 196 keys, all of them real keys:
// the zero-padded rows of edge windows carry k = v = qkv.bias and are attended to, :168-172, 243-264).
//
// CTA = QT query tiles of 128 rows of one (group, head):
//   global   (GRID 64): QT = 2, 640 threads, one CTA per SM, 32 key tiles of 128 keys (2 grid rows)
//   windowed (GRID 14): QT = 1, 320 threads, two CTAs per SM, 2 key tiles of 112 keys (8 grid rows)
// Warp roles: 8 softmax warps per query tile (TWO threads per query row, each owning half of the S tile's
// columns, so 4-5 warps per SM sub-partition hide the MUFU / TMEM latencies with <= 96 registers and no spills),
// one TMA producer warp, one tcgen05.mma issuer warp (single thread) that also owns the TMEM allocation.
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
#include "host_util.h"

#include <stdlib.h>
// 1: global attention with ONE thread per query row (8 softmax warps, row-max bound instead of an exchange);
// 0: the two-threads-per-row layout that the windowed kernel uses.
// 1: (debug builds only, tools/build_variant.sh) print a per-phase cycle breakdown of two softmax warps of CTA (0,0)
#ifndef ULLSAM_ATT_TIMELINE
#define ULLSAM_ATT_TIMELINE 0
#endif
#if ULLSAM_ATT_TIMELINE
#define TL_STAMP(i) do { const long long t_ = clock64(); tl[i] += t_ - tl_last; tl_last = t_; } while (0)
#else
#define TL_STAMP(i) do { } while (0)
#endif
#ifndef ULLSAM_ATT_SOLO
#define ULLSAM_ATT_SOLO 0
#endif
#ifndef ULLSAM_ATT_GLOBAL_ONE_THREAD
#define ULLSAM_ATT_GLOBAL_ONE_THREAD 1
#endif

#include "common.cuh"
#include "gemm.h"

namespace ullsam {

namespace {

constexpr float kLog2e = 1.4426950408889634f;

template <int GRID>
struct AttCfg {
  static constexpr bool GLOBAL = (GRID == 64);
  static constexpr int QT = GLOBAL ? 2 : 1;               // query tiles per CTA
  static constexpr int SPLIT = (GLOBAL && ULLSAM_ATT_GLOBAL_ONE_THREAD) ? 1 : 2;   // threads per query row
  static constexpr int SM_WARPS = QT * 4 * SPLIT;         // softmax warps
  static constexpr int CTRL_WARP0 = SM_WARPS;             // first control warp (TMA); +1 = MMA
  static constexpr int THREADS = (SM_WARPS + (GLOBAL ? 4 : 2)) * 32;   // 384 (640 with SPLIT 2) / 320
  static constexpr int T = GRID * GRID;
  static constexpr int RPT = GLOBAL ? 2 : 8;              // key-grid rows per tile
  static constexpr int BN = RPT * GRID;                   // keys per tile: 128 / 112
  static constexpr int NT = (T + BN - 1) / BN;            // 32 / 2
  static constexpr int NITEMS = NT * QT;                  // (tile, q-tile) work items in issue order
  static constexpr int NSLOT = GLOBAL ? 3 : 2;            // S/P ring in TMEM, 128 columns per slot
  static constexpr int NREL = GLOBAL ? 128 : 32;          // padded rows of the rel-pos tables
  static constexpr int KSTAGES = GLOBAL ? 3 : 2;          // K is needed one tile ahead of V (the S ring runs ahead)
  static constexpr int VSTAGES = 2;
  static constexpr int QBYTES = 128 * 128;                // one 128 x 64 bf16 tile
  static constexpr int KVSLOT = BN * 128;                 // one K (or V) tile: 16 KB / 14 KB
  static constexpr int NROWS = QT * 128;
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = QT * QBYTES;               // KSTAGES K tiles
  static constexpr int OFF_V = OFF_K + KSTAGES * KVSLOT;  // VSTAGES V tiles
  static constexpr int OFF_RELW = OFF_V + VSTAGES * KVSLOT;
  //   global  : rel_w fp32 [16][NROWS][4] (64 KB) then rel_h half [64][NROWS] (32 KB)
  //   windowed: two scatter scratches fp32 [14][NROWS] (2 x 8 KB) then the two rel-pos tables (2 x 4 KB)
  static constexpr int OFF_RELH = OFF_RELW + (GLOBAL ? 16 * NROWS * 16 : 16384);
  static constexpr int OFF_XCH = OFF_RELH + (GLOBAL ? 64 * NROWS * 2 : 8192);   // row-max / row-sum exchange [2][2][NROWS] fp32
  static constexpr int OFF_BAR = OFF_XCH + 2 * 2 * NROWS * 4;
  static constexpr int NBAR = 3 + 2 * KSTAGES + 2 * VSTAGES + NSLOT + 3 * QT;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = GLOBAL ? 512 : 256;
  // TMEM columns.  global: slots 0/128/256, O 384 + qt*64.  windowed: slots 0/128, O at 64 (S(0) is dead by then,
  // P(0) only covers columns 0..55).
  static constexpr int COL_O = GLOBAL ? 384 : 64;
  // where the rel-pos tables ride until the prologue MMAs retire
  // (global: the last K slot and the last V slot; windowed: a private 8 KB region)
  static constexpr int OFF_RELTAB_W = GLOBAL ? OFF_K + (KSTAGES - 1) * KVSLOT : OFF_RELH;
  static constexpr int OFF_RELTAB_H = GLOBAL ? OFF_V + (VSTAGES - 1) * KVSLOT : OFF_RELTAB_W + 4096;
};

// size-overloaded TMEM accessors (the piece size of the softmax is a compile-time knob)
ULLSAM_DEVICE void tmem_ld_n(uint32_t a, uint32_t (&v)[16]) { tmem_ld16(a, v); }
ULLSAM_DEVICE void tmem_st_n(uint32_t a, const uint32_t (&v)[8]) { tmem_st8(a, v); }

#if ULLSAM_ATT_TIMELINE
ULLSAM_DEVICE unsigned tl_smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}
#endif

ULLSAM_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial, relative error 7.5e-5, far below the
// bf16 rounding of P): the MUFU unit issues one warp-wide ex2 per 8 cycles per sub-partition and is the bottleneck of
// the softmax, so a fraction of the exponentials is computed here instead (the FlashAttention-4 trick).
ULLSAM_DEVICE float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);                         // keep the result a normal number
  const float xr = x + 12582912.0f;              // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (xr - 12582912.0f);        // in [-0.5, 0.5]
  float p = fmaf(f, 0.0551716648f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));   // p * 2^round(x)
}
// bit e set: element e of every group of 8 consecutive keys uses ex2_poly instead of MUFU.EX2
// Register re-allocation between warpgroups (global kernel, one thread per row): the kernel is launched with 168
// registers per thread (384 threads); the control warpgroup gives most of its share back and the two softmax
// warpgroups grow, so each softmax thread can keep its 64 rel_w bias values in registers for the whole kernel.
#ifndef ULLSAM_ATT_REGS_SOFTMAX
#define ULLSAM_ATT_REGS_SOFTMAX 224
#endif
#ifndef ULLSAM_ATT_REGS_CTRL
#define ULLSAM_ATT_REGS_CTRL 56
#endif
#define ULLSAM_ATT_PIECE 16   // keys per piece of the exponentiation pass (32 measured slower: register pressure)
#ifndef ULLSAM_ATT_POLY_MASK
#define ULLSAM_ATT_POLY_MASK 0x00
#endif

template <int GRID>
__global__ void __launch_bounds__(AttCfg<GRID>::THREADS, AttCfg<GRID>::GLOBAL ? 1 : 2)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmRh,
                 const __grid_constant__ CUtensorMap tmRw, int heads, const float* __restrict__ kmax_norm,
                 __nv_bfloat16* __restrict__ out) {
  using A = AttCfg<GRID>;
  constexpr int KS = A::KSTAGES, VS = A::VSTAGES, BN = A::BN, NT = A::NT, QT = A::QT, NSLOT = A::NSLOT, NITEMS = A::NITEMS;
  constexpr int NROWS = A::NROWS;
#if ULLSAM_ATT_TIMELINE
  const long long tl_entry = clock64();
  const bool tl_cta = (blockIdx.x == 0 && blockIdx.y == 0) || (blockIdx.x == 5 && blockIdx.y == gridDim.y / 2);
#endif
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms, computed as an offset so the pointer stays in the shared window
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* pro_mma_done = bars + 1;
  uint64_t* pro_consumed = bars + 2;
  uint64_t* k_full = bars + 3;          // [KS]
  uint64_t* k_empty = k_full + KS;
  uint64_t* v_full = k_empty + KS;      // [VS]
  uint64_t* v_empty = v_full + VS;
  uint64_t* s_full = v_empty + VS;      // [NSLOT]
  uint64_t* p_ready = s_full + NSLOT;   // [QT]
  uint64_t* o_done = p_ready + QT;      // [QT][2]: PV(j) of query tile qt commits to o_done[qt*2 + (j & 1)]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 2 * QT);

  pdl_wait();      // (the first TMA loads below already read q / k / v)
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gh = blockIdx.y;                       // group * heads + head
  const int qblock = blockIdx.x;                   // block of QT*128 queries inside the group
  const int row_base = gh * A::T;                  // first row of this (group, head) in q/k/v
  const int q_row0 = row_base + qblock * NROWS;

  // Barrier setup and the first TMA loads come from the same elected lane of the producer warp, BEFORE the CTA-wide
  // sync: Q, the rel-pos tables and the K / V tiles whose ring slots are trivially free are already in flight while
  // tensor memory is being allocated (the windowed kernel is a per-CTA latency chain, every hop counts).
  constexpr int PRE_K = A::GLOBAL ? KS - 1 : (NT < KS ? NT : KS);   // global: the last slot still carries a table
  constexpr int PRE_V = A::GLOBAL ? 0 : (NT < VS ? NT : VS);        // global: V has its own producer warp
  if (warp == A::CTRL_WARP0) {
    if (elect_one_sync()) {
      mbar_init(q_full, 1);
      mbar_init(pro_mma_done, 1);
      mbar_init(pro_consumed, A::SM_WARPS);
      for (int i = 0; i < KS; ++i) {
        mbar_init(&k_full[i], 1);
        mbar_init(&k_empty[i], 1);
      }
      for (int i = 0; i < VS; ++i) {
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], 1);
      }
      for (int i = 0; i < NSLOT; ++i) mbar_init(&s_full[i], 1);
      for (int i = 0; i < QT; ++i) {
        mbar_init(&p_ready[i], 4 * A::SPLIT);
        mbar_init(&o_done[2 * i], 1);
        mbar_init(&o_done[2 * i + 1], 1);
      }
      fence_barrier_init();
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      mbar_arrive_expect_tx(q_full, QT * A::QBYTES + 2 * A::NREL * 128);
#pragma unroll
      for (int qt = 0; qt < QT; ++qt) tma_load_2d(smem + A::OFF_Q + qt * A::QBYTES, &tmQ, q_full, 0, q_row0 + qt * 128);
      tma_load_2d(smem + A::OFF_RELTAB_W, &tmRw, q_full, 0, 0);
      tma_load_2d(smem + A::OFF_RELTAB_H, &tmRh, q_full, 0, 0);
#pragma unroll
      for (int j = 0; j < PRE_K; ++j) {
        mbar_arrive_expect_tx(&k_full[j], A::KVSLOT);
        tma_load_2d(smem + A::OFF_K + j * A::KVSLOT, &tmK, &k_full[j], 0, row_base + j * BN);
      }
#pragma unroll
      for (int j = 0; j < PRE_V; ++j) {
        mbar_arrive_expect_tx(&v_full[j], A::KVSLOT);
        tma_load_2d(smem + A::OFF_V + j * A::KVSLOT, &tmV, &v_full[j], 0, row_base + j * BN);
      }
    }
    __syncwarp();
  }
  if (warp == A::CTRL_WARP0 + 1) {
    tmem_alloc(tmem_ptr_smem, A::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // Role dispatch.  Global kernel with one thread per row: warps 0..7 are the two softmax warpgroups, warps 8..11 the
  // control warpgroup; setmaxnreg is warpgroup-uniform and is the first thing each side executes.
  if (warp >= A::CTRL_WARP0) {
  if constexpr (A::SPLIT == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(ULLSAM_ATT_REGS_CTRL));

  if (warp == A::CTRL_WARP0) {
    // =========================================== TMA producers ==========================================
    // the rest of the K stream from this warp; the V stream has its own ring because V(j) is consumed a full tile
    // later than K(j) — global: a separate warp, windowed (2 tiles): everything was issued before the sync.
    // One lane chosen by elect.sync issues (see the note at the MMA issuer).
    if (elect_one_sync()) {
      int stage = PRE_K % KS;
      uint32_t phase = (PRE_K / KS) & 1;
      for (int j = PRE_K; j < NT; ++j) {
        if (A::GLOBAL && j == KS - 1) mbar_wait(pro_mma_done, 0);  // that slot still holds the rel_w table
        mbar_wait(&k_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&k_full[stage], A::KVSLOT);
        tma_load_2d(smem + A::OFF_K + stage * A::KVSLOT, &tmK, &k_full[stage], 0, row_base + j * BN);
        if (++stage == KS) { stage = 0; phase ^= 1; }
      }
    }
  }
  if (warp == (A::GLOBAL ? A::CTRL_WARP0 + 2 : A::CTRL_WARP0)) {
    if (elect_one_sync()) {
      int stage = PRE_V % VS;
      uint32_t phase = (PRE_V / VS) & 1;
      for (int j = PRE_V; j < NT; ++j) {
        if (A::GLOBAL && j == VS - 1) mbar_wait(pro_mma_done, 0);  // that slot still holds the rel_h table
        mbar_wait(&v_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&v_full[stage], A::KVSLOT);
        tma_load_2d(smem + A::OFF_V + stage * A::KVSLOT, &tmV, &v_full[stage], 0, row_base + j * BN);
        if (++stage == VS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == A::CTRL_WARP0 + 1) {
    // =========================================== MMA issuer =============================================
    // The whole warp walks the control flow and waits on the barriers; the tcgen05 instructions are issued by the
    // lane picked by elect.sync.  (With a plain `lane == 0` branch ptxas wraps EVERY tcgen05.mma in an
    // ELECT / BRA.U.ANY loop over the active lanes, ~20 instructions per MMA, which made this single warp the
    // bottleneck of the whole kernel: 24 small MMAs per key tile.)
    {
      const bool leader = elect_one_sync();
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, BN, 0, 0);
      constexpr uint32_t idesc_rel = make_idesc_bf16(128, A::NREL, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // A = P (TMEM, K-major), B = V (smem, MN-major)
      const uint32_t sq = smem_u32(smem + A::OFF_Q);
      const uint32_t sk = smem_u32(smem + A::OFF_K);
      const uint32_t sv = smem_u32(smem + A::OFF_V);

      auto issue_qk = [&](int n) {  // S[slot n % NSLOT] = Q[qt] K(j)^T
        const int j = n / QT, qt = n % QT, stage = j % KS;
        if (qt == 0) {
          mbar_wait(&k_full[stage], (j / KS) & 1);
          tc_fence_after();
        }
        const uint64_t dq = make_smem_desc_sw128(sq + qt * A::QBYTES, 16, 1024);
        const uint64_t dk = make_smem_desc_sw128(sk + stage * A::KVSLOT, 16, 1024);
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + (n % NSLOT) * 128, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[n % NSLOT]);
          if (qt == QT - 1) umma_commit(&k_empty[stage]);  // K(j) is free once both query tiles' QK have retired
        }
        __syncwarp();
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
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tmem_base + col_w, dq + 2 * k, drw + 2 * k, idesc_rel, k != 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(tmem_base + col_h, dq + 2 * k, drh + 2 * k, idesc_rel, k != 0);
          }
        }
      }
      if (leader) umma_commit(pro_mma_done);
      __syncwarp();
      // windowed: S(0) (columns 0..111) does not touch the prologue tables (columns 128..191), so QK(0) is in flight
      // while the softmax warps gather the bias; S(1) shares its columns with the tables and has to wait.  (global: the
      // tables cover every S slot.)  test_attention_is_deterministic guards this hand-off.
      if constexpr (!A::GLOBAL) issue_qk(0);
      mbar_wait(pro_consumed, 0);
      tc_fence_after();
      for (int n = A::GLOBAL ? 0 : 1; n < NSLOT && n < NITEMS; ++n) issue_qk(n);
      // Steady state, item n: O[qt] += P(n) V(j), then S(n + NSLOT) into the slot P(n) vacates (tcgen05 ops of one
      // thread execute in issue order).  The hand-over P(n) -> S(n + NSLOT) sits on the critical path of the softmax
      // warps, so everything that does not depend on P(n) — the operand barriers and all descriptor arithmetic —
      // is done BEFORE waiting for it; after the wait there is only the fence and the MMA issue itself.
#if ULLSAM_ATT_TIMELINE
      long long mma_tl[3] = {0, 0, 0};
      long long mma_last = clock64();
      int mma_p[NITEMS], mma_e[NITEMS];
#endif
      for (int n = 0; n < NITEMS; ++n) {
        const int j = n / QT, qt = n % QT, vstage = j % VS;
        const int n2 = n + NSLOT, j2 = n2 / QT, qt2 = n2 % QT, kstage = j2 % KS;
        const bool has_qk = n2 < NITEMS;
        if (qt == 0) mbar_wait(&v_full[vstage], (j / VS) & 1);
        if (has_qk && qt2 == 0) mbar_wait(&k_full[kstage], (j2 / KS) & 1);
        const uint64_t dv = make_smem_desc_sw128(sv + vstage * A::KVSLOT, 16, 1024);
        const uint64_t dq = make_smem_desc_sw128(sq + qt2 * A::QBYTES, 16, 1024);
        const uint64_t dk = make_smem_desc_sw128(sk + kstage * A::KVSLOT, 16, 1024);
        const uint32_t t_slot = tmem_base + (n % NSLOT) * 128;   // P(n) now, S(n2) next
        const uint32_t t_o = tmem_base + A::COL_O + qt * 64;
        uint64_t* bar_o = &o_done[qt * 2 + (j & 1)];
        uint64_t* bar_s = &s_full[n % NSLOT];
#if ULLSAM_ATT_TIMELINE
        const long long mt0 = clock64();
#endif
        mbar_wait(&p_ready[qt], j & 1);  // P(n) is in its slot
        tc_fence_after();
#if ULLSAM_ATT_TIMELINE
        const long long mt1 = clock64();
#endif
        if (leader) {
#pragma unroll
          for (int s2 = 0; s2 < BN / 16; ++s2) {
            // P(keys 16 s2 .. + 15) as 8 packed columns.  One thread per row: contiguous.  Two threads per row (global):
            // every 32-key quarter of the tile sits at the START of the 32 score columns its thread read it from
            // (key-grid row s2 / 4, column half (s2 / 2) & 1), so no thread ever writes columns the other one still reads.
            const uint32_t pcol = (A::GLOBAL && A::SPLIT == 2) ? (s2 >> 2) * 64 + ((s2 >> 1) & 1) * 32 + (s2 & 1) * 8 : s2 * 8;
            umma_ts(t_o, t_slot + pcol, dv + s2 * 128, idesc_pv, (j | s2) != 0);
          }
          umma_commit(bar_o);
          if (qt == QT - 1) umma_commit(&v_empty[vstage]);
          if (has_qk) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss(t_slot, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
            umma_commit(bar_s);
            if (qt2 == QT - 1) umma_commit(&k_empty[kstage]);
          }
        }
        __syncwarp();
#if ULLSAM_ATT_TIMELINE
        mma_tl[0] += mt0 - mma_last;      // operand waits + descriptor prep
        mma_tl[1] += mt1 - mt0;           // waiting for P(n)
        mma_last = clock64();
        mma_tl[2] += mma_last - mt1;      // issuing the MMAs
        mma_p[n] = static_cast<int>(mt1 - tl_entry);
        mma_e[n] = static_cast<int>(mma_last - tl_entry);
#endif
      }
#if ULLSAM_ATT_TIMELINE
      if (tl_cta && blockIdx.y != 0 && leader)
        for (int n = 16; n < 32; ++n) printf("  at %7d item %2d: mma warp saw P, issue done at %7d\n", mma_p[n], n, mma_e[n]);
      if (blockIdx.x == 0 && blockIdx.y == 0 && leader)
        printf("mma warp per item: prep+operand waits %lld | wait P %lld | issue %lld\n", mma_tl[0] / NITEMS, mma_tl[1] / NITEMS,
               mma_tl[2] / NITEMS);
#endif
    }
  }
  }  // control warpgroup
  else if (A::SPLIT == 1) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(ULLSAM_ATT_REGS_SOFTMAX));
    // =========================================== softmax warps, one thread per query row =================
    // (global attention) 4 warps per query tile; thread = row = TMEM lane.  Nothing is exchanged between threads.
    // The reference point of the exponentials only has to be an UPPER BOUND of the row's biased scores that is not
    // absurdly loose (softmax is invariant to it; P is bf16 with an 8-bit exponent, O and the row sum are fp32), so
    // instead of the exact maximum of s*scale + rel_w + rel_h the row uses
    //     max_k(s) * scale + max_kw(rel_w) + max(rel_h of the tile's two key-grid rows)
    // which needs one FMNMX3 per two scores in a first pass over the tile; the second pass re-reads the scores from
    // tensor memory in 32-column pieces (tcgen05.ld bandwidth is plentiful) so that no more than 64 of them are ever
    // live in registers.  The bound can exceed the true maximum by at most the spread of the bias terms, a few
    // units in log2 for any real relative-position table.
    if constexpr (A::SPLIT == 1) {
      const int qt = warp >> 2;
      const int quad = warp & 3;
      const int r = quad * 32 + lane;           // row inside the query tile == TMEM lane
      const int tidq = qt * 128 + r;            // row inside the CTA
      const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      const int tq = qblock * NROWS + tidq;     // query index inside the group
      const int qh = tq / GRID, qw = tq % GRID;
      float* relw_s = reinterpret_cast<float*>(smem + A::OFF_RELW);
      __half* relh_s = reinterpret_cast<__half*>(smem + A::OFF_RELH);

      // ---- prologue: this row's 64 rel_h values -> shared memory, its 64 rel_w values -> registers (log2 units) ----
      // Row (qh, qw) needs T[row][q0 + 63 - k] for k = 0..63 (q0 = qh resp. qw): a 64-wide window of its 127-entry table row,
      // reversed.  rel_h: qh is warp-uniform (a warp's 32 rows lie in one image row), so the window is read straight from
      // tensor memory at a warp-uniform column offset and stored unconditionally.  rel_w: the window start differs lane by
      // lane, which registers cannot index — each thread parks its table row in a thread-private shared scratch
      // ([word][thread]: conflict-free for any per-lane word index), 64 columns at a time, and reads it back at its own
      // offset.  No predicated scatter, no divergence (the first version spent 8.5 K of the CTA's ~98 K cycles here).
#if ULLSAM_ATT_TIMELINE
      const long long tp0 = clock64();
#endif
      // |q| for the a-priori score bound (below), computed while the table MMAs run; the key-norm load is issued first
      float qnorm_k = INFINITY;
      if (kmax_norm != nullptr) {
        const float kmaxv = __ldg(kmax_norm + gh);
        mbar_wait(q_full, 0);
        const uint8_t* qrow = smem + A::OFF_Q + qt * A::QBYTES + r * 128;   // 128-byte swizzle: 16-byte chunk c at c ^ (r & 7)
        float qq = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(qrow + ((c ^ (r & 7)) << 4));
          const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f2 = __bfloat1622float2(b2[e]);
            qq = fmaf(f2.x, f2.x, fmaf(f2.y, f2.y, qq));
          }
        }
        qnorm_k = sqrtf(qq * kmaxv) * 1.004f;   // kmaxv = max |k|^2, possibly of the fp32 values before their bf16 rounding (+0.4 %)
      }
      mbar_wait(pro_mma_done, 0);
      tc_fence_after();
#if ULLSAM_ATT_TIMELINE
      const long long tp1 = clock64();
#endif
      float wmax = -INFINITY, hmax_all = -INFINITY;
      {
        const int qh_u = __shfl_sync(0xffffffffu, qh, 0);            // warp-uniform by construction
        const uint32_t col_h = lane_addr + 256 + qt * 128 + qh_u;    // Th[row][qh + i] = rel_h[row][kh = 63 - i]
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld32(col_h + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float val = __uint_as_float(v[i]) * kLog2e;
            relh_s[(63 - (c * 32 + i)) * NROWS + tidq] = __float2half_rn(val);
            hmax_all = fmaxf(hmax_all, val);
          }
        }
      }
      float wreg[64];   // the row's 64 rel_w values stay in registers for the whole kernel (they only depend on (q, kw))
      {
        float* scratch = relw_s + tidq;                              // word w of this thread at scratch[w * NROWS]
        const uint32_t col_w = lane_addr + qt * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {                                // table columns [64h, 64h + 64)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(col_w + h * 64 + c * 32, v);
            tmem_wait_ld();
            if (h == 1 && c == 1) {                                  // last read of the prologue's tensor-memory columns
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(pro_consumed);              // they may now be overwritten: QK(0..2) start
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) scratch[(c * 32 + i) * NROWS] = __uint_as_float(v[i]) * kLog2e;
          }
#pragma unroll
          for (int kw = 0; kw < 64; ++kw) {
            const int rho = qw + 63 - kw;                            // table column of key-grid column kw
            const float t = scratch[(rho & 63) * NROWS];
            wreg[kw] = (h == 0 || rho >= 64) ? t : wreg[kw];         // h = 0 fills everything, h = 1 corrects kw < qw
          }
        }
#pragma unroll
        for (int kw = 0; kw < 64; kw += 2) wmax = fmax3(wmax, wreg[kw], wreg[kw + 1]);
      }
#if ULLSAM_ATT_TIMELINE
      const long long tp2 = clock64();
#endif

      // ---- main loop ----
      const float cs = 0.125f * kLog2e;  // head_dim 64: scale = 1/8, folded with log2(e)
      float m_used = -INFINITY;          // reference point of the exponentials (log2 units)
      float l0 = 0.f, l1 = 0.f;          // row sum (two accumulation chains)
      const uint32_t o_addr = lane_addr + A::COL_O + qt * 64;
      // A-priori bound of EVERY biased score of the row (Cauchy-Schwarz): |q| * max_k |k| * scale + max rel_w + max rel_h,
      // with max_k |k| of this (image, head) from a tiny pre-pass over K (kmax_norm, may be null).  Once the reference
      // point chosen on the first tile is within 2^100 of that bound nothing can overflow any more (P is bf16 with an
      // 8-bit exponent, O and the row sum are fp32), so the maximum pass and the rescale check are skipped for the
      // remaining 31 tiles; otherwise (huge norms) the kernel keeps checking tile by tile as before.
      const float bound_rest = kmax_norm != nullptr ? fmaf(qnorm_k, cs, wmax + hmax_all) + 0.0625f   // (+ slack for the half-rounded rel_h)
                                                    : INFINITY;
      bool skip_max = false;

#if ULLSAM_ATT_TIMELINE
      long long tl[7] = {0, 0, 0, 0, 0, 0, 0};
      long long tl_last = clock64();
      const long long tl_begin = tl_last;
      int tl_wait[NT], tl_at[NT], tl_pr[NT];
#endif
      constexpr int PIECE = ULLSAM_ATT_PIECE, NP = 128 / PIECE;
#pragma unroll 1
      for (int j = 0; j < NT; ++j) {
        const int n = j * QT + qt;
        const uint32_t s_addr = lane_addr + (n % NSLOT) * 128;
        TL_STAMP(5);
        mbar_wait(&s_full[n % NSLOT], (n / NSLOT) & 1);
        tc_fence_after();
        uint32_t pb[2][PIECE];
        const float rh0 = __half2float(relh_s[(2 * j) * NROWS + tidq]);
        const float rh1 = __half2float(relh_s[(2 * j + 1) * NROWS + tidq]);
#if ULLSAM_ATT_TIMELINE
        tl_wait[j] = static_cast<int>(clock64() - tl_last);
        tl_at[j] = static_cast<int>(clock64() - tl_entry);
#endif
        TL_STAMP(0);
#if ULLSAM_ATT_SOLO   // timing experiment only (wrong results): query tile 1 does no softmax work, so tile 0's warps have the MUFU alone
        if (qt == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_ready[qt]);
          mbar_wait(&o_done[qt * 2 + (j & 1)], (j >> 1) & 1);   // keep pace with the MMA warp (it must observe every p_ready phase)
          continue;
        }
#endif

        if (skip_max) {
          tmem_ld_n(s_addr, pb[0]);   // first piece of pass B
        } else {
        // ---- pass A: maximum of the raw scores of the row (all four tcgen05.ld in flight at once) ----
        float mxs;
        {
          uint32_t a0[32], a1[32], a2[32], a3[32];
          tmem_ld32(s_addr, a0);
          tmem_ld32(s_addr + 32, a1);
          tmem_ld32(s_addr + 64, a2);
          tmem_ld32(s_addr + 96, a3);
          tmem_wait_ld();
          tmem_ld_n(s_addr, pb[0]);   // first piece of pass B, in flight while the maximum is folded
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
          float m4 = -INFINITY, m5 = -INFINITY, m6 = -INFINITY, m7 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            m0 = fmax3(m0, __uint_as_float(a0[i]), __uint_as_float(a0[i + 1]));
            m1 = fmax3(m1, __uint_as_float(a0[i + 2]), __uint_as_float(a0[i + 3]));
            m2 = fmax3(m2, __uint_as_float(a1[i]), __uint_as_float(a1[i + 1]));
            m3 = fmax3(m3, __uint_as_float(a1[i + 2]), __uint_as_float(a1[i + 3]));
            m4 = fmax3(m4, __uint_as_float(a2[i]), __uint_as_float(a2[i + 1]));
            m5 = fmax3(m5, __uint_as_float(a2[i + 2]), __uint_as_float(a2[i + 3]));
            m6 = fmax3(m6, __uint_as_float(a3[i]), __uint_as_float(a3[i + 1]));
            m7 = fmax3(m7, __uint_as_float(a3[i + 2]), __uint_as_float(a3[i + 3]));
          }
          mxs = fmaxf(fmaxf(fmax3(m0, m1, m2), fmax3(m3, m4, m5)), fmaxf(m6, m7));
        }
        const float mx = fmaf(mxs, cs, wmax + fmaxf(rh0, rh1));

        // ---- lazy move of the reference point (whole warp: tcgen05.ld/st are collective) ----
        // PV(j) commits to o_done[qt][j & 1]; O is only touched here, after waiting for PV(j-1).  That parity wait is
        // unambiguous without observing the earlier phases: S(n) was issued after PV(n-3), tcgen05 ops of one thread
        // complete in order, so having seen s_full for this tile means PV(j-2) and PV(j-3) of this query tile have
        // retired (the barrier cannot be a phase behind), and PV(j+1) cannot be issued before this warp hands P(j+1)
        // over (it cannot be a phase ahead).
        const bool need = mx > m_used + 8.0f;
        TL_STAMP(1);
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2(m_used - mx) : 1.0f;  // first tile: 2^-inf = 0, and l, O are empty
          l0 *= alpha;
          l1 *= alpha;
          if (j > 0) {
            mbar_wait(&o_done[qt * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);  // PV(j-1) retired: O is quiescent
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t b1[32];
              tmem_ld32(o_addr + h * 32, b1);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) b1[i] = __float_as_uint(__uint_as_float(b1[i]) * alpha);
              tmem_st32(o_addr + h * 32, b1);
            }
          }
          if (need) m_used = mx;
        }

          // (the row may stop looking at maxima: see bound_rest)
          skip_max = __all_sync(0xffffffffu, bound_rest <= m_used + 100.0f);
        }
        TL_STAMP(2);
        // ---- pass B: P = 2^(s*scale + rel_h + rel_w - m), PIECE keys at a time, double-buffered TMEM reads ----
        // (small pieces keep the register pressure low enough for ptxas to issue the next tcgen05.ld early)
        const float c0 = rh0 - m_used, c1 = rh1 - m_used;
        auto piece = [&](const uint32_t (&b)[PIECE], int c) {   // keys [PIECE*c, PIECE*c + PIECE) of the tile
          const float cc = (c < NP / 2) ? c0 : c1;               // key-grid row 2j + (c >= NP/2)
          const int kw0 = (c % (NP / 2)) * PIECE;                // first key-grid column of the piece
          uint32_t pk[PIECE / 2];
#pragma unroll
          for (int k4 = 0; k4 < PIECE / 4; ++k4) {
            const int kw = kw0 + 4 * k4;
            const float x0 = fmaf(__uint_as_float(b[4 * k4]), cs, cc) + wreg[kw];
            const float x1 = fmaf(__uint_as_float(b[4 * k4 + 1]), cs, cc) + wreg[kw + 1];
            const float x2 = fmaf(__uint_as_float(b[4 * k4 + 2]), cs, cc) + wreg[kw + 2];
            const float x3 = fmaf(__uint_as_float(b[4 * k4 + 3]), cs, cc) + wreg[kw + 3];
            constexpr int M = ULLSAM_ATT_POLY_MASK;
            const int e = (k4 & 1) * 4;
            const float p0 = ((M >> (e + 0)) & 1) ? ex2_poly(x0) : ex2(x0);
            const float p1 = ((M >> (e + 1)) & 1) ? ex2_poly(x1) : ex2(x1);
            const float p2 = ((M >> (e + 2)) & 1) ? ex2_poly(x2) : ex2(x2);
            const float p3 = ((M >> (e + 3)) & 1) ? ex2_poly(x3) : ex2(x3);
            l0 += p0 + p2;
            l1 += p1 + p3;
            pk[2 * k4] = pack_bf16(p0, p1);
            pk[2 * k4 + 1] = pack_bf16(p2, p3);
          }
          tmem_st_n(s_addr + c * (PIECE / 2), pk);   // overwrites score columns already consumed
        };
        auto load_piece = [&](uint32_t (&b)[PIECE], int c) {
          tmem_ld_n(s_addr + c * PIECE, b);
        };
#pragma unroll
        for (int c = 0; c < NP; ++c) {
          tmem_wait_ld();
          if (c + 1 < NP) load_piece(pb[(c + 1) & 1], c + 1);
          piece(pb[c & 1], c);
        }
        TL_STAMP(3);
        tmem_wait_st();
        // Back-pressure on the hand-over: P(j) must not be announced before the MMA warp has OBSERVED P(j-1) of this query
        // tile.  Nothing else enforces that: S(j+1) of this query tile goes out with the PV of the OTHER query tile's previous
        // item (S(n+3) is issued in iteration n), i.e. before the MMA warp looks at p_ready for P(j-1), and that warp waits
        // for its K / V operands first.  If a K tile arrives later than one softmax tile takes (~1.5 us), this warp announces
        // P(j-1) and P(j) back to back, p_ready[qt] is two phases ahead of the parity the MMA warp then waits for, and that
        // wait never returns.  Seen once, in bench.py's end-to-end loop (host-to-device copies competing for HBM): qt 1 had
        // announced P(1) and P(3) while the MMA warp was still waiting for K tile 2 (watchdog dump:
        // profiles/r2_attention_handover_race.txt).  PV(j-1) is committed by the iteration that observed P(j-1), so its
        // o_done phase is the proof; in the normal case it retired long ago (one successful try_wait).  The parity is
        // unambiguous for the same reason as in the rescale above: having seen S(j), PV(j-2) of this query tile has
        // retired, and PV(j+1) cannot exist before this warp announces P(j+1).
        if (j > 0) mbar_wait(&o_done[qt * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[qt]);
#if ULLSAM_ATT_TIMELINE
        tl_pr[j] = static_cast<int>(clock64() - tl_entry);
#endif
        TL_STAMP(4);
#if ULLSAM_ATT_TIMELINE == 2
        mbar_wait(&o_done[qt * 2 + (j & 1)], (j >> 1) & 1);  // probe: P(n) handed over -> PV(n) retired
        TL_STAMP(6);
#elif ULLSAM_ATT_TIMELINE == 3
        if (n + NSLOT < NITEMS) mbar_wait(&s_full[n % NSLOT], ((n + NSLOT) / NSLOT) & 1);  // probe: -> S(n + NSLOT) ready
        TL_STAMP(6);
#endif
      }
#if ULLSAM_ATT_TIMELINE
      if (tl_cta && lane == 0 && quad == 0)
        printf("qt %d: per tile: wait S %lld | pass A %lld | rescale %lld | pass B %lld | tail %lld | loop %lld | PV probe %lld | total/tile %lld\n",
               qt, tl[0] / NT, tl[1] / NT, tl[2] / NT, tl[3] / NT, tl[4] / NT, tl[5] / NT, tl[6] / NT, (clock64() - tl_begin) / NT);
      if (tl_cta && lane == 0 && quad == 0 && blockIdx.y != 0)
        for (int j = 8; j < 16; ++j)
          printf("  at %7d item %2d: qt %d tile %2d starts (waited %4d), P handed over at %7d\n", tl_at[j], 2 * j + qt, qt, j,
                 tl_wait[j], tl_pr[j]);
#endif

      // ---- epilogue: O / l -> out ----
      mbar_wait(&o_done[qt * 2 + ((NT - 1) & 1)], ((NT - 1) >> 1) & 1);
      tc_fence_after();
      const float inv_l = 1.0f / (l0 + l1);
      const long out_row = static_cast<long>(gh / heads) * A::T + tq;
      const int head = gh % heads;
      uint4* dst = reinterpret_cast<uint4*>(out + (out_row * heads + head) * 64);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t o[32];
        tmem_ld32(o_addr + h * 32, o);
        tmem_wait_ld();
#pragma unroll
        for (int u = 0; u < 4; ++u)
          dst[h * 4 + u] = make_uint4(pack_bf16(__uint_as_float(o[8 * u]) * inv_l, __uint_as_float(o[8 * u + 1]) * inv_l),
                                      pack_bf16(__uint_as_float(o[8 * u + 2]) * inv_l, __uint_as_float(o[8 * u + 3]) * inv_l),
                                      pack_bf16(__uint_as_float(o[8 * u + 4]) * inv_l, __uint_as_float(o[8 * u + 5]) * inv_l),
                                      pack_bf16(__uint_as_float(o[8 * u + 6]) * inv_l, __uint_as_float(o[8 * u + 7]) * inv_l));
      }
#if ULLSAM_ATT_TIMELINE
      if (tl_cta && lane == 0 && quad == 0)
        printf("cta (%d,%d) qt %d: entry -> role start %lld | -> tables computed %lld | -> scattered %lld | -> loop begin %lld\n", blockIdx.x, blockIdx.y, qt,
               tp0 - tl_entry, tp1 - tp0, tp2 - tp1, tl_begin - tp2);
      if (tl_cta && lane == 0 && quad == 0)
        printf("cta (%d,%d) qt %d: entry -> loop begin %lld | loop %lld | loop end -> stored %lld | smid %d\n", blockIdx.x, blockIdx.y, qt,
               tl_begin - tl_entry, tl_last - tl_begin, clock64() - tl_last, (int)tl_smid());
#endif
    }
  } else if (A::SPLIT == 2) {
    // =========================================== softmax warps, two threads per query row ================
    // Two threads per query row: thread (row, hs) owns columns [hs*HALF, hs*HALF + HALF) of every S tile, i.e. one
    // key-grid row per tile for the global grid.  The pair lives in two warps of the same TMEM lane quadrant and
    // exchanges its partial row max (every tile) and row sum (once, at the end) through shared memory.
    const int qt = warp >> 3;
    const int hs = (warp >> 2) & 1;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;           // row inside the query tile == TMEM lane
    const int tidq = qt * 128 + r;            // row inside the CTA
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int tq = qblock * NROWS + tidq;     // query index inside the group
    const int qh = tq / GRID, qw = tq % GRID;
    constexpr int HALF = BN / 2;              // 64 / 56 columns per thread
    float* relw_s = reinterpret_cast<float*>(smem + A::OFF_RELW);
    __half* relh_s = reinterpret_cast<__half*>(smem + A::OFF_RELH);
    float* xch = reinterpret_cast<float*>(smem + A::OFF_XCH);  // [parity][hs][NROWS]
    const int bar_id = 1 + qt;                // named barrier of this query tile's 8 warps

    // ------------------------------------ prologue: gather the bias rows ------------------------------
    // hs == 0 scatters rel_w (Tw[row][qw - kw + GRID-1]), hs == 1 scatters rel_h (Th[row][qh - kh + GRID-1]).
    mbar_wait(pro_mma_done, 0);
    tc_fence_after();
    {
      const uint32_t col_w = A::GLOBAL ? qt * 128 : 128;
      const uint32_t col_h = A::GLOBAL ? 256 + qt * 128 : 160;
      const uint32_t col_t = hs ? col_h : col_w;
      const int q0 = hs ? qh : qw;
#pragma unroll
      for (int c = 0; c < A::NREL / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + col_t + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kk = q0 + (GRID - 1) - (c * 32 + i);
          if (kk >= 0 && kk < GRID) {
            const float val = __uint_as_float(v[i]) * kLog2e;
            if constexpr (A::GLOBAL) {
              if (hs) relh_s[kk * NROWS + tidq] = __float2half_rn(val);
              else relw_s[((kk >> 2) * NROWS + tidq) * 4 + (kk & 3)] = val;
            } else {
              relw_s[(hs * 14 + kk) * NROWS + tidq] = val;
            }
          }
        }
      }
    }
    tc_fence_before();
    asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
    // (windowed: the 2 x 14 bias values of the row stay in the scatter scratch, [14 + 14][NROWS] fp32, and are read
    //  from there tile by tile — with 96 registers per thread a register copy ends up in local memory)
    if (lane == 0) mbar_arrive(pro_consumed);  // the prologue's TMEM columns may now be overwritten

    // ------------------------------------ main loop ---------------------------------------------------
    const float cs = 0.125f * kLog2e;  // head_dim 64: scale = 1/8, folded with log2(e)
    float m_used = -INFINITY;          // reference point of the exponentials (log2 units), identical in the pair
    float l = 0.f;                     // partial row sum (this thread's columns)
    const uint32_t o_addr = lane_addr + A::COL_O + qt * 64 + hs * 32;  // this thread's 32 of the 64 O columns
    // Global grid, fast mode (the a-priori bound of the one-thread-per-row path): once the reference point chosen on the
    // first tile is within 2^100 of an upper bound of EVERY biased score of the row, nothing can overflow any more, the
    // reference point never moves again and the two threads of a row have nothing left to exchange: each exponentiates
    // its 2 x 32 scores of a tile (its 32 rel_w values live in registers) and hands its half of P over.  Four softmax
    // warps per sub-partition instead of two, at the same instruction count per element.
    float bound_rest = INFINITY;
    bool fast = false;
    if constexpr (A::GLOBAL) {
      float wmax = -INFINITY, hmax = -INFINITY;
#pragma unroll
      for (int k4 = 0; k4 < 16; ++k4) {
        const float4 w = *reinterpret_cast<const float4*>(relw_s + (k4 * NROWS + tidq) * 4);
        wmax = fmaxf(fmaxf(wmax, fmaxf(w.x, w.y)), fmaxf(w.z, w.w));
      }
#pragma unroll 8
      for (int kh = 0; kh < 64; ++kh) hmax = fmaxf(hmax, __half2float(relh_s[kh * NROWS + tidq]));
      if (kmax_norm != nullptr) {
        const uint8_t* qrow = smem + A::OFF_Q + qt * A::QBYTES + r * 128;   // 128-byte swizzle: 16-byte chunk c at c ^ (r & 7)
        float qq = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = *reinterpret_cast<const uint4*>(qrow + ((c ^ (r & 7)) << 4));
          const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f2 = __bfloat1622float2(b2[e]);
            qq = fmaf(f2.x, f2.x, fmaf(f2.y, f2.y, qq));
          }
        }
        bound_rest = fmaf(sqrtf(qq * __ldg(kmax_norm + gh)) * 1.004f, cs, wmax + hmax) + 0.0625f;
      }
    }

    int j = 0;
#pragma unroll 1
    for (; j < NT && !fast; ++j) {
      const int n = j * QT + qt;
      const uint32_t s_addr = lane_addr + (n % NSLOT) * 128;
      mbar_wait(&s_full[n % NSLOT], (n / NSLOT) & 1);
      tc_fence_after();
      // global: thread hs owns key-grid columns [hs*32, hs*32+32) of BOTH key-grid rows of the tile (S columns
      // hs*32.. and 64+hs*32..), so it needs only 32 of the 64 rel_w values; windowed: columns [hs*56, hs*56+56).
      uint32_t v[2][32];
      if constexpr (A::GLOBAL) {
        tmem_ld32(s_addr + hs * 32, v[0]);
        tmem_ld32(s_addr + 64 + hs * 32, v[1]);
      } else {
        tmem_ld32(s_addr + hs * HALF, v[0]);
        tmem_ld32(s_addr + hs * HALF + 32, v[1]);  // only the first HALF - 32 = 24 columns are used
      }
      tmem_wait_ld();

      // ---- biased scores (log2 units) and this thread's partial row max ----
      float mxp = -INFINITY;
      float rh0 = 0.f, rh1 = 0.f;
      if constexpr (A::GLOBAL) {
        rh0 = __half2float(relh_s[(2 * j) * NROWS + tidq]);
        rh1 = __half2float(relh_s[(2 * j + 1) * NROWS + tidq]);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 w = *reinterpret_cast<const float4*>(relw_s + ((hs * 8 + k4) * NROWS + tidq) * 4);
          const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = k4 * 4 + u;  // key-grid column hs*32 + i
            const float t0 = fmaf(__uint_as_float(v[0][i]), cs, wv[u]);
            const float t1 = fmaf(__uint_as_float(v[1][i]), cs, wv[u]);
            v[0][i] = __float_as_uint(t0);
            v[1][i] = __float_as_uint(t1);
            mx0 = fmaxf(mx0, t0);
            mx1 = fmaxf(mx1, t1);
          }
        }
        mxp = fmaxf(mx0 + rh0, mx1 + rh1);
      } else {
        // columns hs*56 + i: key-grid row j*8 + hs*4 + i/14, key-grid column i % 14 (56 = 4 * 14)
        float rht[4], relw_reg[14];
#pragma unroll
        for (int a4 = 0; a4 < 4; ++a4) {
          const int kh = j * 8 + hs * 4 + a4;   // (rows 14, 15 of the last tile are padding: masked below)
          rht[a4] = relw_s[(14 + (kh < 14 ? kh : 13)) * NROWS + tidq];
        }
#pragma unroll
        for (int k = 0; k < 14; ++k) relw_reg[k] = relw_s[k * NROWS + tidq];
        const int valid = ((j == NT - 1) ? (A::T - (NT - 1) * BN) : BN) - hs * HALF;  // valid columns of this half
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
          float t = fmaf(__uint_as_float(v[i >> 5][i & 31]), cs, relw_reg[i % 14] + rht[i / 14]);
          t = (i < valid) ? t : -INFINITY;
          v[i >> 5][i & 31] = __float_as_uint(t);
          mxp = fmaxf(mxp, t);
        }
      }
      // ---- exchange with the other half of the row ----
      xch[((j & 1) * 2 + hs) * NROWS + tidq] = mxp;
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      const float mx = fmaxf(mxp, xch[((j & 1) * 2 + (hs ^ 1)) * NROWS + tidq]);

      // ---- lazy move of the reference point (whole warp: tcgen05.ld/st are collective) ----
      // PV(j) commits to o_done[qt][j & 1].  An mbarrier parity wait may never lag two phases behind, so every
      // phase of both barriers is observed in order: routinely two tiles late (end of this loop body, where it never
      // stalls), and exactly on time only when O really has to be touched here.
      const bool need = mx > m_used + 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2(m_used - mx) : 1.0f;  // first tile: 2^-inf = 0, and l, O are empty
        l *= alpha;
        if (j > 0) {
          mbar_wait(&o_done[qt * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);  // PV(j-1) retired: O is quiescent
          tc_fence_after();
          uint32_t o[32];
          tmem_ld32(o_addr, o);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(o_addr, o);
        }
        if (need) m_used = mx;
      }

      // ---- P = 2^(t - m), partial row sum, packed bf16 pairs back into the first half of the slot ----
      float ls0 = 0.f, ls1 = 0.f;
      if constexpr (A::GLOBAL) {
        const float c0 = rh0 - m_used, c1 = rh1 - m_used;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2(__uint_as_float(v[0][i]) + c0);
          const float p1 = ex2(__uint_as_float(v[0][i + 1]) + c0);
          const float p2 = ex2(__uint_as_float(v[1][i]) + c1);
          const float p3 = ex2(__uint_as_float(v[1][i + 1]) + c1);
          ls0 += p0 + p2;
          ls1 += p1 + p3;
          v[0][i >> 1] = pack_bf16(p0, p1);  // element i>>1 <= i has already been consumed
          v[1][i >> 1] = pack_bf16(p2, p3);
        }
        l += ls0 + ls1;
        // keys hs*32.. -> 16 packed columns at hs*32; keys 64+hs*32.. -> 16 packed columns at 64+hs*32 (the layout the MMA
        // warp reads, see the PV issue; here every score of the row was read before the exchange barrier above)
        tmem_st16(s_addr + hs * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[0][0]));
        tmem_st16(s_addr + 64 + hs * 32, *reinterpret_cast<uint32_t(*)[16]>(&v[1][0]));
        // fast mode from the next tile on?  The answer has to be the same in all eight warps of the query tile (the slow
        // path meets at a named barrier every tile): a barrier with an AND reduction.
        {
          const int ok = bound_rest <= m_used + 100.0f;
          int all_ok;
          asm volatile(
              "{\n\t"
              ".reg .pred p, q;\n\t"
              "setp.ne.b32 q, %2, 0;\n\t"
              "barrier.cta.red.and.pred p, %1, 256, q;\n\t"
              "selp.b32 %0, 1, 0, p;\n\t"
              "}\n"
              : "=r"(all_ok)
              : "r"(bar_id), "r"(ok)
              : "memory");
          fast = all_ok != 0;
        }
      } else {
#pragma unroll
        for (int i = 0; i < HALF; i += 2) {
          const float p0 = ex2(__uint_as_float(v[i >> 5][i & 31]) - m_used);
          const float p1 = ex2(__uint_as_float(v[(i + 1) >> 5][(i + 1) & 31]) - m_used);
          ls0 += p0;
          ls1 += p1;
          v[0][i >> 1] = pack_bf16(p0, p1);  // element i>>1 <= i has already been consumed
        }
        l += ls0 + ls1;
        // 28 packed columns per half
        tmem_st16(s_addr + hs * 28, *reinterpret_cast<uint32_t(*)[16]>(&v[0][0]));
        tmem_st8(s_addr + hs * 28 + 16, *reinterpret_cast<uint32_t(*)[8]>(&v[0][16]));
        tmem_st4(s_addr + hs * 28 + 24, *reinterpret_cast<uint32_t(*)[4]>(&v[0][24]));
      }
      tmem_wait_st();
      // (hand-over back-pressure, see the one-thread-per-row path: P(j) is not announced before the MMA warp has observed
      //  P(j-1); waiting for PV(j-2) only, as this path did, still lets p_ready run two phases ahead)
      if (j > 0) mbar_wait(&o_done[qt * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[qt]);
    }

    if constexpr (A::GLOBAL) {
      // ---- fast mode: the rest of the tiles (a separate loop, so the 32 rel_w registers are not live in the slow one) ----
      float wreg[32];
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4) {
        const float4 w = *reinterpret_cast<const float4*>(relw_s + ((hs * 8 + k4) * NROWS + tidq) * 4);
        wreg[k4 * 4] = w.x; wreg[k4 * 4 + 1] = w.y; wreg[k4 * 4 + 2] = w.z; wreg[k4 * 4 + 3] = w.w;
      }
#pragma unroll 1
      for (; j < NT; ++j) {
        const int n = j * QT + qt;
        const uint32_t s_addr = lane_addr + (n % NSLOT) * 128;
        mbar_wait(&s_full[n % NSLOT], (n / NSLOT) & 1);
        tc_fence_after();
        // pieces of 16 scores: piece p = key-grid row p / 2, columns hs * 32 + (p & 1) * 16 .. + 15; its 8 packed P columns
        // go to the start of the same 32-column block (already consumed: the block's two pieces are read in order)
        const float c0 = __half2float(relh_s[(2 * j) * NROWS + tidq]) - m_used;
        const float c1 = __half2float(relh_s[(2 * j + 1) * NROWS + tidq]) - m_used;
        uint32_t pb[2][16];
        float la = 0.f, lb = 0.f;
        tmem_ld16(s_addr + hs * 32, pb[0]);
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {
          const uint32_t blk = s_addr + (pc >> 1) * 64 + hs * 32;
          tmem_wait_ld();
          if (pc + 1 < 4) tmem_ld16(s_addr + ((pc + 1) >> 1) * 64 + hs * 32 + ((pc + 1) & 1) * 16, pb[(pc + 1) & 1]);
          const float cc = (pc >> 1) ? c1 : c0;
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float x0 = fmaf(__uint_as_float(pb[pc & 1][i]), cs, cc) + wreg[(pc & 1) * 16 + i];
            const float x1 = fmaf(__uint_as_float(pb[pc & 1][i + 1]), cs, cc) + wreg[(pc & 1) * 16 + i + 1];
            const float p0 = ex2(x0), p1 = ex2(x1);
            la += p0;
            lb += p1;
            pk[i >> 1] = pack_bf16(p0, p1);
          }
          tmem_st8(blk + (pc & 1) * 8, pk);
        }
        l += la + lb;
        tmem_wait_st();
        // hand-over back-pressure (see the one-thread-per-row path): P(j) only after the MMA warp has observed P(j-1)
        if (j > 0) mbar_wait(&o_done[qt * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[qt]);
      }
    }

    // ------------------------------------ epilogue: O / l -> out ---------------------------------------
    xch[((NT & 1) * 2 + hs) * NROWS + tidq] = l;
    asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
    l += xch[((NT & 1) * 2 + (hs ^ 1)) * NROWS + tidq];
    mbar_wait(&o_done[qt * 2 + ((NT - 1) & 1)], ((NT - 1) >> 1) & 1);
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
    {
      uint32_t o[32];
      tmem_ld32(o_addr, o);
      tmem_wait_ld();
      if (out_row >= 0) {
        uint4* dst = reinterpret_cast<uint4*>(out + (out_row * heads + head) * 64 + hs * 32);
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

// max_t |k[gh][t][:]| per (group, head): the pre-pass behind the global kernel's a-priori score bound.
// k: bf16 [GH][T][64].  One CTA per (group, head); 128 bytes per row, 16-byte loads.
__global__ void __launch_bounds__(256) k_norm_max_kernel(const __nv_bfloat16* __restrict__ k, int T, float* __restrict__ out) {
  __shared__ float red[8];
  pdl_wait();
  pdl_trigger();
  const uint4* base = reinterpret_cast<const uint4*>(k + static_cast<size_t>(blockIdx.x) * T * 64);
  // 8 lanes per key row (16 bytes each): every load instruction of a warp covers four whole 128-byte rows; the squared norm
  // of a row is the sum over its 8 lanes.  Four independent loads per lane in flight.
  const int sub = threadIdx.x & 7, rgrp = threadIdx.x >> 3;   // 32 row groups per CTA
  float mx = 0.f;
  auto sq8 = [](const uint4& u) {
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u);
    float qq = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f2 = __bfloat1622float2(b2[e]);
      qq = fmaf(f2.x, f2.x, fmaf(f2.y, f2.y, qq));
    }
    return qq;
  };
  for (int t0 = rgrp; t0 < T; t0 += 128) {      // rows t0, t0 + 32, t0 + 64, t0 + 96
    uint4 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      u[i] = (t0 + 32 * i < T) ? __ldg(base + static_cast<size_t>(t0 + 32 * i) * 8 + sub) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float qq = sq8(u[i]);
      qq += __shfl_xor_sync(0xffffffffu, qq, 1);
      qq += __shfl_xor_sync(0xffffffffu, qq, 2);
      qq += __shfl_xor_sync(0xffffffffu, qq, 4);
      mx = fmaxf(mx, qq);
    }
  }
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    out[blockIdx.x] = m;   // squared, like the QKV epilogue's running maximum (GemmParams::kmax_sq)
  }
}

template <int GRID>
const char* launch_t(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const __nv_bfloat16* relh,
                     const __nv_bfloat16* relw, int G, int heads, const float* kmax, __nv_bfloat16* out, cudaStream_t stream) {
  using A = AttCfg<GRID>;
  const uint64_t rows = static_cast<uint64_t>(G) * heads * A::T;
  CUtensorMap tq, tk, tv, trh, trw;
  const char* e;
  if ((e = make_tmap_2d_bf16(&tq, q, rows, 64, 128, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tk, k, rows, 64, A::BN, 64))) return e;
  if ((e = make_tmap_2d_bf16(&tv, v, rows, 64, A::BN, 64))) return e;
  if ((e = make_tmap_2d_bf16(&trh, relh, A::NREL, 64, A::NREL, 64))) return e;
  if ((e = make_tmap_2d_bf16(&trw, relw, A::NREL, 64, A::NREL, 64))) return e;
  static PerDeviceOnce configured;
  auto kern = attention_kernel<GRID>;
  if ((e = set_max_smem_once(configured, kern, A::TOTAL))) return e;
  dim3 grid((A::T + A::NROWS - 1) / A::NROWS, G * heads);
  cudaError_t ce = launch_pdl(kern, grid, dim3(A::THREADS), A::TOTAL, stream, 1, tq, tk, tv, trh, trw, heads, kmax, out);
  return ce == cudaSuccess ? nullptr : cudaGetErrorString(ce);
}

}  // namespace

const char* launch_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v,
                             const __nv_bfloat16* rel_h, const __nv_bfloat16* rel_w, int G, int heads, int grid,
                             __nv_bfloat16* out, cudaStream_t stream, float* kmax_scratch, bool kmax_valid) {
  if (G <= 0) return nullptr;
  if (grid == 64) {
    // pre-pass: max key norm per (image, head) for the a-priori score bound (see the kernel); scratch of G * heads floats
    // from the caller, or stream-ordered here
    float* kmax = kmax_scratch;
    if (kmax == nullptr && cudaMallocAsync(&kmax, static_cast<size_t>(G) * heads * sizeof(float), stream) != cudaSuccess)
      return "attention: cudaMallocAsync failed";
    if (kmax_valid && kmax_scratch == nullptr) return "attention: kmax_valid without a kmax buffer";
    if (!kmax_valid && launch_pdl(k_norm_max_kernel, dim3(G * heads), dim3(256), 0, stream, 1, k, 4096, kmax) != cudaSuccess)
      return "attention: key-norm pre-pass launch failed";
    const char* e = launch_t<64>(q, k, v, rel_h, rel_w, G, heads, kmax, out, stream);
    if (kmax_scratch == nullptr) cudaFreeAsync(kmax, stream);
    return e;
  }
  if (grid == 14) {
    if (G % 25 != 0) return "attention: windowed mode needs 25 windows per image";
    // default: the four-query-tiles-per-SM design (attention_win.cu); ULLSAM_WIN_ATT=1 selects the first design.
    // (A third, whole-window one-shot design was measured on B200 in round 2: parity-green but 0.475 ms against
    //  0.414 ms per launch for this one, so it was removed — profiles/README.md.)
    static const int win_ver = getenv("ULLSAM_WIN_ATT") ? atoi(getenv("ULLSAM_WIN_ATT")) : 2;
    if (win_ver == 2) return launch_attention_windowed_v2(q, k, v, rel_h, rel_w, G, heads, out, stream);
    return launch_t<14>(q, k, v, rel_h, rel_w, G, heads, nullptr, out, stream);
  }
  return "attention: grid must be 64 (global) or 14 (windowed)";
}

}  // namespace ullsam
