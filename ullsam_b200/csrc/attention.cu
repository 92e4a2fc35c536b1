// ullsam_b200 — fused attention with in-kernel decomposed relative-position bias (sm_100a, tcgen05).
//
// Reference semantics (modeling/image_encoder.py:224-240, 325-361):
//   attn = softmax( (q*scale) k^T + rel_h[q, kh] + rel_w[q, kw] ) v        rel_* use the UNSCALED q
//   rel_h[q, kh] = q . rel_pos_h[qh - kh + GRID-1],   rel_w[q, kw] = q . rel_pos_w[qw - kw + GRID-1]
// for the global 64x64 grid (T = 4096 keys) and for 14x14 windows (T = 196 keys, all of them real keys:
// the zero-padded rows of edge windows carry k = v = qkv.bias and are attended to, :168-172, 243-264).
//
// One CTA = 256 query rows (two 128-row tiles) of one (group, head):
//   warps 0-3 / 4-7 : softmax warpgroup for query tile 0 / 1 (thread == query row == TMEM lane)
//   warp 8          : TMA producer (Q once; K/V tiles through an mbarrier ring; the rel-pos tables)
//   warp 9          : tcgen05.mma issuer (single thread) + TMEM allocator
// Flash-style loop over key tiles of RPT grid rows (global: 2 x 64 = 128 keys; windowed: 8 x 14 = 112):
//   S = Q K^T        (SS MMA, fp32 in TMEM)            -> softmax warpgroup adds the bias, exponentiates,
//   P (bf16, smem, 128B-swizzled K-major)                 writes P and signals the MMA warp
//   O += P V         (SS MMA, V is the MN-major operand, O accumulates in TMEM)
// The two query tiles ping-pong on the tensor core, so one tile's softmax overlaps the other's MMAs.
// The running max is only refreshed when it grows by more than 2^8 (exact: softmax is invariant to the
// reference point; O and the row sum are rescaled together), so the O round-trip through registers is rare.
//
// Bias: the prologue computes Tw = Q rel_pos_w^T and Th = Q rel_pos_h^T with two extra MMAs per query tile
// (tables zero-padded to NREL rows), each thread gathers its own 2*GRID values through a shared-memory
// scatter, keeps rel_w in registers and (global) parks rel_h as packed half2 in spare TMEM columns.
#include "attention.h"

#include "common.cuh"
#include "gemm.h"

namespace ullsam {

namespace {

constexpr int ATT_THREADS = 384;  // 2 softmax warpgroups + 1 control warpgroup (TMA, MMA, 2 idle warps)
constexpr float kLog2e = 1.4426950408889634f;

template <int GRID>
struct AttCfg {
  static constexpr int T = GRID * GRID;
  static constexpr int RPT = (GRID == 64) ? 2 : 8;        // key-grid rows per tile
  static constexpr int BN = RPT * GRID;                   // keys per tile: 128 / 112
  static constexpr int NT = (T + BN - 1) / BN;            // 32 / 2
  static constexpr int NREL = (GRID == 64) ? 128 : 32;    // padded rows of the rel-pos tables
  static constexpr int STAGES = 3;
  static constexpr int QBYTES = 128 * 128;                // one 128 x 64 bf16 tile
  static constexpr int KVBYTES = 128 * 128;               // stage slot for K (and for V), BN <= 128 rows
  static constexpr int OFF_Q = 0;                         // 2 tiles
  static constexpr int OFF_KV = 2 * QBYTES;               // STAGES x (K slot, V slot)
  static constexpr int OFF_P = OFF_KV + STAGES * 2 * KVBYTES;  // 2 q-tiles x 2 k-blocks x 16 KB
  static constexpr int OFF_BAR = OFF_P + 4 * 16384;
  static constexpr int NBAR = 3 + 2 * STAGES + 6;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16 + 1024;
  // TMEM columns
  static constexpr int COL_S = 0;       // + qt*128
  static constexpr int COL_O = 256;     // + qt*64
  static constexpr int COL_RELH = 384;  // + qt*32 (global only: packed half2 rel_h, column j = rows 2j, 2j+1)
};

ULLSAM_DEVICE float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

ULLSAM_DEVICE void tmem_ld1(uint32_t taddr, uint32_t& v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}

template <int GRID>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmRh,
                 const __grid_constant__ CUtensorMap tmRw, int heads, __nv_bfloat16* __restrict__ out) {
  using A = AttCfg<GRID>;
  constexpr int S = A::STAGES;
  constexpr int BN = A::BN;
  constexpr int NT = A::NT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* pro_mma_done = bars + 1;
  uint64_t* pro_consumed = bars + 2;
  uint64_t* kv_full = bars + 3;
  uint64_t* kv_empty = kv_full + S;
  uint64_t* s_full = kv_empty + S;   // [2]
  uint64_t* p_ready = s_full + 2;    // [2]
  uint64_t* o_done = p_ready + 2;    // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gh = blockIdx.y;                       // group * heads + head
  const int qblock = blockIdx.x;                   // 256-row block of queries inside the group
  const int row_base = gh * A::T;                  // first row of this (group, head) in q/k/v
  const int q_row0 = row_base + qblock * 256;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    mbar_init(pro_mma_done, 1);
    mbar_init(pro_consumed, 8);
    for (int i = 0; i < S; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= 8) {
    // (register rebalancing via setmaxnreg measured worse with ptxas 12.9 here)
  } else {
    ;
  }

  if (warp == 8) {
    // =========================================== TMA producer ===========================================
    if (lane == 0) {
      tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
      uint8_t* kv = smem + A::OFF_KV;
      // Q tiles + rel-pos tables (tables ride in the LAST ring stage until the prologue MMAs retire)
      mbar_arrive_expect_tx(q_full, 2 * A::QBYTES + 2 * A::NREL * 128);
      tma_load_2d(smem + A::OFF_Q, &tmQ, q_full, 0, q_row0);
      tma_load_2d(smem + A::OFF_Q + A::QBYTES, &tmQ, q_full, 0, q_row0 + 128);
      tma_load_2d(kv + (S - 1) * 2 * A::KVBYTES, &tmRw, q_full, 0, 0);
      tma_load_2d(kv + (S - 1) * 2 * A::KVBYTES + A::KVBYTES, &tmRh, q_full, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < NT; ++j) {
        if (j == S - 1) mbar_wait(pro_mma_done, 0);
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], 2 * BN * 128);
        tma_load_2d(kv + stage * 2 * A::KVBYTES, &tmK, &kv_full[stage], 0, row_base + j * BN);
        tma_load_2d(kv + stage * 2 * A::KVBYTES + A::KVBYTES, &tmV, &kv_full[stage], 0, row_base + j * BN);
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {
    // =========================================== MMA issuer =============================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, BN, 0, 0);
      constexpr uint32_t idesc_rel = make_idesc_bf16(128, A::NREL, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // V is the MN-major B operand
      const uint32_t sq = smem_u32(smem + A::OFF_Q);
      const uint32_t skv = smem_u32(smem + A::OFF_KV);
      const uint32_t sp = smem_u32(smem + A::OFF_P);

      // ---- prologue: Tw[qt] = Q[qt] Rw^T, Th[qt] = Q[qt] Rh^T ----
      mbar_wait(q_full, 0);
      tc_fence_after();
      {
        const uint64_t drw = make_smem_desc_sw128(skv + (S - 1) * 2 * A::KVBYTES, 16, 1024);
        const uint64_t drh = make_smem_desc_sw128(skv + (S - 1) * 2 * A::KVBYTES + A::KVBYTES, 16, 1024);
#pragma unroll
        for (int qt = 0; qt < 2; ++qt) {
          const uint64_t dq = make_smem_desc_sw128(sq + qt * A::QBYTES, 16, 1024);
          const uint32_t col_w = (GRID == 64) ? qt * 128 : qt * 64;
          const uint32_t col_h = (GRID == 64) ? 256 + qt * 128 : qt * 64 + 32;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + col_w, dq + 2 * k, drw + 2 * k, idesc_rel, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss(tmem_base + col_h, dq + 2 * k, drh + 2 * k, idesc_rel, k != 0);
        }
      }
      umma_commit(pro_mma_done);
      mbar_wait(pro_consumed, 0);
      tc_fence_after();

      auto issue_qk = [&](int qt, int stage) {
        const uint64_t dq = make_smem_desc_sw128(sq + qt * A::QBYTES, 16, 1024);
        const uint64_t dk = make_smem_desc_sw128(skv + stage * 2 * A::KVBYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem_base + A::COL_S + qt * 128, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[qt]);
      };
      auto issue_pv = [&](int qt, int stage, int j) {
        const uint64_t dv = make_smem_desc_sw128(skv + stage * 2 * A::KVBYTES + A::KVBYTES, 16, 1024);
#pragma unroll
        for (int s = 0; s < BN / 16; ++s) {
          const uint64_t dp = make_smem_desc_sw128(sp + qt * 32768 + (s >> 2) * 16384, 16, 1024) + 2 * (s & 3);
          // V tile: 16 keys further down = 16 rows x 128 B = 2048 B
          umma_ss(tmem_base + A::COL_O + qt * 64, dp, dv + s * 128, idesc_pv, (j | s) != 0);
        }
        umma_commit(&o_done[qt]);
      };

      int stage = 0;
      uint32_t phase = 0;
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      issue_qk(1, 0);
      for (int j = 0; j < NT; ++j) {
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == S) { nstage = 0; nphase ^= 1; }
#pragma unroll
        for (int qt = 0; qt < 2; ++qt) {
          mbar_wait(&p_ready[qt], j & 1);
          tc_fence_after();
          issue_pv(qt, stage, j);
          if (j + 1 < NT) {
            if (qt == 0) {
              mbar_wait(&kv_full[nstage], nphase);
              tc_fence_after();
            }
            issue_qk(qt, nstage);
          }
        }
        umma_commit(&kv_empty[stage]);
        stage = nstage;
        phase = nphase;
      }
    }
  } else if (warp < 8) {
    // =========================================== softmax warpgroups =====================================
    const int qt = warp >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;           // row inside the query tile == TMEM lane
    const int tid256 = qt * 128 + r;          // row inside the CTA
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int tq = qblock * 256 + tid256;     // query index inside the group
    const int qh = tq / GRID, qw = tq % GRID;
    float* scratch = reinterpret_cast<float*>(smem + A::OFF_P);  // [NRELV][256] fp32, free until the first P

    constexpr int NW = (GRID == 64) ? 64 : 14;
    float relw[NW];
    float relh_reg[(GRID == 64) ? 1 : 14];

    // ------------------------------------ prologue: gather the bias rows ------------------------------
    mbar_wait(pro_mma_done, 0);
    tc_fence_after();
    {
      const uint32_t col_w = (GRID == 64) ? qt * 128 : qt * 64;
      const uint32_t col_h = (GRID == 64) ? 256 + qt * 128 : qt * 64 + 32;
      // rel_w: Tw[row][i], i = qw - kw + GRID-1  ->  scratch[kw][row]
#pragma unroll
      for (int c = 0; c < A::NREL / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + col_w + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kw = qw + (GRID - 1) - (c * 32 + i);
          if (kw >= 0 && kw < GRID) scratch[kw * 256 + tid256] = __uint_as_float(v[i]) * kLog2e;
        }
      }
#pragma unroll
      for (int kw = 0; kw < NW; ++kw) relw[kw] = scratch[kw * 256 + tid256];
      // rel_h: Th[row][i], i = qh - kh + GRID-1  ->  scratch[kh][row]   (same thread re-uses its own column)
#pragma unroll
      for (int c = 0; c < A::NREL / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + col_h + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kh = qh + (GRID - 1) - (c * 32 + i);
          if (kh >= 0 && kh < GRID) scratch[kh * 256 + tid256] = __uint_as_float(v[i]) * kLog2e;
        }
      }
      if constexpr (GRID == 64) {
        uint32_t packed[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const __half2 h = __floats2half2_rn(scratch[(2 * j) * 256 + tid256], scratch[(2 * j + 1) * 256 + tid256]);
          packed[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        // every prologue TMEM read of both warpgroups must retire before columns 384.. are overwritten
        asm volatile("bar.sync 2, 256;" ::: "memory");
        tmem_st32(lane_addr + A::COL_RELH + qt * 32, packed);
        tmem_wait_st();
        relh_reg[0] = 0.f;
      } else {
#pragma unroll
        for (int kh = 0; kh < 14; ++kh) relh_reg[kh] = scratch[kh * 256 + tid256];
      }
    }
    tc_fence_before();
    // the scratch region becomes the P buffers: make sure both warpgroups are done with it
    asm volatile("bar.sync 3, 256;" ::: "memory");
    if (lane == 0) mbar_arrive(pro_consumed);

    // ------------------------------------ main loop ---------------------------------------------------
    const float cs = 0.125f * kLog2e;  // head_dim 64: scale = 1/8, folded with log2(e)
    float m_used = -INFINITY;
    float l = 0.f;
    uint8_t* prow = smem + A::OFF_P + qt * 32768 + r * 128;
    const int rx = r & 7;

#pragma unroll 1
    for (int j = 0; j < NT; ++j) {
      mbar_wait(&s_full[qt], j & 1);
      tc_fence_after();
      float rh0 = 0.f, rh1 = 0.f;
      if constexpr (GRID == 64) {
        uint32_t hp;
        tmem_ld1(lane_addr + A::COL_RELH + qt * 32 + j, hp);
        tmem_wait_ld();
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hp));
        rh0 = f.x;
        rh1 = f.y;
      }
      float rht[8];  // windowed: rel_h of the 8 key-grid rows of this tile
      if constexpr (GRID != 64) {
#pragma unroll
        for (int a = 0; a < 8; ++a) rht[a] = (j == 0) ? relh_reg[a] : ((8 + a < 14) ? relh_reg[(8 + a) % 14] : 0.f);
      } else {
        rht[0] = 0.f;
      }
      // bias of column `col` (0..BN-1) of tile j, in log2 units, excluding the running max
      auto bias_of = [&](int col) -> float {
        if constexpr (GRID == 64) {
          return relw[col & 63] + ((col < 64) ? rh0 : rh1);
        } else {
          // keys of tile j: t = j*112 + col, kh = t / 14, kw = t % 14  (col / 14 and col % 14 are static)
          return relw[col % 14] + rht[col / 14];
        }
      };
      const int valid_cols = (GRID == 64) ? BN : ((j == NT - 1) ? (A::T - (NT - 1) * BN) : BN);

      // ---- pass 1: row max of the biased scores ----
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_addr + A::COL_S + qt * 128 + c * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = c * 32 + i;
          if (col < BN) {
            float t = fmaf(__uint_as_float(v[i]), cs, bias_of(col));
            if (GRID != 64 && col >= valid_cols) t = -INFINITY;
            mx = fmaxf(mx, t);
          }
        }
      }
      // ---- running max, refreshed lazily ----
      const bool need = mx > m_used + 8.0f;
      if (j > 0) {
        mbar_wait(&o_done[qt], (j - 1) & 1);  // PV(j-1) retired: O is quiescent and P may be overwritten
        tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2(m_used - mx) : 1.0f;
          l *= alpha;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tmem_ld32(lane_addr + A::COL_O + qt * 64 + c * 32, o);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(lane_addr + A::COL_O + qt * 64 + c * 32, o);
          }
          tmem_wait_st();
        }
      }
      if (need) m_used = mx;

      // ---- pass 2: P = exp2(t - m), row sum, bf16 P to shared memory (swizzled K-major) ----
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c * 32 >= BN) break;
        uint32_t v[32];
        tmem_ld32(lane_addr + A::COL_S + qt * 128 + c * 32, v);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = 0.f, p1 = 0.f;
          const int col = c * 32 + i;
          if (col < BN) {
            p0 = ex2(fmaf(__uint_as_float(v[i]), cs, bias_of(col) - m_used));
            if (GRID != 64 && col >= valid_cols) p0 = 0.f;
          }
          if (col + 1 < BN) {
            p1 = ex2(fmaf(__uint_as_float(v[i + 1]), cs, bias_of(col + 1) - m_used));
            if (GRID != 64 && col + 1 >= valid_cols) p1 = 0.f;
          }
          l += p0 + p1;
          pk[i >> 1] = pack_bf16(p0, p1);
        }
        uint8_t* dst = prow + (c >> 1) * 16384;
        const int cc0 = (c & 1) * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (c * 32 + u * 8 < BN)
            *reinterpret_cast<uint4*>(dst + (((cc0 + u) ^ rx) << 4)) =
                make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[qt]);
    }

    // ------------------------------------ epilogue: O / l -> out ---------------------------------------
    mbar_wait(&o_done[qt], (NT - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    long out_row = -1;
    if constexpr (GRID == 64) {
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
      tmem_ld32(lane_addr + A::COL_O + qt * 64 + c * 32, o);
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
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
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
  dim3 grid((A::T + 255) / 256, G * heads);
  kern<<<grid, ATT_THREADS, A::TOTAL, stream>>>(tq, tk, tv, trh, trw, heads, out);
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
