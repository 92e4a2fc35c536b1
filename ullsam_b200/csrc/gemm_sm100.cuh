// ullsam_b200 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = A[M,K] (bf16, K-major) x W[N,K]^T (bf16, K-major, i.e. an nn.Linear weight as stored)
//
// One CTA per SM, static tile scheduler (tile = blockIdx.x + i*gridDim.x, N fastest).
//   warp 0      : TMA producer  (A and W tiles, 128-byte swizzle, NSTAGE-deep mbarrier ring)
//   warp 1      : MMA issuer    (one elected lane, tcgen05.mma cta_group::1, 128 x BN x 16)
//   warp 2      : TMEM allocator
//   warps 4..11 : epilogue      (tcgen05.ld 32x32b.x32; thread == accumulator row; two warps per
//                                TMEM lane quadrant split the BN columns between them)
// The fp32 accumulator lives in TMEM and is double-buffered (2 x BN columns) so the epilogue of
// tile i overlaps the main loop of tile i+1.
//
// The epilogue is a compile-time policy (EPI_*) — this is where LayerNorm2d, GELU, residual adds,
// the q/k/v head split and the NCHW transpose of the reference's modules are fused.
#pragma once

#include "common.cuh"
#include "gemm.h"

namespace ullsam {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps (the EW = 16 variant runs 640 threads)
constexpr int GEMM_EPI_WARP0 = 4;
#ifndef ULLSAM_GEMM_RES_PREFETCH
// L2 prefetch of the next tile's addend, one tile period ahead (0 = off); only for K <= 1024 (with
// K = 3072 it cost lin2 3 %).  On top of the register-level prefetch (one chunk ahead, shorter than a DRAM round
// trip under load): proj 213.5 -> 206.1 us.
#define ULLSAM_GEMM_RES_PREFETCH 1
#endif
// bit (j & 3) set: element j of a chunk takes the 16th power of the GELU's erf by RCP + squarings instead of LG2/EX2
#ifndef ULLSAM_GELU_RCP_MASK
#define ULLSAM_GELU_RCP_MASK 0x1  // measured on lin1: 0x0 512 us, 0x1 503, 0x5 506, 0xF 519.  (Round 2: a MUFU-free degree-7
                                  // polynomial form on 1/4, 1/2, 3/4 or all of the elements: 489, 494, 502, 501 us against 490 —
                                  // the epilogue is bound by its instruction count, not by the special-function unit.)
#endif
#ifndef ULLSAM_GELU_LOGISTIC
#define ULLSAM_GELU_LOGISTIC 1   // lin1's GELU as gelu_logistic() (common.cuh: |error| <= 2.6e-5, 9 instructions) instead of the A&S erf (13)
#endif
#ifndef ULLSAM_GELU_PIPE
#define ULLSAM_GELU_PIPE 0   // 1: the 8-warp GELU epilogue also double-buffers its TMEM loads (fully unrolled chunk loop)
#endif
#ifndef ULLSAM_GEMM_EARLY_RELEASE
#define ULLSAM_GEMM_EARLY_RELEASE 1
#endif
#ifndef ULLSAM_GEMM_BIAS_SMEM
#define ULLSAM_GEMM_BIAS_SMEM 1
#endif
#ifndef ULLSAM_GEMM_RELAXED_ARRIVE
#define ULLSAM_GEMM_RELAXED_ARRIVE 1
#endif
// timing experiments only (wrong results): FOLD_DBG bit 0 = no row-statistics loads, bit 1 = no fold arithmetic;
// STATS_DBG bit 0 = no bf16 copy, bit 1 = no partial sums
#ifndef ULLSAM_FOLD_DBG
#define ULLSAM_FOLD_DBG 0
#endif
#ifndef ULLSAM_STATS_DBG
#define ULLSAM_STATS_DBG 0
#endif
#ifndef ULLSAM_FOLD_L2PF
#define ULLSAM_FOLD_L2PF 1   // L2 prefetch of the next tile's row statistics, one tile period ahead
#endif
#ifndef ULLSAM_GEMM_WIDE_F32
#define ULLSAM_GEMM_WIDE_F32 1  // fp32 epilogues: full 128-byte lines per row + addend loads one chunk ahead
#endif

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile —
// each CTA stages its own 128 rows of A and its own BN/2 rows of W, the leader issues one 256 x BN x 16 MMA that
// reads both CTAs' shared memory, and each CTA's TMEM receives its 128 accumulator rows.  Halving the per-CTA W
// traffic is what lifts the shared-memory / L2 operand-feed bound of the single-CTA kernel.
// RED: the TMA reduce-add epilogue (EPI_RES_RED) — two 4 KB staging boxes per epilogue warp, paid for with two pipeline
// stages (its K = 768 main loop has 12 k-blocks; four stages cover it).
template <int BN, int CG = 1, int EW = 8, bool RED = false>
struct GemmSmem {
  static constexpr int kStages = RED ? 4 : ((BN == 256 && CG == 1) ? 4 : (EW == 16 ? 5 : 6));
  static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;
  static constexpr int kBBytes = (BN / CG) * GEMM_BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  // full[S], empty[S], tmem_full[2], tmem_empty[2], tmem ptr, LN scratch (2 x 2 x 128 floats)
  static constexpr int kScratchOffset = kBarOffset + (2 * kStages + 4) * 8 + 16;
  // warp-private transpose tiles of the generic epilogues: 32 rows x 80-byte pitch per epilogue warp
  static constexpr int kStageBytesPerWarp = 32 * 80;
  static constexpr int kStageOffset = (kScratchOffset + 4 * 128 * 4 + 15) / 16 * 16;
  // warp-private copy of the bias slice of the current tile (generic epilogues): up to 128 floats per epilogue warp
  static constexpr int kBiasBytesPerWarp = 1024;   // (second half: the column sums of the LayerNorm-folded epilogues)
  static constexpr int kBiasOffset = kStageOffset + EW * kStageBytesPerWarp;
  static constexpr int kRedBytesPerWarp = RED ? 2 * 4096 : 0;
  static constexpr int kRedOffset = (kBiasOffset + EW * kBiasBytesPerWarp + 1023) / 1024 * 1024;
  static constexpr int kTotal = kRedOffset + EW * kRedBytesPerWarp + 1024 /* alignment slack */;
  static_assert(kTotal <= 232448, "more than 227 KB of shared memory");
};


// Second pass of EPI_RES_LN for one warp: rows [row0, row0 + 32) x its 128-column half of each of the NT 256-wide N
// tiles (exactly the elements this warp has just stored).  mean / rstd of row r live in lanes (r % 8) * 4 .. + 3 at
// index r / 8.  The L2 reads are latency-bound, so the loads of RB rows x NT tiles are issued before any is consumed.
template <int NT>
ULLSAM_DEVICE void res_ln_second_pass(const GemmParams& p, int row0, int half, int lane, const float (&mean)[4],
                                      const float (&rstd)[4]) {
  constexpr int RB = 4;
  const float* xsrc = reinterpret_cast<const float*>(p.out);
  __nv_bfloat16* xn = reinterpret_cast<__nv_bfloat16*>(p.out2);
  const int col0 = half * 128 + lane * 4;
  float4 g4[NT], b4[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    g4[nt] = __ldg(reinterpret_cast<const float4*>(p.ln_w + nt * 256 + col0));
    b4[nt] = __ldg(reinterpret_cast<const float4*>(p.ln_b + nt * 256 + col0));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int r0 = 0; r0 < 8; r0 += RB) {
      float4 v4[RB][NT];
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        const int grow = row0 + i * 8 + r0 + rb;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
          v4[rb][nt] = grow < p.M ? *reinterpret_cast<const float4*>(xsrc + static_cast<size_t>(grow) * p.ldc + nt * 256 + col0)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        const float m = __shfl_sync(0xffffffffu, mean[i], (r0 + rb) * 4);
        const float rs = __shfl_sync(0xffffffffu, rstd[i], (r0 + rb) * 4);
        const int grow = row0 + i * 8 + r0 + rb;
        if (grow >= p.M) continue;
        size_t orow = grow;
        if (p.ln_partition) {
          const int b = grow >> 12, t = grow & 4095, y = t >> 6, x = t & 63;
          const int wy = y / 14, wx = x / 14;
          orow = static_cast<size_t>(b * 25 + wy * 5 + wx) * 196 + (y - wy * 14) * 14 + (x - wx * 14);
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const float4 x4 = v4[rb][nt];
          const uint32_t lo = pack_bf16(fmaf((x4.x - m) * rs, g4[nt].x, b4[nt].x), fmaf((x4.y - m) * rs, g4[nt].y, b4[nt].y));
          const uint32_t hi = pack_bf16(fmaf((x4.z - m) * rs, g4[nt].z, b4[nt].z), fmaf((x4.w - m) * rs, g4[nt].w, b4[nt].w));
          *reinterpret_cast<uint2*>(xn + orow * p.N + nt * 256 + col0) = make_uint2(lo, hi);
        }
      }
    }
  }
}

// EW epilogue warps: 8 (two per TMEM lane quadrant, BN/2 columns each) or 16 (four per quadrant, BN/4 columns each:
// twice the warps to hide the TMEM / global latencies of the epilogue and to spread the GELU arithmetic).
template <int BN, int EPI, int CG = 1, int EW = 8>
__global__ void __launch_bounds__((GEMM_EPI_WARP0 + EW) * 32, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC /* EPI_RES_RED: the fp32 output, 32 x 32 boxes; otherwise unused */, const GemmParams p) {
  using L = GemmSmem<BN, CG, EW, EPI == EPI_RES_RED>;
  static_assert(EPI != EPI_RES_RED || (BN == 256 && EW == 8), "EPI_RES_RED: 256-wide tiles, 8 epilogue warps");
  static_assert(EW == 8 || (EPI != EPI_LN_BF16 && EPI != EPI_LN_NCHW && EPI != EPI_LN_RES && EPI != EPI_UP1 && EPI != EPI_MASK),
                "the LayerNorm / decoder epilogues assume two warps per row");
  static_assert(CG == 1 || BN == 256, "the CTA-pair kernel is built for 256-wide N tiles");
  static_assert(EPI != EPI_RES_LN || (BN == 256 && EW == 8), "EPI_RES_LN: 256-wide tiles, two 128-column halves per row");
  // LayerNorm folded into the neighbouring GEMMs (gemm.h): kStats = producer side, kFold = consumer side
  constexpr bool kStats = (EPI == EPI_RES_STATS || EPI == EPI_PATCH_STATS);
  constexpr bool kFold = (EPI == EPI_QKV_FOLD || EPI == EPI_GELU_FOLD);
  constexpr bool kIsQkv = (EPI == EPI_QKV || EPI == EPI_QKV_FOLD);
  constexpr bool kIsGelu = (EPI == EPI_GELU_BF16 || EPI == EPI_GELU_FOLD);
  constexpr bool kIsPatch = (EPI == EPI_PATCH || EPI == EPI_PATCH_STATS);
  static_assert(!kStats || (BN == 256 && EW == 8), "statistics epilogues: 256-wide tiles, one 128-column slice per warp");
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  constexpr int S = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms, computed as an offset so the pointer stays in the shared window
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* scratch = reinterpret_cast<float*>(smem + L::kScratchOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = (p.M + CG * GEMM_BM - 1) / (CG * GEMM_BM);  // in units of CG*128 rows
  const int tiles_n = p.N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = p.K / GEMM_BK;
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;    // tiles are dealt to CTAs (CG=1) or CTA pairs (CG=2)
  // i-th tile of this CTA (pair).  Default: round robin over the N-fastest tile list.  EPI_RES_LN: M-major — all N tiles
  // of one row block are processed back to back by the same CTA pair, so its epilogue threads see whole rows.
  constexpr bool kMMajor = (EPI == EPI_RES_LN);
  constexpr bool kRes = (EPI == EPI_RES_F32 || EPI == EPI_RES_LN || EPI == EPI_RES_STATS);
  auto tile_at = [&](int i) -> int {
    if constexpr (kMMajor) return (tile0 + (i / tiles_n) * tile_step) * tiles_n + (i % tiles_n);
    else return tile0 + i * tile_step;
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EW * CG);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) {
      tmem_alloc_2cta(tmem_ptr_smem, 512);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_ptr_smem, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync();  // peer barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // (barriers, tensor memory and descriptors are set up; from here on the kernel reads what its predecessor wrote)
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ================================ TMA producer ================================
    // (one lane chosen by elect.sync: with a plain `lane == 0` branch ptxas wraps every TMA / tcgen05 instruction
    //  in an ELECT / BRA.U.ANY loop over the active lanes)
    if (elect_one_sync()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int pit = 0, tile = tile_at(0); tile < num_tiles; tile = tile_at(++pit)) {
        const int m_blk = (tile / tiles_n) * CG + cta_rank, n_blk = tile % tiles_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if constexpr (CG == 2) {
            // both CTAs' bytes are accounted on the LEADER's barrier (the MMA issuer lives there)
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            const uint32_t bar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_2cta(sa, &tmA, bar, kb * GEMM_BK, m_blk * GEMM_BM);
            tma_load_2d_2cta(sb, &tmB, bar, kb * GEMM_BK, n_blk * BN + cta_rank * (BN / 2));
            if (++stage == S) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          if (p.a_conv3) {
            const int cblocks = p.conv_c / GEMM_BK;
            const int tap = kb / cblocks, cb = kb % cblocks;
            const int dy = tap / 3, dx = tap % 3;
            const int b = m_blk / 32, y0 = (m_blk % 32) * 2;
            tma_load_4d(sa, &tmA, &full_bar[stage], cb * GEMM_BK, dx - 1, y0 + dy - 1, b);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
          }
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (cta_rank == 0 && elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_bf16(CG * GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0, tile = tile_at(0); tile < num_tiles; tile = tile_at(++it)) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sb = sa + L::kABytes;
          const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // +32 bytes along K inside the 128-byte swizzle row == +2 in the (addr >> 4) field
            if constexpr (CG == 2) umma_ss_2cta(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            else umma_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          if constexpr (CG == 2) umma_commit_2cta(&empty_bar[stage]);  // frees the slot in BOTH CTAs
          else umma_commit(&empty_bar[stage]);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        if constexpr (CG == 2) umma_commit_2cta(&tfull_bar[acc]);
        else umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= GEMM_EPI_WARP0) {
    // ================================ epilogue ====================================
    const int quad = warp & 3;
    const int half = (warp - GEMM_EPI_WARP0) >> 2;  // which column slice of the tile this warp owns (0 .. EW/4-1)
    const int row_in_tile = quad * 32 + lane;
    constexpr int COLS_PER_WARP = BN / (EW / 4);
    constexpr int NCHUNK = COLS_PER_WARP / 32;
    float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};  // EPI_RES_LN: row sums / sums of squares
    // Residual epilogues: every lane asks L2 for its row's slice of the addend of the tile AFTER the one being drained
    // (one tile period, ~10 us, ahead), so the epilogue's addend loads are L2 hits instead of DRAM round trips.
    auto prefetch_addend = [&](int t) {
#if ULLSAM_GEMM_RES_PREFETCH
      if constexpr (kRes) {
        if (t < num_tiles && p.K <= 1024) {  // (short main loops only: with K = 3072 it competes with the operand loads)
          const int r = ((t / tiles_n) * CG + cta_rank) * GEMM_BM + row_in_tile;
          if (r < p.M) {
            const float* src = p.residual + static_cast<size_t>(r) * p.ldc + (t % tiles_n) * BN + half * COLS_PER_WARP;
            // (per-thread prefetch.global.L2; the bulk form, UBLKPF.L2, takes uniform registers and is serialised
            //  over the 32 lanes by ptxas)
#pragma unroll
            for (int b = 0; b < COLS_PER_WARP * 4; b += 128) l2_prefetch_line(reinterpret_cast<const uint8_t*>(src) + b);
          }
        }
      }
#endif
    };
    prefetch_addend(tile_at(0));
    // fp32 epilogues (kWide): a 32-column chunk is one 128-byte line per row, staged 16 rows at a time, so each global
    // instruction covers 4 whole lines (8 lanes x 16 bytes per row).  The addend (residual stream / pos_embed) of chunk
    // c+1 — across the tile boundary, of the next tile's first chunk — is requested before chunk c is processed: its
    // DRAM latency is hidden behind a chunk of work or the wait for the accumulator (ncu before: 50 % of the
    // projection GEMM's epilogue samples were long-scoreboard stalls on these loads).
    constexpr bool kAdd = kRes || kIsPatch;
    constexpr bool kWide = ULLSAM_GEMM_WIDE_F32 && EW == 8 && (EPI == EPI_RES_F32 || EPI == EPI_PATCH || EPI == EPI_F32 || kStats);
    static_assert(!kStats || kWide, "statistics epilogues are written for the full-line fp32 path");
    // kStats: per-lane partial (sum, sum of squares) of the 8 rows this lane touches in the transposed phase ([h][i]: row
    // h * 16 + i * 4 + lane / 8, columns (lane & 7) * 4 .. + 3 of every chunk), reduced over the 8 lanes at the end of the tile
    float rs1[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, rs2[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    constexpr bool kAddPipe = kWide && kAdd;
    static_assert(!kAddPipe || NCHUNK % 2 == 0, "addend double buffer: the buffer index must wrap at the tile boundary");
    float4 addbuf[2][2][4];
    auto load_addend = [&](float4 (&dst)[2][4], int m_blk_, int n0_) {
      const int base_row = m_blk_ * GEMM_BM + quad * 32 + (lane >> 3);
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int grow = base_row + h * 16 + i * 4;
          dst[h][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (grow < p.M) {
            if constexpr (kRes)
              dst[h][i] = *reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(grow) * p.ldc + n0_ + (lane & 7) * 4);
            else
              dst[h][i] = __ldg(reinterpret_cast<const float4*>(
                  p.pos + static_cast<size_t>(grow % p.pos_rows) * p.N + n0_ + (lane & 7) * 4));
          }
        }
    };
    // Generic epilogues: the bias slice of the NEXT tile is fetched into a register while the current tile is drained
    // and parked in shared memory when that tile starts; the per-chunk reads are broadcast LDS.  (Loading it with
    // __ldg inside the chunk loop exposed one L2 round trip per chunk: ncu, 600 of 2400 samples per epilogue warp.)
#if ULLSAM_GEMM_BIAS_SMEM
    float* bias_s = reinterpret_cast<float*>(smem + L::kBiasOffset + (warp - GEMM_EPI_WARP0) * L::kBiasBytesPerWarp);
    float4 bias_next = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool bias_lane = p.bias != nullptr && lane * 4 < COLS_PER_WARP;
    static_assert(COLS_PER_WARP * 4 <= L::kBiasBytesPerWarp, "bias slice does not fit its shared-memory slot");
    float4 folds_next = make_float4(0.f, 0.f, 0.f, 0.f);
    float* folds_s = bias_s + 128;
    static_assert(COLS_PER_WARP <= 128, "bias / column-sum slices: 128 floats each");
    if (bias_lane && tile_at(0) < num_tiles) {
      bias_next = __ldg(reinterpret_cast<const float4*>(p.bias + (tile_at(0) % tiles_n) * BN + half * COLS_PER_WARP + lane * 4));
      if constexpr (kFold)
        folds_next = __ldg(reinterpret_cast<const float4*>(p.fold_s + (tile_at(0) % tiles_n) * BN + half * COLS_PER_WARP + lane * 4));
    }
#else
    static_assert(!kFold, "the folded epilogues read their vectors from the shared-memory slots");
#endif
    if constexpr (kAddPipe) {
      const int t0 = tile_at(0);
      if (t0 < num_tiles) load_addend(addbuf[0], (t0 / tiles_n) * CG + cta_rank, (t0 % tiles_n) * BN + half * COLS_PER_WARP);
    }
    for (int it = 0, tile = tile_at(0); tile < num_tiles; tile = tile_at(++it)) {
      const int m_blk = (tile / tiles_n) * CG + cta_rank, n_blk = tile % tiles_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      prefetch_addend(tile_at(it + 1));
      const int row = m_blk * GEMM_BM + row_in_tile;
      const bool row_ok = row < p.M;
      // kFold: the row's partial sums are requested before the wait for the accumulator (their latency hides behind it)
      float fold_r = 1.f, fold_t = 0.f;   // rstd and -rstd * mean of this thread's row
      float4 fst[kStatSlots / 2];
      if constexpr (kFold) {
        const float4* sp = reinterpret_cast<const float4*>(p.row_stats) + static_cast<size_t>(row_ok ? row : 0) * (kStatSlots / 2);
        const int nsl = p.K >> 8;          // pairs of 128-column slices (K = LayerNorm dim, a multiple of 256)
#pragma unroll
        for (int j = 0; j < kStatSlots / 2; ++j)
          fst[j] = (j < nsl && !(ULLSAM_FOLD_DBG & 1)) ? sp[j] : make_float4(0.f, 0.f, 0.f, 0.f);
#if ULLSAM_FOLD_L2PF
        const int tn = tile_at(it + 1);
        if (tn < num_tiles) {
          const int rn = ((tn / tiles_n) * CG + cta_rank) * GEMM_BM + row_in_tile;
          if (rn < p.M && rn != row) l2_prefetch_line(reinterpret_cast<const uint8_t*>(p.row_stats) + static_cast<size_t>(rn) * (kStatSlots * 8));
        }
#endif
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if constexpr (kFold) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < kStatSlots / 2; ++j) {   // fixed order: bit-reproducible whatever the producer's schedule was
          s1 += fst[j].x; s2 += fst[j].y;
          s1 += fst[j].z; s2 += fst[j].w;
        }
        const float inv_k = 1.0f / static_cast<float>(p.K);
        const float mean = s1 * inv_k;
        const float var = fmaxf(s2 * inv_k - mean * mean, 0.f);
        fold_r = rsqrtf(var + p.ln_eps);
        fold_t = -fold_r * mean;
      }
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN + half * COLS_PER_WARP;
      const int col0 = n_blk * BN + half * COLS_PER_WARP;
      // Hands this accumulator buffer back to the MMA warp.  The generic epilogues call it as soon as their last
      // tcgen05.ld has landed in registers (before the math and the stores of the last chunk): the buffer is held for
      // a shorter time, which matters when the epilogue's latency is close to the main loop's (lin1's GELU).
      constexpr bool kEarlyRelease = ULLSAM_GEMM_EARLY_RELEASE && !(EPI == EPI_LN_BF16 || EPI == EPI_LN_NCHW ||
                                                                  EPI == EPI_LN_RES || EPI == EPI_UP1 || EPI == EPI_MASK);
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) {
#if ULLSAM_GEMM_RELAXED_ARRIVE
            mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tempty_bar[acc]), 0));  // leader's barrier
#else
            mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
#endif
          } else {
            mbar_arrive(&tempty_bar[acc]);
          }
        }
      };

      if constexpr (EPI == EPI_LN_BF16 || EPI == EPI_LN_NCHW) {
        // ---- LayerNorm2d across the BN (== N) channels of this token; two warps share a row ----
        float* s_sum = scratch;             // [2][128]
        float* s_sq = scratch + 2 * 128;    // [2][128]
        float part = 0.f;
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr0 + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) part += __uint_as_float(v[j]);
        }
        s_sum[half * 128 + row_in_tile] = part;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float mean = (s_sum[row_in_tile] + s_sum[128 + row_in_tile]) * (1.0f / BN);
        part = 0.f;
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr0 + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = __uint_as_float(v[j]) - mean;
            part += d * d;
          }
        }
        s_sq[half * 128 + row_in_tile] = part;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float var = (s_sq[row_in_tile] + s_sq[128 + row_in_tile]) * (1.0f / BN);
        const float rstd = 1.0f / sqrtf(var + p.ln_eps);
#pragma unroll 1
        for (int c = 0; c < NCHUNK; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr0 + c * 32, v);
          tmem_wait_ld();
          const int n0 = col0 + c * 32;
          if constexpr (EPI == EPI_LN_BF16) {
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = (__uint_as_float(v[2 * j]) - mean) * rstd * __ldg(p.ln_w + n0 + 2 * j) + __ldg(p.ln_b + n0 + 2 * j);
              const float b = (__uint_as_float(v[2 * j + 1]) - mean) * rstd * __ldg(p.ln_w + n0 + 2 * j + 1) + __ldg(p.ln_b + n0 + 2 * j + 1);
              o[j] = pack_bf16(a, b);
            }
            if (row_ok) {
              uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldc + n0);
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
          } else {
            // NCHW: out[b][n][pix]; consecutive lanes hold consecutive pixels -> coalesced per channel
            if (row_ok) {
              const int b = row / p.hw, pix = row % p.hw;
              float* dst = reinterpret_cast<float*>(p.out) + (static_cast<size_t>(b) * p.N + n0) * p.hw + pix;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                dst[static_cast<size_t>(j) * p.hw] =
                    (__uint_as_float(v[j]) - mean) * rstd * __ldg(p.ln_w + n0 + j) + __ldg(p.ln_b + n0 + j);
              }
            }
          }
        }
      } else if constexpr (EPI == EPI_LN_RES) {
        // ---- x = acc + bias + residual, LayerNorm over the 256 channels (two warps share a row) ----
        float* s_sum = scratch;
        float* s_sq = scratch + 2 * 128;
        const size_t rbase = static_cast<size_t>(row_ok ? row : 0) * p.ldc;
        auto load_x = [&](int c, float (&f)[32]) {
          uint32_t (&v)[32] = reinterpret_cast<uint32_t (&)[32]>(f);   // accumulator bits land in place
          tmem_ld32(taddr0 + c * 32, v);
          tmem_wait_ld();
          const int n0 = col0 + c * 32;
          const float4* res = reinterpret_cast<const float4*>(p.residual + rbase + n0);
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r = res[j];
            const float4 b = __ldg(b4 + j);
            f[4 * j] = __uint_as_float(v[4 * j]) + r.x + b.x;
            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + r.y + b.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + r.z + b.z;
            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + r.w + b.w;
          }
        };
        // One pass for the statistics (sum and sum of squares together, one exchange with the other column half) and
        // one for the output; the first KEEP chunks of the row stay in registers in between, the others are read again
        // (TMEM + residual) — instead of three full passes.
        constexpr int KEEP = NCHUNK < 2 ? NCHUNK : 2;
        float xk[KEEP][32];
        float s1 = 0.f, s2 = 0.f;
        auto stats = [&](const float (&f)[32]) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            s1 += f[j];
            s2 = fmaf(f[j], f[j], s2);
          }
        };
#pragma unroll
        for (int c = 0; c < KEEP; ++c) {
          load_x(c, xk[c]);
          stats(xk[c]);
        }
#pragma unroll 1
        for (int c = KEEP; c < NCHUNK; ++c) {
          float f[32];
          load_x(c, f);
          stats(f);
        }
        s_sum[half * 128 + row_in_tile] = s1;
        s_sq[half * 128 + row_in_tile] = s2;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float mean = (s_sum[row_in_tile] + s_sum[128 + row_in_tile]) * (1.0f / BN);
        const float var = fmaxf((s_sq[row_in_tile] + s_sq[128 + row_in_tile]) * (1.0f / BN) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.ln_eps);
        auto emit = [&](float (&f)[32], int c) {
          const int n0 = col0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = (f[j] - mean) * rstd * __ldg(p.ln_w + n0 + j) + __ldg(p.ln_b + n0 + j);
          if (row_ok) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + rbase + n0);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            uint4* d2 = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + rbase + n0);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              d2[j] = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                 pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
          }
        };
#pragma unroll 1
        for (int c = KEEP; c < NCHUNK; ++c) {   // (these read the residual again: before the kept chunks overwrite it?
          float f[32];                          //  no — every chunk only touches its own columns)
          load_x(c, f);
          emit(f, c);
        }
#pragma unroll
        for (int c = 0; c < KEEP; ++c) emit(xk[c], c);
        // (the scratch is written again only after the next tile's statistics pass, which every warp of the row pair
        //  reaches after this barrier: add one more so a fast warp cannot overwrite s_sum before the slow one has read it)
        asm volatile("bar.sync 1, 256;" ::: "memory");
      } else if constexpr (EPI == EPI_UP1) {
        // ---- columns = (ky,kx,co): each 64-wide group is one output pixel; LayerNorm2d(64) + GELU ----
        static_assert(EPI != EPI_UP1 || BN == 256, "EPI_UP1 needs BN == 256");
        const int pr = row / 4096, pix = row % 4096;
        const int y = pix >> 6, x = pix & 63;
#pragma unroll 1
        for (int gg = 0; gg < 2; ++gg) {
          const int grp = half * 2 + gg;  // ky*2 + kx
          float f[64];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(taddr0 + gg * 64 + c * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) f[c * 32 + j] = __uint_as_float(v[j]) + __ldg(p.bias + grp * 64 + c * 32 + j);
          }
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 64; ++j) s += f[j];
          const float mean = s * (1.0f / 64);
          float sq = 0.f;
#pragma unroll
          for (int j = 0; j < 64; ++j) sq += (f[j] - mean) * (f[j] - mean);
          const float rstd = 1.0f / sqrtf(sq * (1.0f / 64) + p.ln_eps);
          if (row_ok) {
            const int oy = 2 * y + (grp >> 1), ox = 2 * x + (grp & 1);
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) +
                                                  ((static_cast<size_t>(pr) * 128 + oy) * 128 + ox) * 64);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float o[8];
#pragma unroll
              for (int u = 0; u < 8; ++u)
                o[u] = gelu_erf((f[8 * j + u] - mean) * rstd * __ldg(p.ln_w + 8 * j + u) + __ldg(p.ln_b + 8 * j + u));
              dst[j] = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
            }
          }
        }
      } else if constexpr (EPI == EPI_MASK) {
        // ---- rows = pixels of the 128x128 map; columns = (ky,kx,c2): GELU then contract c2 with hyper_in ----
        static_assert(EPI != EPI_MASK || BN == 128, "EPI_MASK needs BN == 128");
        const int pr = row / 16384, pix = row % 16384;
        const int y = pix >> 7, x = pix & 127;
        const float* hyp = p.hyper + static_cast<size_t>(row_ok ? pr : 0) * 4 * 32;
#pragma unroll 1
        for (int gg = 0; gg < 2; ++gg) {
          const int grp = half * 2 + gg;
          uint32_t v[32];
          tmem_ld32(taddr0 + gg * 32, v);
          tmem_wait_ld();
          float u[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) u[j] = gelu_erf(__uint_as_float(v[j]) + __ldg(p.bias + grp * 32 + j));
          if (row_ok) {
            const int oy = 2 * y + (grp >> 1), ox = 2 * x + (grp & 1);
            for (int m = 0; m < p.mask_cnt; ++m) {
              const float* hm = hyp + (p.mask_m0 + m) * 32;
              float acc = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) acc = fmaf(u[j], __ldg(hm + j), acc);
              reinterpret_cast<float*>(p.out)[((static_cast<size_t>(pr) * p.mask_cnt + m) * 256 + oy) * 256 + ox] = acc;
            }
          }
        }
      } else {
        // ---- generic epilogues.  The accumulator arrives row-per-thread (TMEM lane == row); it is transposed
        // through a warp-private shared-memory tile so that every global access is a run of full 32-byte sectors
        // (4 lanes cover 64 contiguous bytes of one row, 8 rows per instruction) instead of 32 scattered rows. ----
        uint8_t* stage = smem + L::kStageOffset + (warp - GEMM_EPI_WARP0) * L::kStageBytesPerWarp;
        const int tile_row0 = m_blk * GEMM_BM + quad * 32;
#if ULLSAM_GEMM_BIAS_SMEM
        if (p.bias != nullptr) {
          __syncwarp();  // every lane is done with the previous tile's slice
          if (bias_lane) {
            *reinterpret_cast<float4*>(bias_s + lane * 4) = bias_next;
            if constexpr (kFold) *reinterpret_cast<float4*>(folds_s + lane * 4) = folds_next;
            const int tn = tile_at(it + 1);
            if (tn < num_tiles) {
              bias_next = __ldg(reinterpret_cast<const float4*>(p.bias + (tn % tiles_n) * BN + half * COLS_PER_WARP + lane * 4));
              if constexpr (kFold)
                folds_next = __ldg(reinterpret_cast<const float4*>(p.fold_s + (tn % tiles_n) * BN + half * COLS_PER_WARP + lane * 4));
            }
          }
          __syncwarp();
        }
#endif
        const int rr = lane >> 2, cc = lane & 3;  // transposed phase: row-in-group, 16-byte column slot
        float ksq = 0.f;                          // EPI_QKV with kmax_sq: |k row|^2 of the first chunk of a head
        size_t qkv_row_off[4] = {0, 0, 0, 0};     // EPI_QKV: (g * heads * T + t) * 64 of the 4 rows this lane stores
        if constexpr (kIsQkv) {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int grow = tile_row0 + it * 8 + rr;
            int g, t;
            if (p.ln_partition) {
              // rows are image tokens (b, y, x) of the 64 x 64 grid: window_partition (image_encoder.py:243-264) happens
              // here, on the store — the GEMM never sees the 804 padding tokens per image (their k / v rows are the bias,
              // written by kv_pad_fill_kernel)
              const int b = grow >> 12, y = (grow >> 6) & 63, x = grow & 63;
              const int wy = y / 14, wx = x / 14;
              g = (b * 5 + wy) * 5 + wx;
              t = (y - wy * 14) * 14 + (x - wx * 14);
            } else {
              g = grow / p.T;
              t = grow - g * p.T;
            }
            qkv_row_off[it] = (static_cast<size_t>(g) * p.heads * p.T + t) * 64;
          }
          // The k / v rows of the window-PADDING tokens (804 per image: the normalised map is zero-padded to 70 x 70 before the
          // projection, so they equal the bias and are attended to, image_encoder.py:168-172) ride on this epilogue: every
          // 32-row unit of the image (128 of them) writes its 1/128 share of the pad tokens for the columns it owns —
          // six or seven 256-byte rows next to the unit's own 32, no extra kernel and no extra pass over memory.
          if (p.ln_partition && p.bias != nullptr) {
            constexpr int CHUNKS = COLS_PER_WARP / 8;              // 16-byte chunks per pad row in this warp's columns
            constexpr int TOK_PER_INST = 32 / CHUNKS;
            static_assert(CHUNKS == 8 || CHUNKS == 16, "pad rows: 64 or 128 columns per epilogue warp");
            const int dim = p.heads * 64;
            const int unit = (tile_row0 & 4095) >> 5, img = tile_row0 >> 12;
            const int id0 = unit * 804 / 128, id1 = (unit + 1) * 804 / 128;
            const int chunk = lane % CHUNKS;
            const int ncol = col0 + chunk * 8;
            const int which = ncol / dim, rem = ncol - which * dim;
            if (which != 0 && tile_row0 < p.M) {
              // (folded LayerNorm: p.bias holds W beta + b; the padding tokens are zero AFTER the norm, so they see b alone)
              float4 b0, b1;
              if constexpr (kFold) {
                b0 = __ldg(reinterpret_cast<const float4*>(p.pad_bias + ncol));
                b1 = __ldg(reinterpret_cast<const float4*>(p.pad_bias + ncol + 4));
              } else {
#if ULLSAM_GEMM_BIAS_SMEM
                b0 = *reinterpret_cast<const float4*>(bias_s + chunk * 8);
                b1 = *reinterpret_cast<const float4*>(bias_s + chunk * 8 + 4);
#else
                b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol));
                b1 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol + 4));
#endif
              }
              const uint4 val = make_uint4(pack_bf16(b0.x, b0.y), pack_bf16(b0.z, b0.w), pack_bf16(b1.x, b1.y), pack_bf16(b1.z, b1.w));
              __nv_bfloat16* base = (which == 1 ? p.k : p.v) + static_cast<size_t>(rem >> 6) * p.T * 64 + (rem & 63);
              for (int id = id0 + lane / CHUNKS; id < id1; id += TOK_PER_INST) {
                // pad token id -> (yy, xx) of the 70 x 70 padded map: 6 full bottom rows (420 tokens), then the right strip
                int yy, xx;
                if (id < 420) {
                  yy = 64 + id / 70;
                  xx = id % 70;
                } else {
                  yy = (id - 420) / 6;
                  xx = 64 + (id - 420) % 6;
                }
                const int wy = yy / 14, wx = xx / 14;
                const int g = (img * 5 + wy) * 5 + wx, t = (yy - wy * 14) * 14 + (xx - wx * 14);
                *reinterpret_cast<uint4*>(base + (static_cast<size_t>(g) * p.heads * p.T + t) * 64) = val;
              }
            }
          }
        }
        // The TMEM load of chunk c+1 is in flight while chunk c is processed (two register buffers, loop fully
        // unrolled).  The GELU variant keeps the rolled loop: its per-chunk body is ~600 instructions and four
        // copies of it fall out of the instruction cache (measured: 972 -> 850 TFLOP/s when unrolled).
        // (16 epilogue warps: 96 registers, no room for two buffers; with the addend double buffer neither)
        // (keeping both double buffers was measured slower: proj 222 -> 234 us)
        constexpr bool kPipe = (!kIsGelu || ULLSAM_GELU_PIPE) && EW == 8 && !kAddPipe;
        uint32_t vbuf[2][32];
        if constexpr (kPipe) {
          tmem_ld32(taddr0, vbuf[0]);
          tmem_wait_ld();
          if constexpr (kEarlyRelease && NCHUNK == 1) release_acc();
        }
#pragma unroll((kPipe || kAddPipe) ? NCHUNK : 1)
        for (int c = 0; c < NCHUNK; ++c) {
          const int n0 = col0 + c * 32;
          uint32_t (&v)[32] = vbuf[kPipe ? (c & 1) : 0];
          if constexpr (kPipe) {
            if (c + 1 < NCHUNK) tmem_ld32(taddr0 + (c + 1) * 32, vbuf[(c + 1) & 1]);
          } else {
            tmem_ld32(taddr0 + c * 32, v);
            tmem_wait_ld();
            if constexpr (kEarlyRelease) {
              if (c == NCHUNK - 1) release_acc();
            }
          }
          // The addend (residual stream / pos_embed) of the whole 32-column chunk is requested up front,
          // so 8 independent 16-byte loads per lane are in flight while the accumulator is transposed.
          float4 add[2][4];
          if constexpr (kAddPipe) {
            if (c + 1 < NCHUNK) {
              load_addend(addbuf[(c + 1) & 1], m_blk, n0 + 32);
            } else {
              const int tn = tile_at(it + 1);
              if (tn < num_tiles)
                load_addend(addbuf[(c + 1) & 1], (tn / tiles_n) * CG + cta_rank, (tn % tiles_n) * BN + half * COLS_PER_WARP);
            }
          } else if constexpr (kAdd) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int grow = tile_row0 + it * 8 + rr;
                add[h][it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (grow < p.M) {
                  if constexpr (kRes)
                    add[h][it] = *reinterpret_cast<const float4*>(p.residual + static_cast<size_t>(grow) * p.ldc + n0 + h * 16 + cc * 4);
                  else
                    add[h][it] = __ldg(reinterpret_cast<const float4*>(
                        p.pos + static_cast<size_t>(grow % p.pos_rows) * p.N + n0 + h * 16 + cc * 4));
                }
              }
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if constexpr (kFold && !(ULLSAM_FOLD_DBG & 2)) {
            // LayerNorm finished here: rstd * (acc - mean * s_n) + c_n  (gemm.h)
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + j);
              const float4 s4 = *reinterpret_cast<const float4*>(folds_s + c * 32 + j);
              f[j] = fmaf(f[j], fold_r, fmaf(fold_t, s4.x, b4.x));
              f[j + 1] = fmaf(f[j + 1], fold_r, fmaf(fold_t, s4.y, b4.y));
              f[j + 2] = fmaf(f[j + 2], fold_r, fmaf(fold_t, s4.z, b4.z));
              f[j + 3] = fmaf(f[j + 3], fold_r, fmaf(fold_t, s4.w, b4.w));
            }
          } else if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
#if ULLSAM_GEMM_BIAS_SMEM
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + j);
#else
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
#endif
              f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
          }
          if constexpr (kIsGelu) {
#if ULLSAM_GELU_LOGISTIC
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_logistic(f[j]);
#else
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = ((ULLSAM_GELU_RCP_MASK >> (j & 3)) & 1) ? gelu_erf_rcp(f[j]) : gelu_erf(f[j]);
#endif
          }
          if constexpr (EPI == EPI_BF16 || EPI == EPI_F32) {
            if (p.pos != nullptr) {
              const float4* pos = reinterpret_cast<const float4*>(p.pos + static_cast<size_t>(row % p.pos_rows) * p.N + n0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 r = __ldg(pos + j);
                f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
              }
            }
            if (p.act == ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
            } else if (p.act == ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
          }
          if constexpr (kIsQkv) {
            // key-norm bound for the global attention: max over the warp's 32 rows of |k row|^2, one head = two chunks
            if (p.kmax_sq != nullptr) {
              const int dim = p.heads * 64;
              if (n0 >= dim && n0 < 2 * dim) {          // (warp-uniform: a chunk lies inside one of q / k / v)
                float sq = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) sq = fmaf(f[j], f[j], sq);
                ksq = (c & 1) ? ksq + sq : sq;
                if (c & 1) {
                  const float m = warp_max(row_ok ? ksq : 0.f);
                  if (lane == 0 && tile_row0 < p.M)
                    atomicMax(reinterpret_cast<int*>(p.kmax_sq) + (tile_row0 / p.T) * p.heads + ((n0 - dim) >> 6), __float_as_int(m));
                }
              }
            }
          }
          if constexpr (EPI == EPI_RES_RED) {
            // out[32 rows][32 columns] += acc + bias, the add done at the memory side: the chunk goes to a 4 KB box in shared
            // memory (row = lane, 128-byte rows, 16-byte pieces XOR-swizzled with the row like the tensor map says: no bank
            // conflicts) and leaves as ONE TMA reduce-add store.  Two boxes per warp; a box is reused once the store issued two
            // chunks ago has read it.  Nothing here waits for global memory.
            uint8_t* rbuf = smem + L::kRedOffset + (warp - GEMM_EPI_WARP0) * L::kRedBytesPerWarp + (c & 1) * 4096;
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(rbuf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && tile_row0 < p.M) {
              tma_reduce_add_2d(&tmC, rbuf, n0, tile_row0);   // rows beyond M are clipped by the tensor map
              tma_store_commit();
            }
          } else if constexpr (EPI == EPI_BF16 || kIsGelu || kIsQkv) {
            // 32 bf16 per row = 64 bytes; staged with an 80-byte row pitch (conflict-free both ways)
            uint4* srow = reinterpret_cast<uint4*>(stage + lane * 80);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              srow[j] = make_uint4(pack_bf16(f[8 * j], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                   pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
            __syncwarp();
            size_t qkv_col_off = 0;
            __nv_bfloat16* qkv_base = nullptr;
            if constexpr (kIsQkv) {
              const int dim = p.heads * 64;
              const int which = n0 / dim;
              const int rem = n0 - which * dim;
              qkv_base = which == 0 ? p.q : (which == 1 ? p.k : p.v);
              qkv_col_off = static_cast<size_t>(rem >> 6) * p.T * 64 + (rem & 63);  // head * T * 64 + d0
            }
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int r_in = it * 8 + rr;
              const int grow = tile_row0 + r_in;
              const uint4 val = *reinterpret_cast<const uint4*>(stage + r_in * 80 + cc * 16);
              if (grow < p.M) {
                if constexpr (EPI == EPI_BF16 || kIsGelu) {
                  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(grow) * p.ldc + n0 + cc * 8) = val;
                } else {
                  *reinterpret_cast<uint4*>(qkv_base + qkv_row_off[it] + qkv_col_off + cc * 8) = val;
                }
              }
            }
            __syncwarp();
          } else if constexpr (kWide) {
            // fp32 outputs, full lines: rows [16 h, 16 h + 16) of the warp's 32, 128 bytes each at a 144-byte pitch
            // (conflict-free both ways); read back as 8 lanes x 16 bytes per row, 4 rows per instruction
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if ((lane >> 4) == h) {
                float4* srow = reinterpret_cast<float4*>(stage + (lane & 15) * 144);
#pragma unroll
                for (int j = 0; j < 8; ++j) srow[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
              }
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r_in = i * 4 + (lane >> 3);
                const int grow = tile_row0 + h * 16 + r_in;
                float4 val = *reinterpret_cast<const float4*>(stage + r_in * 144 + (lane & 7) * 16);
                if (grow < p.M) {
                  if constexpr (kAdd) {
                    const float4 a = addbuf[c & 1][h][i];
                    val.x += a.x; val.y += a.y; val.z += a.z; val.w += a.w;
                  }
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(grow) * p.ldc + n0 + (lane & 7) * 4) = val;
                  if constexpr (kStats) {
                    // bf16 copy of the stream (the A operand of the GEMM behind the LayerNorm) + the row's partial sums
                    if (!(ULLSAM_STATS_DBG & 1))
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out2) + static_cast<size_t>(grow) * p.ldc + n0 + (lane & 7) * 4) =
                        make_uint2(pack_bf16(val.x, val.y), pack_bf16(val.z, val.w));
                    if (!(ULLSAM_STATS_DBG & 2)) {
                    rs1[h][i] += (val.x + val.y) + (val.z + val.w);
                    rs2[h][i] = fmaf(val.x, val.x, fmaf(val.y, val.y, fmaf(val.z, val.z, fmaf(val.w, val.w, rs2[h][i]))));
                    }
                  }
                }
              }
              __syncwarp();
            }
          } else {
            // fp32 outputs: two 16-column halves, 64 bytes per row each, same 80-byte pitch
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float4* srow = reinterpret_cast<float4*>(stage + lane * 80);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                srow[j] = make_float4(f[h * 16 + 4 * j], f[h * 16 + 4 * j + 1], f[h * 16 + 4 * j + 2], f[h * 16 + 4 * j + 3]);
              __syncwarp();
#pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int r_in = it * 8 + rr;
                const int grow = tile_row0 + r_in;
                float4 val = *reinterpret_cast<const float4*>(stage + r_in * 80 + cc * 16);
                if (grow < p.M) {
                  const size_t off = static_cast<size_t>(grow) * p.ldc + n0 + h * 16 + cc * 4;
                  if constexpr (kRes || kIsPatch) {
                    val.x += add[h][it].x; val.y += add[h][it].y; val.z += add[h][it].z; val.w += add[h][it].w;
                  }
                  if constexpr (EPI == EPI_RES_LN) {
                    st1[it] += (val.x + val.y) + (val.z + val.w);
                    st2[it] = fmaf(val.x, val.x, fmaf(val.y, val.y, fmaf(val.z, val.z, fmaf(val.w, val.w, st2[it]))));
                  }
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + off) = val;
                }
              }
              __syncwarp();
            }
          }
          if constexpr (kPipe) {
            if (c + 1 < NCHUNK) tmem_wait_ld();  // chunk c+1 has landed (it was in flight during the work above)
            if constexpr (kEarlyRelease) {
              if (c + 2 == NCHUNK) release_acc();
            }
          }
        }
      }
      if constexpr (!kEarlyRelease) release_acc();
      if constexpr (kStats && !(ULLSAM_STATS_DBG & 2)) {
        // the 8 lanes that share a row (lane & 7 = 16-byte column slot) add up their partials; lane & 7 == 0 stores the pair
        float2* dst = reinterpret_cast<float2*>(p.row_stats);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a = rs1[h][i], b = rs2[h][i];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              a += __shfl_xor_sync(0xffffffffu, a, o);
              b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            const int grow = m_blk * GEMM_BM + quad * 32 + h * 16 + i * 4 + (lane >> 3);
            if ((lane & 7) == 0 && grow < p.M) dst[static_cast<size_t>(grow) * kStatSlots + n_blk * 2 + half] = make_float2(a, b);
            rs1[h][i] = 0.f;
            rs2[h][i] = 0.f;
          }
      }
      if constexpr (EPI == EPI_RES_LN) {
        // ---- the LayerNorm that follows the residual add, once the last N tile of this row block is stored ----
        if (n_blk == tiles_n - 1) {
          const int rr = lane >> 2, cc = lane & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // the four 16-byte column slots of a row live in four adjacent lanes
            st1[i] += __shfl_xor_sync(0xffffffffu, st1[i], 1);
            st2[i] += __shfl_xor_sync(0xffffffffu, st2[i], 1);
            st1[i] += __shfl_xor_sync(0xffffffffu, st1[i], 2);
            st2[i] += __shfl_xor_sync(0xffffffffu, st2[i], 2);
          }
          float* s1 = scratch;        // [2 column halves][128 rows]
          float* s2 = scratch + 256;
          if (cc == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              s1[half * 128 + quad * 32 + i * 8 + rr] = st1[i];
              s2[half * 128 + quad * 32 + i * 8 + rr] = st2[i];
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps (stats of the other column half)
          float mean[4], rstd[4];
          const float inv_n = 1.0f / static_cast<float>(p.N);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = quad * 32 + i * 8 + rr;
            mean[i] = (s1[r] + s1[128 + r]) * inv_n;
            const float var = fmaxf((s2[r] + s2[128 + r]) * inv_n - mean[i] * mean[i], 0.f);
            rstd[i] = rsqrtf(var + p.ln_eps);
            st1[i] = 0.f;
            st2[i] = 0.f;
          }
          __syncwarp();  // every global store of this warp to its 32 rows (all N tiles) is visible to all its lanes
          // second pass: re-read (L2), normalise, store bf16
          const int row0 = m_blk * GEMM_BM + quad * 32;
          if (tiles_n == 3) res_ln_second_pass<3>(p, row0, half, lane, mean, rstd);
          else if (tiles_n == 4) res_ln_second_pass<4>(p, row0, half, lane, mean, rstd);
          else if (tiles_n == 2) res_ln_second_pass<2>(p, row0, half, lane, mean, rstd);
          else res_ln_second_pass<1>(p, row0, half, lane, mean, rstd);
        }
      }
    }
  }

  if constexpr (EPI == EPI_RES_RED) {
    if (warp >= GEMM_EPI_WARP0 && lane == 0) tma_store_wait_all();   // every reduce-add of this warp has been performed
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync();  // the peer's shared memory / barriers stay alive until both are done
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_2cta(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ullsam
