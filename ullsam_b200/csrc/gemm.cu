// ullsam_b200 — host side of the tcgen05 GEMM: TMA tensor-map construction and launch.
#include "gemm.h"
#include "host_util.h"

#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "gemm_sm100.cuh"

namespace ullsam {

static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static const char* load_encode() {
  static std::once_flag once;
  static const char* err = nullptr;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      err = "cuTensorMapEncodeTiled not available from the CUDA driver";
      return;
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  });
  return err;
}

const char* make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                              uint32_t box_cols) {
  if (const char* e = load_encode()) return e;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled(2d) failed";
}

const char* make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                             uint32_t box_cols) {
  if (const char* e = load_encode()) return e;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled(2d fp32) failed";
}

const char* make_tmap_nhwc_bf16(CUtensorMap* out, const void* base, int B, int H, int W, int C, uint32_t box_c,
                                uint32_t box_w, uint32_t box_h) {
  if (const char* e = load_encode()) return e;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? nullptr : "cuTensorMapEncodeTiled(4d) failed";
}

// CTA-pair variant: cluster (2,1,1), one pair per 256 x 256 tile
template <int EPI, int EW>
static const char* launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                               cudaStream_t stream, const CUtensorMap* tc = nullptr) {
  using L = GemmSmem<256, 2, EW, EPI == EPI_RES_RED>;
  static PerDeviceOnce configured;
  auto kern = gemm_kernel<256, EPI, 2, EW>;
  if (const char* ce = set_max_smem_once(configured, kern, L::kTotal)) return ce;
  const int tiles = ((p.M + 255) / 256) * (p.N / 256);
  const int pairs = tiles < num_sms / 2 ? tiles : num_sms / 2;
  cudaError_t e = launch_pdl(kern, dim3(2 * pairs), dim3((GEMM_EPI_WARP0 + EW) * 32), L::kTotal, stream, 2, ta, tb, tc ? *tc : ta, p);
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

template <int BN, int EPI>
static const char* launch_t(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                            cudaStream_t stream) {
  using L = GemmSmem<BN>;
  static PerDeviceOnce configured;
  auto kern = gemm_kernel<BN, EPI>;
  if (const char* ce = set_max_smem_once(configured, kern, L::kTotal)) return ce;
  const int tiles = ((p.M + GEMM_BM - 1) / GEMM_BM) * (p.N / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  cudaError_t e = launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), L::kTotal, stream, 1, ta, tb, ta, p);
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

const char* launch_gemm(int epi, const void* A, const void* W, const GemmParams& p, int num_sms, cudaStream_t stream) {
  if (p.M <= 0) return nullptr;
  if (epi == EPI_BF16 && p.act == ACT_GELU && p.pos == nullptr) epi = EPI_GELU_BF16;
  else if (epi == EPI_GELU_BF16) return "gemm: EPI_GELU_BF16 is selected internally";
  const bool stats = (epi == EPI_RES_STATS || epi == EPI_PATCH_STATS), fold = (epi == EPI_QKV_FOLD || epi == EPI_GELU_FOLD);
  if (stats && (p.N % 256 != 0 || p.N > 128 * kStatSlots || !p.out2 || !p.row_stats || p.a_conv3))
    return "gemm: the statistics epilogues need N % 256 == 0, N <= 1024, a plain A matrix, out2 and row_stats";
  if (fold && (p.N % 256 != 0 || p.K % 256 != 0 || p.K > 128 * kStatSlots || !p.bias || !p.fold_s || !p.row_stats || p.a_conv3))
    return "gemm: the folded-LayerNorm epilogues need N % 256 == 0, K % 256 == 0, K <= 1024, bias, fold_s and row_stats";
  if (epi == EPI_QKV_FOLD && p.ln_partition && !p.pad_bias) return "gemm: EPI_QKV_FOLD with window scatter needs pad_bias";
  if (epi == EPI_RES_STATS && !p.residual) return "gemm: EPI_RES_STATS needs residual";
  if (epi == EPI_RES_RED && (p.N % 256 != 0 || p.a_conv3 || p.M < 4096 || p.residual != p.out || p.ldc != p.N || !p.out))
    return "gemm: EPI_RES_RED needs N % 256 == 0, M >= 4096, a plain A matrix and out == residual (in place, ldc == N)";
  if (p.kmax_sq != nullptr && (p.T % 32 != 0 || p.ln_partition || p.heads % 2 != 0))
    return "gemm: kmax_sq needs T % 32 == 0, rows in group order and an even number of heads";
  if (p.K % GEMM_BK != 0) return "gemm: K must be a multiple of 64";
  const bool ln = (epi == EPI_LN_BF16 || epi == EPI_LN_NCHW || epi == EPI_LN_RES || epi == EPI_UP1);
  int bn = (p.N % 256 == 0) ? 256 : 128;
  if (p.N % bn != 0) return "gemm: N must be a multiple of 128";
  if (ln && p.N != 256) return "gemm: this epilogue needs N == 256";
  CUtensorMap ta, tb;
  const char* e;
  if (p.a_conv3) {
    if (p.M % (64 * 64) != 0) return "gemm: conv3x3 mode needs whole 64x64 images";
    e = make_tmap_nhwc_bf16(&ta, A, p.M / 4096, 64, 64, p.conv_c, 64, 64, 2);
  } else {
    e = make_tmap_2d_bf16(&ta, A, (uint64_t)p.M, (uint64_t)p.K, GEMM_BM, GEMM_BK);
  }
  if (e) return e;
  // Large plain GEMMs with 256-wide N tiles run on CTA pairs (each CTA loads only half of the W tile)
  static const bool pair_disabled = getenv("ULLSAM_GEMM_NO_PAIR") != nullptr;
  if (epi == EPI_RES_LN) {
    if (bn != 256 || p.a_conv3 || p.M < 4096 || p.N > 1024)
      return "gemm: EPI_RES_LN needs N % 256 == 0, N <= 1024, a plain A matrix and M >= 4096";
    if (!p.out2 || !p.ln_w || !p.ln_b || !p.residual) return "gemm: EPI_RES_LN needs residual, ln_w, ln_b and out2";
    if (p.ln_partition && p.M % 4096 != 0) return "gemm: EPI_RES_LN window partition needs whole 64x64 token grids";
  }
  if ((epi == EPI_QKV || epi == EPI_QKV_FOLD) && p.ln_partition && (p.M % 4096 != 0 || p.T != 196))
    return "gemm: EPI_QKV window scatter needs whole 64x64 token grids and T == 196";
  if (epi == EPI_RES_RED) {
    CUtensorMap tc;
    if ((e = make_tmap_2d_bf16(&tb, W, (uint64_t)p.N, (uint64_t)p.K, 128, GEMM_BK))) return e;
    if ((e = make_tmap_2d_f32(&tc, p.out, (uint64_t)p.M, (uint64_t)p.N, 32, 32))) return e;
    return launch_pair<EPI_RES_RED, 8>(ta, tb, p, num_sms, stream, &tc);
  }
  const bool pair = (epi == EPI_RES_LN) ||
                    (!pair_disabled && bn == 256 && !p.a_conv3 && p.M >= 4096 &&
                     (epi == EPI_BF16 || epi == EPI_GELU_BF16 || epi == EPI_RES_F32 || epi == EPI_QKV || epi == EPI_RES_STATS ||
                      epi == EPI_QKV_FOLD || epi == EPI_GELU_FOLD || epi == EPI_PATCH || epi == EPI_PATCH_STATS));
  e = make_tmap_2d_bf16(&tb, W, (uint64_t)p.N, (uint64_t)p.K, pair ? 128 : bn, GEMM_BK);
  if (e) return e;
  if (pair) {
    static const int ew16 = getenv("ULLSAM_GEMM_EW16") ? atoi(getenv("ULLSAM_GEMM_EW16")) : 0x1;  // bit mask per epilogue
    switch (epi) {
      case EPI_BF16: return launch_pair<EPI_BF16, 8>(ta, tb, p, num_sms, stream);
      case EPI_GELU_BF16:
        return (ew16 & 1) ? launch_pair<EPI_GELU_BF16, 16>(ta, tb, p, num_sms, stream)
                          : launch_pair<EPI_GELU_BF16, 8>(ta, tb, p, num_sms, stream);
      case EPI_RES_F32:
        return (ew16 & 2) ? launch_pair<EPI_RES_F32, 16>(ta, tb, p, num_sms, stream)
                          : launch_pair<EPI_RES_F32, 8>(ta, tb, p, num_sms, stream);
      case EPI_RES_LN: return launch_pair<EPI_RES_LN, 8>(ta, tb, p, num_sms, stream);
      case EPI_RES_STATS: return launch_pair<EPI_RES_STATS, 8>(ta, tb, p, num_sms, stream);
      case EPI_PATCH: return launch_pair<EPI_PATCH, 8>(ta, tb, p, num_sms, stream);
      case EPI_PATCH_STATS: return launch_pair<EPI_PATCH_STATS, 8>(ta, tb, p, num_sms, stream);
      case EPI_QKV_FOLD: return launch_pair<EPI_QKV_FOLD, 8>(ta, tb, p, num_sms, stream);
      case EPI_GELU_FOLD:
        return (ew16 & 1) ? launch_pair<EPI_GELU_FOLD, 16>(ta, tb, p, num_sms, stream)
                          : launch_pair<EPI_GELU_FOLD, 8>(ta, tb, p, num_sms, stream);
      default:
        return (ew16 & 4) ? launch_pair<EPI_QKV, 16>(ta, tb, p, num_sms, stream)
                          : launch_pair<EPI_QKV, 8>(ta, tb, p, num_sms, stream);
    }
  }

#define ULLSAM_GEMM_CASE(EPI)                                                            \
  case EPI:                                                                              \
    return bn == 256 ? launch_t<256, EPI>(ta, tb, p, num_sms, stream)                    \
                     : launch_t<128, EPI>(ta, tb, p, num_sms, stream);
  switch (epi) {
    ULLSAM_GEMM_CASE(EPI_BF16)
    ULLSAM_GEMM_CASE(EPI_GELU_BF16)
    ULLSAM_GEMM_CASE(EPI_RES_F32)
    ULLSAM_GEMM_CASE(EPI_QKV)
    ULLSAM_GEMM_CASE(EPI_PATCH)
    ULLSAM_GEMM_CASE(EPI_F32)
    case EPI_RES_STATS: return launch_t<256, EPI_RES_STATS>(ta, tb, p, num_sms, stream);
    case EPI_PATCH_STATS: return launch_t<256, EPI_PATCH_STATS>(ta, tb, p, num_sms, stream);
    case EPI_QKV_FOLD: return launch_t<256, EPI_QKV_FOLD>(ta, tb, p, num_sms, stream);
    case EPI_GELU_FOLD: return launch_t<256, EPI_GELU_FOLD>(ta, tb, p, num_sms, stream);
    case EPI_LN_RES:
      return launch_t<256, EPI_LN_RES>(ta, tb, p, num_sms, stream);
    case EPI_UP1:
      return launch_t<256, EPI_UP1>(ta, tb, p, num_sms, stream);
    case EPI_MASK:
      return p.N == 128 ? launch_t<128, EPI_MASK>(ta, tb, p, num_sms, stream) : "gemm: EPI_MASK needs N == 128";
    case EPI_LN_BF16:
      return launch_t<256, EPI_LN_BF16>(ta, tb, p, num_sms, stream);
    case EPI_LN_NCHW:
      return launch_t<256, EPI_LN_NCHW>(ta, tb, p, num_sms, stream);
    default:
      return "gemm: unknown epilogue";
  }
#undef ULLSAM_GEMM_CASE
}

}  // namespace ullsam
