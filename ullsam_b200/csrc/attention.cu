#include "attention.h"
namespace ullsam {
const char* launch_attention(const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, const float*,
                             const float*, int, int, int, __nv_bfloat16*, cudaStream_t) {
  return "attention: not implemented yet";
}
}  // namespace ullsam
