// prefill.cu — prompt (prefill) path.  Placeholder until the tcgen05 GEMM lands in this file's
// siblings (gemm_tcgen05.cu / prefill_kernels.cu): prompts are processed token by token through
// the decode kernels, which is exact but HBM-bound.
#include "engine.h"

namespace cl {
bool Engine::prefill_path_ok() const { return false; }
int Engine::prefill_chunked(cl_seq_t, const int32_t*, int, float*) { return CL_ERR_INTERNAL; }
}  // namespace cl

extern "C" {
int cl_op_gemm_bf16(int, const uint16_t*, const uint16_t*, float*, int32_t, int32_t, int32_t, int32_t, float*) {
  cl::set_last_error("tcgen05 GEMM not built yet");
  return CL_ERR_INTERNAL;
}
int cl_op_attn_prefill(int, const uint16_t*, const uint16_t*, const uint16_t*, int32_t, int32_t, int32_t, int32_t, float*) {
  cl::set_last_error("prefill attention not built yet");
  return CL_ERR_INTERNAL;
}
}
