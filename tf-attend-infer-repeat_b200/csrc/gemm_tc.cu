// tcgen05 / TMEM / TMA TF32 GEMM (throughput mode of air_gemm).  Placeholder until the
// tensor-core kernel lands: reports AIR_ERR_UNSUPPORTED rather than silently falling back.
#include "air_common.cuh"

namespace air {
int gemm_tf32(const float *, const float *, float *, const float *, const float *, const float *, int, int, int, int,
              int, int, int, int, int, cudaStream_t) {
  set_error("air_gemm: AIR_GEMM_TF32 (tcgen05) is not built yet; use AIR_GEMM_FP32_EXACT");
  return AIR_ERR_UNSUPPORTED;
}
}  // namespace air
