// Epilogue functions shared by the GEMM kernels and the elementwise kernels.
#pragma once
#include "air_common.cuh"

namespace air {

// tf.nn.softplus (Eigen functor): thresholds at +-(log(eps_f32) + 2) = -+13.9424
__device__ __forceinline__ float softplus_tf(float x) {
  const float thr = -13.942385f;  // logf(FLT_EPSILON) + 2
  if (x > -thr) return x;
  const float e = expf(x);
  if (x < thr) return e;
  return logf(add_rn(e, 1.0f));
}

__device__ __forceinline__ float sigmoid_f(float x) { return __fdiv_rn(1.0f, add_rn(1.0f, expf(-x))); }

__device__ __forceinline__ float apply_epilogue(float v, int epi, float aux) {
  switch (epi) {
    case AIR_EPI_RELU: return fmaxf(v, 0.0f);
    case AIR_EPI_SOFTPLUS: return softplus_tf(v);
    case AIR_EPI_MUL_DRELU: return aux > 0.0f ? v : 0.0f;
    case AIR_EPI_MUL_DSOFTPLUS: return v * (-expm1f(-aux));  // sigmoid(x) == 1 - exp(-softplus(x)), accurate for tiny aux
    default: return v;
  }
}

}  // namespace air
