// Epilogue functions shared by the GEMM kernels and the elementwise kernels.
#pragma once
#include "air_common.cuh"
#include "rng.cuh"

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

__device__ __forceinline__ float apply_epilogue(float v, int epi, float aux, float param = 0.0f) {
  switch (epi) {
    case AIR_EPI_SIGMOID_RNG:  // (the caller passes the generated N(0,1) sample as aux: epi_aux_value)
    case AIR_EPI_SIGMOID_NOISE: return sigmoid_f(add_rn(v, mul_rn(aux, param)));
    case AIR_EPI_RELU: return fmaxf(v, 0.0f);
    case AIR_EPI_SOFTPLUS: return softplus_tf(v);
    case AIR_EPI_MUL_DRELU: return aux > 0.0f ? v : 0.0f;
    case AIR_EPI_MUL_DSOFTPLUS: return v * (-expm1f(-aux));  // sigmoid(x) == 1 - exp(-softplus(x)), accurate for tiny aux
    default: return v;
  }
}

// The second operand of an epilogue at C offset `o` (row * ldc + col) / logical element `idx` (row * N + col): an
// element of the aux array, or -- AIR_EPI_SIGMOID_RNG, where `aux` points at the device RNG state {seed, counter} --
// the N(0,1) sample of stream kRngLike for that element, generated on the spot.
__device__ __forceinline__ float epi_aux_value(const float *aux, int epi, int64_t o, int64_t idx) {
  if (epi == AIR_EPI_SIGMOID_RNG) {
    const unsigned long long *st = reinterpret_cast<const unsigned long long *>(aux);
    return rng_normal_elem(st[0], st[1], kRngLike, static_cast<unsigned long long>(idx));
  }
  return aux ? aux[o] : 0.0f;
}

// 16 independent element chains with the activation selected OUTSIDE the loop (ILP for the
// tensor-core epilogue; branch-free softplus so the compiler can interleave the exp/log chains)
__device__ __forceinline__ float softplus_tf_branchless(float x) {
  const float thr = -13.942385f;
  const float e = expf(x);
  const float mid = logf(add_rn(e, 1.0f));
  return x > -thr ? x : (x < thr ? e : mid);
}

__device__ __forceinline__ void apply_epilogue16(float (&v)[16], const float (&ax)[16], int epi, float param) {
  switch (epi) {
    case AIR_EPI_SIGMOID_RNG:
    case AIR_EPI_SIGMOID_NOISE:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = sigmoid_f(add_rn(v[j], mul_rn(ax[j], param)));
      break;
    case AIR_EPI_RELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
      break;
    case AIR_EPI_SOFTPLUS:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = softplus_tf_branchless(v[j]);
      break;
    case AIR_EPI_MUL_DRELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = ax[j] > 0.0f ? v[j] : 0.0f;
      break;
    case AIR_EPI_MUL_DSOFTPLUS:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = v[j] * (-expm1f(-ax[j]));
      break;
    default:
      break;
  }
}

}  // namespace air
