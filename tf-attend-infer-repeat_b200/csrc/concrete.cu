// Fused Concrete / Gumbel-softmax z_pres step with its MC KL term and the ACT-style
// running-sum stopping bookkeeping (one elementwise kernel, externally supplied noise).
//
// Replaces, per loop iteration of the reference:
//   air/concrete.py:20-27   concrete_binary_pre_sigmoid_sample
//   air/concrete.py:30-43   concrete_binary_kl_mc_sample
//   air/air_model.py:385-390  sigmoid / round
//   air/air_model.py:395      z_pres_prob
//   air/air_model.py:411-415  running_loss += where(stop_prev < thr, kl, 0)
//   air/air_model.py:424-427  stopping_sum, running_digits
// Arithmetic follows the reference op by op (no FMA contraction); logf/expf are the CUDA
// libm versions (<= 1-2 ulp from the CPU libm), so floats agree to ~1e-6 relative and the
// integer outputs (digit counts, masks) are exact away from knife-edge thresholds.
#include <algorithm>

#include "air_common.cuh"

namespace air {

constexpr float kEps = 1e-9f;  // the reference's 10e-10

// log(tau+eps) - y*tau + alpha - 2*log(1 + exp(-y*tau + alpha) + eps)      (concrete.py:35-37)
__device__ __forceinline__ float log_density(float y, float alpha, float tau) {
  const float yt = mul_rn(y, tau);
  const float e = expf(add_rn(-yt, alpha));
  const float l = logf(add_rn(add_rn(1.0f, e), kEps));
  return sub_rn(add_rn(sub_rn(logf(add_rn(tau, kEps)), yt), alpha), mul_rn(2.0f, l));
}

__device__ __forceinline__ float sigmoid_tf(float x) { return __fdiv_rn(1.0f, add_rn(1.0f, expf(-x))); }

__global__ void __launch_bounds__(256)
    concrete_step_fwd(const float *__restrict__ log_odds, const float *__restrict__ u, const float *stop_prev,
                      const float *loss_prev, const int32_t *digits_prev, const float *__restrict__ prior_log_odds,
                      float tau, float thr, int train, float *__restrict__ y, float *__restrict__ z,
                      float *__restrict__ z_prob, float *__restrict__ kl, float *stop_new, float *loss_new,
                      int32_t *digits_new, int64_t B) {
  const float prior = __ldg(prior_log_odds);
  for (int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; b < B;
       b += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float lo = log_odds[b], uu = u[b];
    const float noise = sub_rn(logf(add_rn(uu, kEps)), logf(add_rn(sub_rn(1.0f, uu), kEps)));
    const float yy = __fdiv_rn(add_rn(lo, noise), tau);
    float zz = sigmoid_tf(yy);
    if (!train) zz = rintf(zz);  // tf.round: half to even
    const float k = sub_rn(log_density(yy, lo, tau), log_density(yy, prior, tau));
    const float sp = stop_prev[b];
    const float sn = add_rn(sp, sub_rn(1.0f, zz));
    const float ln = add_rn(loss_prev[b], sp < thr ? k : 0.0f);
    const int32_t dn = digits_prev[b] + (sn < thr ? 1 : 0);
    y[b] = yy;
    z[b] = zz;
    z_prob[b] = sigmoid_tf(lo);
    kl[b] = k;
    stop_new[b] = sn;
    loss_new[b] = ln;
    digits_new[b] = dn;
  }
}

// d/dy and d/dalpha of log_density (ignoring the eps inside the log, as autodiff would
// not: d log(1+e+eps) = e/(1+e+eps)) -- kept exactly as autodiff gives it.
__global__ void __launch_bounds__(256)
    concrete_step_bwd(const float *__restrict__ log_odds, const float *__restrict__ y, const float *__restrict__ z,
                      const float *__restrict__ dz, const float *__restrict__ dkl,
                      const float *__restrict__ prior_log_odds, float tau, int train, float *__restrict__ dlog_odds,
                      int64_t B) {
  const float prior = __ldg(prior_log_odds);
  for (int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; b < B;
       b += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float lo = log_odds[b], yy = y[b], gk = dkl ? dkl[b] : 0.0f;
    const float yt = yy * tau;
    // posterior: alpha = lo ;  f = -yt + alpha - 2 log(1 + exp(-yt+alpha) + eps)
    const float eq = expf(-yt + lo), sq = eq / (1.0f + eq + kEps);
    const float ep = expf(-yt + prior), sp = ep / (1.0f + ep + kEps);
    // d kl / d y      = tau * [(-1 + 2 sq) - (-1 + 2 sp)]
    // d kl / d lo|y   = 1 - 2 sq
    const float dkl_dy = tau * (2.0f * sq - 2.0f * sp);
    const float dkl_dlo = 1.0f - 2.0f * sq;
    float gy = gk * dkl_dy;
    if (train && dz) {
      const float zz = z[b];
      gy += dz[b] * zz * (1.0f - zz);  // sigmoid'
    }
    // y = (lo + noise)/tau
    dlog_odds[b] = gy / tau + gk * dkl_dlo;
  }
}

}  // namespace air

extern "C" int air_concrete_step_fwd(const float *log_odds, const float *u, const float *stop_prev,
                                     const float *loss_prev, const int32_t *digits_prev, const float *prior_log_odds,
                                     float temperature, float thr, int train, float *y, float *z, float *z_prob,
                                     float *kl, float *stop_new, float *loss_new, int32_t *digits_new, int64_t B,
                                     air_stream_t stream) {
  AIR_REQUIRE(log_odds && u && stop_prev && loss_prev && digits_prev && prior_log_odds && y && z && z_prob && kl &&
                  stop_new && loss_new && digits_new,
              AIR_ERR_NULL, "concrete_step_fwd: null pointer");
  AIR_REQUIRE(B >= 0, AIR_ERR_BAD_SHAPE, "concrete_step_fwd: B < 0");
  if (B == 0) return AIR_OK;
  const int blocks = static_cast<int>(std::min<int64_t>((B + 255) / 256, static_cast<int64_t>(air::sm_count()) * 8));
  air::concrete_step_fwd<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      log_odds, u, stop_prev, loss_prev, digits_prev, prior_log_odds, temperature, thr, train, y, z, z_prob, kl,
      stop_new, loss_new, digits_new, B);
  air::count_launch();
  return air::check_launch("concrete_step_fwd");
}

extern "C" int air_concrete_step_bwd(const float *log_odds, const float *y, const float *z, const float *dz,
                                     const float *dkl, const float *prior_log_odds, float temperature, int train,
                                     float *dlog_odds, int64_t B, air_stream_t stream) {
  AIR_REQUIRE(log_odds && y && z && prior_log_odds && dlog_odds, AIR_ERR_NULL, "concrete_step_bwd: null pointer");
  AIR_REQUIRE(B >= 0, AIR_ERR_BAD_SHAPE, "concrete_step_bwd: B < 0");
  if (B == 0) return AIR_OK;
  const int blocks = static_cast<int>(std::min<int64_t>((B + 255) / 256, static_cast<int64_t>(air::sm_count()) * 8));
  air::concrete_step_bwd<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      log_odds, y, z, dz, dkl, prior_log_odds, temperature, train, dlog_odds, B);
  air::count_launch();
  return air::check_launch("concrete_step_bwd");
}
