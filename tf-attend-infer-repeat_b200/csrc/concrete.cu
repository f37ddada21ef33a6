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
#include "concrete.cuh"

namespace air {

__global__ void __launch_bounds__(256)
    concrete_step_fwd(const float *__restrict__ log_odds, const float *__restrict__ u, const float *stop_prev,
                      const float *loss_prev, const int32_t *digits_prev, const float *__restrict__ prior_log_odds,
                      float tau, float thr, int train, float *__restrict__ y, float *__restrict__ z,
                      float *__restrict__ z_prob, float *__restrict__ kl, float *stop_new, float *loss_new,
                      int32_t *digits_new, int64_t B) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  const float prior = __ldg(prior_log_odds);
  for (int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; b < B;
       b += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const ConcreteOut o = concrete_step_one(log_odds[b], u[b], stop_prev[b], loss_prev[b], prior, tau, thr, train);
    const int32_t dn = digits_prev[b] + o.digit_inc;
    y[b] = o.y;
    z[b] = o.z;
    z_prob[b] = o.z_prob;
    kl[b] = o.kl;
    stop_new[b] = o.stop_new;
    loss_new[b] = o.loss_new;
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
  pdl_sync();  // PDL: no global access before the previous grid has completed
  const float prior = __ldg(prior_log_odds);
  for (int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; b < B;
       b += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    dlog_odds[b] = concrete_bwd_one(log_odds[b], y[b], z[b], (train && dz) ? dz[b] : 0.0f, dkl ? dkl[b] : 0.0f, prior,
                                    tau, train);
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
  AIR_LAUNCH(air::concrete_step_fwd, blocks, 256, 0, static_cast<cudaStream_t>(stream), 
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
  AIR_LAUNCH(air::concrete_step_bwd, blocks, 256, 0, static_cast<cudaStream_t>(stream), 
      log_odds, y, z, dz, dkl, prior_log_odds, temperature, train, dlog_odds, B);
  air::count_launch();
  return air::check_launch("concrete_step_bwd");
}
