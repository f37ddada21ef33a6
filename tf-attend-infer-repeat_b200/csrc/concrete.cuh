// Device functions of the binary Concrete distribution shared by concrete.cu and model_ops.cu.
#pragma once
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


// One Concrete/ACT step for one batch item (concrete.py:20-43, air_model.py:385-427).
struct ConcreteOut {
  float y, z, z_prob, kl, stop_new, loss_new;
  int digit_inc;
};
__device__ __forceinline__ ConcreteOut concrete_step_one(float lo, float uu, float stop_prev, float loss_prev,
                                                         float prior, float tau, float thr, int train) {
  ConcreteOut o;
  const float noise = sub_rn(logf(add_rn(uu, kEps)), logf(add_rn(sub_rn(1.0f, uu), kEps)));
  o.y = __fdiv_rn(add_rn(lo, noise), tau);
  o.z = sigmoid_tf(o.y);
  if (!train) o.z = rintf(o.z);  // tf.round: half to even
  o.kl = sub_rn(log_density(o.y, lo, tau), log_density(o.y, prior, tau));
  o.z_prob = sigmoid_tf(lo);
  o.stop_new = add_rn(stop_prev, sub_rn(1.0f, o.z));
  o.loss_new = add_rn(loss_prev, stop_prev < thr ? o.kl : 0.0f);
  o.digit_inc = o.stop_new < thr ? 1 : 0;
  return o;
}

// d(loss)/d(log_odds) given gy_z = d/dz (train only) and gk = d/dkl (concrete.py:30-43 autodiff)
__device__ __forceinline__ float concrete_bwd_one(float lo, float yy, float zz, float gz, float gk, float prior,
                                                  float tau, int train) {
  const float yt = yy * tau;
  const float eq = expf(-yt + lo), sq = eq / (1.0f + eq + kEps);
  const float ep = expf(-yt + prior), sp = ep / (1.0f + ep + kEps);
  const float dkl_dy = tau * (2.0f * sq - 2.0f * sp);
  const float dkl_dlo = 1.0f - 2.0f * sq;
  float gy = gk * dkl_dy;
  if (train) gy += gz * zz * (1.0f - zz);
  return gy / tau + gk * dkl_dlo;
}

}  // namespace air
