// Counter-based sampling noise, generated where it is consumed (no noise tensors in HBM, no library RNG kernels in the
// step): Philox4x32-10 keyed by the model's seed, counter = (element block, stream, step), and a Box-Muller transform
// that uses ONLY correctly rounded fp32 +, -, *, /, sqrt (no libm, no FMA contraction), so that the host restatement
// in oracle/rng_oracle.py reproduces every bit (tests/test_gpu_ops.py).  (The likelihood stream, generated inside a GEMM
// epilogue, uses the SFU form of the same transform: rng_normal_pair_fast.)  The reference draws tf.random_normal /
// tf.random_uniform (air_model.py:123-128, vae.py:23, 37, concrete.py:23); its Philox streams cannot be reproduced
// outside TensorFlow, parity runs inject noise -- this generator only has to be N(0,1) / U[0,1) and deterministic.
//
// Device RNG state (caller-owned, 4 x uint64): [0] seed, [1] step counter, [2] ticket of air_noise_fill, [3] unused.
// Element i of a stream uses Philox block i >> 2 and word pair (i & 2): words (0,1) -> elements 4j, 4j+1; (2,3) -> 4j+2, 4j+3.
#pragma once
#include "air_common.cuh"

namespace air {

enum RngStream : uint32_t { kRngScale = 1, kRngShift = 2, kRngLatent = 3, kRngConcrete = 4, kRngLike = 5 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ uint4 rng_block(unsigned long long seed, unsigned long long counter, uint32_t stream,
                                           unsigned long long block) {
  return philox4x32_10(make_uint4(static_cast<uint32_t>(block), static_cast<uint32_t>(block >> 32), stream,
                                  static_cast<uint32_t>(counter)),
                       static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
}

__device__ __forceinline__ float rng_uniform(uint32_t w) {  // [0, 1): the top 24 bits
  return __fmul_rn(static_cast<float>(w >> 8), 5.9604644775390625e-08f);
}

// two N(0,1) samples from two words; every operation is one correctly rounded fp32 operation (see rng_oracle.py)
__device__ __forceinline__ void rng_normal_pair(uint32_t w0, uint32_t w1, float &n0, float &n1) {
  const uint32_t j = (w0 >> 8) + 1u;                                    // u1 = j * 2^-24 in (0, 1]
  const int e = 31 - __clz(j);
  const float m = __fmul_rn(static_cast<float>(j), __int_as_float((127 - e) << 23));  // j / 2^e in [1, 2), exact
  const float s = __fdiv_rn(__fsub_rn(m, 1.0f), __fadd_rn(m, 1.0f));
  const float s2 = __fmul_rn(s, s);
  float p = static_cast<float>(1.0 / 11.0);                              // ln m = 2 (s + s^3/3 + ... + s^11/11)
  p = __fadd_rn(__fmul_rn(p, s2), static_cast<float>(1.0 / 9.0));
  p = __fadd_rn(__fmul_rn(p, s2), static_cast<float>(1.0 / 7.0));
  p = __fadd_rn(__fmul_rn(p, s2), static_cast<float>(1.0 / 5.0));
  p = __fadd_rn(__fmul_rn(p, s2), static_cast<float>(1.0 / 3.0));
  p = __fadd_rn(__fmul_rn(p, s2), 1.0f);
  const float ln_m = __fmul_rn(__fmul_rn(2.0f, s), p);
  const float ln_u = __fadd_rn(__fmul_rn(static_cast<float>(e - 24), static_cast<float>(0.6931471805599453)), ln_m);
  const float r = __fsqrt_rn(__fmul_rn(-2.0f, ln_u));
  const uint32_t k = w1 >> 8;                                           // angle = 2 pi k / 2^24 = (q + t) pi / 4
  const uint32_t q = k >> 21;
  const float t = __fmul_rn(static_cast<float>(k & 0x1FFFFFu), 4.76837158203125e-07f);  // 2^-21
  const bool odd = q & 1u;
  const float x = __fmul_rn(odd ? __fsub_rn(1.0f, t) : t, static_cast<float>(0.7853981633974483));
  const float x2 = __fmul_rn(x, x);
  float ps = static_cast<float>(1.0 / 362880.0);
  ps = __fadd_rn(__fmul_rn(ps, x2), static_cast<float>(-1.0 / 5040.0));
  ps = __fadd_rn(__fmul_rn(ps, x2), static_cast<float>(1.0 / 120.0));
  ps = __fadd_rn(__fmul_rn(ps, x2), static_cast<float>(-1.0 / 6.0));
  ps = __fadd_rn(__fmul_rn(ps, x2), 1.0f);
  float sx = __fmul_rn(x, ps);
  float pc = static_cast<float>(-1.0 / 3628800.0);
  pc = __fadd_rn(__fmul_rn(pc, x2), static_cast<float>(1.0 / 40320.0));
  pc = __fadd_rn(__fmul_rn(pc, x2), static_cast<float>(-1.0 / 720.0));
  pc = __fadd_rn(__fmul_rn(pc, x2), static_cast<float>(1.0 / 24.0));
  pc = __fadd_rn(__fmul_rn(pc, x2), -0.5f);
  pc = __fadd_rn(__fmul_rn(pc, x2), 1.0f);
  const float cx = pc;
  if (odd) sx = -sx;                                                    // theta = base * pi/2 + (+x | -x)
  const uint32_t base = ((q + 1u) >> 1) & 3u;
  const float ct = base == 0 ? cx : base == 1 ? -sx : base == 2 ? -cx : sx;
  const float st = base == 0 ? sx : base == 1 ? cx : base == 2 ? -sx : -cx;
  n0 = __fmul_rn(r, ct);
  n1 = __fmul_rn(r, st);
}

// The same Box-Muller pair from the SFU approximations (lg2 / rsqrt / sin / cos: ~15 instructions instead of ~80): used
// for the one stream that is consumed millions of times per step inside a GEMM epilogue -- the VAE likelihood noise
// (kRngLike), which only ever enters as sigmoid(gen + 0.3 n).  Same Philox words, same radius / angle mapping up to a
// half turn; the samples agree with the host restatement to ~1e-5 absolute instead of bit for bit (measured: the exact
// transform in the gen_mean epilogue cost 75 us per train step, this one ~10).
__device__ __forceinline__ void rng_normal_pair_fast(uint32_t w0, uint32_t w1, float &n0, float &n1) {
  const float u1 = __fmul_rn(static_cast<float>((w0 >> 8) + 1u), 5.9604644775390625e-08f);      // (0, 1]
  const float r2 = -1.3862943611198906f * __log2f(u1);                                           // -2 ln u1 >= 0
  const float r = r2 > 0.0f ? r2 * rsqrtf(r2) : 0.0f;
  const float a = (static_cast<float>(w1 >> 8) - 8388608.0f) * 3.7450702829238413e-07f;          // 2 pi (k - 2^23) / 2^24 in [-pi, pi)
  float sa, ca;
  __sincosf(a, &sa, &ca);
  n0 = r * ca;
  n1 = r * sa;
}

template <bool FAST = false>
__device__ __forceinline__ float4 rng_normal4_t(unsigned long long seed, unsigned long long counter, uint32_t stream,
                                                unsigned long long block) {
  const uint4 w = rng_block(seed, counter, stream, block);
  float4 n;
  if (FAST) {
    rng_normal_pair_fast(w.x, w.y, n.x, n.y);
    rng_normal_pair_fast(w.z, w.w, n.z, n.w);
  } else {
    rng_normal_pair(w.x, w.y, n.x, n.y);
    rng_normal_pair(w.z, w.w, n.z, n.w);
  }
  return n;
}
// exact (bit-reproducible) transform for the small per-image streams, SFU transform for the likelihood stream
__device__ __forceinline__ float4 rng_normal4(unsigned long long seed, unsigned long long counter, uint32_t stream,
                                              unsigned long long block) {
  return stream == kRngLike ? rng_normal4_t<true>(seed, counter, stream, block) : rng_normal4_t<false>(seed, counter, stream, block);
}

// one element (slow paths only: recomputes the block of four)
__device__ __forceinline__ float rng_normal_elem(unsigned long long seed, unsigned long long counter, uint32_t stream,
                                                 unsigned long long i) {
  const float4 n = rng_normal4(seed, counter, stream, i >> 2);
  const int l = static_cast<int>(i & 3);
  return l == 0 ? n.x : l == 1 ? n.y : l == 2 ? n.z : n.w;
}

}  // namespace air
