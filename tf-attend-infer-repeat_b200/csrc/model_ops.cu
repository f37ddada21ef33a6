// Fused model-specific kernels of the AIR loop body (air/air_model.py:278-508, 580-611,
// 651-694): LSTM pointwise, the Gaussian heads (sampling, squashing, KLs, theta, theta^-1)
// fused with the Concrete/ACT step, the VAE latent, the noisy-sigmoid output, the BCE
// reconstruction loss, deterministic bias-gradient column sums, and global-norm clipping +
// TF-flavoured Adam on a flat parameter buffer.  All HBM-bound elementwise / reduction work.
// Compiled with --fmad=false so that each reference op is rounded separately.
#include <algorithm>
#include <math.h>

#include "air_common.cuh"
#include "concrete.cuh"
#include "epilogue.cuh"

namespace air {

static inline int grid_for(int64_t n, int threads, int per_sm = 8) {
  return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads,
                                                                  static_cast<int64_t>(sm_count()) * per_sm)));
}

#define AIR_GRID_STRIDE(i, n)                                                                    \
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < (n);         \
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)

// =========================================================================================
// LSTM pointwise (BasicLSTMCell: i, j, f, o; forget bias 1.0)
// =========================================================================================
__device__ __forceinline__ void lstm_cell_fwd(float gi, float gj, float gf, float go, float cp, float &c, float &h) {
  c = cp * sigmoid_f(gf + 1.0f) + sigmoid_f(gi) * tanhf(gj);
  h = tanhf(c) * sigmoid_f(go);
}

// VEC = 4: four consecutive cells per thread, 16-byte loads/stores (H % 4 == 0, 16-byte aligned rows)
template <int VEC>
__global__ void __launch_bounds__(256)
    lstm_fwd_k(const float *__restrict__ gates, const float *__restrict__ c_prev, float *__restrict__ c_new,
               float *__restrict__ h_new, int64_t B, int H) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  const int HV = H / VEC;
  AIR_GRID_STRIDE(ev, B * HV) {
    const int64_t b = ev / HV;
    const int k = static_cast<int>(ev - b * HV) * VEC;
    const float *g = gates + b * 4 * H + k;
    const int64_t e = b * H + k;
    if (VEC == 4) {
      const float4 gi = *reinterpret_cast<const float4 *>(g), gj = *reinterpret_cast<const float4 *>(g + H);
      const float4 gf = *reinterpret_cast<const float4 *>(g + 2 * H), go = *reinterpret_cast<const float4 *>(g + 3 * H);
      const float4 cp = c_prev ? *reinterpret_cast<const float4 *>(c_prev + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 c, h;
      lstm_cell_fwd(gi.x, gj.x, gf.x, go.x, cp.x, c.x, h.x);
      lstm_cell_fwd(gi.y, gj.y, gf.y, go.y, cp.y, c.y, h.y);
      lstm_cell_fwd(gi.z, gj.z, gf.z, go.z, cp.z, c.z, h.z);
      lstm_cell_fwd(gi.w, gj.w, gf.w, go.w, cp.w, c.w, h.w);
      *reinterpret_cast<float4 *>(c_new + e) = c;
      *reinterpret_cast<float4 *>(h_new + e) = h;
    } else {
      float c, h;
      lstm_cell_fwd(g[0], g[H], g[2 * H], g[3 * H], c_prev ? c_prev[e] : 0.0f, c, h);
      c_new[e] = c;
      h_new[e] = h;
    }
  }
}

struct LstmGrad { float di, dj, df, dd, dcp; };
__device__ __forceinline__ LstmGrad lstm_cell_bwd(float gi, float gj, float gf, float go, float cp, float cn, float gh,
                                                  float dcn) {
  const float i = sigmoid_f(gi), j = tanhf(gj), f = sigmoid_f(gf + 1.0f), o = sigmoid_f(go);
  const float tc = tanhf(cn);
  const float dc = dcn + gh * o * (1.0f - tc * tc);
  LstmGrad r;
  r.di = dc * j * i * (1.0f - i);
  r.dj = dc * i * (1.0f - j * j);
  r.df = dc * cp * f * (1.0f - f);
  r.dd = gh * tc * o * (1.0f - o);
  r.dcp = dc * f;
  return r;
}

template <int VEC>
__global__ void __launch_bounds__(256)
    lstm_bwd_k(const float *__restrict__ gates, const float *__restrict__ c_prev, const float *__restrict__ c_new,
               const float *__restrict__ dh, const float *dc_new, float *__restrict__ dgates, float *dc_prev,
               float *dgates_sum, int64_t B, int H) {
  pdl_sync();  // PDL: no global access before the previous grid has completed  // dc_prev may alias dc_new (same index, read before write)
  const int HV = H / VEC;
  AIR_GRID_STRIDE(ev, B * HV) {
    const int64_t b = ev / HV;
    const int k = static_cast<int>(ev - b * HV) * VEC;
    const float *g = gates + b * 4 * H + k;
    float *dg = dgates + b * 4 * H + k;
    const int64_t e = b * H + k;
    if (VEC == 4) {
      typedef const float4 *cf4;
      typedef float4 *f4;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 gi = *cf4(g), gj = *cf4(g + H), gf = *cf4(g + 2 * H), go = *cf4(g + 3 * H);
      const float4 cp = c_prev ? *cf4(c_prev + e) : z4, cn = *cf4(c_new + e), gh = *cf4(dh + e);
      const float4 dcn = dc_new ? *cf4(dc_new + e) : z4;
      const LstmGrad r0 = lstm_cell_bwd(gi.x, gj.x, gf.x, go.x, cp.x, cn.x, gh.x, dcn.x);
      const LstmGrad r1 = lstm_cell_bwd(gi.y, gj.y, gf.y, go.y, cp.y, cn.y, gh.y, dcn.y);
      const LstmGrad r2 = lstm_cell_bwd(gi.z, gj.z, gf.z, go.z, cp.z, cn.z, gh.z, dcn.z);
      const LstmGrad r3 = lstm_cell_bwd(gi.w, gj.w, gf.w, go.w, cp.w, cn.w, gh.w, dcn.w);
      const float4 di = make_float4(r0.di, r1.di, r2.di, r3.di), dj = make_float4(r0.dj, r1.dj, r2.dj, r3.dj);
      const float4 df = make_float4(r0.df, r1.df, r2.df, r3.df), dd = make_float4(r0.dd, r1.dd, r2.dd, r3.dd);
      *f4(dg) = di; *f4(dg + H) = dj; *f4(dg + 2 * H) = df; *f4(dg + 3 * H) = dd;
      if (dgates_sum) {
        float *ds = dgates_sum + b * 4 * H + k;
        float4 a = *cf4(ds), bq = *cf4(ds + H), c = *cf4(ds + 2 * H), d = *cf4(ds + 3 * H);
        a.x += di.x; a.y += di.y; a.z += di.z; a.w += di.w;
        bq.x += dj.x; bq.y += dj.y; bq.z += dj.z; bq.w += dj.w;
        c.x += df.x; c.y += df.y; c.z += df.z; c.w += df.w;
        d.x += dd.x; d.y += dd.y; d.z += dd.z; d.w += dd.w;
        *f4(ds) = a; *f4(ds + H) = bq; *f4(ds + 2 * H) = c; *f4(ds + 3 * H) = d;
      }
      *f4(dc_prev + e) = make_float4(r0.dcp, r1.dcp, r2.dcp, r3.dcp);
    } else {
      const LstmGrad r = lstm_cell_bwd(g[0], g[H], g[2 * H], g[3 * H], c_prev ? c_prev[e] : 0.0f, c_new[e], dh[e],
                                       dc_new ? dc_new[e] : 0.0f);
      dg[0] = r.di; dg[H] = r.dj; dg[2 * H] = r.df; dg[3 * H] = r.dd;
      if (dgates_sum) {
        float *ds = dgates_sum + b * 4 * H + k;
        ds[0] += r.di; ds[H] += r.dj; ds[2 * H] += r.df; ds[3 * H] += r.dd;
      }
      dc_prev[e] = r.dcp;
    }
  }
}

// =========================================================================================
// Heads: 8 lanes per image.  Lane j < 7 owns head output j:
//   0 scale/mean  1 scale/log_variance  2,3 shift/mean x,y  4,5 shift/log_variance x,y  6 z_pres/log_odds
// =========================================================================================
__device__ __forceinline__ int head_block(int j) { return j < 2 ? j : (j < 4 ? 2 : (j < 6 ? 3 : 4)); }

// 0.5 * (((plv - lv) - 1) + var/pv + (mean-pm)^2/pv) for one dimension (air_model.py:443-447)
__device__ __forceinline__ float gauss_kl_term(float mean, float lv, float var, float pm, float pv, float plv) {
  const float d = mean - pm;
  return (((plv - lv) - 1.0f) + var / pv) + (d * d) / pv;
}

__global__ void __launch_bounds__(256)
    heads_fwd_k(const float *__restrict__ hidden, const float *__restrict__ w_out, const float *__restrict__ b_out,
                const float *__restrict__ n_scale, const float *__restrict__ n_shift, const float *__restrict__ u,
                const float *__restrict__ prior_p, air_hyper_t hp, float *stop, float *loss, int32_t *digits,
                float *__restrict__ fields, float *__restrict__ theta, float *__restrict__ theta_inv, int64_t B, int HU,
                int vec4) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  const int64_t gt = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t b = gt >> 3;
  const int j = static_cast<int>(gt & 7);
  const bool valid = b < B;
  float out = 0.0f;
  if (valid && j < 7) {
    const float *h = hidden + b * 5 * HU + head_block(j) * HU;
    const float *w = w_out + j * HU;
    float acc = 0.0f;  // k-sequential FMA, like air_gemm's exact mode
    if (vec4) {        // 16-byte loads, same order of the 64 FMAs
      for (int k = 0; k < HU; k += 4) {
        const float4 hv = *reinterpret_cast<const float4 *>(h + k), wv = __ldg(reinterpret_cast<const float4 *>(w + k));
        acc = __fmaf_rn(hv.x, wv.x, acc);
        acc = __fmaf_rn(hv.y, wv.y, acc);
        acc = __fmaf_rn(hv.z, wv.z, acc);
        acc = __fmaf_rn(hv.w, wv.w, acc);
      }
    } else {
      for (int k = 0; k < HU; ++k) acc = __fmaf_rn(h[k], __ldg(w + k), acc);
    }
    out = acc + __ldg(b_out + j);
  }
  const unsigned full = 0xffffffffu;
  const int base = (threadIdx.x & 31) & ~7;
  float o[7];
#pragma unroll
  for (int q = 0; q < 7; ++q) o[q] = __shfl_sync(full, out, base + q);
  if (!valid || j != 0) return;

  const float thr = hp.stopping_threshold;
  // ---- scale (air_model.py:300-303) and shift (:317-320)
  const float s_var = expf(o[1]);
  const float s = sigmoid_f(o[0] + n_scale[b] * sqrtf(s_var));
  const float vx = expf(o[4]), vy = expf(o[5]);
  const float x = tanhf(o[2] + n_shift[2 * b] * sqrtf(vx));
  const float y = tanhf(o[3] + n_shift[2 * b + 1] * sqrtf(vy));
  // ---- theta (:324-327) and theta^-1 (:353-356, three separate divisions)
  float *th = theta + b * 6;
  th[0] = s; th[1] = 0.0f; th[2] = x; th[3] = 0.0f; th[4] = s; th[5] = y;
  float *ti = theta_inv + b * 6;
  const float inv = 1.0f / s;
  ti[0] = inv; ti[1] = 0.0f; ti[2] = -x / s; ti[3] = 0.0f; ti[4] = inv; ti[5] = -y / s;
  // ---- Concrete / ACT step (:380-427)
  const float stop_prev = stop[b];
  const ConcreteOut c = concrete_step_one(o[6], u[b], stop_prev, loss[b], __ldg(prior_p), hp.z_pres_temperature, thr,
                                          hp.train);
  const bool live = c.stop_new < thr;
  // ---- Gaussian KLs masked by the NEW stopping sum (:441-477)
  const float kl_s = 0.5f * gauss_kl_term(o[0], o[1], s_var, hp.scale_prior_mean, hp.scale_prior_variance,
                                          logf(hp.scale_prior_variance));
  const float plv = logf(hp.shift_prior_variance);
  const float kl_t = 0.5f * (gauss_kl_term(o[2], o[4], vx, hp.shift_prior_mean, hp.shift_prior_variance, plv) +
                             gauss_kl_term(o[3], o[5], vy, hp.shift_prior_mean, hp.shift_prior_variance, plv));
  float l = c.loss_new;
  l = l + (live ? kl_s : 0.0f);
  l = l + (live ? kl_t : 0.0f);
  stop[b] = c.stop_new;
  loss[b] = l;
  digits[b] += c.digit_inc;
  float *f = fields + b;
#pragma unroll
  for (int q = 0; q < 7; ++q) f[static_cast<int64_t>(q) * B] = o[q];
  f[AIR_F_S * B] = s; f[AIR_F_X * B] = x; f[AIR_F_Y * B] = y;
  f[AIR_F_YPRE * B] = c.y; f[AIR_F_Z * B] = c.z; f[AIR_F_ZPROB * B] = c.z_prob;
  f[AIR_F_KL_Z * B] = c.kl; f[AIR_F_KL_SCALE * B] = kl_s; f[AIR_F_KL_SHIFT * B] = kl_t;
  f[AIR_F_STOP_PREV * B] = stop_prev; f[AIR_F_STOP_NEW * B] = c.stop_new;
}

constexpr int kHeadsImgs = 8;  // images per 256-thread CTA in heads_bwd (4x the CTAs of the first version:
                               // the kernel is latency-bound, ~5 MB of traffic)

__global__ void __launch_bounds__(256)
    heads_bwd_k(const float *__restrict__ hidden, const float *__restrict__ w_out, const float *__restrict__ n_scale,
                const float *__restrict__ n_shift, const float *__restrict__ fields, const float *__restrict__ dtheta,
                const float *__restrict__ dtheta_inv, const float *__restrict__ dz, const float *__restrict__ prior_p,
                air_hyper_t hp, float dloss, float *__restrict__ dhidden, float *__restrict__ partials, int64_t B,
                int HU) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  __shared__ float sOut[kHeadsImgs][8];
  const int tid = threadIdx.x;
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * kHeadsImgs;
  const int n_img = static_cast<int>((B - b0 < kHeadsImgs ? B - b0 : kHeadsImgs));
  if (tid < kHeadsImgs) {  // one thread per image: gradients w.r.t. the 7 head outputs
    const int64_t b = b0 + tid;
    float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (b < B) {
      const float *f = fields + b;
      const float thr = hp.stopping_threshold;
      const float m_s = f[AIR_F_SCALE_MEAN * B], lv_s = f[AIR_F_SCALE_LV * B];
      const float m_x = f[AIR_F_SHIFT_MEAN_X * B], m_y = f[AIR_F_SHIFT_MEAN_Y * B];
      const float lv_x = f[AIR_F_SHIFT_LV_X * B], lv_y = f[AIR_F_SHIFT_LV_Y * B];
      const float s = f[AIR_F_S * B], x = f[AIR_F_X * B], y = f[AIR_F_Y * B];
      const bool live = f[AIR_F_STOP_NEW * B] < thr;
      const float gl = live ? dloss : 0.0f;                                // weight of the KLs masked by the new sum
      const float gk = f[AIR_F_STOP_PREV * B] < thr ? dloss : 0.0f;        // weight of the z_pres KL (previous sum)
      const float *dt = dtheta + b * 6, *di = dtheta_inv + b * 6;
      // theta = [s,0,x;0,s,y]; theta_inv = [1/s,0,-x/s;0,1/s,-y/s]
      const float is = 1.0f / s, is2 = is * is;
      const float ds = (dt[0] + dt[4]) - (di[0] + di[4]) * is2 + (di[2] * x + di[5] * y) * is2;
      const float dx = dt[2] - di[2] * is;
      const float dy = dt[5] - di[5] * is;
      // s = sigmoid(a), a = mean + n*sqrt(var), var = exp(lv)
      const float var_s = expf(lv_s), var_x = expf(lv_x), var_y = expf(lv_y);
      const float da_s = ds * s * (1.0f - s);
      const float da_x = dx * (1.0f - x * x), da_y = dy * (1.0f - y * y);
      const float pvs = hp.scale_prior_variance, pvt = hp.shift_prior_variance;
      d[0] = da_s + gl * (m_s - hp.scale_prior_mean) / pvs;
      d[1] = da_s * n_scale[b] * 0.5f * sqrtf(var_s) + gl * 0.5f * (var_s / pvs - 1.0f);
      d[2] = da_x + gl * (m_x - hp.shift_prior_mean) / pvt;
      d[3] = da_y + gl * (m_y - hp.shift_prior_mean) / pvt;
      d[4] = da_x * n_shift[2 * b] * 0.5f * sqrtf(var_x) + gl * 0.5f * (var_x / pvt - 1.0f);
      d[5] = da_y * n_shift[2 * b + 1] * 0.5f * sqrtf(var_y) + gl * 0.5f * (var_y / pvt - 1.0f);
      d[6] = concrete_bwd_one(f[AIR_F_LOG_ODDS * B], f[AIR_F_YPRE * B], f[AIR_F_Z * B], dz[b], gk, __ldg(prior_p),
                              hp.z_pres_temperature, hp.train);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) sOut[tid][q] = d[q];
  }
  __syncthreads();
  // ---- dhidden[b, blk*HU + k] = relu'(hidden) * sum_{j in blk} dOut[j] * w_out[j][k]; the CTA's images are
  //      contiguous rows of [B, 5*HU], so consecutive threads touch consecutive floats
  const int n = 5 * HU;
  for (int e = tid; e < n_img * n; e += 256) {
    const int img = e / n, q = e - img * n;
    const int blk = q / HU, k = q - blk * HU;
    const float *so = sOut[img];
    float g;
    if (blk < 2) g = so[blk] * __ldg(w_out + blk * HU + k);
    else if (blk == 2) g = so[2] * __ldg(w_out + 2 * HU + k) + so[3] * __ldg(w_out + 3 * HU + k);
    else if (blk == 3) g = so[4] * __ldg(w_out + 4 * HU + k) + so[5] * __ldg(w_out + 5 * HU + k);
    else g = so[6] * __ldg(w_out + 6 * HU + k);
    const int64_t o = b0 * n + e;
    dhidden[o] = hidden[o] > 0.0f ? g : 0.0f;
  }
  // ---- per-CTA partial d(w_out) [7,HU] and d(b_out) [7]: fixed order over the CTA's images
  float *part = partials + static_cast<int64_t>(blockIdx.x) * (7 * HU + 7);
  for (int e = tid; e < 7 * HU + 7; e += 256) {
    float acc = 0.0f;
    if (e < 7 * HU) {
      const int jj = e / HU, k = e - jj * HU;
      const float *h = hidden + b0 * 5 * HU + head_block(jj) * HU + k;
      float hv[kHeadsImgs];
#pragma unroll
      for (int i = 0; i < kHeadsImgs; ++i) hv[i] = i < n_img ? h[static_cast<int64_t>(i) * 5 * HU] : 0.0f;  // loads in flight together
#pragma unroll
      for (int i = 0; i < kHeadsImgs; ++i) acc += sOut[i][jj] * hv[i];   // sOut rows beyond n_img are zero
    } else {
      const int jj = e - 7 * HU;
#pragma unroll
      for (int i = 0; i < kHeadsImgs; ++i) acc += sOut[i][jj];
    }
    part[e] = acc;
  }
}

// out[e] (+)= sum_r partials[r*stride + e]: one warp per output element, lane l adds rows l, l+32, ...
// in increasing order and the 32 lane sums are combined by a fixed shuffle tree (deterministic).
__global__ void __launch_bounds__(256)
    reduce_rows_k(const float *__restrict__ partials, int R, int stride, int n, float *out, int accumulate) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t e = warp0; e < n; e += nwarps) {
    float acc = 0.0f;
    for (int r = lane; r < R; r += 32) acc += partials[static_cast<int64_t>(r) * stride + e];
    acc = warp_sum(acc);
    if (lane == 0) out[e] = accumulate ? out[e] + acc : acc;
  }
}

int reduce_rows_launch(const float *partials, int R, int stride, int n, float *out, int accumulate, cudaStream_t s) {
  AIR_LAUNCH(reduce_rows_k, grid_for(static_cast<int64_t>(n) * 32, 256), 256, 0, s, partials, R, stride, n, out, accumulate);
  count_launch();
  return check_launch("reduce_rows");
}

// =========================================================================================
// VAE latent: one warp per image
// =========================================================================================
__global__ void __launch_bounds__(256)
    vae_latent_fwd_k(const float *__restrict__ ml, const float *__restrict__ noise, air_hyper_t hp,
                     float *__restrict__ sample, int lds, float *__restrict__ fields, float *loss, int64_t B, int L) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  const int64_t b = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float pm = hp.vae_prior_mean, pv = hp.vae_prior_variance, plv = logf(pv);
  float acc = 0.0f;
  for (int d = lane; d < L; d += 32) {
    const float mean = ml[b * 2 * L + d], lv = ml[b * 2 * L + L + d];
    const float var = expf(lv);
    sample[b * lds + d] = mean + noise[b * L + d] * sqrtf(var);
    acc += gauss_kl_term(mean, lv, var, pm, pv, plv);
  }
  acc = 0.5f * warp_sum(acc);
  if (lane == 0) {
    const bool live = fields[AIR_F_STOP_NEW * B + b] < hp.stopping_threshold;
    fields[AIR_F_KL_VAE * B + b] = acc;
    loss[b] = loss[b] + (live ? acc : 0.0f);
  }
}

__global__ void __launch_bounds__(256)
    vae_latent_bwd_k(const float *__restrict__ ml, const float *__restrict__ noise, const float *__restrict__ dsample,
                     const float *__restrict__ fields, air_hyper_t hp, float dloss, float *__restrict__ dml, int64_t B,
                     int L) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  AIR_GRID_STRIDE(e, B * L) {
    const int64_t b = e / L;
    const int d = static_cast<int>(e - b * L);
    const float gl = fields[AIR_F_STOP_NEW * B + b] < hp.stopping_threshold ? dloss : 0.0f;
    const float mean = ml[b * 2 * L + d], lv = ml[b * 2 * L + L + d];
    const float var = expf(lv);
    const float ds = dsample[e];
    dml[b * 2 * L + d] = ds + gl * (mean - hp.vae_prior_mean) / hp.vae_prior_variance;
    dml[b * 2 * L + L + d] = ds * noise[e] * 0.5f * sqrtf(var) + gl * 0.5f * (var / hp.vae_prior_variance - 1.0f);
  }
}

__global__ void __launch_bounds__(256)
    sigmoid_noise_fwd_k(const float *__restrict__ gen, const float *__restrict__ noise, float sd, float *__restrict__ out,
                        int64_t n) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  AIR_GRID_STRIDE(e, n) out[e] = sigmoid_f(gen[e] + noise[e] * sd);
}

__global__ void __launch_bounds__(256)
    sigmoid_bwd_k(const float *out, const float *dout, float *dgen, int64_t n) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  AIR_GRID_STRIDE(e, n) {
    const float o = out[e];
    dgen[e] = dout[e] * o * (1.0f - o);
  }
}

// =========================================================================================
// BCE reconstruction loss: one CTA per image
// =========================================================================================
__device__ __forceinline__ float bce_one(float cv, float xv, float dscale, bool want_grad, float &r, float &dc) {
  r = fmaxf(fminf(cv, 1.0f), 0.0f);
  const float a = r + kEps, q = (1.0f - r) + kEps;
  if (want_grad) {
    const bool pass = cv <= 1.0f && cv >= 0.0f;  // minimum passes x<=y, maximum passes x>=y
    // a, q in [1e-9, 1 + 1e-9]: the 2-ulp fast division is far inside the 1e-4 gradient tolerance
    dc = pass ? dscale * (__fdividef(1.0f - xv, q) - __fdividef(xv, a)) : 0.0f;
  }
  return xv * logf(a) + (1.0f - xv) * logf(q);
}

// VEC = 4: N % 4 == 0 and 16-byte aligned rows
template <int VEC>
__global__ void __launch_bounds__(256)
    bce_loss_k(const float *__restrict__ canvas, const float *__restrict__ x, float *__restrict__ recon,
               float *__restrict__ rec_loss, float *__restrict__ dcanvas, float dscale, int N) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  __shared__ float red[8];
  const int64_t b = blockIdx.x;
  const float *c = canvas + b * N, *xi = x + b * N;
  const bool wg = dcanvas != nullptr;
  float acc = 0.0f;
  if (VEC == 4) {
    for (int p = threadIdx.x * 4; p < N; p += 1024) {
      const float4 cv = *reinterpret_cast<const float4 *>(c + p), xv = *reinterpret_cast<const float4 *>(xi + p);
      float4 r, d;
      // summed in pixel order within the thread, like the scalar path sums its own pixels
      acc += bce_one(cv.x, xv.x, dscale, wg, r.x, d.x);
      acc += bce_one(cv.y, xv.y, dscale, wg, r.y, d.y);
      acc += bce_one(cv.z, xv.z, dscale, wg, r.z, d.z);
      acc += bce_one(cv.w, xv.w, dscale, wg, r.w, d.w);
      if (recon) *reinterpret_cast<float4 *>(recon + b * N + p) = r;
      if (wg) *reinterpret_cast<float4 *>(dcanvas + b * N + p) = d;
    }
  } else {
    for (int p = threadIdx.x; p < N; p += 256) {
      float r, d;
      acc += bce_one(c[p], xi[p], dscale, wg, r, d);
      if (recon) recon[b * N + p] = r;
      if (wg) dcanvas[b * N + p] = d;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int k = 0; k < 8; ++k) t += red[k];
    rec_loss[b] = -t;
  }
}

__global__ void __launch_bounds__(1024)
    finalize_loss_k(const float *__restrict__ running_loss, const float *__restrict__ rec_loss,
                    const int32_t *__restrict__ digits, const int32_t *__restrict__ target, float *__restrict__ out,
                    float *__restrict__ loss_per_item, int64_t B) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  __shared__ float r0[32], r1[32];
  float a = 0.0f, c = 0.0f;
  for (int64_t b = threadIdx.x; b < B; b += 1024) {
    const float l = running_loss[b] + rec_loss[b];
    if (loss_per_item) loss_per_item[b] = l;
    a += l;
    c += (digits[b] == target[b]) ? 1.0f : 0.0f;
  }
  a = warp_sum(a);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { r0[threadIdx.x >> 5] = a; r1[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.0f, tc = 0.0f;
    for (int k = 0; k < 32; ++k) { ta += r0[k]; tc += r1[k]; }
    out[0] = ta / static_cast<float>(B);
    out[1] = tc / static_cast<float>(B);
  }
}

// =========================================================================================
// Column sums (bias gradients): stage 1 -> partials[R][N], stage 2 = reduce_rows_k
// =========================================================================================
__global__ void __launch_bounds__(256)
    colsum_partial_k(const float *__restrict__ X, int ld, float *__restrict__ partials, int64_t B, int N,
                     int rows_per_chunk) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_chunk;
  const int64_t r1 = (r0 + rows_per_chunk < B) ? r0 + rows_per_chunk : B;
  float acc = 0.0f;
  if (col < N) {
    int64_t r = r0 + w;
    for (; r + 56 < r1; r += 64) {  // 8 independent loads in flight per lane, summed in row order
      const float v0 = X[r * ld + col], v1 = X[(r + 8) * ld + col], v2 = X[(r + 16) * ld + col], v3 = X[(r + 24) * ld + col];
      const float v4 = X[(r + 32) * ld + col], v5 = X[(r + 40) * ld + col], v6 = X[(r + 48) * ld + col], v7 = X[(r + 56) * ld + col];
      acc = (((((((acc + v0) + v1) + v2) + v3) + v4) + v5) + v6) + v7;
    }
    for (; r < r1; r += 8) acc += X[r * ld + col];
  }
  red[w][lane] = acc;
  __syncthreads();
  if (w == 0 && col < N) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    partials[static_cast<int64_t>(blockIdx.y) * N + col] = t;
  }
}

// Several column sums in ONE launch (the bias gradients of a train step: ~10 tensors of very different widths).
struct ColsumItem { const float *X; float *out; int64_t rows; int ld, N, accumulate, R, rows_per_chunk, cta0, ws0, out0; };
constexpr int kMaxColsumItems = 16;
struct ColsumBatch { ColsumItem it[kMaxColsumItems]; int n; };

__global__ void __launch_bounds__(256) colsum_multi_partial_k(ColsumBatch batch, float *__restrict__ workspace) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  __shared__ float red[8][33];
  int i = 0;
  while (i + 1 < batch.n && static_cast<int>(blockIdx.x) >= batch.it[i + 1].cta0) ++i;
  const ColsumItem &q = batch.it[i];
  const int local = blockIdx.x - q.cta0, cb = (q.N + 31) / 32;
  const int chunk = local / cb, col = (local - chunk * cb) * 32 + (threadIdx.x & 31);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t r0 = static_cast<int64_t>(chunk) * q.rows_per_chunk;
  const int64_t r1 = (r0 + q.rows_per_chunk < q.rows) ? r0 + q.rows_per_chunk : q.rows;
  const float *X = q.X;
  const int64_t ld = q.ld;
  float acc = 0.0f;
  if (col < q.N) {
    int64_t r = r0 + w;
    for (; r + 56 < r1; r += 64) {  // 8 independent loads in flight per lane, summed in row order
      const float v0 = X[r * ld + col], v1 = X[(r + 8) * ld + col], v2 = X[(r + 16) * ld + col], v3 = X[(r + 24) * ld + col];
      const float v4 = X[(r + 32) * ld + col], v5 = X[(r + 40) * ld + col], v6 = X[(r + 48) * ld + col], v7 = X[(r + 56) * ld + col];
      acc = (((((((acc + v0) + v1) + v2) + v3) + v4) + v5) + v6) + v7;
    }
    for (; r < r1; r += 8) acc += X[r * ld + col];
  }
  red[w][lane] = acc;
  __syncthreads();
  if (w == 0 && col < q.N) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][lane];
    workspace[q.ws0 + static_cast<int64_t>(chunk) * q.N + col] = t;
  }
}

// second stage for all items: one warp per output element (fixed order -> deterministic)
__global__ void __launch_bounds__(256) colsum_multi_reduce_k(ColsumBatch batch, const float *__restrict__ workspace, int total) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t e = warp0; e < total; e += nwarps) {
    int i = 0;
    while (i + 1 < batch.n && e >= batch.it[i + 1].out0) ++i;
    const ColsumItem &q = batch.it[i];
    const int col = static_cast<int>(e - q.out0);
    float acc = 0.0f;
    for (int r = lane; r < q.R; r += 32) acc += workspace[q.ws0 + static_cast<int64_t>(r) * q.N + col];
    acc = warp_sum(acc);
    if (lane == 0) q.out[col] = q.accumulate ? q.out[col] + acc : acc;
  }
}

static int colsum_chunks(int64_t B, int N) {
  const int cb = (N + 31) / 32;
  int R = (2 * sm_count() + cb - 1) / cb;
  R = std::max(1, std::min(R, 64));
  R = static_cast<int>(std::min<int64_t>(R, (B + 7) / 8));
  return std::max(R, 1);
}

// =========================================================================================
// Global-norm clip + TF Adam on the flat parameter buffer
// =========================================================================================
constexpr int kAdamPartials = 1024;

__global__ void __launch_bounds__(256)
    sumsq_partial_k(const float *__restrict__ g, float gs, float *__restrict__ partials, int64_t n) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  __shared__ float red[8];
  float acc = 0.0f;
  AIR_GRID_STRIDE(e, n) {
    const float v = g[e] * gs;
    acc += v * v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int k = 0; k < 8; ++k) t += red[k];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256)
    adam_apply_k(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                 const float *__restrict__ state, const float *__restrict__ partials, int n_partials, float clip,
                 float beta1, float beta2, float eps, float gs, float *__restrict__ norm_out, int64_t n, int vec4,
                 int skip_nonfinite) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  __shared__ float red[8];
  __shared__ float s_scale;
  // every CTA reduces the same partials in the same order -> identical clip scale everywhere
  float acc = 0.0f;
  for (int k = threadIdx.x; k < n_partials; k += 256) acc += partials[k];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int k = 0; k < 8; ++k) t += red[k];
    const float norm = sqrtf((t / 2.0f) * 2.0f);  // sqrt(2 * sum(l2_loss))
    float sc = 1.0f;
    if (clip > 0.0f) {  // tf.clip_by_global_norm; tf.minimum PROPAGATES a NaN norm (fminf would drop it): every gradient NaN
      const float inv = 1.0f / norm;
      sc = clip * ((inv < 1.0f / clip || inv != inv) ? inv : 1.0f / clip);
    }
    // AIR_ADAM_SKIP_NONFINITE: a non-finite global norm (an overflowed gradient) turns the whole step into a no-op
    // instead of writing NaN into every parameter the way clip_by_global_norm + ApplyAdam do
    if (skip_nonfinite && !(norm <= 3.0e38f)) sc = -1.0f;
    s_scale = sc;
    if (blockIdx.x == 0) *norm_out = norm;
  }
  __syncthreads();
  if (s_scale < 0.0f) return;  // (every CTA reads the same partials: uniform across the grid)
  const float sc = s_scale * gs;
  const float b1p = state[0], b2p = state[1], lr = state[4];
  const float alpha = lr * sqrtf(1.0f - b2p) / (1.0f - b1p);
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  auto one = [&](float gq, float &mm, float &vv, float &pp) {
    const float gv = gq * sc;
    mm = mm + (gv - mm) * omb1;
    vv = vv + (gv * gv - vv) * omb2;
    pp = pp - (mm * alpha) / (sqrtf(vv) + eps);
  };
  if (vec4) {  // 16-byte accesses: 7 streams of 16 MB each, HBM-bound
    AIR_GRID_STRIDE(q, n >> 2) {
      const float4 gq = reinterpret_cast<const float4 *>(g)[q];
      float4 mm = reinterpret_cast<float4 *>(m)[q], vv = reinterpret_cast<float4 *>(v)[q], pp = reinterpret_cast<float4 *>(p)[q];
      one(gq.x, mm.x, vv.x, pp.x);
      one(gq.y, mm.y, vv.y, pp.y);
      one(gq.z, mm.z, vv.z, pp.z);
      one(gq.w, mm.w, vv.w, pp.w);
      reinterpret_cast<float4 *>(m)[q] = mm;
      reinterpret_cast<float4 *>(v)[q] = vv;
      reinterpret_cast<float4 *>(p)[q] = pp;
    }
    for (int64_t e = (n & ~int64_t(3)) + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < n;
         e += static_cast<int64_t>(gridDim.x) * blockDim.x)
      one(g[e], m[e], v[e], p[e]);
  } else {
    AIR_GRID_STRIDE(e, n) one(g[e], m[e], v[e], p[e]);
  }
}

__global__ void adam_advance_k(float *state, float beta1, float beta2, const float *norm_in, int skip_nonfinite) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  state[3] = *norm_in;
  if (skip_nonfinite && !(*norm_in <= 3.0e38f)) {  // skipped step: optimizer state untouched, the skip is counted
    state[5] += 1.0f;
    return;
  }
  state[0] *= beta1;
  state[1] *= beta2;
  state[2] += 1.0f;
  state[3] = *norm_in;
}

__global__ void anneal_k(const float *state, float init, float factor, float iters, int staircase, float vmin,
                         float vmax, int take_log, float *out) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  float p = state[2] / iters;
  if (staircase) p = floorf(p);
  float v = init * powf(factor, p);
  if (vmin == vmin) v = fmaxf(v, vmin);
  if (vmax == vmax) v = fminf(v, vmax);
  if (take_log) v = logf(v + kEps);
  *out = v;
}

// =========================================================================================
// Synthetic multi-digit canvases on the device (stand-in for multi_mnist.py:82-183, whose MNIST source is not
// available offline): 0..max_digits stroke-like blobs per canvas, uniform placement with pixel-overlap rejection
// (generate_multi_image, :141-160), background exactly 0.  Counter-based RNG: image i of a seed is the same
// whatever the batch or the sharding.  Restated bit for bit in oracle/synth_oracle.py.
// =========================================================================================
__device__ __forceinline__ uint64_t synth_rnd(uint64_t seed, uint64_t image, uint32_t draw) {
  uint64_t z = seed + (image + 1) * 0x9E3779B97F4A7C15ull + static_cast<uint64_t>(draw) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Placement semantics of multi_mnist.py:82-183 (generate_multi_image with use_pixel_overlap, gap = margin = 0):
// every digit image is cropped to its non-empty bounding box (crop_non_empty, :36-43), placed at a uniform position
// x in [0, cs - w], y in [0, cs - h] (x drawn before y, :143-144); the first digit always fits, later ones are
// re-drawn up to 100 times until no pixel overlaps what is already on the canvas (pixels_overlap, :61-65); if a
// digit cannot be placed the WHOLE canvas is started over with fresh digits (:95-171), so the label always equals the
// number of digits drawn.  positions = [x, y] and boxes = [w, h] per placed digit (:165-166).
constexpr int kSynthAttempts = 100;  // multi_mnist.py:141
constexpr int kSynthRestarts = 64;   // the reference retries forever; a canvas that fails 64 times stays empty (label 0)

__global__ void __launch_bounds__(256)
    synth_canvases_k(uint64_t seed, int64_t first, float *__restrict__ images, int32_t *__restrict__ counts,
                     int32_t *__restrict__ positions, int32_t *__restrict__ boxes, int cs, int max_digits) {
  extern __shared__ float sC[];  // [cs*cs]
  __shared__ int bb[4];          // y0, y1, x0, x1 of the blob's non-empty pixels
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  const uint64_t img = static_cast<uint64_t>(first + b);
  int count = static_cast<int>((synth_rnd(seed, img, 0) >> 33) % static_cast<uint64_t>(max_digits + 1));
  uint32_t draw = 1;
  bool ready = count == 0;
  for (int p = tid; p < cs * cs; p += 256) sC[p] = 0.0f;
  for (int p = tid; p < 2 * max_digits; p += 256) {
    if (positions) positions[b * 2 * max_digits + p] = 0;
    if (boxes) boxes[b * 2 * max_digits + p] = 0;
  }
  __syncthreads();
  for (int restart = 0; restart < kSynthRestarts && !ready; ++restart) {
    if (restart) {
      for (int p = tid; p < cs * cs; p += 256) sC[p] = 0.0f;
      __syncthreads();
    }
    bool failed = false;
    for (int k = 0; k < count && !failed; ++k) {
      // ---- the "digit": a stroke-like blob in an hh x ww frame (MNIST itself is not available offline)
      const int hh = 14 + static_cast<int>((synth_rnd(seed, img, draw + 0) >> 33) % 11u);
      const int ww = 10 + static_cast<int>((synth_rnd(seed, img, draw + 1) >> 33) % 15u);
      const float bar_on = static_cast<float>((synth_rnd(seed, img, draw + 2) >> 33) & 1u);
      const float u = static_cast<float>(synth_rnd(seed, img, draw + 3) >> 40) * 5.9604644775390625e-08f;  // 2^-24
      const float bar_off = -2.0f + 4.0f * u;
      draw += 4;
      const float cy = static_cast<float>(hh - 1) / 2.0f, cx = static_cast<float>(ww - 1) / 2.0f;
      const float ry = fmaxf(static_cast<float>(hh) / 2.0f - 1.5f, 2.0f), rx = fmaxf(static_cast<float>(ww) / 2.0f - 1.5f, 2.0f);
      if (tid == 0) { bb[0] = hh; bb[1] = -1; bb[2] = ww; bb[3] = -1; }
      __syncthreads();
      float v[3];  // hh * ww <= 24 * 24 = 576 <= 3 * 256
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int p = tid + 256 * q;
        v[q] = 0.0f;
        if (p < hh * ww) {
          const int y = p / ww, x = p - y * ww;
          const float dy = (static_cast<float>(y) - cy) / ry, dx = (static_cast<float>(x) - cx) / rx;
          const float r = sqrtf(dy * dy + dx * dx);
          const float ring = fminf(fmaxf(1.0f - fabsf(r - 0.8f) * 3.0f, 0.0f), 1.0f);
          const float bar = fminf(fmaxf(1.0f - fabsf((static_cast<float>(x) - cx) - bar_off) / 1.6f, 0.0f), 1.0f) * bar_on;
          float val = fmaxf(ring, bar);
          if (val < 0.15f) val = 0.0f;
          v[q] = val;
          if (val > 0.0f) {  // crop_non_empty: integer min / max are order independent
            atomicMin(&bb[0], y); atomicMax(&bb[1], y); atomicMin(&bb[2], x); atomicMax(&bb[3], x);
          }
        }
      }
      __syncthreads();
      const int y0 = bb[0], x0 = bb[2], h = bb[1] - bb[0] + 1, w = bb[3] - bb[2] + 1;
      // ---- position: x then y, up to 100 draws
      bool found = false;
      int px = 0, py = 0;
      for (int attempt = 0; attempt < kSynthAttempts && !found; ++attempt) {
        px = static_cast<int>((synth_rnd(seed, img, draw + 0) >> 33) % static_cast<uint64_t>(cs - w + 1));
        py = static_cast<int>((synth_rnd(seed, img, draw + 1) >> 33) % static_cast<uint64_t>(cs - h + 1));
        draw += 2;
        int clash = 0;
        if (k > 0) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const int p = tid + 256 * q;
            if (p < hh * ww && v[q] > 0.0f) {
              const int y = p / ww, x = p - y * ww;
              if (sC[(py + y - y0) * cs + px + x - x0] > 0.0f) clash = 1;
            }
          }
        }
        found = !__syncthreads_or(clash);
      }
      if (!found) {
        failed = true;  // uniform: start the canvas over
        break;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int p = tid + 256 * q;
        if (p < hh * ww && v[q] > 0.0f) {
          const int y = p / ww, x = p - y * ww;
          sC[(py + y - y0) * cs + px + x - x0] += v[q];  // canvas[y:y+h, x:x+w] += image (no overlap: 0 + v)
        }
      }
      if (tid == 0) {
        if (positions) { positions[(b * max_digits + k) * 2] = px; positions[(b * max_digits + k) * 2 + 1] = py; }
        if (boxes) { boxes[(b * max_digits + k) * 2] = w; boxes[(b * max_digits + k) * 2 + 1] = h; }
      }
      __syncthreads();
    }
    ready = !failed;
  }
  if (!ready) {  // never observed; keeps label == digits drawn even then
    count = 0;
    for (int p = tid; p < cs * cs; p += 256) sC[p] = 0.0f;
    for (int p = tid; p < 2 * max_digits; p += 256) {
      if (positions) positions[b * 2 * max_digits + p] = 0;
      if (boxes) boxes[b * 2 * max_digits + p] = 0;
    }
  }
  __syncthreads();
  for (int p = tid; p < cs * cs; p += 256) images[b * cs * cs + p] = sC[p];
  if (tid == 0) counts[b] = count;
}

// ---- counter-based noise (rng.cuh) -------------------------------------------------------------------------------------
// One launch fills the four small noise tensors of a step.  Work item = one Philox block (4 samples) of one stream;
// the streams are laid end to end.  The step counter is advanced here: every CTA reads the old value c and samples
// with c + 1; the last CTA to finish (ticket) publishes c + 1, so the kernels that follow in the stream -- the
// AIR_EPI_SIGMOID_RNG epilogue -- see the counter this step was sampled with.
__global__ void __launch_bounds__(256)
    noise_fill_k(unsigned long long *state, float *__restrict__ scale, float *__restrict__ shift, float *__restrict__ latent,
                 float *__restrict__ conc, int64_t TB, int L) {
  pdl_sync();
  const unsigned long long seed = state[0], ctr = *reinterpret_cast<volatile unsigned long long *>(state + 1) + 1ull;
  const int64_t n[4] = {scale ? TB : 0, shift ? 2 * TB : 0, latent ? TB * L : 0, conc ? TB : 0};
  float *const out[4] = {scale, shift, latent, conc};
  const uint32_t stream[4] = {kRngScale, kRngShift, kRngLatent, kRngConcrete};
  int64_t nb[4], total = 0;
  for (int q = 0; q < 4; ++q) { nb[q] = (n[q] + 3) >> 2; total += nb[q]; }
  for (int64_t w = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; w < total; w += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int q = 0;
    int64_t blk = w;
    while (blk >= nb[q]) { blk -= nb[q]; ++q; }
    float4 v;
    if (q == 3) {
      const uint4 r = rng_block(seed, ctr, stream[q], static_cast<unsigned long long>(blk));
      v = make_float4(rng_uniform(r.x), rng_uniform(r.y), rng_uniform(r.z), rng_uniform(r.w));
    } else {
      v = rng_normal4(seed, ctr, stream[q], static_cast<unsigned long long>(blk));
    }
    float *o = out[q] + 4 * blk;
    if (4 * blk + 4 <= n[q] && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
      *reinterpret_cast<float4 *>(o) = v;
    } else {
      const float vv[4] = {v.x, v.y, v.z, v.w};
      for (int l = 0; l < 4 && 4 * blk + l < n[q]; ++l) o[l] = vv[l];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(state + 2, 1ull);
    if (t == gridDim.x - 1) {
      state[2] = 0;
      state[1] = ctr;
    }
  }
}

__global__ void __launch_bounds__(256) rng_samples_k(const unsigned long long *state, uint32_t stream, int uniform, float *out, int64_t n) {
  pdl_sync();
  const unsigned long long seed = state[0], ctr = state[1];
  for (int64_t blk = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; 4 * blk < n; blk += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float vv[4];
    if (uniform) {
      const uint4 r = rng_block(seed, ctr, stream, static_cast<unsigned long long>(blk));
      vv[0] = rng_uniform(r.x); vv[1] = rng_uniform(r.y); vv[2] = rng_uniform(r.z); vv[3] = rng_uniform(r.w);
    } else {
      const float4 v = rng_normal4(seed, ctr, stream, static_cast<unsigned long long>(blk));
      vv[0] = v.x; vv[1] = v.y; vv[2] = v.z; vv[3] = v.w;
    }
    for (int l = 0; l < 4 && 4 * blk + l < n; ++l) out[4 * blk + l] = vv[l];
  }
}

// several small buffers zeroed by ONE launch (the accumulators a step starts from), 16 bytes per thread and iteration
struct ZeroBatch { void *ptr[8]; int64_t n16[8]; int n; };
__global__ void __launch_bounds__(256) zero_many_k(ZeroBatch zb) {
  pdl_sync();
  for (int q = 0; q < zb.n; ++q) {
    uint4 *p = reinterpret_cast<uint4 *>(zb.ptr[q]);
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < zb.n16[q]; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
      p[i] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// uint8 canvases -> fp32, bit for bit what the MNIST loader behind multi_mnist.py produces: tensorflow's
// input_data.read_data_sets scales pixels with numpy.multiply(images.astype(float32), 1.0 / 255.0), i.e. in fp32
// fl(float(k) * fl(1/255)).  16 pixels per thread: one 128-bit load, four 128-bit stores.
__global__ void __launch_bounds__(256) expand_u8_k(const uint8_t *__restrict__ src, float *__restrict__ dst, int64_t n) {
  pdl_sync();
  const float inv = 1.0f / 255.0f;  // 0x3b808081, the fp32 constant numpy uses
  const int64_t nvec = n >> 4;
  for (int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v < nvec; v += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src) + v);
    const uint32_t wds[4] = {q.x, q.y, q.z, q.w};
    float4 *o = reinterpret_cast<float4 *>(dst) + 4 * v;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o[j] = make_float4(__fmul_rn(static_cast<float>(wds[j] & 0xffu), inv), __fmul_rn(static_cast<float>((wds[j] >> 8) & 0xffu), inv),
                         __fmul_rn(static_cast<float>((wds[j] >> 16) & 0xffu), inv), __fmul_rn(static_cast<float>(wds[j] >> 24), inv));
  }
  if (blockIdx.x == 0)  // tail
    for (int64_t i = (nvec << 4) + threadIdx.x; i < n; i += blockDim.x) dst[i] = __fmul_rn(static_cast<float>(src[i]), inv);
}

}  // namespace air

// ---- C ABI --------------------------------------------------------------------------------
using namespace air;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int air_lstm_fwd(const float *gates, const float *c_prev, float *c_new, float *h_new, int64_t B, int H,
                            air_stream_t stream) {
  AIR_REQUIRE(B >= 0 && H > 0, AIR_ERR_BAD_SHAPE, "lstm_fwd: bad shape");
  if (B == 0) return AIR_OK;
  AIR_REQUIRE(gates && c_new && h_new, AIR_ERR_NULL, "lstm_fwd: null pointer");
  if (H % 4 == 0 && aligned16(gates) && aligned16(c_new) && aligned16(h_new) && (!c_prev || aligned16(c_prev)))
    AIR_LAUNCH(lstm_fwd_k<4>, grid_for(B * H / 4, 256), 256, 0, ST(stream), gates, c_prev, c_new, h_new, B, H);
  else
    AIR_LAUNCH(lstm_fwd_k<1>, grid_for(B * H, 256), 256, 0, ST(stream), gates, c_prev, c_new, h_new, B, H);
  count_launch();
  return check_launch("lstm_fwd");
}

extern "C" int air_lstm_bwd(const float *gates, const float *c_prev, const float *c_new, const float *dh,
                            const float *dc_new, float *dgates, float *dc_prev, float *dgates_sum, int64_t B, int H,
                            air_stream_t stream) {
  AIR_REQUIRE(B >= 0 && H > 0, AIR_ERR_BAD_SHAPE, "lstm_bwd: bad shape");
  if (B == 0) return AIR_OK;
  AIR_REQUIRE(gates && c_new && dh && dgates && dc_prev, AIR_ERR_NULL, "lstm_bwd: null pointer");
  const bool vec = H % 4 == 0 && aligned16(gates) && aligned16(c_new) && aligned16(dh) && aligned16(dgates) &&
                   aligned16(dc_prev) && (!c_prev || aligned16(c_prev)) && (!dc_new || aligned16(dc_new)) &&
                   (!dgates_sum || aligned16(dgates_sum));
  if (vec)
    AIR_LAUNCH(lstm_bwd_k<4>, grid_for(B * H / 4, 256), 256, 0, ST(stream), gates, c_prev, c_new, dh, dc_new, dgates, dc_prev,
               dgates_sum, B, H);
  else
    AIR_LAUNCH(lstm_bwd_k<1>, grid_for(B * H, 256), 256, 0, ST(stream), gates, c_prev, c_new, dh, dc_new, dgates, dc_prev,
               dgates_sum, B, H);
  count_launch();
  return check_launch("lstm_bwd");
}

extern "C" int air_heads_fwd(const float *hidden, const float *w_out, const float *b_out, const float *noise_scale,
                             const float *noise_shift, const float *u, const float *prior_log_odds,
                             const air_hyper_t *hyper, float *stop, float *loss, int32_t *digits, float *fields,
                             float *theta, float *theta_inv, int64_t B, int HU, air_stream_t stream) {
  AIR_REQUIRE(B >= 0 && HU > 0, AIR_ERR_BAD_SHAPE, "heads_fwd: bad shape");
  if (B == 0) return AIR_OK;
  AIR_REQUIRE(hidden && w_out && b_out && noise_scale && noise_shift && u && prior_log_odds && hyper && stop && loss &&
                  digits && fields && theta && theta_inv,
              AIR_ERR_NULL, "heads_fwd: null pointer");
  const int64_t threads = B * 8;
  AIR_LAUNCH(heads_fwd_k, static_cast<unsigned>((threads + 255) / 256), 256, 0, ST(stream), 
      hidden, w_out, b_out, noise_scale, noise_shift, u, prior_log_odds, *hyper, stop, loss, digits, fields, theta,
      theta_inv, B, HU, (HU % 4 == 0 && aligned16(hidden) && aligned16(w_out)) ? 1 : 0);
  count_launch();
  return check_launch("heads_fwd");
}

extern "C" int64_t air_heads_bwd_workspace(int64_t B, int HU) {
  return ((B + kHeadsImgs - 1) / kHeadsImgs) * (7 * static_cast<int64_t>(HU) + 7);
}

extern "C" int air_heads_bwd(const float *hidden, const float *w_out, const float *noise_scale,
                             const float *noise_shift, const float *fields, const float *dtheta,
                             const float *dtheta_inv, const float *dz, const float *prior_log_odds,
                             const air_hyper_t *hyper, float dloss, float *dhidden, float *dw_out, float *db_out,
                             int accumulate, float *workspace, int64_t B, int HU, air_stream_t stream) {
  AIR_REQUIRE(B >= 0 && HU > 0, AIR_ERR_BAD_SHAPE, "heads_bwd: bad shape");
  if (B == 0) return AIR_OK;
  AIR_REQUIRE(hidden && w_out && noise_scale && noise_shift && fields && dtheta && dtheta_inv && dz && prior_log_odds &&
                  hyper && dhidden && workspace && ((dw_out != nullptr) == (db_out != nullptr)),
              AIR_ERR_NULL, "heads_bwd: null pointer");
  const int R = static_cast<int>((B + kHeadsImgs - 1) / kHeadsImgs);
  AIR_LAUNCH(heads_bwd_k, R, 256, 0, ST(stream), hidden, w_out, noise_scale, noise_shift, fields, dtheta, dtheta_inv, dz,
                                         prior_log_odds, *hyper, dloss, dhidden, workspace, B, HU);
  count_launch();
  int rc = check_launch("heads_bwd");
  if (rc || !dw_out) return rc;  // dw_out == NULL: the caller reduces the per-CTA partials itself (air_reduce_rows)
  const int n = 7 * HU;
  AIR_LAUNCH(reduce_rows_k, grid_for(static_cast<int64_t>(n) * 32, 256), 256, 0, ST(stream), workspace, R, n + 7, n, dw_out, accumulate);
  AIR_LAUNCH(reduce_rows_k, 1, 256, 0, ST(stream), workspace + n, R, n + 7, 7, db_out, accumulate);
  count_launch(2);
  return check_launch("heads_bwd reduce");
}

extern "C" int air_vae_latent_fwd(const float *ml, const float *noise, const air_hyper_t *hyper, float *sample,
                                  int ld_sample, float *fields, float *loss, int64_t B, int L, air_stream_t stream) {
  AIR_REQUIRE(B >= 0 && L > 0 && ld_sample >= L, AIR_ERR_BAD_SHAPE, "vae_latent_fwd: bad shape");
  if (B == 0) return AIR_OK;
  AIR_REQUIRE(ml && noise && hyper && sample && fields && loss, AIR_ERR_NULL, "vae_latent_fwd: null pointer");
  AIR_LAUNCH(vae_latent_fwd_k, static_cast<unsigned>((B * 32 + 255) / 256), 256, 0, ST(stream), ml, noise, *hyper, sample, ld_sample,
                                                                                       fields, loss, B, L);
  count_launch();
  return check_launch("vae_latent_fwd");
}

extern "C" int air_vae_latent_bwd(const float *ml, const float *noise, const float *dsample, const float *fields,
                                  const air_hyper_t *hyper, float dloss, float *dml, int64_t B, int L,
                                  air_stream_t stream) {
  AIR_REQUIRE(B >= 0 && L > 0, AIR_ERR_BAD_SHAPE, "vae_latent_bwd: bad shape");
  if (B == 0) return AIR_OK;
  AIR_REQUIRE(ml && noise && dsample && fields && hyper && dml, AIR_ERR_NULL, "vae_latent_bwd: null pointer");
  AIR_LAUNCH(vae_latent_bwd_k, grid_for(B * L, 256), 256, 0, ST(stream), ml, noise, dsample, fields, *hyper, dloss, dml, B, L);
  count_launch();
  return check_launch("vae_latent_bwd");
}

extern "C" int air_sigmoid_noise_fwd(const float *gen, const float *noise, float sd, float *out, int64_t n,
                                     air_stream_t stream) {
  AIR_REQUIRE(n >= 0, AIR_ERR_BAD_SHAPE, "sigmoid_noise_fwd: n < 0");
  if (n == 0) return AIR_OK;
  AIR_REQUIRE(gen && noise && out, AIR_ERR_NULL, "sigmoid_noise_fwd: null pointer");
  AIR_LAUNCH(sigmoid_noise_fwd_k, grid_for(n, 256, 16), 256, 0, ST(stream), gen, noise, sd, out, n);
  count_launch();
  return check_launch("sigmoid_noise_fwd");
}

extern "C" int air_sigmoid_bwd(const float *out, const float *dout, float *dgen, int64_t n, air_stream_t stream) {
  AIR_REQUIRE(n >= 0, AIR_ERR_BAD_SHAPE, "sigmoid_bwd: n < 0");
  if (n == 0) return AIR_OK;
  AIR_REQUIRE(out && dout && dgen, AIR_ERR_NULL, "sigmoid_bwd: null pointer");
  AIR_LAUNCH(sigmoid_bwd_k, grid_for(n, 256, 16), 256, 0, ST(stream), out, dout, dgen, n);
  count_launch();
  return check_launch("sigmoid_bwd");
}

extern "C" int air_bce_loss(const float *canvas, const float *x, float *recon, float *rec_loss, float *dcanvas,
                            float dscale, int64_t B, int N, air_stream_t stream) {
  AIR_REQUIRE(B >= 0 && N > 0 && B < (int64_t(1) << 31), AIR_ERR_BAD_SHAPE, "bce_loss: bad shape");
  if (B == 0) return AIR_OK;
  AIR_REQUIRE(canvas && x && rec_loss, AIR_ERR_NULL, "bce_loss: null pointer");
  if (N % 4 == 0 && aligned16(canvas) && aligned16(x) && (!recon || aligned16(recon)) && (!dcanvas || aligned16(dcanvas)))
    AIR_LAUNCH(bce_loss_k<4>, static_cast<unsigned>(B), 256, 0, ST(stream), canvas, x, recon, rec_loss, dcanvas, dscale, N);
  else
    AIR_LAUNCH(bce_loss_k<1>, static_cast<unsigned>(B), 256, 0, ST(stream), canvas, x, recon, rec_loss, dcanvas, dscale, N);
  count_launch();
  return check_launch("bce_loss");
}

extern "C" int air_finalize_loss(const float *running_loss, const float *rec_loss, const int32_t *digits,
                                 const int32_t *target, float *out, float *loss_per_item, int64_t B,
                                 air_stream_t stream) {
  AIR_REQUIRE(B > 0, AIR_ERR_BAD_SHAPE, "finalize_loss: B <= 0");
  AIR_REQUIRE(running_loss && rec_loss && digits && target && out, AIR_ERR_NULL, "finalize_loss: null pointer");
  AIR_LAUNCH(finalize_loss_k, 1, 1024, 0, ST(stream), running_loss, rec_loss, digits, target, out, loss_per_item, B);
  count_launch();
  return check_launch("finalize_loss");
}

extern "C" int64_t air_colsum_workspace(int64_t B, int N) { return static_cast<int64_t>(64) * N + 64 + 0 * B; }

extern "C" int air_colsum(const float *X, int ld, float *out, int accumulate, float *workspace, int64_t B, int N,
                          air_stream_t stream) {
  AIR_REQUIRE(B >= 0 && N > 0 && ld >= N, AIR_ERR_BAD_SHAPE, "colsum: bad shape");
  AIR_REQUIRE(X && out && workspace, AIR_ERR_NULL, "colsum: null pointer");
  if (B == 0) {
    if (!accumulate) cudaMemsetAsync(out, 0, sizeof(float) * N, ST(stream));
    return AIR_OK;
  }
  const int R = colsum_chunks(B, N);
  const int rows = static_cast<int>((B + R - 1) / R);
  dim3 grid((N + 31) / 32, R);
  AIR_LAUNCH(colsum_partial_k, grid, 256, 0, ST(stream), X, ld, workspace, B, N, rows);
  count_launch();
  int rc = check_launch("colsum_partial");
  if (rc) return rc;
  AIR_LAUNCH(reduce_rows_k, grid_for(static_cast<int64_t>(N) * 32, 256), 256, 0, ST(stream), workspace, R, N, N, out, accumulate);
  count_launch();
  return check_launch("colsum reduce");
}

extern "C" int64_t air_colsum_multi_workspace(const air_colsum_item_t *items, int n_items) {
  int64_t tot = 0;
  for (int i = 0; items && i < n_items; ++i) tot += static_cast<int64_t>(64) * items[i].N;
  return tot + 64;
}

extern "C" int air_colsum_multi(const air_colsum_item_t *items, int n_items, float *workspace, air_stream_t stream) {
  AIR_REQUIRE(items && n_items > 0 && n_items <= kMaxColsumItems, AIR_ERR_BAD_SHAPE, "colsum_multi: 1..%d items", kMaxColsumItems);
  AIR_REQUIRE(workspace, AIR_ERR_NULL, "colsum_multi: null workspace");
  ColsumBatch batch;
  int ctas = 0, ws = 0, outs = 0;
  for (int i = 0; i < n_items; ++i) {
    const air_colsum_item_t &a = items[i];
    AIR_REQUIRE(a.X && a.out && a.rows > 0 && a.N > 0 && a.ld >= a.N, AIR_ERR_BAD_SHAPE, "colsum_multi: bad item %d", i);
    ColsumItem &q = batch.it[i];
    q.X = a.X; q.out = a.out; q.rows = a.rows; q.ld = a.ld; q.N = a.N; q.accumulate = a.accumulate;
    q.R = colsum_chunks(a.rows, a.N);
    q.rows_per_chunk = static_cast<int>((a.rows + q.R - 1) / q.R);
    q.cta0 = ctas; q.ws0 = ws; q.out0 = outs;
    ctas += q.R * ((a.N + 31) / 32);
    ws += q.R * a.N;
    outs += a.N;
  }
  batch.n = n_items;
  AIR_LAUNCH(colsum_multi_partial_k, ctas, 256, 0, ST(stream), batch, workspace);
  AIR_LAUNCH(colsum_multi_reduce_k, grid_for(static_cast<int64_t>(outs) * 32, 256), 256, 0, ST(stream), batch,
             static_cast<const float *>(workspace), outs);
  count_launch(2);
  return check_launch("colsum_multi");
}

extern "C" int air_synth_canvases_ex(uint64_t seed, int64_t first_index, float *images, int32_t *counts, int32_t *positions,
                                     int32_t *boxes, int64_t B, int canvas_size, int max_digits, air_stream_t stream) {
  AIR_REQUIRE(B >= 0 && canvas_size >= 24 && canvas_size <= 104 && max_digits >= 0 && first_index >= 0, AIR_ERR_BAD_SHAPE,
              "synth_canvases: bad arguments (24 <= canvas_size <= 104)");
  if (B == 0) return AIR_OK;
  AIR_REQUIRE(images && counts, AIR_ERR_NULL, "synth_canvases: null pointer");
  AIR_LAUNCH(synth_canvases_k, static_cast<unsigned>(B), 256, static_cast<size_t>(canvas_size) * canvas_size * 4, ST(stream), seed,
             first_index, images, counts, positions, boxes, canvas_size, max_digits);
  count_launch();
  return check_launch("synth_canvases");
}

extern "C" int air_synth_canvases(uint64_t seed, int64_t first_index, float *images, int32_t *counts, int64_t B,
                                  int canvas_size, int max_digits, air_stream_t stream) {
  return air_synth_canvases_ex(seed, first_index, images, counts, nullptr, nullptr, B, canvas_size, max_digits, stream);
}

extern "C" int air_noise_fill(air_rng_state_t *state, float *scale, float *shift, float *vae_latent, float *concrete_u,
                              int64_t TB, int L, air_stream_t stream) {
  AIR_REQUIRE(TB >= 0 && L >= 0, AIR_ERR_BAD_SHAPE, "noise_fill: bad shape");
  AIR_REQUIRE(state, AIR_ERR_NULL, "noise_fill: null RNG state");
  const int64_t blocks = ((scale ? TB : 0) + 3) / 4 + ((shift ? 2 * TB : 0) + 3) / 4 + ((vae_latent ? TB * L : 0) + 3) / 4 +
                         ((concrete_u ? TB : 0) + 3) / 4;
  AIR_LAUNCH(noise_fill_k, grid_for(std::max<int64_t>(blocks, 1), 256), 256, 0, ST(stream),
             reinterpret_cast<unsigned long long *>(state), scale, shift, vae_latent, concrete_u, TB, L);
  count_launch();
  return check_launch("noise_fill");
}

static int rng_samples(const air_rng_state_t *state, int rng_stream, int uniform, float *out, int64_t n, air_stream_t stream) {
  AIR_REQUIRE(n >= 0 && rng_stream > 0, AIR_ERR_BAD_SHAPE, "rng_samples: bad arguments");
  if (n == 0) return AIR_OK;
  AIR_REQUIRE(state && out, AIR_ERR_NULL, "rng_samples: null pointer");
  AIR_LAUNCH(rng_samples_k, grid_for((n + 3) / 4, 256), 256, 0, ST(stream), reinterpret_cast<const unsigned long long *>(state),
             static_cast<uint32_t>(rng_stream), uniform, out, n);
  count_launch();
  return check_launch("rng_samples");
}

extern "C" int air_rng_normals(const air_rng_state_t *state, int rng_stream, float *out, int64_t n, air_stream_t stream) {
  return rng_samples(state, rng_stream, 0, out, n, stream);
}

extern "C" int air_rng_uniforms(const air_rng_state_t *state, int rng_stream, float *out, int64_t n, air_stream_t stream) {
  return rng_samples(state, rng_stream, 1, out, n, stream);
}

extern "C" int air_zero_buffers(void *const *buffers, const int64_t *nbytes, int n, air_stream_t stream) {
  AIR_REQUIRE(n >= 0 && n <= 8, AIR_ERR_BAD_SHAPE, "zero_buffers: at most 8 buffers per call");
  if (n == 0) return AIR_OK;
  AIR_REQUIRE(buffers && nbytes, AIR_ERR_NULL, "zero_buffers: null pointer");
  ZeroBatch zb;
  zb.n = n;
  int64_t most = 0;
  for (int q = 0; q < n; ++q) {
    AIR_REQUIRE(buffers[q] && aligned16(buffers[q]) && nbytes[q] >= 0 && nbytes[q] % 16 == 0, AIR_ERR_BAD_ALIGN,
                "zero_buffers: buffers must be 16-byte aligned with sizes that are multiples of 16 bytes");
    zb.ptr[q] = buffers[q];
    zb.n16[q] = nbytes[q] / 16;
    most = std::max(most, zb.n16[q]);
  }
  AIR_LAUNCH(zero_many_k, grid_for(std::max<int64_t>(most, 1), 256), 256, 0, ST(stream), zb);
  count_launch();
  return check_launch("zero_buffers");
}

extern "C" int air_expand_u8(const uint8_t *src, float *dst, int64_t n, air_stream_t stream) {
  AIR_REQUIRE(n >= 0, AIR_ERR_BAD_SHAPE, "expand_u8: n < 0");
  if (n == 0) return AIR_OK;
  AIR_REQUIRE(src && dst, AIR_ERR_NULL, "expand_u8: null pointer");
  AIR_REQUIRE(aligned16(src) && aligned16(dst), AIR_ERR_BAD_ALIGN, "expand_u8: 16-byte aligned buffers required");
  AIR_LAUNCH(expand_u8_k, grid_for(std::max<int64_t>(n >> 4, 1), 256), 256, 0, ST(stream), src, dst, n);
  count_launch();
  return check_launch("expand_u8");
}

extern "C" int air_reduce_rows(const float *partials, int R, int stride, int n, float *out, int accumulate, air_stream_t stream) {
  AIR_REQUIRE(partials && out && R > 0 && n > 0 && stride >= n, AIR_ERR_BAD_SHAPE, "reduce_rows: bad arguments");
  return reduce_rows_launch(partials, R, stride, n, out, accumulate, ST(stream));
}

extern "C" int64_t air_adam_workspace(int64_t n) { return kAdamPartials + 8 + 0 * n; }

extern "C" int air_adam_step(float *params, const float *grads, float *m, float *v, float *state, float clip_norm,
                             float beta1, float beta2, float epsilon, float grad_scale, float *workspace, int64_t n,
                             air_stream_t stream) {
  return air_adam_step_ex(params, grads, m, v, state, clip_norm, beta1, beta2, epsilon, grad_scale, workspace, n, 0, stream);
}

extern "C" int air_adam_step_ex(float *params, const float *grads, float *m, float *v, float *state, float clip_norm,
                                float beta1, float beta2, float epsilon, float grad_scale, float *workspace, int64_t n,
                                int flags, air_stream_t stream) {
  AIR_REQUIRE(n > 0, AIR_ERR_BAD_SHAPE, "adam_step: n <= 0");
  AIR_REQUIRE(params && grads && m && v && state && workspace, AIR_ERR_NULL, "adam_step: null pointer");
  const int np = static_cast<int>(std::min<int64_t>(kAdamPartials, (n + 1023) / 1024));
  AIR_LAUNCH(sumsq_partial_k, np, 256, 0, ST(stream), grads, grad_scale, workspace, n);
  AIR_LAUNCH(adam_apply_k, grid_for(n, 256, 4), 256, 0, ST(stream), params, grads, m, v, state, workspace, np, clip_norm, beta1,
                                                            beta2, epsilon, grad_scale, workspace + kAdamPartials, n,
             (aligned16(params) && aligned16(grads) && aligned16(m) && aligned16(v)) ? 1 : 0, flags & AIR_ADAM_SKIP_NONFINITE);
  AIR_LAUNCH(adam_advance_k, 1, 1, 0, ST(stream), state, beta1, beta2, workspace + kAdamPartials, flags & AIR_ADAM_SKIP_NONFINITE);
  count_launch(3);
  return check_launch("adam_step");
}

extern "C" int air_anneal(const float *state, float init, float factor, float iters, int staircase, float vmin,
                          float vmax, int take_log, float *out, air_stream_t stream) {
  AIR_REQUIRE(state && out, AIR_ERR_NULL, "anneal: null pointer");
  AIR_LAUNCH(anneal_k, 1, 1, 0, ST(stream), state, init, factor, iters, staircase, vmin, vmax, take_log, out);
  count_launch();
  return check_launch("anneal");
}
