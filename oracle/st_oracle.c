/* CPU oracle (plain C) for the Spatial-Transformer / canvas / Concrete arithmetic of the
 * AIR hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, smoke() and
 * bench.py's cpu_baseline leg may load this library.
 *
 * PARITY PIN (see oracle/air_oracle.py header): TensorFlow cannot run here and the reference
 * has no golden vectors; this file is a second, independent restatement of the same
 * reference lines, used to cross-check the torch oracle bit-for-bit.  The ST forward of
 * this file reproduces the reference's own graph (model/air-model.meta run by
 * oracle/tfgraph) bit for bit: tests/test_reference_graph.py.  The cnn front-end at the
 * end of the file is not in that graph and is restated-only.
 *
 * Build with -ffp-contract=off: TF 1.3 executes one op at a time, so every product and
 * sum below is rounded separately (no FMA), except oracle_gemm_seq_fma which *defines*
 * the k-sequential FMA order of the CUDA "exact" GEMM mode.
 *
 * Reference lines (relative to the reference checkout):
 *   air/transformer.py:119-136  _meshgrid        -> linspace_tf()
 *   air/transformer.py:138-171  _transform       -> oracle_st_forward()
 *   air/transformer.py:56-117   _interpolate     -> oracle_st_forward()/oracle_st_backward()
 *   air/air_model.py:429-439    canvas update    -> oracle_canvas_update()
 *   air/concrete.py:20-43, air/air_model.py:380-427 -> oracle_concrete_step()
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EPSF 1e-9f

static float linspace_tf(int i, int n) {
  /* tf.linspace(-1, 1, n): start + step*i, step = (stop-start)/(n-1), all fp32 */
  if (n == 1) return -1.0f;
  float step = (1.0f - (-1.0f)) / (float)(n - 1);
  return -1.0f + step * (float)i;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

typedef struct {
  int x0, x1, y0, y1;
  float x, y; /* unclipped pixel coordinates */
} samp_t;

static samp_t sample_point(const float *th, float xt, float yt, int H, int W) {
  samp_t s;
  /* transformer.py:159 BatchMatMul k=3: fl(fl(fl(a0 b0)+fl(a1 b1))+fl(a2 b2)) */
  float xs = (th[0] * xt + th[1] * yt) + th[2] * 1.0f;
  float ys = (th[3] * xt + th[4] * yt) + th[5] * 1.0f;
  /* :75-76 */
  float wf = (float)W - 1.001f, hf = (float)H - 1.001f;
  s.x = (xs + 1.0f) * wf / 2.0f;
  s.y = (ys + 1.0f) * hf / 2.0f;
  /* :79-87; floor of a huge/NaN coordinate is UB in TF as well -- clamp through double */
  double fx = floor((double)s.x), fy = floor((double)s.y);
  if (!(fx > -2.0e9)) fx = -2.0e9;
  if (fx > 2.0e9) fx = 2.0e9;
  if (!(fy > -2.0e9)) fy = -2.0e9;
  if (fy > 2.0e9) fy = 2.0e9;
  int x0 = (int)fx, y0 = (int)fy;
  s.x0 = clampi(x0, 0, W - 1);
  s.x1 = clampi(x0 + 1, 0, W - 1);
  s.y0 = clampi(y0, 0, H - 1);
  s.y1 = clampi(y0 + 1, 0, H - 1);
  return s;
}

/* U [B,H,W,C] NHWC, theta [B,6], out [B,oh,ow,C] */
void oracle_st_forward(const float *U, const float *theta, float *out, int64_t B, int H, int W, int C,
                       int oh, int ow) {
  for (int64_t b = 0; b < B; ++b) {
    const float *th = theta + b * 6;
    const float *im = U + b * (int64_t)H * W * C;
    for (int r = 0; r < oh; ++r) {
      float yt = linspace_tf(r, oh);
      for (int c = 0; c < ow; ++c) {
        float xt = linspace_tf(c, ow);
        samp_t s = sample_point(th, xt, yt, H, W);
        float x0f = (float)s.x0, x1f = (float)s.x1, y0f = (float)s.y0, y1f = (float)s.y1;
        /* :108-115 */
        float wa = (x1f - s.x) * (y1f - s.y);
        float wb = (x1f - s.x) * (s.y - y0f);
        float wc = (s.x - x0f) * (y1f - s.y);
        float wd = (s.x - x0f) * (s.y - y0f);
        const float *pa = im + ((int64_t)s.y0 * W + s.x0) * C;
        const float *pb = im + ((int64_t)s.y1 * W + s.x0) * C;
        const float *pc = im + ((int64_t)s.y0 * W + s.x1) * C;
        const float *pd = im + ((int64_t)s.y1 * W + s.x1) * C;
        float *o = out + ((b * oh + r) * (int64_t)ow + c) * C;
        for (int ch = 0; ch < C; ++ch) {
          /* :116 add_n left to right */
          o[ch] = ((wa * pa[ch] + wb * pb[ch]) + wc * pc[ch]) + wd * pd[ch];
        }
      }
    }
  }
}

/* TF autodiff of the graph above.  dU may be NULL (crop ST: U is data).
 * dU accumulation order: corner a for every output pixel, then b, c, d (four
 * IndexedSlices concatenated into one sequential UnsortedSegmentSum). */
void oracle_st_backward(const float *U, const float *theta, const float *dout, float *dU, float *dtheta,
                        int64_t B, int H, int W, int C, int oh, int ow) {
  if (dU) memset(dU, 0, sizeof(float) * (size_t)(B * H * W * C));
  float wf = (float)W - 1.001f, hf = (float)H - 1.001f;
  for (int64_t b = 0; b < B; ++b) {
    const float *th = theta + b * 6;
    const float *im = U + b * (int64_t)H * W * C;
    float *dim = dU ? dU + b * (int64_t)H * W * C : NULL;
    float acc[6] = {0, 0, 0, 0, 0, 0};
    for (int corner = 0; corner < (dU ? 4 : 0); ++corner) {
      for (int r = 0; r < oh; ++r) {
        float yt = linspace_tf(r, oh);
        for (int c = 0; c < ow; ++c) {
          float xt = linspace_tf(c, ow);
          samp_t s = sample_point(th, xt, yt, H, W);
          float x0f = (float)s.x0, x1f = (float)s.x1, y0f = (float)s.y0, y1f = (float)s.y1;
          float w;
          int yy, xx;
          switch (corner) {
            case 0: w = (x1f - s.x) * (y1f - s.y); yy = s.y0; xx = s.x0; break;
            case 1: w = (x1f - s.x) * (s.y - y0f); yy = s.y1; xx = s.x0; break;
            case 2: w = (s.x - x0f) * (y1f - s.y); yy = s.y0; xx = s.x1; break;
            default: w = (s.x - x0f) * (s.y - y0f); yy = s.y1; xx = s.x1; break;
          }
          const float *g = dout + ((b * oh + r) * (int64_t)ow + c) * C;
          float *d = dim + ((int64_t)yy * W + xx) * C;
          for (int ch = 0; ch < C; ++ch) d[ch] += w * g[ch];
        }
      }
    }
    for (int r = 0; r < oh; ++r) {
      float yt = linspace_tf(r, oh);
      for (int c = 0; c < ow; ++c) {
        float xt = linspace_tf(c, ow);
        samp_t s = sample_point(th, xt, yt, H, W);
        float x0f = (float)s.x0, x1f = (float)s.x1, y0f = (float)s.y0, y1f = (float)s.y1;
        const float *pa = im + ((int64_t)s.y0 * W + s.x0) * C;
        const float *pb = im + ((int64_t)s.y1 * W + s.x0) * C;
        const float *pc = im + ((int64_t)s.y0 * W + s.x1) * C;
        const float *pd = im + ((int64_t)s.y1 * W + s.x1) * C;
        const float *g = dout + ((b * oh + r) * (int64_t)ow + c) * C;
        float dwa = 0, dwb = 0, dwc = 0, dwd = 0; /* reduce over channels (expand_dims bcast) */
        for (int ch = 0; ch < C; ++ch) {
          dwa += g[ch] * pa[ch];
          dwb += g[ch] * pb[ch];
          dwc += g[ch] * pc[ch];
          dwd += g[ch] * pd[ch];
        }
        float wx1 = x1f - s.x, wx0 = s.x - x0f, wy1 = y1f - s.y, wy0 = s.y - y0f;
        /* d/dx: (x1f-x) has -1, (x-x0f) has +1; no gradient through floor/clip/cast */
        float dx = ((-(dwa * wy1) - dwb * wy0) + dwc * wy1) + dwd * wy0;
        float dy = ((-(dwa * wx1) + dwb * wx1) - dwc * wx0) + dwd * wx0;
        float dxs = dx / 2.0f * wf;
        float dys = dy / 2.0f * hf;
        acc[0] += dxs * xt; acc[1] += dxs * yt; acc[2] += dxs;
        acc[3] += dys * xt; acc[4] += dys * yt; acc[5] += dys;
      }
    }
    for (int k = 0; k < 6; ++k) dtheta[b * 6 + k] = acc[k];
  }
}

/* air_model.py:429-439: canvas += where(stop_new < thr, z[:,None]*window_recon, 0) */
void oracle_canvas_update(const float *canvas_in, const float *window_recon, const float *z,
                          const float *stop_new, float thr, float *canvas_out, int64_t B, int N) {
  for (int64_t b = 0; b < B; ++b) {
    int live = stop_new[b] < thr;
    for (int i = 0; i < N; ++i) {
      float add = live ? z[b] * window_recon[b * N + i] : 0.0f;
      canvas_out[b * N + i] = canvas_in[b * N + i] + add;
    }
  }
}

static float log_density(float y, float alpha, float tau) {
  /* concrete.py:35-37: log(tau+eps) - y*tau + alpha - 2*log(1 + exp(-y*tau + alpha) + eps) */
  float yt = y * tau;
  return logf(tau + EPSF) - yt + alpha - 2.0f * logf(1.0f + expf(-yt + alpha) + EPSF);
}

/* One z_pres / ACT step: air_model.py:380-427 with concrete.py:20-43.
 * in : log_odds[B], u[B], stop_prev[B], loss_prev[B], digits_prev[B]
 * out: y (pre-sigmoid), z, z_prob, kl, stop_new, loss_new (+= masked kl), digits_new */
void oracle_concrete_step(const float *log_odds, const float *u, const float *stop_prev,
                          const float *loss_prev, const int32_t *digits_prev, float prior_log_odds,
                          float temperature, float thr, int train, float *y, float *z, float *z_prob,
                          float *kl, float *stop_new, float *loss_new, int32_t *digits_new, int64_t B) {
  for (int64_t b = 0; b < B; ++b) {
    float noise = logf(u[b] + EPSF) - logf(1.0f - u[b] + EPSF);
    float yy = (log_odds[b] + noise) / temperature;
    float zz = 1.0f / (1.0f + expf(-yy));
    if (!train) zz = nearbyintf(zz); /* tf.round, half-to-even */
    float k = log_density(yy, log_odds[b], temperature) - log_density(yy, prior_log_odds, temperature);
    y[b] = yy;
    z[b] = zz;
    z_prob[b] = 1.0f / (1.0f + expf(-log_odds[b]));
    kl[b] = k;
    loss_new[b] = loss_prev[b] + (stop_prev[b] < thr ? k : 0.0f);
    float sn = stop_prev[b] + (1.0f - zz);
    stop_new[b] = sn;
    digits_new[b] = digits_prev[b] + (sn < thr ? 1 : 0);
  }
}

/* Defines the "exact" GEMM mode: C[m,n] = bias[n] + (Cinit[m,n] + sum_k fma(A[m,k], Bm[k,n]))
 * with the k loop strictly sequential and one FMA per term.  A [M,K] row-major (lda),
 * Bm [K,N] row-major (ldb).  Cinit/bias may be NULL. */
void oracle_gemm_seq_fma(const float *A, const float *Bm, const float *Cinit, const float *bias, float *Cout,
                         int64_t M, int N, int K, int lda, int ldb, int ldc) {
  float *acc = (float *)malloc(sizeof(float) * (size_t)N);
  for (int64_t m = 0; m < M; ++m) {
    for (int n = 0; n < N; ++n) acc[n] = Cinit ? Cinit[m * ldc + n] : 0.0f;
    for (int k = 0; k < K; ++k) {
      float a = A[m * lda + k];
      const float *brow = Bm + (int64_t)k * ldb;
      for (int n = 0; n < N; ++n) acc[n] = fmaf(a, brow[n], acc[n]);
    }
    for (int n = 0; n < N; ++n) Cout[m * ldc + n] = bias ? acc[n] + bias[n] : acc[n];
  }
  free(acc);
}

/* ---- CNN front-end of AIRModel(cnn=True): air_model.py:510-535 -----------------------------------------------
 * tf.layers.conv2d(kernel 5x5, strides 1, padding "same", activation relu) optionally followed by
 * tf.layers.max_pooling2d(pool 2, strides 2, "valid"); NHWC activations, HWIO kernels, 'same' = two zero pixels on
 * every side for a 5x5 kernel.  Plain loops in (ky, kx, ci) order, multiply and add rounded separately.
 * Second, independent restatement next to oracle/air_oracle.py::cnn_frontend (which goes through torch's conv2d). */
static float conv_at(const float *in, const float *w, const float *bias, int H, int W, int cin, int cout, int y, int x,
                     int co) {
  float acc = bias[co];
  for (int ky = 0; ky < 5; ++ky)
    for (int kx = 0; kx < 5; ++kx) {
      const int yy = y + ky - 2, xx = x + kx - 2;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      for (int ci = 0; ci < cin; ++ci) {
        const float prod = in[(yy * W + xx) * cin + ci] * w[((ky * 5 + kx) * cin + ci) * cout + co];
        acc = acc + prod;
      }
    }
  return acc > 0.0f ? acc : 0.0f; /* relu */
}

void oracle_conv5x5_relu_pool(const float *in, const float *w, const float *bias, float *out, int64_t B, int H, int W,
                              int cin, int cout, int pool) {
  const int PH = pool ? H / 2 : H, PW = pool ? W / 2 : W;
  for (int64_t b = 0; b < B; ++b) {
    const float *ib = in + b * H * W * cin;
    float *ob = out + b * PH * PW * cout;
    for (int py = 0; py < PH; ++py)
      for (int px = 0; px < PW; ++px)
        for (int co = 0; co < cout; ++co) {
          float m;
          if (!pool) {
            m = conv_at(ib, w, bias, H, W, cin, cout, py, px, co);
          } else {
            m = conv_at(ib, w, bias, H, W, cin, cout, 2 * py, 2 * px, co);
            for (int p = 1; p < 4; ++p) {
              const float v = conv_at(ib, w, bias, H, W, cin, cout, 2 * py + (p >> 1), 2 * px + (p & 1), co);
              if (v > m) m = v;
            }
          }
          ob[(py * PW + px) * cout + co] = m;
        }
  }
}

/* Backward: ReluGrad (passes where the layer output is > 0), MaxPoolGrad (first maximum of the window in row-major
 * order), Conv2DBackpropFilter / Conv2DBackpropInput.  dw [5,5,cin,cout], db [cout] and din (if non-NULL) are
 * overwritten.  Sums are accumulated in double and rounded once (this is a checker, not a bit-level statement). */
void oracle_conv5x5_relu_pool_bwd(const float *in, const float *w, const float *bias, const float *dout, float *din,
                                  float *dw, float *db, int64_t B, int H, int W, int cin, int cout, int pool) {
  const int PH = pool ? H / 2 : H, PW = pool ? W / 2 : W;
  const int nw = 25 * cin * cout;
  double *aw = (double *)calloc((size_t)nw + cout, sizeof(double));
  double *ai = din ? (double *)calloc((size_t)(B * H * W * cin), sizeof(double)) : NULL;
  for (int64_t b = 0; b < B; ++b) {
    const float *ib = in + b * H * W * cin;
    for (int py = 0; py < PH; ++py)
      for (int px = 0; px < PW; ++px)
        for (int co = 0; co < cout; ++co) {
          int y = py, x = px;
          float m = 0.0f;
          if (!pool) {
            m = conv_at(ib, w, bias, H, W, cin, cout, py, px, co);
          } else {
            m = conv_at(ib, w, bias, H, W, cin, cout, 2 * py, 2 * px, co);
            y = 2 * py; x = 2 * px;
            for (int p = 1; p < 4; ++p) {
              const float v = conv_at(ib, w, bias, H, W, cin, cout, 2 * py + (p >> 1), 2 * px + (p & 1), co);
              if (v > m) { m = v; y = 2 * py + (p >> 1); x = 2 * px + (p & 1); }
            }
          }
          if (!(m > 0.0f)) continue;
          const double g = dout[((b * PH + py) * PW + px) * cout + co];
          aw[nw + co] += g;
          for (int ky = 0; ky < 5; ++ky)
            for (int kx = 0; kx < 5; ++kx) {
              const int yy = y + ky - 2, xx = x + kx - 2;
              if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
              for (int ci = 0; ci < cin; ++ci) {
                aw[((ky * 5 + kx) * cin + ci) * cout + co] += g * ib[(yy * W + xx) * cin + ci];
                if (ai) ai[((b * H + yy) * W + xx) * cin + ci] += g * w[((ky * 5 + kx) * cin + ci) * cout + co];
              }
            }
        }
  }
  for (int e = 0; e < nw; ++e) dw[e] = (float)aw[e];
  for (int co = 0; co < cout; ++co) db[co] = (float)aw[nw + co];
  if (ai) {
    for (int64_t e = 0; e < B * H * W * cin; ++e) din[e] = (float)ai[e];
    free(ai);
  }
  free(aw);
}
