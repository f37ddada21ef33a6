// CNN front-end of AIRModel(cnn=True): air/air_model.py:510-535 of the reference.
//   conv1 5x5x1x8 'same' + ReLU -> max-pool 2x2/2 -> conv2 5x5x8x8 + ReLU -> max-pool 2x2/2 (valid: 25 -> 12)
//   -> conv3 5x5x8x8 + ReLU -> reshape [B, 12*12*8] (NHWC order) = the LSTM input.
// Direct convolutions on the FP32 pipe (8 output channels: far too narrow for a tcgen05 tile), NHWC, one
// CTA per image with the zero-padded input tile and the weights in shared memory.  The pooled layers store a
// one-byte argmax per pooled element, so the backward never needs the un-pooled activations.
//   forward : thread = output (pooled) pixel, all 8 output channels in registers; weights are warp-broadcast
//             128-bit shared loads, ~10 FMAs per shared load.
//   backward: dW/db accumulate in registers across the images of a CTA (thread = one (ky,kx,ci) tap, 8 output
//             channels), written once per CTA and combined by the fixed-order reduce_rows kernel -> deterministic;
//             dX is the gather form over a dense, zero-padded d(conv) tile rebuilt in shared memory.
// Gradient semantics follow TF: ReluGrad passes where the output is > 0, MaxPoolGrad routes to the first maximum
// of the window in row-major order (ties only occur at 0, where ReluGrad blocks the gradient anyway).
#include <algorithm>
#include <cstdint>

#include "air_common.cuh"

namespace air {


// model_ops.cu: out[e] (+)= sum_r partials[r * stride + e], fixed order (deterministic)
int reduce_rows_launch(const float *partials, int R, int stride, int n, float *out, int accumulate, cudaStream_t s);

template <int CIN, int H, int W>
__device__ __forceinline__ void load_padded(const float *__restrict__ src, float *sIn, int tid, int nthreads) {
  constexpr int IW = W + 4, IH = H + 4;
  for (int e = tid; e < IH * IW * CIN; e += nthreads) {
    const int ci = e % CIN, p = e / CIN;
    const int x = p % IW - 2, y = p / IW - 2;
    sIn[e] = (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(src + (y * W + x) * CIN + ci) : 0.0f;
  }
}

template <int CIN, int COUT, int H, int W, bool POOL, int THREADS>
__global__ void __launch_bounds__(THREADS)
    conv5x5_fwd_k(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ bias,
                  float *__restrict__ out, uint8_t *__restrict__ arg, int64_t B) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  static_assert(COUT % 4 == 0, "output channels travel as float4 / packed argmax bytes");
  constexpr int IW = W + 4, IH = H + 4, PH = POOL ? H / 2 : H, PW = POOL ? W / 2 : W, NP = POOL ? 4 : 1, CV = COUT / 4;
  extern __shared__ __align__(16) float smem[];
  float *sIn = smem;                                   // [IH][IW][CIN] zero-padded input tile
  float *sW = sIn + ((IH * IW * CIN + 3) & ~3);        // [25*CIN][COUT]  (HWIO, as stored)
  float *sB = sW + 25 * CIN * COUT;                    // [COUT]
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  load_padded<CIN, H, W>(in + b * H * W * CIN, sIn, tid, THREADS);
  for (int e = tid; e < 25 * CIN * COUT; e += THREADS) sW[e] = __ldg(w + e);
  if (tid < COUT) sB[tid] = __ldg(bias + tid);
  __syncthreads();
  for (int item = tid; item < PH * PW; item += THREADS) {
    const int py = item / PW, px = item - py * PW;
    const int y0 = POOL ? 2 * py : py, x0 = POOL ? 2 * px : px;
    float acc[NP][COUT];
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
      for (int co = 0; co < COUT; ++co) acc[p][co] = sB[co];
#pragma unroll 1
    for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          float v[NP];
#pragma unroll
          for (int p = 0; p < NP; ++p) v[p] = sIn[((y0 + (p >> 1) + ky) * IW + (x0 + (p & 1) + kx)) * CIN + ci];
#pragma unroll
          for (int q = 0; q < CV; ++q) {
            const float4 wq = *reinterpret_cast<const float4 *>(sW + ((ky * 5 + kx) * CIN + ci) * COUT + 4 * q);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              acc[p][4 * q + 0] = fmaf(v[p], wq.x, acc[p][4 * q + 0]); acc[p][4 * q + 1] = fmaf(v[p], wq.y, acc[p][4 * q + 1]);
              acc[p][4 * q + 2] = fmaf(v[p], wq.z, acc[p][4 * q + 2]); acc[p][4 * q + 3] = fmaf(v[p], wq.w, acc[p][4 * q + 3]);
            }
          }
        }
      }
    }
    float r[COUT];
    uint8_t a[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      float m = fmaxf(acc[0][co], 0.0f);  // ReLU, then max-pool (first maximum wins)
      int am = 0;
#pragma unroll
      for (int p = 1; p < NP; ++p) {
        const float v = fmaxf(acc[p][co], 0.0f);
        if (v > m) { m = v; am = p; }
      }
      r[co] = m;
      a[co] = static_cast<uint8_t>(am);
    }
    float *o = out + (b * PH * PW + item) * COUT;
#pragma unroll
    for (int q = 0; q < CV; ++q) *reinterpret_cast<float4 *>(o + 4 * q) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    if (POOL) {
      uint32_t *ao = reinterpret_cast<uint32_t *>(arg + (b * PH * PW + item) * COUT);
#pragma unroll
      for (int q = 0; q < CV; ++q) ao[q] = a[4 * q] | (a[4 * q + 1] << 8) | (a[4 * q + 2] << 16) | (a[4 * q + 3] << 24);
    }
  }
}

template <int CIN, int COUT, int H, int W, bool POOL>
constexpr size_t conv_fwd_smem() {
  return (static_cast<size_t>(((H + 4) * (W + 4) * CIN + 3) & ~3) + 25 * CIN * COUT + COUT) * sizeof(float);
}

template <int CIN, int COUT, int H, int W, bool POOL, bool NEED_DX>
constexpr size_t conv_bwd_smem() {
  constexpr size_t pin = ((H + 4) * (W + 4) * CIN + 3) & ~3;
  constexpr size_t pd = NEED_DX ? (H + 4) * (W + 4) * COUT : 0;
  constexpr size_t pg = (POOL ? (H / 2) * (W / 2) : H * W) * COUT;
  constexpr size_t pw = NEED_DX ? 25 * CIN * COUT : 0;
  return (pin + pd + pg + pw) * sizeof(float) + (POOL ? pg : 0);
}

// GS threads form a group that owns the 25*CIN weight taps + 8 biases; the THREADS / GS groups of a CTA split the
// output pixels among themselves and write one partial row each (fixed assignment -> deterministic).
template <int CIN, int COUT, int H, int W, bool POOL, bool NEED_DX, int THREADS, int GS>
__global__ void __launch_bounds__(THREADS)
    conv5x5_bwd_k(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ out,
                  const uint8_t *__restrict__ arg, const float *__restrict__ dout, float *__restrict__ din,
                  float *__restrict__ partials, int64_t B) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  static_assert(25 * CIN + COUT <= GS && THREADS % GS == 0, "one thread per weight tap + one per bias in every group");
  static_assert(COUT % 4 == 0, "output channels travel as float4");
  constexpr int NG = THREADS / GS, CV = COUT / 4;
  constexpr int IW = W + 4, IH = H + 4, PH = POOL ? H / 2 : H, PW = POOL ? W / 2 : W, NPIX = PH * PW;
  extern __shared__ __align__(16) float smem[];
  float *sIn = smem;                                              // [IH][IW][CIN]  zero-padded layer input
  float *sD = sIn + ((IH * IW * CIN + 3) & ~3);                   // [IH][IW][COUT] zero-padded dense d(conv)  (NEED_DX)
  float *sG = sD + (NEED_DX ? IH * IW * COUT : 0);                // [NPIX][COUT]   d(out) * (out > 0)
  float *sW = sG + NPIX * COUT;                                   // [25*CIN][COUT] (NEED_DX)
  uint8_t *sA = reinterpret_cast<uint8_t *>(sW + (NEED_DX ? 25 * CIN * COUT : 0));  // [NPIX][COUT] argmax (POOL)
  const int tid = threadIdx.x;
  if (NEED_DX)
    for (int e = tid; e < 25 * CIN * COUT; e += THREADS) sW[e] = __ldg(w + e);
  // this thread's weight tap (ky, kx, ci) and its COUT gradient accumulators, or a bias accumulator
  const int group = tid / GS, lt = tid - group * GS;
  const bool tap = lt < 25 * CIN, isb = lt >= 25 * CIN && lt < 25 * CIN + COUT;
  const int ci = lt % CIN, kk = lt / CIN, ky = kk / 5, kx = kk - ky * 5;
  float gw[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) gw[co] = 0.0f;
  float gb = 0.0f;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();  // the previous image's tiles are no longer read
    load_padded<CIN, H, W>(in + b * H * W * CIN, sIn, tid, THREADS);
    for (int e = tid; e < NPIX * COUT; e += THREADS) {
      const int64_t o = b * NPIX * COUT + e;
      sG[e] = out[o] > 0.0f ? dout[o] : 0.0f;  // ReluGrad
      if (POOL) sA[e] = arg[o];
    }
    if (NEED_DX)
      for (int e = tid; e < IH * IW * COUT; e += THREADS) sD[e] = 0.0f;
    __syncthreads();
    if (NEED_DX) {  // un-pool: scatter the pooled gradients to their argmax positions (distinct targets)
      for (int e = tid; e < NPIX * COUT; e += THREADS) {
        const int pp = e / COUT, co = e - pp * COUT;
        const int py = pp / PW, px = pp - py * PW;
        const int a = POOL ? sA[e] : 0;
        const int y = POOL ? 2 * py + (a >> 1) : py, x = POOL ? 2 * px + (a & 1) : px;
        sD[((y + 2) * IW + (x + 2)) * COUT + co] = sG[e];
      }
      __syncthreads();
    }
    if (tap && NEED_DX) {
      // dW[ky][kx][ci][:] += sum over conv pixels of in[y+ky-2][x+kx-2][ci] * d(conv)[y][x][:], from the dense
      // (un-pooled) tile that the dX pass needs anyway: one input load + COUT/4 broadcast 128-bit loads per COUT FMAs
      constexpr int CH = POOL ? 2 * PH : H, CW = POOL ? 2 * PW : W;  // conv pixels that can carry a gradient
      for (int q = group; q < CH * CW; q += NG) {
        const int y = q / CW, x = q - y * CW;
        const float *dp = sD + ((y + 2) * IW + (x + 2)) * COUT;
        const float v = sIn[((y + ky) * IW + (x + kx)) * CIN + ci];
#pragma unroll
        for (int u = 0; u < CV; ++u) {
          const float4 g4 = *reinterpret_cast<const float4 *>(dp + 4 * u);
          gw[4 * u + 0] = fmaf(v, g4.x, gw[4 * u + 0]); gw[4 * u + 1] = fmaf(v, g4.y, gw[4 * u + 1]);
          gw[4 * u + 2] = fmaf(v, g4.z, gw[4 * u + 2]); gw[4 * u + 3] = fmaf(v, g4.w, gw[4 * u + 3]);
        }
      }
    } else if (tap) {  // (first layer: no dense tile) the same sum over the pooled elements and their argmax positions
      for (int pp = group; pp < NPIX; pp += NG) {
        const int py = pp / PW, px = pp - py * PW;
#pragma unroll
        for (int u = 0; u < CV; ++u) {
          const float4 g4 = *reinterpret_cast<const float4 *>(sG + pp * COUT + 4 * u);
          const float g[4] = {g4.x, g4.y, g4.z, g4.w};
          const uint32_t pk = POOL ? *reinterpret_cast<const uint32_t *>(sA + pp * COUT + 4 * u) : 0u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const unsigned a = (pk >> (8 * j)) & 3u;
            const float v = POOL ? sIn[((2 * py + (a >> 1) + ky) * IW + (2 * px + (a & 1) + kx)) * CIN + ci]
                                 : sIn[((py + ky) * IW + (px + kx)) * CIN + ci];
            gw[4 * u + j] = fmaf(v, g[j], gw[4 * u + j]);
          }
        }
      }
    } else if (isb) {
      const int co = lt - 25 * CIN;
      for (int pp = group; pp < NPIX; pp += NG) gb += sG[pp * COUT + co];
    }
    if (NEED_DX) {  // d(in)[y][x][ci] = sum_{ky,kx,co} d(conv)[y+2-ky][x+2-kx][co] * W[ky][kx][ci][co]
      for (int p = tid; p < H * W; p += THREADS) {
        const int y = p / W, x = p - y * W;
        float acc[CIN];
#pragma unroll
        for (int c = 0; c < CIN; ++c) acc[c] = 0.0f;
#pragma unroll 1
        for (int t = 0; t < 25; ++t) {
          const int ty = t / 5, tx = t - ty * 5;
          const float *dp = sD + ((y + 4 - ty) * IW + (x + 4 - tx)) * COUT;
#pragma unroll
          for (int u = 0; u < CV; ++u) {
            const float4 d4 = *reinterpret_cast<const float4 *>(dp + 4 * u);
#pragma unroll
            for (int c = 0; c < CIN; ++c) {
              const float4 w4 = *reinterpret_cast<const float4 *>(sW + (t * CIN + c) * COUT + 4 * u);
              float sacc = acc[c];
              sacc = fmaf(d4.x, w4.x, sacc); sacc = fmaf(d4.y, w4.y, sacc); sacc = fmaf(d4.z, w4.z, sacc); sacc = fmaf(d4.w, w4.w, sacc);
              acc[c] = sacc;
            }
          }
        }
        float *o = din + (b * H * W + p) * CIN;
#pragma unroll
        for (int c = 0; c < CIN; ++c) o[c] = acc[c];
      }
    }
  }
  float *part = partials + (static_cast<int64_t>(blockIdx.x) * NG + group) * (25 * CIN * COUT + COUT);
  if (tap) {
#pragma unroll
    for (int co = 0; co < COUT; ++co) part[lt * COUT + co] = gw[co];  // HWIO order: ((ky*5+kx)*CIN+ci)*COUT+co
  } else if (isb) {
    part[25 * CIN * COUT + (lt - 25 * CIN)] = gb;
  }
}

static int conv_bwd_ctas(int64_t B) { return static_cast<int>(std::min<int64_t>(B, static_cast<int64_t>(sm_count()) * 4)); }

template <int CIN, int COUT, int H, int W, bool POOL, int THREADS>
static int launch_conv_fwd(const float *in, const float *w, const float *bias, float *out, uint8_t *arg, int64_t B,
                           cudaStream_t s) {
  auto kern = conv5x5_fwd_k<CIN, COUT, H, W, POOL, THREADS>;
  constexpr size_t smem = conv_fwd_smem<CIN, COUT, H, W, POOL>();
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    AIR_REQUIRE(e == cudaSuccess, AIR_ERR_CUDA, "cudaFuncSetAttribute(conv5x5_fwd): %s", cudaGetErrorString(e));
  }
  AIR_LAUNCH(kern, static_cast<unsigned>(B), THREADS, smem, s, in, w, bias, out, arg, B);
  count_launch();
  return check_launch("conv5x5_fwd");
}

template <int CIN, int COUT, int H, int W, bool POOL, bool NEED_DX, int THREADS, int GS>
static int launch_conv_bwd(const float *in, const float *w, const float *out, const uint8_t *arg, const float *dout,
                           float *din, float *dw, float *db, int accumulate, float *workspace, int64_t B, cudaStream_t s) {
  auto kern = conv5x5_bwd_k<CIN, COUT, H, W, POOL, NEED_DX, THREADS, GS>;
  constexpr size_t smem = conv_bwd_smem<CIN, COUT, H, W, POOL, NEED_DX>();
  static_assert(smem <= 220 * 1024, "the layer's tiles must fit the shared memory of one SM");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    AIR_REQUIRE(e == cudaSuccess, AIR_ERR_CUDA, "cudaFuncSetAttribute(conv5x5_bwd): %s", cudaGetErrorString(e));
  }
  // one wave of CTAs: as many as are co-resident (shared memory limits the 25x25x8 layer to 3 per SM), each looping
  // over its share of the images -- more CTAs than that would run as a second, partial wave
  static int per_sm = 0;
  if (per_sm == 0) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    AIR_REQUIRE(e == cudaSuccess && per_sm > 0, AIR_ERR_CUDA, "conv5x5_bwd occupancy query: %s", cudaGetErrorString(e));
    per_sm = std::min(per_sm, 4);  // the workspace is sized for 4 CTAs per SM
  }
  const int ctas = static_cast<int>(std::min<int64_t>(B, static_cast<int64_t>(sm_count()) * per_sm)), R = ctas * (THREADS / GS);
  AIR_LAUNCH(kern, ctas, THREADS, smem, s, in, w, out, arg, dout, din, workspace, B);
  count_launch();
  int rc = check_launch("conv5x5_bwd");
  if (rc) return rc;
  const int nw = 25 * CIN * COUT;
  rc = reduce_rows_launch(workspace, R, nw + COUT, nw, dw, accumulate, s);
  if (rc) return rc;
  rc = reduce_rows_launch(workspace + nw, R, nw + COUT, COUT, db, accumulate, s);
  if (rc) return rc;
  return check_launch("conv5x5_bwd reduce");
}

// the three layers of air_model.py:510-535 for F = cnn_filters output channels (conv2 / conv3 have F input channels)
template <int F, int T2, int GS2>
static int conv_fwd_dispatch(const float *in, const float *w, const float *bias, float *out, uint8_t *argmax, int64_t B, int H,
                             int W, int cin, int pool, cudaStream_t s) {
  if (cin == 1 && H == 50 && W == 50 && pool) return launch_conv_fwd<1, F, 50, 50, true, 256>(in, w, bias, out, argmax, B, s);
  if (cin == F && H == 25 && W == 25 && pool) return launch_conv_fwd<F, F, 25, 25, true, 160>(in, w, bias, out, argmax, B, s);
  if (cin == F && H == 12 && W == 12 && !pool) return launch_conv_fwd<F, F, 12, 12, false, 160>(in, w, bias, out, argmax, B, s);
  return AIR_ERR_UNSUPPORTED;
}

template <int F, int T2, int GS2>
static int conv_bwd_dispatch(const float *in, const float *w, const float *out, const uint8_t *argmax, const float *dout,
                             float *din, float *dw, float *db, int accumulate, float *workspace, int64_t B, int H, int W, int cin,
                             int pool, cudaStream_t s) {
  if (cin == 1 && H == 50 && W == 50 && pool && !din)
    return launch_conv_bwd<1, F, 50, 50, true, false, 256, 64>(in, w, out, argmax, dout, din, dw, db, accumulate, workspace, B, s);
  if (cin == F && H == 25 && W == 25 && pool && din)
    return launch_conv_bwd<F, F, 25, 25, true, true, T2, GS2>(in, w, out, argmax, dout, din, dw, db, accumulate, workspace, B, s);
  if (cin == F && H == 12 && W == 12 && !pool && din)
    return launch_conv_bwd<F, F, 12, 12, false, true, T2, GS2>(in, w, out, argmax, dout, din, dw, db, accumulate, workspace, B, s);
  return AIR_ERR_UNSUPPORTED;
}

}  // namespace air

using namespace air;

extern "C" int air_conv5x5_fwd(const float *in, const float *w, const float *bias, float *out, uint8_t *argmax, int64_t B,
                               int H, int W, int cin, int cout, int pool, air_stream_t stream) {
  AIR_REQUIRE(B >= 0 && H > 0 && W > 0, AIR_ERR_BAD_SHAPE, "conv5x5_fwd: bad shape");
  if (B == 0) return AIR_OK;
  AIR_REQUIRE(in && w && bias && out && (!pool || argmax), AIR_ERR_NULL, "conv5x5_fwd: null pointer");
  AIR_REQUIRE(aligned16(out) && (!pool || (reinterpret_cast<uintptr_t>(argmax) & 7u) == 0), AIR_ERR_UNSUPPORTED,
              "conv5x5_fwd: out must be 16-byte and argmax 8-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = AIR_ERR_UNSUPPORTED;
  if (cout == 8) rc = conv_fwd_dispatch<8, 256, 256>(in, w, bias, out, argmax, B, H, W, cin, pool, s);
  else if (cout == 4) rc = conv_fwd_dispatch<4, 256, 128>(in, w, bias, out, argmax, B, H, W, cin, pool, s);
  else if (cout == 16) rc = conv_fwd_dispatch<16, 512, 512>(in, w, bias, out, argmax, B, H, W, cin, pool, s);
  if (rc == AIR_ERR_UNSUPPORTED)
    set_error("conv5x5_fwd: only the three layers of air_model.py:510-535 are built, for 4, 8 or 16 filters "
              "(cin=%d cout=%d %dx%d pool=%d)", cin, cout, H, W, pool);
  return rc;
}

extern "C" int64_t air_conv5x5_bwd_workspace(int64_t B, int cin, int cout) {
  return static_cast<int64_t>(conv_bwd_ctas(B > 0 ? B : 1)) * 4 * (25 * static_cast<int64_t>(cin) * cout + cout);  // <= 4 groups / CTA
}

extern "C" int air_conv5x5_bwd(const float *in, const float *w, const float *out, const uint8_t *argmax, const float *dout,
                               float *din, float *dw, float *db, int accumulate, float *workspace, int64_t B, int H, int W,
                               int cin, int cout, int pool, air_stream_t stream) {
  AIR_REQUIRE(B >= 0 && H > 0 && W > 0, AIR_ERR_BAD_SHAPE, "conv5x5_bwd: bad shape");
  if (B == 0) return AIR_OK;
  AIR_REQUIRE(in && w && out && dout && dw && db && workspace && (!pool || argmax), AIR_ERR_NULL, "conv5x5_bwd: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = AIR_ERR_UNSUPPORTED;
  if (cout == 8) rc = conv_bwd_dispatch<8, 256, 256>(in, w, out, argmax, dout, din, dw, db, accumulate, workspace, B, H, W, cin, pool, s);
  else if (cout == 4) rc = conv_bwd_dispatch<4, 256, 128>(in, w, out, argmax, dout, din, dw, db, accumulate, workspace, B, H, W, cin, pool, s);
  else if (cout == 16) rc = conv_bwd_dispatch<16, 512, 512>(in, w, out, argmax, dout, din, dw, db, accumulate, workspace, B, H, W, cin, pool, s);
  if (rc == AIR_ERR_UNSUPPORTED)
    set_error("conv5x5_bwd: only the three layers of air_model.py:510-535 are built, for 4, 8 or 16 filters "
              "(cin=%d cout=%d %dx%d pool=%d din=%d)", cin, cout, H, W, pool, din != nullptr);
  return rc;
}
