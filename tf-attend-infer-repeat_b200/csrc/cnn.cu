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

constexpr int kCout = 8;

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

template <int CIN, int H, int W, bool POOL, int THREADS>
__global__ void __launch_bounds__(THREADS)
    conv5x5_fwd_k(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ bias,
                  float *__restrict__ out, uint8_t *__restrict__ arg, int64_t B) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  constexpr int IW = W + 4, IH = H + 4, PH = POOL ? H / 2 : H, PW = POOL ? W / 2 : W, NP = POOL ? 4 : 1;
  extern __shared__ __align__(16) float smem[];
  float *sIn = smem;                                   // [IH][IW][CIN] zero-padded input tile
  float *sW = sIn + ((IH * IW * CIN + 3) & ~3);        // [25*CIN][8]  (HWIO, as stored)
  float *sB = sW + 25 * CIN * kCout;                   // [8]
  const int tid = threadIdx.x;
  const int64_t b = blockIdx.x;
  load_padded<CIN, H, W>(in + b * H * W * CIN, sIn, tid, THREADS);
  for (int e = tid; e < 25 * CIN * kCout; e += THREADS) sW[e] = __ldg(w + e);
  if (tid < kCout) sB[tid] = __ldg(bias + tid);
  __syncthreads();
  for (int item = tid; item < PH * PW; item += THREADS) {
    const int py = item / PW, px = item - py * PW;
    const int y0 = POOL ? 2 * py : py, x0 = POOL ? 2 * px : px;
    float acc[NP][kCout];
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
      for (int co = 0; co < kCout; ++co) acc[p][co] = sB[co];
#pragma unroll 1
    for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 5; ++kx) {
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float4 wa = *reinterpret_cast<const float4 *>(sW + ((ky * 5 + kx) * CIN + ci) * kCout);
          const float4 wb = *reinterpret_cast<const float4 *>(sW + ((ky * 5 + kx) * CIN + ci) * kCout + 4);
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const float v = sIn[((y0 + (p >> 1) + ky) * IW + (x0 + (p & 1) + kx)) * CIN + ci];
            acc[p][0] = fmaf(v, wa.x, acc[p][0]); acc[p][1] = fmaf(v, wa.y, acc[p][1]);
            acc[p][2] = fmaf(v, wa.z, acc[p][2]); acc[p][3] = fmaf(v, wa.w, acc[p][3]);
            acc[p][4] = fmaf(v, wb.x, acc[p][4]); acc[p][5] = fmaf(v, wb.y, acc[p][5]);
            acc[p][6] = fmaf(v, wb.z, acc[p][6]); acc[p][7] = fmaf(v, wb.w, acc[p][7]);
          }
        }
      }
    }
    float r[kCout];
    uint8_t a[kCout];
#pragma unroll
    for (int co = 0; co < kCout; ++co) {
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
    float *o = out + (b * PH * PW + item) * kCout;
    *reinterpret_cast<float4 *>(o) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4 *>(o + 4) = make_float4(r[4], r[5], r[6], r[7]);
    if (POOL) {
      uint2 pk;
      pk.x = a[0] | (a[1] << 8) | (a[2] << 16) | (a[3] << 24);
      pk.y = a[4] | (a[5] << 8) | (a[6] << 16) | (a[7] << 24);
      *reinterpret_cast<uint2 *>(arg + (b * PH * PW + item) * kCout) = pk;
    }
  }
}

template <int CIN, int H, int W, bool POOL>
constexpr size_t conv_fwd_smem() {
  return (static_cast<size_t>(((H + 4) * (W + 4) * CIN + 3) & ~3) + 25 * CIN * kCout + kCout) * sizeof(float);
}

template <int CIN, int H, int W, bool POOL, bool NEED_DX>
constexpr size_t conv_bwd_smem() {
  constexpr size_t pin = ((H + 4) * (W + 4) * CIN + 3) & ~3;
  constexpr size_t pd = NEED_DX ? (H + 4) * (W + 4) * kCout : 0;
  constexpr size_t pg = (POOL ? (H / 2) * (W / 2) : H * W) * kCout;
  constexpr size_t pw = NEED_DX ? 25 * CIN * kCout : 0;
  return (pin + pd + pg + pw) * sizeof(float) + (POOL ? pg : 0);
}

// GS threads form a group that owns the 25*CIN weight taps + 8 biases; the THREADS / GS groups of a CTA split the
// output pixels among themselves and write one partial row each (fixed assignment -> deterministic).
template <int CIN, int H, int W, bool POOL, bool NEED_DX, int THREADS, int GS>
__global__ void __launch_bounds__(THREADS)
    conv5x5_bwd_k(const float *__restrict__ in, const float *__restrict__ w, const float *__restrict__ out,
                  const uint8_t *__restrict__ arg, const float *__restrict__ dout, float *__restrict__ din,
                  float *__restrict__ partials, int64_t B) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  static_assert(25 * CIN + kCout <= GS && THREADS % GS == 0, "one thread per weight tap + one per bias in every group");
  constexpr int NG = THREADS / GS;
  constexpr int IW = W + 4, IH = H + 4, PH = POOL ? H / 2 : H, PW = POOL ? W / 2 : W, NPIX = PH * PW;
  extern __shared__ __align__(16) float smem[];
  float *sIn = smem;                                              // [IH][IW][CIN] zero-padded layer input
  float *sD = sIn + ((IH * IW * CIN + 3) & ~3);                   // [IH][IW][8]   zero-padded dense d(conv)  (NEED_DX)
  float *sG = sD + (NEED_DX ? IH * IW * kCout : 0);               // [NPIX][8]     d(out) * (out > 0)
  float *sW = sG + NPIX * kCout;                                  // [25*CIN][8]   (NEED_DX)
  uint8_t *sA = reinterpret_cast<uint8_t *>(sW + (NEED_DX ? 25 * CIN * kCout : 0));  // [NPIX][8] argmax (POOL)
  const int tid = threadIdx.x;
  if (NEED_DX)
    for (int e = tid; e < 25 * CIN * kCout; e += THREADS) sW[e] = __ldg(w + e);
  // this thread's weight tap (ky, kx, ci) and its 8 gradient accumulators, or a bias accumulator
  const int group = tid / GS, lt = tid - group * GS;
  const bool tap = lt < 25 * CIN, isb = lt >= 25 * CIN && lt < 25 * CIN + kCout;
  const int ci = lt % CIN, kk = lt / CIN, ky = kk / 5, kx = kk - ky * 5;
  float gw[kCout] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float gb = 0.0f;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();  // the previous image's tiles are no longer read
    load_padded<CIN, H, W>(in + b * H * W * CIN, sIn, tid, THREADS);
    for (int e = tid; e < NPIX * kCout; e += THREADS) {
      const int64_t o = b * NPIX * kCout + e;
      sG[e] = out[o] > 0.0f ? dout[o] : 0.0f;  // ReluGrad
      if (POOL) sA[e] = arg[o];
    }
    if (NEED_DX)
      for (int e = tid; e < IH * IW * kCout; e += THREADS) sD[e] = 0.0f;
    __syncthreads();
    if (NEED_DX) {  // un-pool: scatter the pooled gradients to their argmax positions (distinct targets)
      for (int e = tid; e < NPIX * kCout; e += THREADS) {
        const int pp = e / kCout, co = e - pp * kCout;
        const int py = pp / PW, px = pp - py * PW;
        const int a = POOL ? sA[e] : 0;
        const int y = POOL ? 2 * py + (a >> 1) : py, x = POOL ? 2 * px + (a & 1) : px;
        sD[((y + 2) * IW + (x + 2)) * kCout + co] = sG[e];
      }
      __syncthreads();
    }
    if (tap && NEED_DX) {
      // dW[ky][kx][ci][:] += sum over conv pixels of in[y+ky-2][x+kx-2][ci] * d(conv)[y][x][:], from the dense
      // (un-pooled) tile that the dX pass needs anyway: one input load + two broadcast 128-bit loads per 8 FMAs
      constexpr int CH = POOL ? 2 * PH : H, CW = POOL ? 2 * PW : W;  // conv pixels that can carry a gradient
      for (int q = group; q < CH * CW; q += NG) {
        const int y = q / CW, x = q - y * CW;
        const float *dp = sD + ((y + 2) * IW + (x + 2)) * kCout;
        const float4 ga = *reinterpret_cast<const float4 *>(dp), gq = *reinterpret_cast<const float4 *>(dp + 4);
        const float v = sIn[((y + ky) * IW + (x + kx)) * CIN + ci];
        gw[0] = fmaf(v, ga.x, gw[0]); gw[1] = fmaf(v, ga.y, gw[1]); gw[2] = fmaf(v, ga.z, gw[2]); gw[3] = fmaf(v, ga.w, gw[3]);
        gw[4] = fmaf(v, gq.x, gw[4]); gw[5] = fmaf(v, gq.y, gw[5]); gw[6] = fmaf(v, gq.z, gw[6]); gw[7] = fmaf(v, gq.w, gw[7]);
      }
    } else if (tap) {  // (first layer: no dense tile) the same sum over the pooled elements and their argmax positions
      for (int pp = group; pp < NPIX; pp += NG) {
        const int py = pp / PW, px = pp - py * PW;
        const float4 ga = *reinterpret_cast<const float4 *>(sG + pp * kCout), gq = *reinterpret_cast<const float4 *>(sG + pp * kCout + 4);
        const float g[kCout] = {ga.x, ga.y, ga.z, ga.w, gq.x, gq.y, gq.z, gq.w};
        if (POOL) {
          const uint2 pk = *reinterpret_cast<const uint2 *>(sA + pp * kCout);
#pragma unroll
          for (int co = 0; co < kCout; ++co) {
            const unsigned a = ((co < 4 ? pk.x : pk.y) >> (8 * (co & 3))) & 3u;
            const float v = sIn[((2 * py + (a >> 1) + ky) * IW + (2 * px + (a & 1) + kx)) * CIN + ci];
            gw[co] = fmaf(v, g[co], gw[co]);
          }
        } else {
          const float v = sIn[((py + ky) * IW + (px + kx)) * CIN + ci];
#pragma unroll
          for (int co = 0; co < kCout; ++co) gw[co] = fmaf(v, g[co], gw[co]);
        }
      }
    } else if (isb) {
      const int co = lt - 25 * CIN;
      for (int pp = group; pp < NPIX; pp += NG) gb += sG[pp * kCout + co];
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
          const float *dp = sD + ((y + 4 - ty) * IW + (x + 4 - tx)) * kCout;
          const float4 da = *reinterpret_cast<const float4 *>(dp), db = *reinterpret_cast<const float4 *>(dp + 4);
#pragma unroll
          for (int c = 0; c < CIN; ++c) {
            const float4 wa = *reinterpret_cast<const float4 *>(sW + (t * CIN + c) * kCout);
            const float4 wb = *reinterpret_cast<const float4 *>(sW + (t * CIN + c) * kCout + 4);
            float s = acc[c];
            s = fmaf(da.x, wa.x, s); s = fmaf(da.y, wa.y, s); s = fmaf(da.z, wa.z, s); s = fmaf(da.w, wa.w, s);
            s = fmaf(db.x, wb.x, s); s = fmaf(db.y, wb.y, s); s = fmaf(db.z, wb.z, s); s = fmaf(db.w, wb.w, s);
            acc[c] = s;
          }
        }
        float *o = din + (b * H * W + p) * CIN;
#pragma unroll
        for (int c = 0; c < CIN; ++c) o[c] = acc[c];
      }
    }
  }
  float *part = partials + (static_cast<int64_t>(blockIdx.x) * NG + group) * (25 * CIN * kCout + kCout);
  if (tap) {
#pragma unroll
    for (int co = 0; co < kCout; ++co) part[lt * kCout + co] = gw[co];  // HWIO order: ((ky*5+kx)*CIN+ci)*8+co
  } else if (isb) {
    part[25 * CIN * kCout + (lt - 25 * CIN)] = gb;
  }
}

static int conv_bwd_ctas(int64_t B) { return static_cast<int>(std::min<int64_t>(B, static_cast<int64_t>(sm_count()) * 4)); }

template <int CIN, int H, int W, bool POOL, int THREADS>
static int launch_conv_fwd(const float *in, const float *w, const float *bias, float *out, uint8_t *arg, int64_t B,
                           cudaStream_t s) {
  auto kern = conv5x5_fwd_k<CIN, H, W, POOL, THREADS>;
  constexpr size_t smem = conv_fwd_smem<CIN, H, W, POOL>();
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    AIR_REQUIRE(e == cudaSuccess, AIR_ERR_CUDA, "cudaFuncSetAttribute(conv5x5_fwd): %s", cudaGetErrorString(e));
  }
  AIR_LAUNCH(kern, static_cast<unsigned>(B), THREADS, smem, s, in, w, bias, out, arg, B);
  count_launch();
  return check_launch("conv5x5_fwd");
}

template <int CIN, int H, int W, bool POOL, bool NEED_DX, int THREADS, int GS>
static int launch_conv_bwd(const float *in, const float *w, const float *out, const uint8_t *arg, const float *dout,
                           float *din, float *dw, float *db, int accumulate, float *workspace, int64_t B, cudaStream_t s) {
  auto kern = conv5x5_bwd_k<CIN, H, W, POOL, NEED_DX, THREADS, GS>;
  constexpr size_t smem = conv_bwd_smem<CIN, H, W, POOL, NEED_DX>();
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
  const int nw = 25 * CIN * kCout;
  rc = reduce_rows_launch(workspace, R, nw + kCout, nw, dw, accumulate, s);
  if (rc) return rc;
  rc = reduce_rows_launch(workspace + nw, R, nw + kCout, kCout, db, accumulate, s);
  if (rc) return rc;
  return check_launch("conv5x5_bwd reduce");
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
  if (cout == 8 && cin == 1 && H == 50 && W == 50 && pool) return launch_conv_fwd<1, 50, 50, true, 256>(in, w, bias, out, argmax, B, s);
  if (cout == 8 && cin == 8 && H == 25 && W == 25 && pool) return launch_conv_fwd<8, 25, 25, true, 160>(in, w, bias, out, argmax, B, s);
  if (cout == 8 && cin == 8 && H == 12 && W == 12 && !pool) return launch_conv_fwd<8, 12, 12, false, 160>(in, w, bias, out, argmax, B, s);
  set_error("conv5x5_fwd: only the three layers of air_model.py:510-535 are built (cin=%d cout=%d %dx%d pool=%d)", cin, cout, H, W, pool);
  return AIR_ERR_UNSUPPORTED;
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
  if (cout == 8 && cin == 1 && H == 50 && W == 50 && pool && !din)
    return launch_conv_bwd<1, 50, 50, true, false, 256, 64>(in, w, out, argmax, dout, din, dw, db, accumulate, workspace, B, s);
  if (cout == 8 && cin == 8 && H == 25 && W == 25 && pool && din)
    return launch_conv_bwd<8, 25, 25, true, true, 256, 256>(in, w, out, argmax, dout, din, dw, db, accumulate, workspace, B, s);
  if (cout == 8 && cin == 8 && H == 12 && W == 12 && !pool && din)
    return launch_conv_bwd<8, 12, 12, false, true, 256, 256>(in, w, out, argmax, dout, din, dw, db, accumulate, workspace, B, s);
  set_error("conv5x5_bwd: only the three layers of air_model.py:510-535 are built (cin=%d cout=%d %dx%d pool=%d din=%d)", cin,
            cout, H, W, pool, din != nullptr);
  return AIR_ERR_UNSUPPORTED;
}
