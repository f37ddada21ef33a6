// FP32 "exact" GEMM: SIMT FFMA, k strictly sequential per output element, so the result is
// bit-identical to oracle_gemm_seq_fma() whatever the tiling.  This is the parity mode of
// air_gemm(); the throughput mode (tcgen05, TF32) lives in gemm_tc.cu.
//
// Replaces MatMul + BiasAdd (+ ReLU/softplus) of tf.contrib.layers.fully_connected
// (air/vae.py:13-34, air/air_model.py:292-376), the BasicLSTMCell linear map
// (air/air_model.py:286) and the MatMul gradient ops of TF autodiff.
#include <algorithm>

#include "air_common.cuh"
#include "epilogue.cuh"

namespace air {

// 256 threads, BMxBN tile, each thread TMxTN outputs split into two halves per dimension
// (rows ty*TM/2.. and BM/2 + ty*TM/2..; same for columns) so that the float4 shared-memory
// reads of a quarter-warp are contiguous (no bank conflicts).
template <int BM, int BN, int BK, int TM, int TN, bool TA, bool TB>
__global__ void __launch_bounds__(256)
    gemm_simt(const float *__restrict__ A, const float *__restrict__ Bm, float *C, const float *Cinit,
              const float *__restrict__ bias, const float *aux, int M, int N, int K, int lda, int ldb, int ldc,
              int epi, float epi_param) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  static_assert(TM % 2 == 0 && TN % 2 == 0, "split tiles");
  constexpr int HM = TM / 2, HN = TN / 2;
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i < HM ? ty * HM + i : BM / 2 + ty * HM + (i - HM));
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j < HN ? tx * HN + j : BN / 2 + tx * HN + (j - HN));
      acc[i][j] = (Cinit != nullptr && m < M && n < N) ? Cinit[static_cast<int64_t>(m) * ldc + n] : 0.0f;
    }
  }

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- A tile -> As[k][m]
    if (TA) {  // stored [K, M], m contiguous
      for (int e = tid; e < BK * BM; e += 256) {
        const int k = e / BM, m = e % BM;
        const int gk = k0 + k, gm = m0 + m;
        As[k][m] = (gk < K && gm < M) ? __ldg(A + static_cast<int64_t>(gk) * lda + gm) : 0.0f;
      }
    } else {  // stored [M, K], k contiguous: consecutive threads take consecutive m, 4 k's each
      for (int e = tid; e < (BK / 4) * BM; e += 256) {
        const int m = e % BM, kq = (e / BM) * 4;
        const int gm = m0 + m;
        const float *src = A + static_cast<int64_t>(gm) * lda + k0 + kq;
#pragma unroll
        for (int j = 0; j < 4; ++j) As[kq + j][m] = (gm < M && k0 + kq + j < K) ? __ldg(src + j) : 0.0f;
      }
    }
    // ---- B tile -> Bs[k][n]
    if (!TB) {  // stored [K, N], n contiguous
      for (int e = tid; e < BK * BN; e += 256) {
        const int k = e / BN, n = e % BN;
        const int gk = k0 + k, gn = n0 + n;
        Bs[k][n] = (gk < K && gn < N) ? __ldg(Bm + static_cast<int64_t>(gk) * ldb + gn) : 0.0f;
      }
    } else {  // stored [N, K], k contiguous
      for (int e = tid; e < (BK / 4) * BN; e += 256) {
        const int n = e % BN, kq = (e / BN) * 4;
        const int gn = n0 + n;
        const float *src = Bm + static_cast<int64_t>(gn) * ldb + k0 + kq;
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[kq + j][n] = (gn < N && k0 + kq + j < K) ? __ldg(src + j) : 0.0f;
      }
    }
    __syncthreads();
    // zero padding beyond K contributes fma(0, 0, acc) == acc exactly
    const int kmax = min(BK, K - k0);
    for (int k = 0; k < kmax; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < HM; ++i) {
        a[i] = As[k][ty * HM + i];
        a[HM + i] = As[k][BM / 2 + ty * HM + i];
      }
#pragma unroll
      for (int j = 0; j < HN; ++j) {
        b[j] = Bs[k][tx * HN + j];
        b[HN + j] = Bs[k][BN / 2 + tx * HN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i < HM ? ty * HM + i : BM / 2 + ty * HM + (i - HM));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j < HN ? tx * HN + j : BN / 2 + tx * HN + (j - HN));
      if (n >= N) continue;
      const int64_t o = static_cast<int64_t>(m) * ldc + n;
      float v = acc[i][j];
      if (bias) v = add_rn(v, __ldg(bias + n));
      C[o] = apply_epilogue(v, epi, epi_aux_value(aux, epi, o, static_cast<int64_t>(m) * N + n), epi_param);
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN>
static int launch_simt(const float *A, const float *B, float *C, const float *Cinit, const float *bias,
                       const float *aux, int M, int N, int K, int lda, int ldb, int ldc, int tA, int tB, int epi,
                       float epi_param, cudaStream_t s) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  if (!tA && !tB)
    AIR_LAUNCH((gemm_simt<BM, BN, BK, TM, TN, false, false>), grid, 256, 0, s, A, B, C, Cinit, bias, aux, M, N, K, lda, ldb, ldc, epi, epi_param);
  else if (!tA && tB)
    AIR_LAUNCH((gemm_simt<BM, BN, BK, TM, TN, false, true>), grid, 256, 0, s, A, B, C, Cinit, bias, aux, M, N, K, lda, ldb, ldc, epi, epi_param);
  else if (tA && !tB)
    AIR_LAUNCH((gemm_simt<BM, BN, BK, TM, TN, true, false>), grid, 256, 0, s, A, B, C, Cinit, bias, aux, M, N, K, lda, ldb, ldc, epi, epi_param);
  else
    AIR_LAUNCH((gemm_simt<BM, BN, BK, TM, TN, true, true>), grid, 256, 0, s, A, B, C, Cinit, bias, aux, M, N, K, lda, ldb, ldc, epi, epi_param);
  count_launch();
  return check_launch("gemm_simt");
}

// K == 0: no products; C = epi((Cinit + bias)) elementwise (the first LSTM step, whose initial state is zero:
// air_model.py:540-542).  Same rounding as the GEMM kernels' epilogue: acc = Cinit, then + bias, then the activation.
template <int VEC>
__global__ void __launch_bounds__(256)
    gemm_k0_kernel(float *C, const float *Cinit, const float *__restrict__ bias, const float *aux, int M, int N, int ldc,
                   int epi, float epi_param) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  const int NV = N / VEC;
  const int64_t total = static_cast<int64_t>(M) * NV;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t m = e / NV;
    const int n = static_cast<int>(e - m * NV) * VEC;
    const int64_t o = m * ldc + n;
    if (VEC == 4) {
      float4 v = Cinit ? *reinterpret_cast<const float4 *>(Cinit + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) {
        const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + n));
        v.x = add_rn(v.x, bv.x); v.y = add_rn(v.y, bv.y); v.z = add_rn(v.z, bv.z); v.w = add_rn(v.w, bv.w);
      }
      float4 ax = make_float4(0.f, 0.f, 0.f, 0.f);
      if (epi == AIR_EPI_SIGMOID_RNG) {  // N % 4 == 0 here: element m * N + n starts a Philox block
        const unsigned long long *st = reinterpret_cast<const unsigned long long *>(aux);
        ax = rng_normal4(st[0], st[1], kRngLike, static_cast<unsigned long long>(m * N + n) >> 2);
      } else if (aux) {
        ax = *reinterpret_cast<const float4 *>(aux + o);
      }
      v.x = apply_epilogue(v.x, epi, ax.x, epi_param); v.y = apply_epilogue(v.y, epi, ax.y, epi_param);
      v.z = apply_epilogue(v.z, epi, ax.z, epi_param); v.w = apply_epilogue(v.w, epi, ax.w, epi_param);
      *reinterpret_cast<float4 *>(C + o) = v;
    } else {
      float v = Cinit ? Cinit[o] : 0.0f;
      if (bias) v = add_rn(v, __ldg(bias + n));
      C[o] = apply_epilogue(v, epi, epi_aux_value(aux, epi, o, static_cast<int64_t>(m) * N + n), epi_param);
    }
  }
}

int gemm_k0(float *C, const float *Cinit, const float *bias, const float *aux, int M, int N, int ldc, int epi,
            float epi_param, cudaStream_t s) {
  const bool vec = N % 4 == 0 && ldc % 4 == 0 && aligned16(C) && (!Cinit || aligned16(Cinit)) && (!bias || aligned16(bias)) &&
                   (!aux || epi == AIR_EPI_SIGMOID_RNG || aligned16(aux));
  const int64_t work = static_cast<int64_t>(M) * (vec ? N / 4 : N);
  const int grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, static_cast<int64_t>(sm_count()) * 8)));
  if (vec)
    AIR_LAUNCH(gemm_k0_kernel<4>, grid, 256, 0, s, C, Cinit, bias, aux, M, N, ldc, epi, epi_param);
  else
    AIR_LAUNCH(gemm_k0_kernel<1>, grid, 256, 0, s, C, Cinit, bias, aux, M, N, ldc, epi, epi_param);
  count_launch();
  return check_launch("gemm_k0");
}

int gemm_fp32_exact(const float *A, const float *B, float *C, const float *Cinit, const float *bias, const float *aux,
                    int M, int N, int K, int lda, int ldb, int ldc, int tA, int tB, int epi, float epi_param, cudaStream_t s) {
  // big tiles when they still fill the machine, small tiles otherwise
  const int64_t big_ctas = static_cast<int64_t>((M + 127) / 128) * ((N + 127) / 128);
  if (N > 64 && M > 64 && big_ctas >= sm_count())
    return launch_simt<128, 128, 16, 8, 8>(A, B, C, Cinit, bias, aux, M, N, K, lda, ldb, ldc, tA, tB, epi, epi_param, s);
  return launch_simt<64, 64, 16, 4, 4>(A, B, C, Cinit, bias, aux, M, N, K, lda, ldb, ldc, tA, tB, epi, epi_param, s);
}

int gemm_tf32(const float *A, const float *B, float *C, const float *Cinit, const float *bias, const float *aux, int M,
              int N, int K, int lda, int ldb, int ldc, int tA, int tB, int epi, float epi_param, bool x3, float *workspace,
              int64_t ws_floats, cudaStream_t s);

}  // namespace air

extern "C" int air_gemm_ws(const float *A, const float *B, float *C, const float *Cinit, const float *bias,
                           const float *aux, int64_t M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                           int epilogue, float epi_param, int mode, float *workspace, int64_t workspace_floats,
                           air_stream_t stream) {
  AIR_REQUIRE(M >= 0 && N >= 0 && K >= 0 && M < (int64_t(1) << 31), AIR_ERR_BAD_SHAPE, "air_gemm: bad shape M=%lld N=%d K=%d",
              (long long)M, N, K);
  if (M == 0 || N == 0) return AIR_OK;
  AIR_REQUIRE(((A && B) || K == 0) && C, AIR_ERR_NULL, "air_gemm: null pointer");  // K == 0: C = epi(Cinit + bias)
  AIR_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, AIR_ERR_BAD_SHAPE,
              "air_gemm: leading dimension too small (lda=%d ldb=%d ldc=%d)", lda, ldb, ldc);
  AIR_REQUIRE(epilogue >= AIR_EPI_NONE && epilogue <= AIR_EPI_SIGMOID_RNG, AIR_ERR_BAD_SHAPE, "air_gemm: bad epilogue %d",
              epilogue);
  AIR_REQUIRE((epilogue != AIR_EPI_MUL_DRELU && epilogue != AIR_EPI_MUL_DSOFTPLUS && epilogue != AIR_EPI_SIGMOID_NOISE &&
               epilogue != AIR_EPI_SIGMOID_RNG) || aux,
              AIR_ERR_NULL, "air_gemm: epilogue %d needs aux", epilogue);
  AIR_REQUIRE(epilogue != AIR_EPI_SIGMOID_RNG || (reinterpret_cast<uintptr_t>(aux) & 7) == 0, AIR_ERR_BAD_ALIGN,
              "air_gemm: AIR_EPI_SIGMOID_RNG takes the 8-byte aligned device RNG state as aux");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  AIR_REQUIRE(mode == AIR_GEMM_FP32_EXACT || mode == AIR_GEMM_TF32 || mode == AIR_GEMM_TF32X3, AIR_ERR_UNSUPPORTED,
              "air_gemm: unknown mode %d", mode);
  if (K == 0) return air::gemm_k0(C, Cinit, bias, aux, static_cast<int>(M), N, ldc, epilogue, epi_param, s);
  if (mode == AIR_GEMM_FP32_EXACT)
    return air::gemm_fp32_exact(A, B, C, Cinit, bias, aux, static_cast<int>(M), N, K, lda, ldb, ldc, transA, transB,
                                epilogue, epi_param, s);
  return air::gemm_tf32(A, B, C, Cinit, bias, aux, static_cast<int>(M), N, K, lda, ldb, ldc, transA, transB, epilogue,
                        epi_param, mode == AIR_GEMM_TF32X3, workspace, workspace ? workspace_floats : 0, s);
}

extern "C" int air_gemm_ex(const float *A, const float *B, float *C, const float *Cinit, const float *bias,
                           const float *aux, int64_t M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                           int epilogue, float epi_param, int mode, air_stream_t stream) {
  return air_gemm_ws(A, B, C, Cinit, bias, aux, M, N, K, lda, ldb, ldc, transA, transB, epilogue, epi_param, mode, nullptr, 0,
                     stream);
}

extern "C" int air_gemm(const float *A, const float *B, float *C, const float *Cinit, const float *bias,
                        const float *aux, int64_t M, int N, int K, int lda, int ldb, int ldc, int transA, int transB,
                        int epilogue, int mode, air_stream_t stream) {
  return air_gemm_ws(A, B, C, Cinit, bias, aux, M, N, K, lda, ldb, ldc, transA, transB, epilogue, 0.0f, mode, nullptr, 0, stream);
}
