// Shared pieces of the tcgen05 GEMM kernels (gemm_tc.cu: one CTA per tile, gemm_tc2.cu: CTA pairs, cta_group::2):
// PTX wrappers for TMA / mbarrier / tcgen05, UMMA descriptors, launch parameters, tensor-map construction.
#pragma once
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <stdlib.h>

#include "air_common.cuh"
#include "epilogue.cuh"

namespace air {

constexpr int kBM = 128;      // UMMA M (cta_group::1)
constexpr int kBK = 32;       // 32 tf32 = 128 bytes = one swizzle row
constexpr int kMaxStages = 8;  // ring depth is chosen per launch: deep when one CTA owns the SM, shallow when two share it
constexpr int kTcThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// same load, delivered to the same shared-memory offsets (tile and mbarrier) of every CTA of the cluster in cta_mask
__device__ __forceinline__ void tma_load_2d_mc(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// the same arrival, on the mbarrier at this offset in every CTA of cta_mask (frees a stage that both CTAs fill)
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// lo part of the 3xTF32 split: a - trunc_tf32(a) is exact in fp32 (<= 13 significant bits); rounding it to
// TF32 here (RN) keeps the tensor core's own truncation of the operand from biasing it
__device__ __forceinline__ float tf32_lo(float a) {
  const float hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  float lo = __fsub_rn(a, hi);
  if ((__float_as_uint(a) & 0x7F800000u) == 0x7F800000u) lo = 0.0f;  // inf / NaN travel in the hi part only
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(lo));
  return __uint_as_float(r);
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout).
// layout_type 2 = SWIZZLE_128B (K-major operands), 1 = SWIZZLE_128B_BASE32B: the only layout the
// tensor core accepts for MN-major 32-bit (TF32) operands -- 32-byte swizzle granules, 4-row atoms.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);       // start address      [0,14)
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;  // leading byte offset [16,30)
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;  // stride byte offset  [32,46)
  d |= static_cast<uint64_t>(1) << 46;                           // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

// cute::UMMA::InstrDescriptor for kind::tf32, FP32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4)                                   // c_format = F32
         | (2u << 7) | (2u << 10)                    // a_format = b_format = TF32
         | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16)
         | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

struct TcParams {
  float *C;             // output (or split-K workspace [splits][M][N] when splits > 1)
  const float *Cinit;
  const float *bias;
  const float *aux;
  int M, N, K, ldc, epi;
  float epi_param;
  int kb_per_split, num_kb, splits;
  int stages;  // TMA->MMA ring depth (<= kMaxStages)
  int chains;  // X3: number of hi*hi accumulators used round-robin over the k-blocks (1..kMaxChains)
  int flags;   // experiment switches (AIR_TC_FLAGS), 0 in production
  float *Cout;             // the real output (C above is the split-K workspace when splits > 1)
};

constexpr int kMaxChains = 3;
constexpr int kSplitWarps = 8;  // the epilogue warps write the lo tiles during the main loop (X3)

__host__ __device__ constexpr uint32_t tmem_cols_for(int BN, bool x3) {
  const int want = x3 ? (kMaxChains + 1) * BN : BN;
  return want <= 32 ? 32u : want <= 64 ? 64u : want <= 128 ? 128u : want <= 256 ? 256u : 512u;
}

// Epilogue of one 128-row output tile, run by the eight epilogue warps (warps 2..9): TMEM lane quarter = warp % 4, two
// warps per quarter each taking half of the tile's BN columns.  X3: the accumulator is the FP32 (RN) sum of `chains`
// hi*hi accumulators (columns c * BN) and the correction accumulator (column corr_col), in that fixed order.
// Split-K (p.splits > 1): the raw partial tile goes to the workspace slice of this split (p.C); splitk_reduce_kernel
// adds the slices in increasing split order and applies Cinit / bias / epilogue.
template <int BN, bool X3>
__device__ __forceinline__ void tc_epilogue(const TcParams &p, uint32_t tmem_acc, int m0, int n0, int split, int warp, int lane,
                                            int chains, int corr_col) {
  const int q = warp & 3;            // TMEM lane quarter this warp may access
  const int half = (warp - 2) >> 2;  // which half of the tile's columns
  const int row = m0 + q * 32 + lane;
  const bool partial = p.splits > 1;
  const int64_t slice = static_cast<int64_t>(p.M) * p.N;
  float *Cpart = p.C + static_cast<int64_t>(split) * slice;  // (partial only: p.C is the workspace)
  const bool vec_part = (p.N % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
  const bool vec_out = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.Cout) & 15) == 0) &&
                       (!p.Cinit || (reinterpret_cast<uintptr_t>(p.Cinit) & 15) == 0) &&
                       (!p.aux || p.epi == AIR_EPI_SIGMOID_RNG || (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0) &&
                       (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
  const uint32_t tbase = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  const int cbeg = half * (BN / 2), cend = cbeg + BN / 2;
  const bool rng = p.epi == AIR_EPI_SIGMOID_RNG;  // aux = device RNG state, the noise is generated here
  const unsigned long long *rng_st = reinterpret_cast<const unsigned long long *>(p.aux);

  // v[16] = product terms of (row, nb .. nb + 15)  ->  C = epi((Cinit + v) + bias)
  auto finalize = [&](float (&v)[16], int nb, bool fast, const float (&ci)[16], float (&ax)[16], const float (&bi)[16]) {
    if (fast) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (v[j] + ci[j]) + bi[j];
      if (rng) {  // 16 consecutive elements of row `row`: four Philox blocks when the row starts on a block boundary
        const unsigned long long idx = static_cast<unsigned long long>(row) * p.N + nb, seed = rng_st[0], ctr = rng_st[1];
        if ((idx & 3) == 0) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) *reinterpret_cast<float4 *>(ax + 4 * j4) = rng_normal4(seed, ctr, kRngLike, (idx >> 2) + j4);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) ax[j] = rng_normal_elem(seed, ctr, kRngLike, idx + j);
        }
      }
      apply_epilogue16(v, ax, p.epi, p.epi_param);
      float4 *dst = reinterpret_cast<float4 *>(p.Cout + static_cast<int64_t>(row) * p.ldc + nb);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) dst[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = nb + j;
        if (n < p.N) {
          const int64_t oo = static_cast<int64_t>(row) * p.ldc + n;
          float x = v[j];
          if (p.Cinit) x += p.Cinit[oo];
          if (p.bias) x += __ldg(p.bias + n);
          p.Cout[oo] = apply_epilogue(x, p.epi, epi_aux_value(p.aux, p.epi, oo, static_cast<int64_t>(row) * p.N + n), p.epi_param);
        }
      }
    }
  };
  auto fetch_inputs = [&](int nb, float (&ci)[16], float (&ax)[16], float (&bi)[16]) {
    const int64_t o = static_cast<int64_t>(row) * p.ldc + nb;
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      if (p.Cinit) *reinterpret_cast<float4 *>(ci + 4 * j4) = *reinterpret_cast<const float4 *>(p.Cinit + o + 4 * j4);
      if (p.aux && !rng) *reinterpret_cast<float4 *>(ax + 4 * j4) = *reinterpret_cast<const float4 *>(p.aux + o + 4 * j4);
      if (p.bias) *reinterpret_cast<float4 *>(bi + 4 * j4) = __ldg(reinterpret_cast<const float4 *>(p.bias + nb) + j4);
    }
  };

#pragma unroll 1
  for (int c0 = cbeg; c0 < cend; c0 += 16) {
    uint32_t r[16];
    tmem_ld_x16(tbase + c0, r);
    const int nb = n0 + c0;
    const bool row_ok = row < p.M && nb < p.N;
    const bool fast = (partial ? vec_part : vec_out) && row_ok && nb + 16 <= p.N;
    float ci[16], ax[16], bi[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) { ci[j] = 0.0f; ax[j] = 0.0f; bi[j] = 0.0f; }
    if (fast && !partial) fetch_inputs(nb, ci, ax, bi);  // epilogue inputs are fetched while the TMEM load is in flight
    tmem_ld_wait();
    if (X3) {  // product = ((chain_0 + chain_1) + ...) + correction, FP32 round-to-nearest, fixed order
      uint32_t r2[16];
      for (int c = 1; c < chains; ++c) {
        tmem_ld_x16(tbase + c * BN + c0, r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])));
      }
      tmem_ld_x16(tbase + corr_col + c0, r2);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])));
    }
    if (!row_ok) continue;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    if (!partial) {
      finalize(v, nb, fast, ci, ax, bi);
    } else if (fast) {
      float4 *dst = reinterpret_cast<float4 *>(Cpart + static_cast<int64_t>(row) * p.N + nb);
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) dst[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (nb + j < p.N) Cpart[static_cast<int64_t>(row) * p.N + nb + j] = v[j];
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor map: dim0 = contiguous dimension (extent d0), dim1 has stride ld elements (extent d1)
inline int make_tmap(CUtensorMap *map, const float *base, int64_t d0, int64_t d1, int64_t ld, int box0, int box1,
                     bool mn_major) {
  EncodeTiledFn enc = get_encode();
  AIR_REQUIRE(enc != nullptr, AIR_ERR_CUDA, "air_gemm(TF32): cuTensorMapEncodeTiled is unavailable");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(d0), static_cast<cuuint64_t>(d1)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AIR_REQUIRE(r == CUDA_SUCCESS, AIR_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (d0=%lld d1=%lld ld=%lld)", (int)r,
              (long long)d0, (long long)d1, (long long)ld);
  return AIR_OK;
}

// Tuning / diagnostics knobs, read ONCE when the library is first used (never on the launch path):
//   AIR_TC_STAGES   ring depth override            AIR_TC_BN       force 64- or 128-wide tiles
//   AIR_TC_CLUSTER  0 = never, 2 = always multicast pairs (one-CTA kernel)
//   AIR_TC_CHAINS   hi*hi accumulator chains of the X3 mode (1..3)
//   AIR_TC_PAIR     0 = never use the cta_group::2 kernel, 128 / 256 = force that pair-tile width, 1 = automatic
//   AIR_TC_PERSIST  0 = never use the persistent one-CTA kernel, 1 / 2 = use it everywhere with that many CTAs per SM
struct TcEnv {
  int stages = 0, bn = 0, cluster = 1, chains = 3, pair = 1, flags = 0, persist = -1, split_waves = 2, cluster4 = -1;
  TcEnv() {
    if (const char *e = getenv("AIR_TC_STAGES")) stages = std::max(1, atoi(e));
    if (const char *e = getenv("AIR_TC_BN")) bn = atoi(e);
    if (const char *e = getenv("AIR_TC_CLUSTER")) cluster = atoi(e);
    if (const char *e = getenv("AIR_TC_PAIR")) pair = atoi(e);
    if (const char *e = getenv("AIR_TC_FLAGS")) flags = atoi(e);
    if (const char *e = getenv("AIR_TC_PERSIST")) persist = atoi(e);
    if (const char *e = getenv("AIR_TC_CLUSTER4")) cluster4 = atoi(e);
    if (const char *e = getenv("AIR_TC_SPLIT_WAVES")) split_waves = std::max(1, atoi(e));
    if (const char *e = getenv("AIR_TC_CHAINS")) chains = std::max(1, std::min(atoi(e), kMaxChains));
  }
};
inline const TcEnv &tc_env() {
  static const TcEnv env;
  return env;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, dim3 cluster,
                                      Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster.x;
  attr[0].val.clusterDim.y = cluster.y;
  attr[0].val.clusterDim.z = cluster.z;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}


}  // namespace air
