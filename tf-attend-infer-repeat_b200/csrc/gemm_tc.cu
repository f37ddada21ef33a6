// TF32 tensor-core GEMM for sm_100a: tcgen05.mma (kind::tf32) issued by one thread, FP32
// accumulators in TMEM, operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a
// 3/4-stage mbarrier ring, epilogue warps read the accumulator back with tcgen05.ld and fuse
// Cinit / bias / activation (or activation-derivative) before the global store.
//
// This is the throughput mode of air_gemm() (mode AIR_GEMM_TF32); it replaces the MatMul +
// BiasAdd (+ activation) ops of tf.contrib.layers.fully_connected / BasicLSTMCell
// (air/vae.py:13-34, air/air_model.py:286-376) and their MatMul gradients.  Operands stay
// FP32 in HBM (the tensor core reads the top 19 bits), so no conversion pass is needed.
//
// All four operand layouts are native: a K-contiguous operand is a "K-major" UMMA operand
// (TMA box 32 x rows), an M/N-contiguous one is "MN-major" (TMA box 32 x BLOCK_K per 32-wide
// chunk).  Weight-gradient GEMMs (reduction over the batch, few output tiles) use split-K
// with a deterministic second pass.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue: TMEM lane quarter = warp_idx % 4, two warps per quarter each taking half
// of the tile's columns (the epilogue, not the MMA pipe, bounds the small-K GEMMs of this model).
//
// Mode AIR_GEMM_TF32X3 (template flag X3) is the FP32-grade tensor-core mode ("3xTF32"): every fp32
// operand a is used as a = hi + lo with hi = the top 19 bits of a (exactly what the tensor core reads
// from the raw fp32 tile, so the TMA-loaded tile IS the hi operand) and lo = rn_tf32(a - hi), written by
// the eight epilogue warps -- idle during the main loop otherwise -- into a second shared-memory tile of
// the same swizzled layout.  Three MMAs per k-step: hi*hi into the main accumulator(s), lo*hi + hi*lo
// into a separate correction accumulator (its terms are 2^-11 smaller, so its own rounding is
// irrelevant); the dropped lo*lo term is <= 2^-20 relative per product.  The hi*hi products are
// accumulated in `chains` TMEM accumulators used round-robin over the k-blocks and summed in FP32 (RN)
// by the epilogue: the tensor core's accumulate step rounds toward zero, so shorter chains mean less of
// that bias (measured in tests/diag_tc_rounding.py).
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
};

constexpr int kMaxChains = 3;
constexpr int kSplitWarps = 8;  // the epilogue warps write the lo tiles during the main loop (X3)

__host__ __device__ constexpr uint32_t tmem_cols_for(int BN, bool x3) {
  const int want = x3 ? (kMaxChains + 1) * BN : BN;
  return want <= 32 ? 32u : want <= 64 ? 64u : want <= 128 ? 128u : want <= 256 ? 256u : 512u;
}

// CL: launched as clusters of two CTAs that are neighbours along M (same N tile).  They need the same B tile:
// each loads half of it and multicasts it to both, which halves the L2 -> SM traffic of that operand (the kernel
// is L2 -> SM bandwidth bound with fp32 operands, profiles/r1_gemm_tf32_big_ncu_full.md).  A stage is refilled only
// after BOTH CTAs' MMAs have read it: the empty barriers count two arrivals, one multicast commit from each CTA.
template <int BN, bool A_MN, bool B_MN, bool CL, bool X3>
__global__ void __launch_bounds__(kTcThreads)
    gemm_tf32_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcParams p) {
  constexpr uint32_t kABytes = kBM * kBK * 4, kBBytes = BN * kBK * 4;
  constexpr uint32_t kRawBytes = kABytes + kBBytes;              // what TMA delivers per stage (= the hi operands)
  constexpr uint32_t kStageBytes = X3 ? 2 * kRawBytes : kRawBytes;  // X3: + the lo tiles, same layout
  constexpr uint32_t kTmemCols = tmem_cols_for(BN, X3);
  const int kStages = p.stages;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // 1024-byte aligned tiles (SWIZZLE_128B atoms)
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], split_bar[kMaxStages], tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int split = blockIdx.z;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;
  const int chains = X3 ? min(p.chains, nkb) : 1;  // every chain used below has received at least one k-block

  if (threadIdx.x == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL ? 2 : 1);
      mbar_init(&split_bar[s], kSplitWarps);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;
  const uint32_t crank = CL ? cluster_ctarank() : 0;
  if (CL) cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
  pdl_sync();  // PDL: barriers, TMEM and descriptors were set up while the previous grid drained

  // ================= TMA producers =================
  // Every stage is loaded as 4 + 4 quarter boxes.  Plain TF32: issued by four threads, lane 0 of warp 0 and of the
  // (otherwise idle until the accumulator is complete) epilogue warps 2..4 -- measured neutral against a single
  // issuing thread (TMA issue is not the limit).  X3: the epilogue warps are busy splitting, warp 0 issues all eight.
  if (warp == 0 || (!X3 && warp >= 2 && warp <= 4)) {
    if (lane == 0) {
      const int pi_begin = X3 ? 0 : (warp == 0 ? 0 : warp - 1), pi_end = X3 ? 4 : pi_begin + 1;
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait_spin(&empty_bar[s], ph ^ 1);
        unsigned char *sa = tiles + s * kStageBytes, *sb = sa + kABytes;
        if (warp == 0) mbar_expect_tx(&full_bar[s], kRawBytes);
        const int k0 = (kb_begin + i) * kBK;
        for (int pi = pi_begin; pi < pi_end; ++pi) {
          if (!A_MN) {  // box {32 k, 32 m}: rows [32 pi, 32 pi + 32)
            tma_load_2d(sa + pi * (32 * 128), &mapA, &full_bar[s], k0, m0 + pi * 32);
          } else {      // box {32 m, 32 k}: 32-wide chunk pi
            tma_load_2d(sa + pi * (kBK * 128), &mapA, &full_bar[s], m0 + pi * 32, k0);
          }
          unsigned char *bdst;
          int bc0, bc1;
          if (!B_MN) {  // box {32 k, BN/4 n}
            bdst = sb + pi * (BN / 4 * 128); bc0 = k0; bc1 = n0 + pi * (BN / 4);
          } else if (BN == 128) {  // box {32 n, 32 k}: chunk pi
            bdst = sb + pi * (kBK * 128); bc0 = n0 + pi * 32; bc1 = k0;
          } else {      // BN == 64: box {32 n, 16 k}: chunk pi/2, k-half pi%2
            bdst = sb + (pi >> 1) * (kBK * 128) + (pi & 1) * (16 * 128); bc0 = n0 + (pi >> 1) * 32; bc1 = k0 + (pi & 1) * 16;
          }
          if (!CL) {
            tma_load_2d(bdst, &mapB, &full_bar[s], bc0, bc1);
          } else if (static_cast<uint32_t>(pi >> 1) == crank) {  // this CTA's half of the shared B tile, to both CTAs
            tma_load_2d_mc(bdst, &mapB, &full_bar[s], bc0, bc1, static_cast<uint16_t>(3));
          }
        }
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  }
  if (X3 && warp >= 2) {
    // ================= 3xTF32 splitter: lo = rn_tf32(a - trunc_tf32(a)) for both raw tiles of every stage =========
    // The lo tile has the raw tile's (swizzled) layout, so this is elementwise on float4s.  A stage's lo tile is
    // rewritten only after full_bar[s] of the NEXT ring pass, i.e. after the MMAs that read it have completed.
    constexpr int kVec = kRawBytes / 16, kThreads = kSplitWarps * 32;
    static_assert(kVec % kThreads == 0, "split loop assumes a whole number of float4 per thread");
    const int tid = threadIdx.x - 64;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&full_bar[s], ph);
      const float4 *raw = reinterpret_cast<const float4 *>(tiles + s * kStageBytes);
      float4 *lo = reinterpret_cast<float4 *>(tiles + s * kStageBytes + kRawBytes);
      float4 v[kVec / kThreads];
#pragma unroll
      for (int j = 0; j < kVec / kThreads; ++j) v[j] = raw[j * kThreads + tid];
#pragma unroll
      for (int j = 0; j < kVec / kThreads; ++j)
        lo[j * kThreads + tid] = make_float4(tf32_lo(v[j].x), tf32_lo(v[j].y), tf32_lo(v[j].z), tf32_lo(v[j].w));
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&split_bar[s]);
      if (++s == kStages) { s = 0; ph ^= 1; }
    }
  }
  if (warp == 0) {
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, A_MN, B_MN);
      // K-major  (SWIZZLE_128B):         8-row groups 1024 B apart (SBO), K advance 32 B per MMA (8 tf32).
      // MN-major (SWIZZLE_128B_BASE32B): 32-wide chunks kBK*128 B apart (LBO), 4-k-row atoms 512 B apart
      //                                  (SBO), K advance 8 rows = 1024 B per MMA.
      constexpr uint32_t a_lbo = A_MN ? kBK * 128 : 16, a_sbo = A_MN ? 512 : 1024, a_adv = A_MN ? 1024 : 32;
      constexpr uint32_t b_lbo = B_MN ? kBK * 128 : 16, b_sbo = B_MN ? 512 : 1024, b_adv = B_MN ? 1024 : 32;
      constexpr uint32_t a_lt = A_MN ? 1 : 2, b_lt = B_MN ? 1 : 2;
      const uint32_t tmem_corr = tmem_acc + kMaxChains * BN;  // X3: lo*hi + hi*lo
      int s = 0, chain = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait_spin(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(tiles + s * kStageBytes), sb = sa + kABytes;
        const uint32_t tmem_main = tmem_acc + (X3 ? chain * BN : 0);
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) {
          const uint64_t da = make_smem_desc(sa + k * a_adv, a_lbo, a_sbo, a_lt);
          const uint64_t db = make_smem_desc(sb + k * b_adv, b_lbo, b_sbo, b_lt);
          umma_tf32(tmem_main, da, db, idesc, (i >= chains || k != 0) ? 1u : 0u);  // chain c is first used by k-block c
        }
        if (X3) {
          if (++chain == chains) chain = 0;
          mbar_wait_spin(&split_bar[s], ph);  // the hi*hi MMAs above run while the lo tiles are being written
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            const uint64_t da = make_smem_desc(sa + k * a_adv, a_lbo, a_sbo, a_lt);
            const uint64_t db = make_smem_desc(sb + k * b_adv, b_lbo, b_sbo, b_lt);
            const uint64_t dal = make_smem_desc(sa + kRawBytes + k * a_adv, a_lbo, a_sbo, a_lt);
            const uint64_t dbl = make_smem_desc(sb + kRawBytes + k * b_adv, b_lbo, b_sbo, b_lt);
            umma_tf32(tmem_corr, dal, db, idesc, (i | k) != 0 ? 1u : 0u);
            umma_tf32(tmem_corr, da, dbl, idesc, 1u);
          }
        }
        if (CL) umma_commit_mc(&empty_bar[s], static_cast<uint16_t>(3));
        else umma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
      umma_commit(&tmem_full_bar);  // accumulator complete
    }
  }
  if (warp >= 2) {
    // ================= epilogue: TMEM -> registers -> global (8 warps) =================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;  // which half of the tile's columns
    const int row = m0 + q * 32 + lane;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    float *Cout = p.C;
    int ldo = p.ldc;
    const bool partial = p.splits > 1;
    if (partial) {
      Cout = p.C + static_cast<int64_t>(split) * p.M * p.N;
      ldo = p.N;
    }
    const bool vec_ok = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(Cout) & 15) == 0) &&
                        (partial || ((p.ldc % 4 == 0) && (!p.Cinit || (reinterpret_cast<uintptr_t>(p.Cinit) & 15) == 0) &&
                                     (!p.aux || (reinterpret_cast<uintptr_t>(p.aux) & 15) == 0) &&
                                     (!p.bias || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)));
    const uint32_t tbase = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
    const int cbeg = half * (BN / 2), cend = cbeg + BN / 2;
#pragma unroll 1
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
      uint32_t r[16];
      tmem_ld_x16(tbase + c0, r);
      const int nb = n0 + c0;
      const bool row_ok = row < p.M && nb < p.N;
      const bool fast = vec_ok && row_ok && nb + 16 <= p.N;
      float ci[16], ax[16], bi[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { ci[j] = 0.0f; ax[j] = 0.0f; bi[j] = 0.0f; }
      const int64_t o = static_cast<int64_t>(row) * p.ldc + nb;
      if (fast && !partial) {  // epilogue inputs are fetched while the TMEM load is in flight
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          if (p.Cinit) *reinterpret_cast<float4 *>(ci + 4 * j4) = *reinterpret_cast<const float4 *>(p.Cinit + o + 4 * j4);
          if (p.aux) *reinterpret_cast<float4 *>(ax + 4 * j4) = *reinterpret_cast<const float4 *>(p.aux + o + 4 * j4);
          if (p.bias) *reinterpret_cast<float4 *>(bi + 4 * j4) = __ldg(reinterpret_cast<const float4 *>(p.bias + nb) + j4);
        }
      }
      tmem_ld_wait();
      if (X3) {  // product = ((chain_0 + chain_1) + ...) + correction, FP32 round-to-nearest, fixed order
        uint32_t r2[16];
        for (int c = 1; c < chains; ++c) {
          tmem_ld_x16(tbase + c * BN + c0, r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])));
        }
        tmem_ld_x16(tbase + kMaxChains * BN + c0, r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])));
      }
      if (!row_ok) continue;
      if (fast) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        if (!partial) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (v[j] + ci[j]) + bi[j];
          apply_epilogue16(v, ax, p.epi, p.epi_param);
        }
        float4 *dst = reinterpret_cast<float4 *>(Cout + static_cast<int64_t>(row) * ldo + nb);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) dst[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = nb + j;
          if (n < p.N) {
            float v = __uint_as_float(r[j]);
            if (!partial) {
              const int64_t oo = static_cast<int64_t>(row) * p.ldc + n;
              if (p.Cinit) v += p.Cinit[oo];
              if (p.bias) v += __ldg(p.bias + n);
              v = apply_epilogue(v, p.epi, p.aux ? p.aux[oo] : 0.0f, p.epi_param);
            }
            Cout[static_cast<int64_t>(row) * ldo + n] = v;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, kTmemCols);
  }
  if (CL) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers
}

// second pass of split-K: C = epi((Cinit + sum_s partial[s]) + bias), s in increasing order.
// float4 per thread, splits loop unrolled 4x so the partial loads are in flight together.
__global__ void __launch_bounds__(256)
    splitk_reduce_kernel(const float *__restrict__ ws, int splits, float *C, const float *Cinit,
                         const float *__restrict__ bias, const float *aux, int M, int N, int ldc, int epi, float epi_param) {
  pdl_sync();  // PDL: no global access before the previous grid has completed
  const int64_t total = static_cast<int64_t>(M) * N;  // N % 4 == 0 on this path
  const int64_t nvec = total >> 2;
  for (int64_t v4 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; v4 < nvec;
       v4 += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t e = v4 << 2;
    const float4 *src = reinterpret_cast<const float4 *>(ws + e);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 4 <= splits; s += 4) {
      const float4 a = __ldg(src + (static_cast<int64_t>(s) * total >> 2));
      const float4 b = __ldg(src + (static_cast<int64_t>(s + 1) * total >> 2));
      const float4 c = __ldg(src + (static_cast<int64_t>(s + 2) * total >> 2));
      const float4 d = __ldg(src + (static_cast<int64_t>(s + 3) * total >> 2));
      acc.x = (((acc.x + a.x) + b.x) + c.x) + d.x; acc.y = (((acc.y + a.y) + b.y) + c.y) + d.y;
      acc.z = (((acc.z + a.z) + b.z) + c.z) + d.z; acc.w = (((acc.w + a.w) + b.w) + c.w) + d.w;
    }
    for (; s < splits; ++s) {
      const float4 a = __ldg(src + (static_cast<int64_t>(s) * total >> 2));
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    const int m = static_cast<int>(e / N), n = static_cast<int>(e - static_cast<int64_t>(m) * N);
    const float vals[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t o = static_cast<int64_t>(m) * ldc + n + j;
      float v = vals[j];
      if (Cinit) v += Cinit[o];
      if (bias) v += __ldg(bias + n + j);
      C[o] = apply_epilogue(v, epi, aux ? aux[o] : 0.0f, epi_param);
    }
  }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
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
static int make_tmap(CUtensorMap *map, const float *base, int64_t d0, int64_t d1, int64_t ld, int box0, int box1,
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
//   AIR_TC_CLUSTER  0 = never, 2 = always pair CTAs AIR_TC_CHAINS   hi*hi accumulator chains of the X3 mode (1..3)
struct TcEnv {
  int stages = 0, bn = 0, cluster = 1, chains = 2;
  TcEnv() {
    if (const char *e = getenv("AIR_TC_STAGES")) stages = std::max(1, atoi(e));
    if (const char *e = getenv("AIR_TC_BN")) bn = atoi(e);
    if (const char *e = getenv("AIR_TC_CLUSTER")) cluster = atoi(e);
    if (const char *e = getenv("AIR_TC_CHAINS")) chains = std::max(1, std::min(atoi(e), kMaxChains));
  }
};
static const TcEnv &tc_env() {
  static const TcEnv env;
  return env;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_cluster_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, dim3 cluster,
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

template <int BN, bool A_MN, bool B_MN, bool CL, bool X3>
static int launch_tc(const CUtensorMap &ma, const CUtensorMap &mb, TcParams p, cudaStream_t s) {
  auto kern = gemm_tf32_kernel<BN, A_MN, B_MN, CL, X3>;
  const size_t stage_bytes = static_cast<size_t>(kBM + BN) * kBK * 4 * (X3 ? 2 : 1);
  const int max_stages = static_cast<int>(std::min<size_t>(kMaxStages, 200 * 1024 / stage_bytes));
  // Keep as many operand bytes in flight per SM as shared memory allows: a deep ring when the grid gives each SM
  // one CTA, half of it when two CTAs will share an SM.  (tests/diag_gemm_single_cta.py: one CTA per SM with
  // operands streaming from HBM runs 1.5 / 2.3 / 2.35 k-blocks per us with 2 / 4 / 6 stages.)  X3 stages are twice
  // as large (raw + lo tiles) and its CTAs own all 512 TMEM columns: always one CTA per SM, the deepest ring that fits.
  const int64_t ctas = static_cast<int64_t>((p.N + BN - 1) / BN) * ((p.M + kBM - 1) / kBM) * p.splits;
  const size_t budget = (X3 || ctas <= sm_count() ? 200 : 100) * 1024;
  int stages = static_cast<int>(budget / stage_bytes);
  stages = std::max(2, std::min({stages, max_stages, std::max(p.kb_per_split, 2)}));
  if (tc_env().stages) stages = std::min(tc_env().stages, max_stages);
  p.stages = stages;
  p.chains = tc_env().chains;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024;
  const size_t smem_max = static_cast<size_t>(max_stages) * stage_bytes + 1024;
  // the opt-in shared-memory limit is a per-device function attribute: set it once per (instantiation, device)
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = uint64_t(1) << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_max));
    AIR_REQUIRE(e == cudaSuccess, AIR_ERR_CUDA, "cudaFuncSetAttribute(gemm_tf32): %s", cudaGetErrorString(e));
    configured.fetch_or(bit, std::memory_order_release);
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + kBM - 1) / kBM, p.splits);
  if (CL) {
    cudaError_t e = launch_cluster_pdl(kern, grid, dim3(kTcThreads), smem, s, dim3(1, 2, 1), ma, mb, p);
    AIR_REQUIRE(e == cudaSuccess, AIR_ERR_CUDA, "gemm_tf32 cluster launch: %s", cudaGetErrorString(e));
  } else {
    AIR_LAUNCH(kern, grid, kTcThreads, smem, s, ma, mb, p);
  }
  count_launch();
  return check_launch(X3 ? "gemm_tf32x3" : "gemm_tf32");
}

int gemm_fp32_exact(const float *A, const float *B, float *C, const float *Cinit, const float *bias, const float *aux,
                    int M, int N, int K, int lda, int ldb, int ldc, int tA, int tB, int epi, float epi_param, cudaStream_t s);

// x3: the FP32-grade 3xTF32 mode.  workspace (nullable) / ws_floats: caller-owned split-K scratch; without one (or with
// one that is too small for a useful split) the GEMM runs unsplit -- nothing is ever allocated or cached here.
int gemm_tf32(const float *A, const float *B, float *C, const float *Cinit, const float *bias, const float *aux, int M,
              int N, int K, int lda, int ldb, int ldc, int tA, int tB, int epi, float epi_param, bool x3, float *workspace,
              int64_t ws_floats, cudaStream_t s) {
  if (K == 0)  // no products at all: C = epi(Cinit + bias); nothing for the tensor cores to do
    return gemm_fp32_exact(A, B, C, Cinit, bias, aux, M, N, K, lda, ldb, ldc, tA, tB, epi, epi_param, s);
  AIR_REQUIRE(aligned16(A) && aligned16(B) && (lda % 4 == 0) && (ldb % 4 == 0), AIR_ERR_BAD_ALIGN,
              "air_gemm(TF32): TMA needs 16-byte aligned operands and leading dimensions that are multiples of 4 "
              "(lda=%d ldb=%d)", lda, ldb);
  const bool a_mn = tA != 0;   // A stored [K,M]: M contiguous
  const bool b_mn = tB == 0;   // B stored [K,N]: N contiguous
  // Tile / split selection.  With fp32 operands a 128x128 tile moves 32 KB per k-block for 1 MFLOP (32 FLOP/B) and
  // the L2 -> SM fabric delivers ~12.5 TB/s, so the wide tile is worth it whenever it still gives every SM a CTA
  // (most get two: the short-K GEMMs of this model are fill / epilogue bound and want co-resident CTAs);
  // split-K for the very-long-K weight gradients with few output tiles, otherwise 64-wide tiles.
  // (Splitting K on medium-K shapes was measured slower: the second pass costs more than it saves.)
  const int sms = sm_count();
  const int mt = (M + kBM - 1) / kBM;
  const int64_t tiles128 = static_cast<int64_t>(mt) * ((N + 127) / 128), tiles64 = static_cast<int64_t>(mt) * ((N + 63) / 64);
  TcParams p;
  p.C = C; p.Cinit = Cinit; p.bias = bias; p.aux = aux;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.epi = epi; p.epi_param = epi_param;
  p.num_kb = (K + kBK - 1) / kBK;
  const bool can_split = (N % 4 == 0) && workspace != nullptr && aligned16(workspace);
  const int64_t ws_splits = can_split ? ws_floats / (static_cast<int64_t>(M) * N) : 1;
  auto want_splits = [&](int64_t tiles) {
    int sp = static_cast<int>(std::min<int64_t>((2 * sms + tiles - 1) / tiles, p.num_kb / 8));
    sp = static_cast<int>(std::min<int64_t>(sp, ws_splits));
    return std::max(1, std::min(sp, 32));
  };
  int BN, splits = 1;
  if (N > 64 && tiles128 >= sms) {
    BN = 128;  // at least one wide CTA per SM (most SMs get two)
  } else if (N > 64 && can_split && p.num_kb >= 64) {
    BN = 128;  // very long K, few tiles (weight gradients): wide tiles + split-K
    splits = want_splits(tiles128);
  } else {
    BN = 64;   // twice the CTAs of the wide tiling
    if (tiles64 < sms && can_split && p.num_kb >= 64) splits = want_splits(tiles64);
  }
  if (tc_env().bn) BN = (tc_env().bn == 64 || N <= 64) ? 64 : 128;  // tuning / diagnostics override
  p.kb_per_split = (p.num_kb + splits - 1) / splits;
  splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;  // no empty split
  p.splits = splits;
  if (splits > 1) p.C = workspace;

  CUtensorMap ma, mb;
  int rc;
  if (!a_mn) rc = make_tmap(&ma, A, K, M, lda, kBK, kBM / 4, false);  // [M,K] K contiguous: box {32 k, 32 m}
  else       rc = make_tmap(&ma, A, M, K, lda, 32, kBK, true);        // [K,M] M contiguous: box {32 m, 32 k}
  if (rc) return rc;
  if (!b_mn) rc = make_tmap(&mb, B, K, N, ldb, kBK, BN / 4, false);              // [N,K] K contiguous: box {32 k, BN/4 n}
  else       rc = make_tmap(&mb, B, N, K, ldb, 32, BN == 128 ? kBK : 16, true);  // [K,N] N contiguous: box {32 n, 32|16 k}
  if (rc) return rc;

  // Pairs of M-neighbour CTAs share their B tile through TMA multicast when the M-tile count is even and the
  // mainloop is long: measured (profiles/r1_gemm_cluster_multicast.md) +6.5 % / +9 % on the two biggest GEMMs
  // (K = 2500, 4096), neutral at ~77 k-blocks per CTA, and 3-5 % slower on the short-K GEMMs, where the two
  // cluster barriers and the gang launch cost more than the operand traffic saves.  AIR_TC_CLUSTER=0/2 forces off/on.
  const int cl_env = tc_env().cluster;
  const bool cl = cl_env != 0 && (mt % 2 == 0) && (cl_env == 2 || p.kb_per_split >= 64);
#define AIR_TC_DISPATCH2(BNv, CLv, X3v)                                                                                 \
  (a_mn ? (b_mn ? launch_tc<BNv, true, true, CLv, X3v>(ma, mb, p, s) : launch_tc<BNv, true, false, CLv, X3v>(ma, mb, p, s)) \
        : (b_mn ? launch_tc<BNv, false, true, CLv, X3v>(ma, mb, p, s) : launch_tc<BNv, false, false, CLv, X3v>(ma, mb, p, s)))
#define AIR_TC_DISPATCH(BNv, CLv) (x3 ? AIR_TC_DISPATCH2(BNv, CLv, true) : AIR_TC_DISPATCH2(BNv, CLv, false))
  if (cl) rc = (BN == 128) ? AIR_TC_DISPATCH(128, true) : AIR_TC_DISPATCH(64, true);
  else rc = (BN == 128) ? AIR_TC_DISPATCH(128, false) : AIR_TC_DISPATCH(64, false);
#undef AIR_TC_DISPATCH
#undef AIR_TC_DISPATCH2
  if (rc) return rc;
  if (splits > 1) {
    const int64_t total = static_cast<int64_t>(M) * N / 4;
    const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(sm_count()) * 8));
    AIR_LAUNCH(splitk_reduce_kernel, blocks, 256, 0, s, workspace, splits, C, Cinit, bias, aux, M, N, ldc, epi, epi_param);
    count_launch();
    rc = check_launch("splitk_reduce");
  }
  return rc;
}

}  // namespace air
