// CTA-pair tensor-core GEMM for sm_100a: tcgen05.mma.cta_group::2 -- two CTAs on the two SMs of a TPC compute one
// 256 x BNP output tile (BNP = 128 or 256).  Each CTA stages its own 128 rows of A and HALF of the B tile
// (BNP/2 columns) in its shared memory; the leader CTA (cluster rank 0) issues every MMA, the tensor cores of both
// SMs read both halves of B, and each CTA's TMEM receives its 128 accumulator rows.  Against the one-CTA kernel of
// gemm_tc.cu (128 x 128 tiles) a 256 x 256 pair tile moves half the operand bytes per FLOP from L2 into the SMs --
// the measured limit of the one-CTA kernel with fp32 operands (profiles/r1_gemm_tf32_big_ncu_full.md: 705 MB of
// L2 -> SM traffic at the fabric's ~12.5 TB/s) -- and halves the shared-memory reads of B.
//
// Same operand layouts, epilogues, split-K and 3xTF32 mode (X3, see gemm_tc.cu) as the one-CTA kernel.  Pipeline:
//   warp 0 lane 0 (both CTAs)   TMA: own A tile + own half of B into stage s, completes the local full_bar[s]
//   X3: warps 2..9 (both CTAs)  write the lo tiles of stage s, then arrive (remotely for the peer) on the LEADER's
//                               ready_bar[s] (16 arrivals = 8 warps x 2 CTAs)
//   TF32: the TMA loads of BOTH CTAs complete on the leader's full_bar[s] (cp.async.bulk.tensor.cta_group::2)
//   warp 1 lane 0, leader       waits full_bar[s] (TF32) / ready_bar[s] (X3), issues the MMAs, commits with a multicast
//                               arrive on empty_bar[s] of both CTAs; at the end on tmem_full_bar of both
//   warps 2..9 (both CTAs)      epilogue of the CTA's own 128 rows (tc_epilogue)
#include "tc_common.cuh"

namespace air {

__device__ __forceinline__ void tmem_alloc2(uint32_t *smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this offset in every CTA of cta_mask once all MMAs issued so far have completed
__device__ __forceinline__ void umma_commit2_mc(uint64_t *bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// shared::cluster address of the leader CTA's (cluster rank 0) copy of a shared-memory variable
__device__ __forceinline__ uint32_t leader_addr(const void *p) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(0));
  return r;
}
// Arrive on a (possibly remote) mbarrier by its shared::cluster address.  Default semantics, as CUTLASS's
// ClusterBarrier::arrive(cta_id): an explicit .release.cluster here / .acquire.cluster on the waiting side compile to
// cluster-scope memory barriers that cost ~1 us per k-block on the MMA thread (measured: 96 us -> see
// profiles/r2_gemm_pair.md); the operands are consumed by the tensor core through the async proxy, ordered behind
// the barrier observation by tcgen05.fence::after_thread_sync, and every writer fences its own shared memory
// (TMA completion / fence.proxy.async) before it arrives.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// TMA tile load whose completion is signalled on an mbarrier given by its shared::cluster address -- with
// .cta_group::2 that may be the PEER CTA's barrier: both CTAs of the pair report their bytes to the leader's full
// barrier directly, no forwarding thread in between.
__device__ __forceinline__ void tma_load_2d_2sm(void *smem_dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

constexpr int kMaxStages2 = 8;

template <int BNP, bool A_MN, bool B_MN, bool X3>
__global__ void __launch_bounds__(kTcThreads)
    gemm_tf32_pair_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcParams p) {
  constexpr int HB = BNP / 2;  // columns of B staged by each CTA
  constexpr uint32_t kABytes = kBM * kBK * 4, kBBytes = HB * kBK * 4;
  constexpr uint32_t kRawBytes = kABytes + kBBytes;
  constexpr uint32_t kStageBytes = X3 ? 2 * kRawBytes : kRawBytes;
  constexpr int kChainCap = X3 ? 512 / BNP - 1 : 1;  // hi*hi chains that fit beside the correction accumulator
  constexpr uint32_t kTmemCols = X3 ? 512u : static_cast<uint32_t>(BNP);
  const int kStages = p.stages;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages2], empty_bar[kMaxStages2], ready_bar[kMaxStages2], tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int m0 = blockIdx.x * kBM;                 // this CTA's 128 rows (the pair = two consecutive M tiles, adjacent in x)
  const int n0 = blockIdx.y * BNP;                 // the pair's BNP columns
  const int nb0 = n0 + static_cast<int>(crank) * HB;  // the half of B this CTA stages
  const int split = blockIdx.z;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;
  const int chains = X3 ? min(min(p.chains, kChainCap), nkb) : 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&ready_bar[s], (p.flags & 1) ? 2 : 2 * kSplitWarps);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc2(&tmem_base_slot, kTmemCols);  // collective: one warp of each CTA of the pair
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;
  cluster_sync_all();  // both CTAs' barriers exist before anything is signalled across the pair
  pdl_sync();

  if (warp == 0) {
    // ================= TMA producer (each CTA: its A rows, its half of B) =================
    // TF32: both CTAs' loads complete on the LEADER's full_bar[s] (which expects the bytes of both).
    // X3:   each CTA's loads complete on its own full_bar[s]: its splitter warps wait there.
    if (lane == 0) {
      const uint32_t full0 = X3 ? smem_u32(&full_bar[0]) : leader_addr(&full_bar[0]);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait_spin(&empty_bar[s], ph ^ 1);
        unsigned char *sa = tiles + s * kStageBytes, *sb = sa + kABytes;
        if (X3) mbar_expect_tx(&full_bar[s], kRawBytes);
        else if (leader) mbar_expect_tx(&full_bar[s], 2 * kRawBytes);
        const uint32_t bar = full0 + s * 8;
        const int k0 = (kb_begin + i) * kBK;
        if (!A_MN) {
          tma_load_2d_2sm(sa, &mapA, bar, k0, m0);  // box {32 k, 128 m}
        } else {
#pragma unroll
          for (int c = 0; c < kBM / 32; ++c)        // box {32 m, 32 k} per 32-wide chunk
            tma_load_2d_2sm(sa + c * (kBK * 128), &mapA, bar, m0 + c * 32, k0);
        }
        if (!B_MN) {
          tma_load_2d_2sm(sb, &mapB, bar, k0, nb0);  // box {32 k, HB n}
        } else {
#pragma unroll
          for (int c = 0; c < HB / 32; ++c)          // box {32 n, 32 k} per 32-wide chunk
            tma_load_2d_2sm(sb + c * (kBK * 128), &mapB, bar, nb0 + c * 32, k0);
        }
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // ================= MMA issuer (leader CTA only) =================
      constexpr uint32_t idesc = make_idesc_tf32(2 * kBM, BNP, A_MN, B_MN);
      constexpr uint32_t a_lbo = A_MN ? kBK * 128 : 16, a_sbo = A_MN ? 512 : 1024, a_adv = A_MN ? 1024 : 32;
      constexpr uint32_t b_lbo = B_MN ? kBK * 128 : 16, b_sbo = B_MN ? 512 : 1024, b_adv = B_MN ? 1024 : 32;
      constexpr uint32_t a_lt = A_MN ? 1 : 2, b_lt = B_MN ? 1 : 2;
      const uint32_t tmem_corr = tmem_acc + kChainCap * BNP;
      int s = 0, chain = 0;
      uint32_t ph = 0;
      for (int i = 0; i < nkb; ++i) {
        if (X3) mbar_wait_spin(&ready_bar[s], ph);  // both CTAs' lo tiles are written (which implies their operands landed)
        else mbar_wait_spin(&full_bar[s], ph);      // both CTAs' operands landed
        tc_fence_after();
        const uint32_t sa = smem_u32(tiles + s * kStageBytes), sb = sa + kABytes;
        const uint32_t tmem_main = tmem_acc + (X3 ? chain * BNP : 0);
#pragma unroll
        for (int k = 0; k < kBK / 8; ++k) {
          const uint64_t da = make_smem_desc(sa + k * a_adv, a_lbo, a_sbo, a_lt);
          const uint64_t db = make_smem_desc(sb + k * b_adv, b_lbo, b_sbo, b_lt);
          umma_tf32_2sm(tmem_main, da, db, idesc, (i >= chains || k != 0) ? 1u : 0u);
        }
        if (X3) {
          if (++chain == chains) chain = 0;
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            const uint64_t da = make_smem_desc(sa + k * a_adv, a_lbo, a_sbo, a_lt);
            const uint64_t db = make_smem_desc(sb + k * b_adv, b_lbo, b_sbo, b_lt);
            const uint64_t dal = make_smem_desc(sa + kRawBytes + k * a_adv, a_lbo, a_sbo, a_lt);
            const uint64_t dbl = make_smem_desc(sb + kRawBytes + k * b_adv, b_lbo, b_sbo, b_lt);
            umma_tf32_2sm(tmem_corr, dal, db, idesc, (i | k) != 0 ? 1u : 0u);
            umma_tf32_2sm(tmem_corr, da, dbl, idesc, 1u);
          }
        }
        umma_commit2_mc(&empty_bar[s], static_cast<uint16_t>(3));  // frees stage s in both CTAs
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
      umma_commit2_mc(&tmem_full_bar, static_cast<uint16_t>(3));  // accumulators complete, both CTAs
    }
    __syncwarp();
  } else {
    if (X3) {
      // ================= 3xTF32 splitter (both CTAs): lo tiles of stage s, then arrive on the leader's ready_bar[s] ===
      constexpr int kVec = kRawBytes / 16, kThreads = kSplitWarps * 32;
      static_assert(kVec % kThreads == 0, "split loop assumes a whole number of float4 per thread");
      const uint32_t ready0 = leader_addr(&ready_bar[0]);
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
        fence_proxy_async();
        if (p.flags & 1) {  // experiment: one remote arrival per CTA and stage instead of one per warp
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (tid == 0) mbar_arrive_cluster(ready0 + s * 8);
        } else {
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(ready0 + s * 8);
        }
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
    // ================= epilogue: this CTA's 128 rows =================
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    tc_epilogue<BNP, X3>(p, tmem_acc, m0, n0, split, warp, lane, chains, kChainCap * BNP);
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();  // both CTAs have read their accumulators and nothing is in flight across the pair
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_acc, kTmemCols);
  }
}

template <int BNP, bool A_MN, bool B_MN, bool X3>
static int launch_pair(const CUtensorMap &ma, const CUtensorMap &mb, TcParams p, cudaStream_t s) {
  auto kern = gemm_tf32_pair_kernel<BNP, A_MN, B_MN, X3>;
  const size_t stage_bytes = static_cast<size_t>(kBM + BNP / 2) * kBK * 4 * (X3 ? 2 : 1);
  const int max_stages = static_cast<int>(std::min<size_t>(kMaxStages2, 200 * 1024 / stage_bytes));
  int stages = std::max(2, std::min(max_stages, std::max(p.kb_per_split, 2)));
  if (tc_env().stages) stages = std::min(tc_env().stages, max_stages);
  p.stages = stages;
  p.chains = tc_env().chains;
  p.flags = tc_env().flags;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024;
  const size_t smem_max = static_cast<size_t>(max_stages) * stage_bytes + 1024;
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = uint64_t(1) << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_max));
    AIR_REQUIRE(e == cudaSuccess, AIR_ERR_CUDA, "cudaFuncSetAttribute(gemm_tf32_pair): %s", cudaGetErrorString(e));
    configured.fetch_or(bit, std::memory_order_release);
  }
  // the CTA pair of cta_group::2 is a cluster of two CTAs adjacent in x
  dim3 grid(2 * ((p.M + 2 * kBM - 1) / (2 * kBM)), (p.N + BNP - 1) / BNP, p.splits);
  cudaError_t e = launch_cluster_pdl(kern, grid, dim3(kTcThreads), smem, s, dim3(2, 1, 1), ma, mb, p);
  AIR_REQUIRE(e == cudaSuccess, AIR_ERR_CUDA, "gemm_tf32_pair cluster launch: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch(X3 ? "gemm_tf32x3_pair" : "gemm_tf32_pair");
}

// Launches the CTA-pair kernel for an already chosen tiling (p.splits, p.kb_per_split, p.C = workspace when split).
int gemm_tf32_pair(const float *A, const float *B, TcParams p, int lda, int ldb, bool a_mn, bool b_mn, int BNP, bool x3,
                   cudaStream_t s) {
  CUtensorMap ma, mb;
  int rc;
  const int HB = BNP / 2;
  if (!a_mn) rc = make_tmap(&ma, A, p.K, p.M, lda, kBK, kBM, false);  // [M,K]: box {32 k, 128 m}
  else       rc = make_tmap(&ma, A, p.M, p.K, lda, 32, kBK, true);    // [K,M]: box {32 m, 32 k}
  if (rc) return rc;
  if (!b_mn) rc = make_tmap(&mb, B, p.K, p.N, ldb, kBK, HB, false);   // [N,K]: box {32 k, HB n}
  else       rc = make_tmap(&mb, B, p.N, p.K, ldb, 32, kBK, true);    // [K,N]: box {32 n, 32 k}
  if (rc) return rc;
#define AIR_PAIR_DISPATCH2(BNv, X3v)                                                                          \
  (a_mn ? (b_mn ? launch_pair<BNv, true, true, X3v>(ma, mb, p, s) : launch_pair<BNv, true, false, X3v>(ma, mb, p, s)) \
        : (b_mn ? launch_pair<BNv, false, true, X3v>(ma, mb, p, s) : launch_pair<BNv, false, false, X3v>(ma, mb, p, s)))
#define AIR_PAIR_DISPATCH(BNv) (x3 ? AIR_PAIR_DISPATCH2(BNv, true) : AIR_PAIR_DISPATCH2(BNv, false))
  rc = (BNP == 256) ? AIR_PAIR_DISPATCH(256) : AIR_PAIR_DISPATCH(128);
#undef AIR_PAIR_DISPATCH
#undef AIR_PAIR_DISPATCH2
  return rc;
}

}  // namespace air
