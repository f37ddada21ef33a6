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
#include "tc_common.cuh"

namespace air {

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
          // B: producer pi loads a quarter of the tile (one box; two 32-wide chunks of the 256-wide MN-major tile)
#pragma unroll
          for (int part = 0; part < ((B_MN && BN == 256) ? 2 : 1); ++part) {
            unsigned char *bdst;
            int bc0, bc1;
            if (!B_MN) {  // box {32 k, BN/4 n}
              bdst = sb + pi * (BN / 4 * 128); bc0 = k0; bc1 = n0 + pi * (BN / 4);
            } else if (BN >= 128) {  // box {32 n, 32 k}: chunk pi (and pi + 4)
              const int c = pi + 4 * part;
              bdst = sb + c * (kBK * 128); bc0 = n0 + c * 32; bc1 = k0;
            } else {      // BN == 64: box {32 n, 16 k}: chunk pi/2, k-half pi%2
              bdst = sb + (pi >> 1) * (kBK * 128) + (pi & 1) * (16 * 128); bc0 = n0 + (pi >> 1) * 32; bc1 = k0 + (pi & 1) * 16;
            }
            if (!CL) {
              tma_load_2d(bdst, &mapB, &full_bar[s], bc0, bc1);
            } else if (static_cast<uint32_t>(pi >> 1) == crank) {  // this CTA's half of the shared B tile, to both CTAs
              tma_load_2d_mc(bdst, &mapB, &full_bar[s], bc0, bc1, static_cast<uint16_t>(3));
            }
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
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    tc_epilogue<BN, X3>(p, tmem_acc, m0, n0, split, warp, lane, chains, kMaxChains * BN);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, kTmemCols);
  }
  if (CL) cluster_sync_all();  // no CTA leaves while its peer may still signal its barriers
}

// =========================================================================================
// Persistent variant (plain TF32, no cluster): one CTA per SM (or two) walks tiles t = blockIdx.x, + gridDim.x, ... with
//   * ONE operand ring that keeps running across tiles: the producer is already fetching the next tile's k-blocks while
//     the last MMAs of the current tile execute, so only the first tile of a CTA pays the TMA round trip;
//   * TWO accumulator stages in TMEM (2 x BN columns): the epilogue warps drain tile i (tcgen05.ld, Cinit / bias /
//     activation, global stores) while the tensor core already accumulates tile i + 1 into the other stage
//     (tmem_full / tmem_empty mbarriers, one phase bit per stage);
//   * barrier init, TMEM allocation, descriptor prefetch and the PDL wait once per CTA instead of once per tile.
// The short-K GEMMs of this model (K = 256 ... 784: 8 - 25 k-blocks per tile) spend as long in the epilogue as in the
// main loop; this overlaps the two.  Tile order: n fastest (the CTAs of a wave share their A rows through L2).
// =========================================================================================
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kTcThreads)
    gemm_tf32_persist_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcParams p,
                             int tiles_n, int tiles_mn, int tiles_total) {
  constexpr uint32_t kABytes = kBM * kBK * 4, kBBytes = BN * kBK * 4, kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = tmem_cols_for(2 * BN, false);
  const int kStages = p.stages;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kTcThreads / 32 - 2);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;
  pdl_sync();  // PDL: barriers, TMEM and descriptors were set up while the previous grid drained

  // tile t -> (split, m tile, n tile), its k-block range
  auto tile_coords = [&](int t, int &m0, int &n0, int &split, int &kb_begin, int &nkb) {
    split = t / tiles_mn;
    const int r = t - split * tiles_mn;
    const int mt = r / tiles_n;
    m0 = mt * kBM;
    n0 = (r - mt * tiles_n) * BN;
    kb_begin = split * p.kb_per_split;
    nkb = min(p.num_kb, kb_begin + p.kb_per_split) - kb_begin;
  };

  if (warp == 0) {
    // ================= TMA producer (one thread) =================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
        int m0, n0, split, kb_begin, nkb;
        tile_coords(t, m0, n0, split, kb_begin, nkb);
        for (int i = 0; i < nkb; ++i) {
          mbar_wait_spin(&empty_bar[s], ph ^ 1);
          unsigned char *sa = tiles + s * kStageBytes, *sb = sa + kABytes;
          mbar_expect_tx(&full_bar[s], kStageBytes);
          const int k0 = (kb_begin + i) * kBK;
#pragma unroll
          for (int pi = 0; pi < 4; ++pi) {
            if (!A_MN) tma_load_2d(sa + pi * (32 * 128), &mapA, &full_bar[s], k0, m0 + pi * 32);      // box {32 k, 32 m}
            else       tma_load_2d(sa + pi * (kBK * 128), &mapA, &full_bar[s], m0 + pi * 32, k0);     // box {32 m, 32 k}
            if (!B_MN) {            // box {32 k, BN/4 n}
              tma_load_2d(sb + pi * (BN / 4 * 128), &mapB, &full_bar[s], k0, n0 + pi * (BN / 4));
            } else if (BN >= 128) {  // box {32 n, 32 k}: chunk pi
              tma_load_2d(sb + pi * (kBK * 128), &mapB, &full_bar[s], n0 + pi * 32, k0);
            } else {                // BN == 64: box {32 n, 16 k}: chunk pi/2, k-half pi%2
              tma_load_2d(sb + (pi >> 1) * (kBK * 128) + (pi & 1) * (16 * 128), &mapB, &full_bar[s], n0 + (pi >> 1) * 32,
                          k0 + (pi & 1) * 16);
            }
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, A_MN, B_MN);
      constexpr uint32_t a_lbo = A_MN ? kBK * 128 : 16, a_sbo = A_MN ? 512 : 1024, a_adv = A_MN ? 1024 : 32;
      constexpr uint32_t b_lbo = B_MN ? kBK * 128 : 16, b_sbo = B_MN ? 512 : 1024, b_adv = B_MN ? 1024 : 32;
      constexpr uint32_t a_lt = A_MN ? 1 : 2, b_lt = B_MN ? 1 : 2;
      int s = 0, acc = 0;
      uint32_t ph = 0, acc_ph = 0;
      for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
        int m0, n0, split, kb_begin, nkb;
        tile_coords(t, m0, n0, split, kb_begin, nkb);
        mbar_wait_spin(&tmem_empty_bar[acc], acc_ph ^ 1);  // the epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_acc + acc * BN;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait_spin(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + s * kStageBytes), sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            const uint64_t da = make_smem_desc(sa + k * a_adv, a_lbo, a_sbo, a_lt);
            const uint64_t db = make_smem_desc(sb + k * b_adv, b_lbo, b_sbo, b_lt);
            umma_tf32(tmem_d, da, db, idesc, (i | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator of this tile complete
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: TMEM -> registers -> global (8 warps), one tile behind the MMAs =================
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
      int m0, n0, split, kb_begin, nkb;
      tile_coords(t, m0, n0, split, kb_begin, nkb);
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      tc_epilogue<BN, false>(p, tmem_acc + acc * BN, m0, n0, split, warp, lane, 1, 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, kTmemCols);
  }
}

// =========================================================================================
// Persistent kernel on 2 x 2 CLUSTERS with TMA multicast (plain TF32, 128 x 128 tiles).
// The short-K GEMMs of this model are bound by L2 -> SM operand traffic: with fp32 operands a 128 x 128 tile moves
// 32 KB per k-block for 1 MFLOP (profiles/r2_gemm_short_k_ncu_full.md: 340 MB through the fabric at ~7 TB/s while DRAM
// supplies 65 MB, tensor pipe 20 %).  Here four CTAs own a 256 x 256 block of the output as 2 x 2 tiles of 128 x 128:
// the two CTAs of a block ROW need the same A rows -- each loads half of that A tile and multicasts it to both; the two
// CTAs of a block COLUMN need the same B columns -- likewise.  Every SM then receives its full 32 KB per k-block but only
// 16 KB of it cross the fabric on its behalf: half the L2 -> SM bytes per FLOP of the one-CTA kernel, the traffic of a
// 256 x 256 tile with the scheduling granularity of 128 x 128 ones.  A stage may be refilled once the three CTAs its
// loads write to (self, row peer, column peer) have consumed it: the empty barriers count three multicast commits.
// Persistence and the two TMEM accumulator stages are those of gemm_tf32_persist_kernel; the cluster walks block
// g = cluster id, + number of clusters, ...; tiles outside the matrix (odd tile counts) load zeros and store nothing.
// =========================================================================================
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kTcThreads)
    gemm_tf32_cluster4_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, TcParams p,
                              int blocks_n, int blocks_total) {
  constexpr int BN = 128;
  constexpr uint32_t kABytes = kBM * kBK * 4, kBBytes = BN * kBK * 4, kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = tmem_cols_for(2 * BN, false);
  const int kStages = p.stages;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *tiles = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();          // 0..3: (block row, block column) = (crank >> 1, crank & 1)
  const int rm = static_cast<int>(crank >> 1), rn = static_cast<int>(crank & 1);
  const uint16_t mask_a = static_cast<uint16_t>((1u << crank) | (1u << (crank ^ 1u)));   // the CTAs that share my A rows
  const uint16_t mask_b = static_cast<uint16_t>((1u << crank) | (1u << (crank ^ 2u)));   // ... my B columns
  const uint16_t mask_done = static_cast<uint16_t>(mask_a | mask_b);                     // whom my loads write to
  const int cluster_id = blockIdx.x >> 2, n_clusters = gridDim.x >> 2;

  if (threadIdx.x == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 3);  // self + row peer + column peer
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], kTcThreads / 32 - 2);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_slot;
  cluster_sync_all();  // every CTA's barriers exist before anything is multicast to them
  pdl_sync();          // PDL: barriers, TMEM and descriptors were set up while the previous grid drained

  auto block_coords = [&](int g, int &m0, int &n0) {
    const int gm = g / blocks_n, gn = g - gm * blocks_n;
    m0 = (2 * gm + rm) * kBM;
    n0 = (2 * gn + rn) * BN;
  };
  const int nkb = p.num_kb;

  if (warp == 0) {
    // ================= TMA producer (one thread): my half of my A tile -> row pair, my half of my B tile -> column pair
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int g = cluster_id; g < blocks_total; g += n_clusters) {
        int m0, n0;
        block_coords(g, m0, n0);
        for (int i = 0; i < nkb; ++i) {
          mbar_wait_spin(&empty_bar[s], ph ^ 1);
          unsigned char *sa = tiles + s * kStageBytes, *sb = sa + kABytes;
          mbar_expect_tx(&full_bar[s], kStageBytes);  // both halves of both tiles land here (two of them from the peers)
          const int k0 = i * kBK;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int pa = 2 * rn + h;  // quarter boxes of the A tile this CTA fetches (its row peer fetches the others)
            if (!A_MN) tma_load_2d_mc(sa + pa * (32 * 128), &mapA, &full_bar[s], k0, m0 + pa * 32, mask_a);
            else       tma_load_2d_mc(sa + pa * (kBK * 128), &mapA, &full_bar[s], m0 + pa * 32, k0, mask_a);
            const int pb = 2 * rm + h;  // ... of the B tile (its column peer fetches the others)
            if (!B_MN) tma_load_2d_mc(sb + pb * (BN / 4 * 128), &mapB, &full_bar[s], k0, n0 + pb * (BN / 4), mask_b);
            else       tma_load_2d_mc(sb + pb * (kBK * 128), &mapB, &full_bar[s], n0 + pb * 32, k0, mask_b);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(kBM, BN, A_MN, B_MN);
      constexpr uint32_t a_lbo = A_MN ? kBK * 128 : 16, a_sbo = A_MN ? 512 : 1024, a_adv = A_MN ? 1024 : 32;
      constexpr uint32_t b_lbo = B_MN ? kBK * 128 : 16, b_sbo = B_MN ? 512 : 1024, b_adv = B_MN ? 1024 : 32;
      constexpr uint32_t a_lt = A_MN ? 1 : 2, b_lt = B_MN ? 1 : 2;
      int s = 0, acc = 0;
      uint32_t ph = 0, acc_ph = 0;
      for (int g = cluster_id; g < blocks_total; g += n_clusters) {
        mbar_wait_spin(&tmem_empty_bar[acc], acc_ph ^ 1);  // the epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_acc + acc * BN;
        for (int i = 0; i < nkb; ++i) {
          mbar_wait_spin(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + s * kStageBytes), sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 8; ++k) {
            const uint64_t da = make_smem_desc(sa + k * a_adv, a_lbo, a_sbo, a_lt);
            const uint64_t db = make_smem_desc(sb + k * b_adv, b_lbo, b_sbo, b_lt);
            umma_tf32(tmem_d, da, db, idesc, (i | k) != 0 ? 1u : 0u);
          }
          umma_commit_mc(&empty_bar[s], mask_done);  // this stage of mine is free for me and for the peers that write into it
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);  // accumulator of this tile complete
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue (8 warps), one tile behind the MMAs =================
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int g = cluster_id; g < blocks_total; g += n_clusters) {
      int m0, n0;
      block_coords(g, m0, n0);
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      if (m0 < p.M && n0 < p.N) tc_epilogue<BN, false>(p, tmem_acc + acc * BN, m0, n0, 0, warp, lane, 1, 0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, kTmemCols);
  }
  cluster_sync_all();  // no CTA leaves while a peer may still multicast into its shared memory or signal its barriers
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
      C[o] = apply_epilogue(v, epi, epi_aux_value(aux, epi, o, e + j), epi_param);
    }
  }
}

static int launch_splitk_reduce(const float *ws, int splits, float *C, const float *Cinit, const float *bias, const float *aux,
                                int M, int N, int ldc, int epi, float epi_param, cudaStream_t s) {
  const int64_t total = static_cast<int64_t>(M) * N / 4;
  const int blocks = static_cast<int>(std::min<int64_t>((total + 255) / 256, static_cast<int64_t>(sm_count()) * 8));
  AIR_LAUNCH(splitk_reduce_kernel, blocks, 256, 0, s, ws, splits, C, Cinit, bias, aux, M, N, ldc, epi, epi_param);
  count_launch();
  return check_launch("splitk_reduce");
}

int gemm_tf32_pair(const float *A, const float *B, TcParams p, int lda, int ldb, bool a_mn, bool b_mn, int BNP, bool x3,
                   cudaStream_t s);  // gemm_tc2.cu

// 128 x 256 tiles of the one-CTA kernel (plain TF32; AIR_TC_BN=256 forces them).  Measured per shape
// (profiles/r2_gemm_shapes.md): they win on the long-K split-K weight gradients that the pair kernel does not take
// (dW g2 / r2 / Kh: 18 / 18 / 23 us against 27 / 27 / 30 with 128-wide tiles) and lose on the short-K GEMMs, where
// fewer, fatter CTAs leave SMs idle.
static bool use_bn256(int M, int N, int num_kb, int splits, int sms) {
  (void)M; (void)splits; (void)sms;
  return num_kb >= 64 && N >= 256;
}

// Width of the CTA-pair tile (0 = use the one-CTA kernel).  From the per-shape measurements of the model's 26 GEMMs
// (profiles/r2_gemm_shapes.md): in plain TF32 the pair kernel wins where the main loop is long and the grid is wide
// (the image projection and its weight gradient: 34 vs 48 us, 54 vs 69 us); the short-K GEMMs are bound by per-CTA
// fixed costs and prefer two co-resident one-CTA tiles.  In the 3xTF32 mode every k-block costs three MMAs and two
// extra shared-memory passes, so halving the operand traffic pays almost everywhere (-35 % over the step).
constexpr int kX3MaxChainKb = 26;  // longest hi*hi accumulation chain (k-blocks) of the 256-wide X3 pair tile, see below
static int pair_width(int M, int N, int num_kb, bool x3) {
  const int env = tc_env().pair;
  if (env == 0 || M <= kBM || N < 128) return 0;
  if (env == 128 || env == 256) return N > 128 ? env : 128;
  if (!x3) return (num_kb >= 64 && N >= 512 && M >= 512) ? 256 : 0;
  if (num_kb <= 4) return 0;
  if (num_kb >= 64) return N > 128 ? 256 : 128;
  return N >= 768 ? 256 : 128;
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
  p.flags = tc_env().flags;
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

// Persistent launch: ctas_per_sm resident CTAs per SM (1: the whole shared memory as one deep ring; 2: two shallower
// rings whose prologues / epilogues interleave), never more CTAs than tiles.
template <int BN, bool A_MN, bool B_MN>
static int launch_tc_persist(const CUtensorMap &ma, const CUtensorMap &mb, TcParams p, int ctas_per_sm, cudaStream_t s) {
  auto kern = gemm_tf32_persist_kernel<BN, A_MN, B_MN>;
  const size_t stage_bytes = static_cast<size_t>(kBM + BN) * kBK * 4;
  const int max_stages = static_cast<int>(std::min<size_t>(kMaxStages, 200 * 1024 / stage_bytes));
  const int tiles_n = (p.N + BN - 1) / BN, tiles_mn = tiles_n * ((p.M + kBM - 1) / kBM), total = tiles_mn * p.splits;
  const size_t budget = (ctas_per_sm == 1 ? 200 : 100) * 1024;
  int stages = std::max(2, std::min(static_cast<int>(budget / stage_bytes), max_stages));
  if (tc_env().stages) stages = std::min(tc_env().stages, max_stages);
  p.stages = stages;
  p.chains = 1;
  p.flags = tc_env().flags;
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + 1024;
  const size_t smem_max = static_cast<size_t>(max_stages) * stage_bytes + 1024;
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = uint64_t(1) << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_max));
    AIR_REQUIRE(e == cudaSuccess, AIR_ERR_CUDA, "cudaFuncSetAttribute(gemm_tf32_persist): %s", cudaGetErrorString(e));
    configured.fetch_or(bit, std::memory_order_release);
  }
  const int grid = std::min(total, sm_count() * ctas_per_sm);
  AIR_LAUNCH(kern, grid, kTcThreads, smem, s, ma, mb, p, tiles_n, tiles_mn, total);
  count_launch();
  return check_launch("gemm_tf32_persist");
}

// 2 x 2 cluster launch: as many clusters as the device can keep resident (one CTA per SM: the whole shared memory as one
// six-stage ring), never more than there are 256 x 256 blocks.
template <bool A_MN, bool B_MN>
static int launch_tc_cluster4(const CUtensorMap &ma, const CUtensorMap &mb, TcParams p, cudaStream_t s) {
  auto kern = gemm_tf32_cluster4_kernel<A_MN, B_MN>;
  constexpr int BN = 128;
  const size_t stage_bytes = static_cast<size_t>(kBM + BN) * kBK * 4;
  const int max_stages = static_cast<int>(std::min<size_t>(kMaxStages, 200 * 1024 / stage_bytes));
  int stages = max_stages;
  if (tc_env().stages) stages = std::min(tc_env().stages, max_stages);
  p.stages = std::max(2, stages);
  p.chains = 1;
  p.flags = tc_env().flags;
  const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + 1024;
  const size_t smem_max = static_cast<size_t>(max_stages) * stage_bytes + 1024;
  const int blocks_n = ((p.N + BN - 1) / BN + 1) / 2, blocks_m = ((p.M + kBM - 1) / kBM + 1) / 2, total = blocks_n * blocks_m;
  static std::atomic<uint64_t> configured{0};
  static int max_clusters[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = uint64_t(1) << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_max));
    AIR_REQUIRE(e == cudaSuccess, AIR_ERR_CUDA, "cudaFuncSetAttribute(gemm_tf32_cluster4): %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * 64);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem_max;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    AIR_REQUIRE(e == cudaSuccess && n > 0, AIR_ERR_CUDA, "cudaOccupancyMaxActiveClusters(gemm_tf32_cluster4): %s (%d)",
                cudaGetErrorString(e), n);
    max_clusters[dev & 63] = n;
    configured.fetch_or(bit, std::memory_order_release);
  }
  const int clusters = std::min(total, max_clusters[dev & 63]);
  cudaError_t e = launch_cluster_pdl(kern, dim3(4 * clusters), dim3(kTcThreads), smem, s, dim3(4, 1, 1), ma, mb, p, blocks_n, total);
  AIR_REQUIRE(e == cudaSuccess, AIR_ERR_CUDA, "gemm_tf32_cluster4 launch: %s", cudaGetErrorString(e));
  count_launch();
  return check_launch("gemm_tf32_cluster4");
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
  p.C = C; p.Cout = C; p.Cinit = Cinit; p.bias = bias; p.aux = aux;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.epi = epi; p.epi_param = epi_param;
  p.num_kb = (K + kBK - 1) / kBK;
  // Split-K needs the caller's workspace for the partial tiles [splits][M][N]; a second, machine-wide pass adds them in
  // increasing split order.  (Letting the last CTA of each tile do that sum inside the GEMM kernel -- one launch less --
  // was measured 3-4x SLOWER on the weight-gradient shapes: one CTA per tile reads 2 MB of partials at L2 latency while
  // the other SMs idle; profiles/r2_gemm_shapes.md.)
  const bool can_split = (N % 4 == 0) && workspace != nullptr && aligned16(workspace);
  const int64_t ws_splits = can_split ? ws_floats / (static_cast<int64_t>(M) * N) : 1;
  auto want_splits = [&](int64_t tiles) {
    int sp = static_cast<int>(std::min<int64_t>((tc_env().split_waves * sms + tiles - 1) / tiles, p.num_kb / 8));
    sp = static_cast<int>(std::min<int64_t>(sp, ws_splits));
    return std::max(1, std::min(sp, 32));
  };
  // CTA pairs (gemm_tc2.cu, 256 x 128|256 tiles): half the L2 -> SM operand bytes per FLOP.  AIR_TC_PAIR: 0 = never,
  // 128 / 256 = force that pair width wherever a pair fits; default = automatic (see pair_width()).
  int BNP = pair_width(M, N, p.num_kb, x3);
  if (BNP) {
    int64_t ctas = 2 * static_cast<int64_t>((M + 2 * kBM - 1) / (2 * kBM)) * ((N + BNP - 1) / BNP);
    int splits = 1;
    if (2 * ctas <= sms && can_split && p.num_kb >= 64) splits = want_splits(ctas);  // (a nearly full wave stays unsplit)
    if (x3 && BNP == 256 && (p.num_kb + splits - 1) / splits > kX3MaxChainKb) {
      // The 256-wide X3 tile has TMEM columns for ONE hi*hi accumulator (256 + 256 correction = 512), and the tensor
      // core's accumulate step rounds toward zero: a chain of n k-blocks is biased by ~2 n ulp.  Keep chains short by
      // splitting K (the partial sums are added in FP32 round-to-nearest by the second pass), or use the 128-wide
      // tile with its three round-robin accumulators when there is no workspace for that.
      const int need = (p.num_kb + kX3MaxChainKb - 1) / kX3MaxChainKb;
      if (can_split && need <= ws_splits && need <= 32) splits = std::max(splits, need);
      else { BNP = 128; ctas *= 2; splits = 1; if (2 * ctas <= sms && can_split && p.num_kb >= 64) splits = want_splits(ctas); }
    }
    p.kb_per_split = (p.num_kb + splits - 1) / splits;
    splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
    p.splits = splits;
    if (splits > 1) p.C = workspace;
    int rc = gemm_tf32_pair(A, B, p, lda, ldb, a_mn, b_mn, BNP, x3, s);
    if (rc) return rc;
    return splits > 1 ? launch_splitk_reduce(workspace, splits, C, Cinit, bias, aux, M, N, ldc, epi, epi_param, s) : AIR_OK;
  }
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
  // 128 x 256 tiles (plain TF32 only: the X3 accumulators would not fit TMEM): one MMA reads A 4 KB + B 8 KB for twice
  // the FLOPs of the 128-wide one -- 96 instead of 128 bytes of shared-memory operand reads per tensor-core clock
  if (!x3 && BN == 128 && use_bn256(M, N, p.num_kb, splits, sms)) BN = 256;
  if (tc_env().bn) BN = (tc_env().bn == 64 || N <= 64) ? 64 : (tc_env().bn == 256 && !x3 && N > 128) ? 256 : 128;  // diagnostics override
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
  else       rc = make_tmap(&mb, B, N, K, ldb, 32, BN >= 128 ? kBK : 16, true);  // [K,N] N contiguous: box {32 n, 32|16 k}
  if (rc) return rc;

  // Pairs of M-neighbour CTAs share their B tile through TMA multicast when the M-tile count is even and the
  // mainloop is long: measured (profiles/r1_gemm_cluster_multicast.md) +6.5 % / +9 % on the two biggest GEMMs
  // (K = 2500, 4096), neutral at ~77 k-blocks per CTA, and 3-5 % slower on the short-K GEMMs, where the two
  // cluster barriers and the gang launch cost more than the operand traffic saves.  AIR_TC_CLUSTER=0/2 forces off/on.
  const int cl_env = tc_env().cluster;
  const bool cl = cl_env != 0 && (mt % 2 == 0) && (cl_env == 2 || p.kb_per_split >= 64);
  // Persistent kernel (two TMEM accumulator stages: the epilogue of tile i overlaps the main loop of tile i + 1) for
  // the plain-TF32 GEMMs whose tiles outnumber the resident CTAs.  AIR_TC_PERSIST: 0 = never, 1 / 2 = that many CTAs
  // per SM wherever it applies, default = automatic.
  // 2 x 2 multicast clusters (AIR_TC_CLUSTER4: 0 = never, 1 = wherever it applies; default = automatic)
  if (!x3 && !cl && BN == 128 && splits == 1 && tc_env().cluster4 != 0) {
    const int64_t blocks4 = static_cast<int64_t>((mt + 1) / 2) * (((N + 127) / 128 + 1) / 2);
    const bool auto_on = false;
    if (tc_env().cluster4 > 0 || auto_on) {
      (void)blocks4;
      rc = a_mn ? (b_mn ? launch_tc_cluster4<true, true>(ma, mb, p, s) : launch_tc_cluster4<true, false>(ma, mb, p, s))
                : (b_mn ? launch_tc_cluster4<false, true>(ma, mb, p, s) : launch_tc_cluster4<false, false>(ma, mb, p, s));
      return rc;
    }
  }
  if (!x3 && !cl && BN <= 128 && tc_env().persist != 0) {
    const int64_t total = (BN == 128 ? tiles128 : tiles64) * splits;
    const int per_sm = tc_env().persist > 0 ? std::min(tc_env().persist, 2) : 1;
    // Automatic choice, from the per-shape measurements at the model's sizes (profiles/r2_gemm_shapes.md, M = 12288):
    // one persistent CTA per SM wins where every CTA gets 2+ wide tiles AND the epilogue is heavy -- the dX GEMMs with
    // a K-major weight operand (dX gm 49 -> 41 us, dX r2 30 -> 26, dX r1 39 -> 37);
    // it loses 5-25 % on the 256-wide layers and on fwd r1 / g2 (the non-persistent kernel keeps two CTAs per
    // SM there, whose prologues and epilogues already interleave).
    // (not the gen_mean layer with its in-epilogue noise generator, AIR_EPI_SIGMOID_RNG: eight epilogue warps per SM
    // instead of sixteen doubled that GEMM in the real step, 40 -> 79 us, although the same shape with a plain epilogue
    // gained 4 us in the per-shape table)
    const bool auto_on = BN == 128 && splits == 1 && total >= 2 * static_cast<int64_t>(sms) && p.num_kb >= 8 && !b_mn &&
                         N >= 512 && epi != AIR_EPI_SIGMOID_RNG;
    if (tc_env().persist > 0 || auto_on) {
#define AIR_TC_PERSIST_DISPATCH(BNv)                                                                                             \
  (a_mn ? (b_mn ? launch_tc_persist<BNv, true, true>(ma, mb, p, per_sm, s) : launch_tc_persist<BNv, true, false>(ma, mb, p, per_sm, s)) \
        : (b_mn ? launch_tc_persist<BNv, false, true>(ma, mb, p, per_sm, s) : launch_tc_persist<BNv, false, false>(ma, mb, p, per_sm, s)))
      rc = BN == 128 ? AIR_TC_PERSIST_DISPATCH(128) : AIR_TC_PERSIST_DISPATCH(64);
#undef AIR_TC_PERSIST_DISPATCH
      if (rc) return rc;
      return splits > 1 ? launch_splitk_reduce(workspace, splits, C, Cinit, bias, aux, M, N, ldc, epi, epi_param, s) : AIR_OK;
    }
  }
#define AIR_TC_DISPATCH2(BNv, CLv, X3v)                                                                                 \
  (a_mn ? (b_mn ? launch_tc<BNv, true, true, CLv, X3v>(ma, mb, p, s) : launch_tc<BNv, true, false, CLv, X3v>(ma, mb, p, s)) \
        : (b_mn ? launch_tc<BNv, false, true, CLv, X3v>(ma, mb, p, s) : launch_tc<BNv, false, false, CLv, X3v>(ma, mb, p, s)))
#define AIR_TC_DISPATCH(BNv, CLv) (x3 ? AIR_TC_DISPATCH2(BNv, CLv, true) : AIR_TC_DISPATCH2(BNv, CLv, false))
  if (BN == 256) rc = cl ? AIR_TC_DISPATCH2(256, true, false) : AIR_TC_DISPATCH2(256, false, false);
  else if (cl) rc = (BN == 128) ? AIR_TC_DISPATCH(128, true) : AIR_TC_DISPATCH(64, true);
  else rc = (BN == 128) ? AIR_TC_DISPATCH(128, false) : AIR_TC_DISPATCH(64, false);
#undef AIR_TC_DISPATCH
#undef AIR_TC_DISPATCH2
  if (rc) return rc;
  return splits > 1 ? launch_splitk_reduce(workspace, splits, C, Cinit, bias, aux, M, N, ldc, epi, epi_param, s) : AIR_OK;
}

}  // namespace air
