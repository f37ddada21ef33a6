// Shared host/device helpers for libair_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/air_b200.h"

namespace air {

// ---- host-side error plumbing ---------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int check_launch(const char *what);  // cudaGetLastError -> AIR_ERR_CUDA

#define AIR_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::air::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

// ---- programmatic dependent launch (PDL) ---------------------------------------------------
// Every kernel of this library is launched with programmaticStreamSerialization = 1 (unless
// AIR_PDL=0): its CTAs may be scheduled while the previous kernel in the stream drains, run their
// prologue (shared-memory carve-up, mbarrier init, TMEM allocation, tensor-map prefetch) and then
// block in pdl_wait() until the previous grid has completed and its memory is visible.  The rule
// that keeps this safe: NO global-memory access (read or write) before pdl_wait().
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define AIR_LAUNCH(kern, grid, block, smem, stream, ...) \
  (void)::air::launch_pdl(kern, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
int sm_count();  // cached per device

// device side of PDL: wait for the previous grid, then let the next grid start its own prologue
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() {
  pdl_wait();
  pdl_trigger();
}

// ---- exact (never FMA-contracted) fp32 arithmetic -------------------------------------
// The oracle rounds every TF op separately; these intrinsics are immune to -fmad.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

// tf.linspace(-1, 1, n)[i] = start + step*i in fp32 (transformer.py:127-130)
__device__ __forceinline__ float linspace_pm1(int i, int n) {
  if (n == 1) return -1.0f;
  const float step = __fdiv_rn(2.0f, static_cast<float>(n - 1));
  return add_rn(-1.0f, mul_rn(step, static_cast<float>(i)));
}

// floor -> int with the oracle's saturation, then the two clipped corners
__device__ __forceinline__ void floor_clip(float v, int maxi, int &i0, int &i1) {
  float f = floorf(v);
  f = fminf(fmaxf(f, -2.0e9f), 2.0e9f);  // NaN -> -2e9 like the C oracle
  if (!(f > -2.0e9f)) f = -2.0e9f;
  const int i = static_cast<int>(f);
  i0 = min(max(i, 0), maxi);
  i1 = min(max(i + 1, 0), maxi);
}

// ---- warp / block reductions (fixed order => deterministic) ---------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "AIR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra AIR_DONE;\n"
      "bra AIR_WAIT;\n"
      "AIR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// non-suspending poll (mbarrier.test_wait): for single-thread producer/consumer roles, where the
// suspend/resume latency of try_wait would be paid once per pipeline hand-off
__device__ __forceinline__ void mbar_wait_spin(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "AIR_SPIN:\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra AIR_SPIN_DONE;\n"
      "bra AIR_SPIN;\n"
      "AIR_SPIN_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace air
