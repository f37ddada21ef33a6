// Library plumbing for libair_b200.so: error strings, launch counter, device info.
#include <nmmintrin.h>
#include <stdarg.h>
#include <stdlib.h>

#include "air_common.cuh"

namespace air {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches += n; }

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return AIR_ERR_CUDA;
  }
  return AIR_OK;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("AIR_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace air

extern "C" int air_abi_version(void) { return 1; }

extern "C" const char *air_last_error(void) { return air::g_err; }

extern "C" int64_t air_launch_count(void) { return air::g_launches; }

extern "C" int air_device_info(int *sm_count, int *cc_major, int *cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) {
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e == cudaSuccess) {
      if (sm_count) *sm_count = p.multiProcessorCount;
      if (cc_major) *cc_major = p.major;
      if (cc_minor) *cc_minor = p.minor;
      return AIR_OK;
    }
  }
  air::set_error("air_device_info: %s (no CUDA device; this library has no CPU fallback)", cudaGetErrorString(e));
  return AIR_ERR_CUDA;
}

extern "C" uint32_t air_crc32c(const void *data, uint64_t nbytes, uint32_t crc) {
  const unsigned char *p = static_cast<const unsigned char *>(data);
  uint64_t c = crc ^ 0xFFFFFFFFu;
  while (nbytes && (reinterpret_cast<uintptr_t>(p) & 7)) {
    c = _mm_crc32_u8(static_cast<uint32_t>(c), *p++);
    --nbytes;
  }
  for (; nbytes >= 8; nbytes -= 8, p += 8) c = _mm_crc32_u64(c, *reinterpret_cast<const uint64_t *>(p));
  for (; nbytes; --nbytes) c = _mm_crc32_u8(static_cast<uint32_t>(c), *p++);
  return static_cast<uint32_t>(c) ^ 0xFFFFFFFFu;
}
