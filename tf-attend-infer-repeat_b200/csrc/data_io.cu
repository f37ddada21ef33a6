// Host-side input path for the reference's multi-MNIST TFRecord files (multi_mnist.py:186-251): the reference feeds
// the model from TFRecordReader + parse_single_example + shuffle_batch with 4 reader threads (training.py:28, 76-81).
// Here the file is memory-mapped by the caller and
//   air_tfrecord_index   scans it ONCE: record framing (length + masked CRC-32C), the tf.train.Example wire format, and
//                        records where every example's `image` payload lies and its `digits` value;
//   air_shuffle_order    the order in which a shuffle queue of a given capacity emits record indices;
//   air_gather_rows      copies the selected payloads into one contiguous (pinned) batch buffer with several threads.
// No CUDA here: these run on the host cores next to the GPU (a 4096-image batch is 41 MB).
#include <nmmintrin.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "air_common.cuh"

namespace air {
namespace {

uint32_t masked_crc(const uint8_t *p, uint64_t n) {
  const uint32_t c = air_crc32c(p, n, 0);
  return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

// protobuf wire-format cursor over [p, end)
struct Cur {
  const uint8_t *p, *end;
  bool ok = true;
  bool more() const { return ok && p < end; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 70 && p < end; shift += 7) {
      const uint8_t b = *p++;
      v |= static_cast<uint64_t>(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  // next field: returns field number, sets wire type and (for length-delimited) the payload range
  int field(int *wt, Cur *payload, uint64_t *scalar) {
    const uint64_t tag = varint();
    *wt = static_cast<int>(tag & 7);
    if (!ok) return -1;
    if (*wt == 0) {
      *scalar = varint();
    } else if (*wt == 2) {
      const uint64_t n = varint();
      if (!ok || n > static_cast<uint64_t>(end - p)) { ok = false; return -1; }
      payload->p = p; payload->end = p + n; payload->ok = true;
      p += n;
    } else if (*wt == 5) {
      if (end - p < 4) { ok = false; return -1; }
      p += 4;
    } else if (*wt == 1) {
      if (end - p < 8) { ok = false; return -1; }
      p += 8;
    } else {
      ok = false;
      return -1;
    }
    return static_cast<int>(tag >> 3);
  }
};

// Example{1: Features{1: map entry{1: key, 2: Feature{1: BytesList{1: bytes} | 3: Int64List{1: varint (packed or not)}}}}}
bool parse_example(const uint8_t *rec, uint64_t n, const uint8_t **image, uint32_t *image_len, int64_t *digits) {
  *image = nullptr; *image_len = 0; *digits = -1;
  Cur ex{rec, rec + n};
  while (ex.more()) {
    int wt; Cur feats{nullptr, nullptr}; uint64_t sc;
    const int f = ex.field(&wt, &feats, &sc);
    if (f < 0) return false;
    if (f != 1 || wt != 2) continue;
    while (feats.more()) {
      Cur entry{nullptr, nullptr};
      const int f2 = feats.field(&wt, &entry, &sc);
      if (f2 < 0) return false;
      if (f2 != 1 || wt != 2) continue;
      const uint8_t *key = nullptr; uint64_t key_len = 0; Cur feat{nullptr, nullptr}; bool have_feat = false;
      while (entry.more()) {
        Cur v{nullptr, nullptr};
        const int f3 = entry.field(&wt, &v, &sc);
        if (f3 < 0) return false;
        if (f3 == 1 && wt == 2) { key = v.p; key_len = static_cast<uint64_t>(v.end - v.p); }
        if (f3 == 2 && wt == 2) { feat = v; have_feat = true; }
      }
      if (!key || !have_feat) continue;
      const bool is_image = key_len == 5 && memcmp(key, "image", 5) == 0;
      const bool is_digits = key_len == 6 && memcmp(key, "digits", 6) == 0;
      if (!is_image && !is_digits) continue;
      while (feat.more()) {
        Cur lst{nullptr, nullptr};
        const int kind = feat.field(&wt, &lst, &sc);
        if (kind < 0) return false;
        if (wt != 2) continue;
        while (lst.more()) {
          Cur v{nullptr, nullptr};
          const int f4 = lst.field(&wt, &v, &sc);
          if (f4 < 0) return false;
          if (f4 != 1) continue;
          if (is_image && kind == 1 && wt == 2) { *image = v.p; *image_len = static_cast<uint32_t>(v.end - v.p); }
          if (is_digits && kind == 3) {
            if (wt == 0) *digits = static_cast<int64_t>(sc);
            else if (wt == 2 && v.more()) { *digits = static_cast<int64_t>(v.varint()); if (!v.ok) return false; }
          }
        }
      }
    }
  }
  return ex.ok;
}

inline uint64_t splitmix(uint64_t &s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace
}  // namespace air

using namespace air;

extern "C" int64_t air_tfrecord_index(const uint8_t *buf, uint64_t nbytes, int verify_crc, int64_t max_records,
                                      uint64_t *image_off, uint32_t *image_len, int32_t *digits) {
  if (!buf && nbytes) { set_error("tfrecord_index: null buffer"); return AIR_ERR_NULL; }
  uint64_t pos = 0;
  int64_t n = 0;
  while (pos < nbytes) {
    if (nbytes - pos < 12) { set_error("tfrecord_index: truncated record header at byte %llu", (unsigned long long)pos); return AIR_ERR_BAD_SHAPE; }
    uint64_t len; uint32_t hcrc;
    memcpy(&len, buf + pos, 8);
    memcpy(&hcrc, buf + pos + 8, 4);
    if (verify_crc && masked_crc(buf + pos, 8) != hcrc) { set_error("tfrecord_index: corrupt record length at byte %llu", (unsigned long long)pos); return AIR_ERR_BAD_SHAPE; }
    if (len > nbytes - pos - 12 || nbytes - pos - 12 - len < 4) { set_error("tfrecord_index: truncated record at byte %llu", (unsigned long long)pos); return AIR_ERR_BAD_SHAPE; }
    const uint8_t *rec = buf + pos + 12;
    if (verify_crc) {
      uint32_t dcrc;
      memcpy(&dcrc, rec + len, 4);
      if (masked_crc(rec, len) != dcrc) { set_error("tfrecord_index: corrupt record data at byte %llu", (unsigned long long)pos); return AIR_ERR_BAD_SHAPE; }
    }
    if (image_off || image_len || digits) {
      if (n >= max_records) { set_error("tfrecord_index: more than %lld records", (long long)max_records); return AIR_ERR_BAD_SHAPE; }
      const uint8_t *img; uint32_t ilen; int64_t d;
      if (!parse_example(rec, len, &img, &ilen, &d) || !img || d < 0) {
        set_error("tfrecord_index: record %lld is not a tf.train.Example with `image` and `digits` features", (long long)n);
        return AIR_ERR_BAD_SHAPE;
      }
      if (image_off) image_off[n] = static_cast<uint64_t>(img - buf);
      if (image_len) image_len[n] = ilen;
      if (digits) digits[n] = static_cast<int32_t>(d);
    }
    ++n;
    pos += 12 + len + 4;
  }
  return n;  // (with all three outputs NULL: just the record count, for sizing them)
}

extern "C" int air_shuffle_order(int64_t n, int64_t buffer, int epochs, uint64_t seed, int64_t *order) {
  if (n < 0 || epochs < 0 || (n * epochs > 0 && !order)) { set_error("shuffle_order: bad arguments"); return AIR_ERR_BAD_SHAPE; }
  // tf.train.shuffle_batch (multi_mnist.py:240-249): records enter a queue; once it holds more than `buffer`
  // (min_after_dequeue) elements a uniformly chosen one leaves; at the end of the input the queue drains the same way.
  std::vector<int64_t> q;
  q.reserve(static_cast<size_t>(std::max<int64_t>(buffer, 0)) + 1);
  uint64_t s = seed;
  int64_t out = 0;
  auto emit = [&]() {
    const size_t i = static_cast<size_t>(splitmix(s) % q.size());
    order[out++] = q[i];
    q[i] = q.back();
    q.pop_back();
  };
  for (int e = 0; e < epochs; ++e)
    for (int64_t r = 0; r < n; ++r) {
      q.push_back(r);
      if (static_cast<int64_t>(q.size()) > buffer) emit();
    }
  while (!q.empty()) emit();
  return AIR_OK;
}

extern "C" int air_gather_rows(const uint8_t *src, const uint64_t *offsets, int64_t n, uint64_t row_bytes, uint8_t *dst,
                               int n_threads) {
  if (n < 0 || (n > 0 && (!src || !offsets || !dst))) { set_error("gather_rows: null pointer"); return AIR_ERR_NULL; }
  const int64_t per_thread_min = std::max<int64_t>(1, static_cast<int64_t>((1 << 20) / std::max<uint64_t>(row_bytes, 1)));  // >= 1 MB each
  int t = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(std::max(n_threads, 1), n / per_thread_min)));
  auto work = [=](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) memcpy(dst + static_cast<uint64_t>(i) * row_bytes, src + offsets[i], row_bytes);
  };
  if (t == 1) {
    work(0, n);
    return AIR_OK;
  }
  std::vector<std::thread> pool;
  const int64_t chunk = (n + t - 1) / t;
  for (int k = 0; k < t; ++k) {
    const int64_t lo = k * chunk, hi = std::min<int64_t>(n, lo + chunk);
    if (lo < hi) pool.emplace_back(work, lo, hi);
  }
  for (auto &th : pool) th.join();
  return AIR_OK;
}
