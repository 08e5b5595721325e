// libctr_feed.so -- TFRecord framing, Example / SequenceExample wire parsing and vocabulary mapping on the host
// (include/ctr_feed.h).  C++17, no dependencies beyond the standard library.
#include "../../include/ctr_feed.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <new>
#include <system_error>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <nmmintrin.h>
#endif

namespace {

thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ------------------------------------------------------------------------------------------------ CRC-32C
struct CrcTables {
  uint32_t t[8][256];
  CrcTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
};
const CrcTables& tables() {
  static const CrcTables T;
  return T;
}

uint32_t crc32c_sw(uint32_t c, const uint8_t* p, uint64_t n) {     // slice-by-8
  const CrcTables& T = tables();
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = T.t[7][w & 0xFF] ^ T.t[6][(w >> 8) & 0xFF] ^ T.t[5][(w >> 16) & 0xFF] ^ T.t[4][(w >> 24) & 0xFF] ^
        T.t[3][(w >> 32) & 0xFF] ^ T.t[2][(w >> 40) & 0xFF] ^ T.t[1][(w >> 48) & 0xFF] ^ T.t[0][w >> 56];
    p += 8; n -= 8;
  }
  while (n--) c = T.t[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c;
}

#if defined(__x86_64__)
__attribute__((target("sse4.2"))) uint32_t crc32c_hw(uint32_t c, const uint8_t* p, uint64_t n) {
  uint64_t c64 = c;
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    c64 = _mm_crc32_u64(c64, w);
    p += 8; n -= 8;
  }
  c = (uint32_t)c64;
  while (n--) c = _mm_crc32_u8(c, *p++);
  return c;
}
#endif

uint32_t crc32c(const uint8_t* p, uint64_t n) {
#if defined(__x86_64__)
  static const bool hw = __builtin_cpu_supports("sse4.2");
  if (hw) return crc32c_hw(0xFFFFFFFFu, p, n) ^ 0xFFFFFFFFu;
#endif
  return crc32c_sw(0xFFFFFFFFu, p, n) ^ 0xFFFFFFFFu;
}
uint32_t masked(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xA282EAD8u; }

// ------------------------------------------------------------------------------------------------ string -> int table
// Multiply-mix hash over 8-byte words (the keys are short tokens such as "userid_8": one or two words).
inline uint64_t hash_bytes(const uint8_t* p, uint64_t n) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ (n * 0xFF51AFD7ED558CCDull);
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    h = (h ^ w) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
    p += 8; n -= 8;
  }
  if (n) {
    uint64_t w = 0;
    memcpy(&w, p, n);
    h = (h ^ w) * 0xC4CEB9FE1A85EC53ull;
    h ^= h >> 29;
  }
  h *= 0x9E3779B97F4A7C15ull;
  return h ^ (h >> 32);
}

// Open addressing, linear probing, load <= 1/2.  A slot holds the key's place in `base` (the table does not own the bytes), its
// length, 32 bits of its hash and the value: one cache line decides most probes, the key bytes are touched once, to confirm.
// The first insert of a key wins (categorical_column_with_vocabulary_file: id = line of the FIRST occurrence).
struct FlatMap {
  struct Slot {
    uint64_t off;
    uint32_t len;
    uint32_t tag;        // 0 = empty; hashes are forced non-zero
    int64_t value;
  };
  std::vector<Slot> slots;
  uint64_t mask = 0;
  const uint8_t* base = nullptr;

  static uint32_t tag_of(uint64_t h) { const uint32_t t = (uint32_t)(h >> 32); return t ? t : 1u; }
  void init(const uint8_t* b, size_t n_keys) {
    base = b;
    size_t cap = 8;
    while (cap < 2 * n_keys) cap <<= 1;
    slots.assign(cap, Slot{0, 0, 0, 0});
    mask = cap - 1;
  }
  void insert(uint64_t off, uint32_t len, int64_t value) {
    const uint64_t h = hash_bytes(base + off, len);
    const uint32_t tag = tag_of(h);
    for (uint64_t i = h & mask;; i = (i + 1) & mask) {
      Slot& s = slots[i];
      if (!s.tag) { s = Slot{off, len, tag, value}; return; }
      if (s.tag == tag && s.len == len && memcmp(base + s.off, base + off, len) == 0) return;     // first occurrence wins
    }
  }
  int64_t find(const uint8_t* p, uint64_t n, int64_t missing) const {
    if (slots.empty() || n > 0xFFFFFFFFull) return missing;
    const uint64_t h = hash_bytes(p, n);
    const uint32_t tag = tag_of(h);
    for (uint64_t i = h & mask;; i = (i + 1) & mask) {
      const Slot& s = slots[i];
      if (!s.tag) return missing;
      if (s.tag == tag && s.len == (uint32_t)n && memcmp(base + s.off, p, n) == 0) return s.value;
    }
  }
};

// ------------------------------------------------------------------------------------------------ vocabulary
struct Vocab {
  std::string blob;
  FlatMap map;                                            // keys live in `blob`
  int64_t size = 0;
  void build(std::vector<std::pair<uint64_t, uint64_t>>& spans) {
    size = (int64_t)spans.size();
    map.init(reinterpret_cast<const uint8_t*>(blob.data()), spans.size());
    for (int64_t i = 0; i < size; ++i)
      if (spans[i].second <= 0xFFFFFFFFull) map.insert(spans[i].first, (uint32_t)spans[i].second, i);
  }
  int64_t find(const uint8_t* p, uint64_t n) const { return map.find(p, n, -1); }
};

// ------------------------------------------------------------------------------------------------ protobuf wire helpers
struct Span {
  const uint8_t* p = nullptr;
  uint64_t n = 0;
};

inline bool varint(const uint8_t*& p, const uint8_t* end, uint64_t& v) {
  v = 0;
  for (int shift = 0; shift < 64 && p < end; shift += 7) {
    const uint8_t b = *p++;
    v |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) return true;
  }
  return false;
}

// next field of a message: number, wire type and payload (length-delimited: the bytes; varint: value in `val`; fixed: bytes)
inline bool next_field(const uint8_t*& p, const uint8_t* end, uint32_t& fnum, uint32_t& wt, Span& payload, uint64_t& val) {
  uint64_t tag;
  if (!varint(p, end, tag)) return false;
  fnum = (uint32_t)(tag >> 3);
  wt = (uint32_t)(tag & 7);
  switch (wt) {
    case 0: return varint(p, end, val);
    case 1: if (end - p < 8) return false; payload = {p, 8}; p += 8; return true;
    case 2: {
      uint64_t len;
      if (!varint(p, end, len) || (uint64_t)(end - p) < len) return false;
      payload = {p, len}; p += len; return true;
    }
    case 5: if (end - p < 4) return false; payload = {p, 4}; p += 4; return true;
    default: return false;     // groups are not used by these protos
  }
}

enum Kind { K_NONE = 0, K_BYTES = 1, K_FLOAT = 2, K_INT64 = 3 };

// Feature{ oneof{ BytesList bytes_list=1; FloatList float_list=2; Int64List int64_list=3 } }: which list, and its bytes
inline bool feature_list_of(Span feat, Kind& kind, Span& list) {
  const uint8_t* p = feat.p;
  const uint8_t* end = feat.p + feat.n;
  kind = K_NONE;
  list = {};
  uint32_t f, wt; Span pl; uint64_t v;
  while (p < end) {
    if (!next_field(p, end, f, wt, pl, v)) return false;
    if (wt == 2 && f >= 1 && f <= 3) { kind = (Kind)f; list = pl; }      // last one wins (oneof)
  }
  return true;
}

struct KeySpec {
  std::string_view key;
  bool is_cat;
  int index;          // into cats[] or dense[]
};

struct RecordView {                  // the final Feature bytes of every spec key in one record (proto map: last entry wins)
  std::vector<Span> ctx;             // per spec key
  std::vector<char> has_ctx;
  std::vector<Span> flist;           // per spec key: FeatureList bytes (only when read_feature_lists)
  std::vector<char> has_fl;
  // Writers emit the keys of every record in the same order: the j-th map entry of this record is most likely the key the j-th
  // entry of the previous record was.  Remember that key's bytes (they stay valid: they point into the caller's buffer) and what it
  // resolved to, so that a match costs one memcmp instead of a hash + probe.  Two maps (context, feature_lists), numbered apart.
  struct Hint { Span key; int index; };
  std::vector<Hint> hints[2];
};

struct Parser {
  const std::vector<KeySpec>& keys;
  std::string key_bytes;             // the spec keys back to back: FlatMap does not own its keys
  FlatMap index;
  bool read_fl;
  Parser(const std::vector<KeySpec>& k, bool fl) : keys(k), read_fl(fl) {
    std::vector<uint64_t> at;
    for (const KeySpec& ks : k) { at.push_back(key_bytes.size()); key_bytes.append(ks.key.data(), ks.key.size()); }
    index.init(reinterpret_cast<const uint8_t*>(key_bytes.data()), k.size());
    for (int i = 0; i < (int)k.size(); ++i) index.insert(at[i], (uint32_t)k[i].key.size(), i);
  }

  // map<string, X> entries of `msg` (field 1 = entry{key=1, value=2}); records the value span of every spec key
  bool scan_map(Span msg, std::vector<Span>& out, std::vector<char>& has, std::vector<RecordView::Hint>& hints) const {
    const uint8_t* p = msg.p;
    const uint8_t* end = msg.p + msg.n;
    uint32_t f, wt; Span pl; uint64_t v;
    size_t j = 0;
    while (p < end) {
      if (!next_field(p, end, f, wt, pl, v)) return false;
      if (f != 1 || wt != 2) continue;
      const uint8_t* q = pl.p;
      const uint8_t* qe = pl.p + pl.n;
      Span key{}, val{};
      uint32_t f2, wt2; Span pl2; uint64_t v2;
      while (q < qe) {
        if (!next_field(q, qe, f2, wt2, pl2, v2)) return false;
        if (wt2 == 2 && f2 == 1) key = pl2;
        else if (wt2 == 2 && f2 == 2) val = pl2;
      }
      int idx;
      if (j < hints.size() && hints[j].key.n == key.n && (key.n == 0 || memcmp(hints[j].key.p, key.p, key.n) == 0)) {
        idx = hints[j].index;
        hints[j].key = key;                                  // keep pointing at recent bytes (warm cache lines)
      } else {
        idx = (int)index.find(key.p, key.n, -1);
        if (j < hints.size()) hints[j] = {key, idx};
        else if (j == hints.size() && j < 256) hints.push_back({key, idx});
      }
      ++j;
      if (idx >= 0) { out[idx] = val; has[idx] = 1; }
    }
    return true;
  }

  bool view(Span rec, RecordView& rv) const {
    const size_t K = keys.size();
    rv.ctx.assign(K, Span{}); rv.has_ctx.assign(K, 0);
    rv.flist.assign(K, Span{}); rv.has_fl.assign(K, 0);
    const uint8_t* p = rec.p;
    const uint8_t* end = rec.p + rec.n;
    uint32_t f, wt; Span pl; uint64_t v;
    while (p < end) {
      if (!next_field(p, end, f, wt, pl, v)) return false;
      if (wt != 2) continue;
      if (f == 1) {                          // Example.features == SequenceExample.context
        if (!scan_map(pl, rv.ctx, rv.has_ctx, rv.hints[0])) return false;
      } else if (f == 2 && read_fl) {        // SequenceExample.feature_lists (unknown field 2 when parsed as an Example)
        if (!scan_map(pl, rv.flist, rv.has_fl, rv.hints[1])) return false;
      }
    }
    return true;
  }
};

// values of a BytesList: calls fn(ptr, len) per value
template <typename Fn>
inline bool for_each_bytes(Span list, Fn&& fn) {
  const uint8_t* p = list.p;
  const uint8_t* end = list.p + list.n;
  uint32_t f, wt; Span pl; uint64_t v;
  while (p < end) {
    if (!next_field(p, end, f, wt, pl, v)) return false;
    if (f == 1 && wt == 2) fn(pl.p, pl.n);
  }
  return true;
}

// a categorical key of one record: context Feature first, else (optionally) every step of its FeatureList
template <typename Fn>
inline int cat_values(const RecordView& rv, int k, bool read_fl, const char* key, Fn&& fn) {
  if (rv.has_ctx[k]) {
    Kind kind; Span list;
    if (!feature_list_of(rv.ctx[k], kind, list)) return fail(CTR_FEED_ERR_PROTO, "Key: %s. Can't parse serialized Example.", key);
    if (kind == K_NONE) return CTR_FEED_OK;
    if (kind != K_BYTES) return fail(CTR_FEED_ERR_PROTO, "Key: %s. Data types don't match. Expected type: string", key);
    if (!for_each_bytes(list, fn)) return fail(CTR_FEED_ERR_PROTO, "Key: %s. Can't parse serialized Example.", key);
    return CTR_FEED_OK;
  }
  if (read_fl && rv.has_fl[k]) {             // FeatureList{ repeated Feature feature = 1 }
    const uint8_t* p = rv.flist[k].p;
    const uint8_t* end = p + rv.flist[k].n;
    uint32_t f, wt; Span pl; uint64_t v;
    while (p < end) {
      if (!next_field(p, end, f, wt, pl, v)) return fail(CTR_FEED_ERR_PROTO, "Key: %s. Can't parse serialized SequenceExample.", key);
      if (f != 1 || wt != 2) continue;
      Kind kind; Span list;
      if (!feature_list_of(pl, kind, list)) return fail(CTR_FEED_ERR_PROTO, "Key: %s. Can't parse serialized SequenceExample.", key);
      if (kind == K_NONE) continue;
      if (kind != K_BYTES) return fail(CTR_FEED_ERR_PROTO, "Key: %s. Data types don't match. Expected type: string", key);
      if (!for_each_bytes(list, fn)) return fail(CTR_FEED_ERR_PROTO, "Key: %s. Can't parse serialized SequenceExample.", key);
    }
  }
  return CTR_FEED_OK;
}

// FloatList{ repeated float value = 1 [packed] }: packed (wire type 2) or one fixed32 per value (wire type 5)
inline int dense_values(const RecordView& rv, int k, const ctr_feed_dense_t& d, int64_t b) {
  float* dst = d.out + b * d.width;
  for (int64_t i = 0; i < d.width; ++i) dst[i] = d.default_value;
  if (!rv.has_ctx[k]) return CTR_FEED_OK;
  Kind kind; Span list;
  if (!feature_list_of(rv.ctx[k], kind, list)) return fail(CTR_FEED_ERR_PROTO, "Key: %s. Can't parse serialized Example.", d.key);
  if (kind == K_NONE) return CTR_FEED_OK;
  if (kind != K_FLOAT) return fail(CTR_FEED_ERR_PROTO, "Key: %s. Data types don't match. Expected type: float", d.key);
  int64_t n = 0;
  float tmp;
  const uint8_t* p = list.p;
  const uint8_t* end = list.p + list.n;
  uint32_t f, wt; Span pl; uint64_t v;
  while (p < end) {
    if (!next_field(p, end, f, wt, pl, v)) return fail(CTR_FEED_ERR_PROTO, "Key: %s. Can't parse serialized Example.", d.key);
    if (f != 1) continue;
    if (wt == 2) {
      if (pl.n % 4) return fail(CTR_FEED_ERR_PROTO, "Key: %s. Can't parse serialized Example.", d.key);
      for (uint64_t o = 0; o < pl.n; o += 4, ++n)
        if (n < d.width) { memcpy(&tmp, pl.p + o, 4); dst[n] = tmp; }
    } else if (wt == 5) {
      if (n < d.width) { memcpy(&tmp, pl.p, 4); dst[n] = tmp; }
      ++n;
    }
  }
  if (n == 0) { for (int64_t i = 0; i < d.width; ++i) dst[i] = d.default_value; return CTR_FEED_OK; }
  if (n != d.width)
    return fail(CTR_FEED_ERR_PROTO, "Key: %s. Can't parse serialized Example: expected %lld values, got %lld", d.key, (long long)d.width,
                (long long)n);
  return CTR_FEED_OK;
}

// dataset.shuffle(buffer_size): the buffer holds input positions; `next` is the first position that has not entered it yet.
struct ShuffleState {
  std::vector<int64_t> slots;
  int64_t cap, filled = 0, next = 0;
  bool started = false;                                    // emission begins once the buffer is full or the input is done
  explicit ShuffleState(int64_t capacity) : cap(capacity) {}
  int64_t emit(int64_t n_available, bool done, const double* draws, int64_t max_out, int64_t* out) {
    if (!started) {
      if (slots.size() < (size_t)std::min(cap, n_available)) slots.resize((size_t)std::min(cap, n_available));
      while (filled < cap && next < n_available) slots[filled++] = next++;
      if (filled < cap && !done) return 0;
      started = true;
    }
    int64_t k = 0;
    while (k < max_out && filled > 0) {
      const bool more = next < n_available;
      if (!more && !done) break;                           // cannot tell yet whether position `next` exists
      const double u = draws[k];
      if (!(u >= 0.0 && u < 1.0)) return fail(CTR_FEED_ERR_ARG, "shuffle: draw %lld of this call is outside [0,1)", (long long)k);
      const int64_t j = std::min<int64_t>((int64_t)(u * (double)filled), filled - 1);
      out[k++] = slots[j];
      if (more) slots[j] = next++;
      else slots[j] = slots[--filled];
    }
    return k;
  }
};

// No C++ exception may cross the C ABI (or leave a std::thread: that is std::terminate): turn them into error codes.
template <typename Fn>
int guarded(Fn&& fn) {
  try {
    return fn();
  } catch (const std::bad_alloc&) {
    return fail(CTR_FEED_ERR_NOMEM, "out of memory");
  } catch (const std::exception& e) {
    return fail(CTR_FEED_ERR_NOMEM, "unexpected C++ exception: %s", e.what());
  } catch (...) {
    return fail(CTR_FEED_ERR_NOMEM, "unexpected C++ exception");
  }
}

template <typename Fn>
int run_chunks(int64_t B, int nthreads, Fn&& fn) {              // fn(thread, b0, b1) -> rc; first error wins
  std::vector<int> rcs(nthreads, CTR_FEED_OK);
  std::vector<std::string> msgs(nthreads);
  std::vector<std::thread> th;
  const int64_t per = (B + nthreads - 1) / nthreads;
  auto body = [&](int t) {
    const int64_t b0 = t * per, b1 = std::min<int64_t>(B, b0 + per);
    if (b0 < b1) rcs[t] = guarded([&] { return fn(t, b0, b1); });
    if (rcs[t] != CTR_FEED_OK) msgs[t] = g_err;                 // g_err is thread-local: carry the text to the caller
  };
  int created = 1;
  try {
    for (int t = 1; t < nthreads; ++t, ++created) th.emplace_back(body, t);
  } catch (const std::system_error&) {                          // the process ran out of threads: finish the rest here
  }
  body(0);
  for (int t = created; t < nthreads; ++t) body(t);
  for (auto& x : th) x.join();
  for (int t = 0; t < nthreads; ++t)
    if (rcs[t] != CTR_FEED_OK) return fail(rcs[t], "%s", msgs[t].c_str());
  return CTR_FEED_OK;
}

// Two phases over the same partition with ONE set of threads: every thread runs phase1 on its run of records, the last one to
// arrive runs `middle` (which needs all of phase 1: prefix sums), then every thread runs phase2 on the same run -- unless phase 1
// or `middle` failed.  Saves the second round of thread creation, which costs as much as the work itself at 32 threads.
// The threads are created parked; if the process cannot create them all, they are dismissed and the call runs on the caller alone.
template <typename F1, typename M, typename F2>
int run_two_phase(int64_t B, int nthreads, F1&& phase1, M&& middle, F2&& phase2) {
  std::vector<int> rcs(nthreads, CTR_FEED_OK);
  std::vector<std::string> msgs(nthreads);
  std::vector<std::thread> th;
  const int64_t per = (B + nthreads - 1) / nthreads;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0, middle_rc = CTR_FEED_OK;
  bool released = false, go = false, cancel = false;
  std::string middle_msg;
  auto body = [&](int t) {
    if (t != 0) {                                               // parked until every thread exists
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return go || cancel; });
      if (cancel) return;
    }
    const int64_t b0 = t * per, b1 = std::min<int64_t>(B, b0 + per);
    if (b0 < b1) rcs[t] = guarded([&] { return phase1(t, b0, b1); });
    if (rcs[t] != CTR_FEED_OK) msgs[t] = g_err;                 // g_err is thread-local: carry the text to the caller
    {
      std::unique_lock<std::mutex> lk(mu);
      if (++arrived == nthreads) {
        bool ok = true;
        for (int i = 0; i < nthreads; ++i) ok = ok && rcs[i] == CTR_FEED_OK;
        if (ok) {
          middle_rc = guarded([&] { return middle(); });
          if (middle_rc != CTR_FEED_OK) middle_msg = g_err;
        } else {
          middle_rc = CTR_FEED_ERR_ARG;                          // placeholder: phase 2 is skipped, phase 1's error is reported
        }
        released = true;
        cv.notify_all();
      } else {
        cv.wait(lk, [&] { return released; });
      }
    }
    if (middle_rc == CTR_FEED_OK && b0 < b1) {
      const int r = guarded([&] { return phase2(t, b0, b1); });
      if (r != CTR_FEED_OK) { rcs[t] = r; msgs[t] = g_err; }
    }
  };
  bool all_created = true;
  try {
    for (int t = 1; t < nthreads; ++t) th.emplace_back(body, t);
  } catch (const std::system_error&) {
    all_created = false;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    (all_created ? go : cancel) = true;
  }
  cv.notify_all();
  if (!all_created) {
    for (auto& x : th) x.join();
    return run_two_phase(B, 1, phase1, middle, phase2);          // nthreads = 1 creates no thread: cannot recurse further
  }
  body(0);
  for (auto& x : th) x.join();
  for (int t = 0; t < nthreads; ++t)
    if (rcs[t] != CTR_FEED_OK) return fail(rcs[t], "%s", msgs[t].c_str());
  if (middle_rc != CTR_FEED_OK) return fail(middle_rc, "%s", middle_msg.c_str());
  return CTR_FEED_OK;
}

}  // namespace

extern "C" {

const char* ctr_feed_last_error(void) { return g_err; }
int ctr_feed_version(void) { return 1; }

uint32_t ctr_feed_crc32c(const uint8_t* data, uint64_t n) { return (data || n == 0) ? crc32c(data, n) : 0; }
uint32_t ctr_feed_masked_crc32c(const uint8_t* data, uint64_t n) { return masked(ctr_feed_crc32c(data, n)); }

int64_t ctr_feed_tfrecord_index_from(const uint8_t* buf, uint64_t n, uint64_t start, int verify_crc, int allow_partial_tail,
                                     uint64_t* offsets, uint64_t* lengths, int64_t max_records, uint64_t* consumed) {
  if ((!buf && n) || start > n || max_records < 0 || (max_records > 0 && (!offsets || !lengths)))
    return fail(CTR_FEED_ERR_ARG, "ctr_feed_tfrecord_index: bad arguments");
  uint64_t pos = start;
  int64_t count = 0;
  while (pos < n) {
    if (consumed) *consumed = pos;                         // also on an error return: the damaged record starts here
    if (max_records > 0 && count == max_records) break;
    if (n - pos < 12) {
      if (allow_partial_tail) break;
      return fail(CTR_FEED_ERR_TRUNCATED, "truncated record header at byte %llu", (unsigned long long)pos);
    }
    uint64_t len;
    uint32_t len_crc;
    memcpy(&len, buf + pos, 8);
    memcpy(&len_crc, buf + pos + 8, 4);
    if (verify_crc && masked(crc32c(buf + pos, 8)) != len_crc)
      return fail(CTR_FEED_ERR_CRC, "corrupted record length at byte %llu (crc mismatch)", (unsigned long long)pos);
    if (len > n - pos - 12 || n - pos - 12 - len < 4) {
      if (allow_partial_tail) break;
      return fail(CTR_FEED_ERR_TRUNCATED, "truncated record at byte %llu", (unsigned long long)pos);
    }
    const uint8_t* data = buf + pos + 12;
    uint32_t data_crc;
    memcpy(&data_crc, data + len, 4);
    if (verify_crc == 1 && masked(crc32c(data, len)) != data_crc)          // verify_crc == 2: length CRCs only (the scan needs them to
      return fail(CTR_FEED_ERR_CRC, "corrupted record data at byte %llu (crc mismatch)", (unsigned long long)pos);   // trust `len`);
                                                                               // the payload CRCs go to ctr_feed_tfrecord_verify
    if (max_records > 0) { offsets[count] = pos + 12; lengths[count] = len; }
    ++count;
    pos += 12 + len + 4;
  }
  if (consumed) *consumed = pos;
  return count;
}

int64_t ctr_feed_tfrecord_index(const uint8_t* buf, uint64_t n, int verify_crc, uint64_t* offsets, uint64_t* lengths,
                                int64_t max_records, uint64_t* consumed) {
  return ctr_feed_tfrecord_index_from(buf, n, 0, verify_crc, 0, offsets, lengths, max_records, consumed);
}

static int ctr_feed_tfrecord_verify_impl(const uint8_t* buf, uint64_t n, const uint64_t* offsets, const uint64_t* lengths, int64_t count,
                             int num_threads) {
  if (count < 0 || (count > 0 && (!buf || !offsets || !lengths)))
    return fail(CTR_FEED_ERR_ARG, "ctr_feed_tfrecord_verify: bad arguments");
  if (count == 0) return CTR_FEED_OK;
  int nt = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min({nt, 64, (int)std::max<int64_t>(1, count / 256)}));
  // every chunk stops at ITS first bad record; run_chunks reports the chunk with the lowest index first = the first bad record
  return run_chunks(count, nt, [&](int, int64_t b0, int64_t b1) -> int {
    for (int64_t b = b0; b < b1; ++b) {
      const uint64_t off = offsets[b], len = lengths[b];
      if (off < 12 || off > n || len > n - off || n - off - len < 4)
        return fail(CTR_FEED_ERR_ARG, "ctr_feed_tfrecord_verify: record %lld lies outside the buffer", (long long)b);
      uint32_t data_crc;
      memcpy(&data_crc, buf + off + len, 4);
      if (masked(crc32c(buf + off, len)) != data_crc)
        return fail(CTR_FEED_ERR_CRC, "corrupted record data at byte %llu (crc mismatch)", (unsigned long long)(off - 12));
    }
    return CTR_FEED_OK;
  });
}

static int ctr_feed_shuffle_order_impl(int64_t n, int64_t buffer_size, const double* draws, int64_t* out) {
  if (n < 0 || (n > 0 && !out)) return fail(CTR_FEED_ERR_ARG, "ctr_feed_shuffle_order: bad arguments");
  if (buffer_size <= 1 || n <= 1) {
    for (int64_t i = 0; i < n; ++i) out[i] = i;
    return CTR_FEED_OK;
  }
  if (!draws) return fail(CTR_FEED_ERR_ARG, "ctr_feed_shuffle_order: draws is null");
  ShuffleState st(std::min(buffer_size, n));
  const int64_t got = st.emit(n, true, draws, n, out);
  return got == n ? CTR_FEED_OK : (int)got;               // a negative `got` is the error code (draw outside [0,1))
}

void* ctr_feed_shuffle_create(int64_t buffer_size) {
  void* h = nullptr;
  guarded([&] { h = new ShuffleState(std::max<int64_t>(1, buffer_size)); return CTR_FEED_OK; });
  return h;
}
void ctr_feed_shuffle_destroy(void* shuffle) { delete static_cast<ShuffleState*>(shuffle); }
static int64_t ctr_feed_shuffle_emit_impl(void* shuffle, int64_t n_available, int input_done, const double* draws, int64_t max_out,
                              int64_t* out) {
  if (!shuffle || n_available < 0 || max_out < 0 || (max_out > 0 && (!out || !draws)))
    return fail(CTR_FEED_ERR_ARG, "ctr_feed_shuffle_emit: bad arguments");
  return static_cast<ShuffleState*>(shuffle)->emit(n_available, input_done != 0, draws, max_out, out);
}

static void* ctr_feed_vocab_create_impl(const uint8_t* blob, const uint64_t* offsets, int64_t n_tokens) {
  if (n_tokens < 0 || (n_tokens > 0 && (!offsets || (!blob && offsets[n_tokens] > 0)))) {
    fail(CTR_FEED_ERR_ARG, "ctr_feed_vocab_create: bad arguments");
    return nullptr;
  }
  std::unique_ptr<Vocab> v(new Vocab());                   // owned until it is handed to the caller: an exception frees it
  std::vector<std::pair<uint64_t, uint64_t>> spans((size_t)n_tokens);
  if (n_tokens > 0) {
    v->blob.assign(reinterpret_cast<const char*>(blob) + offsets[0], offsets[n_tokens] - offsets[0]);
    for (int64_t i = 0; i < n_tokens; ++i) spans[i] = {offsets[i] - offsets[0], offsets[i + 1] - offsets[i]};
  }
  v->build(spans);
  return v.release();
}

static void* ctr_feed_vocab_load_impl(const char* path) {
  if (!path) { fail(CTR_FEED_ERR_ARG, "ctr_feed_vocab_load: null path"); return nullptr; }
  std::unique_ptr<FILE, int (*)(FILE*)> f(fopen(path, "rb"), fclose);
  if (!f) { fail(CTR_FEED_ERR_IO, "ctr_feed_vocab_load: cannot open %s", path); return nullptr; }
  std::unique_ptr<Vocab> v(new Vocab());
  char chunk[1 << 16];
  size_t got;
  while ((got = fread(chunk, 1, sizeof(chunk), f.get())) > 0) v->blob.append(chunk, got);
  f.reset();
  std::vector<std::pair<uint64_t, uint64_t>> spans;
  uint64_t start = 0;
  const uint64_t n = v->blob.size();
  for (uint64_t i = 0; i <= n; ++i) {
    if (i == n || v->blob[i] == '\n') {
      if (i == n && start == n) break;                      // no trailing empty line after the final newline
      uint64_t end = i;
      while (end > start && (v->blob[end - 1] == '\r' || v->blob[end - 1] == '\n')) --end;
      spans.emplace_back(start, end - start);
      start = i + 1;
    }
  }
  v->build(spans);
  return v.release();
}

int64_t ctr_feed_vocab_size(const void* vocab) { return vocab ? static_cast<const Vocab*>(vocab)->size : -1; }
void ctr_feed_vocab_destroy(void* vocab) { delete static_cast<Vocab*>(vocab); }

int ctr_feed_vocab_lookup(const void* vocab, const uint8_t* blob, const uint64_t* offsets, int64_t n_keys, int64_t* ids_out) {
  if (!vocab || n_keys < 0 || (n_keys > 0 && (!offsets || !ids_out))) return fail(CTR_FEED_ERR_ARG, "ctr_feed_vocab_lookup: bad arguments");
  const Vocab* v = static_cast<const Vocab*>(vocab);
  for (int64_t i = 0; i < n_keys; ++i) ids_out[i] = v->find(blob + offsets[i], offsets[i + 1] - offsets[i]);
  return CTR_FEED_OK;
}

static int ctr_feed_parse_examples_impl(const uint8_t* buf, const uint64_t* offsets, const uint64_t* lengths, int64_t B, ctr_feed_cat_t* cats,
                            int64_t n_cat, ctr_feed_dense_t* dense, int64_t n_dense, int read_feature_lists, int num_threads) {
  if (B < 0 || n_cat < 0 || n_dense < 0 || (B > 0 && (!buf || !offsets || !lengths)) || (n_cat > 0 && !cats) || (n_dense > 0 && !dense))
    return fail(CTR_FEED_ERR_ARG, "ctr_feed_parse_examples: bad arguments");
  std::vector<KeySpec> keys;
  for (int64_t k = 0; k < n_cat; ++k) {
    if (!cats[k].key || !cats[k].vocab || !cats[k].row_offsets || cats[k].capacity < 0 || (cats[k].capacity > 0 && !cats[k].ids))
      return fail(CTR_FEED_ERR_ARG, "ctr_feed_parse_examples: categorical spec %lld is incomplete", (long long)k);
    keys.push_back({std::string_view(cats[k].key), true, (int)k});
  }
  for (int64_t k = 0; k < n_dense; ++k) {
    if (!dense[k].key || dense[k].width < 1 || (B > 0 && !dense[k].out))
      return fail(CTR_FEED_ERR_ARG, "ctr_feed_parse_examples: dense spec %lld is incomplete", (long long)k);
    keys.push_back({std::string_view(dense[k].key), false, (int)k});
  }
  for (size_t a = 0; a < keys.size(); ++a)
    for (size_t b = a + 1; b < keys.size(); ++b)
      if (keys[a].key == keys[b].key)
        return fail(CTR_FEED_ERR_ARG, "ctr_feed_parse_examples: key %s appears twice in the spec", std::string(keys[a].key).c_str());
  int nt = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  nt = std::max(1, std::min({nt, 32, (int)std::max<int64_t>(1, B / 64)}));
  const Parser parser(keys, read_feature_lists != 0);

  // One parse per record.  Every thread owns a contiguous run of records: it writes the dense outputs and the per-record value
  // counts in place and collects the vocabulary ids of its run per categorical key; after the prefix sum over the counts a run's
  // ids are one contiguous range of the output and are copied there (8 bytes per value -- the wire parse is not repeated).
  std::vector<std::vector<std::vector<int64_t>>> local((size_t)nt, std::vector<std::vector<int64_t>>((size_t)n_cat));
  return run_two_phase(
      B, nt,
      [&](int t, int64_t b0, int64_t b1) -> int {
        RecordView rv;
        for (int64_t k = 0; k < n_cat; ++k) local[t][k].reserve((size_t)(b1 - b0) + 16);
        for (int64_t b = b0; b < b1; ++b) {
          if (!parser.view(Span{buf + offsets[b], lengths[b]}, rv)) return fail(CTR_FEED_ERR_PROTO, "Could not parse example input, record %lld", (long long)b);
          for (size_t i = 0; i < keys.size(); ++i) {
            if (keys[i].is_cat) {
              const ctr_feed_cat_t& c = cats[keys[i].index];
              const Vocab* v = static_cast<const Vocab*>(c.vocab);
              std::vector<int64_t>& dst = local[t][keys[i].index];
              const size_t before = dst.size();
              int r = cat_values(rv, (int)i, read_feature_lists != 0, c.key, [&](const uint8_t* p, uint64_t n) { dst.push_back(v->find(p, n)); });
              if (r) return r;
              c.row_offsets[b + 1] = (int64_t)(dst.size() - before);
            } else {
              int r = dense_values(rv, (int)i, dense[keys[i].index], b);
              if (r) return r;
            }
          }
        }
        return CTR_FEED_OK;
      },
      [&]() -> int {                                                   // counts -> offsets; are the callers' buffers large enough?
        bool short_buf = false;
        for (int64_t k = 0; k < n_cat; ++k) {
          int64_t* ro = cats[k].row_offsets;
          ro[0] = 0;
          for (int64_t b = 0; b < B; ++b) ro[b + 1] += ro[b];
          cats[k].needed = ro[B];
          if (ro[B] > cats[k].capacity) short_buf = true;
        }
        if (short_buf) return fail(CTR_FEED_ERR_CAPACITY, "ctr_feed_parse_examples: an ids buffer is too small (see `needed`)");
        return CTR_FEED_OK;
      },
      [&](int t, int64_t b0, int64_t) -> int {
        for (int64_t k = 0; k < n_cat; ++k)
          if (!local[t][k].empty()) memcpy(cats[k].ids + cats[k].row_offsets[b0], local[t][k].data(), local[t][k].size() * sizeof(int64_t));
        return CTR_FEED_OK;
      });
}

// ---- the exported entry points of everything above that allocates: exceptions become CTR_FEED_ERR_NOMEM / a null handle
int ctr_feed_tfrecord_verify(const uint8_t* buf, uint64_t n, const uint64_t* offsets, const uint64_t* lengths, int64_t count,
                             int num_threads) {
  return guarded([&] { return ctr_feed_tfrecord_verify_impl(buf, n, offsets, lengths, count, num_threads); });
}

int ctr_feed_shuffle_order(int64_t n, int64_t buffer_size, const double* draws, int64_t* out) {
  return guarded([&] { return ctr_feed_shuffle_order_impl(n, buffer_size, draws, out); });
}

int64_t ctr_feed_shuffle_emit(void* shuffle, int64_t n_available, int input_done, const double* draws, int64_t max_out,
                              int64_t* out) {
  int64_t r = 0;
  const int rc = guarded([&] { r = ctr_feed_shuffle_emit_impl(shuffle, n_available, input_done, draws, max_out, out); return CTR_FEED_OK; });
  return rc != CTR_FEED_OK ? rc : r;
}

void* ctr_feed_vocab_create(const uint8_t* blob, const uint64_t* offsets, int64_t n_tokens) {
  void* h = nullptr;
  guarded([&] { h = ctr_feed_vocab_create_impl(blob, offsets, n_tokens); return CTR_FEED_OK; });
  return h;
}

void* ctr_feed_vocab_load(const char* path) {
  void* h = nullptr;
  guarded([&] { h = ctr_feed_vocab_load_impl(path); return CTR_FEED_OK; });
  return h;
}

int ctr_feed_parse_examples(const uint8_t* buf, const uint64_t* offsets, const uint64_t* lengths, int64_t B, ctr_feed_cat_t* cats,
                            int64_t n_cat, ctr_feed_dense_t* dense, int64_t n_dense, int read_feature_lists, int num_threads) {
  return guarded([&] { return ctr_feed_parse_examples_impl(buf, offsets, lengths, B, cats, n_cat, dense, n_dense, read_feature_lists, num_threads); });
}

}  // extern "C"
