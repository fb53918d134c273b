// TFRecord files of tokenised windows, byte-compatible with what the reference's encoding stage leaves in
// <prefix>_encoded_sequences/ (genomad/modules/nn_classification.py:43-52: one tf.train.Example per window,
// features {"sequence": Int64List(5997 tokens)}, written with tf.io.TFRecordWriter) and what its
// parse_tfrecord (:87-91, FixedLenFeature([5997], int64)) reads back.  SURVEY.md §8f rank 4: nothing downstream reads
// these files; they exist for directory-level compatibility with the reference and are off by default.
//
// The two formats involved are public and small, so both are written out by hand (no protobuf / TensorFlow dependency):
//
//   TFRecord framing (tensorflow/core/lib/io/record_writer.cc):
//       uint64 length (LE) | uint32 masked_crc32c(length bytes) | data[length] | uint32 masked_crc32c(data)
//       masked(c) = ((c >> 15) | (c << 17)) + 0xa282ead8;  crc32c = CRC-32C (Castagnoli, RFC 3720 appendix B.4)
//   Example protobuf (tensorflow/core/example/{example,feature}.proto), wire format:
//       Example   { Features features = 1; }                    0A <len>
//       Features  { map<string, Feature> feature = 1; }         0A <len>   (one map entry message:)
//                   entry { string key = 1; Feature value = 2 } 0A 08 "sequence"  12 <len>
//       Feature   { oneof { ... Int64List int64_list = 3; } }   1A <len>
//       Int64List { repeated int64 value = 1 [packed = true]; } 0A <len> varint*
//   Tokens are 0..256, so a token is one varint byte (< 128) or two.
//
// Records are serialised by worker threads into per-thread buffers and written in order; reading verifies both CRCs
// of every record and every length on the way down and fails loudly on anything else than the layout above.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gnm.h"

namespace {

constexpr int kTok = 5997;

thread_local std::string g_err;

uint32_t g_crc_table[8][256];
bool g_crc_ready = [] {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0x82F63B78u : 0u);     // reflected Castagnoli polynomial
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xff];
  return true;
}();

uint32_t crc32c(const uint8_t* p, size_t n) {
  uint32_t c = 0xffffffffu;
  while (n >= 8) {                                    // slicing-by-8
    uint64_t v;
    std::memcpy(&v, p, 8);
    v ^= c;
    c = g_crc_table[7][v & 0xff] ^ g_crc_table[6][(v >> 8) & 0xff] ^ g_crc_table[5][(v >> 16) & 0xff] ^
        g_crc_table[4][(v >> 24) & 0xff] ^ g_crc_table[3][(v >> 32) & 0xff] ^ g_crc_table[2][(v >> 40) & 0xff] ^
        g_crc_table[1][(v >> 48) & 0xff] ^ g_crc_table[0][(v >> 56) & 0xff];
    p += 8; n -= 8;
  }
  while (n--) c = (c >> 8) ^ g_crc_table[0][(c ^ *p++) & 0xff];
  return ~c;
}

inline uint32_t masked(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }

inline int varint_len(uint64_t v) { int n = 1; while (v >= 128) { v >>= 7; ++n; } return n; }
inline uint8_t* put_varint(uint8_t* p, uint64_t v) {
  while (v >= 128) { *p++ = static_cast<uint8_t>(v) | 0x80; v >>= 7; }
  *p++ = static_cast<uint8_t>(v);
  return p;
}
inline void put_le32(uint8_t* p, uint32_t v) { for (int i = 0; i < 4; ++i) p[i] = static_cast<uint8_t>(v >> (8 * i)); }
inline void put_le64(uint8_t* p, uint64_t v) { for (int i = 0; i < 8; ++i) p[i] = static_cast<uint8_t>(v >> (8 * i)); }
inline uint32_t get_le32(const uint8_t* p) { uint32_t v = 0; for (int i = 0; i < 4; ++i) v |= static_cast<uint32_t>(p[i]) << (8 * i); return v; }
inline uint64_t get_le64(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 8; ++i) v |= static_cast<uint64_t>(p[i]) << (8 * i); return v; }

// one framed record for one window, appended to `out`
void append_record(const uint16_t* tok, std::vector<uint8_t>& out) {
  size_t packed = 0;
  for (int i = 0; i < kTok; ++i) packed += tok[i] < 128 ? 1 : (tok[i] < 16384 ? 2 : 3);
  const size_t list_len = 1 + varint_len(packed) + packed;             // 0A <len> varints
  const size_t feat_len = 1 + varint_len(list_len) + list_len;         // 1A <len> Int64List
  const size_t entry_len = 2 + 8 + 1 + varint_len(feat_len) + feat_len;  // 0A 08 "sequence" 12 <len> Feature
  const size_t feats_len = 1 + varint_len(entry_len) + entry_len;      // 0A <len> entry
  const size_t ex_len = 1 + varint_len(feats_len) + feats_len;         // 0A <len> Features
  const size_t at = out.size();
  out.resize(at + 12 + ex_len + 4);
  uint8_t* p = out.data() + at;
  put_le64(p, ex_len);
  put_le32(p + 8, masked(crc32c(p, 8)));
  uint8_t* d = p + 12;
  uint8_t* q = d;
  *q++ = 0x0A; q = put_varint(q, feats_len);
  *q++ = 0x0A; q = put_varint(q, entry_len);
  *q++ = 0x0A; *q++ = 8; std::memcpy(q, "sequence", 8); q += 8;
  *q++ = 0x12; q = put_varint(q, feat_len);
  *q++ = 0x1A; q = put_varint(q, list_len);
  *q++ = 0x0A; q = put_varint(q, packed);
  for (int i = 0; i < kTok; ++i) q = put_varint(q, tok[i]);
  put_le32(q, masked(crc32c(d, ex_len)));
}

bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t& v) {
  v = 0;
  for (int shift = 0; shift < 64 && p < end; shift += 7) {
    const uint8_t b = *p++;
    v |= static_cast<uint64_t>(b & 0x7f) << shift;
    if (!(b & 0x80)) return true;
  }
  return false;
}

// expects tag byte `tag`, then a length that must run exactly to `end`
bool open_field(const uint8_t*& p, const uint8_t* end, uint8_t tag) {
  uint64_t len;
  if (p >= end || *p++ != tag || !get_varint(p, end, len)) return false;
  return static_cast<uint64_t>(end - p) == len;
}

bool parse_example(const uint8_t* p, const uint8_t* end, uint16_t* tok) {
  if (!open_field(p, end, 0x0A) || !open_field(p, end, 0x0A)) return false;          // Example.features, Features.feature entry
  if (end - p < 10 || p[0] != 0x0A || p[1] != 8 || std::memcmp(p + 2, "sequence", 8) != 0) return false;
  p += 10;
  if (!open_field(p, end, 0x12) || !open_field(p, end, 0x1A) || !open_field(p, end, 0x0A)) return false;
  for (int i = 0; i < kTok; ++i) {
    uint64_t v;
    if (!get_varint(p, end, v) || v > 0xffff) return false;
    tok[i] = static_cast<uint16_t>(v);
  }
  return p == end;
}

int fail(const std::string& msg) { g_err = msg; return 1; }

}  // namespace

extern "C" {

const char* gnm_tfrecord_last_error(void) { return g_err.c_str(); }

uint32_t gnm_crc32c(const void* data, size_t n) { return crc32c(static_cast<const uint8_t*>(data), n); }

int gnm_tfrecord_write(const char* path, const uint16_t* tokens, int64_t n, int threads) {
  if (!path || (n > 0 && !tokens) || n < 0) return fail("gnm_tfrecord_write: bad arguments");
  std::FILE* f = std::fopen(path, "wb");
  if (!f) return fail(std::string("gnm_tfrecord_write: cannot open ") + path);
  const int64_t kBlock = 256;                                   // windows per work item (~1.7 MB of output)
  const int64_t n_blocks = (n + kBlock - 1) / kBlock;
  const int T = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(threads > 0 ? threads : 1, n_blocks)));
  bool ok = true;
  // rounds of T blocks: serialise in parallel, then write in order
  std::vector<std::vector<uint8_t>> bufs(T);
  for (int64_t b0 = 0; b0 < n_blocks && ok; b0 += T) {
    const int live = static_cast<int>(std::min<int64_t>(T, n_blocks - b0));
    std::vector<std::thread> pool;
    for (int t = 0; t < live; ++t)
      pool.emplace_back([&, t] {
        auto& buf = bufs[t];
        buf.clear();
        const int64_t lo = (b0 + t) * kBlock, hi = std::min(n, lo + kBlock);
        buf.reserve(static_cast<size_t>(hi - lo) * 7100);
        for (int64_t w = lo; w < hi; ++w) append_record(tokens + w * kTok, buf);
      });
    for (auto& th : pool) th.join();
    for (int t = 0; t < live && ok; ++t) ok = std::fwrite(bufs[t].data(), 1, bufs[t].size(), f) == bufs[t].size();
  }
  if (std::fclose(f) != 0) ok = false;
  return ok ? 0 : fail(std::string("gnm_tfrecord_write: short write to ") + path);
}

int gnm_tfrecord_read(const char* path, uint16_t* tokens, int64_t capacity, int64_t* n_records) {
  if (!path || !n_records || capacity < 0 || (capacity > 0 && !tokens)) return fail("gnm_tfrecord_read: bad arguments");
  std::FILE* f = std::fopen(path, "rb");
  if (!f) return fail(std::string("gnm_tfrecord_read: cannot open ") + path);
  std::vector<uint8_t> data;
  uint8_t head[12], tail[4];
  int64_t n = 0;
  int rc = 0;
  for (;;) {
    const size_t got = std::fread(head, 1, 12, f);
    if (got == 0) break;
    if (got != 12) { rc = fail("gnm_tfrecord_read: truncated record header"); break; }
    if (get_le32(head + 8) != masked(crc32c(head, 8))) { rc = fail("gnm_tfrecord_read: length CRC mismatch"); break; }
    const uint64_t len = get_le64(head);
    if (len > (1u << 20)) { rc = fail("gnm_tfrecord_read: implausible record length"); break; }
    data.resize(len);
    if (std::fread(data.data(), 1, len, f) != len || std::fread(tail, 1, 4, f) != 4) {
      rc = fail("gnm_tfrecord_read: truncated record"); break;
    }
    if (get_le32(tail) != masked(crc32c(data.data(), len))) { rc = fail("gnm_tfrecord_read: data CRC mismatch"); break; }
    if (tokens) {                                        // tokens == nullptr: count (and verify) only
      if (n >= capacity) { rc = fail("gnm_tfrecord_read: more records than capacity"); break; }
      if (!parse_example(data.data(), data.data() + len, tokens + n * kTok)) {
        rc = fail("gnm_tfrecord_read: record is not Example{sequence: Int64List[5997]}"); break;
      }
    }
    ++n;
  }
  std::fclose(f);
  *n_records = n;
  return rc;
}

}  // extern "C"
