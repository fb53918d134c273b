// Native FASTA -> window-matrix front end of the nn-classification path (host code, part of libgnm.so).
//
// Behavioural mirror (not a translation) of what the reference does per record in Python:
//   read_fasta(strip_n=True)   genomad/sequence.py:96-121   text mode => universal newlines: "\r\n" and "\r" end a
//                              line like "\n"; text before the first line starting with '>' is ignored; the record's
//                              sequence is every byte of its lines except the line terminators; leading/trailing
//                              'n'/'N' of the whole contig are stripped; records that are then empty are dropped
//   seq_windows(6000, 2500)    genomad/sequence.py:150-167   consecutive 6000-nt slices; a shorter last slice is kept
//                              only if >= 2500 nt, except that the first window is always kept; max_windows=1 for
//                              --single-window
//   N rule / pad / upper-case  genomad/modules/nn_classification.py:70-72   a window other than the first is skipped if
//                              its RAW text holds more than 4000 upper-case 'N'; windows are upper-cased (ASCII) and
//                              right-padded with 'N' to 6000 bytes
// Golden vectors produced with the real reference code pin all of this (tests/golden/encoder_golden.json,
// tests/test_host_cpu.py::test_native_fasta_*).  Output: one dense uint8 matrix [n_windows][6000] written straight into
// caller memory (pinned, ready for the H2D copy), per-contig window offsets, and the raw header lines (the accession
// = first whitespace-delimited token is taken by the Python caller with str.split(), exactly as the reference does).
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "../../include/gnm.h"

namespace {

constexpr int64_t kWin = GNM_WINDOW, kMinTail = 2500, kMaxN = 4000;

struct Record {
  int64_t hdr_begin, hdr_end;      // header text (without '>' and terminator) in the input buffer
  int64_t body_begin, body_end;    // raw body bytes (with terminators) in the input buffer
  int64_t seq_off, seq_len;        // stripped, compacted sequence inside gnm_fasta::comp
  int64_t raw_len;                 // compacted length before stripping
  int64_t n_windows, first_window;
};

}  // namespace

struct gnm_fasta {
  const uint8_t* text = nullptr;
  int64_t len = 0;
  int single_window = 0;
  std::vector<Record> recs;        // every record found (before dropping empties)
  std::vector<int64_t> kept;       // indices of records whose stripped sequence is non-empty
  std::unique_ptr<uint8_t[]> comp; // compacted sequences (uninitialised), record i inside its own raw body span
  std::vector<std::vector<int32_t>> win_start;   // per kept record: start (in nt) of each kept window
  std::vector<int32_t> flat_rec, flat_start;     // per window (global order): kept-record index, start in nt
  int64_t n_windows = 0;
  int64_t n_nonempty_raw = 0;      // records with a non-empty sequence before stripping (what check_fasta counts)
  int has_dup = 0;
};

static thread_local std::string g_fasta_err;
extern "C" const char* gnm_fasta_last_error(void) { return g_fasta_err.c_str(); }

static inline bool is_eol(uint8_t c) { return c == '\n' || c == '\r'; }

template <class F>
static void parallel_for(int64_t n, int threads, F fn) {
  threads = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(threads, n)));
  if (threads == 1) { for (int64_t i = 0; i < n; ++i) fn(i); return; }
  std::atomic<int64_t> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] { for (int64_t i; (i = next.fetch_add(1)) < n;) fn(i); });
  for (auto& th : pool) th.join();
}

extern "C" int gnm_fasta_parse(const uint8_t* text, size_t len_, int single_window, int threads, gnm_fasta** out) {
  if (!out || (!text && len_)) { g_fasta_err = "gnm_fasta_parse: null argument"; return 1; }
  gnm_fasta* f = new gnm_fasta();
  f->text = text; f->len = static_cast<int64_t>(len_); f->single_window = single_window;
  const int64_t len = f->len;
  // ---- pass 1 (serial, memchr speed): header lines = '>' at offset 0 or right after a line terminator
  int64_t pos = 0;
  std::vector<int64_t> starts;
  while (pos < len) {
    const void* p = std::memchr(text + pos, '>', static_cast<size_t>(len - pos));
    if (!p) break;
    const int64_t i = static_cast<const uint8_t*>(p) - text;
    if (i == 0 || is_eol(text[i - 1])) starts.push_back(i);
    pos = i + 1;
  }
  f->recs.resize(starts.size());
  for (size_t r = 0; r < starts.size(); ++r) {
    Record& R = f->recs[r];
    const int64_t rec_end = r + 1 < starts.size() ? starts[r + 1] : len;
    int64_t e = starts[r] + 1;
    while (e < rec_end && !is_eol(text[e])) ++e;
    R.hdr_begin = starts[r] + 1; R.hdr_end = e;
    if (e < rec_end && text[e] == '\r' && e + 1 < rec_end && text[e + 1] == '\n') ++e;   // "\r\n" is one terminator
    R.body_begin = std::min(e + 1, rec_end); R.body_end = rec_end;
    R.seq_off = R.seq_len = R.raw_len = R.n_windows = R.first_window = 0;
  }
  // ---- pass 2 (parallel per record): compact (drop terminators), strip n/N, enumerate windows
  f->comp.reset(new uint8_t[static_cast<size_t>(len) + 1]);
  std::vector<std::vector<int32_t>> wins(f->recs.size());
  parallel_for(static_cast<int64_t>(f->recs.size()), threads, [&](int64_t r) {
    Record& R = f->recs[r];
    uint8_t* dst = f->comp.get() + R.body_begin;
    int64_t n = 0, i = R.body_begin;
    while (i < R.body_end) {                             // copy line by line, skipping '\n' / '\r'
      int64_t j = i;
      while (j < R.body_end && !is_eol(text[j])) ++j;
      std::memcpy(dst + n, text + i, static_cast<size_t>(j - i));
      n += j - i;
      i = j + 1;
    }
    R.raw_len = n;
    int64_t b = 0, e = n;
    while (b < e && (dst[b] == 'n' || dst[b] == 'N')) ++b;
    while (e > b && (dst[e - 1] == 'n' || dst[e - 1] == 'N')) --e;
    R.seq_off = R.body_begin + b; R.seq_len = e - b;
    const uint8_t* s = f->comp.get() + R.seq_off;
    for (int64_t w = 0; w * kWin < R.seq_len; ++w) {
      const int64_t ws = w * kWin, we = std::min(ws + kWin, R.seq_len);
      if (we - ws < kMinTail) {
        if (w == 0) wins[r].push_back(static_cast<int32_t>(ws));
        break;
      }
      bool keep = true;
      if (w > 0) {
        int64_t nn = 0;
        for (int64_t k = ws; k < we; ++k) nn += (s[k] == 'N');
        keep = nn <= kMaxN;
      }
      if (keep) wins[r].push_back(static_cast<int32_t>(ws));
      if (single_window) break;
    }
    R.n_windows = static_cast<int64_t>(wins[r].size());
  });
  // ---- bookkeeping: kept records, window offsets, duplicate identifiers (first whitespace-delimited token)
  std::unordered_set<std::string> ids;
  for (size_t r = 0; r < f->recs.size(); ++r) {
    Record& R = f->recs[r];
    if (R.raw_len > 0) {
      ++f->n_nonempty_raw;
      int64_t a = R.hdr_begin, e = R.hdr_end;
      auto ws = [](uint8_t c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 28 && c <= 31); };
      while (a < e && ws(text[a])) ++a;
      int64_t b = a;
      while (b < e && !ws(text[b])) ++b;
      if (!ids.emplace(reinterpret_cast<const char*>(text + a), static_cast<size_t>(b - a)).second) f->has_dup = 1;
    }
    if (R.seq_len > 0) {
      R.first_window = f->n_windows;
      f->n_windows += R.n_windows;
      for (int32_t st : wins[r]) { f->flat_rec.push_back(static_cast<int32_t>(f->kept.size())); f->flat_start.push_back(st); }
      f->kept.push_back(static_cast<int64_t>(r));
      f->win_start.push_back(std::move(wins[r]));
    }
  }
  *out = f;
  return 0;
}

extern "C" int gnm_fasta_info(const gnm_fasta* f, int64_t* n_records_nonempty, int* has_duplicate_ids, int64_t* n_contigs,
                              int64_t* n_windows, int64_t* header_bytes) {
  if (!f) { g_fasta_err = "gnm_fasta_info: null handle"; return 1; }
  if (n_records_nonempty) *n_records_nonempty = f->n_nonempty_raw;
  if (has_duplicate_ids) *has_duplicate_ids = f->has_dup;
  if (n_contigs) *n_contigs = static_cast<int64_t>(f->kept.size());
  if (n_windows) *n_windows = f->n_windows;
  if (header_bytes) {
    int64_t b = 0;
    for (int64_t r : f->kept) b += f->recs[r].hdr_end - f->recs[r].hdr_begin + 1;
    *header_bytes = b;
  }
  return 0;
}

extern "C" int gnm_fasta_export(const gnm_fasta* f, uint8_t* windows, int32_t* offsets, char* headers, int threads) {
  if (!f) { g_fasta_err = "gnm_fasta_export: null handle"; return 1; }
  if (f->n_windows > INT32_MAX) { g_fasta_err = "gnm_fasta_export: more than 2^31-1 windows"; return 1; }
  const int64_t nk = static_cast<int64_t>(f->kept.size());
  if (offsets) {
    for (int64_t i = 0; i < nk; ++i) offsets[i] = static_cast<int32_t>(f->recs[f->kept[i]].first_window);
    offsets[nk] = static_cast<int32_t>(f->n_windows);
  }
  if (headers) {
    char* h = headers;
    for (int64_t i = 0; i < nk; ++i) {
      const Record& R = f->recs[f->kept[i]];
      const int64_t n = R.hdr_end - R.hdr_begin;
      std::memcpy(h, f->text + R.hdr_begin, static_cast<size_t>(n));
      h[n] = '\n';
      h += n + 1;
    }
  }
  if (windows) {
    parallel_for(nk, threads, [&](int64_t i) {
      const Record& R = f->recs[f->kept[i]];
      const uint8_t* s = f->comp.get() + R.seq_off;
      const std::vector<int32_t>& ws = f->win_start[i];
      for (size_t k = 0; k < ws.size(); ++k) {
        uint8_t* dst = windows + (R.first_window + static_cast<int64_t>(k)) * kWin;
        const int64_t b = ws[k], n = std::min<int64_t>(kWin, R.seq_len - b);
        for (int64_t j = 0; j < n; ++j) {
          const uint8_t c = s[b + j];
          dst[j] = (c >= 'a' && c <= 'z') ? static_cast<uint8_t>(c - 32) : c;     // ASCII upper(), as bytes.upper()
        }
        if (n < kWin) std::memset(dst + n, 'N', static_cast<size_t>(kWin - n));
      }
    });
  }
  return 0;
}

// windows [first, first + count) of the global window list -> dst [count][6000]  (streaming export: the driver fills one
// pinned chunk while the GPU classifies the previous one)
extern "C" int gnm_fasta_export_windows(const gnm_fasta* f, int64_t first, int64_t count, uint8_t* dst, int threads) {
  if (!f || !dst) { g_fasta_err = "gnm_fasta_export_windows: null argument"; return 1; }
  if (first < 0 || count < 0 || first + count > f->n_windows) { g_fasta_err = "gnm_fasta_export_windows: range out of bounds"; return 1; }
  constexpr int64_t kBlock = 64;                            // windows per work item
  parallel_for((count + kBlock - 1) / kBlock, threads, [&](int64_t b) {
    const int64_t lo = first + b * kBlock, hi = std::min(first + count, lo + kBlock);
    for (int64_t wdx = lo; wdx < hi; ++wdx) {
      const Record& R = f->recs[f->kept[f->flat_rec[wdx]]];
      const uint8_t* s = f->comp.get() + R.seq_off;
      uint8_t* out = dst + (wdx - first) * kWin;
      const int64_t st = f->flat_start[wdx], n = std::min<int64_t>(kWin, R.seq_len - st);
      for (int64_t j = 0; j < n; ++j) {
        const uint8_t c = s[st + j];
        out[j] = (c >= 'a' && c <= 'z') ? static_cast<uint8_t>(c - 32) : c;
      }
      if (n < kWin) std::memset(out + n, 'N', static_cast<size_t>(kWin - n));
    }
  });
  return 0;
}

extern "C" void gnm_fasta_free(gnm_fasta* f) { delete f; }
