/*
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): plain-C restatement of the integer part of the hot path, used by tests/ as a
 * second, independent checker next to the NumPy closed form in oracle/tokenizer.py (and by nothing under genomad_b200/).
 * PINNED: tests/test_oracle_golden.py runs it against the golden vectors made by the real reference code
 * (tests/golden/encoder_golden.json, encoder_batch_tokens.npz, reference_module/run_default/encoded_tokens.npz).
 *
 *   gnm_oracle_tokenize          reference genomad/sequence.py:170-193 (tokenize_dna, word size 4), any length, incl. the
 *                                Python-slice behaviour of the final `[:final_length]` for inputs shorter than 4
 *   gnm_oracle_tokenize_windows  the same for a matrix of equal-length windows (what nn_classification.py:72-73 feeds it)
 *   gnm_oracle_window_plan       which windows of one sequence are classified: 6,000-nt windows, a tail shorter than 2,500 is
 *                                dropped unless it is the first window (sequence.py:150-167), windows after the first with more
 *                                than 4,000 'N' are skipped (nn_classification.py:68-71), optional single-window mode
 *
 * Built by __graft_entry__.build():  gcc -O2 -shared -fPIC oracle/tokenizer_c.c -o oracle/_build/liboracle_tok.so
 */
#include <stddef.h>
#include <stdint.h>

static int base_code(uint8_t b) {           /* only upper-case A C G T are bases; everything else breaks the k-mer */
  switch (b) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
  }
}

/* out must hold len + 4 values; returns the number of tokens */
int64_t gnm_oracle_tokenize(const uint8_t* seq, int64_t len, uint16_t* out) {
  int64_t n = 0;
  unsigned kmer = 0;
  int missing = 3;                          /* bases still to come before the running 4-mer is complete */
  for (int64_t i = 0; i < len; ++i) {
    const int c = base_code(seq[i]);
    if (c >= 0) {
      kmer = ((kmer << 2) | (unsigned)c) & 0xFFu;
    } else {
      for (int k = 0; k < 4 - missing; ++k) out[n++] = 0;   /* the windows this byte spoils that were not emitted yet */
      missing = 4;
    }
    if (missing == 0) out[n++] = (uint16_t)(kmer + 1);
    else --missing;
  }
  const int64_t final_length = len - 3;
  if (final_length >= 0) return n < final_length ? n : final_length;
  return n + final_length > 0 ? n + final_length : 0;       /* list[:negative] drops from the end */
}

/* ascii [n][len] -> tokens [n][len - 3]; returns 0, or -(1 + index) of the first window whose token count is not len - 3 */
int64_t gnm_oracle_tokenize_windows(const uint8_t* ascii, int64_t n, int64_t len, uint16_t* tokens, uint16_t* scratch /* len + 4 */) {
  for (int64_t w = 0; w < n; ++w) {
    const int64_t got = gnm_oracle_tokenize(ascii + w * len, len, scratch);
    if (got != len - 3) return -(1 + w);
    for (int64_t i = 0; i < got; ++i) tokens[w * (len - 3) + i] = scratch[i];
  }
  return 0;
}

/* seq: one record's sequence AFTER the reader stripped leading / trailing n/N.  Writes the start offset and length of every
 * window that is classified (capacity cap); returns their number, or -1 if cap is too small. */
int64_t gnm_oracle_window_plan(const uint8_t* seq, int64_t len, int single_window, int64_t* start, int64_t* length, int64_t cap) {
  const int64_t W = 6000, MIN_TAIL = 2500, MAX_N = 4000;
  int64_t kept = 0;
  for (int64_t win = 0; win * W < len; ++win) {
    const int64_t s = win * W, l = (len - s < W) ? len - s : W;
    if (l < MIN_TAIL && win > 0) break;                     /* a short tail is dropped; a short FIRST window is kept */
    int skip = 0;
    if (win > 0) {
      int64_t n_count = 0;
      for (int64_t i = 0; i < l; ++i) n_count += seq[s + i] == 'N';     /* the reference counts upper-case 'N' only */
      skip = n_count > MAX_N;
    }
    if (!skip) {
      if (kept >= cap) return -1;
      start[kept] = s; length[kept] = l; ++kept;
    }
    if (l < MIN_TAIL) break;
    if (single_window) break;                               /* max_windows = 1: only the first window is produced */
  }
  return kept;
}
