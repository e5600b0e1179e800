/*
 * TEST INFRASTRUCTURE ONLY -- the CPU oracle.
 *
 * A plain-C restatement of the reference's sketch-and-compare path
 * (bensonlzl/spaced-kmer-sketching, mounted at /root/reference while developing).  It exists to
 * CHECK the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (spaced_kmer_sketching_b200/, include/) never
 * links, loads or calls anything in oracle/.
 *
 * Pinning: the reference ships no tests or golden vectors of its own.  This file is
 * pinned (tests/test_oracle.py) against (a) the README example (README.md:35-41), (b) the
 * survey-derived known answers KAT-1..4, and (c) fixtures in tests/golden/ produced
 * by running the UNMODIFIED reference sources (oracle/_ref, built against oracle/shim) on seeded
 * inputs -- see tests/golden/make_golden.py.  Everything that does not depend on Boost is therefore
 * pinned by the reference's own code.  The FracMinHash filter depends on boost::hash, which is not
 * vendored and not installed here: both published hash_combine algorithms are restated (variants
 * 171 and 181) and are "pinned to the restatement", not to a real Boost build.
 *
 * Every function cites the reference lines it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

#define ORC_PRED_ALL 0
#define ORC_PRED_FMH 1

/* ------------------------------------------------------------------------------------------
 * 128-bit helpers: a kmer_bitset is always 128 bits (src/kmer.hpp:27,37,53); block 0 = bits 0..63.
 * ---------------------------------------------------------------------------------------- */
static inline u128 mk128(uint64_t lo, uint64_t hi) { return ((u128)hi << 64) | lo; }
static inline uint64_t lo64(u128 v) { return (uint64_t)v; }
static inline uint64_t hi64(u128 v) { return (uint64_t)(v >> 64); }

/* ------------------------------------------------------------------------------------------
 * Boost hash restatement (third-party, un-vendored; call sites src/kmer.hpp:137-148).
 *   hash_value(dynamic_bitset) = hc(hash(num_bits)=128, hash_range(blocks))
 *   hash_range(b0, b1)         = hc(hc(0, b0), b1)
 *   integers hash to themselves.
 * variant 171: Boost 1.71..1.80 64-bit hash_combine_impl; variant 181: Boost >= 1.81 hash_mix.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t hc171(uint64_t h, uint64_t k) {
  const uint64_t m = 0xc6a4a7935bd1e995ULL;
  k *= m;
  k ^= k >> 47;
  k *= m;
  h ^= k;
  h *= m;
  h += 0xe6546b64ULL;
  return h;
}
static inline uint64_t hc181(uint64_t h, uint64_t k) {
  const uint64_t M = 0x0e9846af9b1a615dULL;
  uint64_t x = h + 0x9e3779b9ULL + k;
  x ^= x >> 32;
  x *= M;
  x ^= x >> 32;
  x *= M;
  x ^= x >> 28;
  return x;
}
static inline uint64_t hc(uint64_t h, uint64_t k, int variant) {
  return variant == 171 ? hc171(h, k) : hc181(h, k);
}

uint64_t orc_boost_hash_bitset(uint64_t lo, uint64_t hi, int variant) {
  uint64_t inner = hc(hc(0, lo, variant), hi, variant);
  return hc(128, inner, variant);
}

/* frac_min_hash::operator(), src/kmer.hpp:144-148: H(masked) ^ H(mask) ^ (size_t)window ^ nonce.
 * `nonce` is an int member (src/kmer.hpp:139,141) => sign-extended when XORed into a size_t. */
uint64_t orc_fmh(uint64_t masked_lo, uint64_t masked_hi, uint64_t mask_lo, uint64_t mask_hi, int window,
                 int nonce, int variant) {
  return orc_boost_hash_bitset(masked_lo, masked_hi, variant) ^
         orc_boost_hash_bitset(mask_lo, mask_hi, variant) ^ (uint64_t)(int64_t)window ^
         (uint64_t)(int64_t)nonce;
}

/* sketching_condition, src/kmer-sketching.cpp:30-34: fmh(k) % c == 0 (c promoted to size_t). */
static inline int pred_pass(int pred_kind, u128 masked, u128 mask, int window, int nonce, uint64_t modulus,
                            int variant) {
  if (pred_kind == ORC_PRED_ALL) return 1;
  return orc_fmh(lo64(masked), hi64(masked), lo64(mask), hi64(mask), window, nonce, variant) % modulus == 0;
}

/* ------------------------------------------------------------------------------------------
 * FASTA ingest.  nucleotide_to_bits, src/fasta_processing.cpp:35-69.
 * ---------------------------------------------------------------------------------------- */
static inline uint8_t nucleotide_code(char c) {
  switch (c) {
    case 'a': case 'A': return 0;
    case 'c': case 'C': return 1;
    case 'g': case 'G': return 2;
    case 't': case 'T': return 3;
    default: return 4;
  }
}

/* add_nucleotide_strings, src/fasta_processing.cpp:144-179: split one record at every non-ACGT
 * byte; empty pieces are dropped.  Appends codes to out_codes and lengths to out_seg_len. */
static void split_record(const char *rec, int64_t n, uint8_t *out_codes, int64_t *n_codes,
                         int64_t *out_seg_len, int64_t *n_segs) {
  int64_t cur = 0;
  for (int64_t i = 0; i < n; ++i) {
    uint8_t code = nucleotide_code(rec[i]);
    if (code & 4) {
      if (cur > 0) {
        if (out_seg_len) out_seg_len[*n_segs] = cur;
        ++*n_segs;
      }
      cur = 0;
    } else {
      if (out_codes) out_codes[*n_codes] = code;
      ++*n_codes;
      ++cur;
    }
  }
  if (cur > 0) {
    if (out_seg_len) out_seg_len[*n_segs] = cur;
    ++*n_segs;
  }
}

/* strings_from_fasta + cut_nucleotide_strings, src/fasta_processing.cpp:79-133,190-211, on the
 * bytes of a file already in memory.  std::getline semantics: lines end at '\n' (a '\r' stays in
 * the line), a final unterminated line is still a line, a trailing '\n' adds no empty line.
 *   - a line that is empty or starts with '>' flushes the current record if `name` is non-empty;
 *     a '>' line sets name = rest of line (possibly empty); an empty line keeps `name`  (:98-110)
 *   - other lines are appended to the record only while `name` is non-empty            (:112)
 *   - a sequence line containing ' ' clears both name and content                      (:114-118)
 *   - at EOF the record is flushed if `name` is non-empty                              (:125-130)
 * Pass out_codes/out_seg_len = NULL to size the outputs.  Returns 0. */
int orc_fasta_parse(const char *text, int64_t n, uint8_t *out_codes, int64_t *out_n_codes,
                    int64_t *out_seg_len, int64_t *out_n_segs) {
  char *content = (char *)malloc((size_t)(n > 0 ? n : 1));
  int64_t content_len = 0;
  int name_nonempty = 0;
  int64_t n_codes = 0, n_segs = 0;
  int64_t pos = 0;
  while (pos < n) {
    int64_t eol = pos;
    while (eol < n && text[eol] != '\n') ++eol;
    const char *line = text + pos;
    int64_t len = eol - pos;
    pos = eol + 1; /* if eol == n there was no terminator; loop ends either way */
    if (len == 0 || line[0] == '>') {
      if (name_nonempty) split_record(content, content_len, out_codes, &n_codes, out_seg_len, &n_segs);
      if (len != 0) name_nonempty = (len > 1);
      content_len = 0;
    } else if (name_nonempty) {
      if (memchr(line, ' ', (size_t)len) != NULL) {
        name_nonempty = 0;
        content_len = 0;
      } else {
        memcpy(content + content_len, line, (size_t)len);
        content_len += len;
      }
    }
  }
  if (name_nonempty) split_record(content, content_len, out_codes, &n_codes, out_seg_len, &n_segs);
  free(content);
  *out_n_codes = n_codes;
  *out_n_segs = n_segs;
  return 0;
}

/* add_nucleotide_strings alone, on one raw string (src/fasta_processing.cpp:144-179). */
int orc_split_raw(const char *raw, int64_t n, uint8_t *out_codes, int64_t *out_n_codes, int64_t *out_seg_len,
                  int64_t *out_n_segs) {
  int64_t n_codes = 0, n_segs = 0;
  split_record(raw, n, out_codes, &n_codes, out_seg_len, &n_segs);
  *out_n_codes = n_codes;
  *out_n_segs = n_segs;
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * The hot loop: nucleotide_string_to_kmers, src/kmer_sliding.cpp:112-186, for a list of segments
 * (nucleotide_string_list_to_kmers_by_reference, :199-213).
 *
 * codes: concatenated 1-byte codes (0..3); seg_len[n_segs]: segment lengths in order.
 * For every kept k-mer, in sequence order with duplicates, writes masked_bits (2 words) and, if
 * out_bits != NULL, kmer_bits (2 words, including the "history" above bit 2w for the forward
 * strand, :28).  Returns the number of kept k-mers; writes at most `cap` of them.
 * ---------------------------------------------------------------------------------------- */
int64_t orc_kmers(const uint8_t *codes, const int64_t *seg_len, int64_t n_segs, uint64_t mask_lo,
                  uint64_t mask_hi, int window, int pred_kind, int nonce, uint64_t modulus, int variant,
                  uint64_t *out_masked, uint64_t *out_bits, int64_t cap) {
  const u128 mask = mk128(mask_lo, mask_hi);
  int64_t count = 0;
  const uint8_t *s = codes;
  for (int64_t seg = 0; seg < n_segs; s += seg_len[seg], ++seg) {
    const int64_t n = seg_len[seg];
    if (n < window) continue; /* :121-125 */
    u128 cur = 0, rc = 0;     /* :128 */
    const int top = 2 * window - 2;
    for (int64_t t = 0; t < n; ++t) {
      const uint8_t b = s[t];
      /* update_kmer_window :26-31: shift left by 2 (truncated to 128 bits), code into bits 0..1 */
      cur = (cur << 2) | (u128)b;
      /* update_complement_kmer_window :42-47 with b^3 (:140,151): shift right, write bits 2w-2.. */
      rc = (rc >> 2);
      rc &= ~((u128)3 << top);
      rc |= (u128)(b ^ 3) << top;
      if (t + 1 < window) continue; /* warm-up :134-141 */
      const u128 f = cur & mask, r = rc & mask; /* :159-160, same mask on both strands */
      u128 bits, masked;
      if (f < r) { bits = cur; masked = f; } else { bits = rc; masked = r; } /* :166-175, ties -> rc */
      if (pred_pass(pred_kind, masked, mask, window, nonce, modulus, variant)) { /* :183-184 */
        if (count < cap) {
          if (out_masked) { out_masked[2 * count] = lo64(masked); out_masked[2 * count + 1] = hi64(masked); }
          if (out_bits) { out_bits[2 * count] = lo64(bits); out_bits[2 * count + 1] = hi64(bits); }
        }
        ++count;
      }
    }
  }
  return count;
}

/* ------------------------------------------------------------------------------------------
 * Legacy canonicalisation, src/kmers.cpp:16-35 + reverse_kmer_bitset src/kmer_bitset.cpp:105-119.
 * ---------------------------------------------------------------------------------------- */
/* Log-step reversal of the 64 two-bit groups of a 128-bit value (gap sizes 2,4,...,64). */
static u128 reverse_groups(u128 v) {
  for (int gap = 2; gap < 128; gap *= 2) {
    /* reversing_kmer_array: blocks of `gap` bits alternating 0s,1s starting with 0s at bit 0 */
    u128 odd = 0;
    for (int blk = 0; blk < 128 / gap; ++blk)
      if (blk & 1) odd |= ((gap == 128 ? (u128)0 : (((u128)1 << gap) - 1))) << (blk * gap);
    v = ((v & odd) >> gap) | ((v & ~odd) << gap);
  }
  return v;
}
void orc_reverse_bitset(uint64_t lo, uint64_t hi, uint64_t out[2]) {
  u128 r = reverse_groups(mk128(lo, hi));
  out[0] = lo64(r);
  out[1] = hi64(r);
}
/* canonical_kmer(k) for k = {window, bits, mask, bits & mask}: returns kmer_bits and masked_bits. */
void orc_legacy_canonical(int window, uint64_t bits_lo, uint64_t bits_hi, uint64_t mask_lo, uint64_t mask_hi,
                          uint64_t out_bits[2], uint64_t out_masked[2]) {
  const u128 bits = mk128(bits_lo, bits_hi), mask = mk128(mask_lo, mask_hi);
  const int sh = (64 - window) * 2;
  u128 rc = ~reverse_groups(bits);
  rc = sh >= 128 ? 0 : rc >> sh; /* :18 */
  const u128 f = bits & mask, r = rc & mask;
  const u128 cb = (f < r) ? bits : rc, cm = (f < r) ? f : r; /* :34 */
  out_bits[0] = lo64(cb); out_bits[1] = hi64(cb);
  out_masked[0] = lo64(cm); out_masked[1] = hi64(cm);
}

/* ------------------------------------------------------------------------------------------
 * Masks.  contiguous_kmer, src/kmer_bitset.cpp:28-56 (k > 64 is an error: returns -1).
 * ---------------------------------------------------------------------------------------- */
int orc_contiguous_mask(int k, uint64_t out[2]) {
  if (k > 64) return -1;
  u128 m = (k <= 0) ? 0 : (k == 64 ? ~(u128)0 : (((u128)1 << (2 * k)) - 1));
  out[0] = lo64(m);
  out[1] = hi64(m);
  return 0;
}

/* std::mt19937 (the standard's parameters) */
typedef struct { uint32_t mt[624]; int idx; } mt19937_t;
static void mt_seed(mt19937_t *g, uint32_t seed) {
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}
static uint32_t mt_next(mt19937_t *g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t y = g->mt[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
/* libstdc++ (GCC 13) uniform_int_distribution<unsigned long>{0, range-1} over a 32-bit URBG:
 * Lemire's nearly-divisionless method on 64-bit products (bits/uniform_int_dist.h, _S_nd). */
static uint64_t libstdcxx_uniform_below(mt19937_t *g, uint32_t range) {
  uint64_t product = (uint64_t)mt_next(g) * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    uint32_t threshold = (uint32_t)(-range) % range;
    while (low < threshold) {
      product = (uint64_t)mt_next(g) * (uint64_t)range;
      low = (uint32_t)product;
    }
  }
  return product >> 32;
}
/* generate_random_spaced_seed_mask, src/kmer_bitset.cpp:132-152: iota(0..window-1), libstdc++
 * std::shuffle with std::mt19937(seed) (two swap positions per draw, bits/stl_algo.h), first
 * kmer_size shuffled positions get both of their bits set. */
void orc_random_mask(int window, int kmer_size, uint64_t seed, uint64_t out[2]) {
  int idx[128];
  for (int i = 0; i < window; ++i) idx[i] = i;
  mt19937_t g;
  mt_seed(&g, (uint32_t)seed);
  if (window > 1) {
    /* urngrange / urange >= urange always holds for window <= 64 */
    int i = 1;
    if ((window % 2) == 0) {
      int j = (int)libstdcxx_uniform_below(&g, 2);
      int t = idx[i]; idx[i] = idx[j]; idx[j] = t;
      ++i;
    }
    while (i != window) {
      const uint32_t swap_range = (uint32_t)i + 1;
      const uint64_t x = libstdcxx_uniform_below(&g, swap_range * (swap_range + 1));
      const int p0 = (int)(x / (swap_range + 1)), p1 = (int)(x % (swap_range + 1));
      int t = idx[i]; idx[i] = idx[p0]; idx[p0] = t; ++i;
      t = idx[i]; idx[i] = idx[p1]; idx[p1] = t; ++i;
    }
  }
  u128 m = 0;
  for (int i = 0; i < kmer_size && i < window; ++i) m |= (u128)3 << (2 * idx[i]);
  out[0] = lo64(m);
  out[1] = hi64(m);
}

/* ------------------------------------------------------------------------------------------
 * Sets.  kmer_set is an unordered_map keyed by (masked_bits, mask) (src/kmer.hpp:82-85,160-190);
 * with one mask per set the set is exactly the set of distinct masked_bits.  Canonical form here:
 * ascending array of 128-bit keys (lo, hi pairs).
 * ---------------------------------------------------------------------------------------- */
static int cmp_key(const void *a, const void *b) {
  const uint64_t *x = (const uint64_t *)a, *y = (const uint64_t *)b;
  if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
  if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
  return 0;
}
/* insert_kmers (src/kmer.hpp:170-178) + kmer_set_size (:186-189): in-place sort + dedup. */
int64_t orc_sort_unique(uint64_t *keys, int64_t n) {
  if (n <= 0) return 0;
  qsort(keys, (size_t)n, 16, cmp_key);
  int64_t m = 1;
  for (int64_t i = 1; i < n; ++i) {
    if (keys[2 * i] != keys[2 * m - 2] || keys[2 * i + 1] != keys[2 * m - 1]) {
      keys[2 * m] = keys[2 * i];
      keys[2 * m + 1] = keys[2 * i + 1];
      ++m;
    }
  }
  return m;
}
/* kmer_set_intersection, src/kmer_set.cpp:23-41: |A n B| (symmetric; the reference's swap only
 * chooses which side to iterate). */
int64_t orc_intersection(const uint64_t *a, int64_t na, const uint64_t *b, int64_t nb) {
  int64_t i = 0, j = 0, c = 0;
  while (i < na && j < nb) {
    int r = cmp_key(a + 2 * i, b + 2 * j);
    if (r == 0) { ++c; ++i; ++j; } else if (r < 0) ++i; else ++j;
  }
  return c;
}

/* containment / binomial_estimator, src/ani_estimation.cpp:24-28,38-42. */
double orc_containment(int intersection, int set_size) {
  if (intersection == 0) return 0;
  return ((double)intersection) / ((double)set_size);
}
double orc_binomial_estimator(double containment, int kmer_num_ones) {
  if (containment <= 0) return 0;
  return pow(containment, ((double)1.0) / ((double)kmer_num_ones));
}

/* generate_all_pairs_from_vector / generate_pairwise_from_vector on indices, src/generators.hpp:20-58. */
void orc_all_pairs(int n, int *first, int *second) {
  int64_t k = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) { first[k] = i; second[k] = j; ++k; }
}
void orc_ring_pairs(int n, int *first, int *second) {
  for (int i = 0; i < n; ++i) { first[i] = i; second[i] = (i + 1) % n; }
}

/* ------------------------------------------------------------------------------------------
 * Synthetic inputs (not in the reference; defined by SURVEY.md 4.2 KAT-3): splitmix64 stream.
 * ---------------------------------------------------------------------------------------- */
static inline uint64_t splitmix64_at(uint64_t seed, uint64_t i) { /* i-th output, i from 1 */
  uint64_t z = seed + i * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
void orc_gen(int64_t L, uint64_t seed, uint8_t *out) {
  for (int64_t i = 0; i < L; ++i) out[i] = (uint8_t)(splitmix64_at(seed, (uint64_t)i + 1) >> 62);
}
void orc_mutate(const uint8_t *in, int64_t L, uint64_t seed, uint64_t D, uint8_t *out) {
  for (int64_t i = 0; i < L; ++i) {
    uint64_t u = splitmix64_at(seed, (uint64_t)i + 1);
    uint8_t b = in[i];
    if (u % D == 0) b = (uint8_t)((b + 1 + ((u >> 32) % 3)) & 3);
    out[i] = b;
  }
}
