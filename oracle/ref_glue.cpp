// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or called from the product.
//
// extern "C" glue over the UNMODIFIED reference sources (compiled where they lie under
// /root/reference/src against oracle/shim).  It lets tests/ and bench.py's CPU-baseline
// legs drive the reference's own public functions (src/kmer.hpp:57-64,89-103,192-218,
// src/fasta_processing.hpp:18-23, src/ani_estimator.hpp:13-14) from ctypes.
// Built by oracle/Makefile into oracle/_ref/libref.so (git-ignored).
#include "kmer.hpp"
#include "ani_estimator.hpp"
#include "fasta_processing.hpp"
#include "generators.hpp"

#include <algorithm>
#include <cstring>

// Defined (external linkage) in src/kmer_sliding.cpp:61-98, not declared in kmer.hpp.
void nucleotide_string_to_kmers_OLD_reverse(std::vector<kmer> &kmer_list,
                                            const acgt_string &nucleotide_string,
                                            const kmer_bitset &mask, const int window_length,
                                            const std::function<bool(const kmer)> &sketching_cond);

namespace {

kmer_bitset bitset_from_words(const uint64_t w[2]) {
  kmer_bitset b(KMER_BITSET_SIZE);
  b.shim_blocks()[0] = w[0];
  b.shim_blocks()[1] = w[1];
  return b;
}
void bitset_to_words(const kmer_bitset &b, uint64_t w[2]) {
  w[0] = b.shim_blocks()[0];
  w[1] = b.shim_blocks()[1];
}

// pred_kind 0: every k-mer.  pred_kind 1: the driver's FracMinHash condition,
// src/kmer-sketching.cpp:29-34 generalised over (nonce, c): fmh(k) % c == 0.
std::function<bool(const kmer)> make_pred(int pred_kind, int nonce, int modulus) {
  if (pred_kind == 0) return [](const kmer) { return true; };
  frac_min_hash fmh(nonce);
  const int c = modulus;
  return [fmh, c](const kmer k) { return fmh(k) % c == 0; };
}

struct key128 {
  uint64_t lo, hi;
  bool operator<(const key128 &o) const { return hi != o.hi ? hi < o.hi : lo < o.lo; }
};

}  // namespace

extern "C" {

void ref_set_boost_variant(int v) { boost::shim::hash_variant() = v; }
int ref_get_boost_variant() { return boost::shim::hash_variant(); }

void ref_init() {
  initialise_contiguous_kmer_array();
  initialise_reversing_kmer_array();
}

void ref_random_mask(int window, int k, uint64_t seed, uint64_t out[2]) {
  bitset_to_words(generate_random_spaced_seed_mask(window, k, seed), out);
}
int ref_contiguous_mask(int k, uint64_t out[2]) {
  try {
    bitset_to_words(contiguous_kmer(k), out);
  } catch (const std::runtime_error &) {
    return -1;
  }
  return 0;
}
void ref_reverse_bitset(const uint64_t in[2], uint64_t out[2]) {
  bitset_to_words(reverse_kmer_bitset(bitset_from_words(in)), out);
}
uint64_t ref_fmh(int nonce, int window, const uint64_t masked[2], const uint64_t mask[2]) {
  frac_min_hash fmh(nonce);
  kmer k{window, bitset_from_words(masked), bitset_from_words(mask), bitset_from_words(masked)};
  return fmh(k);
}
uint64_t ref_boost_hash_bitset(const uint64_t w[2]) {
  return boost::hash<kmer_bitset>()(bitset_from_words(w));
}
// reverse_complement / canonical_kmer of src/kmers.cpp:16-35 on (window, kmer_bits, mask).
void ref_canonical_kmer(int window, const uint64_t bits[2], const uint64_t mask[2],
                        uint64_t out_bits[2], uint64_t out_masked[2]) {
  kmer_bitset b = bitset_from_words(bits), m = bitset_from_words(mask);
  kmer k{window, b, m, b & m};
  kmer c = canonical_kmer(k);
  bitset_to_words(c.kmer_bits, out_bits);
  bitset_to_words(c.masked_bits, out_masked);
}

// ---- nucleotide string lists -------------------------------------------------------
typedef std::vector<acgt_string> strings_t;

void *ref_strings_from_fasta(const char *path) {
  return new strings_t(nucleotide_strings_from_fasta_file(path));
}
void *ref_strings_from_codes(const uint8_t *codes, const int64_t *lens, int n) {
  strings_t *s = new strings_t();
  int64_t off = 0;
  for (int i = 0; i < n; ++i) {
    s->emplace_back(codes + off, codes + off + lens[i]);
    off += lens[i];
  }
  return s;
}
// add_nucleotide_strings (src/fasta_processing.cpp:144-179) on one raw text string.
void *ref_strings_from_raw(const char *raw, int64_t n) {
  strings_t *s = new strings_t();
  add_nucleotide_strings(*s, std::string(raw, raw + n));
  return s;
}
int64_t ref_strings_count(void *h) { return (int64_t) static_cast<strings_t *>(h)->size(); }
int64_t ref_string_len(void *h, int64_t i) { return (int64_t)(*static_cast<strings_t *>(h))[i].size(); }
void ref_string_copy(void *h, int64_t i, uint8_t *out) {
  const acgt_string &s = (*static_cast<strings_t *>(h))[i];
  std::memcpy(out, s.data(), s.size());
}
void ref_strings_free(void *h) { delete static_cast<strings_t *>(h); }

// ---- ordered k-mer lists -----------------------------------------------------------
typedef std::vector<kmer> kmers_t;

void *ref_kmers(void *strings, const uint64_t mask[2], int window, int pred_kind, int nonce,
                int modulus, int legacy) {
  const strings_t &s = *static_cast<strings_t *>(strings);
  kmer_bitset m = bitset_from_words(mask);
  auto pred = make_pred(pred_kind, nonce, modulus);
  if (!legacy) return new kmers_t(nucleotide_string_list_to_kmers(s, m, window, pred));
  kmers_t *out = new kmers_t();
  for (const acgt_string &str : s) nucleotide_string_to_kmers_OLD_reverse(*out, str, m, window, pred);
  return out;
}
int64_t ref_kmers_count(void *h) { return (int64_t) static_cast<kmers_t *>(h)->size(); }
// out_masked / out_bits: 2 words per k-mer (lo, hi), in list order.
void ref_kmers_copy(void *h, uint64_t *out_masked, uint64_t *out_bits) {
  const kmers_t &v = *static_cast<kmers_t *>(h);
  for (size_t i = 0; i < v.size(); ++i) {
    if (out_masked) bitset_to_words(v[i].masked_bits, out_masked + 2 * i);
    if (out_bits) bitset_to_words(v[i].kmer_bits, out_bits + 2 * i);
  }
}
void ref_kmers_free(void *h) { delete static_cast<kmers_t *>(h); }

// ---- sets --------------------------------------------------------------------------
void *ref_set_from_kmers(void *kmers) {
  kmer_set *ks = new kmer_set();
  ks->insert_kmers(*static_cast<kmers_t *>(kmers));
  return ks;
}
void *ref_set_from_fasta(const char *path, const uint64_t mask[2], int window, int pred_kind,
                         int nonce, int modulus) {
  return new kmer_set(kmer_set_from_fasta_file(path, bitset_from_words(mask), window,
                                               make_pred(pred_kind, nonce, modulus)));
}
// parallel_kmer_sets_from_fasta_files (src/kmer_set.cpp:112-133); out_handles[n].
void ref_sets_from_fasta_files(int n, char **paths, const uint64_t mask[2], int window,
                               int pred_kind, int nonce, int modulus, int parallel,
                               void **out_handles) {
  auto pred = make_pred(pred_kind, nonce, modulus);
  kmer_bitset m = bitset_from_words(mask);
  std::vector<kmer_set> sets = parallel ? parallel_kmer_sets_from_fasta_files(n, paths, m, window, pred)
                                        : kmer_sets_from_fasta_files(n, paths, m, window, pred);
  for (int i = 0; i < n; ++i) out_handles[i] = new kmer_set(std::move(sets[i]));
}
int ref_set_size(void *h) { return static_cast<kmer_set *>(h)->kmer_set_size(); }
// Sorted (as unsigned 128-bit) masked_bits of the set members; 2 words per key.
void ref_set_keys(void *h, uint64_t *out) {
  const kmer_set &ks = *static_cast<kmer_set *>(h);
  std::vector<key128> keys;
  keys.reserve(ks.kmer_hashes.size());
  for (const auto &it : ks.kmer_hashes) {
    uint64_t w[2];
    bitset_to_words(it.first.masked_bits, w);
    keys.push_back({w[0], w[1]});
  }
  std::sort(keys.begin(), keys.end());
  for (size_t i = 0; i < keys.size(); ++i) { out[2 * i] = keys[i].lo; out[2 * i + 1] = keys[i].hi; }
}
void ref_set_free(void *h) { delete static_cast<kmer_set *>(h); }
int ref_intersection(void *a, void *b) {
  return kmer_set_intersection(*static_cast<kmer_set *>(a), *static_cast<kmer_set *>(b));
}
// (parallel_)compute_pairwise_kmer_set_intersections, src/kmer_set.cpp:143-184.
// Returns 0, or -1 if the reference threw std::runtime_error (length mismatch).
int ref_pairwise_intersections(void **a, int na, void **b, int nb, int parallel, int *out) {
  std::vector<kmer_set *> va(na), vb(nb);
  for (int i = 0; i < na; ++i) va[i] = static_cast<kmer_set *>(a[i]);
  for (int i = 0; i < nb; ++i) vb[i] = static_cast<kmer_set *>(b[i]);
  try {
    std::vector<int> r = parallel ? parallel_compute_pairwise_kmer_set_intersections(va, vb)
                                  : compute_pairwise_kmer_set_intersections(va, vb);
    std::copy(r.begin(), r.end(), out);
  } catch (const std::runtime_error &) {
    return -1;
  }
  return 0;
}
// generate_all_pairs_from_vector / generate_pairwise_from_vector on indices 0..n-1.
void ref_all_pairs(int n, int *first, int *second) {
  std::vector<int> v(n);
  for (int i = 0; i < n; ++i) v[i] = i;
  auto p = generate_all_pairs_from_vector<int>(v);
  std::copy(p.first.begin(), p.first.end(), first);
  std::copy(p.second.begin(), p.second.end(), second);
}
void ref_ring_pairs(int n, int *first, int *second) {
  std::vector<int> v(n);
  for (int i = 0; i < n; ++i) v[i] = i;
  auto p = generate_pairwise_from_vector<int>(v);
  std::copy(p.first.begin(), p.first.end(), first);
  std::copy(p.second.begin(), p.second.end(), second);
}

double ref_containment(int intersection, int set_size) { return containment(intersection, set_size); }
double ref_binomial_estimator(double c, int k) { return binomial_estimator(c, k); }

}  // extern "C"
