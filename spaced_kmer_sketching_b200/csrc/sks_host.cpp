// Host-only part of libsks.so: error reporting, seed masks, 2-bit packing, FASTA ingest, the Boost
// hash restatement used to fold per-launch constants, and the ANI arithmetic.  None of this needs a
// device.  Each function names the reference code it stands in for.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sks.h"
#include "sks_internal.cuh"

namespace sks {

static thread_local char g_error[512] = "";

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

// ---- Boost hash (un-vendored third party; call sites src/kmer.hpp:137-148) -------------------------
// hash_value(dynamic_bitset) = hc(128, hc(hc(0, block0), block1)); integers hash to themselves.
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
uint64_t boost_hash_bitset(uint64_t lo, uint64_t hi, int variant) {
  if (variant == SKS_HASH_BOOST_171) return hc171(128, hc171(hc171(0, lo), hi));
  return hc181(128, hc181(hc181(0, lo), hi));
}

// x % modulus == 0  <=>  ror(x * minv, mshift) <= mbound, where modulus = 2^mshift * d (d odd),
// minv = d^-1 mod 2^64 and mbound = floor((2^64 - 1) / modulus).
void modulus_magic(uint64_t modulus, uint64_t *minv, uint64_t *mbound, int *mshift) {
  int s = 0;
  uint64_t d = modulus;
  while ((d & 1) == 0) {
    d >>= 1;
    ++s;
  }
  uint64_t inv = d;  // Newton iteration: 3 correct bits double every step
  for (int i = 0; i < 6; ++i) inv *= 2 - d * inv;
  *minv = inv;
  *mbound = ~0ull / modulus;
  *mshift = s;
}

// PEXT(masked_bits, mask) as a table of rotate-and-mask pieces, up to kPiecesPerLimb per 32-bit limb.
int build_pext_table(const uint64_t mask[2], int n_limbs, PextTable *out, int *n_index_bits) {
  memset(out, 0, sizeof(*out));
  int dst = 0;
  for (int k = 0; k < n_limbs && k < 4; ++k) {
    const uint32_t limb = (uint32_t)(mask[k >> 1] >> (32 * (k & 1)));
    int b = 0, n = 0;
    while (b < 32) {
      if (!((limb >> b) & 1)) {
        ++b;
        continue;
      }
      int e = b;
      while (e < 32 && ((limb >> e) & 1)) ++e;
      const int len = e - b;
      if (n >= kPiecesPerLimb) return set_error(SKS_ERR_INVALID, "mask has too many runs in one limb");
      if (dst + len > 32) return set_error(SKS_ERR_INVALID, "mask weight exceeds 16: no 32-bit bitset index");
      out->rot[k][n] = (uint32_t)((b - dst) & 31);
      out->dmask[k][n] = (len == 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << dst;
      ++n;
      dst += len;
      b = e;
    }
    out->n_pieces[k] = (uint32_t)n;
  }
  *n_index_bits = dst;
  return SKS_OK;
}

}  // namespace sks

using sks::set_error;

extern "C" {

int sks_version(void) { return SKS_VERSION; }
const char *sks_last_error(void) { return sks::g_error; }
void sks_free(void *p) { free(p); }

int sks_seed_to_mask(const char *seed, uint64_t out_mask[2], int *out_window) {
  if (!seed || !out_mask) return set_error(SKS_ERR_INVALID, "null argument");
  const size_t w = strlen(seed);
  if (w < 1 || w > 64) return set_error(SKS_ERR_INVALID, "seed length %zu outside 1..64", w);
  unsigned __int128 m = 0;
  for (size_t i = 0; i < w; ++i) {
    if (seed[i] == '1')
      m |= (unsigned __int128)3 << (2 * (w - 1 - i));
    else if (seed[i] != '0')
      return set_error(SKS_ERR_INVALID, "seed strings hold only '0' and '1'");
  }
  out_mask[0] = (uint64_t)m;
  out_mask[1] = (uint64_t)(m >> 64);
  if (out_window) *out_window = (int)w;
  return SKS_OK;
}

int sks_mask_weight(const uint64_t mask[2]) {
  return (__builtin_popcountll(mask[0]) + __builtin_popcountll(mask[1])) / 2;
}

int sks_contiguous_mask(int k, uint64_t out_mask[2]) {
  // src/kmer_bitset.cpp:53-54 throws std::runtime_error for k > MAX_KMER_LENGTH
  if (k > 64) return set_error(SKS_ERR_INVALID, "Given k-mer length exceeds maximum k-mer length");
  unsigned __int128 m = (k <= 0) ? 0 : (k == 64 ? ~(unsigned __int128)0 : (((unsigned __int128)1 << (2 * k)) - 1));
  out_mask[0] = (uint64_t)m;
  out_mask[1] = (uint64_t)(m >> 64);
  return SKS_OK;
}

int sks_random_mask(int window, int k, uint64_t seed, uint64_t out_mask[2]) {
  // src/kmer_bitset.cpp:132-152 verbatim in behaviour: the platform's std::shuffle + std::mt19937
  // decide the mask, so this must run on libstdc++ to reproduce the reference's masks.
  if (window < 0 || window > 64 || k < 0 || k > window)
    return set_error(SKS_ERR_INVALID, "need 0 <= k <= window <= 64 (got window %d, k %d)", window, k);
  std::vector<int> idx(window);
  std::iota(idx.begin(), idx.end(), 0);
  std::shuffle(idx.begin(), idx.end(), std::mt19937(seed));
  unsigned __int128 m = 0;
  for (int i = 0; i < k; ++i) m |= (unsigned __int128)3 << (2 * idx[i]);
  out_mask[0] = (uint64_t)m;
  out_mask[1] = (uint64_t)(m >> 64);
  return SKS_OK;
}

void sks_reverse_bitset(const uint64_t in[2], uint64_t out[2]) {
  // src/kmer_bitset.cpp:105-119: reverse the 64 two-bit groups of the 128-bit value.
  auto rev64 = [](uint64_t v) {
    v = ((v >> 2) & 0x3333333333333333ULL) | ((v & 0x3333333333333333ULL) << 2);
    v = ((v >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((v & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(v);
  };
  const uint64_t lo = in[0], hi = in[1];
  out[0] = rev64(hi);
  out[1] = rev64(lo);
}

uint64_t sks_boost_hash_bitset(const uint64_t value[2], int hash_variant) {
  return sks::boost_hash_bitset(value[0], value[1], hash_variant == SKS_HASH_BOOST_171 ? SKS_HASH_BOOST_171 : SKS_HASH_BOOST_181);
}

uint64_t sks_fmh_hash(const uint64_t masked[2], const uint64_t mask[2], int window, int nonce, int hash_variant) {
  const int v = hash_variant == SKS_HASH_BOOST_171 ? SKS_HASH_BOOST_171 : SKS_HASH_BOOST_181;
  return sks::boost_hash_bitset(masked[0], masked[1], v) ^ sks::boost_hash_bitset(mask[0], mask[1], v) ^
         (uint64_t)(int64_t)window ^ (uint64_t)(int64_t)nonce;
}

double sks_containment(int intersection, int set_size) {  // src/ani_estimation.cpp:24-28
  if (intersection == 0) return 0;
  return ((double)intersection) / ((double)set_size);
}
double sks_binomial_estimator(double containment, int kmer_num_ones) {  // src/ani_estimation.cpp:38-42
  if (containment <= 0) return 0;
  return std::pow(containment, ((double)1.0) / ((double)kmer_num_ones));
}
void sks_ani_from_counts(const int32_t *intersections, const int32_t *first_set_sizes, int64_t n_pairs, int weight,
                         double *out_ani) {  // src/kmer-sketching.cpp:196-200
  auto span = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i)
      out_ani[i] = sks_binomial_estimator(sks_containment(intersections[i], first_set_sizes[i]), weight);
  };
  // a plain loop in the reference; a large matrix (10^6 pairs at C4) is split over host threads here, the
  // arithmetic per pair is unchanged (host double division + std::pow)
  const unsigned hw = std::thread::hardware_concurrency();
  const int64_t n_threads = std::min<int64_t>(std::min<int64_t>(hw ? hw : 1, 16), n_pairs / 16384);
  if (n_threads <= 1) {
    span(0, n_pairs);
    return;
  }
  std::vector<std::thread> pool;
  const int64_t per = (n_pairs + n_threads - 1) / n_threads;
  for (int64_t t = 0; t < n_threads; ++t) pool.emplace_back(span, t * per, std::min(n_pairs, (t + 1) * per));
  for (auto &th : pool) th.join();
}

size_t sks_packed_words(uint64_t n_bases) { return (size_t)((n_bases + 15) / 16); }

int sks_pack_codes(const uint8_t *codes, uint64_t n_bases, uint32_t *out_words) {
  const uint64_t full = n_bases / 16;
  for (uint64_t wi = 0; wi < full; ++wi) {
    const uint8_t *c = codes + wi * 16;
    uint32_t v = 0;
    for (int b = 0; b < 16; ++b) v |= (uint32_t)(c[b] & 3) << (2 * b);
    out_words[wi] = v;
  }
  if (n_bases % 16) {
    uint32_t v = 0;
    for (uint64_t b = 0; b < n_bases % 16; ++b) v |= (uint32_t)(codes[full * 16 + b] & 3) << (2 * b);
    out_words[full] = v;
  }
  return SKS_OK;
}

int sks_unpack_codes(const uint32_t *words, uint64_t n_bases, uint8_t *out_codes) {
  for (uint64_t i = 0; i < n_bases; ++i) out_codes[i] = (uint8_t)((words[i / 16] >> (2 * (i % 16))) & 3);
  return SKS_OK;
}

namespace {

// nucleotide_to_bits, src/fasta_processing.cpp:35-69, as a table: A/a 0, C/c 1, G/g 2, T/t 3, else 4.
struct CodeTable {
  uint8_t t[256];
  CodeTable() {
    memset(t, 4, sizeof(t));
    t[(unsigned char)'A'] = t[(unsigned char)'a'] = 0;
    t[(unsigned char)'C'] = t[(unsigned char)'c'] = 1;
    t[(unsigned char)'G'] = t[(unsigned char)'g'] = 2;
    t[(unsigned char)'T'] = t[(unsigned char)'t'] = 3;
  }
};
const CodeTable g_codes;

// Streaming 2-bit packer + segment table (the host side of the north star's fasta_processing).
struct Packer {
  uint32_t *words;  // may be null (sizing pass)
  uint64_t *seg_len;
  uint64_t n_bases = 0, n_segs = 0, cur = 0;
  uint32_t acc = 0;
  void flush_segment() {
    if (cur > 0) {
      if (seg_len) seg_len[n_segs] = cur;
      ++n_segs;
    }
    cur = 0;
  }
  // add_nucleotide_strings (src/fasta_processing.cpp:144-179) on a piece of a record; a record's
  // pieces are concatenated lines, so runs continue across calls until end_record().
  void add(const char *s, size_t n) {
    for (size_t i = 0; i < n; ++i) {
      const uint8_t code = g_codes.t[(unsigned char)s[i]];
      if (code & 4) {
        flush_segment();
      } else {
        acc |= (uint32_t)code << (2 * (n_bases & 15));
        if ((n_bases & 15) == 15) {
          if (words) words[n_bases >> 4] = acc;
          acc = 0;
        }
        ++n_bases;
        ++cur;
      }
    }
  }
  void end_record() { flush_segment(); }
  void finish() {
    if ((n_bases & 15) && words) words[n_bases >> 4] = acc;
  }
};

// strings_from_fasta, src/fasta_processing.cpp:79-133, restated as a one-pass state machine over the
// file bytes with std::getline's line rules.  A record's lines are buffered until the record is known
// to be kept (a later line with a space discards it, :114-118).
void parse_fasta_text(const char *text, size_t n, Packer &pk) {
  bool name_nonempty = false;
  std::vector<std::pair<size_t, size_t>> lines;  // (offset, length) of the current record's lines
  auto flush = [&]() {
    for (auto &ln : lines) pk.add(text + ln.first, ln.second);
    pk.end_record();
    lines.clear();
  };
  size_t pos = 0;
  while (pos < n) {
    const char *nl = static_cast<const char *>(memchr(text + pos, '\n', n - pos));
    const size_t eol = nl ? (size_t)(nl - text) : n;
    const char *line = text + pos;
    const size_t len = eol - pos;
    if (len == 0 || line[0] == '>') {
      if (name_nonempty) flush(); else lines.clear();
      if (len != 0) name_nonempty = len > 1;
    } else if (name_nonempty) {
      if (memchr(line, ' ', len)) {
        name_nonempty = false;
        lines.clear();
      } else {
        lines.emplace_back(pos, len);
      }
    }
    pos = eol + 1;
  }
  if (name_nonempty) flush();
  pk.finish();
}

}  // namespace

int sks_fasta_parse(const char *text, size_t n, uint64_t *n_bases, uint64_t *n_segs, uint32_t *out_words,
                    uint64_t *out_seg_len) {
  if (!n_bases || !n_segs || (!text && n)) return set_error(SKS_ERR_INVALID, "null argument");
  Packer pk{out_words, out_seg_len};
  parse_fasta_text(text, n, pk);
  *n_bases = pk.n_bases;
  *n_segs = pk.n_segs;
  return SKS_OK;
}

int sks_fasta_parse_file(const char *path, uint64_t *n_bases, uint64_t *n_segs, uint32_t **out_words,
                         uint64_t **out_seg_len) {
  if (!path || !n_bases || !n_segs || !out_words || !out_seg_len) return set_error(SKS_ERR_INVALID, "null argument");
  FILE *f = fopen(path, "rb");
  if (!f) return set_error(SKS_ERR_IO, "Unable to open %s", path);
  std::string text;
  if (fseek(f, 0, SEEK_END) == 0) {
    const long sz = ftell(f);
    if (sz > 0) text.reserve((size_t)sz);
    rewind(f);
  }
  char buf[1 << 16];
  size_t got;
  while ((got = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, got);
  fclose(f);
  // one pass: a file of n bytes holds at most n bases and n/2 + 1 segments (a run needs a separator)
  uint32_t *words = static_cast<uint32_t *>(malloc((text.size() / 16 + 2) * sizeof(uint32_t)));
  uint64_t *segs = static_cast<uint64_t *>(malloc((text.size() / 2 + 2) * sizeof(uint64_t)));
  if (!words || !segs) {
    free(words);
    free(segs);
    return set_error(SKS_ERR_INVALID, "out of host memory");
  }
  uint64_t nb = 0, ns = 0;
  sks_fasta_parse(text.data(), text.size(), &nb, &ns, words, segs);
  if (uint64_t *shrunk = static_cast<uint64_t *>(realloc(segs, (ns + 1) * sizeof(uint64_t)))) segs = shrunk;
  *n_bases = nb;
  *n_segs = ns;
  *out_words = words;
  *out_seg_len = segs;
  return SKS_OK;
}

}  // extern "C"
