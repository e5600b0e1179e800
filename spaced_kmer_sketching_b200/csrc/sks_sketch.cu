// The sketch kernel: fused K1 (stage 2-bit bases) + K2 (spaced-seed gather, canonical strand) +
// K3 (Boost-compatible FracMinHash filter, compaction) + K4 (presence-bitset insert).
//
// Replaces the reference's hot loop nucleotide_string_to_kmers (src/kmer_sliding.cpp:112-186):
//   per base  : update_kmer_window / update_complement_kmer_window        (:26-47,144-151)
//   per window: fwd & mask, rc & mask (same mask), canonical = min, ties -> rc (:159-175)
//               sketching_cond(kmer) -> push_back                          (:182-184)
// and, for OUT_BITSET, kmer_set::insert_kmers (src/kmer.hpp:170-178).
//
// Formulation (DESIGN.md has the derivation): with bases packed 16 per word, base i in bits
// 2(i%16), the reverse-complement window of start i is a plain right funnel shift of the
// complemented words, and the forward window is a left funnel shift of the 2-bit-group-reversed
// words.  A thread owns 16 consecutive window starts whose first is word aligned, so after one
// per-group alignment shift every per-window shift amount is a compile-time constant.
#include <type_traits>
#include <utility>

#include "sks_internal.cuh"

namespace sks {
namespace {

// ---- mbarrier + 1-D bulk copy (TMA engine, no tensor map needed) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// Reverse the sixteen 2-bit groups of a word: LE-packed bases -> MSB-first bases.
__device__ __forceinline__ uint32_t rev_groups(uint32_t x) {
  x = __brev(x);
  return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}

// ---- Boost hash restatement (src/kmer.hpp:137-148 call sites; see oracle/oracle.c header) --------
// x * m mod 2^64 in three instructions (IMAD.WIDE + two IMADs that accumulate into the upper half); the compiler's
// own schedule for `x * m` spends a fourth on a separate add, and the sketch kernel is bound by exactly these.
__device__ __forceinline__ uint64_t mul64(uint64_t x, uint64_t m) {
  uint64_t r;
  asm("{\n"
      " .reg .u32 xl, xh, ml, mh, pl, ph;\n"
      " .reg .u64 p;\n"
      " mov.b64 {xl, xh}, %1;\n"
      " mov.b64 {ml, mh}, %2;\n"
      " mul.wide.u32 p, xl, ml;\n"
      " mov.b64 {pl, ph}, p;\n"
      " mad.lo.u32 ph, xl, mh, ph;\n"
      " mad.lo.u32 ph, xh, ml, ph;\n"
      " mov.b64 %0, {pl, ph};\n"
      "}"
      : "=l"(r)
      : "l"(x), "l"(m));
  return r;
}
__device__ __forceinline__ uint64_t hc181(uint64_t h, uint64_t k) {
  const uint64_t M = 0x0e9846af9b1a615dULL;
  uint64_t x = h + 0x9e3779b9ULL + k;
  x ^= x >> 32;
  x = mul64(x, M);
  x ^= x >> 32;
  x = mul64(x, M);
  x ^= x >> 28;
  return x;
}
__device__ __forceinline__ uint64_t hc171(uint64_t h, uint64_t k) {
  const uint64_t m = 0xc6a4a7935bd1e995ULL;
  k *= m;
  k ^= k >> 47;
  k *= m;
  h ^= k;
  h *= m;
  h += 0xe6546b64ULL;
  return h;
}
template <int PRED>
__device__ __forceinline__ uint64_t bitset_hash(uint64_t b0, uint64_t b1) {
  if (PRED == PRED_FMH171) return hc171(128, hc171(hc171(0, b0), b1));
  return hc181(128, hc181(hc181(0, b0), b1));
}

// PEXT(masked, mask) as rotate-and-mask pieces (the table is built on the host, sks_api.cu)
template <int NL>
__device__ __forceinline__ uint32_t pext_index(const uint32_t (&c)[NL], const PextTable &pext) {
  uint32_t idx = 0;
#pragma unroll
  for (int k = 0; k < NL; ++k) {
#pragma unroll
    for (int p = 0; p < 4; ++p) idx |= __funnelshift_r(c[k], c[k], pext.rot[k][p]) & pext.dmask[k][p];
    if (pext.n_pieces[k] > 4) {  // uniform
#pragma unroll
      for (int p = 4; p < kPiecesPerLimb; ++p) idx |= __funnelshift_r(c[k], c[k], pext.rot[k][p]) & pext.dmask[k][p];
    }
  }
  return idx;
}

// Bucket of an 8-byte key in the bucket sort (sks_sets.cu, SortKey<1>::bucket): its top mask-selected bits.
__device__ __forceinline__ uint32_t kpart_bucket(unsigned long long k, const SortPlan &p) {
  uint32_t b = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
    if (i < p.n_pieces) b |= ((uint32_t)(k >> p.s[i]) & p.m[i]) << p.o[i];
  return b;
}

struct TileMeta {
  uint32_t genome;
  uint32_t t0;       // first window start of the tile, relative to the genome
  uint32_t n_bases;  // of the genome
  uint32_t seg_first;
  uint32_t n_segs;
  uint32_t pad[3];
};

// Staged / emitted element: the masked key (8 B up to 32 bases, else 16 B), or -- OUT_INDEX -- the
// 32-bit PEXT index of the key.
template <int NL, int OUT>
struct KeyType {
  using type = unsigned long long;
};
template <int OUT>
struct KeyType<3, OUT> {
  using type = ulonglong2;
};
template <int OUT>
struct KeyType<4, OUT> {
  using type = ulonglong2;
};
template <>
struct KeyType<1, OUT_PART> {
  using type = uint32_t;
};
template <>
struct KeyType<2, OUT_PART> {
  using type = uint32_t;
};
template <>
struct KeyType<3, OUT_PART> {
  using type = uint32_t;
};
template <>
struct KeyType<4, OUT_PART> {
  using type = uint32_t;
};
template <>
struct KeyType<1, OUT_INDEX> {
  using type = uint32_t;
};
template <>
struct KeyType<2, OUT_INDEX> {
  using type = uint32_t;
};
template <>
struct KeyType<3, OUT_INDEX> {
  using type = uint32_t;
};
template <>
struct KeyType<4, OUT_INDEX> {
  using type = uint32_t;
};

constexpr int kStageSlots = kSketchThreads * kGroup;  // kept k-mers staged per round (4096)
// OUT_BITSET: presence bitsets of up to 2^18 bits (weight <= 9) are accumulated per CTA in shared memory and OR-ed
// into HBM when the CTA moves on to another genome: millions of windows hammering a few thousand words with global
// atomics took 1.96 ms at C1 (weight 5: a 1024-bit bitset).
constexpr int kSmallBitsetWords = 8192;

template <int NL, int OUT>
constexpr size_t sketch_smem_bytes() {
  size_t b = 2 * kStageWords * 4 + 2 * sizeof(TileMeta) + 64;
  if (OUT != OUT_BITSET) b += kStageSlots * sizeof(typename KeyType<NL, OUT>::type);
  if (OUT == OUT_BITSET) b += kSmallBitsetWords * 4;  // CTA-private copy of a small presence bitset
  if (OUT == OUT_LIST) b += kStageSlots * 4;
  if (OUT == OUT_PART) b += kStageSlots * 2 + 2 * kMaxParts * 4;  // ranks, bucket histogram, bucket bases
  return b;
}

// Sparse-predicate instantiations run their 16 hash chains branch-free; asking for 4 resident CTAs lets the compiler
// spend up to 64 registers on interleaving them (measured at C3, 250 Mbp: 0.787 ms at 32 registers, 0.770 ms at 64).
template <int NL, int PRED, int OUT>
__global__ void __launch_bounds__(kSketchThreads, (PRED != PRED_ALL && OUT != OUT_PART) ? (NL <= 2 ? 4 : 3) : 1) sketch_kernel(const __grid_constant__ SketchParams P,
                                                               const uint32_t *__restrict__ tile_genome) {
  using key_t = typename KeyType<NL, OUT>::type;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t *s_words = reinterpret_cast<uint32_t *>(smem_raw);                        // [2][kStageWords]
  TileMeta *s_meta = reinterpret_cast<TileMeta *>(smem_raw + 2 * kStageWords * 4);    // [2]
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + 2 * kStageWords * 4 + 2 * sizeof(TileMeta));  // [2]
  uint32_t *s_count = reinterpret_cast<uint32_t *>(s_bar + 2);
  unsigned long long *s_base = reinterpret_cast<unsigned long long *>(s_bar + 3);
  key_t *s_keys = reinterpret_cast<key_t *>(smem_raw + 2 * kStageWords * 4 + 2 * sizeof(TileMeta) + 64);
  uint32_t *s_pos = reinterpret_cast<uint32_t *>(s_keys + kStageSlots);
  // OUT_PART: slot = j * 256 + tid is fixed; s_rank = rank of the index inside its bucket this round
  uint32_t *s_hist = reinterpret_cast<uint32_t *>(s_keys + kStageSlots);  // [kMaxParts]
  uint32_t *s_pbase = s_hist + kMaxParts;                                 // [kMaxParts]
  uint16_t *s_rank = reinterpret_cast<uint16_t *>(s_pbase + kMaxParts);   // [kStageSlots]

  // OUT_BITSET with a small bitset: CTA-private accumulation (the stage area of the other modes is unused here)
  uint32_t *s_bits = reinterpret_cast<uint32_t *>(s_keys);
  const bool small_bitset = OUT == OUT_BITSET && P.bitset_words <= (uint64_t)kSmallBitsetWords;
  uint32_t bits_genome = 0xFFFFFFFFu;  // genome whose bits s_bits holds
  auto flush_bits = [&](uint32_t genome) {  // uniform call; OR the CTA's bits into the genome's bitset, clear the copy
    __syncthreads();
    uint32_t *dst = P.bitset + (uint64_t)genome * P.bitset_words;
    for (uint32_t i = threadIdx.x; i < (uint32_t)P.bitset_words; i += kSketchThreads) {
      const uint32_t v = s_bits[i];
      if (v) {
        atomicOr(dst + i, v);
        s_bits[i] = 0;
      }
    }
    __syncthreads();
  };
  // few survivors per tile: stage across tiles (see the flush below); OUT_PART always works round by round
  constexpr bool kSparse = PRED != PRED_ALL && OUT != OUT_PART;
  const int tid = threadIdx.x;
  const int w = P.window;
  const int d2 = 2 * (w & 15);  // alignment shift of the forward stream
  const int wq = w >> 4;

  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    *s_count = 0;
    fence_barrier_init();
  }
  if (OUT == OUT_PART)
    for (uint32_t b = tid; b < P.n_parts; b += kSketchThreads) s_hist[b] = 0;
  if (small_bitset)
    for (uint32_t i = tid; i < (uint32_t)P.bitset_words; i += kSketchThreads) s_bits[i] = 0;
  __syncthreads();

  // Producer: describe tile `tile` in s_meta[stage] and start its bulk copy into s_words[stage].
  auto issue = [&](uint32_t tile, int stage) {
    const uint32_t g = (P.n_genomes == 1) ? 0u : tile_genome[tile];
    const GenomeDesc gd = P.genomes[g];
    const uint32_t tw = (tile - gd.tile_first) * (uint32_t)kTileWords;  // first data word of the tile
    TileMeta m;
    m.genome = g;
    m.t0 = tw * kBasesPerWord;
    m.n_bases = gd.n_bases;
    m.seg_first = gd.seg_first;
    m.n_segs = gd.n_segs;
    s_meta[stage] = m;
    if (!P.host_words) {  // uniform
      uint32_t avail = gd.n_words + kPreWords - tw;  // words readable from (word_off + tw - kPreWords)
      uint32_t copy_words = avail < (uint32_t)kStageWords ? avail : (uint32_t)kStageWords;
      const uint32_t *src = P.words + gd.word_off + tw - kPreWords;
      fence_proxy_async();
      mbar_expect_tx(&s_bar[stage], copy_words * 4);
      bulk_g2s(s_words + stage * kStageWords, src, copy_words * 4, &s_bar[stage]);
    } else {
      // The genome lies in the caller's pinned host buffer (sks_pair_ani): word_off is the absolute word address of
      // its first data word (16-byte aligned), n_words the exact number of data words, and nothing around them may
      // be read.  The bulk copy pulls the tile over PCIe while the other resident CTAs compute -- there is no
      // separate host-to-device copy.  The history words of the first tile are zeros; the last 1-3 words of a
      // genome whose length is not a multiple of 16 bytes are fetched with plain loads.
      uint32_t *dst = s_words + stage * kStageWords;
      const uint32_t skip = (tw == 0) ? (uint32_t)kPreWords : 0u;
      const uint32_t first = tw - kPreWords + skip;
      const uint32_t want = (uint32_t)kStageWords - skip, avail = gd.n_words - first;
      const uint32_t n = avail < want ? avail : want, bulk = n & ~3u;
      const uint32_t *src = P.words + gd.word_off + first;
      for (uint32_t i = 0; i < skip; ++i) dst[i] = 0;
      for (uint32_t i = bulk; i < n; ++i) dst[skip + i] = __ldcv(src + i);
      fence_proxy_async();
      mbar_expect_tx(&s_bar[stage], bulk * 4);
      if (bulk) bulk_g2s(dst + skip, src, bulk * 4, &s_bar[stage]);
    }
  };

  // Copies the staged survivors of `genome` to its output region: one global reservation per flush.
  // `staged` is the value of *s_count every thread read between two barriers (so it is uniform and nobody
  // is staging any more).
  // the first level of the bucket sort folded into the emit (8-byte keys only): see SketchParams::kpart_*
  auto kpart_put = [&](uint32_t genome, unsigned long long key) {
    const uint32_t region = (genome << P.kpart_bits) + kpart_bucket(key, P.kpart_plan);
    const uint32_t pos = atomicAdd(P.kpart_cursor + region, 1u);
    if (pos < P.kpart_cap) reinterpret_cast<unsigned long long *>(P.out_keys)[(size_t)region * P.kpart_cap + pos] = key;
    else *P.kpart_overflow = 1u;
  };
  auto flush = [&](uint32_t genome, uint32_t staged) {
    const uint32_t n = staged < (uint32_t)kStageSlots ? staged : (uint32_t)kStageSlots;  // the rest was spilled
    if (OUT == OUT_KEYS && NL <= 2 && P.kpart_bits != 0) {  // uniform
      if (n > 0) {
        if (tid == 0) *s_count = 0;
        __syncthreads();
        for (uint32_t i = tid; i < n; i += kSketchThreads)
          if (NL <= 2) kpart_put(genome, reinterpret_cast<const unsigned long long *>(s_keys)[i]);
        __syncthreads();
      }
      return;
    }
    if (n > 0) {      // uniform
      if (tid == 0) {
        *s_base = atomicAdd(P.out_count + genome, (unsigned long long)n);
        *s_count = 0;
      }
      __syncthreads();
      const unsigned long long base = *s_base;
      const unsigned long long cap = P.out_cap[genome], off = P.out_off[genome];
      key_t *out = reinterpret_cast<key_t *>(P.out_keys);
      for (uint32_t i = tid; i < n; i += kSketchThreads) {
        if (base + i < cap) {
          out[off + base + i] = s_keys[i];
          if (OUT == OUT_LIST) P.out_pos[off + base + i] = s_pos[i];
        }
      }
      __syncthreads();
    }
  };

  if (tid == 0 && P.tile_begin + blockIdx.x < P.n_tiles) issue(P.tile_begin + blockIdx.x, 0);

  for (uint32_t it = 0;; ++it) {
    const uint32_t tile = P.tile_begin + blockIdx.x + it * gridDim.x;
    if (tile >= P.n_tiles) break;
    const int stage = it & 1;
    if (tid == 0 && tile + gridDim.x < P.n_tiles) issue(tile + gridDim.x, stage ^ 1);
    mbar_wait(&s_bar[stage], (it >> 1) & 1);

    const TileMeta tm = s_meta[stage];
    if (small_bitset && tm.genome != bits_genome) {  // uniform: the tile belongs to another genome
      if (bits_genome != 0xFFFFFFFFu) flush_bits(bits_genome);
      bits_genome = tm.genome;
    }
    const uint32_t *sm = s_words + stage * kStageWords;
    const uint32_t n_bases = tm.n_bases;

    // Per-thread segment cursor: [.., seg_e) is the end of the segment holding the current position.
    uint32_t seg_i = tm.seg_first, seg_last = tm.seg_first + tm.n_segs - 1, seg_e = n_bases;
    const uint32_t first_pos = tm.t0 + tid * 32;
    if (tm.n_segs > 1 && first_pos < n_bases) {
      uint32_t lo = tm.seg_first, hi = seg_last;  // first segment whose end is > first_pos
      while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(P.seg_end + mid) > first_pos) hi = mid; else lo = mid + 1;
      }
      seg_i = lo;
      seg_e = __ldg(P.seg_end + lo);
    }

#pragma unroll 1
    for (int gi = 0; gi < 2; ++gi) {
      const uint32_t p0 = first_pos + gi * kGroup;
      const int wl = kPreWords + tid * 2 + gi;

      // ---- which of the 16 windows exist (inside one segment, fully) --------------------------
      uint32_t vmask = 0;
      if (p0 < n_bases) {
        while (seg_i < seg_last && p0 >= seg_e) seg_e = __ldg(P.seg_end + (++seg_i));
        if (p0 + (kGroup - 1) + w <= seg_e) {
          vmask = 0xFFFFu;
        } else {
          for (int j = 0; j < kGroup; ++j) {
            const uint32_t p = p0 + j;
            while (seg_i < seg_last && p >= seg_e) seg_e = __ldg(P.seg_end + (++seg_i));
            if (p < seg_e && p + w <= seg_e) vmask |= 1u << j;
          }
        }
      }

      if (vmask != 0) {
        // ---- stage the words of this group in registers ------------------------------------
        uint32_t rw[NL + 1], fa[NL + 1];
#pragma unroll
        for (int x = 0; x <= NL; ++x) rw[x] = sm[wl + x];
        {
          const int fb = wl + wq - NL;
          uint32_t be_prev = rev_groups(sm[fb]);
#pragma unroll
          for (int x = 0; x <= NL; ++x) {
            const uint32_t be = rev_groups(sm[fb + x + 1]);
            fa[x] = __funnelshift_l(be, be_prev, d2);
            be_prev = be;
          }
        }

        // K2: both strands of window j of the group through the same mask, canonical = smaller, ties -> rc.
        // `sh` = 2 * j: a compile-time constant in the unrolled loops, a run-time value when a candidate is re-derived.
        auto canonical = [&](const int sh, uint32_t (&c)[NL], bool &lt) {
          uint32_t f[NL], r[NL];
#pragma unroll
          for (int k = 0; k < NL; ++k) {
            f[k] = __funnelshift_l(fa[NL - k], fa[NL - 1 - k], sh) & P.mask[k];
            r[k] = ~__funnelshift_r(rw[k], rw[k + 1], sh) & P.mask[k];
          }
          if (NL == 1) {
            lt = f[0] < r[0];
          } else if (NL == 2) {
            lt = (((uint64_t)f[1] << 32) | f[0]) < (((uint64_t)r[1] << 32) | r[0]);
          } else if (NL == 3) {
            const uint64_t fl = ((uint64_t)f[1] << 32) | f[0], rl = ((uint64_t)r[1] << 32) | r[0];
            lt = (f[2] < r[2]) || (f[2] == r[2] && fl < rl);
          } else {
            const uint64_t fl = ((uint64_t)f[1] << 32) | f[0], rl = ((uint64_t)r[1] << 32) | r[0];
            const uint64_t fh = ((uint64_t)f[NL - 1] << 32) | f[NL > 2 ? 2 : 0];
            const uint64_t rh = ((uint64_t)r[NL - 1] << 32) | r[NL > 2 ? 2 : 0];
            lt = (fh < rh) || (fh == rh && fl < rl);
          }
#pragma unroll
          for (int k = 0; k < NL; ++k) c[k] = lt ? f[k] : r[k];
        };
        auto blocks = [&](const uint32_t (&c)[NL], uint64_t &b0, uint64_t &b1) {
          b0 = (NL >= 2) ? ((((uint64_t)c[NL >= 2 ? 1 : 0]) << 32) | c[0]) : (uint64_t)c[0];
          b1 = (NL == 3)   ? (uint64_t)c[NL >= 3 ? 2 : 0]
               : (NL == 4) ? ((((uint64_t)c[NL >= 4 ? 3 : 0]) << 32) | c[NL >= 3 ? 2 : 0])
                           : 0ull;
        };
        // K3/K4: emit the kept k-mer of window j (all modes but OUT_PART, which ranks inside the loop below)
        auto emit = [&](const int j, const uint32_t (&c)[NL], const bool lt, const uint64_t b0, const uint64_t b1) {
          uint32_t idx = 0;
          if (OUT == OUT_BITSET || OUT == OUT_INDEX) idx = pext_index<NL>(c, P.pext);
          if (OUT == OUT_BITSET) {
            const uint32_t bit = 1u << (idx & 31);
            if (small_bitset) {
              if (!(s_bits[idx >> 5] & bit)) atomicOr(&s_bits[idx >> 5], bit);  // mostly set already
            } else {
              // fire-and-forget RED into the L2-resident bitset (looking at the word first was measured: the
              // load's latency costs 2-4x more than the redundant atomics it saves)
              atomicOr(P.bitset + (uint64_t)tm.genome * P.bitset_words + (idx >> 5), bit);
            }
          } else {
            const uint32_t slot = atomicAdd(s_count, 1u);
            if (kSparse && slot >= (uint32_t)kStageSlots && OUT == OUT_KEYS && NL <= 2 && P.kpart_bits != 0) {
              kpart_put(tm.genome, b0);
            } else if (kSparse && slot >= (uint32_t)kStageSlots) {
              // the stage is full (dense survivors under a sparse-mode predicate): write this one directly
              const unsigned long long gs = atomicAdd(P.out_count + tm.genome, 1ull);
              if (gs < P.out_cap[tm.genome]) {
                const unsigned long long at = P.out_off[tm.genome] + gs;
                if (OUT == OUT_INDEX) {
                  reinterpret_cast<uint32_t *>(P.out_keys)[at] = idx;
                } else if (NL <= 2) {
                  reinterpret_cast<unsigned long long *>(P.out_keys)[at] = b0;
                } else {
                  reinterpret_cast<ulonglong2 *>(P.out_keys)[at] = make_ulonglong2(b0, b1);
                }
                if (OUT == OUT_LIST) P.out_pos[at] = (p0 + j) | (lt ? 0u : 0x80000000u);
              }
            } else if (OUT == OUT_INDEX) {
              reinterpret_cast<uint32_t *>(s_keys)[slot] = idx;
            } else if (NL <= 2) {
              reinterpret_cast<unsigned long long *>(s_keys)[slot] = b0;
            } else {
              reinterpret_cast<ulonglong2 *>(s_keys)[slot] = make_ulonglong2(b0, b1);
            }
            if (OUT == OUT_LIST && !(kSparse && slot >= (uint32_t)kStageSlots)) s_pos[slot] = (p0 + j) | (lt ? 0u : 0x80000000u);
          }
        };

        if (kSparse) {
          // ---- K3, sparse predicate: the 16 hash chains run branch-free (independent chains for the scheduler, no
          // control flow per window) and leave one candidate bit each; the few survivors (~1/c of the windows) are
          // re-derived from the staged words and emitted afterwards.
          // h % c == 0 with c = 2^s * d, d odd:  t = h * d^-1 (mod 2^64) has its low s bits clear and
          // t >> s <= floor((2^64 - 1) / c), i.e. t <= mbound << s.  The low-bit test looks at the lower word only,
          // which is the whole test for s <= 32; a modulus with more factors of two is re-tested below.
          uint32_t cand = 0;
          auto window = [&](auto J) {
            constexpr int j = decltype(J)::value;
            uint32_t c[NL];
            bool lt;
            uint64_t b0, b1;
            canonical(2 * j, c, lt);
            blocks(c, b0, b1);
            const uint64_t t = mul64(bitset_hash<PRED>(b0, b1) ^ P.hconst, P.minv);
            // cand |= ((t_lo & mlow_lo) == 0 && t <= mbound) << j in four instructions: the low-bit test feeds the
            // 64-bit compare as its combining predicate, the bit is ORed in under the result (the compiler's own
            // version spends two selects and an OR on it)
            asm("{\n .reg .pred q, p;\n .reg .b32 lowt;\n and.b32 lowt, %1, %2;\n setp.eq.u32 q, lowt, 0;\n"
                " setp.le.and.u64 p, %3, %4, q;\n @p or.b32 %0, %0, %5;\n}"
                : "+r"(cand)
                : "r"((uint32_t)t), "r"((uint32_t)P.mlow), "l"(t), "l"(P.mbound), "n"(1u << j));
          };
          static_assert(kGroup == 16, "the window calls below are written out");
#define SKS_WINDOW(j) window(std::integral_constant<int, j>{})
          SKS_WINDOW(0); SKS_WINDOW(1); SKS_WINDOW(2); SKS_WINDOW(3); SKS_WINDOW(4); SKS_WINDOW(5); SKS_WINDOW(6); SKS_WINDOW(7);
          SKS_WINDOW(8); SKS_WINDOW(9); SKS_WINDOW(10); SKS_WINDOW(11); SKS_WINDOW(12); SKS_WINDOW(13); SKS_WINDOW(14); SKS_WINDOW(15);
#undef SKS_WINDOW
          cand &= vmask;
          const bool wide_low = (P.mlow >> 32) != 0;  // uniform; s > 32
#pragma unroll 1
          while (cand) {
            const int j = __ffs(cand) - 1;
            cand &= cand - 1;
            uint32_t c[NL];
            bool lt;
            uint64_t b0, b1;
            canonical(2 * j, c, lt);
            blocks(c, b0, b1);
            if (wide_low) {
              const uint64_t t = mul64(bitset_hash<PRED>(b0, b1) ^ P.hconst, P.minv);
              if ((t & P.mlow) != 0) continue;
            }
            emit(j, c, lt, b0, b1);
          }
        } else if (OUT == OUT_PART && PRED == PRED_ALL && vmask == 0xFFFFu) {
          // ---- every window of the group exists and is kept (the common case of the dense bitset build): no control
          // flow per window, so the 16 rank atomics are in flight together
#pragma unroll
          for (int j = 0; j < kGroup; ++j) {
            uint32_t c[NL];
            bool lt;
            canonical(2 * j, c, lt);
            const uint32_t idx = pext_index<NL>(c, P.pext);
            reinterpret_cast<uint32_t *>(s_keys)[j * kSketchThreads + tid] = idx;
            s_rank[j * kSketchThreads + tid] = (uint16_t)atomicAdd(&s_hist[idx >> P.part_shift], 1u);
          }
        } else {
#pragma unroll
          for (int j = 0; j < kGroup; ++j) {
            uint32_t c[NL];
            bool lt;
            uint64_t b0, b1;
            canonical(2 * j, c, lt);
            blocks(c, b0, b1);

            // ---- K3: predicate -----------------------------------------------------------------
            bool pass = (vmask >> j) & 1u;
            if (PRED != PRED_ALL) {
              const uint64_t t = mul64(bitset_hash<PRED>(b0, b1) ^ P.hconst, P.minv);
              pass = pass && ((t & P.mlow) == 0) && (t <= P.mbound);
            }

            // ---- K3/K4: emit -------------------------------------------------------------------
            if (OUT == OUT_PART) {
              if (pass) {
                const uint32_t idx = pext_index<NL>(c, P.pext);
                reinterpret_cast<uint32_t *>(s_keys)[j * kSketchThreads + tid] = idx;
                s_rank[j * kSketchThreads + tid] = (uint16_t)atomicAdd(&s_hist[idx >> P.part_shift], 1u);
              } else {
                s_rank[j * kSketchThreads + tid] = 0xFFFFu;
              }
            } else if (pass) {
              emit(j, c, lt, b0, b1);
            }
          }
        }
      }

      if (OUT == OUT_PART) {
        // ---- K4a fused: scatter the round's indices into their (genome, bucket) regions -----------------
        if (vmask == 0) {
#pragma unroll
          for (int j = 0; j < kGroup; ++j) s_rank[j * kSketchThreads + tid] = 0xFFFFu;
        }
        __syncthreads();
        uint32_t *cursor = P.part_cursor + (size_t)tm.genome * P.n_parts;
        for (uint32_t b = tid; b < P.n_parts; b += kSketchThreads) {
          const uint32_t h = s_hist[b];
          s_pbase[b] = h ? atomicAdd(cursor + b, h) : 0u;  // one reservation per non-empty bucket per round
          s_hist[b] = 0;
        }
        __syncthreads();
        uint32_t *out = reinterpret_cast<uint32_t *>(P.out_keys);
        const uint32_t region0 = tm.genome * P.n_parts;
#pragma unroll
        for (int j = 0; j < kGroup; ++j) {
          const uint32_t r = s_rank[j * kSketchThreads + tid];
          if (r != 0xFFFFu) {
            const uint32_t idx = reinterpret_cast<uint32_t *>(s_keys)[j * kSketchThreads + tid];
            const uint32_t b = idx >> P.part_shift;
            const uint32_t at = s_pbase[b] + r;  // slot inside the (genome, bucket) region
            if (at < P.part_cap) out[(size_t)(region0 + b) * P.part_cap + at] = idx;
            else *P.part_overflow = 1u;
          }
        }
        __syncthreads();
      }
      // ---- dense mode: flush the round's kept k-mers (every window may survive) ---------------------
      if (OUT != OUT_BITSET && OUT != OUT_PART && !kSparse) {
        __syncthreads();
        const uint32_t staged = *s_count;
        __syncthreads();
        flush(tm.genome, staged);
      }
    }
    __syncthreads();  // everyone is done with s_words[stage] / s_meta[stage] before it is refilled
    // ---- sparse mode (a FracMinHash filter keeps ~1/c of the windows): survivors of many tiles share the
    // stage; flush when it is half full, when the next tile belongs to another genome, or at the end
    if (OUT != OUT_BITSET && OUT != OUT_PART && kSparse) {
      const uint32_t staged = *s_count;
      const bool last = tile + gridDim.x >= P.n_tiles;
      const bool do_flush = staged >= (uint32_t)kStageSlots / 2 || last || s_meta[stage ^ 1].genome != tm.genome;
      __syncthreads();  // everyone has read s_count / s_meta before the next tile touches them
      if (do_flush) flush(tm.genome, staged);
    }
  }
  if (small_bitset && bits_genome != 0xFFFFFFFFu) flush_bits(bits_genome);
}

template <int NL, int PRED, int OUT>
int launch_one(sks_ctx *ctx, const SketchParams &p, const uint32_t *tile_genome) {
  auto kern = sketch_kernel<NL, PRED, OUT>;
  constexpr size_t smem = sketch_smem_bytes<NL, OUT>();
  static thread_local int ctas_per_sm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int &occ = ctas_per_sm[ctx->device & 7];
  if (occ == 0) {
    SKS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SKS_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSketchThreads, smem));
    if (occ < 1) occ = 1;
  }
  const uint32_t n_tiles = p.n_tiles - p.tile_begin;
  if (p.n_tiles <= p.tile_begin) return SKS_OK;
  // Persistent CTAs: a whole number of waves of resident CTAs, capped by the tile count.
  uint32_t grid = (uint32_t)(ctx->sm_count * occ);
  // fewer than two tiles per resident CTA: one CTA per tile balances better than a 1-or-2 split (C2: 88 -> 83 us)
  if (n_tiles <= 2 * grid) grid = n_tiles;
  KernelTimer timer(ctx, SKS_KERNEL_SKETCH);
  kern<<<grid, kSketchThreads, smem, ctx->stream>>>(p, tile_genome);
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches++;
  return SKS_OK;
}

template <int NL, int PRED>
int launch_out(sks_ctx *ctx, const SketchParams &p, const uint32_t *tg, int out_mode) {
  switch (out_mode) {
    case OUT_KEYS: return launch_one<NL, PRED, OUT_KEYS>(ctx, p, tg);
    case OUT_BITSET: return launch_one<NL, PRED, OUT_BITSET>(ctx, p, tg);
    case OUT_LIST: return launch_one<NL, PRED, OUT_LIST>(ctx, p, tg);
    case OUT_INDEX: return launch_one<NL, PRED, OUT_INDEX>(ctx, p, tg);
    case OUT_PART: return launch_one<NL, PRED, OUT_PART>(ctx, p, tg);
  }
  return set_error(SKS_ERR_INVALID, "bad output mode %d", out_mode);
}
template <int NL>
int launch_pred(sks_ctx *ctx, const SketchParams &p, const uint32_t *tg, int pred_mode, int out_mode) {
  switch (pred_mode) {
    case PRED_ALL: return launch_out<NL, PRED_ALL>(ctx, p, tg, out_mode);
    case PRED_FMH181: return launch_out<NL, PRED_FMH181>(ctx, p, tg, out_mode);
    case PRED_FMH171: return launch_out<NL, PRED_FMH171>(ctx, p, tg, out_mode);
  }
  return set_error(SKS_ERR_INVALID, "bad predicate mode %d", pred_mode);
}

}  // namespace

int launch_sketch(sks_ctx *ctx, const SketchParams &p, const uint32_t *tile_genome, int n_limbs, int pred_mode,
                  int out_mode) {
  switch (n_limbs) {
    case 1: return launch_pred<1>(ctx, p, tile_genome, pred_mode, out_mode);
    case 2: return launch_pred<2>(ctx, p, tile_genome, pred_mode, out_mode);
    case 3: return launch_pred<3>(ctx, p, tile_genome, pred_mode, out_mode);
    case 4: return launch_pred<4>(ctx, p, tile_genome, pred_mode, out_mode);
  }
  return set_error(SKS_ERR_INVALID, "window needs %d limbs", n_limbs);
}

}  // namespace sks
