// All-vs-all intersection counts of many sorted k-mer sets through a dictionary of their SHARED k-mers.
//
// The reference evaluates every ordered pair on its own: `parallel_compute_pairwise_kmer_set_intersections` probes
// the larger hash map with every element of the smaller one (src/kmer_set.cpp:23-41,167-184), n^2 * |sketch| probes
// for `generate_all_pairs_from_vector` (src/generators.hpp:44-58).  Only k-mers that occur in at least two sets can
// ever be counted; of those, most are either shared by a handful of genomes or by a large part of them.  So:
//
//   D1  every key of every set goes into one device-wide open-addressing table that counts its occurrences (the sets
//       hold distinct keys, so occurrences == sets that hold the key); the second occurrence of a key gives it the
//       next dense 32-bit id, and every entry remembers which occurrence it was;
//   D2  keys held by 2 .. kPostMax sets become POSTING LISTS (the sets that hold the key); keys held by more sets
//       are re-coded per set as id bitmaps: the set's ids grouped by id range (2^16 ids), a range the set is dense
//       in (>= 1024 ids) as an 8 KB bitmap, a sparse one as a 16-bit id list; private keys are dropped;
//   X1  a posting list of c sets adds 1 to each of its c (c - 1) ordered pairs;
//   X2  |A n B| += sum over the ranges both sets are present in of popc(bitmap_A & bitmap_B) (or of list probes
//       into A's bitmap, which sits in shared memory for a whole chunk of columns);
//   F   counts (+ the diagonal |A n A| = |A|), set sizes and ANI = (|A n B| / |A|)^(1/weight)
//       (src/ani_estimation.cpp:24-42, src/kmer-sketching.cpp:196-200) are finalised on the device.
//
// Nothing returns to the host in between; the counts are bit-identical to the pairwise kernels' (tests compare
// both with the oracle).  16-byte keys are first compacted through the mask (PEXT) to 64 bits, which needs
// weight <= 32.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cub/cub.cuh>

#include "sks_internal.cuh"

namespace sks {
namespace {

constexpr int kRangeBits = 16;                      // ids per range
constexpr uint32_t kRangeIds = 1u << kRangeBits;
constexpr uint32_t kRangeWords = kRangeIds / 32;    // 2048 words = 8 KB bitmap = 512 x 16 B
constexpr uint32_t kRange16 = kRangeWords / 4;      // the same in 16-byte units
constexpr uint32_t kDenseMin = 1024;                // ids of a (set, range) group from which it is stored as a bitmap
constexpr uint32_t kNoId = 0xFFFFFFFFu;
constexpr int kDictThreads = 256;
constexpr int kDictPer = 4;                         // entries per thread
constexpr int kDictChunk = kDictThreads * kDictPer;
constexpr int kPairsThreads = 256;
constexpr int kPairsWarps = kPairsThreads / 32;
constexpr int kPairsCols = 64;                      // columns per task

// PEXT of a 16-byte key through the mask as runs of mask ones (n_runs == 0: 8-byte keys are used as they are).
struct Compact {
  int n_runs;
  uint8_t src[32], len[32], dst[32];
};

struct Slot {  // 16 bytes: one sector holds the key and its counters
  unsigned long long key;  // 0 = empty (key 0 itself lives in the extra slot `cap`)
  uint32_t cnt;            // occurrences == sets that hold the key
  uint32_t id;             // dense id, given by the second occurrence
};
constexpr uint32_t kPostMax = 16;          // keys held by at most this many sets are posting lists
constexpr uint32_t kPostFlag = 0x80000000u;  // entry = kPostFlag | id: the key is a posting list

// C_IDS numbers the keys held by at least two sets (their posting lists), C_HIGH those held by more than kPostMax
// sets (their bits in the sets' bitmaps): a key gets the first kind of id at its second occurrence and the second
// kind at occurrence kPostMax + 1.
enum Counter : int { C_IDS = 0, C_OVERFLOW = 1, C_ROWS = 2, C_TASK = 3, C_MINE = 4, C_HIGH = 5, C_COUNT = 16 };
constexpr uint32_t kHighFlag = 0x80000000u;  // Slot::id = kHighFlag | bitmap id once the key has one (atomicMax keeps it)
// entry[q] after D1: the slot of the key (< 2^31), or kHighFlag | bitmap id when the probe already saw the key with more
// than kPostMax occurrences and its bitmap id (most occurrences of widely shared keys): D2a then has nothing to look up
// for them.

struct DictView {
  const void *const *set_ptr;  // [n] keys of set s
  const uint32_t *set_off;     // [n + 1] first entry of set s in the flattened numbering of all keys
  uint32_t n_sets;
  uint32_t n_entries;
  Slot *tab;                   // [cap + 1]
  uint32_t cap;
  // Per entered key occurrence q.  One rank: q = the entry's number e, its set is found from set_off.  Several ranks
  // (n_parts > 1): the occurrences of this rank's keys are compacted, q counts them (counters[C_MINE]) and set_of[q]
  // is the set.
  uint32_t *entry;             // [n_entries] D1: slot of the key; D2: its id (| kPostFlag), kNoId: private key
  uint8_t *occ;                // [n_entries] which occurrence of its key the entry was (saturating at 255)
  uint16_t *set_of;            // [n_entries] (n_parts > 1 only)
  uint32_t *counters;          // [C_COUNT]
  uint32_t part, n_parts;      // only keys with owner(key) == part are entered (several ranks split the key space)
  // FLAT source: the occurrences arrive as arrays (the keys this rank owns, routed here by the other ranks):
  // flat_keys[q] is the 64-bit key, set_of[q] its set; set_ptr / set_off are not used.
  const unsigned long long *flat_keys;
};

// A (set, range) group becomes a bitmap from this many ids on: 1024 in a range that is in full use (a list of 16-bit
// ids is then a quarter of the 8 KB bitmap), less in a range of which only the first few thousand ids exist (the
// last range; the only one when several ranks split the key space), where the bitmap is short as well.
__device__ __forceinline__ uint32_t dense_min(uint32_t n_high, uint32_t t) {
  const uint32_t first = t * kRangeIds;
  const uint32_t ids_here = n_high > first ? min(kRangeIds, n_high - first) : 0u;
  return max(64u, min(kDenseMin, ids_here / 16));
}
__device__ __forceinline__ uint32_t size16(uint32_t cnt, uint32_t dmin) {  // payload of a group in 16-byte units
  return cnt == 0 ? 0u : (cnt >= dmin ? kRange16 : (2 * cnt + 15) / 16);
}
struct GroupSize16 {  // group g -> its payload size (for the prefix sum over all groups)
  const uint32_t *group_cnt, *counters;
  uint32_t n_sets;
  __device__ __forceinline__ uint32_t operator()(uint32_t g) const {
    return size16(group_cnt[g], dense_min(counters[5 /* C_HIGH */], g / n_sets));
  }
};

template <int KW>
__device__ __forceinline__ unsigned long long load_key(const void *base, uint32_t i, const Compact &c) {
  if (KW == 1) return static_cast<const unsigned long long *>(base)[i];
  const ulonglong2 k = static_cast<const ulonglong2 *>(base)[i];
  unsigned long long out = 0;
  for (int r = 0; r < c.n_runs; ++r) {
    const int s = c.src[r];
    const unsigned long long v = s >= 64 ? (k.y >> (s - 64)) : (s == 0 ? k.x : ((k.x >> s) | (k.y << (64 - s))));
    const unsigned long long m = c.len[r] >= 64 ? ~0ull : ((1ull << c.len[r]) - 1);
    out |= (v & m) << c.dst[r];
  }
  return out;
}

__device__ __forceinline__ unsigned long long hash_key(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}
// the table slot comes from the upper half of the hash, the owning rank from the lower half
__device__ __forceinline__ uint32_t hash_slot(unsigned long long h, uint32_t cap) { return __umulhi((uint32_t)(h >> 32), cap); }
__device__ __forceinline__ uint32_t hash_owner(unsigned long long h, uint32_t n_parts) { return (uint32_t)h % n_parts; }

// Largest s with set_off[s] <= e (empty sets share their offset with their successor and are never returned).
__device__ __forceinline__ uint32_t find_set(const uint32_t *__restrict__ set_off, uint32_t n_sets, uint32_t e) {
  uint32_t lo = 0, hi = n_sets;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(set_off + mid) <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// D1: insert every key (of this rank's share of the key space), count its occurrences, remember its slot and which
// occurrence the entry was.  A key that is already known to be held by more than kPostMax sets needs neither: its
// entries only look it up.
template <int KW>
__global__ void __launch_bounds__(kDictThreads) dict_insert_kernel(const __grid_constant__ DictView D,
                                                                   const __grid_constant__ Compact C) {
  const uint32_t base = blockIdx.x * (uint32_t)kDictChunk;
  const uint32_t lane = threadIdx.x & 31;
  const bool flat = D.flat_keys != nullptr;
  const bool compact = D.n_parts > 1 && !flat;
  uint32_t s = 0, slot[kDictPer], set_id[kDictPer];
  unsigned long long key[kDictPer];
  uint4 cur[kDictPer];
  bool have = false;
#pragma unroll
  for (int u = 0; u < kDictPer; ++u) {
    const uint32_t e = base + u * kDictThreads + threadIdx.x;
    slot[u] = kNoId;
    set_id[u] = 0;
    if (e < D.n_entries && flat) {
      key[u] = D.flat_keys[e];
      slot[u] = key[u] ? hash_slot(hash_key(key[u]), D.cap) : D.cap;
    } else if (e < D.n_entries) {
      if (!have) {
        s = find_set(D.set_off, D.n_sets, e);
        have = true;
      } else {
        while (e >= __ldg(D.set_off + s + 1)) ++s;
      }
      set_id[u] = s;
      key[u] = load_key<KW>(D.set_ptr[s], e - __ldg(D.set_off + s), C);
      const unsigned long long h = hash_key(key[u]);
      if (!compact || hash_owner(h, D.n_parts) == D.part) slot[u] = key[u] ? hash_slot(h, D.cap) : D.cap;
    }
  }
#pragma unroll
  for (int u = 0; u < kDictPer; ++u)  // the first probes of a thread's entries are in flight together
    if (slot[u] != kNoId) cur[u] = __ldcg(reinterpret_cast<const uint4 *>(D.tab + slot[u]));
  // compact list: the CTA's entered occurrences take consecutive places, one reservation per CTA
  uint32_t q_next = 0;
  if (compact) {
    __shared__ uint32_t s_warp[kDictThreads / 32], s_base;
    uint32_t mine = 0;
#pragma unroll
    for (int u = 0; u < kDictPer; ++u) mine += slot[u] != kNoId ? 1u : 0u;
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += x;
    }
    if (lane == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t total = 0;
      for (int w = 0; w < kDictThreads / 32; ++w) {
        const uint32_t c = s_warp[w];
        s_warp[w] = total;
        total += c;
      }
      s_base = total ? atomicAdd(D.counters + C_MINE, total) : 0u;
    }
    __syncthreads();
    q_next = s_base + s_warp[threadIdx.x >> 5] + incl - mine;
  }
#pragma unroll
  for (int u = 0; u < kDictPer; ++u) {
    uint32_t sl = slot[u], k = 255, known = 0;
    if (sl != kNoId) {
      uint32_t seen = cur[u].z;   // the key's count when it was probed (only meaningful if the probe found the key)
      uint32_t idw = cur[u].w;    // and its id word
      if (sl != D.cap) {
        unsigned long long c = ((unsigned long long)cur[u].y << 32) | cur[u].x;
        for (uint32_t probes = 0;; ++probes) {
          if (c == key[u]) break;
          seen = 0;
          if (c == 0) {
            c = atomicCAS(&D.tab[sl].key, 0ull, key[u]);
            if (c == 0 || c == key[u]) break;
          }
          if (probes >= D.cap) {  // the table is full (a very uneven split of the key space): the caller starts over
            D.counters[C_OVERFLOW] = 1u;
            sl = kNoId;
            break;
          }
          sl = sl + 1 == D.cap ? 0 : sl + 1;
          const uint4 nx = __ldcg(reinterpret_cast<const uint4 *>(D.tab + sl));
          c = ((unsigned long long)nx.y << 32) | nx.x;
          seen = nx.z;
          idw = nx.w;
        }
      }
      if (sl != kNoId && seen <= kPostMax) k = atomicAdd(&D.tab[sl].cnt, 1u);
      else if (sl != kNoId && (idw & kHighFlag)) known = idw;   // widely shared, and already numbered
    }
    // new ids: one reservation per warp and kind (the counters are single words: a million lone atomics on one
    // address would take longer than the rest of the kernel)
#pragma unroll
    for (int kind = 0; kind < 2; ++kind) {
      const bool want = sl != kNoId && k == (kind ? kPostMax : 1u);
      const uint32_t m = __ballot_sync(0xffffffffu, want);
      if (m == 0) continue;
      uint32_t first = 0;
      const uint32_t leader = (uint32_t)(__ffs(m) - 1);
      if (lane == leader) first = atomicAdd(D.counters + (kind ? C_HIGH : C_IDS), (uint32_t)__popc(m));
      first = __shfl_sync(0xffffffffu, first, (int)leader);
      if (want) atomicMax(&D.tab[sl].id, (kind ? kHighFlag : 0u) | (first + (uint32_t)__popc(m & ((1u << lane) - 1))));
    }
    const uint32_t e = base + u * kDictThreads + threadIdx.x;
    if (known) sl = known;
    if (!compact) {
      if (e < D.n_entries) {
        D.entry[e] = sl;
        D.occ[e] = (uint8_t)min(k, 255u);
      }
    } else if (slot[u] != kNoId) {
      D.entry[q_next] = sl;   // kNoId if the table overflowed
      D.occ[q_next] = (uint8_t)min(k, 255u);
      D.set_of[q_next] = (uint16_t)set_id[u];
      ++q_next;
    }
  }
}

// R: several ranks, before the exchange: every key of the rank's own sets, compacted to 64 bits, with the global number
// of its set, grouped by the rank that owns the key (hash_owner).  kScatter == false counts, true scatters to
// out[offset[owner] + position] (positions handed out by cursor[owner]; the order inside a group does not matter).
// `region_cap` > 0 (one pass): group r has room for region_cap entries at r * region_cap; the counts keep running past
// it, which is how the caller notices an overflow and comes back with the exact two passes.
template <int KW, bool kScatter>
__global__ void __launch_bounds__(kDictThreads)
    route_owner_kernel(const __grid_constant__ DictView D, const __grid_constant__ Compact C, uint32_t world, uint32_t set_base,
                       unsigned long long *__restrict__ counts, const unsigned long long *__restrict__ offset,
                       unsigned long long *__restrict__ cursor, unsigned long long *__restrict__ out_keys,
                       uint16_t *__restrict__ out_sets, unsigned long long region_cap) {
  const uint32_t base = blockIdx.x * (uint32_t)kDictChunk;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t s = 0;
  bool have = false;
#pragma unroll
  for (int u = 0; u < kDictPer; ++u) {
    const uint32_t e = base + u * kDictThreads + threadIdx.x;
    int owner = -1;
    unsigned long long key = 0;
    if (e < D.n_entries) {
      if (!have) {
        s = find_set(D.set_off, D.n_sets, e);
        have = true;
      } else {
        while (e >= __ldg(D.set_off + s + 1)) ++s;
      }
      key = load_key<KW>(D.set_ptr[s], e - __ldg(D.set_off + s), C);
      owner = (int)hash_owner(hash_key(key), world);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, owner);
    if (owner < 0) continue;
    const uint32_t leader = (uint32_t)(__ffs(peers) - 1);
    if (!kScatter) {
      if (lane == leader) atomicAdd(counts + owner, (unsigned long long)__popc(peers));
    } else {
      unsigned long long first = 0;
      if (lane == leader) first = atomicAdd(cursor + owner, (unsigned long long)__popc(peers));
      first = __shfl_sync(peers, first, (int)leader);
      const unsigned long long pos = first + (unsigned long long)__popc(peers & ((1u << lane) - 1));
      if (region_cap && pos >= region_cap) continue;
      const unsigned long long at = (region_cap ? (unsigned long long)owner * region_cap : offset[owner]) + pos;
      out_keys[at] = key;
      out_sets[at] = (uint16_t)(set_base + s);
    }
  }
}

__global__ void route_owner_offsets_kernel(const unsigned long long *__restrict__ counts, int world,
                                           unsigned long long *__restrict__ offset, unsigned long long *__restrict__ cursor) {
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int r = 0; r < world; ++r) {
      offset[r] = run;
      cursor[r] = 0;
      run += counts[r];
    }
  }
}

// D2a: entry -> id; sizes of the posting lists and of the (range, set) groups.
__global__ void __launch_bounds__(kDictThreads) dict_ids_kernel(const __grid_constant__ DictView D,
                                                                uint32_t *__restrict__ group_cnt,
                                                                uint32_t *__restrict__ post_cnt) {
  const uint32_t base = blockIdx.x * (uint32_t)kDictChunk;
  const uint32_t lane = threadIdx.x & 31;
  const bool flat = D.flat_keys != nullptr;
  const bool compact = D.n_parts > 1 && !flat;
  const uint32_t n_occ = compact ? D.counters[C_MINE] : D.n_entries;
  if (base >= n_occ) return;
  uint32_t s = 0;
  bool have = false;
#pragma unroll
  for (int u = 0; u < kDictPer; ++u) {
    const uint32_t e = base + u * kDictThreads + threadIdx.x;
    uint32_t bin = kNoId;
    if (e < n_occ) {
      if (compact || flat) {
        s = D.set_of[e];
      } else if (!have) {
        s = find_set(D.set_off, D.n_sets, e);
        have = true;
      } else {
        while (e >= __ldg(D.set_off + s + 1)) ++s;
      }
      const uint32_t at = D.entry[e];
      Slot sl;
      sl.cnt = 0;
      if (at != kNoId && (at & kHighFlag)) {  // numbered at insertion
        sl.cnt = kPostMax + 1;
        sl.id = at;
      } else if (at != kNoId) {
        sl = D.tab[at];   // final: the insertion kernel has completed
      }
      uint32_t id = kNoId;
      if (sl.cnt >= 2) {
        if (sl.cnt <= kPostMax) {
          id = kPostFlag | sl.id;
          if (D.occ[e] == 0) post_cnt[sl.id] = sl.cnt;
        } else {
          id = sl.id & ~kHighFlag;
          bin = (id >> kRangeBits) * D.n_sets + s;
        }
      }
      D.entry[e] = id;
    }
    // one atomic per distinct group of the warp (a warp's entries mostly belong to one set and few ranges)
    const uint32_t peers = __match_any_sync(0xffffffffu, bin);
    if (bin != kNoId && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(group_cnt + bin, (uint32_t)__popc(peers));
  }
}

__global__ void __launch_bounds__(256) zero16_kernel(uint4 *__restrict__ p, const uint32_t *__restrict__ n16_ptr) {
  const uint32_t n16 = *n16_ptr;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) p[i] = make_uint4(0, 0, 0, 0);
}

// D2b: the members of the posting lists; the groups' payloads: bitmap bits (dense) or 16-bit ids in arrival order (sparse).
__global__ void __launch_bounds__(kDictThreads)
    dict_payload_kernel(const __grid_constant__ DictView D, const uint32_t *__restrict__ group_cnt,
                        const uint32_t *__restrict__ group_end, uint32_t *__restrict__ cursor, uint4 *__restrict__ payload,
                        const uint32_t *__restrict__ post_begin, uint16_t *__restrict__ postings) {
  const uint32_t base = blockIdx.x * (uint32_t)kDictChunk;
  const bool flat = D.flat_keys != nullptr;
  const bool compact = D.n_parts > 1 && !flat;
  const uint32_t n_occ = compact ? D.counters[C_MINE] : D.n_entries;
  uint32_t s = 0;
  bool have = false;
#pragma unroll
  for (int u = 0; u < kDictPer; ++u) {
    const uint32_t e = base + u * kDictThreads + threadIdx.x;
    if (e >= n_occ) continue;
    if (compact || flat) {
      s = D.set_of[e];
    } else if (!have) {
      s = find_set(D.set_off, D.n_sets, e);
      have = true;
    } else {
      while (e >= __ldg(D.set_off + s + 1)) ++s;
    }
    const uint32_t id = D.entry[e];
    if (id == kNoId) continue;
    if (id & kPostFlag) {  // the occurrences of a key fill its posting list
      postings[post_begin[id & ~kPostFlag] + D.occ[e]] = (uint16_t)s;
      continue;
    }
    const uint32_t g = (id >> kRangeBits) * D.n_sets + s, c = group_cnt[g];
    const uint32_t dmin = dense_min(D.counters[C_HIGH], id >> kRangeBits);
    const uint32_t off = group_end[g] - size16(c, dmin), low = id & (kRangeIds - 1);
    if (c >= dmin) {
      atomicOr(reinterpret_cast<uint32_t *>(payload + off) + (low >> 5), 1u << (low & 31));
    } else {
      const uint32_t pos = atomicAdd(cursor + g, 1u);
      reinterpret_cast<uint16_t *>(payload + off)[pos] = (uint16_t)low;
    }
  }
}

// X1: every posting list adds 1 to each ordered pair of its sets (rows of this call only; j > i only when the
// finalisation mirrors).
__global__ void __launch_bounds__(256)
    posting_pairs_kernel(const uint32_t *__restrict__ post_cnt, const uint32_t *__restrict__ post_begin,
                         const uint16_t *__restrict__ postings, const uint32_t *__restrict__ counters, uint32_t n_sets,
                         uint32_t row_begin, uint32_t row_end, int symmetric, int32_t *__restrict__ out) {
  const uint32_t n_ids = counters[C_IDS];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t id = blockIdx.x * blockDim.x + threadIdx.x; id < n_ids; id += stride) {
    const uint32_t c = post_cnt[id];
    if (c < 2) continue;
    const uint16_t *p = postings + post_begin[id];
    uint32_t m[kPostMax];
#pragma unroll
    for (uint32_t a = 0; a < kPostMax; ++a) m[a] = a < c ? p[a] : 0u;
#pragma unroll
    for (uint32_t a = 0; a < kPostMax; ++a) {
      if (a >= c) break;
      const uint32_t i = m[a];
      if (i < row_begin || i >= row_end) continue;
#pragma unroll
      for (uint32_t b = 0; b < kPostMax; ++b) {
        if (b >= c) break;
        const uint32_t j = m[b];
        if (j == i || (symmetric && j < i)) continue;
        atomicAdd(out + (size_t)(i - row_begin) * n_sets + j, 1);
      }
    }
  }
}

// X0: the (range, row) groups that exist, for the rows of this call.
__global__ void __launch_bounds__(256) rowlist_kernel(const uint32_t *__restrict__ group_cnt, uint32_t n_groups, uint32_t n_sets,
                                                      uint32_t row_begin, uint32_t row_end, uint32_t *__restrict__ rowlist,
                                                      uint32_t *__restrict__ counters) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups || group_cnt[g] == 0) return;
  const uint32_t s = g % n_sets;
  if (s < row_begin || s >= row_end) return;
  rowlist[atomicAdd(counters + C_ROWS, 1u)] = g;
}

struct PairsView {
  const uint32_t *group_cnt, *group_end;
  const uint4 *payload;
  const uint32_t *rowlist;
  uint32_t *counters;
  uint32_t n_sets, row_begin, n_chunks;
  int symmetric;   // rows cover all sets: only j > i is evaluated, the finalisation mirrors
  int32_t *out;    // [(row_end - row_begin) * n_sets], zero on entry
};

// X: persistent CTAs pull (range, row, column chunk) tasks.  The row's bitmap of that range sits in shared memory.
// A range that is in full use: every warp takes one column at a time, AND/popcount against a dense column, bit probes
// for a sparse one.  A range with few ids (the last one; the only one when several ranks split the key space): every
// THREAD takes a column -- the bitmaps are a few hundred bytes and a warp per column would be all overhead.
constexpr uint32_t kSmall16 = 32;  // ranges of up to 4096 ids take the thread-per-column route
__global__ void __launch_bounds__(kPairsThreads) allpairs_kernel(const __grid_constant__ PairsView V) {
  __shared__ __align__(16) uint32_t s_bits[kRangeWords];
  __shared__ uint32_t s_task;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_tasks = V.counters[C_ROWS] * V.n_chunks;
  const uint32_t n_high = V.counters[C_HIGH];
  uint4 *s_bits4 = reinterpret_cast<uint4 *>(s_bits);
  for (;;) {
    __syncthreads();  // the previous task is done with s_bits and s_task
    if (tid == 0) s_task = atomicAdd(V.counters + C_TASK, 1u);
    __syncthreads();
    const uint32_t task = s_task;
    if (task >= n_tasks) break;
    const uint32_t g = V.rowlist[task / V.n_chunks], chunk = task % V.n_chunks;
    const uint32_t t = g / V.n_sets, i = g % V.n_sets;
    // bitmap ids are dense from 0: the last range is only partly used, and nothing is set beyond its last id
    const uint32_t ids_here = min(kRangeIds, n_high - t * kRangeIds);
    const uint32_t used16 = (ids_here + 127) / 128;
    const bool small = used16 <= kSmall16;
    const uint32_t dmin = dense_min(n_high, t);
    if (small && (chunk & 3)) continue;  // thread-per-column tasks cover four chunks
    const uint32_t j0 = chunk * kPairsCols, j1 = min(V.n_sets, j0 + (small ? 4 * kPairsCols : kPairsCols));
    if (V.symmetric && j1 <= i + 1) continue;
    const uint32_t ca = V.group_cnt[g], offa = V.group_end[g] - size16(ca, dmin);
    if (ca >= dmin) {
      for (uint32_t k = tid; k < used16; k += kPairsThreads) s_bits4[k] = __ldg(V.payload + offa + k);
    } else {
      for (uint32_t k = tid; k < used16; k += kPairsThreads) s_bits4[k] = make_uint4(0, 0, 0, 0);
      __syncthreads();
      const uint16_t *la = reinterpret_cast<const uint16_t *>(V.payload + offa);
      for (uint32_t k = tid; k < ca; k += kPairsThreads) {
        const uint32_t id = la[k];
        atomicOr(s_bits + (id >> 5), 1u << (id & 31));
      }
    }
    __syncthreads();
    if (small) {
      const uint32_t j = j0 + tid;
      if (j >= j1 || j == i || (V.symmetric && j < i)) continue;
      const uint32_t gb = t * V.n_sets + j, cb = V.group_cnt[gb];
      if (cb == 0) continue;
      const uint32_t offb = V.group_end[gb] - size16(cb, dmin);
      uint32_t acc = 0;
      if (cb >= dmin) {
        const uint4 *B = V.payload + offb;
#pragma unroll 4
        for (uint32_t k = 0; k < used16; ++k) {
          const uint4 b = __ldg(B + k), a = s_bits4[k];
          acc += __popc(a.x & b.x) + __popc(a.y & b.y) + __popc(a.z & b.z) + __popc(a.w & b.w);
        }
      } else {
        const uint16_t *lb = reinterpret_cast<const uint16_t *>(V.payload + offb);
        for (uint32_t k = 0; k < cb; ++k) {
          const uint32_t id = lb[k];
          acc += (s_bits[id >> 5] >> (id & 31)) & 1u;
        }
      }
      if (acc) atomicAdd(V.out + (size_t)(i - V.row_begin) * V.n_sets + j, (int32_t)acc);
      continue;
    }
    for (uint32_t j = j0 + warp; j < j1; j += kPairsWarps) {
      if (j == i || (V.symmetric && j < i)) continue;
      const uint32_t gb = t * V.n_sets + j, cb = V.group_cnt[gb];
      if (cb == 0) continue;
      const uint32_t offb = V.group_end[gb] - size16(cb, dmin);
      uint32_t acc = 0;
      if (cb >= dmin) {
        const uint4 *B = V.payload + offb;
#pragma unroll 4
        for (uint32_t k = lane; k < used16; k += 32) {
          const uint4 b = __ldg(B + k), a = s_bits4[k];
          acc += __popc(a.x & b.x) + __popc(a.y & b.y) + __popc(a.z & b.z) + __popc(a.w & b.w);
        }
      } else {
        const uint16_t *lb = reinterpret_cast<const uint16_t *>(V.payload + offb);
        for (uint32_t k = lane; k < cb; k += 32) {
          const uint32_t id = lb[k];
          acc += (s_bits[id >> 5] >> (id & 31)) & 1u;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
      if (lane == 0 && acc) atomicAdd(V.out + (size_t)(i - V.row_begin) * V.n_sets + j, (int32_t)acc);
    }
  }
}

// F: full rows (mirror + diagonal) and ANI.  containment(I, |first|) = I == 0 ? 0 : I / |first|
// (src/ani_estimation.cpp:24-28); binomial_estimator(c, weight) = c <= 0 ? 0 : pow(c, 1.0 / weight) (:38-42).
__global__ void __launch_bounds__(256)
    finalize_kernel(const int32_t *__restrict__ raw, const int32_t *__restrict__ sizes, uint32_t n_sets, uint32_t row_begin,
                    uint32_t n_rows, int symmetric, double inv_weight, int32_t *__restrict__ out_counts,
                    double *__restrict__ out_ani) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_rows * n_sets) return;
  const uint32_t r = (uint32_t)(idx / n_sets), j = (uint32_t)(idx % n_sets), i = row_begin + r;
  int32_t c;
  if (i == j) c = sizes[i];
  else if (symmetric && j < i) c = raw[(size_t)(j - row_begin) * n_sets + i];
  else c = raw[idx];
  if (out_counts) out_counts[idx] = c;
  if (out_ani) {
    double a = 0.0;
    if (c != 0) {
      const double cont = (double)c / (double)sizes[i];
      a = cont <= 0.0 ? 0.0 : pow(cont, inv_weight);
    }
    out_ani[idx] = a;
  }
}

bool make_compact(const uint64_t mask[2], Compact *out) {  // false: more than 64 mask bits or 32 runs
  Compact c = {};
  int dst = 0;
  auto bit = [&](int b) { return (mask[b >> 6] >> (b & 63)) & 1; };
  for (int b = 0; b < 128;) {
    if (!bit(b)) {
      ++b;
      continue;
    }
    int e = b;
    while (e < 128 && bit(e) && (e >> 6) == (b >> 6) ) ++e;  // a run stays inside one 64-bit word of the source
    if (c.n_runs == 32 || dst + (e - b) > 64) return false;
    c.src[c.n_runs] = (uint8_t)b;
    c.len[c.n_runs] = (uint8_t)(e - b);
    c.dst[c.n_runs] = (uint8_t)dst;
    dst += e - b;
    ++c.n_runs;
    b = e;
  }
  *out = c;
  return true;
}

}  // namespace

bool all_pairs_dict_eligible(sks_set *const *sets, int64_t n) {
  static const bool enabled = getenv("SKS_DICT_INTERSECT") ? atoi(getenv("SKS_DICT_INTERSECT")) != 0 : true;
  if (!enabled || n < 2 || n > 65536) return false;
  const sks_set *ref = sets[0];
  if (!ref || ref->repr != SKS_REPR_SORTED) return false;
  uint64_t total = 0;
  for (int64_t i = 0; i < n; ++i) {
    const sks_set *s = sets[i];
    if (!s || s->repr != SKS_REPR_SORTED || s->key_words != ref->key_words || s->mask[0] != ref->mask[0] ||
        s->mask[1] != ref->mask[1] || s->device != ref->device || s->count < 0)
      return false;
    total += (uint64_t)s->count;
  }
  if (total >= (1ull << 30)) return false;
  if (ref->key_words == 2) {
    Compact c;
    if (!make_compact(ref->mask, &c)) return false;
  }
  return true;
}

// Raw counts of the rows [row_begin, row_end) of the n x n matrix of |sets[i] n sets[j]| (off-diagonal entries; with
// `symmetric` -- which needs all rows -- only j > i), from the keys of share `part` of `n_parts` of the key space:
// the shares of all parts add up to the counts.  *raw receives raw_rows x n int32 (rows beyond the range stay zero),
// *sizes the n set sizes.
// *raw_out may come in holding the matrix of earlier calls (other shares of the key space): the counts are added to
// it.  *h_overflow points at a pinned word that is non-zero, once the stream has been synchronised, if the table was
// too small (then the counts are incomplete and the caller takes another route).
// With `flat` the occurrences do not come from `sets` (unused then) but as arrays: the keys this rank owns, routed here
// by all ranks (all_pairs_route + an all-to-all), each with the global number of its set.
int all_pairs_raw(sks_ctx *ctx, sks_set *const *sets, int64_t n, int64_t row_begin, int64_t row_end, int part, int n_parts,
                  bool symmetric, int64_t raw_rows, BufferRef *raw_out, BufferRef *sizes_out, const uint32_t **h_overflow,
                  const FlatKeys *flat) {
  const uint32_t n_sets = (uint32_t)n;
  if (symmetric && !(row_begin == 0 && row_end == n)) return set_error(SKS_ERR_INVALID, "mirroring needs all rows");
  if (raw_rows < row_end - row_begin) return set_error(SKS_ERR_INVALID, "raw matrix too small");
  const int kw = flat ? 1 : sets[0]->key_words;
  Compact compact = {};
  if (kw == 2 && !make_compact(sets[0]->mask, &compact)) return set_error(SKS_ERR_INVALID, "mask too wide for the dictionary");
  uint64_t total = 0;
  if (flat) {
    total = flat->n;
    part = 0;
    n_parts = 1;
  } else {
    for (int64_t i = 0; i < n; ++i) total += (uint64_t)sets[i]->count;
  }
  if (total >= (1ull << 30)) return set_error(SKS_ERR_CAPACITY, "too many keys for the dictionary");   // slots stay below 2^31
  const uint32_t K = (uint32_t)total;
  // this rank enters about K / n_parts keys (the hash spreads the DISTINCT keys evenly; a widely shared key brings all
  // its occurrences to one rank, but they take one slot); 1.5 slots per entered key, and room for an uneven split
  const uint32_t n_parts_u = (uint32_t)std::max(n_parts, 1);
  const uint32_t k_here = n_parts_u == 1 ? K : (uint32_t)std::min<uint64_t>(K, (uint64_t)K / n_parts_u + K / (4 * n_parts_u) + 65536);
  const uint32_t cap = std::max<uint32_t>(1024u, k_here + k_here / 2);
  // at most K / 2 keys can occur twice
  const uint32_t max_ids = K / 2 + 1;
  const uint32_t n_ranges = K / (kPostMax + 1) / kRangeIds + 1;   // bitmap ids: keys held by more than kPostMax sets
  const uint64_t n_groups64 = (uint64_t)n_ranges * n_sets;
  if (n_groups64 >= (1ull << 28)) return set_error(SKS_ERR_CAPACITY, "too many (range, set) groups for the dictionary");
  const uint32_t n_groups = (uint32_t)n_groups64;
  const uint32_t n_chunks = (n_sets + kPairsCols - 1) / kPairsCols;
  if (n_groups64 * n_chunks >= (1ull << 32)) return set_error(SKS_ERR_CAPACITY, "too many tasks for the dictionary");

  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  // one control block: counters | set_off | set_ptr | sizes | group_cnt | cursor | group_end | rowlist
  const size_t sz_cnt = align(4 * C_COUNT), sz_off = align(4 * ((size_t)n_sets + 1)), sz_ptr = align(8 * (size_t)n_sets);
  const size_t sz_sizes = align(4 * (size_t)n_sets), sz_grp = align(4 * (size_t)n_groups);
  BufferRef ctl, tab, entry, occ, payload, raw, post;
  SKS_TRY(alloc_buffer(ctx, sz_cnt + sz_off + sz_ptr + sz_sizes + 4 * sz_grp, &ctl));
  char *cb = static_cast<char *>(ctl->ptr);
  uint32_t *d_counters = reinterpret_cast<uint32_t *>(cb);
  uint32_t *d_off = reinterpret_cast<uint32_t *>(cb + sz_cnt);
  const void **d_ptr = reinterpret_cast<const void **>(cb + sz_cnt + sz_off);
  int32_t *d_sizes = reinterpret_cast<int32_t *>(cb + sz_cnt + sz_off + sz_ptr);
  uint32_t *d_gcnt = reinterpret_cast<uint32_t *>(cb + sz_cnt + sz_off + sz_ptr + sz_sizes);
  uint32_t *d_cursor = d_gcnt + sz_grp / 4, *d_gend = d_cursor + sz_grp / 4, *d_rowlist = d_gend + sz_grp / 4;
  SKS_TRY(alloc_buffer(ctx, sizeof(Slot) * ((size_t)cap + 1), &tab));
  SKS_TRY(alloc_buffer(ctx, 4 * (size_t)std::max<uint32_t>(K, 1), &entry));
  SKS_TRY(alloc_buffer(ctx, (size_t)std::max<uint32_t>(K, 16) * (n_parts > 1 ? 3 : 1), &occ));   // occ | set_of
  // posting lists: sizes and starts per id, members (every entry belongs to at most one list)
  const size_t sz_ids = align(4 * (size_t)max_ids);
  SKS_TRY(alloc_buffer(ctx, 2 * sz_ids + 2 * (size_t)K + 16, &post));
  uint32_t *d_pcnt = static_cast<uint32_t *>(post->ptr), *d_pbegin = d_pcnt + sz_ids / 4;
  uint16_t *d_postings = reinterpret_cast<uint16_t *>(static_cast<char *>(post->ptr) + 2 * sz_ids);
  // payload: a dense group holds >= kDenseMin ids in 8 KB, a sparse one 2 bytes per id rounded up to 16
  // (a bitmap group holds at least 64 ids, and there are at most n_groups groups)
  const size_t payload16 = std::min<size_t>((size_t)n_groups * kRange16, (size_t)K / 64 * kRange16 + n_groups) + (size_t)K / 8 +
                           std::min<size_t>(n_groups, K) + 16;
  SKS_TRY(alloc_buffer(ctx, payload16 * 16, &payload));
  const bool accumulate = raw_out && *raw_out && (*raw_out)->bytes >= 4 * (size_t)raw_rows * n_sets;
  if (accumulate) raw = *raw_out;
  else SKS_TRY(alloc_buffer(ctx, 4 * (size_t)raw_rows * n_sets, &raw));

  // host tables through the pinned ring: offsets | pointers | sizes
  char *stage = nullptr;
  SKS_TRY(ctx_pinned(ctx, sz_off + sz_ptr + sz_sizes, reinterpret_cast<void **>(&stage)));
  uint32_t *h_off = reinterpret_cast<uint32_t *>(stage);
  const void **h_ptr = reinterpret_cast<const void **>(stage + sz_off);
  int32_t *h_sizes = reinterpret_cast<int32_t *>(stage + sz_off + sz_ptr);
  uint32_t at = 0;
  for (uint32_t s = 0; s < n_sets; ++s) {
    h_off[s] = at;
    h_ptr[s] = flat ? nullptr : static_cast<const char *>(sets[s]->buf->ptr) + sets[s]->byte_off;
    h_sizes[s] = flat ? flat->h_sizes[s] : (int32_t)sets[s]->count;
    at += flat ? 0u : (uint32_t)sets[s]->count;
  }
  h_off[n_sets] = at;
  SKS_CUDA_TRY(cudaMemcpyAsync(d_off, stage, sz_off + sz_ptr + sz_sizes, cudaMemcpyHostToDevice, ctx->stream));
  SKS_CUDA_TRY(cudaMemsetAsync(d_counters, 0, sz_cnt, ctx->stream));
  SKS_CUDA_TRY(cudaMemsetAsync(d_gcnt, 0, 2 * sz_grp, ctx->stream));  // group_cnt and cursor
  SKS_CUDA_TRY(cudaMemsetAsync(tab->ptr, 0, sizeof(Slot) * ((size_t)cap + 1), ctx->stream));
  if (!accumulate) SKS_CUDA_TRY(cudaMemsetAsync(raw->ptr, 0, 4 * (size_t)raw_rows * n_sets, ctx->stream));
  SKS_CUDA_TRY(cudaMemsetAsync(d_pcnt, 0, sz_ids, ctx->stream));

  DictView D;
  D.set_ptr = d_ptr;
  D.set_off = d_off;
  D.n_sets = n_sets;
  D.n_entries = K;
  D.tab = static_cast<Slot *>(tab->ptr);
  D.cap = cap;
  D.entry = static_cast<uint32_t *>(entry->ptr);
  D.occ = static_cast<uint8_t *>(occ->ptr);
  D.set_of = reinterpret_cast<uint16_t *>(static_cast<uint8_t *>(occ->ptr) + (((size_t)std::max<uint32_t>(K, 16) + 1) & ~(size_t)1));
  D.counters = d_counters;
  D.part = (uint32_t)part;
  D.n_parts = (uint32_t)std::max(n_parts, 1);
  D.flat_keys = flat ? flat->keys : nullptr;
  if (flat) D.set_of = const_cast<uint16_t *>(flat->sets);
  const unsigned entry_grid = (K + kDictChunk - 1) / kDictChunk;
  const unsigned wide_grid = (unsigned)ctx->sm_count * 8;
  {
    KernelTimer timer(ctx, SKS_KERNEL_DICT);
    if (K > 0) {
      if (kw == 1) dict_insert_kernel<1><<<entry_grid, kDictThreads, 0, ctx->stream>>>(D, compact);
      else dict_insert_kernel<2><<<entry_grid, kDictThreads, 0, ctx->stream>>>(D, compact);
      dict_ids_kernel<<<entry_grid, kDictThreads, 0, ctx->stream>>>(D, d_gcnt, d_pcnt);
      ctx->launches += 2;
    }
    // group_end = inclusive prefix sum of the groups' payload sizes; post_begin = exclusive prefix sum of the list sizes
    size_t temp_bytes = 0, temp2 = 0;
    cub::CountingInputIterator<uint32_t> group_ids(0);
    cub::TransformInputIterator<uint32_t, GroupSize16, cub::CountingInputIterator<uint32_t>> sizes_in(
        group_ids, GroupSize16{d_gcnt, d_counters, n_sets});
    cub::DeviceScan::InclusiveSum(nullptr, temp_bytes, sizes_in, d_gend, (int)n_groups, ctx->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, temp2, d_pcnt, d_pbegin, (int)max_ids, ctx->stream);
    temp_bytes = std::max(temp_bytes, temp2);
    void *temp = nullptr;
    SKS_TRY(ctx_scratch(ctx, temp_bytes, &temp));
    cub::DeviceScan::InclusiveSum(temp, temp_bytes, sizes_in, d_gend, (int)n_groups, ctx->stream);
    cub::DeviceScan::ExclusiveSum(temp, temp_bytes, d_pcnt, d_pbegin, (int)max_ids, ctx->stream);
    zero16_kernel<<<wide_grid, 256, 0, ctx->stream>>>(static_cast<uint4 *>(payload->ptr), d_gend + n_groups - 1);
    ctx->launches += 3;
    if (K > 0) {
      dict_payload_kernel<<<entry_grid, kDictThreads, 0, ctx->stream>>>(D, d_gcnt, d_gend, d_cursor,
                                                                        static_cast<uint4 *>(payload->ptr), d_pbegin, d_postings);
      ctx->launches++;
    }
    SKS_CUDA_TRY(cudaGetLastError());
  }
  {
    KernelTimer timer(ctx, SKS_KERNEL_ALLPAIRS);
    if (K > 0) {
      posting_pairs_kernel<<<wide_grid, 256, 0, ctx->stream>>>(d_pcnt, d_pbegin, d_postings, d_counters, n_sets, (uint32_t)row_begin,
                                                               (uint32_t)row_end, symmetric ? 1 : 0, static_cast<int32_t *>(raw->ptr));
      ctx->launches++;
    }
    rowlist_kernel<<<(n_groups + 255) / 256, 256, 0, ctx->stream>>>(d_gcnt, n_groups, n_sets, (uint32_t)row_begin,
                                                                    (uint32_t)row_end, d_rowlist, d_counters);
    PairsView V;
    V.group_cnt = d_gcnt;
    V.group_end = d_gend;
    V.payload = static_cast<const uint4 *>(payload->ptr);
    V.rowlist = d_rowlist;
    V.counters = d_counters;
    V.n_sets = n_sets;
    V.row_begin = (uint32_t)row_begin;
    V.n_chunks = n_chunks;
    V.symmetric = symmetric ? 1 : 0;
    V.out = static_cast<int32_t *>(raw->ptr);
    allpairs_kernel<<<wide_grid, kPairsThreads, 0, ctx->stream>>>(V);
    ctx->launches += 2;
    SKS_CUDA_TRY(cudaGetLastError());
  }
  if (sizes_out) {  // [n] int32 set sizes, for the finalisation
    SKS_TRY(alloc_buffer(ctx, 4 * (size_t)std::max<uint32_t>(n_sets, 4), sizes_out));
    SKS_CUDA_TRY(cudaMemcpyAsync((*sizes_out)->ptr, d_sizes, 4 * (size_t)n_sets, cudaMemcpyDeviceToDevice, ctx->stream));
  }
  if (h_overflow) {
    uint32_t *h = nullptr;
    SKS_TRY(ctx_pinned(ctx, 64, reinterpret_cast<void **>(&h)));
    *h = 0;
    SKS_CUDA_TRY(cudaMemcpyAsync(h, d_counters + C_OVERFLOW, 4, cudaMemcpyDeviceToHost, ctx->stream));
    *h_overflow = h;
  }
  *raw_out = raw;
  return SKS_OK;
}

// R (several ranks): the keys of this rank's sets (compacted to 64 bits) and the global numbers of their sets, grouped
// by owning rank: *out_keys / *out_sets hold all of them, group r at offset sum(counts[0..r)); d_counts = `world`
// 64-bit counts on the device (for the ranks' exchange of counts).
// region_cap > 0: one pass into fixed regions of that many entries per owner (group r at r * region_cap; counts beyond
// the capacity mean the pass has to be repeated exactly); region_cap == 0: count, then scatter (two passes).
size_t all_pairs_route_cap(uint64_t n_keys, int world) {
  static const bool tight = getenv("SKS_ROUTE_TIGHT") != nullptr;  // tests: regions that are sure to overflow
  if (tight) return (size_t)(n_keys / (2 * (uint64_t)world) + 1);
  return (size_t)(n_keys / world + n_keys / (4 * (uint64_t)world) + 4096);
}
int all_pairs_route(sks_ctx *ctx, sks_set *const *sets, int64_t n_local, int64_t set_base, int world, size_t region_cap,
                    BufferRef *out_keys, BufferRef *out_sets, BufferRef *ctl_out, unsigned long long **d_counts) {
  uint64_t total = 0;
  for (int64_t i = 0; i < n_local; ++i) total += (uint64_t)sets[i]->count;
  if (total >= (1ull << 31)) return set_error(SKS_ERR_CAPACITY, "too many keys for the dictionary");
  const uint32_t K = (uint32_t)total, n_sets = (uint32_t)n_local;
  const int kw = n_local ? sets[0]->key_words : 1;
  Compact compact = {};
  if (kw == 2 && !make_compact(sets[0]->mask, &compact)) return set_error(SKS_ERR_INVALID, "mask too wide for the dictionary");
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t sz_w = align(8 * (size_t)world), sz_off = align(4 * ((size_t)n_sets + 1)), sz_ptr = align(8 * (size_t)std::max<uint32_t>(n_sets, 1));
  SKS_TRY(alloc_buffer(ctx, 3 * sz_w + sz_off + sz_ptr, ctl_out));
  char *cb = static_cast<char *>((*ctl_out)->ptr);
  unsigned long long *d_cnt = reinterpret_cast<unsigned long long *>(cb), *d_offs = d_cnt + sz_w / 8, *d_cur = d_offs + sz_w / 8;
  uint32_t *d_off = reinterpret_cast<uint32_t *>(cb + 3 * sz_w);
  const void **d_ptr = reinterpret_cast<const void **>(cb + 3 * sz_w + sz_off);
  const size_t slots = region_cap ? region_cap * (size_t)world : (size_t)K;
  SKS_TRY(alloc_buffer(ctx, 8 * std::max<size_t>(slots, 2), out_keys));
  SKS_TRY(alloc_buffer(ctx, 2 * std::max<size_t>(slots, 8), out_sets));
  char *stage = nullptr;
  SKS_TRY(ctx_pinned(ctx, sz_off + sz_ptr, reinterpret_cast<void **>(&stage)));
  uint32_t *h_off = reinterpret_cast<uint32_t *>(stage);
  const void **h_ptr = reinterpret_cast<const void **>(stage + sz_off);
  uint32_t at = 0;
  for (uint32_t s = 0; s < n_sets; ++s) {
    h_off[s] = at;
    h_ptr[s] = static_cast<const char *>(sets[s]->buf->ptr) + sets[s]->byte_off;
    at += (uint32_t)sets[s]->count;
  }
  h_off[n_sets] = at;
  SKS_CUDA_TRY(cudaMemcpyAsync(d_off, stage, sz_off + sz_ptr, cudaMemcpyHostToDevice, ctx->stream));
  SKS_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, sz_w, ctx->stream));
  *d_counts = d_cnt;
  if (K == 0) return SKS_OK;
  DictView D = {};
  D.set_ptr = d_ptr;
  D.set_off = d_off;
  D.n_sets = n_sets;
  D.n_entries = K;
  const unsigned grid = (K + kDictChunk - 1) / kDictChunk;
  unsigned long long *ok = static_cast<unsigned long long *>((*out_keys)->ptr);
  uint16_t *os = static_cast<uint16_t *>((*out_sets)->ptr);
  KernelTimer timer(ctx, SKS_KERNEL_DICT);
  const uint32_t w = (uint32_t)world, sb = (uint32_t)set_base;
  if (region_cap) {  // one pass: the cursors are the counts
    if (kw == 1) route_owner_kernel<1, true><<<grid, kDictThreads, 0, ctx->stream>>>(D, compact, w, sb, nullptr, nullptr, d_cnt, ok, os, region_cap);
    else route_owner_kernel<2, true><<<grid, kDictThreads, 0, ctx->stream>>>(D, compact, w, sb, nullptr, nullptr, d_cnt, ok, os, region_cap);
    ctx->launches += 1;
  } else if (kw == 1) {
    route_owner_kernel<1, false><<<grid, kDictThreads, 0, ctx->stream>>>(D, compact, w, sb, d_cnt, nullptr, nullptr, nullptr, nullptr, 0);
    route_owner_offsets_kernel<<<1, 32, 0, ctx->stream>>>(d_cnt, world, d_offs, d_cur);
    route_owner_kernel<1, true><<<grid, kDictThreads, 0, ctx->stream>>>(D, compact, w, sb, nullptr, d_offs, d_cur, ok, os, 0);
    ctx->launches += 3;
  } else {
    route_owner_kernel<2, false><<<grid, kDictThreads, 0, ctx->stream>>>(D, compact, w, sb, d_cnt, nullptr, nullptr, nullptr, nullptr, 0);
    route_owner_offsets_kernel<<<1, 32, 0, ctx->stream>>>(d_cnt, world, d_offs, d_cur);
    route_owner_kernel<2, true><<<grid, kDictThreads, 0, ctx->stream>>>(D, compact, w, sb, nullptr, d_offs, d_cur, ok, os, 0);
    ctx->launches += 3;
  }
  SKS_CUDA_TRY(cudaGetLastError());
  return SKS_OK;
}

// Can the sets of a sharded run take the dictionary (decided from what every rank knows after the header exchange)?
bool all_pairs_dict_usable(int key_words, const uint64_t mask[2], int64_t n_total, uint64_t total_keys) {
  static const bool enabled = getenv("SKS_DICT_INTERSECT") ? atoi(getenv("SKS_DICT_INTERSECT")) != 0 : true;
  if (!enabled || n_total < 2 || n_total > 65536 || total_keys >= (1ull << 30)) return false;
  if (key_words == 2) {
    Compact c;
    if (!make_compact(mask, &c)) return false;
  }
  return key_words == 1 || key_words == 2;
}

// F: from raw off-diagonal counts of the rows [row_begin, row_begin + n_rows) to full rows (mirror, diagonal) and ANI.
int all_pairs_finalize(sks_ctx *ctx, const int32_t *raw_rows, const int32_t *d_sizes, int64_t n, int64_t row_begin,
                       int64_t n_rows, bool symmetric, int weight, BufferRef *counts, BufferRef *ani) {
  SKS_TRY(alloc_buffer(ctx, 4 * (size_t)n_rows * n, counts));
  if (ani) SKS_TRY(alloc_buffer(ctx, 8 * (size_t)n_rows * n, ani));
  KernelTimer timer(ctx, SKS_KERNEL_ANI);
  const size_t cells = (size_t)n_rows * n;
  if (cells > 0) {
    finalize_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, ctx->stream>>>(
        raw_rows, d_sizes, (uint32_t)n, (uint32_t)row_begin, (uint32_t)n_rows, symmetric ? 1 : 0, 1.0 / (double)weight,
        static_cast<int32_t *>((*counts)->ptr), ani ? static_cast<double *>((*ani)->ptr) : nullptr);
    ctx->launches++;
  }
  SKS_CUDA_TRY(cudaGetLastError());
  return SKS_OK;
}

// Rows [row_begin, row_end) of the n x n matrix of |sets[i] n sets[j]|, device resident: *counts receives
// (row_end - row_begin) * n int32 (full rows incl. the diagonal), *ani (optional) as many doubles.
int all_pairs_dict(sks_ctx *ctx, sks_set *const *sets, int64_t n, int64_t row_begin, int64_t row_end, BufferRef *counts,
                   BufferRef *ani, BufferRef *sizes_out, const uint32_t **h_overflow) {
  BufferRef raw, sizes;
  // SKS_DICT_PARTS=p (tests, profiling): the key space is entered in p shares one after the other, as p ranks would
  static const int parts = [] { const char *e = getenv("SKS_DICT_PARTS"); return e ? std::max(1, atoi(e)) : 1; }();
  const bool symmetric = row_begin == 0 && row_end == n && parts == 1;
  for (int p = 0; p < parts; ++p)
    SKS_TRY(all_pairs_raw(ctx, sets, n, row_begin, row_end, p, parts, symmetric, row_end - row_begin, &raw, p == 0 ? &sizes : nullptr,
                          p == parts - 1 ? h_overflow : nullptr, nullptr));
  SKS_TRY(all_pairs_finalize(ctx, static_cast<const int32_t *>(raw->ptr), static_cast<const int32_t *>(sizes->ptr), n, row_begin,
                             row_end - row_begin, symmetric, sets[0]->weight, counts, ani));
  if (sizes_out) *sizes_out = sizes;
  return SKS_OK;
}

}  // namespace sks
