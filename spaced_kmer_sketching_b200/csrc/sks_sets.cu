// Set-side kernels: presence-bitset clear / popcount / AND-popcount (K4 clear, K5 bitset),
// sort + unique of sketched keys (the sorted-keys representation of kmer_set), and the
// sorted-set intersection count (K5 sorted).
//
// Reference semantics reproduced:
//   kmer_set::insert_kmers   src/kmer.hpp:170-178   (set of distinct (masked_bits, mask))
//   kmer_set::kmer_set_size  src/kmer.hpp:186-189
//   kmer_set_intersection    src/kmer_set.cpp:23-41 (|A n B|)
#include <algorithm>

#include <cub/cub.cuh>

#include "sks_internal.cuh"

namespace sks {
namespace {

constexpr int kStreamThreads = 256;

// ---- streaming clear ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStreamThreads) fill_zero_kernel(uint4 *__restrict__ p, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 z = make_uint4(0, 0, 0, 0);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    p[i] = z;
    p[i + stride] = z;
    p[i + 2 * stride] = z;
    p[i + 3 * stride] = z;
  }
  for (; i < n16; i += stride) p[i] = z;
}

__device__ __forceinline__ uint32_t popc4(uint4 v) { return __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }

__device__ __forceinline__ void block_add3(unsigned long long a, unsigned long long b, unsigned long long c,
                                           unsigned long long *out, int n_out) {
  __shared__ unsigned long long s[3][kStreamThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_down_sync(0xffffffffu, a, o);
    b += __shfl_down_sync(0xffffffffu, b, o);
    c += __shfl_down_sync(0xffffffffu, c, o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    s[0][wid] = a;
    s[1][wid] = b;
    s[2][wid] = c;
  }
  __syncthreads();
  if (threadIdx.x < 3 && threadIdx.x < n_out) {
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < kStreamThreads / 32; ++i) t += s[threadIdx.x][i];
    if (t) atomicAdd(out + threadIdx.x, t);
  }
}

// |A|, |B|, |A n B| of two presence bitsets in ONE pass over both (2 * n16 * 16 bytes of traffic).
__global__ void __launch_bounds__(kStreamThreads)
    bitset_pair_counts_kernel(const uint4 *__restrict__ a, const uint4 *__restrict__ b, size_t n16,
                              unsigned long long *__restrict__ out3) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t ca = 0, cb = 0, ci = 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    uint4 va[4], vb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      va[u] = __ldcs(a + i + u * stride);
      vb[u] = __ldcs(b + i + u * stride);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ca += popc4(va[u]);
      cb += popc4(vb[u]);
      ci += popc4(make_uint4(va[u].x & vb[u].x, va[u].y & vb[u].y, va[u].z & vb[u].z, va[u].w & vb[u].w));
    }
  }
  for (; i < n16; i += stride) {
    const uint4 va = __ldcs(a + i), vb = __ldcs(b + i);
    ca += popc4(va);
    cb += popc4(vb);
    ci += popc4(make_uint4(va.x & vb.x, va.y & vb.y, va.z & vb.z, va.w & vb.w));
  }
  block_add3(ca, cb, ci, out3, 3);
}

__global__ void __launch_bounds__(kStreamThreads)
    bitset_popcount_kernel(const uint4 *__restrict__ a, size_t n16, unsigned long long *__restrict__ out1) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t ca = 0;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    uint4 va[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) va[u] = __ldcs(a + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) ca += popc4(va[u]);
  }
  for (; i < n16; i += stride) ca += popc4(__ldcs(a + i));
  block_add3(ca, 0, 0, out1, 1);
}

// Small bitsets (fewer than 4 words, e.g. weight 1..3): scalar word loop.
__global__ void bitset_small_counts_kernel(const uint32_t *a, const uint32_t *b, uint64_t n_words,
                                           unsigned long long *out3) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long ca = 0, cb = 0, ci = 0;
    for (uint64_t i = 0; i < n_words; ++i) {
      const uint32_t x = a[i], y = b ? b[i] : 0u;
      ca += __popc(x);
      cb += __popc(y);
      ci += __popc(x & y);
    }
    out3[0] += ca;
    if (b) {
      out3[1] += cb;
      out3[2] += ci;
    }
  }
}

int stream_grid(const sks_ctx *ctx, size_t n16) {
  size_t want = (n16 + kStreamThreads - 1) / kStreamThreads;
  size_t cap = (size_t)ctx->sm_count * 8;  // 8 resident CTAs of 256 threads per SM
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

// ---- bucketed bitset build --------------------------------------------------------------------------
// Random atomicOr into a 512 MiB bitset costs a 64 B DRAM read + a 32 B write per k-mer (measured:
// 1 GB of DRAM traffic for 10 M k-mers at 29 % of HBM, latency bound).  Instead:
//   1. the 32-bit PEXT indices of a genome are partitioned into <= 1024 buckets by their top bits (a
//      counting pass, a scan and a scatter pass over an L2-resident 20 MB list); a bucket covers a GROUP
//      of kGroupSlices consecutive 64 KB slices of the bitset;
//   2. a CTA takes one bucket, copies its indices into shared memory, and assembles each of the group's
//      slices there before streaming it to HBM exactly once.
// The clear and the insert become ONE sequential write pass, and the slice popcounts give
// kmer_set_size() for free.  (Measured alternatives: radix sort by slice with cub, 205 us of sort; one
// bucket per slice with a per-index global cursor bump, 276 us of scatter; 256 coarse buckets with the
// group's indices filtered out of the bucket by every CTA, 265 us of build of which ~60 us was the filter.)
constexpr int kBuildThreads = 512;
constexpr int kKeyCap = 11264;                       // indices of one bucket kept in shared memory (44 KB)

// 1-D bulk store shared -> global (TMA engine), tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct BuildGenome {
  const uint32_t *raw;      // PEXT indices as emitted by the sketch kernel
  uint32_t *bucketed;       // the same, grouped by bucket
  uint32_t *starts;         // [n_parts + 1] first position of every bucket in `bucketed`
  uint32_t *cursor;         // [n_parts] histogram, then scatter cursors (= bucket ends afterwards)
  uint32_t *bitset;
  unsigned long long *set_count;
  uint32_t n;               // number of indices
  uint32_t cap;             // > 0: the buckets are fixed regions of `cap` slots (filled by the sketch kernel)
};

__global__ void __launch_bounds__(256)
    part_hist_kernel(const BuildGenome *__restrict__ genomes, int part_shift, uint32_t n_parts) {
  __shared__ uint32_t s_hist[kMaxParts];
  const BuildGenome g = genomes[blockIdx.y];
  for (uint32_t i = threadIdx.x; i < n_parts; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += stride)
    atomicAdd(&s_hist[__ldg(g.raw + i) >> part_shift], 1u);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n_parts; i += blockDim.x)
    if (s_hist[i]) atomicAdd(g.cursor + i, s_hist[i]);  // cursor doubles as the histogram until the scan
}

// One 1024-thread CTA per genome: exclusive scan of the histogram -> starts[0..n_parts], cursor = starts.
__global__ void __launch_bounds__(kMaxParts) part_scan_kernel(const BuildGenome *__restrict__ genomes, uint32_t n_parts) {
  __shared__ uint32_t s_warp[32];
  const BuildGenome g = genomes[blockIdx.x];
  const uint32_t t = threadIdx.x, lane = t & 31;
  const uint32_t v = t < n_parts ? g.cursor[t] : 0u;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += x;
  }
  if (lane == 31) s_warp[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    uint32_t w = s_warp[t];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, w, o);
      if (t >= (uint32_t)o) w += x;
    }
    s_warp[t] = w;
  }
  __syncthreads();
  const uint32_t excl = incl - v + ((t >> 5) ? s_warp[(t >> 5) - 1] : 0u);
  if (t < n_parts) {
    g.starts[t] = excl;
    g.cursor[t] = excl;
  }
  if (t == 0) g.starts[n_parts] = g.n;
}

constexpr int kScatterKeys = 4096;  // indices per CTA pass

__global__ void __launch_bounds__(256)
    part_scatter_kernel(const BuildGenome *__restrict__ genomes, int part_shift, uint32_t n_parts) {
  __shared__ uint32_t s_hist[kMaxParts];
  __shared__ uint32_t s_base[kMaxParts];
  const BuildGenome g = genomes[blockIdx.y];
  const uint32_t n_chunks = (g.n + kScatterKeys - 1) / kScatterKeys;
  for (uint32_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    for (uint32_t i = threadIdx.x; i < n_parts; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    uint32_t key[kScatterKeys / 256], rank[kScatterKeys / 256];
    const uint32_t base = chunk * kScatterKeys;
#pragma unroll
    for (int u = 0; u < kScatterKeys / 256; ++u) {
      const uint32_t i = base + u * 256 + threadIdx.x;
      if (i < g.n) key[u] = __ldg(g.raw + i);
    }
#pragma unroll
    for (int u = 0; u < kScatterKeys / 256; ++u) {
      const uint32_t i = base + u * 256 + threadIdx.x;
      if (i < g.n) rank[u] = atomicAdd(&s_hist[key[u] >> part_shift], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_parts; i += blockDim.x)
      s_base[i] = s_hist[i] ? atomicAdd(g.cursor + i, s_hist[i]) : 0u;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kScatterKeys / 256; ++u) {
      const uint32_t i = base + u * 256 + threadIdx.x;
      if (i < g.n) g.bucketed[s_base[key[u] >> part_shift] + rank[u]] = key[u];
    }
    __syncthreads();
  }
}

// Work item = (genome, bucket), handed out by an atomic counter.  Two ~108 KB CTAs per SM: one assembles
// while the other streams its slice out.  A slice leaves shared memory through registers
// (LDS.128 -> popcount -> STG.128, the buffer zeroed behind the read): plain vector stores are
// fire-and-forget, whereas a 64 KB bulk (TMA) store per slice measured ~6 us of latency with one store in
// flight per CTA.
__global__ void __launch_bounds__(kBuildThreads, 2)
    bitset_build_kernel(const BuildGenome *__restrict__ genomes, uint32_t n_genomes, uint32_t n_parts,
                        uint32_t group_slices, unsigned int *__restrict__ work_counter) {
  extern __shared__ __align__(128) uint32_t s_dyn[];
  uint32_t *s_slice = s_dyn;                 // [kSliceWords]
  uint32_t *s_keys = s_dyn + kSliceWords;    // [kKeyCap]
  __shared__ uint32_t s_item;
  __shared__ unsigned long long s_tot[kBuildThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t n_items = n_genomes * n_parts;
  const uint32_t slice_mask = group_slices - 1;
  unsigned long long total = 0;
  uint32_t cur_genome = 0xFFFFFFFFu;
  uint4 *b4 = reinterpret_cast<uint4 *>(s_slice);
  for (int i = tid; i < kSliceWords / 4; i += kBuildThreads) b4[i] = make_uint4(0, 0, 0, 0);  // invariant: zero between slices

  auto flush_total = [&](uint32_t genome) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_down_sync(0xffffffffu, total, o);
    if (lane == 0) s_tot[tid >> 5] = total;
    __syncthreads();
    if (tid == 0) {
      unsigned long long t = 0;
      for (int i = 0; i < kBuildThreads / 32; ++i) t += s_tot[i];
      if (t) atomicAdd(genomes[genome].set_count, t);
    }
    __syncthreads();
    total = 0;
  };

  for (;;) {
    __syncthreads();  // everyone is done with s_item / s_keys of the previous item
    if (tid == 0) s_item = atomicAdd(work_counter, 1u);
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    const uint32_t genome = item / n_parts, part = item - genome * n_parts;
    if (genome != cur_genome) {
      if (cur_genome != 0xFFFFFFFFu) flush_total(cur_genome);
      cur_genome = genome;
    }
    const BuildGenome g = genomes[genome];
    size_t lo;
    uint32_t n;
    if (g.cap) {  // fixed regions filled by the sketch kernel: cursor = slots used (may exceed the region: overflow,
                  // the caller then redoes the genome exactly)
      lo = (size_t)part * g.cap;
      n = __ldg(g.cursor + part);
      if (n > g.cap) n = g.cap;
    } else {      // counting partition: cursor = bucket end after the scatter
      lo = __ldg(g.starts + part);
      n = __ldg(g.cursor + part) - lo;
    }
    const uint32_t *__restrict__ bk = g.bucketed + lo;
    const bool direct = n > kKeyCap;  // more indices than shared memory holds (heavily skewed genome)
    uint32_t n4 = 0;
    if (!direct && n > 0) {
      // straight copy; the tail is padded with copies of the last index (OR is idempotent) so that the
      // scans below can read whole uint4s
      n4 = (n + 3) / 4;
      for (uint32_t i = tid; i < n4 * 4; i += kBuildThreads) s_keys[i] = __ldg(bk + (i < n ? i : n - 1));
    }
    __syncthreads();
    const uint4 *k4 = reinterpret_cast<const uint4 *>(s_keys);
    // kmer_set_size() comes from the atomics themselves: an index is new iff its bit was still clear.  (A POPC
    // per streamed word costs more than the atomics: POPC issues at a quarter of the ALU rate through the same
    // queue as the shared-memory traffic -- ncu: every POPC of the old stream loop stalled on mio_throttle.)
    uint32_t fresh = 0;
    auto put = [&](uint32_t key, uint32_t slice) {
      if (((key >> kSliceBits) & slice_mask) == slice) {
        const uint32_t bit = key & ((1u << kSliceBits) - 1), m = 1u << (bit & 31);
        fresh += (atomicOr(&s_slice[bit >> 5], m) & m) == 0;
      }
    };
    auto clear = [&](uint32_t key, uint32_t slice) {
      if (((key >> kSliceBits) & slice_mask) == slice) s_slice[(key & ((1u << kSliceBits) - 1)) >> 5] = 0;
    };
    for (uint32_t slice = 0; slice < group_slices; ++slice) {
      if (!direct) {
        // the padding copies of the last index find their bit already set: they never count
        for (uint32_t i = tid; i < n4; i += kBuildThreads) {
          const uint4 v = k4[i];
          put(v.x, slice);
          put(v.y, slice);
          put(v.z, slice);
          put(v.w, slice);
        }
      } else {
        for (uint32_t i = tid; i < n; i += kBuildThreads) put(__ldg(bk + i), slice);
      }
      // the slice leaves through one bulk store (TMA engine, SASS UBLKCP); once the engine has read it, the
      // words that were touched are cleared by a re-visit.  The other CTA of the SM assembles meanwhile.
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        bulk_s2g(g.bitset + ((size_t)part * group_slices + slice) * kSliceWords, s_slice, kSliceWords * 4);
        bulk_commit();
        bulk_wait_read<0>();
      }
      __syncthreads();
      if (!direct) {
        for (uint32_t i = tid; i < n4; i += kBuildThreads) {
          const uint4 v = k4[i];
          clear(v.x, slice);
          clear(v.y, slice);
          clear(v.z, slice);
          clear(v.w, slice);
        }
      } else {
        for (uint32_t i = tid; i < n; i += kBuildThreads) clear(__ldg(bk + i), slice);
      }
      __syncthreads();
    }
    total += fresh;
  }
  if (tid == 0) bulk_wait_all();
  if (cur_genome != 0xFFFFFFFFu) flush_total(cur_genome);
}

// ---- fused pair build: K4b + K5 in one pass ---------------------------------------------------------
// The one-call pair pipeline (sks_pair_ani*) needs |A|, |B| and |A n B| of two genomes sketched in the same
// launch.  A CTA takes one bucket of BOTH genomes, assembles slice s of A and slice s of B side by side in
// shared memory with atomicOr, takes |A|, |B|, |A n B| from the atomics' return values, and streams the two
// slices out: the 1 GiB re-read of bitset_pair_counts_kernel disappears.  kStore = false keeps the bitsets on
// chip altogether (nothing but the three counts leaves the SM).  One 1024-thread CTA per SM (216 KB of shared
// memory).  The slices leave through TMA bulk stores, A's while B's is assembled and vice versa.  The kernel
// is bound by the latency of its lock-step phases (a phase is one barrier-to-barrier LDS -> ATOMS chain over
// ~600 indices), not by HBM.  Measured and rejected (DESIGN.md has the numbers): register-staged LDS/STG
// stream-out, 32 KB sub-slices double-buffered through TMA, a counting sort of the bucket by sub-slice,
// 2-4 smaller CTAs per SM, independent warp groups on named barriers, per-lane match queues, and
// zero-fill + global atomics on L2-resident lines (the lines are evicted first: 0.6 GB of DRAM reads).
constexpr int kPairThreads = 1024;
constexpr int kPairSmemBytes = (2 * kSliceWords + 2 * kKeyCap) * 4;

struct PairBuild {
  const uint32_t *regions;  // (genome, bucket) regions of `cap` slots, genome-major
  const uint32_t *cursor;   // [2 * n_parts] absolute end slot of every region (may exceed the region: overflow)
  uint32_t cap;
  uint32_t n_parts;
  uint32_t group_slices;
  uint32_t *bitset[2];               // kStore only
  unsigned long long *out3;          // |A|, |B|, |A n B|
  unsigned int *work_counter;
};

template <bool kStore>
__global__ void __launch_bounds__(kPairThreads, 1) bitset_pair_build_kernel(const __grid_constant__ PairBuild P) {
  extern __shared__ __align__(128) uint32_t s_dyn[];
  uint32_t *s_slice[2] = {s_dyn, s_dyn + kSliceWords};
  uint32_t *s_keys[2] = {s_dyn + 2 * kSliceWords, s_dyn + 2 * kSliceWords + kKeyCap};
  __shared__ uint32_t s_item;
  __shared__ unsigned long long s_tot[3][kPairThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t slice_mask = P.group_slices - 1;
  uint32_t ca = 0, cb = 0, ci = 0;  // < 2^32 bits per genome
  uint32_t n_stored = 0;            // kStore: slices stored so far by this CTA (0 or 1 is all that matters)
  uint4 *a4 = reinterpret_cast<uint4 *>(s_slice[0]);
  for (int i = tid; i < kSliceWords / 2; i += kPairThreads) a4[i] = make_uint4(0, 0, 0, 0);  // both slices (contiguous)

  for (;;) {
    __syncthreads();  // everyone is done with s_item / s_keys of the previous bucket
    if (tid == 0) s_item = atomicAdd(P.work_counter, 1u);
    __syncthreads();
    const uint32_t part = s_item;
    if (part >= P.n_parts) break;
    const uint32_t *bk[2];
    uint32_t n[2], n4[2];
    bool direct[2];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const uint32_t region = g * P.n_parts + part;
      n[g] = __ldg(P.cursor + region);        // slots used in the region
      if (n[g] > P.cap) n[g] = P.cap;          // overflowed region: the caller redoes the pair exactly
      bk[g] = P.regions + (size_t)region * P.cap;
      direct[g] = n[g] > (uint32_t)kKeyCap;
      n4[g] = 0;
      if (!direct[g] && n[g] > 0) {
        n4[g] = (n[g] + 3) / 4;  // tail padded with copies of the last index (OR is idempotent)
        for (uint32_t i = tid; i < n4[g] * 4; i += kPairThreads) s_keys[g][i] = __ldg(bk[g] + (i < n[g] ? i : n[g] - 1));
      }
    }
    __syncthreads();
    // |A|, |B|, |A n B| come from the atomics: an index is new iff its bit was still clear, and a new index of B
    // is shared iff A's finished slice has the bit (no POPC over the streamed words, see bitset_build_kernel).
    auto scan = [&](int g, uint32_t slice, auto &&visit) {
      if (!direct[g]) {
        const uint4 *k4 = reinterpret_cast<const uint4 *>(s_keys[g]);
        for (uint32_t i = tid; i < n4[g]; i += kPairThreads) {
          const uint4 v = k4[i];
          const uint32_t k[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (((k[u] >> kSliceBits) & slice_mask) == slice) visit(k[u] & ((1u << kSliceBits) - 1));
        }
      } else {
        for (uint32_t i = tid; i < n[g]; i += kPairThreads) {
          const uint32_t k = __ldg(bk[g] + i);
          if (((k >> kSliceBits) & slice_mask) == slice) visit(k & ((1u << kSliceBits) - 1));
        }
      }
    };
    if (kStore) {
      // The two slice buffers double-buffer each other: A's slice leaves through a bulk store (TMA engine, SASS
      // UBLKCP) while B's is being assembled and vice versa, so the LSU/MIO queue carries only the atomics.
      // Before a buffer is reused its last store must have read it (wait_group.read 1: only the other buffer's
      // store may still be pending); it is then cleared by re-visiting the indices of the slice it held (by
      // 128-bit zero stores at the first slice of a bucket, whose predecessor's indices are gone).
      for (uint32_t slice = 0; slice < P.group_slices; ++slice) {
        const size_t slice_off = ((size_t)part * P.group_slices + slice) * kSliceWords;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t *buf = s_slice[g];
          if (n_stored) {  // CTA-uniform: this buffer has been stored before
            if (tid == 0) bulk_wait_read<1>();
            __syncthreads();
            if (slice == 0) {
              uint4 *z = reinterpret_cast<uint4 *>(buf);
#pragma unroll
              for (int i = tid; i < kSliceWords / 4; i += kPairThreads) z[i] = make_uint4(0, 0, 0, 0);
            } else {
              scan(g, slice - 1, [&](uint32_t bit) { buf[bit >> 5] = 0; });
            }
            __syncthreads();
          }
          if (g == 0) {
            scan(0, slice, [&](uint32_t bit) {
              const uint32_t m = 1u << (bit & 31);
              ca += (atomicOr(&buf[bit >> 5], m) & m) == 0;
            });
          } else {
            scan(1, slice, [&](uint32_t bit) {
              const uint32_t m = 1u << (bit & 31);
              if ((atomicOr(&buf[bit >> 5], m) & m) == 0) {
                ++cb;
                ci += (s_slice[0][bit >> 5] & m) != 0;
              }
            });
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA engine
          __syncthreads();
          if (tid == 0) {
            bulk_s2g(P.bitset[g] + slice_off, buf, kSliceWords * 4);
            bulk_commit();
          }
        }
        n_stored = 1;
      }
    } else {
      for (uint32_t slice = 0; slice < P.group_slices; ++slice) {
        scan(0, slice, [&](uint32_t bit) {
          const uint32_t m = 1u << (bit & 31);
          ca += (atomicOr(&s_slice[0][bit >> 5], m) & m) == 0;
        });
        __syncthreads();
        scan(1, slice, [&](uint32_t bit) {
          const uint32_t m = 1u << (bit & 31);
          if ((atomicOr(&s_slice[1][bit >> 5], m) & m) == 0) {
            ++cb;
            ci += (s_slice[0][bit >> 5] & m) != 0;
          }
        });
        __syncthreads();
        // nothing leaves the SM: clear only the words that were touched
        scan(0, slice, [&](uint32_t bit) { s_slice[0][bit >> 5] = 0; });
        scan(1, slice, [&](uint32_t bit) { s_slice[1][bit >> 5] = 0; });
        __syncthreads();
      }
    }
  }
  if (kStore && tid == 0) bulk_wait_all();
  unsigned long long t[3] = {ca, cb, ci};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t[k] += __shfl_down_sync(0xffffffffu, t[k], o);
    if (lane == 0) s_tot[k][tid >> 5] = t[k];
  }
  __syncthreads();
  if (tid < 3) {
    unsigned long long sum = 0;
    for (int i = 0; i < kPairThreads / 32; ++i) sum += s_tot[tid][i];
    if (sum) atomicAdd(P.out3 + tid, sum);
  }
}

// ---- sort + unique --------------------------------------------------------------------------------
struct Region {
  unsigned long long begin, end;  // slots
};

template <int KW>
__device__ __forceinline__ bool key_ne(const unsigned long long *k, unsigned long long i, unsigned long long j) {
  if (KW == 1) return k[i] != k[j];
  return k[2 * i] != k[2 * j] || k[2 * i + 1] != k[2 * j + 1];
}

// flags[i] = 1 when slot i is the first occurrence of its key inside its region.
template <int KW>
__global__ void mark_heads_kernel(const unsigned long long *__restrict__ keys, const Region *__restrict__ regions,
                                  uint32_t *__restrict__ flags) {
  const Region r = regions[blockIdx.y];
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = r.begin + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < r.end;
       i += stride)
    flags[i] = (i == r.begin || key_ne<KW>(keys, i, i - 1)) ? 1u : 0u;
}

template <int KW>
__global__ void scatter_heads_kernel(const unsigned long long *__restrict__ keys, const Region *__restrict__ regions,
                                     const uint32_t *__restrict__ flags, const uint32_t *__restrict__ pos,
                                     unsigned long long *__restrict__ out, unsigned long long *__restrict__ uoff,
                                     unsigned long long *__restrict__ ucount) {
  const Region r = regions[blockIdx.y];
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = r.begin + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < r.end;
       i += stride) {
    if (flags[i]) {
      const unsigned long long d = pos[i];
      if (KW == 1) {
        out[d] = keys[i];
      } else {
        out[2 * d] = keys[2 * i];
        out[2 * d + 1] = keys[2 * i + 1];
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (r.end > r.begin) {
      uoff[blockIdx.y] = pos[r.begin];
      ucount[blockIdx.y] = (unsigned long long)pos[r.end - 1] + flags[r.end - 1] - pos[r.begin];
    } else {
      uoff[blockIdx.y] = 0;
      ucount[blockIdx.y] = 0;
    }
  }
}

// Split / join 16-byte keys into (lo, hi) planes for the two-pass stable radix sort.
__global__ void split_keys_kernel(const ulonglong2 *__restrict__ in, unsigned long long *__restrict__ lo,
                                  unsigned long long *__restrict__ hi, unsigned long long n) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const ulonglong2 v = in[i];
    lo[i] = v.x;
    hi[i] = v.y;
  }
}
__global__ void join_keys_kernel(const unsigned long long *__restrict__ lo, const unsigned long long *__restrict__ hi,
                                 ulonglong2 *__restrict__ out, unsigned long long n) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_ulonglong2(lo[i], hi[i]);
}

// ---- sorted-set intersection ------------------------------------------------------------------------
template <int KW>
__device__ __forceinline__ int key_cmp(const unsigned long long *a, long long i, const unsigned long long *b,
                                       long long j) {
  if (KW == 1) {
    const unsigned long long x = a[i], y = b[j];
    return x < y ? -1 : (x > y ? 1 : 0);
  }
  const unsigned long long xh = a[2 * i + 1], yh = b[2 * j + 1];
  if (xh != yh) return xh < yh ? -1 : 1;
  const unsigned long long xl = a[2 * i], yl = b[2 * j];
  return xl < yl ? -1 : (xl > yl ? 1 : 0);
}

// One CTA per pair, one slice of the smaller set per warp.  A warp walks its slice of A and the matching
// range of B in chunks of 32 keys (coalesced 256 B loads, one key per lane): every lane locates its A key in
// the B chunk with a 5-step shuffle search, then whichever chunk has the smaller maximum is advanced
// (both on a tie).  The kernel is issue bound (~65 instructions per step of 32 + 32 keys; a register
// prefetch queue for the next chunks was measured and changed nothing), the sets are L2 resident.
template <int KW>
struct WarpKey;
template <>
struct WarpKey<1> {
  unsigned long long v;
  __device__ __forceinline__ static WarpKey load(const unsigned long long *p, long long i) { return {p[i]}; }
  __device__ __forceinline__ static WarpKey max() { return {~0ull}; }
  __device__ __forceinline__ WarpKey shfl(int src) const { return {__shfl_sync(0xffffffffu, v, src)}; }
  __device__ __forceinline__ bool lt(const WarpKey &o) const { return v < o.v; }
  __device__ __forceinline__ bool eq(const WarpKey &o) const { return v == o.v; }
};
template <>
struct WarpKey<2> {
  unsigned long long lo, hi;
  __device__ __forceinline__ static WarpKey load(const unsigned long long *p, long long i) {
    const ulonglong2 t = reinterpret_cast<const ulonglong2 *>(p)[i];
    return {t.x, t.y};
  }
  __device__ __forceinline__ static WarpKey max() { return {~0ull, ~0ull}; }
  __device__ __forceinline__ WarpKey shfl(int src) const {
    return {__shfl_sync(0xffffffffu, lo, src), __shfl_sync(0xffffffffu, hi, src)};
  }
  __device__ __forceinline__ bool lt(const WarpKey &o) const { return hi != o.hi ? hi < o.hi : lo < o.lo; }
  __device__ __forceinline__ bool eq(const WarpKey &o) const { return lo == o.lo && hi == o.hi; }
};

constexpr int kIntersectThreads = 256;

template <int KW>
__global__ void __launch_bounds__(kIntersectThreads)
    sorted_intersect_kernel(const void *const *__restrict__ pa, const long long *__restrict__ na,
                            const void *const *__restrict__ pb, const long long *__restrict__ nb,
                            int32_t *__restrict__ out, const uint32_t *__restrict__ pair_idx) {
  using K = WarpKey<KW>;
  const long long pair = pair_idx ? (long long)pair_idx[blockIdx.x] : (long long)blockIdx.x;
  const unsigned long long *A = static_cast<const unsigned long long *>(pa[pair]);
  const unsigned long long *B = static_cast<const unsigned long long *>(pb[pair]);
  long long nA = na[pair], nB = nb[pair];
  if (nA > nB) {  // walk the smaller set (src/kmer_set.cpp:26-27)
    const unsigned long long *t = A; A = B; B = t;
    const long long tn = nA; nA = nB; nB = tn;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kIntersectThreads / 32;
  uint32_t cnt = 0;
  if (nA > 0 && nB > 0) {
    // gridDim.y CTAs share one pair (large sets): every warp of every CTA takes its own slice of A
    const long long n_slices = (long long)kWarps * gridDim.y;
    const long long per = ((nA + n_slices - 1) / n_slices + 31) & ~31ll;
    long long ai = ((long long)blockIdx.y * kWarps + warp) * per;
    const long long a_end = ai + per < nA ? ai + per : nA;
    if (ai < a_end) {
      // bi = lower_bound(B, A[ai]) by a 32-ary search: every lane probes one splitter
      const K first = K::load(A, ai);
      long long lo = 0, hi = nB;  // answer in [lo, hi]
      while (hi - lo > 0) {
        const long long span = hi - lo, step = (span + 31) / 32;
        const long long probe = lo + (long long)lane * step;  // splitters lo, lo+step, ...
        const bool less = probe < hi && K::load(B, probe).lt(first);
        const uint32_t m = __ballot_sync(0xffffffffu, less);
        const int k = __popc(m);  // splitters 0..k-1 are < first (monotone)
        if (k == 0) { hi = lo; break; }
        const long long nlo = lo + (long long)(k - 1) * step + 1;
        const long long nhi = (lo + (long long)k * step) < hi ? lo + (long long)k * step : hi;
        lo = nlo;
        hi = nhi;
      }
      long long bi = lo;
      K a = K::max(), b = K::max();
      bool load_a = true, load_b = true;
      while (ai < a_end && bi < nB) {
        if (load_a) a = ai + lane < a_end ? K::load(A, ai + lane) : K::max();
        if (load_b) b = bi + lane < nB ? K::load(B, bi + lane) : K::max();
        // position of a among the 32 b's: number of b_j < a, capped at 31 (an a above every b cannot match)
        int pos = 0;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
          const K probe = b.shfl(pos + s - 1);
          if (probe.lt(a)) pos += s;
        }
        const K hit = b.shfl(pos);
        if (ai + lane < a_end && hit.eq(a) && bi + pos < nB) ++cnt;
        const long long a_last = (a_end - ai < 32 ? a_end - ai : 32) - 1, b_last = (nB - bi < 32 ? nB - bi : 32) - 1;
        const K a_max = a.shfl((int)a_last), b_max = b.shfl((int)b_last);
        load_a = !b_max.lt(a_max);  // a_max <= b_max: this A chunk is done
        load_b = !a_max.lt(b_max);  // b_max <= a_max: this B chunk is done
        if (load_a) ai += 32;
        if (load_b) bi += 32;
      }
    }
  }
  __shared__ uint32_t s[kWarps];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
  if (lane == 0) s[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int k = 0; k < kWarps; ++k) t += s[k];
    if (gridDim.y == 1) out[pair] = (int32_t)t;
    else if (t) atomicAdd(out + pair, (int32_t)t);  // the launcher zeroed out[]
  }
}

}  // namespace

int launch_fill_zero(sks_ctx *ctx, void *ptr, size_t bytes) {
  if (bytes == 0) return SKS_OK;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (bytes & 15)) {
    SKS_CUDA_TRY(cudaMemsetAsync(ptr, 0, bytes, ctx->stream));
    return SKS_OK;
  }
  const size_t n16 = bytes / 16;
  KernelTimer timer(ctx, SKS_KERNEL_FILL);
  fill_zero_kernel<<<stream_grid(ctx, n16), kStreamThreads, 0, ctx->stream>>>(static_cast<uint4 *>(ptr), n16);
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches++;
  return SKS_OK;
}

int launch_bitset_pair_counts(sks_ctx *ctx, const uint32_t *a, const uint32_t *b, uint64_t n_words,
                              unsigned long long *out3) {
  SKS_CUDA_TRY(cudaMemsetAsync(out3, 0, 3 * sizeof(unsigned long long), ctx->stream));
  KernelTimer timer(ctx, SKS_KERNEL_PAIR_COUNTS);
  if (n_words % 4 != 0 || n_words < 4) {
    bitset_small_counts_kernel<<<1, 32, 0, ctx->stream>>>(a, b, n_words, out3);
  } else {
    const size_t n16 = n_words / 4;
    bitset_pair_counts_kernel<<<stream_grid(ctx, n16), kStreamThreads, 0, ctx->stream>>>(
        reinterpret_cast<const uint4 *>(a), reinterpret_cast<const uint4 *>(b), n16, out3);
  }
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches++;
  return SKS_OK;
}

// Builds n_genomes bitsets of 2^index_bits bits each (index_bits > kSliceBits) from the raw PEXT
// indices of every genome: raw_idx + h_off[g] holds h_count[g] indices; bucketed_idx is scratch of the
// same size.  d_set_count[g] receives |set g| (zeroed here).
int launch_bitset_build(sks_ctx *ctx, const uint32_t *raw_idx, uint32_t *bucketed_idx, const uint64_t *h_off,
                        const uint64_t *h_count, int n_genomes, int index_bits, uint32_t *bitset, uint64_t bitset_words,
                        unsigned long long *d_set_count) {
  if (n_genomes == 0) return SKS_OK;
  if (index_bits <= kSliceBits || index_bits > 32)
    return set_error(SKS_ERR_INVALID, "bucketed bitset build handles 20..32 index bits, not %d", index_bits);
  const PartGeometry geo = part_geometry(index_bits);
  const int part_shift = geo.part_shift;
  const uint32_t n_parts = geo.n_parts, group_slices = geo.group_slices;
  uint64_t max_n = 0;
  for (int g = 0; g < n_genomes; ++g) max_n = std::max(max_n, h_count[g]);
  if (max_n >= (1ull << 32)) return set_error(SKS_ERR_CAPACITY, "too many k-mers in one genome for the bucketed build");

  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t sz_desc = align(sizeof(BuildGenome) * n_genomes);
  const size_t sz_tab = align((size_t)(2 * n_parts + 1) * 4);  // starts | cursor per genome
  char *base = nullptr;
  SKS_TRY(ctx_scratch(ctx, sz_desc + sz_tab * n_genomes + 256, reinterpret_cast<void **>(&base)));
  BuildGenome *d_desc = reinterpret_cast<BuildGenome *>(base);
  char *d_tabs = base + sz_desc;
  unsigned int *d_counter = reinterpret_cast<unsigned int *>(d_tabs + sz_tab * n_genomes);
  BuildGenome *h_desc = nullptr;
  SKS_TRY(ctx_pinned(ctx, sizeof(BuildGenome) * n_genomes, reinterpret_cast<void **>(&h_desc)));
  for (int g = 0; g < n_genomes; ++g) {
    uint32_t *tab = reinterpret_cast<uint32_t *>(d_tabs + sz_tab * g);
    h_desc[g].raw = raw_idx + h_off[g];
    h_desc[g].bucketed = bucketed_idx + h_off[g];
    h_desc[g].starts = tab;
    h_desc[g].cursor = tab + n_parts + 1;
    h_desc[g].bitset = bitset + (size_t)g * bitset_words;
    h_desc[g].set_count = d_set_count + g;
    h_desc[g].n = (uint32_t)h_count[g];
    h_desc[g].cap = 0;
  }

  KernelTimer timer(ctx, SKS_KERNEL_BITSET_BUILD);
  SKS_CUDA_TRY(cudaMemsetAsync(d_set_count, 0, sizeof(unsigned long long) * n_genomes, ctx->stream));
  SKS_CUDA_TRY(cudaMemsetAsync(d_tabs, 0, sz_tab * n_genomes + 256, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(d_desc, h_desc, sizeof(BuildGenome) * n_genomes, cudaMemcpyHostToDevice, ctx->stream));
  constexpr int smem = (kSliceWords + kKeyCap) * 4;
  static thread_local bool attr_set[8] = {false, false, false, false, false, false, false, false};
  if (!attr_set[ctx->device & 7]) {
    SKS_CUDA_TRY(cudaFuncSetAttribute(bitset_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[ctx->device & 7] = true;
  }
  const unsigned gx_stream = (unsigned)std::min<uint64_t>((max_n + kScatterKeys - 1) / kScatterKeys + 1,
                                                          (uint64_t)ctx->sm_count * 8);
  for (int g0 = 0; g0 < n_genomes; g0 += 32768) {
    const int ng = std::min(n_genomes - g0, 32768);
    part_hist_kernel<<<dim3(gx_stream, ng), 256, 0, ctx->stream>>>(d_desc + g0, part_shift, n_parts);
    part_scan_kernel<<<ng, kMaxParts, 0, ctx->stream>>>(d_desc + g0, n_parts);
    part_scatter_kernel<<<dim3(gx_stream, ng), 256, 0, ctx->stream>>>(d_desc + g0, part_shift, n_parts);
    ctx->launches += 3;
  }
  const uint64_t n_items = (uint64_t)n_genomes * n_parts;
  if (n_items >= (1ull << 31)) return set_error(SKS_ERR_CAPACITY, "too many bitset slices in one batch");
  const unsigned gx = (unsigned)std::min<uint64_t>(n_items, (uint64_t)ctx->sm_count * 2);  // two ~108 KB CTAs per SM
  bitset_build_kernel<<<gx, kBuildThreads, smem, ctx->stream>>>(d_desc, (uint32_t)n_genomes, n_parts, group_slices,
                                                               d_counter);
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches++;
  return SKS_OK;
}

// The assemble half of the bucketed build when the sketch kernel (OUT_PART) has already scattered the PEXT
// indices into fixed regions: region (g, b) = regions[(g * n_parts + b) * part_cap ...), filled up to
// d_cursor[g * n_parts + b] (absolute slot).
int launch_bitset_assemble(sks_ctx *ctx, const uint32_t *regions, const uint32_t *d_cursor, uint32_t part_cap, int n_genomes,
                           int index_bits, uint32_t *bitset, uint64_t bitset_words, unsigned long long *d_set_count,
                           unsigned int *d_work_counter) {
  if (n_genomes == 0) return SKS_OK;
  const PartGeometry geo = part_geometry(index_bits);
  BuildGenome *d_desc = nullptr, *h_desc = nullptr;
  SKS_TRY(ctx_scratch(ctx, sizeof(BuildGenome) * n_genomes, reinterpret_cast<void **>(&d_desc)));
  SKS_TRY(ctx_pinned(ctx, sizeof(BuildGenome) * n_genomes, reinterpret_cast<void **>(&h_desc)));
  for (int g = 0; g < n_genomes; ++g) {
    h_desc[g].raw = nullptr;
    h_desc[g].bucketed = const_cast<uint32_t *>(regions) + (size_t)g * geo.n_parts * part_cap;
    h_desc[g].starts = nullptr;
    h_desc[g].cursor = const_cast<uint32_t *>(d_cursor) + (size_t)g * geo.n_parts;
    h_desc[g].bitset = bitset + (size_t)g * bitset_words;
    h_desc[g].set_count = d_set_count + g;
    h_desc[g].n = 0;
    h_desc[g].cap = part_cap;
  }
  // d_set_count and d_work_counter were zeroed by the caller together with the cursors
  KernelTimer timer(ctx, SKS_KERNEL_BITSET_BUILD);
  SKS_CUDA_TRY(cudaMemcpyAsync(d_desc, h_desc, sizeof(BuildGenome) * n_genomes, cudaMemcpyHostToDevice, ctx->stream));
  constexpr int smem = (kSliceWords + kKeyCap) * 4;
  SKS_CUDA_TRY(cudaFuncSetAttribute(bitset_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const uint64_t n_items = (uint64_t)n_genomes * geo.n_parts;
  const unsigned gx = (unsigned)std::min<uint64_t>(n_items, (uint64_t)ctx->sm_count * 2);
  bitset_build_kernel<<<gx, kBuildThreads, smem, ctx->stream>>>(d_desc, (uint32_t)n_genomes, geo.n_parts, geo.group_slices,
                                                               d_work_counter);
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches++;
  return SKS_OK;
}

// Fused K4b + K5 for a 2-genome batch whose indices the sketch kernel (OUT_PART) scattered into fixed regions.
// bitset_a / bitset_b may both be NULL: the bitsets then never leave shared memory.  d_out3 = |A|, |B|, |A n B|.
int launch_bitset_pair_build(sks_ctx *ctx, const uint32_t *regions, const uint32_t *d_cursor, uint32_t part_cap, int index_bits,
                             uint32_t *bitset_a, uint32_t *bitset_b, unsigned long long *d_out3, unsigned int *d_work_counter) {
  const PartGeometry geo = part_geometry(index_bits);
  PairBuild p;
  p.regions = regions;
  p.cursor = d_cursor;
  p.cap = part_cap;
  p.n_parts = geo.n_parts;
  p.group_slices = geo.group_slices;
  p.bitset[0] = bitset_a;
  p.bitset[1] = bitset_b;
  p.out3 = d_out3;
  p.work_counter = d_work_counter;  // zeroed by the caller together with d_out3 and the cursors
  const bool store = bitset_a != nullptr;
  KernelTimer timer(ctx, SKS_KERNEL_PAIR_BUILD);
  const unsigned gx = std::min<unsigned>(geo.n_parts, (unsigned)ctx->sm_count);
  if (store) {
    SKS_CUDA_TRY(cudaFuncSetAttribute(bitset_pair_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    bitset_pair_build_kernel<true><<<gx, kPairThreads, kPairSmemBytes, ctx->stream>>>(p);
  } else {
    SKS_CUDA_TRY(cudaFuncSetAttribute(bitset_pair_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes));
    bitset_pair_build_kernel<false><<<gx, kPairThreads, kPairSmemBytes, ctx->stream>>>(p);
  }
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches++;
  return SKS_OK;
}

int launch_bitset_popcount(sks_ctx *ctx, const uint32_t *a, uint64_t n_words, unsigned long long *out1) {
  SKS_CUDA_TRY(cudaMemsetAsync(out1, 0, sizeof(unsigned long long), ctx->stream));
  KernelTimer timer(ctx, SKS_KERNEL_POPCOUNT);
  if (n_words % 4 != 0 || n_words < 4) {
    bitset_small_counts_kernel<<<1, 32, 0, ctx->stream>>>(a, nullptr, n_words, out1);
  } else {
    const size_t n16 = n_words / 4;
    bitset_popcount_kernel<<<stream_grid(ctx, n16), kStreamThreads, 0, ctx->stream>>>(
        reinterpret_cast<const uint4 *>(a), n16, out1);
  }
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches++;
  return SKS_OK;
}

// Sorts and de-duplicates n_regions independent key regions that live in one buffer.
//   keys      : device buffer of `span` slots (key_words uint64 per slot); region g occupies slots
//               [h_off[g], h_off[g] + h_count[g]).
//   out_buf   : receives the distinct keys of all regions back to back (region order kept);
//   out_off / out_count: per-region slot offset and count inside out_buf.
// ---- sort + unique of 8-byte keys without the library sort -----------------------------------------------
// A FracMinHash sketch is small (25 k keys per 5 Mbp genome, 1.25 M for 250 Mbp) and an 8-pass radix sort
// spends most of its time in per-pass latency (0.19 ms for 250 k keys, 0.28 ms for 1.25 M).  Instead: one MSD
// partition on the top key bits below the mask's highest bit into buckets of <= 4096 keys (shared-memory
// histogram, scan, ranked scatter -- the scheme of the bitset build), then every bucket is sorted on its
// remaining bits and made unique by one CTA and copied to its final place.  Buckets that come out larger than a
// CTA holds (skewed keys) raise a flag and the call falls back to the device-wide radix sort.
constexpr int kSortCap = 4096;          // keys per bucket (32 KB of shared memory)
constexpr int kSortMaxBucketBits = 12;

// Key of the bucket sort: 8 bytes (windows <= 32) or 16 bytes as {lo, hi}.
template <int KW>
struct SortKey;
template <>
struct SortKey<1> {
  using T = unsigned long long;
  __device__ __forceinline__ static T pad() { return ~0ull; }
  __device__ __forceinline__ static bool gt(T a, T b) { return a > b; }
  __device__ __forceinline__ static bool ne(T a, T b) { return a != b; }
  __device__ __forceinline__ static uint32_t bucket(T k, const SortPlan &p) {
    uint32_t b = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (i < p.n_pieces) b |= ((uint32_t)(k >> p.s[i]) & p.m[i]) << p.o[i];
    return b;
  }
};
template <>
struct SortKey<2> {
  using T = ulonglong2;  // x = low word, y = high word
  __device__ __forceinline__ static T pad() { return make_ulonglong2(~0ull, ~0ull); }
  __device__ __forceinline__ static bool gt(T a, T b) { return a.y != b.y ? a.y > b.y : a.x > b.x; }
  __device__ __forceinline__ static bool ne(T a, T b) { return a.x != b.x || a.y != b.y; }
  __device__ __forceinline__ static uint32_t bucket(T k, const SortPlan &p) {
    uint32_t b = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if (i < p.n_pieces) b |= ((uint32_t)((p.word[i] ? k.y : k.x) >> p.s[i]) & p.m[i]) << p.o[i];
    return b;
  }
};

template <int KW>
__global__ void __launch_bounds__(256)
    sortp_hist_kernel(const typename SortKey<KW>::T *__restrict__ keys, const Region *__restrict__ regions,
                      uint32_t *__restrict__ hist, int bb, const __grid_constant__ SortPlan plan) {
  __shared__ uint32_t s_hist[1 << kSortMaxBucketBits];
  const Region r = regions[blockIdx.y];
  const uint32_t nb = 1u << bb;
  for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = r.begin + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < r.end; i += stride)
    atomicAdd(&s_hist[SortKey<KW>::bucket(keys[i], plan)], 1u);
  __syncthreads();
  uint32_t *h = hist + ((size_t)blockIdx.y << bb);
  for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x)
    if (s_hist[i]) atomicAdd(h + i, s_hist[i]);
}

// One 1024-thread CTA per row: out[i] = exclusive prefix of in[i] over the row's `n` entries (n <= 4096);
// row_total[row] = the sum; *flag is raised when an entry exceeds `limit`.
__global__ void __launch_bounds__(1024)
    sortp_scan_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t *__restrict__ out2, uint32_t n,
                      unsigned long long *__restrict__ row_total, uint32_t limit, uint32_t *__restrict__ flag) {
  __shared__ uint32_t s_warp[32];
  const uint32_t t = threadIdx.x, lane = t & 31;
  const uint32_t *row = in + (size_t)blockIdx.x * n;
  uint32_t v[4], sum = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t i = t * 4 + k;
    v[k] = i < n ? row[i] : 0u;
    if (v[k] > limit) *flag = 1u;
    sum += v[k];
  }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += x;
  }
  if (lane == 31) s_warp[t >> 5] = incl;
  __syncthreads();
  if (t < 32) {
    uint32_t w = s_warp[t];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, w, o);
      if (t >= (uint32_t)o) w += x;
    }
    s_warp[t] = w;
  }
  __syncthreads();
  uint32_t run = incl - sum + ((t >> 5) ? s_warp[(t >> 5) - 1] : 0u);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t i = t * 4 + k;
    if (i < n) {
      out[(size_t)blockIdx.x * n + i] = run;
      if (out2) out2[(size_t)blockIdx.x * n + i] = run;
    }
    run += v[k];
  }
  if (t == 1023 && row_total) row_total[blockIdx.x] = run;
}

constexpr int kSortChunk = 4096;  // keys per CTA pass of the scatter

template <int KW>
__global__ void __launch_bounds__(256)
    sortp_scatter_kernel(const typename SortKey<KW>::T *__restrict__ keys, const Region *__restrict__ regions,
                         uint32_t *__restrict__ cursor, typename SortKey<KW>::T *__restrict__ tmp, int bb,
                         const __grid_constant__ SortPlan plan, const uint32_t *__restrict__ flag) {
  __shared__ uint32_t s_hist[1 << kSortMaxBucketBits];
  __shared__ uint32_t s_base[1 << kSortMaxBucketBits];
  __shared__ uint32_t s_abort;
  // the scan found an oversized bucket: the call is going to the library sort anyway.  The decision is made once per CTA:
  // the flag can be raised while this CTA runs, and threads that left on their own would leave the others with
  // half-filled shared tables
  if (threadIdx.x == 0) s_abort = *flag;
  __syncthreads();
  if (s_abort) return;
  const Region r = regions[blockIdx.y];
  const uint32_t nb = 1u << bb;
  uint32_t *cur = cursor + ((size_t)blockIdx.y << bb);
  const unsigned long long n = r.end - r.begin;
  const unsigned long long n_chunks = (n + kSortChunk - 1) / kSortChunk;
  for (unsigned long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    typename SortKey<KW>::T key[kSortChunk / 256];
    uint32_t rank[kSortChunk / 256];
    const unsigned long long base = r.begin + chunk * kSortChunk;
#pragma unroll
    for (int u = 0; u < kSortChunk / 256; ++u) {
      const unsigned long long i = base + u * 256 + threadIdx.x;
      if (i < r.end) key[u] = keys[i];
    }
#pragma unroll
    for (int u = 0; u < kSortChunk / 256; ++u) {
      const unsigned long long i = base + u * 256 + threadIdx.x;
      if (i < r.end) rank[u] = atomicAdd(&s_hist[SortKey<KW>::bucket(key[u], plan)], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < nb; i += blockDim.x) s_base[i] = s_hist[i] ? atomicAdd(cur + i, s_hist[i]) : 0u;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kSortChunk / 256; ++u) {
      const unsigned long long i = base + u * 256 + threadIdx.x;
      if (i < r.end) tmp[r.begin + s_base[SortKey<KW>::bucket(key[u], plan)] + rank[u]] = key[u];
    }
    __syncthreads();
  }
}

// One CTA per (region, bucket): bitonic sort in shared memory over the next power of two above the bucket's
// size (the partition aims at ~300 keys per bucket: the network costs log^2 per key), then the distinct keys go to
// tmp2 at the bucket's place and their number to ucount.  (cub::BlockRadixSort over the full 4096-key capacity
// was measured too: 0.21 ms at C3 and 0.59 ms at C4 against 0.19 / 0.33 ms for a 2048-key network.)
template <int KW, int kSortThreads, int kCap>
__global__ void __launch_bounds__(kSortThreads)
    sortp_bucket_kernel(const typename SortKey<KW>::T *__restrict__ tmp, const Region *__restrict__ regions,
                        const uint32_t *__restrict__ boff, const uint32_t *__restrict__ hist,
                        typename SortKey<KW>::T *__restrict__ tmp2, uint32_t *__restrict__ ucount, int bb,
                        uint32_t *__restrict__ flag, uint32_t uniform_cap) {
  // uniform_cap != 0: bucket q occupies the slots [q * uniform_cap, q * uniform_cap + hist[q]) (the sketch kernel wrote
  // the kept k-mers straight into per-bucket regions); regions / boff are not used then
  using K = SortKey<KW>;
  __shared__ typename K::T s[kCap];
  __shared__ uint32_t s_warp[kSortThreads / 32];
  __shared__ uint32_t s_abort;
  // one decision per CTA (another CTA may raise the flag while this one runs: threads leaving on their own would leave
  // holes in the shared arrays and garbage in the compaction's prefix sums)
  if (threadIdx.x == 0) s_abort = *flag;
  __syncthreads();
  if (s_abort) return;
  const uint32_t n = hist[blockIdx.x];
  if (n == 0 || n > (uint32_t)kCap) {  // uniform; an oversized bucket sends the whole call to the library sort
    if (threadIdx.x == 0) {
      ucount[blockIdx.x] = 0;
      if (n) *flag = 1u;
    }
    return;
  }
  const unsigned long long lo = uniform_cap ? (unsigned long long)blockIdx.x * uniform_cap
                                            : regions[blockIdx.x >> bb].begin + boff[blockIdx.x];
  uint32_t P = 32;
  while (P < n) P <<= 1;
  const uint32_t tid = threadIdx.x;
  // the tail is padded with copies of the largest possible key; index < n decides what is real afterwards
  uint32_t k_first = 2;
  if constexpr (KW == 1) {
    // the first five phases (runs of 32 keys, 15 of the network's substages) never leave the registers: a warp holds
    // a run, one key per lane, and exchanges through shuffles -- the shared-memory network below is bound by the
    // shared-memory bandwidth (ncu profiles/r01_t_*: 81 %), and these substages would be a third of its traffic.
    // (More keys per lane -- runs of 64 to 256 with register compare-exchanges for the short strides -- were
    // measured too: no faster, the longer dependent chains leave the warps waiting.)
    for (uint32_t i = tid; i < P; i += kSortThreads) {  // P is a multiple of 32: whole warps
      unsigned long long v = i < n ? tmp[lo + i] : K::pad();
      const uint32_t ln = tid & 31;
#pragma unroll
      for (uint32_t k = 2; k <= 32; k <<= 1) {
        const bool up = (i & k) == 0;
#pragma unroll
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, (int)j);
          const bool want_min = ((ln & j) == 0) == up;
          const bool take = want_min ? (o < v) : (o > v);
          v = take ? o : v;
        }
      }
      s[i] = v;
    }
    k_first = 64;
  } else {
    for (uint32_t i = tid; i < P; i += kSortThreads) s[i] = i < n ? tmp[lo + i] : K::pad();
  }
  __syncthreads();
  // 8-byte keys: only the strides from 32 up go through shared memory; the five short strides of every phase are
  // one pass in registers (a warp loads 32 consecutive keys, exchanges through shuffles, stores them back)
  constexpr uint32_t j_min = KW == 1 ? 32u : 1u;
  for (uint32_t k = k_first; k <= P; k <<= 1) {
    for (uint32_t j = k >> 1; j >= j_min; j >>= 1) {
      for (uint32_t t = tid; t < P / 2; t += kSortThreads) {
        const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the pair
        const uint32_t l = i | j;
        const typename K::T a = s[i], b = s[l];
        const bool up = (i & k) == 0;
        if (K::gt(a, b) == up) {
          s[i] = b;
          s[l] = a;
        }
      }
      __syncthreads();
    }
    if constexpr (KW == 1) {
      for (uint32_t i = tid; i < P; i += kSortThreads) {
        unsigned long long v = s[i];
        const bool up = (i & k) == 0;
        const uint32_t ln = tid & 31;
#pragma unroll
        for (uint32_t j = 16; j > 0; j >>= 1) {
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, (int)j);
          const bool want_min = ((ln & j) == 0) == up;
          const bool take = want_min ? (o < v) : (o > v);
          v = take ? o : v;
        }
        s[i] = v;
      }
      __syncthreads();
    }
  }
  // distinct keys: heads, block-wide exclusive scan, compact (per = consecutive elements per thread, <= 16)
  const uint32_t per = (P + kSortThreads - 1) / kSortThreads;
  uint32_t head = 0, cnt = 0;
  for (uint32_t e = 0; e < per; ++e) {
    const uint32_t i = tid * per + e;
    if (i < n && (i == 0 || K::ne(s[i], s[i - 1]))) {
      head |= 1u << e;
      ++cnt;
    }
  }
  uint32_t incl = cnt;
  const uint32_t lane = tid & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += x;
  }
  if (lane == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  uint32_t before = incl - cnt;
  for (uint32_t w = 0; w < (tid >> 5); ++w) before += s_warp[w];
  for (uint32_t e = 0; e < per; ++e)
    if (head & (1u << e)) tmp2[lo + before++] = s[tid * per + e];
  if (tid == kSortThreads - 1) ucount[blockIdx.x] = before;
}

// Exclusive prefix over the regions' distinct totals (one CTA; n_regions is at most a few thousand).
__global__ void __launch_bounds__(1024)
    sortp_region_offsets_kernel(const unsigned long long *__restrict__ utot, unsigned long long *__restrict__ uoff, int n_regions,
                                unsigned long long bound = ~0ull, uint32_t *__restrict__ flag = nullptr) {
  __shared__ unsigned long long s_carry;
  __shared__ unsigned long long s_warp[32];
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_regions; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned long long v = i < n_regions ? utot[i] : 0ull;
    unsigned long long incl = v;
    const uint32_t lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += x;
    }
    if (lane == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned long long before = s_carry + incl - v;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) before += s_warp[w];
    if (i < n_regions) uoff[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = before + v;
    __syncthreads();
  }
  if (flag && threadIdx.x == 0 && s_carry > bound) *flag = 1u;  // more distinct keys than the output holds: the caller starts over
}

template <int KW>
__global__ void __launch_bounds__(256)
    sortp_copy_kernel(const typename SortKey<KW>::T *__restrict__ tmp2, const Region *__restrict__ regions,
                      const uint32_t *__restrict__ boff, const uint32_t *__restrict__ uboff, const uint32_t *__restrict__ ucount,
                      const unsigned long long *__restrict__ uoff, typename SortKey<KW>::T *__restrict__ out, int bb,
                      const uint32_t *__restrict__ flag, uint32_t uniform_cap) {
  if (*flag) return;
  const uint32_t n = ucount[blockIdx.x];
  const uint32_t region = blockIdx.x >> bb;
  const unsigned long long src = uniform_cap ? (unsigned long long)blockIdx.x * uniform_cap : regions[region].begin + boff[blockIdx.x];
  const unsigned long long dst = uoff[region] + uboff[blockIdx.x];
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) out[dst + i] = tmp2[src + i];
}

// Returns SKS_OK with *handled = false when the keys are too skewed (or too many) for the bucket sort.
template <int KW>
int sort_unique_buckets(sks_ctx *ctx, typename SortKey<KW>::T *keys, const uint64_t *h_off, const uint64_t *h_count, int n_regions,
                        uint64_t span, uint64_t total, const uint64_t mask[2], BufferRef *out_buf,
                        std::vector<uint64_t> *out_off, std::vector<uint64_t> *out_count, bool *handled, int skip_bits) {
  *handled = false;
  uint64_t max_count = 0;
  for (int g = 0; g < n_regions; ++g) max_count = std::max(max_count, h_count[g]);
  int bb = 0;
  // ~300 keys per bucket (a 512-key network) while the partition stays coarse enough for its ranked scatter
  // (>= 4 keys per bucket and 4096-key chunk); beyond 1024 buckets only as far as the network's capacity demands
  while (bb < 10 && (max_count >> bb) > 320) ++bb;
  while (bb < kSortMaxBucketBits && (max_count >> bb) > 1536) ++bb;
  // 16-byte keys: half as many fit a CTA
  const uint64_t big_cap = KW == 1 ? kSortCap : kSortCap / 2;
  if (KW == 2)
    while (bb < kSortMaxBucketBits && (max_count >> bb) > 768) ++bb;
  if ((max_count >> bb) > big_cap / 2 || ((uint64_t)n_regions << bb) > (1u << 22)) return SKS_OK;  // too large for this scheme
  const int mask_bits = __builtin_popcountll(mask[0]) + __builtin_popcountll(mask[1]);
  // fewer possible keys than 4 per raw key: the input is mostly duplicates and the partition would be all contention
  if (mask_bits - skip_bits < 62 && ((uint64_t)1 << std::max(mask_bits - skip_bits, 0)) < 4 * max_count) return SKS_OK;
  SortPlan plan;
  if (!sort_plan(mask, bb, &plan, skip_bits)) return SKS_OK;
  using KT = typename SortKey<KW>::T;
  constexpr size_t kb = sizeof(KT);
  const size_t n_b = (size_t)n_regions << bb;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t sz_regions = align(sizeof(Region) * n_regions), sz_tab = align(4 * n_b), sz_plane = align(kb * span);
  const size_t sz_u64 = align(8 * (size_t)n_regions);
  // scratch: regions | hist | boff | cursor | ucount | uboff | flag | utot | uoff | ucnt64 | tmp | tmp2
  char *base = nullptr;
  SKS_TRY(ctx_scratch(ctx, sz_regions + 5 * sz_tab + 256 + 3 * sz_u64 + 2 * sz_plane, reinterpret_cast<void **>(&base)));
  size_t o = 0;
  Region *d_regions = reinterpret_cast<Region *>(base + o); o += sz_regions;
  uint32_t *d_hist = reinterpret_cast<uint32_t *>(base + o); o += sz_tab;
  uint32_t *d_boff = reinterpret_cast<uint32_t *>(base + o); o += sz_tab;
  uint32_t *d_cursor = reinterpret_cast<uint32_t *>(base + o); o += sz_tab;
  uint32_t *d_ucount = reinterpret_cast<uint32_t *>(base + o); o += sz_tab;
  uint32_t *d_uboff = reinterpret_cast<uint32_t *>(base + o); o += sz_tab;
  uint32_t *d_flag = reinterpret_cast<uint32_t *>(base + o); o += 256;
  unsigned long long *d_utot = reinterpret_cast<unsigned long long *>(base + o); o += sz_u64;
  unsigned long long *d_uoff = reinterpret_cast<unsigned long long *>(base + o); o += sz_u64;
  o += sz_u64;
  KT *d_tmp = reinterpret_cast<KT *>(base + o); o += sz_plane;
  KT *d_tmp2 = reinterpret_cast<KT *>(base + o);

  Region *h_regions = nullptr;
  SKS_TRY(ctx_pinned(ctx, sizeof(Region) * n_regions, reinterpret_cast<void **>(&h_regions)));
  for (int g = 0; g < n_regions; ++g) h_regions[g] = {h_off[g], h_off[g] + h_count[g]};
  SKS_CUDA_TRY(cudaMemcpyAsync(d_regions, h_regions, sizeof(Region) * n_regions, cudaMemcpyHostToDevice, ctx->stream));
  SKS_TRY(alloc_buffer(ctx, kb * total, out_buf));
  KT *d_out = static_cast<KT *>((*out_buf)->ptr);

  KernelTimer timer(ctx, SKS_KERNEL_SORT_UNIQUE);
  SKS_CUDA_TRY(cudaMemsetAsync(d_hist, 0, sz_tab, ctx->stream));
  SKS_CUDA_TRY(cudaMemsetAsync(d_flag, 0, 256, ctx->stream));
  const unsigned per_region = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((max_count + kSortChunk - 1) / kSortChunk,
                                                                               (uint64_t)ctx->sm_count * 8));
  dim3 grid(per_region, (unsigned)n_regions);
  const uint32_t nb = 1u << bb;
  sortp_hist_kernel<KW><<<grid, 256, 0, ctx->stream>>>(keys, d_regions, d_hist, bb, plan);
  SKS_CUDA_TRY(cudaGetLastError());
  sortp_scan_kernel<<<n_regions, 1024, 0, ctx->stream>>>(d_hist, d_boff, d_cursor, nb, nullptr, (uint32_t)big_cap, d_flag);
  SKS_CUDA_TRY(cudaGetLastError());
  sortp_scatter_kernel<KW><<<grid, 256, 0, ctx->stream>>>(keys, d_regions, d_cursor, d_tmp, bb, plan, d_flag);
  SKS_CUDA_TRY(cudaGetLastError());
  if ((max_count >> bb) > 600)  // 2048-key networks: more threads per bucket
    sortp_bucket_kernel<KW, 512, kSortCap / KW><<<(unsigned)n_b, 512, 0, ctx->stream>>>(d_tmp, d_regions, d_boff, d_hist, d_tmp2,
                                                                             d_ucount, bb, d_flag, 0);
  else  // ~300 keys per bucket on average: a 256/512-key network keeps 128 threads busy (half of 256 would idle);
        // 8 KB of shared memory per CTA keeps many buckets in flight per SM
    sortp_bucket_kernel<KW, 128, 1024><<<(unsigned)n_b, 128, 0, ctx->stream>>>(d_tmp, d_regions, d_boff, d_hist, d_tmp2, d_ucount,
                                                                         bb, d_flag, 0);
  SKS_CUDA_TRY(cudaGetLastError());
  sortp_scan_kernel<<<n_regions, 1024, 0, ctx->stream>>>(d_ucount, d_uboff, nullptr, nb, d_utot, 0xFFFFFFFFu, d_flag);
  SKS_CUDA_TRY(cudaGetLastError());
  sortp_region_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(d_utot, d_uoff, n_regions);
  SKS_CUDA_TRY(cudaGetLastError());
  sortp_copy_kernel<KW><<<(unsigned)n_b, 256, 0, ctx->stream>>>(d_tmp2, d_regions, d_boff, d_uboff, d_ucount, d_uoff, d_out, bb,
                                                            d_flag, 0);
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches += 7;

  unsigned long long *h_back = nullptr;
  SKS_TRY(ctx_pinned(ctx, 16 * (size_t)n_regions + 64, reinterpret_cast<void **>(&h_back)));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_back, d_utot, 8 * (size_t)n_regions, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_back + n_regions, d_uoff, 8 * (size_t)n_regions, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_back + 2 * n_regions, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if ((uint32_t)h_back[2 * n_regions] != 0) {  // a bucket outgrew the network: the caller sorts with the library
    out_buf->reset();
    return SKS_OK;
  }
  for (int g = 0; g < n_regions; ++g) {
    (*out_count)[g] = h_back[g];
    (*out_off)[g] = h_back[n_regions + g];
  }
  *handled = true;
  return SKS_OK;
}

// Geometry of the bucket sort when its first level is folded into the sketch kernel's emit (SketchParams::kpart_*):
// bucket bits and slots per (genome, bucket) region for genomes of up to `max_count` kept k-mers each.  False when the
// input is not for this scheme (the caller then emits per genome and partitions afterwards).
bool bucket_regions_plan(const uint64_t mask[2], int n_genomes, uint64_t max_count, int *bb_out, uint32_t *cap_out, SortPlan *plan) {
  static const bool enabled = getenv("SKS_SKETCH_PART") ? atoi(getenv("SKS_SKETCH_PART")) != 0 : true;
  static const bool bucket_sort = getenv("SKS_BUCKET_SORT") ? atoi(getenv("SKS_BUCKET_SORT")) != 0 : true;
  if (!enabled || !bucket_sort || n_genomes < 1 || max_count < 64) return false;
  int bb = 1;  // SketchParams::kpart_bits == 0 means "route off" to the kernel: at least two buckets per genome
  while (bb < 10 && (max_count >> bb) > 320) ++bb;
  while (bb < kSortMaxBucketBits && (max_count >> bb) > 1536) ++bb;
  const uint64_t avg = max_count >> bb;
  if (avg > (uint64_t)kSortCap / 4 || ((uint64_t)n_genomes << bb) > (1u << 22)) return false;
  const int mask_bits = __builtin_popcountll(mask[0]) + __builtin_popcountll(mask[1]);
  if (mask_bits < 62 && ((uint64_t)1 << mask_bits) < 4 * max_count) return false;  // mostly duplicates: see sort_unique_buckets
  if (!sort_plan(mask, bb, plan)) return false;
  // three times the mean load of a bucket: uniform keys (random sequence) stay far below, a skewed genome overflows and
  // takes the exact route
  const uint64_t cap = std::min<uint64_t>(3 * avg + 64, avg > 300 ? kSortCap : 1024);
  if (((uint64_t)n_genomes << bb) * cap * 8 > ((uint64_t)4 << 30)) return false;
  *bb_out = bb;
  *cap_out = (uint32_t)cap;
  return true;
}

// Second half of the bucket sort on regions the sketch kernel filled: region q = (genome << bb) + bucket holds
// d_cursor[q] keys at q * cap.  *handled == false: a region overflowed or a bucket was too large -- nothing was produced.
int sort_unique_from_buckets(sks_ctx *ctx, unsigned long long *regions, int n_genomes, int bb, uint32_t cap, const uint32_t *d_cursor,
                             uint32_t *d_flag, uint64_t total_bound, BufferRef *out_buf, std::vector<uint64_t> *out_off,
                             std::vector<uint64_t> *out_count, bool *handled) {
  *handled = false;
  out_off->assign(n_genomes, 0);
  out_count->assign(n_genomes, 0);
  const size_t n_b = (size_t)n_genomes << bb;
  const uint32_t nb = 1u << bb;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t sz_tab = align(4 * n_b), sz_u64 = align(8 * (size_t)n_genomes);
  char *base = nullptr;
  SKS_TRY(ctx_scratch(ctx, 2 * sz_tab + 2 * sz_u64, reinterpret_cast<void **>(&base)));
  uint32_t *d_ucount = reinterpret_cast<uint32_t *>(base), *d_uboff = reinterpret_cast<uint32_t *>(base + sz_tab);
  unsigned long long *d_utot = reinterpret_cast<unsigned long long *>(base + 2 * sz_tab);
  unsigned long long *d_uoff = reinterpret_cast<unsigned long long *>(base + 2 * sz_tab + sz_u64);
  SKS_TRY(alloc_buffer(ctx, 8 * std::max<uint64_t>(total_bound, 2), out_buf));
  unsigned long long *d_out = static_cast<unsigned long long *>((*out_buf)->ptr);
  {
    KernelTimer timer(ctx, SKS_KERNEL_SORT_UNIQUE);
    if (cap > 1024)
      sortp_bucket_kernel<1, 512, kSortCap><<<(unsigned)n_b, 512, 0, ctx->stream>>>(regions, nullptr, nullptr, d_cursor, regions, d_ucount,
                                                                              bb, d_flag, cap);
    else
      sortp_bucket_kernel<1, 128, 1024><<<(unsigned)n_b, 128, 0, ctx->stream>>>(regions, nullptr, nullptr, d_cursor, regions, d_ucount, bb,
                                                                            d_flag, cap);
    SKS_CUDA_TRY(cudaGetLastError());
    sortp_scan_kernel<<<n_genomes, 1024, 0, ctx->stream>>>(d_ucount, d_uboff, nullptr, nb, d_utot, 0xFFFFFFFFu, d_flag);
    SKS_CUDA_TRY(cudaGetLastError());
    sortp_region_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(d_utot, d_uoff, n_genomes, total_bound, d_flag);
    SKS_CUDA_TRY(cudaGetLastError());
    sortp_copy_kernel<1><<<(unsigned)n_b, 256, 0, ctx->stream>>>(regions, nullptr, nullptr, d_uboff, d_ucount, d_uoff, d_out, bb, d_flag, cap);
    SKS_CUDA_TRY(cudaGetLastError());
    ctx->launches += 4;
  }
  unsigned long long *h_back = nullptr;
  SKS_TRY(ctx_pinned(ctx, 16 * (size_t)n_genomes + 64, reinterpret_cast<void **>(&h_back)));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_back, d_utot, 8 * (size_t)n_genomes, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_back + n_genomes, d_uoff, 8 * (size_t)n_genomes, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_back + 2 * n_genomes, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if ((uint32_t)h_back[2 * n_genomes] != 0) {
    out_buf->reset();
    return SKS_OK;
  }
  for (int g = 0; g < n_genomes; ++g) {
    (*out_count)[g] = h_back[g];
    (*out_off)[g] = h_back[n_genomes + g];
  }
  *handled = true;
  return SKS_OK;
}

int sort_unique_regions(sks_ctx *ctx, int key_words, void *keys, const uint64_t *h_off, const uint64_t *h_count,
                        int n_regions, uint64_t span, BufferRef *out_buf, std::vector<uint64_t> *out_off,
                        std::vector<uint64_t> *out_count, const uint64_t *mask, int skip_bits) {
  out_off->assign(n_regions, 0);
  out_count->assign(n_regions, 0);
  uint64_t total = 0;
  for (int g = 0; g < n_regions; ++g) total += h_count[g];
  if (total == 0) {
    SKS_TRY(alloc_buffer(ctx, 16, out_buf));
    return SKS_OK;
  }
  if (span >= (1ull << 31)) return set_error(SKS_ERR_CAPACITY, "sort span of %llu slots exceeds 2^31", (unsigned long long)span);
  static const bool bucket_sort = getenv("SKS_BUCKET_SORT") ? atoi(getenv("SKS_BUCKET_SORT")) != 0 : true;
  if (bucket_sort && mask) {
    bool handled = false;
    if (key_words == 1)
      SKS_TRY(sort_unique_buckets<1>(ctx, static_cast<unsigned long long *>(keys), h_off, h_count, n_regions, span, total,
                                     mask, out_buf, out_off, out_count, &handled, skip_bits));
    else
      SKS_TRY(sort_unique_buckets<2>(ctx, static_cast<ulonglong2 *>(keys), h_off, h_count, n_regions, span, total, mask,
                                     out_buf, out_off, out_count, &handled, skip_bits));
    if (handled) return SKS_OK;
  }

  const size_t kb = (size_t)key_words * 8;
  uint64_t largest = 0;
  for (int g = 0; g < n_regions; ++g) largest = std::max(largest, h_count[g]);
  const bool per_region = n_regions == 1 || (n_regions <= 64 && largest >= 65536);
  // scratch: regions | begin/end offsets | alt keys (| lo/hi planes) | flags | pos | uoff | ucount | cub temp
  std::vector<Region> h_regions(n_regions);
  std::vector<long long> h_begin(n_regions), h_end(n_regions);
  for (int g = 0; g < n_regions; ++g) {
    h_regions[g] = {h_off[g], h_off[g] + h_count[g]};
    h_begin[g] = (long long)h_off[g];
    h_end[g] = (long long)(h_off[g] + h_count[g]);
  }
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t cub_bytes = 0, t = 0;
  {
    unsigned long long *kp = nullptr;
    long long *op = nullptr;
    uint32_t *fp = nullptr;
    if (per_region) {
      cub::DeviceRadixSort::SortPairs(nullptr, t, kp, kp, kp, kp, (int)span, 0, 64, ctx->stream);
    } else {
      cub::DeviceSegmentedRadixSort::SortPairs(nullptr, t, kp, kp, kp, kp, (int)span, n_regions, op, op, 0, 64,
                                               ctx->stream);
    }
    cub_bytes = t;
    cub::DeviceScan::ExclusiveSum(nullptr, t, fp, fp, (int)span, ctx->stream);
    if (t > cub_bytes) cub_bytes = t;
  }
  const size_t sz_regions = align(sizeof(Region) * n_regions), sz_offs = align(sizeof(long long) * n_regions);
  const size_t sz_plane = align(8 * span), sz_keys = align(kb * span), sz_u32 = align(4 * span);
  const size_t n_planes = key_words == 1 ? 1 : 4;  // K64: alt keys; K128: lo, hi, lo', hi'
  const size_t need = sz_regions + 2 * sz_offs + n_planes * sz_plane + 2 * sz_u32 + 2 * sz_offs + align(cub_bytes);
  (void)sz_keys;
  char *base = nullptr;
  SKS_TRY(ctx_scratch(ctx, need, reinterpret_cast<void **>(&base)));
  size_t o = 0;
  Region *d_regions = reinterpret_cast<Region *>(base + o); o += sz_regions;
  long long *d_begin = reinterpret_cast<long long *>(base + o); o += sz_offs;
  long long *d_end = reinterpret_cast<long long *>(base + o); o += sz_offs;
  unsigned long long *plane[4];
  for (size_t i = 0; i < n_planes; ++i) { plane[i] = reinterpret_cast<unsigned long long *>(base + o); o += sz_plane; }
  uint32_t *d_flags = reinterpret_cast<uint32_t *>(base + o); o += sz_u32;
  uint32_t *d_pos = reinterpret_cast<uint32_t *>(base + o); o += sz_u32;
  unsigned long long *d_uoff = reinterpret_cast<unsigned long long *>(base + o); o += sz_offs;
  unsigned long long *d_ucount = reinterpret_cast<unsigned long long *>(base + o); o += sz_offs;
  void *d_cub = base + o;

  SKS_CUDA_TRY(cudaMemcpyAsync(d_regions, h_regions.data(), sizeof(Region) * n_regions, cudaMemcpyHostToDevice, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(d_begin, h_begin.data(), sizeof(long long) * n_regions, cudaMemcpyHostToDevice, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(d_end, h_end.data(), sizeof(long long) * n_regions, cudaMemcpyHostToDevice, ctx->stream));

  KernelTimer timer(ctx, SKS_KERNEL_SORT_UNIQUE);
  unsigned long long *sorted = nullptr;  // [span] slots of key_words
  unsigned long long *d_keys = static_cast<unsigned long long *>(keys);
  const int nblk = (int)((span + 255) / 256);
  if (key_words == 1) {
    size_t tb = cub_bytes;
    if (per_region) {  // few large regions: a device-wide sort each (the segmented sort gives a segment to ONE CTA)
      for (int g = 0; g < n_regions; ++g) {
        tb = cub_bytes;
        if (h_count[g])
          cub::DeviceRadixSort::SortKeys(d_cub, tb, d_keys + h_off[g], plane[0] + h_off[g], (int)h_count[g], 0, 64, ctx->stream);
      }
    } else {
      cub::DeviceSegmentedRadixSort::SortKeys(d_cub, tb, d_keys, plane[0], (int)span, n_regions, d_begin, d_end, 0, 64,
                                              ctx->stream);
    }
    ctx->launches += 8;
    sorted = plane[0];
  } else {
    split_keys_kernel<<<nblk, 256, 0, ctx->stream>>>(static_cast<const ulonglong2 *>(keys), plane[0], plane[1], span);
    size_t tb = cub_bytes;
    if (per_region) {
      for (int g = 0; g < n_regions; ++g) {
        const uint64_t b = h_off[g];
        const int n = (int)h_count[g];
        if (!n) continue;
        tb = cub_bytes;
        cub::DeviceRadixSort::SortPairs(d_cub, tb, plane[0] + b, plane[2] + b, plane[1] + b, plane[3] + b, n, 0, 64, ctx->stream);
        tb = cub_bytes;
        cub::DeviceRadixSort::SortPairs(d_cub, tb, plane[3] + b, plane[1] + b, plane[2] + b, plane[0] + b, n, 0, 64, ctx->stream);
      }
    } else {
      cub::DeviceSegmentedRadixSort::SortPairs(d_cub, tb, plane[0], plane[2], plane[1], plane[3], (int)span, n_regions,
                                               d_begin, d_end, 0, 64, ctx->stream);
      tb = cub_bytes;
      cub::DeviceSegmentedRadixSort::SortPairs(d_cub, tb, plane[3], plane[1], plane[2], plane[0], (int)span, n_regions,
                                               d_begin, d_end, 0, 64, ctx->stream);
    }
    // now plane[0] = lo, plane[1] = hi, sorted by (hi, lo); re-interleave into the caller's buffer
    join_keys_kernel<<<nblk, 256, 0, ctx->stream>>>(plane[0], plane[1], static_cast<ulonglong2 *>(keys), span);
    ctx->launches += 18;
    sorted = d_keys;
  }
  SKS_CUDA_TRY(cudaGetLastError());

  SKS_CUDA_TRY(cudaMemsetAsync(d_flags, 0, 4 * span, ctx->stream));
  uint64_t max_count = 0;
  for (int g = 0; g < n_regions; ++g) max_count = h_count[g] > max_count ? h_count[g] : max_count;
  dim3 grid((unsigned)std::min<uint64_t>((max_count + 255) / 256, 65535 / 8), (unsigned)n_regions);
  if (grid.x < 1) grid.x = 1;
  if (key_words == 1) mark_heads_kernel<1><<<grid, 256, 0, ctx->stream>>>(sorted, d_regions, d_flags);
  else mark_heads_kernel<2><<<grid, 256, 0, ctx->stream>>>(sorted, d_regions, d_flags);
  {
    size_t tb = cub_bytes;
    cub::DeviceScan::ExclusiveSum(d_cub, tb, d_flags, d_pos, (int)span, ctx->stream);
  }
  SKS_TRY(alloc_buffer(ctx, kb * total, out_buf));
  unsigned long long *d_out = static_cast<unsigned long long *>((*out_buf)->ptr);
  if (key_words == 1)
    scatter_heads_kernel<1><<<grid, 256, 0, ctx->stream>>>(sorted, d_regions, d_flags, d_pos, d_out, d_uoff, d_ucount);
  else
    scatter_heads_kernel<2><<<grid, 256, 0, ctx->stream>>>(sorted, d_regions, d_flags, d_pos, d_out, d_uoff, d_ucount);
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches += 4;

  std::vector<unsigned long long> h_uoff(n_regions), h_ucount(n_regions);
  SKS_CUDA_TRY(cudaMemcpyAsync(h_uoff.data(), d_uoff, 8 * n_regions, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_ucount.data(), d_ucount, 8 * n_regions, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  for (int g = 0; g < n_regions; ++g) {
    (*out_off)[g] = h_uoff[g];
    (*out_count)[g] = h_ucount[g];
  }
  return SKS_OK;
}

// Imported keys (files, host lists, device buffers of other producers) are only trusted after this check: every key
// must be a subset of the mask (the row-resident intersection and the bucket sort index keys by their mask-selected
// bits) and, where the caller claims sorted distinct sets, strictly ascending inside every set.
template <int KW>
__global__ void __launch_bounds__(256)
    validate_keys_kernel(const unsigned long long *__restrict__ keys, unsigned long long n, unsigned long long nm_lo,
                         unsigned long long nm_hi, int need_sorted, const long long *__restrict__ starts, int n_starts,
                         uint32_t *__restrict__ flag) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long lo = keys[KW * i], hi = KW == 2 ? keys[KW * i + 1] : 0ull;
    if ((lo & nm_lo) || (hi & nm_hi)) atomicOr(flag, 1u);
    if (need_sorted && i > 0) {
      const unsigned long long plo = keys[KW * (i - 1)], phi = KW == 2 ? keys[KW * (i - 1) + 1] : 0ull;
      const bool less = phi != hi ? phi < hi : plo < lo;
      if (!less) {  // allowed only where a new set starts
        int a = 0, b = n_starts;
        bool is_start = false;
        while (a < b) {
          const int m = (a + b) >> 1;
          const long long v = starts[m];
          if (v == (long long)i) { is_start = true; break; }
          if (v < (long long)i) a = m + 1; else b = m;
        }
        if (!is_start) atomicOr(flag, 2u);
      }
    }
  }
}

// h_starts: first key of every set (ascending), only used with need_sorted.  Synchronises the stream.
int validate_keys(sks_ctx *ctx, const void *d_keys, int64_t n, int key_words, const uint64_t mask[2], bool need_sorted,
                  const int64_t *h_starts, int n_starts) {
  if (n <= 0) return SKS_OK;
  char *scratch = nullptr;
  const size_t sz_starts = ((size_t)std::max(n_starts, 1) * 8 + 255) & ~(size_t)255;
  SKS_TRY(ctx_scratch(ctx, sz_starts + 256, reinterpret_cast<void **>(&scratch)));
  long long *d_starts = reinterpret_cast<long long *>(scratch);
  uint32_t *d_flag = reinterpret_cast<uint32_t *>(scratch + sz_starts);
  uint32_t *h_flag = nullptr;
  SKS_TRY(ctx_pinned(ctx, sz_starts + 64, reinterpret_cast<void **>(&h_flag)));
  if (need_sorted && n_starts > 0) {
    memcpy(h_flag + 16, h_starts, (size_t)n_starts * 8);
    SKS_CUDA_TRY(cudaMemcpyAsync(d_starts, h_flag + 16, (size_t)n_starts * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  SKS_CUDA_TRY(cudaMemsetAsync(d_flag, 0, 4, ctx->stream));
  const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
  const unsigned long long *k = static_cast<const unsigned long long *>(d_keys);
  if (key_words == 1)
    validate_keys_kernel<1><<<grid, 256, 0, ctx->stream>>>(k, (unsigned long long)n, ~mask[0], ~0ull, need_sorted ? 1 : 0, d_starts,
                                                           need_sorted ? n_starts : 0, d_flag);
  else
    validate_keys_kernel<2><<<grid, 256, 0, ctx->stream>>>(k, (unsigned long long)n, ~mask[0], ~mask[1], need_sorted ? 1 : 0, d_starts,
                                                           need_sorted ? n_starts : 0, d_flag);
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches++;
  SKS_CUDA_TRY(cudaMemcpyAsync(h_flag, d_flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (*h_flag & 1u) return set_error(SKS_ERR_INVALID, "a key has bits outside the mask");
  if (*h_flag & 2u) return set_error(SKS_ERR_INVALID, "keys are not ascending and distinct");
  return SKS_OK;
}

// `slices` > 1: that many CTAs share every pair and add their partial counts (d_out must be zero on entry).
int launch_sorted_intersect_pairs(sks_ctx *ctx, int key_words, const void *const *d_a, const int64_t *d_na,
                                  const void *const *d_b, const int64_t *d_nb, int64_t n_pairs, int32_t *d_out,
                                  const uint32_t *d_pair_idx, int slices) {
  if (n_pairs == 0) return SKS_OK;
  KernelTimer timer(ctx, SKS_KERNEL_INTERSECT);
  if (slices < 1) slices = 1;
  for (int64_t done = 0; done < n_pairs;) {
    const int64_t chunk = std::min<int64_t>(n_pairs - done, 1 << 30);
    const uint32_t *idx = d_pair_idx ? d_pair_idx + done : nullptr;
    const int64_t base = d_pair_idx ? 0 : done;  // with an index list the tables are addressed through it
    const dim3 grid((unsigned)chunk, (unsigned)slices);
    if (key_words == 1)
      sorted_intersect_kernel<1><<<grid, kIntersectThreads, 0, ctx->stream>>>(
          d_a + base, reinterpret_cast<const long long *>(d_na + base), d_b + base,
          reinterpret_cast<const long long *>(d_nb + base), d_out + base, idx);
    else
      sorted_intersect_kernel<2><<<grid, kIntersectThreads, 0, ctx->stream>>>(
          d_a + base, reinterpret_cast<const long long *>(d_na + base), d_b + base,
          reinterpret_cast<const long long *>(d_nb + base), d_out + base, idx);
    SKS_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
    done += chunk;
  }
  return SKS_OK;
}

// ---- sorted-set intersection with the row set resident in shared memory ------------------------------
// All-vs-all compares one set against many.  A 1024-thread CTA loads the row set A (up to ~26 k 8-byte keys)
// into shared memory once, indexes it by the top 12 bits of the key (bucket starts), and then every warp
// takes one column set B at a time: its lanes stream B's keys from L2 (8 independent loads in flight per lane)
// and look each one up in A's bucket (a handful of LDS.64 instead of a merge step).  |A n B| is the number of
// hits.  Pairs whose row set does not fit, or rows with too few columns to pay for the load, go through
// sorted_intersect_kernel.
constexpr int kRowThreads = 1024;
constexpr int kRowWarps = kRowThreads / 32;
constexpr int kRowTableBits = 12;
constexpr int kRowBuckets = 1 << kRowTableBits;
constexpr int kRowSmemBytes = 225 * 1024;
constexpr int kRowUnroll = 8;

template <int KW>
struct RowKey;
template <>
struct RowKey<1> {
  unsigned long long v;
  __device__ __forceinline__ static RowKey load(const unsigned long long *p, uint32_t i) { return {p[i]}; }
  __device__ __forceinline__ uint32_t bucket(const SortPlan &p) const { return SortKey<1>::bucket(v, p); }
  __device__ __forceinline__ bool eq(const RowKey &o) const { return v == o.v; }
  __device__ __forceinline__ bool lt(const RowKey &o) const { return v < o.v; }
};
template <>
struct RowKey<2> {
  unsigned long long lo, hi;
  __device__ __forceinline__ static RowKey load(const unsigned long long *p, uint32_t i) { return {p[2 * i], p[2 * i + 1]}; }
  __device__ __forceinline__ uint32_t bucket(const SortPlan &p) const { return SortKey<2>::bucket(make_ulonglong2(lo, hi), p); }
  __device__ __forceinline__ bool eq(const RowKey &o) const { return lo == o.lo && hi == o.hi; }
  __device__ __forceinline__ bool lt(const RowKey &o) const { return hi != o.hi ? hi < o.hi : lo < o.lo; }
};

struct RowTask {
  const void *a;
  uint32_t n_a;
  uint32_t first;   // the task's columns are the pairs [first, first + n_cols) of the pair tables
  uint32_t n_cols;
  uint32_t pad;
};

template <int KW>
__global__ void __launch_bounds__(kRowThreads, 1)
    row_intersect_wide_kernel(const RowTask *__restrict__ tasks, const void *const *__restrict__ pb,
                         const long long *__restrict__ nb, int32_t *__restrict__ out, const __grid_constant__ SortPlan plan,
                         int tb) {
  using K = RowKey<KW>;
  extern __shared__ __align__(16) unsigned long long s_row[];
  const RowTask t = tasks[blockIdx.x];
  uint32_t *s_start = reinterpret_cast<uint32_t *>(s_row + (size_t)t.n_a * KW);  // [kRowBuckets + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  {
    const unsigned long long *A = static_cast<const unsigned long long *>(t.a);
    for (uint32_t i = tid; i < t.n_a * KW; i += kRowThreads) s_row[i] = __ldg(A + i);
  }
  __syncthreads();
  // bucket starts: s_start[b] = first index whose bucket is >= b
  for (uint32_t i = tid; i <= t.n_a; i += kRowThreads) {
    const uint32_t bi = i < t.n_a ? K::load(s_row, i).bucket(plan) : 1u << tb;
    const uint32_t bp = i > 0 ? K::load(s_row, i - 1).bucket(plan) + 1 : 0u;
    for (uint32_t b = bp; b <= bi; ++b) s_start[b] = i;
  }
  __syncthreads();

  auto hit = [&](const K &k) -> uint32_t {
    const uint32_t b = k.bucket(plan);
    uint32_t lo = s_start[b], hi = s_start[b + 1];
    while (hi - lo > 4) {  // a crowded bucket: halve [lo, hi), which keeps containing k's position if k is in A
      const uint32_t mid = (lo + hi) >> 1;
      if (K::load(s_row, mid).lt(k)) lo = mid + 1; else hi = mid + 1;
    }
    uint32_t f = 0;
    for (uint32_t p = lo; p < hi; ++p) f |= K::load(s_row, p).eq(k) ? 1u : 0u;
    return f;
  };

  // a unit = (column, part of its keys); with fewer than 32 columns several warps share one
  const uint32_t parts = t.n_cols >= (uint32_t)kRowWarps ? 1u : (uint32_t)kRowWarps / t.n_cols;
  for (uint32_t unit = warp; unit < t.n_cols * parts; unit += kRowWarps) {
    const uint32_t col = unit % t.n_cols, part = unit / t.n_cols;
    const unsigned long long *B = static_cast<const unsigned long long *>(pb[t.first + col]);
    const uint32_t n_b = (uint32_t)nb[t.first + col];
    const uint32_t per = ((n_b + parts - 1) / parts + 31) & ~31u;
    const uint32_t begin = part * per, end = begin + per < n_b ? begin + per : n_b;
    uint32_t cnt = 0;
    uint32_t i = begin + lane;
    for (; i + 32 * (kRowUnroll - 1) < end; i += 32 * kRowUnroll) {
      K k[kRowUnroll];
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u) k[u] = K::load(B, i + 32 * u);
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u) cnt += hit(k[u]);
    }
    for (; i < end; i += 32) cnt += hit(K::load(B, i));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if (lane == 0 && cnt) atomicAdd(out + t.first + col, (int32_t)cnt);
  }
}

// 8-byte keys: the row set is stored as two planes -- the 32 bits right below the bucket bits and the remaining
// low bits (16-bit when they fit) -- which leaves room for a bucket index on the top 13-15 key bits (16-bit
// starts, built from the sorted order without atomics).  A bucket then holds ~1 key: a lookup is two 16-bit
// table reads and one or two (hi, lo) compares, ~25 instructions instead of the ~65 of a merge step per key.
// All bit surgery is on 32-bit halves with launch-uniform shift amounts -- variable 64-bit shifts and masks made
// an earlier version of this kernel spend ~175 warp instructions per 32 lookups.  The bucket index is built
// from mask-SELECTED bits only (up to four runs of the mask's top set bits, concatenated): the positions a spaced
// seed skips are always zero in a key and would leave most buckets of a contiguous bit field empty.  Keys are
// subsets of the mask, so nothing lies above its highest bit.
struct RowPlan {
  int n_pieces;        // runs of mask bits that form the bucket index, all inside the key's upper half
  int s[4];            // piece i = ((khi >> s[i]) & m[i]) << o[i]  ==  (khi >> rs[i]) & pm[i]
  uint32_t m[4];
  int o[4];
  int rs[4];           // s[i] - o[i] (never negative: the index only closes gaps)
  uint32_t pm[4];      // m[i] << o[i]
  int tb;              // bucket bits
  int shift_low;       // lowest bucket bit (>= 32): the bits below it go to the two planes
};

template <typename LoT, int NP>  // NP = number of pieces of the bucket index (compile time: no selects)
__global__ void __launch_bounds__(kRowThreads, 1)
    row_intersect_kernel(const RowTask *__restrict__ tasks, const void *const *__restrict__ pb,
                         const long long *__restrict__ nb, int32_t *__restrict__ out, const __grid_constant__ RowPlan R) {
  extern __shared__ __align__(16) uint32_t s_u32[];
  const RowTask t = tasks[blockIdx.x];
  uint32_t *s_hi = s_u32;                                                    // [n_a] bits [shift_low-1 .. pshift]
  LoT *s_lo = reinterpret_cast<LoT *>(s_u32 + t.n_a);                        // [n_a] bits [pshift-1 .. 0]
  uint16_t *s_tab = reinterpret_cast<uint16_t *>(s_u32 + t.n_a + (t.n_a * sizeof(LoT) + 3) / 4);  // [2^tb + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pshift = R.shift_low - 32;  // 0 <= pshift <= 24
  const uint32_t n_buckets = 1u << R.tb;
  const uint32_t lo_mask = pshift ? (1u << pshift) - 1 : 0u;
  // (bucket, the 32 bits below the bucket bits, the rest) of a key given as two halves
  auto split = [&](uint32_t klo, uint32_t khi, uint32_t &b, uint32_t &h, uint32_t &l) {
    b = (khi >> R.rs[0]) & R.pm[0];
    if (NP > 1) b |= (khi >> R.rs[1]) & R.pm[1];
    if (NP > 2) b |= (khi >> R.rs[2]) & R.pm[2];
    if (NP > 3) b |= (khi >> R.rs[3]) & R.pm[3];
    h = __funnelshift_r(klo, khi, pshift);
    l = klo & lo_mask;
  };
  const uint2 *A = static_cast<const uint2 *>(t.a);
  for (uint32_t i = tid; i <= t.n_a; i += kRowThreads) {
    uint32_t bi = n_buckets, h, l;
    if (i < t.n_a) {
      const uint2 k = __ldg(A + i);
      split(k.x, k.y, bi, h, l);
      s_hi[i] = h;
      s_lo[i] = (LoT)l;
    }
    // s_tab[b] = first index whose bucket is >= b: this key owns the entries (bucket of its predecessor, bi]
    // (the bucket index is monotonic in the key: it concatenates the key's top mask bits in order)
    uint32_t bp = 0;
    if (i > 0) {
      const uint2 k = __ldg(A + i - 1);
      split(k.x, k.y, bp, h, l);
      ++bp;
    }
    for (uint32_t b = bp; b <= bi; ++b) s_tab[b] = (uint16_t)i;
  }
  __syncthreads();

  // a lookup in two steps, so that the table reads of a group of keys are in flight together
  struct Probe {
    uint32_t p, end, kh, kl;
  };
  auto prep = [&](uint2 k) -> Probe {
    uint32_t b;
    Probe q;
    split(k.x, k.y, b, q.kh, q.kl);
    q.p = s_tab[b];
    q.end = s_tab[b + 1];
    return q;
  };
  auto probe = [&](const Probe &q) -> uint32_t {
    // probe the first plane only (a bucket holds ~1 key; the warp runs the longest lane's bucket); the second
    // plane is compared once afterwards.  Two keys of one bucket that agree in the first plane are rare but
    // possible: then the bucket is walked again with both planes.
    uint32_t cand = 0, n_match = 0;
#pragma unroll 1  // an unrolled probe loop (the compiler's choice: by 8) is all overhead
    for (uint32_t p = q.p; p < q.end; ++p) {
      const bool m = s_hi[p] == q.kh;
      cand = m ? p : cand;
      n_match += m ? 1u : 0u;
    }
    if (n_match == 0) return 0u;
    if (n_match == 1) return s_lo[cand] == (LoT)q.kl ? 1u : 0u;
    uint32_t f = 0;
    for (uint32_t p = q.p; p < q.end; ++p) f |= (s_hi[p] == q.kh && s_lo[p] == (LoT)q.kl) ? 1u : 0u;
    return f;
  };

  const uint32_t parts = t.n_cols >= (uint32_t)kRowWarps ? 1u : (uint32_t)kRowWarps / t.n_cols;
  for (uint32_t unit = warp; unit < t.n_cols * parts; unit += kRowWarps) {
    const uint32_t col = unit % t.n_cols, part = unit / t.n_cols;
    const uint2 *B = static_cast<const uint2 *>(pb[t.first + col]);
    const uint32_t n_b = (uint32_t)nb[t.first + col];
    const uint32_t per = ((n_b + parts - 1) / parts + 31) & ~31u;
    const uint32_t begin = part * per, end = begin + per < n_b ? begin + per : n_b;
    uint32_t cnt = 0;
    uint32_t i = begin + lane;
    for (; i + 32 * (kRowUnroll - 1) < end; i += 32 * kRowUnroll) {
      uint2 k[kRowUnroll];
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u) k[u] = __ldg(B + i + 32 * u);
#pragma unroll
      for (int h = 0; h < kRowUnroll; h += 4) {
        Probe q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = prep(k[h + u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) cnt += probe(q[u]);
      }
    }
    for (; i < end; i += 32) cnt += probe(prep(__ldg(B + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if (lane == 0 && cnt) atomicAdd(out + t.first + col, (int32_t)cnt);
  }
}

// Bucket plan of the resident kernel for 8-byte keys under `mask`, for rows of up to n_a keys: the top mask bits
// (at most 15, in at most four runs, all at or above bit 32) index the buckets.  Returns false when a row does
// not fit shared memory or the mask's top bits lie too low / are too fragmented to be worth it.
bool row_plan(uint64_t mask, int64_t n_a, RowPlan *plan, bool *lo16) {
  if (n_a > 65535 || mask == 0) return false;
  for (int want = 15; want >= 10; --want) {
    RowPlan p = {};
    int bit = 63, got = 0;
    while (got < want && p.n_pieces < 4 && bit >= 32) {
      while (bit >= 32 && !((mask >> bit) & 1)) --bit;   // next run of set bits, from the top
      if (bit < 32) break;
      int hi = bit;
      while (bit >= 32 && ((mask >> bit) & 1) && hi - bit + 1 <= want - got) --bit;
      const int len = hi - bit;  // bits (bit, hi]
      p.s[p.n_pieces] = bit + 1 - 32;
      p.m[p.n_pieces] = (1u << len) - 1;
      got += len;
      p.o[p.n_pieces] = want - got;  // filled from the top of the index downwards
      p.rs[p.n_pieces] = p.s[p.n_pieces] - p.o[p.n_pieces];
      p.pm[p.n_pieces] = p.m[p.n_pieces] << p.o[p.n_pieces];
      if (p.rs[p.n_pieces] < 0) { got = -1; break; }  // cannot happen (see RowPlan); refuse the plan if it does
      ++p.n_pieces;
      p.shift_low = bit + 1;
    }
    if (got != want) continue;  // not enough mask bits in the upper half within four runs
    p.tb = want;
    const int lo_bits = p.shift_low - 32;
    if (lo_bits > 24) continue;
    const bool l16 = lo_bits <= 16;
    const int64_t bytes = n_a * 4 + ((n_a * (l16 ? 2 : 4) + 3) & ~(int64_t)3) + 2 * (((int64_t)1 << want) + 1) + 16;
    if (bytes > kRowSmemBytes) continue;
    *plan = p;
    *lo16 = l16;
    return true;
  }
  return false;
}

// Can a row of n_a keys be made resident (8-byte keys: depends on the mask too)?
bool row_intersect_fits(int key_words, const uint64_t mask[2], int64_t n_a) {
  if (key_words == 1) {
    RowPlan p;
    bool lo16;
    return row_plan(mask[0], n_a, &p, &lo16);
  }
  return n_a <= ((int64_t)kRowSmemBytes - (kRowBuckets + 1) * 4 - 64) / (8 * key_words);
}

// All tasks of one launch share the plan of the largest row (`max_n_a`).
int launch_row_intersect(sks_ctx *ctx, int key_words, const void *d_tasks, int64_t n_tasks, const void *const *d_b,
                         const int64_t *d_nb, int32_t *d_out, const uint64_t mask[2], int64_t max_n_a) {
  if (n_tasks == 0) return SKS_OK;
  KernelTimer timer(ctx, SKS_KERNEL_INTERSECT);
  const RowTask *tasks = static_cast<const RowTask *>(d_tasks);
  const long long *nb = reinterpret_cast<const long long *>(d_nb);
  if (key_words == 1) {
    RowPlan plan;
    bool lo16 = true;
    if (!row_plan(mask[0], max_n_a, &plan, &lo16)) return set_error(SKS_ERR_INVALID, "row set does not fit the resident kernel");
    auto go = [&](auto kernel) -> int {
      SKS_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmemBytes));
      kernel<<<(unsigned)n_tasks, kRowThreads, kRowSmemBytes, ctx->stream>>>(tasks, d_b, nb, d_out, plan);
      return SKS_OK;
    };
#define SKS_ROW_GO(NP)                                                    \
  case NP:                                                                \
    if (lo16) SKS_TRY(go(row_intersect_kernel<uint16_t, NP>));            \
    else SKS_TRY(go(row_intersect_kernel<uint32_t, NP>));                 \
    break;
    switch (plan.n_pieces) {
      SKS_ROW_GO(1)
      SKS_ROW_GO(2)
      SKS_ROW_GO(3)
      default:
        SKS_ROW_GO(4)
    }
#undef SKS_ROW_GO
  } else {
    // bucket index: the top (up to 12) mask-selected key bits
    int tb = std::min(kRowTableBits, __builtin_popcountll(mask[0]) + __builtin_popcountll(mask[1]));
    SortPlan plan = {};
    while (tb > 0 && !sort_plan(mask, tb, &plan)) --tb;
    if (tb == 0) plan = SortPlan{};
    SKS_CUDA_TRY(cudaFuncSetAttribute(row_intersect_wide_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmemBytes));
    row_intersect_wide_kernel<2><<<(unsigned)n_tasks, kRowThreads, kRowSmemBytes, ctx->stream>>>(tasks, d_b, nb, d_out, plan, tb);
  }
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches++;
  return SKS_OK;
}

}  // namespace sks
