// extern "C" entry points of libsks.so that touch the device: contexts, batches, sketching, sets,
// intersections.  Host-only entry points live in sks_host.cpp.  See include/sks.h for the contract
// and the reference interfaces each call replaces.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
#include <new>
#include <string>

#include "sks_internal.cuh"

namespace sks {

// Device blocks (bitsets, index regions, key buffers, control blocks) are recycled by exact size per (device, stream) instead
// of going back to the stream-ordered pool: after a mix of small and large requests the pool satisfies a
// 1 GiB cudaMallocAsync by remapping physical chunks, which was measured at 2-85 ms per call.  Reuse on the
// same stream is safe by stream order.  The cache of a stream is dropped when its context goes away.
namespace {
constexpr size_t kCacheMinBytes = 0;  // every size: the same few sizes recur call after call
constexpr size_t kCacheMaxBytes = (size_t)12 << 30;
struct BlockKey {
  int device;
  cudaStream_t stream;
  size_t bytes;
  bool operator<(const BlockKey &o) const {
    if (device != o.device) return device < o.device;
    if (stream != o.stream) return stream < o.stream;
    return bytes < o.bytes;
  }
};
struct BlockCache {
  std::mutex mu;
  std::multimap<BlockKey, void *> blocks;
  size_t cached = 0;
  // streams that belong to a live context (with the number of contexts using them).  A block can outlive the
  // context whose stream it was allocated on (a kmer_set sketched by a worker thread that has exited): such a
  // block must neither be cached under nor freed in the order of a stream that no longer exists.
  std::map<std::pair<int, cudaStream_t>, int> live;
};
BlockCache &block_cache() {
  static BlockCache *c = new BlockCache();  // never destroyed: the CUDA runtime may be gone at exit
  return *c;
}
void *cache_take(int device, cudaStream_t stream, size_t bytes) {
  BlockCache &c = block_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  auto it = c.blocks.find(BlockKey{device, stream, bytes});
  if (it == c.blocks.end()) return nullptr;
  void *p = it->second;
  c.blocks.erase(it);
  c.cached -= bytes;
  return p;
}
bool cache_put(int device, cudaStream_t stream, size_t bytes, void *p) {
  if (bytes < kCacheMinBytes) return false;
  BlockCache &c = block_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  if (c.live.find({device, stream}) == c.live.end()) return false;
  if (c.cached + bytes > kCacheMaxBytes) return false;
  c.blocks.emplace(BlockKey{device, stream, bytes}, p);
  c.cached += bytes;
  return true;
}
bool stream_is_live(int device, cudaStream_t stream) {
  BlockCache &c = block_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  return c.live.find({device, stream}) != c.live.end();
}
void stream_attach(int device, cudaStream_t stream) {
  BlockCache &c = block_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  ++c.live[{device, stream}];
}
void stream_detach(int device, cudaStream_t stream) {
  BlockCache &c = block_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  auto it = c.live.find({device, stream});
  if (it != c.live.end() && --it->second <= 0) c.live.erase(it);
}
}  // namespace

// Frees the cached blocks of one stream (all streams of the device when `stream_only` is false).
void cache_release(int device, cudaStream_t stream, bool stream_only) {
  BlockCache &c = block_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  for (auto it = c.blocks.begin(); it != c.blocks.end();) {
    if (it->first.device == device && (!stream_only || it->first.stream == stream)) {
      if (cudaFreeAsync(it->second, it->first.stream) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(it->second);
      }
      c.cached -= it->first.bytes;
      it = c.blocks.erase(it);
    } else {
      ++it;
    }
  }
}

DeviceBuffer::~DeviceBuffer() {
  if (ptr) {
    if (cache_put(device, stream, bytes, ptr)) return;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != device) cudaSetDevice(device);
    // the owning context (and with it the stream) may be gone: cudaFree waits for the device and is always valid
    if (!stream_is_live(device, stream) || cudaFreeAsync(ptr, stream) != cudaSuccess) {
      cudaGetLastError();
      cudaFree(ptr);
    }
    if (cur != device) cudaSetDevice(cur);
  }
}

int alloc_buffer(sks_ctx *ctx, size_t bytes, BufferRef *out) {
  auto buf = std::make_shared<DeviceBuffer>();
  if (bytes == 0) bytes = 16;
  if (bytes >= kCacheMinBytes) buf->ptr = cache_take(ctx->device, ctx->stream, bytes);
  if (!buf->ptr) {
    cudaError_t e = cudaMallocAsync(&buf->ptr, bytes, ctx->stream);
    if (e == cudaErrorMemoryAllocation) {  // give the recycled blocks back and try once more
      cudaGetLastError();
      cache_release(ctx->device, nullptr, false);
      cudaStreamSynchronize(ctx->stream);
      e = cudaMallocAsync(&buf->ptr, bytes, ctx->stream);
    }
    if (e != cudaSuccess) {
      buf->ptr = nullptr;
      return set_error(SKS_ERR_CUDA, "cudaMallocAsync of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
  }
  buf->bytes = bytes;
  buf->device = ctx->device;
  buf->stream = ctx->stream;
  *out = buf;
  return SKS_OK;
}

int ctx_scratch(sks_ctx *ctx, size_t bytes, void **out) {
  if (bytes > ctx->scratch_bytes) {
    if (ctx->scratch) SKS_CUDA_TRY(cudaFreeAsync(ctx->scratch, ctx->stream));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    SKS_CUDA_TRY(cudaMallocAsync(&ctx->scratch, want, ctx->stream));
    ctx->scratch_bytes = want;
  }
  *out = ctx->scratch;
  return SKS_OK;
}

// Pinned staging is a ring: blocks handed out since the last wrap stay untouched, so several
// asynchronous H2D parameter uploads can be in flight without a stream sync between them.
int ctx_pinned(sks_ctx *ctx, size_t bytes, void **out) {
  bytes = (bytes + 63) & ~(size_t)63;
  if (bytes > ctx->pinned_bytes / 2) {
    if (ctx->pinned) {
      SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      SKS_CUDA_TRY(cudaFreeHost(ctx->pinned));
    }
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    const size_t want = std::max<size_t>(2 * bytes, (size_t)1 << 20);
    SKS_CUDA_TRY(cudaMallocHost(&ctx->pinned, want));
    ctx->pinned_bytes = want;
    ctx->pinned_off = 0;
  }
  if (ctx->pinned_off + bytes > ctx->pinned_bytes) {
    SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // everything staged so far has been consumed
    ctx->pinned_off = 0;
  }
  *out = static_cast<char *>(ctx->pinned) + ctx->pinned_off;
  ctx->pinned_off += bytes;
  return SKS_OK;
}

static cudaEvent_t take_event(sks_ctx *ctx) {
  if (!ctx->event_pool.empty()) {
    cudaEvent_t e = ctx->event_pool.back();
    ctx->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
KernelTimer::KernelTimer(sks_ctx *c, int k) : ctx(c), kind(k) {
  if (!ctx->profile) return;
  e0 = take_event(ctx);
  e1 = take_event(ctx);
  cudaEventRecord(e0, ctx->stream);
}
KernelTimer::~KernelTimer() {
  if (!e0) return;
  cudaEventRecord(e1, ctx->stream);
  ctx->prof[kind].emplace_back(e0, e1);
}

namespace {

inline uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

// Lays genomes out in one word buffer: [kPreWords zero][genome 0 data, zero padded to 4 words]
// [kPreWords zero][genome 1 ...] ...  Fills h_genomes / h_seg_end / n_tiles; returns total words.
int layout_batch(sks_batch *b, int n_genomes, const uint64_t *n_bases, const uint64_t *const *seg_len,
                 const uint64_t *n_segs, uint64_t *total_words) {
  b->n_genomes = n_genomes;
  b->h_genomes.resize(n_genomes);
  b->h_seg_end.clear();
  uint64_t off = 0, tiles = 0;
  b->total_bases = 0;
  for (int g = 0; g < n_genomes; ++g) {
    if (n_bases[g] >= 0xFFFF0000ull)
      return set_error(SKS_ERR_INVALID, "genome %d has %llu bases; the limit per genome is 2^32 - 2^16", g,
                       (unsigned long long)n_bases[g]);
    GenomeDesc &gd = b->h_genomes[g];
    off += kPreWords;
    gd.word_off = off;
    gd.n_bases = (uint32_t)n_bases[g];
    gd.n_words = (uint32_t)round_up((n_bases[g] + 15) / 16, 4);
    off += gd.n_words;
    gd.seg_first = (uint32_t)b->h_seg_end.size();
    const uint64_t ns = (seg_len && seg_len[g] && n_segs) ? n_segs[g] : 0;
    if (ns == 0) {
      b->h_seg_end.push_back(gd.n_bases);
      gd.n_segs = 1;
    } else {
      uint64_t e = 0;
      for (uint64_t s = 0; s < ns; ++s) {
        e += seg_len[g][s];
        b->h_seg_end.push_back((uint32_t)e);
      }
      if (e != n_bases[g])
        return set_error(SKS_ERR_INVALID, "genome %d: segment lengths sum to %llu, expected %llu bases", g,
                         (unsigned long long)e, (unsigned long long)n_bases[g]);
      gd.n_segs = (uint32_t)ns;
    }
    gd.tile_first = (uint32_t)tiles;
    gd.n_tiles = (uint32_t)((n_bases[g] + kTileWindows - 1) / kTileWindows);
    tiles += gd.n_tiles;
    b->total_bases += n_bases[g];
  }
  off += kPreWords + kPostWords;
  if (tiles >= 0xFFFFFFFFull) return set_error(SKS_ERR_INVALID, "batch too large (%llu tiles)", (unsigned long long)tiles);
  b->n_tiles = (uint32_t)tiles;
  *total_words = off;
  return SKS_OK;
}

int upload_tables(sks_ctx *ctx, sks_batch *b) {
  // descriptor tables go through the pinned ring, so that nothing here waits for the stream
  const int G = b->n_genomes;
  const size_t sz_g = sizeof(GenomeDesc) * (size_t)G, sz_s = 4 * b->h_seg_end.size();
  const size_t sz_t = (G > 1) ? 4 * (size_t)b->n_tiles : 0;
  SKS_TRY(alloc_buffer(ctx, std::max<size_t>(sz_g, 16), &b->genomes));
  SKS_TRY(alloc_buffer(ctx, std::max<size_t>(sz_s, 16), &b->seg_end));
  char *stage = nullptr;
  SKS_TRY(ctx_pinned(ctx, sz_g + sz_s + sz_t + 64, reinterpret_cast<void **>(&stage)));
  if (G > 0) {
    memcpy(stage, b->h_genomes.data(), sz_g);
    SKS_CUDA_TRY(cudaMemcpyAsync(b->genomes->ptr, stage, sz_g, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (sz_s) {
    memcpy(stage + sz_g, b->h_seg_end.data(), sz_s);
    SKS_CUDA_TRY(cudaMemcpyAsync(b->seg_end->ptr, stage + sz_g, sz_s, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (sz_t) {
    uint32_t *tg = reinterpret_cast<uint32_t *>(stage + sz_g + sz_s);
    for (int g = 0; g < G; ++g)
      std::fill(tg + b->h_genomes[g].tile_first, tg + b->h_genomes[g].tile_first + b->h_genomes[g].n_tiles, (uint32_t)g);
    SKS_TRY(alloc_buffer(ctx, sz_t, &b->tile_genome));
    SKS_CUDA_TRY(cudaMemcpyAsync(b->tile_genome->ptr, tg, sz_t, cudaMemcpyHostToDevice, ctx->stream));
  }
  return SKS_OK;
}

// Number of windows a genome can yield: sum over segments of max(0, len - w + 1)
// (src/kmer_sliding.cpp:121-125,144).
uint64_t genome_windows(const sks_batch *b, int g, int w) {
  const GenomeDesc &gd = b->h_genomes[g];
  uint64_t total = 0, prev = 0;
  for (uint32_t s = 0; s < gd.n_segs; ++s) {
    const uint64_t e = b->h_seg_end[gd.seg_first + s], len = e - prev;
    if (len >= (uint64_t)w) total += len - w + 1;
    prev = e;
  }
  return total;
}

inline int mask_top_bit(const uint64_t mask[2]) {  // -1 for an empty mask
  if (mask[1]) return 127 - __builtin_clzll(mask[1]);
  return mask[0] ? 63 - __builtin_clzll(mask[0]) : -1;
}

struct SketchPlan {
  int n_limbs, pred_mode, weight;
  SketchParams p;
};

int make_plan(const sks_batch *batch, const uint64_t mask[2], int window, const sks_pred *pred, SketchPlan *plan) {
  if (!batch || !mask || !pred) return set_error(SKS_ERR_INVALID, "null argument");
  if (window < 1 || window > 64) return set_error(SKS_ERR_INVALID, "window length %d outside 1..64", window);
  if (window < 64) {
    const unsigned __int128 m = ((unsigned __int128)mask[1] << 64) | mask[0];
    if (m >> (2 * window)) return set_error(SKS_ERR_INVALID, "mask has bits at or above 2*window");
  }
  memset(&plan->p, 0, sizeof(plan->p));
  SketchParams &p = plan->p;
  plan->n_limbs = (2 * window + 31) / 32;
  plan->weight = sks_mask_weight(mask);
  p.words = batch->host_words ? nullptr : static_cast<const uint32_t *>(batch->words->ptr);
  p.host_words = batch->host_words ? 1u : 0u;
  p.genomes = static_cast<const GenomeDesc *>(batch->genomes->ptr);
  p.seg_end = static_cast<const uint32_t *>(batch->seg_end->ptr);
  p.n_genomes = batch->n_genomes;
  p.n_tiles = batch->n_tiles;
  p.window = window;
  p.mask[0] = (uint32_t)mask[0];
  p.mask[1] = (uint32_t)(mask[0] >> 32);
  p.mask[2] = (uint32_t)mask[1];
  p.mask[3] = (uint32_t)(mask[1] >> 32);
  if (pred->kind == SKS_PRED_ALL) {
    plan->pred_mode = PRED_ALL;
  } else if (pred->kind == SKS_PRED_FMH) {
    if (pred->modulus == 0) return set_error(SKS_ERR_INVALID, "FracMinHash modulus must be non-zero");
    const int variant = pred->hash_variant == 0 ? SKS_HASH_BOOST_181 : pred->hash_variant;
    if (variant != SKS_HASH_BOOST_171 && variant != SKS_HASH_BOOST_181)
      return set_error(SKS_ERR_INVALID, "unknown hash variant %d", pred->hash_variant);
    plan->pred_mode = variant == SKS_HASH_BOOST_171 ? PRED_FMH171 : PRED_FMH181;
    // frac_min_hash::operator(), src/kmer.hpp:146: everything but H(masked_bits) is constant per launch
    p.hconst = boost_hash_bitset(mask[0], mask[1], variant) ^ (uint64_t)(int64_t)window ^ (uint64_t)(int64_t)pred->nonce;
    int mshift = 0;
    modulus_magic(pred->modulus, &p.minv, &p.mbound, &mshift);
    p.mbound <<= mshift;  // floor((2^64-1) / modulus) < 2^(64-s): no overflow
    p.mlow = (1ull << mshift) - 1;  // a non-zero modulus has s <= 63
  } else {
    return set_error(SKS_ERR_INVALID, "unknown predicate kind %d", pred->kind);
  }
  return SKS_OK;
}

}  // namespace
// SKS_REPR_AUTO: the presence bitset pays off for an unfiltered sketch when it is not much larger than the genome's
// own k-mer list (a 4^16-bit set is 512 MiB: right for a 5 Mbp genome, absurd for a 100 kbp one) and the bitsets of
// the whole batch fit a memory budget; everything else becomes sorted distinct keys.
int auto_repr(const sks_batch *batch, const SketchPlan &plan, const sks_pred *pred, int window) {
  if (pred->kind != SKS_PRED_ALL || plan.weight > 16) return SKS_REPR_SORTED;
  const uint64_t bitset_bytes = std::max<uint64_t>(((uint64_t)1 << (2 * plan.weight)) / 8, 4);
  uint64_t max_windows = 0;
  for (int g = 0; g < batch->n_genomes; ++g) max_windows = std::max(max_windows, genome_windows(batch, g, window));
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
    cudaGetLastError();
    free_b = (size_t)8 << 30;
  }
  const uint64_t budget = std::min<uint64_t>(free_b / 2, (uint64_t)32 << 30);
  if (bitset_bytes > 65536 && bitset_bytes > 128 * max_windows) return SKS_REPR_SORTED;
  if (bitset_bytes * (uint64_t)std::max(batch->n_genomes, 1) > budget) return SKS_REPR_SORTED;
  return SKS_REPR_BITSET;
}

sks_set *new_set(const sks_ctx *ctx, int repr, const uint64_t mask[2], int window, int weight) {
  sks_set *s = new (std::nothrow) sks_set();
  if (!s) return nullptr;
  s->device = ctx->device;
  s->repr = repr;
  s->window = window;
  s->weight = weight;
  s->mask[0] = mask[0];
  s->mask[1] = mask[1];
  return s;
}
namespace {

int sketch_raw_keys(sks_ctx *ctx, const sks_batch *batch, SketchPlan &plan, const sks_pred *pred, int window,
                    int out_mode, BufferRef *keys, BufferRef *pos, std::vector<uint64_t> *off,
                    std::vector<uint64_t> *count, uint64_t *span);

// Request of the one-call pair pipeline: count |A|, |B|, |A n B| inside the bitset build of a 2-genome batch.
struct PairFuse {
  bool store = true;   // materialise both bitsets in HBM (false: the slices never leave shared memory)
  bool done = false;   // the fused kernel ran and counts[] is valid
  int64_t counts[3] = {0, 0, 0};
};

int sketch_bitset(sks_ctx *ctx, const sks_batch *batch, SketchPlan &plan, const sks_pred *pred, const uint64_t mask[2],
                  int window, sks_set **out_sets, PairFuse *fuse = nullptr) {
  int index_bits = 0;
  SKS_TRY(build_pext_table(mask, plan.n_limbs, &plan.p.pext, &index_bits));
  const int G = batch->n_genomes;
  const uint64_t bits = 1ull << index_bits;           // 4^weight
  const uint64_t words = bits < 32 ? 1 : bits / 32;   // per genome
  BufferRef buf, count_buf;
  const uint32_t *tg = batch->tile_genome ? static_cast<const uint32_t *>(batch->tile_genome->ptr) : nullptr;
  const bool bucketed = index_bits > kSliceBits && index_bits >= ctx->bucket_min_bits;
  if (fuse && G != 2) fuse = nullptr;
  int64_t known_count[2] = {-1, -1};
  if (bucketed) {
    // One control block, zeroed by one memset: per-genome popcounts | |A n B| (pair pipeline) | overflow flag |
    // work-queue counter | per-(genome, bucket) cursors.  The sets keep it alive for their sizes; one 32-byte
    // copy brings counts and flag back.
    const PartGeometry geo = part_geometry(index_bits);
    const size_t ctl_head = (sizeof(unsigned long long) * ((size_t)G + 2) + 63) & ~(size_t)63;
    const size_t ctl_bytes = ctl_head + 64 + 4 * (size_t)geo.n_parts * G;
    SKS_TRY(alloc_buffer(ctx, ctl_bytes, &count_buf));
    unsigned long long *d_set_count = static_cast<unsigned long long *>(count_buf->ptr);
    uint32_t *d_overflow = reinterpret_cast<uint32_t *>(d_set_count + G + 1);
    unsigned int *d_work = reinterpret_cast<unsigned int *>(static_cast<char *>(count_buf->ptr) + ctl_head);
    uint32_t *d_cursor = reinterpret_cast<uint32_t *>(static_cast<char *>(count_buf->ptr) + ctl_head + 64);
    // Fast path: the sketch kernel scatters the PEXT indices straight into fixed per-(genome, bucket) regions
    // sized at 4x the mean bucket load (K2/K3/K4a fused), then the slices are assembled (K4b).  A genome whose
    // index distribution overflows a region (heavy skew) is detected by a flag and redone through the exact
    // counting partition below.
    uint64_t max_windows = 0;
    for (int g = 0; g < G; ++g) max_windows = std::max(max_windows, genome_windows(batch, g, window));
    const uint64_t cap = 4 * ((max_windows + geo.n_parts - 1) / geo.n_parts) + 1024;
    const uint64_t slots = cap * geo.n_parts * (uint64_t)G;
    bool done = false;
    if (!ctx->exact_partition && cap < (1ull << 32) && slots * 4 <= (16ull << 30)) {
      BufferRef regions;
      SKS_TRY(alloc_buffer(ctx, (size_t)slots * 4, &regions));
      SKS_CUDA_TRY(cudaMemsetAsync(count_buf->ptr, 0, ctl_bytes, ctx->stream));
      plan.p.out_keys = regions->ptr;
      plan.p.part_cursor = d_cursor;
      plan.p.part_overflow = d_overflow;
      plan.p.n_parts = geo.n_parts;
      plan.p.part_cap = (uint32_t)cap;
      plan.p.part_shift = geo.part_shift;
      SKS_TRY(launch_sketch(ctx, plan.p, tg, plan.n_limbs, plan.pred_mode, OUT_PART));
      unsigned long long *h_back = nullptr;
      SKS_TRY(ctx_pinned(ctx, 64, reinterpret_cast<void **>(&h_back)));
      if (fuse) {
        // K4b + K5 fused: both genomes' slices side by side in shared memory, counted while they are assembled
        uint32_t *ba = nullptr, *bb = nullptr;
        if (fuse->store) {
          SKS_TRY(alloc_buffer(ctx, (size_t)words * 4 * G, &buf));
          ba = static_cast<uint32_t *>(buf->ptr);
          bb = ba + words;
        }
        SKS_TRY(launch_bitset_pair_build(ctx, static_cast<const uint32_t *>(regions->ptr), d_cursor, (uint32_t)cap, index_bits,
                                         ba, bb, d_set_count, d_work));
        SKS_CUDA_TRY(cudaMemcpyAsync(h_back, d_set_count, 32, cudaMemcpyDeviceToHost, ctx->stream));
        SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        done = (uint32_t)h_back[3] == 0;
        if (done) {
          fuse->done = true;
          for (int k = 0; k < 3; ++k) fuse->counts[k] = (int64_t)h_back[k];
          known_count[0] = fuse->counts[0];
          known_count[1] = fuse->counts[1];
          if (!fuse->store) return SKS_OK;  // nothing was materialised: there are no sets to hand out
        }
      } else {
        SKS_TRY(alloc_buffer(ctx, (size_t)words * 4 * G, &buf));
        SKS_TRY(launch_bitset_assemble(ctx, static_cast<const uint32_t *>(regions->ptr), d_cursor, (uint32_t)cap, G, index_bits,
                                       static_cast<uint32_t *>(buf->ptr), words, d_set_count, d_work));
        SKS_CUDA_TRY(cudaMemcpyAsync(h_back, d_overflow, 8, cudaMemcpyDeviceToHost, ctx->stream));
        SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        done = (uint32_t)h_back[0] == 0;
      }
    }
    if (!done) {
      // Exact path: K2/K3 emit the indices, a counting partition orders them by bucket, K4b assembles.
      BufferRef raw, pos, sorted;
      std::vector<uint64_t> off, count;
      uint64_t span = 0;
      if (!buf) SKS_TRY(alloc_buffer(ctx, (size_t)words * 4 * G, &buf));
      SKS_TRY(sketch_raw_keys(ctx, batch, plan, pred, window, OUT_INDEX, &raw, &pos, &off, &count, &span));
      SKS_TRY(alloc_buffer(ctx, (size_t)span * 4, &sorted));
      SKS_TRY(launch_bitset_build(ctx, static_cast<const uint32_t *>(raw->ptr), static_cast<uint32_t *>(sorted->ptr),
                                  off.data(), count.data(), G, index_bits, static_cast<uint32_t *>(buf->ptr), words,
                                  d_set_count));
    }
  } else {
    SKS_TRY(alloc_buffer(ctx, (size_t)words * 4 * G, &buf));
    SKS_TRY(launch_fill_zero(ctx, buf->ptr, (size_t)words * 4 * G));
    plan.p.bitset = static_cast<uint32_t *>(buf->ptr);
    plan.p.bitset_words = words;

    SKS_TRY(launch_sketch(ctx, plan.p, tg, plan.n_limbs, plan.pred_mode, OUT_BITSET));
  }
  for (int g = 0; g < G; ++g) {
    sks_set *s = new_set(ctx, SKS_REPR_BITSET, mask, window, plan.weight);
    if (!s) return set_error(SKS_ERR_INVALID, "out of host memory");
    s->buf = buf;
    s->byte_off = (size_t)g * words * 4;
    s->bitset_words = words;
    s->count = (fuse && fuse->done) ? known_count[g] : -1;
    s->count_buf = count_buf;
    s->count_off = (size_t)g * sizeof(unsigned long long);
    out_sets[g] = s;
  }
  return SKS_OK;
}

// Runs the sketch kernel in OUT_KEYS / OUT_LIST mode into per-genome regions; re-runs once with
// exact capacities when a region overflowed.  Leaves raw keys (unsorted, duplicated) in *keys.
int sketch_raw_keys(sks_ctx *ctx, const sks_batch *batch, SketchPlan &plan, const sks_pred *pred, int window,
                    int out_mode, BufferRef *keys, BufferRef *pos, std::vector<uint64_t> *off,
                    std::vector<uint64_t> *count, uint64_t *span) {
  const int G = batch->n_genomes;
  const size_t key_bytes = out_mode == OUT_INDEX ? 4 : (plan.n_limbs <= 2 ? 8 : 16);
  std::vector<uint64_t> cap(G);
  for (int g = 0; g < G; ++g) {
    const uint64_t wins = genome_windows(batch, g, window);
    uint64_t c = wins;
    if (plan.pred_mode != PRED_ALL && pred->modulus > 1) {
      const double expect = (double)wins / (double)pred->modulus;
      c = std::min<uint64_t>(wins, (uint64_t)(expect * 1.25) + 1024);
    }
    cap[g] = c;
  }
  off->assign(G, 0);
  count->assign(G, 0);
  for (int attempt = 0; attempt < 2; ++attempt) {
    uint64_t total = 0;
    for (int g = 0; g < G; ++g) {
      (*off)[g] = total;
      total += cap[g];
    }
    *span = total;
    SKS_TRY(alloc_buffer(ctx, (size_t)total * key_bytes, keys));
    if (out_mode == OUT_LIST) SKS_TRY(alloc_buffer(ctx, (size_t)total * 4, pos));
    // device tables: off | cap | count
    char *tab = nullptr;
    SKS_TRY(ctx_scratch(ctx, (size_t)G * 24 + 64, reinterpret_cast<void **>(&tab)));
    uint64_t *d_off = reinterpret_cast<uint64_t *>(tab), *d_cap = d_off + G;
    unsigned long long *d_count = reinterpret_cast<unsigned long long *>(d_cap + G);
    char *stage = nullptr;
    SKS_TRY(ctx_pinned(ctx, (size_t)G * 24, reinterpret_cast<void **>(&stage)));
    memcpy(stage, off->data(), (size_t)G * 8);
    memcpy(stage + (size_t)G * 8, cap.data(), (size_t)G * 8);
    memset(stage + (size_t)G * 16, 0, (size_t)G * 8);
    SKS_CUDA_TRY(cudaMemcpyAsync(tab, stage, (size_t)G * 24, cudaMemcpyHostToDevice, ctx->stream));
    plan.p.out_keys = (*keys)->ptr;
    plan.p.out_pos = out_mode == OUT_LIST ? static_cast<uint32_t *>((*pos)->ptr) : nullptr;
    plan.p.out_off = d_off;
    plan.p.out_cap = d_cap;
    plan.p.out_count = d_count;
    SKS_TRY(launch_sketch(ctx, plan.p, batch->tile_genome ? static_cast<const uint32_t *>(batch->tile_genome->ptr) : nullptr,
                          plan.n_limbs, plan.pred_mode, out_mode));
    if (plan.pred_mode == PRED_ALL) {  // every window is kept: the counts are known without a read-back
      for (int g = 0; g < G; ++g) (*count)[g] = cap[g];
      return SKS_OK;
    }
    SKS_CUDA_TRY(cudaMemcpyAsync(stage, d_count, (size_t)G * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memcpy(count->data(), stage, (size_t)G * 8);
    bool overflow = false;
    for (int g = 0; g < G; ++g)
      if ((*count)[g] > cap[g]) overflow = true;
    if (!overflow) return SKS_OK;
    for (int g = 0; g < G; ++g) cap[g] = std::max(cap[g], (*count)[g]);  // counts are exact even past cap
  }
  return set_error(SKS_ERR_CAPACITY, "sketch output overflowed twice");
}

}  // namespace
// The kept k-mers of a single-genome batch as they leave the sketch kernel: unsorted, duplicates included.
int sketch_raw_one(sks_ctx *ctx, const sks_batch *batch, const uint64_t mask[2], int window, const sks_pred *pred,
                   BufferRef *keys, uint64_t *count, int *key_words) {
  if (!batch || batch->n_genomes != 1) return set_error(SKS_ERR_INVALID, "a single-genome batch is needed");
  if (batch->device != ctx->device) return set_error(SKS_ERR_INVALID, "batch lives on another device");
  SketchPlan plan;
  SKS_TRY(make_plan(batch, mask, window, pred, &plan));
  BufferRef pos;
  std::vector<uint64_t> off, cnt;
  uint64_t span = 0;
  SKS_TRY(sketch_raw_keys(ctx, batch, plan, pred, window, OUT_KEYS, keys, &pos, &off, &cnt, &span));
  *count = cnt[0];
  *key_words = plan.n_limbs <= 2 ? 1 : 2;
  return SKS_OK;
}
namespace {

// The context's stream waits for the last copy of a batch that is still arriving (no-op for every other batch).
// Chunk c of a batch that is still arriving: its copy and its event are queued (pageable sources: the feeder thread
// gets there in its own time), the context's stream waits for the event.
int batch_chunk_ready(sks_ctx *ctx, const sks_batch *batch, size_t c) {
  if (batch->feed) {
    sks_batch::Feed *feed = batch->feed.get();
    std::unique_lock<std::mutex> lk(feed->m);
    feed->cv.wait(lk, [&] { return feed->n_queued > c || feed->status != SKS_OK; });
    if (feed->status != SKS_OK) return set_error(feed->status, "copying the host genomes to the device failed");
  }
  SKS_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, batch->arriving[c].ready, 0));
  return SKS_OK;
}
int batch_arrived(sks_ctx *ctx, const sks_batch *batch) {
  return batch->arriving.empty() ? (int)SKS_OK : batch_chunk_ready(ctx, batch, batch->arriving.size() - 1);
}

int sketch_sorted(sks_ctx *ctx, const sks_batch *batch, SketchPlan &plan, const sks_pred *pred,
                  const uint64_t mask[2], int window, sks_set **out_sets) {
  const int G = batch->n_genomes;
  const int key_words = plan.n_limbs <= 2 ? 1 : 2;
  BufferRef raw, pos, uniq;
  std::vector<uint64_t> off, count, uoff, ucount;
  uint64_t span = 0;
  bool done = false;
  // (a context whose last attempt overflowed skips the next 63: the genomes of one workload tend to look alike, and a
  // failed attempt costs a second pass of the sketch kernel)
  if (ctx->kpart_skip > 0) {
    --ctx->kpart_skip;
  } else if (key_words == 1 && plan.pred_mode != PRED_ALL && pred->modulus > 1) {
    // A sparse condition on 8-byte keys: the first level of the bucket sort (histogram + scatter by the keys' top mask
    // bits) is folded into the sketch kernel's emit -- the kept k-mers go straight into per-(genome, bucket) regions,
    // and the sort starts at its bucket kernel.  Nothing is read back in between.  A region that overflows (a genome
    // whose k-mers crowd a few buckets) sends the call to the route below.
    uint64_t max_expect = 0, total_bound = 0;
    for (int g = 0; g < G; ++g) {
      const uint64_t wins = genome_windows(batch, g, window);
      const uint64_t e = std::min<uint64_t>(wins, (uint64_t)((double)wins / (double)pred->modulus * 1.25) + 1024);
      max_expect = std::max(max_expect, e);
      total_bound += e;   // room of the output; more distinct keys than that raise the flag (sortp_region_offsets_kernel)
    }
    int bb = 0;
    uint32_t cap = 0;
    SortPlan kplan;
    if (bucket_regions_plan(mask, G, max_expect, &bb, &cap, &kplan)) {
      BufferRef regions, ctl;
      const size_t n_regions = (size_t)G << bb;
      SKS_TRY(alloc_buffer(ctx, n_regions * cap * 8, &regions));
      SKS_TRY(alloc_buffer(ctx, 4 * n_regions + 256, &ctl));
      SKS_CUDA_TRY(cudaMemsetAsync(ctl->ptr, 0, 4 * n_regions + 256, ctx->stream));
      uint32_t *d_cursor = static_cast<uint32_t *>(ctl->ptr), *d_flag = d_cursor + n_regions;
      plan.p.out_keys = regions->ptr;
      plan.p.out_pos = nullptr;
      plan.p.out_off = nullptr;
      plan.p.out_cap = nullptr;
      plan.p.out_count = nullptr;
      plan.p.kpart_bits = (uint32_t)bb;
      plan.p.kpart_cap = cap;
      plan.p.kpart_cursor = d_cursor;
      plan.p.kpart_overflow = d_flag;
      plan.p.kpart_plan = kplan;
      const uint32_t *tile_genome = batch->tile_genome ? static_cast<const uint32_t *>(batch->tile_genome->ptr) : nullptr;
      if (batch->arriving.empty()) {
        SKS_TRY(launch_sketch(ctx, plan.p, tile_genome, plan.n_limbs, plan.pred_mode, OUT_KEYS));
      } else {
        // the genomes are still on their way from the host: every chunk is sketched as soon as it is there, into the same
        // regions, while the copy engine brings the next one
        for (size_t c = 0; c < batch->arriving.size(); ++c) {
          const sks_batch::Arriving &chunk = batch->arriving[c];
          SKS_TRY(batch_chunk_ready(ctx, batch, c));
          plan.p.tile_begin = chunk.tile_begin;
          plan.p.n_tiles = chunk.tile_end;
          SKS_TRY(launch_sketch(ctx, plan.p, tile_genome, plan.n_limbs, plan.pred_mode, OUT_KEYS));
        }
        plan.p.tile_begin = 0;
        plan.p.n_tiles = batch->n_tiles;
      }
      plan.p.kpart_bits = 0;
      SKS_TRY(sort_unique_from_buckets(ctx, static_cast<unsigned long long *>(regions->ptr), G, bb, cap, d_cursor, d_flag, total_bound,
                                       &uniq, &uoff, &ucount, &done));
      if (!done) ctx->kpart_skip = 63;
    }
  }
  if (!done) {
    SKS_TRY(batch_arrived(ctx, batch));
    SKS_TRY(sketch_raw_keys(ctx, batch, plan, pred, window, OUT_KEYS, &raw, &pos, &off, &count, &span));
    SKS_TRY(sort_unique_regions(ctx, key_words, raw->ptr, off.data(), count.data(), G, span, &uniq, &uoff, &ucount,
                                mask));
  }
  for (int g = 0; g < G; ++g) {
    sks_set *s = new_set(ctx, SKS_REPR_SORTED, mask, window, plan.weight);
    if (!s) return set_error(SKS_ERR_INVALID, "out of host memory");
    s->buf = uniq;
    s->key_words = key_words;
    s->byte_off = (size_t)uoff[g] * 8 * key_words;
    s->count = (int64_t)ucount[g];
    out_sets[g] = s;
  }
  return SKS_OK;
}

}  // namespace
int check_pair(const sks_set *a, const sks_set *b) {
  if (!a || !b) return set_error(SKS_ERR_INVALID, "null set");
  // window_length is not part of k-mer equality (src/kmer.hpp:82-85); mask and layout are
  if (a->repr != b->repr || a->mask[0] != b->mask[0] || a->mask[1] != b->mask[1] || a->key_words != b->key_words ||
      a->device != b->device)
    return set_error(SKS_ERR_MISMATCH, "sets were built with different masks, representations or devices");
  return SKS_OK;
}

}  // namespace sks

using namespace sks;

extern "C" {

int sks_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int sks_ctx_create(int device, sks_ctx **out) {
  if (!out) return set_error(SKS_ERR_INVALID, "null argument");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_error(SKS_ERR_CUDA, "no CUDA device is available; libsks has no CPU fallback");
  }
  if (device < 0 || device >= n) return set_error(SKS_ERR_INVALID, "device %d outside 0..%d", device, n - 1);
  SKS_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  SKS_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return set_error(SKS_ERR_CUDA, "device %d is sm_%d%d; libsks is built for sm_100a (B200) only", device, prop.major,
                     prop.minor);
  sks_ctx *ctx = new (std::nothrow) sks_ctx();
  if (!ctx) return set_error(SKS_ERR_INVALID, "out of host memory");
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  SKS_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  stream_attach(device, ctx->stream);
  SKS_CUDA_TRY(cudaEventCreate(&ctx->ev0));
  SKS_CUDA_TRY(cudaEventCreate(&ctx->ev1));
  // keep freed blocks cached in the stream-ordered pool: steady-state calls allocate nothing
  cudaMemPool_t pool;
  SKS_CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t threshold = UINT64_MAX;
  SKS_CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold));
  if (const char *e = getenv("SKS_BUCKET_MIN_BITS")) ctx->bucket_min_bits = atoi(e);
  if (const char *e = getenv("SKS_EXACT_PARTITION")) ctx->exact_partition = atoi(e) != 0;
  *out = ctx;
  return SKS_OK;
}

void sks_ctx_destroy(sks_ctx *ctx) {
  if (!ctx) return;
  DeviceGuard guard(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->scratch) cudaFreeAsync(ctx->scratch, ctx->stream);
  stream_detach(ctx->device, ctx->stream);
  if (!stream_is_live(ctx->device, ctx->stream)) cache_release(ctx->device, ctx->stream, true);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (cudaEvent_t e : ctx->sync_events) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stage_ring) cudaFreeHost(ctx->stage_ring);
  for (auto &v : ctx->prof)
    for (auto &pr : v) {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  if (ctx->owns_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int sks_ctx_set_stream(sks_ctx *ctx, void *cuda_stream) {
  if (!ctx) return set_error(SKS_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  stream_detach(ctx->device, ctx->stream);
  if (!stream_is_live(ctx->device, ctx->stream)) cache_release(ctx->device, ctx->stream, true);
  if (ctx->scratch) {
    SKS_CUDA_TRY(cudaFreeAsync(ctx->scratch, ctx->stream));
    SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
  }
  if (ctx->owns_stream && ctx->stream) SKS_CUDA_TRY(cudaStreamDestroy(ctx->stream));
  ctx->stream = static_cast<cudaStream_t>(cuda_stream);
  ctx->owns_stream = false;
  stream_attach(ctx->device, ctx->stream);
  return SKS_OK;
}

int sks_ctx_sync(sks_ctx *ctx) {
  if (!ctx) return set_error(SKS_ERR_INVALID, "null context");
  DeviceGuard guard(ctx->device);
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return SKS_OK;
}

int sks_timer_begin(sks_ctx *ctx) {
  DeviceGuard guard(ctx->device);
  SKS_CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
  return SKS_OK;
}
int sks_timer_end(sks_ctx *ctx, float *out_ms) {
  DeviceGuard guard(ctx->device);
  SKS_CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
  SKS_CUDA_TRY(cudaEventSynchronize(ctx->ev1));
  SKS_CUDA_TRY(cudaEventElapsedTime(out_ms, ctx->ev0, ctx->ev1));
  return SKS_OK;
}
int64_t sks_ctx_launch_count(const sks_ctx *ctx) { return ctx ? ctx->launches : 0; }
int64_t sks_ctx_in_place_count(const sks_ctx *ctx) { return ctx ? ctx->in_place_calls : 0; }
int64_t sks_ctx_streamed_count(const sks_ctx *ctx) { return ctx ? ctx->streamed_calls : 0; }

int sks_ctx_profile(sks_ctx *ctx, int enable) {
  if (!ctx) return set_error(SKS_ERR_INVALID, "null context");
  ctx->profile = enable != 0;
  return SKS_OK;
}
int sks_ctx_kernel_stats(sks_ctx *ctx, int kind, int64_t *out_launches, double *out_total_ms) {
  if (!ctx || kind < 0 || kind >= SKS_KERNEL_KINDS) return set_error(SKS_ERR_INVALID, "bad argument");
  DeviceGuard guard(ctx->device);
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  double total = 0;
  for (auto &pr : ctx->prof[kind]) {
    float ms = 0;
    SKS_CUDA_TRY(cudaEventElapsedTime(&ms, pr.first, pr.second));
    total += ms;
    ctx->event_pool.push_back(pr.first);
    ctx->event_pool.push_back(pr.second);
  }
  if (out_launches) *out_launches = (int64_t)ctx->prof[kind].size();
  if (out_total_ms) *out_total_ms = total;
  ctx->prof[kind].clear();
  return SKS_OK;
}
const char *sks_kernel_name(int kind) {
  static const char *names[SKS_KERNEL_KINDS] = {"sketch_kernel", "fill_zero_kernel", "bitset_pair_counts_kernel",
                                                "bitset_popcount_kernel", "sort_unique", "sorted_intersect_kernel",
                                                "synth_kernel", "list_finalize", "bitset_build", "fasta_parse",
                                                "bitset_pair_build_kernel", "dict_build", "allpairs_kernel",
                                                "ani_finalize_kernel", "nccl_exchange"};
  return (kind >= 0 && kind < SKS_KERNEL_KINDS) ? names[kind] : "?";
}

// ---- batches -------------------------------------------------------------------------------------
int sks_batch_upload(sks_ctx *ctx, int n_genomes, const uint32_t *const *packed, const uint64_t *n_bases,
                     const uint64_t *const *seg_len, const uint64_t *n_segs, sks_batch **out) {
  if (!ctx || !out || n_genomes < 0 || (n_genomes > 0 && (!packed || !n_bases)))
    return set_error(SKS_ERR_INVALID, "bad argument");
  DeviceGuard guard(ctx->device);
  sks_batch *b = new (std::nothrow) sks_batch();
  if (!b) return set_error(SKS_ERR_INVALID, "out of host memory");
  b->device = ctx->device;
  uint64_t total_words = 0;
  int st = layout_batch(b, n_genomes, n_bases, seg_len, n_segs, &total_words);
  if (st == SKS_OK) st = alloc_buffer(ctx, (size_t)total_words * 4, &b->words);
  if (st == SKS_OK) {
    uint32_t *d = static_cast<uint32_t *>(b->words->ptr);
    // zero only the pads (history words, 16-byte rounding, halo); the data words are overwritten
    uint64_t prev_end = 0;
    auto fail = [&](cudaError_t e) { return e == cudaSuccess ? SKS_OK : set_error(SKS_ERR_CUDA, "batch upload: %s", cudaGetErrorString(e)); };
    for (int g = 0; g < n_genomes && st == SKS_OK; ++g) {
      const GenomeDesc &gd = b->h_genomes[g];
      const uint64_t data_words = (n_bases[g] + 15) / 16;
      st = fail(cudaMemsetAsync(d + prev_end, 0, (gd.word_off - prev_end) * 4, ctx->stream));
      if (st == SKS_OK && data_words)
        st = fail(cudaMemcpyAsync(d + gd.word_off, packed[g], data_words * 4, cudaMemcpyHostToDevice, ctx->stream));
      prev_end = gd.word_off + data_words;
    }
    if (st == SKS_OK) st = fail(cudaMemsetAsync(d + prev_end, 0, (total_words - prev_end) * 4, ctx->stream));
  }
  if (st == SKS_OK) st = upload_tables(ctx, b);
  if (st != SKS_OK) {
    delete b;
    return st;
  }
  *out = b;
  return SKS_OK;
}

// A batch whose genomes stay where the caller has them: pinned (page-locked, device-mapped) host buffers, 16-byte
// aligned.  The sketch kernel's bulk copies then read the 2-bit words over PCIe tile by tile while it computes, and the
// separate host-to-device copy (63 us for a 5 Mbp pair) disappears from the call.  Only for calls that hold the batch
// themselves and return after the stream has drained (sks_pair_ani); fails without side effects when a buffer does
// not qualify.  SKS_ZERO_COPY=0 turns it off.
static int batch_in_place(sks_ctx *ctx, int n_genomes, const uint32_t *const *packed, const uint64_t *n_bases,
                          sks_batch **out) {
  static const bool enabled = [] { const char *e = getenv("SKS_ZERO_COPY"); return !e || atoi(e) != 0; }();
  if (!enabled || n_genomes <= 0) return SKS_ERR_INVALID;
  DeviceGuard guard(ctx->device);
  std::vector<uint64_t> dev_word(n_genomes);
  for (int g = 0; g < n_genomes; ++g) {
    if (!packed[g] || n_bases[g] == 0) return SKS_ERR_INVALID;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, packed[g]) != cudaSuccess) {
      cudaGetLastError();
      return SKS_ERR_INVALID;
    }
    if (attr.type != cudaMemoryTypeHost || !attr.devicePointer) return SKS_ERR_INVALID;
    const uintptr_t a = reinterpret_cast<uintptr_t>(attr.devicePointer);
    if (a & 15) return SKS_ERR_INVALID;
    dev_word[g] = (uint64_t)(a >> 2);
  }
  sks_batch *b = new (std::nothrow) sks_batch();
  if (!b) return SKS_ERR_INVALID;
  b->device = ctx->device;
  uint64_t total_words = 0;
  int st = layout_batch(b, n_genomes, n_bases, nullptr, nullptr, &total_words);
  if (st == SKS_OK) {
    b->host_words = true;
    for (int g = 0; g < n_genomes; ++g) {
      b->h_genomes[g].word_off = dev_word[g];
      b->h_genomes[g].n_words = (uint32_t)((n_bases[g] + 15) / 16);  // exact: nothing beyond may be read
    }
    st = upload_tables(ctx, b);
  }
  if (st != SKS_OK) {
    delete b;
    return st;
  }
  *out = b;
  return SKS_OK;
}

// Zeroes what lies between the genomes of a batch buffer (history words, 16-byte rounding, the halo behind the last
// genome): a warp per gap.
__global__ void zero_pads_kernel(uint32_t *words, const GenomeDesc *genomes, int n_genomes, unsigned long long total_words) {
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g > n_genomes) return;
  const unsigned long long lo = g == 0 ? 0ull : genomes[g - 1].word_off + (genomes[g - 1].n_bases + 15ull) / 16;
  const unsigned long long hi = g == n_genomes ? total_words : genomes[g].word_off;
  for (unsigned long long i = lo + (threadIdx.x & 31); i < hi; i += 32) words[i] = 0;
}

// A device batch for genomes in host buffers that is filled WHILE it is sketched: the copy engine brings the genomes
// chunk by chunk (SKS_HOST_CHUNK_MB, default 16 MB: large copies reach 55 GB/s on PCIe 5 where the sketch kernel's own
// 2 KB reads of host memory stay at 44 GB/s; 8 to 64 MB measure the same, 27.3 ms against 31.1 ms for 1000 x 5 Mbp), every
// chunk records an event, and sketch_sorted launches the kernel chunk by chunk behind them.
//   pinned buffers   : all copies are queued here, straight from the caller's memory (a run of equally long, equally
//                      spaced genomes is one strided copy);
//   pageable buffers : a feeder thread and its helpers copy chunk after chunk into a ring of three pinned staging
//                      buffers -- laid out like the device buffer, so that a chunk is one copy -- and queue the copies as
//                      they go; the consumers wait for "chunk c is queued" (sks_batch::Feed) before they wait for its
//                      event.  cudaMemcpyAsync from pageable memory goes through the driver's own staging at ~10 GB/s:
//                      136 ms for 1000 x 5 Mbp, against 28.5 ms this way with 8 threads (59 ms with 2, 35 ms with 4).
// Only for sks_all_vs_all_from_host, which keeps the batch to itself, and only from SKS_HOST_STREAM_MIN_MB (default 64)
// on: below that the in-place route / a plain upload have less to set up.  SKS_HOST_STREAM=0: off.
namespace {
struct CopyPiece {
  char *dst;
  const char *src;
  size_t bytes;
};
// The feeder's helpers: run() hands a list of host copies to all threads (the caller takes part) and returns when done.
struct CopyTeam {
  std::vector<std::thread> helpers;
  std::mutex m;
  std::condition_variable cv_start, cv_done;
  const std::vector<CopyPiece> *pieces = nullptr;
  std::atomic<size_t> next{0};
  int generation = 0, running = 0;
  bool stop = false;
  explicit CopyTeam(int n_helpers) {
    helpers.reserve(n_helpers > 0 ? n_helpers : 0);
    for (int i = 0; i < n_helpers; ++i) {
      try {
        helpers.emplace_back([this] { loop(); });
      } catch (...) {  // no more threads to be had: fewer helpers
        break;
      }
    }
  }
  ~CopyTeam() {
    {
      std::lock_guard<std::mutex> lk(m);
      stop = true;
    }
    cv_start.notify_all();
    for (std::thread &t : helpers) t.join();
  }
  void work() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= pieces->size()) return;
      memcpy((*pieces)[i].dst, (*pieces)[i].src, (*pieces)[i].bytes);
    }
  }
  void loop() {
    int seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(m);
      cv_start.wait(lk, [&] { return stop || generation != seen; });
      if (stop) return;
      seen = generation;
      lk.unlock();
      work();
      lk.lock();
      if (--running == 0) cv_done.notify_one();
    }
  }
  void run(const std::vector<CopyPiece> &p) {
    {
      std::lock_guard<std::mutex> lk(m);
      pieces = &p;
      next = 0;
      running = (int)helpers.size();
      ++generation;
    }
    cv_start.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m);
    cv_done.wait(lk, [&] { return running == 0; });
  }
};
}  // namespace

static int batch_streamed(sks_ctx *ctx, int n_genomes, const uint32_t *const *packed, const uint64_t *n_bases, int world,
                          sks_batch **out) {
  // (read at every call: the tests change them)
  const char *e_on = getenv("SKS_HOST_STREAM"), *e_min = getenv("SKS_HOST_STREAM_MIN_MB"), *e_chunk = getenv("SKS_HOST_CHUNK_MB");
  const char *e_threads = getenv("SKS_HOST_THREADS");
  const bool enabled = !e_on || atoi(e_on) != 0;
  const uint64_t min_bytes = (uint64_t)std::max<long long>(e_min ? atoll(e_min) : 64, 0) << 20;
  const uint64_t chunk_bytes = (uint64_t)std::max<long long>(e_chunk ? atoll(e_chunk) : 16, 1) << 20;
  if (!enabled || n_genomes <= 0) return SKS_ERR_INVALID;
  DeviceGuard guard(ctx->device);
  uint64_t total_bytes = 0;
  for (int g = 0; g < n_genomes; ++g) {
    if (!packed[g] || n_bases[g] == 0) return SKS_ERR_INVALID;
    total_bytes += (n_bases[g] + 15) / 16 * 4;
  }
  if (total_bytes < min_bytes) return SKS_ERR_INVALID;
  int n_pinned = 0, n_pageable = 0;
  for (int g = 0; g < n_genomes; ++g) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, packed[g]) != cudaSuccess) {
      cudaGetLastError();
      return SKS_ERR_INVALID;
    }
    if (attr.type == cudaMemoryTypeHost) ++n_pinned;
    else if (attr.type == cudaMemoryTypeUnregistered) ++n_pageable;
  }
  const bool staged = n_pageable == n_genomes;
  if (!staged && n_pinned != n_genomes) return SKS_ERR_INVALID;  // device, managed or mixed memory: the plain upload
  sks_batch *b = new (std::nothrow) sks_batch();
  if (!b) return SKS_ERR_INVALID;
  b->device = ctx->device;
  uint64_t total_words = 0;
  int st = layout_batch(b, n_genomes, n_bases, nullptr, nullptr, &total_words);
  // chunks: whole genomes, at least chunk_bytes each (but for the last); the device words of chunk [g0, g1) are
  // [word_off(g0) - kPreWords, word_off(g1 - 1) + n_words(g1 - 1)), the last chunk runs to the end of the buffer
  struct Chunk {
    int g0, g1;
    uint64_t w0, w1;
  };
  std::vector<Chunk> chunks;
  uint64_t slot_words = 0;
  if (st == SKS_OK) {
    for (int g = 0; g < n_genomes;) {
      Chunk c{g, g, 0, 0};
      uint64_t bytes = 0;
      while (c.g1 < n_genomes && bytes < chunk_bytes) bytes += (n_bases[c.g1++] + 15) / 16 * 4;
      c.w0 = b->h_genomes[c.g0].word_off - kPreWords;
      c.w1 = c.g1 == n_genomes ? total_words : b->h_genomes[c.g1 - 1].word_off + b->h_genomes[c.g1 - 1].n_words;
      slot_words = std::max(slot_words, c.w1 - c.w0);
      chunks.push_back(c);
      g = c.g1;
    }
    if (staged && slot_words * 4 > ((uint64_t)256 << 20)) st = SKS_ERR_INVALID;  // one huge genome: not worth a 768 MB ring
  }
  if (st != SKS_OK) {
    delete b;
    return SKS_ERR_INVALID;
  }
  st = alloc_buffer(ctx, (size_t)total_words * 4, &b->words);
  if (st == SKS_OK) st = upload_tables(ctx, b);
  auto fail = [&](cudaError_t e) { return e == cudaSuccess ? SKS_OK : set_error(SKS_ERR_CUDA, "streamed batch: %s", cudaGetErrorString(e)); };
  if (st == SKS_OK && !ctx->copy_stream) st = fail(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  while (st == SKS_OK && ctx->sync_events.size() < chunks.size() + 1) {
    cudaEvent_t e;
    st = fail(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (st == SKS_OK) ctx->sync_events.push_back(e);
  }
  if (st == SKS_OK && staged && ctx->stage_ring_bytes < 3 * slot_words * 4) {
    if (ctx->stage_ring) cudaFreeHost(ctx->stage_ring);
    ctx->stage_ring = nullptr;
    ctx->stage_ring_bytes = 0;
    st = fail(cudaMallocHost(&ctx->stage_ring, 3 * slot_words * 4));
    if (st == SKS_OK) ctx->stage_ring_bytes = 3 * slot_words * 4;
  }
  uint32_t *d = st == SKS_OK ? static_cast<uint32_t *>(b->words->ptr) : nullptr;
  if (st == SKS_OK) {
    // what lies between the genomes is zeroed here (the pinned route copies the data words only); the buffer may have
    // been in use by earlier work of the context's stream: the copies start behind all that
    zero_pads_kernel<<<(unsigned)((n_genomes + 1 + 7) / 8), 256, 0, ctx->stream>>>(d, static_cast<const GenomeDesc *>(b->genomes->ptr),
                                                                               n_genomes, (unsigned long long)total_words);
    st = fail(cudaGetLastError());
    ctx->launches++;
    if (st == SKS_OK) st = fail(cudaEventRecord(ctx->sync_events[chunks.size()], ctx->stream));
    if (st == SKS_OK) st = fail(cudaStreamWaitEvent(ctx->copy_stream, ctx->sync_events[chunks.size()], 0));
  }
  if (st == SKS_OK)
    for (size_t c = 0; c < chunks.size(); ++c) {
      const GenomeDesc &first = b->h_genomes[chunks[c].g0], &last = b->h_genomes[chunks[c].g1 - 1];
      b->arriving.push_back({first.tile_first, last.tile_first + last.n_tiles, ctx->sync_events[c]});
    }
  if (st == SKS_OK && !staged) {
    for (size_t c = 0; c < chunks.size() && st == SKS_OK; ++c) {
      for (int g = chunks[c].g0; g < chunks[c].g1 && st == SKS_OK;) {
        // a run of equally long, equally spaced genomes is one strided copy
        const uint64_t row = (n_bases[g] + 15) / 16 * 4;
        int run = 1;
        if (g + 1 < chunks[c].g1 && n_bases[g + 1] == n_bases[g] && packed[g + 1] > packed[g]) {
          const uint64_t pitch = (uint64_t)(reinterpret_cast<const char *>(packed[g + 1]) - reinterpret_cast<const char *>(packed[g]));
          if (pitch >= row && pitch < ((uint64_t)1 << 31)) {
            while (g + run < chunks[c].g1 && n_bases[g + run] == n_bases[g] &&
                   reinterpret_cast<const char *>(packed[g + run]) == reinterpret_cast<const char *>(packed[g]) + (uint64_t)run * pitch)
              ++run;
            if (run > 1)
              st = fail(cudaMemcpy2DAsync(d + b->h_genomes[g].word_off, (size_t)(b->h_genomes[g + 1].word_off - b->h_genomes[g].word_off) * 4,
                                          packed[g], (size_t)pitch, (size_t)row, (size_t)run, cudaMemcpyHostToDevice, ctx->copy_stream));
          }
        }
        if (run == 1) st = fail(cudaMemcpyAsync(d + b->h_genomes[g].word_off, packed[g], (size_t)row, cudaMemcpyHostToDevice, ctx->copy_stream));
        g += run;
      }
      if (st == SKS_OK) st = fail(cudaEventRecord(b->arriving[c].ready, ctx->copy_stream));
    }
  }
  if (st == SKS_OK && staged) {
    int n_threads = e_threads ? atoi(e_threads) : (int)std::thread::hardware_concurrency() / std::max(world, 1);  // the ranks of a node share its cores
    n_threads = std::min(std::max(n_threads, 1), 8);
    b->feed.reset(new (std::nothrow) sks_batch::Feed());
    if (!b->feed) st = set_error(SKS_ERR_INVALID, "out of host memory");
  if (st == SKS_OK) {
    sks_batch::Feed *feed = b->feed.get();
    const int device = ctx->device;
    cudaStream_t copy_stream = ctx->copy_stream;
    char *ring = static_cast<char *>(ctx->stage_ring);
    const size_t slot_bytes = (size_t)slot_words * 4;
    std::vector<const uint32_t *> src(packed, packed + n_genomes);
    const sks_batch *batch = b;
    auto feeder = [=]() {
      auto done = [&](size_t n, int status) {
        {
          std::lock_guard<std::mutex> lk(feed->m);
          feed->n_queued = n;
          if (status != SKS_OK) feed->status = status;
        }
        feed->cv.notify_all();
      };
      if (cudaSetDevice(device) != cudaSuccess) return done(chunks.size(), SKS_ERR_CUDA);
      CopyTeam team(n_threads - 1);
      std::vector<CopyPiece> pieces;
      const size_t kPiece = (size_t)256 << 10;
      for (size_t c = 0; c < chunks.size(); ++c) {
        char *slot = ring + (c % 3) * slot_bytes;
        // the copy that read this slot three chunks ago must be over
        if (c >= 3 && cudaEventSynchronize(batch->arriving[c - 3].ready) != cudaSuccess) return done(chunks.size(), SKS_ERR_CUDA);
        pieces.clear();
        uint64_t at = chunks[c].w0;  // device word the next zero gap starts at
        for (int g = chunks[c].g0; g < chunks[c].g1; ++g) {
          const GenomeDesc &gd = batch->h_genomes[g];
          const size_t bytes = (size_t)((gd.n_bases + 15ull) / 16) * 4;
          memset(slot + (at - chunks[c].w0) * 4, 0, (size_t)(gd.word_off - at) * 4);
          char *dst = slot + (gd.word_off - chunks[c].w0) * 4;
          const char *from = reinterpret_cast<const char *>(src[g]);
          for (size_t o = 0; o < bytes; o += kPiece) pieces.push_back({dst + o, from + o, std::min(kPiece, bytes - o)});
          at = gd.word_off + bytes / 4;
        }
        memset(slot + (at - chunks[c].w0) * 4, 0, (size_t)(chunks[c].w1 - at) * 4);
        team.run(pieces);
        if (cudaMemcpyAsync(d + chunks[c].w0, slot, (size_t)(chunks[c].w1 - chunks[c].w0) * 4, cudaMemcpyHostToDevice, copy_stream) != cudaSuccess ||
            cudaEventRecord(batch->arriving[c].ready, copy_stream) != cudaSuccess)
          return done(chunks.size(), SKS_ERR_CUDA);
        done(c + 1, SKS_OK);
      }
    };
    try {
      feed->worker = std::thread(feeder);
    } catch (...) {
      st = set_error(SKS_ERR_INVALID, "could not start the thread that stages the host genomes");
    }
  }
  }
  if (st != SKS_OK) {
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);  // nothing of a failed call may touch the caller's buffers later
    delete b;
    return st;
  }
  *out = b;
  return SKS_OK;
}

int sks_batch_synth(sks_ctx *ctx, int n_genomes, uint64_t n_bases, const uint64_t *gen_seed, const uint64_t *mut_seed,
                    const uint64_t *mut_D, sks_batch **out) {
  std::vector<uint64_t> zeros(n_genomes > 0 ? n_genomes : 1, 0);
  return sks_batch_synth_at(ctx, n_genomes, n_bases, zeros.data(), gen_seed, mut_seed, mut_D, out);
}

int sks_batch_synth_at(sks_ctx *ctx, int n_genomes, uint64_t n_bases, const uint64_t *first_base,
                       const uint64_t *gen_seed, const uint64_t *mut_seed, const uint64_t *mut_D, sks_batch **out) {
  if (!first_base && n_genomes > 0) return set_error(SKS_ERR_INVALID, "bad argument");
  if (!ctx || !out || n_genomes < 0 || (n_genomes > 0 && (!gen_seed || !mut_seed || !mut_D)))
    return set_error(SKS_ERR_INVALID, "bad argument");
  DeviceGuard guard(ctx->device);
  sks_batch *b = new (std::nothrow) sks_batch();
  if (!b) return set_error(SKS_ERR_INVALID, "out of host memory");
  b->device = ctx->device;
  std::vector<uint64_t> nb(n_genomes, n_bases);
  uint64_t total_words = 0;
  int st = layout_batch(b, n_genomes, nb.data(), nullptr, nullptr, &total_words);
  if (st == SKS_OK) st = alloc_buffer(ctx, (size_t)total_words * 4, &b->words);
  if (st == SKS_OK) st = launch_fill_zero(ctx, b->words->ptr, (size_t)total_words * 4);
  if (st == SKS_OK) st = upload_tables(ctx, b);
  if (st == SKS_OK && n_genomes > 0) {
    uint64_t *d_seeds = nullptr;
    st = ctx_scratch(ctx, (size_t)n_genomes * 32, reinterpret_cast<void **>(&d_seeds));
    auto cp = [&](uint64_t *dst, const uint64_t *src) {
      return cudaMemcpyAsync(dst, src, (size_t)n_genomes * 8, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
    };
    if (st == SKS_OK && !(cp(d_seeds, gen_seed) && cp(d_seeds + n_genomes, mut_seed) && cp(d_seeds + 2 * n_genomes, mut_D) &&
                            cp(d_seeds + 3 * n_genomes, first_base)))
      st = set_error(SKS_ERR_CUDA, "seed upload failed");
    if (st == SKS_OK)
      st = launch_synth(ctx, static_cast<uint32_t *>(b->words->ptr), static_cast<const GenomeDesc *>(b->genomes->ptr),
                        n_genomes, b->h_genomes[0].n_words, d_seeds, d_seeds + n_genomes, d_seeds + 2 * n_genomes,
                        d_seeds + 3 * n_genomes);
    if (st == SKS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = set_error(SKS_ERR_CUDA, "synth failed");
  }
  if (st != SKS_OK) {
    delete b;
    return st;
  }
  *out = b;
  return SKS_OK;
}

int sks_batch_from_fasta_text(sks_ctx *ctx, int n_files, const char *const *text, const uint64_t *n_bytes, sks_batch **out) {
  if (!ctx || !out || n_files < 0 || (n_files > 0 && (!text || !n_bytes))) return set_error(SKS_ERR_INVALID, "bad argument");
  DeviceGuard guard(ctx->device);
  std::vector<uint64_t> file_off((size_t)n_files + 1, 0);
  for (int f = 0; f < n_files; ++f) file_off[(size_t)f + 1] = file_off[(size_t)f] + n_bytes[f];
  const uint64_t total = file_off.back();
  BufferRef d_text;
  SKS_TRY(alloc_buffer(ctx, (size_t)total + 16, &d_text));
  for (int f = 0; f < n_files; ++f)
    if (n_bytes[f])
      SKS_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(d_text->ptr) + file_off[(size_t)f], text[f], (size_t)n_bytes[f],
                                   cudaMemcpyHostToDevice, ctx->stream));
  std::vector<uint64_t> nb;
  std::vector<std::vector<uint64_t>> segs;
  BufferRef codes, gbase;
  SKS_TRY(fasta_parse_device(ctx, static_cast<const unsigned char *>(d_text->ptr), file_off, &nb, &segs, &codes, &gbase));
  sks_batch *b = new (std::nothrow) sks_batch();
  if (!b) return set_error(SKS_ERR_INVALID, "out of host memory");
  b->device = ctx->device;
  std::vector<const uint64_t *> seg_ptr((size_t)n_files);
  std::vector<uint64_t> n_segs((size_t)n_files);
  for (int f = 0; f < n_files; ++f) {
    seg_ptr[(size_t)f] = segs[(size_t)f].data();
    n_segs[(size_t)f] = segs[(size_t)f].size();
  }
  uint64_t total_words = 0;
  int st = layout_batch(b, n_files, nb.data(), seg_ptr.data(), n_segs.data(), &total_words);
  if (st == SKS_OK) st = alloc_buffer(ctx, (size_t)total_words * 4, &b->words);
  if (st == SKS_OK) st = launch_fill_zero(ctx, b->words->ptr, (size_t)total_words * 4);
  if (st == SKS_OK) st = upload_tables(ctx, b);
  if (st == SKS_OK && n_files > 0) {
    uint32_t max_words = 0;
    for (const GenomeDesc &gd : b->h_genomes) max_words = std::max(max_words, gd.n_words);
    st = launch_pack_codes(ctx, static_cast<const uint8_t *>(codes->ptr), static_cast<const uint32_t *>(gbase->ptr),
                           static_cast<const GenomeDesc *>(b->genomes->ptr), n_files, max_words,
                           static_cast<uint32_t *>(b->words->ptr));
  }
  if (st != SKS_OK) {
    delete b;
    return st;
  }
  *out = b;
  return SKS_OK;
}

int sks_batch_from_fasta_files(sks_ctx *ctx, int n_files, const char *const *paths, sks_batch **out) {
  if (!ctx || !out || n_files < 0 || (n_files > 0 && !paths)) return set_error(SKS_ERR_INVALID, "bad argument");
  std::vector<std::string> texts((size_t)n_files);
  for (int f = 0; f < n_files; ++f) {
    FILE *fp = fopen(paths[f], "rb");
    if (!fp) return set_error(SKS_ERR_IO, "Unable to open %s", paths[f]);
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), fp)) > 0) texts[(size_t)f].append(buf, got);
    fclose(fp);
  }
  std::vector<const char *> ptr((size_t)n_files);
  std::vector<uint64_t> len((size_t)n_files);
  for (int f = 0; f < n_files; ++f) {
    ptr[(size_t)f] = texts[(size_t)f].data();
    len[(size_t)f] = texts[(size_t)f].size();
  }
  const int st = sks_batch_from_fasta_text(ctx, n_files, ptr.data(), len.data(), out);
  if (st == SKS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) return set_error(SKS_ERR_CUDA, "FASTA upload failed");
  return st;
}

int sks_batch_segments(const sks_batch *b, int genome, uint64_t *n_segs, uint64_t *out_seg_len) {
  if (!b || !n_segs || genome < 0 || genome >= b->n_genomes) return set_error(SKS_ERR_INVALID, "bad argument");
  const GenomeDesc &gd = b->h_genomes[genome];
  uint64_t prev = 0, n = 0;
  for (uint32_t s = 0; s < gd.n_segs; ++s) {
    const uint64_t e = b->h_seg_end[gd.seg_first + s];
    if (e > prev) {
      if (out_seg_len) out_seg_len[n] = e - prev;
      ++n;
    }
    prev = e;
  }
  *n_segs = n;
  return SKS_OK;
}

int sks_batch_n_genomes(const sks_batch *b) { return b ? b->n_genomes : 0; }
uint64_t sks_batch_n_bases(const sks_batch *b, int genome) {
  return (b && genome >= 0 && genome < b->n_genomes) ? b->h_genomes[genome].n_bases : 0;
}

int sks_batch_download(sks_ctx *ctx, const sks_batch *b, int genome, uint32_t *out_words) {
  if (!ctx || !b || genome < 0 || genome >= b->n_genomes || !out_words) return set_error(SKS_ERR_INVALID, "bad argument");
  DeviceGuard guard(ctx->device);
  const GenomeDesc &gd = b->h_genomes[genome];
  const uint64_t data_words = ((uint64_t)gd.n_bases + 15) / 16;
  SKS_CUDA_TRY(cudaMemcpyAsync(out_words, static_cast<const uint32_t *>(b->words->ptr) + gd.word_off, data_words * 4,
                               cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return SKS_OK;
}

int sks_batch_slice(sks_ctx *ctx, const sks_batch *src, int genome, uint64_t first_base, uint64_t n_starts, int window,
                    sks_batch **out) {
  if (!ctx || !src || !out || genome < 0 || genome >= src->n_genomes || window < 1 || window > 64)
    return set_error(SKS_ERR_INVALID, "bad argument");
  if (first_base % 16 != 0) return set_error(SKS_ERR_INVALID, "slice start must be a multiple of 16 bases");
  DeviceGuard guard(ctx->device);
  const GenomeDesc &sg = src->h_genomes[genome];
  if (first_base > sg.n_bases) first_base = sg.n_bases - sg.n_bases % 16;
  uint64_t end = first_base + n_starts + window - 1;  // exclusive: the (w-1)-base halo after the last start
  if (end > sg.n_bases || n_starts == 0) end = n_starts == 0 ? first_base : sg.n_bases;
  const uint64_t nb = end - first_base;
  // segments clipped to [first_base, end)
  std::vector<uint64_t> segs;
  uint64_t prev = 0;
  for (uint32_t s = 0; s < sg.n_segs; ++s) {
    const uint64_t e = src->h_seg_end[sg.seg_first + s];
    const uint64_t lo = std::max<uint64_t>(prev, first_base), hi = std::min<uint64_t>(e, end);
    if (hi > lo) segs.push_back(hi - lo);
    prev = e;
  }
  sks_batch *b = new (std::nothrow) sks_batch();
  if (!b) return set_error(SKS_ERR_INVALID, "out of host memory");
  b->device = ctx->device;
  const uint64_t *seg_ptr = segs.data();
  const uint64_t ns = segs.size();
  uint64_t total_words = 0;
  int st = layout_batch(b, 1, &nb, ns ? &seg_ptr : nullptr, ns ? &ns : nullptr, &total_words);
  if (st == SKS_OK) st = alloc_buffer(ctx, (size_t)total_words * 4, &b->words);
  if (st == SKS_OK) st = launch_fill_zero(ctx, b->words->ptr, (size_t)total_words * 4);
  if (st == SKS_OK && nb > 0) {
    const uint64_t data_words = (nb + 15) / 16;
    if (cudaMemcpyAsync(static_cast<uint32_t *>(b->words->ptr) + b->h_genomes[0].word_off,
                        static_cast<const uint32_t *>(src->words->ptr) + sg.word_off + first_base / 16, data_words * 4,
                        cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
      st = set_error(SKS_ERR_CUDA, "slice copy failed");
  }
  if (st == SKS_OK) st = upload_tables(ctx, b);
  if (st != SKS_OK) {
    delete b;
    return st;
  }
  *out = b;
  return SKS_OK;
}

void sks_batch_destroy(sks_ctx *ctx, sks_batch *b) {
  (void)ctx;
  delete b;
}

// ---- sketching -----------------------------------------------------------------------------------
int sks_sketch(sks_ctx *ctx, const sks_batch *batch, const uint64_t mask[2], int window, const sks_pred *pred, int repr,
               sks_set **out_sets) {
  if (!ctx || !out_sets) return set_error(SKS_ERR_INVALID, "null argument");
  DeviceGuard guard(ctx->device);
  SketchPlan plan;
  SKS_TRY(make_plan(batch, mask, window, pred, &plan));
  if (batch->device != ctx->device) return set_error(SKS_ERR_INVALID, "batch lives on another device");
  if (repr == SKS_REPR_AUTO) repr = auto_repr(batch, plan, pred, window);
  for (int g = 0; g < batch->n_genomes; ++g) out_sets[g] = nullptr;
  int st;
  if (repr == SKS_REPR_BITSET) {
    if (plan.weight > 16)
      return set_error(SKS_ERR_INVALID, "bitset representation needs weight <= 16 (4^%d bits do not fit)", plan.weight);
    st = sketch_bitset(ctx, batch, plan, pred, mask, window, out_sets);
  } else if (repr == SKS_REPR_SORTED) {
    st = sketch_sorted(ctx, batch, plan, pred, mask, window, out_sets);
  } else {
    return set_error(SKS_ERR_INVALID, "unknown representation %d", repr);
  }
  if (st != SKS_OK)
    for (int g = 0; g < batch->n_genomes; ++g) {
      delete out_sets[g];
      out_sets[g] = nullptr;
    }
  return st;
}

// ---- sets ----------------------------------------------------------------------------------------
int sks_set_repr(const sks_set *s) { return s ? s->repr : 0; }
int sks_set_window(const sks_set *s) { return s ? s->window : 0; }
int sks_set_weight(const sks_set *s) { return s ? s->weight : 0; }

int sks_set_size(sks_ctx *ctx, sks_set *s, int64_t *out) {
  if (!ctx || !s || !out) return set_error(SKS_ERR_INVALID, "null argument");
  if (s->count < 0) {
    DeviceGuard guard(ctx->device);
    unsigned long long *d_cnt = nullptr, *h_cnt = nullptr;
    SKS_TRY(ctx_scratch(ctx, 64, reinterpret_cast<void **>(&d_cnt)));
    SKS_TRY(ctx_pinned(ctx, 64, reinterpret_cast<void **>(&h_cnt)));
    if (s->count_buf) {  // the bucketed build already counted the bits
      d_cnt = reinterpret_cast<unsigned long long *>(static_cast<char *>(s->count_buf->ptr) + s->count_off);
    } else {
      SKS_TRY(launch_bitset_popcount(ctx, reinterpret_cast<const uint32_t *>(static_cast<const char *>(s->buf->ptr) + s->byte_off),
                                     s->bitset_words, d_cnt));
    }
    SKS_CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    s->count = (int64_t)h_cnt[0];
  }
  *out = s->count;
  return SKS_OK;
}

int sks_set_keys(sks_ctx *ctx, sks_set *s, uint64_t *out_lohi, uint64_t capacity) {
  if (!ctx || !s || (!out_lohi && capacity)) return set_error(SKS_ERR_INVALID, "null argument");
  DeviceGuard guard(ctx->device);
  int64_t n = 0;
  SKS_TRY(sks_set_size(ctx, s, &n));
  if ((uint64_t)n > capacity) return set_error(SKS_ERR_CAPACITY, "set has %lld keys, capacity %llu", (long long)n, (unsigned long long)capacity);
  const char *base = static_cast<const char *>(s->buf->ptr) + s->byte_off;
  if (s->repr == SKS_REPR_SORTED) {
    if (n == 0) return SKS_OK;
    if (s->key_words == 2) {
      SKS_CUDA_TRY(cudaMemcpyAsync(out_lohi, base, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
      SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    } else {
      std::vector<uint64_t> tmp((size_t)n);
      SKS_CUDA_TRY(cudaMemcpyAsync(tmp.data(), base, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
      SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      for (int64_t i = 0; i < n; ++i) {
        out_lohi[2 * i] = tmp[(size_t)i];
        out_lohi[2 * i + 1] = 0;
      }
    }
    return SKS_OK;
  }
  // BITSET: download and expand set bits through the mask (PDEP); ascending index == ascending key
  std::vector<uint32_t> words((size_t)s->bitset_words);
  SKS_CUDA_TRY(cudaMemcpyAsync(words.data(), base, (size_t)s->bitset_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  int bitpos[128], nb = 0;
  for (int b = 0; b < 128; ++b)
    if ((s->mask[b >> 6] >> (b & 63)) & 1) bitpos[nb++] = b;
  uint64_t k = 0;
  for (uint64_t wi = 0; wi < s->bitset_words; ++wi) {
    uint32_t v = words[(size_t)wi];
    while (v) {
      const int bit = __builtin_ctz(v);
      v &= v - 1;
      const uint64_t idx = wi * 32 + bit;
      uint64_t lo = 0, hi = 0;
      for (int t = 0; t < nb; ++t)
        if ((idx >> t) & 1) {
          if (bitpos[t] < 64) lo |= 1ull << bitpos[t]; else hi |= 1ull << (bitpos[t] - 64);
        }
      if (k < capacity) {
        out_lohi[2 * k] = lo;
        out_lohi[2 * k + 1] = hi;
      }
      ++k;
    }
  }
  return SKS_OK;
}

int sks_set_device_keys(sks_ctx *ctx, sks_set *s, const void **dptr, int64_t *n_keys, int *words_per_key) {
  if (!ctx || !s || !dptr || !n_keys || !words_per_key) return set_error(SKS_ERR_INVALID, "null argument");
  if (s->repr != SKS_REPR_SORTED) return set_error(SKS_ERR_INVALID, "only sorted sets expose device keys");
  *dptr = static_cast<const char *>(s->buf->ptr) + s->byte_off;
  *n_keys = s->count;
  *words_per_key = s->key_words;
  return SKS_OK;
}

// window in 1..64, no mask bits at or above 2 * window, and the key width the sketch kernel uses for that window
static int check_import(const uint64_t mask[2], int window, int words_per_key) {
  if (window < 1 || window > 64) return set_error(SKS_ERR_INVALID, "window length %d outside 1..64", window);
  if (window < 64) {
    const unsigned __int128 m = ((unsigned __int128)mask[1] << 64) | mask[0];
    if (m >> (2 * window)) return set_error(SKS_ERR_INVALID, "mask has bits at or above 2*window");
  }
  if (words_per_key != (window <= 32 ? 1 : 2))
    return set_error(SKS_ERR_INVALID, "window %d takes %d-word keys, not %d", window, window <= 32 ? 1 : 2, words_per_key);
  return SKS_OK;
}

int sks_set_from_device_keys(sks_ctx *ctx, const void *dptr, int64_t n_keys, int words_per_key, const uint64_t mask[2],
                             int window, sks_set **out) {
  if (!ctx || !out || !mask || n_keys < 0 || (n_keys > 0 && !dptr) || (words_per_key != 1 && words_per_key != 2))
    return set_error(SKS_ERR_INVALID, "bad argument");
  SKS_TRY(check_import(mask, window, words_per_key));
  DeviceGuard guard(ctx->device);
  sks_set *s = new_set(ctx, SKS_REPR_SORTED, mask, window, sks_mask_weight(mask));
  if (!s) return set_error(SKS_ERR_INVALID, "out of host memory");
  s->key_words = words_per_key;
  s->count = n_keys;
  int st = alloc_buffer(ctx, (size_t)n_keys * 8 * words_per_key, &s->buf);
  if (st == SKS_OK && n_keys > 0 &&
      cudaMemcpyAsync(s->buf->ptr, dptr, (size_t)n_keys * 8 * words_per_key, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
    st = set_error(SKS_ERR_CUDA, "device copy failed");
  // foreign keys: subsets of the mask, ascending and distinct (synchronises: the caller's buffer is free on return)
  const int64_t start0 = 0;
  if (st == SKS_OK) st = validate_keys(ctx, s->buf->ptr, n_keys, words_per_key, mask, true, &start0, 1);
  if (st == SKS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = set_error(SKS_ERR_CUDA, "device copy failed");
  if (st != SKS_OK) {
    delete s;
    return st;
  }
  *out = s;
  return SKS_OK;
}

int sks_sets_from_device_keys(sks_ctx *ctx, const void *dptr, int64_t n_sets, const int64_t *counts, int words_per_key,
                              const uint64_t mask[2], int window, sks_set **out) {
  if (!ctx || !out || !mask || n_sets < 0 || (n_sets > 0 && !counts) || (words_per_key != 1 && words_per_key != 2))
    return set_error(SKS_ERR_INVALID, "bad argument");
  SKS_TRY(check_import(mask, window, words_per_key));
  DeviceGuard guard(ctx->device);
  int64_t total = 0;
  for (int64_t i = 0; i < n_sets; ++i) {
    if (counts[i] < 0) return set_error(SKS_ERR_INVALID, "negative key count");
    total += counts[i];
  }
  if (total > 0 && !dptr) return set_error(SKS_ERR_INVALID, "null key pointer");
  BufferRef buf;
  const size_t kb = (size_t)8 * words_per_key;
  SKS_TRY(alloc_buffer(ctx, (size_t)total * kb, &buf));
  if (total > 0)
    SKS_CUDA_TRY(cudaMemcpyAsync(buf->ptr, dptr, (size_t)total * kb, cudaMemcpyDeviceToDevice, ctx->stream));
  {
    std::vector<int64_t> starts((size_t)n_sets);
    int64_t at = 0;
    for (int64_t i = 0; i < n_sets; ++i) {
      starts[(size_t)i] = at;
      at += counts[i];
    }
    starts.erase(std::unique(starts.begin(), starts.end()), starts.end());  // empty sets share their start
    SKS_TRY(validate_keys(ctx, buf->ptr, total, words_per_key, mask, true, starts.data(), (int)starts.size()));
    SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // the caller's buffer is free on return
  }
  const int weight = sks_mask_weight(mask);
  int64_t off = 0;
  for (int64_t i = 0; i < n_sets; ++i) {
    sks_set *s = new_set(ctx, SKS_REPR_SORTED, mask, window, weight);
    if (!s) {
      for (int64_t j = 0; j < i; ++j) delete out[j];
      return set_error(SKS_ERR_INVALID, "out of host memory");
    }
    s->buf = buf;
    s->key_words = words_per_key;
    s->byte_off = (size_t)off * kb;
    s->count = counts[i];
    out[i] = s;
    off += counts[i];
  }
  return SKS_OK;
}

int sks_set_from_unsorted_device_keys(sks_ctx *ctx, const void *dptr, int64_t n_keys, int words_per_key,
                                      const uint64_t mask[2], int window, sks_set **out) {
  if (!ctx || !out || !mask || n_keys < 0 || (n_keys > 0 && !dptr) || (words_per_key != 1 && words_per_key != 2))
    return set_error(SKS_ERR_INVALID, "bad argument");
  SKS_TRY(check_import(mask, window, words_per_key));
  DeviceGuard guard(ctx->device);
  BufferRef raw, uniq;
  SKS_TRY(alloc_buffer(ctx, (size_t)n_keys * 8 * words_per_key, &raw));
  if (n_keys > 0)
    SKS_CUDA_TRY(cudaMemcpyAsync(raw->ptr, dptr, (size_t)n_keys * 8 * words_per_key, cudaMemcpyDeviceToDevice, ctx->stream));
  SKS_TRY(validate_keys(ctx, raw->ptr, n_keys, words_per_key, mask, false, nullptr, 0));  // also: the caller's buffer is free
  const uint64_t off = 0, cnt = (uint64_t)n_keys;
  std::vector<uint64_t> uoff, ucount;
  SKS_TRY(sort_unique_regions(ctx, words_per_key, raw->ptr, &off, &cnt, 1, cnt, &uniq, &uoff, &ucount, mask));
  sks_set *s = new_set(ctx, SKS_REPR_SORTED, mask, window, sks_mask_weight(mask));
  if (!s) return set_error(SKS_ERR_INVALID, "out of host memory");
  s->buf = uniq;
  s->key_words = words_per_key;
  s->byte_off = (size_t)uoff[0] * 8 * words_per_key;
  s->count = (int64_t)ucount[0];
  *out = s;
  return SKS_OK;
}

int sks_set_from_host_keys(sks_ctx *ctx, const uint64_t *keys_lohi, int64_t n_keys, const uint64_t mask[2], int window,
                           sks_set **out) {
  if (!ctx || !out || !mask || n_keys < 0 || (n_keys > 0 && !keys_lohi)) return set_error(SKS_ERR_INVALID, "bad argument");
  if (window < 1 || window > 64) return set_error(SKS_ERR_INVALID, "window length %d outside 1..64", window);
  DeviceGuard guard(ctx->device);
  const int kw = window <= 32 ? 1 : 2;  // same key width the sketch kernel uses for this window
  BufferRef raw;
  SKS_TRY(alloc_buffer(ctx, (size_t)std::max<int64_t>(n_keys, 1) * 8 * kw, &raw));
  if (n_keys > 0) {
    if (kw == 2) {
      SKS_CUDA_TRY(cudaMemcpyAsync(raw->ptr, keys_lohi, (size_t)n_keys * 16, cudaMemcpyHostToDevice, ctx->stream));
    } else {
      std::vector<uint64_t> lo((size_t)n_keys);
      for (int64_t i = 0; i < n_keys; ++i) {
        if (keys_lohi[2 * i + 1]) return set_error(SKS_ERR_INVALID, "key %lld has bits above 2*window", (long long)i);
        lo[(size_t)i] = keys_lohi[2 * i];
      }
      SKS_CUDA_TRY(cudaMemcpyAsync(raw->ptr, lo.data(), (size_t)n_keys * 8, cudaMemcpyHostToDevice, ctx->stream));
      SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // lo is a local
    }
  }
  return sks_set_from_unsorted_device_keys(ctx, raw->ptr, n_keys, kw, mask, window, out);
}

int sks_set_device_index(const sks_set *s) { return s ? s->device : -1; }

int sks_set_clone_to(sks_ctx *dst, sks_set *src, sks_set **out) {
  if (!dst || !src || !out) return set_error(SKS_ERR_INVALID, "null argument");
  DeviceGuard guard(dst->device);
  size_t bytes = 0;
  if (src->repr == SKS_REPR_SORTED) {
    if (src->count < 0) return set_error(SKS_ERR_INVALID, "set without a size");
    bytes = (size_t)src->count * 8 * src->key_words;
  } else {
    bytes = (size_t)src->bitset_words * 4;
  }
  sks_set *c = new (std::nothrow) sks_set(*src);
  if (!c) return set_error(SKS_ERR_INVALID, "out of host memory");
  c->device = dst->device;
  c->byte_off = 0;
  c->buf.reset();
  c->count_buf.reset();   // a bitset's device-side popcount stays behind: the copy counts again if asked
  c->count_off = 0;
  if (src->repr != SKS_REPR_SORTED && src->count < 0 && src->count_buf) {
    // bring the size along instead (the source's control block holds it once its build has run)
    c->count = -1;
  }
  int st = alloc_buffer(dst, std::max<size_t>(bytes, 16), &c->buf);
  if (st == SKS_OK && bytes) {
    const char *sp = static_cast<const char *>(src->buf->ptr) + src->byte_off;
    // the source's stream may still be writing the set: wait for that device first
    if (src->device != dst->device) {
      DeviceGuard g2(src->device);
      cudaDeviceSynchronize();
    }
    cudaError_t e = cudaMemcpyPeerAsync(c->buf->ptr, dst->device, sp, src->device, bytes, dst->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dst->stream);
    if (e != cudaSuccess) st = set_error(SKS_ERR_CUDA, "peer copy of a set failed: %s", cudaGetErrorString(e));
  }
  if (st != SKS_OK) {
    delete c;
    return st;
  }
  *out = c;
  return SKS_OK;
}

void sks_set_destroy(sks_ctx *ctx, sks_set *s) {
  (void)ctx;
  delete s;
}

// ---- sketch files -----------------------------------------------------------------------------------
namespace {
struct SketchFileHeader {
  char magic[8];
  uint32_t version, window;
  uint64_t mask[2];
  uint32_t pred_kind;
  int32_t nonce;
  uint64_t modulus;
  uint32_t hash_variant, key_words;
  uint64_t n_keys;
};
static_assert(sizeof(SketchFileHeader) == 64, "sketch file header layout");
}  // namespace

int sks_set_save(sks_ctx *ctx, sks_set *s, const sks_pred *pred, const char *path) {
  if (!ctx || !s || !path) return set_error(SKS_ERR_INVALID, "null argument");
  int64_t n = 0;
  SKS_TRY(sks_set_size(ctx, s, &n));
  std::vector<uint64_t> keys((size_t)n * 2);
  SKS_TRY(sks_set_keys(ctx, s, keys.data(), (uint64_t)n));
  SketchFileHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "SKSKETCH", 8);
  h.version = 1;
  h.window = (uint32_t)s->window;
  h.mask[0] = s->mask[0];
  h.mask[1] = s->mask[1];
  h.pred_kind = pred ? (uint32_t)pred->kind : 0xFFFFFFFFu;
  h.nonce = pred ? pred->nonce : 0;
  h.modulus = pred ? pred->modulus : 0;
  h.hash_variant = pred ? (uint32_t)(pred->hash_variant ? pred->hash_variant : SKS_HASH_BOOST_181) : 0;
  h.key_words = s->repr == SKS_REPR_SORTED ? (uint32_t)s->key_words : (s->window <= 32 ? 1u : 2u);
  h.n_keys = (uint64_t)n;
  FILE *f = fopen(path, "wb");
  if (!f) return set_error(SKS_ERR_IO, "Unable to open %s for writing", path);
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  if (h.key_words == 2) {
    ok = ok && (n == 0 || fwrite(keys.data(), 16, (size_t)n, f) == (size_t)n);
  } else {
    std::vector<uint64_t> lo((size_t)n);
    for (int64_t i = 0; i < n; ++i) lo[(size_t)i] = keys[(size_t)i * 2];
    ok = ok && (n == 0 || fwrite(lo.data(), 8, (size_t)n, f) == (size_t)n);
  }
  ok = (fclose(f) == 0) && ok;
  return ok ? SKS_OK : set_error(SKS_ERR_IO, "short write to %s", path);
}

int sks_set_load(sks_ctx *ctx, const char *path, sks_set **out, sks_pred *out_pred) {
  if (!ctx || !path || !out) return set_error(SKS_ERR_INVALID, "null argument");
  FILE *f = fopen(path, "rb");
  if (!f) return set_error(SKS_ERR_IO, "Unable to open %s", path);
  SketchFileHeader h;
  int st = SKS_OK;
  std::vector<uint64_t> keys;
  long file_bytes = 0;
  if (fseek(f, 0, SEEK_END) == 0) file_bytes = ftell(f);
  rewind(f);
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "SKSKETCH", 8) != 0 || h.version != 1 ||
      (h.key_words != 1 && h.key_words != 2) || h.window < 1 || h.window > 64) {
    st = set_error(SKS_ERR_IO, "%s is not a version-1 sketch file", path);
  } else if (check_import(h.mask, (int)h.window, (int)h.key_words) != SKS_OK) {
    st = set_error(SKS_ERR_IO, "%s: mask, window and key width do not fit together", path);
  } else if (file_bytes < (long)sizeof(h) || h.n_keys > (uint64_t)(file_bytes - (long)sizeof(h)) / (8 * h.key_words)) {
    st = set_error(SKS_ERR_IO, "%s is truncated", path);   // the header promises more keys than the file holds
  } else {
    try {
      keys.resize((size_t)h.n_keys * h.key_words);
    } catch (const std::exception &) {
      st = set_error(SKS_ERR_IO, "%s: out of host memory for %llu keys", path, (unsigned long long)h.n_keys);
    }
    if (st == SKS_OK && h.n_keys && fread(keys.data(), 8 * h.key_words, (size_t)h.n_keys, f) != (size_t)h.n_keys)
      st = set_error(SKS_ERR_IO, "%s is truncated", path);
  }
  fclose(f);
  SKS_TRY(st);
  for (uint64_t i = 0; i < h.n_keys; ++i) {  // members are subsets of the mask (the intersection kernels rely on it)
    const uint64_t *k = &keys[(size_t)i * h.key_words];
    if ((k[0] & ~h.mask[0]) || (h.key_words == 2 ? (k[1] & ~h.mask[1]) : 0))
      return set_error(SKS_ERR_IO, "%s: a key has bits outside the mask", path);
  }
  for (uint64_t i = 1; i < h.n_keys; ++i) {  // the file must hold ascending distinct keys
    const uint64_t *a = &keys[(size_t)(i - 1) * h.key_words], *b = &keys[(size_t)i * h.key_words];
    const bool less = h.key_words == 1 ? a[0] < b[0] : (a[1] != b[1] ? a[1] < b[1] : a[0] < b[0]);
    if (!less) return set_error(SKS_ERR_IO, "%s: keys are not ascending and distinct", path);
  }
  DeviceGuard guard(ctx->device);
  BufferRef buf;
  SKS_TRY(alloc_buffer(ctx, keys.size() * 8, &buf));
  if (!keys.empty()) {
    SKS_CUDA_TRY(cudaMemcpyAsync(buf->ptr, keys.data(), keys.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  sks_set *s = new_set(ctx, SKS_REPR_SORTED, h.mask, (int)h.window, sks_mask_weight(h.mask));
  if (!s) return set_error(SKS_ERR_INVALID, "out of host memory");
  s->buf = buf;
  s->key_words = (int)h.key_words;
  s->count = (int64_t)h.n_keys;
  *out = s;
  if (out_pred) {
    memset(out_pred, 0, sizeof(*out_pred));
    out_pred->kind = (int32_t)h.pred_kind;
    out_pred->nonce = h.nonce;
    out_pred->modulus = h.modulus;
    out_pred->hash_variant = (int32_t)h.hash_variant;
  }
  return SKS_OK;
}

// ---- comparison ----------------------------------------------------------------------------------
// `uniform`: the caller has already checked that every set of both lists has the mask, representation, key width and
// device of a[0] (sks_intersect_rects checks each distinct set once instead of once per pair).
static int intersect_pairs(sks_ctx *ctx, sks_set *const *a, int64_t na, sks_set *const *b, int64_t nb, int32_t *out,
                           bool uniform) {
  if (!ctx || (na > 0 && (!a || !b || !out))) return set_error(SKS_ERR_INVALID, "null argument");
  // src/kmer_set.cpp:147-150,174-177
  if (na != nb) return set_error(SKS_ERR_MISMATCH, "Lists of kmer sets for intersection computation have different lengths");
  if (na == 0) return SKS_OK;
  DeviceGuard guard(ctx->device);
  // A pair list that covers a good part of all pairs of its distinct sets (generate_all_pairs_from_vector,
  // src/generators.hpp:44-58, is all of them) goes through the all-vs-all dictionary: one pass over the sets instead
  // of one lookup pass per pair.
  if (!uniform && na >= 16) {
    std::map<const sks_set *, int32_t> index;
    std::vector<sks_set *> distinct;
    std::vector<int32_t> ia((size_t)na), ib((size_t)na);
    bool ok = true;
    for (int64_t i = 0; i < na && ok; ++i)
      for (int side = 0; side < 2; ++side) {
        sks_set *s = side ? b[i] : a[i];
        if (!s) {
          ok = false;
          break;
        }
        auto it = index.find(s);
        if (it == index.end()) {
          it = index.emplace(s, (int32_t)distinct.size()).first;
          distinct.push_back(s);
        }
        (side ? ib : ia)[(size_t)i] = it->second;
      }
    const int64_t d = (int64_t)distinct.size();
    if (ok && d >= 4 && na * 4 >= d * d && all_pairs_dict_eligible(distinct.data(), d)) {
      for (int64_t i = 0; i < na; ++i) SKS_TRY(check_pair(a[i], b[i]));
      std::vector<int32_t> full((size_t)d * d);
      const int st = sks_all_vs_all(ctx, distinct.data(), d, 0, d, full.data(), nullptr, nullptr);
      if (st == SKS_OK) {
        for (int64_t i = 0; i < na; ++i) out[i] = full[(size_t)ia[(size_t)i] * d + ib[(size_t)i]];
        return SKS_OK;
      }
      if (st != SKS_ERR_CAPACITY) return st;
    }
  }
  const int repr = a[0]->repr;
  if (!uniform) {
    for (int64_t i = 0; i < na; ++i) SKS_TRY(check_pair(a[i], b[i]));
    for (int64_t i = 0; i < na; ++i)
      if (a[i]->repr != repr || a[i]->key_words != a[0]->key_words)
        return set_error(SKS_ERR_MISMATCH, "pair list mixes set representations");
  }

  if (repr == SKS_REPR_BITSET) {
    unsigned long long *d_cnt = nullptr, *h_cnt = nullptr;
    SKS_TRY(ctx_scratch(ctx, (size_t)na * 24, reinterpret_cast<void **>(&d_cnt)));
    SKS_TRY(ctx_pinned(ctx, (size_t)na * 24, reinterpret_cast<void **>(&h_cnt)));
    for (int64_t i = 0; i < na; ++i) {
      const uint32_t *pa = reinterpret_cast<const uint32_t *>(static_cast<const char *>(a[i]->buf->ptr) + a[i]->byte_off);
      const uint32_t *pb = reinterpret_cast<const uint32_t *>(static_cast<const char *>(b[i]->buf->ptr) + b[i]->byte_off);
      SKS_TRY(launch_bitset_pair_counts(ctx, pa, pb, a[i]->bitset_words, d_cnt + 3 * i));
    }
    SKS_CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, (size_t)na * 24, cudaMemcpyDeviceToHost, ctx->stream));
    SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int64_t i = 0; i < na; ++i) {
      a[i]->count = (int64_t)h_cnt[3 * i];  // the same pass yields both set sizes
      b[i]->count = (int64_t)h_cnt[3 * i + 1];
      out[i] = (int32_t)h_cnt[3 * i + 2];
    }
    return SKS_OK;
  }

  // SORTED.  Runs of pairs that share their first set (the rows of an all-vs-all block; a single pair is a run of
  // one) go to the kernel that keeps that set resident in shared memory; pairs whose first set does not fit go to
  // the merge kernel.  One pass over the pair tables either way.
  const int kw = a[0]->key_words;
  static const bool row_enabled = getenv("SKS_ROW_INTERSECT") ? atoi(getenv("SKS_ROW_INTERSECT")) != 0 : true;
  bool one_mask = na < (int64_t)1 << 31;
  for (int64_t i = 1; i < na && one_mask && !uniform; ++i)
    one_mask = a[i]->mask[0] == a[0]->mask[0] && a[i]->mask[1] == a[0]->mask[1];
  std::vector<RowTaskHost> tasks;
  std::vector<uint32_t> rest;
  int64_t max_row = 0;
  {
    int64_t row_pairs = 0;
    std::vector<std::pair<int64_t, int64_t>> runs;
    for (int64_t i = 0; i < na;) {
      int64_t j = i + 1;
      while (j < na && a[j] == a[i]) ++j;
      if (row_enabled && one_mask && a[i]->count > 0 && row_intersect_fits(kw, a[0]->mask, a[i]->count)) {
        max_row = std::max<int64_t>(max_row, a[i]->count);
        runs.emplace_back(i, j);
        row_pairs += j - i;
      } else {
        for (int64_t k = i; k < j; ++k) rest.push_back((uint32_t)k);
      }
      i = j;
    }
    const int64_t per_task = row_pairs > (int64_t)ctx->sm_count * 32 * 4 ? 32 : 16;  // columns per CTA
    for (auto &run : runs)
      for (int64_t c = run.first; c < run.second; c += per_task) {
        RowTaskHost t;
        t.a = static_cast<const char *>(a[run.first]->buf->ptr) + a[run.first]->byte_off;
        t.n_a = (uint32_t)a[run.first]->count;
        t.first = (uint32_t)c;
        t.n_cols = (uint32_t)std::min<int64_t>(per_task, run.second - c);
        t.pad = 0;
        tasks.push_back(t);
      }
  }
  const size_t tab_bytes = (size_t)na * 32, out_bytes = ((size_t)na * 4 + 15) & ~(size_t)15;
  const size_t task_bytes = tasks.size() * sizeof(RowTaskHost), rest_bytes = rest.size() * 4;
  char *d_tab = nullptr, *h_tab = nullptr;
  SKS_TRY(ctx_scratch(ctx, tab_bytes + out_bytes + task_bytes + rest_bytes, reinterpret_cast<void **>(&d_tab)));
  SKS_TRY(ctx_pinned(ctx, tab_bytes + out_bytes + task_bytes + rest_bytes, reinterpret_cast<void **>(&h_tab)));
  const void **h_pa = reinterpret_cast<const void **>(h_tab);
  const void **h_pb = h_pa + na;
  int64_t *h_na = reinterpret_cast<int64_t *>(h_pb + na), *h_nb = h_na + na;
  for (int64_t i = 0; i < na; ++i) {
    h_pa[i] = static_cast<const char *>(a[i]->buf->ptr) + a[i]->byte_off;
    h_pb[i] = static_cast<const char *>(b[i]->buf->ptr) + b[i]->byte_off;
    h_na[i] = a[i]->count;
    h_nb[i] = b[i]->count;
  }
  // layout (host and device): tables | out | tasks | rest
  if (task_bytes) memcpy(h_tab + tab_bytes + out_bytes, tasks.data(), task_bytes);
  if (rest_bytes) memcpy(h_tab + tab_bytes + out_bytes + task_bytes, rest.data(), rest_bytes);
  SKS_CUDA_TRY(cudaMemcpyAsync(d_tab, h_tab, tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
  if (task_bytes + rest_bytes)
    SKS_CUDA_TRY(cudaMemcpyAsync(d_tab + tab_bytes + out_bytes, h_tab + tab_bytes + out_bytes, task_bytes + rest_bytes,
                                 cudaMemcpyHostToDevice, ctx->stream));
  const void **d_pa = reinterpret_cast<const void **>(d_tab);
  const void **d_pb = d_pa + na;
  int64_t *d_na = reinterpret_cast<int64_t *>(d_pb + na), *d_nb = d_na + na;
  int32_t *d_out = reinterpret_cast<int32_t *>(d_tab + tab_bytes);
  // merge kernel: few pairs of large sets are cut into slices so that the whole GPU works on them
  int slices = 1;
  if (!rest.empty()) {
    int64_t max_small = 0;
    for (uint32_t k : rest) max_small = std::max<int64_t>(max_small, std::min(a[k]->count, b[k]->count));
    const int64_t by_size = max_small / 16384;                                    // >= 2048 keys per warp
    const int64_t by_fill = (int64_t)ctx->sm_count * 8 / (int64_t)rest.size();     // enough CTAs to fill the GPU
    slices = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(by_size, by_fill), 1024));
  }
  if (!tasks.empty() || slices > 1) {
    SKS_CUDA_TRY(cudaMemsetAsync(d_out, 0, out_bytes, ctx->stream));  // both kernels add partial counts
    SKS_TRY(launch_row_intersect(ctx, kw, d_tab + tab_bytes + out_bytes, (int64_t)tasks.size(), d_pb, d_nb, d_out, a[0]->mask,
                                 max_row));
  }
  if (rest.size() == (size_t)na)
    SKS_TRY(launch_sorted_intersect_pairs(ctx, kw, d_pa, d_na, d_pb, d_nb, na, d_out, nullptr, slices));
  else if (!rest.empty())
    SKS_TRY(launch_sorted_intersect_pairs(ctx, kw, d_pa, d_na, d_pb, d_nb, (int64_t)rest.size(), d_out,
                                          reinterpret_cast<const uint32_t *>(d_tab + tab_bytes + out_bytes + task_bytes), slices));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_tab + tab_bytes, d_out, (size_t)na * 4, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  memcpy(out, h_tab + tab_bytes, (size_t)na * 4);
  return SKS_OK;
}

int sks_intersect_pairs(sks_ctx *ctx, sks_set *const *a, int64_t na, sks_set *const *b, int64_t nb, int32_t *out) {
  return intersect_pairs(ctx, a, na, b, nb, out, false);
}

int sks_intersect(sks_ctx *ctx, sks_set *a, sks_set *b, int64_t *out) {
  if (!out) return set_error(SKS_ERR_INVALID, "null argument");
  int32_t r = 0;
  sks_set *pa[1] = {a}, *pb[1] = {b};
  SKS_TRY(sks_intersect_pairs(ctx, pa, 1, pb, 1, &r));
  *out = r;
  return SKS_OK;
}

int sks_intersect_rects(sks_ctx *ctx, sks_set *const *sets, int64_t n, const int64_t *rects, int64_t n_rects, int32_t *out) {
  if (!ctx || (n > 0 && (!sets || !out)) || (n_rects > 0 && !rects)) return set_error(SKS_ERR_INVALID, "null argument");
  // every set that occurs is checked once against the first one; the pair loop below then needs no checks
  size_t cap = 0;
  const sks_set *ref = nullptr;
  for (int64_t q = 0; q < n_rects; ++q) {
    const int64_t row_begin = rects[4 * q], row_end = rects[4 * q + 1], col_begin = rects[4 * q + 2], col_end = rects[4 * q + 3];
    if (row_begin < 0 || row_end > n || row_begin > row_end) return set_error(SKS_ERR_INVALID, "bad row range");
    if (col_begin < 0 || col_end > n || col_begin > col_end) return set_error(SKS_ERR_INVALID, "bad column range");
    if (row_begin == row_end || col_begin == col_end) continue;
    cap += (size_t)(row_end - row_begin) * (size_t)(col_end - col_begin);
    for (int side = 0; side < 2; ++side)
      for (int64_t i = side ? col_begin : row_begin; i < (side ? col_end : row_end); ++i) {
        if (!ref) ref = sets[i];
        SKS_TRY(check_pair(ref, sets[i]));
      }
  }
  std::vector<sks_set *> pa, pb;
  std::vector<int64_t> slot;  // pair k -> out[slot[2k]] and, unless negative, out[slot[2k+1]] (its mirror)
  pa.reserve(cap);
  pb.reserve(cap);
  slot.reserve(2 * cap);
  for (int64_t q = 0; q < n_rects; ++q) {
    const int64_t row_begin = rects[4 * q], row_end = rects[4 * q + 1], col_begin = rects[4 * q + 2], col_end = rects[4 * q + 3];
    // |A n B| is symmetric: an unordered pair that lies in the rectangle with both orientations is evaluated once
    // and mirrored inside it.  (j, i) lies in the rectangle iff j is one of its rows and i one of its columns.
    for (int64_t i = row_begin; i < row_end; ++i) {
      const bool i_is_col = i >= col_begin && i < col_end;
      for (int64_t j = col_begin; j < col_end; ++j) {
        if (j == i) continue;
        const bool both = i_is_col && j >= row_begin && j < row_end;
        if (j < i && both) continue;  // (j, i) is in the rectangle too and comes first
        pa.push_back(sets[i]);
        pb.push_back(sets[j]);
        slot.push_back(i * n + j);
        slot.push_back(both ? j * n + i : -1);
      }
    }
  }
  std::vector<int32_t> r(pa.size());
  SKS_TRY(intersect_pairs(ctx, pa.data(), (int64_t)pa.size(), pb.data(), (int64_t)pb.size(), r.data(), true));
  for (size_t k = 0; k < r.size(); ++k) {
    out[slot[2 * k]] = r[k];
    if (slot[2 * k + 1] >= 0) out[slot[2 * k + 1]] = r[k];
  }
  for (int64_t q = 0; q < n_rects; ++q) {  // |A n A| = |A|
    const int64_t lo = std::max(rects[4 * q], rects[4 * q + 2]), hi = std::min(rects[4 * q + 1], rects[4 * q + 3]);
    for (int64_t i = lo; i < hi; ++i) {
      int64_t sz = 0;
      SKS_TRY(sks_set_size(ctx, sets[i], &sz));
      out[i * n + i] = (int32_t)sz;
    }
  }
  return SKS_OK;
}

int sks_intersect_block(sks_ctx *ctx, sks_set *const *sets, int64_t n, int64_t row_begin, int64_t row_end,
                        int64_t col_begin, int64_t col_end, int32_t *out) {
  const int64_t rect[4] = {row_begin, row_end, col_begin, col_end};
  return sks_intersect_rects(ctx, sets, n, rect, 1, out);
}

int sks_intersect_all_pairs(sks_ctx *ctx, sks_set *const *sets, int64_t n, int64_t row_begin, int64_t row_end,
                            int32_t *out) {
  if (ctx && sets && out && n >= 4 && row_begin >= 0 && row_begin < row_end && row_end <= n && all_pairs_dict_eligible(sets, n))
    return sks_all_vs_all(ctx, sets, n, row_begin, row_end, out + row_begin * n, nullptr, nullptr);
  return sks_intersect_block(ctx, sets, n, row_begin, row_end, 0, n, out);
}

int sks_all_vs_all(sks_ctx *ctx, sks_set *const *sets, int64_t n, int64_t row_begin, int64_t row_end, int32_t *out_counts,
                   int32_t *out_sizes, double *out_ani) {
  if (!ctx || (n > 0 && !sets)) return set_error(SKS_ERR_INVALID, "null argument");
  if (row_begin < 0 || row_end > n || row_begin > row_end) return set_error(SKS_ERR_INVALID, "bad row range");
  for (int64_t i = 0; i < n; ++i) {
    if (!sets[i]) return set_error(SKS_ERR_INVALID, "null set");
    SKS_TRY(check_pair(sets[0], sets[i]));
  }
  DeviceGuard guard(ctx->device);
  const int64_t n_rows = row_end - row_begin;
  if (out_sizes)
    for (int64_t i = 0; i < n; ++i) {
      int64_t sz = 0;
      SKS_TRY(sks_set_size(ctx, sets[i], &sz));
      out_sizes[i] = (int32_t)sz;
    }
  if (n_rows == 0 || n == 0) return SKS_OK;
  if (all_pairs_dict_eligible(sets, n)) {
    BufferRef counts, ani;
    const uint32_t *overflow = nullptr;
    int st = all_pairs_dict(ctx, sets, n, row_begin, row_end, &counts, out_ani ? &ani : nullptr, nullptr, &overflow);
    if (st == SKS_OK) {
      if (out_counts)
        SKS_CUDA_TRY(cudaMemcpyAsync(out_counts, counts->ptr, 4 * (size_t)n_rows * n, cudaMemcpyDeviceToHost, ctx->stream));
      if (out_ani)
        SKS_CUDA_TRY(cudaMemcpyAsync(out_ani, ani->ptr, 8 * (size_t)n_rows * n, cudaMemcpyDeviceToHost, ctx->stream));
      SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
      if (!overflow || *overflow == 0) return SKS_OK;
      st = SKS_ERR_CAPACITY;
    }
    if (st != SKS_ERR_CAPACITY) return st;  // too large for the dictionary: the pairwise kernels below
  }
  std::vector<int32_t> full((size_t)n * n), sizes((size_t)n);
  SKS_TRY(sks_intersect_block(ctx, sets, n, row_begin, row_end, 0, n, full.data()));
  for (int64_t i = 0; i < n; ++i) {
    int64_t sz = 0;
    SKS_TRY(sks_set_size(ctx, sets[i], &sz));
    sizes[i] = (int32_t)sz;
  }
  if (out_counts) memcpy(out_counts, full.data() + row_begin * n, 4 * (size_t)n_rows * n);
  if (out_ani) {
    std::vector<int32_t> first((size_t)n_rows * n);
    for (int64_t r = 0; r < n_rows; ++r) std::fill(first.begin() + r * n, first.begin() + (r + 1) * n, sizes[row_begin + r]);
    sks_ani_from_counts(full.data() + row_begin * n, first.data(), n_rows * n, sets[0]->weight, out_ani);
  }
  return SKS_OK;
}

// ---- one-call pair pipeline ----------------------------------------------------------------------
int sks_pair_ani_resident(sks_ctx *ctx, const sks_batch *batch, const uint64_t mask[2], int window, const sks_pred *pred,
                          int repr, sks_pair_result *out) {
  if (!ctx || !batch || !out) return set_error(SKS_ERR_INVALID, "null argument");
  if (batch->n_genomes != 2) return set_error(SKS_ERR_INVALID, "pair pipeline needs a 2-genome batch");
  if (batch->device != ctx->device) return set_error(SKS_ERR_INVALID, "batch lives on another device");
  DeviceGuard guard(ctx->device);
  SketchPlan plan;
  SKS_TRY(make_plan(batch, mask, window, pred, &plan));
  if (repr == SKS_REPR_AUTO) repr = auto_repr(batch, plan, pred, window);
  sks_set *sets[2] = {nullptr, nullptr};
  int64_t inter = 0, sa = 0, sb = 0;
  int st = SKS_OK;
  PairFuse fuse;
  if (repr == SKS_REPR_BITSET || repr == SKS_REPR_BITSET_ONCHIP) {
    if (plan.weight > 16)
      return set_error(SKS_ERR_INVALID, "bitset representation needs weight <= 16 (4^%d bits do not fit)", plan.weight);
    // large bitsets: the fused build counts |A|, |B|, |A n B| while the slices are still in shared memory
    fuse.store = repr == SKS_REPR_BITSET;
    st = sketch_bitset(ctx, batch, plan, pred, mask, window, sets, &fuse);
  } else {
    st = sks_sketch(ctx, batch, mask, window, pred, repr, sets);
  }
  if (st == SKS_OK) {
    if (fuse.done) {
      sa = fuse.counts[0];
      sb = fuse.counts[1];
      inter = fuse.counts[2];
    } else {
      st = sks_intersect(ctx, sets[0], sets[1], &inter);
      if (st == SKS_OK) st = sks_set_size(ctx, sets[0], &sa);
      if (st == SKS_OK) st = sks_set_size(ctx, sets[1], &sb);
    }
  }
  const int weight = plan.weight;
  if (sets[0]) sks_set_destroy(ctx, sets[0]);
  if (sets[1]) sks_set_destroy(ctx, sets[1]);
  if (st != SKS_OK) return st;
  out->size_a = sa;
  out->size_b = sb;
  out->intersection = inter;
  out->ani_ab = sks_binomial_estimator(sks_containment((int)inter, (int)sa), weight);
  out->ani_ba = sks_binomial_estimator(sks_containment((int)inter, (int)sb), weight);
  return SKS_OK;
}

int sks_pair_ani(sks_ctx *ctx, const uint32_t *packed_a, uint64_t n_bases_a, const uint32_t *packed_b, uint64_t n_bases_b,
                 const uint64_t mask[2], int window, const sks_pred *pred, int repr, sks_pair_result *out) {
  const uint32_t *packed[2] = {packed_a, packed_b};
  const uint64_t nb[2] = {n_bases_a, n_bases_b};
  sks_batch *batch = nullptr;
  if (!ctx) return set_error(SKS_ERR_INVALID, "null context");
  if (batch_in_place(ctx, 2, packed, nb, &batch) != SKS_OK)  // not pinned / not aligned: copy the genomes up
    SKS_TRY(sks_batch_upload(ctx, 2, packed, nb, nullptr, nullptr, &batch));
  if (batch->host_words) ctx->in_place_calls++;
  const int st = sks_pair_ani_resident(ctx, batch, mask, window, pred, repr, out);
  if (batch->host_words) cudaStreamSynchronize(ctx->stream);  // nothing may read the caller's buffers after the return
  sks_batch_destroy(ctx, batch);
  return st;
}

int sks_all_vs_all_resident(sks_ctx *ctx, sks_comm *comm, const sks_batch *batch, int64_t n_total, const uint64_t mask[2],
                            int window, const sks_pred *pred, int32_t *out_counts, int32_t *out_sizes, double *out_ani) {
  if (!ctx) return set_error(SKS_ERR_INVALID, "null context");
  const int n_local = batch ? batch->n_genomes : 0;
  std::vector<sks_set *> sets((size_t)std::max(n_local, 1), nullptr);
  int st = SKS_OK;
  if (n_local > 0) st = sks_sketch(ctx, batch, mask, window, pred, SKS_REPR_SORTED, sets.data());
  if (st == SKS_OK) st = sks_all_vs_all_sharded(ctx, comm, sets.data(), n_local, n_total, out_counts, out_sizes, out_ani);
  for (sks_set *s : sets)
    if (s) sks_set_destroy(ctx, s);
  return st;
}

int sks_all_vs_all_from_host(sks_ctx *ctx, sks_comm *comm, int n_local, const uint32_t *const *packed, const uint64_t *n_bases,
                             int64_t n_total, const uint64_t mask[2], int window, const sks_pred *pred, int32_t *out_counts,
                             int32_t *out_sizes, double *out_ani) {
  if (!ctx || n_local < 0 || (n_local > 0 && (!packed || !n_bases))) return set_error(SKS_ERR_INVALID, "null argument");
  sks_batch *batch = nullptr;
  std::vector<sks_set *> sets((size_t)std::max(n_local, 1), nullptr);
  int st = SKS_OK;
  if (n_local > 0) {
    // pinned and large: copied chunk by chunk under the sketch kernel; pinned and small: read in place by the kernel;
    // pageable (or misaligned): copied up first
    if (batch_streamed(ctx, n_local, packed, n_bases, comm ? sks_comm_world(comm) : 1, &batch) == SKS_OK)
      ctx->streamed_calls++;
    else if (batch_in_place(ctx, n_local, packed, n_bases, &batch) != SKS_OK)
      st = sks_batch_upload(ctx, n_local, packed, n_bases, nullptr, nullptr, &batch);
    if (st == SKS_OK && batch->host_words) ctx->in_place_calls++;
    if (st == SKS_OK) st = sks_sketch(ctx, batch, mask, window, pred, SKS_REPR_SORTED, sets.data());
    // sks_sketch has read the counts back: the copies and the kernel are done with the caller's buffers -- unless it
    // failed on the way
    if (st != SKS_OK && batch && !batch->arriving.empty()) {
      if (batch->feed && batch->feed->worker.joinable()) batch->feed->worker.join();
      cudaStreamSynchronize(ctx->copy_stream);
    }
  }
  if (st == SKS_OK) st = sks_all_vs_all_sharded(ctx, comm, sets.data(), n_local, n_total, out_counts, out_sizes, out_ani);
  for (sks_set *s : sets)
    if (s) sks_set_destroy(ctx, s);
  if (batch) sks_batch_destroy(ctx, batch);
  return st;
}

// nucleotide_string_list_to_kmers, src/kmer_sliding.cpp:224-238: ordered list with duplicates.
int sks_kmer_list(sks_ctx *ctx, const sks_batch *batch, int genome, const uint64_t mask[2], int window,
                  const sks_pred *pred, uint64_t *out_n, uint64_t *out_masked, uint64_t *out_bits, uint64_t capacity) {
  if (!ctx || !batch || !out_n) return set_error(SKS_ERR_INVALID, "null argument");
  if (genome < 0 || genome >= batch->n_genomes) return set_error(SKS_ERR_INVALID, "genome %d outside the batch", genome);
  DeviceGuard guard(ctx->device);
  const sks_batch *one = batch;
  sks_batch *tmp = nullptr;
  if (batch->n_genomes > 1) {  // the kernel numbers tiles batch-wide: work on a single-genome copy
    SKS_TRY(sks_batch_slice(ctx, batch, genome, 0, batch->h_genomes[genome].n_bases, 1, &tmp));
    one = tmp;
  }
  SketchPlan plan;
  BufferRef raw, pos, outbuf;
  std::vector<uint64_t> off, count;
  uint64_t span = 0;
  // list entries carry (strand << 31 | start position) in 32 bits and the list is finalised with 32-bit counts
  if (one->h_genomes[0].n_bases >= (1u << 31)) {
    if (tmp) sks_batch_destroy(ctx, tmp);
    return set_error(SKS_ERR_CAPACITY, "the ordered k-mer list is limited to genomes below 2^31 bases (this one has %u)",
                     one->h_genomes[0].n_bases);
  }
  int st = make_plan(one, mask, window, pred, &plan);
  if (st == SKS_OK) st = sketch_raw_keys(ctx, one, plan, pred, window, OUT_LIST, &raw, &pos, &off, &count, &span);
  if (st == SKS_OK) {
    const uint64_t n = count[0];
    *out_n = n;
    if (out_masked || out_bits) {
      if (n > capacity) {
        st = set_error(SKS_ERR_CAPACITY, "list has %llu k-mers, capacity %llu", (unsigned long long)n, (unsigned long long)capacity);
      } else if (n > 0) {
        const GenomeDesc &gd = one->h_genomes[0];
        st = alloc_buffer(ctx, (size_t)n * 32, &outbuf);
        unsigned long long *d_masked = nullptr, *d_bits = nullptr;
        if (st == SKS_OK) {
          d_masked = static_cast<unsigned long long *>(outbuf->ptr);
          d_bits = d_masked + 2 * n;
          st = launch_list_finalize(ctx, static_cast<const uint32_t *>(one->words->ptr) + gd.word_off,
                                    static_cast<const uint32_t *>(one->seg_end->ptr) + gd.seg_first, gd.n_segs, window,
                                    plan.n_limbs <= 2 ? 1 : 2, raw->ptr, static_cast<const uint32_t *>(pos->ptr),
                                    (uint32_t)n, d_masked, d_bits);
        }
        auto d2h = [&](uint64_t *dst, const unsigned long long *src) {
          if (!dst || st != SKS_OK) return;
          if (cudaMemcpyAsync(dst, src, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
            st = set_error(SKS_ERR_CUDA, "list download failed");
        };
        d2h(out_masked, d_masked);
        d2h(out_bits, d_bits);
        if (st == SKS_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = set_error(SKS_ERR_CUDA, "list download failed");
      }
    }
  }
  if (tmp) sks_batch_destroy(ctx, tmp);
  return st;
}

}  // extern "C"
