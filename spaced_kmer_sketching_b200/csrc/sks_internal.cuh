// Internal declarations shared by the CUDA translation units of libsks.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <string>
#include <vector>

#include "../../include/sks.h"

namespace sks {

// ------------------------------------------------------------------------------------------------
// Errors.  The C ABI never throws: internal code returns a status and records a thread-local text.
// ------------------------------------------------------------------------------------------------
int set_error(int code, const char *fmt, ...);
#define SKS_CUDA_TRY(expr)                                                                         \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::sks::set_error(SKS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                              __FILE__, __LINE__);                                                 \
  } while (0)
#define SKS_TRY(expr)              \
  do {                             \
    int _s = (expr);               \
    if (_s != SKS_OK) return _s;   \
  } while (0)

// ------------------------------------------------------------------------------------------------
// Geometry of the sketch kernel (see DESIGN.md "K1-K4").
// A genome is cut into tiles of kTileWindows window start positions; one CTA pass handles one tile.
// ------------------------------------------------------------------------------------------------
constexpr int kBasesPerWord = 16;    // 2-bit bases in a uint32 word
constexpr int kSketchThreads = 256;  // threads per CTA
constexpr int kGroup = 16;           // windows per unrolled group == bases per word
constexpr int kTileWindows = 8192;   // window starts per tile (all instantiations)
constexpr int kTileWords = kTileWindows / kBasesPerWord;  // 512
constexpr int kPreWords = 4;         // 16 B of history before the tile (only 1 word is ever read)
constexpr int kPostWords = 8;        // halo after the tile: <= 64+16 bases -> 5 words, rounded to 16 B
constexpr int kStageWords = kPreWords + kTileWords + kPostWords;  // 524 words = 2096 B per stage

// Per-genome descriptor in HBM.  Bases of a genome are packed 16 per word, all of its ACGT segments
// back to back; the words live at batch.words + word_off, preceded by >= kPreWords readable words
// and padded with zero words to a multiple of 4 (so every bulk copy is 16-byte granular).
struct GenomeDesc {
  uint64_t word_off;    // multiple of 4
  uint32_t n_bases;     // < 2^32 - 2^16
  uint32_t n_words;     // readable data words from word_off, multiple of 4
  uint32_t seg_first;   // first entry of this genome in the segment tables
  uint32_t n_segs;      // >= 1 (a genome without bases has one empty segment)
  uint32_t tile_first;  // id of the genome's first tile in the batch-wide tile numbering
  uint32_t n_tiles;
};

enum OutMode : int { OUT_KEYS = 0, OUT_BITSET = 1, OUT_LIST = 2, OUT_INDEX = 3, OUT_PART = 4 };
enum PredMode : int { PRED_ALL = 0, PRED_FMH181 = 1, PRED_FMH171 = 2 };

// Bucketed bitset build: the 4^weight-bit bitset is cut into slices of 2^kSliceBits bits (64 KB) that
// are assembled in shared memory and streamed out once (sks_sets.cu, bitset_build_kernel).
constexpr int kSliceBits = 19;
constexpr int kSliceWords = (1 << kSliceBits) / 32;  // 16384
constexpr int kGroupSliceBits = 3;                   // 8 slices (512 KB of bitset) per partition bucket
constexpr int kMaxPartBits = 32 - kSliceBits - kGroupSliceBits;  // <= 1024 buckets
constexpr int kMaxParts = 1 << kMaxPartBits;
// Partition geometry of the bucketed bitset build for a given index width.
struct PartGeometry {
  int part_shift;         // bucket = index >> part_shift
  uint32_t n_parts;       // buckets per genome
  uint32_t group_slices;  // slices per bucket
};
inline PartGeometry part_geometry(int index_bits) {
  // as many buckets as the sketch kernel's tables hold (kMaxParts): a bucket's indices have to fit shared memory
  // in the build kernels, and at 2^30 bits 256 buckets of 8 slices were twice too full for that
  const int group_bits = index_bits - kSliceBits - kMaxPartBits > 0 ? index_bits - kSliceBits - kMaxPartBits : 0;
  PartGeometry g;
  g.part_shift = kSliceBits + group_bits;
  g.n_parts = 1u << (index_bits - g.part_shift);
  g.group_slices = 1u << group_bits;
  return g;
}

// PEXT(masked_bits, mask) as rotate-and-mask pieces: one piece per run of mask ones inside a 32-bit limb
// (a limb holds 16 base positions, hence at most 8 runs).  The compacted index takes the bits selected
// by (rotr(limb, rot) & dmask); unused slots have dmask == 0.  The table is indexed with compile-time
// constants only, so every entry is a direct constant-bank operand (a dynamically indexed table went
// through LDC and kept the ADU pipe 80 % busy).
constexpr int kPiecesPerLimb = 8;
struct PextTable {
  uint32_t dmask[4][kPiecesPerLimb];
  uint32_t rot[4][kPiecesPerLimb];
  uint32_t n_pieces[4];
};

// Bucket index of the bucket sort: the top `bb` MASK-SELECTED bits of the key (up to six runs of the mask,
// concatenated from the top, so the index is monotone in the key).  A plain bit field below the mask's highest bit
// would include the positions a spaced seed skips, which are zero in every key: a quarter of the buckets would
// get all the keys.
struct SortPlan {
  int n_pieces;
  int word[6];       // 0: low 64 bits of the key, 1: high 64 bits
  int s[6];          // piece i = ((word >> s[i]) & m[i]) << o[i]
  uint32_t m[6];
  int o[6];
};
// Plan for `bb` bucket bits under `mask`; false when the mask has fewer set bits than that or needs more than six
// pieces.
// `skip`: that many of the mask's top set bits are passed over first (they are the same in every key of the region:
// keys routed to the owner of a key range).
inline bool sort_plan(const uint64_t mask[2], int bb, SortPlan *out, int skip = 0) {
  SortPlan p = {};
  int got = 0, bit = 127;
  auto set = [&](int b) { return b >= 0 && ((mask[b >> 6] >> (b & 63)) & 1); };
  for (int s = 0; s < skip; ++s) {
    while (bit >= 0 && !set(bit)) --bit;
    if (bit < 0) return false;
    --bit;
  }
  while (got < bb) {
    while (bit >= 0 && !set(bit)) --bit;
    if (bit < 0 || p.n_pieces == 6) return false;
    const int hi = bit;
    // a piece stays inside one 64-bit word and takes at most what is still missing
    while (bit >= 0 && set(bit) && (bit >> 6) == (hi >> 6) && hi - bit + 1 <= bb - got) --bit;
    const int len = hi - bit;
    p.word[p.n_pieces] = hi >> 6;
    p.s[p.n_pieces] = (bit + 1) & 63;
    p.m[p.n_pieces] = (len >= 32) ? 0xFFFFFFFFu : ((1u << len) - 1);
    got += len;
    p.o[p.n_pieces] = bb - got;
    ++p.n_pieces;
  }
  *out = p;
  return true;
}


struct SketchParams {
  const uint32_t *words;       // batch buffer (nullptr when host_words: GenomeDesc::word_off is then absolute)
  const GenomeDesc *genomes;   // [n_genomes]
  const uint32_t *seg_end;     // exclusive end of every segment, relative to its genome start
  int n_genomes;
  uint32_t n_tiles;            // over the whole batch
  uint32_t tile_begin;         // the launch works on tiles [tile_begin, n_tiles): 0 but for a batch that is still arriving
  int window;                  // w, 1..64
  uint32_t mask[4];            // 128-bit mask as four 32-bit limbs
  // predicate (FMH), modulus = 2^s * d with d odd: pass <=> t = (H(masked) ^ hconst) * minv has (t & mlow) == 0
  // and t <= mbound
  uint64_t hconst;             // H(mask) ^ window ^ (int64)nonce
  uint64_t minv;               // inverse of d mod 2^64
  uint64_t mbound;             // floor((2^64-1) / modulus) << s
  uint64_t mlow;               // 2^s - 1
  // OUT_KEYS / OUT_LIST: per-genome output regions
  void *out_keys;                       // uint64 (NL<=2) / ulonglong2 (NL>2) slots; uint32 for OUT_INDEX
  uint32_t *out_pos;                    // OUT_LIST only: (strand<<31 | start position) per slot
  const uint64_t *out_off;              // [n_genomes] first slot of the genome's region
  const uint64_t *out_cap;              // [n_genomes] slots in the region
  unsigned long long *out_count;        // [n_genomes] kept k-mers (keeps counting past out_cap)
  // OUT_PART: PEXT indices scattered straight into per-(genome, bucket) regions of part_cap slots each
  // (out_keys = region buffer); part_cursor[genome * n_parts + bucket] = slots used in the region, starts at 0
  uint32_t *part_cursor;
  uint32_t *part_overflow;     // set to 1 when a region was too small (the caller then takes the exact path)
  uint32_t n_parts;            // <= kMaxParts
  uint32_t part_cap;
  int part_shift;              // bucket = index >> part_shift
  uint32_t host_words;         // genomes are read in place from pinned host memory (see the kernel's tile fetch)
  // OUT_KEYS under a sparse predicate, 8-byte keys: the first level of the bucket sort folded into the emit.  The kept
  // k-mers of genome g go straight to the region of (g, bucket of the key) -- kpart_cap slots at
  // ((g << kpart_bits) + bucket) * kpart_cap of out_keys, positions handed out by kpart_cursor (zero on entry); a full
  // region raises kpart_overflow and the caller falls back to emit + histogram + scatter.  kpart_bits == 0: off.
  uint32_t kpart_bits, kpart_cap;
  uint32_t *kpart_cursor, *kpart_overflow;
  SortPlan kpart_plan;
  // OUT_BITSET
  uint32_t *bitset;            // n_genomes consecutive bitsets
  uint64_t bitset_words;       // words per genome
  PextTable pext;
};

// ------------------------------------------------------------------------------------------------
// Host-side objects behind the opaque C handles.
// ------------------------------------------------------------------------------------------------
struct DeviceBuffer {  // ref-counted cudaMalloc block shared by the sets of one sketch call
  void *ptr = nullptr;
  size_t bytes = 0;
  int device = 0;
  cudaStream_t stream = nullptr;  // stream the block was allocated on; it is freed in that stream's order
  ~DeviceBuffer();
};
using BufferRef = std::shared_ptr<DeviceBuffer>;
int alloc_buffer(sks_ctx *ctx, size_t bytes, BufferRef *out);
void cache_release(int device, cudaStream_t stream, bool stream_only);

}  // namespace sks

struct sks_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = true;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  int64_t in_place_calls = 0;  // sks_pair_ani calls that read the genomes from pinned host memory
  int kpart_skip = 0;          // sketch_sorted: calls that still skip the folded sort partition after it overflowed
  int64_t streamed_calls = 0;  // sks_all_vs_all_from_host calls whose genomes arrived chunk by chunk under the sketch kernel
  cudaStream_t copy_stream = nullptr;    // host-to-device copies of a batch that is sketched while it arrives
  std::vector<cudaEvent_t> sync_events;  // one per chunk of such a batch (no timing), reused by the next call
  void *stage_ring = nullptr;            // three pinned staging buffers for pageable sources of such a batch
  size_t stage_ring_bytes = 0;
  // reusable scratch (grown on demand)
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  bool exact_partition = false;  // SKS_EXACT_PARTITION=1: always take the counting partition of the bucketed build
  // bitsets of >= 2^this bits are built by the bucketed path (SKS_BUCKET_MIN_BITS).  Below, the bitsets of a pair
  // fit the L2 (2 x 2^28 bits = 64 MB) and direct atomicOr is faster: 0.11-0.12 ms against 3.2 / 0.9 ms for the
  // slice assembly at 2^26 / 2^28 bits, whose buckets are far too full to be staged in shared memory.
  int bucket_min_bits = 30;
  // pinned staging for small D2H/H2D traffic
  void *pinned = nullptr;
  size_t pinned_bytes = 0;
  size_t pinned_off = 0;
  // optional per-kernel event timing (sks_ctx_profile)
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof[SKS_KERNEL_KINDS];
  std::vector<cudaEvent_t> event_pool;
};

struct sks_batch {
  int device = 0;
  int n_genomes = 0;
  sks::BufferRef words;      // uint32 words of all genomes (empty for a host-resident batch)
  bool host_words = false;   // internal to sks_pair_ani: the genomes stay in the caller's pinned host buffers
  sks::BufferRef genomes;    // GenomeDesc[n_genomes]
  sks::BufferRef seg_end;    // uint32[total_segs]
  sks::BufferRef tile_genome;  // uint32[n_tiles] (only when n_genomes > 1)
  std::vector<sks::GenomeDesc> h_genomes;
  std::vector<uint32_t> h_seg_end;
  uint32_t n_tiles = 0;
  uint64_t total_bases = 0;
  // sks_all_vs_all_from_host: the words of tiles [tile_begin, tile_end) are there once `ready` has happened (the copies run
  // on the context's copy stream, in this order); empty for every other batch
  struct Arriving {
    uint32_t tile_begin, tile_end;
    cudaEvent_t ready;
  };
  std::vector<Arriving> arriving;
  // pageable sources: a feeder thread stages and queues the copies; chunk c may be waited for once n_queued > c
  struct Feed {
    std::mutex m;
    std::condition_variable cv;
    size_t n_queued = 0;
    int status = 0;
    std::thread worker;
    ~Feed() {
      if (worker.joinable()) worker.join();
    }
  };
  std::unique_ptr<Feed> feed;
};

struct sks_set {
  int device = 0;
  int repr = SKS_REPR_SORTED;
  int window = 0;
  int weight = 0;
  uint64_t mask[2] = {0, 0};
  // SORTED: `count` ascending distinct keys of `key_words` uint64 each at buf + byte_off
  // BITSET: 4^weight bits at buf + byte_off
  sks::BufferRef buf;
  size_t byte_off = 0;
  int key_words = 1;
  int64_t count = -1;  // -1: not yet known (BITSET before the first popcount)
  uint64_t bitset_words = 0;
  // BITSET built by the bucketed path: the build kernel left the popcount on the device
  sks::BufferRef count_buf;
  size_t count_off = 0;
};

namespace sks {
// Makes `dev` the current device for the scope.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};
sks_set *new_set(const sks_ctx *ctx, int repr, const uint64_t mask[2], int window, int weight);
// SKS_ERR_MISMATCH unless the two sets share mask, representation, key width and device.
int check_pair(const sks_set *a, const sks_set *b);

// Brackets the kernel launches of one scope with events when the context is being profiled.
struct KernelTimer {
  sks_ctx *ctx;
  int kind;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  KernelTimer(sks_ctx *c, int k);
  ~KernelTimer();
};

// scratch / pinned helpers
int ctx_scratch(sks_ctx *ctx, size_t bytes, void **out);
int ctx_pinned(sks_ctx *ctx, size_t bytes, void **out);

// kernels' host launchers (sks_sketch.cu, sks_sets.cu, sks_synth.cu)
int launch_sketch(sks_ctx *ctx, const SketchParams &p, const uint32_t *tile_genome, int n_limbs, int pred_mode,
                  int out_mode);
int launch_fill_zero(sks_ctx *ctx, void *ptr, size_t bytes);
int launch_synth(sks_ctx *ctx, uint32_t *words, const GenomeDesc *genomes, int n_genomes, uint32_t max_words,
                 const uint64_t *gen_seed, const uint64_t *mut_seed, const uint64_t *mut_D, const uint64_t *first_base);
int launch_bitset_pair_counts(sks_ctx *ctx, const uint32_t *a, const uint32_t *b, uint64_t n_words,
                              unsigned long long *out3);
int launch_bitset_build(sks_ctx *ctx, const uint32_t *raw_idx, uint32_t *sorted_idx, const uint64_t *h_off,
                        const uint64_t *h_count, int n_genomes, int index_bits, uint32_t *bitset, uint64_t bitset_words,
                        unsigned long long *d_set_count);
int launch_bitset_assemble(sks_ctx *ctx, const uint32_t *regions, const uint32_t *d_cursor, uint32_t part_cap, int n_genomes,
                           int index_bits, uint32_t *bitset, uint64_t bitset_words, unsigned long long *d_set_count,
                           unsigned int *d_work_counter);
int launch_bitset_pair_build(sks_ctx *ctx, const uint32_t *regions, const uint32_t *d_cursor, uint32_t part_cap, int index_bits,
                             uint32_t *bitset_a, uint32_t *bitset_b, unsigned long long *d_out3, unsigned int *d_work_counter);
int launch_bitset_popcount(sks_ctx *ctx, const uint32_t *a, uint64_t n_words, unsigned long long *out1);
int sort_unique_regions(sks_ctx *ctx, int key_words, void *keys, const uint64_t *h_off, const uint64_t *h_count,
                        int n_regions, uint64_t span, BufferRef *out_buf, std::vector<uint64_t> *out_off,
                        std::vector<uint64_t> *out_count, const uint64_t *mask = nullptr,  // mask: enables the bucket sort
                        int skip_bits = 0);  // the mask's top skip_bits set bits are the same in every key of a region
bool bucket_regions_plan(const uint64_t mask[2], int n_genomes, uint64_t max_count, int *bb_out, uint32_t *cap_out, SortPlan *plan);
int sort_unique_from_buckets(sks_ctx *ctx, unsigned long long *regions, int n_genomes, int bb, uint32_t cap, const uint32_t *d_cursor,
                             uint32_t *d_flag, uint64_t total_bound, BufferRef *out_buf, std::vector<uint64_t> *out_off,
                             std::vector<uint64_t> *out_count, bool *handled);
int launch_sorted_intersect_pairs(sks_ctx *ctx, int key_words, const void *const *d_a, const int64_t *d_na,
                                  const void *const *d_b, const int64_t *d_nb, int64_t n_pairs, int32_t *d_out,
                                  const uint32_t *d_pair_idx = nullptr, int slices = 1);
int validate_keys(sks_ctx *ctx, const void *d_keys, int64_t n, int key_words, const uint64_t mask[2], bool need_sorted,
                  const int64_t *h_starts, int n_starts);
// Row-resident variant: tasks (RowTask, 24 bytes: a, n_a, first, n_cols, pad) over the same pair tables.
struct RowTaskHost {
  const void *a;
  uint32_t n_a, first, n_cols, pad;
};
bool row_intersect_fits(int key_words, const uint64_t mask[2], int64_t n_a);
int launch_row_intersect(sks_ctx *ctx, int key_words, const void *d_tasks, int64_t n_tasks, const void *const *d_b,
                         const int64_t *d_nb, int32_t *d_out, const uint64_t mask[2], int64_t max_n_a);

int sketch_raw_one(sks_ctx *ctx, const sks_batch *batch, const uint64_t mask[2], int window, const sks_pred *pred,
                   BufferRef *keys, uint64_t *count, int *key_words);
// all-vs-all through the dictionary of shared k-mers (sks_allpairs.cu)
bool all_pairs_dict_eligible(sks_set *const *sets, int64_t n);
int all_pairs_dict(sks_ctx *ctx, sks_set *const *sets, int64_t n, int64_t row_begin, int64_t row_end, BufferRef *counts,
                   BufferRef *ani, BufferRef *sizes_out, const uint32_t **h_overflow);
struct FlatKeys {  // occurrences as arrays instead of sets: 64-bit keys, the global number of each one's set
  const unsigned long long *keys;
  const uint16_t *sets;
  uint32_t n;
  const int32_t *h_sizes;  // [n sets] set sizes (host)
};
int all_pairs_raw(sks_ctx *ctx, sks_set *const *sets, int64_t n, int64_t row_begin, int64_t row_end, int part, int n_parts,
                  bool symmetric, int64_t raw_rows, BufferRef *raw_out, BufferRef *sizes_out, const uint32_t **h_overflow,
                  const FlatKeys *flat);
size_t all_pairs_route_cap(uint64_t n_keys, int world);
int all_pairs_route(sks_ctx *ctx, sks_set *const *sets, int64_t n_local, int64_t set_base, int world, size_t region_cap,
                    BufferRef *out_keys, BufferRef *out_sets, BufferRef *ctl_out, unsigned long long **d_counts);
bool all_pairs_dict_usable(int key_words, const uint64_t mask[2], int64_t n_total, uint64_t total_keys);
int all_pairs_finalize(sks_ctx *ctx, const int32_t *raw_rows, const int32_t *d_sizes, int64_t n, int64_t row_begin,
                       int64_t n_rows, bool symmetric, int weight, BufferRef *counts, BufferRef *ani);

int launch_list_finalize(sks_ctx *ctx, const uint32_t *words, const uint32_t *seg_end, uint32_t n_segs, int window,
                         int key_words, const void *raw_keys, const uint32_t *raw_pos, uint32_t n,
                         unsigned long long *out_masked, unsigned long long *out_bits);

int fasta_parse_device(sks_ctx *ctx, const unsigned char *d_text, const std::vector<uint64_t> &h_file_off,
                       std::vector<uint64_t> *n_bases, std::vector<std::vector<uint64_t>> *seg_len, BufferRef *codes_buf,
                       BufferRef *genome_base_buf);
int launch_pack_codes(sks_ctx *ctx, const uint8_t *d_codes, const uint32_t *d_genome_base, const GenomeDesc *d_genomes,
                      int n_genomes, uint32_t max_words, uint32_t *d_words);

// host-only helpers (sks_host.cpp)
uint64_t boost_hash_bitset(uint64_t lo, uint64_t hi, int variant);
void modulus_magic(uint64_t modulus, uint64_t *minv, uint64_t *mbound, int *mshift);
int build_pext_table(const uint64_t mask[2], int n_limbs, PextTable *out, int *n_index_bits);
}  // namespace sks
