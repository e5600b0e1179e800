/*
 * sks.h -- C ABI of the B200-native spaced k-mer sketch + ANI engine (libsks.so).
 *
 * This is the drop-in boundary for the reference's sketch-and-compare path.  The reference
 * (bensonlzl/spaced-kmer-sketching) has no FFI of its own: its boundary is the C++ free-function
 * API of src/kmer.hpp / src/fasta_processing.hpp / src/ani_estimator.hpp.  The C++ headers next to
 * this file (include/kmer.hpp, ...) re-export that API with the reference's signatures and are
 * implemented on top of the entry points below; every entry point names the reference interface it
 * replaces (file:line under /root/reference).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types.
 *   - every function that can fail returns an int status (SKS_OK == 0); sks_last_error() gives a
 *     thread-local message.  Nothing here throws or exits.
 *   - a 128-bit kmer_bitset value (src/kmer.hpp:27,37) crosses the ABI as uint64_t[2] = {bits 0..63,
 *     bits 64..127}, the block order of boost::dynamic_bitset<unsigned long>.
 *   - there is NO CPU fallback: compute entry points fail with SKS_ERR_CUDA when no device is usable.
 */
#ifndef SKS_H
#define SKS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKS_VERSION 100

/* status codes */
#define SKS_OK 0
#define SKS_ERR_INVALID 1   /* bad argument (the message says which)                         */
#define SKS_ERR_CUDA 2      /* CUDA runtime / launch failure, or no device                   */
#define SKS_ERR_CAPACITY 3  /* an output buffer was too small                                */
#define SKS_ERR_MISMATCH 4  /* list lengths / set kinds differ (src/kmer_set.cpp:147-150)    */
#define SKS_ERR_IO 5        /* unreadable file (src/fasta_processing.cpp:86-90)              */

/* Sketching predicate: the recognised forms of `std::function<bool(const kmer)>`
 * (src/kmer.hpp:93-103,195-212).  SKS_PRED_FMH is the driver's condition
 * `frac_min_hash(nonce)(k) % modulus == 0` (src/kmer-sketching.cpp:29-34, src/kmer.hpp:135-149). */
#define SKS_PRED_ALL 0
#define SKS_PRED_FMH 1

/* boost::hash_combine flavour behind frac_min_hash (the reference pins no Boost version). */
#define SKS_HASH_BOOST_171 171 /* Boost 1.71 .. 1.80 */
#define SKS_HASH_BOOST_181 181 /* Boost >= 1.81 (default) */

/* Device representation of a kmer_set (src/kmer.hpp:160-190). */
#define SKS_REPR_AUTO 0   /* BITSET for predicate ALL with weight <= 16 when the bitset is at most ~128 B per window of
                             the largest genome and the batch's bitsets fit half the free memory; SORTED otherwise */
#define SKS_REPR_SORTED 1 /* ascending distinct masked_bits, 8 B (window <= 32) or 16 B per key  */
#define SKS_REPR_BITSET 2 /* 4^weight-bit presence bitset indexed by PEXT(masked_bits, mask)     */
/* sks_pair_ani* only: the two presence bitsets are assembled and compared slice by slice in shared memory
 * and are never written to HBM (sks_sketch rejects it: there would be no set to return). */
#define SKS_REPR_BITSET_ONCHIP 3

typedef struct sks_pred {
  int32_t kind;         /* SKS_PRED_*                                   */
  int32_t nonce;        /* frac_min_hash(int n), src/kmer.hpp:141        */
  uint64_t modulus;     /* c in `fmh(k) % c == 0`; must be non-zero      */
  int32_t hash_variant; /* SKS_HASH_BOOST_*; 0 selects the default (181) */
  int32_t reserved;
} sks_pred;

typedef struct sks_ctx sks_ctx;     /* one per (thread, device): stream, scratch, timers */
typedef struct sks_batch sks_batch; /* genomes resident in HBM, 2-bit packed + segment tables */
typedef struct sks_set sks_set;     /* one device-resident kmer_set */

/* ---- library / context --------------------------------------------------------------------- */
int sks_version(void);
const char *sks_last_error(void);
int sks_device_count(void);
/* Creates a context on CUDA device `device` with its own non-blocking stream. */
int sks_ctx_create(int device, sks_ctx **out);
void sks_ctx_destroy(sks_ctx *ctx);
/* Use an externally owned cudaStream_t (e.g. torch's current stream) for all later work. */
int sks_ctx_set_stream(sks_ctx *ctx, void *cuda_stream);
int sks_ctx_sync(sks_ctx *ctx);
/* CUDA-event timing on the context's stream: begin/end bracket a region, end returns ms. */
int sks_timer_begin(sks_ctx *ctx);
int sks_timer_end(sks_ctx *ctx, float *out_ms);
/* Number of kernels launched by this context so far (bench.py's gpu_launches). */
int64_t sks_ctx_launch_count(const sks_ctx *ctx);
/* Number of sks_pair_ani calls of this context that read the genomes in place from the caller's pinned host
 * buffers (no host-to-device copy; see sks_pair_ani). */
int64_t sks_ctx_in_place_count(const sks_ctx *ctx);
/* Number of sks_all_vs_all_from_host calls of this context whose genomes were copied from the caller's pinned host
 * buffers chunk by chunk while the sketch kernel already worked on the chunks that had arrived. */
int64_t sks_ctx_streamed_count(const sks_ctx *ctx);

/* Per-kernel CUDA-event timing (bench.py's roofline numbers).  While enabled, every kernel launch of
 * the context is bracketed by a pair of events on the context's stream; sks_ctx_kernel_stats
 * synchronises, returns launches and summed duration of one kernel kind since the last query and
 * resets it. */
#define SKS_KERNEL_SKETCH 0       /* fused unpack / gather / hash-filter / emit (K1-K4)  */
#define SKS_KERNEL_FILL 1         /* bitset clear                                        */
#define SKS_KERNEL_PAIR_COUNTS 2  /* bitset AND/popcount (K5, bitset)                    */
#define SKS_KERNEL_POPCOUNT 3     /* bitset popcount (set size)                          */
#define SKS_KERNEL_SORT_UNIQUE 4  /* radix sort + unique of sketched keys                */
#define SKS_KERNEL_INTERSECT 5    /* sorted-set intersection (K5, sorted)                */
#define SKS_KERNEL_SYNTH 6        /* synthetic genome generator                          */
#define SKS_KERNEL_LIST 7         /* ordered-list finalisation                           */
#define SKS_KERNEL_BITSET_BUILD 8 /* bucket sort + slice-wise bitset assembly (K4, bucketed) */
#define SKS_KERNEL_FASTA 9        /* device-side FASTA parse + 2-bit pack                */
#define SKS_KERNEL_PAIR_BUILD 10  /* fused slice assembly + AND/popcount of a genome pair (K4b + K5) */
#define SKS_KERNEL_DICT 11       /* all-vs-all: dictionary of shared k-mers, sets re-coded as id bitmaps / lists */
#define SKS_KERNEL_ALLPAIRS 12   /* all-vs-all: AND/popcount of the re-coded sets (K5, many sets)               */
#define SKS_KERNEL_ANI 13        /* all-vs-all: mirror + diagonal + containment^(1/weight) on the device        */
#define SKS_KERNEL_EXCHANGE 14   /* several GPUs: NCCL exchange of the sketches (header all-gather + grouped send/recv) */
#define SKS_KERNEL_KINDS 15
int sks_ctx_profile(sks_ctx *ctx, int enable);
int sks_ctx_kernel_stats(sks_ctx *ctx, int kind, int64_t *out_launches, double *out_total_ms);
const char *sks_kernel_name(int kind);

/* ---- host-side helpers: masks, packing, FASTA, ANI (no device needed) ------------------------ */
/* Seed string in README notation ("11001011", README.md:25-41) -> 128-bit mask with two bits per
 * used position; s[i]=='1' sets bits 2(w-1-i), 2(w-1-i)+1.  The reference has no parser. */
int sks_seed_to_mask(const char *seed, uint64_t out_mask[2], int *out_window);
/* mask.count() / NUCLEOTIDE_BIT_SIZE, src/kmer-sketching.cpp:164 */
int sks_mask_weight(const uint64_t mask[2]);
/* contiguous_kmer(k), src/kmer_bitset.cpp:51-56 (k > 64 -> SKS_ERR_INVALID) */
int sks_contiguous_mask(int k, uint64_t out_mask[2]);
/* generate_random_spaced_seed_mask, src/kmer_bitset.cpp:132-152 (libstdc++ shuffle + mt19937) */
int sks_random_mask(int window, int k, uint64_t seed, uint64_t out_mask[2]);
/* reverse_kmer_bitset, src/kmer_bitset.cpp:105-119 */
void sks_reverse_bitset(const uint64_t in[2], uint64_t out[2]);
/* boost::hash<boost::dynamic_bitset<>> of a 128-bit value (call sites src/kmer.hpp:137,146). */
uint64_t sks_boost_hash_bitset(const uint64_t value[2], int hash_variant);
/* frac_min_hash::operator(), src/kmer.hpp:144-148, on the host (for the C++ functor type). */
uint64_t sks_fmh_hash(const uint64_t masked[2], const uint64_t mask[2], int window, int nonce,
                      int hash_variant);
/* containment / binomial_estimator, src/ani_estimation.cpp:24-42 (host double arithmetic) */
double sks_containment(int intersection, int set_size);
double sks_binomial_estimator(double containment, int kmer_num_ones);

/* 2-bit packing: 16 bases per uint32 word, base i in bits 2(i%16)..2(i%16)+1 of word i/16.
 * `codes` are the reference's 1-byte codes 0..3 (acgt_string, src/fasta_processing.hpp:18). */
size_t sks_packed_words(uint64_t n_bases);
int sks_pack_codes(const uint8_t *codes, uint64_t n_bases, uint32_t *out_words);
int sks_unpack_codes(const uint32_t *words, uint64_t n_bases, uint8_t *out_codes);
/* FASTA text -> packed bases + segment table with the reference's record and split rules
 * (strings_from_fasta + cut_nucleotide_strings, src/fasta_processing.cpp:79-211).  Two-call
 * protocol: with out_* NULL it only reports *n_bases / *n_segs.  Segments are the maximal ACGT
 * runs in file order; their bases are packed back to back. */
int sks_fasta_parse(const char *text, size_t n, uint64_t *n_bases, uint64_t *n_segs,
                    uint32_t *out_words, uint64_t *out_seg_len);
/* Same, reading the file; an unreadable file returns SKS_ERR_IO (the C++ wrapper turns that into
 * the reference's stderr message + exit(1)). */
int sks_fasta_parse_file(const char *path, uint64_t *n_bases, uint64_t *n_segs, uint32_t **out_words,
                         uint64_t **out_seg_len); /* malloc'ed; release with sks_free */
void sks_free(void *p);

/* ---- batches: genomes in HBM ---------------------------------------------------------------- */
/* Uploads n_genomes packed genomes (HOST buffers).  packed[g] holds sks_packed_words(n_bases[g])
 * words; seg_len[g][0..n_segs[g]) are the ACGT-run lengths (NULL / 0 => one segment of n_bases).
 * Replaces the host side of kmer_set_from_fasta_file up to the sliding loop
 * (src/kmer_set.cpp:54-68). */
int sks_batch_upload(sks_ctx *ctx, int n_genomes, const uint32_t *const *packed, const uint64_t *n_bases,
                     const uint64_t *const *seg_len, const uint64_t *n_segs, sks_batch **out);
/* FASTA ingest ON THE DEVICE: the raw bytes of n_files FASTA files (HOST buffers) are uploaded and parsed,
 * split at non-ACGT bytes and 2-bit packed by kernels, with the reference's record rules
 * (strings_from_fasta + cut_nucleotide_strings, src/fasta_processing.cpp:79-211).  Same result as
 * sks_fasta_parse + sks_batch_upload; at most 2 GiB of text per call. */
int sks_batch_from_fasta_text(sks_ctx *ctx, int n_files, const char *const *text, const uint64_t *n_bytes, sks_batch **out);
/* Same, reading the files; an unreadable file returns SKS_ERR_IO. */
int sks_batch_from_fasta_files(sks_ctx *ctx, int n_files, const char *const *paths, sks_batch **out);
/* Segment (ACGT-run) lengths of one genome of a batch; out_seg_len NULL => only *n_segs. */
int sks_batch_segments(const sks_batch *b, int genome, uint64_t *n_segs, uint64_t *out_seg_len);
/* Synthetic genomes generated ON DEVICE (benchmark inputs, SURVEY.md 4.2 KAT-3 generator):
 * genome g = mutate(gen(n_bases, gen_seed[g]), mut_seed[g], mut_D[g]); mut_D[g]==0 => no mutation. */
int sks_batch_synth(sks_ctx *ctx, int n_genomes, uint64_t n_bases, const uint64_t *gen_seed,
                    const uint64_t *mut_seed, const uint64_t *mut_D, sks_batch **out);
/* Same, but genome g holds bases [first_base[g], first_base[g] + n_bases) of its (unbounded) stream:
 * a rank's slice of one long synthetic sequence without generating the rest. */
int sks_batch_synth_at(sks_ctx *ctx, int n_genomes, uint64_t n_bases, const uint64_t *first_base,
                       const uint64_t *gen_seed, const uint64_t *mut_seed, const uint64_t *mut_D, sks_batch **out);
/* A window [first_base, first_base + n_starts + window - 1) of one genome of `src` as a new
 * single-genome batch (position sharding of one long sequence with a (w-1)-base halo). */
int sks_batch_slice(sks_ctx *ctx, const sks_batch *src, int genome, uint64_t first_base, uint64_t n_starts,
                    int window, sks_batch **out);
int sks_batch_n_genomes(const sks_batch *b);
uint64_t sks_batch_n_bases(const sks_batch *b, int genome);
/* D2H copy of one genome's packed words (tests). */
int sks_batch_download(sks_ctx *ctx, const sks_batch *b, int genome, uint32_t *out_words);
void sks_batch_destroy(sks_ctx *ctx, sks_batch *b);

/* ---- sketching: the hot path ----------------------------------------------------------------- */
/* For every genome of the batch: slide the spaced seed, canonicalise, filter, and build the set.
 * Replaces nucleotide_string_list_to_kmers + kmer_set::insert_kmers
 * (src/kmer_sliding.cpp:112-238, src/kmer.hpp:170-178) == the body of kmer_set_from_fasta_file
 * (src/kmer_set.cpp:54-68) and its loop/cilk_for over files (:81-133).
 * out_sets[0..n_genomes) receive new sets (release each with sks_set_destroy). */
int sks_sketch(sks_ctx *ctx, const sks_batch *batch, const uint64_t mask[2], int window, const sks_pred *pred,
               int repr, sks_set **out_sets);
/* The ordered, duplicate-preserving list of nucleotide_string_list_to_kmers
 * (src/kmer_sliding.cpp:224-238) for one genome: masked_bits and kmer_bits (2 words each) in sequence
 * order.  Two-call protocol: out_* NULL => only *out_n is written. */
int sks_kmer_list(sks_ctx *ctx, const sks_batch *batch, int genome, const uint64_t mask[2], int window,
                  const sks_pred *pred, uint64_t *out_n, uint64_t *out_masked, uint64_t *out_bits,
                  uint64_t capacity);

/* ---- sets ------------------------------------------------------------------------------------ */
int sks_set_repr(const sks_set *s);
int sks_set_window(const sks_set *s);
int sks_set_weight(const sks_set *s);
/* kmer_set::kmer_set_size, src/kmer.hpp:186-189 */
int sks_set_size(sks_ctx *ctx, sks_set *s, int64_t *out);
/* Members as ascending 128-bit masked_bits (2 words per key), D2H.  For BITSET sets the indices are
 * expanded back through the mask (PDEP).  capacity in keys. */
int sks_set_keys(sks_ctx *ctx, sks_set *s, uint64_t *out_lohi, uint64_t capacity);
/* Raw device view of a SORTED set (for collectives): pointer, key count, words (uint64) per key. */
int sks_set_device_keys(sks_ctx *ctx, sks_set *s, const void **dptr, int64_t *n_keys, int *words_per_key);
/* Imports.  Stream contract: the keys are copied on the context's stream and validated (subsets of the mask; where
 * sortedness is claimed, strictly ascending inside every set) before the call returns, which synchronises the stream:
 * the caller's buffer is free on return.  A caller that produced the keys on another stream must order that stream
 * before the call (event or sync).  words_per_key must be 1 for window <= 32 and 2 above. */
/* Builds a SORTED set from ascending distinct keys already on the device (copied). */
int sks_set_from_device_keys(sks_ctx *ctx, const void *dptr, int64_t n_keys, int words_per_key,
                             const uint64_t mask[2], int window, sks_set **out);
/* n_sets SORTED sets whose keys lie back to back at dptr (set i has counts[i] keys): one device copy, the
 * sets share the buffer.  The receive side of an all-gather of sketches. */
int sks_sets_from_device_keys(sks_ctx *ctx, const void *dptr, int64_t n_sets, const int64_t *counts, int words_per_key,
                              const uint64_t mask[2], int window, sks_set **out);
/* Builds a SORTED set from UNSORTED, possibly duplicated device keys (sort + unique on device);
 * the merge step after an all-gather of per-rank partial sketches of one sequence. */
int sks_set_from_unsorted_device_keys(sks_ctx *ctx, const void *dptr, int64_t n_keys, int words_per_key,
                                      const uint64_t mask[2], int window, sks_set **out);
/* Same from HOST keys (2 words per key: lo, hi), e.g. kmer_set::insert_kmers (src/kmer.hpp:170-178) on a
 * host-built list, or the survivors of a host-evaluated sketching condition. */
int sks_set_from_host_keys(sks_ctx *ctx, const uint64_t *keys_lohi, int64_t n_keys, const uint64_t mask[2], int window,
                           sks_set **out);
/* Device the set lives on, and a copy of it on the device of `dst` (peer-to-peer over NVLink where available): what a
 * single-process, several-GPU caller needs to compare sets that were sketched on different devices. */
int sks_set_device_index(const sks_set *s);
int sks_set_clone_to(sks_ctx *dst, sks_set *src, sks_set **out);
void sks_set_destroy(sks_ctx *ctx, sks_set *s);

/* ---- sketch files (not in the reference: it never persists a sketch) ---------- */
/* Little-endian file: "SKSKETCH", u32 version (1), u32 window, u64 mask[2], u32 pred kind, i32 nonce,
 * u64 modulus, u32 hash variant, u32 words per key (1 or 2), u64 n_keys, then the ascending distinct
 * masked_bits.  A BITSET set is written as its keys.  `pred` may be NULL (recorded as kind 0xFFFFFFFF). */
int sks_set_save(sks_ctx *ctx, sks_set *s, const sks_pred *pred, const char *path);
/* Loads a file written by sks_set_save as a SORTED set; out_pred (may be NULL) receives the recorded condition. */
int sks_set_load(sks_ctx *ctx, const char *path, sks_set **out, sks_pred *out_pred);

/* ---- comparison ------------------------------------------------------------------------------ */
/* kmer_set_intersection, src/kmer_set.cpp:23-41 */
int sks_intersect(sks_ctx *ctx, sks_set *a, sks_set *b, int64_t *out);
/* (parallel_)compute_pairwise_kmer_set_intersections, src/kmer_set.cpp:143-184: out[i] =
 * |a[i] n b[i]|; na != nb returns SKS_ERR_MISMATCH (the reference throws std::runtime_error). */
int sks_intersect_pairs(sks_ctx *ctx, sks_set *const *a, int64_t na, sks_set *const *b, int64_t nb, int32_t *out);
/* All n*n ordered pairs in generate_all_pairs_from_vector order (src/generators.hpp:44-58):
 * out[i*n + j] = |sets[i] n sets[j]|.  Rows [row_begin, row_end) only (rank tiling); other
 * entries are left untouched. */
int sks_intersect_all_pairs(sks_ctx *ctx, sks_set *const *sets, int64_t n, int64_t row_begin, int64_t row_end,
                            int32_t *out);
/* Same for the block rows [row_begin, row_end) x columns [col_begin, col_end) of that matrix: the unit of the
 * multi-GPU tiling (every unordered block pair is evaluated by exactly one rank and mirrored after the gather,
 * since |A n B| = |B n A|). */
int sks_intersect_block(sks_ctx *ctx, sks_set *const *sets, int64_t n, int64_t row_begin, int64_t row_end,
                        int64_t col_begin, int64_t col_end, int32_t *out);
/* Several such rectangles (rects[4*q .. 4*q+3] = row_begin, row_end, col_begin, col_end) in one pass over one pair
 * table: what a rank of the multi-GPU tiling evaluates. */
int sks_intersect_rects(sks_ctx *ctx, sks_set *const *sets, int64_t n, const int64_t *rects, int64_t n_rects, int32_t *out);
/* The comparison phase of the reference driver in one device-resident pass (src/kmer-sketching.cpp:185-200:
 * generate_all_pairs_from_vector -> parallel_compute_pairwise_kmer_set_intersections -> containment ->
 * binomial_estimator), for the rows [row_begin, row_end) of the n x n matrix of ordered pairs:
 *   out_counts[(i - row_begin) * n + j] = |sets[i] n sets[j]|   (NULL: not wanted)
 *   out_sizes[i]                        = sets[i]->kmer_set_size() for all n sets (NULL: not wanted)
 *   out_ani[(i - row_begin) * n + j]    = binomial_estimator(containment(count, |sets[i]|), weight)  (NULL: not wanted)
 * SORTED sets of one mask go through a dictionary of their shared k-mers (csrc/sks_allpairs.cu) and the ANI is
 * evaluated on the device in double precision (CUDA pow, <= 2 ulp: within 1e-15 of the host's libm); any other
 * input takes sks_intersect_all_pairs + sks_ani_from_counts. */
int sks_all_vs_all(sks_ctx *ctx, sks_set *const *sets, int64_t n, int64_t row_begin, int64_t row_end, int32_t *out_counts,
                   int32_t *out_sizes, double *out_ani);
/* ANI matrix from counts, src/kmer-sketching.cpp:196-200: containment on the FIRST set of the
 * ordered pair, then ^(1/weight).  Host double arithmetic. */
void sks_ani_from_counts(const int32_t *intersections, const int32_t *first_set_sizes, int64_t n_pairs,
                         int weight, double *out_ani);

/* ---- several GPUs: one rank per GPU, NCCL over NVLink underneath -------------------------------------- */
/* The reference parallelises over FASTA files and over set pairs with cilk_for (src/kmer_set.cpp:124-131,179-182).
 * Here genome g of n belongs to rank g / ceil(n / world) (sks_shard_range); every rank sketches its block, every k-mer
 * travels once to the rank that owns it, and every rank returns its own complete block rows of the pair matrix.  A rank is a
 * process (sks_comm_init_rank: rank 0 makes the id, the caller's launcher carries its 128 bytes to the others) or a
 * thread of one process that drives one GPU (sks_comm_init_all).  NCCL is loaded at run time (libnccl.so.2, or
 * $SKS_NCCL_LIB); without it these calls fail with SKS_ERR_CUDA and the single-GPU API is unaffected. */
typedef struct sks_comm sks_comm;
#define SKS_COMM_ID_BYTES 128
void sks_shard_range(int64_t n, int rank, int world, int64_t *begin, int64_t *end);
int sks_comm_unique_id(void *out_id /* SKS_COMM_ID_BYTES */);
int sks_comm_init_rank(sks_ctx *ctx, const void *id, int rank, int world, sks_comm **out);
/* One process, n GPUs: ctxs[i] is a context on the i-th device; out[i] its communicator (rank i of n).  The sharded
 * calls below are then made by n threads, one per context. */
int sks_comm_init_all(sks_ctx *const *ctxs, int n, sks_comm **out);
void sks_comm_destroy(sks_comm *c);
int sks_comm_rank(const sks_comm *c);
int sks_comm_world(const sks_comm *c);
int sks_comm_nccl_version(void); /* 0 when NCCL cannot be loaded */
/* Every rank passes the sets of its block of the n_total genomes (SORTED, one mask) and receives all n_total sets
 * in genome order: one small all-gather of the key counts (the only host synchronisation) and one in-place
 * ncclAllGather of the keys, straight into the buffer the returned sets alias.  out_all[i] must be released with
 * sks_set_destroy; the rank's own sets come back as second handles on the same keys.  comm NULL = one rank. */
int sks_comm_allgather_sets(sks_ctx *ctx, sks_comm *comm, sks_set *const *local, int64_t n_local, int64_t n_total,
                            sks_set **out_all);
/* parallel_compute_pairwise_kmer_set_intersections over generate_all_pairs_from_vector + containment +
 * binomial_estimator (src/kmer-sketching.cpp:185-200), sharded.  The ranks split the KEY SPACE of the all-vs-all
 * dictionary, not the rows: a k-mer belongs to the rank its hash names; every rank sends each of its keys (with the
 * number of its set) to the owner -- one small all-gather of counts and one grouped send/receive, 1/world of what an
 * all-gather of the sketches would move (with two ranks the sketches are all-gathered instead and each rank enters
 * the keys it owns) --, the owner enters what arrives into its dictionary and counts what its keys contribute to
 * every pair, one NCCL reduce-scatter of the n x n partial counts adds the shares up, and the rank finalises its own
 * block rows [begin, end) = sks_shard_range(n_total, rank, world).  Every rank must make the call (it is collective)
 * with the same n_total.  (Sets the dictionary cannot take -- 16-byte keys of weight > 32 -- are all-gathered and
 * compared row block by row block instead.)  Outputs as in
 * sks_all_vs_all: out_counts / out_ani hold (end - begin) * n_total entries, out_sizes n_total. */
int sks_all_vs_all_sharded(sks_ctx *ctx, sks_comm *comm, sks_set *const *local, int64_t n_local, int64_t n_total,
                           int32_t *out_counts, int32_t *out_sizes, double *out_ani);
/* Sketch + exchange + comparison in one call for genomes already resident in HBM: `batch` holds the rank's block of the
 * n_total genomes (NULL: the rank has none).  = sks_sketch(SKS_REPR_SORTED) + sks_all_vs_all_sharded. */
int sks_all_vs_all_resident(sks_ctx *ctx, sks_comm *comm, const sks_batch *batch, int64_t n_total, const uint64_t mask[2],
                            int window, const sks_pred *pred, int32_t *out_counts, int32_t *out_sizes, double *out_ani);
/* The whole path from HOST buffers: packed[g] / n_bases[g] are the rank's n_local genomes (2-bit packed, one segment
 * each); they are brought to the device, sketched, exchanged and compared.  = parallel_kmer_sets_from_fasta_files + the
 * comparison loop of src/kmer-sketching.cpp:163-200.  Pinned host memory of 64 MB and more is copied chunk by chunk
 * (16 MB) by the copy engine while the sketch kernel works on the chunks that are there (sks_ctx_streamed_count);
 * pageable memory of that size goes the same way through pinned staging buffers that a few host threads fill
 * (SKS_HOST_THREADS, default: the cores divided by the ranks, at most 8).  Smaller pinned inputs are read in place by
 * the kernel (sks_ctx_in_place_count), smaller pageable ones are copied up first.  The call returns after the device
 * has finished with the buffers. */
int sks_all_vs_all_from_host(sks_ctx *ctx, sks_comm *comm, int n_local, const uint32_t *const *packed, const uint64_t *n_bases,
                             int64_t n_total, const uint64_t mask[2], int window, const sks_pred *pred, int32_t *out_counts,
                             int32_t *out_sizes, double *out_ani);
/* One long sequence split by position (BASELINE configs[2]): `slice` holds this rank's window starts plus a
 * (window - 1)-base halo (sks_batch_slice / sks_batch_synth_at).  The rank sketches its slice, the partial sketches
 * are routed by key range (rank r owns the r-th share of the key space under the mask: one grouped send/receive),
 * and every rank sort-uniques its range only.  gather == 0: *out is the rank's range of the global set (the ranges
 * are disjoint and ordered by rank); gather != 0: every rank receives the whole set.  *out_global_size = |set|. */
int sks_sketch_sequence_sharded(sks_ctx *ctx, sks_comm *comm, const sks_batch *slice, const uint64_t mask[2], int window,
                                const sks_pred *pred, int gather, sks_set **out, int64_t *out_global_size);

/* ---- one-call pair pipeline (bench / e2e) ---------------------------------------------------- */
typedef struct sks_pair_result {
  int64_t size_a, size_b, intersection;
  double ani_ab, ani_ba; /* binomial_estimator(containment(I, |A|), weight), and with |B| */
} sks_pair_result;
/* HOST packed genomes in, counts + ANI out: upload, sketch both, intersect, sizes, ANI.
 * = kmer_sets_from_fasta_files on two genomes + kmer_set_intersection + containment / binomial_estimator
 * (src/kmer-sketching.cpp:168-200 for n = 2).  With SKS_REPR_BITSET both 4^weight-bit sets are materialised in
 * HBM and |A|, |B|, |A n B| are counted while their slices stream out (one fused kernel instead of a build
 * and a re-read); SKS_REPR_BITSET_ONCHIP skips the HBM copy of the sets.
 * Genomes in pinned (page-locked, device-mapped) host memory at 16-byte aligned addresses are not copied at all:
 * the sketch kernel's bulk copies read them in place, tile by tile, while it computes (sks_ctx_in_place_count
 * counts such calls; SKS_ZERO_COPY=0 turns it off).  Any other host memory is copied to the device first.  The call
 * returns after the device has finished with the buffers either way. */
int sks_pair_ani(sks_ctx *ctx, const uint32_t *packed_a, uint64_t n_bases_a, const uint32_t *packed_b,
                 uint64_t n_bases_b, const uint64_t mask[2], int window, const sks_pred *pred, int repr,
                 sks_pair_result *out);
/* Same on a resident 2-genome batch (inputs already in HBM). */
int sks_pair_ani_resident(sks_ctx *ctx, const sks_batch *batch, const uint64_t mask[2], int window,
                          const sks_pred *pred, int repr, sks_pair_result *out);

#ifdef __cplusplus
}
#endif
#endif /* SKS_H */
