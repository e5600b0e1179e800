/**
 * kmer.hpp -- drop-in for the reference's main header (src/kmer.hpp): same types, constants and free
 * functions, implemented on the B200 through the C ABI of libsks.so (include/sks.h).
 *
 * What differs from the reference, deliberately:
 *   - kmer_bitset is this repository's own 128-bit value type (kmer_bitset.hpp), not
 *     boost::dynamic_bitset; Boost and OpenCilk are not dependencies.
 *   - kmer_set keeps its members on the device (sorted distinct keys or a 4^weight-bit presence
 *     bitset); the `kmer_hashes` table is materialised on the host only when a caller touches it.
 *   - the sketching condition is still a std::function<bool(const kmer)>.  Recognised forms run on the
 *     device: sks::all_kmers() and sks::fmh_condition(nonce, c).  Any other callable is opaque: it is
 *     evaluated on the host, exactly once per real k-mer and in sequence order as in the reference
 *     (src/kmer_sliding.cpp:183), over the canonical k-mer list the device produces (window sliding and
 *     canonicalisation stay on the GPU).  Opt-in (sks::enable_predicate_probe(true) or SKS_PREDICATE_PROBE=1):
 *     an opaque callable is first PROBED -- called a few thousand times with scripted frac_min_hash values --
 *     and, if it behaves as `frac_min_hash(n)(k) % c == 0` like the reference driver's free function
 *     (src/kmer-sketching.cpp:29-34), run on the device as FMH{n, c}.  Probing is off by default because a
 *     stateful callable observes the extra calls and a condition of the form `fmh(k) % c == 0 && extra(k)`
 *     can pass it when `extra` holds on the probe k-mers.
 */
#ifndef SKS_KMER_HPP
#define SKS_KMER_HPP

// STL includes
#include "stl_includes.hpp"

#include <memory>

#include "kmer_bitset.hpp"
#include "logging.hpp"

/** src/kmer.hpp:35-37: 7 -> 128-bit bitsets -> up to 64-mers.  Fixed in this implementation. */
constexpr int LOG_KMER_BITSET_SIZE = 7;

/** src/kmer.hpp:43.  There is no Cilk here; the flag only selects the serial entry points' behaviour. */
constexpr int PARALLEL_DISABLE = DEBUG | 0;

/** src/kmer.hpp:52-54 */
constexpr int NUCLEOTIDE_BIT_SIZE = 2;
constexpr int KMER_BITSET_SIZE = (1 << LOG_KMER_BITSET_SIZE);
constexpr int MAX_KMER_LENGTH = (KMER_BITSET_SIZE / NUCLEOTIDE_BIT_SIZE);

// Masks (src/kmer.hpp:57-64).  The two initialise_* functions are kept for source compatibility; the
// tables they filled in the reference are not needed (both are no-ops).
void initialise_contiguous_kmer_array();
kmer_bitset contiguous_kmer(const int kmer_length); // throws std::runtime_error for kmer_length > 64
void initialise_reversing_kmer_array();
kmer_bitset reverse_kmer_bitset(const kmer_bitset &kbs);
kmer_bitset generate_random_spaced_seed_mask(
    const int window_size,
    const int kmer_size,
    size_t random_seed = 0);

/**
 * src/kmer.hpp:75-86.
 * @param window_length Length of the whole kmer window
 * @param kmer_bits Raw bits of the chosen strand (the forward strand keeps earlier bases above 2w)
 * @param mask Mask used for the kmer
 * @param masked_bits kmer_bits & mask
 */
struct kmer
{
    int window_length;
    kmer_bitset kmer_bits;
    kmer_bitset mask;
    kmer_bitset masked_bits;

    bool operator==(const kmer &other) const
    {
        return (masked_bits == other.masked_bits) && (mask == other.mask);
    }
};

// Legacy canonicalisation (src/kmer.hpp:89-90, src/kmers.cpp:16-35); host only.
kmer reverse_complement(kmer k);
kmer canonical_kmer(kmer k);

// Ordered, duplicate-preserving k-mer lists (src/kmer.hpp:93-103, src/kmer_sliding.cpp:199-238).
std::vector<kmer> nucleotide_string_list_to_kmers(
    const std::vector<std::vector<uint8_t>> &nucleotide_strings,
    const kmer_bitset &mask,
    const int window_length,
    const std::function<bool(const kmer)> &sketching_cond);
void nucleotide_string_list_to_kmers_by_reference(
    std::vector<kmer> &kmer_list,
    const std::vector<std::vector<uint8_t>> &nucleotide_strings,
    const kmer_bitset &mask,
    const int window_length,
    const std::function<bool(const kmer)> &sketching_cond);

/** src/kmer.hpp:113-124: orders the host hash table only. */
struct kmer_hash
{
    std::hash<kmer_bitset> kmer_bitset_std_hash;
    std::hash<int> int_std_hash;
    inline size_t operator()(const kmer &k) const
    {
        return kmer_bitset_std_hash(k.masked_bits) ^ kmer_bitset_std_hash(k.mask) ^ int_std_hash(k.window_length);
    }
};

namespace sks
{
/** boost::hash<kmer_bitset> / boost::hash<int> stand-ins used by frac_min_hash. */
struct boost_bitset_hash
{
    size_t operator()(const kmer_bitset &b) const { return boost_hash_value(b); }
};
struct boost_int_hash
{
    size_t operator()(int v) const { return static_cast<size_t>(v); } // Boost hashes integers to themselves
};
/** When a sketching condition is being probed, frac_min_hash returns scripted values (see kmer.hpp top). */
bool fmh_probe_active();
size_t fmh_probe_value(int nonce);
} // namespace sks

/**
 * src/kmer.hpp:135-149: FracMinHash functor, H(masked_bits) ^ H(mask) ^ window_length ^ nonce with Boost's
 * hash_value(dynamic_bitset) (both hash_combine generations are implemented; see sks::set_boost_hash_variant).
 */
struct frac_min_hash
{
    sks::boost_bitset_hash kmer_bitset_boost_hash;
    sks::boost_int_hash int_boost_hash;
    int nonce;

    frac_min_hash(int n) : nonce(static_cast<int>(int_boost_hash(n))) {}

    inline size_t operator()(const kmer &k) const
    {
        if (sks::fmh_probe_active()) return sks::fmh_probe_value(nonce);
        return kmer_bitset_boost_hash(k.masked_bits) ^ kmer_bitset_boost_hash(k.mask) ^
               int_boost_hash(k.window_length) ^ static_cast<size_t>(nonce);
    }
};

// hash table for kmers (src/kmer.hpp:152)
typedef std::unordered_map<kmer, int, kmer_hash> kmer_hash_table;

struct kmer_set;

namespace sks
{
struct device_set; // RAII owner of an sks_set handle (sks_cpp.cpp)

/**
 * `kmer_set::kmer_hashes`: behaves like the reference's kmer_hash_table member, but its contents live on
 * the device until someone looks; any access through table() / the forwarding members downloads the
 * keys once and rebuilds the host map.
 */
class lazy_kmer_table
{
public:
    typedef kmer_hash_table::iterator iterator;
    typedef kmer_hash_table::const_iterator const_iterator;

    kmer_hash_table &table();
    const kmer_hash_table &table() const;
    operator kmer_hash_table &() { return table(); }
    operator const kmer_hash_table &() const { return table(); }

    size_t size() const;
    bool empty() const { return size() == 0; }
    iterator begin() { return table().begin(); }
    iterator end() { return table().end(); }
    const_iterator begin() const { return table().begin(); }
    const_iterator end() const { return table().end(); }
    iterator find(const kmer &k) { return table().find(k); }
    const_iterator find(const kmer &k) const { return table().find(k); }
    size_t count(const kmer &k) const { return table().count(k); }
    int &operator[](const kmer &k);
    void clear();

private:
    friend struct ::kmer_set;
    friend struct set_access;
    mutable kmer_hash_table host_;
    mutable bool host_valid_ = true;           // host_ mirrors the set
    mutable std::shared_ptr<device_set> dev_;  // device copy; null when only the host map exists
    mutable bool dev_valid_ = false;
    void materialise() const;
    void touch_host(); // the host map is about to be modified: the device copy becomes stale
};
} // namespace sks

/** src/kmer.hpp:160-190 */
struct kmer_set
{
    sks::lazy_kmer_table kmer_hashes;

    /** Inserts k-mers given on the host (src/kmer.hpp:170-178). */
    void insert_kmers(const std::vector<kmer> &kmers)
    {
        for (const kmer &k : kmers)
        {
            if (DEBUG)
                std::cout << "Inserting kmer " << k.masked_bits << std::endl;
            kmer_hashes[k] = 1;
        }
    }

    /** Number of distinct k-mers (src/kmer.hpp:186-189); answered by the device when the set lives there. */
    inline int kmer_set_size() const
    {
        return static_cast<int>(kmer_hashes.size());
    }
};

// |ks1 n ks2| (src/kmer.hpp:192, src/kmer_set.cpp:23-41)
int kmer_set_intersection(const kmer_set &ks1, const kmer_set &ks2);

// FASTA -> sets (src/kmer.hpp:195-212, src/kmer_set.cpp:54-133).  All files of a call are sketched by one
// batched launch; the "parallel" variant is the same code (there is no cilk_for to switch off).
kmer_set kmer_set_from_fasta_file(
    const char fasta_filename[],
    const kmer_bitset &mask,
    const int window_length,
    const std::function<bool(const kmer)> &sketching_cond);
std::vector<kmer_set> kmer_sets_from_fasta_files(
    const int num_files,
    char *fasta_filenames[],
    const kmer_bitset &mask,
    const int window_length,
    const std::function<bool(const kmer)> &sketching_cond);
std::vector<kmer_set> parallel_kmer_sets_from_fasta_files(
    const int num_files,
    char *fasta_filenames[],
    const kmer_bitset &mask,
    const int window_length,
    const std::function<bool(const kmer)> &sketching_cond);
// Pair lists (src/kmer.hpp:213-218, src/kmer_set.cpp:143-184): one launch over the whole list; lists of
// different lengths throw std::runtime_error.
std::vector<int> compute_pairwise_kmer_set_intersections(
    const std::vector<kmer_set *> &kmer_sets_1,
    const std::vector<kmer_set *> &kmer_sets_2);
std::vector<int> parallel_compute_pairwise_kmer_set_intersections(
    const std::vector<kmer_set *> &kmer_sets_1,
    const std::vector<kmer_set *> &kmer_sets_2);

// ---- additions (not in the reference) -----------------------------------------------------------
namespace sks
{
/** Functor types the set builders recognise through std::function::target<T>(). */
struct all_kmers
{
    bool operator()(const kmer) const { return true; }
};
struct fmh_condition
{
    int nonce;
    uint64_t modulus;
    fmh_condition(int n = 1, uint64_t c = 200) : nonce(n), modulus(c) {}
    bool operator()(const kmer k) const { return frac_min_hash(nonce)(k) % modulus == 0; }
};

/** README seed-string notation ("11001011", README.md:25-41) -> mask; window_length = strlen. */
kmer_bitset seed_string_to_mask(const std::string &seed);

/** Device and representation used by the calling thread's implicit context. */
void set_device(int device);
enum set_representation { REPR_AUTO = 0, REPR_SORTED = 1, REPR_BITSET = 2 };
void set_representation_hint(set_representation r);

/** How the last sketching condition seen by this thread was executed: "device:all", "device:fmh" or "host". */
const char *last_predicate_path();
/** Probing of opaque sketching conditions (see the top of this header); process-wide, off by default. */
void enable_predicate_probe(bool on);
/**
 * Several GPUs in one process (default 1; also SKS_DEVICES=n|all).  With n > 1, kmer_sets_from_fasta_files and
 * parallel_kmer_sets_from_fasta_files sketch contiguous blocks of files on devices 0 .. n-1 (one worker thread each --
 * the reference's cilk_for over files, src/kmer_set.cpp:124-131), and (parallel_)compute_pairwise_kmer_set_intersections
 * runs an all-pairs list (generate_all_pairs_from_vector) over such sets as the sharded all-vs-all: one NCCL exchange
 * of the sketches, every device fills its block rows (the reference's cilk_for over pairs, src/kmer_set.cpp:179-182).
 * Any other pair list that mixes devices is evaluated on the calling thread's device after peer copies.
 */
void set_devices(int n);
int devices();
} // namespace sks

#endif // SKS_KMER_HPP
