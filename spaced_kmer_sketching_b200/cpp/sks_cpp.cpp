// libsks_cpp.so -- the reference-shaped C++ API (include/kmer.hpp, fasta_processing.hpp,
// ani_estimator.hpp) on top of the C ABI of libsks.so (include/sks.h).  Host logic only: every hot
// stage (window sliding, canonicalisation, FracMinHash filter, set build, intersection) is a CUDA
// kernel behind the sks_* calls.  C-ABI failures surface as std::runtime_error, like the reference's
// own errors (src/kmer_bitset.cpp:53-54, src/kmer_set.cpp:147-150).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <thread>
#include <sstream>
#include <string>

#include "../../include/ani_estimator.hpp"
#include "../../include/fasta_processing.hpp"
#include "../../include/kmer.hpp"
#include "../../include/sks.h"

namespace sks
{
namespace
{
[[noreturn]] void fail(const char *what)
{
    std::string msg = std::string(what) + ": " + sks_last_error();
    throw std::runtime_error(msg);
}
inline void check(int status, const char *what)
{
    if (status != SKS_OK) fail(what);
}

// ---- per-thread implicit context (the reference's functions take none) --------------------------
struct thread_state
{
    int device = -1;
    sks_ctx *ctx = nullptr;
    bool owns_ctx = true; // false: a worker thread of run_on_devices borrows a context of the device pool
    int repr = SKS_REPR_AUTO;
    const char *pred_path = "none";
    ~thread_state()
    {
        if (ctx && owns_ctx) sks_ctx_destroy(ctx);
    }
};
thread_local thread_state g_ts;

sks_ctx *ctx()
{
    if (!g_ts.ctx)
    {
        int dev = g_ts.device;
        if (dev < 0)
        {
            const char *e = getenv("SKS_DEVICE");
            dev = e ? atoi(e) : 0;
        }
        check(sks_ctx_create(dev, &g_ts.ctx), "sks_ctx_create");
        g_ts.device = dev;
    }
    return g_ts.ctx;
}

int g_hash_variant = SKS_HASH_BOOST_181;

// ---- several GPUs in one process ----------------------------------------------------------------------
// The reference spreads files and pairs over Cilk workers (src/kmer_set.cpp:124-131,179-182); with sks::set_devices(n)
// (or SKS_DEVICES=n|all) the same two loops are spread over n GPUs: file i of m goes to device
// i / ceil(m / n) (sks_shard_range), one worker thread per device, and an all-pairs comparison of sets that were
// sketched that way runs as the sharded all-vs-all of the C ABI (sks_all_vs_all_sharded over sks_comm_init_all).
std::atomic<int> g_devices{0}; // 0: follow SKS_DEVICES
int devices_wanted()
{
    int n = g_devices.load();
    if (n <= 0)
    {
        const char *e = getenv("SKS_DEVICES");
        if (!e || !*e) return 1;
        n = (strcmp(e, "all") == 0) ? sks_device_count() : atoi(e);
    }
    return std::max(1, std::min(n, sks_device_count()));
}

struct device_pool
{
    std::mutex mu;
    std::vector<sks_ctx *> ctxs;
    std::vector<sks_comm *> comms;
    // contexts (and, on demand, communicators) for devices 0 .. n-1; grown under the lock, never shrunk
    void ensure(int n, bool with_comms)
    {
        std::lock_guard<std::mutex> lock(mu);
        while ((int)ctxs.size() < n)
        {
            sks_ctx *c = nullptr;
            check(sks_ctx_create((int)ctxs.size(), &c), "sks_ctx_create");
            ctxs.push_back(c);
        }
        if (with_comms && (int)comms.size() != n)
        {
            for (sks_comm *c : comms) sks_comm_destroy(c);
            comms.assign((size_t)n, nullptr);
            check(sks_comm_init_all(ctxs.data(), n, comms.data()), "sks_comm_init_all");
        }
    }
};
device_pool &pool()
{
    static device_pool *p = new device_pool(); // never destroyed: the CUDA runtime may be gone at exit
    return *p;
}

// fn(rank) on one thread per device; the worker's implicit context is the pool's context of that device.
template <typename F>
void run_on_devices(int n, F fn)
{
    std::vector<std::thread> workers;
    std::vector<std::string> errors((size_t)n);
    const int repr = g_ts.repr;
    for (int r = 0; r < n; ++r)
        workers.emplace_back([&, r]() {
            g_ts.device = r;
            g_ts.ctx = pool().ctxs[(size_t)r];
            g_ts.owns_ctx = false;
            g_ts.repr = repr;
            try
            {
                fn(r);
            }
            catch (const std::exception &e)
            {
                errors[(size_t)r] = e.what()[0] ? e.what() : "error";
            }
            g_ts.ctx = nullptr;
        });
    for (std::thread &t : workers) t.join();
    for (const std::string &e : errors)
        if (!e.empty()) throw std::runtime_error(e);
}

// ---- probing of opaque sketching conditions ------------------------------------------------------
struct probe_state
{
    bool active = false;
    size_t value = 0;
    int calls = 0;
    int nonce = 0;
};
thread_local probe_state g_probe;

struct pred_plan
{
    enum kind_t { HOST, DEVICE } kind = HOST;
    sks_pred pred{};
};

std::atomic<int> g_probe_switch{-1};  // -1: follow the environment, 0 / 1: set by sks::enable_predicate_probe
bool probe_enabled()
{
    const int sw = g_probe_switch.load();
    if (sw >= 0) return sw != 0;
    static const bool env = [] { const char *e = getenv("SKS_PREDICATE_PROBE"); return e && atoi(e) != 0; }();
    return env;
}

uint64_t splitmix(uint64_t &s)
{
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Decides how a std::function<bool(const kmer)> is executed (kmer.hpp, top comment).
pred_plan classify(const std::function<bool(const kmer)> &f, const kmer_bitset &mask, int window)
{
    pred_plan plan;
    plan.pred.hash_variant = g_hash_variant;
    if (!f) throw std::bad_function_call();
    if (f.target<all_kmers>())
    {
        plan.kind = pred_plan::DEVICE;
        plan.pred.kind = SKS_PRED_ALL;
        g_ts.pred_path = "device:all";
        return plan;
    }
    if (const fmh_condition *c = f.target<fmh_condition>())
    {
        if (c->modulus == 0) throw std::runtime_error("fmh_condition: modulus must be non-zero");
        plan.kind = pred_plan::DEVICE;
        plan.pred.kind = SKS_PRED_FMH;
        plan.pred.nonce = c->nonce;
        plan.pred.modulus = c->modulus;
        g_ts.pred_path = "device:fmh";
        return plan;
    }
    g_ts.pred_path = "host";
    // Any other callable is opaque.  By default it is evaluated on the host, exactly once per real k-mer, as the
    // reference does (src/kmer_sliding.cpp:183).  Probing is opt-in (sks::enable_predicate_probe / SKS_PREDICATE_PROBE=1):
    // it calls the callable a few thousand times with scripted hash values, which a stateful callable observes, and a
    // condition that merely contains `fmh(k) % c == 0` can pass it.
    if (!probe_enabled()) return plan;

    // Probe: while g_probe.active, frac_min_hash::operator() returns g_probe.value.  The callable is
    // accepted as `frac_min_hash(n)(k) % c == 0` when it calls the hash exactly once per k-mer, accepts
    // 0, first accepts again at c, and then agrees with (v % c == 0) on scripted values for several
    // different k-mers.
    uint64_t rng = 0x5eed5eedull;
    kmer probes[8];
    for (kmer &p : probes)
    {
        kmer_bitset bits = kmer_bitset::from_words(splitmix(rng), splitmix(rng));
        p = kmer{window, bits & mask, mask, bits & mask};
    }
    struct guard
    {
        guard() { g_probe = probe_state{true, 0, 0, 0}; }
        ~guard() { g_probe.active = false; }
    } on;
    auto ask = [&](size_t v, const kmer &k, bool &ok) {
        g_probe.value = v;
        const int before = g_probe.calls;
        const bool r = f(k);
        if (g_probe.calls != before + 1) ok = false;
        return r;
    };
    bool ok = true;
    if (!ask(0, probes[0], ok) || !ok) return plan;
    const int nonce = g_probe.nonce;
    uint64_t c = 0;
    for (uint64_t v = 1; v <= (1u << 22) && ok; ++v)
        if (ask(v, probes[0], ok))
        {
            c = v;
            break;
        }
    if (!ok || c == 0) return plan;
    for (int t = 0; t < 4096 && ok; ++t)
    {
        uint64_t v = splitmix(rng);
        if (t & 1) v = (v % (~0ull / c)) * c;       // multiples of c
        if ((t & 3) == 2) v = (v % (1u << 16));      // small values
        if (ask(v, probes[t & 7], ok) != (v % c == 0)) ok = false;
        if (g_probe.nonce != nonce) ok = false;
    }
    if (!ok) return plan;
    plan.kind = pred_plan::DEVICE;
    plan.pred.kind = SKS_PRED_FMH;
    plan.pred.nonce = nonce;
    plan.pred.modulus = c;
    g_ts.pred_path = "device:fmh";
    return plan;
}

void mask_words(const kmer_bitset &mask, uint64_t out[2])
{
    out[0] = mask.word(0);
    out[1] = mask.word(1);
}

kmer make_kmer(int window, const kmer_bitset &mask, const uint64_t masked[2], const uint64_t bits[2])
{
    return kmer{window, kmer_bitset::from_words(bits[0], bits[1]), mask, kmer_bitset::from_words(masked[0], masked[1])};
}

struct batch_guard
{
    sks_batch *b = nullptr;
    ~batch_guard()
    {
        if (b) sks_batch_destroy(ctx(), b);
    }
};
} // namespace

// ---- device_set -------------------------------------------------------------------------------------
struct device_set
{
    sks_set *h = nullptr;
    int window = 0;
    kmer_bitset mask;
    explicit device_set(sks_set *s, int w, const kmer_bitset &m) : h(s), window(w), mask(m) {}
    ~device_set()
    {
        if (h) sks_set_destroy(nullptr, h);
    }
    device_set(const device_set &) = delete;
    device_set &operator=(const device_set &) = delete;
};

struct set_access
{
    static void adopt(kmer_set &ks, sks_set *h, int window, const kmer_bitset &mask)
    {
        ks.kmer_hashes.dev_ = std::make_shared<device_set>(h, window, mask);
        ks.kmer_hashes.dev_valid_ = true;
        ks.kmer_hashes.host_valid_ = false;
        ks.kmer_hashes.host_.clear();
    }
    // Returns the device copy of a set, uploading the host table first when it is the newer one.
    static device_set *device(const kmer_set &ks)
    {
        const lazy_kmer_table &t = ks.kmer_hashes;
        if (t.dev_valid_ && t.dev_) return t.dev_.get();
        // host -> device: all members must share one mask (sets built by this API always do)
        if (t.host_.empty()) return nullptr;
        const kmer &first = t.host_.begin()->first;
        std::vector<uint64_t> keys;
        keys.reserve(t.host_.size() * 2);
        for (const auto &kv : t.host_)
        {
            if (!(kv.first.mask == first.mask))
                throw std::runtime_error("kmer_set mixes masks: such a set cannot be moved to the device");
            keys.push_back(kv.first.masked_bits.word(0));
            keys.push_back(kv.first.masked_bits.word(1));
        }
        uint64_t m[2];
        mask_words(first.mask, m);
        sks_set *h = nullptr;
        check(sks_set_from_host_keys(ctx(), keys.data(), (int64_t)t.host_.size(), m, first.window_length, &h),
              "sks_set_from_host_keys");
        t.dev_ = std::make_shared<device_set>(h, first.window_length, first.mask);
        t.dev_valid_ = true;
        return t.dev_.get();
    }
};

// ---- lazy_kmer_table ----------------------------------------------------------------------------------
void lazy_kmer_table::materialise() const
{
    if (host_valid_) return;
    host_.clear();
    if (dev_ && dev_valid_)
    {
        int64_t n = 0;
        check(sks_set_size(ctx(), dev_->h, &n), "sks_set_size");
        std::vector<uint64_t> keys((size_t)n * 2);
        check(sks_set_keys(ctx(), dev_->h, keys.data(), (uint64_t)n), "sks_set_keys");
        host_.reserve((size_t)n);
        for (int64_t i = 0; i < n; ++i)
        {
            // the device keeps (masked_bits, mask) -- what equality and hashing use; kmer_bits is not kept
            const uint64_t *k = &keys[(size_t)i * 2];
            host_[make_kmer(dev_->window, dev_->mask, k, k)] = 1;
        }
    }
    host_valid_ = true;
}
kmer_hash_table &lazy_kmer_table::table()
{
    materialise();
    touch_host();
    return host_;
}
const kmer_hash_table &lazy_kmer_table::table() const
{
    materialise();
    return host_;
}
size_t lazy_kmer_table::size() const
{
    if (host_valid_) return host_.size();
    int64_t n = 0;
    check(sks_set_size(ctx(), dev_->h, &n), "sks_set_size");
    return (size_t)n;
}
int &lazy_kmer_table::operator[](const kmer &k)
{
    materialise();
    touch_host();
    return host_[k];
}
void lazy_kmer_table::clear()
{
    host_.clear();
    host_valid_ = true;
    dev_.reset();
    dev_valid_ = false;
}
void lazy_kmer_table::touch_host()
{
    dev_valid_ = false; // a non-const view may change the table: rebuild the device copy on next use
    dev_.reset();
}

// ---- misc additions ----------------------------------------------------------------------------------------
int boost_hash_variant() { return g_hash_variant; }
void set_boost_hash_variant(int variant)
{
    if (variant != BOOST_HASH_171 && variant != BOOST_HASH_181) throw std::runtime_error("unknown Boost hash variant");
    g_hash_variant = variant;
}
std::size_t boost_hash_value(const kmer_bitset &b)
{
    return (std::size_t)sks_boost_hash_bitset(b.words(), g_hash_variant);
}
bool fmh_probe_active() { return g_probe.active; }
size_t fmh_probe_value(int nonce)
{
    g_probe.calls++;
    g_probe.nonce = nonce;
    return g_probe.value;
}
kmer_bitset seed_string_to_mask(const std::string &seed)
{
    uint64_t m[2];
    int w = 0;
    check(sks_seed_to_mask(seed.c_str(), m, &w), "seed_string_to_mask");
    return kmer_bitset::from_words(m[0], m[1], KMER_BITSET_SIZE);
}
void set_device(int device)
{
    if (g_ts.ctx && g_ts.device != device)
    {
        sks_ctx_destroy(g_ts.ctx);
        g_ts.ctx = nullptr;
    }
    g_ts.device = device;
}
void set_representation_hint(set_representation r) { g_ts.repr = (int)r; }
void set_devices(int n) { g_devices.store(n); }
int devices() { return devices_wanted(); }
const char *last_predicate_path() { return g_ts.pred_path; }
void enable_predicate_probe(bool on) { g_probe_switch.store(on ? 1 : 0); }
} // namespace sks

using sks::check;
using sks::ctx;

// ---- masks ----------------------------------------------------------------------------------------------------
void initialise_contiguous_kmer_array() {}
void initialise_reversing_kmer_array() {}

kmer_bitset contiguous_kmer(const int kmer_length)
{
    uint64_t m[2];
    if (sks_contiguous_mask(kmer_length, m) != SKS_OK)
        throw std::runtime_error("Given k-mer length exceeds maximum k-mer length");
    return kmer_bitset::from_words(m[0], m[1], KMER_BITSET_SIZE);
}

kmer_bitset reverse_kmer_bitset(const kmer_bitset &kbs)
{
    uint64_t out[2];
    sks_reverse_bitset(kbs.words(), out);
    return kmer_bitset::from_words(out[0], out[1], KMER_BITSET_SIZE);
}

kmer_bitset generate_random_spaced_seed_mask(const int window_size, const int kmer_size, size_t random_seed)
{
    uint64_t m[2];
    check(sks_random_mask(window_size, kmer_size, random_seed, m), "generate_random_spaced_seed_mask");
    return kmer_bitset::from_words(m[0], m[1], KMER_BITSET_SIZE);
}

// ---- legacy canonicalisation (src/kmers.cpp:16-35) --------------------------------------------------------------
kmer reverse_complement(kmer k)
{
    kmer_bitset rc = reverse_kmer_bitset(k.kmer_bits);
    rc.flip();
    rc >>= (size_t)((MAX_KMER_LENGTH - k.window_length) * NUCLEOTIDE_BIT_SIZE);
    return kmer{k.window_length, rc, k.mask, rc & k.mask};
}
kmer canonical_kmer(kmer k)
{
    kmer rc = reverse_complement(k);
    return (k.masked_bits < rc.masked_bits) ? k : rc;
}

// ---- ANI ------------------------------------------------------------------------------------------------------------
double containment(int intersection, int set_size) { return sks_containment(intersection, set_size); }
double binomial_estimator(double c, int kmer_num_ones) { return sks_binomial_estimator(c, kmer_num_ones); }

// ---- FASTA (host only; same rules as src/fasta_processing.cpp:79-211) ---------------------------------------------------
std::vector<std::string> strings_from_fasta(const char fasta_filename[])
{
    std::ifstream in(fasta_filename);
    if (!in.good())
    {
        std::cerr << "Unable to open " << fasta_filename << ". \n Exiting..." << std::endl;
        exit(1);
    }
    std::vector<std::string> records;
    std::string name, body, line;
    auto flush = [&]() {
        if (!name.empty())
        {
            if (LOGGING) std::clog << INFO_LOG << "Read " << name << " from file " << fasta_filename << std::endl;
            records.push_back(body);
        }
    };
    while (std::getline(in, line))
    {
        const bool header = !line.empty() && line[0] == '>';
        if (line.empty() || header)
        {
            flush();                          // a blank line ends the record but keeps its name
            if (header) name = line.substr(1);
            body.clear();
        }
        else if (!name.empty())
        {
            if (line.find(' ') != std::string::npos)
            { // a space drops the record, and everything up to the next header
                name.clear();
                body.clear();
            }
            else
                body += line;
        }
    }
    flush();
    return records;
}

void add_nucleotide_strings(std::vector<acgt_string> &return_strings, const std::string &raw_string)
{
    static const struct table_t
    {
        uint8_t t[256];
        table_t()
        {
            memset(t, 4, sizeof(t));
            t[(unsigned char)'A'] = t[(unsigned char)'a'] = 0;
            t[(unsigned char)'C'] = t[(unsigned char)'c'] = 1;
            t[(unsigned char)'G'] = t[(unsigned char)'g'] = 2;
            t[(unsigned char)'T'] = t[(unsigned char)'t'] = 3;
        }
    } codes;
    acgt_string run;
    for (const char ch : raw_string)
    {
        const uint8_t code = codes.t[(unsigned char)ch];
        if (code & 0x4)
        {
            if (!run.empty()) return_strings.push_back(run);
            run.clear();
        }
        else
            run.push_back(code);
    }
    if (!run.empty()) return_strings.push_back(run);
}

std::vector<acgt_string> cut_nucleotide_strings(const std::vector<std::string> &raw_strings)
{
    std::vector<acgt_string> out;
    for (const std::string &s : raw_strings) add_nucleotide_strings(out, s);
    return out;
}

std::vector<acgt_string> nucleotide_strings_from_fasta_file(const char fasta_filename[])
{
    return cut_nucleotide_strings(strings_from_fasta(fasta_filename));
}

// ---- ordered k-mer lists ----------------------------------------------------------------------------------------------
namespace
{
// masked / bits lists of one resident genome -> vector<kmer>, filtered on the host when the plan says so
void append_list(std::vector<kmer> &out, sks_batch *batch, int genome, const kmer_bitset &mask, int window,
                 const sks::pred_plan &plan, const std::function<bool(const kmer)> &cond)
{
    uint64_t m[2];
    sks::mask_words(mask, m);
    sks_pred all{};
    all.kind = SKS_PRED_ALL;
    const sks_pred *p = plan.kind == sks::pred_plan::DEVICE ? &plan.pred : &all;
    uint64_t n = 0;
    check(sks_kmer_list(ctx(), batch, genome, m, window, p, &n, nullptr, nullptr, 0), "sks_kmer_list");
    if (n == 0) return;
    std::vector<uint64_t> masked(n * 2), bits(n * 2);
    check(sks_kmer_list(ctx(), batch, genome, m, window, p, &n, masked.data(), bits.data(), n), "sks_kmer_list");
    out.reserve(out.size() + n);
    for (uint64_t i = 0; i < n; ++i)
    {
        kmer k = sks::make_kmer(window, mask, &masked[i * 2], &bits[i * 2]);
        if (plan.kind == sks::pred_plan::DEVICE || cond(k)) out.push_back(k);
    }
}

sks_batch *upload_strings(const std::vector<std::vector<uint8_t>> &strings)
{
    std::vector<uint8_t> codes;
    std::vector<uint64_t> seg;
    for (const auto &s : strings)
    {
        if (s.empty()) continue;
        codes.insert(codes.end(), s.begin(), s.end());
        seg.push_back(s.size());
    }
    std::vector<uint32_t> words(sks_packed_words(codes.size()) + 1);
    check(sks_pack_codes(codes.data(), codes.size(), words.data()), "sks_pack_codes");
    const uint32_t *pw = words.data();
    const uint64_t nb = codes.size(), ns = seg.size();
    const uint64_t *ps = seg.data();
    sks_batch *b = nullptr;
    check(sks_batch_upload(ctx(), 1, &pw, &nb, ns ? &ps : nullptr, ns ? &ns : nullptr, &b), "sks_batch_upload");
    return b;
}
} // namespace

void nucleotide_string_list_to_kmers_by_reference(std::vector<kmer> &kmer_list,
                                                  const std::vector<std::vector<uint8_t>> &nucleotide_strings,
                                                  const kmer_bitset &mask, const int window_length,
                                                  const std::function<bool(const kmer)> &sketching_cond)
{
    const sks::pred_plan plan = sks::classify(sketching_cond, mask, window_length);
    sks::batch_guard bg;
    bg.b = upload_strings(nucleotide_strings);
    append_list(kmer_list, bg.b, 0, mask, window_length, plan, sketching_cond);
}

std::vector<kmer> nucleotide_string_list_to_kmers(const std::vector<std::vector<uint8_t>> &nucleotide_strings,
                                                  const kmer_bitset &mask, const int window_length,
                                                  const std::function<bool(const kmer)> &sketching_cond)
{
    std::vector<kmer> out;
    nucleotide_string_list_to_kmers_by_reference(out, nucleotide_strings, mask, window_length, sketching_cond);
    return out;
}

// ---- FASTA -> sets ------------------------------------------------------------------------------------------------------
namespace
{
std::vector<kmer_set> kmer_sets_on_this_device(const int num_files, char *fasta_filenames[], const kmer_bitset &mask,
                                               const int window_length, const sks::pred_plan &plan,
                                               const std::function<bool(const kmer)> &sketching_cond);
}

std::vector<kmer_set> kmer_sets_from_fasta_files(const int num_files, char *fasta_filenames[], const kmer_bitset &mask,
                                                 const int window_length,
                                                 const std::function<bool(const kmer)> &sketching_cond)
{
    if (num_files <= 0) return std::vector<kmer_set>();
    const sks::pred_plan plan = sks::classify(sketching_cond, mask, window_length);
    const int nd = std::min(sks::devices_wanted(), num_files);
    // an opaque condition is evaluated by the caller's callable: keep that on the calling thread
    if (nd <= 1 || plan.kind != sks::pred_plan::DEVICE)
        return kmer_sets_on_this_device(num_files, fasta_filenames, mask, window_length, plan, sketching_cond);
    // contiguous blocks of files per device, one worker thread each (the reference: cilk_for over files)
    sks::pool().ensure(nd, false);
    std::vector<kmer_set> out((size_t)num_files);
    sks::run_on_devices(nd, [&](int r) {
        int64_t b = 0, e = 0;
        sks_shard_range(num_files, r, nd, &b, &e);
        if (e <= b) return;
        std::vector<kmer_set> part =
            kmer_sets_on_this_device((int)(e - b), fasta_filenames + b, mask, window_length, plan, sketching_cond);
        for (int64_t i = b; i < e; ++i) out[(size_t)i] = std::move(part[(size_t)(i - b)]);
    });
    return out;
}

namespace
{
std::vector<kmer_set> kmer_sets_on_this_device(const int num_files, char *fasta_filenames[], const kmer_bitset &mask,
                                               const int window_length, const sks::pred_plan &plan,
                                               const std::function<bool(const kmer)> &sketching_cond)
{
    std::vector<kmer_set> out((size_t)std::max(num_files, 0));
    if (num_files <= 0) return out;
    uint64_t m[2];
    sks::mask_words(mask, m);

    // host: read the files concurrently (the reference's cilk_for over files, src/kmer_set.cpp:124-131; its
    // exit(1) on an unreadable file is kept).  Parsing, splitting at non-ACGT bytes and 2-bit packing happen
    // on the device (sks_batch_from_fasta_text), at most 2 GiB of text per batch.
    std::vector<std::string> texts((size_t)num_files);
    std::vector<char> unreadable((size_t)num_files, 0);
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const int n_threads = (int)std::min<unsigned>(hw, (unsigned)num_files);
        std::atomic<int> next{0};
        auto work = [&]() {
            for (int i = next.fetch_add(1); i < num_files; i = next.fetch_add(1))
            {
                FILE *fp = fopen(fasta_filenames[i], "rb");
                if (!fp)
                {
                    unreadable[(size_t)i] = 1;
                    continue;
                }
                std::string &t = texts[(size_t)i];
                if (fseek(fp, 0, SEEK_END) == 0)
                {
                    const long sz = ftell(fp);
                    if (sz > 0) t.reserve((size_t)sz);
                    rewind(fp);
                }
                char buf[1 << 16];
                size_t got;
                while ((got = fread(buf, 1, sizeof(buf), fp)) > 0) t.append(buf, got);
                fclose(fp);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
        work();
        for (std::thread &t : pool) t.join();
    }
    for (int i = 0; i < num_files; ++i)
        if (unreadable[(size_t)i])
        {
            std::cerr << "Unable to open " << fasta_filenames[i] << ". \n Exiting..." << std::endl;
            exit(1);
        }

    const uint64_t batch_limit = (1ull << 31) - (1ull << 20);
    for (int first = 0; first < num_files;)
    {
        int last = first;
        uint64_t bytes = 0;
        std::vector<const char *> ptr;
        std::vector<uint64_t> len;
        while (last < num_files && (last == first || bytes + texts[(size_t)last].size() <= batch_limit))
        {
            ptr.push_back(texts[(size_t)last].data());
            len.push_back(texts[(size_t)last].size());
            bytes += texts[(size_t)last].size();
            ++last;
        }
        const int n = last - first;
        sks::batch_guard bg;
        if (bytes <= batch_limit)
        {
            check(sks_batch_from_fasta_text(ctx(), n, ptr.data(), len.data(), &bg.b), "sks_batch_from_fasta_text");
        }
        else
        { // a single file beyond the device parser's limit: parse and pack it on the host
            uint64_t nb = 0, ns = 0;
            check(sks_fasta_parse(ptr[0], (size_t)len[0], &nb, &ns, nullptr, nullptr), "sks_fasta_parse");
            std::vector<uint32_t> words(sks_packed_words(nb) + 1);
            std::vector<uint64_t> segs((size_t)ns + 1);
            check(sks_fasta_parse(ptr[0], (size_t)len[0], &nb, &ns, words.data(), segs.data()), "sks_fasta_parse");
            const uint32_t *pw = words.data();
            const uint64_t *ps = segs.data();
            check(sks_batch_upload(ctx(), 1, &pw, &nb, &ps, &ns, &bg.b), "sks_batch_upload");
            check(sks_ctx_sync(ctx()), "sks_ctx_sync");
        }
        if (plan.kind == sks::pred_plan::DEVICE)
        {
            std::vector<sks_set *> sets((size_t)n, nullptr);
            check(sks_sketch(ctx(), bg.b, m, window_length, &plan.pred, sks::g_ts.repr, sets.data()), "sks_sketch");
            for (int i = 0; i < n; ++i) sks::set_access::adopt(out[(size_t)(first + i)], sets[(size_t)i], window_length, mask);
        }
        else
        {
            // Opaque condition: the device slides, canonicalises and returns every k-mer; the callable runs here.
            for (int i = 0; i < n; ++i)
            {
                std::vector<kmer> kept;
                append_list(kept, bg.b, i, mask, window_length, plan, sketching_cond);
                std::vector<uint64_t> keys;
                keys.reserve(kept.size() * 2);
                for (const kmer &k : kept)
                {
                    keys.push_back(k.masked_bits.word(0));
                    keys.push_back(k.masked_bits.word(1));
                }
                sks_set *h = nullptr;
                check(sks_set_from_host_keys(ctx(), keys.data(), (int64_t)kept.size(), m, window_length, &h),
                      "sks_set_from_host_keys");
                sks::set_access::adopt(out[(size_t)(first + i)], h, window_length, mask);
            }
        }
        check(sks_ctx_sync(ctx()), "sks_ctx_sync"); // the host text buffers of this batch may go away
        first = last;
    }
    return out;
}
} // namespace

std::vector<kmer_set> parallel_kmer_sets_from_fasta_files(const int num_files, char *fasta_filenames[],
                                                          const kmer_bitset &mask, const int window_length,
                                                          const std::function<bool(const kmer)> &sketching_cond)
{
    return kmer_sets_from_fasta_files(num_files, fasta_filenames, mask, window_length, sketching_cond);
}

kmer_set kmer_set_from_fasta_file(const char fasta_filename[], const kmer_bitset &mask, const int window_length,
                                  const std::function<bool(const kmer)> &sketching_cond)
{
    char *names[1] = {const_cast<char *>(fasta_filename)};
    std::vector<kmer_set> v = kmer_sets_from_fasta_files(1, names, mask, window_length, sketching_cond);
    return std::move(v[0]);
}

namespace sks
{
// generate_all_pairs_from_vector (src/generators.hpp:44-58) over sets that were sketched in contiguous blocks on
// several devices (kmer_sets_from_fasta_files with sks::set_devices(n)): the sharded all-vs-all of the C ABI, one
// worker thread per device.  Returns false when the lists are anything else (the caller then takes the general route).
bool sharded_all_pairs(const std::vector<kmer_set *> &v1, const std::vector<kmer_set *> &v2, std::vector<int> &out)
{
    const size_t len = v1.size();
    size_t n = 0;
    while (n * n < len) ++n;
    if (n < 2 || n * n != len) return false;
    std::vector<sks_set *> sets(n, nullptr);
    int nd = 0;
    for (size_t i = 0; i < n; ++i)
    {
        kmer_set *s = v1[i * n];
        if (!s || s != v2[i]) return false;
        device_set *d = set_access::device(*s);
        if (!d || sks_set_repr(d->h) != SKS_REPR_SORTED) return false;
        sets[i] = d->h;
        nd = std::max(nd, sks_set_device_index(d->h) + 1);
    }
    if (nd < 2) return false;
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 0; j < n; ++j)
            if (v1[i * n + j] != v1[i * n] || v2[i * n + j] != v2[j]) return false;
    for (int r = 0; r < nd; ++r)
    { // the blocks must be the ones sks_shard_range assigns
        int64_t b = 0, e = 0;
        sks_shard_range((int64_t)n, r, nd, &b, &e);
        for (int64_t i = b; i < e; ++i)
            if (sks_set_device_index(sets[(size_t)i]) != r) return false;
    }
    pool().ensure(nd, true);
    run_on_devices(nd, [&](int r) {
        int64_t b = 0, e = 0;
        sks_shard_range((int64_t)n, r, nd, &b, &e);
        std::vector<int32_t> rows((size_t)std::max<int64_t>(e - b, 1) * n);
        check(sks_all_vs_all_sharded(pool().ctxs[(size_t)r], pool().comms[(size_t)r], sets.data() + b, e - b, (int64_t)n,
                                     rows.data(), nullptr, nullptr),
              "sks_all_vs_all_sharded");
        for (int64_t i = b; i < e; ++i)
            for (size_t j = 0; j < n; ++j) out[(size_t)i * n + j] = rows[(size_t)(i - b) * n + j];
    });
    return true;
}
} // namespace sks

// ---- intersections --------------------------------------------------------------------------------------------------------
std::vector<int> compute_pairwise_kmer_set_intersections(const std::vector<kmer_set *> &kmer_sets_1,
                                                         const std::vector<kmer_set *> &kmer_sets_2)
{
    if (kmer_sets_1.size() != kmer_sets_2.size())
        throw std::runtime_error("Lists of kmer sets for intersection computation have different lengths");
    const size_t n = kmer_sets_1.size();
    std::vector<int> out(n, 0);
    if (sks::sharded_all_pairs(kmer_sets_1, kmer_sets_2, out)) return out;
    // pairs the device can take in one launch: both sides resident, same mask and representation
    std::vector<sks_set *> a, b;
    std::vector<size_t> where;
    std::vector<std::shared_ptr<sks::device_set>> keep_alive;
    std::map<sks_set *, sks_set *> here; // sets of other devices -> their copies on this thread's device
    auto local_handle = [&](sks::device_set *d) {
        const int dev = sks_set_device_index(d->h);
        (void)ctx();
        if (dev == sks::g_ts.device) return d->h;
        auto it = here.find(d->h);
        if (it != here.end()) return it->second;
        sks_set *copy = nullptr;
        check(sks_set_clone_to(ctx(), d->h, &copy), "sks_set_clone_to");
        keep_alive.push_back(std::make_shared<sks::device_set>(copy, d->window, d->mask));
        here[d->h] = copy;
        return copy;
    };
    for (size_t i = 0; i < n; ++i)
    {
        sks::device_set *da = sks::set_access::device(*kmer_sets_1[i]);
        sks::device_set *db = sks::set_access::device(*kmer_sets_2[i]);
        if (!da || !db) continue;                 // an empty set intersects nothing
        if (!(da->mask == db->mask)) continue;    // k-mers under different masks are never equal (src/kmer.hpp:82-85)
        sks_set *ha = local_handle(da), *hb = local_handle(db);
        if (sks_set_repr(ha) != sks_set_repr(hb))
        {
            // mixed representations (bitset vs sorted keys): re-key the bitset side as sorted keys
            sks::device_set *bit = sks_set_repr(ha) == SKS_REPR_BITSET ? da : db;
            int64_t cnt = 0;
            check(sks_set_size(ctx(), bit->h, &cnt), "sks_set_size");
            std::vector<uint64_t> keys((size_t)cnt * 2);
            check(sks_set_keys(ctx(), bit->h, keys.data(), (uint64_t)cnt), "sks_set_keys");
            uint64_t m[2];
            sks::mask_words(bit->mask, m);
            sks_set *conv = nullptr;
            check(sks_set_from_host_keys(ctx(), keys.data(), cnt, m, bit->window, &conv), "sks_set_from_host_keys");
            keep_alive.push_back(std::make_shared<sks::device_set>(conv, bit->window, bit->mask));
            (bit == da ? ha : hb) = conv;
        }
        a.push_back(ha);
        b.push_back(hb);
        where.push_back(i);
    }
    if (!a.empty())
    {
        std::vector<int32_t> r(a.size());
        check(sks_intersect_pairs(ctx(), a.data(), (int64_t)a.size(), b.data(), (int64_t)b.size(), r.data()),
              "sks_intersect_pairs");
        for (size_t k = 0; k < where.size(); ++k) out[where[k]] = r[k];
    }
    return out;
}

std::vector<int> parallel_compute_pairwise_kmer_set_intersections(const std::vector<kmer_set *> &kmer_sets_1,
                                                                  const std::vector<kmer_set *> &kmer_sets_2)
{
    return compute_pairwise_kmer_set_intersections(kmer_sets_1, kmer_sets_2);
}

int kmer_set_intersection(const kmer_set &ks1, const kmer_set &ks2)
{
    std::vector<kmer_set *> a{const_cast<kmer_set *>(&ks1)}, b{const_cast<kmer_set *>(&ks2)};
    return compute_pairwise_kmer_set_intersections(a, b)[0];
}
