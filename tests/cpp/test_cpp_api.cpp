// Exercises the reference-shaped C++ API (include/kmer.hpp ...) exactly as a user of the reference
// would, and prints one JSON object; tests/test_gpu_cpp_api.py compares it with the CPU oracle.
//   test_cpp_api A.fna B.fna
#include <iomanip>
#include <iostream>
#include <sstream>
#include <thread>

#include "ani_estimator.hpp"
#include "fasta_processing.hpp"
#include "generators.hpp"
#include "kmer.hpp"

static std::string hex128(const kmer_bitset &b)
{
    std::ostringstream os;
    os << std::hex << std::setfill('0') << std::setw(16) << b.word(1) << std::setw(16) << b.word(0);
    return os.str();
}

static frac_min_hash fmh(1);
static bool driver_condition(const kmer &k) { return fmh(k) % 200 == 0; } // src/kmer-sketching.cpp:29-34

int main(int argc, char *argv[])
{
    if (argc < 3) return 2;
    initialise_contiguous_kmer_array();
    initialise_reversing_kmer_array();
    std::cout << std::setprecision(17) << "{";

    // masks
    std::cout << "\"mask_24_16\":\"" << hex128(generate_random_spaced_seed_mask(24, 16)) << "\",";
    std::cout << "\"mask_31_21_s3\":\"" << hex128(generate_random_spaced_seed_mask(31, 21, 3)) << "\",";
    std::cout << "\"contig_33\":\"" << hex128(contiguous_kmer(33)) << "\",";
    bool threw = false;
    try { contiguous_kmer(65); } catch (const std::runtime_error &) { threw = true; }
    std::cout << "\"contig_65_throws\":" << (threw ? "true" : "false") << ",";
    const kmer_bitset seed_mask = sks::seed_string_to_mask("11001011");
    std::cout << "\"seed_mask\":\"" << hex128(seed_mask) << "\",\"seed_weight\":" << seed_mask.count() / NUCLEOTIDE_BIT_SIZE << ",";
    { std::ostringstream os; os << seed_mask; std::cout << "\"seed_mask_printed_len\":" << os.str().size() << ","; }

    // ordered list on strings (segments): KAT-1 plus a second segment
    {
        std::vector<std::vector<uint8_t>> strings = {{0, 0, 0, 1, 2, 3, 0, 1, 2, 3, 3, 3}, {}, {3, 2, 1, 0, 0, 1, 2, 3, 1}};
        std::vector<kmer> ks = nucleotide_string_list_to_kmers(strings, seed_mask, 8, sks::all_kmers());
        std::cout << "\"list_masked\":[";
        for (size_t i = 0; i < ks.size(); ++i) std::cout << (i ? "," : "") << "\"" << hex128(ks[i].masked_bits) << "\"";
        std::cout << "],\"list_bits\":[";
        for (size_t i = 0; i < ks.size(); ++i) std::cout << (i ? "," : "") << "\"" << hex128(ks[i].kmer_bits) << "\"";
        std::cout << "],";
        // legacy canonicalisation agrees with the sliding version (src/kmers.cpp:31-35)
        bool legacy_ok = true;
        for (const kmer &k : ks) legacy_ok = legacy_ok && canonical_kmer(k).masked_bits == k.masked_bits;
        std::cout << "\"legacy_ok\":" << (legacy_ok ? "true" : "false") << ",";
    }

    // FASTA -> sets under four kinds of sketching condition
    const kmer_bitset mask = generate_random_spaced_seed_mask(24, 16);
    char *files[2] = {argv[1], argv[2]};
    struct cond_case { const char *name; std::function<bool(const kmer)> f; };
    const cond_case cases[] = {
        {"all", sks::all_kmers()},
        {"fmh_struct", sks::fmh_condition(1, 50)},
        {"driver", driver_condition},
        {"lambda_fmh7", [](const kmer k) { return frac_min_hash(-2)(k) % 7 == 0; }},
        {"parity", [](const kmer k) { return k.masked_bits.count() % 2 == 0; }},
    };
    // Opaque callables are evaluated on the host by default, exactly once per real k-mer (src/kmer_sliding.cpp:183);
    // probing is opt-in.  A counting callable shows both.
    {
        long calls = 0;
        std::function<bool(const kmer)> counting = [&calls](const kmer k) { ++calls; return frac_min_hash(1)(k) % 200 == 0; };
        kmer_set plain = kmer_set_from_fasta_file(argv[1], mask, 24, counting);
        const std::string path_off = sks::last_predicate_path();
        const long calls_off = calls;
        sks::enable_predicate_probe(true);
        calls = 0;
        kmer_set probed = kmer_set_from_fasta_file(argv[1], mask, 24, counting);
        const std::string path_on = sks::last_predicate_path();
        sks::enable_predicate_probe(false);
        std::cout << "\"counting\":{\"path_default\":\"" << path_off << "\",\"calls_default\":" << calls_off
                  << ",\"size_default\":" << plain.kmer_set_size() << ",\"path_probe\":\"" << path_on << "\",\"calls_probe\":" << calls
                  << ",\"size_probe\":" << probed.kmer_set_size() << ",\"same\":" << (kmer_set_intersection(plain, probed) == plain.kmer_set_size() ? "true" : "false")
                  << "},";
    }
    std::cout << "\"cases\":{";
    bool first_case = true;
    for (int pass = 0; pass < 2; ++pass)
    for (const cond_case &c : cases)
    {
        const bool opaque = std::string(c.name) == "driver" || std::string(c.name) == "lambda_fmh7";
        if (pass == 1 && !opaque) continue;
        sks::enable_predicate_probe(pass == 1);
        std::vector<kmer_set> sets = parallel_kmer_sets_from_fasta_files(2, files, mask, 24, c.f);
        sks::enable_predicate_probe(false);
        const std::string path = sks::last_predicate_path();
        std::vector<kmer_set *> ptrs = {&sets[0], &sets[1]};
        const auto pairs = generate_all_pairs_from_vector(ptrs);
        const std::vector<int> inter = compute_pairwise_kmer_set_intersections(pairs.first, pairs.second);
        std::cout << (first_case ? "" : ",") << "\"" << c.name << (pass ? "_probed" : "") << "\":{\"path\":\"" << path << "\",\"sizes\":["
                  << sets[0].kmer_set_size() << "," << sets[1].kmer_set_size() << "],\"inter\":[";
        for (size_t i = 0; i < inter.size(); ++i) std::cout << (i ? "," : "") << inter[i];
        std::cout << "],\"ani\":[";
        for (size_t i = 0; i < inter.size(); ++i)
            std::cout << (i ? "," : "") << binomial_estimator(containment(inter[i], pairs.first[i]->kmer_set_size()), 16);
        std::cout << "]}";
        first_case = false;
    }
    std::cout << "},";

    // host-side sets: insert_kmers on a host list, the lazily materialised table, mixed use
    {
        const kmer_bitset m8 = sks::seed_string_to_mask("110101101");
        std::vector<acgt_string> strings = nucleotide_strings_from_fasta_file(argv[1]);
        std::vector<kmer> list = nucleotide_string_list_to_kmers(strings, m8, 9, sks::all_kmers());
        kmer_set host_set;
        host_set.insert_kmers(list);
        kmer_set dev_set = kmer_set_from_fasta_file(argv[1], m8, 9, sks::all_kmers());
        size_t found = 0, walked = 0;
        for (const auto &kv : dev_set.kmer_hashes) { ++walked; found += host_set.kmer_hashes.count(kv.first); }
        std::cout << "\"host_list_len\":" << list.size() << ",\"host_set_size\":" << host_set.kmer_set_size()
                  << ",\"dev_set_size\":" << dev_set.kmer_set_size() << ",\"walked\":" << walked << ",\"found\":" << found
                  << ",\"host_dev_inter\":" << kmer_set_intersection(host_set, dev_set)
                  << ",\"n_strings\":" << strings.size() << ",";
        kmer_set other = kmer_set_from_fasta_file(argv[2], mask, 24, sks::all_kmers());
        std::cout << "\"different_mask_inter\":" << kmer_set_intersection(dev_set, other) << ",";
        kmer_set empty;
        std::cout << "\"empty_inter\":" << kmer_set_intersection(empty, dev_set) << ",\"empty_size\":" << empty.kmer_set_size() << ",";
    }
    // Concurrent callers on distinct outputs, as the reference's Cilk workers are (src/kmer_set.cpp:124-131,179-182):
    // every worker thread sketches into its own kmer_set and exits; the sets outlive the workers' implicit
    // contexts and are then intersected from other threads.
    {
        std::vector<kmer_set> sets(8);
        std::vector<std::thread> pool;
        for (int t = 0; t < 8; ++t)
            pool.emplace_back([&, t]() {
                if (t < 4) sets[(size_t)t] = kmer_set_from_fasta_file(files[t % 2], mask, 24, sks::all_kmers());
                else sets[(size_t)t] = kmer_set_from_fasta_file(files[t % 2], mask, 24, sks::fmh_condition(1, 50));
            });
        for (std::thread &t : pool) t.join();
        pool.clear();
        std::vector<int> inter(8, -1);
        for (int t = 0; t < 8; ++t)
            pool.emplace_back([&, t]() { inter[(size_t)t] = kmer_set_intersection(sets[(size_t)t], sets[(size_t)(t ^ 1)]); });
        for (std::thread &t : pool) t.join();
        std::cout << "\"threads\":{\"sizes\":[";
        for (int t = 0; t < 8; ++t) std::cout << (t ? "," : "") << sets[(size_t)t].kmer_set_size();
        std::cout << "],\"inter\":[";
        for (int t = 0; t < 8; ++t) std::cout << (t ? "," : "") << inter[(size_t)t];
        std::cout << "]},";
    }
    // Several GPUs in one process (sks::set_devices): files in contiguous blocks per device, the all-pairs list as
    // the sharded all-vs-all, any other list after peer copies -- all equal to the single-device results.
    {
        char *files4[5] = {argv[1], argv[2], argv[2], argv[1], argv[2]};
        auto run = [&](int nd, std::vector<int> &all, std::vector<int> &ring, std::vector<int> &sizes) {
            sks::set_devices(nd);
            std::vector<kmer_set> sets = parallel_kmer_sets_from_fasta_files(5, files4, mask, 24, sks::fmh_condition(1, 50));
            std::vector<kmer_set *> ptrs;
            for (kmer_set &s : sets) ptrs.push_back(&s);
            const auto pairs = generate_all_pairs_from_vector(ptrs);
            all = parallel_compute_pairwise_kmer_set_intersections(pairs.first, pairs.second);
            const auto rp = generate_pairwise_from_vector(ptrs);
            ring = compute_pairwise_kmer_set_intersections(rp.first, rp.second);
            for (kmer_set &s : sets) sizes.push_back(s.kmer_set_size());
            sks::set_devices(1);
        };
        sks::set_devices(2);
        const int nd = sks::devices();
        sks::set_devices(1);
        std::cout << "\"multi\":{\"devices\":" << nd;
        std::vector<int> all1, ring1, sizes1, alln, ringn, sizesn;
        run(1, all1, ring1, sizes1);
        if (nd >= 2) run(nd, alln, ringn, sizesn);
        auto dump = [&](const char *name, const std::vector<int> &v) {
            std::cout << ",\"" << name << "\":[";
            for (size_t i = 0; i < v.size(); ++i) std::cout << (i ? "," : "") << v[i];
            std::cout << "]";
        };
        dump("all_1", all1);
        dump("ring_1", ring1);
        dump("sizes_1", sizes1);
        dump("all_n", alln);
        dump("ring_n", ringn);
        dump("sizes_n", sizesn);
        std::cout << "},";
    }
    bool mismatch_threw = false;
    try
    {
        kmer_set a, b;
        std::vector<kmer_set *> one = {&a}, two = {&a, &b};
        compute_pairwise_kmer_set_intersections(one, two);
    }
    catch (const std::runtime_error &) { mismatch_threw = true; }
    std::cout << "\"mismatch_throws\":" << (mismatch_threw ? "true" : "false") << ",";
    std::cout << "\"containment_0\":" << containment(0, 10) << ",\"estimator_neg\":" << binomial_estimator(-1.0, 5) << "}" << std::endl;
    return 0;
}
