// sks_cli -- the experiment driver of the sketch-and-compare path, with the reference driver's
// command line, sweep, timing lines and CSV layout (src/kmer-sketching.cpp:46-81,151-240):
//
//     sks_cli OUT.csv GENOME1.fna GENOME2.fna ...
//
// For each of the 62 (window, k) configurations -- contiguous k = 10..40, then random spaced seeds of
// weight k in a window of k + 10 for k = 10..40 -- every genome is sketched with the FracMinHash
// condition frac_min_hash(1)(kmer) % 200 == 0, all n*n ordered pairs are intersected, and
// ANI = containment(|A n B|, |A|)^(1/weight) is appended to the CSV.
// Written against include/kmer.hpp only, so it also builds against the reference's own headers (where the sketching
// condition is the free function below; on this repository's headers it is the recognised functor
// sks::fmh_condition, which runs on the device without probing).
#include <chrono>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "ani_estimator.hpp"
#include "generators.hpp"
#include "kmer.hpp"

namespace
{
const frac_min_hash sketch_hash(1);
const int sketch_modulus = 200;
[[maybe_unused]] bool keep_kmer(const kmer &k) { return sketch_hash(k) % sketch_modulus == 0; }

using clock_type = std::chrono::high_resolution_clock;
double ms_between(clock_type::time_point a, clock_type::time_point b)
{
    return std::chrono::duration<double, std::milli>(b - a).count();
}

struct csv_sink
{
    std::string path;
    bool started = false;
    void append(const std::vector<std::string> &first, const std::vector<std::string> &second,
                const std::vector<double> &values, int window, const kmer_bitset &mask)
    {
        std::ofstream out(path, started ? std::ios_base::app : std::ios_base::out);
        if (!out.is_open())
        {
            std::cerr << "Error: Unable to open file " << path << " for writing." << std::endl;
            return;
        }
        if (!started) out << "File 1,File 2,Estimated Value,Window Size,Mask" << std::endl;
        started = true;
        for (size_t i = 0; i < values.size() && i < first.size() && i < second.size(); ++i)
            out << first[i] << "," << second[i] << "," << values[i] << "," << window << "," << mask << std::endl;
    }
};

void run_configuration(int window, int weight_wanted, int n_files, char *files[], csv_sink &csv)
{
    const kmer_bitset mask = generate_random_spaced_seed_mask(window, weight_wanted);
    const int weight = static_cast<int>(mask.count() / NUCLEOTIDE_BIT_SIZE);

    const auto t0 = clock_type::now();
#ifdef SKS_KMER_HPP
    const std::function<bool(const kmer)> condition = sks::fmh_condition(1, sketch_modulus);
#else
    const std::function<bool(const kmer)> condition = keep_kmer;
#endif
    std::vector<kmer_set> sets = parallel_kmer_sets_from_fasta_files(n_files, files, mask, window, condition);
    const auto t1 = clock_type::now();
    std::cout << "Time taken for sketching = " << ms_between(t0, t1) << " ms" << std::endl;

    std::vector<kmer_set *> handles;
    std::vector<std::string> names;
    for (size_t i = 0; i < sets.size(); ++i)
    {
        handles.push_back(&sets[i]);
        names.push_back(files[i]);
    }
    const auto set_pairs = generate_all_pairs_from_vector(handles);
    const auto name_pairs = generate_all_pairs_from_vector(names);
    const std::vector<int> shared = parallel_compute_pairwise_kmer_set_intersections(set_pairs.first, set_pairs.second);

    std::vector<double> ani(shared.size());
    for (size_t i = 0; i < shared.size(); ++i)
        ani[i] = binomial_estimator(containment(shared[i], set_pairs.first[i]->kmer_set_size()), weight);
    const auto t2 = clock_type::now();
    std::cout << "Time taken for comparison = " << ms_between(t1, t2) << " ms" << std::endl;
    csv.append(name_pairs.first, name_pairs.second, ani, window, mask);
}
} // namespace

int main(int argc, char *argv[])
{
    if (argc < 2)
    {
        std::cerr << "usage: " << argv[0] << " OUT.csv GENOME.fna [GENOME.fna ...]" << std::endl;
        return 2;
    }
    initialise_contiguous_kmer_array();
    initialise_reversing_kmer_array();
    csv_sink csv{argv[1]};
    const int n_files = argc - 2;
    for (int k = 10; k <= 40; ++k) run_configuration(k, k, n_files, argv + 2, csv);       // contiguous
    for (int k = 10; k <= 40; ++k) run_configuration(k + 10, k, n_files, argv + 2, csv);  // spaced
    return 0;
}
