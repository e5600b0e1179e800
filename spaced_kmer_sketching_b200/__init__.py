"""B200-native spaced k-mer sketching + ANI engine (drop-in for the sketch-and-compare path of
bensonlzl/spaced-kmer-sketching).  CUDA kernels and the C ABI live in csrc/ -> libsks.so; the
reference-shaped C++ API is in include/; this package is the Python host mirror."""
from ._lib import (HASH_BOOST_171, HASH_BOOST_181, LIB_PATH, PRED_ALL, PRED_FMH, PROTOTYPES, REPR_AUTO, REPR_BITSET, REPR_BITSET_ONCHIP,
                   REPR_SORTED, SksError, load)
from .engine import (Batch, Comm, Context, KmerSet, Predicate, all_kmers, ani_from_counts, binomial_estimator,
                     compute_pairwise_kmer_set_intersections, containment, contiguous_kmer, fasta_parse,
                     fasta_parse_file, fmh_hash, frac_min_hash, generate_all_pairs_from_vector,
                     generate_pairwise_from_vector, generate_random_spaced_seed_mask, kmer_set_from_fasta_file,
                     kmer_set_intersection, kmer_sets_from_fasta_files, mask_weight, pack_codes, reverse_kmer_bitset,
                     seed_to_mask, shard_range, comm_unique_id, unpack_codes)
