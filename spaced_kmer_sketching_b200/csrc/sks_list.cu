// Ordered k-mer list (the std::vector<kmer> of nucleotide_string_list_to_kmers,
// src/kmer_sliding.cpp:224-238).  Placeholder translation unit: implemented below.
#include "sks_internal.cuh"
extern "C" int sks_kmer_list(sks_ctx *ctx, const sks_batch *batch, int genome, const uint64_t mask[2], int window,
                             const sks_pred *pred, uint64_t *out_n, uint64_t *out_masked, uint64_t *out_bits,
                             uint64_t capacity) {
  (void)ctx; (void)batch; (void)genome; (void)mask; (void)window; (void)pred; (void)out_n; (void)out_masked;
  (void)out_bits; (void)capacity;
  return sks::set_error(SKS_ERR_INVALID, "sks_kmer_list: not implemented yet");
}
