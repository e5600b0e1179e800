// Synthetic genomes generated directly in HBM, 2-bit packed (benchmark inputs).
// The generator is the one SURVEY.md 4.2 (KAT-3) defines -- the reference ships none:
//   splitmix64 stream; base_i = next() >> 62; mutate: u = next(); if (u % D == 0)
//   code = (code + 1 + ((u >> 32) % 3)) & 3.  The i-th output depends only on seed + i*gamma, so
//   every word is generated independently.  oracle/oracle.c (orc_gen / orc_mutate) is the CPU twin.
#include "sks_internal.cuh"

namespace sks {
namespace {

__device__ __forceinline__ uint64_t splitmix64_at(uint64_t seed, uint64_t i) {
  uint64_t z = seed + i * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
    synth_kernel(uint32_t *__restrict__ words, const GenomeDesc *__restrict__ genomes,
                 const uint64_t *__restrict__ gen_seed, const uint64_t *__restrict__ mut_seed,
                 const uint64_t *__restrict__ mut_D, const uint64_t *__restrict__ first_base) {
  const GenomeDesc gd = genomes[blockIdx.y];
  const uint64_t first = first_base[blockIdx.y];
  const uint64_t gs = gen_seed[blockIdx.y], ms = mut_seed[blockIdx.y], D = mut_D[blockIdx.y];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < gd.n_words; wi += stride) {
    uint32_t word = 0;
#pragma unroll 4
    for (int b = 0; b < 16; ++b) {
      const uint64_t i = (uint64_t)wi * 16 + b;
      if (i < gd.n_bases) {
        uint32_t code = (uint32_t)(splitmix64_at(gs, first + i + 1) >> 62);
        if (D != 0) {
          const uint64_t u = splitmix64_at(ms, first + i + 1);
          if (u % D == 0) code = (code + 1 + (uint32_t)((u >> 32) % 3)) & 3;
        }
        word |= code << (2 * b);
      }
    }
    words[gd.word_off + wi] = word;
  }
}

}  // namespace

int launch_synth(sks_ctx *ctx, uint32_t *words, const GenomeDesc *genomes, int n_genomes, uint32_t max_words,
                 const uint64_t *gen_seed, const uint64_t *mut_seed, const uint64_t *mut_D, const uint64_t *first_base) {
  if (n_genomes == 0 || max_words == 0) return SKS_OK;
  unsigned gx = (max_words + 255) / 256;
  if (gx > (unsigned)ctx->sm_count * 8) gx = (unsigned)ctx->sm_count * 8;
  KernelTimer timer(ctx, SKS_KERNEL_SYNTH);
  for (int g0 = 0; g0 < n_genomes; g0 += 65535) {
    const int ng = n_genomes - g0 < 65535 ? n_genomes - g0 : 65535;
    synth_kernel<<<dim3(gx, ng), 256, 0, ctx->stream>>>(words, genomes + g0, gen_seed + g0, mut_seed + g0, mut_D + g0, first_base + g0);
    SKS_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
  }
  return SKS_OK;
}

}  // namespace sks
