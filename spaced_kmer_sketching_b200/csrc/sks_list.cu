// Ordered k-mer list: the std::vector<kmer> of nucleotide_string_list_to_kmers
// (src/kmer_sliding.cpp:199-238), in sequence order with duplicates, including the kmer_bits field
// with its history bits (src/kmer_sliding.cpp:28 never truncates the forward window).
//
// The sketch kernel (OUT_LIST) emits (masked_bits, strand<<31 | start) in CTA-local order; here the
// entries are put back into sequence order (radix sort on the start position) and kmer_bits is
// rebuilt from the packed bases:
//   forward strand : group g (bits 2g, 2g+1) = s[p + w - 1 - g] while that base is inside the
//                    window's segment, g = 0..63; 0 before the segment start
//   reverse strand : group g = 3 - s[p + g] for g < w, 0 above (src/kmer_sliding.cpp:44-46)
#include <cub/cub.cuh>

#include "sks_internal.cuh"

namespace sks {
namespace {

__global__ void iota_kernel(uint32_t *out, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

__device__ __forceinline__ uint32_t base_at(const uint32_t *words, uint32_t i) {
  return (words[i >> 4] >> (2 * (i & 15))) & 3u;
}

template <int KW>
__global__ void __launch_bounds__(256)
    list_finalize_kernel(const uint32_t *__restrict__ words, const uint32_t *__restrict__ seg_end, uint32_t n_segs,
                         int window, const unsigned long long *__restrict__ raw_keys,
                         const uint32_t *__restrict__ pos_sorted, const uint32_t *__restrict__ idx_sorted, uint32_t n,
                         unsigned long long *__restrict__ out_masked, unsigned long long *__restrict__ out_bits) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t src = idx_sorted[i];
  const uint32_t ps = pos_sorted[i];
  const uint32_t p = ps & 0x7FFFFFFFu;
  const bool rc = (ps >> 31) != 0;
  out_masked[2 * i] = raw_keys[(size_t)KW * src];
  out_masked[2 * i + 1] = KW == 2 ? raw_keys[(size_t)KW * src + 1] : 0ull;
  unsigned long long lo = 0, hi = 0;
  if (rc) {
    for (int g = 0; g < window; ++g) {
      const unsigned long long v = 3u - base_at(words, p + g);
      if (g < 32) lo |= v << (2 * g); else hi |= v << (2 * (g - 32));
    }
  } else {
    // start of the segment that holds p: the end of the previous segment
    uint32_t a = 0, b = n_segs - 1;
    while (a < b) {
      const uint32_t mid = (a + b) >> 1;
      if (seg_end[mid] > p) b = mid; else a = mid + 1;
    }
    const uint32_t seg_start = a == 0 ? 0u : seg_end[a - 1];
    const uint32_t last = p + window - 1;
    for (int g = 0; g < 64; ++g) {
      if (last < seg_start + g) break;
      const unsigned long long v = base_at(words, last - g);
      if (g < 32) lo |= v << (2 * g); else hi |= v << (2 * (g - 32));
    }
  }
  out_bits[2 * i] = lo;
  out_bits[2 * i + 1] = hi;
}

}  // namespace

// raw_keys / raw_pos: n entries as written by the OUT_LIST sketch kernel for ONE genome whose data
// words start at `words` and whose segment ends are seg_end[0..n_segs).  Writes n entries of
// (masked lo, hi) and (kmer_bits lo, hi) to the DEVICE buffers out_masked / out_bits.
int launch_list_finalize(sks_ctx *ctx, const uint32_t *words, const uint32_t *seg_end, uint32_t n_segs, int window,
                         int key_words, const void *raw_keys, const uint32_t *raw_pos, uint32_t n,
                         unsigned long long *out_masked, unsigned long long *out_bits) {
  if (n == 0) return SKS_OK;
  size_t cub_bytes = 0;
  {
    uint32_t *k = nullptr;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, k, k, k, k, (int)n, 0, 31, ctx->stream);
  }
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t sz = align((size_t)n * 4);
  char *base = nullptr;
  SKS_TRY(ctx_scratch(ctx, 3 * sz + align(cub_bytes), reinterpret_cast<void **>(&base)));
  uint32_t *idx = reinterpret_cast<uint32_t *>(base);
  uint32_t *pos_sorted = reinterpret_cast<uint32_t *>(base + sz);
  uint32_t *idx_sorted = reinterpret_cast<uint32_t *>(base + 2 * sz);
  void *d_cub = base + 3 * sz;
  const unsigned nblk = (n + 255) / 256;
  KernelTimer timer(ctx, SKS_KERNEL_LIST);
  iota_kernel<<<nblk, 256, 0, ctx->stream>>>(idx, n);
  // start positions are distinct, so sorting the low 31 bits orders the list; bit 31 is the strand
  cub::DeviceRadixSort::SortPairs(d_cub, cub_bytes, raw_pos, pos_sorted, idx, idx_sorted, (int)n, 0, 31, ctx->stream);
  if (key_words == 1)
    list_finalize_kernel<1><<<nblk, 256, 0, ctx->stream>>>(words, seg_end, n_segs, window,
                                                           static_cast<const unsigned long long *>(raw_keys),
                                                           pos_sorted, idx_sorted, n, out_masked, out_bits);
  else
    list_finalize_kernel<2><<<nblk, 256, 0, ctx->stream>>>(words, seg_end, n_segs, window,
                                                           static_cast<const unsigned long long *>(raw_keys),
                                                           pos_sorted, idx_sorted, n, out_masked, out_bits);
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches += 6;
  return SKS_OK;
}

}  // namespace sks
