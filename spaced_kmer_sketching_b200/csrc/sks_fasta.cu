// Device-side FASTA ingest: raw file bytes in HBM -> 2-bit packed bases + segment tables (a batch).
//
// Reproduces strings_from_fasta + cut_nucleotide_strings (src/fasta_processing.cpp:79-211) as data-parallel
// passes instead of a getline loop.  The reference's rules, per file:
//   * lines are what std::getline yields ('\n' stripped, '\r' kept); a line starting with '>' is a header,
//     an empty line is blank, anything else is content;
//   * `name` is set by a header to the rest of the line (so ">" alone makes it empty), cleared by a content
//     line that holds a space, and unchanged by blank lines;
//   * content lines are collected while `name` is non-empty; a header or a blank line pushes what was
//     collected as one string, a content line with a space DISCARDS what was collected since the last push;
//   * every string is cut at every non-ACGT byte (case-insensitive) into the ACGT runs we call segments.
// Formulation: (1) line table; (2) a forward "last event wins" scan over lines gives `name` before each line;
// (3) a backward "nearest terminator" scan tells a collected line whether its string is pushed or discarded;
// (4) per byte: keep flag + break counter, two prefix sums, scatter of 1-byte codes; (5) segment ends from
// the break counters; (6) 2-bit packing into the batch layout.
#include <algorithm>

#include <cub/cub.cuh>

#include "sks_internal.cuh"

namespace sks {
namespace {

constexpr int kT = 256;

enum LineKind : uint8_t { LINE_HEADER = 0, LINE_BLANK = 1, LINE_CONTENT = 2 };
constexpr uint8_t kFirstOfFile = 0x80, kKindMask = 0x7F;

__device__ __forceinline__ uint32_t code_of(unsigned char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
  }
}

// index of the last element of a[0..n) that is <= x (a ascending, a[0] <= x)
__device__ __forceinline__ uint32_t upper_index(const uint32_t *__restrict__ a, uint32_t n, uint32_t x) {
  uint32_t lo = 0, hi = n;  // a[lo] <= x < a[hi]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// flag[i] = 1 when a line starts at byte i (file start or after '\n', and not at the file's end)
__global__ void line_start_flags(const unsigned char *__restrict__ text, const uint32_t *__restrict__ file_off,
                                 uint32_t n_files, uint32_t n, uint8_t *__restrict__ flag) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t f = upper_index(file_off, n_files + 1, i);
  flag[i] = (i == file_off[f] || text[i - 1] == '\n') ? 1 : 0;
}

struct Iota {
  __host__ __device__ uint32_t operator()(uint32_t i) const { return i; }
};
struct Widen {
  __host__ __device__ uint32_t operator()(uint8_t v) const { return v; }
};
using WideFlags = cub::TransformInputIterator<uint32_t, Widen, const uint8_t *>;

// Per line: kind, header value, first-of-file; length = up to the '\n' or the file end.
__global__ void line_info_kernel(const unsigned char *__restrict__ text, const uint32_t *__restrict__ file_off,
                                 uint32_t n_files, const uint32_t *__restrict__ line_off, uint32_t n_lines,
                                 uint8_t *__restrict__ kind, uint8_t *__restrict__ event, uint32_t *__restrict__ line_len) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lines) return;
  const uint32_t start = line_off[l];
  const uint32_t f = upper_index(file_off, n_files + 1, start);
  const uint32_t fend = file_off[f + 1];
  uint32_t end = (l + 1 < n_lines && line_off[l + 1] <= fend) ? line_off[l + 1] : fend;  // one past the line incl. '\n'
  if (end > start && text[end - 1] == '\n') --end;
  const uint32_t len = end - start;
  line_len[l] = len;
  uint8_t k = LINE_CONTENT, ev = 0;  // event: 0 none, 2 name := empty, 3 name := non-empty
  if (len == 0) {
    k = LINE_BLANK;
  } else if (text[start] == '>') {
    k = LINE_HEADER;
    ev = len > 1 ? 3 : 2;
  }
  const bool first = start == file_off[f];
  if (first && ev == 0) ev = 2;  // a new file starts with an empty name
  kind[l] = k | (first ? kFirstOfFile : 0);
  event[l] = ev;
}

// A space inside a content line clears `name` (and discards the collected record).
__global__ void space_mark_kernel(const unsigned char *__restrict__ text, uint32_t n, const uint32_t *__restrict__ line_off,
                                  uint32_t n_lines, const uint8_t *__restrict__ kind, uint8_t *__restrict__ has_space,
                                  uint8_t *__restrict__ event) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || text[i] != ' ') return;
  const uint32_t l = upper_index(line_off, n_lines, i);
  if ((kind[l] & kKindMask) == LINE_CONTENT) {
    has_space[l] = 1;
    event[l] = 2;
  }
}

struct LastEvent {  // forward scan: the most recent event wins
  __host__ __device__ uint8_t operator()(uint8_t a, uint8_t b) const { return b ? b : a; }
};

// acc = content line without a space, collected (name non-empty before it).  term = 0 for collected lines,
// else the terminator type: 2 = discards the collected record (content line with a space while name was
// non-empty), 1 = pushes it (everything else).
__global__ void line_collect_kernel(const uint8_t *__restrict__ kind, const uint8_t *__restrict__ has_space,
                                    const uint8_t *__restrict__ state_incl, uint32_t n_lines,
                                    uint8_t *__restrict__ term_rev) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lines) return;
  // name before this line = scan value after the previous line, unless this line is itself first of its file
  // (then its own forced event says "empty", which only matters for the lines after it)
  const bool first_of_file = (kind[l] & kFirstOfFile) != 0;
  const bool content = (kind[l] & kKindMask) == LINE_CONTENT;
  const bool name_before = (l > 0 && !first_of_file) ? state_incl[l - 1] == 3 : false;
  uint8_t t;
  if (content && !has_space[l] && name_before) t = 0;
  else if (content && has_space[l] && name_before) t = 2;
  else t = 1;
  term_rev[n_lines - 1 - l] = t;  // reversed, so that a forward scan finds the NEXT terminator
}

struct NearestTerm {  // on the reversed array: the first non-zero seen so far (= nearest terminator ahead)
  __host__ __device__ uint8_t operator()(uint8_t a, uint8_t b) const { return b ? b : a; }
};

// kept[l] = collected and its string is pushed; starts[l] = kept and the previous line is not kept
__global__ void line_keep_kernel(const uint8_t *__restrict__ term_rev, const uint8_t *__restrict__ next_rev, uint32_t n_lines,
                                 uint8_t *__restrict__ kept) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lines) return;
  const uint32_t r = n_lines - 1 - l;
  const bool acc = term_rev[r] == 0;
  // nearest terminator strictly after l: scan value at the reversed position before r
  const uint8_t nxt = r > 0 ? next_rev[r - 1] : 0;  // 0: end of input = push
  kept[l] = (acc && nxt != 2) ? 1 : 0;
}

// Per byte: keep = ACGT byte of a kept line; brk = an event that separates segments at this byte.
__global__ void byte_flags_kernel(const unsigned char *__restrict__ text, uint32_t n, const uint32_t *__restrict__ line_off,
                                  const uint32_t *__restrict__ line_len, uint32_t n_lines, const uint8_t *__restrict__ kept,
                                  uint8_t *__restrict__ keep, uint8_t *__restrict__ brk) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t k = 0, b = 0;
  if (n_lines > 0 && i >= line_off[0]) {
    const uint32_t l = upper_index(line_off, n_lines, i);
    const uint32_t pos = i - line_off[l];
    if (kept[l] && pos < line_len[l]) {
      const bool valid = code_of(text[i]) < 4;
      k = valid ? 1 : 0;
      b = (!valid || (pos == 0 && !(l > 0 && kept[l - 1]))) ? 1 : 0;
    }
  }
  keep[i] = k;
  brk[i] = b;
}

// codes[K[i]] = code, bval[K[i]] = inclusive break count, for kept bytes
__global__ void scatter_codes_kernel(const unsigned char *__restrict__ text, uint32_t n, const uint8_t *__restrict__ keep,
                                     const uint8_t *__restrict__ brk, const uint32_t *__restrict__ kpos,
                                     const uint32_t *__restrict__ bcount, uint8_t *__restrict__ codes,
                                     uint32_t *__restrict__ bval) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  const uint32_t k = kpos[i];
  codes[k] = (uint8_t)code_of(text[i]);
  bval[k] = bcount[i] + brk[i];  // bcount is the exclusive sum
}

// genome_base[g] = compacted index of the first kept byte at or after file_off[g]  (g = 0..n_files)
__global__ void genome_base_kernel(const uint32_t *__restrict__ file_off, uint32_t n_files, uint32_t n,
                                   const uint32_t *__restrict__ kpos, uint32_t total_kept, uint32_t *__restrict__ genome_base) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > n_files) return;
  genome_base[g] = file_off[g] < n ? kpos[file_off[g]] : total_kept;
}

// end_flag[k] = 1 when compacted base k is the last of its segment
__global__ void seg_end_flags_kernel(const uint32_t *__restrict__ bval, uint32_t total, const uint32_t *__restrict__ genome_base,
                                     uint32_t n_files, uint8_t *__restrict__ end_flag) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= total) return;
  bool last = k + 1 == total || bval[k + 1] != bval[k];
  if (!last) {  // genome boundary (a new file always starts with a break, so this is only a safety net)
    const uint32_t g = upper_index(genome_base, n_files + 1, k);
    last = genome_base[g + 1] == k + 1;
  }
  end_flag[k] = last ? 1 : 0;
}

struct PlusOne {
  __host__ __device__ uint32_t operator()(uint32_t k) const { return k + 1; }
};

// words of genome g: 16 codes per word, into the batch layout
__global__ void pack_kernel(const uint8_t *__restrict__ codes, const uint32_t *__restrict__ genome_base,
                            const GenomeDesc *__restrict__ genomes, uint32_t *__restrict__ words) {
  const GenomeDesc gd = genomes[blockIdx.y];
  const uint32_t base = genome_base[blockIdx.y];
  const uint32_t n_words = (gd.n_bases + 15) / 16;
  for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += gridDim.x * blockDim.x) {
    uint32_t v = 0;
    const uint32_t first = wi * 16, cnt = gd.n_bases - first < 16 ? gd.n_bases - first : 16;
    for (uint32_t b = 0; b < cnt; ++b) v |= (uint32_t)codes[base + first + b] << (2 * b);
    words[gd.word_off + wi] = v;
  }
}

template <typename T>
T *carve(char *&p, size_t count) {
  T *r = reinterpret_cast<T *>(p);
  p += (count * sizeof(T) + 255) & ~(size_t)255;
  return r;
}

}  // namespace

// Parses n_files FASTA texts that lie back to back in the DEVICE buffer d_text (file f = bytes
// [h_file_off[f], h_file_off[f+1])).  Produces, on the host, the number of bases and the segment lengths of
// every genome, and on the device the 1-byte codes of all genomes back to back (d_codes, owned by *codes_buf).
int fasta_parse_device(sks_ctx *ctx, const unsigned char *d_text, const std::vector<uint64_t> &h_file_off,
                       std::vector<uint64_t> *n_bases, std::vector<std::vector<uint64_t>> *seg_len, BufferRef *codes_buf,
                       BufferRef *genome_base_buf) {
  const uint32_t n_files = (uint32_t)h_file_off.size() - 1;
  const uint64_t n64 = h_file_off.back();
  n_bases->assign(n_files, 0);
  seg_len->assign(n_files, {});
  if (n64 >= (1ull << 31)) return set_error(SKS_ERR_CAPACITY, "FASTA batch of %llu bytes exceeds 2 GiB", (unsigned long long)n64);
  const uint32_t n = (uint32_t)n64;
  SKS_TRY(alloc_buffer(ctx, 4 * (size_t)(n_files + 1), genome_base_buf));
  uint32_t *d_genome_base = static_cast<uint32_t *>((*genome_base_buf)->ptr);
  if (n == 0) {
    SKS_TRY(alloc_buffer(ctx, 16, codes_buf));
    SKS_CUDA_TRY(cudaMemsetAsync(d_genome_base, 0, 4 * (size_t)(n_files + 1), ctx->stream));
    return SKS_OK;
  }
  KernelTimer timer(ctx, SKS_KERNEL_FASTA);
  cudaStream_t st = ctx->stream;
  const unsigned nb_bytes = (n + kT - 1) / kT;

  // ---- pass 1: line table ----------------------------------------------------------------------------
  size_t cub_a = 0, cub_b = 0, cub_c = 0;
  {
    uint8_t *f8 = nullptr;
    uint32_t *u = nullptr;
    cub::TransformInputIterator<uint32_t, Iota, cub::CountingInputIterator<uint32_t>> it(cub::CountingInputIterator<uint32_t>(0), Iota());
    cub::DeviceSelect::Flagged(nullptr, cub_a, it, f8, u, u, (int)n, st);
    cub::DeviceScan::ExclusiveSum(nullptr, cub_b, WideFlags(f8, Widen()), u, (int)n, st);
    cub::DeviceScan::InclusiveScan(nullptr, cub_c, f8, f8, LastEvent(), (int)n, st);
  }
  const size_t cub_bytes = std::max(cub_a, std::max(cub_b, cub_c)) + 256;
  // scratch for the byte-level passes (line-level arrays are carved after the line count is known)
  const size_t byte_scratch = 3 * (((size_t)n + 255) & ~(size_t)255) + 2 * (((size_t)n * 4 + 255) & ~(size_t)255) +
                              (((size_t)(n_files + 2) * 4 + 255) & ~(size_t)255) + 512 + cub_bytes;
  BufferRef tmp;
  SKS_TRY(alloc_buffer(ctx, byte_scratch, &tmp));
  char *p = static_cast<char *>(tmp->ptr);
  uint8_t *d_flag = carve<uint8_t>(p, n);     // line-start flags, later keep flags
  uint8_t *d_brk = carve<uint8_t>(p, n);
  uint8_t *d_endflag = carve<uint8_t>(p, n);
  uint32_t *d_u32a = carve<uint32_t>(p, n);   // line offsets, later kpos
  uint32_t *d_u32b = carve<uint32_t>(p, n);   // bcount
  uint32_t *d_file_off = carve<uint32_t>(p, n_files + 2);
  uint32_t *d_count = carve<uint32_t>(p, 64);
  void *d_cub = p;

  uint32_t *h_stage = nullptr;
  SKS_TRY(ctx_pinned(ctx, 4 * (size_t)(n_files + 2) + 64, reinterpret_cast<void **>(&h_stage)));
  for (uint32_t f = 0; f <= n_files; ++f) h_stage[f] = (uint32_t)h_file_off[f];
  SKS_CUDA_TRY(cudaMemcpyAsync(d_file_off, h_stage, 4 * (size_t)(n_files + 1), cudaMemcpyHostToDevice, st));

  line_start_flags<<<nb_bytes, kT, 0, st>>>(d_text, d_file_off, n_files, n, d_flag);
  {
    cub::TransformInputIterator<uint32_t, Iota, cub::CountingInputIterator<uint32_t>> it(cub::CountingInputIterator<uint32_t>(0), Iota());
    size_t tb = cub_bytes;
    cub::DeviceSelect::Flagged(d_cub, tb, it, d_flag, d_u32a, d_count, (int)n, st);
  }
  uint32_t *h_count = h_stage + n_files + 2;
  SKS_CUDA_TRY(cudaMemcpyAsync(h_count, d_count, 4, cudaMemcpyDeviceToHost, st));
  SKS_CUDA_TRY(cudaStreamSynchronize(st));
  const uint32_t n_lines = *h_count;

  // ---- pass 2/3: line-level scans ---------------------------------------------------------------------
  BufferRef ltmp;
  const size_t lsz = ((size_t)n_lines + 255) & ~(size_t)255;
  SKS_TRY(alloc_buffer(ctx, 8 * lsz + 2 * (((size_t)n_lines * 4 + 255) & ~(size_t)255) + 256, &ltmp));
  char *q = static_cast<char *>(ltmp->ptr);
  uint32_t *d_line_off = carve<uint32_t>(q, n_lines);
  uint32_t *d_line_len = carve<uint32_t>(q, n_lines);
  uint8_t *d_kind = carve<uint8_t>(q, n_lines);
  uint8_t *d_event = carve<uint8_t>(q, n_lines);
  uint8_t *d_space = carve<uint8_t>(q, n_lines);
  uint8_t *d_state = carve<uint8_t>(q, n_lines);
  uint8_t *d_term = carve<uint8_t>(q, n_lines);
  uint8_t *d_next = carve<uint8_t>(q, n_lines);
  uint8_t *d_kept = carve<uint8_t>(q, n_lines);
  if (n_lines > 0) {
    const unsigned nb_lines = (n_lines + kT - 1) / kT;
    SKS_CUDA_TRY(cudaMemcpyAsync(d_line_off, d_u32a, 4 * (size_t)n_lines, cudaMemcpyDeviceToDevice, st));
    SKS_CUDA_TRY(cudaMemsetAsync(d_space, 0, n_lines, st));
    line_info_kernel<<<nb_lines, kT, 0, st>>>(d_text, d_file_off, n_files, d_line_off, n_lines, d_kind, d_event, d_line_len);
    space_mark_kernel<<<nb_bytes, kT, 0, st>>>(d_text, n, d_line_off, n_lines, d_kind, d_space, d_event);
    size_t tb = cub_bytes;
    cub::DeviceScan::InclusiveScan(d_cub, tb, d_event, d_state, LastEvent(), (int)n_lines, st);
    line_collect_kernel<<<nb_lines, kT, 0, st>>>(d_kind, d_space, d_state, n_lines, d_term);
    tb = cub_bytes;
    cub::DeviceScan::InclusiveScan(d_cub, tb, d_term, d_next, NearestTerm(), (int)n_lines, st);
    line_keep_kernel<<<nb_lines, kT, 0, st>>>(d_term, d_next, n_lines, d_kept);
  }

  // ---- pass 4: bytes -> kept codes ----------------------------------------------------------------------
  byte_flags_kernel<<<nb_bytes, kT, 0, st>>>(d_text, n, d_line_off, d_line_len, n_lines, d_kept, d_flag, d_brk);
  {
    size_t tb = cub_bytes;
    cub::DeviceScan::ExclusiveSum(d_cub, tb, WideFlags(d_flag, Widen()), d_u32a, (int)n, st);
    tb = cub_bytes;
    cub::DeviceScan::ExclusiveSum(d_cub, tb, WideFlags(d_brk, Widen()), d_u32b, (int)n, st);
  }
  // total kept = kpos[n-1] + keep[n-1]
  {
    uint32_t *h_tot = h_stage + n_files + 4;
    SKS_CUDA_TRY(cudaMemcpyAsync(h_tot, d_u32a + (n - 1), 4, cudaMemcpyDeviceToHost, st));
    SKS_CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<uint8_t *>(h_tot + 1), d_flag + (n - 1), 1, cudaMemcpyDeviceToHost, st));
    SKS_CUDA_TRY(cudaStreamSynchronize(st));
    const uint32_t total = h_tot[0] + (*reinterpret_cast<uint8_t *>(h_tot + 1) ? 1u : 0u);
    SKS_TRY(alloc_buffer(ctx, (size_t)total + 16, codes_buf));
    uint8_t *d_codes = static_cast<uint8_t *>((*codes_buf)->ptr);
    BufferRef bvbuf, endbuf;
    SKS_TRY(alloc_buffer(ctx, 4 * (size_t)total + 16, &bvbuf));
    uint32_t *d_bval = static_cast<uint32_t *>(bvbuf->ptr);
    scatter_codes_kernel<<<nb_bytes, kT, 0, st>>>(d_text, n, d_flag, d_brk, d_u32a, d_u32b, d_codes, d_bval);
    genome_base_kernel<<<(n_files + 1 + kT - 1) / kT, kT, 0, st>>>(d_file_off, n_files, n, d_u32a, total, d_genome_base);
    std::vector<uint32_t> h_base(n_files + 1, 0), h_ends;
    uint32_t n_ends = 0;
    if (total > 0) {
      // ---- pass 5: segment ends -----------------------------------------------------------------------
      const unsigned nb_tot = (total + kT - 1) / kT;
      seg_end_flags_kernel<<<nb_tot, kT, 0, st>>>(d_bval, total, d_genome_base, n_files, d_endflag);
      SKS_TRY(alloc_buffer(ctx, 4 * (size_t)total + 16, &endbuf));
      uint32_t *d_ends = static_cast<uint32_t *>(endbuf->ptr);
      cub::TransformInputIterator<uint32_t, PlusOne, cub::CountingInputIterator<uint32_t>> it(cub::CountingInputIterator<uint32_t>(0), PlusOne());
      size_t tb = cub_bytes;
      cub::DeviceSelect::Flagged(d_cub, tb, it, d_endflag, d_ends, d_count, (int)total, st);
      SKS_CUDA_TRY(cudaMemcpyAsync(h_count, d_count, 4, cudaMemcpyDeviceToHost, st));
      SKS_CUDA_TRY(cudaMemcpyAsync(h_base.data(), d_genome_base, 4 * (size_t)(n_files + 1), cudaMemcpyDeviceToHost, st));
      SKS_CUDA_TRY(cudaStreamSynchronize(st));
      n_ends = *h_count;
      h_ends.resize(n_ends);
      if (n_ends) {
        SKS_CUDA_TRY(cudaMemcpyAsync(h_ends.data(), d_ends, 4 * (size_t)n_ends, cudaMemcpyDeviceToHost, st));
        SKS_CUDA_TRY(cudaStreamSynchronize(st));
      }
    } else {
      SKS_CUDA_TRY(cudaStreamSynchronize(st));
    }
    // host: split the global list of segment ends by genome
    size_t e = 0;
    for (uint32_t g = 0; g < n_files; ++g) {
      const uint32_t lo = h_base[g], hi = h_base[g + 1];
      (*n_bases)[g] = hi - lo;
      uint32_t prev = lo;
      while (e < h_ends.size() && h_ends[e] <= hi) {
        (*seg_len)[g].push_back(h_ends[e] - prev);
        prev = h_ends[e];
        ++e;
      }
    }
    ctx->launches += 16;
  }
  return SKS_OK;
}

int launch_pack_codes(sks_ctx *ctx, const uint8_t *d_codes, const uint32_t *d_genome_base, const GenomeDesc *d_genomes,
                      int n_genomes, uint32_t max_words, uint32_t *d_words) {
  if (n_genomes == 0 || max_words == 0) return SKS_OK;
  KernelTimer timer(ctx, SKS_KERNEL_FASTA);
  unsigned gx = std::min<unsigned>((max_words + kT - 1) / kT, (unsigned)ctx->sm_count * 8);
  for (int g0 = 0; g0 < n_genomes; g0 += 65535) {
    const int ng = std::min(n_genomes - g0, 65535);
    pack_kernel<<<dim3(gx, ng), kT, 0, ctx->stream>>>(d_codes, d_genome_base + g0, d_genomes + g0, d_words);
    SKS_CUDA_TRY(cudaGetLastError());
    ctx->launches++;
  }
  return SKS_OK;
}

}  // namespace sks
