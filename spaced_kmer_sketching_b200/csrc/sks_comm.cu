// Multi-GPU entry points of the C ABI: one communicator per (rank, GPU), NCCL over NVLink / NVSwitch underneath.
//
// The reference's only parallelism is `cilk_for` over FASTA files (src/kmer_set.cpp:124-131) and over set pairs
// (src/kmer_set.cpp:179-182).  Sharded over GPUs that becomes: genomes are split into contiguous blocks by rank,
// every rank sketches its block, the sketches are exchanged once (sks_comm_allgather_sets), the ranks then split the KEY
// SPACE of the all-vs-all dictionary (csrc/sks_allpairs.cu): each enters its share of the distinct k-mers and counts
// what they contribute to every pair, and one reduce-scatter of the n x n partial counts leaves every rank with its
// own complete block rows of the pair matrix (generate_all_pairs_from_vector order, src/generators.hpp:44-58).  One long sequence (BASELINE
// configs[2]) is split by position instead; its partial sketches are routed by key range, so that every rank
// sort-uniques 1/world of the keys (sks_sketch_sequence_sharded).
//
// NCCL is bound at run time (dlopen): libsks.so has no link-time dependency on it, and inside a process that has
// already loaded a libnccl.so.2 (e.g. through torch) that same copy is used.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "sks_internal.cuh"

struct sks_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};

namespace sks {
namespace {

struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclReduceScatter) ReduceScatter = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  std::string error;
};

NcclApi *load_nccl() {
  NcclApi *api = new NcclApi();
  const char *env = getenv("SKS_NCCL_LIB");
  if (env && *env) api->handle = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  if (!api->handle) api->handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process
  if (!api->handle) api->handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!api->handle) api->handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!api->handle) {
    api->error = std::string("libnccl.so.2 could not be loaded: ") + (dlerror() ? dlerror() : "?");
    return api;
  }
#define SKS_NCCL_SYM(name)                                                          \
  api->name = reinterpret_cast<decltype(api->name)>(dlsym(api->handle, "nccl" #name)); \
  if (!api->name) api->error = "libnccl lacks nccl" #name;
  SKS_NCCL_SYM(GetUniqueId)
  SKS_NCCL_SYM(CommInitRank)
  SKS_NCCL_SYM(CommInitAll)
  SKS_NCCL_SYM(CommDestroy)
  SKS_NCCL_SYM(AllGather)
  SKS_NCCL_SYM(ReduceScatter)
  SKS_NCCL_SYM(Send)
  SKS_NCCL_SYM(Recv)
  SKS_NCCL_SYM(GroupStart)
  SKS_NCCL_SYM(GroupEnd)
  SKS_NCCL_SYM(GetErrorString)
  SKS_NCCL_SYM(GetVersion)
#undef SKS_NCCL_SYM
  return api;
}

NcclApi *nccl() {
  static NcclApi *api = load_nccl();
  return api;
}

#define SKS_NCCL_TRY(expr)                                                                                     \
  do {                                                                                                         \
    ncclResult_t _r = (expr);                                                                                  \
    if (_r != ncclSuccess)                                                                                     \
      return set_error(SKS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, nccl()->GetErrorString(_r), __FILE__, __LINE__); \
  } while (0)

int need_nccl() {
  NcclApi *api = nccl();
  if (!api->error.empty()) return set_error(SKS_ERR_CUDA, "%s", api->error.c_str());
  return SKS_OK;
}

// lower_bound of every splitter in an ascending key array: out[r] = number of keys < split[r]
template <int KW>
__global__ void split_offsets_kernel(const unsigned long long *__restrict__ keys, uint32_t n,
                                     const unsigned long long *__restrict__ split, int n_split,
                                     unsigned long long *__restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_split) return;
  const unsigned long long slo = split[2 * r], shi = split[2 * r + 1];
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    bool less;
    if (KW == 1) less = keys[mid] < slo;
    else less = keys[2 * mid + 1] != shi ? keys[2 * mid + 1] < shi : keys[2 * mid] < slo;
    if (less) lo = mid + 1; else hi = mid;
  }
  out[r] = lo;
}

// PDEP of a (2 * weight)-bit value into the mask's set bits: the r-th quantile of the key space under the mask
void deposit(unsigned __int128 v, const uint64_t mask[2], uint64_t out[2]) {
  out[0] = out[1] = 0;
  for (int b = 0; b < 128; ++b)
    if ((mask[b >> 6] >> (b & 63)) & 1) {
      if (v & 1) out[b >> 6] |= 1ull << (b & 63);
      v >>= 1;
    }
}

}  // namespace

// The sets of `local` as one contiguous key range on the device: in place when they already lie back to back
// (the sets of one sks_sketch call do), packed into *packed otherwise.
int contiguous_keys(sks_ctx *ctx, sks_set *const *local, int64_t n_local, const void **ptr, int64_t *n_keys, BufferRef *packed) {
  int64_t total = 0;
  bool contiguous = true;
  const char *expect = nullptr;
  const size_t kb = n_local ? (size_t)local[0]->key_words * 8 : 8;
  for (int64_t i = 0; i < n_local; ++i) {
    const char *p = static_cast<const char *>(local[i]->buf->ptr) + local[i]->byte_off;
    if (local[i]->count > 0) {
      if (expect && p != expect) contiguous = false;
      if (!expect) *ptr = p;
      expect = p + (size_t)local[i]->count * kb;
    }
    total += local[i]->count;
  }
  *n_keys = total;
  if (total == 0) {
    *ptr = nullptr;
    return SKS_OK;
  }
  if (contiguous) return SKS_OK;
  SKS_TRY(alloc_buffer(ctx, (size_t)total * kb, packed));
  char *dst = static_cast<char *>((*packed)->ptr);
  for (int64_t i = 0; i < n_local; ++i) {
    const size_t bytes = (size_t)local[i]->count * kb;
    if (bytes)
      SKS_CUDA_TRY(cudaMemcpyAsync(dst, static_cast<const char *>(local[i]->buf->ptr) + local[i]->byte_off, bytes,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
    dst += bytes;
  }
  *ptr = (*packed)->ptr;
  return SKS_OK;
}

}  // namespace sks

using namespace sks;

extern "C" {

void sks_shard_range(int64_t n, int rank, int world, int64_t *begin, int64_t *end) {
  if (world < 1) world = 1;
  const int64_t per = (n + world - 1) / world;
  if (begin) *begin = std::min<int64_t>((int64_t)rank * per, n);
  if (end) *end = std::min<int64_t>((int64_t)(rank + 1) * per, n);
}

int sks_comm_unique_id(void *out_id) {
  if (!out_id) return set_error(SKS_ERR_INVALID, "null argument");
  SKS_TRY(need_nccl());
  ncclUniqueId id;
  SKS_NCCL_TRY(nccl()->GetUniqueId(&id));
  static_assert(sizeof(id) == SKS_COMM_ID_BYTES, "SKS_COMM_ID_BYTES must match ncclUniqueId");
  memcpy(out_id, &id, sizeof(id));
  return SKS_OK;
}

int sks_comm_init_rank(sks_ctx *ctx, const void *id, int rank, int world, sks_comm **out) {
  if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return set_error(SKS_ERR_INVALID, "bad argument");
  SKS_TRY(need_nccl());
  DeviceGuard guard(ctx->device);
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  sks_comm *c = new (std::nothrow) sks_comm();
  if (!c) return set_error(SKS_ERR_INVALID, "out of host memory");
  c->rank = rank;
  c->world = world;
  c->device = ctx->device;
  ncclResult_t r = nccl()->CommInitRank(&c->comm, world, uid, rank);
  if (r != ncclSuccess) {
    delete c;
    return set_error(SKS_ERR_CUDA, "ncclCommInitRank failed: %s", nccl()->GetErrorString(r));
  }
  *out = c;
  return SKS_OK;
}

int sks_comm_init_all(sks_ctx *const *ctxs, int n, sks_comm **out) {
  if (!ctxs || !out || n < 1) return set_error(SKS_ERR_INVALID, "bad argument");
  SKS_TRY(need_nccl());
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) return set_error(SKS_ERR_INVALID, "null context");
    devs[i] = ctxs[i]->device;
  }
  std::vector<ncclComm_t> comms(n);
  SKS_NCCL_TRY(nccl()->CommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; ++i) {
    sks_comm *c = new sks_comm();
    c->comm = comms[i];
    c->rank = i;
    c->world = n;
    c->device = devs[i];
    out[i] = c;
  }
  return SKS_OK;
}

void sks_comm_destroy(sks_comm *c) {
  if (!c) return;
  if (c->comm && nccl()->CommDestroy) {
    DeviceGuard guard(c->device);
    nccl()->CommDestroy(c->comm);
  }
  delete c;
}

int sks_comm_rank(const sks_comm *c) { return c ? c->rank : 0; }
int sks_comm_world(const sks_comm *c) { return c ? c->world : 1; }
int sks_comm_nccl_version(void) {
  if (need_nccl() != SKS_OK) return 0;
  int v = 0;
  nccl()->GetVersion(&v);
  return v;
}

int sks_comm_allgather_sets(sks_ctx *ctx, sks_comm *comm, sks_set *const *local, int64_t n_local, int64_t n_total,
                            sks_set **out_all) {
  if (!ctx || !out_all || n_local < 0 || n_total < 0 || (n_local > 0 && !local)) return set_error(SKS_ERR_INVALID, "bad argument");
  const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
  int64_t begin = 0, end = 0;
  sks_shard_range(n_total, rank, world, &begin, &end);
  if (end - begin != n_local)
    return set_error(SKS_ERR_MISMATCH, "rank %d holds %lld sets, its block of %lld sets over %d ranks has %lld", rank,
                     (long long)n_local, (long long)n_total, world, (long long)(end - begin));
  for (int64_t i = 0; i < n_local; ++i) {
    if (!local[i] || local[i]->repr != SKS_REPR_SORTED) return set_error(SKS_ERR_INVALID, "the exchange needs SORTED sets");
    SKS_TRY(check_pair(local[0], local[i]));
    if (local[i]->device != ctx->device) return set_error(SKS_ERR_INVALID, "set lives on another device");
  }
  DeviceGuard guard(ctx->device);
  auto share = [&](const sks_set *s) {  // a second handle on the same keys
    sks_set *c = new sks_set(*s);
    return c;
  };
  if (world == 1) {
    for (int64_t i = 0; i < n_local; ++i) out_all[i] = share(local[i]);
    return SKS_OK;
  }
  if (comm->device != ctx->device) return set_error(SKS_ERR_INVALID, "communicator and context are on different devices");
  SKS_TRY(need_nccl());
  // header of every rank: [n_keys, key_words, window, mask lo, mask hi, count of each of its `per` sets]
  const int64_t per = (n_total + world - 1) / world;
  const size_t hdr = 5 + (size_t)per;
  BufferRef d_hdr;
  SKS_TRY(alloc_buffer(ctx, 8 * hdr * ((size_t)world + 1), &d_hdr));
  long long *h_mine = nullptr, *h_all = nullptr;
  SKS_TRY(ctx_pinned(ctx, 8 * hdr * ((size_t)world + 1), reinterpret_cast<void **>(&h_mine)));
  h_all = h_mine + hdr;
  const void *my_keys = nullptr;
  int64_t my_n = 0;
  BufferRef packed;
  SKS_TRY(contiguous_keys(ctx, local, n_local, &my_keys, &my_n, &packed));
  memset(h_mine, 0, 8 * hdr);
  h_mine[0] = my_n;
  h_mine[1] = n_local ? local[0]->key_words : 0;
  h_mine[2] = n_local ? local[0]->window : 0;
  h_mine[3] = n_local ? (long long)local[0]->mask[0] : 0;
  h_mine[4] = n_local ? (long long)local[0]->mask[1] : 0;
  for (int64_t i = 0; i < n_local; ++i) h_mine[5 + i] = local[i]->count;
  long long *d_mine = static_cast<long long *>(d_hdr->ptr), *d_all = d_mine + hdr;
  KernelTimer timer(ctx, SKS_KERNEL_EXCHANGE);
  SKS_CUDA_TRY(cudaMemcpyAsync(d_mine, h_mine, 8 * hdr, cudaMemcpyHostToDevice, ctx->stream));
  SKS_NCCL_TRY(nccl()->AllGather(d_mine, d_all, hdr, ncclInt64, comm->comm, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_all, d_all, 8 * hdr * world, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // the one host synchronisation of the exchange
  // every non-empty rank must agree on the key layout
  int kw = 0, window = 0;
  uint64_t mask[2] = {0, 0};
  int64_t total_remote = 0;
  for (int r = 0; r < world; ++r) {
    const long long *h = h_all + hdr * r;
    if (h[1] == 0) continue;  // a rank without sets
    if (kw == 0) {
      kw = (int)h[1];
      window = (int)h[2];
      mask[0] = (uint64_t)h[3];
      mask[1] = (uint64_t)h[4];
    } else if (kw != (int)h[1] || mask[0] != (uint64_t)h[3] || mask[1] != (uint64_t)h[4]) {
      return set_error(SKS_ERR_MISMATCH, "ranks sketched with different masks or key widths");
    }
    if (r != rank) total_remote += h[0];
  }
  const size_t kb = (size_t)std::max(kw, 1) * 8;
  BufferRef inbox;
  SKS_TRY(alloc_buffer(ctx, (size_t)total_remote * kb, &inbox));
  // one grouped exchange: my keys to every peer, every peer's keys into its place of the inbox
  std::vector<size_t> at(world, 0);
  {
    size_t off = 0;
    for (int r = 0; r < world; ++r) {
      at[r] = off;
      if (r != rank) off += (size_t)h_all[hdr * r] * kb;
    }
  }
  SKS_NCCL_TRY(nccl()->GroupStart());
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    if (my_n > 0) SKS_NCCL_TRY(nccl()->Send(my_keys, (size_t)my_n * kb, ncclUint8, r, comm->comm, ctx->stream));
    const size_t bytes = (size_t)h_all[hdr * r] * kb;
    if (bytes) SKS_NCCL_TRY(nccl()->Recv(static_cast<char *>(inbox->ptr) + at[r], bytes, ncclUint8, r, comm->comm, ctx->stream));
  }
  SKS_NCCL_TRY(nccl()->GroupEnd());
  const int weight = sks_mask_weight(mask);
  for (int r = 0; r < world; ++r) {
    int64_t b = 0, e = 0;
    sks_shard_range(n_total, r, world, &b, &e);
    size_t off = at[r];
    for (int64_t i = b; i < e; ++i) {
      if (r == rank) {
        out_all[i] = share(local[i - b]);
        continue;
      }
      sks_set *s = new_set(ctx, SKS_REPR_SORTED, mask, window, weight);
      s->buf = inbox;
      s->key_words = kw;
      s->byte_off = off;
      s->count = h_all[hdr * r + 5 + (i - b)];
      off += (size_t)s->count * kb;
      out_all[i] = s;
    }
  }
  // the packed copy of non-contiguous local keys must outlive the sends: they are stream ordered, and so is its release
  return SKS_OK;
}

int sks_all_vs_all_sharded(sks_ctx *ctx, sks_comm *comm, sks_set *const *local, int64_t n_local, int64_t n_total,
                           int32_t *out_counts, int32_t *out_sizes, double *out_ani) {
  if (!ctx) return set_error(SKS_ERR_INVALID, "null context");
  const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
  std::vector<sks_set *> all((size_t)std::max<int64_t>(n_total, 1), nullptr);
  struct Release {
    sks_ctx *c;
    std::vector<sks_set *> &v;
    ~Release() {
      for (sks_set *s : v)
        if (s) sks_set_destroy(c, s);
    }
  } release{ctx, all};
  SKS_TRY(sks_comm_allgather_sets(ctx, comm, local, n_local, n_total, all.data()));
  int64_t begin = 0, end = 0;
  sks_shard_range(n_total, rank, world, &begin, &end);
  if (world == 1 || n_total < 2 || !all_pairs_dict_eligible(all.data(), n_total))
    return sks_all_vs_all(ctx, all.data(), n_total, begin, end, out_counts, out_sizes, out_ani);
  // Several ranks split the KEY SPACE of the dictionary, not the rows: every rank enters 1 / world of the distinct
  // keys (all sets are here after the exchange), counts what its keys contribute to every pair, and one reduce-scatter
  // adds the shares up and leaves every rank with its own block rows.
  DeviceGuard guard(ctx->device);
  const int64_t per = (n_total + world - 1) / world, n_rows = end - begin;
  BufferRef raw, sizes, mine, counts, ani;
  int st = all_pairs_raw(ctx, all.data(), n_total, 0, n_total, rank, world, false, (int64_t)world * per, &raw, &sizes);
  if (st == SKS_ERR_CAPACITY) return sks_all_vs_all(ctx, all.data(), n_total, begin, end, out_counts, out_sizes, out_ani);
  SKS_TRY(st);
  SKS_TRY(alloc_buffer(ctx, 4 * (size_t)per * n_total, &mine));
  {
    KernelTimer timer(ctx, SKS_KERNEL_EXCHANGE);
    SKS_NCCL_TRY(nccl()->ReduceScatter(raw->ptr, mine->ptr, (size_t)per * n_total, ncclInt32, ncclSum, comm->comm, ctx->stream));
  }
  SKS_TRY(all_pairs_finalize(ctx, static_cast<const int32_t *>(mine->ptr), static_cast<const int32_t *>(sizes->ptr), n_total, begin,
                             n_rows, false, all[0]->weight, &counts, out_ani ? &ani : nullptr));
  if (out_counts && n_rows)
    SKS_CUDA_TRY(cudaMemcpyAsync(out_counts, counts->ptr, 4 * (size_t)n_rows * n_total, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_ani && n_rows)
    SKS_CUDA_TRY(cudaMemcpyAsync(out_ani, ani->ptr, 8 * (size_t)n_rows * n_total, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_sizes)
    for (int64_t i = 0; i < n_total; ++i) out_sizes[i] = (int32_t)all[i]->count;
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return SKS_OK;
}

int sks_sketch_sequence_sharded(sks_ctx *ctx, sks_comm *comm, const sks_batch *slice, const uint64_t mask[2], int window,
                                const sks_pred *pred, int gather, sks_set **out, int64_t *out_global_size) {
  if (!ctx || !slice || !mask || !pred || !out) return set_error(SKS_ERR_INVALID, "null argument");
  if (sks_batch_n_genomes(slice) != 1) return set_error(SKS_ERR_INVALID, "a sequence slice is a single-genome batch");
  const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
  if (world > 64) return set_error(SKS_ERR_INVALID, "at most 64 ranks");
  sks_set *loc = nullptr;
  SKS_TRY(sks_sketch(ctx, slice, mask, window, pred, SKS_REPR_SORTED, &loc));
  if (world == 1) {
    *out = loc;
    if (out_global_size) *out_global_size = loc->count;
    return SKS_OK;
  }
  struct Drop {  // the local sketch is an intermediate from here on
    sks_ctx *c;
    sks_set *s;
    ~Drop() { sks_set_destroy(c, s); }
  } drop{ctx, loc};
  SKS_TRY(need_nccl());
  DeviceGuard guard(ctx->device);
  const int kw = loc->key_words;
  const size_t kb = (size_t)kw * 8;
  const int weight = sks_mask_weight(mask);
  // key range r = [split[r], split[r + 1]): equal shares of the 4^weight possible keys under the mask
  unsigned long long *h_tab = nullptr;
  const size_t n_tab = 2 * (size_t)world + (size_t)world * world + 2 * (size_t)world;
  SKS_TRY(ctx_pinned(ctx, 8 * n_tab, reinterpret_cast<void **>(&h_tab)));
  BufferRef d_tab;
  SKS_TRY(alloc_buffer(ctx, 8 * n_tab, &d_tab));
  unsigned long long *d_split = static_cast<unsigned long long *>(d_tab->ptr), *d_lb = d_split + 2 * world,
                     *d_all = d_lb + world;
  for (int r = 0; r < world; ++r) {
    uint64_t s[2] = {0, 0};
    if (r > 0) {
      const int bits = 2 * weight;
      unsigned __int128 q;
      if (bits >= 128) q = (~(unsigned __int128)0) / world * r;
      else q = (((unsigned __int128)1 << bits) / world) * r;
      deposit(q, mask, s);
    }
    h_tab[2 * r] = s[0];
    h_tab[2 * r + 1] = s[1];
  }
  SKS_CUDA_TRY(cudaMemcpyAsync(d_split, h_tab, 16 * (size_t)world, cudaMemcpyHostToDevice, ctx->stream));
  const unsigned long long *keys = reinterpret_cast<const unsigned long long *>(static_cast<const char *>(loc->buf->ptr) + loc->byte_off);
  if (kw == 1) split_offsets_kernel<1><<<1, 64, 0, ctx->stream>>>(keys, (uint32_t)loc->count, d_split, world, d_lb);
  else split_offsets_kernel<2><<<1, 64, 0, ctx->stream>>>(keys, (uint32_t)loc->count, d_split, world, d_lb);
  SKS_CUDA_TRY(cudaGetLastError());
  ctx->launches++;
  // every rank's lower bounds to every rank: row r of the table = where rank r's key ranges start
  SKS_NCCL_TRY(nccl()->AllGather(d_lb, d_all, (size_t)world, ncclUint64, comm->comm, ctx->stream));
  unsigned long long *h_all = h_tab + 2 * world;
  unsigned long long *h_cnt = h_all + (size_t)world * world;  // [world] local key counts, all-gathered below
  // the counts travel in the same table: rank r's total = its last range's end; all-gather them as well
  BufferRef d_cnt;
  SKS_TRY(alloc_buffer(ctx, 8 * ((size_t)world + 1), &d_cnt));
  unsigned long long *d_mycnt = static_cast<unsigned long long *>(d_cnt->ptr), *d_cnts = d_mycnt + 1;
  h_cnt[world] = (unsigned long long)loc->count;
  SKS_CUDA_TRY(cudaMemcpyAsync(d_mycnt, h_cnt + world, 8, cudaMemcpyHostToDevice, ctx->stream));
  SKS_NCCL_TRY(nccl()->AllGather(d_mycnt, d_cnts, 1, ncclUint64, comm->comm, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_all, d_all, 8 * (size_t)world * world, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnts, 8 * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  // (the pinned ring may hand the staging area out again inside the calls below: keep what is needed)
  const std::vector<unsigned long long> all_lb(h_all, h_all + (size_t)world * world), all_cnt(h_cnt, h_cnt + world);
  auto range_begin = [&](int src, int r) { return (int64_t)all_lb[(size_t)src * world + r]; };
  auto range_end = [&](int src, int r) { return r + 1 < world ? (int64_t)all_lb[(size_t)src * world + r + 1] : (int64_t)all_cnt[src]; };
  // inbox: the keys of my range from every rank (mine by a device copy)
  int64_t n_in = 0;
  std::vector<int64_t> in_at(world);
  for (int src = 0; src < world; ++src) {
    in_at[src] = n_in;
    n_in += range_end(src, rank) - range_begin(src, rank);
  }
  BufferRef inbox;
  SKS_TRY(alloc_buffer(ctx, (size_t)std::max<int64_t>(n_in, 1) * kb, &inbox));
  SKS_NCCL_TRY(nccl()->GroupStart());
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    const int64_t sb = range_begin(rank, r), se = range_end(rank, r);
    if (se > sb)
      SKS_NCCL_TRY(nccl()->Send(reinterpret_cast<const char *>(keys) + (size_t)sb * kb, (size_t)(se - sb) * kb, ncclUint8, r,
                                comm->comm, ctx->stream));
    const int64_t rn = range_end(r, rank) - range_begin(r, rank);
    if (rn > 0)
      SKS_NCCL_TRY(nccl()->Recv(static_cast<char *>(inbox->ptr) + (size_t)in_at[r] * kb, (size_t)rn * kb, ncclUint8, r,
                                comm->comm, ctx->stream));
  }
  SKS_NCCL_TRY(nccl()->GroupEnd());
  {
    const int64_t sb = range_begin(rank, rank), se = range_end(rank, rank);
    if (se > sb)
      SKS_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(inbox->ptr) + (size_t)in_at[rank] * kb,
                                   reinterpret_cast<const char *>(keys) + (size_t)sb * kb, (size_t)(se - sb) * kb,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
  }
  // my range of the global set: sort + unique of what arrived (the same k-mer can occur in several slices)
  sks_set *mine = nullptr;
  SKS_TRY(sks_set_from_unsorted_device_keys(ctx, inbox->ptr, n_in, kw, mask, window, &mine));
  // sizes of all ranges (for the global size, and for the optional gather)
  SKS_TRY(ctx_pinned(ctx, 8 * ((size_t)world + 1), reinterpret_cast<void **>(&h_cnt)));
  h_cnt[world] = (unsigned long long)mine->count;
  SKS_CUDA_TRY(cudaMemcpyAsync(d_mycnt, h_cnt + world, 8, cudaMemcpyHostToDevice, ctx->stream));
  SKS_NCCL_TRY(nccl()->AllGather(d_mycnt, d_cnts, 1, ncclUint64, comm->comm, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnts, 8 * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  int64_t global = 0;
  for (int r = 0; r < world; ++r) global += (int64_t)h_cnt[r];
  if (out_global_size) *out_global_size = global;
  if (!gather) {
    *out = mine;
    return SKS_OK;
  }
  // the ranges are globally ordered: concatenated in rank order they are the sorted global set
  sks_set *full = new_set(ctx, SKS_REPR_SORTED, mask, window, weight);
  if (!full) {
    sks_set_destroy(ctx, mine);
    return set_error(SKS_ERR_INVALID, "out of host memory");
  }
  full->key_words = kw;
  full->count = global;
  int st = alloc_buffer(ctx, (size_t)std::max<int64_t>(global, 1) * kb, &full->buf);
  if (st == SKS_OK) {
    ncclResult_t nr = nccl()->GroupStart();
    int64_t at = 0;
    const char *src = static_cast<const char *>(mine->buf->ptr) + mine->byte_off;
    for (int r = 0; r < world && nr == ncclSuccess; ++r) {
      char *dst = static_cast<char *>(full->buf->ptr) + (size_t)at * kb;
      if (r == rank) {
        for (int p = 0; p < world && nr == ncclSuccess; ++p)
          if (p != rank && mine->count > 0) nr = nccl()->Send(src, (size_t)mine->count * kb, ncclUint8, p, comm->comm, ctx->stream);
      } else if (h_cnt[r] > 0) {
        nr = nccl()->Recv(dst, (size_t)h_cnt[r] * kb, ncclUint8, r, comm->comm, ctx->stream);
      }
      at += (int64_t)h_cnt[r];
    }
    const ncclResult_t ne = nccl()->GroupEnd();
    if (nr == ncclSuccess) nr = ne;
    if (nr != ncclSuccess) st = set_error(SKS_ERR_CUDA, "NCCL gather of the key ranges failed: %s", nccl()->GetErrorString(nr));
    if (st == SKS_OK && mine->count > 0) {
      int64_t my_at = 0;
      for (int r = 0; r < rank; ++r) my_at += (int64_t)h_cnt[r];
      if (cudaMemcpyAsync(static_cast<char *>(full->buf->ptr) + (size_t)my_at * kb, src, (size_t)mine->count * kb,
                          cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
        st = set_error(SKS_ERR_CUDA, "device copy failed");
    }
    // `mine` may be recycled as soon as it is destroyed: its block returns in stream order, after the sends
  }
  sks_set_destroy(ctx, mine);
  if (st != SKS_OK) {
    sks_set_destroy(ctx, full);
    return st;
  }
  *out = full;
  return SKS_OK;
}

}  // extern "C"
