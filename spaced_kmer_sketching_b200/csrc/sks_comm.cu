// Multi-GPU entry points of the C ABI: one communicator per (rank, GPU), NCCL over NVLink / NVSwitch underneath.
//
// The reference's only parallelism is `cilk_for` over FASTA files (src/kmer_set.cpp:124-131) and over set pairs
// (src/kmer_set.cpp:179-182).  Sharded over GPUs that becomes: genomes are split into contiguous blocks by rank,
// every rank sketches its block, the ranks split the KEY SPACE of the all-vs-all dictionary (csrc/sks_allpairs.cu) by a
// hash, every key travels once to the rank that owns it, each owner counts what its k-mers contribute to every pair, and
// one reduce-scatter of the n x n partial counts leaves every rank with its own complete block rows of the pair matrix
// (generate_all_pairs_from_vector order, src/generators.hpp:44-58).  sks_comm_allgather_sets puts every sketch on every
// rank for callers that want that.  One long sequence (BASELINE
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

// Key range of a key: the number of splitters (ascending, the first is 0) that are <= key, minus one.
template <int KW>
__device__ __forceinline__ int key_range(const unsigned long long *s_split, int world, unsigned long long lo, unsigned long long hi) {
  int a = 0, b = world;  // split[a] <= key < split[b]
  while (b - a > 1) {
    const int m = (a + b) >> 1;
    const unsigned long long slo = s_split[2 * m], shi = s_split[2 * m + 1];
    const bool le = KW == 1 ? slo <= lo : (shi != hi ? shi < hi : slo <= lo);
    if (le) a = m; else b = m;
  }
  return a;
}

// Routing of unsorted keys by range.  kScatter == false: counts[r] += keys of range r.  kScatter == true: the keys go
// to out[offset[r] + position], positions handed out by cursor[r] (order inside a range does not matter: the
// receiver sorts).
template <int KW, bool kScatter>
__global__ void __launch_bounds__(256)
    route_kernel(const unsigned long long *__restrict__ keys, unsigned long long n, const unsigned long long *__restrict__ split,
                 int world, unsigned long long *__restrict__ counts, const unsigned long long *__restrict__ offset,
                 unsigned long long *__restrict__ cursor, unsigned long long *__restrict__ out) {
  __shared__ unsigned long long s_split[128];
  for (int i = threadIdx.x; i < 2 * world; i += blockDim.x) s_split[i] = split[i];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long n_round = (n + 31) & ~31ull;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    unsigned long long lo = 0, hi = 0;
    int r = -1;
    if (i < n) {
      lo = keys[KW * i];
      hi = KW == 2 ? keys[KW * i + 1] : 0ull;
      r = key_range<KW>(s_split, world, lo, hi);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, r);
    if (r < 0) continue;
    const uint32_t leader = (uint32_t)(__ffs(peers) - 1), rank_in = (uint32_t)__popc(peers & ((1u << lane) - 1));
    if (!kScatter) {
      if (lane == leader) atomicAdd(counts + r, (unsigned long long)__popc(peers));
    } else {
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(cursor + r, (unsigned long long)__popc(peers));
      base = __shfl_sync(peers, base, (int)leader);
      const unsigned long long at = offset[r] + base + rank_in;
      out[KW * at] = lo;
      if (KW == 2) out[KW * at + 1] = hi;
    }
  }
}

// offset[r] = counts[0] + ... + counts[r - 1]; cursor[r] = 0
__global__ void route_offsets_kernel(const unsigned long long *__restrict__ counts, int world, unsigned long long *__restrict__ offset,
                                     unsigned long long *__restrict__ cursor) {
  if (threadIdx.x == 0) {
    unsigned long long run = 0;
    for (int r = 0; r < world; ++r) {
      offset[r] = run;
      cursor[r] = 0;
      run += counts[r];
    }
  }
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

namespace sks {
namespace {
// What every rank knows about every rank's sets after one small all-gather (the only host synchronisation of the
// exchange): per rank [n_keys, key_words, window, mask lo, mask hi, count of each of its `per` sets].
struct Headers {
  std::vector<long long> all;
  size_t hdr = 0;
  int kw = 0, window = 0;
  uint64_t mask[2] = {0, 0};
  uint64_t total_keys = 0;
  size_t n_extra = 0;
  const long long *of(int r) const { return all.data() + hdr * (size_t)r; }
  const long long *extra(int r) const { return of(r) + (hdr - n_extra); }
};

int check_local_sets(sks_ctx *ctx, sks_comm *comm, sks_set *const *local, int64_t n_local, int64_t n_total) {
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
  if (comm && comm->device != ctx->device) return set_error(SKS_ERR_INVALID, "communicator and context are on different devices");
  return SKS_OK;
}

// d_extra / n_extra: device words every rank appends to its record (H->extra(r) afterwards).
int exchange_headers(sks_ctx *ctx, sks_comm *comm, sks_set *const *local, int64_t n_local, int64_t n_total, Headers *H,
                     const unsigned long long *d_extra = nullptr, size_t n_extra = 0) {
  const int world = comm->world;
  const int64_t per = (n_total + world - 1) / world;
  const size_t hdr = 5 + (size_t)per + n_extra;
  H->n_extra = n_extra;
  BufferRef d_hdr;
  SKS_TRY(alloc_buffer(ctx, 8 * hdr * ((size_t)world + 1), &d_hdr));
  long long *h_mine = nullptr, *h_all = nullptr;
  SKS_TRY(ctx_pinned(ctx, 8 * hdr * ((size_t)world + 1), reinterpret_cast<void **>(&h_mine)));
  h_all = h_mine + hdr;
  memset(h_mine, 0, 8 * hdr);
  int64_t my_n = 0;
  for (int64_t i = 0; i < n_local; ++i) my_n += local[i]->count;
  h_mine[0] = my_n;
  h_mine[1] = n_local ? local[0]->key_words : 0;
  h_mine[2] = n_local ? local[0]->window : 0;
  h_mine[3] = n_local ? (long long)local[0]->mask[0] : 0;
  h_mine[4] = n_local ? (long long)local[0]->mask[1] : 0;
  for (int64_t i = 0; i < n_local; ++i) h_mine[5 + i] = local[i]->count;
  long long *d_mine = static_cast<long long *>(d_hdr->ptr), *d_all = d_mine + hdr;
  SKS_CUDA_TRY(cudaMemcpyAsync(d_mine, h_mine, 8 * (hdr - n_extra), cudaMemcpyHostToDevice, ctx->stream));
  if (n_extra)
    SKS_CUDA_TRY(cudaMemcpyAsync(d_mine + (hdr - n_extra), d_extra, 8 * n_extra, cudaMemcpyDeviceToDevice, ctx->stream));
  SKS_NCCL_TRY(nccl()->AllGather(d_mine, d_all, hdr, ncclInt64, comm->comm, ctx->stream));
  SKS_CUDA_TRY(cudaMemcpyAsync(h_all, d_all, 8 * hdr * world, cudaMemcpyDeviceToHost, ctx->stream));
  SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  H->all.assign(h_all, h_all + hdr * world);  // (the pinned ring may hand the staging area out again)
  H->hdr = hdr;
  H->kw = H->window = 0;
  H->total_keys = 0;
  for (int r = 0; r < world; ++r) {  // every non-empty rank must agree on the key layout
    const long long *h = H->of(r);
    H->total_keys += (uint64_t)h[0];
    if (h[1] == 0) continue;  // a rank without sets
    if (H->kw == 0) {
      H->kw = (int)h[1];
      H->window = (int)h[2];
      H->mask[0] = (uint64_t)h[3];
      H->mask[1] = (uint64_t)h[4];
    } else if (H->kw != (int)h[1] || H->mask[0] != (uint64_t)h[3] || H->mask[1] != (uint64_t)h[4]) {
      return set_error(SKS_ERR_MISMATCH, "ranks sketched with different masks or key widths");
    }
  }
  return SKS_OK;
}
}  // namespace
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
  SKS_TRY(check_local_sets(ctx, comm, local, n_local, n_total));
  DeviceGuard guard(ctx->device);
  auto share = [&](const sks_set *s) {  // a second handle on the same keys
    sks_set *c = new sks_set(*s);
    return c;
  };
  if (world == 1) {
    for (int64_t i = 0; i < n_local; ++i) out_all[i] = share(local[i]);
    return SKS_OK;
  }
  SKS_TRY(need_nccl());
  KernelTimer timer(ctx, SKS_KERNEL_EXCHANGE);
  Headers H;
  SKS_TRY(exchange_headers(ctx, comm, local, n_local, n_total, &H));
  const size_t hdr = H.hdr;
  const long long *h_all = H.all.data();
  const int kw = H.kw, window = H.window;
  uint64_t mask[2] = {H.mask[0], H.mask[1]};
  const void *my_keys = nullptr;
  int64_t my_n = 0, total_remote = 0;
  BufferRef packed;
  SKS_TRY(contiguous_keys(ctx, local, n_local, &my_keys, &my_n, &packed));
  const size_t kb = (size_t)std::max(kw, 1) * 8;
  // One in-place ncclAllGather of equal slots (the ranks' key counts differ by a few per cent at most: the slot is the
  // largest): slot r of the inbox receives rank r's keys, mine go there by a device copy first.
  int64_t slot_keys = 0;
  for (int r = 0; r < world; ++r) slot_keys = std::max<int64_t>(slot_keys, h_all[hdr * r]);
  const size_t slot = ((size_t)slot_keys * kb + 255) & ~(size_t)255;
  (void)total_remote;
  BufferRef inbox;
  SKS_TRY(alloc_buffer(ctx, std::max<size_t>(slot * world, 16), &inbox));
  std::vector<size_t> at(world, 0);
  for (int r = 0; r < world; ++r) at[r] = slot * r;
  if (slot > 0) {
    char *mine = static_cast<char *>(inbox->ptr) + at[rank];
    if (my_n > 0) SKS_CUDA_TRY(cudaMemcpyAsync(mine, my_keys, (size_t)my_n * kb, cudaMemcpyDeviceToDevice, ctx->stream));
    SKS_NCCL_TRY(nccl()->AllGather(mine, inbox->ptr, slot, ncclUint8, comm->comm, ctx->stream));
  }
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
  if (!ctx || n_local < 0 || n_total < 0 || (n_local > 0 && !local)) return set_error(SKS_ERR_INVALID, "bad argument");
  const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
  int64_t begin = 0, end = 0;
  sks_shard_range(n_total, rank, world, &begin, &end);
  auto by_rows = [&]() {  // every set on every rank, then the rank's rows through the single-device call
    std::vector<sks_set *> all((size_t)std::max<int64_t>(n_total, 1), nullptr);
    int st = sks_comm_allgather_sets(ctx, comm, local, n_local, n_total, all.data());
    if (st == SKS_OK) st = sks_all_vs_all(ctx, all.data(), n_total, begin, end, out_counts, out_sizes, out_ani);
    for (sks_set *s : all)
      if (s) sks_set_destroy(ctx, s);
    return st;
  };
  if (world == 1) return by_rows();
  SKS_TRY(check_local_sets(ctx, comm, local, n_local, n_total));
  SKS_TRY(need_nccl());
  DeviceGuard guard(ctx->device);
  const size_t W = (size_t)world;
  const int64_t per = (n_total + world - 1) / world, n_rows = end - begin;
  BufferRef raw, sizes, mine, counts, ani;
  const uint32_t *overflow = nullptr;
  Headers H;
  std::vector<int32_t> h_sizes;
  auto sizes_from_headers = [&]() {  // the sizes of all sets (ANI denominators, the diagonal) come with the headers
    h_sizes.assign((size_t)n_total, 0);
    for (int r = 0; r < world; ++r) {
      int64_t b = 0, e = 0;
      sks_shard_range(n_total, r, world, &b, &e);
      for (int64_t i = b; i < e; ++i) h_sizes[(size_t)i] = (int32_t)H.of(r)[5 + (i - b)];
    }
  };
  auto finish = [&]() -> int {  // partial counts of all rows -> my complete rows -> counts / ANI on the host
    SKS_TRY(alloc_buffer(ctx, 4 * (size_t)per * n_total, &mine));
    {
      KernelTimer timer(ctx, SKS_KERNEL_EXCHANGE);
      SKS_NCCL_TRY(nccl()->ReduceScatter(raw->ptr, mine->ptr, (size_t)per * n_total, ncclInt32, ncclSum, comm->comm, ctx->stream));
    }
    SKS_TRY(all_pairs_finalize(ctx, static_cast<const int32_t *>(mine->ptr), static_cast<const int32_t *>(sizes->ptr), n_total, begin,
                               n_rows, false, sks_mask_weight(H.mask), &counts, out_ani ? &ani : nullptr));
    if (out_counts && n_rows)
      SKS_CUDA_TRY(cudaMemcpyAsync(out_counts, counts->ptr, 4 * (size_t)n_rows * n_total, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_ani && n_rows)
      SKS_CUDA_TRY(cudaMemcpyAsync(out_ani, ani->ptr, 8 * (size_t)n_rows * n_total, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_sizes) memcpy(out_sizes, h_sizes.data(), 4 * (size_t)n_total);
    SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (overflow && *overflow)  // (every rank completes the collective first; a full table is reported, not papered over)
      return set_error(SKS_ERR_CAPACITY, "the dictionary of rank %d overflowed", rank);
    return SKS_OK;
  };
  // The ranks split the KEY SPACE of the all-vs-all dictionary (csrc/sks_allpairs.cu), not the rows: a key belongs to the
  // rank its hash names, the owner counts what its keys contribute to EVERY pair, and one reduce-scatter of the n x n
  // partial counts leaves every rank with its own complete block rows.  Two ways for the keys to reach their owners:
  //   gather (two ranks): every sketch to every rank (sks_comm_allgather_sets), each rank enters only the keys it owns;
  //   owner (more ranks): every rank sends each of its keys, with the number of its set, to the owner -- 1 / world of
  //                       the bytes, and nobody scans keys that are not theirs.
  static const int forced = [] {
    const char *e = getenv("SKS_SHARD_ROUTE");
    return !e ? 0 : (strcmp(e, "gather") == 0 ? 1 : (strcmp(e, "owner") == 0 ? 2 : 0));
  }();
  const bool by_owner = forced ? forced == 2 : world > 2;
  if (!by_owner) {
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
    if (!all_pairs_dict_eligible(all.data(), n_total))  // the same decision on every rank: all hold the same sets
      return sks_all_vs_all(ctx, all.data(), n_total, begin, end, out_counts, out_sizes, out_ani);
    H.mask[0] = all[0]->mask[0];
    H.mask[1] = all[0]->mask[1];
    h_sizes.resize((size_t)n_total);
    for (int64_t i = 0; i < n_total; ++i) h_sizes[(size_t)i] = (int32_t)all[i]->count;
    int st = all_pairs_raw(ctx, all.data(), n_total, 0, n_total, rank, world, false, (int64_t)world * per, &raw, &sizes, &overflow, nullptr);
    if (st == SKS_ERR_CAPACITY)  // too large for the dictionary (decided from global sizes: on every rank alike)
      return sks_all_vs_all(ctx, all.data(), n_total, begin, end, out_counts, out_sizes, out_ani);
    if (st != SKS_OK) return st;
    return finish();
  }
  // owner route: group my keys by owner first (one pass into regions with 25 % slack; the exact two passes if that
  // overflows anywhere), then ONE all-gather carries the set sizes and the group counts of every rank
  uint64_t my_keys_n = 0;
  for (int64_t i = 0; i < n_local; ++i) my_keys_n += (uint64_t)local[i]->count;
  const bool local_ok = n_local == 0 || all_pairs_dict_usable(local[0]->key_words, local[0]->mask, 2, my_keys_n);
  BufferRef r_keys, r_sets, r_ctl, zero_counts, in_keys, in_sets;
  unsigned long long *d_counts = nullptr;
  size_t my_cap = all_pairs_route_cap(my_keys_n, world);
  if (local_ok) {
    SKS_TRY(all_pairs_route(ctx, local, n_local, begin, world, my_cap, &r_keys, &r_sets, &r_ctl, &d_counts));
  } else {  // these sets cannot take the dictionary: say so with counts of ~0 (all ranks then fall back together)
    SKS_TRY(alloc_buffer(ctx, 8 * W, &zero_counts));
    SKS_CUDA_TRY(cudaMemsetAsync(zero_counts->ptr, 0xFF, 8 * W, ctx->stream));
    d_counts = static_cast<unsigned long long *>(zero_counts->ptr);
  }
  {
    KernelTimer timer(ctx, SKS_KERNEL_EXCHANGE);
    SKS_TRY(exchange_headers(ctx, comm, local, n_local, n_total, &H, d_counts, W));
  }
  bool usable = all_pairs_dict_usable(H.kw, H.mask, n_total, H.total_keys), refused = false, spilled = false;
  for (int src = 0; src < world; ++src)
    for (int dst = 0; dst < world; ++dst) {
      const unsigned long long c = (unsigned long long)H.extra(src)[dst];
      if (c == ~0ull) refused = true;
      else if (c > all_pairs_route_cap((uint64_t)H.of(src)[0], world)) spilled = true;
    }
  if (!usable || refused) return by_rows();  // the same decision on every rank
  std::vector<unsigned long long> cnt(W * W);
  for (int src = 0; src < world; ++src)
    for (int dst = 0; dst < world; ++dst) cnt[(size_t)src * W + dst] = (unsigned long long)H.extra(src)[dst];
  if (spilled) {  // somebody's region was too small: everybody groups exactly (count, then scatter) and says so again
    my_cap = 0;
    SKS_TRY(all_pairs_route(ctx, local, n_local, begin, world, 0, &r_keys, &r_sets, &r_ctl, &d_counts));
    BufferRef d_all_buf;
    SKS_TRY(alloc_buffer(ctx, 8 * W * W, &d_all_buf));
    unsigned long long *h_all = nullptr;
    SKS_TRY(ctx_pinned(ctx, 8 * W * W, reinterpret_cast<void **>(&h_all)));
    KernelTimer timer(ctx, SKS_KERNEL_EXCHANGE);
    SKS_NCCL_TRY(nccl()->AllGather(d_counts, d_all_buf->ptr, W, ncclUint64, comm->comm, ctx->stream));
    SKS_CUDA_TRY(cudaMemcpyAsync(h_all, d_all_buf->ptr, 8 * W * W, cudaMemcpyDeviceToHost, ctx->stream));
    SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    cnt.assign(h_all, h_all + W * W);
  }
  auto sends = [&](int src, int dst) { return (size_t)cnt[(size_t)src * W + dst]; };
  size_t n_in = 0;
  std::vector<size_t> in_at(W), out_at(W);
  for (int src = 0; src < world; ++src) {
    in_at[src] = n_in;
    n_in += sends(src, rank);
  }
  {
    size_t at = 0;
    for (int dst = 0; dst < world; ++dst) {
      out_at[dst] = my_cap ? my_cap * (size_t)dst : at;   // fixed regions, or packed groups after the exact passes
      at += sends(rank, dst);
    }
  }
  if (n_in >= (1ull << 31)) return set_error(SKS_ERR_CAPACITY, "too many keys arrive at rank %d", rank);
  SKS_TRY(alloc_buffer(ctx, 8 * std::max<size_t>(n_in, 2), &in_keys));
  SKS_TRY(alloc_buffer(ctx, 2 * std::max<size_t>(n_in, 8), &in_sets));
  {
    KernelTimer timer(ctx, SKS_KERNEL_EXCHANGE);
    const char *sk = static_cast<const char *>(r_keys->ptr), *ss = static_cast<const char *>(r_sets->ptr);
    char *ik = static_cast<char *>(in_keys->ptr), *is = static_cast<char *>(in_sets->ptr);
    SKS_NCCL_TRY(nccl()->GroupStart());
    for (int r = 0; r < world; ++r) {
      if (r == rank) continue;
      if (sends(rank, r)) {
        SKS_NCCL_TRY(nccl()->Send(sk + 8 * out_at[r], 8 * sends(rank, r), ncclUint8, r, comm->comm, ctx->stream));
        SKS_NCCL_TRY(nccl()->Send(ss + 2 * out_at[r], 2 * sends(rank, r), ncclUint8, r, comm->comm, ctx->stream));
      }
      if (sends(r, rank)) {
        SKS_NCCL_TRY(nccl()->Recv(ik + 8 * in_at[r], 8 * sends(r, rank), ncclUint8, r, comm->comm, ctx->stream));
        SKS_NCCL_TRY(nccl()->Recv(is + 2 * in_at[r], 2 * sends(r, rank), ncclUint8, r, comm->comm, ctx->stream));
      }
    }
    SKS_NCCL_TRY(nccl()->GroupEnd());
    if (sends(rank, rank)) {
      SKS_CUDA_TRY(cudaMemcpyAsync(ik + 8 * in_at[rank], sk + 8 * out_at[rank], 8 * sends(rank, rank), cudaMemcpyDeviceToDevice, ctx->stream));
      SKS_CUDA_TRY(cudaMemcpyAsync(is + 2 * in_at[rank], ss + 2 * out_at[rank], 2 * sends(rank, rank), cudaMemcpyDeviceToDevice, ctx->stream));
    }
  }
  sizes_from_headers();
  const FlatKeys flat = {static_cast<const unsigned long long *>(in_keys->ptr), static_cast<const uint16_t *>(in_sets->ptr),
                         (uint32_t)n_in, h_sizes.data()};
  SKS_TRY(all_pairs_raw(ctx, nullptr, n_total, 0, n_total, 0, 1, false, (int64_t)world * per, &raw, &sizes, &overflow, &flat));
  return finish();
}

int sks_sketch_sequence_sharded(sks_ctx *ctx, sks_comm *comm, const sks_batch *slice, const uint64_t mask[2], int window,
                                const sks_pred *pred, int gather, sks_set **out, int64_t *out_global_size) {
  if (!ctx || !slice || !mask || !pred || !out) return set_error(SKS_ERR_INVALID, "null argument");
  if (sks_batch_n_genomes(slice) != 1) return set_error(SKS_ERR_INVALID, "a sequence slice is a single-genome batch");
  const int world = comm ? comm->world : 1, rank = comm ? comm->rank : 0;
  if (world > 64) return set_error(SKS_ERR_INVALID, "at most 64 ranks");
  if (world == 1) {
    sks_set *loc = nullptr;
    SKS_TRY(sks_sketch(ctx, slice, mask, window, pred, SKS_REPR_SORTED, &loc));
    *out = loc;
    if (out_global_size) *out_global_size = loc->count;
    return SKS_OK;
  }
  SKS_TRY(need_nccl());
  DeviceGuard guard(ctx->device);
  // the slice's kept k-mers as the sketch kernel leaves them (unsorted, duplicates included): they are sorted once,
  // by the rank that owns their key range
  BufferRef raw;
  uint64_t n_raw = 0;
  int kw = 1;
  SKS_TRY(sketch_raw_one(ctx, slice, mask, window, pred, &raw, &n_raw, &kw));
  const size_t kb = (size_t)kw * 8;
  const int weight = sks_mask_weight(mask);
  // key range r = [split[r], split[r + 1]): equal shares of the 4^weight possible keys under the mask
  const size_t W = (size_t)world;
  unsigned long long *h_tab = nullptr;
  SKS_TRY(ctx_pinned(ctx, 8 * (2 * W + W * W + W + 8), reinterpret_cast<void **>(&h_tab)));
  BufferRef d_tab, routed;
  SKS_TRY(alloc_buffer(ctx, 8 * (2 * W + 3 * W + W * W + 8), &d_tab));
  unsigned long long *d_split = static_cast<unsigned long long *>(d_tab->ptr), *d_cnt = d_split + 2 * W, *d_off = d_cnt + W,
                     *d_cur = d_off + W, *d_all = d_cur + W;
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
  SKS_CUDA_TRY(cudaMemcpyAsync(d_split, h_tab, 16 * W, cudaMemcpyHostToDevice, ctx->stream));
  SKS_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 8 * W, ctx->stream));
  SKS_TRY(alloc_buffer(ctx, std::max<size_t>((size_t)n_raw * kb, 16), &routed));
  const unsigned long long *keys = static_cast<const unsigned long long *>(raw->ptr);
  unsigned long long *d_routed = static_cast<unsigned long long *>(routed->ptr);
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_raw + 255) / 256, (uint64_t)ctx->sm_count * 8));
  {
    KernelTimer timer(ctx, SKS_KERNEL_SORT_UNIQUE);
    if (kw == 1) {
      route_kernel<1, false><<<grid, 256, 0, ctx->stream>>>(keys, n_raw, d_split, world, d_cnt, nullptr, nullptr, nullptr);
      route_offsets_kernel<<<1, 32, 0, ctx->stream>>>(d_cnt, world, d_off, d_cur);
      route_kernel<1, true><<<grid, 256, 0, ctx->stream>>>(keys, n_raw, d_split, world, nullptr, d_off, d_cur, d_routed);
    } else {
      route_kernel<2, false><<<grid, 256, 0, ctx->stream>>>(keys, n_raw, d_split, world, d_cnt, nullptr, nullptr, nullptr);
      route_offsets_kernel<<<1, 32, 0, ctx->stream>>>(d_cnt, world, d_off, d_cur);
      route_kernel<2, true><<<grid, 256, 0, ctx->stream>>>(keys, n_raw, d_split, world, nullptr, d_off, d_cur, d_routed);
    }
    SKS_CUDA_TRY(cudaGetLastError());
    ctx->launches += 3;
  }
  // every rank's counts per range to every rank (row r of the table = what rank r sends where)
  unsigned long long *h_all = h_tab + 2 * W;
  {
    KernelTimer timer(ctx, SKS_KERNEL_EXCHANGE);
    SKS_NCCL_TRY(nccl()->AllGather(d_cnt, d_all, W, ncclUint64, comm->comm, ctx->stream));
    SKS_CUDA_TRY(cudaMemcpyAsync(h_all, d_all, 8 * W * W, cudaMemcpyDeviceToHost, ctx->stream));
    SKS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  // (the pinned ring may hand the staging area out again inside the calls below: keep what is needed)
  const std::vector<unsigned long long> cnt(h_all, h_all + W * W);
  auto sends = [&](int src, int dst) { return (int64_t)cnt[(size_t)src * W + dst]; };
  // inbox: the keys of my range from every rank (mine by a device copy)
  int64_t n_in = 0;
  std::vector<int64_t> in_at(world), out_at(world);
  for (int src = 0; src < world; ++src) {
    in_at[src] = n_in;
    n_in += sends(src, rank);
  }
  {
    int64_t at = 0;
    for (int dst = 0; dst < world; ++dst) {
      out_at[dst] = at;
      at += sends(rank, dst);
    }
  }
  BufferRef inbox;
  SKS_TRY(alloc_buffer(ctx, (size_t)std::max<int64_t>(n_in, 1) * kb, &inbox));
  {
    KernelTimer timer(ctx, SKS_KERNEL_EXCHANGE);
    SKS_NCCL_TRY(nccl()->GroupStart());
    for (int r = 0; r < world; ++r) {
      if (r == rank) continue;
      if (sends(rank, r) > 0)
        SKS_NCCL_TRY(nccl()->Send(reinterpret_cast<const char *>(d_routed) + (size_t)out_at[r] * kb, (size_t)sends(rank, r) * kb,
                                  ncclUint8, r, comm->comm, ctx->stream));
      if (sends(r, rank) > 0)
        SKS_NCCL_TRY(nccl()->Recv(static_cast<char *>(inbox->ptr) + (size_t)in_at[r] * kb, (size_t)sends(r, rank) * kb, ncclUint8, r,
                                  comm->comm, ctx->stream));
    }
    SKS_NCCL_TRY(nccl()->GroupEnd());
    if (sends(rank, rank) > 0)
      SKS_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(inbox->ptr) + (size_t)in_at[rank] * kb,
                                   reinterpret_cast<const char *>(d_routed) + (size_t)out_at[rank] * kb, (size_t)sends(rank, rank) * kb,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
  }
  unsigned long long *h_cnt = nullptr;
  BufferRef d_cntbuf;
  SKS_TRY(alloc_buffer(ctx, 8 * (W + 1), &d_cntbuf));
  unsigned long long *d_mycnt = static_cast<unsigned long long *>(d_cntbuf->ptr), *d_cnts = d_mycnt + 1;
  // my range of the global set: sort + unique of what arrived (the same k-mer can occur in several slices)
  sks_set *mine = nullptr;
  {
    const uint64_t off = 0, cnt_in = (uint64_t)n_in;
    std::vector<uint64_t> uoff, ucount;
    BufferRef uniq;
    // a power-of-two world cuts the key space at bit boundaries: the top log2(world) mask bits are the same in all of
    // my keys, and the bucket sort indexes by the bits below them
    int skip = 0;
    if ((world & (world - 1)) == 0)
      while ((1 << skip) < world) ++skip;
    SKS_TRY(sort_unique_regions(ctx, kw, inbox->ptr, &off, &cnt_in, 1, cnt_in, &uniq, &uoff, &ucount, mask, skip));
    mine = new_set(ctx, SKS_REPR_SORTED, mask, window, weight);
    if (!mine) return set_error(SKS_ERR_INVALID, "out of host memory");
    mine->buf = uniq;
    mine->key_words = kw;
    mine->byte_off = (size_t)uoff[0] * kb;
    mine->count = (int64_t)ucount[0];
  }
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
