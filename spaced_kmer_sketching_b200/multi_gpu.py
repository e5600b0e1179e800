"""One process per GPU: sharding of the sketch-and-compare path over the GPUs of one node.

The reference's only parallelism is `cilk_for` over files (src/kmer_set.cpp:124-131) and over set pairs
(src/kmer_set.cpp:179-182).  Here:

  * many genomes  : genome g belongs to rank g // ceil(n / world); every rank sketches its shard, the
                    sketches (sorted distinct keys) are exchanged with ONE NCCL all-gather over NVLink,
                    and the n x n pair matrix (generate_all_pairs_from_vector order, src/generators.hpp:44-58)
                    is tiled by contiguous row blocks; the int counts are gathered back.
  * one long sequence (C3): window starts are split into `world` contiguous ranges, each rank gets its
                    range plus a (w-1)-base halo; the global set is the sort-unique of the gathered keys.

torch.distributed is plumbing only (rendezvous, NCCL all-gather); the compute is libsks.so.  The
helpers that do not touch the device (`position_shard`, `row_tile`, `allgather_varlen`) run under
gloo on CPU tensors, which is how tests/test_multi_rank_cpu.py covers the N > 1 logic.
"""
from __future__ import annotations

from typing import Optional, List, Sequence, Tuple

import numpy as np


def position_shard(n_bases: int, window: int, rank: int, world: int) -> Tuple[int, int]:
    """(first window start, number of window starts) of `rank`; starts are multiples of 16 bases."""
    n_starts = max(n_bases - window + 1, 0)
    per = -(-n_starts // world)
    per = -(-per // 16) * 16
    first = rank * per
    count = max(min(per, n_starts - first), 0)
    return (first, count) if count else (0, 0)


def genome_shard(n_genomes: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the genomes owned by `rank` (contiguous blocks)."""
    per = -(-n_genomes // world)
    return min(rank * per, n_genomes), min((rank + 1) * per, n_genomes)


def row_tile(n_sets: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the pair-matrix rows computed by `rank`."""
    return genome_shard(n_sets, rank, world)


def _all_gather_flat(dist, out, inp, world: int):
    """out[world * n] <- every rank's inp[n]; NCCL takes the fused form, gloo the list form."""
    if inp.is_cuda and hasattr(dist, "all_gather_into_tensor"):
        dist.all_gather_into_tensor(out, inp)
    else:
        dist.all_gather(list(out.view(world, -1).unbind(0)), inp)


def block_rects(n_sets: int, rank: int, world: int):
    """Rectangles ((row_begin, row_end), (col_begin, col_end)) of the n x n pair matrix that `rank` evaluates, all
    inside its own block row: the cyclic half of the column blocks r, r+1, ..., so that every unordered pair
    of blocks {r, c} is evaluated exactly once (|A n B| is symmetric; the mirror entries are filled in after
    the gather).  For an even world the opposite block {r, r + world/2} is split between its two ranks: the
    lower one takes the first half of its own rows against the whole block, the upper one the other half of
    those rows as columns."""
    rows = row_tile(n_sets, rank, world)
    half = world // 2
    rects = [(rows, row_tile(n_sets, (rank + k) % world, world)) for k in range(half + (world % 2))]
    if world % 2 == 0 and world > 1:
        if rank < half:
            mid = rows[0] + (rows[1] - rows[0] + 1) // 2
            rects.append(((rows[0], mid), row_tile(n_sets, rank + half, world)))
        else:
            other = row_tile(n_sets, rank - half, world)
            mid = other[0] + (other[1] - other[0] + 1) // 2
            rects.append((rows, (mid, other[1])))
    return [r for r in rects if r[0][1] > r[0][0] and r[1][1] > r[1][0]]


def tiled_counts(ctx, sets: Sequence, rank: int, world: int, out: Optional[np.ndarray] = None) -> np.ndarray:
    """This rank's share of the n x n intersection counts (entries not evaluated here are -1).  `out`: an [n, n]
    int32 array to reuse -- a fresh 4 MB matrix costs more in page faults (1-1.6 ms at n = 1000) than the host side
    of the whole intersection call."""
    n = len(sets)
    if out is None or out.shape != (n, n) or out.dtype != np.int32 or not out.flags.c_contiguous:
        out = np.empty((n, n), dtype=np.int32)
    out.fill(-1)
    ctx.intersect_rects(sets, block_rects(n, rank, world), out)   # one pair table, one launch
    return out


def mirror_counts(counts: np.ndarray) -> np.ndarray:
    """Fills the entries no rank evaluated (-1) from their transposes."""
    return np.where(counts < 0, counts.T, counts)


def exchange_blocks(counts: np.ndarray, rank: int, world: int) -> np.ndarray:
    """From this rank's share of the counts (tiled_counts: entries not evaluated here are -1) to its complete block
    rows of the mirrored matrix, without assembling the whole matrix anywhere: every rank sends rank c the block
    (own rows x c's rows) and fills its own gaps from the transposes it receives.  Point-to-point sends
    (batch_isend_irecv): world - 1 blocks of n^2 / world^2 int32 per rank instead of an all-gather of n^2."""
    n = counts.shape[0]
    rows = row_tile(n, rank, world)
    mine = counts[rows[0]:rows[1]].copy()
    if world == 1:
        return mirror_rows(counts, rows)
    import torch
    import torch.distributed as dist
    on_gpu = dist.get_backend() == "nccl"
    # one host<->device copy each way: the rank's rows go up once, the received blocks come back in one buffer
    src = torch.from_numpy(mine)
    if on_gpu:
        src = src.cuda()
    peers = [c for c in range(world) if c != rank and row_tile(n, c, world)[1] > row_tile(n, c, world)[0]]
    if rows[1] == rows[0]:
        peers = []
    n_rows = rows[1] - rows[0]
    sizes = [(row_tile(n, c, world)[1] - row_tile(n, c, world)[0]) * n_rows for c in peers]
    inbox = torch.empty(sum(sizes), dtype=torch.int32, device=src.device)
    ops, keep, at = [], [], 0
    for c, size in zip(peers, sizes):
        cols = row_tile(n, c, world)
        send = src[:, cols[0]:cols[1]].contiguous()
        keep.append(send)
        ops += [dist.P2POp(dist.isend, send, c), dist.P2POp(dist.irecv, inbox[at:at + size], c)]
        at += size
    if ops:
        for work in dist.batch_isend_irecv(ops):
            work.wait()
    got, at = inbox.cpu().numpy(), 0
    for c, size in zip(peers, sizes):
        cols = row_tile(n, c, world)
        theirs = got[at:at + size].reshape(cols[1] - cols[0], n_rows).T
        at += size
        block = mine[:, cols[0]:cols[1]]
        mine[:, cols[0]:cols[1]] = np.where(block < 0, theirs, block)
    return mine


def mirror_rows(counts: np.ndarray, rows: Tuple[int, int]) -> np.ndarray:
    """The block rows [rows) of the mirrored matrix only (what one rank needs for the ANI of its own rows)."""
    block = counts[rows[0]:rows[1]]
    return np.where(block < 0, counts[:, rows[0]:rows[1]].T, block)


def allgather_varlen_many(locals_, world: int, dist=None):
    """All-gather of several 1-D tensors of rank-dependent lengths with ONE host synchronisation: a small
    all-gather of all the lengths, then one padded all-gather per tensor.  Returns, per input tensor, the list
    of every rank's contribution."""
    import torch
    if world == 1:
        return [[t] for t in locals_]
    dev = locals_[0].device
    n = torch.tensor([t.numel() for t in locals_], dtype=torch.int64, device=dev)
    lens = torch.empty(world * len(locals_), dtype=torch.int64, device=dev)
    _all_gather_flat(dist, lens, n, world)
    lens = lens.view(world, len(locals_)).cpu()      # the one sync
    out = []
    for k, local in enumerate(locals_):
        cap = max(int(lens[:, k].max()), 1)
        if local.numel() == cap:
            padded = local.contiguous()
        else:
            padded = torch.zeros(cap, dtype=local.dtype, device=dev)
            padded[: local.numel()] = local
        buf = torch.empty(world * cap, dtype=local.dtype, device=dev)
        _all_gather_flat(dist, buf, padded, world)
        out.append([buf.view(world, cap)[r, : int(lens[r, k])] for r in range(world)])
    return out


def allgather_varlen(local, world: int, dist=None):
    """All-gather of 1-D tensors of different lengths: returns the list of every rank's tensor."""
    return allgather_varlen_many([local], world, dist)[0]


class _DevPtr:
    """Zero-copy torch view of a device pointer owned by libsks (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def keys_as_tensor(sets: Sequence, torch):
    """The keys of `sets` (all from one sketch call) as one int64 device tensor + per-set key counts."""
    metas = [s.device_keys() for s in sets]
    kw = metas[0][2] if metas else 1
    counts = [m[1] for m in metas]
    contiguous = all(metas[i + 1][0] == metas[i][0] + metas[i][1] * 8 * kw or metas[i + 1][1] == 0 or metas[i][1] == 0
                     for i in range(len(metas) - 1))
    total = sum(counts) * kw
    if total == 0:
        return torch.zeros(0, dtype=torch.int64, device="cuda"), counts, kw
    if contiguous and all(c > 0 for c in counts):
        return torch.as_tensor(_DevPtr(metas[0][0], total), device="cuda"), counts, kw
    parts = [torch.as_tensor(_DevPtr(m[0], m[1] * kw), device="cuda") for m in metas if m[1] > 0]
    return torch.cat(parts), counts, kw


def allgather_sets(ctx, local_sets: Sequence, mask: int, window: int, rank: int, world: int, stream=None) -> List:
    """Every rank ends up with the sets of ALL ranks, in global genome order."""
    if world == 1:
        return list(local_sets)
    import torch
    import torch.distributed as dist
    keys, counts, kw = keys_as_tensor(local_sets, torch)
    cnt = torch.tensor(counts, dtype=torch.int64, device="cuda")
    all_counts, all_keys = allgather_varlen_many([cnt, keys], world, dist)   # one length sync, two NCCL all-gathers
    counts_host = torch.cat(all_counts).cpu()    # syncs the stream: the gathered buffers are complete
    out, at = [], 0
    for r in range(world):   # one device copy + one buffer per source rank
        n_r = all_counts[r].numel()
        out.extend(ctx.sets_from_device_keys(all_keys[r].data_ptr(), counts_host[at:at + n_r].tolist(), kw, mask, window))
        at += n_r
    return out


def gather_rows(counts: np.ndarray, rows: Tuple[int, int], world: int) -> np.ndarray:
    """Collects every rank's row block of the n x n count matrix (all ranks get the full matrix)."""
    if world == 1:
        return counts
    import torch
    import torch.distributed as dist
    n = counts.shape[0]
    block = torch.from_numpy(np.ascontiguousarray(counts[rows[0]:rows[1]])).reshape(-1)
    if dist.get_backend() == "nccl":
        block = block.cuda()
    parts = allgather_varlen(block, world, dist)
    full = torch.cat(parts).reshape(n, n)
    return full.cpu().numpy()


def synth_slice(ctx, n_bases_total: int, gen_seed: int, shard: Tuple[int, int], window: int = 0):
    """Bases [first, first + count + window - 1) of gen(n_bases_total, gen_seed), generated on the device."""
    first, count = shard
    n = 0 if count == 0 else min(count + max(window - 1, 0), n_bases_total - first)
    return ctx.synth_at(n, [first], [gen_seed], [0], [0])


def all_vs_all(ctx, local_batch, mask: int, window: int, pred, rank: int, world: int):
    """Sharded all-vs-all: returns (counts[n, n] int32, sizes[n] int32, ani[n, n] float64)."""
    from . import engine
    local_sets = ctx.sketch(local_batch, mask, window, pred, engine.REPR_SORTED)
    sets = allgather_sets(ctx, local_sets, mask, window, rank, world)
    n = len(sets)
    rows = row_tile(n, rank, world)
    counts = mirror_counts(gather_rows(tiled_counts(ctx, sets, rank, world), rows, world))
    sizes = np.array([s.kmer_set_size() for s in sets], dtype=np.int32)
    ani = engine.ani_from_counts(counts.ravel(), np.repeat(sizes, n), engine.mask_weight(mask)).reshape(n, n)
    return counts, sizes, ani
