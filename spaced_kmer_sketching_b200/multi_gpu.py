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

from typing import List, Sequence, Tuple

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


def allgather_varlen(local, world: int, dist=None):
    """All-gather of 1-D tensors of different lengths: returns the list of every rank's tensor.
    One small all-gather of lengths, one padded all-gather of payload."""
    import torch
    if world == 1:
        return [local]
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    lens = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(lens, n)
    lens = [int(x.item()) for x in lens]
    cap = max(max(lens), 1)
    padded = torch.zeros(cap, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = torch.empty(world * cap, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded) if hasattr(dist, "all_gather_into_tensor") and local.is_cuda else \
        dist.all_gather(list(out.view(world, cap).unbind(0)), padded)
    return [out.view(world, cap)[r, : lens[r]] for r in range(world)]


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
    all_counts = allgather_varlen(cnt, world, dist)
    all_keys = allgather_varlen(keys, world, dist)
    torch.cuda.current_stream().synchronize()   # the gathered buffers are read by libsks on the same stream
    out = []
    for r in range(world):   # one device copy + one buffer per source rank
        out.extend(ctx.sets_from_device_keys(all_keys[r].data_ptr(), all_counts[r].tolist(), kw, mask, window))
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
    counts = np.zeros((n, n), dtype=np.int32)
    ctx.intersect_all_pairs(sets, rows[0], rows[1], counts)
    counts = gather_rows(counts, rows, world)
    sizes = np.array([s.kmer_set_size() for s in sets], dtype=np.int32)
    ani = engine.ani_from_counts(counts.ravel(), np.repeat(sizes, n), engine.mask_weight(mask)).reshape(n, n)
    return counts, sizes, ani
