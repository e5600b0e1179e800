"""One process per GPU: rendezvous for the sharded entry points of the C ABI (include/sks.h, `sks_comm_*`).

The sharding itself -- the exchange of sketches, the rank's block rows of the pair matrix, the key-range routing of a
position-split sequence -- lives in libsks.so (csrc/sks_comm.cu) and is driven through `Context.all_vs_all_sharded`,
`Context.all_vs_all_from_host` and `Context.sketch_sequence_sharded`.  What is left here is what a launcher has to do:
carry rank 0's NCCL id to the other ranks (over whatever process group torchrun set up: NCCL on GPUs, gloo in the CPU
tests) and the shard arithmetic both sides agree on.

The reference's only parallelism is `cilk_for` over files (src/kmer_set.cpp:124-131) and over set pairs
(src/kmer_set.cpp:179-182); genome g of n belongs to rank g // ceil(n / world), and that rank returns rows
[begin, end) of the n x n matrix in generate_all_pairs_from_vector order (src/generators.hpp:44-58).
"""
from __future__ import annotations

from typing import Optional, Tuple

from . import engine


def genome_shard(n_genomes: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the genomes (and of the pair-matrix rows) owned by `rank`: sks_shard_range."""
    return engine.shard_range(n_genomes, rank, world)


def position_shard(n_bases: int, window: int, rank: int, world: int) -> Tuple[int, int]:
    """(first window start, number of window starts) of `rank` in one long sequence; starts are multiples of 16 bases
    (one packed word), the rank reads a (window - 1)-base halo beyond its last start."""
    n_starts = max(n_bases - window + 1, 0)
    per = -(-n_starts // world)
    per = -(-per // 16) * 16
    first = rank * per
    count = max(min(per, n_starts - first), 0)
    return (first, count) if count else (0, 0)


def synth_slice(ctx, n_bases_total: int, gen_seed: int, shard: Tuple[int, int], window: int = 0):
    """Bases [first, first + count + window - 1) of gen(n_bases_total, gen_seed), generated on the device."""
    first, count = shard
    n = 0 if count == 0 else min(count + max(window - 1, 0), n_bases_total - first)
    return ctx.synth_at(n, [first], [gen_seed], [0], [0])


def exchange_comm_id(rank: int, world: int, dist=None, device: Optional[str] = None) -> bytes:
    """Rank 0 makes the 128-byte id (sks_comm_unique_id), every rank returns it.  `dist`: an initialised
    torch.distributed module (any backend); the id travels as a uint8 tensor on `device` ("cuda" under NCCL)."""
    if world == 1:
        return engine.comm_unique_id()
    import torch
    if dist is None:
        import torch.distributed as dist
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros(engine.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        buf = torch.frombuffer(bytearray(engine.comm_unique_id()), dtype=torch.uint8).clone()
    buf = buf.to(device)
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def init_comm(ctx, rank: int, world: int, dist=None):
    """The rank's communicator (None for a single rank)."""
    if world == 1:
        return None
    return engine.Comm(ctx, exchange_comm_id(rank, world, dist), rank, world)
