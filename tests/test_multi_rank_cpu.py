"""World-size-2 (gloo, CPU) coverage of the N > 1 host logic in spaced_kmer_sketching_b200/multi_gpu.py:
shard arithmetic, the variable-length all-gather, the rank-tiled pair matrix and its gather.  The
"device" work of each rank is done by the CPU oracle here, so the only thing under test is the sharding
and exchange plumbing that the GPU ranks use unchanged (over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spaced_kmer_sketching_b200 import multi_gpu


def test_position_shards_cover_every_window_once():
    for n_bases, w, world in ((250_000_000, 31, 8), (1000, 31, 4), (17, 8, 2), (5, 8, 2), (100_003, 24, 3)):
        starts = []
        for r in range(world):
            first, count = multi_gpu.position_shard(n_bases, w, r, world)
            assert first % 16 == 0
            starts.extend(range(first, first + count) if count < 10_000 else [first, first + count - 1])
            if count >= 10_000:
                nxt = multi_gpu.position_shard(n_bases, w, r + 1, world)[0] if r + 1 < world else None
                assert nxt is None or nxt in (first + count, 0)
        total = sum(multi_gpu.position_shard(n_bases, w, r, world)[1] for r in range(world))
        assert total == max(n_bases - w + 1, 0)
        if n_bases < 10_000:
            assert starts == list(range(max(n_bases - w + 1, 0)))


def test_genome_and_row_tiles_partition():
    for n, world in ((1000, 8), (7, 2), (3, 4), (0, 2)):
        spans = [multi_gpu.genome_shard(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert spans == [multi_gpu.row_tile(n, r, world) for r in range(world)]


def test_block_rects_cover_every_unordered_pair_once():
    for world in range(1, 10):
        for n in (0, 1, 5, 16, 37):
            seen = np.zeros((n, n), dtype=np.int32)
            work = []
            for r in range(world):
                rows = multi_gpu.row_tile(n, r, world)
                cells = 0
                for (r0, r1), (c0, c1) in multi_gpu.block_rects(n, r, world):
                    assert rows[0] <= r0 and r1 <= rows[1]       # results stay in the rank's own block row
                    seen[r0:r1, c0:c1] += 1
                    cells += (r1 - r0) * (c1 - c0)
                work.append(cells)
            both = seen + seen.T - np.diag(np.diag(seen))
            off = ~np.eye(n, dtype=bool)
            # every unordered pair once; within a diagonal block both orientations appear (one is evaluated, the
            # C side mirrors it), so count those through `both` = 2
            for i in range(n):
                for j in range(i + 1, n):
                    same_block = any(multi_gpu.row_tile(n, r, world)[0] <= i < multi_gpu.row_tile(n, r, world)[1] and
                                     multi_gpu.row_tile(n, r, world)[0] <= j < multi_gpu.row_tile(n, r, world)[1]
                                     for r in range(world))
                    assert both[i, j] == (2 if same_block else 1), (world, n, i, j)
            assert (np.diag(seen) == 1).all()
            if n >= 4 * world:
                assert max(work) <= 1.35 * (sum(work) / world) + n, (world, n, work)
    # mirror: entries no rank evaluated come from the transpose
    m = np.array([[5, -1, 2], [1, 6, -1], [-1, 3, 7]], dtype=np.int32)
    assert multi_gpu.mirror_counts(m).tolist() == [[5, 1, 2], [1, 6, 3], [2, 3, 7]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, q):
    from oracle import port
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # variable-length all-gather, incl. an empty contribution
        local = torch.arange(5 * rank, dtype=torch.int64) + 100 * rank
        parts = multi_gpu.allgather_varlen(local, world, dist)
        assert [p.tolist() for p in parts] == [(torch.arange(5 * r, dtype=torch.int64) + 100 * r).tolist() for r in range(world)]

        # sharded all-vs-all with the oracle standing in for the device
        n_genomes, L = 5, 20_000
        mask, w = port.seed_to_mask("0011111011010111111011001011101")
        base = port.gen(L, 1000)
        genome = lambda g: base if g == 0 else port.mutate(base, 2000 + g, [0, 1000, 200, 100, 50][g])
        lo, hi = multi_gpu.genome_shard(n_genomes, rank, world)
        mine = [port.sketch_set(genome(g), [L], mask, w, port.FMH, 1, 20, 181) for g in range(lo, hi)]
        keys = torch.from_numpy(np.concatenate([m[:, 0] for m in mine]).astype(np.int64)) if mine else torch.zeros(0, dtype=torch.int64)
        counts = torch.tensor([len(m) for m in mine], dtype=torch.int64)
        all_counts = multi_gpu.allgather_varlen(counts, world, dist)
        all_keys = multi_gpu.allgather_varlen(keys, world, dist)
        sets = []
        for r in range(world):
            off = 0
            for c in all_counts[r].tolist():
                k = all_keys[r][off:off + c].numpy().astype(np.uint64)
                sets.append(np.stack([k, np.zeros_like(k)], axis=1))
                off += c
        assert len(sets) == n_genomes
        rows = multi_gpu.row_tile(n_genomes, rank, world)
        mat = np.zeros((n_genomes, n_genomes), dtype=np.int32)
        for i in range(rows[0], rows[1]):
            for j in range(n_genomes):
                mat[i, j] = port.intersection(sets[i], sets[j])
        full = multi_gpu.gather_rows(mat, rows, world)
        # the block exchange of the tiled evaluation: keep only what block_rects assigns to this rank, swap blocks
        # point to point, and the rank's rows must come out complete
        share = np.full((n_genomes, n_genomes), -1, dtype=np.int32)
        for (r0, r1), (c0, c1) in multi_gpu.block_rects(n_genomes, rank, world):
            share[r0:r1, c0:c1] = mat[r0:r1, c0:c1]
        for i in range(rows[0], rows[1]):      # a diagonal block is mirrored inside by sks_intersect_block
            for j in range(rows[0], rows[1]):
                share[i, j] = mat[i, j]
        got = multi_gpu.exchange_blocks(share, rank, world)
        assert np.array_equal(got, full[rows[0]:rows[1]]), (rank, got.tolist())
        q.put((rank, full.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_all_vs_all_matches_single_process():
    from oracle import port
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_genomes, L = 5, 20_000
    mask, w = port.seed_to_mask("0011111011010111111011001011101")
    base = port.gen(L, 1000)
    sets = [port.sketch_set(base if g == 0 else port.mutate(base, 2000 + g, [0, 1000, 200, 100, 50][g]), [L], mask, w,
                            port.FMH, 1, 20, 181) for g in range(n_genomes)]
    want = [[port.intersection(a, b) for b in sets] for a in sets]
    assert results[0] == want and results[1] == want
