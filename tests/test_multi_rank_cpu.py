"""World-size-2 (gloo, CPU) coverage of the N > 1 host side: the rendezvous that carries rank 0's NCCL id to the other
ranks (spaced_kmer_sketching_b200/multi_gpu.py), the shard arithmetic the C ABI and the launcher agree on
(sks_shard_range, position_shard), and the contract of the sharded all-vs-all -- every rank returns its own complete
block rows, stacked in rank order they are the matrix a single process computes.  The device work of a rank is done
by the CPU oracle here (libsks has no CPU path); the GPU run of the same contract is tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from spaced_kmer_sketching_b200 import engine, multi_gpu


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


def test_genome_shards_partition_and_match_the_c_abi():
    for n, world in ((1000, 8), (7, 2), (3, 4), (0, 2), (10, 3), (1, 1)):
        spans = [multi_gpu.genome_shard(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        per = -(-n // world)
        assert spans == [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]
        assert spans == [engine.shard_range(n, r, world) for r in range(world)]


def test_comm_id_is_made_without_a_gpu():
    a, b = engine.comm_unique_id(), engine.comm_unique_id()
    assert len(a) == engine.COMM_ID_BYTES == 128 and a != b


def test_torch_still_imports_after_the_first_communicator_call():
    """libsks binds NCCL at run time; a process that touches a communicator entry point before it imports torch must not
    end up with the system's older libnccl where libtorch_cuda.so needs the wheel's (engine -> _lib.preload_nccl)."""
    import subprocess
    import sys
    code = ("from spaced_kmer_sketching_b200 import engine\n"
            "assert len(engine.comm_unique_id()) == 128\n"
            "import torch\n"
            "import torch.distributed\n"
            "print('ok', torch.__version__)\n")
    env = {k: v for k, v in os.environ.items() if k != "SKS_NCCL_LIB"}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


N_GENOMES, L = 5, 20_000


def _genome(port, g):
    base = port.gen(L, 1000)
    return base if g == 0 else port.mutate(base, 2000 + g, [0, 1000, 200, 100, 50][g])


def _worker(rank, world, port_no, q):
    from oracle import port
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm_id = multi_gpu.exchange_comm_id(rank, world, dist)     # the launcher's part of sks_comm_init_rank
        # the sharded all-vs-all with the oracle standing in for the device: sketch the rank's block, exchange the
        # sketches (sks_comm_allgather_sets does this over NCCL), fill the rank's block rows
        mask, w = port.seed_to_mask("0011111011010111111011001011101")
        lo, hi = multi_gpu.genome_shard(N_GENOMES, rank, world)
        mine = [port.sketch_set(_genome(port, g), [L], mask, w, port.FMH, 1, 20, 181) for g in range(lo, hi)]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        sets = [s for part in gathered for s in part]
        assert len(sets) == N_GENOMES
        rows = np.array([[port.intersection(sets[i], sets[j]) for j in range(N_GENOMES)] for i in range(lo, hi)],
                        dtype=np.int32).reshape(hi - lo, N_GENOMES)
        q.put((rank, comm_id, (lo, hi), rows.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_rendezvous_and_block_rows_match_single_process():
    from oracle import port
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][1] == results[1][1] and len(results[0][1]) == 128      # both ranks hold rank 0's id
    assert [r[2] for r in results] == [(0, 3), (3, 5)]
    mask, w = port.seed_to_mask("0011111011010111111011001011101")
    sets = [port.sketch_set(_genome(port, g), [L], mask, w, port.FMH, 1, 20, 181) for g in range(N_GENOMES)]
    want = [[port.intersection(a, b) for b in sets] for a in sets]
    assert results[0][3] + results[1][3] == want
