"""GPU tests of the sharded entry points of the C ABI (include/sks.h, sks_comm_* / *_sharded): the rank's block rows of
the all-vs-all matrix and the key-range routed sketch of one long sequence, against the single-GPU result and the
oracle.  With several GPUs visible the ranks are threads of this process, one per GPU (sks_comm_init_all); on a
one-GPU box the same entry points run with a single rank."""
import threading

import numpy as np
import pytest

import spaced_kmer_sketching_b200 as sks
from spaced_kmer_sketching_b200 import _lib, multi_gpu
from oracle import port

pytestmark = pytest.mark.gpu

C3_SEED = "0011111011010111111011001011101"
DS = [0, 1000, 200, 100, 50, 20]


def _world():
    return max(1, min(int(_lib.load().sks_device_count()), 4))


def _run_ranks(world, fn):
    """fn(rank, ctx, comm) on one thread per GPU; returns the results in rank order."""
    ctxs = [sks.Context(r) for r in range(world)]
    comms = sks.Comm.init_all(ctxs) if world > 1 else [None]
    out, err = [None] * world, []

    def body(r):
        try:
            out[r] = fn(r, ctxs[r], comms[r])
        except Exception as e:   # noqa: BLE001 -- reported by the main thread
            err.append((r, repr(e)))

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    for c in comms:
        if c is not None:
            c.close()
    for c in ctxs:
        c.close()
    assert not err, err
    return out


@pytest.mark.parametrize("n_total", [11, 3])
def test_sharded_all_vs_all_rows_match_single_gpu_and_oracle(n_total):
    world = _world()
    L = 300_000
    mask, w = sks.seed_to_mask(C3_SEED)
    pred = sks.frac_min_hash(1, 20)
    ids = list(range(n_total))

    def rank_body(rank, ctx, comm):
        b, e = sks.shard_range(n_total, rank, world)
        batch = ctx.synth(L, [1000] * (e - b), [2000 + g for g in ids[b:e]], [DS[g % 6] for g in ids[b:e]])
        sets = ctx.sketch(batch, mask, w, pred, sks.REPR_SORTED) if e > b else []
        counts, sizes, ani = ctx.all_vs_all_sharded(comm, sets, n_total)
        everything = ctx.allgather_sets(comm, sets, n_total)
        digest = [x.keys()[:, 0].tolist() for x in everything] if rank == 0 else None
        for x in everything + list(sets):
            x.close()
        return (b, e), counts, sizes, ani, digest

    res = _run_ranks(world, rank_body)
    assert [r[0] for r in res] == [sks.shard_range(n_total, r, world) for r in range(world)]
    counts = np.concatenate([r[1] for r in res])
    ani = np.concatenate([r[3] for r in res])
    # single GPU, all genomes
    ctx = sks.Context(0)
    batch = ctx.synth(L, [1000] * n_total, [2000 + g for g in ids], [DS[g % 6] for g in ids])
    sets = ctx.sketch(batch, mask, w, pred, sks.REPR_SORTED)
    want = ctx.intersect_block(sets, (0, n_total), (0, n_total), np.full((n_total, n_total), -1, dtype=np.int32))
    sizes = np.array([s.kmer_set_size() for s in sets], dtype=np.int32)
    assert np.array_equal(counts, want)
    for r in res:
        assert np.array_equal(r[2], sizes)
    assert res[0][4] == [s.keys()[:, 0].tolist() for s in sets]     # the gathered sets are the sets, in genome order
    wani = sks.ani_from_counts(want.ravel(), np.repeat(sizes, n_total), sks.mask_weight(mask)).reshape(n_total, n_total)
    assert np.max(np.abs(ani - wani)) <= 1e-12
    # oracle: two genomes from the far ends of the shard order
    base = port.gen(L, 1000)
    for g in (0, n_total - 1):
        codes = base if DS[g % 6] == 0 else port.mutate(base, 2000 + g, DS[g % 6])
        assert np.array_equal(sets[g].keys(), port.sketch_set(codes, [L], mask, w, port.FMH, 1, 20, 181))
    ctx.close()


@pytest.mark.parametrize("seed,L", [(C3_SEED, 3_000_017), ("1110110111011011101101110110111011011101", 400_000)])
def test_sharded_sequence_sketch_equals_whole_sequence_sketch(seed, L):
    """Position-split sketch routed by key range (8- and 16-byte keys): the gathered set, and the concatenation of
    the ranks' disjoint ordered ranges, equal the single-GPU sketch of the whole sequence and the oracle's."""
    world = _world()
    mask, w = sks.seed_to_mask(seed)
    pred = sks.frac_min_hash(1, 50)

    def rank_body(rank, ctx, comm):
        shard = multi_gpu.position_shard(L, w, rank, world)
        sl = multi_gpu.synth_slice(ctx, L, 7, shard, w)
        full, n_full = ctx.sketch_sequence_sharded(comm, sl, mask, w, pred, gather=True)
        part, n_part = ctx.sketch_sequence_sharded(comm, sl, mask, w, pred, gather=False)
        out = (full.keys(), n_full, part.keys(), n_part)
        full.close()
        part.close()
        sl.close()
        return out

    res = _run_ranks(world, rank_body)
    ctx = sks.Context(0)
    (whole,) = ctx.sketch(ctx.synth(L, [7], [0], [0]), mask, w, pred, sks.REPR_SORTED)
    want = whole.keys()
    assert np.array_equal(want, port.sketch_set(port.gen(L, 7), [L], mask, w, port.FMH, 1, 50, 181))
    for full, n_full, part, n_part in res:
        assert n_full == n_part == len(want)
        assert np.array_equal(full, want)
    assert np.array_equal(np.concatenate([r[2] for r in res]), want)
    ctx.close()


def test_both_exchange_routes_in_child_processes():
    """The sharded all-vs-all has two ways to bring the keys to the ranks that own them (csrc/sks_comm.cu): all sketches
    to every rank (default for two ranks) and every key to its owner only (default beyond two).  The switch is read
    once per process: the all-vs-all test again, in child processes, with each route forced."""
    import os
    import subprocess
    import sys
    if os.environ.get("SKS_SHARD_ROUTE"):
        pytest.skip("already inside the child process")
    if _world() < 2:
        pytest.skip("one GPU: a single rank exchanges nothing")
    # (SKS_ROUTE_TIGHT: the one-pass grouping by owner gets regions that are sure to overflow, so that every rank has
    # to come back with the exact two passes)
    for env in ({"SKS_SHARD_ROUTE": "gather"}, {"SKS_SHARD_ROUTE": "owner"}, {"SKS_SHARD_ROUTE": "owner", "SKS_ROUTE_TIGHT": "1"}):
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k",
                            "sharded_all_vs_all"], env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, (env, r.stdout[-2000:] + r.stderr[-2000:])


def test_one_call_entry_points_from_resident_and_host_buffers():
    """sks_all_vs_all_resident / sks_all_vs_all_from_host (what bench.py times) against sketch + sks_all_vs_all: genomes in
    a resident batch, in pageable host memory (copied up) and in pinned host memory (read in place by the sketch
    kernel's bulk copies), single rank."""
    import torch
    ctx = sks.Context(0)
    n, L = 9, 400_003      # not a multiple of 16 bases: the in-place path fetches the last words with plain loads
    mask, w = sks.seed_to_mask(C3_SEED)
    pred = sks.frac_min_hash(1, 20)
    batch = ctx.synth(L, [1000] * n, [2000 + g for g in range(n)], [DS[g % 6] for g in range(n)])
    sets = ctx.sketch(batch, mask, w, pred, sks.REPR_SORTED)
    want = ctx.all_vs_all(sets)
    for s in sets:
        s.close()

    def fresh():
        return np.zeros((n, n), np.int32), np.zeros(n, np.int32), np.zeros((n, n), np.float64)

    got = ctx.all_vs_all_resident(None, batch, n, mask, w, pred, fresh())
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
    words = (L + 15) // 16
    stride = (words + 3) // 4 * 4
    pageable = np.zeros(n * stride, dtype=np.uint32)
    for g in range(n):
        pageable[g * stride:g * stride + words] = batch.download(g)
    pinned = torch.from_numpy(pageable.view(np.int32).copy()).pin_memory()
    before = ctx.in_place_calls
    for base_ptr, in_place in ((pageable.ctypes.data, 0), (pinned.data_ptr(), 1)):
        ptrs = [base_ptr + 4 * g * stride for g in range(n)]
        got = ctx.all_vs_all_from_host(None, ptrs, [L] * n, n, mask, w, pred, fresh())
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), in_place
        assert ctx.in_place_calls - before == in_place
    ctx.close()


def test_pinned_host_genomes_are_streamed_under_the_sketch_kernel():
    """sks_all_vs_all_from_host on pinned host buffers above the streaming threshold (csrc/sks_api.cu, batch_streamed): the
    copy engine brings the genomes chunk by chunk, every chunk is sketched behind its own event.  Equally long, equally
    spaced genomes travel as one strided copy, the others one by one; thresholds lowered so that 30 small genomes make
    several chunks.  Pageable buffers go the same way through a feeder thread and pinned staging buffers.  Same matrix as
    from a resident batch; SKS_HOST_STREAM=0 takes the in-place route / the plain upload again."""
    import os
    import torch
    ctx = sks.Context(0)
    mask, w = sks.seed_to_mask(C3_SEED)
    pred = sks.frac_min_hash(1, 20)
    lens = [400_003] * 12 + [150_000 + 7919 * g for g in range(10)] + [262_144] * 8      # a run, ragged ones, a run
    n = len(lens)
    batches = [ctx.synth(L, [1000 + g % 3], [2000 + g], [DS[g % 6]]) for g, L in enumerate(lens)]
    words = [(L + 15) // 16 for L in lens]
    offs, at = [], 0
    for g in range(n):
        offs.append(at)
        at += (words[g] + 3) // 4 * 4 + (8 if g == 15 else 0)       # 16-byte aligned starts; one irregular gap
    host = torch.zeros(at, dtype=torch.int32).pin_memory()
    hnp = host.numpy().view(np.uint32)
    for g in range(n):
        hnp[offs[g]:offs[g] + words[g]] = batches[g].download(0)
    ptrs = [host.data_ptr() + 4 * o for o in offs]
    # expected: every genome sketched on its own, all-vs-all of the sets
    sets = [ctx.sketch(b, mask, w, pred, sks.REPR_SORTED)[0] for b in batches]
    want = ctx.all_vs_all(sets)
    for g in (0, 13, 29):
        assert np.array_equal(sets[g].keys(), port.sketch_set(port.mutate(port.gen(lens[g], 1000 + g % 3), 2000 + g, DS[g % 6]) if DS[g % 6] else port.gen(lens[g], 1000 + g % 3),
                                                              [lens[g]], mask, w, port.FMH, 1, 20, 181))
    for s_ in sets:
        s_.close()

    def fresh():
        return np.zeros((n, n), np.int32), np.zeros(n, np.int32), np.zeros((n, n), np.float64)

    saved = {k: os.environ.get(k) for k in ("SKS_HOST_STREAM", "SKS_HOST_STREAM_MIN_MB", "SKS_HOST_CHUNK_MB", "SKS_HOST_THREADS")}
    try:
        os.environ["SKS_HOST_STREAM_MIN_MB"] = "0"
        os.environ["SKS_HOST_CHUNK_MB"] = "1"
        os.environ.pop("SKS_HOST_STREAM", None)
        s0, p0 = ctx.streamed_calls, ctx.in_place_calls
        for _ in range(3):
            got = ctx.all_vs_all_from_host(None, ptrs, lens, n, mask, w, pred, fresh())
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
        assert ctx.streamed_calls - s0 == 3 and ctx.in_place_calls == p0
        # pageable memory: a feeder thread stages the chunks through pinned buffers and queues the copies as it goes
        pageable = np.array(hnp, copy=True)
        pptrs = [pageable.ctypes.data + 4 * o for o in offs]
        for threads in ("1", "3", None):
            if threads is None:
                os.environ.pop("SKS_HOST_THREADS", None)
            else:
                os.environ["SKS_HOST_THREADS"] = threads
            got = ctx.all_vs_all_from_host(None, pptrs, lens, n, mask, w, pred, fresh())
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), threads
        assert ctx.streamed_calls - s0 == 6 and ctx.in_place_calls == p0
        os.environ["SKS_HOST_STREAM"] = "0"
        got = ctx.all_vs_all_from_host(None, ptrs, lens, n, mask, w, pred, fresh())
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
        assert ctx.streamed_calls - s0 == 6 and ctx.in_place_calls == p0 + 1
        got = ctx.all_vs_all_from_host(None, pptrs, lens, n, mask, w, pred, fresh())      # pageable, plain upload
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
        assert ctx.streamed_calls - s0 == 6 and ctx.in_place_calls == p0 + 1
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    ctx.close()


def test_configs3_full_size_matrix_digest():
    """BASELINE configs[3] at full size (1000 x 5 Mbp, all 10^6 ordered pairs) on one GPU: the sha256 of the count matrix and
    of the sizes equal the recorded values (bench.py; first produced by the pairwise kernels, which the parity tests pin
    to the oracle), the matrix is symmetric with the sizes on its diagonal, and the ANI follows the substitution rate."""
    import hashlib
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ctx = sks.Context(0)
    n = bench.C4_N
    mask, w = sks.seed_to_mask(bench.C3_SEED)
    batch = ctx.synth(bench.C4_L, [1000] * n, [2000 + g for g in range(n)], [bench.C4_DS[g % 6] for g in range(n)])
    out = (np.zeros((n, n), np.int32), np.zeros(n, np.int32), np.zeros((n, n), np.float64))
    ctx.all_vs_all_resident(None, batch, n, mask, w, sks.frac_min_hash(1, 200), out)
    counts, sizes, ani = out
    assert hashlib.sha256(counts.astype("<i4").tobytes()).hexdigest() == bench.C4_MATRIX_SHA256
    assert hashlib.sha256(sizes.astype("<i4").tobytes()).hexdigest() == bench.C4_SIZES_SHA256
    assert (counts == counts.T).all() and (np.diag(counts) == sizes).all()
    for g in range(1, 60):
        d = bench.C4_DS[g % 6]
        want = 1.0 if d == 0 else 1.0 - 1.0 / d
        assert abs(ani[0, g] - want) < 0.002, (g, ani[0, g])
    ctx.close()
