"""Dev tool: randomised stress of the sharded all-vs-all (one thread per visible GPU, sks_comm_init_all) against the
single-GPU result: random numbers of genomes (also fewer than ranks), sizes, sharing patterns, masks incl. 16-byte keys.
`python tests/devtools/stress_sharded.py [seconds] [seed]`; SKS_SHARD_ROUTE / SKS_ROUTE_TIGHT select the exchange."""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import spaced_kmer_sketching_b200 as sks
from spaced_kmer_sketching_b200 import _lib

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
world = max(1, min(int(_lib.load().sks_device_count()), 8))
ctxs = [sks.Context(r) for r in range(world)]
comms = sks.Comm.init_all(ctxs) if world > 1 else [None]
t_end = time.time() + budget
trial = 0
while time.time() < t_end:
    rng = np.random.default_rng(seed0 * 104729 + trial)
    n = int(rng.choice([1, 2, 3, world - 1 if world > 1 else 1, world, world + 1, 17, 40]))
    n = max(n, 1)
    base = rng.integers(0, 4, int(rng.integers(3000, 150000)), dtype=np.uint8)
    genomes = []
    for g in range(n):
        x = base.copy() if rng.random() < 0.8 else rng.integers(0, 4, int(rng.integers(50, 90000)), dtype=np.uint8)
        d = int(rng.choice([0, 1000, 100, 20, 6, 3]))
        if d:
            idx = rng.integers(0, len(x), max(len(x) // d, 1))
            x[idx] = (x[idx] + rng.integers(1, 4, len(idx))) & 3
        if rng.random() < 0.08:
            x = x[: int(rng.integers(1, 70))]
        genomes.append(x)
    w = int(rng.integers(6, 65))
    k = int(rng.integers(max(3, w // 3), w + 1))
    mask = sks.generate_random_spaced_seed_mask(w, k, int(rng.integers(0, 1000)))
    pred = sks.all_kmers() if rng.random() < 0.25 else sks.frac_min_hash(int(rng.integers(0, 4)), int(rng.integers(2, 50)))
    out, err = [None] * world, []

    def body(r):
        try:
            b, e = sks.shard_range(n, r, world)
            sets = []
            if e > b:
                batch = ctxs[r].upload_codes(genomes[b:e])
                sets = ctxs[r].sketch(batch, mask, w, pred, sks.REPR_SORTED)
            out[r] = ctxs[r].all_vs_all_sharded(comms[r], sets, n)
            for s in sets:
                s.close()
        except Exception as ex:   # noqa: BLE001
            err.append((r, repr(ex)))

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if err:
        print("ERROR trial %d n %d w %d k %d: %r" % (trial, n, w, k, err))
        sys.exit(1)
    batch = ctxs[0].upload_codes(genomes)
    sets = ctxs[0].sketch(batch, mask, w, pred, sks.REPR_SORTED)
    want = ctxs[0].intersect_block(sets, (0, n), (0, n), np.full((n, n), -1, dtype=np.int32))
    sizes = np.array([s.kmer_set_size() for s in sets], dtype=np.int32)
    counts = np.concatenate([o[0] for o in out])
    if not (np.array_equal(counts, want) and all(np.array_equal(o[1], sizes) for o in out)):
        print("MISMATCH trial %d n %d w %d k %d mask %x" % (trial, n, w, k, mask))
        sys.exit(1)
    wani = sks.ani_from_counts(want.ravel(), np.repeat(sizes, n), sks.mask_weight(mask)).reshape(n, n)
    assert np.max(np.abs(np.concatenate([o[2] for o in out]) - wani)) <= 1e-12
    for s in sets:
        s.close()
    trial += 1
print("stress ok: %d trials on %d GPUs" % (trial, world))
