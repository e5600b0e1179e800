"""Dev tool: randomised stress of the all-vs-all dictionary route against the pairwise kernels (different numbers of
sets, sharing patterns, masks, moduli, row ranges), for a given number of seconds."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = sks.Context(0)
t_end = time.time() + budget
trial = 0
while time.time() < t_end:
    rng = np.random.default_rng(seed0 * 100003 + trial)
    n = int(rng.integers(4, 80))
    n_bases = int(rng.integers(1, 4))
    bases = [rng.integers(0, 4, int(rng.integers(2000, 120000)), dtype=np.uint8) for _ in range(n_bases)]
    genomes = []
    for g in range(n):
        x = bases[int(rng.integers(0, n_bases))].copy()
        d = int(rng.choice([0, 0, 1000, 100, 30, 10, 4, 2]))
        if d:
            idx = rng.integers(0, len(x), max(len(x) // d, 1))
            x[idx] = (x[idx] + rng.integers(1, 4, len(idx))) & 3
        if rng.random() < 0.1:
            x = x[: int(rng.integers(1, 80))]
        if rng.random() < 0.05:
            x = np.zeros(int(rng.integers(100, 5000)), dtype=np.uint8)
        genomes.append(x)
    batch = ctx.upload_codes(genomes)
    w = int(rng.integers(5, 65))
    k = int(rng.integers(max(3, w // 3), w + 1))
    mask = sks.generate_random_spaced_seed_mask(w, k, int(rng.integers(0, 1000)))
    pred = sks.all_kmers() if rng.random() < 0.3 else sks.frac_min_hash(int(rng.integers(0, 4)), int(rng.integers(2, 60)))
    if os.environ.get("STRESS_VERBOSE"):
        print("trial %d n %d lens %r w %d k %d mask %x pred %r/%r/%r" % (trial, n, [len(g) for g in genomes], w, k, mask, pred.kind,
                                                                     getattr(pred, "nonce", None), getattr(pred, "modulus", None)), flush=True)
    sets = ctx.sketch(batch, mask, w, pred, sks.REPR_SORTED)
    want = ctx.intersect_block(sets, (0, n), (0, n), np.full((n, n), -1, dtype=np.int32))
    r0, r1 = sorted(int(v) for v in rng.integers(0, n + 1, 2))
    for rows in ((0, n), (r0, r1)):
        cnt, sizes, ani = ctx.all_vs_all(sets, rows[0], rows[1])
        if not np.array_equal(cnt, want[rows[0]:rows[1]]):
            bad = np.argwhere(cnt != want[rows[0]:rows[1]])
            print("MISMATCH trial %d n %d w %d k %d rows %r first bad %r got %d want %d" % (
                trial, n, w, k, rows, bad[0].tolist(), cnt[tuple(bad[0])], want[rows[0]:rows[1]][tuple(bad[0])]))
            sys.exit(1)
        assert sizes.tolist() == [s.kmer_set_size() for s in sets]
    for s in sets:
        s.close()
    batch.close()
    trial += 1
print("stress ok: %d trials" % trial)
