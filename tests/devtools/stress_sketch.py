"""Dev tool (test infrastructure: it uses the CPU oracle): randomised stress of the sketch path -- random segmented
genomes, masks of every span, ALL / FracMinHash conditions, both set representations -- against oracle/port.py, for a
given number of seconds.  `python tests/devtools/stress_sketch.py [seconds] [seed]`"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import spaced_kmer_sketching_b200 as sks
from oracle import port

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = sks.Context(0)
t_end = time.time() + budget
trial = 0
while time.time() < t_end:
    rng = np.random.default_rng(seed0 * 7919 + trial)
    n = int(rng.integers(1, 12))
    genomes, segs = [], []
    for g in range(n):
        L = int(rng.choice([0, 5, 70, 1000, 8191, 8192, 8193, 20000, 70000]))
        L = max(0, L + int(rng.integers(-3, 4)))
        x = rng.integers(0, 4, L, dtype=np.uint8)
        kind = rng.random()
        if kind < 0.1:
            x[:] = int(rng.integers(0, 4))                     # homopolymer
        elif kind < 0.2 and L > 8:
            unit = rng.integers(0, 4, int(rng.integers(1, 7)), dtype=np.uint8)
            x = np.resize(unit, L)                             # short tandem repeat: heavy duplication
        cuts = sorted(set(int(c) for c in rng.integers(0, L + 1, int(rng.integers(0, 6))))) if L else []
        bounds = [0] + cuts + [L]
        seg = [b - a for a, b in zip(bounds[:-1], bounds[1:]) if b > a] or ([L] if L else [])
        genomes.append(x)
        segs.append(seg)
    w = int(rng.integers(1, 65))
    k = int(rng.integers(1, w + 1))
    mask = sks.generate_random_spaced_seed_mask(w, k, int(rng.integers(0, 1000)))
    weight = sks.mask_weight(mask)
    if rng.random() < 0.35:
        pred, opred = sks.all_kmers(), (port.ALL,)
    else:
        nonce, mod = int(rng.integers(-3, 5)), int(rng.choice([1, 2, 3, 7, 8, 25, 200, 1024, 3000]))
        var = int(rng.choice([171, 181]))
        pred, opred = sks.frac_min_hash(nonce, mod, var), (port.FMH, nonce, mod, var)
    reprs = [sks.REPR_SORTED] + ([sks.REPR_BITSET] if pred.kind == sks.PRED_ALL and weight <= 12 else [])
    if os.environ.get("STRESS_VERBOSE"):
        print("trial %d lens %r w %d k %d mask %x pred %r" % (trial, [len(g) for g in genomes], w, k, mask, opred), flush=True)
    batch = ctx.upload_codes(genomes, segs)
    want = [port.sketch_set(g, s, mask, w, *opred) for g, s in zip(genomes, segs)]
    for r in reprs:
        sets = ctx.sketch(batch, mask, w, pred, r)
        for i, (s, o) in enumerate(zip(sets, want)):
            if not np.array_equal(s.keys(), o):
                print("MISMATCH trial %d genome %d repr %d w %d k %d mask %x pred %r segs %r" % (trial, i, r, w, k, mask, opred, segs[i]))
                sys.exit(1)
        if len(sets) >= 2:
            a, b = int(rng.integers(0, len(sets))), int(rng.integers(0, len(sets)))
            assert ctx.intersect(sets[a], sets[b]) == port.intersection(want[a], want[b]), (trial, a, b, r)
        for s in sets:
            s.close()
    batch.close()
    trial += 1
print("stress ok: %d trials" % trial)
