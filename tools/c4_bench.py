"""Dev tool: per-kernel times of a C4-style all-vs-all (G genomes of 5 Mbp, FMH(200)) on one GPU."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks

G = int(sys.argv[1]) if len(sys.argv) > 1 else 125
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("0011111011010111111011001011101")
pred = sks.frac_min_hash(1, 200)
Ds = [[0, 1000, 200, 100, 50, 20][g % 6] for g in range(G)]
batch = ctx.synth(5_000_000, [1000] * G, [2000 + g for g in range(G)], Ds)
for it in range(3):
    ctx.profile(True)
    ctx.kernel_stats()
    t0 = time.perf_counter()
    sets = ctx.sketch(batch, mask, w, pred)
    t1 = time.perf_counter()
    counts = ctx.intersect_all_pairs(sets)
    t2 = time.perf_counter()
    ks = ctx.kernel_stats()
    print("G=%d sketch %.2f ms intersect %.2f ms (wall) kernels:" % (G, (t1 - t0) * 1e3, (t2 - t1) * 1e3),
          {k: round(v[1], 3) for k, v in ks.items()}, "checksum", int(counts.sum()), "diag ok",
          bool((np.diag(counts) == [s.kmer_set_size() for s in sets]).all()), "symmetric", bool((counts == counts.T).all()))
    for s in sets:
        s.close()
