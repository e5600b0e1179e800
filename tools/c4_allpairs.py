"""Dev tool: the C4 all-vs-all (G genomes of 5 Mbp, FMH(200)) on one GPU through sks_all_vs_all (dictionary route) and,
for comparison and as a cross-check, the pairwise kernels; per-kernel times."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
check = len(sys.argv) > 2 and sys.argv[2] == "check"
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("0011111011010111111011001011101")
pred = sks.frac_min_hash(1, 200)
Ds = [[0, 1000, 200, 100, 50, 20][g % 6] for g in range(G)]
batch = ctx.synth(5_000_000, [1000] * G, [2000 + g for g in range(G)], Ds)
for it in range(4):
    ctx.profile(True)
    ctx.kernel_stats()
    t0 = time.perf_counter()
    sets = ctx.sketch(batch, mask, w, pred)
    ctx.sync()
    t1 = time.perf_counter()
    counts, sizes, ani = ctx.all_vs_all(sets)
    t2 = time.perf_counter()
    ks = ctx.kernel_stats()
    print("G=%d sketch %.2f ms all_vs_all %.2f ms (wall) kernels:" % (G, (t1 - t0) * 1e3, (t2 - t1) * 1e3),
          {k: (v[0], round(v[1], 3)) for k, v in ks.items()}, "checksum", int(counts.sum()), "ani[0,1]=%.12f" % ani[0, 1],
          flush=True)
    if check and it == 0:
        t3 = time.perf_counter()
        want = ctx.intersect_block(sets, (0, G), (0, G), np.full((G, G), -1, dtype=np.int32))
        t4 = time.perf_counter()
        print("pairwise kernels %.2f ms, equal: %s" % ((t4 - t3) * 1e3, bool(np.array_equal(want, counts))), flush=True)
        wani = sks.ani_from_counts(counts.ravel(), np.repeat(sizes, G), sks.mask_weight(mask))
        print("max |ani - host ani| = %.3g" % float(np.max(np.abs(ani.ravel() - wani))))
    for s in sets:
        s.close()
