"""Dev tool: the bench's C4 step (sks_all_vs_all_resident) on one GPU, G genomes; per-kernel times; for ncu captures."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks

G = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("0011111011010111111011001011101")
pred = sks.frac_min_hash(1, 200)
Ds = [[0, 1000, 200, 100, 50, 20][g % 6] for g in range(G)]
batch = ctx.synth(5_000_000, [1000] * G, [2000 + g for g in range(G)], Ds)
out = (np.zeros((G, G), np.int32), np.zeros(G, np.int32), np.zeros((G, G), np.float64))
for it in range(reps):
    ctx.profile(True)
    ctx.kernel_stats()
    t0 = time.perf_counter()
    ctx.all_vs_all_resident(None, batch, G, mask, w, pred, out)
    t1 = time.perf_counter()
    ks = ctx.kernel_stats()
    print("G=%d step %.2f ms (wall) kernels:" % (G, (t1 - t0) * 1e3), {k: (v[0], round(v[1], 3)) for k, v in ks.items()},
          "checksum", int(out[0].sum()), flush=True)
