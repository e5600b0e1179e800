"""Dev tool: the work of ONE rank of the 8-GPU all-vs-all over 1000 genomes (configs[3]) on a single GPU -- all 1000
sketches are made here (8 batches of 125, as an all-gather would deliver them), then rank `r` evaluates its rectangles
of the pair matrix and the ANI of its rows.  Host-side wall time against kernel time."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks
from spaced_kmer_sketching_b200 import multi_gpu

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
G = int(sys.argv[2]) if len(sys.argv) > 2 else 125
rank = 0
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("0011111011010111111011001011101")
pred = sks.frac_min_hash(1, 200)
sets = []
for r in range(world):
    ids = list(range(r * G, (r + 1) * G))
    Ds = [[0, 1000, 200, 100, 50, 20][g % 6] for g in ids]
    b = ctx.synth(5_000_000, [1000] * G, [2000 + g for g in ids], Ds)
    sets += ctx.sketch(b, mask, w, pred)
    b.close()
n = len(sets)
rows = multi_gpu.row_tile(n, rank, world)
for it in range(4):
    ctx.profile(True)
    ctx.kernel_stats()
    t0 = time.perf_counter()
    counts = buf = np.empty((n, n), dtype=np.int32) if it == 0 else buf
    counts.fill(-1)
    ta = time.perf_counter()
    rects = multi_gpu.block_rects(n, rank, world)
    import ctypes as C
    ps = (C.c_void_p * n)(*[s.h for s in sets])
    flat = np.array([[r[0][0], r[0][1], r[1][0], r[1][1]] for r in rects], dtype=np.int64).reshape(-1)
    tb = time.perf_counter()
    ctx._L.sks_intersect_rects(ctx.h, ps, n, flat.ctypes.data, len(rects), counts.ctypes.data)
    t1 = time.perf_counter()
    print("   fill %.2f ms, handles+rects %.2f ms, sks_intersect_rects %.2f ms" % ((ta - t0) * 1e3, (tb - ta) * 1e3, (t1 - tb) * 1e3))
    ks = ctx.kernel_stats()
    mine = multi_gpu.mirror_rows(counts, rows)      # stand-in for exchange_blocks (no peers here)
    first_sizes = np.repeat(np.array([s.kmer_set_size() for s in sets[rows[0]:rows[1]]], dtype=np.int32), n)
    t2 = time.perf_counter()
    ani = sks.ani_from_counts(np.ascontiguousarray(mine).ravel(), first_sizes, sks.mask_weight(mask))
    t3 = time.perf_counter()
    n_eval = int((counts >= 0).sum())
    print("world=%d n=%d: tiled_counts %.2f ms wall (kernels %s), rows+sizes %.2f ms, ANI of %d values %.2f ms; %d entries evaluated"
          % (world, n, (t1 - t0) * 1e3, {k: round(v[1], 3) for k, v in ks.items()}, (t2 - t1) * 1e3, ani.size, (t3 - t2) * 1e3, n_eval))
