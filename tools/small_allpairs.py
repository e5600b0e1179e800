"""Dev tool: a small all-vs-all through the dictionary route (one share; the key space in three shares with
SKS_DICT_PARTS=3), checked against the pairwise kernels -- small enough for a run under a memory checker where one is
available (compute-sanitizer is closed on the shared GPU pool)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks

ctx = sks.Context(0)
rng = np.random.default_rng(3)
base = rng.integers(0, 4, 30_000, dtype=np.uint8)
genomes = []
for g in range(24):
    x = base.copy()
    d = [0, 500, 50, 12, 5, 3][g % 6]
    if d:
        idx = rng.integers(0, len(x), len(x) // d)
        x[idx] = (x[idx] + rng.integers(1, 4, len(idx))) & 3
    genomes.append(x)
genomes[5] = np.zeros(2000, dtype=np.uint8)
batch = ctx.upload_codes(genomes)
n = len(genomes)
for seed, pred in (("0011111011010111111011001011101", sks.all_kmers()), ("0011111011010111111011001011101", sks.frac_min_hash(1, 5)),
                   ("1110110111011011101101110110111011011101", sks.frac_min_hash(1, 3))):
    mask, w = sks.seed_to_mask(seed)
    sets = ctx.sketch(batch, mask, w, pred, sks.REPR_SORTED)
    want = ctx.intersect_block(sets, (0, n), (0, n), np.full((n, n), -1, dtype=np.int32))
    for rows in ((0, n), (3, 17)):
        cnt, sizes, ani = ctx.all_vs_all(sets, rows[0], rows[1])
        assert np.array_equal(cnt, want[rows[0]:rows[1]]), (seed, rows)
    for s in sets:
        s.close()
print("sanitize_allpairs ok")
