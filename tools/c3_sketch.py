"""Dev tool: the C3 sketch (FMH, weight-21 span-31 seed) on a synthetic sequence, for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks

L = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("0011111011010111111011001011101")
batch = ctx.synth(L, [7], [0], [0])
ctx.profile(True)
for i in range(4):
    (s,) = ctx.sketch(batch, mask, w, sks.frac_min_hash(1, 200))
    n = s.kmer_set_size()
    s.close()
print(L, n, {k: (v[0], round(v[1] / v[0], 4)) for k, v in ctx.kernel_stats().items()})
