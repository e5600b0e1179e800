"""Dev tool: the C2 pair pipeline (fused build + AND/popcount), stored and on-chip, for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks

L = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("011101110010111110011011")
batch = ctx.synth(L, [42, 42], [0, 43], [0, 100])
ctx.profile(True)
for repr_ in (sks.REPR_BITSET, sks.REPR_BITSET_ONCHIP):  # 2 = stored, 3 = on chip
    for i in range(reps):
        r = ctx.pair_ani_resident(batch, mask, w, sks.all_kmers(), repr_)
    print(repr_, (r.size_a, r.size_b, r.intersection), {k: (v[0], round(v[1] / v[0], 4)) for k, v in ctx.kernel_stats().items()})
