"""Dev tool: per-kernel times of the C2 bitset sketch under SKS_BUILD_DEBUG ablations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("011101110010111110011011")
batch = ctx.synth(5_000_000, [42, 42], [0, 43], [0, 100])
for i in range(3):
    s = ctx.sketch(batch, mask, w, sks.all_kmers(), sks.REPR_BITSET); [x.close() for x in s]
ctx.profile(True); ctx.kernel_stats()
for i in range(10):
    s = ctx.sketch(batch, mask, w, sks.all_kmers(), sks.REPR_BITSET); [x.close() for x in s]
print(os.environ.get("SKS_BUILD_DEBUG"), {k: round(v[1] / v[0] * 1e3, 1) for k, v in ctx.kernel_stats().items()})
