"""Dev tool: intersection of two LARGE sorted sets (5 Mbp, predicate ALL, weight-21 seed: 5 M keys each)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("0011111011010111111011001011101")
batch = ctx.synth(5_000_000, [42, 42], [0, 43], [0, 100])
sa, sb = ctx.sketch(batch, mask, w, sks.all_kmers(), sks.REPR_SORTED)
ctx.profile(True)
for i in range(4):
    t0 = time.perf_counter(); n = ctx.intersect(sa, sb); t1 = time.perf_counter()
print(sa.kmer_set_size(), sb.kmer_set_size(), n, "wall %.3f ms" % ((t1 - t0) * 1e3), {k: (v[0], round(v[1] / v[0], 4)) for k, v in ctx.kernel_stats().items()})
