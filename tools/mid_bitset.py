"""Dev tool: predicate ALL -> bitset for weights 13..16 under different SKS_BUCKET_MIN_BITS (set in the environment)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks
ctx = sks.Context(0)
batch = ctx.synth(5_000_000, [42, 42], [0, 43], [0, 100])
ctx.profile(True)
for k in range(10, 17):
    mask = sks.contiguous_kmer(k)
    for i in range(3):
        t0 = time.perf_counter(); r = ctx.pair_ani_resident(batch, mask, k, sks.all_kmers(), sks.REPR_BITSET); t1 = time.perf_counter()
    print(os.environ.get("SKS_BUCKET_MIN_BITS"), "k=%2d" % k, (r.size_a, r.intersection), "wall %.3f ms" % ((t1 - t0) * 1e3), {n: round(v[1] / v[0], 4) for n, v in ctx.kernel_stats().items()})
