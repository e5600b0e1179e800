"""Dev tool: predicate ALL -> bitset for contiguous seeds of weight 4..16, and FMH(200) -> sorted for weights 8..32
(5 Mbp pair): looks for pathologies between the headline configurations."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks
ctx = sks.Context(0)
batch = ctx.synth(5_000_000, [42, 42], [0, 43], [0, 100])
ctx.profile(True)
for k in range(4, 17):
    mask = sks.contiguous_kmer(k)
    for i in range(3):
        t0 = time.perf_counter(); r = ctx.pair_ani_resident(batch, mask, k, sks.all_kmers(), sks.REPR_BITSET); t1 = time.perf_counter()
    print("ALL  k=%2d" % k, (r.size_a, r.intersection), "wall %.3f ms" % ((t1 - t0) * 1e3), {n: round(v[1] / v[0], 4) for n, v in ctx.kernel_stats().items()})
for k in (8, 12, 16, 20, 24, 28, 32):
    mask = sks.contiguous_kmer(k)
    for i in range(3):
        t0 = time.perf_counter(); r = ctx.pair_ani_resident(batch, mask, k, sks.frac_min_hash(1, 200), sks.REPR_SORTED); t1 = time.perf_counter()
    print("FMH  k=%2d" % k, (r.size_a, r.intersection), "wall %.3f ms" % ((t1 - t0) * 1e3), {n: round(v[1] / v[0], 4) for n, v in ctx.kernel_stats().items()})
for k in (17, 20, 24):
    mask = sks.contiguous_kmer(k)
    for i in range(3):
        t0 = time.perf_counter(); r = ctx.pair_ani_resident(batch, mask, k, sks.all_kmers(), sks.REPR_SORTED); t1 = time.perf_counter()
    print("ALLs k=%2d" % k, (r.size_a, r.intersection), "wall %.3f ms" % ((t1 - t0) * 1e3), {n: round(v[1] / v[0], 4) for n, v in ctx.kernel_stats().items()})
