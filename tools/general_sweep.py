"""Dev tool: the general route (sks_sketch -> sets -> all pairs) on 8 x 5 Mbp genomes for several weights and both
representations: looks for pathologies outside the benchmarked configurations."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks
ctx = sks.Context(0)
n = 8
batch = ctx.synth(5_000_000, [1000] * n, [2000 + g for g in range(n)], [0, 1000, 200, 100, 50, 20, 1000, 200])
ctx.profile(True)
def run(tag, mask, w, pred, repr_):
    for i in range(3):
        t0 = time.perf_counter()
        sets = ctx.sketch(batch, mask, w, pred, repr_)
        t1 = time.perf_counter()
        cnt = ctx.intersect_all_pairs(sets)
        t2 = time.perf_counter()
        sizes = [s.kmer_set_size() for s in sets]
        for s in sets: s.close()
    print(tag, "sketch %.3f ms  all-pairs %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), sizes[0], int(cnt[0, 1]),
          {k: (v[0], round(v[1] / v[0], 4)) for k, v in ctx.kernel_stats().items()})
for k in (8, 12, 14, 15, 16):
    run("ALL bitset k=%d" % k, sks.contiguous_kmer(k), k, sks.all_kmers(), sks.REPR_BITSET)
for k in (16, 21, 32):
    run("ALL sorted k=%d" % k, sks.contiguous_kmer(k), k, sks.all_kmers(), sks.REPR_SORTED)
for w, k in ((31, 21), (40, 30), (50, 40), (64, 33)):
    run("FMH sorted (%d,%d)" % (w, k), sks.generate_random_spaced_seed_mask(w, k), w, sks.frac_min_hash(1, 200), sks.REPR_SORTED)
