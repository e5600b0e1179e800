"""Dev tool: BASELINE configs[0] (C1): 5 Mbp pair, seed 11001011 (weight 5), predicates ALL and FMH(200)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("11001011")
batch = ctx.synth(5_000_000, [42, 42], [0, 43], [0, 100])
ctx.profile(True)
for pred in (sks.all_kmers(), sks.frac_min_hash(1, 200)):
    for repr_ in (sks.REPR_AUTO, sks.REPR_SORTED):
        for i in range(4):
            t0 = time.perf_counter(); r = ctx.pair_ani_resident(batch, mask, w, pred, repr_); t1 = time.perf_counter()
        print(pred.kind, repr_, (r.size_a, r.size_b, r.intersection, r.ani_ab), "wall %.3f ms" % ((t1 - t0) * 1e3),
              {k: (v[0], round(v[1] / v[0], 4)) for k, v in ctx.kernel_stats().items()})
