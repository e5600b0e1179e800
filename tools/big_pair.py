"""Dev tool: a 50 Mbp pair through the bitset pair pipeline (bucket loads far above the staging capacity: the
kernels' direct paths), checked against the sorted route."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks
ctx = sks.Context(0)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
mask, w = sks.seed_to_mask("011101110010111110011011")
batch = ctx.synth(L, [42, 42], [0, 43], [0, 100])
ctx.profile(True)
out = {}
for name, r in (("bitset", sks.REPR_BITSET), ("onchip", sks.REPR_BITSET_ONCHIP), ("sorted", sks.REPR_SORTED)):
    t0 = time.perf_counter(); res = ctx.pair_ani_resident(batch, mask, w, sks.all_kmers(), r); t1 = time.perf_counter()
    out[name] = (res.size_a, res.size_b, res.intersection)
    print(name, out[name], "%.2f ms" % ((t1 - t0) * 1e3), {k: round(v[1] / v[0], 3) for k, v in ctx.kernel_stats().items()})
assert out["bitset"] == out["sorted"] == out["onchip"], out
print("ok")
