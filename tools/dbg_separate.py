"""Dev tool: host-side timing of the general API route on C2 (sketch -> two bitset sets -> intersect)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spaced_kmer_sketching_b200 as sks
stream = torch.cuda.Stream()
ctx = sks.Context(0)
ctx.set_stream(stream.cuda_stream)
mask, w = sks.seed_to_mask("011101110010111110011011")
batch = ctx.synth(5_000_000, [42, 42], [0, 43], [0, 100])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(tag, prof, use_flush, events):
    ctx.profile(prof)
    for i in range(6):
        if events:
            e0 = torch.cuda.Event(enable_timing=True); e0.record(stream)
        t0 = time.perf_counter(); sa, sb = ctx.sketch(batch, mask, w, sks.all_kmers(), sks.REPR_BITSET)
        t1 = time.perf_counter(); n = ctx.intersect(sa, sb)
        t2 = time.perf_counter(); sa.close(); sb.close()
        t3 = time.perf_counter()
        if use_flush:
            flush.zero_()
        torch.cuda.synchronize()
        print("%s sketch %.3f ms  intersect %.3f ms  close %.3f ms" % (tag, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    ctx.kernel_stats(); ctx.profile(False)
with torch.cuda.stream(stream):
    for _ in range(3):
        ctx.pair_ani_resident(batch, mask, w, sks.all_kmers(), sks.REPR_BITSET)
    run("plain     ", False, False, False)
    run("profile   ", True, False, False)
    run("flush     ", False, True, False)
    run("prof+flush", True, True, True)
