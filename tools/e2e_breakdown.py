"""Dev tool: where the end-to-end pair call spends its time (pinned host buffers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spaced_kmer_sketching_b200 as sks
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("011101110010111110011011")
L = 5_000_000
batch = ctx.synth(L, [42, 42], [0, 43], [0, 100])
wa = torch.from_numpy(batch.download(0).view(np.int32)).pin_memory()
wb = torch.from_numpy(batch.download(1).view(np.int32)).pin_memory()
na, nb = wa.numpy().view(np.uint32), wb.numpy().view(np.uint32)
def t(fn, n=30):
    for _ in range(5): fn()
    ctx.sync(); t0 = time.perf_counter()
    for _ in range(n): fn()
    ctx.sync(); return (time.perf_counter() - t0) / n * 1e3
def up():
    b = ctx.upload([(na, L, None), (nb, L, None)]); ctx.sync(); b.close()
print("upload + sync            %.3f ms" % t(up))
print("resident pair            %.3f ms" % t(lambda: ctx.pair_ani_resident(batch, mask, w, sks.all_kmers(), sks.REPR_BITSET)))
print("e2e pair (pinned)        %.3f ms" % t(lambda: ctx.pair_ani_ptr(wa.data_ptr(), L, wb.data_ptr(), L, mask, w, sks.all_kmers(), sks.REPR_BITSET)))
d = torch.empty(2 * wa.numel(), dtype=torch.int32, device="cuda")
def copy_only():
    d[:wa.numel()].copy_(wa, non_blocking=True); d[wa.numel():].copy_(wb, non_blocking=True); torch.cuda.synchronize()
print("2 x 1.25 MB H2D (torch)  %.3f ms" % t(copy_only))
