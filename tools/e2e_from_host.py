"""Dev tool: sks_all_vs_all_from_host on 1000 x 5 Mbp pinned host genomes, in place against streamed, by chunk size."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import spaced_kmer_sketching_b200 as sks
G = 1000; L = 5_000_000
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("0011111011010111111011001011101")
pred = sks.frac_min_hash(1, 200)
Ds = [[0, 1000, 200, 100, 50, 20][g % 6] for g in range(G)]
batch = ctx.synth(L, [1000] * G, [2000 + g for g in range(G)], Ds)
words = L // 16 + (1 if L % 16 else 0); stride = (words + 3) // 4 * 4
out = (np.zeros((G, G), np.int32), np.zeros(G, np.int32), np.zeros((G, G), np.float64))
host = torch.empty(G * stride, dtype=torch.int32).pin_memory()
hnp = host.numpy().view(np.uint32)
for g in range(G):
    hnp[g * stride:g * stride + words] = batch.download(g)
ptrs = [host.data_ptr() + 4 * g * stride for g in range(G)]
ref = None
def run(tag):
    global ref
    ts = []
    for it in range(5):
        t0 = time.perf_counter()
        ctx.all_vs_all_from_host(None, ptrs, [L] * G, G, mask, w, pred, out)
        ts.append((time.perf_counter() - t0) * 1e3)
    if ref is None: ref = out[0].copy()
    assert np.array_equal(ref, out[0])
    print(tag, "e2e ms", [round(t, 2) for t in ts], "streamed", ctx.streamed_calls, "in place", ctx.in_place_calls, flush=True)
os.environ["SKS_HOST_STREAM"] = "0"; run("in place")
os.environ["SKS_HOST_STREAM"] = "1"
for mb in (8, 16, 32):
    os.environ["SKS_HOST_CHUNK_MB"] = str(mb); run("chunk %d MB" % mb)
pageable = np.array(hnp, copy=True)
pinned_ptrs = ptrs
ptrs = [pageable.ctypes.data + 4 * g * stride for g in range(G)]
os.environ["SKS_HOST_CHUNK_MB"] = "16"
os.environ["SKS_HOST_STREAM"] = "0"; run("pageable, plain upload")
os.environ["SKS_HOST_STREAM"] = "1"
for th in (1, 2, 4, 8, 16):
    os.environ["SKS_HOST_THREADS"] = str(th); run("pageable, staged, %d threads" % th)
os.environ.pop("SKS_HOST_THREADS")
ptrs = pinned_ptrs
ctx.profile(True); ctx.kernel_stats()
os.environ["SKS_HOST_CHUNK_MB"] = "16"
ctx.all_vs_all_from_host(None, ptrs, [L] * G, G, mask, w, pred, out)
print({k: (v[0], round(v[1], 3)) for k, v in ctx.kernel_stats().items()})
