import os, sys, time, glob, subprocess
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import spaced_kmer_sketching_b200 as sks

print(subprocess.run("nvidia-smi topo -m 2>&1 | head -20; lscpu | grep -i 'numa\\|socket\\|model name' ; nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader", shell=True, capture_output=True, text=True).stdout)
bdf = subprocess.run("nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0", shell=True, capture_output=True, text=True).stdout.strip().lower()
bdf = bdf[4:] if len(bdf) > 12 else bdf
try:
    print("gpu0 numa_node", open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
except Exception as ex:
    print("no numa_node", ex)
nodes = {}
for d in sorted(glob.glob("/sys/devices/system/node/node*")):
    cl = open(d + "/cpulist").read().strip()
    cpus = []
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-"); cpus += list(range(int(a), int(b) + 1))
        elif part:
            cpus.append(int(part))
    nodes[int(d.rsplit("node", 1)[1])] = cpus
print("nodes", {k: (v[0], v[-1], len(v)) for k, v in nodes.items() if v}, "affinity now", len(os.sched_getaffinity(0)))

G = 1000; L = 5_000_000
ctx = sks.Context(0)
mask, w = sks.seed_to_mask("0011111011010111111011001011101")
pred = sks.frac_min_hash(1, 200)
Ds = [[0, 1000, 200, 100, 50, 20][g % 6] for g in range(G)]
batch = ctx.synth(L, [1000] * G, [2000 + g for g in range(G)], Ds)
words = L // 16 + (1 if L % 16 else 0); stride = (words + 3) // 4 * 4
out = (np.zeros((G, G), np.int32), np.zeros(G, np.int32), np.zeros((G, G), np.float64))
allowed = sorted(os.sched_getaffinity(0))
def run(tag):
    host = torch.empty(G * stride, dtype=torch.int32).pin_memory()
    hnp = host.numpy().view(np.uint32)
    for g in range(G):
        hnp[g * stride:g * stride + words] = batch.download(g)
    ptrs = [host.data_ptr() + 4 * g * stride for g in range(G)]
    ts = []
    for it in range(5):
        t0 = time.perf_counter()
        ctx.all_vs_all_from_host(None, ptrs, [L] * G, G, mask, w, pred, out)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(tag, "e2e ms", [round(t, 2) for t in ts], flush=True)
    del host
run("default")
for k, cpus in nodes.items():
    use = [c for c in cpus if c in allowed]
    if not use: 
        print("node", k, "no allowed cpus"); continue
    os.sched_setaffinity(0, use)
    run("node %d (%d cpus)" % (k, len(use)))
os.sched_setaffinity(0, allowed)
