"""Dev tool (torchrun, N >= 2): where the all-gather of sketches and the exchange of count blocks spend their time."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spaced_kmer_sketching_b200 as sks
from spaced_kmer_sketching_b200 import multi_gpu

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
ctx = sks.Context(torch.cuda.current_device())
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
mask, w = sks.seed_to_mask("0011111011010111111011001011101")
pred = sks.frac_min_hash(1, 200)
G = 125
ids = list(range(rank * G, (rank + 1) * G))
b = ctx.synth(5_000_000, [1000] * G, [2000 + g for g in ids], [[0, 1000, 200, 100, 50, 20][g % 6] for g in ids])
local_sets = ctx.sketch(b, mask, w, pred)


def timed(label, fn, acc):
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    acc.append((label, (time.perf_counter() - t) * 1e3))
    return r


buf = None
for it in range(4):
    acc = []
    dist.barrier()
    keys, counts, kw = timed("keys_as_tensor", lambda: multi_gpu.keys_as_tensor(local_sets, torch), acc)
    cnt = timed("counts tensor", lambda: torch.tensor(counts, dtype=torch.int64, device="cuda"), acc)
    all_counts, all_keys = timed("allgather_varlen_many", lambda: multi_gpu.allgather_varlen_many([cnt, keys], world, dist), acc)
    counts_host = timed("counts to host", lambda: torch.cat(all_counts).cpu(), acc)

    def make_sets():
        out, at = [], 0
        for r in range(world):
            n_r = all_counts[r].numel()
            out.extend(ctx.sets_from_device_keys(all_keys[r].data_ptr(), counts_host[at:at + n_r].tolist(), kw, mask, w))
            at += n_r
        return out
    all_sets = timed("sets_from_device_keys", make_sets, acc)
    cm = buf = timed("tiled_counts", lambda: multi_gpu.tiled_counts(ctx, all_sets, rank, world, buf), acc)
    mine = timed("exchange_blocks", lambda: multi_gpu.exchange_blocks(cm, rank, world), acc)
    rows = multi_gpu.row_tile(len(all_sets), rank, world)
    fs = timed("first_sizes", lambda: np.repeat(np.diagonal(mine[:, rows[0]:rows[1]]), len(all_sets)).astype(np.int32), acc)
    ani = timed("ani_from_counts", lambda: sks.ani_from_counts(np.ascontiguousarray(mine).ravel(), fs, sks.mask_weight(mask)), acc)
    if rank == 0:
        print("it %d: " % it + ", ".join("%s %.2f" % x for x in acc))
    for s in all_sets:
        s.close()
dist.destroy_process_group()
