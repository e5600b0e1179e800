"""Dev tool: host-to-device copy bandwidth of the box by chunk size (pinned memory)."""
import torch, time
n = 1_250_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk in (n, n // 16, n // 64, n // 1000):
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        off = 0
        while off < n:
            c = min(chunk, n - off)
            d[off:off + c].copy_(h[off:off + c], non_blocking=True)
            off += c
        e1.record()
        torch.cuda.synchronize()
    print("chunk %d MB: %.2f ms = %.1f GB/s" % (chunk // 1_000_000, e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e6), flush=True)
# D2H
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); h.copy_(d, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("D2H %.1f GB/s" % (n / e0.elapsed_time(e1) / 1e6))
