#!/usr/bin/env python
"""bench.py -- the hot path (sketch -> set -> intersect -> ANI) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" is one pass of the hot path over one batch of synthetic input.  The workload is BASELINE.json
configs[1] (C2): a synthetic 5 Mbp genome and its 1%-mutated copy, the reference's own (24,16,seed 0)
spaced seed, predicate ALL, a full 4^16-bit presence bitset per genome (512 MiB), AND/popcount and the
containment^(1/weight) ANI.  For N > 1 every rank runs its own C2 pair (weak scaling, genomes are
independent: no data-path collective); the extra legs (C3 sketching, C4-style all-vs-all with an NCCL
all-gather of sketches) are reported under "extra".

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libref.so: the unmodified
reference sources compiled against the Boost/Cilk shim; the oracle port if that is absent) on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C2_SEED = "011101110010111110011011"          # generate_random_spaced_seed_mask(24, 16, 0)
C3_SEED = "0011111011010111111011001011101"   # generate_random_spaced_seed_mask(31, 21, 0)
C2_L = 5_000_000
KAT4_C2 = (4994572, 4994591, 4244791)         # |A|, |B|, |A n B| from the reference
WORKLOAD = ("C2: synthetic 5 Mbp genome vs 1%-mutated copy per GPU, weight-16 span-24 seed " + C2_SEED +
            ", predicate ALL, one kmer_set per genome (4^16-bit presence bitset, 512 MiB, on the GPU), "
            "intersection (AND/popcount), containment^(1/16) ANI")
METRIC = "spaced_kmers_per_s_sketch_plus_ani"
UNIT = "kmers/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU path on the box's host cores
# ------------------------------------------------------------------------------------------------
def cpu_c2_step(L, workdir):
    """One C2 pass of the reference on a pair of L-base genomes.  Returns (seconds, kind, cores, counts)."""
    from oracle import port, ref
    A = port.gen(L, 42)
    B = port.mutate(A, 43, 100)
    mask, w = port.seed_to_mask(C2_SEED)
    if ref.available():
        fa, fb = os.path.join(workdir, "a.fna"), os.path.join(workdir, "b.fna")
        if not os.path.exists(fa):
            port.write_fasta(fa, A, "a")
            port.write_fasta(fb, B, "b")
        t0 = time.perf_counter()
        # parallel_kmer_sets_from_fasta_files (cilk_for over files -> 2 threads) + kmer_set_intersection
        sa, sb = ref.sets_from_fasta_files([fa, fb], mask, w, ref.ALL, 1, 200, parallel=True)
        inter = ref.intersection(sa, sb)
        ani = ref.binomial_estimator(ref.containment(inter, sa.size()), port.mask_weight(mask))
        dt = time.perf_counter() - t0
        return dt, "reference", min(2, os.cpu_count() or 1), (sa.size(), sb.size(), inter, ani)
    t0 = time.perf_counter()
    sa = port.sketch_set(A, [L], mask, w)
    sb = port.sketch_set(B, [L], mask, w)
    inter = port.intersection(sa, sb)
    ani = port.ani(inter, len(sa), port.mask_weight(mask))
    dt = time.perf_counter() - t0
    return dt, "port", 1, (len(sa), len(sb), inter, ani)


def cpu_fmh_rates():
    """Per-unit rates of the reference's CPU path on the C3 / C4 kind of work (FracMinHash sketches, sketch-sized
    intersections), on a small sample: `parallel_kmer_sets_from_fasta_files` over 8 x 1 Mbp genomes (one thread
    per file, as its cilk_for does) and `parallel_compute_pairwise_kmer_set_intersections` over all 64 ordered
    pairs.  Context for the extra legs, not a target."""
    from oracle import port, ref
    if not ref.available():
        return {"unavailable": "oracle/_ref/libref.so was not built"}
    L, n = 1_000_000, 8
    mask, w = port.seed_to_mask(C3_SEED)
    base = port.gen(L, 1000)
    with tempfile.TemporaryDirectory() as d:
        paths = []
        for g in range(n):
            codes = base if g == 0 else port.mutate(base, 2000 + g, [1000, 200, 100, 50, 20][g % 5])
            paths.append(os.path.join(d, "g%d.fna" % g))
            port.write_fasta(paths[-1], codes, "g%d" % g)
        t0 = time.perf_counter()
        sets = ref.sets_from_fasta_files(paths, mask, w, ref.FMH, 1, 200, parallel=True)
        t1 = time.perf_counter()
        first = [a for a in sets for _ in sets]
        second = [b for _ in sets for b in sets]
        reps = 20
        for _ in range(reps):
            ref.pairwise_intersections(first, second, parallel=True)
        t2 = time.perf_counter()
    threads = min(n, os.cpu_count() or 1)
    return {"sample": "%d x %d-base genomes, seed %s, FMH(200): sketch through parallel_kmer_sets_from_fasta_files, "
                      "all %d ordered pairs x %d through parallel_compute_pairwise_kmer_set_intersections "
                      "(sketches of ~%d k-mers)" % (n, L, C3_SEED, n * n, reps, sets[0].size()),
            "threads": threads, "sketch_bases_per_s": n * L / (t1 - t0),
            "intersect_pairs_per_s": n * n * reps / (t2 - t1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.steps + args.warmup
    # the full C2 pair costs the reference ~30 s; size the sample so the whole run ends in ~3 minutes
    L = int(min(C2_L, max(100_000, C2_L * 150.0 / (30.0 * total))))
    L -= L % 1000
    w = len(C2_SEED)
    with tempfile.TemporaryDirectory() as d:
        for _ in range(args.warmup):
            cpu_c2_step(L, d)
        t = 0.0
        for _ in range(args.steps):
            dt, kind, cores, counts = cpu_c2_step(L, d)
            t += dt
    value = 2 * (L - w + 1) * args.steps / t
    sample = "C2 pair at %d bases per genome (full workload: %d), predicate ALL, seed %s" % (L, C2_L, C2_SEED)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample,
                   "note": "reference CPU path: parallel_kmer_sets_from_fasta_files (unordered_map sets) + "
                           "kmer_set_intersection + containment/binomial_estimator"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "result": {"size_a": counts[0], "size_b": counts[1], "intersection": counts[2], "ani_ab": counts[3]},
    }))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import spaced_kmer_sketching_b200 as sks

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsks has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # one rank per GPU shares the host's few cores with the other ranks: no intra-op thread pools
        torch.set_num_threads(1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.Stream()
    ctx = sks.Context(local)
    ctx.set_stream(stream.cuda_stream)   # kernels launch on this torch stream: torch events time them
    mask, w = sks.seed_to_mask(C2_SEED)
    pred = sks.all_kmers()
    weight = sks.mask_weight(mask)
    L = C2_L
    gseed = 42 + 1000 * rank             # rank 0 is the KAT-4 pair
    batch = ctx.synth(L, [gseed, gseed], [0, 43 + 1000 * rank], [0, 100])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    kmers_per_step = 2 * (L - w + 1)
    launches0 = ctx.launches

    def step():
        return ctx.pair_ani_resident(batch, mask, w, pred, sks.REPR_BITSET)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()   # nvidia-smi needs ~0.2 s to its first sample: run it from warm-up to the last leg

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            r = step()
            flush.zero_()
        if rank == 0:
            got = (r.size_a, r.size_b, r.intersection)
            if got != KAT4_C2:
                raise SystemExit("C2 result %r differs from the reference's counts %r" % (got, KAT4_C2))
        barrier()
        ctx.profile(True)
        ctx.kernel_stats()
        launches1 = ctx.launches
        evs = []
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            r = step()
            e1.record(stream)
            evs.append((e0, e1))
            flush.zero_()                # L2 flush between timed steps (outside the event pairs)
        barrier()
        kstats = ctx.kernel_stats()
        ctx.profile(False)
        gpu_launches = ctx.launches - launches1
        total_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
        value = world * kmers_per_step * args.steps / (total_ms / 1e3)

        # ---- e2e: the C-ABI call with HOST buffers (pinned), H2D + compute + D2H inside the timed region
        wa = torch.from_numpy(batch.download(0).view(np.int32)).pin_memory()
        wb = torch.from_numpy(batch.download(1).view(np.int32)).pin_memory()

        def e2e_step():
            return ctx.pair_ani_ptr(wa.data_ptr(), L, wb.data_ptr(), L, mask, w, pred, sks.REPR_BITSET)

        for _ in range(args.warmup):
            e2e_step()
        barrier()
        in_place0 = ctx.in_place_calls
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            r2 = e2e_step()
        e1.record(stream)
        t1 = time.perf_counter()     # every call has returned its counts to the host: the K steps end here
        barrier()
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
        e2e_wall_ms = max_over_ranks((t1 - t0) * 1e3)
        e2e_ms = max(e2e_ms, e2e_wall_ms)  # the call returns counts to the host: wall clock is the honest one
        assert (r2.size_a, r2.size_b, r2.intersection) == (r.size_a, r.size_b, r.intersection)
        e2e_value = world * kmers_per_step * args.steps / (e2e_ms / 1e3)
        h2d = int(wa.numel() * 4 + wb.numel() * 4 + 2 * 32 + 8)    # packed bases + genome descriptors + segment ends
        d2h = 32                                                  # |A|, |B|, |A n B| and the region-overflow flag, four uint64

        # ---- roofline of the dominant kernel (per-launch CUDA-event times over the timed region) ----------
        peak, peak_src = measured_peak()
        bitset_bytes = (1 << (2 * weight)) // 8
        algo = {   # algorithmic bytes per launch, DESIGN.md "Kernels"
            "sketch_kernel": 2 * L * (0.25 + 4.0),         # 2-bit bases read + one 4-byte PEXT index written per k-mer
            "bitset_pair_build_kernel": 2 * L * 4.0 + 2 * bitset_bytes,  # indices read + both bitsets written once
            "fill_zero_kernel": 2 * bitset_bytes,          # both bitsets cleared by one launch
            "bitset_pair_counts_kernel": 2 * bitset_bytes,  # both bitsets read once (|A|, |B|, |A n B| in one pass)
        }
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath))
            except Exception:
                traffic = {}
        kernels = {}
        for name, (n, ms) in kstats.items():
            per = ms / n
            ent = {"launches_per_step": n / args.steps, "ms_per_launch": per, "share_of_step": ms / (total_ms if world == 1 else sum(a.elapsed_time(b) for a, b in evs))}
            if name in algo:
                ent["algorithmic_bytes"] = algo[name]
                ent["achieved_gbs"] = algo[name] / (per * 1e-3) / 1e9
                ent["frac"] = ent["achieved_gbs"] / peak
            kernels[name] = ent
        top = max((k for k in kernels if k in algo), key=lambda k: kstats[k][1])
        roofline = {"bound": "hbm", "kernel": top, "achieved": kernels[top]["achieved_gbs"], "peak": peak,
                    "unit": "GB/s", "frac": kernels[top]["frac"], "traffic": traffic.get(top), "peak_source": peak_src,
                    "kernels": kernels}

        # ---- cpu baseline (rank 0, N = 1 only): the reference on a bounded sample ------------------------
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            Ls = args.cpu_sample_bases
            with tempfile.TemporaryDirectory() as d:
                dt, kind, cores, counts = cpu_c2_step(Ls, d)
            if Ls == L and tuple(counts[:3]) != (r.size_a, r.size_b, r.intersection):
                raise SystemExit("CPU reference counts %r differ from the GPU's" % (counts[:3],))
            cpu = {"value": 2 * (Ls - w + 1) / dt, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "one C2 pass on a %d-base pair (full workload %d), %.1f s of wall time; the reference "
                             "parallelises over files only, so a pair uses 2 threads" % (Ls, L, dt)}

        extra = {}
        if cpu is not None and not args.no_extra:
            extra["reference_cpu_rates"] = cpu_fmh_rates()
        if not args.no_extra:
            extra.update(c2_variants(ctx, sks, torch, batch, mask, w, stream, barrier, max_over_ranks, peak, flush))
            extra.update(extra_legs(ctx, sks, torch, dist, rank, world, stream, barrier, max_over_ranks, peak))

    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "bases_per_step_per_gpu": 2 * L, "l2": "1 GiB of bitsets per step (> 126 MB L2) and a 256 MiB "
                       "flush write between timed steps", "parallelism": "genome pairs sharded over ranks, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "call": "sks_pair_ani (host packed genomes in, counts + ANI out)",
                    "host_to_device": ("the sketch kernel's bulk copies read the pinned host buffers in place, tile by tile "
                                       "(no separate copy)") if ctx.in_place_calls - in_place0 == args.steps
                                      else "cudaMemcpyAsync before the sketch kernel"},
            "gpu_launches": int(gpu_launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "result": {"size_a": r.size_a, "size_b": r.size_b, "intersection": r.intersection, "ani_ab": r.ani_ab,
                       "ani_ba": r.ani_ba}, "extra": extra,
        }))
    if world > 1:
        dist.destroy_process_group()


def c2_variants(ctx, sks, torch, batch, mask, w, stream, barrier, max_over_ranks, peak, flush):
    """The C2 pair through the other two routes of the library: (a) the general API -- sks_sketch builds the two
    bitset sets, sks_intersect re-reads them (bitset_pair_counts_kernel, K5 on its own) -- and (b) the pair pipeline
    with the bitsets kept on chip (SKS_REPR_BITSET_ONCHIP)."""
    out = {}
    pred = sks.all_kmers()
    bitset_bytes = (1 << (2 * sks.mask_weight(mask))) // 8

    def separate():
        sa, sb = ctx.sketch(batch, mask, w, pred, sks.REPR_BITSET)
        n = ctx.intersect(sa, sb)
        sa.close()
        sb.close()
        return n

    def onchip():
        return ctx.pair_ani_resident(batch, mask, w, pred, sks.REPR_BITSET_ONCHIP).intersection

    for name, fn in (("separate_build_then_intersect", separate), ("pair_pipeline_onchip", onchip)):
        for _ in range(3):
            fn()
            flush.zero_()
        barrier()
        ctx.profile(True)
        ctx.kernel_stats()
        evs = []
        reps = 10
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            n = fn()
            e1.record(stream)
            evs.append((e0, e1))
            flush.zero_()
        barrier()
        ks = ctx.kernel_stats()
        ctx.profile(False)
        ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs)) / reps
        per_rep = [a.elapsed_time(b) for a, b in evs]
        ent = {"ms_per_step": ms, "ms_per_step_median": sorted(per_rep)[len(per_rep) // 2], "ms_per_rep": [round(x, 3) for x in per_rep],
               "intersection": int(n),
               "kernels_ms": {k: v[1] / v[0] for k, v in ks.items()}}
        if "bitset_pair_counts_kernel" in ks:
            per = ks["bitset_pair_counts_kernel"][1] / ks["bitset_pair_counts_kernel"][0]
            ent["pair_counts_gbs"] = 2 * bitset_bytes / (per * 1e-3) / 1e9
            ent["pair_counts_frac_of_hbm"] = ent["pair_counts_gbs"] / peak
        out[name] = ent
    return {"c2_variants": out}


def extra_legs(ctx, sks, torch, dist, rank, world, stream, barrier, max_over_ranks, peak):
    """C3 (250 Mbp FMH sketching, position-sharded) and a C4-style all-vs-all (genomes sharded, NCCL
    all-gather of sketches, pair matrix tiled by rank).  Reported beside the headline, not as it."""
    import numpy as np
    from spaced_kmer_sketching_b200 import multi_gpu
    out = {}
    mask3, w3 = sks.seed_to_mask(C3_SEED)
    pred = sks.frac_min_hash(1, 200)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # C3: each rank sketches its slice (with a (w-1)-base halo) of ONE 250 Mbp sequence; weak-free (strong) scaling
    L3 = 250_000_000
    shard = multi_gpu.position_shard(L3, w3, rank, world)
    b3 = multi_gpu.synth_slice(ctx, L3, 7, shard, w3)
    def c3_step(mid=None):
        """Local sketch of the rank's slice; for N > 1 the global set on every rank: all-gather of the partial
        sketches' keys, sort + unique of the union (slices overlap by the halo only, duplicates are k-mers that
        occur in two slices)."""
        (loc,) = ctx.sketch(b3, mask3, w3, pred)
        if mid is not None:
            mid.record(stream)
        if world == 1:
            return loc, loc
        keys, _, kw = multi_gpu.keys_as_tensor([loc], torch)
        parts = multi_gpu.allgather_varlen(keys, world, dist)
        allk = torch.cat(parts)
        glob = ctx.set_from_device_keys(allk.data_ptr(), allk.numel() // kw, kw, mask3, w3, sorted_unique=False)
        return loc, glob

    for _ in range(2):
        loc, glob = c3_step()
        glob.close()
        if world > 1:
            loc.close()
    flush.zero_()
    barrier()
    ctx.profile(True)
    ctx.kernel_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(stream)
    c3_ev = []
    for _ in range(reps):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        s, glob = c3_step(eb)
        c3_ev.append((ea, eb))
        n_global = glob.kmer_set_size()
        if world > 1:
            glob.close()
        if _ != reps - 1:
            s.close()
    e1.record(stream)
    barrier()
    ks = ctx.kernel_stats()
    ctx.profile(False)
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    local_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in c3_ev)) / reps   # partial sketches only, no exchange
    n_local = s.kmer_set_size()
    sk_ms = ks["sketch_kernel"][1] / ks["sketch_kernel"][0]
    bases_local = shard[1] + w3 - 1
    out["c3_sketch"] = {"workload": "250 Mbp sequence, seed " + C3_SEED + ", FMH(200, nonce 1, Boost>=1.81), position-sharded, global set on every rank",
                        "bases_per_s": L3 / (ms / 1e3), "ms": ms, "scaling": "strong",
                        "bases_per_s_partial_sketches": L3 / (local_ms / 1e3), "ms_partial_sketches": local_ms,
                        "sketch_kernel_ms": sk_ms,
                        "sketch_kernel_gbs": bases_local * (0.25 + 8.0 / 200) / (sk_ms * 1e-3) / 1e9,
                        "sketch_kernel_frac_of_hbm": bases_local * (0.25 + 8.0 / 200) / (sk_ms * 1e-3) / 1e9 / peak,
                        "sketch_kernel_bases_per_s": bases_local / (sk_ms * 1e-3),
                        "local_sketch_size": int(n_local), "global_sketch_size": int(n_global),
                        "includes": "local sketch" + (" + NCCL all-gather of the partial sketches + sort-unique of the union on every rank" if world > 1 else "")}
    s.close()
    b3.close()

    # C4-style: G genomes per rank, all-gather, rank-tiled all-vs-all
    G = 125   # 125 genomes per GPU: 1000 genomes on 8 GPUs is BASELINE.json configs[3]
    Lg = 5_000_000
    ids = list(range(rank * G, (rank + 1) * G))
    Ds = [[0, 1000, 200, 100, 50, 20][g % 6] for g in ids]
    bg = ctx.synth(Lg, [1000] * G, [2000 + g for g in ids], Ds)
    res = None
    c4_counts = None     # the n x n count matrix is reused from one iteration to the next
    ctx.profile(True)
    c4_runs = []
    for it in range(6):   # one warm-up pass, then five timed ones: the pass with the median total is reported
        barrier()
        ctx.kernel_stats()
        t = {}
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(stream)
        local_sets = ctx.sketch(bg, mask3, w3, pred)
        e[1].record(stream)
        all_sets = multi_gpu.allgather_sets(ctx, local_sets, mask3, w3, rank, world, stream)
        e[2].record(stream)
        counts = c4_counts = multi_gpu.tiled_counts(ctx, all_sets, rank, world, c4_counts)   # every unordered block pair on one rank
        e[3].record(stream)
        # counts of every rank -> the full matrix everywhere, mirror, ANI of this rank's rows (host double pow)
        torch.cuda.synchronize()
        t_host = time.perf_counter()
        rows = multi_gpu.row_tile(len(all_sets), rank, world)
        mine = multi_gpu.exchange_blocks(counts, rank, world)   # this rank's complete rows: blocks swapped point to point
        first_sizes = np.repeat(np.diagonal(mine[:, rows[0]:rows[1]]), len(all_sets)).astype(np.int32)
        ani = sks.ani_from_counts(np.ascontiguousarray(mine).ravel(), first_sizes, sks.mask_weight(mask3))
        t_host = (time.perf_counter() - t_host) * 1e3
        if it == 0:   # the whole matrix is consistent: every entry evaluated exactly once, symmetric counts
            full = multi_gpu.mirror_counts(multi_gpu.gather_rows(counts, rows, world))
            assert (full >= 0).all() and (full == full.T).all() and np.array_equal(full[rows[0]:rows[1]], mine)
        assert ani.shape[0] == (rows[1] - rows[0]) * len(all_sets)
        barrier()
        res = [max_over_ranks(e[i].elapsed_time(e[i + 1])) for i in range(3)] + [max_over_ranks(t_host)]
        c4_kernels = {k: {"launches": v[0], "ms": v[1]} for k, v in ctx.kernel_stats().items()}
        if it > 0:
            c4_runs.append((sum(res), res, c4_kernels))
        n_total = len(all_sets)
        sizes = [x.kmer_set_size() for x in all_sets]
        # algorithmic bytes of this rank's intersections: (|A| + |B|) * key_bytes per unordered pair it
        # evaluated; a diagonal block is evaluated above the diagonal only (|A n A| = |A| costs nothing)
        sz = np.asarray(sizes, dtype=np.int64)
        key_bytes = 8 * (2 if w3 > 32 else 1)
        c4_pairs, c4_bytes = 0, 0
        for (r0, r1), (c0, c1) in multi_gpu.block_rects(n_total, rank, world):
            nr, nc = r1 - r0, c1 - c0
            if (r0, r1) == (c0, c1):
                c4_pairs += nr * (nr - 1) // 2
                c4_bytes += int(sz[r0:r1].sum()) * (nr - 1) * key_bytes
            else:
                c4_pairs += nr * nc
                c4_bytes += (int(sz[r0:r1].sum()) * nc + int(sz[c0:c1].sum()) * nr) * key_bytes
        for x in all_sets:
            x.close()
        if world > 1:
            for x in local_sets:
                x.close()
    ctx.profile(False)
    bg.close()

    # C5 (rank 0 of a 1-GPU run only): several random spaced seeds of weight 12..28 over 100 graded mutants of one
    # genome -- per seed: sketch, all-vs-all, and the error of the ANI estimate against the true substitution rate
    if world == 1:
        n5, L5 = 100, 5_000_000
        D5 = [[0, 1000, 200, 100, 50, 20][g % 6] for g in range(n5)]
        b5 = ctx.synth(L5, [1000] * n5, [2000 + g for g in range(n5)], D5)
        rows5 = []
        for k in (12, 16, 20, 24, 28):
            w5 = k + 10
            m5 = sks.generate_random_spaced_seed_mask(w5, k)      # the reference's (k+10, k, seed 0) masks
            for rep in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                sets5 = ctx.sketch(b5, m5, w5, pred)
                cnt5 = ctx.intersect_all_pairs(sets5)
                e1.record(stream)
                torch.cuda.synchronize()
                ms5 = e0.elapsed_time(e1)
                sizes5 = np.diag(cnt5).astype(np.int32)
                for x in sets5:
                    x.close()
            ani5 = sks.ani_from_counts(cnt5.ravel(), np.repeat(sizes5, n5), k).reshape(n5, n5)
            err = [abs(ani5[0, g] - (1.0 - 1.0 / D5[g])) for g in range(1, n5) if D5[g]]
            rows5.append({"weight": k, "window": w5, "ms_sketch_plus_all_pairs": ms5,
                          "mean_abs_ani_error_vs_true": float(np.mean(err)), "max_abs_ani_error": float(np.max(err)),
                          "mean_sketch_size": float(np.mean(sizes5))})
        b5.close()
        out["c5_multi_seed"] = {"workload": "%d synthetic 5 Mbp genomes (graded mutants of one base), random spaced seeds "
                                            "(k+10, k, seed 0) for k = 12..28, FMH(200); ANI(base, mutant) against 1 - 1/D" % n5,
                                "per_seed": rows5}
    c4_runs.sort(key=lambda r: r[0])
    total, res, c4_kernels = c4_runs[len(c4_runs) // 2]
    out["c4_all_vs_all"] = {"kernels": c4_kernels, "passes": "median of %d passes (totals %s ms)" % (
                                len(c4_runs), ", ".join("%.2f" % r[0] for r in c4_runs)),"workload": "%d synthetic 5 Mbp genomes (%d per GPU) at graded mutation rates, seed %s, "
                                        "FMH(200), all n^2 ordered pairs; every unordered block pair on one rank" % (n_total, G, C3_SEED),
                            "ani_pairs_per_s": n_total * n_total / (total / 1e3),
                            "ani_pairs_per_s_compare_only": n_total * n_total / (res[2] / 1e3),
                            "sketch_bases_per_s": n_total * Lg / (res[0] / 1e3),
                            "ms": {"sketch": res[0], "allgather": res[1], "intersect": res[2], "exchange_counts_and_ani": res[3]},
                            "mean_sketch_size": float(np.mean(sizes)), "scaling": "weak"}
    ik = c4_kernels.get("sorted_intersect_kernel")
    if ik and ik["ms"] > 0:
        gbs = c4_bytes / (ik["ms"] * 1e-3) / 1e9
        out["c4_all_vs_all"]["intersect_roofline"] = {
            "kernel": "row_intersect_kernel (+ sorted_intersect_kernel for rows that do not fit shared memory)",
            "unordered_pairs_on_rank0": c4_pairs, "algorithmic_bytes": c4_bytes, "ms": ik["ms"],
            "achieved_gbs": gbs, "frac_of_hbm": gbs / peak,
            "note": "(|A|+|B|) * key_bytes per unordered pair / kernel time; the sets are L2-resident, so this is "
                    "algorithmic bytes per second against the HBM peak, not DRAM traffic"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the C3 / C4 legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-bases", type=int, default=C2_L, help="genome length of the cpu_baseline sample")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
