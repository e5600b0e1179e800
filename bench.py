#!/usr/bin/env python
"""bench.py -- the hot path (sketch -> set -> intersect -> ANI) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" is one pass of the hot path over one batch of synthetic input.  The workload is the one BASELINE.json's
metric "all-vs-all ANI pairs/sec at 1/2/4/8 B200" is quoted on, configs[3] (C4): 1000 synthetic 5 Mbp genomes at
graded mutation rates, the reference's own (31,21,seed 0) spaced seed, FracMinHash(200, nonce 1), all n^2 ordered
pairs -- sketch every genome, exchange the sketches, intersect every pair, containment^(1/weight) ANI
(/root/reference/src/kmer-sketching.cpp:163-203: the two timed phases of the reference driver).  The SAME 1000
genomes are run at every N (strong scaling): genome g belongs to rank g // ceil(1000 / N), every rank returns its own
complete block rows of the 1000 x 1000 matrix, and the sha256 of the assembled count matrix must be the same at every N.

`value` times the step with the genomes already resident in HBM; `e2e` is the same step through the C-ABI call that
takes HOST buffers (sks_all_vs_all_from_host: packed genomes in pinned host memory in, counts + sizes + ANI rows out).
The other configurations are reported under "extra": C2 (configs[1], the HBM-bound bitset pair with its own roofline
block), C3 (configs[2], 250 Mbp position-sharded), C5 (configs[4], multi-seed sweep), the reference CLI end to end.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/libref.so: the unmodified reference
sources compiled against the Boost/Cilk shim; the oracle port if that is absent) with all host threads on a bounded
sample of the same workload.
"""
import argparse
import hashlib
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
C4_N, C4_L = 1000, 5_000_000
C4_DS = [0, 1000, 200, 100, 50, 20]           # genome g = mutate(gen(5 Mbp, 1000), 2000 + g, C4_DS[g % 6])
# sha256 of the 1000 x 1000 int32 count matrix and of the 1000 int32 set sizes (row-major, little endian), first
# produced by the single-GPU pairwise kernels (row_intersect / sorted_intersect), which the tests pin to the oracle
C4_MATRIX_SHA256 = "ec7dc05ecede49acb21cfc7ecec53acf2eca73dd6cab9bc10e0c8d25f28cbf6b"
C4_SIZES_SHA256 = "7ba010cc645a18f5b77855d0d23d40f5b4bb0d280d21e6280aff577985ca8e84"
WORKLOAD = ("C4 = BASELINE configs[3]: %d synthetic %d-base genomes at graded mutation rates (D = %s), weight-21 span-31 seed %s, "
            "FracMinHash(200, nonce 1, Boost >= 1.81 hash), all n^2 ordered pairs: sketch + exchange + intersect + "
            "containment^(1/21) ANI; the same genomes at every N (genome g on rank g // ceil(n/N))" % (C4_N, C4_L, C4_DS, C3_SEED))
METRIC = "all_vs_all_ani_pairs_per_s"
UNIT = "pairs/s"


def bench_config(n=C4_N):
    """The `config` of both arms (ours and --impl reference): what is computed, and how the GPU arm keeps the L2 cold."""
    return {"workload": WORKLOAD if n == C4_N else WORKLOAD.replace("%d synthetic" % C4_N, "%d (of the %d) synthetic" % (n, C4_N)),
            "genomes": n, "bases_per_step": n * C4_L, "ordered_pairs_per_step": n * n,
            "l2": "%.2f GB of packed genomes per step (> 126 MB L2) and a 256 MiB flush write between timed steps" % (n * C4_L / 4e9),
            "parallelism": "genomes sharded over ranks in contiguous blocks; every key travels once to the rank that owns it "
                           "(NCCL), one reduce-scatter of the partial counts; every rank returns its own block rows"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_constants():
    """ncu-derived per-kernel constants (DRAM traffic per launch, executed instructions per unit), profiles/traffic.json."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


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


def host_threads():
    return {"nproc": os.cpu_count(), "OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS")}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU path on the box's host cores
# ------------------------------------------------------------------------------------------------
def c4_genome(port, base, g):
    return base if C4_DS[g % 6] == 0 else port.mutate(base, 2000 + g, C4_DS[g % 6])


class CpuC4Sample:
    """A bounded sample of the C4 workload for the reference: n_s of its genomes (full length), written once as FASTA.
    One pass = parallel_kmer_sets_from_fasta_files (cilk_for over files) + parallel_compute_pairwise_kmer_set_
    intersections over generate_all_pairs_from_vector + containment / binomial_estimator, as the reference driver
    does (/root/reference/src/kmer-sketching.cpp:163-203), with all the host threads it can use."""

    def __init__(self, workdir, n_s, L=C4_L):
        from oracle import port, ref
        self.port, self.ref, self.n_s, self.L = port, ref, n_s, L
        self.mask, self.w = port.seed_to_mask(C3_SEED)
        self.weight = port.mask_weight(self.mask)
        self.threads = min(os.cpu_count() or 1, int(os.environ.get("OMP_NUM_THREADS") or (os.cpu_count() or 1)))
        base = port.gen(L, 1000)
        self.codes = [c4_genome(port, base, g) for g in range(n_s)]
        self.paths = []
        for g in range(n_s):
            self.paths.append(os.path.join(workdir, "g%d.fna" % g))
            port.write_fasta(self.paths[-1], self.codes[g], "g%d" % g)

    def step(self):
        """(seconds sketching, seconds comparing, counts[n_s, n_s], sizes[n_s])."""
        import numpy as np
        ref, port, n = self.ref, self.port, self.n_s
        if ref.available():
            t0 = time.perf_counter()
            sets = ref.sets_from_fasta_files(self.paths, self.mask, self.w, ref.FMH, 1, 200, parallel=True)
            t1 = time.perf_counter()
            first = [a for a in sets for _ in sets]
            second = [b for _ in sets for b in sets]
            ints = ref.pairwise_intersections(first, second, parallel=True).reshape(n, n)
            sizes = np.array([s.size() for s in sets], dtype=np.int32)
            for i in range(n):
                for j in range(n):
                    ref.binomial_estimator(ref.containment(int(ints[i, j]), int(sizes[i])), self.weight)
            t2 = time.perf_counter()
            return t1 - t0, t2 - t1, ints, sizes, "reference"
        t0 = time.perf_counter()
        sets = [port.sketch_set(c, [len(c)], self.mask, self.w, port.FMH, 1, 200, 181) for c in self.codes]
        t1 = time.perf_counter()
        ints = np.array([[port.intersection(a, b) for b in sets] for a in sets], dtype=np.int32)
        sizes = np.array([len(s) for s in sets], dtype=np.int32)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, ints, sizes, "port"

    def projected_pairs_per_s(self, t_sketch, t_compare):
        """The full workload's ordered pairs per second from the sample's per-genome and per-pair times (same threads)."""
        t_full = t_sketch / self.n_s * C4_N + t_compare / (self.n_s * self.n_s) * (C4_N * C4_N)
        return C4_N * C4_N / t_full, t_full

    def describe(self, t_sketch, t_compare, t_full):
        return ("%d of the %d genomes (full %d-base length) per pass: sketching %.2f s, all %d ordered pairs %.3f s on %d threads; "
                "value = n^2 / (n * t_genome + n^2 * t_pair) for n = %d, i.e. %.0f s for the full workload"
                % (self.n_s, C4_N, self.L, t_sketch, self.n_s ** 2, t_compare, self.threads, C4_N, t_full))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_s = max(2, min(16, cores))
    with tempfile.TemporaryDirectory() as d:
        sample = CpuC4Sample(d, n_s)
        for _ in range(args.warmup):
            sample.step()
        ts = tc = 0.0
        for _ in range(args.steps):
            a, b, ints, sizes, kind = sample.step()
            ts += a
            tc += b
    ts /= args.steps
    tc /= args.steps
    value, t_full = sample.projected_pairs_per_s(ts, tc)
    text = sample.describe(ts, tc, t_full)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_full, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": bench_config(),
        "reference_note": "reference CPU path: parallel_kmer_sets_from_fasta_files (FASTA parse, sliding window, FracMinHash, "
                          "unordered_map sets) + parallel_compute_pairwise_kmer_set_intersections + containment/binomial_estimator; "
                          "bounded sample: " + text,
        "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": sample.threads, "kind": kind, "sample": text}, **host_threads()),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "sample_ms_per_step": 1e3 * (ts + tc),
        "result": {"sample_counts_sha256": hashlib.sha256(ints.astype("<i4").tobytes()).hexdigest(), "sample_sizes": sizes.tolist()},
    }))


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


def cli_wall_times():
    """The reference's own main() (62 configurations, CSV out; /root/reference/src/kmer-sketching.cpp:214-240) on 4 x 2 Mbp
    FASTA files: wall time of oracle/_ref/ref_cli (the reference) against dropin_cli (the reference's unmodified main()
    compiled on include/*.hpp + libsks) and sks_cli (the rewritten harness); the CSVs must be byte-identical."""
    from oracle import port, ref
    exe = {"ref_cli": ref.CLI_PATH, "dropin_cli": os.path.join(ROOT, "oracle", "_ref", "dropin_cli"),
           "sks_cli": os.path.join(ROOT, "spaced_kmer_sketching_b200", "sks_cli")}
    if not all(os.path.exists(p) for p in exe.values()):
        return {"unavailable": "needs " + ", ".join(k for k, p in exe.items() if not os.path.exists(p))}
    out = {"files": "4 x 2 Mbp (graded mutants of one genome), 62 configurations x 16 ordered pairs"}
    with tempfile.TemporaryDirectory() as d:
        base = port.gen(2_000_000, 1000)
        paths = []
        for g in range(4):
            paths.append(os.path.join(d, "g%d.fna" % g))
            port.write_fasta(paths[-1], base if g == 0 else port.mutate(base, 2000 + g, [0, 1000, 100, 20][g]), "g%d" % g)
        csv = {}
        for name, path in exe.items():
            dst = os.path.join(d, name + ".csv")
            t0 = time.perf_counter()
            r = subprocess.run([path, dst] + paths, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                               env=dict(os.environ, SKS_PREDICATE_PROBE="1"))
            out[name + "_wall_s"] = time.perf_counter() - t0
            if r.returncode != 0:
                out[name + "_error"] = (r.stderr or r.stdout)[-300:]
                continue
            csv[name] = open(dst, "rb").read()
            phases = [float(ln.split("=")[1].split("ms")[0]) for ln in r.stdout.splitlines() if "Time taken" in ln and "=" in ln]
            out[name + "_sketching_ms"] = sum(phases[0::2])
            out[name + "_comparison_ms"] = sum(phases[1::2])
        out["csv_identical"] = len(csv) == 3 and len(set(csv.values())) == 1
        out["csv_sha256"] = hashlib.sha256(csv.get("ref_cli", b"")).hexdigest()
    return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import spaced_kmer_sketching_b200 as sks
    from spaced_kmer_sketching_b200 import multi_gpu

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsks has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        torch.set_num_threads(1)   # one rank per GPU shares the host's cores with the other ranks
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
    comm = multi_gpu.init_comm(ctx, rank, world, dist if world > 1 else None)   # the C ABI's communicator (NCCL inside libsks)
    peak, peak_src = measured_peak()
    consts = profile_constants()
    sampler = ClockSampler(local)

    mask, w = sks.seed_to_mask(C3_SEED)
    weight = sks.mask_weight(mask)
    pred = sks.frac_min_hash(1, 200)
    n = args.genomes
    b, e = sks.shard_range(n, rank, world)
    n_loc = e - b
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    with torch.cuda.stream(stream):
        ids = list(range(b, e))
        batch = ctx.synth(C4_L, [1000] * n_loc, [2000 + g for g in ids], [C4_DS[g % 6] for g in ids])
        # results land in pinned host memory (the D2H of the rank's rows is part of the step)
        pin_counts = torch.empty((max(n_loc, 1), n), dtype=torch.int32).pin_memory()
        pin_sizes = torch.empty(n, dtype=torch.int32).pin_memory()
        pin_ani = torch.empty((max(n_loc, 1), n), dtype=torch.float64).pin_memory()
        out = (pin_counts.numpy()[:n_loc], pin_sizes.numpy(), pin_ani.numpy()[:n_loc])

        def step():   # sks_sketch + sks_comm_allgather_sets + sks_all_vs_all for the rank's rows, one C call
            return ctx.all_vs_all_resident(comm, batch if n_loc else None, n, mask, w, pred, out)

        if rank == 0:
            sampler.start()   # nvidia-smi needs ~0.2 s to its first sample: from the warm-up to the end of the timed regions
        for _ in range(args.warmup):
            step()
            flush.zero_()
        barrier()
        ctx.profile(True)
        ctx.kernel_stats()
        launches1 = ctx.launches
        evs = []
        t_wall = time.perf_counter()
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step()
            e1.record(stream)
            evs.append((e0, e1))
            flush.zero_()                # L2 flush between timed steps (outside the event pairs)
        barrier()
        t_wall = time.perf_counter() - t_wall
        kstats = ctx.kernel_stats()
        ctx.profile(False)
        gpu_launches = ctx.launches - launches1
        own_ms = sum(a.elapsed_time(b) for a, b in evs)
        total_ms = max_over_ranks(own_ms)
        value = n * n * args.steps / (total_ms / 1e3)
        counts0, sizes0, ani0 = out[0].copy(), out[1].copy(), out[2].copy()

        # ---- parity: the assembled matrix must be the same at every N; rank 0 re-derives its rows with the pairwise
        # kernels of the single-GPU path and two sketches with the oracle
        parity = parity_check(args, sks, np, torch, dist, ctx, comm, batch, mask, w, pred, rank, world, n, b, e, counts0, sizes0, ani0)

        # ---- e2e: the C-ABI call with HOST buffers (pinned), H2D + compute + D2H inside the timed region
        words = C4_L // 16 + (1 if C4_L % 16 else 0)
        stride = (words + 3) // 4 * 4               # 16-byte aligned genome starts
        host = torch.empty(max(n_loc, 1) * stride, dtype=torch.int32).pin_memory()
        hnp = host.numpy().view(np.uint32)
        for g in range(n_loc):
            hnp[g * stride:g * stride + words] = batch.download(g)
        ptrs = [host.data_ptr() + 4 * g * stride for g in range(n_loc)]
        nb = [C4_L] * n_loc

        def e2e_step():
            return ctx.all_vs_all_from_host(comm, ptrs, nb, n, mask, w, pred, out)

        for _ in range(args.warmup):
            e2e_step()
            flush.zero_()
        barrier()
        in_place0, streamed0 = ctx.in_place_calls, ctx.streamed_calls
        e2e_ev, e2e_wall = [], 0.0
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(stream)
            e2e_step()
            e1.record(stream)
            e2e_wall += time.perf_counter() - t0     # the call has returned the rank's rows to the host
            e2e_ev.append((e0, e1))
            flush.zero_()
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        e2e_ms = max_over_ranks(max(sum(a.elapsed_time(b) for a, b in e2e_ev), e2e_wall * 1e3))
        assert np.array_equal(out[0], counts0) and np.array_equal(out[1], sizes0)
        e2e_value = n * n * args.steps / (e2e_ms / 1e3)
        in_place = ctx.in_place_calls - in_place0 == args.steps
        streamed = ctx.streamed_calls - streamed0 == args.steps
        h2d = int(n_loc * words * 4 + n_loc * (32 + 4 + 4))       # packed bases + genome descriptors + segment ends + tile map (per rank)
        d2h = int(n_loc * n * 12 + n * 4)                         # the rank's rows: int32 counts + float64 ANI, and the n sizes
        # pageable host buffers: staged through pinned ring buffers by host threads, chunk by chunk under the sketch kernel
        pageable = np.array(hnp, copy=True)
        pptrs = [pageable.ctypes.data + 4 * g * stride for g in range(n_loc)]
        pg_wall = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            ctx.all_vs_all_from_host(comm, pptrs, nb, n, mask, w, pred, out)
            pg_wall.append(time.perf_counter() - t0)
            flush.zero_()
        e2e_pageable_ms = max_over_ranks(sorted(pg_wall)[1] * 1e3)
        del pageable, host

        # ---- roofline of the step's kernels (per-launch CUDA-event times over the timed region) ----------
        bases_rank = n_loc * C4_L
        n_keys_all = int(sizes0.sum())
        n_keys_rank = int(sizes0[b:e].sum())
        pairs_rank_unordered = (n * (n - 1) // 2) if world == 1 else n_loc * (n - 1)
        sz = sizes0.astype(np.int64)
        if world == 1:
            pair_bytes = int(8 * (sz.sum() * (n - 1)))                       # sum over unordered pairs of (|A| + |B|) * 8
        else:
            pair_bytes = int(8 * (sz[b:e].sum() * (n - 1) + n_loc * (sz.sum()) - (sz[b:e].sum())))
        algo = {   # algorithmic bytes per launch on this rank (DESIGN.md "Kernels")
            "sketch_kernel": bases_rank * 0.25 + n_keys_rank * 8.0 * 1.0,     # 2-bit bases read + 8 B per kept k-mer written
            "sort_unique": n_keys_rank * 8.0 * 2,                             # raw keys read, distinct keys written
            "dict_build": n_keys_all * (8.0 + 4.0 + 4.0 + 2.0),               # keys read, slot/id per key written + read, re-coded payload written
            "allpairs_kernel": float(pair_bytes),                             # (|A| + |B|) * key_bytes per unordered pair evaluated
            "ani_finalize_kernel": n_loc * n * (4.0 + 4.0 + 8.0),
        }
        step_ms_own = own_ms / args.steps
        kernels = {}
        for name, (cnt, ms) in kstats.items():
            per = ms / cnt
            ent = {"launches_per_step": cnt / args.steps, "ms_per_launch": per, "ms_per_step": ms / args.steps,
                   "share_of_step": ms / own_ms}
            if name in algo and cnt == args.steps:
                ent["algorithmic_bytes"] = algo[name]
                ent["achieved_gbs"] = algo[name] / (per * 1e-3) / 1e9
                ent["frac"] = ent["achieved_gbs"] / peak
            tr = consts.get(name + ":c4") or consts.get(name)
            if isinstance(tr, dict):
                ent.update({k: v for k, v in tr.items() if k in ("dram_bytes_per_launch", "note")})
            kernels[name] = ent
        top = max((k for k in kernels if "frac" in kernels[k]), key=lambda k: kstats[k][1])
        clocks_mhz = 1965.0
        roofline = {"bound": "hbm", "kernel": top, "achieved": kernels[top]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[top]["frac"], "traffic": (consts.get(top + ":c4") or {}).get("dram_bytes_per_launch"),
                    "peak_source": peak_src, "kernels": kernels}
        ipw = (consts.get("sketch_kernel:c4") or {}).get("warp_instructions_per_window")
        if "sketch_kernel" in kernels and ipw:
            # the FMH sketch kernel is bound by integer issue, not by HBM: executed warp-instructions
            # against the issue capacity of 148 SMs x 4 schedulers
            windows = n_loc * (C4_L - w + 1)
            t = kernels["sketch_kernel"]["ms_per_launch"] * 1e-3
            roofline["issue_frac"] = (windows / 32.0 * ipw) / (148 * 4 * clocks_mhz * 1e6 * t)
            roofline["issue_note"] = ("sketch_kernel: %.0f executed warp-instructions per 32 windows (ncu, profiles/) over "
                                      "148 SMs x 4 schedulers x %.0f MHz" % (ipw, clocks_mhz))
        phases = {"sketch": sum(kernels.get(k, {}).get("ms_per_step", 0.0) for k in ("sketch_kernel", "sort_unique")),
                  "exchange": kernels.get("nccl_exchange", {}).get("ms_per_step", 0.0),
                  "dictionary": kernels.get("dict_build", {}).get("ms_per_step", 0.0),
                  "intersect": kernels.get("allpairs_kernel", {}).get("ms_per_step", 0.0),
                  "ani": kernels.get("ani_finalize_kernel", {}).get("ms_per_step", 0.0)}
        phases["host_and_copies"] = max(step_ms_own - sum(phases.values()), 0.0)
        headline_detail = {"ms_phases_rank0": phases, "limiting_phase": max(phases, key=phases.get),
                           "sketch_bases_per_s": n * C4_L / (max_over_ranks(phases["sketch"]) / 1e3),
                           "compare_only_pairs_per_s": n * n / (max_over_ranks(step_ms_own - phases["sketch"]) / 1e3),
                           "mean_sketch_size": float(sizes0.mean()), "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps}

        # ---- cpu baseline (rank 0, N = 1 only): the reference on a bounded sample ------------------------
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            n_s = max(2, min(8, os.cpu_count() or 1))
            with tempfile.TemporaryDirectory() as d:
                sample = CpuC4Sample(d, n_s)
                ts, tc, ints, sizes_s, kind = sample.step()
            if n >= n_s and not (np.array_equal(ints, counts0[:n_s, :n_s]) and np.array_equal(sizes_s, sizes0[:n_s])):
                raise SystemExit("the reference's counts on the first %d genomes differ from the GPU's" % n_s)
            v, t_full = sample.projected_pairs_per_s(ts, tc)
            cpu = dict({"value": v, "unit": UNIT, "cores": sample.threads, "kind": kind, "sample": sample.describe(ts, tc, t_full),
                        "counts_equal_gpu": True}, **host_threads())

        extra = {"c4_detail": headline_detail}
        if not args.no_extra:
            if world == 1:
                extra.update(c2_pair_leg(args, ctx, sks, torch, stream, barrier, peak, peak_src, flush, consts))
                if not args.no_cpu_baseline:
                    extra["reference_cli"] = cli_wall_times()
                    extra.update(fasta_leg(ctx, sks, np))
            extra.update(c3_leg(ctx, comm, sks, multi_gpu, torch, rank, world, stream, barrier, max_over_ranks, peak))
            # the same path where the sequence is long enough for the sketching itself to dominate the fixed costs
            extra.update(c3_leg(ctx, comm, sks, multi_gpu, torch, rank, world, stream, barrier, max_over_ranks, peak,
                                L3=2_000_000_000, name="c3_sketch_2gbp"))
            if world == 1:
                extra.update(c5_leg(ctx, sks, np, torch, stream))

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": bench_config(n),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": e2e_ms / args.steps,
                    "call": "sks_all_vs_all_from_host (host packed genomes in; counts, sizes and ANI rows out)",
                    "host_to_device": ("the copy engine brings the pinned host buffers 16 MB at a time while the sketch kernel "
                                       "works on the chunks that have arrived") if streamed else
                                      ("the sketch kernel's bulk copies read the pinned host buffers in place, tile by tile "
                                       "(no separate copy)") if in_place else "cudaMemcpyAsync before the sketch kernel",
                    "pageable_host_buffers_ms_per_step": e2e_pageable_ms,
                    "l2": "256 MiB flush write between steps"},
            "gpu_launches": int(gpu_launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "parity_check": parity,
            "result": {"counts_0_1": int(counts0[0, 1]) if n > 1 and rank == 0 else None, "size_0": int(sizes0[0]),
                       "ani_0_1": float(ani0[0, 1]) if n > 1 else None},
            "extra": extra,
        }))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def parity_check(args, sks, np, torch, dist, ctx, comm, batch, mask, w, pred, rank, world, n, b, e, counts0, sizes0, ani0):
    """On-hardware parity of the run that was just timed."""
    n_loc = e - b
    # assemble the matrix from every rank's rows (blocks are padded to the largest)
    per = sks.shard_range(n, 0, world)[1]
    mine = torch.zeros((per, n), dtype=torch.int32, device="cuda")
    mine[:n_loc] = torch.from_numpy(counts0).cuda()
    if world > 1:
        full = torch.empty((world * per, n), dtype=torch.int32, device="cuda")
        dist.all_gather_into_tensor(full, mine)
        rows = []
        for r in range(world):
            rb, re_ = sks.shard_range(n, r, world)
            rows.append(full[r * per:r * per + (re_ - rb)])
        matrix = torch.cat(rows).cpu().numpy()
    else:
        matrix = counts0
    out = {"matrix_sha256": hashlib.sha256(np.ascontiguousarray(matrix, dtype="<i4").tobytes()).hexdigest(),
           "sizes_sha256": hashlib.sha256(np.ascontiguousarray(sizes0, dtype="<i4").tobytes()).hexdigest(),
           "symmetric": bool((matrix == matrix.T).all()), "diagonal_is_sizes": bool((np.diag(matrix) == sizes0).all())}
    if n == C4_N and C4_MATRIX_SHA256:
        out["matches_recorded_sha256"] = out["matrix_sha256"] == C4_MATRIX_SHA256 and out["sizes_sha256"] == C4_SIZES_SHA256
    # rank 0 (every rank, in fact): its block rows again, by the pairwise kernels on the gathered sets
    sets = ctx.sketch(batch, mask, w, pred, sks.REPR_SORTED) if n_loc else []
    everything = ctx.allgather_sets(comm, sets, n)
    sub = min(n_loc, 32)     # 32 rows x n columns through row_intersect_kernel / sorted_intersect_kernel
    want = np.full((n, n), -1, dtype=np.int32)
    if sub:
        ctx.intersect_block(everything, (b, b + sub), (0, n), want)
    rows_ok = bool(np.array_equal(want[b:b + sub], counts0[:sub]))
    host_ani = sks.ani_from_counts(counts0.ravel(), np.repeat(sizes0[b:e], n), sks.mask_weight(mask)) if n_loc else np.zeros(0)
    ani_err = float(np.max(np.abs(ani0.ravel() - host_ani))) if n_loc else 0.0
    oracle_ok = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import port
        base = port.gen(C4_L, 1000)
        oracle_ok = True
        for g in sorted({0, min(5, n - 1)}):
            okeys = port.sketch_set(c4_genome(port, base, g), [C4_L], mask, w, port.FMH, 1, 200, 181)
            oracle_ok = oracle_ok and bool(np.array_equal(everything[g].keys(), okeys))
    for s in everything + list(sets):
        s.close()
    if world > 1:
        flags = torch.tensor([int(rows_ok), int(ani_err <= 1e-12)], dtype=torch.int32, device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        rows_ok, ani_ok = bool(flags[0].item()), bool(flags[1].item())
    else:
        ani_ok = ani_err <= 1e-12
    out.update({"rows_match_pairwise_kernels": rows_ok, "rows_checked_per_rank": sub, "ani_within_1e-12_of_host_libm": ani_ok,
                "max_abs_ani_error_rank0": ani_err, "oracle_sets_match": oracle_ok})
    bad = [k for k in ("symmetric", "diagonal_is_sizes", "rows_match_pairwise_kernels", "ani_within_1e-12_of_host_libm") if not out[k]]
    if oracle_ok is False:
        bad.append("oracle_sets_match")
    if out.get("matches_recorded_sha256") is False:
        bad.append("matches_recorded_sha256")
    if bad:
        raise SystemExit("parity check failed: %s (%r)" % (", ".join(bad), out))
    return out


def c2_pair_leg(args, ctx, sks, torch, stream, barrier, peak, peak_src, flush, consts):
    """BASELINE configs[1] (C2): a 5 Mbp genome and its 1%-mutated copy, the (24,16,seed 0) seed, predicate ALL, a full
    4^16-bit presence bitset per genome (512 MiB each, written to HBM), AND/popcount, ANI -- the HBM-bound leg, with its
    own roofline block.  One step = sks_pair_ani_resident; e2e = sks_pair_ani from pinned host buffers."""
    import numpy as np
    mask, w = sks.seed_to_mask(C2_SEED)
    pred = sks.all_kmers()
    weight = sks.mask_weight(mask)
    L = C2_L
    steps = max(args.steps, 10)
    batch = ctx.synth(L, [42, 42], [0, 43], [0, 100])
    kmers_per_step = 2 * (L - w + 1)

    def step():
        return ctx.pair_ani_resident(batch, mask, w, pred, sks.REPR_BITSET)

    for _ in range(3):
        r = step()
        flush.zero_()
    got = (r.size_a, r.size_b, r.intersection)
    if got != KAT4_C2:
        raise SystemExit("C2 result %r differs from the reference's counts %r" % (got, KAT4_C2))
    barrier()
    ctx.profile(True)
    ctx.kernel_stats()
    evs = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r = step()
        e1.record(stream)
        evs.append((e0, e1))
        flush.zero_()
    barrier()
    kstats = ctx.kernel_stats()
    ctx.profile(False)
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    wa = torch.from_numpy(batch.download(0).view(np.int32)).pin_memory()
    wb = torch.from_numpy(batch.download(1).view(np.int32)).pin_memory()

    def e2e_step():
        return ctx.pair_ani_ptr(wa.data_ptr(), L, wb.data_ptr(), L, mask, w, pred, sks.REPR_BITSET)

    for _ in range(3):
        e2e_step()
        flush.zero_()
    barrier()
    e2e_wall = 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        r2 = e2e_step()
        e2e_wall += time.perf_counter() - t0
        flush.zero_()            # L2 flush between e2e steps as well
    assert (r2.size_a, r2.size_b, r2.intersection) == got
    pa, pb = np.array(wa.numpy(), copy=True), np.array(wb.numpy(), copy=True)     # pageable copies
    pg = []
    for _ in range(5):
        t0 = time.perf_counter()
        ctx.pair_ani_ptr(pa.ctypes.data, L, pb.ctypes.data, L, mask, w, pred, sks.REPR_BITSET)
        pg.append(time.perf_counter() - t0)
        flush.zero_()
    bitset_bytes = (1 << (2 * weight)) // 8
    algo = {"sketch_kernel": 2 * L * (0.25 + 4.0), "bitset_pair_build_kernel": 2 * L * 4.0 + 2 * bitset_bytes}
    kernels = {}
    for name, (cnt, ms) in kstats.items():
        per = ms / cnt
        ent = {"launches_per_step": cnt / steps, "ms_per_launch": per, "share_of_step": ms / total_ms}
        if name in algo:
            ent["algorithmic_bytes"] = algo[name]
            ent["achieved_gbs"] = algo[name] / (per * 1e-3) / 1e9
            ent["frac"] = ent["achieved_gbs"] / peak
        kernels[name] = ent
    top = max((k for k in kernels if k in algo), key=lambda k: kstats[k][1])
    out = {"workload": "C2 = BASELINE configs[1]: synthetic 5 Mbp genome vs 1%-mutated copy, weight-16 span-24 seed " + C2_SEED +
                       ", predicate ALL, one 4^16-bit presence bitset per genome (512 MiB, written to HBM), AND/popcount, containment^(1/16) ANI",
           "metric": "spaced_kmers_per_s_sketch_plus_ani", "value": kmers_per_step * steps / (total_ms / 1e3), "unit": "kmers/s",
           "ms_per_step": total_ms / steps, "steps": steps,
           "e2e": {"value": kmers_per_step * steps / e2e_wall, "ms_per_step": 1e3 * e2e_wall / steps,
                   "call": "sks_pair_ani (pinned host buffers read in place by the sketch kernel; L2 flushed between steps)",
                   "pageable_host_buffers_ms_per_step": 1e3 * sorted(pg)[len(pg) // 2],
                   "h2d_bytes_per_step": int(wa.numel() * 4 + wb.numel() * 4 + 72), "d2h_bytes_per_step": 32},
           "roofline": {"bound": "hbm", "kernel": top, "achieved": kernels[top]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                        "frac": kernels[top]["frac"], "traffic": (consts.get(top) if not isinstance(consts.get(top), dict) else
                                                                  consts.get(top, {}).get("dram_bytes_per_launch")),
                        "peak_source": peak_src, "kernels": kernels},
           "result": {"size_a": r.size_a, "size_b": r.size_b, "intersection": r.intersection, "ani_ab": r.ani_ab,
                      "matches_reference_counts": True}}
    # the route a drop-in user of the reference API gets: sks_sketch x 2 -> two sets -> sks_intersect
    def separate():
        sa, sb = ctx.sketch(batch, mask, w, pred, sks.REPR_BITSET)
        m = ctx.intersect(sa, sb)
        sa.close()
        sb.close()
        return m

    def onchip():
        return ctx.pair_ani_resident(batch, mask, w, pred, sks.REPR_BITSET_ONCHIP).intersection

    variants = {}
    for name, fn in (("api_route_sketch_then_intersect", separate), ("pair_pipeline_onchip", onchip)):
        for _ in range(3):
            fn()
            flush.zero_()
        barrier()
        ctx.profile(True)
        ctx.kernel_stats()
        evs, reps = [], 10
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            m = fn()
            e1.record(stream)
            evs.append((e0, e1))
            flush.zero_()
        barrier()
        ks = ctx.kernel_stats()
        ctx.profile(False)
        per_rep = sorted(a.elapsed_time(b) for a, b in evs)
        ent = {"ms_per_step": sum(per_rep) / reps, "ms_per_step_median": per_rep[len(per_rep) // 2], "intersection": int(m),
               "kernels_ms": {k: v[1] / v[0] for k, v in ks.items()}}
        if "bitset_pair_counts_kernel" in ks:
            per = ks["bitset_pair_counts_kernel"][1] / ks["bitset_pair_counts_kernel"][0]
            ent["pair_counts_gbs"] = 2 * bitset_bytes / (per * 1e-3) / 1e9
            ent["pair_counts_frac_of_hbm"] = ent["pair_counts_gbs"] / peak
        variants[name] = ent
    out["variants"] = variants
    cpu = None
    if not args.no_cpu_baseline and args.c2_cpu_sample_bases > 0:
        Ls = args.c2_cpu_sample_bases
        with tempfile.TemporaryDirectory() as d:
            dt, kind, cores, counts = cpu_c2_step(Ls, d)
        if Ls == L and tuple(counts[:3]) != got:
            raise SystemExit("CPU reference counts %r differ from the GPU's" % (counts[:3],))
        cpu = dict({"value": 2 * (Ls - w + 1) / dt, "unit": "kmers/s", "cores": cores, "kind": kind,
                    "sample": "one C2 pass on a %d-base pair (full workload %d), %.1f s of wall time; the reference parallelises "
                              "over files only, so a pair uses 2 threads" % (Ls, L, dt)}, **host_threads())
    out["cpu_baseline"] = cpu
    batch.close()
    return {"c2_pair": out}


def c3_leg(ctx, comm, sks, multi_gpu, torch, rank, world, stream, barrier, max_over_ranks, peak, L3=250_000_000, name="c3_sketch"):
    """BASELINE configs[2] (C3): ONE 250 Mbp sequence, weight-21 span-31 seed, FMH(200): window starts split over the
    ranks (a (w-1)-base halo each); the partial sketches are routed by key range, every rank sort-uniques its range
    (sks_sketch_sequence_sharded).  Strong scaling."""
    mask3, w3 = sks.seed_to_mask(C3_SEED)
    pred = sks.frac_min_hash(1, 200)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shard = multi_gpu.position_shard(L3, w3, rank, world)
    b3 = multi_gpu.synth_slice(ctx, L3, 7, shard, w3)
    res = {}
    for gather in (False, True):
        for _ in range(2):
            s, n_global = ctx.sketch_sequence_sharded(comm, b3, mask3, w3, pred, gather=gather)
            s.close()
        flush.zero_()
        barrier()
        ctx.profile(True)
        ctx.kernel_stats()
        reps, evs = 5, []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            s, n_global = ctx.sketch_sequence_sharded(comm, b3, mask3, w3, pred, gather=gather)
            e1.record(stream)
            evs.append((e0, e1))
            n_mine = s.kmer_set_size()
            s.close()
            flush.zero_()
        barrier()
        ks = ctx.kernel_stats()
        ctx.profile(False)
        ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs)) / reps
        res[gather] = (ms, n_global, n_mine, ks)
    ms_r, n_global, n_mine, ks = res[False]
    sk_ms = ks["sketch_kernel"][1] / ks["sketch_kernel"][0]
    bases_local = shard[1] + w3 - 1
    out = {"workload": ("C3 = BASELINE configs[2]: one 250 Mbp sequence" if L3 == 250_000_000 else
                        "the C3 kind of work at %.1f Gbp (one sequence)" % (L3 / 1e9)) + ", seed " + C3_SEED +
                       ", FMH(200, nonce 1, Boost>=1.81), window starts split over the ranks, the kept k-mers routed by key range, "
                       "every rank sort-uniques its range (sks_sketch_sequence_sharded)",
           "scaling": "strong", "bases_per_s": L3 / (ms_r / 1e3), "ms": ms_r,
           "result": "every rank holds its key range of the global set (disjoint, ordered by rank)",
           "bases_per_s_global_set_on_every_rank": L3 / (res[True][0] / 1e3), "ms_global_set_on_every_rank": res[True][0],
           "sketch_kernel_ms": sk_ms, "sketch_kernel_bases_per_s": bases_local / (sk_ms * 1e-3),
           "sketch_kernel_gbs": bases_local * (0.25 + 8.0 / 200) / (sk_ms * 1e-3) / 1e9,
           "sketch_kernel_frac_of_hbm": bases_local * (0.25 + 8.0 / 200) / (sk_ms * 1e-3) / 1e9 / peak,
           "kernels_ms_per_step": {k: v[1] / 5 for k, v in ks.items()},
           "global_sketch_size": int(n_global), "keys_in_rank0_range": int(n_mine)}
    b3.close()
    del flush
    return {name: out}


def fasta_leg(ctx, sks, np):
    """FASTA ingest.  8 x 5 Mbp files: parse + split at non-ACGT + 2-bit pack on the device
    (sks_batch_from_fasta_text: raw bytes up, kernels do the rest), libsks' own host parser, and the reference's
    nucleotide_strings_from_fasta_file (/root/reference/src/fasta_processing.cpp:79-211), one thread each; then files ->
    kmer_sets (device parse + sketch) against the reference's parallel_kmer_sets_from_fasta_files on all threads."""
    from oracle import port, ref
    n, L = 8, 5_000_000
    mask, w = sks.seed_to_mask(C3_SEED)
    out = {"files": "%d x %d bases, 80 columns, LF" % (n, L)}
    with tempfile.TemporaryDirectory() as d:
        base = port.gen(L, 1000)
        paths = []
        for g in range(n):
            paths.append(os.path.join(d, "g%d.fna" % g))
            port.write_fasta(paths[-1], c4_genome(port, base, g), "g%d" % g)
        texts = [open(p, "rb").read() for p in paths]
        best = {}
        for rep in range(3):
            ctx.profile(True)
            ctx.kernel_stats()
            t0 = time.perf_counter()
            b = ctx.batch_from_fasta_text(texts)
            ctx.sync()
            t1 = time.perf_counter()
            ks = ctx.kernel_stats()
            ctx.profile(False)
            t2 = time.perf_counter()
            sets = sks.kmer_sets_from_fasta_files(ctx, paths, mask, w, sks.frac_min_hash(1, 200))
            t3 = time.perf_counter()
            sizes = [s.kmer_set_size() for s in sets]
            for s in sets:
                s.close()
            b.close()
            best["device_ingest_ms"] = min(best.get("device_ingest_ms", 1e9), (t1 - t0) * 1e3)
            best["device_ingest_kernels_ms"] = min(best.get("device_ingest_kernels_ms", 1e9), sum(v[1] for v in ks.values()))
            best["files_to_sets_ms"] = min(best.get("files_to_sets_ms", 1e9), (t3 - t2) * 1e3)
        t0 = time.perf_counter()
        for p in paths[:2]:
            sks.fasta_parse_file(p)
        best["libsks_host_parser_bases_per_s_1_thread"] = 2 * L / (time.perf_counter() - t0)
        if ref.available():
            t0 = time.perf_counter()
            ref.Strings.from_fasta(paths[0])
            best["reference_parser_bases_per_s_1_thread"] = L / (time.perf_counter() - t0)
            t0 = time.perf_counter()
            rsets = ref.sets_from_fasta_files(paths, mask, w, ref.FMH, 1, 200, parallel=True)
            best["reference_files_to_sets_ms"] = (time.perf_counter() - t0) * 1e3
            best["sizes_equal_reference"] = [s.size() for s in rsets] == sizes
    best["device_ingest_bases_per_s"] = n * L / (best["device_ingest_ms"] / 1e3)
    best["device_ingest_kernels_bases_per_s"] = n * L / (best["device_ingest_kernels_ms"] / 1e3)
    out.update(best)
    return {"fasta_ingest": out}


def c5_leg(ctx, sks, np, torch, stream):
    """BASELINE configs[4] (C5): several random spaced seeds of weight 12..28 over 100 graded mutants of one genome --
    per seed: sketch, all-vs-all, and the error of the ANI estimate against the true substitution rate."""
    pred = sks.frac_min_hash(1, 200)
    n5, L5 = 100, 5_000_000
    D5 = [C4_DS[g % 6] for g in range(n5)]
    b5 = ctx.synth(L5, [1000] * n5, [2000 + g for g in range(n5)], D5)
    rows5 = []
    for k in (12, 16, 20, 24, 28):
        w5 = k + 10
        m5 = sks.generate_random_spaced_seed_mask(w5, k)      # the reference's (k+10, k, seed 0) masks
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sets5 = ctx.sketch(b5, m5, w5, pred)
            cnt5, sizes5, ani5 = ctx.all_vs_all(sets5)
            e1.record(stream)
            torch.cuda.synchronize()
            ms5 = e0.elapsed_time(e1)
            for x in sets5:
                x.close()
        err = [abs(ani5[0, g] - (1.0 - 1.0 / D5[g])) for g in range(1, n5) if D5[g]]
        rows5.append({"weight": k, "window": w5, "ms_sketch_plus_all_pairs": ms5,
                      "mean_abs_ani_error_vs_true": float(np.mean(err)), "max_abs_ani_error": float(np.max(err)),
                      "mean_sketch_size": float(np.mean(sizes5))})
    b5.close()
    return {"c5_multi_seed": {"workload": "C5 = BASELINE configs[4]: %d synthetic 5 Mbp genomes (graded mutants of one base), random spaced "
                                          "seeds (k+10, k, seed 0) for k = 12..28, FMH(200); ANI(base, mutant) against 1 - 1/D" % n5,
                              "per_seed": rows5}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genomes", type=int, default=C4_N, help="genomes of the all-vs-all (default: the full configs[3])")
    ap.add_argument("--no-extra", action="store_true", help="skip the C2 / C3 / C5 legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c2-cpu-sample-bases", type=int, default=C2_L, help="genome length of the C2 cpu_baseline sample (0: skip)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
