"""The reference arm of bench.py (`--impl reference`) runs on the host alone: its JSON line must keep the contract the
driver reads (same metric, unit and config as the GPU arm; `impl`, `cpu_baseline`, an `e2e` block without copies), and
the sketches of its sample must be the oracle's."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_line_keeps_the_contract():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    bench = _bench_module()
    assert d["impl"] == "reference" and "unavailable" not in d
    assert d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"] == bench.bench_config(d["config"]["genomes"]) and d["config"]["genomes"] == 1000
    cpu = d["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["sample"] and cpu["value"] == d["value"]
    assert "nproc" in cpu and "OMP_NUM_THREADS" in cpu
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the sample is the first genomes of the workload: genome 0 is gen(5 Mbp, 1000) itself, sketched as the oracle does
    from oracle import port
    mask, w = port.seed_to_mask("0011111011010111111011001011101")
    want = port.sketch_set(port.gen(bench.C4_L, 1000), [bench.C4_L], mask, w, port.FMH, 1, 200, 181)
    assert d["result"]["sample_sizes"][0] == len(want)
    assert all(abs(s - len(want)) < 0.05 * len(want) for s in d["result"]["sample_sizes"])
