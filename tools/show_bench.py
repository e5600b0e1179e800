"""Dev tool: one-screen summary of bench.py JSON lines (files given on the command line)."""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    n = d["n_gpus"]
    print("== %s  N=%d  value %.4g %s  %.3f ms/step | e2e %.4g (%.3f ms, pageable %.2f ms) | sha ok %s rows ok %s" % (
        path, n, d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"],
        d["e2e"].get("pageable_host_buffers_ms_per_step", float("nan")), d["parity_check"].get("matches_recorded_sha256"),
        d["parity_check"]["rows_match_pairwise_kernels"]))
    print("   phases", {k: round(v, 3) for k, v in d["extra"]["c4_detail"]["ms_phases_rank0"].items()})
    print("   kernels", {k: (v["launches_per_step"], round(v["ms_per_step"], 3)) for k, v in d["roofline"]["kernels"].items()})
    print("   roofline", {k: d["roofline"].get(k) for k in ("kernel", "achieved", "frac", "issue_frac")})
    for name in ("c3_sketch", "c3_sketch_2gbp"):
        c3 = d["extra"].get(name)
        if c3:
            print("   %s: %.3g bases/s %.3f ms ; gathered %.3g %.3f ms; kernels %s" % (
                name, c3["bases_per_s"], c3["ms"], c3["bases_per_s_global_set_on_every_rank"], c3["ms_global_set_on_every_rank"],
                {k: round(v, 3) for k, v in c3["kernels_ms_per_step"].items()}))
    c2 = d["extra"].get("c2_pair")
    if c2:
        print("   c2: %.4g kmers/s %.4f ms/step e2e %.4f ms (pageable %.3f) frac %.3f" % (
            c2["value"], c2["ms_per_step"], c2["e2e"]["ms_per_step"], c2["e2e"]["pageable_host_buffers_ms_per_step"], c2["roofline"]["frac"]))
    if d.get("cpu_baseline"):
        print("   cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:120])
