"""Dev tool: the reference's own main() on this repository's headers (oracle/_ref/dropin_cli) over 4 x 2 Mbp FASTA
files: its 62 configurations with the reference's timing lines, slowest first."""
import os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import port   # FASTA writer only
exe = os.path.join(ROOT, "oracle", "_ref", "dropin_cli")
with tempfile.TemporaryDirectory() as d:
    base = port.gen(2_000_000, 5)
    files = []
    for g, D in enumerate((0, 200, 50, 20)):
        f = os.path.join(d, "g%d.fna" % g)
        port.write_fasta(f, base if D == 0 else port.mutate(base, 10 + g, D), "g%d" % g)
        files.append(f)
    t0 = time.perf_counter()
    out = subprocess.run([exe, os.path.join(d, "out.csv")] + files, capture_output=True, text=True, timeout=900)
    dt = time.perf_counter() - t0
    lines = [l for l in out.stdout.splitlines() if "Time taken" in l]
    vals = [(float(re.findall(r"[-+0-9.eE]+", l)[-1]), i, l) for i, l in enumerate(lines)]
    print("total %.2f s, %d timing lines, csv rows %d" % (dt, len(lines), sum(1 for _ in open(os.path.join(d, "out.csv")))))
    for v, i, l in sorted(vals, reverse=True)[:12]:
        print(i, l)
    print(out.stderr[-500:])
