"""Dev tool: host vs device FASTA ingest of n synthetic 5 Mbp genomes."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import spaced_kmer_sketching_b200 as sks
from oracle import port

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d = tempfile.mkdtemp()
paths = []
for g in range(n):
    p = os.path.join(d, "g%d.fna" % g)
    port.write_fasta(p, port.gen(5_000_000, 100 + g))
    paths.append(p)
ctx = sks.Context(0)
for rep in range(3):
    t0 = time.perf_counter()
    genomes = [sks.fasta_parse_file(p) for p in paths]
    t1 = time.perf_counter()
    b1 = ctx.upload(genomes)
    ctx.sync()
    t2 = time.perf_counter()
    ctx.profile(True); ctx.kernel_stats()
    b2 = ctx.batch_from_fasta_files(paths)
    ctx.sync()
    t3 = time.perf_counter()
    ks = ctx.kernel_stats()
    texts = [open(p, "rb").read() for p in paths]
    t4 = time.perf_counter()
    b3 = ctx.batch_from_fasta_text(texts)
    ctx.sync()
    t5 = time.perf_counter()
    print("n=%d host parse %.1f ms (+upload %.1f) | device from files %.1f ms, from text in RAM %.1f ms; fasta kernels %s"
          % (n, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t5 - t4) * 1e3, {k: round(v[1], 3) for k, v in ks.items()}))
