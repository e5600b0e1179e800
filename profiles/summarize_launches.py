#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list per kernel."""
import csv
import sys
from collections import defaultdict

lines = open(sys.argv[1]).read().splitlines()
start = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[start:]))
agg, cnt = defaultdict(lambda: defaultdict(float)), defaultdict(int)
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("sks::<unnamed>::", "")[-56:]
    agg[name][r["Metric Name"]] += float(r["Metric Value"].replace(",", ""))
    if r["Metric Name"] == "gpu__time_duration.sum":
        cnt[name] += 1
tot = sum(v["gpu__time_duration.sum"] for v in agg.values())
print("%-58s %5s %10s %10s %7s %10s %10s" % ("kernel", "n", "total_us", "avg_us", "share", "rd_MB/l", "wr_MB/l"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    t = v["gpu__time_duration.sum"]
    print("%-58s %5d %10.1f %10.1f %7.3f %10.1f %10.1f" % (k, cnt[k], t / 1e3, t / 1e3 / cnt[k], t / tot,
                                                        v["dram__bytes_read.sum"] / cnt[k] / 1e6,
                                                        v["dram__bytes_write.sum"] / cnt[k] / 1e6))
