"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...): total device time
and the share of each kernel.  Usage: python tools/launch_summary.py X.csv [top]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lines = [l for l in open(path, newline="") if not l.startswith("==")]
rows = list(csv.reader(lines))
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, agg = 0.0, defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    if len(r) <= mv:
        continue
    us = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
    name = r[kn]
    if "mb_" in name:                      # hb_microbench peaks, not part of the step
        continue
    tot += us
    agg[name][0] += us
    agg[name][1] += 1
print(f"total (excl. hb_microbench mb_* kernels) {tot:.1f} us over all captured launches")
for name, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{us:10.1f} us {100 * us / tot:5.1f}%  n={n:4d}  avg={us / n:8.2f} us  {name[:150]}")
