"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python profiles/summarize_launches.py launches.csv [last_n_launches]
       python profiles/summarize_launches.py launches.csv start count     (a window of the list)"""
import collections
import csv
import re
import sys

path = sys.argv[1]
last = int(sys.argv[2]) if len(sys.argv) == 3 else None
window = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else None
import gzip
op = gzip.open if path.endswith(".gz") else open
with op(path, "rt") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
rows = []
for r in rd:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    rows.append((re.sub(r"\(.*", "", r[ki]), v))
print(f"{len(rows)} launches in file")
if last:
    rows = rows[-last:]
    print(f"summarising the last {last} launches (the timed region)")
if window:
    rows = rows[window[0]:window[0] + window[1]]
    print(f"summarising launches [{window[0]}, {window[0] + window[1]}) (the timed region)")
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in rows:
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in rows)
print(f"total kernel time {tot:.2f} ms")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:40]:40s} n={n:6d} {t:10.3f} ms {100 * t / tot:6.2f}%")
