"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and, with
--seq, every launch of the LAST step in order (kernel, grid, us).  Usage:
    python scripts/summarize_launches.py gpurun_out/launches.csv [--last N] [--seq]"""
import csv, io, re, sys
from collections import OrderedDict

path = sys.argv[1]
last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else None
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(io.StringIO("".join(lines))):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*$", "", r["Kernel Name"])
    name = re.sub(r"mmvae::\(anonymous namespace\)::|mmvae::", "", name)
    rows.append((name, r.get("Grid Size", ""), r.get("Block Size", ""), us))
if last:
    rows = rows[-last:]
tot = sum(r[3] for r in rows)
agg = OrderedDict()
for n, g, b, us in rows:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us
print(f"Total {tot:.1f} us over {len(rows)} launches\n")
print("| kernel | launches | us | share |\n|---|---:|---:|---:|")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {us:.1f} | {100 * us / tot:.1f}% |")
if "--seq" in sys.argv:
    print("\nlaunch sequence:")
    for i, (n, g, b, us) in enumerate(rows):
        print(f"{i:4d} {us:9.1f} us  {n[:60]:60s} grid {g} block {b}")
