"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and, with
--seq, every launch of the LAST step in order (kernel, grid, us).  Usage:
    python scripts/summarize_launches.py gpurun_out/launches.csv [--last N | --step K] [--seq]"""
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
if "--step" in sys.argv:
    # one training step = the launches from a pack_weights_kernel up to the next one (bench.py's roofline block, which
    # follows the last step, starts with an L2-flush fill and is cut off); --step K picks the K-th step from the end
    k = int(sys.argv[sys.argv.index("--step") + 1])
    starts = [i for i, r in enumerate(rows) if "pack_weights_kernel" in r[0]]
    b = starts[-k]
    e = starts[-k + 1] if k > 1 else len(rows)
    seg = rows[b:e]
    for j, r in enumerate(seg):
        if "FillFunctor<unsigned char>" in r[0]:
            seg = seg[:j]
            break
    rows = seg
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
