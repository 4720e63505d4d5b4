"""Sum dram__bytes_read / write over the launches of ONE step of an ncu CSV launch list (see scripts/r02_step_dram.sh)."""
import csv, io, re, sys
from collections import OrderedDict
lines = [l for l in open(sys.argv[1], newline="") if not l.startswith("==")]
rows = OrderedDict()
for r in csv.DictReader(io.StringIO("".join(lines))):
    k = int(r["ID"])
    name = re.sub(r"\(.*$", "", r["Kernel Name"]); name = re.sub(r"mmvae::\(anonymous namespace\)::|mmvae::|<unnamed>::|void ", "", name)
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
    rows.setdefault(k, {"name": name})[r["Metric Name"]] = v * scale
seq = list(rows.values())
starts = [i for i, r in enumerate(seq) if "pack_weights_kernel" in r["name"]]
step = None
for k in range(len(starts) - 1, 0, -1):              # the last complete step without the optimizer (bench.py's resident / e2e graphs)
    cand = seq[starts[k - 1]:starts[k]]
    if not any("adam_kernel" in r["name"] or "prepare_input" in r["name"] for r in cand):
        step = cand
        break
assert step is not None
rd = sum(r.get("dram__bytes_read.sum", 0) for r in step); wr = sum(r.get("dram__bytes_write.sum", 0) for r in step)
print(f"one step: {len(step)} launches, dram read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB = {(rd + wr) / 1e6:.1f} MB "
      f"(algorithmic 2.79 MB/frame x 256 = 714 MB: x{(rd + wr) / 714.2e6:.2f})")
agg = OrderedDict()
for r in step:
    a = agg.setdefault(r["name"], [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += r.get("dram__bytes_read.sum", 0); a[2] += r.get("dram__bytes_write.sum", 0); a[3] += r.get("gpu__time_duration.sum", 0)
print("\n| kernel | launches | dram read MB | dram write MB | us |\n|---|---:|---:|---:|---:|")
for n, (c, a, w, t) in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
    print(f"| `{n[:60]}` | {c} | {a / 1e6:.1f} | {w / 1e6:.1f} | {t:.1f} |")
