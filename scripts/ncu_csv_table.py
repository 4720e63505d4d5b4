"""Key metrics of every kernel in a `ncu --page raw --csv` export as a markdown table (see scripts/r02_profile.sh)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("lts__t_bytes.sum", "L2 MB"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor inst"), ("smsp__inst_executed.avg.per_cycle_active", "IPC/smsp"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid")]
def val(r, k):
    if k not in col: return ""
    v = r[col[k]].replace(",", "")
    try: f = float(v)
    except ValueError: return v
    u = units[col[k]]
    if u == "byte": f /= 1e6
    if u == "Kbyte": f /= 1e3
    if u == "Gbyte": f *= 1e3
    if u in ("ns", "nsecond"): f /= 1e3
    if u in ("ms", "msecond"): f *= 1e3
    return f"{f:.2f}"
print("| kernel | " + " | ".join(w[1] for w in want) + " | GB/s (dram) |")
print("|---|" + "---:|" * (len(want) + 1))
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("mmvae::<unnamed>::", "").replace("<unnamed>::", "")[:44]
    try:
        us = float(val(r, "gpu__time_duration.sum")); mb = float(val(r, "dram__bytes_read.sum")) + float(val(r, "dram__bytes_write.sum"))
        gbs = f"{mb / us * 1e3:.0f}"
    except ValueError:
        gbs = ""
    print(f"| `{name}` | " + " | ".join(val(r, k) for k, _ in want) + f" | {gbs} |")
