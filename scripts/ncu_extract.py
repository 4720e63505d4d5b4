"""Key metrics of every kernel in an .ncu-rep (ncu --set full) as a markdown table."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("lts__t_bytes.sum", "L2 MB"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid")]
units = rows[1]
def val(r, k):
    if k not in col: return ""
    v = r[col[k]].replace(",", "")
    try: f = float(v)
    except ValueError: return v
    u = units[col[k]]
    if u in ("byte",): f /= 1e6
    if u in ("Kbyte",): f /= 1e3
    if u in ("Gbyte",): f *= 1e3
    if u in ("ns", "nsecond"): f /= 1e3
    if u in ("ms", "msecond"): f *= 1e3
    return f"{f:.2f}"
print("| kernel | " + " | ".join(w[1] for w in want) + " |")
print("|---|" + "---:|" * len(want))
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")[:48]
    print(f"| `{name}` | " + " | ".join(val(r, k) for k, _ in want) + " |")
