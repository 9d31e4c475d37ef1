"""Markdown table of the headline metrics of every launch in an .ncu-rep (`ncu -i rep --page raw --csv`).
usage: ncu_table.py file.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, k, scale_to=None):
    try:
        v = float(r[ix[k]])
    except Exception:
        return float("nan")
    u = units[ix[k]]
    if scale_to == "us":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    if scale_to == "MB":
        v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    return v


print("| kernel | grid x block | regs | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | DRAM % peak | FP64 pipe % | DMMA pipe % | warps active % | top stalls (per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    t = val(r, "gpu__time_duration.sum", "us")
    rd, wr = val(r, "dram__bytes_read.sum", "MB"), val(r, "dram__bytes_write.sum", "MB")
    st = [(float(v), h) for h, v in zip(hdr, r) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and v not in ("", "n/a")]
    top = ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.1f}" for v, h in sorted(st, reverse=True)[:3])
    print(f"| `{name}` | {int(val(r, 'launch__grid_size'))} x {int(val(r, 'launch__block_size'))} | {int(val(r, 'launch__registers_per_thread'))} | {t:.1f} | {rd:.1f} | {wr:.1f} | "
          f"{(rd + wr) / t * 1e3 if t > 0 else 0:.0f} | {val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{val(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f} | {val(r, 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {top} |")
