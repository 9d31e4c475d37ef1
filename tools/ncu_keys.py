"""Print the headline metrics of every kernel in an .ncu-rep (ncu --page raw --csv).  usage: ncu_keys.py file.ncu-rep [extra substrings]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum"] + sys.argv[2:]
for r in rows[2:]:
    print("----")
    for h, u, v in zip(hdr, units, r):
        if any(h == k or (k in sys.argv[2:] and k in h) for k in KEYS):
            print(f"{h} [{u}] = {v}")
    st = [(float(v), h) for h, v in zip(hdr, r) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and v not in ("", "n/a")]
    for v, h in sorted(st, reverse=True)[:6]:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}: {v:.2f}")
