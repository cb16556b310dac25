"""Summarise an .ncu-rep (read here, no GPU): python profiles/ncu_summary.py <file.ncu-rep> [kernel-index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2 + idx]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct", "launch__shared_mem_per_block_dynamic",
        "smsp__average_warps_issue_stalled", "sm__inst_executed_pipe_fp64", "smsp__inst_executed_pipe_lsu",
        "sm__pipe_fp64_cycles_active", "smsp__cycles_active.avg", "l1tex__throughput.avg.pct"]
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w) for w in want):
        if h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) < 0.05:
            continue
        print(f"{h} [{u}] = {v}")
