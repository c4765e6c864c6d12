"""Flatten `.ncu-rep` captures into the per-launch CSV kept under profiles/ (the reports themselves are 10-20 MB each
and stay out of git).  usage: python profiles/ncu_summarize.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/x.csv"""
import csv
import io
import os
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed.sum"]

out = csv.writer(sys.stdout)
out.writerow(["report", "Kernel Name"] + COLS)
wrote_units = False
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    if not wrote_units:                                  # ncu picks a unit per column and report; keep it next to the data
        out.writerow([os.path.basename(rep), "(units)"] + [units[idx[c]] if c in idx else "" for c in COLS])
        wrote_units = True
    for r in body:
        out.writerow([os.path.basename(rep), r[idx["Kernel Name"]]] + [r[idx[c]] if c in idx else "" for c in COLS])
