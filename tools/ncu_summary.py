"""Compact per-kernel summary of an .ncu-rep (the columns profiles/*_summary.csv carry):
   python tools/ncu_summary.py report.ncu-rep > profiles/rNN_ncu_full_<what>_summary.csv"""
import csv
import subprocess
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "sm__cycles_elapsed.max", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = [hdr.index(c) for c in COLS if c in hdr]
w = csv.writer(sys.stdout)
w.writerow([units[i] for i in idx])
w.writerow([hdr[i] for i in idx])
for r in rows[2:]:
    w.writerow([r[i][:60] if hdr[i] == "Kernel Name" else r[i] for i in idx])
