#!/usr/bin/env python
"""Condense `ncu -i <rep> --page raw --csv` into the columns the roofline discussion uses, one row per captured launch.
usage: ncu_summary.py <rep> [<rep> ...] > profiles/<name>.csv"""
import csv, io, subprocess, sys

COLS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"]
w = csv.writer(sys.stdout)
first = True
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u = rows[0], rows[1]
    cols = [c for c in COLS if c in h]
    if first:
        w.writerow(["capture", "kernel"] + cols)
        w.writerow(["ncu --set full --clock-control none", ""] + [u[h.index(c)] for c in cols])
        first = False
    for r in rows[2:]:
        d = dict(zip(h, r))
        w.writerow([rep.split("/")[-1], d["Kernel Name"][:70]] + [d[c] for c in cols])
