#!/usr/bin/env python
"""Headline metrics of EVERY launch in an .ncu-rep, one column per launch.  Usage: ncu_table.py file.ncu-rep"""
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers", "launch__shared_mem_config_size",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "smsp__inst_executed_op_branch.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")


def short(k):
    k2 = re.sub(r'\(.*', '', k).replace('void ', '').replace('gort::', '')
    m = re.search(r'<\(?(?:gort::PoolSrc\))?(\d)', k)
    return re.sub(r'<.*', '', k2) + (("<%s>" % m.group(1)) if (m and 'pool_trace' in k) else '')


print("%-78s %-10s %s" % ("launch", "", " | ".join("%-16s" % short(r[ki])[:16] for r in data)))
for h, u in zip(hdr, units):
    if h in KEYS or ("issue_stalled" in h and "per_issue_active" in h and "warps" in h and any(k in h for k in ("long_sc", "wait", "branch", "not_sel", "no_inst", "short", "math_pipe", "lg_thr"))):
        i = hdr.index(h)
        print("%-78s %-10s %s" % (h[:78], u[:10], " | ".join("%-16s" % r[i][:16] for r in data)))
