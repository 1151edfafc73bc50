#!/usr/bin/env python3
"""Print the headline metrics of the first kernel in an ncu report: scripts/ncu_metrics.py <report> [kernel-regex]"""
import csv, subprocess, sys
rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"]
if len(sys.argv) > 2:
    cmd += ["-k", f"regex:{sys.argv[2]}"]
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True, check=True).stdout.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'sm__cycles_active.avg', 'smsp__thread_inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in keys or ('issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h):
        print(f"- {h} = {vals[i]} {units[i]}")
