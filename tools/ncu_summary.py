#!/usr/bin/env python
"""Print the metrics we track from an `ncu --page raw --csv` export (one kernel per row)."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('==', d.get('Kernel Name', '?')[:60])
    for k in KEYS:
        if k in d:
            print(f"  {k:82s} {d[k]:>16s} {units[hdr.index(k)]}")
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v > 0.08:
                print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:.3f}")
