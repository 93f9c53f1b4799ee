#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py rep [pattern...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
pats = sys.argv[2:] or ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__throughput.avg.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit', 'sm__warps_active.avg.pct', 'smsp__issue_active.avg.pct', 'sm__inst_executed_pipe',
    'smsp__inst_executed.sum', 'sm__pipe_fma', 'sm__pipe_alu', 'sm__pipe_xu', 'sm__pipe_tensor', 'pipe_fmaheavy', 'pipe_fmalite',
    'issue_stalled', 'bank_conflicts', 'sm__cycles_elapsed.avg', 'lts__t_bytes.sum', 'l1tex__t_bytes', 'sm__throughput.avg.pct',
    'smsp__cycles_active.avg', 'shared_mem_per_block', 'sm__ctas_launched']
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, unit = rows[0], rows[1]
for r in rows[2:]:
    print('=== kernel:', r[hdr.index('Kernel Name')][:90])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats) and r[i] not in ('', 'n/a'):
            if 'issue_stalled' in h and not h.endswith('.pct'): continue
            print(f'  {h} [{unit[i]}] = {r[i]}')
