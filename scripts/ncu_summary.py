#!/usr/bin/env python
"""Summarise an .ncu-rep (read here without a GPU): key metrics + stall reasons per kernel launch."""
import csv, subprocess, sys, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units = r[0], r[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'lts__t_bytes.sum', 'sm__inst_executed_pipe_lsu.sum']
seen = set()
for row in r[2:]:
    name = re.sub(r'\(.*', '', row[hdr.index('Kernel Name')])
    flt = [a for a in sys.argv[2:] if not a.startswith('--')]
    if flt and flt[0] not in name: continue
    if name in seen and '--all' not in sys.argv: continue
    seen.add(name)
    print('--- kernel', name[:70])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f'  {w:78s} {row[i]:>16s} {units[i]}')
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
            try:
                v = float(row[i])
                if v > 0.3: print(f'  STALL {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]:40s} {v:8.2f}')
            except ValueError: pass
