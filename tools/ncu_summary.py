#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the few metrics this project tracks.  usage: tools_ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum',
        'smsp__sass_inst_executed_op_global_ld.sum', 'smsp__sass_inst_executed_op_global_st.sum']
out = []
for r in rows[2:]:
    out.append('---- ' + r[idx['Kernel Name']][:70])
    for w in want:
        if w in idx:
            out.append(f"  {w} [{units[idx[w]]}] = {r[idx[w]]}")
    st = []
    for h in hdr:
        if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
            try:
                v = float(r[idx[h]])
            except ValueError:
                continue
            if v > 0.25:
                st.append((v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    out.append('  stalls (warps per issue-active cycle): ' + ', '.join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)))
txt = '\n'.join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(txt + '\n')
