#!/usr/bin/env python
"""Key metrics + stall breakdown of one kernel from an .ncu-rep (ncu -i ... --page raw --csv): the numbers quoted in
DESIGN.md / profiles/*_ncu_summary.txt.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [row]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2 + row]
d = {n: (u[i], v[i]) for i, n in enumerate(h)}
print('kernel:', d.get('Kernel Name', ('', ''))[1])
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
for k in keys:
    if k in d:
        print('%-75s %-10s %s' % (k, d[k][0], d[k][1]))
st = {}
for n in d:
    if n.startswith('smsp__average_warps_issue_stalled_') and n.endswith('_per_issue_active.ratio'):
        st[n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]] = float(d[n][1])
tot = sum(st.values())
print('warp states per issue-active cycle (share): ' + ', '.join('%s %.1f%%' % (k, 100 * x / tot) for k, x in sorted(st.items(), key=lambda t: -t[1]) if x / tot > 0.004))
