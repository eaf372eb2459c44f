"""Summarise an ncu report for profiles/: per kernel launch the duration, DRAM bytes, L2->L1 bytes, pipe utilisation."""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__cluster_dim_x']
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
units = rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('== %s  (id %s)' % (d.get('Kernel Name', '?')[:110], d.get('ID')))
    for k in want:
        if k in d:
            print('   %-70s %s %s' % (k, d[k], units[hdr.index(k)]))
