"""Summarise an .ncu-rep (raw metrics + per-opcode stall samples) into text; used to fill profiles/."""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
STALLS = ['stall_barrier', 'stall_long_sb', 'stall_math', 'stall_short_sb', 'stall_wait', 'stall_not_selected',
          'stall_selected', 'stall_mio', 'stall_branch_resolving', 'stall_no_inst', 'stall_dispatch', 'stall_lg',
          'stall_sleep', 'stall_membar', 'stall_tex', 'stall_drain', 'stall_misc']


def run(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = run(rep, 'raw')
    hdr = raw[0]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        print('== kernel:', d.get('Kernel Name', '?')[:110])
        for k in KEYS:
            if k in d:
                print('  %-85s %s %s' % (k, d[k], raw[1][hdr.index(k)]))
    src = run(rep, 'source')
    # the source page may hold several kernels back to back; split on header rows
    blocks, cur = [], None
    for r in src:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            blocks.append(cur)
        elif cur is not None:
            cur['rows'].append(r)
    for b in blocks:
        rows = b['rows']
        if len(rows) < 2:
            continue
        h = rows[0]
        ix = {x: i for i, x in enumerate(h)}
        data = [r for r in rows[1:] if len(r) == len(h)]
        tot = sum(int(r[ix['# Samples']] or 0) for r in data)
        print('== stall samples:', b['name'][:100], 'total', tot)
        allst = collections.Counter()
        agg = collections.Counter()
        per = collections.defaultdict(collections.Counter)
        for r in data:
            t = r[ix['Source']].split()
            op = (t[1] if t and t[0].startswith('@') and len(t) > 1 else (t[0] if t else '?')).split('.')[0]
            n = int(r[ix['# Samples']] or 0)
            agg[op] += n
            for k in STALLS:
                if k in ix:
                    v = int(r[ix[k]] or 0)
                    allst[k] += v
                    per[op][k] += v
        print('  by reason:', ', '.join('%s=%.1f%%' % (k[6:], 100.0 * v / max(tot, 1)) for k, v in allst.most_common(9)))
        for op, n in agg.most_common(10):
            print('  %-8s %5.1f%%  %s' % (op, 100.0 * n / max(tot, 1),
                                          ' '.join('%s=%d' % (k[6:], v) for k, v in per[op].most_common(3))))


if __name__ == '__main__':
    main()
