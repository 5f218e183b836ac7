"""Key numbers of every kernel of an exported ncu report, and the report's SASS source page split per kernel.
usage: ncu_multi.py <base>   (reads <base>.raw.csv and <base>.source.csv written by tools/ncu_export.sh)"""
import csv, sys
base = sys.argv[1]
rows = list(csv.reader(open(base + '.raw.csv')))
h = rows[0]
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    d = dict(zip(h, r))
    print(d.get('Kernel Name'))
    for k in WANT:
        if k in d: print('   %-72s %s' % (k, d[k]))
    st = [(n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], float(v)) for n, v in d.items()
          if n.startswith('smsp__average_warps_issue_stalled') and n.endswith('per_issue_active.ratio')]
    print('   stalls per issue:', ', '.join(f'{k}={x:.2f}' for k, x in sorted(st, key=lambda kv: -kv[1])[:8]))
try:
    rows = list(csv.reader(open(base + '.source.csv')))
except FileNotFoundError:
    sys.exit(0)
hdrs = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
for n, i in enumerate(hdrs):
    j = hdrs[n + 1] if n + 1 < len(hdrs) else len(rows)
    end = j
    while end > i and len(rows[end - 1]) < 5: end -= 1
    with open(f'{base}_k{n}.source.csv', 'w', newline='') as f:
        csv.writer(f).writerows(rows[i:end])
