"""Key numbers of an exported ncu report (tools/ncu_export.sh): usage ncu_summary.py <base>.raw.csv [...]"""
import csv, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'launch__grid_size']
for f in sys.argv[1:]:
    rows = list(csv.reader(open(f)))
    h, u, v = rows[0], rows[1], rows[2]
    print("==", f)
    for n, uu, vv in zip(h, u, v):
        if n in WANT:
            print(f"  {n:75s} {uu:16s} {vv}")
    st = [(n[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], float(vv)) for n, vv in zip(h, v)
          if n.startswith('smsp__average_warps_issue_stalled') and n.endswith('per_issue_active.ratio')]
    print("  stalls per issue:", ", ".join(f"{k}={x:.2f}" for k, x in sorted(st, key=lambda kv: -kv[1])[:8]))
