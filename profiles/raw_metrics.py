"""Key metrics from `ncu -i X.ncu-rep --page raw --csv` (stdin).  usage: ncu -i rep --page raw --csv | python profiles/raw_metrics.py"""
import csv, sys
rows = list(csv.reader(sys.stdin)); h = rows[0]; u = dict(zip(h, rows[1]))
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    d = dict(zip(h, r))
    print(d.get('Kernel Name', '')[:70])
    for k in keys:
        if k in d: print("  %-62s %s %s" % (k, d[k], u[k]))
    for k in h:
        if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio'):
            try: v = float(d[k])
            except ValueError: continue
            if v >= 0.1: print("  stall %-56s %.2f" % (k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], v))
