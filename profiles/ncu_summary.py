"""Print the handful of ncu metrics we track from a `--page raw --csv` export."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_shared_atom.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('=====', d['Kernel Name'][:70])
    for w in WANT:
        if w in d:
            print('  %-62s %s %s' % (w, d[w], units[hdr.index(w)]))
    st = sorted(((k, float(v or 0)) for k, v in d.items()
                 if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')),
                key=lambda kv: -kv[1])[:6]
    for k, v in st:
        print('  stall %-40s %.2f' % (k.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), v))
