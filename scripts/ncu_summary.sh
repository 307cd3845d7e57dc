#!/bin/bash
# key counters of an ncu report: scripts/ncu_summary.sh <file.ncu-rep>
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
for v in vals:
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print('%-90s %s %s' % (h, v[i][:80], units[i]))
    print()
"
