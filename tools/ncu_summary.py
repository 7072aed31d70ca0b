#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report -> markdown (profiles/<name>.md).  usage: ncu_summary.py rep.ncu-rep out.md"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max']


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        f.write('# ncu --set full summary: %s\n\n(captured with --clock-control none; one replayed launch per row, '
                'cold caches -- use for traffic / stall shares, not for bench numbers)\n\n' % rep)
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')]
            f.write('## `%s`  grid %s block %s\n\n| metric | value | unit |\n|---|---|---|\n' % (
                name[:90], r[hdr.index('Grid Size')], r[hdr.index('Block Size')]))
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write('| %s | %s | %s |\n' % (k, r[i], units[i]))
            f.write('\n')
    print(open(out).read())


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
