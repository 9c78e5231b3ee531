#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the numbers DESIGN.md / bench.py quote.

    python scripts/summarize_ncu.py gpurun_out/prof_X.ncu-rep > profiles/X_summary.txt
"""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
    'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
    'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print('kernel:', d.get('Kernel Name', '?'))
        for k in KEYS:
            if k in d:
                print('  %-82s %s %s' % (k, d[k], u.get(k, '')))
        print()


if __name__ == '__main__':
    main()
