#!/bin/bash
# One `ncu --set full` capture of the dominant kernels per workload.  The reports are summarised on
# the box (scripts/summarize_ncu.py -> gpurun_out/<tag>_<wl>_ncu_summary.txt); only the reports of the
# default workload travel back whole (gpurun returns at most 64 MiB).
# Usage (under gpurun, after the same commands ran clean without ncu): bash scripts/gpu_final_ncu.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out /tmp/ncu
python -c "import __graft_entry__ as g; g.build()" || exit 1
cap() { # workload, kernel regex, skip, count, keep-report
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --workload $1"
  $CMD > gpurun_out/${TAG}_plain_full_$1.log 2>&1 || { echo "$1: plain run failed"; return; }
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/ncu/prof_${TAG}_$1 $CMD > gpurun_out/${TAG}_ncu_full_$1.log 2>&1
  python scripts/summarize_ncu.py /tmp/ncu/prof_${TAG}_$1.ncu-rep > gpurun_out/${TAG}_$1_ncu_summary.txt
  ncu -i /tmp/ncu/prof_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_ncu_raw.csv 2>/dev/null
  [ -n "$5" ] && cp /tmp/ncu/prof_${TAG}_$1.ncu-rep gpurun_out/
  grep -c "^kernel" gpurun_out/${TAG}_$1_ncu_summary.txt
}
cap C4 'resident_fwd2r_kernel|pull_bwd_kernel|pull_tables_kernel' 9 3 keep
cap C2 'resident_fwd2r_kernel|pull_bwd_kernel|pull_tables_kernel' 9 3
cap C3 'resident_fwd2p_kernel|pull_bwd_kernel|pull_tables_kernel' 9 3
cap C5 'sweep_fwd_kernel|pull_bwd_kernel' 6 2
du -sh gpurun_out
