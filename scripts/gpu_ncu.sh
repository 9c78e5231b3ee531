#!/bin/bash
# one ncu --set full capture of the kernels matching $1 (regex), on the C2 bench
mkdir -p gpurun_out
PAT=${1:-plane}
TAG=${2:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e ${3:-}"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$PAT -s ${4:-6} -c ${5:-4} -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
ls -la gpurun_out/prof_$TAG.ncu-rep
