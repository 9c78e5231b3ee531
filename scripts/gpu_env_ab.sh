#!/bin/bash
# A/B over values of one environment variable: gpu_env_ab.sh VAR v1 v2 ...   (WORKLOADS as in gpu_ab.sh)
VAR=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "unset" ]; then env -u $VAR TAG="$VAR unset" bash scripts/gpu_ab.sh; else env $VAR=$v TAG="$VAR=$v" bash scripts/gpu_ab.sh; fi
done | tee gpurun_out/env_ab_$VAR.log
