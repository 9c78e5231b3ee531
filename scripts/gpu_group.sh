#!/bin/bash
# grouped pull backward: parity (per-pixel lists vs grouped lists, oracle) per main-kernel variant, then A/B bench lines
mkdir -p gpurun_out
: > gpurun_out/group_ab.log
for v in ${VARIANTS:-0 1 2 3}; do
  MOBULA_PULL8_VARIANT=$v timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "grouped" 2>&1 | tail -3 | tee -a gpurun_out/group_pytest.log
done
MOBULA_PULL_GROUP=0 TAG="group=0" WORKLOADS="${WORKLOADS:-C2 C3 C4 C5}" bash scripts/gpu_ab.sh 2>&1 | tee -a gpurun_out/group_ab.log
for v in ${VARIANTS:-0 1 2 3}; do
  MOBULA_PULL8_VARIANT=$v MOBULA_PULL_GROUP=1 TAG="group=1 variant=$v" WORKLOADS="${WORKLOADS:-C2 C3 C4 C5}" bash scripts/gpu_ab.sh 2>&1 | tee -a gpurun_out/group_ab.log
done
