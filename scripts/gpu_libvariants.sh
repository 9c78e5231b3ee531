#!/bin/bash
# times prebuilt library variants (mobulaop_b200/lib/variants/*.so, linked in the container) against the shipped one
mkdir -p gpurun_out
L=mobulaop_b200/lib/libmobula_roialign_sm100.so
cp $L /tmp/base.so
TAG="base" bash scripts/gpu_ab.sh | tee gpurun_out/libvariants.log
for v in mobulaop_b200/lib/variants/*.so; do
  cp $v $L
  if [ -n "$PYTEST_K" ]; then timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$PYTEST_K" 2>&1 | tail -2 | tee -a gpurun_out/libvariants.log; fi
  TAG="$(basename $v)" bash scripts/gpu_ab.sh | tee -a gpurun_out/libvariants.log
done
cp /tmp/base.so $L
