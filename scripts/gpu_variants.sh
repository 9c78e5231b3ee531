#!/bin/bash
# builds the library with different -D variants on the GPU box and times the C2 bench
mkdir -p gpurun_out
for v in "$@"; do
  MOBULA_NVCC_EXTRA="$v" python -m mobulaop_b200.build --force > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  for wl in ${WORKLOADS:-C2}; do
  python bench.py --steps 10 --warmup 3 --workload $wl --algo ${ALGO:-resident} --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$v $wl', 'fwd %.1f us (%s)'%(r['fwd']['ms']*1e3, r['fwd']['algo']), 'bwd %.1f us (%s)'%(r['bwd']['ms']*1e3, r['bwd']['algo']))
"
  done
done
python -m mobulaop_b200.build --force > /dev/null 2>&1
