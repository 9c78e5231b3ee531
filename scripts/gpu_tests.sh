#!/bin/bash
# parity suite + one bench line per algorithm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for wl in ${WORKLOADS:-C2 C3 C5 C1}; do
 for algo in ${ALGOLIST:-auto}; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $wl --algo $algo --no-cpu-baseline --no-e2e > gpurun_out/bench_${algo}_$wl.json 2> gpurun_out/bench_${algo}_$wl.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${algo}_$wl.json'))
    r=d['roofline']
    print('$wl', '$algo', 'ROIs/s=%.3g'%d['value'], 'fwd %.1f us (%s, %.1f%%)'%(r['fwd']['ms']*1e3, r['fwd']['algo'], 100*r['fwd']['frac']), 'bwd %.1f us (%s, %.1f%%)'%(r['bwd']['ms']*1e3, r['bwd']['algo'], 100*r['bwd']['frac']), 'launches', d['gpu_launches'])
except Exception as e:
    print('$wl $algo failed', e); print(open('gpurun_out/bench_${algo}_$wl.err').read()[-1500:])
PY
 done
done
