#!/bin/bash
# sweep kernels: parity first (hard time limits: a ring-protocol bug would hang), then bench lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout -s KILL ${PYTEST_LIMIT:-200} python -m pytest tests -m gpu -v -x --timeout 60 -k "${PYTEST_K:-sweep or deterministic_golden}" > gpurun_out/sw_pytest.log 2>&1
echo "pytest rc=$?"; tail -${PYTEST_TAIL:-30} gpurun_out/sw_pytest.log
for spec in ${BENCHES:-C2:sweep C2:auto C4:sweep C3:sweep C5:sweep}; do
  wl=${spec%%:*}; algo=${spec##*:}
  timeout -s KILL 300 python bench.py --steps ${STEPS:-20} --warmup 3 --workload $wl --algo $algo --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/sw_bench_${wl}_${algo}.json 2> gpurun_out/sw_bench_${wl}_${algo}.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/sw_bench_${wl}_${algo}.json'))
    r=d['roofline']
    print('$wl', '$algo', 'ROIs/s=%.4g'%d['value'], 'fwd %.1f us (%s, %.1f%%)'%(r['fwd']['ms']*1e3, r['fwd']['algo'], 100*r['fwd']['frac']), 'bwd %.1f us (%s, %.1f%%)'%(r['bwd']['ms']*1e3, r['bwd']['algo'], 100*r['bwd']['frac']), 'launches', d['gpu_launches'])
except Exception as e:
    print('$wl $algo failed', e); print(open('gpurun_out/sw_bench_${wl}_${algo}.err').read()[-1500:])
PY
done
