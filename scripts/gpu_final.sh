#!/bin/bash
# Round-end evidence on one GPU: parity suite, the bench lines of every workload, the ncu
# launch list and one --set full capture of the two dominant kernels (C2).
# Usage (under gpurun): bash scripts/gpu_final.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/bench_C2.err || tail -20 gpurun_out/bench_C2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_C2_reference.json 2> gpurun_out/bench_ref.err
python bench.py --mode deterministic --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_C2_deterministic.json 2>> gpurun_out/bench_C2.err
for wl in C1 C3 C4 C5; do
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/bench_$wl.err
done
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches_C2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'resident_(fwd2r?|bwd2)_kernel' -s 6 -c 2 -f -o gpurun_out/prof_${TAG}_final $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
for f in gpurun_out/${TAG}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'],
          'e2e=%s' % (d.get('e2e') and round(d['e2e']['value'])), 'frac=%s' % r.get('frac'),
          'fwd=%s' % (r.get('fwd') and round(r['fwd']['ms'] * 1e3, 1)), 'bwd=%s' % (r.get('bwd') and round(r['bwd']['ms'] * 1e3, 1)))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
