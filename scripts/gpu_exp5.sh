#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
run() { # tag, args
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e $2 > gpurun_out/x5_$1.json 2> gpurun_out/x5_$1.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/x5_$1.json')); r=d['roofline']
    print('$1', 'ROIs/s=%.4g'%d['value'], 'fwd %.1f (%s %.1f%%)'%(r['fwd']['ms']*1e3, r['fwd']['algo'], 100*r['fwd']['frac']), 'bwd %.1f (%s %.1f%%)'%(r['bwd']['ms']*1e3, r['bwd']['algo'], 100*r['bwd']['frac']))
except Exception as ex:
    print('$1 failed', ex); print(open('gpurun_out/x5_$1.err').read()[-800:])
PY
}
run C2 "--workload C2"
run C3 "--workload C3"
run C4 "--workload C4"
run C1 "--workload C1"
run C5 "--workload C5"
run C5det "--workload C5 --mode deterministic"
run C3_bf16 "--workload C3 --top-dtype bfloat16"
( time python bench.py > gpurun_out/x5_default.json 2> gpurun_out/x5_default.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/x5_default.json')); print('default', d['value'], d['ms_per_step'], d['e2e'], d['e2e_torch']['value'], d['cpu_baseline'])"
