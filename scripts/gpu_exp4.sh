#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests/test_plan.py tests/test_typed_top.py -m gpu -q -x 2>&1 | tail -8
run() { # tag, args
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e $2 > gpurun_out/x4_$1.json 2> gpurun_out/x4_$1.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/x4_$1.json')); r=d['roofline']
    print('$1', 'ROIs/s=%.4g'%d['value'], 'fwd %.1f (%s %.1f%%)'%(r['fwd']['ms']*1e3, r['fwd']['algo'], 100*r['fwd']['frac']), 'bwd %.1f (%s %.1f%%)'%(r['bwd']['ms']*1e3, r['bwd']['algo'], 100*r['bwd']['frac']), d['clocks'])
except Exception as ex:
    print('$1 failed', ex); print(open('gpurun_out/x4_$1.err').read()[-800:])
PY
}
run C3_f32 "--workload C3"
run C3_bf16 "--workload C3 --top-dtype bfloat16"
run C3_f16_rhwc "--workload C3 --top-dtype float16 --layout RHWC"
run C3_f32_rhwc "--workload C3 --layout RHWC"
run C4_f32_rhwc "--workload C4 --layout RHWC"
run C4_bf16_rhwc "--workload C4 --layout RHWC --top-dtype bfloat16"
run C4_f32 "--workload C4"
