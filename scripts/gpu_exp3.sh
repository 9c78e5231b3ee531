#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests/test_plan.py tests/test_typed_top.py tests/test_conv.py -m gpu -q -x 2>&1 | tail -15
python scripts/dispatch_overhead.py > gpurun_out/x3_dispatch.json 2> gpurun_out/x3_dispatch.err; python -c "
import json; d=json.load(open('gpurun_out/x3_dispatch.json')); print({k:v['host_us'] for k,v in d.items() if isinstance(v,dict) and 'host_us' in v})"
run() { # tag, args
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary $2 > gpurun_out/x3_$1.json 2> gpurun_out/x3_$1.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/x3_$1.json')); r=d['roofline']
    e=d.get('e2e') or {}; et=d.get('e2e_torch') or {}
    print('$1', 'ROIs/s=%.4g'%d['value'], 'fwd %.1f (%s %.1f%%)'%(r['fwd']['ms']*1e3, r['fwd']['algo'], 100*r['fwd']['frac']), 'bwd %.1f (%s %.1f%%)'%(r['bwd']['ms']*1e3, r['bwd']['algo'], 100*r['bwd']['frac']), 'e2e %.4g (%.1f ms)'%(e.get('value',0), e.get('ms_per_step',0)), 'e2e_torch %.4g host %.0f us'%(et.get('value',0), et.get('host_issue_us_per_step',0)))
except Exception as ex:
    print('$1 failed', ex); print(open('gpurun_out/x3_$1.err').read()[-800:])
PY
}
run C3_f32 "--workload C3"
run C3_bf16 "--workload C3 --top-dtype bfloat16"
run C3_f16_rhwc "--workload C3 --top-dtype float16 --layout RHWC"
run C3_f32_rhwc "--workload C3 --layout RHWC"
run C2_f16 "--workload C2 --top-dtype float16"
run C2_f32 "--workload C2"
