#!/bin/bash
# experiment: forward with L2 prefetch of the next plane; pull backward unroll 4 vs 8
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 900 python -m pytest tests -m gpu -q -x -k "conv or golden or full_size" 2>&1 | tail -4
run() { # tag, env, args
  env $2 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary $3 > gpurun_out/x1_$1.json 2> gpurun_out/x1_$1.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/x1_$1.json')); r=d['roofline']; sp=d['details']['rank0_step_spread']
    print('$1', 'ROIs/s=%.4g'%d['value'], 'fwd %.1f (%s) min %.1f'%(r['fwd']['ms']*1e3, r['fwd']['algo'], sp['fwd_ms_min']*1e3), 'bwd %.1f (%s) min %.1f max %.1f'%(r['bwd']['ms']*1e3, r['bwd']['algo'], sp['bwd_ms_min']*1e3, sp['bwd_ms_max']*1e3))
except Exception as e:
    print('$1 failed', e); print(open('gpurun_out/x1_$1.err').read()[-800:])
PY
}
run C2_u8 "MOBULA_PULL_UNROLL=8" "--workload C2"
run C2_u4 "MOBULA_PULL_UNROLL=4" "--workload C2"
run C4_u8 "MOBULA_PULL_UNROLL=8" "--workload C4"
run C4_u4 "MOBULA_PULL_UNROLL=4" "--workload C4"
run C3_u8 "MOBULA_PULL_UNROLL=8" "--workload C3"
run C3_u4 "MOBULA_PULL_UNROLL=4" "--workload C3"
run C4_u8_again "MOBULA_PULL_UNROLL=8" "--workload C4"
