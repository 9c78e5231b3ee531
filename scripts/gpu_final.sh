#!/bin/bash
# Round-end evidence on one GPU (everything lands in gpurun_out/, copied to profiles/ afterwards):
# parity suite, smoke, the default bench line + reference arm, one line per workload / mode / tensor
# format, the ncu launch lists (with DRAM bytes -> profiles/traffic.json), one --set full capture of
# the dominant kernels, the dispatch microbenchmark.
# Usage (under gpurun): bash scripts/gpu_final.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${TAG}_pytest_gpu.log
cat gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt
python bench.py > gpurun_out/${TAG}_bench_default_C4.json 2> gpurun_out/${TAG}_bench_default.err || tail -20 gpurun_out/${TAG}_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_default_reference.json 2> gpurun_out/${TAG}_bench_ref.err
one() { # name, args
  timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary $2 > gpurun_out/${TAG}_bench_$1.json 2> gpurun_out/${TAG}_bench_$1.err
}
one C2 "--workload C2"
one C2_deterministic "--workload C2 --mode deterministic --no-e2e"
one C3 "--workload C3"
one C3_deterministic "--workload C3 --mode deterministic --no-e2e"
one C4_deterministic "--workload C4 --mode deterministic --no-e2e"
one C1 "--workload C1 --no-e2e"
one C5 "--workload C5 --no-e2e"
one C5_deterministic "--workload C5 --mode deterministic --no-e2e"
one C3_bf16 "--workload C3 --top-dtype bfloat16 --no-e2e"
one C3_f16_host "--workload C3 --top-dtype float16"
one C3_f32_channel_last "--workload C3 --layout RHWC --no-e2e"
one C4_f32_channel_last "--workload C4 --layout RHWC --no-e2e"
python scripts/dispatch_overhead.py > gpurun_out/${TAG}_dispatch_overhead.json 2> gpurun_out/${TAG}_dispatch.err
# launch lists with DRAM bytes (C4 = the default bench command's workload, C2, C3)
for wl in C4 C2 C3 C5; do
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --workload $wl"
  $CMD > gpurun_out/${TAG}_plain_$wl.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum \
      --clock-control none -c 90 --csv --log-file gpurun_out/${TAG}_launches_$wl.csv $CMD > gpurun_out/${TAG}_ncu_l_$wl.log 2>&1
done
python scripts/ncu_traffic.py C4-batch-sharded gpurun_out/${TAG}_launches_C4.csv C2-box-head gpurun_out/${TAG}_launches_C2.csv \
    C3-mask-head gpurun_out/${TAG}_launches_C3.csv C5-stress-adaptive gpurun_out/${TAG}_launches_C5.csv > gpurun_out/${TAG}_traffic.json
if [ -n "$SKIP_FULL" ]; then echo "(full ncu captures skipped: run scripts/gpu_final_ncu.sh)"; else bash scripts/gpu_final_ncu.sh $TAG; fi
for f in gpurun_out/${TAG}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d.get('roofline') or {}
    print(sys.argv[1].split('/')[-1], 'value=%.4g' % d['value'], 'ms=%.3f' % d['ms_per_step'],
          'e2e=%s' % (d.get('e2e') and round(d['e2e']['value'])), 'frac=%.3f' % (r.get('frac') or 0),
          'fwd=%s' % (r.get('fwd') and round(r['fwd']['ms'] * 1e3, 1)), 'bwd=%s' % (r.get('bwd') and round(r['bwd']['ms'] * 1e3, 1)))
except Exception as e:
    print(sys.argv[1], 'unreadable', e)
PY
done
