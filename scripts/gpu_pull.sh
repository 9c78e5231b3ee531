#!/bin/bash
# pull backward: parity first, then bench lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout -s KILL ${PYTEST_LIMIT:-400} python -m pytest tests -m gpu -q -x --timeout 120 -k "${PYTEST_K:-pull or golden}" > gpurun_out/pl_pytest.log 2>&1
echo "pytest rc=$?"; tail -${PYTEST_TAIL:-30} gpurun_out/pl_pytest.log
for spec in ${BENCHES-C2:pull C2:resident C3:pull C4:pull C1:pull}; do
  wl=${spec%%:*}; algo=${spec##*:}
  for det in ${DETS:-0 1}; do
  extra=""; [ "$det" = 1 ] && extra="--mode deterministic"
  timeout -s KILL 300 python bench.py --steps ${STEPS:-20} --warmup 3 --workload $wl --algo $algo $extra --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/pl_bench_${wl}_${algo}_$det.json 2> gpurun_out/pl_bench_${wl}_${algo}_$det.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/pl_bench_${wl}_${algo}_$det.json'))
    r=d['roofline']
    print('$wl', '$algo', 'det=$det', 'ROIs/s=%.4g'%d['value'], 'fwd %.1f us (%s, %.1f%%)'%(r['fwd']['ms']*1e3, r['fwd']['algo'], 100*r['fwd']['frac']), 'bwd %.1f us (%s, %.1f%%)'%(r['bwd']['ms']*1e3, r['bwd']['algo'], 100*r['bwd']['frac']), 'launches', d['gpu_launches'])
except Exception as e:
    print('$wl $algo failed', e); print(open('gpurun_out/pl_bench_${wl}_${algo}_$det.err').read()[-1500:])
PY
  done
done
if [ -n "$NCU_WL" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --workload $NCU_WL --algo ${NCU_ALGO:-pull}"
  $CMD > gpurun_out/pl_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 80 --csv --log-file gpurun_out/pl_launches_${NCU_WL}.csv $CMD > gpurun_out/pl_ncu_l.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/pl_launches_${NCU_WL}.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=r[vi]
for (i,k),m in list(cur.items())[-16:]:
    print(i,k,' '.join('%s=%s'%(a.split('__')[-1][:18],b) for a,b in m.items()))
PY
fi
