#!/bin/bash
# quick GPU visit: build, selected tests, bench lines per workload/algo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 1200 python -m pytest tests -m gpu -q -x ${PYTEST_K:+-k "$PYTEST_K"} 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for wl in ${WORKLOADS:-C2}; do
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
if [ -n "$NCU_LAUNCHES" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --workload ${NCU_WL:-C2} --algo ${NCU_ALGO:-auto}"
  $CMD > gpurun_out/plain_l.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max --clock-control none -c 60 --csv --log-file gpurun_out/launches_$NCU_LAUNCHES.csv $CMD > gpurun_out/ncu_l.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_$NCU_LAUNCHES.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii],r[ki][:70]),{})[r[mi]]=r[vi]
for (i,k),m in list(cur.items())[-14:]:
    print(i,k,' '.join('%s=%s'%(a.split('__')[-1][:22],b) for a,b in m.items()))
PY
fi
