#!/bin/bash
# per-kernel times of the backward, per-pixel vs grouped lists (ncu launch list; cold-cache, serialised)
mkdir -p gpurun_out
for wl in ${WORKLOADS:-C2 C4}; do
for g in ${GROUPS_AB:-0 1}; do
  export MOBULA_PULL_GROUP=$g
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --workload $wl"
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,sm__inst_executed_pipe_lsu.sum --clock-control none --csv --log-file gpurun_out/group_launches_${wl}_g$g.csv $CMD > gpurun_out/ncu_l.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/group_launches_${wl}_g$g.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=r[vi]
print('== $wl group=$g')
for (i,k),m in list(cur.items())[-10:]:
    print(i,k,' '.join('%s=%s'%(a.split('__')[-1][:26],b) for a,b in m.items()))
PY
done
done
