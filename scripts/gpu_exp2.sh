#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python -m pytest tests/test_conv.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/x2_conv.log; tail -30 gpurun_out/x2_conv.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --workload C4"
$CMD > gpurun_out/x2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 80 --csv --log-file gpurun_out/x2_launches_C4.csv $CMD > gpurun_out/x2_ncu_l.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/x2_launches_C4.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=r[vi]
for (i,k),m in list(cur.items())[-12:]:
    print(i,k,' '.join('%s=%s'%(a.split('__')[-1][:18],b) for a,b in m.items()))
PY
