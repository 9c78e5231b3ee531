#!/bin/bash
# per-launch device times of one short bench run (cold-cache, serialised: compare shares)
mkdir -p gpurun_out
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e ${2:-}"
$CMD > gpurun_out/plain_l_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/launches_$TAG.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
from collections import OrderedDict
d=OrderedDict()
for r in rows[h+1:]:
    d.setdefault(r[0],{'name':r[4][:70]})[r[12]]=r[14]
for k,v in list(d.items())[-14:]:
    print(k, v['name'], v.get('gpu__time_duration.sum'), 'rd',v.get('dram__bytes_read.sum'),'wr',v.get('dram__bytes_write.sum'))
PY
