#!/bin/bash
# launch list + one --set full capture of a kernel.  Usage: gpu_ncu.sh <tag> <kernel regex> <bench args...>
TAG=$1; KREGEX=$2; shift 2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary $*"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_l.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${TAG}_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
for r in rows[-14:]: print(r[ki][:60], r[vi])
PY
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s ${NCU_SKIP:-3} -c 1 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
