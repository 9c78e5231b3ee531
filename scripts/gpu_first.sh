#!/bin/bash
# first GPU visit: gather kernels only
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
python -c "import __graft_entry__ as g; g.build()" || exit 1
python -m pytest tests -m gpu -q -k "not plane and not deterministic and not full_size and not auto" -x 2>&1 | tail -15 > gpurun_out/pytest_gather.log
cat gpurun_out/pytest_gather.log
for wl in C2 C3 C5 C1; do
  python bench.py --steps 10 --warmup 3 --algo gather --workload $wl --no-cpu-baseline --no-e2e > gpurun_out/bench_gather_$wl.json 2> gpurun_out/bench_gather_$wl.err
  cat gpurun_out/bench_gather_$wl.json
done
python bench.py --steps 20 --warmup 3 --algo gather > gpurun_out/bench_gather_full.json 2> gpurun_out/bench_gather_full.err
cat gpurun_out/bench_gather_full.json
python bench.py --steps 2 --warmup 3 --algo gather --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -c 40 --csv --log-file gpurun_out/launches_gather.csv python bench.py --steps 2 --warmup 3 --algo gather --no-cpu-baseline --no-e2e > gpurun_out/ncu.log 2>&1
tail -5 gpurun_out/launches_gather.csv
