#!/bin/bash
# round-2 first GPU visit: whole parity suite (new full-size C4/C5, sharded, CuPy-stub tests) and
# the new default bench line (C4 on one GPU) + the reference arm on the same config
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
free -g | head -2 > gpurun_out/p1_host.txt; nproc >> gpurun_out/p1_host.txt
( time timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 ) > gpurun_out/p1_pytest.log 2>&1
tail -40 gpurun_out/p1_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/p1_bench_C4.json 2> gpurun_out/p1_bench_C4.err || tail -30 gpurun_out/p1_bench_C4.err
tail -c 3000 gpurun_out/p1_bench_C4.json
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/p1_bench_C4_ref.json 2> gpurun_out/p1_bench_ref.err
tail -c 1200 gpurun_out/p1_bench_C4_ref.json; tail -5 gpurun_out/p1_bench_ref.err
