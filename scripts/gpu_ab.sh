#!/bin/bash
# A/B helper: bench lines of the backward for the workloads given (the library is whatever was built)
for wl in ${WORKLOADS:-C2 C3 C4}; do python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline --no-e2e --no-secondary 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); r=d['roofline']; print('$wl ${TAG}', 'fwd %.1f (%s)'%(r['fwd']['ms']*1e3, r['fwd']['algo']), 'bwd %.1f (%s)'%(r['bwd']['ms']*1e3, r['bwd']['algo']))"; done
