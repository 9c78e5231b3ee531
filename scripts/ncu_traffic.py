#!/usr/bin/env python
"""profiles/traffic.json from an ncu launch list of bench.py.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --csv --log-file LIST.csv python bench.py --steps 2 --warmup 3 --workload W ...
    python scripts/ncu_traffic.py WORKLOAD_NAME LIST.csv [WORKLOAD_NAME LIST.csv ...] > profiles/traffic.json

One forward / backward CALL of the library is several kernels (tables, then the main kernel); the
DRAM traffic of a call is summed over its kernels -- the last complete call of the list -- and keyed
the way bench.py's roofline record names it ('forward (resident)', 'backward (pull)', ...).
"""
import csv
import json
import sys

FWD_MAIN = {'resident_fwd2r_kernel': 'resident', 'resident_fwd2_kernel': 'resident', 'resident_fwd2p_kernel': 'resident',
            'resident_fwd2pr_kernel': 'resident', 'resident_fwd_kernel': 'resident', 'sweep_fwd_kernel': 'sweep'}
BWD_MAIN = {'pull_bwd_kernel': 'pull', 'resident_bwd2_kernel': 'resident', 'tile_bwd_kernel': 'tile'}


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    out, cur = [], None
    for r in rows[1:]:
        if cur is None or cur['id'] != r[ii]:
            cur = {'id': r[ii], 'name': r[ki]}
            out.append(cur)
        cur[r[mi]] = float(r[vi].replace(',', ''))
    return out


def calls(ls):
    """Groups of our kernels, each ending in a main kernel: (direction, family, kernels)."""
    group = []
    for l in ls:
        if 'mobula_b200' not in l['name']:
            continue
        group.append(l)
        for table, direction in ((FWD_MAIN, 'forward'), (BWD_MAIN, 'backward')):
            for key, fam in table.items():
                if key + '<' in l['name'] or key + '(' in l['name']:
                    yield direction, fam, group
                    group = []
                    break
            if not group:
                break


def main():
    out = {'_comment': 'dram__bytes_read.sum + dram__bytes_write.sum of ONE forward / backward call of the library '
                       '(all of its kernels, the last complete call of the ncu launch list named in `source`); '
                       'written by scripts/ncu_traffic.py from the same gpurun visit as the bench lines',
           '_note': 'ncu serialises the kernels and starts each one on a cold L1 / warm-ish L2; writes still resident '
                    'in the 126 MB L2 when a kernel ends are not counted by dram__bytes_write'}
    args = sys.argv[1:]
    for wl, path in zip(args[0::2], args[1::2]):
        last = {}
        for direction, fam, group in calls(launches(path)):
            last[direction] = (fam, group)
        rec = {}
        for direction, (fam, group) in last.items():
            rd = sum(k.get('dram__bytes_read.sum', 0.0) for k in group)
            wr = sum(k.get('dram__bytes_write.sum', 0.0) for k in group)
            rec['%s (%s)' % (direction, fam)] = {
                'bytes': int(rd + wr), 'read': int(rd), 'write': int(wr), 'source': path,
                'kernels': [{'name': k['name'].split('(')[0].split('::')[-1][:48],
                             'us': round(k.get('gpu__time_duration.sum', 0.0) / 1e3, 1),
                             'dram_mb': round((k.get('dram__bytes_read.sum', 0.0) + k.get('dram__bytes_write.sum', 0.0)) / 1e6, 1)}
                            for k in group]}
        out[wl] = rec
    json.dump(out, sys.stdout, indent=1)
    sys.stdout.write('\n')


if __name__ == '__main__':
    main()
