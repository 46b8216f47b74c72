#!/usr/bin/env python
"""Condense `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k "regex:gemm_bf16x3|ffn_chain" --csv`
(one fwd+bwd step of bench.py) into profiles/r02_gemm_traffic.json: measured DRAM bytes per GEMM launch (bench.py reports
it as roofline.traffic next to the algorithmic bytes).  usage: python tools/gemm_traffic.py launches.csv out.json"""
import collections
import csv
import json
import sys


def main(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    ix = {k: i for i, k in enumerate(h)}
    per = collections.defaultdict(dict)
    for r in rows[hi + 1:]:
        if len(r) != len(h):
            continue
        v, u = float(r[ix['Metric Value']].replace(',', '')), r[ix['Metric Unit']]
        mul = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 'usecond': 1e-6,
               'nsecond': 1e-9, 'msecond': 1e-3}.get(u, 1)
        per[r[ix['ID']]][r[ix['Metric Name']]] = v * mul
    n = len(per)
    rd = sum(d.get('dram__bytes_read.sum', 0) for d in per.values())
    wr = sum(d.get('dram__bytes_write.sum', 0) for d in per.values())
    t = sum(d.get('gpu__time_duration.sum', 0) for d in per.values())
    json.dump({'launches': n, 'dram_read_bytes': rd, 'dram_write_bytes': wr, 'bytes_per_launch': (rd + wr) / max(n, 1),
               'time_s_under_ncu': t,
               'note': f'ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the {n} tcgen05 (gemm_bf16x3 + ffn_chain) launches captured '
                       f'(first fwd+bwd micro-batch of bench.py, 8 x 1024^2), divided by the launch count'}, open(out, 'w'), indent=1)
    print(open(out).read())


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
