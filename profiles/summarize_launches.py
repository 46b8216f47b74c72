#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised:
compare SHARES, not absolutes).  usage: python profiles/summarize_launches.py launches.csv [passes]"""
import collections
import csv
import sys


def main(path, passes=1):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    ix = {k: i for i, k in enumerate(h)}
    data = [r for r in rows[hi + 1:] if len(r) == len(h)]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in data:
        name = r[ix['Kernel Name']].split('(')[0][:70]
        v, u = float(r[ix['Metric Value']]), r[ix['Metric Unit']]
        v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f'{len(data)} launches, {T / 1e3:.2f} ms total over {passes} pass(es) of the step -> {T / 1e3 / passes:.2f} ms per pass')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:30]:
        print(f'{v / 1e3 / passes:9.3f} ms/pass {100 * v / T:5.1f}%  n/pass={cnt[k] / passes:7.1f}  {k}')


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1)
