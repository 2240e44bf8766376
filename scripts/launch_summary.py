#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python scripts/launch_summary.py file.csv [n_proofs]"""
import collections
import csv
import sys


def main(path, div=1.0, detail=None):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, iv, ig = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        k = r[ik].split('(')[0].replace('zg::', '').replace('<unnamed>::', '').replace('void ', '')[:48]
        v = float(r[iv].replace(',', ''))
        agg[k][0] += 1
        agg[k][1] += v
        if detail and detail in k:
            print('   %-40s grid %-16s %9.1f us' % (k, r[ig], v / 1e3))
    tot = sum(v[1] for v in agg.values())
    print('total %.1f us over %d launches (divided by %g: %.1f us)' % (tot / 1e3, len(rows) - 1, div, tot / 1e3 / div))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-50s %6d %10.1f us %5.1f%%' % (k, v[0], v[1] / 1e3 / div, 100 * v[1] / tot))


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0, sys.argv[3] if len(sys.argv) > 3 else None)
