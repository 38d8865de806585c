#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import collections
import csv
import sys


def main(path, per_step_div=None):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            d[r[ki]][0] += 1
            d[r[ki]][1] += float(r[vi].replace(",", ""))
        except (ValueError, IndexError):
            pass
    tot = sum(v[1] for v in d.values())
    print("total %.1f us over %d launches" % (tot / 1e3, sum(v[0] for v in d.values())))
    for k, v in sorted(d.items(), key=lambda x: -x[1][1]):
        print("%10.1f us %4d x %8.1f us %5.1f%%  %s" % (v[1] / 1e3, v[0], v[1] / 1e3 / v[0], 100 * v[1] / tot, k[:100]))


if __name__ == "__main__":
    main(sys.argv[1])
