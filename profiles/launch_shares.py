#!/usr/bin/env python
"""Per-kernel share of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python profiles/launch_shares.py gpurun_out/launches.csv"""
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].split("(")[0][:44]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':46s} {'n':>5s} {'total us':>10s} {'avg us':>9s} {'share':>6s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:46s} {a[0]:5d} {a[1] / 1e3:10.1f} {a[1] / a[0] / 1e3:9.1f} {100 * a[1] / tot:5.1f}%")


if __name__ == "__main__":
    main()
