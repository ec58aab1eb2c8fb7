"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.

    python scripts/launch_summary.py gpurun_out/launches_final.csv > profiles/r1_final_launches_summary.txt
"""
import csv
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("<unnamed>::", "")
        ns = float(r[-1].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print("# %s: %d launches, %.3f ms of kernel time (serialised, cold-cache under ncu)" % (sys.argv[1].split("/")[-1], len(rows), tot / 1e6))
    print("%-34s %8s %12s %10s %8s" % ("kernel", "launches", "total us", "avg us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-34s %8d %12.1f %10.2f %7.1f%%" % (k, n, t / 1e3, t / 1e3 / n, 100 * t / tot))


if __name__ == "__main__":
    main()
