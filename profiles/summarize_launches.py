"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"][:64]
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0, row["Metric Unit"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-66s %5s %12s %7s %10s" % ("kernel", "n", "total", "share", "avg"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-66s %5d %12.0f %7.3f %10.0f %s" % (k, a[0], a[1], a[1] / tot, a[1] / a[0], a[2]))


if __name__ == "__main__":
    main(sys.argv[1])
