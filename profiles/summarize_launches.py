"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel (shares, not absolutes)."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0][:48]
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0, row["Metric Unit"]])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("%-50s %5s %14s %8s" % ("kernel", "n", "total", "share"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-50s %5d %12.1f %s %7.3f" % (k, a[0], a[1], a[2], a[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
