"""Per-source-line hot spots from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` output."""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Line No"][0]
    hdr = rows[hi]
    ci = {}
    for i, h in enumerate(hdr):
        ci.setdefault(h, i)
    data = [r for r in rows[hi + 1:] if len(r) > ci["Instructions Executed"] and r[2] == "-"]   # source-line rows
    tot = sum(float(r[ci["Instructions Executed"]] or 0) for r in data)
    tsamp = sum(float(r[ci["# Samples"]] or 0) for r in data)
    print("total warp instructions %.3e, samples %d" % (tot, tsamp))
    data.sort(key=lambda r: -float(r[ci["# Samples"]] or 0))
    for r in data[:top]:
        print("L%-5s inst %5.1f%%  samples %5.1f%% | %s" % (r[0], 100 * float(r[ci["Instructions Executed"]]) / tot,
                                                          100 * float(r[ci["# Samples"]]) / tsamp, r[1].strip()[:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
