"""Per-source-line hot spots from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` output.
The file holds one section per (launch, source file); only the sections of the FIRST launch are read, rows of one source line summed."""
import collections
import csv
import os
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    agg = collections.OrderedDict()
    seen = set()
    i = 0
    while i < len(rows):
        r = rows[i]
        if len(r) == 2 and r[0] == "File Path":
            fpath = r[1]
            func = rows[i + 1][1] if i + 1 < len(rows) else ""
            if (fpath, func) in seen:
                break                                    # second launch of the kernel
            seen.add((fpath, func))
            hdr = rows[i + 2]
            ci = {}
            for k, h in enumerate(hdr):
                ci.setdefault(h, k)
            i += 3
            while i < len(rows) and not (len(rows[i]) == 2 and rows[i][0] == "File Path"):
                q = rows[i]
                if len(q) > ci["Instructions Executed"] and q[2] == "-":
                    a = agg.setdefault((os.path.basename(fpath), q[0]), [q[1].strip(), 0.0, 0.0])
                    a[1] += float(q[ci["Instructions Executed"]] or 0)
                    a[2] += float(q[ci["# Samples"]] or 0)
                i += 1
            continue
        i += 1
    tot = sum(a[1] for a in agg.values())
    tsamp = sum(a[2] for a in agg.values())
    print("total warp instructions %.3e, samples %d, source lines %d" % (tot, tsamp, len(agg)))
    key = (lambda kv: -kv[1][1]) if "--by-inst" in sys.argv else (lambda kv: -kv[1][2])
    for (f, ln), a in sorted(agg.items(), key=key)[:top]:
        print("%-18s L%-5s inst %5.1f%%  samples %5.1f%% | %s" % (f[:18], ln, 100 * a[1] / tot, 100 * a[2] / max(tsamp, 1), a[0][:110]))


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    main(args[0], int(args[1]) if len(args) > 1 else 30)
