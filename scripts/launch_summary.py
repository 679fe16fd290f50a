"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): python scripts/launch_summary.py FILE..."""
import collections
import csv
import sys


def summary(path, top=16):
    hdr = None
    agg = collections.OrderedDict()
    for r in csv.reader(open(path, errors="replace")):
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr is None or not r or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3}.get(d["Metric Unit"], 1e-6)
        a = agg.setdefault(d["Kernel Name"].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = ["%s: %d launches, %.2f ms" % (path, sum(a[0] for a in agg.values()), tot)]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        out.append("  %-28s n=%3d %9.2f ms %5.1f%%" % (k, a[0], a[1], 100 * a[1] / tot))
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(summary(p))
