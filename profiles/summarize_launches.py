"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (launches, total us, share)."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,total_us,avg_us,share")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%s,%d,%.1f,%.1f,%.3f" % (k, a[0], a[1], a[1] / a[0], a[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
