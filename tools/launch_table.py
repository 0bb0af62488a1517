"""Per-launch table (time, DRAM read/write) of the LAST forest build found in an ncu launch-list CSV."""
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, mi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("ID")
    by = {}
    for r in rows[1:]:
        by.setdefault(int(r[ii]), {"name": r[ki].split("(")[0]})[r[mi]] = float(r[vi].replace(",", ""))
    ids = sorted(by)
    last = max(i for i in ids if "k_project" in by[i]["name"] and ", 1>" in by[i]["name"])
    print("id,kernel,us,dram_read_MB,dram_write_MB")
    tot = 0.0
    for i in ids:
        if i < last:
            continue
        d = by[i]
        us = d["gpu__time_duration.sum"] / 1e3
        tot += us
        print("%d,%s,%.1f,%.1f,%.1f" % (i, d["name"][:40], us, d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6))
        if "k_bottom" in d["name"]:
            break
    print("total_us,%.1f" % tot)


if __name__ == "__main__":
    main(sys.argv[1])
