"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every profiled kernel in one or more
.ncu-rep files -> profiles/r02_traffic.json (argument --out NAME: another file), keyed by the engine's phase names (what bench.py's `roofline.traffic` reads).

  python tools/ncu_traffic.py gpurun_out/prof_v8.ncu-rep gpurun_out/prof_v7.ncu-rep
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PHASE = [("k_knn_f32", "q_knn"), ("k_knn_tma", "q_knn"), ("k_knn", "q_knn"), ("k_project_t<", "project"), ("k_project<", "project"), ("k_bottom", "bottom"), ("k_top_hist", "top_hist"),
         ("k_top_compact", "top_compact"), ("k_top_relabel", "top_relabel"), ("k_top_scatter", "top_relabel")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME_MS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main(paths):
    acc = {}
    for path in paths:
        out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        ir, iw, it, ik = (hdr.index(x) for x in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "Kernel Name"))
        ig = hdr.index("Grid Size") if "Grid Size" in hdr else None
        for r in rows[2:]:
            name = r[ik]
            ph = next((p for key, p in PHASE if key in name), None)
            if ph is None:
                continue
            if ph == "project" and float(r[ir].replace(",", "")) * UNIT[units[ir]] < 1e8:
                continue            # the query-projection launch of the same template
            b = float(r[ir].replace(",", "")) * UNIT[units[ir]] + float(r[iw].replace(",", "")) * UNIT[units[iw]]
            # top-phase / bottom kernels: grid.y = trees of the launch (the graph build runs them per concurrent branch)
            trees = None
            if ig is not None and (ph.startswith("top_") or ph == "bottom"):
                try:
                    trees = int(r[ig].strip("() ").split(",")[1])
                except Exception:
                    trees = None
            acc.setdefault(ph, []).append((b, float(r[it].replace(",", "")) * TIME_MS.get(units[it], 1.0), os.path.basename(path), name[:48], trees))
    res = {}
    for ph, lst in acc.items():
        res[ph] = dict(dram_bytes_per_launch=sum(x[0] for x in lst) / len(lst), launches_profiled=len(lst),
                       ncu_ms_per_launch=sum(x[1] for x in lst) / len(lst), source=sorted({x[2] for x in lst}), kernel=lst[0][3])
        if lst[0][4]:
            res[ph]["trees_per_launch"] = lst[0][4]      # bench.py scales the bytes to the trees of ITS launch
    dst = os.path.join(ROOT, "profiles", OUT)
    with open(dst, "w") as fh:
        json.dump(res, fh, indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True))


OUT = "r02_traffic.json"
if __name__ == "__main__":
    args = sys.argv[1:]
    if "--out" in args:
        i = args.index("--out"); OUT = args[i + 1]; del args[i:i + 2]
    main(args)
