"""A/B of one engine option on configs[1] data: mean device build ms, per-phase profile, and a check that every setting gives
the same forest.  Usage: python tools/sweep_opt.py <option> <v0,v1,...> [trees,...]
       or: python tools/sweep_opt.py set "a=1,b=0;a=0,b=0;..." [trees,...]   (several options per setting)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rp_tree_b200 as R  # noqa: E402

opt = sys.argv[1]
if opt == "set":
    vals = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in st.split(",")) for st in sys.argv[2].split(";")]
else:
    vals = [{opt: int(v)} for v in sys.argv[2].split(",")]
trees = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [32, 4]
W = bench.WORKLOAD
n, d = W["n"], W["d"]
X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
hp_all = R.sampleHyperplanes(W["forest_seed"], 32, maxd, W["pnz"], d)
for T in trees:
    hp = R.slice_hyperplanes(hp_all, maxd, 0, T)
    f = R.RPForest(0)
    f.setHyperplanes(hp, T, maxd)
    f.setPoints(X)
    ref = None
    for v in vals:
        for ok, ov in v.items():
            f.setOption(ok, ov)
        ms = []
        for i in range(9):
            f.build(maxd, W["min_leaf"])
            if i >= 3:
                ms.append(f.lastDeviceMs())
        sig = []
        for t in (0, T - 1):
            e = f.treeExport(t)
            sig.append((e["perm"].tobytes(), e["thr"].tobytes(), e["mlo"].tobytes(), e["mhi"].tobytes()))
        if ref is None:
            ref = sig
        f.setProfiling(True); f.build(maxd, W["min_leaf"]); prof = f.profile(); f.setProfiling(False)
        print(json.dumps(dict(T=T, setting=v, build_ms=round(float(np.mean(ms)), 3), build_min=round(float(np.min(ms)), 3),
                              phases={k: round(p[0], 3) for k, p in prof.items() if p[1]}, same_forest=sig == ref)), flush=True)
    f.close()
