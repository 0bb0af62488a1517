"""bottom_cap sweep on configs[1]: build ms and the top / bottom split.  Usage: python tools/sweep_cap.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rp_tree_b200 as R  # noqa: E402

W = bench.WORKLOAD
n, d, T = W["n"], W["d"], W["ntrees"]
X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
hp = R.sampleHyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
f = R.RPForest(0)
f.setHyperplanes(hp, T, maxd)
f.setPoints(X)
ref = None
for cap in (1024, 512, 256, 2048):
    for br in (2, 4):
        f.setBottomCap(cap); f.setOption("branches", br)
        ms = []
        for i in range(8):
            f.build(maxd, W["min_leaf"])
            if i >= 3:
                ms.append(f.lastDeviceMs())
        e = f.treeExport(T - 1)
        sig = (e["perm"].tobytes(), e["thr"].tobytes())
        ref = ref or sig
        assert sig == ref
        f.setProfiling(True); f.build(maxd, W["min_leaf"]); p = f.profile(); f.setProfiling(False)
        top = sum(v[0] for k, v in p.items() if k.startswith("top_"))
        print(json.dumps(dict(cap=cap, branches=br, build_ms=round(float(np.mean(ms)), 3), project=round(p["project"][0], 3),
                              top=round(top, 3), bottom=round(p["bottom"][0], 3))), flush=True)
