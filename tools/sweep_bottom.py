"""Bottom-phase kernels on configs[1] data: build / bottom ms with option `bottom_select` (1 = k_bottom4 warp-per-node select /
partition + k_bottom3 on flagged nodes, 0 = k_bottom3 everywhere), and a check that both give the same forest.
Usage: python tools/sweep_bottom.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rp_tree_b200 as R  # noqa: E402

W = bench.WORKLOAD
n, d = W["n"], W["d"]
X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
hp_all = R.sampleHyperplanes(W["forest_seed"], 32, maxd, W["pnz"], d)
for T in (32, 4):
    hp = R.slice_hyperplanes(hp_all, maxd, 0, T)
    f = R.RPForest(0)
    f.setHyperplanes(hp, T, maxd)
    f.setPoints(X)
    ref = None
    for bs in (0, 1):
        f.setOption("bottom_select", bs)
        ms = []
        for i in range(8):
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
        print(json.dumps(dict(T=T, bottom_select=bs, build_ms=round(float(np.mean(ms)), 3), bottom_ms=round(prof["bottom"][0], 3),
                              bottom_launches=prof["bottom"][1], same_forest=sig == ref)), flush=True)
    f.close()
