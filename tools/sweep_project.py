"""Projection kernel variants on configs[1] data: build / projection ms for tree counts {32, 4} under `project_variant`
(0 = default choice, 7 = k_project single buffer, 8 = k_project_pipe, 10 / 11 / 12 = k_project_t with 128-point tiles x 1024
threads, 64 x 512, 64 x 256) and a check that every variant yields the same forest.  Usage: python tools/sweep_project.py [variants...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rp_tree_b200 as R  # noqa: E402

variants = [int(a) for a in sys.argv[1:]] or [7, 8, 10, 11, 12, 0]
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
    for pv in variants:
        f.setOption("project_variant", pv)
        ms = []
        for i in range(8):
            f.build(maxd, W["min_leaf"])
            if i >= 3:
                ms.append(f.lastDeviceMs())
        e = f.treeExport(T - 1)
        sig = (e["perm"].tobytes(), e["thr"].tobytes(), e["mlo"].tobytes(), e["mhi"].tobytes())
        if ref is None:
            ref = sig
        same = sig == ref
        f.setProfiling(True); f.build(maxd, W["min_leaf"]); prof = f.profile(); f.setProfiling(False)
        print(json.dumps(dict(T=T, project_variant=pv, build_ms=round(float(np.mean(ms)), 3), project_ms=round(prof["project"][0], 3),
                              same_forest_as_first_variant=same)), flush=True)
    f.close()
