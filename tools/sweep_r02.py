"""Round-2 tuning sweep on configs[1] data: build ms for tree counts {32, 4} under the options `branches`,
`project_variant` (7 = single-buffer kernel, 8 = pipelined kernel), `project_prefetch`.  Usage: python tools/sweep_r02.py"""
import itertools
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
rows = []
for T in (32, 4):
    hp = R.slice_hyperplanes(hp_all, maxd, 0, T)
    f = R.RPForest(0)
    f.setHyperplanes(hp, T, maxd)
    f.setPoints(X)
    ref = None
    for br, fu in itertools.product((2,), (0, 1, 2, 3)):
        pv, pf = 0, 1
        f.setOption("branches", br); f.setOption("fused_top", fu)
        ms = []
        for i in range(8):
            f.build(maxd, W["min_leaf"])
            if i >= 3:
                ms.append(f.lastDeviceMs())
        e = f.treeExport(T - 1)
        sig = (e["perm"][:1000].tobytes(), e["thr"][:100].tobytes())
        if ref is None:
            ref = sig
        assert sig == ref, "result changed with the options"
        f.setProfiling(True); f.build(maxd, W["min_leaf"]); prof = f.profile(); f.setProfiling(False)
        rows.append(dict(T=T, branches=br, fused_top=fu, build_ms=round(float(np.mean(ms)), 3),
                         project_ms=round(prof["project"][0], 3), top_ms=round(sum(v[0] for k, v in prof.items() if k.startswith("top_")), 3),
                         launches=sum(v[1] for v in prof.values())))
        print(json.dumps(rows[-1]), flush=True)
    f.close()
