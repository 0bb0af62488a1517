"""k_knn_f32 configuration sweep on configs[1] (ring stages, rows per stage, entry buffer, side region).  Usage: python tools/sweep_knn_cfg.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rp_tree_b200 as R  # noqa: E402

W = bench.WORKLOAD
n, d, nq, k = W["n"], W["d"], W["nq"], W["k"]
X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
Q = bench.make_points(nq, d, W["query_seed"], W["clusters"], W["sigma"])
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
hp_all = R.sampleHyperplanes(W["forest_seed"], 32, maxd, W["pnz"], d)
CFGS = [(4, 32, 1280, 512), (3, 30, 768, 256), (2, 32, 768, 256), (2, 28, 768, 256), (3, 20, 768, 256), (4, 16, 768, 256), (2, 16, 512, 128), (4, 24, 1024, 256)]
for T in (32, 4):
    hp = R.slice_hyperplanes(hp_all, maxd, 0, T)
    f = R.RPForest(0)
    f.setHyperplanes(hp, T, maxd)
    f.setPoints(X)
    f.build(maxd, W["min_leaf"])
    ref = None
    for st, rows, buf, sreg in CFGS:
        for name, v in (("knn_f32_stages", st), ("knn_f32_rows", rows), ("knn_f32_buf", buf), ("knn_f32_sreg", sreg)):
            f.setOption(name, v)
        f.setProfiling(True)
        ms = []
        for i in range(5):
            dist, ids, cnt = f.knnBatch(Q, k)
            if i >= 2:
                ms.append(f.profile()["q_knn"][0])
        f.setProfiling(False)
        sig = (dist.tobytes(), ids.tobytes(), cnt.tobytes())
        if ref is None:
            ref = sig
        print(json.dumps(dict(T=T, stages=st, rows=rows, buf=buf, sreg=sreg, q_knn_ms=round(float(np.mean(ms)), 3), same_answers=sig == ref)), flush=True)
    f.close()
