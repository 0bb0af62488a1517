"""knn kernels on configs[1]: batch ms with option `knn_filter32` (1 = fp32 filter pass k_knn_f32 + exact re-rank of the survivors,
0 = exact TMA gather kernel k_knn_tma only) and a check that ids and distance bits agree.  Usage: python tools/sweep_knn.py [T ...]"""
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
for T in [int(a) for a in sys.argv[1:]] or [32, 4]:
    hp = R.slice_hyperplanes(hp_all, maxd, 0, T)
    f = R.RPForest(0)
    f.setHyperplanes(hp, T, maxd)
    f.setPoints(X)
    f.build(maxd, W["min_leaf"])
    ref = None
    for flt in (0, 1):
        f.setOption("knn_filter32", flt)
        ms = []
        for i in range(6):
            dist, ids, cnt = f.knnBatch(Q, k)
            if i >= 2:
                ms.append(f.lastDeviceMs())
        f.setProfiling(True); f.knnBatch(Q, k); prof = f.profile(); f.setProfiling(False)
        sig = (dist.tobytes(), ids.tobytes(), cnt.tobytes())
        if ref is None:
            ref = sig
        print(json.dumps(dict(T=T, knn_filter32=flt, knn_ms=round(float(np.mean(ms)), 3), q_knn_ms=round(prof["q_knn"][0], 3),
                              q_knn_launches=prof["q_knn"][1], same_answers=sig == ref)), flush=True)
    f.close()
