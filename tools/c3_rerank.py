"""configs[2] (1M x 960, 64 trees, 100k queries, k = 10) query step with and without the leaf-grouped tensor-core re-rank.
Usage: python tools/c3_rerank.py [ntrees] [nq]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rp_tree_b200 as R  # noqa: E402

W = bench.CONFIGS["c3"]
T = int(sys.argv[1]) if len(sys.argv) > 1 else W["ntrees"]
nq = int(sys.argv[2]) if len(sys.argv) > 2 else W["nq"]
n, d, k = W["n"], W["d"], W["k"]
X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
Q = bench.make_points(nq, d, W["query_seed"], W["clusters"], W["sigma"])
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
hp = R.slice_hyperplanes(R.sampleHyperplanes(W["forest_seed"], W["ntrees"], maxd, W["pnz"], d), maxd, 0, T)
f = R.RPForest(0)
f.setHyperplanes(hp, T, maxd); f.setPoints(X)
f.build(maxd, W["min_leaf"]); f.build(maxd, W["min_leaf"])
print(json.dumps(dict(build_ms=f.lastDeviceMs(), T=T, nq=nq)), flush=True)
res = {}
for mode in (2, 0):
    f.setOption("rerank_gemm", mode)
    ms = []
    for i in range(3):
        out = f.knnBatch(Q, k)
        ms.append(f.lastDeviceMs())
    res[mode] = out
    f.setProfiling(True); f.knnBatch(Q, k); prof = f.profile(); f.setProfiling(False)
    print(json.dumps(dict(rerank_gemm=mode, knn_ms=[round(x, 2) for x in ms], queries_per_s=nq / (min(ms) * 1e-3),
                          phases={a: round(b[0], 2) for a, b in prof.items() if b[1] > 0})), flush=True)
same = all(np.array_equal(a.view(np.uint64) if a.dtype == np.float64 else a, b.view(np.uint64) if b.dtype == np.float64 else b)
           for a, b in zip(res[2], res[0]))
print(json.dumps(dict(identical_results=bool(same))))
