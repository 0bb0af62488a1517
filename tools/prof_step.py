"""One warm pass + one measured pass of the hot path on the bench workload (for ncu captures)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rp_tree_b200 as R  # noqa: E402

W = bench.WORKLOAD
n, d, T, k, nq = W["n"], W["d"], W["ntrees"], W["k"], W["nq"]
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
Q = bench.make_points(nq, d, W["query_seed"], W["clusters"], W["sigma"])
hp = R.sampleHyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
f = R.RPForest(0)
if os.environ.get("RPF_PROJECT_VARIANT"):
    f.setOption("project_variant", int(os.environ["RPF_PROJECT_VARIANT"]))
if os.environ.get("RPF_LEAN_TOP"):
    f.setOption("lean_top", int(os.environ["RPF_LEAN_TOP"]))
if os.environ.get("RPF_TOP_CH"):       # "hist,compact,relabel" points per CTA
    for name, v in zip(("top_chunk_hist", "top_chunk_compact", "top_chunk_relabel"), os.environ["RPF_TOP_CH"].split(",")):
        f.setOption(name, int(v))
if os.environ.get("RPF_BOTTOM_CAP"):
    f.setBottomCap(int(os.environ["RPF_BOTTOM_CAP"]))
f.setHyperplanes(hp, T, maxd)
f.setPoints(X)
for i in range(passes):
    f.build(maxd, W["min_leaf"])
    b = f.lastDeviceMs()
    f.knnBatch(Q, k)
    print("pass %d: build %.3f ms, knn %.3f ms" % (i, b, f.lastDeviceMs()))
f.setProfiling(True)
f.build(maxd, W["min_leaf"])
print({k: v for k, v in f.profile().items() if v[1]})
