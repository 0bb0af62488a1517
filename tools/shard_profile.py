"""Phase profile of the batch build for an N-GPU shard (T/N trees) of the bench workload (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, rp_tree_b200 as R
W = bench.WORKLOAD
n, d = W["n"], W["d"]
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
for T in [int(a) for a in sys.argv[1:]] or [4, 8, 16]:
    hp = R.sampleHyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
    f = R.RPForest(0); f.setHyperplanes(hp, T, maxd); f.setPoints(X)
    if os.environ.get('RPF_PROJECT_VARIANT'):
        f.setOption('project_variant', int(os.environ['RPF_PROJECT_VARIANT']))
    for i in range(4):
        f.build(maxd, W["min_leaf"])
    ms = f.lastDeviceMs()
    f.setProfiling(True); f.build(maxd, W["min_leaf"])
    print("T=%d build %.3f ms |" % (T, ms), {k: (round(v[0], 3), v[1]) for k, v in f.profile().items() if v[1]})
    f.close()
