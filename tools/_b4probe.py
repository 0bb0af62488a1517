import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
import rp_tree_b200 as R
W = bench.WORKLOAD
n, d = W["n"], W["d"]
X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
hp = R.sampleHyperplanes(W["forest_seed"], 32, maxd, W["pnz"], d)
f = R.RPForest(0); f.setHyperplanes(hp, 32, maxd); f.setPoints(X)
f.setOption("cuda_graph", 0)
for i in range(2): f.build(maxd, W["min_leaf"])
print("ok")
