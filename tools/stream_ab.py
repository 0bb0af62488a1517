import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import bench, rp_tree_b200 as R
W = bench.WORKLOAD
n, d, T = W["n"], W["d"], W["ntrees"]
cfg = R.rpTreeCfg(W["min_leaf"], n, d)
maxd = cfg.fpMaxTreeDepth
X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
hp = R.sampleHyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
f = R.RPForest(0); f.setHyperplanes(hp, T, maxd); f.setPoints(X)
chunk = cfg.fpDataChunkSize
for setting in ({"fuse_relabel_hist": 1, "hist_big_chunk": 1}, {"fuse_relabel_hist": 0, "hist_big_chunk": 1}, {"fuse_relabel_hist": 0, "hist_big_chunk": 0}, {"fuse_relabel_hist": 1, "hist_big_chunk": 1}):
    for k, v in setting.items(): f.setOption(k, v)
    ms = []
    for i in range(4):
        f.build(maxd, W["min_leaf"], chunk=chunk); ms.append(round(f.lastDeviceMs(), 3))
    print(setting, ms, flush=True)
