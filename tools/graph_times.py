"""Batch build device time with and without CUDA-graph replay, for the full forest and for an 8-GPU-sized shard (diagnostic)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, rp_tree_b200 as R
W = bench.WORKLOAD
n, d = W["n"], W["d"]
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
for T in (32, 4):
    hp = R.sampleHyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
    for graph in (1, 0):
        f = R.RPForest(0); f.setOption("cuda_graph", graph); f.setHyperplanes(hp, T, maxd); f.setPoints(X)
        ms = []
        for i in range(8):
            t0 = time.perf_counter(); f.build(maxd, W["min_leaf"]); ms.append((f.lastDeviceMs(), (time.perf_counter() - t0) * 1e3))
        e = f.treeExport(T - 1)
        print("T=%d graph=%d device ms %s | wall ms %s | launches %d | thr checksum %r" % (
            T, graph, [round(a, 3) for a, _ in ms], [round(b, 3) for _, b in ms], f.launchCount(), float(np.sum(e["thr"]))))
        f.close()
