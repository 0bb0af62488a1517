"""Streaming (chunked) forest build on the bench workload: device / wall time per chunk size, phase profile (diagnostic)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, rp_tree_b200 as R
W = bench.WORKLOAD
n, d, T = W["n"], W["d"], W["ntrees"]
cfg = R.rpTreeCfg(W["min_leaf"], n, d)
maxd = cfg.fpMaxTreeDepth
Xp = torch.empty((n, d), dtype=torch.float64, pin_memory=True); X = Xp.numpy(); X[:] = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
hp = R.sampleHyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
f = R.RPForest(0); f.setHyperplanes(hp, T, maxd); f.setPoints(X)
chunks = [int(c) for c in sys.argv[1:]] or [cfg.fpDataChunkSize, 100000]
for chunk in chunks:
    for i in range(3):
        l0 = f.launchCount()
        t0 = time.perf_counter(); f.build(maxd, W["min_leaf"], chunk=chunk); t1 = time.perf_counter()
        print("chunk %d pass %d: device %.3f ms, wall %.3f ms, launches %d, lost %d, nodes %d" % (
            chunk, i, f.lastDeviceMs(), (t1 - t0) * 1e3, f.launchCount() - l0, f.pointsLost(), len(f.topology()["child"])))
    f.setProfiling(True)
    f.build(maxd, W["min_leaf"], chunk=chunk)
    print({k: (round(v[0], 3), v[1]) for k, v in f.profile().items() if v[1]})
    f.setProfiling(False)
for i in range(2):
    f.build(maxd, W["min_leaf"])
    print("batch build: device %.3f ms" % f.lastDeviceMs())
