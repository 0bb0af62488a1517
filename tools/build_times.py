"""Per-step device and wall time of build / knn in the bench loop order (diagnostic)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, rp_tree_b200 as R
W = bench.WORKLOAD
n, d, T, k, nq = W["n"], W["d"], W["ntrees"], W["k"], W["nq"]
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
Xp = torch.empty((n, d), dtype=torch.float64, pin_memory=True); X = Xp.numpy(); X[:] = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
Qp = torch.empty((nq, d), dtype=torch.float64, pin_memory=True); Q = Qp.numpy(); Q[:] = bench.make_points(nq, d, W["query_seed"], W["clusters"], W["sigma"])
hp = R.sampleHyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
f = R.RPForest(0); f.setHyperplanes(hp, T, maxd); f.setPoints(X)
for i in range(8):
    t0 = time.perf_counter(); f.build(maxd, W["min_leaf"]); t1 = time.perf_counter(); b = f.lastDeviceMs()
    f.knnBatch(Q, k); t2 = time.perf_counter()
    print("step %d: build dev %.3f wall %.3f | knn dev %.3f wall %.3f" % (i, b, (t1 - t0) * 1e3, f.lastDeviceMs(), (t2 - t1) * 1e3))
for i in range(4):
    t0 = time.perf_counter(); f.build(maxd, W["min_leaf"]); t1 = time.perf_counter()
    print("build only %d: dev %.3f wall %.3f" % (i, f.lastDeviceMs(), (t1 - t0) * 1e3))
g = R.RPForest(0); g.setHyperplanes(hp, T, maxd)
bufs = {}
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); g.setPoints(X); t1 = time.perf_counter(); g.build(maxd, W["min_leaf"]); t2 = time.perf_counter()
    if not bufs:
        nn_ = len(g.topology()["child"])
        for key in ("thr", "mlo", "mhi"):
            bufs[key] = torch.empty((T, nn_), dtype=torch.float64, pin_memory=True).numpy()
        bufs["perm"] = torch.empty((T, n), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    t3 = time.perf_counter(); g.forestExport(bufs); t4 = time.perf_counter()
    print("e2e %d: setPoints %.2f build %.2f (dev %.2f) export %.2f" % (i, (t1 - t0) * 1e3, (t2 - t1) * 1e3, g.lastDeviceMs(), (t4 - t3) * 1e3))
