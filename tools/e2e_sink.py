import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import bench, rp_tree_b200 as R
W = bench.WORKLOAD; n, d, T = W["n"], W["d"], W["ntrees"]
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
Xp = torch.empty((n, d), dtype=torch.float64, pin_memory=True); X = Xp.numpy(); X[:] = bench.make_points(n, d, 1234, 256, 0.25)
hp = R.sampleHyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
g = R.RPForest(0); g.setHyperplanes(hp, T, maxd)
g.buildFromHost(X, maxd, W["min_leaf"]); nn = len(g.topology()["child"])
bufs = {k: torch.empty((T, nn), dtype=torch.float64, pin_memory=True).numpy() for k in ("thr", "mlo", "mhi")}
bufs["perm"] = torch.empty((T, n), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
for sink in (False, True, False, True):
    g.setExportSink(bufs if sink else None)
    ts = []
    for i in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        g.buildFromHost(X, maxd, W["min_leaf"]); g.forestExport(bufs)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print("sink", sink, "e2e build+export ms:", [round(t, 2) for t in ts])
