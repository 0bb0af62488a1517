"""Where the end-to-end build time goes: H2D of the points, device build, D2H of the forest (per-tree vs whole-forest export)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rp_tree_b200 as R  # noqa: E402

W = bench.WORKLOAD
n, d, T, k, nq = W["n"], W["d"], W["ntrees"], W["k"], W["nq"]
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
Xp = torch.empty((n, d), dtype=torch.float64, pin_memory=True)
X = Xp.numpy()
X[:] = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
Qp = torch.empty((nq, d), dtype=torch.float64, pin_memory=True)
Q = Qp.numpy()
Q[:] = bench.make_points(nq, d, W["query_seed"], W["clusters"], W["sigma"])
hp = R.sampleHyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
f = R.RPForest(0)
f.setHyperplanes(hp, T, maxd)


def tm(fn, reps=4):
    fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), float(np.mean(ts))


print("setPoints (pinned H2D %.0f MB): min %.2f ms mean %.2f ms" % ((n * d * 8 / 1e6,) + tm(lambda: f.setPoints(X))))
print("build: min %.2f mean %.2f ms (device %.2f)" % (tm(lambda: f.build(maxd, W["min_leaf"])) + (f.lastDeviceMs(),)))
print("treeExport x%d (pageable): min %.2f mean %.2f ms" % ((T,) + tm(lambda: [f.treeExport(t) for t in range(T)])))
print("forestExport (pageable, fresh arrays): min %.2f mean %.2f ms" % tm(lambda: f.forestExport()))
nn = len(f.topology()["child"])
out = {key: torch.empty((T, nn), dtype=torch.float64, pin_memory=True).numpy() for key in ("thr", "mlo", "mhi")}
out["perm"] = torch.empty((T, n), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
print("forestExport (pinned, reused): min %.2f mean %.2f ms" % tm(lambda: f.forestExport(out)))
for opt in (1, 0):
    f.setOption("no_query_order", opt)
    r = tm(lambda: f.knnBatch(Q, k))
    print("knnBatch no_query_order=%d: min %.2f mean %.2f ms (device %.2f)" % ((opt,) + r + (f.lastDeviceMs(),)))
    f.setProfiling(True); f.knnBatch(Q, k); print({a: b for a, b in f.profile().items() if b[1]}); f.setProfiling(False)
