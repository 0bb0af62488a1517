"""configs[4]-like probe (10M x 96, k = 100): one GPU's shard at a reduced tree count -- shape / partition checks, oracle
parity on one tree, phase profile (diagnostic)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, rp_tree_b200 as R
from oracle import orc
from helpers import compare_tree
n, d, T, nq, k, minl = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 96, 4, 10000, 100, 38
t0 = time.time()
X = bench.make_points(n, d, 1234, 256, 0.25)
Q = bench.make_points(nq, d, 4321, 256, 0.25)
print("data %.1f s" % (time.time() - t0), flush=True)
cfg = R.rpTreeCfg(minl, n, d)
maxd = cfg.fpMaxTreeDepth
hp = R.sampleHyperplanes(1235137, T, maxd, 0.1, d)
f = R.RPForest(0); f.setHyperplanes(hp, T, maxd); f.setPoints(X)
for i in range(3):
    f.build(maxd, minl); b = f.lastDeviceMs()
    f.knnBatch(Q, k); q = f.lastDeviceMs()
    print("pass %d: build %.3f ms (%.3g points/s for %d trees), knn %.3f ms" % (i, b, n / (b * 1e-3), T, q), flush=True)
f.setProfiling(True); f.build(maxd, minl)
print({k2: (round(v[0], 3), v[1]) for k2, v in f.profile().items() if v[1]})
f.setProfiling(False)
plan = R.topologyPlan(n, maxd, minl)
tp = f.topology()
assert all(np.array_equal(tp[x], plan[x]) for x in ("child", "depth", "seg_start", "seg_size"))
e = f.treeExport(T - 1)
cnt = np.bincount(e["perm"], minlength=n)
assert cnt.min() == 1 and cnt.max() == 1
print("leaf order exact:", f.leafOrderExact(), "| levels", maxd, "| nodes", len(tp["child"]), flush=True)
dist, ids, c = f.knnBatch(Q[:256], k)
off, _ = f.candidatesBatch(Q[:256], -1)
assert np.array_equal(c, np.minimum(k, np.diff(off))) and all(np.all(np.diff(dist[i, :c[i]]) >= 0) for i in range(256))
if "--oracle" in sys.argv:
    t0 = time.time()
    of = orc.Forest(X, R.slice_hyperplanes(hp, maxd, T - 1, 1), 1, maxd, minl)
    print("oracle tree %.1f s" % (time.time() - t0), flush=True)
    bad = compare_tree(e, of.export(0))
    print("parity with the oracle on tree %d:" % (T - 1), "OK" if not bad else bad)
    t0 = time.time()
    f.build(maxd, minl, chunk=cfg.fpDataChunkSize)
    print("streamed build (chunk %d): first call %.1f ms wall, lost %d" % (cfg.fpDataChunkSize, (time.time() - t0) * 1e3, f.pointsLost()), flush=True)
    f.build(maxd, minl, chunk=cfg.fpDataChunkSize); print("streamed build device %.3f ms" % f.lastDeviceMs())
    oc = orc.Forest(X, R.slice_hyperplanes(hp, maxd, T - 1, 1), 1, maxd, minl, chunk=cfg.fpDataChunkSize)
    bad = compare_tree(f.treeExport(T - 1), oc.export(0))
    print("streamed parity with the oracle on tree %d:" % (T - 1), "OK" if not bad else bad)
