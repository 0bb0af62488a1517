"""configs[2]-like probe (d = 960): build + knn phase profile at a reduced point count (diagnostic)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, rp_tree_b200 as R
n, d, T, nq, k, minl = int(sys.argv[1]) if len(sys.argv) > 1 else 200000, int(os.environ.get('RPF_D', '960')), 16, 2000, 10, 64
X = bench.make_points(n, d, 1234, 256, 0.25)
Q = bench.make_points(nq, d, 4321, 256, 0.25)
maxd = R.rpTreeCfg(minl, n, d).fpMaxTreeDepth
hp = R.sampleHyperplanes(1235137, T, maxd, 0.1, d)
f = R.RPForest(0); f.setHyperplanes(hp, T, maxd); f.setPoints(X)
if os.environ.get('RPF_PROJECT_VARIANT'):
    f.setOption('project_variant', int(os.environ['RPF_PROJECT_VARIANT']))
for i in range(3):
    f.build(maxd, minl); b = f.lastDeviceMs()
    f.knnBatch(Q, k); q = f.lastDeviceMs()
    print("pass %d: build %.3f ms, knn %.3f ms" % (i, b, q))
f.setProfiling(True); f.build(maxd, minl)
print({k2: (round(v[0], 3), v[1]) for k2, v in f.profile().items() if v[1]})
f.knnBatch(Q, k)
print({k2: (round(v[0], 3), v[1]) for k2, v in f.profile().items() if v[1]})
