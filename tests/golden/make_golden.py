"""Regenerates tests/golden/oracle_small.json from the oracle (self-golden: a drift guard, not reference output;
the reference is Haskell and cannot be executed in this image)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_data  # noqa: E402
from oracle import orc  # noqa: E402

c = dict(n=600, d=6, T=2, maxd=5, minl=12, pnz=0.5, kind="gauss", data_seed=21, hp_seed=1235137, k=7)
X = make_data(c["n"], c["d"], c["data_seed"], c["kind"])
hp = orc.gen_hyperplanes(c["hp_seed"], c["T"], c["maxd"], c["pnz"], c["d"])
f = orc.Forest(X, hp, c["T"], c["maxd"], c["minl"])
trees = []
for t in range(c["T"]):
    e = f.export(t)
    internal = e["child"] >= 0
    trees.append(dict(thr_bits=e["thr"][internal].view(np.uint64).tolist(), mlo_bits=e["mlo"][internal].view(np.uint64).tolist(),
                      mhi_bits=e["mhi"][internal].view(np.uint64).tolist(), perm=e["perm"].tolist()))
q = (X[17] + 0.03).tolist()
d, i = f.knn(np.array(q), c["k"])
out = dict(config=c, hp_off=hp[0].tolist(), hp_idx=hp[1].tolist(), hp_val_bits=hp[2].view(np.uint64).tolist(), trees=trees,
           query=q, knn_ids=i.tolist(), knn_dist_bits=d.view(np.uint64).tolist(), recall=f.recall(np.array(q), c["k"]))
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.json"), "w") as fh:
    json.dump(out, fh)
print("wrote oracle_small.json")
