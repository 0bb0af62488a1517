"""Candidate multiplicity on the bench workload: how many of a query's candidates (all trees) are distinct rows (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, rp_tree_b200 as R
W = bench.WORKLOAD
n, d, T, k = W["n"], W["d"], W["ntrees"], W["k"]
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
Q = bench.make_points(512, d, W["query_seed"], W["clusters"], W["sigma"])
f = R.forestBatch(W["forest_seed"], maxd, W["min_leaf"], T, W["pnz"], d, X)
off, ids = f.candidatesBatch(Q, -1)
tot = np.diff(off)
uniq = np.array([len(np.unique(ids[off[i]:off[i + 1]])) for i in range(len(Q))])
print("candidates/query mean %.1f, distinct rows mean %.1f (%.1f %%), min %.1f %%, max %.1f %%" % (
    tot.mean(), uniq.mean(), 100 * uniq.sum() / tot.sum(), 100 * (uniq / tot).min(), 100 * (uniq / tot).max()))
