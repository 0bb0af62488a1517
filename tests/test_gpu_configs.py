"""Full-size checks on the other BASELINE.json configurations (the headline configs[1] lives in test_gpu_fullsize.py):

  configs[2]  GIST-like 1M x 960, 64 trees: column-blocked projection (d > one shared-memory tile), tree groups, the
              re-rank kernel at 7.7 KB rows
  configs[4]  Deep1B-like 10M x 96, k = 100: the per-GPU shard of the 8-GPU run (a block of trees on the full data),
              18 levels, n > 2^23, k = 100 lists

Per configuration: every tree is a permutation with the planned leaf sizes, thresholds / margins satisfy the
positional-median properties (Internal.hs:495-503) with keys recomputed by the oracle's innerSD, one whole tree equals
the oracle's build bit for bit, and knn ids / distance bits on a 2-tree forest equal the oracle's.
Sized so the file runs in a few minutes on one B200 (the oracle builds two full-size trees per configuration)."""
import numpy as np
import pytest

from helpers import compare_tree, bits
from test_gpu_fullsize import _check_node_properties

pytestmark = pytest.mark.gpu


def _setup(cfg_name, ntrees):
    import bench
    import rp_tree_b200 as R
    from oracle import orc
    W = bench.CONFIGS[cfg_name]
    n, d, minl = W["n"], W["d"], W["min_leaf"]
    X = bench.make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
    Q = bench.make_points(128, d, W["query_seed"], W["clusters"], W["sigma"])
    maxd = R.rpTreeCfg(minl, n, d).fpMaxTreeDepth
    hp_all = R.sampleHyperplanes(W["forest_seed"], W["ntrees"], maxd, W["pnz"], d)
    hp = R.slice_hyperplanes(hp_all, maxd, 0, ntrees)          # the first trees of the configuration's own draw
    f = R.RPForest(0)
    f.setHyperplanes(hp, ntrees, maxd)
    f.setPoints(X)
    f.build(maxd, minl)
    return dict(R=R, orc=orc, W=W, X=X, Q=Q, hp=hp, f=f, maxd=maxd, T=ntrees)


def _check_config(c, k, prop_trees, prop_levels):
    R, orc, f, X, Q, W, T, maxd = c["R"], c["orc"], c["f"], c["X"], c["Q"], c["W"], c["T"], c["maxd"]
    n, minl = W["n"], W["min_leaf"]
    plan = R.topologyPlan(n, maxd, minl)
    tp = f.topology()
    for key in ("child", "depth", "seg_start", "seg_size"):
        assert np.array_equal(tp[key], plan[key])
    assert f.leafOrderExact()
    for t in range(T):                                          # every tree holds every point exactly once
        e = f.treeExport(t)
        cnt = np.bincount(e["perm"], minlength=n)
        assert len(cnt) == n and cnt.min() == 1 and cnt.max() == 1, "tree %d: perm is not a permutation" % t
        if t in prop_trees:
            _check_node_properties(orc, X, c["hp"], maxd, t, e, levels=prop_levels)
    sizes = R.leafSizes(f)
    assert sizes.min() >= minl // 2 - 1 and sizes.max() <= minl
    # two whole trees against the oracle (built on two host threads), then knn on exactly that 2-tree forest
    hp2 = R.slice_hyperplanes(c["hp"], maxd, 0, 2)
    og = orc.Forest(X, hp2, 2, maxd, minl, threads=2)
    for t in range(2):
        bad = compare_tree(f.treeExport(t), og.export(t))
        assert not bad, (t, bad)
    g = R.RPForest(0)
    g.setHyperplanes(hp2, 2, maxd); g.setPoints(X); g.build(maxd, minl)
    d2, i2, c2 = g.knnBatch(Q[:48], k)
    for i in range(48):
        od, oi = og.knn(Q[i], k)
        assert np.array_equal(i2[i, :c2[i]], oi) and np.array_equal(bits(d2[i, :c2[i]]), bits(od)), "knn of query %d" % i
    rs = g.recallSumBatch(Q[:4], k) / 2
    for i in range(4):
        assert abs(rs[i] - og.recall_shared(Q[i], k)) <= 1e-12
    g.close()
    # the forest's own knn: sorted, k results, drawn from the candidate sets
    dist, ids, cnt = f.knnBatch(Q, k)
    assert np.all(cnt == k) and np.all(np.diff(dist, axis=1) >= 0)
    off, cand = f.candidatesBatch(Q[:8], -1)
    for i in range(8):
        assert np.all(np.isin(ids[i], cand[off[i]:off[i + 1]]))
    f.close()


def test_c3_gist_like_1m_x_960():
    """configs[2]: 8 of the 64 trees at full n and d (the forest of the bench is these trees' draw continued)."""
    c = _setup("c3", 8)
    assert (c["W"]["n"], c["W"]["d"], c["W"]["ntrees"]) == (1_000_000, 960, 64) and c["maxd"] == 14
    _check_config(c, k=10, prop_trees=(0, 7), prop_levels=(0, 5, 9, 10, 13))


def test_c5_deep1b_like_shard_10m_x_96_k100():
    """configs[4]: 4 trees of one GPU's block at full n = 10M (18 levels), lists of k = 100."""
    c = _setup("c5", 4)
    assert (c["W"]["n"], c["W"]["d"], c["W"]["ntrees"], c["W"]["k"]) == (10_000_000, 96, 256, 100) and c["maxd"] == 18
    _check_config(c, k=100, prop_trees=(0, 3), prop_levels=(0, 8, 13, 14, 17))
