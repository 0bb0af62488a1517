"""Full-size checks on BASELINE.json configs[1] (1M x 128 fp64, 32 trees, pnz 0.1, minLeaf 64, maxDepth 14):
size-independent properties for every tree, plus direct parity with the oracle on a sample of trees (the oracle
builds one 1M-point tree in about two seconds).

Properties checked per tree (they pin the result without a second implementation of the build):
  * perm is a permutation of 0..n-1 and the leaf sizes are the positional-median sizes (Internal.hs:495,503)
  * at every node of every level: thr == min key of the right part, max key of the left part <= thr,
    mlo == max key of the left part, mhi == second smallest key of the right part (Internal.hs:496-503),
    with keys recomputed by the oracle's innerSD restatement -- bit exact
"""
import numpy as np
import pytest

from helpers import compare_tree, bits

pytestmark = pytest.mark.gpu

N, D, T, MINL, PNZ, NQ, K = 1_000_000, 128, 32, 64, 0.1, 256, 10


@pytest.fixture(scope="module")
def c2(built):
    import bench
    import rp_tree_b200 as R
    from oracle import orc
    W = bench.WORKLOAD
    assert (W["n"], W["d"], W["ntrees"], W["min_leaf"]) == (N, D, T, MINL)
    X = bench.make_points(N, D, W["data_seed"], W["clusters"], W["sigma"])
    Q = bench.make_points(NQ, D, W["query_seed"], W["clusters"], W["sigma"])
    cfg = R.rpTreeCfg(MINL, N, D)
    maxd = cfg.fpMaxTreeDepth
    hp = R.sampleHyperplanes(W["forest_seed"], T, maxd, PNZ, D)
    f = R.RPForest(0)
    f.setHyperplanes(hp, T, maxd)
    f.setPoints(X)
    f.build(maxd, MINL)
    return dict(R=R, orc=orc, X=X, Q=Q, hp=hp, f=f, maxd=maxd, chunk=cfg.fpDataChunkSize)


def _check_node_properties(orc, X, hp, maxd, t_global, e, levels):
    """thr / margins / partition of every node on the given levels, from recomputed keys."""
    child, depth, ss, sz, perm = e["child"], e["depth"], e["seg_start"], e["seg_size"], e["perm"].astype(np.int64)
    for l in levels:
        g = np.flatnonzero((depth == l) & (child >= 0))
        if len(g) == 0:
            continue
        keys = orc.project_all(hp, t_global * maxd + l, X)[perm]          # keys in leaf-concatenation order
        s, z = ss[g], sz[g]
        nh = z // 2
        assert np.all(z >= 3)
        order = np.argsort(s)
        g, s, z, nh = g[order], s[order], z[order], nh[order]
        # boundaries: [s, s+nh) left, [s+nh, s+z) right, interleaved so one reduceat serves both
        bnd = np.stack([s, s + nh], axis=1).ravel()
        covered_end = s[-1] + z[-1]
        assert np.all(bnd[1:] > bnd[:-1])
        mx = np.maximum.reduceat(keys[:covered_end], bnd)
        mn = np.minimum.reduceat(keys[:covered_end], bnd)
        # reduceat segments run to the next boundary: left = [s, s+nh) exactly; right = [s+nh, next s) which is
        # [s+nh, s+z) only when the nodes of the level are contiguous -- they are (level order == segment order)
        assert np.array_equal(s[1:], (s + z)[:-1])
        left_max, right_min = mx[0::2], mn[1::2]
        assert np.array_equal(bits(e["thr"][g]), bits(right_min)), "level %d: thr != min(right keys)" % l
        assert np.all(left_max <= e["thr"][g]), "level %d: a left key exceeds the threshold" % l
        assert np.array_equal(bits(e["mlo"][g]), bits(left_max)), "level %d: mlo != max(left keys)" % l
        # mhi = sorted[nh+1] = second smallest of the right part
        k2 = keys[:covered_end].copy()
        first_min_pos = np.array([a + np.argmin(k2[a:b]) for a, b in zip(s + nh, s + z)]) if len(g) <= 4096 else None
        if first_min_pos is not None:
            k2[first_min_pos] = np.inf
            second = np.minimum.reduceat(k2, bnd)[1::2]
            assert np.array_equal(bits(e["mhi"][g]), bits(second)), "level %d: mhi != second smallest right key" % l


def test_c2_every_tree_is_a_partition_with_planned_leaf_sizes(c2):
    R, f = c2["R"], c2["f"]
    plan = R.topologyPlan(N, c2["maxd"], MINL)
    tp = f.topology()
    for key in ("child", "depth", "seg_start", "seg_size"):
        assert np.array_equal(tp[key], plan[key])
    assert f.leafOrderExact()
    out = f.forestExport()
    for t in range(T):
        cnt = np.bincount(out["perm"][t], minlength=N)
        assert len(cnt) == N and cnt.min() == 1 and cnt.max() == 1, "tree %d: perm is not a permutation" % t
    assert R.treeSize(f) == N
    sizes = R.leafSizes(f)
    assert sizes.min() >= MINL // 2 - 1 and sizes.max() <= MINL


@pytest.mark.parametrize("t", [0, 13, 31])
def test_c2_threshold_and_margin_properties(c2, t):
    e = c2["f"].treeExport(t)
    _check_node_properties(c2["orc"], c2["X"], c2["hp"], c2["maxd"], t, e, levels=range(c2["maxd"]))


@pytest.mark.parametrize("t", [0, 31])
def test_c2_full_size_tree_matches_oracle(c2, t):
    R, orc = c2["R"], c2["orc"]
    of = orc.Forest(c2["X"], R.slice_hyperplanes(c2["hp"], c2["maxd"], t, 1), 1, c2["maxd"], MINL)
    bad = compare_tree(c2["f"].treeExport(t), of.export(0))
    assert not bad, bad


def test_c2_knn_properties_and_two_tree_parity(c2):
    R, orc, f, X, Q = c2["R"], c2["orc"], c2["f"], c2["X"], c2["Q"]
    dist, ids, cnt = f.knnBatch(Q, K)
    assert np.all(cnt == K)
    assert np.all(np.diff(dist, axis=1) >= 0)                                  # sortedness
    off, cand = f.candidatesBatch(Q[:32], -1)
    for i in range(32):
        cs = cand[off[i]:off[i + 1]]
        assert np.all(np.isin(ids[i], cs))                                     # results come from the candidate set
        for j in range(K):                                                     # distances are the reference metric, bit exact
            assert bits(np.array([dist[i, j]]))[0] == bits(np.array([orc.metric_l2(X[ids[i, j]], Q[i])]))[0]
        # stable top-k of the candidate list: recompute from the candidates with the oracle's metric
        dd = np.array([orc.metric_l2(X[c], Q[i]) for c in cs])
        o = np.argsort(dd, kind="stable")[:K]
        assert np.array_equal(cs[o], ids[i]) and np.array_equal(bits(dd[o]), bits(dist[i]))
    # direct parity on a 2-tree forest (trees 0 and 1 of the same draw)
    hp2 = R.slice_hyperplanes(c2["hp"], c2["maxd"], 0, 2)
    g = R.RPForest(0)
    g.setHyperplanes(hp2, 2, c2["maxd"]); g.setPoints(X); g.build(c2["maxd"], MINL)
    og = orc.Forest(X, hp2, 2, c2["maxd"], MINL)
    d2, i2, c2_ = g.knnBatch(Q[:64], K)
    for i in range(64):
        od, oi = og.knn(Q[i], K)
        assert np.array_equal(i2[i, :c2_[i]], oi) and np.array_equal(bits(d2[i, :c2_[i]]), bits(od))
    g.close()


def test_c2_streaming_build_full_size(c2):
    """forest with the rpTreeCfg chunk size (n/100 = 10000 points): shape == plan, nothing lost, every tree a
    permutation, node properties hold for the last chunk... and tree 0 equals the oracle's chunked insert."""
    R, orc, X = c2["R"], c2["orc"], c2["X"]
    chunk = c2["chunk"]
    assert chunk == 10000
    g = R.RPForest(0)
    g.setHyperplanes(c2["hp"], T, c2["maxd"]); g.setPoints(X)
    g.build(c2["maxd"], MINL, chunk=chunk)
    plan = R.topologyPlan(N, c2["maxd"], MINL, chunk=chunk)
    tp = g.topology()
    for key in ("child", "depth", "seg_start", "seg_size"):
        assert np.array_equal(tp[key], plan[key])
    assert g.pointsLost() == 0 == plan["points_lost"] and g.leafOrderExact()
    out = g.forestExport()
    for t in range(T):
        cnt = np.bincount(out["perm"][t], minlength=N)
        assert len(cnt) == N and cnt.min() == 1 and cnt.max() == 1, "tree %d: perm is not a permutation" % t
    of = orc.Forest(X, R.slice_hyperplanes(c2["hp"], c2["maxd"], 0, 1), 1, c2["maxd"], MINL, chunk=chunk)
    bad = compare_tree(g.treeExport(0), of.export(0))
    assert not bad, bad
    g.close()
