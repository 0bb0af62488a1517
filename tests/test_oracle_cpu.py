"""CPU tests: the oracle against every known answer the reference (and its pinned dependencies) provides,
plus self-consistency of the restatement.  No GPU, no /root/reference at run time."""
import json
import os

import numpy as np

from helpers import make_data

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_vector_space_kats():
    """test/Data/RPTreeSpec.hs:22-46 -- the only numeric known answers in the reference's test-suite."""
    from oracle import orc
    vs0 = ([1, 4], [3.4, 2.1]); vs1 = ([0, 3], [6.7, 5.5]); v1 = [1, 2, 3, 4, 5]
    assert np.array_equal(orc.sum_sd(vs0[0], vs0[1], v1), np.array([1, 5.4, 3, 4, 7.1]))
    assert np.array_equal(orc.diff_sd(vs0[0], vs0[1], v1), np.array([-1, 1.4, -3, -4, -2.9]))
    assert orc.inner_ss(vs0[0], vs0[1], vs1[0], vs1[1]) == 0
    assert orc.inner_sd(vs0[0], vs0[1], v1) == 17.3


def test_splitmix_known_answers():
    """splitmix haddock: mkSMGen 42 == SMGen 9297814886316923340 13679457532755275413; plus the vectors SURVEY.md 8c lists."""
    from oracle import orc
    g = orc.mk_smgen(42)
    assert (g.seed, g.gamma) == (9297814886316923340, 13679457532755275413)
    g = orc.mk_smgen(1337)
    assert [orc.next_word64(g) for _ in range(4)] == [0xb5c19e300e8b07b3, 0xd600e0e216c0ac76, 0xc54efc3b3cc5af29, 0x6899ec7461f13294]
    g = orc.mk_smgen(1234)
    assert (g.seed, g.gamma) == (17144079483850646185, 13478418381427711195)
    assert [orc.next_double(g) for _ in range(4)] == [0.11217429899578746, 0.11037222596102358, 0.9210381494952997, 0.7325535387573261]
    g = orc.mk_smgen(1235137)
    assert (g.seed, g.gamma) == (16009143799767038469, 2986435545200735767)
    assert [orc.next_double(g) for _ in range(4)] == [0.08687686605096445, 0.21769893896371983, 0.24825594610219293, 0.9871211537283644]


def test_inner_sd_is_a_right_fold():
    """Internal.hs:375-382: x0*y0 + (x1*y1 + (... + 0)); differs from the left fold in the last bits."""
    from oracle import orc
    rng = np.random.default_rng(0)
    nright = 0
    for _ in range(200):
        d = 40
        idx = np.sort(rng.choice(d, size=12, replace=False)).astype(np.int32)
        val = rng.normal(size=12); x = rng.normal(size=d) * 1e3
        acc = 0.0
        for j in range(len(idx) - 1, -1, -1):
            acc = float(np.float64(val[j]) * np.float64(x[idx[j]])) + acc
        assert orc.inner_sd(idx, val, x) == acc
        left = 0.0
        for j in range(len(idx)):
            left = left + float(np.float64(val[j]) * np.float64(x[idx[j]]))
        nright += left != acc
    assert nright > 0       # the order matters, so the test above really pins it


def test_rptree_cfg():
    """Conduit.hs:132-141; values quoted in SURVEY.md 3.4."""
    from oracle import orc
    assert orc.rptree_cfg(20, 10000, 2) == (9, 100, 1.0)
    maxd, chunk, pnz = orc.rptree_cfg(64, 1000000, 128)
    assert maxd == 14 and chunk == 10000 and abs(pnz - 0.4746) < 1e-4


def test_reference_invariants_batch_and_conduit():
    """RPTreeSpec.hs:51-107 shape: n=10000 2-d points, 10 trees, minLeaf 20, rpTreeCfg depth, pnz 1.0:
    every tree holds all points; chunked build with equal chunks (n/100) keeps them too; knn is sorted."""
    from oracle import orc
    n, d, T, minl = 10000, 2, 10, 20
    rng = np.random.default_rng(1)
    th = rng.uniform(0, 2 * np.pi, n); r = np.sqrt(rng.uniform(0, 1, n))
    X = np.stack([r * np.cos(th), r * np.sin(th)], 1) + np.where(rng.uniform(size=(n, 1)) < 0.5, 0.0, 1.0) * np.array([2.0, 3.0])
    maxd, chunk, _ = orc.rptree_cfg(minl, n, d)
    hp = orc.gen_hyperplanes(99, T, maxd, 1.0, d)
    fb = orc.Forest(X, hp, T, maxd, minl)
    fc = orc.Forest(X, hp, T, maxd, minl, chunk=chunk)
    for t in range(T):
        assert fb.tree_size(t) == n and fc.tree_size(t) == n
    for f in (fb, fc):
        dist, ids = f.knn(np.zeros(2), 5)
        assert len(dist) == 5 and np.all(np.diff(dist) >= 0) and dist.max() < 1
        dist, ids = f.knn(np.zeros(2), 5, dedup=True)
        assert len(set(dist.tolist())) == len(dist) and dist.max() < 1


def test_single_chunk_equals_batch():
    """forest with chunk >= n is literally create (Internal.hs:223-225 vs Conduit.hs:157-160)."""
    from oracle import orc
    X = make_data(1500, 6, 3)
    hp = orc.gen_hyperplanes(5, 3, 7, 0.5, 6)
    a = orc.Forest(X, hp, 3, 7, 10)
    b = orc.Forest(X, hp, 3, 7, 10, chunk=1500)
    for t in range(3):
        ea, eb = a.export(t), b.export(t)
        for k in ea:
            assert np.array_equal(ea[k], eb[k])


def test_topology_is_data_independent():
    """node sizes depend only on (n, minLeaf, maxDepth)."""
    from oracle import orc
    hp = orc.gen_hyperplanes(5, 2, 8, 0.7, 5)
    a = orc.Forest(make_data(3001, 5, 1), hp, 2, 8, 9).export(0)
    b = orc.Forest(make_data(3001, 5, 2, "mixture"), hp, 2, 8, 9).export(1)
    for k in ("child", "depth", "seg_start", "seg_size"):
        assert np.array_equal(a[k], b[k])


def test_margins_and_threshold_are_order_statistics():
    from oracle import orc
    n, d = 501, 4
    X = make_data(n, d, 9)
    hp = orc.gen_hyperplanes(1, 1, 1, 1.0, d)
    e = orc.Forest(X, hp, 1, 1, 3).export(0)
    keys = np.sort([orc.inner_sd(hp[1], hp[2], X[i]) for i in range(n)])
    assert (e["mlo"][0], e["thr"][0], e["mhi"][0]) == (keys[n // 2 - 1], keys[n // 2], keys[n // 2 + 1])
    assert e["seg_size"][1] == n // 2 and e["seg_size"][2] == n - n // 2


def test_candidates_fork_rule():
    """RPTree.hs:309-314 on a hand-built 1-level tree: proj in (mid, thr) or (thr, mid) forks, proj == thr goes right."""
    from oracle import orc
    X = np.array([[0.0], [1.0], [2.0], [10.0]])          # keys = x; sorted: 0,1,2,10 ; nh=2: thr=2, mlo=1, mhi=10
    hp = (np.array([0, 1], np.int64), np.array([0], np.int32), np.array([1.0]))
    f = orc.Forest(X, hp, 1, 1, 1)
    e = f.export(0)
    assert (e["mlo"][0], e["thr"][0], e["mhi"][0]) == (1.0, 2.0, 10.0)
    assert f.candidates(0, [0.5]).tolist() == [0, 1]              # left only
    assert f.candidates(0, [1.9]).tolist() == [0, 1]              # |1-1.9| < |10-1.9| -> left only
    assert f.candidates(0, [2.0]).tolist() == [2, 3]              # == thr -> right
    assert f.candidates(0, [3.0]).tolist() == [0, 1, 2, 3]        # > thr and dl=2 < dr=7 -> both
    assert f.candidates(0, [7.0]).tolist() == [2, 3]              # dl=6 > dr=3 -> right only
    X2 = np.array([[0.0], [9.0], [10.0], [11.0]])                 # thr=10, mlo=9, mhi=11
    f2 = orc.Forest(X2, hp, 1, 1, 1)
    assert f2.candidates(0, [9.9]).tolist() == [0, 1]
    X3 = np.array([[0.0], [1.0], [10.0], [10.5]])                 # thr=10, mlo=1, mhi=10.5 ; proj 9 : dl=8 > dr=1.5 -> both
    f3 = orc.Forest(X3, hp, 1, 1, 1)
    assert f3.candidates(0, [9.0]).tolist() == [0, 1, 2, 3]


def test_recall_definition():
    """RPTree.hs:265-282: mean over trees of per-tree candidate recall; a tree that is a single Tip has recall 1."""
    from oracle import orc
    X = make_data(300, 4, 2)
    hp = orc.gen_hyperplanes(3, 2, 5, 1.0, 4)
    f = orc.Forest(X, hp, 2, 0, 10)         # maxDepth 0 -> every tree is one Tip holding everything
    assert f.recall(X[3], 10) == 1.0
    g = orc.Forest(X, hp, 2, 5, 10)
    r = g.recall(X[3] + 0.01, 10)
    per = []
    bd, bi = orc.brute_knn(X, X[3] + 0.01, 10)
    for t in range(2):
        per.append(len(set(g.candidates(t, X[3] + 0.01).tolist()) & set(bi.tolist())) / 10)
    assert abs(r - np.mean(per)) < 1e-15


def test_golden_fixtures():
    """Self-generated fixtures (tests/golden/make_golden.py): guard the oracle against silent drift.
    They are NOT reference outputs (the reference cannot be run here: no GHC) -- parity stays 'unpinned'."""
    from oracle import orc
    with open(os.path.join(GOLD, "oracle_small.json")) as fh:
        G = json.load(fh)
    c = G["config"]
    X = make_data(c["n"], c["d"], c["data_seed"], c["kind"])
    hp = orc.gen_hyperplanes(c["hp_seed"], c["T"], c["maxd"], c["pnz"], c["d"])
    assert hp[0].tolist() == G["hp_off"] and hp[1].tolist() == G["hp_idx"]
    assert np.array_equal(hp[2].view(np.uint64), np.array(G["hp_val_bits"], np.uint64))
    f = orc.Forest(X, hp, c["T"], c["maxd"], c["minl"])
    for t in range(c["T"]):
        e = f.export(t)
        gt = G["trees"][t]
        internal = e["child"] >= 0
        assert e["thr"][internal].view(np.uint64).tolist() == gt["thr_bits"]
        assert e["mlo"][internal].view(np.uint64).tolist() == gt["mlo_bits"]
        assert e["mhi"][internal].view(np.uint64).tolist() == gt["mhi_bits"]
        assert e["perm"].tolist() == gt["perm"]
    q = np.array(G["query"])
    d, i = f.knn(q, c["k"])
    assert i.tolist() == G["knn_ids"] and d.view(np.uint64).tolist() == G["knn_dist_bits"]
    assert f.recall(q, c["k"]) == G["recall"]


def test_sparse_metrics_keep_the_reference_truncation_quirk():
    """metricSSL2 / metricSDL2 (Internal.hs:389-400) go through binSS / binSDD (Internal.hs:432-470), which stop as soon as
    EITHER operand is exhausted: components past the sparse operand's last stored index are dropped."""
    from oracle import orc
    # u = {1: 3, 4: 1}, v dense
    ui, uv = np.array([1, 4], np.int32), np.array([3.0, 1.0])
    v = np.array([1.0, 1.0, 2.0, 0.0, 5.0, 7.0, 7.0])
    # visited positions 0..4: (0-1), (3-1), (0-2), (0-0), (1-5); positions 5, 6 are never reached
    assert orc.metric_sd_l2(ui, uv, v) == np.sqrt(1.0 + 4.0 + 4.0 + 0.0 + 16.0)
    # sparse-sparse: w = {0: 2, 1: 1, 6: 9}; the merge stops when u is exhausted (after index 4), so w's 6 is dropped
    wi, wv = np.array([0, 1, 6], np.int32), np.array([2.0, 1.0, 9.0])
    # emitted: (0: 0-2), (1: 3-1), (4: 1-0); then u is exhausted
    assert orc.metric_ss_l2(ui, uv, wi, wv) == 3.0
    assert orc.metric_ss_l2(wi, wv, ui, uv) == 3.0
    # empty operand -> nothing emitted -> distance 0
    e_i, e_v = np.zeros(0, np.int32), np.zeros(0)
    assert orc.metric_ss_l2(e_i, e_v, wi, wv) == 0.0 and orc.metric_sd_l2(e_i, e_v, v) == 0.0


def test_sparse_forest_equals_dense_forest_on_the_dense_image():
    """innerSS over SVector points == innerSD over their dense image (missing components contribute exact zeros), so the
    two forests have identical structure; only the metric differs."""
    from oracle import orc
    rng = np.random.default_rng(0)
    n, d, T, maxd, minl = 600, 12, 3, 7, 8
    M = rng.normal(size=(n, d)) * (rng.random((n, d)) < 0.3)
    r, c = np.nonzero(M)
    off = np.zeros(n + 1, np.int64); np.add.at(off, r + 1, 1); off = np.cumsum(off)
    hp = orc.gen_hyperplanes(3, T, maxd, 0.5, d)
    fs = orc.SparseForest((off, c.astype(np.int32), M[r, c]), d, hp, T, maxd, minl)
    fd = orc.Forest(M, hp, T, maxd, minl)
    for t in range(T):
        a, b = fs.export(t), fd.export(t)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    q = M[5] + 0.1
    assert np.array_equal(fs.candidates(0, q), fd.candidates(0, q))
    ds, _ = fs.knn(q, 5)
    dd, _ = fd.knn(q, 5)
    assert not np.array_equal(ds, dd)          # metricSDL2 ignores q's components past each row's last index
