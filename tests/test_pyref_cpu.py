"""The C oracle (oracle/rpt_oracle.c) against a second, independently written pure-Python restatement of the reference
(oracle/pyref.py): thresholds and margins bit for bit, leaf contents in order, candidate lists, knn, knnPQ distances and
recallWith -- for the batch build, the streaming multi-chunk build (incl. shapes that drop points) and data with ties."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_data, bits  # noqa: E402


def _hyperplanes(hp, T, maxd):
    off, idx, val = hp
    return [[(idx[off[t * maxd + l]:off[t * maxd + l + 1]].tolist(), val[off[t * maxd + l]:off[t * maxd + l + 1]].tolist())
             for l in range(maxd)] for t in range(T)]


def _compare(pt, e, g=0, path="root"):
    """python tree vs the oracle's BFS export; returns a list of mismatches"""
    bad = []
    if pt[0] == "tip":
        if e["child"][g] >= 0:
            return ["%s: python Tip, oracle Bin" % path]
        s, z = int(e["seg_start"][g]), int(e["seg_size"][g])
        if e["perm"][s:s + z].tolist() != list(pt[1]):
            bad.append("%s: leaf contents / order differ" % path)
        return bad
    if e["child"][g] < 0:
        return ["%s: python Bin, oracle Tip" % path]
    _, thr, (lo, hi), lt, rt = pt
    for name, a, b in (("thr", thr, e["thr"][g]), ("mlo", lo, e["mlo"][g]), ("mhi", hi, e["mhi"][g])):
        if bits(np.array([a]))[0] != bits(np.array([b]))[0]:
            bad.append("%s: %s %r vs %r" % (path, name, a, b))
    c = int(e["child"][g])
    return bad + _compare(lt, e, c, path + "L") + _compare(rt, e, c + 1, path + "R")


CASES = [
    # n, d, T, maxd, minl, pnz, kind, chunk
    pytest.param(300, 5, 2, 6, 8, 0.6, "gauss", None, id="batch"),
    pytest.param(257, 3, 2, 9, 1, 1.0, "gauss", None, id="batch-minleaf1"),
    pytest.param(200, 4, 2, 5, 6, 0.7, "integer", None, id="batch-ties"),
    pytest.param(180, 3, 2, 6, 5, 0.8, "dupes", None, id="batch-duplicate-rows"),
    pytest.param(2, 3, 1, 4, 0, 1.0, "gauss", None, id="two-points"),
    pytest.param(400, 4, 2, 6, 10, 0.6, "gauss", 64, id="stream-64"),
    pytest.param(333, 3, 2, 7, 4, 0.9, "gauss", 50, id="stream-ragged-last-chunk"),
    pytest.param(240, 4, 2, 6, 6, 0.7, "integer", 40, id="stream-ties"),
    pytest.param(100, 2, 2, 8, 2, 1.0, "gauss", 7, id="stream-tiny-chunks-drop-points"),
]


@pytest.mark.parametrize("n,d,T,maxd,minl,pnz,kind,chunk", CASES)
def test_c_oracle_equals_python_restatement(n, d, T, maxd, minl, pnz, kind, chunk):
    from oracle import orc, pyref
    X = make_data(n, d, 13, kind)
    hp = orc.gen_hyperplanes(77, T, maxd, pnz, d)
    H = _hyperplanes(hp, T, maxd)
    Xl = X.tolist()
    trees = pyref.forest(Xl, H, maxd, minl, chunk)
    of = orc.Forest(X, hp, T, maxd, minl, chunk=chunk)
    for t in range(T):
        e = of.export(t)
        assert of.tree_size(t) == len(pyref.points(trees[t]))
        if chunk is not None and chunk <= 7:
            assert of.tree_size(t) < n          # the reference's empty-piece rule (Internal.hs:274-276) really dropped subtrees
        bad = _compare(trees[t], e)
        assert not bad, "tree %d: %s" % (t, bad[:5])
    rng = np.random.default_rng(3)
    Q = np.concatenate([X[:4] + 0.01, rng.normal(size=(4, d))]) if n >= 4 else rng.normal(size=(3, d))
    k = 5
    for q in Q:
        ql = q.tolist()
        for t in range(T):
            assert of.candidates(t, q).tolist() == pyref.candidates(trees[t], H[t], ql)
        od, oi = of.knn(q, k)
        pk = pyref.knn(trees, H, Xl, k, ql)
        assert oi.tolist() == [i for _, i in pk]
        assert np.array_equal(bits(od), bits(np.array([dd for dd, _ in pk])))
        if kind == "gauss":            # distinct distances: knnPQ and recallWith carry no package-internal tie rule
            pd_, _ = of.knn(q, k, dedup=True)
            assert np.array_equal(bits(pd_), bits(np.array(pyref.knn_pq_distances(trees, H, Xl, k, ql))))
            if all(len(pyref.points(tr)) >= k for tr in trees):
                assert abs(of.recall(q, k) - pyref.recall_with(trees, H, Xl, k, ql)) < 1e-12
