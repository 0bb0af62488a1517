"""GPU parity for SVector data points (Embed SVector Double x -- the reference bench's own data type,
bench/time/Main.hs:77,113-122): build by innerSS projections, queries as DVectors (metricSDL2) and as SVectors
(metricSSL2), both with the reference's early-stop quirk (Internal.hs:432-470).  Oracle: oracle.orc.SparseForest."""
import numpy as np
import pytest

from helpers import compare_tree, bits

pytestmark = pytest.mark.gpu


def _sparse_data(n, d, density, seed, kind="gauss"):
    rng = np.random.default_rng(seed)
    M = rng.normal(size=(n, d)) if kind == "gauss" else rng.integers(-2, 3, size=(n, d)).astype(np.float64)
    M = M * (rng.random((n, d)) < density)
    if n > 5:
        M[3] = 0.0                     # an SVector with no components
        M[4, : d // 2] = 0.0
    return M


CASES = [
    # n, d, T, maxd, minl, density, kind, chunk
    pytest.param(3000, 24, 4, 8, 10, 0.3, "gauss", None, id="batch-d24"),
    pytest.param(5000, 100, 3, 9, 12, 0.1, "gauss", None, id="batch-d100-sparse10pct"),
    pytest.param(4000, 16, 3, 9, 8, 0.4, "integer", None, id="batch-integer-ties"),
    pytest.param(6000, 32, 3, 9, 10, 0.2, "gauss", 500, id="streamed-chunks"),
    pytest.param(2000, 785, 2, 6, 20, 0.19, "gauss", None, id="mnist-like-d785-odd"),
]


@pytest.mark.parametrize("n,d,T,maxd,minl,density,kind,chunk", CASES)
def test_sparse_points_parity(built, n, d, T, maxd, minl, density, kind, chunk):
    import rp_tree_b200 as R
    from oracle import orc
    M = _sparse_data(n, d, density, 3, kind)
    rows = R.SparseRows.fromDense(M)
    hp = orc.gen_hyperplanes(77, T, maxd, 0.3, d)
    if chunk is None:
        f = R.forestBatch(0, maxd, minl, T, 0.3, d, rows, hyperplanes=hp)
    else:
        f = R.forest(0, maxd, minl, T, chunk, 0.3, d, rows, hyperplanes=hp)
    assert f.pointsAreSparse()
    of = orc.SparseForest((rows.off, rows.idx, rows.val), d, hp, T, maxd, minl, chunk=chunk)
    order = f.leafOrderExact()
    for t in range(T):
        bad = compare_tree(f.treeExport(t), of.export(t), check_order=order)
        assert not bad, "tree %d: %s" % (t, bad)
    rng = np.random.default_rng(9)
    nq = 24
    Qd = M[rng.integers(0, n, size=nq)] + (0.05 * rng.normal(size=(nq, d)) if kind == "gauss" else 0.0)   # DVector queries
    Qs_dense = _sparse_data(nq, d, density, 11, kind)                                                     # SVector queries
    Qs_dense[0] = M[7]
    Qs = R.SparseRows.fromDense(Qs_dense)
    # candidates (projection only)
    for Qx, sparse in ((Qd, False), (Qs, True)):
        off, ids = f.candidatesBatch(Qx, -1)
        for i in range(nq):
            q = (Qs.idx[Qs.off[i]:Qs.off[i + 1]], Qs.val[Qs.off[i]:Qs.off[i + 1]]) if sparse else Qd[i]
            exp = np.concatenate([of.candidates(t, q) for t in range(T)])
            got = ids[off[i]:off[i + 1]]
            assert np.array_equal(got, exp) if order else np.array_equal(np.sort(got), np.sort(exp))
        # knn / knnPQ: distances bit exact (the truncated metrics), ids identical
        for dedup in (False, True):
            for k in (1, 10):
                dist, idk, cnt = f.knnBatch(Qx, k, dedup=dedup)
                for i in range(nq):
                    q = (Qs.idx[Qs.off[i]:Qs.off[i + 1]], Qs.val[Qs.off[i]:Qs.off[i + 1]]) if sparse else Qd[i]
                    od, oi = of.knn(q, k, dedup=dedup)
                    assert cnt[i] == len(od), (sparse, dedup, k, i)
                    assert np.array_equal(bits(dist[i, :cnt[i]]), bits(od)), (sparse, dedup, k, i)
                    if kind == "gauss" and not dedup:
                        # equal distances (e.g. 0 against rows with few components) are ordered by candidate position
                        assert np.array_equal(idk[i, :cnt[i]], oi), (sparse, dedup, k, i)
        # recallWith with the sparse metrics
        r = R.recallWith(R.metricL2, f, 5, Qx)
        if kind == "gauss":
            ro = np.array([of.recall((Qs.idx[Qs.off[i]:Qs.off[i + 1]], Qs.val[Qs.off[i]:Qs.off[i + 1]]) if sparse else Qd[i], 5)
                           for i in range(nq)])
            # exact-zero distances tie (rows whose last component precedes the query's first): the truth set then depends
            # on the tie order (row id here, leaf order in the oracle) -- compare only queries without a tie at the k-th
            bd, _ = f.bruteKnnBatch(Qx, 6)
            clean = bd[:, 4] != bd[:, 5]
            assert np.allclose(r[clean], ro[clean], rtol=0, atol=1e-12)


def test_sparse_queries_need_sparse_points(built):
    import rp_tree_b200 as R
    X = np.random.default_rng(0).normal(size=(500, 8))
    f = R.forestBatch(1, 5, 10, 2, 0.5, 8, X)
    with pytest.raises(R.RPForestError, match="SVector"):
        f.knnBatch(R.SparseRows.fromDense(X[:3]), 3)


def test_sparse_points_reject_malformed_rows(built):
    import rp_tree_b200 as R
    f = R.RPForest(0)
    with pytest.raises(R.RPForestError, match="ascending"):
        f.setPointsSparse(R.SparseRows([0, 2], [3, 1], [1.0, 2.0], 5))
    with pytest.raises(R.RPForestError, match="range"):
        f.setPointsSparse(R.SparseRows([0, 1], [7], [1.0], 5))


def _write_idx3(path, img):
    import struct
    with open(path, "wb") as fh:
        fh.write(bytes([0, 0, 8, 3]) + struct.pack(">3I", *img.shape) + img.tobytes())


@pytest.mark.gpu
def test_c1_mnist_like_single_tree(built, tmp_path):
    """BASELINE configs[0] shape: 784-d MNIST-like IDX file -> single RPTree build + knn k=10, checked against the oracle
    (massive ties at 0: most pixels are background)."""
    import rp_tree_b200 as R
    from oracle import orc
    rng = np.random.default_rng(5)
    n = 3000
    img = np.zeros((n, 28, 28), np.uint8)
    for i in range(n):                                  # blobs of ink on a black background
        cx, cy, r = rng.integers(6, 22), rng.integers(6, 22), rng.integers(2, 6)
        yy, xx = np.ogrid[:28, :28]
        m = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        img[i][m] = rng.integers(1, 256, size=int(m.sum()))
    p = tmp_path / "train-images-idx3-ubyte"
    _write_idx3(p, img)
    rows = R.idx.mnistSparse(p)
    d, minl, k = 784, 10, 10
    cfg = R.rpTreeCfg(minl, n, d)
    maxd = cfg.fpMaxTreeDepth
    hp = orc.gen_hyperplanes(1235137, 1, maxd, cfg.fpProjNzDensity, d)
    f = R.RPForest(0)
    f.setHyperplanes(hp, 1, maxd)
    f.setPointsSparse(rows)
    f.build(maxd, minl)
    of = orc.SparseForest((rows.off, rows.idx, rows.val), d, hp, 1, maxd, minl)
    assert not compare_tree(f.treeExport(0), of.export(0))
    Qd, _ = rows.densify()
    dist, ids, cnt = f.knnBatch(Qd[:24], k)
    for i in range(24):
        od, oi = of.knn(Qd[i], k)
        assert np.array_equal(ids[i, :cnt[i]], oi) and np.array_equal(bits(dist[i, :cnt[i]]), bits(od))
