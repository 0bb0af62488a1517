"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: thresholds / margins bit-exact (uint64 view of the doubles), leaf sets AND leaf order identical,
candidate id lists identical, knn ids identical and distances bit-exact, recall identical up to fp summation.
"""
import numpy as np
import pytest

from helpers import make_data, compare_tree, bits

pytestmark = pytest.mark.gpu


def _mods():
    import rp_tree_b200 as R
    from oracle import orc
    return R, orc


def _build_pair(n, d, T, maxd, minl, pnz, kind="gauss", cap=None, seed=3, generic=False):
    R, orc = _mods()
    X = make_data(n, d, seed, kind)
    hp = orc.gen_hyperplanes(1235137 + seed, T, maxd, pnz, d)
    opts = {"force_generic_bottom": 1} if generic is True else ({"bottom_words64": 1} if generic == "words64" else None)
    f = R.forestBatch(0, maxd, minl, T, pnz, d, X, hyperplanes=hp, bottom_cap=cap, options=opts)
    of = orc.Forest(X, hp, T, maxd, minl)
    return X, hp, f, of


BUILD_CASES = [
    # n, d, T, maxd, minl, pnz, kind, cap
    pytest.param(1000, 8, 3, 6, 10, 0.5, "gauss", None, id="bottom-only"),
    pytest.param(10000, 16, 4, 9, 20, 0.3, "gauss", 1024, id="top4-bottom"),
    pytest.param(10000, 2, 10, 9, 20, 1.0, "gauss", None, id="reference-test-shape"),
    pytest.param(20000, 32, 3, 11, 16, 0.3, "mixture", 256, id="top7-cap256"),
    pytest.param(5000, 2, 6, 8, 10, 0.5, "gauss", 256, id="empty-hyperplanes-ties"),
    pytest.param(6000, 6, 4, 9, 8, 0.5, "integer", 256, id="integer-data-massive-ties"),
    pytest.param(4000, 5, 3, 9, 7, 0.6, "dupes", 256, id="duplicate-rows"),
    pytest.param(3000, 4, 3, 12, 1, 0.7, "gauss", 256, id="minleaf1-deep"),
    pytest.param(777, 3, 2, 20, 0, 1.0, "gauss", 256, id="minleaf0-maxdepth20"),
    pytest.param(9000, 12, 2, 3, 5, 0.4, "gauss", 1024, id="shallow-big-leaves"),
    pytest.param(40000, 24, 2, 12, 12, 0.25, "mixture", 4096, id="cap4096-top4"),
    pytest.param(70000, 8, 2, 13, 10, 0.5, "gauss", 8192, id="cap8192-top4"),
    pytest.param(50000, 3, 2, 12, 9, 0.5, "integer", 4096, id="cap4096-integer-ties"),
    pytest.param(30000, 2, 4, 11, 11, 0.4, "gauss", 4096, id="cap4096-empty-hyperplanes"),
    pytest.param(3000, 1100, 2, 5, 40, 0.02, "gauss", 1024, id="large-d-direct-projection"),
    pytest.param(2500, 300, 2, 6, 20, 0.05, "gauss", 1024, id="d300-r1-tile"),
    pytest.param(4000, 960, 2, 7, 30, 0.1, "gauss", 1024, id="d960-column-blocked-projection"),
    pytest.param(3000, 769, 2, 6, 30, 0.1, "mixture", 1024, id="d769-odd-column-blocks"),
    pytest.param(30000, 8, 3, 12, 8, 0.5, "outlier", 1024, id="one-outlier-collapses-key-prefixes"),
    pytest.param(30000, 8, 3, 12, 8, 0.5, "outlier", 2048, id="one-outlier-collapses-key-prefixes-cap2048"),
]


@pytest.mark.parametrize("generic", [False, True, "words64"], ids=["fast-bottom", "generic-bottom", "fast-bottom-64bit-words"])
@pytest.mark.parametrize("n,d,T,maxd,minl,pnz,kind,cap", BUILD_CASES)
def test_build_parity(built, n, d, T, maxd, minl, pnz, kind, cap, generic):
    X, hp, f, of = _build_pair(n, d, T, maxd, minl, pnz, kind, cap, generic=generic)
    order = f.leafOrderExact()
    problems = []
    for t in range(T):
        bad = compare_tree(f.treeExport(t), of.export(t), check_order=order)
        problems += ["tree %d: %s" % (t, b) for b in bad]
    assert not problems, "\n".join(problems[:20])
    if kind == "gauss" and minl >= 5 and maxd >= 6:
        assert order


SELECT_CASES = [
    # n, d, T, maxd, minl, pnz, kind   (bottom_cap 1024: nodes of 513 .. 1024 points with <= 5 levels below them are eligible)
    pytest.param(20000, 16, 3, 12, 40, 0.3, "gauss", id="625-point-nodes-4-levels"),
    pytest.param(31000, 8, 2, 14, 64, 0.5, "mixture", id="968-point-nodes-4-levels"),
    pytest.param(16000, 6, 3, 12, 70, 0.5, "integer", id="integer-ties-every-node-goes-to-the-second-pass"),
    pytest.param(16000, 5, 3, 12, 70, 0.6, "dupes", id="duplicate-rows"),
    pytest.param(30000, 8, 3, 12, 120, 0.5, "outlier", id="outlier-collapses-prefixes"),
    pytest.param(18000, 8, 2, 12, 150, 0.5, "gauss", id="tips-of-more-than-64-points"),
    pytest.param(1000, 8, 3, 6, 10, 0.5, "gauss", id="bottom-only-root-node"),
]


@pytest.mark.parametrize("n,d,T,maxd,minl,pnz,kind", SELECT_CASES)
def test_build_parity_select_bottom(built, n, d, T, maxd, minl, pnz, kind):
    """Option bottom_select = 1: the warp-per-node kernel (median select + partition where the children split again, sort where
    Tips form) followed by k_bottom3 on the nodes it flagged (prefix ties, equal keys, odd shapes) gives the oracle's forest."""
    R, orc = _mods()
    X = make_data(n, d, 3, kind)
    hp = orc.gen_hyperplanes(77, T, maxd, pnz, d)
    f = R.forestBatch(0, maxd, minl, T, pnz, d, X, hyperplanes=hp, bottom_cap=1024, options={"bottom_select": 1})
    of = orc.Forest(X, hp, T, maxd, minl)
    problems = []
    for t in range(T):
        problems += ["tree %d: %s" % (t, b) for b in compare_tree(f.treeExport(t), of.export(t))]
    assert not problems, "\n".join(problems[:20])
    assert f.leafOrderExact()


GENERIC_TOP_CASES = [c for c in BUILD_CASES if c.id in (
    "top4-bottom", "top7-cap256", "integer-data-massive-ties", "duplicate-rows", "cap4096-integer-ties",
    "one-outlier-collapses-key-prefixes", "cap8192-top4")]


@pytest.mark.parametrize("n,d,T,maxd,minl,pnz,kind,cap", GENERIC_TOP_CASES)
def test_build_parity_generic_top_kernels(built, n, d, T, maxd, minl, pnz, kind, cap):
    """n % 8 == 0 selects the lean top-phase kernels (k_top_compact_lean / k_top_relabel_lean / k_top_scatter_lean);
    option lean_top = 0 keeps the generic streaming kernels covered on the same shapes."""
    R, orc = _mods()
    X = make_data(n, d, 3, kind)
    hp = orc.gen_hyperplanes(1235137 + 3, T, maxd, pnz, d)
    f = R.forestBatch(0, maxd, minl, T, pnz, d, X, hyperplanes=hp, bottom_cap=cap, options={"lean_top": 0})
    of = orc.Forest(X, hp, T, maxd, minl)
    for t in range(T):
        bad = compare_tree(f.treeExport(t), of.export(t), check_order=f.leafOrderExact())
        assert not bad, "tree %d: %s" % (t, bad)


@pytest.mark.parametrize("n", [65536 + 8, 32768, 98304 + 4096])
def test_lean_top_chunk_edges(built, n):
    """chunk boundaries of the lean kernels (32768-point chunks: exact multiples, an 8-point tail, a partial chunk),
    clustered keys that overflow the per-chunk hit list of k_top_compact_lean (integer data) and a wide forest."""
    R, orc = _mods()
    for kind, d, T in (("mixture", 12, 5), ("integer", 3, 2)):
        maxd, minl, pnz = 11, 12, 0.5
        X = make_data(n, d, 17, kind)
        hp = orc.gen_hyperplanes(77, T, maxd, pnz, d)
        f = R.forestBatch(0, maxd, minl, T, pnz, d, X, hyperplanes=hp, bottom_cap=512)
        of = orc.Forest(X, hp, T, maxd, minl)
        for t in range(T):
            bad = compare_tree(f.treeExport(t), of.export(t), check_order=f.leafOrderExact())
            assert not bad, "%s n=%d tree %d: %s" % (kind, n, t, bad)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5])
def test_tiny_inputs(built, n):
    R, orc = _mods()
    d, T, maxd = 3, 2, 4
    X = make_data(max(n, 1), d, 11)[:n].reshape(n, d)
    hp = orc.gen_hyperplanes(5, T, maxd, 1.0, d)
    for minl in (0, 1, 2):
        f = R.forestBatch(0, maxd, minl, T, 1.0, d, X, hyperplanes=hp)
        of = orc.Forest(X, hp, T, maxd, minl)
        for t in range(T):
            bad = compare_tree(f.treeExport(t), of.export(t))
            assert not bad, "n=%d minl=%d tree %d: %s" % (n, minl, t, bad)


def test_all_points_in_every_tree(built):
    """The reference's own invariant (test/Data/RPTreeSpec.hs:67-68): treeSize t == n for every tree."""
    R, orc = _mods()
    n, d, T, minl = 10000, 2, 10, 20
    X = make_data(n, d, 1)
    cfg = R.rpTreeCfg(minl, n, d)
    f = R.forestBatch(42, cfg.fpMaxTreeDepth, minl, T, 1.0, d, X)     # built-in sampler
    assert R.treeSize(f) == n
    for t in range(T):
        assert np.array_equal(np.sort(R.points(f, t)), np.arange(n, dtype=np.uint32))
    # knn of the origin is close (RPTreeSpec.hs:69-74 checks < 1 on its 2-cluster data; here: finds the true 1-NN region)
    dist, ids = R.knn(R.metricL2, 5, f, np.zeros(d))
    assert len(dist) == 5 and np.all(np.diff(dist) >= 0)


def test_direct_projection_kernel_still_exact(built):
    """The per-row fallback kernel (no shared-memory tile; selected here with project_variant = 3)."""
    R, orc = _mods()
    n, d, T, maxd, minl = 3000, 1100, 2, 5, 40
    X = make_data(n, d, 3)
    hp = orc.gen_hyperplanes(4, T, maxd, 0.02, d)
    f = R.forestBatch(0, maxd, minl, T, 0.02, d, X, hyperplanes=hp, options={"project_variant": 3})
    of = orc.Forest(X, hp, T, maxd, minl)
    for t in range(T):
        assert not compare_tree(f.treeExport(t), of.export(t))
    Q = X[:8] + 0.01
    off, ids = f.candidatesBatch(Q, -1)
    for i in range(8):
        assert np.array_equal(ids[off[i]:off[i + 1]], np.concatenate([of.candidates(t, Q[i]) for t in range(T)]))


def test_projection_matches_oracle_fold_order(built):
    """thresholds are projections of data points: bit equality of thr already pins innerSD's right fold; this
    checks a level-0 threshold explicitly against the oracle's inner_sd of the median point."""
    R, orc = _mods()
    n, d = 2001, 64
    X = make_data(n, d, 5) * 1e3
    hp = orc.gen_hyperplanes(9, 1, 1, 0.9, d)
    f = R.forestBatch(0, 1, 10, 1, 0.9, d, X, hyperplanes=hp)
    e = f.treeExport(0)
    keys = np.array([orc.inner_sd(hp[1], hp[2], X[i]) for i in range(n)])
    ks = np.sort(keys, kind="stable")
    assert bits(e["thr"])[0] == bits(ks[n // 2 : n // 2 + 1])[0]
    assert bits(e["mlo"])[0] == bits(ks[n // 2 - 1 : n // 2])[0]
    assert bits(e["mhi"])[0] == bits(ks[n // 2 + 1 : n // 2 + 2])[0]


QUERY_CASES = [
    pytest.param(10000, 16, 4, 9, 20, 0.3, "gauss", 1024, id="gauss"),
    pytest.param(20000, 32, 6, 10, 16, 0.3, "mixture", None, id="mixture"),
    pytest.param(4000, 5, 3, 9, 7, 0.6, "dupes", 256, id="dupes"),
    pytest.param(3000, 3, 8, 8, 10, 0.7, "integer", 256, id="integer"),
    pytest.param(6000, 128, 40, 7, 30, 0.1, "mixture", None, id="d128-T40-chunked"),
    pytest.param(3000, 600, 4, 6, 30, 0.05, "gauss", None, id="d600-small-ring"),
]


@pytest.mark.parametrize("simple_knn", [False, True], ids=["tma-knn", "gather-knn"])
@pytest.mark.parametrize("n,d,T,maxd,minl,pnz,kind,cap", QUERY_CASES)
def test_candidates_knn_recall_parity(built, n, d, T, maxd, minl, pnz, kind, cap, simple_knn):
    R, orc = _mods()
    X, hp, f, of = _build_pair(n, d, T, maxd, minl, pnz, kind, cap)
    f.setOption("force_simple_knn", int(simple_knn))
    rng = np.random.default_rng(17)
    nq = 40
    Q = X[rng.integers(0, n, size=nq)] + (0.05 * rng.normal(size=(nq, d)) if kind != "integer" else 0.0)
    Q[0] = X[0]                       # an exact data point (distance 0, proj == some key)
    order = f.leafOrderExact()
    # candidates, per tree and concatenated
    for t in list(range(T)) + [-1]:
        off, ids = f.candidatesBatch(Q, t)
        for i in range(nq):
            got = ids[off[i]:off[i + 1]]
            exp = of.candidates(t, Q[i]) if t >= 0 else np.concatenate([of.candidates(tt, Q[i]) for tt in range(T)])
            if order:
                assert np.array_equal(got, exp), "candidates differ: tree %d query %d" % (t, i)
            else:
                assert np.array_equal(np.sort(got), np.sort(exp)), "candidate sets differ: tree %d query %d" % (t, i)
    # knn and knnPQ
    for dedup in (False, True):
        for k in (1, 10, 37, 100):
            dist, ids, cnt = f.knnBatch(Q, k, dedup=dedup)
            for i in range(nq):
                od, oi = of.knn(Q[i], k, dedup=dedup)
                assert cnt[i] == len(od), "count: dedup=%s k=%d q=%d: %d vs %d" % (dedup, k, i, cnt[i], len(od))
                assert np.array_equal(bits(dist[i, :cnt[i]]), bits(od)), "distances: dedup=%s k=%d q=%d" % (dedup, k, i)
                if order or kind in ("gauss", "mixture"):
                    assert np.array_equal(ids[i, :cnt[i]], oi), "ids: dedup=%s k=%d q=%d" % (dedup, k, i)
    # recallWith
    r = R.recallWith(R.metricL2, f, 10, Q)
    if kind in ("gauss", "mixture"):        # row identity == vector identity (no duplicate rows)
        ro = np.array([of.recall(Q[i], 10) for i in range(nq)])
        assert np.allclose(r, ro, rtol=0, atol=1e-12), (r, ro)
    # brute force
    bd, bi = f.bruteKnnBatch(Q[:8], 10)
    for i in range(8):
        od, oi = orc.brute_knn(X, Q[i], 10)
        assert np.array_equal(bits(bd[i]), bits(od)) and np.array_equal(bi[i], oi)


def test_fork_heavy_queries(built):
    """Queries sitting exactly between threshold and margin force `candidates` to take both branches."""
    R, orc = _mods()
    X, hp, f, of = _build_pair(2000, 2, 3, 7, 10, 1.0, "gauss", 256)
    Q = []
    for t in range(3):
        e = of.export(t)
        # build queries whose projection on level 0 lies just beside the root threshold
        h0 = (hp[1][hp[0][t * 7]:hp[0][t * 7 + 1]], hp[2][hp[0][t * 7]:hp[0][t * 7 + 1]])
        v = np.zeros(2); v[h0[0]] = h0[1]
        for target in (np.nextafter(e["thr"][0], -np.inf), np.nextafter(e["thr"][0], np.inf), e["thr"][0],
                       0.5 * (e["thr"][0] + e["mlo"][0]), 0.5 * (e["thr"][0] + e["mhi"][0])):
            Q.append(v * (target / np.dot(v, v)))
    Q = np.array(Q)
    off, ids = f.candidatesBatch(Q, -1)
    for i in range(len(Q)):
        exp = np.concatenate([of.candidates(t, Q[i]) for t in range(3)])
        assert np.array_equal(ids[off[i]:off[i + 1]], exp)
    dist, idk, cnt = f.knnBatch(Q, 5)
    for i in range(len(Q)):
        od, oi = of.knn(Q[i], 5)
        assert np.array_equal(idk[i, :cnt[i]], oi)


def test_merge_topk_equals_single_forest(built):
    """Tree-sharded build on one device emulating G ranks: merged per-shard top-k == whole-forest top-k."""
    R, orc = _mods()
    n, d, T, maxd, minl, pnz = 8000, 12, 8, 9, 12, 0.4
    X = make_data(n, d, 2, "mixture")
    hp = orc.gen_hyperplanes(77, T, maxd, pnz, d)
    whole = R.forestBatch(0, maxd, minl, T, pnz, d, X, hyperplanes=hp)
    Q = X[:32] + 0.01
    for dedup in (False, True):
        wd, wi, wc = whole.knnBatch(Q, 10, dedup=dedup)
        for G in (2, 4):
            per = T // G
            ds, is_, cs = [], [], []
            for g in range(G):
                sh = R.forestBatch(0, maxd, minl, T, pnz, d, X, hyperplanes=hp, t_first=g * per, t_local=per)
                a, b, c = sh.knnBatch(Q, 10, dedup=dedup)
                ds.append(a); is_.append(b); cs.append(c)
            md, mi, mc = whole.mergeTopk(np.stack(ds), np.stack(is_), np.stack(cs), dedup=dedup)
            assert np.array_equal(mc, wc)
            assert np.array_equal(bits(md), bits(wd)) and np.array_equal(mi, wi)


def test_chunked_single_chunk_equals_batch(built):
    R, orc = _mods()
    n, d, T, maxd, minl = 3000, 6, 2, 7, 10
    X = make_data(n, d, 8)
    hp = orc.gen_hyperplanes(3, T, maxd, 0.5, d)
    f = R.forest(0, maxd, minl, T, n, 0.5, d, X, hyperplanes=hp)
    of = orc.Forest(X, hp, T, maxd, minl, chunk=n)
    for t in range(T):
        assert not compare_tree(f.treeExport(t), of.export(t))
    assert f.pointsLost() == 0


STREAM_CASES = [
    # n, d, T, maxd, minl, chunk, pnz, kind, cap
    pytest.param(1000, 8, 3, 8, 10, 100, 0.5, "gauss", None, id="small-chunks"),
    pytest.param(10000, 16, 4, 10, 20, 1000, 0.3, "gauss", 256, id="chunk1000-top-phase-cap256"),
    pytest.param(20000, 16, 3, 11, 16, 5000, 0.3, "mixture", 1024, id="chunk5000-top-phase"),
    pytest.param(10007, 12, 3, 10, 12, 1001, 0.4, "gauss", 256, id="ragged-chunks-unaligned"),
    pytest.param(6000, 6, 4, 10, 8, 600, 0.5, "integer", 256, id="integer-ties"),
    pytest.param(4000, 5, 3, 9, 7, 512, 0.6, "dupes", 256, id="duplicate-rows"),
    pytest.param(5000, 2, 6, 9, 10, 700, 0.5, "gauss", 256, id="empty-hyperplanes"),
    pytest.param(1000, 4, 3, 10, 5, 64, 0.7, "gauss", None, id="reference-drops-subtrees"),
    pytest.param(3000, 4, 2, 6, 40, 999, 0.7, "gauss", None, id="tiny-last-chunk-wipes-tree"),
    pytest.param(997, 3, 2, 20, 6, 101, 1.0, "gauss", None, id="deep-maxdepth20"),
    pytest.param(1000, 3, 2, 5, 2, 1, 1.0, "gauss", None, id="chunk-of-one"),
    pytest.param(500, 3, 2, 7, 64, 20, 1.0, "gauss", None, id="chunks-accumulate-in-root-tip"),
    pytest.param(3000, 4, 3, 12, 1, 300, 0.7, "gauss", 256, id="minleaf1"),
    pytest.param(600, 3, 2, 8, 0, 50, 1.0, "gauss", 256, id="minleaf0"),
    pytest.param(40000, 8, 2, 12, 1000, 10000, 0.5, "gauss", 1024, id="big-leaves-resplit-2000"),
    pytest.param(100000, 16, 2, 12, 32, 1000, 0.3, "mixture", 1024, id="rptreecfg-like-100-chunks"),
    pytest.param(10003, 6, 3, 10, 10, 5000, 0.5, "gauss", 256, id="three-point-last-chunk-in-top-phase"),
    pytest.param(10001, 6, 3, 10, 10, 2500, 0.5, "integer", 256, id="one-point-last-chunk-in-top-phase-ties"),
    pytest.param(60000, 8, 2, 13, 10, 3000, 0.5, "gauss", 512, id="20-chunks-top-phase-cap512"),
]


@pytest.mark.parametrize("generic", [False, True], ids=["fast-bottom", "generic-bottom"])
@pytest.mark.parametrize("n,d,T,maxd,minl,chunk,pnz,kind,cap", STREAM_CASES)
def test_streaming_build_parity(built, n, d, T, maxd, minl, chunk, pnz, kind, cap, generic):
    """forest / tree with chunk < n (Conduit.hs:58-121, insert Bin + Tip cases Internal.hs:257-297) against the oracle's
    chunked insert: shape, thresholds (running averages), margins (semigroup), leaf sets and leaf order."""
    R, orc = _mods()
    X = make_data(n, d, 21, kind)
    hp = orc.gen_hyperplanes(99, T, maxd, pnz, d)
    f = R.forest(0, maxd, minl, T, chunk, pnz, d, X, hyperplanes=hp, bottom_cap=cap,
                 options={"force_generic_bottom": 1} if generic else None)
    of = orc.Forest(X, hp, T, maxd, minl, chunk=chunk)
    plan = R.topologyPlan(n, maxd, minl, chunk=chunk)
    assert f.pointsLost() == plan["points_lost"] == n - of.tree_size(0)
    order = f.leafOrderExact()
    problems = []
    for t in range(T):
        bad = compare_tree(f.treeExport(t), of.export(t), check_order=order)
        problems += ["tree %d: %s" % (t, b) for b in bad]
    assert not problems, "\n".join(problems[:20])
    assert order
    # queries against the streamed forest
    rng = np.random.default_rng(5)
    nq = 16
    Q = X[rng.integers(0, n, size=nq)] + (0.05 * rng.normal(size=(nq, d)) if kind != "integer" else 0.0)
    off, ids = f.candidatesBatch(Q, -1)
    for i in range(nq):
        exp = np.concatenate([of.candidates(t, Q[i]) for t in range(T)])
        assert np.array_equal(ids[off[i]:off[i + 1]], exp), "candidates differ: query %d" % i
    dist, idk, cnt = f.knnBatch(Q, 10)
    for i in range(nq):
        od, oi = of.knn(Q[i], 10)
        assert cnt[i] == len(od) and np.array_equal(bits(dist[i, :cnt[i]]), bits(od))
        if kind in ("gauss", "mixture"):
            assert np.array_equal(idk[i, :cnt[i]], oi)


def test_streaming_rebuild_and_batch_on_same_handle(built):
    """A handle can go batch -> streamed -> batch; the streamed result does not depend on what ran before."""
    R, orc = _mods()
    n, d, T, maxd, minl = 6000, 8, 3, 9, 12
    X = make_data(n, d, 4)
    hp = orc.gen_hyperplanes(5, T, maxd, 0.5, d)
    f = R.forestBatch(0, maxd, minl, T, 0.5, d, X, hyperplanes=hp)
    for chunk in (500, None, 750):
        f.build(maxd, minl, chunk=chunk)
        of = orc.Forest(X, hp, T, maxd, minl, chunk=chunk if chunk else n)
        for t in range(T):
            assert not compare_tree(f.treeExport(t), of.export(t)), "chunk=%r tree %d" % (chunk, t)


INSERT_CASES = [
    # n, d, T, maxd, minl, chunk, pnz, kind, cap
    pytest.param(1000, 8, 3, 8, 10, 100, 0.5, "gauss", None, id="small-chunks"),
    pytest.param(10007, 12, 3, 10, 12, 1001, 0.4, "gauss", 256, id="ragged-chunks-top-phase-cap256"),
    pytest.param(3000, 4, 2, 6, 40, 999, 0.7, "gauss", None, id="tiny-last-chunk-wipes-tree"),
    pytest.param(500, 3, 2, 7, 64, 20, 1.0, "gauss", None, id="chunks-accumulate-in-root-tip"),
    pytest.param(10001, 6, 3, 10, 10, 2500, 0.5, "integer", 256, id="one-point-last-chunk-ties"),
    pytest.param(40000, 16, 2, 12, 32, 4000, 0.3, "mixture", 1024, id="ten-chunks-top-phase"),
]


@pytest.mark.parametrize("n,d,T,maxd,minl,chunk,pnz,kind,cap", INSERT_CASES)
def test_incremental_insert_equals_fold_after_every_chunk(built, n, d, T, maxd, minl, chunk, pnz, kind, cap):
    """rpf_insert_begin / rpf_insert_chunk: the forest after EVERY chunk equals the oracle's chunked insert over the rows seen
    so far (Conduit.hs:157-176: the fold's accumulator after each `insertMulti`, Internal.hs:243-255) -- neither n nor the
    number of chunks is known to the engine in advance -- and it answers queries between chunks."""
    R, orc = _mods()
    X = make_data(n, d, 21, kind)
    hp = orc.gen_hyperplanes(99, T, maxd, pnz, d)
    f = R.RPForest(0)
    if cap is not None:
        f.setBottomCap(cap)
    f.setHyperplanes(hp, T, maxd)
    f.insertBegin(d, maxd, minl)
    assert f.topology()["child"].tolist() == [-1] and f.topology()["seg_size"].tolist() == [0]      # Tip () mempty
    nchunks = (n + chunk - 1) // chunk
    check_at = {1, 2, nchunks // 2, nchunks - 1, nchunks}
    rng = np.random.default_rng(5)
    for c in range(nchunks):
        f.insertChunk(X[c * chunk:(c + 1) * chunk])
        seen = min(n, (c + 1) * chunk)
        if c + 1 not in check_at:
            continue
        of = orc.Forest(X[:seen], hp, T, maxd, minl, chunk=chunk)
        assert f.pointsLost() == seen - of.tree_size(0)
        problems = []
        for t in range(T):
            problems += ["after chunk %d tree %d: %s" % (c + 1, t, b) for b in compare_tree(f.treeExport(t), of.export(t))]
        assert not problems, "\n".join(problems[:20])
        assert f.leafOrderExact()
        Q = X[rng.integers(0, seen, size=8)] + (0.05 * rng.normal(size=(8, d)) if kind != "integer" else 0.0)
        dist, idk, cnt = f.knnBatch(Q, 5)
        for i in range(len(Q)):
            od, oi = of.knn(Q[i], 5)
            assert cnt[i] == len(od) and np.array_equal(bits(dist[i, :cnt[i]]), bits(od))
            if kind in ("gauss", "mixture"):
                assert np.array_equal(idk[i, :cnt[i]], oi)
    f.insertEnd()
    # closed session: the forest and the points stay
    of = orc.Forest(X, hp, T, maxd, minl, chunk=chunk)
    for t in range(T):
        assert not compare_tree(f.treeExport(t), of.export(t))
    # ... and the handle is an ordinary one again
    f.build(maxd, minl)
    ob = orc.Forest(X, hp, T, maxd, minl)
    for t in range(T):
        assert not compare_tree(f.treeExport(t), ob.export(t))


def test_incremental_insert_from_a_row_source(built):
    """`forest` fed by a generator of single rows (the conduit source of Conduit.hs:104-121) == `forest` over the matrix;
    unequal chunk sizes are accepted by rpf_insert_chunk (the oracle has no counterpart: checked for shape + membership)."""
    R, orc = _mods()
    n, d, T, maxd, minl, chunk = 2503, 6, 3, 9, 8, 250
    X = make_data(n, d, 3)
    hp = orc.gen_hyperplanes(7, T, maxd, 0.5, d)
    f = R.forest(0, maxd, minl, T, chunk, 0.5, d, (row for row in X), hyperplanes=hp)
    of = orc.Forest(X, hp, T, maxd, minl, chunk=chunk)
    assert f.n == n
    for t in range(T):
        assert not compare_tree(f.treeExport(t), of.export(t))
    g = R.RPForest(0)
    g.setHyperplanes(hp, T, maxd)
    g.insertBegin(d, maxd, minl)
    a = 0
    for m in (700, 3, 0, 1200, 600):
        g.insertChunk(X[a:a + m]); a += m
    assert a == n and g.n == n
    tp = g.topology()
    assert tp["seg_size"][0] + g.pointsLost() == n
    e = g.treeExport(0)
    kept = e["perm"][:tp["seg_size"][0]]
    assert len(np.unique(kept)) == len(kept) and kept.max() < n


def test_incremental_insert_state_errors(built):
    R, orc = _mods()
    d, T, maxd = 4, 2, 5
    hp = orc.gen_hyperplanes(5, T, maxd, 1.0, d)
    f = R.RPForest(0)
    with pytest.raises(R.RPForestError, match="hyperplanes"):
        f.insertBegin(d, maxd, 4)
    f.setHyperplanes(hp, T, maxd)
    with pytest.raises(R.RPForestError, match="insert_begin"):
        f.insertChunk(np.zeros((3, d)))
    f.insertBegin(d, maxd, 4)
    f.insertChunk(make_data(100, d, 1))
    with pytest.raises(R.RPForestError, match="insert session"):
        f.build(maxd, 4)
    f.setPoints(make_data(50, d, 2))               # replaces the points: the session is gone
    with pytest.raises(R.RPForestError, match="insert_begin"):
        f.insertChunk(np.zeros((3, d)))
    f.build(maxd, 4)
    assert f.topology()["seg_size"][0] == 50


@pytest.mark.parametrize("kind,scale", [("gauss", 1.0), ("mixture", 1.0), ("integer", 1.0), ("dupes", 1.0), ("allsame", 1.0),
                                        ("gauss", 1e19), ("gauss", 1e-30), ("offset", 1.0)])
def test_knn_fp32_filter_equals_exact_kernel_and_oracle(built, kind, scale):
    """k_knn_f32 (fp32 filter pass + exact re-rank of the survivors) against the exact kernel and the oracle: well separated data
    (a handful of survivors), massive ties (integer data, duplicate rows, all rows equal: the query is flagged and answered by
    the exact kernel), norms beyond fp32's range (filter refused), tiny magnitudes (squares underflow in fp32) and data far from
    the origin (the fp32 image loses the low bits the distances live in: wide margin, more survivors, same answer)."""
    R, orc = _mods()
    n, d, T, maxd, minl, k = 6000, 16, 6, 8, 12, 10
    if kind == "allsame":
        X = np.tile(make_data(1, d, 2), (n, 1))
    elif kind == "offset":
        X = make_data(n, d, 2) * 1e-3 + 1000.0
    else:
        X = make_data(n, d, 2, kind) * scale
    hp = orc.gen_hyperplanes(11, T, maxd, 0.5, d)
    f = R.forestBatch(0, maxd, minl, T, 0.5, d, X, hyperplanes=hp)
    of = orc.Forest(X, hp, T, maxd, minl)
    rng = np.random.default_rng(9)
    Q = X[rng.integers(0, n, size=96)] + (0.0 if kind in ("integer", "allsame") else 0.01 * scale * rng.normal(size=(96, d)))
    f.setOption("knn_filter32", 0)
    d0, i0, c0 = f.knnBatch(Q, k)
    f.setOption("knn_filter32", 1)
    d1, i1, c1 = f.knnBatch(Q, k)
    assert np.array_equal(c0, c1) and np.array_equal(bits(d0), bits(d1)) and np.array_equal(i0, i1)
    for q in range(0, 96, 8):
        od, oi = of.knn(Q[q], k)
        assert c1[q] == len(od) and np.array_equal(bits(d1[q, :c1[q]]), bits(od))
        if kind in ("gauss", "mixture", "offset"):
            assert np.array_equal(i1[q, :c1[q]], oi)


def test_streaming_unsupported_shape_reports(built):
    """A Tip of more than 8192 points that must be re-split is outside the streaming path's limits."""
    R, orc = _mods()
    n, d = 30000, 4
    X = make_data(n, d, 4)
    hp = orc.gen_hyperplanes(5, 1, 4, 1.0, d)
    with pytest.raises(R.RPForestError, match="8192"):
        R.forest(0, 4, 9000, 1, 6000, 1.0, d, X, hyperplanes=hp)


def test_profile_and_launch_count(built):
    R, orc = _mods()
    X = make_data(30000, 16, 4)
    f = R.RPForest(0)
    f.setProfiling(True)
    f.setPoints(X)
    f.genHyperplanes(1, 4, 11, 0.3, 16)
    f.build(11, 16)
    prof = f.profile()
    assert prof["project"][1] >= 1 and prof["bottom"][1] >= 1
    assert f.lastDeviceMs() > 0 and f.launchCount() > 0


@pytest.mark.parametrize("n,d,T,maxd,minl,pnz,kind,cap", QUERY_CASES[:4])
def test_knnH_parity(built, n, d, T, maxd, minl, pnz, kind, cap):
    """knnH (RPTree.hs:199-217): margin-priority leaf search.  Same leaves, same (reverse pop) order, bit-exact distances;
    equal priorities are ordered by (tree, leaf position) on both sides (unpinned in the reference: heaps internals)."""
    R, orc = _mods()
    X, hp, f, of = _build_pair(n, d, T, maxd, minl, pnz, kind, cap)
    rng = np.random.default_rng(23)
    nq = 40
    Q = X[rng.integers(0, n, size=nq)] + (0.05 * rng.normal(size=(nq, d)) if kind != "integer" else 0.0)
    for k in (1, 10, 50, 200):
        dist, ids, cnt = f.knnHBatch(Q, k)
        for i in range(nq):
            od, oi = of.knn_h(Q[i], k)
            assert cnt[i] == len(oi), "count: k=%d q=%d: %d vs %d" % (k, i, cnt[i], len(oi))
            assert np.array_equal(ids[i, :cnt[i]], oi), "ids: k=%d q=%d" % (k, i)
            assert np.array_equal(bits(dist[i, :cnt[i]]), bits(od)), "distances: k=%d q=%d" % (k, i)
    d1, i1 = R.knnH(R.metricL2, 10, f, Q[0])
    od, oi = of.knn_h(Q[0], 10)
    assert np.array_equal(i1, oi) and np.array_equal(bits(d1), bits(od))


@pytest.mark.parametrize("with_points", [True, False])
@pytest.mark.parametrize("chunk", [None, 700])
def test_checkpoint_round_trip(built, tmp_path, with_points, chunk):
    """serialise / deserialise counterpart (Internal.hs:185-196): a restored forest answers exactly like the original."""
    R, orc = _mods()
    n, d, T, maxd, minl = 5000, 10, 3, 9, 12
    X = make_data(n, d, 6)
    hp = orc.gen_hyperplanes(8, T, maxd, 0.5, d)
    f = R.forest(0, maxd, minl, T, chunk if chunk else n, 0.5, d, X, hyperplanes=hp)
    path = tmp_path / "forest.rpf"
    R.serialiseRPForest(f, path, with_points=with_points)
    if with_points:
        g = R.deserialiseRPForest(path)
    else:
        g = R.RPForest(0)
        with pytest.raises(R.RPForestError, match="carries no points"):
            g.load(path)
        g.setPoints(X)
        g.load(path)
    assert g.ntrees == T and g.n == n and g.d == d and g.leafOrderExact() == f.leafOrderExact()
    for t in range(T):
        a, b = f.treeExport(t), g.treeExport(t)
        for key in a:
            assert np.array_equal(a[key], b[key]), key
    Q = X[:20] + 0.01
    for fn in (lambda h: h.knnBatch(Q, 7), lambda h: h.candidatesBatch(Q, -1), lambda h: h.knnHBatch(Q, 7)):
        ra, rb = fn(f), fn(g)
        for u, v in zip(ra, rb):
            assert np.array_equal(u, v)
    assert np.array_equal(f.recallSumBatch(Q, 5), g.recallSumBatch(Q, 5))
    # a garbage file is refused
    bad = tmp_path / "bad.rpf"
    bad.write_bytes(b"not a checkpoint")
    with pytest.raises(R.RPForestError):
        g.load(bad)
    # ... and so is a truncated, padded or internally corrupt one (deserialiseRPForest returns Left, Internal.hs:191-196):
    # nothing a query kernel indexes with may come from the file unchecked
    raw = bytearray(path.read_bytes())
    nn = len(f.topology()["child"])
    hdr = 80                                            # CkptHeader
    o_hpoff = hdr
    o_hpidx = o_hpoff + (T * maxd + 1) * 8
    nnz = len(hp[1])
    o_start = o_hpidx + nnz * 12
    o_child = o_start + nn * 8
    o_perm = o_child + nn * 8 + (int(f.topology()["depth"].max()) + 2) * 8 + (int(f.topology()["depth"].max()) + 1) * 4 + T * nn * 24

    def variant(name, mutate):
        b = bytearray(raw)
        b = mutate(b) or b
        q = tmp_path / (name + ".rpf")
        q.write_bytes(bytes(b))
        return q

    def put(b, off, val, fmt):
        import struct
        b[off:off + struct.calcsize(fmt)] = struct.pack(fmt, val)

    cases = {
        "truncated": lambda b: b[:-100],
        "padded": lambda b: b + b"\0" * 8,
        "hp_index": lambda b: put(b, o_hpidx + 4 * (nnz // 2), d + 3, "<i"),
        "hp_offset": lambda b: put(b, o_hpoff + 8 * 3, 10 ** 9, "<q"),
        "child": lambda b: put(b, o_child + 4 * 1, nn + 7, "<i"),
        "segment": lambda b: put(b, o_start + 4 * (nn - 1), n + 1, "<I"),
        "perm_row": lambda b: put(b, o_perm + 4 * 17, n + 5, "<I"),
        "header_T": lambda b: put(b, 8 + 8 + 32 + 4, 10 ** 6, "<i"),
    }
    for name, mut in cases.items():
        q = variant(name, mut)
        with pytest.raises(R.RPForestError):
            g.load(q)
    # the handle is still usable after the refusals
    if with_points:
        g.load(path)
    else:
        g.setPoints(X)
        g.load(path)
    assert np.array_equal(g.knnBatch(Q, 7)[1], f.knnBatch(Q, 7)[1])


def test_repeated_builds_replay_the_graph_and_follow_new_data(built):
    """From the third build of an unchanged configuration the launch sequence is replayed as a CUDA graph: every
    rebuild must still equal the oracle, also after the points / hyperplanes / capacity changed in between."""
    R, orc = _mods()
    n, d, T, maxd, minl = 20000, 16, 3, 11, 16
    hp = orc.gen_hyperplanes(5, T, maxd, 0.4, d)
    f = R.RPForest(0)
    f.setHyperplanes(hp, T, maxd)
    for seed in (1, 2):
        X = make_data(n, d, seed, "mixture")
        f.setPoints(X)
        of = orc.Forest(X, hp, T, maxd, minl)
        exp = [of.export(t) for t in range(T)]
        for rep in range(5):
            f.build(maxd, minl)
            for t in range(T):
                assert not compare_tree(f.treeExport(t), exp[t]), "seed %d build %d tree %d" % (seed, rep, t)
    # same points, other hyperplanes and another capacity: the cached plan / graph must not survive
    hp2 = orc.gen_hyperplanes(6, T, maxd, 0.4, d)
    f.setHyperplanes(hp2, T, maxd)
    f.setBottomCap(256)
    of = orc.Forest(X, hp2, T, maxd, minl)
    for rep in range(4):
        f.build(maxd, minl)
        for t in range(T):
            assert not compare_tree(f.treeExport(t), of.export(t)), "hp2 build %d tree %d" % (rep, t)
    # graphs off gives the same result
    f.setOption("cuda_graph", 0)
    f.build(maxd, minl)
    assert not compare_tree(f.treeExport(0), of.export(0))


def test_device_resident_exchange_equals_host_exchange(built):
    """rpf_knn_dev + rpf_merge_topk_dev (lists stay on the device between the two calls, as in the NCCL exchange) give the
    same merged result as rpf_knn + rpf_merge_topk with host lists."""
    import torch
    R, orc = _mods()
    n, d, T, maxd, minl, pnz, k = 8000, 12, 8, 9, 12, 0.4, 10
    X = make_data(n, d, 2, "mixture")
    hp = orc.gen_hyperplanes(77, T, maxd, pnz, d)
    whole = R.forestBatch(0, maxd, minl, T, pnz, d, X, hyperplanes=hp)
    Q = X[:64] + 0.01
    dev = torch.device("cuda", 0)
    for dedup in (False, True):
        wd, wi, wc = whole.knnBatch(Q, k, dedup=dedup)
        G, per = 4, T // 4
        gd = torch.empty((G, len(Q), k), dtype=torch.float64, device=dev)
        gi = torch.empty((G, len(Q), k), dtype=torch.int32, device=dev)
        gc = torch.empty((G, len(Q)), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        for g in range(G):
            sh = R.forestBatch(0, maxd, minl, T, pnz, d, X, hyperplanes=hp, t_first=g * per, t_local=per)
            sh.knnBatchDevice(Q, k, gd[g].data_ptr(), gi[g].data_ptr(), gc[g].data_ptr(), dedup=dedup)
        md, mi, mc = whole.mergeTopkDevice(G, len(Q), k, gd.data_ptr(), gi.data_ptr(), gc.data_ptr(), dedup=dedup)
        assert np.array_equal(mc, wc) and np.array_equal(bits(md), bits(wd)) and np.array_equal(mi, wi)


def test_reference_spec_conduit_shape(built):
    """test/Data/RPTreeSpec.hs:87-107 through the engine: `forest` with the rpTreeCfg chunk size on the two-disc data keeps
    every point in every tree (the reference's own assertion), and knn / knnPQ / knnH of the origin stay below 1."""
    R, orc = _mods()
    n, d, T, minl, k = 10000, 2, 10, 20, 5
    rng = np.random.default_rng(1)
    th = rng.uniform(0, 2 * np.pi, n); r = np.sqrt(rng.uniform(0, 1, n))
    X = np.stack([r * np.cos(th), r * np.sin(th)], 1) + np.where(rng.uniform(size=(n, 1)) < 0.5, 0.0, 1.0) * np.array([2.0, 3.0])
    cfg = R.rpTreeCfg(minl, n, d)
    assert R.topologyPlan(n, cfg.fpMaxTreeDepth, minl, chunk=cfg.fpDataChunkSize)["points_lost"] == 0
    f = R.forest(42, cfg.fpMaxTreeDepth, minl, T, cfg.fpDataChunkSize, 1.0, d, X)
    assert R.treeSize(f) == n and f.pointsLost() == 0
    for t in range(T):
        assert np.array_equal(np.sort(R.points(f, t)[:n]), np.arange(n, dtype=np.uint32))
    q = np.zeros(d)
    for fn in (R.knn, R.knnPQ, R.knnH):
        dist, ids = fn(R.metricL2, k, f, q)
        assert len(dist) >= 1 and dist.max() < 1, fn.__name__


@pytest.mark.parametrize("kind", ["gauss", "all-equal", "sorted-far-first"])
def test_brute_force_topk_sampled_threshold_path(built, kind):
    """n >= 4 * 65536 rows select the truth top-k through k_topk_tau / k_topk_filter / k_topk_final (one pass over the
    distances, threshold from a strided sample); results must equal the nine-pass radix select (option
    force_simple_topk) and the oracle.  `all-equal` (every distance identical: all rows pass the filter) and
    `sorted-far-first` exercise the overflow flag -> k_select_topk fallback and the (distance, row id) tie order."""
    R, orc = _mods()
    n, d = 300_000, 4
    rng = np.random.default_rng(8)
    if kind == "gauss":
        X = rng.normal(size=(n, d))
    elif kind == "all-equal":
        X = np.ones((n, d))
    else:                                   # few distinct distances, many ties
        X = np.repeat(rng.integers(0, 3, size=(n, 1)).astype(np.float64), d, axis=1)
    Q = rng.normal(size=(11, d))
    hp = orc.gen_hyperplanes(3, 1, 2, 1.0, d)
    f = R.RPForest(0)
    f.setHyperplanes(hp, 1, 2); f.setPoints(X); f.build(2, 1000)
    for k in (1, 10, 100):
        f.setOption("force_simple_topk", 0)
        fd, fi = f.bruteKnnBatch(Q, k)
        f.setOption("force_simple_topk", 1)
        sd, si = f.bruteKnnBatch(Q, k)
        assert np.array_equal(bits(fd), bits(sd)) and np.array_equal(fi, si), (kind, k)
        for i in (0, 10):
            od, oi = orc.brute_knn(X, Q[i], k)
            assert np.array_equal(bits(fd[i]), bits(od)) and np.array_equal(fi[i], oi), (kind, k, i)
    f.close()


@pytest.mark.parametrize("T,cap", [(8, 256), (16, 1024), (3, 256)])
def test_export_sink_streams_the_same_forest(built, T, cap):
    """rpf_set_export_sink: buildFromHost downloads perm per bottom-phase tree group while the build runs; the sink must
    hold exactly what a plain forestExport returns (and the oracle's trees), also across repeated builds and for forests
    too small to be split (T < 8)."""
    import torch
    R, orc = _mods()
    n, d, maxd, minl, pnz = 24000, 12, 11, 10, 0.4
    X = make_data(n, d, 21, "mixture")
    hp = orc.gen_hyperplanes(31, T, maxd, pnz, d)
    f = R.RPForest(0)
    f.setHyperplanes(hp, T, maxd)
    if cap:
        f.setBottomCap(cap)
    f.buildFromHost(X, maxd, minl)
    ref = f.forestExport()
    nn = ref["thr"].shape[1]
    sink = {k: torch.zeros((T, nn), dtype=torch.float64, pin_memory=True).numpy() for k in ("thr", "mlo", "mhi")}
    sink["perm"] = torch.zeros((T, n), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    f.setExportSink(sink)
    for rep in range(2):
        for a in sink.values():
            a[...] = 0
        f.buildFromHost(X, maxd, minl)
        out = f.forestExport(sink)
        assert out is sink
        for key in ("thr", "mlo", "mhi"):
            assert np.array_equal(bits(sink[key]), bits(ref[key])), (key, rep)
        assert np.array_equal(sink["perm"], ref["perm"]), rep
    other = f.forestExport()                       # different buffers: the plain download still works with a sink set
    assert np.array_equal(other["perm"], ref["perm"])
    f.build(maxd, minl)                            # resident-data build: no streaming, export copies
    again = f.forestExport(sink)
    assert np.array_equal(again["perm"], ref["perm"]) and np.array_equal(bits(again["thr"]), bits(ref["thr"]))
    f.buildFromHost(X, maxd, minl)                 # the sink holds this forest ...
    f.forestExport(sink)
    f.build(maxd, minl, chunk=5000)                # ... and must not be taken for the streamed one built afterwards
    fresh = f.forestExport()
    stale = f.forestExport(sink)
    assert np.array_equal(stale["perm"], fresh["perm"]) and np.array_equal(bits(stale["thr"]), bits(fresh["thr"]))
    f.build(maxd, minl)
    f.setExportSink(None)
    of = orc.Forest(X, hp, T, maxd, minl)
    for t in range(T):
        assert not compare_tree(f.treeExport(t), of.export(t))
    f.close()


@pytest.mark.parametrize("lean", [1, 0], ids=["lean-then-generic-levels", "generic-levels-only"])
def test_top_phase_beyond_1024_nodes_per_level(built, lean):
    """600k points with bottom_cap 256: the top phase runs 12 levels, the last ones with 2048 nodes per tree -- more than
    the lean kernels' shared-memory tables (1024) and than the shared-memory histogram (512: global-atomic histogram) --
    so one build mixes lean levels, generic levels and both histogram forms.  Integer-valued columns add straddling ties."""
    R, orc = _mods()
    n, d, T, maxd, minl, pnz = 600_000, 4, 2, 13, 100, 0.75
    rng = np.random.default_rng(5)
    X = rng.normal(size=(n, d))
    X[:, 1] = np.round(X[:, 1] * 3)
    hp = orc.gen_hyperplanes(41, T, maxd, pnz, d)
    f = R.forestBatch(0, maxd, minl, T, pnz, d, X, hyperplanes=hp, bottom_cap=256, options={"lean_top": lean})
    assert f.leafOrderExact()
    of = orc.Forest(X, hp, T, maxd, minl)
    for t in range(T):
        bad = compare_tree(f.treeExport(t), of.export(t))
        assert not bad, "tree %d: %s" % (t, bad)


@pytest.mark.parametrize("n,d,T,minl,kind,nq", [
    (30000, 64, 6, 32, "mixture", 900),      # ~1 query per leaf and tree
    (8000, 128, 5, 48, "gauss", 3000),       # many queries per leaf: several passes of 32 queries per (tree, leaf)
    (5000, 20, 3, 64, "dupes", 400),         # exact duplicate rows: massive distance ties -> some queries fall back to the gather kernel
    (4000, 36, 4, 16, "integer", 500),       # ties everywhere, forks
    (300, 8, 2, 64, "gauss", 50),            # the whole data set is a handful of leaves
])
def test_leaf_grouped_tensor_core_rerank_equals_gather_path_and_oracle(built, n, d, T, minl, kind, nq):
    """rerank.cu: (tree, leaf)-grouped FP64 DMMA GEMM selects, survivors recomputed in the reference's arithmetic.  The
    result must be identical -- ids, distance bits, tie order, counts -- to the gather kernel and to the oracle."""
    R, orc = _mods()
    maxd = R.rpTreeCfg(minl, n, d).fpMaxTreeDepth
    X = make_data(n, d, 21, kind)
    hp = orc.gen_hyperplanes(4242, T, maxd, 0.4, d)
    rng = np.random.default_rng(3)
    Q = np.concatenate([X[rng.integers(0, n, nq // 2)] + 0.01 * rng.normal(size=(nq // 2, d)), make_data(nq - nq // 2, d, 22, kind)])
    f = R.RPForest(0)
    f.setHyperplanes(hp, T, maxd); f.setPoints(X); f.build(maxd, minl)
    of = orc.Forest(X, hp, T, maxd, minl)
    for k in (1, 10, 37, 100):
        f.setOption("rerank_gemm", 0)
        d0, i0, c0 = f.knnBatch(Q, k)
        f.setOption("rerank_gemm", 2)
        d1, i1, c1 = f.knnBatch(Q, k)
        assert np.array_equal(c0, c1), k
        for i in range(nq):
            assert np.array_equal(i0[i, :c0[i]], i1[i, :c1[i]]), (k, i)
            assert np.array_equal(bits(d0[i, :c0[i]]), bits(d1[i, :c1[i]])), (k, i)
        for i in range(0, nq, max(1, nq // 40)):
            od, oi = of.knn(Q[i], k)
            assert np.array_equal(i1[i, :c1[i]], oi) and np.array_equal(bits(d1[i, :c1[i]]), bits(od)), (k, i)
    # knnPQ (dedup) is not regrouped: same answers with the option on
    f.setOption("rerank_gemm", 2)
    da, ia, ca = f.knnBatch(Q[:64], 10, dedup=True)
    f.setOption("rerank_gemm", 0)
    db, ib, cb = f.knnBatch(Q[:64], 10, dedup=True)
    assert np.array_equal(ia, ib) and np.array_equal(ca, cb)
    f.close()


@pytest.mark.parametrize("opts", [{"project_variant": 5}, {"fused_top": 3}, {"branches": 1}, {"branches": 4, "fused_top": 0},
                                  {"fuse_relabel_hist": 0, "hist_big_chunk": 0}, {"fuse_relabel_hist": 1, "top_chunk_hist": 57344, "fused_pick_min_tg": 1}],
                         ids=["register-accumulator-projection", "fused-top-chain", "no-branches", "four-branches",
                              "separate-relabel-and-histogram", "fused-relabel-histogram-largest-chunk"])
def test_build_variants_give_the_same_forest(built, opts):
    """A/B hooks of the build: the register-accumulator projection kernel for long rows (k_project_wide) instead of the
    column-blocked launches, the fused top-phase kernels, the number of concurrent branches, the relabel pass with and without
    the next level's histogram (k_top_relabel_hist), the histogram chunk at its 16-bit-counter limit."""
    R, orc = _mods()
    for (n, d, T, minl, kind) in [(70000, 960 if "project_variant" in opts else 24, 3 if "project_variant" in opts else 6, 16, "mixture"),
                                  (66000, 12, 5, 8, "integer")]:
        if "project_variant" in opts and d != 960:
            continue
        maxd = R.rpTreeCfg(minl, n, d).fpMaxTreeDepth
        X = make_data(n, d, 11, kind)
        hp = orc.gen_hyperplanes(99, T, maxd, 0.3 if d < 100 else 0.05, d)
        f = R.forestBatch(0, maxd, minl, T, 0.3, d, X, hyperplanes=hp, options=opts)
        g = R.forestBatch(0, maxd, minl, T, 0.3, d, X, hyperplanes=hp)
        a, b = f.forestExport(), g.forestExport()
        for key in ("thr", "mlo", "mhi"):
            assert np.array_equal(bits(a[key]), bits(b[key])), (opts, key)
        assert np.array_equal(a["perm"], b["perm"]), opts
        of = orc.Forest(X, R.slice_hyperplanes(hp, maxd, 0, 1), 1, maxd, minl)
        assert not compare_tree(f.treeExport(0), of.export(0))
        f.close(); g.close()
