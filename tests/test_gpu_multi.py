"""Multi-GPU behind the C ABI: ONE handle over several GPUs (rpf_create_multi) must return the same
forest, candidates, knn lists and recall sums as a single-GPU handle -- bit for bit, because the trees are independent
(Internal.hs:234-240) and the merge order (distance, GPU, position) is the reference's tree order (RPTree.hs:174-176).
Skipped with fewer than 2 GPUs (the driver's default GPU test box has one); run with `gpurun --gpus 2`."""
import numpy as np
import pytest

from helpers import make_data, bits

pytestmark = pytest.mark.gpu


def _mods():
    import rp_tree_b200 as R
    from oracle import orc
    return R, orc


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("T,kind", [(5, "mixture"), (8, "gauss"), (3, "dupes")])
def test_multi_handle_equals_single_gpu(built, T, kind):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    R, orc = _mods()
    n, d, minl = 40000, 24, 16
    maxd = R.rpTreeCfg(minl, n, d).fpMaxTreeDepth
    X = make_data(n, d, 3, kind)
    hp = orc.gen_hyperplanes(77, T, maxd, 0.3, d)
    Q = np.concatenate([X[:40] + 0.01, make_data(40, d, 9, kind)])
    W = min(_ngpu(), T, 4)

    one = R.RPForest(0)
    one.setHyperplanes(hp, T, maxd)
    one.setPoints(X)
    one.build(maxd, minl)

    for mode in ("set_points+build", "build_from_host", "build_from_host+sink"):
        m = R.RPForest(devices=list(range(W)))
        assert m.numGpus() == W
        m.setHyperplanes(hp, T, maxd)
        if mode == "set_points+build":
            m.setPoints(X)
            m.build(maxd, minl)
        else:
            if mode.endswith("sink"):
                nn = len(one.topology()["child"])
                bufs = dict(thr=np.zeros((T, nn)), mlo=np.zeros((T, nn)), mhi=np.zeros((T, nn)), perm=np.zeros((T, n), np.uint32))
                m.setExportSink(bufs)
            m.buildFromHost(X, maxd, minl)
            m.buildFromHost(X, maxd, minl)          # second call: buffers reused, sink active
        assert m.ntrees == T
        # forest: every tree, forest-wide index
        a, b = one.forestExport(), m.forestExport(bufs if mode.endswith("sink") else None)
        for key in ("thr", "mlo", "mhi", "perm"):
            assert np.array_equal(bits(a[key]) if key != "perm" else a[key], bits(b[key]) if key != "perm" else b[key]), (mode, key)
        for t in (0, T - 1):
            ea, eb = one.treeExport(t), m.treeExport(t)
            assert np.array_equal(ea["perm"], eb["perm"]) and np.array_equal(bits(ea["thr"]), bits(eb["thr"]))
        # queries
        for dedup in (False, True):
            da, ia, ca = one.knnBatch(Q, 10, dedup=dedup)
            db, ib, cb = m.knnBatch(Q, 10, dedup=dedup)
            assert np.array_equal(ca, cb) and np.array_equal(ia, ib) and np.array_equal(bits(da), bits(db)), (mode, dedup)
        oa, ca_ = one.candidatesBatch(Q, -1)
        ob, cb_ = m.candidatesBatch(Q, -1)
        assert np.array_equal(oa, ob) and np.array_equal(ca_, cb_)
        o1, c1 = one.candidatesBatch(Q, T - 1)
        o2, c2 = m.candidatesBatch(Q, T - 1)
        assert np.array_equal(o1, o2) and np.array_equal(c1, c2)
        ra, rb = one.recallSumBatch(Q, 10), m.recallSumBatch(Q, 10)
        assert np.allclose(ra, rb, rtol=0, atol=1e-12), (mode, np.abs(ra - rb).max())
        ba, bb = one.bruteKnnBatch(Q[:8], 5), m.bruteKnnBatch(Q[:8], 5)
        assert np.array_equal(ba[1], bb[1])
        m.close()
    # oracle spot check of the multi-GPU path itself
    of = orc.Forest(X, hp, T, maxd, minl)
    m = R.RPForest(devices=list(range(W)))
    m.setHyperplanes(hp, T, maxd)
    m.buildFromHost(X, maxd, minl)
    dist, ids, cnt = m.knnBatch(Q[:16], 7)
    for i in range(16):
        od, oi = of.knn(Q[i], 7)
        assert np.array_equal(ids[i, :cnt[i]], oi) and np.array_equal(bits(dist[i, :cnt[i]]), bits(od))
    rs = m.recallSumBatch(Q[:4], 5) / T
    for i in range(4):
        assert abs(rs[i] - of.recall(Q[i], 5)) <= 1e-12
    m.close(); one.close()


def test_multi_handle_rejects_what_it_cannot_shard(built):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    R, orc = _mods()
    m = R.RPForest(devices=[0, 1])
    hp = orc.gen_hyperplanes(1, 1, 4, 0.5, 8)
    with pytest.raises(R.RPForestError, match="fewer trees than GPUs"):
        m.setHyperplanes(hp, 1, 4)
    m.close()


def test_multi_handle_checkpoint_round_trip(built, tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    R, orc = _mods()
    n, d, T, minl = 6000, 12, 4, 10
    maxd = R.rpTreeCfg(minl, n, d).fpMaxTreeDepth
    X = make_data(n, d, 5)
    hp = orc.gen_hyperplanes(3, T, maxd, 0.5, d)
    m = R.RPForest(devices=[0, 1])
    m.setHyperplanes(hp, T, maxd)
    m.buildFromHost(X, maxd, minl)
    m.save(tmp_path / "f.rpf")
    g = R.RPForest(devices=[0, 1])
    g.load(tmp_path / "f.rpf")
    Q = X[:10] + 0.02
    for u, v in zip(m.knnBatch(Q, 5), g.knnBatch(Q, 5)):
        assert np.array_equal(u, v)
    m.close(); g.close()
