"""CPU tests of the product library's host-side logic: the .so loads, exports every symbol the header declares,
and its host-only entry points (sampler, topology, rpTreeCfg) agree with the oracle.  No compute calls."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    import rp_tree_b200 as R
    L = R.lib()
    hdr = open(os.path.join(ROOT, "include", "rpforest.h")).read()
    declared = set(re.findall(r"\b(rpf_[a-z0-9_]+)\s*\(", hdr)) - {"rpf_status"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(L, name), "librpforest.so does not export " + name
    assert declared == set(R.SIGNATURES), declared ^ set(R.SIGNATURES)
    assert L.rpf_abi_version() == 1


def test_no_cpu_fallback(built):
    """Without a CUDA device the engine refuses to come up (it must never route through the oracle)."""
    import torch
    import rp_tree_b200 as R
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(R.RPForestError):
        R.RPForest(0)
    src = "".join(open(os.path.join(ROOT, "rp-tree_b200", f)).read() for f in ("api.py", "_lib.py", "__init__.py", "dist.py"))
    assert "oracle" not in src.replace("no CPU fallback", "")


def test_sampler_matches_oracle_restatement(built):
    import rp_tree_b200 as R
    from oracle import orc
    for seed, T, maxd, pnz, d in [(1235137, 4, 6, 0.3, 16), (42, 2, 14, 0.1, 128), (7, 3, 5, 1.0, 2), (9, 2, 3, 0.0, 10)]:
        a = R.sampleHyperplanes(seed, T, maxd, pnz, d)
        b = orc.gen_hyperplanes(seed, T, maxd, pnz, d)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.array_equal(a[2].view(np.uint64), b[2].view(np.uint64))


def test_rptree_cfg_matches(built):
    import rp_tree_b200 as R
    from oracle import orc
    for minl, n, d in [(20, 10000, 2), (64, 1000000, 128), (10, 1000, 1000), (64, 10**7, 96)]:
        assert tuple(R.rpTreeCfg(minl, n, d)) == orc.rptree_cfg(minl, n, d)


@pytest.mark.parametrize("n,maxd,minl", [(10000, 9, 20), (1000, 6, 10), (777, 20, 0), (3000, 12, 1), (1, 3, 0), (0, 3, 1),
                                         (5, 2, 1), (100003, 14, 64), (4096, 30, 3)])
def test_topology_plan_matches_oracle_tree_shape(built, n, maxd, minl):
    """The engine's arithmetic topology == the shape the oracle's recursion actually produces on data."""
    import rp_tree_b200 as R
    from oracle import orc
    d = 3
    X = np.random.default_rng(n + maxd).normal(size=(max(n, 1), d))[:n].reshape(n, d)
    hp = orc.gen_hyperplanes(1, 1, maxd, 1.0, d)
    e = orc.Forest(X, hp, 1, maxd, minl).export(0)
    tp = R.topologyPlan(n, maxd, minl)
    for k in ("child", "depth", "seg_start", "seg_size"):
        assert np.array_equal(tp[k], e[k]), k


def test_slice_hyperplanes(built):
    import rp_tree_b200 as R
    hp = R.sampleHyperplanes(3, 8, 5, 0.4, 20)
    parts = [R.slice_hyperplanes(hp, 5, g * 2, 2) for g in range(4)]
    assert sum(len(p[1]) for p in parts) == len(hp[1])
    assert np.array_equal(np.concatenate([p[2] for p in parts]), hp[2])
    for p in parts:
        assert p[0][0] == 0 and len(p[0]) == 11 and p[0][-1] == len(p[1])


@pytest.mark.parametrize("n,chunk,maxd,minl", [(1000, 100, 8, 10), (1000, 64, 10, 5), (5000, 50, 12, 8), (300, 7, 6, 3), (2000, 300, 4, 100),
                                               (1000, 1, 5, 2), (1000, 3, 9, 1), (997, 101, 20, 6), (4096, 512, 12, 1), (500, 20, 7, 64),
                                               (10000, 100, 14, 8), (3000, 999, 6, 40), (600, 50, 8, 0), (1000, 1000, 6, 10), (1000, 5000, 6, 10)])
def test_streaming_plan_matches_oracle_chunked_shape(built, n, chunk, maxd, minl):
    """The planner of the streaming build (sizes only, host) == the shape the oracle's chunked insert produces,
    including the subtrees the reference drops when an empty piece reaches a Bin (Internal.hs:279)."""
    import rp_tree_b200 as R
    from oracle import orc
    d = 3
    X = np.random.default_rng(n + chunk).normal(size=(n, d))
    hp = orc.gen_hyperplanes(1, 1, maxd, 1.0, d)
    of = orc.Forest(X, hp, 1, maxd, minl, chunk=chunk)
    e = of.export(0)
    tp = R.topologyPlan(n, maxd, minl, chunk=chunk)
    for k in ("child", "depth", "seg_start", "seg_size"):
        assert np.array_equal(tp[k], e[k]), k
    assert tp["points_lost"] == n - of.tree_size(0)


def test_streaming_plan_keeps_every_point_at_the_reference_spec_shape(built):
    """test/Data/RPTreeSpec.hs:87-92 asserts treeSize == n for `forest` with the rpTreeCfg parameters (n = 10000,
    minLeaf = 20): the planner must not drop anything there."""
    import rp_tree_b200 as R
    cfg = R.rpTreeCfg(20, 10000, 2)
    tp = R.topologyPlan(10000, cfg.fpMaxTreeDepth, 20, chunk=cfg.fpDataChunkSize)
    assert tp["points_lost"] == 0 and tp["seg_size"][0] == 10000


def test_streaming_plan_rejects_oversized_resplit(built):
    import rp_tree_b200 as R
    with pytest.raises(R.RPForestError):
        R.topologyPlan(30000, 4, 9000, chunk=6000)


def test_densify_rows_helper(built):
    import rp_tree_b200 as R
    M = np.array([[0, 1.5, 0, -2.0], [0, 0, 0, 0], [3.0, 0, 0, 0]])
    rows = R.SparseRows.fromDense(M)
    assert rows.n == 3 and rows.d == 4 and list(rows.off) == [0, 2, 2, 3]
    Q, last = rows.densify()
    assert np.array_equal(Q, M) and list(last) == [3, -1, 0]
    bad = R.SparseRows([0, 2], [2, 1], [1.0, 2.0], 4)        # indices not ascending
    with pytest.raises(R.RPForestError):
        bad.densify()


def _write_idx3(path, img):
    import struct
    with open(path, "wb") as fh:
        fh.write(bytes([0, 0, 8, 3]) + struct.pack(">3I", *img.shape) + img.tobytes())


def test_idx_loader_cpu_roundtrip(tmp_path):
    """IDX3 ubyte file -> SVectors of the nonzero pixels / 255 (`mnist`, bench/time/Main.hs:113-125); no GPU needed."""
    import rp_tree_b200 as R
    rng = np.random.default_rng(0)
    img = (rng.integers(0, 256, size=(37, 6, 5)) * (rng.random((37, 6, 5)) < 0.3)).astype(np.uint8)
    p = tmp_path / "train-images-idx3-ubyte"
    _write_idx3(p, img)
    rows = R.idx.mnistSparse(p, 20)
    assert rows.n == 20 and rows.d == 30
    flat = img.reshape(37, -1)[:20]
    for i in range(20):
        a, b = rows.off[i], rows.off[i + 1]
        nzc = np.flatnonzero(flat[i])
        assert np.array_equal(rows.idx[a:b], nzc) and np.array_equal(rows.val[a:b], flat[i, nzc] / 255.0)
    with pytest.raises(ValueError):
        bad = tmp_path / "bad"
        bad.write_bytes(b"\x00\x00\x0d\x03" + b"\x00" * 12)
        R.idx.read_idx_ubyte(bad)




def test_chunksOf_groups_a_row_source_like_the_conduit(built):
    """chunkedAccum's `C.chunksOf n` (Conduit.hs:168-176): n rows per chunk, a shorter last chunk, nothing for an empty source;
    the source may yield single rows or blocks of rows."""
    import rp_tree_b200 as R
    X = np.arange(23 * 3, dtype=np.float64).reshape(23, 3)
    chunks = list(R.chunksOf(5, (row for row in X), 3))
    assert [c.shape[0] for c in chunks] == [5, 5, 5, 5, 3] and np.array_equal(np.concatenate(chunks), X)
    chunks = list(R.chunksOf(5, iter([X[:7], X[7:8], X[8:23]]), 3))
    assert [c.shape[0] for c in chunks] == [5, 5, 5, 5, 3] and np.array_equal(np.concatenate(chunks), X)
    assert list(R.chunksOf(4, iter([]), 3)) == []
    assert [c.shape[0] for c in R.chunksOf(23, iter([X]), 3)] == [23]


def test_insert_entry_points_reject_bad_state_without_a_gpu(built):
    """Argument / state errors of the insert entry points are decided on the host (no device needed): NULL handle."""
    import ctypes as C
    import rp_tree_b200 as R
    L = R.lib()
    assert L.rpf_insert_begin(None, 4, 3, 2) != 0
    assert L.rpf_insert_chunk(None, None, 0) != 0
    assert L.rpf_insert_end(None) != 0
