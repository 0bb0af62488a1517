"""Shared helpers for the parity tests: seeded data, oracle-vs-engine comparison."""
import numpy as np


def make_data(n, d, seed, kind="gauss"):
    rng = np.random.default_rng(seed)
    if kind == "gauss":
        return rng.normal(size=(n, d))
    if kind == "mixture":           # clustered, like the benchmark data
        nc = 16
        c = rng.normal(size=(nc, d))
        return c[rng.integers(0, nc, size=n)] + 0.25 * rng.normal(size=(n, d))
    if kind == "integer":           # many exact projection ties
        return rng.integers(-2, 3, size=(n, d)).astype(np.float64)
    if kind == "dupes":             # exact duplicate rows
        base = rng.normal(size=(max(n // 4, 1), d))
        return base[rng.integers(0, len(base), size=n)]
    if kind == "outlier":           # one huge point stretches every key range: nearly all keys share their leading bits
        X = rng.normal(size=(n, d))
        X[n // 3] *= 1e9
        return X
    raise ValueError(kind)


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def compare_tree(eng, orc_tree, check_order=True):
    """eng: RPForest.treeExport(t) dict; orc_tree: oracle Forest.export(t) dict.  Returns a list of mismatch strings."""
    bad = []
    for key in ("child", "depth", "seg_start", "seg_size"):
        if not np.array_equal(np.asarray(eng[key], np.int64), np.asarray(orc_tree[key], np.int64)):
            bad.append("topology." + key)
    if bad:
        return bad
    internal = orc_tree["child"] >= 0
    for key in ("thr", "mlo", "mhi"):
        a, b = bits(eng[key])[internal], bits(orc_tree[key])[internal]
        if not np.array_equal(a, b):
            w = np.flatnonzero(a != b)
            bad.append("%s: %d/%d nodes differ, first node %d (depth %d): %r vs %r" % (
                key, len(w), internal.sum(), np.flatnonzero(internal)[w[0]], orc_tree["depth"][np.flatnonzero(internal)[w[0]]],
                eng[key][internal][w[0]], orc_tree[key][internal][w[0]]))
    pe, po = eng["perm"], orc_tree["perm"]
    leaf = np.flatnonzero(orc_tree["child"] < 0)
    nset = nord = 0
    for g in leaf:
        s, z = orc_tree["seg_start"][g], orc_tree["seg_size"][g]
        a, b = pe[s:s + z], po[s:s + z]
        if not np.array_equal(np.sort(a), np.sort(b)):
            nset += 1
        elif check_order and not np.array_equal(a, b):
            nord += 1
    if nset:
        bad.append("leaf sets differ in %d/%d leaves" % (nset, len(leaf)))
    if nord:
        bad.append("leaf order differs in %d/%d leaves" % (nord, len(leaf)))
    return bad
