"""Second, independent restatement of the reference algorithm in plain Python (lists and floats, no numpy arithmetic).

TEST INFRASTRUCTURE ONLY -- imported by tests/ to cross-check the C oracle (oracle/rpt_oracle.c), which was written
separately from the same Haskell sources.  Two restatements that agree bit for bit on thresholds, margins, leaf order,
candidate lists, knn results and recall do not pin parity to the real reference (no GHC in this image: parity stays
"unpinned"), but they make a transcription slip in either one very unlikely.  Small inputs only (pure Python loops).

Every function cites the reference lines it follows (paths relative to /root/reference/src/Data/).  Python floats are
IEEE doubles, `a * b + c` is two roundings (CPython never fuses), `sorted` is stable -- the properties the reference's
`Double` arithmetic and `Data.Vector.Algorithms.Merge.sortBy` have.
"""
import math


def inner_sd(sv, x):
    """innerSD (RPTree/Internal.hs:369-382): right fold `xl * xr + go (i+1)`, stops at i >= nnz or i >= length x."""
    idx, val = sv
    m = min(len(idx), len(x))
    acc = 0.0
    for i in range(m - 1, -1, -1):          # innermost term of the right fold first
        acc = val[i] * x[idx[i]] + acc
    return acc


def metric_l2(u, v):
    """metricDDL2 (RPTree/Internal.hs:403-406): sqrt (sum (map (**2) (zipWith (-) u v))), sum = left fold from 0.
    `** 2` is written d * d here, as in the C oracle (DESIGN.md section 5 states the 1-ulp tolerance on libm pow)."""
    s = 0.0
    for a, b in zip(u, v):
        d = a - b
        s = s + d * d
    return math.sqrt(s)


def partition_at_median(r, ids, X):
    """partitionAtMedian (RPTree/Internal.hs:484-505) on the points `ids` (row numbers into X), in their given order."""
    n = len(ids)
    if n < 1:
        return None
    projs = sorted(((i, inner_sd(r, X[i])) for i in ids), key=lambda p: p[1])      # stable, Internal.hs:504,508-512
    xs = [p[0] for p in projs]
    inns = [p[1] for p in projs]
    nh = n // 2
    if n >= 3:
        mgl, mgr = inns[nh - 1], inns[nh + 1]
    elif n == 2:
        mgl, mgr = inns[0], inns[1]
    else:
        mgl = mgr = inns[0]
    return inns[nh], (mgl, mgr), xs[:nh], xs[nh:]


TIP_EMPTY = ("tip", [])


def insert(max_depth, min_leaf, rvs, tree, chunk, X):
    """insert (RPTree/Internal.hs:257-297): Bin case = partition the CHUNK, average the thresholds, combine the margins
    (Margin semigroup, Internal.hs:75-89: Max of the lows, Min of the highs); Tip case = chunk <> old contents, split once
    the leaf outgrows minLeaf; an empty piece reaching a Bin replaces the subtree by an empty Tip."""
    def loop(lev, tt, xs):
        if tt[0] == "bin":
            _, thr0, (lo0, hi0), tl0, tr0 = tt
            if lev >= max_depth:
                return tt
            p = partition_at_median(rvs[lev], xs, X)
            if p is None:
                return TIP_EMPTY
            thr, (lo, hi), ll, rr = p
            return ("bin", (thr0 + thr) / 2, (max(lo0, lo), min(hi0, hi)), loop(lev + 1, tl0, ll), loop(lev + 1, tr0, rr))
        xs2 = list(xs) + list(tt[1])
        if lev >= max_depth or len(xs2) <= min_leaf:
            return ("tip", xs2)
        p = partition_at_median(rvs[lev], xs2, X)
        if p is None:
            return TIP_EMPTY
        thr, mg, ll, rr = p
        return ("bin", thr, mg, loop(lev + 1, TIP_EMPTY, ll), loop(lev + 1, TIP_EMPTY, rr))
    return loop(0, tree, chunk)


def forest(X, hyperplanes, max_depth, min_leaf, chunk=None):
    """forestBatch (RPTree/Batch.hs:48-63: one chunk = the whole data set) or forest (RPTree/Conduit.hs:104-121:
    chunksOf chunk .| foldl insertMulti).  hyperplanes[t][level] = (idx list, val list)."""
    n = len(X)
    chunks = [list(range(n))] if chunk is None else [list(range(a, min(n, a + chunk))) for a in range(0, n, chunk)]
    trees = []
    for rvs in hyperplanes:
        t = TIP_EMPTY
        for c in chunks:
            t = insert(max_depth, min_leaf, rvs, t, c, X)
        trees.append(t)
    return trees


def points(tree):
    """leaves left to right (RPTree/Internal.hs:199-208)."""
    if tree[0] == "tip":
        return list(tree[1])
    return points(tree[3]) + points(tree[4])


def candidates(tree, rvs, q):
    """candidates (RPTree.hs:293-314): the four-way rule on (proj, thr, margins)."""
    def go(lev, tt):
        if tt[0] == "tip":
            return list(tt[1])
        _, thr, (mglo, mghi), lt, rt = tt
        proj = inner_sd(rvs[lev], q)
        dl, dr = abs(mglo - proj), abs(mghi - proj)
        if proj < thr and dl > dr:
            return go(lev + 1, lt) + go(lev + 1, rt)
        if proj < thr:
            return go(lev + 1, lt)
        if proj > thr and dl < dr:
            return go(lev + 1, lt) + go(lev + 1, rt)
        return go(lev + 1, rt)
    return go(0, tree)


def knn(trees, hyperplanes, X, k, q):
    """knn (RPTree.hs:168-176): all candidates of all trees in tree order, stable sort by distance, take k."""
    cs = []
    for t, rvs in zip(trees, hyperplanes):
        cs += candidates(t, rvs, q)
    ds = sorted(((metric_l2(X[i], q), i) for i in cs), key=lambda p: p[0])
    return ds[:k]


def knn_pq_distances(trees, hyperplanes, X, k, q):
    """knnPQ (RPTree.hs:181-194,224-227): one entry per distinct distance (Entry compares on the priority only), ascending.
    Which of several equidistant points survives is an internal of the `heaps` package: only the distances are stated."""
    return sorted({d for d, _ in knn(trees, hyperplanes, X, 10 ** 9, q)})[:k]


def recall_with(trees, hyperplanes, X, k, q):
    """recallWith / recallWith1 (RPTree.hs:259-285): per tree |candidates ∩ true top-k| / k, truth = stable sort of the
    tree's points (leaves left to right) by distance; mean over trees."""
    rs = []
    for t, rvs in zip(trees, hyperplanes):
        aa = set(candidates(t, rvs, q))
        dists = sorted(((i, metric_l2(X[i], q)) for i in points(t)), key=lambda p: p[1])
        kk = {i for i, _ in dists[:k]}
        rs.append(len(aa & kk) / k)
    return sum(rs) / len(trees)
