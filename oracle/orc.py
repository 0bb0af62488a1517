"""ctypes binding of the CPU parity oracle (oracle/rpt_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under rp-tree_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_f64p = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = [os.path.join(_HERE, f) for f in ("rpt_oracle.c", "rpt_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_mix64.restype = C.c_uint64
        L.orc_mix64.argtypes = [C.c_uint64]
        L.orc_next_word64.restype = C.c_uint64
        L.orc_next_double.restype = C.c_double
        L.orc_invnormcdf.restype = C.c_double
        L.orc_invnormcdf.argtypes = [C.c_double]
        L.orc_std_normal.restype = C.c_double
        L.orc_inner_sd.restype = C.c_double
        L.orc_inner_sd.argtypes = [C.c_int64, c_i32p, c_f64p, c_f64p, C.c_int64]
        L.orc_project_all.argtypes = [C.c_int64, c_i32p, c_f64p, c_f64p, C.c_int64, C.c_int64, c_f64p]
        L.orc_inner_ss.restype = C.c_double
        L.orc_inner_ss.argtypes = [C.c_int64, c_i32p, c_f64p, C.c_int64, c_i32p, c_f64p]
        L.orc_inner_dd.restype = C.c_double
        L.orc_inner_dd.argtypes = [c_f64p, c_f64p, C.c_int64]
        L.orc_metric_dd_l2.restype = C.c_double
        L.orc_metric_dd_l2.argtypes = [c_f64p, c_f64p, C.c_int64]
        for f in (L.orc_sum_sd, L.orc_diff_sd):
            f.restype = C.c_int64
            f.argtypes = [C.c_int64, c_i32p, c_f64p, c_f64p, C.c_int64, c_f64p]
        L.orc_rptree_cfg.argtypes = [C.c_int64, C.c_int64, C.c_int64, c_i64p, c_i64p, c_f64p]
        L.orc_gen_hyperplanes.restype = C.c_int64
        L.orc_gen_hyperplanes.argtypes = [C.c_uint64, C.c_int32, C.c_int32, C.c_double, C.c_int32, c_i64p, c_i32p, c_f64p]
        L.orc_forest_new.restype = C.c_void_p
        L.orc_forest_new.argtypes = [c_f64p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_i64p, c_i32p, c_f64p]
        L.orc_forest_new_mt.restype = C.c_void_p
        L.orc_forest_new_mt.argtypes = [c_f64p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_i64p, c_i32p, c_f64p, C.c_int32]
        L.orc_forest_new_chunked.restype = C.c_void_p
        L.orc_forest_new_chunked.argtypes = [c_f64p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, c_i64p, c_i32p, c_f64p]
        L.orc_forest_new_sparse.restype = C.c_void_p
        L.orc_forest_new_sparse.argtypes = [C.c_int64, C.c_int32, c_i64p, c_i32p, c_f64p, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                            c_i64p, c_i32p, c_f64p]
        L.orc_candidates_sq.restype = C.c_int64
        L.orc_candidates_sq.argtypes = [C.c_void_p, C.c_int32, C.c_int64, c_i32p, c_f64p, c_u32p, C.c_int64]
        L.orc_knn_sq.restype = C.c_int64
        L.orc_knn_sq.argtypes = [C.c_void_p, C.c_int64, c_i32p, c_f64p, C.c_int32, C.c_int32, c_f64p, c_u32p]
        L.orc_recall_sq.restype = C.c_double
        L.orc_recall_sq.argtypes = [C.c_void_p, C.c_int64, c_i32p, c_f64p, C.c_int32]
        L.orc_metric_ss_l2.restype = C.c_double
        L.orc_metric_ss_l2.argtypes = [C.c_int64, c_i32p, c_f64p, C.c_int64, c_i32p, c_f64p]
        L.orc_metric_sd_l2.restype = C.c_double
        L.orc_metric_sd_l2.argtypes = [C.c_int64, c_i32p, c_f64p, c_f64p, C.c_int64]
        L.orc_forest_free.argtypes = [C.c_void_p]
        L.orc_tree_export.restype = C.c_int64
        L.orc_tree_export.argtypes = [C.c_void_p, C.c_int32, c_i64p, c_i32p, c_f64p, c_f64p, c_f64p, c_i64p, c_i64p, c_u32p]
        L.orc_tree_size.restype = C.c_int64
        L.orc_tree_size.argtypes = [C.c_void_p, C.c_int32]
        L.orc_candidates.restype = C.c_int64
        L.orc_candidates.argtypes = [C.c_void_p, C.c_int32, c_f64p, c_u32p, C.c_int64]
        L.orc_knn.restype = C.c_int64
        L.orc_knn.argtypes = [C.c_void_p, c_f64p, C.c_int32, C.c_int32, c_f64p, c_u32p]
        L.orc_knn_h.restype = C.c_int64
        L.orc_knn_h.argtypes = [C.c_void_p, c_f64p, C.c_int32, c_f64p, c_u32p, C.c_int64]
        L.orc_knn_h_sq.restype = C.c_int64
        L.orc_knn_h_sq.argtypes = [C.c_void_p, C.c_int64, c_i32p, c_f64p, C.c_int32, c_f64p, c_u32p, C.c_int64]
        L.orc_recall.restype = C.c_double
        L.orc_recall.argtypes = [C.c_void_p, c_f64p, C.c_int32]
        L.orc_recall_shared.restype = C.c_double
        L.orc_recall_shared.argtypes = [C.c_void_p, c_f64p, C.c_int32]
        L.orc_brute_knn.argtypes = [c_f64p, C.c_int64, C.c_int32, c_f64p, C.c_int32, c_f64p, c_u32p]
        L.orc_set_use_pow.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t)


class SMGen(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("gamma", C.c_uint64)]


def mk_smgen(seed):
    g = SMGen()
    lib().orc_mk_smgen(C.c_uint64(seed), C.byref(g))
    return g


def next_word64(g):
    return lib().orc_next_word64(C.byref(g))


def next_double(g):
    return lib().orc_next_double(C.byref(g))


def inner_sd(idx, val, x):
    idx = np.ascontiguousarray(idx, np.int32)
    val = np.ascontiguousarray(val, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    return lib().orc_inner_sd(len(idx), _p(idx, c_i32p), _p(val, c_f64p), _p(x, c_f64p), len(x))


def project_all(hp, row, X):
    """innerSD of CSR row `row` of hp = (off, idx, val) against every row of X."""
    off, idx, val = hp
    X = np.ascontiguousarray(X, np.float64)
    a, b = int(off[row]), int(off[row + 1])
    ii = np.ascontiguousarray(idx[a:b] if b > a else np.zeros(1), np.int32)
    vv = np.ascontiguousarray(val[a:b] if b > a else np.zeros(1), np.float64)
    out = np.zeros(X.shape[0])
    lib().orc_project_all(b - a, _p(ii, c_i32p), _p(vv, c_f64p), _p(X, c_f64p), X.shape[0], X.shape[1], _p(out, c_f64p))
    return out


def inner_ss(i1, v1, i2, v2):
    i1 = np.ascontiguousarray(i1, np.int32); v1 = np.ascontiguousarray(v1, np.float64)
    i2 = np.ascontiguousarray(i2, np.int32); v2 = np.ascontiguousarray(v2, np.float64)
    return lib().orc_inner_ss(len(i1), _p(i1, c_i32p), _p(v1, c_f64p), len(i2), _p(i2, c_i32p), _p(v2, c_f64p))


def _bin_sd(fn, idx, val, x):
    idx = np.ascontiguousarray(idx, np.int32)
    val = np.ascontiguousarray(val, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros(len(idx) + len(x), np.float64)
    m = fn(len(idx), _p(idx, c_i32p), _p(val, c_f64p), _p(x, c_f64p), len(x), _p(out, c_f64p))
    return out[:m]


def sum_sd(idx, val, x):
    return _bin_sd(lib().orc_sum_sd, idx, val, x)


def diff_sd(idx, val, x):
    return _bin_sd(lib().orc_diff_sd, idx, val, x)


def metric_l2(u, v):
    u = np.ascontiguousarray(u, np.float64); v = np.ascontiguousarray(v, np.float64)
    return lib().orc_metric_dd_l2(_p(u, c_f64p), _p(v, c_f64p), len(u))


def rptree_cfg(minl, n, d):
    maxd = C.c_int64(); nchunk = C.c_int64(); pnz = C.c_double()
    lib().orc_rptree_cfg(minl, n, d, C.byref(maxd), C.byref(nchunk), C.byref(pnz))
    return maxd.value, nchunk.value, pnz.value


def gen_hyperplanes(seed, T, maxd, pnz, dim):
    """CSR (off[T*maxd+1] int64, idx int32, val float64) in the reference's draw order (Batch.hs:57-63)."""
    L = lib()
    off = np.zeros(T * maxd + 1, np.int64)
    nnz = L.orc_gen_hyperplanes(seed, T, maxd, pnz, dim, _p(off, c_i64p), None, None)
    idx = np.zeros(max(nnz, 1), np.int32)
    val = np.zeros(max(nnz, 1), np.float64)
    L.orc_gen_hyperplanes(seed, T, maxd, pnz, dim, _p(off, c_i64p), _p(idx, c_i32p), _p(val, c_f64p))
    return off, idx[:nnz].copy(), val[:nnz].copy()


class Forest:
    """Oracle forest built with forestBatch semantics (chunk=None) or forest/conduit semantics (chunk=int)."""

    def __init__(self, X, hp, T, maxd, minl, chunk=None, threads=1):
        self.X = np.ascontiguousarray(X, np.float64)
        self.n, self.d = self.X.shape
        off, idx, val = hp
        self.off = np.ascontiguousarray(off, np.int64)
        self.idx = np.ascontiguousarray(idx if len(idx) else np.zeros(1), np.int32)
        self.val = np.ascontiguousarray(val if len(val) else np.zeros(1), np.float64)
        self.T, self.maxd, self.minl = T, maxd, minl
        L = lib()
        if chunk is None and threads > 1:
            self.h = L.orc_forest_new_mt(_p(self.X, c_f64p), self.n, self.d, T, maxd, minl,
                                         _p(self.off, c_i64p), _p(self.idx, c_i32p), _p(self.val, c_f64p), threads)
        elif chunk is None:
            self.h = L.orc_forest_new(_p(self.X, c_f64p), self.n, self.d, T, maxd, minl,
                                      _p(self.off, c_i64p), _p(self.idx, c_i32p), _p(self.val, c_f64p))
        else:
            self.h = L.orc_forest_new_chunked(_p(self.X, c_f64p), self.n, self.d, T, maxd, minl, chunk,
                                              _p(self.off, c_i64p), _p(self.idx, c_i32p), _p(self.val, c_f64p))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_forest_free(self.h)
            self.h = None

    def tree_size(self, t):
        return lib().orc_tree_size(self.h, t)

    def export(self, t):
        L = lib()
        nn = L.orc_tree_export(self.h, t, None, None, None, None, None, None, None, None)
        child = np.zeros(nn, np.int64); depth = np.zeros(nn, np.int32)
        thr = np.zeros(nn); mlo = np.zeros(nn); mhi = np.zeros(nn)
        ss = np.zeros(nn, np.int64); sz = np.zeros(nn, np.int64)
        perm = np.zeros(max(self.n, 1), np.uint32)
        L.orc_tree_export(self.h, t, _p(child, c_i64p), _p(depth, c_i32p), _p(thr, c_f64p), _p(mlo, c_f64p), _p(mhi, c_f64p),
                          _p(ss, c_i64p), _p(sz, c_i64p), _p(perm, c_u32p))
        return dict(child=child, depth=depth, thr=thr, mlo=mlo, mhi=mhi, seg_start=ss, seg_size=sz,
                    perm=perm[: self.tree_size(t)])

    def candidates(self, t, q):
        q = np.ascontiguousarray(q, np.float64)
        L = lib()
        c = L.orc_candidates(self.h, t, _p(q, c_f64p), None, 0)
        ids = np.zeros(max(c, 1), np.uint32)
        L.orc_candidates(self.h, t, _p(q, c_f64p), _p(ids, c_u32p), c)
        return ids[:c]

    def knn(self, q, k, dedup=False):
        q = np.ascontiguousarray(q, np.float64)
        dist = np.zeros(k); ids = np.zeros(k, np.uint32)
        m = lib().orc_knn(self.h, _p(q, c_f64p), k, int(dedup), _p(dist, c_f64p), _p(ids, c_u32p))
        return dist[:m], ids[:m]

    def recall(self, q, k):
        q = np.ascontiguousarray(q, np.float64)
        return lib().orc_recall(self.h, _p(q, c_f64p), k)

    def recall_shared(self, q, k):
        """recall() with the brute-force distances evaluated once per query instead of once per tree (same value)."""
        q = np.ascontiguousarray(q, np.float64)
        return lib().orc_recall_shared(self.h, _p(q, c_f64p), k)

    def knn_h(self, q, k):
        """knnH: (distances, ids) in the reference's result order (not sorted by distance, not cut to k)."""
        cap = max(k, self.n) + 1
        dist = np.zeros(cap); ids = np.zeros(cap, np.uint32)
        if isinstance(q, tuple):
            nz, ii, vv = _sv(*q)
            m = lib().orc_knn_h_sq(self.h, nz, _p(ii, c_i32p), _p(vv, c_f64p), k, _p(dist, c_f64p), _p(ids, c_u32p), cap)
        else:
            q = np.ascontiguousarray(q, np.float64)
            m = lib().orc_knn_h(self.h, _p(q, c_f64p), k, _p(dist, c_f64p), _p(ids, c_u32p), cap)
        return dist[:m], ids[:m]


def _sv(idx, val):
    ii = np.ascontiguousarray(idx if len(idx) else np.zeros(1), np.int32)
    vv = np.ascontiguousarray(val if len(val) else np.zeros(1), np.float64)
    return len(idx), ii, vv


def metric_ss_l2(i1, v1, i2, v2):
    n1, a, b = _sv(i1, v1); n2, c, d = _sv(i2, v2)
    return lib().orc_metric_ss_l2(n1, _p(a, c_i32p), _p(b, c_f64p), n2, _p(c, c_i32p), _p(d, c_f64p))


def metric_sd_l2(idx, val, x):
    n1, a, b = _sv(idx, val)
    x = np.ascontiguousarray(x, np.float64)
    return lib().orc_metric_sd_l2(n1, _p(a, c_i32p), _p(b, c_f64p), _p(x, c_f64p), len(x))


class SparseForest(Forest):
    """Oracle forest over SVector data points: csr = (off int64[n+1], idx int32, val float64), indices ascending per
    row.  Queries: a dense vector (metricSDL2) or an (idx, val) pair (metricSSL2)."""

    def __init__(self, csr, d, hp, T, maxd, minl, chunk=None):
        off, idx, val = csr
        self.sp_off = np.ascontiguousarray(off, np.int64)
        self.sp_idx = np.ascontiguousarray(idx if len(idx) else np.zeros(1), np.int32)
        self.sp_val = np.ascontiguousarray(val if len(val) else np.zeros(1), np.float64)
        self.n, self.d = len(off) - 1, d
        hoff, hidx, hval = hp
        self.off = np.ascontiguousarray(hoff, np.int64)
        self.idx = np.ascontiguousarray(hidx if len(hidx) else np.zeros(1), np.int32)
        self.val = np.ascontiguousarray(hval if len(hval) else np.zeros(1), np.float64)
        self.T, self.maxd, self.minl = T, maxd, minl
        self.h = lib().orc_forest_new_sparse(self.n, d, _p(self.sp_off, c_i64p), _p(self.sp_idx, c_i32p), _p(self.sp_val, c_f64p),
                                             T, maxd, minl, chunk if chunk else 0,
                                             _p(self.off, c_i64p), _p(self.idx, c_i32p), _p(self.val, c_f64p))

    @staticmethod
    def _is_sparse(q):
        return isinstance(q, tuple)

    def candidates(self, t, q):
        if not self._is_sparse(q):
            return Forest.candidates(self, t, q)
        nz, ii, vv = _sv(*q)
        L = lib()
        c = L.orc_candidates_sq(self.h, t, nz, _p(ii, c_i32p), _p(vv, c_f64p), None, 0)
        ids = np.zeros(max(c, 1), np.uint32)
        L.orc_candidates_sq(self.h, t, nz, _p(ii, c_i32p), _p(vv, c_f64p), _p(ids, c_u32p), c)
        return ids[:c]

    def knn(self, q, k, dedup=False):
        if not self._is_sparse(q):
            return Forest.knn(self, q, k, dedup)
        nz, ii, vv = _sv(*q)
        dist = np.zeros(k); ids = np.zeros(k, np.uint32)
        m = lib().orc_knn_sq(self.h, nz, _p(ii, c_i32p), _p(vv, c_f64p), k, int(dedup), _p(dist, c_f64p), _p(ids, c_u32p))
        return dist[:m], ids[:m]

    def recall(self, q, k):
        if not self._is_sparse(q):
            return Forest.recall(self, q, k)
        nz, ii, vv = _sv(*q)
        return lib().orc_recall_sq(self.h, nz, _p(ii, c_i32p), _p(vv, c_f64p), k)


def slice_hyperplanes(hp, maxd, t_first, t_local):
    """CSR rows of trees [t_first, t_first + t_local) of a forest-wide hyperplane set, re-based to offset 0."""
    off, idx, val = hp
    a, b = t_first * maxd, (t_first + t_local) * maxd
    lo, hi = int(off[a]), int(off[b])
    return (np.ascontiguousarray(off[a:b + 1] - off[a], np.int64), np.ascontiguousarray(idx[lo:hi], np.int32),
            np.ascontiguousarray(val[lo:hi], np.float64))


def brute_knn(X, q, k):
    X = np.ascontiguousarray(X, np.float64); q = np.ascontiguousarray(q, np.float64)
    dist = np.zeros(k); ids = np.zeros(k, np.uint32)
    lib().orc_brute_knn(_p(X, c_f64p), X.shape[0], X.shape[1], _p(q, c_f64p), k, _p(dist, c_f64p), _p(ids, c_u32p))
    return dist, ids
