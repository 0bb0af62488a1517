"""Host-side mirror of the reference's public interface for the hot path (Data.RPTree, src/Data/RPTree.hs:50-113).

Same names and argument meaning as the Haskell functions; every call goes through the C ABI of
librpforest.so (include/rpforest.h) -- exactly what the Haskell FFI shim would bind.
Data points are rows of a float64 matrix (one DVector per row); the `Embed` payload is the row number.

    forestBatch seed maxd minl ntrees pnz dim xs      src/Data/RPTree/Batch.hs:48-63
    treeBatch   seed maxd minl pnz dim xs             src/Data/RPTree/Batch.hs:29-41
    forest / tree (chunked conduit versions)          src/Data/RPTree/Conduit.hs:58-121
    rpTreeCfg minl n d                                src/Data/RPTree/Conduit.hs:132-141
    knn / knnPQ / knnH distf k tts q                  src/Data/RPTree.hs:168-217
    candidates tree q                                 src/Data/RPTree.hs:293-314
    recallWith distf tts k q                          src/Data/RPTree.hs:259-268
    treeSize / leafSizes / levels / points            src/Data/RPTree.hs:351-367, Internal.hs:199-208
"""
import ctypes as C
from collections import namedtuple

import numpy as np

from ._lib import lib, RPForestError, i64p, i32p, u32p, f64p, H

RPTreeConfig = namedtuple("RPTreeConfig", ["fpMaxTreeDepth", "fpDataChunkSize", "fpProjNzDensity"])  # Conduit.hs:123-128


class _MetricL2:
    """Stand-in for `metricL2` (Internal.hs:318,337-339): the only distance the engine implements."""

    def __repr__(self):
        return "metricL2"


metricL2 = _MetricL2()


def _p(a, t):
    return a.ctypes.data_as(t)


class SparseRows:
    """A batch of SVectors (Internal.hs:92-97) as CSR rows: off int64[n+1], idx int32 (strictly ascending per row), val."""

    def __init__(self, off, idx, val, d):
        self.off = np.ascontiguousarray(off, np.int64)
        self.idx = np.ascontiguousarray(idx if len(idx) else np.zeros(1), np.int32)
        self.val = np.ascontiguousarray(val if len(val) else np.zeros(1), np.float64)
        self.n, self.d = len(self.off) - 1, int(d)

    @classmethod
    def fromDense(cls, M):
        """fromListSv-style: keep the nonzero components of every row of M."""
        M = np.asarray(M, np.float64)
        if M.ndim == 1:
            M = M[None, :]
        r, c = np.nonzero(M)
        off = np.zeros(M.shape[0] + 1, np.int64)
        np.add.at(off, r + 1, 1)
        return cls(np.cumsum(off), c.astype(np.int32), M[r, c], M.shape[1])

    def densify(self):
        """(dense n x d image, last stored component per row) through the library's host helper."""
        Q = np.zeros((self.n, self.d)); last = np.zeros(max(self.n, 1), np.int32)
        rc = lib().rpf_densify_rows(self.n, self.d, _p(self.off, i64p), _p(self.idx, i32p), _p(self.val, f64p), _p(Q, f64p), _p(last, i32p))
        if rc != 0:
            raise RPForestError("rpf_densify_rows: malformed CSR rows (rc=%d)" % rc)
        return Q, last[: self.n]


def _as_q(q, d):
    """-> (dense image nq x d, single?, q_last or None)"""
    if isinstance(q, SparseRows):
        if q.d != d:
            raise ValueError("query must have dimension %d" % d)
        Q, last = q.densify()
        return Q, False, last
    Q, single = _as_qd(q, d)
    return Q, single, None


def _as_qd(q, d):
    Q = np.ascontiguousarray(q, dtype=np.float64)
    single = Q.ndim == 1
    if single:
        Q = Q[None, :]
    if Q.ndim != 2 or Q.shape[1] != d:
        raise ValueError("query must have dimension %d" % d)
    return Q, single


def rpTreeCfg(minl, n, d):
    """rpTreeCfg (Conduit.hs:132-141): natural defaults for maxDepth, chunk size, projection density."""
    maxd = C.c_int64(); chunk = C.c_int64(); pnz = C.c_double()
    lib().rpf_rptree_cfg(minl, n, d, C.byref(maxd), C.byref(chunk), C.byref(pnz))
    return RPTreeConfig(maxd.value, chunk.value, pnz.value)


def sampleHyperplanes(seed, ntrees, maxd, pnz, dim):
    """`sample seed (replicateM ntrees (V.replicateM maxd (sparse pnz dim stdNormal)))` (Batch.hs:59-61) as CSR.

    Host-only (no GPU).  The SplitMix64 core is verified; the normal sampler is an unverified restatement.
    """
    L = lib()
    off = np.zeros(ntrees * maxd + 1, np.int64)
    nnz = L.rpf_sample_hyperplanes(seed, ntrees, maxd, pnz, dim, _p(off, i64p), None, None)
    if nnz < 0:
        raise RPForestError("rpf_sample_hyperplanes: bad arguments")
    idx = np.zeros(max(nnz, 1), np.int32); val = np.zeros(max(nnz, 1), np.float64)
    L.rpf_sample_hyperplanes(seed, ntrees, maxd, pnz, dim, _p(off, i64p), _p(idx, i32p), _p(val, f64p))
    return off, idx[:nnz].copy(), val[:nnz].copy()


def topologyPlan(n, maxd, minl, chunk=None):
    """The data-independent tree shape for (n, maxDepth, minLeaf[, chunk size]): BFS arrays child/depth/seg_start/
    seg_size.  With `chunk` (the streaming `forest`/`tree`, Conduit.hs:58-121) also `points_lost`."""
    L = lib()
    if chunk is None:
        nn = L.rpf_topology_plan(n, maxd, minl, None, None, None, None)
    else:
        nn = L.rpf_topology_plan_chunked(n, maxd, minl, chunk, None, None, None, None, None)
    if nn < 0:
        raise RPForestError("rpf_topology_plan: bad or unsupported arguments (rc=%d)" % nn)
    child = np.zeros(nn, np.int64); depth = np.zeros(nn, np.int32); ss = np.zeros(nn, np.int64); sz = np.zeros(nn, np.int64)
    if chunk is None:
        L.rpf_topology_plan(n, maxd, minl, _p(child, i64p), _p(depth, i32p), _p(ss, i64p), _p(sz, i64p))
        return dict(child=child, depth=depth, seg_start=ss, seg_size=sz)
    lost = C.c_int64()
    L.rpf_topology_plan_chunked(n, maxd, minl, chunk, _p(child, i64p), _p(depth, i32p), _p(ss, i64p), _p(sz, i64p), C.byref(lost))
    return dict(child=child, depth=depth, seg_start=ss, seg_size=sz, points_lost=lost.value)


def slice_hyperplanes(hp, maxd, t_first, t_local):
    """CSR rows of trees [t_first, t_first+t_local) -- what one GPU rank passes to rpf_set_hyperplanes."""
    off, idx, val = hp
    r0, r1 = t_first * maxd, (t_first + t_local) * maxd
    lo, hi = off[r0], off[r1]
    return (off[r0:r1 + 1] - lo).astype(np.int64), idx[lo:hi].copy(), val[lo:hi].copy()


class RPForest:
    """RPForest Double (V.Vector (Embed DVector Double Int)) held on one B200 (Internal.hs:182).

    Wraps one `rpf_handle`.  `t_first` is the global index of this shard's first tree (multi-GPU sharding).
    """

    def __init__(self, device=0, devices=None):
        """device: one CUDA device (rpf_create).  devices=[g0, g1, ...]: ONE handle over several GPUs of this process
        (rpf_create_multi): trees sharded in contiguous blocks, data replicated, NCCL exchanges inside the engine -- every
        method below keeps its single-GPU meaning."""
        self._L = lib()
        self._h = H()
        if devices is not None and len(devices) > 1:
            ids = np.ascontiguousarray(devices, np.int32)
            rc = self._L.rpf_create_multi(C.byref(self._h), _p(ids, i32p), len(ids))
            device = int(ids[0])
        else:
            if devices is not None:
                device = int(devices[0])
            rc = self._L.rpf_create(C.byref(self._h), device)
        if rc != 0:
            self._h = None
            raise RPForestError("rpf_create failed (rc=%d): no usable CUDA device %r -- the engine has no CPU fallback" % (
                rc, devices if devices is not None else device))
        self.device = device
        self.n = 0; self.d = 0; self.ntrees = 0; self.maxDepth = 0; self.minLeaf = 0
        self.t_first = 0; self.ntrees_total = 0
        self._topo = None

    # -- plumbing
    def _ck(self, rc, what):
        if rc != 0:
            raise RPForestError("%s failed (rc=%d): %s" % (what, rc, self._L.rpf_last_error(self._h).decode()))

    def close(self):
        if getattr(self, "_h", None):
            self._L.rpf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- multi-GPU
    @staticmethod
    def commUniqueId():
        """128-byte NCCL id (rpf_comm_unique_id): create on one rank, hand to every rank's commInitRank."""
        buf = C.create_string_buffer(128)
        rc = lib().rpf_comm_unique_id(buf)
        if rc != 0:
            raise RPForestError("rpf_comm_unique_id failed (rc=%d): NCCL not available" % rc)
        return buf.raw

    def commInitRank(self, world, rank, uid):
        """One process per GPU: make this handle rank `rank` of a `world`-rank tree-sharded forest (before setPoints).
        From then on setPoints / buildFromHost / knnBatch / recallSumBatch are collective and forest-wide."""
        buf = C.create_string_buffer(bytes(uid), 128)
        self._ck(self._L.rpf_comm_init_rank(self._h, world, rank, buf), "rpf_comm_init_rank")

    def numGpus(self):
        return int(self._L.rpf_num_gpus(self._h))

    # -- build
    def setPoints(self, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("X must be n x d")
        self._borrowed_points = None          # dist.buildFromHostSharded's cache: the handle owns its points again
        self._ck(self._L.rpf_set_points(self._h, _p(X, f64p), X.shape[0], X.shape[1]), "rpf_set_points")
        self.n, self.d = X.shape

    def setPointsSparse(self, rows):
        """Data points as SVectors (Embed SVector Double x): rows is a SparseRows."""
        self._borrowed_points = None
        self._ck(self._L.rpf_set_points_sparse(self._h, rows.n, rows.d, _p(rows.off, i64p), _p(rows.idx, i32p), _p(rows.val, f64p)),
                 "rpf_set_points_sparse")
        self.n, self.d = rows.n, rows.d

    def pointsAreSparse(self):
        return bool(self._L.rpf_points_are_sparse(self._h))

    def setPointsRaw(self, addr, n, d):
        """rpf_set_points with a raw host address for row 0 (rank of a communicator: only this rank's rows
        [r*per, (r+1)*per), per = ceil(n/world), are read, so `addr` may be a virtual base whose other rows are unmapped)."""
        self._borrowed_points = None
        self._ck(self._L.rpf_set_points(self._h, C.cast(C.c_void_p(addr), f64p), n, d), "rpf_set_points")
        self.n, self.d = n, d

    def setPointsDevice(self, ptr, n, d):
        self._ck(self._L.rpf_set_points_device(self._h, C.c_void_p(ptr), n, d), "rpf_set_points_device")
        self.n, self.d = n, d

    def setHyperplanes(self, hp, ntrees, maxd):
        off, idx, val = hp
        off = np.ascontiguousarray(off, np.int64)
        idx = np.ascontiguousarray(idx if len(idx) else np.zeros(1), np.int32)
        val = np.ascontiguousarray(val if len(val) else np.zeros(1), np.float64)
        self._ck(self._L.rpf_set_hyperplanes(self._h, ntrees, maxd, _p(off, i64p), _p(idx, i32p), _p(val, f64p)), "rpf_set_hyperplanes")
        self.ntrees = ntrees
        self._hp_depth = maxd

    def genHyperplanes(self, seed, ntrees_total, maxd, pnz, dim, t_first=0, t_local=None):
        t_local = ntrees_total - t_first if t_local is None else t_local
        self._ck(self._L.rpf_gen_hyperplanes(self._h, seed, ntrees_total, maxd, pnz, dim, t_first, t_local), "rpf_gen_hyperplanes")
        self.ntrees = t_local
        self._hp_depth = maxd

    def hyperplanes(self):
        nnz = self._L.rpf_hyperplane_nnz(self._h)
        rows = self.ntrees * self._hp_depth
        off = np.zeros(rows + 1, np.int64); idx = np.zeros(max(nnz, 1), np.int32); val = np.zeros(max(nnz, 1), np.float64)
        self._ck(self._L.rpf_get_hyperplanes(self._h, _p(off, i64p), _p(idx, i32p), _p(val, f64p)), "rpf_get_hyperplanes")
        return off, idx[:nnz], val[:nnz]

    def buildFromHost(self, X, maxd, minl):
        """setPoints + build in one call, the upload overlapped with the projection (rpf_build_from_host)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("X must be n x d")
        self._borrowed_points = None
        self._ck(self._L.rpf_build_from_host(self._h, _p(X, f64p), X.shape[0], X.shape[1], maxd, minl), "rpf_build_from_host")
        self.n, self.d = X.shape
        self.maxDepth, self.minLeaf = maxd, minl
        self._topo = None

    def build(self, maxd, minl, chunk=None):
        if chunk is None:
            self._ck(self._L.rpf_build(self._h, maxd, minl), "rpf_build")
        else:
            self._ck(self._L.rpf_build_chunked(self._h, maxd, minl, chunk), "rpf_build_chunked")
        self.maxDepth, self.minLeaf = maxd, minl
        self._topo = None

    # -- incremental build: the fold of Conduit.hs:157-176, one insertMulti per call
    def insertBegin(self, d, maxd, minl):
        """Every tree = `Tip () mempty` (rpf_insert_begin); the hyperplanes must be set."""
        self._borrowed_points = None
        self._ck(self._L.rpf_insert_begin(self._h, d, maxd, minl), "rpf_insert_begin")
        self.n, self.d = 0, d
        self.maxDepth, self.minLeaf = maxd, minl
        self._topo = None

    def insertChunk(self, Xc):
        """insertMulti (Internal.hs:243-255) of one chunk of rows into every tree; the forest is queryable afterwards."""
        Xc = np.ascontiguousarray(Xc, dtype=np.float64)
        if Xc.ndim != 2 or (self.d and Xc.shape[0] and Xc.shape[1] != self.d):
            raise ValueError("chunk must be m x %d" % self.d)
        self._ck(self._L.rpf_insert_chunk(self._h, _p(Xc, f64p), Xc.shape[0]), "rpf_insert_chunk")
        self.n += Xc.shape[0]
        self._topo = None

    def insertEnd(self):
        self._ck(self._L.rpf_insert_end(self._h), "rpf_insert_end")

    # -- structure
    def topology(self):
        if self._topo is None:
            nn = self._L.rpf_num_nodes(self._h)
            child = np.zeros(nn, np.int64); depth = np.zeros(nn, np.int32); ss = np.zeros(nn, np.int64); sz = np.zeros(nn, np.int64)
            self._ck(self._L.rpf_topology(self._h, _p(child, i64p), _p(depth, i32p), _p(ss, i64p), _p(sz, i64p)), "rpf_topology")
            self._topo = dict(child=child, depth=depth, seg_start=ss, seg_size=sz)
        return self._topo

    def treeExport(self, t):
        """Flat image of tree t: what the shim turns back into Bin/Tip (Internal.hs:139-148)."""
        tp = self.topology()
        nn = len(tp["child"])
        thr = np.zeros(nn); mlo = np.zeros(nn); mhi = np.zeros(nn); perm = np.zeros(max(self.n, 1), np.uint32)
        self._ck(self._L.rpf_tree_export(self._h, t, _p(thr, f64p), _p(mlo, f64p), _p(mhi, f64p), _p(perm, u32p)), "rpf_tree_export")
        out = dict(tp)
        out.update(thr=thr, mlo=mlo, mhi=mhi, perm=perm[: self.n])
        return out

    def forestExport(self, out=None):
        """Flat image of every tree in one call: thr/mlo/mhi [T][nodes], perm [T][n].  `out` may hold preallocated
        (ideally page-locked) arrays of those shapes under the same keys; they are filled in place."""
        tp = self.topology()
        nn, T, n = len(tp["child"]), self.ntrees, self.n
        out = {} if out is None else out
        for key in ("thr", "mlo", "mhi"):
            if key not in out:
                out[key] = np.zeros((T, nn))
            assert out[key].shape == (T, nn) and out[key].dtype == np.float64 and out[key].flags.c_contiguous
        if "perm" not in out:
            out["perm"] = np.zeros((T, max(n, 1)), np.uint32)[:, :n] if n == 0 else np.zeros((T, n), np.uint32)
        assert out["perm"].shape == (T, n) and out["perm"].dtype == np.uint32
        self._ck(self._L.rpf_forest_export(self._h, _p(out["thr"], f64p), _p(out["mlo"], f64p), _p(out["mhi"], f64p),
                                           _p(out["perm"], u32p) if n else None), "rpf_forest_export")
        return out

    def setExportSink(self, out):
        """Register `out` (dict of thr/mlo/mhi [T][nodes] float64 and perm [T][n] uint32, ideally page-locked; None clears)
        as the buffers buildFromHost streams the forest into while it builds; forestExport(out) then only waits."""
        if out is None:
            self._ck(self._L.rpf_set_export_sink(self._h, None, None, None, None), "rpf_set_export_sink")
            self._sink = None
            return
        self._sink = out                     # keeps the arrays alive
        self._ck(self._L.rpf_set_export_sink(self._h, _p(out["thr"], f64p), _p(out["mlo"], f64p), _p(out["mhi"], f64p),
                                             _p(out["perm"], u32p)), "rpf_set_export_sink")

    def save(self, path, with_points=True):
        """Checkpoint of the built forest (counterpart of serialiseRPForest, Internal.hs:185-190)."""
        self._ck(self._L.rpf_forest_save(self._h, str(path).encode(), int(with_points)), "rpf_forest_save")

    def load(self, path):
        """Restore a checkpoint into this handle (counterpart of deserialiseRPForest, Internal.hs:192-196)."""
        self._borrowed_points = None
        self._ck(self._L.rpf_forest_load(self._h, str(path).encode()), "rpf_forest_load")
        self._topo = None
        self.ntrees = int(self._L.rpf_num_trees(self._h))
        tp = self.topology()
        self.maxDepth = int(tp["depth"].max()) if len(tp["depth"]) else 0
        n = C.c_int64(); d = C.c_int32()
        self._L.rpf_points_shape(self._h, C.byref(n), C.byref(d))
        self.n, self.d = n.value, d.value
        self._hp_depth = int(self._L.rpf_hyperplane_depth(self._h))
        self.ntrees_total = self.ntrees

    def leafOrderExact(self):
        return bool(self._L.rpf_leaf_order_exact(self._h))

    def pointsLost(self):
        """Points the reference's streaming insert drops (empty piece reaching a Bin, Internal.hs:279); 0 for batch builds."""
        return int(self._L.rpf_points_lost(self._h))

    # -- queries (batched: Q is nq x d)
    def candidatesBatch(self, Q, t=-1):
        Q, _, _ = _as_q(Q, self.d)
        nq = Q.shape[0]
        off = np.zeros(nq + 1, np.int64)
        self._ck(self._L.rpf_candidates_count(self._h, _p(Q, f64p), nq, t, _p(off, i64p)), "rpf_candidates_count")
        ids = np.zeros(max(int(off[-1]), 1), np.uint32)
        self._ck(self._L.rpf_candidates(self._h, _p(Q, f64p), nq, t, _p(off, i64p), _p(ids, u32p)), "rpf_candidates")
        return off, ids[: off[-1]]

    def knnBatch(self, Q, k, dedup=False, out=None):
        """knn / knnPQ for a batch: (dist nq x k, ids nq x k, count nq).  `out` may hold preallocated (ideally page-locked)
        arrays of those shapes.  On a multi-GPU handle / communicator rank the result covers the whole forest."""
        Q, _, ql = _as_q(Q, self.d)
        nq = Q.shape[0]
        if out is None:
            dist = np.zeros((nq, k)); ids = np.zeros((nq, k), np.uint32); cnt = np.zeros(nq, np.int32)
        else:
            dist, ids, cnt = out
            assert dist.shape == (nq, k) and dist.dtype == np.float64 and ids.shape == (nq, k) and ids.dtype == np.uint32
            assert cnt.shape == (nq,) and cnt.dtype == np.int32
        if ql is None:
            self._ck(self._L.rpf_knn(self._h, _p(Q, f64p), nq, k, int(dedup), _p(dist, f64p), _p(ids, u32p), _p(cnt, i32p)), "rpf_knn")
        else:
            self._ck(self._L.rpf_knn_s(self._h, _p(Q, f64p), _p(ql, i32p), nq, k, int(dedup), _p(dist, f64p), _p(ids, u32p), _p(cnt, i32p)), "rpf_knn_s")
        return dist, ids, cnt

    def knnHBatch(self, Q, k):
        """knnH for a batch: (dist, ids) nq x cap and cnt; row i holds cnt[i] results in the reference's order."""
        Q, _, ql = _as_q(Q, self.d)
        nq = Q.shape[0]
        cap = int(self._L.rpf_knn_h_capacity(self._h, k))
        if cap < 0:
            raise RPForestError("rpf_knn_h_capacity: forest not built")
        dist = np.zeros((nq, cap)); ids = np.zeros((nq, cap), np.uint32); cnt = np.zeros(nq, np.int32)
        self._ck(self._L.rpf_knn_h(self._h, _p(Q, f64p), _p(ql, i32p) if ql is not None else None, nq, k, cap,
                                   _p(dist, f64p), _p(ids, u32p), _p(cnt, i32p)), "rpf_knn_h")
        return dist, ids, cnt

    def recallSumBatch(self, Q, k):
        Q, _, ql = _as_q(Q, self.d)
        nq = Q.shape[0]
        r = np.zeros(nq)
        if ql is None:
            self._ck(self._L.rpf_recall(self._h, _p(Q, f64p), nq, k, _p(r, f64p)), "rpf_recall")
        else:
            self._ck(self._L.rpf_recall_s(self._h, _p(Q, f64p), _p(ql, i32p), nq, k, _p(r, f64p)), "rpf_recall_s")
        return r

    def bruteKnnBatch(self, Q, k):
        Q, _, ql = _as_q(Q, self.d)
        nq = Q.shape[0]
        dist = np.zeros((nq, k)); ids = np.zeros((nq, k), np.uint32)
        if ql is None:
            self._ck(self._L.rpf_brute_knn(self._h, _p(Q, f64p), nq, k, _p(dist, f64p), _p(ids, u32p)), "rpf_brute_knn")
        else:
            self._ck(self._L.rpf_brute_knn_s(self._h, _p(Q, f64p), _p(ql, i32p), nq, k, _p(dist, f64p), _p(ids, u32p)), "rpf_brute_knn_s")
        return dist, ids

    def mergeTopk(self, dist, ids, cnt, dedup=False):
        """dist/ids: G x nq x k, cnt: G x nq (rank-major) -> merged nq x k."""
        dist = np.ascontiguousarray(dist, np.float64); ids = np.ascontiguousarray(ids, np.uint32); cnt = np.ascontiguousarray(cnt, np.int32)
        G, nq, k = dist.shape
        od = np.zeros((nq, k)); oi = np.zeros((nq, k), np.uint32); oc = np.zeros(nq, np.int32)
        self._ck(self._L.rpf_merge_topk(self._h, G, nq, k, int(dedup), _p(dist, f64p), _p(ids, u32p), _p(cnt, i32p),
                                        _p(od, f64p), _p(oi, u32p), _p(oc, i32p)), "rpf_merge_topk")
        return od, oi, oc

    def knnBatchDevice(self, Q, k, dist_ptr, ids_ptr, cnt_ptr, dedup=False):
        """knn whose (dist nq x k f64, ids nq x k u32, count nq i32) land in DEVICE buffers of the caller (raw pointers on
        this forest's device); complete when the call returns."""
        Q, _, ql = _as_q(Q, self.d)
        self._ck(self._L.rpf_knn_dev(self._h, _p(Q, f64p), _p(ql, i32p) if ql is not None else None, Q.shape[0], k, int(dedup),
                                     C.c_void_p(dist_ptr), C.c_void_p(ids_ptr), C.c_void_p(cnt_ptr)), "rpf_knn_dev")

    def mergeTopkDevice(self, G, nq, k, dist_ptr, ids_ptr, cnt_ptr, dedup=False):
        """Merge of gathered rank-major DEVICE lists (G x nq x k, G x nq x k, G x nq) -> merged host arrays."""
        od = np.zeros((nq, k)); oi = np.zeros((nq, k), np.uint32); oc = np.zeros(nq, np.int32)
        self._ck(self._L.rpf_merge_topk_dev(self._h, G, nq, k, int(dedup), C.c_void_p(dist_ptr), C.c_void_p(ids_ptr), C.c_void_p(cnt_ptr),
                                            _p(od, f64p), _p(oi, u32p), _p(oc, i32p)), "rpf_merge_topk_dev")
        return od, oi, oc

    # -- measurement
    def lastDeviceMs(self):
        return self._L.rpf_last_device_ms(self._h)

    def setProfiling(self, on):
        self._ck(self._L.rpf_set_profiling(self._h, int(on)), "rpf_set_profiling")

    def profile(self):
        n = self._L.rpf_get_profile(self._h, None, None, 0)
        ms = np.zeros(n); la = np.zeros(n, np.int64)
        self._L.rpf_get_profile(self._h, _p(ms, f64p), _p(la, i64p), n)
        return {self._L.rpf_phase_name(i).decode(): (float(ms[i]), int(la[i])) for i in range(n)}

    def launchCount(self):
        return self._L.rpf_launch_count(self._h)

    def setOption(self, name, value):
        self._ck(self._L.rpf_set_option(self._h, name.encode(), int(value)), "rpf_set_option")

    def setBottomCap(self, cap):
        self._ck(self._L.rpf_set_bottom_cap(self._h, cap), "rpf_set_bottom_cap")


# ---------------------------------------------------------------------------------------------------------
# the reference's function names
# ---------------------------------------------------------------------------------------------------------
def _set_points(f, xs, dim):
    if isinstance(xs, SparseRows):
        if xs.d != dim:
            raise ValueError("dataset must have dimension %d" % dim)
        f.setPointsSparse(xs)
        return
    xs = np.ascontiguousarray(xs, dtype=np.float64)
    if xs.ndim != 2 or xs.shape[1] != dim:
        raise ValueError("dataset must be n x %d" % dim)
    f.setPoints(xs)


def forestBatch(seed, maxd, minl, ntrees, pnz, dim, xs, *, hyperplanes=None, device=0, t_first=0, t_local=None, bottom_cap=None, options=None):
    """forestBatch (Batch.hs:48-63).  `hyperplanes` = CSR (off, idx, val) drawn by the Haskell host overrides the
    built-in sampler (bit-exact parity path).  t_first/t_local shard the trees for multi-GPU runs.
    xs: n x dim float64 matrix (DVector points) or SparseRows (SVector points)."""
    t_local = ntrees - t_first if t_local is None else t_local
    f = RPForest(device)
    if bottom_cap is not None:
        f.setBottomCap(bottom_cap)
    for name, value in (options or {}).items():
        f.setOption(name, value)
    dense = not isinstance(xs, SparseRows)
    if dense:
        xs = np.ascontiguousarray(xs, dtype=np.float64)
        if xs.ndim != 2 or xs.shape[1] != dim:
            raise ValueError("dataset must be n x %d" % dim)
        f.d = dim                    # component range check of the hyperplanes
    else:
        _set_points(f, xs, dim)
    if hyperplanes is not None:
        hp = hyperplanes if (t_first == 0 and t_local == ntrees) else slice_hyperplanes(hyperplanes, maxd, t_first, t_local)
        f.setHyperplanes(hp, t_local, maxd)
    else:
        f.genHyperplanes(seed, ntrees, maxd, pnz, dim, t_first, t_local)
    f.t_first, f.ntrees_total = t_first, ntrees
    if dense:
        f.buildFromHost(xs, maxd, minl)      # upload overlapped with the projection
    else:
        f.build(maxd, minl)
    return f


def treeBatch(seed, maxd, minl, pnz, dim, xs, **kw):
    """treeBatch (Batch.hs:29-41): a forest of one tree."""
    return forestBatch(seed, maxd, minl, 1, pnz, dim, xs, **kw)


def forest(seed, maxd, minl, ntrees, chunksize, pnz, dim, xs, *, hyperplanes=None, device=0, t_first=0, t_local=None,
           bottom_cap=None, options=None):
    """forest (Conduit.hs:104-121): the rows of xs arrive in chunks of `chunksize` (the conduit's `chunksOf`), each chunk
    updates every tree (insertMulti, Internal.hs:243-255).  chunksize >= n is forestBatch."""
    t_local = ntrees - t_first if t_local is None else t_local
    f = RPForest(device)
    if bottom_cap is not None:
        f.setBottomCap(bottom_cap)
    for name, value in (options or {}).items():
        f.setOption(name, value)
    source = not isinstance(xs, (np.ndarray, SparseRows, list, tuple))     # a conduit-like source: rows arrive one at a time
    if not source:
        _set_points(f, xs, dim)
    if hyperplanes is not None:
        hp = hyperplanes if (t_first == 0 and t_local == ntrees) else slice_hyperplanes(hyperplanes, maxd, t_first, t_local)
        f.setHyperplanes(hp, t_local, maxd)
    else:
        f.genHyperplanes(seed, ntrees, maxd, pnz, dim, t_first, t_local)
    f.t_first, f.ntrees_total = t_first, ntrees
    if source:
        # `src .| chunksOf n .| foldl insertMulti`: n is unknown until the source is exhausted
        f.insertBegin(dim, maxd, minl)
        for chunk in chunksOf(chunksize, xs, dim):
            f.insertChunk(chunk)
        f.insertEnd()
    else:
        f.build(maxd, minl, chunk=chunksize)
    return f


def chunksOf(n, src, dim):
    """chunkedAccum's `C.chunksOf n` (Conduit.hs:168-176): groups a stream of rows (1-d arrays, or 2-d blocks of rows) into n x dim chunks, last one shorter."""
    buf = np.empty((n, dim), np.float64)
    fill = 0
    for item in src:
        rows = np.asarray(item, np.float64)
        rows = rows.reshape(1, dim) if rows.ndim == 1 else rows
        a = 0
        while a < rows.shape[0]:
            take = min(n - fill, rows.shape[0] - a)
            buf[fill:fill + take] = rows[a:a + take]
            fill += take; a += take
            if fill == n:
                yield buf.copy()
                fill = 0
    if fill:
        yield buf[:fill].copy()


def tree(seed, maxd, minl, chunksize, pnz, dim, xs, **kw):
    """tree (Conduit.hs:58-75)."""
    return forest(seed, maxd, minl, 1, chunksize, pnz, dim, xs, **kw)


def _q_for_call(q, d):
    """One query vector / a batch (dense), or SparseRows (passed through): -> (queries, single?)"""
    if isinstance(q, SparseRows):
        return q, False
    Q, single = _as_qd(q, d)
    return Q, single


def serialiseRPForest(tts, path, with_points=True):
    """serialiseRPForest (Internal.hs:185-190) counterpart: writes the engine's flat checkpoint."""
    tts.save(path, with_points)


def deserialiseRPForest(path, device=0):
    """deserialiseRPForest (Internal.hs:192-196) counterpart: a queryable forest from a checkpoint that carries its points."""
    f = RPForest(device)
    f.load(path)
    return f


def _need_l2(distf):
    if distf is not metricL2:
        raise RPForestError("only distf = metricL2 is implemented on the GPU path")


def knn(distf, k, tts, q):
    """knn (RPTree.hs:168-176): (distances, row ids) in increasing distance; duplicates across trees are kept."""
    _need_l2(distf)
    Q, single = _q_for_call(q, tts.d)
    dist, ids, cnt = tts.knnBatch(Q, k, dedup=False)
    if single:
        return dist[0, : cnt[0]], ids[0, : cnt[0]]
    return dist, ids, cnt


def knnPQ(distf, k, tts, q):
    """knnPQ (RPTree.hs:181-194): one result per distinct distance."""
    _need_l2(distf)
    Q, single = _q_for_call(q, tts.d)
    dist, ids, cnt = tts.knnBatch(Q, k, dedup=True)
    if single:
        return dist[0, : cnt[0]], ids[0, : cnt[0]]
    return dist, ids, cnt


def knnH(distf, k, tts, q):
    """knnH (RPTree.hs:199-217): leaves in margin-priority order, accumulated while the total stays <= k.  Like the
    reference the result is neither sorted by distance nor cut to k."""
    _need_l2(distf)
    Q, single = _q_for_call(q, tts.d)
    dist, ids, cnt = tts.knnHBatch(Q, k)
    if single:
        return dist[0, : cnt[0]], ids[0, : cnt[0]]
    return dist, ids, cnt


def candidates(tts, t, q):
    """candidates (RPTree.hs:293-314) of tree `t` of the forest for query q: row ids in the reference's order."""
    Q, single = _q_for_call(q, tts.d)
    off, ids = tts.candidatesBatch(Q, t)
    if single:
        return ids
    return off, ids


def recallWith(distf, tts, k, q):
    """recallWith (RPTree.hs:259-268): mean over the forest's trees of the per-tree candidate recall@k."""
    _need_l2(distf)
    Q, single = _q_for_call(q, tts.d)
    r = tts.recallSumBatch(Q, k) / float(tts.ntrees)
    return float(r[0]) if single else r


def levels(tts):
    """levels (Internal.hs:203-204): number of projection vectors per tree."""
    return tts.maxDepth


def leafSizes(tts):
    """leafSizes (RPTree.hs:366-367): sizes of the Tips, left to right (identical for every tree)."""
    tp = tts.topology()
    leaf = tp["child"] < 0
    order = np.argsort(tp["seg_start"][leaf], kind="stable")
    return tp["seg_size"][leaf][order]


def treeSize(tts):
    """treeSize (RPTree.hs:362-363)."""
    return int(leafSizes(tts).sum())


def points(tts, t):
    """points (Internal.hs:207-208): row ids of tree t, leaves left to right."""
    return tts.treeExport(t)["perm"]
