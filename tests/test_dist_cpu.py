"""World-size-2 gloo test (CPU) of the multi-GPU host path: tree sharding, the all-gather of per-rank top-k lists and
the rank-major merge rule.  Per-rank lists come from the oracle here (no GPU in this container); the merge applied is
the rule rpf_merge_topk implements on the device (order by (distance, rank, position)), restated in numpy for the test.
The GPU tests check the CUDA merge kernel itself against the whole-forest result."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_numpy(D, I, C, k, dedup):
    G, nq, _ = D.shape
    od = np.full((nq, k), np.inf); oi = np.full((nq, k), 0xFFFFFFFF, np.uint32); oc = np.zeros(nq, np.int32)
    for q in range(nq):
        ent = [(D[g, q, j], g, j, I[g, q, j]) for g in range(G) for j in range(C[g, q])]
        ent.sort(key=lambda e: (e[0], e[1], e[2]))
        out = []
        for e in ent:
            if dedup and out and out[-1][0] == e[0]:
                continue
            out.append(e)
            if len(out) == k:
                break
        oc[q] = len(out)
        for j, e in enumerate(out):
            od[q, j] = e[0]; oi[q, j] = e[3]
    return od, oi, oc


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        import rp_tree_b200 as R
        from oracle import orc
        from helpers import make_data
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        n, d, T, maxd, minl, pnz, k = 3000, 8, 6, 7, 12, 0.5, 10
        X = make_data(n, d, 5, "mixture")
        hp = orc.gen_hyperplanes(11, T, maxd, pnz, d)
        t_first, t_local = R.dist.shard_trees(T, world, rank)
        shard_hp = R.slice_hyperplanes(hp, maxd, t_first, t_local)
        local = orc.Forest(X, shard_hp, t_local, maxd, minl)          # this rank's trees
        Q = X[:16] + 0.01
        ok = True
        for dedup in (False, True):
            dd = np.full((len(Q), k), np.inf); ii = np.zeros((len(Q), k), np.uint32); cc = np.zeros(len(Q), np.int32)
            for i in range(len(Q)):
                a, b = local.knn(Q[i], k, dedup=dedup)
                dd[i, :len(a)] = a; ii[i, :len(b)] = b; cc[i] = len(a)
            D, I, Cn = R.dist.gather_topk(dd, ii, cc)
            assert D.shape == (world, len(Q), k) and Cn.shape == (world, len(Q))
            md, mi, mc = _merge_numpy(D, I, Cn, k, dedup)
            whole = orc.Forest(X, hp, T, maxd, minl)
            for i in range(len(Q)):
                a, b = whole.knn(Q[i], k, dedup=dedup)
                ok &= mc[i] == len(a) and np.array_equal(md[i, :mc[i]].view(np.uint64), a.view(np.uint64)) and np.array_equal(mi[i, :mc[i]], b)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


def test_shard_trees_contiguous_and_complete():
    import rp_tree_b200 as R
    for T in (1, 7, 32, 33):
        for world in (1, 2, 3, 4, 8):
            blocks = [R.dist.shard_trees(T, world, r) for r in range(world)]
            covered = []
            for t0, tl in blocks:
                covered += list(range(t0, t0 + tl))
            assert covered == list(range(T))


def test_world2_gloo_gather_and_merge(built):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in res:
        assert ok, "rank %d failed: %s" % (rank, msg)


def _rows_worker(rank, world, port, q):
    try:
        import torch
        import torch.distributed as dist
        import rp_tree_b200 as R
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        ok = True
        for n, slices in ((1001, 4), (64, 4), (3, 1), (10000, 4), (8200, 3), (4097, 2)):
            X = np.random.default_rng(3).normal(size=(n, 5))
            rp = R.dist.ReplicatedPoints(torch.device("cpu"), slices=slices)
            S, sub, nsl = rp.layout(n, world)
            ok &= nsl * S >= n and S == sub * world and (nsl == 1 or n >= world * 1024 * nsl)
            ptr, nn, d = rp.upload(X)
            ok &= (nn, d) == (n, 5) and ptr == rp.buf.data_ptr()
            ok &= np.array_equal(rp.buf[:n].numpy().view(np.uint64), X.view(np.uint64))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bool(ok), ""))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


def test_shard_rows_cover_all_rows():
    import rp_tree_b200 as R
    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                per, r0, rl = R.dist.shard_rows(n, world, r)
                assert rl <= per
                rows += list(range(r0, r0 + rl))
            assert rows == list(range(n))


def test_world2_gloo_row_sharded_replica(built):
    """ReplicatedPoints: each rank contributes its row block, every rank ends up with the whole matrix (bit-identical)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, msg in res:
        assert ok, "rank %d failed: %s" % (rank, msg)
