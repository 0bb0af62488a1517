"""Tree-sharded multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in the
CPU tests).  Trees are independent (createMulti maps over the IntMap of trees, src/Data/RPTree/Internal.hs:234-240; knn
folds the per-tree candidates in ascending tree order, src/Data/RPTree.hs:176), so the forest is partitioned in
CONTIGUOUS blocks of trees, the data is replicated, the build needs no communication, and a query needs exactly one
exchange: an all-gather of the per-rank (dist, id, count) top-k lists followed by the engine's merge kernel
(rpf_merge_topk).  Getting the replicated data onto the GPUs is the other exchange: every rank uploads only its row block
over its own PCIe link and the blocks are all-gathered over NVLink (ReplicatedPoints), instead of every rank pulling
the whole n x d matrix through the host.  Contiguous blocks make (rank, position) order equal to the reference's tree order, so the merged
result is identical to the single-GPU result, ties included.
"""
import numpy as np


def shard_trees(ntrees, world, rank):
    """Contiguous block of trees owned by `rank`: (t_first, t_local).  Balanced: the first ntrees % world ranks hold one
    tree more, so every rank owns at least one tree whenever ntrees >= world (contiguous blocks keep the merge order)."""
    base, rem = divmod(ntrees, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist


def gather_topk(dist_arr, ids_arr, cnt_arr, group=None, device=None):
    """All-gather per-rank top-k lists.  Inputs: nq x k float64, nq x k uint32, nq int32 (numpy).
    Returns rank-major stacks (G x nq x k, G x nq x k, G x nq) as numpy arrays on every rank."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    dev = device if device is not None else torch.device("cpu")
    d = torch.from_numpy(np.ascontiguousarray(dist_arr, np.float64)).to(dev)
    i = torch.from_numpy(np.ascontiguousarray(ids_arr, np.uint32).view(np.int32)).to(dev)
    c = torch.from_numpy(np.ascontiguousarray(cnt_arr, np.int32)).to(dev)
    gd = [torch.empty_like(d) for _ in range(world)]
    gi = [torch.empty_like(i) for _ in range(world)]
    gc = [torch.empty_like(c) for _ in range(world)]
    dist.all_gather(gd, d, group=group)
    dist.all_gather(gi, i, group=group)
    dist.all_gather(gc, c, group=group)
    return (torch.stack(gd).cpu().numpy(), torch.stack(gi).cpu().numpy().view(np.uint32), torch.stack(gc).cpu().numpy())


def forestBatchSharded(seed, maxd, minl, ntrees, pnz, dim, xs, *, hyperplanes=None, device=0, group=None, bottom_cap=None):
    """forestBatch (Batch.hs:48-63) with this rank's contiguous block of trees; data replicated on every GPU."""
    from .api import forestBatch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if ntrees < world:                        # the same verdict on every rank, before any collective
        raise ValueError("more ranks (%d) than trees (%d)" % (world, ntrees))
    t_first, t_local = shard_trees(ntrees, world, rank)
    return forestBatch(seed, maxd, minl, ntrees, pnz, dim, xs, hyperplanes=hyperplanes, device=device,
                       t_first=t_first, t_local=t_local, bottom_cap=bottom_cap)


def knnSharded(forest, k, Q, dedup=False, group=None, device=None):
    """knn / knnPQ over the whole forest: local top-k on this rank's trees, all-gather, merge kernel (on every rank)."""
    d, i, c = forest.knnBatch(Q, k, dedup=dedup)
    D, I, Cn = gather_topk(d, i, c, group=group, device=device)
    return forest.mergeTopk(D, I, Cn, dedup=dedup)


def knnShardedDevice(forest, k, Q, dedup=False, group=None, device=None, timed=False):
    """Same result as knnSharded with the exchange kept on the devices: this rank's lists are written into CUDA tensors,
    all-gathered over NCCL (NVLink), merged by the engine's merge kernel from the gathered device buffer; only the merged
    nq x k result crosses PCIe.  Returns the merged lists; with timed=True also (device ms of the engine's two calls,
    device ms of the all-gather from CUDA events on torch's stream)."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    Q = np.ascontiguousarray(Q, np.float64)
    nq = Q.shape[0]
    d = torch.empty((nq, k), dtype=torch.float64, device=device)
    i = torch.empty((nq, k), dtype=torch.int32, device=device)
    c = torch.empty((nq,), dtype=torch.int32, device=device)
    gd = torch.empty((world, nq, k), dtype=torch.float64, device=device)
    gi = torch.empty((world, nq, k), dtype=torch.int32, device=device)
    gc = torch.empty((world, nq), dtype=torch.int32, device=device)
    stream = torch.cuda.current_stream(device)
    stream.synchronize()                                         # the buffers exist before the engine's stream writes them
    forest.knnBatchDevice(Q, k, d.data_ptr(), i.data_ptr(), c.data_ptr(), dedup=dedup)
    ms = forest.lastDeviceMs()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    dist.all_gather_into_tensor(gd, d, group=group)
    dist.all_gather_into_tensor(gi, i, group=group)
    dist.all_gather_into_tensor(gc, c, group=group)
    ev1.record(stream)
    stream.synchronize()                                         # gathered lists complete before the engine's stream reads them
    out = forest.mergeTopkDevice(world, nq, k, gd.data_ptr(), gi.data_ptr(), gc.data_ptr(), dedup=dedup)
    ms += forest.lastDeviceMs()
    if timed:
        return out, ms, ev0.elapsed_time(ev1)
    return out


def recallSharded(forest, k, Q, group=None, device=None):
    """recallWith (RPTree.hs:259-268): per-rank sums over local trees, all-reduced, divided by the forest size."""
    import torch
    dist = _dist()
    dev = device if device is not None else torch.device("cpu")
    r = torch.from_numpy(forest.recallSumBatch(Q, k)).to(dev)
    dist.all_reduce(r, group=group)
    return r.cpu().numpy() / float(forest.ntrees_total)


def shard_rows(n, world, rank):
    """Equal row blocks for the all-gather of the replicated points: (rows per rank, first row, rows of this rank)."""
    per = (n + world - 1) // world
    r0 = min(rank * per, n)
    return per, r0, max(0, min(n, r0 + per) - r0)


def gather_rows(full, block, group=None):
    """All-gather equal row blocks into `full` (world*per x d); `block` may alias full[rank*per:(rank+1)*per] (in place)."""
    dist = _dist()
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(full, block, group=group)
    else:                                                        # gloo (CPU tests): list form
        import torch
        world = dist.get_world_size(group)
        parts = [torch.empty_like(block) for _ in range(world)]
        dist.all_gather(parts, block.clone(), group=group)
        per = block.shape[0]
        for r, p in enumerate(parts):
            full[r * per:(r + 1) * per].copy_(p)


class ReplicatedPoints:
    """Device-resident replica of the host data set, filled by a row-sharded upload: the rows are cut into `slices`
    contiguous slices; inside slice s rank r copies ITS sub-block of rows from (pinned) host memory over its own PCIe
    link, and the sub-blocks of a slice are all-gathered in place over NCCL (NVLink / NVSwitch) -- the all-gather of slice s
    overlaps the PCIe upload of slice s+1 (uploads on a side stream, one event per slice).  Every GPU ends up with all n
    rows; H2D traffic per rank is n*d*8 / world bytes instead of n*d*8.  The buffer is kept across calls."""

    def __init__(self, device, group=None, slices=4):
        self.device, self.group, self.slices = device, group, max(1, int(slices))
        self.buf = None
        self.n = self.d = 0
        self._copy_stream = None

    def layout(self, n, world):
        """(rows per slice, rows per rank inside a slice, number of slices): slice s covers rows [s*S, (s+1)*S)."""
        nsl = max(1, min(self.slices, n // max(world * 1024, 1)))       # no slicing for small inputs
        sub = (n + nsl * world - 1) // (nsl * world)
        return sub * world, sub, nsl

    def upload(self, X):
        """X: n x d float64 host array or CPU tensor (ideally pinned), identical on every rank.  Returns (device pointer, n, d)."""
        import torch
        dist = _dist()
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if not isinstance(X, torch.Tensor):
            X = torch.from_numpy(np.ascontiguousarray(X, np.float64))
        n, d = X.shape
        S, sub, nsl = self.layout(n, world)
        if self.buf is None or self.buf.shape != (nsl * S, d):
            self.buf = torch.empty((nsl * S, d), dtype=torch.float64, device=self.device)
        cuda = self.buf.is_cuda
        if cuda and self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.buf.device)
        main = torch.cuda.current_stream(self.buf.device) if cuda else None
        if cuda:
            self._copy_stream.wait_stream(main)                  # the previous contents are no longer read
        events = []
        for s in range(nsl):                                     # this rank's sub-block of every slice: PCIe, side stream
            r0 = s * S + rank * sub
            rl = max(0, min(n, r0 + sub) - r0)
            mine = self.buf[r0:r0 + sub]
            if cuda:
                with torch.cuda.stream(self._copy_stream):
                    if rl > 0:
                        mine[:rl].copy_(X[r0:r0 + rl], non_blocking=True)
                    if rl < sub:
                        mine[rl:].zero_()
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                events.append(ev)
            else:
                if rl > 0:
                    mine[:rl].copy_(X[r0:r0 + rl])
                if rl < sub:
                    mine[rl:].zero_()
        for s in range(nsl):                                     # slice by slice: NVLink all-gather as soon as the slice is up
            if cuda:
                main.wait_event(events[s])
            r0 = s * S + rank * sub
            gather_rows(self.buf[s * S:(s + 1) * S], self.buf[r0:r0 + sub], group=self.group)
        if cuda:
            main.synchronize()                                   # complete before the engine's stream reads it
        self.n, self.d = n, d
        return self.buf.data_ptr(), n, d


def buildFromHostSharded(forest, points, X, maxd, minl):
    """forestBatch on this rank's trees from HOST data: row-sharded upload + NVLink all-gather (ReplicatedPoints), then
    the batch build on the borrowed device replica.  Same forest as forest.buildFromHost(X, ...), bit for bit."""
    ptr, n, d = points.upload(X)
    if getattr(forest, "_borrowed_points", None) != (ptr, n, d):
        forest.setPointsDevice(ptr, n, d)
        forest._borrowed_points = (ptr, n, d)
    forest.build(maxd, minl)
