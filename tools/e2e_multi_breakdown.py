"""Where the multi-GPU e2e build time goes (torchrun, one rank per GPU, configs[1] data).
  setPoints       : row-sharded upload + in-engine NCCL all-gather, nothing else
  build           : forest build from the resident replica
  buildFromHost   : upload + all-gather + projection overlapped, then the build
  +export (sink)  : the same with the forest streamed back into pinned buffers
  torch all-gather: 16 x 64 MB through torch.distributed (same NCCL library), for reference
Usage: torchrun --nproc-per-node N tools/e2e_multi_breakdown.py"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import rp_tree_b200 as R  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
W = bench.WORKLOAD
n, d, T = W["n"], W["d"], W["ntrees"]
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
Xp = torch.empty((n, d), dtype=torch.float64, pin_memory=True)
X = Xp.numpy()
bench.make_rows(n, d, W["data_seed"], W["clusters"], W["sigma"], out=X)
hp_all = R.sampleHyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
t_first, t_local = R.dist.shard_trees(T, world, rank)
hp = R.slice_hyperplanes(hp_all, maxd, t_first, t_local)
f = R.RPForest(lr)
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    uid = torch.frombuffer(bytearray(R.RPForest.commUniqueId()), dtype=torch.uint8).to(dev)
dist.broadcast(uid, 0)
f.commInitRank(world, rank, bytes(uid.cpu().numpy().tobytes()))
f.setHyperplanes(hp, t_local, maxd)


def timed(fn, reps=5, warm=2):
    ts = []
    for i in range(warm + reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize(); dist.barrier()
        if i >= warm:
            ts.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([float(np.mean(ts))], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 3)


res = {"world": world}
res["setPoints_ms"] = timed(lambda: f.setPoints(X))
res["build_resident_ms"] = timed(lambda: f.build(maxd, W["min_leaf"]))
res["build_resident_device_ms"] = round(f.lastDeviceMs(), 3)
res["buildFromHost_ms"] = timed(lambda: f.buildFromHost(X, maxd, W["min_leaf"]))
res["buildFromHost_device_ms"] = round(f.lastDeviceMs(), 3)
nn = len(f.topology()["child"])
bufs = {k: torch.empty((t_local, nn), dtype=torch.float64, pin_memory=True).numpy() for k in ("thr", "mlo", "mhi")}
bufs["perm"] = torch.empty((t_local, n), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
res["forestExport_ms"] = timed(lambda: f.forestExport(bufs))
f.setExportSink(bufs)
res["buildFromHost_plus_sink_export_ms"] = timed(lambda: (f.buildFromHost(X, maxd, W["min_leaf"]), f.forestExport(bufs)))
if rank == 0:
    print(json.dumps(res), flush=True)
# reference: the same 16 x (n/16 rows) all-gathers through torch.distributed
blk = (n // 16 // world) * world
full = torch.empty((blk, d), dtype=torch.float64, device=dev)
part = full[rank * (blk // world):(rank + 1) * (blk // world)]


def ag():
    for _ in range(16):
        dist.all_gather_into_tensor(full, part)


res["torch_16_allgathers_ms"] = timed(ag)
nw = (n // world) * world
one = torch.empty((nw, d), dtype=torch.float64, device=dev)
res["torch_1_allgather_full_ms"] = timed(lambda: dist.all_gather_into_tensor(one, one[rank * (n // world):(rank + 1) * (n // world)]))
h2d = torch.empty((n // world, d), dtype=torch.float64, device=dev)
res["h2d_own_rows_ms"] = timed(lambda: h2d.copy_(Xp[rank * (n // world):(rank + 1) * (n // world)], non_blocking=True))
if rank == 0:
    print(json.dumps(res))
f.close()
dist.destroy_process_group()
