"""torchrun check (N GPUs): the forest built from the row-sharded upload (ReplicatedPoints: n/N rows per rank over PCIe +
NCCL all-gather) is bit-identical to the forest built from a full per-rank upload; prints the timing of both.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_upload.py
"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench, rp_tree_b200 as R

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
W = bench.WORKLOAD
n, d, T = 200_000 + 8 * 13, W["d"], 8
maxd = R.rpTreeCfg(W["min_leaf"], n, d).fpMaxTreeDepth
Xp = torch.empty((n, d), dtype=torch.float64, pin_memory=True)
Xp.numpy()[:] = bench.make_points(n, d, 7, 64, 0.25)
t0, tl = R.dist.shard_trees(T, world, rank)
hp = R.slice_hyperplanes(R.sampleHyperplanes(5, T, maxd, W["pnz"], d), maxd, t0, tl)
a = R.RPForest(lr); a.setHyperplanes(hp, tl, maxd)
b = R.RPForest(lr); b.setHyperplanes(hp, tl, maxd)
rp = R.dist.ReplicatedPoints(dev)
for it in range(3):
    dist.barrier(); torch.cuda.synchronize(); t = time.perf_counter()
    a.buildFromHost(Xp.numpy(), maxd, W["min_leaf"])
    torch.cuda.synchronize(); dist.barrier(); ta = time.perf_counter() - t
    t = time.perf_counter()
    R.dist.buildFromHostSharded(b, rp, Xp, maxd, W["min_leaf"])
    torch.cuda.synchronize(); dist.barrier(); tb = time.perf_counter() - t
ok = True
for tr in range(tl):
    ea, eb = a.treeExport(tr), b.treeExport(tr)
    for key in ("thr", "mlo", "mhi"):
        ok &= np.array_equal(ea[key].view(np.uint64), eb[key].view(np.uint64))
    ok &= np.array_equal(ea["perm"], eb["perm"])
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded upload identical on all %d ranks: %s | full upload + build %.2f ms, sharded upload + all-gather + build %.2f ms"
          % (world, bool(flag.item()), ta * 1e3, tb * 1e3))
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
