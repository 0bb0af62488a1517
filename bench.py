#!/usr/bin/env python
"""bench.py -- forest-build points/s (+ kNN queries/s, recall@10) on BASELINE.json configs[1]:
synthetic SIFT-like 1M x 128 fp64, 32-tree forest, pnz = 0.1, 10k queries, k = 10.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over the batch: build the whole forest (points already resident in HBM)
and answer the whole query batch.  `value` = points/s of the forest build (device time, CUDA events on the
engine's stream, max over ranks); knn throughput and recall ride along as extra keys.  `e2e` = the same through
the public API with HOST buffers (pinned), H2D of the points and D2H of the forest inside the timed region.
For N > 1 (launched by torchrun) the 32 trees are sharded in contiguous blocks over the ranks, the data is
replicated, and the per-rank top-k lists are all-gathered over NCCL and merged by the engine's merge kernel.
`--impl reference` times the reference's CPU algorithm (the oracle port: no GHC in this image) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_COMMON = dict(data_seed=1234, query_seed=4321, forest_seed=1235137, clusters=256, sigma=0.25, pnz=0.1, min_leaf=64)
CONFIGS = {   # BASELINE.json configs[1..4]; configs[0] (MNIST) is the reference's own CPU case (data file absent from the checkout)
    "c2": dict(_COMMON, name="configs[1]: synthetic SIFT-like 1Mx128 fp64, 32 trees, pnz 0.1, 10k queries k=10",
               n=1_000_000, d=128, ntrees=32, nq=10_000, k=10),
    "c3": dict(_COMMON, name="configs[2]: synthetic GIST-like 1Mx960 fp64, 64 trees, pnz 0.1, 100k queries k=10",
               n=1_000_000, d=960, ntrees=64, nq=100_000, k=10),
    "c4": dict(_COMMON, name="configs[3]: synthetic 768-d embeddings, 5M points, 128 trees sharded across the GPUs, 100k queries k=10",
               n=5_000_000, d=768, ntrees=128, nq=100_000, k=10),
    "c5": dict(_COMMON, name="configs[4]: synthetic Deep1B-like 10Mx96, 256 trees, 1M queries k=100",
               n=10_000_000, d=96, ntrees=256, nq=1_000_000, k=100),
}
WORKLOAD = CONFIGS["c2"]          # the headline configuration (the one BASELINE.json's metric is quoted on)
GEN_BLOCK = 16384                 # rows per independently seeded block: any rank can generate any row range


def make_rows(n, d, seed, clusters, sigma, r0=0, r1=None, out=None, center_seed=99):
    """Rows [r0, r1) of the clustered Gaussian mixture: centres ~ N(0,1)^d, points = centre + N(0, sigma^2)^d.
    Block b of GEN_BLOCK rows is drawn from its own generator seeded (seed, b), so a row range is the same on every rank."""
    r1 = n if r1 is None else r1
    cen = np.random.default_rng(center_seed).normal(size=(clusters, d))
    X = np.empty((r1 - r0, d)) if out is None else out
    for b in range(r0 // GEN_BLOCK, (max(r1, r0 + 1) - 1) // GEN_BLOCK + 1):
        lo, hi = b * GEN_BLOCK, min(n, (b + 1) * GEN_BLOCK)
        rng = np.random.default_rng([seed, b])
        blk = rng.normal(size=(hi - lo, d))
        blk *= sigma
        blk += cen[rng.integers(0, clusters, size=hi - lo)]
        a, z = max(lo, r0), min(hi, r1)
        if z > a:
            X[a - r0:z - r0] = blk[a - lo:z - lo]
    return X


def make_points(n, d, seed, clusters, sigma):
    return make_rows(n, d, seed, clusters, sigma)


def config_dict(W, maxd, gpus):
    """The `config` object of the JSON line -- built by this one function for BOTH arms so they print the same dict."""
    return {"workload": W["name"], "n": W["n"], "d": W["d"], "ntrees": W["ntrees"], "pnz": W["pnz"], "max_depth": int(maxd),
            "min_leaf": W["min_leaf"], "queries": W["nq"], "k": W["k"],
            "parallelism": "trees sharded over %d GPU(s) in contiguous blocks, data replicated" % gpus,
            "l2": "inputs larger than L2 (X %.2f GB, keys %.1f GB per build over all trees)" % (
                W["n"] * W["d"] * 8 / 1e9, W["ntrees"] * maxd * W["n"] * 8 / 1e9)}


def slice_hp(hp, maxd, t_first, t_local):
    """CSR rows of trees [t_first, t_first + t_local) of a forest-wide hyperplane set, re-based to offset 0."""
    off, idx, val = hp
    a, b = t_first * maxd, (t_first + t_local) * maxd
    lo, hi = int(off[a]), int(off[b])
    return (np.ascontiguousarray(off[a:b + 1] - off[a], np.int64), np.ascontiguousarray(idx[lo:hi], np.int32),
            np.ascontiguousarray(val[lo:hi], np.float64))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML every few ms DURING the timed region (the in-process
    equivalent of the nvidia-smi --query-gpu=clocks.sm,...,clocks_event_reasons.* line in B200_PROFILING.md)."""

    def __init__(self, gpu_index, period_s=0.004):
        self.idx, self.period = gpu_index, period_s
        self.sm, self.reasons, self.power = [], set(), []
        self.max_sm = None
        self._stop = threading.Event()
        self.th = None
        self.err = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception as e:      # pragma: no cover
            self.err = "nvml unavailable: %r" % (e,)
            return
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in names.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(self.period)

    def stop(self):
        self._stop.set()
        if self.th:
            self.th.join(timeout=2)
        if not self.sm:
            return dict(sm_mhz=None, sm_max_mhz=self.max_sm, reasons=[self.err or "no samples"], samples=0)
        return dict(sm_mhz=float(np.median(self.sm)), sm_max_mhz=self.max_sm, reasons=sorted(self.reasons), samples=len(self.sm),
                    power_w_max=max(self.power) if self.power else None)


def algorithmic_bytes(W, tlocal, L, s_top, C_mean, fused_levels=0):
    """Algorithmic HBM bytes per launch for every kernel class.  fused_levels: top levels whose relabel
    pass also produced the next level's histogram (k_top_relabel_hist: bin + label + next key read, label + bin write = 16
    bytes per point instead of the 6 of a plain relabel; those levels have no k_top_hist launch)."""
    n, d = W["n"], W["d"]
    nrel = max(s_top, 1)
    return {
        "project": 8 * d * n + 8 * tlocal * L * n,                     # read X once, write every (tree, level) key
        "top_hist": tlocal * n * (8 + 2 + 2),                          # key + label read, 2-byte bin write
        "top_compact": tlocal * n * (2 + 2),                           # bin + label (keys only for the median bin)
        "top_relabel": tlocal * n * (6 * (nrel - fused_levels) + 16 * fused_levels) // nrel,   # bin + label read, label write (mean over the launches)
        "bottom": tlocal * n * (4 + 4 + 8 * (L - s_top) + (8 if s_top > 0 else 0)),   # perm r/w + one key per bottom level
        # fp32 filter pass (k_knn_f32: d % 4 == 0, plain knn): 4d bytes per candidate + ~(k + 4) exact rows; else the exact kernel
        "q_knn": W["nq"] * ((C_mean * (4 * d + 4) + (W["k"] + 4) * 8 * d + 8 * d + 12 * W["k"]) if d % 4 == 0 and d < 512
                            else (C_mean * (8 * d + 4) + 8 * d + 12 * W["k"])),
    }


def run_ours(args):
    import torch
    import rp_tree_b200 as R

    W = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    dist = None
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # torch.distributed is only the launcher-side plumbing here (barrier, max over ranks, handing the NCCL id to the
        # ranks); every data-path exchange runs inside the engine on its own communicator (rpf_comm_init_rank).
        # NCCL writes its banner / debug lines to stdout by default; stdout is reserved for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    n, d, T, k, nq = W["n"], W["d"], W["ntrees"], W["k"], W["nq"]
    cfg = R.rpTreeCfg(W["min_leaf"], n, d)
    maxd = cfg.fpMaxTreeDepth
    assert T >= world, "more ranks than trees"
    t_first, t_local = R.dist.shard_trees(T, world, rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def new_forest():
        f_ = R.RPForest(local_rank)
        if os.environ.get("RPF_BOTTOM_CAP"):
            f_.setBottomCap(int(os.environ["RPF_BOTTOM_CAP"]))
        if world > 1:                                  # the engine's own communicator: id made on rank 0, handed out once
            uid = torch.zeros(128, dtype=torch.uint8, device=dev)
            if rank == 0:
                uid = torch.frombuffer(bytearray(R.RPForest.commUniqueId()), dtype=torch.uint8).to(dev)
            dist.broadcast(uid, 0)
            f_.commInitRank(world, rank, bytes(uid.cpu().numpy().tobytes()))
        f_.setHyperplanes(hp, t_local, maxd)
        return f_

    # synthetic inputs.  Small enough (or one GPU): the whole matrix in PINNED host memory on every rank (so the e2e H2D is a
    # real DMA).  Large multi-GPU configurations: every rank generates and holds only the rows it uploads (rpf_set_points on
    # a communicator rank reads rows [r*per, (r+1)*per) only); the e2e arm is then not run.
    full_host = world == 1 or n * d * 8 <= 2.2e9
    per = -(-n // world)
    r0, r1 = (0, n) if full_host else (min(n, rank * per), min(n, (rank + 1) * per))
    Xp = torch.empty((r1 - r0, d), dtype=torch.float64, pin_memory=full_host)
    X = Xp.numpy()
    make_rows(n, d, W["data_seed"], W["clusters"], W["sigma"], r0, r1, out=X)
    Qp = torch.empty((nq, d), dtype=torch.float64, pin_memory=True)
    Q = Qp.numpy()
    make_rows(nq, d, W["query_seed"], W["clusters"], W["sigma"], out=Q)
    hp_all = R.sampleHyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
    hp = slice_hp(hp_all, maxd, t_first, t_local)

    f = new_forest()
    if full_host:
        f.setPoints(X)                  # resident in HBM before the timed region (the `value` arm); N > 1: row-sharded upload + all-gather
    else:
        f.setPointsRaw(X.ctypes.data - r0 * d * 8, n, d)
    kout = (torch.empty((nq, k), dtype=torch.float64, pin_memory=True).numpy(),
            torch.empty((nq, k), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32),
            torch.empty((nq,), dtype=torch.int32, pin_memory=True).numpy())

    def knn_step(forest):
        """knn over the whole forest: local top-k on this rank's trees; N > 1: ONE packed NCCL all-gather over NVLink and the
        merge kernel, all inside rpf_knn on the engine's stream.  Device ms from the engine's events around the whole call."""
        forest.knnBatch(Q, k, dedup=False, out=kout)
        return forest.lastDeviceMs()

    # ---- warm-up
    for _ in range(args.warmup):
        f.build(maxd, W["min_leaf"])
        knn_step(f)

    # ---- timed: device-resident arm
    lc0 = f.launchCount()
    cs = ClockSampler(local_rank)
    cs.start()
    barrier()
    b_ms, q_ms = [], []
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        f.build(maxd, W["min_leaf"])
        b_ms.append(f.lastDeviceMs())
        q_ms.append(knn_step(f))
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3 / args.steps
    clocks = cs.stop()
    launches = f.launchCount() - lc0
    build_ms = allmax(float(np.mean(b_ms)))
    knn_ms = allmax(float(np.mean(q_ms)))

    # ---- timed: e2e arm -- host (pinned) buffers through the public API; every step copies the points H2D (N > 1: 1/N of
    # every row block per rank + NVLink all-gather, overlapped with the projection), builds, and streams the forest back D2H
    # (export sink); the query step copies the queries H2D and the merged results D2H
    e2e = None
    nn = len(f.topology()["child"])
    if full_host:
        e2e_b, e2e_q = [], []
        g = new_forest()
        exp_bufs = {key: torch.empty((t_local, nn), dtype=torch.float64, pin_memory=True).numpy() for key in ("thr", "mlo", "mhi")}
        exp_bufs["perm"] = torch.empty((t_local, n), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
        g.setExportSink(exp_bufs)                                # builds stream the forest into these buffers while they run

        def e2e_build():
            g.buildFromHost(X, maxd, W["min_leaf"])              # forestBatch from pinned host memory
            return g.forestExport(exp_bufs)                      # waits for the streamed download (thr/mlo/mhi + perm of every local tree)

        for _ in range(2):                                       # warm-up (workspace allocation)
            e2e_build(); knn_step(g)
        for _ in range(args.steps):
            barrier()
            t0 = time.perf_counter()
            e2e_build()
            barrier()
            e2e_b.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            knn_step(g)
            barrier()
            e2e_q.append(time.perf_counter() - t0)
        g.close()
        if rank == 0:
            print("per-step build device ms: %s | e2e build s: %s | e2e knn s: %s" % (
                [round(x, 3) for x in b_ms], [round(x, 4) for x in e2e_b], [round(x, 4) for x in e2e_q]), file=sys.stderr)
        e2e_build_s = allmax(float(np.mean(e2e_b)))
        e2e_knn_s = allmax(float(np.mean(e2e_q)))
        e2e = {"value": n / e2e_build_s, "unit": "points/s",
               "h2d_bytes_per_step": int(n * d * 8 + world * (len(hp[1]) * 12 + len(hp[0]) * 8)),     # all ranks together
               "d2h_bytes_per_step": int(T * (nn * 24 + n * 4)), "build_s": e2e_build_s,
               "upload": ("one H2D of the n x d points (row blocks overlapped with the projection)" if world == 1 else
                          "row-sharded: 1/%d of every row block per rank over its own PCIe link + in-engine NCCL all-gather over NVLink, "
                          "overlapped with the projection" % world),
               "knn_queries_per_s": nq / e2e_knn_s, "knn_s": e2e_knn_s,
               "knn_h2d_bytes": int(world * nq * d * 8), "knn_d2h_bytes": int(world * (nq * k * 12 + nq * 4))}

    # ---- per-kernel profile (separate pass: event pairs around every launch) -> roofline of the dominant kernel
    f.setProfiling(True)
    f.build(maxd, W["min_leaf"])
    prof = f.profile()
    f.knnBatch(Q, k, out=kout)
    prof_q = f.profile()
    f.setProfiling(False)
    for name in ("q_project", "q_traverse", "q_knn", "merge"):
        if name in prof_q:
            prof[name] = prof_q[name]
    # candidates per query over the WHOLE forest (local trees of every rank added up): the re-rank's algorithmic bytes
    qs = min(512, nq)
    off, _ = f.candidatesBatch(Q[:qs], -1)
    c_local = float(off[-1]) / qs
    C_mean = c_local
    if dist is not None:
        t = torch.tensor([c_local], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        C_mean = float(t.item())
    tp = f.topology()
    L = int(tp["depth"][tp["child"] >= 0].max()) + 1
    cap = int(os.environ.get("RPF_BOTTOM_CAP", "1024"))      # engine default (rpf_set_bottom_cap)
    lvl_max = [int(tp["seg_size"][tp["depth"] == l].max()) for l in range(int(tp["depth"].max()) + 1)]
    s_top = next((l for l, m in enumerate(lvl_max) if m <= cap), len(lvl_max))
    s_top = min(s_top, L)
    nhist = prof["top_hist"][1] if "top_hist" in prof else 0
    nrel = prof["top_relabel"][1] if "top_relabel" in prof else 0
    # levels whose histogram came out of the previous level's fused pass (launch counts scale with the number of tree groups)
    fused_levels = int(round(s_top * max(0, nrel - nhist) / nrel)) if nrel > 0 and nhist > 0 else 0
    ab = algorithmic_bytes(W, t_local, L, s_top, c_local, fused_levels)    # per-rank kernels: this rank's trees / candidates
    kern = {kname: v for kname, v in prof.items() if v[1] > 0 and kname in ab}
    traffic_tab = {}
    for tf in ("r02_traffic.json", "r01_traffic.json"):      # DRAM bytes per launch from the committed `ncu --set full` captures
        try:
            with open(os.path.join(ROOT, "profiles", tf)) as fh:
                traffic_tab = json.load(fh)
            break
        except Exception:
            pass
    if args.config != "c2" or world != 1:
        traffic_tab = {}                                     # the captures were taken on configs[1] at one GPU

    def kernel_roofline(kname):
        """Roofline record of one kernel class.  top_* entries of `ab` are bytes per level launch (every launch streams all
        local trees' points once); the others are bytes per step, issued in `launches` launches (one per tree group)."""
        ms, nl = kern[kname]
        if kname == "q_knn":
            nl = 1      # the filter kernel + the (nearly empty) exact second pass over its flagged queries: one logical launch
        per_launch = ab[kname] if kname.startswith("top_") else ab[kname] / nl
        avg_ms = ms / nl
        ach = per_launch / (avg_ms * 1e-3) / 1e9
        tj = traffic_tab.get(kname)
        traffic = int(tj["dram_bytes_per_launch"]) if tj else None
        if tj and tj.get("trees_per_launch"):        # captured on one branch of the graph build (e.g. 16 of 32 trees per launch)
            traffic = int(traffic * t_local / tj["trees_per_launch"])
        r = dict(bound="hbm", kernel=kname, achieved=round(ach, 1), peak=peak, unit="GB/s", frac=round(ach / peak, 4),
                 traffic=traffic, traffic_source=("ncu dram__bytes_read+write: profiles/%s" % ",".join(tj["source"])) if tj else None,
                 peak_source=peak_src, launches_per_step=nl, avg_launch_ms=round(avg_ms, 4),
                 algorithmic_bytes_per_launch=int(per_launch))
        if traffic:
            r["frac_by_dram_traffic"] = round(traffic / (avg_ms * 1e-3) / 1e9 / peak, 4)
        return r

    peak, peak_src = peaks()
    # `roofline` describes the dominant kernel of the HEADLINE metric (the forest build); the re-rank kernel of the query
    # step has its own record (`roofline_knn_kernel`), where the DRAM-traffic fraction is the honest one: its algorithmic
    # bytes count every candidate row once per (query, tree) while rows shared by co-scheduled queries come from L2.
    bk = {kname: v for kname, v in kern.items() if not kname.startswith("q_")}
    bdom = max(bk, key=lambda kname: bk[kname][0])
    roofline = kernel_roofline(bdom)
    roofline_knn_kernel = kernel_roofline("q_knn") if "q_knn" in kern else None
    roofline_build_kernels = {kname: kernel_roofline(kname) for kname in bk}
    build_bytes = 8 * d * n + t_local * L * n * 24
    knn_bytes = ab["q_knn"]
    phases = {kname: dict(ms=round(v[0], 3), launches=v[1]) for kname, v in prof.items() if v[1] > 0}

    # ---- quality: recallWith (reference definition) and forest-level recall@k on a query sample (untimed)
    ns = 64
    recall_ref_def = float(np.mean(f.recallSumBatch(Q[:ns], k) / T))       # N > 1: forest-wide sums, reduced inside the engine
    bd, bi = f.bruteKnnBatch(Q[:ns], k)
    pd_, pi_, pc_ = f.knnBatch(Q[:ns], k, dedup=True)
    forest_recall = float(np.mean([len(set(pi_[i, :pc_[i]].tolist()) & set(bi[i].tolist())) / k for i in range(ns)]))

    # ---- streaming build (`forest` with the rpTreeCfg chunk size, Conduit.hs:104-141): reported beside the batch build
    stream = None
    if args.config == "c2":
        chunk = cfg.fpDataChunkSize
        t0 = time.perf_counter()
        f.build(maxd, W["min_leaf"], chunk=chunk)                    # first call: plans the 100 chunks on the host
        stream_first_ms = (time.perf_counter() - t0) * 1e3
        s_ms = []
        for _ in range(3):
            f.build(maxd, W["min_leaf"], chunk=chunk)
            s_ms.append(f.lastDeviceMs())
        stream_ms = allmax(float(np.mean(s_ms)))
        stream = {"chunk": int(chunk), "chunks": int(-(-n // chunk)), "device_ms": stream_ms, "points_per_s": n / (stream_ms * 1e-3),
                  "first_call_ms_incl_host_plan": stream_first_ms, "points_lost_by_reference_rule": f.pointsLost()}
        f.build(maxd, W["min_leaf"])

    out = None
    if rank == 0:
        value = n / (build_ms * 1e-3)
        out = {
            "metric": "forest_build_points_per_s", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": build_ms + knn_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(W, maxd, world), "trees_per_gpu": t_local,
            "multi_gpu": None if world == 1 else "rpf_comm_init_rank: NCCL communicator owned by the engine (C ABI); exchanges inside rpf_set_points / rpf_build_from_host / rpf_knn / rpf_recall",
            "build_ms": build_ms, "knn_ms": knn_ms, "knn_queries_per_s": nq / (knn_ms * 1e-3),
            "recall_at_k_recallWith": recall_ref_def, "recall_at_k_forest": forest_recall, "recall_queries": ns,
            "candidates_per_query": C_mean, "stream_build": stream,
            "e2e": e2e if e2e is not None else {"value": None, "unit": "points/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                                                "note": "not measured: at this configuration every rank's host holds only the rows it uploads"},
            "gpu_launches": int(launches), "wall_ms_per_step": wall_ms,
            "roofline": roofline, "roofline_knn_kernel": roofline_knn_kernel, "roofline_build_kernels": roofline_build_kernels,
            "roofline_build": dict(bound="hbm", achieved=round(build_bytes / (build_ms * 1e-3) / 1e9, 1), peak=peak, unit="GB/s",
                                   frac=round(build_bytes / (build_ms * 1e-3) / 1e9 / peak, 4), algorithmic_bytes=int(build_bytes)),
            "roofline_knn": dict(bound="hbm", achieved=round(knn_bytes / (knn_ms * 1e-3) / 1e9, 1), peak=peak, unit="GB/s",
                                 frac=round(knn_bytes / (knn_ms * 1e-3) / 1e9 / peak, 4), algorithmic_bytes=int(knn_bytes)),
            "phases": phases, "clocks": clocks,
        }
        if not args.no_cpu and world == 1:
            # CPU leg (rank 0, N=1 only): the oracle builds the whole forest on all host threads, answers a query sample, and
            # is the CHECKER of the GPU results on that sample (knn ids / distance bits, recallWith within 0.005)
            threads = min(os.cpu_count() or 1, T)
            cb, (ores, orec) = cpu_baseline(X, Q, hp_all, W, maxd, threads, full_forest=True, nq_knn=args.cpu_queries)
            gd, gi, gc = f.knnBatch(Q[:len(ores)], k)
            for i, (od, oi) in enumerate(ores):
                assert np.array_equal(gi[i, :gc[i]], oi) and np.array_equal(gd[i, :gc[i]].view(np.uint64), od.view(np.uint64)), \
                    "knn of query %d differs from the oracle" % i
            grec = f.recallSumBatch(Q[:len(orec)], k) / T
            cb["recall_gpu_same_queries"] = float(grec.mean())
            cb["recall_abs_diff"] = float(abs(grec.mean() - orec.mean()))
            cb["knn_ids_and_distance_bits_equal_oracle"] = True
            assert cb["recall_abs_diff"] <= 0.005, "recallWith@%d differs from the oracle by %g" % (k, cb["recall_abs_diff"])
            out["cpu_baseline"] = cb
    barrier()
    f.close()
    if dist is not None:
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def cpu_build(X, hp_all, W, maxd, trees, threads):
    """Trees [0, trees) of the forest built by the C oracle (reference algorithm: per-node stable merge sort), the
    independent trees spread over `threads` host threads.  Returns (oracle forest, seconds)."""
    from oracle import orc
    orc.lib()
    hp = slice_hp(hp_all, maxd, 0, trees)
    t0 = time.perf_counter()
    f = orc.Forest(X, hp, trees, maxd, W["min_leaf"], threads=threads)
    dt = time.perf_counter() - t0
    assert all(f.tree_size(t) == W["n"] for t in range(trees))
    return f, dt


def cpu_queries(of, Q, k, threads, nq_knn, nq_recall):
    """The reference's query path on the oracle forest: knn (RPTree.hs:168-176) on the first nq_knn queries -> queries/s,
    recallWith (RPTree.hs:259-282; the harness of bench/time/Main.hs:66-84 times exactly this action) on the first
    nq_recall.  Queries are spread over `threads` host threads (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=threads) as ex:
        t0 = time.perf_counter()
        res = list(ex.map(lambda i: of.knn(Q[i], k), range(nq_knn)))
        knn_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        rec = list(ex.map(lambda i: of.recall_shared(Q[i], k), range(nq_recall)))
        rec_s = time.perf_counter() - t0
    return res, knn_s, np.asarray(rec), rec_s


def cpu_baseline(X, Q, hp_all, W, maxd, threads, full_forest, nq_knn=2048, nq_recall=64):
    """The reference algorithm (oracle port, `kind: port`) on the host cores, bounded sample.  Build: `threads` trees at once
    (one tree per thread; the reference itself is single threaded) on the full data, the forest time extrapolated from
    trees built / T -- or, with full_forest, every tree of the forest (no extrapolation).  Queries: see cpu_queries; only
    with full_forest (the candidate sets of a partial forest are not the workload's)."""
    T = W["ntrees"]
    ntr = T if full_forest else min(threads, T)
    of, dt = cpu_build(X, hp_all, W, maxd, ntr, threads)
    forest_s = dt * T / ntr
    out = {"value": W["n"] / forest_s, "unit": "points/s", "cores": threads, "kind": "port",
           "sample": "build: %d of %d trees on the full %dx%d data by the C oracle (reference algorithm: per-node stable merge sort) on "
                     "%d threads, %.1f s%s" % (ntr, T, W["n"], W["d"], threads, dt, "" if ntr == T else "; forest time extrapolated x%g" % (T / ntr)),
           "seconds": dt}
    oracle_q = None
    if full_forest:
        res, knn_s, rec, rec_s = cpu_queries(of, Q, W["k"], threads, min(nq_knn, len(Q)), min(nq_recall, len(Q)))
        out.update({"knn_queries_per_s": len(res) / knn_s, "knn_seconds": knn_s, "knn_queries": len(res),
                    "recall_oracle": float(rec.mean()), "recall_queries": len(rec), "recall_seconds": rec_s,
                    "recall_queries_per_s": len(rec) / rec_s})
        out["sample"] += "; knn: first %d queries, %.2f s; recallWith@%d: first %d queries, %.1f s (distances evaluated once per query)" % (
            len(res), knn_s, W["k"], len(rec), rec_s)
        oracle_q = (res, rec)
    return out, oracle_q


def run_reference(args):
    """--impl reference: the reference's CPU algorithm for the same metric/config on all host threads.  Loads only the
    oracle (oracle/liborc.so); nothing of the product is imported on this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    W = CONFIGS[args.config]
    n, d, T = W["n"], W["d"], W["ntrees"]
    maxd = orc.rptree_cfg(W["min_leaf"], n, d)[0]
    X = make_points(n, d, W["data_seed"], W["clusters"], W["sigma"])
    Q = make_points(W["nq"], d, W["query_seed"], W["clusters"], W["sigma"])
    hp_all = orc.gen_hyperplanes(W["forest_seed"], T, maxd, W["pnz"], d)
    threads = min(os.cpu_count() or 1, T)
    vals = []
    for i in range(args.warmup + args.steps):                    # each step: one round of `threads` trees (bounded sample)
        cb, _ = cpu_baseline(X, Q, hp_all, W, maxd, threads, full_forest=False)
        if i >= args.warmup:
            vals.append(cb)
    full, _ = cpu_baseline(X, Q, hp_all, W, maxd, threads, full_forest=True, nq_knn=args.cpu_queries)   # untimed for `value`: whole forest once + the query legs
    v = float(np.mean([c["value"] for c in vals]))
    cb = dict(full)
    cb.update({"value": v, "seconds": float(np.mean([c["seconds"] for c in vals])), "full_forest_points_per_s": full["value"],
               "full_forest_seconds": full["seconds"]})
    cb["sample"] = ("per step: %d of %d trees on the full %dx%d data on %d threads, forest time extrapolated x%g; after the timed steps "
                    "the whole forest once (%.1f s) for the query legs: %s" % (
                        min(threads, T), T, n, d, threads, T / min(threads, T), full["seconds"], full["sample"].split("; ", 1)[1]))
    out = {"impl": "reference", "metric": "forest_build_points_per_s", "value": v, "unit": "points/s", "n_gpus": args.gpus,
           "steps": len(vals), "warmup": args.warmup, "ms_per_step": float(np.mean([c["seconds"] for c in vals])) * 1e3,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(W, maxd, args.gpus),
           "knn_queries_per_s": full["knn_queries_per_s"], "recall_at_k_recallWith": full["recall_oracle"],
           "cpu_baseline": cb,
           "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "reference = C port of the Haskell algorithm (no GHC toolchain in this image), trees spread over all host threads "
                   "(the Haskell reference is single threaded)"}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json configuration (c2 = configs[1], the headline)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-queries", type=int, default=10000, help="queries answered by the CPU knn leg (and checked against the GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
