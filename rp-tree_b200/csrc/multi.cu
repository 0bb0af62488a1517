// multi.cu -- tree-sharded multi-GPU forests behind the C ABI.
//
// Trees are independent (createMulti maps over the IntMap of trees, src/Data/RPTree/Internal.hs:234-240; knn folds the
// per-tree candidates in ascending tree order, src/Data/RPTree.hs:174-176; recallWith is a mean over trees, :265-268), so
// a forest is partitioned in CONTIGUOUS blocks of trees over W GPUs with the data replicated.  Two ways to get there:
//   * rpf_create_multi(out, gpu_ids, n): ONE process drives n GPUs (what a Haskell host does).  The returned handle is a
//     parent over one sub-handle per GPU; every entry point of rpforest.h accepts it.  One host thread per GPU issues the
//     work, so the GPUs build / answer concurrently.
//   * rpf_comm_init_rank(h, world, rank, id): one process per GPU (torchrun); every process calls the same entry points
//     (SPMD) on its own single-device handle.
// Either way each rank's handle owns an NCCL communicator and the data-path exchanges happen INSIDE the engine, on the
// engine's own streams, without host synchronisation in between:
//   - replicated points: every rank uploads a 1/W slice of each row block over its own PCIe link, the slices are
//     all-gathered in place over NVLink while the projection kernel already runs on the blocks that are complete
//     (rpf_upload_rows, build.cu);
//   - knn: the local top-k lists are written straight into the rank's chunk of one packed buffer, ONE all-gather, k_merge;
//   - recallWith: brute-force truth sharded by query, all-gather, per-rank hit sums added in rank (= tree) order.
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy a host process already loaded -- e.g. PyTorch's -- is
// reused, and the library itself loads on machines without NCCL).
#include "rpf_internal.h"
#include <nccl.h>
#include <dlfcn.h>
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

// ---------------------------------------------------------------------------------------------------
// NCCL, resolved lazily
// ---------------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (api.so) break;
        }
        if (!api.so) { api.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char* n) { void* p = dlsym(api.so, n); if (!p && api.err.empty()) api.err = std::string("libnccl: missing symbol ") + n; return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return &api;
}
}  // namespace

struct RpfComm {
    int world = 1, rank = 0;
    ncclComm_t comm = nullptr;
};

int rpf_comm_rank(const rpf_handle* h) { return (h && h->comm) ? h->comm->rank : 0; }
int rpf_comm_world(const rpf_handle* h) { return (h && h->comm) ? h->comm->world : 1; }

void rpf_comm_free(rpf_handle* h) {
    if (!h || !h->comm) return;
    NcclApi* N = nccl_api();
    if (h->comm->comm && N->CommDestroy) N->CommDestroy(h->comm->comm);
    delete h->comm;
    h->comm = nullptr;
}

int rpf_comm_allgather(rpf_handle* h, void* buf, size_t bytes, cudaStream_t stream) {
    if (!h->comm || h->comm->world <= 1 || bytes == 0) return RPF_OK;
    NcclApi* N = nccl_api();
    ncclResult_t r = N->AllGather((const char*)buf + (size_t)h->comm->rank * bytes, buf, bytes, ncclChar, h->comm->comm, stream);
    if (r != ncclSuccess) return rpf_fail(h, RPF_ERR_CUDA, std::string("ncclAllGather: ") + N->GetErrorString(r));
    return RPF_OK;
}

// ---------------------------------------------------------------------------------------------------
// in-process group: one worker thread per GPU
// ---------------------------------------------------------------------------------------------------
struct RpfGroup {
    std::vector<rpf_handle*> subs;
    std::vector<int> t0, tl;                 // contiguous block of trees of every sub-handle
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    std::function<int(int)> job;
    uint64_t gen = 0;
    int pending = 0;
    bool quit = false;
    std::vector<int> rc;
    // forest-wide export sink (rpf_set_export_sink on the parent)
    double *sink_thr = nullptr, *sink_mlo = nullptr, *sink_mhi = nullptr; uint32_t* sink_perm = nullptr;

    void worker(int r) {
        cudaSetDevice(subs[r]->device);
        uint64_t seen = 0;
        for (;;) {
            std::function<int(int)> f;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_go.wait(lk, [&] { return quit || gen != seen; });
                if (quit) return;
                seen = gen;
                f = job;
            }
            int c;
            try { c = f(r); } catch (const std::bad_alloc&) { c = RPF_ERR_NOMEM; } catch (...) { c = RPF_ERR_CUDA; }
            {
                std::lock_guard<std::mutex> lk(mu);
                rc[r] = c;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    // runs f(r) on the worker of every GPU at once; first failing rank's code (its message is copied to the parent)
    int run(rpf_handle* parent, const std::function<int(int)>& f) {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = f; pending = (int)subs.size(); ++gen;
            std::fill(rc.begin(), rc.end(), 0);
        }
        cv_go.notify_all();
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return pending == 0; });
        }
        for (size_t r = 0; r < subs.size(); ++r)
            if (rc[r]) { parent->err = "gpu " + std::to_string(subs[r]->device) + ": " + subs[r]->err; return rc[r]; }
        return RPF_OK;
    }
};

void rpf_group_free(rpf_handle* h) {
    RpfGroup* G = h->group;
    if (!G) return;
    {
        std::lock_guard<std::mutex> lk(G->mu);
        G->quit = true;
    }
    G->cv_go.notify_all();
    for (auto& t : G->th) if (t.joinable()) t.join();
    for (rpf_handle* s : G->subs) rpf_destroy(s);     // also frees the sub-handle's communicator
    delete G;
    h->group = nullptr;
}

static void group_mirror_shape(rpf_handle* h) {
    RpfGroup* G = h->group;
    rpf_handle* s0 = G->subs[0];
    h->topo = s0->topo; h->n = s0->n; h->d = s0->d;
    h->stream_lost = s0->stream_lost;
    h->leaf_order_exact = true; h->built = true;
    for (rpf_handle* s : G->subs) { h->leaf_order_exact = h->leaf_order_exact && s->leaf_order_exact; h->built = h->built && s->built; }
}

static int group_of_tree(const RpfGroup* G, int t) {
    for (size_t r = 0; r < G->subs.size(); ++r) if (t >= G->t0[r] && t < G->t0[r] + G->tl[r]) return (int)r;
    return -1;
}

// body of a query-type call on one sub-handle (what the public entry point does around its *_impl)
template <typename F>
static int sub_call(rpf_handle* s, F&& f) {
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) return rpf_fail(s, RPF_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    s->call_begin();
    int rc = f();
    int rc2 = s->call_end();
    return rc ? rc : rc2;
}

// ---- group versions of the entry points (capi.cu forwards here when h->group is set) ------------------------------
int rpfg_set_hyperplanes(rpf_handle* h, int32_t T, int32_t maxDepth, const int64_t* off, const int32_t* idx, const double* val) {
    RpfGroup* G = h->group;
    const int W = (int)G->subs.size();
    if (T < W) return rpf_fail(h, RPF_ERR_ARG, "set_hyperplanes: fewer trees than GPUs (every GPU of a multi-GPU handle owns at least one tree)");
    const int64_t nrow = (int64_t)T * maxDepth, nnz = off[nrow];
    h->T = T; h->hpDepth = maxDepth;
    h->hp_off.assign(off, off + nrow + 1); h->hp_idx.assign(idx, idx + nnz); h->hp_val.assign(val, val + nnz);
    h->built = false;
    const int base = T / W, rem = T % W;      // balanced contiguous blocks: rank order == tree order
    for (int r = 0; r < W; ++r) { G->t0[r] = r * base + std::min(r, rem); G->tl[r] = base + (r < rem ? 1 : 0); }
    return G->run(h, [&](int r) {
        const int64_t a = (int64_t)G->t0[r] * maxDepth, b = a + (int64_t)G->tl[r] * maxDepth;
        std::vector<int64_t> o(off + a, off + b + 1);
        const int64_t lo = o[0];
        for (auto& x : o) x -= lo;
        return rpf_set_hyperplanes(G->subs[r], G->tl[r], maxDepth, o.data(), idx + lo, val + lo);
    });
}

int rpfg_after_gen_hyperplanes(rpf_handle* h) {     // h->hp_* hold the whole forest (drawn by capi.cu): shard them
    std::vector<int64_t> off = h->hp_off; std::vector<int32_t> idx = h->hp_idx; std::vector<double> val = h->hp_val;
    idx.push_back(0); val.push_back(0.0);           // non-NULL pointers for an all-empty forest
    return rpfg_set_hyperplanes(h, h->T, h->hpDepth, off.data(), idx.data(), val.data());
}

int rpfg_set_points(rpf_handle* h, const double* X, int64_t n, int32_t d) {
    RpfGroup* G = h->group;
    h->built = false;
    int rc = G->run(h, [&](int r) { return rpf_set_points(G->subs[r], X, n, d); });
    if (!rc) { h->n = n; h->d = d; }
    return rc;
}

int rpfg_set_points_sparse(rpf_handle* h, int64_t n, int32_t d, const int64_t* off, const int32_t* idx, const double* val) {
    RpfGroup* G = h->group;
    h->built = false;
    int rc = G->run(h, [&](int r) { return rpf_set_points_sparse(G->subs[r], n, d, off, idx, val); });
    if (!rc) { h->n = n; h->d = d; }
    return rc;
}

static void group_apply_sink(rpf_handle* h) {       // per-GPU slices of the forest-wide sink buffers
    RpfGroup* G = h->group;
    const size_t nn = (size_t)G->subs[0]->topo.nnodes(), n = (size_t)G->subs[0]->n;
    for (size_t r = 0; r < G->subs.size(); ++r) {
        const size_t t0 = (size_t)G->t0[r];
        if (G->sink_perm && nn > 0)
            rpf_set_export_sink(G->subs[r], G->sink_thr ? G->sink_thr + t0 * nn : nullptr, G->sink_mlo ? G->sink_mlo + t0 * nn : nullptr,
                                G->sink_mhi ? G->sink_mhi + t0 * nn : nullptr, G->sink_perm + t0 * n);
        else
            rpf_set_export_sink(G->subs[r], nullptr, nullptr, nullptr, nullptr);
    }
}

int rpfg_build(rpf_handle* h, int32_t maxDepth, int32_t minLeaf, int64_t chunk) {
    RpfGroup* G = h->group;
    h->built = false;
    int rc = G->run(h, [&](int r) {
        return chunk > 0 ? rpf_build_chunked(G->subs[r], maxDepth, minLeaf, chunk) : rpf_build(G->subs[r], maxDepth, minLeaf);
    });
    if (rc) return rc;
    group_mirror_shape(h);
    h->last_ms = 0; for (rpf_handle* s : G->subs) h->last_ms = std::max(h->last_ms, s->last_ms);
    return RPF_OK;
}

int rpfg_build_from_host(rpf_handle* h, const double* X, int64_t n, int32_t d, int32_t maxDepth, int32_t minLeaf) {
    RpfGroup* G = h->group;
    h->built = false;
    if (G->sink_perm) {
        // the sink slices depend on the shape of this build; the sub-handles compute the same topology the parent plans here
        Topology tp; build_topology(tp, n, maxDepth, minLeaf);
        const size_t nn = (size_t)tp.nnodes();
        for (size_t r = 0; r < G->subs.size(); ++r) {
            const size_t t0 = (size_t)G->t0[r];
            rpf_set_export_sink(G->subs[r], G->sink_thr ? G->sink_thr + t0 * nn : nullptr, G->sink_mlo ? G->sink_mlo + t0 * nn : nullptr,
                                G->sink_mhi ? G->sink_mhi + t0 * nn : nullptr, G->sink_perm + t0 * (size_t)n);
        }
    }
    int rc = G->run(h, [&](int r) { return rpf_build_from_host(G->subs[r], X, n, d, maxDepth, minLeaf); });
    if (rc) return rc;
    group_mirror_shape(h);
    h->last_ms = 0; for (rpf_handle* s : G->subs) h->last_ms = std::max(h->last_ms, s->last_ms);
    return RPF_OK;
}

int rpfg_tree_export(rpf_handle* h, int32_t t, double* thr, double* mlo, double* mhi, uint32_t* perm) {
    RpfGroup* G = h->group;
    const int r = group_of_tree(G, t);
    if (r < 0) return rpf_fail(h, RPF_ERR_ARG, "tree_export: tree index out of range");
    int rc = rpf_tree_export(G->subs[r], t - G->t0[r], thr, mlo, mhi, perm);
    if (rc) h->err = G->subs[r]->err;
    return rc;
}

int rpfg_forest_export(rpf_handle* h, double* thr, double* mlo, double* mhi, uint32_t* perm) {
    RpfGroup* G = h->group;
    const size_t nn = (size_t)h->topo.nnodes(), n = (size_t)h->n;
    return G->run(h, [&](int r) {
        const size_t t0 = (size_t)G->t0[r];
        return rpf_forest_export(G->subs[r], thr ? thr + t0 * nn : nullptr, mlo ? mlo + t0 * nn : nullptr, mhi ? mhi + t0 * nn : nullptr,
                                 perm ? perm + t0 * n : nullptr);
    });
}

int rpfg_set_export_sink(rpf_handle* h, double* thr, double* mlo, double* mhi, uint32_t* perm) {
    RpfGroup* G = h->group;
    G->sink_thr = thr; G->sink_mlo = mlo; G->sink_mhi = mhi; G->sink_perm = perm;
    if (h->built || !perm) group_apply_sink(h);      // otherwise applied by the next rpf_build_from_host (shape not known yet)
    return RPF_OK;
}

int rpfg_knn(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, int32_t dedup, double* dist, uint32_t* ids, int32_t* count) {
    RpfGroup* G = h->group;
    int rc = G->run(h, [&](int r) {
        rpf_handle* s = G->subs[r];
        return sub_call(s, [&] { return rpf_knn_impl(s, Q, q_last, nq, k, dedup, r == 0 ? dist : nullptr, r == 0 ? ids : nullptr, r == 0 ? count : nullptr, false); });
    });
    h->last_ms = 0; for (rpf_handle* s : G->subs) h->last_ms = std::max(h->last_ms, s->last_ms);
    return rc;
}

int rpfg_recall(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, double* recall_sum) {
    RpfGroup* G = h->group;
    int rc = G->run(h, [&](int r) {
        rpf_handle* s = G->subs[r];
        return sub_call(s, [&] { return rpf_recall_impl(s, Q, q_last, nq, k, r == 0 ? recall_sum : nullptr); });
    });
    h->last_ms = 0; for (rpf_handle* s : G->subs) h->last_ms = std::max(h->last_ms, s->last_ms);
    return rc;
}

int rpfg_brute_knn(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, double* dist, uint32_t* ids) {
    rpf_handle* s = h->group->subs[0];                // every GPU holds all the points
    int rc = rpf_brute_knn_s(s, Q, q_last, nq, k, dist, ids);
    if (rc) h->err = s->err;
    h->last_ms = s->last_ms;
    return rc;
}

// candidates over all trees: per-GPU CSR lists concatenated per query in GPU (= tree) order
int rpfg_candidates(rpf_handle* h, const double* Q, int64_t nq, int32_t t, int64_t* off_out, const int64_t* off_in, uint32_t* ids) {
    RpfGroup* G = h->group;
    if (t >= 0) {
        const int r = group_of_tree(G, t);
        if (r < 0) return rpf_fail(h, RPF_ERR_ARG, "candidates: tree index out of range");
        int rc = off_out ? rpf_candidates_count(G->subs[r], Q, nq, t - G->t0[r], off_out) : rpf_candidates(G->subs[r], Q, nq, t - G->t0[r], off_in, ids);
        if (rc) h->err = G->subs[r]->err;
        return rc;
    }
    const int W = (int)G->subs.size();
    std::vector<std::vector<int64_t>> off(W, std::vector<int64_t>((size_t)nq + 1));
    int rc = G->run(h, [&](int r) { return rpf_candidates_count(G->subs[r], Q, nq, -1, off[r].data()); });
    if (rc) return rc;
    if (off_out) {
        for (int64_t q = 0; q <= nq; ++q) { int64_t s = 0; for (int r = 0; r < W; ++r) s += off[r][q]; off_out[q] = s; }
        return RPF_OK;
    }
    std::vector<std::vector<uint32_t>> loc(W);
    for (int r = 0; r < W; ++r) loc[r].resize((size_t)std::max<int64_t>(off[r][nq], 1));
    rc = G->run(h, [&](int r) { return rpf_candidates(G->subs[r], Q, nq, -1, off[r].data(), loc[r].data()); });
    if (rc) return rc;
    for (int64_t q = 0; q < nq; ++q) {
        int64_t w = off_in[q];
        for (int r = 0; r < W; ++r) {
            const int64_t a = off[r][q], b = off[r][q + 1];
            if (w + (b - a) > off_in[q + 1]) return rpf_fail(h, RPF_ERR_ARG, "candidates: offsets do not match rpf_candidates_count");
            if (b > a) std::memcpy(ids + w, loc[r].data() + a, (size_t)(b - a) * 4);
            w += b - a;
        }
    }
    return RPF_OK;
}

int rpfg_forest_save(rpf_handle* h, const char* path, int32_t with_points) {
    RpfGroup* G = h->group;
    const int W = (int)G->subs.size();
    return G->run(h, [&](int r) {
        const std::string p = std::string(path) + ".gpu" + std::to_string(r) + "of" + std::to_string(W);
        return rpf_forest_save(G->subs[r], p.c_str(), with_points);
    });
}

int rpfg_forest_load(rpf_handle* h, const char* path) {
    RpfGroup* G = h->group;
    const int W = (int)G->subs.size();
    h->built = false;
    int rc = G->run(h, [&](int r) {
        const std::string p = std::string(path) + ".gpu" + std::to_string(r) + "of" + std::to_string(W);
        return rpf_forest_load(G->subs[r], p.c_str());
    });
    if (rc) return rc;
    // rebuild the parent's view: tree blocks and the forest-wide hyperplane CSR from the sub-handles
    h->T = 0; h->hpDepth = G->subs[0]->hpDepth;
    h->hp_off.assign(1, 0); h->hp_idx.clear(); h->hp_val.clear();
    for (int r = 0; r < W; ++r) {
        rpf_handle* s = G->subs[r];
        if (s->hpDepth != h->hpDepth || s->n != G->subs[0]->n || s->d != G->subs[0]->d || s->topo.nnodes() != G->subs[0]->topo.nnodes())
            return rpf_fail(h, RPF_ERR_ARG, "forest_load: the per-GPU checkpoints do not belong to one forest");
        G->t0[r] = h->T; G->tl[r] = s->T; h->T += s->T;
        const int64_t base = h->hp_off.back();
        for (size_t i = 1; i < s->hp_off.size(); ++i) h->hp_off.push_back(base + s->hp_off[i]);
        h->hp_idx.insert(h->hp_idx.end(), s->hp_idx.begin(), s->hp_idx.end());
        h->hp_val.insert(h->hp_val.end(), s->hp_val.begin(), s->hp_val.end());
    }
    group_mirror_shape(h);
    return RPF_OK;
}

int rpfg_set_option(rpf_handle* h, const char* name, int64_t value, bool is_cap) {
    RpfGroup* G = h->group;
    for (rpf_handle* s : G->subs) {
        int rc = is_cap ? rpf_set_bottom_cap(s, (int32_t)value) : rpf_set_option(s, name, value);
        if (rc) { h->err = s->err; return rc; }
    }
    return RPF_OK;
}

int rpfg_set_profiling(rpf_handle* h, int on) {
    for (rpf_handle* s : h->group->subs) rpf_set_profiling(s, on);
    return RPF_OK;
}

int rpfg_get_profile(const rpf_handle* h, double* ms, int64_t* launches, int cap) {
    for (int i = 0; i < PH_COUNT && i < cap; ++i) {
        double m = 0; int64_t l = 0;
        for (const rpf_handle* s : h->group->subs) { m = std::max(m, s->phase_ms[i]); l += s->phase_launches[i]; }
        if (ms) ms[i] = m;                           // the GPUs run concurrently: slowest GPU per phase, launches of all
        if (launches) launches[i] = l;
    }
    return PH_COUNT;
}

int64_t rpfg_launch_count(const rpf_handle* h) {
    int64_t l = 0;
    for (const rpf_handle* s : h->group->subs) l += s->launches;
    return l;
}

extern "C" {

int rpf_num_gpus(const rpf_handle* h) {
    if (!h) return -1;
    if (h->group) return (int)h->group->subs.size();
    return rpf_comm_world(h);
}

int rpf_comm_unique_id(void* id128) {
    if (!id128) return RPF_ERR_ARG;
    NcclApi* N = nccl_api();
    if (!N->GetUniqueId) return RPF_ERR_UNSUPPORTED;
    ncclUniqueId id;
    if (N->GetUniqueId(&id) != ncclSuccess) return RPF_ERR_CUDA;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(id128, &id, 128);
    return RPF_OK;
}

int rpf_comm_init_rank(rpf_handle* h, int32_t world, int32_t rank, const void* id128) {
    if (!h) return RPF_ERR_ARG;
    if (h->group) return rpf_fail(h, RPF_ERR_ARG, "comm_init_rank: the handle already drives several GPUs (rpf_create_multi)");
    if (world < 1 || rank < 0 || rank >= world || !id128) return rpf_fail(h, RPF_ERR_ARG, "comm_init_rank: bad world / rank / id");
    if (h->dX || h->built) return rpf_fail(h, RPF_ERR_STATE, "comm_init_rank: call it before any points are set");
    NcclApi* N = nccl_api();
    if (!N->CommInitRank) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "comm_init_rank: " + N->err);
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return rpf_fail(h, RPF_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    rpf_comm_free(h);
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    RpfComm* c = new RpfComm();
    c->world = world; c->rank = rank;
    ncclResult_t r = N->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) { delete c; return rpf_fail(h, RPF_ERR_CUDA, std::string("ncclCommInitRank: ") + N->GetErrorString(r)); }
    h->comm = c;
    ++h->cfg_epoch;
    return RPF_OK;
}

int rpf_create_multi(rpf_handle** out, const int* gpu_ids, int n_gpus) {
    if (!out) return RPF_ERR_ARG;
    *out = nullptr;
    if (n_gpus < 1 || !gpu_ids) return RPF_ERR_ARG;
    if (n_gpus == 1) return rpf_create(out, gpu_ids[0]);
    for (int i = 0; i < n_gpus; ++i) for (int j = 0; j < i; ++j) if (gpu_ids[i] == gpu_ids[j]) return RPF_ERR_ARG;
    NcclApi* N = nccl_api();
    if (!N->CommInitAll) return RPF_ERR_UNSUPPORTED;
    RpfGroup* G = new RpfGroup();
    bool ok = true;
    for (int i = 0; i < n_gpus && ok; ++i) {
        rpf_handle* s = nullptr;
        ok = rpf_create(&s, gpu_ids[i]) == RPF_OK;
        if (ok) G->subs.push_back(s);
    }
    std::vector<ncclComm_t> comms((size_t)n_gpus, nullptr);
    if (ok) ok = N->CommInitAll(comms.data(), n_gpus, gpu_ids) == ncclSuccess;
    if (!ok) {
        for (rpf_handle* s : G->subs) rpf_destroy(s);
        delete G;
        return RPF_ERR_CUDA;
    }
    for (int i = 0; i < n_gpus; ++i) {
        RpfComm* c = new RpfComm();
        c->world = n_gpus; c->rank = i; c->comm = comms[i];
        G->subs[i]->comm = c;
    }
    G->t0.assign((size_t)n_gpus, 0); G->tl.assign((size_t)n_gpus, 0); G->rc.assign((size_t)n_gpus, 0);
    for (int i = 0; i < n_gpus; ++i) G->th.emplace_back([G, i] { G->worker(i); });
    rpf_handle* h = new rpf_handle();
    h->device = gpu_ids[0];
    h->group = G;
    *out = h;
    return RPF_OK;
}

}  // extern "C"
