// capi.cu -- the C ABI declared in include/rpforest.h: handle lifecycle, data upload, the data-independent
// topology, hyperplane regeneration (SplitMix64), result export and the measurement hooks.
#include "rpf_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdio>

int rpf_fail(rpf_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

// ---------------------------------------------------------------------------------------------------
// profiling / timing
// ---------------------------------------------------------------------------------------------------
static cudaEvent_t get_event(rpf_handle* h) {
    if (!h->event_pool.empty()) { cudaEvent_t e = h->event_pool.back(); h->event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void rpf_handle::prof_reset() {
    for (int i = 0; i < PH_COUNT; ++i) { phase_ms[i] = 0; phase_launches[i] = 0; }
}
void rpf_handle::prof_begin(int phase) {
    ++launches;
    ++phase_launches[phase];
    if (!profiling) return;
    ProfEvent pe; pe.phase = phase; pe.a = get_event(this); pe.b = get_event(this);
    cudaEventRecord(pe.a, stream);
    pending.push_back(pe);
}
void rpf_handle::prof_end(int) {
    if (!profiling) return;
    cudaEventRecord(pending.back().b, stream);
}
void rpf_handle::call_begin() {
    prof_reset();
    cudaEventRecord(ev_begin, stream);
}
int rpf_handle::call_end() {
    cudaEventRecord(ev_end, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return rpf_fail(this, RPF_ERR_CUDA, std::string("stream sync: ") + cudaGetErrorString(e));
    float ms = 0; cudaEventElapsedTime(&ms, ev_begin, ev_end); last_ms = ms;
    for (auto& pe : pending) {
        float m = 0; cudaEventElapsedTime(&m, pe.a, pe.b);
        phase_ms[pe.phase] += m;
        event_pool.push_back(pe.a); event_pool.push_back(pe.b);
    }
    pending.clear();
    return RPF_OK;
}

void* rpf_handle::ws_get(int slot, size_t bytes) {
    WsBuf& b = ws[slot];
    if (bytes < 16) bytes = 16;
    if (b.cap >= bytes) return b.p;
    if (capturing) { err = "workspace growth during graph capture"; return nullptr; }
    ++cfg_epoch;
    if (b.p) { cudaStreamSynchronize(stream); cudaFree(b.p); ws_bytes -= b.cap; b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 16;          // a little slack so slowly growing requests do not thrash
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&b.p, want); }
    if (e != cudaSuccess) { b.p = nullptr; err = std::string("workspace cudaMalloc: ") + cudaGetErrorString(e); cudaGetLastError(); return nullptr; }
    b.cap = want; ws_bytes += want;
    return b.p;
}
void rpf_handle::ws_free_all() {
    ++cfg_epoch;
    for (int i = 0; i < WS_COUNT; ++i) if (ws[i].p) { cudaFree(ws[i].p); ws[i].p = nullptr; ws[i].cap = 0; }
    ws_bytes = 0;
}

// ---------------------------------------------------------------------------------------------------
// staging ring (page-locked host tables -> device, one async copy per flush)
// ---------------------------------------------------------------------------------------------------
int rpf_handle::stage_begin(size_t bytes) {
    stage_cur = (stage_cur + 1) % RPF_STAGE_SLOTS;
    StageSlot& S = stage[stage_cur];
    if (!S.ev && cudaEventCreateWithFlags(&S.ev, cudaEventDisableTiming) != cudaSuccess) return rpf_fail(this, RPF_ERR_CUDA, "stage: event");
    if (S.pending) { cudaEventSynchronize(S.ev); S.pending = false; }      // the slot's previous upload has left host memory
    bytes += 4096;
    if (S.cap < bytes) {
        // Grow EVERY slot now (one burst of page-locked allocations on the first build of a shape) instead of one slot per
        // call, which would put an allocation into each of the next RPF_STAGE_SLOTS builds.  Kernels of earlier jobs may
        // still read the old device blocks: synchronise first (growth is rare).
        cudaStreamSynchronize(stream);
        const size_t want = bytes + bytes / 4;
        for (auto& Q : stage) {
            if (Q.cap >= want) continue;
            if (Q.h) cudaFreeHost(Q.h);
            if (Q.d) cudaFree(Q.d);
            Q.h = nullptr; Q.d = nullptr; Q.cap = 0; Q.pending = false;
            if (cudaMallocHost(&Q.h, want) != cudaSuccess || cudaMalloc(&Q.d, want) != cudaSuccess) {
                cudaGetLastError();
                return rpf_fail(this, RPF_ERR_NOMEM, "stage: allocation failed");
            }
            Q.cap = want;
        }
    }
    stage_off = 0; stage_flushed = 0;
    return RPF_OK;
}
void* rpf_handle::stage_put_raw(const void* src, size_t bytes) {
    StageSlot& S = stage[stage_cur];
    const size_t off = (stage_off + 255) & ~(size_t)255;
    if (off + bytes > S.cap) { err = "stage: overflow (internal sizing error)"; return nullptr; }
    if (bytes) std::memcpy(S.h + off, src, bytes);
    stage_off = off + bytes;
    return S.d + off;
}
int rpf_handle::stage_flush() {
    StageSlot& S = stage[stage_cur];
    if (stage_off > stage_flushed) {
        RPF_CUDA(this, cudaMemcpyAsync(S.d + stage_flushed, S.h + stage_flushed, stage_off - stage_flushed, cudaMemcpyHostToDevice, stream));
        RPF_CUDA(this, cudaEventRecord(S.ev, stream));
        S.pending = true;
        stage_flushed = stage_off;
    }
    return RPF_OK;
}
void rpf_handle::stage_free_all() {
    for (auto& S : stage) {
        if (S.h) cudaFreeHost(S.h);
        if (S.d) cudaFree(S.d);
        if (S.ev) cudaEventDestroy(S.ev);
        S = StageSlot();
    }
    stage_cur = -1;
}

static const char* kPhaseNames[PH_COUNT] = {
    "project", "top_hist", "top_pick", "top_compact", "top_finish", "top_ties", "top_relabel", "bottom",
    "q_project", "q_traverse", "q_knn", "q_candidates", "truth", "recall", "merge", "stream", "stream_concat", "misc"};

// ---------------------------------------------------------------------------------------------------
// topology: Internal.hs:289 (Tip iff ixLev >= maxDepth || length xs' <= minLeaf), :495/:503 (nh = n div 2)
// ---------------------------------------------------------------------------------------------------
void build_topology(Topology& tp, int64_t n, int maxDepth, int minLeaf) {
    tp = Topology();
    tp.n = n; tp.maxDepth = maxDepth; tp.minLeaf = minLeaf;
    tp.start.push_back(0); tp.size.push_back((uint32_t)n); tp.child.push_back(-1); tp.depth.push_back(0);
    tp.level_off.push_back(0);
    int64_t lo = 0, hi = 1;
    int lev = 0;
    while (lo < hi) {
        uint32_t mx = 0; bool any_internal = false;
        for (int64_t g = lo; g < hi; ++g) {
            const uint32_t sz = tp.size[g];
            mx = std::max(mx, sz);
            const bool leaf = lev >= maxDepth || (int64_t)sz <= (int64_t)minLeaf;
            if (!leaf) {
                any_internal = true;
                const uint32_t nh = sz / 2;
                tp.child[g] = (int32_t)tp.start.size();
                tp.start.push_back(tp.start[g]);      tp.size.push_back(nh);      tp.child.push_back(-1); tp.depth.push_back(lev + 1);
                tp.start.push_back(tp.start[g] + nh); tp.size.push_back(sz - nh); tp.child.push_back(-1); tp.depth.push_back(lev + 1);
            }
        }
        tp.lvl_maxsize.push_back(mx);
        tp.level_off.push_back(hi);
        ++lev;
        if (any_internal) tp.L_eff = lev;
        lo = hi; hi = (int64_t)tp.start.size();
    }
    tp.nlevels = lev;
}

// ---------------------------------------------------------------------------------------------------
// SplitMix64 + sparse hyperplane sampler (host side; replaces Gen.hs:148-195 under Batch.hs:59-61).
// splitmix: mkSMGen s = SMGen (mix64 s) (mixGamma (s + goldenGamma)); nextWord64 advances seed by gamma.
// ---------------------------------------------------------------------------------------------------
namespace {
struct SMGen { uint64_t seed, gamma; };
inline uint64_t sx(int n, uint64_t w) { return w ^ (w >> n); }
inline uint64_t mix64(uint64_t z) { z = sx(33, z) * 0xff51afd7ed558ccdULL; z = sx(33, z) * 0xc4ceb9fe1a85ec53ULL; return sx(33, z); }
inline uint64_t mix_gamma(uint64_t z) {
    z = sx(30, z) * 0xbf58476d1ce4e5b9ULL; z = sx(27, z) * 0x94d049bb133111ebULL; z = sx(31, z) | 1ULL;
    return __builtin_popcountll(z ^ (z >> 1)) >= 24 ? z : z ^ 0xaaaaaaaaaaaaaaaaULL;
}
inline SMGen mk_smgen(uint64_t s) { return SMGen{mix64(s), mix_gamma(s + 0x9e3779b97f4a7c15ULL)}; }
inline double next_double(SMGen& g) { g.seed += g.gamma; return (double)(mix64(g.seed) >> 11) * 0x1.0p-53; }
// inverse normal CDF: Acklam's rational approximation + one Halley step (erf package's invnormcdf, as
// recalled; UNVERIFIED against Hackage -- use rpf_set_hyperplanes for bit-exact parity with a Haskell host).
double inv_norm_cdf(double p) {
    static const double a[6] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02, 1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[5] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02, 6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00, -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double dd[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
    if (p == 0) return -INFINITY;
    if (p == 1) return INFINITY;
    const double plow = 0.02425, phigh = 1 - plow;
    double x;
    if (p < plow) { double q = std::sqrt(-2 * std::log(p)); x = (((((c[0]*q+c[1])*q+c[2])*q+c[3])*q+c[4])*q+c[5]) / ((((dd[0]*q+dd[1])*q+dd[2])*q+dd[3])*q+1); }
    else if (p <= phigh) { double q = p - 0.5, r = q*q; x = (((((a[0]*r+a[1])*r+a[2])*r+a[3])*r+a[4])*r+a[5])*q / (((((b[0]*r+b[1])*r+b[2])*r+b[3])*r+b[4])*r+1); }
    else { double q = std::sqrt(-2 * std::log(1 - p)); x = -(((((c[0]*q+c[1])*q+c[2])*q+c[3])*q+c[4])*q+c[5]) / ((((dd[0]*q+dd[1])*q+dd[2])*q+dd[3])*q+1); }
    const double e = 0.5 * std::erfc(-x / std::sqrt(2.0)) - p;
    const double u = e * std::sqrt(2 * M_PI) * std::exp(x * x / 2);
    return x - u / (1 + x * u / 2);
}
}  // namespace

static void free_hp_dev(rpf_handle* h) {
    if (h->d_hp_off) cudaFree(h->d_hp_off);
    if (h->d_hp_idx) cudaFree(h->d_hp_idx);
    if (h->d_hp_val) cudaFree(h->d_hp_val);
    if (h->d_hp_pack) cudaFree(h->d_hp_pack);
    if (h->d_hp_chunk) cudaFree(h->d_hp_chunk);
    h->d_hp_chunk = nullptr; h->hp_chunk_rows = 0;
    if (h->proj_progs && h->proj_progs_free) { cudaStreamSynchronize(h->stream); h->proj_progs_free(h->proj_progs); }
    h->proj_progs = nullptr;
    h->d_hp_off = nullptr; h->d_hp_idx = nullptr; h->d_hp_val = nullptr; h->d_hp_pack = nullptr;
}
static void free_topo_dev(rpf_handle* h) {
    if (h->d_node_start) cudaFree(h->d_node_start);
    if (h->d_node_size) cudaFree(h->d_node_size);
    if (h->d_node_child) cudaFree(h->d_node_child);
    if (h->d_node_depth) cudaFree(h->d_node_depth);
    h->d_node_start = h->d_node_size = nullptr; h->d_node_child = h->d_node_depth = nullptr;
    h->topo_dev_nn = 0; h->topo_dev_child.clear(); h->topo_dev_size.clear();
}
static void free_forest_dev(rpf_handle* h) {
    if (h->d_thr) cudaFree(h->d_thr);
    if (h->d_mlo) cudaFree(h->d_mlo);
    if (h->d_mhi) cudaFree(h->d_mhi);
    if (h->d_perm) cudaFree(h->d_perm);
    h->d_thr = h->d_mlo = h->d_mhi = nullptr; h->d_perm = nullptr; h->built = false; h->sink_pending = false;
    h->res_node_bytes = h->res_perm_bytes = 0;
}

static int upload_hyperplanes(rpf_handle* h) {
    ++h->cfg_epoch;
    free_hp_dev(h);
    free_forest_dev(h);
    const size_t nrow = h->hp_off.size(), nnz = h->hp_idx.size();
    RPF_CUDA(h, cudaMalloc(&h->d_hp_off, nrow * 8));
    RPF_CUDA(h, cudaMalloc(&h->d_hp_idx, std::max<size_t>(nnz, 1) * 4));
    RPF_CUDA(h, cudaMalloc(&h->d_hp_val, std::max<size_t>(nnz, 1) * 8));
    RPF_CUDA(h, cudaMemcpy(h->d_hp_off, h->hp_off.data(), nrow * 8, cudaMemcpyHostToDevice));
    RPF_CUDA(h, cudaMalloc(&h->d_hp_pack, std::max<size_t>(nnz, 1) * 16));
    if (nnz) {
        RPF_CUDA(h, cudaMemcpy(h->d_hp_idx, h->hp_idx.data(), nnz * 4, cudaMemcpyHostToDevice));
        RPF_CUDA(h, cudaMemcpy(h->d_hp_val, h->hp_val.data(), nnz * 8, cudaMemcpyHostToDevice));
        std::vector<double> pack(nnz * 2);
        for (size_t q = 0; q < nnz; ++q) {
            pack[2 * q] = h->hp_val[q];
            const int64_t ib = (int64_t)h->hp_idx[q] * 8;
            std::memcpy(&pack[2 * q + 1], &ib, 8);      // byte offset of the component (8 * index) in the bit pattern of the second double
        }
        RPF_CUDA(h, cudaMemcpy(h->d_hp_pack, pack.data(), nnz * 16, cudaMemcpyHostToDevice));
    }
    return RPF_OK;
}

#define RPF_SETDEV(h) RPF_CUDA(h, cudaSetDevice((h)->device))

// group parent (rpf_create_multi, multi.cu): every entry point forwards to the per-GPU sub-handles
int rpfg_set_hyperplanes(rpf_handle* h, int32_t T, int32_t maxDepth, const int64_t* off, const int32_t* idx, const double* val);
int rpfg_after_gen_hyperplanes(rpf_handle* h);
int rpfg_set_points(rpf_handle* h, const double* X, int64_t n, int32_t d);
int rpfg_set_points_sparse(rpf_handle* h, int64_t n, int32_t d, const int64_t* off, const int32_t* idx, const double* val);
int rpfg_build(rpf_handle* h, int32_t maxDepth, int32_t minLeaf, int64_t chunk);
int rpfg_build_from_host(rpf_handle* h, const double* X, int64_t n, int32_t d, int32_t maxDepth, int32_t minLeaf);
int rpfg_tree_export(rpf_handle* h, int32_t t, double* thr, double* mlo, double* mhi, uint32_t* perm);
int rpfg_forest_export(rpf_handle* h, double* thr, double* mlo, double* mhi, uint32_t* perm);
int rpfg_set_export_sink(rpf_handle* h, double* thr, double* mlo, double* mhi, uint32_t* perm);
int rpfg_knn(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, int32_t dedup, double* dist, uint32_t* ids, int32_t* count);
int rpfg_recall(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, double* recall_sum);
int rpfg_brute_knn(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, double* dist, uint32_t* ids);
int rpfg_candidates(rpf_handle* h, const double* Q, int64_t nq, int32_t t, int64_t* off_out, const int64_t* off_in, uint32_t* ids);
int rpfg_forest_save(rpf_handle* h, const char* path, int32_t with_points);
int rpfg_forest_load(rpf_handle* h, const char* path);
int rpfg_set_option(rpf_handle* h, const char* name, int64_t value, bool is_cap);
int rpfg_set_profiling(rpf_handle* h, int on);
int rpfg_get_profile(const rpf_handle* h, double* ms, int64_t* launches, int cap);
int64_t rpfg_launch_count(const rpf_handle* h);

// device copies of h->topo (start, size, child, depth per BFS node)
int rpf_upload_topology(rpf_handle* h) {
    const Topology& tp = h->topo;
    const size_t nn = (size_t)tp.nnodes();
    // the device copy is kept while the shape stays the same (a rebuild of the same shape does no allocation and no copy)
    if (h->d_node_start && h->topo_dev_nn == nn && h->topo_dev_child == tp.child && h->topo_dev_size == tp.size) return RPF_OK;
    free_topo_dev(h);
    ++h->cfg_epoch;
    RPF_CUDA(h, cudaMalloc(&h->d_node_start, nn * 4));
    RPF_CUDA(h, cudaMalloc(&h->d_node_size, nn * 4));
    RPF_CUDA(h, cudaMalloc(&h->d_node_child, nn * 4));
    RPF_CUDA(h, cudaMalloc(&h->d_node_depth, nn * 4));
    RPF_CUDA(h, cudaMemcpy(h->d_node_start, tp.start.data(), nn * 4, cudaMemcpyHostToDevice));
    RPF_CUDA(h, cudaMemcpy(h->d_node_size, tp.size.data(), nn * 4, cudaMemcpyHostToDevice));
    RPF_CUDA(h, cudaMemcpy(h->d_node_child, tp.child.data(), nn * 4, cudaMemcpyHostToDevice));
    RPF_CUDA(h, cudaMemcpy(h->d_node_depth, tp.depth.data(), nn * 4, cudaMemcpyHostToDevice));
    h->topo_dev_nn = nn; h->topo_dev_child = tp.child; h->topo_dev_size = tp.size;
    return RPF_OK;
}

namespace {
struct CkptHeader {
    char magic[8];              // "RPFB200\0"
    uint32_t version, flags;    // flags bit 0: points included, bit 1: SVector data (xlast present)
    int64_t n, nn, hp_nnz, lost;
    int32_t d, T, hpDepth, maxDepth, minLeaf, nlevels, L_eff, leaf_order_exact;
};
template <typename T> bool wr(FILE* f, const T* p, size_t n) { return n == 0 || fwrite(p, sizeof(T), n, f) == n; }
template <typename T> bool rd(FILE* f, T* p, size_t n) { return n == 0 || fread(p, sizeof(T), n, f) == n; }
template <typename T> bool wr_dev(FILE* f, const T* d, size_t n, std::vector<char>& tmp) {
    if (n == 0) return true;
    tmp.resize(n * sizeof(T));
    if (cudaMemcpy(tmp.data(), d, n * sizeof(T), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
    return fwrite(tmp.data(), sizeof(T), n, f) == n;
}
template <typename T> bool rd_dev(FILE* f, T* d, size_t n, std::vector<char>& tmp) {
    if (n == 0) return true;
    tmp.resize(n * sizeof(T));
    if (fread(tmp.data(), sizeof(T), n, f) != n) return false;
    return cudaMemcpy(d, tmp.data(), n * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}
}  // namespace

extern "C" {

int rpf_abi_version(void) { return 1; }

int rpf_create(rpf_handle** out, int device) {
    if (!out) return RPF_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return RPF_ERR_CUDA;   // no CPU fallback
    if (device < 0 || device >= ndev) return RPF_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return RPF_ERR_CUDA;
    rpf_handle* h = new rpf_handle();
    h->device = device;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev_begin) != cudaSuccess || cudaEventCreate(&h->ev_end) != cudaSuccess) {
        delete h;
        return RPF_ERR_CUDA;
    }
    *out = h;
    return RPF_OK;
}

void rpf_destroy(rpf_handle* h) {
    if (!h) return;
    if (h->group) { rpf_group_free(h); delete h; return; }
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    rpf_insert_drop(h);
    if (h->ownX && h->dX) cudaFree((void*)h->dX);
    if (h->d_xlast) cudaFree(h->d_xlast);
    if (h->dX32) cudaFree(h->dX32);
    if (h->d_xmax) cudaFree(h->d_xmax);
    free_hp_dev(h); free_topo_dev(h); free_forest_dev(h);
    if (h->stream_plan && h->stream_plan_free) h->stream_plan_free(h->stream_plan);
    if (h->build_graph) cudaGraphExecDestroy(h->build_graph);
    if (h->batch_plan && h->batch_plan_free) h->batch_plan_free(h->batch_plan);
    for (auto e : h->branch_ev) if (e) cudaEventDestroy(e);
    for (auto st : h->branch_stream) if (st) cudaStreamDestroy(st);
    for (auto e : h->copy_ev) if (e) cudaEventDestroy(e);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (auto e : h->up_ev) if (e) cudaEventDestroy(e);
    if (h->gather_stream) cudaStreamDestroy(h->gather_stream);
    for (auto e : h->sink_ev) if (e) cudaEventDestroy(e);
    if (h->d2h_stream) { cudaStreamSynchronize(h->d2h_stream); cudaStreamDestroy(h->d2h_stream); }
    h->ws_free_all();
    h->stage_free_all();
    for (auto e : h->event_pool) cudaEventDestroy(e);
    if (h->ev_begin) cudaEventDestroy(h->ev_begin);
    if (h->ev_end) cudaEventDestroy(h->ev_end);
    cudaStreamDestroy(h->stream);
    rpf_comm_free(h);
    delete h;
}

const char* rpf_last_error(const rpf_handle* h) { return h ? h->err.c_str() : "null handle"; }

int rpf_set_points(rpf_handle* h, const double* X, int64_t n, int32_t d) {
    if (!h) return RPF_ERR_ARG;
    if (n < 0 || d < 1 || (n > 0 && !X)) return rpf_fail(h, RPF_ERR_ARG, "set_points: bad n/d/X");
    if (n >= (int64_t)1 << 31) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "set_points: n must be < 2^31");
    if (h->group) return rpfg_set_points(h, X, n, d);
    RPF_SETDEV(h);
    rpf_insert_drop(h);
    ++h->cfg_epoch; ++h->x_version;
    if (h->d_xlast) { cudaFree(h->d_xlast); h->d_xlast = nullptr; }
    double* p = nullptr;
    const int64_t pad = rpf_comm_world(h) > 1 ? rpf_comm_world(h) : 0;      // spill rows of the in-place all-gather
    const size_t bytes = std::max<size_t>((size_t)(n + pad) * d * 8, 16);
    if (h->ownX && h->dX && h->x_bytes == bytes) {
        p = (double*)h->dX;                       // same footprint: reuse the device buffer
        h->built = false; h->sink_pending = false;
    } else {
        if (h->ownX && h->dX) cudaFree((void*)h->dX);
        h->dX = nullptr; h->ownX = false;
        free_forest_dev(h);
        RPF_CUDA(h, cudaMalloc(&p, bytes));
        h->x_bytes = bytes;
    }
    h->dX = p; h->ownX = true; h->n = n; h->d = d; h->x_pad_rows = pad;
    // rank of a tree-sharded forest: only rows [rank * per, (rank + 1) * per), per = ceil(n / world), are read from X; the
    // rest arrives over NVLink (rpf_upload_rows)
    int rcu = n > 0 ? rpf_upload_rows(h, X, 0, n, h->stream) : RPF_OK;
    if (!rcu && cudaStreamSynchronize(h->stream) != cudaSuccess) rcu = rpf_fail(h, RPF_ERR_CUDA, std::string("set_points: ") + cudaGetErrorString(cudaGetLastError()));
    if (rcu) {           // no half-filled buffer is left behind
        cudaFree(p); h->dX = nullptr; h->ownX = false; h->n = 0; h->x_bytes = 0;
        return rcu;
    }
    return RPF_OK;
}

int rpf_set_points_device(rpf_handle* h, const double* X_dev, int64_t n, int32_t d) {
    if (!h) return RPF_ERR_ARG;
    if (n < 0 || d < 1 || (n > 0 && !X_dev)) return rpf_fail(h, RPF_ERR_ARG, "set_points_device: bad n/d/X");
    if (n >= (int64_t)1 << 31) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "set_points_device: n must be < 2^31");
    if (h->group) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "set_points_device: a multi-GPU handle replicates the points itself (rpf_set_points)");
    RPF_SETDEV(h);
    rpf_insert_drop(h);
    ++h->cfg_epoch; ++h->x_version;
    if (h->d_xlast) { cudaFree(h->d_xlast); h->d_xlast = nullptr; }
    if (h->ownX && h->dX) cudaFree((void*)h->dX);
    free_forest_dev(h);
    h->dX = X_dev; h->ownX = false; h->n = n; h->d = d; h->x_pad_rows = 0;
    return RPF_OK;
}

// CSR rows -> dense n x d image (zero elsewhere) + last stored component per row
__global__ void k_densify(const int64_t* __restrict__ off, const int32_t* __restrict__ idx, const double* __restrict__ val,
                          int64_t n, int d, double* __restrict__ X, int32_t* __restrict__ xlast) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t a = off[i], b = off[i + 1];
    for (int64_t q = a; q < b; ++q) X[i * d + idx[q]] = val[q];
    xlast[i] = b > a ? idx[b - 1] : -1;
}

static int check_csr_rows(rpf_handle* h, int64_t n, int32_t d, const int64_t* off, const int32_t* idx, const char* what) {
    if (off[0] != 0) return rpf_fail(h, RPF_ERR_ARG, std::string(what) + ": off[0] must be 0");
    for (int64_t i = 0; i < n; ++i) {
        if (off[i + 1] < off[i]) return rpf_fail(h, RPF_ERR_ARG, std::string(what) + ": offsets not monotone");
        for (int64_t q = off[i]; q < off[i + 1]; ++q) {
            if (idx[q] < 0 || idx[q] >= d) return rpf_fail(h, RPF_ERR_ARG, std::string(what) + ": component index out of range");
            if (q > off[i] && idx[q] <= idx[q - 1]) return rpf_fail(h, RPF_ERR_ARG, std::string(what) + ": component indices must be strictly ascending per row");
        }
    }
    return RPF_OK;
}

int rpf_set_points_sparse(rpf_handle* h, int64_t n, int32_t d, const int64_t* off, const int32_t* idx, const double* val) {
    if (!h) return RPF_ERR_ARG;
    if (n < 0 || d < 1 || !off || (off[n] > 0 && (!idx || !val))) return rpf_fail(h, RPF_ERR_ARG, "set_points_sparse: bad n/d/CSR");
    if (n >= (int64_t)1 << 31) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "set_points_sparse: n must be < 2^31");
    int rc = check_csr_rows(h, n, d, off, idx, "set_points_sparse");
    if (rc) return rc;
    if (h->group) return rpfg_set_points_sparse(h, n, d, off, idx, val);
    RPF_SETDEV(h);
    rpf_insert_drop(h);
    ++h->cfg_epoch; ++h->x_version;
    if (h->d_xlast) { cudaFree(h->d_xlast); h->d_xlast = nullptr; }
    if (h->ownX && h->dX) cudaFree((void*)h->dX);
    h->dX = nullptr; h->ownX = false; h->x_bytes = 0;
    free_forest_dev(h);
    const int64_t nnz = off[n];
    const size_t bytes = std::max<size_t>((size_t)n * d * 8, 16);
    double* X = nullptr; int64_t* doff = nullptr; int32_t* didx = nullptr; double* dval = nullptr; int32_t* xl = nullptr;
    cudaError_t e = cudaMalloc(&X, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&xl, std::max<size_t>((size_t)n * 4, 16));
    if (e == cudaSuccess) e = cudaMalloc(&doff, (size_t)(n + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc(&didx, std::max<size_t>((size_t)nnz * 4, 16));
    if (e == cudaSuccess) e = cudaMalloc(&dval, std::max<size_t>((size_t)nnz * 8, 16));
    if (e == cudaSuccess) e = cudaMemsetAsync(X, 0, bytes, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(doff, off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(didx, idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(dval, val, (size_t)nnz * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && n > 0) {
        k_densify<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(doff, didx, dval, n, d, X, xl);
        ++h->launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (doff) cudaFree(doff);
    if (didx) cudaFree(didx);
    if (dval) cudaFree(dval);
    if (e != cudaSuccess) {
        if (X) cudaFree(X);
        if (xl) cudaFree(xl);
        return rpf_fail(h, e == cudaErrorMemoryAllocation ? RPF_ERR_NOMEM : RPF_ERR_CUDA, std::string("set_points_sparse: ") + cudaGetErrorString(e));
    }
    h->dX = X; h->ownX = true; h->x_bytes = bytes; h->n = n; h->d = d; h->d_xlast = xl; h->x_pad_rows = 0;
    return RPF_OK;
}

int rpf_points_are_sparse(const rpf_handle* h) { return h ? (h->d_xlast ? 1 : 0) : -1; }

int rpf_densify_rows(int64_t nq, int32_t d, const int64_t* off, const int32_t* idx, const double* val, double* Q, int32_t* q_last) {
    if (nq < 0 || d < 1 || !off || !Q) return RPF_ERR_ARG;
    std::memset(Q, 0, (size_t)nq * d * 8);
    for (int64_t i = 0; i < nq; ++i) {
        if (off[i + 1] < off[i]) return RPF_ERR_ARG;
        for (int64_t q = off[i]; q < off[i + 1]; ++q) {
            if (idx[q] < 0 || idx[q] >= d || (q > off[i] && idx[q] <= idx[q - 1])) return RPF_ERR_ARG;
            Q[i * d + idx[q]] = val[q];
        }
        if (q_last) q_last[i] = off[i + 1] > off[i] ? idx[off[i + 1] - 1] : -1;
    }
    return RPF_OK;
}

int rpf_set_hyperplanes(rpf_handle* h, int32_t T, int32_t maxDepth, const int64_t* off, const int32_t* idx, const double* val) {
    if (!h) return RPF_ERR_ARG;
    if (T < 1 || maxDepth < 0 || !off) return rpf_fail(h, RPF_ERR_ARG, "set_hyperplanes: bad T/maxDepth/off");
    if (!h->group) RPF_SETDEV(h);
    const int64_t nrow = (int64_t)T * maxDepth;
    const int64_t nnz = off[nrow];
    if (off[0] != 0 || nnz < 0 || (nnz > 0 && (!idx || !val))) return rpf_fail(h, RPF_ERR_ARG, "set_hyperplanes: bad CSR");
    for (int64_t r = 0; r < nrow; ++r) if (off[r + 1] < off[r]) return rpf_fail(h, RPF_ERR_ARG, "set_hyperplanes: offsets not monotone");
    if (h->d > 0) for (int64_t q = 0; q < nnz; ++q) if (idx[q] < 0 || idx[q] >= h->d) return rpf_fail(h, RPF_ERR_ARG, "set_hyperplanes: component index out of range");
    if (h->group) return rpfg_set_hyperplanes(h, T, maxDepth, off, idx, val);
    rpf_insert_drop(h);
    h->T = T; h->hpDepth = maxDepth;
    h->hp_off.assign(off, off + nrow + 1);
    h->hp_idx.assign(idx, idx + nnz);
    h->hp_val.assign(val, val + nnz);
    return upload_hyperplanes(h);
}

int rpf_gen_hyperplanes(rpf_handle* h, uint64_t seed, int32_t T_total, int32_t maxDepth, double pnz, int32_t d,
                        int32_t t_first, int32_t T_local) {
    if (!h) return RPF_ERR_ARG;
    if (T_total < 1 || maxDepth < 0 || d < 1 || t_first < 0 || T_local < 1 || t_first + T_local > T_total)
        return rpf_fail(h, RPF_ERR_ARG, "gen_hyperplanes: bad arguments");
    if (!h->group) RPF_SETDEV(h);
    // One sequential generator for the whole forest: tree-major, level-major, component-minor; per component
    // one uniform (bernoulli p = u < p) and, on a hit, one more for the normal (Gen.hs:183-195).
    SMGen g = mk_smgen(seed);
    h->hp_off.clear(); h->hp_idx.clear(); h->hp_val.clear();
    for (int t = 0; t < T_total; ++t) {
        const bool keep = t >= t_first && t < t_first + T_local;
        for (int l = 0; l < maxDepth; ++l) {
            if (keep) h->hp_off.push_back((int64_t)h->hp_idx.size());
            for (int i = 0; i < d; ++i) {
                const double u = next_double(g);
                if (u < pnz) {
                    const double x = inv_norm_cdf(next_double(g)) * 1.0 + 0.0;
                    if (keep) { h->hp_idx.push_back(i); h->hp_val.push_back(x); }
                }
            }
        }
    }
    h->hp_off.push_back((int64_t)h->hp_idx.size());
    if (!h->group) rpf_insert_drop(h);
    h->T = T_local; h->hpDepth = maxDepth;
    if (h->group) return rpfg_after_gen_hyperplanes(h);
    return upload_hyperplanes(h);
}

// host-only variants (no handle, no GPU): used by the shim for sizing and by the CPU test-suite
int64_t rpf_sample_hyperplanes(uint64_t seed, int32_t T, int32_t maxDepth, double pnz, int32_t d,
                               int64_t* off, int32_t* idx, double* val) {
    if (T < 0 || maxDepth < 0 || d < 0) return RPF_ERR_ARG;
    SMGen g = mk_smgen(seed);
    int64_t nnz = 0;
    for (int t = 0; t < T; ++t)
        for (int l = 0; l < maxDepth; ++l) {
            if (off) off[(int64_t)t * maxDepth + l] = nnz;
            for (int i = 0; i < d; ++i) {
                const double u = next_double(g);
                if (u < pnz) {
                    const double x = inv_norm_cdf(next_double(g)) * 1.0 + 0.0;
                    if (idx) idx[nnz] = i;
                    if (val) val[nnz] = x;
                    ++nnz;
                }
            }
        }
    if (off) off[(int64_t)T * maxDepth] = nnz;
    return nnz;
}

int64_t rpf_topology_plan(int64_t n, int32_t maxDepth, int32_t minLeaf, int64_t* child, int32_t* depth,
                          int64_t* seg_start, int64_t* seg_size) {
    if (n < 0 || maxDepth < 0 || minLeaf < 0 || n >= ((int64_t)1 << 31)) return RPF_ERR_ARG;
    Topology tp;
    build_topology(tp, n, maxDepth, minLeaf);
    for (int64_t g = 0; g < tp.nnodes(); ++g) {
        if (child) child[g] = tp.child[g];
        if (depth) depth[g] = tp.depth[g];
        if (seg_start) seg_start[g] = tp.start[g];
        if (seg_size) seg_size[g] = tp.size[g];
    }
    return tp.nnodes();
}

/* rpTreeCfg, src/Data/RPTree/Conduit.hs:132-141 */
void rpf_rptree_cfg(int64_t minLeaf, int64_t n, int64_t d, int64_t* maxDepth, int64_t* chunk, double* pnz) {
    if (maxDepth) *maxDepth = (int64_t)std::ceil(std::log((double)n / (double)minLeaf) / std::log(2.0));
    if (chunk) *chunk = (int64_t)std::ceil((double)n / 100.0);
    if (pnz) { const double p = 1.0 / (std::log((double)d) / std::log(10.0)); *pnz = p < 1.0 ? p : 1.0; }
}

int rpf_leaf_order_exact(const rpf_handle* h) { return h ? (h->leaf_order_exact ? 1 : 0) : -1; }
int64_t rpf_points_lost(const rpf_handle* h) { return h ? h->stream_lost : -1; }

int64_t rpf_hyperplane_nnz(const rpf_handle* h) { return h ? (int64_t)h->hp_idx.size() : -1; }

int rpf_get_hyperplanes(const rpf_handle* h, int64_t* off, int32_t* idx, double* val) {
    if (!h) return RPF_ERR_ARG;
    if (off) std::memcpy(off, h->hp_off.data(), h->hp_off.size() * 8);
    if (idx && !h->hp_idx.empty()) std::memcpy(idx, h->hp_idx.data(), h->hp_idx.size() * 4);
    if (val && !h->hp_val.empty()) std::memcpy(val, h->hp_val.data(), h->hp_val.size() * 8);
    return RPF_OK;
}

static int check_build_args(rpf_handle* h, int32_t maxDepth, int32_t minLeaf) {
    if (!h->dX && h->n != 0) return rpf_fail(h, RPF_ERR_STATE, "build: call rpf_set_points first");
    if (h->d < 1) return rpf_fail(h, RPF_ERR_STATE, "build: call rpf_set_points first");
    if (h->T < 1) return rpf_fail(h, RPF_ERR_STATE, "build: call rpf_set_hyperplanes / rpf_gen_hyperplanes first");
    if (maxDepth < 0 || minLeaf < 0) return rpf_fail(h, RPF_ERR_ARG, "build: maxDepth and minLeaf must be >= 0");
    if (maxDepth > h->hpDepth) return rpf_fail(h, RPF_ERR_ARG, "build: maxDepth exceeds the number of hyperplanes per tree");
    if (maxDepth > 62) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "build: maxDepth > 62");
    for (int32_t q : h->hp_idx) if (q < 0 || q >= h->d) return rpf_fail(h, RPF_ERR_ARG, "build: hyperplane component index out of range");
    return RPF_OK;
}

int rpf_build(rpf_handle* h, int32_t maxDepth, int32_t minLeaf) {
    if (!h) return RPF_ERR_ARG;
    if (h->group) return rpfg_build(h, maxDepth, minLeaf, 0);
    if (h->insert_session) return rpf_fail(h, RPF_ERR_STATE, "build: an insert session is open (rpf_insert_end first)");
    int rc = check_build_args(h, maxDepth, minLeaf);
    if (rc) return rc;
    RPF_SETDEV(h);
    if (!(h->topo_key_n == h->n && h->topo_key_maxd == maxDepth && h->topo_key_minl == minLeaf)) {
        build_topology(h->topo, h->n, maxDepth, minLeaf);
        h->topo_key_n = h->n; h->topo_key_maxd = maxDepth; h->topo_key_minl = minLeaf;
    }
    rc = rpf_upload_topology(h);
    if (rc) return rc;
    h->built = false; h->sink_pending = false;
    h->stream_lost = 0;
    h->call_begin();
    rc = rpf_build_impl(h, nullptr);
    int rc2 = h->call_end();
    if (rc) return rc;
    if (rc2) return rc2;
    h->built = true;
    return RPF_OK;
}

// forestBatch straight from host memory: rpf_set_points + rpf_build with the upload overlapped with the projection
int rpf_build_from_host(rpf_handle* h, const double* X, int64_t n, int32_t d, int32_t maxDepth, int32_t minLeaf) {
    if (!h) return RPF_ERR_ARG;
    if (n < 0 || d < 1 || (n > 0 && !X)) return rpf_fail(h, RPF_ERR_ARG, "build_from_host: bad n/d/X");
    if (n >= (int64_t)1 << 31) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "build_from_host: n must be < 2^31");
    if (h->group) return rpfg_build_from_host(h, X, n, d, maxDepth, minLeaf);
    if (h->insert_session) return rpf_fail(h, RPF_ERR_STATE, "build: an insert session is open (rpf_insert_end first)");
    // validate against the NEW shape before the handle's state is touched
    if (h->T < 1) return rpf_fail(h, RPF_ERR_STATE, "build: call rpf_set_hyperplanes / rpf_gen_hyperplanes first");
    if (maxDepth < 0 || minLeaf < 0) return rpf_fail(h, RPF_ERR_ARG, "build: maxDepth and minLeaf must be >= 0");
    if (maxDepth > h->hpDepth) return rpf_fail(h, RPF_ERR_ARG, "build: maxDepth exceeds the number of hyperplanes per tree");
    if (maxDepth > 62) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "build: maxDepth > 62");
    for (int32_t q : h->hp_idx) if (q < 0 || q >= d) return rpf_fail(h, RPF_ERR_ARG, "build: hyperplane component index out of range");
    RPF_SETDEV(h);
    ++h->cfg_epoch; ++h->x_version;
    if (h->d_xlast) { cudaFree(h->d_xlast); h->d_xlast = nullptr; }
    const int64_t pad = rpf_comm_world(h) > 1 ? rpf_comm_world(h) : 0;      // spill rows of the in-place all-gather
    const size_t bytes = std::max<size_t>((size_t)(n + pad) * d * 8, 16);
    h->built = false; h->sink_pending = false;
    if (!(h->ownX && h->dX && h->x_bytes == bytes)) {
        if (h->ownX && h->dX) cudaFree((void*)h->dX);
        h->dX = nullptr; h->ownX = false; h->n = 0; h->x_bytes = 0;     // "no points" until the new buffer exists
        free_forest_dev(h);
        double* p = nullptr;
        RPF_CUDA(h, cudaMalloc(&p, bytes));
        h->dX = p; h->ownX = true; h->x_bytes = bytes;
    }
    h->n = n; h->d = d; h->x_pad_rows = pad;
    int rc = check_build_args(h, maxDepth, minLeaf);
    if (rc) return rc;
    if (!(h->topo_key_n == h->n && h->topo_key_maxd == maxDepth && h->topo_key_minl == minLeaf)) {
        build_topology(h->topo, h->n, maxDepth, minLeaf);
        h->topo_key_n = h->n; h->topo_key_maxd = maxDepth; h->topo_key_minl = minLeaf;
    }
    rc = rpf_upload_topology(h);
    if (rc) return rc;
    h->stream_lost = 0;
    h->call_begin();
    rc = rpf_build_impl(h, X);
    int rc2 = h->call_end();
    if (rc || rc2) {     // the buffer may hold a partial upload: a later rpf_build must not run on it
        if (h->ownX && h->dX) cudaFree((void*)h->dX);
        h->dX = nullptr; h->ownX = false; h->n = 0; h->x_bytes = 0;
        return rc ? rc : rc2;
    }
    h->built = true;
    return RPF_OK;
}

int rpf_build_chunked(rpf_handle* h, int32_t maxDepth, int32_t minLeaf, int64_t chunk) {
    if (!h) return RPF_ERR_ARG;
    if (chunk < 1) return rpf_fail(h, RPF_ERR_ARG, "build_chunked: chunk must be >= 1");
    if (h->group) return rpfg_build(h, maxDepth, minLeaf, chunk);
    if (h->insert_session) return rpf_fail(h, RPF_ERR_STATE, "build: an insert session is open (rpf_insert_end first)");
    if (chunk >= h->n) return rpf_build(h, maxDepth, minLeaf);   // one chunk == insert into an empty Tip == forestBatch
    int rc = check_build_args(h, maxDepth, minLeaf);
    if (rc) return rc;
    RPF_SETDEV(h);
    h->built = false; h->sink_pending = false;
    h->call_begin();
    rc = rpf_build_stream_impl(h, maxDepth, minLeaf, chunk);     // sets h->topo and the device topology itself
    int rc2 = h->call_end();
    if (rc) return rc;
    if (rc2) return rc2;
    h->built = true;
    return RPF_OK;
}

// ---------------------------------------------------------------------------------------------------
// incremental insert (stream.cu): forest = foldl insertMulti over the chunks as they arrive (Conduit.hs:157-176)
// ---------------------------------------------------------------------------------------------------
int rpf_insert_begin(rpf_handle* h, int32_t d, int32_t maxDepth, int32_t minLeaf) {
    if (!h) return RPF_ERR_ARG;
    if (h->group) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "insert: not available on a multi-GPU handle (use one handle per tree shard)");
    if (d < 1) return rpf_fail(h, RPF_ERR_ARG, "insert_begin: d must be >= 1");
    if (h->T < 1) return rpf_fail(h, RPF_ERR_STATE, "insert_begin: call rpf_set_hyperplanes / rpf_gen_hyperplanes first");
    if (maxDepth < 0 || minLeaf < 0) return rpf_fail(h, RPF_ERR_ARG, "insert_begin: maxDepth and minLeaf must be >= 0");
    if (maxDepth > h->hpDepth) return rpf_fail(h, RPF_ERR_ARG, "insert_begin: maxDepth exceeds the number of hyperplanes per tree");
    if (maxDepth > 62) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "insert_begin: maxDepth > 62");
    for (int32_t q : h->hp_idx) if (q < 0 || q >= d) return rpf_fail(h, RPF_ERR_ARG, "insert_begin: hyperplane component index out of range");
    if (rpf_comm_world(h) > 1) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "insert: every rank inserts the full chunk itself; detach the communicator first");
    RPF_SETDEV(h);
    free_forest_dev(h);
    return rpf_insert_begin_impl(h, d, maxDepth, minLeaf);
}

int rpf_insert_chunk(rpf_handle* h, const double* X_chunk, int64_t m) {
    if (!h) return RPF_ERR_ARG;
    if (!h->insert_session) return rpf_fail(h, RPF_ERR_STATE, "insert_chunk: call rpf_insert_begin first");
    if (m < 0 || (m > 0 && !X_chunk)) return rpf_fail(h, RPF_ERR_ARG, "insert_chunk: bad m / X_chunk");
    if (m == 0) return RPF_OK;
    RPF_SETDEV(h);
    h->sink_pending = false;
    h->call_begin();
    int rc = rpf_insert_chunk_impl(h, X_chunk, m);
    int rc2 = h->call_end();
    if (rc || rc2) {
        rpf_insert_drop(h);                        // the device state no longer matches the planner's
        h->built = false;
        return rc ? rc : rc2;
    }
    h->built = true;
    return RPF_OK;
}

int rpf_insert_end(rpf_handle* h) {
    if (!h) return RPF_ERR_ARG;
    if (!h->insert_session) return RPF_OK;
    RPF_SETDEV(h);
    return rpf_insert_end_impl(h);
}

// ---------------------------------------------------------------------------------------------------
// checkpoint: the device image of a built forest (serialiseRPForest / deserialiseRPForest, Internal.hs:185-196, store the
// Haskell value as CBOR; a Haskell host still has that by rebuilding the value from rpf_forest_export.  This is the
// engine-side equivalent: one flat little-endian file, restored without rebuilding.)
// ---------------------------------------------------------------------------------------------------

int rpf_forest_save(rpf_handle* h, const char* path, int32_t with_points) {
    if (!h || !path) return RPF_ERR_ARG;
    if (!h->built) return rpf_fail(h, RPF_ERR_STATE, "forest_save: forest not built");
    if (h->group) return rpfg_forest_save(h, path, with_points);
    RPF_SETDEV(h);
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    FILE* f = fopen(path, "wb");
    if (!f) return rpf_fail(h, RPF_ERR_ARG, std::string("forest_save: cannot open ") + path);
    const Topology& tp = h->topo;
    CkptHeader H{};
    std::memcpy(H.magic, "RPFB200", 8);
    H.version = 1; H.flags = (with_points ? 1u : 0u) | (h->d_xlast ? 2u : 0u);
    H.n = h->n; H.nn = tp.nnodes(); H.hp_nnz = (int64_t)h->hp_idx.size(); H.lost = h->stream_lost;
    H.d = h->d; H.T = h->T; H.hpDepth = h->hpDepth; H.maxDepth = tp.maxDepth; H.minLeaf = tp.minLeaf; H.nlevels = tp.nlevels; H.L_eff = tp.L_eff;
    H.leaf_order_exact = h->leaf_order_exact ? 1 : 0;
    std::vector<char> tmp;
    const size_t nn = (size_t)H.nn, T = (size_t)H.T, n = (size_t)H.n;
    bool ok = wr(f, &H, 1) && wr(f, h->hp_off.data(), h->hp_off.size()) && wr(f, h->hp_idx.data(), h->hp_idx.size()) &&
              wr(f, h->hp_val.data(), h->hp_val.size()) && wr(f, tp.start.data(), nn) && wr(f, tp.size.data(), nn) &&
              wr(f, tp.child.data(), nn) && wr(f, tp.depth.data(), nn) && wr(f, tp.level_off.data(), (size_t)tp.nlevels + 1) &&
              wr(f, tp.lvl_maxsize.data(), (size_t)tp.nlevels) &&
              wr_dev(f, h->d_thr, T * nn, tmp) && wr_dev(f, h->d_mlo, T * nn, tmp) && wr_dev(f, h->d_mhi, T * nn, tmp) &&
              wr_dev(f, h->d_perm, T * n, tmp);
    if (ok && with_points) {
        ok = wr_dev(f, h->dX, n * (size_t)h->d, tmp);
        if (ok && h->d_xlast) ok = wr_dev(f, h->d_xlast, n, tmp);
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) return rpf_fail(h, RPF_ERR_CUDA, std::string("forest_save: write failed: ") + path);
    return RPF_OK;
}

// every id in perm must be a row number (< n) or the "dropped slot" marker of a streaming build
__global__ void k_check_perm(const uint32_t* __restrict__ perm, size_t cnt, uint32_t n, int* __restrict__ bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int b = 0;
    for (; i < cnt; i += stride) { const uint32_t v = perm[i]; if (v >= n && v != 0xffffffffu) b = 1; }
    if (b) *bad = 1;
}

// Everything a query kernel later indexes with is validated here, BEFORE the handle is touched (deserialiseRPForest
// returns Left on malformed input, Internal.hs:191-196): table sizes against the file length, CSR offsets, component
// indices, the BFS topology, and -- on the device -- the row ids of perm.
static int forest_load_impl(rpf_handle* h, const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return rpf_fail(h, RPF_ERR_ARG, std::string("forest_load: cannot open ") + path);
    struct Closer { FILE* f; ~Closer() { if (f) fclose(f); } } closer{f};     // also on an exception (bad_alloc)
    CkptHeader H{};
    auto fail = [&](const char* why) { return rpf_fail(h, RPF_ERR_ARG, std::string("forest_load: ") + why); };
    if (!rd(f, &H, 1) || std::memcmp(H.magic, "RPFB200", 8) != 0 || H.version != 1) return fail("not a forest checkpoint");
    if (H.n < 0 || H.n >= ((int64_t)1 << 31) || H.nn < 1 || H.nn >= ((int64_t)1 << 31) || H.T < 1 || H.d < 1 || H.hpDepth < 0 ||
        H.hpDepth > 4096 || H.nlevels < 1 || H.nlevels > 64 || H.hp_nnz < 0 || H.lost < 0 || H.maxDepth < 0 || H.minLeaf < 0 || (H.flags & ~3u))
        return fail("corrupt header");
    const bool has_points = H.flags & 1u, sparse = H.flags & 2u;
    // the header fixes the file length exactly; nothing below allocates more than the file holds
    const long here = ftell(f);
    if (fseek(f, 0, SEEK_END) != 0) return fail("cannot seek");
    const long flen = ftell(f);
    if (here < 0 || flen < 0 || fseek(f, here, SEEK_SET) != 0) return fail("cannot seek");
    const unsigned __int128 nn = (unsigned __int128)H.nn, T = (unsigned __int128)H.T, n = (unsigned __int128)H.n, d = (unsigned __int128)H.d;
    unsigned __int128 want = (unsigned __int128)here + (T * (unsigned __int128)H.hpDepth + 1) * 8 + (unsigned __int128)H.hp_nnz * 12 +
                             nn * 16 + ((unsigned __int128)H.nlevels + 1) * 8 + (unsigned __int128)H.nlevels * 4 + T * nn * 24 + T * n * 4;
    if (has_points) want += n * d * 8 + (sparse ? n * 4 : 0);
    if (want != (unsigned __int128)flen) return fail("file length does not match the header (truncated or corrupt)");
    if (!has_points) {
        if (!h->dX || h->n != H.n || h->d != H.d) return fail("the checkpoint carries no points: call rpf_set_points with the same data first");
        if (sparse != (h->d_xlast != nullptr)) return fail("point representation (SVector / DVector) differs from the checkpoint's");
    }
    const size_t snn = (size_t)H.nn, sT = (size_t)H.T, sn = (size_t)H.n;
    std::vector<int64_t> hp_off((size_t)H.T * H.hpDepth + 1); std::vector<int32_t> hp_idx((size_t)H.hp_nnz); std::vector<double> hp_val((size_t)H.hp_nnz);
    Topology tp;
    tp.n = H.n; tp.maxDepth = H.maxDepth; tp.minLeaf = H.minLeaf; tp.nlevels = H.nlevels;
    tp.start.resize(snn); tp.size.resize(snn); tp.child.resize(snn); tp.depth.resize(snn); tp.level_off.resize((size_t)H.nlevels + 1); tp.lvl_maxsize.resize((size_t)H.nlevels);
    if (!(rd(f, hp_off.data(), hp_off.size()) && rd(f, hp_idx.data(), hp_idx.size()) && rd(f, hp_val.data(), hp_val.size()) &&
          rd(f, tp.start.data(), snn) && rd(f, tp.size.data(), snn) && rd(f, tp.child.data(), snn) && rd(f, tp.depth.data(), snn) &&
          rd(f, tp.level_off.data(), (size_t)H.nlevels + 1) && rd(f, tp.lvl_maxsize.data(), (size_t)H.nlevels)))
        return fail("truncated file");
    // hyperplanes: CSR rows over (tree, level)
    if (hp_off[0] != 0 || hp_off.back() != H.hp_nnz) return fail("hyperplane offsets do not match the header");
    for (size_t r = 0; r + 1 < hp_off.size(); ++r) if (hp_off[r + 1] < hp_off[r]) return fail("hyperplane offsets not monotone");
    for (int32_t q : hp_idx) if (q < 0 || q >= H.d) return fail("hyperplane component index out of range");
    // topology: BFS numbering, level table, segments inside perm
    if (tp.level_off[0] != 0 || tp.level_off.back() != H.nn) return fail("level table does not match the node count");
    for (int l = 0; l < H.nlevels; ++l) if (tp.level_off[l + 1] <= tp.level_off[l]) return fail("level table not increasing");
    int L_eff = 0;
    for (int l = 0; l < H.nlevels; ++l) {
        uint32_t mx = 0;
        for (int64_t g = tp.level_off[l]; g < tp.level_off[l + 1]; ++g) {
            if (tp.depth[g] != l) return fail("node depth does not match its level");
            if ((uint64_t)tp.start[g] + tp.size[g] > (uint64_t)H.n) return fail("node segment outside the point array");
            mx = std::max(mx, tp.size[g]);
            const int32_t c = tp.child[g];
            if (c == -1) continue;
            if (l + 1 >= H.nlevels || c < tp.level_off[l + 1] || (int64_t)c + 1 >= tp.level_off[l + 2]) return fail("child id outside the next level");
            if (l >= H.hpDepth) return fail("internal node deeper than the stored hyperplanes");
            if ((uint64_t)tp.size[c] + tp.size[c + 1] > (uint64_t)tp.size[g] || tp.start[c] < tp.start[g] ||
                (uint64_t)tp.start[c + 1] + tp.size[c + 1] > (uint64_t)tp.start[g] + tp.size[g])
                return fail("child segments outside the parent's");
            L_eff = l + 1;
        }
        tp.lvl_maxsize[l] = mx;          // derived, not trusted: sizes shared-memory buffers of the query kernels
    }
    tp.L_eff = L_eff;
    if (H.lost > H.n) return fail("corrupt header");

    // ---- the file is sane: now replace the handle's state
    h->built = false; h->sink_pending = false;
    ++h->cfg_epoch; ++h->x_version;
    h->T = H.T; h->hpDepth = H.hpDepth;
    h->hp_off.swap(hp_off); h->hp_idx.swap(hp_idx); h->hp_val.swap(hp_val);
    auto fail_reset = [&](int code, const char* why) {      // the previous forest is gone; leave a consistent "not built" handle
        h->built = false; h->sink_pending = false;
        return rpf_fail(h, code, std::string("forest_load: ") + why);
    };
    if (has_points) {      // the data set travels with the forest: replace whatever the handle holds
        if (h->d_xlast) { cudaFree(h->d_xlast); h->d_xlast = nullptr; }
        if (h->ownX && h->dX) cudaFree((void*)h->dX);
        h->dX = nullptr; h->ownX = false; h->n = 0; h->x_bytes = 0;
        double* X = nullptr;
        const size_t bytes = std::max<size_t>(sn * (size_t)H.d * 8, 16);
        if (cudaMalloc(&X, bytes) != cudaSuccess) { cudaGetLastError(); return fail_reset(RPF_ERR_NOMEM, "out of device memory"); }
        h->dX = X; h->ownX = true; h->x_bytes = bytes; h->n = H.n; h->d = H.d;
    }
    int rc = upload_hyperplanes(h);      // also drops the previous forest arrays
    if (rc) return rc;
    h->topo = tp; h->topo_key_n = -1;
    rc = rpf_upload_topology(h);
    if (!rc) rc = rpf_alloc_forest(h, H.nn, H.n);
    if (rc) return rc;
    std::vector<char> tmp;
    bool ok = rd_dev(f, h->d_thr, sT * snn, tmp) && rd_dev(f, h->d_mlo, sT * snn, tmp) && rd_dev(f, h->d_mhi, sT * snn, tmp) && rd_dev(f, h->d_perm, sT * sn, tmp);
    if (ok && has_points) {
        ok = rd_dev(f, (double*)h->dX, sn * (size_t)H.d, tmp);
        if (ok && sparse) {
            if (cudaMalloc(&h->d_xlast, std::max<size_t>(sn * 4, 16)) != cudaSuccess) { cudaGetLastError(); return fail_reset(RPF_ERR_NOMEM, "out of device memory"); }
            ok = rd_dev(f, h->d_xlast, sn, tmp);
        }
    }
    if (!ok) return fail_reset(RPF_ERR_ARG, "truncated file");
    if (sT * sn > 0) {
        int* bad = (int*)h->ws_get(WS_MAXCNT, 16);
        if (!bad) return RPF_ERR_NOMEM;
        RPF_CUDA(h, cudaMemsetAsync(bad, 0, 4, h->stream));
        k_check_perm<<<1184, 256, 0, h->stream>>>(h->d_perm, sT * sn, (uint32_t)H.n, bad);
        ++h->launches;
        int hb = 0;
        RPF_CUDA(h, cudaMemcpyAsync(&hb, bad, 4, cudaMemcpyDeviceToHost, h->stream));
        RPF_CUDA(h, cudaStreamSynchronize(h->stream));
        if (hb) return rpf_fail(h, RPF_ERR_ARG, "forest_load: perm holds a row id >= n");
    }
    h->stream_lost = H.lost; h->leaf_order_exact = H.leaf_order_exact != 0;
    h->built = true;
    return RPF_OK;
}

int rpf_forest_load(rpf_handle* h, const char* path) {
    if (!h || !path) return RPF_ERR_ARG;
    if (h->group) return rpfg_forest_load(h, path);
    RPF_SETDEV(h);
    rpf_insert_drop(h);
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    try {
        return forest_load_impl(h, path);
    } catch (const std::bad_alloc&) {
        h->built = false; h->sink_pending = false;
        return rpf_fail(h, RPF_ERR_NOMEM, "forest_load: out of host memory");
    } catch (const std::exception& e) {
        h->built = false; h->sink_pending = false;
        return rpf_fail(h, RPF_ERR_ARG, std::string("forest_load: ") + e.what());
    }
}

int64_t rpf_num_nodes(const rpf_handle* h) { return h ? h->topo.nnodes() : -1; }
int32_t rpf_num_trees(const rpf_handle* h) { return h ? h->T : -1; }
int32_t rpf_hyperplane_depth(const rpf_handle* h) { return h ? h->hpDepth : -1; }
int rpf_points_shape(const rpf_handle* h, int64_t* n, int32_t* d) { if (!h) return RPF_ERR_ARG; if (n) *n = h->n; if (d) *d = h->d; return RPF_OK; }

int rpf_topology(const rpf_handle* h, int64_t* child, int32_t* depth, int64_t* seg_start, int64_t* seg_size) {
    if (!h) return RPF_ERR_ARG;
    const Topology& tp = h->topo;
    for (int64_t g = 0; g < tp.nnodes(); ++g) {
        if (child) child[g] = tp.child[g];
        if (depth) depth[g] = tp.depth[g];
        if (seg_start) seg_start[g] = tp.start[g];
        if (seg_size) seg_size[g] = tp.size[g];
    }
    return RPF_OK;
}

int rpf_tree_export(rpf_handle* h, int32_t t, double* thr, double* mlo, double* mhi, uint32_t* perm) {
    if (!h) return RPF_ERR_ARG;
    if (!h->built) return rpf_fail(h, RPF_ERR_STATE, "tree_export: forest not built");
    if (h->group) return rpfg_tree_export(h, t, thr, mlo, mhi, perm);
    if (t < 0 || t >= h->T) return rpf_fail(h, RPF_ERR_ARG, "tree_export: tree index out of range");
    RPF_SETDEV(h);
    const size_t nn = (size_t)h->topo.nnodes();
    if (thr) RPF_CUDA(h, cudaMemcpy(thr, h->d_thr + (size_t)t * nn, nn * 8, cudaMemcpyDeviceToHost));
    if (mlo) RPF_CUDA(h, cudaMemcpy(mlo, h->d_mlo + (size_t)t * nn, nn * 8, cudaMemcpyDeviceToHost));
    if (mhi) RPF_CUDA(h, cudaMemcpy(mhi, h->d_mhi + (size_t)t * nn, nn * 8, cudaMemcpyDeviceToHost));
    if (perm && h->n > 0) RPF_CUDA(h, cudaMemcpy(perm, h->d_perm + (size_t)t * h->n, (size_t)h->n * 4, cudaMemcpyDeviceToHost));
    return RPF_OK;
}

int rpf_forest_export(rpf_handle* h, double* thr, double* mlo, double* mhi, uint32_t* perm) {
    if (!h) return RPF_ERR_ARG;
    if (!h->built) return rpf_fail(h, RPF_ERR_STATE, "forest_export: forest not built");
    if (h->group) return rpfg_forest_export(h, thr, mlo, mhi, perm);
    RPF_SETDEV(h);
    if (h->sink_pending && perm == h->sink_perm && thr == h->sink_thr && mlo == h->sink_mlo && mhi == h->sink_mhi) {
        RPF_CUDA(h, cudaEventSynchronize(h->sink_ev[9]));        // the build streamed the forest into these buffers already
        return RPF_OK;
    }
    // the device arrays are already [T][nodes] / [T][n]: four straight copies on the engine's stream, one sync
    const size_t nb = (size_t)h->T * (size_t)h->topo.nnodes() * 8, pb = (size_t)h->T * (size_t)h->n * 4;
    if (thr && nb) RPF_CUDA(h, cudaMemcpyAsync(thr, h->d_thr, nb, cudaMemcpyDeviceToHost, h->stream));
    if (mlo && nb) RPF_CUDA(h, cudaMemcpyAsync(mlo, h->d_mlo, nb, cudaMemcpyDeviceToHost, h->stream));
    if (mhi && nb) RPF_CUDA(h, cudaMemcpyAsync(mhi, h->d_mhi, nb, cudaMemcpyDeviceToHost, h->stream));
    if (perm && pb) RPF_CUDA(h, cudaMemcpyAsync(perm, h->d_perm, pb, cudaMemcpyDeviceToHost, h->stream));
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RPF_OK;
}

int rpf_set_export_sink(rpf_handle* h, double* thr, double* mlo, double* mhi, uint32_t* perm) {
    if (!h) return RPF_ERR_ARG;
    if (h->group) return rpfg_set_export_sink(h, thr, mlo, mhi, perm);
    RPF_SETDEV(h);
    if (h->d2h_stream) RPF_CUDA(h, cudaStreamSynchronize(h->d2h_stream));
    h->sink_thr = thr; h->sink_mlo = mlo; h->sink_mhi = mhi; h->sink_perm = perm;
    h->sink_pending = false;
    return RPF_OK;
}

int rpf_candidates_count(rpf_handle* h, const double* Q, int64_t nq, int32_t t, int64_t* off_out) {
    if (!h) return RPF_ERR_ARG;
    if (!h->built) return rpf_fail(h, RPF_ERR_STATE, "candidates: forest not built");
    if (nq < 0 || (nq > 0 && !Q) || !off_out || t < -1 || t >= h->T) return rpf_fail(h, RPF_ERR_ARG, "candidates_count: bad arguments");
    if (h->group) return rpfg_candidates(h, Q, nq, t, off_out, nullptr, nullptr);
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_candidates_impl(h, Q, nq, t, off_out, nullptr, nullptr);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

int rpf_candidates(rpf_handle* h, const double* Q, int64_t nq, int32_t t, const int64_t* off, uint32_t* ids) {
    if (!h) return RPF_ERR_ARG;
    if (!h->built) return rpf_fail(h, RPF_ERR_STATE, "candidates: forest not built");
    if (nq < 0 || (nq > 0 && !Q) || !off || t < -1 || t >= h->T) return rpf_fail(h, RPF_ERR_ARG, "candidates: bad arguments");
    if (h->group) return rpfg_candidates(h, Q, nq, t, nullptr, off, ids);
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_candidates_impl(h, Q, nq, t, nullptr, off, ids);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

int rpf_knn(rpf_handle* h, const double* Q, int64_t nq, int32_t k, int32_t dedup, double* dist, uint32_t* ids, int32_t* count) {
    if (!h) return RPF_ERR_ARG;
    if (!h->built) return rpf_fail(h, RPF_ERR_STATE, "knn: forest not built");
    if (nq < 0 || (nq > 0 && (!Q || !dist || !ids)) || k < 1 || k > 1024) return rpf_fail(h, RPF_ERR_ARG, "knn: bad arguments (1 <= k <= 1024)");
    if (h->group) return rpfg_knn(h, Q, nullptr, nq, k, dedup, dist, ids, count);
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_knn_impl(h, Q, nullptr, nq, k, dedup, dist, ids, count, false);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

int rpf_knn_s(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, int32_t dedup, double* dist, uint32_t* ids, int32_t* count) {
    if (!h) return RPF_ERR_ARG;
    if (!h->built) return rpf_fail(h, RPF_ERR_STATE, "knn: forest not built");
    if (nq < 0 || (nq > 0 && (!Q || !dist || !ids)) || k < 1 || k > 1024) return rpf_fail(h, RPF_ERR_ARG, "knn: bad arguments (1 <= k <= 1024)");
    if (h->group) return rpfg_knn(h, Q, q_last, nq, k, dedup, dist, ids, count);
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_knn_impl(h, Q, q_last, nq, k, dedup, dist, ids, count, false);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

int64_t rpf_knn_h_capacity(const rpf_handle* h, int32_t k) {
    if (!h || !h->built || k < 1) return -1;
    uint32_t mx = 0;
    const Topology& tp = h->topo;
    for (int64_t g = 0; g < tp.nnodes(); ++g) if (tp.child[g] < 0) mx = std::max(mx, tp.size[g]);
    return std::max<int64_t>(k, mx);
}

int rpf_knn_h(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, int64_t cap, double* dist, uint32_t* ids, int32_t* count) {
    if (!h) return RPF_ERR_ARG;
    if (!h->built) return rpf_fail(h, RPF_ERR_STATE, "knnH: forest not built");
    if (nq < 0 || (nq > 0 && (!Q || !dist || !ids || !count)) || k < 1) return rpf_fail(h, RPF_ERR_ARG, "knnH: bad arguments");
    if (cap < rpf_knn_h_capacity(h, k) || cap > 0x7fffffff) return rpf_fail(h, RPF_ERR_ARG, "knnH: cap must be >= rpf_knn_h_capacity(h, k)");
    if (h->group || rpf_comm_world(h) > 1)
        return rpf_fail(h, RPF_ERR_UNSUPPORTED, "knnH: the margin-priority search ranks the leaves of ALL trees in one heap; not sharded across GPUs yet");
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_knn_h_impl(h, Q, q_last, nq, k, (int)cap, dist, ids, count);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

int rpf_recall(rpf_handle* h, const double* Q, int64_t nq, int32_t k, double* recall_sum) {
    if (!h) return RPF_ERR_ARG;
    if (!h->built) return rpf_fail(h, RPF_ERR_STATE, "recall: forest not built");
    if (nq < 0 || (nq > 0 && (!Q || !recall_sum)) || k < 1 || k > 1024) return rpf_fail(h, RPF_ERR_ARG, "recall: bad arguments (1 <= k <= 1024)");
    if (h->group) return rpfg_recall(h, Q, nullptr, nq, k, recall_sum);
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_recall_impl(h, Q, nullptr, nq, k, recall_sum);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

int rpf_recall_s(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, double* recall_sum) {
    if (!h) return RPF_ERR_ARG;
    if (!h->built) return rpf_fail(h, RPF_ERR_STATE, "recall: forest not built");
    if (nq < 0 || (nq > 0 && (!Q || !recall_sum)) || k < 1 || k > 1024) return rpf_fail(h, RPF_ERR_ARG, "recall: bad arguments (1 <= k <= 1024)");
    if (h->group) return rpfg_recall(h, Q, q_last, nq, k, recall_sum);
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_recall_impl(h, Q, q_last, nq, k, recall_sum);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

int rpf_brute_knn(rpf_handle* h, const double* Q, int64_t nq, int32_t k, double* dist, uint32_t* ids) {
    if (!h) return RPF_ERR_ARG;
    if (h->group) return rpfg_brute_knn(h, Q, nullptr, nq, k, dist, ids);
    if (!h->dX) return rpf_fail(h, RPF_ERR_STATE, "brute_knn: call rpf_set_points first");
    if (nq < 0 || (nq > 0 && (!Q || !dist || !ids)) || k < 1 || k > 1024) return rpf_fail(h, RPF_ERR_ARG, "brute_knn: bad arguments (1 <= k <= 1024)");
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_brute_knn_impl(h, Q, nullptr, nq, k, dist, ids);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

int rpf_brute_knn_s(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, double* dist, uint32_t* ids) {
    if (!h) return RPF_ERR_ARG;
    if (h->group) return rpfg_brute_knn(h, Q, q_last, nq, k, dist, ids);
    if (!h->dX) return rpf_fail(h, RPF_ERR_STATE, "brute_knn: call rpf_set_points first");
    if (nq < 0 || (nq > 0 && (!Q || !dist || !ids)) || k < 1 || k > 1024) return rpf_fail(h, RPF_ERR_ARG, "brute_knn: bad arguments (1 <= k <= 1024)");
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_brute_knn_impl(h, Q, q_last, nq, k, dist, ids);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

int rpf_merge_topk(rpf_handle* h, int32_t G, int64_t nq, int32_t k, int32_t dedup, const double* dist, const uint32_t* ids,
                   const int32_t* count, double* dist_out, uint32_t* ids_out, int32_t* count_out) {
    if (!h) return RPF_ERR_ARG;
    if (G < 1 || nq < 0 || k < 1 || k > 1024 || (nq > 0 && (!dist || !ids || !count || !dist_out || !ids_out)))
        return rpf_fail(h, RPF_ERR_ARG, "merge_topk: bad arguments");
    if (h->group) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "merge_topk: a multi-GPU handle merges inside rpf_knn");
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_merge_impl(h, G, nq, k, dedup, dist, ids, count, dist_out, ids_out, count_out, false);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

int rpf_knn_dev(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int32_t k, int32_t dedup,
                double* dist_dev, uint32_t* ids_dev, int32_t* count_dev) {
    if (!h) return RPF_ERR_ARG;
    if (!h->built) return rpf_fail(h, RPF_ERR_STATE, "knn: forest not built");
    if (nq < 0 || (nq > 0 && (!Q || !dist_dev || !ids_dev || !count_dev)) || k < 1 || k > 1024) return rpf_fail(h, RPF_ERR_ARG, "knn_dev: bad arguments (1 <= k <= 1024)");
    if (h->group) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "knn_dev: a multi-GPU handle exchanges and merges inside rpf_knn");
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_knn_impl(h, Q, q_last, nq, k, dedup, dist_dev, ids_dev, count_dev, true);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

int rpf_merge_topk_dev(rpf_handle* h, int32_t G, int64_t nq, int32_t k, int32_t dedup, const double* dist_dev, const uint32_t* ids_dev,
                       const int32_t* count_dev, double* dist_out, uint32_t* ids_out, int32_t* count_out) {
    if (!h) return RPF_ERR_ARG;
    if (G < 1 || nq < 0 || k < 1 || k > 1024 || (nq > 0 && (!dist_dev || !ids_dev || !count_dev || !dist_out || !ids_out)))
        return rpf_fail(h, RPF_ERR_ARG, "merge_topk_dev: bad arguments");
    if (h->group) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "merge_topk_dev: a multi-GPU handle merges inside rpf_knn");
    RPF_SETDEV(h);
    h->call_begin();
    int rc = rpf_merge_impl(h, G, nq, k, dedup, dist_dev, ids_dev, count_dev, dist_out, ids_out, count_out, true);
    int rc2 = h->call_end();
    return rc ? rc : rc2;
}

double rpf_last_device_ms(const rpf_handle* h) { return h ? h->last_ms : -1.0; }
int rpf_set_profiling(rpf_handle* h, int on) { if (!h) return RPF_ERR_ARG; if (h->group) return rpfg_set_profiling(h, on); h->profiling = on != 0; ++h->cfg_epoch; return RPF_OK; }
int rpf_get_profile(const rpf_handle* h, double* ms, int64_t* launches, int cap) {
    if (!h) return RPF_ERR_ARG;
    if (h->group) return rpfg_get_profile(h, ms, launches, cap);
    for (int i = 0; i < PH_COUNT && i < cap; ++i) {
        if (ms) ms[i] = h->phase_ms[i];
        if (launches) launches[i] = h->phase_launches[i];
    }
    return PH_COUNT;
}
const char* rpf_phase_name(int i) { return (i >= 0 && i < PH_COUNT) ? kPhaseNames[i] : ""; }
int64_t rpf_launch_count(const rpf_handle* h) { return h ? (h->group ? rpfg_launch_count(h) : h->launches) : -1; }
int rpf_set_option(rpf_handle* h, const char* name, int64_t value) {
    if (!h || !name) return RPF_ERR_ARG;
    if (h->group) return rpfg_set_option(h, name, value, false);
    const std::string s(name);
    ++h->cfg_epoch;
    if (s == "cuda_graph") { h->use_graphs = value != 0; return RPF_OK; }
    if (s == "top_chunk_hist") { h->top_chunk[0] = (int)value; return RPF_OK; }
    if (s == "top_chunk_compact") { h->top_chunk[1] = (int)value; return RPF_OK; }
    if (s == "top_chunk_relabel") { h->top_chunk[2] = (int)value; return RPF_OK; }
    if (s == "lean_top") { h->lean_top = value != 0; return RPF_OK; }
    if (s == "force_generic_bottom") { h->force_generic_bottom = value != 0; return RPF_OK; }
    if (s == "bottom_words64") { h->bottom_words64 = value != 0; return RPF_OK; }
    if (s == "force_simple_topk") { h->force_simple_topk = value != 0; return RPF_OK; }
    if (s == "knn_f32_stages") { h->knn_f32_cfg[0] = (int)value; return RPF_OK; }
    if (s == "knn_f32_rows") { h->knn_f32_cfg[1] = (int)value; return RPF_OK; }
    if (s == "knn_f32_buf") { h->knn_f32_cfg[2] = (int)value; return RPF_OK; }
    if (s == "knn_f32_sreg") { h->knn_f32_cfg[3] = (int)value; return RPF_OK; }
    if (s == "knn_filter32") { h->knn_filter32 = (int)value; return RPF_OK; }
    if (s == "force_simple_knn") { h->force_simple_knn = value != 0; return RPF_OK; }
    if (s == "no_query_order") { h->no_query_order = value != 0; return RPF_OK; }
    if (s == "project_variant") { h->project_variant = (int)value; return RPF_OK; }
    if (s == "bottom_select") { h->bottom_select = (int)value; return RPF_OK; }
    if (s == "hist_big_chunk") { h->hist_big_chunk = (int)value; return RPF_OK; }
    if (s == "fused_pick_min_tg") { h->fused_pick_min_tg = (int)value; return RPF_OK; }
    if (s == "fuse_relabel_hist") { h->fuse_relabel_hist = (int)value; return RPF_OK; }
    if (s == "fused_top") { h->fused_top = (int)value; return RPF_OK; }
    if (s == "rerank_gemm") { h->rerank_gemm = (int)value; return RPF_OK; }
    if (s == "project_prefetch") { h->project_prefetch = (int)value; return RPF_OK; }
    if (s == "project_pipe_maxh") { h->project_pipe_maxh = (int)value; return RPF_OK; }
    if (s == "branches") { h->branches = (int)value; h->tg_cached = 0; return RPF_OK; }
    if (s == "release_workspace") { cudaStreamSynchronize(h->stream); h->ws_free_all(); h->tg_cached = 0; return RPF_OK; }
    return rpf_fail(h, RPF_ERR_ARG, "unknown option " + s);
}
int rpf_set_bottom_cap(rpf_handle* h, int32_t cap) {
    if (!h) return RPF_ERR_ARG;
    if (cap != 256 && cap != 512 && cap != 1024 && cap != 2048 && cap != 4096 && cap != 8192)
        return rpf_fail(h, RPF_ERR_ARG, "bottom_cap must be a power of two in [256, 8192]");
    if (h->group) return rpfg_set_option(h, "", cap, true);
    h->bottom_cap = cap;
    h->tg_cached = 0;
    ++h->cfg_epoch;
    return RPF_OK;
}

}  // extern "C"
