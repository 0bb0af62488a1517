// rpf_internal.h -- engine state shared by the C ABI (capi.cu), the build path (build.cu) and the
// query path (query.cu).  Not part of the public interface (that is include/rpforest.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <string.h>
#include <vector>
#include "../../include/rpforest.h"

// ---- phases (profiling) -----------------------------------------------------------------------------
enum rpf_phase {
    PH_PROJECT = 0,     // K1: sparse-hyperplane . dense-row projections of all points (innerSD)
    PH_TOP_HIST,        // top phase: per-node histogram of keys
    PH_TOP_PICK,        // top phase: locate the median bin
    PH_TOP_COMPACT,     // top phase: gather the median bin's keys
    PH_TOP_FINISH,      // top phase: exact median inside the bin
    PH_TOP_TIES,        // top phase: lexicographic tie resolution (rare)
    PH_TOP_RELABEL,     // top phase: relabel / scatter points to children
    PH_BOTTOM,          // bottom phase: whole subtrees in shared memory
    PH_Q_PROJECT,       // query projections
    PH_Q_TRAVERSE,      // candidates descent
    PH_Q_KNN,           // leaf re-rank + top-k
    PH_Q_CAND,          // candidates materialisation
    PH_TRUTH,           // brute-force top-k
    PH_RECALL,          // candidate-set /\ truth
    PH_MERGE,           // multi-GPU top-k merge
    PH_STREAM,          // streaming build: threshold/margin updates, export
    PH_STREAM_CONCAT,   // streaming build: Tip concatenation (leaf arena rewrite)
    PH_MISC,            // memsets, setup kernels
    PH_COUNT
};

// ---- data-independent tree topology (pure function of n, minLeaf, maxDepth) ---------------------------
// Internal.hs:289 (leaf iff lev >= maxDepth || size <= minLeaf), :495,:503 (left = n div 2, right = rest).
struct Topology {
    int64_t n = 0;
    int maxDepth = 0, minLeaf = 0;
    int nlevels = 0;                    // depths 0 .. nlevels-1 hold at least one node
    int L_eff = 0;                      // depths 0 .. L_eff-1 hold at least one internal node (need projections)
    std::vector<uint32_t> start, size;  // per node (BFS id)
    std::vector<int32_t> child;         // left child BFS id or -1
    std::vector<int32_t> depth;
    std::vector<int64_t> level_off;     // nlevels+1
    std::vector<uint32_t> lvl_maxsize;  // per level
    int64_t nnodes() const { return (int64_t)start.size(); }
};
void build_topology(Topology& tp, int64_t n, int maxDepth, int minLeaf);

struct ProfEvent { int phase; cudaEvent_t a, b; };

// One level-synchronous build over an explicit topology (build.cu: rpf_run_job).  The batch build is one job over all
// points; the streaming build (stream.cu) runs one job per data chunk over the part of the current tree the chunk
// descends through.  Key row (t, l) of the job lives at keys + (t * Lk + l) * ks and holds the job's points 0..n-1.
struct BuildJob {
    const Topology* tp = nullptr;                                             // host topology (BFS ids, root at level 0)
    const uint32_t* d_start = nullptr; const uint32_t* d_size = nullptr; const int32_t* d_child = nullptr;   // device copies
    int64_t n = 0, ks = 0, ps = 0, ns = 0;      // points; key-row stride; stride between trees in perm; nodes-per-tree stride
    int Lk = 0;                                 // key rows per tree
    const unsigned long long* keys = nullptr;   // order-preserving images of the projections
    const unsigned long long* kmin = nullptr;   // [tg][Lk] range of the job's keys per (tree, level)
    const unsigned long long* kmax = nullptr;
    uint32_t* perm = nullptr;                   // [tg][ps] out: point ids 0..n-1, leaves left to right
    double *thr = nullptr, *mlo = nullptr, *mhi = nullptr;   // [..][ns] out, indexed (gt0 + t) * ns + node
    int gt0 = 0, tg = 0;
    bool stream_to_sink = false;                // batch build from host with an export sink: copy perm per bottom tree group
    int sink_ev0 = 0, sink_groups = 8;          // ... using the handle's events sink_ev[sink_ev0 .. sink_ev0 + sink_groups)
    // concurrent branches (rpf_build_impl): this job uses slice ws_part of ws_parts equal workspace slices, each sized
    // for ws_tg trees (0: tg)
    int ws_part = 0, ws_parts = 1, ws_tg = 0;
    bool order_exact = true;                    // out: every leaf is in the reference's order
};
struct JobGeom {
    int CAP = 0, L = 0, s = 0, s_top = 0, MAXTD = 0, bottom_levels = 0;
    int64_t NTOP = 0, HSZ = 1;
    bool order_exact = true, fast_bottom = false;
    std::vector<int> nb_level, smem_level;
};

// host tables of a plan, addressed by offset (256-byte aligned) so the block can live in the staging ring or in a cached
// device buffer
struct TableBuf {
    std::vector<char> bytes;
    size_t put(const void* p, size_t n) {
        const size_t off = (bytes.size() + 255) & ~(size_t)255;
        bytes.resize(off + n);
        if (n) memcpy(bytes.data() + off, p, n);
        return off;
    }
};
struct JobPlan {
    JobGeom G;
    int64_t n = 0, nn = 0;
    int nlevels = 0, nroots = 1, nnodes_s = 0, nlb = 0;
    uint32_t maxsize_s = 0;
    std::vector<int64_t> level_off;
    std::vector<char> lvl_all_internal;
    size_t off_range = (size_t)-1, off_lvlpv = (size_t)-1, off_nb = 0;   // offsets into the job's table block
};

// host -> device table staging: tables are written into page-locked memory and travel with one async copy per flush,
// so planning the next job overlaps the kernels of the previous one (no stream synchronisation in the build loop)
#define RPF_STAGE_SLOTS 8
#define RPF_MAX_BRANCH 4
struct StageSlot { char* h = nullptr; char* d = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; bool pending = false; };


// persistent device workspace (grow-only): no cudaMalloc/cudaFree inside the steady-state hot path
enum rpf_ws_slot {
    WS_KEYS = 0, WS_LABEL, WS_HIST, WS_SEL, WS_CAND, WS_CANDTOT, WS_PIVOTS, WS_FILL, WS_KMIN, WS_KMAX, WS_BINLO, WS_BINSC,
    WS_NBDEV, WS_RANGE, WS_LVLPV, WS_HPPACK,
    WS_Q, WS_KEYSQ, WS_SEGS, WS_CNT, WS_MAXCNT, WS_OUT_D, WS_OUT_I, WS_OUT_C, WS_BF_D, WS_TRUTH_D, WS_TRUTH_I, WS_RECALL,
    WS_CANDCNT, WS_CANDOFF, WS_CANDOUT, WS_MRG_D, WS_MRG_I, WS_MRG_C, WS_QHIST, WS_QORDER,
    WS_S_ARENA0, WS_S_ARENA1, WS_S_CPERM, WS_S_TMPN, WS_S_POOL, WS_QLAST, WS_PRIO, WS_WORKLIST, WS_PBIN, WS_BF_CV, WS_BF_CI, WS_BF_AUX,
    WS_BOT_REDO, WS_KNN_FB, WS_RR_LEAVES, WS_RR_XN, WS_RR_QN, WS_RR_HIST, WS_RR_START, WS_RR_CC, WS_RR_QOFF, WS_RR_AUX, WS_RR_FB, WS_RR_ENTQ, WS_RR_ENTD, WS_RR_DAP,
    WS_COUNT
};
struct WsBuf { void* p = nullptr; size_t cap = 0; };

struct RpfComm;     // multi.cu: this handle is one rank of a tree-sharded forest (NCCL communicator + rank / world)
struct RpfGroup;    // multi.cu: this handle is the in-process parent of one sub-handle per GPU (rpf_create_multi)

struct rpf_handle {
    int device = 0;
    RpfComm* comm = nullptr;
    RpfGroup* group = nullptr;
    int64_t x_pad_rows = 0;              // rows allocated behind row n of dX (the in-place all-gather of the last row block may spill)
    cudaStream_t stream = nullptr;
    // Batch build: the trees of a group are independent, so the group is cut into up to RPF_MAX_BRANCH contiguous parts whose
    // top / bottom phases run as concurrent branches on these streams (joined on `stream`): the many small kernels of one
    // branch (median pick / finish / ties) and the partial last waves of its big ones are filled by the other branches.
    cudaStream_t branch_stream[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t branch_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // [0..3] branch done, [4] fork point
    int branches = 0;                     // option "branches": 0 = chosen per build, 1 = off, 2..4 = forced
    cudaStream_t copy_stream = nullptr;   // rpf_build_from_host: row-block uploads overlapped with the projection
    cudaEvent_t copy_ev[17] = {nullptr};  // one per upload block + the "previous contents of dX are no longer read" event
    cudaStream_t gather_stream = nullptr; // communicator rank: the NVLink all-gathers of the row blocks (PCIe copies stay on copy_stream)
    cudaEvent_t up_ev[16] = {nullptr};    // block b's PCIe part has landed
    // export sink (rpf_set_export_sink): host buffers the forest is streamed into while rpf_build_from_host still runs --
    // the bottom phase is launched in tree groups and every group's slice of perm starts its D2H as soon as it is final
    double *sink_thr = nullptr, *sink_mlo = nullptr, *sink_mhi = nullptr; uint32_t* sink_perm = nullptr;
    cudaStream_t d2h_stream = nullptr;
    cudaEvent_t sink_ev[10] = {nullptr};  // [0..7] bottom groups, [8] whole build, [9] sink complete
    bool sink_pending = false;            // the sink holds (or is receiving) the forest of the last build
    bool sink_nodes_streamed = false;     // ... and the thr / mlo / mhi copies, tree group by tree group
    bool sink_perm_streamed = false;      // set by the job when it issued the perm copies itself
    std::string err;

    // points
    int64_t n = 0; int d = 0;
    const double* dX = nullptr; bool ownX = false; size_t x_bytes = 0;
    int32_t* d_xlast = nullptr;          // SVector data (rpf_set_points_sparse): last stored component of every row; else NULL

    // hyperplanes: CSR over (tree, level); host copy + device copy
    int T = 0, hpDepth = 0;
    std::vector<int64_t> hp_off; std::vector<int32_t> hp_idx; std::vector<double> hp_val;
    int64_t* d_hp_off = nullptr; int32_t* d_hp_idx = nullptr; double* d_hp_val = nullptr;
    void* d_hp_pack = nullptr;   // (val, idx) pairs, 16 bytes each, CSR order
    int32_t* d_hp_chunk = nullptr; int hp_chunk_d = 0; int64_t hp_chunk_rows = 0;   // long rows: nonzeros per column chunk (k_project_wide)
    void* proj_progs = nullptr;          // build.cu: ProjProgCache, fold programs of k_project_t per (tree group, levels, d, tile)
    void (*proj_progs_free)(void*) = nullptr;

    // topology
    Topology topo;
    uint32_t* d_node_start = nullptr; uint32_t* d_node_size = nullptr; int32_t* d_node_child = nullptr; int32_t* d_node_depth = nullptr;
    size_t topo_dev_nn = 0; std::vector<int32_t> topo_dev_child; std::vector<uint32_t> topo_dev_size;   // what the device copy holds
    // tree-group size of the last batch build (cudaMemGetInfo is only asked again when the shape changes)
    int64_t tg_key_n = -1; int tg_key_L = -1, tg_key_T = -1, tg_cached = 0;

    // forest
    bool built = false;
    double *d_thr = nullptr, *d_mlo = nullptr, *d_mhi = nullptr;   // [T][nnodes]
    uint32_t* d_perm = nullptr;                                       // [T][n]
    bool leaf_order_exact = true;
    int64_t stream_lost = 0;             // points dropped by the reference's empty-piece rule during a streaming build
    void* insert_session = nullptr;      // stream.cu: InsertSession of rpf_insert_begin / rpf_insert_chunk (owns dX while it lives)
    void (*insert_session_free)(void*) = nullptr;
    void* stream_plan = nullptr;         // cached plan of the last streaming build shape (stream.cu: StreamPlanAll)
    void (*stream_plan_free)(void*) = nullptr;
    size_t res_node_bytes = 0, res_perm_bytes = 0;
    int fused_top = 1;                   // option, bits: 1 = median-bin pick fused into the histogram kernel (jobs of >= 16 trees), 2 = one
                                         //         finish kernel per level instead of finish_warp -> finish -> ties (measured slower: off)
    int rerank_gemm = 1;                 // option: leaf-grouped FP64 tensor-core re-rank (rerank.cu): 0 = never, 1 = when it pays (d >= 512,
                                         //         >= 2 queries per leaf), 2 = whenever applicable (tests)
    int fused_pick_min_tg = 16;          // option: trees per job from which the histogram kernels also pick the median bins (ticket counter)
    int hist_big_chunk = 1;              // option: histogram kernels may take up to 57 344 points per CTA when that saves a wave
    int fuse_relabel_hist = 1;           // option: top-phase relabel of level l fused with the histogram of level l + 1 (k_top_relabel_hist)
    int bottom_select = 0;               // option: warp-per-node bottom kernel (k_bottom4: median select + partition on the levels whose
                                         //         children split again, sort only where Tips form); 0 = k_bottom3 everywhere
    int project_prefetch = 1;            // option: L2 prefetch of a later tile in the single-buffer projection kernel
    int project_pipe_maxh = 128;         // option: the pipelined projection kernel is used up to this many hyperplanes per launch
    int project_variant = 0;             // tuning hook: 0 = 1024 threads x 4 points/lane, 1 = 1024 x 2 (two CTAs/SM), 2 = 512 x 4
    bool no_query_order = false;         // test/tuning hook: answer queries in input order (no locality grouping)
    bool force_simple_topk = false;      // test hook: brute-force truth through the nine-pass radix select only
    int knn_f32_cfg[4] = {0, 0, 0, 0};  // tuning hook: ring stages, rows per stage, entry buffer, side region of k_knn_f32 (0 = default)
    int knn_filter32 = 1;                // option: fp32 filter pass in front of the exact re-rank (k_knn_f32): 0 = exact gather kernel only
    float* dX32 = nullptr; size_t x32_bytes = 0;     // fp32 image of X for the filter pass (built lazily by the first knn after the points change)
    const double* x32_src = nullptr; int64_t x32_n = -1; int x32_d = -1; uint64_t x32_version = 0;
    uint64_t x_version = 1;              // bumped by every entry point that changes the CONTENT of dX
    double* d_xmax = nullptr;            // [1] largest row norm of X (error margin of the filter)
    bool force_simple_knn = false;       // test hook: per-thread gather knn kernel instead of the TMA ring
    int top_chunk[3] = {0, 0, 0};        // tuning hook: points per CTA of the top-phase hist / compact / relabel kernels (0 = chosen per launch)
    bool lean_top = true;                // option "lean_top": 0 = generic top-phase compact / relabel kernels only (test hook)
    bool force_generic_bottom = false;   // test hook: run the generic (entry-table) bottom kernel
    bool bottom_words64 = false;         // test hook: 64-bit sort words in the fast bottom kernel even for <= 2048 slots

    // staging ring
    StageSlot stage[RPF_STAGE_SLOTS];
    int stage_cur = -1; size_t stage_off = 0, stage_flushed = 0;
    int stage_begin(size_t bytes);                 // next slot, sized for `bytes` of tables (+ alignment slack)
    void* stage_put_raw(const void* src, size_t bytes);   // returns the DEVICE address the bytes will have after the flush
    template <typename T> T* stage_put(const T* src, size_t count) { return (T*)stage_put_raw(src, count * sizeof(T)); }
    int stage_flush();
    void stage_free_all();

    // Every mutation that can move a device buffer or change the build's shape bumps cfg_epoch; the cached batch plan /
    // captured build graph are only replayed while the epoch they were made under is still current.
    uint64_t cfg_epoch = 1;
    uint64_t last_build_epoch = 0;        // epoch at the end of the previous batch build (0: none)
    void* batch_plan = nullptr;           // build.cu: BatchPlan (job plan + device-resident tables of the current shape)
    void (*batch_plan_free)(void*) = nullptr;
    cudaGraphExec_t build_graph = nullptr;   // the whole batch build (memsets, projection, top and bottom phases) as one graph
    uint64_t graph_epoch = 0;
    int64_t graph_launches = 0;
    bool capturing = false;
    bool use_graphs = true;               // option "cuda_graph"
    int64_t topo_key_n = -1; int topo_key_maxd = -1, topo_key_minl = -1;   // h->topo == build_topology(these)

    // workspace
    WsBuf ws[WS_COUNT];
    size_t ws_bytes = 0;
    void* ws_get(int slot, size_t bytes);   // nullptr on allocation failure (err is set)
    void ws_free_all();
    int64_t hp_pack_rows = 0;

    // tuning
    int bottom_cap = 1024;   // measured optimum on B200 for 1M x 128: 512..2048 are within 10%

    // measurement
    bool profiling = false;
    double last_ms = 0.0;
    double phase_ms[PH_COUNT] = {0};
    int64_t phase_launches[PH_COUNT] = {0};
    int64_t launches = 0;
    std::vector<ProfEvent> pending;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;

    void prof_reset();
    void prof_begin(int phase);
    void prof_end(int phase);
    void call_begin();
    int  call_end();    // syncs the stream, accumulates timings; returns RPF_OK or error
};

int rpf_fail(rpf_handle* h, int code, const std::string& msg);

#define RPF_CUDA(h, expr)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess)                                                                         \
            return rpf_fail((h), RPF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
    } while (0)

// Launch wrapper: counts launches, brackets with profiling events when enabled.
#define RPF_LAUNCH(h, phase, kern, grid, block, smem, ...)                                             \
    do {                                                                                               \
        (h)->prof_begin(phase);                                                                        \
        kern<<<(grid), (block), (smem), (h)->stream>>>(__VA_ARGS__);                                   \
        (h)->prof_end(phase);                                                                          \
        cudaError_t _e = cudaGetLastError();                                                           \
        if (_e != cudaSuccess)                                                                         \
            return rpf_fail((h), RPF_ERR_CUDA, std::string(#kern) + " launch: " + cudaGetErrorString(_e)); \
    } while (0)

// implemented in build.cu / stream.cu / query.cu
int rpf_build_impl(rpf_handle* h, const double* hostX);
int rpf_alloc_forest(rpf_handle* h, int64_t nn, int64_t n);
void rpf_job_geometry(const Topology& tp, int cap_cfg, int Lk, JobGeom& G);
size_t rpf_job_ws_per_tree(const JobGeom& G, int64_t n);
int rpf_run_job(rpf_handle* h, BuildJob& J);
void rpf_plan_job(const Topology& tp, int cap_cfg, int Lk, bool force_generic, TableBuf& TB, JobPlan& P);
int rpf_launch_job(rpf_handle* h, BuildJob& J, const JobPlan& P, const char* tab);
int rpf_bottom_fast_levels();
int rpf_project_launch(rpf_handle* h, int phase, const double* dX, int64_t n, int t0, int Tg, int L, bool ord, void* out,
                       int64_t ostride, unsigned long long* kmin, unsigned long long* kmax);
int rpf_upload_topology(rpf_handle* h);
int rpf_build_stream_impl(rpf_handle* h, int maxDepth, int minLeaf, int64_t chunk);
int rpf_insert_begin_impl(rpf_handle* h, int d, int maxDepth, int minLeaf);
int rpf_insert_chunk_impl(rpf_handle* h, const double* Xc, int64_t m);
int rpf_insert_end_impl(rpf_handle* h);
void rpf_insert_drop(rpf_handle* h);      // closes an insert session (every other way of setting points / building calls it)
int rpf_project_queries(rpf_handle* h, const double* dQ, int64_t nq, double* d_keysQ);
int rpf_knn_impl(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int k, int dedup, double* dist, uint32_t* ids, int32_t* count, bool out_dev);
int rpf_knn_h_impl(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int k, int cap, double* dist, uint32_t* ids, int32_t* count);
int rpf_candidates_impl(rpf_handle* h, const double* Q, int64_t nq, int t, int64_t* off_out, const int64_t* off_in, uint32_t* ids);
int rpf_recall_impl(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int k, double* recall_sum);
int rpf_brute_knn_impl(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int k, double* dist, uint32_t* ids);
int rpf_merge_impl(rpf_handle* h, int G, int64_t nq, int k, int dedup, const double* dist, const uint32_t* ids,
                   const int32_t* count, double* dist_out, uint32_t* ids_out, int32_t* count_out, bool in_dev);

// ---- leaf-grouped tensor-core re-rank (rerank.cu) ------------------------------------------------------------
bool rpf_rerank_gemm_wanted(const rpf_handle* h, int64_t nq, int k, int dedup);
int rpf_rerank_gemm(rpf_handle* h, const double* dQ, int64_t nq, int S, const uint32_t* segs, const uint32_t* cnt, int k,
                    double* ddist, uint32_t* dids, int32_t* dcount, uint32_t* n_fallback);

// ---- multi-GPU (multi.cu) ---------------------------------------------------------------------------
// rank / world of a handle (0 / 1 without a communicator)
int rpf_comm_rank(const rpf_handle* h);
int rpf_comm_world(const rpf_handle* h);
void rpf_comm_free(rpf_handle* h);
// in-place all-gather on `stream`: every rank contributed `bytes` bytes at buf + rank * bytes
int rpf_comm_allgather(rpf_handle* h, void* buf, size_t bytes, cudaStream_t stream);
// the rows rank r supplies to a row-sharded upload of n rows: [r * per, min(n, (r + 1) * per)), per = ceil(n / world)
int rpf_upload_rows(rpf_handle* h, const double* hostX, int64_t r0, int64_t nr, cudaStream_t stream,
                    cudaStream_t gather_stream = nullptr, cudaEvent_t up_ev = nullptr);   // build.cu
// group parent (rpf_create_multi): every public entry point forwards here when h->group is set
void rpf_group_free(rpf_handle* h);

// ---- bottom phase launch arguments (build.cu kernels; also filled by stream.cu for Tip re-splits) --------
typedef unsigned long long ull;
struct BottomArgs {
    int64_t ks, ps, nn_all;              // key-row stride, stride between the trees' slices of perm, nodes per tree
    int L, s, nlb, gt0, first_gid;      // nlb = levels recorded per node in `range`
    int given_order;                     // 1: the incoming order of perm IS the reference's order (streaming re-split of
                                         //    a Tip, Internal.hs:287-297); 0: establish (key_{s-1}, ..., key_0, row id)
    const ull* keys;                     // [Tg][L][n]
    uint32_t* perm;                      // [Tg][n]  (in: node segments in any order; out: final leaf order)
    const int32_t* child;
    const uint32_t* nstart;
    const uint32_t* nsize;
    const int2* range;                   // [nodes at level s][nlb]: BFS id range of the descendants that split
    const uint32_t* lvl_pv;              // per level: next_pow2(max node size)
    const ull* kmin;                     // [Tg][L] key range per (tree, level) (may be a sample's range, may be NULL):
    const ull* kmax;                     //         seeds the key prefixes of 32-bit sort words
    double *thr, *mlo, *mhi;
    const uint8_t* only;                 // k_bottom3 as the second pass of k_bottom4: [tg][nroots], run a node only if its flag is set
};
int rpf_bottom_launch(rpf_handle* h, const BottomArgs& B, int nroots, int tg, bool fast, unsigned max_root, int levels);

// ---- device helpers ---------------------------------------------------------------------------------
#ifdef __CUDACC__
// order-preserving map double -> uint64 (total order == IEEE order on non-NaN; -0 canonicalised to +0)
__device__ __forceinline__ uint64_t f2ord(double x) {
    uint64_t b = (uint64_t)__double_as_longlong(x);
    if (x == 0.0) b = 0ull;
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord2f(uint64_t o) {
    uint64_t b = (o & 0x8000000000000000ull) ? (o & 0x7fffffffffffffffull) : ~o;
    return __longlong_as_double((long long)b);
}
#define ORD_NONE_LO 0ull                       /* below every finite key */
#define ORD_NONE_HI 0xffffffffffffffffull      /* above every finite key */
#endif
