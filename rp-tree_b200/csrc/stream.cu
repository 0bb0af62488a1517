// stream.cu -- streaming (multi-chunk) forest construction on sm_100a.
//
// Replaces, for dense Double data, the reference's incremental path
//   forest / tree           src/Data/RPTree/Conduit.hs:58-121   (chunksOf n .| foldl insertMulti)
//   insertMulti / insert    src/Data/RPTree/Internal.hs:243-297 (Bin case :274-285, Tip case :287-297)
//   Margin semigroup        src/Data/RPTree/Internal.hs:75-89   (Max on the low side, Min on the high side)
//
// What the reference does per chunk xs and per tree (Internal.hs:265-297):
//   Bin thr0 margin0 l r : partitionAtMedian r_lev xs (the CHUNK's own positional median, not thr0) ->
//                          thr' = (thr0 + thr) / 2, margin' = margin0 <> margin, recurse with the two halves;
//                          an EMPTY piece reaching a Bin yields `Tip () mempty`: the whole subtree is dropped.
//   Tip xs0              : xs' = xs <> xs0 (new points first); Tip again if lev >= maxDepth or |xs'| <= minLeaf,
//                          else split xs' recursively as in the batch build (children start from empty Tips).
// The routing of a chunk never looks at the stored thresholds, only at piece sizes, so -- exactly like the batch
// build -- the SHAPE of the tree after every chunk is a pure function of (n, chunk, maxDepth, minLeaf) and is the same
// for every tree.  The host therefore plans each chunk on sizes alone (StreamPlanner) and the device executes the plan
// for all trees at once:
//   1. k_project          keys of the chunk's points for every (tree, level) -> persistent key store [Tg][Lk][n]
//   2. rpf_run_job        the chunk descends through the current Bin structure: the batch machinery (top + bottom
//                         phases of build.cu) on the "chunk tree" CT = current tree cut at its Tips
//   3. k_pool_update      thr' / margin' of every Bin the chunk went through (mode avg) or of new Bins (mode set)
//   4. k_tip_concat       every Tip's new content = piece ++ old content, written to the other leaf arena in the
//                         final left-to-right layout
//   5. rpf_bottom_launch  Tips that outgrew minLeaf are split in place (given_order: the concatenation order is the
//                         stable sort's incoming order), grouped by depth
// No host synchronisation happens inside the chunk loop: plans travel through the page-locked staging ring.
#include "rpf_internal.h"
#include "rpf_device.cuh"
#include <algorithm>
#include <cstring>

// =====================================================================================================
// host planner (sizes only; shared by the GPU path and by rpf_topology_plan_chunked)
// =====================================================================================================
namespace {

struct SNode {
    int32_t l = -1, r = -1;     // children (pool ids); l < 0: Tip
    int32_t depth = 0;
    int64_t cnt = 0;            // Tip: points held
    int64_t off = 0;            // Tip: offset of its content in the current leaf arena
};

struct TipCopy { uint32_t dst, src_old, n_old, src_new, n_new; };

struct RtGroup { int depth = 0, first = 0, count = 0, nlb = 0; uint32_t max_root = 0; bool fast = true; };

struct ChunkPlan {
    Topology ct;                         // chunk tree (BFS); leaves = Tips (or wiped Bins) of the current tree
    std::vector<int32_t> upd_g, upd_u;   // CT nodes that split -> pool ids
    bool ct_set = false;                 // CT results replace (fresh tree) instead of averaging
    std::vector<TipCopy> copies;
    // re-split forest RT: roots first (grouped by depth), then descendants level by level; siblings adjacent
    std::vector<int32_t> rt_child, rt_pool, rt_depth;
    std::vector<uint32_t> rt_start, rt_size;
    std::vector<int32_t> rt_upd_g, rt_upd_u;
    std::vector<RtGroup> groups;
    int64_t kept = 0;                    // points held by the tree after this chunk
};

struct StreamPlanner {
    int64_t n = 0; int maxDepth = 0, minLeaf = 0;
    std::vector<SNode> pool;
    int32_t root = 0;
    int64_t lost = 0;                    // points dropped by the reference's empty-piece rule (Internal.hs:279)
    std::string err;

    void begin(int64_t n_, int maxd, int minl) {
        n = n_; maxDepth = maxd; minLeaf = minl; pool.clear(); pool.emplace_back(); root = 0; lost = 0; err.clear();
    }
    int32_t new_node(int depth) { pool.emplace_back(); pool.back().depth = depth; return (int32_t)pool.size() - 1; }
    int64_t subtree_points(int32_t u) const {
        int64_t tot = 0; std::vector<int32_t> st{u};
        while (!st.empty()) { const int32_t v = st.back(); st.pop_back(); if (pool[v].l < 0) tot += pool[v].cnt; else { st.push_back(pool[v].l); st.push_back(pool[v].r); } }
        return tot;
    }

    // Plans the insertion of a chunk of m points.  Returns false on an unsupported shape (err is set).
    bool plan_chunk(int64_t m, ChunkPlan& P) {
        P = ChunkPlan();
        Topology& ct = P.ct;
        ct.n = m; ct.maxDepth = maxDepth; ct.minLeaf = minLeaf;
        std::vector<int32_t> ct_pool;
        struct Pend { int32_t u; uint32_t src_new, n_new; };
        std::vector<Pend> tips;            // CT leaves sitting on Tips: piece location inside the chunk's perm
        const bool fresh = pool[root].l < 0 && pool[root].cnt == 0 && m > minLeaf && maxDepth > 0;
        std::vector<int32_t> resplit;      // Tips that split after this chunk (pool ids)
        std::vector<char> is_new_tip;
        if (fresh) {
            // first chunk into an empty tree: the Tip case splits xs recursively == the batch build of the chunk
            build_topology(ct, m, maxDepth, minLeaf);
            ct_pool.assign((size_t)ct.nnodes(), -1);
            ct_pool[0] = root;
            for (int64_t g = 0; g < ct.nnodes(); ++g) {
                const int32_t u = ct_pool[g];
                pool[u].depth = ct.depth[g];
                if (ct.child[g] >= 0) {
                    const int32_t a = new_node(ct.depth[g] + 1), b = new_node(ct.depth[g] + 1);
                    pool[u].l = a; pool[u].r = b; pool[u].cnt = 0;
                    ct_pool[ct.child[g]] = a; ct_pool[ct.child[g] + 1] = b;
                    P.upd_g.push_back((int32_t)g); P.upd_u.push_back(u);
                } else {
                    pool[u].cnt = 0;       // filled by the layout pass below
                    tips.push_back(Pend{u, ct.start[g], ct.size[g]});
                }
            }
            P.ct_set = true;
        } else {
            ct.start.push_back(0); ct.size.push_back((uint32_t)m); ct.child.push_back(-1); ct.depth.push_back(0);
            ct_pool.push_back(root);
            ct.level_off.push_back(0);
            int64_t lo = 0, hi = 1; int lev = 0;
            while (lo < hi) {
                uint32_t mx = 0; bool any_internal = false;
                for (int64_t g = lo; g < hi; ++g) {
                    const uint32_t sz = ct.size[g];
                    mx = std::max(mx, sz);
                    const int32_t u = ct_pool[g];
                    if (pool[u].l >= 0) {
                        if (lev >= maxDepth) { err = "internal: Bin below maxDepth"; return false; }   // Internal.hs:275 (unreachable)
                        if (sz < 1) {
                            // partitionAtMedian r [] = Nothing -> Tip () mempty: the subtree and its points vanish
                            lost += subtree_points(u);
                            pool[u].l = pool[u].r = -1; pool[u].cnt = 0; pool[u].off = 0;
                            tips.push_back(Pend{u, ct.start[g], 0});
                        } else {
                            any_internal = true;
                            const uint32_t nh = sz / 2;
                            ct.child[g] = (int32_t)ct.start.size();
                            ct.start.push_back(ct.start[g]);      ct.size.push_back(nh);      ct.child.push_back(-1); ct.depth.push_back(lev + 1);
                            ct.start.push_back(ct.start[g] + nh); ct.size.push_back(sz - nh); ct.child.push_back(-1); ct.depth.push_back(lev + 1);
                            ct_pool.push_back(pool[u].l); ct_pool.push_back(pool[u].r);
                            P.upd_g.push_back((int32_t)g); P.upd_u.push_back(u);
                        }
                    } else {
                        tips.push_back(Pend{u, ct.start[g], sz});
                    }
                }
                ct.lvl_maxsize.push_back(mx);
                ct.level_off.push_back(hi);
                ++lev;
                if (any_internal) ct.L_eff = lev;
                lo = hi; hi = (int64_t)ct.start.size();
            }
            ct.nlevels = lev;
        }

        // ---- Tip case: xs' = xs <> xs0; split when it outgrew minLeaf (Internal.hs:287-297)
        std::vector<uint32_t> pend_src(pool.size(), 0), pend_n(pool.size(), 0);
        std::vector<char> touched(pool.size(), 0), is_root(pool.size(), 0);
        std::vector<int64_t> newcnt(pool.size(), 0);
        for (const Pend& t : tips) { pend_src[t.u] = t.src_new; pend_n[t.u] = t.n_new; touched[t.u] = 1; newcnt[t.u] = pool[t.u].cnt + t.n_new; }
        if (!fresh) {
            for (const Pend& t : tips) {
                const int32_t u = t.u;
                const int64_t tot = newcnt[u];
                if (pool[u].depth >= maxDepth || tot <= minLeaf) continue;
                if (tot > 8192) {
                    err = "build_chunked: a Tip of more than 8192 points must be re-split (minLeaf > 4095 with chunk < n is not supported)";
                    return false;
                }
                resplit.push_back(u);
            }
        }
        // new subtrees: sizes follow Internal.hs:289,495,503 (leaf iff lev >= maxDepth or size <= minLeaf; left = size div 2)
        // RT ids: roots sorted by depth (stable: left-to-right inside a depth), then breadth first over all roots.
        std::stable_sort(resplit.begin(), resplit.end(), [&](int32_t a, int32_t b) { return pool[a].depth < pool[b].depth; });
        {
            const size_t R = resplit.size();
            P.rt_child.assign(R, -1); P.rt_pool.assign(resplit.begin(), resplit.end()); P.rt_size.resize(R); P.rt_depth.resize(R);
            for (size_t i = 0; i < R; ++i) { P.rt_size[i] = (uint32_t)newcnt[resplit[i]]; P.rt_depth[i] = pool[resplit[i]].depth; }
            size_t lo = 0, hi = R;
            while (lo < hi) {
                for (size_t g = lo; g < hi; ++g) {
                    const uint32_t sz = P.rt_size[g]; const int dl = P.rt_depth[g];
                    if (dl >= maxDepth || (int64_t)sz <= (int64_t)minLeaf) continue;
                    const uint32_t nh = sz / 2;
                    const int32_t u = P.rt_pool[g];
                    const int32_t a = new_node(dl + 1), b = new_node(dl + 1);
                    pool[u].l = a; pool[u].r = b;
                    pool[a].cnt = nh; pool[b].cnt = sz - nh;
                    P.rt_child[g] = (int32_t)P.rt_size.size();
                    P.rt_size.push_back(nh);      P.rt_child.push_back(-1); P.rt_pool.push_back(a); P.rt_depth.push_back(dl + 1);
                    P.rt_size.push_back(sz - nh); P.rt_child.push_back(-1); P.rt_pool.push_back(b); P.rt_depth.push_back(dl + 1);
                    P.rt_upd_g.push_back((int32_t)g); P.rt_upd_u.push_back(u);
                }
                lo = hi; hi = P.rt_size.size();
            }
            P.rt_start.assign(P.rt_size.size(), 0);
            for (size_t i = 0; i < R; ++i) is_root[resplit[i]] = 1;
        }

        // ---- new left-to-right layout of the leaf arena + copy descriptors
        std::vector<int64_t> root_region(pool.size(), -1);
        {
            int64_t cur = 0;
            struct Fr { int32_t u; bool inside; };
            std::vector<Fr> st; st.push_back(Fr{root, false});
            while (!st.empty()) {
                const Fr f = st.back(); st.pop_back();
                SNode& N = pool[f.u];
                bool inside = f.inside;
                if (!inside && f.u < (int32_t)touched.size() && touched[f.u]) {
                    // a Tip of the tree as the chunk found it (possibly the root of a new subtree now)
                    if (cur + newcnt[f.u] > (int64_t)0xffffffffu) { err = "arena offset overflow"; return false; }
                    P.copies.push_back(TipCopy{(uint32_t)cur, (uint32_t)N.off, (uint32_t)(newcnt[f.u] - pend_n[f.u]), pend_src[f.u], pend_n[f.u]});
                    if (N.l >= 0) { root_region[f.u] = cur; inside = true; }
                    else { N.off = cur; N.cnt = newcnt[f.u]; cur += N.cnt; continue; }
                }
                if (N.l >= 0) { st.push_back(Fr{N.r, inside}); st.push_back(Fr{N.l, inside}); }
                else if (inside) { N.off = cur; cur += N.cnt; }      // new Tip inside a re-split region (cnt set above)
                else { err = "internal: Tip not reached by the chunk"; return false; }
            }
            P.kept = cur;
        }
        // RT starts (parents precede children in RT order)
        for (size_t g = 0; g < P.rt_size.size(); ++g) {
            if (g < resplit.size()) P.rt_start[g] = (uint32_t)root_region[resplit[g]];
            if (P.rt_child[g] >= 0) {
                const int32_t c = P.rt_child[g];
                P.rt_start[c] = P.rt_start[g];
                P.rt_start[c + 1] = P.rt_start[g] + P.rt_size[c];
            }
        }
        // groups of roots by depth
        for (size_t i = 0; i < resplit.size();) {
            RtGroup G; G.depth = P.rt_depth[i]; G.first = (int)i;
            size_t j = i;
            while (j < resplit.size() && P.rt_depth[j] == G.depth) { G.max_root = std::max(G.max_root, P.rt_size[j]); ++j; }
            G.count = (int)(j - i);
            P.groups.push_back(G);
            i = j;
        }
        return true;
    }

    // final tree in canonical BFS form; pool_of[g] = pool id of BFS node g
    void final_topology(Topology& tp, std::vector<int32_t>& pool_of) const {
        tp = Topology();
        tp.n = n; tp.maxDepth = maxDepth; tp.minLeaf = minLeaf;
        pool_of.clear();
        pool_of.push_back(root);
        tp.child.push_back(-1); tp.depth.push_back(0); tp.start.push_back(0); tp.size.push_back(0);
        tp.level_off.push_back(0);
        int64_t lo = 0, hi = 1; int lev = 0;
        while (lo < hi) {
            bool any_internal = false;
            for (int64_t g = lo; g < hi; ++g) {
                const SNode& N = pool[pool_of[g]];
                if (N.l >= 0) {
                    any_internal = true;
                    tp.child[g] = (int32_t)pool_of.size();
                    pool_of.push_back(N.l); pool_of.push_back(N.r);
                    for (int q = 0; q < 2; ++q) { tp.child.push_back(-1); tp.depth.push_back(lev + 1); tp.start.push_back(0); tp.size.push_back(0); }
                } else {
                    tp.start[g] = (uint32_t)N.off; tp.size[g] = (uint32_t)N.cnt;
                }
            }
            tp.level_off.push_back(hi);
            ++lev;
            if (any_internal) tp.L_eff = lev;
            lo = hi; hi = (int64_t)pool_of.size();
        }
        tp.nlevels = lev;
        for (int64_t g = (int64_t)pool_of.size() - 1; g >= 0; --g)
            if (tp.child[g] >= 0) { const int32_t c = tp.child[g]; tp.start[g] = tp.start[c]; tp.size[g] = tp.size[c] + tp.size[c + 1]; }
        tp.lvl_maxsize.assign(tp.nlevels, 0);
        for (int l = 0; l < tp.nlevels; ++l)
            for (int64_t g = tp.level_off[l]; g < tp.level_off[l + 1]; ++g) tp.lvl_maxsize[l] = std::max(tp.lvl_maxsize[l], tp.size[g]);
    }
};

// BFS id ranges, per root and relative level, of the descendants that exist at that level (generic bottom kernel)
void make_ranges(const std::vector<int32_t>& child, int first, int nroots, int nlb, std::vector<int2>& rg) {
    rg.assign((size_t)nroots * nlb, make_int2(0, 0));
    for (int e = 0; e < nroots; ++e) {
        int64_t lo = first + e, hi = lo + 1;
        for (int j = 0; j < nlb; ++j) {
            int64_t fi = -1, li = -1;
            for (int64_t g = lo; g < hi; ++g) if (child[g] >= 0) { if (fi < 0) fi = g; li = g; }
            if (fi < 0) break;
            rg[(size_t)e * nlb + j] = make_int2((int)lo, (int)hi);
            lo = child[fi]; hi = (int64_t)child[li] + 2;
        }
    }
}

}  // namespace

// =====================================================================================================
// kernels
// =====================================================================================================
// thr/margins of the nodes a job produced (tree-major temp arrays) -> node-major pool arrays.
// mode 0: set (new Bin).  mode 1: streaming update of an existing Bin (Internal.hs:281-282):
//   thr' = (thr0 + thr) / 2;  margin' = margin0 <> margin = Margin (max lo0 lo) (min hi0 hi)  with Haskell's
//   max x y = if x <= y then y else x, min x y = if x <= y then x else y.
__global__ void k_pool_update(const int32_t* __restrict__ src_g, const int32_t* __restrict__ pool_u, int cnt, int tg, int64_t ns,
                              const double* __restrict__ thr, const double* __restrict__ mlo, const double* __restrict__ mhi,
                              double* __restrict__ pthr, double* __restrict__ pmlo, double* __restrict__ pmhi, int mode) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)cnt * tg) return;
    const int i = (int)(idx / tg), t = (int)(idx % tg);
    const int64_t s = (int64_t)t * ns + src_g[i], o = (int64_t)pool_u[i] * tg + t;
    const double v = thr[s], lo = mlo[s], hi = mhi[s];
    if (mode == 0) { pthr[o] = v; pmlo[o] = lo; pmhi[o] = hi; return; }
    const double v0 = pthr[o], lo0 = pmlo[o], hi0 = pmhi[o];
    pthr[o] = __ddiv_rn(__dadd_rn(v0, v), 2.0);
    pmlo[o] = (lo0 <= lo) ? lo : lo0;
    pmhi[o] = (hi0 <= hi) ? hi0 : hi;
}

// new content of every Tip: the chunk's piece (ids local to the chunk, + row0) followed by the old content
// (Internal.hs:288 `xs <> xs0`).  One warp per (Tip, tree).
__global__ void __launch_bounds__(256) k_tip_concat(const TipCopy* __restrict__ cp, int ncopy, int tg,
                                                     const uint32_t* __restrict__ piece, int64_t piece_stride, uint32_t row0,
                                                     const uint32_t* __restrict__ old_arena, uint32_t* __restrict__ new_arena, int64_t astride) {
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= (int64_t)ncopy * tg) return;
    const int i = (int)(wid / tg), t = (int)(wid % tg);
    const TipCopy c = cp[i];
    uint32_t* dst = new_arena + (int64_t)t * astride + c.dst;
    const uint32_t* pn = piece + (int64_t)t * piece_stride + c.src_new;
    const uint32_t* po = old_arena + (int64_t)t * astride + c.src_old;
    for (uint32_t j = lane; j < c.n_new; j += 32) dst[j] = pn[j] + row0;
    for (uint32_t j = lane; j < c.n_old; j += 32) dst[c.n_new + j] = po[j];
}

// node-major pool -> canonical [T][nodes] forest arrays
__global__ void k_pool_export(const int32_t* __restrict__ pool_of, const int32_t* __restrict__ child, int64_t nn, int tg, int gt0,
                              const double* __restrict__ pthr, const double* __restrict__ pmlo, const double* __restrict__ pmhi,
                              double* __restrict__ thr, double* __restrict__ mlo, double* __restrict__ mhi) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nn * tg) return;
    const int64_t g = idx % nn; const int t = (int)(idx / nn);
    const int64_t o = (int64_t)(gt0 + t) * nn + g;
    if (child[g] < 0) { thr[o] = 0.0; mlo[o] = 0.0; mhi[o] = 0.0; return; }
    const int64_t s = (int64_t)pool_of[g] * tg + t;
    thr[o] = pthr[s]; mlo[o] = pmlo[s]; mhi[o] = pmhi[s];
}

// =====================================================================================================
// host orchestration
// =====================================================================================================
#define WSX(h, var, type, slot, bytes)                                  \
    type* var = (type*)(h)->ws_get((slot), (bytes));                    \
    if (!var) return RPF_ERR_NOMEM;

int rpf_build_stream_impl(rpf_handle* h, int maxDepth, int minLeaf, int64_t chunk) {
    const int64_t n = h->n;
    const int T = h->T;
    const int Lk = std::max(maxDepth, 1);
    h->leaf_order_exact = true;

    // ---- tree group size: persistent key store [Tg][Lk][n] + two leaf arenas + per-chunk job workspace
    size_t freeB = 0, totalB = 0;
    RPF_CUDA(h, cudaMemGetInfo(&freeB, &totalB));
    const int64_t mmax = std::min(chunk, n);
    const size_t per_tree = (size_t)Lk * n * 8 + (size_t)n * 8 + (size_t)mmax * 32 + ((size_t)1 << 20);
    const size_t budget = (size_t)((double)(freeB + h->ws_bytes) * 0.7);
    if (per_tree > budget) return rpf_fail(h, RPF_ERR_NOMEM, "not enough device memory for one tree's keys");
    const int Tg = (int)std::min<size_t>((size_t)T, std::max<size_t>(1, budget / per_tree));

    WSX(h, keys, ull, WS_KEYS, (size_t)Tg * Lk * n * 8);
    WSX(h, kmin, ull, WS_KMIN, (size_t)Tg * Lk * 8);
    WSX(h, kmax, ull, WS_KMAX, (size_t)Tg * Lk * 8);
    WSX(h, arena0, uint32_t, WS_S_ARENA0, (size_t)Tg * n * 4);
    WSX(h, arena1, uint32_t, WS_S_ARENA1, (size_t)Tg * n * 4);
    WSX(h, cperm, uint32_t, WS_S_CPERM, (size_t)Tg * mmax * 4);

    if (h->stream_pool) { cudaFree(h->stream_pool); h->stream_pool = nullptr; }     // left over from a failed call
    StreamPlanner SP;
    Topology final_tp; std::vector<int32_t> pool_of;
    ChunkPlan P;
    for (int t0 = 0; t0 < T; t0 += Tg) {
        const int tg = std::min(Tg, T - t0);
        SP.begin(n, maxDepth, minLeaf);
        uint32_t* arena_old = arena0; uint32_t* arena_new = arena1;
        RPF_CUDA(h, cudaMemsetAsync(arena0, 0xff, (size_t)tg * n * 4, h->stream));     // slots past the kept points stay 0xffffffff
        RPF_CUDA(h, cudaMemsetAsync(arena1, 0xff, (size_t)tg * n * 4, h->stream));
        double* pthr = nullptr; double* pmlo = nullptr; double* pmhi = nullptr; size_t pool_cap = 0;
        for (int64_t row0 = 0; row0 < n; row0 += chunk) {
            const int64_t m = std::min(chunk, n - row0);
            if (!SP.plan_chunk(m, P)) return rpf_fail(h, RPF_ERR_UNSUPPORTED, SP.err);
            const Topology& ct = P.ct;
            const int64_t nnct = ct.nnodes(), nnrt = (int64_t)P.rt_size.size();

            // ---- pool arrays (node-major [cap][tg]); growth copies the prefix
            if (SP.pool.size() > pool_cap) {
                // without dropped subtrees the pool never exceeds min(2^(maxDepth+1), 2n) nodes: one allocation
                const size_t bound = (size_t)std::min<int64_t>(maxDepth < 40 ? ((int64_t)1 << (maxDepth + 1)) : (int64_t)1 << 41, 2 * n + 2) + 2;
                const size_t want = std::max<size_t>(std::max<size_t>(SP.pool.size() * 2, bound), 1024);
                double* nb = nullptr;
                RPF_CUDA(h, cudaMalloc(&nb, want * tg * 8 * 3));
                if (pthr) {
                    RPF_CUDA(h, cudaMemcpyAsync(nb, pthr, pool_cap * tg * 8, cudaMemcpyDeviceToDevice, h->stream));
                    RPF_CUDA(h, cudaMemcpyAsync(nb + want * tg, pmlo, pool_cap * tg * 8, cudaMemcpyDeviceToDevice, h->stream));
                    RPF_CUDA(h, cudaMemcpyAsync(nb + 2 * want * tg, pmhi, pool_cap * tg * 8, cudaMemcpyDeviceToDevice, h->stream));
                    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
                    cudaFree(pthr);
                }
                pthr = nb; pmlo = nb + want * tg; pmhi = nb + 2 * want * tg; pool_cap = want;
                h->stream_pool = nb;      // owned by the handle until the end of the call (freed on error paths too)
            }

            // ---- stage this chunk's tables (CT topology, update lists, copy descriptors, RT forest)
            size_t bytes = (size_t)nnct * 12 + (P.upd_g.size() + P.rt_upd_g.size()) * 8 + P.copies.size() * sizeof(TipCopy) + (size_t)nnrt * 12;
            std::vector<std::vector<int2>> g_rg(P.groups.size());
            std::vector<std::vector<uint32_t>> g_pv(P.groups.size());
            for (size_t gi = 0; gi < P.groups.size(); ++gi) {
                RtGroup& G = P.groups[gi];
                // depth of the deepest descendant below this group's roots
                int maxrel = 0;
                {
                    std::vector<int32_t> fr;
                    for (int e = 0; e < G.count; ++e) fr.push_back(G.first + e);
                    int rel = 0;
                    while (!fr.empty()) {
                        std::vector<int32_t> nx;
                        for (int32_t g : fr) if (P.rt_child[g] >= 0) { nx.push_back(P.rt_child[g]); nx.push_back(P.rt_child[g] + 1); }
                        if (!nx.empty()) maxrel = ++rel;
                        fr.swap(nx);
                    }
                }
                G.nlb = std::max(1, maxrel + 1);
                G.fast = minLeaf >= 1 && maxrel <= rpf_bottom_fast_levels() && !h->force_generic_bottom;
                if (!G.fast) {
                    make_ranges(P.rt_child, G.first, G.count, G.nlb, g_rg[gi]);
                    // next_pow2(max size) per absolute level among this group's descendants
                    g_pv[gi].assign((size_t)G.depth + G.nlb, 1);
                    std::vector<int32_t> fr;
                    for (int e = 0; e < G.count; ++e) fr.push_back(G.first + e);
                    for (int rel = 0; !fr.empty(); ++rel) {
                        uint32_t mx = 1; std::vector<int32_t> nx;
                        for (int32_t g : fr) { mx = std::max(mx, P.rt_size[g]); if (P.rt_child[g] >= 0) { nx.push_back(P.rt_child[g]); nx.push_back(P.rt_child[g] + 1); } }
                        uint32_t p2 = 1; while (p2 < mx) p2 <<= 1;
                        g_pv[gi][(size_t)G.depth + rel] = p2;
                        fr.swap(nx);
                    }
                    bytes += g_rg[gi].size() * sizeof(int2) + g_pv[gi].size() * 4 + 512;
                }
            }
            int rc = h->stage_begin(bytes + 16 * 256);
            if (rc) return rc;
            const uint32_t* d_ct_start = h->stage_put(ct.start.data(), ct.start.size());
            const uint32_t* d_ct_size = h->stage_put(ct.size.data(), ct.size.size());
            const int32_t* d_ct_child = h->stage_put(ct.child.data(), ct.child.size());
            const int32_t* d_upd_g = h->stage_put(P.upd_g.data(), P.upd_g.size());
            const int32_t* d_upd_u = h->stage_put(P.upd_u.data(), P.upd_u.size());
            const TipCopy* d_copies = h->stage_put(P.copies.data(), P.copies.size());
            const uint32_t* d_rt_start = h->stage_put(P.rt_start.data(), P.rt_start.size());
            const uint32_t* d_rt_size = h->stage_put(P.rt_size.data(), P.rt_size.size());
            const int32_t* d_rt_child = h->stage_put(P.rt_child.data(), P.rt_child.size());
            const int32_t* d_rt_upd_g = h->stage_put(P.rt_upd_g.data(), P.rt_upd_g.size());
            const int32_t* d_rt_upd_u = h->stage_put(P.rt_upd_u.data(), P.rt_upd_u.size());
            std::vector<const int2*> d_rg(P.groups.size(), nullptr);
            std::vector<const uint32_t*> d_pv(P.groups.size(), nullptr);
            for (size_t gi = 0; gi < P.groups.size(); ++gi)
                if (!P.groups[gi].fast) { d_rg[gi] = h->stage_put(g_rg[gi].data(), g_rg[gi].size()); d_pv[gi] = h->stage_put(g_pv[gi].data(), g_pv[gi].size()); }
            if (!d_ct_start || !d_ct_size || !d_ct_child) return rpf_fail(h, RPF_ERR_NOMEM, h->err);
            rc = h->stage_flush();
            if (rc) return rc;

            // ---- temp node arrays of the chunk job and of the re-splits
            WSX(h, tmpn, double, WS_S_TMPN, (size_t)tg * (nnct + nnrt + 2) * 8 * 3);
            double* cthr = tmpn; double* cmlo = cthr + (size_t)tg * nnct; double* cmhi = cmlo + (size_t)tg * nnct;
            double* rthr = cmhi + (size_t)tg * nnct; double* rmlo = rthr + (size_t)tg * nnrt; double* rmhi = rmlo + (size_t)tg * nnrt;

            // ---- 1. keys of the chunk's points
            RPF_CUDA(h, cudaMemsetAsync(kmin, 0xff, (size_t)tg * Lk * 8, h->stream));
            RPF_CUDA(h, cudaMemsetAsync(kmax, 0x00, (size_t)tg * Lk * 8, h->stream));
            if (maxDepth > 0) {
                rc = rpf_project_launch(h, PH_PROJECT, h->dX + row0 * (int64_t)h->d, m, t0, tg, Lk, true, keys + row0, n, kmin, kmax);
                if (rc) return rc;
            }
            // ---- 2. the chunk descends through the Bins of the current tree
            BuildJob J{};
            J.tp = &ct; J.d_start = d_ct_start; J.d_size = d_ct_size; J.d_child = d_ct_child;
            J.n = m; J.ks = n; J.ps = mmax; J.ns = nnct; J.Lk = Lk;
            J.keys = keys + row0; J.kmin = kmin; J.kmax = kmax;
            J.perm = cperm; J.thr = cthr; J.mlo = cmlo; J.mhi = cmhi; J.gt0 = 0; J.tg = tg;
            rc = rpf_run_job(h, J);
            if (rc) return rc;
            if (!J.order_exact) h->leaf_order_exact = false;
            // ---- 3. thr' = (thr0 + thr) / 2, margin' = margin0 <> margin
            if (!P.upd_g.empty()) {
                const int64_t tot = (int64_t)P.upd_g.size() * tg;
                RPF_LAUNCH(h, PH_STREAM, k_pool_update, (unsigned)((tot + 255) / 256), 256, 0, d_upd_g, d_upd_u, (int)P.upd_g.size(), tg, nnct,
                           cthr, cmlo, cmhi, pthr, pmlo, pmhi, P.ct_set ? 0 : 1);
            }
            // ---- 4. Tip contents: piece ++ old
            if (!P.copies.empty()) {
                const int64_t warps = (int64_t)P.copies.size() * tg;
                RPF_LAUNCH(h, PH_STREAM, k_tip_concat, (unsigned)((warps + 7) / 8), 256, 0, d_copies, (int)P.copies.size(), tg,
                           cperm, mmax, (uint32_t)row0, arena_old, arena_new, n);
            }
            // ---- 5. Tips that outgrew minLeaf split in place
            for (size_t gi = 0; gi < P.groups.size(); ++gi) {
                const RtGroup& G = P.groups[gi];
                BottomArgs B{};
                B.ks = n; B.ps = n; B.nn_all = nnrt; B.L = Lk; B.s = G.depth; B.nlb = G.nlb; B.gt0 = 0; B.first_gid = G.first;
                B.given_order = 1;
                B.keys = keys; B.perm = arena_new; B.child = d_rt_child; B.nstart = d_rt_start; B.nsize = d_rt_size;
                B.range = d_rg[gi]; B.lvl_pv = d_pv[gi]; B.thr = rthr; B.mlo = rmlo; B.mhi = rmhi;
                rc = rpf_bottom_launch(h, B, G.count, tg, G.fast, G.max_root);
                if (rc) return rc;
            }
            if (!P.rt_upd_g.empty()) {
                const int64_t tot = (int64_t)P.rt_upd_g.size() * tg;
                RPF_LAUNCH(h, PH_STREAM, k_pool_update, (unsigned)((tot + 255) / 256), 256, 0, d_rt_upd_g, d_rt_upd_u, (int)P.rt_upd_g.size(), tg, nnrt,
                           rthr, rmlo, rmhi, pthr, pmlo, pmhi, 0);
            }
            std::swap(arena_old, arena_new);
        }

        // ---- canonical export of this tree group: BFS topology, [T][nodes] node arrays, perm = final arena
        SP.final_topology(final_tp, pool_of);
        const int64_t nn = final_tp.nnodes();
        if (t0 == 0) {
            h->topo = final_tp;
            int rc = rpf_upload_topology(h);
            if (rc) return rc;
            rc = rpf_alloc_forest(h, nn, n);
            if (rc) return rc;
            h->stream_lost = SP.lost;
        }
        int rc = h->stage_begin((size_t)nn * 4 + 1024);
        if (rc) return rc;
        const int32_t* d_pool_of = h->stage_put(pool_of.data(), pool_of.size());
        rc = h->stage_flush();
        if (rc) return rc;
        RPF_LAUNCH(h, PH_STREAM, k_pool_export, (unsigned)((nn * tg + 255) / 256), 256, 0, d_pool_of, h->d_node_child, nn, tg, t0,
                   pthr, pmlo, pmhi, h->d_thr, h->d_mlo, h->d_mhi);
        for (int t = 0; t < tg; ++t)
            RPF_CUDA(h, cudaMemcpyAsync(h->d_perm + (int64_t)(t0 + t) * n, arena_old + (int64_t)t * n, (size_t)n * 4, cudaMemcpyDeviceToDevice, h->stream));
        RPF_CUDA(h, cudaStreamSynchronize(h->stream));
        if (pthr) cudaFree(pthr);
        h->stream_pool = nullptr;
    }
    return RPF_OK;
}

// host-only: the tree shape `forest`/`tree` produce for n points arriving in chunks (no GPU needed)
extern "C" int64_t rpf_topology_plan_chunked(int64_t n, int32_t maxDepth, int32_t minLeaf, int64_t chunk, int64_t* child, int32_t* depth,
                                             int64_t* seg_start, int64_t* seg_size, int64_t* points_lost) {
    if (n < 0 || maxDepth < 0 || minLeaf < 0 || chunk < 1 || n >= ((int64_t)1 << 31)) return RPF_ERR_ARG;
    Topology tp;
    int64_t lost = 0;
    if (chunk >= n) {
        build_topology(tp, n, maxDepth, minLeaf);
    } else {
        StreamPlanner SP; ChunkPlan P;
        SP.begin(n, maxDepth, minLeaf);
        for (int64_t row0 = 0; row0 < n; row0 += chunk)
            if (!SP.plan_chunk(std::min(chunk, n - row0), P)) return RPF_ERR_UNSUPPORTED;
        std::vector<int32_t> pool_of;
        SP.final_topology(tp, pool_of);
        lost = SP.lost;
    }
    for (int64_t g = 0; g < tp.nnodes(); ++g) {
        if (child) child[g] = tp.child[g];
        if (depth) depth[g] = tp.depth[g];
        if (seg_start) seg_start[g] = tp.start[g];
        if (seg_size) seg_size[g] = tp.size[g];
    }
    if (points_lost) *points_lost = lost;
    return tp.nnodes();
}
