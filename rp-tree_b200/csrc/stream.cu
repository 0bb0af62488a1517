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

struct RtGroup { int depth = 0, first = 0, count = 0, nlb = 0, levels = 0; uint32_t max_root = 0; bool fast = true; };

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
    void reset() {                       // keeps the vectors' storage (a fresh plan per chunk would page-fault it in again)
        ct.n = 0; ct.nlevels = 0; ct.L_eff = 0;
        ct.start.clear(); ct.size.clear(); ct.child.clear(); ct.depth.clear(); ct.level_off.clear(); ct.lvl_maxsize.clear();
        upd_g.clear(); upd_u.clear(); ct_set = false; copies.clear();
        rt_child.clear(); rt_pool.clear(); rt_depth.clear(); rt_start.clear(); rt_size.clear(); rt_upd_g.clear(); rt_upd_u.clear();
        groups.clear(); kept = 0;
    }
};

struct StreamPlanner {
    int64_t n = 0; int maxDepth = 0, minLeaf = 0;
    std::vector<SNode> pool;
    int32_t root = 0;
    int64_t lost = 0;                    // points dropped by the reference's empty-piece rule (Internal.hs:279)
    std::string err;

    // per-node scratch of the chunk being planned (valid where the stamp equals the chunk counter)
    std::vector<uint32_t> pend_src, pend_n;
    std::vector<int32_t> tstamp;
    std::vector<int64_t> newcnt, root_region;
    int32_t cur_chunk = 0;
    size_t last_ct_nodes = 0;
    struct Pend { int32_t u; uint32_t src_new, n_new; };
    std::vector<Pend> tips;              // chunk-tree leaves sitting on Tips: piece location inside the chunk's perm
    std::vector<int32_t> ct_pool;        // pool id of every chunk-tree node
    std::vector<int32_t> resplit;        // Tips that split after this chunk (pool ids)
    struct Fr { int32_t u; bool inside; };
    std::vector<Fr> lay_stack;

    void begin(int64_t n_, int maxd, int minl) {
        n = n_; maxDepth = maxd; minLeaf = minl; pool.clear(); root = 0; lost = 0; err.clear(); cur_chunk = 0; last_ct_nodes = 0;
        pend_src.clear(); pend_n.clear(); tstamp.clear(); newcnt.clear(); root_region.clear();
        new_node(0);
    }
    int32_t new_node(int depth) {
        pool.emplace_back(); pool.back().depth = depth;
        pend_src.push_back(0); pend_n.push_back(0); tstamp.push_back(-1); newcnt.push_back(0); root_region.push_back(-1);
        return (int32_t)pool.size() - 1;
    }
    int64_t subtree_points(int32_t u) const {
        int64_t tot = 0; std::vector<int32_t> st{u};
        while (!st.empty()) { const int32_t v = st.back(); st.pop_back(); if (pool[v].l < 0) tot += pool[v].cnt; else { st.push_back(pool[v].l); st.push_back(pool[v].r); } }
        return tot;
    }

    // Plans the insertion of a chunk of m points.  Returns false on an unsupported shape (err is set).
    bool plan_chunk(int64_t m, ChunkPlan& P) {
        P.reset();
        ++cur_chunk;
        Topology& ct = P.ct;
        if (last_ct_nodes) {
            const size_t r = last_ct_nodes + last_ct_nodes / 8 + 64;
            ct.start.reserve(r); ct.size.reserve(r); ct.child.reserve(r); ct.depth.reserve(r);
            P.upd_g.reserve(r / 2 + 8); P.upd_u.reserve(r / 2 + 8); P.copies.reserve(r / 2 + 8);
        }
        ct.n = m; ct.maxDepth = maxDepth; ct.minLeaf = minLeaf;
        ct_pool.clear(); tips.clear(); resplit.clear();
        const bool fresh = pool[root].l < 0 && pool[root].cnt == 0 && m > minLeaf && maxDepth > 0;
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
        last_ct_nodes = (size_t)ct.nnodes();
        for (const Pend& t : tips) { pend_src[t.u] = t.src_new; pend_n[t.u] = t.n_new; tstamp[t.u] = cur_chunk; newcnt[t.u] = pool[t.u].cnt + t.n_new; }
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
        }

        // ---- new left-to-right layout of the leaf arena + copy descriptors
        {
            int64_t cur = 0;
            std::vector<Fr>& st = lay_stack;
            st.clear(); st.push_back(Fr{root, false});
            while (!st.empty()) {
                const Fr f = st.back(); st.pop_back();
                SNode& N = pool[f.u];
                bool inside = f.inside;
                if (!inside && tstamp[f.u] == cur_chunk) {
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

// new content of every Tip: the chunk's piece (ids local to the batch, + row0) followed by the old content
// (Internal.hs:288 `xs <> xs0`).  The descriptors are sorted by destination (the arena is laid out left to right), so
// a thread owns one arena slot: it finds the slot's Tip by binary search once and moves that slot for every tree --
// consecutive threads touch consecutive words of both arenas.
__global__ void __launch_bounds__(256) k_tip_concat(const TipCopy* __restrict__ cp, int ncopy, int tg, uint32_t kept,
                                                     const uint32_t* __restrict__ piece, int64_t piece_stride, uint32_t row0,
                                                     const uint32_t* __restrict__ old_arena, uint32_t* __restrict__ new_arena, int64_t astride) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= kept) return;
    int lo = 0, hi = ncopy - 1;                   // last descriptor with dst <= o
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (__ldg(&cp[mid].dst) <= o) lo = mid; else hi = mid - 1; }
    const TipCopy c = cp[lo];
    const uint32_t j = o - c.dst;
    if (j >= c.n_new + c.n_old) return;           // (cannot happen: the descriptors tile [0, kept))
    // blockIdx.y owns 8 trees; the 8 moves are independent (all loads issued before the stores)
    const int tA = blockIdx.y * 8, tB = min(tg, tA + 8);
    const uint32_t* src; int64_t sstride; uint32_t add;
    if (j < c.n_new) { src = piece + c.src_new + j; sstride = piece_stride; add = row0; }
    else { src = old_arena + c.src_old + (j - c.n_new); sstride = astride; add = 0; }
    uint32_t v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) if (tA + q < tB) v[q] = __ldg(src + (int64_t)(tA + q) * sstride);
#pragma unroll
    for (int q = 0; q < 8; ++q) if (tA + q < tB) new_arena[(int64_t)(tA + q) * astride + o] = v[q] + add;
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
// the plan of a whole streaming build (pure function of the shape: cached in the handle across builds)
// =====================================================================================================
namespace {

struct RtGroupL { RtGroup g; size_t off_rg = (size_t)-1, off_pv = (size_t)-1; };
struct StreamChunk {
    int batch = 0; int64_t row0 = 0;
    int n_upd = 0; size_t off_upd_g = 0, off_upd_u = 0; bool ct_set = false;
    int n_copies = 0; size_t off_copies = 0; int64_t kept = 0;
    int nnrt = 0; size_t off_rt_start = 0, off_rt_size = 0, off_rt_child = 0;
    int n_rt_upd = 0; size_t off_rt_upd_g = 0, off_rt_upd_u = 0;
    std::vector<RtGroupL> groups;
};
struct StreamBatch {
    int64_t row0 = 0, rows = 0, nn = 0;
    int first_chunk = 0, nchunks = 0;
    JobPlan JP;                                  // the descent of all the batch's chunks, one level-synchronous job
    size_t off_start = 0, off_size = 0, off_child = 0;
};
struct StreamPlanAll {
    int64_t n = 0, chunk = 0; int maxDepth = 0, minLeaf = 0, cap = 0, Lk = 0; bool force_generic = false;
    TableBuf TB;
    char* d_tab = nullptr;
    std::vector<StreamBatch> batches;
    std::vector<StreamChunk> chunks;
    Topology final_tp;
    size_t off_pool_of = 0;
    int64_t lost = 0, max_nn = 0, max_nnrt = 0, max_rows = 0;
    size_t pool_nodes = 0;
    bool order_exact = true;
    ~StreamPlanAll() { if (d_tab) cudaFree(d_tab); }
};
void free_stream_plan(void* p) { delete (StreamPlanAll*)p; }

// levels / kernel choice of one group of re-split roots (same depth); rg / pv: tables of the generic bottom kernel
void size_group(const ChunkPlan& P, RtGroup& G, bool force_generic, std::vector<int2>& rg, std::vector<uint32_t>& pv) {
    int maxrel = 0;                   // depth of the deepest descendant below this group's roots
    std::vector<int32_t> fr, nx;
    pv.assign((size_t)G.depth, 1);
    for (int e = 0; e < G.count; ++e) fr.push_back(G.first + e);
    for (int rl = 0; !fr.empty(); ++rl) {
        uint32_t mx = 1; nx.clear();
        for (int32_t g : fr) { mx = std::max(mx, P.rt_size[g]); if (P.rt_child[g] >= 0) { nx.push_back(P.rt_child[g]); nx.push_back(P.rt_child[g] + 1); } }
        uint32_t p2 = 1; while (p2 < mx) p2 <<= 1;
        pv.push_back(p2);              // next_pow2(max size) at absolute level G.depth + rl
        if (!nx.empty()) maxrel = rl + 1;
        fr.swap(nx);
    }
    G.nlb = std::max(1, maxrel + 1);
    G.levels = maxrel;
    G.fast = maxrel <= rpf_bottom_fast_levels() && !force_generic;
    if (G.fast) {
        unsigned slots = 256; while (slots < G.max_root) slots <<= 1;
        while (maxrel > 0 && (slots >> maxrel) == 0) slots <<= 1;
        if (slots > 8192) G.fast = false;
    }
    rg.clear();
    if (!G.fast) make_ranges(P.rt_child, G.first, G.count, G.nlb, rg);
}

// Plans every chunk (sequentially: the shape after chunk c depends on chunks < c) and groups the chunks into batches
// whose descents run as ONE job: the routing of a chunk depends only on the tree SHAPE the chunk finds, which the
// planner knows, never on the thresholds or leaf contents earlier chunks produced.
bool build_stream_plan(StreamPlanAll& S, std::string& err) {
    StreamPlanner SP;
    SP.begin(S.n, S.maxDepth, S.minLeaf);
    ChunkPlan P;
    const int64_t node_cap = 6000000;            // temp node arrays of a batch: tg * nodes * 24 bytes
    const int64_t top_per_chunk = 4 * ((S.chunk + S.cap - 1) / S.cap) + 2;      // top-phase nodes a chunk can contribute
    const int64_t chunk_cap = S.chunk <= S.cap ? ((int64_t)1 << 40) : std::max<int64_t>(1, 60000 / top_per_chunk);
    int64_t row = 0;
    std::vector<Topology> cts;
    std::vector<std::vector<int32_t>> upd_gs;
    while (row < S.n) {
        StreamBatch B;
        B.row0 = row; B.first_chunk = (int)S.chunks.size();
        size_t ncts = 0;
        int64_t nodes = 0;
        do {
            const int64_t m = std::min(S.chunk, S.n - row);
            if (!SP.plan_chunk(m, P)) { err = SP.err; return false; }
            StreamChunk C;
            C.batch = (int)S.batches.size(); C.row0 = row; C.ct_set = P.ct_set;
            const uint32_t rel = (uint32_t)(row - B.row0);
            for (TipCopy& c : P.copies) c.src_new += rel;
            C.kept = P.kept;
            C.n_copies = (int)P.copies.size(); C.off_copies = S.TB.put(P.copies.data(), P.copies.size() * sizeof(TipCopy));
            C.n_upd = (int)P.upd_g.size(); C.off_upd_u = S.TB.put(P.upd_u.data(), P.upd_u.size() * 4);
            C.nnrt = (int)P.rt_size.size();
            C.off_rt_start = S.TB.put(P.rt_start.data(), P.rt_start.size() * 4);
            C.off_rt_size = S.TB.put(P.rt_size.data(), P.rt_size.size() * 4);
            C.off_rt_child = S.TB.put(P.rt_child.data(), P.rt_child.size() * 4);
            C.n_rt_upd = (int)P.rt_upd_g.size();
            C.off_rt_upd_g = S.TB.put(P.rt_upd_g.data(), P.rt_upd_g.size() * 4);
            C.off_rt_upd_u = S.TB.put(P.rt_upd_u.data(), P.rt_upd_u.size() * 4);
            S.max_nnrt = std::max<int64_t>(S.max_nnrt, C.nnrt);
            for (RtGroup& G : P.groups) {
                RtGroupL GL;
                std::vector<int2> rg; std::vector<uint32_t> pv;
                size_group(P, G, S.force_generic, rg, pv);
                GL.g = G;
                if (!G.fast) {
                    GL.off_rg = S.TB.put(rg.data(), rg.size() * sizeof(int2));
                    GL.off_pv = S.TB.put(pv.data(), pv.size() * 4);
                }
                C.groups.push_back(GL);
            }
            S.chunks.push_back(std::move(C));
            nodes += P.ct.nnodes();
            if (ncts == cts.size()) { cts.emplace_back(); upd_gs.emplace_back(); }
            std::swap(cts[ncts], P.ct);            // P gets an old chunk tree's storage back
            upd_gs[ncts].assign(P.upd_g.begin(), P.upd_g.end());
            ++ncts;
            row += m;
        } while (row < S.n && nodes < node_cap && (int64_t)ncts < chunk_cap);
        B.rows = row - B.row0; B.nchunks = (int)ncts;

        // ---- super topology of the batch: level l = the level-l nodes of every chunk tree, chunk-major
        Topology sup;
        sup.n = B.rows; sup.maxDepth = S.maxDepth; sup.minLeaf = S.minLeaf;
        int nl = 0;
        for (size_t j = 0; j < ncts; ++j) nl = std::max(nl, cts[j].nlevels);
        std::vector<std::vector<int64_t>> base(ncts, std::vector<int64_t>(nl + 1, 0));
        sup.level_off.assign(nl + 1, 0);
        {
            int64_t tot = 0;
            for (int l = 0; l < nl; ++l) {
                sup.level_off[l] = tot;
                for (size_t j = 0; j < ncts; ++j) {
                    base[j][l] = tot;
                    if (l < cts[j].nlevels) tot += cts[j].level_off[l + 1] - cts[j].level_off[l];
                }
            }
            sup.level_off[nl] = tot;
            sup.start.resize(tot); sup.size.resize(tot); sup.child.resize(tot); sup.depth.resize(tot);
        }
        sup.nlevels = nl; sup.lvl_maxsize.assign(nl, 0); sup.L_eff = 0;
        for (size_t j = 0; j < ncts; ++j) {
            const Topology& ct = cts[j];
            const uint32_t rel = (uint32_t)(S.chunks[B.first_chunk + j].row0 - B.row0);
            sup.L_eff = std::max(sup.L_eff, ct.L_eff);
            for (int l = 0; l < ct.nlevels; ++l) {
                const int64_t lo = ct.level_off[l], hi = ct.level_off[l + 1];
                for (int64_t g = lo; g < hi; ++g) {
                    const int64_t sg = base[j][l] + (g - lo);
                    sup.start[sg] = rel + ct.start[g]; sup.size[sg] = ct.size[g]; sup.depth[sg] = l;
                    sup.child[sg] = ct.child[g] < 0 ? -1 : (int32_t)(base[j][l + 1] + (ct.child[g] - ct.level_off[l + 1]));
                    sup.lvl_maxsize[l] = std::max(sup.lvl_maxsize[l], ct.size[g]);
                }
            }
            // pool updates of this chunk: chunk-tree id -> id in the batch job
            std::vector<int32_t>& ug = upd_gs[j];
            for (int32_t& g : ug) { const int l = ct.depth[g]; g = (int32_t)(base[j][l] + (g - ct.level_off[l])); }
            S.chunks[B.first_chunk + j].off_upd_g = S.TB.put(ug.data(), ug.size() * 4);
        }
        B.nn = sup.nnodes();
        B.off_start = S.TB.put(sup.start.data(), sup.start.size() * 4);
        B.off_size = S.TB.put(sup.size.data(), sup.size.size() * 4);
        B.off_child = S.TB.put(sup.child.data(), sup.child.size() * 4);
        TableBuf jt;
        rpf_plan_job(sup, S.cap, S.Lk, S.force_generic, jt, B.JP);
        {   // re-base the job's table offsets into the plan's block
            const size_t o = S.TB.put(jt.bytes.data(), jt.bytes.size());
            if (B.JP.off_range != (size_t)-1) B.JP.off_range += o;
            if (B.JP.off_lvlpv != (size_t)-1) B.JP.off_lvlpv += o;
            B.JP.off_nb += o;
        }
        if (B.JP.G.s_top > 0 && B.JP.G.NTOP > 65535) { err = "internal: batch exceeds 16-bit labels"; return false; }
        if (!B.JP.G.order_exact) S.order_exact = false;
        S.max_nn = std::max(S.max_nn, B.nn); S.max_rows = std::max(S.max_rows, B.rows);
        S.batches.push_back(std::move(B));
    }
    std::vector<int32_t> pool_of;
    SP.final_topology(S.final_tp, pool_of);
    S.off_pool_of = S.TB.put(pool_of.data(), pool_of.size() * 4);
    S.lost = SP.lost; S.pool_nodes = SP.pool.size();
    return true;
}

}  // namespace

#define WSX(h, var, type, slot, bytes)                                  \
    type* var = (type*)(h)->ws_get((slot), (bytes));                    \
    if (!var) return RPF_ERR_NOMEM;

int rpf_build_stream_impl(rpf_handle* h, int maxDepth, int minLeaf, int64_t chunk) {
    const int64_t n = h->n;
    const int T = h->T;
    const int Lk = std::max(maxDepth, 1);

    // ---- the plan: cached per shape (like the batch topology, it does not depend on the data)
    StreamPlanAll* S = (StreamPlanAll*)h->stream_plan;
    if (!S || S->n != n || S->chunk != chunk || S->maxDepth != maxDepth || S->minLeaf != minLeaf || S->cap != h->bottom_cap ||
        S->force_generic != h->force_generic_bottom) {
        if (S) { cudaStreamSynchronize(h->stream); delete S; h->stream_plan = nullptr; }
        S = new StreamPlanAll();
        S->n = n; S->chunk = chunk; S->maxDepth = maxDepth; S->minLeaf = minLeaf; S->cap = h->bottom_cap; S->Lk = Lk;
        S->force_generic = h->force_generic_bottom;
        std::string err;
        if (!build_stream_plan(*S, err)) { delete S; return rpf_fail(h, RPF_ERR_UNSUPPORTED, err); }
        if (cudaMalloc(&S->d_tab, std::max<size_t>(S->TB.bytes.size(), 256)) != cudaSuccess) { cudaGetLastError(); delete S; return rpf_fail(h, RPF_ERR_NOMEM, "stream plan: device tables"); }
        cudaError_t e = cudaMemcpyAsync(S->d_tab, S->TB.bytes.data(), S->TB.bytes.size(), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { delete S; return rpf_fail(h, RPF_ERR_CUDA, std::string("stream plan upload: ") + cudaGetErrorString(e)); }
        S->TB.bytes.clear(); S->TB.bytes.shrink_to_fit();      // the tables live on the device from here on
        h->stream_plan = S; h->stream_plan_free = free_stream_plan;
    }
    h->leaf_order_exact = S->order_exact;
    h->stream_lost = S->lost;
    const char* tab = S->d_tab;

    // ---- tree group size: persistent key store [Tg][Lk][n] + two leaf arenas + the batch jobs' workspace
    size_t freeB = 0, totalB = 0;
    RPF_CUDA(h, cudaMemGetInfo(&freeB, &totalB));
    size_t job_ws = 0;
    for (const StreamBatch& B : S->batches) job_ws = std::max(job_ws, rpf_job_ws_per_tree(B.JP.G, B.rows));
    const size_t per_tree = (size_t)Lk * n * 8 + (size_t)n * 8 + (size_t)S->max_rows * 4 + job_ws +
                            (size_t)(S->max_nn + S->max_nnrt + 2 + (int64_t)S->pool_nodes) * 24 + ((size_t)1 << 16);
    const size_t budget = (size_t)((double)(freeB + h->ws_bytes) * 0.7);
    if (per_tree > budget) return rpf_fail(h, RPF_ERR_NOMEM, "not enough device memory for one tree's keys");
    const int Tg = (int)std::min<size_t>((size_t)T, std::max<size_t>(1, budget / per_tree));

    WSX(h, keys, ull, WS_KEYS, (size_t)Tg * Lk * n * 8);
    WSX(h, kmin, ull, WS_KMIN, (size_t)Tg * Lk * 8);
    WSX(h, kmax, ull, WS_KMAX, (size_t)Tg * Lk * 8);
    WSX(h, arena0, uint32_t, WS_S_ARENA0, (size_t)Tg * n * 4);
    WSX(h, arena1, uint32_t, WS_S_ARENA1, (size_t)Tg * n * 4);
    WSX(h, cperm, uint32_t, WS_S_CPERM, (size_t)Tg * S->max_rows * 4);
    WSX(h, tmpn, double, WS_S_TMPN, (size_t)Tg * (S->max_nn + S->max_nnrt + 2) * 8 * 3);
    WSX(h, pool, double, WS_S_POOL, (size_t)Tg * (S->pool_nodes + 1) * 8 * 3);

    // ---- canonical result shape
    h->topo = S->final_tp; h->topo_key_n = -1;
    int rc = rpf_upload_topology(h);
    if (rc) return rc;
    const int64_t nn = h->topo.nnodes();
    rc = rpf_alloc_forest(h, nn, n);
    if (rc) return rc;

    for (int t0 = 0; t0 < T; t0 += Tg) {
        const int tg = std::min(Tg, T - t0);
        uint32_t* arena_old = arena0; uint32_t* arena_new = arena1;
        RPF_CUDA(h, cudaMemsetAsync(arena0, 0xff, (size_t)tg * n * 4, h->stream));     // slots past the kept points stay 0xffffffff
        RPF_CUDA(h, cudaMemsetAsync(arena1, 0xff, (size_t)tg * n * 4, h->stream));
        double* pthr = pool; double* pmlo = pthr + (size_t)tg * S->pool_nodes; double* pmhi = pmlo + (size_t)tg * S->pool_nodes;
        for (const StreamBatch& B : S->batches) {
            const int64_t nnb = B.nn;
            double* cthr = tmpn; double* cmlo = cthr + (size_t)tg * nnb; double* cmhi = cmlo + (size_t)tg * nnb;
            double* rthr = cmhi + (size_t)tg * nnb; double* rmlo = rthr + (size_t)tg * S->max_nnrt; double* rmhi = rmlo + (size_t)tg * S->max_nnrt;
            // ---- 1. keys of the batch's points
            RPF_CUDA(h, cudaMemsetAsync(kmin, 0xff, (size_t)tg * Lk * 8, h->stream));
            RPF_CUDA(h, cudaMemsetAsync(kmax, 0x00, (size_t)tg * Lk * 8, h->stream));
            if (maxDepth > 0) {
                rc = rpf_project_launch(h, PH_PROJECT, h->dX + B.row0 * (int64_t)h->d, B.rows, t0, tg, Lk, true, keys + B.row0, n, kmin, kmax);
                if (rc) return rc;
            }
            // ---- 2. every chunk of the batch descends through the Bins of the tree it finds (one job)
            BuildJob J{};
            J.d_start = (const uint32_t*)(tab + B.off_start); J.d_size = (const uint32_t*)(tab + B.off_size); J.d_child = (const int32_t*)(tab + B.off_child);
            J.n = B.rows; J.ks = n; J.ps = S->max_rows; J.ns = nnb; J.Lk = Lk;
            J.keys = keys + B.row0; J.kmin = kmin; J.kmax = kmax;
            J.perm = cperm; J.thr = cthr; J.mlo = cmlo; J.mhi = cmhi; J.gt0 = 0; J.tg = tg;
            rc = rpf_launch_job(h, J, B.JP, tab);
            if (rc) return rc;
            // ---- 3..5 per chunk, in arrival order
            for (int ci = B.first_chunk; ci < B.first_chunk + B.nchunks; ++ci) {
                const StreamChunk& C = S->chunks[ci];
                if (C.n_upd) {        // thr' = (thr0 + thr) / 2, margin' = margin0 <> margin
                    const int64_t tot = (int64_t)C.n_upd * tg;
                    RPF_LAUNCH(h, PH_STREAM, k_pool_update, (unsigned)((tot + 255) / 256), 256, 0, (const int32_t*)(tab + C.off_upd_g),
                               (const int32_t*)(tab + C.off_upd_u), C.n_upd, tg, nnb, cthr, cmlo, cmhi, pthr, pmlo, pmhi, C.ct_set ? 0 : 1);
                }
                if (C.n_copies && C.kept) {     // Tip contents: piece ++ old
                    RPF_LAUNCH(h, PH_STREAM_CONCAT, k_tip_concat, dim3((unsigned)((C.kept + 255) / 256), (unsigned)((tg + 7) / 8)), 256, 0, (const TipCopy*)(tab + C.off_copies), C.n_copies, tg,
                               (uint32_t)C.kept, cperm, S->max_rows, (uint32_t)B.row0, arena_old, arena_new, n);
                }
                for (const RtGroupL& GL : C.groups) {      // Tips that outgrew minLeaf split in place
                    const RtGroup& G = GL.g;
                    BottomArgs A{};
                    A.ks = n; A.ps = n; A.nn_all = S->max_nnrt; A.L = Lk; A.s = G.depth; A.nlb = G.nlb; A.gt0 = 0; A.first_gid = G.first;
                    A.given_order = 1;
                    A.keys = keys; A.perm = arena_new; A.child = (const int32_t*)(tab + C.off_rt_child);
                    A.nstart = (const uint32_t*)(tab + C.off_rt_start); A.nsize = (const uint32_t*)(tab + C.off_rt_size);
                    A.range = GL.off_rg == (size_t)-1 ? nullptr : (const int2*)(tab + GL.off_rg);
                    A.lvl_pv = GL.off_pv == (size_t)-1 ? nullptr : (const uint32_t*)(tab + GL.off_pv);
                    A.thr = rthr; A.mlo = rmlo; A.mhi = rmhi;
                    A.kmin = kmin; A.kmax = kmax;                 // range of the current batch's keys (a sample is enough)
                    rc = rpf_bottom_launch(h, A, G.count, tg, G.fast, G.max_root, G.levels);
                    if (rc) return rc;
                }
                if (C.n_rt_upd) {
                    const int64_t tot = (int64_t)C.n_rt_upd * tg;
                    RPF_LAUNCH(h, PH_STREAM, k_pool_update, (unsigned)((tot + 255) / 256), 256, 0, (const int32_t*)(tab + C.off_rt_upd_g),
                               (const int32_t*)(tab + C.off_rt_upd_u), C.n_rt_upd, tg, S->max_nnrt, rthr, rmlo, rmhi, pthr, pmlo, pmhi, 0);
                }
                std::swap(arena_old, arena_new);
            }
        }
        // ---- canonical export of this tree group: [T][nodes] node arrays, perm = final arena
        RPF_LAUNCH(h, PH_STREAM, k_pool_export, (unsigned)((nn * tg + 255) / 256), 256, 0, (const int32_t*)(tab + S->off_pool_of), h->d_node_child, nn, tg, t0,
                   pthr, pmlo, pmhi, h->d_thr, h->d_mlo, h->d_mhi);
        for (int t = 0; t < tg; ++t)
            RPF_CUDA(h, cudaMemcpyAsync(h->d_perm + (int64_t)(t0 + t) * n, arena_old + (int64_t)t * n, (size_t)n * 4, cudaMemcpyDeviceToDevice, h->stream));
    }
    return RPF_OK;
}

// =====================================================================================================
// true incremental insert: the forest is a fold over the chunks AS THEY ARRIVE
//   forest src = src .| chunksOf n .| foldl (insertMulti ...) im0        Conduit.hs:157-176
//   insertMulti maxd minl rvss tts xs = per tree: insert ... tt xs       Internal.hs:243-255
// rpf_insert_begin makes every tree `Tip () mempty`; every rpf_insert_chunk is ONE insertMulti: the chunk is uploaded,
// projected, descends through the Bins it finds (thr' = (thr0 + thr)/2, margin' = margin0 <> margin), is prepended to the
// Tips it reaches, and Tips that outgrow minLeaf split -- the per-chunk half of rpf_build_stream_impl, with the planner
// (sizes only) kept alive between calls instead of replaying a known n.  After every call the handle holds a complete,
// queryable forest over all points inserted so far.  Chunks may have any sizes; neither n nor the chunk count is known
// in advance (per-point arrays grow geometrically).
// =====================================================================================================
namespace {
struct InsertSession {
    StreamPlanner SP; ChunkPlan P;
    int maxDepth = 0, minLeaf = 0, Lk = 1, d = 0, T = 0;
    int64_t n = 0, cap = 0;
    double* X = nullptr;                 // [cap][d]
    ull* keys = nullptr;                 // [T][Lk][cap]
    uint32_t* arena[2] = {nullptr, nullptr};   // [T][cap]: leaf contents, left to right (slots past the kept points: 0xffffffff)
    int cur = 0; bool order_exact = true;
    double* pool = nullptr; size_t pool_cap = 0;     // thr | mlo | mhi, each [pool_cap][T] (node-major)
    ~InsertSession() {
        if (X) cudaFree(X);
        if (keys) cudaFree(keys);
        if (arena[0]) cudaFree(arena[0]);
        if (arena[1]) cudaFree(arena[1]);
        if (pool) cudaFree(pool);
    }
};
void free_insert_session(void* p) { delete (InsertSession*)p; }
}  // namespace

void rpf_insert_drop(rpf_handle* h) {
    if (!h->insert_session) return;
    cudaStreamSynchronize(h->stream);
    InsertSession* S = (InsertSession*)h->insert_session;
    if (h->dX == S->X) { h->dX = nullptr; h->ownX = false; h->n = 0; h->x_bytes = 0; h->built = false; }
    if (h->insert_session_free) h->insert_session_free(h->insert_session);
    h->insert_session = nullptr;
}

// closes the session but keeps what it built: the points become the handle's own (as after rpf_set_points) and the forest
// stays queryable; the key store, the arenas and the node pool are released
int rpf_insert_end_impl(rpf_handle* h) {
    InsertSession* S = (InsertSession*)h->insert_session;
    if (!S) return RPF_OK;
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    if (S->X) { h->dX = S->X; h->ownX = true; h->x_bytes = std::max<size_t>((size_t)S->cap * S->d * 8, 16); h->n = S->n; h->x_pad_rows = 0; S->X = nullptr; }
    else { h->dX = nullptr; h->ownX = false; h->x_bytes = 0; h->n = 0; }
    if (h->insert_session_free) h->insert_session_free(h->insert_session);
    h->insert_session = nullptr;
    ++h->cfg_epoch;
    return RPF_OK;
}

int rpf_insert_begin_impl(rpf_handle* h, int d, int maxDepth, int minLeaf) {
    rpf_insert_drop(h);
    if (h->d_xlast) { cudaFree(h->d_xlast); h->d_xlast = nullptr; }
    if (h->ownX && h->dX) cudaFree((void*)h->dX);
    h->dX = nullptr; h->ownX = false; h->x_bytes = 0; h->n = 0; h->d = d; h->x_pad_rows = 0;
    h->built = false; h->sink_pending = false;
    ++h->cfg_epoch; ++h->x_version;
    InsertSession* S = new InsertSession();
    S->maxDepth = maxDepth; S->minLeaf = minLeaf; S->Lk = std::max(maxDepth, 1); S->d = d; S->T = h->T;
    S->SP.begin(0, maxDepth, minLeaf);
    h->insert_session = S; h->insert_session_free = free_insert_session;
    // the empty forest: one empty Tip per tree
    h->stream_lost = 0;
    S->SP.n = 0;
    std::vector<int32_t> pool_of;
    S->SP.final_topology(h->topo, pool_of);
    h->topo_key_n = -1;
    int rc = rpf_upload_topology(h);
    if (!rc) rc = rpf_alloc_forest(h, h->topo.nnodes(), 0);
    if (rc) return rc;
    h->built = true;
    return RPF_OK;
}

// grows a [rows][cap] device array to [rows][ncap], keeping the first `used` elements of every row
template <typename E>
static int grow_rows(rpf_handle* h, E** buf, size_t rows, int64_t cap, int64_t ncap, int64_t used, int fill) {
    E* nb = nullptr;
    if (cudaMalloc(&nb, std::max<size_t>(rows * (size_t)ncap * sizeof(E), 16)) != cudaSuccess) { cudaGetLastError(); return rpf_fail(h, RPF_ERR_NOMEM, "insert: out of device memory"); }
    if (fill >= 0) RPF_CUDA(h, cudaMemsetAsync(nb, fill, rows * (size_t)ncap * sizeof(E), h->stream));
    if (*buf && used > 0)
        RPF_CUDA(h, cudaMemcpy2DAsync(nb, (size_t)ncap * sizeof(E), *buf, (size_t)cap * sizeof(E), (size_t)used * sizeof(E), rows, cudaMemcpyDeviceToDevice, h->stream));
    if (*buf) { RPF_CUDA(h, cudaStreamSynchronize(h->stream)); cudaFree(*buf); }
    *buf = nb;
    return RPF_OK;
}

int rpf_insert_chunk_impl(rpf_handle* h, const double* Xc, int64_t m) {
    InsertSession* S = (InsertSession*)h->insert_session;
    const int T = S->T, Lk = S->Lk, d = S->d;
    const int64_t n0 = S->n, n1 = n0 + m;
    if (n1 >= ((int64_t)1 << 31)) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "insert: more than 2^31 - 1 points");
    // ---- plan on sizes alone (the planner's state is the shape of the tree so far)
    ChunkPlan& P = S->P;
    S->SP.n = n1;
    if (!S->SP.plan_chunk(m, P)) {
        const std::string e = S->SP.err;
        rpf_insert_drop(h);                      // the planner's state is no longer the device's
        return rpf_fail(h, RPF_ERR_UNSUPPORTED, e + " (the insert session was closed)");
    }
    ++h->cfg_epoch; ++h->x_version;
    h->built = false;
    // ---- capacity of the per-point arrays
    if (n1 > S->cap) {
        const int64_t ncap = ((std::max<int64_t>({n1, 2 * S->cap, 65536}) + 255) / 256) * 256;
        int rc = grow_rows<double>(h, &S->X, 1, S->cap * d, ncap * d, n0 * d, -1);
        if (!rc) rc = grow_rows<ull>(h, &S->keys, (size_t)T * Lk, S->cap, ncap, n0, -1);
        if (!rc) rc = grow_rows<uint32_t>(h, &S->arena[0], (size_t)T, S->cap, ncap, n0, 0xff);
        if (!rc) rc = grow_rows<uint32_t>(h, &S->arena[1], (size_t)T, S->cap, ncap, n0, 0xff);
        if (rc) return rc;
        S->cap = ncap;
    }
    const int64_t cap = S->cap;
    if (S->SP.pool.size() + 1 > S->pool_cap) {
        const size_t npc = std::max<size_t>(2 * S->pool_cap, S->SP.pool.size() + 1024);
        double* np = nullptr;
        if (cudaMalloc(&np, npc * T * 24) != cudaSuccess) { cudaGetLastError(); return rpf_fail(h, RPF_ERR_NOMEM, "insert: out of device memory"); }
        for (int a = 0; a < 3 && S->pool; ++a)
            RPF_CUDA(h, cudaMemcpyAsync(np + (size_t)a * npc * T, S->pool + (size_t)a * S->pool_cap * T, S->pool_cap * T * 8, cudaMemcpyDeviceToDevice, h->stream));
        if (S->pool) { RPF_CUDA(h, cudaStreamSynchronize(h->stream)); cudaFree(S->pool); }
        S->pool = np; S->pool_cap = npc;
    }
    double* pthr = S->pool; double* pmlo = pthr + S->pool_cap * T; double* pmhi = pmlo + S->pool_cap * T;
    h->dX = S->X; h->ownX = false; h->n = n0; h->d = d;
    if (m > 0) RPF_CUDA(h, cudaMemcpyAsync(S->X + n0 * d, Xc, (size_t)m * d * 8, cudaMemcpyHostToDevice, h->stream));

    // ---- everything the device needs to know about this chunk, in ONE slot of the staging ring (a second stage_begin
    //      inside this call could grow the ring and move the first slot's device block)
    const int64_t nnct = P.ct.nnodes(), nnrt = (int64_t)P.rt_size.size();
    std::vector<int32_t> pool_of;
    S->SP.final_topology(h->topo, pool_of);       // canonical shape of the forest AFTER this chunk (sizes only)
    h->topo.n = n1; h->topo_key_n = -1;
    TableBuf JT; JobPlan JP;
    if (m > 0) rpf_plan_job(P.ct, h->bottom_cap, Lk, h->force_generic_bottom, JT, JP);
    std::vector<std::vector<int2>> rgs(P.groups.size()); std::vector<std::vector<uint32_t>> pvs(P.groups.size());
    size_t tab_bytes = (size_t)nnct * 12 + P.upd_g.size() * 8 + P.copies.size() * sizeof(TipCopy) + (size_t)nnrt * 12 + P.rt_upd_g.size() * 8 +
                       pool_of.size() * 4 + JT.bytes.size() + 64 * 256;
    for (size_t i = 0; i < P.groups.size(); ++i) {
        size_group(P, P.groups[i], h->force_generic_bottom, rgs[i], pvs[i]);
        tab_bytes += rgs[i].size() * sizeof(int2) + pvs[i].size() * 4 + 512;
    }
    int rc = h->stage_begin(tab_bytes);
    if (rc) return rc;
    const uint32_t* d_start = h->stage_put(P.ct.start.data(), P.ct.start.size());
    const uint32_t* d_size = h->stage_put(P.ct.size.data(), P.ct.size.size());
    const int32_t* d_child = h->stage_put(P.ct.child.data(), P.ct.child.size());
    const int32_t* d_upd_g = h->stage_put(P.upd_g.data(), P.upd_g.size());
    const int32_t* d_upd_u = h->stage_put(P.upd_u.data(), P.upd_u.size());
    const TipCopy* d_copies = h->stage_put(P.copies.data(), P.copies.size());
    const uint32_t* d_rt_start = h->stage_put(P.rt_start.data(), P.rt_start.size());
    const uint32_t* d_rt_size = h->stage_put(P.rt_size.data(), P.rt_size.size());
    const int32_t* d_rt_child = h->stage_put(P.rt_child.data(), P.rt_child.size());
    const int32_t* d_rt_upd_g = h->stage_put(P.rt_upd_g.data(), P.rt_upd_g.size());
    const int32_t* d_rt_upd_u = h->stage_put(P.rt_upd_u.data(), P.rt_upd_u.size());
    const int32_t* d_pool_of = h->stage_put(pool_of.data(), pool_of.size());
    const char* d_jtab = (const char*)h->stage_put_raw(JT.bytes.data(), JT.bytes.size());
    std::vector<const int2*> d_rg(P.groups.size(), nullptr); std::vector<const uint32_t*> d_pv(P.groups.size(), nullptr);
    bool ok = d_start && d_size && d_child && d_upd_g && d_upd_u && d_copies && d_rt_start && d_rt_size && d_rt_child && d_rt_upd_g && d_rt_upd_u &&
              d_pool_of && d_jtab;
    for (size_t i = 0; i < P.groups.size(); ++i)
        if (!P.groups[i].fast) {
            d_rg[i] = h->stage_put(rgs[i].data(), rgs[i].size()); d_pv[i] = h->stage_put(pvs[i].data(), pvs[i].size());
            ok = ok && d_rg[i] && d_pv[i];
        }
    if (!ok) return rpf_fail(h, RPF_ERR_NOMEM, h->err);
    rc = h->stage_flush();
    if (rc) return rc;

    WSX(h, kmin, ull, WS_KMIN, (size_t)T * Lk * 8);
    WSX(h, kmax, ull, WS_KMAX, (size_t)T * Lk * 8);
    WSX(h, cperm, uint32_t, WS_S_CPERM, (size_t)T * std::max<int64_t>(m, 1) * 4);
    WSX(h, tmpn, double, WS_S_TMPN, (size_t)T * (nnct + nnrt + 2) * 8 * 3);
    double* cthr = tmpn; double* cmlo = cthr + (size_t)T * nnct; double* cmhi = cmlo + (size_t)T * nnct;
    double* rthr = cmhi + (size_t)T * nnct; double* rmlo = rthr + (size_t)T * nnrt; double* rmhi = rmlo + (size_t)T * nnrt;
    uint32_t* arena_old = S->arena[S->cur]; uint32_t* arena_new = S->arena[S->cur ^ 1];
    if (m > 0) {
        // ---- 1. keys of the chunk's points, appended to the persistent key store
        RPF_CUDA(h, cudaMemsetAsync(kmin, 0xff, (size_t)T * Lk * 8, h->stream));
        RPF_CUDA(h, cudaMemsetAsync(kmax, 0x00, (size_t)T * Lk * 8, h->stream));
        if (S->maxDepth > 0) {
            rc = rpf_project_launch(h, PH_PROJECT, S->X + n0 * d, m, 0, T, Lk, true, S->keys + n0, cap, kmin, kmax);
            if (rc) return rc;
        }
        // ---- 2. the chunk descends through the Bins of the tree it finds
        BuildJob J{};
        J.tp = &P.ct; J.d_start = d_start; J.d_size = d_size; J.d_child = d_child;
        J.n = m; J.ks = cap; J.ps = m; J.ns = nnct; J.Lk = Lk;
        J.keys = S->keys + n0; J.kmin = kmin; J.kmax = kmax;
        J.perm = cperm; J.thr = cthr; J.mlo = cmlo; J.mhi = cmhi; J.gt0 = 0; J.tg = T;
        rc = rpf_launch_job(h, J, JP, d_jtab);
        if (rc) return rc;
        // ---- 3. thr' = (thr0 + thr) / 2, margin' = margin0 <> margin (or plain set for the first chunk of an empty tree)
        if (!P.upd_g.empty()) {
            const int64_t tot = (int64_t)P.upd_g.size() * T;
            RPF_LAUNCH(h, PH_STREAM, k_pool_update, (unsigned)((tot + 255) / 256), 256, 0, d_upd_g, d_upd_u, (int)P.upd_g.size(), T, nnct,
                       cthr, cmlo, cmhi, pthr, pmlo, pmhi, P.ct_set ? 0 : 1);
        }
    }
    // ---- 4. Tip contents: piece ++ old, in the new left-to-right layout
    RPF_CUDA(h, cudaMemsetAsync(arena_new, 0xff, (size_t)T * cap * 4, h->stream));
    if (!P.copies.empty() && P.kept > 0)
        RPF_LAUNCH(h, PH_STREAM_CONCAT, k_tip_concat, dim3((unsigned)((P.kept + 255) / 256), (unsigned)((T + 7) / 8)), 256, 0, d_copies, (int)P.copies.size(), T,
                   (uint32_t)P.kept, cperm, (int64_t)m, (uint32_t)n0, arena_old, arena_new, cap);
    // ---- 5. Tips that outgrew minLeaf split in place
    for (size_t i = 0; i < P.groups.size(); ++i) {
        const RtGroup& G = P.groups[i];
        BottomArgs A{};
        A.ks = cap; A.ps = cap; A.nn_all = nnrt; A.L = Lk; A.s = G.depth; A.nlb = G.nlb; A.gt0 = 0; A.first_gid = G.first;
        A.given_order = 1;
        A.keys = S->keys; A.perm = arena_new; A.child = d_rt_child; A.nstart = d_rt_start; A.nsize = d_rt_size;
        A.range = d_rg[i]; A.lvl_pv = d_pv[i];
        A.thr = rthr; A.mlo = rmlo; A.mhi = rmhi;
        A.kmin = m > 0 ? kmin : nullptr; A.kmax = m > 0 ? kmax : nullptr;
        rc = rpf_bottom_launch(h, A, G.count, T, G.fast, G.max_root, G.levels);
        if (rc) return rc;
    }
    if (!P.rt_upd_g.empty()) {
        const int64_t tot = (int64_t)P.rt_upd_g.size() * T;
        RPF_LAUNCH(h, PH_STREAM, k_pool_update, (unsigned)((tot + 255) / 256), 256, 0, d_rt_upd_g, d_rt_upd_u, (int)P.rt_upd_g.size(), T, nnrt,
                   rthr, rmlo, rmhi, pthr, pmlo, pmhi, 0);
    }
    S->cur ^= 1;
    S->n = n1;

    // ---- canonical forest over everything inserted so far
    h->n = n1;
    h->stream_lost = S->SP.lost;
    if (m > 0 && !JP.G.order_exact) S->order_exact = false;
    h->leaf_order_exact = S->order_exact;
    rc = rpf_upload_topology(h);
    if (rc) return rc;
    const int64_t nn = h->topo.nnodes();
    rc = rpf_alloc_forest(h, nn, n1);
    if (rc) return rc;
    RPF_LAUNCH(h, PH_STREAM, k_pool_export, (unsigned)((nn * T + 255) / 256), 256, 0, d_pool_of, h->d_node_child, nn, T, 0,
               pthr, pmlo, pmhi, h->d_thr, h->d_mlo, h->d_mhi);
    if (n1 > 0)
        RPF_CUDA(h, cudaMemcpy2DAsync(h->d_perm, (size_t)n1 * 4, S->arena[S->cur], (size_t)cap * 4, (size_t)n1 * 4, T, cudaMemcpyDeviceToDevice, h->stream));
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
