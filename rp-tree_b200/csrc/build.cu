// build.cu -- forest construction on sm_100a.
//
// Replaces, for dense Double data, the reference's
//   forestBatch / createMulti / create / insert (Tip case)   src/Data/RPTree/Batch.hs:57-63, Internal.hs:217-297
//   partitionAtMedian + sortByVG                              src/Data/RPTree/Internal.hs:484-512
//   innerSD                                                   src/Data/RPTree/Internal.hs:369-382
//
// Design: the reference uses ONE hyperplane per tree LEVEL and a POSITIONAL median split,
// so (a) every projection key[t][l][i] can be computed up front in one pass over X (k_project), and (b) the
// tree topology is data independent (Topology).  The build is level synchronous over all trees of a group:
//   top phase   (nodes larger than the shared-memory capacity): streaming, gather-free.  Every point carries
//               the BFS id of its node (label); per level an exact median is found by histogram -> median-bin
//               compaction -> in-bin sort, and labels are rewritten.  Ties at the threshold are resolved by the
//               reference's rule (stable sort => order of the previous levels' keys, then row id) with a
//               lexicographic select that only runs when a tie straddles the split.
//   bottom phase (nodes <= capacity): one CTA owns a node and sorts its whole subtree level by level in shared
//               memory with a (key, incoming position) bitonic network == the reference's stable sort.
// All comparisons are on order-preserving uint64 images of the doubles; projections use __dmul_rn/__dadd_rn in
// the reference's right-fold order, so thresholds/margins/leaf sets are bit exact.
#include "rpf_internal.h"
#include <algorithm>
#include <cstdio>

typedef unsigned long long ull;

// =====================================================================================================
// K1  projections: key[h][i] = hp[h] . X[i]   (innerSD, right fold, no FMA)
// =====================================================================================================
// Tile of P points staged in shared memory (row stride ld = d|1 doubles -> conflict-free column gathers),
// thread (p, g) walks hyperplanes g, g+G, ...  Output row j corresponds to CSR row
// (t0 + j / L) * hpDepth + (j % L).  ORD: write order-preserving uint64 + track per-row min/max.
template <int NT, bool ORD>
__global__ void __launch_bounds__(NT) k_project(const double* __restrict__ X, int64_t n, int d, int ld, int P,
                                                 const int64_t* __restrict__ hp_off, const int32_t* __restrict__ hp_idx,
                                                 const double* __restrict__ hp_val, int t0, int L, int hpDepth, int H,
                                                 void* __restrict__ out, ull* __restrict__ kmin, ull* __restrict__ kmax) {
    extern __shared__ double xs[];
    const int tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * P;
    const int rows = (int)min((int64_t)P, n - i0);
    // coalesced tile load
    for (int r = 0; r < rows; ++r) {
        const double* src = X + (i0 + r) * (int64_t)d;
        for (int c = tid; c < d; c += NT) xs[r * ld + c] = src[c];
    }
    __syncthreads();
    const int p = tid % P, g = tid / P, G = NT / P;
    const bool live = p < rows;
    const double* xr = xs + p * ld;
    for (int jb = 0; jb < H; jb += G) {      // uniform trip count: the warp shuffles below need every lane
        const int j = jb + g;
        const bool hv = j < H;
        double acc = 0.0;
        if (hv && live) {
            const int row = (t0 + j / L) * hpDepth + (j % L);
            const int64_t s = hp_off[row];
            int64_t e = hp_off[row + 1];
            if (e - s > d) e = s + d;   // innerSD's `i >= nz2` guard (Internal.hs:376)
            for (int64_t q = e - 1; q >= s; --q) acc = __dadd_rn(__dmul_rn(__ldg(hp_val + q), xr[__ldg(hp_idx + q)]), acc);
        }
        if (ORD) {
            const ull o = f2ord(acc);
            if (hv && live) ((ull*)out)[(int64_t)j * n + i0 + p] = o;
            ull vmin = (hv && live) ? o : ORD_NONE_HI, vmax = (hv && live) ? o : ORD_NONE_LO;
            const int w = P < 32 ? P : 32;
            for (int off = w >> 1; off > 0; off >>= 1) {
                ull a = __shfl_xor_sync(0xffffffffu, vmin, off);
                ull b = __shfl_xor_sync(0xffffffffu, vmax, off);
                vmin = a < vmin ? a : vmin;
                vmax = b > vmax ? b : vmax;
            }
            if (hv && (tid & (w - 1)) == 0 && vmin != ORD_NONE_HI) {
                if (vmin < kmin[j]) atomicMin(&kmin[j], vmin);
                if (vmax > kmax[j]) atomicMax(&kmax[j], vmax);
            }
        } else {
            if (hv && live) ((double*)out)[(int64_t)j * n + i0 + p] = acc;
        }
    }
}

static int project_tile_points(int d) {
    // shared bytes = P * (d|1) * 8 ; keep <= ~72 KB so 3 CTAs fit an SM
    int ld = d | 1;
    int P = 64;
    while (P > 1 && (size_t)P * ld * 8 > 72 * 1024) P >>= 1;
    return P;
}

int rpf_project_launch(rpf_handle* h, int phase, const double* dX, int64_t n, int t0, int Tg, int L, bool ord, void* out,
                       ull* kmin, ull* kmax) {
    const int d = h->d, ld = d | 1;
    const int P = project_tile_points(d);
    const size_t smem = (size_t)P * ld * sizeof(double);
    if (smem > 200 * 1024) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "dimension too large for the projection tile");
    const int H = Tg * L;
    const int64_t grid = (n + P - 1) / P;
    if (grid <= 0 || H <= 0) return RPF_OK;
    constexpr int NT = 256;
    if (ord) {
        auto kfn = k_project<NT, true>;
        RPF_CUDA(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RPF_LAUNCH(h, phase, kfn, (unsigned)grid, NT, smem, dX, n, d, ld, P, h->d_hp_off, h->d_hp_idx,
                   h->d_hp_val, t0, L, h->hpDepth, H, out, kmin, kmax);
    } else {
        auto kfn = k_project<NT, false>;
        RPF_CUDA(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RPF_LAUNCH(h, phase, kfn, (unsigned)grid, NT, smem, dX, n, d, ld, P, h->d_hp_off, h->d_hp_idx,
                   h->d_hp_val, t0, L, h->hpDepth, H, out, kmin, kmax);
    }
    return RPF_OK;
}

// =====================================================================================================
// shared bitonic helpers (normalised network: every comparator puts the smaller element at the lower
// index, so positions >= m behave as +inf padding and comparators touching them are skipped)
// =====================================================================================================
__device__ __forceinline__ int ilog2_pow2(unsigned v) { return 31 - __clz(v); }
__device__ __forceinline__ unsigned next_pow2_u32(unsigned v) { return v <= 1 ? 1u : 1u << (32 - __clz(v - 1)); }

// sort m uint64 keys ascending in shared memory
template <int NT>
__device__ void bitonic_keys(ull* buf, unsigned m) {
    const unsigned Pv = next_pow2_u32(m), half = Pv >> 1;
    for (unsigned k = 2; k <= Pv; k <<= 1) {
        const int lk = ilog2_pow2(k);
        // flip stage
        for (unsigned c = threadIdx.x; c < half; c += NT) {
            const unsigned blk = c >> (lk - 1), w = c & ((k >> 1) - 1);
            const unsigned i = (blk << lk) + w, p = (blk << lk) + (k - 1 - w);
            if (p < m) { ull a = buf[i], b = buf[p]; if (a > b) { buf[i] = b; buf[p] = a; } }
        }
        __syncthreads();
        for (unsigned j = k >> 2; j > 0; j >>= 1) {
            const int lj = ilog2_pow2(j);
            for (unsigned c = threadIdx.x; c < half; c += NT) {
                const unsigned i = ((c >> lj) << (lj + 1)) + (c & (j - 1)), p = i + j;
                if (p < m) { ull a = buf[i], b = buf[p]; if (a > b) { buf[i] = b; buf[p] = a; } }
            }
            __syncthreads();
        }
    }
}

// =====================================================================================================
// top phase
// =====================================================================================================
struct NodeSel {
    ull thr, pred, succ;
    int32_t sel_bin;
    uint32_t below, cand_off, cand_cnt, cand_fill, cless, ceq, tie_r;
    int32_t tie_depth;
    int32_t pad_;
};

struct TopArgs {
    int64_t n;
    int Tg, L, l, node0, nnodes, NTOP, NB, HSZ, MAXTD, smem_hist, gt0;   // gt0: global tree id of the group's first tree
    int64_t nn_all;                                                         // nodes per tree (stride of thr/mlo/mhi)
    const ull* keys;
    uint16_t* label;
    const int32_t* child;
    const uint32_t* nstart;
    const uint32_t* nsize;
    double* binlo;
    double* binscale;
    const ull* kmin;
    const ull* kmax;
    uint32_t* hist;
    NodeSel* sel;
    ull* cand;
    uint32_t* cand_total;
    ull* pivots;
    uint32_t* fill;
    uint32_t* perm;
    double *thr, *mlo, *mhi;
};

#define TOP_CH 32768      /* points per CTA in the streaming top-phase kernels */
#define TOP_NT 512
#define HBINS 8192        /* shared-memory histogram counters */
#define FIN_CAP 4096      /* in-bin sort capacity */
#define SMEM_NODES 256    /* relabel keeps per-node state in shared memory up to this many nodes */

__device__ __forceinline__ int key_bin(ull o, double lo, double sc, int NB) {
    double v = (ord2f(o) - lo) * sc;
    int b = (int)v;
    return b < 0 ? 0 : (b >= NB ? NB - 1 : b);
}

// per (tree, level): linear bin map from the key range
__global__ void k_bin_setup(TopArgs A, const int* __restrict__ nb_per_level, int s_top) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= A.Tg * A.L) return;
    int l = idx % A.L;
    double lo = 0, sc = 0;
    if (l < s_top) {
        lo = ord2f(A.kmin[idx]);
        double hi = ord2f(A.kmax[idx]);
        double w = hi - lo;
        sc = (w > 0 && isfinite(w)) ? (double)nb_per_level[l] / w : 0.0;
        if (!isfinite(sc)) sc = 0.0;
    }
    A.binlo[idx] = lo;
    A.binscale[idx] = sc;
}

// point -> (local node index at level l) or -1 when the point does not sit in an internal node of level l
__device__ __forceinline__ int point_node(const TopArgs& A, const uint16_t* lab, int64_t i) {
    int g = A.l == 0 ? 0 : (int)lab[i];
    int nl = g - A.node0;
    if ((unsigned)nl >= (unsigned)A.nnodes) return -1;
    if (__ldg(A.child + g) < 0) return -1;
    return nl;
}

__global__ void __launch_bounds__(TOP_NT) k_top_hist(TopArgs A) {
    extern __shared__ uint32_t sh[];
    const int t = blockIdx.y, tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * TOP_CH, i1 = min(A.n, i0 + TOP_CH);
    const ull* keys = A.keys + ((int64_t)t * A.L + A.l) * A.n;
    const uint16_t* lab = A.label + (int64_t)t * A.n;
    const double lo = A.binlo[t * A.L + A.l], sc = A.binscale[t * A.L + A.l];
    const int NB = A.NB, tot = A.nnodes * NB;
    uint32_t* gh = A.hist + (int64_t)t * A.HSZ;
    if (A.smem_hist) {
        for (int j = tid; j < tot; j += TOP_NT) sh[j] = 0;
        __syncthreads();
    }
    for (int64_t i = i0 + tid; i < i1; i += TOP_NT) {
        int nl = point_node(A, lab, i);
        if (nl < 0) continue;
        int b = key_bin(keys[i], lo, sc, NB);
        if (A.smem_hist) atomicAdd(&sh[nl * NB + b], 1u);
        else atomicAdd(&gh[nl * NB + b], 1u);
    }
    if (A.smem_hist) {
        __syncthreads();
        for (int j = tid; j < tot; j += TOP_NT) { uint32_t v = sh[j]; if (v) atomicAdd(&gh[j], v); }
    }
}

// one CTA per (node, tree): find the bin holding rank nh = size/2
__global__ void __launch_bounds__(256) k_top_pick(TopArgs A) {
    __shared__ uint32_t part[256];
    __shared__ uint32_t excl[257];
    const int t = blockIdx.y, nl = blockIdx.x, g = A.node0 + nl, tid = threadIdx.x;
    if (A.child[g] < 0) return;
    const uint32_t k = A.nsize[g] >> 1;
    const uint32_t* hr = A.hist + (int64_t)t * A.HSZ + (int64_t)nl * A.NB;
    const int per = (A.NB + 255) / 256;
    const int b0 = tid * per, b1 = min(A.NB, b0 + per);
    uint32_t s = 0;
    for (int b = b0; b < b1; ++b) s += hr[b];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) { uint32_t c = 0; for (int j = 0; j < 256; ++j) { excl[j] = c; c += part[j]; } excl[256] = c; }
    __syncthreads();
    if (k >= excl[tid] && k < excl[tid] + part[tid]) {
        uint32_t c = excl[tid];
        for (int b = b0; b < b1; ++b) {
            uint32_t hb = hr[b];
            if (k < c + hb) {
                NodeSel& S = A.sel[(int64_t)t * A.NTOP + g];
                S.sel_bin = b; S.below = c; S.cand_cnt = hb; S.cand_fill = 0;
                S.cand_off = atomicAdd(&A.cand_total[t], hb);
                break;
            }
            c += hb;
        }
    }
}

__global__ void __launch_bounds__(TOP_NT) k_top_compact(TopArgs A) {
    __shared__ int s_bin[SMEM_NODES];
    const int t = blockIdx.y, tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * TOP_CH, i1 = min(A.n, i0 + TOP_CH);
    const ull* keys = A.keys + ((int64_t)t * A.L + A.l) * A.n;
    const uint16_t* lab = A.label + (int64_t)t * A.n;
    const double lo = A.binlo[t * A.L + A.l], sc = A.binscale[t * A.L + A.l];
    NodeSel* sel = A.sel + (int64_t)t * A.NTOP + A.node0;
    const bool cached = A.nnodes <= SMEM_NODES;
    if (cached) {
        for (int j = tid; j < A.nnodes; j += TOP_NT) s_bin[j] = sel[j].sel_bin;
        __syncthreads();
    }
    ull* cand = A.cand + (int64_t)t * A.n;
    for (int64_t i = i0 + tid; i < i1; i += TOP_NT) {
        int nl = point_node(A, lab, i);
        if (nl < 0) continue;
        ull kv = keys[i];
        int b = key_bin(kv, lo, sc, A.NB);
        int sb = cached ? s_bin[nl] : sel[nl].sel_bin;
        if (b == sb) {
            uint32_t pos = atomicAdd(&sel[nl].cand_fill, 1u);
            cand[sel[nl].cand_off + pos] = kv;
        }
    }
}

// 8-bit MSD radix select of rank r over `c` values fetched by `get(i)`; all threads of the CTA call it.
// Returns the selected value; cl = #values < it, ce = #values == it.
template <int NT, typename Get>
__device__ ull cta_radix_select(uint32_t c, uint32_t r, Get get, uint32_t* sh /*256+*/, ull* sh64 /*1*/, uint32_t& cl, uint32_t& ce) {
    ull prefix = 0;
    uint32_t rr = r, below = 0;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        const ull mask_hi = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int j = threadIdx.x; j < 256; j += NT) sh[j] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < c; i += NT) {
            ull v = get(i);
            if ((v & mask_hi) == prefix) atomicAdd(&sh[(v >> shift) & 255], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t cum = 0; int dg = 255;
            for (int b = 0; b < 256; ++b) { if (rr < cum + sh[b]) { dg = b; break; } cum += sh[b]; }
            sh[256] = cum; sh[257] = sh[dg];
            *sh64 = prefix | ((ull)dg << shift);
        }
        __syncthreads();
        prefix = *sh64;
        rr -= sh[256];
        below += sh[256];
        ce = sh[257];
        __syncthreads();
    }
    cl = below;
    return prefix;
}

// one CTA per (node, tree): exact order statistic inside the median bin
__global__ void __launch_bounds__(512) k_top_finish(TopArgs A) {
    __shared__ ull buf[FIN_CAP];
    __shared__ uint32_t sh[260];
    __shared__ ull sh64[3];
    const int t = blockIdx.y, nl = blockIdx.x, g = A.node0 + nl, tid = threadIdx.x;
    if (A.child[g] < 0) return;
    NodeSel& S = A.sel[(int64_t)t * A.NTOP + g];
    const uint32_t c = S.cand_cnt, nh = A.nsize[g] >> 1, r = nh - S.below;
    const ull* seg = A.cand + (int64_t)t * A.n + S.cand_off;
    ull thr, pred = ORD_NONE_LO, succ = ORD_NONE_HI;
    uint32_t lower, ceq;
    if (c <= FIN_CAP) {
        for (uint32_t i = tid; i < c; i += 512) buf[i] = seg[i];
        __syncthreads();
        bitonic_keys<512>(buf, c);
        thr = buf[r];
        if (tid == 0) { sh[0] = 0xffffffffu; sh[1] = 0; }
        __syncthreads();
        for (uint32_t i = tid; i < c; i += 512)
            if (buf[i] == thr) { atomicMin(&sh[0], i); atomicMax(&sh[1], i + 1); }
        __syncthreads();
        lower = sh[0];
        ceq = sh[1] - sh[0];
        if (lower > 0) pred = buf[lower - 1];
        if (sh[1] < c) succ = buf[sh[1]];
    } else {
        uint32_t cl, ce;
        thr = cta_radix_select<512>(c, r, [&](uint32_t i) { return seg[i]; }, sh, sh64, cl, ce);
        lower = cl; ceq = ce;
        if (tid == 0) { sh64[1] = ORD_NONE_LO; sh64[2] = ORD_NONE_HI; }
        __syncthreads();
        ull lp = ORD_NONE_LO, ls = ORD_NONE_HI;
        for (uint32_t i = tid; i < c; i += 512) {
            ull v = seg[i];
            if (v < thr && v > lp) lp = v;
            if (v > thr && v < ls) ls = v;
        }
        if (lp != ORD_NONE_LO) atomicMax(&sh64[1], lp);
        if (ls != ORD_NONE_HI) atomicMin(&sh64[2], ls);
        __syncthreads();
        pred = sh64[1]; succ = sh64[2];
    }
    if (tid == 0) {
        S.thr = thr; S.pred = pred; S.succ = succ;
        S.cless = S.below + lower; S.ceq = ceq;
        S.tie_r = nh - S.cless;      // tied points that must go left; > 0 => the split cuts through a tie
        S.tie_depth = 0;
    }
}

// one CTA per (node, tree): only does work when a tie straddles the split.  Finds the composite pivot
// (key_{l-1}, key_{l-2}, ..., key_0, row id) such that exactly tie_r tied points are lexicographically below it:
// this is the order the reference's stable merge sort leaves tied points in (Internal.hs:504-512).
__global__ void __launch_bounds__(512) k_top_ties(TopArgs A) {
    __shared__ uint32_t sh[260];
    __shared__ ull sh64[1];
    __shared__ uint32_t cnt;
    const int t = blockIdx.y, nl = blockIdx.x, g = A.node0 + nl, tid = threadIdx.x;
    if (A.child[g] < 0) return;
    NodeSel& S = A.sel[(int64_t)t * A.NTOP + g];
    if (S.tie_r == 0) return;
    const ull* keys_t = A.keys + (int64_t)t * A.L * A.n;
    const ull* keys_l = keys_t + (int64_t)A.l * A.n;
    const uint16_t* lab = A.label + (int64_t)t * A.n;
    const ull thr = S.thr;
    uint32_t* la = (uint32_t*)(A.cand + (int64_t)t * A.n) + 2 * (int64_t)A.nstart[g];
    uint32_t* lb = la + A.nsize[g];
    if (tid == 0) cnt = 0;
    __syncthreads();
    for (int64_t i = tid; i < A.n; i += 512) {
        int gi = A.l == 0 ? 0 : (int)lab[i];
        if (gi == g && keys_l[i] == thr) { uint32_t p = atomicAdd(&cnt, 1u); la[p] = (uint32_t)i; }
    }
    __syncthreads();
    uint32_t c = cnt, rr = S.tie_r;
    int depth = 0;
    ull* piv = A.pivots + ((int64_t)t * A.NTOP + g) * A.MAXTD;
    while (true) {
        const int lvl = A.l - 1 - depth;
        const ull* kk = lvl >= 0 ? keys_t + (int64_t)lvl * A.n : nullptr;
        uint32_t cl, ce;
        ull pv = cta_radix_select<512>(c, rr, [&](uint32_t i) { uint32_t id = la[i]; return kk ? kk[id] : (ull)id; }, sh, sh64, cl, ce);
        if (tid == 0) piv[depth] = pv;
        ++depth;
        const uint32_t r2 = rr - cl;
        if (r2 == 0 || lvl < 0) break;
        if (tid == 0) cnt = 0;
        __syncthreads();
        for (uint32_t i = tid; i < c; i += 512) {
            uint32_t id = la[i];
            if (kk[id] == pv) { uint32_t p = atomicAdd(&cnt, 1u); lb[p] = id; }
        }
        __syncthreads();
        c = cnt; rr = r2;
        uint32_t* tmp = la; la = lb; lb = tmp;
        __syncthreads();
    }
    if (tid == 0) S.tie_depth = depth;
}

// relabel every point of an internal level-l node to its child; track the keys adjacent to the threshold
// (margins); at the last top level also scatter the points into the per-node segments of perm.
__global__ void __launch_bounds__(TOP_NT) k_top_relabel(TopArgs A, int last) {
    __shared__ ull s_thr[SMEM_NODES], s_pred[SMEM_NODES], s_succ[SMEM_NODES];
    __shared__ uint32_t s_tie[SMEM_NODES];
    const int t = blockIdx.y, tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * TOP_CH, i1 = min(A.n, i0 + TOP_CH);
    const ull* keys_t = A.keys + (int64_t)t * A.L * A.n;
    const ull* keys = keys_t + (int64_t)A.l * A.n;
    uint16_t* lab = A.label + (int64_t)t * A.n;
    NodeSel* sel = A.sel + (int64_t)t * A.NTOP + A.node0;
    const bool cached = A.nnodes <= SMEM_NODES;
    if (cached) {
        for (int j = tid; j < A.nnodes; j += TOP_NT) {
            s_thr[j] = sel[j].thr; s_pred[j] = sel[j].pred; s_succ[j] = sel[j].succ; s_tie[j] = sel[j].tie_r;
        }
        __syncthreads();
    }
    uint32_t* fill = A.fill + (int64_t)t * A.NTOP;
    uint32_t* perm = A.perm + (int64_t)t * A.n;
    for (int64_t i = i0 + tid; i < i1; i += TOP_NT) {
        int g = A.l == 0 ? 0 : (int)lab[i];
        int nl = g - A.node0;
        int ch = ((unsigned)nl < (unsigned)A.nnodes) ? __ldg(A.child + g) : -1;
        if (ch >= 0) {
            const ull kv = keys[i];
            const ull thr = cached ? s_thr[nl] : sel[nl].thr;
            bool left = kv < thr;
            if (kv == thr) {
                const uint32_t tr = cached ? s_tie[nl] : sel[nl].tie_r;
                if (tr > 0) {   // composite compare against the tie pivot
                    const int td = sel[nl].tie_depth;
                    const ull* piv = A.pivots + ((int64_t)t * A.NTOP + g) * A.MAXTD;
                    for (int j = 0; j < td; ++j) {
                        const int lvl = A.l - 1 - j;
                        const ull kq = lvl >= 0 ? keys_t[(int64_t)lvl * A.n + i] : (ull)i;
                        const ull pv = piv[j];
                        if (kq != pv) { left = kq < pv; break; }
                    }
                }
            } else if (kv < thr) {
                if (cached) { if (kv > s_pred[nl]) atomicMax(&s_pred[nl], kv); }
                else if (kv > sel[nl].pred) atomicMax(&sel[nl].pred, kv);
            } else {
                if (cached) { if (kv < s_succ[nl]) atomicMin(&s_succ[nl], kv); }
                else if (kv < sel[nl].succ) atomicMin(&sel[nl].succ, kv);
            }
            g = ch + (left ? 0 : 1);
            lab[i] = (uint16_t)g;
        }
        if (last) {
            uint32_t pos = atomicAdd(&fill[g], 1u);
            perm[A.nstart[g] + pos] = (uint32_t)i;
        }
    }
    if (cached) {
        __syncthreads();
        for (int j = tid; j < A.nnodes; j += TOP_NT) {
            if (s_pred[j] > sel[j].pred) atomicMax(&sel[j].pred, s_pred[j]);
            if (s_succ[j] < sel[j].succ) atomicMin(&sel[j].succ, s_succ[j]);
        }
    }
}

// thr / margins of the level's nodes (Internal.hs:496-501); top-phase nodes always have size >= 3
__global__ void k_top_finalize(TopArgs A) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= A.nnodes * A.Tg) return;
    int t = idx / A.nnodes, nl = idx % A.nnodes, g = A.node0 + nl;
    if (A.child[g] < 0) return;
    const NodeSel& S = A.sel[(int64_t)t * A.NTOP + g];
    const uint32_t nh = A.nsize[g] >> 1;
    const ull lo = (S.cless == nh) ? S.pred : S.thr;                 // sorted[nh-1]
    const ull hi = (S.cless + S.ceq >= nh + 2) ? S.thr : S.succ;     // sorted[nh+1]
    const int64_t o = (int64_t)(A.gt0 + t) * A.nn_all + g;
    A.thr[o] = ord2f(S.thr);
    A.mlo[o] = ord2f(lo);
    A.mhi[o] = ord2f(hi);
}

__global__ void k_iota_perm(uint32_t* perm, int64_t n, int Tg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * Tg) perm[i] = (uint32_t)(i % n);
}

// =====================================================================================================
// bottom phase: one CTA = one level-s node and its whole subtree, resident in shared memory
// =====================================================================================================
struct BottomArgs {
    int64_t n, nn_all;
    int L, s, nlb, gt0, first_gid;      // nlb = levels recorded per node in `range`
    const ull* keys;                     // [Tg][L][n]
    uint32_t* perm;                      // [Tg][n]  (in: node segments in any order; out: final leaf order)
    const int32_t* child;
    const uint32_t* nstart;
    const uint32_t* nsize;
    const int2* range;                   // [nodes at level s][nlb]: BFS id range of the descendants that split
    const uint32_t* lvl_pv;              // per level: next_pow2(max node size)
    double *thr, *mlo, *mhi;
};

#define BOT_EMAX 1024
#define POS_NONE 0xffffffffu

template <int CAP, int NT>
__global__ void __launch_bounds__(NT) k_bottom(BottomArgs A) {
    extern __shared__ unsigned char smraw[];
    ull* skey = (ull*)smraw;
    uint32_t* spos = (uint32_t*)(skey + CAP);
    uint32_t* sidx = spos + CAP;
    __shared__ uint16_t s_off[BOT_EMAX], s_sz[BOT_EMAX];
    const int t = blockIdx.y, tid = threadIdx.x;
    const int e0 = A.first_gid + blockIdx.x;
    const uint32_t m = A.nsize[e0], start = A.nstart[e0];
    if (m == 0) return;
    const int64_t n = A.n;
    const ull* keys_t = A.keys + (int64_t)t * A.L * n;
    uint32_t* perm = A.perm + (int64_t)t * n + start;

    for (uint32_t p = tid; p < m; p += NT) sidx[p] = perm[p];
    __syncthreads();

    if (A.s > 0) {
        // Establish the reference's incoming order O_s: lexicographic (key_{s-1}, ..., key_0, row id).
        const ull* k1 = keys_t + (int64_t)(A.s - 1) * n;
        for (uint32_t p = tid; p < m; p += NT) skey[p] = k1[sidx[p]];
        __syncthreads();
        auto after = [&](ull ka, uint32_t ia, ull kb, uint32_t ib) -> bool {   // a sorts after b
            if (ka != kb) return ka > kb;
            for (int lvl = A.s - 2; lvl >= 0; --lvl) {
                ull xa = keys_t[(int64_t)lvl * n + ia], xb = keys_t[(int64_t)lvl * n + ib];
                if (xa != xb) return xa > xb;
            }
            return ia > ib;
        };
        const unsigned Pv = next_pow2_u32(m), half = Pv >> 1;
        for (unsigned k = 2; k <= Pv; k <<= 1) {
            const int lk = ilog2_pow2(k);
            for (unsigned c = tid; c < half; c += NT) {
                const unsigned blk = c >> (lk - 1), w = c & ((k >> 1) - 1);
                const unsigned i = (blk << lk) + w, p = (blk << lk) + (k - 1 - w);
                if (p < m) {
                    ull a = skey[i], b = skey[p]; uint32_t ia = sidx[i], ib = sidx[p];
                    if (after(a, ia, b, ib)) { skey[i] = b; skey[p] = a; sidx[i] = ib; sidx[p] = ia; }
                }
            }
            __syncthreads();
            for (unsigned j = k >> 2; j > 0; j >>= 1) {
                const int lj = ilog2_pow2(j);
                for (unsigned c = tid; c < half; c += NT) {
                    const unsigned i = ((c >> lj) << (lj + 1)) + (c & (j - 1)), p = i + j;
                    if (p < m) {
                        ull a = skey[i], b = skey[p]; uint32_t ia = sidx[i], ib = sidx[p];
                        if (after(a, ia, b, ib)) { skey[i] = b; skey[p] = a; sidx[i] = ib; sidx[p] = ia; }
                    }
                }
                __syncthreads();
            }
        }
    }

    const int2* rng = A.range + (int64_t)blockIdx.x * A.nlb;
    for (int j = 0; j < A.nlb; ++j) {
        const int l = A.s + j;
        const int2 r = rng[j];
        const int lo = r.x, nent = r.y - r.x;
        if (nent <= 0) break;
        const ull* kl = keys_t + (int64_t)l * n;
        const bool tab = nent <= BOT_EMAX;
        if (tab) {
            for (int e = tid; e < nent; e += NT) {
                const int g = lo + e;
                s_off[e] = (uint16_t)(A.nstart[g] - start);
                s_sz[e] = (uint16_t)(A.child[g] >= 0 ? A.nsize[g] : 0);
            }
        }
        for (uint32_t p = tid; p < m; p += NT) spos[p] = POS_NONE;
        __syncthreads();
        auto entry = [&](int e, uint32_t& off, uint32_t& sz) {
            if (tab) { off = s_off[e]; sz = s_sz[e]; }
            else { const int g = lo + e; off = A.nstart[g] - start; sz = A.child[g] >= 0 ? A.nsize[g] : 0; }
        };
        const unsigned Pv = A.lvl_pv[l], half = Pv >> 1;
        const int lpv = ilog2_pow2(Pv);
        // gather this level's keys for the points of the nodes that split
        for (unsigned v = tid; v < (unsigned)nent * Pv; v += NT) {
            const int e = v >> lpv; const unsigned i = v & (Pv - 1);
            uint32_t off, sz; entry(e, off, sz);
            if (i < sz) { const uint32_t p = off + i; skey[p] = kl[sidx[p]]; spos[p] = p; }
        }
        __syncthreads();
        // segmented stable sort: (key, incoming position) ascending == Merge.sortBy (comparing snd)
        if (half > 0) {
            const int lh = ilog2_pow2(half);
            const unsigned ncmp = (unsigned)nent * half;
            for (unsigned k = 2; k <= Pv; k <<= 1) {
                const int lk = ilog2_pow2(k);
                for (unsigned c = tid; c < ncmp; c += NT) {
                    const int e = c >> lh; const unsigned cc = c & (half - 1);
                    const unsigned blk = cc >> (lk - 1), w = cc & ((k >> 1) - 1);
                    const unsigned i = (blk << lk) + w, p = (blk << lk) + (k - 1 - w);
                    uint32_t off, sz; entry(e, off, sz);
                    if (p < sz) {
                        const uint32_t xi = off + i, xp = off + p;
                        ull a = skey[xi], b = skey[xp]; uint32_t pa = spos[xi], pb = spos[xp];
                        if (a > b || (a == b && pa > pb)) { skey[xi] = b; skey[xp] = a; spos[xi] = pb; spos[xp] = pa; }
                    }
                }
                __syncthreads();
                for (unsigned jj = k >> 2; jj > 0; jj >>= 1) {
                    const int lj = ilog2_pow2(jj);
                    for (unsigned c = tid; c < ncmp; c += NT) {
                        const int e = c >> lh; const unsigned cc = c & (half - 1);
                        const unsigned i = ((cc >> lj) << (lj + 1)) + (cc & (jj - 1)), p = i + jj;
                        uint32_t off, sz; entry(e, off, sz);
                        if (p < sz) {
                            const uint32_t xi = off + i, xp = off + p;
                            ull a = skey[xi], b = skey[xp]; uint32_t pa = spos[xi], pb = spos[xp];
                            if (a > b || (a == b && pa > pb)) { skey[xi] = b; skey[xp] = a; spos[xi] = pb; spos[xp] = pa; }
                        }
                    }
                    __syncthreads();
                }
            }
        }
        // thresholds / margins at the sorted positions (Internal.hs:496-503)
        for (int e = tid; e < nent; e += NT) {
            uint32_t off, sz; entry(e, off, sz);
            if (sz == 0) continue;
            const uint32_t nh = sz >> 1;
            ull th = skey[off + nh], ml, mh;
            if (sz >= 3) { ml = skey[off + nh - 1]; mh = skey[off + nh + 1]; }
            else if (sz == 2) { ml = skey[off]; mh = skey[off + 1]; }
            else { ml = skey[off]; mh = ml; }
            const int64_t o = (int64_t)(A.gt0 + t) * A.nn_all + (lo + e);
            A.thr[o] = ord2f(th); A.mlo[o] = ord2f(ml); A.mhi[o] = ord2f(mh);
        }
        // apply the permutation to the row ids
        uint32_t tmp[CAP / NT];
#pragma unroll
        for (int q = 0; q < CAP / NT; ++q) {
            const uint32_t p = tid + q * NT;
            tmp[q] = POS_NONE;
            if (p < m) { const uint32_t src = spos[p]; if (src != POS_NONE) tmp[q] = sidx[src]; }
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < CAP / NT; ++q) {
            const uint32_t p = tid + q * NT;
            if (p < m && spos[p] != POS_NONE) sidx[p] = tmp[q];
        }
        __syncthreads();
    }
    for (uint32_t p = tid; p < m; p += NT) perm[p] = sidx[p];
}

// =====================================================================================================
// host orchestration
// =====================================================================================================
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <typename T> T* as() { return (T*)p; }
};

static unsigned next_pow2_host(unsigned v) { unsigned p = 1; while (p < v) p <<= 1; return p; }

template <int CAP, int NT>
static int launch_bottom(rpf_handle* h, const BottomArgs& B, int nnodes_s, int tg) {
    const size_t smem = (size_t)CAP * 16;
    auto kfn = k_bottom<CAP, NT>;
    RPF_CUDA(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)nnodes_s, (unsigned)tg);
    RPF_LAUNCH(h, PH_BOTTOM, kfn, grid, NT, smem, B);
    return RPF_OK;
}

int rpf_build_impl(rpf_handle* h) {
    const Topology& tp = h->topo;
    const int64_t n = h->n, nn = tp.nnodes();
    const int T = h->T, L = tp.L_eff, CAP = h->bottom_cap;
    h->leaf_order_exact = true;

    // ---- result arrays
    if (h->d_thr) { cudaFree(h->d_thr); cudaFree(h->d_mlo); cudaFree(h->d_mhi); cudaFree(h->d_perm); h->d_thr = h->d_mlo = h->d_mhi = nullptr; h->d_perm = nullptr; }
    RPF_CUDA(h, cudaMalloc(&h->d_thr, sizeof(double) * (size_t)(T * nn)));
    RPF_CUDA(h, cudaMalloc(&h->d_mlo, sizeof(double) * (size_t)(T * nn)));
    RPF_CUDA(h, cudaMalloc(&h->d_mhi, sizeof(double) * (size_t)(T * nn)));
    RPF_CUDA(h, cudaMalloc(&h->d_perm, sizeof(uint32_t) * (size_t)std::max<int64_t>(T * n, 1)));
    RPF_CUDA(h, cudaMemsetAsync(h->d_thr, 0, sizeof(double) * (size_t)(T * nn), h->stream));
    RPF_CUDA(h, cudaMemsetAsync(h->d_mlo, 0, sizeof(double) * (size_t)(T * nn), h->stream));
    RPF_CUDA(h, cudaMemsetAsync(h->d_mhi, 0, sizeof(double) * (size_t)(T * nn), h->stream));

    // ---- phase split: first level whose nodes all fit the shared-memory capacity
    int s = tp.nlevels;
    for (int l = 0; l < tp.nlevels; ++l) if ((int64_t)tp.lvl_maxsize[l] <= CAP) { s = l; break; }
    const int s_top = std::min(s, L);                       // levels 0..s_top-1 are split by the top phase
    const int64_t NTOP = tp.level_off[std::min(s_top + 1, tp.nlevels)];
    if (s_top > 0 && NTOP > 65535) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "too many top-phase nodes for 16-bit labels");
    if (s >= tp.nlevels) h->leaf_order_exact = false;     // leaves larger than the capacity: membership exact, order not
    for (int l = 0; l < s && l < tp.nlevels; ++l)
        for (int64_t g = tp.level_off[l]; g < tp.level_off[l + 1]; ++g) if (tp.child[g] < 0) h->leaf_order_exact = false;

    if (L == 0 || n == 0) {   // every tree is a single Tip holding the points in input order
        if (n > 0) {
            const int64_t tot = n * T;
            RPF_LAUNCH(h, PH_MISC, k_iota_perm, (unsigned)((tot + 255) / 256), 256, 0, h->d_perm, n, T);
        }
        return RPF_OK;
    }

    // ---- per-level histogram geometry for the top phase
    std::vector<int> nb_level(std::max(L, 1), 0), smem_level(std::max(L, 1), 0);
    int64_t HSZ = 1;
    for (int l = 0; l < s_top; ++l) {
        const int nodes = (int)(tp.level_off[l + 1] - tp.level_off[l]);
        int nb;
        if (nodes <= 128) { nb = HBINS / (int)next_pow2_host((unsigned)nodes); smem_level[l] = 1; }
        else { nb = 256; smem_level[l] = 0; }
        nb_level[l] = nb;
        HSZ = std::max<int64_t>(HSZ, (int64_t)nodes * nb);
    }
    const int MAXTD = L + 1;

    // ---- tree group size from free memory
    size_t freeB = 0, totalB = 0;
    RPF_CUDA(h, cudaMemGetInfo(&freeB, &totalB));
    const size_t per_tree = (size_t)L * n * 8 + (s_top > 0 ? (size_t)n * 10 : 0) + (size_t)HSZ * 4 +
                            (size_t)NTOP * (sizeof(NodeSel) + 4 + (size_t)MAXTD * 8) + 4096;
    int Tg = (int)std::min<size_t>((size_t)T, std::max<size_t>(1, (size_t)(freeB * 0.7) / per_tree));
    if ((size_t)per_tree > freeB) return rpf_fail(h, RPF_ERR_NOMEM, "not enough device memory for one tree's keys");

    DevBuf keys, label, hist, sel, cand, cand_total, pivots, fill, kmin, kmax, binlo, binscale, nbdev, range, lvlpv;
    RPF_CUDA(h, keys.alloc((size_t)Tg * L * n * 8));
    RPF_CUDA(h, kmin.alloc((size_t)Tg * L * 8));
    RPF_CUDA(h, kmax.alloc((size_t)Tg * L * 8));
    if (s_top > 0) {
        RPF_CUDA(h, label.alloc((size_t)Tg * n * 2));
        RPF_CUDA(h, hist.alloc((size_t)Tg * HSZ * 4));
        RPF_CUDA(h, sel.alloc((size_t)Tg * NTOP * sizeof(NodeSel)));
        RPF_CUDA(h, cand.alloc((size_t)Tg * n * 8));
        RPF_CUDA(h, cand_total.alloc((size_t)Tg * 4));
        RPF_CUDA(h, pivots.alloc((size_t)Tg * NTOP * MAXTD * 8));
        RPF_CUDA(h, fill.alloc((size_t)Tg * NTOP * 4));
        RPF_CUDA(h, binlo.alloc((size_t)Tg * L * 8));
        RPF_CUDA(h, binscale.alloc((size_t)Tg * L * 8));
        RPF_CUDA(h, nbdev.alloc((size_t)L * 4));
        RPF_CUDA(h, cudaMemcpyAsync(nbdev.p, nb_level.data(), (size_t)L * 4, cudaMemcpyHostToDevice, h->stream));
    }

    // ---- bottom-phase tables: per level-s node, the BFS id range of its descendants at each deeper level
    int nnodes_s = 0, nlb = 0;
    if (s < tp.nlevels) {
        nnodes_s = (int)(tp.level_off[s + 1] - tp.level_off[s]);
        nlb = std::max(1, tp.nlevels - s);
        std::vector<int2> rg((size_t)nnodes_s * nlb, make_int2(0, 0));
        for (int e = 0; e < nnodes_s; ++e) {
            int64_t lo = tp.level_off[s] + e, hi = lo + 1;
            for (int j = 0; j < nlb; ++j) {
                int64_t fi = -1, li = -1;
                for (int64_t g = lo; g < hi; ++g) if (tp.child[g] >= 0) { if (fi < 0) fi = g; li = g; }
                if (fi < 0) break;
                rg[(size_t)e * nlb + j] = make_int2((int)lo, (int)hi);
                lo = tp.child[fi]; hi = (int64_t)tp.child[li] + 2;
            }
        }
        std::vector<uint32_t> pv(tp.nlevels);
        for (int l = 0; l < tp.nlevels; ++l) pv[l] = next_pow2_host(std::max<uint32_t>(tp.lvl_maxsize[l], 1));
        RPF_CUDA(h, range.alloc(rg.size() * sizeof(int2)));
        RPF_CUDA(h, lvlpv.alloc(pv.size() * 4));
        RPF_CUDA(h, cudaMemcpyAsync(range.p, rg.data(), rg.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
        RPF_CUDA(h, cudaMemcpyAsync(lvlpv.p, pv.data(), pv.size() * 4, cudaMemcpyHostToDevice, h->stream));
        RPF_CUDA(h, cudaStreamSynchronize(h->stream));   // host vectors go out of scope below
    }

    for (int t0 = 0; t0 < T; t0 += Tg) {
        const int tg = std::min(Tg, T - t0);
        // K1
        RPF_CUDA(h, cudaMemsetAsync(kmin.p, 0xff, (size_t)tg * L * 8, h->stream));
        RPF_CUDA(h, cudaMemsetAsync(kmax.p, 0x00, (size_t)tg * L * 8, h->stream));
        int rc = rpf_project_launch(h, PH_PROJECT, h->dX, n, t0, tg, L, true, keys.p, kmin.as<ull>(), kmax.as<ull>());
        if (rc) return rc;

        uint32_t* perm_g = h->d_perm + (int64_t)t0 * n;
        if (s_top > 0) {
            TopArgs A{};
            A.n = n; A.Tg = tg; A.L = L; A.NTOP = (int)NTOP; A.HSZ = (int)HSZ; A.MAXTD = MAXTD; A.gt0 = t0; A.nn_all = nn;
            A.keys = keys.as<ull>(); A.label = label.as<uint16_t>(); A.child = h->d_node_child; A.nstart = h->d_node_start;
            A.nsize = h->d_node_size; A.binlo = binlo.as<double>(); A.binscale = binscale.as<double>();
            A.kmin = kmin.as<ull>(); A.kmax = kmax.as<ull>(); A.hist = hist.as<uint32_t>(); A.sel = sel.as<NodeSel>();
            A.cand = cand.as<ull>(); A.cand_total = cand_total.as<uint32_t>(); A.pivots = pivots.as<ull>();
            A.fill = fill.as<uint32_t>(); A.perm = perm_g; A.thr = h->d_thr; A.mlo = h->d_mlo; A.mhi = h->d_mhi;
            RPF_LAUNCH(h, PH_MISC, k_bin_setup, (unsigned)((tg * L + 127) / 128), 128, 0, A, nbdev.as<int>(), s_top);
            RPF_CUDA(h, cudaMemsetAsync(fill.p, 0, (size_t)tg * NTOP * 4, h->stream));
            const unsigned nchunks = (unsigned)((n + TOP_CH - 1) / TOP_CH);
            for (int l = 0; l < s_top; ++l) {
                A.l = l; A.node0 = (int)tp.level_off[l]; A.nnodes = (int)(tp.level_off[l + 1] - tp.level_off[l]);
                A.NB = nb_level[l]; A.smem_hist = smem_level[l];
                RPF_CUDA(h, cudaMemsetAsync(hist.p, 0, (size_t)tg * HSZ * 4, h->stream));
                RPF_CUDA(h, cudaMemsetAsync(cand_total.p, 0, (size_t)tg * 4, h->stream));
                dim3 gs(nchunks, (unsigned)tg), gn((unsigned)A.nnodes, (unsigned)tg);
                const size_t hs = A.smem_hist ? (size_t)A.nnodes * A.NB * 4 : 0;
                RPF_LAUNCH(h, PH_TOP_HIST, k_top_hist, gs, TOP_NT, hs, A);
                RPF_LAUNCH(h, PH_TOP_PICK, k_top_pick, gn, 256, 0, A);
                RPF_LAUNCH(h, PH_TOP_COMPACT, k_top_compact, gs, TOP_NT, 0, A);
                RPF_LAUNCH(h, PH_TOP_FINISH, k_top_finish, gn, 512, 0, A);
                RPF_LAUNCH(h, PH_TOP_TIES, k_top_ties, gn, 512, 0, A);
                RPF_LAUNCH(h, PH_TOP_RELABEL, k_top_relabel, gs, TOP_NT, 0, A, (int)(l == s_top - 1));
                RPF_LAUNCH(h, PH_MISC, k_top_finalize, (unsigned)((A.nnodes * tg + 127) / 128), 128, 0, A);
            }
        } else {
            const int64_t tot = n * tg;
            RPF_LAUNCH(h, PH_MISC, k_iota_perm, (unsigned)((tot + 255) / 256), 256, 0, perm_g, n, tg);
        }

        if (s < tp.nlevels) {
            BottomArgs B{};
            B.n = n; B.nn_all = nn; B.L = L; B.s = s; B.nlb = nlb; B.gt0 = t0; B.first_gid = (int)tp.level_off[s];
            B.keys = keys.as<ull>(); B.perm = perm_g; B.child = h->d_node_child; B.nstart = h->d_node_start; B.nsize = h->d_node_size;
            B.range = range.as<int2>(); B.lvl_pv = lvlpv.as<uint32_t>(); B.thr = h->d_thr; B.mlo = h->d_mlo; B.mhi = h->d_mhi;
            int rc2;
            switch (CAP) {
                case 256: rc2 = launch_bottom<256, 128>(h, B, nnodes_s, tg); break;
                case 1024: rc2 = launch_bottom<1024, 256>(h, B, nnodes_s, tg); break;
                case 4096: rc2 = launch_bottom<4096, 512>(h, B, nnodes_s, tg); break;
                case 8192: rc2 = launch_bottom<8192, 1024>(h, B, nnodes_s, tg); break;
                default: return rpf_fail(h, RPF_ERR_ARG, "bottom_cap must be 256, 1024, 4096 or 8192");
            }
            if (rc2) return rc2;
        }
    }
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));   // group buffers are freed on return
    return RPF_OK;
}

// projections of a query batch onto every (tree, level) hyperplane: keysQ[(t*L + l) * nq + q]
int rpf_project_queries(rpf_handle* h, const double* dQ, int64_t nq, double* d_keysQ) {
    return rpf_project_launch(h, PH_Q_PROJECT, dQ, nq, 0, h->T, h->topo.L_eff, false, d_keysQ, nullptr, nullptr);
}
