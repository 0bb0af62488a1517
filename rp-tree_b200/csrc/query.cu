// query.cu -- candidates / knn / knnPQ / recallWith on sm_100a.
//
// Replaces, for dense Double data and distf = metricL2, the reference's
//   candidates   src/Data/RPTree.hs:289-314   (margin-aware descent, may fork into both children)
//   knn          src/Data/RPTree.hs:168-176   (all candidates of all trees, duplicates kept, stable sort, take k)
//   knnPQ        src/Data/RPTree.hs:181-194,224-227 (one result per distinct distance)
//   recallWith   src/Data/RPTree.hs:259-282   (mean over trees of |candidates /\ true top-k| / k)
//   metricDDL2   src/Data/RPTree/Internal.hs:403-406
//
// Queries are projected onto every (tree, level) hyperplane with the same exact-order kernel as the build
// (k_project), so the descent is a pure table walk over the 24-byte node records (L2 resident).  The re-rank is
// HBM bound: one thread per candidate streams the candidate's row with 256-bit loads and accumulates
// sum (x-q)^2 strictly left to right (no FMA) so distances, hence orderings, match the reference.
#include "rpf_internal.h"
#include "rpf_device.cuh"
#include <algorithm>
#include <vector>

int rpf_project_launch(rpf_handle* h, int phase, const double* dX, int64_t n, int t0, int Tg, int L, bool ord, void* out,
                       int64_t ostride, ull* kmin, ull* kmax);

__device__ __forceinline__ int q_ilog2(unsigned v) { return ilog2_pow2(v); }
__device__ __forceinline__ unsigned q_next_pow2(unsigned v) { return next_pow2_u32(v); }

// ---------------------------------------------------------------------------------------------------
// descent
// ---------------------------------------------------------------------------------------------------
// One thread per (query, tree).  Emits the BFS ids of the reached leaves, left to right, into
// segs[(q*Tq + tt)*S ..]; cnt holds the true count (may exceed S -> caller retries with a larger S).
// PRIO (candidatesH, RPTree.hs:318-341): every reached leaf also gets its search priority = the smallest margin distance
// met on the way down (p starts at +inf; going left p = min p dl, going right p = min p dr; Haskell's min x y = if x <= y
// then x else y).
template <bool PRIO>
__global__ void k_traverse(const double* __restrict__ keysQ, int64_t nq, int T, int L, int64_t nn,
                           const int32_t* __restrict__ child, const int32_t* __restrict__ depth,
                           const double* __restrict__ thr, const double* __restrict__ mlo, const double* __restrict__ mhi,
                           int S, int t_only, uint32_t* __restrict__ segs, uint32_t* __restrict__ cnt, uint32_t* __restrict__ maxcnt,
                           double* __restrict__ prio) {
    const int Tq = t_only >= 0 ? 1 : T;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nq * Tq) return;
    const int64_t q = idx % nq;
    const int tt = (int)(idx / nq);
    const int t = t_only >= 0 ? t_only : tt;
    const double* thr_t = thr + (int64_t)t * nn;
    const double* mlo_t = mlo + (int64_t)t * nn;
    const double* mhi_t = mhi + (int64_t)t * nn;
    const double* kq = keysQ + (int64_t)t * L * nq + q;
    int stk[64];
    double pstk[PRIO ? 64 : 1];
    int sp = 0, g = 0;
    uint32_t c = 0;
    double p = __longlong_as_double(0x7ff0000000000000LL);
    uint32_t* out = segs + (q * Tq + tt) * (int64_t)S;
    double* pout = PRIO ? prio + (q * Tq + tt) * (int64_t)S : nullptr;
    while (true) {
        int ch;
        while ((ch = __ldg(child + g)) >= 0) {
            const int lev = __ldg(depth + g);
            const double proj = kq[(int64_t)lev * nq];
            const double th = thr_t[g], lo = mlo_t[g], hi = mhi_t[g];
            const double dl = fabs(__dsub_rn(lo, proj)), dr = fabs(__dsub_rn(hi, proj));
            const double pl = (p <= dl) ? p : dl, pr = (p <= dr) ? p : dr;
            if (proj < th && dl > dr) { if (PRIO) pstk[sp] = pr; stk[sp++] = ch + 1; g = ch; p = pl; }
            else if (proj < th) { g = ch; p = pl; }
            else if (proj > th && dl < dr) { if (PRIO) pstk[sp] = pr; stk[sp++] = ch + 1; g = ch; p = pl; }
            else { g = ch + 1; p = pr; }
        }
        if (c < (uint32_t)S) { out[c] = (uint32_t)g; if (PRIO) pout[c] = p; }
        ++c;
        if (sp == 0) break;
        g = stk[--sp];
        if (PRIO) p = pstk[sp];
    }
    cnt[q * Tq + tt] = c;
    if (c > (uint32_t)S) atomicMax(maxcnt, c);
}

// ---------------------------------------------------------------------------------------------------
// exact distance
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// four 256-bit loads in ONE asm statement: the compiler cannot interleave their uses, so every thread keeps
// 128 bytes in flight (the kernel is DRAM-latency bound otherwise)
__device__ __forceinline__ void ld1024(const double* p, double* x) {
    asm volatile(
        "ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%16];\n\t"
        "ld.global.nc.v4.f64 {%4,%5,%6,%7}, [%16+32];\n\t"
        "ld.global.nc.v4.f64 {%8,%9,%10,%11}, [%16+64];\n\t"
        "ld.global.nc.v4.f64 {%12,%13,%14,%15}, [%16+96];"
        : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]), "=d"(x[4]), "=d"(x[5]), "=d"(x[6]), "=d"(x[7]),
          "=d"(x[8]), "=d"(x[9]), "=d"(x[10]), "=d"(x[11]), "=d"(x[12]), "=d"(x[13]), "=d"(x[14]), "=d"(x[15])
        : "l"(p));
}
// sqrt (sum_j (x_j - q_j)^2), left fold from 0, separate roundings (Internal.hs:403-406).
// `d` may be a prefix length (SVector data: the reference's diffSD / diffSS stop at the sparse operand's last component).
__device__ __forceinline__ double dist_exact(const double* __restrict__ row, const double* __restrict__ sq, int d, bool vec) {
    double acc = 0.0;
    if (vec) {
        int j = 0;
        if (d >= 16) {
            double x[16], y[16];
            ld1024(row, x);
            for (; j + 32 <= d; j += 16) {          // software pipeline: next 128 B requested before this batch is consumed
                ld1024(row + j + 16, y);
#pragma unroll
                for (int u = 0; u < 16; ++u) { const double df = __dsub_rn(x[u], sq[j + u]); acc = __dadd_rn(acc, __dmul_rn(df, df)); }
#pragma unroll
                for (int u = 0; u < 16; ++u) x[u] = y[u];
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) { const double df = __dsub_rn(x[u], sq[j + u]); acc = __dadd_rn(acc, __dmul_rn(df, df)); }
            j += 16;
        }
        for (; j < d; j += 4) {
            double x[4];
            ld256(row + j, x[0], x[1], x[2], x[3]);
#pragma unroll
            for (int u = 0; u < 4; ++u) { const double df = __dsub_rn(x[u], sq[j + u]); acc = __dadd_rn(acc, __dmul_rn(df, df)); }
        }
    } else {
        for (int j = 0; j < d; ++j) {
            const double df = __dsub_rn(__ldg(row + j), sq[j]);
            acc = __dadd_rn(acc, __dmul_rn(df, df));
        }
    }
    return __dsqrt_rn(acc);
}

// ---------------------------------------------------------------------------------------------------
// (key, pos, id) sort + keep-k in shared memory.  Keys are the raw bits of non-negative doubles (order
// preserving); pos is the candidate's position in the reference's concatenation order, which makes the
// bitonic network reproduce the reference's STABLE sort (RPTree.hs:174).
// ---------------------------------------------------------------------------------------------------
#define KNN_NT 256
#define KNN_BUF 2304   /* candidate entries per chunk (incl. the running best) */
#define KNN_SREG 768   /* side region for the selected front */

template <int NT>
__device__ void sort3(ull* skey, uint32_t* spos, uint32_t* sid, unsigned m) {
    const unsigned Pv = q_next_pow2(m), half = Pv >> 1;
    for (unsigned k = 2; k <= Pv; k <<= 1) {
        const int lk = q_ilog2(k);
        for (unsigned c = threadIdx.x; c < half; c += NT) {
            const unsigned blk = c >> (lk - 1), w = c & ((k >> 1) - 1);
            const unsigned i = (blk << lk) + w, p = (blk << lk) + (k - 1 - w);
            if (p < m) {
                ull a = skey[i], b = skey[p]; uint32_t pa = spos[i], pb = spos[p];
                if (a > b || (a == b && pa > pb)) { skey[i] = b; skey[p] = a; spos[i] = pb; spos[p] = pa; uint32_t x = sid[i]; sid[i] = sid[p]; sid[p] = x; }
            }
        }
        __syncthreads();
        for (unsigned j = k >> 2; j > 0; j >>= 1) {
            const int lj = q_ilog2(j);
            for (unsigned c = threadIdx.x; c < half; c += NT) {
                const unsigned i = ((c >> lj) << (lj + 1)) + (c & (j - 1)), p = i + j;
                if (p < m) {
                    ull a = skey[i], b = skey[p]; uint32_t pa = spos[i], pb = spos[p];
                    if (a > b || (a == b && pa > pb)) { skey[i] = b; skey[p] = a; spos[i] = pb; spos[p] = pa; uint32_t x = sid[i]; sid[i] = sid[p]; sid[p] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// after sort3: keep the first k entries (dedup: first of every run of equal keys). Returns kept count.
template <int NT>
__device__ unsigned keep_k(ull* skey, uint32_t* spos, uint32_t* sid, unsigned tot, unsigned k, int dedup, unsigned* s_n) {
    if (!dedup) return tot < k ? tot : k;
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        unsigned w = 0;
        for (unsigned base = 0; base < tot && w < k; base += 32) {
            const unsigned i = base + lane;
            ull kv = 0; uint32_t pv = 0, iv = 0; bool keep = false;
            if (i < tot) { kv = skey[i]; pv = spos[i]; iv = sid[i]; keep = (i == 0) || (skey[i - 1] != kv); }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            const unsigned dst = w + __popc(bal & ((1u << lane) - 1));
            __syncwarp();
            if (keep && dst < k) { skey[dst] = kv; spos[dst] = pv; sid[dst] = iv; }
            __syncwarp();
            w += __popc(bal);
        }
        if (lane == 0) *s_n = w < k ? w : k;
    }
    __syncthreads();
    return *s_n;
}

// Front selection: instead of sorting all `tot` entries, radix-select the value of rank r-1 (r = k, or 32k when
// de-duplicating), compact the entries <= it into a side region, sort only those and keep k.  Falls back to the
// full sort whenever that is not provably sufficient.  On exit entries [0, ret) of (skey,spos,sid) hold the result.
template <int NT, int SREG = KNN_SREG>
__device__ unsigned topk_front(ull* skey, uint32_t* spos, uint32_t* sid, unsigned tot, unsigned k, int dedup,
                               ull* rkey, uint32_t* rpos, uint32_t* rid, uint32_t* sh, ull* sh64, unsigned* s_n) {
    if (tot <= 32) {
        // one warp, no block barriers: rank every entry by (key, pos) with shuffles and write it to its rank
        if (threadIdx.x < 32) {
            const unsigned lane = threadIdx.x;
            const bool have = lane < tot;
            const ull kv = have ? skey[lane] : ~0ull;
            const uint32_t pv = have ? spos[lane] : 0xffffffffu, iv = have ? sid[lane] : 0u;
            unsigned rank = 0;
            for (unsigned o = 0; o < tot; ++o) {
                const ull ko = __shfl_sync(0xffffffffu, kv, o);
                const uint32_t po = __shfl_sync(0xffffffffu, pv, o);
                rank += (ko < kv || (ko == kv && po < pv)) ? 1u : 0u;
            }
            __syncwarp();
            if (have) { skey[rank] = kv; spos[rank] = pv; sid[rank] = iv; }
        }
        __syncthreads();
        const unsigned nb = keep_k<NT>(skey, spos, sid, tot, k, dedup, s_n);
        __syncthreads();
        return nb;
    }
    unsigned r = dedup ? min(tot, 32u * k) : min(tot, k);
    bool full = (r >= tot) || (r > SREG / 2) || tot <= 128;
    unsigned cnt = 0;
    if (!full) {
        uint32_t cl, ce;
        const ull v = cta_radix_select<NT>(tot, r - 1, [&](uint32_t i) { return skey[i]; }, sh, sh64, cl, ce);
        cnt = cl + ce;
        full = cnt > SREG;
        if (!full) {
            if (threadIdx.x == 0) *s_n = 0;
            __syncthreads();
            for (unsigned i = threadIdx.x; i < tot; i += NT) {
                const ull kv = skey[i];
                if (kv <= v) { const unsigned p = atomicAdd(s_n, 1u); rkey[p] = kv; rpos[p] = spos[i]; rid[p] = sid[i]; }
            }
            __syncthreads();
            sort3<NT>(rkey, rpos, rid, cnt);
            const unsigned nb = keep_k<NT>(rkey, rpos, rid, cnt, k, dedup, s_n);
            __syncthreads();
            if (dedup && nb < k && cnt < tot) full = true;      // not enough distinct distances in the front
            else {
                for (unsigned i = threadIdx.x; i < nb; i += NT) { skey[i] = rkey[i]; spos[i] = rpos[i]; sid[i] = rid[i]; }
                __syncthreads();
                return nb;
            }
        }
    }
    sort3<NT>(skey, spos, sid, tot);
    const unsigned nb = keep_k<NT>(skey, spos, sid, tot, k, dedup, s_n);
    __syncthreads();
    return nb;
}

// exclusive prefix of slot sizes (nslots entries) -> pre[0..nslots]; all threads participate
template <int NT>
__device__ void slot_prefix(uint32_t* pre, unsigned nslots, uint32_t* part /*NT+1*/) {
    const unsigned per = (nslots + NT - 1) / NT, b0 = threadIdx.x * per, b1 = min(nslots, b0 + per);
    uint32_t s = 0;
    for (unsigned i = b0; i < b1; ++i) s += pre[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t c = 0; for (int j = 0; j < NT; ++j) { uint32_t v = part[j]; part[j] = c; c += v; } part[NT] = c; }
    __syncthreads();
    uint32_t c = part[threadIdx.x];
    for (unsigned i = b0; i < b1; ++i) { uint32_t v = pre[i]; pre[i] = c; c += v; }
    if (threadIdx.x == 0) pre[nslots] = part[NT];
    __syncthreads();
}

struct QArgs {
    int64_t n, nq, nn;
    int d, T, S, k, dedup, vec;
    const double* X;
    const double* Q;
    const uint32_t* perm;
    const uint32_t* nstart;
    const uint32_t* nsize;
    const uint32_t* segs;
    const uint32_t* cnt;
    const uint32_t* order;     // knn only: CTA b answers query order[b] (queries grouped by their tree-0 leaf, see k_qorder_*)
    const int32_t* xlast;      // SVector data only: index of each row's last stored component (-1: none); NULL for DVector data
    const int32_t* qlast;      // SVector queries only: same per query; NULL for DVector queries
    double* dist;
    uint32_t* ids;
    int32_t* count;
    const float* X32;          // k_knn_f32: fp32 image of X (row i at X32 + i * d)
    const double* xmax;        // k_knn_f32: [1] largest row norm of X
    uint8_t* fb;               // k_knn_f32: [nq] set when the query must be answered by the exact kernel
    const uint8_t* only;       // k_knn_tma as the second pass: answer only the queries whose flag is set
};

// number of leading components the metric visits for data row `id` and query q:
//   DVector data: all d (metricDDL2, Internal.hs:403-406)
//   SVector data, DVector query: up to the row's last stored component (metricSDL2 / binSDD, Internal.hs:396-400,455-470)
//   SVector data, SVector query: up to the smaller of the two last components (metricSSL2 / binSS, Internal.hs:389-393,432-453)
__device__ __forceinline__ int metric_len(const QArgs& A, uint32_t id, int64_t q) {
    if (!A.xlast) return A.d;
    int m = __ldg(A.xlast + id);
    if (A.qlast) m = min(m, __ldg(A.qlast + q));
    return m + 1;
}

// slot sizes for query q into pre[0..nslots): slot = tt*S + j
__device__ __forceinline__ void load_slots(const QArgs& A, int64_t q, int Tq, uint32_t* pre, int NT) {
    const unsigned nslots = (unsigned)Tq * A.S;
    for (unsigned s = threadIdx.x; s < nslots; s += NT) {
        const int tt = s / A.S, j = s % A.S;
        const uint32_t c = A.cnt[q * Tq + tt];
        pre[s] = (uint32_t)j < c ? A.nsize[A.segs[(q * Tq + tt) * (int64_t)A.S + j]] : 0u;
    }
    __syncthreads();
}
__device__ __forceinline__ unsigned find_slot(const uint32_t* pre, unsigned nslots, uint32_t c) {
    unsigned lo = 0, hi = nslots;     // last slot with pre[slot] <= c and non-empty
    while (hi - lo > 1) { unsigned mid = (lo + hi) >> 1; if (pre[mid] <= c) lo = mid; else hi = mid; }
    return lo;
}

// One CTA per query: distances of every candidate (re-rank) + stable top-k.
__global__ void __launch_bounds__(KNN_NT) k_knn(QArgs A) {
    __shared__ uint32_t part[KNN_NT + 1];
    __shared__ unsigned s_n;
    __shared__ uint32_t sh[264];
    __shared__ ull sh64;
    extern __shared__ unsigned char dyn[];
    ull* skey = (ull*)dyn;
    ull* rkey = skey + KNN_BUF;
    uint32_t* spos = (uint32_t*)(rkey + KNN_SREG);
    uint32_t* sid = spos + KNN_BUF;
    uint32_t* rpos = sid + KNN_BUF;
    uint32_t* rid = rpos + KNN_SREG;
    double* sq = (double*)(rid + KNN_SREG);
    uint32_t* pre = (uint32_t*)(sq + ((A.d + 3) & ~3));
    const int64_t q = A.order ? (int64_t)A.order[blockIdx.x] : (int64_t)blockIdx.x;
    const int tid = threadIdx.x;
    const unsigned nslots = (unsigned)A.T * A.S;
    for (int j = tid; j < A.d; j += KNN_NT) sq[j] = A.Q[q * A.d + j];
    load_slots(A, q, A.T, pre, KNN_NT);
    slot_prefix<KNN_NT>(pre, nslots, part);
    const uint32_t C = pre[nslots];
    const unsigned k = (unsigned)A.k, CHK = KNN_BUF - k;
    unsigned nbest = 0;
    for (uint32_t base = 0; base < C; base += CHK) {
        const unsigned m = min((uint32_t)CHK, C - base);
        for (unsigned j = tid; j < m; j += KNN_NT) {
            const uint32_t c = base + j;
            const unsigned slot = find_slot(pre, nslots, c);
            const int tt = slot / A.S;
            const uint32_t g = A.segs[(q * A.T + tt) * (int64_t)A.S + (slot % A.S)];
            const uint32_t id = A.perm[(int64_t)tt * A.n + A.nstart[g] + (c - pre[slot])];
            const int len = metric_len(A, id, q);
            const double dist = dist_exact(A.X + (int64_t)id * A.d, sq, len, A.vec && len == A.d);
            skey[nbest + j] = (ull)__double_as_longlong(dist);
            spos[nbest + j] = c;
            sid[nbest + j] = id;
        }
        __syncthreads();
        const unsigned tot = nbest + m;
        nbest = topk_front<KNN_NT>(skey, spos, sid, tot, k, A.dedup, rkey, rpos, rid, sh, &sh64, &s_n);
    }
    for (unsigned i = tid; i < k; i += KNN_NT) {
        const bool ok = i < nbest;
        A.dist[q * k + i] = ok ? __longlong_as_double((long long)skey[i]) : __longlong_as_double(0x7ff0000000000000LL);
        A.ids[q * k + i] = ok ? sid[i] : 0xffffffffu;
    }
    if (tid == 0 && A.count) A.count[q] = (int32_t)nbest;
}

// ---------------------------------------------------------------------------------------------------
// knn, B200 path: candidate rows are gathered by the TMA engine (cp.async.bulk, one 8d-byte bulk copy per row)
// into a 2-stage shared-memory ring guarded by mbarriers; a producer warp resolves candidate -> row id and issues
// the copies, one consumer warp per stage folds each staged row into its exact distance (one lane per row, strictly
// left to right).  Per-thread 32-byte gathers are limited by the number of outstanding L1 requests (~1.7 TB/s
// measured); bulk copies move whole rows and bypass that limit.
// ---------------------------------------------------------------------------------------------------
#define KT_NT 256
#ifndef KT_STAGES
#define KT_STAGES 2
#endif
#ifndef KT_NPROD
#define KT_NPROD 4     /* producer warps (must divide 32) */
#endif
#define KT_BUF 1280      /* candidate entries per chunk (incl. the running best) */
#define KT_SREG 512

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// sum_j (x_j - q_j)^2 of one staged row, strictly left to right with separate roundings (Internal.hs:403-406).
// Software pipelined: the loads, differences and squares of the NEXT eight components are issued before the eight
// dependent adds of the current ones, so the only loop-carried latency is one DADD per component.
__device__ __forceinline__ double row_dist2(const double2* __restrict__ row, const double2* __restrict__ q2, int d2) {
    double acc = 0.0;
    int jj = 0;
    const int nblk = d2 >> 2;
    if (nblk > 0) {
        double2 x0 = row[0], x1 = row[1], x2 = row[2], x3 = row[3];
        double2 y0 = q2[0], y1 = q2[1], y2 = q2[2], y3 = q2[3];
        for (int b = 0; b < nblk; ++b) {
            const int nx = (b + 1 < nblk) ? 4 * (b + 1) : 4 * b;      // last block re-reads itself (values unused)
            const double2 nx0 = row[nx], nx1 = row[nx + 1], nx2 = row[nx + 2], nx3 = row[nx + 3];
            const double2 ny0 = q2[nx], ny1 = q2[nx + 1], ny2 = q2[nx + 2], ny3 = q2[nx + 3];
            const double a0 = __dsub_rn(x0.x, y0.x), a1 = __dsub_rn(x0.y, y0.y), a2 = __dsub_rn(x1.x, y1.x), a3 = __dsub_rn(x1.y, y1.y);
            const double a4 = __dsub_rn(x2.x, y2.x), a5 = __dsub_rn(x2.y, y2.y), a6 = __dsub_rn(x3.x, y3.x), a7 = __dsub_rn(x3.y, y3.y);
            const double s0 = __dmul_rn(a0, a0), s1 = __dmul_rn(a1, a1), s2 = __dmul_rn(a2, a2), s3 = __dmul_rn(a3, a3);
            const double s4 = __dmul_rn(a4, a4), s5 = __dmul_rn(a5, a5), s6 = __dmul_rn(a6, a6), s7 = __dmul_rn(a7, a7);
            acc = __dadd_rn(acc, s0); acc = __dadd_rn(acc, s1); acc = __dadd_rn(acc, s2); acc = __dadd_rn(acc, s3);
            acc = __dadd_rn(acc, s4); acc = __dadd_rn(acc, s5); acc = __dadd_rn(acc, s6); acc = __dadd_rn(acc, s7);
            x0 = nx0; x1 = nx1; x2 = nx2; x3 = nx3; y0 = ny0; y1 = ny1; y2 = ny2; y3 = ny3;
        }
        jj = nblk * 4;
    }
    for (; jj < d2; ++jj) {
        const double2 x = row[jj], y = q2[jj];
        const double d0 = __dsub_rn(x.x, y.x), d1 = __dsub_rn(x.y, y.y);
        acc = __dadd_rn(acc, __dmul_rn(d0, d0));
        acc = __dadd_rn(acc, __dmul_rn(d1, d1));
    }
    return acc;
}

#define KT_INF_BITS 0x7ff0000000000000ull

// Roles per chunk of CH = KT_BUF - k candidates: warp 0 = TMA producer, warps 1..KT_STAGES = consumers (one per ring
// stage), the remaining warps resolve candidate -> row id for the NEXT chunk (two dependent global reads, segs ->
// perm) while the ring runs.  Consumers keep only candidates whose distance does not exceed the running k-th best
// (tau): after the first select nearly everything is filtered, so a query costs ~one radix select however many
// candidates it has.  Survivors carry their position in the reference's concatenation order; the final
// (distance, position) sort therefore equals the reference's stable sort (RPTree.hs:174).
__global__ void __launch_bounds__(KT_NT) k_knn_tma(QArgs A, int rows_per_stage, int pitch /* bytes, multiple of 16 */) {
    __shared__ uint32_t part[KT_NT + 1];
    __shared__ unsigned s_n, s_nsurv;
    __shared__ uint32_t sh[264];
    __shared__ ull sh64, s_tau;
    __shared__ __align__(8) uint64_t full_bar[KT_STAGES], empty_bar[KT_STAGES];
    extern __shared__ __align__(16) unsigned char dyn[];
    unsigned char* stage_buf = dyn;                                              // [KT_STAGES][rows][pitch]
    ull* skey = (ull*)(dyn + (size_t)KT_STAGES * rows_per_stage * pitch);
    ull* rkey = skey + KT_BUF;
    uint32_t* spos = (uint32_t*)(rkey + KT_SREG);
    uint32_t* sid = spos + KT_BUF;
    uint32_t* rpos = sid + KT_BUF;
    uint32_t* rid = rpos + KT_SREG;
    uint32_t* cid = rid + KT_SREG;                                               // [2][KT_BUF] row ids of the current / next chunk
    double* sq = (double*)(cid + 2 * KT_BUF);
    uint32_t* pre = (uint32_t*)(sq + ((A.d + 3) & ~3));
    const int64_t q = A.order ? (int64_t)A.order[blockIdx.x] : (int64_t)blockIdx.x;
    if (A.only && !A.only[q]) return;                                            // second pass of k_knn_f32: flagged queries only
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nslots = (unsigned)A.T * A.S;
    const int R = rows_per_stage;
    const uint32_t row_bytes = (uint32_t)A.d * 8u;

    if (tid == 0) {
        for (int s2 = 0; s2 < KT_STAGES; ++s2) { mbar_init(&full_bar[s2], KT_NPROD); mbar_init(&empty_bar[s2], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        s_nsurv = 0; s_tau = KT_INF_BITS;
    }
    for (int j = tid; j < A.d; j += KT_NT) sq[j] = A.Q[q * A.d + j];
    load_slots(A, q, A.T, pre, KT_NT);
    slot_prefix<KT_NT>(pre, nslots, part);
    const uint32_t C = pre[nslots];
    const unsigned k = (unsigned)A.k, CHK = KT_BUF - k;
    auto resolve = [&](uint32_t base, unsigned m, uint32_t* dst, unsigned r, unsigned nthr) {
        for (unsigned j0 = r; j0 < m; j0 += 4 * nthr) {          // four independent segs -> perm chains in flight per thread
            const uint32_t* src[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned j = j0 + u * nthr;
                src[u] = nullptr;
                if (j < m) {
                    const uint32_t c = base + j;
                    const unsigned slot = find_slot(pre, nslots, c);
                    const int tt = slot / A.S;
                    const uint32_t g = A.segs[(q * A.T + tt) * (int64_t)A.S + (slot % A.S)];
                    src[u] = A.perm + (int64_t)tt * A.n + A.nstart[g] + (c - pre[slot]);
                }
            }
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = src[u] ? __ldg(src[u]) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (src[u]) dst[j0 + u * nthr] = v[u];
        }
    };
    resolve(0, min((uint32_t)CHK, C), cid, tid, KT_NT);
    __syncthreads();
    unsigned nbest = 0;
    uint32_t uses = 0;            // tiles issued so far (all roles advance it identically)
    int cb = 0;
    for (uint32_t base = 0; base < C; base += CHK, cb ^= 1) {
        const unsigned m = min((uint32_t)CHK, C - base);
        const unsigned ntiles = (m + R - 1) / R;
        const uint32_t nbase = base + CHK;
        const unsigned next_m = nbase < C ? min((uint32_t)CHK, C - nbase) : 0u;
        const uint32_t* ids = cid + cb * KT_BUF;
        if (warp < KT_NPROD) {
            // ---- producers: warp p issues rows [p*RP, (p+1)*RP) of every tile.  cp.async.bulk takes uniform operands, so
            // the per-lane copies of one warp are issued one after the other (~60 cycles each); several producer warps
            // issue concurrently.
            constexpr unsigned RP = 32 / KT_NPROD;
            for (unsigned ti = 0; ti < ntiles; ++ti) {
                const uint32_t u = uses + ti, st = u % KT_STAGES, round = u / KT_STAGES;
                if (round > 0) mbar_wait(&empty_bar[st], (round - 1) & 1);
                const unsigned rloc = (unsigned)warp * RP + lane;                 // row inside the tile
                const unsigned j = ti * R + rloc;
                const bool valid = lane < RP && rloc < (unsigned)R && j < m;
                const uint32_t id = valid ? ids[j] : 0u;
                const unsigned nrows = min((unsigned)R, m - ti * R);
                const unsigned lo = min(nrows, (unsigned)warp * RP), hi = min(nrows, (unsigned)(warp + 1) * RP);
                if (lane == 0) {
                    if (hi > lo) mbar_arrive_expect_tx(&full_bar[st], (hi - lo) * row_bytes);
                    else mbar_arrive(&full_bar[st]);
                }
                __syncwarp();
                if (valid) bulk_g2s(stage_buf + ((size_t)st * R + rloc) * pitch, A.X + (int64_t)id * A.d, row_bytes, &full_bar[st]);
            }
        } else if (warp < KT_NPROD + KT_STAGES) {
            // ---- consumer of stage warp-KT_NPROD
            const uint32_t st = warp - KT_NPROD;
            const ull tau = s_tau;
            for (unsigned ti = 0; ti < ntiles; ++ti) {
                const uint32_t u = uses + ti;
                if (u % KT_STAGES != st) continue;
                mbar_wait(&full_bar[st], (u / KT_STAGES) & 1);
                const unsigned j = ti * R + lane;
                if (lane < R && j < m) {
                    const double acc = row_dist2((const double2*)(stage_buf + ((size_t)st * R + lane) * pitch), (const double2*)sq, A.d >> 1);
                    const ull bits = (ull)__double_as_longlong(__dsqrt_rn(acc));
                    if (bits <= tau) {
                        const unsigned p = atomicAdd(&s_nsurv, 1u);
                        skey[p] = bits; spos[p] = base + j; sid[p] = ids[j];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[st]);
            }
        } else if (next_m) {
            // ---- the other warps look up the next chunk's row ids while the ring runs
            resolve(nbase, next_m, cid + (cb ^ 1) * KT_BUF, (unsigned)tid - 32u * (KT_NPROD + KT_STAGES), (unsigned)KT_NT - 32u * (KT_NPROD + KT_STAGES));
        }
        uses += ntiles;
        __syncthreads();
        const unsigned nsurv = s_nsurv;
        if (next_m == 0 || nsurv + next_m > KT_BUF) {
            nbest = topk_front<KT_NT, KT_SREG>(skey, spos, sid, nsurv, k, A.dedup, rkey, rpos, rid, sh, &sh64, &s_n);
            if (tid == 0) { s_nsurv = nbest; s_tau = nbest >= k ? skey[k - 1] : KT_INF_BITS; }
            __syncthreads();
        }
    }
    for (unsigned i = tid; i < k; i += KT_NT) {
        const bool ok = i < nbest;
        A.dist[q * k + i] = ok ? __longlong_as_double((long long)skey[i]) : __longlong_as_double(0x7ff0000000000000LL);
        A.ids[q * k + i] = ok ? sid[i] : 0xffffffffu;
    }
    if (tid == 0 && A.count) A.count[q] = (int32_t)nbest;
}


// ---------------------------------------------------------------------------------------------------
// knn with an fp32 FILTER pass in front of the exact re-rank.
// k_knn_tma moves 8d bytes per candidate and is bound by how many rows fit the shared-memory ring (two 32-row stages per CTA:
// ~90 KB in flight per SM against ~2 us of latency = 45 GB/s per SM).  The reference's answer only needs the exact distance
// of the FEW candidates that can be among the k nearest.  So: (A) every candidate row is streamed from an fp32 image of X
// (half the bytes, twice the rows in flight, four ring stages) and gets an approximate distance d~ in fp32 arithmetic;
// with u = 2^-24, |d~ - d| <= delta d + eta for delta = 2 (d + 8) u, eta = 4 u (|q| + max|x|) + 2^-70 sqrt(d)  (rounding of x and
// q to fp32, of the differences, of the d-term sum, of the square root; underflow of squares) -- a 2x over-estimate of the
// standard bounds.  If tau~ is the k-th smallest d~ seen so far, every candidate of the exact top-k has
// d~ <= tau~ (1 + 4 delta) + 4 eta; the others are dropped.  (B) the survivors (k plus whatever lies within the margin: a
// handful) get their 8d-byte rows staged and the exact distance of metricDDL2 (Internal.hs:403-406: left to right,
// separate roundings), and the final order is the (distance, position in the reference's concatenation order) sort of
// k_knn_tma -- same ids, same distance bits, same tie order.  Queries with more than KT_SREG survivors (duplicate rows,
// integer data) or non-finite / huge norms are flagged and answered by k_knn_tma in a second launch.
// ---------------------------------------------------------------------------------------------------
#define KF_STAGES 4
#define KF_NPROD 3
__global__ void k_x32_convert(const double* __restrict__ X, int64_t n, int d, float* __restrict__ X32, ull* __restrict__ xmax_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double best = 0.0;
    for (int64_t i = w; i < n; i += nw) {
        double s = 0.0;
        for (int j = lane; j < d; j += 32) { const double x = X[i * d + j]; X32[i * d + j] = __double2float_rn(x); s += x * x; }
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        s = sqrt(s);
        if (!(s <= 1e300)) s = __longlong_as_double(0x7ff0000000000000LL);      // NaN / Inf rows: the filter is not used
        best = fmax(best, s);
    }
    if (lane == 0) atomicMax(xmax_bits, (ull)__double_as_longlong(best));
}

__global__ void __launch_bounds__(KT_NT) k_knn_f32(QArgs A, int rows_per_stage, int pitch32, int pitch64, int nb_exact, int nstages, int BUF, int SREG) {
    __shared__ uint32_t part[KT_NT + 1];
    __shared__ unsigned s_n, s_nsurv;
    __shared__ uint32_t sh[264];
    __shared__ ull sh64;
    __shared__ double s_thr, s_qn;
    __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], xbar;
    extern __shared__ __align__(16) unsigned char dyn[];
    unsigned char* stage_buf = dyn;                                              // phase A: [nstages][rows][pitch32]; phase B: [nb_exact][pitch64]
    const size_t stage_bytes = (size_t)nstages * rows_per_stage * pitch32;   // >= nb_exact * pitch64 (host)
    ull* skey = (ull*)(dyn + ((stage_bytes + 15) & ~(size_t)15));
    ull* rkey = skey + BUF;
    uint32_t* spos = (uint32_t*)(rkey + SREG);
    uint32_t* sid = spos + BUF;
    uint32_t* rpos = sid + BUF;
    uint32_t* rid = rpos + SREG;
    uint32_t* cid = rid + SREG;                                               // [2][BUF] row ids of the current / next chunk
    double* sq = (double*)(cid + 2 * BUF);
    float* sqf = (float*)(sq + ((A.d + 3) & ~3));
    uint32_t* pre = (uint32_t*)(sqf + ((A.d + 3) & ~3));
    const int64_t q = A.order ? (int64_t)A.order[blockIdx.x] : (int64_t)blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned nslots = (unsigned)A.T * A.S;
    const int R = rows_per_stage;
    const uint32_t row_bytes = (uint32_t)A.d * 4u;

    if (tid == 0) {
        for (int s2 = 0; s2 < nstages; ++s2) { mbar_init(&full_bar[s2], KF_NPROD); mbar_init(&empty_bar[s2], 1); }
        mbar_init(&xbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        s_nsurv = 0; s_thr = __longlong_as_double(0x7ff0000000000000LL);
    }
    for (int j = tid; j < A.d; j += KT_NT) { const double v = A.Q[q * A.d + j]; sq[j] = v; sqf[j] = __double2float_rn(v); }
    load_slots(A, q, A.T, pre, KT_NT);
    slot_prefix<KT_NT>(pre, nslots, part);
    if (warp == 0) {                                                             // |q|
        double s = 0.0;
        for (int j = lane; j < A.d; j += 32) s += sq[j] * sq[j];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) s_qn = sqrt(s);
    }
    __syncthreads();
    const uint32_t C = pre[nslots];
    const unsigned k = (unsigned)A.k, CHK = (unsigned)(BUF - SREG);                   // a chunk's survivors always fit next to the kept front
    const double xm = *A.xmax, qn = s_qn;
    if (!(xm < 1e18) || !(qn < 1e18)) {                                          // fp32 would overflow: exact kernel
        if (tid == 0) A.fb[q] = 1;
        return;
    }
    const double delta = 2.0 * (double)(A.d + 8) * 5.9604644775390625e-8;       // 2 (d + 8) u
    const double eta = 4.0 * 5.9604644775390625e-8 * (qn + xm) + 8.470329472543003e-22 * sqrt((double)A.d);
    auto resolve = [&](uint32_t base, unsigned m, uint32_t* dst, unsigned r, unsigned nthr) {
        for (unsigned j0 = r; j0 < m; j0 += 4 * nthr) {          // four independent segs -> perm chains in flight per thread
            const uint32_t* src[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const unsigned j = j0 + u * nthr;
                src[u] = nullptr;
                if (j < m) {
                    const uint32_t c = base + j;
                    const unsigned slot = find_slot(pre, nslots, c);
                    const int tt = slot / A.S;
                    const uint32_t g = A.segs[(q * A.T + tt) * (int64_t)A.S + (slot % A.S)];
                    src[u] = A.perm + (int64_t)tt * A.n + A.nstart[g] + (c - pre[slot]);
                }
            }
            uint32_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = src[u] ? __ldg(src[u]) : 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) if (src[u]) dst[j0 + u * nthr] = v[u];
        }
    };
    resolve(0, min((uint32_t)CHK, C), cid, tid, KT_NT);
    __syncthreads();
    // ================= phase A: approximate distances of all candidates; survivors = (fp32 bits of d~ << 32, position, row id)
    uint32_t uses = 0;
    int cb = 0;
    for (uint32_t base = 0; base < C; base += CHK, cb ^= 1) {
        const unsigned m = min((uint32_t)CHK, C - base);
        const unsigned ntiles = (m + R - 1) / R;
        const uint32_t nbase = base + CHK;
        const unsigned next_m = nbase < C ? min((uint32_t)CHK, C - nbase) : 0u;
        const uint32_t* ids = cid + cb * BUF;
        if (warp < KF_NPROD) {
            // ---- producers: warp p issues the rows p, p + NPROD, ... of every tile (lane l: row l * NPROD + p)
            for (unsigned ti = 0; ti < ntiles; ++ti) {
                const uint32_t u = uses + ti, st = u % (unsigned)nstages, round = u / (unsigned)nstages;
                if (round > 0) mbar_wait(&empty_bar[st], (round - 1) & 1);
                const unsigned nrows = min((unsigned)R, m - ti * R);
                const unsigned rloc = (unsigned)lane * KF_NPROD + (unsigned)warp;
                const bool valid = rloc < nrows;
                const uint32_t id = valid ? ids[ti * R + rloc] : 0u;
                const unsigned mine = nrows > (unsigned)warp ? (nrows - (unsigned)warp + KF_NPROD - 1) / KF_NPROD : 0u;
                if (lane == 0) {
                    if (mine) mbar_arrive_expect_tx(&full_bar[st], mine * row_bytes);
                    else mbar_arrive(&full_bar[st]);
                }
                __syncwarp();
                if (valid) bulk_g2s(stage_buf + ((size_t)st * R + rloc) * pitch32, A.X32 + (int64_t)id * A.d, row_bytes, &full_bar[st]);
            }
        } else if (warp < KF_NPROD + nstages) {
            // ---- consumer of stage warp - NPROD: one lane per staged row, fp32, four partial sums
            const uint32_t st = warp - KF_NPROD;
            const float thr = (float)s_thr;                                      // (rounded up below: s_thr holds a float value)
            const int d4 = A.d >> 2;
            for (unsigned ti = 0; ti < ntiles; ++ti) {
                const uint32_t u = uses + ti;
                if (u % (unsigned)nstages != st) continue;
                mbar_wait(&full_bar[st], (u / (unsigned)nstages) & 1);
                const unsigned j = ti * R + lane;
                if (lane < R && j < m) {
                    const float4* row = (const float4*)(stage_buf + ((size_t)st * R + lane) * pitch32);
                    const float4* qf = (const float4*)sqf;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
                    for (int jj = 0; jj < d4; ++jj) {
                        const float4 x = row[jj], y = qf[jj];
                        const float t0 = x.x - y.x, t1 = x.y - y.y, t2 = x.z - y.z, t3 = x.w - y.w;
                        a0 = fmaf(t0, t0, a0); a1 = fmaf(t1, t1, a1); a2 = fmaf(t2, t2, a2); a3 = fmaf(t3, t3, a3);
                    }
                    const float da = sqrtf((a0 + a1) + (a2 + a3));
                    if (da <= thr) {
                        const unsigned p = atomicAdd(&s_nsurv, 1u);
                        skey[p] = (ull)__float_as_uint(da) << 32; spos[p] = base + j; sid[p] = ids[j];
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[st]);
            }
        } else if (next_m) {
            resolve(nbase, next_m, cid + (cb ^ 1) * BUF, (unsigned)tid - 32u * (unsigned)(KF_NPROD + nstages), (unsigned)KT_NT - 32u * (unsigned)(KF_NPROD + nstages));
        }
        uses += ntiles;
        __syncthreads();
        const unsigned nsurv = s_nsurv;
        if (next_m == 0 || nsurv + next_m > (unsigned)BUF) {
            // tau~ = k-th smallest approximate distance so far (4 radix passes: the keys are 32-bit); keep what lies within the margin
            if (nsurv > k) {
                uint32_t cl, ce;
                const ull v = cta_radix_select<KT_NT, 4>(nsurv, k - 1, [&](uint32_t i) { return skey[i]; }, sh, &sh64, cl, ce);
                const double tau = (double)__uint_as_float((unsigned)(v >> 32));
                const float thr2 = __double2float_ru(tau * (1.0 + 4.0 * delta) + 4.0 * eta);
                if (tid == 0) s_n = 0;
                __syncthreads();
                const ull tb = (ull)__float_as_uint(thr2) << 32;
                for (unsigned i = tid; i < nsurv; i += KT_NT) {
                    const ull kv = skey[i];
                    if (kv <= tb) { const unsigned p = atomicAdd(&s_n, 1u); if (p < (unsigned)SREG) { rkey[p] = kv; rpos[p] = spos[i]; rid[p] = sid[i]; } }
                }
                __syncthreads();
                const unsigned cnt = s_n;
                if (cnt > (unsigned)SREG) {                                             // too many candidates within the margin: exact kernel
                    if (tid == 0) A.fb[q] = 1;
                    return;
                }
                for (unsigned i = tid; i < cnt; i += KT_NT) { skey[i] = rkey[i]; spos[i] = rpos[i]; sid[i] = rid[i]; }
                if (tid == 0) { s_nsurv = cnt; s_thr = (double)thr2; }
                __syncthreads();
            }
        }
    }
    // ================= phase B: exact distances of the survivors, then the reference's (distance, position) order
    const unsigned ns = s_nsurv;
    if (ns > (unsigned)SREG) { if (tid == 0) A.fb[q] = 1; return; }
    const uint32_t row_bytes64 = (uint32_t)A.d * 8u;
    uint32_t xphase = 0;
    for (unsigned b0 = 0; b0 < ns; b0 += (unsigned)nb_exact, xphase ^= 1) {
        const unsigned nb = min((unsigned)nb_exact, ns - b0);
        __syncthreads();                                                         // the buffer's previous readers are done
        if (warp == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (lane == 0) mbar_arrive_expect_tx(&xbar, nb * row_bytes64);
            __syncwarp();
            for (unsigned r = lane; r < nb; r += 32) bulk_g2s(stage_buf + (size_t)r * pitch64, A.X + (int64_t)sid[b0 + r] * A.d, row_bytes64, &xbar);
        }
        mbar_wait(&xbar, xphase);
        if ((unsigned)tid < nb) {
            const double acc = row_dist2((const double2*)(stage_buf + (size_t)tid * pitch64), (const double2*)sq, A.d >> 1);
            skey[b0 + tid] = (ull)__double_as_longlong(__dsqrt_rn(acc));
        }
    }
    __syncthreads();
    const unsigned nbest = topk_front<KT_NT, 128>(skey, spos, sid, ns, k, 0, rkey, rpos, rid, sh, &sh64, &s_n);
    for (unsigned i = tid; i < k; i += KT_NT) {
        const bool ok = i < nbest;
        A.dist[q * k + i] = ok ? __longlong_as_double((long long)skey[i]) : __longlong_as_double(0x7ff0000000000000LL);
        A.ids[q * k + i] = ok ? sid[i] : 0xffffffffu;
    }
    if (tid == 0 && A.count) A.count[q] = (int32_t)nbest;
}

// fp32 image of X + largest row norm, rebuilt when the points changed
static int ensure_x32(rpf_handle* h) {
    // (a borrowed buffer -- rpf_set_points_device -- can change under the handle: its image is rebuilt on every call)
    const bool stable = h->ownX || h->insert_session;
    if (stable && h->dX32 && h->x32_src == h->dX && h->x32_n == h->n && h->x32_d == h->d && h->x32_version == h->x_version) return RPF_OK;
    const size_t bytes = std::max<size_t>((size_t)h->n * h->d * 4, 16);
    if (!h->dX32 || h->x32_bytes < bytes) {
        if (h->dX32) { cudaStreamSynchronize(h->stream); cudaFree(h->dX32); h->dX32 = nullptr; h->x32_bytes = 0; }
        if (cudaMalloc(&h->dX32, bytes) != cudaSuccess) { cudaGetLastError(); return 1; }      // no room: the caller uses the exact kernel
        h->x32_bytes = bytes;
    }
    if (!h->d_xmax && cudaMalloc(&h->d_xmax, 16) != cudaSuccess) { cudaGetLastError(); return 1; }
    RPF_CUDA(h, cudaMemsetAsync(h->d_xmax, 0, 8, h->stream));
    RPF_LAUNCH(h, PH_Q_PROJECT, k_x32_convert, 148 * 8, 256, 0, h->dX, h->n, h->d, h->dX32, (ull*)h->d_xmax);
    h->x32_src = h->dX; h->x32_n = h->n; h->x32_d = h->d; h->x32_version = h->x_version;
    return RPF_OK;
}

// ---------------------------------------------------------------------------------------------------
// query scheduling order: CTAs that run at the same time should re-rank the same rows, so that a candidate row is
// fetched from HBM once and then served from the 126 MB L2.  Queries are grouped (counting sort) by the position of
// their first tree-0 leaf in that tree's left-to-right leaf order: neighbours in that order share the top of the
// tree, i.e. lie on the same side of the same hyperplanes.  Only the SCHEDULE changes; results are per query.
// ---------------------------------------------------------------------------------------------------
#define QO_BUCKETS 4096
__device__ __forceinline__ unsigned q_bucket(const QArgs& A, int64_t q) {
    const uint32_t g = A.segs[q * (int64_t)A.T * A.S];
    const unsigned long long st = A.nstart[g];
    return (unsigned)((st * QO_BUCKETS) / (unsigned long long)(A.n > 0 ? A.n : 1)) & (QO_BUCKETS - 1);
}
__global__ void k_qorder_hist(QArgs A, uint32_t* __restrict__ hist) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < A.nq) atomicAdd(&hist[q_bucket(A, q)], 1u);
}
__global__ void __launch_bounds__(1024) k_qorder_scan(uint32_t* __restrict__ hist) {   // exclusive scan of QO_BUCKETS counters
    __shared__ uint32_t part[1024];
    const int tid = threadIdx.x;
    uint32_t v[QO_BUCKETS / 1024], s = 0;
#pragma unroll
    for (int e = 0; e < QO_BUCKETS / 1024; ++e) { v[e] = hist[tid * (QO_BUCKETS / 1024) + e]; s += v[e]; }
    part[tid] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const uint32_t y = tid >= off ? part[tid - off] : 0u;
        __syncthreads();
        part[tid] += y;
        __syncthreads();
    }
    uint32_t c = part[tid] - s;
#pragma unroll
    for (int e = 0; e < QO_BUCKETS / 1024; ++e) { hist[tid * (QO_BUCKETS / 1024) + e] = c; c += v[e]; }
}
__global__ void k_qorder_scatter(QArgs A, uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < A.nq) order[atomicAdd(&cursor[q_bucket(A, q)], 1u)] = (uint32_t)q;
}

// candidate counts per query (sum of reached leaf sizes over Tq trees)
__global__ void k_cand_count(QArgs A, int Tq, unsigned long long* out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= A.nq) return;
    unsigned long long tot = 0;
    for (int tt = 0; tt < Tq; ++tt) {
        const uint32_t c = A.cnt[q * Tq + tt];
        for (uint32_t j = 0; j < c; ++j) tot += A.nsize[A.segs[(q * Tq + tt) * (int64_t)A.S + j]];
    }
    out[q] = tot;
}

// candidate ids per query in the reference's order (tree-major, leaves left to right, leaf order)
__global__ void __launch_bounds__(KNN_NT) k_cand_fill(QArgs A, int Tq, int t_only, const int64_t* __restrict__ off, uint32_t* __restrict__ out) {
    __shared__ uint32_t part[KNN_NT + 1];
    extern __shared__ unsigned char dyn[];
    uint32_t* pre = (uint32_t*)dyn;
    const int64_t q = blockIdx.x;
    const unsigned nslots = (unsigned)Tq * A.S;
    load_slots(A, q, Tq, pre, KNN_NT);
    slot_prefix<KNN_NT>(pre, nslots, part);
    const uint32_t C = pre[nslots];
    uint32_t* o = out + off[q];
    for (uint32_t c = threadIdx.x; c < C; c += KNN_NT) {
        const unsigned slot = find_slot(pre, nslots, c);
        const int tt = slot / A.S;
        const int t = t_only >= 0 ? t_only : tt;
        const uint32_t g = A.segs[(q * Tq + tt) * (int64_t)A.S + (slot % A.S)];
        o[c] = A.perm[(int64_t)t * A.n + A.nstart[g] + (c - pre[slot])];
    }
}

// ---------------------------------------------------------------------------------------------------
// brute force: all distances for a tile of queries, then exact per-query top-k (ties by row id)
// ---------------------------------------------------------------------------------------------------
#define BF_TQ 8
#define BF_NT 256
// D[qi][i] = raw bits of dist(X[i], Q[q0+qi]);  thread per point, query tile in shared memory
__global__ void __launch_bounds__(BF_NT) k_dist_all(const double* __restrict__ X, int64_t n, int d, const double* __restrict__ Q,
                                                    int64_t q0, int nqt, ull* __restrict__ D, int vec,
                                                    const int32_t* __restrict__ xlast, const int32_t* __restrict__ qlast) {
    extern __shared__ double sqt[];   // [BF_TQ][d]
    for (int e = threadIdx.x; e < nqt * d; e += BF_NT) sqt[e] = Q[q0 * d + e];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * BF_NT + threadIdx.x;
    if (i >= n) return;
    const double* row = X + i * d;
    double acc[BF_TQ];
#pragma unroll
    for (int qi = 0; qi < BF_TQ; ++qi) acc[qi] = 0.0;
    if (xlast) {          // SVector data: per (row, query) prefix length (see metric_len)
        const int xl = xlast[i];
        int len[BF_TQ];
#pragma unroll
        for (int qi = 0; qi < BF_TQ; ++qi) len[qi] = (qi < nqt ? (qlast ? min(xl, qlast[q0 + qi]) : xl) : -1) + 1;
        for (int j = 0; j <= xl; ++j) {
            const double x = __ldg(row + j);
#pragma unroll
            for (int qi = 0; qi < BF_TQ; ++qi)
                if (j < len[qi]) { const double df = __dsub_rn(x, sqt[qi * d + j]); acc[qi] = __dadd_rn(acc[qi], __dmul_rn(df, df)); }
        }
    } else if (vec) {
        for (int j = 0; j < d; j += 4) {
            double x[4];
            ld256(row + j, x[0], x[1], x[2], x[3]);
#pragma unroll
            for (int qi = 0; qi < BF_TQ; ++qi) {
                if (qi < nqt) {
                    const double* s = sqt + qi * d + j;
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const double df = __dsub_rn(x[u], s[u]); acc[qi] = __dadd_rn(acc[qi], __dmul_rn(df, df)); }
                }
            }
        }
    } else {
        for (int j = 0; j < d; ++j) {
            const double x = __ldg(row + j);
#pragma unroll
            for (int qi = 0; qi < BF_TQ; ++qi)
                if (qi < nqt) { const double df = __dsub_rn(x, sqt[qi * d + j]); acc[qi] = __dadd_rn(acc[qi], __dmul_rn(df, df)); }
        }
    }
#pragma unroll
    for (int qi = 0; qi < BF_TQ; ++qi)
        if (qi < nqt) D[(int64_t)qi * n + i] = (ull)__double_as_longlong(__dsqrt_rn(acc[qi]));
}

// One CTA per query of the tile: k smallest of D[qi][0..n) ordered by (distance, row id)
// (the general path: nine passes over the n distances; `need` != NULL: only the queries flagged by k_topk_final)
__global__ void __launch_bounds__(512) k_select_topk(const ull* __restrict__ D, int64_t n, int k, int64_t q0,
                                                     double* __restrict__ dist, uint32_t* __restrict__ ids, const uint32_t* __restrict__ need) {
    __shared__ uint32_t sh[260];
    __shared__ ull s_pref;
    __shared__ ull skey[1024];
    __shared__ uint32_t sidv[1024];
    __shared__ unsigned s_cnt, s_cnt_eq;
    const int qi = blockIdx.x, tid = threadIdx.x;
    if (need && need[qi] == 0) return;
    const ull* Dq = D + (int64_t)qi * n;
    const uint32_t kk = (uint32_t)min((int64_t)k, n);
    // radix select of rank kk-1
    ull prefix = 0; uint32_t rr = kk - 1;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        const ull mask_hi = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int j = tid; j < 256; j += 512) sh[j] = 0;
        __syncthreads();
        for (int64_t i = tid; i < n; i += 512) { ull v = Dq[i]; if ((v & mask_hi) == prefix) atomicAdd(&sh[(v >> shift) & 255], 1u); }
        __syncthreads();
        if (tid == 0) {
            uint32_t cum = 0; int dg = 255;
            for (int b = 0; b < 256; ++b) { if (rr < cum + sh[b]) { dg = b; break; } cum += sh[b]; }
            sh[256] = cum; s_pref = prefix | ((ull)dg << shift);
        }
        __syncthreads();
        prefix = s_pref; rr -= sh[256];
        __syncthreads();
    }
    const ull kth = prefix;       // rr = rank of the wanted element among the values equal to kth
    const uint32_t need_eq = rr + 1;
    if (tid == 0) { s_cnt = 0; s_cnt_eq = 0; }
    __syncthreads();
    // everything strictly below kth
    for (int64_t i = tid; i < n; i += 512) {
        ull v = Dq[i];
        if (v < kth) { unsigned p = atomicAdd(&s_cnt, 1u); skey[p] = v; sidv[p] = (uint32_t)i; }
    }
    __syncthreads();
    const unsigned nlt = s_cnt;   // == kk - need_eq
    // the need_eq smallest row ids among the values equal to kth: ordered scan by one warp (ties are rare)
    if (tid < 32) {
        unsigned w = 0;
        for (int64_t base = 0; base < n && w < need_eq; base += 32) {
            const int64_t i = base + tid;
            const bool eq = i < n && Dq[i] == kth;
            const unsigned bal = __ballot_sync(0xffffffffu, eq);
            const unsigned dst = w + __popc(bal & ((1u << tid) - 1));
            if (eq && dst < need_eq) { skey[nlt + dst] = kth; sidv[nlt + dst] = (uint32_t)i; }
            w += __popc(bal);
        }
    }
    __syncthreads();
    // order the kk results by (distance, row id)
    {
        const unsigned m = kk, Pv = q_next_pow2(m), half = Pv >> 1;
        for (unsigned kq = 2; kq <= Pv; kq <<= 1) {
            const int lk = q_ilog2(kq);
            for (unsigned c = tid; c < half; c += 512) {
                const unsigned blk = c >> (lk - 1), w = c & ((kq >> 1) - 1);
                const unsigned i = (blk << lk) + w, p = (blk << lk) + (kq - 1 - w);
                if (p < m) {
                    ull a = skey[i], b = skey[p]; uint32_t ia = sidv[i], ib = sidv[p];
                    if (a > b || (a == b && ia > ib)) { skey[i] = b; skey[p] = a; sidv[i] = ib; sidv[p] = ia; }
                }
            }
            __syncthreads();
            for (unsigned j = kq >> 2; j > 0; j >>= 1) {
                const int lj = q_ilog2(j);
                for (unsigned c = tid; c < half; c += 512) {
                    const unsigned i = ((c >> lj) << (lj + 1)) + (c & (j - 1)), p = i + j;
                    if (p < m) {
                        ull a = skey[i], b = skey[p]; uint32_t ia = sidv[i], ib = sidv[p];
                        if (a > b || (a == b && ia > ib)) { skey[i] = b; skey[p] = a; sidv[i] = ib; sidv[p] = ia; }
                    }
                }
                __syncthreads();
            }
        }
    }
    for (unsigned i = tid; i < (unsigned)k; i += 512) {
        const bool ok = i < kk;
        dist[(q0 + qi) * k + i] = ok ? __longlong_as_double((long long)skey[i]) : __longlong_as_double(0x7ff0000000000000LL);
        ids[(q0 + qi) * k + i] = ok ? sidv[i] : 0xffffffffu;
    }
}

// ---- fast exact top-k: one pass over the distances instead of nine ---------------------------------------------
// tau = the kk-th smallest of a strided SAMPLE of the query's distances is an upper bound of the kk-th smallest of all n,
// so every member of the exact top-k (by (distance, row id)) passes the filter v <= tau; with S samples about kk * n / S
// rows pass.  The survivors are sorted by (distance, row id) in shared memory.  A query whose survivors exceed the
// buffer (adversarial distributions) is flagged and answered by k_select_topk.
#define TK_CAP 8192       /* survivors per query (96 KB of shared memory in k_topk_final) */
__global__ void __launch_bounds__(512) k_topk_tau(const ull* __restrict__ D, int64_t n, int kk, int64_t S, int64_t stride,
                                                  ull* __restrict__ tau, uint32_t* __restrict__ cnt, uint32_t* __restrict__ need) {
    __shared__ uint32_t sh[264];
    __shared__ ull sh64[1];
    const int qi = blockIdx.x;
    const ull* Dq = D + (int64_t)qi * n;
    uint32_t cl, ce;
    const ull t = cta_radix_select<512>((uint32_t)S, (uint32_t)(kk - 1), [&](uint32_t i) { return Dq[(int64_t)i * stride]; }, sh, sh64, cl, ce);
    if (threadIdx.x == 0) { tau[qi] = t; cnt[qi] = 0; need[qi] = 0; }
}
__global__ void __launch_bounds__(512) k_topk_filter(const ull* __restrict__ D, int64_t n, const ull* __restrict__ tau,
                                                     ull* __restrict__ cv, uint32_t* __restrict__ ci, uint32_t* __restrict__ cnt) {
    const int qi = blockIdx.y;
    const ull* Dq = D + (int64_t)qi * n;
    const ull t = tau[qi];
    const int64_t base = (int64_t)blockIdx.x * (512 * 16) + threadIdx.x;
    ull v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) { const int64_t i = base + e * 512; v[e] = i < n ? __ldcs(Dq + i) : ~0ull; }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int64_t i = base + e * 512;
        if (i < n && v[e] <= t) {
            const uint32_t p = atomicAdd(&cnt[qi], 1u);
            if (p < TK_CAP) { cv[(int64_t)qi * TK_CAP + p] = v[e]; ci[(int64_t)qi * TK_CAP + p] = (uint32_t)i; }
        }
    }
}
__global__ void __launch_bounds__(512) k_topk_final(const ull* __restrict__ cv, const uint32_t* __restrict__ ci, const uint32_t* __restrict__ cnt,
                                                    int kk, int k, int64_t q0, double* __restrict__ dist, uint32_t* __restrict__ ids,
                                                    uint32_t* __restrict__ need) {
    extern __shared__ ull tk_key[];                       // [Pv] keys, then [Pv] row ids
    const int qi = blockIdx.x, tid = threadIdx.x;
    const unsigned m = cnt[qi];
    if (m > TK_CAP) { if (tid == 0) need[qi] = 1; return; }
    const unsigned Pv = q_next_pow2(m < 2 ? 2 : m), half = Pv >> 1;
    uint32_t* tk_id = (uint32_t*)(tk_key + Pv);
    for (unsigned i = tid; i < Pv; i += 512) {
        tk_key[i] = i < m ? cv[(int64_t)qi * TK_CAP + i] : ~0ull;
        tk_id[i] = i < m ? ci[(int64_t)qi * TK_CAP + i] : 0xffffffffu;
    }
    __syncthreads();
    for (unsigned kq = 2; kq <= Pv; kq <<= 1) {           // bitonic network on (distance, row id)
        const int lk = q_ilog2(kq);
        for (unsigned c = tid; c < half; c += 512) {
            const unsigned blk = c >> (lk - 1), w = c & ((kq >> 1) - 1);
            const unsigned i = (blk << lk) + w, p = (blk << lk) + (kq - 1 - w);
            ull a = tk_key[i], b = tk_key[p]; uint32_t ia = tk_id[i], ib = tk_id[p];
            if (a > b || (a == b && ia > ib)) { tk_key[i] = b; tk_key[p] = a; tk_id[i] = ib; tk_id[p] = ia; }
        }
        __syncthreads();
        for (unsigned j = kq >> 2; j > 0; j >>= 1) {
            const int lj = q_ilog2(j);
            for (unsigned c = tid; c < half; c += 512) {
                const unsigned i = ((c >> lj) << (lj + 1)) + (c & (j - 1)), p = i + j;
                ull a = tk_key[i], b = tk_key[p]; uint32_t ia = tk_id[i], ib = tk_id[p];
                if (a > b || (a == b && ia > ib)) { tk_key[i] = b; tk_key[p] = a; tk_id[i] = ib; tk_id[p] = ia; }
            }
            __syncthreads();
        }
    }
    for (unsigned i = tid; i < (unsigned)k; i += 512) {
        const bool ok = i < (unsigned)kk;
        dist[(q0 + qi) * k + i] = ok ? __longlong_as_double((long long)tk_key[i]) : __longlong_as_double(0x7ff0000000000000LL);
        ids[(q0 + qi) * k + i] = ok ? tk_id[i] : 0xffffffffu;
    }
}

// ---------------------------------------------------------------------------------------------------
// knnH (RPTree.hs:199-217): the reached leaves of all trees go into one min-heap keyed by their margin priority
// (candidatesH); leaves are popped in increasing priority and PREPENDED to the result while the running total stays
// <= k (the first non-empty pop is always taken).  The result is NOT sorted by distance and NOT cut to k -- it is what
// the reference returns.  Order among EQUAL priorities is an internal of the `heaps` package (unpinned); here: tree
// index, then leaf position left to right.
// One CTA per query; the (tree, leaf) slots of a query (T * S of them, most unused) live in dynamic shared memory.
// ---------------------------------------------------------------------------------------------------
#define HK_NT 128
#define HK_MAX 9216      /* slots: 20 bytes each next to the query vector */
__global__ void __launch_bounds__(HK_NT) k_knn_h(QArgs A, const double* __restrict__ prio, int cap, int nslot_cap) {
    __shared__ unsigned s_m, s_total, s_nacc;
    extern __shared__ unsigned char dyn[];
    ull* skey = (ull*)dyn;
    double* sq = (double*)(skey + nslot_cap);
    uint32_t* spos = (uint32_t*)(sq + ((A.d + 3) & ~3));
    uint32_t* sleaf = spos + nslot_cap;
    uint32_t* soff = sleaf + nslot_cap;
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    if (tid == 0) s_m = 0;
    for (int j = tid; j < A.d; j += HK_NT) sq[j] = A.Q[q * A.d + j];
    __syncthreads();
    const unsigned nslots = (unsigned)A.T * A.S;
    for (unsigned s = tid; s < nslots; s += HK_NT) {
        const int tt = s / A.S, j = s % A.S;
        if ((uint32_t)j < A.cnt[q * A.T + tt]) {
            const unsigned p = atomicAdd(&s_m, 1u);
            const int64_t e = (q * A.T + tt) * (int64_t)A.S + j;
            skey[p] = (ull)__double_as_longlong(prio[e]);      // priorities are >= +0 or +inf: raw bits are order preserving
            spos[p] = s;
            sleaf[p] = A.segs[e];
        }
    }
    __syncthreads();
    const unsigned m = s_m;
    sort3<HK_NT>(skey, spos, sleaf, m);                       // (priority, slot) ascending
    if (tid == 0) {
        // go acc n hh (RPTree.hs:207-217): stop at the first pop that would exceed k once something has been taken
        unsigned n = 0, nacc = 0;
        for (unsigned i = 0; i < m; ++i) {
            const unsigned nels = A.nsize[sleaf[i]], ntot = n + nels;
            if (ntot > (unsigned)A.k && n > 0) break;
            soff[i] = n;                                         // elements taken before this leaf
            n = ntot; nacc = i + 1;
        }
        s_total = n; s_nacc = nacc;
    }
    __syncthreads();
    const unsigned total = s_total, nacc = s_nacc;
    for (unsigned i = 0; i < nacc; ++i) {
        const uint32_t g = sleaf[i], sz = A.nsize[g];
        const int tt = spos[i] / A.S;
        const unsigned base = total - soff[i] - sz;              // xsh <> acc: the latest pop comes first
        for (unsigned e = tid; e < sz; e += HK_NT) {
            if (base + e >= (unsigned)cap) continue;
            const uint32_t id = A.perm[(int64_t)tt * A.n + A.nstart[g] + e];
            const int len = metric_len(A, id, q);
            A.dist[q * cap + base + e] = dist_exact(A.X + (int64_t)id * A.d, sq, len, A.vec && len == A.d);
            A.ids[q * cap + base + e] = id;
        }
    }
    for (unsigned e = total + tid; e < (unsigned)cap; e += HK_NT) {          // unused tail of the row
        A.dist[q * cap + e] = __longlong_as_double(0x7ff0000000000000LL);
        A.ids[q * cap + e] = 0xffffffffu;
    }
    if (tid == 0 && A.count) A.count[q] = (int32_t)min(total, (unsigned)cap);
}

// recallWith: per query, per tree: |candidates /\ truth| ; recall_sum = fold over trees of hits/k
__global__ void __launch_bounds__(KNN_NT) k_recall(QArgs A, const uint32_t* __restrict__ truth, double* __restrict__ recall_sum) {
    __shared__ uint32_t part[KNN_NT + 1];
    __shared__ uint32_t s_truth[1024];
    extern __shared__ unsigned char dyn[];
    uint32_t* pre = (uint32_t*)dyn;
    uint32_t* hits = pre + (unsigned)A.T * A.S + 1;
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    const unsigned nslots = (unsigned)A.T * A.S;
    const unsigned kk = (unsigned)min((int64_t)A.k, A.n);
    for (unsigned i = tid; i < kk; i += KNN_NT) s_truth[i] = truth[q * A.k + i];
    for (int t = tid; t < A.T; t += KNN_NT) hits[t] = 0;
    load_slots(A, q, A.T, pre, KNN_NT);
    slot_prefix<KNN_NT>(pre, nslots, part);
    const uint32_t C = pre[nslots];
    for (uint32_t c = tid; c < C; c += KNN_NT) {
        const unsigned slot = find_slot(pre, nslots, c);
        const int tt = slot / A.S;
        const uint32_t g = A.segs[(q * A.T + tt) * (int64_t)A.S + (slot % A.S)];
        const uint32_t id = A.perm[(int64_t)tt * A.n + A.nstart[g] + (c - pre[slot])];
        bool hit = false;
        for (unsigned i = 0; i < kk; ++i) hit |= (s_truth[i] == id);
        if (hit) atomicAdd(&hits[tt], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int t = 0; t < A.T; ++t) s = s + (double)hits[t] / (double)A.k;   // left fold in tree order
        recall_sum[q] = s;
    }
}

// multi-GPU merge: G rank-major lists of up to k (dist, id) per query -> global top-k by (dist, rank, position).
// Rank g's lists start at dist + g * sd, ids + g * si, count + g * sc (elements): three rank-major arrays (sd = si = nq * k,
// sc = nq) or the packed per-rank chunks of the in-engine NCCL exchange.
__global__ void __launch_bounds__(KNN_NT) k_merge(int G, int64_t nq, int k, int dedup, const double* __restrict__ dist,
                                                  const uint32_t* __restrict__ ids, const int32_t* __restrict__ count,
                                                  int64_t sd, int64_t si, int64_t sc,
                                                  double* __restrict__ dist_out, uint32_t* __restrict__ ids_out, int32_t* __restrict__ count_out) {
    extern __shared__ unsigned char dyn[];
    ull* skey = (ull*)dyn;
    uint32_t* spos = (uint32_t*)(skey + KNN_BUF);
    uint32_t* sid = spos + KNN_BUF;
    __shared__ unsigned s_n;
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    unsigned nbest = 0;
    const unsigned per = (KNN_BUF - k) / k;     // ranks per chunk (k <= 1024 -> per >= 3)
    for (int g0 = 0; g0 < G; g0 += per) {
        const int g1 = min(G, g0 + (int)per);
        unsigned m = 0;
        for (int g = g0; g < g1; ++g) {
            const unsigned c = (unsigned)count[(int64_t)g * sc + q];
            for (unsigned j = tid; j < c; j += KNN_NT) {
                skey[nbest + m + j] = (ull)__double_as_longlong(dist[(int64_t)g * sd + q * k + j]);
                spos[nbest + m + j] = (uint32_t)(g * k + j);
                sid[nbest + m + j] = ids[(int64_t)g * si + q * k + j];
            }
            m += c;
        }
        __syncthreads();
        const unsigned tot = nbest + m;
        sort3<KNN_NT>(skey, spos, sid, tot);
        nbest = keep_k<KNN_NT>(skey, spos, sid, tot, (unsigned)k, dedup, &s_n);
        __syncthreads();
    }
    for (unsigned i = tid; i < (unsigned)k; i += KNN_NT) {
        const bool ok = i < nbest;
        dist_out[q * k + i] = ok ? __longlong_as_double((long long)skey[i]) : __longlong_as_double(0x7ff0000000000000LL);
        ids_out[q * k + i] = ok ? sid[i] : 0xffffffffu;
    }
    if (tid == 0 && count_out) count_out[q] = (int32_t)nbest;
}

// The same merge without a sort, for dedup = 0: every rank's list is already ordered by (distance, position), so the place
// of element (g, j) in the merged order is j + sum over the other lists of the number of their elements that precede it --
// all elements with a smaller distance, and on equal distances those of LOWER ranks (rank order = tree order,
// RPTree.hs:174-176).  Two binary searches per (element, other list) in shared memory; one CTA per query.
#define MR_NT 128
__global__ void __launch_bounds__(MR_NT) k_merge_rank(int G, int64_t nq, int k, const double* __restrict__ dist,
                                                      const uint32_t* __restrict__ ids, const int32_t* __restrict__ count,
                                                      int64_t sd, int64_t si, int64_t sc,
                                                      double* __restrict__ dist_out, uint32_t* __restrict__ ids_out, int32_t* __restrict__ count_out) {
    extern __shared__ unsigned char dyn[];
    ull* sk = (ull*)dyn;                         // [G][k] distance bits (non-negative doubles: bit order == value order)
    int* scnt = (int*)(sk + (size_t)G * k);      // [G]
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    for (int g = tid; g < G; g += MR_NT) scnt[g] = min(k, max(0, count[(int64_t)g * sc + q]));
    __syncthreads();
    for (int e = tid; e < G * k; e += MR_NT) {
        const int g = e / k, j = e % k;
        sk[e] = j < scnt[g] ? (ull)__double_as_longlong(dist[(int64_t)g * sd + q * k + j]) : ~0ull;
    }
    __syncthreads();
    int total = 0;
    for (int g = 0; g < G; ++g) total += scnt[g];
    const int nres = min(total, k);
    for (int e = tid; e < G * k; e += MR_NT) {
        const int g = e / k, j = e % k;
        if (j >= scnt[g]) continue;
        const ull key = sk[e];
        int rank = j;
        for (int g2 = 0; g2 < G && rank < k; ++g2) {
            if (g2 == g) continue;
            const ull* l = sk + (size_t)g2 * k;
            int lo = 0, hi = scnt[g2];
            if (g2 < g) { while (lo < hi) { const int mid = (lo + hi) >> 1; if (l[mid] <= key) lo = mid + 1; else hi = mid; } }   // # <= key
            else        { while (lo < hi) { const int mid = (lo + hi) >> 1; if (l[mid] <  key) lo = mid + 1; else hi = mid; } }   // # <  key
            rank += lo;
        }
        if (rank < k) {
            dist_out[q * k + rank] = __longlong_as_double((long long)key);
            ids_out[q * k + rank] = ids[(int64_t)g * si + q * k + j];
        }
    }
    for (int i = nres + tid; i < k; i += MR_NT) {
        dist_out[q * k + i] = __longlong_as_double(0x7ff0000000000000LL);
        ids_out[q * k + i] = 0xffffffffu;
    }
    if (tid == 0 && count_out) count_out[q] = nres;
}

// both merges behind one launch site
static int launch_merge(rpf_handle* h, int G, int64_t nq, int k, int dedup, const double* dd, const uint32_t* di, const int32_t* dc,
                        int64_t sd, int64_t si, int64_t sc, double* od, uint32_t* oi, int32_t* oc) {
    const size_t dynr = (size_t)G * k * 8 + (size_t)G * 4 + 16;
    if (!dedup && dynr <= 96 * 1024) {
        RPF_CUDA(h, cudaFuncSetAttribute(k_merge_rank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynr));
        RPF_LAUNCH(h, PH_MERGE, k_merge_rank, (unsigned)nq, MR_NT, dynr, G, nq, k, dd, di, dc, sd, si, sc, od, oi, oc);
        return RPF_OK;
    }
    const size_t dynm = (size_t)KNN_BUF * 16;
    RPF_CUDA(h, cudaFuncSetAttribute(k_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dynm));
    RPF_LAUNCH(h, PH_MERGE, k_merge, (unsigned)nq, KNN_NT, dynm, G, nq, k, dedup, dd, di, dc, sd, si, sc, od, oi, oc);
    return RPF_OK;
}

// recall sums of the ranks (rank-major [G][stride]) added in rank order = tree order (left fold, like the reference's sum)
__global__ void k_sum_ranks(const double* __restrict__ part, int G, int64_t stride, int64_t nq, double* __restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    double s = 0.0;
    for (int g = 0; g < G; ++g) s = s + part[(int64_t)g * stride + q];
    out[q] = s;
}

// ---------------------------------------------------------------------------------------------------
// host side (all device buffers come from the handle's persistent workspace)
// ---------------------------------------------------------------------------------------------------
#define QWS(h, var, type, slot, bytes)                                  \
    type* var = (type*)(h)->ws_get((slot), (bytes));                    \
    if (!var) return RPF_ERR_NOMEM;

// shared front half of every query entry point: upload Q, project, descend (with retry on fork overflow)
struct QState {
    double* dQ = nullptr; double* keysQ = nullptr; uint32_t* segs = nullptr; uint32_t* cnt = nullptr; uint32_t* maxcnt = nullptr;
    const int32_t* dqlast = nullptr;
    double* prio = nullptr;      // candidatesH priorities, parallel to segs (only when want_prio)
    bool want_prio = false;
    int S = 2, Tq = 0;
};

// SVector queries: device copy of q_last (host), after checking it against the data representation
static int upload_qlast(rpf_handle* h, const int32_t* q_last, int64_t nq, const int32_t** out) {
    *out = nullptr;
    if (!q_last) return RPF_OK;
    if (!h->d_xlast) return rpf_fail(h, RPF_ERR_ARG, "SVector queries need SVector data (the reference has no Inner DVector SVector instance)");
    for (int64_t i = 0; i < nq; ++i) if (q_last[i] < -1 || q_last[i] >= h->d) return rpf_fail(h, RPF_ERR_ARG, "q_last out of range");
    int32_t* dq = (int32_t*)h->ws_get(WS_QLAST, (size_t)nq * 4);
    if (!dq) return RPF_ERR_NOMEM;
    RPF_CUDA(h, cudaMemcpyAsync(dq, q_last, (size_t)nq * 4, cudaMemcpyHostToDevice, h->stream));
    *out = dq;
    return RPF_OK;
}

static int run_descent(rpf_handle* h, const double* Q, int64_t nq, int t_only, QState& st) {
    const int L = h->topo.L_eff, T = h->T;
    st.Tq = t_only >= 0 ? 1 : T;
    QWS(h, dQ, double, WS_Q, (size_t)nq * h->d * 8);
    RPF_CUDA(h, cudaMemcpyAsync(dQ, Q, (size_t)nq * h->d * 8, cudaMemcpyHostToDevice, h->stream));
    QWS(h, keysQ, double, WS_KEYSQ, (size_t)std::max(1, T * L) * nq * 8);
    if (L > 0) { int rc = rpf_project_queries(h, dQ, nq, keysQ); if (rc) return rc; }
    QWS(h, cnt, uint32_t, WS_CNT, (size_t)nq * st.Tq * 4);
    QWS(h, maxcnt, uint32_t, WS_MAXCNT, 4);
    st.dQ = dQ; st.keysQ = keysQ; st.cnt = cnt; st.maxcnt = maxcnt;
    while (true) {
        QWS(h, segs, uint32_t, WS_SEGS, (size_t)nq * st.Tq * st.S * 4);
        st.segs = segs;
        RPF_CUDA(h, cudaMemsetAsync(maxcnt, 0, 4, h->stream));
        const int64_t tot = nq * st.Tq;
        if (st.want_prio) {
            QWS(h, prio, double, WS_PRIO, (size_t)nq * st.Tq * st.S * 8);
            st.prio = prio;
            RPF_LAUNCH(h, PH_Q_TRAVERSE, k_traverse<true>, (unsigned)((tot + 127) / 128), 128, 0, keysQ, nq, T, L, h->topo.nnodes(),
                       h->d_node_child, h->d_node_depth, h->d_thr, h->d_mlo, h->d_mhi, st.S, t_only, segs, cnt, maxcnt, prio);
        } else {
            RPF_LAUNCH(h, PH_Q_TRAVERSE, k_traverse<false>, (unsigned)((tot + 127) / 128), 128, 0, keysQ, nq, T, L, h->topo.nnodes(),
                       h->d_node_child, h->d_node_depth, h->d_thr, h->d_mlo, h->d_mhi, st.S, t_only, segs, cnt, maxcnt, (double*)nullptr);
        }
        uint32_t mx = 0;
        RPF_CUDA(h, cudaMemcpyAsync(&mx, maxcnt, 4, cudaMemcpyDeviceToHost, h->stream));
        RPF_CUDA(h, cudaStreamSynchronize(h->stream));
        if (mx <= (uint32_t)st.S) break;
        st.S = (int)mx;          // a query forked into more leaves than the stride: redo with the exact maximum
    }
    return RPF_OK;
}

static QArgs make_qargs(rpf_handle* h, int64_t nq, const QState& st) {
    QArgs A{};
    A.n = h->n; A.nq = nq; A.nn = h->topo.nnodes(); A.d = h->d; A.T = h->T; A.S = st.S;
    A.vec = (h->d % 4 == 0) && (((uintptr_t)h->dX & 31) == 0);
    A.X = h->dX; A.Q = st.dQ; A.perm = h->d_perm; A.nstart = h->d_node_start; A.nsize = h->d_node_size;
    A.segs = st.segs; A.cnt = st.cnt;
    A.xlast = h->d_xlast; A.qlast = st.dqlast;
    return A;
}

// out_dev: dist / ids / count are DEVICE buffers of the caller (multi-GPU exchange without a host round trip)
int rpf_knn_impl(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int k, int dedup, double* dist, uint32_t* ids, int32_t* count, bool out_dev) {
    if (nq == 0) return RPF_OK;
    QState st;
    int rc = upload_qlast(h, q_last, nq, &st.dqlast);
    if (rc) return rc;
    rc = run_descent(h, Q, nq, -1, st);
    if (rc) return rc;
    QWS(h, wdist, double, WS_OUT_D, (size_t)nq * k * 8);
    QWS(h, wids, uint32_t, WS_OUT_I, (size_t)nq * k * 4);
    QWS(h, wcount, int32_t, WS_OUT_C, (size_t)nq * 4);
    double* ddist = out_dev ? dist : wdist; uint32_t* dids = out_dev ? ids : wids; int32_t* dcount = (out_dev && count) ? count : wcount;
    // rank of a tree-sharded forest (rpf_comm_init_rank / rpf_create_multi): the local lists are written straight into this
    // rank's chunk of ONE packed exchange buffer [dist | ids | count], all-gathered in place over NCCL on the engine's
    // stream (no host synchronisation in between), and merged by k_merge in (distance, rank, position) order -- which is
    // the reference's tree order because the ranks hold contiguous blocks of trees (RPTree.hs:174-176)
    const int W = out_dev ? 1 : rpf_comm_world(h);
    char* xbuf = nullptr; size_t chunk = 0, off_i = 0, off_c = 0;
    if (W > 1) {
        off_i = ((size_t)nq * k * 8 + 15) & ~(size_t)15;
        off_c = off_i + (((size_t)nq * k * 4 + 15) & ~(size_t)15);
        chunk = off_c + (((size_t)nq * 4 + 15) & ~(size_t)15);
        xbuf = (char*)h->ws_get(WS_MRG_D, chunk * W);
        if (!xbuf) return RPF_ERR_NOMEM;
        char* mine = xbuf + (size_t)rpf_comm_rank(h) * chunk;
        ddist = (double*)mine; dids = (uint32_t*)(mine + off_i); dcount = (int32_t*)(mine + off_c);
    }
    QArgs A = make_qargs(h, nq, st);
    A.k = k; A.dedup = dedup; A.dist = ddist; A.ids = dids; A.count = dcount;
    // long rows, many queries per leaf: leaf-grouped GEMM on the FP64 tensor cores selects, exact recompute of the survivors
    // (rerank.cu); the few queries it cannot settle (massive distance ties) go through the gather kernel below
    unsigned grid_q = (unsigned)nq;
    bool run_gather = true;
    if (rpf_rerank_gemm_wanted(h, nq, k, dedup)) {
        uint32_t nfb = 0;
        int rcg = rpf_rerank_gemm(h, st.dQ, nq, st.S, st.segs, st.cnt, k, ddist, dids, dcount, &nfb);
        if (rcg) return rcg;
        run_gather = nfb > 0;
        grid_q = nfb;
        A.order = (const uint32_t*)h->ws[WS_RR_FB].p;
    } else if (nq >= 64 && !h->no_query_order) {
        QWS(h, qhist, uint32_t, WS_QHIST, (size_t)QO_BUCKETS * 4);
        QWS(h, qorder, uint32_t, WS_QORDER, (size_t)nq * 4);
        RPF_CUDA(h, cudaMemsetAsync(qhist, 0, (size_t)QO_BUCKETS * 4, h->stream));
        RPF_LAUNCH(h, PH_Q_TRAVERSE, k_qorder_hist, (unsigned)((nq + 255) / 256), 256, 0, A, qhist);
        RPF_LAUNCH(h, PH_Q_TRAVERSE, k_qorder_scan, 1, 1024, 0, qhist);
        RPF_LAUNCH(h, PH_Q_TRAVERSE, k_qorder_scatter, (unsigned)((nq + 255) / 256), 256, 0, A, qhist, qorder);
        A.order = qorder;
    }
    // TMA path: rows are 16-byte multiples and a 2-stage ring of up to 32 rows fits next to the selection buffers
    const size_t tail = (size_t)((h->d + 3) & ~3) * 8 + ((size_t)h->T * st.S + 1) * 4;
    const int pitch = h->d * 8 + 16;
    int rows = (int)std::min<size_t>(32, (size_t)(72 * 1024) / ((size_t)KT_STAGES * pitch));
    const size_t dyn_tma = (size_t)KT_STAGES * rows * pitch + (size_t)(KT_BUF + KT_SREG) * 16 + (size_t)2 * KT_BUF * 4 + tail;
    const bool use_tma = (h->d % 2 == 0) && rows >= 1 && k <= KT_BUF / 2 && dyn_tma <= 112 * 1024 && !h->force_simple_knn &&
                         (((uintptr_t)h->dX & 15) == 0) && !h->d_xlast;     // SVector data: per-candidate prefix lengths -> gather kernel
    // fp32 filter pass + exact re-rank of the survivors (plain knn, whole-forest batches); queries it flags, and every other
    // case, go through the exact TMA kernel
    const int pitch32 = h->d * 4 + 16;
    // Occupancy decides here (sweep: profiles/r02_knn_f32_sweep.txt): per query the CTA drains its ring at every chunk end and
    // runs selects behind barriers, so three or four resident CTAs per SM hide one another's bubbles -- 3 stages x 30 rows
    // (72 KB per CTA, 3 per SM) for forests with many candidates per query, 2 x 28 (54 KB, 4 per SM) for few.
    const bool many = h->T >= 16;
    const int nst = std::max(1, std::min(4, h->knn_f32_cfg[0] ? h->knn_f32_cfg[0] : (many ? 3 : 2)));
    // (the side region holds the candidates within the margin of the k-th best: at least 4 k entries)
    const int fbuf = h->knn_f32_cfg[2] ? h->knn_f32_cfg[2] : (k <= 64 ? 768 : 1280), fsreg = h->knn_f32_cfg[3] ? h->knn_f32_cfg[3] : (k <= 64 ? 256 : 512);
    const size_t stage_budget = h->knn_f32_cfg[1] ? (size_t)(72 * 1024) : (many ? (size_t)47600 : (size_t)29700);
    const int rows32 = (int)std::min<size_t>(h->knn_f32_cfg[1] ? h->knn_f32_cfg[1] : 32, stage_budget / ((size_t)nst * pitch32));
    const int nb_exact = rows32 >= 1 ? (int)std::min<size_t>(KT_NT, ((size_t)nst * rows32 * pitch32) / (size_t)pitch) : 0;
    const size_t dyn_f32 = (((size_t)nst * std::max(rows32, 1) * pitch32 + 15) & ~(size_t)15) + (size_t)(fbuf + fsreg) * 16 + (size_t)2 * fbuf * 4 +
                           (size_t)((h->d + 3) & ~3) * 12 + ((size_t)h->T * st.S + 1) * 4;
    if (run_gather && use_tma && h->knn_filter32 && !dedup && grid_q == (unsigned)nq && (h->d % 4 == 0) && rows32 >= 4 && rows32 <= 32 && nb_exact >= 1 &&
        k <= fsreg / 4 && fsreg >= 128 && fbuf >= 2 * fsreg && dyn_f32 <= 112 * 1024 && !h->capturing && ensure_x32(h) == RPF_OK) {
        uint8_t* fb = (uint8_t*)h->ws_get(WS_KNN_FB, (size_t)nq);
        if (!fb) return RPF_ERR_NOMEM;
        RPF_CUDA(h, cudaMemsetAsync(fb, 0, (size_t)nq, h->stream));
        QArgs F = A;
        F.X32 = h->dX32; F.xmax = h->d_xmax; F.fb = fb;
        RPF_CUDA(h, cudaFuncSetAttribute(k_knn_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_f32));
        RPF_CUDA(h, cudaFuncSetAttribute(k_knn_f32, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        RPF_LAUNCH(h, PH_Q_KNN, k_knn_f32, grid_q, KT_NT, dyn_f32, F, rows32, pitch32, pitch, nb_exact, nst, fbuf, fsreg);
        A.only = fb;
    }
    if (!run_gather) {
    } else if (use_tma) {
        RPF_CUDA(h, cudaFuncSetAttribute(k_knn_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_tma));
        RPF_CUDA(h, cudaFuncSetAttribute(k_knn_tma, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        RPF_LAUNCH(h, PH_Q_KNN, k_knn_tma, grid_q, KT_NT, dyn_tma, A, rows, pitch);
    } else {
        const size_t dyn = (size_t)(KNN_BUF + KNN_SREG) * 16 + tail;
        if (dyn > 200 * 1024) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "knn: d / tree count too large for the query kernel's shared memory");
        RPF_CUDA(h, cudaFuncSetAttribute(k_knn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        RPF_LAUNCH(h, PH_Q_KNN, k_knn, grid_q, KNN_NT, dyn, A);
    }
    if (W > 1) {
        int rcx = rpf_comm_allgather(h, xbuf, chunk, h->stream);
        if (rcx) return rcx;
        if (dist) {      // ranks that want the result (all of them in a multi-process job, rank 0 of an in-process group)
            int rcm = launch_merge(h, W, nq, k, dedup, (const double*)xbuf, (const uint32_t*)(xbuf + off_i), (const int32_t*)(xbuf + off_c),
                                   (int64_t)(chunk / 8), (int64_t)(chunk / 4), (int64_t)(chunk / 4), wdist, wids, wcount);
            if (rcm) return rcm;
            ddist = wdist; dids = wids; dcount = wcount;
        }
    }
    if (!out_dev && dist) {
        RPF_CUDA(h, cudaMemcpyAsync(dist, ddist, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, h->stream));
        RPF_CUDA(h, cudaMemcpyAsync(ids, dids, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, h->stream));
        if (count) RPF_CUDA(h, cudaMemcpyAsync(count, dcount, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RPF_OK;
}

int rpf_knn_h_impl(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int k, int cap, double* dist, uint32_t* ids, int32_t* count) {
    if (nq == 0) return RPF_OK;
    QState st;
    st.want_prio = true;
    int rc = upload_qlast(h, q_last, nq, &st.dqlast);
    if (rc) return rc;
    rc = run_descent(h, Q, nq, -1, st);
    if (rc) return rc;
    const int64_t nslot = (int64_t)h->T * st.S;
    if (nslot > HK_MAX) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "knnH: more than 9216 (tree, leaf) slots per query");
    QWS(h, ddist, double, WS_OUT_D, (size_t)nq * cap * 8);
    QWS(h, dids, uint32_t, WS_OUT_I, (size_t)nq * cap * 4);
    QWS(h, dcount, int32_t, WS_OUT_C, (size_t)nq * 4);
    QArgs A = make_qargs(h, nq, st);
    A.k = k; A.dist = ddist; A.ids = dids; A.count = dcount;
    const int nslot_cap = (int)((nslot + 1) & ~(int64_t)1);
    const size_t dyn = (size_t)((h->d + 3) & ~3) * 8 + (size_t)nslot_cap * 20;
    if (dyn > 200 * 1024) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "knnH: dimension / leaf slots too large for shared memory");
    RPF_CUDA(h, cudaFuncSetAttribute(k_knn_h, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    RPF_LAUNCH(h, PH_Q_KNN, k_knn_h, (unsigned)nq, HK_NT, dyn, A, st.prio, cap, nslot_cap);
    RPF_CUDA(h, cudaMemcpyAsync(dist, ddist, (size_t)nq * cap * 8, cudaMemcpyDeviceToHost, h->stream));
    RPF_CUDA(h, cudaMemcpyAsync(ids, dids, (size_t)nq * cap * 4, cudaMemcpyDeviceToHost, h->stream));
    if (count) RPF_CUDA(h, cudaMemcpyAsync(count, dcount, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RPF_OK;
}

int rpf_candidates_impl(rpf_handle* h, const double* Q, int64_t nq, int t, int64_t* off_out, const int64_t* off_in, uint32_t* ids) {
    if (nq == 0) { if (off_out) off_out[0] = 0; return RPF_OK; }
    QState st;
    int rc = run_descent(h, Q, nq, t, st);
    if (rc) return rc;
    QArgs A = make_qargs(h, nq, st);
    if (off_out) {
        QWS(h, dc, unsigned long long, WS_CANDCNT, (size_t)nq * 8);
        RPF_LAUNCH(h, PH_Q_CAND, k_cand_count, (unsigned)((nq + 127) / 128), 128, 0, A, st.Tq, dc);
        std::vector<unsigned long long> c((size_t)nq);
        RPF_CUDA(h, cudaMemcpyAsync(c.data(), dc, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
        RPF_CUDA(h, cudaStreamSynchronize(h->stream));
        int64_t acc = 0;
        for (int64_t q = 0; q < nq; ++q) { off_out[q] = acc; acc += (int64_t)c[q]; }
        off_out[nq] = acc;
        return RPF_OK;
    }
    const int64_t total = off_in[nq];
    if (total == 0) return RPF_OK;
    if (!ids) return rpf_fail(h, RPF_ERR_ARG, "candidates: ids is NULL");
    QWS(h, doff, int64_t, WS_CANDOFF, (size_t)(nq + 1) * 8);
    QWS(h, dout, uint32_t, WS_CANDOUT, (size_t)total * 4);
    RPF_CUDA(h, cudaMemcpyAsync(doff, off_in, (size_t)(nq + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    const size_t dyn = ((size_t)st.Tq * st.S + 1) * 4;
    RPF_CUDA(h, cudaFuncSetAttribute(k_cand_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(dyn, 1024)));
    RPF_LAUNCH(h, PH_Q_CAND, k_cand_fill, (unsigned)nq, KNN_NT, dyn, A, st.Tq, t, doff, dout);
    RPF_CUDA(h, cudaMemcpyAsync(ids, dout, (size_t)total * 4, cudaMemcpyDeviceToHost, h->stream));
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RPF_OK;
}

// exact top-k of all n rows for queries dQ[0..nq) -> device arrays d_dist/d_ids (nq x k)
static int brute_device(rpf_handle* h, const double* dQ, const int32_t* dqlast, int64_t nq, int k, double* d_dist, uint32_t* d_ids) {
    const int64_t n = h->n;
    const int d = h->d;
    if (n == 0) return rpf_fail(h, RPF_ERR_STATE, "brute_knn: no points");
    QWS(h, D, ull, WS_BF_D, (size_t)BF_TQ * n * 8);
    const int vec = (d % 4 == 0) && (((uintptr_t)h->dX & 31) == 0);
    const size_t smem = (size_t)BF_TQ * d * 8;
    if (smem > 160 * 1024) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "brute_knn: dimension too large");
    RPF_CUDA(h, cudaFuncSetAttribute(k_dist_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // sampled-threshold select (k_topk_*): needs a sample much larger than k and much smaller than n
    const int kk = (int)std::min<int64_t>(k, n);
    const int64_t S = std::min<int64_t>(n, std::max<int64_t>(16384, 256 * (int64_t)kk));    // ~ kk * n / S survivors per query
    const int64_t stride = n / S;
    const bool fast = !h->force_simple_topk && kk >= 1 && n >= 4 * S;
    ull* cv = nullptr; uint32_t* ci = nullptr; ull* tau = nullptr; uint32_t *cnt = nullptr, *need = nullptr;
    if (fast) {
        cv = (ull*)h->ws_get(WS_BF_CV, (size_t)BF_TQ * TK_CAP * 8);
        ci = (uint32_t*)h->ws_get(WS_BF_CI, (size_t)BF_TQ * TK_CAP * 4);
        tau = (ull*)h->ws_get(WS_BF_AUX, (size_t)BF_TQ * 16);
        if (!cv || !ci || !tau) return RPF_ERR_NOMEM;
        cnt = (uint32_t*)(tau + BF_TQ); need = cnt + BF_TQ;
        RPF_CUDA(h, cudaFuncSetAttribute(k_topk_final, cudaFuncAttributeMaxDynamicSharedMemorySize, TK_CAP * 12));
    }
    for (int64_t q0 = 0; q0 < nq; q0 += BF_TQ) {
        const int nqt = (int)std::min<int64_t>(BF_TQ, nq - q0);
        RPF_LAUNCH(h, PH_TRUTH, k_dist_all, (unsigned)((n + BF_NT - 1) / BF_NT), BF_NT, smem, h->dX, n, d, dQ, q0, nqt, D, vec, h->d_xlast, dqlast);
        if (fast) {
            RPF_LAUNCH(h, PH_TRUTH, k_topk_tau, (unsigned)nqt, 512, 0, D, n, kk, S, stride, tau, cnt, need);
            RPF_LAUNCH(h, PH_TRUTH, k_topk_filter, dim3((unsigned)((n + 512 * 16 - 1) / (512 * 16)), (unsigned)nqt), 512, 0, D, n, tau, cv, ci, cnt);
            RPF_LAUNCH(h, PH_TRUTH, k_topk_final, (unsigned)nqt, 512, (size_t)TK_CAP * 12, cv, ci, cnt, kk, k, q0, d_dist, d_ids, need);
            RPF_LAUNCH(h, PH_TRUTH, k_select_topk, (unsigned)nqt, 512, 0, D, n, k, q0, d_dist, d_ids, need);     // flagged queries only
        } else {
            RPF_LAUNCH(h, PH_TRUTH, k_select_topk, (unsigned)nqt, 512, 0, D, n, k, q0, d_dist, d_ids, (const uint32_t*)nullptr);
        }
    }
    return RPF_OK;
}

int rpf_brute_knn_impl(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int k, double* dist, uint32_t* ids) {
    if (nq == 0) return RPF_OK;
    const int32_t* dqlast = nullptr;
    int rc0 = upload_qlast(h, q_last, nq, &dqlast);
    if (rc0) return rc0;
    QWS(h, dQ, double, WS_Q, (size_t)nq * h->d * 8);
    QWS(h, dd, double, WS_TRUTH_D, (size_t)nq * k * 8);
    QWS(h, di, uint32_t, WS_TRUTH_I, (size_t)nq * k * 4);
    RPF_CUDA(h, cudaMemcpyAsync(dQ, Q, (size_t)nq * h->d * 8, cudaMemcpyHostToDevice, h->stream));
    int rc = brute_device(h, dQ, dqlast, nq, k, dd, di);
    if (rc) return rc;
    RPF_CUDA(h, cudaMemcpyAsync(dist, dd, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, h->stream));
    RPF_CUDA(h, cudaMemcpyAsync(ids, di, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, h->stream));
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RPF_OK;
}

int rpf_recall_impl(rpf_handle* h, const double* Q, const int32_t* q_last, int64_t nq, int k, double* recall_sum) {
    if (nq == 0) return RPF_OK;
    QState st;
    int rc = upload_qlast(h, q_last, nq, &st.dqlast);
    if (rc) return rc;
    rc = run_descent(h, Q, nq, -1, st);
    if (rc) return rc;
    // Rank of a tree-sharded forest: the brute-force truth (the expensive part: n x d per query) is computed for a 1/W
    // slice of the queries per rank and all-gathered; every rank then counts the hits of ITS trees and the per-rank sums
    // are added in rank (= tree) order, so recall_sum is the sum over the WHOLE forest (RPTree.hs:265-268).
    const int W = rpf_comm_world(h), R = rpf_comm_rank(h);
    const int64_t per = (nq + W - 1) / W;
    QWS(h, dd, double, WS_TRUTH_D, (size_t)per * W * k * 8);
    QWS(h, di, uint32_t, WS_TRUTH_I, (size_t)per * W * k * 4);
    const size_t res_n = (size_t)per * W;      // >= nq
    QWS(h, dr, double, WS_RECALL, (res_n + (W > 1 ? (size_t)W * res_n : 0)) * 8);   // result, then W > 1: [W][res_n] per-rank sums
    double* part = dr + res_n;
    double* mine = W > 1 ? part + (size_t)R * res_n : dr;
    const int64_t q0 = std::min<int64_t>(nq, R * per), q1 = std::min<int64_t>(nq, q0 + per);
    if (q1 > q0) {
        rc = brute_device(h, st.dQ + q0 * h->d, st.dqlast ? st.dqlast + q0 : nullptr, q1 - q0, k, dd + q0 * k, di + q0 * k);
        if (rc) return rc;
    }
    if (W > 1) { rc = rpf_comm_allgather(h, di, (size_t)per * k * 4, h->stream); if (rc) return rc; }
    QArgs A = make_qargs(h, nq, st);
    A.k = k;
    const size_t dyn = ((size_t)h->T * st.S + 1) * 4 + (size_t)h->T * 4;
    RPF_CUDA(h, cudaFuncSetAttribute(k_recall, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(dyn, 1024)));
    RPF_LAUNCH(h, PH_RECALL, k_recall, (unsigned)nq, KNN_NT, dyn, A, di, mine);
    if (W > 1) {
        rc = rpf_comm_allgather(h, part, res_n * 8, h->stream);
        if (rc) return rc;
        RPF_LAUNCH(h, PH_RECALL, k_sum_ranks, (unsigned)((nq + 255) / 256), 256, 0, part, W, (int64_t)res_n, nq, dr);
    }
    if (recall_sum) RPF_CUDA(h, cudaMemcpyAsync(recall_sum, dr, (size_t)nq * 8, cudaMemcpyDeviceToHost, h->stream));
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RPF_OK;
}

// in_dev: dist / ids / count (the gathered rank-major lists) already live on this handle's device
int rpf_merge_impl(rpf_handle* h, int G, int64_t nq, int k, int dedup, const double* dist, const uint32_t* ids,
                   const int32_t* count, double* dist_out, uint32_t* ids_out, int32_t* count_out, bool in_dev) {
    if (nq == 0) return RPF_OK;
    const size_t ne = (size_t)G * nq * k;
    const double* dd; const uint32_t* di; const int32_t* dc;
    if (in_dev) { dd = dist; di = ids; dc = count; }
    else {
        QWS(h, wd, double, WS_MRG_D, ne * 8);
        QWS(h, wi, uint32_t, WS_MRG_I, ne * 4);
        QWS(h, wc, int32_t, WS_MRG_C, (size_t)G * nq * 4);
        RPF_CUDA(h, cudaMemcpyAsync(wd, dist, ne * 8, cudaMemcpyHostToDevice, h->stream));
        RPF_CUDA(h, cudaMemcpyAsync(wi, ids, ne * 4, cudaMemcpyHostToDevice, h->stream));
        RPF_CUDA(h, cudaMemcpyAsync(wc, count, (size_t)G * nq * 4, cudaMemcpyHostToDevice, h->stream));
        dd = wd; di = wi; dc = wc;
    }
    QWS(h, od, double, WS_OUT_D, (size_t)nq * k * 8);
    QWS(h, oi, uint32_t, WS_OUT_I, (size_t)nq * k * 4);
    QWS(h, oc, int32_t, WS_OUT_C, (size_t)nq * 4);
    int rcm = launch_merge(h, G, nq, k, dedup, dd, di, dc, nq * k, nq * k, nq, od, oi, oc);
    if (rcm) return rcm;
    RPF_CUDA(h, cudaMemcpyAsync(dist_out, od, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, h->stream));
    RPF_CUDA(h, cudaMemcpyAsync(ids_out, oi, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, h->stream));
    if (count_out) RPF_CUDA(h, cudaMemcpyAsync(count_out, oc, (size_t)nq * 4, cudaMemcpyDeviceToHost, h->stream));
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RPF_OK;
}
