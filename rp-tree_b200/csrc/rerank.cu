// rerank.cu -- leaf-grouped re-rank on the FP64 tensor cores (DMMA) for long rows.
//
// knn (src/Data/RPTree.hs:174-176) evaluates metricDDL2 (Internal.hs:403-406) between the query and EVERY candidate of
// every tree.  With d = 960 (configs[2]) a query gathers ~3900 rows of 7.7 KB = 30 MB; but the candidates of a query are
// whole leaves, and with Q >> #leaves many queries visit the same leaf of the same tree.  So the work is regrouped by
// (tree, leaf): the ~61 rows of the leaf are staged ONCE and multiplied with all the queries that reached it --
//   dot[m][r] = q_m . x_r   as an m x d x 64 GEMM on the FP64 tensor cores (mma.sync.m8n8k4.f64; tcgen05 has no f64 kind)
//   approx d^2 = |q|^2 + |x|^2 - 2 dot
// This form rounds differently from the reference's left-to-right sum of squared differences, so it only SELECTS: per
// query the k-th smallest approximate distance plus a rigorous error margin defines the survivors (about k + a few),
// whose distances are then recomputed in the reference's order (dist_exact) and sorted by (distance, position in the
// reference's concatenation order).  Every candidate that can be among the exact top-k survives (margin argument in
// k_rr_select), so ids, distance bits and tie order are identical to the gather path / the oracle.
#include "rpf_internal.h"
#include "rpf_device.cuh"
#include <algorithm>

// ---- small device-wide exclusive scan (uint32 counts -> uint64 offsets): block scan, scan of the block sums, add -----
#define SC_NT 1024
#define SC_PER 4
__global__ void __launch_bounds__(SC_NT) k_rr_scan_block(const uint32_t* __restrict__ in, int64_t n, ull* __restrict__ out, ull* __restrict__ bsum) {
    __shared__ ull wsum[32];
    const int64_t base = ((int64_t)blockIdx.x * SC_NT + threadIdx.x) * SC_PER;
    ull v[SC_PER], s = 0;
#pragma unroll
    for (int i = 0; i < SC_PER; ++i) { v[i] = base + i < n ? in[base + i] : 0u; s += v[i]; }
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    ull incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const ull y = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= (unsigned)off) incl += y; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    if (w == 0) {
        ull x = wsum[lane], in2 = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const ull y = __shfl_up_sync(0xffffffffu, in2, off); if (lane >= (unsigned)off) in2 += y; }
        wsum[lane] = in2 - x;
        if (lane == 31) bsum[blockIdx.x] = in2;
    }
    __syncthreads();
    ull run = wsum[w] + incl - s;
#pragma unroll
    for (int i = 0; i < SC_PER; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
}
__global__ void __launch_bounds__(1024) k_rr_scan_top(ull* __restrict__ bsum, int nb, ull* __restrict__ total) {
    __shared__ ull carry;
    __shared__ ull wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + (int)threadIdx.x;
        const ull x = i < nb ? bsum[i] : 0ull;
        ull incl = x;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const ull y = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= (unsigned)off) incl += y; }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        if (w == 0) {
            ull y = wsum[lane], in2 = y;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const ull z = __shfl_up_sync(0xffffffffu, in2, off); if (lane >= (unsigned)off) in2 += z; }
            wsum[lane] = in2 - y;
        }
        __syncthreads();
        const ull excl = carry + wsum[w] + incl - x;
        if (i < nb) bsum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}
__global__ void __launch_bounds__(SC_NT) k_rr_scan_add(ull* __restrict__ out, int64_t n, const ull* __restrict__ bsum) {
    const int64_t base = ((int64_t)blockIdx.x * SC_NT + threadIdx.x) * SC_PER;
    const ull add = bsum[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SC_PER; ++i) if (base + i < n) out[base + i] += add;
}

// ---- |x|^2 per row (any summation order: it only feeds the approximate distances) and its maximum -------------------------
__global__ void __launch_bounds__(256) k_rr_norms(const double* __restrict__ X, int64_t n, int d, double* __restrict__ xn, ull* __restrict__ xmax) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const double* r = X + row * (int64_t)d;
    double s = 0.0;
    for (int c = lane; c < d; c += 32) { const double v = __ldg(r + c); s = fma(v, v, s); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) {
        xn[row] = s;
        if (xmax) atomicMax(xmax, (ull)__double_as_longlong(s));      // non-negative doubles: bit order == value order
    }
}

// ---- per query: candidate count, and the histogram of (tree, leaf) visits ---------------------------------------------------------
struct RRArgs {
    int64_t n, nq, nn;
    int d, T, S, k;
    const double* X; const double* Q;
    const uint32_t* perm; const uint32_t* nstart; const uint32_t* nsize;
    const uint32_t* segs; const uint32_t* cnt;
    uint32_t* ccount;          // [nq] candidates per query
    ull* qoff;                 // [nq] first entry of the query in the approximate-distance buffer
    uint32_t* hist;            // [T * nn] visits per (tree, leaf)
    ull* start;                // [T * nn] first entry of the (tree, leaf) group
    uint32_t* cursor;          // [T * nn]
    uint32_t* ent_q; ull* ent_dst;
    const double* xn; const double* qn; const ull* xmax;
    float* dap;                // approximate squared distances, query-major in the reference's concatenation order
    const uint32_t* leaves; int nleaves;
    double* dist; uint32_t* ids; int32_t* count;
    uint32_t* fb_list; uint32_t* fb_count;     // queries the selection could not settle (answered by the gather kernel)
};

template <bool FILL>
__global__ void __launch_bounds__(256) k_rr_group(RRArgs A) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (q >= A.nq) return;
    const unsigned nslots = (unsigned)A.T * A.S;
    uint32_t run = 0;
    for (unsigned s0 = 0; s0 < nslots; s0 += 32) {
        const unsigned s = s0 + lane;
        uint32_t sz = 0, g = 0; int tt = 0;
        if (s < nslots) {
            tt = s / A.S; const unsigned j = s % A.S;
            if (j < A.cnt[q * A.T + tt]) { g = A.segs[(q * A.T + tt) * (int64_t)A.S + j]; sz = A.nsize[g]; }
        }
        uint32_t incl = sz;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += y; }
        if (sz) {
            const int64_t key = (int64_t)tt * A.nn + g;
            if (!FILL) atomicAdd(&A.hist[key], 1u);
            else {
                const uint32_t p = atomicAdd(&A.cursor[key], 1u);
                const ull e = A.start[key] + p;
                A.ent_q[e] = (uint32_t)q;
                A.ent_dst[e] = A.qoff[q] + run + (incl - sz);
            }
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (!FILL && lane == 0) A.ccount[q] = run;
}

// ---- the GEMM: one CTA per visited (tree, leaf) ---------------------------------------------------------------------------------------
#define RG_NT 256
#define RG_KC 64              /* columns per staged chunk */
#define RG_LD (RG_KC + 4)     /* row stride in doubles: 544 B, rows 32 B apart modulo 128 -> the fragment loads are conflict free */
#define RG_ROWS 64            /* leaf rows per CTA (leaves are <= minLeaf points; the host checks) */
#define RG_MQ 32              /* queries per pass */

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src, bool live) {
    const unsigned nb = live ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(nb) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(RG_NT, 2) k_rr_gemm(RRArgs A) {
    extern __shared__ __align__(16) double rg_sm[];
    double* Xs = rg_sm;                                   // [2][RG_ROWS][RG_LD]
    double* Qs = rg_sm + 2 * RG_ROWS * RG_LD;             // [2][RG_MQ][RG_LD]
    __shared__ uint32_t s_row[RG_ROWS];
    __shared__ uint32_t s_q[RG_MQ];
    __shared__ ull s_dst[RG_MQ];
    const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t g = A.leaves[blockIdx.x];
    const int64_t key = (int64_t)t * A.nn + g;
    const uint32_t m = A.hist[key];
    if (m == 0) return;
    const ull ebase = A.start[key];
    const uint32_t R = A.nsize[g];
    const int d = A.d;
    if (tid < RG_ROWS) s_row[tid] = tid < (int)R ? A.perm[(int64_t)t * A.n + A.nstart[g] + tid] : 0u;
    const int nchunk = (d + RG_KC - 1) / RG_KC;
    const int gm = lane >> 2, gk = lane & 3;              // fragment coordinates of this lane
    for (uint32_t m0 = 0; m0 < m; m0 += RG_MQ) {
        const uint32_t mc = min((uint32_t)RG_MQ, m - m0);
        __syncthreads();                                  // previous pass done with s_q / the stage buffers
        if (tid < RG_MQ) {
            s_q[tid] = tid < (int)mc ? A.ent_q[ebase + m0 + tid] : 0u;
            s_dst[tid] = tid < (int)mc ? A.ent_dst[ebase + m0 + tid] : 0ull;
        }
        __syncthreads();
        auto stage = [&](int ch, int b) {
            const int c0 = ch * RG_KC;
            double* xd = Xs + (size_t)b * RG_ROWS * RG_LD;
            double* qd = Qs + (size_t)b * RG_MQ * RG_LD;
            for (int i = tid; i < RG_ROWS * (RG_KC / 2); i += RG_NT) {       // 16-byte pieces
                const int r = i / (RG_KC / 2), c = (i % (RG_KC / 2)) * 2;
                const bool live = r < (int)R && c0 + c < d;
                cp_async16((unsigned)__cvta_generic_to_shared(xd + r * RG_LD + c), A.X + (int64_t)s_row[r] * d + (live ? c0 + c : 0), live);
            }
            for (int i = tid; i < RG_MQ * (RG_KC / 2); i += RG_NT) {
                const int r = i / (RG_KC / 2), c = (i % (RG_KC / 2)) * 2;
                const bool live = r < (int)mc && c0 + c < d;
                cp_async16((unsigned)__cvta_generic_to_shared(qd + r * RG_LD + c), A.Q + (int64_t)s_q[r] * d + (live ? c0 + c : 0), live);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        double acc[RG_MQ / 8][2];
#pragma unroll
        for (int mt = 0; mt < RG_MQ / 8; ++mt) { acc[mt][0] = 0.0; acc[mt][1] = 0.0; }
        const int mtiles = (int)((mc + 7) >> 3);
        stage(0, 0);
        for (int ch = 0; ch < nchunk; ++ch) {
            const int b = ch & 1;
            if (ch + 1 < nchunk) { stage(ch + 1, b ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            const double* xb = Xs + (size_t)b * RG_ROWS * RG_LD + (8 * w + gm) * RG_LD + gk;      // warp w: leaf rows 8w .. 8w+7
            const double* qb = Qs + (size_t)b * RG_MQ * RG_LD + gm * RG_LD + gk;
            if (8 * w < (int)R) {
#pragma unroll 4
                for (int kk = 0; kk < RG_KC; kk += 4) {
                    const double bv = xb[kk];
#pragma unroll
                    for (int mt = 0; mt < RG_MQ / 8; ++mt)
                        if (mt < mtiles) dmma884(acc[mt][0], acc[mt][1], qb[mt * 8 * RG_LD + kk], bv);
                }
            }
            __syncthreads();
        }
        // approx |x - q|^2 = |q|^2 + |x|^2 - 2 q.x  for (query gm of tile mt, leaf rows 8w + 2 gk, + 1)
#pragma unroll
        for (int mt = 0; mt < RG_MQ / 8; ++mt) {
            const int mi = mt * 8 + gm;
            if (mi < (int)mc) {
                const double qn = A.qn[s_q[mi]];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int r = 8 * w + 2 * gk + u;
                    if (r < (int)R) A.dap[s_dst[mi] + r] = (float)(qn + A.xn[s_row[r]] - 2.0 * acc[mt][u]);
                }
            }
        }
    }
}

// ---- per query: select the survivors of the approximate distances, recompute them exactly, stable top-k -----------------
#define RS_NT 256
#define RS_CACHE 8192         /* approximate distances cached in shared memory (more: read from global in every pass) */
#define RS_SURV 1024          /* survivors per query (more: the query goes to the gather kernel) */

__device__ __forceinline__ uint32_t f2ord32(float x) {
    uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f32(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// sqrt(sum (x_j - q_j)^2), left fold, separate roundings -- the same statement as query.cu's dist_exact (Internal.hs:403-406)
__device__ __forceinline__ double rr_dist_exact(const double* __restrict__ row, const double* __restrict__ sq, int d) {
    double acc = 0.0;
    for (int j = 0; j < d; j += 4) {
        double x0, x1, x2, x3;
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(x3) : "l"(row + j));
        double df = __dsub_rn(x0, sq[j]);     acc = __dadd_rn(acc, __dmul_rn(df, df));
        df = __dsub_rn(x1, sq[j + 1]);        acc = __dadd_rn(acc, __dmul_rn(df, df));
        df = __dsub_rn(x2, sq[j + 2]);        acc = __dadd_rn(acc, __dmul_rn(df, df));
        df = __dsub_rn(x3, sq[j + 3]);        acc = __dadd_rn(acc, __dmul_rn(df, df));
    }
    return __dsqrt_rn(acc);
}

__global__ void __launch_bounds__(RS_NT) k_rr_select(RRArgs A) {
    extern __shared__ __align__(16) unsigned char rs_sm[];
    float* cache = (float*)rs_sm;                               // [RS_CACHE]
    ull* skey = (ull*)(cache + RS_CACHE);                       // [RS_SURV]
    uint32_t* spos = (uint32_t*)(skey + RS_SURV);               // [RS_SURV]
    uint32_t* sid = spos + RS_SURV;                             // [RS_SURV]
    double* sq = (double*)(sid + RS_SURV);                      // [d]
    uint32_t* pre = (uint32_t*)(sq + A.d);                      // [T*S + 1]
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_pref, s_rr, s_ns;
    const int64_t q = blockIdx.x;
    const int tid = threadIdx.x;
    const uint32_t C = A.ccount[q];
    const ull base = A.qoff[q];
    const unsigned k = (unsigned)A.k, nslots = (unsigned)A.T * A.S;
    const bool cached = C <= RS_CACHE;
    for (int j = tid; j < A.d; j += RS_NT) sq[j] = A.Q[q * A.d + j];
    for (unsigned s = tid; s < nslots; s += RS_NT) {
        const int tt = s / A.S; const unsigned j = s % A.S;
        pre[s] = j < A.cnt[q * A.T + tt] ? A.nsize[A.segs[(q * A.T + tt) * (int64_t)A.S + j]] : 0u;
    }
    if (cached) for (uint32_t i = tid; i < C; i += RS_NT) cache[i] = A.dap[base + i];
    if (tid == 0) { s_pref = 0; s_rr = min(k, C) - (C ? 1u : 0u); s_ns = 0; }
    __syncthreads();
    if (tid == 0) {        // serial prefix of the slot sizes (T*S entries; once per query)
        uint32_t c = 0;
        for (unsigned s = 0; s < nslots; ++s) { const uint32_t v = pre[s]; pre[s] = c; c += v; }
        pre[nslots] = c;
    }
    auto val = [&](uint32_t i) -> float { return cached ? cache[i] : A.dap[base + i]; };
    unsigned nbest = 0;
    if (C > 0) {
        // ---- value of rank min(k, C) - 1 among the approximate distances: 4 passes of an 8-bit MSD radix select
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            const uint32_t mask_hi = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
            hist[tid] = 0;
            __syncthreads();
            const uint32_t pref = s_pref;
            for (uint32_t i = tid; i < C; i += RS_NT) {
                const uint32_t v = f2ord32(val(i));
                if ((v & mask_hi) == pref) atomicAdd(&hist[(v >> shift) & 255], 1u);
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t rr = s_rr, cum = 0; int dg = 255;
                for (int b = 0; b < 256; ++b) { if (rr < cum + hist[b]) { dg = b; break; } cum += hist[b]; }
                s_rr = rr - cum; s_pref = pref | ((uint32_t)dg << shift);
            }
            __syncthreads();
        }
        // ---- survivors.  Let e_i = |approx_i - exact_i| (both roundings of the same real number): e_i <= E with
        //   E = 2^-20 (|q|^2 + max|x|^2): the float store contributes 2^-24 relative to a value <= 2(|q|^2 + |x|^2) plus the
        //   cancellation of the three-term form, the fp64 dot products ~d 2^-53 -- orders of magnitude below.
        // If i is among the exact top-k then exact_i <= (k-th smallest exact) <= (k-th smallest approx) + E, hence
        // approx_i <= tau + 2E: every such i is kept.  (Compared as squared distances; sqrt is monotone.)
        const float tau = ord2f32(s_pref);
        const double E = ldexp(A.qn[q] + __longlong_as_double((long long)*A.xmax), -20);
        const float thr = (float)((double)tau + 2.0 * E + (double)fabsf(tau) * 1e-6);
        for (uint32_t i = tid; i < C; i += RS_NT) {
            if (val(i) <= thr) { const uint32_t p = atomicAdd(&s_ns, 1u); if (p < RS_SURV) spos[p] = i; }
        }
        __syncthreads();
        const uint32_t ns = s_ns;
        if (ns > RS_SURV) {          // massive ties (duplicate rows): the gather kernel answers this query
            if (tid == 0) A.fb_list[atomicAdd(A.fb_count, 1u)] = (uint32_t)q;
            return;
        }
        // ---- exact distances of the survivors, in the reference's arithmetic
        for (uint32_t j = tid; j < ns; j += RS_NT) {
            const uint32_t c = spos[j];
            unsigned lo = 0, hi = nslots;
            while (hi - lo > 1) { const unsigned mid = (lo + hi) >> 1; if (pre[mid] <= c) lo = mid; else hi = mid; }
            const int tt = lo / A.S;
            const uint32_t g = A.segs[(q * A.T + tt) * (int64_t)A.S + (lo % A.S)];
            const uint32_t id = A.perm[(int64_t)tt * A.n + A.nstart[g] + (c - pre[lo])];
            sid[j] = id;
            skey[j] = (ull)__double_as_longlong(rr_dist_exact(A.X + (int64_t)id * A.d, sq, A.d));
        }
        __syncthreads();
        // ---- (distance, position) order == the reference's stable sort; bitonic network over the survivors
        const unsigned Pv = next_pow2_u32(ns), half = Pv >> 1;
        for (unsigned kk = 2; kk <= Pv; kk <<= 1) {
            const int lk = ilog2_pow2(kk);
            for (unsigned c = tid; c < half; c += RS_NT) {
                const unsigned blk = c >> (lk - 1), x = c & ((kk >> 1) - 1);
                const unsigned i = (blk << lk) + x, p = (blk << lk) + (kk - 1 - x);
                if (p < ns) {
                    const ull a = skey[i], b = skey[p]; const uint32_t pa = spos[i], pb = spos[p];
                    if (a > b || (a == b && pa > pb)) { skey[i] = b; skey[p] = a; spos[i] = pb; spos[p] = pa; const uint32_t t2 = sid[i]; sid[i] = sid[p]; sid[p] = t2; }
                }
            }
            __syncthreads();
            for (unsigned j = kk >> 2; j > 0; j >>= 1) {
                const int lj = ilog2_pow2(j);
                for (unsigned c = tid; c < half; c += RS_NT) {
                    const unsigned i = ((c >> lj) << (lj + 1)) + (c & (j - 1)), p = i + j;
                    if (p < ns) {
                        const ull a = skey[i], b = skey[p]; const uint32_t pa = spos[i], pb = spos[p];
                        if (a > b || (a == b && pa > pb)) { skey[i] = b; skey[p] = a; spos[i] = pb; spos[p] = pa; const uint32_t t2 = sid[i]; sid[i] = sid[p]; sid[p] = t2; }
                    }
                }
                __syncthreads();
            }
        }
        nbest = min(ns, k);
    }
    for (unsigned i = tid; i < k; i += RS_NT) {
        const bool ok = i < nbest;
        A.dist[q * k + i] = ok ? __longlong_as_double((long long)skey[i]) : __longlong_as_double(0x7ff0000000000000LL);
        A.ids[q * k + i] = ok ? sid[i] : 0xffffffffu;
    }
    if (tid == 0 && A.count) A.count[q] = (int32_t)nbest;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
#define RWS(h, var, type, slot, bytes)                                  \
    type* var = (type*)(h)->ws_get((slot), (bytes));                    \
    if (!var) return RPF_ERR_NOMEM;

static int rr_scan(rpf_handle* h, const uint32_t* in, int64_t n, ull* out, ull* bsum, ull* total) {
    const int nb = (int)((n + SC_NT * SC_PER - 1) / (SC_NT * SC_PER));
    RPF_LAUNCH(h, PH_Q_KNN, k_rr_scan_block, (unsigned)nb, SC_NT, 0, in, n, out, bsum);
    RPF_LAUNCH(h, PH_Q_KNN, k_rr_scan_top, 1, 1024, 0, bsum, nb, total);
    RPF_LAUNCH(h, PH_Q_KNN, k_rr_scan_add, (unsigned)nb, SC_NT, 0, out, n, bsum);
    return RPF_OK;
}

// Is the leaf-grouped path applicable / worthwhile for this call?  (d long enough that a query's gather dominates, enough
// queries per leaf for the regrouping to save traffic, leaves that fit the 64-row tile, DVector data, plain knn.)
bool rpf_rerank_gemm_wanted(const rpf_handle* h, int64_t nq, int k, int dedup) {
    if (h->rerank_gemm == 0 || dedup || h->d_xlast || (h->d & 3) || k > RS_SURV / 2 || h->n < 1) return false;
    if ((((uintptr_t)h->dX) & 15) != 0) return false;
    const Topology& tp = h->topo;
    int64_t nleaves = 0; uint32_t maxleaf = 0;
    for (int64_t g = 0; g < tp.nnodes(); ++g) if (tp.child[g] < 0) { ++nleaves; maxleaf = std::max(maxleaf, tp.size[g]); }
    if (maxleaf > RG_ROWS || nleaves < 1) return false;
    if (h->rerank_gemm == 2) return true;                       // forced (tests)
    return h->d >= 512 && nq >= 2 * nleaves;
}

// dQ / segs / cnt: the state run_descent left on the device.  Writes the local top-k lists to ddist / dids / dcount
// (device).  *fallback receives the number of queries left to the gather kernel; their ids are in WS_RR_FB.
int rpf_rerank_gemm(rpf_handle* h, const double* dQ, int64_t nq, int S, const uint32_t* segs, const uint32_t* cnt, int k,
                    double* ddist, uint32_t* dids, int32_t* dcount, uint32_t* n_fallback) {
    const Topology& tp = h->topo;
    const int64_t nn = tp.nnodes(), n = h->n;
    const int T = h->T, d = h->d;
    *n_fallback = 0;
    // leaf list of the current topology (cached on the device with the topology's epoch)
    std::vector<uint32_t> leaves;
    for (int64_t g = 0; g < nn; ++g) if (tp.child[g] < 0 && tp.size[g] > 0) leaves.push_back((uint32_t)g);
    if (leaves.empty()) return RPF_OK;
    RWS(h, d_leaves, uint32_t, WS_RR_LEAVES, leaves.size() * 4);
    RPF_CUDA(h, cudaMemcpyAsync(d_leaves, leaves.data(), leaves.size() * 4, cudaMemcpyHostToDevice, h->stream));
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));           // `leaves` is a local
    const size_t nkeys = (size_t)T * nn;
    RWS(h, xn, double, WS_RR_XN, (size_t)n * 8 + 16);
    RWS(h, qn, double, WS_RR_QN, (size_t)nq * 8 + 16);
    RWS(h, hist, uint32_t, WS_RR_HIST, nkeys * 8);            // hist + cursor
    RWS(h, start, ull, WS_RR_START, nkeys * 8);
    RWS(h, ccount, uint32_t, WS_RR_CC, (size_t)nq * 4);
    RWS(h, qoff, ull, WS_RR_QOFF, (size_t)nq * 8);
    const int nb_max = (int)((std::max<size_t>(nkeys, (size_t)nq) + SC_NT * SC_PER - 1) / (SC_NT * SC_PER));
    RWS(h, aux, ull, WS_RR_AUX, (size_t)(nb_max + 8) * 8);
    ull* xmax = aux + nb_max; ull* tot_c = xmax + 1; ull* tot_e = xmax + 2; uint32_t* fbc = (uint32_t*)(xmax + 3);
    RWS(h, fb, uint32_t, WS_RR_FB, (size_t)nq * 4);
    uint32_t* cursor = hist + nkeys;
    RPF_CUDA(h, cudaMemsetAsync(hist, 0, nkeys * 8, h->stream));
    RPF_CUDA(h, cudaMemsetAsync(xmax, 0, 32, h->stream));
    RRArgs A{};
    A.n = n; A.nq = nq; A.nn = nn; A.d = d; A.T = T; A.S = S; A.k = k;
    A.X = h->dX; A.Q = dQ; A.perm = h->d_perm; A.nstart = h->d_node_start; A.nsize = h->d_node_size; A.segs = segs; A.cnt = cnt;
    A.ccount = ccount; A.qoff = qoff; A.hist = hist; A.start = start; A.cursor = cursor;
    A.xn = xn; A.qn = qn; A.xmax = xmax; A.leaves = d_leaves; A.nleaves = (int)leaves.size();
    A.dist = ddist; A.ids = dids; A.count = dcount; A.fb_list = fb; A.fb_count = fbc;
    RPF_LAUNCH(h, PH_Q_KNN, k_rr_norms, (unsigned)((n + 7) / 8), 256, 0, h->dX, n, d, xn, xmax);
    RPF_LAUNCH(h, PH_Q_KNN, k_rr_norms, (unsigned)((nq + 7) / 8), 256, 0, dQ, nq, d, qn, (ull*)nullptr);
    RPF_LAUNCH(h, PH_Q_KNN, k_rr_group<false>, (unsigned)((nq + 7) / 8), 256, 0, A);
    int rc = rr_scan(h, ccount, nq, qoff, aux, tot_c);
    if (rc) return rc;
    rc = rr_scan(h, hist, (int64_t)nkeys, start, aux, tot_e);
    if (rc) return rc;
    ull tot[2];
    RPF_CUDA(h, cudaMemcpyAsync(tot, tot_c, 16, cudaMemcpyDeviceToHost, h->stream));
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    const ull ncand = tot[0], nent = tot[1];
    if (ncand == 0) {       // no candidates at all: every list is empty
        RPF_CUDA(h, cudaMemsetAsync(dcount, 0, (size_t)nq * 4, h->stream));
        return RPF_OK;
    }
    RWS(h, ent_q, uint32_t, WS_RR_ENTQ, (size_t)nent * 4 + 16);
    RWS(h, ent_dst, ull, WS_RR_ENTD, (size_t)nent * 8 + 16);
    RWS(h, dap, float, WS_RR_DAP, (size_t)ncand * 4 + 16);
    A.ent_q = ent_q; A.ent_dst = ent_dst; A.dap = dap;
    RPF_LAUNCH(h, PH_Q_KNN, k_rr_group<true>, (unsigned)((nq + 7) / 8), 256, 0, A);
    const size_t smem_g = (size_t)2 * (RG_ROWS + RG_MQ) * RG_LD * 8;
    RPF_CUDA(h, cudaFuncSetAttribute(k_rr_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
    RPF_LAUNCH(h, PH_Q_KNN, k_rr_gemm, dim3((unsigned)leaves.size(), (unsigned)T), RG_NT, smem_g, A);
    const size_t smem_s = (size_t)RS_CACHE * 4 + (size_t)RS_SURV * 16 + (size_t)d * 8 + ((size_t)T * S + 2) * 4 + 16;
    if (smem_s > 200 * 1024) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "rerank: dimension / leaf slots too large for the selection kernel");
    RPF_CUDA(h, cudaFuncSetAttribute(k_rr_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    RPF_LAUNCH(h, PH_Q_KNN, k_rr_select, (unsigned)nq, RS_NT, smem_s, A);
    RPF_CUDA(h, cudaMemcpyAsync(n_fallback, fbc, 4, cudaMemcpyDeviceToHost, h->stream));
    RPF_CUDA(h, cudaStreamSynchronize(h->stream));
    return RPF_OK;
}
