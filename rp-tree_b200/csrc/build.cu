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
#include "rpf_device.cuh"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <type_traits>

// =====================================================================================================
// K1  projections: key[h][i] = hp[h] . X[i]   (innerSD, right fold, no FMA)
// =====================================================================================================
// A tile of P = 32*R points is staged in shared memory (row stride ld = d|1 doubles -> conflict-free column
// gathers).  One warp owns one hyperplane at a time; every lane carries R points, so each (val, idx) pair is
// fetched once (one 16-byte uniform load from the packed CSR) and feeds R independent mul/add chains.
// Output row j corresponds to CSR row (t0 + j / L) * hpDepth + (j % L).
// ORD: write order-preserving uint64 keys and track the per-row min/max (bin ranges of the top phase).
template <int OFF>
__device__ __forceinline__ double lds_f64_off(unsigned a) {
    double x;
    asm("ld.shared.f64 %0, [%1+%2];" : "=d"(x) : "r"(a), "n"(OFF));
    return x;
}

// LD: compile-time row stride in doubles (0 = use the run-time ld): with it the R row addresses of a term are ONE add plus
// immediate offsets.  hp_pack holds (value, byte offset of the component inside a row = 8 * idx).
// Column blocks (rows too long for one tile, e.g. d = 768 / 960): a launch covers the columns [c0, c1) only.  The right fold
// runs from the LAST nonzero to the first, so the host launches the blocks from the highest columns down; every launch
// but the first resumes from the partial sum the previous one left in `out` (as a raw double), every launch but the
// last stores the partial sum back, the last one stores the finished key.  The arithmetic -- order and roundings -- is
// that of the single-tile kernel.  flags: bit 0 = first block of the fold (start from 0), bit 1 = last block.
// The fold of one staged tile: xs holds the P = 32 * R points i0 .. i0 + P - 1 (columns [c0, c1), row stride ld doubles).
template <int NT, int R, bool ORD, int LD>
__device__ __forceinline__ void project_fold(const double* xs, const int ld, const int64_t i0, const int64_t n, const int d,
                                             const int64_t* __restrict__ hp_off, const double2* __restrict__ hp_pack,
                                             const int t0, const int L, const int hpDepth, const int H,
                                             void* __restrict__ out, const int64_t ostride, ull* __restrict__ kmin, ull* __restrict__ kmax,
                                             const int c0, const int c1, const int flags, const bool track) {
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int dc = c1 - c0;
    const bool blocked = dc != d, first = flags & 1, last = flags & 2;
    // 32-bit shared-window addresses of this lane's R rows (one IADD + LDS per term instead of 64-bit pointer math)
    unsigned rowaddr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) rowaddr[r] = (unsigned)__cvta_generic_to_shared(xs + (size_t)(lane + 32 * r) * ld) - 8u * (unsigned)c0;
    // (the packed CSR stores byte offsets 8 * idx from column 0; the tile starts at column c0)
    // The per-row key range only seeds the bin map of the top phase (keys outside it fall into the first / last bin, the
    // exact select inside the median bin does the rest), so it is taken from every 8th tile: the warp reduction below
    // costs about as much as four terms of the dot product.  (`track`: chosen by the caller.)
    // write one output row (and fold its min/max) -- warp-uniform call
    auto emit = [&](int j, const double (&acc)[R]) {
        if (ORD && last) {
            ull vmin = ORD_NONE_HI, vmax = ORD_NONE_LO;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int64_t i = i0 + lane + 32 * r;
                if (i < n) {
                    const ull o = f2ord(acc[r]);
                    ((ull*)out)[(int64_t)j * ostride + i] = o;
                    vmin = o < vmin ? o : vmin;
                    vmax = o > vmax ? o : vmax;
                }
            }
            if (!track) return;
            for (int off = 16; off > 0; off >>= 1) {
                const ull a2 = __shfl_xor_sync(0xffffffffu, vmin, off), b2 = __shfl_xor_sync(0xffffffffu, vmax, off);
                vmin = a2 < vmin ? a2 : vmin;
                vmax = b2 > vmax ? b2 : vmax;
            }
            if (lane == 0 && vmin != ORD_NONE_HI) {
                if (vmin < kmin[j]) atomicMin(&kmin[j], vmin);
                if (vmax > kmax[j]) atomicMax(&kmax[j], vmax);
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int64_t i = i0 + lane + 32 * r;
                if (i < n) ((double*)out)[(int64_t)j * ostride + i] = acc[r];
            }
        }
    };
    // one term of one hyperplane for this lane's R points: acc = val * x[idx] + acc  (separate roundings, right fold)
    auto term = [&](const double2 hv, double (&acc)[R]) {
        const unsigned c = (unsigned)__double_as_longlong(hv.y);
        double x[R];
        if constexpr (LD != 0) {
            const unsigned a0 = rowaddr[0] + c;
            x[0] = lds_f64_off<0>(a0);
            if constexpr (R > 1) x[1] = lds_f64_off<1 * 32 * LD * 8>(a0);
            if constexpr (R > 2) x[2] = lds_f64_off<2 * 32 * LD * 8>(a0);
            if constexpr (R > 3) x[3] = lds_f64_off<3 * 32 * LD * 8>(a0);
            static_assert(R <= 4, "extend the immediate-offset loads");
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) asm("ld.shared.f64 %0, [%1];" : "=d"(x[r]) : "r"(rowaddr[r] + c));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = __dadd_rn(__dmul_rn(hv.x, x[r]), acc[r]);
    };
    // two hyperplanes per warp in flight (2R independent accumulate chains per lane)
    const unsigned magicL = 0xffffffffu / (unsigned)L + 1u;      // j / L = umulhi(j, magicL) for j * L < 2^32 (L > 1)
    auto csr_row = [&](int j) { const int q = L > 1 ? (int)__umulhi((unsigned)j, magicL) : j; return (t0 + q) * hpDepth + (j - q * L); };
    for (int j = w; j < H; j += 2 * NW) {          // warp-uniform
        const int jB = j + NW;
        const bool hasB = jB < H;
        const int rowA = csr_row(j);
        const int rowB = hasB ? csr_row(jB) : rowA;
        const int64_t sA = hp_off[rowA], sB = hp_off[rowB];
        int64_t eA = hp_off[rowA + 1], eB = hp_off[rowB + 1];
        if (eA - sA > d) eA = sA + d;               // innerSD's `i >= nz2` guard (Internal.hs:376)
        if (eB - sB > d) eB = sB + d;
        int64_t bA = sA, bB = sB;                   // nonzeros of this column block: [bA, eA) / [bB, eB)
        if (blocked) {
            // component offsets (8 * idx) ascend inside a row: cut [s, e) down to the offsets in [8 c0, 8 c1)
            auto lower = [&](int64_t lo, int64_t hi, long long key) {
                while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__double_as_longlong(__ldg(hp_pack + mid).y) < key) lo = mid + 1; else hi = mid; }
                return lo;
            };
            eA = lower(sA, eA, 8ll * c1); bA = lower(sA, eA, 8ll * c0);
            eB = lower(sB, eB, 8ll * c1); bB = lower(sB, eB, 8ll * c0);
        }
        int cA = (int)(eA - bA), cB = hasB ? (int)(eB - bB) : 0;
        const double2* hA = hp_pack + eA - 1;       // right fold: innermost (last) term first
        const double2* hB = hp_pack + eB - 1;
        double accA[R], accB[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { accA[r] = 0.0; accB[r] = 0.0; }
        if (!first) {                               // resume from the partial sums of the higher column blocks
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int64_t i = i0 + lane + 32 * r;
                if (i < n) {
                    accA[r] = ((const double*)out)[(int64_t)j * ostride + i];
                    if (hasB) accB[r] = ((const double*)out)[(int64_t)jB * ostride + i];
                }
            }
        }
        // the (value, offset) pair of the NEXT term is requested before the current term's loads and adds are issued
        double2 a = cA > 0 ? __ldg(hA) : make_double2(0.0, 0.0), b2 = cB > 0 ? __ldg(hB) : make_double2(0.0, 0.0);
        // two terms per stream and iteration: the consumed pair's registers are reloaded with the term after next
        // (no register rotation, half the loop control per term)
        while (cA > 2 && cB > 2) {
            const double2 a1 = __ldg(hA - 1), b1 = __ldg(hB - 1);
            term(a, accA);
            term(b2, accB);
            a = __ldg(hA - 2); b2 = __ldg(hB - 2);
            term(a1, accA);
            term(b1, accB);
            hA -= 2; hB -= 2; cA -= 2; cB -= 2;
        }
        while (cA > 1 && cB > 1) {
            const double2 an = __ldg(hA - 1), bn = __ldg(hB - 1);
            --hA; --hB; --cA; --cB;
            term(a, accA);
            term(b2, accB);
            a = an; b2 = bn;
        }
        while (cA > 1) { const double2 an = __ldg(hA - 1); --hA; --cA; term(a, accA); a = an; }
        while (cB > 1) { const double2 bn = __ldg(hB - 1); --hB; --cB; term(b2, accB); b2 = bn; }
        if (cA > 0) term(a, accA);
        if (cB > 0) term(b2, accB);
        emit(j, accA);
        if (hasB) emit(jB, accB);
    }
}

template <int NT, int R, bool ORD, int LD>
__global__ void __launch_bounds__(NT) k_project(const double* __restrict__ X, int64_t n, int d, int ld_rt,
                                                 const int64_t* __restrict__ hp_off, const double2* __restrict__ hp_pack,
                                                 int t0, int L, int hpDepth, int H,
                                                 void* __restrict__ out, int64_t ostride, ull* __restrict__ kmin, ull* __restrict__ kmax,
                                                 int c0, int c1, int flags, int pf_ahead) {
    constexpr int P = 32 * R, NW = NT / 32;
    extern __shared__ double xs[];
    const int ld = LD ? LD : ld_rt;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * P;
    const int rows = (int)min((int64_t)P, n - i0);
    const int dc = c1 - c0;
    if (pf_ahead > 0 && tid == 0) {
        // the tile a later wave will stage (its rows are contiguous in X): pulled into L2 now, so that CTA's load phase
        // -- during which its SM does nothing else -- is served from L2 instead of HBM
        const int64_t j0 = i0 + (int64_t)pf_ahead * P;
        if (j0 + P <= n) {
            const double* pf = X + j0 * (int64_t)d;
            const unsigned bytes = (unsigned)((size_t)P * d * 8);
            if ((((uintptr_t)pf) & 15) == 0 && (bytes & 15) == 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(pf), "r"(bytes) : "memory");
        }
    }
    for (int r = w; r < P; r += NW) {
        if (r < rows) {
            const double* src = X + (i0 + r) * (int64_t)d + c0;
            for (int c = lane; c < dc; c += 32) xs[r * ld + c] = src[c];
        } else {
            for (int c = lane; c < dc; c += 32) xs[r * ld + c] = 0.0;
        }
    }
    __syncthreads();
    project_fold<NT, R, ORD, LD>(xs, ld, i0, n, d, hp_off, hp_pack, t0, L, hpDepth, H, out, ostride, kmin, kmax, c0, c1, flags,
                                 (blockIdx.x & 7) == 0);
}

// Pipelined form for rows that fit one tile twice (d <= ~145 at 96 points, d <= ~107 at 128 points): persistent CTAs walk
// the tiles with a stride of gridDim.x; while the warps fold tile i out of one shared-memory buffer, the rows of tile i + 1
// stream into the other one with asynchronous 8-byte copies (cp.async -> LDGSTS; the odd row stride that keeps the column
// gathers conflict-free rules out 16-byte bulk copies), so the HBM read of X overlaps the fold instead of preceding it.
// At small tree groups (the shard of an 8-GPU run: a few dozen hyperplanes) the load IS most of the kernel.
template <int NT, int R, bool ORD, int LD>
__global__ void __launch_bounds__(NT, 1) k_project_pipe(const double* __restrict__ X, int64_t n, int d, int ld_rt,
                                                         const int64_t* __restrict__ hp_off, const double2* __restrict__ hp_pack,
                                                         int t0, int L, int hpDepth, int H,
                                                         void* __restrict__ out, int64_t ostride, ull* __restrict__ kmin, ull* __restrict__ kmax,
                                                         int64_t ntiles) {
    constexpr int P = 32 * R, NW = NT / 32;
    extern __shared__ double xs[];
    const int ld = LD ? LD : ld_rt;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const size_t bufsz = (size_t)P * ld;                       // doubles per buffer
    auto stage = [&](int64_t tile, int b) {                    // rows of `tile` -> buffer b (rows past n: zero fill)
        const int64_t i0 = tile * P;
        double* dst = xs + (size_t)b * bufsz;
        for (int r = w; r < P; r += NW) {
            const bool live = i0 + r < n;
            const double* src = X + (live ? (i0 + r) : 0) * (int64_t)d;
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + (size_t)r * ld);
            const unsigned nb = live ? 8u : 0u;                 // src-size 0: the 8 bytes are written as zeros
            for (int c = lane; c < d; c += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(sa + 8u * (unsigned)c), "l"(src + c), "r"(nb) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int64_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    stage(tile, 0);
    int b = 0;
    for (; tile < ntiles; tile += gridDim.x, b ^= 1) {
        const int64_t nxt = tile + gridDim.x;
        if (nxt < ntiles) {
            stage(nxt, b ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                                        // every thread's copies of this tile have landed
        project_fold<NT, R, ORD, LD>(xs + (size_t)b * bufsz, ld, tile * P, n, d, hp_off, hp_pack, t0, L, hpDepth, H, out, ostride,
                                     kmin, kmax, 0, d, 3, (tile & 7) == 0);
        __syncthreads();                                        // buffer b is free for the tile after next
    }
}

// Long rows (d > one tile: 768, 960, ...).  The column-blocked launches of k_project carry every partial sum through the key
// array (2 x 8 bytes per key and block: at 1M x 960 x 64 trees that is 100 GB next to 7.7 GB of X).  Here the partial sums
// never leave the registers: a warp OWNS up to HPW hyperplanes for the whole tile (HPW x R accumulators per lane), the tile's
// columns arrive in chunks of PW_KC from the HIGHEST columns down (the right fold's order) through a double-buffered
// cp.async pipeline that runs across tiles (persistent CTAs), and for every chunk the warp folds the chunk's terms of each of
// its hyperplanes onto the accumulator it already holds -- same order, same roundings as the single-tile kernel.
// ctab[row * (nch + 1) + c] = number of nonzeros of CSR row `row` with column < c * PW_KC (host-built, rpf_project_launch).
// A launch covers the H output rows jbase .. jbase + H - 1 (H <= NW * HPW); more hyperplanes take more launches (X re-read).
#define PW_KC 128
template <int NT, int R, int HPW, bool ORD>
__global__ void __launch_bounds__(NT, 1) k_project_wide(const double* __restrict__ X, int64_t n, int d,
                                                         const int64_t* __restrict__ hp_off, const double2* __restrict__ hp_pack,
                                                         const int32_t* __restrict__ ctab, int nch, int t0, int L, int hpDepth, int jbase, int H,
                                                         void* __restrict__ out, int64_t ostride, ull* __restrict__ kmin, ull* __restrict__ kmax,
                                                         int64_t ntiles) {
    constexpr int P = 32 * R, NW = NT / 32, LD = PW_KC + 1;
    static_assert(HPW % 2 == 0, "hyperplanes are folded in pairs");
    extern __shared__ double xs[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr size_t bufsz = (size_t)P * LD;
    auto stage = [&](int64_t tile, int cc, int b) {            // columns [cc * KC, ..) of the rows of `tile` -> buffer b
        const int64_t i0 = tile * P;
        const int c0 = cc * PW_KC, dc = min(PW_KC, d - c0);
        double* dst = xs + (size_t)b * bufsz;
        for (int r = w; r < P; r += NW) {
            const bool live = i0 + r < n;
            const double* src = X + (live ? (i0 + r) : 0) * (int64_t)d + c0;
            const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + (size_t)r * LD);
            const unsigned nb = live ? 8u : 0u;
            for (int c = lane; c < dc; c += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(sa + 8u * (unsigned)c), "l"(src + c), "r"(nb) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const unsigned magicL = 0xffffffffu / (unsigned)L + 1u;
    auto csr_row = [&](int j) { const int q = L > 1 ? (int)__umulhi((unsigned)j, magicL) : j; return (t0 + q) * hpDepth + (j - q * L); };
    int64_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    stage(tile, nch - 1, 0);
    int b = 0;
    for (; tile < ntiles; tile += gridDim.x) {
        double acc[HPW][R];
#pragma unroll
        for (int s2 = 0; s2 < HPW; ++s2)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[s2][r] = 0.0;
        for (int cc = nch - 1; cc >= 0; --cc, b ^= 1) {
            // next stage: the chunk below, or the top chunk of this CTA's next tile
            const bool more = cc > 0 || tile + gridDim.x < ntiles;
            if (more) {
                if (cc > 0) stage(tile, cc - 1, b ^ 1); else stage(tile + gridDim.x, nch - 1, b ^ 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();
            const unsigned a0 = (unsigned)__cvta_generic_to_shared(xs + (size_t)b * bufsz + (size_t)lane * LD) - 8u * (unsigned)(cc * PW_KC);
            auto term = [&](const double2 hv, double (&ac)[R]) {
                const unsigned a = a0 + (unsigned)__double_as_longlong(hv.y);
                double x[R];
                x[0] = lds_f64_off<0>(a);
                if constexpr (R > 1) x[1] = lds_f64_off<1 * 32 * LD * 8>(a);
                if constexpr (R > 2) x[2] = lds_f64_off<2 * 32 * LD * 8>(a);
                if constexpr (R > 3) x[3] = lds_f64_off<3 * 32 * LD * 8>(a);
                static_assert(R <= 4, "extend the immediate-offset loads");
#pragma unroll
                for (int r = 0; r < R; ++r) ac[r] = __dadd_rn(__dmul_rn(hv.x, x[r]), ac[r]);
            };
#pragma unroll
            for (int jj = 0; jj < HPW / 2; ++jj) {
                const int jA = jj * NW + w, jB = (jj + HPW / 2) * NW + w;        // output rows of this warp inside the pass
                if (jA >= H) continue;                                           // warp-uniform
                const bool hasB = jB < H;
                const int rowA = csr_row(jbase + jA), rowB = hasB ? csr_row(jbase + jB) : rowA;
                const int32_t* tA = ctab + (int64_t)rowA * (nch + 1) + cc;
                const int32_t* tB = ctab + (int64_t)rowB * (nch + 1) + cc;
                const int lA = __ldg(tA), lB = __ldg(tB);
                int cA = __ldg(tA + 1) - lA, cB = hasB ? __ldg(tB + 1) - lB : 0;
                const double2* hA = hp_pack + hp_off[rowA] + lA + cA - 1;       // right fold: last term of the chunk first
                const double2* hB = hp_pack + hp_off[rowB] + lB + cB - 1;
                double2 a = cA > 0 ? __ldg(hA) : make_double2(0.0, 0.0), b2 = cB > 0 ? __ldg(hB) : make_double2(0.0, 0.0);
                while (cA > 2 && cB > 2) {
                    const double2 a1 = __ldg(hA - 1), b1 = __ldg(hB - 1);
                    term(a, acc[jj]);
                    term(b2, acc[jj + HPW / 2]);
                    a = __ldg(hA - 2); b2 = __ldg(hB - 2);
                    term(a1, acc[jj]);
                    term(b1, acc[jj + HPW / 2]);
                    hA -= 2; hB -= 2; cA -= 2; cB -= 2;
                }
                while (cA > 1 && cB > 1) {
                    const double2 an = __ldg(hA - 1), bn = __ldg(hB - 1);
                    --hA; --hB; --cA; --cB;
                    term(a, acc[jj]);
                    term(b2, acc[jj + HPW / 2]);
                    a = an; b2 = bn;
                }
                while (cA > 1) { const double2 an = __ldg(hA - 1); --hA; --cA; term(a, acc[jj]); a = an; }
                while (cB > 1) { const double2 bn = __ldg(hB - 1); --hB; --cB; term(b2, acc[jj + HPW / 2]); b2 = bn; }
                if (cA > 0) term(a, acc[jj]);
                if (cB > 0) term(b2, acc[jj + HPW / 2]);
            }
            __syncthreads();                                    // buffer b is free for the stage after next
        }
        // ---- the finished keys of this tile
        const int64_t i0 = tile * P;
        const bool track = (tile & 7) == 0;
#pragma unroll
        for (int s2 = 0; s2 < HPW; ++s2) {
            const int j = s2 * NW + w;
            if (j >= H) continue;
            const int jo = jbase + j;
            if (ORD) {
                ull vmin = ORD_NONE_HI, vmax = ORD_NONE_LO;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int64_t i = i0 + lane + 32 * r;
                    if (i < n) {
                        const ull o = f2ord(acc[s2][r]);
                        ((ull*)out)[(int64_t)jo * ostride + i] = o;
                        vmin = o < vmin ? o : vmin;
                        vmax = o > vmax ? o : vmax;
                    }
                }
                if (track) {
                    for (int off = 16; off > 0; off >>= 1) {
                        const ull a2 = __shfl_xor_sync(0xffffffffu, vmin, off), b3 = __shfl_xor_sync(0xffffffffu, vmax, off);
                        vmin = a2 < vmin ? a2 : vmin;
                        vmax = b3 > vmax ? b3 : vmax;
                    }
                    if (lane == 0 && vmin != ORD_NONE_HI) {
                        if (vmin < kmin[jo]) atomicMin(&kmin[jo], vmin);
                        if (vmax > kmax[jo]) atomicMax(&kmax[jo], vmax);
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int64_t i = i0 + lane + 32 * r;
                    if (i < n) ((double*)out)[(int64_t)jo * ostride + i] = acc[s2][r];
                }
            }
        }
    }
}

// Fallback for very large d (tile does not fit shared memory): same arithmetic, rows read through L1/L2.
template <bool ORD>
__global__ void __launch_bounds__(256) k_project_direct(const double* __restrict__ X, int64_t n, int d,
                                                         const int64_t* __restrict__ hp_off, const double2* __restrict__ hp_pack,
                                                         int t0, int L, int hpDepth, int H,
                                                         void* __restrict__ out, int64_t ostride, ull* __restrict__ kmin, ull* __restrict__ kmax) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + lane;
    const bool live = i < n;
    const double* xr = X + (live ? i : 0) * (int64_t)d;
    for (int j = w; j < H; j += 8) {
        const int row = (t0 + j / L) * hpDepth + (j % L);
        const int64_t s = hp_off[row];
        int64_t e = hp_off[row + 1];
        if (e - s > d) e = s + d;
        double acc = 0.0;
        for (int64_t q = e - 1; q >= s; --q) {
            const double2 hv = __ldg(hp_pack + q);
            acc = __dadd_rn(__dmul_rn(hv.x, __ldg(xr + (int)(__double_as_longlong(hv.y) >> 3))), acc);
        }
        if (ORD) {
            const ull o = f2ord(acc);
            if (live) ((ull*)out)[(int64_t)j * ostride + i] = o;
            ull vmin = live ? o : ORD_NONE_HI, vmax = live ? o : ORD_NONE_LO;
            for (int off = 16; off > 0; off >>= 1) {
                const ull a2 = __shfl_xor_sync(0xffffffffu, vmin, off), b2 = __shfl_xor_sync(0xffffffffu, vmax, off);
                vmin = a2 < vmin ? a2 : vmin;
                vmax = b2 > vmax ? b2 : vmax;
            }
            if (lane == 0 && vmin != ORD_NONE_HI) {
                if (vmin < kmin[j]) atomicMin(&kmin[j], vmin);
                if (vmax > kmax[j]) atomicMax(&kmax[j], vmax);
            }
        } else if (live) {
            ((double*)out)[(int64_t)j * ostride + i] = acc;
        }
    }
}

// nblk column blocks of at most dcmax columns each, launched from the highest columns down (see k_project)
template <int NT, int R, bool ORD, int LD = 0>
static int launch_project(rpf_handle* h, int phase, const double* dX, int64_t n, int t0, int L, int H, void* out, int64_t ostride, ull* kmin, ull* kmax,
                          int nblk = 1) {
    const int d = h->d;
    const int dcmax = (d + nblk - 1) / nblk, ld = dcmax | 1;
    const size_t smem = (size_t)32 * R * ld * sizeof(double);
    auto kfn = k_project<NT, R, ORD, LD>;
    RPF_CUDA(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (n + 32 * R - 1) / (32 * R);
    for (int b = nblk - 1; b >= 0; --b) {
        const int c0 = b * dcmax, c1 = std::min(d, c0 + dcmax);
        if (c1 <= c0) continue;
        const int flags = (b == nblk - 1 || c1 == d ? 1 : 0) | (b == 0 ? 2 : 0);
        // whole-row tiles: prefetch the tile `resident CTAs` ahead into L2 (option project_prefetch: 0 = off)
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, NT, smem);
        const int pf = (nblk == 1 && h->project_prefetch) ? std::max(1, occ) * 148 : 0;
        RPF_LAUNCH(h, phase, kfn, (unsigned)grid, NT, smem, dX, n, d, ld, h->d_hp_off, (const double2*)h->d_hp_pack, t0, L, h->hpDepth, H, out, ostride,
                   kmin, kmax, c0, c1, flags, pf);
    }
    return RPF_OK;
}

template <int NT, int R, bool ORD, int LD = 0>
static int launch_project_pipe(rpf_handle* h, int phase, const double* dX, int64_t n, int t0, int L, int H, void* out, int64_t ostride, ull* kmin, ull* kmax) {
    const int d = h->d, ld = d | 1;
    const size_t smem = (size_t)2 * 32 * R * ld * sizeof(double);
    auto kfn = k_project_pipe<NT, R, ORD, LD>;
    RPF_CUDA(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (n + 32 * R - 1) / (32 * R);
    static int nsm = 0;
    if (!nsm) { cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device); if (nsm <= 0) nsm = 148; }
    const int64_t grid = std::min<int64_t>(ntiles, nsm);
    RPF_LAUNCH(h, phase, kfn, (unsigned)grid, NT, smem, dX, n, d, ld, h->d_hp_off, (const double2*)h->d_hp_pack, t0, L, h->hpDepth, H, out, ostride,
               kmin, kmax, ntiles);
    return RPF_OK;
}

// long rows: chunk table (per CSR row: nonzeros below every PW_KC column boundary), built once per hyperplane set and d
static int ensure_chunk_table(rpf_handle* h, int nch) {
    if (h->d_hp_chunk && h->hp_chunk_d == h->d && h->hp_chunk_rows == (int64_t)h->hp_off.size() - 1) return RPF_OK;
    if (h->capturing) return rpf_fail(h, RPF_ERR_STATE, "projection chunk table missing during graph capture");
    const int64_t rows = (int64_t)h->hp_off.size() - 1;
    std::vector<int32_t> tab((size_t)rows * (nch + 1));
    for (int64_t r = 0; r < rows; ++r) {
        int64_t q = h->hp_off[r];
        const int64_t e = std::min(h->hp_off[r + 1], h->hp_off[r] + (int64_t)h->d);     // innerSD's `i >= nz2` guard (Internal.hs:376)
        for (int c = 0; c <= nch; ++c) {
            while (q < e && h->hp_idx[q] < c * PW_KC) ++q;
            tab[(size_t)r * (nch + 1) + c] = (int32_t)((c == nch ? e : q) - h->hp_off[r]);
        }
    }
    if (h->d_hp_chunk) { cudaStreamSynchronize(h->stream); cudaFree(h->d_hp_chunk); h->d_hp_chunk = nullptr; }
    RPF_CUDA(h, cudaMalloc(&h->d_hp_chunk, std::max<size_t>(tab.size() * 4, 16)));
    RPF_CUDA(h, cudaMemcpy(h->d_hp_chunk, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    h->hp_chunk_d = h->d; h->hp_chunk_rows = rows;
    return RPF_OK;
}

template <int NT, int R, int HPW, bool ORD>
static int launch_project_wide(rpf_handle* h, int phase, const double* dX, int64_t n, int t0, int L, int H, void* out, int64_t ostride, ull* kmin, ull* kmax) {
    const int d = h->d, nch = (d + PW_KC - 1) / PW_KC;
    int rc = ensure_chunk_table(h, nch);
    if (rc) return rc;
    const size_t smem = (size_t)2 * 32 * R * (PW_KC + 1) * sizeof(double);
    auto kfn = k_project_wide<NT, R, HPW, ORD>;
    RPF_CUDA(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (n + 32 * R - 1) / (32 * R);
    const int64_t grid = std::min<int64_t>(ntiles, 148);
    const int per = (NT / 32) * HPW;
    // balanced passes (e.g. 896 hyperplanes at 288 per pass: 4 x 224 instead of 3 x 288 + 32)
    const int npass = (H + per - 1) / per, hp = (H + npass - 1) / npass;
    for (int j0 = 0; j0 < H; j0 += hp)
        RPF_LAUNCH(h, phase, kfn, (unsigned)grid, NT, smem, dX, n, d, h->d_hp_off, (const double2*)h->d_hp_pack, (const int32_t*)h->d_hp_chunk, nch,
                   t0, L, h->hpDepth, j0, std::min(hp, H - j0), out, ostride, kmin, kmax, ntiles);
    return RPF_OK;
}

// =====================================================================================================
// K1t  projections, transposed tile + fold program (the default for rows that fit one tile)
// =====================================================================================================
// What bounds this fold is shared-memory bandwidth: 8 bytes per (point, term) cannot be avoided (the sparsity pattern is
// run-time data, so the x values cannot sit in registers) -- 1.23 ms of the LDS pipe at configs[1].  k_project spends 35 warp
// instructions per 128-point term on it (ncu: 1.57e9 instructions for 4.5e7 warp-terms) although its inner loop has 17.5:
// the rest is per-hyperplane control (CSR offsets, tail loops for unequal pair lengths, 64-bit address arithmetic and
// bounds checks in the store of the keys), paid once per ~13 terms.  Here
//   * the tile is stored TRANSPOSED, xs[column][point] (row of P points = 8P bytes, the 16-byte point pairs XOR-swizzled
//     with the column's low bits so the transposing stores of the loader are conflict-free too): the x values of a term
//     arrive as 16-byte loads of two points each instead of 8-byte ones, and no padding column is needed (128 x 128
//     doubles = 128 KB exactly: the carve-out leaves 124 KB of L1 for the program);
//   * the host turns the hyperplanes of a launch into a FOLD PROGRAM: output rows sorted by their number of nonzeros and
//     paired; both streams of a pair padded to the same even length with (0.0, column 0) terms -- exact: the fold's
//     accumulator is never -0 (it starts at +0 and RN addition gives -0 only for (-0) + (-0)), so adding 0.0 * x = +-0
//     leaves every bit of it unchanged for finite x -- and stored interleaved, 48 bytes per iteration (2 terms of each
//     stream: four values + four packed smem offsets), so the inner loop has no tails, no per-stream pointer and 3 uniform
//     loads per 4 terms; the pairs are dealt to the warps in serpentine order (equal work per warp, no run-time balancing);
//   * keys are stored as 16-byte pairs from a precomputed row pointer; bounds checks only in the last tile.
// Arithmetic: acc = fl(fl(val * x) + acc) from the LAST nonzero to the first, as innerSD (Internal.hs:369-382) -- the same
// operations in the same order as k_project; the results are bit-identical (tests: both kernels against the oracle).
struct ProjProg {
    int t0 = 0, H = 0, L = 0, hpDepth = 0, d = 0, P = 0, NW = 0, c0 = 0, c1 = 0;       // columns [c0, c1) of the rows (whole rows: 0, d)
    int jpw = 0;                        // jobs (pairs of output rows) per warp
    int4* d_jobs = nullptr;             // [NW][jpw]: (first 16-byte unit of the pair's terms, iterations, row A, row B); row < 0: none
    uint4* d_terms = nullptr;
    size_t term_bytes = 0;
    ~ProjProg() { if (d_jobs) cudaFree(d_jobs); if (d_terms) cudaFree(d_terms); }
};
struct ProjProgCache { std::vector<ProjProg*> v; ~ProjProgCache() { for (auto* p : v) delete p; } };
static void free_proj_progs(void* p) { delete (ProjProgCache*)p; }

static int get_proj_prog(rpf_handle* h, int t0, int H, int L, int P, int NW, int c0, int c1, ProjProg** out) {
    if (!h->proj_progs) { h->proj_progs = new ProjProgCache(); h->proj_progs_free = free_proj_progs; }
    ProjProgCache* C = (ProjProgCache*)h->proj_progs;
    for (ProjProg* q : C->v)
        if (q->t0 == t0 && q->H == H && q->L == L && q->hpDepth == h->hpDepth && q->d == h->d && q->P == P && q->NW == NW && q->c0 == c0 && q->c1 == c1) { *out = q; return RPF_OK; }
    if (h->capturing) return rpf_fail(h, RPF_ERR_STATE, "projection program missing during graph capture");
    if (C->v.size() >= 256) { cudaStreamSynchronize(h->stream); for (auto* q : C->v) delete q; C->v.clear(); }
    const int d = h->d;
    struct Row { int j; int64_t s; int cnt; };
    std::vector<Row> rows((size_t)H);
    for (int j = 0; j < H; ++j) {
        const int64_t r = (int64_t)(t0 + j / L) * h->hpDepth + (j % L);
        int64_t s0 = h->hp_off[r], e0 = std::min(h->hp_off[r + 1], s0 + (int64_t)d);      // innerSD's `i >= nz2` guard (Internal.hs:376)
        // column block: the nonzeros with c0 <= column < c1 (the indices ascend inside a row)
        while (s0 < e0 && h->hp_idx[s0] < c0) ++s0;
        while (e0 > s0 && h->hp_idx[e0 - 1] >= c1) --e0;
        rows[j] = Row{j, s0, (int)(e0 - s0)};
    }
    std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) { return a.cnt > b.cnt; });
    const int npairs = (H + 1) / 2;
    const int jpw = (npairs + NW - 1) / NW;
    std::vector<int4> jobs((size_t)NW * jpw, make_int4(0, 0, -1, -1));
    std::vector<uint4> terms;
    const unsigned rowb = (unsigned)P * 8u;
    auto enc = [&](int c) { return (unsigned)c * rowb + (((unsigned)c & 7u) << 4); };
    for (int q = 0; q < npairs; ++q) {
        const Row& A = rows[2 * q];
        const bool hasB = 2 * q + 1 < H;
        const Row B = hasB ? rows[2 * q + 1] : Row{-1, 0, 0};
        const int niter = (std::max(A.cnt, B.cnt) + 1) / 2;
        const int g = q / NW, i = q % NW, w = (g & 1) ? NW - 1 - i : i;            // serpentine deal
        jobs[(size_t)w * jpw + g] = make_int4((int)terms.size(), niter, A.j, B.j);
        auto val = [&](const Row& R, int k) { return k < R.cnt ? h->hp_val[R.s + R.cnt - 1 - k] : 0.0; };     // fold order: last nonzero first
        auto rec = [&](const Row& R, int k) { return k < R.cnt ? enc(h->hp_idx[R.s + R.cnt - 1 - k] - c0) : enc(0); };
        for (int it = 0; it < niter; ++it) {
            // 48 bytes per iteration: [A_k0, A_k1] [B_k0, B_k1] [recA_k0, recA_k1] [recB_k0, recB_k1] -- a lane of half-warp s reads
            // its stream's 16 bytes of values at +16 s and its 8 bytes of offsets at +32 + 8 s
            const int k0 = 2 * it, k1 = k0 + 1;
            const double v4[4] = {val(A, k0), val(A, k1), val(B, k0), val(B, k1)};
            uint4 u0, u1;
            std::memcpy(&u0, &v4[0], 16); std::memcpy(&u1, &v4[2], 16);
            terms.push_back(u0); terms.push_back(u1);
            terms.push_back(make_uint4(rec(A, k0), rec(A, k1), rec(B, k0), rec(B, k1)));
        }
    }
    ProjProg* Q = new ProjProg();
    Q->t0 = t0; Q->H = H; Q->L = L; Q->hpDepth = h->hpDepth; Q->d = d; Q->P = P; Q->NW = NW; Q->jpw = jpw; Q->c0 = c0; Q->c1 = c1;
    Q->term_bytes = terms.size() * sizeof(uint4);
    if (cudaMalloc(&Q->d_jobs, jobs.size() * sizeof(int4)) != cudaSuccess || cudaMalloc(&Q->d_terms, std::max<size_t>(Q->term_bytes, 64)) != cudaSuccess ||
        cudaMemcpy(Q->d_jobs, jobs.data(), jobs.size() * sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess ||
        (Q->term_bytes && cudaMemcpy(Q->d_terms, terms.data(), Q->term_bytes, cudaMemcpyHostToDevice) != cudaSuccess)) {
        cudaGetLastError(); delete Q;
        return rpf_fail(h, RPF_ERR_NOMEM, "projection program: device tables");
    }
    C->v.push_back(Q);
    *out = Q;
    return RPF_OK;
}

__device__ __forceinline__ double2 lds_f64x2(unsigned a) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ double ldg_stream_f64(const double* p) {
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <int NT, int R, bool ORD, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_project_t(const double* __restrict__ X, int64_t n, int d,
                                                         const int4* __restrict__ jobs, int jpw, const uint4* __restrict__ terms,
                                                         void* __restrict__ out, int64_t ostride, ull* __restrict__ kmin, ull* __restrict__ kmax,
                                                         int pf_ahead, int c0, int dc, int flags) {
    // Column blocks (rows too long for one tile: d = 768, 960, ...): a launch covers the columns [c0, c0 + dc) only.  The right
    // fold runs from the LAST nonzero to the first, so the host launches the blocks from the highest columns down; every launch
    // but the first (flags bit 0) resumes from the partial sums the previous one left in `out` as raw doubles, every launch but
    // the last (bit 1) stores them back; the last one stores the finished keys.  Same operations, same order, same roundings.
    static_assert(R == 2 || R == 4, "tile of 64 or 128 points");
    const bool first = flags & 1, last = flags & 2;
    constexpr int P = 32 * R, NW = NT / 32;
    constexpr int K = R;                                           // 16-byte point pairs per lane: a half-warp covers the tile
    constexpr unsigned ROWB = P * 8;                               // bytes of one tile row (one column, P points)
    extern __shared__ double xs_raw[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // the swizzle XORs address bits 4..6: the tile base must not carry into them
    const unsigned base = ((unsigned)__cvta_generic_to_shared(xs_raw) + 511u) & ~511u;
    const int64_t i0 = (int64_t)blockIdx.x * P;
    if (pf_ahead > 0 && dc == d && tid == 0) {
        // the tile a later wave will stage (its rows are contiguous in X) is pulled into L2 now
        const int64_t j0 = i0 + (int64_t)pf_ahead * P;
        if (j0 + P <= n) {
            const double* pf = X + j0 * (int64_t)d;
            const unsigned bytes = (unsigned)((size_t)P * d * 8);
            if ((((uintptr_t)pf) & 15) == 0 && (bytes & 15) == 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(pf), "r"(bytes) : "memory");
        }
    }
    // ---- stage the tile transposed: rows (2q, 2q+1) of X -> one 16-byte pair per column
    for (int q = w; q < P / 2; q += NW) {
        const int64_t r0 = i0 + 2 * q;
        const bool l0 = r0 < n, l1 = r0 + 1 < n;
        const double* s0 = X + (l0 ? r0 : 0) * (int64_t)d + c0;
        const double* s1 = X + (l1 ? r0 + 1 : 0) * (int64_t)d + c0;
        for (int cb = lane; cb < dc; cb += 128) {
            double v0[4], v1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = cb + 32 * u;
                v0[u] = (c < dc && l0) ? ldg_stream_f64(s0 + c) : 0.0;
                v1[u] = (c < dc && l1) ? ldg_stream_f64(s1 + c) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = cb + 32 * u;
                if (c < dc) {
                    const unsigned a = base + (unsigned)c * ROWB + ((unsigned)(q ^ (c & 7)) << 4);
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a), "d"(v0[u]), "d"(v1[u]) : "memory");
                }
            }
        }
    }
    __syncthreads();

    // Half-warp s = lane >> 4 folds stream s of the warp's current pair of output rows; lane m = lane & 15 of it carries the
    // point pairs m, m + 16, ... (points 2m, 2m+1, 2m+32, 2m+33, ...).  The (value, offset) words of a term are then fetched
    // by 16 lanes, not 32: the LSU writes back lanes x bytes whatever the address pattern, so a warp-uniform 16-byte load
    // costs four 128-byte wavefronts of the same data pipe the x gathers saturate (ncu of the first version, every lane
    // loading both streams' terms: 3 wavefronts per term next to the 8 of the x values, pipe 92 % busy).
    const int sgrp = lane >> 4, m = lane & 15;
    const unsigned m16 = (unsigned)m << 4;
    const bool full = i0 + P <= n;
    const bool al16 = ((((uintptr_t)out) | ((uintptr_t)ostride << 3)) & 15) == 0;
    const bool track = ORD && last && (blockIdx.x & 7) == 0;       // the key range only seeds bin maps: every 8th tile is enough
    const int4* jw = jobs + (size_t)w * jpw;
    for (int k = 0; k < jpw; ++k) {                                // warp-uniform
        const int4 jb = __ldg(jw + k);
        if (jb.z < 0) break;
        double acc[2 * K];
#pragma unroll
        for (int r = 0; r < 2 * K; ++r) acc[r] = 0.0;
        const int j = sgrp ? jb.w : jb.z;                                      // this half-warp's output row (< 0: none)
        if (!first && j >= 0) {                                                // resume from the partial sums of the higher column blocks
            const double* row = (const double*)out + (int64_t)j * ostride + i0 + 2 * m;
#pragma unroll
            for (int r = 0; r < 2 * K; ++r) {
                const int off = (r >> 1) * 32 + (r & 1);
                if (i0 + 2 * m + off < n) acc[r] = row[off];
            }
        }
        auto term = [&](const double v, const unsigned rec) {
            const unsigned a = (rec ^ m16) + base;
            double2 x[K];
#pragma unroll
            for (int q = 0; q < K; ++q) x[q] = lds_f64x2(a + 256u * q);
#pragma unroll
            for (int q = 0; q < K; ++q) {
                acc[2 * q] = __dadd_rn(__dmul_rn(v, x[q].x), acc[2 * q]);
                acc[2 * q + 1] = __dadd_rn(__dmul_rn(v, x[q].y), acc[2 * q + 1]);
            }
        };
        int it = jb.y;
        if (it > 0) {
            const char* tp = (const char*)(terms + jb.x) + 16 * sgrp;      // this half-warp's values; its offsets: + 32 - 8 s
            const int ro = 32 - 8 * sgrp;
            double2 v = __ldg((const double2*)tp);
            uint2 rc = __ldg((const uint2*)(tp + ro));
            while (true) {
                // the next iteration's words are requested before this one's loads and adds are issued
                const char* tn = tp + (it > 1 ? 48 : 0);
                const double2 nv = __ldg((const double2*)tn);
                const uint2 nr = __ldg((const uint2*)(tn + ro));
                term(v.x, rc.x);
                term(v.y, rc.y);
                if (--it == 0) break;
                tp = tn; v = nv; rc = nr;
            }
        }
        // ---- this half-warp's output row: points i0 + 2m + 32q (+1)
        if (j < 0) continue;                                           // odd row count: the last pair has no second stream
        if (ORD && last) {
            ull o[2 * K];
#pragma unroll
            for (int r = 0; r < 2 * K; ++r) o[r] = f2ord(acc[r]);
            ull* row = (ull*)out + (int64_t)j * ostride + i0 + 2 * m;
            if (full && al16) {
#pragma unroll
                for (int q = 0; q < K; ++q) *reinterpret_cast<ulonglong2*>(row + 32 * q) = make_ulonglong2(o[2 * q], o[2 * q + 1]);
            } else {
#pragma unroll
                for (int r = 0; r < 2 * K; ++r) {
                    const int off = (r >> 1) * 32 + (r & 1);
                    if (i0 + 2 * m + off < n) row[off] = o[r];
                }
            }
            if (track) {
                ull vmin = ORD_NONE_HI, vmax = ORD_NONE_LO;
#pragma unroll
                for (int r = 0; r < 2 * K; ++r) {
                    const int off = (r >> 1) * 32 + (r & 1);
                    if (i0 + 2 * m + off < n) { vmin = o[r] < vmin ? o[r] : vmin; vmax = o[r] > vmax ? o[r] : vmax; }
                }
                const unsigned hm = sgrp ? 0xffff0000u : 0x0000ffffu;
                for (int off = 8; off > 0; off >>= 1) {
                    const ull a2 = __shfl_xor_sync(hm, vmin, off), b2 = __shfl_xor_sync(hm, vmax, off);
                    vmin = a2 < vmin ? a2 : vmin;
                    vmax = b2 > vmax ? b2 : vmax;
                }
                if (m == 0 && vmin != ORD_NONE_HI) {
                    if (vmin < kmin[j]) atomicMin(&kmin[j], vmin);
                    if (vmax > kmax[j]) atomicMax(&kmax[j], vmax);
                }
            }
        } else {
            double* row = (double*)out + (int64_t)j * ostride + i0 + 2 * m;
            if (full && al16) {
#pragma unroll
                for (int q = 0; q < K; ++q) *reinterpret_cast<double2*>(row + 32 * q) = make_double2(acc[2 * q], acc[2 * q + 1]);
            } else {
#pragma unroll
                for (int r = 0; r < 2 * K; ++r) {
                    const int off = (r >> 1) * 32 + (r & 1);
                    if (i0 + 2 * m + off < n) row[off] = acc[r];
                }
            }
        }
    }
}

template <int NT, int R, bool ORD, int MINB>
static int launch_project_t(rpf_handle* h, int phase, const double* dX, int64_t n, int t0, int L, int H, void* out, int64_t ostride, ull* kmin, ull* kmax,
                            int nblk = 1) {
    constexpr int P = 32 * R;
    const int d = h->d, dcmax = (d + nblk - 1) / nblk;
    const size_t smem = (size_t)dcmax * P * 8 + 512;
    auto kfn = k_project_t<NT, R, ORD, MINB>;
    RPF_CUDA(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (n + P - 1) / P;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, NT, smem);
    const int pf = (h->project_prefetch && nblk == 1) ? std::max(1, occ) * 148 : 0;
    for (int b = nblk - 1; b >= 0; --b) {                          // highest columns first: the right fold's order
        const int c0 = b * dcmax, c1 = std::min(d, c0 + dcmax);
        if (c1 <= c0) continue;
        ProjProg* Q = nullptr;
        int rc = get_proj_prog(h, t0, H, L, P, NT / 32, c0, c1, &Q);
        if (rc) return rc;
        const int flags = (b == nblk - 1 || c1 == d ? 1 : 0) | (b == 0 ? 2 : 0);
        RPF_LAUNCH(h, phase, kfn, (unsigned)grid, NT, smem, dX, n, d, (const int4*)Q->d_jobs, Q->jpw, (const uint4*)Q->d_terms, out, ostride, kmin, kmax, pf,
                   c0, c1 - c0, flags);
    }
    return RPF_OK;
}

// out row j (= tree-in-group * L + level) starts at out + j * ostride; point i of dX lands at column i
int rpf_project_launch(rpf_handle* h, int phase, const double* dX, int64_t n, int t0, int Tg, int L, bool ord, void* out,
                       int64_t ostride, ull* kmin, ull* kmax) {
    const int d = h->d, ld = d | 1;
    const int H = Tg * L;
    if (n <= 0 || H <= 0) return RPF_OK;
    const size_t row = (size_t)ld * 8;
    // transposed tile + fold program: whole rows in one tile (variant 7 / 8 and 1..6: the earlier kernels, kept as test hooks)
    if (h->project_variant == 0 || h->project_variant >= 10) {
        const int v = h->project_variant;
        const bool fitA = (size_t)d * 1024 + 512 <= 164 * 1024, fitB = (size_t)d * 512 + 512 <= 200 * 1024;
        const int pick = v >= 10 ? v : (fitA && H > h->project_pipe_maxh ? 10 : (fitB ? 11 : 0));
        if (pick == 10 && fitA)
            return ord ? launch_project_t<1024, 4, true, 1>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax)
                       : launch_project_t<1024, 4, false, 1>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax);
        if (pick == 11 && fitB)
            return ord ? launch_project_t<512, 2, true, 2>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax)
                       : launch_project_t<512, 2, false, 2>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax);
        if (pick == 12 && fitB)
            return ord ? launch_project_t<256, 2, true, 3>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax)
                       : launch_project_t<256, 2, false, 3>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax);
        if ((pick == 0 || pick == 13) && !fitB) {
            // long rows (d = 768, 960, ...): 128-point tiles over column blocks of <= 160 columns, partial sums carried in `out`
            const int nblk = (d + 159) / 160;
            if (nblk <= 64)
                return ord ? launch_project_t<1024, 4, true, 1>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax, nblk)
                           : launch_project_t<1024, 4, false, 1>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax, nblk);
        }
    }
    if (h->project_variant == 1 && 64 * row <= 110 * 1024)
        return ord ? launch_project<1024, 2, true>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax)
                   : launch_project<1024, 2, false>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax);
    if (h->project_variant == 2 && 128 * row <= 140 * 1024)
        return ord ? launch_project<512, 4, true>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax)
                   : launch_project<512, 4, false>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax);
    // rows short enough for two tiles in shared memory: the pipelined kernel (variant 7 = the single-buffer kernels below)
    // (measured at 1M x 128: at 448 hyperplanes the two 96-point buffers leave 30 KB of L1 for the 92 KB of packed
    //  hyperplanes and the fold slows down more than the overlap gains -- 2.92 vs 2.54 ms; with few hyperplanes, e.g. the
    //  4-tree shard of an 8-GPU run, the load dominates and the pipeline wins)
    const bool want_pipe = h->project_variant == 8 || (h->project_variant == 0 && H <= h->project_pipe_maxh);
    if (want_pipe && n >= 4096) {
        const size_t lim = 225 * 1024;
        if (ld == 129)
            return ord ? launch_project_pipe<1024, 3, true, 129>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax)
                       : launch_project_pipe<1024, 3, false, 129>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax);
        if ((size_t)2 * 128 * row <= lim)
            return ord ? launch_project_pipe<1024, 4, true>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax)
                       : launch_project_pipe<1024, 4, false>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax);
        if ((size_t)2 * 96 * row <= lim)
            return ord ? launch_project_pipe<1024, 3, true>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax)
                       : launch_project_pipe<1024, 3, false>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax);
    }
    if (ld == 129)      // d = 128: compile-time row stride
        return ord ? launch_project<1024, 4, true, 129>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax)
                   : launch_project<1024, 4, false, 129>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax);
    // long rows: register-accumulator kernel -- tuning hook only (variant 5).  Measured at 1M x 960 x 64 trees: 82.7 ms against
    // 57.5 ms for the column-blocked launches below: with 2 points per lane and 16 warps per SM it keeps half as many
    // accumulate chains in flight, and the shared-memory gathers -- the real bound of this fold -- are hidden less well;
    // the 100 GB of partial sums the blocked launches move through HBM cost less than that.
    if (h->project_variant == 5 && (size_t)64 * row > 140 * 1024 && n >= 2048)
        return ord ? launch_project_wide<512, 2, 18, true>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax)
                   : launch_project_wide<512, 2, 18, false>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax);
    if (h->project_variant != 3) {
        // 128-point tiles (4 points per lane); rows too long for one tile (d > 139: 200, 768, 960, ...) go in column
        // blocks of <= 139 columns with the partial sums carried in `out`.  Tuning hooks: variant 4 / 6 = 32- / 64-point
        // tiles (fewer, wider blocks) -- measured slower: 4 chains per lane matter more than the extra passes over `out`.
        // (measured, 400 k points x 16 trees: d = 200: one 64-point tile 1.12 ms vs 128 x 2 blocks 1.62; d = 500: 128 x 4 blocks
        //  3.85 ms vs 64 x 2 blocks 6.24; d = 960: 128 x 8 blocks 3.2 ms vs 64 x 4 blocks 6.0)
        const bool one64 = (size_t)64 * row <= 140 * 1024;      // a single 64-point tile beats two passes of 128-point tiles
        const int pts = h->project_variant == 4 ? 32 : ((h->project_variant == 6 || (h->project_variant == 0 && one64)) ? 64 : 128);
        int nblk = 1;
        while ((size_t)pts * ((size_t)((d + nblk - 1) / nblk) | 1) * 8 > 140 * 1024) ++nblk;
        if (nblk <= 128) {
            if (pts == 32)
                return ord ? launch_project<256, 1, true>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax, nblk)
                           : launch_project<256, 1, false>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax, nblk);
            if (pts == 64)
                return ord ? launch_project<512, 2, true>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax, nblk)
                           : launch_project<512, 2, false>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax, nblk);
            return ord ? launch_project<1024, 4, true>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax, nblk)
                       : launch_project<1024, 4, false>(h, phase, dX, n, t0, L, H, out, ostride, kmin, kmax, nblk);
        }
    }
    const int64_t grid = (n + 31) / 32;
    if (ord) {
        RPF_LAUNCH(h, phase, k_project_direct<true>, (unsigned)grid, 256, 0, dX, n, d, h->d_hp_off, (const double2*)h->d_hp_pack, t0, L, h->hpDepth, H, out, ostride, kmin, kmax);
    } else {
        RPF_LAUNCH(h, phase, k_project_direct<false>, (unsigned)grid, 256, 0, dX, n, d, h->d_hp_off, (const double2*)h->d_hp_pack, t0, L, h->hpDepth, H, out, ostride, kmin, kmax);
    }
    return RPF_OK;
}

// =====================================================================================================
// shared bitonic helpers (normalised network: every comparator puts the smaller element at the lower
// index, so positions >= m behave as +inf padding and comparators touching them are skipped)
// =====================================================================================================

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

// ---- register-blocked bitonic network --------------------------------------------------------------------------
// Thread t owns the 8 consecutive slots [8t, 8t+8) in registers.  Comparator distances 1,2,4 are register-to-register,
// 8..128 are lane-to-lane shuffles inside the warp, >= 256 go through shared memory in a transposed layout
// (slot x lives at (x & 7) * NT + (x >> 3), so consecutive threads touch consecutive words: no bank conflicts).
template <typename W>
__device__ __forceinline__ void ce_min_first(W& a, W& b) { if (a > b) { const W t = a; a = b; b = t; } }

template <int K, typename W>
__device__ __forceinline__ void flip_in(W (&v)[8]) {      // mirror pairs inside blocks of K <= 8 registers
#pragma unroll
    for (int e = 0; e < 8; ++e) { const int p = e ^ (K - 1); if (e < p) ce_min_first(v[e], v[p]); }
}
template <int J, typename W>
__device__ __forceinline__ void half_in(W (&v)[8]) {      // pairs (e, e+J), J in {1,2,4}
#pragma unroll
    for (int e = 0; e < 8; ++e) if ((e & J) == 0) ce_min_first(v[e], v[e + J]);
}
template <typename W>
__device__ __forceinline__ void flip_shfl(W (&v)[8], unsigned g, unsigned lane) {   // block of g lanes (8g slots)
    const bool lower = (lane & (g >> 1)) == 0;
    W o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = __shfl_xor_sync(0xffffffffu, v[7 - e], g - 1);
#pragma unroll
    for (int e = 0; e < 8; ++e) { const bool lt = v[e] < o[e]; v[e] = (lt == lower) ? v[e] : o[e]; }
}
template <typename W>
__device__ __forceinline__ void half_shfl(W (&v)[8], unsigned jl, unsigned lane) {  // partner lane = lane ^ jl
    const bool lower = (lane & jl) == 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const W o = __shfl_xor_sync(0xffffffffu, v[e], jl);
        const bool lt = v[e] < o;
        v[e] = (lt == lower) ? v[e] : o;
    }
}
template <typename W>
__device__ __forceinline__ void tail_in(W (&v)[8]) { half_in<4>(v); half_in<2>(v); half_in<1>(v); }

// sorts every aligned block of Pv slots (Pv a power of two, 2 <= Pv <= 8*NT) ascending; all threads must call.
// W = uint64 or uint32 sort words (the 32-bit form halves the compare / select / shuffle / shared-memory work).
template <int NT, typename W>
__device__ void sort_regs(W (&v)[8], W* w, unsigned Pv) {
    const unsigned tid = threadIdx.x, lane = tid & 31;
    flip_in<2>(v);
    if (Pv >= 4) { flip_in<4>(v); half_in<1>(v); }
    if (Pv >= 8) { flip_in<8>(v); half_in<2>(v); half_in<1>(v); }
    for (unsigned g = 2; g <= 32 && 8 * g <= Pv; g <<= 1) {           // k = 8g = 16 .. 256: inside one warp
        flip_shfl(v, g, lane);
        for (unsigned jl = g >> 2; jl >= 1; jl >>= 1) half_shfl(v, jl, lane);
        tail_in(v);
    }
    if (NT > 32) {
        for (unsigned k = 512; k <= Pv; k <<= 1) {                    // cross-warp merges
            const unsigned kt = k >> 3;                                // block size in threads
            const int lkt = ilog2_pow2(kt);
#pragma unroll
            for (int e = 0; e < 8; ++e) w[e * NT + tid] = v[e];
            __syncthreads();
            // flip: slot (e, u) with u in the lower half of its kt-block <-> (7-e, mirrored u)
            for (unsigned c = tid; c < 4u * NT; c += NT) {
                const unsigned e = c / (NT / 2), cc = c % (NT / 2);
                const unsigned blk = cc >> (lkt - 1), uo = cc & ((kt >> 1) - 1);
                const unsigned u = (blk << lkt) + uo, up = (blk << lkt) + (kt - 1 - uo);
                const unsigned ia = e * NT + u, ib = (7 - e) * NT + up;
                const W a = w[ia], b = w[ib];
                if (a > b) { w[ia] = b; w[ib] = a; }
            }
            __syncthreads();
            for (unsigned jt = kt >> 2; jt >= 32; jt >>= 1) {          // distances j = 8*jt >= 256
                const int lj = ilog2_pow2(jt);
                for (unsigned c = tid; c < 4u * NT; c += NT) {
                    const unsigned e = c / (NT / 2), cc = c % (NT / 2);
                    const unsigned u = ((cc >> lj) << (lj + 1)) + (cc & (jt - 1));
                    const unsigned ia = e * NT + u, ib = ia + jt;
                    const W a = w[ia], b = w[ib];
                    if (a > b) { w[ia] = b; w[ib] = a; }
                }
                __syncthreads();
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = w[e * NT + tid];
            for (unsigned jl = 16; jl >= 1; jl >>= 1) half_shfl(v, jl, lane);
            tail_in(v);
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
    uint16_t lo_bin, hi_bin;         // nearest non-empty bins below / above the median bin when sorted[nh-1] / sorted[nh+1]
                                     // lie outside it (0xffff: not needed); read by the lean relabel kernels
};

struct TopArgs {
    int64_t n;                       // points of this job (labels, cand are [Tg][n])
    int64_t ks, ps;                  // element stride between key rows / between the trees' slices of perm
    int ch;                          // points per CTA of the streaming kernels (8192 .. TOP_CH, histogram kernels up to HIST_CH_MAX; small tree groups use small chunks to fill the SMs)
    int vec;                         // key rows and labels are 32-byte / 8-byte aligned: 4-point vector accesses allowed
    int haslab;                      // labels are valid (level > 0, or a job with several roots); else every point sits in node 0
    int Tg, L, l, node0, nnodes, NTOP, NB, HSZ, MAXTD, smem_hist, gt0;   // gt0: global tree id of the group's first tree
    int all_internal, child0;        // every node of level l splits; BFS id of the first node of level l+1
    int scatter_fast;                // last top level: every point lands in a child of this level (CTA-aggregated scatter)
    uint32_t* track_any;             // [Tg] set by k_top_finish when a margin neighbour lies outside the median bin
    int64_t nn_all;                                                         // nodes per tree (stride of thr/mlo/mhi)
    const ull* keys;
    uint16_t* label;
    uint16_t* pbin;                  // [Tg][n] bin of every point's key at the current level (written by k_top_hist): lets
                                     // k_top_compact / k_top_relabel stream 2 bytes per point instead of the 8-byte key
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
    uint32_t* done;                  // [Tg] CTAs of k_top_hist that finished the tree (NULL: separate pick kernels)
    int NB2, nnodes2;                // k_top_relabel_hist: bins per node / nodes of level l + 1 (its first node is child0)
    uint32_t* done2;                 //   and its ticket counters (NULL: separate pick kernels), see `done`
    int fused_finish;                // k_top_finish_all instead of finish_warp -> finish -> ties
    uint32_t* wl_cnt;                // [2] work-list lengths of this level: big median bins (k_top_finish), straddling ties (k_top_ties)
    uint32_t* wl_big;                // [Tg * nnodes] entries t * nnodes + nl
    uint32_t* wl_tie;
    ull* pivots;
    uint32_t* fill;
    uint32_t* perm;
    double *thr, *mlo, *mhi;
};

#ifndef TOP_CH
#define TOP_CH 32768      /* points per CTA in the streaming top-phase kernels (<= 65535: 16-bit histogram counters) */
#endif
#ifndef TOP_NT
#define TOP_NT 512
#endif
#define HIST_CH_MAX 57344 /* points per CTA of the histogram kernels: multiples of 8192 below 65 536 (16-bit counters) */
#define HBINS 32768       /* shared-memory histogram counters (16-bit, two per word; a CTA streams at most HIST_CH_MAX <= 65535 points) */
#define HBINS_MAXNB 16384
#define FIN_CAP 4096      /* in-bin sort capacity */
#define SMEM_NODES 1024   /* compact/relabel keep per-node state in shared memory up to this many nodes */
#define SCAT_MAX 2048     /* children handled by the CTA-aggregated scatter of the last top level */

// (monotone: difference in fp64 -- keys far from zero keep their resolution --, scaling and conversion in fp32: NB <= 65536
//  bins against a 24-bit mantissa; three cheap instructions instead of an fp64 multiply and an fp64 -> int conversion)
__device__ __forceinline__ int key_bin(ull o, double lo, float scf, int NB) {
    // (the unsigned conversion saturates: negative and NaN -> 0, so one min() clamps both ends)
    return (int)min(__float2uint_rz(__double2float_rz(ord2f(o) - lo) * scf), (unsigned)(NB - 1));
}

// Called by one whole warp.  When the margin neighbours sorted[nh-1] / sorted[nh+1] lie outside the median bin they are
// the largest key of the nearest non-empty bin below / the smallest key of the nearest non-empty bin above (monotone
// binning): find those bins in the node's histogram so that the relabel pass only reads the keys of these bins.
__device__ __forceinline__ void warp_track_bins(const TopArgs& A, NodeSel& S, int t, int nl, bool need_pred, bool need_succ) {
    const int lane = threadIdx.x & 31;
    const uint32_t* hr = A.hist + (int64_t)t * A.HSZ + (int64_t)nl * A.NB;
    const int sb = S.sel_bin;
    int lo = 0xffff, hi = 0xffff;
    if (need_pred) {
        for (int b0 = sb - 1; b0 >= 0; b0 -= 32) {
            const int b = b0 - lane;
            const unsigned m = __ballot_sync(0xffffffffu, b >= 0 && hr[b] != 0);
            if (m) { lo = b0 - (__ffs(m) - 1); break; }
        }
    }
    if (need_succ) {
        for (int b0 = sb + 1; b0 < A.NB; b0 += 32) {
            const int b = b0 + lane;
            const unsigned m = __ballot_sync(0xffffffffu, b < A.NB && hr[b] != 0);
            if (m) { hi = b0 + (__ffs(m) - 1); break; }
        }
    }
    if (lane == 0) { S.lo_bin = (uint16_t)lo; S.hi_bin = (uint16_t)hi; }
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
    int g = A.haslab ? (int)lab[i] : 0;
    int nl = g - A.node0;
    if ((unsigned)nl >= (unsigned)A.nnodes) return -1;
    if (__ldg(A.child + g) < 0) return -1;
    return nl;
}

// ---- streaming access to the points of one tree: 4 points per thread per step (one 32-byte key load + one 8-byte
// label load), two steps in flight.  Falls back to scalar accesses when the rows are not 32-byte aligned (n % 4 != 0).
struct Pt4 { ull k[4]; uint16_t g[4]; };
__device__ __forceinline__ void load_pt4(const TopArgs& A, const ull* __restrict__ keys, const uint16_t* lab, int64_t i, Pt4& p) {
    asm("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(p.k[0]), "=l"(p.k[1]), "=l"(p.k[2]), "=l"(p.k[3]) : "l"(keys + i));
    if (A.haslab) {
        const uint2 q = *(const uint2*)(lab + i);
        p.g[0] = (uint16_t)(q.x & 0xffff); p.g[1] = (uint16_t)(q.x >> 16); p.g[2] = (uint16_t)(q.y & 0xffff); p.g[3] = (uint16_t)(q.y >> 16);
    } else {
        p.g[0] = p.g[1] = p.g[2] = p.g[3] = 0;
    }
}
// local node index of a point sitting in node g, or -1 when g is not an internal node of level A.l
__device__ __forceinline__ int node_of(const TopArgs& A, int g) {
    const int nl = g - A.node0;
    if ((unsigned)nl >= (unsigned)A.nnodes) return -1;
    if (!A.all_internal && __ldg(A.child + g) < 0) return -1;
    return nl;
}
template <typename F>
__device__ __forceinline__ void stream_points(const TopArgs& A, const ull* __restrict__ keys, const uint16_t* lab, int64_t i0, int64_t i1, F&& f) {
    if (A.vec) {
        const int64_t step = 4 * TOP_NT;
        int64_t i = i0 + 4 * (int64_t)threadIdx.x;
        for (; i + step < i1; i += 2 * step) {
            Pt4 a, b;
            load_pt4(A, keys, lab, i, a);
            load_pt4(A, keys, lab, i + step, b);
#pragma unroll
            for (int u = 0; u < 4; ++u) f(i + u, a.k[u], (int)a.g[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) f(i + step + u, b.k[u], (int)b.g[u]);
        }
        if (i < i1) {
            Pt4 a;
            load_pt4(A, keys, lab, i, a);
#pragma unroll
            for (int u = 0; u < 4; ++u) f(i + u, a.k[u], (int)a.g[u]);
        }
    } else {
        for (int64_t i = i0 + threadIdx.x; i < i1; i += TOP_NT) f(i, keys[i], A.haslab ? (int)lab[i] : 0);
    }
}

// streaming access to (bin, label) of the points of one tree: 4 points per thread per step (two 8-byte loads), two steps in
// flight; f(i, bin, node)
template <typename F>
__device__ __forceinline__ void stream_bins(const TopArgs& A, const uint16_t* __restrict__ pb, const uint16_t* lab, int64_t i0, int64_t i1, F&& f) {
    if ((A.n & 3) == 0) {
        const int64_t step = 4 * TOP_NT;
        auto ld4 = [&](int64_t i, unsigned (&b)[4], int (&g)[4]) {
            const uint2 q = *(const uint2*)(pb + i);
            b[0] = q.x & 0xffff; b[1] = q.x >> 16; b[2] = q.y & 0xffff; b[3] = q.y >> 16;
            if (A.haslab) { const uint2 r = *(const uint2*)(lab + i); g[0] = r.x & 0xffff; g[1] = r.x >> 16; g[2] = r.y & 0xffff; g[3] = r.y >> 16; }
            else { g[0] = g[1] = g[2] = g[3] = 0; }
        };
        int64_t i = i0 + 4 * (int64_t)threadIdx.x;
        for (; i + step < i1; i += 2 * step) {
            unsigned ba[4], bb[4]; int ga[4], gb[4];
            ld4(i, ba, ga); ld4(i + step, bb, gb);
#pragma unroll
            for (int u = 0; u < 4; ++u) f(i + u, ba[u], ga[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) f(i + step + u, bb[u], gb[u]);
        }
        if (i < i1) {
            unsigned ba[4]; int ga[4];
            ld4(i, ba, ga);
#pragma unroll
            for (int u = 0; u < 4; ++u) f(i + u, ba[u], ga[u]);
        }
    } else {
        for (int64_t i = i0 + threadIdx.x; i < i1; i += TOP_NT) f(i, (unsigned)pb[i], A.haslab ? (int)lab[i] : 0);
    }
}

template <int NT> __device__ void pick_node_block(const TopArgs& A, int t, int nl, uint32_t* wsum, uint32_t* own);
__device__ __forceinline__ void pick_node_warp(const TopArgs& A, int t, int nl);

__global__ void __launch_bounds__(TOP_NT) k_top_hist(TopArgs A) {
    extern __shared__ uint32_t sh[];
    const int t = blockIdx.y, tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * A.ch, i1 = min(A.n, i0 + A.ch);
    const ull* keys = A.keys + ((int64_t)t * A.L + A.l) * A.ks;
    const uint16_t* lab = A.label + (int64_t)t * A.n;
    const double lo = A.binlo[t * A.L + A.l];
    const float sc = __double2float_rz(A.binscale[t * A.L + A.l]);
    const int NB = A.NB, tot = A.nnodes * NB;
    uint32_t* gh = A.hist + (int64_t)t * A.HSZ;
    if (A.smem_hist) {
        for (int j = tid; j < (tot + 1) / 2; j += TOP_NT) sh[j] = 0;
        __syncthreads();
    }
    uint16_t* pb = A.pbin + (int64_t)t * A.n;
    auto one = [&](ull kv, int g) -> unsigned {        // counts the point, returns its bin (0 when it sits in no splitting node)
        const int nl = node_of(A, g);
        if (nl < 0) return 0u;
        const int b = key_bin(kv, lo, sc, NB);
        const int j = nl * NB + b;
        if (A.smem_hist) atomicAdd(&sh[j >> 1], 1u << ((j & 1) << 4));
        else atomicAdd(&gh[j], 1u);
        return (unsigned)b;
    };
    if (A.vec) {
        const int64_t step = 4 * TOP_NT;
        int64_t i = i0 + 4 * (int64_t)tid;
        for (; i + step < i1; i += 2 * step) {
            Pt4 a, b;
            load_pt4(A, keys, lab, i, a);
            load_pt4(A, keys, lab, i + step, b);
            unsigned ba[4], bb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) ba[u] = one(a.k[u], (int)a.g[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) bb[u] = one(b.k[u], (int)b.g[u]);
            *(uint2*)(pb + i) = make_uint2(ba[0] | (ba[1] << 16), ba[2] | (ba[3] << 16));
            *(uint2*)(pb + i + step) = make_uint2(bb[0] | (bb[1] << 16), bb[2] | (bb[3] << 16));
        }
        if (i < i1) {
            Pt4 a;
            load_pt4(A, keys, lab, i, a);
            unsigned ba[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) ba[u] = one(a.k[u], (int)a.g[u]);
            *(uint2*)(pb + i) = make_uint2(ba[0] | (ba[1] << 16), ba[2] | (ba[3] << 16));
        }
    } else {
        for (int64_t i = i0 + tid; i < i1; i += TOP_NT) pb[i] = (uint16_t)one(keys[i], A.haslab ? (int)lab[i] : 0);
    }
    if (A.smem_hist) {
        __syncthreads();
        // one 64-bit RED per pair of adjacent 32-bit counters (the low one never carries: counts stay below 2^32)
        for (int w2 = tid; w2 < (tot + 1) / 2; w2 += TOP_NT) {
            const uint32_t v = sh[w2];
            if (v) atomicAdd((ull*)gh + w2, (ull)(v & 0xffffu) | ((ull)(v >> 16) << 32));
        }
    }
    if (!A.done) return;
    // ---- fused pick: the CTA that completes the tree's histogram (ticket counter) locates every node's median bin right
    // away -- one launch and one kernel-boundary less per level
    __shared__ uint32_t s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&A.done[t], 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int nl = tid >> 5; nl < A.nnodes; nl += TOP_NT / 32) pick_node_warp(A, t, nl);      // 16 nodes at a time
}

// Median bin of node (t, nl): the bin holding rank nh = size/2.  All NT threads of the CTA call it; thread j sums a contiguous
// range of bins, a shuffle scan over the partial sums names the range that covers the rank, warp 0 then scans that range.
// (__ldcg: the counters were written by other CTAs' atomics -- read them where the atomics landed, in L2.)
template <int NT>
__device__ void pick_node_block(const TopArgs& A, int t, int nl, uint32_t* wsum /*[NT/32]*/, uint32_t* own /*[2]*/) {
    const int g = A.node0 + nl, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (A.child[g] < 0) return;                    // block-uniform
    const uint32_t k = A.nsize[g] >> 1;
    const uint32_t* hr = A.hist + (int64_t)t * A.HSZ + (int64_t)nl * A.NB;
    const int per = (A.NB + NT - 1) / NT;
    const int b0 = min(A.NB, tid * per), b1 = min(A.NB, b0 + per);
    uint32_t s = 0;
#pragma unroll 8
    for (int b = b0; b < b1; ++b) s += __ldcg(hr + b);
    uint32_t incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += y; }
    __syncthreads();                               // scratch reuse across calls
    if (lane == 31) wsum[wid] = incl;
    if (tid == 0) { own[0] = 0; own[1] = 0; }
    __syncthreads();
    uint32_t wofs = 0;
    for (int j = 0; j < wid; ++j) wofs += wsum[j];
    const uint32_t excl = wofs + incl - s;
    if (k >= excl && k < excl + s) { own[0] = (uint32_t)b0; own[1] = excl; }     // exactly one thread
    __syncthreads();
    if (wid == 0) {
        const int ob0 = (int)own[0], ob1 = min(A.NB, ob0 + per);
        uint32_t c = own[1];
        for (int base = ob0; base < ob1; base += 32) {
            const int b = base + lane;
            const uint32_t hb = b < ob1 ? __ldcg(hr + b) : 0u;
            uint32_t in2 = hb;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, in2, off); if (lane >= off) in2 += y; }
            const uint32_t ex2 = c + in2 - hb;
            const bool hit = hb > 0 && k >= ex2 && k < ex2 + hb;
            if (hit) {
                NodeSel& S = A.sel[(int64_t)t * A.NTOP + g];
                S.sel_bin = b; S.below = ex2; S.cand_cnt = hb; S.cand_fill = 0;
                S.cand_off = atomicAdd(&A.cand_total[t], hb);
            }
            if (__any_sync(0xffffffffu, hit)) break;
            c += __shfl_sync(0xffffffffu, in2, 31);
        }
    }
}
// same, one WARP per (node, tree): the deep top levels have thousands of nodes.  The warp walks the node's bins in chunks
// of 32 (one coalesced load + one shuffle scan per chunk, the next chunk's load already in flight) with a running prefix.
__device__ __forceinline__ void pick_node_warp(const TopArgs& A, int t, int nl) {
    const int lane = threadIdx.x & 31, g = A.node0 + nl;
    if (A.child[g] < 0) return;
    const uint32_t k = A.nsize[g] >> 1;
    const uint32_t* hr = A.hist + (int64_t)t * A.HSZ + (int64_t)nl * A.NB;
    uint32_t c = 0;
    uint32_t nxt = lane < A.NB ? __ldcg(hr + lane) : 0u;
    for (int base = 0; base < A.NB; base += 32) {
        const uint32_t hb = nxt;
        const int bn = base + 32 + lane;
        nxt = bn < A.NB ? __ldcg(hr + bn) : 0u;
        uint32_t in2 = hb;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, in2, off); if (lane >= off) in2 += y; }
        const uint32_t ex2 = c + in2 - hb;
        const bool hit = hb > 0 && k >= ex2 && k < ex2 + hb;
        if (hit) {
            NodeSel& S = A.sel[(int64_t)t * A.NTOP + g];
            S.sel_bin = base + lane; S.below = ex2; S.cand_cnt = hb; S.cand_fill = 0;
            S.cand_off = atomicAdd(&A.cand_total[t], hb);
        }
        if (__any_sync(0xffffffffu, hit)) break;
        c += __shfl_sync(0xffffffffu, in2, 31);
    }
}
__global__ void __launch_bounds__(256) k_top_pick(TopArgs A) {
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t own[2];
    pick_node_block<256>(A, blockIdx.y, blockIdx.x, wsum, own);
}
__global__ void __launch_bounds__(256) k_top_pick_warp(TopArgs A) {
    const int64_t item = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (item >= (int64_t)A.nnodes * A.Tg) return;
    pick_node_warp(A, (int)(item / A.nnodes), (int)(item % A.nnodes));
}

__global__ void __launch_bounds__(TOP_NT) k_top_compact(TopArgs A) {
    __shared__ int s_bin[SMEM_NODES];
    const int t = blockIdx.y, tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * A.ch, i1 = min(A.n, i0 + A.ch);
    const ull* keys = A.keys + ((int64_t)t * A.L + A.l) * A.ks;
    const uint16_t* lab = A.label + (int64_t)t * A.n;
    const double lo = A.binlo[t * A.L + A.l], sc = A.binscale[t * A.L + A.l];
    NodeSel* sel = A.sel + (int64_t)t * A.NTOP + A.node0;
    const bool cached = A.nnodes <= SMEM_NODES;
    if (cached) {
        for (int j = tid; j < A.nnodes; j += TOP_NT) s_bin[j] = sel[j].sel_bin;
        __syncthreads();
    }
    ull* cand = A.cand + (int64_t)t * A.n;
    const uint16_t* pb = A.pbin + (int64_t)t * A.n;
    (void)lo; (void)sc;
    stream_bins(A, pb, lab, i0, i1, [&](int64_t i, unsigned b, int g) {
        const int nl = node_of(A, g);
        if (nl < 0) return;
        const int sb = cached ? s_bin[nl] : sel[nl].sel_bin;
        if ((int)b == sb) {                      // about one point in NB: only these keys are read
            const uint32_t pos = atomicAdd(&sel[nl].cand_fill, 1u);
            cand[sel[nl].cand_off + pos] = keys[i];
        }
    });
}

// one WARP per (node, tree) whose median bin holds <= 256 keys: register-blocked bitonic sort, no block barriers.
// Returns bit 0: the bin is larger (CTA-wide path needed), bit 1: a tie straddles the split (lexicographic select needed).
#define FW_MAX 256
#define FW_WARPS 8
__device__ __forceinline__ unsigned top_finish_small(const TopArgs& A, uint32_t item, ull* bufw /*[FW_MAX], this warp's*/) {
    const int lane = threadIdx.x & 31;
    const int t = (int)(item / A.nnodes), nl = (int)(item % A.nnodes), g = A.node0 + nl;
    if (A.child[g] < 0) return 0u;
    NodeSel& S = A.sel[(int64_t)t * A.NTOP + g];
    const uint32_t c = S.cand_cnt, nh = A.nsize[g] >> 1, r = nh - S.below;
    if (c > FW_MAX) return 1u;
    const ull* seg = A.cand + (int64_t)t * A.n + S.cand_off;
    ull v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { const uint32_t i = lane * 8 + e; v[e] = i < c ? seg[i] : 0xffffffffffffffffull; }
    sort_regs<32, ull>(v, (ull*)nullptr, max(8u, next_pow2_u32(c)));      // only the block that holds the c keys needs sorting
#pragma unroll
    for (int e = 0; e < 8; ++e) bufw[lane * 8 + e] = v[e];
    __syncwarp();
    const ull thr = bufw[r];
    uint32_t lt = 0, eq = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) { const uint32_t i = lane * 8 + e; if (i < c) { lt += v[e] < thr; eq += v[e] == thr; } }
    for (int off = 16; off > 0; off >>= 1) { lt += __shfl_xor_sync(0xffffffffu, lt, off); eq += __shfl_xor_sync(0xffffffffu, eq, off); }
    const uint32_t lower = lt, upper = lt + eq;
    const ull pred = lower > 0 ? bufw[lower - 1] : ORD_NONE_LO;
    const ull succ = upper < c ? bufw[upper] : ORD_NONE_HI;
    const uint32_t cless = S.below + lower;
    const bool need_pred = (cless == nh) && pred == ORD_NONE_LO;
    const bool need_succ = (cless + eq == nh + 1) && succ == ORD_NONE_HI;
    if (lane == 0) {
        S.thr = thr; S.pred = pred; S.succ = succ;
        S.cless = cless; S.ceq = eq;
        S.tie_r = nh - cless;
        S.tie_depth = 0;
        if (need_pred || need_succ) atomicOr(&A.track_any[t], 1u);
    }
    warp_track_bins(A, S, t, nl, need_pred, need_succ);
    __syncwarp();
    return nh > cless ? 2u : 0u;
}
__global__ void __launch_bounds__(FW_WARPS * 32) k_top_finish_warp(TopArgs A) {
    __shared__ ull buf[FW_WARPS][FW_MAX];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int64_t item = (int64_t)blockIdx.x * FW_WARPS + wi;
    if (item >= (int64_t)A.nnodes * A.Tg) return;
    const unsigned f = top_finish_small(A, (uint32_t)item, buf[wi]);
    if (lane == 0) {
        if (f & 1u) A.wl_big[atomicAdd(&A.wl_cnt[0], 1u)] = (uint32_t)item;      // rare: handed to k_top_finish through the work list
        if (f & 2u) A.wl_tie[atomicAdd(&A.wl_cnt[1], 1u)] = (uint32_t)item;
    }
}

// exact order statistic inside the median bin for a bin with more than FW_MAX keys; all 512 threads of the CTA.
// Returns (to every thread) whether a tie straddles the split.
__device__ bool top_finish_big(const TopArgs& A, uint32_t item, ull* buf /*[FIN_CAP]*/, uint32_t* sh /*[264]*/, ull* sh64 /*[3]*/) {
    const int tid = threadIdx.x;
    const int t = (int)(item / A.nnodes), nl = (int)(item % A.nnodes), g = A.node0 + nl;
    __syncthreads();                              // shared buffers are reused from the previous work item
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
    // sorted[nh-1] / sorted[nh+1] are taken from the bin's sorted keys; only when they fall outside the bin must
    // k_top_relabel track the nearest keys below / above the threshold (they sit in the nearest non-empty bins)
    const uint32_t cless = S.below + lower;
    const bool need_pred = (cless == nh) && pred == ORD_NONE_LO;
    const bool need_succ = (cless + ceq == nh + 1) && succ == ORD_NONE_HI;
    if (tid == 0) {
        S.thr = thr; S.pred = pred; S.succ = succ;
        S.cless = cless; S.ceq = ceq;
        S.tie_r = nh - cless;        // tied points that must go left; > 0 => the split cuts through a tie
        S.tie_depth = 0;
        if (need_pred || need_succ) atomicOr(&A.track_any[t], 1u);
    }
    if (tid < 32) warp_track_bins(A, S, t, nl, need_pred, need_succ);
    return nh > cless;
}
// the CTAs walk the work list k_top_finish_warp filled (a grid of one CTA per node would spend the deep levels launching
// CTAs that return at once)
__global__ void __launch_bounds__(512) k_top_finish(TopArgs A) {
    __shared__ ull buf[FIN_CAP];
    __shared__ uint32_t sh[264];
    __shared__ ull sh64[3];
    const uint32_t nwork = A.wl_cnt[0];
    for (uint32_t wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
        const uint32_t item = A.wl_big[wi];
        if (top_finish_big(A, item, buf, sh, sh64) && threadIdx.x == 0) A.wl_tie[atomicAdd(&A.wl_cnt[1], 1u)] = item;
    }
}

// Only does work when a tie straddles the split.  Finds the composite pivot (key_{l-1}, key_{l-2}, ..., key_0, row id) such
// that exactly tie_r tied points are lexicographically below it: this is the order the reference's stable merge sort
// leaves tied points in (Internal.hs:504-512).  All 512 threads of the CTA.
__device__ void top_ties_item(const TopArgs& A, uint32_t item, uint32_t* sh /*[264]*/, ull* sh64 /*[1]*/, uint32_t* cnt) {
    const int tid = threadIdx.x;
    const int t = (int)(item / A.nnodes), nl = (int)(item % A.nnodes), g = A.node0 + nl;
    __syncthreads();
    NodeSel& S = A.sel[(int64_t)t * A.NTOP + g];
    const ull* keys_t = A.keys + (int64_t)t * A.L * A.ks;
    const ull* keys_l = keys_t + (int64_t)A.l * A.ks;
    const uint16_t* lab = A.label + (int64_t)t * A.n;
    const ull thr = S.thr;
    uint32_t* la = (uint32_t*)(A.cand + (int64_t)t * A.n) + 2 * (int64_t)A.nstart[g];
    uint32_t* lb = la + A.nsize[g];
    if (tid == 0) *cnt = 0;
    __syncthreads();
    for (int64_t i = tid; i < A.n; i += 512) {
        int gi = A.haslab ? (int)lab[i] : 0;
        if (gi == g && keys_l[i] == thr) { uint32_t p = atomicAdd(cnt, 1u); la[p] = (uint32_t)i; }
    }
    __syncthreads();
    uint32_t c = *cnt, rr = S.tie_r;
    int depth = 0;
    ull* piv = A.pivots + ((int64_t)t * A.NTOP + g) * A.MAXTD;
    while (true) {
        const int lvl = A.l - 1 - depth;
        const ull* kk = lvl >= 0 ? keys_t + (int64_t)lvl * A.ks : nullptr;
        uint32_t cl, ce;
        ull pv = cta_radix_select<512>(c, rr, [&](uint32_t i) { uint32_t id = la[i]; return kk ? kk[id] : (ull)id; }, sh, sh64, cl, ce);
        if (tid == 0) piv[depth] = pv;
        ++depth;
        const uint32_t r2 = rr - cl;
        if (r2 == 0 || lvl < 0) break;
        if (tid == 0) *cnt = 0;
        __syncthreads();
        for (uint32_t i = tid; i < c; i += 512) {
            uint32_t id = la[i];
            if (kk[id] == pv) { uint32_t p = atomicAdd(cnt, 1u); lb[p] = id; }
        }
        __syncthreads();
        c = *cnt; rr = r2;
        uint32_t* tmp = la; la = lb; lb = tmp;
        __syncthreads();
    }
    if (tid == 0) S.tie_depth = depth;
}
__global__ void __launch_bounds__(512) k_top_ties(TopArgs A) {
    __shared__ uint32_t sh[264];
    __shared__ ull sh64[1];
    __shared__ uint32_t cnt;
    const uint32_t nwork = A.wl_cnt[1];
    for (uint32_t wi = blockIdx.x; wi < nwork; wi += gridDim.x) top_ties_item(A, A.wl_tie[wi], sh, sh64, &cnt);
}

// finish_warp -> finish -> ties of one level in ONE launch: the 16 warps of a CTA each settle one node (median bin <= 256
// keys: the usual case); the nodes of the CTA that need the CTA-wide sort / the tie select (rare) are then taken by the
// whole CTA, one after the other.  Replaces three launches -- two of them fixed-size grids that mostly found empty work
// lists -- per level.
#define FA_WARPS 16
__global__ void __launch_bounds__(FA_WARPS * 32) k_top_finish_all(TopArgs A) {
    __shared__ ull buf[FIN_CAP];                          // phase 1: [FA_WARPS][FW_MAX]; phase 2: the CTA-wide sort buffer
    __shared__ uint32_t sh[264];
    __shared__ ull sh64[3];
    __shared__ uint32_t cnt;
    __shared__ uint32_t s_flag[FA_WARPS];
    static_assert(FA_WARPS * FW_MAX <= FIN_CAP, "phase-1 buffers alias the CTA-wide sort buffer");
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int64_t total = (int64_t)A.nnodes * A.Tg, item0 = (int64_t)blockIdx.x * FA_WARPS;
    unsigned f = 0;
    if (item0 + wi < total) f = top_finish_small(A, (uint32_t)(item0 + wi), buf + (size_t)wi * FW_MAX);
    if (lane == 0) s_flag[wi] = f;
    __syncthreads();
    for (int w2 = 0; w2 < FA_WARPS; ++w2) {
        if (s_flag[w2] & 1u) {                            // block-uniform
            const bool tie = top_finish_big(A, (uint32_t)(item0 + w2), buf, sh, sh64);
            __syncthreads();
            if (threadIdx.x == 0) s_flag[w2] = tie ? 2u : 0u;
            __syncthreads();
        }
    }
    for (int w2 = 0; w2 < FA_WARPS; ++w2)
        if (s_flag[w2] & 2u) top_ties_item(A, (uint32_t)(item0 + w2), sh, sh64, &cnt);
}

// relabel every point of an internal level-l node to its child; track the keys adjacent to the threshold
// (margins); at the last top level also scatter the points into the per-node segments of perm.
#ifndef RELABEL_MINB
#define RELABEL_MINB 2
#endif
__global__ void __launch_bounds__(TOP_NT, RELABEL_MINB) k_top_relabel(TopArgs A, int last) {
    __shared__ ull s_thr[SMEM_NODES], s_pred[SMEM_NODES], s_succ[SMEM_NODES];
    __shared__ uint32_t s_tie[SMEM_NODES];
    __shared__ uint16_t s_sbin[SMEM_NODES];
    __shared__ uint32_t s_cnt[SCAT_MAX], s_base[SCAT_MAX];
    const int t = blockIdx.y, tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * A.ch, i1 = min(A.n, i0 + A.ch);
    const ull* keys_t = A.keys + (int64_t)t * A.L * A.ks;
    const ull* keys = keys_t + (int64_t)A.l * A.ks;
    uint16_t* lab = A.label + (int64_t)t * A.n;
    const bool sf = last && A.scatter_fast;
    if (sf) for (int j = tid; j < 2 * A.nnodes; j += TOP_NT) s_cnt[j] = 0;
    NodeSel* sel = A.sel + (int64_t)t * A.NTOP + A.node0;
    const bool cached = A.nnodes <= SMEM_NODES;
    if (cached) {
        for (int j = tid; j < A.nnodes; j += TOP_NT) {
            s_thr[j] = sel[j].thr; s_pred[j] = sel[j].pred; s_succ[j] = sel[j].succ; s_tie[j] = sel[j].tie_r;
            s_sbin[j] = (uint16_t)sel[j].sel_bin;
        }
    }
    __syncthreads();
    uint32_t* fill = A.fill + (int64_t)t * A.NTOP;
    uint32_t* perm = A.perm + (int64_t)t * A.ps;
    const bool fast = cached && A.all_internal && A.track_any[t] == 0;     // block-uniform
    // returns the point's node after this level's split (unchanged when it does not sit in a splitting node)
    auto relabel_one = [&](int64_t i, ull kv, int g) -> int {
        const int nl = g - A.node0;
        if (fast) {
            if ((unsigned)nl < (unsigned)A.nnodes) {
                const ull thr = s_thr[nl];
                bool left = kv < thr;
                if (kv == thr && s_tie[nl] > 0) {   // composite compare against the tie pivot (rare)
                    const int td = sel[nl].tie_depth;
                    const ull* piv = A.pivots + ((int64_t)t * A.NTOP + g) * A.MAXTD;
                    for (int j = 0; j < td; ++j) {
                        const int lvl = A.l - 1 - j;
                        const ull kq = lvl >= 0 ? keys_t[(int64_t)lvl * A.ks + i] : (ull)i;
                        const ull pv = piv[j];
                        if (kq != pv) { left = kq < pv; break; }
                    }
                }
                g = A.child0 + 2 * nl + (left ? 0 : 1);
            }
        } else {
            const int ch = ((unsigned)nl < (unsigned)A.nnodes) ? __ldg(A.child + g) : -1;
            if (ch >= 0) {
                const ull thr = cached ? s_thr[nl] : sel[nl].thr;
                bool left = kv < thr;
                if (kv == thr) {
                    const uint32_t tr = cached ? s_tie[nl] : sel[nl].tie_r;
                    if (tr > 0) {   // composite compare against the tie pivot
                        const int td = sel[nl].tie_depth;
                        const ull* piv = A.pivots + ((int64_t)t * A.NTOP + g) * A.MAXTD;
                        for (int j = 0; j < td; ++j) {
                            const int lvl = A.l - 1 - j;
                            const ull kq = lvl >= 0 ? keys_t[(int64_t)lvl * A.ks + i] : (ull)i;
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
            }
        }
        if (last) {
            if (sf) atomicAdd(&s_cnt[g - A.child0], 1u);
            else { const uint32_t pos = atomicAdd(&fill[g], 1u); perm[A.nstart[g] + pos] = (uint32_t)i; }
        }
        return g;
    };
    // fast path: the side of the split follows from the point's BIN unless it sits in the median bin (monotone binning:
    // bin < median bin => key < thr, bin > median bin => key > thr), so only ~1 key in NB is read
    const uint16_t* pb = A.pbin + (int64_t)t * A.n;
    auto relabel_bin = [&](int64_t i, unsigned b, int g) -> int {
        const int nl = g - A.node0;
        if ((unsigned)nl < (unsigned)A.nnodes) {
            const unsigned sb = s_sbin[nl];
            bool left = b < sb;
            if (b == sb) {
                const ull kv = keys[i], thr = s_thr[nl];
                left = kv < thr;
                if (kv == thr && s_tie[nl] > 0) {   // composite compare against the tie pivot (rare)
                    const int td = sel[nl].tie_depth;
                    const ull* piv = A.pivots + ((int64_t)t * A.NTOP + g) * A.MAXTD;
                    for (int j = 0; j < td; ++j) {
                        const int lvl = A.l - 1 - j;
                        const ull kq = lvl >= 0 ? keys_t[(int64_t)lvl * A.ks + i] : (ull)i;
                        const ull pv = piv[j];
                        if (kq != pv) { left = kq < pv; break; }
                    }
                }
            }
            g = A.child0 + 2 * nl + (left ? 0 : 1);
        }
        if (last) {
            if (sf) atomicAdd(&s_cnt[g - A.child0], 1u);
            else { const uint32_t pos = atomicAdd(&fill[g], 1u); perm[A.nstart[g] + pos] = (uint32_t)i; }
        }
        return g;
    };
    if (fast && (A.n & 3) == 0) {
        const int64_t step = 4 * TOP_NT;
        auto ld4 = [&](int64_t i, unsigned (&b)[4], int (&g)[4]) {
            const uint2 q = *(const uint2*)(pb + i);
            b[0] = q.x & 0xffff; b[1] = q.x >> 16; b[2] = q.y & 0xffff; b[3] = q.y >> 16;
            if (A.haslab) { const uint2 r = *(const uint2*)(lab + i); g[0] = r.x & 0xffff; g[1] = r.x >> 16; g[2] = r.y & 0xffff; g[3] = r.y >> 16; }
            else { g[0] = g[1] = g[2] = g[3] = 0; }
        };
        int64_t i = i0 + 4 * (int64_t)tid;
        for (; i + step < i1; i += 2 * step) {
            unsigned ba[4], bb[4]; int ia[4], ib[4];
            ld4(i, ba, ia); ld4(i + step, bb, ib);
            uint32_t ga[4], gb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) ga[u] = (uint32_t)relabel_bin(i + u, ba[u], ia[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) gb[u] = (uint32_t)relabel_bin(i + step + u, bb[u], ib[u]);
            *(uint2*)(lab + i) = make_uint2(ga[0] | (ga[1] << 16), ga[2] | (ga[3] << 16));
            *(uint2*)(lab + i + step) = make_uint2(gb[0] | (gb[1] << 16), gb[2] | (gb[3] << 16));
        }
        if (i < i1) {
            unsigned ba[4]; int ia[4];
            ld4(i, ba, ia);
            uint32_t ga[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) ga[u] = (uint32_t)relabel_bin(i + u, ba[u], ia[u]);
            *(uint2*)(lab + i) = make_uint2(ga[0] | (ga[1] << 16), ga[2] | (ga[3] << 16));
        }
    } else if (fast) {
        for (int64_t i = i0 + tid; i < i1; i += TOP_NT) lab[i] = (uint16_t)relabel_bin(i, (unsigned)pb[i], A.haslab ? (int)lab[i] : 0);
    } else if (A.vec) {
        const int64_t step = 4 * TOP_NT;
        int64_t i = i0 + 4 * (int64_t)tid;
        for (; i + step < i1; i += 2 * step) {
            Pt4 a, b;
            load_pt4(A, keys, lab, i, a);
            load_pt4(A, keys, lab, i + step, b);
            uint32_t ga[4], gb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) ga[u] = (uint32_t)relabel_one(i + u, a.k[u], (int)a.g[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) gb[u] = (uint32_t)relabel_one(i + step + u, b.k[u], (int)b.g[u]);
            *(uint2*)(lab + i) = make_uint2(ga[0] | (ga[1] << 16), ga[2] | (ga[3] << 16));
            *(uint2*)(lab + i + step) = make_uint2(gb[0] | (gb[1] << 16), gb[2] | (gb[3] << 16));
        }
        if (i < i1) {
            Pt4 a;
            load_pt4(A, keys, lab, i, a);
            uint32_t ga[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) ga[u] = (uint32_t)relabel_one(i + u, a.k[u], (int)a.g[u]);
            *(uint2*)(lab + i) = make_uint2(ga[0] | (ga[1] << 16), ga[2] | (ga[3] << 16));
        }
    } else {
        for (int64_t i = i0 + tid; i < i1; i += TOP_NT) lab[i] = (uint16_t)relabel_one(i, keys[i], A.haslab ? (int)lab[i] : 0);
    }
    if (sf) {
        // reserve a range per child for this CTA (one global atomic per non-empty child), then place the points
        __syncthreads();
        for (int j = tid; j < 2 * A.nnodes; j += TOP_NT) {
            const uint32_t c = s_cnt[j];
            s_base[j] = c ? atomicAdd(&fill[A.child0 + j], c) : 0u;
            s_cnt[j] = 0;
        }
        __syncthreads();
        auto place = [&](int64_t i, int g) {
            const int j = g - A.child0;
            const uint32_t pos = s_base[j] + atomicAdd(&s_cnt[j], 1u);
            perm[A.nstart[g] + pos] = (uint32_t)i;
        };
        if ((A.n & 3) == 0 && (fast || A.vec)) {
            for (int64_t i = i0 + 4 * (int64_t)tid; i < i1; i += 4 * TOP_NT) {
                const uint2 q = *(const uint2*)(lab + i);          // labels written by this same thread above
                place(i, (int)(q.x & 0xffff)); place(i + 1, (int)(q.x >> 16)); place(i + 2, (int)(q.y & 0xffff)); place(i + 3, (int)(q.y >> 16));
            }
        } else {
            for (int64_t i = i0 + tid; i < i1; i += TOP_NT) place(i, (int)lab[i]);
        }
    }
    if (cached && !fast) {
        __syncthreads();
        for (int j = tid; j < A.nnodes; j += TOP_NT) {
            if (s_pred[j] > sel[j].pred) atomicMax(&sel[j].pred, s_pred[j]);
            if (s_succ[j] < sel[j].succ) atomicMin(&sel[j].succ, s_succ[j]);
        }
    }
}

// ---- lean top-phase kernels --------------------------------------------------------------------------------------
// Used when the level's nodes all split, fit the shared-memory tables (<= SMEM_NODES) and n % 8 == 0 (16-byte rows of
// 2-byte bins / labels).  They stream 8 points per load (one 16-byte load of bins + one of labels, LEAN_G such pairs in
// flight per thread) and touch an 8-byte key only for a point in the median bin or in a margin-tracking bin
// (NodeSel::lo_bin / hi_bin), so no level has to fall back to streaming the keys.
// bins: x = median bin | lo_bin << 16, y = hi_bin -- ONE 8-byte shared-memory read per point (three 2-byte table reads were 17 %
// of the fused kernel's issue slots, ncu source page)
struct LeanTabs { ull thr[SMEM_NODES]; uint2 bins[SMEM_NODES]; };

__device__ __forceinline__ void lean_load_tabs(const TopArgs& A, const NodeSel* sel, LeanTabs& T, int nthreads) {
    for (int j = threadIdx.x; j < A.nnodes; j += nthreads) {
        T.thr[j] = sel[j].thr;
        T.bins[j] = make_uint2(((unsigned)sel[j].sel_bin & 0xffffu) | ((unsigned)sel[j].lo_bin << 16), (unsigned)sel[j].hi_bin);
    }
}

// Sides of the split for 8 consecutive points i..i+7 (bins b, local node indices nl; nl outside [0, nnodes) = the point
// does not sit in a node of this level).  Bit u of the result: point u goes to the RIGHT child.  The bin decides unless
// the point sits in the median bin; the keys of median-bin and margin-tracking-bin points are fetched together (all
// loads issued before the first use: one memory latency per 8 points instead of one per hit).
__device__ __forceinline__ unsigned lean_sides8(const TopArgs& A, NodeSel* sel, const LeanTabs& T, const ull* __restrict__ keys,
                                                const ull* __restrict__ keys_t, int t, int64_t i, const unsigned (&b)[8], const int (&nl)[8]) {
    unsigned right = 0, kind = 0;               // kind: 2 bits per point (1 median bin, 2 tracked bin below, 3 tracked bin above)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if ((unsigned)nl[u] < (unsigned)A.nnodes) {
            const uint2 tb = T.bins[nl[u]];
            const unsigned sb = tb.x & 0xffffu;
            if (b[u] >= sb) right |= 1u << u;
            const unsigned k = b[u] == sb ? 1u : (b[u] == (tb.x >> 16) ? 2u : (b[u] == tb.y ? 3u : 0u));
            kind |= k << (2 * u);
        }
    }
    if (kind == 0) return right;
    ull kv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) kv[u] = ((kind >> (2 * u)) & 3u) ? __ldg(keys + i + u) : 0ull;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const unsigned k = (kind >> (2 * u)) & 3u;
        if (k == 1u) {
            const ull thr = T.thr[nl[u]];
            bool left = kv[u] < thr;
            if (kv[u] == thr && sel[nl[u]].tie_r > 0) {   // composite compare against the tie pivot (rare)
                const int td = sel[nl[u]].tie_depth;
                const ull* piv = A.pivots + ((int64_t)t * A.NTOP + A.node0 + nl[u]) * A.MAXTD;
                for (int j = 0; j < td; ++j) {
                    const int lvl = A.l - 1 - j;
                    const ull kq = lvl >= 0 ? keys_t[(int64_t)lvl * A.ks + i + u] : (ull)(i + u);
                    const ull pv = piv[j];
                    if (kq != pv) { left = kq < pv; break; }
                }
            }
            if (left) right &= ~(1u << u);
        } else if (k == 2u) {                   // every key of a lower bin is < thr; the largest one is sorted[nh-1]
            if (kv[u] > *(volatile ull*)&sel[nl[u]].pred) atomicMax(&sel[nl[u]].pred, kv[u]);
        } else if (k == 3u) {
            if (kv[u] < *(volatile ull*)&sel[nl[u]].succ) atomicMin(&sel[nl[u]].succ, kv[u]);
        }
    }
    return right;
}

__device__ __forceinline__ void unpack8(const uint4 q, unsigned (&v)[8]) {
    v[0] = q.x & 0xffff; v[1] = q.x >> 16; v[2] = q.y & 0xffff; v[3] = q.y >> 16;
    v[4] = q.z & 0xffff; v[5] = q.z >> 16; v[6] = q.w & 0xffff; v[7] = q.w >> 16;
}
__device__ __forceinline__ uint4 pack8(const unsigned (&g)[8]) {
    return make_uint4(g[0] | (g[1] << 16), g[2] | (g[3] << 16), g[4] | (g[5] << 16), g[6] | (g[7] << 16));
}

// relabel (levels before the last top level): 6 bytes of traffic per point
#define LEAN_IT (A.ch / (8 * TOP_NT))       /* A.ch is a multiple of 8 * TOP_NT * LEAN_G */
#ifndef LEAN_G
#define LEAN_G 2          /* 16-byte load pairs in flight per thread */
#endif
__global__ void __launch_bounds__(TOP_NT, 2) k_top_relabel_lean(TopArgs A) {
    __shared__ LeanTabs T;
    const int t = blockIdx.y, tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * A.ch, i1 = min(A.n, i0 + A.ch);
    const ull* keys_t = A.keys + (int64_t)t * A.L * A.ks;
    const ull* keys = keys_t + (int64_t)A.l * A.ks;
    uint16_t* lab = A.label + (int64_t)t * A.n;
    const uint16_t* pb = A.pbin + (int64_t)t * A.n;
    NodeSel* sel = A.sel + (int64_t)t * A.NTOP + A.node0;
    lean_load_tabs(A, sel, T, TOP_NT);
    __syncthreads();
    for (int it0 = 0; it0 < LEAN_IT; it0 += LEAN_G) {
        uint4 qb[LEAN_G], ql[LEAN_G];
#pragma unroll
        for (int k = 0; k < LEAN_G; ++k) {
            const int64_t i = i0 + ((int64_t)(it0 + k) * TOP_NT + tid) * 8;
            if (i < i1) {
                qb[k] = __ldcs((const uint4*)(pb + i));
                ql[k] = A.haslab ? *(const uint4*)(lab + i) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int k = 0; k < LEAN_G; ++k) {
            const int64_t i = i0 + ((int64_t)(it0 + k) * TOP_NT + tid) * 8;
            if (i < i1) {
                unsigned b[8], g[8];
                int nl[8];
                unpack8(qb[k], b); unpack8(ql[k], g);
#pragma unroll
                for (int u = 0; u < 8; ++u) nl[u] = (int)g[u] - A.node0;
                const unsigned right = lean_sides8(A, sel, T, keys, keys_t, t, i, b, nl);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if ((unsigned)nl[u] < (unsigned)A.nnodes) g[u] = A.child0 + 2 * nl[u] + ((right >> u) & 1u);
                *(uint4*)(lab + i) = pack8(g);
            }
        }
    }
}

// relabel of level l FUSED with the histogram of level l + 1 (both levels lean, the next one with a shared-memory histogram):
// one pass reads (bin_l, label_l, key_{l+1}) and writes (label_{l+1}, bin_{l+1}) -- 16 bytes per point instead of the 6 of the
// relabel pass plus the 12 of a separate histogram pass, one launch (ramp, table loads, tail) less per level, and the key load of
// the next level is in flight while the sides are decided.  The tail (flush, ticket, fused pick) is k_top_hist's.
__global__ void __launch_bounds__(TOP_NT, 2) k_top_relabel_hist(TopArgs A) {
    extern __shared__ uint32_t sh[];                             // level l + 1: 16-bit counters, two per word; then the key stage
    __shared__ LeanTabs T;
    const int t = blockIdx.y, tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * A.ch, i1 = min(A.n, i0 + A.ch);
    const ull* keys_t = A.keys + (int64_t)t * A.L * A.ks;
    const ull* keys = keys_t + (int64_t)A.l * A.ks;
    const ull* keys2 = keys + A.ks;
    uint16_t* lab = A.label + (int64_t)t * A.n;
    uint16_t* pb = A.pbin + (int64_t)t * A.n;
    NodeSel* sel = A.sel + (int64_t)t * A.NTOP + A.node0;
    const double lo2 = A.binlo[t * A.L + A.l + 1];
    const float sc2 = __double2float_rz(A.binscale[t * A.L + A.l + 1]);
    const int NB2 = A.NB2, tot2 = A.nnodes2 * NB2;
    // key stage: the 8 next-level keys of a thread's 8 points (64 bytes) travel global -> shared by cp.async one iteration
    // ahead, as four 16-byte pieces at [piece][tid] (conflict-free both ways); held in registers they would pin 16 of the 64
    // registers across the side decision and the compiler sinks the loads to their use -- three dependent latencies per iteration
    uint4* stage = (uint4*)(sh + HBINS / 2);
    const unsigned st0 = (unsigned)__cvta_generic_to_shared(stage + tid);
    auto fetch_keys = [&](int64_t i) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(st0 + (unsigned)c * TOP_NT * 16u), "l"(keys2 + i + 2 * c) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int64_t ifirst = i0 + (int64_t)tid * 8;
    if (ifirst < i1) fetch_keys(ifirst);
    uint4 qb = make_uint4(0, 0, 0, 0), ql = make_uint4(0, 0, 0, 0);
    if (ifirst < i1) { qb = __ldcs((const uint4*)(pb + ifirst)); if (A.haslab) ql = *(const uint4*)(lab + ifirst); }
    lean_load_tabs(A, sel, T, TOP_NT);
    for (int j = tid; j < (tot2 + 1) / 2; j += TOP_NT) sh[j] = 0;
    __syncthreads();
    const int nit = A.ch / (8 * TOP_NT);
    for (int it = 0; it < nit; ++it) {
        const int64_t i = i0 + ((int64_t)it * TOP_NT + tid) * 8;
        if (i >= i1) break;
        const int64_t inext = i + (int64_t)TOP_NT * 8;
        const bool more = it + 1 < nit && inext < i1;
        unsigned b[8], g[8];
        int nl[8];
        unpack8(qb, b); unpack8(ql, g);
        if (more) { qb = __ldcs((const uint4*)(pb + inext)); if (A.haslab) ql = *(const uint4*)(lab + inext); }   // next iteration's rows
#pragma unroll
        for (int u = 0; u < 8; ++u) nl[u] = (int)g[u] - A.node0;
        const unsigned right = lean_sides8(A, sel, T, keys, keys_t, t, i, b, nl);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        ull k2[8];
#pragma unroll
        for (int c = 0; c < 4; ++c) { const uint4 q = stage[c * TOP_NT + tid]; k2[2 * c] = (ull)q.x | ((ull)q.y << 32); k2[2 * c + 1] = (ull)q.z | ((ull)q.w << 32); }
        if (more) fetch_keys(inext);                              // the stage words of this thread are in registers now
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            unsigned b2 = 0;
            if ((unsigned)nl[u] < (unsigned)A.nnodes) {           // the point moves to child 2 nl + side: a node of level l + 1
                const int c = 2 * nl[u] + (int)((right >> u) & 1u);
                g[u] = (unsigned)(A.child0 + c);
                b2 = (unsigned)key_bin(k2[u], lo2, sc2, NB2);
                const int j = c * NB2 + (int)b2;
                atomicAdd(&sh[j >> 1], 1u << ((j & 1) << 4));
            }
            b[u] = b2;
        }
        *(uint4*)(lab + i) = pack8(g);
        *(uint4*)(pb + i) = pack8(b);
    }
    __syncthreads();
    uint32_t* gh = A.hist + (int64_t)t * A.HSZ;
    for (int w2 = tid; w2 < (tot2 + 1) / 2; w2 += TOP_NT) {
        const uint32_t v = sh[w2];
        if (v) atomicAdd((ull*)gh + w2, (ull)(v & 0xffffu) | ((ull)(v >> 16) << 32));
    }
    if (!A.done2) return;
    __shared__ uint32_t s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&A.done2[t], 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    TopArgs A2 = A;
    A2.l = A.l + 1; A2.node0 = A.child0; A2.nnodes = A.nnodes2; A2.NB = NB2;
    for (int nl2 = tid >> 5; nl2 < A2.nnodes; nl2 += TOP_NT / 32) pick_node_warp(A2, t, nl2);
}

// last top level: children instead of labels, and the points are placed into their child's slice of perm.  The CTA's
// chunk is bucketed by child in shared memory first, so a warp writes runs of consecutive perm entries (full sectors)
// instead of 32 scattered 4-byte stores.  Order inside a slice is irrelevant: the bottom phase sorts it.
#ifndef SCAT_NT
#define SCAT_NT 512
#endif
#ifndef SCAT_CH
#define SCAT_CH 16384     /* points per CTA: 64 KB staging buffer, two CTAs per SM */
#endif
#define SCAT_IT (SCAT_CH / (8 * SCAT_NT))
#define SCAT_PER (SCAT_MAX / SCAT_NT)
__global__ void __launch_bounds__(SCAT_NT, 2) k_top_scatter_lean(TopArgs A) {
    extern __shared__ uint32_t s_stage[];                       // [SCAT_CH] (child << 15) | local point index
    __shared__ LeanTabs T;
    __shared__ uint32_t s_cnt[SCAT_MAX], s_off[SCAT_MAX], s_dst[SCAT_MAX];
    __shared__ uint32_t s_wsum[32];
    const int t = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * SCAT_CH, i1 = min(A.n, i0 + SCAT_CH);
    const ull* keys_t = A.keys + (int64_t)t * A.L * A.ks;
    const ull* keys = keys_t + (int64_t)A.l * A.ks;
    const uint16_t* lab = A.label + (int64_t)t * A.n;
    const uint16_t* pb = A.pbin + (int64_t)t * A.n;
    NodeSel* sel = A.sel + (int64_t)t * A.NTOP + A.node0;
    const int nch = 2 * A.nnodes;
    uint4 qb[SCAT_IT], ql[SCAT_IT];
#pragma unroll
    for (int it = 0; it < SCAT_IT; ++it) {
        const int64_t i = i0 + ((int64_t)it * SCAT_NT + tid) * 8;
        if (i < i1) {
            qb[it] = __ldcs((const uint4*)(pb + i));
            ql[it] = A.haslab ? __ldcs((const uint4*)(lab + i)) : make_uint4(0, 0, 0, 0);
        }
    }
    lean_load_tabs(A, sel, T, SCAT_NT);
    for (int j = tid; j < nch; j += SCAT_NT) s_cnt[j] = 0;
    if (tid < 32) s_wsum[tid] = 0;
    __syncthreads();
    // pass 1: child of every point (kept in registers, 16 bits each), per-child counts of this chunk
    uint32_t cj[SCAT_IT][4];
#pragma unroll
    for (int it = 0; it < SCAT_IT; ++it) {
        const int64_t i = i0 + ((int64_t)it * SCAT_NT + tid) * 8;
        if (i < i1) {
            unsigned b[8], g[8];
            int nl[8];
            unpack8(qb[it], b); unpack8(ql[it], g);
#pragma unroll
            for (int u = 0; u < 8; ++u) nl[u] = (int)g[u] - A.node0;       // scatter_fast: every point sits in a node of this level
            const unsigned right = lean_sides8(A, sel, T, keys, keys_t, t, i, b, nl);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                g[u] = min(2u * (unsigned)nl[u] + ((right >> u) & 1u), (unsigned)(nch - 1));
                atomicAdd(&s_cnt[g[u]], 1u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) cj[it][u] = g[2 * u] | (g[2 * u + 1] << 16);
        }
    }
    __syncthreads();
    // exclusive scan of the counts (SCAT_PER consecutive children per thread); reserve the chunk's range in every child's slice
    {
        uint32_t c[SCAT_PER], s = 0;
#pragma unroll
        for (int e = 0; e < SCAT_PER; ++e) { const int j = SCAT_PER * tid + e; c[e] = j < nch ? s_cnt[j] : 0u; s += c[e]; }
        uint32_t incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) s_wsum[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t w = s_wsum[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
            s_wsum[lane] = wi - w;
        }
        __syncthreads();
        uint32_t ex = s_wsum[wid] + incl - s;
        uint32_t* fill = A.fill + (int64_t)t * A.NTOP;
#pragma unroll
        for (int e = 0; e < SCAT_PER; ++e) {
            const int j = SCAT_PER * tid + e;
            if (j < nch) {
                s_off[j] = ex;
                s_dst[j] = A.nstart[A.child0 + j] + (c[e] ? atomicAdd(&fill[A.child0 + j], c[e]) : 0u) - ex;
                s_cnt[j] = 0;
            }
            ex += c[e];
        }
    }
    __syncthreads();
    // pass 2: bucket the chunk by child
#pragma unroll
    for (int it = 0; it < SCAT_IT; ++it) {
        const int64_t i = i0 + ((int64_t)it * SCAT_NT + tid) * 8;
        if (i < i1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const unsigned j = (cj[it][u >> 1] >> ((u & 1) << 4)) & 0xffffu;
                const uint32_t slot = s_off[j] + atomicAdd(&s_cnt[j], 1u);
                s_stage[slot] = (j << 15) | (uint32_t)((it * SCAT_NT + tid) * 8 + u);
            }
        }
    }
    __syncthreads();
    // pass 3: consecutive threads write consecutive entries of a child's slice
    uint32_t* perm = A.perm + (int64_t)t * A.ps;
    const int npts = (int)(i1 - i0);
    for (int sl = tid; sl < npts; sl += SCAT_NT) {
        const uint32_t e = s_stage[sl];
        perm[s_dst[e >> 15] + sl] = (uint32_t)i0 + (e & 0x7fffu);
    }
}

// gather the median bins' keys: the chunk's hits are listed in shared memory, every node's range is reserved with one
// atomic per (CTA, node), then all keys are fetched at once (no atomic -> store chain inside the streaming loop)
#define CL_CAP 8192
__global__ void __launch_bounds__(TOP_NT, 3) k_top_compact_lean(TopArgs A) {
    __shared__ uint16_t s_sbin[SMEM_NODES];
    __shared__ uint32_t s_cnt[SMEM_NODES], s_base[SMEM_NODES];
    extern __shared__ uint32_t s_list[];                         // [CL_CAP] (node << 15) | local point index
    uint16_t* s_rank = (uint16_t*)(s_list + CL_CAP);             // [CL_CAP] position among the chunk's hits of that node
    __shared__ uint32_t s_n;
    const int t = blockIdx.y, tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * A.ch, i1 = min(A.n, i0 + A.ch);
    const ull* keys = A.keys + ((int64_t)t * A.L + A.l) * A.ks;
    const uint16_t* lab = A.label + (int64_t)t * A.n;
    const uint16_t* pb = A.pbin + (int64_t)t * A.n;
    NodeSel* sel = A.sel + (int64_t)t * A.NTOP + A.node0;
    ull* cand = A.cand + (int64_t)t * A.n;
    for (int j = tid; j < A.nnodes; j += TOP_NT) { s_sbin[j] = (uint16_t)sel[j].sel_bin; s_cnt[j] = 0; }
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int it0 = 0; it0 < LEAN_IT; it0 += LEAN_G) {
        uint4 qb[LEAN_G], ql[LEAN_G];
#pragma unroll
        for (int k = 0; k < LEAN_G; ++k) {
            const int64_t i = i0 + ((int64_t)(it0 + k) * TOP_NT + tid) * 8;
            if (i < i1) {
                qb[k] = *(const uint4*)(pb + i);
                ql[k] = A.haslab ? *(const uint4*)(lab + i) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int k = 0; k < LEAN_G; ++k) {
            const int64_t i = i0 + ((int64_t)(it0 + k) * TOP_NT + tid) * 8;
            if (i < i1) {
                unsigned b[8], g[8];
                unpack8(qb[k], b); unpack8(ql[k], g);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int nl = (int)g[u] - A.node0;
                    if ((unsigned)nl < (unsigned)A.nnodes && b[u] == s_sbin[nl]) {
                        const uint32_t q = atomicAdd(&s_n, 1u);
                        if (q < CL_CAP) {
                            s_rank[q] = (uint16_t)atomicAdd(&s_cnt[nl], 1u);
                            s_list[q] = ((uint32_t)nl << 15) | (uint32_t)(((it0 + k) * TOP_NT + tid) * 8 + u);
                        } else {                             // list full (a chunk dominated by median-bin points)
                            const uint32_t pos = atomicAdd(&sel[nl].cand_fill, 1u);
                            cand[sel[nl].cand_off + pos] = keys[i + u];
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int j = tid; j < A.nnodes; j += TOP_NT) {
        const uint32_t c = s_cnt[j];
        if (c) s_base[j] = sel[j].cand_off + atomicAdd(&sel[j].cand_fill, c);
    }
    __syncthreads();
    const uint32_t nq = min(s_n, (uint32_t)CL_CAP);
    for (uint32_t q = tid; q < nq; q += TOP_NT) {
        const uint32_t e = s_list[q];
        cand[s_base[e >> 15] + s_rank[q]] = keys[i0 + (e & 0x7fffu)];
    }
}

// thr / margins of the level's nodes (Internal.hs:496-501).  A batch build's top-phase nodes hold >= 3 points; the
// streaming build's chunk trees can carry tiny pieces (a short last chunk) through the same levels
__global__ void k_top_finalize(TopArgs A) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= A.nnodes * A.Tg) return;
    int t = idx / A.nnodes, nl = idx % A.nnodes, g = A.node0 + nl;
    if (A.child[g] < 0) return;
    const NodeSel& S = A.sel[(int64_t)t * A.NTOP + g];
    const uint32_t sz = A.nsize[g], nh = sz >> 1;
    ull lo = (S.cless == nh) ? S.pred : S.thr;                       // sorted[nh-1]
    ull hi = (S.cless + S.ceq >= nh + 2) ? S.thr : S.succ;           // sorted[nh+1]
    if (sz == 2) hi = S.thr;                                         // (sorted[0], sorted[1]), Internal.hs:499
    if (sz == 1) { lo = S.thr; hi = S.thr; }                         // (z, z), Internal.hs:500
    const int64_t o = (int64_t)(A.gt0 + t) * A.nn_all + g;
    A.thr[o] = ord2f(S.thr);
    A.mlo[o] = ord2f(lo);
    A.mhi[o] = ord2f(hi);
}

// jobs with several roots (one per data chunk): label[t][i] = root whose row range holds point i
__global__ void k_label_roots(uint16_t* __restrict__ label, int64_t n, int Tg, const uint32_t* __restrict__ nstart, int nroots) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = nroots - 1;                  // last root with start <= i (roots are consecutive row ranges)
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (__ldg(nstart + mid) <= (uint32_t)i) lo = mid; else hi = mid - 1; }
    for (int t = 0; t < Tg; ++t) label[(int64_t)t * n + i] = (uint16_t)lo;
}

__global__ void k_iota_perm(uint32_t* perm, int64_t n, int64_t ps, int Tg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * Tg) perm[(i / n) * ps + (i % n)] = (uint32_t)(i % n);
}

// =====================================================================================================
// bottom phase: one CTA = one level-s node and its whole subtree, resident in shared memory
// =====================================================================================================

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
    const int64_t n = A.ks;
    const ull* keys_t = A.keys + (int64_t)t * A.L * n;
    uint32_t* perm = A.perm + (int64_t)t * A.ps + start;

    for (uint32_t p = tid; p < m; p += NT) sidx[p] = perm[p];
    __syncthreads();

    if (A.s > 0 && !A.given_order) {
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
// bottom phase, fast path: uniform padded layout + packed (key48 | slot16) sort words
// =====================================================================================================
// The node's P0 = next_pow2(size) slots form a complete binary layout: at relative depth j a segment owns Pv = P0>>j
// slots, real elements first, then padding.  Every level is ONE uniform bitonic network over all slots on 64-bit
// words w = (order-preserving key with its low 16 bits replaced by the element's slot).  Comparing w compares the
// top 48 key bits, then the incoming slot -- i.e. the reference's stable sort whenever the low 16 key bits do not
// decide.  Adjacent words with equal top-48 bits (rare) are re-checked with the full keys and fixed by odd-even
// transposition, so the result is exactly Merge.sortBy (comparing snd) (Internal.hs:504-512).
// No entry tables, no bounds checks and no payload in the network: 2 LDS.64 + compare + 2 STS.64 per comparator.
#define BOT2_TAB 512        /* segment-table entries of the default instance: subtrees of <= 9 splitting levels */
#define BOT2_TAB_SHALLOW 32 /* shallow instance (<= 5 splitting levels, the usual batch build): 0.6 KB of tables instead of 9 KB per CTA */
#define BOT2_TAB_DEEP 1024  /* deep instance (streaming chunk trees, minLeaf 0/1 chains): <= 10 splitting levels */
#define W_SENT 0xffffffffffffffffull

template <int NT>
__device__ __forceinline__ void bitonic_uniform(ull* w, unsigned nslots, unsigned Pv) {
    const unsigned half = nslots >> 1;
    for (unsigned k = 2; k <= Pv; k <<= 1) {
        const int lk = ilog2_pow2(k);
        for (unsigned c = threadIdx.x; c < half; c += NT) {
            const unsigned blk = c >> (lk - 1), x = c & ((k >> 1) - 1);
            const unsigned i = (blk << lk) + x, p = (blk << lk) + (k - 1 - x);
            const ull a = w[i], b = w[p];
            if (a > b) { w[i] = b; w[p] = a; }
        }
        __syncthreads();
        for (unsigned j = k >> 2; j > 0; j >>= 1) {
            const int lj = ilog2_pow2(j);
            for (unsigned c = threadIdx.x; c < half; c += NT) {
                const unsigned i = ((c >> lj) << (lj + 1)) + (c & (j - 1)), p = i + j;
                const ull a = w[i], b = w[p];
                if (a > b) { w[i] = b; w[p] = a; }
            }
            __syncthreads();
        }
    }
}

// Sort word W: uint64 = (top 48 key bits | 16-bit slot), or -- for P0 <= 2048 -- uint32 = (PB-bit key prefix | lp0-bit slot)
// with the prefix taken from the key's position inside the (tree, level) key range.  Elements whose prefixes collide are
// put in exact (full key, incoming slot) order by the neighbour fix-up below, so both forms give the same result; the
// 32-bit form halves the compare/select, shuffle and shared-memory work of the network.
#ifndef RPF_BOT_MINB128
#define RPF_BOT_MINB128 10     /* resident CTAs per SM asked of the 128-thread instances (caps them at 48 registers) */
#endif
#ifndef RPF_BOT_MINB128S
#define RPF_BOT_MINB128S 12    /* ... of the shallow one: 40 registers (16 bytes of spills); the kernel waits on its MIO queue (shuffles + shared
                                  memory), two more resident CTAs cover more of that: bottom 1.98 -> 1.93 ms (8: 2.08 ms; 16 would spill 248 bytes) */
#endif
template <int NT, int TAB, typename W>
__global__ void __launch_bounds__(NT, (NT == 128 ? (TAB == BOT2_TAB_SHALLOW ? RPF_BOT_MINB128S : RPF_BOT_MINB128) : 1)) k_bottom3(BottomArgs A) {
    constexpr unsigned P0 = 8 * NT;               // slots (>= node size), 8 per thread
    constexpr int lp0 = (NT == 32 ? 8 : NT == 64 ? 9 : NT == 128 ? 10 : NT == 256 ? 11 : NT == 512 ? 12 : 13);
    constexpr bool W32 = sizeof(W) == 4;
    constexpr int SB = W32 ? lp0 : 16;            // slot bits of a sort word
    constexpr int PB = 32 - lp0;                  // key-prefix bits of a 32-bit word
    constexpr unsigned SMASK = (1u << SB) - 1u;
    constexpr W WSENT = (W)~(W)0;
    static_assert(!W32 || lp0 <= 11, "32-bit sort words need at least 21 prefix bits");
    // Shallow instance (<= 5 splitting levels): a segment keeps >= 16 slots and a child >= 8, so the 8 slots of a thread always
    // share one segment and one child -- segment size, median rank and table lookups are per THREAD, not per slot (the
    // per-slot form spent 60 % of the kernel's instructions outside the sort network: ncu source page, round 2).
    constexpr bool UNI = (TAB == BOT2_TAB_SHALLOW);
    static_assert(!UNI || (P0 >> 5) >= 8, "uniform-thread path needs 8 slots per child at the deepest level");
    extern __shared__ unsigned char smraw[];
    W* w = (W*)smraw;                             // [P0] sort words (the block is sized for 8-byte words: composite_sort
    ull* w64 = (ull*)smraw;                       //      uses it as linear uint64 scratch)
    uint32_t* sidx = (uint32_t*)(w64 + P0);       // [P0] row id held by each slot
    __shared__ uint16_t t_sz[2][TAB];        // size of the segment if it splits at this level, else 0
    __shared__ uint16_t t_ps[2][TAB];        // offset of the segment inside this CTA's slice of perm
    __shared__ int32_t t_gid[2][TAB];        // BFS id
    __shared__ uint16_t t_lsz[TAB];          // size if the segment just became a Tip (to be emitted), else 0
    __shared__ double s_plo[16];             // per level of this subtree: low end of the (tree, level) key range and the fp32
    __shared__ float s_psc[16];              //   prefix scale of the 32-bit sort words (one fp64 division per level and CTA)
    __shared__ uint32_t s_trash[UNI ? NT : 1];
    const int t = blockIdx.y, tid = threadIdx.x;
    if (A.only && !A.only[(size_t)t * gridDim.x + blockIdx.x]) return;      // second pass of k_bottom4: flagged nodes only
    const int e0 = A.first_gid + blockIdx.x;
    const uint32_t m = A.nsize[e0], start = A.nstart[e0];
    if (m == 0) return;
    const int64_t n = A.ks;
    const ull* keys_t = A.keys + (int64_t)t * A.L * n;
    uint32_t* perm = A.perm + (int64_t)t * A.ps + start;
    const bool root_internal = A.child[e0] >= 0;
    const bool unordered = A.s > 0 && !A.given_order;   // the slots do not arrive in the reference's order

    // Shared-memory layout: slot x of `w` and `sidx` lives at TI(x) = (x & 7) * NT + (x >> 3).  Thread t owns slots
    // 8t .. 8t+7, i.e. TI = q * NT + t: every warp-wide access to "my q-th slot" touches 32 consecutive words (no bank
    // conflicts; the linear layout made each of those a 4- to 16-way conflict -- 47 % of all wavefronts in ncu).
    auto TI = [](unsigned x) -> unsigned { return (x & 7u) * NT + (x >> 3); };
#pragma unroll
    for (int q = 0; q < 8; ++q) { const uint32_t p = 8u * tid + q; sidx[q * NT + tid] = p < m ? perm[p] : 0u; }
    if (W32 && tid < 16) {
        double lo = 0.0; float scf = 0.f;
        const int l = A.s + tid;
        if (A.kmin && l < A.L) {
            lo = ord2f(A.kmin[t * A.L + l]);
            const double wdt = ord2f(A.kmax[t * A.L + l]) - lo;
            const double psc = (wdt > 0.0 && isfinite(wdt)) ? (double)((1u << PB) - 2u) / wdt : 0.0;
            scf = isfinite(psc) ? __double2float_rz(psc) : 0.f;
            if (!isfinite(scf)) scf = 0.f;
        }
        s_plo[tid] = lo; s_psc[tid] = scf;
    }
    __syncthreads();

    // composite order (key_{s-1}, ..., key_0, row id) of the whole node: only needed when the node arrives unordered
    // (s > 0) AND either it is a Tip itself or its first sort hits a tie in the top 48 key bits.
    auto composite_sort = [&]() {
        const ull* k1 = keys_t + (int64_t)(A.s - 1) * n;
        for (uint32_t p = tid; p < m; p += NT) w64[p] = k1[sidx[TI(p)]];    // linear uint64 scratch inside this (rare) path
        __syncthreads();
        auto after = [&](ull ka, uint32_t ia, ull kb, uint32_t ib) -> bool {
            if (ka != kb) return ka > kb;
            for (int lvl = A.s - 2; lvl >= 0; --lvl) {
                const ull xa = keys_t[(int64_t)lvl * n + ia], xb = keys_t[(int64_t)lvl * n + ib];
                if (xa != xb) return xa > xb;
            }
            return ia > ib;
        };
        const unsigned Pm = next_pow2_u32(m), half = Pm >> 1;
        for (unsigned k = 2; k <= Pm; k <<= 1) {
            const int lk = ilog2_pow2(k);
            for (unsigned c = tid; c < half; c += NT) {
                const unsigned blk = c >> (lk - 1), x = c & ((k >> 1) - 1);
                const unsigned i = (blk << lk) + x, p = (blk << lk) + (k - 1 - x);
                if (p < m) {
                    const ull a = w64[i], b = w64[p]; const uint32_t ia = sidx[TI(i)], ib = sidx[TI(p)];
                    if (after(a, ia, b, ib)) { w64[i] = b; w64[p] = a; sidx[TI(i)] = ib; sidx[TI(p)] = ia; }
                }
            }
            __syncthreads();
            for (unsigned j = k >> 2; j > 0; j >>= 1) {
                const int lj = ilog2_pow2(j);
                for (unsigned c = tid; c < half; c += NT) {
                    const unsigned i = ((c >> lj) << (lj + 1)) + (c & (j - 1)), p = i + j;
                    if (p < m) {
                        const ull a = w64[i], b = w64[p]; const uint32_t ia = sidx[TI(i)], ib = sidx[TI(p)];
                        if (after(a, ia, b, ib)) { w64[i] = b; w64[p] = a; sidx[TI(i)] = ib; sidx[TI(p)] = ia; }
                    }
                }
                __syncthreads();
            }
        }
    };

    if (!root_internal) {        // the node is a Tip: its points must simply be in the reference's order
        if (unordered) { composite_sort(); for (uint32_t p = tid; p < m; p += NT) perm[p] = sidx[TI(p)]; }
        return;
    }
    if (tid == 0) { t_sz[0][0] = (uint16_t)m; t_ps[0][0] = 0; t_gid[0][0] = e0; }
    __syncthreads();

    int cur = 0;
    bool need_check_order = unordered;     // slots are not yet in the reference's incoming order
    const unsigned x0 = 8u * tid;          // first slot owned by this thread
    for (int j = 0;; ++j) {
        const int l = A.s + j;
        const unsigned Pv = P0 >> j, nseg = 1u << j;
        const int lpv = lp0 - j;
        const ull* kl = keys_t + (int64_t)l * n;
        const uint16_t* sz = t_sz[cur];

        // ---- gather keys into registers, build sort words, sort, store back in slot order
        W v[8];                                           // this thread's 8 sort words (kept in registers across the level)
        // 32-bit words: prefix = position of the key's VALUE inside the (tree, level) key range, PB bits (monotone: every
        // step below is monotone non-decreasing in the key).  A prefix taken from the key's bit pattern would spend
        // almost all codes on magnitudes near zero, where no projections live.
        // (the difference is taken in fp64 -- keys far from zero keep their resolution -- the scaling in fp32: a 24-bit mantissa
        //  against <= 24 prefix bits, 3 instructions instead of an fp64 multiply, compare / select and fp64 -> int conversion)
        const double plo = W32 ? s_plo[j & 15] : 0.0;
        const float pscf = W32 ? s_psc[j & 15] : 0.f;
        auto make_word = [&](ull key, unsigned slot) -> W {
            if (W32) {
                const float fv = __double2float_rz(ord2f(key) - plo) * pscf;
                const unsigned p = min(__float2uint_rz(fv), (1u << PB) - 2u);      // the conversion saturates: negative (and NaN) -> 0
                return (W)((p << SB) | slot);
            }
            return (W)((key & ~0xffffull) | slot);
        };
        auto gather_and_sort = [&]() {
            const unsigned e = x0 >> lpv;                 // all 8 slots share a segment when Pv >= 8
            if (Pv >= 8) {
                const unsigned se = sz[e], i0 = x0 & (Pv - 1);
                uint32_t ids[8];
                ull kk[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) ids[q] = sidx[q * NT + tid];
#pragma unroll
                for (int q = 0; q < 8; ++q) kk[q] = (i0 + q < se) ? kl[ids[q]] : 0ull;      // 8 independent loads in flight
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = (i0 + q < se) ? make_word(kk[q], x0 + q) : WSENT;
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned x = x0 + q, eq = x >> lpv, i = x & (Pv - 1);
                    v[q] = (i < sz[eq]) ? make_word(kl[sidx[q * NT + tid]], x) : WSENT;
                }
            }
            __syncthreads();                              // every thread has read sidx/w before w is reused
            sort_regs<NT, W>(v, w, Pv);
            __syncthreads();                              // other threads may still be reading the transposed staging
#pragma unroll
            for (int q = 0; q < 8; ++q) w[q * NT + tid] = v[q];
            __syncthreads();
        };
        gather_and_sort();

        // ---- neighbours with equal key prefixes: re-check with the full keys
        // (UNI) pairs (x0+q, x0+q+1) of this thread whose prefixes collide, bit q; swaps between equal prefixes leave it unchanged
        unsigned cmask = 0;
        auto tie_flag = [&]() -> int {
            if (UNI) {
                const int lim = (int)sz[x0 >> lpv] - (int)(x0 & (Pv - 1));      // slots q < lim of this thread hold elements
                cmask = 0;
#pragma unroll
                for (int q = 0; q < 7; ++q) cmask |= (unsigned)((q + 1 < lim) & ((v[q] >> SB) == (v[q + 1] >> SB))) << q;
                if (8 < lim) cmask |= (unsigned)((v[7] >> SB) == (w[tid + 1] >> SB)) << 7;   // slot x0+8: same segment, thread tid+1
                return __syncthreads_or((int)cmask);
            }
            int f = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const unsigned slot = x0 + q;
                if (slot + 1 < P0) {
                    const unsigned e = slot >> lpv, i = slot & (Pv - 1);
                    const W nxt = q < 7 ? v[q < 7 ? q + 1 : 7] : w[tid + 1];            // slot + 1 = first slot of thread tid + 1
                    if (i + 1 < sz[e]) f |= ((v[q] >> SB) == (nxt >> SB));
                }
            }
            return __syncthreads_or(f);
        };
        // exact comparator on neighbours with equal prefixes: (full key, incoming slot); odd-even transposition until
        // stable.  Returns whether two compared elements had EQUAL full keys (only then does the incoming order matter).
        auto fix_up = [&]() -> int {
            int eqfull = 0;
            while (UNI) {              // each thread walks its own colliding pairs only (most threads have none)
                int swapped = 0;
                for (int par = 0; par < 2; ++par) {
                    if (cmask & (0x55u << par)) {
#pragma unroll
                        for (int q2 = 0; q2 < 4; ++q2) {
                            const int q = 2 * q2 + par;
                            if (cmask & (1u << q)) {
                                const unsigned slot = x0 + q;
                                const W a = w[TI(slot)], b = w[TI(slot + 1)];
                                const unsigned sa = (unsigned)(a & SMASK), sb2 = (unsigned)(b & SMASK);
                                const ull fa = kl[sidx[TI(sa)]], fb = kl[sidx[TI(sb2)]];
                                eqfull |= (fa == fb);
                                if (fa > fb || (fa == fb && sa > sb2)) { w[TI(slot)] = b; w[TI(slot + 1)] = a; swapped = 1; }
                            }
                        }
                    }
                    __syncthreads();
                }
                if (!__syncthreads_or(swapped)) return __syncthreads_or(eqfull);
            }
            while (!UNI) {
                int swapped = 0;
                for (int par = 0; par < 2; ++par) {
                    for (unsigned c = tid; c < (P0 >> 1); c += NT) {
                        const unsigned slot = 2 * c + par;
                        if (slot + 1 < P0) {
                            const unsigned e = slot >> lpv, i = slot & (Pv - 1);
                            if (i + 1 < sz[e]) {
                                const W a = w[TI(slot)], b = w[TI(slot + 1)];
                                if ((a >> SB) == (b >> SB)) {
                                    const unsigned sa = (unsigned)(a & SMASK), sb2 = (unsigned)(b & SMASK);
                                    const ull fa = kl[sidx[TI(sa)]], fb = kl[sidx[TI(sb2)]];
                                    eqfull |= (fa == fb);
                                    if (fa > fb || (fa == fb && sa > sb2)) { w[TI(slot)] = b; w[TI(slot + 1)] = a; swapped = 1; }
                                }
                            }
                        }
                    }
                    __syncthreads();
                }
                if (!__syncthreads_or(swapped)) break;
            }
            return __syncthreads_or(eqfull);
        };
        if (tie_flag()) {
            const int eqfull = fix_up();
            if (eqfull && need_check_order) {
                // equal keys, and the incoming slot order was arbitrary: establish the reference's order, redo the level
                composite_sort();
                gather_and_sort();
                if (tie_flag()) fix_up();
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = w[q * NT + tid];      // the registers follow the repaired order
        }
        need_check_order = false;

        // ---- thresholds / margins at the sorted positions (Internal.hs:496-503)
        for (unsigned e = tid; e < nseg; e += NT) {
            const unsigned se = sz[e];
            if (!se) continue;
            const unsigned off = e << lpv, nh = se >> 1;
            auto full = [&](unsigned slot) { return kl[sidx[TI((unsigned)(w[TI(slot)] & SMASK))]]; };
            const ull th = full(off + nh);
            ull ml, mh;
            if (se >= 3) { ml = full(off + nh - 1); mh = full(off + nh + 1); }
            else if (se == 2) { ml = full(off); mh = full(off + 1); }
            else { ml = th; mh = th; }
            const int64_t o = (int64_t)(A.gt0 + t) * A.nn_all + t_gid[cur][e];
            A.thr[o] = ord2f(th); A.mlo[o] = ord2f(ml); A.mhi[o] = ord2f(mh);
        }

        // ---- children tables
        const int nxt = cur ^ 1;
        const bool can_grow = (nseg * 2 <= (unsigned)TAB) && Pv >= 2;
        int any_internal = 0;
        if (can_grow) {
            for (unsigned c = tid; c < nseg * 2; c += NT) {
                const unsigned e = c >> 1, ps = sz[e];
                uint16_t csz = 0, lsz = 0, cps = 0; int32_t cg = -1;
                if (ps) {
                    const unsigned nh = ps >> 1;
                    const unsigned s2 = (c & 1) ? ps - nh : nh;
                    cps = (uint16_t)(t_ps[cur][e] + ((c & 1) ? nh : 0));
                    cg = A.child[t_gid[cur][e]] + (int)(c & 1);
                    if (A.child[cg] >= 0) { csz = (uint16_t)s2; any_internal = 1; }
                    else lsz = (uint16_t)s2;
                }
                t_sz[nxt][c] = csz; t_ps[nxt][c] = cps; t_gid[nxt][c] = cg; t_lsz[c] = lsz;
            }
        }
        // ---- move the row ids: left half stays, right half starts at the middle of the segment
        if (UNI) {
            const unsigned i0 = x0 & (Pv - 1), se = sz[x0 >> lpv], nh = se >> 1;
            const int lim = (int)se - (int)i0;
            const unsigned hd = (Pv >> 1) - nh;             // an element of rank i >= nh moves hd slots up
            uint32_t val[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) val[q] = sidx[TI((unsigned)(v[q] & SMASK) & (P0 - 1u))];   // (padding words: any slot in range)
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 8; ++q) {       // branch-free: padding slots store to a per-thread dummy word
                const unsigned di = (i0 + q < nh) ? (unsigned)(q * NT + tid) : TI(x0 + q + hd);
                uint32_t* dp = (q < lim) ? sidx + di : s_trash + tid;
                *dp = val[q];
            }
            // (tried: prefetch.global.L2 of the keys two levels ahead while the ids move -- the node's point set is the same at
            //  every level of the subtree; bottom 1.95 -> 1.98 ms, the gather latency is already hidden by the 10 resident CTAs)
        } else {
            uint32_t val[8]; uint32_t dst[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const unsigned slot = x0 + q;
                dst[q] = 0xffffffffu;
                const unsigned e = slot >> lpv, i = slot & (Pv - 1), se = sz[e];
                if (i < se) {
                    const unsigned nh = se >> 1;
                    val[q] = sidx[TI((unsigned)(v[q] & SMASK))];
                    dst[q] = i < nh ? slot : (e << lpv) + (Pv >> 1) + (i - nh);
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 8; ++q) if (dst[q] != 0xffffffffu) sidx[TI(dst[q])] = val[q];
        }
        any_internal = __syncthreads_or(any_internal);
        if (!can_grow) break;      // unreachable for the shapes the host routes here
        // ---- emit the children that are Tips
        if (UNI) {
            const unsigned Pc = Pv >> 1, c = x0 >> (lpv - 1), ic0 = x0 & (Pc - 1), lsz = t_lsz[c];
            if (ic0 < lsz) {
                uint32_t* dstp = perm + t_ps[nxt][c] + ic0;
#pragma unroll
                for (int q = 0; q < 8; ++q) if (ic0 + q < lsz) dstp[q] = sidx[q * NT + tid];
            }
        } else {
            const unsigned Pc = Pv >> 1;
            const int lpc = lpv - 1;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const unsigned slot = x0 + q;
                const unsigned c = slot >> lpc, i = slot & (Pc - 1);
                if (i < t_lsz[c]) perm[t_ps[nxt][c] + i] = sidx[q * NT + tid];
            }
        }
        __syncthreads();
        cur = nxt;
        if (!any_internal) break;
    }
}

// =====================================================================================================
// k_bottom4: one WARP owns a level-s node (<= 1024 points) and its subtree -- no block barriers at all
// =====================================================================================================
// What k_bottom3 pays for is a full sort of every segment at every level (164 network stages for the four bottom levels of
// configs[1], ~270 thread instructions per point and level, 3.7 barrier stalls per issue).  The reference only needs
//   * at a level whose children all split again: the elements of rank nh-1, nh, nh+1 of the stable sort (threshold and margin,
//     Internal.hs:496-501) and WHICH elements go left -- the order inside a child matters only as the tie-break of a later
//     sort, i.e. only if two full keys are equal later on;
//   * at a level where Tips form: the sorted order (it is the Tip's content order).
// The warp keeps three 1024-entry arrays in shared memory: the row ids in slot order (two buffers, ping-pong) and the level's
// 32-bit sort words (22-bit prefix = position of the key's value in the (tree, level) key range | 10-bit slot, as in
// k_bottom3).  A segment of Pv = 1024 >> j slots at relative depth j is read in chunks of 128 slots, four consecutive slots
// per lane (one 16-byte load).
//   select level (Pv >= 256): bisection on the prefix for the class of rank nh -- 22 steps, each one compare-and-count pass
//     over the words + one REDUX per segment, the segments of the level interleaved -- then the median / predecessor / successor
//     words by REDUX.MIN/MAX and a ballot compaction of the row ids into the children's slots.  Exact iff the prefix classes of
//     ranks nh-1, nh, nh+1 are singletons (different prefixes order like the full keys: the map is monotone) -- checked.
//   sort level (Pv <= 128): bitonic network per 128-slot chunk (distances 1, 2 in registers, 4 .. 64 by shuffles); neighbours
//     with equal prefixes are put in full-key order afterwards.
// Anything else -- a prefix class with several members where it matters, equal FULL keys (the reference's incoming order is
// needed, which the select levels do not keep), a node that is a Tip itself but arrives unordered, segments of fewer than 3
// points, Tips larger than 64 points -- sets the node's `redo` flag after putting the row ids of the unfinished segments back
// into their ranges of perm (the node's slice is again a permutation of its points); k_bottom3 then runs on the flagged nodes
// only and, being exact on any input, produces the same forest.
#define B4_WPC 4
struct B4Warp {
    uint32_t ids[2][1024];
    uint32_t words[1024];
    int32_t gid[2][32];
    uint16_t sz[2][32], ps[2][32], lsz[32];
};

__global__ void __launch_bounds__(B4_WPC * 32, 4) k_bottom4(BottomArgs A, uint8_t* __restrict__ redo, int nroots) {
    extern __shared__ __align__(16) unsigned char b4raw[];
    constexpr unsigned PAD = 0xffffffffu, FULL = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int node = blockIdx.x * B4_WPC + (int)wib, t = blockIdx.y;
    if (node >= nroots) return;
    B4Warp& S = ((B4Warp*)b4raw)[wib];
    const int e0 = A.first_gid + node;
    const uint32_t m = A.nsize[e0], start = A.nstart[e0];
    if (m == 0) return;
    const int64_t n = A.ks;
    const ull* keys_t = A.keys + (int64_t)t * A.L * n;
    uint32_t* perm = A.perm + (int64_t)t * A.ps + start;
    uint8_t* flag = redo + (size_t)t * nroots + node;
    if (A.child[e0] < 0) {                       // the node is a Tip: only its order may be open (k_bottom3 knows how)
        if (A.s > 0 && lane == 0) *flag = 1;
        return;
    }
    for (unsigned x = lane; x < 1024u; x += 32) S.ids[0][x] = x < m ? perm[x] : 0u;
    if (lane == 0) { S.sz[0][0] = (uint16_t)m; S.ps[0][0] = 0; S.gid[0][0] = e0; }
    __syncwarp();
    uint32_t* words = S.words;
    const unsigned lt = (1u << lane) - 1u;
    int cur = 0, rc = 1;
    for (int J = 0; J < 5 && rc == 1; ++J) {
        const int nseg = 1 << J, lpv = 10 - J, l = A.s + J, nxt = cur ^ 1;
        const unsigned Pv = 1024u >> J;
        const ull* kl = keys_t + (int64_t)l * n;
        const uint16_t* sz = S.sz[cur];
        const uint32_t* ids = S.ids[cur];
        uint32_t* idn = S.ids[nxt];
        // ---- children tables (sizes only)
        int anyint = 0, anytip = 0, bail = 0;
        if (lane < 2u * nseg) {
            const unsigned c = lane, e = c >> 1, psz = sz[e];
            uint16_t csz = 0, lsz = 0, cps = 0; int32_t cg = -1;
            if (psz) {
                const unsigned nh = psz >> 1, s2 = (c & 1) ? psz - nh : nh;
                cps = (uint16_t)(S.ps[cur][e] + ((c & 1) ? nh : 0));
                cg = A.child[S.gid[cur][e]] + (int)(c & 1);
                if (A.child[cg] >= 0) { csz = (uint16_t)s2; anyint = 1; } else { lsz = (uint16_t)s2; anytip = 1; }
                if (psz < 3) bail = 1;
            }
            S.sz[nxt][c] = csz; S.ps[nxt][c] = cps; S.gid[nxt][c] = cg; S.lsz[c] = lsz;
        }
        anyint = __any_sync(FULL, anyint); anytip = __any_sync(FULL, anytip); bail = __any_sync(FULL, bail);
        const bool do_sort = anytip || Pv <= 128;
        if (do_sort && Pv > 128) bail = 1;        // Tips of more than 64 points: k_bottom3
        __syncwarp();
        if (!bail) {
            // ---- sort words of the level: prefix = floor((key - lo) * scale), the difference in fp64 (keys far from zero keep
            //      their resolution), the scaling in fp32 (24-bit mantissa against 22 prefix bits); every step is monotone
            const double plo = ord2f(A.kmin[t * A.L + l]);
            float scf = 0.f;
            {
                const double wdt = ord2f(A.kmax[t * A.L + l]) - plo;
                const double sc = (wdt > 0.0 && isfinite(wdt)) ? (double)((1u << 22) - 2u) / wdt : 0.0;
                scf = isfinite(sc) ? __double2float_rz(sc) : 0.f;
                if (!isfinite(scf)) scf = 0.f;
            }
#pragma unroll 1
            for (unsigned c = 0; c < 8; c += 2) {             // 8 key gathers in flight per lane
                uint4 idv[2]; ull kk[8]; bool ok[8];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) idv[h2] = *reinterpret_cast<const uint4*>(ids + 128u * (c + h2) + 4u * lane);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned x = 128u * (c + (q >> 2)) + 4u * lane + (q & 3);
                    const uint32_t id = (q & 3) == 0 ? idv[q >> 2].x : (q & 3) == 1 ? idv[q >> 2].y : (q & 3) == 2 ? idv[q >> 2].z : idv[q >> 2].w;
                    ok[q] = (x & (Pv - 1u)) < sz[x >> lpv];
                    kk[q] = ok[q] ? kl[id] : ~0ull;
                }
                uint32_t wv[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const unsigned x = 128u * (c + (q >> 2)) + 4u * lane + (q & 3);
                    const float fv = __double2float_rz(ord2f(kk[q]) - plo) * scf;
                    const unsigned p = min(__float2uint_rz(fv), (1u << 22) - 2u);      // the conversion saturates: negative (and NaN) -> 0
                    wv[q] = ok[q] ? ((p << 10) | x) : PAD;
                }
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2)
                    *reinterpret_cast<uint4*>(words + 128u * (c + h2) + 4u * lane) = make_uint4(wv[4 * h2], wv[4 * h2 + 1], wv[4 * h2 + 2], wv[4 * h2 + 3]);
            }
            __syncwarp();
        }
        if (!bail && !do_sort) {
            // ================= select level: nseg <= 4 segments of cps = Pv / 128 >= 2 chunks each
            const unsigned cps = Pv >> 7;
            unsigned lo[4], hi[4], cl[4], chi[4], nhv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { const unsigned se = e < nseg ? sz[e] : 0u; lo[e] = 0u; hi[e] = (1u << 22) - 1u; cl[e] = 0u; chi[e] = se; nhv[e] = se >> 1; }
#pragma unroll 1
            for (int it = 0; it < 22; ++it) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (e < nseg) {
                        const unsigned mid = (lo[e] + hi[e]) >> 1, lim = mid << 10;
                        unsigned c0 = 0, c1 = 0;
                        const uint4* wp = reinterpret_cast<const uint4*>(words + (unsigned)e * Pv) + lane;
#pragma unroll 2
                        for (unsigned q = 0; q < cps; ++q) {
                            const uint4 v = wp[32u * q];
                            c0 += (v.x < lim) + (v.z < lim);
                            c1 += (v.y < lim) + (v.w < lim);
                        }
                        const unsigned c = __reduce_add_sync(FULL, c0 + c1);
                        if (c <= nhv[e]) { lo[e] = mid; cl[e] = c; } else { hi[e] = mid; chi[e] = c; }
                    }
                }
            }
            unsigned medw[4], predw[4], succw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                medw[e] = PAD; predw[e] = 0u; succw[e] = PAD;
                if (e < nseg && sz[e]) {
                    const unsigned lim0 = lo[e] << 10, lim1 = (lo[e] + 1u) << 10;
                    unsigned a = PAD, b = 0u, c2 = PAD;
                    const uint4* wp = reinterpret_cast<const uint4*>(words + (unsigned)e * Pv) + lane;
#pragma unroll 2
                    for (unsigned q = 0; q < cps; ++q) {
                        const uint4 v = wp[32u * q];
                        const unsigned vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            a = min(a, vv[u] >= lim0 ? vv[u] : PAD);
                            b = max(b, vv[u] < lim0 ? vv[u] : 0u);
                            c2 = min(c2, vv[u] >= lim1 ? vv[u] : PAD);
                        }
                    }
                    medw[e] = __reduce_min_sync(FULL, a); predw[e] = __reduce_max_sync(FULL, b); succw[e] = __reduce_min_sync(FULL, c2);
                    // singleton prefix classes at ranks nh-1, nh, nh+1
                    const unsigned limp = predw[e] & ~1023u, lims = (succw[e] | 1023u) + 1u;
                    unsigned cp = 0, cs = 0;
#pragma unroll 2
                    for (unsigned q = 0; q < cps; ++q) {
                        const uint4 v = wp[32u * q];
                        cp += (v.x < limp) + (v.y < limp) + (v.z < limp) + (v.w < limp);
                        cs += (v.x < lims) + (v.y < lims) + (v.z < lims) + (v.w < lims);
                    }
                    cp = __reduce_add_sync(FULL, cp); cs = __reduce_add_sync(FULL, cs);
                    if (chi[e] - cl[e] != 1u || cl[e] != nhv[e] || succw[e] == PAD || cp + 1u != cl[e] || cs != chi[e] + 1u) bail = 1;
                }
            }
            if (!bail) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (e < nseg && sz[e]) {
                        if (lane == 0) {
                            const int64_t o = (int64_t)(A.gt0 + t) * A.nn_all + S.gid[cur][e];
                            A.thr[o] = ord2f(kl[ids[medw[e] & 1023u]]);
                            A.mlo[o] = ord2f(kl[ids[predw[e] & 1023u]]);
                            A.mhi[o] = ord2f(kl[ids[succw[e] & 1023u]]);
                        }
                        unsigned baseL = (unsigned)e * Pv, baseR = (unsigned)e * Pv + Pv / 2;
                        const uint4* wp = reinterpret_cast<const uint4*>(words + (unsigned)e * Pv) + lane;
                        const uint4* ip = reinterpret_cast<const uint4*>(ids + (unsigned)e * Pv) + lane;
                        const unsigned mw = medw[e];
#pragma unroll 1
                        for (unsigned q = 0; q < cps; ++q) {
                            const uint4 v = wp[32u * q], iv = ip[32u * q];
                            const unsigned vv[4] = {v.x, v.y, v.z, v.w}, ii[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const bool valid = vv[u] != PAD, isL = vv[u] < mw;
                                const unsigned bL = __ballot_sync(FULL, isL), bR = __ballot_sync(FULL, valid && !isL);
                                const unsigned dst = isL ? baseL + __popc(bL & lt) : baseR + __popc(bR & lt);
                                if (valid) idn[dst] = ii[u];
                                baseL += __popc(bL); baseR += __popc(bR);
                            }
                        }
                    }
                }
            }
        } else if (!bail) {
            // ================= sort level: every aligned block of Pv <= 128 slots ascending, one 128-slot chunk at a time
#pragma unroll 1
            for (unsigned c = 0; c < 8; ++c) {
                const uint4 v = *reinterpret_cast<const uint4*>(words + 128u * c + 4u * lane);
                if (__all_sync(FULL, v.x == PAD)) continue;                  // chunk without elements (real elements come first)
                unsigned w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (unsigned k = 2; k <= 128; k <<= 1) {
                    if (k <= Pv) {
#pragma unroll
                        for (unsigned j = k >> 1; j >= 1; j >>= 1) {
                            if (j >= 4) {
                                const unsigned lj = j >> 2;
                                const bool lower = (lane & lj) == 0, asc = k == Pv || ((4u * lane) & k) == 0;
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const unsigned o = __shfl_xor_sync(FULL, w4[u], lj);
                                    w4[u] = (asc == lower) ? min(w4[u], o) : max(w4[u], o);
                                }
                            } else {
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int u2 = u ^ (int)j;
                                    if (u < u2) {
                                        const bool asc = k == Pv || ((4u * lane + u) & k) == 0;
                                        const unsigned a = w4[u], b = w4[u2];
                                        w4[u] = asc ? min(a, b) : max(a, b);
                                        w4[u2] = asc ? max(a, b) : min(a, b);
                                    }
                                }
                            }
                        }
                    }
                }
                // the row ids in sorted order, left / right halves at the children's slots; tie candidates flagged
                const unsigned nx0 = __shfl_down_sync(FULL, w4[0], 1);
                int tie = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned x = 128u * c + 4u * lane + u, e = x >> lpv, i = x & (Pv - 1u), se = sz[e], nh = se >> 1;
                    if (i < se) idn[i < nh ? x : e * Pv + Pv / 2 + (i - nh)] = ids[w4[u] & 1023u];
                    const unsigned nx = u < 3 ? w4[u < 3 ? u + 1 : 3] : (lane < 31 ? nx0 : PAD);
                    if (i + 1u < se && (w4[u] >> 10) == (nx >> 10)) tie = 1;
                }
                // the sorted words go back (the fix-up below finds the tie positions in them)
                *reinterpret_cast<uint4*>(words + 128u * c + 4u * lane) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                if (tie) bail = 2;                                           // (2: not a bail yet)
            }
            __syncwarp();
            if (__any_sync(FULL, bail == 2)) {
                // neighbours with equal prefixes: their order is decided by the full keys (a swap exchanges two elements of the
                // SAME prefix, so the tie positions stay what they are); equal FULL keys need the reference's incoming order
                bail = 0;
                auto slot_of = [&](unsigned e, unsigned i, unsigned se) { const unsigned nh = se >> 1; return i < nh ? e * Pv + i : e * Pv + Pv / 2 + (i - nh); };
                for (int guard = 0; guard < 128 && !bail; ++guard) {
                    int swapped = 0;
                    for (int par = 0; par < 2; ++par) {
                        for (unsigned x = 2u * lane + (unsigned)par; x + 1u < 1024u; x += 64u) {
                            const unsigned e = x >> lpv, i = x & (Pv - 1u), se = sz[e];
                            if (i + 1u < se && (words[x] >> 10) == (words[x + 1u] >> 10)) {
                                const unsigned sa = slot_of(e, i, se), sb = slot_of(e, i + 1u, se);
                                const uint32_t ia = idn[sa], ib = idn[sb];
                                const ull fa = kl[ia], fb = kl[ib];
                                if (fa == fb) bail = 1;
                                else if (fa > fb) { idn[sa] = ib; idn[sb] = ia; swapped = 1; }
                            }
                        }
                        __syncwarp();
                    }
                    bail = __any_sync(FULL, bail);
                    if (!__any_sync(FULL, swapped)) break;
                }
            }
            if (!bail) {
                if (lane < (unsigned)nseg && sz[lane]) {
                    const unsigned e = lane, se = sz[e], nh = se >> 1, oL = e * Pv, oR = e * Pv + Pv / 2;
                    const ull th = kl[idn[oR]];
                    const ull ml = kl[idn[oL + nh - 1]], mh = kl[idn[oR + 1]];      // se >= 3 (checked above)
                    const int64_t o = (int64_t)(A.gt0 + t) * A.nn_all + S.gid[cur][e];
                    A.thr[o] = ord2f(th); A.mlo[o] = ord2f(ml); A.mhi[o] = ord2f(mh);
                }
            } else {
                // the segments' row ids sit in idn (sorted by prefix): hand them back from there
                for (int e = 0; e < nseg; ++e) {
                    const unsigned se = sz[e], nh = se >> 1;
                    for (unsigned i = lane; i < se; i += 32) perm[S.ps[cur][e] + i] = idn[i < nh ? (unsigned)e * Pv + i : (unsigned)e * Pv + Pv / 2 + (i - nh)];
                }
                rc = 2;
                break;
            }
        }
        bail = __any_sync(FULL, bail);
        if (bail) {
            // the unfinished segments' row ids go back to their ranges of perm: the slice is a permutation of the node's points again
            for (int e = 0; e < nseg; ++e)
                for (unsigned i = lane; i < sz[e]; i += 32) perm[S.ps[cur][e] + i] = ids[(unsigned)e * Pv + i];
            rc = 2;
            break;
        }
        __syncwarp();
        // ---- children that are Tips: final
        if (anytip) {
            for (int c = 0; c < 2 * nseg; ++c) {
                const unsigned ls = S.lsz[c];
                for (unsigned i = lane; i < ls; i += 32) perm[S.ps[nxt][c] + i] = idn[(unsigned)c * (Pv / 2) + i];
            }
        }
        __syncwarp();
        cur = nxt;
        rc = anyint ? 1 : 0;
    }
    if (rc == 1) {                               // deeper than five levels (the host never routes such a subtree here): hand back what is left
        for (int e = 0; e < 32; ++e)
            for (unsigned i = lane; i < S.sz[cur][e]; i += 32) perm[S.ps[cur][e] + i] = S.ids[cur][(unsigned)e * 32u + i];
    }
    if (rc != 0 && lane == 0) *flag = 1;         // tie / tiny segment (rc 2)
}

// =====================================================================================================
// host orchestration
// =====================================================================================================
static unsigned next_pow2_host(unsigned v) { unsigned p = 1; while (p < v) p <<= 1; return p; }

template <int CAP, int NT>
static int launch_bottom_generic(rpf_handle* h, const BottomArgs& B, int nnodes_s, int tg) {
    dim3 grid((unsigned)nnodes_s, (unsigned)tg);
    const size_t smem = (size_t)CAP * 16;
    auto kfn = k_bottom<CAP, NT>;
    RPF_CUDA(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RPF_LAUNCH(h, PH_BOTTOM, kfn, grid, NT, smem, B);
    return RPF_OK;
}
template <int NT, int TAB>
static int launch_bottom_fast(rpf_handle* h, const BottomArgs& B, int nnodes_s, int tg) {
    dim3 grid((unsigned)nnodes_s, (unsigned)tg);
    const size_t smem = (size_t)NT * 8 * 12;
    typedef typename std::conditional<(NT <= 256), uint32_t, ull>::type W;      // <= 2048 slots: 32-bit sort words
    auto kfn = h->bottom_words64 ? k_bottom3<NT, TAB, ull> : k_bottom3<NT, TAB, W>;
    RPF_CUDA(h, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RPF_CUDA(h, cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    RPF_LAUNCH(h, PH_BOTTOM, kfn, grid, NT, smem, B);
    return RPF_OK;
}

// Bottom phase over `nroots` consecutive nodes (BFS ids B.first_gid ..) of `tg` trees; every root holds at most
// max_root points and its subtree has `levels` splitting levels.  fast: uniform-layout kernel -- a segment at relative
// depth j owns slots >> j slots and must keep >= 2 of them while it splits, so slots >= 2^levels and levels <= 10;
// otherwise the generic entry-table kernel (needs B.range / B.lvl_pv).
int rpf_bottom_launch(rpf_handle* h, const BottomArgs& B, int nroots, int tg, bool fast, unsigned max_root, int levels) {
    if (nroots <= 0 || tg <= 0) return RPF_OK;
    if (max_root > 8192) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "bottom phase: node larger than 8192 points");
    if (fast) {
        unsigned slots = std::max(256u, next_pow2_host(std::max(max_root, 1u)));
        while (levels > 0 && (slots >> levels) == 0) slots <<= 1;          // slots >= 2^levels
        if (slots > 8192 || levels > rpf_bottom_fast_levels()) return rpf_fail(h, RPF_ERR_ARG, "internal: subtree too deep for the fast bottom kernel");
        const bool deep = levels > 9, shallow = levels <= 5;
        BottomArgs B2 = B;
        if (h->bottom_select && slots == 1024 && levels >= 1 && levels <= 5 && !B.given_order && B.kmin && B.kmax && !B.only) {
            // warp-per-node select / partition kernel first; k_bottom3 below then only runs the nodes it flagged
            uint8_t* redo = (uint8_t*)h->ws_get(WS_BOT_REDO, (size_t)std::max(h->T, B.gt0 + tg) * nroots);
            if (!redo) return rpf_fail(h, RPF_ERR_NOMEM, h->err);
            redo += (size_t)B.gt0 * nroots;                                 // concurrent branches work on disjoint trees
            RPF_CUDA(h, cudaMemsetAsync(redo, 0, (size_t)tg * nroots, h->stream));
            const size_t smem4 = sizeof(B4Warp) * B4_WPC;
            RPF_CUDA(h, cudaFuncSetAttribute(k_bottom4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
            RPF_LAUNCH(h, PH_BOTTOM, k_bottom4, dim3((unsigned)((nroots + B4_WPC - 1) / B4_WPC), (unsigned)tg), B4_WPC * 32, smem4, B, redo, nroots);
            B2.only = redo;
        }
        const BottomArgs& B = B2;
#define RPF_BOT_CASE(SL, NT_)                                                                                         \
        case SL: return deep ? launch_bottom_fast<NT_, BOT2_TAB_DEEP>(h, B, nroots, tg)                              \
                             : (shallow ? launch_bottom_fast<NT_, BOT2_TAB_SHALLOW>(h, B, nroots, tg) : launch_bottom_fast<NT_, BOT2_TAB>(h, B, nroots, tg));
        switch (slots) {
            RPF_BOT_CASE(256, 32)
            RPF_BOT_CASE(512, 64)
            RPF_BOT_CASE(1024, 128)
            RPF_BOT_CASE(2048, 256)
            RPF_BOT_CASE(4096, 512)
            RPF_BOT_CASE(8192, 1024)
            default: return rpf_fail(h, RPF_ERR_ARG, "internal: bad bottom slot count");
        }
#undef RPF_BOT_CASE
    }
    if (max_root <= 256) return launch_bottom_generic<256, 128>(h, B, nroots, tg);
    if (max_root <= 1024) return launch_bottom_generic<1024, 256>(h, B, nroots, tg);
    if (max_root <= 4096) return launch_bottom_generic<4096, 512>(h, B, nroots, tg);
    return launch_bottom_generic<8192, 1024>(h, B, nroots, tg);
}
int rpf_bottom_fast_levels() { int v = BOT2_TAB_DEEP, l = 0; while (v > 1) { v >>= 1; ++l; } return l; }

// ---- geometry of a job: phase split and top-phase histogram shapes (pure host arithmetic on the topology) ----------
void rpf_job_geometry(const Topology& tp, int cap_cfg, int Lk, JobGeom& G) {
    G = JobGeom();
    const int64_t nn = tp.nnodes();
    int CAP = cap_cfg;
    {   // Tips larger than the configured capacity still get the reference's internal order as long as they fit the
        // largest bottom-phase instance (8192 points): raise the capacity for this job.
        uint32_t maxleaf = 0;
        for (int64_t g = 0; g < nn; ++g) if (tp.child[g] < 0) maxleaf = std::max(maxleaf, tp.size[g]);
        if ((int64_t)maxleaf > CAP && maxleaf <= 8192) { CAP = 256; while ((uint32_t)CAP < maxleaf) CAP <<= 1; }
    }
    G.CAP = CAP;
    const int L = std::min(tp.L_eff, Lk);
    G.L = L;
    int s = tp.nlevels;
    for (int l = 0; l < tp.nlevels; ++l) if ((int64_t)tp.lvl_maxsize[l] <= CAP) { s = l; break; }
    G.s = s;
    G.s_top = std::min(s, L);                        // levels 0..s_top-1 are split by the top phase
    G.NTOP = tp.level_off[std::min(G.s_top + 1, tp.nlevels)];
    G.order_exact = true;
    if (s >= tp.nlevels) G.order_exact = false;     // leaves larger than the capacity: membership exact, order not
    for (int l = 0; l < s && l < tp.nlevels; ++l)
        for (int64_t g = tp.level_off[l]; g < tp.level_off[l + 1]; ++g) if (tp.child[g] < 0 && tp.size[g] > 1) G.order_exact = false;
    G.nb_level.assign(std::max(L, 1), 0);
    G.smem_level.assign(std::max(L, 1), 0);
    G.HSZ = 1;
    for (int l = 0; l < G.s_top; ++l) {
        const int nodes = (int)(tp.level_off[l + 1] - tp.level_off[l]);
        int nb;
        if (nodes <= 512) { nb = std::min(HBINS_MAXNB, HBINS / (int)next_pow2_host((unsigned)nodes)); G.smem_level[l] = 1; }   // >= 64 bins per node
        else { nb = 256; G.smem_level[l] = 0; }                                                         // global-atomic histogram
        G.nb_level[l] = nb;
        G.HSZ = std::max<int64_t>(G.HSZ, (int64_t)nodes * nb);
    }
    G.HSZ = (G.HSZ + 1) & ~(int64_t)1;          // 64-bit REDs on counter pairs (k_top_hist flush)
    G.MAXTD = L + 1;
    // fast bottom kernel: at most 10 splitting levels below s, and 2^levels slots fit one CTA (8192)
    G.bottom_levels = std::max(0, L - s);
    G.fast_bottom = G.bottom_levels <= rpf_bottom_fast_levels();
    if (G.fast_bottom && s < tp.nlevels) {
        unsigned slots = std::max(256u, next_pow2_host(std::max<uint32_t>(tp.lvl_maxsize[s], 1)));
        while (G.bottom_levels > 0 && (slots >> G.bottom_levels) == 0) slots <<= 1;
        if (slots > 8192) G.fast_bottom = false;
    }
}
size_t rpf_job_ws_per_tree(const JobGeom& G, int64_t n) {
    return (G.s_top > 0 ? (size_t)n * 12 : 0) + (size_t)G.HSZ * 4 + (size_t)G.NTOP * (sizeof(NodeSel) + 4 + (size_t)G.MAXTD * 8) + 4096;
}

// Host half of a job: geometry, per-level flags and the device tables (bottom-phase ranges, histogram shapes) appended
// to TB.  Pure function of the topology, so callers may cache the result (the streaming build does).
void rpf_plan_job(const Topology& tp, int cap_cfg, int Lk, bool force_generic, TableBuf& TB, JobPlan& P) {
    P = JobPlan();
    JobGeom& G = P.G;
    rpf_job_geometry(tp, cap_cfg, Lk, G);
    if (force_generic) G.fast_bottom = false;
    P.n = tp.n; P.nn = tp.nnodes(); P.nlevels = tp.nlevels;
    P.level_off = tp.level_off;
    P.nroots = (int)tp.level_off[1];
    P.lvl_all_internal.assign(std::max(tp.nlevels, 1), 1);
    for (int l = 0; l < tp.nlevels; ++l)
        for (int64_t g = tp.level_off[l]; g < tp.level_off[l + 1]; ++g) if (tp.child[g] < 0) { P.lvl_all_internal[l] = 0; break; }
    const int s = G.s;
    // ---- bottom-phase tables: per level-s node, the BFS id range of its descendants at each deeper level
    std::vector<int2> rg; std::vector<uint32_t> pv;
    if (s < tp.nlevels) {
        P.nnodes_s = (int)(tp.level_off[s + 1] - tp.level_off[s]);
        P.nlb = std::max(1, tp.nlevels - s);
        P.maxsize_s = tp.lvl_maxsize[s];
        pv.resize(tp.nlevels);
        for (int l = 0; l < tp.nlevels; ++l) pv[l] = next_pow2_host(std::max<uint32_t>(tp.lvl_maxsize[l], 1));
        if (!G.fast_bottom) {
            rg.assign((size_t)P.nnodes_s * P.nlb, make_int2(0, 0));
            for (int e = 0; e < P.nnodes_s; ++e) {
                int64_t lo = tp.level_off[s] + e, hi = lo + 1;
                for (int j = 0; j < P.nlb; ++j) {
                    int64_t fi = -1, li = -1;
                    for (int64_t g = lo; g < hi; ++g) if (tp.child[g] >= 0) { if (fi < 0) fi = g; li = g; }
                    if (fi < 0) break;
                    rg[(size_t)e * P.nlb + j] = make_int2((int)lo, (int)hi);
                    lo = tp.child[fi]; hi = (int64_t)tp.child[li] + 2;
                }
            }
        }
    }
    P.off_range = rg.empty() ? (size_t)-1 : TB.put(rg.data(), rg.size() * sizeof(int2));
    P.off_lvlpv = pv.empty() ? (size_t)-1 : TB.put(pv.data(), pv.size() * 4);
    P.off_nb = TB.put(G.nb_level.data(), G.nb_level.size() * sizeof(int));
}

// Device half: enqueues every kernel of the job on the engine's stream.  `tab` = device address of the table block
// the plan's offsets refer to.  No host synchronisation.
int rpf_launch_job(rpf_handle* h, BuildJob& J, const JobPlan& P, const char* tab) {
    const JobGeom& G = P.G;
    const int64_t n = J.n;
    const int tg = J.tg;
    const int L = G.L, s = G.s, s_top = G.s_top;
    const int64_t NTOP = G.NTOP, HSZ = G.HSZ;
    const int MAXTD = G.MAXTD;
    J.order_exact = G.order_exact;
    if (s_top > 0 && NTOP > 65535) return rpf_fail(h, RPF_ERR_UNSUPPORTED, "too many top-phase nodes for 16-bit labels");

    if (L == 0 || n == 0) {   // every root is a single Tip holding its points in input order
        if (n > 0) {
            const int64_t tot = n * tg;
            RPF_LAUNCH(h, PH_MISC, k_iota_perm, (unsigned)((tot + 255) / 256), 256, 0, J.perm, n, J.ps, tg);
        }
        return RPF_OK;
    }
    const int2* range = P.off_range == (size_t)-1 ? nullptr : (const int2*)(tab + P.off_range);
    const uint32_t* lvlpv = P.off_lvlpv == (size_t)-1 ? nullptr : (const uint32_t*)(tab + P.off_lvlpv);
    const int* nbdev = (const int*)(tab + P.off_nb);

    if (s_top > 0) {
        // workspace: slice ws_part of ws_parts equal slices (one per concurrent branch), each sized for tgw trees
        const int tgw = J.ws_tg > 0 ? J.ws_tg : tg;
        bool ws_ok = true;
        auto WSP = [&](int slot, size_t bytes) -> char* {
            const size_t one = (bytes + 255) & ~(size_t)255;
            char* b = (char*)h->ws_get(slot, one * (size_t)J.ws_parts);
            if (!b) { ws_ok = false; return nullptr; }
            return b + one * (size_t)J.ws_part;
        };
        uint32_t max_lvl_nodes = 1;
        for (int l = 0; l < s_top; ++l) max_lvl_nodes = std::max<uint32_t>(max_lvl_nodes, (uint32_t)(P.level_off[l + 1] - P.level_off[l]));
        uint16_t* label = (uint16_t*)WSP(WS_LABEL, (size_t)tgw * n * 2);
        uint16_t* pbin = (uint16_t*)WSP(WS_PBIN, (size_t)tgw * n * 2);
        uint32_t* hist = (uint32_t*)WSP(WS_HIST, (size_t)tgw * HSZ * 4);
        NodeSel* sel = (NodeSel*)WSP(WS_SEL, (size_t)tgw * NTOP * sizeof(NodeSel));
        ull* cand = (ull*)WSP(WS_CAND, (size_t)tgw * n * 8);
        uint32_t* cand_total = (uint32_t*)WSP(WS_CANDTOT, (size_t)tgw * 12 + 8);  // [tg] candidate totals + [tg] margin-tracking flags + 2 work-list lengths + [tg] hist tickets
        uint32_t* wl = (uint32_t*)WSP(WS_WORKLIST, (size_t)tgw * max_lvl_nodes * 8);
        ull* pivots = (ull*)WSP(WS_PIVOTS, (size_t)tgw * NTOP * MAXTD * 8);
        uint32_t* fill = (uint32_t*)WSP(WS_FILL, (size_t)tgw * NTOP * 4);
        double* binlo = (double*)WSP(WS_BINLO, (size_t)tgw * J.Lk * 8);
        double* binscale = (double*)WSP(WS_BINSC, (size_t)tgw * J.Lk * 8);
        if (!ws_ok) return RPF_ERR_NOMEM;
        TopArgs A{};
        A.n = n; A.ks = J.ks; A.ps = J.ps; A.Tg = tg; A.L = J.Lk; A.NTOP = (int)NTOP; A.HSZ = (int)HSZ; A.MAXTD = MAXTD; A.gt0 = J.gt0; A.nn_all = J.ns;
        A.vec = ((n & 3) == 0 && (J.ks & 3) == 0 && ((uintptr_t)J.keys & 31) == 0) ? 1 : 0;
        A.keys = J.keys; A.label = label; A.pbin = pbin; A.child = J.d_child; A.nstart = J.d_start;
        A.nsize = J.d_size; A.binlo = binlo; A.binscale = binscale;
        A.kmin = J.kmin; A.kmax = J.kmax; A.hist = hist; A.sel = sel;
        A.cand = cand; A.cand_total = cand_total; A.track_any = cand_total + tg; A.pivots = pivots;
        A.wl_cnt = cand_total + 2 * tg; A.wl_big = wl; A.wl_tie = wl + (size_t)tg * max_lvl_nodes;
        uint32_t* done = cand_total + 2 * tg + 2;
        A.fused_finish = (h->fused_top & 2) ? 1 : 0;
        A.fill = fill; A.perm = J.perm; A.thr = J.thr; A.mlo = J.mlo; A.mhi = J.mhi;
        RPF_CUDA(h, cudaFuncSetAttribute(k_top_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, HBINS * 2));
        RPF_CUDA(h, cudaFuncSetAttribute(k_top_compact_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, CL_CAP * 6));
        RPF_CUDA(h, cudaFuncSetAttribute(k_top_scatter_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, SCAT_CH * 4));
        RPF_LAUNCH(h, PH_MISC, k_bin_setup, (unsigned)((tg * J.Lk + 127) / 128), 128, 0, A, nbdev, s_top);
        RPF_CUDA(h, cudaMemsetAsync(fill, 0, (size_t)tg * NTOP * 4, h->stream));
        if (P.nroots > 1) RPF_LAUNCH(h, PH_MISC, k_label_roots, (unsigned)((n + 255) / 256), 256, 0, label, n, tg, J.d_start, P.nroots);
        // Chunk per CTA, chosen per kernel class.  The streaming kernels do uniform work per CTA, so a launch takes
        // ceil(CTAs / resident CTAs) waves of (chunk + fixed per-CTA cost: histogram zero / flush, table loads); a 32-tree
        // group at 32 768 points per CTA runs 2.2 waves of the histogram kernel (the third wave is a quarter full), a
        // 4-tree shard of an 8-GPU run would not even fill the SMs once.  Candidates: multiples of 8192 up to TOP_CH.
        auto pick_chunk = [&](int resident_per_sm, int fixed, int forced) {
            if (forced >= 8192 && forced <= TOP_CH && forced % 8192 == 0) return forced;
            int best = TOP_CH; double best_cost = 1e300;
            for (int ch = TOP_CH; ch >= 8192; ch -= 8192) {
                const int64_t ctas = ((n + ch - 1) / ch) * tg;
                const int64_t waves = (ctas + 148 * resident_per_sm - 1) / (148 * resident_per_sm);
                const double cost = (double)waves * (ch + fixed);
                if (cost < best_cost * 0.98) { best_cost = cost; best = ch; }      // prefer the larger chunk on near ties
            }
            return best;
        };
        // (measured at 32 trees: the histogram kernel prefers the largest chunk -- its flush of 32 768 counters per CTA
        // outweighs the partial last wave: 0.99 ms at 32 768, 1.02 at 24 576, 1.09 at 16 384 -- so it only shrinks the
        // chunk to reach 2 CTAs per SM; compact / relabel gain 3-5 % from the wave model)
        int ch_hist = TOP_CH;
        while (ch_hist > 8192 && ((n + ch_hist - 1) / ch_hist) * tg < 2 * 148) ch_hist >>= 1;
        // ... and grows it (up to 57 344 points: 16-bit counters) when that saves a wave: the histogram kernels hold 2 CTAs per SM
        // and pay a fixed ~16 K points' worth per CTA (zero + flush of the 64 KB histogram), so 16 trees x 1M points run as ONE
        // wave of 288 CTAs instead of 496 CTAs in 1.7 waves
        {
            double best = 1e300; int best_ch = ch_hist;
            for (int ch = ch_hist; ch <= HIST_CH_MAX; ch += 8192) {
                const int64_t ctas = ((n + ch - 1) / ch) * tg;
                const int64_t waves = (ctas + 2 * 148 - 1) / (2 * 148);
                const double cost = (double)waves * (ch + 16384);
                if (cost < best * 0.98) { best = cost; best_ch = ch; }
            }
            if (h->hist_big_chunk) ch_hist = best_ch;
        }
        if (h->top_chunk[0] >= 8192 && h->top_chunk[0] <= HIST_CH_MAX && h->top_chunk[0] % 8192 == 0) ch_hist = h->top_chunk[0];
        const int ch_compact = pick_chunk(3, 2048, h->top_chunk[1]);
        const int ch_relabel = pick_chunk(2, 2048, h->top_chunk[2]);
        auto grid_for = [&](int ch) { return dim3((unsigned)((n + ch - 1) / ch), (unsigned)tg); };
        bool all_top_internal = true;
        auto done_for = [&](int nnodes_l) -> uint32_t* { return ((h->fused_top & 1) && nnodes_l >= 16 && tg >= h->fused_pick_min_tg) ? done : nullptr; };
        auto lean_level = [&](int l) -> bool {
            return h->lean_top && (n & 7) == 0 && (int)(P.level_off[l + 1] - P.level_off[l]) <= SMEM_NODES && P.lvl_all_internal[l];
        };
        RPF_CUDA(h, cudaFuncSetAttribute(k_top_relabel_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, HBINS * 2 + TOP_NT * 64));
        RPF_CUDA(h, cudaFuncSetAttribute(k_top_relabel_hist, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        bool hist_prefused = false;        // this level's histogram (and pick) came out of the previous level's k_top_relabel_hist
        for (int l = 0; l < s_top; ++l) {
            A.l = l; A.node0 = (int)P.level_off[l]; A.nnodes = (int)(P.level_off[l + 1] - P.level_off[l]);
            A.haslab = (l > 0 || P.nroots > 1) ? 1 : 0;
            A.NB = G.nb_level[l]; A.smem_hist = G.smem_level[l];
            A.child0 = (int)P.level_off[l + 1];
            A.all_internal = P.lvl_all_internal[l];
            all_top_internal = all_top_internal && A.all_internal;
            A.scatter_fast = (all_top_internal && 2 * A.nnodes <= SCAT_MAX) ? 1 : 0;
            if (!hist_prefused) {
                RPF_CUDA(h, cudaMemsetAsync(hist, 0, (size_t)tg * HSZ * 4, h->stream));
                RPF_CUDA(h, cudaMemsetAsync(cand_total, 0, (size_t)tg * 12 + 8, h->stream));
            }
            dim3 gn((unsigned)A.nnodes, (unsigned)tg);
            const size_t hs = A.smem_hist ? ((size_t)A.nnodes * A.NB + 1) / 2 * 4 : 0;     // 16-bit counters
            A.ch = ch_hist;
            // pick fused into the histogram kernel where ONE CTA settles a tree's level quickly (>= 16 nodes: one warp per
            // node); the first levels have a handful of nodes with up to 16384 bins each and keep their own pick launch
            // (measured, profiles/r02_top_fusion_sweep.txt: the ticket's __threadfence also waits for the CTA's streaming bin
            //  stores; with >= 16 trees per job other CTAs hide that and the fusion gains 0.15 ms at 32 trees, with 4 trees it
            //  costs 0.13 ms -- so it is only used for large jobs)
            A.done = done_for(A.nnodes);
            if (!hist_prefused) RPF_LAUNCH(h, PH_TOP_HIST, k_top_hist, grid_for(ch_hist), TOP_NT, hs, A);
            hist_prefused = false;
            if (!A.done) {
                if (A.NB <= 256) RPF_LAUNCH(h, PH_TOP_PICK, k_top_pick_warp, (unsigned)(((int64_t)A.nnodes * tg + 7) / 8), 256, 0, A);
                else RPF_LAUNCH(h, PH_TOP_PICK, k_top_pick, gn, 256, 0, A);
            }
            // lean kernels (16-byte rows of bins / labels, shared-memory node tables): see k_top_relabel_lean
            const bool lean = h->lean_top && (n & 7) == 0 && A.nnodes <= SMEM_NODES && A.all_internal;
            const bool last = l == s_top - 1;
            A.ch = ch_compact;
            if (lean) RPF_LAUNCH(h, PH_TOP_COMPACT, k_top_compact_lean, grid_for(ch_compact), TOP_NT, CL_CAP * 6, A);
            else RPF_LAUNCH(h, PH_TOP_COMPACT, k_top_compact, grid_for(ch_compact), TOP_NT, 0, A);
            if (A.fused_finish) {
                RPF_LAUNCH(h, PH_TOP_FINISH, k_top_finish_all, (unsigned)(((int64_t)A.nnodes * tg + FA_WARPS - 1) / FA_WARPS), FA_WARPS * 32, 0, A);
            } else {
                RPF_LAUNCH(h, PH_TOP_FINISH, k_top_finish_warp, (unsigned)(((int64_t)A.nnodes * tg + FW_WARPS - 1) / FW_WARPS), FW_WARPS * 32, 0, A);
                const unsigned gw = (unsigned)std::min<int64_t>((int64_t)A.nnodes * tg, 592);      // work-list walkers
                RPF_LAUNCH(h, PH_TOP_FINISH, k_top_finish, gw, 512, 0, A);
                RPF_LAUNCH(h, PH_TOP_TIES, k_top_ties, gw, 512, 0, A);
            }
            A.ch = ch_relabel;
            // relabel fused with the next level's histogram when that level is lean too and its histogram fits shared memory
            if (lean && !last && h->fuse_relabel_hist && A.vec && lean_level(l + 1) && G.smem_level[l + 1]) {
                A.nnodes2 = (int)(P.level_off[l + 2] - P.level_off[l + 1]); A.NB2 = G.nb_level[l + 1]; A.done2 = done_for(A.nnodes2);
                RPF_CUDA(h, cudaMemsetAsync(hist, 0, (size_t)tg * HSZ * 4, h->stream));
                RPF_CUDA(h, cudaMemsetAsync(cand_total, 0, (size_t)tg * 12 + 8, h->stream));
                A.ch = ch_hist;
                RPF_LAUNCH(h, PH_TOP_RELABEL, k_top_relabel_hist, grid_for(ch_hist), TOP_NT, (size_t)HBINS * 2 + TOP_NT * 64, A);
                hist_prefused = true;
            } else
            if (lean && !last) RPF_LAUNCH(h, PH_TOP_RELABEL, k_top_relabel_lean, grid_for(ch_relabel), TOP_NT, 0, A);
            else if (lean && A.scatter_fast)
                RPF_LAUNCH(h, PH_TOP_RELABEL, k_top_scatter_lean, dim3((unsigned)((n + SCAT_CH - 1) / SCAT_CH), (unsigned)tg), SCAT_NT, (size_t)SCAT_CH * 4, A);
            else RPF_LAUNCH(h, PH_TOP_RELABEL, k_top_relabel, grid_for(ch_relabel), TOP_NT, 0, A, (int)last);
        }
        // thr / margins of every top-phase node in one launch (NodeSel entries are per node and stay valid)
        A.node0 = 0; A.nnodes = (int)P.level_off[s_top];
        RPF_LAUNCH(h, PH_MISC, k_top_finalize, (unsigned)(((int64_t)A.nnodes * tg + 127) / 128), 128, 0, A);
    } else {
        const int64_t tot = n * tg;
        RPF_LAUNCH(h, PH_MISC, k_iota_perm, (unsigned)((tot + 255) / 256), 256, 0, J.perm, n, J.ps, tg);
    }

    if (s < P.nlevels) {
        BottomArgs B{};
        B.ks = J.ks; B.ps = J.ps; B.nn_all = J.ns; B.L = J.Lk; B.s = s; B.nlb = P.nlb; B.gt0 = J.gt0; B.first_gid = (int)P.level_off[s];
        B.given_order = 0;
        B.keys = J.keys; B.perm = J.perm; B.child = J.d_child; B.nstart = J.d_start; B.nsize = J.d_size;
        B.range = range; B.lvl_pv = lvlpv; B.thr = J.thr; B.mlo = J.mlo; B.mhi = J.mhi;
        B.kmin = J.kmin; B.kmax = J.kmax;
        // With an export sink the trees go through the bottom phase in groups: a group's slice of perm is final when its
        // launch ends, so its D2H runs on the download stream while the next group is still sorting.
        const bool to_sink = J.stream_to_sink && h->sink_perm && J.ps == n;
        const int groups = to_sink ? std::max(1, std::min(J.sink_groups, tg)) : 1;
        for (int gi = 0; gi < groups; ++gi) {
            const int ta = (int)((int64_t)tg * gi / groups), tb = (int)((int64_t)tg * (gi + 1) / groups);
            if (tb <= ta) continue;
            BottomArgs Bg = B;
            Bg.keys += (int64_t)ta * J.Lk * J.ks; Bg.perm += (int64_t)ta * J.ps; Bg.gt0 += ta;
            if (Bg.kmin) { Bg.kmin += (int64_t)ta * J.Lk; Bg.kmax += (int64_t)ta * J.Lk; }
            int rc = rpf_bottom_launch(h, Bg, P.nnodes_s, tb - ta, G.fast_bottom, P.maxsize_s, G.bottom_levels);
            if (rc) return rc;
            if (to_sink) {
                RPF_CUDA(h, cudaEventRecord(h->sink_ev[J.sink_ev0 + gi], h->stream));
                RPF_CUDA(h, cudaStreamWaitEvent(h->d2h_stream, h->sink_ev[J.sink_ev0 + gi], 0));
                RPF_CUDA(h, cudaMemcpyAsync(h->sink_perm + (int64_t)(J.gt0 + ta) * n, J.perm + (int64_t)ta * J.ps, (size_t)(tb - ta) * n * 4,
                                            cudaMemcpyDeviceToHost, h->d2h_stream));
                // the group's node arrays are final too (top-phase nodes: k_top_finalize above; the rest: this bottom launch), so
                // they leave with the group instead of as a 25 MB tail after the last one
                const size_t o = (size_t)(J.gt0 + ta) * (size_t)J.ns, nbn = (size_t)(tb - ta) * (size_t)J.ns * 8;
                if (h->sink_thr) RPF_CUDA(h, cudaMemcpyAsync(h->sink_thr + o, J.thr + o, nbn, cudaMemcpyDeviceToHost, h->d2h_stream));
                if (h->sink_mlo) RPF_CUDA(h, cudaMemcpyAsync(h->sink_mlo + o, J.mlo + o, nbn, cudaMemcpyDeviceToHost, h->d2h_stream));
                if (h->sink_mhi) RPF_CUDA(h, cudaMemcpyAsync(h->sink_mhi + o, J.mhi + o, nbn, cudaMemcpyDeviceToHost, h->d2h_stream));
            }
        }
        if (to_sink) { h->sink_perm_streamed = true; h->sink_nodes_streamed = true; }
    }
    return RPF_OK;
}

// plan + stage + launch (the batch build: the plan is cheap and the tables travel through the staging ring)
int rpf_run_job(rpf_handle* h, BuildJob& J) {
    TableBuf TB; JobPlan P;
    rpf_plan_job(*J.tp, h->bottom_cap, J.Lk, h->force_generic_bottom, TB, P);
    int rc = h->stage_begin(TB.bytes.size() + 1024);
    if (rc) return rc;
    const char* tab = (const char*)h->stage_put_raw(TB.bytes.data(), TB.bytes.size());
    if (!tab) return rpf_fail(h, RPF_ERR_NOMEM, h->err);
    rc = h->stage_flush();
    if (rc) return rc;
    return rpf_launch_job(h, J, P, tab);
}

#define WS(h, var, type, slot, bytes)                                   \
    type* var = (type*)(h)->ws_get((slot), (bytes));                    \
    if (!var) { (h)->tg_cached = 0; return RPF_ERR_NOMEM; }

// The plan of a batch build (geometry + device-resident tables) only depends on the shape: kept per handle.
namespace {
struct BatchPlan {
    int64_t n = -1; int maxDepth = -1, minLeaf = -1, cap = -1, Lk = -1; bool force_generic = false;
    JobPlan P; char* d_tab = nullptr;
    ~BatchPlan() { if (d_tab) cudaFree(d_tab); }
};
void free_batch_plan(void* p) { delete (BatchPlan*)p; }
}  // namespace

static int get_batch_plan(rpf_handle* h, int L, BatchPlan** out) {
    const Topology& tp = h->topo;
    BatchPlan* B = (BatchPlan*)h->batch_plan;
    if (B && B->n == tp.n && B->maxDepth == tp.maxDepth && B->minLeaf == tp.minLeaf && B->cap == h->bottom_cap && B->Lk == L &&
        B->force_generic == h->force_generic_bottom && B->P.nn == tp.nnodes()) { *out = B; return RPF_OK; }
    if (B) { cudaStreamSynchronize(h->stream); delete B; h->batch_plan = nullptr; }
    ++h->cfg_epoch;
    B = new BatchPlan();
    B->n = tp.n; B->maxDepth = tp.maxDepth; B->minLeaf = tp.minLeaf; B->cap = h->bottom_cap; B->Lk = L; B->force_generic = h->force_generic_bottom;
    TableBuf TB;
    rpf_plan_job(tp, h->bottom_cap, L, h->force_generic_bottom, TB, B->P);
    if (cudaMalloc(&B->d_tab, std::max<size_t>(TB.bytes.size(), 256)) != cudaSuccess ||
        cudaMemcpy(B->d_tab, TB.bytes.data(), TB.bytes.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError(); delete B;
        return rpf_fail(h, RPF_ERR_NOMEM, "batch plan: device tables");
    }
    h->batch_plan = B; h->batch_plan_free = free_batch_plan;
    *out = B;
    return RPF_OK;
}

// Rows [r0, r0 + nr) of the host matrix -> dX on `stream`.  Single GPU: one H2D copy.  Rank R of a W-rank tree-sharded
// forest (data replicated): this rank copies only ITS 1/W sub-block over its own PCIe link and the sub-blocks are
// all-gathered in place over NCCL / NVLink (the last sub-block may spill past row r0 + nr: dX carries x_pad_rows spare rows,
// and a later block overwrites its own range afterwards on the same stream).
int rpf_upload_rows(rpf_handle* h, const double* hostX, int64_t r0, int64_t nr, cudaStream_t stream, cudaStream_t gather_stream, cudaEvent_t up_ev) {
    const int W = rpf_comm_world(h), R = rpf_comm_rank(h);
    const size_t rb = (size_t)h->d * 8;
    if (W <= 1) {
        RPF_CUDA(h, cudaMemcpyAsync((void*)(h->dX + r0 * h->d), hostX + r0 * h->d, (size_t)nr * rb, cudaMemcpyHostToDevice, stream));
        return RPF_OK;
    }
    const int64_t sub = (nr + W - 1) / W;
    if (h->x_pad_rows < W) return rpf_fail(h, RPF_ERR_STATE, "upload_rows: point buffer has no all-gather padding");
    const int64_t a = std::min(nr, R * sub), b = std::min(nr, a + sub);
    if (b > a) RPF_CUDA(h, cudaMemcpyAsync((void*)(h->dX + (r0 + a) * h->d), hostX + (r0 + a) * h->d, (size_t)(b - a) * rb, cudaMemcpyHostToDevice, stream));
    if (gather_stream && gather_stream != stream) {      // PCIe copies and NVLink all-gathers on separate streams: block b + 1
        RPF_CUDA(h, cudaEventRecord(up_ev, stream));     // crosses PCIe while block b is being gathered
        RPF_CUDA(h, cudaStreamWaitEvent(gather_stream, up_ev, 0));
        stream = gather_stream;
    }
    return rpf_comm_allgather(h, (void*)(h->dX + r0 * h->d), (size_t)sub * rb, stream);
}

// forestBatch: the whole data set is one chunk (Batch.hs:48-63).  hostX != NULL: the points still live in host memory
// (h->dX is allocated but empty): they are uploaded in row blocks on a second stream while the projection kernel
// already runs on the blocks that have arrived (the rest of the build needs all keys and follows on the engine's stream).
// A rebuild of an unchanged shape with the points resident replays the whole launch sequence as one CUDA graph.
int rpf_build_impl(rpf_handle* h, const double* hostX) {
    const Topology& tp = h->topo;
    const int64_t n = h->n, nn = tp.nnodes();
    const int T = h->T, L = tp.L_eff;

    // ---- result arrays (kept across builds of the same shape)
    const uint64_t ep_before = h->cfg_epoch;
    int rc = rpf_alloc_forest(h, nn, n);
    if (rc) return rc;

    JobGeom G;
    rpf_job_geometry(tp, h->bottom_cap, L, G);
    h->leaf_order_exact = G.order_exact;

    // ---- tree group size from the memory budget (free memory + what the workspace already holds)
    int Tg;
    if (h->tg_cached > 0 && h->tg_key_n == n && h->tg_key_L == L && h->tg_key_T == T) {
        Tg = h->tg_cached;
    } else {
        size_t freeB = 0, totalB = 0;
        RPF_CUDA(h, cudaMemGetInfo(&freeB, &totalB));
        const size_t per_tree = (size_t)L * n * 8 + rpf_job_ws_per_tree(G, n);
        const size_t budget = (size_t)((double)(freeB + h->ws_bytes) * 0.7);
        if (per_tree > budget) return rpf_fail(h, RPF_ERR_NOMEM, "not enough device memory for one tree's keys");
        Tg = (int)std::min<size_t>((size_t)T, std::max<size_t>(1, budget / per_tree));
        h->tg_key_n = n; h->tg_key_L = L; h->tg_key_T = T; h->tg_cached = Tg;
    }

    WS(h, keys, ull, WS_KEYS, (size_t)Tg * std::max(L, 1) * std::max<int64_t>(n, 1) * 8);
    WS(h, kmin, ull, WS_KMIN, (size_t)Tg * std::max(L, 1) * 8);
    WS(h, kmax, ull, WS_KMAX, (size_t)Tg * std::max(L, 1) * 8);
    BatchPlan* BP = nullptr;
    rc = get_batch_plan(h, L, &BP);
    if (rc) return rc;

    const bool pipelined = hostX && Tg == T && L > 0 && n > 0;
    // export sink: only the host-data build streams into it (never captured into the build graph)
    const bool sink = hostX && h->sink_perm != nullptr;
    h->sink_pending = false; h->sink_perm_streamed = false; h->sink_nodes_streamed = false;
    if (h->d2h_stream) RPF_CUDA(h, cudaStreamSynchronize(h->d2h_stream));   // a previous build's download is complete before its source is rewritten
    if (sink) {
        if (!h->d2h_stream) RPF_CUDA(h, cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
        if (!h->sink_ev[0]) for (auto& e : h->sink_ev) RPF_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const bool graphable = h->use_graphs && !hostX && Tg == T && L > 0 && n > 0 && !h->profiling;
    if (graphable && h->build_graph && h->graph_epoch == h->cfg_epoch) {
        RPF_CUDA(h, cudaGraphLaunch(h->build_graph, h->stream));
        h->launches += h->graph_launches;
        h->phase_launches[PH_MISC] += h->graph_launches;
        h->last_build_epoch = h->cfg_epoch;
        return RPF_OK;
    }
    if (h->build_graph) { cudaGraphExecDestroy(h->build_graph); h->build_graph = nullptr; }

    // concurrent branches per tree group (see rpf_handle::branch_stream): off while profiling (per-kernel event pairs want
    // one stream) and for tiny inputs
    int NB = 1;
    if (!h->profiling && h->branches != 1 && Tg >= 2 && n >= 65536 && L > 0)
        NB = h->branches >= 2 ? std::min(h->branches, RPF_MAX_BRANCH) : (Tg >= 16 ? 2 : std::min(Tg, RPF_MAX_BRANCH));
    if (NB > 1 && !h->branch_stream[0]) {
        for (int b = 0; b < RPF_MAX_BRANCH; ++b) RPF_CUDA(h, cudaStreamCreateWithFlags(&h->branch_stream[b], cudaStreamNonBlocking));
        for (int b = 0; b <= RPF_MAX_BRANCH; ++b) RPF_CUDA(h, cudaEventCreateWithFlags(&h->branch_ev[b], cudaEventDisableTiming));
    }
    // everything below only enqueues work on the engine's stream (no allocation once the workspace has its size)
    auto enqueue = [&]() -> int {
        RPF_CUDA(h, cudaMemsetAsync(h->d_thr, 0, h->res_node_bytes, h->stream));
        RPF_CUDA(h, cudaMemsetAsync(h->d_mlo, 0, h->res_node_bytes, h->stream));
        RPF_CUDA(h, cudaMemsetAsync(h->d_mhi, 0, h->res_node_bytes, h->stream));
        if (hostX && !pipelined && n > 0) {    // several tree groups (or nothing to project): plain upload first
            int rcu = rpf_upload_rows(h, hostX, 0, n, h->stream);
            if (rcu) return rcu;
        }
        for (int t0 = 0; t0 < T; t0 += Tg) {
            const int tg = std::min(Tg, T - t0);
            if (L > 0 && n > 0) {   // K1
                RPF_CUDA(h, cudaMemsetAsync(kmin, 0xff, (size_t)tg * L * 8, h->stream));
                RPF_CUDA(h, cudaMemsetAsync(kmax, 0x00, (size_t)tg * L * 8, h->stream));
                if (pipelined) {
                    if (!h->copy_stream) RPF_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
                    const int NBLK = 16;
                    int64_t rows = (n + NBLK - 1) / NBLK;
                    rows = (rows + 4095) / 4096 * 4096;              // whole projection tiles, 32-byte aligned key columns
                    // the upload may only start once the engine's stream is done with the previous contents of dX
                    if (!h->copy_ev[0]) for (int i = 0; i <= NBLK; ++i) RPF_CUDA(h, cudaEventCreateWithFlags(&h->copy_ev[i], cudaEventDisableTiming));
                    RPF_CUDA(h, cudaEventRecord(h->copy_ev[NBLK], h->stream));
                    RPF_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->copy_ev[NBLK], 0));
                    const bool multi = rpf_comm_world(h) > 1;
                    if (multi && !h->gather_stream) {
                        RPF_CUDA(h, cudaStreamCreateWithFlags(&h->gather_stream, cudaStreamNonBlocking));
                        for (auto& e : h->up_ev) RPF_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                    }
                    cudaStream_t gs = multi ? h->gather_stream : h->copy_stream;
                    int bi = 0;
                    for (int64_t r0 = 0; r0 < n; r0 += rows, ++bi) {
                        const int64_t nr = std::min(rows, n - r0);
                        // rank of a sharded forest: 1/W of the block over PCIe (copy stream), NVLink all-gather (gather stream)
                        int rcu = rpf_upload_rows(h, hostX, r0, nr, h->copy_stream, gs, multi ? h->up_ev[bi] : nullptr);
                        if (rcu) return rcu;
                        RPF_CUDA(h, cudaEventRecord(h->copy_ev[bi], gs));
                        RPF_CUDA(h, cudaStreamWaitEvent(h->stream, h->copy_ev[bi], 0));
                        int rc2 = rpf_project_launch(h, PH_PROJECT, h->dX + r0 * h->d, nr, t0, tg, L, true, keys + r0, n, kmin, kmax);
                        if (rc2) return rc2;
                    }
                }
            }
            // the trees [ta, tb) of the group as one job
            auto job = [&](int ta, int tb, int part, int parts, int ws_tg) -> int {
                BuildJob J{};
                J.tp = &tp; J.d_start = h->d_node_start; J.d_size = h->d_node_size; J.d_child = h->d_node_child;
                J.n = n; J.ks = n; J.ps = n; J.ns = nn; J.Lk = L;
                J.keys = keys + (int64_t)ta * L * n; J.kmin = kmin + (int64_t)ta * L; J.kmax = kmax + (int64_t)ta * L;
                J.perm = h->d_perm + (int64_t)(t0 + ta) * n; J.thr = h->d_thr; J.mlo = h->d_mlo; J.mhi = h->d_mhi;
                J.gt0 = t0 + ta; J.tg = tb - ta;
                J.stream_to_sink = sink;
                J.ws_part = part; J.ws_parts = parts; J.ws_tg = ws_tg;
                J.sink_groups = std::max(1, 8 / parts); J.sink_ev0 = part * J.sink_groups;
                return rpf_launch_job(h, J, BP->P, BP->d_tab);
            };
            // (tried: every branch projecting its own trees on its own stream, one branch after the other, so that the LSU-bound
            //  projection of branch b + 1 runs beside the DRAM / issue-bound top phase of branch b: 6.5 -> 6.6 ms with 2 branches,
            //  6.9 ms with 4 -- X is read once per branch and the kernels do not share an SM well)
            if (L > 0 && n > 0 && !pipelined) {       // one pass over X for all trees of the group
                int rc2 = rpf_project_launch(h, PH_PROJECT, h->dX, n, t0, tg, L, true, keys, n, kmin, kmax);
                if (rc2) return rc2;
            }
            if (NB <= 1 || tg < 2) {
                int rc2 = job(0, tg, 0, 1, 0);
                if (rc2) return rc2;
            } else {
                // concurrent branches: contiguous parts of the group, each with its top and bottom phase on its own stream;
                // forked from / joined on the engine's stream (captured into the build graph like everything else)
                const int nb = std::min(NB, tg), per = (tg + nb - 1) / nb;
                cudaStream_t main = h->stream;
                RPF_CUDA(h, cudaEventRecord(h->branch_ev[RPF_MAX_BRANCH], main));
                int rcb = RPF_OK, used = 0;
                for (int b = 0; b < nb && rcb == RPF_OK; ++b) {
                    const int ta = b * per, tb = std::min(tg, ta + per);
                    if (tb <= ta) break;
                    ++used;
                    h->stream = h->branch_stream[b];
                    cudaError_t e = cudaStreamWaitEvent(h->stream, h->branch_ev[RPF_MAX_BRANCH], 0);
                    if (e != cudaSuccess) { rcb = rpf_fail(h, RPF_ERR_CUDA, std::string("branch fork: ") + cudaGetErrorString(e)); break; }
                    rcb = job(ta, tb, b, nb, per);
                    if (rcb == RPF_OK && cudaEventRecord(h->branch_ev[b], h->stream) != cudaSuccess) rcb = rpf_fail(h, RPF_ERR_CUDA, "branch join: event record");
                }
                h->stream = main;
                for (int b = 0; b < used; ++b)       // always rejoin (a capture must end with every forked stream joined)
                    if (cudaStreamWaitEvent(main, h->branch_ev[b], 0) != cudaSuccess && rcb == RPF_OK) rcb = rpf_fail(h, RPF_ERR_CUDA, "branch join: wait");
                if (rcb) return rcb;
            }
        }
        if (sink) {     // node arrays (and perm, unless the job streamed it group by group) once everything is final
            RPF_CUDA(h, cudaEventRecord(h->sink_ev[8], h->stream));
            RPF_CUDA(h, cudaStreamWaitEvent(h->d2h_stream, h->sink_ev[8], 0));
            const size_t nb = (size_t)T * (size_t)nn * 8;
            if (!h->sink_nodes_streamed) {
                if (h->sink_thr && nb) RPF_CUDA(h, cudaMemcpyAsync(h->sink_thr, h->d_thr, nb, cudaMemcpyDeviceToHost, h->d2h_stream));
                if (h->sink_mlo && nb) RPF_CUDA(h, cudaMemcpyAsync(h->sink_mlo, h->d_mlo, nb, cudaMemcpyDeviceToHost, h->d2h_stream));
                if (h->sink_mhi && nb) RPF_CUDA(h, cudaMemcpyAsync(h->sink_mhi, h->d_mhi, nb, cudaMemcpyDeviceToHost, h->d2h_stream));
            }
            if (!h->sink_perm_streamed && n > 0)
                RPF_CUDA(h, cudaMemcpyAsync(h->sink_perm, h->d_perm, (size_t)T * n * 4, cudaMemcpyDeviceToHost, h->d2h_stream));
            RPF_CUDA(h, cudaEventRecord(h->sink_ev[9], h->d2h_stream));
            h->sink_pending = true;
        }
        return RPF_OK;
    };

    // second build of an unchanged configuration: capture the launch sequence, then run it as a graph from now on
    const bool capture = graphable && h->last_build_epoch == h->cfg_epoch && ep_before == h->cfg_epoch;
    if (capture) {
        const int64_t l0 = h->launches;
        int64_t pl[PH_COUNT];
        for (int i = 0; i < PH_COUNT; ++i) pl[i] = h->phase_launches[i];
        cudaGraph_t graph = nullptr;
        h->capturing = true;
        cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
        int rcq = e == cudaSuccess ? enqueue() : RPF_ERR_CUDA;
        cudaError_t e2 = e == cudaSuccess ? cudaStreamEndCapture(h->stream, &graph) : e;
        h->capturing = false;
        cudaGraphExec_t exec = nullptr;
        if (rcq == RPF_OK && e2 == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
            cudaGraphDestroy(graph);
            h->build_graph = exec; h->graph_epoch = h->cfg_epoch; h->graph_launches = h->launches - l0;
            RPF_CUDA(h, cudaGraphLaunch(h->build_graph, h->stream));
            h->last_build_epoch = h->cfg_epoch;
            return RPF_OK;
        }
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        h->launches = l0;
        for (int i = 0; i < PH_COUNT; ++i) h->phase_launches[i] = pl[i];
        h->use_graphs = false;                 // capture is not possible in this environment: plain launches from now on
    }
    rc = enqueue();
    if (rc) return rc;
    h->last_build_epoch = h->cfg_epoch;
    return RPF_OK;
}

// (re)allocate the forest arrays thr/mlo/mhi [T][nn] and perm [T][n]
int rpf_alloc_forest(rpf_handle* h, int64_t nn, int64_t n) {
    const int T = h->T;
    const size_t node_bytes = sizeof(double) * (size_t)std::max<int64_t>(T * nn, 1), perm_bytes = sizeof(uint32_t) * (size_t)std::max<int64_t>(T * n, 1);
    if (h->res_node_bytes != node_bytes || h->res_perm_bytes != perm_bytes || !h->d_thr) {
        if (h->d_thr) { cudaFree(h->d_thr); cudaFree(h->d_mlo); cudaFree(h->d_mhi); cudaFree(h->d_perm); h->d_thr = h->d_mlo = h->d_mhi = nullptr; h->d_perm = nullptr; }
        h->res_node_bytes = h->res_perm_bytes = 0;
        ++h->cfg_epoch;
        RPF_CUDA(h, cudaMalloc(&h->d_thr, node_bytes));
        RPF_CUDA(h, cudaMalloc(&h->d_mlo, node_bytes));
        RPF_CUDA(h, cudaMalloc(&h->d_mhi, node_bytes));
        RPF_CUDA(h, cudaMalloc(&h->d_perm, perm_bytes));
        h->res_node_bytes = node_bytes; h->res_perm_bytes = perm_bytes;
    }
    return RPF_OK;
}

// projections of a query batch onto every (tree, level) hyperplane: keysQ[(t*L + l) * nq + q]
int rpf_project_queries(rpf_handle* h, const double* dQ, int64_t nq, double* d_keysQ) {
    return rpf_project_launch(h, PH_Q_PROJECT, dQ, nq, 0, h->T, h->topo.L_eff, false, d_keysQ, nq, nullptr, nullptr);
}
