// rpf_device.cuh -- device helpers shared by build.cu and query.cu
#pragma once
#include <stdint.h>

typedef unsigned long long ull;

__device__ __forceinline__ int ilog2_pow2(unsigned v) { return 31 - __clz(v); }
__device__ __forceinline__ unsigned next_pow2_u32(unsigned v) { return v <= 1 ? 1u : 1u << (32 - __clz(v - 1)); }

// 8-bit MSD radix select of rank r (0-based) over `c` uint64 values fetched by get(i); every thread of the CTA
// must call it.  Returns the selected value; cl = #values < it, ce = #values == it.
// sh: >= 264 uint32 of shared memory, sh64: 1 uint64 of shared memory.
// PASSES < 8: only the top 8 * PASSES bits are resolved (keys whose low bits are zero, e.g. 32-bit values << 32).
template <int NT, int PASSES = 8, typename Get>
__device__ ull cta_radix_select(uint32_t c, uint32_t r, Get get, uint32_t* sh, ull* sh64, uint32_t& cl, uint32_t& ce) {
    ull prefix = 0;
    uint32_t rr = r, below = 0;
    for (int pass = 0; pass < PASSES; ++pass) {
        const int shift = 56 - 8 * pass;
        const ull mask_hi = pass == 0 ? 0ull : (~0ull << (shift + 8));
        for (int j = threadIdx.x; j < 256; j += NT) sh[j] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < c; i += NT) {
            const ull v = get(i);
            if ((v & mask_hi) == prefix) atomicAdd(&sh[(v >> shift) & 255], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 32) {          // warp 0: locate the digit whose cumulative count covers rr
            const unsigned lane = threadIdx.x;
            uint32_t loc[8], s = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { loc[b] = sh[lane * 8 + b]; s += loc[b]; }
            uint32_t incl = s;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= (unsigned)off) incl += y; }
            const uint32_t excl = incl - s;
            const bool mine = rr >= excl && rr < incl;
            if (mine) {
                uint32_t cum = excl; int dg = 7;
#pragma unroll
                for (int b = 0; b < 8; ++b) { if (rr < cum + loc[b]) { dg = b; break; } cum += loc[b]; }
                sh[256] = cum; sh[257] = loc[dg];
                *sh64 = prefix | ((ull)(lane * 8 + dg) << shift);
            }
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
