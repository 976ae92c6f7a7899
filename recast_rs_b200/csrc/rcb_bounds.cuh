// rcb_bounds.cuh — the -DRCB_BOUNDS debug build: every hand-laid-out array of the tile-resident kernels and of the clipper
// (shared-memory regions carved out of one dynamic allocation, per-band ranges of the batch-wide arenas) is accessed through a
// checked accessor that knows how many elements the array holds.  compute-sanitizer is closed on this GPU pool; this build is
// its stand-in: `RCB_DEBUG_SO=1 pytest -m gpu` runs the parity suite on librecast_b200_dbg.so and fails if any index of any
// launch left its array (rcb_debug_bounds_report).  In the normal build the accessors ARE raw pointers: zero cost.
#pragma once
#include <cuda_runtime.h>

#ifdef RCB_BOUNDS
// one counter block per translation unit (no relocatable device code in this build): [0] violations, [1] file << 32 | line
// of the first, [2] its index, [3] its array size.  RCB_BOUNDS_EXPORT(name) gives the host a reader for this unit's block.
static __device__ unsigned long long g_rcb_bounds[4];
#define RCB_BOUNDS_EXPORT(name)                                                                                  \
    void name(unsigned long long out[4], int reset) {                                                            \
        cudaMemcpyFromSymbol(out, g_rcb_bounds, 4 * sizeof(unsigned long long));                                 \
        if (reset) {                                                                                             \
            const unsigned long long z[4] = {0, 0, 0, 0};                                                        \
            cudaMemcpyToSymbol(g_rcb_bounds, z, sizeof z);                                                       \
        }                                                                                                        \
    }

__device__ __noinline__ static void rcb_bounds_hit(int where, unsigned long long i, unsigned long long n) {
    if (atomicAdd(&g_rcb_bounds[0], 1ull) == 0) {
        g_rcb_bounds[1] = ((unsigned long long)(where >> 16) << 32) | (unsigned long long)(where & 0xffff);
        g_rcb_bounds[2] = i;
        g_rcb_bounds[3] = n;
    }
}

template <class T>
struct BArr {
    T* p;
    unsigned long long n;
    int where;
    __device__ __forceinline__ T& operator[](unsigned long long i) const {
        if (i >= n) {
            rcb_bounds_hit(where, i, n);
            return p[0];
        }
        return p[i];
    }
    __device__ __forceinline__ BArr operator+(unsigned long long k) const {
        if (k > n) rcb_bounds_hit(where, k, n);
        return BArr{p + k, k <= n ? n - k : 0, where};
    }
    __device__ __forceinline__ T* raw() const { return p; }
};
#define RCB_A(T) BArr<T>
#define RCB_AR(T) BArr<T>
#define RCB_MK(T, ptr, n) (BArr<T>{(T*)(ptr), (unsigned long long)(n), (RCB_FILE_ID << 16) | __LINE__})
#define RCB_RAW(a) ((a).raw())
#define RCB_CHECK(cond)                                                              \
    do {                                                                             \
        if (!(cond)) rcb_bounds_hit((RCB_FILE_ID << 16) | __LINE__, 0ull, 0ull);    \
    } while (0)
__device__ __forceinline__ unsigned rcb_ix(unsigned long long i, unsigned long long n, int where) {
    if (i >= n) {
        rcb_bounds_hit(where, i, n);
        return 0u;
    }
    return (unsigned)i;
}
#define RCB_IX(i, n) rcb_ix((unsigned long long)(i), (unsigned long long)(n), (RCB_FILE_ID << 16) | __LINE__)  // index i of an array of n
#else
#define RCB_BOUNDS_EXPORT(name)                               \
    void name(unsigned long long out[4], int) { out[0] = out[1] = out[2] = out[3] = 0; }
#define RCB_A(T) T*
#define RCB_AR(T) T* __restrict__
#define RCB_MK(T, ptr, n) ((T*)(ptr))
#define RCB_RAW(a) (a)
#define RCB_CHECK(cond) ((void)0)
#define RCB_IX(i, n) (i)
#endif
