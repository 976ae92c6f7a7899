// k_scan.cu — exclusive prefix sum over u32 (cell -> first fragment / first span / first connection).
// Three launches: per-block reduce, single-block scan of the block sums, per-block scan + offset.
// HBM traffic: 2 reads + 1 write of the array (12 B per cell); the arrays are L2-resident at the
// configured sizes (8 M cells = 32 MB).
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "rcb_common.cuh"

thread_local RcbProfHook g_rcb_prof = {nullptr, nullptr, nullptr};

bool rcb_sync_check() {
    static const bool on = getenv("RCB_SYNC_CHECK") != nullptr;
    return on;
}

void rcb_sync_check_kernel(const char* name, cudaStream_t s) {
    cudaError_t e = cudaPeekAtLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        int dev = -1;
        cudaGetDevice(&dev);
        fprintf(stderr, "[rcb sync check] device %d stream %p: %s: %s\n", dev, (void*)s, name, cudaGetErrorString(e));
    }
}

void rcb_raise_dynamic_smem(const void* kernel, size_t bytes) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    // the current limit is read back from the driver (per kernel and current device) rather than remembered here: a process
    // that resets a device between two lives of the library starts from the default again
    cudaFuncAttributes a{};
    if (cudaFuncGetAttributes(&a, kernel) == cudaSuccess && (size_t)a.maxDynamicSharedSizeBytes >= bytes) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

namespace {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += n;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    __shared__ uint32_t block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        uint32_t winc = warp_inclusive_scan(w);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = winc - w;
        if (lane == 31) block_total = winc;
    }
    __syncthreads();
    uint32_t r = inc - v + warp_sums[warp];
    *total = block_total;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const uint32_t* __restrict__ in, size_t n,
                                                              uint32_t* __restrict__ block_sums) {
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocksums(uint32_t* __restrict__ block_sums, size_t nblocks,
                                                                 uint32_t* __restrict__ out_total) {
    uint32_t carry = 0;
    for (size_t base = 0; base < nblocks; base += SCAN_THREADS) {
        size_t idx = base + threadIdx.x;
        uint32_t v = idx < nblocks ? block_sums[idx] : 0;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (idx < nblocks) block_sums[idx] = ex + carry;
        carry += total;
    }
    if (threadIdx.x == 0) *out_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                             size_t n, const uint32_t* __restrict__ block_sums) {
    // thread owns SCAN_ITEMS consecutive items so the scan order equals the memory order; full tiles move them as two
    // 16-byte vectors (eight scalar accesses at a 32-byte lane stride were eight times the LSU transactions)
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    static_assert(SCAN_ITEMS == 8, "two uint4 per thread");
    const bool full = (size_t)(blockIdx.x + 1) * SCAN_TILE <= n && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
    if (full) {
        const uint4 a = *reinterpret_cast<const uint4*>(in + base), b = *reinterpret_cast<const uint4*>(in + base + 4);
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) s += v[i];
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            v[i] = (base + i < n) ? in[base + i] : 0;
            s += v[i];
        }
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(s, &total) + block_sums[blockIdx.x];
    if (full) {
        uint32_t o[SCAN_ITEMS];
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            o[i] = ex;
            ex += v[i];
        }
        *reinterpret_cast<uint4*>(out + base) = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4*>(out + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; ++i) {
            if (base + i < n) out[base + i] = ex;
            ex += v[i];
        }
    }
}

}  // namespace

size_t scan_tmp_elems(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

void launch_exclusive_scan(const uint32_t* in, uint32_t* out, size_t n, uint32_t* tmp, cudaStream_t s) {
    if (n == 0) {
        cudaMemsetAsync(out, 0, sizeof(uint32_t), s);
        return;
    }
    const size_t nblocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    {
        KScope _k("k_scan_reduce", s);
        k_scan_reduce<<<(unsigned)nblocks, SCAN_THREADS, 0, s>>>(in, n, tmp);
    }
    {
        KScope _k("k_scan_blocksums", s);
        k_scan_blocksums<<<1, SCAN_THREADS, 0, s>>>(tmp, nblocks, out + n);
    }
    {
        KScope _k("k_scan_apply", s);
        k_scan_apply<<<(unsigned)nblocks, SCAN_THREADS, 0, s>>>(in, out, n, tmp);
    }
}
