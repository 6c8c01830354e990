// scan.cuh -- exclusive prefix sum u32 -> u64 over up to 2^31 items, shared by the JPEG encoder (bit offsets of blocks, byte offsets
// after stuffing) and the device entropy decoder (block index of every subsequence, DC predictors).  Three plain passes: no CTA ever
// waits for another one, so nothing can hang.  Included by the translation units that use it (kernels have internal linkage).
#pragma once
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_PER_THREAD 8
#define SCAN_TILE (SCAN_THREADS * SCAN_PER_THREAD)

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long *s_warp, unsigned long long &block_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += up;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    unsigned long long before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) { const unsigned long long x = s_warp[w]; if (w < warp) before += x; total += x; }
    __syncthreads();
    block_total = total;
    return before + inc - v;
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_partials_kernel(const uint32_t *__restrict__ in, long long n, unsigned long long *__restrict__ partial)
{
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_PER_THREAD;
    unsigned long long v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) if (base + k < n) v += in[base + k];
    unsigned long long total;
    block_exclusive_scan(v, s_warp, total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

// one CTA: partial[] -> exclusive prefix in place, grand total to *total
static __global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(unsigned long long *partial, int n_tiles, unsigned long long *total)
{
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    unsigned long long carry = 0;
    for (int base = 0; base < n_tiles; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < n_tiles ? partial[i] : 0;
        unsigned long long chunk_total;
        const unsigned long long ex = block_exclusive_scan(v, s_warp, chunk_total);
        if (i < n_tiles) partial[i] = carry + ex;
        carry += chunk_total;
    }
    if (threadIdx.x == 0) *total = carry;
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_final_kernel(const uint32_t *__restrict__ in, long long n, const unsigned long long *__restrict__ partial,
                                                                  unsigned long long *__restrict__ out)
{
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_PER_THREAD;
    uint32_t x[SCAN_PER_THREAD];
    unsigned long long v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) { x[k] = base + k < n ? in[base + k] : 0; v += x[k]; }
    unsigned long long total;
    unsigned long long run = partial[blockIdx.x] + block_exclusive_scan(v, s_warp, total);
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) { if (base + k < n) out[base + k] = run; run += x[k]; }
}

static int exclusive_scan(vfsms_ctx *ctx, DevBuf &partial_buf, const uint32_t *in, long long n, unsigned long long *out, unsigned long long *total_dev, cudaStream_t st)
{
    const int n_tiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    int rc;
    if ((rc = partial_buf.reserve((size_t)n_tiles * 8))) return rc;
    scan_partials_kernel<<<n_tiles, SCAN_THREADS, 0, st>>>(in, n, partial_buf.as<unsigned long long>()); LAUNCH_CHECK(ctx);
    scan_spine_kernel<<<1, SCAN_THREADS, 0, st>>>(partial_buf.as<unsigned long long>(), n_tiles, total_dev); LAUNCH_CHECK(ctx);
    scan_final_kernel<<<n_tiles, SCAN_THREADS, 0, st>>>(in, n, partial_buf.as<unsigned long long>(), out); LAUNCH_CHECK(ctx);
    return 0;
}

