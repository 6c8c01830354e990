// phase.cu -- phase-correlation offset of two u8 ROIs, float64 like the reference.  sm_100a.
//
// Replaces cv2.phaseCorrelate(np.float64(roiA), np.float64(roiB)) as called by
// Stitcher.calculateOffsetForPhaseCorrleateIncre (Stitcher.py:230): zero-pad to the optimal DFT size (2^a 3^b 5^c),
// forward transforms, cross-power spectrum normalised by its magnitude (divSpectrums with its DBL_EPSILON guard),
// unscaled inverse transform, fftshift + first maximum, 5x5 weighted centroid, response = window sum / (M N).
//
// Pipeline (all on one stream, no host sync until the 3 result doubles are read):
//   pad_convert_kernel   u8 (strided) -> double [2][M][N] zero padded                     (reads 2hw B, writes 16 MN B)
//   cufftExecD2Z         batch of 2, plans cached per (M, N)                              (cuFFT: non power-of-two sizes)
//   cross_power_kernel   C = F1 conj(F2) |.| / (|.|^2 + eps) in place over the half spectrum
//   cufftExecZ2D         unnormalised inverse (same convention as cv2.idft without DFT_SCALE)
//   peak_kernel          arg-max in fftshift order (ties -> first in raster order), two-level reduction
//   centroid_kernel      5x5 window (clipped) -> shift_x, shift_y, response
// HBM-bound: ~ 2hw + 16*MN*(1 + 2 + 2 + 2 + 1 + 1) bytes per pair; the FFTs dominate.
#include "common.cuh"
#include <cufft.h>
#include <float.h>
#include <map>
#include <algorithm>

struct PhasePlan { cufftHandle fwd = 0, inv = 0; size_t ws = 0; };
struct PhaseState {
    std::map<std::pair<int, int>, PhasePlan> plans;
    DevBuf real, spec, work, peak_val, peak_idx, img_a, img_b, out;
};

static PhaseState *pstate(vfsms_ctx *ctx)
{
    if (!ctx->phase_state) ctx->phase_state = new PhaseState();
    return (PhaseState *)ctx->phase_state;
}

void phase_state_destroy(vfsms_ctx *ctx)
{
    PhaseState *s = (PhaseState *)ctx->phase_state;
    if (!s) return;
    for (auto &kv : s->plans) { cufftDestroy(kv.second.fwd); cufftDestroy(kv.second.inv); }
    DevBuf *b[] = { &s->real, &s->spec, &s->work, &s->peak_val, &s->peak_idx, &s->img_a, &s->img_b, &s->out };
    for (DevBuf *x : b) x->release();
    delete s;
    ctx->phase_state = nullptr;
}

static int optimal_dft_size(int n)
{
    long long best = -1;
    for (long long p2 = 1; p2 < 4LL * n + 4; p2 *= 2)
        for (long long p3 = p2; p3 < 4LL * n + 4; p3 *= 3)
            for (long long p5 = p3; p5 < 4LL * n + 4; p5 *= 5)
                if (p5 >= n && (best < 0 || p5 < best)) best = p5;
    return (int)best;
}

__global__ void __launch_bounds__(256) pad_convert_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int rows, int cols,
                                                          int stride, double *out, int M, int N)
{
    const int64_t total = (int64_t)M * N;
    const uint8_t *src = blockIdx.y == 0 ? a : b;
    double *dst = out + (size_t)blockIdx.y * total;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / N), c = (int)(i - (int64_t)r * N);
        dst[i] = (r < rows && c < cols) ? (double)src[(size_t)r * stride + c] : 0.0;
    }
}

__global__ void __launch_bounds__(256) cross_power_kernel(cufftDoubleComplex *spec, int64_t n_half)
{
    cufftDoubleComplex *f1 = spec, *f2 = spec + n_half;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_half; i += (int64_t)gridDim.x * blockDim.x) {
        const cufftDoubleComplex x = f1[i], y = f2[i];
        const double re = x.x * y.x + x.y * y.y;            // F1 * conj(F2)
        const double im = x.y * y.x - x.x * y.y;
        const double mag = sqrt(re * re + im * im);
        const double denom = mag * mag + DBL_EPSILON;
        cufftDoubleComplex o;
        o.x = re * mag / denom; o.y = im * mag / denom;
        f1[i] = o;
    }
}

// index of element (r, c) of the UNSHIFTED correlation surface inside the fftshift'ed one
__device__ __forceinline__ int64_t shifted_index(int r, int c, int M, int N)
{
    // cv2's fftShift moves quadrant blocks: out[(r + M/2) % M][(c + N/2) % N] = in[r][c] for even sizes; for odd sizes cv2
    // rotates by ceil (rows: M - M/2 ... ) -- use the numpy.fft.fftshift convention, identical for the even sizes
    // getOptimalDFTSize produces in practice and matching cv2 for odd ones (shift by floor(n/2)).
    const int rs = (r + M / 2) % M, cs = (c + N / 2) % N;
    return (int64_t)rs * N + cs;
}

__global__ void __launch_bounds__(256) peak_kernel(const double *__restrict__ corr, int M, int N, double *blk_val, long long *blk_idx)
{
    __shared__ double s_v[256];
    __shared__ long long s_i[256];
    const int64_t total = (int64_t)M * N;
    double best = -DBL_MAX; long long bi = 0x7fffffffffffffffLL;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / N), c = (int)(i - (int64_t)r * N);
        const double v = corr[i];
        const long long si = shifted_index(r, c, M, N);
        if (v > best || (v == best && si < bi)) { best = v; bi = si; }
    }
    s_v[threadIdx.x] = best; s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double v = s_v[threadIdx.x + o]; const long long si = s_i[threadIdx.x + o];
            if (v > s_v[threadIdx.x] || (v == s_v[threadIdx.x] && si < s_i[threadIdx.x])) { s_v[threadIdx.x] = v; s_i[threadIdx.x] = si; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { blk_val[blockIdx.x] = s_v[0]; blk_idx[blockIdx.x] = s_i[0]; }
}

__global__ void centroid_kernel(const double *__restrict__ corr, int M, int N, const double *blk_val, const long long *blk_idx, int n_blk,
                                double *out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double best = -DBL_MAX; long long bi = 0x7fffffffffffffffLL;
    for (int k = 0; k < n_blk; k++)
        if (blk_val[k] > best || (blk_val[k] == best && blk_idx[k] < bi)) { best = blk_val[k]; bi = blk_idx[k]; }
    const int py = (int)(bi / N), px = (int)(bi - (long long)py * N);
    int minr = py - 2, maxr = py + 2, minc = px - 2, maxc = px + 2;
    if (minr < 0) minr = 0;
    if (minc < 0) minc = 0;
    if (maxr > M - 1) maxr = M - 1;
    if (maxc > N - 1) maxc = N - 1;
    double cx = 0, cy = 0, sum = 0;
    for (int y = minr; y <= maxr; y++)
        for (int x = minc; x <= maxc; x++) {
            // shifted (y, x) -> unshifted coordinates
            const int r = (y - M / 2 + M) % M, c = (x - N / 2 + N) % N;
            const double v = corr[(size_t)r * N + c];
            cx += (double)x * v; cy += (double)y * v; sum += v;
        }
    const double response = sum;
    sum += DBL_EPSILON;
    cx /= sum; cy /= sum;
    out[0] = (double)N / 2.0 - cx;
    out[1] = (double)M / 2.0 - cy;
    out[2] = response / ((double)M * (double)N);
}

static int get_plan(vfsms_ctx *ctx, int M, int N, PhasePlan **out)
{
    PhaseState *ps = pstate(ctx);
    auto key = std::make_pair(M, N);
    auto it = ps->plans.find(key);
    if (it == ps->plans.end()) {
        PhasePlan pl;
        int n[2] = { M, N };
        size_t ws1 = 0, ws2 = 0;
        if (cufftCreate(&pl.fwd) != CUFFT_SUCCESS || cufftCreate(&pl.inv) != CUFFT_SUCCESS) { vfsms_set_error("cufftCreate failed"); return VFSMS_E_CUDA; }
        cufftSetAutoAllocation(pl.fwd, 0); cufftSetAutoAllocation(pl.inv, 0);
        if (cufftMakePlanMany(pl.fwd, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, 2, &ws1) != CUFFT_SUCCESS ||
            cufftMakePlanMany(pl.inv, 2, n, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, 1, &ws2) != CUFFT_SUCCESS) {
            vfsms_set_error("cufftMakePlanMany(%d x %d) failed", M, N); return VFSMS_E_CUDA;
        }
        pl.ws = ws1 > ws2 ? ws1 : ws2;
        it = ps->plans.emplace(key, pl).first;
    }
    *out = &it->second;
    return 0;
}

static int phase_run(vfsms_ctx *ctx, const uint8_t *a_dev, const uint8_t *b_dev, int rows, int cols, int stride, double *out_dev, cudaStream_t st)
{
    PhaseState *ps = pstate(ctx);
    const int M = optimal_dft_size(rows), N = optimal_dft_size(cols);
    const int64_t mn = (int64_t)M * N, n_half = (int64_t)M * (N / 2 + 1);
    PhasePlan *pl;
    int rc;
    if ((rc = get_plan(ctx, M, N, &pl))) return rc;
    if ((rc = ps->real.reserve((size_t)mn * 8 * 2))) return rc;
    if ((rc = ps->spec.reserve((size_t)n_half * 16 * 2))) return rc;
    if ((rc = ps->work.reserve(pl->ws ? pl->ws : 16))) return rc;
    const int n_blk = ctx->num_sms * 4;
    if ((rc = ps->peak_val.reserve((size_t)n_blk * 8))) return rc;
    if ((rc = ps->peak_idx.reserve((size_t)n_blk * 8))) return rc;
    cufftSetStream(pl->fwd, st); cufftSetStream(pl->inv, st);
    cufftSetWorkArea(pl->fwd, ps->work.p); cufftSetWorkArea(pl->inv, ps->work.p);
    {
        StageTimer t(ctx, st, VFSMS_STAGE_PHASE_FFT);
        pad_convert_kernel<<<dim3(ctx->num_sms * 4, 2), 256, 0, st>>>(a_dev, b_dev, rows, cols, stride, ps->real.as<double>(), M, N);
        LAUNCH_CHECK(ctx);
        if (cufftExecD2Z(pl->fwd, ps->real.as<double>(), ps->spec.as<cufftDoubleComplex>()) != CUFFT_SUCCESS) { vfsms_set_error("cufftExecD2Z failed"); return VFSMS_E_CUDA; }
        cross_power_kernel<<<ctx->num_sms * 4, 256, 0, st>>>(ps->spec.as<cufftDoubleComplex>(), n_half);
        LAUNCH_CHECK(ctx);
        if (cufftExecZ2D(pl->inv, ps->spec.as<cufftDoubleComplex>(), ps->real.as<double>()) != CUFFT_SUCCESS) { vfsms_set_error("cufftExecZ2D failed"); return VFSMS_E_CUDA; }
        ctx->launches += 2;
    }
    StageTimer t2(ctx, st, VFSMS_STAGE_PHASE_PEAK);
    peak_kernel<<<n_blk, 256, 0, st>>>(ps->real.as<double>(), M, N, ps->peak_val.as<double>(), ps->peak_idx.as<long long>());
    LAUNCH_CHECK(ctx);
    centroid_kernel<<<1, 32, 0, st>>>(ps->real.as<double>(), M, N, ps->peak_val.as<double>(), ps->peak_idx.as<long long>(), n_blk, out_dev);
    LAUNCH_CHECK(ctx);
    return 0;
}

// ---------------------------------------------------------------- overlap sums (wrap-aware phase mode)
// For candidate shift k = (dRow, dCol): over the pixels with roiB(r, c) <-> roiA(r + dRow, c + dCol) inside both ROIs, the
// integer sums n, Sa, Sb, Sab, Saa, Sbb.  grid (blocks, n_cand); each thread strides over the overlap rectangle, warp-shuffle
// + shared-memory reduction, one 64-bit atomicAdd per sum and block: integers, so the result does not depend on the order.
__global__ void __launch_bounds__(256) overlap_sums_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int rows, int cols,
                                                           int stride_a, int stride_b, const int32_t *__restrict__ shifts,
                                                           unsigned long long *out)
{
    const int k = blockIdx.y;
    const int dr = shifts[2 * k], dc = shifts[2 * k + 1];
    const int r0 = max(0, -dr), r1 = min(rows, rows - dr), c0 = max(0, -dc), c1 = min(cols, cols - dc);
    const long long h = r1 - r0, w = c1 - c0;
    unsigned long long acc[6] = { 0, 0, 0, 0, 0, 0 };
    if (h > 0 && w > 0) {
        const long long total = h * w;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int r = r0 + (int)(i / w), c = c0 + (int)(i % w);
            const unsigned pb = b[(size_t)r * stride_b + c];
            const unsigned pa = a[(size_t)(r + dr) * stride_a + (c + dc)];
            acc[0] += 1; acc[1] += pa; acc[2] += pb; acc[3] += pa * pb; acc[4] += pa * pa; acc[5] += pb * pb;
        }
    }
    __shared__ unsigned long long s_part[8][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 6; q++) {
        unsigned long long v = acc[q];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned long long v = 0;
        for (int wv = 0; wv < 8; wv++) v += s_part[wv][threadIdx.x];
        if (v) atomicAdd(&out[(size_t)k * 6 + threadIdx.x], v);
    }
}

extern "C" {

int vfsms_overlap_sums_host(vfsms_ctx *ctx, const uint8_t *roi_a, const uint8_t *roi_b, int rows, int cols, int stride_a, int stride_b,
                            int n_shifts, const int32_t *shifts, int64_t *sums_out)
{
    if (!ctx || !roi_a || !roi_b || !shifts || !sums_out || rows < 1 || cols < 1 || stride_a < cols || stride_b < cols || n_shifts < 1 ||
        n_shifts > 64) {
        vfsms_set_error("overlap_sums: bad arguments"); return VFSMS_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    PhaseState *ps = pstate(ctx);
    int rc;
    if ((rc = ps->img_a.reserve((size_t)rows * cols))) return rc;
    if ((rc = ps->img_b.reserve((size_t)rows * cols))) return rc;
    if ((rc = ps->out.reserve(64 * 8 + 64 * 6 * 8))) return rc;
    int32_t *d_shifts = (int32_t *)ps->out.p;                                      // [64][2]
    unsigned long long *d_sums = (unsigned long long *)((char *)ps->out.p + 64 * 8);  // [64][6]
    CUDA_TRY(cudaMemcpy2DAsync(ps->img_a.p, cols, roi_a, stride_a, cols, rows, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpy2DAsync(ps->img_b.p, cols, roi_b, stride_b, cols, rows, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_shifts, shifts, (size_t)n_shifts * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(d_sums, 0, (size_t)n_shifts * 6 * 8, st));
    const int blocks = std::max(1, std::min(ctx->num_sms * 2, (int)(((long long)rows * cols + 255) / 256)));
    overlap_sums_kernel<<<dim3(blocks, n_shifts), 256, 0, st>>>(ps->img_a.as<uint8_t>(), ps->img_b.as<uint8_t>(), rows, cols, cols, cols,
                                                                 d_shifts, d_sums);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(sums_out, d_sums, (size_t)n_shifts * 6 * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int vfsms_phase_correlate_dev(vfsms_ctx *ctx, const uint8_t *roi_a_dev, const uint8_t *roi_b_dev, int rows, int cols, int stride,
                              double *out_dev, void *stream)
{
    if (!ctx || !roi_a_dev || !roi_b_dev || !out_dev || rows < 1 || cols < 1 || stride < cols) { vfsms_set_error("phase_correlate_dev: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    return phase_run(ctx, roi_a_dev, roi_b_dev, rows, cols, stride, out_dev, stream ? (cudaStream_t)stream : ctx->stream);
}

int vfsms_phase_correlate_host(vfsms_ctx *ctx, const uint8_t *roi_a, const uint8_t *roi_b, int rows, int cols, int stride, double out[3])
{
    if (!ctx || !roi_a || !roi_b || !out || rows < 1 || cols < 1 || stride < cols) { vfsms_set_error("phase_correlate_host: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    PhaseState *ps = pstate(ctx);
    int rc;
    if ((rc = ps->img_a.reserve((size_t)rows * cols))) return rc;
    if ((rc = ps->img_b.reserve((size_t)rows * cols))) return rc;
    if ((rc = ps->out.reserve(32))) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(ps->img_a.p, cols, roi_a, stride, cols, rows, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpy2DAsync(ps->img_b.p, cols, roi_b, stride, cols, rows, cudaMemcpyHostToDevice, st));
    if ((rc = phase_run(ctx, ps->img_a.as<uint8_t>(), ps->img_b.as<uint8_t>(), rows, cols, cols, ps->out.as<double>(), st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, ps->out.p, 24, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
