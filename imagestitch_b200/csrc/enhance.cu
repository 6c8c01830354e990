// enhance.cu -- optional ROI pre-processing of the alignment path: histogram equalisation and CLAHE, sm_100a.
//
// Replaces cv2.equalizeHist and cv2.createCLAHE(clipLimit, (tile, tile)).apply as called at Stitcher.py:269-276 and
// 327-334 when isEnhance is set (off by default).  Integer histograms / LUTs follow the
// published OpenCV algorithms; the CLAHE bilinear LUT interpolation keeps OpenCV's float expression order.
//   hist_kernel        per-tile 256-bin histograms (shared-memory privatised), tiles over a REFLECT_101 padded image
//   lut_kernel         one warp-sized CTA per tile: clip + redistribute (CLAHE) or plain CDF (equalizeHist) -> u8 LUT
//   apply_kernel       per pixel LUT lookup (equalizeHist) or 4-LUT bilinear blend (CLAHE)
// HBM-bound: 1 B/px read for the histogram, 1 B/px read + 1 B/px written for apply.
#include "common.cuh"

__device__ __forceinline__ int refl101_e(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * n - 2 - i; }
    return i;
}

// grid (tiles_x, tiles_y): histogram of one tile of the (virtually) padded image
__global__ void __launch_bounds__(256) enh_hist_kernel(const uint8_t *__restrict__ img, int rows, int cols, int stride, int tile_w, int tile_h, int *hist)
{
    __shared__ int s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const int x0 = blockIdx.x * tile_w, y0 = blockIdx.y * tile_h;
    for (int i = threadIdx.x; i < tile_w * tile_h; i += 256) {
        const int y = refl101_e(y0 + i / tile_w, rows), x = refl101_e(x0 + i % tile_w, cols);
        atomicAdd(&s_h[img[(size_t)y * stride + x]], 1);
    }
    __syncthreads();
    hist[(blockIdx.y * gridDim.x + blockIdx.x) * 256 + threadIdx.x] = s_h[threadIdx.x];
}

// one thread per tile does the (tiny, sequential) OpenCV logic
__global__ void enh_lut_kernel(const int *__restrict__ hist_all, int n_tiles, int tile_total, int clip_limit, int mode, uint8_t *lut_all, int *flat_value)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    int hist[256];
    for (int i = 0; i < 256; i++) hist[i] = hist_all[t * 256 + i];
    uint8_t *lut = lut_all + t * 256;
    if (mode == 0) {                                   // cv::equalizeHist
        int i = 0;
        while (!hist[i]) ++i;
        if (hist[i] == tile_total) { *flat_value = i; for (int k = 0; k < 256; k++) lut[k] = (uint8_t)i; return; }
        *flat_value = -1;
        const float scale = (256 - 1.f) / (tile_total - hist[i]);
        int sum = 0;
        for (int k = 0; k <= i; k++) lut[k] = 0;
        for (++i; i < 256; ++i) { sum += hist[i]; const int v = __float2int_rn(sum * scale); lut[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
        return;
    }
    if (clip_limit > 0) {                              // CLAHE_CalcLut_Body
        int clipped = 0;
        for (int i = 0; i < 256; ++i) if (hist[i] > clip_limit) { clipped += hist[i] - clip_limit; hist[i] = clip_limit; }
        const int redistBatch = clipped / 256;
        int residual = clipped - redistBatch * 256;
        for (int i = 0; i < 256; ++i) hist[i] += redistBatch;
        if (residual != 0) {
            const int residualStep = max(256 / residual, 1);
            for (int i = 0; i < 256 && residual > 0; i += residualStep, residual--) hist[i]++;
        }
    }
    const float lutScale = (float)(256 - 1) / tile_total;
    int sum = 0;
    for (int i = 0; i < 256; ++i) { sum += hist[i]; const int v = __float2int_rn(sum * lutScale); lut[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
}

__global__ void __launch_bounds__(256) enh_apply_kernel(const uint8_t *__restrict__ img, int rows, int cols, int stride, const uint8_t *__restrict__ lut,
                                                        int mode, int tiles_x, int tiles_y, int tile_w, int tile_h, uint8_t *out)
{
    const float inv_tw = 1.0f / tile_w, inv_th = 1.0f / tile_h;
    const int64_t total = (int64_t)rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / cols), x = (int)(i - (int64_t)y * cols);
        const int v = img[(size_t)y * stride + x];
        if (mode == 0) { out[i] = lut[v]; continue; }
        // CLAHE_Interpolation_Body
        const float tyf = y * inv_th - 0.5f;
        int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
        const float ya = tyf - ty1, ya1 = 1.0f - ya;
        ty1 = max(ty1, 0); ty2 = min(ty2, tiles_y - 1);
        const float txf = x * inv_tw - 0.5f;
        int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
        const float xa = txf - tx1, xa1 = 1.0f - xa;
        tx1 = max(tx1, 0); tx2 = min(tx2, tiles_x - 1);
        const uint8_t *l1 = lut + (size_t)ty1 * tiles_x * 256, *l2 = lut + (size_t)ty2 * tiles_x * 256;
        const float res = (l1[tx1 * 256 + v] * xa1 + l1[tx2 * 256 + v] * xa) * ya1 + (l2[tx1 * 256 + v] * xa1 + l2[tx2 * 256 + v] * xa) * ya;
        const int r = __float2int_rn(res);
        out[i] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
    }
}

extern "C" int vfsms_enhance_host(vfsms_ctx *ctx, const uint8_t *image, int rows, int cols, int stride, int mode, double clip_limit,
                                  int tile_grid, uint8_t *out)
{
    if (!ctx || !image || !out || rows < 1 || cols < 1 || stride < cols || (mode != 0 && mode != 1) || (mode == 1 && tile_grid < 1)) {
        vfsms_set_error("enhance: bad arguments"); return VFSMS_E_ARG;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    int tiles_x = 1, tiles_y = 1, tile_w = cols, tile_h = rows;
    if (mode == 1) {
        tiles_x = tiles_y = tile_grid;
        // OpenCV (CLAHE_Impl::apply) pads (REFLECT_101) to a multiple of the grid and takes the tiles over the padded size.  It pads
        // BOTH dimensions as soon as EITHER is not divisible: right = tilesX - cols % tilesX, bottom = tilesY - rows % tilesY, so an
        // already divisible side grows by one whole tile count.
        const bool exact = cols % tiles_x == 0 && rows % tiles_y == 0;
        const int pc = exact ? cols : cols + (tiles_x - cols % tiles_x);
        const int pr = exact ? rows : rows + (tiles_y - rows % tiles_y);
        tile_w = pc / tiles_x; tile_h = pr / tiles_y;
    }
    const int n_tiles = tiles_x * tiles_y, tile_total = tile_w * tile_h;
    int clip = 0;
    if (mode == 1 && clip_limit > 0.0) { clip = (int)(clip_limit * tile_total / 256); if (clip < 1) clip = 1; }
    int rc;
    if ((rc = ctx->img_a.reserve((size_t)rows * cols))) return rc;
    if ((rc = ctx->img_b.reserve((size_t)rows * cols))) return rc;
    if ((rc = ctx->scratch0.reserve((size_t)n_tiles * 256 * 4 + 16))) return rc;
    if ((rc = ctx->scratch1.reserve((size_t)n_tiles * 256))) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(ctx->img_a.p, cols, image, stride, cols, rows, cudaMemcpyHostToDevice, st));
    int *hist = ctx->scratch0.as<int>(), *flat = hist + (size_t)n_tiles * 256;
    enh_hist_kernel<<<dim3(tiles_x, tiles_y), 256, 0, st>>>(ctx->img_a.as<uint8_t>(), rows, cols, cols, tile_w, tile_h, hist);
    LAUNCH_CHECK(ctx);
    enh_lut_kernel<<<ceil_div(n_tiles, 32), 32, 0, st>>>(hist, n_tiles, tile_total, clip, mode, ctx->scratch1.as<uint8_t>(), flat);
    LAUNCH_CHECK(ctx);
    const int64_t n = (int64_t)rows * cols;
    const int grid = (int)((n + 255) / 256 < (int64_t)ctx->num_sms * 8 ? (n + 255) / 256 : (int64_t)ctx->num_sms * 8);
    enh_apply_kernel<<<grid, 256, 0, st>>>(ctx->img_a.as<uint8_t>(), rows, cols, cols, ctx->scratch1.as<uint8_t>(), mode, tiles_x, tiles_y, tile_w, tile_h,
                                           ctx->img_b.as<uint8_t>());
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(out, ctx->img_b.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}
