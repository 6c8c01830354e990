// orb.cu -- ORB detect + describe (oFAST + Harris ranking + intensity-centroid angle + steered rBRIEF-256), sm_100a.
//
// Replaces cv2.ORB_create(...).detectAndCompute (ImageUtility.py:260,262) and myGpuFeatures.detectAndDescribeByOrb
// (appendix/myGpuFeatures.cpp:106-146; descriptors handed out as byte values in float32, cpp:118).
// Pipeline (published ORB algorithm as OpenCV implements it):
//   pyramid_kernel      8 levels, scale 1.2, bilinear from the previous level
//   fast_score_kernel   FAST-9/16 corner test + corner score per pixel (smem tile with a 3-px halo)
//   fast_nms_kernel     3x3 strict non-max suppression, border filter (edgeThreshold), per-level candidate lists + score histogram
//   select_fast_kernel  retainBest(2N) per level by FAST score (histogram threshold; ties kept like KeyPointsFilter::retainBest)
//   harris_kernel       7x7 Harris response of the survivors
//   select_harris_kernel retainBest(N) per level by Harris response (rank counting), deterministic order: level, response desc, y, x
//   blur_kernel         7x7 Gaussian (sigma 2) per level for the descriptor
//   describe_kernel     IC angle (integer moments over the circular patch) + 256 steered binary tests
// Integer stages (FAST test/score, NMS, Harris sums, moments) are exact; the pyramid / blur use float arithmetic with
// rounding to u8, which differs from OpenCV's fixed-point resize / blur by at most one gray level on a few pixels
// (parity target of the survey for ORB: keypoint-set overlap and offsets within 1 px).
#include "common.cuh"
#include <math.h>
#include <float.h>

#define ORB_MAX_LEVELS 12
#define HARRIS_K 0.04f

struct OrbLevel { int rows, cols; long long offset; float scale; int n_features; };
struct OrbPlan { int n_levels; int edge, fast_thr, half_patch; OrbLevel lv[ORB_MAX_LEVELS]; int umax[32]; };

struct OrbState {
    DevBuf pyr, blur, score, cand, cand2, counts, hist, kp, desc, thr;
};

__constant__ signed char c_orb_pattern[1024] = {
    8, -3, 9, 5, 4, 2, 7, -12, -11, 9, -8, 2, 7, -12, 12, -13,
    2, -13, 2, 12, 1, -7, 1, 6, -2, -10, -2, -4, -13, -13, -11, -8,
    -13, -3, -12, -9, 10, 4, 11, 9, -13, -8, -8, -9, -11, 7, -9, 12,
    7, 7, 12, 6, -4, -5, -3, 0, -13, 2, -12, -3, -9, 0, -7, 5,
    12, -6, 12, -1, -3, 6, -2, 12, -6, -13, -4, -8, 11, -13, 12, -8,
    4, 7, 5, 1, 5, -3, 10, -3, 3, -7, 6, 12, -8, -7, -6, -2,
    -2, 11, -1, -10, -13, 12, -8, 10, -7, 3, -5, -3, -4, 2, -3, 7,
    -10, -12, -6, 11, 5, -12, 6, -7, 5, -6, 7, -1, 1, 0, 4, -5,
    9, 11, 11, -13, 4, 7, 4, 12, 2, -1, 4, 4, -4, -12, -2, 7,
    -8, -5, -7, -10, 4, 11, 9, 12, 0, -8, 1, -13, -13, -2, -8, 2,
    -3, -2, -2, 3, -6, 9, -4, -9, 8, 12, 10, 7, 0, 9, 1, 3,
    7, -5, 11, -10, -13, -6, -11, 0, 10, 7, 12, 1, -6, -3, -6, 12,
    10, -9, 12, -4, -13, 8, -8, -12, -13, 0, -8, -4, 3, 3, 7, 8,
    5, 7, 10, -7, -1, 7, 1, -12, 3, -10, 5, 6, 2, -4, 3, -10,
    -13, 0, -13, 5, -13, -7, -12, 12, -13, 3, -11, 8, -7, 12, -4, 7,
    6, -10, 12, 8, -9, -1, -7, -6, -2, -5, 0, 12, -12, 5, -7, 5,
    3, -10, 8, -13, -7, -7, -4, 5, -3, -2, -1, -7, 2, 9, 5, -11,
    -11, -13, -5, -13, -1, 6, 0, -1, 5, -3, 5, 2, -4, -13, -4, 12,
    -9, -6, -9, 6, -12, -10, -8, -4, 10, 2, 12, -3, 7, 12, 12, 12,
    -7, -13, -6, 5, -4, 9, -3, 4, 7, -1, 12, 2, -7, 6, -5, 1,
    -13, 11, -12, 5, -3, 7, -2, -6, 7, -8, 12, -7, -13, -7, -11, -12,
    1, -3, 12, 12, 2, -6, 3, 0, -4, 3, -2, -13, -1, -13, 1, 9,
    7, 1, 8, -6, 1, -1, 3, 12, 9, 1, 12, 6, -1, -9, -1, 3,
    -13, -13, -10, 5, 7, 7, 10, 12, 12, -5, 12, 9, 6, 3, 7, 11,
    5, -13, 6, 10, 2, -12, 2, 3, 3, 8, 4, -6, 2, 6, 12, -13,
    9, -12, 10, 3, -8, 4, -7, 9, -11, 12, -4, -6, 1, 12, 2, -8,
    6, -9, 7, -4, 2, 3, 3, -2, 6, 3, 11, 0, 3, -3, 8, -8,
    7, 8, 9, 3, -11, -5, -6, -4, -10, 11, -5, 10, -5, -8, -3, 12,
    -10, 5, -9, 0, 8, -1, 12, -6, 4, -6, 6, -11, -10, 12, -8, 7,
    4, -2, 6, 7, -2, 0, -2, 12, -5, -8, -5, 2, 7, -6, 10, 12,
    -9, -13, -8, -8, -5, -13, -5, -2, 8, -8, 9, -13, -9, -11, -9, 0,
    1, -8, 1, -2, 7, -4, 9, 1, -2, 1, -1, -4, 11, -6, 12, -11,
    -12, -9, -6, 4, 3, 7, 7, 12, 5, 5, 10, 8, 0, -4, 2, 8,
    -9, 12, -5, -13, 0, 7, 2, 12, -1, 2, 1, 7, 5, 11, 7, -9,
    3, 5, 6, -8, -13, -4, -8, 9, -5, 9, -3, -3, -4, -7, -3, -12,
    6, 5, 8, 0, -7, 6, -6, 12, -13, 6, -5, -2, 1, -10, 3, 10,
    4, 1, 8, -4, -2, -2, 2, -13, 2, -12, 12, 12, -2, -13, 0, -6,
    4, 1, 9, 3, -6, -10, -3, -5, -3, -13, -1, 1, 7, 5, 12, -11,
    4, -2, 5, -7, -13, 9, -9, -5, 7, 1, 8, 6, 7, -8, 7, 6,
    -7, -4, -7, 1, -8, 11, -7, -8, -13, 6, -12, -8, 2, 4, 3, 9,
    10, -5, 12, 3, -6, -5, -6, 7, 8, -3, 9, -8, 2, -12, 2, 8,
    -11, -2, -10, 3, -12, -13, -7, -9, -11, 0, -10, -5, 5, -3, 11, 8,
    -2, -13, -1, 12, -1, -8, 0, 9, -13, -11, -12, -5, -10, -2, -10, 11,
    -3, 9, -2, -13, 2, -3, 3, 2, -9, -13, -4, 0, -4, 6, -3, -10,
    -4, 12, -2, -7, -6, -11, -4, 9, 6, -3, 6, 11, -13, 11, -5, 5,
    11, 11, 12, 6, 7, -5, 12, -2, -1, 12, 0, 7, -4, -8, -3, -2,
    -7, 1, -6, 7, -13, -12, -8, -13, -7, -2, -6, -8, -8, 5, -6, -9,
    -5, -1, -4, 5, -13, 7, -8, 10, 1, 5, 5, -13, 1, 0, 10, -13,
    9, 12, 10, -1, 5, -8, 10, -9, -1, 11, 1, -13, -9, -3, -6, 2,
    -1, -10, 1, 12, -13, 1, -8, -10, 8, -11, 10, -6, 2, -13, 3, -6,
    7, -13, 12, -9, -10, -10, -5, -7, -10, -8, -8, -13, 4, -6, 8, 5,
    3, 12, 8, -13, -4, 2, -3, -3, 5, -13, 10, -12, 4, -13, 5, -1,
    -9, 9, -4, 3, 0, 3, 3, -9, -12, 1, -6, 1, 3, 2, 4, -8,
    -10, -10, -10, 9, 8, -13, 12, 12, -8, -12, -6, -5, 2, 2, 3, 7,
    10, 6, 11, -8, 6, 8, 8, -12, -7, 10, -6, 5, -3, -9, -3, 9,
    -1, -13, -1, 5, -3, -7, -3, 4, -8, -2, -8, 3, 4, 2, 12, 12,
    2, -5, 3, 11, 6, -9, 11, -13, 3, -1, 7, 12, 11, -1, 12, 4,
    -3, 0, -3, 6, 4, -11, 4, 12, 2, -4, 2, 1, -10, -6, -8, 1,
    -13, 7, -11, 1, -13, 12, -11, -13, 6, 0, 11, -13, 0, -1, 1, 4,
    -13, 3, -9, -2, -9, 8, -6, -3, -13, -6, -8, -2, 5, -9, 8, 10,
    2, 7, 3, -9, -1, -6, -1, -1, 9, 5, 11, -2, 11, -3, 12, -8,
    3, 0, 3, 5, -1, 4, 0, 10, 3, -6, 4, 5, -13, 0, -10, 5,
    5, 8, 12, 11, 8, 9, 9, -6, 7, -4, 8, -12, -10, 4, -10, 9,
    7, 3, 12, 4, 9, -7, 10, -2, 7, 0, 12, -2, -1, -6, 0, -11,
};

static OrbState *ostate(vfsms_ctx *ctx)
{
    if (!ctx->orb_state) ctx->orb_state = new OrbState();
    return (OrbState *)ctx->orb_state;
}

void orb_state_destroy(vfsms_ctx *ctx)
{
    OrbState *s = (OrbState *)ctx->orb_state;
    if (!s) return;
    DevBuf *b[] = { &s->pyr, &s->blur, &s->score, &s->cand, &s->cand2, &s->counts, &s->hist, &s->kp, &s->desc, &s->thr };
    for (DevBuf *x : b) x->release();
    delete s;
    ctx->orb_state = nullptr;
}

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * n - 2 - i; }
    return i;
}

// level l (> 0) from level l-1: bilinear at pixel centres, float weights, round to nearest
__global__ void __launch_bounds__(256) pyramid_kernel(const uint8_t *__restrict__ src, int srows, int scols, uint8_t *dst, int drows, int dcols)
{
    const float sx = (float)scols / dcols, sy = (float)srows / drows;
    const int64_t total = (int64_t)drows * dcols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / dcols), x = (int)(i - (int64_t)y * dcols);
        float fx = (x + 0.5f) * sx - 0.5f, fy = (y + 0.5f) * sy - 0.5f;
        int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
        fx -= x0; fy -= y0;
        if (x0 < 0) { x0 = 0; fx = 0; }
        if (y0 < 0) { y0 = 0; fy = 0; }
        int x1 = x0 + 1, y1 = y0 + 1;
        if (x1 >= scols) { x1 = scols - 1; if (x0 >= scols) x0 = scols - 1; }
        if (y1 >= srows) { y1 = srows - 1; if (y0 >= srows) y0 = srows - 1; }
        const float p00 = src[(size_t)y0 * scols + x0], p01 = src[(size_t)y0 * scols + x1];
        const float p10 = src[(size_t)y1 * scols + x0], p11 = src[(size_t)y1 * scols + x1];
        const float v = (p00 * (1.f - fx) + p01 * fx) * (1.f - fy) + (p10 * (1.f - fx) + p11 * fx) * fy;
        dst[i] = (uint8_t)__float2int_rn(v);
    }
}

__global__ void __launch_bounds__(256) copy_strided_kernel(const uint8_t *__restrict__ src, int stride, uint8_t *dst, int rows, int cols)
{
    const int64_t total = (int64_t)rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / cols), x = (int)(i - (int64_t)y * cols);
        dst[i] = src[(size_t)y * stride + x];
    }
}

// FAST-9/16.  circle offsets in OpenCV's order
__constant__ int c_fast_dx[16] = { 0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1 };
__constant__ int c_fast_dy[16] = { 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3 };

#define FT 32
__global__ void __launch_bounds__(FT * 8) fast_score_kernel(const uint8_t *__restrict__ img, int rows, int cols, int thr, uint8_t *score)
{
    __shared__ uint8_t tile[8 + 6][FT + 6];
    const int x0 = blockIdx.x * FT, y0 = blockIdx.y * 8;
    for (int i = threadIdx.y * FT + threadIdx.x; i < (8 + 6) * (FT + 6); i += FT * 8) {
        const int ty = i / (FT + 6), tx = i - ty * (FT + 6);
        const int y = min(max(y0 + ty - 3, 0), rows - 1), x = min(max(x0 + tx - 3, 0), cols - 1);
        tile[ty][tx] = img[(size_t)y * cols + x];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= cols || y >= rows) return;
    int sc = 0;
    if (x >= 3 && y >= 3 && x < cols - 3 && y < rows - 3) {
        const int v = tile[threadIdx.y + 3][threadIdx.x + 3];
        int d[25];
#pragma unroll
        for (int k = 0; k < 16; k++) d[k] = v - (int)tile[threadIdx.y + 3 + c_fast_dy[k]][threadIdx.x + 3 + c_fast_dx[k]];
#pragma unroll
        for (int k = 16; k < 25; k++) d[k] = d[k - 16];
        // corner test: 9 contiguous circle pixels all darker than v - thr (d > thr) or all brighter (d < -thr)
        unsigned mb = 0, md = 0;
#pragma unroll
        for (int k = 0; k < 16; k++) { md |= (d[k] > thr ? 1u : 0u) << k; mb |= (d[k] < -thr ? 1u : 0u) << k; }
        md |= md << 16; mb |= mb << 16;
        bool corner = false;
#pragma unroll
        for (int k = 0; k < 16; k++) { const unsigned m9 = 0x1ffu << k; corner |= ((md & m9) == m9) | ((mb & m9) == m9); }
        if (corner) {
            // cornerScore<16>: the largest threshold for which the pixel is still a corner
            int a0 = thr;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                int a = min(d[k + 1], d[k + 2]); a = min(a, d[k + 3]);
                if (a <= a0) continue;
                a = min(a, d[k + 4]); a = min(a, d[k + 5]); a = min(a, d[k + 6]); a = min(a, d[k + 7]); a = min(a, d[k + 8]);
                a0 = max(a0, min(a, d[k])); a0 = max(a0, min(a, d[k + 9]));
            }
            int b0 = -a0;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                int b = max(d[k + 1], d[k + 2]); b = max(b, d[k + 3]); b = max(b, d[k + 4]); b = max(b, d[k + 5]);
                if (b >= b0) continue;
                b = max(b, d[k + 6]); b = max(b, d[k + 7]); b = max(b, d[k + 8]);
                b0 = min(b0, max(b, d[k])); b0 = min(b0, max(b, d[k + 9]));
            }
            sc = -b0 - 1;
        }
    }
    score[(size_t)y * cols + x] = (uint8_t)min(sc, 255);
}

// candidates: int4 (x, y, level, score)
__global__ void __launch_bounds__(256) fast_nms_kernel(const uint8_t *__restrict__ score, int rows, int cols, int level, int edge,
                                                       int4 *cand, int cand_cap, int *counts, int *hist)
{
    const int64_t total = (int64_t)rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / cols), x = (int)(i - (int64_t)y * cols);
        const int s = score[i];
        if (s == 0 || x < 4 || y < 4 || x >= cols - 4 || y >= rows - 4) continue;      // FAST itself skips a 3-px frame; NMS needs neighbours
        const uint8_t *p = score + i;
        if (!(s > p[-1] && s > p[1] && s > p[-cols - 1] && s > p[-cols] && s > p[-cols + 1] && s > p[cols - 1] && s > p[cols] && s > p[cols + 1])) continue;
        // KeyPointsFilter::runByImageBorder: keep points inside Rect(edge, edge, cols - 2 edge, rows - 2 edge)
        if (x < edge || y < edge || x >= cols - edge || y >= rows - edge) continue;
        const int slot = atomicAdd(&counts[level], 1);
        if (slot < cand_cap) cand[(size_t)level * cand_cap + slot] = make_int4(x, y, level, s);
        atomicAdd(&hist[level * 256 + s], 1);
    }
}

// per level: FAST-score threshold keeping the best 2N (ties kept).  one thread per level.
__global__ void select_fast_kernel(const OrbPlan plan, const int *hist, const int *counts, int *thr_out, int cand_cap, int harris)
{
    const int l = threadIdx.x;
    if (l >= plan.n_levels) return;
    const int want = (harris ? 2 : 1) * plan.lv[l].n_features;
    const int n = min(counts[l], cand_cap);
    int thr = 0;
    if (n > want && want > 0) {
        int cum = 0;
        for (int s = 255; s >= 0; s--) { cum += hist[l * 256 + s]; if (cum >= want) { thr = s; break; } }
    } else if (want <= 0) thr = 256;
    thr_out[l] = thr;
}

// Harris response of candidates that pass the FAST threshold; writes float4 (x, y, response, level) compacted per level
__global__ void __launch_bounds__(128) harris_kernel(const OrbPlan plan, const uint8_t *__restrict__ pyr, const int4 *__restrict__ cand, int cand_cap,
                                                     const int *counts, const int *thr, float4 *cand2, int *counts2)
{
    const int l = blockIdx.y;
    const int n = min(counts[l], cand_cap);
    const OrbLevel L = plan.lv[l];
    const uint8_t *img = pyr + L.offset;
    const int step = L.cols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = cand[(size_t)l * cand_cap + i];
        if (c.w < thr[l]) continue;
        int a = 0, b = 0, cc = 0;
        for (int dy = -3; dy <= 3; dy++)
            for (int dx = -3; dx <= 3; dx++) {
                const int y = reflect101(c.y + dy, L.rows), x = reflect101(c.x + dx, L.cols);
                const int ym = reflect101(y - 1, L.rows), yp = reflect101(y + 1, L.rows), xm = reflect101(x - 1, L.cols), xp = reflect101(x + 1, L.cols);
                const int Ix = ((int)img[(size_t)y * step + xp] - img[(size_t)y * step + xm]) * 2 + ((int)img[(size_t)ym * step + xp] - img[(size_t)ym * step + xm]) +
                               ((int)img[(size_t)yp * step + xp] - img[(size_t)yp * step + xm]);
                const int Iy = ((int)img[(size_t)yp * step + x] - img[(size_t)ym * step + x]) * 2 + ((int)img[(size_t)yp * step + xm] - img[(size_t)ym * step + xm]) +
                               ((int)img[(size_t)yp * step + xp] - img[(size_t)ym * step + xp]);
                a += Ix * Ix; b += Iy * Iy; cc += Ix * Iy;
            }
        const float scale = 1.f / ((1 << 2) * 7 * 255.f);
        const float s4 = scale * scale * scale * scale;
        const float resp = ((float)a * b - (float)cc * cc - HARRIS_K * ((float)a + b) * ((float)a + b)) * s4;
        const int slot = atomicAdd(&counts2[l], 1);
        cand2[(size_t)l * cand_cap + slot] = make_float4((float)c.x, (float)c.y, resp, (float)c.w);
    }
}

// per level: rank by (response desc, y asc, x asc); keep rank < N plus ties with the N-th response; write to the final list.
// grid (chunks, levels).  final order: level-major, rank order.
__global__ void __launch_bounds__(256) select_harris_kernel(const OrbPlan plan, const float4 *__restrict__ cand2, int cand_cap, const int *counts2,
                                                            float4 *sorted, int *counts3)
{
    __shared__ float4 s_c[256];
    const int l = blockIdx.y;
    const int n = counts2[l];
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (blockIdx.x * 256 >= n) return;
    const float4 *C = cand2 + (size_t)l * cand_cap;
    float4 me = make_float4(0, 0, 0, 0);
    if (i < n) me = C[i];
    int rank = 0;
    for (int base = 0; base < n; base += 256) {
        __syncthreads();
        if (base + (int)threadIdx.x < n) s_c[threadIdx.x] = C[base + threadIdx.x];
        __syncthreads();
        const int m = min(256, n - base);
        if (i < n)
            for (int q = 0; q < m; q++) {
                const float4 o = s_c[q];
                const bool before = o.z > me.z || (o.z == me.z && (o.y < me.y || (o.y == me.y && o.x < me.x)));
                rank += before ? 1 : 0;
            }
    }
    if (i < n) { sorted[(size_t)l * cand_cap + rank] = me; }
    if (blockIdx.x == 0 && threadIdx.x == 0) counts3[l] = n;
}

// keep = min(n, N) + ties at the boundary; emits the global prefix.  one thread.
__global__ void finalize_counts_kernel(const OrbPlan plan, const float4 *__restrict__ sorted, int cand_cap, const int *counts3, int *keep, int *prefix, int harris)
{
    if (threadIdx.x || blockIdx.x) return;
    int acc = 0;
    for (int l = 0; l < plan.n_levels; l++) {
        const int n = counts3[l], N = plan.lv[l].n_features;
        int k = min(n, N);
        if (k > 0 && k < n) { const float t = sorted[(size_t)l * cand_cap + k - 1].z; while (k < n && sorted[(size_t)l * cand_cap + k].z >= t) k++; }
        keep[l] = k; prefix[l] = acc; acc += k;
    }
    prefix[plan.n_levels] = acc;
}

// 7x7 Gaussian, sigma 2 (separable taps of cv::getGaussianKernel(7, 2)), BORDER_REFLECT_101, float accumulate, round
__constant__ float c_g7[7];
__global__ void __launch_bounds__(256) blur_kernel(const uint8_t *__restrict__ src, uint8_t *dst, int rows, int cols)
{
    const int64_t total = (int64_t)rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / cols), x = (int)(i - (int64_t)y * cols);
        float acc = 0;
        for (int dy = -3; dy <= 3; dy++) {
            const int yy = reflect101(y + dy, rows);
            float row = 0;
            for (int dx = -3; dx <= 3; dx++) row += c_g7[dx + 3] * (float)src[(size_t)yy * cols + reflect101(x + dx, cols)];
            acc += c_g7[dy + 3] * row;
        }
        dst[i] = (uint8_t)min(max(__float2int_rn(acc), 0), 255);
    }
}

__device__ __forceinline__ float orb_fast_atan2(float y, float x)
{
    const float p1 = 0.9997878412794807f * (float)(180 / M_PI), p3 = -0.3258083974640975f * (float)(180 / M_PI);
    const float p5 = 0.1555786518463281f * (float)(180 / M_PI), p7 = -0.04432655554792128f * (float)(180 / M_PI);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) { c = ay / (ax + (float)DBL_EPSILON); c2 = c * c; a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    else { c = ax / (ay + (float)DBL_EPSILON); c2 = c * c; a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c; }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// one warp per keypoint: IC angle, then 256 steered tests (8 per lane)
__global__ void __launch_bounds__(256) orb_describe_kernel(const OrbPlan plan, const uint8_t *__restrict__ pyr, const uint8_t *__restrict__ blur,
                                                           const float4 *__restrict__ sorted, int cand_cap, const int *keep, const int *prefix,
                                                           float *kp_out, float *desc_out, int out_cap)
{
    const int lane = threadIdx.x & 31;
    const int total = min(prefix[plan.n_levels], out_cap);
    for (int item = blockIdx.x * 8 + (threadIdx.x >> 5); item < total; item += gridDim.x * 8) {
        int l = 0;
        while (l + 1 < plan.n_levels && prefix[l + 1] <= item) l++;
        const int k = item - prefix[l];
        const float4 c = sorted[(size_t)l * cand_cap + k];
        const OrbLevel L = plan.lv[l];
        const uint8_t *img = pyr + L.offset, *bl = blur + L.offset;
        const int cx = (int)c.x, cy = (int)c.y, step = L.cols, hp = plan.half_patch;
        // intensity centroid over the circular patch (integer moments)
        int m01 = 0, m10 = 0;
        for (int v = lane; v <= hp; v += 32) {
            const int d = v == 0 ? hp : plan.umax[v];
            int vsum = 0;
            for (int u = -d; u <= d; u++) {
                const int x = reflect101(cx + u, L.cols);
                if (v == 0) m10 += u * (int)img[(size_t)cy * step + x];
                else {
                    const int vp = img[(size_t)reflect101(cy + v, L.rows) * step + x], vm = img[(size_t)reflect101(cy - v, L.rows) * step + x];
                    vsum += vp - vm; m10 += u * (vp + vm);
                }
            }
            m01 += v * vsum;
        }
        for (int o = 16; o; o >>= 1) { m01 += __shfl_xor_sync(0xffffffffu, m01, o); m10 += __shfl_xor_sync(0xffffffffu, m10, o); }
        const float angle = orb_fast_atan2((float)m01, (float)m10);
        const float rad = angle * (float)(M_PI / 180.f);
        const float a = cosf(rad), b = sinf(rad);
        unsigned byte = 0;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            const signed char *pt = c_orb_pattern + (lane * 8 + t) * 4;
            const int x0 = __float2int_rn(pt[0] * a - pt[1] * b), y0 = __float2int_rn(pt[0] * b + pt[1] * a);
            const int x1 = __float2int_rn(pt[2] * a - pt[3] * b), y1 = __float2int_rn(pt[2] * b + pt[3] * a);
            const int t0 = bl[(size_t)reflect101(cy + y0, L.rows) * step + reflect101(cx + x0, L.cols)];
            const int t1 = bl[(size_t)reflect101(cy + y1, L.rows) * step + reflect101(cx + x1, L.cols)];
            byte |= (t0 < t1 ? 1u : 0u) << t;
        }
        desc_out[(size_t)item * 32 + lane] = (float)byte;
        if (lane == 0) {
            float *kp = kp_out + (size_t)item * KP_STRIDE;
            kp[KP_X] = c.x * L.scale; kp[KP_Y] = c.y * L.scale; kp[KP_SIZE] = (float)(2 * hp + 1) * L.scale; kp[KP_ANGLE] = angle;
            kp[KP_RESPONSE] = c.z; kp[KP_OCTAVE] = (float)l; kp[KP_LAPLACIAN] = 0.f; kp[7] = 0.f;
        }
    }
}

static int grid_of(vfsms_ctx *ctx, int64_t n) { int64_t g = (n + 255) / 256; const int64_t m = (int64_t)ctx->num_sms * 8; return (int)(g < 1 ? 1 : (g < m ? g : m)); }

extern "C" int vfsms_orb_detect_and_describe(vfsms_ctx *ctx, const uint8_t *image, int rows, int cols, int stride, int n_features,
                                             float scale_factor, int n_levels, int edge_threshold, int first_level, int wta_k,
                                             int patch_size, int fast_threshold, float *kp_out, float *desc_out, int cap, int *n_out)
{
    if (!ctx || !image || !n_out || rows < 1 || cols < 1 || stride < cols || n_features < 1) { vfsms_set_error("orb: bad arguments"); return VFSMS_E_ARG; }
    if (wta_k != 2 || patch_size != 31 || first_level != 0 || n_levels < 1 || n_levels > ORB_MAX_LEVELS || scale_factor <= 1.f) {
        vfsms_set_error("orb: only WTA_K=2, patchSize=31, firstLevel=0, 1..%d levels are implemented", ORB_MAX_LEVELS); return VFSMS_E_UNSUPPORTED;
    }
    *n_out = 0;
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    OrbState *os = ostate(ctx);
    static bool g7 = false;
    if (!g7) {
        float k[7]; double sum = 0, t[7];
        for (int i = 0; i < 7; i++) { const double x = i - 3; t[i] = exp(-0.5 * x * x / 4.0); sum += t[i]; }
        for (int i = 0; i < 7; i++) k[i] = (float)(t[i] / sum);
        CUDA_TRY(cudaMemcpyToSymbol(c_g7, k, sizeof(k)));
        g7 = true;
    }
    OrbPlan plan; memset(&plan, 0, sizeof(plan));
    plan.n_levels = n_levels; plan.edge = edge_threshold; plan.fast_thr = fast_threshold; plan.half_patch = patch_size / 2;
    {   // features per level (orb.cpp computeKeyPoints) and the circular patch rows
        const float factor = (float)(1.0 / scale_factor);
        float nd = n_features * (1 - factor) / (1 - (float)pow((double)factor, (double)n_levels));
        int sum = 0; long long off = 0;
        for (int l = 0; l < n_levels; l++) {
            const double sc = pow((double)scale_factor, (double)l);
            OrbLevel &L = plan.lv[l];
            L.scale = (float)sc; L.cols = (int)lrint(cols / (float)sc); L.rows = (int)lrint(rows / (float)sc);
            if (L.cols < 1) L.cols = 1;
            if (L.rows < 1) L.rows = 1;
            L.offset = off; off += (long long)L.rows * L.cols;
            if (l < n_levels - 1) { L.n_features = (int)lrint(nd); sum += L.n_features; nd *= factor; }
            else L.n_features = n_features - sum > 0 ? n_features - sum : 0;
        }
        const int hp = plan.half_patch;
        const int vmax = (int)floor(hp * sqrt(2.f) / 2 + 1), vmin = (int)ceil(hp * sqrt(2.f) / 2);
        for (int v = 0; v <= vmax; v++) plan.umax[v] = (int)lrint(sqrt((double)hp * hp - v * v));
        for (int v = hp, v0 = 0; v >= vmin; --v) { while (plan.umax[v0] == plan.umax[v0 + 1]) ++v0; plan.umax[v] = v0; ++v0; }
    }
    const OrbLevel &last = plan.lv[n_levels - 1];
    const long long pyr_bytes = last.offset + (long long)last.rows * last.cols;
    const int cand_cap = (int)(((long long)rows * cols / 16 > 65536 ? (long long)rows * cols / 16 : 65536));
    int rc;
    if ((rc = os->pyr.reserve((size_t)pyr_bytes))) return rc;
    if ((rc = os->blur.reserve((size_t)pyr_bytes))) return rc;
    if ((rc = os->score.reserve((size_t)pyr_bytes))) return rc;
    if ((rc = os->cand.reserve((size_t)n_levels * cand_cap * 16))) return rc;
    if ((rc = os->cand2.reserve((size_t)n_levels * cand_cap * 16 * 2))) return rc;
    if ((rc = os->counts.reserve(ORB_MAX_LEVELS * 4 * 8))) return rc;
    if ((rc = os->hist.reserve(ORB_MAX_LEVELS * 256 * 4))) return rc;
    const int out_cap = n_features * 2 + 1024;
    if ((rc = os->kp.reserve((size_t)out_cap * KP_STRIDE * 4))) return rc;
    if ((rc = os->desc.reserve((size_t)out_cap * 32 * 4))) return rc;
    int *counts = os->counts.as<int>(), *counts2 = counts + ORB_MAX_LEVELS, *counts3 = counts2 + ORB_MAX_LEVELS, *thr = counts3 + ORB_MAX_LEVELS,
        *keep = thr + ORB_MAX_LEVELS, *prefix = keep + ORB_MAX_LEVELS;
    CUDA_TRY(cudaMemsetAsync(os->counts.p, 0, ORB_MAX_LEVELS * 4 * 8, st));
    CUDA_TRY(cudaMemsetAsync(os->hist.p, 0, ORB_MAX_LEVELS * 256 * 4, st));
    if ((rc = ctx->img_a.reserve((size_t)rows * cols))) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(ctx->img_a.p, cols, image, stride, cols, rows, cudaMemcpyHostToDevice, st));
    uint8_t *pyr = os->pyr.as<uint8_t>(), *blur = os->blur.as<uint8_t>(), *score = os->score.as<uint8_t>();
    copy_strided_kernel<<<grid_of(ctx, (int64_t)rows * cols), 256, 0, st>>>(ctx->img_a.as<uint8_t>(), cols, pyr, rows, cols);
    LAUNCH_CHECK(ctx);
    int4 *cand = os->cand.as<int4>();
    float4 *cand2 = os->cand2.as<float4>(), *sorted = cand2 + (size_t)n_levels * cand_cap;
    for (int l = 0; l < n_levels; l++) {
        const OrbLevel &L = plan.lv[l];
        if (l > 0) {
            const OrbLevel &P = plan.lv[l - 1];
            pyramid_kernel<<<grid_of(ctx, (int64_t)L.rows * L.cols), 256, 0, st>>>(pyr + P.offset, P.rows, P.cols, pyr + L.offset, L.rows, L.cols);
            LAUNCH_CHECK(ctx);
        }
        fast_score_kernel<<<dim3(ceil_div(L.cols, FT), ceil_div(L.rows, 8)), dim3(FT, 8), 0, st>>>(pyr + L.offset, L.rows, L.cols, fast_threshold, score + L.offset);
        LAUNCH_CHECK(ctx);
        fast_nms_kernel<<<grid_of(ctx, (int64_t)L.rows * L.cols), 256, 0, st>>>(score + L.offset, L.rows, L.cols, l, edge_threshold, cand, cand_cap, counts, os->hist.as<int>());
        LAUNCH_CHECK(ctx);
        blur_kernel<<<grid_of(ctx, (int64_t)L.rows * L.cols), 256, 0, st>>>(pyr + L.offset, blur + L.offset, L.rows, L.cols);
        LAUNCH_CHECK(ctx);
    }
    select_fast_kernel<<<1, 32, 0, st>>>(plan, os->hist.as<int>(), counts, thr, cand_cap, 1);
    LAUNCH_CHECK(ctx);
    harris_kernel<<<dim3(ctx->num_sms, n_levels), 128, 0, st>>>(plan, pyr, cand, cand_cap, counts, thr, cand2, counts2);
    LAUNCH_CHECK(ctx);
    select_harris_kernel<<<dim3(ceil_div(cand_cap, 256), n_levels), 256, 0, st>>>(plan, cand2, cand_cap, counts2, sorted, counts3);
    LAUNCH_CHECK(ctx);
    finalize_counts_kernel<<<1, 32, 0, st>>>(plan, sorted, cand_cap, counts3, keep, prefix, 1);
    LAUNCH_CHECK(ctx);
    orb_describe_kernel<<<ctx->num_sms * 2, 256, 0, st>>>(plan, pyr, blur, sorted, cand_cap, keep, prefix, os->kp.as<float>(), os->desc.as<float>(), out_cap);
    LAUNCH_CHECK(ctx);
    int h_prefix[ORB_MAX_LEVELS + 1], h_counts[ORB_MAX_LEVELS];
    CUDA_TRY(cudaMemcpyAsync(h_prefix, prefix, sizeof(int) * (n_levels + 1), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h_counts, counts, sizeof(int) * n_levels, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int l = 0; l < n_levels; l++) if (h_counts[l] > cand_cap) { vfsms_set_error("orb: candidate overflow on level %d (%d > %d)", l, h_counts[l], cand_cap); return VFSMS_E_OVERFLOW; }
    int n = h_prefix[n_levels];
    if (n > out_cap) n = out_cap;
    *n_out = n;
    if (n > cap) { vfsms_set_error("orb: %d keypoints but capacity %d", n, cap); return VFSMS_E_CAPACITY; }
    if (n > 0) {
        if (kp_out) CUDA_TRY(cudaMemcpyAsync(kp_out, os->kp.p, (size_t)n * KP_STRIDE * 4, cudaMemcpyDeviceToHost, st));
        if (desc_out) CUDA_TRY(cudaMemcpyAsync(desc_out, os->desc.p, (size_t)n * 32 * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return 0;
}
