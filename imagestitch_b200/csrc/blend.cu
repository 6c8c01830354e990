// blend.cu -- overlap blending and mosaic assembly on a device-resident canvas, sm_100a.
//
// Replaces Stitcher.fuseImage (Stitcher.py:488-525), ImageFusion.fuseByAverage / Maximum / Minimum / FadeInAndFadeOut /
// Trigonometric and getWeightsMatrix (ImageFusion.py:12-293), and the paste/blend loop of Stitcher.getStitchByOffset
// (Stitcher.py:433-486).  Data model: the canvas holds -1 for "empty"; the reference uses int64,
// here int16 (values are -1..255) -- 4x less HBM traffic, identical results.
//
// Per overlap ROI, three launches and no host synchronisation:
//   blend_stats_kernel   count(A > -1), the four quadrant counts of (A > 0), first/last non-empty row of every column
//   blend_plan_kernel    "normal" vs "corner" decision (ImageFusion.py:209), and for corners a literal transcription of
//                        the data-dependent scans of getWeightsMatrix (ImageFusion.py:62-187, including its early-exit and
//                        index-wrap quirks) producing the two 1-D ramps weightMatB_1 / weightMatB_2
//   blend_apply_kernel   A[A<0] = B, w_A*A + w_B*B in float64, clip, truncate (ImageFusion.py:240-243, 289-291)
// All HBM-bound: ~ (2+2) B/px read for stats, (2+2) B/px read + 1-2 B/px written for apply.
#include "common.cuh"
#include <math.h>

struct BlendPlan {
    int corner;            // 0: 1-D ramp, 1: corner weights
    int index;             // quadrant case 0..3 (ImageFusion.py:62)
    int row_index, col_index;
    long long count_valid; // elements with A > -1
    long long quad[4];     // elements with A > 0 per quadrant: TL, BL, BR, TR (ImageFusion.py:57-60)
};

struct BlendState {
    DevBuf a16, b16, out8, plan, col_top, col_bot, w1, w2, wa_out, wb_out, canvas, tile, tiles_all, pyr;
};

static BlendState *bstate(vfsms_ctx *ctx)
{
    if (!ctx->blend_state) ctx->blend_state = new BlendState();
    return (BlendState *)ctx->blend_state;
}

void blend_state_destroy(vfsms_ctx *ctx)
{
    BlendState *s = (BlendState *)ctx->blend_state;
    if (!s) return;
    DevBuf *b[] = { &s->a16, &s->b16, &s->out8, &s->plan, &s->col_top, &s->col_bot, &s->w1, &s->w2, &s->wa_out, &s->wb_out,
                    &s->canvas, &s->tile, &s->tiles_all, &s->pyr };
    for (DevBuf *x : b) x->release();
    delete s;
    ctx->blend_state = nullptr;
}

static int grid_for(vfsms_ctx *ctx, int64_t n) { int64_t g = (n + 255) / 256; int64_t m = (int64_t)ctx->num_sms * 8; return (int)(g < m ? (g < 1 ? 1 : g) : m); }

// pixel "non-empty" test of getWeightsMatrix: gray: != -1; colour: channel sum != -3 (ImageFusion.py:72,93,...)
__device__ __forceinline__ bool px_nonempty(const int16_t *p, int ch)
{
    if (ch == 1) return p[0] != -1;
    return (int)p[0] + (int)p[1] + (int)p[2] != -3;
}

// A: rows x cols x ch int16 with row stride a_rs (elements).  grid (ceil(cols / 256), ceil(rows / STATS_ROWS)): a thread walks
// STATS_ROWS rows of ONE column (coalesced across the warp), so the first / last non-empty row of the column costs one atomic per
// thread instead of two per pixel (a 2048^2 corner ROI made 8 M contended atomics).
#define STATS_ROWS 32
__global__ void __launch_bounds__(256) blend_stats_kernel(const int16_t *__restrict__ A, int64_t a_rs, int rows, int cols, int ch,
                                                          BlendPlan *plan, int *col_top, int *col_bot)
{
    long long cv = 0, q0 = 0, q1 = 0, q2 = 0, q3 = 0;
    const int hr = rows / 2, hc = cols / 2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r_begin = blockIdx.y * STATS_ROWS, r_end = min(rows, r_begin + STATS_ROWS);
    if (c < cols) {
        int top = rows, bot = -1;
#pragma unroll 8
        for (int r = r_begin; r < r_end; r++) {
            const int16_t *p = A + r * a_rs + (int64_t)c * ch;
            int pos = 0, valid = 0;
            for (int k = 0; k < ch; k++) { pos += p[k] > 0; valid += p[k] > -1; }
            cv += valid;
            if (r < hr) { if (c < hc) q0 += pos; else q3 += pos; }
            else { if (c < hc) q1 += pos; else q2 += pos; }
            if (px_nonempty(p, ch)) { top = min(top, r); bot = r; }
        }
        if (bot >= 0) { atomicMin(&col_top[c], top); atomicMax(&col_bot[c], bot); }
    }
    for (int o = 16; o; o >>= 1) {
        cv += __shfl_xor_sync(0xffffffffu, cv, o); q0 += __shfl_xor_sync(0xffffffffu, q0, o);
        q1 += __shfl_xor_sync(0xffffffffu, q1, o); q2 += __shfl_xor_sync(0xffffffffu, q2, o);
        q3 += __shfl_xor_sync(0xffffffffu, q3, o);
    }
    if ((threadIdx.x & 31) == 0 && (cv | q0 | q1 | q2 | q3)) {
        atomicAdd((unsigned long long *)&plan->count_valid, (unsigned long long)cv);
        atomicAdd((unsigned long long *)&plan->quad[0], (unsigned long long)q0);
        atomicAdd((unsigned long long *)&plan->quad[1], (unsigned long long)q1);
        atomicAdd((unsigned long long *)&plan->quad[2], (unsigned long long)q2);
        atomicAdd((unsigned long long *)&plan->quad[3], (unsigned long long)q3);
    }
}

// python-style index into a length-n axis (negative wraps once, like numpy); clamps what numpy would reject
__device__ __forceinline__ int py_index(int i, int n)
{
    if (i < 0) i += n;
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

__global__ void __launch_bounds__(256) blend_plan_kernel(const int16_t *__restrict__ A, int64_t a_rs, int rows, int cols, int ch,
                                                         BlendPlan *plan, const int *__restrict__ col_top, const int *__restrict__ col_bot,
                                                         float *w1, float *w2, int force_corner)
{
    __shared__ int s_ri_start, s_ri, s_ci_start, s_ci, s_index, s_corner, s_red[8];
    const int row = rows, col = cols;
    // block-wide max of v (v >= -1; the scans below look for the first hit in ascending order as a max of -index)
    auto block_max = [&](int v) {
        for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
        __syncthreads();
        int m = s_red[0];
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) m = max(m, s_red[k]);
        return m;
    };
    if (threadIdx.x == 0) {
        const double ratio = (double)plan->count_valid / ((double)rows * cols * ch);
        int index = 0;
        { long long best = plan->quad[0]; for (int k = 1; k < 4; k++) if (plan->quad[k] < best) { best = plan->quad[k]; index = k; } }
        s_corner = force_corner || !(ratio > 0.65);
        s_index = index;
    }
    __syncthreads();
    const int corner = s_corner, idx = s_index;
    int rowIndex = 0, colIndex = 0;
    if (corner) {
        // The reference scans the columns for the first one with data (ImageFusion.py:63-90 / 92-120 / 122-152 / 154-186): cases 2, 3
        // from the right (columns col-1 .. 1), cases 0, 1 from the left; cases 2, 1 take rowIndex = last non-empty row + 1, cases
        // 3, 0 take first non-empty row - 1 and KEEP SCANNING when that is 0 (first row == 1).  The whole CTA looks for that column.
        const bool from_right = idx == 2 || idx == 3, use_bot = idx == 2 || idx == 1;
        int best = -0x7fffffff;
        for (int c = threadIdx.x + (from_right ? 1 : 0); c < col; c += blockDim.x) {
            const bool ok = use_bot ? col_bot[c] >= 0 : (col_top[c] < row && col_top[c] != 1);
            if (ok) best = max(best, from_right ? c : -c);
        }
        best = block_max(best);
        if (best != -0x7fffffff) { const int c = from_right ? best : -best; rowIndex = use_bot ? col_bot[c] + 1 : col_top[c] - 1; }
        // then the row rowIndex (python indexing) for its last (cases 2, 3) / first (cases 0, 1) non-empty pixel
        const int16_t *rp = A + py_index(rowIndex, row) * a_rs;
        best = -0x7fffffff;
        for (int i = threadIdx.x; i < col; i += blockDim.x)
            if (px_nonempty(rp + (int64_t)i * ch, ch)) best = max(best, from_right ? i : -i);
        best = block_max(best);
        if (best != -0x7fffffff) colIndex = from_right ? best + 1 : -best - 1;
    }
    if (threadIdx.x == 0) {
        plan->corner = corner; plan->index = idx; plan->row_index = rowIndex; plan->col_index = colIndex;
        s_ri_start = rowIndex; s_ci_start = colIndex;
        // "if rowIndex == 0: rowIndex = 1" happens inside the assignment loops (first iteration), ImageFusion.py:85-90 etc.
        s_ri = rowIndex == 0 ? 1 : rowIndex;
        s_ci = colIndex == 0 ? 1 : colIndex;
    }
    __syncthreads();
    if (!s_corner) return;
    const int index = s_index, ri0 = s_ri_start, ri = s_ri, ci0 = s_ci_start, ci = s_ci;
    for (int r = threadIdx.x; r < row; r += blockDim.x) w1[r] = 1.f;
    for (int c = threadIdx.x; c < col; c += blockDim.x) w2[c] = 1.f;
    __syncthreads();
    // row ramp weightMatB_1
    if (index == 2 || index == 1) {
        // for i in range(rowIndex + 1): w[rowIndex - i] = (rowIndex - i) / rowIndex   (rowIndex bumped 0 -> 1 first)
        // loop bound uses the ORIGINAL rowIndex, targets use the bumped one
        for (int i = threadIdx.x; i <= ri0; i += blockDim.x) {
            const int t = ri - i;
            if (ri0 < 0) break;
            const int tt = t < 0 ? t + row : t;
            if (tt >= 0 && tt < row) w1[tt] = (float)((double)(ri - i) / (double)ri);
        }
    } else {
        // for i in range(rowIndex, row): w[i] = (row - i - 1) / (row - rowIndex - 1)
        // (a negative start index wraps to the last rows in numpy, which later iterations overwrite: skipping i < 0 is equivalent)
        for (int i = (ri0 < 0 ? 0 : ri0) + (int)threadIdx.x; i < row; i += blockDim.x) {
            const int den = row - ri - 1;
            w1[i] = den != 0 ? (float)((double)(row - i - 1) / (double)den) : 1.f;
        }
    }
    // column ramp weightMatB_2
    if (index == 2 || index == 3) {
        for (int i = threadIdx.x; i <= ci0; i += blockDim.x) {
            if (ci0 < 0) break;
            const int t = ci - i;
            const int tt = t < 0 ? t + col : t;
            if (tt >= 0 && tt < col) w2[tt] = (float)((double)(ci - i) / (double)ci);
        }
    } else {
        for (int i = (ci0 < 0 ? 0 : ci0) + (int)threadIdx.x; i < col; i += blockDim.x) {
            const int den = col - ci - 1;
            w2[i] = den != 0 ? (float)((double)(col - i - 1) / (double)den) : 1.f;
        }
    }
}

// method_flags: VFSMS_FUSE_* | 0x200 (raw: average / maximum / minimum without the -1 -> 0 and mutual zero fill of Stitcher.fuseImage --
// the semantics of a direct ImageFusion.fuseBy* call).  out16 may be B's own memory (the mosaic blends in place): A and B carry no
// __restrict__, and every element is read before the same thread writes it.  wa_out / wb_out optional (per pixel, float32).
__global__ void __launch_bounds__(256) blend_apply_kernel(const int16_t *A, int64_t a_rs, const int16_t *B, int64_t b_rs,
                                                          int rows, int cols, int ch, int method_flags, int d_row, int d_col,
                                                          const BlendPlan *__restrict__ plan, const float *__restrict__ w1, const float *__restrict__ w2,
                                                          uint8_t *out8, int64_t o_rs, int16_t *out16, int64_t o16_rs,
                                                          float *wa_out, float *wb_out)
{
    const int64_t total = (int64_t)rows * cols;
    const int method = method_flags & 0xff;
    const bool raw = (method_flags & 0x200) != 0;
    const int corner = (method == VFSMS_FUSE_FADE || method == VFSMS_FUSE_TRIG) ? plan->corner : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
        const int16_t *pa = A + r * a_rs + (int64_t)c * ch, *pb = B + r * b_rs + (int64_t)c * ch;
        double wa = 1.0, wb = 1.0;
        if (method == VFSMS_FUSE_FADE) {
            float fa = 1.f, fb = 1.f;
            if (corner) { fb = w1[r] * w2[c]; fa = 1.f - fb; }
            else if (cols <= rows) {                                       // horizontal ramp, ImageFusion.py:213-225
                const float fc = (float)cols;
                if (d_col >= 0) { fa = (float)(cols - 1 - c) / fc; fb = (float)c / fc; }
                else { fa = (float)(c + 1) / fc; fb = (float)(cols - c) / fc; }
            } else {                                                       // vertical ramp, ImageFusion.py:227-235
                const float fr = (float)rows;
                if (d_row <= 0) { fa = (float)r / fr; fb = (float)(rows - 1 - r) / fr; }
                else { fa = (float)(rows - r) / fr; fb = (float)(r + 1) / fr; }
            }
            wa = (double)fa; wb = (double)fb;
        } else if (method == VFSMS_FUSE_TRIG) {
            if (corner) {
                // getWeightsMatrix returns float32, so sin / square / 1-x stay in float32 (numpy keeps the array dtype)
                const float fb = w1[r] * w2[c];
                const float x = ((1.f - fb) * (float)M_PI) / 2.f;
                const float sv = sinf(x);
                const float fwa = sv * sv;
                wa = (double)fwa; wb = (double)(1.f - fwa);
            } else {
                double ta;
                if (cols <= rows) {                                        // ImageFusion.py:263-271 (orientation differs from fade)
                    if (d_col >= 0) ta = (double)c / (double)cols; else ta = (double)(cols - c) / (double)cols;
                } else {
                    if (d_row <= 0) ta = (double)r / (double)rows; else ta = (double)(rows - r) / (double)rows;
                }
                const double sv = sin(ta * M_PI / 2);
                wa = sv * sv; wb = 1.0 - wa;                               // ImageFusion.py:286-287
            }
        }
        if (wa_out) { wa_out[i] = (float)wa; wb_out[i] = (float)wb; }
        for (int k = 0; k < ch; k++) {
            int a = pa[k], b = pb[k];
            int res;
            if (method == VFSMS_FUSE_FADE || method == VFSMS_FUSE_TRIG) {
                if (a < 0) a = b;                                          // imageA[imageA < 0] = imageB[imageA < 0]
                double v = wa * (double)a + wb * (double)b;
                v = v < 0 ? 0 : (v > 255 ? 255 : v);
                res = (int)v;                                              // np.uint8 truncation
            } else {
                // Stitcher.py:498-504: -1 -> 0, then mutual zero fill
                if (!raw) {
                    if (a == -1) a = 0;
                    if (b == -1) b = 0;
                    if (a == 0) a = b;
                    if (b == 0) b = a;
                }
                if (method == VFSMS_FUSE_AVERAGE) res = (a + b) / 2;      // uint8((A + B) / 2): truncation of a non-negative value
                else if (method == VFSMS_FUSE_MAXIMUM) res = a > b ? a : b;
                else if (method == VFSMS_FUSE_MINIMUM) res = a < b ? a : b;
                else res = b;                                              // notFuse
            }
            if (out8) out8[r * o_rs + (int64_t)c * ch + k] = (uint8_t)res;
            if (out16) out16[r * o16_rs + (int64_t)c * ch + k] = (int16_t)(res & 255);
        }
    }
}

__global__ void fill_i32_kernel(int *p, int n, int v) { for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v; }
__global__ void fill_i16_kernel(int16_t *p, int64_t n, int16_t v)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
// tile (u8) -> int16 buffer, rectangular copy
__global__ void u8_to_i16_kernel(const uint8_t *__restrict__ src, int64_t s_rs, int16_t *dst, int64_t d_rs, int rows, int row_elems)
{
    const int64_t total = (int64_t)rows * row_elems;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / row_elems), e = (int)(i - (int64_t)r * row_elems);
        dst[r * d_rs + e] = src[r * s_rs + e];
    }
}
__global__ void canvas_to_u8_kernel(const int16_t *__restrict__ src, uint8_t *dst, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = src[i];
        dst[i] = (uint8_t)(v < 0 ? 0 : v);                                // stitchResult[stitchResult == -1] = 0 (Stitcher.py:485)
    }
}


// Core: fuse ROI (A, B int16 on device) -> out8 and/or out16 (may alias A's memory: every pixel is read before written by
// the same thread).  Asynchronous on st.
static int fuse_roi_dev(vfsms_ctx *ctx, const int16_t *A, int64_t a_rs, const int16_t *B, int64_t b_rs, int rows, int cols, int ch,
                        int method, int d_row, int d_col, uint8_t *out8, int64_t o_rs, int16_t *out16, int64_t o16_rs,
                        float *wa_out, float *wb_out, cudaStream_t st)
{
    const int force_corner = (method & 0x100) != 0;      // getWeightsMatrix parity hook: always take the corner path
    const int raw = method & 0x200;                      // direct ImageFusion.fuseByAverage / Maximum / Minimum: no zero fill
    method &= 0xff;
    BlendState *bs = bstate(ctx);
    int rc;
    if ((rc = bs->plan.reserve(sizeof(BlendPlan)))) return rc;
    if ((rc = bs->col_top.reserve((size_t)cols * 4))) return rc;
    if ((rc = bs->col_bot.reserve((size_t)cols * 4))) return rc;
    if ((rc = bs->w1.reserve((size_t)rows * 4))) return rc;
    if ((rc = bs->w2.reserve((size_t)cols * 4))) return rc;
    StageTimer t(ctx, st, VFSMS_STAGE_BLEND);
    const int64_t n = (int64_t)rows * cols;
    if (method == VFSMS_FUSE_FADE || method == VFSMS_FUSE_TRIG) {
        CUDA_TRY(cudaMemsetAsync(bs->plan.p, 0, sizeof(BlendPlan), st));
        fill_i32_kernel<<<grid_for(ctx, cols), 256, 0, st>>>(bs->col_top.as<int>(), cols, rows);     // "none" = rows
        LAUNCH_CHECK(ctx);
        CUDA_TRY(cudaMemsetAsync(bs->col_bot.p, 0xff, (size_t)cols * 4, st));                        // "none" = -1
        blend_stats_kernel<<<dim3(ceil_div(cols, 256), ceil_div(rows, STATS_ROWS)), 256, 0, st>>>(A, a_rs, rows, cols, ch, bs->plan.as<BlendPlan>(),
                                                                                                     bs->col_top.as<int>(), bs->col_bot.as<int>());
        LAUNCH_CHECK(ctx);
        blend_plan_kernel<<<1, 256, 0, st>>>(A, a_rs, rows, cols, ch, bs->plan.as<BlendPlan>(), bs->col_top.as<int>(), bs->col_bot.as<int>(),
                                             bs->w1.as<float>(), bs->w2.as<float>(), force_corner);
        LAUNCH_CHECK(ctx);
    }
    blend_apply_kernel<<<grid_for(ctx, n), 256, 0, st>>>(A, a_rs, B, b_rs, rows, cols, ch, method | raw, d_row, d_col, bs->plan.as<BlendPlan>(),
                                                         bs->w1.as<float>(), bs->w2.as<float>(), out8, o_rs, out16, o16_rs, wa_out, wb_out);
    LAUNCH_CHECK(ctx);
    return 0;
}

// ---------------------------------------------------------------- multi-band blending (ImageFusion.py:296-367)
// BlendArbitrary2(A, B, 4): Gaussian pyramids (cv2.pyrDown, float64), Laplacian levels through cv2.pyrUp + resize(INTER_CUBIC)
// to the finer level's size, constant 0.5 / 0.5 level weights, reconstruction, np.uint8 truncation.  Border rules pinned
// against cv2 here: pyrDown BORDER_REFLECT_101; pyrUp reflects at the left / top and clamps at the right / bottom.
__device__ __forceinline__ int refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * n - 2 - i; }
    return i;
}

__global__ void __launch_bounds__(256) pyr_down_kernel(const double *__restrict__ src, int h, int w, double *dst)
{
    const int H = (h + 1) / 2, W = (w + 1) / 2;
    const double k[5] = { 1, 4, 6, 4, 1 };
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i - y * W;
        double acc = 0;
        for (int ty = 0; ty < 5; ty++) {
            const double *row = src + (size_t)refl101(2 * y + ty - 2, h) * w;
            double r = 0;
            for (int tx = 0; tx < 5; tx++) r += k[tx] * row[refl101(2 * x + tx - 2, w)];
            acc += k[ty] * r;
        }
        dst[i] = acc / 256.0;
    }
}

__device__ __forceinline__ int up_idx(int i, int n) { return i < 0 ? refl101(i, n) : (i > n - 1 ? n - 1 : i); }

// cv2.pyrUp: (h, w) -> (2h, 2w)
__global__ void __launch_bounds__(256) pyr_up_kernel(const double *__restrict__ src, int h, int w, double *dst)
{
    const int H = 2 * h, W = 2 * w;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        const int Y = i / W, X = i - Y * W;
        const int y = Y >> 1, x = X >> 1;
        // horizontal pass value at source row r, output column X
        auto hval = [&](int r) -> double {
            const double *row = src + (size_t)r * w;
            if (X & 1) return 4.0 * (row[x] + row[up_idx(x + 1, w)]);
            return row[up_idx(x - 1, w)] + 6.0 * row[x] + row[up_idx(x + 1, w)];
        };
        double v;
        if (Y & 1) v = 4.0 * (hval(y) + hval(up_idx(y + 1, h)));
        else v = hval(up_idx(y - 1, h)) + 6.0 * hval(y) + hval(up_idx(y + 1, h));
        dst[i] = v / 64.0;
    }
}

__device__ __forceinline__ void cubic_coeffs(float t, float c[4])
{
    const float A = -0.75f;
    c[0] = ((A * (t + 1) - 5 * A) * (t + 1) + 8 * A) * (t + 1) - 4 * A;
    c[1] = ((A + 2) * t - (A + 3)) * t * t + 1;
    c[2] = ((A + 2) * (1 - t) - (A + 3)) * (1 - t) * (1 - t) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}

// cv2.resize(src, (W, H), INTER_CUBIC) for float64 (identity when the size is unchanged is handled by the caller)
__global__ void __launch_bounds__(256) resize_cubic_kernel(const double *__restrict__ src, int h, int w, double *dst, int H, int W)
{
    const double sx = (double)w / W, sy = (double)h / H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        const int Y = i / W, X = i - Y * W;
        float fx = (float)((X + 0.5) * sx - 0.5), fy = (float)((Y + 0.5) * sy - 0.5);
        const int x0 = (int)floorf(fx), y0 = (int)floorf(fy);
        float cx[4], cy[4];
        cubic_coeffs(fx - x0, cx); cubic_coeffs(fy - y0, cy);
        double acc = 0;
        for (int ky = 0; ky < 4; ky++) {
            const double *row = src + (size_t)min(max(y0 - 1 + ky, 0), h - 1) * w;
            double r = 0;
            for (int kx = 0; kx < 4; kx++) r += (double)cx[kx] * row[min(max(x0 - 1 + kx, 0), w - 1)];
            acc += (double)cy[ky] * r;
        }
        dst[i] = acc;
    }
}

// dst = alpha * a + beta * b   (b may be null)
__global__ void __launch_bounds__(256) axpby_kernel(double *dst, const double *a, double alpha, const double *b, double beta, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = alpha * a[i] + (b ? beta * b[i] : 0.0);
}

// Stitcher.fuseImage pre-processing (-1 -> 0, mutual zero fill, Stitcher.py:498-504) + conversion to float64
__global__ void __launch_bounds__(256) multiband_prepare_kernel(const int16_t *__restrict__ A, int64_t a_rs, const int16_t *__restrict__ B, int64_t b_rs,
                                                                int rows, int cols, double *da, double *db)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += gridDim.x * blockDim.x) {
        const int r = i / cols, c = i - r * cols;
        int a = A[r * a_rs + c], b = B[r * b_rs + c];
        if (a == -1) a = 0;
        if (b == -1) b = 0;
        if (a == 0) a = b;
        if (b == 0) b = a;
        da[i] = a; db[i] = b;
    }
}

__global__ void __launch_bounds__(256) multiband_store_kernel(const double *__restrict__ src, int rows, int cols, uint8_t *out8, int64_t o_rs,
                                                              int16_t *out16, int64_t o16_rs)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += gridDim.x * blockDim.x) {
        const int r = i / cols, c = i - r * cols;
        const uint8_t v = (uint8_t)(long long)src[i];          // np.uint8(float64): truncation toward zero, wraps like the C cast
        if (out8) out8[r * o_rs + c] = v;
        if (out16) out16[r * o16_rs + c] = v;
    }
}

static int multiband_dev(vfsms_ctx *ctx, const int16_t *A, int64_t a_rs, const int16_t *B, int64_t b_rs, int rows, int cols,
                         uint8_t *out8, int64_t o_rs, int16_t *out16, int64_t o16_rs, cudaStream_t st)
{
    BlendState *bs = bstate(ctx);
    const int LEVELS = 4;
    int hs[LEVELS], wsz[LEVELS];
    size_t off[LEVELS], tot = 0;
    hs[0] = rows; wsz[0] = cols;
    for (int l = 0; l < LEVELS; l++) {
        if (l) { hs[l] = (hs[l - 1] + 1) / 2; wsz[l] = (wsz[l - 1] + 1) / 2; }
        off[l] = tot; tot += (size_t)hs[l] * wsz[l];
    }
    const size_t big = (size_t)rows * cols * 4 + 64;        // room for a pyrUp of the finest level
    // layout: gpA | gpB | LC | tmp1 | tmp2
    int rc;
    if ((rc = bs->pyr.reserve((tot * 3 + big * 2) * 8))) return rc;
    double *gA = bs->pyr.as<double>(), *gB = gA + tot, *LC = gB + tot, *t1 = LC + tot, *t2 = t1 + big;
    StageTimer t(ctx, st, VFSMS_STAGE_BLEND);
    auto G = [&](int n) { return grid_for(ctx, n); };
    multiband_prepare_kernel<<<G(rows * cols), 256, 0, st>>>(A, a_rs, B, b_rs, rows, cols, gA, gB);
    LAUNCH_CHECK(ctx);
    for (int l = 1; l < LEVELS; l++) {
        pyr_down_kernel<<<G(hs[l] * wsz[l]), 256, 0, st>>>(gA + off[l - 1], hs[l - 1], wsz[l - 1], gA + off[l]); LAUNCH_CHECK(ctx);
        pyr_down_kernel<<<G(hs[l] * wsz[l]), 256, 0, st>>>(gB + off[l - 1], hs[l - 1], wsz[l - 1], gB + off[l]); LAUNCH_CHECK(ctx);
    }
    // expand(src at level l+1) -> size of level l, into `dst`
    auto expand = [&](const double *src, int l, double *dst) -> int {
        const int h = hs[l + 1], w = wsz[l + 1];
        if (2 * h == hs[l] && 2 * w == wsz[l]) { pyr_up_kernel<<<G(4 * h * w), 256, 0, st>>>(src, h, w, dst); LAUNCH_CHECK(ctx); }
        else {
            pyr_up_kernel<<<G(4 * h * w), 256, 0, st>>>(src, h, w, t2); LAUNCH_CHECK(ctx);
            resize_cubic_kernel<<<G(hs[l] * wsz[l]), 256, 0, st>>>(t2, 2 * h, 2 * w, dst, hs[l], wsz[l]); LAUNCH_CHECK(ctx);
        }
        return 0;
    };
    // LC[coarsest] = 0.5 gA[3] + 0.5 gB[3];  LC[l] = 0.5 (gA[l] - E(gA[l+1])) + 0.5 (gB[l] - E(gB[l+1]))
    axpby_kernel<<<G(hs[LEVELS - 1] * wsz[LEVELS - 1]), 256, 0, st>>>(LC + off[LEVELS - 1], gA + off[LEVELS - 1], 0.5, gB + off[LEVELS - 1], 0.5, hs[LEVELS - 1] * wsz[LEVELS - 1]);
    LAUNCH_CHECK(ctx);
    for (int l = LEVELS - 2; l >= 0; l--) {
        const int n = hs[l] * wsz[l];
        if ((rc = expand(gA + off[l + 1], l, t1))) return rc;
        axpby_kernel<<<G(n), 256, 0, st>>>(t1, gA + off[l], 1.0, t1, -1.0, n); LAUNCH_CHECK(ctx);             // LA[l]
        axpby_kernel<<<G(n), 256, 0, st>>>(LC + off[l], t1, 0.5, nullptr, 0.0, n); LAUNCH_CHECK(ctx);
        if ((rc = expand(gB + off[l + 1], l, t1))) return rc;
        axpby_kernel<<<G(n), 256, 0, st>>>(t1, gB + off[l], 1.0, t1, -1.0, n); LAUNCH_CHECK(ctx);             // LB[l]
        axpby_kernel<<<G(n), 256, 0, st>>>(LC + off[l], LC + off[l], 1.0, t1, 0.5, n); LAUNCH_CHECK(ctx);
    }
    // reconstruct: out = LC[3]; out = E(out) + LC[l]
    double *cur = gA;      // reuse gA storage level by level
    CUDA_TRY(cudaMemcpyAsync(cur + off[LEVELS - 1], LC + off[LEVELS - 1], (size_t)hs[LEVELS - 1] * wsz[LEVELS - 1] * 8, cudaMemcpyDeviceToDevice, st));
    for (int l = LEVELS - 2; l >= 0; l--) {
        const int n = hs[l] * wsz[l];
        if ((rc = expand(cur + off[l + 1], l, t1))) return rc;
        axpby_kernel<<<G(n), 256, 0, st>>>(cur + off[l], t1, 1.0, LC + off[l], 1.0, n); LAUNCH_CHECK(ctx);
    }
    multiband_store_kernel<<<G(rows * cols), 256, 0, st>>>(cur, rows, cols, out8, o_rs, out16, o16_rs);
    LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" {

int vfsms_fuse_roi_host(vfsms_ctx *ctx, const int16_t *a, const int16_t *b, int rows, int cols, int channels, int method,
                        int d_row, int d_col, uint8_t *out, float *wa_out, float *wb_out)
{
    if (!ctx || !a || !b || !out || rows < 1 || cols < 1 || (channels != 1 && channels != 3)) { vfsms_set_error("fuse_roi: bad arguments"); return VFSMS_E_ARG; }
    if ((method & 0xff) < VFSMS_FUSE_NONE || (method & 0xff) > VFSMS_FUSE_MULTIBAND) { vfsms_set_error("fuse_roi: unknown method %d", method); return VFSMS_E_ARG; }
    if ((method & 0xff) == VFSMS_FUSE_MULTIBAND && channels != 1) {
        vfsms_set_error("fuse_roi: multi-band blending is gray only (Stitcher.py:520)"); return VFSMS_E_UNSUPPORTED;
    }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    BlendState *bs = bstate(ctx);
    const size_t n = (size_t)rows * cols * channels;
    int rc;
    if ((rc = bs->a16.reserve(n * 2))) return rc;
    if ((rc = bs->b16.reserve(n * 2))) return rc;
    if ((rc = bs->out8.reserve(n))) return rc;
    float *dwa = nullptr, *dwb = nullptr;
    if (wa_out && wb_out) {
        if ((rc = bs->wa_out.reserve((size_t)rows * cols * 4))) return rc;
        if ((rc = bs->wb_out.reserve((size_t)rows * cols * 4))) return rc;
        dwa = bs->wa_out.as<float>(); dwb = bs->wb_out.as<float>();
    }
    CUDA_TRY(cudaMemcpyAsync(bs->a16.p, a, n * 2, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(bs->b16.p, b, n * 2, cudaMemcpyHostToDevice, st));
    const int64_t rs = (int64_t)cols * channels;
    if ((method & 0xff) == VFSMS_FUSE_MULTIBAND) {
        if ((rc = multiband_dev(ctx, bs->a16.as<int16_t>(), rs, bs->b16.as<int16_t>(), rs, rows, cols, bs->out8.as<uint8_t>(), rs, nullptr, 0, st))) return rc;
    } else if ((rc = fuse_roi_dev(ctx, bs->a16.as<int16_t>(), rs, bs->b16.as<int16_t>(), rs, rows, cols, channels, method, d_row, d_col,
                                  bs->out8.as<uint8_t>(), rs, nullptr, 0, dwa, dwb, st))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, bs->out8.p, n, cudaMemcpyDeviceToHost, st));
    if (dwa) {
        CUDA_TRY(cudaMemcpyAsync(wa_out, dwa, (size_t)rows * cols * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(wb_out, dwb, (size_t)rows * cols * 4, cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

// tiles_host: n_tiles x (tile_rows x tile_cols x channels) in host memory, or nullptr with tiles_dev pointing at the same layout in HBM
// Band extension (vfsms_mosaic_band_host): `halo_in` pre-fills the canvas rectangle halo_in_rect (what earlier bands left
// there, -1 = empty), `fuse_first` blends tile 0 too (it is not the first tile of the sequence), `halo_out` reads the
// rectangle halo_out_rect back as int16 (holes kept) after the last tile.
struct MosaicBand {
    int fuse_first = 0;
    const int16_t *halo_in = nullptr;  const int32_t *halo_in_rect = nullptr;     // r0, c0, rows, cols
    int16_t *halo_out = nullptr;       const int32_t *halo_out_rect = nullptr;
};

static bool rect_inside(const int32_t *r, int canvas_rows, int canvas_cols)
{
    return r && r[0] >= 0 && r[1] >= 0 && r[2] >= 1 && r[3] >= 1 && r[0] + r[2] <= canvas_rows && r[1] + r[3] <= canvas_cols;
}

static int mosaic_run(vfsms_ctx *ctx, const uint8_t *tiles_host, const uint8_t *tiles_dev, int n_tiles, int tile_rows, int tile_cols, int channels,
                      const int32_t *tile_origin, const int32_t *roi_rect, const int32_t *pair_offset, int method,
                      int canvas_rows, int canvas_cols, uint8_t *canvas_out, const MosaicBand *band = nullptr)
{
    const uint8_t *tiles = tiles_host ? tiles_host : tiles_dev;
    if (!ctx || !tiles || !tile_origin || !roi_rect || !pair_offset || !canvas_out || n_tiles < 1 || (channels != 1 && channels != 3)) {
        vfsms_set_error("mosaic: bad arguments"); return VFSMS_E_ARG;
    }
    if (method < VFSMS_FUSE_NONE || method > VFSMS_FUSE_TRIG) { vfsms_set_error("mosaic: method %d not available", method); return VFSMS_E_UNSUPPORTED; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    BlendState *bs = bstate(ctx);
    const int64_t crs = (int64_t)canvas_cols * channels;
    const int64_t cn = (int64_t)canvas_rows * crs;
    const size_t tile_bytes = (size_t)tile_rows * tile_cols * channels;
    int rc;
    if ((rc = bs->canvas.reserve((size_t)cn * 2))) return rc;
    if ((rc = bs->tile.reserve(tile_bytes * 2))) return rc;          // int16 copy of the current tile
    if ((rc = bs->tiles_all.reserve(tile_bytes))) return rc;          // u8 staging of the current tile
    if ((rc = bs->out8.reserve((size_t)cn))) return rc;
    int16_t *canvas = bs->canvas.as<int16_t>();
    fill_i16_kernel<<<grid_for(ctx, cn), 256, 0, st>>>(canvas, cn, (int16_t)-1);
    LAUNCH_CHECK(ctx);
    if (band && band->halo_in) {
        const int32_t *r = band->halo_in_rect;
        if (!rect_inside(r, canvas_rows, canvas_cols)) { vfsms_set_error("mosaic: halo_in rectangle outside the canvas"); return VFSMS_E_ARG; }
        const size_t wbytes = (size_t)r[3] * channels * 2;
        CUDA_TRY(cudaMemcpy2DAsync(canvas + r[0] * crs + (int64_t)r[1] * channels, (size_t)crs * 2, band->halo_in, wbytes, wbytes, r[2],
                                   cudaMemcpyHostToDevice, st));
    }
    if (band && band->halo_out && !rect_inside(band->halo_out_rect, canvas_rows, canvas_cols)) {
        vfsms_set_error("mosaic: halo_out rectangle outside the canvas"); return VFSMS_E_ARG;
    }
    const bool fuse_first = band && band->fuse_first;
    const int64_t trs = (int64_t)tile_cols * channels;
    for (int i = 0; i < n_tiles; i++) {
        const int r0 = tile_origin[2 * i], c0 = tile_origin[2 * i + 1];
        if (r0 < 0 || c0 < 0 || r0 + tile_rows > canvas_rows || c0 + tile_cols > canvas_cols) { vfsms_set_error("mosaic: tile %d outside the canvas", i); return VFSMS_E_ARG; }
        const uint8_t *t8 = tiles_dev ? tiles_dev + (size_t)i * tile_bytes : bs->tiles_all.as<uint8_t>();
        if (tiles_host) CUDA_TRY(cudaMemcpyAsync(bs->tiles_all.p, tiles_host + (size_t)i * tile_bytes, tile_bytes, cudaMemcpyHostToDevice, st));
        int16_t *t16 = bs->tile.as<int16_t>();
        u8_to_i16_kernel<<<grid_for(ctx, (int64_t)tile_rows * trs), 256, 0, st>>>(t8, trs, t16, trs, tile_rows, (int)trs);
        LAUNCH_CHECK(ctx);
        int16_t *dst = canvas + r0 * crs + (int64_t)c0 * channels;
        const int rr0 = roi_rect[4 * i], rc0 = roi_rect[4 * i + 1], rr1 = roi_rect[4 * i + 2], rc1 = roi_rect[4 * i + 3];
        const bool fuse = (i > 0 || fuse_first) && method != VFSMS_FUSE_NONE && rr1 > rr0 && rc1 > rc0;
        if (fuse) {
            // ROI: A = canvas before the paste, B = the tile there; result written over the canvas ROI (Stitcher.py:466-483)
            if (rr0 < r0 || rc0 < c0 || rr1 > r0 + tile_rows || rc1 > c0 + tile_cols) { vfsms_set_error("mosaic: ROI %d outside its tile", i); return VFSMS_E_ARG; }
            int16_t *Aroi = canvas + rr0 * crs + (int64_t)rc0 * channels;
            const int16_t *Broi = t16 + (rr0 - r0) * trs + (int64_t)(rc0 - c0) * channels;
            // the fused values go to a scratch copy of the tile ROI (so that the plain paste below finishes the job)
            if ((rc = fuse_roi_dev(ctx, Aroi, crs, Broi, trs, rr1 - rr0, rc1 - rc0, channels, method, pair_offset[2 * i], pair_offset[2 * i + 1],
                                   nullptr, 0, (int16_t *)Broi, trs, nullptr, nullptr, st))) return rc;
        }
        CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)crs * 2, t16, (size_t)trs * 2, (size_t)trs * 2, tile_rows, cudaMemcpyDeviceToDevice, st));
    }
    if (band && band->halo_out) {
        const int32_t *r = band->halo_out_rect;
        const size_t wbytes = (size_t)r[3] * channels * 2;
        CUDA_TRY(cudaMemcpy2DAsync(band->halo_out, wbytes, canvas + r[0] * crs + (int64_t)r[1] * channels, (size_t)crs * 2, wbytes, r[2],
                                   cudaMemcpyDeviceToHost, st));
    }
    canvas_to_u8_kernel<<<grid_for(ctx, cn), 256, 0, st>>>(canvas, bs->out8.as<uint8_t>(), cn);
    LAUNCH_CHECK(ctx);
    CUDA_TRY(cudaMemcpyAsync(canvas_out, bs->out8.p, (size_t)cn, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int vfsms_mosaic_host(vfsms_ctx *ctx, const uint8_t *tiles, int n_tiles, int tile_rows, int tile_cols, int channels,
                      const int32_t *tile_origin, const int32_t *roi_rect, const int32_t *pair_offset, int method,
                      int canvas_rows, int canvas_cols, uint8_t *canvas_out)
{
    if (!tiles) { vfsms_set_error("mosaic: bad arguments"); return VFSMS_E_ARG; }
    return mosaic_run(ctx, tiles, nullptr, n_tiles, tile_rows, tile_cols, channels, tile_origin, roi_rect, pair_offset, method, canvas_rows,
                      canvas_cols, canvas_out);
}

int vfsms_mosaic_band_host(vfsms_ctx *ctx, const uint8_t *tiles, int n_tiles, int tile_rows, int tile_cols, int channels,
                           const int32_t *tile_origin, const int32_t *roi_rect, const int32_t *pair_offset, int method,
                           int canvas_rows, int canvas_cols, int fuse_first, const int16_t *halo_in, const int32_t *halo_in_rect,
                           int16_t *halo_out, const int32_t *halo_out_rect, uint8_t *canvas_out)
{
    if (!tiles || (halo_in && !halo_in_rect) || (halo_out && !halo_out_rect)) { vfsms_set_error("mosaic_band: bad arguments"); return VFSMS_E_ARG; }
    MosaicBand band;
    band.fuse_first = fuse_first;
    band.halo_in = halo_in; band.halo_in_rect = halo_in_rect;
    band.halo_out = halo_out; band.halo_out_rect = halo_out_rect;
    return mosaic_run(ctx, tiles, nullptr, n_tiles, tile_rows, tile_cols, channels, tile_origin, roi_rect, pair_offset, method, canvas_rows,
                      canvas_cols, canvas_out, &band);
}

int vfsms_tiles_mosaic(vfsms_ctx *ctx, int first, int n_tiles, const int32_t *tile_origin, const int32_t *roi_rect, const int32_t *pair_offset,
                       int method, int canvas_rows, int canvas_cols, uint8_t *canvas_out)
{
    if (!ctx || !ctx->tiles.p || first < 0 || n_tiles < 1 || first + n_tiles > ctx->tiles_n) {
        vfsms_set_error("tiles_mosaic: tiles [%d, %d) outside the reserved stack", first, first + n_tiles); return VFSMS_E_ARG;
    }
    const size_t img = (size_t)ctx->tiles_rows * ctx->tiles_cols;
    return mosaic_run(ctx, nullptr, ctx->tiles.as<uint8_t>() + first * img, n_tiles, ctx->tiles_rows, ctx->tiles_cols, 1, tile_origin, roi_rect,
                      pair_offset, method, canvas_rows, canvas_cols, canvas_out);
}

int vfsms_tiles_mosaic_bgr(vfsms_ctx *ctx, int first, int n_tiles, const int32_t *tile_origin, const int32_t *roi_rect, const int32_t *pair_offset,
                           int method, int canvas_rows, int canvas_cols, uint8_t *canvas_out)
{
    if (!ctx || !ctx->tiles_bgr.p || first < 0 || n_tiles < 1 || first + n_tiles > ctx->tiles_n) {
        vfsms_set_error("tiles_mosaic_bgr: tiles [%d, %d) outside the colour stack", first, first + n_tiles); return VFSMS_E_ARG;
    }
    for (int k = first; k < first + n_tiles; k++)
        if (!ctx->tiles_has_bgr[k]) { vfsms_set_error("tiles_mosaic_bgr: slot %d holds no colour tile", k); return VFSMS_E_ARG; }
    const size_t img = (size_t)ctx->tiles_rows * ctx->tiles_cols * 3;
    return mosaic_run(ctx, nullptr, ctx->tiles_bgr.as<uint8_t>() + first * img, n_tiles, ctx->tiles_rows, ctx->tiles_cols, 3, tile_origin, roi_rect,
                      pair_offset, method, canvas_rows, canvas_cols, canvas_out);
}

}  // extern "C"
