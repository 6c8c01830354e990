// surf_describe.cuh -- SURF orientation + descriptor, one WARP per keypoint (included by surf.cu; shares its constant tables).
//
// Replaces the per-keypoint loop of OpenCV's SURFInvoker (opencv_contrib 3.3.1 modules/xfeatures2d/src/surf.cpp, reached from
// ImageUtility.py:258-264 / appendix/myGpuFeatures.cpp:77-83) as restated in oracle/surf_oracle.c describe_one():
// 113 Haar samples -> 72 sliding windows -> direction; rotated 20s x 20s window sampled bilinearly, each sample rounded to u8;
// INTER_AREA resize to 21 x 21; Gaussian-weighted gradients; 4 x 4 x (4 | 8) bins; L2 normalisation.
// Results are bit-identical to the oracle: every rounding of the scalar CPU code is reproduced (this file is compiled with
// -fmad=false), only the schedule differs.
//
// Two kernels:
//   describe_fixed_kernel      the product path.  Window positions in 32.32 fixed point (exact, see below), rows sampled in CHUNKS
//                              of up to 32 warp-rounds whose texture gathers are issued four at a time, INTER_AREA folded row by row.
//                              Small batches add a cooperative pass: the eight warps of a CTA share each window of >= 320 px by
//                              output rows (describe_giants_kernel lists them).  Options describe = 2 / 3 swap the window pixel
//                              for a tolerance-mode sampler (NOT bit-exact; sample_rounds_tol).
//   describe_reference_kernel  row-at-a-time sampler in double precision straight from the u8 image (no texture, no exactness
//                              preconditions).  Describes what the fixed kernel hands over (a work list: degenerate directions,
//                              sub-2^-9 row starts), everything when vfsms_set_option("describe", 0), and upright keypoints.
//
// Why fixed point is exact.  The CPU walks a window row as  pixel_x = (double)start_x + j * (double)cos_dir  (accumulated, every
// partial sum exact in double) with start_x, cos_dir floats; start_x itself is the float chain start_x += sin_dir per row, which
// both kernels advance with the same float additions.  When |cos_dir| is 0 or in [2^-9, 1) it is a multiple of 2^-32 below 1, and
// when |start_x| is 0 or >= 2^-9 so is the row start: then X = pixel_x * 2^32 is an exact 64-bit integer for every sample,
// floor(pixel_x) is its high word, and the CPU's a = (float)(pixel_x - ix) -- ONE rounding of an exact difference -- equals
// RN((float)low word) * 2^-32.  The power-of-two scale commutes with every later rounding of the bilinear expression, so it is folded
// into the texture: the float copy of the image holds p * 2^-64 and the kernel evaluates
//     p00' * (T - A) * (T - B) + p01' * A * (T - B) + p10' * (T - A) * B + p11' * A * B,   T = 2^32, A = RN(lo(X)), B = RN(lo(Y))
// in the CPU's association order: the same float, no scaling instruction.
#pragma once

#define DESC_THREADS 256
#define SURF_MAX_DESC_CHUNKS 64   // launches per batch (one stacked texture each)
#define WK_MAX_WIN 768            // windows up to this size are described by one warp (n_octaves <= 4 never exceeds 739)
#define WK_WARPS 8
#define DESC_BUF 1024             // floats of row storage per warp: one chunk of sampled rows (fixed kernel), one row (reference kernel)
#define DESC_SLOTS 32             // warp-rounds per chunk
#define DESC_FIXED_MINB 3          // CTAs per SM the fixed kernel's register budget is cut for
#define DESC_FIXED_U 4             // gathers in flight per lane
#define DESC_COOP_SPLIT 320        // windows at least this wide are shared by the warps of a CTA ...
#define DESC_COOP_MAX_BATCH 24     // ... in batches of at most this many images (larger batches fill the GPU without it)
#define DESC_COOP_SPLIT_TINY 240   // threshold for batches of up to four images (measured: 1.53 vs 1.73 ms for a single pair)

// bilinear / clamped sample of the rotated window, rounded to u8 exactly like the CPU loop (reference sampler)
__device__ __forceinline__ int window_pixel(const uint8_t *__restrict__ img, int stride, int ncols1, int nrows1,
                                            double pixel_x, double pixel_y)
{
    const int ix = __double2int_rd(pixel_x), iy = __double2int_rd(pixel_y);
    if ((unsigned)ix < (unsigned)ncols1 && (unsigned)iy < (unsigned)nrows1) {
        const float a = (float)(pixel_x - ix), bq = (float)(pixel_y - iy);
        const uint8_t *p = img + (size_t)iy * stride + ix;
        const float p00 = p[0], p01 = p[1], p10 = p[stride], p11 = p[stride + 1];
        const float v = p00 * (1.f - a) * (1.f - bq) + p01 * a * (1.f - bq) + p10 * (1.f - a) * bq + p11 * a * bq;
        return __float2int_rn(v) & 255;
    }
    int x = __double2int_rn(pixel_x), y = __double2int_rn(pixel_y);
    x = min(max(x, 0), ncols1); y = min(max(y, 0), nrows1);
    return img[(size_t)y * stride + x];
}

__device__ __forceinline__ int round_half_even_u8(float v)      // cvRound for 0 <= v < 2^22 without a conversion instruction
{
    return __float_as_int(v + 12582912.0f) - 0x4B400000;
}

// float copy of the batch's images for the texture path (pitch in floats), scaled by a power of two (2^-64 for the fixed kernel)
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t *base_a, const uint8_t *base_b, int split, int64_t img_stride,
                                                        int rows, int cols, int stride, float *dst, int pitch_f, float scale,
                                                        const int64_t *__restrict__ img_off)
{
    const int b = blockIdx.y;
    const uint8_t *img = image_ptr(base_a, base_b, split, b, img_stride, img_off);
    float *D = dst + (size_t)b * rows * pitch_f;
    const int total = rows * cols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / cols, x = i - y * cols;
        D[(size_t)y * pitch_f + x] = (float)img[(size_t)y * stride + x] * scale;
    }
}

struct __align__(16) DescScratch {
    float buf[DESC_BUF];         // orientation: X[0..127] Y[128..255] A[256..383] (int) ; window phase: sampled rows as exact floats ;
                                 // descriptor: DX[0..399] DY[400..799]
    uint4 slot[DESC_SLOTS];      // fixed kernel: 32.32 start (x lo, x hi, y lo, y hi) of each warp-round of the current chunk
    float vec[168];              // descriptor bins (128) ; before that: INTER_AREA column table (21 x 8 words)
    uint8_t patch[448];
};

// ---- dominant orientation (OpenCV's 72 windows of 60 degrees over the 113 Gaussian-weighted Haar responses); warp-collective.
// Window i = 5w holds sample j when |round(angle_j) - i| < 30 or > 330, i.e. when w lies in the circular range of 11 (12 when the
// angle is not a multiple of 5) windows starting at angle / 5 - 5 (mod 72).  Each sample's owner turns that range into three 24-bit
// membership words (w = 24 g + l); lane l < 24 then walks the samples ONCE, in order, adding into its windows l, l + 24, l + 48
// under one bit test each: the CPU's summation order per window with a third of the compares.
__device__ __forceinline__ float orient_keypoint(DescScratch &S, int lane, const int32_t *__restrict__ I, int W, int srows, int scols,
                                                 float cx, float cy, float s, int gws)
{
    float2 *sXY = (float2 *)S.buf;                    // [128]
    uint4 *sM = (uint4 *)(S.buf + 256);               // [128]
    const unsigned lt_mask = (1u << lane) - 1;
    const int h2 = __float2int_rn(((float)gws / 4) * 2);
    const int h4 = __float2int_rn(((float)gws / 4) * 4);
    const float wgt = 1.f / ((float)(h2) * (float)(h4));
    const float half = (float)(gws - 1) / 2;
    int nangle = 0;
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
        const int kk = r * 32 + lane;
        bool have = false; float vX = 0, vY = 0;
        if (kk < ORI_SAMPLES) {
            const int x = __float2int_rn(cx + c_apt_x[kk] * s - half);
            const int y = __float2int_rn(cy + c_apt_y[kk] * s - half);
            if (!(y < 0 || y >= srows - gws || x < 0 || x >= scols - gws)) {
                const int32_t *o = I + (size_t)y * W + x;
                // 3x3 grid of integral corners shared by the four half boxes
                const int a00 = __ldg(o), a01 = __ldg(o + h2), a02 = __ldg(o + h4);
                const int32_t *o1 = o + (size_t)h2 * W, *o2 = o + (size_t)h4 * W;
                const int a10 = __ldg(o1), a12 = __ldg(o1 + h4);
                const int a20 = __ldg(o2), a21 = __ldg(o2 + h2), a22 = __ldg(o2 + h4);
                const int bl = a00 + a21 - a20 - a01;          // x in [0,h2), y in [0,h4)
                const int br = a01 + a22 - a21 - a02;          // x in [h2,h4)
                const int bt = a00 + a12 - a10 - a02;          // y in [0,h2)
                const int bb = a10 + a22 - a20 - a12;          // y in [h2,h4)
                double d = 0; d += (double)((float)bl * (-wgt)); d += (double)((float)br * wgt);
                const float vx = (float)d;
                d = 0; d += (double)((float)bt * wgt); d += (double)((float)bb * (-wgt));
                const float vy = (float)d;
                vX = vx * c_aptw[kk]; vY = vy * c_aptw[kk];
                have = true;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, have);
        if (have) {
            const int pos = nangle + __popc(bal & lt_mask);
            int A = __float2int_rn(fast_atan2_deg(vY, vX));       // 0 .. 360
            if (A >= 360) A -= 360;
            const int a = A / 5, bq = A - 5 * a;
            int lo = a - 5; if (lo < 0) lo += 72;
            const unsigned M = (1u << (bq > 0 ? 12 : 11)) - 1;
            unsigned wd[3];
#pragma unroll
            for (int g = 0; g < 3; g++) {
                int sh = lo - 24 * g;
                if (sh >= 36) sh -= 72; else if (sh < -36) sh += 72;
                wd[g] = sh >= 0 ? ((M << min(sh, 24)) & 0xFFFFFFu) : (M >> min(-sh, 31));
            }
            sXY[pos] = make_float2(vX, vY);
            sM[pos] = make_uint4(wd[0], wd[1], wd[2], 0u);
        }
        nangle += __popc(bal);
    }
    __syncwarp();
    const unsigned lbit = lane < 24 ? 1u << lane : 0u;
    float s0x = 0, s0y = 0, s1x = 0, s1y = 0, s2x = 0, s2y = 0;
#pragma unroll 4
    for (int j = 0; j < nangle; j++) {
        const uint4 m = sM[j];
        const float2 v = sXY[j];
        if (m.x & lbit) { s0x += v.x; s0y += v.y; }
        if (m.y & lbit) { s1x += v.x; s1y += v.y; }
        if (m.z & lbit) { s2x += v.x; s2y += v.y; }
    }
    float bmod = 0, bx = 0, by = 0; int bw = 1 << 30;
    if (lane < 24) {                                  // this lane's windows in increasing order: the first maximum wins
        float m = s0x * s0x + s0y * s0y;
        if (m > bmod) { bmod = m; bx = s0x; by = s0y; bw = lane; }
        m = s1x * s1x + s1y * s1y;
        if (m > bmod) { bmod = m; bx = s1x; by = s1y; bw = lane + 24; }
        m = s2x * s2x + s2y * s2y;
        if (m > bmod) { bmod = m; bx = s2x; by = s2y; bw = lane + 48; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, bmod, o), ox = __shfl_xor_sync(0xffffffffu, bx, o),
                    oy = __shfl_xor_sync(0xffffffffu, by, o);
        const int ow = __shfl_xor_sync(0xffffffffu, bw, o);
        if (om > bmod || (om == bmod && ow < bw)) { bmod = om; bx = ox; by = oy; bw = ow; }
    }
    __syncwarp();
    return fast_atan2_deg(-by, bx);
}

// ---- S.patch (21 x 21 u8) -> descriptor; warp-collective
__device__ __forceinline__ void patch_to_descriptor(DescScratch &S, int lane, int extended, float *__restrict__ dst)
{
    constexpr int PD = PATCH_SZ + 1;
    const int dsize = extended ? 128 : 64;
    float *sDX = S.buf, *sDY = S.buf + 400;
    for (int p = lane; p < PATCH_SZ * PATCH_SZ; p += 32) {
        const int i = p / PATCH_SZ, j = p - i * PATCH_SZ;
        const float dw = c_DW[p];
        const int p00 = S.patch[i * PD + j], p01 = S.patch[i * PD + j + 1];
        const int p10 = S.patch[(i + 1) * PD + j], p11 = S.patch[(i + 1) * PD + j + 1];
        sDX[p] = (float)(p01 - p00 + p11 - p10) * dw;
        sDY[p] = (float)(p10 - p00 + p11 - p01) * dw;
    }
    __syncwarp();
    {   // lane = (cell, half): 4 running sums each, raster order inside the 5x5 cell
        const int cell = lane >> 1, half = lane & 1;
        const int ci = cell >> 2, cj = cell & 3;
        float v0 = 0, v1 = 0, v2 = 0, v3 = 0;
        for (int y = ci * 5; y < ci * 5 + 5; y++)
#pragma unroll
            for (int x5 = 0; x5 < 5; x5++) {
                const int x = cj * 5 + x5;
                const float tx = sDX[y * PATCH_SZ + x], ty = sDY[y * PATCH_SZ + x];
                if (extended) {
                    // half 0: tx sums split by sign(ty); half 1: ty sums split by sign(tx)
                    const float u = half ? ty : tx, g = half ? tx : ty;
                    if (g >= 0) { v0 += u; v1 += fabsf(u); } else { v2 += u; v3 += fabsf(u); }
                } else {
                    // 64-d: (sum tx, sum ty, sum |tx|, sum |ty|); half 0 -> (v0, v2) from tx, half 1 -> from ty
                    const float u = half ? ty : tx;
                    v0 += u; v1 += fabsf(u);
                }
            }
        if (extended) {
            float *d = S.vec + cell * 8 + half * 4;
            d[0] = v0; d[1] = v1; d[2] = v2; d[3] = v3;
        } else {
            float *d = S.vec + cell * 4;
            d[half] = v0; d[2 + half] = v1;
        }
    }
    __syncwarp();
    // sum of squares: float products accumulated in double (order differences are far below float resolution)
    double sq = 0;
    for (int t = lane; t < dsize; t += 32) sq += (double)(S.vec[t] * S.vec[t]);
#pragma unroll
    for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float nscale = (float)(1. / (sqrt(sq) + (double)FLT_EPSILON));
    for (int t = lane; t < dsize; t += 32) dst[t] = S.vec[t] * nscale;
    __syncwarp();
}

// ---- INTER_AREA (cv::resize of the win x win u8 window to 21 x 21) folded over window rows arriving in chunks, in order.
// cv::resize's decimation table of output column c -- first interior tap sx1, n interior taps of weight axm, optional edge taps
// sx1 - 1 (axl) and sx1 + n (axr) -- is built once per keypoint by lane c into S.vec (free until the descriptor); the row table of
// output row dy is the same entry (square window, one scale).  Per chunk:
//   horizontal pass: the (row, column) pairs of the chunk are spread over ALL 32 lanes; each sums its taps in table order and the row
//       sum replaces element `column` of its row in place (a column's taps lie right of every column index written before it);
//   vertical pass: lane dx < 21 folds the row sums of column dx into `sum` with the row weights, emitting patch rows as they complete
//       (spreading the (output row, column) pairs over all lanes was measured: slower, the per-item set-up outweighs the idle lanes).
// Integer scales (OpenCV's box-sum fast path) and win == 21 use the same passes with unit weights and their own rounding at the end.
struct __align__(16) AreaCol {                  // nf = n | flags << 16; flags: 1 = left edge tap, 2 = right edge tap
    int sx1, nf, ya, yb;                        // ya / yb: first / last source row of output row c (the entry read as a row table)
    float axl, axm, axr, pad;
};

struct AreaFold {
    int iscale, mode;            // mode 0: general decimation tables, 1: integer scale (box sums), 2: win == 21 (copy)
    int nmax, dy; float sum;
    int dy_end; uint8_t *patch;  // output rows [dy at init, dy_end) go to patch[row * 21 + col] (default: all 21 rows into S.patch)

    // The output rows [dy_begin, dy_end) only: their source rows start at first_row() (call after init), every output row is
    // still summed over its own source rows in order, so several warps can fold disjoint row ranges of one window exactly.
    __device__ __forceinline__ void restrict_rows(const DescScratch &S, int dy_begin, int dy_stop, uint8_t *out)
    {
        dy = dy_begin; dy_end = dy_stop; patch = out;
    }
    __device__ __forceinline__ int first_row(const DescScratch &S) const
    {
        const AreaCol c = ((const AreaCol *)S.vec)[dy];
        return c.sx1 - ((c.nf >> 16) & 1);
    }

    __device__ __forceinline__ void init(DescScratch &S, int win, int lane)
    {
        constexpr int PD = PATCH_SZ + 1;
        const double inv_scale = (double)PD / win;
        const double scale = 1. / inv_scale;
        iscale = __double2int_rn(scale);
        mode = win == PD ? 2 : (fabs(scale - iscale) < DBL_EPSILON ? 1 : 0);
        dy = 0; sum = 0; dy_end = PD; patch = S.patch;
        const int dxc = lane < PD ? lane : PD - 1;
        AreaCol c;
        c.sx1 = dxc * iscale; c.nf = iscale; c.axl = c.axr = c.pad = 0; c.axm = 1.f; c.ya = c.yb = 0;
        if (mode == 0) {
            // decimation table entries of output column dxc, in table order
            const double fsx1 = dxc * scale, fsx2 = fsx1 + scale, cwx = fmin(scale, win - fsx1);
            int sx2 = __double2int_rd(fsx2);
            c.sx1 = __double2int_ru(fsx1);
            sx2 = min(sx2, win - 1); c.sx1 = min(c.sx1, sx2);
            c.nf = (sx2 - c.sx1) | (c.sx1 - fsx1 > 1e-3 ? 1 << 16 : 0) | (fsx2 - sx2 > 1e-3 ? 2 << 16 : 0);
            c.axl = (float)((c.sx1 - fsx1) / cwx); c.axm = (float)(1.0 / cwx);
            c.axr = (float)(fmin(fmin(fsx2 - sx2, 1.), cwx) / cwx);
        }
        c.ya = c.sx1 - ((c.nf >> 16) & 1); c.yb = c.sx1 + (c.nf & 0xffff) - 1 + ((c.nf >> 17) & 1);
        nmax = c.nf & 0xffff;
#pragma unroll
        for (int o = 16; o; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));
        if (lane < PD) ((AreaCol *)S.vec)[lane] = c;
        __syncwarp();
    }

    __device__ __forceinline__ bool done() const { return dy >= dy_end; }

    // rows r0 .. r0 + nrows - 1 of the window (exact floats 0..255 in S.buf, `pitch` floats apart, overwritten); chunks arrive in order
    __device__ __forceinline__ void rows(DescScratch &S, int pitch, int r0, int nrows, int lane)
    {
        constexpr int PD = PATCH_SZ + 1;
        if (dy >= dy_end) return;
        const AreaCol *tab = (const AreaCol *)S.vec;
        float *buf = S.buf;
        // ---- horizontal pass
        const int nitems = nrows * PD;
        for (int id0 = 0; id0 < nitems; id0 += 32) {
            const int id = min(id0 + lane, nitems - 1);
            const int row = (id * 3121) >> 16, col = id - row * PD;       // id / 21 for id < 2 ^ 12
            const AreaCol c = tab[col];
            const int n = c.nf & 0xffff;
            const float *p = buf + row * pitch + c.sx1;
            float bufv = (c.nf & (1 << 16)) ? p[-1] * c.axl : 0.f;
            // taps 0..3 straight-line (all there is for windows up to ~100 px), the rest in a loop to the warp maximum
            if (0 < n) bufv += p[0] * c.axm;
            if (1 < n) bufv += p[1] * c.axm;
            if (2 < n) bufv += p[2] * c.axm;
            if (3 < n) bufv += p[3] * c.axm;
            if (nmax > 4) {
#pragma unroll 4
                for (int t = 4; t < nmax; t++) if (t < n) bufv += p[t] * c.axm;
            }
            if (c.nf & (2 << 16)) bufv += p[n] * c.axr;
            __syncwarp();
            if (id0 + lane < nitems) buf[row * pitch + col] = bufv;
        }
        __syncwarp();
        // ---- vertical pass (sum starts at 0: 0 + x == x, the CPU's first assignment)
        const int rend = r0 + nrows;
        const float *colp = buf + min(lane, PD - 1) - r0 * pitch;          // row sum of source row sy at colp[sy * pitch]
        while (true) {
            const AreaCol c = tab[dy];
            int sy = max(c.ya, r0);
            if ((c.nf & (1 << 16)) && sy == c.ya) { sum += c.axl * colp[sy * pitch]; sy++; }        // left-edge row (weight axl)
            const bool ends = c.yb < rend, tail = ends && (c.nf & (2 << 16));
            const int last = tail ? c.yb - 1 : min(c.yb, rend - 1);
            for (; sy <= last; sy++) sum += c.axm * colp[sy * pitch];
            if (tail && sy == c.yb) sum += c.axr * colp[sy * pitch];                                  // right-edge row (weight axr)
            if (!ends) return;                                // this output row continues in the next chunk
            int out;
            if (mode == 0) out = min(max(round_half_even_u8(sum), 0), 255);
            else if (mode == 2) out = (int)sum;
            else if (iscale == 2) out = (int)((sum + 2.f) * 0.25f);          // (sum + 2) >> 2
            else out = min(max(__float2int_rn(sum * (1.f / (float)(iscale * iscale))), 0), 255);
            if (lane < PD) patch[dy * PD + lane] = (uint8_t)out;
            dy++; sum = 0;
            if (dy >= dy_end) return;
            // (written with the whole entry loaded: `dy >= PD || tab[dy].ya >= rend` gave wrong results on the B200 with nvcc 12.9)
            { const AreaCol cn = tab[dy]; if ((cn.sx1 - ((cn.nf >> 16) & 1)) >= rend) return; }     // the next output row starts in a later chunk
        }
    }
};

// ---------------------------------------------------------------- K4a: reference sampler (row at a time, double precision, u8 image)
// work_list == nullptr: every keypoint of the images [0, batch); else the items listed (flattened indices), *work_count of them.
__global__ void __launch_bounds__(WK_WARPS * 32, 4) describe_reference_kernel(
    const uint8_t *base_a, const uint8_t *base_b, int split, int64_t img_stride, int rows, int cols, int stride,
    const int32_t *__restrict__ integral, float *kp_all, float *desc_all, const int32_t *__restrict__ prefix,
    int batch, int kp_cap, int extended, int upright, int *work_counter, int *big_flag,
    const int *__restrict__ work_list, const int *__restrict__ work_count, const int64_t *__restrict__ img_off)
{
    __shared__ DescScratch s_ws[WK_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    DescScratch &S = s_ws[warp];
    const int total = work_list ? min(*work_count, prefix[batch]) : prefix[batch];
    const int W = cols + 1, srows = rows + 1, scols = cols + 1;
    const int dsize = extended ? 128 : 64;
    while (true) {
        int item = 0;
        if (lane == 0) item = atomicAdd(work_counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= total) break;
        if (work_list) item = work_list[item];
        int lo = 0, hi = batch;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(prefix + mid) <= item) lo = mid; else hi = mid; }
        const int b = lo, k = item - __ldg(prefix + lo);
        float *kp = kp_all + ((size_t)b * kp_cap + k) * KP_STRIDE;
        const float size = kp[KP_SIZE], cx = kp[KP_X], cy = kp[KP_Y];
        const float s = size * 1.2f / 9.0f;
        const int win = (int)((PATCH_SZ + 1) * s);
        if (win > WK_MAX_WIN) { if (lane == 0) *big_flag = 1; continue; }     // warp-uniform: flag work for the CTA kernel
        const uint8_t *img = image_ptr(base_a, base_b, split, b, img_stride, img_off);
        const int32_t *I = integral + (size_t)b * srows * W;
        const int gws = 2 * __float2int_rn(2 * s);
        float descriptor_dir = 360.f - 90.f;
        if (!upright) descriptor_dir = orient_keypoint(S, lane, I, W, srows, scols, cx, cy, s, gws);
        if (lane == 0) kp[KP_ANGLE] = descriptor_dir;

        const int ncols1 = cols - 1, nrows1 = rows - 1;
        float sin_dir = 0, cos_dir = 0, chain_x = 0, chain_y = 0;
        int ustart_x = 0, ustart_y = 0;
        const float win_offset = -(float)(win - 1) / 2;
        if (!upright) {
            const float dir_rad = descriptor_dir * (float)(M_PI / 180);
            double sd, cd;
            sincos((double)dir_rad, &sd, &cd);
            sin_dir = -(float)sd; cos_dir = (float)cd;
            chain_x = cx + win_offset * cos_dir + win_offset * sin_dir;
            chain_y = cy - win_offset * sin_dir + win_offset * cos_dir;
        } else {
            ustart_x = __float2int_rn(cx + win_offset);
            ustart_y = __float2int_rn(cy - win_offset);
        }
        AreaFold F;
        F.init(S, win, lane);
        float *rowf = S.buf;
        for (int r = 0; r < win && !F.done(); r++) {
            if (!upright) {
                // per-lane positions advance by 32 columns per round; all terms are exact in double (24-bit increments,
                // |x| < 2^13), so the running sum equals the CPU's column-by-column accumulation
                const double rx = (double)chain_x, ry = (double)chain_y;
                double px = rx + (double)lane * (double)cos_dir, py = ry - (double)lane * (double)sin_dir;
                const double dpx = 32.0 * (double)cos_dir, dpy = 32.0 * (double)sin_dir;
                for (int j = lane; j < win; j += 32, px += dpx, py -= dpy)
                    rowf[j] = (float)window_pixel(img, stride, ncols1, nrows1, px, py);
                chain_x += sin_dir; chain_y += cos_dir;
            } else {
                const int x = min(max(ustart_x + r, 0), cols - 1);
                for (int j = lane; j < win; j += 32) {
                    const int y = min(max(ustart_y - j, 0), rows - 1);
                    rowf[j] = (float)img[(size_t)y * stride + x];
                }
            }
            __syncwarp();
            F.rows(S, 0, r, 1, lane);
            __syncwarp();
        }
        patch_to_descriptor(S, lane, extended, desc_all + ((size_t)b * kp_cap + k) * dsize);
    }
}

// ---------------------------------------------------------------- K4b: fixed-point chunked sampler (the product path)
// One launch per group of images sharing a stacked texture (image b at texture rows [(b - b_first) * rows, ...)).
// Keypoints whose direction or row starts are not multiples of 2^-32 go to `fb_list` for describe_reference_kernel.
__device__ __forceinline__ float bilinear_scaled(const float4 g, float A, float B)
{
    const float T = 4294967296.0f;
    const float ia = T - A, ib = T - B;
    const float p00 = g.w, p01 = g.z, p10 = g.x, p11 = g.y;
    const float v = p00 * ia * ib + p01 * A * ib + p10 * ia * B + p11 * A * B;
    return (v + 12582912.0f) - 12582912.0f;          // cvRound, ties to even, as an exact float
}

// U warp-rounds of a chunk: the U gathers are issued before the first is consumed.  slot: their entries of the slot table, out: this
// lane's element of the first round's row buffer, mxc / myc: this lane's column offset inside a round times the per-column step.
// CHECK: some sample of these rounds may lie outside the image -- such samples take the CPU's clamped nearest pixel.
// FINE: 16.48 instead of 32.32 fixed point; the fraction then takes the slow 64-bit conversion.
template <bool CHECK, bool FINE, int U>
__device__ __forceinline__ void sample_rounds(const ulonglong2 *__restrict__ slot, float *__restrict__ out, const cudaTextureObject_t tex,
                                              unsigned long long mxc, unsigned long long myc, const uint8_t *__restrict__ img, int stride,
                                              int ncols1, int nrows1, int row_off)
{
    constexpr unsigned long long FMASK = (1ULL << 48) - 1, FHALF = 1ULL << 47;
    float4 g[U]; float A[U], B[U]; unsigned oob = 0;
#pragma unroll
    for (int u = 0; u < U; u++) {
        const ulonglong2 e = slot[u];
        const unsigned long long X = e.x + mxc, Y = e.y + myc;
        int ix1, iy1;
        if (!FINE) {
            ix1 = (int)(X >> 32); iy1 = (int)(Y >> 32);
            A[u] = __uint2float_rn((unsigned)X); B[u] = __uint2float_rn((unsigned)Y);
        } else {
            ix1 = (int)((long long)X >> 48); iy1 = (int)((long long)Y >> 48) + row_off;
            A[u] = __ull2float_rn(X & FMASK) * 1.52587890625e-05f;      // one rounding of the 48-bit fraction, then an exact 2^-16
            B[u] = __ull2float_rn(Y & FMASK) * 1.52587890625e-05f;
        }
        g[u] = tex2Dgather<float4>(tex, __int2float_rn(ix1), __int2float_rn(iy1), 0);
        if (CHECK && !((unsigned)(ix1 - 1) < (unsigned)ncols1 && (unsigned)(iy1 - 1 - row_off) < (unsigned)nrows1)) oob |= 1u << u;
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        float v = bilinear_scaled(g[u], A[u], B[u]);
        if (CHECK && (oob >> u & 1)) {
            // outside the image: the CPU takes the clamped nearest pixel, cvRound(pixel) = ties to even
            const ulonglong2 e = slot[u];
            const unsigned long long X = e.x + mxc, Y = e.y + myc;
            int ix, iy; bool upx, upy;
            if (!FINE) {
                ix = (int)(X >> 32) - 1; iy = (int)(Y >> 32) - 1 - row_off;
                const unsigned fx = (unsigned)X, fy = (unsigned)Y;
                upx = fx > 0x80000000u || (fx == 0x80000000u && (ix & 1));
                upy = fy > 0x80000000u || (fy == 0x80000000u && (iy & 1));
            } else {
                ix = (int)((long long)X >> 48) - 1; iy = (int)((long long)Y >> 48) - 1;
                const unsigned long long fx = X & FMASK, fy = Y & FMASK;
                upx = fx > FHALF || (fx == FHALF && (ix & 1));
                upy = fy > FHALF || (fy == FHALF && (iy & 1));
            }
            const int x = min(max(ix + (upx ? 1 : 0), 0), ncols1), y = min(max(iy + (upy ? 1 : 0), 0), nrows1);
            v = (float)img[(size_t)y * stride + x];
        }
        out[u * 32] = v;
    }
}

// Samples the warp-rounds [0, Q4) of the chunk described by S.slot into S.buf, round q at buf[32 q + lane] (Q4: Q rounded up to a
// multiple of U, the padding rounds repeat the last one).  need: bit q set when round q may hold a sample outside the image (0 for
// windows that lie inside it); groups of U rounds without such a bit skip the test.
template <bool FINE, int U>
__device__ __forceinline__ void sample_chunk(DescScratch &S, const cudaTextureObject_t tex, int Q4, unsigned need, unsigned long long mxc,
                                             unsigned long long myc, int lane, const uint8_t *__restrict__ img, int stride,
                                             int ncols1, int nrows1, int row_off)
{
    const ulonglong2 *slot = (const ulonglong2 *)S.slot;
    float *out = S.buf + lane;
#pragma unroll 1
    for (int q0 = 0; q0 < Q4; q0 += U, slot += U, out += 32 * U, need >>= U) {
        if (need & ((1u << U) - 1)) sample_rounds<true, FINE, U>(slot, out, tex, mxc, myc, img, stride, ncols1, nrows1, row_off);
        else sample_rounds<false, FINE, U>(slot, out, tex, mxc, myc, img, stride, ncols1, nrows1, row_off);
    }
}

// ---- tolerance modes (option describe = 2 / 3; NOT bit-exact, DESIGN.md "tolerance modes"), both at float positions:
// TOL = 1 (describe = 2): the 2x2 footprint through a gather, fp32 weights, three fused lerps -- the real bilinear value to a few
//   1e-6 of a grey level, so a window pixel differs from the CPU's only when its value sits that close to a .5 boundary;
// TOL = 2 (describe = 3): the texture unit's own bilinear filter (one TEX; the unit interpolates with 8-bit weights).  slot: float2 position of lane 0 of every warp-round, texel-centre convention, image b's rows
// included; mxf / myf: this lane's offset along the window row.  Everything after the window pixel (rounding to u8, INTER_AREA,
// gradients, bins, normalisation) is the exact path's code.
template <bool CHECK, int U, int TOL>
__device__ __forceinline__ void sample_rounds_tol(const float2 *__restrict__ slot, float *__restrict__ out, const cudaTextureObject_t tex,
                                                  float mxf, float myf, const uint8_t *__restrict__ img, int stride, int ncols1, int nrows1,
                                                  int row_off)
{
    float t[U], xs[U], ys[U];
    float4 g[U]; float fa[U], fb[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const float2 e = slot[u];
        xs[u] = e.x + mxf; ys[u] = e.y + myf;
        if (TOL == 2) t[u] = tex2D<float>(tex, xs[u], ys[u]);          // the texture unit's filter: 8-bit weights
        else {
            // fp32 weights, fused lerps: the footprint through a gather, no emulation of the CPU's operation order
            const float fx = floorf(xs[u] - 0.5f), fy = floorf(ys[u] - 0.5f);
            fa[u] = (xs[u] - 0.5f) - fx; fb[u] = (ys[u] - 0.5f) - fy;
            g[u] = tex2Dgather<float4>(tex, fx + 1.0f, fy + 1.0f, 0);
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        if (TOL != 2) {
            const float p00 = g[u].w, p01 = g[u].z, p10 = g[u].x, p11 = g[u].y;
            const float top = __fmaf_rn(fa[u], p01 - p00, p00), bot = __fmaf_rn(fa[u], p11 - p10, p10);
            t[u] = __fmaf_rn(fb[u], bot - top, top);
        }
        float v = (t[u] + 12582912.0f) - 12582912.0f;          // cvRound, ties to even
        if (CHECK) {
            const float px = xs[u] - 0.5f, py = ys[u] - 0.5f - (float)row_off;
            const int ix = (int)floorf(px), iy = (int)floorf(py);
            if (!((unsigned)ix < (unsigned)ncols1 && (unsigned)iy < (unsigned)nrows1)) {
                // outside the image: the CPU takes the clamped nearest pixel
                const int x = min(max(__float2int_rn(px), 0), ncols1), y = min(max(__float2int_rn(py), 0), nrows1);
                v = (float)img[(size_t)y * stride + x];
            }
        }
        out[u * 32] = v;
    }
}

// The window of one keypoint -> S.patch.  false: a row start is not a multiple of one fixed-point unit (FINE = false: the caller
// retries in 16.48; FINE = true: the reference kernel's).  cU / sU: |cos_dir|, |sin_dir| in fixed-point units.
template <bool FINE, int U, int TOL = 0>
__device__ bool window_to_patch(DescScratch &S, const cudaTextureObject_t tex, int lane, int win, float cx, float cy,
                                                float sin_dir, float cos_dir, unsigned long long cU, unsigned long long sU, bool interior,
                                                const uint8_t *__restrict__ img, int stride, int ncols1, int nrows1, int row_off,
                                                int dy_begin = 0, int dy_stop = PATCH_SZ + 1, uint8_t *patch_out = nullptr)
{
    constexpr int FB = FINE ? 48 : 32;
    const float tiny = FINE ? 2.98023223876953125e-08f : 0.001953125f;        // row starts must be 0 or at least this (ulp >= one unit)
    const bool xneg = cos_dir < 0.f;            // x decreases along a row
    const bool yneg = sin_dir > 0.f;            // pixel_y -= sin_dir
    const unsigned long long mxc = (unsigned long long)(xneg ? 31 - lane : lane) * cU;
    const unsigned long long myc = (unsigned long long)(yneg ? 31 - lane : lane) * sU;
    const float win_offset = -(float)(win - 1) / 2;
    float chain_x = cx + win_offset * cos_dir + win_offset * sin_dir;      // start_x / start_y of the next unsampled row
    float chain_y = cy - win_offset * sin_dir + win_offset * cos_dir;
    const int kpr = (win + 31) >> 5;            // warp-rounds per window row; a row occupies 32 * kpr floats of S.buf
    const int R = DESC_SLOTS / kpr;             // rows per chunk
    const int kinv = (65536 + kpr - 1) / kpr;   // (q * kinv) >> 16 == q / kpr for q < 32, kpr <= 24
    AreaFold F;
    F.init(S, win, lane);
    int r_first = 0;
    if (patch_out) {                             // a share of the output rows (cooperative pass): skip to their first source row
        F.restrict_rows(S, dy_begin, dy_stop, patch_out);
        r_first = F.first_row(S);
        for (int i = 0; i < r_first; i++) { chain_x += sin_dir; chain_y += cos_dir; }      // the CPU's float chain up to that row
    }
    for (int r0 = r_first; r0 < win && !F.done(); r0 += R) {
        const int Rc = min(R, win - r0), Q = Rc * kpr;
        // ---- slot table: lane q owns warp-round q = (row r0 + q / kpr, columns 32 * (q % kpr) ...); lanes >= Q repeat round Q - 1
        unsigned need = 0;
        {
            const int q = min(lane, Q - 1);
            const int myrow = (q * kinv) >> 16, myk = q - myrow * kpr;
            float cap_x = 1.f, cap_y = 1.f;
            for (int i = 0; i < Rc; i++) {       // the CPU's float chain, advanced by every lane alike
                if (i == myrow) { cap_x = chain_x; cap_y = chain_y; }
                chain_x += sin_dir; chain_y += cos_dir;
            }
            if constexpr (TOL) {
                const float c0 = (float)(32 * myk);
                float2 e;
                e.x = cap_x + c0 * cos_dir + 0.5f;
                e.y = cap_y - c0 * sin_dir + 0.5f + (float)row_off;
                ((float2 *)S.slot)[lane] = e;
                if (!interior) {
                    const int xa = (int)floorf(e.x - 0.5f), xb = (int)floorf(e.x - 0.5f + 31.f * cos_dir);
                    const int ya = (int)floorf(e.y - 0.5f) - row_off, yb = (int)floorf(e.y - 0.5f - 31.f * sin_dir) - row_off;
                    const bool in = (unsigned)xa < (unsigned)ncols1 && (unsigned)xb < (unsigned)ncols1 &&
                                    (unsigned)ya < (unsigned)nrows1 && (unsigned)yb < (unsigned)nrows1;
                    need = __ballot_sync(0xffffffffu, !in);
                }
            } else {
            const bool ok = (cap_x == 0.f || fabsf(cap_x) >= tiny) && (cap_y == 0.f || fabsf(cap_y) >= tiny);
            if (!__all_sync(0xffffffffu, ok)) return false;
            long long X0, Y0;                    // exact: the row start is a multiple of one unit
            if (FINE) {
                X0 = (long long)((double)cap_x * 281474976710656.0); Y0 = (long long)((double)cap_y * 281474976710656.0);
            } else {
                const float flx = floorf(cap_x), fly = floorf(cap_y);
                X0 = ((long long)(int)flx << 32) | (unsigned)((cap_x - flx) * 4294967296.0f);
                Y0 = ((long long)(int)fly << 32) | (unsigned)((cap_y - fly) * 4294967296.0f);
            }
            const int kx = xneg ? -(32 * myk + 31) : 32 * myk, ky = yneg ? -(32 * myk + 31) : 32 * myk;
            // + 1: tex2Dgather at (ix + 1, iy + 1) returns the footprint (ix, iy) .. (ix + 1, iy + 1); + row_off: image b of the stack
            // (added after the shift in 16.48, whose 16 integer bits do not hold the stack's rows)
            ulonglong2 e;
            if (FINE) {
                e.x = (unsigned long long)(X0 + (long long)kx * (long long)cU + (1LL << FB));
                e.y = (unsigned long long)(Y0 + (long long)ky * (long long)sU + (1LL << FB));
            } else {
                e.x = (unsigned long long)(X0 + (long long)kx * (long long)(unsigned)cU + (1LL << FB));
                e.y = (unsigned long long)(Y0 + (long long)ky * (long long)(unsigned)sU + ((long long)(1 + row_off) << FB));
            }
            ((ulonglong2 *)S.slot)[lane] = e;
            if (!interior) {
                // a round whose two end lanes have their 2x2 footprint inside the image holds no outside sample (collinear positions)
                const unsigned long long Xb = e.x + 31 * cU, Yb = e.y + 31 * sU;
                const int xa = (int)((long long)e.x >> FB) - 1, xb = (int)((long long)Xb >> FB) - 1;
                const int ya = (int)((long long)e.y >> FB) - 1 - (FINE ? 0 : row_off), yb = (int)((long long)Yb >> FB) - 1 - (FINE ? 0 : row_off);
                const bool in = (unsigned)xa < (unsigned)ncols1 && (unsigned)xb < (unsigned)ncols1 &&
                                (unsigned)ya < (unsigned)nrows1 && (unsigned)yb < (unsigned)nrows1;
                need = __ballot_sync(0xffffffffu, !in);
            }
            }
        }
        __syncwarp();
        const int Q4 = (Q + U - 1) / U * U;
        if constexpr (TOL) {
            const float mxf = (float)lane * cos_dir, myf = -(float)lane * sin_dir;
            const float2 *slot = (const float2 *)S.slot;
            float *out = S.buf + lane;
#pragma unroll 1
            for (int q0 = 0; q0 < Q4; q0 += U, slot += U, out += 32 * U, need >>= U) {
                if (need & ((1u << U) - 1)) sample_rounds_tol<true, U, TOL>(slot, out, tex, mxf, myf, img, stride, ncols1, nrows1, row_off);
                else sample_rounds_tol<false, U, TOL>(slot, out, tex, mxf, myf, img, stride, ncols1, nrows1, row_off);
            }
        } else
        sample_chunk<FINE, U>(S, tex, Q4, need, mxc, myc, lane, img, stride, ncols1, nrows1, row_off);
        __syncwarp();
        // ---- fold the chunk's rows into the patch
        F.rows(S, kpr * 32, r0, Rc, lane);
        __syncwarp();
    }
    return F.done();
}

// Work list of the cooperative pass: the keypoints whose window is at least coop_split pixels wide (and that the warp kernel can
// hold at all).  Small batches only: there a single giant window on a single warp is the whole tail of the launch.
__global__ void __launch_bounds__(256) describe_giants_kernel(const float *__restrict__ kp_all, const int32_t *__restrict__ prefix, int batch,
                                                              int kp_cap, int coop_split, int *coop_list, int *coop_count)
{
    const int total = prefix[batch];
    for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < total; item += gridDim.x * blockDim.x) {
        int lo = 0, hi = batch;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(prefix + mid) <= item) lo = mid; else hi = mid; }
        const float size = kp_all[((size_t)lo * kp_cap + (item - __ldg(prefix + lo))) * KP_STRIDE + KP_SIZE];
        const int win = (int)((PATCH_SZ + 1) * (size * 1.2f / 9.0f));
        if (win >= coop_split && win <= WK_MAX_WIN) coop_list[atomicAdd(coop_count, 1)] = item;
    }
}

// coop_split > 0 (COOP instances, single launch group): before the per-warp passes the CTA's warps share every window of at least
// coop_split pixels -- each warp samples and folds its share of the 21 output rows (AreaFold::restrict_rows), warp 0 turns the
// assembled patch into the descriptor.  Same arithmetic per output row, so the descriptor is the per-warp path's.
template <int MINB, int U, int NW, int TOL = 0, bool COOP = false>
__global__ void __launch_bounds__(NW * 32, MINB) describe_fixed_kernel(
    const uint8_t *base_a, const uint8_t *base_b, int split, int64_t img_stride, int rows, int cols, int stride,
    const int32_t *__restrict__ integral, float *kp_all, float *desc_all, const int32_t *__restrict__ prefix,
    int batch, int kp_cap, int extended, const cudaTextureObject_t tex, int b_first, int b_count,
    int *work_counter, int *work_counter_large, int lpt_split, int *big_flag, int *fb_list, int *fb_count,
    const int64_t *__restrict__ img_off, int coop_split = 0, const int *coop_list = nullptr, const int *coop_count = nullptr,
    int *coop_counter = nullptr)
{
    __shared__ DescScratch s_ws[NW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    DescScratch &S = s_ws[warp];
    const int total = prefix[b_first + b_count];
    const int item0 = prefix[b_first];
    const int W = cols + 1, srows = rows + 1, scols = cols + 1;
    const int dsize = extended ? 128 : 64;
    const int ncols1 = cols - 1, nrows1 = rows - 1;

    if constexpr (COOP) {
        constexpr int PD = PATCH_SZ + 1;
        __shared__ int s_item, s_fail, s_fail_fine;
        __shared__ uint8_t s_patch[PD * PD + 7];
        const int n_coop = *coop_count;
        while (true) {
            if (threadIdx.x == 0) { s_item = atomicAdd(coop_counter, 1); s_fail = 0; s_fail_fine = 0; }
            __syncthreads();
            const int ci = s_item;
            if (ci >= n_coop) break;                           // CTA-uniform
            const int item = coop_list[ci];
            int lo = b_first, hi = b_first + b_count;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(prefix + mid) <= item) lo = mid; else hi = mid; }
            const int b = lo, k = item - __ldg(prefix + lo);
            float *kp = kp_all + ((size_t)b * kp_cap + k) * KP_STRIDE;
            const float size = kp[KP_SIZE], cx = kp[KP_X], cy = kp[KP_Y];
            const float s = size * 1.2f / 9.0f;
            const int win = (int)((PATCH_SZ + 1) * s);
            const uint8_t *img = image_ptr(base_a, base_b, split, b, img_stride, img_off);
            const int32_t *I = integral + (size_t)b * srows * W;
            const int gws = 2 * __float2int_rn(2 * s);
            // every warp repeats the orientation (a few thousand instructions against the > 10^5 of its share of the window)
            const float descriptor_dir = orient_keypoint(S, lane, I, W, srows, scols, cx, cy, s, gws);
            const float dir_rad = descriptor_dir * (float)(M_PI / 180);
            double sd, cd;
            sincos((double)dir_rad, &sd, &cd);
            const float sin_dir = -(float)sd, cos_dir = (float)cd;
            const float ac = fabsf(cos_dir), as = fabsf(sin_dir);
            const bool ok32 = (ac == 0.f || ac >= 0.001953125f) && ac < 1.f && (as == 0.f || as >= 0.001953125f) && as < 1.f;
            const int row_off = (b - b_first) * rows;
            const float Rw = (float)(win - 1) * 0.7072f + 2.0f;
            const bool interior = cx - Rw >= 1.f && cx + Rw <= (float)(ncols1 - 1) && cy - Rw >= 1.f && cy + Rw <= (float)(nrows1 - 1);
            const int dy0 = warp * PD / NW, dy1 = (warp + 1) * PD / NW;
            bool mine = TOL != 0 || ok32;
            if (mine)
                mine = window_to_patch<false, U, TOL>(S, tex, lane, win, cx, cy, sin_dir, cos_dir, (unsigned long long)(unsigned)(ac * 4294967296.0f),
                                                      (unsigned long long)(unsigned)(as * 4294967296.0f), interior, img, stride, ncols1, nrows1, row_off,
                                                      dy0, dy1, s_patch);
            if (!mine && lane == 0) s_fail = 1;
            __syncthreads();
            if (TOL == 0 && s_fail) {                          // CTA-uniform: some share needs the finer unit -> all shares again in 16.48
                const bool ok48 = (ac == 0.f || ac >= 2.98023223876953125e-08f) && (as == 0.f || as >= 2.98023223876953125e-08f) &&
                                  rows < 16384 && cols < 16384;
                mine = ok48 && window_to_patch<true, 2>(S, tex, lane, win, cx, cy, sin_dir, cos_dir, (unsigned long long)((double)ac * 281474976710656.0),
                                                        (unsigned long long)((double)as * 281474976710656.0), interior, img, stride, ncols1, nrows1,
                                                        row_off, dy0, dy1, s_patch);
                if (!mine && lane == 0) s_fail_fine = 1;
                __syncthreads();
            }
            if (warp == 0) {
                if (s_fail && (TOL != 0 || s_fail_fine)) {     // neither unit fits: the reference kernel takes the keypoint
                    if (lane == 0) fb_list[atomicAdd(fb_count, 1)] = item;
                } else {
                    if (lane == 0) kp[KP_ANGLE] = descriptor_dir;
                    for (int i = lane; i < PD * PD; i += 32) S.patch[i] = s_patch[i];
                    __syncwarp();
                    patch_to_descriptor(S, lane, extended, desc_all + ((size_t)b * kp_cap + k) * dsize);
                }
            }
            __syncthreads();                                   // s_patch / s_item are reused by the next window
        }
    }

    // lpt_split > 0: longest-processing-time-first in two passes over the same work list -- pass 0 describes only the windows
    // >= lpt_split (a 600-pixel window keeps one warp busy for a long time; met late in the queue it becomes the tail of the
    // launch), pass 1 the rest.  A skipped item costs one keypoint read.
    int pass = lpt_split > 0 ? 0 : 1;
    while (true) {
        int item = 0;
        if (lane == 0) item = atomicAdd(pass == 0 ? work_counter_large : work_counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0) + item0;
        if (item >= total) { if (pass == 0) { pass = 1; continue; } break; }
        int lo = b_first, hi = b_first + b_count;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(prefix + mid) <= item) lo = mid; else hi = mid; }
        const int b = lo, k = item - __ldg(prefix + lo);
        float *kp = kp_all + ((size_t)b * kp_cap + k) * KP_STRIDE;
        const float size = kp[KP_SIZE], cx = kp[KP_X], cy = kp[KP_Y];
        const float s = size * 1.2f / 9.0f;
        const int win = (int)((PATCH_SZ + 1) * s);
        if (lpt_split > 0 && ((win >= lpt_split) != (pass == 0))) continue;   // warp-uniform: the other pass owns this keypoint
        if (win > WK_MAX_WIN) { if (lane == 0) *big_flag = 1; continue; }     // warp-uniform: flag work for the CTA kernel
        if (COOP && win >= coop_split) continue;                              // described by the cooperative pass above
        const uint8_t *img = image_ptr(base_a, base_b, split, b, img_stride, img_off);
        const int32_t *I = integral + (size_t)b * srows * W;
        const int gws = 2 * __float2int_rn(2 * s);
        const float descriptor_dir = orient_keypoint(S, lane, I, W, srows, scols, cx, cy, s, gws);
        if (lane == 0) kp[KP_ANGLE] = descriptor_dir;

        const float dir_rad = descriptor_dir * (float)(M_PI / 180);
        double sd, cd;
        sincos((double)dir_rad, &sd, &cd);
        const float sin_dir = -(float)sd, cos_dir = (float)cd;
        const float ac = fabsf(cos_dir), as = fabsf(sin_dir);
        // both steps multiples of 2^-32 below 1 -> 32.32 positions; multiples of 2^-48 up to 1 -> 16.48 (slower conversions, rare)
        const bool ok32 = (ac == 0.f || ac >= 0.001953125f) && ac < 1.f && (as == 0.f || as >= 0.001953125f) && as < 1.f;
        const bool ok48 = (ac == 0.f || ac >= 2.98023223876953125e-08f) && (as == 0.f || as >= 2.98023223876953125e-08f) &&
                          rows < 16384 && cols < 16384;
        const int row_off = (b - b_first) * rows;
        // every sample of the window (half diagonal + the float chain's drift, 2 px of slack) keeps its 2x2 footprint inside the image
        const float Rw = (float)(win - 1) * 0.7072f + 2.0f;
        const bool interior = cx - Rw >= 1.f && cx + Rw <= (float)(ncols1 - 1) && cy - Rw >= 1.f && cy + Rw <= (float)(nrows1 - 1);

        bool described = false;
        if constexpr (TOL)
            described = window_to_patch<false, U, TOL>(S, tex, lane, win, cx, cy, sin_dir, cos_dir, 0ULL, 0ULL, interior, img, stride, ncols1, nrows1, row_off);
        else if (ok32)
            described = window_to_patch<false, U>(S, tex, lane, win, cx, cy, sin_dir, cos_dir, (unsigned long long)(unsigned)(ac * 4294967296.0f),
                                               (unsigned long long)(unsigned)(as * 4294967296.0f), interior, img, stride, ncols1, nrows1, row_off);
        if (!described && ok48)      // the direction needs the finer unit, or a row of the first attempt started within 2^-9 of an axis
            described = window_to_patch<true, 2>(S, tex, lane, win, cx, cy, sin_dir, cos_dir, (unsigned long long)((double)ac * 281474976710656.0),
                                              (unsigned long long)((double)as * 281474976710656.0), interior, img, stride, ncols1, nrows1, row_off);
        if (!described) {                            // hand over to the reference kernel (it repeats the orientation)
            if (lane == 0) fb_list[atomicAdd(fb_count, 1)] = item;
            continue;
        }
        patch_to_descriptor(S, lane, extended, desc_all + ((size_t)b * kp_cap + k) * dsize);
    }
}
