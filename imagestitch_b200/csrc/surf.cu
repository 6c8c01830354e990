// surf.cu -- SURF detect + describe for a batch of equally-sized u8 images, sm_100a.
//
// Replaces what the reference reaches through cv2.xfeatures2d.SURF_create().detectAndCompute
// (ImageUtility.py:258,262) and myGpuFeatures.detectAndDescribeBySurf (appendix/myGpuFeatures.cpp:67-104):
// integral image -> Hessian box-filter layers -> 3x3x3 NMS -> interpolation -> response order ->
// orientation -> rotated 20s window -> INTER_AREA 21x21 patch -> 4x4x(4|8) descriptor.
// Arithmetic follows the CPU algorithm restated in oracle/surf_oracle.c operation by operation (this file is
// compiled with -fmad=false so float expressions round exactly like the scalar CPU code); the structure does not:
//   * det/trace layers are never materialised in HBM: one CTA computes the (nOctaveLayers+2) det layers of a
//     64x16 sample tile (+1 halo) into shared memory and runs NMS + interpolation straight from there;
//   * no host synchronisation anywhere: counters stay on the device, later kernels size themselves from them;
//   * the response ordering (OpenCV's KeypointGreater) is a rank-by-counting pass, deterministic under any
//     atomic arrival order of the candidates;
//   * everything is batched over images (grid.y / flattened work lists) so one launch serves a whole batch of ROIs.
#include "common.cuh"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <algorithm>
#include <map>
#include <tuple>
#include <vector>

#define ORI_RADIUS 6
#define ORI_WIN 60
#define PATCH_SZ 20
#define ORI_SAMPLES 113
#define MAX_WIN 2048

#define HT_X 64          // Hessian tile (samples) of every octave but 1
#define HT_Y 16
#define HT1_X 32         // octave 1: its staged footprint (step 2) is 4 x larger per sample; 32 x 16 keeps 3 CTAs on an SM instead of 2
                         // (measured per 32-pair step: 64 x 16 1.46 ms, 32 x 16 1.36 ms, 32 x 8 1.59 ms)
#define HT1_Y 16
__host__ __device__ constexpr int ht_x(int o) { return o == 1 ? HT1_X : HT_X; }
__host__ __device__ constexpr int ht_y(int o) { return o == 1 ? HT1_Y : HT_Y; }
#define HT_THREADS 256
#define INT_BAND 32      // integral band height
#define INT_SUB 8        // rows staged per sub-step (one warp per row)

__constant__ int c_apt_x[ORI_SAMPLES];
__constant__ int c_apt_y[ORI_SAMPLES];
__constant__ float c_aptw[ORI_SAMPLES];
__constant__ float c_DW[PATCH_SZ * PATCH_SZ];

// ---------------------------------------------------------------- host: constant tables (same formulas as the oracle)
static void gaussian_kernel_f32(int n, double sigma, float *out)
{
    double sum = 0, scale2X = -0.5 / (sigma * sigma), tmp[64];
    for (int i = 0; i < n; i++) {
        double x = i - (n - 1) * 0.5;
        tmp[i] = (double)(float)exp(scale2X * x * x);
        sum += tmp[i];
    }
    sum = 1. / sum;
    for (int i = 0; i < n; i++) out[i] = (float)(tmp[i] * sum);
}

int surf_init_tables()
{
    static bool done = false;   // per process & device; tables are immutable
    if (done) return 0;
    float G[13], Gd[PATCH_SZ], aptw[ORI_SAMPLES], DW[PATCH_SZ * PATCH_SZ];
    int ax[ORI_SAMPLES], ay[ORI_SAMPLES], n = 0;
    gaussian_kernel_f32(13, 2.5, G);
    for (int i = -ORI_RADIUS; i <= ORI_RADIUS; i++)
        for (int j = -ORI_RADIUS; j <= ORI_RADIUS; j++)
            if (i * i + j * j <= ORI_RADIUS * ORI_RADIUS) { ax[n] = i; ay[n] = j; aptw[n++] = G[i + ORI_RADIUS] * G[j + ORI_RADIUS]; }
    gaussian_kernel_f32(PATCH_SZ, (double)3.3f, Gd);     // OpenCV's DESC_SIGMA is the float constant 3.3f, widened by getGaussianKernel
    for (int i = 0; i < PATCH_SZ; i++)
        for (int j = 0; j < PATCH_SZ; j++) DW[i * PATCH_SZ + j] = Gd[i] * Gd[j];
    CUDA_TRY(cudaMemcpyToSymbol(c_apt_x, ax, sizeof(ax)));
    CUDA_TRY(cudaMemcpyToSymbol(c_apt_y, ay, sizeof(ay)));
    CUDA_TRY(cudaMemcpyToSymbol(c_aptw, aptw, sizeof(aptw)));
    CUDA_TRY(cudaMemcpyToSymbol(c_DW, DW, sizeof(DW)));
    done = true;
    return 0;
}

// ---------------------------------------------------------------- host: plan
static void resize_haar_host(const int src[][5], HaarBox *dst, int n, int oldSize, int newSize)
{
    float ratio = (float)newSize / oldSize;
    for (int k = 0; k < n; k++) {
        int dx1 = (int)lrint(ratio * src[k][0]), dy1 = (int)lrint(ratio * src[k][1]);
        int dx2 = (int)lrint(ratio * src[k][2]), dy2 = (int)lrint(ratio * src[k][3]);
        dst[k].x1 = (short)dx1; dst[k].y1 = (short)dy1; dst[k].x2 = (short)dx2; dst[k].y2 = (short)dy2;
        dst[k].w = src[k][4] / ((float)(dx2 - dx1) * (dy2 - dy1));
    }
}

int surf_build_plan(SurfPlan *plan, int rows, int cols, const vfsms_surf_params *p)
{
    static const int dx_s[3][5] = { {0, 2, 3, 7, 1}, {3, 2, 6, 7, -2}, {6, 2, 9, 7, 1} };
    static const int dy_s[3][5] = { {2, 0, 7, 3, 1}, {2, 3, 7, 6, -2}, {2, 6, 7, 9, 1} };
    static const int dxy_s[4][5] = { {1, 1, 4, 4, 1}, {5, 1, 8, 4, -1}, {1, 5, 4, 8, -1}, {5, 5, 8, 8, 1} };
    if (p->n_octaves < 1 || p->n_octaves > VFSMS_MAX_OCTAVES || p->n_octave_layers < 1 ||
        p->n_octave_layers + 2 > VFSMS_MAX_LAYERS_PER_OCTAVE || rows < 1 || cols < 1) {
        vfsms_set_error("surf: unsupported parameters (octaves %d, layers %d, %dx%d)", p->n_octaves, p->n_octave_layers, rows, cols);
        return VFSMS_E_ARG;
    }
    memset(plan, 0, sizeof(*plan));
    plan->rows = rows; plan->cols = cols;
    plan->n_octaves = p->n_octaves; plan->n_layers = p->n_octave_layers + 2;
    plan->threshold = p->hessian_threshold;
    int tiles = 0;
    for (int o = 0; o < plan->n_octaves; o++) {
        int step = 1 << o;
        int lrows = rows / step, lcols = cols / step;
        for (int l = 0; l < plan->n_layers; l++) {
            SurfLayer &L = plan->layer[o][l];
            L.size = (9 + 6 * l) << o;
            L.margin = (L.size / 2) / step;
            if (L.size > rows || L.size > cols) { L.samples_i = L.samples_j = 0; }
            else { L.samples_i = 1 + (rows - L.size) / step; L.samples_j = 1 + (cols - L.size) / step; }
            resize_haar_host(dx_s, L.dx, 3, 9, L.size);
            resize_haar_host(dy_s, L.dy, 3, 9, L.size);
            resize_haar_host(dxy_s, L.dxy, 4, 9, L.size);
        }
        {   // integral footprint of one tile (haloed) over all layers: offsets relative to (tile origin - 1) * step
            int off_min = 0, off_max = 0;
            for (int l = 0; l < plan->n_layers; l++) {
                const SurfLayer &L = plan->layer[o][l];
                off_min = std::min(off_min, -L.margin * step);
                off_max = std::max(off_max, -L.margin * step + L.size);
            }
            plan->stage_off_min[o] = off_min;
            plan->stage_rows[o] = (ht_y(o) + 1) * step + (off_max - off_min) + 1;
            plan->stage_cols[o] = (ht_x(o) + 1) * step + (off_max - off_min) + 1;
        }
        // NMS positions exist only inside [margin_min, l - margin_min): skip tiles that cannot hold a maximum
        int tx = ceil_div(lcols > 0 ? lcols : 1, ht_x(o)), ty = ceil_div(lrows > 0 ? lrows : 1, ht_y(o));
        if (lrows < 3 || lcols < 3) { tx = 0; ty = 0; }
        plan->tiles_x[o] = tx;
        plan->tile_begin[o] = tiles;
        tiles += tx * ty;
    }
    plan->tile_begin[plan->n_octaves] = tiles;
    return 0;
}

// ---------------------------------------------------------------- device helpers
// image b of a batch: the first `split` images lie img_stride apart from base_a, the others from base_b -- or, when the caller
// passed a table (vfsms_tiles_align_list: arbitrary tiles of the stack, mixed strip positions), at base_a + img_off[b]
__device__ __forceinline__ const uint8_t *image_ptr(const uint8_t *base_a, const uint8_t *base_b, int split, int b, int64_t img_stride,
                                                    const int64_t *__restrict__ img_off = nullptr)
{
    if (img_off) return base_a + img_off[b];
    return b < split ? base_a + (int64_t)b * img_stride : base_b + (int64_t)(b - split) * img_stride;
}

// ---------------------------------------------------------------- K1a: band-local integral
// grid (bands, batch), 256 threads = 8 warps, one warp per staged row.  Output: integral rows of the band hold
// sums over the band's own rows only; the band's column totals go to band_tot[b][band][cols].
__global__ void __launch_bounds__(256) integral_band_kernel(const uint8_t *base_a, const uint8_t *base_b, int split,
                                                            int64_t img_stride, int rows, int cols, int stride,
                                                            int32_t *integral, int32_t *band_tot, int n_bands,
                                                            const int64_t *__restrict__ img_off)
{
    extern __shared__ int32_t s_rows[];   // [INT_SUB][cols_pad]
    const int band = blockIdx.x, b = blockIdx.y;
    const int W = cols + 1;
    const int cols_pad = (min(cols, 4096) + 3) & ~3;
    const uint8_t *img = image_ptr(base_a, base_b, split, b, img_stride, img_off);
    int32_t *I = integral + (size_t)b * (rows + 1) * W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = band * INT_BAND;
    const int r1 = min(rows, r0 + INT_BAND);
    if (band == 0) for (int c = threadIdx.x; c < W; c += blockDim.x) I[c] = 0;   // integral row 0
    // running column sums: thread owns columns threadIdx.x + k*256
    int32_t run[16];
#pragma unroll
    for (int k = 0; k < 16; k++) run[k] = 0;
    const bool aligned4 = ((((uintptr_t)img) | (uintptr_t)stride) & 3) == 0;
    for (int cbase = 0; cbase < cols; cbase += 4096) {          // column super-chunks (cols <= 4096: one pass)
        const int ccount = min(4096, cols - cbase);
#pragma unroll
        for (int k = 0; k < 16; k++) run[k] = 0;
        for (int rs = r0; rs < r1; rs += INT_SUB) {
            const int r = rs + warp;
            if (r < r1) {
                // horizontal inclusive prefix of row r over [0, cbase+ccount), keeping only [cbase, ...) in smem
                const uint8_t *src = img + (size_t)r * stride;
                int32_t carry = 0;
                // prefix of the columns before cbase (only when cols > 4096)
                for (int c = lane * 4; c < cbase; c += 128) {
                    int32_t s = 0;
                    for (int q = 0; q < 4; q++) s += src[c + q];
                    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    carry += s;
                }
                int32_t *dst = s_rows + (size_t)warp * cols_pad;
                for (int c0 = 0; c0 < ccount; c0 += 128) {
                    const int c = c0 + lane * 4;
                    uint32_t px = 0;
                    if (c + 3 < ccount && aligned4) px = *(const uint32_t *)(src + cbase + c);
                    else {
                        for (int q = 0; q < 4; q++) if (c + q < ccount) px |= (uint32_t)src[cbase + c + q] << (8 * q);
                    }
                    int32_t p0 = px & 255, p1 = p0 + ((px >> 8) & 255), p2 = p1 + ((px >> 16) & 255), p3 = p2 + (px >> 24);
                    int32_t incl = p3;
                    for (int o = 1; o < 32; o <<= 1) {
                        int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    const int32_t excl = incl - p3 + carry;
                    if (c < ccount) {
                        int4 v = make_int4(excl + p0, excl + p1, excl + p2, excl + p3);
                        *(int4 *)(dst + c) = v;
                    }
                    carry += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
            __syncthreads();
            // vertical accumulation over the staged rows
            const int nsub = min(INT_SUB, r1 - rs);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int c = threadIdx.x + k * 256;
                if (c < ccount) {
                    int32_t acc = run[k];
                    for (int q = 0; q < nsub; q++) {
                        acc += s_rows[(size_t)q * cols_pad + c];
                        I[(size_t)(rs + q + 1) * W + cbase + c + 1] = acc;
                    }
                    run[k] = acc;
                }
            }
            if (threadIdx.x == 0 && cbase == 0) for (int q = 0; q < nsub; q++) I[(size_t)(rs + q + 1) * W] = 0;
            __syncthreads();
        }
        int32_t *bt = band_tot + ((size_t)b * n_bands + band) * cols + cbase;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int c = threadIdx.x + k * 256;
            if (c < ccount) bt[c] = run[k];
        }
    }
}

// K1b: add the totals of all previous bands.  grid (bands-1, batch): band index = blockIdx.x + 1.
__global__ void __launch_bounds__(256) integral_fix_kernel(int rows, int cols, int32_t *integral,
                                                           const int32_t *band_tot, int n_bands)
{
    const int band = blockIdx.x + 1, b = blockIdx.y;
    const int W = cols + 1;
    int32_t *I = integral + (size_t)b * (rows + 1) * W;
    const int32_t *bt = band_tot + (size_t)b * n_bands * cols;
    const int r0 = band * INT_BAND, r1 = min(rows, r0 + INT_BAND);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        int32_t off = 0;
        for (int q = 0; q < band; q++) off += bt[(size_t)q * cols + c];
        for (int r = r0; r < r1; r++) I[(size_t)(r + 1) * W + c + 1] += off;
    }
}

// ---------------------------------------------------------------- K2: Hessian layers + NMS + interpolation
__device__ __forceinline__ bool solve3(const float a[3][3], const float b[3], float x[3])
{
    float d = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1])
            - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0])
            + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (d == 0) { x[0] = x[1] = x[2] = 0; return false; }
    d = 1 / d;
    x[0] = d * (b[0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1])
              - a[0][1] * (b[1] * a[2][2] - a[1][2] * b[2])
              + a[0][2] * (b[1] * a[2][1] - a[1][1] * b[2]));
    x[1] = d * (a[0][0] * (b[1] * a[2][2] - a[1][2] * b[2])
              - b[0] * (a[1][0] * a[2][2] - a[1][2] * a[2][0])
              + a[0][2] * (a[1][0] * b[2] - b[1] * a[2][0]));
    x[2] = d * (a[0][0] * (a[1][1] * b[2] - b[1] * a[2][1])
              - a[0][1] * (a[1][0] * b[2] - b[1] * a[2][0])
              + b[0] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]));
    return true;
}

// Box responses of one sample from 32 integral reads (the 10 boxes share their corners: Dxx 2x4, Dyy 4x2, Dxy 4x4 grid).
// `o` points at the sample origin inside an int32 array (global integral or its shared-memory copy); L.off[] holds the 32
// corner offsets already combined with that array's row pitch.  With the octave a template parameter and the layer loop
// unrolled, every L.field is a compile-time address in the kernel-parameter constant bank, i.e. a free instruction operand.
// Each box sum is an exact integer; products and the double accumulation keep the CPU expression's rounding.
template <typename Ptr>
__device__ __forceinline__ void hessian_responses(Ptr o, const SurfLayer &L, float &dx, float &dy, float &dxy)
{
    {   // Dxx: rows y1, y2 x column edges x0..x3          off[0..3] = (y1, x0..x3), off[4..7] = (y2, x0..x3)
        const int t0 = o[L.off[0]], t1 = o[L.off[1]], t2 = o[L.off[2]], t3 = o[L.off[3]];
        const int b0 = o[L.off[4]], b1 = o[L.off[5]], b2 = o[L.off[6]], b3 = o[L.off[7]];
        double d = 0;
        d += (double)((float)(t0 + b1 - b0 - t1) * L.dx[0].w);
        d += (double)((float)(t1 + b2 - b1 - t2) * L.dx[1].w);
        d += (double)((float)(t2 + b3 - b2 - t3) * L.dx[2].w);
        dx = (float)d;
    }
    {   // Dyy: row edges y0..y3 x columns x1, x2            off[8..11] = (y0..y3, x1), off[12..15] = (y0..y3, x2)
        const int l0 = o[L.off[8]], l1 = o[L.off[9]], l2 = o[L.off[10]], l3 = o[L.off[11]];
        const int r0 = o[L.off[12]], r1 = o[L.off[13]], r2 = o[L.off[14]], r3 = o[L.off[15]];
        double d = 0;
        d += (double)((float)(l0 + r1 - l1 - r0) * L.dy[0].w);
        d += (double)((float)(l1 + r2 - l2 - r1) * L.dy[1].w);
        d += (double)((float)(l2 + r3 - l3 - r2) * L.dy[2].w);
        dy = (float)d;
    }
    {   // Dxy: 4x4 corner grid, row-major                    off[16 + 4*row + col], rows ya..yd, cols xa..xd
        const int aa = o[L.off[16]], ab = o[L.off[17]], ac = o[L.off[18]], ad = o[L.off[19]];
        const int ba = o[L.off[20]], bb = o[L.off[21]], bc = o[L.off[22]], bd = o[L.off[23]];
        const int ca = o[L.off[24]], cb = o[L.off[25]], cc = o[L.off[26]], cd = o[L.off[27]];
        const int da = o[L.off[28]], db = o[L.off[29]], dc = o[L.off[30]], dd = o[L.off[31]];
        double d = 0;
        d += (double)((float)(aa + bb - ba - ab) * L.dxy[0].w);
        d += (double)((float)(ac + bd - bc - ad) * L.dxy[1].w);
        d += (double)((float)(ca + db - da - cb) * L.dxy[2].w);
        d += (double)((float)(cc + dd - dc - cd) * L.dxy[3].w);
        dxy = (float)d;
    }
}

#include "hessian_tables.inc"

// Shared-memory address of box corner k relative to the sample origin.  Octave 0: plain [row][col] footprint.  Octave 1
// samples every second column, so the footprint is staged as two column-parity planes [parity][row][col / 2] and the lanes
// of a warp (consecutive samples) read consecutive words instead of every second one (2-way bank conflict on all 32 reads).
template <int OCT, int L>
struct HessAddr {
    using F = HessFixed<OCT, L>;
    static constexpr bool deint = OCT == 1;
    static __host__ __device__ constexpr int off(int k)
    {
        return deint ? (F::dx(k) & 1) * F::plane + F::dy(k) * F::pitch2 + (F::dx(k) >> 1) : F::dy(k) * F::pitch + F::dx(k);
    }
    static __host__ __device__ constexpr float w(int k) { return F::w(k); }
};

// det of one sample for the default layer structure: every corner offset and weight is an immediate
template <int OCT, int L>
__device__ __forceinline__ float det_fixed(const int32_t *o)
{
    using T = HessAddr<OCT, L>;
    float dx, dy, dxy;
    {
        const int t0 = o[T::off(0)], t1 = o[T::off(1)], t2 = o[T::off(2)], t3 = o[T::off(3)];
        const int b0 = o[T::off(4)], b1 = o[T::off(5)], b2 = o[T::off(6)], b3 = o[T::off(7)];
        double d = 0;
        d += (double)((float)(t0 + b1 - b0 - t1) * T::w(0));
        d += (double)((float)(t1 + b2 - b1 - t2) * T::w(1));
        d += (double)((float)(t2 + b3 - b2 - t3) * T::w(2));
        dx = (float)d;
    }
    {
        const int l0 = o[T::off(8)], l1 = o[T::off(9)], l2 = o[T::off(10)], l3 = o[T::off(11)];
        const int r0 = o[T::off(12)], r1 = o[T::off(13)], r2 = o[T::off(14)], r3 = o[T::off(15)];
        double d = 0;
        d += (double)((float)(l0 + r1 - l1 - r0) * T::w(3));
        d += (double)((float)(l1 + r2 - l2 - r1) * T::w(4));
        d += (double)((float)(l2 + r3 - l3 - r2) * T::w(5));
        dy = (float)d;
    }
    {
        const int aa = o[T::off(16)], ab = o[T::off(17)], ac = o[T::off(18)], ad = o[T::off(19)];
        const int ba = o[T::off(20)], bb = o[T::off(21)], bc = o[T::off(22)], bd = o[T::off(23)];
        const int ca = o[T::off(24)], cb = o[T::off(25)], cc = o[T::off(26)], cd = o[T::off(27)];
        const int da = o[T::off(28)], db = o[T::off(29)], dc = o[T::off(30)], dd = o[T::off(31)];
        double d = 0;
        d += (double)((float)(aa + bb - ba - ab) * T::w(6));
        d += (double)((float)(ac + bd - bc - ad) * T::w(7));
        d += (double)((float)(ca + db - da - cb) * T::w(8));
        d += (double)((float)(cc + dd - dc - cd) * T::w(9));
        dxy = (float)d;
    }
    float tt = 0.81f * dxy;
    tt = tt * dxy;
    return dx * dy - tt;
}

template <int OCT, int L>
__device__ __forceinline__ void det_layer_fixed(const SurfPlan &plan, const int32_t *s_int, float *s_det, int i0, int j0, int r_base, int c_base)
{
    using T = HessFixed<OCT, L>;
    constexpr int step = 1 << OCT, SW = ht_x(OCT) + 2, SH = ht_y(OCT) + 2, RW = T::pitch;
    const int ni = plan.layer[OCT][L].samples_i, nj = plan.layer[OCT][L].samples_j;
    // the haloed tile flattened over the CTA: (66 x 18 samples as rows per warp and columns per lane left a third of the lane slots idle)
    for (int idx = threadIdx.x; idx < SH * SW; idx += HT_THREADS) {
        const int ly = idx / SW, lx = idx - ly * SW;
        const int si = i0 + ly - T::margin, sj = j0 + lx - T::margin;
        float det = 0.f;
        if (si >= 0 && si < ni && sj >= 0 && sj < nj) {
            if constexpr (HessAddr<OCT, L>::deint) det = det_fixed<OCT, L>(s_int + (si * step - r_base) * T::pitch2 + sj - (c_base >> 1));   // c_base is even
            else det = det_fixed<OCT, L>(s_int + (si * step - r_base) * RW - c_base + sj * step);
        }
        s_det[L * SH * SW + idx] = det;
    }
}

// phase 2b of hessian_nms_kernel for one in-layer maximum (tile sample idx of layer l): the 26-neighbour test against the
// layers below and above, sub-sample interpolation (Brown & Lowe step of the CPU code), laplacian sign, candidate push.
template <int OCT>
__device__ __forceinline__ void nms_finish(const SurfPlan &plan, const int32_t *__restrict__ I, int W, const float *s_det, int SW, int SH,
                                           int tile_j0, int tile_i0, int TX, int idx, int l, int b, float *cand, int32_t *counters, int cand_cap)
{
    constexpr int o = OCT, step = 1 << OCT;
    const int ty = idx / TX, tx = idx - ty * TX;
    const int li = tile_i0 + ty, lj = tile_j0 + tx;
    const SurfLayer &L = plan.layer[o][l];
    const float *c1 = s_det + (l * SH + ty + 1) * SW + tx + 1;
    const float val0 = c1[0];
    const float *c0 = c1 - SH * SW, *c2 = c1 + SH * SW;
    float N9[3][9];
    const float *cs[3] = { c0, c1, c2 };
    bool is_max = true;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        const float *c = cs[q];
        N9[q][0] = c[-SW - 1]; N9[q][1] = c[-SW]; N9[q][2] = c[-SW + 1];
        N9[q][3] = c[-1];      N9[q][4] = c[0];   N9[q][5] = c[1];
        N9[q][6] = c[SW - 1];  N9[q][7] = c[SW];  N9[q][8] = c[SW + 1];
    }
#pragma unroll
    for (int q = 0; q < 3; q++)
#pragma unroll
        for (int k = 0; k < 9; k++)
            if (!(q == 1 && k == 4)) is_max = is_max && (val0 > N9[q][k]);
    if (!is_max) return;
    const int size = L.size;
    const int sum_i = step * (li - (size / 2) / step), sum_j = step * (lj - (size / 2) / step);
    float py = sum_i + (size - 1) * 0.5f, px = sum_j + (size - 1) * 0.5f, psz = (float)size;
    const int ds = size - plan.layer[o][l - 1].size;
    float bb[3] = { -(N9[1][5] - N9[1][3]) / 2, -(N9[1][7] - N9[1][1]) / 2, -(N9[2][4] - N9[0][4]) / 2 };
    float A[3][3];
    A[0][0] = N9[1][3] - 2 * N9[1][4] + N9[1][5];
    A[0][1] = (N9[1][8] - N9[1][6] - N9[1][2] + N9[1][0]) / 4;
    A[0][2] = (N9[2][5] - N9[2][3] - N9[0][5] + N9[0][3]) / 4;
    A[1][0] = A[0][1];
    A[1][1] = N9[1][1] - 2 * N9[1][4] + N9[1][7];
    A[1][2] = (N9[2][7] - N9[2][1] - N9[0][7] + N9[0][1]) / 4;
    A[2][0] = A[0][2]; A[2][1] = A[1][2];
    A[2][2] = N9[0][4] - 2 * N9[1][4] + N9[2][4];
    float x[3];
    solve3(A, bb, x);
    const bool ok = (x[0] != 0 || x[1] != 0 || x[2] != 0) && fabsf(x[0]) <= 1 && fabsf(x[1]) <= 1 && fabsf(x[2]) <= 1;
    if (!ok) return;
    px += x[0] * step;
    py += x[1] * step;
    psz = (float)__float2int_rn(psz + x[2] * ds);
    // laplacian sign: recompute trace at the maximum (rare path)
    // laplacian sign from trace = dxx + dyy, recomputed box by box from the global integral
    const int32_t *org = I + (size_t)((li - L.margin) * step) * W + (lj - L.margin) * step;
    double tdx = 0, tdy = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const HaarBox &fx = L.dx[k], &fy = L.dy[k];
        const int vx = __ldg(org + fx.y1 * W + fx.x1) + __ldg(org + fx.y2 * W + fx.x2) - __ldg(org + fx.y2 * W + fx.x1) - __ldg(org + fx.y1 * W + fx.x2);
        const int vy = __ldg(org + fy.y1 * W + fy.x1) + __ldg(org + fy.y2 * W + fy.x2) - __ldg(org + fy.y2 * W + fy.x1) - __ldg(org + fy.y1 * W + fy.x2);
        tdx += (double)((float)vx * fx.w); tdy += (double)((float)vy * fy.w);
    }
    const float trace = (float)tdx + (float)tdy;
    const int slot = atomicAdd(&counters[b * 4 + 0], 1);
    if (slot < cand_cap) {
        float4 *dst = (float4 *)(cand + ((size_t)b * cand_cap + slot) * KP_STRIDE);
        dst[0] = make_float4(px, py, psz, -1.f);
        dst[1] = make_float4(val0, (float)o, (float)((trace > 0) - (trace < 0)), 0.f);
    }
}

// 4-byte global -> shared copy
__device__ __forceinline__ void stage_copy4(int32_t *dst, const int32_t *src)
{
#ifdef __CUDACC__
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#else
    *dst = *src;        // CPU emulation build (tests/cuda_emu)
#endif
}
__device__ __forceinline__ void stage_copy_wait()
{
#ifdef __CUDACC__
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
#endif
}

// STAGE = true: the integral footprint of the tile (all layers of the octave) is copied to shared memory once and the
// 32 corner reads per sample-layer become LDS (octaves whose footprint fits: 0 and 1); false: reads go to L1/L2.
// One launch per octave (OCT is a template parameter, see hessian_responses).
template <bool STAGE, int OCT, bool FIXED = false>
__global__ void __launch_bounds__(HT_THREADS) hessian_nms_kernel(const __grid_constant__ SurfPlan plan,
                                                                 const int32_t *__restrict__ integral,
                                                                 float *cand, int32_t *counters, int cand_cap)
{
    extern __shared__ float s_dyn[];
    const int b = blockIdx.y;
    constexpr int o = OCT;
    const int t = blockIdx.x;
    const int tile_x = t % plan.tiles_x[o], tile_y = t / plan.tiles_x[o];
    constexpr int step = 1 << OCT;
    const int lrows = plan.rows / step, lcols = plan.cols / step;
    const int W = plan.cols + 1;
    const int32_t *I = integral + (size_t)b * (plan.rows + 1) * W;
    const int nl = plan.n_layers;
    constexpr int TX = ht_x(OCT), TY = ht_y(OCT), SW = TX + 2, SH = TY + 2;
    float *s_det = s_dyn;                                   // [n_layers][TY+2][TX+2]
    int *s_list = (int *)(s_dyn + nl * SW * SH);            // [0] entry count, [4 ..) queue of in-layer maxima (phase 2)
    int32_t *s_int = (int32_t *)(s_list + 4 + TX * TY);     // [stage_rows][stage_cols] (STAGE only)
    const int i0 = tile_y * TY - 1, j0 = tile_x * TX - 1;   // layer coords of the smem origin
    const int RW = plan.stage_cols[o];
    const int r_base = i0 * step + plan.stage_off_min[o], c_base = j0 * step + plan.stage_off_min[o];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) s_list[0] = 0;
    if (STAGE) {
        // global -> shared without a register round trip (cp.async, 4 bytes: the integral rows have an odd pitch); the
        // load -> store dependency of the plain copy was 18 % of this kernel's stall samples.  Tiles whose footprint lies
        // inside the image skip the coordinate clamps.
        const int RH = plan.stage_rows[o];
        const bool inside = r_base >= 0 && r_base + RH - 1 <= plan.rows && c_base >= 0 && c_base + RW - 1 <= plan.cols;
        for (int rr = warp; rr < RH; rr += HT_THREADS / 32) {           // one warp per footprint row: no index division
            const int32_t *src = I + (size_t)min(max(r_base + rr, 0), plan.rows) * W;
            int32_t *dst = s_int + rr * RW;
            if constexpr (FIXED && OCT == 1) {              // column-parity planes (HessAddr)
                using F = HessFixed<1, 0>;
                dst = s_int + rr * F::pitch2;
                for (int cc = lane; cc < RW; cc += 32)
                    stage_copy4(dst + (cc & 1) * F::plane + (cc >> 1), src + (inside ? c_base + cc : min(max(c_base + cc, 0), plan.cols)));
            } else if (inside) for (int cc = lane; cc < RW; cc += 32) stage_copy4(dst + cc, src + c_base + cc);
            else for (int cc = lane; cc < RW; cc += 32) stage_copy4(dst + cc, src + min(max(c_base + cc, 0), plan.cols));
        }
        stage_copy_wait();
        __syncthreads();
    }

    // phase 1: det for every layer over the haloed tile; warp w takes rows w, w+8, ..., lanes take columns
    if constexpr (FIXED) {
        // default layer structure: geometry from compile-time tables (verified against the plan by the host)
        det_layer_fixed<OCT < 2 ? OCT : 0, 0>(plan, s_int, s_det, i0, j0, r_base, c_base);
        det_layer_fixed<OCT < 2 ? OCT : 0, 1>(plan, s_int, s_det, i0, j0, r_base, c_base);
        det_layer_fixed<OCT < 2 ? OCT : 0, 2>(plan, s_int, s_det, i0, j0, r_base, c_base);
        det_layer_fixed<OCT < 2 ? OCT : 0, 3>(plan, s_int, s_det, i0, j0, r_base, c_base);
        det_layer_fixed<OCT < 2 ? OCT : 0, 4>(plan, s_int, s_det, i0, j0, r_base, c_base);
    } else
    for (int idx = threadIdx.x; idx < SH * SW; idx += HT_THREADS) {           // the haloed tile flattened over the CTA
        const int ly = idx / SW, lx = idx - ly * SW;
        const int li = i0 + ly, lj = j0 + lx;
        {
#pragma unroll
            for (int l = 0; l < VFSMS_MAX_LAYERS_PER_OCTAVE; l++) {
                if (l < nl) {
                    const SurfLayer &L = plan.layer[OCT][l];
                    const int si = li - L.margin, sj = lj - L.margin;
                    float det = 0.f;
                    if (si >= 0 && si < L.samples_i && sj >= 0 && sj < L.samples_j) {
                        float dx, dy, dxy;
                        if (STAGE) hessian_responses(s_int + (si * step - r_base) * RW + (sj * step - c_base), L, dx, dy, dxy);
                        else hessian_responses(I + (size_t)(si * step) * W + sj * step, L, dx, dy, dxy);
                        float tt = 0.81f * dxy;
                        tt = tt * dxy;
                        det = dx * dy - tt;
                    }
                    s_det[(l * SH + ly) * SW + lx] = det;
                }
            }
        }
    }
    __syncthreads();

    // phase 2a: threshold and 3x3 maximum inside the own layer.  A warp runs the whole neighbourhood test as soon as one
    // lane needs it, so the cheap in-layer test runs here and the few survivors (in-layer maxima above the threshold) are
    // queued; phase 2b gives every queued sample its own lane for the two neighbour layers, the interpolation and the push.
    for (int idx = threadIdx.x; idx < TX * TY; idx += HT_THREADS) {
        const int ty = idx / TX, tx = idx - ty * TX;
        const int li = tile_y * TY + ty, lj = tile_x * TX + tx;
        if (li >= lrows || lj >= lcols) continue;
        for (int l = 1; l < nl - 1; l++) {
            const float *c = s_det + (l * SH + ty + 1) * SW + tx + 1;
            const float val0 = c[0];
            if (!(val0 > plan.threshold)) continue;                      // cheapest test first: most samples stop here
            const int margin = (plan.layer[o][l + 1].size / 2) / step + 1;
            if (li < margin || li >= lrows - margin || lj < margin || lj >= lcols - margin) continue;
            if (!(val0 > c[-SW - 1] && val0 > c[-SW] && val0 > c[-SW + 1] && val0 > c[-1] && val0 > c[1] &&
                  val0 > c[SW - 1] && val0 > c[SW] && val0 > c[SW + 1])) continue;
            const int slot = atomicAdd(s_list, 1);
            if (slot < TX * TY) s_list[4 + slot] = idx | (l << 16);
            else nms_finish<OCT>(plan, I, W, s_det, SW, SH, tile_x * TX, tile_y * TY, TX, idx, l, b, cand, counters, cand_cap);    // queue full
        }
    }
    __syncthreads();
    const int n_list = min(s_list[0], TX * TY);
    for (int e = threadIdx.x; e < n_list; e += HT_THREADS) {
        const int v = s_list[4 + e];
        nms_finish<OCT>(plan, I, W, s_det, SW, SH, tile_x * TX, tile_y * TY, TX, v & 0xffff, v >> 16, b, cand, counters, cand_cap);
    }
}

// ---------------------------------------------------------------- K3: response ordering by rank counting
// OpenCV's KeypointGreater: response desc, size desc, octave desc, y desc, x asc.
struct SortKey { unsigned long long k1, k2; };

__device__ __forceinline__ unsigned f2ord(float f)   // order-preserving float -> uint
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ SortKey make_key(const float *kp)
{
    SortKey k;
    // size is integral (< 2^24) and octave < 256: both exact in the low word
    const unsigned sz = (unsigned)kp[KP_SIZE], oc = (unsigned)kp[KP_OCTAVE];
    k.k1 = ((unsigned long long)f2ord(kp[KP_RESPONSE]) << 32) | ((unsigned long long)(sz & 0xffffffu) << 8) | (oc & 0xffu);
    k.k2 = ((unsigned long long)f2ord(kp[KP_Y]) << 32) | (unsigned long long)(~f2ord(kp[KP_X]));
    return k;
}

// a sorts before b ?
__device__ __forceinline__ bool key_before(const SortKey &a, const SortKey &b)
{
    return a.k1 > b.k1 || (a.k1 == b.k1 && a.k2 > b.k2);
}

// K3a: when only the strongest max_features survive, most candidates never need a rank.  Histogram the responses on
// their top 11 bits (positive floats order like their bit patterns), pick the bin holding the max_features-th
// largest, and copy everything at or above that bin (K + at most one bin) into the staging buffer.
#define RH_BINS 2048
__global__ void __launch_bounds__(256) response_hist_kernel(const float *__restrict__ cand, const int32_t *__restrict__ counters, int cand_cap,
                                                            int *hist)
{
    __shared__ int s_h[RH_BINS];          // responses of one image crowd into a few exponent bins: privatise per CTA
    const int b = blockIdx.y;
    const int n = min(counters[b * 4 + 0], cand_cap);
    const float *C = cand + (size_t)b * cand_cap * KP_STRIDE;
    for (int i = threadIdx.x; i < RH_BINS; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&s_h[__float_as_uint(C[(size_t)i * KP_STRIDE + KP_RESPONSE]) >> 20], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < RH_BINS; i += blockDim.x) if (s_h[i]) atomicAdd(&hist[b * RH_BINS + i], s_h[i]);
}

// one warp per image: smallest bin index whose suffix count reaches max_features (0 = keep everything)
__global__ void response_threshold_kernel(const int *__restrict__ hist, int32_t *counters, int cand_cap, int max_features, int batch, unsigned *thr_bits)
{
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= batch) return;
    const int n = min(counters[b * 4 + 0], cand_cap);
    unsigned thr = 0;
    if (max_features > 0 && n > max_features) {
        int cum = 0, found = -1;
        for (int base = RH_BINS - 32; base >= 0 && found < 0; base -= 32) {
            const int v = hist[b * RH_BINS + base + lane];
            // suffix sums within the 32-bin group, high bins first
            int suf = v;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += t; }
            const unsigned hit = __ballot_sync(0xffffffffu, cum + suf >= max_features);
            if (hit) found = base + 31 - __clz(hit);          // highest bin of the group that already reaches the target
            cum += __shfl_sync(0xffffffffu, suf, 0);
        }
        thr = found < 0 ? 0u : ((unsigned)found << 20);
    }
    if (lane == 0) thr_bits[b] = thr;
}

__global__ void __launch_bounds__(256) response_filter_kernel(const float *__restrict__ cand, float *staged, int32_t *counters, int cand_cap,
                                                              const unsigned *__restrict__ thr_bits)
{
    const int b = blockIdx.y;
    const int n = min(counters[b * 4 + 0], cand_cap);
    const unsigned thr = thr_bits[b];
    const float *C = cand + (size_t)b * cand_cap * KP_STRIDE;
    float *S = staged + (size_t)b * cand_cap * KP_STRIDE;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (__float_as_uint(C[(size_t)i * KP_STRIDE + KP_RESPONSE]) < thr) continue;
        const int slot = atomicAdd(&counters[b * 4 + 1], 1);
        const float4 *src = (const float4 *)(C + (size_t)i * KP_STRIDE);
        float4 *dst = (float4 *)(S + (size_t)slot * KP_STRIDE);
        dst[0] = src[0]; dst[1] = src[1];
    }
}

// K3: rank of every staged candidate under KeypointGreater (any arrival order in -> the same order out).
// src = staged [n_staged = counters[1]], dst = ordered; n_keep written back to counters[1].
__global__ void __launch_bounds__(256) rank_sort_kernel(const float *__restrict__ src, float *dst, int32_t *counters,
                                                        int cand_cap, int batch, int max_features)
{
    __shared__ unsigned s_r[2048];                 // order-preserving response bits: the hot loop compares 32-bit words only
    const int chunks_per_img = ceil_div(cand_cap, 256);
    for (int item = blockIdx.x; item < batch * chunks_per_img; item += gridDim.x) {
        const int b = item / chunks_per_img, chunk = item - b * chunks_per_img;
        const int n_raw = counters[b * 4 + 0];
        const int n = counters[b * 4 + 1];
        if (chunk * 256 >= n) continue;             // uniform per CTA
        const float *C = src + (size_t)b * cand_cap * KP_STRIDE;
        const int i = chunk * 256 + threadIdx.x;
        SortKey mine; mine.k1 = 0; mine.k2 = 0;
        if (i < n) mine = make_key(C + (size_t)i * KP_STRIDE);
        int rank = 0;
        const unsigned my_r = (unsigned)(mine.k1 >> 32);
        for (int base = 0; base < n; base += 2048) {
            __syncthreads();
            for (int q = threadIdx.x; q < 2048; q += 256) {
                const int j = base + q;
                s_r[q] = j < n ? f2ord(C[(size_t)j * KP_STRIDE + KP_RESPONSE]) : 0u;
            }
            __syncthreads();
            const int m = min(2048, n - base);
            if (i < n) {
#pragma unroll 8
                for (int q = 0; q < m; q++) {
                    const unsigned k = s_r[q];
                    rank += k > my_r ? 1 : 0;
                    if (k == my_r && base + q != i) {
                        // equal response (rare): the full KeypointGreater order -- size, octave, y desc, x asc -- then arrival index
                        const int j = base + q;
                        const SortKey kj = make_key(C + (size_t)j * KP_STRIDE);
                        if (kj.k1 > mine.k1 || (kj.k1 == mine.k1 && (kj.k2 > mine.k2 || (kj.k2 == mine.k2 && j < i)))) rank++;
                    }
                }
            }
        }
        const int n_keep = (max_features > 0) ? min(n, max_features) : n;
        if (i < n && rank < n_keep) {
            const float4 *s4 = (const float4 *)(C + (size_t)i * KP_STRIDE);
            float4 *d4 = (float4 *)(dst + ((size_t)b * cand_cap + rank) * KP_STRIDE);
            d4[0] = s4[0]; d4[1] = s4[1];
        }
        __syncthreads();
        if (chunk == 0 && threadIdx.x == 0 && n_raw > cand_cap) counters[b * 4 + 3] |= 1;
    }
}

// after every chunk of an image has ranked: publish n_keep (separate tiny kernel: chunks of one image run in different CTAs)
__global__ void rank_finalize_kernel(int32_t *counters, int batch, int max_features)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const int n = counters[b * 4 + 1];
    counters[b * 4 + 1] = (max_features > 0) ? min(n, max_features) : n;
}

// ---------------------------------------------------------------- K3 (sort mode 1): rank inside response bins
// Same output as rank_sort_kernel with ~n^2 / (occupied bins) comparisons instead of n^2: a 13-bit histogram of the response
// words (sign, exponent, 4 mantissa bits -- 1/16-octave bins) gives every bin its slot range in descending order, the
// candidates are scattered into their bin's range, and each candidate only counts the members of its own bin that sort
// before it.  Needs positive responses (bit patterns ordered like the values): the host keeps mode 0 for thresholds < 0.
#define RB_BITS 13
#define RB_BINS (1 << RB_BITS)
#define RB_SHIFT (32 - RB_BITS)

__global__ void __launch_bounds__(256) bin_hist_kernel(const float *__restrict__ cand, const int32_t *__restrict__ counters, int cand_cap, int *hist)
{
    __shared__ int s_h[RB_BINS];
    const int b = blockIdx.y;
    const int n = min(counters[b * 4 + 0], cand_cap);
    const float *C = cand + (size_t)b * cand_cap * KP_STRIDE;
    for (int i = threadIdx.x; i < RB_BINS; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&s_h[__float_as_uint(C[(size_t)i * KP_STRIDE + KP_RESPONSE]) >> RB_SHIFT], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < RB_BINS; i += blockDim.x) if (s_h[i]) atomicAdd(&hist[b * RB_BINS + i], s_h[i]);
}

// One warp per image, bins walked from the top: start[bin] = number of candidates in higher bins, fill[bin] = 0, the cut bin
// for max_features (as response_threshold_kernel) and the number of staged candidates (counters[1]).
__global__ void bin_offsets_kernel(const int *__restrict__ hist, int32_t *counters, int cand_cap, int max_features, int batch,
                                   int *start, int *fill, unsigned *thr_bits)
{
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= batch) return;
    const int n = min(counters[b * 4 + 0], cand_cap);
    const bool cut = max_features > 0 && n > max_features;
    int cum = 0, found = -1, n_staged = n;
    for (int base = RB_BINS - 32; base >= 0 && found < 0; base -= 32) {
        const int v = hist[b * RB_BINS + base + lane];
        int suf = v;                                              // candidates in bins base+lane .. base+31
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += t; }
        start[b * RB_BINS + base + lane] = cum + suf - v;
        fill[b * RB_BINS + base + lane] = 0;
        if (cut) {
            const unsigned hit = __ballot_sync(0xffffffffu, cum + suf >= max_features);
            if (hit) {
                const int l = 31 - __clz(hit);                    // highest bin of the group that reaches the target
                found = base + l;
                n_staged = cum + __shfl_sync(0xffffffffu, suf, l);
            }
        }
        cum += __shfl_sync(0xffffffffu, suf, 0);
    }
    if (lane == 0) {
        thr_bits[b] = found < 0 ? 0u : ((unsigned)found << RB_SHIFT);
        counters[b * 4 + 1] = n_staged;
    }
}

__global__ void __launch_bounds__(256) bin_scatter_kernel(const float *__restrict__ cand, float *staged, const int32_t *__restrict__ counters,
                                                          int cand_cap, const unsigned *__restrict__ thr_bits,
                                                          const int *__restrict__ start, int *fill)
{
    const int b = blockIdx.y;
    const int n = min(counters[b * 4 + 0], cand_cap);
    const unsigned thr = thr_bits[b];
    const float *C = cand + (size_t)b * cand_cap * KP_STRIDE;
    float *S = staged + (size_t)b * cand_cap * KP_STRIDE;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned bits = __float_as_uint(C[(size_t)i * KP_STRIDE + KP_RESPONSE]);
        if (bits < thr) continue;
        const int bin = bits >> RB_SHIFT;
        const int slot = start[b * RB_BINS + bin] + atomicAdd(&fill[b * RB_BINS + bin], 1);
        const float4 *src = (const float4 *)(C + (size_t)i * KP_STRIDE);
        float4 *dst = (float4 *)(S + (size_t)slot * KP_STRIDE);
        dst[0] = src[0]; dst[1] = src[1];
    }
}

__global__ void __launch_bounds__(256) bin_rank_kernel(const float *__restrict__ src, float *dst, int32_t *counters, int cand_cap, int batch,
                                                       int max_features, const int *__restrict__ hist, const int *__restrict__ start)
{
    const int chunks_per_img = ceil_div(cand_cap, 256);
    for (int item = blockIdx.x; item < batch * chunks_per_img; item += gridDim.x) {
        const int b = item / chunks_per_img, chunk = item - b * chunks_per_img;
        const int n_raw = counters[b * 4 + 0];
        const int n = counters[b * 4 + 1];
        if (chunk * 256 >= n) continue;
        const float *C = src + (size_t)b * cand_cap * KP_STRIDE;
        const int i = chunk * 256 + threadIdx.x;
        if (i < n) {
            const SortKey mine = make_key(C + (size_t)i * KP_STRIDE);
            const unsigned my_r = __float_as_uint(C[(size_t)i * KP_STRIDE + KP_RESPONSE]);
            const int bin = my_r >> RB_SHIFT;
            const int s = start[b * RB_BINS + bin], e = s + hist[b * RB_BINS + bin];
            int rank = s;                                          // everything in higher bins sorts before this candidate
            for (int j = s; j < e; j++) {
                const unsigned k = __float_as_uint(__ldg(C + (size_t)j * KP_STRIDE + KP_RESPONSE));
                if (k > my_r) rank++;
                else if (k == my_r && j != i) {
                    const SortKey kj = make_key(C + (size_t)j * KP_STRIDE);
                    if (kj.k1 > mine.k1 || (kj.k1 == mine.k1 && (kj.k2 > mine.k2 || (kj.k2 == mine.k2 && j < i)))) rank++;
                }
            }
            const int n_keep = (max_features > 0) ? min(n, max_features) : n;
            if (rank < n_keep) {
                const float4 *s4 = (const float4 *)(C + (size_t)i * KP_STRIDE);
                float4 *d4 = (float4 *)(dst + ((size_t)b * cand_cap + rank) * KP_STRIDE);
                d4[0] = s4[0]; d4[1] = s4[1];
            }
        }
        if (chunk == 0 && threadIdx.x == 0 && n_raw > cand_cap) counters[b * 4 + 3] |= 1;
    }
}

// ---------------------------------------------------------------- K3b: drop keypoints that cannot be oriented, keep order
__device__ __forceinline__ bool keypoint_valid(const float *kp, int rows, int cols, int upright)
{
    const int srows = rows + 1, scols = cols + 1;
    const float size = kp[KP_SIZE];
    const float s = size * 1.2f / 9.0f;
    const int gws = 2 * __float2int_rn(2 * s);
    if (srows < gws || scols < gws) return false;
    if (upright) return true;
    const float cx = kp[KP_X], cy = kp[KP_Y];
    const float half = (float)(gws - 1) / 2;
    for (int kk = 0; kk < ORI_SAMPLES; kk++) {
        const int x = __float2int_rn(cx + c_apt_x[kk] * s - half);
        const int y = __float2int_rn(cy + c_apt_y[kk] * s - half);
        if (y < 0 || y >= srows - gws || x < 0 || x >= scols - gws) continue;
        return true;
    }
    return false;
}

// one CTA (1024 threads) per image
__global__ void __launch_bounds__(1024) validate_compact_kernel(const float *__restrict__ sorted, float *kp_out,
                                                                int32_t *counters, int cand_cap, int kp_cap,
                                                                int rows, int cols, int upright)
{
    __shared__ int s_warp[32];
    __shared__ int s_base, s_total;
    const int b = blockIdx.x;
    const int n = counters[b * 4 + 1];
    const float *S = sorted + (size_t)b * cand_cap * KP_STRIDE;
    float *K = kp_out + (size_t)b * kp_cap * KP_STRIDE;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const bool v = (i < n) && keypoint_valid(S + (size_t)i * KP_STRIDE, rows, cols, upright);
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            const int c = s_warp[lane];
            int incl = c;
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            s_warp[lane] = incl - c;
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
        const int pos = s_base + s_warp[warp] + __popc(bal & ((1u << lane) - 1));
        if (v && pos < kp_cap) {
            const float4 *src = (const float4 *)(S + (size_t)i * KP_STRIDE);
            float4 *dst = (float4 *)(K + (size_t)pos * KP_STRIDE);
            dst[0] = src[0]; dst[1] = src[1];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int m = s_base;
        if (m > kp_cap) { counters[b * 4 + 3] |= 2; m = kp_cap; }
        counters[b * 4 + 2] = m;
    }
}

// prefix over images of n_final -> work list for the describe kernel
__global__ void prefix_kernel(const int32_t *counters, int32_t *prefix, int batch)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < batch; b++) { prefix[b] = acc; acc += counters[b * 4 + 2]; }
        prefix[batch] = acc;
    }
}

// ---------------------------------------------------------------- K4: orientation + descriptor, one CTA per keypoint
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

__device__ __forceinline__ int box_sum(const int32_t *__restrict__ o, int W, int x1, int y1, int x2, int y2)
{
    return __ldg(o + y1 * W + x1) + __ldg(o + y2 * W + x2) - __ldg(o + y2 * W + x1) - __ldg(o + y1 * W + x2);
}

#include "surf_describe.cuh"

// ---------------------------------------------------------------- K4c: one CTA per keypoint (windows > WK_MAX_WIN only)
__global__ void __launch_bounds__(DESC_THREADS) orient_describe_kernel(
    const uint8_t *base_a, const uint8_t *base_b, int split, int64_t img_stride, int rows, int cols, int stride,
    const int32_t *__restrict__ integral, float *kp_all, float *desc_all, const int32_t *__restrict__ prefix,
    int batch, int kp_cap, int extended, int upright, const int *big_flag, const int64_t *__restrict__ img_off)
{
    if (*big_flag == 0) return;        // no window exceeded WK_MAX_WIN (always the case for n_octaves <= 4)
    __shared__ float s_X[ORI_SAMPLES], s_Y[ORI_SAMPLES], s_ang[ORI_SAMPLES];
    __shared__ int s_flag_cnt[4];
    __shared__ float s_sumx[72], s_sumy[72];
    __shared__ float s_startx[MAX_WIN], s_starty[MAX_WIN];
    __shared__ uint8_t s_patch[(PATCH_SZ + 1) * (PATCH_SZ + 1) + 3];
    __shared__ float s_DX[PATCH_SZ * PATCH_SZ], s_DY[PATCH_SZ * PATCH_SZ];
    __shared__ float s_vec[128];
    __shared__ float s_dir, s_scale;
    __shared__ int s_nangle;

    const int total = prefix[batch];
    const int W = cols + 1, srows = rows + 1, scols = cols + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int dsize = extended ? 128 : 64;

    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        // (image, keypoint) from the flattened index
        int lo = 0, hi = batch;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (prefix[mid] <= item) lo = mid; else hi = mid; }
        const int b = lo, k = item - prefix[lo];
        float *kp = kp_all + ((size_t)b * kp_cap + k) * KP_STRIDE;
        const uint8_t *img = image_ptr(base_a, base_b, split, b, img_stride, img_off);
        const int32_t *I = integral + (size_t)b * srows * W;
        const float size = kp[KP_SIZE], cx = kp[KP_X], cy = kp[KP_Y];
        const float s = size * 1.2f / 9.0f;
        const int gws = 2 * __float2int_rn(2 * s);
        if ((int)((PATCH_SZ + 1) * s) <= WK_MAX_WIN) continue;   // small windows: orient_describe_warp_kernel (CTA-uniform)

        float descriptor_dir = 360.f - 90.f;
        if (!upright) {
            // --- 113 Haar samples on a disc of radius 6s, compacted in table order
            const int h2 = __float2int_rn(((float)gws / 4) * 2);   // cvRound(ratio*2), ratio = gws/4
            const int h4 = __float2int_rn(((float)gws / 4) * 4);
            const float wgt = 1.f / ((float)(h2) * (float)(h4));    // |w| of both half boxes: 1/((dx2-dx1)*(dy2-dy1))
            bool have = false; float vX = 0, vY = 0;
            if (tid < ORI_SAMPLES) {
                const float half = (float)(gws - 1) / 2;
                const int x = __float2int_rn(cx + c_apt_x[tid] * s - half);
                const int y = __float2int_rn(cy + c_apt_y[tid] * s - half);
                if (!(y < 0 || y >= srows - gws || x < 0 || x >= scols - gws)) {
                    const int32_t *o = I + (size_t)y * W + x;
                    // dx pattern {0,0,2,4,-1},{2,0,4,4,+1}; dy pattern {0,0,4,2,+1},{0,2,4,4,-1}
                    const int bl = box_sum(o, W, 0, 0, h2, h4), br = box_sum(o, W, h2, 0, h4, h4);
                    const int bt = box_sum(o, W, 0, 0, h4, h2), bb = box_sum(o, W, 0, h2, h4, h4);
                    double d = 0; d += (double)((float)bl * (-wgt)); d += (double)((float)br * wgt);
                    const float vx = (float)d;
                    d = 0; d += (double)((float)bt * wgt); d += (double)((float)bb * (-wgt));
                    const float vy = (float)d;
                    vX = vx * c_aptw[tid]; vY = vy * c_aptw[tid];
                    have = true;
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, have);
            if (lane == 0 && warp < 4) s_flag_cnt[warp] = __popc(bal);
            __syncthreads();
            if (tid < 128) {
                int off = 0;
                for (int q = 0; q < warp; q++) off += s_flag_cnt[q];
                if (have) {
                    const int pos = off + __popc(bal & ((1u << lane) - 1));
                    s_X[pos] = vX; s_Y[pos] = vY; s_ang[pos] = fast_atan2_deg(vY, vX);
                }
            }
            if (tid == 0) s_nangle = s_flag_cnt[0] + s_flag_cnt[1] + s_flag_cnt[2] + s_flag_cnt[3];
            __syncthreads();
            const int nangle = s_nangle;
            // --- 72 sliding windows of 60 degrees, each summed serially in sample order (CPU summation order)
            if (tid < 72) {
                const int i = tid * 5;
                float sumx = 0, sumy = 0;
                for (int j = 0; j < nangle; j++) {
                    const int d = abs(__float2int_rn(s_ang[j]) - i);
                    if (d < ORI_WIN / 2 || d > 360 - ORI_WIN / 2) { sumx += s_X[j]; sumy += s_Y[j]; }
                }
                s_sumx[tid] = sumx; s_sumy[tid] = sumy;
            }
            __syncthreads();
            if (tid == 0) {
                float bestx = 0, besty = 0, best = 0;
                for (int q = 0; q < 72; q++) {
                    const float m = s_sumx[q] * s_sumx[q] + s_sumy[q] * s_sumy[q];
                    if (m > best) { best = m; bestx = s_sumx[q]; besty = s_sumy[q]; }
                }
                s_dir = fast_atan2_deg(-besty, bestx);
            }
            __syncthreads();
            descriptor_dir = s_dir;
        }
        if (tid == 0) kp[KP_ANGLE] = descriptor_dir;

        // --- 20s window -> 21x21 patch (INTER_AREA), sampled on the fly
        const int win = (int)((PATCH_SZ + 1) * s);
        const int ncols1 = cols - 1, nrows1 = rows - 1;
        float sin_dir = 0, cos_dir = 0;
        int ustart_x = 0, ustart_y = 0;
        if (!upright) {
            const float dir_rad = descriptor_dir * (float)(M_PI / 180);
            sin_dir = -(float)sin((double)dir_rad);
            cos_dir = (float)cos((double)dir_rad);
            const float win_offset = -(float)(win - 1) / 2;
            if (tid == 0) {
                float sx = cx + win_offset * cos_dir + win_offset * sin_dir;
                for (int i = 0; i < win && i < MAX_WIN; i++) { s_startx[i] = sx; sx += sin_dir; }
            } else if (tid == 32) {
                float sy = cy - win_offset * sin_dir + win_offset * cos_dir;
                for (int i = 0; i < win && i < MAX_WIN; i++) { s_starty[i] = sy; sy += cos_dir; }
            }
        } else {
            const float win_offset = -(float)(win - 1) / 2;
            ustart_x = __float2int_rn(cx + win_offset);
            ustart_y = __float2int_rn(cy - win_offset);
        }
        __syncthreads();

        // WIN(i, j): row i, column j of the window
        auto WINPIX = [&](int i, int j) -> int {
            if (!upright) {
                const double px = (double)s_startx[i] + (double)j * (double)cos_dir;
                const double py = (double)s_starty[i] - (double)j * (double)sin_dir;
                return window_pixel(img, stride, ncols1, nrows1, px, py);
            }
            int x = ustart_x + i, y = ustart_y - j;
            x = min(max(x, 0), cols - 1); y = min(max(y, 0), rows - 1);
            return img[(size_t)y * stride + x];
        };

        constexpr int PD = PATCH_SZ + 1;
        const double inv_scale = (double)PD / win;
        const double scale = 1. / inv_scale;
        const int iscale = __double2int_rn(scale);
        const bool area_fast = fabs(scale - iscale) < DBL_EPSILON;
        for (int p = tid; p < PD * PD; p += DESC_THREADS) {
            const int dy = p / PD, dx = p - dy * PD;
            int out;
            if (win == PD) out = WINPIX(dy, dx);
            else if (area_fast) {
                int sum = 0;
                for (int yy = 0; yy < iscale; yy++)
                    for (int xx = 0; xx < iscale; xx++) sum += WINPIX(dy * iscale + yy, dx * iscale + xx);
                if (iscale == 2) out = (sum + 2) >> 2;
                else { const float fs = 1.f / (float)(iscale * iscale); out = min(max(__float2int_rn(sum * fs), 0), 255); }
            } else {
                // decimation tables of this output pixel, generated in table order
                double fsx1 = dx * scale, fsx2 = fsx1 + scale, cwx = fmin(scale, win - fsx1);
                int sx1 = __double2int_ru(fsx1), sx2 = __double2int_rd(fsx2);
                sx2 = min(sx2, win - 1); sx1 = min(sx1, sx2);
                double fsy1 = dy * scale, fsy2 = fsy1 + scale, cwy = fmin(scale, win - fsy1);
                int sy1 = __double2int_ru(fsy1), sy2 = __double2int_rd(fsy2);
                sy2 = min(sy2, win - 1); sy1 = min(sy1, sy2);
                const bool xl = (sx1 - fsx1 > 1e-3), xr = (fsx2 - sx2 > 1e-3);
                const bool yl = (sy1 - fsy1 > 1e-3), yr = (fsy2 - sy2 > 1e-3);
                const float axl = (float)((sx1 - fsx1) / cwx), axm = (float)(1.0 / cwx),
                            axr = (float)(fmin(fmin(fsx2 - sx2, 1.), cwx) / cwx);
                const float ayl = (float)((sy1 - fsy1) / cwy), aym = (float)(1.0 / cwy),
                            ayr = (float)(fmin(fmin(fsy2 - sy2, 1.), cwy) / cwy);
                float sum = 0; bool first = true;
                const int ya = yl ? sy1 - 1 : sy1, yb = yr ? sy2 : sy2 - 1;     // inclusive source row range
                for (int sy = ya; sy <= yb; sy++) {
                    const float beta = (yl && sy == sy1 - 1) ? ayl : ((yr && sy == sy2) ? ayr : aym);
                    float buf = 0;
                    if (xl) buf += (float)WINPIX(sy, sx1 - 1) * axl;
                    for (int sx = sx1; sx < sx2; sx++) buf += (float)WINPIX(sy, sx) * axm;
                    if (xr) buf += (float)WINPIX(sy, sx2) * axr;
                    if (first) { sum = beta * buf; first = false; } else sum += beta * buf;
                }
                out = min(max(__float2int_rn(sum), 0), 255);
            }
            s_patch[p] = (uint8_t)out;
        }
        __syncthreads();

        // --- gradients with wavelets of size 2s, Gaussian weighted
        for (int p = tid; p < PATCH_SZ * PATCH_SZ; p += DESC_THREADS) {
            const int i = p / PATCH_SZ, j = p - i * PATCH_SZ;
            const float dw = c_DW[p];
            const int p00 = s_patch[i * PD + j], p01 = s_patch[i * PD + j + 1];
            const int p10 = s_patch[(i + 1) * PD + j], p11 = s_patch[(i + 1) * PD + j + 1];
            s_DX[p] = (float)(p01 - p00 + p11 - p10) * dw;
            s_DY[p] = (float)(p10 - p00 + p11 - p01) * dw;
        }
        __syncthreads();

        // --- 4x4 cells x (4|8) bins, each bin summed serially in the CPU order
        if (tid < dsize) {
            const int nb = extended ? 8 : 4;
            const int cell = tid / nb, bin = tid - cell * nb;
            const int ci = cell >> 2, cj = cell & 3;
            float acc = 0;
            for (int y = ci * 5; y < ci * 5 + 5; y++)
                for (int x = cj * 5; x < cj * 5 + 5; x++) {
                    const float tx = s_DX[y * PATCH_SZ + x], ty = s_DY[y * PATCH_SZ + x];
                    if (extended) {
                        switch (bin) {
                        case 0: if (ty >= 0) acc += tx; break;
                        case 1: if (ty >= 0) acc += fabsf(tx); break;
                        case 2: if (!(ty >= 0)) acc += tx; break;
                        case 3: if (!(ty >= 0)) acc += fabsf(tx); break;
                        case 4: if (tx >= 0) acc += ty; break;
                        case 5: if (tx >= 0) acc += fabsf(ty); break;
                        case 6: if (!(tx >= 0)) acc += ty; break;
                        default: if (!(tx >= 0)) acc += fabsf(ty); break;
                        }
                    } else {
                        switch (bin) {
                        case 0: acc += tx; break;
                        case 1: acc += ty; break;
                        case 2: acc += fabsf(tx); break;
                        default: acc += fabsf(ty); break;
                        }
                    }
                }
            s_vec[tid] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            double sq = 0;
            for (int q = 0; q < dsize; q++) sq += (double)(s_vec[q] * s_vec[q]);
            s_scale = (float)(1. / (sqrt(sq) + (double)FLT_EPSILON));
        }
        __syncthreads();
        if (tid < dsize) desc_all[((size_t)b * kp_cap + k) * dsize + tid] = s_vec[tid] * s_scale;
        __syncthreads();
    }
}

// ---------------------------------------------------------------- host: texture objects over the float copy of the images
// Pitch-2D float textures (point sampled, clamped, used only through tex2Dgather).  Cached by (ptr, shape, pitch).
struct TexKey { const void *p; int rows, cols, stride, linear; bool operator<(const TexKey &o) const {
    return std::tie(p, rows, cols, stride, linear) < std::tie(o.p, o.rows, o.cols, o.stride, o.linear); } };   // stride: pitch in floats
struct TexCache { std::map<TexKey, cudaTextureObject_t> m; int align = 512, pitch_align = 32; bool init = false; };

void surf_tex_destroy(vfsms_ctx *ctx)
{
    TexCache *tc = (TexCache *)ctx->tex_cache;
    if (!tc) return;
    for (auto &kv : tc->m) cudaDestroyTextureObject(kv.second);
    delete tc;
    ctx->tex_cache = nullptr;
}

// One texture over `n_img` consecutive images of the float copy, starting at image `b0` (image b at texture rows
// [(b - b0) * rows, ...)).  Cached by (pointer, shape).
static bool surf_stack_texture(vfsms_ctx *ctx, int b0, int n_img, int rows, int cols, int pitch_f, cudaStream_t st, cudaTextureObject_t *out,
                               bool linear = false)
{
    if (!ctx->tex_cache) ctx->tex_cache = new TexCache();
    TexCache *tc = (TexCache *)ctx->tex_cache;
    const float *p = ctx->surf.img_f32.as<float>() + (size_t)b0 * rows * pitch_f;
    TexKey key{p, n_img * rows, cols, pitch_f, linear ? 1 : 0};
    auto it = tc->m.find(key);
    if (it == tc->m.end()) {
        if (tc->m.size() > 8192) {
            cudaStreamSynchronize(st);
            for (auto &kv : tc->m) cudaDestroyTextureObject(kv.second);
            tc->m.clear();
        }
        cudaResourceDesc rd; memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypePitch2D;
        rd.res.pitch2D.devPtr = (void *)p;
        rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
        rd.res.pitch2D.width = cols; rd.res.pitch2D.height = (size_t)n_img * rows; rd.res.pitch2D.pitchInBytes = (size_t)pitch_f * 4;
        cudaTextureDesc td; memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = linear ? cudaFilterModeLinear : cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
        cudaTextureObject_t t = 0;
        if (cudaCreateTextureObject(&t, &rd, &td, nullptr) != cudaSuccess) { cudaGetLastError(); return false; }
        it = tc->m.emplace(key, t).first;
    }
    *out = it->second;
    return true;
}

// ---------------------------------------------------------------- host driver
int surf_reserve(vfsms_ctx *ctx, int batch, int rows, int cols, const vfsms_surf_params *p)
{
    SurfWorkspace &ws = ctx->surf;
    const int dim = p->extended ? 128 : 64;
    const long long px = (long long)rows * cols;
    int kp_cap, cand_cap, max_features = 0;
    if (p->keypoints_ratio > 0) {
        long long mf = (long long)(p->keypoints_ratio * (float)px);
        if (mf > 65535) mf = 65535;
        if (mf < 1) mf = 1;
        max_features = (int)mf;
        kp_cap = max_features;
    } else {
        kp_cap = (int)(px / 24 > 4096 ? px / 24 : 4096);
    }
    cand_cap = (int)(px / 24 > 8192 ? px / 24 : 8192);
    if (cand_cap < kp_cap) cand_cap = kp_cap;
    if (ws.rows == rows && ws.cols == cols) {          // keep sizes that were regrown for this shape
        if (ws.cand_cap > cand_cap) cand_cap = ws.cand_cap;
        if (max_features == 0 && ws.kp_cap > kp_cap) kp_cap = ws.kp_cap;
    }
    cand_cap = (cand_cap + 255) & ~255;
    kp_cap = (kp_cap + 255) & ~255;                     // matcher tiles: 256 train rows per CTA pair (tcgen05), 64 (SIMT)
    const int n_bands = ceil_div(rows, INT_BAND);
    int rc;
    if ((rc = ws.integral.reserve((size_t)batch * (rows + 1) * (cols + 1) * 4))) return rc;
    if ((rc = ws.band_tot.reserve((size_t)batch * n_bands * cols * 4))) return rc;
    if ((rc = ws.cand.reserve((size_t)batch * cand_cap * KP_STRIDE * 4))) return rc;
    if ((rc = ws.sorted.reserve((size_t)batch * cand_cap * KP_STRIDE * 4))) return rc;
    if ((rc = ws.kp.reserve((size_t)batch * kp_cap * KP_STRIDE * 4))) return rc;
    if ((rc = ws.desc.reserve((size_t)batch * kp_cap * dim * 4))) return rc;
    if ((rc = ws.counters.reserve((size_t)batch * 16 + 32 + 8 * SURF_MAX_DESC_CHUNKS))) return rc;   // + work counters (two per launch group, cooperative pass)
    if ((rc = ws.prefix.reserve((size_t)(batch + 1) * 4))) return rc;
    if ((rc = ws.descT.reserve((size_t)batch * kp_cap * dim * 4))) return rc;
    if ((rc = ws.hist.reserve((size_t)batch * (3 * RB_BINS + 1) * 4))) return rc;     // sort mode 0 uses [RH_BINS + 1] per image
    ws.pitch_f = (cols + 31) & ~31;
    if (!p->upright && (rc = ws.img_f32.reserve((size_t)batch * rows * ws.pitch_f * 4))) return rc;
    if ((rc = ws.fb_list.reserve((size_t)batch * kp_cap * 8))) return rc;      // keypoints handed from the fixed-point sampler to the reference one | giant windows of the cooperative pass
    ws.max_features = max_features;
    ws.batch = batch; ws.rows = rows; ws.cols = cols; ws.cand_cap = cand_cap; ws.kp_cap = kp_cap; ws.dim = dim;
    return 0;
}

int surf_grow(vfsms_ctx *ctx, int grow_cand, int grow_kp)
{
    SurfWorkspace &ws = ctx->surf;
    if (grow_cand) ws.cand_cap *= 2;
    int rc;
    if (grow_kp && ws.max_features == 0) {
        ws.kp_cap *= 2;
        if (ws.cand_cap < ws.kp_cap) ws.cand_cap = ws.kp_cap;
        if ((rc = ws.kp.reserve((size_t)ws.batch * ws.kp_cap * KP_STRIDE * 4))) return rc;
        if ((rc = ws.desc.reserve((size_t)ws.batch * ws.kp_cap * ws.dim * 4))) return rc;
        if ((rc = ws.descT.reserve((size_t)ws.batch * ws.kp_cap * ws.dim * 4))) return rc;
        if ((rc = ws.fb_list.reserve((size_t)ws.batch * ws.kp_cap * 8))) return rc;
    }
    if ((rc = ws.cand.reserve((size_t)ws.batch * ws.cand_cap * KP_STRIDE * 4))) return rc;
    if ((rc = ws.sorted.reserve((size_t)ws.batch * ws.cand_cap * KP_STRIDE * 4))) return rc;
    return 0;
}

int surf_run_batch(vfsms_ctx *ctx, const uint8_t *base_a, const uint8_t *base_b, int split, int batch, int rows,
                   int cols, int stride, int64_t img_stride, const vfsms_surf_params *p, cudaStream_t st)
{
    SurfWorkspace &ws = ctx->surf;
    int rc;
    if ((rc = surf_init_tables())) return rc;
    if (ws.batch < batch || ws.rows != rows || ws.cols != cols || ws.dim != (p->extended ? 128 : 64)) {
        vfsms_set_error("surf_run_batch: workspace not reserved for this shape");
        return VFSMS_E_ARG;
    }
    SurfPlan plan;
    if ((rc = surf_build_plan(&plan, rows, cols, p))) return rc;
    const int n_bands = ceil_div(rows, INT_BAND);
    const int cols_pad = (min(cols, 4096) + 3) & ~3;
    const size_t smem_int = (size_t)INT_SUB * cols_pad * 4;
    static bool attr_done = false;
    if (!attr_done) {
        CUDA_TRY(cudaFuncSetAttribute(integral_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_done = true;
    }
    CUDA_TRY(cudaMemsetAsync(ws.counters.p, 0, (size_t)batch * 16 + 32 + 8 * SURF_MAX_DESC_CHUNKS, st));
    ws.last_batch = batch;
    {
    StageTimer t_int(ctx, st, VFSMS_STAGE_INTEGRAL);
    integral_band_kernel<<<dim3(n_bands, batch), 256, smem_int, st>>>(base_a, base_b, split, img_stride, rows, cols, stride,
                                                                      ws.integral.as<int32_t>(), ws.band_tot.as<int32_t>(), n_bands, ws.img_off);
    LAUNCH_CHECK(ctx);
    if (n_bands > 1) {
        integral_fix_kernel<<<dim3(n_bands - 1, batch), 256, 0, st>>>(rows, cols, ws.integral.as<int32_t>(),
                                                                      ws.band_tot.as<int32_t>(), n_bands);
        LAUNCH_CHECK(ctx);
    }
    }
    const int total_tiles = plan.tile_begin[plan.n_octaves];
    if (total_tiles > 0) {
        StageTimer t_h(ctx, st, VFSMS_STAGE_HESSIAN);
        // det planes of every layer over the haloed tile + the queue of in-layer maxima (count + one entry per tile sample)
        auto smem_det_of = [&](int o) { return (size_t)plan.n_layers * (ht_x(o) + 2) * (ht_y(o) + 2) * 4 + (size_t)(4 + ht_x(o) * ht_y(o)) * 4; };
        // octaves whose integral footprint fits in shared memory next to the det tile (two CTAs per SM)
        int o_split = 0;
        static const bool no_stage = getenv("VFSMS_HESSIAN_UNSTAGED") != nullptr;     // debugging aid
        while (!no_stage && o_split < plan.n_octaves &&
               smem_det_of(o_split) + (size_t)plan.stage_rows[o_split] * plan.stage_cols[o_split] * 4 <= 110 * 1024) o_split++;
        if (o_split > 2) o_split = 2;                  // staged kernels are instantiated for octaves 0 and 1
        // corner offsets combined with the pitch each octave's kernel reads from
        for (int o = 0; o < plan.n_octaves; o++) {
            const int P = o < o_split ? plan.stage_cols[o] : cols + 1;
            for (int l = 0; l < plan.n_layers; l++) {
                SurfLayer &L = plan.layer[o][l];
                const int xs[4] = { L.dx[0].x1, L.dx[1].x1, L.dx[2].x1, L.dx[2].x2 };
                for (int k = 0; k < 4; k++) { L.off[k] = L.dx[0].y1 * P + xs[k]; L.off[4 + k] = L.dx[0].y2 * P + xs[k]; }
                const int ys[4] = { L.dy[0].y1, L.dy[1].y1, L.dy[2].y1, L.dy[2].y2 };
                for (int k = 0; k < 4; k++) { L.off[8 + k] = ys[k] * P + L.dy[0].x1; L.off[12 + k] = ys[k] * P + L.dy[0].x2; }
                const int gx[4] = { L.dxy[0].x1, L.dxy[0].x2, L.dxy[1].x1, L.dxy[1].x2 };
                const int gy[4] = { L.dxy[0].y1, L.dxy[0].y2, L.dxy[2].y1, L.dxy[2].y2 };
                for (int r = 0; r < 4; r++) for (int k = 0; k < 4; k++) L.off[16 + 4 * r + k] = gy[r] * P + gx[k];
            }
        }
        static bool attr_h = false;
        if (!attr_h) {
            CUDA_TRY(cudaFuncSetAttribute(hessian_nms_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(hessian_nms_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(hessian_nms_kernel<true, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
            CUDA_TRY(cudaFuncSetAttribute(hessian_nms_kernel<true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
            attr_h = true;
        }
        // the compile-time tables apply when they reproduce this plan exactly (default nOctaveLayers = 3)
        bool fixed_ok[2] = { plan.n_layers == 5 && o_split > 0, plan.n_layers == 5 && o_split > 1 };
        static const bool no_fixed = getenv("VFSMS_HESSIAN_GENERIC") != nullptr;           // debugging aid
        auto check_fixed = [&](int o, int l, int size, int margin, int pitch, int (*off)(int), float (*w)(int)) {
            const SurfLayer &L = plan.layer[o][l];
            bool ok = L.size == size && L.margin == margin && plan.stage_cols[o] == pitch && plan.stage_rows[o] == HessFixed<0, 0>::rows * (o == 0) + HessFixed<1, 0>::rows * (o == 1) &&
                      (plan.stage_off_min[o] & 1) == 0;
            for (int k = 0; ok && k < 32; k++) ok = L.off[k] == off(k);
            const float pw[10] = { L.dx[0].w, L.dx[1].w, L.dx[2].w, L.dy[0].w, L.dy[1].w, L.dy[2].w, L.dxy[0].w, L.dxy[1].w, L.dxy[2].w, L.dxy[3].w };
            for (int k = 0; ok && k < 10; k++) { const float wk = w(k); ok = memcmp(&pw[k], &wk, 4) == 0; }
            if (!ok) fixed_ok[o] = false;
        };
#define CHK(O, LL) if (fixed_ok[O]) check_fixed(O, LL, HessFixed<O, LL>::size, HessFixed<O, LL>::margin, HessFixed<O, LL>::pitch, \
                                                [](int k) { return HessFixed<O, LL>::off(k); }, [](int k) { return HessFixed<O, LL>::w(k); })
        CHK(0, 0); CHK(0, 1); CHK(0, 2); CHK(0, 3); CHK(0, 4); CHK(1, 0); CHK(1, 1); CHK(1, 2); CHK(1, 3); CHK(1, 4);
#undef CHK
        if (no_fixed) fixed_ok[0] = fixed_ok[1] = false;
        for (int o = 0; o < plan.n_octaves; o++) {          // one launch per octave: its own smem footprint, octave as template parameter
            const int nt = plan.tile_begin[o + 1] - plan.tile_begin[o];
            if (nt <= 0) continue;
            const bool staged = o < o_split;
            size_t sm = smem_det_of(o) + (staged ? (size_t)plan.stage_rows[o] * plan.stage_cols[o] * 4 : 0);
            if (o == 1 && staged && fixed_ok[1]) sm = smem_det_of(1) + (size_t)2 * HessFixed<1, 0>::plane * 4;      // two column-parity planes
            const dim3 grid(nt, batch);
#define HL(S, O) hessian_nms_kernel<S, O><<<grid, HT_THREADS, sm, st>>>(plan, ws.integral.as<int32_t>(), ws.cand.as<float>(), ws.counters.as<int32_t>(), ws.cand_cap)
            switch (o) {
#define HLF(O) hessian_nms_kernel<true, O, true><<<grid, HT_THREADS, sm, st>>>(plan, ws.integral.as<int32_t>(), ws.cand.as<float>(), ws.counters.as<int32_t>(), ws.cand_cap)
            case 0: if (staged && fixed_ok[0]) HLF(0); else if (staged) HL(true, 0); else HL(false, 0); break;
            case 1: if (staged && fixed_ok[1]) HLF(1); else if (staged) HL(true, 1); else HL(false, 1); break;
#undef HLF
            case 2: HL(false, 2); break;
            case 3: HL(false, 3); break;
            default: HL(false, 4); break;
            }
#undef HL
            LAUNCH_CHECK(ctx);
        }
    }
    const int max_features = ws.max_features;
    if (ctx->sort_mode == 1 && p->hessian_threshold >= 0.f) {
        StageTimer t_s(ctx, st, VFSMS_STAGE_SORT);
        int *hist = ws.hist.as<int>();
        int *start = hist + (size_t)batch * RB_BINS, *fill = start + (size_t)batch * RB_BINS;
        unsigned *thr_bits = (unsigned *)(fill + (size_t)batch * RB_BINS);
        CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)batch * RB_BINS * 4, st));
        const int gx = min(ceil_div(ws.cand_cap, 256), 16);
        bin_hist_kernel<<<dim3(gx, batch), 256, 0, st>>>(ws.cand.as<float>(), ws.counters.as<int32_t>(), ws.cand_cap, hist);
        LAUNCH_CHECK(ctx);
        bin_offsets_kernel<<<ceil_div(batch, 8), 256, 0, st>>>(hist, ws.counters.as<int32_t>(), ws.cand_cap, max_features, batch, start, fill, thr_bits);
        LAUNCH_CHECK(ctx);
        bin_scatter_kernel<<<dim3(gx, batch), 256, 0, st>>>(ws.cand.as<float>(), ws.sorted.as<float>(), ws.counters.as<int32_t>(), ws.cand_cap,
                                                            thr_bits, start, fill);
        LAUNCH_CHECK(ctx);
        const int chunks = ceil_div(ws.cand_cap, 256) * batch;
        bin_rank_kernel<<<min(chunks, ctx->num_sms * 8), 256, 0, st>>>(ws.sorted.as<float>(), ws.cand.as<float>(), ws.counters.as<int32_t>(),
                                                                        ws.cand_cap, batch, max_features, hist, start);
        LAUNCH_CHECK(ctx);
        rank_finalize_kernel<<<ceil_div(batch, 256), 256, 0, st>>>(ws.counters.as<int32_t>(), batch, max_features);
        LAUNCH_CHECK(ctx);
    } else {
        StageTimer t_s(ctx, st, VFSMS_STAGE_SORT);
        int *hist = ws.hist.as<int>();
        unsigned *thr_bits = (unsigned *)(hist + (size_t)batch * RH_BINS);
        CUDA_TRY(cudaMemsetAsync(hist, 0, (size_t)batch * RH_BINS * 4, st));
        const int gx = min(ceil_div(ws.cand_cap, 256), 16);
        response_hist_kernel<<<dim3(gx, batch), 256, 0, st>>>(ws.cand.as<float>(), ws.counters.as<int32_t>(), ws.cand_cap, hist);
        LAUNCH_CHECK(ctx);
        response_threshold_kernel<<<ceil_div(batch, 8), 256, 0, st>>>(hist, ws.counters.as<int32_t>(), ws.cand_cap, max_features, batch, thr_bits);
        LAUNCH_CHECK(ctx);
        response_filter_kernel<<<dim3(gx, batch), 256, 0, st>>>(ws.cand.as<float>(), ws.sorted.as<float>(), ws.counters.as<int32_t>(), ws.cand_cap, thr_bits);
        LAUNCH_CHECK(ctx);
        // staged (ws.sorted) -> ordered (ws.cand, whose raw content is no longer needed)
        const int chunks = ceil_div(ws.cand_cap, 256) * batch;
        const int grid = min(chunks, ctx->num_sms * 8);
        rank_sort_kernel<<<grid, 256, 0, st>>>(ws.sorted.as<float>(), ws.cand.as<float>(), ws.counters.as<int32_t>(),
                                              ws.cand_cap, batch, max_features);
        LAUNCH_CHECK(ctx);
        rank_finalize_kernel<<<ceil_div(batch, 256), 256, 0, st>>>(ws.counters.as<int32_t>(), batch, max_features);
        LAUNCH_CHECK(ctx);
    }
    {
    StageTimer t_c(ctx, st, VFSMS_STAGE_COMPACT);
    validate_compact_kernel<<<batch, 1024, 0, st>>>(ws.cand.as<float>(), ws.kp.as<float>(), ws.counters.as<int32_t>(),
                                                    ws.cand_cap, ws.kp_cap, rows, cols, p->upright);
    LAUNCH_CHECK(ctx);
    prefix_kernel<<<1, 32, 0, st>>>(ws.counters.as<int32_t>(), ws.prefix.as<int32_t>(), batch);
    LAUNCH_CHECK(ctx);
    }
    StageTimer t_d(ctx, st, VFSMS_STAGE_DESCRIBE);
    int *work_counter = ws.counters.as<int32_t>() + (size_t)batch * 4;      // extra slots after the per-image counters:
    int *big_flag = work_counter + 1, *fb_count = work_counter + 2;         // [0] reference queue, [1] big-window flag, [2] hand-over count,
    int *fb_list = ws.fb_list.as<int>();                                    // [4 + c], [4 + SURF_MAX_DESC_CHUNKS + c] queues of group c
    const int lpt_split = ctx->describe_lpt == 1 ? 128 : (ctx->describe_lpt == 2 ? 64 : (ctx->describe_lpt == 3 ? 256 : 0));
    // describe = 1 (default): fixed-point chunked sampler over a float copy of the images scaled by 2^-64, read through one stacked
    // texture per group of images whose rows fit the 2-D linear texture height limit; one launch per group with its own queues.
    // Keypoints it hands over (and everything when describe = 0, upright, or no texture) go through the reference sampler.
    // describe = 2 / 3: tolerance modes -- same kernel structure, window pixels from fp32 lerps on a gather / from the texture
    // unit's bilinear filter (NOT bit-exact; surf_describe.cuh).
    bool fixed = false;
    const int tol = ctx->describe_mode >= 2 ? ctx->describe_mode - 1 : 0;      // 1: fp32 lerps on a gather, 2: texture-unit filter
    if (ctx->describe_mode >= 1 && !p->upright) {
        int max_h = 0;
        if (cudaDeviceGetAttribute(&max_h, cudaDevAttrMaxTexture2DLinearHeight, ctx->device) != cudaSuccess) { cudaGetLastError(); max_h = 0; }
        int per = max_h >= rows ? max_h / rows : 0;
        if (per >= 4) per &= ~3;          // keeps every group's base address on the 512-byte texture alignment (pitch is 128-byte aligned)
        const int n_chunks = per > 0 ? ceil_div(batch, per) : 0;
        if (n_chunks > 0 && n_chunks <= SURF_MAX_DESC_CHUNKS) {
            if ((rc = ws.img_f32.reserve((size_t)batch * rows * ws.pitch_f * 4))) return rc;
            u8_to_f32_kernel<<<dim3(std::min(ceil_div(rows * cols, 256), ctx->num_sms * 4), batch), 256, 0, st>>>(
                base_a, base_b, split, img_stride, rows, cols, stride, ws.img_f32.as<float>(), ws.pitch_f,
                tol ? 1.0f : 5.421010862427522e-20f /* 2^-64 */, ws.img_off);
            LAUNCH_CHECK(ctx);
            std::vector<cudaTextureObject_t> ts((size_t)n_chunks);
            fixed = true;
            for (int c = 0; c < n_chunks && fixed; c++)
                fixed = surf_stack_texture(ctx, c * per, std::min(per, batch - c * per), rows, cols, ws.pitch_f, st, &ts[c], tol == 2);
            for (int c = 0; c < n_chunks && fixed; c++) {
                // Small batches: one giant window (up to 768 px, 6 * 10^5 samples) on one warp is the whole tail of the launch
                // (3 ms of a 3.5 ms single-pair call); the CTA's eight warps share such windows (surf_describe.cuh).
                const bool coop = n_chunks == 1 && batch <= DESC_COOP_MAX_BATCH;
                const int coop_split = batch <= 4 ? DESC_COOP_SPLIT_TINY : DESC_COOP_SPLIT;       // one or two pairs: share more windows
                int *coop_list = fb_list + (size_t)batch * ws.kp_cap, *coop_count = work_counter + 3,
                    *coop_counter = work_counter + 4 + 2 * SURF_MAX_DESC_CHUNKS;
                if (coop) {
                    describe_giants_kernel<<<std::min(ceil_div(batch * ws.kp_cap, 256), ctx->num_sms * 2), 256, 0, st>>>(
                        ws.kp.as<float>(), ws.prefix.as<int32_t>(), batch, ws.kp_cap, coop_split, coop_list, coop_count);
                    LAUNCH_CHECK(ctx);
                }
#define LAUNCH_COOP(MB, UU, NW, TT) describe_fixed_kernel<MB, UU, NW, TT, true><<<ctx->num_sms * MB, NW * 32, 0, st>>>(                                  \
                    base_a, base_b, split, img_stride, rows, cols, stride, ws.integral.as<int32_t>(), ws.kp.as<float>(), ws.desc.as<float>(), \
                    ws.prefix.as<int32_t>(), batch, ws.kp_cap, p->extended, ts[c], c * per, std::min(per, batch - c * per),                   \
                    work_counter + 4 + c, work_counter + 4 + SURF_MAX_DESC_CHUNKS + c, lpt_split, big_flag, fb_list, fb_count, ws.img_off,   \
                    coop_split, coop_list, coop_count, coop_counter)
#define LAUNCH_FIXED(MB, UU, NW, TT) describe_fixed_kernel<MB, UU, NW, TT><<<ctx->num_sms * MB, NW * 32, 0, st>>>(                                            \
                    base_a, base_b, split, img_stride, rows, cols, stride, ws.integral.as<int32_t>(), ws.kp.as<float>(), ws.desc.as<float>(), \
                    ws.prefix.as<int32_t>(), batch, ws.kp_cap, p->extended, ts[c], c * per, std::min(per, batch - c * per),                   \
                    work_counter + 4 + c, work_counter + 4 + SURF_MAX_DESC_CHUNKS + c, lpt_split, big_flag, fb_list, fb_count, ws.img_off)
                // (3 CTAs per SM, 4 gathers in flight: measured against 2 / 4 CTAs and 2 / 6 / 8 gathers, profiles/r02)
                {
                    // (8 warps per CTA, 3 CTAs per SM: 4-warp CTAs at 6 / 7 per SM and 2-warp CTAs at 14 measured no better)
                    if (coop && tol == 0) LAUNCH_COOP(DESC_FIXED_MINB, DESC_FIXED_U, WK_WARPS, 0);
                    else if (tol == 1) LAUNCH_FIXED(DESC_FIXED_MINB, DESC_FIXED_U, WK_WARPS, 1);
                    else if (tol == 2) LAUNCH_FIXED(DESC_FIXED_MINB, DESC_FIXED_U, WK_WARPS, 2);      // (8 in flight, 4 or 2 CTAs per SM: no faster -- the fp32 filter rate of the texture unit is the limit)
                    else LAUNCH_FIXED(DESC_FIXED_MINB, DESC_FIXED_U, WK_WARPS, 0);
                }
#undef LAUNCH_FIXED
#undef LAUNCH_COOP
                LAUNCH_CHECK(ctx);
            }
        }
    }
    describe_reference_kernel<<<fixed ? ctx->num_sms : ctx->num_sms * 4, WK_WARPS * 32, 0, st>>>(
        base_a, base_b, split, img_stride, rows, cols, stride, ws.integral.as<int32_t>(), ws.kp.as<float>(), ws.desc.as<float>(),
        ws.prefix.as<int32_t>(), batch, ws.kp_cap, p->extended, p->upright, work_counter, big_flag,
        fixed ? fb_list : nullptr, fixed ? fb_count : nullptr, ws.img_off);
    LAUNCH_CHECK(ctx);
    orient_describe_kernel<<<ctx->num_sms * 2, DESC_THREADS, 0, st>>>(base_a, base_b, split, img_stride, rows, cols, stride,
                                                                       ws.integral.as<int32_t>(), ws.kp.as<float>(), ws.desc.as<float>(),
                                                                       ws.prefix.as<int32_t>(), batch, ws.kp_cap, p->extended, p->upright, big_flag, ws.img_off);
    LAUNCH_CHECK(ctx);
    return 0;
}
