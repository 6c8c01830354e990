/*
 * oracle/surf_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the SURF detect+describe that the reference reaches through
 *   cv2.xfeatures2d.SURF_create().detectAndCompute      (ImageUtility.py:258,262)
 *   myGpuFeatures.detectAndDescribeBySurf               (ImageUtility.py:272, appendix/myGpuFeatures.cpp:67-104)
 * The arithmetic itself lives in a third-party dependency that is NOT vendored in /root/reference:
 *   opencv-contrib-python==3.3.1.11 (requirements.txt:113), modules/xfeatures2d/src/surf.cpp.
 * This file restates that published algorithm (Bay et al. SURF as implemented by OpenCV's CPU path; SURVEY.md
 * Appendix A) from its description: integral image -> box-filter Hessian layers -> 3x3x3 NMS -> quadratic
 * interpolation -> response-sorted keypoints -> dominant orientation -> 20s rotated window -> INTER_AREA
 * 21x21 patch -> 4x4x(4|8) descriptor -> L2 normalisation.
 *
 * PARITY STATUS: "parity unpinned" at keypoint/descriptor level -- the reference holds no SURF golden
 * vectors and cv2 in this image has no xfeatures2d.  What IS pinned (tests/test_oracle_pins.py):
 *   - integral image vs numpy cumsum (exact), INTER_AREA patch resize vs cv2.resize (exact),
 *     fastAtan2 restatement vs cv2.fastAtan2 / cv2.phase (exact),
 *   - offset level: the reference's own Stitcher driven by this SURF reproduces the golden offset list
 *     of Stitcher.py:87 (dendriticCrystal, 89 pairs) within +-1 px (tests/golden/dendritic_offsets.json).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this.
 *
 * Build: make -C oracle   ->  oracle/_build/libsurf_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORI_RADIUS 6
#define ORI_WIN 60
#define PATCH_SZ 20
#define ORI_SEARCH_INC 5
#define ORI_SIGMA 2.5f
#define DESC_SIGMA 3.3f
#define HAAR_SIZE0 9
#define HAAR_SIZE_INC 6
#define MAX_LAYERS 64

typedef struct { int p0, p1, p2, p3; float w; } SurfHF;

/* record layout shared with the CUDA path: 8 floats per keypoint */
enum { KP_X = 0, KP_Y, KP_SIZE, KP_ANGLE, KP_RESPONSE, KP_OCTAVE, KP_LAPLACIAN, KP_PAD, KP_STRIDE };

static inline int cv_round(double v) { return (int)lrint(v); }           /* round-half-even, like cvRound */
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v)  { int i = (int)v; return i + (i < v); }

/* ---------------------------------------------------------------- integral */
/* sum is (rows+1) x (cols+1) int32, first row / column zero (cv::integral, CV_32S). */
void so_integral(const uint8_t *img, int rows, int cols, int stride, int32_t *sum)
{
    int W = cols + 1;
    memset(sum, 0, sizeof(int32_t) * (size_t)W);
    for (int y = 0; y < rows; y++) {
        int32_t *out = sum + (size_t)(y + 1) * W;
        const int32_t *prev = sum + (size_t)y * W;
        const uint8_t *src = img + (size_t)y * stride;
        int32_t s = 0;
        out[0] = 0;
        for (int x = 0; x < cols; x++) { s += src[x]; out[x + 1] = prev[x + 1] + s; }
    }
}

/* ---------------------------------------------------------------- Haar helpers */
static void resize_haar(const int src[][5], SurfHF *dst, int n, int oldSize, int newSize, int widthStep)
{
    float ratio = (float)newSize / oldSize;
    for (int k = 0; k < n; k++) {
        int dx1 = cv_round(ratio * src[k][0]);
        int dy1 = cv_round(ratio * src[k][1]);
        int dx2 = cv_round(ratio * src[k][2]);
        int dy2 = cv_round(ratio * src[k][3]);
        dst[k].p0 = dy1 * widthStep + dx1;
        dst[k].p1 = dy2 * widthStep + dx1;
        dst[k].p2 = dy1 * widthStep + dx2;
        dst[k].p3 = dy2 * widthStep + dx2;
        dst[k].w = src[k][4] / ((float)(dx2 - dx1) * (dy2 - dy1));
    }
}

static inline float calc_haar(const int32_t *origin, const SurfHF *f, int n)
{
    double d = 0;
    for (int k = 0; k < n; k++)
        d += (origin[f[k].p0] + origin[f[k].p3] - origin[f[k].p1] - origin[f[k].p2]) * f[k].w;
    return (float)d;
}

/* ---------------------------------------------------------------- Hessian layer */
/* det/trace are (rows/step) x (cols/step), caller zero-initialises. */
void so_layer_det_trace(const int32_t *sum, int rows, int cols, int size, int step, float *det, float *trace)
{
    static const int dx_s[3][5] = { {0, 2, 3, 7, 1}, {3, 2, 6, 7, -2}, {6, 2, 9, 7, 1} };
    static const int dy_s[3][5] = { {2, 0, 7, 3, 1}, {2, 3, 7, 6, -2}, {2, 6, 7, 9, 1} };
    static const int dxy_s[4][5] = { {1, 1, 4, 4, 1}, {5, 1, 8, 4, -1}, {1, 5, 4, 8, -1}, {5, 5, 8, 8, 1} };
    SurfHF Dx[3], Dy[3], Dxy[4];
    int W = cols + 1;
    if (size > rows || size > cols) return;
    resize_haar(dx_s, Dx, 3, 9, size, W);
    resize_haar(dy_s, Dy, 3, 9, size, W);
    resize_haar(dxy_s, Dxy, 4, 9, size, W);
    int samples_i = 1 + (rows - size) / step;
    int samples_j = 1 + (cols - size) / step;
    int margin = (size / 2) / step;
    int lcols = cols / step;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < samples_i; i++) {
        const int32_t *sp = sum + (size_t)(i * step) * W;
        float *dp = det + (size_t)(i + margin) * lcols + margin;
        float *tp = trace + (size_t)(i + margin) * lcols + margin;
        for (int j = 0; j < samples_j; j++) {
            float dx = calc_haar(sp, Dx, 3);
            float dy = calc_haar(sp, Dy, 3);
            float dxy = calc_haar(sp, Dxy, 4);
            sp += step;
            float t = 0.81f * dxy;
            t = t * dxy;
            float p = dx * dy;
            dp[j] = p - t;
            tp[j] = dx + dy;
        }
    }
}

/* 3x3 solve, Cramer's rule in float (cv::Matx33f::solve(DECOMP_LU) fast path) */
static int solve3(const float a[3][3], const float b[3], float x[3])
{
    float d = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1])
            - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0])
            + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (d == 0) { x[0] = x[1] = x[2] = 0; return 0; }
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
    return 1;
}

/* returns 1 and updates (x, y, size) when the interpolated offset is inside the unit cube */
static int interpolate_keypoint(float N9[3][9], int dx, int dy, int ds, float *px, float *py, float *psize)
{
    float b[3] = { -(N9[1][5] - N9[1][3]) / 2, -(N9[1][7] - N9[1][1]) / 2, -(N9[2][4] - N9[0][4]) / 2 };
    float A[3][3];
    A[0][0] = N9[1][3] - 2 * N9[1][4] + N9[1][5];
    A[0][1] = (N9[1][8] - N9[1][6] - N9[1][2] + N9[1][0]) / 4;
    A[0][2] = (N9[2][5] - N9[2][3] - N9[0][5] + N9[0][3]) / 4;
    A[1][0] = A[0][1];
    A[1][1] = N9[1][1] - 2 * N9[1][4] + N9[1][7];
    A[1][2] = (N9[2][7] - N9[2][1] - N9[0][7] + N9[0][1]) / 4;
    A[2][0] = A[0][2];
    A[2][1] = A[1][2];
    A[2][2] = N9[0][4] - 2 * N9[1][4] + N9[2][4];
    float x[3];
    solve3(A, b, x);
    int ok = (x[0] != 0 || x[1] != 0 || x[2] != 0) && fabsf(x[0]) <= 1 && fabsf(x[1]) <= 1 && fabsf(x[2]) <= 1;
    if (ok) {
        *px += x[0] * dx;
        *py += x[1] * dy;
        *psize = (float)cv_round(*psize + x[2] * ds);
    }
    return ok;
}

typedef struct { float *v; int n, cap; } KpVec;

static void kpvec_push(KpVec *kv, const float *rec)
{
    if (kv->n == kv->cap) {
        kv->cap = kv->cap ? kv->cap * 2 : 4096;
        kv->v = (float *)realloc(kv->v, sizeof(float) * KP_STRIDE * (size_t)kv->cap);
    }
    memcpy(kv->v + (size_t)kv->n * KP_STRIDE, rec, sizeof(float) * KP_STRIDE);
    kv->n++;
}

/* KeypointGreater: response desc, size desc, octave desc, y desc, x asc */
static int kp_cmp(const void *pa, const void *pb)
{
    const float *a = (const float *)pa, *b = (const float *)pb;
    if (a[KP_RESPONSE] > b[KP_RESPONSE]) return -1;
    if (a[KP_RESPONSE] < b[KP_RESPONSE]) return 1;
    if (a[KP_SIZE] > b[KP_SIZE]) return -1;
    if (a[KP_SIZE] < b[KP_SIZE]) return 1;
    if (a[KP_OCTAVE] > b[KP_OCTAVE]) return -1;
    if (a[KP_OCTAVE] < b[KP_OCTAVE]) return 1;
    if (a[KP_Y] > b[KP_Y]) return -1;
    if (a[KP_Y] < b[KP_Y]) return 1;
    if (a[KP_X] < b[KP_X]) return -1;
    if (a[KP_X] > b[KP_X]) return 1;
    return 0;
}

static void find_maxima_layer(int rows, int cols, float *const *dets, float *const *traces, const int *sizes,
                              int octave, int layer, float thr, int step, KpVec *out)
{
    int size = sizes[layer];
    int lrows = rows / step, lcols = cols / step;
    int margin = (sizes[layer + 1] / 2) / step + 1;
    const float *D0 = dets[layer - 1], *D1 = dets[layer], *D2 = dets[layer + 1];
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = margin; i < lrows - margin; i++) {
        const float *det_ptr = D1 + (size_t)i * lcols;
        const float *trace_ptr = traces[layer] + (size_t)i * lcols;
        for (int j = margin; j < lcols - margin; j++) {
            float val0 = det_ptr[j];
            if (!(val0 > thr)) continue;
            int sum_i = step * (i - (size / 2) / step);
            int sum_j = step * (j - (size / 2) / step);
            float N9[3][9];
            const float *L[3] = { D0 + (size_t)i * lcols + j, D1 + (size_t)i * lcols + j, D2 + (size_t)i * lcols + j };
            for (int l = 0; l < 3; l++) {
                const float *d = L[l];
                N9[l][0] = d[-lcols - 1]; N9[l][1] = d[-lcols]; N9[l][2] = d[-lcols + 1];
                N9[l][3] = d[-1];         N9[l][4] = d[0];      N9[l][5] = d[1];
                N9[l][6] = d[lcols - 1];  N9[l][7] = d[lcols];  N9[l][8] = d[lcols + 1];
            }
            int is_max = 1;
            for (int l = 0; l < 3 && is_max; l++)
                for (int k = 0; k < 9; k++) {
                    if (l == 1 && k == 4) continue;
                    if (!(val0 > N9[l][k])) { is_max = 0; break; }
                }
            if (!is_max) continue;
            float center_i = sum_i + (size - 1) * 0.5f;
            float center_j = sum_j + (size - 1) * 0.5f;
            float x = center_j, y = center_i, ksz = (float)size;
            int ds = size - sizes[layer - 1];
            if (!interpolate_keypoint(N9, step, step, ds, &x, &y, &ksz)) continue;
            float rec[KP_STRIDE] = { x, y, ksz, -1.f, val0, (float)octave,
                                     (float)((trace_ptr[j] > 0) - (trace_ptr[j] < 0)), 0.f };
#pragma omp critical(so_push)
            kpvec_push(out, rec);
        }
    }
}

/* ---------------------------------------------------------------- fastAtan2 / Gaussian kernels */
static const float atan2_p1 = 0.9997878412794807f * (float)(180 / M_PI);
static const float atan2_p3 = -0.3258083974640975f * (float)(180 / M_PI);
static const float atan2_p5 = 0.1555786518463281f * (float)(180 / M_PI);
static const float atan2_p7 = -0.04432655554792128f * (float)(180 / M_PI);

float so_fast_atan2(float y, float x)
{
    float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((atan2_p7 * c2 + atan2_p5) * c2 + atan2_p3) * c2 + atan2_p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((atan2_p7 * c2 + atan2_p5) * c2 + atan2_p3) * c2 + atan2_p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* cv::getGaussianKernel(n, sigma, CV_32F): exp in double, normalised, stored float */
void so_gaussian_kernel(int n, double sigma, float *out)
{
    double sum = 0, scale2X = -0.5 / (sigma * sigma);
    double tmp[64];
    for (int i = 0; i < n; i++) {
        double x = i - (n - 1) * 0.5;
        double t = exp(scale2X * x * x);
        tmp[i] = (double)(float)t;          /* CV_32F kernel: value stored as float before summing */
        sum += tmp[i];
    }
    sum = 1. / sum;
    for (int i = 0; i < n; i++) out[i] = (float)(tmp[i] * sum);
}

/* ---------------------------------------------------------------- INTER_AREA resize of the win x win u8 window to 21x21 */
typedef struct { int si, di; float alpha; } DecimateAlpha;

static int area_tab(int ssize, int dsize, double scale, DecimateAlpha *tab)
{
    int k = 0;
    for (int dx = 0; dx < dsize; dx++) {
        double fsx1 = dx * scale;
        double fsx2 = fsx1 + scale;
        double cellWidth = fmin(scale, ssize - fsx1);
        int sx1 = cv_ceil(fsx1), sx2 = cv_floor(fsx2);
        if (sx2 > ssize - 1) sx2 = ssize - 1;
        if (sx1 > sx2) sx1 = sx2;
        if (sx1 - fsx1 > 1e-3) { tab[k].di = dx; tab[k].si = sx1 - 1; tab[k++].alpha = (float)((sx1 - fsx1) / cellWidth); }
        for (int sx = sx1; sx < sx2; sx++) { tab[k].di = dx; tab[k].si = sx; tab[k++].alpha = (float)(1.0 / cellWidth); }
        if (fsx2 - sx2 > 1e-3) {
            tab[k].di = dx; tab[k].si = sx2;
            tab[k++].alpha = (float)(fmin(fmin(fsx2 - sx2, 1.), cellWidth) / cellWidth);
        }
    }
    return k;
}

/* cv::resize(win(u8, n x n), patch(u8, d x d), INTER_AREA) for n >= d.  Exposed for the cv2.resize pin. */
void so_resize_area_u8(const uint8_t *win, int n, uint8_t *patch, int d)
{
    double inv_scale = (double)d / n;
    double scale = 1. / inv_scale;
    int iscale = (int)lrint(scale);
    if (n == d) { memcpy(patch, win, (size_t)n * n); return; }
    if (fabs(scale - iscale) < DBL_EPSILON) {
        /* integer factor: exact integer box sums */
        int area = iscale * iscale;
        float fscale = 1.f / area;
        for (int dy = 0; dy < d; dy++)
            for (int dx = 0; dx < d; dx++) {
                const uint8_t *S = win + (size_t)(dy * iscale) * n + dx * iscale;
                int sum = 0;
                for (int yy = 0; yy < iscale; yy++)
                    for (int xx = 0; xx < iscale; xx++) sum += S[yy * n + xx];
                if (iscale == 2) patch[dy * d + dx] = (uint8_t)((sum + 2) >> 2);
                else {
                    int v = cv_round(sum * fscale);
                    patch[dy * d + dx] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
                }
            }
        return;
    }
    DecimateAlpha *tab = (DecimateAlpha *)malloc(sizeof(DecimateAlpha) * (size_t)n * 2);
    int tabn = area_tab(n, d, scale, tab);
    float *buf = (float *)malloc(sizeof(float) * 2 * (size_t)d);
    float *sum = buf + d;
    for (int dx = 0; dx < d; dx++) sum[dx] = 0;
    int prev_dy = tab[0].di;
    for (int j = 0; j < tabn; j++) {
        float beta = tab[j].alpha;
        int dy = tab[j].di, sy = tab[j].si;
        const uint8_t *S = win + (size_t)sy * n;
        for (int dx = 0; dx < d; dx++) buf[dx] = 0;
        for (int k = 0; k < tabn; k++) buf[tab[k].di] += S[tab[k].si] * tab[k].alpha;
        if (dy != prev_dy) {
            uint8_t *D = patch + (size_t)prev_dy * d;
            for (int dx = 0; dx < d; dx++) {
                int v = cv_round(sum[dx]);
                D[dx] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
                sum[dx] = beta * buf[dx];
            }
            prev_dy = dy;
        } else {
            for (int dx = 0; dx < d; dx++) sum[dx] += beta * buf[dx];
        }
    }
    {
        uint8_t *D = patch + (size_t)prev_dy * d;
        for (int dx = 0; dx < d; dx++) {
            int v = cv_round(sum[dx]);
            D[dx] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(tab); free(buf);
}

/* ---------------------------------------------------------------- orientation + descriptor for one keypoint */
typedef struct {
    int n_ori; int apt_x[169], apt_y[169]; float aptw[169];
    float DW[PATCH_SZ * PATCH_SZ];
} SurfTables;

static void build_tables(SurfTables *t)
{
    float G[13], Gd[PATCH_SZ];
    so_gaussian_kernel(2 * ORI_RADIUS + 1, ORI_SIGMA, G);
    t->n_ori = 0;
    for (int i = -ORI_RADIUS; i <= ORI_RADIUS; i++)
        for (int j = -ORI_RADIUS; j <= ORI_RADIUS; j++)
            if (i * i + j * j <= ORI_RADIUS * ORI_RADIUS) {
                t->apt_x[t->n_ori] = i; t->apt_y[t->n_ori] = j;
                t->aptw[t->n_ori++] = G[i + ORI_RADIUS] * G[j + ORI_RADIUS];
            }
    so_gaussian_kernel(PATCH_SZ, DESC_SIGMA, Gd);
    for (int i = 0; i < PATCH_SZ; i++)
        for (int j = 0; j < PATCH_SZ; j++) t->DW[i * PATCH_SZ + j] = Gd[i] * Gd[j];
}

/* sample the (rotated) win x win window into WIN; exposed so the CUDA path's window can be compared stage-wise */
static void sample_window(const uint8_t *img, int rows, int cols, int stride, float cx, float cy,
                          int win_size, int upright, float dir_deg, uint8_t *WIN)
{
    if (!upright) {
        float descriptor_dir = dir_deg * (float)(M_PI / 180);
        /* OpenCV: -std::sin(float) / std::cos(float).  Which float comes out of sinf for the ~1.3 % of arguments whose
         * result lies close to a rounding boundary depends on the C library (glibc's sinf is faithful, not correctly rounded;
         * 3.3.1's Windows wheels used MSVC's).  The oracle takes the correctly rounded value -- double sin, then one rounding --
         * which is also what the CUDA path computes, so descriptors can be compared bit for bit. */
        float sin_dir = -(float)sin((double)descriptor_dir);
        float cos_dir = (float)cos((double)descriptor_dir);
        float win_offset = -(float)(win_size - 1) / 2;
        float start_x = cx + win_offset * cos_dir + win_offset * sin_dir;
        float start_y = cy - win_offset * sin_dir + win_offset * cos_dir;
        int ncols1 = cols - 1, nrows1 = rows - 1;
        for (int i = 0; i < win_size; i++, start_x += sin_dir, start_y += cos_dir) {
            double pixel_x = start_x, pixel_y = start_y;
            for (int j = 0; j < win_size; j++, pixel_x += cos_dir, pixel_y -= sin_dir) {
                int ix = cv_floor(pixel_x), iy = cv_floor(pixel_y);
                if ((unsigned)ix < (unsigned)ncols1 && (unsigned)iy < (unsigned)nrows1) {
                    float a = (float)(pixel_x - ix), b = (float)(pixel_y - iy);
                    const uint8_t *p = img + (size_t)iy * stride + ix;
                    WIN[i * win_size + j] = (uint8_t)cv_round(p[0] * (1.f - a) * (1.f - b) + p[1] * a * (1.f - b) +
                                                              p[stride] * (1.f - a) * b + p[stride + 1] * a * b);
                } else {
                    int x = cv_round(pixel_x), y = cv_round(pixel_y);
                    x = x < 0 ? 0 : x > ncols1 ? ncols1 : x;
                    y = y < 0 ? 0 : y > nrows1 ? nrows1 : y;
                    WIN[i * win_size + j] = img[(size_t)y * stride + x];
                }
            }
        }
    } else {
        float win_offset = -(float)(win_size - 1) / 2;
        int start_x = cv_round(cx + win_offset);
        int start_y = cv_round(cy - win_offset);
        for (int i = 0; i < win_size; i++, start_x++) {
            int pixel_x = start_x, pixel_y = start_y;
            for (int j = 0; j < win_size; j++, pixel_y--) {
                int x = pixel_x < 0 ? 0 : pixel_x, y = pixel_y < 0 ? 0 : pixel_y;
                if (x > cols - 1) x = cols - 1;
                if (y > rows - 1) y = rows - 1;
                WIN[i * win_size + j] = img[(size_t)y * stride + x];
            }
        }
    }
}

/* returns 0 when the keypoint must be deleted */
static int describe_one(const uint8_t *img, int rows, int cols, int stride, const int32_t *sum,
                        const SurfTables *T, float *kp, float *vec, int extended, int upright, uint8_t **winbuf, size_t *wincap)
{
    static const int dx_s[2][5] = { {0, 0, 2, 4, -1}, {2, 0, 4, 4, 1} };
    static const int dy_s[2][5] = { {0, 0, 4, 2, 1}, {0, 2, 4, 4, -1} };
    int W = cols + 1, srows = rows + 1, scols = cols + 1;
    float size = kp[KP_SIZE];
    float cx = kp[KP_X], cy = kp[KP_Y];
    float s = size * 1.2f / 9.0f;
    int grad_wav_size = 2 * cv_round(2 * s);
    if (srows < grad_wav_size || scols < grad_wav_size) return 0;
    float descriptor_dir = 360.f - 90.f;
    if (!upright) {
        SurfHF dx_t[2], dy_t[2];
        float X[169], Y[169], angle[169];
        int nangle = 0;
        resize_haar(dx_s, dx_t, 2, 4, grad_wav_size, W);
        resize_haar(dy_s, dy_t, 2, 4, grad_wav_size, W);
        for (int kk = 0; kk < T->n_ori; kk++) {
            int x = cv_round(cx + T->apt_x[kk] * s - (float)(grad_wav_size - 1) / 2);
            int y = cv_round(cy + T->apt_y[kk] * s - (float)(grad_wav_size - 1) / 2);
            if (y < 0 || y >= srows - grad_wav_size || x < 0 || x >= scols - grad_wav_size) continue;
            const int32_t *ptr = sum + (size_t)y * W + x;
            float vx = calc_haar(ptr, dx_t, 2);
            float vy = calc_haar(ptr, dy_t, 2);
            X[nangle] = vx * T->aptw[kk];
            Y[nangle] = vy * T->aptw[kk];
            nangle++;
        }
        if (nangle == 0) return 0;
        for (int j = 0; j < nangle; j++) angle[j] = so_fast_atan2(Y[j], X[j]);   /* cv::phase(..., degrees) */
        float bestx = 0, besty = 0, descriptor_mod = 0;
        for (int i = 0; i < 360; i += ORI_SEARCH_INC) {
            float sumx = 0, sumy = 0, temp_mod;
            for (int j = 0; j < nangle; j++) {
                int d = abs(cv_round(angle[j]) - i);
                if (d < ORI_WIN / 2 || d > 360 - ORI_WIN / 2) { sumx += X[j]; sumy += Y[j]; }
            }
            temp_mod = sumx * sumx + sumy * sumy;
            if (temp_mod > descriptor_mod) { descriptor_mod = temp_mod; bestx = sumx; besty = sumy; }
        }
        descriptor_dir = so_fast_atan2(-besty, bestx);
    }
    kp[KP_ANGLE] = descriptor_dir;
    if (!vec) return 1;

    int win_size = (int)((PATCH_SZ + 1) * s);
    size_t need = (size_t)win_size * win_size;
    if (need > *wincap) { *winbuf = (uint8_t *)realloc(*winbuf, need); *wincap = need; }
    uint8_t *WIN = *winbuf;
    sample_window(img, rows, cols, stride, cx, cy, win_size, upright, descriptor_dir, WIN);

    uint8_t PATCH[PATCH_SZ + 1][PATCH_SZ + 1];
    so_resize_area_u8(WIN, win_size, &PATCH[0][0], PATCH_SZ + 1);

    float DX[PATCH_SZ][PATCH_SZ], DY[PATCH_SZ][PATCH_SZ];
    for (int i = 0; i < PATCH_SZ; i++)
        for (int j = 0; j < PATCH_SZ; j++) {
            float dw = T->DW[i * PATCH_SZ + j];
            float vx = (PATCH[i][j + 1] - PATCH[i][j] + PATCH[i + 1][j + 1] - PATCH[i + 1][j]) * dw;
            float vy = (PATCH[i + 1][j] - PATCH[i][j] + PATCH[i + 1][j + 1] - PATCH[i][j + 1]) * dw;
            DX[i][j] = vx; DY[i][j] = vy;
        }
    int dsize = extended ? 128 : 64;
    for (int kk = 0; kk < dsize; kk++) vec[kk] = 0;
    double square_mag = 0;
    float *v = vec;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            for (int y = i * 5; y < i * 5 + 5; y++)
                for (int x = j * 5; x < j * 5 + 5; x++) {
                    float tx = DX[y][x], ty = DY[y][x];
                    if (extended) {
                        if (ty >= 0) { v[0] += tx; v[1] += fabsf(tx); } else { v[2] += tx; v[3] += fabsf(tx); }
                        if (tx >= 0) { v[4] += ty; v[5] += fabsf(ty); } else { v[6] += ty; v[7] += fabsf(ty); }
                    } else {
                        v[0] += tx; v[1] += ty; v[2] += fabsf(tx); v[3] += fabsf(ty);
                    }
                }
            int nb = extended ? 8 : 4;
            for (int kk = 0; kk < nb; kk++) square_mag += v[kk] * v[kk];
            v += nb;
        }
    float scale = (float)(1. / (sqrt(square_mag) + FLT_EPSILON));
    for (int kk = 0; kk < dsize; kk++) vec[kk] *= scale;
    return 1;
}

/* stage hook for tests: window + patch of one keypoint (x, y, size, angle given) */
int so_window_patch(const uint8_t *img, int rows, int cols, int stride, float cx, float cy, float size,
                    float dir_deg, int upright, uint8_t *win_out, int win_cap, uint8_t *patch_out)
{
    float s = size * 1.2f / 9.0f;
    int win_size = (int)((PATCH_SZ + 1) * s);
    if (win_size * win_size > win_cap) return -win_size;
    sample_window(img, rows, cols, stride, cx, cy, win_size, upright, dir_deg, win_out);
    so_resize_area_u8(win_out, win_size, patch_out, PATCH_SZ + 1);
    return win_size;
}

/* ---------------------------------------------------------------- full detect + describe */
/*
 * img: rows x cols u8 with row stride `stride` (bytes).  kp_out: cap x 8 floats, desc_out: cap x (64|128) floats
 * (may be NULL).  max_features <= 0: unlimited (cv2 CPU semantics); > 0: keep the max_features strongest after
 * the response sort (deterministic stand-in for SURF_CUDA's keypointsRatio cap, myGpuFeatures.cpp:77).
 * Returns the number of keypoints written (<= cap), or the negated total if cap was too small.
 */
int so_surf_detect_and_compute(const uint8_t *img, int rows, int cols, int stride,
                               float hessianThreshold, int nOctaves, int nOctaveLayers, int extended, int upright,
                               int max_features, float *kp_out, float *desc_out, int cap)
{
    int W = cols + 1;
    int32_t *sum = (int32_t *)malloc(sizeof(int32_t) * (size_t)(rows + 1) * W);
    so_integral(img, rows, cols, stride, sum);

    int nTotal = (nOctaveLayers + 2) * nOctaves;
    if (nTotal > MAX_LAYERS) { free(sum); return 0; }
    float *dets[MAX_LAYERS], *traces[MAX_LAYERS];
    int sizes[MAX_LAYERS], steps[MAX_LAYERS];
    int idx = 0, step = 1;
    for (int o = 0; o < nOctaves; o++) {
        for (int l = 0; l < nOctaveLayers + 2; l++) {
            size_t n = (size_t)(rows / step) * (cols / step);
            dets[idx] = (float *)calloc(n ? n : 1, sizeof(float));
            traces[idx] = (float *)calloc(n ? n : 1, sizeof(float));
            sizes[idx] = (HAAR_SIZE0 + HAAR_SIZE_INC * l) << o;
            steps[idx] = step;
            idx++;
        }
        step *= 2;
    }
    for (int i = 0; i < nTotal; i++) so_layer_det_trace(sum, rows, cols, sizes[i], steps[i], dets[i], traces[i]);

    KpVec kv = { 0, 0, 0 };
    for (int o = 0; o < nOctaves; o++)
        for (int l = 1; l <= nOctaveLayers; l++) {
            int li = o * (nOctaveLayers + 2) + l;
            find_maxima_layer(rows, cols, dets, traces, sizes, o, li, hessianThreshold, steps[li], &kv);
        }
    for (int i = 0; i < nTotal; i++) { free(dets[i]); free(traces[i]); }

    qsort(kv.v, (size_t)kv.n, sizeof(float) * KP_STRIDE, kp_cmp);
    int N = kv.n;
    if (max_features > 0 && N > max_features) N = max_features;

    SurfTables T;
    build_tables(&T);
    int dsize = extended ? 128 : 64;
    float *desc = desc_out ? (float *)malloc(sizeof(float) * (size_t)dsize * (N ? N : 1)) : NULL;
    uint8_t *keep = (uint8_t *)malloc((size_t)(N ? N : 1));
#pragma omp parallel
    {
        uint8_t *winbuf = NULL; size_t wincap = 0;
#pragma omp for schedule(dynamic, 16)
        for (int k = 0; k < N; k++)
            keep[k] = (uint8_t)describe_one(img, rows, cols, stride, sum, &T, kv.v + (size_t)k * KP_STRIDE,
                                            desc ? desc + (size_t)k * dsize : NULL, extended, upright, &winbuf, &wincap);
        free(winbuf);
    }
    int M = 0;
    for (int k = 0; k < N; k++) if (keep[k]) M++;
    int ret;
    if (M > cap) ret = -M;
    else {
        int j = 0;
        for (int k = 0; k < N; k++) {
            if (!keep[k]) continue;
            memcpy(kp_out + (size_t)j * KP_STRIDE, kv.v + (size_t)k * KP_STRIDE, sizeof(float) * KP_STRIDE);
            if (desc) memcpy(desc_out + (size_t)j * dsize, desc + (size_t)k * dsize, sizeof(float) * dsize);
            j++;
        }
        ret = M;
    }
    free(keep); free(desc); free(kv.v); free(sum);
    return ret;
}

/* ---------------------------------------------------------------- brute-force matcher restatements */
/*
 * cv2 BFMatcher(NORM_L2).knnMatch(A, B, 2) + ratio test as the reference applies it
 * (ImageUtility.py:288-296; myGpuFeatures.cpp:160-173): dist = sqrt(sum((a-b)^2)) in fp32,
 * ties -> lower train index, keep when d0 < ratio * d1.  out: M x 2 int32 (trainIdx, queryIdx), query ascending.
 */
int so_match_l2_ratio(const float *A, int nA, const float *B, int nB, int D, float ratio, int32_t *out,
                      float *dist_out /* nA x 2 or NULL */, int32_t *idx_out /* nA x 2 or NULL */)
{
    int32_t *best = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(nA ? nA : 1));
    float *bd = (float *)malloc(sizeof(float) * 2 * (size_t)(nA ? nA : 1));
#pragma omp parallel for schedule(static)
    for (int q = 0; q < nA; q++) {
        const float *a = A + (size_t)q * D;
        float d0 = FLT_MAX, d1 = FLT_MAX; int i0 = -1, i1 = -1;
        for (int t = 0; t < nB; t++) {
            const float *b = B + (size_t)t * D;
            float s = 0;
            for (int k = 0; k < D; k++) { float df = a[k] - b[k]; s += df * df; }
            if (s < d0) { d1 = d0; i1 = i0; d0 = s; i0 = t; }
            else if (s < d1) { d1 = s; i1 = t; }
        }
        bd[2 * q] = sqrtf(d0); bd[2 * q + 1] = sqrtf(d1);
        best[2 * q] = i0; best[2 * q + 1] = i1;
    }
    int M = 0;
    for (int q = 0; q < nA; q++) {
        if (dist_out) { dist_out[2 * q] = bd[2 * q]; dist_out[2 * q + 1] = bd[2 * q + 1]; }
        if (idx_out) { idx_out[2 * q] = best[2 * q]; idx_out[2 * q + 1] = best[2 * q + 1]; }
        if (best[2 * q + 1] >= 0 && bd[2 * q] < bd[2 * q + 1] * ratio) {
            out[2 * M] = best[2 * q]; out[2 * M + 1] = q; M++;
        }
    }
    free(best); free(bd);
    return M;
}

/*
 * Offset vote, ImageUtility.py:139-178 (getOffsetByMode): per match (int(yA-yB), int(xA-xB)) with truncation
 * toward zero, exact (0,0) dropped, mode by count, ties -> first seen in match order; status = count >= evaluate.
 * kps are (x, y) float32 pairs with element stride `kstride` floats.  Returns status; out = {dRow, dCol, votes}.
 */
int so_offset_by_mode(const float *kpsA, const float *kpsB, int kstride, const int32_t *matches, int M,
                      int evaluate, int32_t out[3])
{
    out[0] = out[1] = out[2] = 0;
    if (M == 0) return 0;
    int32_t *dx = (int32_t *)malloc(sizeof(int32_t) * (size_t)M), *dy = (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
    int n = 0;
    for (int m = 0; m < M; m++) {
        int t = matches[2 * m], q = matches[2 * m + 1];
        float ay = kpsA[(size_t)q * kstride + 1], ax = kpsA[(size_t)q * kstride];
        float by = kpsB[(size_t)t * kstride + 1], bx = kpsB[(size_t)t * kstride];
        int r = (int)(ay - by), c = (int)(ax - bx);      /* float32 subtraction then truncation, as numpy scalars do */
        if (r == 0 && c == 0) continue;
        dx[n] = r; dy[n] = c; n++;
    }
    if (n == 0) { dx[0] = 0; dy[0] = 0; n = 1; }
    int bestc = 0, bi = 0;
    for (int i = 0; i < n; i++) {
        int seen = 0;
        for (int j = 0; j < i; j++) if (dx[j] == dx[i] && dy[j] == dy[i]) { seen = 1; break; }
        if (seen) continue;
        int c = 0;
        for (int j = i; j < n; j++) if (dx[j] == dx[i] && dy[j] == dy[i]) c++;
        if (c > bestc) { bestc = c; bi = i; }
    }
    out[0] = dx[bi]; out[1] = dy[bi]; out[2] = bestc;
    free(dx); free(dy);
    return bestc >= evaluate;
}

int so_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
