// jpeg.cu -- grayscale decode of JPEG tiles.
//
// Replaces `cv2.imdecode(np.fromfile(f), 0)` of the reference (Stitcher.py:68-69, again at :382 / :401): libjpeg with
// out_color_space = JCS_GRAYSCALE reconstructs ONLY the luma component -- entropy decode, dequantise, accurate integer
// IDCT (jidctint.c "islow", the library default), range limit.  Split here as the hardware wants it:
//   host   : marker parsing + Huffman decoding (inherently serial per file; files are decoded in parallel on the host
//            cores) into quantised luma coefficients, int16, natural order, in pinned memory;
//   device : jpeg_idct_luma_kernel -- dequantise + the two islow passes + range limit, one 8x8 block per 8 threads,
//            writing u8 rows straight into the HBM-resident tile.  HBM-bound: 2 B/px in, 1 B/px out.
// Integer arithmetic throughout: the result is bit-identical to cv2.imdecode (tests/test_gpu_jpeg.py).
// Supported: baseline / extended sequential Huffman (SOF0 / SOF1), 8-bit, one interleaved scan or a single component,
// restart intervals -- every JPEG of the reference's demoImages.  Anything else returns VFSMS_E_UNSUPPORTED and the
// caller keeps its own decoder.
#include "common.cuh"
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <thread>

// ---------------------------------------------------------------- host: parsing + Huffman
static const uint8_t kZigzag[64] = { 0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20,
                                     13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59,
                                     52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };

struct HuffTable {
    bool present = false;
    uint16_t fast[512];       // 9-bit prefix -> (length << 8) | symbol, 0xFFFF when the code is longer
    int32_t maxcode[18];      // largest code of each length, -1 when none
    int32_t valptr[17];       // symbol index = code + valptr[length]
    uint8_t vals[256];
    // AC tables: FAST_AC_BITS-bit prefix -> (value << 8) | (run << 4) | (code length + magnitude bits) when the whole
    // (run/size code + magnitude) fits in the prefix and the value in 8 bits; 0 otherwise
    int16_t fast_ac[1 << 12];
};
#define FAST_AC_BITS 12

struct JpegComp { int id, h, v, tq, td, ta; };

struct JpegHeader {
    int rows = 0, cols = 0, ncomp = 0;
    JpegComp comp[4];
    uint16_t quant[4][64];    // natural order
    bool quant_present[4] = { false, false, false, false };
    HuffTable dc[4], ac[4];
    int restart_interval = 0;
    size_t scan_begin = 0;
    int mcux = 0, mcuy = 0, blocks_w = 0, blocks_h = 0;     // luma block grid (padded to whole MCUs)
};

static bool build_huff(HuffTable &t, const uint8_t *bits /* 16 */, const uint8_t *vals, int n)
{
    {   // a valid table never assigns more codes of a length than remain (Kraft); a damaged one would index past the tables
        int code = 0;
        for (int l = 1; l <= 16; l++) { code += bits[l - 1]; if (code > (1 << l)) return false; code <<= 1; }
    }
    t.present = true;
    memcpy(t.vals, vals, (size_t)n);
    for (int i = 0; i < 512; i++) t.fast[i] = 0xFFFF;
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
        t.valptr[l] = k - code;
        for (int i = 0; i < bits[l - 1]; i++, k++, code++)
            if (l <= 9) {
                const int lo = code << (9 - l), cnt = 1 << (9 - l);
                for (int j = 0; j < cnt; j++) t.fast[lo + j] = (uint16_t)((l << 8) | vals[k]);
            }
        t.maxcode[l] = bits[l - 1] ? code - 1 : -1;
        code <<= 1;
    }
    t.maxcode[17] = 0x7fffffff;
    // combined (symbol, magnitude) lookup for short AC codes
    for (int i = 0; i < (1 << FAST_AC_BITS); i++) {
        t.fast_ac[i] = 0;
        const uint16_t e = t.fast[i >> (FAST_AC_BITS - 9)];
        if (e == 0xFFFF) continue;
        const int len = e >> 8, rs = e & 255, run = rs >> 4, mag = rs & 15;
        if (mag == 0 || len + mag > FAST_AC_BITS) continue;
        int v = (i >> (FAST_AC_BITS - len - mag)) & ((1 << mag) - 1);
        if (v < (1 << (mag - 1))) v -= (1 << mag) - 1;
        if (v >= -128 && v <= 127) t.fast_ac[i] = (int16_t)(v * 256 + run * 16 + len + mag);
    }
    return true;
}

static int jpeg_parse(const uint8_t *d, size_t n, JpegHeader &H)
{
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) { vfsms_set_error("jpeg: no SOI marker"); return VFSMS_E_UNSUPPORTED; }
    size_t i = 2;
    bool have_sof = false;
    int adobe_transform = -1;
    while (i + 4 <= n) {
        if (d[i] != 0xFF) { vfsms_set_error("jpeg: marker expected at byte %zu", i); return VFSMS_E_UNSUPPORTED; }
        while (i + 1 < n && d[i + 1] == 0xFF) i++;
        const int m = d[i + 1];
        i += 2;
        if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
        if (m == 0xD9) break;
        if (i + 2 > n) break;
        const size_t L = ((size_t)d[i] << 8) | d[i + 1];
        if (L < 2 || i + L > n) { vfsms_set_error("jpeg: truncated segment"); return VFSMS_E_UNSUPPORTED; }
        const uint8_t *s = d + i + 2;
        const size_t sl = L - 2;
        if (m == 0xDB) {
            size_t k = 0;
            while (k < sl) {
                const int pq = s[k] >> 4, tq = s[k] & 15;
                k++;
                if (tq > 3 || k + (pq ? 128 : 64) > sl) { vfsms_set_error("jpeg: bad DQT"); return VFSMS_E_UNSUPPORTED; }
                for (int z = 0; z < 64; z++) {
                    H.quant[tq][kZigzag[z]] = pq ? (uint16_t)((s[k] << 8) | s[k + 1]) : s[k];
                    k += pq ? 2 : 1;
                }
                H.quant_present[tq] = true;
            }
        } else if (m == 0xC4) {
            size_t k = 0;
            while (k + 17 <= sl) {
                const int tc = s[k] >> 4, th = s[k] & 15;
                int cnt = 0;
                for (int b = 0; b < 16; b++) cnt += s[k + 1 + b];
                if (th > 3 || tc > 1 || cnt > 256 || k + 17 + cnt > sl) { vfsms_set_error("jpeg: bad DHT"); return VFSMS_E_UNSUPPORTED; }
                if (!build_huff(tc ? H.ac[th] : H.dc[th], s + k + 1, s + k + 17, cnt)) { vfsms_set_error("jpeg: inconsistent Huffman table"); return VFSMS_E_UNSUPPORTED; }
                k += 17 + cnt;
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (sl < 6 || s[0] != 8) { vfsms_set_error("jpeg: only 8-bit precision is supported"); return VFSMS_E_UNSUPPORTED; }
            H.rows = (s[1] << 8) | s[2]; H.cols = (s[3] << 8) | s[4]; H.ncomp = s[5];
            if (H.ncomp < 1 || H.ncomp > 4 || sl < (size_t)(6 + 3 * H.ncomp) || H.rows == 0 || H.cols == 0) {
                vfsms_set_error("jpeg: bad SOF"); return VFSMS_E_UNSUPPORTED;
            }
            for (int c = 0; c < H.ncomp; c++) {
                H.comp[c] = { s[6 + 3 * c], s[7 + 3 * c] >> 4, s[7 + 3 * c] & 15, s[8 + 3 * c], 0, 0 };
                if (H.comp[c].h < 1 || H.comp[c].h > 4 || H.comp[c].v < 1 || H.comp[c].v > 4 || H.comp[c].tq > 3) {
                    vfsms_set_error("jpeg: bad sampling factors"); return VFSMS_E_UNSUPPORTED;
                }
            }
            have_sof = true;
        } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xC7) || (m >= 0xC9 && m <= 0xCB) || (m >= 0xCD && m <= 0xCF)) {
            vfsms_set_error("jpeg: SOF%d (progressive / lossless / arithmetic) is not supported", m - 0xC0);
            return VFSMS_E_UNSUPPORTED;
        } else if (m == 0xEE) {
            if (sl >= 12 && memcmp(s, "Adobe", 5) == 0) adobe_transform = s[11];
        } else if (m == 0xDD) {
            if (sl >= 2) H.restart_interval = (s[0] << 8) | s[1];
        } else if (m == 0xDA) {
            if (!have_sof || sl < 1 || s[0] != H.ncomp || sl < (size_t)(1 + 2 * H.ncomp + 3)) {
                vfsms_set_error("jpeg: non-interleaved multi-scan files are not supported"); return VFSMS_E_UNSUPPORTED;
            }
            for (int c = 0; c < H.ncomp; c++) {
                const int cid = s[1 + 2 * c], t = s[2 + 2 * c];
                int which = -1;
                for (int z = 0; z < H.ncomp; z++) if (H.comp[z].id == cid) which = z;
                if (which != c) { vfsms_set_error("jpeg: scan component order differs from the frame"); return VFSMS_E_UNSUPPORTED; }
                H.comp[c].td = t >> 4; H.comp[c].ta = t & 15;
                if (H.comp[c].td > 3 || H.comp[c].ta > 3 || !H.dc[H.comp[c].td].present || !H.ac[H.comp[c].ta].present) {
                    vfsms_set_error("jpeg: scan refers to a missing Huffman table"); return VFSMS_E_UNSUPPORTED;
                }
            }
            if (!H.quant_present[H.comp[0].tq]) { vfsms_set_error("jpeg: missing quantisation table"); return VFSMS_E_UNSUPPORTED; }
            // component 0 must be luma: grayscale or YCbCr files (libjpeg's colour-space guess, jdapimin.c default_decompress_parms)
            const bool rgb_ids = H.ncomp == 3 && H.comp[0].id == 'R' && H.comp[1].id == 'G' && H.comp[2].id == 'B';
            if ((H.ncomp != 1 && H.ncomp != 3) || (H.ncomp == 3 && (adobe_transform == 0 || (adobe_transform < 0 && rgb_ids)))) {
                vfsms_set_error("jpeg: only grayscale and YCbCr files are supported"); return VFSMS_E_UNSUPPORTED;
            }
            H.scan_begin = i + L;
            int hmax = 1, vmax = 1;
            for (int c = 0; c < H.ncomp; c++) { if (H.comp[c].h > hmax) hmax = H.comp[c].h; if (H.comp[c].v > vmax) vmax = H.comp[c].v; }
            if (H.ncomp == 1) { H.comp[0].h = H.comp[0].v = 1; hmax = vmax = 1; }      // a single-component scan is not interleaved
            H.mcux = (H.cols + 8 * hmax - 1) / (8 * hmax); H.mcuy = (H.rows + 8 * vmax - 1) / (8 * vmax);
            H.blocks_w = H.mcux * H.comp[0].h; H.blocks_h = H.mcuy * H.comp[0].v;
            return 0;
        }
        i += L;
    }
    vfsms_set_error("jpeg: no SOS marker");
    return VFSMS_E_UNSUPPORTED;
}

struct BitReader {
    const uint8_t *d; size_t pos, end;
    uint64_t acc = 0; int nbits = 0;
    inline void refill()
    {
        if (pos + 8 <= end) {               // fast path: the next bytes hold no 0xFF (no stuffing, no marker)
            uint64_t w;
            memcpy(&w, d + pos, 8);
            w = __builtin_bswap64(w);
            const int k = (64 - nbits) >> 3;                                  // whole bytes that fit
            const uint64_t top = k == 8 ? w : (w >> (64 - 8 * k)) << (64 - 8 * k);   // only the k bytes we take
            const uint64_t x = ~top;                                          // a byte of x is 0 iff a taken byte was 0xFF
            if (k > 0 && !((x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull)) {
                acc = k == 8 ? w : (acc << (8 * k)) | (w >> (64 - 8 * k));
                nbits += 8 * k; pos += (size_t)k;
                return;
            }
        }
        while (nbits <= 56) {
            uint32_t b = 0;
            if (pos < end) {
                b = d[pos];
                if (b == 0xFF) {
                    if (pos + 1 < end && d[pos + 1] == 0) pos += 2;
                    else b = 0;                      // a marker: feed zeros, stay on it
                } else pos++;
            }
            acc = (acc << 8) | b; nbits += 8;
        }
    }
    inline uint32_t peek(int k) const { return (uint32_t)(acc >> (nbits - k)) & ((1u << k) - 1u); }
    inline void skip(int k) { nbits -= k; }
    inline int symbol(const HuffTable &t)
    {
        if (nbits < 32) refill();
        const uint16_t e = t.fast[peek(9)];
        if (e != 0xFFFF) { skip(e >> 8); return e & 255; }
        for (int l = 10; l <= 16; l++) {
            const int code = (int)peek(l);
            if (code <= t.maxcode[l]) { skip(l); return t.vals[(code + t.valptr[l]) & 255]; }
        }
        skip(16);
        return 0;                                    // corrupt stream: keep going, like libjpeg's warning path
    }
    inline int receive_extend(int s)
    {
        if (s == 0) return 0;
        if (nbits < 32) refill();
        const int v = (int)peek(s);
        skip(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
    void restart()
    {
        acc = 0; nbits = 0;
        while (pos + 1 < end && !(d[pos] == 0xFF && d[pos + 1] >= 0xD0 && d[pos + 1] <= 0xD7)) pos++;
        if (pos + 1 < end) pos += 2;
    }
};

// Entropy-decode one file.  coef[c] = destination of component c (quantised, natural order,
// [mcuy * v_c][mcux * h_c][64]) or nullptr to parse and drop that component (the grayscale path drops chroma).
static void jpeg_entropy_decode(const uint8_t *d, size_t n, const JpegHeader &H, int16_t *const coef[4])
{
    BitReader br{ d, H.scan_begin, n };
    int pred[4] = { 0, 0, 0, 0 };
    int n_mcu = 0;
    for (int my = 0; my < H.mcuy; my++)
        for (int mx = 0; mx < H.mcux; mx++, n_mcu++) {
            if (H.restart_interval && n_mcu && n_mcu % H.restart_interval == 0) { br.restart(); pred[0] = pred[1] = pred[2] = pred[3] = 0; }
            for (int c = 0; c < H.ncomp; c++) {
                const JpegComp &C = H.comp[c];
                const HuffTable &dc = H.dc[C.td], &ac = H.ac[C.ta];
                const int comp_bw = H.mcux * C.h;
                for (int v = 0; v < C.v; v++)
                    for (int h = 0; h < C.h; h++) {
                        const int s = br.symbol(dc) & 15;
                        pred[c] += br.receive_extend(s);
                        if (coef[c]) {
                            int16_t *blk = coef[c] + ((size_t)(my * C.v + v) * comp_bw + (mx * C.h + h)) * 64;
                            memset(blk, 0, 128);
                            blk[0] = (int16_t)pred[c];
                            for (int k = 1; k < 64;) {
                                if (br.nbits < 32) br.refill();
                                const int fe = ac.fast_ac[br.peek(FAST_AC_BITS)];
                                if (fe) {                                   // short code + magnitude in one lookup
                                    k += (fe >> 4) & 15;
                                    br.skip(fe & 15);
                                    if (k < 64) blk[kZigzag[k]] = (int16_t)(fe >> 8);
                                    k++;
                                    continue;
                                }
                                const int rs = br.symbol(ac), r = rs >> 4, sz = rs & 15;
                                if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
                                k += r;
                                const int val = br.receive_extend(sz);
                                if (k < 64) blk[kZigzag[k]] = (int16_t)val;
                                k++;
                            }
                        } else {
                            for (int k = 1; k < 64;) {                  // parse and drop
                                if (br.nbits < 32) br.refill();
                                const int fe = ac.fast_ac[br.peek(FAST_AC_BITS)];
                                if (fe) { k += ((fe >> 4) & 15) + 1; br.skip(fe & 15); continue; }
                                const int rs = br.symbol(ac), r = rs >> 4, sz = rs & 15;
                                if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
                                k += r + 1;
                                if (br.nbits < 32) br.refill();
                                br.skip(sz);
                            }
                        }
                    }
            }
        }
}

// ---------------------------------------------------------------- device: dequantise + islow IDCT + range limit
struct JpegQuant { uint16_t q[64]; };

#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172

// one 1-D pass of jidctint.c (CONST_BITS = 13); the caller descales.  All products fit in 32 bits like in the C code.
__device__ __forceinline__ void islow_1d(const int (&v)[8], int (&o)[8])
{
    int z2 = v[2], z3 = v[6];
    int z1 = (z2 + z3) * FIX_0_541196100;
    int tmp2 = z1 + z3 * (-FIX_1_847759065);
    int tmp3 = z1 + z2 * FIX_0_765366865;
    z2 = v[0]; z3 = v[4];
    int tmp0 = (z2 + z3) << 13;
    int tmp1 = (z2 - z3) << 13;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = v[7]; tmp1 = v[5]; tmp2 = v[3]; tmp3 = v[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * FIX_1_175875602;
    tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
    z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    o[0] = tmp10 + tmp3; o[7] = tmp10 - tmp3;
    o[1] = tmp11 + tmp2; o[6] = tmp11 - tmp2;
    o[2] = tmp12 + tmp1; o[5] = tmp12 - tmp1;
    o[3] = tmp13 + tmp0; o[4] = tmp13 - tmp0;
}

// libjpeg's `range_limit[x & RANGE_MASK]` (table centred on +128, RANGE_MASK = 1023)
__device__ __forceinline__ uint32_t jpeg_range_limit(int x)
{
    const int idx = x & 1023;
    return idx < 128 ? idx + 128 : (idx < 512 ? 255 : (idx < 896 ? 0 : idx - 896));
}

#define JPEG_BLOCKS_PER_CTA 32
#define JPEG_WS_PITCH 72          // ints per block of the pass-1 workspace (64 + padding against bank conflicts)
// 8 threads per 8x8 block: thread c runs the column pass of column c, then thread r the row pass of row r.
__global__ void __launch_bounds__(JPEG_BLOCKS_PER_CTA * 8) jpeg_idct_luma_kernel(const int16_t *__restrict__ coef, JpegQuant Q, int blocks_w,
                                                                                 int n_blocks, uint8_t *__restrict__ out, int rows, int cols,
                                                                                 int64_t stride)
{
    __shared__ __align__(16) int16_t s_c[JPEG_BLOCKS_PER_CTA * 64];
    __shared__ __align__(16) int s_ws[JPEG_BLOCKS_PER_CTA * JPEG_WS_PITCH];
    const int t = threadIdx.x, b = t >> 3, c = t & 7;
    const int base = blockIdx.x * JPEG_BLOCKS_PER_CTA;
    {   // 32 blocks x 128 B, one int4 per thread (coefficient blocks are contiguous)
        const int4 *src = (const int4 *)(coef + (size_t)base * 64);
        const int n_valid = min(JPEG_BLOCKS_PER_CTA, n_blocks - base) * 8;          // int4s
        ((int4 *)s_c)[t] = t < n_valid ? __ldg(src + t) : make_int4(0, 0, 0, 0);
    }
    __syncthreads();
    int v[8], o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = (int)s_c[b * 64 + k * 8 + c] * (int)Q.q[k * 8 + c];
    islow_1d(v, o);
#pragma unroll
    for (int k = 0; k < 8; k++) s_ws[b * JPEG_WS_PITCH + k * 8 + c] = (o[k] + (1 << 10)) >> 11;      // DESCALE(x, CONST_BITS - PASS1_BITS)
    __syncthreads();
    const int g = base + b, r = c;
    if (g >= n_blocks) return;
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = s_ws[b * JPEG_WS_PITCH + r * 8 + k];
    islow_1d(v, o);
    uint32_t px[8];
#pragma unroll
    for (int k = 0; k < 8; k++) px[k] = jpeg_range_limit((o[k] + (1 << 17)) >> 18);                   // DESCALE(x, CONST_BITS + PASS1_BITS + 3)
    const int by = g / blocks_w, bx = g - by * blocks_w;
    const int y = by * 8 + r, x0 = bx * 8;
    if (y >= rows || x0 >= cols) return;
    uint8_t *dst = out + (size_t)y * stride + x0;
    if (x0 + 8 <= cols && (((uintptr_t)dst) & 7) == 0) {
        uint2 w;
        w.x = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
        w.y = px[4] | (px[5] << 8) | (px[6] << 16) | (px[7] << 24);
        *(uint2 *)dst = w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) if (x0 + k < cols) dst[k] = (uint8_t)px[k];
    }
}

// ---------------------------------------------------------------- device: chroma upsampling + YCbCr -> BGR
// libjpeg defaults for cv2.imdecode(data, IMREAD_COLOR) (Stitcher.py:382,401): do_fancy_upsampling (jdsample.c triangle
// filters for 2h1v, 2h2v -- only when the downsampled width exceeds 2 -- and 1h2v; box replication otherwise) and the
// 16-bit fixed-point tables of jdcolor.c, evaluated here directly.  One thread per output pixel.
struct JpegPlanes {
    const uint8_t *y, *cb, *cr;      // IDCT output planes, pitches in bytes
    int pitch_y, pitch_c;
    int hf, vf;                      // chroma upsampling factors (max sampling / chroma sampling)
    int ds_rows, ds_cols;            // true downsampled chroma size (compptr->downsampled_height / _width)
    int gray;                        // single-component file: B = G = R = Y
};

__device__ __forceinline__ int chroma_sample(const uint8_t *__restrict__ p, int pitch, int x, int y, const JpegPlanes &P)
{
    const int hf = P.hf, vf = P.vf, dr = P.ds_rows, dc = P.ds_cols;
    if (hf == 1 && vf == 1) return p[(size_t)y * pitch + x];
    const bool fancy_h = hf == 2 && dc > 2 && (vf == 1 || vf == 2);
    if (vf == 2 && (fancy_h || hf == 1)) {
        // vertical triangle: 3 * nearer row + farther row (edge rows replicated)
        const int r = y >> 1, rf = (y & 1) ? min(r + 1, dr - 1) : max(r - 1, 0);
        const uint8_t *near = p + (size_t)r * pitch, *far = p + (size_t)rf * pitch;
        if (hf == 1) return (3 * near[x] + far[x] + ((y & 1) ? 2 : 1)) >> 2;                    // h1v2_fancy_upsample
        const int i = x >> 1;
        const int cs = 3 * near[i] + far[i];
        if (x & 1) { if (i == dc - 1) return (cs * 4 + 7) >> 4; return (3 * cs + 3 * near[i + 1] + far[i + 1] + 7) >> 4; }
        if (i == 0) return (cs * 4 + 8) >> 4;
        return (3 * cs + 3 * near[i - 1] + far[i - 1] + 8) >> 4;                                // h2v2_fancy_upsample
    }
    if (fancy_h && vf == 1) {                                                                   // h2v1_fancy_upsample
        const uint8_t *row = p + (size_t)y * pitch;
        const int i = x >> 1;
        if (x & 1) return i == dc - 1 ? row[i] : (3 * row[i] + row[i + 1] + 2) >> 2;
        return i == 0 ? row[0] : (3 * row[i] + row[i - 1] + 1) >> 2;
    }
    return p[(size_t)(y / vf) * pitch + x / hf];                                                // int_upsample / h2v1 / h2v2 box
}

__global__ void __launch_bounds__(256) jpeg_upsample_bgr_kernel(JpegPlanes P, uint8_t *__restrict__ out, int rows, int cols, int64_t row_stride)
{
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= cols || y >= rows) return;
    const int Y = P.y[(size_t)y * P.pitch_y + x];
    uint8_t *dst = out + (size_t)y * row_stride + 3 * x;
    if (P.gray) { dst[0] = dst[1] = dst[2] = (uint8_t)Y; return; }
    const int cb = chroma_sample(P.cb, P.pitch_c, x, y, P) - 128, cr = chroma_sample(P.cr, P.pitch_c, x, y, P) - 128;
    // FIX(x) = (int)(x * 65536 + 0.5): 1.40200 -> 91881, 1.77200 -> 116130, 0.71414 -> 46802, 0.34414 -> 22554
    const int r = Y + ((91881 * cr + 32768) >> 16);
    const int g = Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
    const int b = Y + ((116130 * cb + 32768) >> 16);
    dst[0] = (uint8_t)min(max(b, 0), 255); dst[1] = (uint8_t)min(max(g, 0), 255); dst[2] = (uint8_t)min(max(r, 0), 255);
}

// ---------------------------------------------------------------- C ABI
// worker threads for the entropy stage: CPUs this process may run on, capped by the container's CFS quota
// (cgroup v2 cpu.max / v1 cfs_quota_us) and by 32
static int host_threads()
{
    cpu_set_t set;
    int n = 1;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
    long long quota = -1, period = 100000;
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {
        char q[32] = { 0 };
        if (fscanf(f, "%31s %lld", q, &period) == 2 && strcmp(q, "max") != 0) quota = atoll(q);
        fclose(f);
    } else if (FILE *g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) {
        if (fscanf(g, "%lld", &quota) != 1) quota = -1;
        fclose(g);
        if (FILE *h = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (fscanf(h, "%lld", &period) != 1) period = 100000; fclose(h); }
    }
    if (quota > 0 && period > 0 && quota / period < n) n = (int)(quota / period);
    if (n > 32) n = 32;
    return n < 1 ? 1 : n;
}

extern "C" int vfsms_jpeg_info(const uint8_t *data, size_t size, int *rows, int *cols, int *components)
{
    if (!data) { vfsms_set_error("vfsms_jpeg_info: bad arguments"); return VFSMS_E_ARG; }
    JpegHeader H;
    int rc = jpeg_parse(data, size, H);
    if (rc) return rc;
    if (rows) *rows = H.rows;
    if (cols) *cols = H.cols;
    if (components) *components = H.ncomp;
    return 0;
}

extern "C" int vfsms_jpeg_component_coefficients(const uint8_t *data, size_t size, int component, int16_t *coef, size_t coef_capacity,
                                                 int *blocks_h, int *blocks_w, uint16_t *quant, int *h_samp, int *v_samp)
{
    if (!data || component < 0 || component > 3) { vfsms_set_error("vfsms_jpeg_component_coefficients: bad arguments"); return VFSMS_E_ARG; }
    JpegHeader H;
    int rc = jpeg_parse(data, size, H);
    if (rc) return rc;
    if (component >= H.ncomp) { vfsms_set_error("vfsms_jpeg_component_coefficients: the file has %d components", H.ncomp); return VFSMS_E_ARG; }
    const JpegComp &C = H.comp[component];
    if (!H.quant_present[C.tq]) { vfsms_set_error("jpeg: missing quantisation table"); return VFSMS_E_UNSUPPORTED; }
    const int bh = H.mcuy * C.v, bw = H.mcux * C.h;
    if (blocks_h) *blocks_h = bh;
    if (blocks_w) *blocks_w = bw;
    if (h_samp) *h_samp = C.h;
    if (v_samp) *v_samp = C.v;
    if (quant) memcpy(quant, H.quant[C.tq], 128);
    const size_t need = (size_t)bh * bw * 64;
    if (!coef) return 0;
    if (coef_capacity < need) { vfsms_set_error("vfsms_jpeg_component_coefficients: capacity %zu < %zu", coef_capacity, need); return VFSMS_E_CAPACITY; }
    int16_t *dst[4] = { nullptr, nullptr, nullptr, nullptr };
    dst[component] = coef;
    jpeg_entropy_decode(data, size, H, dst);
    return 0;
}

extern "C" int vfsms_jpeg_luma_coefficients(const uint8_t *data, size_t size, int16_t *coef, size_t coef_capacity, int *blocks_h,
                                            int *blocks_w, uint16_t *quant)
{
    return vfsms_jpeg_component_coefficients(data, size, 0, coef, coef_capacity, blocks_h, blocks_w, quant, nullptr, nullptr);
}

static int launch_idct(vfsms_ctx *ctx, const int16_t *cdev, const uint16_t *quant, int blocks_w, int blocks_h, uint8_t *out, int rows, int cols,
                       int64_t stride, cudaStream_t st)
{
    JpegQuant Q;
    memcpy(Q.q, quant, 128);
    const int n_blocks = blocks_h * blocks_w;
    jpeg_idct_luma_kernel<<<(n_blocks + JPEG_BLOCKS_PER_CTA - 1) / JPEG_BLOCKS_PER_CTA, JPEG_BLOCKS_PER_CTA * 8, 0, st>>>(cdev, Q, blocks_w, n_blocks, out, rows,
                                                                                                                   cols, stride);
    LAUNCH_CHECK(ctx);
    return 0;
}

// Decode n files of identical geometry.  channels 1: gray, out[i * image_stride + y * row_stride + x];
// channels 3: BGR interleaved, out[i * image_stride + y * row_stride + 3 * x + c].
static int jpeg_decode_batch(vfsms_ctx *ctx, int n, const uint8_t *const *data, const size_t *sizes, uint8_t *out_dev, int rows, int cols,
                             int channels, int64_t row_stride, int64_t image_stride, cudaStream_t st,
                             uint8_t *gray_out_dev = nullptr, int64_t gray_row_stride = 0, int64_t gray_image_stride = 0)
{
    if (n == 0) return 0;
    std::vector<JpegHeader> H((size_t)n);
    size_t per = 0, per_planes = 0;
    for (int i = 0; i < n; i++) {
        int rc = jpeg_parse(data[i], sizes[i], H[i]);
        if (rc) return rc;
        JpegHeader &h = H[i];
        if (h.rows != rows || h.cols != cols) {
            vfsms_set_error("jpeg: image %d is %d x %d, expected %d x %d", i, h.rows, h.cols, rows, cols); return VFSMS_E_ARG;
        }
        size_t blocks = (size_t)h.blocks_h * h.blocks_w;
        if (channels == 3 && h.ncomp == 3) {
            const JpegComp &cb = h.comp[1], &cr = h.comp[2];
            if (cb.h != cr.h || cb.v != cr.v || h.comp[0].h % cb.h || h.comp[0].v % cb.v || !h.quant_present[cb.tq] || !h.quant_present[cr.tq]) {
                vfsms_set_error("jpeg: unsupported chroma sampling layout"); return VFSMS_E_UNSUPPORTED;
            }
            blocks += 2 * (size_t)(h.mcuy * cb.v) * (h.mcux * cb.h);
        }
        if (blocks * 128 > per) per = blocks * 128;
        if (blocks * 64 > per_planes) per_planes = blocks * 64;
    }
    const int workers = host_threads() < n ? host_threads() : n;
    const int chunk = workers;                       // files entropy-decoded concurrently, then shipped together
    int rc;
    if ((rc = ctx->jpeg_pinned.reserve(per * chunk))) return rc;
    if ((rc = ctx->jpeg_coef.reserve(per * chunk))) return rc;
    if (channels == 3 && (rc = ctx->jpeg_planes.reserve(per_planes))) return rc;
    auto comp_blocks = [](const JpegHeader &h, int c) { return (size_t)(h.mcuy * h.comp[c].v) * (h.mcux * h.comp[c].h); };
    for (int c0 = 0; c0 < n; c0 += chunk) {
        const int cn = n - c0 < chunk ? n - c0 : chunk;
        CUDA_TRY(cudaStreamSynchronize(st));          // the previous chunk has left the pinned buffer
        std::atomic<int> next(0);
        auto work = [&]() {
            for (int j = next.fetch_add(1); j < cn; j = next.fetch_add(1)) {
                const JpegHeader &h = H[c0 + j];
                int16_t *base = (int16_t *)((uint8_t *)ctx->jpeg_pinned.p + per * j);
                int16_t *dst[4] = { base, nullptr, nullptr, nullptr };
                if (channels == 3 && h.ncomp == 3) { dst[1] = base + comp_blocks(h, 0) * 64; dst[2] = dst[1] + comp_blocks(h, 1) * 64; }
                jpeg_entropy_decode(data[c0 + j], sizes[c0 + j], h, dst);
            }
        };
        std::vector<std::thread> pool;
        for (int w = 1; w < (cn < workers ? cn : workers); w++) pool.emplace_back(work);
        work();
        for (auto &th : pool) th.join();
        for (int j = 0; j < cn; j++) {
            const JpegHeader &h = H[c0 + j];
            const bool colour = channels == 3 && h.ncomp == 3;
            size_t blocks = comp_blocks(h, 0) + (colour ? 2 * comp_blocks(h, 1) : 0);
            int16_t *cdev = (int16_t *)((uint8_t *)ctx->jpeg_coef.p + per * j);
            CUDA_TRY(cudaMemcpyAsync(cdev, (uint8_t *)ctx->jpeg_pinned.p + per * j, blocks * 128, cudaMemcpyHostToDevice, st));
            uint8_t *dst = out_dev + (size_t)(c0 + j) * image_stride;
            if (channels == 1) {
                if ((rc = launch_idct(ctx, cdev, h.quant[h.comp[0].tq], h.blocks_w, h.blocks_h, dst, rows, cols, row_stride, st))) return rc;
                continue;
            }
            // planes: full padded IDCT output of every component, then one pass of upsampling + colour conversion
            JpegPlanes P = {};
            uint8_t *py = ctx->jpeg_planes.as<uint8_t>();
            P.y = py; P.pitch_y = h.blocks_w * 8; P.gray = !colour; P.hf = P.vf = 1;
            if ((rc = launch_idct(ctx, cdev, h.quant[h.comp[0].tq], h.blocks_w, h.blocks_h, py, h.blocks_h * 8, h.blocks_w * 8, P.pitch_y, st))) return rc;
            if (gray_out_dev)      // the luma plane IS the grayscale decode of the file (cv2.imdecode(data, 0))
                CUDA_TRY(cudaMemcpy2DAsync(gray_out_dev + (size_t)(c0 + j) * gray_image_stride, (size_t)gray_row_stride, py, (size_t)P.pitch_y,
                                           (size_t)cols, (size_t)rows, cudaMemcpyDeviceToDevice, st));
            if (colour) {
                const JpegComp &Y = h.comp[0], &C = h.comp[1];
                const int cbw = h.mcux * C.h, cbh = h.mcuy * C.v;
                uint8_t *pcb = py + comp_blocks(h, 0) * 64, *pcr = pcb + comp_blocks(h, 1) * 64;
                const int16_t *ccb = cdev + comp_blocks(h, 0) * 64, *ccr = ccb + comp_blocks(h, 1) * 64;
                if ((rc = launch_idct(ctx, ccb, h.quant[h.comp[1].tq], cbw, cbh, pcb, cbh * 8, cbw * 8, cbw * 8, st))) return rc;
                if ((rc = launch_idct(ctx, ccr, h.quant[h.comp[2].tq], cbw, cbh, pcr, cbh * 8, cbw * 8, cbw * 8, st))) return rc;
                P.cb = pcb; P.cr = pcr; P.pitch_c = cbw * 8;
                P.hf = Y.h / C.h; P.vf = Y.v / C.v;
                P.ds_rows = (rows * C.v + Y.v - 1) / Y.v; P.ds_cols = (cols * C.h + Y.h - 1) / Y.h;
            }
            jpeg_upsample_bgr_kernel<<<dim3((cols + 63) / 64, (rows + 3) / 4), 256, 0, st>>>(P, dst, rows, cols, row_stride);
            LAUNCH_CHECK(ctx);
            // the plane scratch is shared by the files of the batch: stream order keeps the next IDCT behind this kernel
        }
    }
    return 0;
}

extern "C" int vfsms_jpeg_decode_gray_dev(vfsms_ctx *ctx, int n_images, const uint8_t *const *data, const size_t *sizes, uint8_t *out_dev,
                                          int rows, int cols, int64_t row_stride, int64_t image_stride, void *stream)
{
    if (!ctx || n_images < 0 || !data || !sizes || !out_dev || row_stride < cols) { vfsms_set_error("vfsms_jpeg_decode_gray_dev: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int rc = jpeg_decode_batch(ctx, n_images, data, sizes, out_dev, rows, cols, 1, row_stride, image_stride, st);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));              // the pinned staging buffer is reusable on return
    return 0;
}

extern "C" int vfsms_jpeg_decode_gray_host(vfsms_ctx *ctx, int n_images, const uint8_t *const *data, const size_t *sizes, uint8_t *out,
                                           int rows, int cols)
{
    if (!ctx || n_images < 0 || !data || !sizes || !out) { vfsms_set_error("vfsms_jpeg_decode_gray_host: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t img = (size_t)rows * cols;
    int rc;
    if ((rc = ctx->jpeg_out.reserve(img * (size_t)n_images))) return rc;
    if ((rc = jpeg_decode_batch(ctx, n_images, data, sizes, ctx->jpeg_out.as<uint8_t>(), rows, cols, 1, cols, (int64_t)img, ctx->stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, ctx->jpeg_out.p, img * (size_t)n_images, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" int vfsms_jpeg_decode_bgr_dev(vfsms_ctx *ctx, int n_images, const uint8_t *const *data, const size_t *sizes, uint8_t *out_dev,
                                         int rows, int cols, int64_t row_stride, int64_t image_stride, void *stream)
{
    if (!ctx || n_images < 0 || !data || !sizes || !out_dev || row_stride < 3 * (int64_t)cols) { vfsms_set_error("vfsms_jpeg_decode_bgr_dev: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int rc = jpeg_decode_batch(ctx, n_images, data, sizes, out_dev, rows, cols, 3, row_stride, image_stride, st);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int jpeg_decode_bgr_gray_dev(vfsms_ctx *ctx, int n, const uint8_t *const *data, const size_t *sizes, uint8_t *out_dev, int rows, int cols,
                             int64_t row_stride, int64_t image_stride, uint8_t *gray_out_dev, int64_t gray_row_stride, int64_t gray_image_stride,
                             cudaStream_t st)
{
    int rc = jpeg_decode_batch(ctx, n, data, sizes, out_dev, rows, cols, 3, row_stride, image_stride, st, gray_out_dev, gray_row_stride, gray_image_stride);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));              // the pinned staging buffer is reusable on return
    return 0;
}

extern "C" int vfsms_jpeg_decode_bgr_host(vfsms_ctx *ctx, int n_images, const uint8_t *const *data, const size_t *sizes, uint8_t *out,
                                          int rows, int cols)
{
    if (!ctx || n_images < 0 || !data || !sizes || !out) { vfsms_set_error("vfsms_jpeg_decode_bgr_host: bad arguments"); return VFSMS_E_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t img = (size_t)rows * cols * 3;
    int rc;
    if ((rc = ctx->jpeg_out.reserve(img * (size_t)n_images))) return rc;
    if ((rc = jpeg_decode_batch(ctx, n_images, data, sizes, ctx->jpeg_out.as<uint8_t>(), rows, cols, 3, (int64_t)cols * 3, (int64_t)img, ctx->stream))) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, ctx->jpeg_out.p, img * (size_t)n_images, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return 0;
}
