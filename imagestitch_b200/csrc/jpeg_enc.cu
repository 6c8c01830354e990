// jpeg_enc.cu -- baseline JPEG encoder on the device (SURVEY.md 8(f) rank 2 "output encode").
//
// Replaces `cv2.imwrite(path, stitchResult)` for .jpg outputs (Stitcher.py:130-131, :196-197: every mosaic the reference produces
// leaves through it) with a byte-identical file: cv2's defaults = libjpeg(-turbo) quality 95, Annex-K Huffman tables, 4:2:0 for
// BGR input, no restart markers, JFIF 1.01 header.  The algorithm is libjpeg's (not vendored in the reference; restated and
// pinned against cv2.imencode in oracle/jpeg_encode_oracle.py, which cites the libjpeg files per step); the schedule is not:
//
//   fdct_quant   8 threads per 8x8 block: BGR -> YCbCr (jccolor.c fixed point) / 2x2 chroma box with alternating bias
//                (jcsample.c) / edge replication evaluated while LOADING, level shift, jfdctint.c islow row pass, column pass,
//                quantisation, zigzag; coefficients int16 in scan (MCU) order.  HBM: 1 B (3 B) read per pixel, 2 B (3 B) written.
//   block_bits   one thread per block: length in bits of its Huffman code (DC difference against the previous block of the
//                component; the dummy-block rules of jccoefct.c for luma blocks beyond the image)
//   scan         exclusive prefix sum of the lengths -> bit offset of every block (three plain passes, no spin-waits)
//   emit         one thread per block: codes OR-ed into the zeroed bit stream at the block's offset
//   stuff        0xFF -> 0xFF 0x00: count per 16 bytes, scan, scatter; final byte padded with one-bits (jchuff.c flush_bits)
// Only the compressed stream crosses PCIe.  The host writes the markers (jcmarker.c order) around it.
#include "common.cuh"
#include "scan.cuh"
#include <string.h>

#define ENC_BLOCKS_PER_CTA 32
#define ENC_WS_PITCH 72

struct EncGeom {
    int rows, cols, channels;
    int64_t stride;            // bytes per image row
    int bh, bw;                // gray: blocks; colour: MCUs (16x16 pixels)
    int ybh, ybw;              // real luma block rows / columns (blocks beyond them are dummies)
    int crows;                 // true downsampled chroma rows: ceil(rows / 2)
    int n_blocks;              // gray: bh * bw; colour: 6 * bh * bw
    uint16_t q[2][64];         // quantisation tables, natural order
};

__constant__ uint8_t c_enc_zigpos[64];          // natural index -> position in zigzag order
__constant__ uint32_t c_enc_dc[2][12];          // (length << 16) | code, by category
__constant__ uint32_t c_enc_ac[2][256];         // (length << 16) | code, by (run << 4) | category

static const uint8_t ZIGZAG_NATURAL[64] = {
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 };
// ITU T.81 Annex K.1 (quantisation) and K.3 - K.6 (Huffman): the tables libjpeg's jpeg_set_defaults installs
static const uint8_t STD_LUMA_Q[64] = {
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99 };
static const uint8_t STD_CHROMA_Q[64] = {
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99 };
static const uint8_t DC_BITS[2][16] = { { 0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0 }, { 0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0 } };
static const uint8_t DC_VALS[12] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 };
static const uint8_t AC_BITS[2][16] = { { 0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d }, { 0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77 } };
static const uint8_t AC_VALS[2][162] = {
    { 0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
      0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
      0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
      0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
      0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
      0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
      0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa },
    { 0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
      0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
      0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
      0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
      0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
      0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
      0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa } };

struct EncState {
    bool tables = false;
    DevBuf img, coef, bits, offs, partial, stream, counts, offs2, out, totals;
};

static EncState *enc_state(vfsms_ctx *ctx)
{
    if (!ctx->jpeg_enc_state) ctx->jpeg_enc_state = new EncState();
    return (EncState *)ctx->jpeg_enc_state;
}

void jpeg_enc_state_destroy(vfsms_ctx *ctx)
{
    EncState *s = (EncState *)ctx->jpeg_enc_state;
    if (!s) return;
    DevBuf *bufs[] = { &s->img, &s->coef, &s->bits, &s->offs, &s->partial, &s->stream, &s->counts, &s->offs2, &s->out, &s->totals };
    for (DevBuf *b : bufs) b->release();
    delete s;
    ctx->jpeg_enc_state = nullptr;
}

// jchuff.c jpeg_make_c_derived_tbl: canonical codes in order of increasing length
static void derive_codes(const uint8_t *bits, const uint8_t *vals, uint32_t *table /* indexed by symbol */)
{
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; len++) {
        for (int i = 0; i < bits[len - 1]; i++) table[vals[k++]] = ((uint32_t)len << 16) | code++;
        code <<= 1;
    }
}

static int enc_init_tables(EncState *s)
{
    if (s->tables) return 0;
    uint8_t zigpos[64];
    for (int k = 0; k < 64; k++) zigpos[ZIGZAG_NATURAL[k]] = (uint8_t)k;
    uint32_t dc[2][12], ac[2][256];
    memset(dc, 0, sizeof(dc)); memset(ac, 0, sizeof(ac));
    for (int t = 0; t < 2; t++) { derive_codes(DC_BITS[t], DC_VALS, dc[t]); derive_codes(AC_BITS[t], AC_VALS[t], ac[t]); }
    CUDA_TRY(cudaMemcpyToSymbol(c_enc_zigpos, zigpos, sizeof(zigpos)));
    CUDA_TRY(cudaMemcpyToSymbol(c_enc_dc, dc, sizeof(dc)));
    CUDA_TRY(cudaMemcpyToSymbol(c_enc_ac, ac, sizeof(ac)));
    s->tables = true;
    return 0;
}

// ---------------------------------------------------------------- samples -> quantised coefficients
// jccolor.c rgb_ycc_convert: FIX(x) = (int)(x * 65536 + 0.5), ONE_HALF = 1 << 15, CBCR_OFFSET = 128 << 16
__device__ __forceinline__ int enc_y(int b, int g, int r) { return (19595 * r + 38470 * g + 7471 * b + 32768) >> 16; }
__device__ __forceinline__ int enc_cb(int b, int g, int r) { return (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16; }
__device__ __forceinline__ int enc_cr(int b, int g, int r) { return (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16; }

// jfdctint.c jpeg_fdct_islow, one dimension.  FIRST: row pass (outputs scaled up by 2^PASS1_BITS), else column pass.
template <bool FIRST>
__device__ __forceinline__ void fdct_islow_1d(const int *d, int *o)
{
    const int tmp0 = d[0] + d[7], tmp7 = d[0] - d[7], tmp1 = d[1] + d[6], tmp6 = d[1] - d[6];
    const int tmp2 = d[2] + d[5], tmp5 = d[2] - d[5], tmp3 = d[3] + d[4], tmp4 = d[3] - d[4];
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    constexpr int SH = FIRST ? 13 - 2 : 13 + 2, RND = 1 << (SH - 1);
    if (FIRST) { o[0] = (tmp10 + tmp11) << 2; o[4] = (tmp10 - tmp11) << 2; }
    else { o[0] = (tmp10 + tmp11 + 2) >> 2; o[4] = (tmp10 - tmp11 + 2) >> 2; }
    int z1 = (tmp12 + tmp13) * 4433;                               // FIX_0_541196100
    o[2] = (z1 + tmp13 * 6270 + RND) >> SH;                         // FIX_0_765366865
    o[6] = (z1 + tmp12 * (-15137) + RND) >> SH;                     // FIX_1_847759065
    z1 = tmp4 + tmp7;
    int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
    const int z5 = (z3 + z4) * 9633;                                // FIX_1_175875602
    const int t4 = tmp4 * 2446, t5 = tmp5 * 16819, t6 = tmp6 * 25172, t7 = tmp7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    o[7] = (t4 + z1 + z3 + RND) >> SH;
    o[5] = (t5 + z2 + z4 + RND) >> SH;
    o[3] = (t6 + z2 + z3 + RND) >> SH;
    o[1] = (t7 + z1 + z4 + RND) >> SH;
}

// block g of the scan -> component (0 Y, 1 Cb, 2 Cr), block row / column inside the component, dummy flag
__device__ __forceinline__ void enc_block_pos(const EncGeom &G, int g, int &comp, int &by, int &bx, bool &dummy)
{
    if (G.channels == 1) { comp = 0; by = g / G.bw; bx = g - by * G.bw; dummy = false; return; }
    const int m = g / 6, j = g - m * 6;
    const int my = m / G.bw, mx = m - my * G.bw;
    if (j < 4) { comp = 0; by = my * 2 + (j >> 1); bx = mx * 2 + (j & 1); dummy = by >= G.ybh || bx >= G.ybw; }
    else { comp = j - 3; by = my; bx = mx; dummy = false; }
}

__global__ void __launch_bounds__(ENC_BLOCKS_PER_CTA * 8) jpeg_fdct_quant_kernel(const uint8_t *__restrict__ img, const EncGeom G, int16_t *__restrict__ coef)
{
    __shared__ int s_ws[ENC_BLOCKS_PER_CTA * ENC_WS_PITCH];
    __shared__ __align__(16) int16_t s_out[ENC_BLOCKS_PER_CTA * 64];
    const int t = threadIdx.x, b = t >> 3, r = t & 7;
    const int base = blockIdx.x * ENC_BLOCKS_PER_CTA;
    const int g = base + b;
    int comp = 0, by = 0, bx = 0;
    bool dummy = true;
    if (g < G.n_blocks) enc_block_pos(G, g, comp, by, bx, dummy);
    int d[8], o[8];
    if (!dummy) {
        if (comp == 0) {
            const int y = min(by * 8 + r, G.rows - 1);
            const uint8_t *row = img + (size_t)y * G.stride;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int x = min(bx * 8 + k, G.cols - 1);
                d[k] = (G.channels == 1 ? (int)row[x] : enc_y(row[3 * x], row[3 * x + 1], row[3 * x + 2])) - 128;
            }
        } else {
            // h2v2_downsample of the converted plane; rows / columns beyond the image replicate the last one, downsampled rows
            // beyond the true chroma height replicate the last downsampled row (jcprepct.c)
            const int cy = min(by * 8 + r, G.crows - 1);
            const uint8_t *r0 = img + (size_t)min(2 * cy, G.rows - 1) * G.stride, *r1 = img + (size_t)min(2 * cy + 1, G.rows - 1) * G.stride;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int cx = bx * 8 + k;
                const int x0 = 3 * min(2 * cx, G.cols - 1), x1 = 3 * min(2 * cx + 1, G.cols - 1);
                int s;
                if (comp == 1) s = enc_cb(r0[x0], r0[x0 + 1], r0[x0 + 2]) + enc_cb(r0[x1], r0[x1 + 1], r0[x1 + 2]) +
                                   enc_cb(r1[x0], r1[x0 + 1], r1[x0 + 2]) + enc_cb(r1[x1], r1[x1 + 1], r1[x1 + 2]);
                else s = enc_cr(r0[x0], r0[x0 + 1], r0[x0 + 2]) + enc_cr(r0[x1], r0[x1 + 1], r0[x1 + 2]) +
                         enc_cr(r1[x0], r1[x0 + 1], r1[x0 + 2]) + enc_cr(r1[x1], r1[x1 + 1], r1[x1 + 2]);
                d[k] = ((s + 1 + (cx & 1)) >> 2) - 128;              // bias 1, 2, 1, 2, ... along the row
            }
        }
        fdct_islow_1d<true>(d, o);
#pragma unroll
        for (int k = 0; k < 8; k++) s_ws[b * ENC_WS_PITCH + r * 8 + k] = o[k];
    }
    __syncthreads();
    const int c = r;        // column pass: thread c owns column c
    if (!dummy) {
#pragma unroll
        for (int k = 0; k < 8; k++) d[k] = s_ws[b * ENC_WS_PITCH + k * 8 + c];
        fdct_islow_1d<false>(d, o);
        const uint16_t *q = G.q[comp ? 1 : 0];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            // jcdctmgr.c: divisor = q << 3 (the islow output carries a factor 8), round half away from zero
            const int qv = (int)q[k * 8 + c] << 3;
            const int a = abs(o[k]) + (qv >> 1);
            const int v = a / qv;
            s_out[b * 64 + c_enc_zigpos[k * 8 + c]] = (int16_t)(o[k] < 0 ? -v : v);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) s_out[b * 64 + r * 8 + k] = 0;     // dummy block: zero AC; its DC is a rule of the entropy stage
    }
    __syncthreads();
    {   // 32 blocks x 128 B, one int4 per thread
        const int n_valid = min(ENC_BLOCKS_PER_CTA, G.n_blocks - base) * 8;
        if (t < n_valid) ((int4 *)(coef + (size_t)base * 64))[t] = ((const int4 *)s_out)[t];
    }
}

// ---------------------------------------------------------------- entropy coding
// jccoefct.c compress_data: a dummy block at the right edge copies the DC of the block before it; a dummy block ROW copies the DC
// of the block before the row.  Only luma (2x2 blocks per MCU) has dummies; block 0 of an MCU is always real.
__device__ __forceinline__ int enc_effective_dc(const int16_t *__restrict__ coef, const EncGeom &G, int m, int j)
{
    if (G.channels == 1 || j >= 4) return coef[((size_t)m * (G.channels == 1 ? 1 : 6) + j) * 64];
    const int my = m / G.bw, mx = m - my * G.bw;
    const bool row1 = my * 2 + 1 < G.ybh, col1 = mx * 2 + 1 < G.ybw;
    const int16_t *M = coef + (size_t)m * 6 * 64;
    const int dc0 = M[0];
    const int dc1 = col1 ? (int)M[64] : dc0;
    if (j == 0) return dc0;
    if (j == 1) return dc1;
    if (!row1) return dc1;                          // dummy row: both blocks take the DC of the block before the row
    const int dc2 = M[128];
    if (j == 2) return dc2;
    return col1 ? (int)M[192] : dc2;
}

struct BitCount {
    unsigned n = 0;
    __device__ __forceinline__ void put(uint32_t code, int len) { (void)code; n += len; }
};

// MSB-first writer into 32-bit words that hold the stream as big-endian numbers; words are shared with the neighbouring blocks
// only at the two ends of a block, every store is an atomicOr into the zeroed stream
struct BitEmit {
    uint32_t *words;
    unsigned long long acc = 0;
    int fill;                                        // bits of the current word that are occupied (by earlier blocks or acc)
    __device__ __forceinline__ BitEmit(uint32_t *stream, unsigned long long bit_offset) : words(stream + (bit_offset >> 5)), fill((int)(bit_offset & 31)) {}
    __device__ __forceinline__ void put(uint32_t code, int len)
    {
        acc = (acc << len) | (unsigned long long)(code & ((1u << len) - 1u));
        fill += len;
        if (fill >= 32) {
            fill -= 32;
            atomicOr(words++, (uint32_t)(acc >> fill));
            acc &= (1ull << fill) - 1ull;
        }
    }
    __device__ __forceinline__ void flush() { if (fill > 0 && acc) atomicOr(words, (uint32_t)(acc << (32 - fill))); }
};

// jchuff.c encode_one_block
template <class Sink>
__device__ __forceinline__ void enc_one_block(Sink &sink, const int16_t *__restrict__ zz, bool dummy, int dc, int last_dc, int tbl)
{
    int temp = dc - last_dc, temp2 = temp;
    if (temp < 0) { temp = -temp; temp2--; }
    int nbits = 32 - __clz(temp);
    uint32_t e = c_enc_dc[tbl][nbits];
    sink.put(e & 0xffffu, (int)(e >> 16));
    if (nbits) sink.put((uint32_t)temp2, nbits);
    int run = 0;
    if (!dummy) {
#pragma unroll 1
        for (int k = 1; k < 64; k++) {
            temp = zz[k];
            if (temp == 0) { run++; continue; }
            while (run > 15) { e = c_enc_ac[tbl][0xF0]; sink.put(e & 0xffffu, (int)(e >> 16)); run -= 16; }
            temp2 = temp;
            if (temp < 0) { temp = -temp; temp2--; }
            nbits = 32 - __clz(temp);
            e = c_enc_ac[tbl][(run << 4) + nbits];
            sink.put(e & 0xffffu, (int)(e >> 16));
            sink.put((uint32_t)temp2, nbits);
            run = 0;
        }
    } else run = 63;
    if (run > 0) { e = c_enc_ac[tbl][0]; sink.put(e & 0xffffu, (int)(e >> 16)); }
}

// DC of block g and of the previous block of the same component in scan order
__device__ __forceinline__ void enc_block_dcs(const int16_t *__restrict__ coef, const EncGeom &G, int g, int &dc, int &last_dc, int &tbl, bool &dummy)
{
    if (G.channels == 1) {
        dc = coef[(size_t)g * 64]; last_dc = g > 0 ? (int)coef[(size_t)(g - 1) * 64] : 0; tbl = 0; dummy = false;
        return;
    }
    const int m = g / 6, j = g - m * 6;
    int comp, by, bx;
    enc_block_pos(G, g, comp, by, bx, dummy);
    tbl = comp ? 1 : 0;
    dc = enc_effective_dc(coef, G, m, j);
    if (j >= 1 && j <= 3) last_dc = enc_effective_dc(coef, G, m, j - 1);
    else last_dc = m > 0 ? enc_effective_dc(coef, G, m - 1, j == 0 ? 3 : j) : 0;
}

__global__ void __launch_bounds__(256) jpeg_block_bits_kernel(const int16_t *__restrict__ coef, const EncGeom G, uint32_t *__restrict__ bits)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G.n_blocks) return;
    int dc, last_dc, tbl; bool dummy;
    enc_block_dcs(coef, G, g, dc, last_dc, tbl, dummy);
    BitCount sink;
    enc_one_block(sink, coef + (size_t)g * 64, dummy, dc, last_dc, tbl);
    bits[g] = sink.n;
}

__global__ void __launch_bounds__(256) jpeg_emit_kernel(const int16_t *__restrict__ coef, const EncGeom G, const unsigned long long *__restrict__ offs,
                                                        uint32_t *stream)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G.n_blocks) return;
    int dc, last_dc, tbl; bool dummy;
    enc_block_dcs(coef, G, g, dc, last_dc, tbl, dummy);
    BitEmit sink(stream, offs[g]);
    enc_one_block(sink, coef + (size_t)g * 64, dummy, dc, last_dc, tbl);
    sink.flush();
}

// ---------------------------------------------------------------- byte stuffing
__device__ __forceinline__ uint32_t stream_byte(const uint32_t *__restrict__ words, long long i, long long n_bytes, int pad_bits)
{
    uint32_t v = (words[i >> 2] >> (24 - 8 * (int)(i & 3))) & 0xffu;
    if (i == n_bytes - 1 && pad_bits) v |= (1u << pad_bits) - 1u;       // jchuff.c flush_bits: fill the last byte with one-bits
    return v;
}

__global__ void __launch_bounds__(256) jpeg_stuff_count_kernel(const uint32_t *__restrict__ words, long long n_bytes, int pad_bits, uint32_t *__restrict__ counts)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long b0 = t * 16;
    if (b0 >= n_bytes) return;
    uint32_t c = 0;
    for (int k = 0; k < 16 && b0 + k < n_bytes; k++) c += stream_byte(words, b0 + k, n_bytes, pad_bits) == 0xffu;
    counts[t] = c;
}

__global__ void __launch_bounds__(256) jpeg_stuff_write_kernel(const uint32_t *__restrict__ words, long long n_bytes, int pad_bits,
                                                               const unsigned long long *__restrict__ offs, uint8_t *__restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long b0 = t * 16;
    if (b0 >= n_bytes) return;
    uint8_t *dst = out + b0 + offs[t];
    for (int k = 0; k < 16 && b0 + k < n_bytes; k++) {
        const uint32_t v = stream_byte(words, b0 + k, n_bytes, pad_bits);
        *dst++ = (uint8_t)v;
        if (v == 0xffu) *dst++ = 0;
    }
}

// ---------------------------------------------------------------- host: tables, markers, orchestration
// jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline)
static void scaled_quant(const uint8_t *std_tbl, int quality, uint16_t *out)
{
    quality = quality < 1 ? 1 : (quality > 100 ? 100 : quality);
    const int scale = quality < 50 ? 5000 / quality : 200 - quality * 2;
    for (int i = 0; i < 64; i++) {
        long v = ((long)std_tbl[i] * scale + 50) / 100;
        out[i] = (uint16_t)(v < 1 ? 1 : (v > 255 ? 255 : v));
    }
}

static void put_segment(std::vector<uint8_t> &h, int marker, const std::vector<uint8_t> &payload)
{
    h.push_back(0xFF); h.push_back((uint8_t)marker);
    const size_t len = payload.size() + 2;
    h.push_back((uint8_t)(len >> 8)); h.push_back((uint8_t)(len & 255));
    h.insert(h.end(), payload.begin(), payload.end());
}

// jcmarker.c write_file_header / write_frame_header / write_scan_header: SOI, APP0 (JFIF 1.01, aspect 1:1), one DQT per table,
// SOF0, one DHT per table (DC0, AC0[, DC1, AC1]), SOS
static std::vector<uint8_t> enc_header(const EncGeom &G)
{
    std::vector<uint8_t> h = { 0xFF, 0xD8 };
    put_segment(h, 0xE0, { 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0 });
    for (int t = 0; t < (G.channels == 3 ? 2 : 1); t++) {
        std::vector<uint8_t> p = { (uint8_t)t };
        for (int k = 0; k < 64; k++) p.push_back((uint8_t)G.q[t][ZIGZAG_NATURAL[k]]);
        put_segment(h, 0xDB, p);
    }
    std::vector<uint8_t> sof = { 8, (uint8_t)(G.rows >> 8), (uint8_t)(G.rows & 255), (uint8_t)(G.cols >> 8), (uint8_t)(G.cols & 255), (uint8_t)G.channels };
    if (G.channels == 1) sof.insert(sof.end(), { 1, 0x11, 0 });
    else sof.insert(sof.end(), { 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1 });
    put_segment(h, 0xC0, sof);
    for (int t = 0; t < (G.channels == 3 ? 2 : 1); t++) {
        std::vector<uint8_t> p = { (uint8_t)t };
        p.insert(p.end(), DC_BITS[t], DC_BITS[t] + 16); p.insert(p.end(), DC_VALS, DC_VALS + 12);
        put_segment(h, 0xC4, p);
        p.assign(1, (uint8_t)(0x10 | t));
        p.insert(p.end(), AC_BITS[t], AC_BITS[t] + 16); p.insert(p.end(), AC_VALS[t], AC_VALS[t] + 162);
        put_segment(h, 0xC4, p);
    }
    if (G.channels == 1) put_segment(h, 0xDA, { 1, 1, 0x00, 0, 63, 0 });
    else put_segment(h, 0xDA, { 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0 });
    return h;
}

static int jpeg_encode_device_image(vfsms_ctx *ctx, const uint8_t *img_dev, int rows, int cols, int channels, int64_t row_stride, int quality,
                                    uint8_t *out, size_t out_capacity, size_t *out_size, cudaStream_t st)
{
    EncState *s = enc_state(ctx);
    int rc;
    if ((rc = enc_init_tables(s))) return rc;
    EncGeom G;
    memset(&G, 0, sizeof(G));
    G.rows = rows; G.cols = cols; G.channels = channels; G.stride = row_stride;
    G.ybh = (rows + 7) / 8; G.ybw = (cols + 7) / 8; G.crows = (rows + 1) / 2;
    if (channels == 1) { G.bh = G.ybh; G.bw = G.ybw; }
    else { G.bh = (rows + 15) / 16; G.bw = (cols + 15) / 16; }
    const long long n_blocks = (long long)G.bh * G.bw * (channels == 1 ? 1 : 6);
    if (n_blocks > 0x7fffffffLL) { vfsms_set_error("jpeg encode: image too large"); return VFSMS_E_ARG; }
    G.n_blocks = (int)n_blocks;
    scaled_quant(STD_LUMA_Q, quality, G.q[0]);
    scaled_quant(STD_CHROMA_Q, quality, G.q[1]);

    if ((rc = s->coef.reserve((size_t)n_blocks * 64 * 2))) return rc;
    if ((rc = s->bits.reserve((size_t)n_blocks * 4))) return rc;
    if ((rc = s->offs.reserve((size_t)n_blocks * 8))) return rc;
    if ((rc = s->totals.reserve(16))) return rc;
    unsigned long long *totals = s->totals.as<unsigned long long>();
    int16_t *coef = s->coef.as<int16_t>();
    const int grid_b = (int)((n_blocks + 255) / 256);
    jpeg_fdct_quant_kernel<<<(int)((n_blocks + ENC_BLOCKS_PER_CTA - 1) / ENC_BLOCKS_PER_CTA), ENC_BLOCKS_PER_CTA * 8, 0, st>>>(img_dev, G, coef);
    LAUNCH_CHECK(ctx);
    jpeg_block_bits_kernel<<<grid_b, 256, 0, st>>>(coef, G, s->bits.as<uint32_t>()); LAUNCH_CHECK(ctx);
    if ((rc = exclusive_scan(ctx, s->partial, s->bits.as<uint32_t>(), n_blocks, s->offs.as<unsigned long long>(), totals, st))) return rc;
    unsigned long long total_bits = 0;
    CUDA_TRY(cudaMemcpyAsync(&total_bits, totals, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));

    const long long n_bytes = (long long)((total_bits + 7) / 8);
    const int pad_bits = (int)((8 - (total_bits & 7)) & 7);
    const size_t n_words = (size_t)(n_bytes + 3) / 4 + 2;
    if ((rc = s->stream.reserve(n_words * 4))) return rc;
    CUDA_TRY(cudaMemsetAsync(s->stream.p, 0, n_words * 4, st));
    jpeg_emit_kernel<<<grid_b, 256, 0, st>>>(coef, G, s->offs.as<unsigned long long>(), s->stream.as<uint32_t>()); LAUNCH_CHECK(ctx);

    const long long n16 = (n_bytes + 15) / 16;
    if ((rc = s->counts.reserve((size_t)n16 * 4))) return rc;
    if ((rc = s->offs2.reserve((size_t)n16 * 8))) return rc;
    const int grid_s = (int)((n16 + 255) / 256);
    jpeg_stuff_count_kernel<<<grid_s, 256, 0, st>>>(s->stream.as<uint32_t>(), n_bytes, pad_bits, s->counts.as<uint32_t>()); LAUNCH_CHECK(ctx);
    if ((rc = exclusive_scan(ctx, s->partial, s->counts.as<uint32_t>(), n16, s->offs2.as<unsigned long long>(), totals + 1, st))) return rc;
    unsigned long long n_ff = 0;
    CUDA_TRY(cudaMemcpyAsync(&n_ff, totals + 1, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));

    const std::vector<uint8_t> header = enc_header(G);
    const size_t body = (size_t)n_bytes + (size_t)n_ff;
    const size_t total = header.size() + body + 2;
    if (out_size) *out_size = total;
    if (!out || out_capacity < total) { vfsms_set_error("jpeg encode: output capacity %zu < %zu", out_capacity, total); return VFSMS_E_CAPACITY; }
    if ((rc = s->out.reserve(body + 16))) return rc;
    jpeg_stuff_write_kernel<<<grid_s, 256, 0, st>>>(s->stream.as<uint32_t>(), n_bytes, pad_bits, s->offs2.as<unsigned long long>(), s->out.as<uint8_t>());
    LAUNCH_CHECK(ctx);
    memcpy(out, header.data(), header.size());
    CUDA_TRY(cudaMemcpyAsync(out + header.size(), s->out.p, body, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    out[header.size() + body] = 0xFF; out[header.size() + body + 1] = 0xD9;
    return 0;
}

static int enc_check_args(const void *img, int rows, int cols, int channels, int64_t row_stride, const size_t *out_size)
{
    if (!img || !out_size || rows < 1 || cols < 1 || rows > 65535 || cols > 65535 || (channels != 1 && channels != 3) ||
        row_stride < (int64_t)cols * channels) {
        vfsms_set_error("jpeg encode: bad arguments (rows, cols in 1..65535, channels 1 or 3, row_stride >= cols * channels)");
        return VFSMS_E_ARG;
    }
    return 0;
}

extern "C" int vfsms_jpeg_encode_dev(vfsms_ctx *ctx, const uint8_t *img_dev, int rows, int cols, int channels, int64_t row_stride, int quality,
                                     uint8_t *out, size_t out_capacity, size_t *out_size, void *stream)
{
    if (!ctx) { vfsms_set_error("jpeg encode: no context"); return VFSMS_E_ARG; }
    int rc = enc_check_args(img_dev, rows, cols, channels, row_stride, out_size);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    return jpeg_encode_device_image(ctx, img_dev, rows, cols, channels, row_stride, quality, out, out_capacity, out_size,
                                    stream ? (cudaStream_t)stream : ctx->stream);
}

extern "C" int vfsms_jpeg_encode_host(vfsms_ctx *ctx, const uint8_t *img, int rows, int cols, int channels, int64_t row_stride, int quality,
                                      uint8_t *out, size_t out_capacity, size_t *out_size)
{
    if (!ctx) { vfsms_set_error("jpeg encode: no context"); return VFSMS_E_ARG; }
    int rc = enc_check_args(img, rows, cols, channels, row_stride, out_size);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    EncState *s = enc_state(ctx);
    const size_t line = (size_t)cols * channels;
    if ((rc = s->img.reserve(line * rows))) return rc;
    CUDA_TRY(cudaMemcpy2DAsync(s->img.p, line, img, (size_t)row_stride, line, (size_t)rows, cudaMemcpyHostToDevice, ctx->stream));
    return jpeg_encode_device_image(ctx, s->img.as<uint8_t>(), rows, cols, channels, (int64_t)line, quality, out, out_capacity, out_size, ctx->stream);
}
