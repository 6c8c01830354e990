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
#include "scan.cuh"
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

// Orientation tag of an APP1 Exif segment (payload `s`, `n` bytes), 0 when absent or malformed.
static int exif_orientation(const uint8_t *s, size_t n)
{
    if (n < 14 || memcmp(s, "Exif\0\0", 6) != 0) return 0;
    const uint8_t *t = s + 6; const size_t tn = n - 6;
    const bool le = t[0] == 'I' && t[1] == 'I', be = t[0] == 'M' && t[1] == 'M';
    if (!le && !be) return 0;
    auto u16 = [&](size_t o) -> unsigned { return le ? (unsigned)(t[o] | (t[o + 1] << 8)) : (unsigned)((t[o] << 8) | t[o + 1]); };
    auto u32 = [&](size_t o) -> unsigned { return le ? (unsigned)(t[o] | (t[o + 1] << 8) | (t[o + 2] << 16) | ((unsigned)t[o + 3] << 24))
                                                     : (unsigned)(((unsigned)t[o] << 24) | (t[o + 1] << 16) | (t[o + 2] << 8) | t[o + 3]); };
    if (u16(2) != 42) return 0;
    const size_t ifd = u32(4);
    if (ifd + 2 > tn) return 0;
    const unsigned cnt = u16(ifd);
    for (unsigned k = 0; k < cnt; k++) {
        const size_t e = ifd + 2 + 12 * (size_t)k;
        if (e + 12 > tn) return 0;
        if (u16(e) == 0x0112) return (int)(u16(e + 2) == 3 ? u16(e + 8) : u32(e + 8));      // SHORT in the value field
    }
    return 0;
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
        } else if (m == 0xE1) {
            // Exif orientation (TIFF tag 0x0112 of IFD0): cv2.imdecode rotates / flips the decoded image for values 2..8 (the
            // reference never sets IMREAD_IGNORE_ORIENTATION).  Such files are left to the caller's cv2 fallback.
            const int orient = exif_orientation(s, sl);
            if (orient > 1 && orient <= 8) {
                vfsms_set_error("jpeg: Exif orientation %d (cv2.imdecode would rotate / flip the image) is not supported", orient);
                return VFSMS_E_UNSUPPORTED;
            }
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


// ---------------------------------------------------------------- device entropy decoding (VFSMS_OPT_ENTROPY = 1)
// The host Huffman stage sets the ingest rate (one thread per file, 20+ ms per 4-MPix tile).  A Huffman stream can be entered at
// any bit: a decoder started in the wrong state re-synchronises with the true symbol boundaries after a few codes
// (self-synchronisation; Weissenberger & Schmidt, "Massively parallel Huffman decoding on GPUs", ICPP 2018, and their JPEG
// decoder, ICPP 2021).  Schedule:
//   host        the scan is copied without its stuffing bytes (0xFF 0x00 -> 0xFF) into pinned memory: 1/6 of the coefficient bytes
//               the host stage ships -- the only part that stays on the CPU (memchr speed)
//   sub_init    the bit stream of every file is cut into subsequences of HUFF_SUB_BITS bits; one thread decodes each from its first
//               bit in state (block 0 of the MCU, coefficient 0) up to its end and records where and in which state it left
//               (exit position, block-in-MCU, coefficient index, blocks completed)
//   sub_sync    subsequence s is decoded again from the exit of s - 1; passes repeat until none changed (subsequence 0 starts in
//               the true state, so the exits become true front to back; a pass without change is the fixed point = the serial
//               decode).  A thread whose predecessor's exit did not change skips the pass.
//   scan        blocks completed per subsequence -> index of the first block of every subsequence (scan.cuh)
//   sub_write   every subsequence is decoded a last time from its true entry and writes its coefficients (int16, natural order,
//               the layout of the host stage) and the DC DIFFERENCES
//   dc_*        per component: sums of the differences per MCU, prefix sum over the MCUs, absolute DC values
// Same result as jpeg_entropy_decode, symbol for symbol (the decode step below mirrors BitReader::symbol / receive_extend,
// including what they do on corrupt data).  Files with restart intervals take the host stage.
__constant__ uint8_t c_huff_zigzag[64];
#define HUFF_SUB_BITS 1024
#define HUFF_MAX_MCU_BLOCKS 10

static int host_threads();

struct DevHuff { uint16_t fast[512]; int32_t maxcode[18]; int32_t valptr[17]; uint8_t vals[256]; };

struct DevScan {
    unsigned long long byte_base;            // this file's unstuffed scan inside the stream buffer (multiple of 4)
    unsigned int n_bits;                     // length of the scan in bits
    int first_sub, n_sub;                    // its subsequences in the flattened arrays
    int mcu_blocks, n_blocks, mcux, n_mcu;
    int first_mcu_slot;                      // its slots in the per-(component, MCU) DC arrays: [first_mcu_slot + c * n_mcu + m]
    int ncomp;
    uint8_t blk_comp[HUFF_MAX_MCU_BLOCKS], blk_v[HUFF_MAX_MCU_BLOCKS], blk_h[HUFF_MAX_MCU_BLOCKS];
    uint8_t comp_h[4], comp_v[4], dc_tab[4], ac_tab[4];          // tables: indices into this file's DevHuff[8] (dc 0-3, ac 4-7)
    int comp_bw[4];                          // blocks per row of component c (mcux * h)
    long long comp_off[4];                   // int16 offset of component c's plane from the file's coefficient base, -1 = parse and drop
    long long coef_base;                     // int16 offset of the file's coefficients in the device buffer
    int table_base;                          // index of this file's first DevHuff
};

__device__ __forceinline__ uint32_t huff_bits32(const uint32_t *__restrict__ words, unsigned int pos)
{
    const uint32_t a = __byte_perm(words[pos >> 5], 0, 0x0123), b = __byte_perm(words[(pos >> 5) + 1], 0, 0x0123);
    const int sh = (int)(pos & 31);
    return sh ? (a << sh) | (b >> (32 - sh)) : a;
}

// One code (+ its magnitude bits).  State: b = block inside the MCU, k = next coefficient (0: the DC code comes next).
// Returns true when a block was completed.  WRITE: blk points at the current block (nullptr when its component is dropped).
template <bool WRITE>
__device__ __forceinline__ bool huff_step(const uint32_t *__restrict__ words, const DevHuff *__restrict__ tabs, const DevScan &F,
                                          unsigned int &pos, int &b, int &k, int16_t *blk)
{
    const int c = F.blk_comp[b];
    const DevHuff &T = tabs[F.table_base + (k == 0 ? F.dc_tab[c] : 4 + F.ac_tab[c])];
    uint32_t w = huff_bits32(words, pos);
    int sym, len;
    const uint16_t e = T.fast[w >> 23];
    if (e != 0xFFFF) { len = e >> 8; sym = e & 255; }
    else {
        len = 16; sym = 0;                                  // corrupt stream: skip 16 bits like the host stage
#pragma unroll 1
        for (int l = 10; l <= 16; l++) {
            const int code = (int)(w >> (32 - l));
            if (code <= T.maxcode[l]) { len = l; sym = T.vals[(code + T.valptr[l]) & 255]; break; }
        }
    }
    pos += len;
    bool done = false;
    if (k == 0) {
        const int s = sym & 15;
        int v = 0;
        if (s) {
            w = huff_bits32(words, pos);
            v = (int)(w >> (32 - s));
            if (v < (1 << (s - 1))) v = v - (1 << s) + 1;
            pos += s;
        }
        if (WRITE && blk) blk[0] = (int16_t)v;              // the DC DIFFERENCE; dc_apply_kernel turns it into the value
        k = 1;
    } else {
        const int r = sym >> 4, sz = sym & 15;
        if (sz == 0) {
            if (r == 15) { k += 16; done = k >= 64; }
            else done = true;                               // end of block
        } else {
            k += r;
            w = huff_bits32(words, pos);
            int v = (int)(w >> (32 - sz));
            if (v < (1 << (sz - 1))) v = v - (1 << sz) + 1;
            pos += sz;
            if (WRITE && blk && k < 64) blk[c_huff_zigzag[k]] = (int16_t)v;
            k++;
            done = k >= 64;
        }
    }
    if (done) { k = 0; b = b + 1 == F.mcu_blocks ? 0 : b + 1; }
    return done;
}


// exit record of a subsequence (one 64-bit word, so that it is read and written in one piece): low word = exit bit position
// (relative to the file), high word = blocks completed << 16 | b << 6 | k
typedef unsigned long long huff_rec;
__device__ __forceinline__ huff_rec huff_make_rec(unsigned int pos, unsigned int meta) { return (huff_rec)pos | ((huff_rec)meta << 32); }
__device__ __forceinline__ huff_rec huff_decode_sub(const uint32_t *__restrict__ words, const DevHuff *__restrict__ tabs, const DevScan &F,
                                                 unsigned int pos, int b, int k, unsigned int end)
{
    int nblk = 0;
    while (pos < end) nblk += huff_step<false>(words, tabs, F, pos, b, k, nullptr) ? 1 : 0;
    return huff_make_rec(pos, ((unsigned)nblk << 16) | ((unsigned)b << 6) | (unsigned)k);
}

__device__ __forceinline__ const DevScan &huff_file_of(const DevScan *__restrict__ files, int n_files, int s)
{
    int lo = 0, hi = n_files;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (files[mid].first_sub <= s) lo = mid; else hi = mid; }
    return files[lo];
}

__global__ void __launch_bounds__(128) huff_sub_init_kernel(const uint8_t *__restrict__ stream, const DevHuff *__restrict__ tabs,
                                                           const DevScan *__restrict__ files, int n_files, int n_sub, huff_rec *exits, huff_rec *used)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sub) return;
    const DevScan &F = huff_file_of(files, n_files, s);
    const int sl = s - F.first_sub;
    const uint32_t *words = (const uint32_t *)(stream + F.byte_base);
    const unsigned int begin = (unsigned int)sl * HUFF_SUB_BITS, end = min(begin + HUFF_SUB_BITS, F.n_bits);
    exits[s] = huff_decode_sub(words, tabs, F, begin, 0, 0, end);
    used[s] = huff_make_rec(begin, 0);                       // the entry this exit was computed from
}

__global__ void __launch_bounds__(128) huff_sub_sync_kernel(const uint8_t *__restrict__ stream, const DevHuff *__restrict__ tabs,
                                                           const DevScan *__restrict__ files, int n_files, int n_sub, huff_rec *exits, huff_rec *used,
                                                           int *changed)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sub) return;
    const DevScan &F = huff_file_of(files, n_files, s);
    const int sl = s - F.first_sub;
    if (sl == 0) return;                                     // the first subsequence of a file starts in the true state
    // the records are single 64-bit words: a concurrent update of the predecessor is seen either before or after, both are exits
    // some pass has produced, and the pass that sees no change anywhere ends the iteration
    const huff_rec prev = *(volatile huff_rec *)(exits + s - 1);
    const huff_rec entry = prev & 0x0000ffffffffffffull;     // position + state, without the block count
    if (used[s] == entry) return;                            // same entry as last time: same exit
    const uint32_t *words = (const uint32_t *)(stream + F.byte_base);
    const unsigned int begin = (unsigned int)sl * HUFF_SUB_BITS, end = min(begin + HUFF_SUB_BITS, F.n_bits);
    const unsigned int meta = (unsigned int)(entry >> 32);
    const unsigned int pos = max((unsigned int)entry, begin);   // (an exit never lies before the next subsequence's first bit)
    const huff_rec e = huff_decode_sub(words, tabs, F, pos, (int)((meta >> 6) & 1023), (int)(meta & 63), end);
    used[s] = entry;
    if (exits[s] != e) { *(volatile huff_rec *)(exits + s) = e; *changed = 1; }
}

__global__ void __launch_bounds__(256) huff_sub_count_kernel(const huff_rec *__restrict__ exits, int n_sub, uint32_t *__restrict__ counts)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_sub) counts[s] = (uint32_t)(exits[s] >> 48);
}

__global__ void __launch_bounds__(128) huff_sub_write_kernel(const uint8_t *__restrict__ stream, const DevHuff *__restrict__ tabs,
                                                            const DevScan *__restrict__ files, int n_files, int n_sub,
                                                            const huff_rec *__restrict__ exits, const unsigned long long *__restrict__ first_block,
                                                            int16_t *__restrict__ coef)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sub) return;
    const DevScan &F = huff_file_of(files, n_files, s);
    const int sl = s - F.first_sub;
    const uint32_t *words = (const uint32_t *)(stream + F.byte_base);
    const unsigned int begin = (unsigned int)sl * HUFF_SUB_BITS, end = min(begin + HUFF_SUB_BITS, F.n_bits);
    unsigned int pos = begin;
    int b = 0, k = 0;
    if (sl > 0) {
        const huff_rec prev = exits[s - 1];
        const unsigned int meta = (unsigned int)(prev >> 32);
        pos = max((unsigned int)prev, begin); b = (int)((meta >> 6) & 1023); k = (int)(meta & 63);
    }
    long long g = (long long)(first_block[s] - first_block[F.first_sub]);     // block of the file this subsequence starts in
    auto block_ptr = [&](long long gi, int bi) -> int16_t * {
        const int c = F.blk_comp[bi];
        if (F.comp_off[c] < 0) return nullptr;
        const long long m = gi / F.mcu_blocks;
        const int my = (int)(m / F.mcux), mx = (int)(m - (long long)my * F.mcux);
        const long long row = (long long)my * F.comp_v[c] + F.blk_v[bi], col = (long long)mx * F.comp_h[c] + F.blk_h[bi];
        return coef + F.coef_base + F.comp_off[c] + (row * F.comp_bw[c] + col) * 64;
    };
    int16_t *blk = g < F.n_blocks ? block_ptr(g, b) : nullptr;
    while (pos < end && g < F.n_blocks) {
        if (huff_step<true>(words, tabs, F, pos, b, k, blk)) { g++; blk = g < F.n_blocks ? block_ptr(g, b) : nullptr; }
    }
}

// DC predictors: blk[0] holds the difference against the previous block of the component in scan order
__global__ void __launch_bounds__(256) huff_dc_sums_kernel(const DevScan *__restrict__ files, int n_files, const int16_t *__restrict__ coef,
                                                          uint32_t *__restrict__ sums)
{
    const int f = blockIdx.y;
    const DevScan &F = files[f];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F.ncomp * F.n_mcu; i += gridDim.x * blockDim.x) {
        const int c = i / F.n_mcu, m = i - c * F.n_mcu;
        int sum = 0;
        if (F.comp_off[c] >= 0) {
            const int my = m / F.mcux, mx = m - my * F.mcux;
            for (int v = 0; v < F.comp_v[c]; v++)
                for (int h = 0; h < F.comp_h[c]; h++)
                    sum += coef[F.coef_base + F.comp_off[c] + ((long long)(my * F.comp_v[c] + v) * F.comp_bw[c] + mx * F.comp_h[c] + h) * 64];
        }
        sums[F.first_mcu_slot + i] = (uint32_t)sum;
    }
}

__global__ void __launch_bounds__(256) huff_dc_apply_kernel(const DevScan *__restrict__ files, int n_files, int16_t *__restrict__ coef,
                                                           const unsigned long long *__restrict__ prefix)
{
    const int f = blockIdx.y;
    const DevScan &F = files[f];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < F.ncomp * F.n_mcu; i += gridDim.x * blockDim.x) {
        const int c = i / F.n_mcu, m = i - c * F.n_mcu;
        if (F.comp_off[c] < 0) continue;
        // two's-complement sums: the low 32 bits of the 64-bit prefix difference are the signed predictor
        int pred = (int)(uint32_t)(prefix[F.first_mcu_slot + i] - prefix[F.first_mcu_slot + c * F.n_mcu]);
        const int my = m / F.mcux, mx = m - my * F.mcux;
        for (int v = 0; v < F.comp_v[c]; v++)
            for (int h = 0; h < F.comp_h[c]; h++) {
                int16_t *p = coef + F.coef_base + F.comp_off[c] + ((long long)(my * F.comp_v[c] + v) * F.comp_bw[c] + mx * F.comp_h[c] + h) * 64;
                pred += *p;
                *p = (int16_t)pred;
            }
    }
}

struct HuffDevState {
    bool zigzag = false;
    DevBuf stream, tabs, files, exits, used, counts, first_block, partial, changed, dc_sums, dc_prefix, totals;
    HostBuf pinned;
};

static HuffDevState *huff_state(vfsms_ctx *ctx)
{
    if (!ctx->jpeg_huff_state) ctx->jpeg_huff_state = new HuffDevState();
    return (HuffDevState *)ctx->jpeg_huff_state;
}

void jpeg_huff_state_destroy(vfsms_ctx *ctx)
{
    HuffDevState *s = (HuffDevState *)ctx->jpeg_huff_state;
    if (!s) return;
    DevBuf *bufs[] = { &s->stream, &s->tabs, &s->files, &s->exits, &s->used, &s->counts, &s->first_block, &s->partial, &s->changed, &s->dc_sums,
                       &s->dc_prefix, &s->totals };
    for (DevBuf *b : bufs) b->release();
    s->pinned.release();
    delete s;
    ctx->jpeg_huff_state = nullptr;
}

// scan bytes of one file without the stuffing zeros; stops at the first marker.  Returns the number of bytes written.
static size_t jpeg_unstuff(const uint8_t *d, size_t begin, size_t n, uint8_t *out)
{
    size_t i = begin, o = 0;
    while (i < n) {
        const uint8_t *ff = (const uint8_t *)memchr(d + i, 0xFF, n - i);
        const size_t run = ff ? (size_t)(ff - (d + i)) : n - i;
        memcpy(out + o, d + i, run);
        o += run; i += run;
        if (!ff) break;
        if (i + 1 < n && d[i + 1] == 0) { out[o++] = 0xFF; i += 2; }
        else break;                                          // a marker (EOI): end of the entropy-coded data
    }
    return o;
}

// Entropy-decode files [0, cn) of a chunk on the device: coefficients of file j at coef_dev + per_i16 * j, laid out like the host stage
// (component planes one after the other; chroma parsed and dropped unless want_chroma).  All files must have restart_interval == 0.
// Returns 0, a negative VFSMS_E_* code, or 1 when a stream ends early and the chunk has to take the host stage.
static int jpeg_entropy_decode_device(vfsms_ctx *ctx, int cn, const uint8_t *const *data, const size_t *sizes, const JpegHeader *H,
                                      bool want_chroma, int16_t *coef_dev, size_t per_i16, cudaStream_t st)
{
    HuffDevState *S = huff_state(ctx);
    int rc;
    if (!S->zigzag) { CUDA_TRY(cudaMemcpyToSymbol(c_huff_zigzag, kZigzag, 64)); S->zigzag = true; }
    // layout of the pinned staging buffer: [DevScan x cn][DevHuff x 8 cn][streams, each padded to 4 bytes + 16 zero bytes]
    std::vector<size_t> base((size_t)cn + 1);
    const size_t files_bytes = (size_t)cn * sizeof(DevScan), tabs_bytes = (size_t)cn * 8 * sizeof(DevHuff);
    size_t total = 0;
    for (int j = 0; j < cn; j++) { base[j] = total; total += ((sizes[j] - H[j].scan_begin + 3) & ~(size_t)3) + 16; }
    base[cn] = total;
    if ((rc = S->pinned.reserve(files_bytes + tabs_bytes + total))) return rc;
    DevScan *files = (DevScan *)S->pinned.p;
    DevHuff *tabs = (DevHuff *)((uint8_t *)S->pinned.p + files_bytes);
    uint8_t *streams = (uint8_t *)S->pinned.p + files_bytes + tabs_bytes;
    memset(streams, 0, total);
    std::vector<size_t> lens((size_t)cn);
    {
        std::atomic<int> next(0);
        auto work = [&]() { for (int j = next.fetch_add(1); j < cn; j = next.fetch_add(1)) lens[j] = jpeg_unstuff(data[j], H[j].scan_begin, sizes[j], streams + base[j]); };
        const int workers = host_threads() < cn ? host_threads() : cn;
        std::vector<std::thread> pool;
        for (int w = 1; w < workers; w++) pool.emplace_back(work);
        work();
        for (auto &th : pool) th.join();
    }
    int n_sub = 0, n_slots = 0;
    for (int j = 0; j < cn; j++) {
        const JpegHeader &h = H[j];
        DevScan &F = files[j];
        memset(&F, 0, sizeof(F));
        if (lens[j] * 8 > 0xfffffff0ull) { vfsms_set_error("jpeg: scan too long for the device entropy stage"); return VFSMS_E_UNSUPPORTED; }
        F.byte_base = base[j]; F.n_bits = (unsigned int)(lens[j] * 8);
        F.first_sub = n_sub; F.n_sub = (int)((F.n_bits + HUFF_SUB_BITS - 1) / HUFF_SUB_BITS);
        if (F.n_sub < 1) F.n_sub = 1;
        n_sub += F.n_sub;
        F.mcux = h.mcux; F.n_mcu = h.mcux * h.mcuy; F.ncomp = h.ncomp; F.table_base = 8 * j;
        F.first_mcu_slot = n_slots; n_slots += F.ncomp * F.n_mcu;
        long long off = 0;
        int nb = 0;
        for (int c = 0; c < h.ncomp; c++) {
            const JpegComp &C = h.comp[c];
            F.comp_h[c] = (uint8_t)C.h; F.comp_v[c] = (uint8_t)C.v; F.dc_tab[c] = (uint8_t)C.td; F.ac_tab[c] = (uint8_t)C.ta;
            F.comp_bw[c] = h.mcux * C.h;
            const long long plane = (long long)(h.mcuy * C.v) * (h.mcux * C.h) * 64;
            if (c == 0 || want_chroma) { F.comp_off[c] = off; off += plane; } else F.comp_off[c] = -1;
            for (int v = 0; v < C.v; v++)
                for (int hh = 0; hh < C.h; hh++) {
                    if (nb >= HUFF_MAX_MCU_BLOCKS) { vfsms_set_error("jpeg: more than 10 blocks per MCU"); return VFSMS_E_UNSUPPORTED; }
                    F.blk_comp[nb] = (uint8_t)c; F.blk_v[nb] = (uint8_t)v; F.blk_h[nb] = (uint8_t)hh; nb++;
                }
        }
        for (int t = 0; t < 4; t++) {
            const HuffTable *src[2] = { &h.dc[t], &h.ac[t] };
            for (int q = 0; q < 2; q++) {
                DevHuff &D = tabs[8 * j + 4 * q + t];
                memcpy(D.fast, src[q]->fast, sizeof(D.fast)); memcpy(D.maxcode, src[q]->maxcode, sizeof(D.maxcode));
                memcpy(D.valptr, src[q]->valptr, sizeof(D.valptr)); memcpy(D.vals, src[q]->vals, sizeof(D.vals));
            }
        }
        F.mcu_blocks = nb; F.n_blocks = nb * F.n_mcu;
        F.coef_base = (long long)(per_i16 * (size_t)j);
    }
    const size_t up = files_bytes + tabs_bytes + total;
    if ((rc = S->stream.reserve(up))) return rc;
    if ((rc = S->exits.reserve((size_t)n_sub * 8))) return rc;
    if ((rc = S->used.reserve((size_t)n_sub * 8))) return rc;
    if ((rc = S->counts.reserve((size_t)n_sub * 4))) return rc;
    if ((rc = S->first_block.reserve((size_t)n_sub * 8))) return rc;
    if ((rc = S->changed.reserve(16))) return rc;
    if ((rc = S->totals.reserve(16))) return rc;
    if ((rc = S->dc_sums.reserve((size_t)n_slots * 4))) return rc;
    if ((rc = S->dc_prefix.reserve((size_t)n_slots * 8))) return rc;
    CUDA_TRY(cudaMemcpyAsync(S->stream.p, S->pinned.p, up, cudaMemcpyHostToDevice, st));
    const DevScan *files_dev = (const DevScan *)S->stream.p;
    const DevHuff *tabs_dev = (const DevHuff *)((const uint8_t *)S->stream.p + files_bytes);
    const uint8_t *stream_dev = (const uint8_t *)S->stream.p + files_bytes + tabs_bytes;
    huff_rec *exits = S->exits.as<huff_rec>(), *used = S->used.as<huff_rec>();
    const int grid = (n_sub + 127) / 128;
    huff_sub_init_kernel<<<grid, 128, 0, st>>>(stream_dev, tabs_dev, files_dev, cn, n_sub, exits, used); LAUNCH_CHECK(ctx);
    int *changed = S->changed.as<int>();
    // Every pass makes at least one more subsequence of each file final, so n_sub passes always suffice; in practice the decoders
    // re-synchronise within a few subsequences.  Passes are launched in rounds (4, 8, 16, 32, 32, ...) with one flag read-back per
    // round: the round in which no thread changed its exit is the fixed point.
    for (int done = 0, per_round = 4;;) {
        if (done > n_sub + 64) { vfsms_set_error("jpeg: device entropy stage did not reach its fixed point"); return VFSMS_E_CUDA; }
        CUDA_TRY(cudaMemsetAsync(changed, 0, 4, st));
        for (int r = 0; r < per_round; r++) {
            huff_sub_sync_kernel<<<grid, 128, 0, st>>>(stream_dev, tabs_dev, files_dev, cn, n_sub, exits, used, changed); LAUNCH_CHECK(ctx);
        }
        int flag = 0;
        CUDA_TRY(cudaMemcpyAsync(&flag, changed, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        done += per_round;
        if (!flag) { ctx->entropy_passes = done; break; }
        if (per_round < 32) per_round *= 2;
    }
    huff_sub_count_kernel<<<(n_sub + 255) / 256, 256, 0, st>>>(exits, n_sub, S->counts.as<uint32_t>()); LAUNCH_CHECK(ctx);
    if ((rc = exclusive_scan(ctx, S->partial, S->counts.as<uint32_t>(), n_sub, S->first_block.as<unsigned long long>(), S->totals.as<unsigned long long>(), st))) return rc;
    {   // a stream that ends before its last block (truncated / damaged file): the host stage, which keeps decoding zero bits like
        // libjpeg's warning path, defines the result for those -- tell the caller to use it for this chunk
        std::vector<unsigned long long> fb((size_t)cn + 1);
        for (int j = 0; j < cn; j++)
            CUDA_TRY(cudaMemcpyAsync(&fb[j], S->first_block.as<unsigned long long>() + files[j].first_sub, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(&fb[cn], S->totals.p, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (int j = 0; j < cn; j++) if (fb[j + 1] - fb[j] < (unsigned long long)files[j].n_blocks) return 1;
    }
    CUDA_TRY(cudaMemsetAsync(coef_dev, 0, per_i16 * 2 * (size_t)cn, st));
    huff_sub_write_kernel<<<grid, 128, 0, st>>>(stream_dev, tabs_dev, files_dev, cn, n_sub, exits, S->first_block.as<unsigned long long>(), coef_dev);
    LAUNCH_CHECK(ctx);
    int max_slots = 0;
    for (int j = 0; j < cn; j++) max_slots = std::max(max_slots, files[j].ncomp * files[j].n_mcu);
    const dim3 dc_grid((unsigned)std::min((max_slots + 255) / 256, 1024), (unsigned)cn);
    huff_dc_sums_kernel<<<dc_grid, 256, 0, st>>>(files_dev, cn, coef_dev, S->dc_sums.as<uint32_t>()); LAUNCH_CHECK(ctx);
    if ((rc = exclusive_scan(ctx, S->partial, S->dc_sums.as<uint32_t>(), n_slots, S->dc_prefix.as<unsigned long long>(), S->totals.as<unsigned long long>() + 1, st))) return rc;
    huff_dc_apply_kernel<<<dc_grid, 256, 0, st>>>(files_dev, cn, coef_dev, S->dc_prefix.as<unsigned long long>()); LAUNCH_CHECK(ctx);
    return 0;
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
    // VFSMS_OPT_ENTROPY = 1: Huffman decoding on the device (files with restart intervals keep the host stage)
    bool dev_entropy = ctx->entropy_mode == 1;
    for (int i = 0; i < n; i++) dev_entropy = dev_entropy && H[i].restart_interval == 0;
    for (int c0 = 0; c0 < n; c0 += chunk) {
        const int cn = n - c0 < chunk ? n - c0 : chunk;
        CUDA_TRY(cudaStreamSynchronize(st));          // the previous chunk has left the pinned buffer
        bool chunk_dev = dev_entropy;
        if (chunk_dev) {
            rc = jpeg_entropy_decode_device(ctx, cn, data + c0, sizes + c0, &H[c0], channels == 3, ctx->jpeg_coef.as<int16_t>(), per / 2, st);
            if (rc < 0) return rc;
            chunk_dev = rc == 0;
        }
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
        if (!chunk_dev) {
            std::vector<std::thread> pool;
            for (int w = 1; w < (cn < workers ? cn : workers); w++) pool.emplace_back(work);
            work();
            for (auto &th : pool) th.join();
        }
        for (int j = 0; j < cn; j++) {
            const JpegHeader &h = H[c0 + j];
            const bool colour = channels == 3 && h.ncomp == 3;
            size_t blocks = comp_blocks(h, 0) + (colour ? 2 * comp_blocks(h, 1) : 0);
            int16_t *cdev = (int16_t *)((uint8_t *)ctx->jpeg_coef.p + per * j);
            if (!chunk_dev) CUDA_TRY(cudaMemcpyAsync(cdev, (uint8_t *)ctx->jpeg_pinned.p + per * j, blocks * 128, cudaMemcpyHostToDevice, st));
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

extern "C" int vfsms_jpeg_last_entropy_passes(vfsms_ctx *ctx, int *passes_out)
{
    if (!ctx || !passes_out) { vfsms_set_error("vfsms_jpeg_last_entropy_passes: bad arguments"); return VFSMS_E_ARG; }
    *passes_out = ctx->entropy_passes;
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
