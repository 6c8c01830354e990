"""TEST INFRASTRUCTURE ONLY (imported by tests/ only; the product never calls it).

CPU restatement of the baseline JPEG encoder behind `cv2.imwrite(path, stitchResult)` (the reference writes every mosaic
through it: Stitcher.py:130-131, :196-197) = OpenCV's grfmt_jpeg.cpp on libjpeg(-turbo) with cv2's defaults: quality 95,
sequential Huffman with the Annex-K tables (no optimisation), no restart markers, 4:2:0 for colour, JFIF 1.01 header.
libjpeg-turbo is a third-party dependency that is not vendored in /root/reference (the reference pins opencv-python
3.3.1.11, requirements.txt); the published algorithm is restated here from the libjpeg sources' documented behaviour:
  jcparam.c   jpeg_set_quality / jpeg_quality_scaling, std_luminance / std_chrominance tables, std Huffman tables (K.3-K.6)
  jccolor.c   rgb_ycc_convert (16-bit fixed point)
  jcsample.c  h2v2_downsample (alternating bias 1, 2), expand_right_edge;  jcprepct.c expand_bottom_edge
  jfdctint.c  jpeg_fdct_islow (CONST_BITS 13, PASS1_BITS 2), level shift 128
  jcdctmgr.c  quantisation: sign * ((|c| + (q*8 >> 1)) / (q*8))
  jccoefct.c  dummy blocks at the right / bottom edge (zero AC, DC of the previous block of the MCU)
  jchuff.c    encode_one_block, byte stuffing, final padding with one-bits
  jcmarker.c  marker order SOI APP0 DQT.. SOF0 DHT.. SOS
PINNED (tests/test_jpeg_encode_cpu.py): byte-identical to cv2.imencode(".jpg", img[, quality]) of this container's cv2 4.13 /
libjpeg-turbo 3.1.2 for gray and BGR images over odd / tiny / large sizes, flat, noise and natural content, several qualities.
"""
import numpy as np

STD_LUMA_Q = np.array([
    16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
    18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99],
    np.int64)
STD_CHROMA_Q = np.array([
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99], np.int64)

# zigzag[k] = natural index of the k-th coefficient in zigzag order
ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])

# Annex K.3 - K.6: (bits[1..16], values)
DC_LUMA = ([0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], list(range(12)))
DC_CHROMA = ([0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0], list(range(12)))
AC_LUMA = ([0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d], [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08,
    0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
    0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa])
AC_CHROMA = ([0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77], [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91,
    0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
    0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa])


def quant_table(std, quality):
    """jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline = TRUE).  Natural order."""
    quality = min(max(int(quality), 1), 100)
    scale = 5000 // quality if quality < 50 else 200 - quality * 2
    return np.clip((std * scale + 50) // 100, 1, 255)


def huff_codes(bits, values):
    """jchuff.c jpeg_make_c_derived_tbl: symbol -> (code, length)."""
    table = {}
    code, k = 0, 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            table[values[k]] = (code, length)
            code += 1
            k += 1
        code <<= 1
    return table


def rgb_to_ycc(bgr):
    """jccolor.c rgb_ycc_convert.  -> three int64 planes."""
    def fix(x):
        return int(x * 65536 + 0.5)
    b = bgr[..., 0].astype(np.int64); g = bgr[..., 1].astype(np.int64); r = bgr[..., 2].astype(np.int64)
    half, off = 1 << 15, 128 << 16
    y = (fix(0.29900) * r + fix(0.58700) * g + fix(0.11400) * b + half) >> 16
    cb = (-fix(0.16874) * r - fix(0.33126) * g + fix(0.50000) * b + off + half - 1) >> 16
    cr = (fix(0.50000) * r - fix(0.41869) * g - fix(0.08131) * b + off + half - 1) >> 16
    return y, cb, cr


def _pad_edge(plane, rows, cols):
    """replicate the last column / row (expand_right_edge, expand_bottom_edge)"""
    h, w = plane.shape
    return np.pad(plane, ((0, rows - h), (0, cols - w)), mode="edge")


def h2v2_downsample(plane, out_rows, out_cols):
    """jcsample.c h2v2_downsample on a plane already padded to (2 out_rows, 2 out_cols): bias 1, 2, 1, 2, ... along a row"""
    p = plane[:2 * out_rows, :2 * out_cols]
    s = p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2]
    bias = np.tile(np.array([1, 2], np.int64), out_cols)[:out_cols]
    return (s + bias[None, :]) >> 2


def fdct_islow(blocks):
    """jfdctint.c jpeg_fdct_islow on int64 [n, 8, 8] level-shifted samples -> [n, 8, 8] (scaled by 8)."""
    C = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137, f1_961=16069,
             f2_053=16819, f2_562=20995, f3_072=25172)

    def descale(x, n):
        return (x + (1 << (n - 1))) >> n

    def one_pass(d, first):
        # d: [n, 8, 8]; transform along the last axis
        t0 = d[..., 0] + d[..., 7]; t7 = d[..., 0] - d[..., 7]
        t1 = d[..., 1] + d[..., 6]; t6 = d[..., 1] - d[..., 6]
        t2 = d[..., 2] + d[..., 5]; t5 = d[..., 2] - d[..., 5]
        t3 = d[..., 3] + d[..., 4]; t4 = d[..., 3] - d[..., 4]
        t10 = t0 + t3; t13 = t0 - t3; t11 = t1 + t2; t12 = t1 - t2
        out = np.empty_like(d)
        if first:
            out[..., 0] = (t10 + t11) << 2
            out[..., 4] = (t10 - t11) << 2
            sh = 13 - 2
        else:
            out[..., 0] = descale(t10 + t11, 2)
            out[..., 4] = descale(t10 - t11, 2)
            sh = 13 + 2
        z1 = (t12 + t13) * C["f0_541"]
        out[..., 2] = descale(z1 + t13 * C["f0_765"], sh)
        out[..., 6] = descale(z1 + t12 * (-C["f1_847"]), sh)
        z1 = t4 + t7; z2 = t5 + t6; z3 = t4 + t6; z4 = t5 + t7
        z5 = (z3 + z4) * C["f1_175"]
        t4 = t4 * C["f0_298"]; t5 = t5 * C["f2_053"]; t6 = t6 * C["f3_072"]; t7 = t7 * C["f1_501"]
        z1 = z1 * (-C["f0_899"]); z2 = z2 * (-C["f2_562"]); z3 = z3 * (-C["f1_961"]); z4 = z4 * (-C["f0_390"])
        z3 = z3 + z5; z4 = z4 + z5
        out[..., 7] = descale(t4 + z1 + z3, sh)
        out[..., 5] = descale(t5 + z2 + z4, sh)
        out[..., 3] = descale(t6 + z2 + z3, sh)
        out[..., 1] = descale(t7 + z1 + z4, sh)
        return out

    rows_done = one_pass(blocks, True)
    return one_pass(rows_done.transpose(0, 2, 1), False).transpose(0, 2, 1)


def quantize(coef, qtable):
    """jcdctmgr.c: divisor = q << 3 (the islow output is scaled by 8); round half away from zero."""
    q = (qtable.reshape(1, 8, 8) << 3)
    a = np.abs(coef) + (q >> 1)
    return np.sign(coef) * (a // q)


def component_blocks(plane, blocks_h, blocks_w, qtable):
    """plane already padded to (8 blocks_h, 8 blocks_w) -> quantised coefficients int64 [blocks_h, blocks_w, 64] in ZIGZAG order."""
    b = plane.reshape(blocks_h, 8, blocks_w, 8).transpose(0, 2, 1, 3).reshape(-1, 8, 8) - 128
    q = quantize(fdct_islow(b), qtable).reshape(blocks_h, blocks_w, 64)
    return q[..., ZIGZAG]


class BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, code, length):
        self.acc = (self.acc << length) | (code & ((1 << length) - 1))
        self.n += length
        while self.n >= 8:
            byte = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(byte)
            if byte == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put(0x7F, 8 - self.n)          # pad with one-bits (jchuff.c flush_bits)


def _nbits(v):
    return int(v).bit_length()


def encode_block(bw, zz, last_dc, dc_tab, ac_tab):
    """jchuff.c encode_one_block; zz = 64 coefficients in zigzag order."""
    t = int(zz[0]) - last_dc
    t2 = t
    if t < 0:
        t = -t
        t2 -= 1
    n = _nbits(t)
    bw.put(*dc_tab[n])
    if n:
        bw.put(t2, n)
    r = 0
    for k in range(1, 64):
        t = int(zz[k])
        if t == 0:
            r += 1
            continue
        while r > 15:
            bw.put(*ac_tab[0xF0])
            r -= 16
        t2 = t
        if t < 0:
            t = -t
            t2 -= 1
        n = _nbits(t)
        bw.put(*ac_tab[(r << 4) + n])
        bw.put(t2, n)
        r = 0
    if r > 0:
        bw.put(*ac_tab[0x00])
    return int(zz[0])


def _segment(marker, payload):
    return bytes([0xFF, marker]) + (len(payload) + 2).to_bytes(2, "big") + payload


def header(rows, cols, channels, qtabs):
    """jcmarker.c: SOI, JFIF APP0 (1.01, aspect 1:1), DQT per table (8-bit, zigzag order), SOF0, DHT per table, SOS"""
    out = bytearray(b"\xFF\xD8")
    out += _segment(0xE0, b"JFIF\x00\x01\x01\x00\x00\x01\x00\x01\x00\x00")
    for i, q in enumerate(qtabs):
        out += _segment(0xDB, bytes([i]) + bytes(int(v) for v in q[ZIGZAG]))
    sof = bytes([8]) + rows.to_bytes(2, "big") + cols.to_bytes(2, "big") + bytes([channels])
    if channels == 1:
        sof += bytes([1, 0x11, 0])
    else:
        sof += bytes([1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1])
    out += _segment(0xC0, sof)
    tabs = [(0x00, DC_LUMA), (0x10, AC_LUMA)] + ([(0x01, DC_CHROMA), (0x11, AC_CHROMA)] if channels == 3 else [])
    for ident, (bits, vals) in tabs:
        out += _segment(0xC4, bytes([ident]) + bytes(bits) + bytes(vals))
    if channels == 1:
        out += _segment(0xDA, bytes([1, 1, 0x00, 0, 63, 0]))
    else:
        out += _segment(0xDA, bytes([3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0]))
    return bytes(out)


def coefficients(img, quality=95):
    """-> list of (coef [block rows, block cols, 64] zigzag int64, h_samp, v_samp, real block rows, real block cols), qtabs.
    Blocks beyond the component's own ceil(size / 8) blocks are DUMMY blocks (jccoefct.c): `encode` applies their rule."""
    img = np.asarray(img)
    rows, cols = img.shape[:2]
    if img.ndim == 2:
        ql = quant_table(STD_LUMA_Q, quality)
        bh, bw = -(-rows // 8), -(-cols // 8)
        y = _pad_edge(img.astype(np.int64), bh * 8, bw * 8)
        return [(component_blocks(y, bh, bw, ql), 1, 1, bh, bw)], [ql]
    ql, qc = quant_table(STD_LUMA_Q, quality), quant_table(STD_CHROMA_Q, quality)
    y, cb, cr = rgb_to_ycc(img)
    mcu_h, mcu_w = -(-rows // 16), -(-cols // 16)
    ybh, ybw = -(-rows // 8), -(-cols // 8)                       # real luma blocks
    crows, ccols = -(-rows // 2), -(-cols // 2)
    cbh, cbw = -(-crows // 8), -(-ccols // 8)                     # real chroma blocks
    comps = []
    # luma: rows replicated to the iMCU height (16 mcu_h), columns to 8 ybw; blocks beyond ybw are dummies
    yp = _pad_edge(y, mcu_h * 16, ybw * 8)
    comps.append((component_blocks(yp, mcu_h * 2, ybw, ql), 2, 2, ybh, ybw))
    for c in (cb, cr):
        # jcprepct: pad the input rows to an even count; h2v2_downsample pads the input columns to 2 * 8 cbw; then the
        # downsampled rows are replicated to the iMCU height (8 mcu_h)
        cp = _pad_edge(c, 2 * crows, 2 * 8 * cbw)
        d = h2v2_downsample(cp, crows, 8 * cbw)
        d = _pad_edge(d, mcu_h * 8, 8 * cbw)
        comps.append((component_blocks(d, mcu_h, cbw, qc), 1, 1, mcu_h, cbw))
    return comps, [ql, qc]


def encode(img, quality=95):
    """-> bytes, equal to cv2.imencode('.jpg', img, [cv2.IMWRITE_JPEG_QUALITY, quality]).tobytes()"""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and (img.ndim == 2 or (img.ndim == 3 and img.shape[2] == 3))
    rows, cols = img.shape[:2]
    channels = 1 if img.ndim == 2 else 3
    comps, qtabs = coefficients(img, quality)
    dcl, acl = huff_codes(*DC_LUMA), huff_codes(*AC_LUMA)
    dcc, acc = huff_codes(*DC_CHROMA), huff_codes(*AC_CHROMA)
    bw = BitWriter()
    zero = np.zeros(64, np.int64)
    if channels == 1:
        coef = comps[0][0]
        last = 0
        for r in range(coef.shape[0]):
            for c in range(coef.shape[1]):
                last = encode_block(bw, coef[r, c], last, dcl, acl)
    else:
        mcu_h, mcu_w = -(-rows // 16), -(-cols // 16)
        last = [0, 0, 0]
        for my in range(mcu_h):
            for mx in range(mcu_w):
                for ci, (coef, hs, vs, nbh, nbw) in enumerate(comps):
                    dct, act = (dcl, acl) if ci == 0 else (dcc, acc)
                    prev_dc = None                                 # DC of MCU_buffer[blkn - 1]
                    for by in range(vs):
                        row_dc = prev_dc                           # a dummy ROW takes the DC of the block before the row
                        for bx in range(hs):
                            r, c = my * vs + by, mx * hs + bx
                            if r < nbh and c < nbw:
                                blk = coef[r, c]
                            else:                                  # jccoefct.c compress_data: zero AC, DC copied
                                blk = zero.copy()
                                blk[0] = prev_dc if r < nbh else row_dc
                            last[ci] = encode_block(bw, blk, last[ci], dct, act)
                            prev_dc = int(blk[0])
    bw.flush()
    return header(rows, cols, channels, qtabs) + bytes(bw.out) + b"\xFF\xD9"
