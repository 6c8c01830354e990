"""CPU restatement of the grayscale JPEG decode the reference performs at Stitcher.py:68-69
(`cv2.imdecode(np.fromfile(f), 0)`): libjpeg(-turbo) with out_color_space = JCS_GRAYSCALE, i.e. ONLY the luma
component is reconstructed -- entropy decode, dequantise, the accurate integer IDCT (`jidctint.c`, JDCT_ISLOW, the
library default) and the post-IDCT range limit.  No colour conversion, no upsampling.

TEST INFRASTRUCTURE ONLY (see oracle/README in DESIGN.md): imported by tests/ and by oracle/check_jpeg_demo.py.
Pinned bit-exactly against cv2.imdecode on the fixtures of tests/golden and on all 140 demo JPEGs of the reference
(oracle/check_jpeg_demo.py, run in the build container).

Baseline / extended-sequential Huffman, 8-bit precision, one interleaved scan (or a single-component image), restart
intervals.  The Huffman stage is a plain Python loop: use it on small images.
"""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                   54, 47, 55, 62, 63], np.int32)      # zigzag position -> natural (row-major) index


class Unsupported(ValueError):
    pass


def parse(data):
    """-> dict(rows, cols, comps=[(id, h, v, tq, td, ta)], quant={tq: int32[64] natural order}, dc / ac = {id: (bits, vals)},
    restart_interval, scan=(start, end) byte range of the entropy-coded segment)."""
    d = bytes(data)
    if d[:2] != b"\xff\xd8":
        raise Unsupported("not a JPEG")
    i = 2
    out = dict(quant={}, dc={}, ac={}, restart_interval=0, comps=None)
    while True:
        if d[i] != 0xFF:
            raise Unsupported("marker expected")
        while d[i + 1] == 0xFF:
            i += 1
        m = d[i + 1]
        i += 2
        if m == 0xD8 or m == 0x01 or 0xD0 <= m <= 0xD7:
            continue
        if m == 0xD9:
            raise Unsupported("EOI before SOS")
        L = (d[i] << 8) | d[i + 1]
        seg = d[i + 2:i + L]
        if m == 0xDB:
            k = 0
            while k < len(seg):
                pq, tq = seg[k] >> 4, seg[k] & 15
                k += 1
                q = np.zeros(64, np.int32)
                for z in range(64):
                    if pq:
                        q[ZIGZAG[z]] = (seg[k] << 8) | seg[k + 1]; k += 2
                    else:
                        q[ZIGZAG[z]] = seg[k]; k += 1
                out["quant"][tq] = q
        elif m == 0xC4:
            k = 0
            while k < len(seg):
                tc, th = seg[k] >> 4, seg[k] & 15
                bits = list(seg[k + 1:k + 17]); n = sum(bits)
                vals = list(seg[k + 17:k + 17 + n]); k += 17 + n
                out["ac" if tc else "dc"][th] = (bits, vals)
        elif m in (0xC0, 0xC1):
            if seg[0] != 8:
                raise Unsupported("precision")
            out["rows"] = (seg[1] << 8) | seg[2]; out["cols"] = (seg[3] << 8) | seg[4]
            nc = seg[5]
            out["comps"] = [[seg[6 + 3 * c], seg[7 + 3 * c] >> 4, seg[7 + 3 * c] & 15, seg[8 + 3 * c], 0, 0] for c in range(nc)]
        elif m in (0xC2, 0xC3, 0xC5, 0xC6, 0xC7, 0xC9, 0xCA, 0xCB, 0xCD, 0xCE, 0xCF):
            raise Unsupported("not baseline / extended sequential Huffman")
        elif m == 0xDD:
            out["restart_interval"] = (seg[0] << 8) | seg[1]
        elif m == 0xDA:
            ns = seg[0]
            if out["comps"] is None or ns != len(out["comps"]):
                raise Unsupported("non-interleaved scans")
            for c in range(ns):
                cid, t = seg[1 + 2 * c], seg[2 + 2 * c]
                comp = [cc for cc in out["comps"] if cc[0] == cid][0]
                comp[4], comp[5] = t >> 4, t & 15
            out["scan"] = (i + L, len(d))
            return out
        i += L


def _huff_table(bits, vals):
    """code -> (length, symbol) dictionary keyed by (length, code)."""
    table = {}
    code = 0; k = 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            table[(length, code)] = vals[k]; k += 1; code += 1
        code <<= 1
    return table


class _Bits:
    def __init__(self, d, pos, end):
        self.d, self.pos, self.end = d, pos, end
        self.acc = 0; self.n = 0

    def _fill(self):
        while self.n <= 24:
            if self.pos >= self.end:
                b = 0
            else:
                b = self.d[self.pos]
                if b == 0xFF:
                    nx = self.d[self.pos + 1] if self.pos + 1 < self.end else 0xD9
                    if nx == 0:
                        self.pos += 2
                    else:
                        b = 0              # a marker: feed zeros, do not advance
                else:
                    self.pos += 1
            self.acc = (self.acc << 8) | b; self.n += 8

    def get(self, k):
        if k == 0:
            return 0
        if self.n < k:
            self._fill()
        v = (self.acc >> (self.n - k)) & ((1 << k) - 1)
        self.n -= k
        self.acc &= (1 << self.n) - 1
        return v

    def decode(self, table):
        code = 0
        for length in range(1, 17):
            code = (code << 1) | self.get(1)
            s = table.get((length, code))
            if s is not None:
                return s
        raise ValueError("bad Huffman code")

    def restart(self):
        self.acc = 0; self.n = 0
        # skip to the RSTn marker and over it
        while not (self.d[self.pos] == 0xFF and 0xD0 <= self.d[self.pos + 1] <= 0xD7):
            self.pos += 1
        self.pos += 2


def _extend(v, s):
    return v if v >= (1 << (s - 1)) else v - (1 << s) + 1


def luma_coefficients(data):
    """Entropy-decode; -> (info, coef int16 [blocks_h, blocks_w, 64] natural order, quantised) for component 0."""
    info = parse(data)
    comps = info["comps"]
    hmax = max(c[1] for c in comps); vmax = max(c[2] for c in comps)
    mcux = -(-info["cols"] // (8 * hmax)); mcuy = -(-info["rows"] // (8 * vmax))
    if len(comps) == 1:                       # a single-component scan is never interleaved: MCU = one block
        comps[0][1] = comps[0][2] = 1
        mcux = -(-info["cols"] // 8); mcuy = -(-info["rows"] // 8)
    y = comps[0]
    bw, bh = mcux * y[1], mcuy * y[2]
    coef = np.zeros((bh, bw, 64), np.int16)
    dc_t = {k: _huff_table(*v) for k, v in info["dc"].items()}
    ac_t = {k: _huff_table(*v) for k, v in info["ac"].items()}
    br = _Bits(bytes(data), info["scan"][0], info["scan"][1])
    pred = [0] * len(comps)
    ri = info["restart_interval"]
    n_mcu = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if ri and n_mcu and n_mcu % ri == 0:
                br.restart(); pred = [0] * len(comps)
            n_mcu += 1
            for ci, c in enumerate(comps):
                for v in range(c[2]):
                    for h in range(c[1]):
                        s = br.decode(dc_t[c[4]])
                        diff = _extend(br.get(s), s) if s else 0
                        pred[ci] += diff
                        blk = coef[my * c[2] + v, mx * c[1] + h] if ci == 0 else None
                        if blk is not None:
                            blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = br.decode(ac_t[c[5]])
                            r, s = rs >> 4, rs & 15
                            if s == 0:
                                if r == 15:
                                    k += 16; continue
                                break
                            k += r
                            val = _extend(br.get(s), s)
                            if blk is not None:
                                blk[ZIGZAG[k]] = val
                            k += 1
    info["blocks"] = (bh, bw)
    return info, coef


# ---- jidctint.c (jpeg_idct_islow): CONST_BITS = 13, PASS1_BITS = 2
_F = dict(f0_298=2446, f0_390=3196, f0_541=4433, f0_765=6270, f0_899=7373, f1_175=9633, f1_501=12299, f1_847=15137,
          f1_961=16069, f2_053=16819, f2_562=20995, f3_072=25172)


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _pass(v, shift):
    """One 1-D pass over the LAST axis of int64 v[..., 8]; result descaled by `shift`."""
    F = _F
    z2, z3 = v[..., 2], v[..., 6]
    z1 = (z2 + z3) * F["f0_541"]
    tmp2 = z1 + z3 * (-F["f1_847"])
    tmp3 = z1 + z2 * F["f0_765"]
    z2, z3 = v[..., 0], v[..., 4]
    tmp0 = (z2 + z3) << 13
    tmp1 = (z2 - z3) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = v[..., 7], v[..., 5], v[..., 3], v[..., 1]
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * F["f1_175"]
    tmp0 = tmp0 * F["f0_298"]; tmp1 = tmp1 * F["f2_053"]; tmp2 = tmp2 * F["f3_072"]; tmp3 = tmp3 * F["f1_501"]
    z1 = z1 * (-F["f0_899"]); z2 = z2 * (-F["f2_562"]); z3 = z3 * (-F["f1_961"]) + z5; z4 = z4 * (-F["f0_390"]) + z5
    tmp0 = tmp0 + z1 + z3; tmp1 = tmp1 + z2 + z4; tmp2 = tmp2 + z2 + z3; tmp3 = tmp3 + z1 + z4
    out = np.stack([tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2, tmp10 - tmp3], axis=-1)
    return _descale(out, shift)


def range_limit(x):
    """libjpeg's post-IDCT table lookup `range_limit[x & RANGE_MASK]` (table centred on +128, RANGE_MASK = 1023)."""
    idx = x & 1023
    return np.where(idx < 128, idx + 128, np.where(idx < 512, 255, np.where(idx < 896, 0, idx - 896))).astype(np.uint8)


def idct_islow(coef, quant):
    """coef int16 [bh, bw, 64] (natural order, quantised), quant int32[64] -> u8 [bh*8, bw*8]."""
    bh, bw, _ = coef.shape
    blk = (coef.astype(np.int64) * quant.astype(np.int64)).reshape(bh, bw, 8, 8)          # [.., row, col]
    ws = _pass(np.swapaxes(blk, -1, -2), 13 - 2)            # pass 1 works on columns: put the row index last
    ws = np.swapaxes(ws, -1, -2)                            # back to [row, col]; ws[row k][col c]
    out = _pass(ws, 13 + 2 + 3)                             # pass 2 on rows
    px = range_limit(out)
    return px.transpose(0, 2, 1, 3).reshape(bh * 8, bw * 8)


def decode_gray(data):
    info, coef = luma_coefficients(data)
    q = info["quant"][info["comps"][0][3]]
    full = idct_islow(coef, q)
    return full[:info["rows"], :info["cols"]]


# ---------------------------------------------------------------- colour decode (cv2.imdecode(data, IMREAD_COLOR), Stitcher.py:382,401)
# libjpeg(-turbo) defaults: islow IDCT per component, do_fancy_upsampling = TRUE (so no merged upsampling), YCbCr -> RGB with
# the 16-bit fixed-point tables of jdcolor.c; cv2 asks for BGR order.
def all_coefficients(data):
    """Entropy-decode every component.  -> (info, [coef int16 [bh_c, bw_c, 64]] per component)."""
    info = parse(data)
    comps = info["comps"]
    hmax = max(c[1] for c in comps); vmax = max(c[2] for c in comps)
    mcux = -(-info["cols"] // (8 * hmax)); mcuy = -(-info["rows"] // (8 * vmax))
    if len(comps) == 1:
        comps[0][1] = comps[0][2] = 1
        hmax = vmax = 1
        mcux = -(-info["cols"] // 8); mcuy = -(-info["rows"] // 8)
    coefs = [np.zeros((mcuy * c[2], mcux * c[1], 64), np.int16) for c in comps]
    dc_t = {k: _huff_table(*v) for k, v in info["dc"].items()}
    ac_t = {k: _huff_table(*v) for k, v in info["ac"].items()}
    br = _Bits(bytes(data), info["scan"][0], info["scan"][1])
    pred = [0] * len(comps)
    ri = info["restart_interval"]
    n_mcu = 0
    for my in range(mcuy):
        for mx in range(mcux):
            if ri and n_mcu and n_mcu % ri == 0:
                br.restart(); pred = [0] * len(comps)
            n_mcu += 1
            for ci, c in enumerate(comps):
                for v in range(c[2]):
                    for h in range(c[1]):
                        s_ = br.decode(dc_t[c[4]])
                        pred[ci] += _extend(br.get(s_), s_) if s_ else 0
                        blk = coefs[ci][my * c[2] + v, mx * c[1] + h]
                        blk[0] = pred[ci]
                        k = 1
                        while k < 64:
                            rs = br.decode(ac_t[c[5]])
                            r, s_ = rs >> 4, rs & 15
                            if s_ == 0:
                                if r == 15:
                                    k += 16; continue
                                break
                            k += r
                            blk[ZIGZAG[k]] = _extend(br.get(s_), s_)
                            k += 1
    info["hmax"], info["vmax"] = hmax, vmax
    return info, coefs


def _h2_fancy(p):
    """jdsample.c h2v1_fancy_upsample on int32 rows [R, n] -> [R, 2n] (triangle filter, 3/4 + 1/4, alternating rounding)."""
    n = p.shape[1]
    out = np.empty((p.shape[0], 2 * n), np.int32)
    if n == 1:
        out[:, 0] = p[:, 0]; out[:, 1] = p[:, 0]
        return out
    left = np.concatenate([p[:, :1], p[:, :-1]], axis=1)
    right = np.concatenate([p[:, 1:], p[:, -1:]], axis=1)
    out[:, 0::2] = (3 * p + left + 1) >> 2
    out[:, 1::2] = (3 * p + right + 2) >> 2
    out[:, 0] = p[:, 0]; out[:, -1] = p[:, -1]
    return out


def _v2_colsums(p):
    """Vertical half of h2v2 / h1v2 fancy upsampling: rows [R, n] -> [2R, n] of 3*near + far (context rows replicated at the edges)."""
    above = np.concatenate([p[:1], p[:-1]], axis=0)
    below = np.concatenate([p[1:], p[-1:]], axis=0)
    out = np.empty((2 * p.shape[0], p.shape[1]), np.int32)
    out[0::2] = 3 * p + above
    out[1::2] = 3 * p + below
    return out


def _h2v2_fancy(p):
    cs = _v2_colsums(p)                       # [2R, n]
    n = cs.shape[1]
    out = np.empty((cs.shape[0], 2 * n), np.int32)
    if n == 1:
        out[:, 0] = (cs[:, 0] * 4 + 8) >> 4; out[:, 1] = (cs[:, 0] * 4 + 7) >> 4
        return out
    left = np.concatenate([cs[:, :1], cs[:, :-1]], axis=1)
    right = np.concatenate([cs[:, 1:], cs[:, -1:]], axis=1)
    out[:, 0::2] = (3 * cs + left + 8) >> 4
    out[:, 1::2] = (3 * cs + right + 7) >> 4
    out[:, 0] = (cs[:, 0] * 4 + 8) >> 4; out[:, -1] = (cs[:, -1] * 4 + 7) >> 4
    return out


def _h1v2_fancy(p):
    """libjpeg-turbo h1v2_fancy_upsample: vertical triangle filter only, bias 1 for the upper and 2 for the lower output row."""
    cs = _v2_colsums(p)
    out = np.empty_like(cs)
    out[0::2] = (cs[0::2] + 1) >> 2
    out[1::2] = (cs[1::2] + 2) >> 2
    return out


def upsample(plane, hf, vf, ds_rows, ds_cols):
    """plane: IDCT output of a component (padded); (hf, vf) = max sampling / component sampling; only the component's
    true downsampled size takes part, like libjpeg (compptr->downsampled_width / _height)."""
    p = plane[:ds_rows, :ds_cols].astype(np.int32)
    if hf == 1 and vf == 1:
        return p
    if hf == 2 and vf == 1 and ds_cols > 2:          # jinit_upsampler: fancy only when downsampled_width > 2
        return _h2_fancy(p)
    if hf == 2 and vf == 2 and ds_cols > 2:
        return _h2v2_fancy(p)
    if hf == 1 and vf == 2:
        return _h1v2_fancy(p)
    return np.repeat(np.repeat(p, vf, axis=0), hf, axis=1)          # other integral factors: box replication (int_upsample)


def ycc_to_bgr(y, cb, cr):
    """jdcolor.c ycc_rgb_convert (SCALEBITS = 16)."""
    def fix(x):
        return int(x * 65536 + 0.5)
    x = np.arange(256, dtype=np.int64) - 128
    cr_r = (fix(1.40200) * x + 32768) >> 16
    cb_b = (fix(1.77200) * x + 32768) >> 16
    cr_g = -fix(0.71414) * x
    cb_g = -fix(0.34414) * x + 32768
    y = y.astype(np.int64)
    r = y + cr_r[cr]
    g = y + ((cb_g[cb] + cr_g[cr]) >> 16)
    b = y + cb_b[cb]
    return np.stack([np.clip(b, 0, 255), np.clip(g, 0, 255), np.clip(r, 0, 255)], axis=-1).astype(np.uint8)


def decode_bgr(data):
    info, coefs = all_coefficients(data)
    rows, cols = info["rows"], info["cols"]
    comps = info["comps"]
    planes = [idct_islow(cf, info["quant"][c[3]]) for cf, c in zip(coefs, comps)]
    if len(comps) == 1:
        g = planes[0][:rows, :cols]
        return np.stack([g, g, g], axis=-1)
    up = []
    for pl, c in zip(planes, comps):
        hf, vf = info["hmax"] // c[1], info["vmax"] // c[2]
        ds_rows = -(-rows * c[2] // info["vmax"]); ds_cols = -(-cols * c[1] // info["hmax"])
        up.append(upsample(pl, hf, vf, ds_rows, ds_cols)[:rows, :cols])
    return ycc_to_bgr(up[0], up[1], up[2])
