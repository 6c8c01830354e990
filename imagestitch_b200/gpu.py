"""Array-level host API over libvfsms.so (numpy in / numpy out, plus torch-tensor device variants).

This is the layer the reference-surface modules (ImageUtility.Method, Stitcher, ImageFusion, myGpuFeatures) call.
Nothing here computes on the CPU: every function ends in a C-ABI call and raises VfsmsError when the library or
the device is missing.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import KP_STRIDE, PAIR_RESULT_DTYPE, PairResult, SurfParams, VfsmsError, check

# reference defaults: ImageUtility.py:23-28 (GPU-SURF) and cv2.xfeatures2d.SURF_create() (ImageUtility.py:258)
SURF_GPU_DEFAULTS = dict(hessian_threshold=100.0, n_octaves=4, n_octave_layers=3, extended=True, keypoints_ratio=0.01,
                         upright=False)
SURF_CPU_DEFAULTS = dict(hessian_threshold=100.0, n_octaves=4, n_octave_layers=3, extended=False, keypoints_ratio=0.0,
                         upright=False)


def surf_params(hessian_threshold=100.0, n_octaves=4, n_octave_layers=3, extended=True, keypoints_ratio=0.01,
                upright=False):
    return SurfParams(float(hessian_threshold), int(n_octaves), int(n_octave_layers), int(bool(extended)),
                      float(keypoints_ratio), int(bool(upright)))


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data)


def _as_u8_image(image):
    img = np.asarray(image)
    if img.dtype != np.uint8 or img.ndim != 2:
        raise TypeError("expected a 2-D uint8 image, got %s %s" % (img.dtype, img.shape))
    if img.strides[1] != 1 or img.strides[0] < img.shape[1]:
        img = np.ascontiguousarray(img)     # Fortran-ordered / negative-stride views (conversion.cpp:197-202)
    return img


def surf_detect_and_describe(image, params=None, device=0, **kw):
    """-> (kp [N, 8] float32 (x, y, size, angle, response, octave, laplacian, 0), desc [N, 64|128] float32)."""
    L = _lib.load()
    ctx = _lib.context(device)
    img = _as_u8_image(image)
    p = params if params is not None else surf_params(**kw)
    h, w = img.shape
    dim = 128 if p.extended else 64
    if p.keypoints_ratio > 0:
        cap = int(min(max(p.keypoints_ratio * h * w, 1), 65535)) + 64
    else:
        cap = max(4096, h * w // 24) + 64
    n = ctypes.c_int(0)
    while True:
        kp = np.empty((cap, KP_STRIDE), np.float32)
        desc = np.empty((cap, dim), np.float32)
        rc = L.vfsms_surf_detect_and_describe(ctx, _vp(img), h, w, img.strides[0], ctypes.byref(p), _vp(kp), _vp(desc), cap,
                                              ctypes.byref(n))
        if rc == _lib.VFSMS_E_CAPACITY:
            cap = n.value + 64
            continue
        check(rc, "vfsms_surf_detect_and_describe")
        return kp[:n.value].copy(), desc[:n.value].copy()


def match_descriptors(desc_a, desc_b, feature_type=2, param=0.75, device=0):
    """-> int32 [M, 2] rows (trainIdx, queryIdx), ascending queryIdx (appendix/myGpuFeatures.cpp:53-65)."""
    L = _lib.load()
    ctx = _lib.context(device)
    A = np.ascontiguousarray(desc_a, np.float32)
    B = np.ascontiguousarray(desc_b, np.float32)
    if A.ndim != 2 or B.ndim != 2 or (A.shape[0] and B.shape[0] and A.shape[1] != B.shape[1]):
        raise ValueError("descriptor arrays must be [n, D] with equal D")
    nA, nB = A.shape[0], B.shape[0]
    out = np.empty((max(nA, 1), 2), np.int32)
    m = ctypes.c_int(0)
    dim = A.shape[1] if nA else (B.shape[1] if nB else 1)
    check(L.vfsms_match_descriptors(ctx, _vp(A), nA, _vp(B), nB, dim, int(feature_type), float(param), _vp(out),
                                    ctypes.byref(m)), "vfsms_match_descriptors")
    return out[:m.value].copy()


def offset_by_mode(kps_a, kps_b, matches, offset_evaluate=3, device=0):
    """-> (status, [dRow, dCol], votes) with the reference's truncation / first-seen tie rule."""
    L = _lib.load()
    ctx = _lib.context(device)
    A = np.ascontiguousarray(kps_a, np.float32).reshape(len(kps_a), -1) if len(kps_a) else np.zeros((0, 2), np.float32)
    B = np.ascontiguousarray(kps_b, np.float32).reshape(len(kps_b), -1) if len(kps_b) else np.zeros((0, 2), np.float32)
    m = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
    if A.shape[1] != B.shape[1]:
        raise ValueError("keypoint arrays must share their row length")
    if len(m) and (m[:, 1].max() >= len(A) or m[:, 0].max() >= len(B) or m.min() < 0):
        raise IndexError("match index out of range")
    r = PairResult()
    check(L.vfsms_offset_by_mode(ctx, _vp(A), len(A), _vp(B), len(B), A.shape[1], _vp(m), len(m), int(offset_evaluate),
                                 ctypes.byref(r)), "vfsms_offset_by_mode")
    return bool(r.status), [int(r.d_row), int(r.d_col)], int(r.votes)


def align_batch(rois_a, rois_b, params=None, ratio=0.75, offset_evaluate=3, device=0, **kw):
    """Fused detect x2 -> match -> vote for P ROI pairs of equal shape (host arrays).  -> structured array [P]."""
    L = _lib.load()
    ctx = _lib.context(device)
    A = np.asarray(rois_a)
    B = np.asarray(rois_b)
    if A.ndim == 2:
        A = A[None]; B = B[None]
    if A.dtype != np.uint8 or B.dtype != np.uint8 or A.shape != B.shape or A.ndim != 3:
        raise TypeError("rois must be uint8 arrays of identical shape [P, h, w]")
    if A.strides[2] != 1 or A.strides[1] < A.shape[2] or A.strides[0] < 0:
        A = np.ascontiguousarray(A)
    if B.strides != A.strides:
        A = np.ascontiguousarray(A); B = np.ascontiguousarray(B)
    P, h, w = A.shape
    p = params if params is not None else surf_params(**kw)
    res = np.zeros(P, PAIR_RESULT_DTYPE)
    check(L.vfsms_align_batch_host(ctx, _vp(A), _vp(B), P, h, w, A.strides[1], A.strides[0] if P > 1 else h * A.strides[1],
                                   ctypes.byref(p), float(ratio), int(offset_evaluate), _vp(res)), "vfsms_align_batch_host")
    return res


def _roi_stack(rois_a, rois_b):
    A = np.asarray(rois_a)
    B = np.asarray(rois_b)
    if A.ndim == 2:
        A = A[None]; B = B[None]
    if A.dtype != np.uint8 or B.dtype != np.uint8 or A.shape != B.shape or A.ndim != 3:
        raise TypeError("rois must be uint8 arrays of identical shape [P, h, w]")
    if A.strides[2] != 1 or A.strides[1] < A.shape[2] or A.strides[0] < 0:
        A = np.ascontiguousarray(A)
    if B.strides != A.strides:
        A = np.ascontiguousarray(A); B = np.ascontiguousarray(B)
    return A, B


def align_batches(batches, params=None, ratio=0.75, offset_evaluate=3, device=0, **kw):
    """align_batch over a stream of batches, double-buffered: `batches` yields (rois_a, rois_b) host arrays [P, h, w]; the
    generator yields one structured result array per batch, in order.  The host -> device copy of batch k+1 is enqueued
    (vfsms_align_batch_upload, asynchronous when the arrays are in pinned memory) before batch k runs, so it overlaps
    batch k's kernels; results are the same as calling align_batch on every batch."""
    L = _lib.load()
    ctx = _lib.context(device)
    p = params if params is not None else surf_params(**kw)

    def upload(slot, ab):
        A, B = _roi_stack(*ab)
        P, h, w = A.shape
        check(L.vfsms_align_batch_upload(ctx, slot, _vp(A), _vp(B), P, h, w, A.strides[1], A.strides[0] if P > 1 else h * A.strides[1]),
              "vfsms_align_batch_upload")
        return slot, P, A, B                    # the arrays stay referenced until their batch has run

    def run(job):
        slot, P = job[0], job[1]
        res = np.zeros(P, PAIR_RESULT_DTYPE)
        check(L.vfsms_align_batch_run(ctx, slot, ctypes.byref(p), float(ratio), int(offset_evaluate), _vp(res)), "vfsms_align_batch_run")
        return res

    pending = None
    for k, ab in enumerate(batches):
        job = upload(k & 1, ab)
        if pending is not None:
            yield run(pending)
        pending = job
    if pending is not None:
        yield run(pending)


def align_batch_dev(rois_a, rois_b, results, params=None, ratio=0.75, offset_evaluate=3, stream=None):
    """Device-resident variant: rois_* are contiguous torch uint8 CUDA tensors [P, h, w]; results is a torch int32
    CUDA tensor [P, 8].  Asynchronous on `stream` (a torch.cuda.Stream) or the current torch stream."""
    import torch
    L = _lib.load()
    dev = rois_a.device.index or 0
    ctx = _lib.context(dev)
    assert rois_a.is_cuda and rois_b.is_cuda and results.is_cuda and rois_a.dtype == torch.uint8
    assert rois_a.is_contiguous() and rois_b.is_contiguous() and results.is_contiguous() and rois_a.shape == rois_b.shape
    P, h, w = rois_a.shape
    assert results.dtype == torch.int32 and results.numel() >= 8 * P
    p = params if params is not None else surf_params()
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    check(L.vfsms_align_batch_dev(ctx, ctypes.c_void_p(rois_a.data_ptr()), ctypes.c_void_p(rois_b.data_ptr()), P, h, w, w, h * w,
                                  ctypes.byref(p), float(ratio), int(offset_evaluate), ctypes.c_void_p(results.data_ptr()),
                                  ctypes.c_void_p(st.cuda_stream)), "vfsms_align_batch_dev")


def align_strips_dev(tiles_a, tiles_b, results, roi_rows, params=None, ratio=0.75, offset_evaluate=3, stream=None):
    """Direction-1 incremental-search body on device-resident FULL tiles [P, H, W] (torch uint8 CUDA): ROI A = bottom
    `roi_rows` rows of tiles_a[p], ROI B = top rows of tiles_b[p] (ImageUtility.py:77-82), read in place by stride."""
    import torch
    L = _lib.load()
    dev = tiles_a.device.index or 0
    ctx = _lib.context(dev)
    assert tiles_a.is_cuda and tiles_a.dtype == torch.uint8 and tiles_a.is_contiguous() and tiles_b.is_contiguous()
    P, H, W = tiles_a.shape
    assert results.dtype == torch.int32 and results.numel() >= 8 * P and results.is_contiguous()
    p = params if params is not None else surf_params()
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    a_ptr = tiles_a.data_ptr() + (H - roi_rows) * W
    check(L.vfsms_align_batch_dev(ctx, ctypes.c_void_p(a_ptr), ctypes.c_void_p(tiles_b.data_ptr()), P, roi_rows, W, W, H * W,
                                  ctypes.byref(p), float(ratio), int(offset_evaluate), ctypes.c_void_p(results.data_ptr()),
                                  ctypes.c_void_p(st.cuda_stream)), "vfsms_align_batch_dev")


FUSE_METHODS = {"notFuse": 0, "average": 1, "maximum": 2, "minimum": 3, "fadeInAndFadeOut": 4, "trigonometric": 5,
                "multiBandBlending": 6}


DEVICE_MOSAIC_METHODS = ("notFuse", "average", "maximum", "minimum", "fadeInAndFadeOut", "trigonometric")


def _as_i16(a):
    a = np.asarray(a)
    if a.size and (a.min() < -1 or a.max() > 255):
        raise ValueError("canvas values must lie in [-1, 255]")
    return np.ascontiguousarray(a, np.int16)


def fuse_roi(image_a, image_b, method, d_row=0, d_col=0, want_weights=False, force_corner=False, device=0, raw=False):
    """Stitcher.fuseImage for one overlap ROI.  image_*: [r, c] or [r, c, 3] integer arrays with -1 = empty.
    -> uint8 array (and the float32 weight matrices when want_weights)."""
    L = _lib.load()
    ctx = _lib.context(device)
    A, B = _as_i16(image_a), _as_i16(image_b)
    if A.shape != B.shape or A.ndim not in (2, 3):
        raise ValueError("ROI arrays must have identical 2-D or 3-D shape")
    rows, cols = A.shape[:2]
    ch = 1 if A.ndim == 2 else A.shape[2]
    m = FUSE_METHODS[method] if isinstance(method, str) else int(method)
    if force_corner:
        m |= 0x100
    if raw:                 # average / maximum / minimum on the arrays as given (no -1 -> 0, no mutual zero fill)
        m |= 0x200
    out = np.empty(A.shape, np.uint8)
    wa = wb = None
    if want_weights:
        wa = np.empty((rows, cols), np.float32); wb = np.empty((rows, cols), np.float32)
    check(L.vfsms_fuse_roi_host(ctx, _vp(A), _vp(B), rows, cols, ch, m, int(d_row), int(d_col), _vp(out),
                                _vp(wa) if want_weights else None, _vp(wb) if want_weights else None), "vfsms_fuse_roi_host")
    return (out, wa, wb) if want_weights else out


def mosaic(tiles, tile_origin, roi_rect, pair_offset, method, canvas_shape, device=0):
    """Paste/blend loop of Stitcher.getStitchByOffset on a device-resident canvas.
    tiles [n, h, w] or [n, h, w, 3] uint8; tile_origin [n, 2]; roi_rect [n, 4] (r0, c0, r1, c1 in canvas coordinates);
    pair_offset [n, 2] original offsets; canvas_shape (rows, cols)."""
    L = _lib.load()
    ctx = _lib.context(device)
    T = np.ascontiguousarray(tiles, np.uint8)
    n, h, w = T.shape[:3]
    ch = 1 if T.ndim == 3 else T.shape[3]
    org = np.ascontiguousarray(tile_origin, np.int32).reshape(n, 2)
    roi = np.ascontiguousarray(roi_rect, np.int32).reshape(n, 4)
    off = np.ascontiguousarray(pair_offset, np.int32).reshape(n, 2)
    m = FUSE_METHODS[method] if isinstance(method, str) else int(method)
    R, C = int(canvas_shape[0]), int(canvas_shape[1])
    out = np.empty((R, C) if ch == 1 else (R, C, ch), np.uint8)
    check(L.vfsms_mosaic_host(ctx, _vp(T), n, h, w, ch, _vp(org), _vp(roi), _vp(off), m, R, C, _vp(out)), "vfsms_mosaic_host")
    return out


def mosaic_band(tiles, tile_origin, roi_rect, pair_offset, method, canvas_shape, fuse_first=False, halo_in=None,
                halo_in_rect=None, halo_out_rect=None, device=0):
    """One band of a mosaic partitioned over GPUs (vfsms_mosaic_band_host; host logic in sharding.mosaic_sharded).
    Band-local coordinates.  halo_in: int16 [rows, cols(, 3)] with -1 = empty, pasted at halo_in_rect = (r0, c0, rows, cols)
    before the first tile; halo_out_rect: rectangle read back as int16 after the last tile.
    -> (canvas uint8, halo_out int16 or None)."""
    L = _lib.load()
    ctx = _lib.context(device)
    T = np.ascontiguousarray(tiles, np.uint8)
    n, h, w = T.shape[:3]
    ch = 1 if T.ndim == 3 else T.shape[3]
    org = np.ascontiguousarray(tile_origin, np.int32).reshape(n, 2)
    roi = np.ascontiguousarray(roi_rect, np.int32).reshape(n, 4)
    off = np.ascontiguousarray(pair_offset, np.int32).reshape(n, 2)
    m = FUSE_METHODS[method] if isinstance(method, str) else int(method)
    R, C = int(canvas_shape[0]), int(canvas_shape[1])
    out = np.empty((R, C) if ch == 1 else (R, C, ch), np.uint8)
    hin = hin_rect = hout = hout_rect = None
    if halo_in is not None:
        hin_rect = np.ascontiguousarray(halo_in_rect, np.int32).reshape(4)
        hin = _as_i16(halo_in)
        if hin.shape[:2] != (hin_rect[2], hin_rect[3]) or (hin.ndim == 3) != (ch == 3):
            raise ValueError("halo_in does not match halo_in_rect / the tile channels")
    if halo_out_rect is not None:
        hout_rect = np.ascontiguousarray(halo_out_rect, np.int32).reshape(4)
        hout = np.empty((hout_rect[2], hout_rect[3]) if ch == 1 else (hout_rect[2], hout_rect[3], ch), np.int16)
    check(L.vfsms_mosaic_band_host(ctx, _vp(T), n, h, w, ch, _vp(org), _vp(roi), _vp(off), m, R, C, int(bool(fuse_first)),
                                   _vp(hin) if hin is not None else None, _vp(hin_rect) if hin is not None else None,
                                   _vp(hout) if hout is not None else None, _vp(hout_rect) if hout is not None else None,
                                   _vp(out)), "vfsms_mosaic_band_host")
    return out, hout


def phase_correlate(roi_a, roi_b, device=0):
    """cv2.phaseCorrelate(np.float64(a), np.float64(b)) as called at Stitcher.py:230 -> ((shift_x, shift_y), response)."""
    L = _lib.load()
    ctx = _lib.context(device)
    A = _as_u8_image(roi_a)
    B = _as_u8_image(roi_b)
    if A.shape != B.shape:
        raise ValueError("phase correlation needs equally sized ROIs")
    if A.strides != B.strides:
        A = np.ascontiguousarray(A); B = np.ascontiguousarray(B)
    out = (ctypes.c_double * 3)()
    check(L.vfsms_phase_correlate_host(ctx, _vp(A), _vp(B), A.shape[0], A.shape[1], A.strides[0], out), "vfsms_phase_correlate_host")
    return (float(out[0]), float(out[1])), float(out[2])


def phase_correlate_dev(roi_a, roi_b, out, stream=None):
    """phase_correlate on device-resident ROIs (torch uint8 CUDA tensors [rows, cols], equal row stride); out: float64 CUDA tensor
    of 3 elements receiving (shift_x, shift_y, response).  Asynchronous on `stream` (default: torch's current stream)."""
    import torch
    L = _lib.load()
    dev = roi_a.device.index or 0
    ctx = _lib.context(dev)
    assert roi_a.is_cuda and roi_b.is_cuda and roi_a.dtype == torch.uint8 and roi_a.shape == roi_b.shape and roi_a.dim() == 2
    assert roi_a.stride(1) == 1 and roi_b.stride(1) == 1 and roi_a.stride(0) == roi_b.stride(0)
    assert out.is_cuda and out.dtype == torch.float64 and out.numel() >= 3
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    check(L.vfsms_phase_correlate_dev(ctx, ctypes.c_void_p(roi_a.data_ptr()), ctypes.c_void_p(roi_b.data_ptr()), roi_a.shape[0],
                                      roi_a.shape[1], roi_a.stride(0), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(st.cuda_stream)),
          "vfsms_phase_correlate_dev")


def overlap_sums(roi_a, roi_b, shifts, device=0):
    """Integer sums (n, Sa, Sb, Sab, Saa, Sbb) over the pixels two ROIs share under each candidate shift (dRow, dCol) --
    the scoring step of the wrap-aware phase mode (phase_wrap.resolve).  -> int64 [n, 6]."""
    L = _lib.load()
    ctx = _lib.context(device)
    a, b = _as_u8_image(roi_a), _as_u8_image(roi_b)
    if a.shape != b.shape:
        raise ValueError("ROIs must have the same shape")
    sh = np.ascontiguousarray(shifts, np.int32).reshape(-1, 2)
    out = np.zeros((len(sh), 6), np.int64)
    check(L.vfsms_overlap_sums_host(ctx, _vp(a), _vp(b), a.shape[0], a.shape[1], a.strides[0], b.strides[0], len(sh), _vp(sh), _vp(out)),
          "vfsms_overlap_sums_host")
    return out


def orb_detect_and_describe(image, n_features=5000, scale_factor=1.2, n_levels=8, edge_threshold=31, first_level=0, wta_k=2,
                            patch_size=31, fast_threshold=20, device=0):
    """-> (kp [N, 8] float32, desc [N, 32] float32 holding byte values) -- appendix/myGpuFeatures.cpp:106-146."""
    L = _lib.load()
    ctx = _lib.context(device)
    img = _as_u8_image(image)
    h, w = img.shape
    cap = int(n_features) + 64
    kp = np.empty((cap, KP_STRIDE), np.float32)
    desc = np.empty((cap, 32), np.float32)
    n = ctypes.c_int(0)
    check(L.vfsms_orb_detect_and_describe(ctx, _vp(img), h, w, img.strides[0], int(n_features), float(scale_factor), int(n_levels),
                                          int(edge_threshold), int(first_level), int(wta_k), int(patch_size), int(fast_threshold),
                                          _vp(kp), _vp(desc), cap, ctypes.byref(n)), "vfsms_orb_detect_and_describe")
    return kp[:n.value].copy(), desc[:n.value].copy()


def enhance(image, clahe=False, clip_limit=20, tile_size=5, device=0):
    """cv2.equalizeHist / cv2.createCLAHE(clip_limit, (tile_size, tile_size)).apply on the device (Stitcher.py:269-276)."""
    L = _lib.load()
    ctx = _lib.context(device)
    img = _as_u8_image(image)
    out = np.empty(img.shape, np.uint8)
    check(L.vfsms_enhance_host(ctx, _vp(img), img.shape[0], img.shape[1], img.strides[0], 1 if clahe else 0, float(clip_limit),
                               int(tile_size), _vp(out)), "vfsms_enhance_host")
    return out


MATCHERS = {"tc": 0, "simt": 1, "tc_1sm": 2, "tc_bf16x3": 3, "tc_f16x2": 4, "tc_f16x1": 5}      # include/vfsms.h vfsms_set_matcher


def set_matcher(mode, device=0):
    """'tc' (default): tcgen05 candidates (CTA pairs, fp16 operands = 'tc_f16x1') + exact rescoring;
    'tc_1sm': the same on single CTAs; 'simt': exact fp32 SIMT kernel; 'tc_bf16x3' / 'tc_f16x2' / 'tc_f16x1': 'tc' with
    split-bf16 (3 terms) / fp16 + query split (2 terms) / plain fp16 (1 term) operands.  Identical results."""
    check(_lib.load().vfsms_set_matcher(_lib.context(device), MATCHERS[mode]), "vfsms_set_matcher")


OPTIONS = {"describe": 0, "sort": 1, "lpt": 2, "entropy": 3}      # include/vfsms.h VFSMS_OPT_*


def set_option(name, value, device=0):
    """Kernel-variant switch (include/vfsms.h VFSMS_OPT_*): every value of an option gives identical results (except the two
    tolerance modes describe=2 / 3, whose stated tolerances include/vfsms.h lists); non-default values are alternative
    schedules kept for A/B measurement.  Also settable through VFSMS_OPTS="describe=0,lpt=1"."""
    check(_lib.load().vfsms_set_option(_lib.context(device), OPTIONS[name], int(value)), "vfsms_set_option")


def get_option(name, device=0):
    v = ctypes.c_int(0)
    check(_lib.load().vfsms_get_option(_lib.context(device), OPTIONS[name], ctypes.byref(v)), "vfsms_get_option")
    return v.value


def last_match_fallbacks(device=0):
    n = ctypes.c_int(0)
    check(_lib.load().vfsms_last_match_fallbacks(_lib.context(device), ctypes.byref(n)), "vfsms_last_match_fallbacks")
    return n.value


def last_match_bound_violations(device=0):
    """Rescored candidates of the last tensor-core match whose GEMM score broke the guard's error bound (expected: 0)."""
    n = ctypes.c_int(0)
    check(_lib.load().vfsms_last_match_bound_violations(_lib.context(device), ctypes.byref(n)), "vfsms_last_match_bound_violations")
    return n.value


def last_describe_handovers(device=0):
    """Keypoints of the last SURF run that the fixed-point window sampler handed to the reference sampler."""
    n = ctypes.c_int(0)
    check(_lib.load().vfsms_last_describe_handovers(_lib.context(device), ctypes.byref(n)), "vfsms_last_describe_handovers")
    return n.value


def profile_enable(on=True, device=0):
    check(_lib.load().vfsms_profile_enable(_lib.context(device), int(on)), "vfsms_profile_enable")


def profile_read(reset=True, device=0):
    """-> {stage_name: (total_ms, calls)} measured with CUDA events on the launching stream."""
    L = _lib.load()
    ms = np.zeros(_lib.STAGE_COUNT, np.float32)
    calls = np.zeros(_lib.STAGE_COUNT, np.int32)
    check(L.vfsms_profile_read(_lib.context(device), _vp(ms), _vp(calls), int(reset)), "vfsms_profile_read")
    return {L.vfsms_stage_name(i).decode(): (float(ms[i]), int(calls[i])) for i in range(_lib.STAGE_COUNT)}


def launch_count(device=0):
    return int(_lib.load().vfsms_launch_count(_lib.context(device)))


def synchronize(device=0):
    check(_lib.load().vfsms_synchronize(_lib.context(device)), "vfsms_synchronize")


def device_count():
    return int(_lib.load().vfsms_device_count())


__all__ = ["surf_params", "surf_detect_and_describe", "match_descriptors", "offset_by_mode", "align_batch",
           "align_batch_dev", "launch_count", "synchronize", "device_count", "VfsmsError"]


# ---------------------------------------------------------------- JPEG tile decode
class JpegUnsupported(_lib.VfsmsError):
    """The file is not a sequential-Huffman 8-bit grayscale / YCbCr JPEG; callers fall back to their own decoder."""


def _jpeg_check(rc, what):
    if rc == -6:
        raise JpegUnsupported("%s: %s" % (what, _lib.load().vfsms_last_error().decode()))
    check(rc, what)


def jpeg_info(data):
    """(rows, cols, components) of a JPEG byte string; host-only."""
    buf = np.frombuffer(data, np.uint8)
    r, c, n = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _jpeg_check(_lib.load().vfsms_jpeg_info(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.size), ctypes.byref(r), ctypes.byref(c),
                                           ctypes.byref(n)), "vfsms_jpeg_info")
    return r.value, c.value, n.value


def jpeg_luma_coefficients(data):
    """Host-only entropy stage: (coef int16 [blocks_h, blocks_w, 64] natural order, quant uint16[64])."""
    L = _lib.load()
    buf = np.frombuffer(data, np.uint8)
    bh, bw = ctypes.c_int(), ctypes.c_int()
    quant = np.zeros(64, np.uint16)
    p = buf.ctypes.data_as(ctypes.c_void_p)
    _jpeg_check(L.vfsms_jpeg_luma_coefficients(p, ctypes.c_size_t(buf.size), None, ctypes.c_size_t(0), ctypes.byref(bh), ctypes.byref(bw),
                                               quant.ctypes.data_as(ctypes.c_void_p)), "vfsms_jpeg_luma_coefficients")
    coef = np.empty((bh.value, bw.value, 64), np.int16)
    _jpeg_check(L.vfsms_jpeg_luma_coefficients(p, ctypes.c_size_t(buf.size), coef.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(coef.size),
                                               ctypes.byref(bh), ctypes.byref(bw), quant.ctypes.data_as(ctypes.c_void_p)),
                "vfsms_jpeg_luma_coefficients")
    return coef, quant


def jpeg_component_coefficients(data, component):
    """Host-only entropy stage for any component: (coef int16 [bh, bw, 64], quant uint16[64], (h_samp, v_samp))."""
    L = _lib.load()
    buf = np.frombuffer(data, np.uint8)
    bh, bw, hs, vs = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    quant = np.zeros(64, np.uint16)
    p = buf.ctypes.data_as(ctypes.c_void_p)
    qp = quant.ctypes.data_as(ctypes.c_void_p)
    _jpeg_check(L.vfsms_jpeg_component_coefficients(p, ctypes.c_size_t(buf.size), int(component), None, ctypes.c_size_t(0), ctypes.byref(bh),
                                                    ctypes.byref(bw), qp, ctypes.byref(hs), ctypes.byref(vs)), "vfsms_jpeg_component_coefficients")
    coef = np.empty((bh.value, bw.value, 64), np.int16)
    _jpeg_check(L.vfsms_jpeg_component_coefficients(p, ctypes.c_size_t(buf.size), int(component), coef.ctypes.data_as(ctypes.c_void_p),
                                                    ctypes.c_size_t(coef.size), ctypes.byref(bh), ctypes.byref(bw), qp, ctypes.byref(hs),
                                                    ctypes.byref(vs)), "vfsms_jpeg_component_coefficients")
    return coef, quant, (hs.value, vs.value)


def _jpeg_args(datas):
    bufs = [np.frombuffer(d, np.uint8) for d in datas]
    n = len(bufs)
    ptrs = (ctypes.c_void_p * n)(*[b.ctypes.data for b in bufs])
    sizes = (ctypes.c_size_t * n)(*[b.size for b in bufs])
    return bufs, ptrs, sizes


def jpeg_decode_gray(datas, device=0):
    """cv2.imdecode(data, cv2.IMREAD_GRAYSCALE) for a list of JPEG byte strings of identical geometry -> u8 [n, rows, cols]
    (bit-identical to cv2).  Entropy decoding on the host cores, IDCT on the device."""
    single = isinstance(datas, (bytes, bytearray, memoryview, np.ndarray))
    if single:
        datas = [datas]
    rows, cols, _ = jpeg_info(datas[0])
    bufs, ptrs, sizes = _jpeg_args(datas)
    out = np.empty((len(bufs), rows, cols), np.uint8)
    _jpeg_check(_lib.load().vfsms_jpeg_decode_gray_host(_lib.context(device), len(bufs), ptrs, sizes, out.ctypes.data_as(ctypes.c_void_p), rows, cols),
                "vfsms_jpeg_decode_gray_host")
    return out[0] if single else out


def jpeg_decode_bgr(datas, device=0):
    """cv2.imdecode(data, cv2.IMREAD_COLOR) for JPEG byte strings of identical geometry -> u8 [n, rows, cols, 3] (BGR,
    bit-identical to cv2: islow IDCT, fancy chroma upsampling, libjpeg's YCbCr tables)."""
    single = isinstance(datas, (bytes, bytearray, memoryview, np.ndarray))
    if single:
        datas = [datas]
    rows, cols, _ = jpeg_info(datas[0])
    bufs, ptrs, sizes = _jpeg_args(datas)
    out = np.empty((len(bufs), rows, cols, 3), np.uint8)
    _jpeg_check(_lib.load().vfsms_jpeg_decode_bgr_host(_lib.context(device), len(bufs), ptrs, sizes, out.ctypes.data_as(ctypes.c_void_p), rows, cols),
                "vfsms_jpeg_decode_bgr_host")
    return out[0] if single else out


def jpeg_last_entropy_passes(device=0):
    """Synchronisation passes of the last device entropy decode (option "entropy" = 1)."""
    n = ctypes.c_int(0)
    check(_lib.load().vfsms_jpeg_last_entropy_passes(_lib.context(device), ctypes.byref(n)), "vfsms_jpeg_last_entropy_passes")
    return n.value


def jpeg_encode(image, quality=95, device=0):
    """cv2.imencode(".jpg", image, [cv2.IMWRITE_JPEG_QUALITY, quality]) on the device -> bytes, identical to cv2's (baseline
    JPEG, Annex-K Huffman tables, 4:2:0 for BGR).  image: u8 [rows, cols] or [rows, cols, 3] (BGR), rows / cols <= 65535."""
    img = np.asarray(image)
    if img.dtype != np.uint8 or not (img.ndim == 2 or (img.ndim == 3 and img.shape[2] == 3)):
        raise TypeError("expected a uint8 image [rows, cols] or [rows, cols, 3], got %s %s" % (img.dtype, img.shape))
    channels = 1 if img.ndim == 2 else 3
    if img.strides[-1] != 1 or (channels == 3 and img.strides[1] != 3) or img.strides[0] < img.shape[1] * channels:
        img = np.ascontiguousarray(img)
    rows, cols = img.shape[:2]
    L, ctx = _lib.load(), _lib.context(device)
    cap = rows * cols * channels // 2 + (64 << 10)
    size = ctypes.c_size_t(0)
    while True:
        out = np.empty(cap, np.uint8)
        rc = L.vfsms_jpeg_encode_host(ctx, _vp(img), rows, cols, channels, ctypes.c_int64(img.strides[0]), int(quality), _vp(out), cap,
                                      ctypes.byref(size))
        if rc == _lib.VFSMS_E_CAPACITY and size.value > cap:
            cap = size.value
            continue
        check(rc, "vfsms_jpeg_encode_host")
        return out[:size.value].tobytes()


def jpeg_encode_dev(image, quality=95, device=0, stream=None):
    """jpeg_encode of a torch.uint8 CUDA tensor [rows, cols] or [rows, cols, 3] (BGR) that stays in HBM; returns the file's bytes."""
    assert image.is_cuda and image.dim() in (2, 3) and image.stride(-1) == 1
    channels = 1 if image.dim() == 2 else 3
    assert channels == 1 or (image.shape[2] == 3 and image.stride(1) == 3)
    rows, cols = int(image.shape[0]), int(image.shape[1])
    L, ctx = _lib.load(), _lib.context(device)
    cap = rows * cols * channels // 2 + (64 << 10)
    size = ctypes.c_size_t(0)
    while True:
        out = np.empty(cap, np.uint8)
        rc = L.vfsms_jpeg_encode_dev(ctx, ctypes.c_void_p(image.data_ptr()), rows, cols, channels, ctypes.c_int64(image.stride(0)), int(quality),
                                     _vp(out), cap, ctypes.byref(size), ctypes.c_void_p(stream.cuda_stream if stream is not None else 0))
        if rc == _lib.VFSMS_E_CAPACITY and size.value > cap:
            cap = size.value
            continue
        check(rc, "vfsms_jpeg_encode_dev")
        return out[:size.value].tobytes()


def jpeg_decode_gray_dev(datas, out, device=0, stream=None):
    """Decode into a device-resident tile stack: `out` is a torch.uint8 CUDA tensor [n, rows, cols] (any row / image stride)."""
    rows, cols, _ = jpeg_info(datas[0])
    assert out.is_cuda and out.dim() == 3 and out.shape[0] == len(datas) and out.shape[1] == rows and out.shape[2] == cols and out.stride(2) == 1
    bufs, ptrs, sizes = _jpeg_args(datas)
    _jpeg_check(_lib.load().vfsms_jpeg_decode_gray_dev(_lib.context(device), len(bufs), ptrs, sizes, ctypes.c_void_p(out.data_ptr()), rows, cols,
                                                       ctypes.c_int64(out.stride(1)), ctypes.c_int64(out.stride(0)),
                                                       ctypes.c_void_p(stream.cuda_stream if stream is not None else 0)),
                "vfsms_jpeg_decode_gray_dev")
    return out


# ---------------------------------------------------------------- device-resident tile stack
def tiles_reserve(n_tiles, rows, cols, device=0):
    check(_lib.load().vfsms_tiles_reserve(_lib.context(device), int(n_tiles), int(rows), int(cols)), "vfsms_tiles_reserve")


def tiles_decode_jpeg(first, datas, device=0):
    """Decode JPEG byte strings into stack slots first .. first + len(datas) - 1 (geometry = the reserved one)."""
    bufs, ptrs, sizes = _jpeg_args(datas)
    _jpeg_check(_lib.load().vfsms_tiles_decode_jpeg(_lib.context(device), int(first), len(bufs), ptrs, sizes), "vfsms_tiles_decode_jpeg")


def tiles_decode_jpeg_bgr(first, datas, device=0):
    """One entropy-decoding pass per file: BGR into the colour twin of the stack, the luma plane (= the gray decode) into the gray stack."""
    bufs, ptrs, sizes = _jpeg_args(datas)
    _jpeg_check(_lib.load().vfsms_tiles_decode_jpeg_bgr(_lib.context(device), int(first), len(bufs), ptrs, sizes), "vfsms_tiles_decode_jpeg_bgr")


def tiles_upload_bgr(first, tiles, device=0):
    t = np.ascontiguousarray(tiles, np.uint8)
    if t.ndim == 3:
        t = t[None]
    assert t.ndim == 4 and t.shape[3] == 3
    check(_lib.load().vfsms_tiles_upload_bgr(_lib.context(device), int(first), t.shape[0], t.ctypes.data_as(ctypes.c_void_p)), "vfsms_tiles_upload_bgr")


def tiles_mosaic_bgr(first, n_tiles, origins, rois, pair_offsets, method, canvas_shape, device=0):
    o = np.ascontiguousarray(origins, np.int32); r = np.ascontiguousarray(rois, np.int32); po = np.ascontiguousarray(pair_offsets, np.int32)
    out = np.empty((int(canvas_shape[0]), int(canvas_shape[1]), 3), np.uint8)
    check(_lib.load().vfsms_tiles_mosaic_bgr(_lib.context(device), int(first), int(n_tiles), o.ctypes.data_as(ctypes.c_void_p),
                                             r.ctypes.data_as(ctypes.c_void_p), po.ctypes.data_as(ctypes.c_void_p), FUSE_METHODS[method],
                                             out.shape[0], out.shape[1], out.ctypes.data_as(ctypes.c_void_p)), "vfsms_tiles_mosaic_bgr")
    return out


def tiles_upload(first, tiles, device=0):
    t = np.ascontiguousarray(tiles, np.uint8)
    if t.ndim == 2:
        t = t[None]
    check(_lib.load().vfsms_tiles_upload(_lib.context(device), int(first), t.shape[0], t.ctypes.data_as(ctypes.c_void_p)), "vfsms_tiles_upload")


def tiles_download(first, n, rows, cols, device=0):
    out = np.empty((n, rows, cols), np.uint8)
    check(_lib.load().vfsms_tiles_download(_lib.context(device), int(first), int(n), out.ctypes.data_as(ctypes.c_void_p)), "vfsms_tiles_download")
    return out


def tiles_align(first, n_pairs, direction, roi_len, params=None, ratio=0.75, offset_evaluate=3, device=0, step=1):
    """Candidate (direction, ROI length) of the incremental search for the pairs (first + p * step, first + p * step + 1), p < n_pairs,
    ROIs read in place from the tile stack.  -> structured array like align_batch."""
    p = params if params is not None else surf_params()
    res = np.zeros(n_pairs, PAIR_RESULT_DTYPE)
    check(_lib.load().vfsms_tiles_align_strided(_lib.context(device), int(first), int(n_pairs), int(step), int(direction), int(roi_len),
                                                ctypes.byref(p), ctypes.c_float(ratio), int(offset_evaluate),
                                                res.ctypes.data_as(ctypes.c_void_p)), "vfsms_tiles_align_strided")
    return res


def tiles_attach(lane=1, device=0):
    """Give the extra context `lane` of this device read access to lane 0's tile stack (vfsms_tiles_attach).  Call again after
    lane 0 re-reserves its stack."""
    check(_lib.load().vfsms_tiles_attach(_lib.context(device, lane), _lib.context(device)), "vfsms_tiles_attach")


def tiles_align_list(first_tiles, directions, roi_len, params=None, ratio=0.75, offset_evaluate=3, device=0, lane=0):
    """One fused call for an arbitrary list of candidates: pair p = stack tiles (first_tiles[p], first_tiles[p] + 1) in
    directions[p]; all directions must cut strips of one shape (1 / 3 or 2 / 4).  -> structured array like align_batch.
    lane: the context that runs the call (lane > 0 after tiles_attach(lane))."""
    p = params if params is not None else surf_params()
    ft = np.ascontiguousarray(first_tiles, np.int32); dr = np.ascontiguousarray(directions, np.int32)
    assert ft.ndim == 1 and ft.shape == dr.shape and len(ft) > 0
    res = np.zeros(len(ft), PAIR_RESULT_DTYPE)
    check(_lib.load().vfsms_tiles_align_list(_lib.context(device, lane), len(ft), ft.ctypes.data_as(ctypes.c_void_p), dr.ctypes.data_as(ctypes.c_void_p),
                                             int(roi_len), ctypes.byref(p), ctypes.c_float(ratio), int(offset_evaluate),
                                             res.ctypes.data_as(ctypes.c_void_p)), "vfsms_tiles_align_list")
    return res


def tiles_mosaic(first, n_tiles, origins, rois, pair_offsets, method, canvas_shape, device=0):
    o = np.ascontiguousarray(origins, np.int32); r = np.ascontiguousarray(rois, np.int32); po = np.ascontiguousarray(pair_offsets, np.int32)
    out = np.empty((int(canvas_shape[0]), int(canvas_shape[1])), np.uint8)
    check(_lib.load().vfsms_tiles_mosaic(_lib.context(device), int(first), int(n_tiles), o.ctypes.data_as(ctypes.c_void_p),
                                         r.ctypes.data_as(ctypes.c_void_p), po.ctypes.data_as(ctypes.c_void_p), FUSE_METHODS[method],
                                         out.shape[0], out.shape[1], out.ctypes.data_as(ctypes.c_void_p)), "vfsms_tiles_mosaic")
    return out
