"""ctypes face of oracle/surf_oracle.c -- TEST INFRASTRUCTURE ONLY (see the C file's header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsurf_oracle.so")
_lib = None

KP_STRIDE = 8
KP_X, KP_Y, KP_SIZE, KP_ANGLE, KP_RESPONSE, KP_OCTAVE, KP_LAPLACIAN = range(7)


def build(force=False):
    """Compile the C restatement (gcc, seconds). Building the checker is not using it."""
    src = os.path.join(_HERE, "surf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        c_u8p = ctypes.POINTER(ctypes.c_uint8)
        c_i32p = ctypes.POINTER(ctypes.c_int32)
        c_f32p = ctypes.POINTER(ctypes.c_float)
        L.so_integral.argtypes = [c_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i32p]
        L.so_integral.restype = None
        L.so_layer_det_trace.argtypes = [c_i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f32p, c_f32p]
        L.so_layer_det_trace.restype = None
        L.so_fast_atan2.argtypes = [ctypes.c_float, ctypes.c_float]
        L.so_fast_atan2.restype = ctypes.c_float
        L.so_gaussian_kernel.argtypes = [ctypes.c_int, ctypes.c_double, c_f32p]
        L.so_gaussian_kernel.restype = None
        L.so_resize_area_u8.argtypes = [c_u8p, ctypes.c_int, c_u8p, ctypes.c_int]
        L.so_resize_area_u8.restype = None
        L.so_window_patch.argtypes = [c_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                      ctypes.c_float, ctypes.c_float, ctypes.c_int, c_u8p, ctypes.c_int, c_u8p]
        L.so_window_patch.restype = ctypes.c_int
        L.so_surf_detect_and_compute.argtypes = [c_u8p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 c_f32p, c_f32p, ctypes.c_int]
        L.so_surf_detect_and_compute.restype = ctypes.c_int
        L.so_match_l2_ratio.argtypes = [c_f32p, ctypes.c_int, c_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                        c_i32p, c_f32p, c_i32p]
        L.so_match_l2_ratio.restype = ctypes.c_int
        L.so_offset_by_mode.argtypes = [c_f32p, c_f32p, ctypes.c_int, c_i32p, ctypes.c_int, ctypes.c_int, c_i32p]
        L.so_offset_by_mode.restype = ctypes.c_int
        L.so_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _img(image):
    img = np.asarray(image)
    assert img.dtype == np.uint8 and img.ndim == 2
    if img.strides[1] != 1:
        img = np.ascontiguousarray(img)
    return img


def integral(image):
    img = _img(image)
    h, w = img.shape
    out = np.empty((h + 1, w + 1), np.int32)
    lib().so_integral(_p(img, ctypes.c_uint8), h, w, img.strides[0], _p(out, ctypes.c_int32))
    return out


def layer_det_trace(sum_, size, step):
    h, w = sum_.shape[0] - 1, sum_.shape[1] - 1
    det = np.zeros((h // step, w // step), np.float32)
    tr = np.zeros_like(det)
    s = np.ascontiguousarray(sum_, np.int32)
    lib().so_layer_det_trace(_p(s, ctypes.c_int32), h, w, size, step, _p(det, ctypes.c_float), _p(tr, ctypes.c_float))
    return det, tr


def fast_atan2(y, x):
    return lib().so_fast_atan2(float(y), float(x))


def gaussian_kernel(n, sigma):
    out = np.empty(n, np.float32)
    lib().so_gaussian_kernel(n, float(sigma), _p(out, ctypes.c_float))
    return out


def resize_area_u8(win, d=21):
    win = np.ascontiguousarray(win, np.uint8)
    n = win.shape[0]
    assert win.shape == (n, n)
    out = np.empty((d, d), np.uint8)
    lib().so_resize_area_u8(_p(win, ctypes.c_uint8), n, _p(out, ctypes.c_uint8), d)
    return out


def window_patch(image, x, y, size, angle, upright=False):
    img = _img(image)
    h, w = img.shape
    s = np.float32(size) * np.float32(1.2) / np.float32(9.0)
    n = int(np.float32(21) * s)
    win = np.empty((n, n), np.uint8)
    patch = np.empty((21, 21), np.uint8)
    r = lib().so_window_patch(_p(img, ctypes.c_uint8), h, w, img.strides[0], x, y, size, angle, int(upright),
                              _p(win, ctypes.c_uint8), n * n, _p(patch, ctypes.c_uint8))
    assert r == n, (r, n)
    return win, patch


def detect_and_compute(image, hessian_threshold=100.0, n_octaves=4, n_octave_layers=3, extended=False,
                       upright=False, max_features=0, want_desc=True):
    """Returns (kp [N,8] float32, desc [N,64|128] float32)."""
    img = _img(image)
    h, w = img.shape
    D = 128 if extended else 64
    cap = 1 << 14
    while True:
        kp = np.zeros((cap, KP_STRIDE), np.float32)
        desc = np.zeros((cap, D), np.float32) if want_desc else None
        n = lib().so_surf_detect_and_compute(_p(img, ctypes.c_uint8), h, w, img.strides[0], hessian_threshold,
                                             n_octaves, n_octave_layers, int(extended), int(upright), int(max_features),
                                             _p(kp, ctypes.c_float),
                                             _p(desc, ctypes.c_float) if want_desc else None, cap)
        if n >= 0:
            return kp[:n].copy(), (desc[:n].copy() if want_desc else None)
        cap = -n


def match_l2_ratio(descA, descB, ratio=0.75, want_raw=False):
    A = np.ascontiguousarray(descA, np.float32)
    B = np.ascontiguousarray(descB, np.float32)
    nA, D = A.shape
    nB = B.shape[0]
    out = np.empty((max(nA, 1), 2), np.int32)
    dist = np.empty((max(nA, 1), 2), np.float32)
    idx = np.empty((max(nA, 1), 2), np.int32)
    m = lib().so_match_l2_ratio(_p(A, ctypes.c_float), nA, _p(B, ctypes.c_float), nB, D, ratio,
                                _p(out, ctypes.c_int32), _p(dist, ctypes.c_float), _p(idx, ctypes.c_int32))
    if want_raw:
        return out[:m].copy(), dist[:nA], idx[:nA]
    return out[:m].copy()


def offset_by_mode(kpsA, kpsB, matches, evaluate=3):
    """kps: [N,>=2] float32 (x, y, ...). matches [M,2] int32 (trainIdx, queryIdx). -> (status, [dRow, dCol], votes)."""
    A = np.ascontiguousarray(kpsA, np.float32)
    B = np.ascontiguousarray(kpsB, np.float32)
    m = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
    out = np.zeros(3, np.int32)
    assert A.shape[1] == B.shape[1]
    st = lib().so_offset_by_mode(_p(A, ctypes.c_float), _p(B, ctypes.c_float), A.shape[1], _p(m, ctypes.c_int32),
                                 m.shape[0], evaluate, _p(out, ctypes.c_int32))
    return bool(st), [int(out[0]), int(out[1])], int(out[2])


def num_threads():
    return lib().so_num_threads()
