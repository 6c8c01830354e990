"""ctypes binding of libvfsms.so (include/vfsms.h).  There is no CPU fallback: a missing library or a missing
CUDA device raises."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libvfsms.so")

KP_STRIDE = 8

VFSMS_E_CAPACITY = -4


class SurfParams(ctypes.Structure):
    _fields_ = [("hessian_threshold", ctypes.c_float), ("n_octaves", ctypes.c_int), ("n_octave_layers", ctypes.c_int),
                ("extended", ctypes.c_int), ("keypoints_ratio", ctypes.c_float), ("upright", ctypes.c_int)]


class PairResult(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int32), ("d_row", ctypes.c_int32), ("d_col", ctypes.c_int32), ("votes", ctypes.c_int32),
                ("n_a", ctypes.c_int32), ("n_b", ctypes.c_int32), ("n_matches", ctypes.c_int32), ("flags", ctypes.c_int32)]


PAIR_RESULT_DTYPE = np.dtype([("status", "<i4"), ("d_row", "<i4"), ("d_col", "<i4"), ("votes", "<i4"),
                              ("n_a", "<i4"), ("n_b", "<i4"), ("n_matches", "<i4"), ("flags", "<i4")])

EXPORTS = [
    "vfsms_version", "vfsms_last_error", "vfsms_device_count", "vfsms_create", "vfsms_destroy", "vfsms_synchronize",
    "vfsms_stream", "vfsms_launch_count", "vfsms_surf_detect_and_describe", "vfsms_match_descriptors",
    "vfsms_orb_detect_and_describe", "vfsms_offset_by_mode", "vfsms_align_batch_host", "vfsms_align_batch_dev",
    "vfsms_match_batch_dev", "vfsms_phase_correlate_host", "vfsms_phase_correlate_dev", "vfsms_fuse_roi_host",
    "vfsms_mosaic_host", "vfsms_profile_enable", "vfsms_profile_read", "vfsms_stage_name",
    "vfsms_set_matcher", "vfsms_last_match_fallbacks", "vfsms_last_match_bound_violations", "vfsms_align_batch_upload", "vfsms_align_batch_run", "vfsms_tiles_attach", "vfsms_last_describe_handovers", "vfsms_enhance_host",
    "vfsms_jpeg_info", "vfsms_jpeg_luma_coefficients", "vfsms_jpeg_decode_gray_dev", "vfsms_jpeg_decode_gray_host",
    "vfsms_jpeg_component_coefficients", "vfsms_jpeg_decode_bgr_dev", "vfsms_jpeg_decode_bgr_host",
    "vfsms_tiles_reserve", "vfsms_tiles_decode_jpeg", "vfsms_tiles_upload", "vfsms_tiles_download", "vfsms_tiles_ptr",
    "vfsms_tiles_align", "vfsms_tiles_align_strided", "vfsms_tiles_align_list", "vfsms_tiles_mosaic", "vfsms_set_option", "vfsms_get_option", "vfsms_option_name",
    "vfsms_mosaic_band_host", "vfsms_overlap_sums_host", "vfsms_jpeg_encode_host", "vfsms_jpeg_encode_dev", "vfsms_jpeg_last_entropy_passes",
    "vfsms_tiles_decode_jpeg_bgr", "vfsms_tiles_upload_bgr", "vfsms_tiles_mosaic_bgr",
]
STAGE_COUNT = 12

_lib = None


class VfsmsError(RuntimeError):
    pass


def load():
    """Load libvfsms.so (building is the job of __graft_entry__.build / imagestitch_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise VfsmsError("libvfsms.so is not built (%s); run `python -m imagestitch_b200.build`. "
                         "There is no CPU fallback." % SO_PATH)
    L = ctypes.CDLL(SO_PATH)
    vp, i32, f32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int64
    L.vfsms_version.restype = i32
    L.vfsms_last_error.restype = ctypes.c_char_p
    L.vfsms_device_count.restype = i32
    L.vfsms_create.argtypes = [i32, ctypes.POINTER(vp)]
    L.vfsms_destroy.argtypes = [vp]
    L.vfsms_destroy.restype = None
    L.vfsms_synchronize.argtypes = [vp]
    L.vfsms_stream.argtypes = [vp]
    L.vfsms_stream.restype = vp
    L.vfsms_launch_count.argtypes = [vp]
    L.vfsms_launch_count.restype = i64
    L.vfsms_surf_detect_and_describe.argtypes = [vp, vp, i32, i32, i32, ctypes.POINTER(SurfParams), vp, vp, i32,
                                                 ctypes.POINTER(i32)]
    L.vfsms_match_descriptors.argtypes = [vp, vp, i32, vp, i32, i32, i32, f32, vp, ctypes.POINTER(i32)]
    L.vfsms_orb_detect_and_describe.argtypes = [vp, vp, i32, i32, i32, i32, f32, i32, i32, i32, i32, i32, i32, vp, vp, i32,
                                                ctypes.POINTER(i32)]
    L.vfsms_offset_by_mode.argtypes = [vp, vp, i32, vp, i32, i32, vp, i32, i32, ctypes.POINTER(PairResult)]
    L.vfsms_align_batch_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, i64, ctypes.POINTER(SurfParams), f32, i32, vp]
    L.vfsms_align_batch_dev.argtypes = [vp, vp, vp, i32, i32, i32, i32, i64, ctypes.POINTER(SurfParams), f32, i32, vp, vp]
    L.vfsms_tiles_attach.argtypes = [vp, vp]
    L.vfsms_align_batch_upload.argtypes = [vp, i32, vp, vp, i32, i32, i32, i32, i64]
    L.vfsms_align_batch_run.argtypes = [vp, i32, ctypes.POINTER(SurfParams), f32, i32, vp]
    L.vfsms_match_batch_dev.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, f32, vp, vp, vp]
    L.vfsms_phase_correlate_host.argtypes = [vp, vp, vp, i32, i32, i32, ctypes.POINTER(ctypes.c_double)]
    L.vfsms_phase_correlate_dev.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]
    L.vfsms_fuse_roi_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]
    L.vfsms_mosaic_host.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, i32, vp]
    L.vfsms_mosaic_band_host.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    L.vfsms_overlap_sums_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
    L.vfsms_jpeg_encode_host.argtypes = [vp, vp, i32, i32, i32, i64, i32, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    L.vfsms_jpeg_encode_dev.argtypes = [vp, vp, i32, i32, i32, i64, i32, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), vp]
    L.vfsms_set_matcher.argtypes = [vp, i32]
    L.vfsms_last_match_fallbacks.argtypes = [vp, ctypes.POINTER(i32)]
    L.vfsms_last_match_bound_violations.argtypes = [vp, ctypes.POINTER(i32)]
    L.vfsms_last_describe_handovers.argtypes = [vp, ctypes.POINTER(i32)]
    L.vfsms_set_option.argtypes = [vp, i32, i32]
    L.vfsms_get_option.argtypes = [vp, i32, ctypes.POINTER(i32)]
    L.vfsms_option_name.argtypes = [i32]
    L.vfsms_option_name.restype = ctypes.c_char_p
    L.vfsms_enhance_host.argtypes = [vp, vp, i32, i32, i32, i32, ctypes.c_double, i32, vp]
    L.vfsms_profile_enable.argtypes = [vp, i32]
    L.vfsms_profile_read.argtypes = [vp, vp, vp, i32]
    L.vfsms_stage_name.argtypes = [i32]
    L.vfsms_stage_name.restype = ctypes.c_char_p
    for name in EXPORTS:
        getattr(L, name)   # AttributeError here = header/library drift
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        raise VfsmsError("%s failed (%d): %s" % (what or "libvfsms call", rc, load().vfsms_last_error().decode()))


_contexts = {}


def context(device=0, lane=0):
    """One vfsms_ctx per (process, device); lane > 0: an extra context of that device (own stream and workspaces) for calls that
    should overlap with lane 0's, e.g. the second strip shape of a search round (gpu.tiles_attach)."""
    key = device if lane == 0 else (device, lane)
    if key not in _contexts:
        L = load()
        h = ctypes.c_void_p()
        check(L.vfsms_create(device, ctypes.byref(h)), "vfsms_create")
        _contexts[key] = h
    return _contexts[key]


def destroy_contexts():
    for h in _contexts.values():
        load().vfsms_destroy(h)
    _contexts.clear()
