"""Import the UNMODIFIED reference modules from /root/reference (read-only) -- TEST INFRASTRUCTURE ONLY.

Used in THIS container only (fixture generation + pin tests that are skipped when /root/reference is absent);
nothing under `-m gpu`, smoke() or bench.py touches it.  Shims follow SURVEY.md Appendix E:
  * stub package `myGpuFeatures` (ImageUtility.py:4 imports it unconditionally),
  * `np.int = int` (removed in NumPy >= 1.24; used at Stitcher.py:198-199,231-232,434,436 ...),
  * `cv2.xfeatures2d` namespace: SIFT from cv2.SIFT_create; SURF_create -> the C restatement in oracle/surf.py
    (cv2 in this image has no contrib/non-free build).
"""
import os
import sys
import types

REFERENCE = "/root/reference"


def available():
    return os.path.isdir(REFERENCE) and os.path.exists(os.path.join(REFERENCE, "Stitcher.py"))


class _OracleSurf:
    """Duck-types the cv2.xfeatures2d.SURF object as the reference uses it (ImageUtility.py:258,262-264)."""

    def __init__(self, hessianThreshold=100, nOctaves=4, nOctaveLayers=3, extended=False, upright=False):
        self.args = (float(hessianThreshold), nOctaves, nOctaveLayers, extended, upright)

    def detectAndCompute(self, image, mask):
        from oracle import surf
        h, o, l, e, u = self.args
        kp, desc = surf.detect_and_compute(image, h, o, l, e, u)
        kps = [types.SimpleNamespace(pt=(float(r[0]), float(r[1]))) for r in kp]
        return kps, desc


def import_reference():
    """Returns the reference's (Stitcher module, ImageUtility module, ImageFusion module)."""
    import numpy as np
    import cv2
    if not available():
        raise RuntimeError("reference tree not present")
    if not hasattr(np, "int"):
        np.int = int
    if "myGpuFeatures" not in sys.modules or not getattr(sys.modules["myGpuFeatures"], "_oracle_stub", False):
        pkg = types.ModuleType("myGpuFeatures")
        pkg._oracle_stub = True
        mod = types.ModuleType("myGpuFeatures.myGpuFeatures")

        def _no(*a, **k):
            raise NotImplementedError("reference GPU plugin is a Windows binary; stubbed for import only")
        mod.detectAndDescribeBySurf = mod.detectAndDescribeByOrb = mod.matchDescriptors = _no
        pkg.myGpuFeatures = mod
        saved = {k: sys.modules.get(k) for k in ("myGpuFeatures", "myGpuFeatures.myGpuFeatures")}
        sys.modules["myGpuFeatures"] = pkg
        sys.modules["myGpuFeatures.myGpuFeatures"] = mod
    else:
        saved = {}
    if not hasattr(cv2, "xfeatures2d"):
        cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create, SURF_create=_OracleSurf)
    saved_mods = {k: sys.modules.pop(k, None) for k in ("Stitcher", "ImageUtility", "ImageFusion")}
    sys.path.insert(0, REFERENCE)
    try:
        import importlib
        ref_util = importlib.import_module("ImageUtility")
        ref_fusion = importlib.import_module("ImageFusion")
        ref_stitcher = importlib.import_module("Stitcher")
    finally:
        sys.path.remove(REFERENCE)
        for k in ("Stitcher", "ImageUtility", "ImageFusion"):
            sys.modules.pop(k, None)
            if saved_mods[k] is not None:
                sys.modules[k] = saved_mods[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
    ref_stitcher.Stitcher.isPrintLog = False
    return ref_stitcher, ref_util, ref_fusion


def golden_offsets():
    """The 89 offsets the author left as a comment at Stitcher.py:87."""
    line = open(os.path.join(REFERENCE, "Stitcher.py"), encoding="utf8").read().splitlines()[86]
    return eval(line.split("=", 1)[1])
