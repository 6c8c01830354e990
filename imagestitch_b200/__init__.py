"""imagestitch_b200 -- B200-native pairwise-alignment hot path of Keep-Passion/ImageStitch (VFSMS).

Layout: csrc/ (hand-written sm_100a CUDA + the C ABI of include/vfsms.h), _lib.py (ctypes binding),
gpu.py (array-level host API), ImageUtility.py / Stitcher.py / ImageFusion.py (the reference's call surface),
myGpuFeatures/ (the reference's three-function plugin), synth.py (seeded synthetic tiles), sharding.py (multi-GPU).
"""
__version__ = "0.1.0"
