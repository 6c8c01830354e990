"""The reference's three-function GPU plugin (appendix/myGpuFeatures.cpp:203-209), re-created over libvfsms.so.

Signatures, positional argument order and array layouts are the plugin's:
  detectAndDescribeBySurf(image, hessianThreshold, nOctaves, nOctaveLayers, isExtended, keypointsRatio, isUpright)
      -> float32 [N, D, 2]: [i, 0, 0] = x, [i, 1, 0] = y, [i, j, 1] = descriptor[j], everything else 0   (cpp:16-51, 67-104)
  detectAndDescribeByOrb(image, nFeatures, scaleFactor, nlevels, edgeThreshold, firstLevel, WTA_K, scoreType, patchSize,
                         fastThreshold, blurForDescriptor)   -> float32 [N, 32, 2], descriptor bytes as floats  (cpp:106-146)
  matchDescriptors(descA, descB, featureType, param) -> int32 [M, 2] rows (trainIdx, queryIdx)              (cpp:148-195)
Differences: empty results are empty arrays ((0, D, 2) / (0, 2)) instead of the plugin's `None`
(appendix/conversion.cpp:247-248), which crashed its callers (ImageUtility.py:231,243).  Nothing is computed on the CPU:
the first call creates the device context and raises if libvfsms.so or a CUDA device is missing.
"""
import numpy as np


def _pack(kp, desc):
    n, d = desc.shape
    out = np.zeros((n, d, 2), np.float32)
    if n:
        out[:, 0, 0] = kp[:, 0]
        out[:, 1, 0] = kp[:, 1]
        out[:, :, 1] = desc
    return out


def detectAndDescribeBySurf(image, hessianThreshold, nOctaves, nOctaveLayers, isExtended, keypointsRatio, isUpright):
    from imagestitch_b200 import gpu
    kp, desc = gpu.surf_detect_and_describe(np.asarray(image), hessian_threshold=hessianThreshold, n_octaves=nOctaves,
                                            n_octave_layers=nOctaveLayers, extended=isExtended, keypoints_ratio=keypointsRatio,
                                            upright=isUpright)
    return _pack(kp, desc)


def detectAndDescribeByOrb(image, nFeatures, scaleFactor, nlevels, edgeThreshold, firstLevel, WTA_K, scoreType, patchSize,
                           fastThreshold, blurForDescriptor):
    from imagestitch_b200 import gpu
    # scoreType is ignored by the plugin too (0 = HARRIS_SCORE hard-coded, cpp:115)
    kp, desc = gpu.orb_detect_and_describe(np.asarray(image), nFeatures, scaleFactor, nlevels, edgeThreshold, firstLevel, WTA_K,
                                           patchSize, fastThreshold)
    return _pack(kp, desc)


def matchDescriptors(descA, descB, featureType, param):
    from imagestitch_b200 import gpu
    return gpu.match_descriptors(np.asarray(descA, np.float32), np.asarray(descB, np.float32), int(featureType), float(param))
