"""ImageUtility.Method -- the reference's utility base class (ImageUtility.py:7-368) on the B200 path.

Same class name, attribute names and method signatures as the reference so that `Stitcher` / `ImageFusion` subclasses
and `Main.py` keep working; the arithmetic of the hot path (SURF / ORB detect+describe, brute-force match + ratio test,
offset vote) runs in libvfsms.so through imagestitch_b200.gpu.  There is no CPU implementation of those here: without
the library or a CUDA device the calls raise.  Two pieces stay on cv2 because the north star does not name them and
SURVEY.md section 8 marks them "secondary": SIFT (a7, ImageUtility.py:256,268) and the RANSAC variant (a10,
ImageUtility.py:180-210).
"""
import math

import numpy as np

from . import gpu


class Method():
    # printing (ImageUtility.py:8-12)
    outputAddress = "result/"
    isEvaluate = False
    evaluateFile = "evaluate.txt"
    isPrintLog = True

    # feature search (ImageUtility.py:14-17)
    featureMethod = "surf"      # "sift", "surf" or "orb"
    roiRatio = 0.1
    searchRatio = 0.75

    # In the reference this flag picks cv2-CPU vs the myGpuFeatures plugin (ImageUtility.py:254,285).  Here every
    # value runs on the GPU; the flag only selects which PARAMETER SET the reference would have used:
    #   False -> cv2.xfeatures2d.SURF_create() defaults (64-d, no keypoint cap), BF match without the ORB distance filter
    #   True  -> the plugin call's arguments below (128-d, keypointsRatio cap, orbMaxDistance)
    isGPUAvailable = False

    # GPU-SURF (ImageUtility.py:22-28)
    surfHessianThreshold = 100.0
    surfNOctaves = 4
    surfNOctaveLayers = 3
    surfIsExtended = True
    surfKeypointsRatio = 0.01
    surfIsUpright = False

    # GPU-ORB (ImageUtility.py:30-40)
    orbNfeatures = 5000
    orbScaleFactor = 1.2
    orbNlevels = 8
    orbEdgeThreshold = 31
    orbFirstLevel = 0
    orbWTA_K = 2
    orbPatchSize = 31
    orbFastThreshold = 20
    orbBlurForDescriptor = False
    orbMaxDistance = 30

    # registration (ImageUtility.py:42-44)
    offsetCaculate = "mode"     # "mode" or "ransac"
    offsetEvaluate = 3

    # enhancement (ImageUtility.py:46-50)
    isEnhance = False
    isClahe = False
    clipLimit = 20
    tileSize = 5

    # ------------------------------------------------------------------ logging
    def printAndWrite(self, content):
        """ImageUtility.py:52-64."""
        if self.isPrintLog:
            print(content)
        if self.isEvaluate:
            with open(self.outputAddress + self.evaluateFile, "a") as f:
                f.write(content)
                f.write("\n")

    # ------------------------------------------------------------------ ROI strips
    def getROIRegionForIncreMethod(self, image, direction=1, order="first", searchRatio=0.1):
        """Edge strip used by the incremental search (ImageUtility.py:66-101).  Returns a view (no copy);
        direction 2 / 4 strips are non-contiguous and are consumed as strided input by the C ABI."""
        row, col = image.shape[:2]
        first = order == "first"
        if direction in (1, 3):
            n = int(np.floor(row * searchRatio))
            bottom = (direction == 1) == first
            return image[row - n:row, :] if bottom else image[0:n, :]
        if direction in (2, 4):
            n = int(np.floor(col * searchRatio))
            right = (direction == 2) == first
            return image[:, col - n:col] if right else image[:, 0:n]
        return np.zeros(image.shape, np.uint8)

    def getROIRegion(self, image, direction="horizontal", order="first", searchLength=150, searchLengthForLarge=-1):
        """Fixed-length strip, deprecated in the reference (ImageUtility.py:103-137)."""
        row, col = image.shape[:2]
        big = searchLengthForLarge
        roi = None
        if direction in ("horizontal", 2):
            if order == "first":
                roi = image[:, col - searchLength:col] if big == -1 else (image[row - big:row, col - searchLength:col] if big > 0 else None)
            elif order == "second":
                roi = image[:, 0:searchLength] if big == -1 else (image[0:big, 0:searchLength] if big > 0 else None)
        elif direction in ("vertical", 1):
            if order == "first":
                roi = image[row - searchLength:row, :] if big == -1 else (image[row - searchLength:row, col - big:col] if big > 0 else None)
            elif order == "second":
                roi = image[0:searchLength, :] if big == -1 else (image[0:searchLength, 0:big] if big > 0 else None)
        return roi

    # ------------------------------------------------------------------ offset estimation
    def getOffsetByMode(self, kpsA, kpsB, matches, offsetEvaluate=10):
        """Mode of the truncated per-match offsets, first-seen tie rule, (0, 0) votes dropped
        (ImageUtility.py:139-178) -- device-side vote table instead of the O(M^2) list.count loop."""
        if len(matches) == 0:
            return (False, [0, 0])
        a = np.asarray(kpsA, np.float32).reshape(len(kpsA), -1)[:, :2]
        b = np.asarray(kpsB, np.float32).reshape(len(kpsB), -1)[:, :2]
        status, offset, _ = gpu.offset_by_mode(a, b, np.asarray(matches, np.int32).reshape(-1, 2), offsetEvaluate)
        return (status, offset)

    def getOffsetByRansac(self, kpsA, kpsB, matches, offsetEvaluate=100):
        """cv2 passthrough of the self-labelled incomplete variant (ImageUtility.py:180-210); not on the hot path.
        The reference calls cv2.getAffineTransform on all N points first, which raises for N != 3 (and for N == 3 the
        findHomography after it raises): the function cannot return there for any non-empty input.  That unused call is
        dropped here, everything else is kept (tests/test_sequencing_reference_cpu.py pins both behaviours)."""
        import cv2
        if len(matches) == 0:
            return (False, [0, 0], 0)
        ptsA = np.float32([kpsA[i] for (_, i) in matches])
        ptsB = np.float32([kpsB[i] for (i, _) in matches])
        (H, status) = cv2.findHomography(ptsA, ptsB, cv2.RANSAC, 3, 0.9)
        if H is None:
            return (False, [0, 0], 0)
        trueCount = int(np.count_nonzero(status))
        if trueCount >= offsetEvaluate:
            adjustH = H.copy()
            adjustH[0, 2] = 0; adjustH[1, 2] = 0
            adjustH[2, 0] = 0; adjustH[2, 1] = 0
            Hi = np.array(H).astype(int)
            return (True, [int(np.round(Hi[1, 2]) * (-1)), int(np.round(Hi[0, 2]) * (-1))], adjustH)
        return (False, [0, 0], 0)

    # ------------------------------------------------------------------ plugin-array helpers (ImageUtility.py:212-246)
    def npToListForKeypoints(self, array):
        return [[array[i, 0], array[i, 1]] for i in range(array.shape[0])]

    def npToListForMatches(self, array):
        return [(array[i, 0], array[i, 1]) for i in range(array.shape[0])]

    def npToKpsAndDescriptors(self, array):
        kps = [[array[i, 0, 0], array[i, 1, 0]] for i in range(array.shape[0])]
        return (kps, array[:, :, 1])

    # ------------------------------------------------------------------ detect / describe / match
    def _surf_params(self):
        if self.isGPUAvailable:
            return gpu.surf_params(self.surfHessianThreshold, self.surfNOctaves, self.surfNOctaveLayers, self.surfIsExtended,
                                   self.surfKeypointsRatio, self.surfIsUpright)
        return gpu.surf_params(**gpu.SURF_CPU_DEFAULTS)

    def detectAndDescribe(self, image, featureMethod):
        """(kps float32 [N, 2] as (x, y), features float32 [N, D]) -- ImageUtility.py:248-276."""
        if featureMethod == "surf":
            kp, desc = gpu.surf_detect_and_describe(image, params=self._surf_params())
            return (np.ascontiguousarray(kp[:, :2]), desc)
        if featureMethod == "orb":
            kp, desc = gpu.orb_detect_and_describe(image, self.orbNfeatures, self.orbScaleFactor, self.orbNlevels,
                                                   self.orbEdgeThreshold, self.orbFirstLevel, self.orbWTA_K, self.orbPatchSize,
                                                   self.orbFastThreshold)
            return (np.ascontiguousarray(kp[:, :2]), desc)
        if featureMethod == "sift":
            import cv2                      # not named by the north star: cv2 passthrough like the reference's GPU mode (:266-269)
            create = getattr(getattr(cv2, "xfeatures2d", None), "SIFT_create", None) or cv2.SIFT_create
            kps, features = create().detectAndCompute(np.ascontiguousarray(image), None)
            return (np.float32([kp.pt for kp in kps]).reshape(-1, 2), features)
        raise ValueError("unknown featureMethod %r" % (featureMethod,))

    def matchDescriptors(self, featuresA, featuresB):
        """List of (trainIdx, queryIdx) in ascending queryIdx -- ImageUtility.py:278-309."""
        fa = np.asarray(featuresA, np.float32)
        fb = np.asarray(featuresB, np.float32)
        if self.featureMethod in ("surf", "sift"):
            m = gpu.match_descriptors(fa, fb, 2, self.searchRatio)
        elif self.featureMethod == "orb":
            # CPU branch: best-1 Hamming without a distance filter (:298-302); plugin branch: distance < orbMaxDistance (:308)
            m = gpu.match_descriptors(fa, fb, 3, float(self.orbMaxDistance) if self.isGPUAvailable else 1e9)
        else:
            raise ValueError("unknown featureMethod %r" % (self.featureMethod,))
        return [(int(t), int(q)) for t, q in m]

    # ------------------------------------------------------------------ misc (not on the hot path)
    def resizeImg(self, image, resizeTimes, interMethod=None):
        import cv2
        (h, w) = image.shape
        return cv2.resize(image, (int(w * resizeTimes), int(h * resizeTimes)),
                          interpolation=cv2.INTER_AREA if interMethod is None else interMethod)

    def rectifyFinalImg(self, image, regionLength=10):
        """Test helper of the reference (ImageUtility.py:324-367): rotate when two opposite corners are empty."""
        import cv2
        (h, w) = image.shape
        ul = np.sum(image[0:regionLength, 0:regionLength]); ur = np.sum(image[0:regionLength, w - regionLength:w])
        bl = np.sum(image[h - regionLength:h, 0:regionLength]); br = np.sum(image[h - regionLength:h, w - regionLength:w])
        if (np.count_nonzero(image[:, 0]) / h) < 0.3:
            return image
        center = (w // 2, h // 2)
        angle = math.atan(center[1] / center[0] * 180 / math.pi)
        if ul == 0 and br == 0 and ur != 0 and bl != 0:
            return cv2.warpAffine(image, cv2.getRotationMatrix2D(center, -1 * angle, 1.0), (w, h))
        if ul != 0 and br != 0 and ur == 0 and bl == 0:
            return cv2.warpAffine(image, cv2.getRotationMatrix2D(center, angle, 1.0), (w, h))
        return image
