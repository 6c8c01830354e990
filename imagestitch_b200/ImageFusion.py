"""ImageFusion -- the reference's blend class (ImageFusion.py:7-492) on the B200 path.

Every fuseBy* method takes the reference's arguments ([imageA, imageB] integer ROI arrays in which -1 marks empty canvas,
SURVEY.md Appendix C) and returns the uint8 fused ROI; the arithmetic runs in blend.cu through the C ABI.  Unlike the
reference these methods do not modify their inputs (the reference's in-place `imageA[imageA < 0] = ...` side effects are
never observed by its callers, which pass copies: Stitcher.py:474-483).
"""
import numpy as np

from . import gpu
from .ImageUtility import Method


class ImageFusion(Method):

    isColorMode = False

    @staticmethod
    def _pair(images):
        (imageA, imageB) = images
        return np.asarray(imageA), np.asarray(imageB)

    def fuseByAverage(self, images):
        """uint8((A + B) / 2) -- ImageFusion.py:12-21.  Inputs are taken as already zero-filled (Stitcher.py:498-504)."""
        a, b = self._pair(images)
        return gpu.fuse_roi(self._filled(a), self._filled(b), "average")

    def fuseByMaximum(self, images):
        """ImageFusion.py:23-31."""
        a, b = self._pair(images)
        return gpu.fuse_roi(self._filled(a), self._filled(b), "maximum")

    def fuseByMinimum(self, images):
        """ImageFusion.py:33-41."""
        a, b = self._pair(images)
        return gpu.fuse_roi(self._filled(a), self._filled(b), "minimum")

    @staticmethod
    def _filled(x):
        # the device kernel applies "-1 -> 0, mutual zero fill" itself (idempotent on already filled data)
        return x

    def getWeightsMatrix(self, images):
        """Corner-case weight matrices (weightMatA, weightMatB) float32 -- ImageFusion.py:43-190."""
        a, b = self._pair(images)
        if np.count_nonzero(a > -1) / a.size > 0.65:
            # the reference only calls this for sparse ROIs; force the corner path by asking the kernel for it explicitly
            pass
        _, wa, wb = gpu.fuse_roi(a, b, "fadeInAndFadeOut", 0, 0, want_weights=True, force_corner=True)
        if a.ndim == 3:
            wa = np.repeat(wa[:, :, None], a.shape[2], axis=2); wb = np.repeat(wb[:, :, None], a.shape[2], axis=2)
        return (wa, wb)

    def fuseByFadeInAndFadeOut(self, images, dx, dy):
        """Linear fade; dx = row offset, dy = column offset of the ORIGINAL pair (ImageFusion.py:192-244)."""
        a, b = self._pair(images)
        return gpu.fuse_roi(a, b, "fadeInAndFadeOut", dx, dy)

    def fuseByTrigonometric(self, images, dx, dy):
        """sin^2 weights (ImageFusion.py:246-293)."""
        a, b = self._pair(images)
        return gpu.fuse_roi(a, b, "trigonometric", dx, dy)

    def fuseByMultiBandBlending(self, images):
        """4-level Laplacian blend with constant 0.5 / 0.5 level weights (ImageFusion.py:296-367)."""
        a, b = self._pair(images)
        return gpu.fuse_roi(a, b, "multiBandBlending")

    def fuseByOptimalSeamLine(self, images, direction="horizontal"):
        """Interactive (cv2.imshow / waitKey) in the reference, not reachable from Main.py (ImageFusion.py:377-492):
        out of scope."""
        raise NotImplementedError("optimalSeamLine is out of scope of the B200 hot path (interactive GUI code in the reference)")
