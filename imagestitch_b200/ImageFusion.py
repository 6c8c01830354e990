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
        """uint8((A + B) / 2) on the arrays as given -- ImageFusion.py:12-21.  (The -1 -> 0 and mutual zero fill belong to
        Stitcher.fuseImage, Stitcher.py:498-504, which goes through gpu.fuse_roi without `raw`.)"""
        a, b = self._pair(images)
        return gpu.fuse_roi(a, b, "average", raw=True)

    def fuseByMaximum(self, images):
        """ImageFusion.py:23-31."""
        a, b = self._pair(images)
        return gpu.fuse_roi(a, b, "maximum", raw=True)

    def fuseByMinimum(self, images):
        """ImageFusion.py:33-41."""
        a, b = self._pair(images)
        return gpu.fuse_roi(a, b, "minimum", raw=True)

    def getWeightsMatrix(self, images):
        """Corner-case weight matrices (weightMatA, weightMatB) float32 -- ImageFusion.py:43-190."""
        a, b = self._pair(images)
        # the reference only calls this for sparse ROIs (ImageFusion.py:209); the corner path is requested explicitly
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
