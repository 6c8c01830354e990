"""Drop-in module name of the reference (`import ImageUtility as Utility`, Stitcher.py:10) -> the B200 implementation."""
from imagestitch_b200.ImageUtility import Method  # noqa: F401
