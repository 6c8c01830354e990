"""Drop-in module name of the reference (`import ImageFusion`, Stitcher.py:11) -> the B200 implementation."""
from imagestitch_b200.ImageFusion import ImageFusion  # noqa: F401
