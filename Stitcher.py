"""Drop-in module name of the reference (`from Stitcher import Stitcher`, Main.py:1) -> the B200 implementation."""
from imagestitch_b200.Stitcher import ImageFeature, Stitcher  # noqa: F401
