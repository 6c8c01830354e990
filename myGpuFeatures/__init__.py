"""Package `myGpuFeatures` -- the reference unpacks its Windows plugin next to the scripts under this name and imports it
unconditionally (`from myGpuFeatures import myGpuFeatures`, ImageUtility.py:4).  Importing never needs a GPU."""
