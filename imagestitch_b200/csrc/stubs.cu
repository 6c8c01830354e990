// stubs.cu -- every entry point of include/vfsms.h is implemented; kept so the build list stays stable.
#include "common.cuh"
