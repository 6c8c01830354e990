// stubs.cu -- entry points declared in include/vfsms.h whose kernels are not written yet.  They fail loudly.
#include "common.cuh"


extern "C" {
int vfsms_orb_detect_and_describe(vfsms_ctx *, const uint8_t *, int, int, int, int, float, int, int, int, int, int, int,
                                  float *, float *, int, int *)
{ vfsms_set_error("vfsms_orb_detect_and_describe: not implemented yet"); return VFSMS_E_UNSUPPORTED; }
}
