// stubs.cu -- entry points declared in include/vfsms.h whose kernels are not written yet.  They fail loudly.
#include "common.cuh"

void phase_state_destroy(vfsms_ctx *) {}

extern "C" {
int vfsms_orb_detect_and_describe(vfsms_ctx *, const uint8_t *, int, int, int, int, float, int, int, int, int, int, int,
                                  float *, float *, int, int *)
{ vfsms_set_error("vfsms_orb_detect_and_describe: not implemented yet"); return VFSMS_E_UNSUPPORTED; }
int vfsms_phase_correlate_host(vfsms_ctx *, const uint8_t *, const uint8_t *, int, int, int, double *)
{ vfsms_set_error("vfsms_phase_correlate_host: not implemented yet"); return VFSMS_E_UNSUPPORTED; }
int vfsms_phase_correlate_dev(vfsms_ctx *, const uint8_t *, const uint8_t *, int, int, int, double *, void *)
{ vfsms_set_error("vfsms_phase_correlate_dev: not implemented yet"); return VFSMS_E_UNSUPPORTED; }
}
