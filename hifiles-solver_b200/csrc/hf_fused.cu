// placeholder until the fused tensor-product kernels land
#include "hf_device.h"
int hf_fused_available(hf_ctx *) { return 0; }
int hf_fused_prepare(hf_ctx *) { return 0; }
int hf_fused_stage(hf_ctx *, int, double, int, int) { hf_set_error("fused path not built"); return 1; }
int hf_fused_extrapolate(hf_ctx *) { hf_set_error("fused path not built"); return 1; }
