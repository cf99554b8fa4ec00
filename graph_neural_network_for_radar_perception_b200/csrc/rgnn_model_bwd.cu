// Backward of the detector (placeholder until the backward tile programs land).
#include "rgnn_model.h"

namespace rgnn {
void plan_detector_bwd(const rgnn_detector&, const rgnn_graph&, const TakeFn&, DetPlan* pl) {
    pl->dx = pl->dx2 = pl->dP = pl->dagg = pl->demb = pl->dh = pl->dg = pl->scratch = nullptr;
    pl->scratch_floats = 0;
}
int launch_bwd(const Program&, cudaStream_t) { set_error("backward not built"); return RGNN_ERR_INVALID; }
}  // namespace rgnn

extern "C" size_t rgnn_ffn_stack_bwd_workspace_bytes(const rgnn_stack*) { return 0; }
extern "C" int rgnn_ffn_stack_bwd(const rgnn_stack*, const float*, const float*, int, float*, void*, size_t, void*) {
    rgnn::set_error("backward not built");
    return RGNN_ERR_INVALID;
}
extern "C" int rgnn_detector_bwd(const rgnn_detector*, const rgnn_graph*, const float*, const float*, const float*,
                                 const float*, const float*, const float*, void*, size_t, void*) {
    rgnn::set_error("backward not built");
    return RGNN_ERR_INVALID;
}
