// placeholder until the BNN kernels land
#include "../../include/pddp_b200.h"
extern "C" int64_t pddp_bnn_workspace_bytes(const pddp_shape*, const pddp_bnn*, int32_t) { return PDDP_E_UNSUPPORTED; }
extern "C" int pddp_linearize_bnn(const pddp_shape*, const pddp_bnn*, const pddp_cost*, const void*, const void*, const void*, const void*, const int32_t*, void*, void*, void*, void*, void*, void*, void*, void*, void*, void*, int32_t*, void*, int64_t, void*) { return PDDP_E_UNSUPPORTED; }
extern "C" int pddp_rollout_bnn(const pddp_shape*, const pddp_bnn*, const pddp_cost*, const void*, const void*, const void*, const void*, const void*, int32_t, const void*, const void*, const int32_t*, const int32_t*, void*, int32_t*, void*, void*, void*, int32_t*, void*, int64_t, void*) { return PDDP_E_UNSUPPORTED; }
