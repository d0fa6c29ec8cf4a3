// placeholder until the tcgen05 kernel lands
#include "pst_internal.h"
int pst_prepare_tc_weights(pst_model*) { return PST_ERR_UNSUPPORTED_CONFIG; }
int pst_launch_edge_mlp_tc(const pst_model*, cudaStream_t, int, int, float*, const float*, const float*,
                           const int32_t*, const int32_t*, int, int, float*) { return PST_ERR_UNSUPPORTED_CONFIG; }
