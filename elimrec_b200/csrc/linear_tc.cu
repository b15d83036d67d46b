// Tensor-core (tcgen05 / TMEM) linear layers - placeholder translation unit.
// Until the UMMA kernels land, both entry points report "unsupported shape" (-2) and the host
// routes the projection layers through the exact-fp32 elimrec_gemm path.
#include "common.cuh"

ELIMREC_API int elimrec_linear_tf32_fwd(int64_t, int64_t, const float*, int64_t, const float*, const float*, float*,
                                        int64_t, elimrec_stream_t) {
    elimrec_set_error("elimrec_linear_tf32_fwd: not built in this revision");
    return -2;
}
ELIMREC_API int elimrec_linear_tf32_wgrad(int64_t, int64_t, const float*, int64_t, const float*, int64_t, float*, float*,
                                          elimrec_stream_t) {
    elimrec_set_error("elimrec_linear_tf32_wgrad: not built in this revision");
    return -2;
}
ELIMREC_API int64_t elimrec_linear_tf32_wgrad_workspace_floats(int64_t, int64_t) { return 0; }
