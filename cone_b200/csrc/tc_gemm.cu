// placeholder until the tcgen05 path lands: CONE_PREC_TC is refused, never silently downgraded
#include "tc_gemm.h"

namespace cone {
struct TcWeights { int unused; };
int tc_weights_create(TcWeights**, cudaStream_t) {
    set_error("CONE_PREC_TC: the tensor-core path is not built in this library");
    return CONE_ERR_INVALID;
}
void tc_weights_destroy(TcWeights* t) { delete t; }
size_t tc_scratch_bytes(int64_t, int) { return 0; }
void tc_set_scratch(TcWeights*, void*, size_t) {}
bool tc_gemm_supported(int64_t, int, int) { return false; }
int tc_gemm(TcWeights*, const float*, int64_t, int64_t, const float*, const float*, int, int, float*, int64_t, int,
            const float*, int64_t, cudaStream_t) {
    set_error("tc_gemm: not built");
    return CONE_ERR_INVALID;
}
}  // namespace cone
