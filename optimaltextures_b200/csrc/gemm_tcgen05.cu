#include "gemm_tc.cuh"
namespace optex {
int gemm_tc_rotate_forward(const float *, const float *, float *, int64_t, int, bool, int, cudaStream_t) {
    return OPTEX_ENOTSUP;
}
int gemm_tc_rotate_inverse(const float *, bool, const float *, float *, int64_t, int, const float *, float, int,
                           cudaStream_t) {
    return OPTEX_ENOTSUP;
}
}  // namespace optex
