// tcgen05 (5th-gen tensor core) rotation GEMMs - see gemm_tcgen05.cu.
#pragma once
#include "common.cuh"

namespace optex {

// internal: "this shape/alignment is not handled by the tensor-core path" (caller falls back to
// the fp32 SIMT tiles when the GEMM mode is AUTO, or reports OPTEX_ESIZE when it was forced)
constexpr int OPTEX_ENOTSUP = -1;

// dst = X R  (transposed: dst[c, n], else dst[n, c]);  terms = 1 (TF32) or 3 (3xTF32 split)
int gemm_tc_rotate_forward(const float *X, const float *R, float *dst, int64_t n, int c, bool transposed,
                           int terms, cudaStream_t st);
// out[n, j] = sum_c M(n, c) R[j, c] (+ content blend);  M channel-major [c, n] or NHWC [n, c]
int gemm_tc_rotate_inverse(const float *M, bool m_channel_major, const float *R, float *out, int64_t n, int c,
                           const float *content, float strength, int terms, cudaStream_t st);

}  // namespace optex
