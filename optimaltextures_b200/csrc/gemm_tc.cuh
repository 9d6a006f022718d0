// tcgen05 (5th-gen tensor core) rotation GEMMs - see gemm_tcgen05.cu.
#pragma once
#include "common.cuh"

namespace optex {

// internal: "this shape/alignment is not handled by the tensor-core path" (caller falls back to
// the fp32 SIMT tiles when the GEMM mode is AUTO, or reports OPTEX_ESIZE when it was forced)
constexpr int OPTEX_ENOTSUP = -1;

// General dense fp32 GEMM on the tensor cores:  D = alpha * op(A) op(B)  (+ bias) (then content blend)
//   a_mn == false : A is row-major [M, K] (K-major);   a_mn == true : A is row-major [K, M] (MN-major, A^T stored)
//   b_mn == false : B is row-major [N, K] (K-major, i.e. D = A B^T);  b_mn == true : B is row-major [K, N]
//   d_trans       : D stored [N, M] (only with a K-major A and an MN-major B: the forward rotation)
//   split_k > 1   : blockIdx.z splits K; partial sums go to D + z * d_z_stride (caller reduces them)
struct TcGemm {
    const float *A;
    bool a_mn;
    const float *B;
    bool b_mn;
    int64_t ldb;  // MN-major B only: row pitch when B is a column block of a wider matrix (0 = dense)
    int64_t b_col0;  // ... and the first column of that block (B points at the full matrix)
    float *D;
    int64_t ldd;
    bool d_trans;
    int64_t M, N, K;
    int terms;  // 1 = TF32, 3 = 3xTF32
    int split_k;
    int64_t d_z_stride;
    const float *blend;
    float strength;
    const float *bias;
    int64_t bias_hw, bias_ld;
    float alpha;
    const int *skip;
    bool stream_a;       // A is not needed again after this GEMM: load it with L2 evict-first priority
    bool presplit_b;     // 3xTF32: split B with an element-wise pass even though M is small (B is the big operand)
    uint32_t *rowrange;  // !d_trans: per output row, the same fold (rows = channels in the fused loop)
    uint32_t *colrange;  // d_trans only: per output column, atomicMin of (f2ord(v), ~f2ord(v)) - see cdf_match.cu
    // Newton-Schulz epilogue (cov_match.cu): D = alpha * acc + diag * I, and atomicMax(resid_max, max |acc - I|)
    float diag;
    float *resid_max;
    const float *skip_below;  // device value: *skip_below < skip_tol -> the whole launch is a no-op
    float skip_tol;
    bool relu;  // max(., 0) after the bias (not with d_trans)
    // batch > 0: `batch` (<= 4) independent problems of the same shape in ONE launch (A K-major, B either major, no
    // split-K / blend / bias).  A, B, D above are ignored; the operands of all problems must lie in one arena each
    // at whole-matrix distances (the tensor maps span the arena).  resid_z / skip_z: per-problem resid_max /
    // skip_below (a skipped problem contributes no tiles).
    int batch;
    const float *A_z[4], *B_z[4];
    float *D_z[4];
    float *resid_z[4];
    const float *skip_z[4];
    // alpha / diag from device memory (coef[0], coef[1]) instead of the host values above; per problem in batch mode
    const float *coef;
    const float *coef_z[4];
    // a second problem with the same B, N, K in the same launch (K-major A only; forward rotation of P and S)
    const float *A2;
    int64_t M2, ldd2;
    float *D2;
};
// implicit-GEMM 3 x 3 convolution (vgg.cu): `padded` = reflection-padded NHWC input [b, h + 2, w + 2, cin], cin % 32 == 0;
// w_hi / w_lo = tf32 halves of the packed weights [cout, 9 * cin] (k = tap * cin + ci); D = NHWC output [b, h, w, ldd]
struct TcConv {
    const float *padded;
    int b, h, w, cin;
    const float *w_hi, *w_lo;
    int cout;
    const float *bias;
    bool relu;
    float *D;
    int64_t ldd;
};
int gemm_tc_conv(const TcConv &c, cudaStream_t st);   // OPTEX_OK, OPTEX_ENOTSUP or an error

// selects which of the two library-owned hi/lo scratch buffers the calling thread's GEMMs use (pipelined callers)
void gemm_tc_set_scratch_slot(int slot);
void gemm_tc_set_trace(unsigned long long *device_buf);  // debug: 64 x u64 clock stamps of CTA 0's first tile
// 3xTF32 with a pre-split B: while registered (thread-local), gemm_tc() calls whose B pointer equals `src` use the
// given hi / lo halves instead of splitting B again (src == nullptr clears it)
void gemm_tc_set_presplit(const float *src, const float *hi, const float *lo);
int gemm_tc_split_batch(const float *x, float *out, int count, int64_t n, cudaStream_t st);
int gemm_tc_split_and_fill(const float *x, float *hi, float *lo, int64_t n, uint32_t *fill, int64_t fill_n,
                           uint32_t fill_v, cudaStream_t st);
// OPTEX_OK, OPTEX_ENOTSUP (shape/alignment outside the TMA constraints) or an error
int gemm_tc(const TcGemm &g, cudaStream_t st);

// dst = X R  (transposed: dst[c, n], else dst[n, c]);  terms = 1 (TF32) or 3 (3xTF32 split)
int gemm_tc_rotate_forward(const float *X, const float *R, float *dst, int64_t n, int c, bool transposed,
                           int terms, cudaStream_t st, int c0 = 0, int nc = -1, uint32_t *colrange = nullptr);
int gemm_tc_rotate_forward2(const float *X1, int64_t n1, const float *X2, int64_t n2, const float *R, float *dst1,
                            float *dst2, int c, int terms, cudaStream_t st, uint32_t *colrange = nullptr);
// out[n, j] = sum_c M(n, c) R[j, c] (+ content blend);  M channel-major [c, n] or NHWC [n, c]
int gemm_tc_rotate_inverse(const float *M, bool m_channel_major, const float *R, float *out, int64_t n, int c,
                           const float *content, float strength, int terms, cudaStream_t st);

}  // namespace optex
