// fp32 FFMA GEMM tiles (SIMT) - the exact-fp32 arithmetic path of the rotation GEMMs
// (optex.py:170-171,175) and the general GEMM used by the covariance modes
// (histmatch.py:18-42).  Handles every shape/alignment (PCA'd C = 23, 85, 181 ...).
// The tensor-core path lives in gemm_tcgen05.cu; this file is what runs when
// OPTEX_GEMM_FP32 is selected or when a shape does not meet the TMA constraints.
#include "common.cuh"

namespace optex {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int LDS_ = BM + 4;  // row pitch of the smem tiles (keeps float4 alignment)

// Load an 8-float strip of a [rows x BK] operand tile into registers.
//   KMAJOR : element(r, k) = p[r*ld + k]  thread -> r = t%128, k = (t/128)*8 + i
//   !KMAJOR: element(r, k) = p[k*ld + r]  thread -> k = t/16,   r = (t%16)*8 + i
template <bool KMAJOR>
__device__ __forceinline__ void load_strip(const float *__restrict__ p, int64_t ld, int64_t r0,
                                           int64_t k0, int64_t R, int64_t K, bool vec_ok, int t,
                                           float (&v)[8]) {
    if (KMAJOR) {
        int64_t r = r0 + (t & 127);
        int64_t k = k0 + (t >> 7) * 8;
        if (r < R && vec_ok && k + 8 <= K) {
            const float4 *q = reinterpret_cast<const float4 *>(p + r * ld + k);
            float4 a = __ldg(q), b = __ldg(q + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (r < R && k + i < K) ? __ldg(p + r * ld + k + i) : 0.f;
        }
    } else {
        int64_t k = k0 + (t >> 4);
        int64_t r = r0 + (t & 15) * 8;
        if (k < K && vec_ok && r + 8 <= R) {
            const float4 *q = reinterpret_cast<const float4 *>(p + k * ld + r);
            float4 a = __ldg(q), b = __ldg(q + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = (k < K && r + i < R) ? __ldg(p + k * ld + r + i) : 0.f;
        }
    }
}

template <bool KMAJOR>
__device__ __forceinline__ void store_strip(float (*tile)[LDS_], int t, const float (&v)[8]) {
    if (KMAJOR) {
        int r = t & 127, k = (t >> 7) * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) tile[k + i][r] = v[i];
    } else {
        int k = t >> 4, r = (t & 15) * 8;
        *reinterpret_cast<float4 *>(&tile[k][r]) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(&tile[k][r + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

struct SimtExtra {
    const float *bias;
    int64_t bias_hw, bias_ld;
    int accumulate;
    int64_t k_per_z, d_z_stride;
    const int *skip;
    const float *skip_below;  // *skip_below < skip_tol -> no-op
    float skip_tol;
    int relu;
};

template <bool A_KMAJOR, bool B_KMAJOR, bool D_TRANS>
__global__ void __launch_bounds__(NT, 2)
sgemm_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ B, int64_t ldb,
             float *__restrict__ D, int64_t ldd, int64_t M, int64_t N, int64_t K,
             const float *__restrict__ blend, float strength, float alpha, int a_vec, int b_vec,
             int d_vec, SimtExtra ex) {
    if (ex.skip && *ex.skip) return;
    if (ex.skip_below && *ex.skip_below < ex.skip_tol) return;
    __shared__ __align__(16) float As[2][BK][LDS_];
    __shared__ __align__(16) float Bs[2][BK][LDS_];

    const int t = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int64_t n0 = (int64_t)blockIdx.y * BN;
    const int tx = t & 15, ty = t >> 4;
    const int mi = D_TRANS ? tx : ty;  // lanes run along the contiguous dim of D
    const int ni = D_TRANS ? ty : tx;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // split-K: blockIdx.z covers [kz0, kz1) and writes its partial product to D + z * d_z_stride
    const int64_t kz0 = ex.k_per_z > 0 ? (int64_t)blockIdx.z * ex.k_per_z : 0;
    const int64_t kz1 = ex.k_per_z > 0 ? (kz0 + ex.k_per_z < K ? kz0 + ex.k_per_z : K) : K;
    D += (int64_t)blockIdx.z * ex.d_z_stride;
    K = kz1;  // loaders bound-check against the end of this slice
    const int64_t nk = (kz1 - kz0 + BK - 1) / BK;
    float ra[8], rb[8];
    load_strip<A_KMAJOR>(A, lda, m0, kz0, M, K, a_vec, t, ra);
    load_strip<B_KMAJOR>(B, ldb, n0, kz0, N, K, b_vec, t, rb);
    store_strip<A_KMAJOR>(As[0], t, ra);
    store_strip<B_KMAJOR>(Bs[0], t, rb);
    __syncthreads();

    for (int64_t kt = 0; kt < nk; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < nk) {
            load_strip<A_KMAJOR>(A, lda, m0, kz0 + (kt + 1) * BK, M, K, a_vec, t, ra);
            load_strip<B_KMAJOR>(B, ldb, n0, kz0 + (kt + 1) * BK, N, K, b_vec, t, rb);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][k][mi * 4]);
            float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][k][64 + mi * 4]);
            float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][k][ni * 4]);
            float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][k][64 + ni * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_strip<A_KMAJOR>(As[cur ^ 1], t, ra);
            store_strip<B_KMAJOR>(Bs[cur ^ 1], t, rb);
        }
        __syncthreads();
    }

    // ---- epilogue
    if (!D_TRANS) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int64_t m = m0 + (i < 4 ? mi * 4 + i : 64 + mi * 4 + (i - 4));
            if (m >= M) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int64_t n = n0 + (h ? 64 + ni * 4 : ni * 4);
                float v[4];
                float *dp = D + m * ldd + n;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = alpha * acc[i][h * 4 + j];
                    if (n + j < N) {
                        if (ex.bias) v[j] = __fadd_rn(v[j], __ldg(ex.bias + (m / ex.bias_hw) * ex.bias_ld + n + j));
                        if (ex.accumulate) v[j] += dp[j];
                        if (ex.relu) v[j] = v[j] < 0.f ? 0.f : v[j];
                    }
                }
                if (d_vec && n + 4 <= N) {
                    if (blend) {
                        float4 c = __ldg(reinterpret_cast<const float4 *>(blend + m * ldd + n));
                        // optex.py:117  p += strength * (content - p), unfused like torch
                        v[0] = __fadd_rn(v[0], __fmul_rn(strength, __fsub_rn(c.x, v[0])));
                        v[1] = __fadd_rn(v[1], __fmul_rn(strength, __fsub_rn(c.y, v[1])));
                        v[2] = __fadd_rn(v[2], __fmul_rn(strength, __fsub_rn(c.z, v[2])));
                        v[3] = __fadd_rn(v[3], __fmul_rn(strength, __fsub_rn(c.w, v[3])));
                    }
                    *reinterpret_cast<float4 *>(dp) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < N) {
                            float x = v[j];
                            if (blend)
                                x = __fadd_rn(x, __fmul_rn(strength, __fsub_rn(__ldg(blend + m * ldd + n + j), x)));
                            dp[j] = x;
                        }
                }
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int64_t n = n0 + (j < 4 ? ni * 4 + j : 64 + ni * 4 + (j - 4));
            if (n >= N) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int64_t m = m0 + (h ? 64 + mi * 4 : mi * 4);
                float *dp = D + n * ldd + m;
                if (d_vec && m + 4 <= M) {
                    *reinterpret_cast<float4 *>(dp) =
                        make_float4(alpha * acc[h * 4 + 0][j], alpha * acc[h * 4 + 1][j],
                                    alpha * acc[h * 4 + 2][j], alpha * acc[h * 4 + 3][j]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (m + i < M) dp[i] = alpha * acc[h * 4 + i][j];
                }
            }
        }
    }
}

__global__ void transpose_kernel(const float *__restrict__ in, float *__restrict__ out, int64_t rows,
                                 int64_t cols) {
    __shared__ float tile[32][33];
    const int64_t tiles_c = (cols + 31) / 32;
    int64_t c0 = ((int64_t)blockIdx.x % tiles_c) * 32, r0 = ((int64_t)blockIdx.x / tiles_c) * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        int64_t r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][i];
    }
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int sgemm_simt(const float *A, int64_t lda, bool a_kmajor, const float *B, int64_t ldb,
               bool b_kmajor, float *D, int64_t ldd, bool d_trans, int64_t M, int64_t N,
               int64_t K, const float *blend, float strength, float alpha, cudaStream_t st) {
    SimtOpts o{};
    return sgemm_simt_ex(A, lda, a_kmajor, B, ldb, b_kmajor, D, ldd, d_trans, M, N, K, blend, strength, alpha, o, st);
}

int sgemm_simt_ex(const float *A, int64_t lda, bool a_kmajor, const float *B, int64_t ldb, bool b_kmajor, float *D,
                  int64_t ldd, bool d_trans, int64_t M, int64_t N, int64_t K, const float *blend, float strength,
                  float alpha, const SimtOpts &o, cudaStream_t st) {
    if (M <= 0 || N <= 0 || K <= 0) return OPTEX_OK;
    if ((blend || o.bias || o.accumulate) && d_trans) {
        set_error("sgemm_simt: blend / bias / accumulate epilogues need a non-transposed D");
        return OPTEX_EINVAL;
    }
    SimtExtra ex{};
    ex.bias = o.bias; ex.bias_hw = o.bias_hw > 0 ? o.bias_hw : 1; ex.bias_ld = o.bias_ld;
    ex.accumulate = o.accumulate ? 1 : 0; ex.skip = o.skip; ex.d_z_stride = o.d_z_stride;
    ex.skip_below = o.skip_below; ex.skip_tol = o.skip_tol; ex.relu = o.relu ? 1 : 0;
    int nz = 1;
    if (o.split_k > 1) {
        ex.k_per_z = ((K + o.split_k - 1) / o.split_k + BK - 1) / BK * BK;
        nz = (int)((K + ex.k_per_z - 1) / ex.k_per_z);
    }
    int a_vec = aligned16(A) && (lda % 4 == 0);
    int b_vec = aligned16(B) && (ldb % 4 == 0);
    int d_vec = aligned16(D) && (ldd % 4 == 0) && (!blend || aligned16(blend));
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN), (unsigned)nz);
    if (grid.y > 65535) {
        set_error("sgemm_simt: N too large for grid.y");
        return OPTEX_ESIZE;
    }
#define GO(AK, BK_, DT)                                                                        \
    sgemm_kernel<AK, BK_, DT><<<grid, NT, 0, st>>>(A, lda, B, ldb, D, ldd, M, N, K, blend,     \
                                                   strength, alpha, a_vec, b_vec, d_vec, ex)
    int sel = (a_kmajor ? 4 : 0) | (b_kmajor ? 2 : 0) | (d_trans ? 1 : 0);
    switch (sel) {
        case 0: GO(false, false, false); break;
        case 1: GO(false, false, true); break;
        case 2: GO(false, true, false); break;
        case 3: GO(false, true, true); break;
        case 4: GO(true, false, false); break;
        case 5: GO(true, false, true); break;
        case 6: GO(true, true, false); break;
        default: GO(true, true, true); break;
    }
#undef GO
    OPTEX_LAUNCH_CHECK("sgemm_kernel");
    return OPTEX_OK;
}

int transpose_f32(const float *in, float *out, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return OPTEX_OK;
    int64_t tiles = ((cols + 31) / 32) * ((rows + 31) / 32);
    if (tiles > 0x7fffffffLL) {
        set_error("transpose_f32: too many tiles");
        return OPTEX_ESIZE;
    }
    dim3 grid((unsigned)tiles);
    transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(in, out, rows, cols);
    OPTEX_LAUNCH_CHECK("transpose_kernel");
    return OPTEX_OK;
}

}  // namespace optex
