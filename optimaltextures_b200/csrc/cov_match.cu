// Closed-form Gaussian matching - the reference's covariance hist modes chol / pca / sym.
//
//   reference: hist_match()  histmatch.py:13-46  called on ROTATED features by
//   optimal_transport()  optex.py:167-177:
//       rp = P R, rs = S R;  m = T (rp - mu_rp) + mu_rs;  out = m R^T
//   with T built from the rotated covariances  Sig' = cov(rp) + eps I,  cov(rs) + eps I.
//
// B200 formulation.  Rotation by an orthogonal R commutes with taking moments:
//       mu_rp = R^T mu_p,      cov(rp) = R^T cov(P) R
// so the three N x C x C rotation GEMMs and the two rotated-data covariance GEMMs of the reference
// collapse into ONE Gram GEMM on the un-rotated pastiche (the style Gram is a second one), a handful of
// C x C products, and ONE N x C x C application GEMM with the means folded into a bias:
//       G = R T R^T,     out = P G^T + (mu_s - G mu_p)          (+ content blend, optex.py:117)
// The result is the reference's, up to fp32 rounding (tolerance stated in tests/test_gpu_cov.py).
// With R == nullptr the same code is hist_match() itself (histmatch.py:5-46, used un-rotated by
// mix_style_features, optex.py:200-201).
//
// C x C matrix functions, all as tensor-core GEMM chains (no eigendecomposition on the device):
//   pca / sym : Sig^(1/2) and Sig^(-1/2) by the coupled Newton-Schulz iteration
//               Y <- Y (3I - ZY)/2,  Z <- (3I - ZY)/2 Z     (Y0 = Sig/|Sig|_F, Z0 = I),
//               which replaces eigh + V sqrt(w) V^T (histmatch.py:30-33, 37-41); iterations that find the
//               residual |ZY - I|_max below 3e-4 switch the remaining launches off through a device flag.
//   chol      : blocked right-looking Cholesky (64-wide panels factored in shared memory, trailing update by
//               GEMM) and a warp-per-row triangular solve for  T = L_s L_t^-1  (histmatch.py:25-27).
#include <mutex>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace optex {
namespace {

constexpr int NS_MAX_ITERS = 24;
constexpr int NS_SCALED_ITERS = 16;  // launch cap of the scaled iteration (ns_prepare_kernel)
constexpr float NS_TOL = 3e-4f;
constexpr int MAX_C = 1024;
constexpr int PANEL = 64;
constexpr int B_MAX = 64;  // samples per batch (per-sample means, histmatch.py:16)

// ------------------------------------------------------------------ moments
// part[b][split][c] = sum over the split's rows of X[b*hw + r, c]
__global__ void colsum_partial_kernel(const float *__restrict__ X, float *__restrict__ part, int64_t hw, int c,
                                      int splits) {
    pdl_wait();
    __shared__ float red[8][33];
    const int ch = blockIdx.x * 32 + threadIdx.x;
    const int split = blockIdx.y, b = blockIdx.z;
    const int64_t rows = (hw + splits - 1) / splits;
    const int64_t r0 = split * rows, r1 = r0 + rows < hw ? r0 + rows : hw;
    float acc = 0.f;
    if (ch < c)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) acc += X[((int64_t)b * hw + r) * c + ch];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        part[((int64_t)b * splits + split) * c + ch] = s;
    }
}
// mu[b][c] = (sum over splits) / hw                       histmatch.py:16,20
__global__ void colmean_final_kernel(const float *__restrict__ part, float *__restrict__ mu, int64_t hw, int c,
                                     int splits, int nb) {
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb * c) return;
    int b = i / c, ch = i % c;
    // four independent accumulators: the loads of a walk over the splits overlap instead of queueing behind one sum
    const float *p = part + (int64_t)b * splits * c + ch;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 3 < splits; k += 4) {
        s0 += p[(int64_t)k * c];
        s1 += p[(int64_t)(k + 1) * c];
        s2 += p[(int64_t)(k + 2) * c];
        s3 += p[(int64_t)(k + 3) * c];
    }
    for (; k < splits; ++k) s0 += p[(int64_t)k * c];
    mu[i] = ((s0 + s1) + (s2 + s3)) / (float)hw;
}
// Xc[b*hw + r, ch] = X[b*hw + r, ch] - mu[b][ch]          histmatch.py:17,21  (x - mu before the product, like
// the reference: the second moment of un-centred data loses |mu|^2 / var digits to cancellation in fp32)
__global__ void center_kernel(const float *__restrict__ X, const float *__restrict__ mu, float *__restrict__ Xc,
                              int64_t hw, int c, int64_t total) {
    pdl_wait();
    if ((c & 3) == 0 && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Xc)) & 15) == 0) {
        // four channels per thread: one 128-bit load / store, one division per quad
        const int64_t quads = total >> 2;
        const int cq = c >> 2;
        const float4 *x4 = reinterpret_cast<const float4 *>(X);
        float4 *o4 = reinterpret_cast<float4 *>(Xc);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads;
             i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t row = i / cq;
            const int q = (int)(i - row * cq);
            const float4 m = *reinterpret_cast<const float4 *>(mu + (row / hw) * c + 4 * q);
            float4 v = x4[i];
            v.x -= m.x;
            v.y -= m.y;
            v.z -= m.z;
            v.w -= m.w;
            o4[i] = v;
        }
        return;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        const int64_t b = i / c / hw;
        Xc[i] = X[i] - mu[b * c + ch];
    }
}
// Sig = (sum_z part_z) / n  (+ eps on the diagonal); `part` is the Gram matrix of the CENTRED data, unless
// sub_mean: then - sum_b (hw/n) mu_b mu_b^T is applied here                         histmatch.py:17-18
__global__ void __launch_bounds__(256) gram_reduce_kernel(const float *__restrict__ part, int nz, int64_t zstride,
                                                          const float *__restrict__ mu, int nb, int64_t hw, int c,
                                                          float eps, float *__restrict__ Sig, int sub_mean) {
    pdl_wait();
    // 32 elements per block; 8 groups of threads take the slices z = g, g + 8, ... and fold through shared memory
    // (up to 256 slices: one thread per element would walk them as a dependent chain of L2 round trips)
    __shared__ float red[8][33];
    const int e = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + e;
    const bool ok = i < (int64_t)c * c;
    float s0 = 0.f, s1 = 0.f;
    if (ok) {
        int z = g;
        for (; z + 8 < nz; z += 16) {
            s0 += part[z * zstride + i];
            s1 += part[(z + 8) * zstride + i];
        }
        if (z < nz) s0 += part[z * zstride + i];
    }
    red[g][e] = s0 + s1;
    __syncthreads();
    if (g != 0 || !ok) return;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][e];
    const int r = (int)(i / c), q = (int)(i % c);
    const float n = (float)(hw * nb);
    float m2 = 0.f;
    if (sub_mean)
        for (int b = 0; b < nb; ++b) m2 = fmaf(mu[b * c + r], mu[b * c + q], m2);
    float v = s / n - m2 * ((float)hw / n);
    if (r == q) v += eps;
    Sig[i] = v;
}
__global__ void add_diag_kernel(float *A, int c, float eps) {
    pdl_wait();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c) A[(int64_t)i * c + i] += eps;
}
// bias[b][j] = mu_s[bs(b)][j] - sum_c G[j][c] mu_p[b][c]            (means folded through the map)
__global__ void bias_kernel(const float *__restrict__ G, const float *__restrict__ mu_p, const float *__restrict__ mu_s,
                            int b_s, int c, float *__restrict__ bias) {
    pdl_wait();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    if (warp >= c) return;
    float acc = 0.f;
    for (int k = lane; k < c; k += 32) acc = fmaf(G[(int64_t)warp * c + k], mu_p[b * c + k], acc);
    acc = warp_sum(acc);
    if (lane == 0) bias[b * c + warp] = mu_s[(b_s == 1 ? 0 : b) * c + warp] - acc;
}

// Inside optex_ot_loop the pastiche of iteration i + 1 is the output of iteration i, whose per-sample mean is known
// without a pass over it: out = (x - mu_p) G^T + mu_s has mean mu_s, and the content blend of optex.py:117 moves it to
// (1 - s) mu_s + s mean(content).  (The rounding of the stored outputs moves the true mean by ~1e-7 of the scale; the
// Gram of data centred by the analytic mean differs by n delta delta^T, 1e-14 of the variance.)
__global__ void mean_from_style_kernel(const float *__restrict__ mu_s, const float *__restrict__ mu_c, float strength,
                                       int b_p, int b_s, int c, float *__restrict__ mu_p) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b_p * c) return;
    const int b = i / c, ch = i - b * c;
    const float ms = mu_s[(b_s == 1 ? 0 : b) * c + ch];
    mu_p[i] = mu_c ? ms + strength * (mu_c[i] - ms) : ms;
}

// ------------------------------------------------------------------ Newton-Schulz helpers
// Coupled Newton-Schulz for A^(1/2), A^(-1/2).  Convergence is tracked on the device without flags: resid[it] is the
// max |Z Y - I| seen by iteration `it` (0 while the iteration has not run); iteration it is a no-op when
// resid[it-1] < NS_TOL - either iteration it-1 converged (its update is still applied) or it was itself skipped (its
// slot kept the initial 0) - and every kernel of an iteration tests that one value.  Y / Z ping-pong between two
// buffers, so the state after the last executed iteration k-1 sits in buffer k & 1.

// one block: norm2[0] = sum A^2 (deterministic), resid[..] = 0, and the coefficient schedule of the SCALED iteration.
//   With M_k = Z_k Y_k (eigenvalues m in [l_k, 1]; M_0 = A / |A|_F, so l_0 = lambda_min(A) / |A|_F >= lmin / |A|_F) the
//   step T_k = a_k I + b_k M_k, a_k = 1.5 sqrt(rho), b_k = -0.5 rho^1.5 is the plain step (rho = 1) applied to rho M_k;
//   rho = 3 / (1 + sqrt(l) + l) maps both ends of [l, 1] to the same value l' = rho l (3 - rho l)^2 / 4 and keeps the
//   maximum at 1: the small eigenvalues grow ~6x per iteration instead of 2.25x (8-9 real iterations instead of 13-16
//   on PCA'd VGG covariances), the fixed point Y = (A/s)^(1/2), Z = (A/s)^(-1/2) is unchanged.  lmin <= 0: rho = 1.
//   coefficients at norm2[8 + 2 it] = b_it (multiplies Z Y), norm2[9 + 2 it] = a_it (on the diagonal).
__global__ void ns_prepare_kernel(const float *__restrict__ A, int64_t n, float *__restrict__ norm2, float *resid,
                                  int iters, float lmin) {
    pdl_wait();
    __shared__ float red[32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc = fmaf(A[i], A[i], acc);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        acc = warp_sum(acc);
        if (threadIdx.x == 0) {
            norm2[0] = acc;
            double l = lmin > 0.f && acc > 0.f ? 0.9 * (double)lmin / sqrt((double)acc) : 1.0;   // 10 % under the bound
            if (l > 1.0) l = 1.0;
            for (int it = 0; it < iters; ++it) {
                const double rho = l < 0.8 ? 3.0 / (1.0 + sqrt(l) + l) : 1.0;   // near convergence: the plain step
                norm2[8 + 2 * it] = (float)(-0.5 * rho * sqrt(rho));
                norm2[9 + 2 * it] = (float)(1.5 * sqrt(rho));
                const double rl = rho * l;
                l = 0.97 * rl * (3.0 - rl) * (3.0 - rl) * 0.25;   // the new lower end, 3 % under it for the rounding
                if (l > 1.0) l = 1.0;
            }
        }
    }
    for (int i = threadIdx.x; i <= iters; i += blockDim.x) resid[i] = 0.f;
}
// Y = A / |A|_F, Z = I
__global__ void ns_init_kernel(const float *__restrict__ A, const float *__restrict__ norm2, float *__restrict__ Y,
                               float *__restrict__ Z, int c) {
    pdl_wait();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)c * c) return;
    const float inv = 1.f / sqrtf(norm2[0]);
    Y[i] = A[i] * inv;
    Z[i] = (i / c == i % c) ? 1.f : 0.f;
}
// T = 1.5 I - 0.5 T0 ;  resid[it] = max |T0 - I|   (only when the GEMM could not fuse it into its epilogue)
__global__ void ns_t_kernel(const float *__restrict__ T0, float *__restrict__ T, int c, float *resid, int it,
                            const float *__restrict__ coef) {
    pdl_wait();
    if (it > 0 && resid[it - 1] < NS_TOL) return;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float d = 0.f;
    if (i < (int64_t)c * c) {
        const float eye = (i / c == i % c) ? 1.f : 0.f;
        const float t0 = T0[i];
        T[i] = coef[1] * eye + coef[0] * t0;
        d = fabsf(t0 - eye);
        if (!(d == d)) d = INFINITY;
    }
    d = warp_max(d);
    if ((threadIdx.x & 31) == 0 && d > 0.f) atomicMax(reinterpret_cast<unsigned int *>(resid + it), __float_as_uint(d));
}
// Y = Y_k |A|_F^(1/2), Z = Z_k / |A|_F^(1/2), k = the first skipped iteration (buffer k & 1 holds the final state)
__global__ void ns_finish_kernel(const float *__restrict__ Y0, const float *__restrict__ Y1,
                                 const float *__restrict__ Z0, const float *__restrict__ Z1, float *__restrict__ Y,
                                 float *__restrict__ Z, const float *__restrict__ norm2,
                                 const float *__restrict__ resid, int iters, int c) {
    pdl_wait();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)c * c) return;
    int k = iters;
    for (int it = 1; it < iters; ++it)
        if (resid[it - 1] < NS_TOL) { k = it; break; }
    const float rs = sqrtf(sqrtf(norm2[0]));
    const float y = (k & 1) ? Y1[i] : Y0[i], z = (k & 1) ? Z1[i] : Z0[i];
    Y[i] = y * rs;
    Z[i] = z / rs;
}

// ------------------------------------------------------------------ Cholesky
// Block r of the panel starting at column j0: factor the diagonal block A[j0:j0+w, j0:j0+w] in shared memory
// (every block redundantly - 64^3/3 flops), then block 0 stores L_jj and block r > 0 solves and stores
// L_rj = A_rj L_jj^-T for its 64 rows below.
__global__ void __launch_bounds__(256)
potrf_panel_kernel(float *__restrict__ A, int c, int j0) {
    pdl_wait();
    __shared__ float Ljj[PANEL][PANEL + 1];
    __shared__ float Arow[PANEL][PANEL + 1];
    __shared__ __align__(16) float colk[PANEL];
    __shared__ float diagk;
    const int w = c - j0 < PANEL ? c - j0 : PANEL;
    const int tid = threadIdx.x;
    static_assert(PANEL == 64, "the register-resident factorisation below is written for 64 x 64 diagonal blocks");
    if (tid < PANEL) {
        // Thread r keeps row r of the diagonal block in registers (identity padding beyond w); per column k: the
        // pivot thread takes the square root, the rows below scale their entry and publish it, then every row
        // applies the rank-1 update to its own registers.  Two 64-thread barriers per column instead of three
        // block-wide ones around shared-memory updates: 89 -> ~12 us per panel.  Same operations in the same order
        // as the textbook right-looking loop, so the factor is bit-identical to it.
        const int r = tid;
        float a[PANEL];
#pragma unroll
        for (int q = 0; q < PANEL; ++q)
            a[q] = (r < w && q < w) ? A[(int64_t)(j0 + r) * c + j0 + q] : (r == q ? 1.f : 0.f);
#pragma unroll
        for (int k = 0; k < PANEL; ++k) {
            if (r == k) {
                a[k] = sqrtf(a[k]);
                diagk = a[k];
            }
            asm volatile("bar.sync 1, 64;" ::: "memory");
            if (r > k) {
                a[k] /= diagk;
                colk[r] = a[k];
            }
            asm volatile("bar.sync 1, 64;" ::: "memory");
            if (r > k) {
#pragma unroll
                for (int q = k + 1; q < PANEL; ++q)
                    if (q <= r) a[q] = fmaf(-a[k], colk[q], a[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < PANEL; ++q) Ljj[r][q] = q <= r ? a[q] : 0.f;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int i = tid; i < PANEL * PANEL; i += 256) {
            int r = i / PANEL, q = i % PANEL;
            if (r < w && q < w) A[(int64_t)(j0 + r) * c + j0 + q] = q <= r ? Ljj[r][q] : 0.f;
        }
        return;
    }
    const int r0 = j0 + PANEL + (blockIdx.x - 1) * PANEL;  // first row of this block (below the diagonal block)
    const int h = c - r0 < PANEL ? c - r0 : PANEL;
    for (int i = tid; i < PANEL * PANEL; i += 256) {
        int r = i / PANEL, q = i % PANEL;
        Arow[r][q] = (r < h && q < w) ? A[(int64_t)(r0 + r) * c + j0 + q] : 0.f;
    }
    __syncthreads();
    if (tid < h) {  // forward substitution along the row: x L_jj^T = a
        for (int k = 0; k < w; ++k) {
            float s = Arow[tid][k];
            for (int m = 0; m < k; ++m) s = fmaf(-Arow[tid][m], Ljj[k][m], s);
            Arow[tid][k] = s / Ljj[k][k];
        }
    }
    __syncthreads();
    for (int i = tid; i < PANEL * PANEL; i += 256) {
        int r = i / PANEL, q = i % PANEL;
        if (r < h && q < w) A[(int64_t)(r0 + r) * c + j0 + q] = Arow[r][q];
    }
}
// zero the strict upper triangle (the trailing updates touch the full square)
__global__ void tril_kernel(float *A, int c) {
    pdl_wait();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)c * c && i % c > i / c) A[i] = 0.f;
}
// T = Ls Lt^-1 : one warp per row solves  x Lt = ls  from the last column backwards; Ut = Lt^T row-major
template <int PL>
__global__ void __launch_bounds__(128)
trsm_right_kernel(const float *__restrict__ Ls, const float *__restrict__ Ut, float *__restrict__ T, int c) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= c) return;
    float x[PL];
#pragma unroll
    for (int r = 0; r < PL; ++r) x[r] = 0.f;
    // the row of U, the diagonal entry and the right-hand side of step j - 1 are fetched while step j computes (the
    // chain acc -> v -> x -> next acc is dependent, so an unprefetched L2 round trip per step was the whole cost)
    float un[PL], ucur[PL];
    const float *lsrow = Ls + (int64_t)row * c;
    float dn, rn;
    {
        const float *u = Ut + (int64_t)row * c;
#pragma unroll
        for (int r = 0; r < PL; ++r) un[r] = 0.f;  // k > row: nothing to the right of the diagonal of the last step
        dn = u[row];
        rn = lsrow[row];
    }
    for (int j = row; j >= 0; --j) {  // x[j] = 0 for j > row (lower triangular product)
        const float dcur = dn, rcur = rn;
#pragma unroll
        for (int r = 0; r < PL; ++r) ucur[r] = un[r];
        if (j > 0) {
            const float *u = Ut + (int64_t)(j - 1) * c;
#pragma unroll
            for (int r = 0; r < PL; ++r) {
                const int k = r * 32 + lane;
                un[r] = (k > j - 1 && k <= row) ? u[k] : 0.f;
            }
            dn = u[j - 1];
            rn = lsrow[j - 1];
        }
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < PL; ++r) acc = fmaf(x[r], ucur[r], acc);
        acc = warp_sum(acc);
        const float v = (rcur - acc) / dcur;
#pragma unroll
        for (int r = 0; r < PL; ++r)
            if (r == (j >> 5) && lane == (j & 31)) x[r] = v;
    }
#pragma unroll
    for (int r = 0; r < PL; ++r) {
        int k = r * 32 + lane;
        if (k < c) T[(int64_t)row * c + k] = x[r];
    }
}

// ------------------------------------------------------------------ GEMM front-end (tensor cores, SIMT fallback)
int g_terms() { return optex_get_gemm_mode() == OPTEX_GEMM_TF32 ? 1 : 3; }
bool g_want_tc() { return optex_get_gemm_mode() != OPTEX_GEMM_FP32; }

// D[c, c] = alpha * op(A) op(B), all row-major [c, c];  ta / tb : use the transpose
int mm(const float *A, bool ta, const float *B, bool tb, float *D, int c, float alpha, const int *skip,
       cudaStream_t st) {
    if (g_want_tc()) {
        TcGemm g{};
        g.A = A; g.a_mn = ta; g.B = B; g.b_mn = !tb; g.D = D; g.ldd = c;
        g.M = g.N = g.K = c; g.terms = g_terms(); g.alpha = alpha; g.skip = skip;
        int rc = gemm_tc(g, st);
        if (rc != OPTEX_ENOTSUP) return rc;
    }
    SimtOpts o;
    o.skip = skip;
    return sgemm_simt_ex(A, c, !ta, B, c, tb, D, c, false, c, c, c, nullptr, 0.f, alpha, o, st);
}

inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

struct Ws {
    float *mu_p, *mu_s, *mu_c, *bias, *part_mean, *part_gram;
    float *centred;  // max(n_t, n_s) x c: the block minus its per-(b, c) means, input of the Gram GEMM
    float *m[19];  // c x c matrices
    float *norm2, *resid;   // Newton-Schulz state of chain 0; chain 1 (side stream) at +NS_STATE
    int *flags;
    float *coop;            // scratch of the cooperative chain kernel (cov_chain.cu)
};
constexpr int NS_STATE = 3 * NS_MAX_ITERS + 8;  // floats / ints per chain (norm2: [0] = |A|_F^2, [8 + 2 it ..] = step coefficients)

// Newton-Schulz state of one chain
struct NsState {
    float *norm2, *resid;
    int *flags;
};
constexpr int kMeanSplits = 64;
constexpr int kGramSplits = 32;   // at least this many K slices may be used; narrow matrices take more (gram_split_cap)
// K slices of the Gram GEMM: a c x c output has few tiles (one at c <= 64 ... 32 at c = 512), so the slices are what
// fills the SMs - two CTAs' worth of slices per SM, 32 ... 256 of them, each at least 256 rows long
int gram_split_cap(int c) {
    const int tiles = ((c + 127) / 128) * ((c + 63) / 64);
    int cap = 2 * 148 / (tiles < 1 ? 1 : tiles);
    if (cap < kGramSplits) cap = kGramSplits;
    if (cap > 256) cap = 256;
    return cap;
}

size_t ws_layout(int64_t n_t, int64_t n_s, int c, int b_max, Ws *w, void *base, size_t cap, bool *ok) {
    Arena ar(base, cap);
    const size_t cc = (size_t)c * c;
    Ws l{};
    l.mu_p = ar.take<float>((size_t)b_max * c);
    l.mu_s = ar.take<float>((size_t)b_max * c);
    l.mu_c = ar.take<float>((size_t)b_max * c);
    l.bias = ar.take<float>((size_t)b_max * c);
    l.part_mean = ar.take<float>((size_t)b_max * kMeanSplits * c);
    l.part_gram = ar.take<float>((size_t)gram_split_cap(c) * cc);
    l.centred = ar.take<float>((size_t)(n_t > n_s ? n_t : n_s) * c);
    for (int i = 0; i < 19; ++i) l.m[i] = ar.take<float>(cc);
    l.norm2 = ar.take<float>(2 * NS_STATE);
    l.resid = ar.take<float>(2 * NS_STATE);
    l.flags = ar.take<int>(2 * NS_STATE);
    l.coop = ar.take<float>(cov_coop_scratch_floats());
    if (w) *w = l;
    if (ok) *ok = ar.ok();
    return ar.off;
}

// pixel-sharded step (optex_ot_step_sharded): while set, moments() sees only this rank's rows of a ONE-sample block -
// column sums and Gram are all-reduced and divided by the total row count
thread_local const ShardCtx *g_shard = nullptr;

// mu[b][c] and Sig = cov + (eps on the diagonal) of X [nb*hw, c]
int moments(const float *X, int nb, int64_t hw, int c, float eps, float *mu, float *Sig, Ws &w, cudaStream_t st,
            int64_t hw_total = 0, bool mu_known = false) {
    const int64_t n = (int64_t)nb * hw;
    const ShardComm *cm = (g_shard && hw_total > 0) ? g_shard->comm : nullptr;
    if (cm && nb != 1) {
        set_error("sharded covariance step: one sample per side (b = 1)");
        return OPTEX_EINVAL;
    }
    const int64_t hw_div = cm ? hw_total : hw;   // sharded: local sums over the TOTAL count, summed over the ranks
    if (!mu_known) {   // (known: the caller filled mu - mean_from_style_kernel)
        int splits = (int)(hw < kMeanSplits ? hw : kMeanSplits);
        launch_pdl(colsum_partial_kernel, dim3(cdiv(c, 32), splits, nb), dim3(32, 8), 0, st, X, w.part_mean, hw, c, splits);
        OPTEX_LAUNCH_CHECK("colsum_partial_kernel");
        launch_pdl(colmean_final_kernel, dim3((unsigned)(cdiv((int64_t)nb * c, 64))), dim3(64), 0, st, w.part_mean, mu, hw_div, c, splits, nb);
        OPTEX_LAUNCH_CHECK("colmean_final_kernel");
        if (cm) OPTEX_TRY(shard_allreduce_f32_sum(cm, mu, (size_t)c, st));
    }
    {
        const int64_t total = n * c;
        int64_t blocks = (total + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        launch_pdl(center_kernel, dim3((unsigned)blocks), dim3(256), 0, st, X, (const float *)mu, w.centred, hw, c, total);
        OPTEX_LAUNCH_CHECK("center_kernel");
        X = w.centred;
    }
    // Gram Xc^T Xc: A = Xc^T (stored [K = n, M = c]) and B = Xc (stored [K = n, N = c]), split over K
    const int zcap = gram_split_cap(c);
    int nz = (int)(n / 256 < 1 ? 1 : (n / 256 > zcap ? zcap : n / 256));
    const int64_t zstride = (int64_t)c * c;
    int rc = OPTEX_ENOTSUP;
    if (g_want_tc()) {
        TcGemm g{};
        g.A = X; g.a_mn = true; g.B = X; g.b_mn = true; g.D = w.part_gram; g.ldd = c;
        g.M = c; g.N = c; g.K = n; g.terms = g_terms(); g.alpha = 1.f; g.split_k = nz; g.d_z_stride = zstride;
        rc = gemm_tc(g, st);
        if (rc != OPTEX_OK && rc != OPTEX_ENOTSUP) return rc;
        if (rc == OPTEX_OK && nz > 1) {  // gemm_tc rounds the slice length up to 32: same count rule as below
            int64_t kz = ((n + nz - 1) / nz + 31) / 32 * 32;
            nz = (int)((n + kz - 1) / kz);
        }
    }
    if (rc == OPTEX_ENOTSUP) {
        SimtOpts o;
        o.split_k = nz;
        o.d_z_stride = zstride;
        OPTEX_TRY(sgemm_simt_ex(X, c, false, X, c, false, w.part_gram, c, false, c, c, n, nullptr, 0.f, 1.f, o, st));
        if (nz > 1) {
            int64_t kz = ((n + nz - 1) / nz + 15) / 16 * 16;
            nz = (int)((n + kz - 1) / kz);
        }
    }
    launch_pdl(gram_reduce_kernel, dim3((unsigned)(cdiv((int64_t)c * c, 32))), dim3(256), 0, st, w.part_gram, nz, zstride, mu, nb, hw_div, c, cm ? 0.f : eps, Sig, 0);
    OPTEX_LAUNCH_CHECK("gram_reduce_kernel");
    if (cm) {
        OPTEX_TRY(shard_allreduce_f32_sum(cm, Sig, (size_t)c * c, st));
        if (eps != 0.f) {
            launch_pdl(add_diag_kernel, dim3((unsigned)(cdiv(c, 256))), dim3(256), 0, st, Sig, c, eps);
            OPTEX_LAUNCH_CHECK("add_diag_kernel");
        }
    }
    return OPTEX_OK;
}

// T = 1.5 I - 0.5 Z Y and resid[it] = max |Z Y - I|: fused into the tensor-core GEMM's epilogue, or GEMM + ns_t_kernel
// OPTEX_NS_SCALED=0: the plain Newton-Schulz iteration (A/B comparisons)
bool g_ns_scaled() {
    static const char *env = getenv("OPTEX_NS_SCALED");
    return !(env && atoi(env) == 0);
}

int ns_t_step(const float *Z, const float *Y, float *t0, float *t, int c, float *resid, int it, const float *coef,
              cudaStream_t st) {
    const float *skip_below = it > 0 ? resid + it - 1 : nullptr;
    if (g_want_tc()) {
        TcGemm g{};
        g.A = Z; g.a_mn = false; g.B = Y; g.b_mn = true; g.D = t; g.ldd = c;
        g.M = g.N = g.K = c; g.terms = g_terms(); g.alpha = -0.5f; g.diag = 1.5f; g.resid_max = resid + it;
        g.coef = coef;   // the scaled step's (b, a) replace (-0.5, 1.5)
        g.skip_below = skip_below; g.skip_tol = NS_TOL;
        int rc = gemm_tc(g, st);
        if (rc != OPTEX_ENOTSUP) return rc;
    }
    SimtOpts o;
    o.skip_below = skip_below;
    o.skip_tol = NS_TOL;
    OPTEX_TRY(sgemm_simt_ex(Z, c, true, Y, c, false, t0, c, false, c, c, c, nullptr, 0.f, 1.f, o, st));
    launch_pdl(ns_t_kernel, dim3(cdiv((int64_t)c * c, 256)), dim3(256), 0, st, (const float *)t0, t, c, resid, it, coef);
    OPTEX_LAUNCH_CHECK("ns_t_kernel");
    return OPTEX_OK;
}

// D = A B (all [c, c] row-major), a no-op once the chain has converged
int ns_mm(const float *A, const float *B, float *D, int c, const float *skip_below, cudaStream_t st) {
    if (g_want_tc()) {
        TcGemm g{};
        g.A = A; g.a_mn = false; g.B = B; g.b_mn = true; g.D = D; g.ldd = c;
        g.M = g.N = g.K = c; g.terms = g_terms(); g.alpha = 1.f; g.skip_below = skip_below; g.skip_tol = NS_TOL;
        int rc = gemm_tc(g, st);
        if (rc != OPTEX_ENOTSUP) return rc;
    }
    SimtOpts o;
    o.skip_below = skip_below;
    o.skip_tol = NS_TOL;
    return sgemm_simt_ex(A, c, true, B, c, false, D, c, false, c, c, c, nullptr, 0.f, 1.f, o, st);
}

// One Newton-Schulz chain: Y = A^(1/2), Z = A^(-1/2) for SPD A; t0, t, yn, zn: scratch
struct NsChain {
    const float *A;
    float *Y, *Z, *t0, *t, *yn, *zn;
    NsState w;
    float lmin;   // a lower bound of A's smallest eigenvalue (eps for cov + eps I); <= 0: unknown -> unscaled iteration
};

// Up to two independent chains in lockstep.  Per iteration: ONE batched launch for the T steps of all chains and ONE
// for their 2 x nch update products (gemm_tc batch mode; a converged chain contributes no tiles), i.e. 2 launches
// instead of 3 per chain; shapes outside the tensor-core path fall back to per-chain launches.
int ns_sqrt_multi(const NsChain *ch, int nch, int c, cudaStream_t st) {
    const int64_t cc = (int64_t)c * c;
    const unsigned nb = cdiv(cc, 256);
    for (int k = 0; k < nch; ++k) {
        launch_pdl(ns_prepare_kernel, dim3(1), dim3(1024), 0, st, ch[k].A, cc, ch[k].w.norm2, ch[k].w.resid,
                   NS_MAX_ITERS, g_ns_scaled() ? ch[k].lmin : 0.f);
        OPTEX_LAUNCH_CHECK("ns_prepare_kernel");
        launch_pdl(ns_init_kernel, dim3(nb), dim3(256), 0, st, ch[k].A, (const float *)ch[k].w.norm2, ch[k].Y, ch[k].Z, c);
        OPTEX_LAUNCH_CHECK("ns_init_kernel");
    }
    bool batched = g_want_tc() && nch >= 1 && 2 * nch <= 4;
    static const char *cap_env = getenv("OPTEX_NS_CAP");   // debug: fewer host-side iterations (results may not converge)
    // the scaled iteration multiplies the lower end of the spectrum by ~6 per step: 16 steps cover |A|_F / lmin up to
    // 1e9 (measured: 9-10 real steps on PCA'd covariances); the plain one keeps the 24
    bool all_scaled = g_ns_scaled();
    for (int k = 0; k < nch; ++k) all_scaled = all_scaled && ch[k].lmin > 0.f;
    int max_it = all_scaled ? NS_SCALED_ITERS : NS_MAX_ITERS;
    if (cap_env && atoi(cap_env) > 0 && atoi(cap_env) < max_it) max_it = atoi(cap_env);
    for (int it = 0; it < max_it; ++it) {
        const int cur = it & 1, nxt = cur ^ 1;
        auto Yb = [&](int k, int i) { return i ? ch[k].yn : ch[k].Y; };
        auto Zb = [&](int k, int i) { return i ? ch[k].zn : ch[k].Z; };
        auto skip_of = [&](int k) -> const float * { return it > 0 ? ch[k].w.resid + it - 1 : nullptr; };
        bool done = false;
        if (batched) {
            TcGemm g{};
            g.batch = nch; g.b_mn = true; g.ldd = c; g.M = g.N = g.K = c; g.terms = g_terms();
            g.alpha = -0.5f; g.diag = 1.5f; g.skip_tol = NS_TOL;
            for (int k = 0; k < nch; ++k) {
                g.A_z[k] = Zb(k, cur); g.B_z[k] = Yb(k, cur); g.D_z[k] = ch[k].t;
                g.resid_z[k] = ch[k].w.resid + it; g.skip_z[k] = skip_of(k);
                g.coef_z[k] = ch[k].w.norm2 + 8 + 2 * it;
            }
            const int rc = gemm_tc(g, st);
            if (rc == OPTEX_OK) done = true;
            else if (rc != OPTEX_ENOTSUP) return rc;
            else batched = false;
        }
        if (!done)
            for (int k = 0; k < nch; ++k)
                OPTEX_TRY(ns_t_step(Zb(k, cur), Yb(k, cur), ch[k].t0, ch[k].t, c, ch[k].w.resid, it,
                                    ch[k].w.norm2 + 8 + 2 * it, st));
        done = false;
        if (batched) {
            TcGemm g{};
            g.batch = 2 * nch; g.b_mn = true; g.ldd = c; g.M = g.N = g.K = c; g.terms = g_terms();
            g.alpha = 1.f; g.skip_tol = NS_TOL;
            for (int k = 0; k < nch; ++k) {
                g.A_z[2 * k] = Yb(k, cur); g.B_z[2 * k] = ch[k].t; g.D_z[2 * k] = Yb(k, nxt);
                g.A_z[2 * k + 1] = ch[k].t; g.B_z[2 * k + 1] = Zb(k, cur); g.D_z[2 * k + 1] = Zb(k, nxt);
                g.skip_z[2 * k] = g.skip_z[2 * k + 1] = skip_of(k);
            }
            const int rc = gemm_tc(g, st);
            if (rc == OPTEX_OK) done = true;
            else if (rc != OPTEX_ENOTSUP) return rc;
            else batched = false;
        }
        if (!done)
            for (int k = 0; k < nch; ++k) {
                OPTEX_TRY(ns_mm(Yb(k, cur), ch[k].t, Yb(k, nxt), c, skip_of(k), st));
                OPTEX_TRY(ns_mm(ch[k].t, Zb(k, cur), Zb(k, nxt), c, skip_of(k), st));
            }
    }
    for (int k = 0; k < nch; ++k) {
        launch_pdl(ns_finish_kernel, dim3(nb), dim3(256), 0, st, (const float *)ch[k].Y, (const float *)ch[k].yn,
                   (const float *)ch[k].Z, (const float *)ch[k].zn, ch[k].Y, ch[k].Z, (const float *)ch[k].w.norm2,
                   (const float *)ch[k].w.resid, max_it, c);
        OPTEX_LAUNCH_CHECK("ns_finish_kernel");
    }
    return OPTEX_OK;
}

int ns_sqrt(const float *A, float *Y, float *Z, float *t0, float *t, float *yn, float *zn, int c, const NsState &w,
            float lmin, cudaStream_t st) {
    const NsChain ch{A, Y, Z, t0, t, yn, zn, w, lmin};
    return ns_sqrt_multi(&ch, 1, c, st);
}

// in-place lower Cholesky factor of SPD A [c, c]
int cholesky(float *A, int c, cudaStream_t st) {
    for (int j0 = 0; j0 < c; j0 += PANEL) {
        const int below = c - j0 - PANEL > 0 ? c - j0 - PANEL : 0;
        launch_pdl(potrf_panel_kernel, dim3((unsigned)(1 + cdiv(below, PANEL))), dim3(256), 0, st, A, c, j0);
        OPTEX_LAUNCH_CHECK("potrf_panel_kernel");
        if (below > 0) {  // A22 -= L21 L21^T  (full square; the upper triangle is discarded at the end)
            const float *L21 = A + (int64_t)(j0 + PANEL) * c + j0;
            float *A22 = A + (int64_t)(j0 + PANEL) * c + j0 + PANEL;
            SimtOpts o;
            o.accumulate = true;
            OPTEX_TRY(sgemm_simt_ex(L21, c, true, L21, c, true, A22, c, false, below, below, PANEL, nullptr, 0.f, -1.f,
                                    o, st));
        }
    }
    launch_pdl(tril_kernel, dim3((unsigned)(cdiv((int64_t)c * c, 256))), dim3(256), 0, st, A, c);
    OPTEX_LAUNCH_CHECK("tril_kernel");
    return OPTEX_OK;
}

int trsm_right(const float *Ls, const float *Ut, float *T, int c, cudaStream_t st) {
    const unsigned grid = cdiv(c, 4);
    const int pl = (c + 31) / 32;
    if (pl <= 2) launch_pdl(trsm_right_kernel<2>, grid, dim3(128), 0, st, Ls, Ut, T, c);
    else if (pl <= 4) launch_pdl(trsm_right_kernel<4>, grid, dim3(128), 0, st, Ls, Ut, T, c);
    else if (pl <= 8) launch_pdl(trsm_right_kernel<8>, grid, dim3(128), 0, st, Ls, Ut, T, c);
    else if (pl <= 16) launch_pdl(trsm_right_kernel<16>, grid, dim3(128), 0, st, Ls, Ut, T, c);
    else launch_pdl(trsm_right_kernel<32>, grid, dim3(128), 0, st, Ls, Ut, T, c);
    OPTEX_LAUNCH_CHECK("trsm_right_kernel");
    return OPTEX_OK;
}

// The pastiche-side and the style-side factorisations are independent, latency-bound chains of small launches:
// the style side runs on a library-owned side stream, forked from and joined back into the caller's stream with
// events (so the call stays asynchronous and ordered on the caller's stream, and remains graph-capturable).
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
// one side stream + event pair per calling host THREAD and device: two threads driving the library concurrently never
// re-record each other's fork / join events, and their side chains do not serialise behind one another
thread_local SideStream t_side[64];

int side_stream(SideStream **out) {
    int dev = 0;
    OPTEX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 63;
    SideStream &s = t_side[dev];
    if (!s.stream) {
        OPTEX_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        OPTEX_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        OPTEX_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    }
    *out = &s;
    return OPTEX_OK;
}

}  // namespace

void cov_set_shard(const ShardCtx *ctx) { g_shard = ctx; }

size_t cov_match_ws_bytes(int64_t n_t, int64_t n_s, int c, int mode) {
    (void)mode;
    if (c < 1) return 0;
    const size_t wide = ws_layout(n_t, n_s, c, B_MAX, nullptr, nullptr, 0, nullptr) + 256;
    const size_t narrow = c <= 64 ? cov_small_ws_bytes(n_t, n_s, c) : 0;
    return wide > narrow ? wide : narrow;
}

// out[n, c] = (X - mu_t) G^T + mu_s with G = R T R^T (R may be null = identity); see the file header
int cov_ot_step(const float *P, const float *S, const float *R, float *out, int b_p, int64_t hw_p, int b_s,
                int64_t hw_s, int c, int mode, float eps, const float *content, float strength, void *workspace,
                size_t workspace_bytes, cudaStream_t st, int style_reuse) {
    if (c > MAX_C) {
        set_error("covariance modes: c=%d > %d", c, MAX_C);
        return OPTEX_ESIZE;
    }
    const int b_max = B_MAX;
    if (b_p > b_max || b_s > b_max) {
        set_error("covariance modes: batch %d exceeds %d", b_p > b_s ? b_p : b_s, b_max);
        return OPTEX_ESIZE;
    }
    // pca / sym are built from matrix square roots, which commute with an orthogonal change of basis:
    // f(R^T S R) = R^T f(S) R, so G = R T(R^T S_t R, R^T S_s R) R^T = T(S_t, S_s) - the rotation cancels and the step
    // is computed in the un-rotated frame (6 C x C products fewer; same result up to rounding).  Cholesky factors do
    // not commute with a rotation, chol keeps it.
    if (mode != OPTEX_MODE_CHOL) R = nullptr;
    // style_reuse is a bit field: 1 = the style side of an earlier call is still in the workspace, 2 = the pastiche is
    // that call's output (its mean is known: mean_from_style_kernel), 4 = a loop follows (keep the content's mean)
    const int loop_flags = style_reuse;
    style_reuse = (loop_flags & 1) ? 1 : 0;
    const bool mean_known = (loop_flags & 3) == 3 && (b_s == 1 || b_s == b_p);
    const bool loop_follows = (loop_flags & 4) != 0;
    if (!g_shard && cov_small_supported(c, mode, b_p, b_s))
        return cov_small_step(P, S, R, out, b_p, hw_p, b_s, hw_s, c, mode, eps, content, strength, workspace,
                              workspace_bytes, st, loop_flags);
    Ws w;
    bool ok = false;
    const int64_t n_p = (int64_t)b_p * hw_p, n_s = (int64_t)b_s * hw_s;
    ws_layout(n_p, n_s, c, b_max, &w, workspace, workspace_bytes, &ok);
    if (!ok) {
        set_error("covariance modes: workspace %zu < %zu bytes", workspace_bytes, cov_match_ws_bytes(n_p, n_s, c, mode));
        return OPTEX_EWORKSPACE;
    }
    float *sig_t = w.m[0], *sig_s = w.m[1], *tmp = w.m[2], *T = w.m[3], *G = w.m[4];
    float *Y = w.m[5], *Z = w.m[6], *t0 = w.m[7], *t = w.m[8], *yn = w.m[9], *zn = w.m[10], *Y2 = w.m[11],
          *Z2 = w.m[12], *aux = w.m[13];
    float *tmp_b = w.m[14], *t0_b = w.m[15], *t_b = w.m[16], *yn_b = w.m[17], *zn_b = w.m[18];  // side-stream scratch
    const NsState ns_a{w.norm2, w.resid, w.flags};
    const NsState ns_b{w.norm2 + NS_STATE, w.resid + NS_STATE, w.flags + NS_STATE};
    // moments in the un-rotated frame; eps is added after the rotation like the reference (histmatch.py:18,22)
    // style_reuse (the iterations of optex_ot_loop after the first: same S, same workspace): the style moments - for
    // pca also the style square root Y2 - are still in the workspace.  With a rotation (chol) the un-rotated style
    // covariance lives in `aux`, because the factorisation overwrites sig_s.
    float *sig_s_src = R ? aux : sig_s;
    if (loop_follows && content && !g_shard) {   // mean(content) per sample, once per loop
        const int sp = (int)(hw_p < kMeanSplits ? hw_p : kMeanSplits);
        launch_pdl(colsum_partial_kernel, dim3(cdiv(c, 32), sp, b_p), dim3(32, 8), 0, st, content, w.part_mean, hw_p, c, sp);
        OPTEX_LAUNCH_CHECK("colsum_partial_kernel");
        launch_pdl(colmean_final_kernel, dim3((unsigned)(cdiv((int64_t)b_p * c, 64))), dim3(64), 0, st, w.part_mean, w.mu_c, hw_p, c, sp, b_p);
        OPTEX_LAUNCH_CHECK("colmean_final_kernel");
    }
    const bool mu_known = mean_known && !g_shard;
    if (mu_known) {
        launch_pdl(mean_from_style_kernel, dim3((unsigned)cdiv((int64_t)b_p * c, 128)), dim3(128), 0, st,
                   (const float *)w.mu_s, content ? (const float *)w.mu_c : (const float *)nullptr, strength, b_p, b_s, c, w.mu_p);
        OPTEX_LAUNCH_CHECK("mean_from_style_kernel");
    }
    OPTEX_TRY(moments(P, b_p, hw_p, c, R ? 0.f : eps, w.mu_p, sig_t, w, st, g_shard ? g_shard->hw_p_total : 0, mu_known));
    if (!style_reuse)
        OPTEX_TRY(moments(S, b_s, hw_s, c, R ? 0.f : eps, w.mu_s, sig_s_src, w, st, g_shard ? g_shard->hw_s_total : 0));
    const bool coop = !R && cov_coop_supported(c, mode);
    if (coop) {
        // pca / sym at the PCA'd layer widths: Newton-Schulz chain(s), closing products and bias in ONE cooperative kernel
        // (a cooperative blocked Cholesky + triangular solve for chol was built and measured: its sequential 32-wide
        // panel steps and row substitutions behind grid barriers came out no faster than the launch path - DESIGN 8)
        OPTEX_TRY(cov_coop_chain(w.m, c, mode, eps, style_reuse, w.mu_p, w.mu_s, b_p, b_s, w.bias, w.coop, st));
    } else if (mode == OPTEX_MODE_CHOL) {
        // ---- fork: the style-side chain (sandwich, factorisation) on the side stream
        SideStream *side;
        OPTEX_TRY(side_stream(&side));
        cudaStream_t sb = side->stream;
        // the moments above were launched with programmatic serialisation: fence before the fork event, so the side stream
        // cannot start on sig_s / mu_s while their producers are still draining (and likewise before the join)
        if (pdl_enabled()) OPTEX_TRY(stream_fence(st));
        OPTEX_CUDA(cudaEventRecord(side->fork, st));
        OPTEX_CUDA(cudaStreamWaitEvent(sb, side->fork, 0));
        int rc_b = OPTEX_OK;
        auto side_chain = [&]() -> int {
            if (R) {  // Sig_s' = R^T Sig_s R + eps I
                OPTEX_TRY(mm(sig_s_src, false, R, false, tmp_b, c, 1.f, nullptr, sb));
                OPTEX_TRY(mm(R, true, tmp_b, false, sig_s, c, 1.f, nullptr, sb));
                launch_pdl(add_diag_kernel, dim3((unsigned)(cdiv(c, 256))), dim3(256), 0, sb, sig_s, c, eps);
                OPTEX_LAUNCH_CHECK("add_diag_kernel");
            }
            if (mode == OPTEX_MODE_CHOL) return cholesky(sig_s, c, sb);
            return OPTEX_OK;  // pca: both square roots run as one batched chain below; sym: the second needs the first
        };
        rc_b = side_chain();
        // always join, also on an error above: the side stream must not be left forked inside a capture
        if (pdl_enabled() && rc_b == OPTEX_OK) rc_b = stream_fence(sb);
        cudaError_t join_err = cudaEventRecord(side->join, sb);
        // ---- the pastiche-side chain on the caller's stream
        auto main_chain = [&]() -> int {
            if (R) {  // Sig_t' = R^T Sig_t R + eps I
                OPTEX_TRY(mm(sig_t, false, R, false, tmp, c, 1.f, nullptr, st));
                OPTEX_TRY(mm(R, true, tmp, false, sig_t, c, 1.f, nullptr, st));
                launch_pdl(add_diag_kernel, dim3((unsigned)(cdiv(c, 256))), dim3(256), 0, st, sig_t, c, eps);
                OPTEX_LAUNCH_CHECK("add_diag_kernel");
            }
            if (mode == OPTEX_MODE_CHOL) return cholesky(sig_t, c, st);
            if (mode == OPTEX_MODE_PCA) return OPTEX_OK;
            return ns_sqrt(sig_t, Y, Z, t0, t, yn, zn, c, ns_a, eps, st);
        };
        const int rc_a = main_chain();
        if (join_err == cudaSuccess) join_err = cudaStreamWaitEvent(st, side->join, 0);
        OPTEX_CUDA(join_err);
        OPTEX_TRY(rc_b);
        OPTEX_TRY(rc_a);
    } else if (mode == OPTEX_MODE_SYM) {   // pca / sym: nothing to run beside the pastiche chain - no fork
        OPTEX_TRY(ns_sqrt(sig_t, Y, Z, t0, t, yn, zn, c, ns_a, eps, st));
    }
    if (coop) {
        // T and the bias are in place
    } else if (mode == OPTEX_MODE_CHOL) {  // T = L_s L_t^-1                      histmatch.py:25-27
        OPTEX_TRY(transpose_f32(sig_t, tmp, c, c, st));
        OPTEX_TRY(trsm_right(sig_s, tmp, T, c, st));
    } else if (mode == OPTEX_MODE_PCA) {  // T = Sig_s^(1/2) Sig_t^(-1/2)   histmatch.py:29-34
        const NsChain chains[2] = {{sig_t, Y, Z, t0, t, yn, zn, ns_a, eps}, {sig_s, Y2, Z2, t0_b, t_b, yn_b, zn_b, ns_b, eps}};
        OPTEX_TRY(ns_sqrt_multi(chains, style_reuse ? 1 : 2, c, st));   // reuse: Y2 = Sig_s^(1/2) is still there
        OPTEX_TRY(mm(Y2, false, Z, false, T, c, 1.f, nullptr, st));
    } else {  // sym: T = Qt^-1 (Qt Sig_s Qt)^(1/2) Qt^-1                   histmatch.py:36-42
        OPTEX_TRY(mm(Y, false, sig_s, false, tmp, c, 1.f, nullptr, st));
        OPTEX_TRY(mm(tmp, false, Y, false, aux, c, 1.f, nullptr, st));
        // Qt Sig_s Qt >= lambda_min(Qt)^2 lambda_min(Sig_s) >= eps * eps
        OPTEX_TRY(ns_sqrt(aux, Y2, Z2, t0, t, yn, zn, c, ns_a, eps * eps, st));
        OPTEX_TRY(mm(Z, false, Y2, false, tmp, c, 1.f, nullptr, st));
        OPTEX_TRY(mm(tmp, false, Z, false, T, c, 1.f, nullptr, st));
    }
    const float *Gp = T;
    if (R) {  // G = R T R^T
        OPTEX_TRY(mm(R, false, T, false, tmp, c, 1.f, nullptr, st));
        OPTEX_TRY(mm(tmp, false, R, true, G, c, 1.f, nullptr, st));
        Gp = G;
    }
    if (!coop) {
        launch_pdl(bias_kernel, dim3(cdiv((int64_t)c * 32, 128), b_p), dim3(128), 0, st, Gp, w.mu_p, w.mu_s, b_s, c, w.bias);
        OPTEX_LAUNCH_CHECK("bias_kernel");
    }
    // out[n, j] = sum_c P[n, c] G[j, c] + bias[b(n), j]   (+ content blend)
    int rc = OPTEX_ENOTSUP;
    if (g_want_tc()) {
        TcGemm g{};
        g.A = P; g.a_mn = false; g.B = Gp; g.b_mn = false; g.D = out; g.ldd = c;
        g.M = n_p; g.N = c; g.K = c; g.terms = g_terms(); g.alpha = 1.f;
        g.blend = content; g.strength = strength; g.bias = w.bias; g.bias_hw = hw_p; g.bias_ld = c;
        rc = gemm_tc(g, st);
        if (rc != OPTEX_ENOTSUP) return rc;
    }
    SimtOpts o;
    o.bias = w.bias;
    o.bias_hw = hw_p;
    o.bias_ld = c;
    return sgemm_simt_ex(P, c, true, Gp, c, true, out, c, false, n_p, c, c, content, strength, 1.f, o, st);
}

int cov_match_nhwc(const float *target, const float *source, float *out, int b_t, int64_t hw_t, int b_s,
                   int64_t hw_s, int c, int mode, float eps, void *workspace, size_t workspace_bytes,
                   cudaStream_t st) {
    return cov_ot_step(target, source, nullptr, out, b_t, hw_t, b_s, hw_s, c, mode, eps, nullptr, 0.f, workspace,
                       workspace_bytes, st, 0);
}

}  // namespace optex
