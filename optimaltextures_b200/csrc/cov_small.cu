// Covariance hist modes chol / pca / sym for NARROW blocks (c <= 64 channels) - hist_match() histmatch.py:13-46 inside
// optimal_transport() optex.py:167-177, the same algebra as cov_match.cu (rotation cancelled, moments of the un-rotated
// block, one application pass with the means folded into a bias), specialised for the shapes where cov_match.cu's
// launch chain is pure latency: a 512^2 synthesis spends 160 of its 493 OT iterations on conv1_1 (262 144 pixels x 23
// PCA'd channels), where the tensor-core path costs 46 launches and 490 us per iteration - 350 us of it in N x C
// passes whose 32-wide k blocks keep ~12 KB in flight per SM, 140 us in 30 launches of 32 x 32 matrix products.
// Here an iteration is FOUR launches:
//   small_colsum_kernel  column sums per sample, split over the rows of the block       (reads X once)
//   small_gram_kernel    Gram matrix of the CENTRED rows (x - mu on the fly, mu from the partial sums), fp32 FFMA
//                        register tiles, split over the rows; the 8 CTAs of a cluster fold their partials through
//                        distributed shared memory (deterministic)                       (reads X once)
//   small_chain_kernel   ONE CTA: Sig = cov + eps I, the coupled Newton-Schulz chain(s) with every matrix in shared
//                        memory and a real early exit on the residual, T, the bias
//   small_apply_kernel   out = X T^T + bias (+ content blend, optex.py:117), rows staged through shared memory
// fp32 FFMA throughout (more accurate than the 3xTF32 products of the wide path).  chol takes the same four launches
// (the rotation sandwiches, both Cholesky factorisations, the triangular solve and G = R T R^T in the chain kernel's
// shared memory); the pixel-sharded step stays on cov_match.cu.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace optex {
namespace {

constexpr int ST = 256;            // threads per CTA (all kernels)
constexpr int SLAB = 64;           // rows per shared-memory slab of the Gram kernel
constexpr int NS_CAP_S = 24;       // iteration cap of a chain (the scaled iteration needs 8-10)
constexpr float NS_TOL_S = 3e-4f;  // max |Z Y - I| at which a chain stops (cov_match.cu: NS_TOL)
constexpr int MAX_SUM_SPLITS = 1024;
constexpr int B_MAX_S = 64;

inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// part[(b * (splits / GC) + split / GC) * c + ch] = sum over the rows of the GC splits of a cluster of X[b * hw + r, ch]
// (grid.x = splits, a multiple of GC; the GC CTAs of a cluster fold their sums through distributed shared memory)
constexpr int GC = 8;
template <int CP>
__global__ void __cluster_dims__(GC, 1, 1) __launch_bounds__(ST)
    small_colsum_kernel(const float *__restrict__ X, float *__restrict__ part, int64_t hw, int c, int splits) {
    pdl_wait();
    constexpr int RP = ST / CP;
    __shared__ float red[RP][CP];
    __shared__ float tot[CP];
    const int ch = threadIdx.x % CP, lr = threadIdx.x / CP;
    const int split = blockIdx.x, b = blockIdx.y;
    const int64_t rows = (hw + splits - 1) / splits;
    const int64_t r0 = split * rows, r1 = r0 + rows < hw ? r0 + rows : hw;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (ch < c) {
        const float *x = X + (int64_t)b * hw * c + ch;
        int64_t r = r0 + lr;
        for (; r + 3 * RP < r1; r += 4 * RP) {
            a0 += x[r * c];
            a1 += x[(r + RP) * c];
            a2 += x[(r + 2 * RP) * c];
            a3 += x[(r + 3 * RP) * c];
        }
        for (; r < r1; r += RP) a0 += x[r * c];
    }
    red[lr][ch] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (lr == 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < RP; ++i) s += red[i][ch];
        tot[ch] = s;
    }
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    if (cluster.block_rank() == 0 && threadIdx.x < c) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < GC; ++r) s += cluster.map_shared_rank(tot, r)[threadIdx.x];
        part[((int64_t)b * (splits / GC) + split / GC) * c + threadIdx.x] = s;
    }
    cluster.sync();  // nobody leaves while its shared memory is still being read
}

// mu[ch] = (sum over the splits of part[(b * splits + k) * c + ch]) / hw for one sample b, by all 256 threads of the
// CTA into shared memory (histmatch.py:16,20).  Fixed summation order: every CTA that calls it gets the same bits.
template <int CP>
__device__ __forceinline__ void block_mean(const float *__restrict__ part, int b, int splits, int c, int64_t hw,
                                           float *smu, float (*scratch)[CP]) {
    constexpr int RP = ST / CP;
    const int ch = threadIdx.x % CP, lr = threadIdx.x / CP;
    if (splits == 0) {   // `part` IS the mean [nb][c] (a loop's later iterations: small_known_mean_kernel)
        if (lr == 0) smu[ch] = ch < c ? part[b * c + ch] : 0.f;
        __syncthreads();
        return;
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (ch < c) {
        const float *p = part + (int64_t)b * splits * c + ch;
        int k = lr;
        for (; k + 3 * RP < splits; k += 4 * RP) {
            a0 += p[(int64_t)k * c];
            a1 += p[(int64_t)(k + RP) * c];
            a2 += p[(int64_t)(k + 2 * RP) * c];
            a3 += p[(int64_t)(k + 3 * RP) * c];
        }
        for (; k < splits; k += RP) a0 += p[(int64_t)k * c];
    }
    scratch[lr][ch] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (lr == 0) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < RP; ++i) s += scratch[i][ch];
        smu[ch] = ch < c ? s / (float)hw : 0.f;
    }
    __syncthreads();
}

// mu[b][ch] from the column-sum partials (the content's mean, once per loop)
template <int CP>
__global__ void __launch_bounds__(ST) small_mean_kernel(const float *__restrict__ part, int splits, int c, int64_t hw,
                                                        float *__restrict__ mu) {
    pdl_wait();
    __shared__ float smu[CP];
    __shared__ float mscr[ST / CP][CP];
    block_mean<CP>(part, blockIdx.x, splits, c, hw, smu, mscr);
    if (threadIdx.x < c) mu[blockIdx.x * c + threadIdx.x] = smu[threadIdx.x];
}
// the pastiche mean of a loop's later iterations (cov_match.cu: mean_from_style_kernel)
__global__ void small_known_mean_kernel(const float *__restrict__ mu_s, const float *__restrict__ mu_c, float strength,
                                        int b_p, int b_s, int c, float *__restrict__ mu_p) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b_p * c) return;
    const int b = i / c, ch = i - b * c;
    const float ms = mu_s[(b_s == 1 ? 0 : b) * c + ch];
    mu_p[i] = mu_c ? ms + strength * (mu_c[i] - ms) : ms;
}

// part[cta][i][j] = sum over the CTA's rows of (x_i - mu_i)(x_j - mu_j)     histmatch.py:17-18 (x - mu BEFORE the product)
// CP = padded channel count (32 / 64); (CP / 4)^2 threads hold one CP x CP accumulator as 4 x 4 register tiles and
// 256 / (CP / 4)^2 such groups take the rows of a slab in turn.
// The GC CTAs of a cluster fold their accumulators through distributed shared memory (rank r sums slice r of all GC
// matrices in rank order), so the chain kernel reads ctas / GC partials instead of ctas.
template <int CP>
__global__ void __cluster_dims__(GC, 1, 1) __launch_bounds__(ST)
    small_gram_kernel(const float *__restrict__ X, const float *__restrict__ part_sum, int splits,
                      float *__restrict__ part, int64_t n, int64_t hw, int c, int64_t rows_per_cta) {
    pdl_wait();
    constexpr int TD = CP / 4, TI = TD * TD, NG = ST / TI;
    constexpr int LPT = SLAB * CP / ST;  // loads per thread and slab
    __shared__ __align__(16) float tile[SLAB][CP];
    __shared__ float smu[CP];
    __shared__ float mscr[ST / CP][CP];
    const int tid = threadIdx.x;
    const int g = tid / TI, t = tid % TI, ty = t / TD, tx = t % TD;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < n ? r0 + rows_per_cta : n;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lch = tid % CP, lrow = tid / CP;
    int64_t cur_b = -1;
    // a slab never straddles two samples' means: it is cut at the sample boundary
    auto slab_end = [&](int64_t a) {
        const int64_t b = a / hw;
        int64_t e = a + SLAB < r1 ? a + SLAB : r1;
        return e > (b + 1) * hw ? (b + 1) * hw : e;
    };
    float v[LPT];
    auto fetch = [&](int64_t a, int64_t e) {   // raw values; the mean is subtracted on the way into the tile
#pragma unroll
        for (int i = 0; i < LPT; ++i) {
            const int64_t r = a + lrow + i * (ST / CP);
            v[i] = (r < e && lch < c) ? X[r * c + lch] : 0.f;
        }
    };
    int64_t s0 = r0, s1 = s0 < r1 ? slab_end(s0) : s0;
    if (s0 < r1) fetch(s0, s1);
    while (s0 < r1) {
        const int64_t b = s0 / hw;
        if (b != cur_b) {
            block_mean<CP>(part_sum, (int)b, splits, c, hw, smu, mscr);
            cur_b = b;
        }
#pragma unroll
        for (int i = 0; i < LPT; ++i) {
            const int64_t r = s0 + lrow + i * (ST / CP);
            tile[lrow + i * (ST / CP)][lch] = (r < s1 && lch < c) ? v[i] - smu[lch] : 0.f;
        }
        __syncthreads();
        const int64_t n0 = s1, n1 = n0 < r1 ? slab_end(n0) : n0;
        if (n0 < r1) fetch(n0, n1);   // the next slab is in flight during the FMAs
        const int rows = (int)(s1 - s0);
#pragma unroll 4
        for (int k = g; k < rows; k += NG) {
            const float4 a = *reinterpret_cast<const float4 *>(&tile[k][4 * ty]);
            const float4 bq = *reinterpret_cast<const float4 *>(&tile[k][4 * tx]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();  // also orders the next slab's smu / tile writes behind this slab's reads
        s0 = n0;
        s1 = n1;
    }
    // fold the NG groups through shared memory (tile is free now; NG * CP * CP floats = SLAB * CP for both CP)
    float *red = &tile[0][0];
    static_assert(CP * CP <= SLAB * CP, "group reduction buffer");
    for (int gg = 0; gg < NG; ++gg) {
        if (g == gg) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float *p = &red[(4 * ty + i) * CP + 4 * tx + j];
                    *p = gg == 0 ? acc[i][j] : *p + acc[i][j];
                }
        }
        __syncthreads();
    }
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    const unsigned rank = cluster.block_rank();
    float *out = part + (int64_t)(blockIdx.x / GC) * c * c;
    constexpr int SLICE = CP * CP / GC;
    for (int e = tid; e < SLICE; e += ST) {
        const int idx = rank * SLICE + e, i = idx / CP, j = idx % CP;
        float sacc = 0.f;
#pragma unroll
        for (int r = 0; r < GC; ++r) sacc += cluster.map_shared_rank(red, r)[idx];
        if (i < c && j < c) out[i * c + j] = sacc;
    }
    cluster.sync();  // nobody leaves while its shared memory is still being read
}

// ---------------------------------------------------------------- the C x C part on one CTA
// C = alpha * A B + diag * I for CP x CP matrices in shared memory (row-major, ld = CP).  True products, no symmetry
// assumed: the operands of the chains are symmetric only up to rounding, and reading A^T for A feeds that asymmetry
// back as a NON-commuting perturbation, which the coupled iteration amplifies by the condition number (it diverged on
// the sym mode's second chain).  256 threads as a 16 x 16 grid of TM x TM register tiles; A is read four k at a time
// along its rows (all lanes of a half-warp share the address), B along its rows.
// resid (optional): max |A B - I| over the real c x c block.
template <int CP>
__device__ __forceinline__ void mm_smem(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C,
                                        float alpha, float diag, int c, float *resid) {
    constexpr int TM = CP / 16;
    const int ty = threadIdx.x / 16, tx = threadIdx.x % 16;
    float acc[TM][TM];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int k0 = 0; k0 < CP; k0 += 4) {
        float av[TM][4];
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const float4 a = *reinterpret_cast<const float4 *>(&A[(TM * ty + i) * CP + k0]);
            av[i][0] = a.x;
            av[i][1] = a.y;
            av[i][2] = a.z;
            av[i][3] = a.w;
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            float bv[TM];
#pragma unroll
            for (int j = 0; j < TM; ++j) bv[j] = B[(k0 + kk) * CP + TM * tx + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(av[i][kk], bv[j], acc[i][j]);
        }
    }
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) {
            const int r = TM * ty + i, q = TM * tx + j;
            const float eye = r == q ? 1.f : 0.f;
            if (resid && r < c && q < c) {
                float e = fabsf(acc[i][j] - eye);
                if (!(e == e)) e = INFINITY;
                d = fmaxf(d, e);
            }
            C[r * CP + q] = fmaf(alpha, acc[i][j], diag * eye);
        }
    if (resid) {
        d = warp_max(d);
        if ((threadIdx.x & 31) == 0 && d > 0.f)
            atomicMax(reinterpret_cast<unsigned int *>(resid), __float_as_uint(d));  // non-negative floats
    }
}

// The coupled Newton-Schulz iteration of cov_match.cu (same scaled step schedule, same stopping rule) with everything
// in shared memory:  Y -> A^(1/2), Z -> A^(-1/2) for the SPD matrix in `A` (overwritten).  Buffers: A/Y, Z, T, Yn, Zn.
// On return *Yout / *Zout point at the buffers holding the results.  `scr`: >= 16 floats of shared scratch.
template <int CP>
__device__ void ns_chain_smem(float *Y, float *Z, float *T, float *Yn, float *Zn, int c, float lmin, float *scr,
                              float **Yout, float **Zout) {
    const int tid = threadIdx.x;
    float *sh = scr;  // [0] = |A|_F^2, [1] = residual of the current iteration
    // s^2 with s >= |A|_2: |A|_F^2 over the whole padded matrix (the padding is eps on the diagonal: part of the
    // spectrum being iterated) ...
    float acc = 0.f;
    for (int i = tid; i < CP * CP; i += ST) acc = fmaf(Y[i], Y[i], acc);
    acc = warp_sum(acc);
    __shared__ float wred[ST / 32];
    if ((tid & 31) == 0) wred[tid >> 5] = acc;
    __syncthreads();
    // the other bound of |A|_2: the largest absolute column sum (thread t walks column t: conflict-free); the smaller
    // of the two scales the iteration (cov_chain.cu: norm_partial)
    __shared__ float cmax[ST / 32];
    float cs = 0.f;
    if (tid < CP)
        for (int k = 0; k < CP; ++k) cs += fabsf(Y[k * CP + tid]);
    cs = warp_max(cs);
    if ((tid & 31) == 0) cmax[tid >> 5] = cs;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f, m = 0.f;
        for (int i = 0; i < ST / 32; ++i) {
            s += wred[i];
            m = fmaxf(m, cmax[i]);
        }
        sh[0] = (m > 0.f && m * m < s) ? m * m : s;
    }
    __syncthreads();
    const float inv = rsqrtf(sh[0]);
    for (int i = tid; i < CP * CP; i += ST) {
        Y[i] *= inv;
        Z[i] = (i / CP == i % CP) ? 1.f : 0.f;
    }
    __syncthreads();
    // the scaled step schedule of cov_match.cu's ns_prepare_kernel, one step per iteration in every thread's registers
    // (fp32: the schedule is a bound with 3-10 % margins, not a quantity the result depends on)
    float l = lmin > 0.f && sh[0] > 0.f ? fminf(0.9f * lmin * inv, 1.f) : 1.f;
    float prev = 1.f;
    bool prev_plain = false;
    for (int it = 0; it < NS_CAP_S; ++it) {
        const float rho = l < 0.8f ? 3.f / (1.f + sqrtf(l) + l) : 1.f;   // near convergence: the plain step
        const float sr = sqrtf(rho), rl = rho * l;
        l = fminf(0.97f * rl * (3.f - rl) * (3.f - rl) * 0.25f, 1.f);     // the new lower end, 3 % under it
        if (tid == 0) sh[1] = 0.f;
        __syncthreads();
        mm_smem<CP>(Z, Y, T, -0.5f * rho * sr, 1.5f * sr, c, &sh[1]);  // T = a I + b Z Y
        __syncthreads();
        const float r = sh[1];
        // under PLAIN steps the residual falls monotonically until it meets the rounding floor of the problem
        // (covariances with |Sig|_F / eps beyond ~1e6 in fp32): an iteration that no longer improves is not applied.
        // (A scaled step on a spectrum that is already tighter than the schedule's bound overshoots - the residual may
        // rise once, legitimately - so the test only looks at residuals produced by plain steps.)
        if (prev_plain && prev < 0.1f && !(r < prev)) break;
        prev = r;
        prev_plain = rho == 1.f;
        mm_smem<CP>(Y, T, Yn, 1.f, 0.f, c, nullptr);
        mm_smem<CP>(T, Z, Zn, 1.f, 0.f, c, nullptr);
        __syncthreads();
        float *tmp = Y;
        Y = Yn;
        Yn = tmp;
        tmp = Z;
        Z = Zn;
        Zn = tmp;
        if (r < NS_TOL_S) break;  // this iteration's update is applied, the next is not needed
    }
    __syncthreads();
    const float rs = sqrtf(sqrtf(sh[0]));
    for (int i = tid; i < CP * CP; i += ST) {
        Y[i] *= rs;
        Z[i] /= rs;
    }
    __syncthreads();
    *Yout = Y;
    *Zout = Z;
}

// Sig[CP x CP] (shared) = (sum of the Gram partials) / n + eps I; padding: eps (or 1) on the diagonal, 0 elsewhere
template <int CP>
__device__ void load_sig(const float *__restrict__ part, int nz, int c, float n, float eps, float *Sig) {
    for (int i = threadIdx.x; i < CP * CP; i += ST) {
        const int r = i / CP, q = i % CP;
        float v = 0.f;
        if (r < c && q < c) {
            const float *p = part + r * c + q;
            const int64_t zs = (int64_t)c * c;
            float sum = 0.f;
            for (int z0 = 0; z0 < nz; z0 += 8) {   // eight independent loads in flight, a fixed association
                float t[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) t[i] = z0 + i < nz ? p[(z0 + i) * zs] : 0.f;
                sum += ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
            }
            v = sum / n;
            if (r == q) v += eps;
        } else if (r == q) {
            v = eps > 0.f ? eps : 1.f;
        }
        Sig[i] = v;
    }
}

// In-place lower Cholesky factor of the SPD matrix A (CP x CP, shared memory); the strict upper triangle is zeroed.
// Right-looking, one column per step: the CTA's 256 threads share the rank-1 update of the trailing block.
template <int CP>
__device__ void chol_smem(float *A) {
    const int tid = threadIdx.x;
    for (int k = 0; k < CP; ++k) {
        const float d = sqrtf(A[k * CP + k]);   // every thread reads the same value (broadcast)
        __syncthreads();
        if (tid == 0) A[k * CP + k] = d;
        const float inv = 1.f / d;
        for (int r = k + 1 + tid; r < CP; r += ST) A[r * CP + k] *= inv;
        __syncthreads();
        const int m = CP - k - 1;               // trailing block: rows / columns k + 1 .. CP - 1, lower triangle
        for (int i = tid; i < m * m; i += ST) {
            const int r = k + 1 + i / m, q = k + 1 + i % m;
            if (q <= r) A[r * CP + q] = fmaf(-A[r * CP + k], A[q * CP + k], A[r * CP + q]);
        }
        __syncthreads();
    }
    for (int i = tid; i < CP * CP; i += ST)
        if (i % CP > i / CP) A[i] = 0.f;
    __syncthreads();
}

// T = Ls Lt^-1 for lower-triangular Ls, Lt (histmatch.py:25-27): row r of T solves  t Lt = ls_r  from the last column
// backwards; one thread per row, T stored TRANSPOSED in Tt while it is built (thread r owns column r: conflict-free),
// every thread reads the same Lt element at the same step (broadcast).  T is lower triangular.
template <int CP>
__device__ void trsm_smem(const float *Ls, const float *Lt, float *Tt, float *T) {
    const int r = threadIdx.x;
    if (r < CP) {
        for (int j = CP - 1; j >= 0; --j) {
            float acc = 0.f;
            if (j <= r) {
                acc = Ls[r * CP + j];
                for (int k = j + 1; k <= r; ++k) acc = fmaf(-Tt[k * CP + r], Lt[k * CP + j], acc);
                acc /= Lt[j * CP + j];
            }
            Tt[j * CP + r] = acc;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CP * CP; i += ST) T[i] = Tt[(i % CP) * CP + i / CP];
    __syncthreads();
}

// mode pca:  T = Sig_s^(1/2) Sig_t^(-1/2)                               histmatch.py:29-34
// mode sym:  T = Qt^-1 (Qt Sig_s Qt)^(1/2) Qt^-1,  Qt = Sig_t^(1/2)     histmatch.py:36-42
// mode chol: T = L_s L_t^-1 with L = chol(R^T Sig R + eps I) in the ROTATED frame (Cholesky factors do not commute
//            with a rotation), G = R T R^T                                histmatch.py:24-27, optex.py:167-177
// style_state: 0 = compute the style side from part_s (and store it: sig_s for sym / chol, Y2 = Sig_s^(1/2) for pca),
//              1 = reuse what an earlier call stored (optex_ot_loop: same S in every iteration)
template <int CP>
__global__ void __launch_bounds__(ST) small_chain_kernel(const float *__restrict__ part_t, int nz_t, float n_t,
                                                         const float *__restrict__ part_s, int nz_s, float n_s,
                                                         float *__restrict__ style_keep, int style_state, int mode,
                                                         float eps, int c, const float *__restrict__ sum_p,
                                                         int splits_p, int64_t hw_p, const float *__restrict__ sum_s,
                                                         int splits_s, int64_t hw_s, float *__restrict__ mu_s,
                                                         int b_p, int b_s, float *__restrict__ G,
                                                         float *__restrict__ bias, const float *__restrict__ R,
                                                         float *__restrict__ sig_keep, int sig_mode) {
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    constexpr int MM = CP * CP;
    float *m0 = sm, *m1 = sm + MM, *m2 = sm + 2 * MM, *m3 = sm + 3 * MM, *m4 = sm + 4 * MM, *m5 = sm + 5 * MM,
          *m6 = sm + 6 * MM, *m7 = sm + 7 * MM, *m8 = sm + 8 * MM;
    float *scr = sm + 9 * MM;
    const int tid = threadIdx.x;
    // sig_mode (inside optex_ot_loop, no content blend): 1 = the pastiche covariance is measured (Gram partials) and the
    // covariance of THIS call's output, Sig' = G (Sig - eps I) G^T + eps I, is left in sig_keep; 2 = Sig is taken from
    // sig_keep (no pass over the pastiche at all) and propagated again.  0 = measured, nothing kept.
    // (scripts/cov_propagation_sim.py: 40 propagated iterations end 1e-6 of the scale from the float64 loop; the loop
    // re-measures every 8th iteration.)
    auto load_sig_t = [&](float eps_in, float *dst) {
        if (sig_mode == 2) {
            for (int i = tid; i < MM; i += ST) dst[i] = sig_keep[i];
        } else {
            load_sig<CP>(part_t, nz_t, c, n_t, eps_in, dst);
        }
        __syncthreads();
        if (sig_mode)
            for (int i = tid; i < MM; i += ST) m8[i] = dst[i];
    };
    float eps_loaded = eps;
    float *Yt, *Zt, *Ys, *Zs;
    const float *Tres;
    if (mode == OPTEX_MODE_PCA) {
        // style side: Y2 = Sig_s^(1/2) into m5
        if (style_state == 0) {
            load_sig<CP>(part_s, nz_s, c, n_s, eps, m0);
            __syncthreads();
            ns_chain_smem<CP>(m0, m1, m2, m3, m4, c, eps, scr, &Ys, &Zs);
            for (int i = tid; i < MM; i += ST) {
                m5[i] = Ys[i];
                style_keep[i] = Ys[i];
            }
        } else {
            for (int i = tid; i < MM; i += ST) m5[i] = style_keep[i];
        }
        __syncthreads();
        load_sig_t(eps, m0);
        __syncthreads();
        ns_chain_smem<CP>(m0, m1, m2, m3, m4, c, eps, scr, &Yt, &Zt);
        mm_smem<CP>(m5, Zt, m6, 1.f, 0.f, c, nullptr);  // T = Y2 Z
        Tres = m6;
    } else if (mode == OPTEX_MODE_CHOL) {
        // moments without eps when a rotation follows (eps is added in the rotated frame, histmatch.py:18,22)
        const float eps_in = R ? 0.f : eps;
        eps_loaded = eps_in;
        if (style_state == 0) {
            load_sig<CP>(part_s, nz_s, c, n_s, eps_in, m1);
            __syncthreads();
            for (int i = tid; i < MM; i += ST) style_keep[i] = m1[i];
        } else {
            for (int i = tid; i < MM; i += ST) m1[i] = style_keep[i];
        }
        load_sig_t(eps_in, m0);
        if (R) {
            // m2 = R padded with the identity, m3 = its transpose
            for (int i = tid; i < MM; i += ST) {
                const int r = i / CP, q = i % CP;
                const float v = (r < c && q < c) ? R[r * c + q] : (r == q ? 1.f : 0.f);
                m2[i] = v;
                m3[q * CP + r] = v;
            }
            __syncthreads();
            mm_smem<CP>(m0, m2, m4, 1.f, 0.f, c, nullptr);   // Sig_t R
            mm_smem<CP>(m1, m2, m5, 1.f, 0.f, c, nullptr);   // Sig_s R
            __syncthreads();
            mm_smem<CP>(m3, m4, m0, 1.f, eps, c, nullptr);   // R^T Sig_t R + eps I
            mm_smem<CP>(m3, m5, m1, 1.f, eps, c, nullptr);   // R^T Sig_s R + eps I
        }
        __syncthreads();
        chol_smem<CP>(m0);                                   // L_t
        chol_smem<CP>(m1);                                   // L_s
        trsm_smem<CP>(m1, m0, m4, m5);                       // m5 = T = L_s L_t^-1
        if (R) {
            mm_smem<CP>(m5, m3, m4, 1.f, 0.f, c, nullptr);   // T R^T
            __syncthreads();
            mm_smem<CP>(m2, m4, m6, 1.f, 0.f, c, nullptr);   // G = R T R^T
            Tres = m6;
        } else {
            Tres = m5;
        }
    } else {
        if (style_state == 0) {
            load_sig<CP>(part_s, nz_s, c, n_s, eps, m5);
            __syncthreads();
            for (int i = tid; i < MM; i += ST) style_keep[i] = m5[i];
        } else {
            for (int i = tid; i < MM; i += ST) m5[i] = style_keep[i];
        }
        load_sig_t(eps, m0);
        __syncthreads();
        ns_chain_smem<CP>(m0, m1, m2, m3, m4, c, eps, scr, &Yt, &Zt);  // Qt, Qt^-1 somewhere in m0..m4
        for (int i = tid; i < MM; i += ST) {
            m7[i] = Yt[i];
            m6[i] = Zt[i];
        }
        __syncthreads();
        mm_smem<CP>(m5, m7, m0, 1.f, 0.f, c, nullptr);  // m0 = Sig_s Qt
        __syncthreads();
        mm_smem<CP>(m7, m0, m1, 1.f, 0.f, c, nullptr);  // m1 = Qt Sig_s Qt
        __syncthreads();
        // Qt Sig_s Qt >= lambda_min(Qt)^2 lambda_min(Sig_s) >= eps * eps
        ns_chain_smem<CP>(m1, m0, m2, m3, m4, c, eps * eps, scr, &Ys, &Zs);
        mm_smem<CP>(Ys, m6, m5, 1.f, 0.f, c, nullptr);  // m5 = (Qt Sig_s Qt)^(1/2) Qt^-1
        __syncthreads();
        mm_smem<CP>(m6, m5, m7, 1.f, 0.f, c, nullptr);  // T = Qt^-1 m5
        Tres = m7;
    }
    __syncthreads();
    if (sig_mode) {   // Sig' = G (Sig - eps I) G^T + eps I with G = Tres (in m5 / m6 / m7); m0 .. m2 are free now
        for (int i = tid; i < MM; i += ST) {
            const int r = i / CP, q = i % CP;
            m0[q * CP + r] = Tres[i];                     // G^T
            if (r == q) m8[i] -= eps_loaded;              // Sig - eps I (padding: eps - eps = 0, or 1 when eps = 0)
        }
        __syncthreads();
        mm_smem<CP>(Tres, m8, m1, 1.f, 0.f, c, nullptr);  // G (Sig - eps I)
        __syncthreads();
        mm_smem<CP>(m1, m0, m2, 1.f, eps_loaded, c, nullptr);
        __syncthreads();
        for (int i = tid; i < MM; i += ST) sig_keep[i] = m2[i];
    }
    // G[j][k] (real c x c) and bias[b][j] = mu_s[bs(b)][j] - sum_k G[j][k] mu_p[b][k]
    for (int i = tid; i < c * c; i += ST) G[i] = Tres[(i / c) * CP + (i % c)];
    // the means, from the column-sum partials (the style's are kept in mu_s for the following iterations of a loop)
    float *smu = scr;                                    // the chains are done with their scratch
    __shared__ float mscr[ST / CP][CP];
    if (style_state == 0)
        for (int b = 0; b < b_s; ++b) {
            block_mean<CP>(sum_s, b, splits_s, c, hw_s, smu, mscr);
            if (tid < c) mu_s[b * c + tid] = smu[tid];
            __syncthreads();
        }
    __threadfence_block();
    for (int b = 0; b < b_p; ++b) {
        block_mean<CP>(sum_p, b, splits_p, c, hw_p, smu, mscr);
        if (tid < c) {
            float a = 0.f;
            for (int k = 0; k < c; ++k) a = fmaf(Tres[tid * CP + k], smu[k], a);
            bias[b * c + tid] = mu_s[(b_s == 1 ? 0 : b) * c + tid] - a;
        }
        __syncthreads();
    }
}

// out[r][j] = sum_k X[r][k] G[j][k] + bias[b(r)][j]   (+ content blend: out += strength (content - out), optex.py:117)
// A warp takes 32 * RPL rows at a time: coalesced load into its shared-memory tile, RPL rows per lane in registers (every
// broadcast load of G feeds RPL rows), results written back to the tile as 128-bit stores and stored coalesced (with
// the blend).
template <int CP, int RPL>
__global__ void __launch_bounds__(ST) small_apply_kernel(const float *__restrict__ X, const float *__restrict__ G,
                                                         const float *__restrict__ bias, float *__restrict__ out,
                                                         int64_t n, int64_t hw, int c,
                                                         const float *__restrict__ content, float strength) {
    pdl_wait();
    constexpr int LDT = CP + 4;  // row stride of a warp's tile: conflict-free 128-bit row accesses
    constexpr int RW = 32 * RPL;  // rows per warp and step
    extern __shared__ __align__(16) float sm[];
    float *sG = sm;                       // [CP][CP]  G[j][k], zero padded
    float *tiles = sm + CP * CP;          // [8 warps][RW][LDT]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < CP * CP; i += ST) {
        const int j = i / CP, k = i % CP;
        sG[i] = (j < c && k < c) ? G[j * c + k] : 0.f;
    }
    __syncthreads();
    float *tile = tiles + warp * RW * LDT;
    const int64_t groups = (n + RW - 1) / RW;
    for (int64_t gi = (int64_t)blockIdx.x * (ST / 32) + warp; gi < groups; gi += (int64_t)gridDim.x * (ST / 32)) {
        const int64_t r0 = gi * RW;
        const int rows = (int)(n - r0 < RW ? n - r0 : RW);
        // coalesced load of rows x c floats (contiguous in memory) into the padded tile
        const float *src = X + r0 * c;
        const int total = rows * c;
        if (c == CP) {
            for (int i = lane; i < total; i += 32) tile[(i / CP) * LDT + (i % CP)] = src[i];
        } else {
            for (int i = lane; i < total; i += 32) tile[(i / c) * LDT + (i % c)] = src[i];
            for (int i = lane; i < rows * (CP - c); i += 32) tile[(i / (CP - c)) * LDT + c + (i % (CP - c))] = 0.f;
        }
        __syncwarp();
        float x[RPL][CP];
        const float *bs[RPL];
#pragma unroll
        for (int rr = 0; rr < RPL; ++rr) {
            const int row = lane + 32 * rr < rows ? lane + 32 * rr : 0;   // rows past the end: computed, not stored
#pragma unroll
            for (int q = 0; q < CP / 4; ++q) {
                const float4 v = *reinterpret_cast<const float4 *>(&tile[row * LDT + 4 * q]);
                x[rr][4 * q] = v.x;
                x[rr][4 * q + 1] = v.y;
                x[rr][4 * q + 2] = v.z;
                x[rr][4 * q + 3] = v.w;
            }
            bs[rr] = bias + ((r0 + row) / hw) * c;
        }
        __syncwarp();   // every lane holds its rows: the tile can take the results
#pragma unroll 1
        for (int j0 = 0; j0 < CP; j0 += 4) {
            if (j0 >= c) break;
            float a[RPL][4];
#pragma unroll
            for (int rr = 0; rr < RPL; ++rr)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) a[rr][jj] = 0.f;
#pragma unroll
            for (int q = 0; q < CP / 4; ++q) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float4 gq = *reinterpret_cast<const float4 *>(&sG[(j0 + jj) * CP + 4 * q]);
#pragma unroll
                    for (int rr = 0; rr < RPL; ++rr) {
                        a[rr][jj] = fmaf(x[rr][4 * q], gq.x, a[rr][jj]);
                        a[rr][jj] = fmaf(x[rr][4 * q + 1], gq.y, a[rr][jj]);
                        a[rr][jj] = fmaf(x[rr][4 * q + 2], gq.z, a[rr][jj]);
                        a[rr][jj] = fmaf(x[rr][4 * q + 3], gq.w, a[rr][jj]);
                    }
                }
            }
#pragma unroll
            for (int rr = 0; rr < RPL; ++rr) {
                if (lane + 32 * rr >= rows) continue;
                float4 o;
                o.x = a[rr][0] + (j0 < c ? bs[rr][j0] : 0.f);
                o.y = a[rr][1] + (j0 + 1 < c ? bs[rr][j0 + 1] : 0.f);
                o.z = a[rr][2] + (j0 + 2 < c ? bs[rr][j0 + 2] : 0.f);
                o.w = a[rr][3] + (j0 + 3 < c ? bs[rr][j0 + 3] : 0.f);
                *reinterpret_cast<float4 *>(&tile[(lane + 32 * rr) * LDT + j0]) = o;   // columns >= c: never stored
            }
        }
        __syncwarp();
        float *dst = out + r0 * c;
        if (content) {
            const float *ct = content + r0 * c;
            for (int i = lane; i < total; i += 32) {
                const float o = c == CP ? tile[(i / CP) * LDT + (i % CP)] : tile[(i / c) * LDT + (i % c)];
                dst[i] = o + strength * (ct[i] - o);
            }
        } else if (c == CP) {
            for (int i = lane; i < total; i += 32) dst[i] = tile[(i / CP) * LDT + (i % CP)];
        } else {
            for (int i = lane; i < total; i += 32) dst[i] = tile[(i / c) * LDT + (i % c)];
        }
        __syncwarp();
    }
}

struct SmallWs {
    float *mu_s, *mu_c, *mu_p, *bias, *sum_p, *sum_s, *G, *style_keep, *sig_keep, *part_t, *part_s;
};

// CTAs of the column-sum kernel per sample: whole clusters, >= 128 rows each, two CTAs per SM over all samples
int sum_splits(int64_t hw, int nb) {
    int64_t s = hw / 128;
    int64_t cap = 2 * (int64_t)sm_count() / (nb < 1 ? 1 : nb);
    if (cap > MAX_SUM_SPLITS / (nb < 1 ? 1 : nb)) cap = MAX_SUM_SPLITS / (nb < 1 ? 1 : nb);
    if (s > cap) s = cap;
    s = s / GC * GC;
    if (s < GC) s = GC;
    return (int)s;
}
int gram_ctas(int64_t n) {
    int64_t s = n / (2 * SLAB);  // >= 2 slabs per CTA
    const int64_t cap = (int64_t)sm_count();   // one per SM: the chain kernel reads ctas / GC partial matrices
    if (s > cap) s = cap;
    if (s < 1) s = 1;
    return (int)s;
}

size_t small_layout(int64_t n_t, int64_t n_s, int c, SmallWs *w, void *base, size_t cap, bool *ok) {
    Arena ar(base, cap);
    SmallWs l{};
    l.mu_s = ar.take<float>((size_t)B_MAX_S * c);
    l.mu_c = ar.take<float>((size_t)B_MAX_S * c);
    l.mu_p = ar.take<float>((size_t)B_MAX_S * c);
    l.bias = ar.take<float>((size_t)B_MAX_S * c);
    l.sum_p = ar.take<float>((size_t)MAX_SUM_SPLITS * c);
    l.sum_s = ar.take<float>((size_t)MAX_SUM_SPLITS * c);
    l.G = ar.take<float>((size_t)c * c);
    l.style_keep = ar.take<float>((size_t)64 * 64);
    l.sig_keep = ar.take<float>((size_t)64 * 64);
    l.part_t = ar.take<float>((size_t)gram_ctas(n_t) * c * c);
    l.part_s = ar.take<float>((size_t)gram_ctas(n_s) * c * c);
    if (w) *w = l;
    if (ok) *ok = ar.ok();
    return ar.off;
}

// known_mu != NULL: the means are given ([nb][c]); no column-sum pass, *splits_out = 0 and the consumers read them directly
template <int CP>
int moments_small(const float *X, int nb, int64_t hw, int c, float *part_sum, float *part_gram, int *splits_out,
                  int *nz_out, cudaStream_t st, const float *known_mu = nullptr) {
    const int64_t n = (int64_t)nb * hw;
    int parts = 0;
    if (known_mu) {
        part_sum = const_cast<float *>(known_mu);
    } else {
        const int splits = sum_splits(hw, nb);
        launch_pdl(small_colsum_kernel<CP>, dim3(splits, nb), dim3(ST), 0, st, X, part_sum, hw, c, splits);
        OPTEX_LAUNCH_CHECK("small_colsum_kernel");
        parts = splits / GC;   // what the cluster fold leaves per sample
    }
    int ctas = gram_ctas(n);
    int64_t rows = ((n + ctas - 1) / ctas + SLAB - 1) / SLAB * SLAB;
    ctas = (int)((n + rows - 1) / rows);
    ctas = (ctas + GC - 1) / GC * GC;   // whole clusters; the CTAs past the last row contribute zeros
    launch_pdl(small_gram_kernel<CP>, dim3(ctas), dim3(ST), 0, st, X, (const float *)part_sum, parts, part_gram, n, hw,
               c, rows);
    OPTEX_LAUNCH_CHECK("small_gram_kernel");
    *splits_out = parts;
    *nz_out = ctas / GC;
    return OPTEX_OK;
}

template <int CP>
int step_small(const float *P, const float *S, const float *R, float *out, int b_p, int64_t hw_p, int b_s,
               int64_t hw_s, int c, int mode, float eps, const float *content, float strength, const SmallWs &w,
               cudaStream_t st, int style_reuse) {
    const int64_t n_p = (int64_t)b_p * hw_p, n_s = (int64_t)b_s * hw_s;
    // style_reuse is cov_match.cu's bit field: 1 = style side kept, 2 = pastiche mean known, 4 = a loop follows
    const int loop_flags = style_reuse;
    style_reuse = (loop_flags & 1) ? 1 : 0;
    const bool mean_known = (loop_flags & 3) == 3 && (b_s == 1 || b_s == b_p);
    int nz_t = 0, nz_s = 0, sp_t = 0, sp_s = 0;
    if ((loop_flags & 4) && content) {   // mean(content) per sample, once per loop
        const int sp = sum_splits(hw_p, b_p);
        launch_pdl(small_colsum_kernel<CP>, dim3(sp, b_p), dim3(ST), 0, st, content, w.sum_p, hw_p, c, sp);
        OPTEX_LAUNCH_CHECK("small_colsum_kernel");
        launch_pdl(small_mean_kernel<CP>, dim3(b_p), dim3(ST), 0, st, (const float *)w.sum_p, sp / GC, c, hw_p, w.mu_c);
        OPTEX_LAUNCH_CHECK("small_mean_kernel");
    }
    if (mean_known) {
        launch_pdl(small_known_mean_kernel, dim3(cdiv((int64_t)b_p * c, 128)), dim3(128), 0, st, (const float *)w.mu_s,
                   content ? (const float *)w.mu_c : (const float *)nullptr, strength, b_p, b_s, c, w.mu_p);
        OPTEX_LAUNCH_CHECK("small_known_mean_kernel");
    }
    // 8 = the loop allows the propagated covariance this iteration (no content blend; it re-measures every 8th)
    const bool cov_known = mean_known && !content && (loop_flags & 8);
    const int sig_mode = cov_known ? 2 : ((loop_flags & 5) && !content ? 1 : 0);
    if (!cov_known)
        OPTEX_TRY(moments_small<CP>(P, b_p, hw_p, c, w.sum_p, w.part_t, &sp_t, &nz_t, st,
                                    mean_known ? (const float *)w.mu_p : (const float *)nullptr));
    if (!style_reuse) OPTEX_TRY(moments_small<CP>(S, b_s, hw_s, c, w.sum_s, w.part_s, &sp_s, &nz_s, st));
    const size_t chain_smem = (size_t)(9 * CP * CP + 16 + CP) * sizeof(float);
    constexpr int RPL = CP <= 32 ? 2 : 1;   // rows per lane of the application kernel
    const size_t apply_smem = (size_t)(CP * CP + (ST / 32) * 32 * RPL * (CP + 4)) * sizeof(float);
    static bool attr_done[64] = {};
    int dev = 0;
    OPTEX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        OPTEX_CUDA(cudaFuncSetAttribute(small_chain_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)chain_smem));
        OPTEX_CUDA(cudaFuncSetAttribute(small_apply_kernel<CP, RPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)apply_smem));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    launch_pdl(small_chain_kernel<CP>, dim3(1), dim3(ST), chain_smem, st, (const float *)w.part_t, nz_t, (float)n_p,
               (const float *)w.part_s, nz_s, (float)n_s, w.style_keep, style_reuse ? 1 : 0, mode, eps, c,
               mean_known ? (const float *)w.mu_p : (const float *)w.sum_p, sp_t, hw_p, (const float *)w.sum_s, sp_s, hw_s,
               w.mu_s, b_p, b_s, w.G, w.bias, R, w.sig_keep, sig_mode);
    OPTEX_LAUNCH_CHECK("small_chain_kernel");
    int64_t grid = (n_p + 32 * RPL * (ST / 32) - 1) / (32 * RPL * (ST / 32));
    const int64_t cap = 4 * (int64_t)sm_count();
    if (grid > cap) grid = cap;
    launch_pdl(small_apply_kernel<CP, RPL>, dim3((unsigned)grid), dim3(ST), apply_smem, st, P, (const float *)w.G,
               (const float *)w.bias, out, n_p, hw_p, c, content, strength);
    OPTEX_LAUNCH_CHECK("small_apply_kernel");
    return OPTEX_OK;
}

bool small_enabled() {
    static const bool v = [] {
        const char *e = getenv("OPTEX_COV_SMALL");
        return !(e && atoi(e) == 0);
    }();
    return v;
}

}  // namespace

bool cov_small_supported(int c, int mode, int b_p, int b_s) {
    return small_enabled() && c >= 1 && c <= 64 &&
           (mode == OPTEX_MODE_PCA || mode == OPTEX_MODE_SYM || mode == OPTEX_MODE_CHOL) && b_p <= B_MAX_S &&
           b_s <= B_MAX_S;
}

size_t cov_small_ws_bytes(int64_t n_t, int64_t n_s, int c) {
    if (c < 1 || c > 64) return 0;
    return small_layout(n_t, n_s, c, nullptr, nullptr, 0, nullptr) + 256;
}

int cov_small_step(const float *P, const float *S, const float *R, float *out, int b_p, int64_t hw_p, int b_s,
                   int64_t hw_s, int c, int mode, float eps, const float *content, float strength, void *workspace,
                   size_t workspace_bytes, cudaStream_t st, int style_reuse) {
    SmallWs w;
    bool ok = false;
    small_layout((int64_t)b_p * hw_p, (int64_t)b_s * hw_s, c, &w, workspace, workspace_bytes, &ok);
    if (!ok) {
        set_error("covariance modes (narrow path): workspace %zu < %zu bytes", workspace_bytes,
                  cov_small_ws_bytes((int64_t)b_p * hw_p, (int64_t)b_s * hw_s, c));
        return OPTEX_EWORKSPACE;
    }
    if (c <= 32)
        return step_small<32>(P, S, R, out, b_p, hw_p, b_s, hw_s, c, mode, eps, content, strength, w, st, style_reuse);
    return step_small<64>(P, S, R, out, b_p, hw_p, b_s, hw_s, c, mode, eps, content, strength, w, st, style_reuse);
}

}  // namespace optex
