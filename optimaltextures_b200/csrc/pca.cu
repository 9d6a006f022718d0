// PCA basis of a feature block - the reference's fit_pca() (optex.py:180-190) and the two projections either side
// of the inner loop (optex.py:110 `pastiche_feature @ style_eigvs[l]`, optex.py:120 `@ style_eigvs[l].T`).
//
//   reference:  A = X - mean(X)  (ONE scalar mean over all elements, optex.py:182)
//               _, sigma, V = torch.svd(A)                      (sigma descending)
//               k = first index with cumsum(sigma / sum(sigma)) > 0.9 ;  eigvecs = V[:, :k]
//               features = X @ eigvecs                           (the un-centred X)
//
// B200 formulation.  The right singular vectors of A are the eigenvectors of the C x C Gram matrix
// G = A^T A and sigma = sqrt(lambda), so the [n, C] SVD (n up to 2.4 M rows) becomes
//   1. column sums + raw Gram X^T X accumulated in FP64 (symmetric 64 x 64 tiles, split over the rows,
//      deterministic reduction), centred algebraically:  G = X^T X - m (s 1^T + 1 s^T) + n m^2
//      - in FP64 the cancellation costs nothing that fp32 outputs can see, and X is read exactly once;
//   2. a parallel one-sided Jacobi eigensolver on G in FP64: the rows w_i of W (W = G at the start) are rotated pairwise
//      until they are mutually orthogonal.  Cold solves (no basis requested) take pca_jacobi_blk_kernel: blocks of 8
//      rows paired round-robin across CTAs (C/8 - 1 grid barriers per sweep), the 8 x 8 cross pairs of a block pair
//      rotated inside the CTA, no V - at convergence w_i = lambda_i v_i, so the eigenvectors are the normalised rows.
//      A `basis` request (optex_fit_pca_warm) takes pca_jacobi_kernel: one pair per CTA, C - 1 barriers per sweep,
//      rows v_i of V^T rotated along.  Blocks with fewer rows than channels are solved in the dual form (the n x n
//      matrix A A^T, eigenvectors mapped through A^T: pca_dual_vecs_kernel);
//   3. lambda_i (|w_i| without V, v_i . w_i with it), sort, the reference's 90 % rule in fp32, sign convention, fp32 store.
// FP64 throughout because sigma = sqrt(lambda) squares the conditioning: an fp32 Gram would leave the small singular
// values (which all enter the 90 % rule's normaliser) with absolute errors of 1e-3 sigma_max.
//
// Sign convention: an SVD fixes each singular vector only up to sign (and the reference's LAPACK gives whatever it
// gives); here the component of largest magnitude of every eigenvector is made positive.  Parity is therefore stated
// on sigma, k, the projector V_k V_k^T and on sign-aligned columns (tests/test_gpu_pca.py).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace cg = cooperative_groups;

namespace optex {
namespace {

constexpr int PCA_MAX_C = 1024;
constexpr int GT = 64;           // Gram tile edge
constexpr int GK = 16;           // rows of X per shared-memory slab
constexpr int SUM_SPLITS = 64;   // row splits of the column-sum pass
constexpr int MAX_GRAM_SPLITS = 64;
constexpr int JT = 128;          // threads per Jacobi pair
constexpr int MAX_SWEEPS = 40;
constexpr double JTOL = 1e-12;   // rotate while |w_p . w_q| > JTOL |w_p| |w_q|
constexpr double JFLOOR = 1e-13; // ... and > (JFLOOR trace G)^2 (keeps null-space noise from rotating for ever)
// Early exit: the cyclic Jacobi iteration converges quadratically, so a sweep whose LARGEST rotated pair had
// |w_p . w_q| <= JEXIT |w_p| |w_q| leaves every pair at ~JEXIT^2 - below JTOL - and the "empty" confirming sweep
// (c - 1 more grid barriers) is not run.  OPTEX_PCA_EXIT=0 restores the run-until-no-rotation rule.
constexpr double JEXIT = 1e-6;

inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// part[z][j] = sum over the rows of split z of X[r, j]
__global__ void pca_colsum_kernel(const float *__restrict__ X, double *__restrict__ part, int64_t n, int c,
                                  int splits) {
    pdl_wait();
    __shared__ double red[8][33];
    const int ch = blockIdx.x * 32 + threadIdx.x;
    const int split = blockIdx.y;
    const int64_t rows = (n + splits - 1) / splits;
    const int64_t r0 = split * rows, r1 = r0 + rows < n ? r0 + rows : n;
    double acc = 0.0;
    if (ch < c)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) acc += (double)X[r * c + ch];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && ch < c) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
        part[(int64_t)split * c + ch] = s;
    }
}

// s[j] = sum_z part[z][j];  stat[0] = mean over ALL elements (optex.py:182 `tensor.mean()`), stat[1] = n
__global__ void __launch_bounds__(1024) pca_mean_kernel(const double *__restrict__ part, int splits, int c, int64_t n,
                                                        double *__restrict__ s, double *__restrict__ stat) {
    pdl_wait();
    __shared__ double red[32];
    const int tid = threadIdx.x;
    double v = 0.0;
    if (tid < c) {
        for (int z = 0; z < splits; ++z) v += part[(int64_t)z * c + tid];
        s[tid] = v;
    }
    v = warp_sum(v);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int i = 0; i < 32; ++i) t += red[i];
        stat[0] = t / ((double)n * (double)c);
        stat[1] = (double)n;
    }
}

// part[z][i][j] = sum over the rows of split z of X[r, i] X[r, j]  for the tile pairs (bi <= bj) only
__global__ void __launch_bounds__(256) pca_gram_kernel(const float *__restrict__ X, double *__restrict__ part,
                                                       int64_t n, int c, int T, int64_t rows_per_split) {
    pdl_wait();
    __shared__ double As[GK][GT], Bs[GK][GT];
    int t = blockIdx.x, bi = 0;
    while (t >= T - bi) {
        t -= T - bi;
        ++bi;
    }
    const int bj = bi + t;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r1 = r0 + rows_per_split < n ? r0 + rows_per_split : n;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int ci = bi * GT, cj = bj * GT;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    float pa[4], pb[4];
    auto fetch = [&](int64_t r) {
        const int64_t row = r + ty;
        const bool rok = row < r1;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int col = tx + 16 * q;
            pa[q] = (rok && ci + col < c) ? X[row * c + ci + col] : 0.f;
            pb[q] = (rok && cj + col < c) ? X[row * c + cj + col] : 0.f;
        }
    };
    if (r0 < r1) fetch(r0);
    for (int64_t r = r0; r < r1; r += GK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            As[ty][tx + 16 * q] = (double)pa[q];
            Bs[ty][tx + 16 * q] = (double)pb[q];
        }
        __syncthreads();
        if (r + GK < r1) fetch(r + GK);  // next slab in flight during the FMAs
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    double *out = part + (int64_t)blockIdx.y * c * c;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = ci + ty + 16 * i;
        if (gi >= c) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = cj + tx + 16 * j;
            if (gj < c) out[(int64_t)gi * c + gj] = acc[i][j];
        }
    }
}

// G = sum_z part_z - m (s_i + s_j) + n m^2 (mirrored from the computed tile pairs);  W = G, V = I
// (warm start: V is NULL and stays what the caller passed in; W then receives G only as the input of pca_warm_kernel)
__global__ void pca_gram_final_kernel(const double *__restrict__ part, int nz, const double *__restrict__ s,
                                      const double *__restrict__ stat, int c, double *__restrict__ W,
                                      double *__restrict__ V) {
    pdl_wait();
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)c * c) return;
    const int i = (int)(idx / c), j = (int)(idx % c);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    double g = 0.0;
    for (int z = 0; z < nz; ++z) g += part[(int64_t)z * c * c + (int64_t)lo * c + hi];
    const double m = stat[0], n = stat[1];
    g = g - m * (s[i] + s[j]) + n * m * m;
    W[idx] = g;
    if (V) V[idx] = i == j ? 1.0 : 0.0;
}

// Warm start: W = V G for an orthogonal V carried over from a previous solve (rows w_i = G v_i; G symmetric).
// 64 x 64 output tiles, 4 x 4 per thread, FP64 FMAs (c^3 of them: 134 M at c = 512, ~20 us).
__global__ void __launch_bounds__(256) pca_warm_kernel(const double *__restrict__ V, const double *__restrict__ G,
                                                       double *__restrict__ W, int c) {
    pdl_wait();
    __shared__ double As[GK][GT + 1], Bs[GK][GT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < c; k0 += GK) {
        // As[k][i] = V[i0 + i][k0 + k]  (16 consecutive k per row i), Bs[k][j] = G[k0 + k][j0 + j]
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = ty + 16 * q, k = tx;
            As[k][i] = (i0 + i < c && k0 + k < c) ? V[(int64_t)(i0 + i) * c + k0 + k] : 0.0;
            const int j = tx + 16 * q, kk = ty;
            Bs[kk][j] = (k0 + kk < c && j0 + j < c) ? G[(int64_t)(k0 + kk) * c + j0 + j] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = i0 + ty + 16 * i;
        if (gi >= c) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = j0 + tx + 16 * j;
            if (gj < c) W[(int64_t)gi * c + gj] = acc[i][j];
        }
    }
}

// One-sided Jacobi on the rows of W (= G V^T) and V^T; see the file header.  Cooperative launch: gridDim.x CTAs are
// co-resident, each takes the pairs pi = blockIdx.x, blockIdx.x + gridDim.x, ... of every round.
// Round-robin schedule of ne = c (+1 if odd) players: round r pairs (ne-1, r) and ((r+i) mod (ne-1), (r-i) mod (ne-1)).
template <int EPT>
__global__ void __launch_bounds__(JT) pca_jacobi_kernel(double *W, double *V, const double *G, int c, int max_sweeps,
                                                        unsigned *rot_count, unsigned *rot_max, float exit_r2,
                                                        int *info) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[2][3][JT / 32];
    __shared__ double s_tr[JT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double tr = 0.0;
    for (int i = tid; i < c; i += JT) tr += __ldcg(&G[(int64_t)i * c + i]);  // G: W itself on a cold start
    tr = warp_sum(tr);
    if (lane == 0) s_tr[warp] = tr;
    __syncthreads();
    tr = 0.0;
#pragma unroll
    for (int i = 0; i < JT / 32; ++i) tr += s_tr[i];
    const double floor_abs = (JFLOOR * tr) * (JFLOOR * tr);
    const int ne = c + (c & 1), n1 = ne - 1, m = ne / 2;
    int par = 0, sweep = 0;
    unsigned local_rot = 0;
    float local_max = 0.f;  // largest (w_p . w_q)^2 / (|w_p|^2 |w_q|^2) among the pairs this CTA rotated in the sweep
    for (; sweep < max_sweeps; ++sweep) {
        for (int r = 0; r < n1; ++r) {
            for (int pi = blockIdx.x; pi < m; pi += gridDim.x) {
                const int a = pi == 0 ? n1 : (r + pi) % n1;
                const int b = pi == 0 ? r : (r - pi + n1) % n1;
                if (a >= c || b >= c) continue;  // the dummy player of an odd c
                double *wa = W + (int64_t)a * c, *wb = W + (int64_t)b * c;
                double *va = V + (int64_t)a * c, *vb = V + (int64_t)b * c;
                double wp[EPT], wq[EPT], vp[EPT], vq[EPT];
                double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    const int i = tid + JT * e;
                    const bool ok = i < c;
                    wp[e] = ok ? __ldcg(wa + i) : 0.0;
                    wq[e] = ok ? __ldcg(wb + i) : 0.0;
                    vp[e] = ok ? __ldcg(va + i) : 0.0;
                    vq[e] = ok ? __ldcg(vb + i) : 0.0;
                }
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    al = fma(wp[e], wp[e], al);
                    be = fma(wq[e], wq[e], be);
                    ga = fma(wp[e], wq[e], ga);
                }
                al = warp_sum(al);
                be = warp_sum(be);
                ga = warp_sum(ga);
                if (lane == 0) {
                    red[par][0][warp] = al;
                    red[par][1][warp] = be;
                    red[par][2][warp] = ga;
                }
                __syncthreads();
                al = be = ga = 0.0;
#pragma unroll
                for (int i = 0; i < JT / 32; ++i) {
                    al += red[par][0][i];
                    be += red[par][1][i];
                    ga += red[par][2][i];
                }
                par ^= 1;
                if (ga * ga > JTOL * JTOL * al * be && fabs(ga) > floor_abs) {
                    const double zeta = (be - al) / (2.0 * ga);
                    const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
                    for (int e = 0; e < EPT; ++e) {
                        const int i = tid + JT * e;
                        if (i < c) {
                            __stcg(wa + i, cs * wp[e] - sn * wq[e]);
                            __stcg(wb + i, sn * wp[e] + cs * wq[e]);
                            __stcg(va + i, cs * vp[e] - sn * vq[e]);
                            __stcg(vb + i, sn * vp[e] + cs * vq[e]);
                        }
                    }
                    ++local_rot;
                    local_max = fmaxf(local_max, (float)((ga * ga) / (al * be)));
                }
            }
            if (r == n1 - 1 && tid == 0 && local_rot) {
                atomicAdd(&rot_count[sweep], local_rot);
                atomicMax(&rot_max[sweep], __float_as_uint(local_max));  // non-negative floats order like their bits
            }
            grid.sync();
        }
        local_rot = 0;
        local_max = 0.f;
        // both words were last written before the barrier above: every CTA takes the same branch
        if (__ldcg(&rot_count[sweep]) == 0u) break;  // a whole sweep without a rotation: converged
        if (__uint_as_float(__ldcg(&rot_max[sweep])) <= exit_r2) {  // quadratic convergence: now at ~exit_r2^2
            break;
        }
    }
    if (blockIdx.x == 0 && tid == 0) info[0] = sweep < max_sweeps ? sweep + 1 : max_sweeps;  // sweeps started
}

// ---------------------------------------------------------------------------------------------------------------
// Blocked-order one-sided Jacobi (round 2; the cold solve).  The same rotations as pca_jacobi_kernel, ordered so
// that most of them need no grid-wide synchronisation: the rows of W form c_pad / BR blocks of BR = 8 rows; a
// "global round" pairs the blocks round-robin (one CTA per block pair, the grid barrier and the trip through L2
// happen HERE: nb - 1 times per sweep instead of c - 1), and inside a global round the CTA rotates all BR x BR
// cross pairs of its two blocks in BR sub-rounds of BR disjoint pairs (one warp per pair, __syncthreads between
// sub-rounds): warp w keeps row w of block A in registers for the whole global round while the rows of block B pass
// by through shared memory.  Global round 0 of every sweep runs the full 2 BR-row tournament instead, which also
// covers the pairs inside each block - every pair of rows meets exactly once per sweep.  Sweep counts are those of
// the round-robin order for full-rank blocks and ~25 % higher for rank-deficient ones (scripts/jacobi_sim.py
// --blocked-order); a sweep costs nb - 1 = 63 barriers at c = 512 instead of 511.
// No V: W starts as G, so at convergence row i is lambda_i v_i - the eigenvectors are the normalised rows and
// lambda_i = |w_i| (pca_finish_rows_kernel); the null-space rows carry no direction, which the 90 % rule never keeps.
// Layout: W is [c_pad][ld] with ld = 32 EPL >= c, zero padded; lane l of a warp holds the double2 chunks l, l + 32, ...
constexpr int BR = 8;

__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        unsigned seen;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while ((int)(seen - target) < 0);
    }
    __syncthreads();
}

// One pair: the rows a, b (this lane's chunks) with their squared norms al, be.  Only w_a . w_b is reduced over the
// warp - the norms are carried along exactly (|a'|^2 = al - t ga, |b'|^2 = be + t ga with t = tan(theta)) and
// recomputed from the data once per global round.  The angle without a division (FP64 div / sqrt are ~30-instruction
// dependent sequences, four of them sat on every sub-round's critical path): with d = be - al, g = 2 ga,
// r = 1 / hypot(d, g):  cos^2 = (1 + |d| r) / 2,  sin = sign(d) g r / (2 cos)  - two rsqrt.
template <int NV>
__device__ __forceinline__ void pair_step(double2 (&a)[NV], double2 (&b)[NV], double &al, double &be,
                                          double floor_abs, double exit_r2, unsigned &rot, unsigned &big) {
    double g0 = 0.0, g1 = 0.0;
#pragma unroll
    for (int e = 0; e < NV; ++e) {
        g0 = fma(a[e].x, b[e].x, g0);
        g1 = fma(a[e].y, b[e].y, g1);
    }
    const double ga = warp_sum(g0 + g1);
    const double g2 = ga * ga, ab = al * be;
    if (!(g2 > JTOL * JTOL * ab && fabs(ga) > floor_abs)) return;
    if (g2 > exit_r2 * ab) big = 1u;
    const double d = be - al, g = 2.0 * ga;
    const double r = rsqrt(fma(d, d, g * g));
    const double x = fma(0.5 * fabs(d), r, 0.5);  // cos^2 in [1/2, 1]
    const double ic = rsqrt(x);
    const double cs = x * ic, sn = (d >= 0.0 ? 0.5 : -0.5) * g * r * ic, t = sn * ic;
#pragma unroll
    for (int e = 0; e < NV; ++e) {
        const double ax = a[e].x, ay = a[e].y, bx = b[e].x, by = b[e].y;
        a[e].x = fma(cs, ax, -sn * bx);
        a[e].y = fma(cs, ay, -sn * by);
        b[e].x = fma(sn, ax, cs * bx);
        b[e].y = fma(sn, ay, cs * by);
    }
    al = fmax(fma(-t, ga, al), 0.0);
    be = fma(t, ga, be);
    ++rot;
}

template <int NV>
__device__ __forceinline__ double row_norm2(const double2 (&a)[NV]) {
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int e = 0; e < NV; ++e) {
        s0 = fma(a[e].x, a[e].x, s0);
        s1 = fma(a[e].y, a[e].y, s1);
    }
    return warp_sum(s0 + s1);
}

template <int EPL>
__global__ void __launch_bounds__(BR * 32) pca_jacobi_blk_kernel(double *W, int c_pad, int max_sweeps,
                                                                 unsigned *rot_count, unsigned *rot_max,
                                                                 float exit_r2f, int *info, const double *stat,
                                                                 unsigned *bar, long long *stamps, int n_stamps) {
    constexpr int LD = 32 * EPL, NV = EPL / 2;
    extern __shared__ __align__(16) unsigned char jsm_raw[];
    double2 *rows = reinterpret_cast<double2 *>(jsm_raw);  // [2 BR][LD / 2]: block A rows 0..7, block B rows 8..15
    __shared__ double nrm[2 * BR];                         // their squared norms
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double tr = stat[2];
    const double floor_abs = (JFLOOR * tr) * (JFLOOR * tr), exit_r2 = (double)exit_r2f;
    const int nb = c_pad / BR, n1 = nb - 1;
    unsigned target = 0, local_rot = 0, local_big = 0;
    int sweep = 0;
    auto srow = [&](int r) { return rows + (size_t)r * (LD / 2); };
    for (; sweep < max_sweeps; ++sweep) {
        for (int gr = 0; gr < n1; ++gr) {
            const int pi = blockIdx.x;
            const int A = pi == 0 ? n1 : (gr + pi) % n1;
            const int B = pi == 0 ? gr : (gr - pi + n1) % n1;
            // debug stamps (optex_debug_pca_stamps): clock64 of CTA 1's thread 0 at 5 points of its first rounds
            const int sidx = sweep * n1 + gr;
            const bool stamp = stamps && blockIdx.x == 1 && tid == 0 && sidx < n_stamps;
            if (stamp) stamps[5 * sidx + 0] = clock64();
            double2 *ga_row = reinterpret_cast<double2 *>(W + (size_t)(A * BR + warp) * LD);
            double2 *gb_row = reinterpret_cast<double2 *>(W + (size_t)(B * BR + warp) * LD);
            double2 a[NV], b[NV];
#pragma unroll
            for (int e = 0; e < NV; ++e) {
                a[e] = __ldcg(ga_row + lane + 32 * e);
                b[e] = __ldcg(gb_row + lane + 32 * e);
            }
#pragma unroll
            for (int e = 0; e < NV; ++e) srow(BR + warp)[lane + 32 * e] = b[e];
            double al = row_norm2<NV>(a);
            {
                const double be = row_norm2<NV>(b);
                if (lane == 0) nrm[BR + warp] = be;
            }
            if (gr == 0) {
                // the full tournament of the 2 BR rows (15 sub-rounds): both rows of a pair through shared memory
#pragma unroll
                for (int e = 0; e < NV; ++e) srow(warp)[lane + 32 * e] = a[e];
                if (lane == 0) nrm[warp] = al;
                __syncthreads();
                constexpr int M1 = 2 * BR - 1;
                for (int sr = 0; sr < M1; ++sr) {
                    const int x = warp == 0 ? M1 : (sr + warp) % M1;
                    const int y = warp == 0 ? sr : (sr - warp + M1) % M1;
#pragma unroll
                    for (int e = 0; e < NV; ++e) {
                        a[e] = srow(x)[lane + 32 * e];
                        b[e] = srow(y)[lane + 32 * e];
                    }
                    double nx = nrm[x], ny = nrm[y];
                    const unsigned before = local_rot;
                    pair_step<NV>(a, b, nx, ny, floor_abs, exit_r2, local_rot, local_big);
                    if (local_rot != before) {
#pragma unroll
                        for (int e = 0; e < NV; ++e) {
                            srow(x)[lane + 32 * e] = a[e];
                            srow(y)[lane + 32 * e] = b[e];
                        }
                        if (lane == 0) {
                            nrm[x] = nx;
                            nrm[y] = ny;
                        }
                    }
                    __syncthreads();
                }
#pragma unroll
                for (int e = 0; e < NV; ++e) {
                    __stcg(ga_row + lane + 32 * e, srow(warp)[lane + 32 * e]);
                    __stcg(gb_row + lane + 32 * e, srow(BR + warp)[lane + 32 * e]);
                }
            } else {
                __syncthreads();
                if (stamp) stamps[5 * sidx + 1] = clock64();
                for (int sr = 0; sr < BR; ++sr) {
                    const int j = BR + ((warp + sr) & (BR - 1));
#pragma unroll
                    for (int e = 0; e < NV; ++e) b[e] = srow(j)[lane + 32 * e];
                    double be = nrm[j];
                    const unsigned before = local_rot;
                    pair_step<NV>(a, b, al, be, floor_abs, exit_r2, local_rot, local_big);
                    if (local_rot != before) {
#pragma unroll
                        for (int e = 0; e < NV; ++e) srow(j)[lane + 32 * e] = b[e];
                        if (lane == 0) nrm[j] = be;
                    }
                    __syncthreads();
                }
                if (stamp) stamps[5 * sidx + 2] = clock64();
#pragma unroll
                for (int e = 0; e < NV; ++e) {
                    __stcg(ga_row + lane + 32 * e, a[e]);
                    __stcg(gb_row + lane + 32 * e, srow(BR + warp)[lane + 32 * e]);
                }
            }
            if (gr == n1 - 1 && lane == 0 && local_rot) {
                atomicAdd(&rot_count[sweep], local_rot);
                if (local_big) atomicMax(&rot_max[sweep], __float_as_uint(1.0f));  // "a pair above the exit level"
            }
            if (stamp) stamps[5 * sidx + 3] = clock64();
            grid_barrier(bar, target);
            if (stamp) stamps[5 * sidx + 4] = clock64();
        }
        local_rot = 0;
        local_big = 0;
        if (__ldcg(&rot_count[sweep]) == 0u) break;
        if (__ldcg(&rot_max[sweep]) == 0u) break;  // every rotated pair was below the exit level: converged
    }
    if (blockIdx.x == 0 && tid == 0) info[0] = sweep < max_sweeps ? sweep + 1 : max_sweeps;
}

// Wp[c_pad][ld] = G zero-padded;  stat[2] = trace(G)
__global__ void __launch_bounds__(256) pca_pad_kernel(const double *__restrict__ G, int c, int c_pad, int ld,
                                                      double *__restrict__ Wp, double *__restrict__ stat) {
    pdl_wait();
    const int64_t total = (int64_t)c_pad * ld;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / ld), j = (int)(idx % ld);
        Wp[idx] = (i < c && j < c) ? G[(int64_t)i * c + j] : 0.0;
    }
    if (blockIdx.x == 0) {
        __shared__ double red[8];
        double tr = 0.0;
        for (int i = threadIdx.x; i < c; i += 256) tr += G[(int64_t)i * c + i];
        tr = warp_sum(tr);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tr;
        __syncthreads();
        if (threadIdx.x == 0) {
            tr = 0.0;
            for (int i = 0; i < 8; ++i) tr += red[i];
            stat[2] = tr;
        }
    }
}

// The no-V finish: lambda_i = |w_i|, descending order, sigma = sqrt(lambda), the 90 % rule (optex.py:184);
// scale[i] = +-1 / |w_i| (the component of largest magnitude of every eigenvector positive), 0 for a null row
__global__ void __launch_bounds__(1024) pca_finish_rows_kernel(const double *__restrict__ W, int ld, int c,
                                                               float *__restrict__ sigma, int32_t *__restrict__ k_out,
                                                               int *__restrict__ perm, double *__restrict__ scale) {
    pdl_wait();
    __shared__ double lam[PCA_MAX_C];
    __shared__ float sg[PCA_MAX_C];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = warp; i < c; i += 32) {
        const double *w = W + (int64_t)i * ld;
        double ww = 0.0, best = -1.0;
        int besti = 0;
        for (int k = lane; k < c; k += 32) {
            const double x = w[k];
            ww = fma(x, x, ww);
            if (fabs(x) > best) {
                best = fabs(x);
                besti = k;
            }
        }
        ww = warp_sum(ww);
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) {
                best = ob;
                besti = oi;
            }
        }
        if (lane == 0) {
            const double nrm = sqrt(ww);
            lam[i] = nrm;
            scale[i] = nrm > 0.0 ? (w[besti] < 0.0 ? -1.0 : 1.0) / nrm : 0.0;
        }
    }
    __syncthreads();
    if (tid < c) {
        const double mine = lam[tid];
        int rank = 0;
        for (int j = 0; j < c; ++j) {
            const double o = lam[j];
            rank += (o > mine || (o == mine && j < tid)) ? 1 : 0;
        }
        perm[rank] = tid;
        sg[rank] = (float)sqrt(mine);
    }
    __syncthreads();
    if (tid < c) sigma[tid] = sg[tid];
    if (tid == 0) {
        float total = 0.f;
        for (int i = 0; i < c; ++i) total = __fadd_rn(total, sg[i]);
        float cum = 0.f;
        int k = 0;
        for (int i = 0; i < c; ++i) {
            cum = __fadd_rn(cum, __fdiv_rn(sg[i], total));
            if (cum > 0.9f) {
                k = i;
                break;
            }
        }
        k_out[0] = k;
    }
}

// eigvecs[r][j] = scale[perm[j]] * W[perm[j]][r]  (fp32; column j = j-th principal direction)
__global__ void pca_vecs_rows_kernel(const double *__restrict__ W, int ld, const int *__restrict__ perm,
                                     const double *__restrict__ scale, int c, float *__restrict__ eigvecs) {
    pdl_wait();
    __shared__ float tile[32][33];
    const int j0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += 8) {
        const int j = j0 + jj, r = r0 + threadIdx.x;
        float v = 0.f;
        if (j < c && r < c) {
            const int src = perm[j];
            v = (float)(scale[src] * W[(int64_t)src * ld + r]);
        }
        tile[jj][threadIdx.x] = v;
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += 8) {
        const int r = r0 + rr, j = j0 + threadIdx.x;
        if (r < c && j < c) eigvecs[(int64_t)r * c + j] = tile[threadIdx.x][rr];
    }
}

// lambda_i = v_i . w_i, descending order, sigma = sqrt(lambda), the 90 % rule (optex.py:184), sign convention
__global__ void __launch_bounds__(1024) pca_finish_kernel(const double *__restrict__ W, const double *__restrict__ V,
                                                          int c, float *__restrict__ sigma, int32_t *__restrict__ k_out,
                                                          int *__restrict__ perm, float *__restrict__ sign) {
    pdl_wait();
    __shared__ double lam[PCA_MAX_C];
    __shared__ float sg[PCA_MAX_C];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = warp; i < c; i += 32) {
        const double *w = W + (int64_t)i * c, *v = V + (int64_t)i * c;
        double vw = 0.0, vv = 0.0, best = -1.0;
        int besti = 0;
        for (int k = lane; k < c; k += 32) {
            const double x = v[k];
            vw = fma(x, w[k], vw);
            vv = fma(x, x, vv);
            if (fabs(x) > best) {
                best = fabs(x);
                besti = k;
            }
        }
        vw = warp_sum(vw);
        vv = warp_sum(vv);
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) {
                best = ob;
                besti = oi;
            }
        }
        if (lane == 0) {
            lam[i] = vv > 0.0 ? vw / vv : 0.0;
            sign[i] = v[besti] < 0.0 ? -1.f : 1.f;
        }
    }
    if (tid < c) perm[tid] = tid;
    __syncthreads();
    if (tid < c) {
        const double mine = lam[tid];
        int rank = 0;
        for (int j = 0; j < c; ++j) {
            const double o = lam[j];
            rank += (o > mine || (o == mine && j < tid)) ? 1 : 0;
        }
        perm[rank] = tid;
        sg[rank] = (float)sqrt(mine > 0.0 ? mine : 0.0);
    }
    __syncthreads();
    if (tid < c) sigma[tid] = sg[tid];
    if (tid == 0) {
        // optex.py:184: k = (cumsum(sigma / sum(sigma)) > 0.9).max(0).indices - the FIRST index over 0.9 (so the
        // component that crosses the threshold is itself dropped), 0 when nothing crosses; all in fp32
        float total = 0.f;
        for (int i = 0; i < c; ++i) total = __fadd_rn(total, sg[i]);
        float cum = 0.f;
        int k = 0;
        for (int i = 0; i < c; ++i) {
            cum = __fadd_rn(cum, __fdiv_rn(sg[i], total));
            if (cum > 0.9f) {
                k = i;
                break;
            }
        }
        k_out[0] = k;
    }
}

// eigvecs[r][j] = sign * V[perm[j]][r]  (fp32; column j = j-th principal direction)
__global__ void pca_vecs_kernel(const double *__restrict__ V, const int *__restrict__ perm,
                                const float *__restrict__ sign, int c, float *__restrict__ eigvecs) {
    pdl_wait();
    __shared__ float tile[32][33];
    const int j0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += 8) {
        const int j = j0 + jj, r = r0 + threadIdx.x;
        float v = 0.f;
        if (j < c && r < c) {
            const int src = perm[j];
            v = (float)((double)sign[src] * V[(int64_t)src * c + r]);
        }
        tile[jj][threadIdx.x] = v;
    }
    __syncthreads();
    for (int rr = threadIdx.y; rr < 32; rr += 8) {
        const int r = r0 + rr, j = j0 + threadIdx.x;
        if (r < c && j < c) eigvecs[(int64_t)r * c + j] = tile[threadIdx.x][rr];
    }
}

// Dual form (n < c): the solver ran on the n x n matrix A A^T (A = X - mean), whose eigenvectors u_j are the LEFT
// singular vectors; the right ones are v_j = A^T u_j / sigma_j.  One CTA per j: u_j = the normalised row perm[j] of W
// in shared memory, every thread walks the rows of X for its channels in FP64, then the vector is normalised by its
// own norm (= sigma_j up to rounding), the component of largest magnitude made positive, and stored as column j.
__global__ void __launch_bounds__(256) pca_dual_vecs_kernel(const float *__restrict__ X, int64_t n, int c,
                                                            const double *__restrict__ W, int ld,
                                                            const int *__restrict__ perm,
                                                            const double *__restrict__ scale,
                                                            const double *__restrict__ stat, float *__restrict__ eigvecs) {
    pdl_wait();
    __shared__ double u[PCA_MAX_C];
    __shared__ double red[8];
    __shared__ double rbest[8];
    __shared__ int ribest[8];
    const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int src = perm[j];
    const double sc = fabs(scale[src]);
    double usum = 0.0;
    for (int r = tid; r < n; r += 256) {
        const double v = W[(int64_t)src * ld + r] * sc;
        u[r] = v;
        usum += v;
    }
    usum = warp_sum(usum);
    if (lane == 0) red[warp] = usum;
    __syncthreads();
    usum = 0.0;
    for (int i = 0; i < 8; ++i) usum += red[i];
    __syncthreads();
    const double m = stat[0];
    constexpr int CPT = PCA_MAX_C / 256;   // channels per thread
    double acc[CPT];
    double nrm = 0.0, best = -1.0;
    int besti = 0;
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
        const int ch = tid + 256 * q;
        acc[q] = 0.0;
        if (ch < c) {
            double a = 0.0;
            for (int64_t r = 0; r < n; ++r) a = fma((double)X[r * c + ch], u[r], a);
            a -= m * usum;
            acc[q] = a;
            nrm = fma(a, a, nrm);
            if (fabs(a) > best) {
                best = fabs(a);
                besti = ch;
            }
        }
    }
    nrm = warp_sum(nrm);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        if (ob > best || (ob == best && oi < besti)) {
            best = ob;
            besti = oi;
        }
    }
    if (lane == 0) {
        red[warp] = nrm;
        rbest[warp] = best;
        ribest[warp] = besti;
    }
    __syncthreads();
    nrm = 0.0;
    best = -1.0;
    besti = 0;
    for (int i = 0; i < 8; ++i) {
        nrm += red[i];
        if (rbest[i] > best || (rbest[i] == best && ribest[i] < besti)) {
            best = rbest[i];
            besti = ribest[i];
        }
    }
    __shared__ double s_sign;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < CPT; ++q)
        if (tid + 256 * q == besti) s_sign = acc[q] < 0.0 ? -1.0 : 1.0;
    __syncthreads();
    const double f = nrm > 0.0 ? s_sign / sqrt(nrm) : 0.0;
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
        const int ch = tid + 256 * q;
        if (ch < c) eigvecs[(int64_t)ch * c + j] = (float)(acc[q] * f);
    }
}

struct PcaWs {
    double *colpart, *s, *stat, *gpart, *W, *V, *Wp, *scale;
    unsigned *rot, *rotmax, *bar;
    int *info, *perm;
    float *sign;
};

// padded shape of the blocked-order solver: rows to a multiple of 2 BR, row length 32 EPL in {64, ..., 1024}
inline int blk_ld(int c) {
    int ld = 64;
    while (ld < c) ld *= 2;
    return ld;
}
inline int blk_rows(int c) { return (c + 2 * BR - 1) / (2 * BR) * (2 * BR); }

int gram_splits(int64_t n, int c) {
    const int T = (c + GT - 1) / GT;
    const int pairs = T * (T + 1) / 2;
    int64_t splits = (4 * 148 + pairs - 1) / pairs;
    if (splits > MAX_GRAM_SPLITS) splits = MAX_GRAM_SPLITS;
    const int64_t by_rows = n / 256 < 1 ? 1 : n / 256;
    if (splits > by_rows) splits = by_rows;
    return (int)splits;
}

size_t pca_layout(int64_t n, int c, PcaWs *w, void *base, size_t cap, bool *ok) {
    Arena ar(base, cap);
    const size_t cc = (size_t)c * c;
    PcaWs l{};
    l.colpart = ar.take<double>((size_t)SUM_SPLITS * c);
    l.s = ar.take<double>(c);
    l.stat = ar.take<double>(8);
    l.gpart = ar.take<double>((size_t)gram_splits(n, c) * cc);
    l.W = ar.take<double>(cc);
    l.V = ar.take<double>(cc);
    l.rot = ar.take<unsigned>(2 * (MAX_SWEEPS + 8));
    l.rotmax = l.rot + MAX_SWEEPS + 8;
    l.info = ar.take<int>(8);
    l.perm = ar.take<int>(c);
    l.sign = ar.take<float>(c);
    l.Wp = ar.take<double>((size_t)blk_rows(c) * blk_ld(c));
    l.scale = ar.take<double>(c);
    l.bar = ar.take<unsigned>(8);
    if (w) *w = l;
    if (ok) *ok = ar.ok();
    return ar.off;
}

float jacobi_exit_r2() {
    static const float v = [] {
        const char *e = getenv("OPTEX_PCA_EXIT");
        const double r = e ? atof(e) : JEXIT;
        return (float)(r * r);
    }();
    return v;
}

template <int EPT>
int launch_jacobi(const PcaWs &w, const double *G, int c, cudaStream_t st) {
    int per_sm = 0;
    OPTEX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pca_jacobi_kernel<EPT>, JT, 0));
    if (per_sm < 1) {
        set_error("optex_fit_pca: the Jacobi kernel does not fit on an SM");
        return OPTEX_ECUDA;
    }
    const int pairs = (c + 1) / 2;
    int grid = per_sm * sm_count();
    if (grid > pairs) grid = pairs;
    if (grid < 1) grid = 1;
    double *W = w.W, *V = w.V;
    int cc = c, ms = MAX_SWEEPS;
    unsigned *rot = w.rot, *rotmax = w.rotmax;
    int *info = w.info;
    float exit_r2 = jacobi_exit_r2();
    void *args[] = {&W, &V, &G, &cc, &ms, &rot, &rotmax, &exit_r2, &info};
    OPTEX_CUDA(cudaLaunchCooperativeKernel((const void *)pca_jacobi_kernel<EPT>, dim3(grid), dim3(JT), args, 0, st));
    count_launch();
    return OPTEX_OK;
}

long long *g_pca_stamps = nullptr;  // optex_debug_pca_stamps
int g_pca_n_stamps = 0;

template <int EPL>
int launch_jacobi_blk(const PcaWs &w, int c, cudaStream_t st) {
    const int c_pad = blk_rows(c);
    const size_t smem = (size_t)2 * BR * 32 * EPL * sizeof(double);
    static bool attr_done[64] = {};
    int dev = 0;
    OPTEX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        OPTEX_CUDA(cudaFuncSetAttribute(pca_jacobi_blk_kernel<EPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    double *W = w.Wp;
    int cp = c_pad, ms = MAX_SWEEPS;
    unsigned *rot = w.rot, *rotmax = w.rotmax, *bar = w.bar;
    int *info = w.info;
    const double *stat = w.stat;
    float exit_r2 = jacobi_exit_r2();
    long long *stamps = g_pca_stamps;
    int n_stamps = g_pca_n_stamps;
    void *args[] = {&W, &cp, &ms, &rot, &rotmax, &exit_r2, &info, &stat, &bar, &stamps, &n_stamps};
    // cooperative launch: the c_pad / 16 CTAs (<= 64) spin on each other in grid_barrier and must be co-resident
    OPTEX_CUDA(cudaLaunchCooperativeKernel((const void *)pca_jacobi_blk_kernel<EPL>, dim3(c_pad / (2 * BR)),
                                           dim3(BR * 32), args, smem, st));
    count_launch();
    return OPTEX_OK;
}

// OPTEX_PCA_DUAL=0: blocks with fewer rows than channels are solved in the primal (c x c) form too
bool dual_enabled() {
    static const bool v = [] {
        const char *e = getenv("OPTEX_PCA_DUAL");
        return !(e && atoi(e) == 0);
    }();
    return v;
}

// OPTEX_PCA_SOLVER=0: the round-robin kernel (one grid barrier per round) also for cold solves
bool blocked_solver_enabled() {
    static const bool v = [] {
        const char *e = getenv("OPTEX_PCA_SOLVER");
        return !(e && atoi(e) == 0);
    }();
    return v;
}

}  // namespace
}  // namespace optex

using namespace optex;

// Debug hook (scripts/pca_stamps.py): device buffer of 5 * n clock64 stamps written by CTA 1 of the blocked-order
// solver (round start / rows loaded / sub-rounds done / rows stored / barrier passed); NULL switches it off.
extern "C" int optex_debug_pca_stamps(long long *device_buffer, int n) {
    g_pca_stamps = device_buffer;
    g_pca_n_stamps = device_buffer ? n : 0;
    return OPTEX_OK;
}

// n < c: the dual solve works on the transposed block (c rows, n columns) behind the primal layout
static size_t pca_dual_offset(int64_t n, int c) { return align_up(pca_layout(n, c, nullptr, nullptr, 0, nullptr), 256); }

extern "C" size_t optex_fit_pca_workspace_bytes(int64_t n, int c) {
    if (n < 1 || c < 1 || c > PCA_MAX_C) return 0;
    size_t bytes = pca_dual_offset(n, c);
    if (n < c)
        bytes += align_up((size_t)n * c * sizeof(float), 256) +
                 align_up(pca_layout(c, (int)n, nullptr, nullptr, 0, nullptr), 256);
    return bytes;
}

extern "C" int optex_fit_pca(const float *X, int64_t n, int c, float *eigvecs, float *sigma, int32_t *k_out,
                             void *workspace, size_t workspace_bytes, void *stream) {
    return optex_fit_pca_warm(X, n, c, eigvecs, sigma, k_out, nullptr, 0, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int optex_fit_pca_warm(const float *X, int64_t n, int c, float *eigvecs, float *sigma, int32_t *k_out,
                                  double *basis, int warm, int32_t *sweeps_out, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    if (!X || !eigvecs || !sigma || !k_out || n < 1 || c < 1) {
        set_error("optex_fit_pca: NULL pointer or empty shape");
        return OPTEX_EINVAL;
    }
    if (warm && !basis) {
        set_error("optex_fit_pca_warm: warm != 0 needs the basis of a previous solve");
        return OPTEX_EINVAL;
    }
    if (c > PCA_MAX_C) {
        set_error("optex_fit_pca: c = %d exceeds %d channels", c, PCA_MAX_C);
        return OPTEX_ESIZE;
    }
    PcaWs w;
    bool ok = false;
    pca_layout(n, c, &w, workspace, workspace_bytes, &ok);
    if (!workspace || !ok) {
        set_error("optex_fit_pca: workspace %zu < %zu bytes", workspace_bytes, optex_fit_pca_workspace_bytes(n, c));
        return OPTEX_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (!basis && blocked_solver_enabled() && n < c && n >= 2 && dual_enabled()) {
        // Fewer rows than channels (conv5_1 of a 256^2 pass: 384 x 512): G = A^T A has c - n zero eigenvalues whose
        // rows keep the Jacobi iteration busy (25 sweeps against 13).  Solve the n x n problem A A^T instead - full rank,
        // (n / 16) CTAs, n / 8 - 1 barriers per sweep - and map its eigenvectors through A^T (pca_dual_vecs_kernel).
        const int nd = (int)n;
        char *base = (char *)workspace + pca_dual_offset(n, c);
        float *Xt = (float *)base;
        base += align_up((size_t)n * c * sizeof(float), 256);
        const size_t left = workspace_bytes - (size_t)(base - (char *)workspace);
        PcaWs d;
        bool okd = false;
        pca_layout(c, nd, &d, base, left, &okd);
        if (!okd) {
            set_error("optex_fit_pca: workspace %zu < %zu bytes", workspace_bytes, optex_fit_pca_workspace_bytes(n, c));
            return OPTEX_EWORKSPACE;
        }
        OPTEX_TRY(transpose_f32(X, Xt, n, c, st));   // Xt [c, n]: the same pipeline on it yields X X^T, centred
        const int64_t rows_t = c;
        const int sp = (int)(rows_t < SUM_SPLITS ? rows_t : SUM_SPLITS);
        launch_pdl(pca_colsum_kernel, dim3(cdiv(nd, 32), sp), dim3(32, 8), 0, st, (const float *)Xt, d.colpart, rows_t,
                   nd, sp);
        OPTEX_LAUNCH_CHECK("pca_colsum_kernel");
        launch_pdl(pca_mean_kernel, dim3(1), dim3(1024), 0, st, (const double *)d.colpart, sp, nd, rows_t, d.s, d.stat);
        OPTEX_LAUNCH_CHECK("pca_mean_kernel");
        const int Td = (nd + GT - 1) / GT;
        int gsd = gram_splits(rows_t, nd);
        int64_t rws = ((rows_t + gsd - 1) / gsd + GK - 1) / GK * GK;
        gsd = (int)((rows_t + rws - 1) / rws);
        launch_pdl(pca_gram_kernel, dim3(Td * (Td + 1) / 2, gsd), dim3(256), 0, st, (const float *)Xt, d.gpart, rows_t,
                   nd, Td, rws);
        OPTEX_LAUNCH_CHECK("pca_gram_kernel");
        OPTEX_CUDA(cudaMemsetAsync(d.rot, 0, sizeof(unsigned) * 2 * (MAX_SWEEPS + 8), st));
        launch_pdl(pca_gram_final_kernel, dim3(cdiv((int64_t)nd * nd, 256)), dim3(256), 0, st, (const double *)d.gpart,
                   gsd, (const double *)d.s, (const double *)d.stat, nd, d.W, (double *)nullptr);
        OPTEX_LAUNCH_CHECK("pca_gram_final_kernel");
        const int ldd = blk_ld(nd), padd = blk_rows(nd);
        OPTEX_CUDA(cudaMemsetAsync(d.bar, 0, sizeof(unsigned) * 8, st));
        pca_pad_kernel<<<cdiv((int64_t)padd * ldd, 256 * 8), 256, 0, st>>>((const double *)d.W, nd, padd, ldd, d.Wp, d.stat);
        OPTEX_LAUNCH_CHECK("pca_pad_kernel");
        switch (ldd) {
            case 64: OPTEX_TRY(launch_jacobi_blk<2>(d, nd, st)); break;
            case 128: OPTEX_TRY(launch_jacobi_blk<4>(d, nd, st)); break;
            case 256: OPTEX_TRY(launch_jacobi_blk<8>(d, nd, st)); break;
            case 512: OPTEX_TRY(launch_jacobi_blk<16>(d, nd, st)); break;
            default: OPTEX_TRY(launch_jacobi_blk<32>(d, nd, st)); break;
        }
        OPTEX_CUDA(cudaMemsetAsync(sigma, 0, sizeof(float) * c, st));             // sigma[n ..] = 0
        OPTEX_CUDA(cudaMemsetAsync(eigvecs, 0, sizeof(float) * (size_t)c * c, st));   // columns n .. stay zero
        pca_finish_rows_kernel<<<1, 1024, 0, st>>>((const double *)d.Wp, ldd, nd, sigma, k_out, d.perm, d.scale);
        OPTEX_LAUNCH_CHECK("pca_finish_rows_kernel");
        pca_dual_vecs_kernel<<<nd, 256, 0, st>>>(X, n, c, (const double *)d.Wp, ldd, (const int *)d.perm,
                                                 (const double *)d.scale, (const double *)d.stat, eigvecs);
        OPTEX_LAUNCH_CHECK("pca_dual_vecs_kernel");
        if (sweeps_out)
            OPTEX_CUDA(cudaMemcpyAsync(sweeps_out, d.info, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        return OPTEX_OK;
    }
    const int splits = (int)(n < SUM_SPLITS ? n : SUM_SPLITS);
    launch_pdl(pca_colsum_kernel, dim3(cdiv(c, 32), splits), dim3(32, 8), 0, st, X, w.colpart, n, c, splits);
    OPTEX_LAUNCH_CHECK("pca_colsum_kernel");
    launch_pdl(pca_mean_kernel, dim3(1), dim3(1024), 0, st, (const double *)w.colpart, splits, c, n, w.s, w.stat);
    OPTEX_LAUNCH_CHECK("pca_mean_kernel");
    const int T = (c + GT - 1) / GT;
    int gs = gram_splits(n, c);
    int64_t rows = ((n + gs - 1) / gs + GK - 1) / GK * GK;
    gs = (int)((n + rows - 1) / rows);
    launch_pdl(pca_gram_kernel, dim3(T * (T + 1) / 2, gs), dim3(256), 0, st, X, w.gpart, n, c, T, rows);
    OPTEX_LAUNCH_CHECK("pca_gram_kernel");
    OPTEX_CUDA(cudaMemsetAsync(w.rot, 0, sizeof(unsigned) * 2 * (MAX_SWEEPS + 8), st));
    if (!basis && blocked_solver_enabled()) {
        // cold solve without a basis to hand back: blocked-order Jacobi on the padded copy of G, no V
        launch_pdl(pca_gram_final_kernel, dim3(cdiv((int64_t)c * c, 256)), dim3(256), 0, st, (const double *)w.gpart,
                   gs, (const double *)w.s, (const double *)w.stat, c, w.W, (double *)nullptr);
        OPTEX_LAUNCH_CHECK("pca_gram_final_kernel");
        const int ld = blk_ld(c), c_pad = blk_rows(c);
        OPTEX_CUDA(cudaMemsetAsync(w.bar, 0, sizeof(unsigned) * 8, st));
        pca_pad_kernel<<<cdiv((int64_t)c_pad * ld, 256 * 8), 256, 0, st>>>((const double *)w.W, c, c_pad, ld, w.Wp,
                                                                             w.stat);
        OPTEX_LAUNCH_CHECK("pca_pad_kernel");
        switch (ld) {
            case 64: OPTEX_TRY(launch_jacobi_blk<2>(w, c, st)); break;
            case 128: OPTEX_TRY(launch_jacobi_blk<4>(w, c, st)); break;
            case 256: OPTEX_TRY(launch_jacobi_blk<8>(w, c, st)); break;
            case 512: OPTEX_TRY(launch_jacobi_blk<16>(w, c, st)); break;
            default: OPTEX_TRY(launch_jacobi_blk<32>(w, c, st)); break;
        }
        pca_finish_rows_kernel<<<1, 1024, 0, st>>>((const double *)w.Wp, ld, c, sigma, k_out, w.perm, w.scale);
        OPTEX_LAUNCH_CHECK("pca_finish_rows_kernel");
        launch_pdl(pca_vecs_rows_kernel, dim3(cdiv(c, 32), cdiv(c, 32)), dim3(32, 8), 0, st, (const double *)w.Wp, ld,
                   (const int *)w.perm, (const double *)w.scale, c, eigvecs);
        OPTEX_LAUNCH_CHECK("pca_vecs_rows_kernel");
    } else {
        // the solver works directly in the caller's basis buffer when there is one (V is both the start and the result)
        double *G = w.V;  // warm: the workspace's V slot is free and holds G until W = V G is formed
        if (basis) w.V = basis;
        launch_pdl(pca_gram_final_kernel, dim3(cdiv((int64_t)c * c, 256)), dim3(256), 0, st, (const double *)w.gpart,
                   gs, (const double *)w.s, (const double *)w.stat, c, warm ? G : w.W,
                   warm ? (double *)nullptr : w.V);
        OPTEX_LAUNCH_CHECK("pca_gram_final_kernel");
        if (warm) {
            launch_pdl(pca_warm_kernel, dim3(T, T), dim3(256), 0, st, (const double *)w.V, (const double *)G, w.W, c);
            OPTEX_LAUNCH_CHECK("pca_warm_kernel");
        }
        if (c <= JT)
            OPTEX_TRY(launch_jacobi<1>(w, warm ? G : w.W, c, st));
        else if (c <= 2 * JT)
            OPTEX_TRY(launch_jacobi<2>(w, warm ? G : w.W, c, st));
        else if (c <= 4 * JT)
            OPTEX_TRY(launch_jacobi<4>(w, warm ? G : w.W, c, st));
        else
            OPTEX_TRY(launch_jacobi<8>(w, warm ? G : w.W, c, st));
        pca_finish_kernel<<<1, 1024, 0, st>>>(w.W, w.V, c, sigma, k_out, w.perm, w.sign);
        OPTEX_LAUNCH_CHECK("pca_finish_kernel");
        launch_pdl(pca_vecs_kernel, dim3(cdiv(c, 32), cdiv(c, 32)), dim3(32, 8), 0, st, (const double *)w.V,
                   (const int *)w.perm, (const float *)w.sign, c, eigvecs);
        OPTEX_LAUNCH_CHECK("pca_vecs_kernel");
    }
    if (sweeps_out)
        OPTEX_CUDA(cudaMemcpyAsync(sweeps_out, w.info, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    return OPTEX_OK;
}

// out[n, k] = X[n, c] V[c, k]            (transpose = 0;  optex.py:110, :188)
// out[n, c] = F[n, k] V[c, k]^T          (transpose = 1;  optex.py:120)
extern "C" int optex_pca_project(const float *X, const float *V, float *out, int64_t n, int c, int k, int transpose,
                                 void *stream) {
    OPTEX_TRY(require_sm100());
    if (!X || !V || !out || n < 1 || c < 1 || k < 1) {
        set_error("optex_pca_project: NULL pointer or empty shape");
        return OPTEX_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int mode = optex_get_gemm_mode();
    if (mode != OPTEX_GEMM_FP32) {
        TcGemm g{};
        g.A = X;
        g.a_mn = false;
        g.B = V;
        g.D = out;
        g.M = n;
        g.alpha = 1.f;
        g.terms = mode == OPTEX_GEMM_TF32 ? 1 : 3;
        if (!transpose) {
            g.b_mn = true;  // V row-major [K = c, N = k]
            g.N = k;
            g.K = c;
            g.ldd = k;
        } else {
            g.b_mn = false;  // V row-major [N = c, K = k]
            g.N = c;
            g.K = k;
            g.ldd = c;
        }
        int rc = gemm_tc(g, st);
        if (rc != OPTEX_ENOTSUP) return rc;
        if (mode == OPTEX_GEMM_TF32X3 || mode == OPTEX_GEMM_TF32) {
            set_error("optex_pca_project: shape (n=%lld, c=%d, k=%d) is outside the tensor-core path's TMA constraints; "
                      "use gemm mode auto or fp32", (long long)n, c, k);
            return OPTEX_ESIZE;
        }
    }
    if (!transpose) return sgemm_simt(X, c, true, V, k, false, out, k, false, n, k, c, nullptr, 0.f, 1.f, st);
    return sgemm_simt(X, k, true, V, k, true, out, c, false, n, c, k, nullptr, 0.f, 1.f, st);
}
