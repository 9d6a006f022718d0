// Haar-random SO(c) rotations generated on the device.
//
//   reference: random_rotation()  optex.py:142-164.  Its live branch calls scipy on the
//   host every iteration (35 ms at c = 512 + an H2D copy); this is the reference's own
//   impl="torch" branch (Stewart's Householder construction, optex.py:151-164) restated
//   for the GPU, in fp64, result stored as fp32 like `.to(pastiche_feature)` (optex.py:168).
//
// Parity: distributional only (the reference draws from numpy's global RNG); with the
// normals injected (`gauss`) it matches oracle/rotation.py::haar_rotation_householder.
//
// Two kernels for a whole BATCH of rotations (one per OT iteration of a layer):
//   rot_vectors : one warp per reflection n - normals -> normalised Householder vector v_n
//   rot_apply   : one warp per row of H - the row stays in registers while the c-1
//                 reflections H[i, n:] -= (H[i, n:] . v_n) v_n are applied in sequence
//                 (rows are independent, so no inter-CTA synchronisation is needed).
#include "common.cuh"

namespace optex {
namespace {

constexpr int MAX_C = 1024;
constexpr int PER_LANE = MAX_C / 32;

// ---- Philox4x32-10 ------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint64_t key) {
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
// Standard normal for element (n, k) of rotation `counter` under `seed` (Box-Muller, fp64).
__device__ __forceinline__ double philox_normal(uint64_t seed, uint64_t counter, uint32_t n, uint32_t k) {
    uint32_t c[4] = {k >> 1, n, (uint32_t)counter, (uint32_t)(counter >> 32)};
    philox4x32_10(c, seed);
    double u1 = ((double)c[0] + 0.5) * (1.0 / 4294967296.0);
    double u2 = ((double)c[1] + 0.5) * (1.0 / 4294967296.0);
    double r = sqrt(-2.0 * log(u1));
    double s, co;
    sincospi(2.0 * u2, &s, &co);
    return (k & 1) ? r * s : r * co;
}

// v[b][n][k] (k >= n; zero for k < n) and signs d[b][n]     optex.py:154-160
__global__ void rot_vectors_kernel(double *__restrict__ V, double *__restrict__ D, int c, uint64_t seed,
                                   uint64_t first_counter, const double *__restrict__ gauss) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (n >= c - 1) return;
    double *v = V + ((int64_t)b * (c - 1) + n) * c;
    double x[PER_LANE];
    double norm2 = 0.0;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
        int k = i * 32 + lane;
        double g = 0.0;
        if (k >= n && k < c)
            g = gauss ? gauss[((int64_t)b * (c - 1) + n) * c + k]
                      : philox_normal(seed, first_counter + b, (uint32_t)n, (uint32_t)k);
        x[i] = g;
        norm2 += g * g;
    }
    norm2 = warp_sum(norm2);
    // x0 = x[n]
    double x0 = __shfl_sync(0xffffffffu, x[0], n & 31);
#pragma unroll
    for (int i = 1; i < PER_LANE; ++i) {
        double cand = __shfl_sync(0xffffffffu, x[i], n & 31);
        if ((n >> 5) == i) x0 = cand;
    }
    const double dn = x0 >= 0.0 ? 1.0 : -1.0;  // sign(sign(x0) + 0.5)
    const double x0n = x0 + dn * sqrt(norm2);
    const double scale = 1.0 / sqrt((norm2 - x0 * x0 + x0n * x0n) / 2.0);
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) {
        int k = i * 32 + lane;
        if (k < c) v[k] = (k == n ? x0n : x[i]) * scale;
    }
    if (lane == 0) D[(int64_t)b * c + n] = dn;
}

// R[b][i][:] = D[i] * (e_i * prod_n (I - v_n v_n^T))        optex.py:153,161-163
template <int PL>
__global__ void __launch_bounds__(128)
rot_apply_kernel(const double *__restrict__ V, const double *__restrict__ D, float *__restrict__ R, int c) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (row >= c) return;
    const double *Vb = V + (int64_t)b * (c - 1) * c;
    double h[PL];
#pragma unroll
    for (int i = 0; i < PL; ++i) h[i] = (i * 32 + lane == row) ? 1.0 : 0.0;
    double vn[PL], vnext[PL];
#pragma unroll
    for (int i = 0; i < PL; ++i) {
        int k = i * 32 + lane;
        vnext[i] = (c > 1 && k < c) ? Vb[k] : 0.0;
    }
    for (int n = 0; n < c - 1; ++n) {
#pragma unroll
        for (int i = 0; i < PL; ++i) vn[i] = vnext[i];
        if (n + 1 < c - 1) {
            const double *vp = Vb + (int64_t)(n + 1) * c;
#pragma unroll
            for (int i = 0; i < PL; ++i) {
                int k = i * 32 + lane;
                vnext[i] = k < c ? vp[k] : 0.0;
            }
        }
        double dot = 0.0;
#pragma unroll
        for (int i = 0; i < PL; ++i) dot = fma(h[i], vn[i], dot);
        dot = warp_sum(dot);
#pragma unroll
        for (int i = 0; i < PL; ++i) h[i] = fma(-dot, vn[i], h[i]);
    }
    // D[-1] = (-1)^(c-1) * prod(D[:-1])                     optex.py:162
    double d;
    if (row < c - 1) {
        d = D[(int64_t)b * c + row];
    } else {
        double p = 1.0;
        for (int k = lane; k < c - 1; k += 32) p *= D[(int64_t)b * c + k];
#pragma unroll
        for (int o = 16; o; o >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, o);
        d = ((c - 1) & 1) ? -p : p;
    }
    float *out = R + ((int64_t)b * c + row) * c;
#pragma unroll
    for (int i = 0; i < PL; ++i) {
        int k = i * 32 + lane;
        if (k < c) out[k] = (float)(d * h[i]);
    }
}

}  // namespace

size_t rotation_ws_bytes(int c, int batch) {
    if (c <= 0 || batch <= 0) return 0;
    return align_up(sizeof(double) * (size_t)batch * (c > 1 ? c - 1 : 1) * c, 256) +
           align_up(sizeof(double) * (size_t)batch * c, 256);
}

int random_rotations(float *R, int c, int batch, uint64_t seed, uint64_t first_counter, const double *gauss,
                     void *workspace, size_t workspace_bytes, cudaStream_t st) {
    if (!R || c < 1 || batch < 1) {
        set_error("random_rotation: NULL output, c < 1 or batch < 1");
        return OPTEX_EINVAL;
    }
    if (c > MAX_C) {
        set_error("random_rotation: c=%d > %d", c, MAX_C);
        return OPTEX_ESIZE;
    }
    if (batch > 65535) {
        set_error("random_rotation: batch > 65535");
        return OPTEX_ESIZE;
    }
    Arena ar(workspace, workspace_bytes);
    double *V = ar.take<double>((size_t)batch * (c > 1 ? c - 1 : 1) * c);
    double *D = ar.take<double>((size_t)batch * c);
    if (!ar.ok()) {
        set_error("random_rotation: workspace %zu < %zu", workspace_bytes, rotation_ws_bytes(c, batch));
        return OPTEX_EWORKSPACE;
    }
    if (c > 1) {
        dim3 g1((unsigned)((c - 1 + 3) / 4), (unsigned)batch);
        rot_vectors_kernel<<<g1, 128, 0, st>>>(V, D, c, seed, first_counter, gauss);
        OPTEX_LAUNCH_CHECK("rot_vectors_kernel");
    }
    dim3 g2((unsigned)((c + 3) / 4), (unsigned)batch);
    int pl = (c + 31) / 32;
    if (pl <= 2) rot_apply_kernel<2><<<g2, 128, 0, st>>>(V, D, R, c);
    else if (pl <= 4) rot_apply_kernel<4><<<g2, 128, 0, st>>>(V, D, R, c);
    else if (pl <= 8) rot_apply_kernel<8><<<g2, 128, 0, st>>>(V, D, R, c);
    else if (pl <= 16) rot_apply_kernel<16><<<g2, 128, 0, st>>>(V, D, R, c);
    else rot_apply_kernel<32><<<g2, 128, 0, st>>>(V, D, R, c);
    OPTEX_LAUNCH_CHECK("rot_apply_kernel");
    return OPTEX_OK;
}

}  // namespace optex

using namespace optex;

extern "C" size_t optex_rotation_workspace_bytes(int c) { return rotation_ws_bytes(c, 1); }

extern "C" int optex_random_rotation(float *R, int c, uint64_t seed, uint64_t counter, const double *gauss,
                                     void *workspace, size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    return random_rotations(R, c, 1, seed, counter, gauss, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" size_t optex_rotations_workspace_bytes(int c, int count) { return rotation_ws_bytes(c, count); }

extern "C" int optex_random_rotations(float *R, int c, int count, uint64_t seed, uint64_t first_counter,
                                      const double *gauss, void *workspace, size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    return random_rotations(R, c, count, seed, first_counter, gauss, workspace, workspace_bytes, (cudaStream_t)stream);
}
