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
// Three kernels for a whole BATCH of rotations (one per OT iteration of a layer):
//   rot_vectors : one warp per reflection n - normals -> normalised Householder vector v_n (zero-padded rows)
//   rot_gram    : one warp per block of NB consecutive reflections - the NB(NB-1)/2 products v_a . v_b
//   rot_apply   : a warp keeps RW rows of H in registers and applies the reflections NB at a time in compact-WY
//                 form (rows are independent, so no inter-CTA synchronisation is needed).
#include <atomic>

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

constexpr int NB = 4;  // reflections per compact-WY block (divides 32, so a block never straddles a 32-column chunk)

// columns of the zero-padded V rows = the width PL * 32 of the rot_apply instantiation that serves c
__host__ __device__ inline int padded_cols(int c) { return c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : c <= 512 ? 512 : 1024; }
__host__ __device__ inline int padded_rows(int c) { return (c - 1 + NB - 1) / NB * NB + NB; }

// v[b][n][k] (k >= n; zero for k < n) and signs d[b][n]     optex.py:154-160
// V is stored zero-padded, [b][padded_rows(c)][padded_cols(c)], so that rot_apply needs no bounds checks.
template <typename T>
__global__ void rot_vectors_kernel(T *__restrict__ V, double *__restrict__ D, int c, uint64_t seed,
                                   uint64_t first_counter, const double *__restrict__ gauss) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    const int cp = padded_cols(c), rp = padded_rows(c);
    if (n >= rp) return;
    T *v = V + ((int64_t)b * rp + n) * cp;
    if (n >= c - 1) {  // padding rows: reflections that do not exist
        for (int k = lane; k < cp; k += 32) v[k] = (T)0;
        return;
    }
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
        if (k < cp) v[k] = (T)((k == n ? x0n : x[i]) * scale);   // x is zero outside [n, c)
    }
    if (lane == 0) D[(int64_t)b * c + n] = dn;
}

// G[b][block][a * NB + a2] = v_{n0+a} . v_{n0+a2}  (a < a2) for every block of NB consecutive reflections
template <typename T>
__global__ void rot_gram_kernel(const T *__restrict__ V, T *__restrict__ G, int c) {
    const int lane = threadIdx.x & 31;
    const int cp = padded_cols(c), rp = padded_rows(c);
    const int blk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (blk >= rp / NB) return;
    const T *v = V + ((int64_t)b * rp + (int64_t)blk * NB) * cp;
    double g[NB * NB];
#pragma unroll
    for (int j = 0; j < NB * NB; ++j) g[j] = 0.0;
    for (int k = lane; k < cp; k += 32) {
        double va[NB];
#pragma unroll
        for (int a = 0; a < NB; ++a) va[a] = (double)v[(int64_t)a * cp + k];
#pragma unroll
        for (int a = 0; a < NB; ++a)
#pragma unroll
            for (int a2 = a + 1; a2 < NB; ++a2) g[a * NB + a2] = fma(va[a], va[a2], g[a * NB + a2]);
    }
#pragma unroll
    for (int a = 0; a < NB; ++a)
#pragma unroll
        for (int a2 = a + 1; a2 < NB; ++a2) {
            const double t = warp_sum(g[a * NB + a2]);
            if (lane == 0) G[((int64_t)b * (rp / NB) + blk) * (NB * NB) + a * NB + a2] = (T)t;
        }
}

// R[b][i][:] = D[i] * (e_i * prod_n (I - v_n v_n^T))        optex.py:153,161-163
// A warp keeps RW rows of H in registers and streams the reflections past them NB at a time.  For a block
// v_0..v_{NB-1} and a row h:   d_a = h . v_a  (all NB dot products of the UNMODIFIED row, one combined reduction),
//   t_0 = d_0,  t_a = d_a - sum_{a' < a} t_a' (v_a' . v_a)      (what the sequential reflections would have seen)
//   h  -= sum_a t_a v_a
// i.e. NB times fewer latency-bound reduction chains than reflection-by-reflection, v fetched once per RW rows
// (one row per warp was bound by L1 bandwidth), and since v_n is zero below k = n the 32-column chunks left of the
// block are skipped (the phase loop is unrolled, so the chunk loops have constant bounds) - half of the FP64 work.
constexpr int ROT_STAGES = 4;  // shared-memory ring of v blocks (cp.async)

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// first 32-column chunk that block j of NB reflections can touch, rounded down to the phase granularity PS
template <int PS>
__device__ __forceinline__ int block_chunk0(int j) { return ((j * NB) >> 5) / PS * PS; }

// stage block j (rows j*NB .. +NB-1 of V, columns from its phase's first chunk on) into ring slot j % ROT_STAGES
template <typename T, int PL, int PS>
__device__ __forceinline__ void rot_fetch(T *ring, const T *__restrict__ Vg, int j, int nblk, int tid) {
    constexpr int cp = PL * 32;
    if (j < nblk) {
        const int col0 = block_chunk0<PS>(j) * 32;
        constexpr int E = 16 / sizeof(T);       // elements per 16-byte piece
        const int per_row = (cp - col0) / E;    // pieces per row
        T *dst = ring + (size_t)(j % ROT_STAGES) * NB * cp;
        const T *src = Vg + (int64_t)j * NB * cp;
#pragma unroll
        for (int a = 0; a < NB; ++a)
            for (int q = tid; q < per_row; q += 128) cp_async16(dst + a * cp + col0 + E * q, src + a * cp + col0 + E * q);
    }
    cp_async_commit();  // one group per block index, empty or not, so that wait_group counts stay uniform
}

// The reflections whose first column lies in chunks [I0, I0 + PS) of 32 columns, NB at a time; the chunk loops
// start at the CONSTANT I0 (v_n is zero left of n) and carry no bounds checks (V is zero-padded to PL * 32 columns),
// so they are straight-line code.  PS trades skipped work for code size: one loop body per phase.
template <typename T, int PL, int RW, int PS, int I0>
__device__ __forceinline__ void rot_phase(T (&h)[RW][PL], T *ring, const T *__restrict__ Vg,
                                          const T *__restrict__ Gb, int c, int nblk, int lane, int tid) {
    if constexpr (I0 < PL) {
        constexpr int cp = PL * 32;
        if (I0 * 32 >= c - 1) return;
        for (int n0 = I0 * 32; n0 < (I0 + PS) * 32 && n0 < c - 1; n0 += NB) {
            const int j = n0 / NB;
            // v_a . v_a' of this block (needed only after the reduction: the latency hides behind the dot pass)
            const T *g = Gb + (int64_t)j * (NB * NB);
            T gv[NB * NB];
#pragma unroll
            for (int a = 1; a < NB; ++a)
#pragma unroll
                for (int a2 = 0; a2 < a; ++a2) gv[a2 * NB + a] = __ldg(g + a2 * NB + a);
            cp_async_wait<ROT_STAGES - 2>();   // block j has landed (this thread's copies) ...
            __syncthreads();                   // ... and everybody's; everybody is also done with block j - 1
            rot_fetch<T, PL, PS>(ring, Vg, j + ROT_STAGES - 1, nblk, tid);
            const T *vb = ring + (size_t)(j % ROT_STAGES) * NB * cp + lane;
            constexpr int NV = RW * NB;
            T d[NV];  // d[r * NB + a]
#pragma unroll
            for (int q = 0; q < NV; ++q) d[q] = (T)0;
#pragma unroll
            for (int i = I0; i < PL; ++i) {
                // keep at most two chunks of v loads in flight (unrestrained hoisting spills the h registers)
                if (((i - I0) & 1) == 0) asm volatile("" ::: "memory");
                T va[NB];
#pragma unroll
                for (int a = 0; a < NB; ++a) va[a] = vb[a * cp + i * 32];
#pragma unroll
                for (int r = 0; r < RW; ++r)
#pragma unroll
                    for (int a = 0; a < NB; ++a) d[r * NB + a] = fma(h[r][i], va[a], d[r * NB + a]);
            }
            // transpose-reduce: halve the number of live values at every butterfly level; value q ends up (summed
            // over the warp) in the lanes with  lane * NV / 32 == q
#pragma unroll
            for (int w = NV / 2, o = 16; w >= 1; w >>= 1, o >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int q = 0; q < w; ++q) {
                    const T send = up ? d[q] : d[q + w];
                    const T keep = up ? d[q + w] : d[q];
                    d[q] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                }
            }
#pragma unroll
            for (int o = 16 / NV; o >= 1; o >>= 1) d[0] += __shfl_xor_sync(0xffffffffu, d[0], o);
            const T mine = d[0];
            T t[NV];
#pragma unroll
            for (int q = 0; q < NV; ++q) t[q] = __shfl_sync(0xffffffffu, mine, q * (32 / NV));
            // what the sequential reflections would have seen
#pragma unroll
            for (int a = 1; a < NB; ++a)
#pragma unroll
                for (int a2 = 0; a2 < a; ++a2)
#pragma unroll
                    for (int r = 0; r < RW; ++r) t[r * NB + a] = fma(-t[r * NB + a2], gv[a2 * NB + a], t[r * NB + a]);
#pragma unroll
            for (int i = I0; i < PL; ++i) {
                // keep at most two chunks of v loads in flight (unrestrained hoisting spills the h registers)
                if (((i - I0) & 1) == 0) asm volatile("" ::: "memory");
                T va[NB];
#pragma unroll
                for (int a = 0; a < NB; ++a) va[a] = vb[a * cp + i * 32];
#pragma unroll
                for (int r = 0; r < RW; ++r)
#pragma unroll
                    for (int a = 0; a < NB; ++a) h[r][i] = fma(-t[r * NB + a], va[a], h[r][i]);
            }
        }
        rot_phase<T, PL, RW, PS, I0 + PS>(h, ring, Vg, Gb, c, nblk, lane, tid);
    }
}

template <typename T, int PL>
struct RotCfg {
    static constexpr int PS = PL >= 16 ? PL / 4 : (PL >= 4 ? PL / 2 : PL);
    static constexpr size_t SMEM = (size_t)ROT_STAGES * NB * PL * 32 * sizeof(T);
};

template <typename T, int PL, int RW>
__global__ void __launch_bounds__(128)
rot_apply_kernel(const T *__restrict__ V, const T *__restrict__ G, const double *__restrict__ D,
                 float *__restrict__ R, int c) {
    static_assert(RW * NB == 32 || RW * NB == 16 || RW * NB == 8 || RW * NB == 4,
                  "the transpose-reduce wants a power of two <= 32");
    extern __shared__ __align__(16) unsigned char rot_ring_raw[];
    T *rot_ring = reinterpret_cast<T *>(rot_ring_raw);
    const int tid = threadIdx.x, lane = tid & 31;
    // no early exit: every warp of the CTA takes part in the cp.async ring and its barriers (rows >= c are not stored)
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + (tid >> 5)) * RW;
    const int b = blockIdx.y;
    constexpr int cp = PL * 32;
    constexpr int PS = RotCfg<T, PL>::PS;
    const int rp = padded_rows(c);
    const int nblk = (c - 1 + NB - 1) / NB;
    const T *Vg = V + (int64_t)b * rp * cp;
    const T *Gb = G + (int64_t)b * (rp / NB) * (NB * NB);
    T h[RW][PL];
#pragma unroll
    for (int r = 0; r < RW; ++r)
#pragma unroll
        for (int i = 0; i < PL; ++i) h[r][i] = (i * 32 + lane == row0 + r) ? (T)1 : (T)0;
#pragma unroll
    for (int j = 0; j < ROT_STAGES - 1; ++j) rot_fetch<T, PL, PS>(rot_ring, Vg, j, nblk, tid);
    rot_phase<T, PL, RW, PS, 0>(h, rot_ring, Vg, Gb, c, nblk, lane, tid);
    cp_async_wait<0>();
    // D[-1] = (-1)^(c-1) * prod(D[:-1])                     optex.py:162
    double plast = 1.0;
    if (row0 + RW > c - 1) {
        for (int k = lane; k < c - 1; k += 32) plast *= D[(int64_t)b * c + k];
#pragma unroll
        for (int o = 16; o; o >>= 1) plast *= __shfl_xor_sync(0xffffffffu, plast, o);
        plast = ((c - 1) & 1) ? -plast : plast;
    }
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        const int row = row0 + r;
        if (row >= c) break;
        const double dd = row < c - 1 ? D[(int64_t)b * c + row] : plast;
        float *out = R + ((int64_t)b * c + row) * c;
#pragma unroll
        for (int i = 0; i < PL; ++i) {
            const int k = i * 32 + lane;
            if (k < c) out[k] = (float)(dd * (double)h[r][i]);
        }
    }
}

template <typename T, int PL, int RW>
int launch_rot_apply(const T *V, const T *G, const double *D, float *R, int c, int batch, cudaStream_t st) {
    auto kern = rot_apply_kernel<T, PL, RW>;
    static PerDeviceOnce attr_once1;
    OPTEX_TRY(ensure_dyn_smem(attr_once1, kern, (int)RotCfg<T, PL>::SMEM));
    dim3 g((unsigned)((c + 4 * RW - 1) / (4 * RW)), (unsigned)batch);
    kern<<<g, 128, RotCfg<T, PL>::SMEM, st>>>(V, G, D, R, c);
    return OPTEX_OK;
}

}  // namespace

// 1 = fp64 arithmetic (default; what the reference's live scipy branch computes in before `.to(pastiche_feature)`
// rounds the result to fp32): |R R^T - I| ~ 1e-8.  0 = fp32 (what its impl="torch" branch computes in,
// optex.py:150-164): 2x faster (the FP64 pipe is the bound), |R R^T - I| ~ 1e-6 at c = 512.
std::atomic<int> g_rot_fp64{1};

size_t rotation_ws_bytes(int c, int batch) {
    if (c <= 0 || batch <= 0) return 0;
    const size_t rp = padded_rows(c), cp = padded_cols(c);
    return align_up(sizeof(double) * (size_t)batch * rp * cp, 256) + align_up(sizeof(double) * (size_t)batch * c, 256) +
           align_up(sizeof(double) * (size_t)batch * (rp / NB) * NB * NB, 256);
}

namespace {
template <typename T>
int random_rotations_t(float *R, int c, int batch, uint64_t seed, uint64_t first_counter, const double *gauss,
                       Arena &ar, cudaStream_t st) {
    const int rp = padded_rows(c), cp = padded_cols(c);
    T *V = ar.take<T>((size_t)batch * rp * cp);
    double *D = ar.take<double>((size_t)batch * c);
    T *G = ar.take<T>((size_t)batch * (rp / NB) * NB * NB);
    if (!ar.ok()) {
        set_error("random_rotation: workspace too small (need %zu bytes)", rotation_ws_bytes(c, batch));
        return OPTEX_EWORKSPACE;
    }
    dim3 g1((unsigned)((rp + 3) / 4), (unsigned)batch);
    rot_vectors_kernel<T><<<g1, 128, 0, st>>>(V, D, c, seed, first_counter, gauss);
    OPTEX_LAUNCH_CHECK("rot_vectors_kernel");
    dim3 g3((unsigned)((rp / NB + 3) / 4), (unsigned)batch);
    rot_gram_kernel<T><<<g3, 128, 0, st>>>(V, G, c);
    OPTEX_LAUNCH_CHECK("rot_gram_kernel");
    const int pl = (c + 31) / 32;
    // rows per warp: many amortise the v fetches, fewer put a small batch on more SMs (latency of a single draw);
    // RWMAX is what the register file takes next to the accumulators (fp32 rows are half the size)
    constexpr int RWMAX = sizeof(T) == 4 ? 8 : 4;
    const int64_t want = sm_count();
    int rw = 1;
    for (int cand : {8, 4, 2})
        if (cand <= RWMAX && (int64_t)((c + 4 * cand - 1) / (4 * cand)) * batch >= want) { rw = cand; break; }
#define OPTEX_ROT(PL_)                                                                                         \
    do {                                                                                                       \
        constexpr int RW_HI = (PL_ <= 16 ? RWMAX : RWMAX / 2);                                                 \
        if (rw >= RW_HI) OPTEX_TRY((launch_rot_apply<T, PL_, RW_HI>(V, G, D, R, c, batch, st)));               \
        else if (rw >= 2) OPTEX_TRY((launch_rot_apply<T, PL_, 2>(V, G, D, R, c, batch, st)));                  \
        else OPTEX_TRY((launch_rot_apply<T, PL_, 1>(V, G, D, R, c, batch, st)));                               \
    } while (0)
    if (pl <= 2) OPTEX_ROT(2);
    else if (pl <= 4) OPTEX_ROT(4);
    else if (pl <= 8) OPTEX_ROT(8);
    else if (pl <= 16) OPTEX_ROT(16);
    else OPTEX_ROT(32);
#undef OPTEX_ROT
    OPTEX_LAUNCH_CHECK("rot_apply_kernel");
    return OPTEX_OK;
}
}  // namespace

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
    if (workspace_bytes < rotation_ws_bytes(c, batch)) {
        set_error("random_rotation: workspace %zu < %zu", workspace_bytes, rotation_ws_bytes(c, batch));
        return OPTEX_EWORKSPACE;
    }
    Arena ar(workspace, workspace_bytes);
    if (g_rot_fp64.load()) return random_rotations_t<double>(R, c, batch, seed, first_counter, gauss, ar, st);
    return random_rotations_t<float>(R, c, batch, seed, first_counter, gauss, ar, st);
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

extern "C" int optex_set_rotation_precision(int fp64) {
    const int prev = g_rot_fp64.exchange(fp64 ? 1 : 0);
    return prev;
}

extern "C" int optex_get_rotation_precision(void) { return g_rot_fp64.load(); }
