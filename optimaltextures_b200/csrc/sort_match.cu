// Exact 1-D optimal transport per rotated channel (the north-star's "sort" mode).
//
// NOT in the reference (its per-channel matcher is the 256-bin cdf_match,
// histmatch.py:49-69); defined by oracle/sort_oracle.py:
//     idx = stable argsort(target_c);  ss = sort(source_c)
//     out[idx[r]] = ss[((2r+1)*n_s) / (2*n_t)]
//
// One CTA per channel.  The whole channel lives in REGISTERS for the duration of
// the sort (512 threads x up to 32 elements): a bitonic network whose
// compare-exchange distance selects the exchange medium -
//     distance <  E        : between registers of one thread (no memory traffic)
//     distance <  32*E     : warp shuffles
//     distance >= 32*E     : one bank-conflict-free round trip through shared memory
// The target is sorted as 64-bit (ordered-key << 32 | pixel index) words, which makes
// the network's result identical to a STABLE sort (ties broken by index, -0.0 == +0.0,
// NaN last - torch.sort semantics), so the permutation is bit-exact.
#include "common.cuh"

namespace optex {
namespace {

constexpr int NT = 512;
constexpr int MAX_LOG_E = 5;  // 512 * 32 = 16384 elements per channel in registers

__device__ __forceinline__ uint32_t sort_key(float x) {
    if (x != x) return 0xffc00000u;  // canonical NaN sorts after +inf
    return f2ord(x + 0.0f);          // -0.0 + 0.0 == +0.0 : ties with +0.0 like torch
}

template <typename T>
__device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <typename T>
__device__ __forceinline__ T tmax(T a, T b) { return a < b ? b : a; }

template <int LOG_E, int J, typename T>
__device__ __forceinline__ void reg_stage(T (&v)[1 << LOG_E], int k, int gbase) {
    constexpr int E = 1 << LOG_E;
#pragma unroll
    for (int r = 0; r < E; ++r) {
        if ((r & J) == 0) {
            bool up = (((gbase | r) & k) == 0);
            T a = v[r], b = v[r | J];
            T lo = tmin(a, b), hi = tmax(a, b);
            v[r] = up ? lo : hi;
            v[r | J] = up ? hi : lo;
        }
    }
}

// Sort N2 = NT << LOG_E elements held as v[r] at global position (tid << LOG_E) | r, ascending.
// xbuf: shared scratch of N2 elements of T.
template <int LOG_E, typename T>
__device__ __forceinline__ void bitonic_sort(T (&v)[1 << LOG_E], T *xbuf) {
    constexpr int E = 1 << LOG_E;
    constexpr int N2 = NT * E;
    const int tid = threadIdx.x;
    const int gbase = tid << LOG_E;
#pragma unroll 1
    for (int k = 2; k <= N2; k <<= 1) {
        int j = k >> 1;
        const bool up = ((gbase & k) == 0);  // valid whenever k >= E (bit k is above the register bits)
#pragma unroll 1
        for (; j >= 32 * E; j >>= 1) {
            const int partner = tid ^ (j >> LOG_E);
            const bool keep_min = (((gbase & j) == 0) == up);
#pragma unroll
            for (int r = 0; r < E; ++r) xbuf[r * NT + tid] = v[r];
            __syncthreads();
#pragma unroll
            for (int r = 0; r < E; ++r) {
                T o = xbuf[r * NT + partner];
                v[r] = keep_min ? tmin(v[r], o) : tmax(v[r], o);
            }
            __syncthreads();
        }
#pragma unroll 1
        for (; j >= E; j >>= 1) {
            const int lm = j >> LOG_E;
            const bool keep_min = (((gbase & j) == 0) == up);
#pragma unroll
            for (int r = 0; r < E; ++r) {
                T o = __shfl_xor_sync(0xffffffffu, v[r], lm);
                v[r] = keep_min ? tmin(v[r], o) : tmax(v[r], o);
            }
        }
        if (LOG_E >= 5) { if (j >= 16) { reg_stage<LOG_E, (LOG_E >= 5 ? 16 : 0)>(v, k, gbase); j >>= 1; } }
        if (LOG_E >= 4) { if (j >= 8) { reg_stage<LOG_E, (LOG_E >= 4 ? 8 : 0)>(v, k, gbase); j >>= 1; } }
        if (LOG_E >= 3) { if (j >= 4) { reg_stage<LOG_E, (LOG_E >= 3 ? 4 : 0)>(v, k, gbase); j >>= 1; } }
        if (LOG_E >= 2) { if (j >= 2) { reg_stage<LOG_E, (LOG_E >= 2 ? 2 : 0)>(v, k, gbase); j >>= 1; } }
        if (LOG_E >= 1) { if (j >= 1) { reg_stage<LOG_E, (LOG_E >= 1 ? 1 : 0)>(v, k, gbase); } }
    }
}

template <int LOG_E>
__global__ void __launch_bounds__(NT, 1)
sort_match_kernel(const float *target, const float *__restrict__ source, float *out, int64_t n_t,
                  int64_t n_s, int32_t *__restrict__ perm) {
    constexpr int E = 1 << LOG_E;
    constexpr int N2 = NT * E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *xbuf = reinterpret_cast<uint64_t *>(smem_raw);                 // N2 x 8 B
    uint32_t *ss = reinterpret_cast<uint32_t *>(smem_raw + (size_t)N2 * 8);  // N2 x 4 B
    const int tid = threadIdx.x;
    const int ch = blockIdx.x;
    const int gbase = tid << LOG_E;

    {   // ---- ascending sort of the source channel (keys only)
        const float *srow = source + (int64_t)ch * n_s;
        uint32_t v[E];
#pragma unroll
        for (int r = 0; r < E; ++r) {
            int i = r * NT + tid;
            v[r] = i < n_s ? sort_key(__ldg(srow + i)) : 0xffffffffu;
        }
        bitonic_sort<LOG_E, uint32_t>(v, reinterpret_cast<uint32_t *>(xbuf));
#pragma unroll
        for (int r = 0; r < E; ++r) ss[gbase + r] = v[r];
    }
    // ---- stable argsort of the target channel
    const float *trow = target + (int64_t)ch * n_t;
    uint64_t v[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        int i = r * NT + tid;
        v[r] = i < n_t ? ((uint64_t)sort_key(trow[i]) << 32) | (uint32_t)i : ~0ull;
    }
    bitonic_sort<LOG_E, uint64_t>(v, xbuf);
    __syncthreads();  // ss visible; xbuf free
    // ---- rank r of the target receives the mid-point quantile of the sorted source
    float *stage = reinterpret_cast<float *>(xbuf);
#pragma unroll
    for (int r = 0; r < E; ++r) {
        int g = gbase + r;
        if (g < n_t) {
            uint32_t idx = (uint32_t)v[r];
            int64_t q = ((2 * (int64_t)g + 1) * n_s) / (2 * n_t);
            stage[idx] = ord2f(ss[q]);
            if (perm) perm[(int64_t)ch * n_t + g] = (int32_t)idx;
        }
    }
    __syncthreads();
    float *orow = out + (int64_t)ch * n_t;
    for (int i = tid; i < n_t; i += NT) orow[i] = stage[i];
}

template <int LOG_E>
int launch_sort(const float *t, const float *s, float *out, int c, int64_t n_t, int64_t n_s, int32_t *perm,
                cudaStream_t st) {
    size_t smem = (size_t)(NT << LOG_E) * 12;
    static bool attr_done = false;
    if (!attr_done) {
        OPTEX_CUDA(cudaFuncSetAttribute(sort_match_kernel<LOG_E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
        attr_done = true;
    }
    sort_match_kernel<LOG_E><<<c, NT, smem, st>>>(t, s, out, n_t, n_s, perm);
    OPTEX_LAUNCH_CHECK("sort_match_kernel");
    return OPTEX_OK;
}

}  // namespace
}  // namespace optex

using namespace optex;

extern "C" size_t optex_sort_match_workspace_bytes(int c, int64_t n_t, int64_t n_s) {
    (void)c; (void)n_t; (void)n_s;
    return 0;  // the in-register path needs no global scratch
}

extern "C" int optex_sort_match(const float *target, const float *source, float *out, int c, int64_t n_t,
                                int64_t n_s, int32_t *perm, void *workspace, size_t workspace_bytes,
                                void *stream) {
    (void)workspace; (void)workspace_bytes;
    OPTEX_TRY(require_sm100());
    if (c < 0 || n_t < 0 || n_s < 0) {
        set_error("optex_sort_match: negative size");
        return OPTEX_EINVAL;
    }
    if (c == 0 || n_t == 0) return OPTEX_OK;
    if (n_s == 0) {
        set_error("optex_sort_match: empty source");
        return OPTEX_EINVAL;
    }
    if (!target || !source || !out) {
        set_error("optex_sort_match: NULL pointer");
        return OPTEX_EINVAL;
    }
    int64_t n = n_t > n_s ? n_t : n_s;
    if (n > (int64_t)NT << MAX_LOG_E) {
        set_error("optex_sort_match: %lld elements per channel exceed the on-chip sort capacity (%d)",
                  (long long)n, NT << MAX_LOG_E);
        return OPTEX_ESIZE;
    }
    int log_e = 0;
    while (((int64_t)NT << log_e) < n) ++log_e;
    cudaStream_t st = (cudaStream_t)stream;
    switch (log_e) {
        case 0: return launch_sort<0>(target, source, out, c, n_t, n_s, perm, st);
        case 1: return launch_sort<1>(target, source, out, c, n_t, n_s, perm, st);
        case 2: return launch_sort<2>(target, source, out, c, n_t, n_s, perm, st);
        case 3: return launch_sort<3>(target, source, out, c, n_t, n_s, perm, st);
        case 4: return launch_sort<4>(target, source, out, c, n_t, n_s, perm, st);
        default: return launch_sort<5>(target, source, out, c, n_t, n_s, perm, st);
    }
}
