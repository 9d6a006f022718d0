// Exact 1-D optimal transport per rotated channel (the north-star's "sort" mode).
//
// NOT in the reference (its per-channel matcher is the 256-bin cdf_match,
// histmatch.py:49-69); defined by oracle/sort_oracle.py:
//     idx = stable argsort(target_c);  ss = sort(source_c)
//     out[idx[r]] = ss[((2r+1)*n_s) / (2*n_t)]
//
// One CTA per channel and per array, the whole channel on chip (512 threads x up to 32 items in registers):
// a stable LSD radix sort, 6 bits per pass -
//   count   : every thread counts the digits of its blocked chunk into private packed 16-bit counters
//             (bank-conflict free, no atomics) and remembers each item's rank among its own equal digits
//   scan    : one block-wide exclusive scan over the packed counters in (digit, thread) order
//   scatter : item -> scanned base + local rank, through a swizzled shared-memory buffer, then back to registers
// The target carries its 16-bit pixel index as a payload and is sorted by the key only; LSD stability on the
// blocked (index-ordered) arrangement makes ties come out in index order, so the permutation is bit-exact with
// torch.sort(stable=True) (-0.0 == +0.0, NaN last).  The source is sorted keys-only, in place (it is scratch).
#include "common.cuh"

namespace optex {
namespace {

// 512 threads x 32 items, 6-bit digits.  (A 1024 x 16 / 5-bit / 7-pass configuration was measured: 489 vs 427 us.)
constexpr int NT = 512;
constexpr int MAX_LOG_E = 5;   // 512 threads x 32 items = 16384 elements per channel
constexpr int RB = 6;          // radix bits per pass: 64 digits, 6 passes over the 32-bit key
constexpr int DIG = 1 << RB;
constexpr int ROWS = DIG / 2;  // packed counter rows: row r holds digit r (low 16 bits) and digit r + ROWS (high 16)
constexpr int PASSES = (32 + RB - 1) / RB;
constexpr int SEG = ROWS;      // linear counter entries (ROWS * NT) raked per thread

__device__ __forceinline__ uint32_t sort_key(float x) {
    if (x != x) return 0xffc00000u;  // canonical NaN sorts after +inf
    return f2ord(x + 0.0f);          // -0.0 + 0.0 == +0.0 : ties with +0.0 like torch
}

// counter array index with one pad word per 32: both the [row][thread] accesses of the counting phase and the
// 32-consecutive-word segments of the raking scan are bank-conflict free
__device__ __forceinline__ int cphys(int L) { return L + (L >> 5); }
// item buffer index: position p = t*E + r is stored at t*E + (r ^ (t mod E)), so that the "thread t reads its
// blocked chunk" accesses (fixed r across a warp) spread over all banks
template <int LOG_E>
__device__ __forceinline__ int bphys(int p) {
    constexpr int E = 1 << LOG_E;
    return (p & ~(E - 1)) | ((p ^ (p >> LOG_E)) & (E - 1));
}

// LSD radix sort of the CTA's NT << LOG_E items (blocked arrangement: thread t holds positions t*E .. t*E+E-1),
// stable, ascending by the 32-bit key k[]; PAYLOAD carries a 16-bit value (the pixel index) along, two per register.
// On return the sorted sequence is in kbuf / pbuf (swizzled by bphys) AND, if `reload_last`, back in k[] / pl[].
//   cnt : (ROWS * NT) * 33 / 32 words of packed 16-bit counters,  wsum : 32 words
template <int LOG_E, bool PAYLOAD>
__device__ __forceinline__ void radix_sort_blocked(uint32_t (&k)[1 << LOG_E], uint32_t (&pl)[((1 << LOG_E) + 1) / 2],
                                                   uint32_t *kbuf, uint16_t *pbuf, uint32_t *cnt, uint32_t *wsum,
                                                   bool reload_last) {
    constexpr int E = 1 << LOG_E;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll 1
    for (int pass = 0; pass < PASSES; ++pass) {
        const int shift = pass * RB;
        // ---- count this thread's digits; remember every item's rank among the thread's equal digits
#pragma unroll
        for (int r = 0; r < ROWS; ++r) cnt[cphys(r * NT + tid)] = 0u;
        uint32_t local[(E + 3) / 4];
#pragma unroll
        for (int i = 0; i < (E + 3) / 4; ++i) local[i] = 0u;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const uint32_t d = (k[i] >> shift) & (DIG - 1);
            const int a = cphys((int)(d & (ROWS - 1)) * NT + tid);
            const uint32_t sh = (d / ROWS) * 16u;
            const uint32_t c = cnt[a];
            local[i >> 2] |= ((c >> sh) & 0xffu) << ((i & 3) * 8);  // < E <= 32
            cnt[a] = c + (1u << sh);
        }
        __syncthreads();
        // ---- exclusive scan of the packed counters in (row, thread) order = (digit, thread) order per 16-bit lane
        {
            const int base = tid * SEG;  // this thread rakes linear entries [base, base + SEG)
            uint32_t sum = 0;
#pragma unroll
            for (int j = 0; j < SEG; ++j) sum += cnt[cphys(base + j)];
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            uint32_t woff = 0, total = 0;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) {
                const uint32_t x = wsum[w];
                if (w < warp) woff += x;
                total += x;
            }
            // the high-lane digits come after all of the low-lane digits (the low-lane total)
            uint32_t run = woff + incl - sum + (total << 16);
#pragma unroll
            for (int j = 0; j < SEG; ++j) {
                const int a = cphys(base + j);
                const uint32_t c = cnt[a];
                cnt[a] = run;
                run += c;
            }
        }
        __syncthreads();
        // ---- scatter to the item's global rank for this digit
#pragma unroll
        for (int i = 0; i < E; ++i) {
            const uint32_t d = (k[i] >> shift) & (DIG - 1);
            const uint32_t basep = (cnt[cphys((int)(d & (ROWS - 1)) * NT + tid)] >> ((d / ROWS) * 16u)) & 0xffffu;
            const int pos = bphys<LOG_E>((int)(basep + ((local[i >> 2] >> ((i & 3) * 8)) & 0xffu)));
            kbuf[pos] = k[i];
            if (PAYLOAD) pbuf[pos] = (uint16_t)(pl[i >> 1] >> ((i & 1) * 16));
        }
        __syncthreads();
        if (pass + 1 < PASSES || reload_last) {
#pragma unroll
            for (int i = 0; i < (E + 1) / 2; ++i) pl[i] = 0u;
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const int pos = bphys<LOG_E>(tid * E + i);
                k[i] = kbuf[pos];
                if (PAYLOAD) pl[i >> 1] |= (uint32_t)pbuf[pos] << ((i & 1) * 16);
            }
            __syncthreads();
        }
    }
}

// smem carve-up: [kbuf: N2 u32][cnt: ROWS*NT*33/32 words][wsum: 32 words][pbuf: N2 u16 (payload only)]
template <int LOG_E, bool PAYLOAD>
__host__ __device__ constexpr size_t radix_smem_bytes() {
    return (size_t)(NT << LOG_E) * (PAYLOAD ? 6 : 4) + (size_t)(ROWS * NT / 32 * 33 + 32) * 4;
}

// ascending sort of every source channel, in place (values rewritten as floats in sorted order)
template <int LOG_E>
__global__ void __launch_bounds__(NT, 1)
sort_source_kernel(float *src, int64_t n_s) {
    constexpr int E = 1 << LOG_E;
    constexpr int N2 = NT * E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *kbuf = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *cnt = kbuf + N2;
    uint32_t *wsum = cnt + ROWS * NT / 32 * 33;
    const int tid = threadIdx.x;
    float *row = src + (int64_t)blockIdx.x * n_s;
    for (int i = tid; i < N2; i += NT) kbuf[bphys<LOG_E>(i)] = i < n_s ? sort_key(row[i]) : 0xffffffffu;
    __syncthreads();
    uint32_t k[E], pl[(E + 1) / 2];
#pragma unroll
    for (int i = 0; i < E; ++i) k[i] = kbuf[bphys<LOG_E>(tid * E + i)];
    __syncthreads();
    radix_sort_blocked<LOG_E, false>(k, pl, kbuf, nullptr, cnt, wsum, false);
    for (int i = tid; i < n_s; i += NT) row[i] = ord2f(kbuf[bphys<LOG_E>(i)]);
}

// stable argsort of every target channel, then rank r receives sorted_source[((2r+1) n_s) / (2 n_t)]
template <int LOG_E>
__global__ void __launch_bounds__(NT, 1)
sort_target_kernel(const float *target, const float *__restrict__ sorted_source, float *out, int64_t n_t, int64_t n_s,
                   int32_t *__restrict__ perm) {
    constexpr int E = 1 << LOG_E;
    constexpr int N2 = NT * E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *kbuf = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *cnt = kbuf + N2;
    uint32_t *wsum = cnt + ROWS * NT / 32 * 33;
    uint16_t *pbuf = reinterpret_cast<uint16_t *>(wsum + 32);
    const int tid = threadIdx.x;
    const int ch = blockIdx.x;
    const float *trow = target + (int64_t)ch * n_t;
    for (int i = tid; i < N2; i += NT) kbuf[bphys<LOG_E>(i)] = i < n_t ? sort_key(trow[i]) : 0xffffffffu;
    __syncthreads();
    uint32_t k[E], pl[(E + 1) / 2];
#pragma unroll
    for (int i = 0; i < (E + 1) / 2; ++i) pl[i] = 0u;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        k[i] = kbuf[bphys<LOG_E>(tid * E + i)];
        pl[i >> 1] |= (uint32_t)((tid * E + i) & 0xffff) << ((i & 1) * 16);  // pixel index < 16384
    }
    __syncthreads();
    radix_sort_blocked<LOG_E, true>(k, pl, kbuf, pbuf, cnt, wsum, true);
    // ---- rank g = tid*E + i ; kbuf is free again (everything was reloaded and synchronised)
    float *stage = reinterpret_cast<float *>(kbuf);
    const float *ss = sorted_source + (int64_t)ch * n_s;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const int g = tid * E + i;
        if (g < n_t) {
            const uint32_t idx = (pl[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
            const int64_t q = ((2 * (int64_t)g + 1) * n_s) / (2 * n_t);
            stage[idx] = __ldg(ss + q);
            if (perm) perm[(int64_t)ch * n_t + g] = (int32_t)idx;
        }
    }
    __syncthreads();
    float *orow = out + (int64_t)ch * n_t;
    for (int i = tid; i < n_t; i += NT) orow[i] = stage[i];
}

template <int LOG_E>
int launch_sort(const float *t, float *s_sorted, float *out, int c, int64_t n_t, int64_t n_s, int32_t *perm,
                cudaStream_t st) {
    constexpr size_t smem_s = radix_smem_bytes<LOG_E, false>();
    constexpr size_t smem_t = radix_smem_bytes<LOG_E, true>();
    static bool attr_done = false;
    if (!attr_done) {
        OPTEX_CUDA(cudaFuncSetAttribute(sort_source_kernel<LOG_E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_s));
        OPTEX_CUDA(cudaFuncSetAttribute(sort_target_kernel<LOG_E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem_t));
        attr_done = true;
    }
    sort_source_kernel<LOG_E><<<c, NT, smem_s, st>>>(s_sorted, n_s);
    OPTEX_LAUNCH_CHECK("sort_source_kernel");
    sort_target_kernel<LOG_E><<<c, NT, smem_t, st>>>(t, s_sorted, out, n_t, n_s, perm);
    OPTEX_LAUNCH_CHECK("sort_target_kernel");
    return OPTEX_OK;
}

}  // namespace
}  // namespace optex

using namespace optex;

extern "C" size_t optex_sort_match_workspace_bytes(int c, int64_t n_t, int64_t n_s) {
    (void)n_t;
    if (c <= 0 || n_s <= 0) return 0;
    return align_up(sizeof(float) * (size_t)c * (size_t)n_s, 256);  // sorted copy of the source
}

namespace optex {
// `source_scratch` is sorted IN PLACE (the OT step hands over its own rotated-style buffer)
int sort_match_inplace(const float *target, float *source_scratch, float *out, int c, int64_t n_t, int64_t n_s,
                       int32_t *perm, cudaStream_t st) {
    int64_t n = n_t > n_s ? n_t : n_s;
    if (n > (int64_t)NT << MAX_LOG_E) {
        set_error("optex_sort_match: %lld elements per channel exceed the on-chip sort capacity (%d)",
                  (long long)n, NT << MAX_LOG_E);
        return OPTEX_ESIZE;
    }
    int log_e = 0;
    while (((int64_t)NT << log_e) < n) ++log_e;
    switch (log_e) {
        case 0: return launch_sort<0>(target, source_scratch, out, c, n_t, n_s, perm, st);
        case 1: return launch_sort<1>(target, source_scratch, out, c, n_t, n_s, perm, st);
        case 2: return launch_sort<2>(target, source_scratch, out, c, n_t, n_s, perm, st);
        case 3: return launch_sort<3>(target, source_scratch, out, c, n_t, n_s, perm, st);
        case 4: return launch_sort<4>(target, source_scratch, out, c, n_t, n_s, perm, st);
        default: return launch_sort<5>(target, source_scratch, out, c, n_t, n_s, perm, st);
    }
}
}  // namespace optex

extern "C" int optex_sort_match(const float *target, const float *source, float *out, int c, int64_t n_t,
                                int64_t n_s, int32_t *perm, void *workspace, size_t workspace_bytes,
                                void *stream) {
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
    const size_t need = optex_sort_match_workspace_bytes(c, n_t, n_s);
    if (!workspace || workspace_bytes < need) {
        set_error("optex_sort_match: workspace %zu < %zu bytes", workspace_bytes, need);
        return OPTEX_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    OPTEX_CUDA(cudaMemcpyAsync(workspace, source, sizeof(float) * (size_t)c * (size_t)n_s, cudaMemcpyDeviceToDevice, st));
    return sort_match_inplace(target, (float *)workspace, out, c, n_t, n_s, perm, st);
}
