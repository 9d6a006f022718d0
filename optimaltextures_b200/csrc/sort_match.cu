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
    pdl_wait();
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
    pdl_wait();
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
int launch_source_sort(float *s_sorted, int c, int64_t n_s, cudaStream_t st) {
    constexpr size_t smem_s = radix_smem_bytes<LOG_E, false>();
    static PerDeviceOnce attr_once1;
        OPTEX_TRY(ensure_dyn_smem(attr_once1, sort_source_kernel<LOG_E>, (int)smem_s));
    launch_pdl(sort_source_kernel<LOG_E>, dim3((unsigned)c), dim3(NT), smem_s, st, s_sorted, n_s);
    OPTEX_LAUNCH_CHECK("sort_source_kernel");
    return OPTEX_OK;
}
template <int LOG_E>
int launch_target_sort(const float *t, const float *s_sorted, float *out, int c, int64_t n_t, int64_t n_s,
                       int32_t *perm, cudaStream_t st) {
    constexpr size_t smem_t = radix_smem_bytes<LOG_E, true>();
    static PerDeviceOnce attr_once2;
        OPTEX_TRY(ensure_dyn_smem(attr_once2, sort_target_kernel<LOG_E>, (int)smem_t));
    launch_pdl(sort_target_kernel<LOG_E>, dim3((unsigned)c), dim3(NT), smem_t, st, t, (const float *)s_sorted, out, n_t, n_s,
               perm);
    OPTEX_LAUNCH_CHECK("sort_target_kernel");
    return OPTEX_OK;
}

// ------------------------------------------------------------------ channels longer than the on-chip capacity
// 16384-element chunks are sorted on chip (same radix kernel), then merged pairwise - run length doubling per
// pass - by a stable merge-path kernel (ties take the earlier run first, so the overall order stays the stable
// argsort).  Keys travel as ordered u32, the pixel index as a second u32 array.
constexpr int CHUNK = NT << MAX_LOG_E;
constexpr int MT = 256, ME = 16, TM = MT * ME;  // merge tile: 4096 outputs per CTA

template <bool PAYLOAD>
__global__ void __launch_bounds__(NT, 1)
sort_chunks_kernel(const float *__restrict__ src, int64_t n, uint32_t *__restrict__ K, uint32_t *__restrict__ I) {
    constexpr int LOG_E = MAX_LOG_E, E = 1 << LOG_E, N2 = NT * E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *kbuf = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *cnt = kbuf + N2;
    uint32_t *wsum = cnt + ROWS * NT / 32 * 33;
    uint16_t *pbuf = reinterpret_cast<uint16_t *>(wsum + 32);
    const int tid = threadIdx.x;
    const int64_t base = (int64_t)blockIdx.x * CHUNK;
    const int64_t rowoff = (int64_t)blockIdx.y * n;
    const int64_t valid = n - base < CHUNK ? n - base : CHUNK;
    for (int i = tid; i < N2; i += NT) kbuf[bphys<LOG_E>(i)] = i < valid ? sort_key(src[rowoff + base + i]) : 0xffffffffu;
    __syncthreads();
    uint32_t k[E], pl[(E + 1) / 2];
#pragma unroll
    for (int i = 0; i < (E + 1) / 2; ++i) pl[i] = 0u;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        k[i] = kbuf[bphys<LOG_E>(tid * E + i)];
        pl[i >> 1] |= (uint32_t)((tid * E + i) & 0xffff) << ((i & 1) * 16);
    }
    __syncthreads();
    radix_sort_blocked<LOG_E, PAYLOAD>(k, pl, kbuf, pbuf, cnt, wsum, false);
    for (int i = tid; i < valid; i += NT) {  // pads sorted last: the first `valid` entries are the real ones
        K[rowoff + base + i] = kbuf[bphys<LOG_E>(i)];
        if (PAYLOAD) I[rowoff + base + i] = (uint32_t)(base + pbuf[bphys<LOG_E>(i)]);
    }
}

// number of elements taken from A among the first d outputs of merge(A, B), ties to A
__device__ __forceinline__ int merge_path(const uint32_t *A, int na, const uint32_t *B, int nb, int d) {
    int lo = d - nb > 0 ? d - nb : 0, hi = d < na ? d : na;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (A[mid] <= B[d - 1 - mid]) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

template <bool PAYLOAD>
__global__ void __launch_bounds__(MT)
merge_pass_kernel(const uint32_t *__restrict__ Kin, const uint32_t *__restrict__ Iin, uint32_t *__restrict__ Kout,
                  uint32_t *__restrict__ Iout, int64_t n, int64_t L) {
    extern __shared__ uint32_t msm[];
    uint32_t *sK = msm, *sI = msm + TM, *oK = msm + 2 * TM, *oI = msm + 3 * TM;
    __shared__ int part[2];
    const int tid = threadIdx.x;
    const int64_t rowoff = (int64_t)blockIdx.y * n;
    const int64_t o0 = (int64_t)blockIdx.x * TM;
    if (o0 >= n) return;
    const int64_t pbase = (o0 / (2 * L)) * (2 * L);
    const int na = (int)(n - pbase < L ? n - pbase : L);
    const int nb = (int)(n - pbase - L < 0 ? 0 : (n - pbase - L < L ? n - pbase - L : L));
    const uint32_t *A = Kin + rowoff + pbase, *B = A + L;
    const int d0 = (int)(o0 - pbase);
    const int d1 = d0 + TM < na + nb ? d0 + TM : na + nb;
    if (tid < 2) part[tid] = merge_path(A, na, B, nb, tid ? d1 : d0);
    __syncthreads();
    const int a0 = part[0], a1 = part[1], b0 = d0 - a0, b1 = d1 - a1;
    const int ca = a1 - a0, cb = b1 - b0, cnt = ca + cb;
    for (int i = tid; i < cnt; i += MT) {
        const int64_t g = i < ca ? pbase + a0 + i : pbase + L + b0 + (i - ca);
        sK[i] = Kin[rowoff + g];
        if (PAYLOAD) sI[i] = Iin[rowoff + g];
    }
    __syncthreads();
    {
        const uint32_t *sA = sK, *sB = sK + ca;
        const int d = tid * ME;
        if (d < cnt) {
            int i = merge_path(sA, ca, sB, cb, d), j = d - i;
#pragma unroll
            for (int e = 0; e < ME; ++e) {
                if (d + e < cnt) {
                    const bool takeA = j >= cb || (i < ca && sA[i] <= sB[j]);
                    const int src = takeA ? i : ca + j;
                    oK[d + e] = sK[src];
                    if (PAYLOAD) oI[d + e] = sI[src];
                    if (takeA) ++i; else ++j;
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < cnt; i += MT) {
        Kout[rowoff + o0 + i] = oK[i];
        if (PAYLOAD) Iout[rowoff + o0 + i] = oI[i];
    }
}

__global__ void keys_to_float_kernel(const uint32_t *__restrict__ K, float *__restrict__ out, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = ord2f(K[i]);
}
// rank r of the target receives sorted_source[((2r+1) n_s) / (2 n_t)]
__global__ void sort_finalize_kernel(const uint32_t *__restrict__ I, const float *__restrict__ ss, float *__restrict__ out,
                                     int32_t *__restrict__ perm, int64_t n_t, int64_t n_s) {
    const int64_t rt = (int64_t)blockIdx.y * n_t, rs = (int64_t)blockIdx.y * n_s;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_t; r += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t idx = I[rt + r];
        const int64_t q = ((2 * r + 1) * n_s) / (2 * n_t);
        out[rt + idx] = __ldg(ss + rs + q);
        if (perm) perm[rt + r] = (int32_t)idx;
    }
}

template <bool PAYLOAD>
int large_sort(const float *src, int c, int64_t n, uint32_t *Ka, uint32_t *Ia, uint32_t *Kb, uint32_t *Ib,
               uint32_t **Kres, uint32_t **Ires, cudaStream_t st) {
    constexpr size_t smem_c = radix_smem_bytes<MAX_LOG_E, true>();
    constexpr size_t smem_m = sizeof(uint32_t) * 4 * TM;
    static PerDeviceOnce attr_once3;
        OPTEX_TRY(ensure_dyn_smem(attr_once3, sort_chunks_kernel<PAYLOAD>, (int)smem_c));
        static PerDeviceOnce attr_once4;
        OPTEX_TRY(ensure_dyn_smem(attr_once4, merge_pass_kernel<PAYLOAD>, (int)smem_m));
    const int64_t chunks = (n + CHUNK - 1) / CHUNK;
    if (chunks > 0x7fffffffLL || c > 65535) {
        set_error("optex_sort_match: problem too large for the grid");
        return OPTEX_ESIZE;
    }
    sort_chunks_kernel<PAYLOAD><<<dim3((unsigned)chunks, (unsigned)c), NT, smem_c, st>>>(src, n, Ka, Ia);
    OPTEX_LAUNCH_CHECK("sort_chunks_kernel");
    uint32_t *kin = Ka, *iin = Ia, *kout = Kb, *iout = Ib;
    const unsigned tiles = (unsigned)((n + TM - 1) / TM);
    for (int64_t L = CHUNK; L < n; L *= 2) {
        merge_pass_kernel<PAYLOAD><<<dim3(tiles, (unsigned)c), MT, smem_m, st>>>(kin, iin, kout, iout, n, L);
        OPTEX_LAUNCH_CHECK("merge_pass_kernel");
        uint32_t *t;
        t = kin; kin = kout; kout = t;
        t = iin; iin = iout; iout = t;
    }
    *Kres = kin;
    if (Ires) *Ires = iin;
    return OPTEX_OK;
}

}  // namespace
}  // namespace optex

using namespace optex;

// scratch of the large-channel path: source phase (two key arrays) and target phase (two key + two index arrays)
// run one after the other and share the workspace
static size_t sort_large_bytes(int c, int64_t n_t, int64_t n_s) {
    size_t src = n_s > CHUNK ? 2 * align_up(sizeof(uint32_t) * (size_t)c * n_s, 256) : 0;
    size_t tgt = n_t > CHUNK ? 4 * align_up(sizeof(uint32_t) * (size_t)c * n_t, 256) : 0;
    return src > tgt ? src : tgt;
}

extern "C" size_t optex_sort_match_workspace_bytes(int c, int64_t n_t, int64_t n_s) {
    if (c <= 0 || n_s <= 0 || n_t < 0) return 0;
    return align_up(sizeof(float) * (size_t)c * (size_t)n_s, 256) + sort_large_bytes(c, n_t, n_s);
}

namespace optex {
size_t sort_match_scratch_bytes(int c, int64_t n_t, int64_t n_s) { return sort_large_bytes(c, n_t, n_s); }

// `source_scratch` is sorted IN PLACE (the OT step hands over its own rotated-style buffer); `ws` is only needed
// when a channel exceeds the on-chip capacity (sort_match_scratch_bytes)
int sort_match_inplace(const float *target, float *source_scratch, float *out, int c, int64_t n_t, int64_t n_s,
                       int32_t *perm, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (n_t >= (1LL << 31) || n_s >= (1LL << 31)) {
        set_error("optex_sort_match: more than 2^31 elements per channel");
        return OPTEX_ESIZE;
    }
    if (ws_bytes < sort_large_bytes(c, n_t, n_s) || (sort_large_bytes(c, n_t, n_s) && !ws)) {
        set_error("optex_sort_match: scratch %zu < %zu bytes", ws_bytes, sort_large_bytes(c, n_t, n_s));
        return OPTEX_EWORKSPACE;
    }
    // ---- source: ascending values, in place
    if (n_s <= CHUNK) {
        int log_e = 0;
        while (((int64_t)NT << log_e) < n_s) ++log_e;
        switch (log_e) {
            case 0: OPTEX_TRY(launch_source_sort<0>(source_scratch, c, n_s, st)); break;
            case 1: OPTEX_TRY(launch_source_sort<1>(source_scratch, c, n_s, st)); break;
            case 2: OPTEX_TRY(launch_source_sort<2>(source_scratch, c, n_s, st)); break;
            case 3: OPTEX_TRY(launch_source_sort<3>(source_scratch, c, n_s, st)); break;
            case 4: OPTEX_TRY(launch_source_sort<4>(source_scratch, c, n_s, st)); break;
            default: OPTEX_TRY(launch_source_sort<5>(source_scratch, c, n_s, st)); break;
        }
    } else {
        const size_t one = align_up(sizeof(uint32_t) * (size_t)c * n_s, 256);
        uint32_t *Ka = (uint32_t *)ws, *Kb = (uint32_t *)((char *)ws + one), *Kres = nullptr;
        OPTEX_TRY(large_sort<false>(source_scratch, c, n_s, Ka, nullptr, Kb, nullptr, &Kres, nullptr, st));
        const int64_t total = (int64_t)c * n_s;
        keys_to_float_kernel<<<(unsigned)(sm_count() * 8), 256, 0, st>>>(Kres, source_scratch, total);
        OPTEX_LAUNCH_CHECK("keys_to_float_kernel");
    }
    // ---- target: stable argsort, then the quantile assignment
    if (n_t <= CHUNK) {
        int log_e = 0;
        while (((int64_t)NT << log_e) < n_t) ++log_e;
        switch (log_e) {
            case 0: return launch_target_sort<0>(target, source_scratch, out, c, n_t, n_s, perm, st);
            case 1: return launch_target_sort<1>(target, source_scratch, out, c, n_t, n_s, perm, st);
            case 2: return launch_target_sort<2>(target, source_scratch, out, c, n_t, n_s, perm, st);
            case 3: return launch_target_sort<3>(target, source_scratch, out, c, n_t, n_s, perm, st);
            case 4: return launch_target_sort<4>(target, source_scratch, out, c, n_t, n_s, perm, st);
            default: return launch_target_sort<5>(target, source_scratch, out, c, n_t, n_s, perm, st);
        }
    }
    const size_t one = align_up(sizeof(uint32_t) * (size_t)c * n_t, 256);
    char *w = (char *)ws;
    uint32_t *Ka = (uint32_t *)w, *Kb = (uint32_t *)(w + one), *Ia = (uint32_t *)(w + 2 * one),
             *Ib = (uint32_t *)(w + 3 * one), *Kres = nullptr, *Ires = nullptr;
    OPTEX_TRY(large_sort<true>(target, c, n_t, Ka, Ia, Kb, Ib, &Kres, &Ires, st));
    const unsigned gx = (unsigned)((n_t + 255) / 256 < 1024 ? (n_t + 255) / 256 : 1024);
    sort_finalize_kernel<<<dim3(gx, (unsigned)c), 256, 0, st>>>(Ires, source_scratch, out, perm, n_t, n_s);
    OPTEX_LAUNCH_CHECK("sort_finalize_kernel");
    return OPTEX_OK;
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
    const size_t copy_bytes = align_up(sizeof(float) * (size_t)c * (size_t)n_s, 256);
    OPTEX_CUDA(cudaMemcpyAsync(workspace, source, sizeof(float) * (size_t)c * (size_t)n_s, cudaMemcpyDeviceToDevice, st));
    return sort_match_inplace(target, (float *)workspace, out, c, n_t, n_s, perm, (char *)workspace + copy_bytes,
                              workspace_bytes - copy_bytes, st);
}
