// Per-channel 256-bin CDF matching on channel-major data, bit-exact with the
// reference's torch-CPU arithmetic.
//
//   reference: cdf_match()  histmatch.py:49-69,  interp()  histmatch.py:72-92
//   arithmetic restated in oracle/cdf_explicit.py (histc / linspace / cumsum /
//   searchsorted rules of torch 2.11 CPU); every fp32 operation below uses the
//   round-to-nearest intrinsics so nvcc can neither fuse nor reorder them.
//
// The reference runs a Python loop over channels with ~25 launches and >= 2 host
// syncs per channel.  Here: 4 launches for ALL channels (init, range, histogram,
// apply), no host sync, grid = (splits, C) sized to fill 148 SMs at any (C, N).
//
// Histogramming avoids shared-memory atomics (2 cycles/lane on Blackwell): every
// thread owns a private byte-counter histogram laid out bank-conflict-free
// (word w of thread t at [w*NT + t]); counters are flushed with packed 16-bit
// adds before they can overflow.
#include "common.cuh"

namespace optex {
namespace {

constexpr int NT = 256;
constexpr int MAX_BINS = 1024;
constexpr int PRIV_MAX_BINS = 512;  // private byte histograms: bins * NT bytes of smem

// ---- element iteration ------------------------------------------------------
// Calls f(x) for every element of row[beg, end) with 128-bit loads when possible (block of NTH threads).
template <int NTH, typename F>
__device__ __forceinline__ void for_each(const float *__restrict__ row, int64_t beg, int64_t end,
                                         bool vec, F f) {
    if (vec) {  // row 16B-aligned, beg % 4 == 0
        int64_t nv = (end - beg) >> 2;
        const float4 *p = reinterpret_cast<const float4 *>(row + beg);
        for (int64_t i = threadIdx.x; i < nv; i += NTH) {
            float4 v = __ldg(p + i);
            f(v.x); f(v.y); f(v.z); f(v.w);
        }
        for (int64_t i = beg + (nv << 2) + threadIdx.x; i < end; i += NTH) f(__ldg(row + i));
    } else {
        for (int64_t i = beg + threadIdx.x; i < end; i += NTH) f(__ldg(row + i));
    }
}

__device__ __forceinline__ void slice_of(int64_t n, int part, int nparts, int64_t &beg, int64_t &end) {
    int64_t len = ((n + nparts - 1) / nparts + 3) & ~int64_t(3);
    beg = len * part;
    end = beg + len;
    if (beg > n) beg = n;
    if (end > n) end = n;
}

// ---- the torch arithmetic ---------------------------------------------------
// torch.histc bin rule (CPU): int64((x - lo) * bins / (hi - lo)), bin == bins -> bins-1;
// a degenerate range [v, v] becomes [v-1, v+1].
struct HistRange {
    float lo, width, fbins, rcp, scale;
    int bins;
    __device__ HistRange(float lo_, float hi_, int bins_) : bins(bins_) {
        if (lo_ == hi_) {
            lo_ = __fsub_rn(lo_, 1.f);
            hi_ = __fadd_rn(hi_, 1.f);
        }
        lo = lo_;
        width = __fsub_rn(hi_, lo_);
        fbins = (float)bins_;
        rcp = 1.f / width;
        scale = fbins * rcp;
    }
    // Exact trunc(fdiv_rn(a, width)) without dividing every element: a * rcp is within a few ulp (< 1e-4 absolute
    // for quotients <= 1024) of the true quotient, so it has the same integer part unless it lies within 1e-3 of an
    // integer - only those elements (about 0.2 %) take the IEEE division.
    __device__ __forceinline__ int bin(float x) const {
        const float a = __fmul_rn(__fsub_rn(x, lo), fbins);
        float q = a * rcp;
        const float r = __fsub_rn(__fadd_rn(q, 12582912.f), 12582912.f);  // nearest integer (|q| < 2^22)
        if (!(fabsf(q - r) >= 1e-3f)) q = __fdiv_rn(a, width);           // also taken for NaN / huge q
        int b = (int)q;  // truncation, like the int64 cast
        b = b >= bins ? bins - 1 : b;
        return b < 0 ? 0 : b;  // memory safety for NaN only
    }
    // same rule, but x == hi is reported as bin `bins` (the caller folds it into bins - 1); values are in [lo, hi]
    // by construction (lo / hi are the data's own extrema), NaN converts to 0
    __device__ __forceinline__ int bin_unclamped(float x) const {
        const float d = __fsub_rn(x, lo);
        float q = d * scale;   // one multiply on the common path (scale = bins / width, a few ulp like a * rcp)
        const float r = __fsub_rn(__fadd_rn(q, 12582912.f), 12582912.f);
        if (!(fabsf(q - r) >= 1e-3f)) q = __fdiv_rn(__fmul_rn(d, fbins), width);
        const unsigned b = (unsigned)(int)q;
        return (int)(b > (unsigned)bins ? (unsigned)bins : b);  // one unsigned min: memory safety only
    }
};

// torch.linspace(lo, hi, bins+1)[i], i in 1..bins  (CPU kernel: symmetric, FMA-rounded)
__device__ __forceinline__ float linspace_at(float lo, float hi, float step, int i, int bins) {
    int half = (bins + 1) / 2;
    return i < half ? fmaf(step, (float)i, lo) : fmaf(-step, (float)(bins - i), hi);
}

// torch.searchsorted(xp, x) (right=False): identical bisection to ATen's cus_lower_bound.
__device__ __forceinline__ int lower_bound(const float *xp, int len, float x) {
    int start = 0, end = len;
    while (start < end) {
        int mid = start + ((end - start) >> 1);
        if (!(xp[mid] >= x)) start = mid + 1;
        else end = mid;
    }
    return start;
}

__device__ __forceinline__ bool finite_f(float f) { return fabsf(f) <= 3.402823466e+38f; }

// histmatch.py:72-92 for one x, given i = searchsorted(xp, x).
__device__ __forceinline__ float interp_at(float x, const float *xp, const float *fp, int len, int i) {
    int last = len - 1;
    i = i > last ? last : i;  // never taken on valid data (x <= xp[last]); memory safety
    int j = i + 1 > last ? last : i + 1;
    float xi = xp[i], xj = xp[j], fi = fp[i], fj = fp[j];
    float slope = __fdiv_rn(__fsub_rn(fj, fi), __fsub_rn(xj, xi));
    float f = __fadd_rn(__fmul_rn(slope, __fsub_rn(x, xi)), fi);
    if (!finite_f(f)) {
        f = __fadd_rn(__fmul_rn(slope, __fsub_rn(x, xj)), fj);
        if (!finite_f(f)) f = fi;
    }
    return f;
}

// ---- kernels ----------------------------------------------------------------
// Per-channel range slots minmax[c][2]: [0] = f2ord(min), [1] = ~f2ord(max).  Both are folded with atomicMin, so
// the whole array initialises with one cudaMemsetAsync(0xFF) and the forward-rotation GEMM can produce the range
// in its epilogue (gemm_tcgen05.cu), saving a full read pass over the rotated features.
__device__ __forceinline__ float range_lo(const uint32_t *minmax, int ch) { return ord2f(minmax[2 * ch]); }
__device__ __forceinline__ float range_hi(const uint32_t *minmax, int ch) { return ord2f(~minmax[2 * ch + 1]); }

// lo = min(t.min(), s.min()); hi = max(t.max(), s.max())     histmatch.py:52-53
__global__ void __launch_bounds__(NT)
cdf_range_kernel(const float *__restrict__ t, const float *__restrict__ s, int64_t n_t, int64_t n_s,
                 uint32_t *__restrict__ minmax, int t_vec, int s_vec) {
    pdl_wait();
    const int ch = blockIdx.y;
    float mn = INFINITY, mx = -INFINITY;
    auto upd = [&](float x) { mn = fminf(mn, x); mx = fmaxf(mx, x); };
    int64_t b, e;
    slice_of(n_t, blockIdx.x, gridDim.x, b, e);
    for_each<NT>(t + (int64_t)ch * n_t, b, e, t_vec, upd);
    slice_of(n_s, blockIdx.x, gridDim.x, b, e);
    for_each<NT>(s + (int64_t)ch * n_s, b, e, s_vec, upd);
    mn = warp_min(mn);
    mx = warp_max(mx);
    __shared__ float smn[NT / 32], smx[NT / 32];
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x < 32) {
        mn = threadIdx.x < NT / 32 ? smn[threadIdx.x] : INFINITY;
        mx = threadIdx.x < NT / 32 ? smx[threadIdx.x] : -INFINITY;
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (threadIdx.x == 0 && mn <= mx) {
            atomicMin(&minmax[2 * ch], f2ord(mn));
            atomicMin(&minmax[2 * ch + 1], ~f2ord(mx));
        }
    }
}

// Accumulate one row slice into the block's u32 histogram `acc` (smem, `bins` entries).
// PRIV: every thread counts into its own byte-counter histogram (word w of thread t at priv[w*NTH + t]: no bank
// conflicts, no atomics); after at most 252 elements per thread the counters are folded into `acc` with packed
// 16-bit adds.  NTH is small (128) on purpose: the fold costs ~bins/4 words per thread, so a thread must count a
// few hundred elements per fold for it to amortise (N = 16384 per channel at conv4_1 1024^2 = 128 per thread).
// `tid` / `bar`: the NTH threads that share `priv` and `acc` synchronise on hardware barrier `bar` (0 = the whole
// CTA), so two groups of one CTA can count two arrays side by side (cdf_channel_kernel).
__device__ __forceinline__ void group_sync(int bar, int nth) {
    asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nth) : "memory");
}

template <bool PRIV, int NTH>
__device__ __forceinline__ void hist_slice(const float *__restrict__ row, int64_t beg, int64_t end,
                                           bool vec, const HistRange &hr, uint32_t *priv,
                                           uint32_t *acc, int tid, int bar) {
    if (!PRIV) {
        for_each<NTH>(row, beg, end, vec, [&](float x) { atomicAdd(&acc[hr.bin(x)], 1u); });
        return;
    }
    // byte counters, [bin][thread] (bins + 1 rows: the closed right edge x == hi lands in row `bins` and is folded
    // into the last bin, like torch).  One LDS.U8 / IADD / STS.U8 per element, no shifts or masks; the 32 lanes of a
    // warp touch 8 consecutive words of one or several rows.
    uint8_t *pb = reinterpret_cast<uint8_t *>(priv);
    const int rows = hr.bins + 1;
    const int64_t chunk = (int64_t)NTH * 252;  // 63 float4 per thread: byte counters stay <= 252
    for (int64_t cb = beg; cb < end; cb += chunk) {
        int64_t ce = cb + chunk < end ? cb + chunk : end;
        for (int i = tid; i < rows * (NTH / 4); i += NTH) priv[i] = 0u;
        group_sync(bar, NTH);
        auto one = [&](float x) {
            const int a = hr.bin_unclamped(x) * NTH + tid;
            pb[a] = (uint8_t)(pb[a] + 1);
        };
        if (vec) {
            // four elements per step with their four counter loads in flight together (a read-modify-write chain per
            // element would expose the shared-memory latency 4x); equal addresses are resolved in registers and
            // the stores keep program order
            const int64_t nv = (ce - cb) >> 2;
            const float4 *p4 = reinterpret_cast<const float4 *>(row + cb);
            // 4 vector loads in flight per thread (the loop is otherwise bound by one L2 / HBM round trip per step)
            constexpr int PF = 4;
            float4 ring[PF];
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int64_t j = tid + (int64_t)k * NTH;
                ring[k] = j < nv ? __ldg(p4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (int64_t i = tid; i < nv; i += (int64_t)PF * NTH) {
#pragma unroll
              for (int k = 0; k < PF; ++k) {
                const int64_t cur = i + (int64_t)k * NTH;
                if (cur >= nv) break;
                const float4 v = ring[k];
                const int64_t nxt = cur + (int64_t)PF * NTH;
                if (nxt < nv) ring[k] = __ldg(p4 + nxt);
                const int a0 = hr.bin_unclamped(v.x) * NTH + tid, a1 = hr.bin_unclamped(v.y) * NTH + tid;
                const int a2 = hr.bin_unclamped(v.z) * NTH + tid, a3 = hr.bin_unclamped(v.w) * NTH + tid;
                const uint32_t c0 = pb[a0], c1 = pb[a1], c2 = pb[a2], c3 = pb[a3];
                const uint32_t n0 = c0 + 1;
                const uint32_t n1 = (a1 == a0 ? n0 : c1) + 1;
                const uint32_t n2 = (a2 == a1 ? n1 : (a2 == a0 ? n0 : c2)) + 1;
                const uint32_t n3 = (a3 == a2 ? n2 : (a3 == a1 ? n1 : (a3 == a0 ? n0 : c3))) + 1;
                pb[a0] = (uint8_t)n0;
                pb[a1] = (uint8_t)n1;
                pb[a2] = (uint8_t)n2;
                pb[a3] = (uint8_t)n3;
              }
            }
            for (int64_t i = cb + (nv << 2) + tid; i < ce; i += NTH) one(__ldg(row + i));
        } else {
            for (int64_t i = cb + tid; i < ce; i += NTH) one(__ldg(row + i));
        }
        group_sync(bar, NTH);
        // fold: thread t sums the NTH byte counters of rows t, t + NTH, ... with packed 16-bit adds
        constexpr int WPR = NTH / 4;  // words per row
        for (int r = tid; r < rows; r += NTH) {
            uint32_t even = 0, odd = 0;
            int w = (tid & 31) % WPR;  // start each lane on a different word: distinct banks
            for (int it = 0; it < WPR; ++it) {
                const uint32_t v = priv[r * WPR + w];
                w = w + 1 == WPR ? 0 : w + 1;
                even += v & 0x00ff00ffu;
                odd += (v >> 8) & 0x00ff00ffu;
            }
            const uint32_t total = (even & 0xffffu) + (even >> 16) + (odd & 0xffffu) + (odd >> 16);
            if (total) atomicAdd(&acc[r < hr.bins ? r : hr.bins - 1], total);
        }
        group_sync(bar, NTH);
    }
}

constexpr int NTH_HIST = 96;  // 24.6 KB of private counters per CTA: all 2C CTAs of C = 512 resident in one wave

// th = histc(t, bins, lo, hi), sh = histc(s, bins, lo, hi)     histmatch.py:57-58
template <bool PRIV>
__global__ void __launch_bounds__(NTH_HIST, 9)
cdf_hist_kernel(const float *__restrict__ t, const float *__restrict__ s, int64_t n_t, int64_t n_s,
                const uint32_t *__restrict__ minmax, uint32_t *__restrict__ hist, int bins, int t_vec,
                int s_vec) {
    pdl_wait();
    extern __shared__ uint32_t smem_u32[];
    uint32_t *acc = smem_u32;          // [bins]
    uint32_t *priv = smem_u32 + bins;  // [bins/4][NTH]  (PRIV only)
    const int ch = blockIdx.y;
    for (int i = threadIdx.x; i < bins; i += NTH_HIST) acc[i] = 0u;
    __syncthreads();
    HistRange hr(range_lo(minmax, ch), range_hi(minmax, ch), bins);
    // blockIdx.z == 0 counts the target slice, 1 the source slice (twice the warps in flight)
    int64_t b, e;
    if (blockIdx.z == 0) {
        slice_of(n_t, blockIdx.x, gridDim.x, b, e);
        hist_slice<PRIV, NTH_HIST>(t + (int64_t)ch * n_t, b, e, t_vec, hr, priv, acc, threadIdx.x, 0);
    } else {
        slice_of(n_s, blockIdx.x, gridDim.x, b, e);
        hist_slice<PRIV, NTH_HIST>(s + (int64_t)ch * n_s, b, e, s_vec, hr, priv, acc, threadIdx.x, 0);
    }
    __syncthreads();
    uint32_t *gh = hist + ((int64_t)ch * 2 + blockIdx.z) * bins;
    if (gridDim.x == 1) {
        for (int i = threadIdx.x; i < bins; i += NTH_HIST) gh[i] = acc[i];
    } else {
        for (int i = threadIdx.x; i < bins; i += NTH_HIST)
            if (acc[i]) atomicAdd(&gh[i], acc[i]);
    }
}

// Build edges / CDFs / remap in shared memory from the channel's two histograms.
//   xs: edges[bins], rm: remap[bins]; scratch tc[bins], sc[bins] (float), cnt[2*bins] (u32)
template <int NT>
__device__ void build_tables(const uint32_t *gh, float lo, float hi, int bins,
                             float *edges, float *remap, float *tc, float *sc, uint32_t *cnt) {
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * bins; i += NT) cnt[i] = gh[i];
    __syncthreads();
    // inclusive prefix sums (exact integers; torch's fp32 cumsum is exact below 2^24)
    for (int off = 1; off < bins; off <<= 1) {
        uint32_t a[(2 * MAX_BINS + NT - 1) / NT];
        int k = 0;
        for (int i = tid; i < 2 * bins; i += NT, ++k) {
            int pos = i >= bins ? i - bins : i;
            a[k] = pos >= off ? cnt[i - off] : 0u;
        }
        __syncthreads();
        k = 0;
        for (int i = tid; i < 2 * bins; i += NT, ++k) cnt[i] += a[k];
        __syncthreads();
    }
    const float tot_t = (float)cnt[bins - 1], tot_s = (float)cnt[2 * bins - 1];
    const float step = __fdiv_rn(__fsub_rn(hi, lo), (float)bins);
    for (int i = tid; i < bins; i += NT) {
        tc[i] = __fdiv_rn((float)cnt[i], tot_t);          // histmatch.py:61-62
        sc[i] = __fdiv_rn((float)cnt[bins + i], tot_s);   // histmatch.py:64-65
        edges[i] = linspace_at(lo, hi, step, i + 1, bins);  // histmatch.py:59
    }
    __syncthreads();
    for (int i = tid; i < bins; i += NT)                   // histmatch.py:67
        remap[i] = interp_at(tc[i], sc, edges, bins, lower_bound(sc, bins, tc[i]));
    __syncthreads();
}

// One CTA per channel: edges / remap / slope tables -> tbl[ch][4][bins]  ([3][0] = edges are non-decreasing)
__global__ void __launch_bounds__(NT)
cdf_tables_kernel(const uint32_t *__restrict__ minmax, const uint32_t *__restrict__ hist, int bins,
                  float *__restrict__ tbl, float *__restrict__ tables_out) {
    pdl_wait();
    extern __shared__ float smem_f32[];
    float *edges = smem_f32, *remap = edges + bins, *tc = remap + bins, *sc = tc + bins;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(sc + bins);
    const int ch = blockIdx.x;
    const float lo = range_lo(minmax, ch), hi = range_hi(minmax, ch);
    build_tables<NT>(hist + (int64_t)ch * 2 * bins, lo, hi, bins, edges, remap, tc, sc, cnt);
    float *o = tbl + (int64_t)ch * 4 * bins;
    const int last = bins - 1;
    int mono = 1;
    for (int i = threadIdx.x; i < bins; i += NT) {
        const int j = i + 1 > last ? last : i + 1;
        o[i] = edges[i];
        o[bins + i] = remap[i];
        o[2 * bins + i] = __fdiv_rn(__fsub_rn(remap[j], remap[i]), __fsub_rn(edges[j], edges[i]));  // histmatch.py:79
        if (!(edges[i] <= edges[j])) mono = 0;
        if (tables_out) {
            tables_out[((int64_t)ch * 2 + 0) * bins + i] = edges[i];
            tables_out[((int64_t)ch * 2 + 1) * bins + i] = remap[i];
        }
    }
    mono = __syncthreads_and(mono);
    if (threadIdx.x == 0) o[3 * bins] = mono ? 1.f : 0.f;
}

// matched = interp(x, edges, remap)     histmatch.py:68
// searchsorted(edges, x) is found from an arithmetic estimate of the bin plus an exact fix-up against the edges
// (identical to the bisection whenever the edges are non-decreasing; otherwise the bisection itself runs).
struct ApplyRule {
    const float *edges, *remap, *slope;
    float lo, inv;
    int bins, last;
    bool mono;
    __device__ __forceinline__ float operator()(float x) const {
        int i;
        float xi;
        if (mono) {
            // the estimate k is the answer iff edges[k-1] < x <= edges[k] (k = last: no upper test, like the clamp of
            // searchsorted's result) - one predictable branch on the common path; everything else (exact ties with an
            // edge, rounding next to one, NaN, repeated edges) walks from k exactly as the bisection would land
            int k = (int)((x - lo) * inv);
            k = k < 0 ? 0 : (k > last ? last : k);
            const float t1 = edges[k];
            const float t0 = edges[k > 0 ? k - 1 : 0];
            if ((k == 0 || t0 < x) && (x <= t1 || k == last)) {
                i = k;
                xi = t1;
            } else {
                i = k;
                while (i > 0 && edges[i - 1] >= x) --i;
                while (i < last && edges[i] < x) ++i;
                if (x != x) i = last;
                xi = edges[i];
            }
        } else {
            i = lower_bound(edges, bins, x);
            i = i > last ? last : i;
            xi = edges[i];
        }
        const float fi = remap[i], sl = slope[i];
        float f = __fadd_rn(__fmul_rn(sl, __fsub_rn(x, xi)), fi);
        if (!finite_f(f)) {
            const int j = i + 1 > last ? last : i + 1;
            f = __fadd_rn(__fmul_rn(sl, __fsub_rn(x, edges[j])), remap[j]);
            if (!finite_f(f)) f = fi;
        }
        return f;
    }
};

// out[b, e) = rule(row[b, e)) for a block of NTH threads; 128-bit accesses with PF loads in flight per thread (one
// L2 round trip per step otherwise).  row may alias out: a thread only ever writes elements it has already read.
template <int NTH>
__device__ __forceinline__ void apply_slice(const ApplyRule &one, const float *row, float *orow, int64_t b, int64_t e,
                                            bool vec) {
    if (vec) {
        const int64_t nv = (e - b) >> 2;
        const float4 *p = reinterpret_cast<const float4 *>(row + b);
        float4 *q = reinterpret_cast<float4 *>(orow + b);
        constexpr int PF = 4;
        float4 ring[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) {
            const int64_t j = threadIdx.x + (int64_t)k * NTH;
            ring[k] = j < nv ? p[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int64_t i = threadIdx.x; i < nv; i += (int64_t)PF * NTH) {
#pragma unroll
            for (int k = 0; k < PF; ++k) {
                const int64_t cur = i + (int64_t)k * NTH;
                if (cur >= nv) break;
                const float4 v = ring[k];
                const int64_t nxt = cur + (int64_t)PF * NTH;
                if (nxt < nv) ring[k] = p[nxt];
                q[cur] = make_float4(one(v.x), one(v.y), one(v.z), one(v.w));
            }
        }
        for (int64_t i = b + (nv << 2) + threadIdx.x; i < e; i += NTH) orow[i] = one(row[i]);
    } else {
        for (int64_t i = b + threadIdx.x; i < e; i += NTH) orow[i] = one(row[i]);
    }
}

__global__ void __launch_bounds__(NT)
cdf_apply_kernel(const float *t, float *out, int64_t n_t, const uint32_t *__restrict__ minmax,
                 const float *__restrict__ tbl, int bins, int vec) {
    pdl_wait();
    __shared__ float tb[3 * MAX_BINS];
    const int ch = blockIdx.y;
    const float *g = tbl + (int64_t)ch * 4 * bins;
    for (int i = threadIdx.x; i < 3 * bins; i += NT) tb[i] = g[i];
    const bool mono = g[3 * bins] != 0.f;
    __syncthreads();
    const float lo = range_lo(minmax, ch), hi = range_hi(minmax, ch);
    const ApplyRule one{tb, tb + bins, tb + 2 * bins, lo, hi > lo ? (float)bins / (hi - lo) : 0.f, bins, bins - 1, mono};
    int64_t b, e;
    slice_of(n_t, blockIdx.x, gridDim.x, b, e);
    apply_slice<NT>(one, t + (int64_t)ch * n_t, out + (int64_t)ch * n_t, b, e, vec != 0);
}

// The whole matcher for ONE channel in one CTA - histograms of both arrays (two groups of NTH_HIST threads side by
// side), tables, apply - for the many-channel / short-row regime (conv4_1, conv5_1): the histograms and tables never
// leave shared memory, the second pass over the target row hits L1 / L2, and the three launches with their
// dependency bubbles become one wave of C CTAs whose phases overlap each other on every SM.
constexpr int NT_CH = 2 * NTH_HIST;
__global__ void __launch_bounds__(NT_CH, 4)
cdf_channel_kernel(const float *t, const float *__restrict__ s, float *out, int64_t n_t, int64_t n_s,
                   const uint32_t *__restrict__ minmax, int bins, int t_vec, int s_vec, int o_vec,
                   float *__restrict__ tables_out, int need_range) {
    pdl_wait();
    extern __shared__ uint32_t smem_u32[];
    uint32_t *acc = smem_u32;                              // [2][bins]: target, source counts
    uint32_t *priv = smem_u32 + 2 * bins;                  // [2][(bins + 1) * NTH_HIST bytes]
    const int priv_words = (bins + 1) * (NTH_HIST / 4);
    const int ch = blockIdx.x;
    for (int i = threadIdx.x; i < 2 * bins; i += NT_CH) acc[i] = 0u;
    float lo, hi;
    if (need_range) {
        // histmatch.py:52-53 in this CTA: the channel's two rows are read three times here anyway (range, histogram,
        // apply) and stay in L1 / L2 - cheaper than the REDUX + atomic fold in the epilogue of the rotation GEMMs
        // (measured: 11 us per forward rotation at the headline shape against ~3 us here)
        __shared__ float s_mn[NT_CH / 32], s_mx[NT_CH / 32];
        float mn = INFINITY, mx = -INFINITY;
        auto upd = [&](float x) { mn = fminf(mn, x); mx = fmaxf(mx, x); };
        for_each<NT_CH>(t + (int64_t)ch * n_t, 0, n_t, t_vec != 0, upd);
        for_each<NT_CH>(s + (int64_t)ch * n_s, 0, n_s, s_vec != 0, upd);
        mn = warp_min(mn);
        mx = warp_max(mx);
        if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = mn; s_mx[threadIdx.x >> 5] = mx; }
        __syncthreads();
        mn = s_mn[0]; mx = s_mx[0];
#pragma unroll
        for (int w = 1; w < NT_CH / 32; ++w) { mn = fminf(mn, s_mn[w]); mx = fmaxf(mx, s_mx[w]); }
        lo = mn; hi = mx;
    } else {
        __syncthreads();
        lo = range_lo(minmax, ch); hi = range_hi(minmax, ch);
    }
    {
        HistRange hr(lo, hi, bins);
        const int g = threadIdx.x >= NTH_HIST ? 1 : 0, tid = threadIdx.x - g * NTH_HIST;
        if (g == 0)
            hist_slice<true, NTH_HIST>(t + (int64_t)ch * n_t, 0, n_t, t_vec, hr, priv, acc, tid, 1);
        else
            hist_slice<true, NTH_HIST>(s + (int64_t)ch * n_s, 0, n_s, s_vec, hr, priv + priv_words, acc + bins, tid, 2);
    }
    __syncthreads();
    // tables in the (now free) private-counter space
    float *edges = reinterpret_cast<float *>(priv), *remap = edges + bins, *slope = remap + bins;
    float *tc = slope + bins, *sc = tc + bins;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(sc + bins);
    build_tables<NT_CH>(acc, lo, hi, bins, edges, remap, tc, sc, cnt);
    const int last = bins - 1;
    int mono = 1;
    for (int i = threadIdx.x; i < bins; i += NT_CH) {
        const int j = i + 1 > last ? last : i + 1;
        slope[i] = __fdiv_rn(__fsub_rn(remap[j], remap[i]), __fsub_rn(edges[j], edges[i]));  // histmatch.py:79
        if (!(edges[i] <= edges[j])) mono = 0;
        if (tables_out) {
            tables_out[((int64_t)ch * 2 + 0) * bins + i] = edges[i];
            tables_out[((int64_t)ch * 2 + 1) * bins + i] = remap[i];
        }
    }
    mono = __syncthreads_and(mono);
    const ApplyRule one{edges, remap, slope, lo, hi > lo ? (float)bins / (hi - lo) : 0.f, bins, last, mono != 0};
    apply_slice<NT_CH>(one, t + (int64_t)ch * n_t, out + (int64_t)ch * n_t, 0, n_t, o_vec != 0);
}

__global__ void interp_kernel(const float *__restrict__ x, const float *__restrict__ xp,
                              const float *__restrict__ fp, float *__restrict__ out, int64_t n, int len) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = interp_at(x[i], xp, fp, len, lower_bound(xp, len, x[i]));
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// the (c, n) regime of cdf_channel_kernel: the whole matcher of a channel, its range included, in one CTA
bool cdf_uses_channel_kernel(int c, int64_t n_t, int64_t n_s, int bins) {
    static const char *no_fused = getenv("OPTEX_NO_CDF_FUSED");
    const int64_t n_big = n_t > n_s ? n_t : n_s;
    const bool priv = (bins % 4 == 0) && bins <= PRIV_MAX_BINS;
    return priv && n_big <= 32768 && c >= 2 * sm_count() && !(no_fused && atoi(no_fused));
}

int cdf_splits(int c, int64_t n) {
    // enough CTAs for ~4 resident blocks on each of the SMs, at least 8K elements per CTA
    int64_t want = (4LL * sm_count() + c - 1) / c;
    int64_t cap = (n + 8191) / 8192;
    int64_t s = want < cap ? want : cap;
    return (int)(s < 1 ? 1 : s);
}

}  // namespace optex

using namespace optex;

extern "C" size_t optex_cdf_match_workspace_bytes(int c, int bins) {
    if (c <= 0 || bins <= 0) return 0;
    return align_up(sizeof(uint32_t) * 2 * (size_t)c, 256) + align_up(sizeof(uint32_t) * 2 * (size_t)c * bins, 256) +
           align_up(sizeof(float) * (4 * (size_t)c * bins + 4), 256);
}

namespace optex {
size_t cdf_minmax_bytes(int c) { return sizeof(uint32_t) * 2 * (size_t)c; }

namespace {
__global__ void fill_u32_kernel(uint32_t *p, int64_t n, uint32_t v) {
    pdl_wait();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
}  // namespace
// (a plain kernel: cudaMemsetAsync on a few KB measured ~10x slower than this launch)
int fill_u32(uint32_t *p, int64_t n, uint32_t v, cudaStream_t st) {
    if (n <= 0) return OPTEX_OK;
    launch_pdl(fill_u32_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, p, n, v);
    OPTEX_LAUNCH_CHECK("fill_u32_kernel");
    return OPTEX_OK;
}

// have_range: minmax (the first 2c words of `workspace`, initialised to 0xFF bytes by the caller) already holds
// the per-channel range - the forward-rotation GEMM folded it in its epilogue.
int cdf_match_core(const float *target, const float *source, float *out, int c, int64_t n_t, int64_t n_s, int bins,
                   float *tables, void *workspace, size_t workspace_bytes, bool have_range, cudaStream_t st) {
    // Counts are exact 32-bit integers here.  Below 2^24 elements per channel that is also what the reference computes
    // (its fp32 histc / cumsum are exact there) and the result is bit-identical; above, the reference itself loses
    // counts to fp32 rounding and this kernel simply stays exact (a 4096^2 colour transfer has 16.7 M pixels per channel).
    if (n_t >= (1LL << 31) || n_s >= (1LL << 31)) {
        set_error("optex_cdf_match: n >= 2^31 per channel");
        return OPTEX_ESIZE;
    }
    if (c > 65535) {
        set_error("optex_cdf_match: c > 65535");
        return OPTEX_ESIZE;
    }
    Arena ar(workspace, workspace_bytes);
    uint32_t *minmax = ar.take<uint32_t>(2 * (size_t)c);
    uint32_t *hist = ar.take<uint32_t>(2 * (size_t)c * bins);
    float *tbl = ar.take<float>(4 * (size_t)c * bins + 4);
    if (!ar.ok()) {
        set_error("optex_cdf_match: workspace %zu < %zu", workspace_bytes, optex_cdf_match_workspace_bytes(c, bins));
        return OPTEX_EWORKSPACE;
    }
    const int t_vec = aligned16(target) && (n_t % 4 == 0);
    const int s_vec = aligned16(source) && (n_s % 4 == 0);
    const int o_vec = t_vec && aligned16(out);
    const int64_t n_big = n_t > n_s ? n_t : n_s;
    const bool priv = (bins % 4 == 0) && bins <= PRIV_MAX_BINS;
    // many channels, short rows: one CTA per channel does the whole match (see cdf_channel_kernel), range included
    if (cdf_uses_channel_kernel(c, n_t, n_s, bins)) {
        const size_t smem = sizeof(uint32_t) * 2 * (size_t)bins + 2 * (size_t)(bins + 1) * NTH_HIST;
        static PerDeviceOnce attr_once1;
        OPTEX_TRY(ensure_dyn_smem(attr_once1, cdf_channel_kernel, (int)(sizeof(uint32_t) * 2 * PRIV_MAX_BINS + 2 * (PRIV_MAX_BINS + 1) * NTH_HIST)));
        launch_pdl(cdf_channel_kernel, dim3((unsigned)c), dim3(NT_CH), smem, st, target, source, out, n_t, n_s,
                   (const uint32_t *)minmax, bins, t_vec, s_vec, o_vec, tables, have_range ? 0 : 1);
        OPTEX_LAUNCH_CHECK("cdf_channel_kernel");
        return OPTEX_OK;
    }
    if (!have_range) {
        OPTEX_TRY(fill_u32(minmax, 2 * (int64_t)c, 0xffffffffu, st));
        dim3 grid((unsigned)cdf_splits(c, n_big), (unsigned)c);
        launch_pdl(cdf_range_kernel, grid, dim3(NT), 0, st, target, source, n_t, n_s, minmax, t_vec, s_vec);
        OPTEX_LAUNCH_CHECK("cdf_range_kernel");
    }
    // histogram grid: one CTA per channel and array unless that leaves SMs idle and the slices stay long.  The kernel
    // holds 8 CTAs of 3 warps per SM (25 KB of private counters each): fill that wave - 2 CTAs per SM left the SMs
    // at 20 % warp occupancy and 34 % issue utilisation (ncu, conv1_1 @ 1024^2: 316 us for 537 MB)
    int64_t hs = (8LL * sm_count()) / (2LL * c), hcap = (n_big + 32767) / 32768;
    dim3 grid_h((unsigned)(hs < hcap ? (hs < 1 ? 1 : hs) : (hcap < 1 ? 1 : hcap)), (unsigned)c, 2);
    if (grid_h.x > 1)  // several CTAs add into one histogram: it has to start at zero
        OPTEX_TRY(fill_u32(hist, 2 * (int64_t)c * bins, 0u, st));
    if (priv) {
        size_t smem = sizeof(uint32_t) * (size_t)bins + (size_t)(bins + 1) * NTH_HIST;
        static PerDeviceOnce attr_once2;
        OPTEX_TRY(ensure_dyn_smem(attr_once2, cdf_hist_kernel<true>, (int)(sizeof(uint32_t) * PRIV_MAX_BINS + (PRIV_MAX_BINS + 1) * NTH_HIST)));
        launch_pdl(cdf_hist_kernel<true>, grid_h, dim3(NTH_HIST), smem, st, target, source, n_t, n_s,
                   (const uint32_t *)minmax, hist, bins, t_vec, s_vec);
    } else {
        launch_pdl(cdf_hist_kernel<false>, grid_h, dim3(NTH_HIST), sizeof(uint32_t) * 2 * bins, st, target, source, n_t,
                   n_s, (const uint32_t *)minmax, hist, bins, t_vec, s_vec);
    }
    OPTEX_LAUNCH_CHECK("cdf_hist_kernel");
    launch_pdl(cdf_tables_kernel, dim3((unsigned)c), dim3(NT), sizeof(float) * 6 * bins, st, (const uint32_t *)minmax,
               (const uint32_t *)hist, bins, tbl, tables);
    OPTEX_LAUNCH_CHECK("cdf_tables_kernel");
    dim3 grid_a((unsigned)cdf_splits(c, n_t), (unsigned)c);
    launch_pdl(cdf_apply_kernel, grid_a, dim3(NT), 0, st, target, out, n_t, (const uint32_t *)minmax, (const float *)tbl,
               bins, o_vec);
    OPTEX_LAUNCH_CHECK("cdf_apply_kernel");
    return OPTEX_OK;
}
// ---- the matcher in three separately callable stages, for the pixel-sharded multi-GPU step (sharded.cu): every rank
// holds a slice of the pixels of ALL channels, so the range (2 c words, min) and the histograms (2 c bins counts, sum)
// are all-reduced between the stages and every rank then builds identical tables.  Workspace layout as above.
int cdf_stage_range(const float *target, const float *source, int c, int64_t n_t, int64_t n_s, uint32_t *minmax,
                    bool reset, cudaStream_t st) {
    const int t_vec = aligned16(target) && (n_t % 4 == 0), s_vec = aligned16(source) && (n_s % 4 == 0);
    const int64_t n_big = n_t > n_s ? n_t : n_s;
    if (reset) OPTEX_TRY(fill_u32(minmax, 2 * (int64_t)c, 0xffffffffu, st));
    dim3 grid((unsigned)cdf_splits(c, n_big), (unsigned)c);
    launch_pdl(cdf_range_kernel, grid, dim3(NT), 0, st, target, source, n_t, n_s, minmax, t_vec, s_vec);
    OPTEX_LAUNCH_CHECK("cdf_range_kernel");
    return OPTEX_OK;
}
int cdf_stage_hist(const float *target, const float *source, int c, int64_t n_t, int64_t n_s, int bins,
                   const uint32_t *minmax, uint32_t *hist, cudaStream_t st) {
    const int t_vec = aligned16(target) && (n_t % 4 == 0), s_vec = aligned16(source) && (n_s % 4 == 0);
    const int64_t n_big = n_t > n_s ? n_t : n_s;
    const bool priv = (bins % 4 == 0) && bins <= PRIV_MAX_BINS;
    int64_t hs = (8LL * sm_count()) / (2LL * c), hcap = (n_big + 32767) / 32768;   // one full wave of 8 CTAs per SM
    dim3 grid_h((unsigned)(hs < hcap ? (hs < 1 ? 1 : hs) : (hcap < 1 ? 1 : hcap)), (unsigned)c, 2);
    if (grid_h.x > 1) OPTEX_TRY(fill_u32(hist, 2 * (int64_t)c * bins, 0u, st));
    if (priv) {
        size_t smem = sizeof(uint32_t) * (size_t)bins + (size_t)(bins + 1) * NTH_HIST;
        static PerDeviceOnce attr_once3;
        OPTEX_TRY(ensure_dyn_smem(attr_once3, cdf_hist_kernel<true>, (int)(sizeof(uint32_t) * PRIV_MAX_BINS + (PRIV_MAX_BINS + 1) * NTH_HIST)));
        launch_pdl(cdf_hist_kernel<true>, grid_h, dim3(NTH_HIST), smem, st, target, source, n_t, n_s, minmax, hist, bins,
                   t_vec, s_vec);
    } else {
        launch_pdl(cdf_hist_kernel<false>, grid_h, dim3(NTH_HIST), sizeof(uint32_t) * 2 * bins, st, target, source, n_t,
                   n_s, minmax, hist, bins, t_vec, s_vec);
    }
    OPTEX_LAUNCH_CHECK("cdf_hist_kernel");
    return OPTEX_OK;
}
int cdf_stage_apply(const float *target, float *out, int c, int64_t n_t, int bins, const uint32_t *minmax,
                    const uint32_t *hist, float *tbl, cudaStream_t st) {
    const int o_vec = aligned16(target) && (n_t % 4 == 0) && aligned16(out);
    launch_pdl(cdf_tables_kernel, dim3((unsigned)c), dim3(NT), sizeof(float) * 6 * bins, st, minmax, hist, bins, tbl,
               (float *)nullptr);
    OPTEX_LAUNCH_CHECK("cdf_tables_kernel");
    dim3 grid_a((unsigned)cdf_splits(c, n_t), (unsigned)c);
    launch_pdl(cdf_apply_kernel, grid_a, dim3(NT), 0, st, target, out, n_t, minmax, (const float *)tbl, bins, o_vec);
    OPTEX_LAUNCH_CHECK("cdf_apply_kernel");
    return OPTEX_OK;
}
}  // namespace optex

extern "C" int optex_cdf_match(const float *target, const float *source, float *out, int c, int64_t n_t,
                               int64_t n_s, int bins, float *tables, void *workspace,
                               size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    if (c < 0 || n_t < 0 || n_s < 0) {
        set_error("optex_cdf_match: negative size");
        return OPTEX_EINVAL;
    }
    if (bins < 1 || bins > MAX_BINS) {
        set_error("optex_cdf_match: bins=%d outside [1, %d]", bins, MAX_BINS);
        return OPTEX_EINVAL;
    }
    if (c == 0 || n_t == 0) return OPTEX_OK;  // empty target: nothing to write (torch: empty loop / empty rows)
    if (n_s == 0) {
        set_error("optex_cdf_match: empty source (the reference raises on min() of an empty tensor)");
        return OPTEX_EINVAL;
    }
    if (!target || !source || !out) {
        set_error("optex_cdf_match: NULL pointer");
        return OPTEX_EINVAL;
    }
    return cdf_match_core(target, source, out, c, n_t, n_s, bins, tables, workspace, workspace_bytes, false,
                          (cudaStream_t)stream);
}

extern "C" int optex_interp(const float *x, const float *xp, const float *fp, float *out, int64_t n, int len,
                            void *stream) {
    OPTEX_TRY(require_sm100());
    if (!x || !xp || !fp || !out || n < 0 || len < 1) {
        set_error("optex_interp: NULL pointer, n < 0 or len < 1");
        return OPTEX_EINVAL;
    }
    if (n == 0) return OPTEX_OK;
    interp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, xp, fp, out, n, len);
    OPTEX_LAUNCH_CHECK("interp_kernel");
    return OPTEX_OK;
}
