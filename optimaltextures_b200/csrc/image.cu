// Image-space and once-per-pass glue around the OT loop, sm_100a.
//
//   optex_resize_bicubic_aa   util.py:105-106  interpolate(mode="bicubic", align_corners=False, antialias=True)
//   optex_rgb_to_hls / optex_hls_to_rgb / optex_lightness_transfer      optex.py:126-128 (kornia.color.hls)
//   optex_mix_features        optex.py:197-204 (mask resize + the blend of mix_style_features)
//   optex_recentre            optex.py:76      (content - mean(content) + mean(style), scalar means)
//
// All of it is streaming fp32 work on a few MB per call, once per pass: HBM/latency-bound, plain coalesced kernels.
#include <math.h>

#include "common.cuh"

namespace optex {
namespace {

// ---------------------------------------------------------------------------------------------- bicubic, antialias
// torch's separable antialiased resize: per output index i a window [xmin, xmin + xsize) of input taps with
// weights cubic((j + xmin - center + 0.5) * invscale), a = -0.5, normalised to sum 1; scale = in / out,
// center = scale (i + 0.5), support = 2 scale (down-sampling) or 2, invscale = 1 / scale (down-sampling) or 1.
__device__ __forceinline__ float cubic_aa(float x) {
    const float a = -0.5f;
    x = fabsf(x);
    if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
    if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
    return 0.f;
}

// table per output index: xmin, xsize, weights[taps]
__global__ void resize_table_kernel(int in_size, int out_size, int taps, int *__restrict__ xmin_out,
                                    int *__restrict__ xsize_out, float *__restrict__ weights) {
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= out_size) return;
    const float scale = (float)in_size / (float)out_size;
    const float support = scale >= 1.f ? 2.f * scale : 2.f;
    const float invscale = scale >= 1.f ? 1.f / scale : 1.f;
    const float center = scale * ((float)i + 0.5f);
    int xmin = (int)(center - support + 0.5f);
    if (xmin < 0) xmin = 0;
    int xend = (int)(center + support + 0.5f);
    if (xend > in_size) xend = in_size;
    int xsize = xend - xmin;
    if (xsize < 0) xsize = 0;
    if (xsize > taps) xsize = taps;
    float *w = weights + (size_t)i * taps;
    float total = 0.f;
    for (int j = 0; j < xsize; ++j) {
        const float v = cubic_aa(((float)(j + xmin) - center + 0.5f) * invscale);
        w[j] = v;
        total += v;
    }
    if (total != 0.f)
        for (int j = 0; j < xsize; ++j) w[j] /= total;
    for (int j = xsize; j < taps; ++j) w[j] = 0.f;
    xmin_out[i] = xmin;
    xsize_out[i] = xsize;
}

// horizontal pass: dst[p, y, x] = sum_j w[x][j] src[p, y, xmin[x] + j]      (p = plane = b * c)
__global__ void resize_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, int64_t rows, int w_in,
                                   int w_out, int taps, const int *__restrict__ xmin, const int *__restrict__ xsize,
                                   const float *__restrict__ weights) {
    pdl_wait();
    const int64_t total = rows * w_out;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % w_out);
        const int64_t r = idx / w_out;
        const float *s = src + r * w_in + xmin[x];
        const float *w = weights + (size_t)x * taps;
        const int n = xsize[x];
        float acc = 0.f;
        for (int j = 0; j < n; ++j) acc += s[j] * w[j];
        dst[idx] = acc;
    }
}

// vertical pass: dst[p, y, x] = sum_j w[y][j] src[p, ymin[y] + j, x]
__global__ void resize_cols_kernel(const float *__restrict__ src, float *__restrict__ dst, int64_t planes, int h_in,
                                   int h_out, int w, int taps, const int *__restrict__ ymin,
                                   const int *__restrict__ ysize, const float *__restrict__ weights) {
    pdl_wait();
    const int64_t total = planes * h_out * w;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % w);
        const int64_t t = idx / w;
        const int y = (int)(t % h_out);
        const int64_t p = t / h_out;
        const float *s = src + (p * h_in + ymin[y]) * (int64_t)w + x;
        const float *wt = weights + (size_t)y * taps;
        const int n = ysize[y];
        float acc = 0.f;
        for (int j = 0; j < n; ++j) acc += s[(int64_t)j * w] * wt[j];
        dst[idx] = acc;
    }
}

inline int resize_taps(int in_size, int out_size) {
    const float scale = (float)in_size / (float)out_size;
    const float support = scale >= 1.f ? 2.f * scale : 2.f;
    return (int)ceilf(support) * 2 + 1;
}

// ---------------------------------------------------------------------------------------------- HLS (kornia)
struct Hls {
    float h, l, s;
};

// kornia.color.hls.rgb_to_hls: h in radians [0, 2 pi), NaN (grey / black / white pixels) -> 0
__device__ __forceinline__ Hls rgb2hls(float r, float g, float b) {
    const float maxc = fmaxf(r, fmaxf(g, b));
    const float minc = fminf(r, fminf(g, b));
    const int imax = (r >= g && r >= b) ? 0 : (g >= b ? 1 : 2);  // first maximum, like torch.max's index
    const float sum = maxc + minc;
    const float l = sum / 2.f;
    const float d = maxc - minc;
    float s = l < 0.5f ? d / sum : d / (2.f - sum);
    float hi;
    if (imax == 0) {
        hi = (g - b) / d;
        hi = hi - 6.f * floorf(hi / 6.f);  // python-style modulo
    } else if (imax == 1) {
        hi = (b - r) / d + 2.f;
    } else {
        hi = (r - g) / d + 4.f;
    }
    float h = 2.f * 3.14159265358979323846f * (60.f * hi) / 360.f;
    Hls o;
    o.h = isnan(h) ? 0.f : h;
    o.l = isnan(l) ? 0.f : l;
    o.s = isnan(s) ? 0.f : s;
    return o;
}

// kornia.color.hls.hls_to_rgb: k = (h * 6 / pi + {0, 8, 4}) mod 12 ; out = l - a * max(min(k - 3, 9 - k, 1), -1)
__device__ __forceinline__ float hls_channel(float h12, float l, float a, float off) {
    float k = h12 + off;
    k = k - 12.f * floorf(k / 12.f);
    const float t = fminf(fminf(k - 3.f, 9.f - k), 1.f);
    return l - a * fmaxf(t, -1.f);
}
__device__ __forceinline__ void hls2rgb(Hls v, float &r, float &g, float &b) {
    const float h12 = v.h * (6.f / 3.14159265358979323846f);
    const float a = v.s * fminf(v.l, 1.f - v.l);
    r = hls_channel(h12, v.l, a, 0.f);
    g = hls_channel(h12, v.l, a, 8.f);
    b = hls_channel(h12, v.l, a, 4.f);
}

// op 0: rgb -> hls, 1: hls -> rgb, 2: out = hls_to_rgb(h(a), l(b2), s(a))  (optex.py:126-128)
__global__ void hls_kernel(const float *__restrict__ a, const float *__restrict__ b2, float *__restrict__ out, int b,
                           int64_t hw, int op) {
    pdl_wait();
    const int64_t total = (int64_t)b * hw;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t img = idx / hw, px = idx % hw;
        const int64_t o = img * 3 * hw + px;
        const float x0 = a[o], x1 = a[o + hw], x2 = a[o + 2 * hw];
        float y0, y1, y2;
        if (op == 0) {
            const Hls v = rgb2hls(x0, x1, x2);
            y0 = v.h; y1 = v.l; y2 = v.s;
        } else if (op == 1) {
            Hls v{x0, x1, x2};
            hls2rgb(v, y0, y1, y2);
        } else {
            Hls v = rgb2hls(x0, x1, x2);
            v.l = rgb2hls(b2[o], b2[o + hw], b2[o + 2 * hw]).l;
            hls2rgb(v, y0, y1, y2);
        }
        out[o] = y0; out[o + hw] = y1; out[o + 2 * hw] = y2;
    }
}

// ---------------------------------------------------------------------------------------------- style mixing
// optex.py:197-204 for one layer.  A, B, AtoB, BtoA, out: [h, w, c]; mask [mh, mw] is resized with
// interpolate(mode="nearest") (source index = min(floor(dst * (in / out)), in - 1), scale in fp32).
__global__ void mix_kernel(const float *__restrict__ A, const float *__restrict__ B, const float *__restrict__ AtoB,
                           const float *__restrict__ BtoA, const float *__restrict__ mask, float *__restrict__ out,
                           int h, int w, int c, int mh, int mw, float alpha) {
    pdl_wait();
    const int64_t total = (int64_t)h * w * c;
    const float sy = (float)mh / (float)h, sx = (float)mw / (float)w;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t px = idx / c;
        const int y = (int)(px / w), x = (int)(px % w);
        int my = (int)floorf((float)y * sy), mx = (int)floorf((float)x * sx);
        if (my > mh - 1) my = mh - 1;
        if (mx > mw - 1) mx = mw - 1;
        const float m = mask[(int64_t)my * mw + mx];
        const float i = alpha;
        out[idx] = (A[idx] * (1.f - i) + AtoB[idx] * i) * m + (BtoA[idx] * (1.f - i) + B[idx] * i) * (1.f - m);
    }
}

// ---------------------------------------------------------------------------------------------- scalar means
// Two-stage deterministic FP64 sum: partials[blockIdx.x] then one block folds them.
__global__ void sum_partial_kernel(const float *__restrict__ x, int64_t n, double *__restrict__ partial) {
    pdl_wait();
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += (double)x[i];
    __shared__ double sh[32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}

// shift[0] = mean(style) - mean(content) from the two partial arrays (one warp)
__global__ void recentre_shift_kernel(const double *__restrict__ pc, int nc, int64_t n_content,
                                      const double *__restrict__ ps, int ns, int64_t n_style,
                                      float *__restrict__ shift) {
    pdl_wait();
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nc; i += 32) a += pc[i];
    for (int i = threadIdx.x; i < ns; i += 32) b += ps[i];
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) {
        // the reference subtracts the fp32 mean and then adds the other fp32 mean (optex.py:76)
        shift[0] = (float)(a / (double)n_content);
        shift[1] = (float)(b / (double)n_style);
    }
}

__global__ void recentre_apply_kernel(const float *__restrict__ x, float *__restrict__ out, int64_t n,
                                      const float *__restrict__ shift) {
    pdl_wait();
    const float mc = shift[0], ms = shift[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = x[i] - mc + ms;
}

inline unsigned grid_for(int64_t items, int threads = 256) {
    int64_t blocks = (items + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

constexpr int kSumBlocks = 592;  // 4 per SM

}  // namespace
}  // namespace optex

using namespace optex;

extern "C" size_t optex_resize_workspace_bytes(int planes, int h_in, int w_in, int h_out, int w_out) {
    if (planes < 1 || h_in < 1 || w_in < 1 || h_out < 1 || w_out < 1) return 0;
    const int tw = resize_taps(w_in, w_out), th = resize_taps(h_in, h_out);
    size_t bytes = align_up((size_t)planes * h_in * w_out * 4, 256);          // after the horizontal pass
    bytes += 2 * align_up((size_t)w_out * 4, 256) + align_up((size_t)w_out * tw * 4, 256);
    bytes += 2 * align_up((size_t)h_out * 4, 256) + align_up((size_t)h_out * th * 4, 256);
    return bytes + 256;
}

extern "C" int optex_resize_bicubic_aa(const float *src, float *dst, int planes, int h_in, int w_in, int h_out,
                                       int w_out, void *workspace, size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    if (!src || !dst || planes < 1 || h_in < 1 || w_in < 1 || h_out < 1 || w_out < 1) {
        set_error("optex_resize_bicubic_aa: NULL pointer or empty shape");
        return OPTEX_EINVAL;
    }
    const int tw = resize_taps(w_in, w_out), th = resize_taps(h_in, h_out);
    Arena ar(workspace, workspace_bytes);
    float *tmp = ar.take<float>((size_t)planes * h_in * w_out);
    int *xmin = ar.take<int>(w_out), *xsize = ar.take<int>(w_out);
    float *xw = ar.take<float>((size_t)w_out * tw);
    int *ymin = ar.take<int>(h_out), *ysize = ar.take<int>(h_out);
    float *yw = ar.take<float>((size_t)h_out * th);
    if (!ar.ok()) {
        set_error("optex_resize_bicubic_aa: workspace %zu < %zu bytes", workspace_bytes,
                  optex_resize_workspace_bytes(planes, h_in, w_in, h_out, w_out));
        return OPTEX_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    launch_pdl(resize_table_kernel, dim3((w_out + 127) / 128), dim3(128), 0, st, w_in, w_out, tw, xmin, xsize, xw);
    OPTEX_LAUNCH_CHECK("resize_table_kernel");
    launch_pdl(resize_table_kernel, dim3((h_out + 127) / 128), dim3(128), 0, st, h_in, h_out, th, ymin, ysize, yw);
    OPTEX_LAUNCH_CHECK("resize_table_kernel");
    // width first, then height (the order of torch's separable CPU kernel)
    const int64_t rows = (int64_t)planes * h_in;
    launch_pdl(resize_rows_kernel, dim3(grid_for(rows * w_out)), dim3(256), 0, st, src, tmp, rows, w_in, w_out, tw,
               (const int *)xmin, (const int *)xsize, (const float *)xw);
    OPTEX_LAUNCH_CHECK("resize_rows_kernel");
    launch_pdl(resize_cols_kernel, dim3(grid_for((int64_t)planes * h_out * w_out)), dim3(256), 0, st,
               (const float *)tmp, dst, (int64_t)planes, h_in, h_out, w_out, th, (const int *)ymin,
               (const int *)ysize, (const float *)yw);
    OPTEX_LAUNCH_CHECK("resize_cols_kernel");
    return OPTEX_OK;
}

static int hls_call(const float *a, const float *b2, float *out, int b, int64_t hw, int op, void *stream,
                    const char *name) {
    OPTEX_TRY(require_sm100());
    if (!a || !out || (op == 2 && !b2) || b < 1 || hw < 1) {
        set_error("%s: NULL pointer or empty shape", name);
        return OPTEX_EINVAL;
    }
    launch_pdl(hls_kernel, dim3(grid_for((int64_t)b * hw)), dim3(256), 0, (cudaStream_t)stream, a, b2, out, b, hw, op);
    OPTEX_LAUNCH_CHECK("hls_kernel");
    return OPTEX_OK;
}

extern "C" int optex_rgb_to_hls(const float *rgb, float *hls, int b, int64_t hw, void *stream) {
    return hls_call(rgb, nullptr, hls, b, hw, 0, stream, "optex_rgb_to_hls");
}
extern "C" int optex_hls_to_rgb(const float *hls, float *rgb, int b, int64_t hw, void *stream) {
    return hls_call(hls, nullptr, rgb, b, hw, 1, stream, "optex_hls_to_rgb");
}
extern "C" int optex_lightness_transfer(const float *content, const float *pastiche, float *out, int b, int64_t hw,
                                        void *stream) {
    return hls_call(content, pastiche, out, b, hw, 2, stream, "optex_lightness_transfer");
}

extern "C" int optex_mix_features(const float *A, const float *B, const float *AtoB, const float *BtoA,
                                  const float *mask, float *out, int h, int w, int c, int mask_h, int mask_w,
                                  float alpha, void *stream) {
    OPTEX_TRY(require_sm100());
    if (!A || !B || !AtoB || !BtoA || !mask || !out || h < 1 || w < 1 || c < 1 || mask_h < 1 || mask_w < 1) {
        set_error("optex_mix_features: NULL pointer or empty shape");
        return OPTEX_EINVAL;
    }
    launch_pdl(mix_kernel, dim3(grid_for((int64_t)h * w * c)), dim3(256), 0, (cudaStream_t)stream, A, B, AtoB, BtoA,
               mask, out, h, w, c, mask_h, mask_w, alpha);
    OPTEX_LAUNCH_CHECK("mix_kernel");
    return OPTEX_OK;
}

extern "C" size_t optex_recentre_workspace_bytes(void) { return (size_t)2 * kSumBlocks * 8 + 512; }

extern "C" int optex_recentre(const float *content, int64_t n_content, const float *style, int64_t n_style,
                              float *out, void *workspace, size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    if (!content || !style || !out || n_content < 1 || n_style < 1) {
        set_error("optex_recentre: NULL pointer or empty tensor");
        return OPTEX_EINVAL;
    }
    Arena ar(workspace, workspace_bytes);
    double *pc = ar.take<double>(kSumBlocks), *ps = ar.take<double>(kSumBlocks);
    float *shift = ar.take<float>(2);
    if (!ar.ok()) {
        set_error("optex_recentre: workspace %zu < %zu bytes", workspace_bytes, optex_recentre_workspace_bytes());
        return OPTEX_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    launch_pdl(sum_partial_kernel, dim3(kSumBlocks), dim3(256), 0, st, content, n_content, pc);
    OPTEX_LAUNCH_CHECK("sum_partial_kernel");
    launch_pdl(sum_partial_kernel, dim3(kSumBlocks), dim3(256), 0, st, style, n_style, ps);
    OPTEX_LAUNCH_CHECK("sum_partial_kernel");
    launch_pdl(recentre_shift_kernel, dim3(1), dim3(32), 0, st, (const double *)pc, kSumBlocks, n_content,
               (const double *)ps, kSumBlocks, n_style, shift);
    OPTEX_LAUNCH_CHECK("recentre_shift_kernel");
    launch_pdl(recentre_apply_kernel, dim3(grid_for(n_content)), dim3(256), 0, st, content, out, n_content,
               (const float *)shift);
    OPTEX_LAUNCH_CHECK("recentre_apply_kernel");
    return OPTEX_OK;
}
