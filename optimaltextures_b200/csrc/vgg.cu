// VGG-19 encoder / feature-inverter layers: the callers either side of the OT loop,
// `encoder(pastiche)` optex.py:107 and `decoder(pastiche_feature)` optex.py:122 (modules: vgg.py:14-136).
//
// Every layer of both networks is   [pre-op] -> ReflectionPad2d(1) -> Conv2d 3x3 -> [ReLU]   with pre-op one of
//   none | MaxPool2d(2, 2, ceil_mode=True) (vgg.py:27,36,51,68) | UpsamplingNearest2d(2) (vgg.py:82,98,114,122)
// (the encoder's leading 1x1 colour conv, vgg.py:16, is folded into conv1_1's weights by the host code).
//
// B200 formulation: IMPLICIT GEMM on the tcgen05 kernel of the rotations (every layer with c_in % 32 == 0).  One pass
// (pad_reflect_kernel) materialises the reflection-padded input - the pooling window or the nearest-neighbour
// up-sampling resolved in its index arithmetic - and the GEMM's TMA producer gathers the nine taps from it with
// shifted 4-D boxes: out[pixel, co] = sum_k A(pixel, k) W[co, k], k = tap * c_in + ci, A never materialised
// (gemm_tcgen05.cu, Params::conv).  Bias and ReLU ride in the epilogue, which writes NHWC - the layout the OT loop
// wants (the reference permutes, vgg.py:153 `to_nhwc`).  The encoder's conv1_1 (c_in = 3) keeps the explicit form: a
// gather kernel writes the im2col matrix col[m, k] in 512 MB row chunks and the same GEMM reads it.
// Layouts: activations NHWC [b, h, w, c] fp32; the very first layer may read the NCHW image directly (src_nchw).
// Arithmetic: the library's GEMM mode (3xTF32 by default: fp32-grade).
#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace optex {
namespace {

struct ConvGeom {
    int b, hs, ws, cin;  // source tensor
    int h, w;            // after the pre-op = output size
    int pre;             // OPTEX_PRE_*
    int kp;              // row pitch of col (9 * cin rounded up to 32)
    int src_nchw;
};

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// source value at (pre-op'd) position (yy, xx), channel c
__device__ __forceinline__ float sample1(const float *__restrict__ src, const ConvGeom &g, int bi, int yy, int xx,
                                         int c) {
    auto at = [&](int y, int x) -> float {
        const int64_t i = g.src_nchw ? (((int64_t)bi * g.cin + c) * g.hs + y) * g.ws + x
                                     : (((int64_t)bi * g.hs + y) * g.ws + x) * g.cin + c;
        return __ldg(src + i);
    };
    if (g.pre == OPTEX_PRE_UPSAMPLE2) return at(yy >> 1, xx >> 1);
    if (g.pre == OPTEX_PRE_MAXPOOL2) {
        const int y0 = 2 * yy, x0 = 2 * xx;
        const bool y1 = y0 + 1 < g.hs, x1 = x0 + 1 < g.ws;  // ceil_mode: the last window may be clipped
        float v = at(y0, x0);
        if (x1) v = fmaxf(v, at(y0, x0 + 1));
        if (y1) v = fmaxf(v, at(y0 + 1, x0));
        if (y1 && x1) v = fmaxf(v, at(y0 + 1, x0 + 1));
        return v;
    }
    return at(yy, xx);
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ float4 sample4(const float *__restrict__ src, const ConvGeom &g, int bi, int yy, int xx,
                                          int c) {
    auto at = [&](int y, int x) -> float4 {
        return __ldg(reinterpret_cast<const float4 *>(src + (((int64_t)bi * g.hs + y) * g.ws + x) * g.cin + c));
    };
    if (g.pre == OPTEX_PRE_UPSAMPLE2) return at(yy >> 1, xx >> 1);
    if (g.pre == OPTEX_PRE_MAXPOOL2) {
        const int y0 = 2 * yy, x0 = 2 * xx;
        const bool y1 = y0 + 1 < g.hs, x1 = x0 + 1 < g.ws;
        float4 v = at(y0, x0);
        if (x1) v = max4(v, at(y0, x0 + 1));
        if (y1) v = max4(v, at(y0 + 1, x0));
        if (y1 && x1) v = max4(v, at(y0 + 1, x0 + 1));
        return v;
    }
    return at(yy, xx);
}

// col[ml, tap * cin + ci] for the rows m0 .. m0 + mc - 1;  VEC: NHWC source with cin % 4 == 0, one float4 per thread
template <bool VEC>
__global__ void __launch_bounds__(256) im2col3x3_kernel(const float *__restrict__ src, float *__restrict__ col,
                                                        ConvGeom g, int64_t m0, int64_t mc) {
    pdl_wait();
    const int per_row = VEC ? 9 * (g.cin >> 2) : g.kp;
    const int64_t total = mc * per_row;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t ml = idx / per_row;
        const int t = (int)(idx - ml * per_row);
        const int64_t m = m0 + ml;
        const int hw = g.h * g.w;
        const int bi = (int)(m / hw);
        const int rem = (int)(m - (int64_t)bi * hw);
        const int y = rem / g.w, x = rem - y * g.w;
        if (VEC) {
            const int c4n = g.cin >> 2;
            const int tap = t / c4n, c = (t - tap * c4n) << 2;
            const int yy = reflect1(y + tap / 3 - 1, g.h), xx = reflect1(x + tap % 3 - 1, g.w);
            *reinterpret_cast<float4 *>(col + ml * g.kp + tap * g.cin + c) = sample4(src, g, bi, yy, xx, c);
        } else {
            float v = 0.f;
            if (t < 9 * g.cin) {
                const int tap = t / g.cin, c = t - tap * g.cin;
                const int yy = reflect1(y + tap / 3 - 1, g.h), xx = reflect1(x + tap % 3 - 1, g.w);
                v = sample1(src, g, bi, yy, xx, c);
            }
            col[ml * g.kp + t] = v;
        }
    }
}

// pad[bi, yp, xp, :] = the pre-op'd source at the reflected position (yp - 1, xp - 1): ReflectionPad2d(1) of the pooled /
// up-sampled / plain NHWC tensor, materialised ONCE (1.0x the layer's input; the im2col matrix is 9x) for the implicit
// GEMM, whose TMA boxes address it with the taps' offsets.  cin % 4 == 0.
__global__ void __launch_bounds__(256) pad_reflect_kernel(const float *__restrict__ src, float *__restrict__ pad,
                                                          ConvGeom g) {
    pdl_wait();
    const int c4n = g.cin >> 2, hp = g.h + 2, wp = g.w + 2;
    const int64_t total = (int64_t)g.b * hp * wp * c4n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx % c4n) << 2;
        int64_t r = idx / c4n;
        const int xp = (int)(r % wp);
        r /= wp;
        const int yp = (int)(r % hp);
        const int bi = (int)(r / hp);
        *reinterpret_cast<float4 *>(pad + idx * 4) = sample4(src, g, bi, reflect1(yp - 1, g.h), reflect1(xp - 1, g.w), c);
    }
}

// dst[b, c, hw] = src[b, hw, 0:c] (row pitch c_src >= c)
__global__ void nhwc_to_nchw_kernel(const float *__restrict__ src, float *__restrict__ dst, int64_t hw, int c_src,
                                    int c, int64_t total) {
    pdl_wait();
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = idx % hw;
        const int64_t bc = idx / hw;
        const int ci = (int)(bc % c);
        const int64_t bi = bc / c;
        dst[idx] = __ldg(src + (bi * hw + p) * c_src + ci);
    }
}

inline int kpad(int cin) { return (9 * cin + 31) / 32 * 32; }
inline void out_size(int hs, int ws, int pre, int *h, int *w) {
    if (pre == OPTEX_PRE_MAXPOOL2) {
        *h = (hs + 1) / 2;
        *w = (ws + 1) / 2;
    } else if (pre == OPTEX_PRE_UPSAMPLE2) {
        *h = 2 * hs;
        *w = 2 * ws;
    } else {
        *h = hs;
        *w = ws;
    }
}
// implicit GEMM applies to NHWC sources with whole 32-channel k blocks (every layer but conv1_1 of the encoder)
bool implicit_ok(int c_in, int src_nchw) {
    static const char *env = getenv("OPTEX_CONV_IMPLICIT");
    return !src_nchw && c_in >= 32 && c_in % 32 == 0 && !(env && atoi(env) == 0);
}
int64_t chunk_rows(int64_t M, int kp) {
    static const char *env = getenv("OPTEX_CONV_CHUNK_MB");
    // Round 1 kept the im2col chunk at 32 MB so it stayed in L2; that made every GEMM tiny (M = 1820 rows for the
    // 512-channel layers: 64-wide tiles, A converted 8 times).  512 MB chunks go through HBM once (a 10-25 % add-on to
    // the GEMM's own time) and let the GEMMs take the paired 256-wide tiles: Encoder(5) @ 1024^2 15.7 -> 8.1 ms,
    // Decoder(5) 19.8 -> 9.9 ms.
    const int64_t budget = (env && atoi(env) > 0 ? (int64_t)atoi(env) : 512) << 20;
    int64_t rows = budget / ((int64_t)kp * 4) / 128 * 128;
    if (rows < 128) rows = 128;
    return rows < M ? rows : M;
}

}  // namespace
}  // namespace optex

using namespace optex;

extern "C" int optex_conv3x3_packed_k(int c_in) { return c_in > 0 ? kpad(c_in) : 0; }

extern "C" size_t optex_conv3x3_workspace_bytes(int b, int h_src, int w_src, int c_in, int c_out, int pre_op) {
    if (b < 1 || h_src < 1 || w_src < 1 || c_in < 1 || c_out < 1 || pre_op < 0 || pre_op > 2) return 0;
    int h, w;
    out_size(h_src, w_src, pre_op, &h, &w);
    const int kp = kpad(c_in);
    const int64_t M = (int64_t)b * h * w;
    size_t col = (size_t)chunk_rows(M, kp) * kp * 4;
    if (implicit_ok(c_in, 0)) {   // implicit GEMM: the padded input instead of an im2col chunk
        const size_t pad = (size_t)b * (h + 2) * (w + 2) * c_in * 4;
        if (pad > col) col = pad;
    }
    return align_up(col, 256) + align_up((size_t)2 * c_out * kp * 4, 256) + 256;
}

extern "C" int optex_conv3x3(const float *src, int src_nchw, int b, int h_src, int w_src, int c_in,
                             const float *weight, const float *bias, int c_out, int pre_op, int relu, float *dst,
                             int64_t ldd, void *workspace, size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    if (!src || !weight || !dst || b < 1 || h_src < 1 || w_src < 1 || c_in < 1 || c_out < 1 || pre_op < 0 ||
        pre_op > 2 || ldd < c_out) {
        set_error("optex_conv3x3: NULL pointer, empty shape, unknown pre-op or ldd < c_out");
        return OPTEX_EINVAL;
    }
    int h, w;
    out_size(h_src, w_src, pre_op, &h, &w);
    if (h < 2 || w < 2) {
        set_error("optex_conv3x3: ReflectionPad2d(1) needs at least 2 x 2 pixels (got %d x %d)", h, w);
        return OPTEX_ESIZE;
    }
    const int kp = kpad(c_in);
    const int64_t M = (int64_t)b * h * w;
    if ((int64_t)b * h_src * w_src * c_in > 0x7fffffffffLL || M > 0x3fffffffLL) {
        set_error("optex_conv3x3: tensor too large");
        return OPTEX_ESIZE;
    }
    const int64_t rows = chunk_rows(M, kp);
    Arena ar(workspace, workspace_bytes);
    size_t col_elems = (size_t)rows * kp;
    if (implicit_ok(c_in, src_nchw)) {   // the same region holds the padded input of the implicit GEMM
        const size_t pad_elems = (size_t)b * (h + 2) * (w + 2) * c_in;
        if (pad_elems > col_elems) col_elems = pad_elems;
    }
    float *col = ar.take<float>(col_elems);
    float *wsplit = ar.take<float>((size_t)2 * c_out * kp);
    if (!ar.ok()) {
        set_error("optex_conv3x3: workspace %zu < %zu bytes", workspace_bytes,
                  optex_conv3x3_workspace_bytes(b, h_src, w_src, c_in, c_out, pre_op));
        return OPTEX_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    ConvGeom g{b, h_src, w_src, c_in, h, w, pre_op, kp, src_nchw ? 1 : 0};
    // the vector gather writes whole taps only: it needs kp == 9 c_in (no padding columns to zero)
    const bool vec = !src_nchw && c_in % 4 == 0 && kp == 9 * c_in && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    const int mode = optex_get_gemm_mode();
    const bool want_tc = mode != OPTEX_GEMM_FP32;
    const int terms = mode == OPTEX_GEMM_TF32 ? 1 : 3;
    bool presplit = false;
    // the tensor-core path wants 32-byte aligned rows of D and whole 32-byte sectors per row
    const bool tc_shape = ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
    if (want_tc && tc_shape && terms == 3 && rows > 4096) {  // weights -> tf32 hi / lo once for all chunks
        const int64_t nw = (int64_t)c_out * kp;
        OPTEX_TRY(gemm_tc_split_and_fill(weight, wsplit, wsplit + nw, nw, nullptr, 0, 0u, st));
        gemm_tc_set_presplit(weight, wsplit, wsplit + nw);
        presplit = true;
    }
    // ---- implicit GEMM: pad once, the GEMM's TMA boxes gather the taps (no im2col matrix)
    if (want_tc && tc_shape && terms == 3 && implicit_ok(c_in, src_nchw) && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        float *padded = col;   // the workspace's first region holds the padded input instead of an im2col chunk
        const int64_t nw = (int64_t)c_out * kp;
        if (!presplit) OPTEX_TRY(gemm_tc_split_and_fill(weight, wsplit, wsplit + nw, nw, nullptr, 0, 0u, st));
        const int64_t items = (int64_t)b * (h + 2) * (w + 2) * (c_in / 4);
        int64_t blocks = (items + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 32;
        if (blocks > cap) blocks = cap;
        launch_pdl(pad_reflect_kernel, dim3((unsigned)blocks), dim3(256), 0, st, src, padded, g);
        OPTEX_LAUNCH_CHECK("pad_reflect_kernel");
        TcConv cv{padded, b, h, w, c_in, wsplit, wsplit + nw, c_out, bias, relu != 0, dst, ldd};
        int rc_i = gemm_tc_conv(cv, st);
        if (rc_i != OPTEX_ENOTSUP) {
            if (presplit) gemm_tc_set_presplit(nullptr, nullptr, nullptr);
            return rc_i;
        }
    }
    int rc = OPTEX_OK;
    for (int64_t m0 = 0; m0 < M && rc == OPTEX_OK; m0 += rows) {
        const int64_t mc = M - m0 < rows ? M - m0 : rows;
        const int64_t items = mc * (vec ? 9 * (c_in / 4) : kp);
        int64_t blocks = (items + 255) / 256;
        const int64_t cap = (int64_t)sm_count() * 32;
        if (blocks > cap) blocks = cap;
        if (vec)
            launch_pdl(im2col3x3_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, st, src, col, g, m0, mc);
        else
            launch_pdl(im2col3x3_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, st, src, col, g, m0, mc);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            rc = cuda_fail(e, "im2col3x3_kernel");
            break;
        }
        float *d = dst + m0 * ldd;
        rc = OPTEX_ENOTSUP;
        if (want_tc && tc_shape) {
            TcGemm t{};
            t.A = col; t.a_mn = false; t.B = weight; t.b_mn = false; t.D = d; t.ldd = ldd;
            t.M = mc; t.N = c_out; t.K = kp; t.terms = terms; t.alpha = 1.f;
            t.bias = bias; t.bias_hw = 1; t.bias_ld = 0; t.relu = relu != 0; t.stream_a = true;
            rc = gemm_tc(t, st);
        }
        if (rc == OPTEX_ENOTSUP) {
            if (mode == OPTEX_GEMM_TF32X3 || mode == OPTEX_GEMM_TF32) {
                set_error("optex_conv3x3: c_out = %d / ldd = %lld is outside the tensor-core path's constraints "
                          "(ldd %% 8, 32-byte aligned dst); use gemm mode auto or fp32", c_out, (long long)ldd);
                rc = OPTEX_ESIZE;
                break;
            }
            SimtOpts o;
            o.bias = bias; o.bias_hw = 1; o.bias_ld = 0; o.relu = relu != 0;
            rc = sgemm_simt_ex(col, kp, true, weight, kp, true, d, ldd, false, mc, c_out, kp, nullptr, 0.f, 1.f, o, st);
        }
    }
    if (presplit) gemm_tc_set_presplit(nullptr, nullptr, nullptr);
    return rc;
}

extern "C" int optex_nhwc_to_nchw(const float *src, float *dst, int b, int64_t hw, int c_src, int c, void *stream) {
    OPTEX_TRY(require_sm100());
    if (!src || !dst || b < 1 || hw < 1 || c < 1 || c_src < c) {
        set_error("optex_nhwc_to_nchw: NULL pointer, empty shape or c_src < c");
        return OPTEX_EINVAL;
    }
    const int64_t total = (int64_t)b * c * hw;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    launch_pdl(nhwc_to_nchw_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, src, dst, hw, c_src, c,
               total);
    OPTEX_LAUNCH_CHECK("nhwc_to_nchw_kernel");
    return OPTEX_OK;
}
