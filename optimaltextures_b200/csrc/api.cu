// C-ABI of liboptex_b200 (include/optex_b200.h): library state, argument checking and the
// composition of the kernels into the reference's call surface
//   optimal_transport()  optex.py:167-177      -> optex_ot_step / optex_ot_step_host
//   inner loop           optex.py:112-117      -> optex_ot_loop
//   hist_match()         histmatch.py:5-46     -> optex_hist_match
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace optex {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_gemm_mode{OPTEX_GEMM_AUTO};
static std::atomic<int> g_pdl{1};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return OPTEX_ECUDA;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static int g_dev_ok[64];  // 0 unknown, 1 sm_100, -1 other
static int g_dev_sms[64];
int require_sm100() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cuda_fail(e, "cudaGetDevice (is there a GPU?)");
        return OPTEX_EDEVICE;
    }
    if (dev < 0 || dev >= 64) dev = 63;
    if (g_dev_ok[dev] == 0) {
        int major = 0, minor = 0, sms = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        g_dev_sms[dev] = sms > 0 ? sms : 148;
        g_dev_ok[dev] = (major == 10 && minor == 0) ? 1 : -1;
        if (g_dev_ok[dev] < 0)
            set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only and has no fallback",
                      dev, major, minor);
    }
    if (g_dev_ok[dev] < 0) {
        if (!g_err[0]) set_error("current device is not sm_100 (B200); no fallback exists");
        return OPTEX_EDEVICE;
    }
    return OPTEX_OK;
}
int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    return g_dev_sms[dev] > 0 ? g_dev_sms[dev] : 148;
}

// ---- rotation GEMMs -----------------------------------------------------------
static bool want_tc(bool &forced) {
    int m = g_gemm_mode.load();
    forced = (m == OPTEX_GEMM_TF32X3 || m == OPTEX_GEMM_TF32);
    return m != OPTEX_GEMM_FP32;
}
static int tc_terms() { return g_gemm_mode.load() == OPTEX_GEMM_TF32 ? 1 : 3; }

// Xt[c, n] = (X R)^T   or  XR[n, c] = X R            optex.py:170-171
// colrange (optional): fold the per-channel range into the GEMM epilogue; *range_done reports whether it happened
static int rotate_forward(const float *X, const float *R, float *dst, int64_t n, int c, bool transposed,
                          cudaStream_t st, int c0 = 0, int nc = -1, uint32_t *colrange = nullptr,
                          bool *range_done = nullptr) {
    if (nc < 0) nc = c;
    if (range_done) *range_done = false;
    bool forced;
    if (want_tc(forced)) {
        int rc = gemm_tc_rotate_forward(X, R, dst, n, c, transposed, tc_terms(), st, c0, nc, colrange);
        if (rc == OPTEX_OK && range_done) *range_done = colrange != nullptr;
        if (rc != OPTEX_ENOTSUP) return rc;
        if (forced) {
            set_error("rotation GEMM: shape (n=%lld, c=%d, block %d+%d) is outside the tensor-core path's TMA "
                      "constraints (c %% 32, 16-byte alignment); use gemm mode auto or fp32", (long long)n, c, c0, nc);
            return OPTEX_ESIZE;
        }
    }
    // D[m = pixel, j] = sum_k X[m, k] R[k, c0 + j]
    return sgemm_simt(X, c, true, R + c0, c, false, dst, transposed ? n : nc, transposed, n, nc, c, nullptr, 0.f, 1.f,
                      st);
}
// out[n, j] = sum_c M(n, c) R[j, c]  (+ blend)         optex.py:175, :117
static int rotate_inverse(const float *M, bool m_channel_major, const float *R, float *out, int64_t n, int c,
                          const float *content, float strength, cudaStream_t st) {
    bool forced;
    if (want_tc(forced)) {
        int rc = gemm_tc_rotate_inverse(M, m_channel_major, R, out, n, c, content, strength, tc_terms(), st);
        if (rc != OPTEX_ENOTSUP) return rc;
        if (forced) {
            set_error("rotation GEMM: shape (n=%lld, c=%d) is outside the tensor-core path's TMA constraints "
                      "(c %% 32, n %% 32, 16-byte alignment); use gemm mode auto or fp32", (long long)n, c);
            return OPTEX_ESIZE;
        }
    }
    return sgemm_simt(M, m_channel_major ? n : c, !m_channel_major, R, c, true, out, c, false, n, c, c, content,
                      strength, 1.f, st);
}

// Stage profiler of optex_ot_step_profile: events between the step's launches (the caller turns PDL off)
struct StageProf {
    cudaEvent_t ev[OPTEX_MAX_STAGES + 1];
    uint64_t launches[OPTEX_MAX_STAGES + 1];
    int n = 0;
};
static thread_local StageProf *g_prof = nullptr;
static void prof_mark(cudaStream_t st) {
    if (!g_prof || g_prof->n > OPTEX_MAX_STAGES) return;
    cudaEventRecord(g_prof->ev[g_prof->n], st);
    g_prof->launches[g_prof->n] = g_launches.load();
    ++g_prof->n;
}

static bool per_channel(int mode) { return mode == OPTEX_MODE_CDF || mode == OPTEX_MODE_SORT; }
static bool valid_mode(int mode) { return mode >= OPTEX_MODE_CHOL && mode <= OPTEX_MODE_SORT; }

static size_t match_ws_bytes(int64_t n_p, int64_t n_s, int c, int mode) {
    if (mode == OPTEX_MODE_CDF) return optex_cdf_match_workspace_bytes(c, 256);
    if (mode == OPTEX_MODE_SORT) return sort_match_scratch_bytes(c, n_p, n_s) + 256;  // rs itself is sorted in place
    return cov_match_ws_bytes(n_p, n_s, c, mode);
}

static int check_step_args(const char *fn, const void *P, const void *S, const void *out, int b_p, int64_t hw_p,
                           int b_s, int64_t hw_s, int c, int mode) {
    if (!P || !S || !out) {
        set_error("%s: NULL feature pointer", fn);
        return OPTEX_EINVAL;
    }
    if (b_p < 1 || b_s < 1 || hw_p < 1 || hw_s < 1 || c < 1) {
        set_error("%s: empty or negative shape (b_p=%d hw_p=%lld b_s=%d hw_s=%lld c=%d)", fn, b_p,
                  (long long)hw_p, b_s, (long long)hw_s, c);
        return OPTEX_EINVAL;
    }
    if (!valid_mode(mode)) {
        set_error("%s: unknown mode %d", fn, mode);
        return OPTEX_EINVAL;
    }
    if (!per_channel(mode) && !(b_s == 1 || b_s == b_p)) {
        set_error("%s: covariance modes need b_s == 1 or b_s == b_p (histmatch.py:44 broadcast), got %d vs %d",
                  fn, b_s, b_p);
        return OPTEX_EINVAL;
    }
    return OPTEX_OK;
}

// tf32 hi / lo planes [2][c][c] of the rotation the next ot_step_impl call receives (optex_ot_steps with R_split)
static thread_local const float *g_r_split = nullptr;

// The step proper.  P may alias out (P is dead once the forward rotation has run).
static int ot_step_impl(const float *P, const float *S, const float *R, float *out, int b_p, int64_t hw_p,
                        int b_s, int64_t hw_s, int c, int mode, float eps, const float *content, float strength,
                        void *ws, size_t ws_bytes, cudaStream_t st, int style_reuse = 0) {
    const int64_t n_p = (int64_t)b_p * hw_p, n_s = (int64_t)b_s * hw_s;
    if (!per_channel(mode)) {  // closed-form modes: rotation folded into C x C products (cov_match.cu)
        prof_mark(st);
        int rc = cov_ot_step(P, S, R, out, b_p, hw_p, b_s, hw_s, c, mode, eps, content, strength, ws, ws_bytes, st,
                             style_reuse);
        prof_mark(st);
        return rc;
    }
    Arena ar(ws, ws_bytes);
    float *rp = ar.take<float>((size_t)n_p * c);
    float *rs = ar.take<float>((size_t)n_s * c);
    float *mt = rp;
    float *r_hi = ar.take<float>((size_t)c * c), *r_lo = ar.take<float>((size_t)c * c);
    size_t mws = match_ws_bytes(n_p, n_s, c, mode);
    void *mw = ar.take<char>(mws);
    if (!ar.ok()) {
        set_error("optex_ot_step: workspace %zu < %zu bytes", ws_bytes, optex_ot_workspace_bytes(n_p, n_s, c, mode));
        return OPTEX_EWORKSPACE;
    }
    // One preparation launch per step: split R into its tf32 hi / lo halves ONCE for the three rotation GEMMs and
    // reset the cdf range slots; the GEMMs pick the halves up through the pre-split registry.
    struct PresplitGuard {
        ~PresplitGuard() { gemm_tc_set_presplit(nullptr, nullptr, nullptr); }
    } presplit_guard;
    bool forced_tc;
    const bool tc3 = want_tc(forced_tc) && tc_terms() == 3 && c % 4 == 0;
    prof_mark(st);
    if (tc3) {
        const bool fold = mode == OPTEX_MODE_CDF && !cdf_uses_channel_kernel(c, n_p, n_s, 256);
        if (g_r_split) {   // the caller split this rotation already (optex_split_rotations): no launch here
            if (fold) OPTEX_TRY(fill_u32((uint32_t *)mw, 2 * (int64_t)c, 0xffffffffu, st));
            gemm_tc_set_presplit(R, g_r_split, g_r_split + (size_t)c * c);
        } else {
            OPTEX_TRY(gemm_tc_split_and_fill(R, r_hi, r_lo, (int64_t)c * c, fold ? (uint32_t *)mw : nullptr,
                                             fold ? 2 * (int64_t)c : 0, 0xffffffffu, st));
            gemm_tc_set_presplit(R, r_hi, r_lo);
        }
    }
    uint32_t *minmax = nullptr;
    if (mode == OPTEX_MODE_CDF && !cdf_uses_channel_kernel(c, n_p, n_s, 256)) {
        // long rows (conv1_1 .. conv3_1): the forward rotations fold the per-channel range (histmatch.py:52-53) into
        // their epilogues, which saves the three-kernel matcher a pass over the rotated block; short rows: the
        // one-CTA-per-channel matcher finds the range itself
        minmax = (uint32_t *)mw;
        if (!tc3) OPTEX_TRY(fill_u32(minmax, 2 * (int64_t)c, 0xffffffffu, st));
    }
    prof_mark(st);  // stage 0: prepare (split R, reset range slots)
    bool r1 = false, r2 = false;
    // both forward rotations in ONE launch when the shapes take the paired tensor-core form (optex.py:170-171)
    int rc2 = OPTEX_ENOTSUP;
    static const char *no_dual = getenv("OPTEX_NO_DUAL_FORWARD");
    if (tc3 && !(no_dual && atoi(no_dual)))
        rc2 = gemm_tc_rotate_forward2(P, n_p, S, n_s, R, rp, rs, c, 3, st, minmax);
    if (rc2 == OPTEX_OK) {
        r1 = r2 = minmax != nullptr;
        prof_mark(st);  // stage 1: forward rotation of the pastiche and the style (one launch)
        prof_mark(st);  // stage 2: (empty)
    } else {
        if (rc2 != OPTEX_ENOTSUP) return rc2;
        OPTEX_TRY(rotate_forward(P, R, rp, n_p, c, true, st, 0, -1, minmax, &r1));
        prof_mark(st);  // stage 1: forward rotation of the pastiche
        OPTEX_TRY(rotate_forward(S, R, rs, n_s, c, true, st, 0, -1, minmax, &r2));
        prof_mark(st);  // stage 2: forward rotation of the style
    }
    if (mode == OPTEX_MODE_CDF)
        OPTEX_TRY(cdf_match_core(rp, rs, mt, c, n_p, n_s, 256, nullptr, mw, mws, r1 && r2, st));
    else
        OPTEX_TRY(sort_match_inplace(rp, rs, mt, c, n_p, n_s, nullptr, mw, mws, st));  // rs is scratch: sorted in place
    prof_mark(st);  // stage 3: matcher
    int rc = rotate_inverse(mt, true, R, out, n_p, c, content, strength, st);
    prof_mark(st);  // stage 4: inverse rotation (+ blend)
    return rc;
}

// library-owned device scratch for the *_host entry points
struct HostScratch {
    std::mutex mu;
    void *buf = nullptr;
    size_t cap = 0;
    int dev = -1;
    int ensure(size_t bytes) {
        int d = 0;
        OPTEX_CUDA(cudaGetDevice(&d));
        if (buf && (d != dev || cap < bytes)) {
            cudaFree(buf);
            buf = nullptr;
            cap = 0;
        }
        if (!buf) {
            OPTEX_CUDA(cudaMalloc(&buf, bytes));
            cap = bytes;
            dev = d;
        }
        return OPTEX_OK;
    }
};
constexpr int kHostSlots = 4;   // upload of step i+1 | compute of step i | download of step i-1 (+ one in reserve)
static HostScratch g_host[kHostSlots];  // one per pipeline slot of optex_ot_step_host_async

// The style block is constant over the iterations of a layer (optex.py:112-113 passes the same style_features[l]
// every time): optex_ot_host_set_style uploads it once, the *_host entry points called with S == NULL reuse it.
struct ResidentStyle {
    std::mutex mu;
    void *buf = nullptr;
    size_t cap = 0;
    int dev = -1, b_s = 0, c = 0;
    int64_t hw_s = 0;
    cudaEvent_t ready = nullptr;
};
static ResidentStyle g_style;

__global__ void fence_kernel() {}
// An ORDINARY launch (no programmatic serialisation): it runs only after every earlier kernel of the stream has fully
// completed, so an event recorded behind it - the fork / join of a side stream, a timing event - cannot be reached
// while a PDL-launched predecessor is still draining.
int stream_fence(cudaStream_t st) {
    fence_kernel<<<1, 32, 0, st>>>();
    OPTEX_LAUNCH_CHECK("fence_kernel");
    return OPTEX_OK;
}

}  // namespace optex

using namespace optex;

extern "C" int optex_abi_version(void) { return OPTEX_ABI_VERSION; }
extern "C" const char *optex_last_error(void) { return g_err; }
extern "C" int optex_device_check(void) { return require_sm100(); }
extern "C" uint64_t optex_launch_count(void) { return g_launches.load(); }
extern "C" int optex_set_gemm_mode(int m) {
    if (m < OPTEX_GEMM_AUTO || m > OPTEX_GEMM_TF32) {
        set_error("optex_set_gemm_mode: unknown mode %d", m);
        return OPTEX_EINVAL;
    }
    g_gemm_mode.store(m);
    return OPTEX_OK;
}
extern "C" int optex_get_gemm_mode(void) { return g_gemm_mode.load(); }
extern "C" int optex_debug_gemm_trace(void *device_buf) {
    gemm_tc_set_trace((unsigned long long *)device_buf);
    return OPTEX_OK;
}
static thread_local int g_user_slot = 0;
extern "C" int optex_set_scratch_slot(int slot) {
    const int prev = g_user_slot;
    g_user_slot = slot & 3;
    gemm_tc_set_scratch_slot(g_user_slot);
    return prev;
}

extern "C" int optex_set_pdl(int enable) {
    g_pdl.store(enable ? 1 : 0);
    return OPTEX_OK;
}

extern "C" size_t optex_ot_workspace_bytes(int64_t n_p, int64_t n_s, int c, int mode) {
    if (n_p < 1 || n_s < 1 || c < 1 || !valid_mode(mode)) return 0;
    if (!per_channel(mode)) return align_up(cov_match_ws_bytes(n_p, n_s, c, mode), 256);
    size_t b = align_up(sizeof(float) * (size_t)n_p * c, 256) + align_up(sizeof(float) * (size_t)n_s * c, 256);
    b += 2 * align_up(sizeof(float) * (size_t)c * c, 256);  // tf32 hi / lo halves of the rotation
    return b + align_up(match_ws_bytes(n_p, n_s, c, mode), 256);
}

extern "C" int optex_ot_step(const float *P, const float *S, const float *R, float *out, int b_p, int64_t hw_p,
                             int b_s, int64_t hw_s, int c, int mode, float eps, const float *content,
                             float content_strength, void *workspace, size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    OPTEX_TRY(check_step_args("optex_ot_step", P, S, out, b_p, hw_p, b_s, hw_s, c, mode));
    if (!R) {
        set_error("optex_ot_step: NULL rotation");
        return OPTEX_EINVAL;
    }
    if (out == P || out == S) {
        set_error("optex_ot_step: out must not alias an input (use optex_ot_loop for in-place iteration)");
        return OPTEX_EINVAL;
    }
    return ot_step_impl(P, S, R, out, b_p, hw_p, b_s, hw_s, c, mode, eps, content, content_strength, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}

static int ot_step_host_impl(const float *P, const float *S, const float *R, float *out, int b_p, int64_t hw_p,
                             int b_s, int64_t hw_s, int c, int mode, float eps, const float *content,
                             float content_strength, uint64_t seed, uint64_t counter, int slot, bool sync,
                             cudaStream_t st) {
    OPTEX_TRY(require_sm100());
    const float *dS_res = nullptr;
    if (!S) {  // resident style (optex_ot_host_set_style)
        std::lock_guard<std::mutex> lock(g_style.mu);
        int d = 0;
        OPTEX_CUDA(cudaGetDevice(&d));
        if (!g_style.buf || g_style.dev != d) {
            set_error("optex_ot_step_host: S == NULL but no style block is resident on device %d "
                      "(call optex_ot_host_set_style first)", d);
            return OPTEX_EINVAL;
        }
        if (g_style.b_s != b_s || g_style.hw_s != hw_s || g_style.c != c) {
            set_error("optex_ot_step_host: resident style is [%d x %lld, %d], call asks for [%d x %lld, %d]", g_style.b_s,
                      (long long)g_style.hw_s, g_style.c, b_s, (long long)hw_s, c);
            return OPTEX_EINVAL;
        }
        dS_res = (const float *)g_style.buf;
        OPTEX_CUDA(cudaStreamWaitEvent(st, g_style.ready, 0));
    }
    OPTEX_TRY(check_step_args("optex_ot_step_host", P, S ? (const void *)S : (const void *)dS_res, out, b_p, hw_p, b_s,
                              hw_s, c, mode));
    const int64_t n_p = (int64_t)b_p * hw_p, n_s = (int64_t)b_s * hw_s;
    const size_t bp = align_up(sizeof(float) * (size_t)n_p * c, 256), bs = align_up(sizeof(float) * (size_t)n_s * c, 256);
    const size_t br = align_up(sizeof(float) * (size_t)c * c, 256);
    const size_t step_ws = optex_ot_workspace_bytes(n_p, n_s, c, mode);
    const size_t rot_ws = R ? 0 : align_up(rotation_ws_bytes(c, 1), 256);
    const size_t ws = step_ws + rot_ws;
    HostScratch &hs = g_host[slot % kHostSlots];
    std::lock_guard<std::mutex> lock(hs.mu);
    OPTEX_TRY(hs.ensure(2 * bp + bs + br + (content ? bp : 0) + ws));
    char *base = (char *)hs.buf;
    float *dP = (float *)base, *dO = (float *)(base + bp), *dS = (float *)(base + 2 * bp);
    float *dR = (float *)(base + 2 * bp + bs);
    float *dC = content ? (float *)(base + 2 * bp + bs + br) : nullptr;
    void *dW = base + 2 * bp + bs + br + (content ? bp : 0);
    gemm_tc_set_scratch_slot(slot);
    OPTEX_CUDA(cudaMemcpyAsync(dP, P, sizeof(float) * n_p * c, cudaMemcpyHostToDevice, st));
    if (S) OPTEX_CUDA(cudaMemcpyAsync(dS, S, sizeof(float) * n_s * c, cudaMemcpyHostToDevice, st));
    const float *dS_use = S ? dS : dS_res;
    if (R)
        OPTEX_CUDA(cudaMemcpyAsync(dR, R, sizeof(float) * (size_t)c * c, cudaMemcpyHostToDevice, st));
    else
        OPTEX_TRY(random_rotations(dR, c, 1, seed, counter, nullptr, (char *)dW + step_ws, rot_ws, st));
    if (content) OPTEX_CUDA(cudaMemcpyAsync(dC, content, sizeof(float) * n_p * c, cudaMemcpyHostToDevice, st));
    int rc = ot_step_impl(dP, dS_use, dR, dO, b_p, hw_p, b_s, hw_s, c, mode, eps, dC, content_strength, dW, step_ws, st);
    gemm_tc_set_scratch_slot(g_user_slot);
    OPTEX_TRY(rc);
    OPTEX_CUDA(cudaMemcpyAsync(out, dO, sizeof(float) * n_p * c, cudaMemcpyDeviceToHost, st));
    if (sync) OPTEX_CUDA(cudaStreamSynchronize(st));
    return OPTEX_OK;
}

extern "C" int optex_ot_step_host(const float *P, const float *S, const float *R, float *out, int b_p,
                                  int64_t hw_p, int b_s, int64_t hw_s, int c, int mode, float eps,
                                  const float *content, float content_strength, uint64_t seed,
                                  uint64_t counter, void *stream) {
    return ot_step_host_impl(P, S, R, out, b_p, hw_p, b_s, hw_s, c, mode, eps, content, content_strength, seed,
                             counter, 0, true, (cudaStream_t)stream);
}

extern "C" int optex_ot_step_host_async(const float *P, const float *S, const float *R, float *out, int b_p,
                                        int64_t hw_p, int b_s, int64_t hw_s, int c, int mode, float eps,
                                        const float *content, float content_strength, uint64_t seed,
                                        uint64_t counter, int slot, void *stream) {
    if (slot < 0 || slot >= kHostSlots) {
        set_error("optex_ot_step_host_async: slot must be 0 .. 3");
        return OPTEX_EINVAL;
    }
    return ot_step_host_impl(P, S, R, out, b_p, hw_p, b_s, hw_s, c, mode, eps, content, content_strength, seed,
                             counter, slot, false, (cudaStream_t)stream);
}

extern "C" int optex_ot_host_set_style(const float *S, int b_s, int64_t hw_s, int c, void *stream) {
    OPTEX_TRY(require_sm100());
    std::lock_guard<std::mutex> lock(g_style.mu);
    if (!S) {  // release
        if (g_style.buf) {
            cudaDeviceSynchronize();
            cudaFree(g_style.buf);
        }
        g_style.buf = nullptr;
        g_style.cap = 0;
        g_style.dev = -1;
        return OPTEX_OK;
    }
    if (b_s < 1 || hw_s < 1 || c < 1) {
        set_error("optex_ot_host_set_style: empty shape");
        return OPTEX_EINVAL;
    }
    int d = 0;
    OPTEX_CUDA(cudaGetDevice(&d));
    const size_t bytes = sizeof(float) * (size_t)b_s * hw_s * c;
    if (g_style.buf && (g_style.dev != d || g_style.cap < bytes)) {
        OPTEX_CUDA(cudaDeviceSynchronize());
        OPTEX_CUDA(cudaFree(g_style.buf));
        g_style.buf = nullptr;
        g_style.cap = 0;
    }
    if (!g_style.buf) {
        OPTEX_CUDA(cudaMalloc(&g_style.buf, bytes));
        g_style.cap = bytes;
        g_style.dev = d;
    }
    if (!g_style.ready) OPTEX_CUDA(cudaEventCreateWithFlags(&g_style.ready, cudaEventDisableTiming));
    cudaStream_t st = (cudaStream_t)stream;
    OPTEX_CUDA(cudaMemcpyAsync(g_style.buf, S, bytes, cudaMemcpyHostToDevice, st));
    OPTEX_CUDA(cudaEventRecord(g_style.ready, st));
    g_style.b_s = b_s;
    g_style.hw_s = hw_s;
    g_style.c = c;
    return OPTEX_OK;
}

static const int kStepsChunk = 32;

extern "C" size_t optex_ot_steps_workspace_bytes(int64_t n_p, int64_t n_s, int c, int mode) {
    const size_t step = optex_ot_workspace_bytes(n_p, n_s, c, mode);
    if (!step) return 0;
    // + a chunk of rotations (R and its tf32 hi / lo planes) and the draw's own scratch
    return step + align_up(sizeof(float) * 3 * (size_t)kStepsChunk * c * c, 256) +
           align_up(rotation_ws_bytes(c, kStepsChunk), 256);
}

// `steps` INDEPENDENT OT steps enqueued by one call (a batch of syntheses at the same layer; the granularity of the
// reference's own loop, optex.py:112-113, without a host round trip per step): step i transports
// P[(first + i) % n_sets] towards S[(first + i) % n_sets] with rotation R_all[i] into out[(first + i) % n_out].
// R_all == NULL: the rotations are drawn on the device from (seed, first_counter + i) like the reference draws one per
// call (optex.py:168) - in batches of up to 32 on the caller's stream (workspace: optex_ot_steps_workspace_bytes).
// (Drawing the next batch on a side stream beside the running steps was measured and buys nothing: the device is busy,
// the draw's FP64 work only competes - 210.4 vs 210.7 us per step at 200 steps, worse at 20.)
extern "C" int optex_ot_steps(const float *const *P, const float *const *S, int n_sets, const float *R_all,
                              const float *R_split, uint64_t seed, uint64_t first_counter, float *const *out,
                              int n_out, int steps, int first, int b_p, int64_t hw_p, int b_s, int64_t hw_s, int c,
                              int mode, float eps, void *workspace, size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    if (!P || !S || !out || n_sets < 1 || n_out < 1 || steps < 0 || first < 0) {
        set_error("optex_ot_steps: NULL table or empty set list");
        return OPTEX_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_p = (int64_t)b_p * hw_p, n_s = (int64_t)b_s * hw_s;
    const size_t cc = (size_t)c * c;
    auto run = [&](int i, const float *R, const float *split) -> int {
        const int k = (first + i) % n_sets, o = (first + i) % n_out;
        OPTEX_TRY(check_step_args("optex_ot_steps", P[k], S[k], out[o], b_p, hw_p, b_s, hw_s, c, mode));
        if (out[o] == P[k] || out[o] == S[k]) {
            set_error("optex_ot_steps: out must not alias an input");
            return OPTEX_EINVAL;
        }
        g_r_split = split;
        int rc = ot_step_impl(P[k], S[k], R, out[o], b_p, hw_p, b_s, hw_s, c, mode, eps, nullptr, 0.f, workspace,
                              optex_ot_workspace_bytes(n_p, n_s, c, mode), st);
        g_r_split = nullptr;
        return rc;
    };
    if (R_all) {
        if (workspace_bytes < optex_ot_workspace_bytes(n_p, n_s, c, mode)) {
            set_error("optex_ot_steps: workspace too small");
            return OPTEX_EWORKSPACE;
        }
        for (int i = 0; i < steps; ++i)
            OPTEX_TRY(run(i, R_all + (size_t)i * cc, R_split ? R_split + (size_t)i * 2 * cc : nullptr));
        return OPTEX_OK;
    }
    // ---- rotations drawn here, a batch at a time
    const size_t step_ws = optex_ot_workspace_bytes(n_p, n_s, c, mode);
    if (!workspace || workspace_bytes < optex_ot_steps_workspace_bytes(n_p, n_s, c, mode)) {
        set_error("optex_ot_steps: workspace %zu < %zu bytes (optex_ot_steps_workspace_bytes)", workspace_bytes,
                  optex_ot_steps_workspace_bytes(n_p, n_s, c, mode));
        return OPTEX_EWORKSPACE;
    }
    Arena ar((char *)workspace + step_ws, workspace_bytes - step_ws);
    float *R = ar.take<float>(3 * (size_t)kStepsChunk * cc);
    const size_t rws_bytes = rotation_ws_bytes(c, kStepsChunk);
    void *rws = ar.take<char>(rws_bytes);
    const bool split = c % 4 == 0;
    for (int i0 = 0; i0 < steps; i0 += kStepsChunk) {
        const int nb = steps - i0 < kStepsChunk ? steps - i0 : kStepsChunk;
        OPTEX_TRY(random_rotations(R, c, nb, seed, first_counter + (uint64_t)i0, nullptr, rws, rws_bytes, st));
        if (split) OPTEX_TRY(gemm_tc_split_batch(R, R + (size_t)kStepsChunk * cc, nb, (int64_t)cc, st));
        for (int j = 0; j < nb; ++j)
            OPTEX_TRY(run(i0 + j, R + (size_t)j * cc, split ? R + (size_t)kStepsChunk * cc + (size_t)j * 2 * cc : nullptr));
    }
    return OPTEX_OK;
}

// tf32 hi / lo planes of `count` rotations in ONE launch: out[i] = [hi(R_i) | lo(R_i)], each [c, c] (the 3xTF32
// operand halves of the rotation GEMMs; c % 4 == 0).  Batched like the draw itself (optex_random_rotations).
extern "C" int optex_split_rotations(const float *R_all, int count, int c, float *out, void *stream) {
    OPTEX_TRY(require_sm100());
    if (!R_all || !out || count < 1 || c < 1 || c % 4 != 0) {
        set_error("optex_split_rotations: NULL pointer, empty batch or c %% 4 != 0");
        return OPTEX_EINVAL;
    }
    return gemm_tc_split_batch(R_all, out, count, (int64_t)c * c, (cudaStream_t)stream);
}

// One OT step with an event between its stages (measurement aid for bench.py: the in-step launches exactly as
// optex_ot_step issues them, which the stand-alone building blocks only approximate).  PDL is off for the call (an
// event between two programmatically serialised kernels does not separate them); synchronises the stream.
// per-channel modes: stage_ms[0..4] = prepare, rotate_forward_P, rotate_forward_S, match, rotate_inverse;
// covariance modes: stage_ms[0] = the whole step.
extern "C" int optex_ot_step_profile(const float *P, const float *S, const float *R, float *out, int b_p, int64_t hw_p,
                                     int b_s, int64_t hw_s, int c, int mode, float eps, void *workspace,
                                     size_t workspace_bytes, void *stream, float *stage_ms, int *stage_launches,
                                     int *n_stages) {
    OPTEX_TRY(require_sm100());
    OPTEX_TRY(check_step_args("optex_ot_step_profile", P, S, out, b_p, hw_p, b_s, hw_s, c, mode));
    if (!R || !stage_ms || !stage_launches || !n_stages) {
        set_error("optex_ot_step_profile: NULL rotation or result pointer");
        return OPTEX_EINVAL;
    }
    StageProf prof;
    for (auto &e : prof.ev) OPTEX_CUDA(cudaEventCreate(&e));
    const int pdl = g_pdl.exchange(0);
    g_prof = &prof;
    int rc = ot_step_impl(P, S, R, out, b_p, hw_p, b_s, hw_s, c, mode, eps, nullptr, 0.f, workspace, workspace_bytes,
                          (cudaStream_t)stream);
    g_prof = nullptr;
    g_pdl.store(pdl);
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    *n_stages = 0;
    if (rc == OPTEX_OK && e == cudaSuccess) {
        for (int i = 0; i + 1 < prof.n; ++i) {
            cudaEventElapsedTime(&stage_ms[i], prof.ev[i], prof.ev[i + 1]);
            stage_launches[i] = (int)(prof.launches[i + 1] - prof.launches[i]);
        }
        *n_stages = prof.n > 0 ? prof.n - 1 : 0;
    }
    for (auto &ev : prof.ev) cudaEventDestroy(ev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    return rc;
}

// An ordinary (not programmatically serialised) empty kernel: it starts only after every earlier kernel of the
// stream - including PDL-launched ones, whose completion an event alone does not imply - has fully drained.
extern "C" int optex_fence(void *stream) {
    OPTEX_TRY(require_sm100());
    return stream_fence((cudaStream_t)stream);
}

// The step on this rank's rows (include/optex_b200.h "multi-GPU"; sharded.cu has the NCCL plumbing).
extern "C" int optex_ot_step_sharded(optex_comm_t *comm, const float *P, const float *S, const float *R, float *out,
                                     int64_t n_p, int64_t n_s, int64_t n_p_total, int64_t n_s_total, int c, int mode,
                                     float eps, const float *content, float strength, void *ws, size_t ws_bytes,
                                     void *stream) {
    OPTEX_TRY(require_sm100());
    const ShardComm *cm = (const ShardComm *)comm;
    if (!cm) {
        set_error("optex_ot_step_sharded: NULL communicator");
        return OPTEX_EINVAL;
    }
    OPTEX_TRY(check_step_args("optex_ot_step_sharded", P, S, out, 1, n_p, 1, n_s, c, mode));
    if (!R && (per_channel(mode) || mode == OPTEX_MODE_CHOL)) {
        set_error("optex_ot_step_sharded: NULL rotation");
        return OPTEX_EINVAL;
    }
    if (n_p_total < n_p || n_s_total < n_s) {
        set_error("optex_ot_step_sharded: totals smaller than the local row counts");
        return OPTEX_EINVAL;
    }
    if (out == P || out == S) {
        set_error("optex_ot_step_sharded: out must not alias an input");
        return OPTEX_EINVAL;
    }
    if (mode == OPTEX_MODE_SORT) {
        set_error("optex_ot_step_sharded: sort needs global ranks per channel - use the channel-sharded form "
                  "(optimaltextures_b200.parallel.optimal_transport_sharded)");
        return OPTEX_EUNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (mode != OPTEX_MODE_CDF) {
        // pixel-sharded moments: cov_match.cu all-reduces the column sums and the centred Gram while the context is set
        ShardCtx ctx{cm, n_p_total, n_s_total};
        cov_set_shard(&ctx);
        int rc = cov_ot_step(P, S, R, out, 1, n_p, 1, n_s, c, mode, eps, content, strength, ws, ws_bytes, st, 0);
        cov_set_shard(nullptr);
        return rc;
    }
    if (n_p_total >= (1LL << 31) || n_s_total >= (1LL << 31)) {
        set_error("optex_ot_step_sharded: n >= 2^31 per channel");
        return OPTEX_ESIZE;
    }
    Arena ar(ws, ws_bytes);
    float *rp = ar.take<float>((size_t)n_p * c);
    float *rs = ar.take<float>((size_t)n_s * c);
    float *r_hi = ar.take<float>((size_t)c * c), *r_lo = ar.take<float>((size_t)c * c);
    const size_t mws = match_ws_bytes(n_p, n_s, c, mode);
    Arena mar(ar.take<char>(mws), mws);
    uint32_t *minmax = mar.take<uint32_t>(2 * (size_t)c);
    uint32_t *hist = mar.take<uint32_t>(2 * (size_t)c * 256);
    float *tbl = mar.take<float>(4 * (size_t)c * 256 + 4);
    if (!ar.ok() || !mar.ok()) {
        set_error("optex_ot_step_sharded: workspace %zu < %zu bytes", ws_bytes, optex_ot_workspace_bytes(n_p, n_s, c, mode));
        return OPTEX_EWORKSPACE;
    }
    struct PresplitGuard {
        ~PresplitGuard() { gemm_tc_set_presplit(nullptr, nullptr, nullptr); }
    } presplit_guard;
    bool forced_tc;
    const bool tc3 = want_tc(forced_tc) && tc_terms() == 3 && c % 4 == 0;
    if (tc3) {
        OPTEX_TRY(gemm_tc_split_and_fill(R, r_hi, r_lo, (int64_t)c * c, minmax, 2 * (int64_t)c, 0xffffffffu, st));
        gemm_tc_set_presplit(R, r_hi, r_lo);
    } else {
        OPTEX_TRY(fill_u32(minmax, 2 * (int64_t)c, 0xffffffffu, st));
    }
    bool r1 = false, r2 = false;
    OPTEX_TRY(rotate_forward(P, R, rp, n_p, c, true, st, 0, -1, minmax, &r1));     // local rows, range folded
    OPTEX_TRY(rotate_forward(S, R, rs, n_s, c, true, st, 0, -1, minmax, &r2));
    if (!(r1 && r2)) OPTEX_TRY(cdf_stage_range(rp, rs, c, n_p, n_s, minmax, true, st));
    OPTEX_TRY(shard_allreduce_u32(cm, minmax, 2 * (size_t)c, true, st));            // histmatch.py:52-53 over all rows
    OPTEX_TRY(cdf_stage_hist(rp, rs, c, n_p, n_s, 256, minmax, hist, st));
    OPTEX_TRY(shard_allreduce_u32(cm, hist, 2 * (size_t)c * 256, false, st));       // histmatch.py:57-58 over all rows
    OPTEX_TRY(cdf_stage_apply(rp, rp, c, n_p, 256, minmax, hist, tbl, st));
    return rotate_inverse(rp, true, R, out, n_p, c, content, strength, st);
}

static const int kRotChunk = 16;

extern "C" size_t optex_ot_loop_workspace_bytes(int64_t n_p, int64_t n_s, int c, int mode) {
    size_t step = optex_ot_workspace_bytes(n_p, n_s, c, mode);
    if (!step) return 0;
    // a second feature block: ping-pong partner (closed-form modes) / second channel-major buffer (fused loop)
    size_t alt = align_up(sizeof(float) * (size_t)n_p * c, 256);
    return step + alt + align_up(sizeof(float) * (size_t)(kRotChunk + 2) * c * c, 256) +
           align_up(rotation_ws_bytes(c, kRotChunk), 256);
}

// Per-channel modes without a content blend: the un-rotated pastiche between two iterations is never needed, so
// "rotate back with R_i, rotate forward with R_{i+1}" (optex.py:175 then :170) becomes ONE rotation of the
// channel-major block by Q = R_i^T R_{i+1} - two N x C x C GEMMs per iteration instead of three.  Same maths as
// the step-by-step loop up to fp32 rounding (tests/test_gpu_ot_step.py::test_ot_loop_fused_rotations).
static int tc_must(int rc) {
    if (rc != OPTEX_ENOTSUP) return rc;
    set_error("optex_ot_loop: internal - fused rotation rejected by the tensor-core path");
    return OPTEX_ESIZE;
}
static int ot_loop_fused(float *feat, const float *S, const float *R_all, int iters, uint64_t seed,
                         uint64_t first_counter, int64_t n_p, int64_t n_s, int c, int mode, void *sw, size_t step_ws,
                         float *alt, float *rbuf, void *rws, size_t rws_bytes, cudaStream_t st) {
    Arena ar(sw, step_ws);
    float *rp = ar.take<float>((size_t)n_p * c);
    float *rs = ar.take<float>((size_t)n_s * c);
    size_t mws = match_ws_bytes(n_p, n_s, c, mode);
    void *mw = ar.take<char>(mws);
    uint32_t *minmax = (uint32_t *)mw;
    float *Q = rbuf + (size_t)(kRotChunk + 1) * c * c;
    const bool cdf = mode == OPTEX_MODE_CDF && !cdf_uses_channel_kernel(c, n_p, n_s, 256);   // fold the range in the GEMMs
    const bool is_cdf = mode == OPTEX_MODE_CDF;
    const int terms = tc_terms();
    float *cur = rp, *nxt = alt;
    const float *R = nullptr;
    for (int i = 0; i < iters; ++i) {
        // R_i lives in slot 1 + (i % chunk); slot 0 keeps the last matrix of the previous chunk
        const float *Rprev = R;
        if (R_all) {
            R = R_all + (size_t)i * c * c;
        } else {
            if (i % kRotChunk == 0) {
                if (i > 0) {
                    OPTEX_CUDA(cudaMemcpyAsync(rbuf, Rprev, sizeof(float) * (size_t)c * c, cudaMemcpyDeviceToDevice, st));
                    Rprev = rbuf;
                }
                int nb = iters - i < kRotChunk ? iters - i : kRotChunk;
                OPTEX_TRY(random_rotations(rbuf + (size_t)c * c, c, nb, seed, first_counter + (uint64_t)i, nullptr, rws,
                                           rws_bytes, st));
            }
            R = rbuf + (size_t)(1 + i % kRotChunk) * c * c;
        }
        if (cdf) OPTEX_TRY(fill_u32(minmax, 2 * (int64_t)c, 0xffffffffu, st));
        bool r1 = false, r2 = false;
        if (i == 0) {
            OPTEX_TRY(rotate_forward(feat, R, cur, n_p, c, true, st, 0, -1, cdf ? minmax : nullptr, &r1));
        } else {
            TcGemm q{};  // Q = R_{i-1}^T R_i
            q.A = Rprev; q.a_mn = true; q.B = R; q.b_mn = true; q.D = Q; q.ldd = c; q.M = q.N = q.K = c;
            q.terms = terms; q.alpha = 1.f;
            OPTEX_TRY(tc_must(gemm_tc(q, st)));
            TcGemm g{};  // next[c_out, n] = sum_k Q[k, c_out] cur[k, n]
            g.A = Q; g.a_mn = true; g.B = cur; g.b_mn = true; g.D = nxt; g.ldd = n_p; g.M = c; g.N = n_p; g.K = c;
            g.terms = terms; g.alpha = 1.f; g.rowrange = cdf ? minmax : nullptr; g.presplit_b = true;
            OPTEX_TRY(tc_must(gemm_tc(g, st)));
            float *t = cur; cur = nxt; nxt = t;
            r1 = true;
        }
        OPTEX_TRY(rotate_forward(S, R, rs, n_s, c, true, st, 0, -1, cdf ? minmax : nullptr, &r2));
        if (is_cdf)
            OPTEX_TRY(cdf_match_core(cur, rs, cur, c, n_p, n_s, 256, nullptr, mw, mws, cdf && r1 && r2, st));
        else
            OPTEX_TRY(sort_match_inplace(cur, rs, cur, c, n_p, n_s, nullptr, mw, mws, st));
    }
    return rotate_inverse(cur, true, R, feat, n_p, c, nullptr, 0.f, st);
}

extern "C" int optex_ot_loop(float *feat, const float *S, const float *R_all, int iters, uint64_t seed,
                             uint64_t first_counter, int b_p, int64_t hw_p, int b_s, int64_t hw_s, int c,
                             int mode, float eps, const float *content, float content_strength, void *workspace,
                             size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    OPTEX_TRY(check_step_args("optex_ot_loop", feat, S, feat, b_p, hw_p, b_s, hw_s, c, mode));
    if (iters < 0) {
        set_error("optex_ot_loop: iters < 0");
        return OPTEX_EINVAL;
    }
    const int64_t n_p = (int64_t)b_p * hw_p, n_s = (int64_t)b_s * hw_s;
    const size_t step_ws = optex_ot_workspace_bytes(n_p, n_s, c, mode);
    Arena ar(workspace, workspace_bytes);
    void *sw = ar.take<char>(step_ws);
    float *alt = ar.take<float>((size_t)n_p * c);
    float *rbuf = ar.take<float>((size_t)(kRotChunk + 2) * c * c);
    size_t rws_bytes = rotation_ws_bytes(c, kRotChunk);
    void *rws = ar.take<char>(rws_bytes);
    if (!ar.ok()) {
        set_error("optex_ot_loop: workspace %zu < %zu bytes", workspace_bytes,
                  optex_ot_loop_workspace_bytes(n_p, n_s, c, mode));
        return OPTEX_EWORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    {
        bool forced;
        static const char *no_fuse = getenv("OPTEX_NO_LOOP_FUSION");
        if (!no_fuse && per_channel(mode) && !content && iters >= 2 && want_tc(forced) && c % 32 == 0 && n_p % 32 == 0 &&
            (mode != OPTEX_MODE_SORT || sort_match_scratch_bytes(c, n_p, n_s) == 0))
            return ot_loop_fused(feat, S, R_all, iters, seed, first_counter, n_p, n_s, c, mode, sw, step_ws, alt, rbuf,
                                 rws, rws_bytes, st);
    }
    if (per_channel(mode)) alt = nullptr;
    // pca / sym do not depend on the rotation (cov_match.cu): no draw at all
    const bool needs_rotation = per_channel(mode) || mode == OPTEX_MODE_CHOL;
    for (int i = 0; i < iters; ++i) {
        const float *R = nullptr;
        if (R_all) {
            R = R_all + (size_t)i * c * c;
        } else if (needs_rotation) {
            if (i % kRotChunk == 0) {
                int nb = iters - i < kRotChunk ? iters - i : kRotChunk;
                OPTEX_TRY(random_rotations(rbuf, c, nb, seed, first_counter + (uint64_t)i, nullptr, rws, rws_bytes, st));
            }
            R = rbuf + (size_t)(i % kRotChunk) * c * c;
        }
        // per-channel modes run in place (P is dead after the forward rotation); the closed-form modes read
        // P in their last GEMM, so they ping-pong between feat and alt
        float *src = (alt && (i & 1)) ? alt : feat;
        float *dst = alt ? ((i & 1) ? feat : alt) : feat;
        // the style side of the closed-form modes (moments, pca square root) is computed by the first iteration only
        // closed-form modes: 4 = a loop follows (first iteration), 3 = style side in the workspace AND the pastiche is
        // the previous iteration's output, whose mean is known analytically (cov_match.cu); 8 = its covariance may be
        // propagated too (cov_small.cu; every 8th iteration measures it again)
        OPTEX_TRY(ot_step_impl(src, S, R, dst, b_p, hw_p, b_s, hw_s, c, mode, eps, content, content_strength, sw,
                               step_ws, st, i > 0 ? (3 | (i % 8 ? 8 : 0)) : 4));
    }
    if (alt && (iters & 1))
        OPTEX_CUDA(cudaMemcpyAsync(feat, alt, sizeof(float) * (size_t)n_p * c, cudaMemcpyDeviceToDevice, st));
    return OPTEX_OK;
}

extern "C" size_t optex_hist_match_workspace_bytes(int64_t n_t, int64_t n_s, int c, int mode) {
    if (n_t < 1 || n_s < 1 || c < 1 || !valid_mode(mode)) return 0;
    if (per_channel(mode))
        return align_up(sizeof(float) * (size_t)n_t * c, 256) + align_up(sizeof(float) * (size_t)n_s * c, 256) +
               align_up(match_ws_bytes(n_t, n_s, c, mode), 256);
    return align_up(cov_match_ws_bytes(n_t, n_s, c, mode), 256);
}

extern "C" int optex_hist_match(const float *target, const float *source, float *out, int b_t, int64_t hw_t,
                                int b_s, int64_t hw_s, int c, int mode, float eps, void *workspace,
                                size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    OPTEX_TRY(check_step_args("optex_hist_match", target, source, out, b_t, hw_t, b_s, hw_s, c, mode));
    const int64_t n_t = (int64_t)b_t * hw_t, n_s = (int64_t)b_s * hw_s;
    cudaStream_t st = (cudaStream_t)stream;
    if (!per_channel(mode))
        return cov_match_nhwc(target, source, out, b_t, hw_t, b_s, hw_s, c, mode, eps, workspace, workspace_bytes, st);
    // per-channel modes work channel-major (histmatch.py:6-8 permutes to [c, b, h, w])
    Arena ar(workspace, workspace_bytes);
    float *tt = ar.take<float>((size_t)n_t * c);
    float *ts = ar.take<float>((size_t)n_s * c);
    size_t mws = match_ws_bytes(n_t, n_s, c, mode);
    void *mw = ar.take<char>(mws);
    if (!ar.ok()) {
        set_error("optex_hist_match: workspace %zu < %zu bytes", workspace_bytes,
                  optex_hist_match_workspace_bytes(n_t, n_s, c, mode));
        return OPTEX_EWORKSPACE;
    }
    OPTEX_TRY(transpose_f32(target, tt, n_t, c, st));
    OPTEX_TRY(transpose_f32(source, ts, n_s, c, st));
    if (mode == OPTEX_MODE_CDF)
        OPTEX_TRY(optex_cdf_match(tt, ts, tt, c, n_t, n_s, 256, nullptr, mw, mws, st));
    else
        OPTEX_TRY(sort_match_inplace(tt, ts, tt, c, n_t, n_s, nullptr, mw, mws, st));
    return transpose_f32(tt, out, c, n_t, st);
}

extern "C" int optex_rotate_forward(const float *X, const float *R, float *Xt, int64_t n, int c, void *stream) {
    OPTEX_TRY(require_sm100());
    if (!X || !R || !Xt || n < 1 || c < 1) {
        set_error("optex_rotate_forward: NULL pointer or empty shape");
        return OPTEX_EINVAL;
    }
    return rotate_forward(X, R, Xt, n, c, true, (cudaStream_t)stream);
}

// Split R into its tf32 hi / lo halves once (3xTF32 arithmetic) for the optex_rotate_* calls that follow on this
// host thread with the same R pointer; optex_ot_step does the same internally for its three GEMMs.
extern "C" size_t optex_rotation_prepare_workspace_bytes(int c) {
    return c > 0 ? align_up(sizeof(float) * 2 * (size_t)c * c, 256) : 0;
}

extern "C" int optex_rotation_prepare(const float *R, int c, void *workspace, size_t workspace_bytes, void *stream) {
    OPTEX_TRY(require_sm100());
    gemm_tc_set_presplit(nullptr, nullptr, nullptr);
    if (!R) return OPTEX_OK;  // release only
    if (c < 1 || c % 4 != 0) {
        set_error("optex_rotation_prepare: c must be a positive multiple of 4 (the tensor-core path's constraint)");
        return OPTEX_EINVAL;
    }
    if (!workspace || workspace_bytes < optex_rotation_prepare_workspace_bytes(c)) {
        set_error("optex_rotation_prepare: workspace %zu < %zu bytes", workspace_bytes,
                  optex_rotation_prepare_workspace_bytes(c));
        return OPTEX_EWORKSPACE;
    }
    bool forced_tc;
    if (!(want_tc(forced_tc) && tc_terms() == 3)) return OPTEX_OK;  // fp32 / single-pass tf32: nothing to split
    float *hi = (float *)workspace, *lo = hi + (size_t)c * c;
    OPTEX_TRY(gemm_tc_split_and_fill(R, hi, lo, (int64_t)c * c, nullptr, 0, 0u, (cudaStream_t)stream));
    gemm_tc_set_presplit(R, hi, lo);
    return OPTEX_OK;
}

extern "C" int optex_rotate_forward_block(const float *X, const float *R, float *Xt, int64_t n, int c, int c0, int nc,
                                          void *stream) {
    OPTEX_TRY(require_sm100());
    if (!X || !R || !Xt || n < 1 || c < 1 || c0 < 0 || nc < 1 || c0 + nc > c) {
        set_error("optex_rotate_forward_block: NULL pointer, empty shape or channel block outside [0, c)");
        return OPTEX_EINVAL;
    }
    return rotate_forward(X, R, Xt, n, c, true, (cudaStream_t)stream, c0, nc);
}

extern "C" int optex_rotate_inverse(const float *Mt, const float *R, float *out, int64_t n, int c,
                                    const float *content, float content_strength, void *stream) {
    OPTEX_TRY(require_sm100());
    if (!Mt || !R || !out || n < 1 || c < 1) {
        set_error("optex_rotate_inverse: NULL pointer or empty shape");
        return OPTEX_EINVAL;
    }
    return rotate_inverse(Mt, true, R, out, n, c, content, content_strength, (cudaStream_t)stream);
}
