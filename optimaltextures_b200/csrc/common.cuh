// Shared helpers for the optex_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/optex_b200.h"

namespace optex {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);  // records + returns OPTEX_ECUDA
void count_launch(int n = 1);
int require_sm100();  // OPTEX_OK or OPTEX_EDEVICE
int sm_count();

#define OPTEX_CUDA(call)                                           \
    do {                                                           \
        cudaError_t _e = (call);                                   \
        if (_e != cudaSuccess) return optex::cuda_fail(_e, #call); \
    } while (0)

#define OPTEX_LAUNCH_CHECK(name)                                   \
    do {                                                           \
        optex::count_launch();                                     \
        cudaError_t _e = cudaGetLastError();                       \
        if (_e != cudaSuccess) return optex::cuda_fail(_e, name);  \
    } while (0)

#define OPTEX_TRY(expr)                   \
    do {                                  \
        int _rc = (expr);                 \
        if (_rc != OPTEX_OK) return _rc;  \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: set it once per DEVICE and
// kernel, not once per process (one process may drive several GPUs).  `done` is a per-call-site flag array.
struct PerDeviceOnce {
    unsigned char done[64] = {};
};
template <typename K>
inline int ensure_dyn_smem(PerDeviceOnce &once, K kern, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= 64) dev = 63;
    if (once.done[dev] && dev != 63) return OPTEX_OK;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    once.done[dev] = 1;
    return OPTEX_OK;
}

// Bump allocator over the caller's workspace.
struct Arena {
    char *base;
    size_t cap, off;
    Arena(void *p, size_t n) : base((char *)p), cap(n), off(0) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        T *r = (T *)(base + off);
        off += bytes;
        return r;
    }
    bool ok() const { return base != nullptr && off <= cap; }
};

// ---- programmatic dependent launch (PDL) --------------------------------------
// Every kernel of the step waits on griddepcontrol.wait before it touches memory and is launched with
// programmaticStreamSerializationAllowed: the next kernel's launch is processed while its predecessor still runs
// (a dependent-launch boundary otherwise costs ~5 us of idle GPU), and the wait returns once the predecessor's
// memory is visible.  Kernels launched this way MUST call pdl_wait() first.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();  // optex_set_pdl(); defined in api.cu
int stream_fence(cudaStream_t st);  // ordinary empty kernel: full completion of everything before it (api.cu)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
    if (!pdl_enabled()) {
        kern<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
        return cudaSuccess;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---- device helpers ---------------------------------------------------------
// Order-preserving float <-> uint32 (for atomicMin/Max and radix/bitonic keys).
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- multi-GPU (sharded.cu): an NCCL communicator handle + the collectives of the sharded steps
struct ShardComm {
    void *comm;   // ncclComm_t
    int rank, world;
    int owned;    // created by optex_comm_init (destroyed with it) vs adopted from the caller
};
int shard_allreduce_u32(const ShardComm *cm, uint32_t *buf, size_t count, bool take_min, cudaStream_t st);
int shard_allreduce_f32_sum(const ShardComm *cm, float *buf, size_t count, cudaStream_t st);
int shard_allgather_f32(const ShardComm *cm, const float *send, float *recv, size_t count_per_rank, cudaStream_t st);
// pixel-sharded covariance step: while set (thread-local), moments() all-reduces its sums over `comm` and divides by
// the TOTAL pixel counts
struct ShardCtx {
    const ShardComm *comm;
    int64_t hw_p_total, hw_s_total;
};
void cov_set_shard(const ShardCtx *ctx);   // cov_match.cu

// ---- internal entry points shared between translation units ------------------
// GEMM:  D[m, n] = sum_k A(m,k) * B(k,n)
//   a_kmajor : A(m,k) = A[m*lda + k]  else A[k*lda + m]
//   b_kmajor : B(k,n) = B[n*ldb + k]  else B[k*ldb + n]
//   d_trans  : D(m,n) -> D[n*ldd + m] else D[m*ldd + n]
//   blend    : if non-null (same indexing as D, not transposed):
//              d += strength * (blend - d)          (optex.py:117)
int sgemm_simt(const float *A, int64_t lda, bool a_kmajor, const float *B, int64_t ldb,
               bool b_kmajor, float *D, int64_t ldd, bool d_trans, int64_t M, int64_t N,
               int64_t K, const float *blend, float strength, float alpha, cudaStream_t st);

struct SimtOpts {
    const float *bias = nullptr;  // + bias[(m / bias_hw) * bias_ld + n]
    int64_t bias_hw = 1, bias_ld = 0;
    bool accumulate = false;      // D += alpha * acc  instead of  D = alpha * acc
    int split_k = 1;              // blockIdx.z splits K; partials at D + z * d_z_stride
    int64_t d_z_stride = 0;
    const int *skip = nullptr;    // device flag: non-zero -> no-op
    const float *skip_below = nullptr;  // device value: *skip_below < skip_tol -> no-op
    float skip_tol = 0.f;
    bool relu = false;            // max(., 0) after bias / accumulate (conv layers)
};
int sgemm_simt_ex(const float *A, int64_t lda, bool a_kmajor, const float *B, int64_t ldb, bool b_kmajor, float *D,
                  int64_t ldd, bool d_trans, int64_t M, int64_t N, int64_t K, const float *blend, float strength,
                  float alpha, const SimtOpts &o, cudaStream_t st);

int transpose_f32(const float *in, float *out, int64_t rows, int64_t cols, cudaStream_t st);

// rotation.cu: a batch of Haar SO(c) matrices R[batch][c][c] from (seed, first_counter + b)
size_t rotation_ws_bytes(int c, int batch);
int random_rotations(float *R, int c, int batch, uint64_t seed, uint64_t first_counter, const double *gauss,
                     void *workspace, size_t workspace_bytes, cudaStream_t st);

// cdf_match.cu: workspace = [minmax: 2c u32][hist][tables]; have_range: the forward GEMMs already folded the
// per-channel range into minmax (which the caller initialised to 0xFF bytes)
size_t cdf_minmax_bytes(int c);
bool cdf_uses_channel_kernel(int c, int64_t n_t, int64_t n_s, int bins);  // that kernel computes the range itself
int fill_u32(uint32_t *p, int64_t n, uint32_t v, cudaStream_t st);
int cdf_match_core(const float *target, const float *source, float *out, int c, int64_t n_t, int64_t n_s, int bins,
                   float *tables, void *workspace, size_t workspace_bytes, bool have_range, cudaStream_t st);

// the same matcher stage by stage (pixel-sharded multi-GPU step: all-reduce minmax after range, hist after hist)
int cdf_stage_range(const float *target, const float *source, int c, int64_t n_t, int64_t n_s, uint32_t *minmax,
                    bool reset, cudaStream_t st);
int cdf_stage_hist(const float *target, const float *source, int c, int64_t n_t, int64_t n_s, int bins,
                   const uint32_t *minmax, uint32_t *hist, cudaStream_t st);
int cdf_stage_apply(const float *target, float *out, int c, int64_t n_t, int bins, const uint32_t *minmax,
                    const uint32_t *hist, float *tbl, cudaStream_t st);

// sort_match.cu: exact 1-D OT per channel; `source_scratch` [c, n_s] is sorted in place
size_t sort_match_scratch_bytes(int c, int64_t n_t, int64_t n_s);  // 0 while channels fit on chip (<= 16384)
int sort_match_inplace(const float *target, float *source_scratch, float *out, int c, int64_t n_t, int64_t n_s,
                       int32_t *perm, void *ws, size_t ws_bytes, cudaStream_t st);

// cov_match.cu: closed-form Gaussian matching (histmatch.py:13-44) on NHWC-flattened data.
//   out[n, c] = (X - mu_t) T^T + mu_s ;  T from chol / pca / sym of the two covariances
size_t cov_match_ws_bytes(int64_t n_t, int64_t n_s, int c, int mode);
// the whole OT step for the covariance modes (rotation folded algebraically, see cov_match.cu); R may be null
int cov_ot_step(const float *P, const float *S, const float *R, float *out, int b_p, int64_t hw_p, int b_s,
                int64_t hw_s, int c, int mode, float eps, const float *content, float strength, void *workspace,
                size_t workspace_bytes, cudaStream_t st, int style_reuse = 0);
// narrow blocks (c <= 64, chol / pca / sym): cov_small.cu - five launches per step, the C x C chain on one CTA
bool cov_small_supported(int c, int mode, int b_p, int b_s);
size_t cov_small_ws_bytes(int64_t n_t, int64_t n_s, int c);
int cov_small_step(const float *P, const float *S, const float *R, float *out, int b_p, int64_t hw_p, int b_s,
                   int64_t hw_s, int c, int mode, float eps, const float *content, float strength, void *workspace,
                   size_t workspace_bytes, cudaStream_t st, int style_reuse);
// 64 < c <= 384 (pca / sym): cov_chain.cu - the whole C x C chain (Newton-Schulz, closing products, bias) as one
// cooperative kernel on the 19 c x c matrices of cov_match.cu's workspace
bool cov_coop_supported(int c, int mode);
size_t cov_coop_scratch_floats();
int cov_coop_chain(float *const *m, int c, int mode, float eps, int style_reuse, const float *mu_p, const float *mu_s,
                   int b_p, int b_s, float *bias, float *scratch, cudaStream_t st);
int cov_match_nhwc(const float *target, const float *source, float *out, int b_t, int64_t hw_t, int b_s,
                   int64_t hw_s, int c, int mode, float eps, void *workspace, size_t workspace_bytes,
                   cudaStream_t st);

}  // namespace optex
