// Standalone probe of the TMA -> smem -> tcgen05.mma -> TMEM path (debug aid, not product code).
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../optimaltextures_b200/csrc/gemm_tcgen05.cu"
namespace optex {
void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
int cuda_fail(cudaError_t e, const char *what) { fprintf(stderr, "CUDA %s at %s\n", cudaGetErrorString(e), what); return 3; }
void count_launch(int) {}
int sm_count() { return 148; }
int require_sm100() { return 0; }
}
using namespace optex;

__global__ void __launch_bounds__(192, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float *dbg, int a_mn, int nkk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar, accum_bar;
    __shared__ uint32_t tmem_base_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
    uint8_t *a_dst = tiles, *b_dst = tiles + 16384;
    if (threadIdx.x == 0) { mbar_init(&full_bar, 1); mbar_init(&accum_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&full_bar, 16384 + 8192);
        if (a_mn) tma_load_3d(&tmA, &full_bar, a_dst, 0, 0, 0, L2_EVICT_NORMAL); else tma_load_2d(&tmA, &full_bar, a_dst, 0, 0, L2_EVICT_NORMAL);
        tma_load_2d(&tmB, &full_bar, b_dst, 0, 0, L2_EVICT_NORMAL);
    }
    mbar_wait(&full_bar, 0);
    // dump smem: first 256 floats of A tile, first 256 floats of B tile
    for (int i = threadIdx.x; i < 256; i += 192) { dbg[i] = ((float *)a_dst)[i]; dbg[256 + i] = ((float *)b_dst)[i]; dbg[512 + i] = ((float *)a_dst)[1024 + i]; }
    if (threadIdx.x == 0) dbg[2000] = __uint_as_float(tmem_base);
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int kk = 0; kk < nkk; ++kk) {
                uint64_t da = a_mn ? make_desc(smem_u32(a_dst) + kk * 1024, 4096, 512, 1) : make_desc(smem_u32(a_dst) + kk * 32, 16, 1024);
                uint64_t db = make_desc(smem_u32(b_dst) + kk * 32, 16, 1024);
                umma_tf32(tmem_base, da, db, idesc, kk != 0);
            }
            umma_commit(&accum_bar);
        }
        __syncwarp();
    }
    if (warp >= 2) {
        mbar_wait(&accum_bar, 0);
        tc_fence_after();
        int q = warp & 3;
        for (int col = 0; col < 64; col += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + col, v);
            for (int j = 0; j < 32; ++j) dbg[4096 + (q * 32 + lane) * 64 + col + j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory"); }
}

int main() {
    const int M = 128, N = 64, K = 32;
    std::vector<float> A(M * K), B(N * K);
    for (int i = 0; i < M * K; ++i) A[i] = (float)((i * 7) % 13) - 6.0f;
    for (int i = 0; i < N * K; ++i) B[i] = (float)((i * 5) % 11) - 5.0f;
    float *dA, *dB, *dbg;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dbg, (4096 + 128 * 64) * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dbg, 0xff, (4096 + 128 * 64) * 4);
    CUtensorMap ta, tb;
    if (make_map_kmajor(&ta, dA, M, K, 128) || make_map_kmajor(&tb, dB, N, K, 64)) return 1;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    std::vector<float> At(K * M);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) At[k * M + m] = A[m * K + k];
    float *dAt; cudaMalloc(&dAt, At.size() * 4); cudaMemcpy(dAt, At.data(), At.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap tat;
    if (make_map_mnmajor(&tat, dAt, K, M, 128)) return 1;
    for (int variant = 0; variant < 4; ++variant) {
        int nkk = (variant & 1) ? 4 : 1; int a_mn = variant >> 1;
        cudaMemset(dbg, 0xff, (4096 + 128 * 64) * 4);
        printf("---- a_mn=%d\n", a_mn);
        probe_kernel<<<1, 192, 64 * 1024>>>(a_mn ? tat : ta, tb, dbg, a_mn, nkk);
        cudaError_t e = cudaDeviceSynchronize();
        printf("nkk=%d kernel: %s\n", nkk, cudaGetErrorString(e));
        std::vector<float> h(4096 + 128 * 64);
        cudaMemcpy(h.data(), dbg, h.size() * 4, cudaMemcpyDeviceToHost);
        printf("A smem[0..8): "); for (int i = 0; i < 8; ++i) printf("%g ", h[i]); printf(" | A gmem: "); for (int i = 0; i < 8; ++i) printf("%g ", A[i]); printf("\n");
        printf("A smem row1[32..40): "); for (int i = 32; i < 40; ++i) printf("%g ", h[i]); printf(" | A gmem row1: "); for (int i = 32; i < 40; ++i) printf("%g ", A[i]); printf("\n");
        printf("A smem +4KB [0..8): "); for (int i = 0; i < 8; ++i) printf("%g ", h[512 + i]); printf(" | At row0 [32..40): "); for (int i = 32; i < 40; ++i) printf("%g ", At[i]); printf("\n");
        printf("B smem[0..8): "); for (int i = 0; i < 8; ++i) printf("%g ", h[256 + i]); printf("\n");
        unsigned tb_; memcpy(&tb_, &h[2000], 4); printf("tmem_base=0x%08x\n", tb_);
        double maxerr = 0; int nz = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0; for (int k = 0; k < 8 * nkk; ++k) ref += (double)A[m * K + k] * B[n * K + k];
            double got = h[4096 + m * 64 + n]; if (got != 0) nz++;
            double er = fabs(ref - got); if (er > maxerr) maxerr = er;
        }
        printf("D: nonzero=%d maxerr=%g  D[0][0..4)=%g %g %g %g  D[1][0]=%g D[64][3]=%g\n", nz, maxerr, h[4096], h[4097], h[4098], h[4099], h[4096 + 64], h[4096 + 64 * 64 + 3]);
    }
    return 0;
}
