// Rotation GEMMs on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   forward  optex.py:170-171   D[pixel, c_out] = sum_k X[pixel, k] R[k, c_out]
//   inverse  optex.py:175       D[pixel, j]     = sum_c M(pixel, c) R[j, c]   (+ blend, optex.py:117)
//
// Structure (one CTA per 128 x BLOCK_N output tile, 192 threads):
//   warp 0   TMA producer : cp.async.bulk.tensor -> 128B-swizzled smem stages, mbarrier expect_tx
//   warp 1   MMA issuer   : one elected lane issues tcgen05.mma.kind::tf32, fp32 accumulator in TMEM,
//                           tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2-5 epilogue    : tcgen05.ld (TMEM -> registers) -> coalesced global stores
//                           (channel-major for the forward rotation, NHWC + content blend for the inverse)
//
// Precision: kind::tf32 keeps 10 mantissa bits.  terms == 3 runs the 3xTF32 split
//   a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi   (a_hi = tf32(a), a_lo = tf32(a - a_hi), fp32 accumulate)
// which restores fp32-grade products (the parity tests hold it to the fp32 tolerance);
// terms == 1 is plain TF32, the reference's own CUDA default (optex.py:248-249).
//
// Both operand majors are used without any transposition pass:
//   K-major  (X[n, k], R[j, c])   : 2-D tensor map, box {32 k, rows}            -> canonical K-major SW128
//   MN-major (R[k, c_out], Mt[c, n]): 3-D tensor map {32 mn, k, mn/32}, box {32, 32, rows/32}
//                                     -> canonical MN-major SW128 (LBO = 4 KB between 32-wide mn blocks)
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <vector>

#include "gemm_tc.cuh"

namespace optex {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;  // K granularity of split-K slices (both stage depths divide it)
constexpr int UMMA_K = 8;    // 32 B of K per tcgen05.mma.kind::tf32
constexpr int MAX_STAGES = 8;
constexpr int MAX_BATCH = 4;     // independent problems per launch (gemm_tc batch mode)
// The epilogue (TMEM -> registers -> global) of a 128 x 256 tile took 20 k cycles with 4 warps - as long as a whole
// plain-TF32 main loop and exposed after the last tile; EPI_GROUPS groups of 4 warps (one warp per TMEM lane quadrant)
// split the tile's columns between them.
constexpr int EPI_GROUPS = 4;
// warp roles are aligned to warpgroups (setmaxnreg works on 4 warps at a time):
//   warps 0-3: TMA producer, MMA issuer, two idle | warps 4-7: converters | warps 8..: 4 x EPI_GROUPS epilogue warps
constexpr int EPI_WARP0 = 8;
constexpr int NTHREADS = EPI_WARP0 * 32 + EPI_GROUPS * 128;
// A-via-TMEM (3xTF32, see the kernel comment): the accumulator keeps columns [0, BLOCK_N), the tf32 hi / lo halves of
// the A tile live in a ring of TMEM stages behind it (32 + 32 columns per stage)
constexpr int A_TMEM_COL0 = 256;
constexpr int A_TMEM_STAGES = 4;
constexpr int SMEM_BUDGET = 224 * 1024;  // tiles; + 1 KB alignment slack + 1 KB static (barriers) = 226 of 227 KB
constexpr int EPI_STAGE_BYTES = 128 * 32 * 4;  // TMA-store epilogue: one 128 x 32 fp32 staging tile per epilogue group

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float lds_f32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void lds_v4(uint32_t saddr, float &a, float &b, float &c, float &d) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(saddr) : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred P;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
// L2 eviction-priority hints for TMA loads (the encodings CUTLASS uses for createpolicy results)
constexpr uint64_t L2_EVICT_NORMAL = 0x1000000000000000ull;
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;  // streamed once: do not displace the step's intermediates
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;   // re-read by every tile (the rotation matrix)
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int x, int y,
                                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap *map, uint64_t *bar, void *dst, int x, int y, int z,
                                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1, int c2,
                                            int c3, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int x, int y) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int x, int y, int z) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from tensor memory (lane = row of the tile, one 32-bit column per k), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
// ---- 2-CTA (cta_group::2) forms: a pair of CTAs (cluster of 2) works on one 256-row tile; each CTA loads HALF of the
// B tile and the tensor cores of both SMs read both halves, so the L2 -> SM traffic per MMA halves
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// (default semantics, release at CTA scope - what CUTLASS's ClusterBarrier::arrive(cta_id) uses: a release at CLUSTER
// scope costs a membar + L1 invalidate per arrive and measured 1000-2000 clk per k block in the converter warps; no
// generic-proxy data crosses the pair here, the TMEM hand-off is ordered by the tcgen05 fences)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred P;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
// TMA load whose completion is signalled on a barrier of EITHER CTA of the pair (bar_cluster: shared::cluster address)
__device__ __forceinline__ void tma_load_2d_cg2(const CUtensorMap *map, uint32_t bar_cluster, void *dst, int x, int y,
                                                uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(x), "r"(y), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_cg2(const CUtensorMap *map, uint32_t bar_cluster, void *dst, int x, int y,
                                                int z, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(x), "r"(y), "r"(z), "l"(hint)
        : "memory");
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_ts_cg2(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// TMA store of a staged tile (shared -> global, bulk async group), and the waits on it
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(x), "r"(y)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), version 1.
//   K-major operands : LayoutType::SWIZZLE_128B (2), 8-row x 128 B atoms, SBO = 1024 B
//   MN-major tf32    : LayoutType::SWIZZLE_128B_BASE32B (1) - the ONLY layout the tensor core accepts for
//                      32-bit MN-major operands: 4 k-rows x 128 B atoms whose 32 B chunks are XOR-ed with the
//                      row index (TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), SBO = 512 B, LBO = mn-block pitch
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;
    return d;
}

// a -> (tf32(a), tf32(a - tf32(a))): round-to-nearest split, both halves exactly representable in tf32.
// tf32(x) = cvt.rna.tf32.f32 = round the magnitude to 10 mantissa bits, ties away from zero: on the bit pattern that
// is "add half an ulp, clear the 13 low bits" (inf stays inf, FLT_MAX rounds to inf, NaN stays NaN - what the cvt
// instruction returns as well); ptxas expands the cvt into four instructions with an explicit inf test, this is two.
__device__ __forceinline__ float rna_tf32(float a) {
    return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float a, float &h, float &l) {
    h = rna_tf32(a);
    l = rna_tf32(__fsub_rn(a, h));
}

struct Params {
    float *D;
    int64_t ldd;
    int64_t M, N, K;
    const float *blend;   // optex.py:117 content blend (same indexing as D), or null
    float strength;
    int terms;
    int stages;
    int64_t k_per_z;      // split-K: blockIdx.z covers K range [z*k_per_z, +k_per_z) (multiple of 32)
    int64_t d_z_stride;   // ... and writes its partial sum to D + z*d_z_stride
    const float *bias;    // + bias[(row / bias_hw) * bias_ld + col]  (per-sample mean term), or null
    int64_t bias_hw, bias_ld;
    float alpha;          // D = alpha * acc (+ bias) (then blend)
    const int *skip;      // device flag: non-zero -> the whole launch is a no-op (converged iterations)
    int conv_a;           // terms == 3: A arrives raw (tmA_hi) and warps 2-5 split it into hi/lo in shared memory
    int nz;               // number of split-K slices (tiles enumerate z as well)
    uint32_t *colrange;   // D_TRANS only: per output column [2]: atomicMin of f2ord(v) and of ~f2ord(v) (range fold)
    uint64_t a_hint, b_hint;    // L2 eviction priority of the operand loads
    unsigned long long *trace;  // debug: CTA 0 records clock stamps of its first tile's pipeline events (or null)
    uint32_t *rowrange;   // !D_TRANS: the same per output ROW (rows = channels in the fused loop's mid rotation)
    int conv_b;           // same for B (small problems, where a pre-split pass per GEMM would dominate)
    float diag;           // !D_TRANS: + diag on the main diagonal (after alpha), Newton-Schulz T = 1.5 I - 0.5 Z Y
    float *resid_max;     // !D_TRANS: atomicMax of max |acc - I| over the output (acc before alpha)
    const float *skip_below;  // *skip_below < skip_tol -> no-op launch (a converged Newton-Schulz chain)
    float skip_tol;
    int relu;             // !D_TRANS: max(., 0) after the bias (conv layers, vgg.py)
    // batch > 0: tiles enumerate `batch` independent problems (z) whose operands are matrices of one arena each:
    // row offsets into the A / B tensor maps, element offset of D, per-problem residual slot and skip value
    int batch;
    int a_off[MAX_BATCH], b_off[MAX_BATCH];
    int64_t d_off[MAX_BATCH];
    float *resid_z[MAX_BATCH];
    const float *skip_z[MAX_BATCH];
    int a_tmem;           // conv_a && !conv_b: the converters write the hi / lo halves of A to TENSOR memory
    int stages_a;         // a_tmem: depth of the raw-A shared-memory ring (p.stages is the B ring then)
    int tma_store;        // CG == 2: the epilogue stages 128 x 32 tiles in shared memory and writes them with TMA
    uint32_t range_off;   // byte offset of that array from the aligned tile base
    int colrange_smem;    // D_TRANS: the range fold accumulates in shared memory (2 * N words behind the tiles) and
                          // reaches global memory once per CTA - per-tile global atomics on 2 * N addresses serialise
                          // in L2 (conv1_1 @ 1024^2: 16 384 tiles x 512 atomics on 128 addresses = 0.65 ms)
    // dual (a_tmem launches): a SECOND problem with the same B, N and K - the forward rotations of the pastiche and of
    // the style block in one launch (optex.py:170-171).  Its A map travels in tmA_lo (unused by the TMEM-A form), its
    // output map in tmD2; tiles of problem 0 come first in the tile list.
    int64_t M2, ldd2;
    float *D2;
    // !D_TRANS: alpha and diag read from DEVICE memory (coef[0], coef[1]) - the scaled Newton-Schulz step, whose
    // coefficients depend on a norm computed on the device (cov_match.cu); per problem in batch mode
    const float *coef;
    const float *coef_z[MAX_BATCH];
    // conv (TMEM-A launches, K-major A): implicit GEMM of a 3 x 3 convolution (vgg.cu).  The A operand is never
    // materialised: an M tile is a CONV_TH x CONV_TW block of output pixels of one image and k block kb = 32 input
    // channels of tap kb * 32 / c_in, loaded by ONE 4-D TMA box {32 ch, CONV_TW, CONV_TH, 1} at the tap's offset from
    // the reflection-padded NHWC input [b, h + 2, w + 2, c_in] - which lands exactly in the K-major SWIZZLE_128B tile
    // layout of an explicit im2col row block.  p.M = (number of pixel tiles) x 128 x CG virtual rows.
    int conv, cv_h, cv_w, cv_cin, cv_tx, cv_ty;   // output height / width, input channels, tiles along x / y per image
};
constexpr int CONV_TW = 16, CONV_TH = 8;   // 128 pixels per CTA tile; a pair stacks two of them along y

// ------------------------------------------------------------------ the kernel
// Persistent, warp-specialised: grid = min(#tiles, #SMs); every CTA walks tiles  t = blockIdx.x, += gridDim.x.
//   warp 0      TMA producer        raw/full barriers per smem stage (expect_tx)
//   warp 1      MMA issuer          tcgen05.mma into a TMEM accumulator
//   warps 4-7   converters (3xTF32) split the raw A tile of every stage into tf32 hi / lo halves
//   warps 8-23  epilogue            tcgen05.ld of the whole tile into registers, accumulator released at once, then
//                                   the global stores run while the MMA warp is already in the next tile
// so the prologue (barrier init, TMEM alloc) is paid once per SM and the epilogue is hidden behind the next tile's
// main loop.  Three operand paths:
//   terms == 1            plain TF32: raw fp32 tiles go straight from TMA to the tensor core
//   conv_a (+ conv_b)     3xTF32, hi / lo halves written back to SHARED memory (small problems: B converted too)
//   a_tmem                3xTF32, hi / lo halves of A written to TENSOR memory and fed to tcgen05.mma as the TMEM A
//                         operand.  The SS form is shared-memory-bandwidth bound: per 32-wide k block of a 128 x 256
//                         tile TMA writes 80 KB, the converters move 48 KB and the three MMAs read 144 KB = 272 KB at
//                         128 B/clk = 2100 clk against 1536 clk of tensor-pipe time; with A in TMEM the converters
//                         only read (16 KB) and the MMAs only fetch B (96 KB): 192 KB = 1500 clk.
// BK = K elements per pipeline stage: 32 (128-byte swizzle rows) or 16 (64-byte rows, half-size stages).
// CG == 2 (a_tmem launches only): clusters of two CTAs share one 256 x BLOCK_N tile.  CTA r of the pair owns rows
// [m0 + 128 r, +128): it streams and converts its own A rows into its own tensor memory and loads the B columns
// [n0 + r BLOCK_N / 2, + BLOCK_N / 2); the leader (r = 0) issues tcgen05.mma.cta_group::2 for both.  Barriers the
// leader's MMA warp waits on (B landed, A in TMEM, accumulator drained) live in the LEADER's shared memory and
// collect arrivals from both CTAs; barriers the MMA releases are signalled in both CTAs by a multicast commit.
template <int BLOCK_N, bool A_MN, bool B_MN, bool D_TRANS, int BK, int CG = 1>
__global__ void __launch_bounds__(NTHREADS, 1)
rotate_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmD2, const Params p) {
    constexpr uint32_t A_TILE = BLOCK_M * BK * 4;
    constexpr uint32_t B_TILE = (BLOCK_N / CG) * BK * 4;   // what THIS CTA loads per stage and half (hi / lo)
    static_assert(BK == 32 || BK == 16, "stage depth");
    static_assert(CG == 1 || CG == 2, "cta_group");
    const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;   // rank in the pair; 0 = leader (issues the MMAs)
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], raw_bar[MAX_STAGES];
    __shared__ __align__(8) uint64_t afree_bar[MAX_STAGES];                            // a_tmem: raw A slot consumed
    __shared__ __align__(8) uint64_t ta_full_bar[A_TMEM_STAGES], ta_empty_bar[A_TMEM_STAGES];  // a_tmem: TMEM A ring
    __shared__ __align__(8) uint64_t tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[64] = clock64();
    pdl_wait();
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[65] = clock64();
    // debug: trace[95] == 0x7ace asks every CTA for its own (clock64, globaltimer) at entry and exit, 4 words per CTA
    // from trace[128] (the buffer must hold 128 + 4 * gridDim.x words)
    const bool cta_stamps = p.trace && threadIdx.x == 0 && p.trace[95] == 0x7aceull;
    if (cta_stamps) {
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        p.trace[128 + 4 * blockIdx.x + 0] = clock64();
        p.trace[128 + 4 * blockIdx.x + 1] = g;
    }
    // ... and CTA 0 stamps the waits of its producer / MMA / converter roles for its first 32 k blocks:
    // trace[1024 + ((role * 32 + kblock) * 4 + j)]
    const bool role_stamps = p.trace && blockIdx.x == 0 && p.trace[95] == 0x7aceull;
    auto stamp = [&](int role, int kbi, int j) {
        if (role_stamps && kbi < 32) p.trace[1024 + ((role * 32 + kbi) * 4 + j)] = clock64();
    };
    if (p.skip && *p.skip) return;  // uniform over the grid
    if (p.skip_below && *p.skip_below < p.skip_tol) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool a_tmem = p.a_tmem != 0;
    const int nterm_tiles = p.terms == 3 ? 2 : 1;  // hi (+ lo) tiles per operand
    // a_tmem: a stage of the main ring holds the hi / lo tiles of B only; raw A tiles have their own ring behind it
    const uint32_t stage_bytes = a_tmem ? 2 * B_TILE : nterm_tiles * (A_TILE + B_TILE);
    const int n_tiles_n = (int)((p.N + BLOCK_N - 1) / BLOCK_N);
    const int n_tiles_m = (int)((p.M + BLOCK_M * CG - 1) / (BLOCK_M * CG));
    const int tiles_per_z = n_tiles_m * n_tiles_n;
    const int tiles_dual = (int)((p.M2 + BLOCK_M * CG - 1) / (BLOCK_M * CG)) * n_tiles_n;   // second problem (or 0)
    const int num_tiles = tiles_per_z * p.nz + tiles_dual;
    // 1024-byte alignment of the dynamic smem base (swizzle atoms)
    uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
    uint8_t *a_ring = tiles + (size_t)p.stages * stage_bytes;
    uint8_t *epi_stage = a_ring + (size_t)p.stages_a * A_TILE;   // tma_store: EPI_GROUPS staging tiles
    // colrange_smem: [2 * N] words at the very end of the dynamic allocation (the host reserved them)
    uint32_t *s_range = reinterpret_cast<uint32_t *>(tiles + p.range_off);
    if (D_TRANS && p.colrange_smem)
        for (int i = threadIdx.x; i < 2 * (int)p.N; i += NTHREADS) s_range[i] = 0xffffffffu;   // before the __syncthreads below
    // a_tmem: ONE accumulator (the other 256 columns hold the A ring); the epilogue frees it as soon as the tile
    // sits in registers.  Otherwise two accumulators alternate.
    const uint32_t tmem_cols = a_tmem ? 512u : (uint32_t)(2 * BLOCK_N);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi);
        prefetch_tmap(&tmB_hi);
        if (CG == 2 && p.tma_store) prefetch_tmap(&tmD);
        if (p.M2 > 0) { prefetch_tmap(&tmA_lo); if (CG == 2 && p.tma_store) prefetch_tmap(&tmD2); }
        if (p.terms == 3) {
            if (!p.conv_a) prefetch_tmap(&tmA_lo);
            if (!p.conv_b) prefetch_tmap(&tmB_lo);
        }
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], (p.conv_a && !a_tmem) ? 4 : 1);  // conv_a: one arrival per converter warp
            mbar_init(&empty_bar[s], 1);
            mbar_init(&raw_bar[s], 1);
        }
        if (a_tmem) {
            for (int s = 0; s < p.stages_a; ++s) {
                mbar_init(&raw_bar[s], 1);
                mbar_init(&afree_bar[s], 4);
            }
            for (int s = 0; s < A_TMEM_STAGES; ++s) {
                mbar_init(&ta_full_bar[s], 4 * CG);   // one arrival per converter warp (of both CTAs of a pair)
                mbar_init(&ta_empty_bar[s], 1);
            }
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);   // tcgen05.commit of the tile's last MMA
            mbar_init(&tempty_bar[a], 4 * EPI_GROUPS * CG);  // one arrival per epilogue warp (of both CTAs)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(&tmem_base_smem)),
                         "r"(tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(&tmem_base_smem)),
                         "r"(tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[66] = clock64();

    // tile -> (z, m, n): n fastest, so CTAs running side by side share the rows of the big A operand in L2
    // (dual launches have nz == 1: z == 1 then means "the second problem")
    auto tile_coords = [&](int tile, int &m0, int &n0, int &z) {
        z = tiles_dual > 0 ? (tile >= tiles_per_z ? 1 : 0) : tile / tiles_per_z;
        const int r = tile - z * tiles_per_z;
        m0 = (r / n_tiles_n) * (BLOCK_M * CG) + (int)crank * BLOCK_M;   // this CTA's rows of the (pair's) tile
        n0 = (r % n_tiles_n) * BLOCK_N;
    };
    // conv: (image, first pixel row / column) of this CTA's block of output pixels
    auto conv_coords = [&](int m0, int &bi, int &y0, int &x0) {
        const int mt = m0 / (BLOCK_M * CG);          // pixel-tile index (of the pair)
        const int per_img = p.cv_tx * p.cv_ty;
        bi = mt / per_img;
        const int rem = mt - bi * per_img;
        y0 = (rem / p.cv_tx) * (CONV_TH * CG) + (int)crank * CONV_TH;
        x0 = (rem % p.cv_tx) * CONV_TW;
    };
    const int tile0 = (int)blockIdx.x / CG, tile_step = (int)gridDim.x / CG;   // pairs walk the tile list together
    auto k_range = [&](int z, int64_t &k_begin, int &num_kb) {
        if (tiles_dual > 0) z = 0;   // both problems of a dual launch span the whole K
        k_begin = (int64_t)z * p.k_per_z;
        const int64_t k_len = p.k_per_z > 0 ? (p.K - k_begin < p.k_per_z ? p.K - k_begin : p.k_per_z) : p.K;
        num_kb = (int)((k_len + BK - 1) / BK);
    };
    // batch mode: a problem whose chain has converged contributes no tiles (every role skips them alike)
    auto z_skipped = [&](int z) -> bool {
        return p.batch > 0 && p.skip_z[z] != nullptr && *p.skip_z[z] < p.skip_tol;
    };
    // which accumulator / barrier phase the i-th tile of this CTA uses
    // (TMEM-A form: the A ring owns columns 256..511, so only tiles up to 128 wide leave room for a second accumulator)
    const bool one_acc = a_tmem && BLOCK_N > 128;
    auto acc_of = [&](int titer, int &acc, uint32_t &aph) {
        acc = one_acc ? 0 : (titer & 1);
        aph = one_acc ? (uint32_t)(titer & 1) : (uint32_t)((titer >> 1) & 1);
    };

    if (warp < 4) {
        setmaxnreg_dec<48>();
        if (warp == 0) {
            // ===== TMA producer
            if (elect_one()) {
                int s = 0, sa = 0;
                uint32_t ph = 0, pha = 0;
                for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                    int m0, n0, z, num_kb;
                    int64_t k_begin;
                    tile_coords(tile, m0, n0, z);
                    k_range(z, k_begin, num_kb);
                    if (z_skipped(z)) continue;
                    // batch mode: the problem's operands sit a_off / b_off rows into the arena the maps describe
                    const int ao = p.batch > 0 ? p.a_off[z] : 0, bo = p.batch > 0 ? p.b_off[z] : 0;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        const int k0 = (int)k_begin + kb * BK;
                        const bool tr = p.trace && blockIdx.x == 0 && tile == tile0 && kb < 16;
                        if (a_tmem) {
                            // raw A -> its own ring (freed by the converters as soon as they have read it)
                            const int kbi = (tile - tile0) / tile_step * num_kb + kb;
                            stamp(0, kbi, 0);
                            mbar_wait(&afree_bar[sa], pha ^ 1);
                            stamp(0, kbi, 1);
                            if (tr) p.trace[kb * 4 + 0] = clock64();
                            uint8_t *a_dst = a_ring + (size_t)sa * A_TILE;
                            mbar_expect_tx(&raw_bar[sa], A_TILE);
                            const CUtensorMap *ma = (tiles_dual > 0 && z == 1) ? &tmA_lo : &tmA_hi;
                            if (!A_MN && p.conv) {
                                int bi, y0, x0;
                                conv_coords(m0, bi, y0, x0);
                                const int tap = k0 / p.cv_cin, ch0 = k0 - tap * p.cv_cin;
                                tma_load_4d(ma, &raw_bar[sa], a_dst, ch0, x0 + tap % 3, y0 + tap / 3, bi, p.a_hint);
                            } else if (A_MN) tma_load_3d(ma, &raw_bar[sa], a_dst, 0, k0 + ao, m0 / 32, p.a_hint);
                            else tma_load_2d(ma, &raw_bar[sa], a_dst, k0, m0 + ao, p.a_hint);
                            if (++sa == p.stages_a) { sa = 0; pha ^= 1; }
                            // pre-split B hi / lo -> the main ring (freed by the MMAs)
                            stamp(0, kbi, 2);
                            mbar_wait(&empty_bar[s], ph ^ 1);
                            stamp(0, kbi, 3);
                            uint8_t *b_dst = tiles + (size_t)s * stage_bytes;
                            if (CG == 2) {
                                // this CTA's half of the B columns; completion is counted on the LEADER's barrier, which
                                // expects the bytes of both halves (the peer's may land before the leader has armed it:
                                // the pending arrival keeps the phase open)
                                const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
                                if (crank == 0) mbar_expect_tx(&full_bar[s], 2 * 2 * B_TILE);
                                const int nb0 = n0 + (int)crank * (BLOCK_N / 2);
                                for (int t = 0; t < 2; ++t) {
                                    const CUtensorMap *mb = t ? &tmB_lo : &tmB_hi;
                                    if (B_MN) tma_load_3d_cg2(mb, fb, b_dst + t * B_TILE, 0, k0 + bo, nb0 / 32, p.b_hint);
                                    else tma_load_2d_cg2(mb, fb, b_dst + t * B_TILE, k0, nb0 + bo, p.b_hint);
                                }
                            } else {
                                mbar_expect_tx(&full_bar[s], 2 * B_TILE);
                                for (int t = 0; t < 2; ++t) {
                                    const CUtensorMap *mb = t ? &tmB_lo : &tmB_hi;
                                    if (B_MN) tma_load_3d(mb, &full_bar[s], b_dst + t * B_TILE, 0, k0 + bo, n0 / 32, p.b_hint);
                                    else tma_load_2d(mb, &full_bar[s], b_dst + t * B_TILE, k0, n0 + bo, p.b_hint);
                                }
                            }
                            if (++s == p.stages) { s = 0; ph ^= 1; }
                            continue;
                        }
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        if (tr) p.trace[kb * 4 + 0] = clock64();
                        uint8_t *st = tiles + (size_t)s * stage_bytes;
                        // conv_a: the raw A tile lands in the hi slot and signals the converter warps (raw_bar);
                        // otherwise hi and lo halves of both operands arrive pre-split and signal the MMA warp directly
                        uint64_t *bar = p.conv_a ? &raw_bar[s] : &full_bar[s];
                        mbar_expect_tx(bar, p.conv_a ? A_TILE + (p.conv_b ? 1 : nterm_tiles) * B_TILE : stage_bytes);
                        for (int t = 0; t < nterm_tiles; ++t) {
                            const CUtensorMap *ma = t ? &tmA_lo : &tmA_hi;
                            const CUtensorMap *mb = t ? &tmB_lo : &tmB_hi;
                            uint8_t *a_dst = st + t * A_TILE;
                            uint8_t *b_dst = st + nterm_tiles * A_TILE + t * B_TILE;
                            if (t == 0 || !p.conv_a) {
                                if (A_MN) tma_load_3d(ma, bar, a_dst, 0, k0 + ao, m0 / 32, p.a_hint);
                                else tma_load_2d(ma, bar, a_dst, k0, m0 + ao, p.a_hint);
                            }
                            if (t == 0 || !p.conv_b) {
                                if (B_MN) tma_load_3d(mb, bar, b_dst, 0, k0 + bo, n0 / 32, p.b_hint);
                                else tma_load_2d(mb, bar, b_dst, k0, n0 + bo, p.b_hint);
                            }
                        }
                        if (++s == p.stages) { s = 0; ph ^= 1; }
                    }
                }
            }
        } else if (warp == 1 && crank == 0) {
            // ===== MMA issuer  (the TMEM A operand is always "K-major": lane = row, column = k); pairs: leader only
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (((A_MN && !a_tmem) ? 1u : 0u) << 15) |
                                   ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                                   ((uint32_t)((BLOCK_M * CG) >> 4) << 24);
            int s = 0, sta = 0;
            uint32_t ph = 0, phta = 0;
            int titer = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                int m0, n0, z, num_kb, acc;
                int64_t k_begin;
                uint32_t aph;
                tile_coords(tile, m0, n0, z);
                k_range(z, k_begin, num_kb);
                if (z_skipped(z)) continue;
                acc_of(titer++, acc, aph);
                if (CG == 2) mbar_wait_cluster(&tempty_bar[acc], aph ^ 1);
                else mbar_wait(&tempty_bar[acc], aph ^ 1);  // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int kbi = (titer - 1) * num_kb + kb;
                    if (lane == 0) stamp(1, kbi, 0);
                    if (CG == 2) {
                        mbar_wait_cluster(&full_bar[s], ph);
                        if (lane == 0) stamp(1, kbi, 1);
                        mbar_wait_cluster(&ta_full_bar[sta], phta);
                    } else {
                        mbar_wait(&full_bar[s], ph);
                        if (lane == 0) stamp(1, kbi, 1);
                        if (a_tmem) mbar_wait(&ta_full_bar[sta], phta);
                    }
                    if (lane == 0) stamp(1, kbi, 2);
                    tc_fence_after();
                    if (p.trace && blockIdx.x == 0 && tile == tile0 && kb < 16 && lane == 0) p.trace[kb * 4 + 3] = clock64();
                    if (p.trace && blockIdx.x == 0 && tile != tile0 && (kb == 0 || kb == num_kb - 1) && lane == 0) p.trace[76 + (kb ? 1 : 0)] = clock64();
                    if (elect_one()) {
                        const uint32_t st = smem_u32(tiles + (size_t)s * stage_bytes);
                        // MN-major: 32-wide mn blocks of BK k-rows x 128 B (LBO), 4-row BASE32B atoms (SBO 512 B)
                        // K-major : 8-row atoms of BK*4-byte rows: SWIZZLE_128B (BK = 32) or SWIZZLE_64B (BK = 16)
                        constexpr uint32_t kmaj_sbo = BK == 32 ? 1024u : 512u, kmaj_lt = BK == 32 ? 2u : 4u;
                        const uint32_t a_lbo = A_MN ? BK * 128u : 16u, b_lbo = B_MN ? BK * 128u : 16u;
                        const uint32_t a_sbo = A_MN ? 512u : kmaj_sbo, b_sbo = B_MN ? 512u : kmaj_sbo;
                        const uint32_t a_lt = A_MN ? 1u : kmaj_lt, b_lt = B_MN ? 1u : kmaj_lt;
                        if (a_tmem) {
                            const uint32_t b_hi = st, b_lo = st + B_TILE;
                            const uint32_t ta_hi = tmem_base + (uint32_t)(A_TMEM_COL0 + sta * 2 * BK), ta_lo = ta_hi + BK;
#pragma unroll
                            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                                const uint32_t b_off = B_MN ? kk * 1024u : kk * 32u;
                                const uint64_t db = make_desc(b_hi + b_off, b_lbo, b_sbo, b_lt);
                                const uint64_t dbl = make_desc(b_lo + b_off, b_lbo, b_sbo, b_lt);
                                if (CG == 2) {
                                    umma_tf32_ts_cg2(tmem_d, ta_hi + kk * UMMA_K, db, idesc, (kb | kk) != 0);
                                    umma_tf32_ts_cg2(tmem_d, ta_hi + kk * UMMA_K, dbl, idesc, 1u);
                                    umma_tf32_ts_cg2(tmem_d, ta_lo + kk * UMMA_K, db, idesc, 1u);
                                } else {
                                    umma_tf32_ts(tmem_d, ta_hi + kk * UMMA_K, db, idesc, (kb | kk) != 0);
                                    umma_tf32_ts(tmem_d, ta_hi + kk * UMMA_K, dbl, idesc, 1u);
                                    umma_tf32_ts(tmem_d, ta_lo + kk * UMMA_K, db, idesc, 1u);
                                }
                            }
                            if (CG == 2) umma_commit_pair(&ta_empty_bar[sta]);
                            else umma_commit(&ta_empty_bar[sta]);  // TMEM A stage free once these MMAs retire
                        } else {
                            const uint32_t a_hi = st, a_lo = st + A_TILE;
                            const uint32_t b_hi = st + nterm_tiles * A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
                            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                                // K-major: +32 B per K step inside the swizzle row; MN-major: +8 k-rows = 1 KB
                                const uint32_t a_off = A_MN ? kk * 1024u : kk * 32u;
                                const uint32_t b_off = B_MN ? kk * 1024u : kk * 32u;
                                uint64_t da = make_desc(a_hi + a_off, a_lbo, a_sbo, a_lt);
                                uint64_t db = make_desc(b_hi + b_off, b_lbo, b_sbo, b_lt);
                                umma_tf32(tmem_d, da, db, idesc, (kb | kk) != 0);
                                if (p.terms == 3) {
                                    uint64_t dal = make_desc(a_lo + a_off, a_lbo, a_sbo, a_lt);
                                    uint64_t dbl = make_desc(b_lo + b_off, b_lbo, b_sbo, b_lt);
                                    umma_tf32(tmem_d, da, dbl, idesc, 1u);
                                    umma_tf32(tmem_d, dal, db, idesc, 1u);
                                }
                            }
                        }
                        if (CG == 2) {
                            umma_commit_pair(&empty_bar[s]);
                            if (kb == num_kb - 1) umma_commit_pair(&tfull_bar[acc]);
                        } else {
                            umma_commit(&empty_bar[s]);                          // smem stage free once these MMAs retire
                            if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete
                        }
                    }
                    __syncwarp();
                    if (lane == 0) stamp(1, kbi, 3);
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                    if (a_tmem && ++sta == A_TMEM_STAGES) { sta = 0; phta ^= 1; }
                }
            }
        }
    } else if (warp < EPI_WARP0) {
        setmaxnreg_dec<72>();
        const int ct = threadIdx.x - 128;  // 0..127
        if (a_tmem) {
            // ===== converters, TMEM form: thread = one row of the A tile (warp w owns TMEM lanes 32*(w%4) .. +31).
            // Read the row's BK raw values from shared memory, split, tcgen05.st the halves into the TMEM ring.
            const int q = warp & 3;
            const int r = q * 32 + lane;
            int sa = 0, sta = 0;
            uint32_t pha = 0, phta = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                int m0, n0, z, num_kb;
                int64_t k_begin;
                tile_coords(tile, m0, n0, z);
                k_range(z, k_begin, num_kb);
                if (z_skipped(z)) continue;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const bool tr = p.trace && blockIdx.x == 0 && tile == tile0 && kb < 16 && ct == 0;
                    const int kbi = (tile - tile0) / tile_step * num_kb + kb;
                    if (ct == 0) stamp(2, kbi, 0);
                    mbar_wait(&raw_bar[sa], pha);
                    if (ct == 0) stamp(2, kbi, 1);
                    if (tr) p.trace[kb * 4 + 1] = clock64();
                    // Order matters (root cause of the round-1 "isolated 32-row groups" corruption): the raw slot used
                    // to be handed back right after the loads were ISSUED - the loads were generic LD.E (the ring pointer
                    // is derived through an integer cast) and ptxas put no scoreboard wait in front of the mbarrier
                    // arrive, so the producer's next TMA could land in the slot while the loads were still in flight.
                    // Now: explicit ld.shared, and the slot is released only after the values have been consumed by the
                    // tcgen05.st below (tcgen05.wait::st retires them).
                    mbar_wait(&ta_empty_bar[sta], phta ^ 1);
                    tc_fence_after();
                    if (ct == 0) stamp(2, kbi, 2);
                    const uint32_t raw = smem_u32(a_ring) + (uint32_t)sa * A_TILE;
                    float v[BK];
                    if (A_MN) {
                        // [mn block q][k][32 mn]: 128-byte k rows, 32-byte chunks XOR-ed with (k & 3)
                        const uint32_t blk = raw + (uint32_t)q * (BK * 128);
#pragma unroll
                        for (int k = 0; k < BK; ++k)
                            v[k] = lds_f32(blk + k * 128 + ((((lane >> 3) ^ (k & 3)) << 5) | ((lane & 7) << 2)));
                    } else {
                        // [row][BK k]: 16-byte chunks XOR-ed with the row index inside the 8-row swizzle atom
                        const uint32_t rowp = raw + (uint32_t)r * (BK * 4);
                        const int sw = BK == 32 ? (r & 7) : ((r >> 1) & 3);
#pragma unroll
                        for (int c = 0; c < BK / 4; ++c)
                            lds_v4(rowp + ((c ^ sw) << 4), v[c * 4 + 0], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
                    }
                    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(A_TMEM_COL0 + sta * 2 * BK);
#pragma unroll
                    for (int h = 0; h < BK / 16; ++h) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float fh, fl;
                            split_tf32(v[h * 16 + j], fh, fl);
                            hi[j] = __float_as_uint(fh);
                            lo[j] = __float_as_uint(fl);
                        }
                        tmem_st16(ta + h * 16, hi);
                        tmem_st16(ta + BK + h * 16, lo);
                    }
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&afree_bar[sa]);   // every value of the slot has been read AND used
                        if (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&ta_full_bar[sta]), 0));   // the leader's
                        else mbar_arrive(&ta_full_bar[sta]);
                    }
                    if (++sa == p.stages_a) { sa = 0; pha ^= 1; }
                    if (ct == 0) stamp(2, kbi, 3);
                    if (tr) p.trace[kb * 4 + 2] = clock64();
                    if (++sta == A_TMEM_STAGES) { sta = 0; phta ^= 1; }
                }
            }
        } else if (p.conv_a) {
            // ===== converters, shared-memory form: split the raw tiles of every stage into tf32 hi / lo halves,
            // element-wise in place (independent of the swizzle), then hand the stage to the MMA warp
            int s = 0;
            uint32_t ph = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                int m0, n0, z, num_kb;
                int64_t k_begin;
                tile_coords(tile, m0, n0, z);
                k_range(z, k_begin, num_kb);
                if (z_skipped(z)) continue;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&raw_bar[s], ph);
                    if (p.trace && blockIdx.x == 0 && tile == tile0 && kb < 16 && ct == 0) p.trace[kb * 4 + 1] = clock64();
                    float4 *hi = reinterpret_cast<float4 *>(tiles + (size_t)s * stage_bytes);
                    float4 *lo = reinterpret_cast<float4 *>(tiles + (size_t)s * stage_bytes + A_TILE);
#pragma unroll
                    for (int i = 0; i < (int)(A_TILE / 16 / 128); ++i) {
                        float4 v = hi[i * 128 + ct], h, l;
                        split_tf32(v.x, h.x, l.x);
                        split_tf32(v.y, h.y, l.y);
                        split_tf32(v.z, h.z, l.z);
                        split_tf32(v.w, h.w, l.w);
                        hi[i * 128 + ct] = h;
                        lo[i * 128 + ct] = l;
                    }
                    if (p.conv_b) {
                        float4 *bh = reinterpret_cast<float4 *>(tiles + (size_t)s * stage_bytes + 2 * A_TILE);
                        float4 *bl = reinterpret_cast<float4 *>(tiles + (size_t)s * stage_bytes + 2 * A_TILE + B_TILE);
#pragma unroll 4
                        for (int i = 0; i < (int)(B_TILE / 16 / 128); ++i) {
                            float4 v = bh[i * 128 + ct], h, l;
                            split_tf32(v.x, h.x, l.x);
                            split_tf32(v.y, h.y, l.y);
                            split_tf32(v.z, h.z, l.z);
                            split_tf32(v.w, h.w, l.w);
                            bh[i * 128 + ct] = h;
                            bl[i * 128 + ct] = l;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> UMMA reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[s]);
                    if (p.trace && blockIdx.x == 0 && tile == tile0 && kb < 16 && ct == 0) p.trace[kb * 4 + 2] = clock64();
                    if (++s == p.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        setmaxnreg_inc<88>();  // pool = 768 x 80 allocated at launch: 128 x 48 + 128 x 72 + 512 x 88 <= 61440
        // ===== epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31
        const int q = warp & 3;
        const int egrp = (warp - EPI_WARP0) >> 2;             // which column slice of the tile this warp drains
        constexpr int EPI_COLS = BLOCK_N / EPI_GROUPS < 32 ? 32 : BLOCK_N / EPI_GROUPS;
        constexpr int NCH = EPI_COLS / 32;
        const bool active = egrp * EPI_COLS < BLOCK_N;
        int titer = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            int m0, n0, z, acc;
            uint32_t aph;
            tile_coords(tile, m0, n0, z);
            if (z_skipped(z)) continue;
            const int tcur = titer++;
            acc_of(tcur, acc, aph);
            const bool second = tiles_dual > 0 && z == 1;
            float *const Dz = second ? p.D2 : p.D + (p.batch > 0 ? p.d_off[z] : (int64_t)z * p.d_z_stride);
            const int64_t Mz = second ? p.M2 : p.M, lddz = second ? p.ldd2 : p.ldd;
            const float *const cf = p.batch > 0 ? p.coef_z[z] : p.coef;
            const float alpha_e = cf ? __ldg(cf) : p.alpha, diag_e = cf ? __ldg(cf + 1) : p.diag;
            const CUtensorMap *const dmap = second ? &tmD2 : &tmD;
            float *const resid_max = p.batch > 0 ? p.resid_z[z] : p.resid_max;
            if (warp == EPI_WARP0 && lane == 0) stamp(3, tcur, 0);
            mbar_wait(&tfull_bar[acc], aph);
            tc_fence_after();
            if (warp == EPI_WARP0 && lane == 0) stamp(3, tcur, 1);
            if (p.trace && blockIdx.x == 0 && warp == EPI_WARP0 && lane == 0 && tcur < 4) p.trace[68 + tcur * 2] = clock64();
            // the warp's whole slice of the tile -> registers, then the accumulator goes back to the MMA warp
            uint32_t vv[NCH][32];
            if (active) {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch)
                    tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) +
                                         (uint32_t)(acc * BLOCK_N + egrp * EPI_COLS + ch * 32), vv[ch]);
                tmem_wait_ld();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));   // the leader's
                else mbar_arrive(&tempty_bar[acc]);
            }
            if (p.trace && blockIdx.x == 0 && warp == EPI_WARP0 && lane == 0 && tcur < 4) p.trace[69 + tcur * 2] = clock64();
            if (warp == EPI_WARP0 && lane == 0) stamp(3, tcur, 2);
            int64_t row = (int64_t)m0 + q * 32 + lane;
            int cbi = 0, cy0 = 0, cx0 = 0;
            if (!D_TRANS && p.conv) {   // tile row r <-> output pixel (y0 + r / 16, x0 + r % 16) of image bi
                conv_coords(m0, cbi, cy0, cx0);
                const int r = q * 32 + lane, yy = cy0 + r / CONV_TW, xx = cx0 + r % CONV_TW;
                row = (yy < p.cv_h && xx < p.cv_w) ? ((int64_t)cbi * p.cv_h + yy) * p.cv_w + xx : Mz;
            }
            uint32_t rmn = 0xffffffffu, rmx = 0xffffffffu;  // this thread's row: f2ord(min), ~f2ord(max)
            float rres = 0.f;                                // max |acc - I| of this thread's part of the tile
            // CG == 2: the tile leaves through shared memory and a TMA store.  (Direct stores of a row-per-thread
            // fragment touch 32 different 128-byte lines per warp instruction: the NHWC epilogue of the inverse
            // rotation took 23 k cycles per tile that way, the channel-major one 8 k - measured with the stamps.)
            const bool staged = CG == 2 && p.tma_store != 0;
            uint8_t *const stg = epi_stage + (size_t)egrp * EPI_STAGE_BYTES;
            const bool issuer = (q == 0) && lane == 0;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int col = egrp * EPI_COLS + ch * 32;
                if (!active || n0 + col >= p.N) break;   // uniform over the group's four warps
                uint32_t (&v)[32] = vv[ch];
                const bool es = warp == EPI_WARP0 && lane == 0;
                if (es) stamp(3, 4 + tcur * 2 + ch, 0);
                if (staged) {
                    if (issuer) tma_store_wait_read();   // the previous store has read the staging tile
                    named_bar_sync(1 + egrp, 128);
                }
                if (es) stamp(3, 4 + tcur * 2 + ch, 1);
                if (D_TRANS) {
                    if (staged) {
                        // staging tile [32 channels][128 pixels]: lanes = consecutive pixels, conflict-free
                        const uint32_t sp = smem_u32(stg) + (uint32_t)(q * 32 + lane) * 4u;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sp + (uint32_t)j * 512u),
                                         "f"(p.alpha * __uint_as_float(v[j]))
                                         : "memory");
                    } else if (row < Mz) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            int64_t n = (int64_t)n0 + col + j;
                            if (n < p.N) Dz[n * lddz + row] = p.alpha * __uint_as_float(v[j]);
                        }
                    }
                    if (p.colrange) {
                        // per-channel min / max of the rotated features (histmatch.py:52-53), folded here so the
                        // matcher does not have to re-read them: warp REDUX over the 32 pixels of every column,
                        // then lane j publishes column j
                        uint32_t mn = 0xffffffffu, mxn = 0xffffffffu;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const uint32_t u = f2ord(p.alpha * __uint_as_float(v[j]));
                            const uint32_t a = __reduce_min_sync(0xffffffffu, row < Mz ? u : 0xffffffffu);
                            const uint32_t b = __reduce_min_sync(0xffffffffu, row < Mz ? ~u : 0xffffffffu);
                            if (lane == j) { mn = a; mxn = b; }
                        }
                        const int64_t n = (int64_t)n0 + col + lane;
                        if (n < p.N) {
                            if (p.colrange_smem) {
                                atomicMin(s_range + 2 * n, mn);
                                atomicMin(s_range + 2 * n + 1, mxn);
                            } else {
                                atomicMin(p.colrange + 2 * n, mn);
                                atomicMin(p.colrange + 2 * n + 1, mxn);
                            }
                        }
                    }
                    if (staged) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        named_bar_sync(1 + egrp, 128);
                        if (issuer) tma_store_2d(dmap, stg, m0, n0 + col);   // x = pixel, y = channel
                    }
                } else if (row < Mz || staged) {
                    float *dp = Dz + row * lddz + n0 + col;
                    const float *bp = p.blend ? p.blend + row * lddz + n0 + col : nullptr;
                    const float *bias = (p.bias && row < Mz) ? p.bias + (row / p.bias_hw) * p.bias_ld + n0 + col : nullptr;
                    // the rotations use none of the optional epilogue terms: their per-element tests (five runtime
                    // conditions x 64 values per thread) cost 9.5 k cycles per 32-column chunk - measured with the stamps
                    const bool extras = resid_max != nullptr || diag_e != 0.f || bias != nullptr || p.relu != 0 ||
                                        p.rowrange != nullptr || alpha_e != 1.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (!extras) break;
                        float o = alpha_e * __uint_as_float(v[j]);
                        if (resid_max && row < Mz && n0 + col + j < p.N) {
                            float d = fabsf(__uint_as_float(v[j]) - (row == (int64_t)n0 + col + j ? 1.f : 0.f));
                            if (!(d == d)) d = INFINITY;
                            rres = fmaxf(rres, d);
                        }
                        if (diag_e != 0.f && row == (int64_t)n0 + col + j) o = __fadd_rn(o, diag_e);
                        if (bias && n0 + col + j < p.N) o = __fadd_rn(o, __ldg(bias + j));
                        if (p.relu) o = o < 0.f ? 0.f : o;
                        v[j] = __float_as_uint(o);
                        if (p.rowrange && n0 + col + j < p.N) {
                            const uint32_t u = f2ord(o);
                            rmn = u < rmn ? u : rmn;
                            rmx = ~u < rmx ? ~u : rmx;
                        }
                    }
                    if (staged) {
                        // staging tile [128 rows][32 columns], 128-byte rows, 16-byte chunks XOR-ed with the row index
                        // (TMA SWIZZLE_128B): a quarter warp's 8 rows hit 8 different bank groups
                        const int r = q * 32 + lane;
                        const uint32_t sp = smem_u32(stg) + (uint32_t)r * 128u;
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sp + (uint32_t)((c4 ^ (r & 7)) << 4)),
                                         "r"(v[c4 * 4 + 0]), "r"(v[c4 * 4 + 1]), "r"(v[c4 * 4 + 2]), "r"(v[c4 * 4 + 3])
                                         : "memory");
                        if (es) stamp(3, 4 + tcur * 2 + ch, 2);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        named_bar_sync(1 + egrp, 128);
                        if (issuer) {
                            if (p.conv) tma_store_4d(dmap, stg, n0 + col, cx0, cy0, cbi);   // box {32 cols, 16 x, 8 y, 1}
                            else tma_store_2d(dmap, stg, n0 + col, m0);                     // x = column, y = row
                        }
                        if (es) stamp(3, 4 + tcur * 2 + ch, 3);
                    } else if (n0 + col + 32 <= p.N) {
                        // 256-bit accesses: every store is one full 32-byte sector of this thread's output row
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            float o[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[j + e]);
                            if (bp) {
                                float c[8];
                                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                             : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]), "=f"(c[4]), "=f"(c[5]),
                                               "=f"(c[6]), "=f"(c[7])
                                             : "l"(bp + j));
#pragma unroll
                                for (int e = 0; e < 8; ++e)
                                    o[e] = __fadd_rn(o[e], __fmul_rn(p.strength, __fsub_rn(c[e], o[e])));
                            }
                            asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dp + j), "f"(o[0]),
                                         "f"(o[1]), "f"(o[2]), "f"(o[3]), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7])
                                         : "memory");
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n0 + col + j < p.N) {
                                float o = __uint_as_float(v[j]);
                                if (bp) o = __fadd_rn(o, __fmul_rn(p.strength, __fsub_rn(__ldg(bp + j), o)));
                                dp[j] = o;
                            }
                    }
                }
            }
            if (!D_TRANS && p.rowrange && row < Mz) {
                atomicMin(p.rowrange + 2 * row, rmn);
                atomicMin(p.rowrange + 2 * row + 1, rmx);
            }
            if (!D_TRANS && resid_max) {
                rres = warp_max(rres);
                if (lane == 0 && rres > 0.f) atomicMax(reinterpret_cast<unsigned int *>(resid_max), __float_as_uint(rres));
            }
            if (warp == EPI_WARP0 && lane == 0) stamp(3, tcur, 3);
        }
        if (CG == 2 && p.tma_store && (warp & 3) == 0 && lane == 0) tma_store_wait_all();   // stores complete before exit
    }
    tc_fence_before();
    __syncthreads();
    if (D_TRANS && p.colrange && p.colrange_smem)
        for (int i = threadIdx.x; i < 2 * (int)p.N; i += NTHREADS)
            if (s_range[i] != 0xffffffffu) atomicMin(p.colrange + i, s_range[i]);
    if (CG == 2) cluster_sync_all();   // neither CTA leaves while its peer may still signal its barriers / read its smem
    if (p.trace && blockIdx.x == 0 && threadIdx.x == 0) p.trace[67] = clock64();
    if (cta_stamps) {
        unsigned long long g;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
        p.trace[128 + 4 * blockIdx.x + 2] = clock64();
        p.trace[128 + 4 * blockIdx.x + 3] = g;
    }
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols)
                         : "memory");
    }
}

// a -> (tf32(a), tf32(a - tf32(a)))   round-to-nearest split, both halves exactly representable in tf32
__global__ void split_tf32_kernel(const float4 *__restrict__ x, float4 *__restrict__ hi, float4 *__restrict__ lo,
                                  int64_t n4) {
    pdl_wait();
    auto split = [](float a, float &h, float &l) { split_tf32(a, h, l); };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = __ldg(x + i), h, l;
        split(v.x, h.x, l.x);
        split(v.y, h.y, l.y);
        split(v.z, h.z, l.z);
        split(v.w, h.w, l.w);
        hi[i] = h;
        lo[i] = l;
    }
}

__global__ void split_fill_kernel(const float4 *__restrict__ x, float4 *__restrict__ hi, float4 *__restrict__ lo,
                                  int64_t n4, uint32_t *__restrict__ fill, int64_t fill_n, uint32_t fill_v) {
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        float4 v = __ldg(x + i), h, l;
        split_tf32(v.x, h.x, l.x);
        split_tf32(v.y, h.y, l.y);
        split_tf32(v.z, h.z, l.z);
        split_tf32(v.w, h.w, l.w);
        hi[i] = h;
        lo[i] = l;
    }
    if (i < fill_n) fill[i] = fill_v;
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
    });
    return fn;
}

// K-major operand: row-major [rows, K]; box {32 k, box_rows}
int make_map_kmajor(CUtensorMap *map, const float *base, int64_t rows, int64_t K, int box_rows, int bk = 32) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE,
                             bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (K-major [%lld, %lld]) failed: %d", (long long)rows, (long long)K, (int)r);
        return OPTEX_ECUDA;
    }
    return OPTEX_OK;
}
// MN-major operand: row-major [K, MN] (MN contiguous, MN % 32 == 0) viewed as {32, K, MN/32}; box {32, 32, box_mn/32}
int make_map_mnmajor(CUtensorMap *map, const float *base, int64_t K, int64_t MN, int box_mn, int64_t ld = 0,
                     int bk = 32) {
    cuuint64_t dims[3] = {32, (cuuint64_t)K, (cuuint64_t)(MN / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)(ld > 0 ? ld : MN) * 4, 128};
    cuuint32_t box[3] = {32, (cuuint32_t)bk, (cuuint32_t)(box_mn / 32)};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (MN-major [%lld, %lld]) failed: %d", (long long)K, (long long)MN, (int)r);
        return OPTEX_ECUDA;
    }
    return OPTEX_OK;
}

// output tile map for the TMA-store epilogue: row-major [rows, cols] with pitch ld; box {box_cols, box_rows}
int make_map_out(CUtensorMap *map, const float *base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                 bool swizzle128) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE,
                             swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (output [%lld, %lld]) failed: %d", (long long)rows, (long long)cols, (int)r);
        return OPTEX_ECUDA;
    }
    return OPTEX_OK;
}

// implicit-GEMM convolution: the padded NHWC input [b, hp, wp, cin] as {cin, wp, hp, b}, box {32, CONV_TW, CONV_TH, 1}
int make_map_conv_in(CUtensorMap *map, const float *base, int b, int hp, int wp, int cin) {
    cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)wp, (cuuint64_t)hp, (cuuint64_t)b};
    cuuint64_t strides[3] = {(cuuint64_t)cin * 4, (cuuint64_t)wp * cin * 4, (cuuint64_t)hp * wp * cin * 4};
    cuuint32_t box[4] = {32, CONV_TW, CONV_TH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (conv input [%d, %d, %d, %d]) failed: %d", b, hp, wp, cin, (int)r);
        return OPTEX_ECUDA;
    }
    return OPTEX_OK;
}
// ... and the NHWC output [b, h, w, ld] (first n channels) as {n, w, h, b}, box {32, CONV_TW, CONV_TH, 1}
int make_map_conv_out(CUtensorMap *map, const float *base, int b, int h, int w, int64_t n, int64_t ld) {
    cuuint64_t dims[4] = {(cuuint64_t)n, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)b};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)w * ld * 4, (cuuint64_t)h * w * ld * 4};
    cuuint32_t box[4] = {32, CONV_TW, CONV_TH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (conv output [%d, %d, %d, %lld]) failed: %d", b, h, w, (long long)n, (int)r);
        return OPTEX_ECUDA;
    }
    return OPTEX_OK;
}

// grow-only device scratch for the hi/lo halves, one buffer per (device, stream, slot): work enqueued on two streams
// never shares scratch (every C-ABI entry takes the stream, and the Python workspace is per stream as well); the slot
// separates the two pipelines of optex_ot_step_host_async when they are given the same stream.
struct Scratch {
    int dev;
    cudaStream_t st;
    int slot;
    void *buf;
    size_t cap;
};
std::vector<Scratch> g_scratch;
std::mutex g_scratch_mu;
thread_local int g_scratch_slot = 0;
struct Presplit {
    const float *src, *hi, *lo;
};
thread_local Presplit g_presplit = {nullptr, nullptr, nullptr};

int scratch(size_t bytes, cudaStream_t st, float **out) {
    int dev = 0;
    OPTEX_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_scratch_mu);
    Scratch *s = nullptr;
    for (auto &e : g_scratch)
        if (e.dev == dev && e.st == st && e.slot == g_scratch_slot) s = &e;
    if (!s) {
        g_scratch.push_back(Scratch{dev, st, g_scratch_slot, nullptr, 0});
        s = &g_scratch.back();
    }
    if (s->cap < bytes) {
        if (s->buf) {
            OPTEX_CUDA(cudaStreamSynchronize(st));  // only this stream's work can still be reading the old buffer
            OPTEX_CUDA(cudaFree(s->buf));
            s->buf = nullptr;
            s->cap = 0;
        }
        OPTEX_CUDA(cudaMalloc(&s->buf, bytes));
        s->cap = bytes;
    }
    *out = (float *)s->buf;
    return OPTEX_OK;
}

int split(const float *x, float *hi, float *lo, int64_t n, cudaStream_t st) {
    int64_t n4 = n / 4;  // callers guarantee n % 4 == 0
    int blocks = (int)((n4 + 255) / 256);
    int cap = sm_count() * 8;
    if (blocks > cap) blocks = cap;
    launch_pdl(split_tf32_kernel, dim3(blocks), dim3(256), 0, st, (const float4 *)x, (float4 *)hi, (float4 *)lo, n4);
    OPTEX_LAUNCH_CHECK("split_tf32_kernel");
    return OPTEX_OK;
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool D_TRANS, int BK, int CG = 1>
int launch(const CUtensorMap &ah, const CUtensorMap &al, const CUtensorMap &bh, const CUtensorMap &bl,
           const CUtensorMap &dm, const CUtensorMap &dm2, Params p, int nz, cudaStream_t st) {
    constexpr int A_TILE = BLOCK_M * BK * 4, B_TILE = (BLOCK_N / CG) * BK * 4;
    size_t smem;
    if (CG != 2) p.tma_store = 0;
    // shared-memory accumulation of the range fold: 2 * N words carved from the tile budget
    const int range_bytes = (D_TRANS && p.colrange && p.N <= 1024) ? (int)((2 * p.N * 4 + 1023) / 1024 * 1024) : 0;
    p.colrange_smem = range_bytes > 0 ? 1 : 0;
    if (p.a_tmem) {
        // main ring: hi + lo tiles of B (CG == 2: this CTA's half of them); the rest of shared memory: raw A tiles
        // (at least 4 in flight) and, with the TMA-store epilogue, one staging tile per epilogue group
        const int epi = (p.tma_store ? EPI_GROUPS * EPI_STAGE_BYTES : 0) + range_bytes;
        int stages = (SMEM_BUDGET - epi - 4 * A_TILE) / (2 * B_TILE);
        if (stages > MAX_STAGES) stages = MAX_STAGES;
        int stages_a = (SMEM_BUDGET - epi - stages * 2 * B_TILE) / A_TILE;
        if (stages_a > MAX_STAGES) stages_a = MAX_STAGES;
        p.stages = stages;
        p.stages_a = stages_a;
        p.range_off = (uint32_t)((size_t)stages * 2 * B_TILE + (size_t)stages_a * A_TILE + (epi - range_bytes));
        smem = (size_t)stages * 2 * B_TILE + (size_t)stages_a * A_TILE + epi + 1024;
    } else {
        const uint32_t stage_bytes = (p.terms == 3 ? 2 : 1) * (A_TILE + B_TILE);
        int stages = (SMEM_BUDGET - range_bytes) / (int)stage_bytes;
        if (stages > MAX_STAGES) stages = MAX_STAGES;
        p.stages = stages;
        p.range_off = (uint32_t)((size_t)stages * stage_bytes);
        smem = (size_t)stages * stage_bytes + range_bytes + 1024;
    }
    auto kern = rotate_gemm_kernel<BLOCK_N, A_MN, B_MN, D_TRANS, BK, CG>;
    static PerDeviceOnce attr_once1;
    OPTEX_TRY(ensure_dyn_smem(attr_once1, kern, (int)(SMEM_BUDGET + 1024)));
    if (!(p.a_tmem && nz == 1 && p.batch == 0)) p.M2 = 0;   // dual problems: TMEM-A form only
    const int64_t num_tiles = (((p.M + BLOCK_M * CG - 1) / (BLOCK_M * CG)) * nz + (p.M2 + BLOCK_M * CG - 1) / (BLOCK_M * CG)) *
                              ((p.N + BLOCK_N - 1) / BLOCK_N);
    const int sms = sm_count();
    if (CG == 1) {
        dim3 grid((unsigned)(num_tiles < sms ? num_tiles : sms));  // persistent: one CTA per SM walks the tile list
        launch_pdl(kern, grid, dim3(NTHREADS), smem, st, ah, al, bh, bl, dm, dm2, p);
    } else {
        // persistent pairs: one cluster of two CTAs per SM pair
        const int64_t pairs = num_tiles < sms / 2 ? num_tiles : sms / 2;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(2 * pairs));
        cfg.blockDim = dim3(NTHREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled() ? 2 : 1;
        OPTEX_CUDA(cudaLaunchKernelEx(&cfg, kern, ah, al, bh, bl, dm, dm2, p));
    }
    OPTEX_LAUNCH_CHECK("rotate_gemm_kernel");
    return OPTEX_OK;
}

template <bool A_MN, bool B_MN, bool D_TRANS>
int launch_n(int block_n, int cg, const CUtensorMap &ah, const CUtensorMap &al, const CUtensorMap &bh,
             const CUtensorMap &bl, const CUtensorMap &dm, const CUtensorMap &dm2, Params p, int nz, cudaStream_t st) {
    // (a BK = 16 / 4-stage variant of the 3xTF32 wide tile exists as a template option; it measured slower: 64 vs 59 us)
    if (block_n == 64) return launch<64, A_MN, B_MN, D_TRANS, 32>(ah, al, bh, bl, dm, dm2, p, nz, st);
    if (block_n == 128) return launch<128, A_MN, B_MN, D_TRANS, 32>(ah, al, bh, bl, dm, dm2, p, nz, st);
    if (cg == 2) return launch<256, A_MN, B_MN, D_TRANS, 32, 2>(ah, al, bh, bl, dm, dm2, p, nz, st);
    return launch<256, A_MN, B_MN, D_TRANS, 32>(ah, al, bh, bl, dm, dm2, p, nz, st);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// widest N tile that still yields about one CTA per SM (small C x C products want many small tiles)
inline int pick_block_n(int64_t M, int64_t N, int nz) {
    const int64_t mt = (M + BLOCK_M - 1) / BLOCK_M;
    const int64_t want = sm_count();
    static const char *force = getenv("OPTEX_FORCE_BN");
    if (force) return atoi(force);
    for (int bn : {256, 128}) {
        if (N <= bn / 2) continue;
        // 128-wide tiles already from 0.7 tiles per SM: one partial wave of them beats 1.4+ waves of 64-wide tiles,
        // which pull twice the A bytes per MMA and sit on the L2 -> SM limit (VGG conv4_x of a 512^2 image: 128 tiles;
        // Encoder(5) 1.92 -> 1.68 ms, Decoder(5) 2.07 -> 1.88 ms)
        const int64_t need = bn == 128 ? (want * 7 + 9) / 10 : want;
        if (mt * ((N + bn - 1) / bn) * nz >= need) return bn;
    }
    if (N > 128 && mt * ((N + 63) / 64) * nz > 4 * want) return 256;
    return N <= 64 ? 64 : (mt * ((N + 127) / 128) * nz >= (want * 7 + 9) / 10 ? 128 : 64);
}

}  // namespace

unsigned long long *g_trace = nullptr;
void gemm_tc_set_trace(unsigned long long *t) { g_trace = t; }

void gemm_tc_set_scratch_slot(int slot) { g_scratch_slot = slot & 3; }

void gemm_tc_set_presplit(const float *src, const float *hi, const float *lo) { g_presplit = {src, hi, lo}; }

// `count` matrices of n floats each: out[i] = [hi_i (n floats) | lo_i (n floats)]
__global__ void split_batch_kernel(const float4 *__restrict__ x, float4 *__restrict__ out, int64_t n4, int count) {
    pdl_wait();
    const int64_t total = n4 * count;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / n4, e = i - m * n4;
        float4 v = __ldg(x + i), h, l;
        split_tf32(v.x, h.x, l.x);
        split_tf32(v.y, h.y, l.y);
        split_tf32(v.z, h.z, l.z);
        split_tf32(v.w, h.w, l.w);
        out[m * 2 * n4 + e] = h;
        out[m * 2 * n4 + n4 + e] = l;
    }
}
int gemm_tc_split_batch(const float *x, float *out, int count, int64_t n, cudaStream_t st) {
    const int64_t n4 = n / 4;
    int64_t blocks = (n4 * count + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    launch_pdl(split_batch_kernel, dim3((unsigned)blocks), dim3(256), 0, st, (const float4 *)x, (float4 *)out, n4, count);
    OPTEX_LAUNCH_CHECK("split_batch_kernel");
    return OPTEX_OK;
}

// hi/lo split of x (n floats, n % 4 == 0) and, in the same launch, fill of a small u32 array (the cdf range slots)
int gemm_tc_split_and_fill(const float *x, float *hi, float *lo, int64_t n, uint32_t *fill, int64_t fill_n,
                           uint32_t fill_v, cudaStream_t st) {
    const int64_t n4 = n / 4;
    int64_t work = n4 > fill_n ? n4 : fill_n;
    int blocks = (int)((work + 255) / 256);
    if (blocks < 1) blocks = 1;
    launch_pdl(split_fill_kernel, dim3(blocks), dim3(256), 0, st, (const float4 *)x, (float4 *)hi, (float4 *)lo, n4, fill,
               fill_n, fill_v);
    OPTEX_LAUNCH_CHECK("split_fill_kernel");
    return OPTEX_OK;
}

// batch mode of gemm_tc (see TcGemm::batch)
static int gemm_tc_batched(const TcGemm &g, cudaStream_t st) {
    const int nb = g.batch;
    if (nb > MAX_BATCH || g.a_mn || g.d_trans || g.blend || g.bias || g.split_k > 1 || g.ldb > 0 || g.b_col0 != 0)
        return OPTEX_ENOTSUP;
    if (g.K % 32 != 0 || g.N % 32 != 0 || g.ldd % 8 != 0 || g.ldd != g.N) return OPTEX_ENOTSUP;
    const float *a0 = g.A_z[0], *b0 = g.B_z[0];
    float *d0 = g.D_z[0];
    for (int i = 1; i < nb; ++i) {
        if (g.A_z[i] < a0) a0 = g.A_z[i];
        if (g.B_z[i] < b0) b0 = g.B_z[i];
        if (g.D_z[i] < d0) d0 = g.D_z[i];
    }
    if (!aligned16(a0) || !aligned16(b0) || (reinterpret_cast<uintptr_t>(d0) & 31) != 0) return OPTEX_ENOTSUP;
    Params p{};
    // A: [rows, K] K-major arena; B: MN-major [k rows, N] or K-major [n rows, K] arena
    const int64_t a_row = g.K, b_row = g.b_mn ? g.N : g.K;
    int64_t a_rows = 0, b_rows = 0;
    for (int i = 0; i < nb; ++i) {
        const int64_t da = g.A_z[i] - a0, db = g.B_z[i] - b0, dd = g.D_z[i] - d0;
        if (da % a_row != 0 || db % b_row != 0 || dd % 8 != 0) return OPTEX_ENOTSUP;
        p.a_off[i] = (int)(da / a_row);
        p.b_off[i] = (int)(db / b_row);
        p.d_off[i] = dd;
        p.resid_z[i] = g.resid_z[i];
        p.skip_z[i] = g.skip_z[i];
        p.coef_z[i] = g.coef_z[i];
        if (p.a_off[i] + g.M > a_rows) a_rows = p.a_off[i] + g.M;
        const int64_t bext = p.b_off[i] + (g.b_mn ? g.K : g.N);
        if (bext > b_rows) b_rows = bext;
    }
    if (a_rows > 0x3fffffffLL || b_rows > 0x3fffffffLL) return OPTEX_ENOTSUP;
    const int bn = pick_block_n(g.M, g.N, nb);
    CUtensorMap ah, bh;
    OPTEX_TRY(make_map_kmajor(&ah, a0, a_rows, g.K, BLOCK_M, 32));
    if (g.b_mn) OPTEX_TRY(make_map_mnmajor(&bh, b0, b_rows, g.N, bn, g.N, 32));
    else OPTEX_TRY(make_map_kmajor(&bh, b0, b_rows, g.K, bn, 32));
    p.D = d0; p.ldd = g.ldd; p.M = g.M; p.N = g.N; p.K = g.K;
    p.terms = g.terms == 3 ? 3 : 1;
    p.alpha = g.alpha; p.diag = g.diag; p.skip = g.skip; p.skip_tol = g.skip_tol;
    // both operands are split inside the kernel (C x C problems: a pre-split pass per launch would dominate)
    p.conv_a = p.terms == 3 ? 1 : 0; p.conv_b = p.conv_a; p.a_tmem = 0;
    p.batch = nb; p.nz = nb;
    p.bias_hw = 1;
    p.a_hint = p.b_hint = L2_EVICT_NORMAL;
    p.trace = nullptr;
    if (!g.b_mn) return launch_n<false, false, false>(bn, 1, ah, ah, bh, bh, ah, ah, p, nb, st);
    return launch_n<false, true, false>(bn, 1, ah, ah, bh, bh, ah, ah, p, nb, st);
}

int gemm_tc(const TcGemm &g, cudaStream_t st) {
    if (g.batch > 0) {
        if (!encode_fn() || g.M < 1 || g.N < 1 || g.K < 1) return OPTEX_ENOTSUP;
        return gemm_tc_batched(g, st);
    }
    if (!encode_fn() || g.M < 1 || g.N < 1 || g.K < 1 || g.M > 0x3fffffffLL || g.N > 0x3fffffffLL || g.K > 0x3fffffffLL)
        return OPTEX_ENOTSUP;
    if (!aligned16(g.A) || !aligned16(g.B) || !aligned16(g.D) || (g.blend && !aligned16(g.blend))) return OPTEX_ENOTSUP;
    // 16-byte global strides for TMA; 32-wide mn blocks for the MN-major 3-D view
    if (g.a_mn ? (g.M % 32 != 0) : (g.K % 4 != 0)) return OPTEX_ENOTSUP;
    if (g.b_mn ? (g.N % 32 != 0) : (g.K % 4 != 0)) return OPTEX_ENOTSUP;
    if (g.d_trans && !(g.b_mn && !g.a_mn)) return OPTEX_ENOTSUP;
    const int64_t ldb = g.ldb > 0 ? g.ldb : g.N;  // B = columns [0, N) of a row-major [K, ldb] matrix
    if ((ldb != g.N || g.b_col0 != 0) && (!g.b_mn || ldb % 4 != 0 || g.b_col0 % 4 != 0 || g.b_col0 + g.N > ldb))
        return OPTEX_ENOTSUP;
    if (g.d_trans && (g.blend || g.bias)) return OPTEX_ENOTSUP;
    if (!g.d_trans && (g.ldd % 8 != 0 || (reinterpret_cast<uintptr_t>(g.D) & 31) != 0 ||
                       (g.blend && (reinterpret_cast<uintptr_t>(g.blend) & 31) != 0)))
        return OPTEX_ENOTSUP;
    int nz = g.split_k > 1 ? g.split_k : 1;
    int64_t k_per_z = 0;
    if (nz > 1) {
        k_per_z = ((g.K + nz - 1) / nz + BLOCK_K - 1) / BLOCK_K * BLOCK_K;
        nz = (int)((g.K + k_per_z - 1) / k_per_z);
    }
    const int bn = pick_block_n(g.M, g.N, nz);
    const int bk = 32;  // must match launch_n()
    const float *ah_p = g.A, *al_p = g.A, *bh_p = g.B, *bl_p = g.B;
    bool conv_b = false;
    if (g.terms == 3) {
        const size_t na = (size_t)g.M * g.K, nb = (size_t)ldb * g.K;
        // A is split inside the kernel (converter warps); only B (the small operand of the rotations: R) is
        // pre-split by an element-wise pass
        (void)na;
        static const char *force_cb = getenv("OPTEX_FORCE_CONVB");
        // few M tiles share B: split it in the kernel - unless B is the big operand (presplit_b), where converting
        // it in shared memory would make the kernel smem-bandwidth bound
        conv_b = force_cb ? atoi(force_cb) != 0 : (g.M <= 4096 && !g.presplit_b);
        if (!conv_b && g_presplit.src == g.B && g.B != nullptr) {
            bh_p = g_presplit.hi;  // the caller split this operand once for several GEMMs (the step's rotation R)
            bl_p = g_presplit.lo;
        } else if (!conv_b) {
            float *buf;
            OPTEX_TRY(scratch(2 * nb * sizeof(float), st, &buf));
            float *b_hi = buf, *b_lo = b_hi + nb;
            OPTEX_TRY(split(g.B, b_hi, b_lo, (int64_t)nb, st));
            bh_p = b_hi; bl_p = b_lo;
        }
    }
    bh_p += g.b_col0;  // column block of a wider B (16-byte aligned: b_col0 % 4 == 0)
    bl_p += g.b_col0;
    // 2-CTA pairs (cta_group::2) for the big 3xTF32 GEMMs: each CTA pulls half of the B tile, which takes the kernel
    // off the L2 -> SM bandwidth limit it sat on (80 KB per k block and SM against 1536 clk of tensor-pipe time)
    static const char *no_atmem = getenv("OPTEX_NO_A_TMEM");
    static const char *cg_env = getenv("OPTEX_CTA_GROUP");
    const bool a_tmem = g.terms == 3 && !conv_b && !(no_atmem && atoi(no_atmem));
    const int cg = (a_tmem && bn == 256 && g.M >= 2 * BLOCK_M && g.N % 64 == 0 && !(cg_env && atoi(cg_env) == 1)) ? 2 : 1;
    CUtensorMap ah, al, bh, bl;
    if (g.a_mn) {
        OPTEX_TRY(make_map_mnmajor(&ah, ah_p, g.K, g.M, BLOCK_M, 0, bk));
        OPTEX_TRY(make_map_mnmajor(&al, al_p, g.K, g.M, BLOCK_M, 0, bk));
    } else {
        OPTEX_TRY(make_map_kmajor(&ah, ah_p, g.M, g.K, BLOCK_M, bk));
        OPTEX_TRY(make_map_kmajor(&al, al_p, g.M, g.K, BLOCK_M, bk));
    }
    if (g.b_mn) {
        OPTEX_TRY(make_map_mnmajor(&bh, bh_p, g.K, g.N, bn / cg, ldb, bk));
        OPTEX_TRY(make_map_mnmajor(&bl, bl_p, g.K, g.N, bn / cg, ldb, bk));
    } else {
        OPTEX_TRY(make_map_kmajor(&bh, bh_p, g.N, g.K, bn / cg, bk));
        OPTEX_TRY(make_map_kmajor(&bl, bl_p, g.N, g.K, bn / cg, bk));
    }
    Params p{};
    p.D = g.D; p.ldd = g.ldd; p.M = g.M; p.N = g.N; p.K = g.K;
    p.blend = g.blend; p.strength = g.strength; p.terms = g.terms == 3 ? 3 : 1;
    p.k_per_z = k_per_z; p.d_z_stride = g.d_z_stride;
    p.bias = g.bias; p.bias_hw = g.bias_hw > 0 ? g.bias_hw : 1; p.bias_ld = g.bias_ld;
    p.diag = g.d_trans ? 0.f : g.diag; p.resid_max = g.d_trans ? nullptr : g.resid_max;
    p.skip_below = g.skip_below; p.skip_tol = g.skip_tol; p.relu = (g.relu && !g.d_trans) ? 1 : 0;
    p.alpha = g.alpha; p.skip = g.skip; p.conv_a = p.terms == 3 ? 1 : 0; p.conv_b = conv_b ? 1 : 0;
    p.coef = g.d_trans ? nullptr : g.coef;
    p.a_tmem = a_tmem ? 1 : 0;
    // Round 1 routed 64-wide multi-tile launches away from the TMEM-A form because of a timing-dependent corruption;
    // the cause was the early release of the raw-A slot in the converters (see there).  OPTEX_ATMEM64=0 restores the
    // old routing for A/B comparisons (scripts/debug_gemm_multi.py).
    {
        static const char *atmem64 = getenv("OPTEX_ATMEM64");
        const int64_t tiles = ((g.M + BLOCK_M - 1) / BLOCK_M) * ((g.N + bn - 1) / bn) * nz;
        if (p.a_tmem && bn == 64 && tiles > sm_count() && atmem64 && atoi(atmem64) == 0) p.a_tmem = 0;
    }
    p.nz = nz;
    p.colrange = g.d_trans ? g.colrange : nullptr;
    p.rowrange = g.d_trans ? nullptr : g.rowrange;
    p.trace = g_trace;
    // a big A operand is streamed exactly once (n_tiles_n CTAs read it at the same time); B is re-read by every tile
    p.a_hint = (g.M > 4096 && g.stream_a) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
    p.b_hint = g.M > 4096 ? L2_EVICT_LAST : L2_EVICT_NORMAL;
    // dual launch (forward rotation of the pastiche AND the style block): second A map in `al`, second output map
    CUtensorMap dm2 = ah;
    const bool dual = g.A2 != nullptr && g.M2 > 0 && a_tmem && nz == 1 && !g.a_mn && aligned16(g.A2) && aligned16(g.D2);
    if (g.A2 != nullptr && !dual) return OPTEX_ENOTSUP;
    if (dual) {
        OPTEX_TRY(make_map_kmajor(&al, g.A2, g.M2, g.K, BLOCK_M, bk));
        p.M2 = g.M2; p.ldd2 = g.ldd2; p.D2 = g.D2;
    }
    // TMA-store epilogue (pairs only): needs 16-byte pitches; the blend reads content with the direct-store code
    CUtensorMap dm = ah;
    static const char *no_tma_store = getenv("OPTEX_NO_TMA_STORE");
    if (cg == 2 && !g.blend && nz == 1 && g.ldd % 4 == 0 && !(no_tma_store && atoi(no_tma_store))) {
        if (g.d_trans) OPTEX_TRY(make_map_out(&dm, g.D, g.N, g.M, g.ldd, BLOCK_M, 32, false));   // D^T [channel][pixel]
        else OPTEX_TRY(make_map_out(&dm, g.D, g.M, g.N, g.ldd, 32, BLOCK_M, true));
        p.tma_store = 1;
        if (dual) {
            if (g.ldd2 % 4 != 0) return OPTEX_ENOTSUP;
            if (g.d_trans) OPTEX_TRY(make_map_out(&dm2, g.D2, g.N, g.M2, g.ldd2, BLOCK_M, 32, false));
            else OPTEX_TRY(make_map_out(&dm2, g.D2, g.M2, g.N, g.ldd2, 32, BLOCK_M, true));
        }
    }
    if (g.d_trans) return launch_n<false, true, true>(bn, cg, ah, al, bh, bl, dm, dm2, p, nz, st);
    if (!g.a_mn && !g.b_mn) return launch_n<false, false, false>(bn, cg, ah, al, bh, bl, dm, dm2, p, nz, st);
    if (!g.a_mn && g.b_mn) return launch_n<false, true, false>(bn, cg, ah, al, bh, bl, dm, dm2, p, nz, st);
    if (g.a_mn && !g.b_mn) return launch_n<true, false, false>(bn, cg, ah, al, bh, bl, dm, dm2, p, nz, st);
    return launch_n<true, true, false>(bn, cg, ah, al, bh, bl, dm, dm2, p, nz, st);
}

// 3 x 3 convolution as an implicit GEMM (Params::conv): out[pixel, co] = sum_k A(pixel, k) W[co, k] (+ bias, ReLU) with
// k = tap * cin + ci and A read straight from the reflection-padded NHWC input.  3xTF32, weights pre-split by the caller.
int gemm_tc_conv(const TcConv &c, cudaStream_t st) {
    if (!encode_fn() || c.b < 1 || c.h < 1 || c.w < 1 || c.cin < 32 || c.cin % 32 != 0 || c.cout < 1) return OPTEX_ENOTSUP;
    if (!aligned16(c.padded) || !aligned16(c.w_hi) || !aligned16(c.w_lo) || c.ldd % 8 != 0 ||
        (reinterpret_cast<uintptr_t>(c.D) & 31) != 0)
        return OPTEX_ENOTSUP;
    const int64_t pixels = (int64_t)c.b * c.h * c.w, K = 9LL * c.cin, N = c.cout;
    if (pixels > 0x3fffffffLL) return OPTEX_ENOTSUP;
    const int bn = pick_block_n(pixels, N, 1);
    static const char *cg_env = getenv("OPTEX_CTA_GROUP");
    const int cg = (bn == 256 && pixels >= 2 * BLOCK_M && N % 64 == 0 && c.h > CONV_TH && !(cg_env && atoi(cg_env) == 1)) ? 2 : 1;
    const int tx = (c.w + CONV_TW - 1) / CONV_TW, ty = (c.h + CONV_TH * cg - 1) / (CONV_TH * cg);
    const int64_t m_tiles = (int64_t)c.b * tx * ty;
    if (m_tiles * BLOCK_M * cg > 0x3fffffffLL) return OPTEX_ENOTSUP;
    CUtensorMap ah, bh, bl, dm;
    OPTEX_TRY(make_map_conv_in(&ah, c.padded, c.b, c.h + 2, c.w + 2, c.cin));
    OPTEX_TRY(make_map_kmajor(&bh, c.w_hi, N, K, bn / cg, 32));
    OPTEX_TRY(make_map_kmajor(&bl, c.w_lo, N, K, bn / cg, 32));
    dm = ah;
    Params p{};
    p.D = c.D; p.ldd = c.ldd; p.M = m_tiles * BLOCK_M * cg; p.N = N; p.K = K;
    p.terms = 3; p.alpha = 1.f; p.conv_a = 1; p.conv_b = 0; p.a_tmem = 1; p.nz = 1;
    p.bias = c.bias; p.bias_hw = 1; p.bias_ld = 0; p.relu = c.relu ? 1 : 0;
    p.conv = 1; p.cv_h = c.h; p.cv_w = c.w; p.cv_cin = c.cin; p.cv_tx = tx; p.cv_ty = ty;
    p.trace = g_trace;
    p.a_hint = L2_EVICT_NORMAL;   // every input pixel is read by 9 taps (and by both N tiles)
    p.b_hint = L2_EVICT_LAST;
    static const char *no_tma_store = getenv("OPTEX_NO_TMA_STORE");
    if (cg == 2 && !(no_tma_store && atoi(no_tma_store))) {
        OPTEX_TRY(make_map_conv_out(&dm, c.D, c.b, c.h, c.w, N, c.ldd));
        p.tma_store = 1;
    }
    return launch_n<false, false, false>(bn, cg, ah, ah, bh, bl, dm, dm, p, 1, st);
}

int gemm_tc_rotate_forward(const float *X, const float *R, float *dst, int64_t n, int c, bool transposed, int terms,
                           cudaStream_t st, int c0, int nc, uint32_t *colrange) {
    // A = X [n, c] K-major; B = R[:, c0:c0+nc] MN-major (needs nc % 32 == 0, c0 % 4 == 0)
    if (nc < 0) nc = c;
    if (c0 % 4 != 0) return OPTEX_ENOTSUP;
    TcGemm g{};
    g.A = X; g.a_mn = false; g.B = R; g.b_col0 = c0; g.b_mn = true; g.ldb = c; g.D = dst;
    g.ldd = transposed ? n : (int64_t)nc;
    g.d_trans = transposed; g.M = n; g.N = nc; g.K = c; g.terms = terms; g.alpha = 1.f; g.colrange = colrange;
    g.stream_a = true;  // the un-rotated block is not touched again in this step
    return gemm_tc(g, st);
}

// the forward rotations of two blocks with the same R in one launch (dst_* channel-major); OPTEX_ENOTSUP when the
// shapes do not take the paired TMEM-A form - the caller then issues two launches
int gemm_tc_rotate_forward2(const float *X1, int64_t n1, const float *X2, int64_t n2, const float *R, float *dst1,
                            float *dst2, int c, int terms, cudaStream_t st, uint32_t *colrange) {
    TcGemm g{};
    g.A = X1; g.a_mn = false; g.B = R; g.b_col0 = 0; g.b_mn = true; g.ldb = c; g.D = dst1; g.ldd = n1;
    g.d_trans = true; g.M = n1; g.N = c; g.K = c; g.terms = terms; g.alpha = 1.f; g.colrange = colrange;
    g.stream_a = true;
    g.A2 = X2; g.M2 = n2; g.D2 = dst2; g.ldd2 = n2;
    return gemm_tc(g, st);
}

int gemm_tc_rotate_inverse(const float *M, bool m_channel_major, const float *R, float *out, int64_t n, int c,
                           const float *content, float strength, int terms, cudaStream_t st) {
    // B(j, k = c) = R[j, c] K-major; A = Mt [c, n] MN-major (needs n % 32 == 0) or M [n, c] K-major
    TcGemm g{};
    g.A = M; g.a_mn = m_channel_major; g.B = R; g.b_mn = false; g.D = out; g.ldd = c;
    g.M = n; g.N = c; g.K = c; g.terms = terms; g.alpha = 1.f; g.blend = content; g.strength = strength;
    g.stream_a = true;  // the matched block is consumed here
    return gemm_tc(g, st);
}

}  // namespace optex
