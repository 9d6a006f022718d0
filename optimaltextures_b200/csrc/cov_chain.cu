// The C x C part of the covariance hist modes pca / sym (histmatch.py:29-42) as ONE cooperative kernel - the coupled
// Newton-Schulz chain(s), the closing products and the bias that cov_match.cu issues as ~40 launches per OT iteration
// (2 dependent tensor-core GEMM launches of ~11 us per Newton-Schulz iteration, of which ~8 us is the fixed cost of a
// launch at these sizes, plus the switched-off tail of the fixed-length launch list).
//
// Every step of the chain is a set of independent products C = alpha A B (+ diag I) of c x c fp32 matrices that live
// in L2; a unit of work is one 32 x 32 output tile, computed by a whole CTA - 256 threads = 4 k-groups x 64 threads
// with 4 x 4 register tiles, operands streamed through a 3-stage cp.async ring (L1 bypassed: other SMs wrote them),
// the four partial tiles folded in shared memory in a fixed order - and the units of a step are dealt round-robin to
// the CTAs of the grid; steps are separated by a grid-wide barrier (a release / acquire counter in L2: ~1.5 us).  The
// residual max |Z Y - I| comes out of the T-step's epilogue (atomicMax on the float's bits), every thread of every CTA
// reads it behind the barrier and takes the same branch: the loop really ends where the iteration converges, and a
// chain that stops improving at its rounding floor stops too.  fp32 FFMA (the launch path multiplies in 3xTF32).
// Used for 64 < c <= 384, c % 32 == 0 (the PCA'd layer widths of a synthesis after OptimalTexture's padding); above
// that the FFMA tiles lose to the tensor cores and cov_match.cu's launch chain stays.
#include "common.cuh"

namespace optex {
namespace {

constexpr int CT = 256;     // threads per CTA
constexpr int TILE = 32;    // output tile edge
constexpr int KC = 64;      // k per pipeline stage
constexpr int STAGES = 3;
constexpr int LDA = KC + 4;  // row stride of an A stage [TILE][KC] (floats; 16-byte aligned rows)
constexpr int NS_CAP_C = 24;
constexpr float NS_TOL_C = 3e-4f;
constexpr int MAX_GRID = 160;

struct Prod {
    const float *A, *B;
    float *C;
    float alpha, diag;
    unsigned *resid;  // optional: max |A B - I| (bits of a non-negative float)
    int sym;          // the result is symmetric (products of commuting symmetric matrices): only the tiles on and
                      // above the diagonal are computed, the others are their mirror images
};

__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        unsigned seen;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while ((int)(seen - target) < 0);
    }
    __syncthreads();
}

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// one 32 x 32 tile of C = alpha A B + diag I at (i0, j0); sA: [STAGES][TILE][LDA], sB: [STAGES][KC][TILE]
__device__ void tile_product(const Prod &p, int c, int i0, int j0, float *sA, float *sB) {
    const bool mirror = p.sym && i0 != j0, diag_tile = p.sym && i0 == j0;
    const int tid = threadIdx.x;
    const int g = tid >> 6, t = tid & 63, ty = t >> 3, tx = t & 7;   // k-group, 8 x 8 threads of 4 x 4 outputs
    const int nchunks = (c + KC - 1) / KC;
    auto load = [&](int ch) {
        const int k0 = ch * KC;
        const int kv = c - k0 < KC ? c - k0 : KC;          // valid k of this chunk (c % 32 == 0: 32 or 64)
        float *a = sA + (ch % STAGES) * TILE * LDA, *b = sB + (ch % STAGES) * KC * TILE;
        // A chunk: rows i0 .. i0 + 31, k0 .. k0 + kv: kv / 4 float4 per row
        const int qa = kv >> 2;
        for (int i = tid; i < TILE * qa; i += CT) {
            const int r = i / qa, q = i - r * qa;
            cp_async16(a + r * LDA + 4 * q, p.A + (int64_t)(i0 + r) * c + k0 + 4 * q);
        }
        // B chunk: rows k0 .. k0 + kv, columns j0 .. j0 + 31: 8 float4 per row
        for (int i = tid; i < kv * (TILE / 4); i += CT) {
            const int r = i >> 3, q = i & 7;
            cp_async16(b + r * TILE + 4 * q, p.B + (int64_t)(k0 + r) * c + j0 + 4 * q);
        }
    };
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    load(0);
    cp_async_commit();
    if (nchunks > 1) load(1);
    cp_async_commit();
    for (int ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<1>();   // chunk ch has landed (one younger group may still be in flight)
        __syncthreads();      // ... for every thread, and chunk ch - 1's buffer is free for the next load
        if (ch + 2 < nchunks) load(ch + 2);
        cp_async_commit();
        const int kv = c - ch * KC < KC ? c - ch * KC : KC;
        const int kq = kv >> 2;                              // this group's quarter of the chunk: kq consecutive k
        const float *a = sA + (ch % STAGES) * TILE * LDA + g * kq, *b = sB + (ch % STAGES) * KC * TILE + g * kq * TILE;
        auto fma4 = [&](int k) {
            float av[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 v = *reinterpret_cast<const float4 *>(a + (4 * ty + i) * LDA + k);
                av[i][0] = v.x;
                av[i][1] = v.y;
                av[i][2] = v.z;
                av[i][3] = v.w;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4 v = *reinterpret_cast<const float4 *>(b + (k + kk) * TILE + 4 * tx);
                const float bv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i][kk], bv[j], acc[i][j]);
            }
        };
        if (kq == KC / 4) {   // the full chunk: a compile-time trip count lets the loads of step k + 4 pass the FMAs of k
#pragma unroll
            for (int k = 0; k < KC / 4; k += 4) fma4(k);
        } else {
            for (int k = 0; k < kq; k += 4) fma4(k);
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // fold the four k-groups in a fixed order through shared memory (the A stages are free: 4 x 32 x 32 floats)
    float *red = sA;
    static_assert(4 * TILE * TILE <= STAGES * TILE * LDA, "reduction buffer");
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) red[g * TILE * TILE + (4 * ty + i) * TILE + 4 * tx + j] = acc[i][j];
    __syncthreads();
    // thread -> row tid / 8, columns 4 (tid % 8) .. + 3: a coalesced 128-byte row per 8 threads
    const int r = tid >> 3, q = (tid & 7) * 4;
    float d = 0.f;
    float4 o;
    float *ov = reinterpret_cast<float *>(&o);
    float *fin = sA + 4 * TILE * TILE;   // [TILE][TILE + 1]: the finished tile, for the mirrored / symmetrised store
    static_assert(4 * TILE * TILE + TILE * (TILE + 1) <= STAGES * TILE * LDA, "finished-tile buffer");
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int idx = r * TILE + q + j;
        const float m = (red[idx] + red[TILE * TILE + idx]) + (red[2 * TILE * TILE + idx] + red[3 * TILE * TILE + idx]);
        const float eye = (i0 + r == j0 + q + j) ? 1.f : 0.f;
        if (p.resid) {
            float e = fabsf(m - eye);
            if (!(e == e)) e = INFINITY;
            d = fmaxf(d, e);
        }
        ov[j] = fmaf(p.alpha, m, p.diag * eye);
        if (p.sym) fin[r * (TILE + 1) + q + j] = ov[j];
    }
    if (p.sym) {
        __syncthreads();
        float4 t;
        float *tv = reinterpret_cast<float *>(&t);
#pragma unroll
        for (int j = 0; j < 4; ++j) tv[j] = fin[(q + j) * (TILE + 1) + r];   // row r of the transposed tile
        if (diag_tile) {   // a diagonal tile is made exactly symmetric itself
#pragma unroll
            for (int j = 0; j < 4; ++j) ov[j] = 0.5f * (ov[j] + tv[j]);
        } else if (mirror) {
            __stcg(reinterpret_cast<float4 *>(p.C + (int64_t)(j0 + r) * c + i0 + q), t);
        }
    }
    __stcg(reinterpret_cast<float4 *>(p.C + (int64_t)(i0 + r) * c + j0 + q), o);
    if (p.resid) {
        d = warp_max(d);
        if ((tid & 31) == 0 && d > 0.f) atomicMax(p.resid, __float_as_uint(d));
    }
    __syncthreads();  // the stages are reused by the CTA's next unit
}

// the units of a step, dealt round-robin to the CTAs
__device__ void run_step(const Prod *prods, int np, int c, float *sA, float *sB) {
    const int tc = c / TILE, full = tc * tc, tri = tc * (tc + 1) / 2;
    int total = 0;
    for (int i = 0; i < np; ++i) total += prods[i].sym ? tri : full;
    for (int u = blockIdx.x; u < total; u += gridDim.x) {
        int pi = 0, tt = u;
        while (tt >= (prods[pi].sym ? tri : full)) {
            tt -= prods[pi].sym ? tri : full;
            ++pi;
        }
        int ti, tj;
        if (prods[pi].sym) {   // tt-th tile of the upper triangle, row by row
            ti = 0;
            while (tt >= tc - ti) {
                tt -= tc - ti;
                ++ti;
            }
            tj = ti + tt;
        } else {
            ti = tt / tc;
            tj = tt - ti * tc;
        }
        tile_product(prods[pi], c, ti * TILE, tj * TILE, sA, sB);
    }
}

struct ChainBuf {
    const float *A;      // SPD input
    float *Y[2], *Z[2], *T;
    float lmin;
};

struct CoopParams {
    int c, mode, style_state, b_p, b_s;
    float eps;
    const float *sig_t, *sig_s;
    ChainBuf a, b;             // a: pastiche chain; b: style chain (pca) / second chain (sym)
    float *keep;               // pca: Sig_s^(1/2) kept for the following iterations of a loop
    float *f, *aux, *w;        // sym scratch
    float *G, *bias;
    const float *mu_p, *mu_s;
    float *norm_part;          // [2 chains][sum of squares, largest row sum][MAX_GRID]
    unsigned *resid;           // [2][NS_CAP_C], zeroed before the launch
    unsigned *bar;             // zeroed before the launch
    long long *stamps;         // debug (optex_debug_chain_stamps): clock64 of CTA 0 behind every grid barrier
    int n_stamps;
};

__device__ int g_stamp_idx;
__device__ __forceinline__ void stamp(const CoopParams &p) {
    if (p.stamps && blockIdx.x == 0 && threadIdx.x == 0) {
        const int i = g_stamp_idx++;
        if (i < p.n_stamps) p.stamps[i] = clock64();
    }
}

// The scale of the iteration: s >= |A|_2 so that A / s has its spectrum in (0, 1].  Two bounds, the smaller wins:
// the Frobenius norm and the largest absolute row sum (|A|_inf >= |A|_2; for the near-diagonal covariances of
// PCA-projected features it is 1.0-1.5 lambda_max where |A|_F is 1.5-4, which is worth about one scaled iteration).
// Identical bits in every thread: per-CTA partials (norm_partial), a grid barrier by the caller, a fixed-order
// fold (norm2_total returns s^2).
__device__ void norm_partial(const float *A, int c, float *part, float *part_inf, float *sred) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n = (int64_t)c * c;
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * CT + tid; i < n; i += (int64_t)gridDim.x * CT) {
        const float v = __ldcg(A + i);
        acc = fmaf(v, v, acc);
    }
    acc = warp_sum(acc);
    float rmax = 0.f;
    for (int r = blockIdx.x * (CT / 32) + warp; r < c; r += gridDim.x * (CT / 32)) {   // a warp per row
        float s = 0.f;
        for (int j = lane; j < c; j += 32) s += fabsf(__ldcg(A + (int64_t)r * c + j));
        rmax = fmaxf(rmax, warp_sum(s));
    }
    if (lane == 0) {
        sred[warp] = acc;
        sred[CT / 32 + warp] = rmax;
    }
    __syncthreads();
    if (tid == 0) {
        float s = 0.f, m = 0.f;
        for (int i = 0; i < CT / 32; ++i) {
            s += sred[i];
            m = fmaxf(m, sred[CT / 32 + i]);
        }
        __stcg(part + blockIdx.x, s);
        __stcg(part_inf + blockIdx.x, m);
    }
    __syncthreads();
}
__device__ float norm2_total(const float *part, const float *part_inf, float *sred) {
    const int tid = threadIdx.x;
    if (tid < 32) {
        float s = 0.f, m = 0.f;
        for (int i = tid; i < (int)gridDim.x; i += 32) {
            s += __ldcg(part + i);
            m = fmaxf(m, __ldcg(part_inf + i));
        }
        s = warp_sum(s);   // xor butterfly: the same association in every CTA
        m = warp_max(m);
        if (tid == 0) sred[0] = (m > 0.f && m * m < s) ? m * m : s;
    }
    __syncthreads();
    const float s = sred[0];
    __syncthreads();
    return s;
}

// The coupled Newton-Schulz iteration for `n` (1 or 2) chains in lockstep.  Returns, per chain, the index of the
// buffer pair holding the result (cur) and |A|_F^2 (norm2); Y -> (A / |A|_F)^(1/2), Z -> (A / |A|_F)^(-1/2).
__device__ void run_chains(const CoopParams &p, const ChainBuf *ch, int n, int c, float *norm_part, unsigned *resid,
                           unsigned *bar, unsigned &target, float *sA, float *sB, float *sred, int *cur_out,
                           float *norm2_out) {
    const int64_t cc = (int64_t)c * c;
    float norm2[2], l[2], prev[2];
    int cur[2];
    bool live[2], prev_plain[2] = {false, false}, plain[2] = {false, false};
    for (int k = 0; k < n; ++k) norm_partial(ch[k].A, c, norm_part + 2 * k * MAX_GRID, norm_part + (2 * k + 1) * MAX_GRID, sred);
    grid_barrier(bar, target);
    stamp(p);
    for (int k = 0; k < n; ++k) {
        norm2[k] = norm2_total(norm_part + 2 * k * MAX_GRID, norm_part + (2 * k + 1) * MAX_GRID, sred);
        cur[k] = 0;
        live[k] = true;
        prev[k] = 1.f;
        const float inv = rsqrtf(norm2[k]);
        l[k] = ch[k].lmin > 0.f && norm2[k] > 0.f ? fminf(0.9f * ch[k].lmin * inv, 1.f) : 1.f;
        for (int64_t i = (int64_t)blockIdx.x * CT + threadIdx.x; i < cc; i += (int64_t)gridDim.x * CT) {
            __stcg(ch[k].Y[0] + i, __ldcg(ch[k].A + i) * inv);
            __stcg(ch[k].Z[0] + i, (i / c == i % c) ? 1.f : 0.f);
        }
    }
    grid_barrier(bar, target);
    stamp(p);
    for (int it = 0; it < NS_CAP_C; ++it) {
        Prod pr[4];
        int np = 0;
        for (int k = 0; k < n; ++k) {
            if (!live[k]) continue;
            // the scaled step (cov_match.cu ns_prepare_kernel's schedule, one step per iteration, fp32)
            const float rho = l[k] < 0.8f ? 3.f / (1.f + sqrtf(l[k]) + l[k]) : 1.f;
            const float sr = sqrtf(rho), rl = rho * l[k];
            l[k] = fminf(0.97f * rl * (3.f - rl) * (3.f - rl) * 0.25f, 1.f);
            plain[k] = rho == 1.f;
            pr[np++] = Prod{ch[k].Z[cur[k]], ch[k].Y[cur[k]], ch[k].T, -0.5f * rho * sr, 1.5f * sr,
                            resid + k * NS_CAP_C + it, 1};
        }
        if (np == 0) break;
        run_step(pr, np, c, sA, sB);   // T = a I + b Z Y, residual
        grid_barrier(bar, target);
        stamp(p);
        np = 0;
        bool stop[2] = {false, false};
        for (int k = 0; k < n; ++k) {
            if (!live[k]) continue;
            const float r = __uint_as_float(__ldcg(resid + k * NS_CAP_C + it));
            // under plain steps the residual falls monotonically down to the problem's rounding floor: an iteration that
            // no longer improves is not applied (a scaled step may legitimately overshoot once: cov_small.cu)
            if (prev_plain[k] && prev[k] < 0.1f && !(r < prev[k])) {
                live[k] = false;
                continue;
            }
            prev[k] = r;
            prev_plain[k] = plain[k];
            pr[np++] = Prod{ch[k].Y[cur[k]], ch[k].T, ch[k].Y[cur[k] ^ 1], 1.f, 0.f, nullptr, 1};
            pr[np++] = Prod{ch[k].T, ch[k].Z[cur[k]], ch[k].Z[cur[k] ^ 1], 1.f, 0.f, nullptr, 1};
            stop[k] = r < NS_TOL_C;   // this iteration's update is applied, the next is not needed
        }
        if (np == 0) break;
        run_step(pr, np, c, sA, sB);
        grid_barrier(bar, target);
        stamp(p);
        for (int k = 0; k < n; ++k) {
            if (!live[k]) continue;
            cur[k] ^= 1;
            if (stop[k]) live[k] = false;
        }
    }
    for (int k = 0; k < n; ++k) {
        cur_out[k] = cur[k];
        norm2_out[k] = norm2[k];
    }
}

__global__ void __launch_bounds__(CT) ns_coop_kernel(CoopParams p) {
    extern __shared__ __align__(16) float sm[];
    float *sA = sm, *sB = sm + STAGES * TILE * LDA;
    __shared__ float sred[2 * CT / 32];
    unsigned target = 0;
    const int c = p.c, tid = threadIdx.x;
    if (p.stamps && blockIdx.x == 0 && tid == 0) {
        g_stamp_idx = 0;
        stamp(p);
    }
    const int64_t cc = (int64_t)c * c;
    int cur[2];
    float n2[2];
    const float *Tfin = p.G;
    if (p.mode == OPTEX_MODE_PCA) {
        // T = Sig_s^(1/2) Sig_t^(-1/2)                                  histmatch.py:29-34
        const ChainBuf chains[2] = {p.a, p.b};
        const int n = p.style_state == 0 ? 2 : 1;
        run_chains(p, chains, n, c, p.norm_part, p.resid, p.bar, target, sA, sB, sred, cur, n2);
        const float rs_a = sqrtf(sqrtf(n2[0]));
        Prod pr;
        if (n == 2) {
            const float rs_b = sqrtf(sqrtf(n2[1]));
            const float *yb = p.b.Y[cur[1]];
            for (int64_t i = (int64_t)blockIdx.x * CT + tid; i < cc; i += (int64_t)gridDim.x * CT)
                __stcg(p.keep + i, __ldcg(yb + i) * rs_b);
            pr = Prod{yb, p.a.Z[cur[0]], p.G, rs_b / rs_a, 0.f, nullptr, 0};
        } else {
            pr = Prod{p.keep, p.a.Z[cur[0]], p.G, 1.f / rs_a, 0.f, nullptr, 0};
        }
        run_step(&pr, 1, c, sA, sB);
        grid_barrier(p.bar, target);
        stamp(p);
    } else {
        // sym: T = Qt^-1 (Qt Sig_s Qt)^(1/2) Qt^-1,  Qt = Sig_t^(1/2)    histmatch.py:36-42
        run_chains(p, &p.a, 1, c, p.norm_part, p.resid, p.bar, target, sA, sB, sred, cur, n2);
        const float rs_a = sqrtf(sqrtf(n2[0]));
        const float *ya = p.a.Y[cur[0]], *za = p.a.Z[cur[0]];
        Prod pr{p.sig_s, ya, p.f, rs_a, 0.f, nullptr, 0};        // f = Sig_s Qt
        run_step(&pr, 1, c, sA, sB);
        grid_barrier(p.bar, target);
        pr = Prod{ya, p.f, p.aux, rs_a, 0.f, nullptr, 1};        // aux = Qt Sig_s Qt (symmetric)
        run_step(&pr, 1, c, sA, sB);
        grid_barrier(p.bar, target);
        int cur2[2];
        float n22[2];
        ChainBuf second = p.b;
        second.A = p.aux;
        second.lmin = p.eps * p.eps;   // Qt Sig_s Qt >= lambda_min(Qt)^2 lambda_min(Sig_s) >= eps^2
        run_chains(p, &second, 1, c, p.norm_part + 2 * MAX_GRID, p.resid + NS_CAP_C, p.bar, target, sA, sB, sred, cur2, n22);
        const float rs_c = sqrtf(sqrtf(n22[0]));
        pr = Prod{second.Y[cur2[0]], za, p.w, rs_c / rs_a, 0.f, nullptr, 0};   // w = (Qt Sig_s Qt)^(1/2) Qt^-1
        run_step(&pr, 1, c, sA, sB);
        grid_barrier(p.bar, target);
        pr = Prod{za, p.w, p.G, 1.f / rs_a, 0.f, nullptr, 1};               // T = Qt^-1 w (symmetric)
        run_step(&pr, 1, c, sA, sB);
        grid_barrier(p.bar, target);
    }
    // bias[b][j] = mu_s[bs(b)][j] - sum_k T[j][k] mu_p[b][k]          (means folded through the map)
    const int lane = tid & 31, gw = blockIdx.x * (CT / 32) + (tid >> 5), nw = gridDim.x * (CT / 32);
    for (int job = gw; job < p.b_p * c; job += nw) {
        const int b = job / c, j = job - b * c;
        float acc = 0.f;
        for (int k = lane; k < c; k += 32) acc = fmaf(__ldcg(Tfin + (int64_t)j * c + k), p.mu_p[b * c + k], acc);
        acc = warp_sum(acc);
        if (lane == 0) p.bias[job] = p.mu_s[(p.b_s == 1 ? 0 : b) * c + j] - acc;
    }
    stamp(p);
}

long long *g_chain_stamps = nullptr;
int g_chain_n_stamps = 0;

bool coop_enabled() {
    static const bool v = [] {
        const char *e = getenv("OPTEX_COV_COOP");
        return !(e && atoi(e) == 0);
    }();
    return v;
}

}  // namespace

bool cov_coop_supported(int c, int mode) {
    return coop_enabled() && (mode == OPTEX_MODE_PCA || mode == OPTEX_MODE_SYM) && c > 64 && c <= 384 && c % 32 == 0;
}

size_t cov_coop_scratch_floats() { return 4 * MAX_GRID + 2 * NS_CAP_C + 64; }

// m: the 19 c x c matrices of cov_match.cu's workspace (roles as there); scratch: cov_coop_scratch_floats() floats.
// Expects sig_t = m[0] and sig_s = m[1] (cov + eps I), mu_p / mu_s; writes T into m[3] and the bias.
int cov_coop_chain(float *const *m, int c, int mode, float eps, int style_reuse, const float *mu_p, const float *mu_s,
                   int b_p, int b_s, float *bias, float *scratch, cudaStream_t st) {
    CoopParams p{};
    p.c = c;
    p.mode = mode;
    p.style_state = style_reuse ? 1 : 0;
    p.b_p = b_p;
    p.b_s = b_s;
    p.eps = eps;
    p.sig_t = m[0];
    p.sig_s = m[1];
    p.a = ChainBuf{m[0], {m[5], m[9]}, {m[6], m[10]}, m[8], eps};
    p.b = ChainBuf{m[1], {m[12], m[17]}, {m[15], m[18]}, m[16], eps};
    p.keep = m[11];   // cov_match.cu's Y2: the style square root a loop reuses
    p.f = m[2];
    p.aux = m[13];
    p.w = m[7];
    p.G = m[3];
    p.bias = bias;
    p.mu_p = mu_p;
    p.mu_s = mu_s;
    p.norm_part = scratch;
    p.resid = reinterpret_cast<unsigned *>(scratch + 4 * MAX_GRID);
    p.bar = reinterpret_cast<unsigned *>(scratch + 4 * MAX_GRID + 2 * NS_CAP_C);
    p.stamps = g_chain_stamps;
    p.n_stamps = g_chain_n_stamps;
    OPTEX_CUDA(cudaMemsetAsync(p.resid, 0, sizeof(unsigned) * (2 * NS_CAP_C + 8), st));
    const int tc = c / TILE, tiles = tc * tc, tri = tc * (tc + 1) / 2;
    // the widest step: the update step (2 symmetric products per chain); the closing product has tc^2 tiles
    int grid = 2 * tri * ((mode == OPTEX_MODE_PCA && !style_reuse) ? 2 : 1);
    if (grid < tiles) grid = tiles;
    const int sms = sm_count();
    if (grid > sms) {
        // whole rounds of the common step (one chain: 2 tri tiles in the update step)
        const int common = 2 * tri > tiles ? 2 * tri : tiles;
        const int rounds = (common + sms - 1) / sms;
        grid = (common + rounds - 1) / rounds;
    }
    if (grid > MAX_GRID) grid = MAX_GRID;
    const size_t smem = (size_t)(STAGES * TILE * LDA + STAGES * KC * TILE) * sizeof(float);
    static bool attr_done[64] = {};
    int dev = 0;
    OPTEX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        OPTEX_CUDA(cudaFuncSetAttribute(ns_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    void *args[] = {&p};
    // cooperative: the CTAs spin on each other in grid_barrier and must be co-resident
    OPTEX_CUDA(cudaLaunchCooperativeKernel((const void *)ns_coop_kernel, dim3(grid), dim3(CT), args, smem, st));
    count_launch();
    return OPTEX_OK;
}

}  // namespace optex

// Debug hook (scripts/chain_stamps.py): a device buffer of n int64 that CTA 0 of the cooperative chain kernel fills
// with clock64 behind every grid barrier of its next launches (kernel entry first, kernel end last); NULL: off.
extern "C" int optex_debug_chain_stamps(long long *device_buffer, int n) {
    optex::g_chain_stamps = device_buffer;
    optex::g_chain_n_stamps = device_buffer ? n : 0;
    return OPTEX_OK;
}
