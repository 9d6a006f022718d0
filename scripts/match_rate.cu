// Throughput of __match_any_sync (MATCH.ANY) vs an 8-step ballot loop on one B200, 16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_match(unsigned *out, int iters) {
    unsigned v = threadIdx.x * 2654435761u + blockIdx.x, acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            unsigned d = (v >> (j * 3)) & 255u;
            acc += __popc(__match_any_sync(0xffffffffu, d));
        }
        v = v * 1664525u + 1013904223u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_ballot(unsigned *out, int iters) {
    unsigned v = threadIdx.x * 2654435761u + blockIdx.x, acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            unsigned d = (v >> (j * 3)) & 255u, peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                unsigned m = __ballot_sync(0xffffffffu, (d >> b) & 1u);
                peers &= ((d >> b) & 1u) ? m : ~m;
            }
            acc += __popc(peers);
        }
        v = v * 1664525u + 1013904223u;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    unsigned *o; cudaMalloc(&o, 148 * 512 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 2000; float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k_match<<<148, 512>>>(o, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("match.any : %.1f cycles per warp-instruction per SM (16 warps)\n", ms * 1e-3 * 1.9e9 / (iters * 8.0 * 16));
        cudaEventRecord(e0); k_ballot<<<148, 512>>>(o, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("8x ballot : %.1f cycles per key-rank per SM (16 warps)\n", ms * 1e-3 * 1.9e9 / (iters * 8.0 * 16));
    }
    return 0;
}
