// FP64 FMA throughput of one B200 (vector pipe): independent DFMA chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, double a, double b) {
    double x[ILP];
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double *o; cudaMalloc(&o, 148 * 8 * 256 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 1; warps <= 8; warps *= 2) {
        int iters = 20000;
        k<8><<<148 * 4, warps * 32>>>(o, 100, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        k<8><<<148 * 4, warps * 32>>>(o, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 148 * 4 * warps * 32 * 8.0 * iters;
        printf("4 CTAs/SM x %d warps, ILP 8: %.2f TFLOP/s fp64\n", warps, fl / ms / 1e9);
    }
    return 0;
}
