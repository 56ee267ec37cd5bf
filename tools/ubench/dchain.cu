// Debug micro-benchmark (not part of the product): latency of DEPENDENT FP64 operations on one warp -- DFMA, DADD, DMUL and
// the separately rounded multiply-then-add pair the WDSP recurrences need -- with and without other busy warps on the SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dchain dchain.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND>
__global__ void k(double *out, double a, double b, int iters, long long *cyc, int busy_warps)
{
    double s = threadIdx.x * 1e-3 + 1.0;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 16; u++) {
                if (KIND == 0) s = fma(s, a, b);
                if (KIND == 1) s = __dadd_rn(s, b);
                if (KIND == 2) s = __dmul_rn(s, a);
                if (KIND == 3) s = __dadd_rn(__dmul_rn(s, a), b);
                if (KIND == 4) { const double d = __dsub_rn(b, s); s = __dadd_rn(s, __dmul_rn(d, a)); }
            }
        }
        long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    } else if (warp <= busy_warps) {
        double x0 = s, x1 = s + 1, x2 = s + 2, x3 = s + 3;
        for (int it = 0; it < iters * 16; it++) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); }
        s = x0 + x1 + x2 + x3;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND> void run(const char *name, int ops, double *d, long long *dc)
{
    for (int busy : {0, 4}) {
        const int iters = 2000;
        k<KIND><<<148, 192>>>(d, 0.9999999, 1e-9, iters, dc, busy);
        cudaDeviceSynchronize();
        k<KIND><<<148, 192>>>(d, 0.9999999, 1e-9, iters, dc, busy);
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s busy warps %d: %.2f cycles per dependent op (%.1f per step)\n", name, busy, c / (iters * 16.0 * ops), c / (iters * 16.0));
    }
}
int main()
{
    double *d; long long *dc;
    cudaMalloc(&d, 148 * 192 * 8); cudaMalloc(&dc, 8);
    run<0>("DFMA chain", 1, d, dc);
    run<1>("DADD chain", 1, d, dc);
    run<2>("DMUL chain", 1, d, dc);
    run<3>("DMUL -> DADD (s = s*a + b)", 2, d, dc);
    run<4>("DADD -> DMUL -> DADD (agc step)", 3, d, dc);
    return 0;
}
