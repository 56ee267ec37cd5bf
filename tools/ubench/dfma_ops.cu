// Debug micro-benchmark (not part of the product): DFMA issue rate with the coefficient in a register, in the constant
// bank / uniform register, and with the multiplicand changing every instruction (the half-band inner loop's shape).
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double cc[16] = {1.0000001, 0.9999999, 1.0000002, 0.9999998, 1.0000003, 0.9999997, 1.0000004, 0.9999996,
                              1.0000005, 0.9999995, 1.0000006, 0.9999994, 1.0000007, 0.9999993, 1.0000008, 0.9999992};
template <int MODE>
__global__ void k(double *out, const double *in, int iters, long long *cyc)
{
    double acc[16], e[8], cr[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { acc[i] = threadIdx.x + i; cr[i] = in[i]; }
#pragma unroll
    for (int i = 0; i < 8; i++) e[i] = in[16 + i + threadIdx.x];
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int r = 0; r < 16; r++) {
                if (MODE == 0) acc[r] = fma(e[j], cr[(r + j) & 15], acc[r]);        // coefficient in a register
                if (MODE == 1) acc[r] = fma(e[j], cc[(r + j) & 15], acc[r]);        // coefficient from the constant bank
                if (MODE == 2) acc[r] = fma(acc[r], cr[0], cr[1]);                  // chain through the multiplicand
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char *name, double *d, double *in, long long *dc)
{
    for (int threads : {128, 256, 512}) {
        k<MODE><<<148, threads>>>(d, in, 2000, dc); cudaDeviceSynchronize();
        k<MODE><<<148, threads>>>(d, in, 2000, dc);
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        printf("%-34s warps/sub-partition %d: %.3f cycles per DFMA per sub-partition\n", name, threads / 128, c / (2000.0 * 128 * (threads / 128)));
    }
}
int main()
{
    double *d, *in; long long *dc;
    cudaMalloc(&d, 148 * 512 * 8); cudaMalloc(&in, 1024 * 8); cudaMemset(in, 0, 1024 * 8); cudaMalloc(&dc, 8);
    run<0>("e[j] * reg coef + acc[r]", d, in, dc);
    run<1>("e[j] * const coef + acc[r]", d, in, dc);
    run<2>("acc[r] * reg + reg", d, in, dc);
    return 0;
}
