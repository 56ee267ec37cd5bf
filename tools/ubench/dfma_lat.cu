// Debug micro-benchmark (not part of the product): FP64 FMA issue rate of one SM sub-partition as a function of the
// number of independent accumulator chains per thread and of resident warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_lat dfma_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, double a, double b, int iters, long long *cyc)
{
    double acc[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < CH; i++) acc[i] = fma(acc[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH> void run(int threads, double *d, long long *dc)
{
    const int iters = 2000;
    k<CH><<<148, threads>>>(d, 1.0000001, 1e-9, iters, dc);
    cudaDeviceSynchronize();
    k<CH><<<148, threads>>>(d, 1.0000001, 1e-9, iters, dc);
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    const double per_warp_dfma = (double)iters * 8 * CH;
    const int warps_per_sub = (threads / 32 + 3) / 4;
    printf("chains %2d warps/SM %2d: %.2f cycles per DFMA per warp, %.2f cycles per DFMA per sub-partition\n", CH, threads / 32,
           c / per_warp_dfma, c / (per_warp_dfma * warps_per_sub));
}
int main()
{
    double *d; long long *dc;
    cudaMalloc(&d, 148 * 1024 * 8); cudaMalloc(&dc, 8);
    for (int threads : {32, 128, 256, 512}) {
        run<1>(threads, d, dc); run<2>(threads, d, dc); run<4>(threads, d, dc); run<8>(threads, d, dc); run<16>(threads, d, dc); run<32>(threads, d, dc);
    }
    return 0;
}
