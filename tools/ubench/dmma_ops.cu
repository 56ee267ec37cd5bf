// Micro-benchmark (evidence for DESIGN.md section 4.5, not part of the product): FP64 tensor-core DMMA (mma.sync m8n8k4 f64)
// against FP64 vector DFMA on this part, alone and together.  The question it answers: would casting the 1121-tap WDSP
// resampler (resample.c:121-157) as a Toeplitz contraction on the tensor cores buy anything over DFMA?
//   * "dfma"  : 8 independent DFMA chains per thread
//   * "dmma"  : 4 independent m8n8k4 accumulator chains per warp (each instruction = 8 x 8 x 4 = 256 FMA = 8 per lane)
//   * "both"  : the two instruction streams interleaved in the same warps
// Reported in T FMA/s over the whole GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_ops dmma_ops.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
    for (int i = 0; i < iters; i++) {
        if (MODE & 1) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
        if (MODE & 2) { dmma(c0, c1, a, b); dmma(c2, c3, a, b); dmma(c4, c5, a, b); dmma(c6, c7, a, b); }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7)) + c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

template <int MODE> void run(const char *name, double *d, int n_sm)
{
    const int blocks = n_sm * 8, iters = 1 << 14;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 256, 0.999999, 1e-9);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (MODE & 1) ? (double)blocks * 256 * 8.0 * iters : 0.0;
    const double mma = (MODE & 2) ? (double)blocks * 256 * 4.0 * 8.0 * iters : 0.0;         // 4 instructions x 8 FMA per lane
    printf("%-5s %.3f ms: DFMA %.2f T FMA/s, DMMA %.2f T FMA/s, total %.2f T FMA/s (%s)\n", name, ms, dfma / ms / 1e9, mma / ms / 1e9, (dfma + mma) / ms / 1e9,
           cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    double *d; cudaMalloc(&d, (size_t)n_sm * 8 * 256 * 8);
    run<1>("dfma", d, n_sm); run<2>("dmma", d, n_sm); run<3>("both", d, n_sm);
    return 0;
}
