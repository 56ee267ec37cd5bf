// Debug micro-benchmark (not part of the product): the two sequential-lane loops of wdsp_rxa_fused.cu in isolation, cycles per
// sample on one warp of an otherwise idle SM, to separate the loops' own cost from what the full kernel adds around them.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o seq_lanes seq_lanes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double lin_block(const double *x, int n, double c1, double c2, double s)
{
    double xa[4], xb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) xa[j] = j < n ? x[j] : 0.0;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) xb[j] = i + 4 + j < n ? x[i + 4 + j] : 0.0;
        const double t0 = __dmul_rn(c1, xa[0]), t1 = __dmul_rn(c1, xa[1]), t2 = __dmul_rn(c1, xa[2]), t3 = __dmul_rn(c1, xa[3]);
        s = __dadd_rn(__dmul_rn(c2, s), t0);
        s = __dadd_rn(__dmul_rn(c2, s), t1);
        s = __dadd_rn(__dmul_rn(c2, s), t2);
        s = __dadd_rn(__dmul_rn(c2, s), t3);
#pragma unroll
        for (int j = 0; j < 4; j++) xa[j] = xb[j];
    }
    return s;
}

__device__ __forceinline__ double agc_run(const double *A, double *RV, int n, double M, double kf, double kof, double minv, bool store, double v, double &fb)
{
    double rmn[8], abn[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { rmn[j] = RV[j]; abn[j] = A[j]; }
    double f = fb;
    int i = 0;
    bool ok = true;
    while (i + 8 <= n && ok) {
        double rm[8], ab[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { rm[j] = rmn[j]; ab[j] = abn[j]; }
        const int nx = i + 8;
#pragma unroll
        for (int j = 0; j < 8; j++) { rmn[j] = nx + j < n ? RV[nx + j] : 0.0; abn[j] = nx + j < n ? A[nx + j] : 0.0; }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const bool p = rm[j] >= v;
            const double d = __dsub_rn(rm[j], v);
            v = __dadd_rn(v, __dmul_rn(d, M));
            f = __dadd_rn(__dmul_rn(kf, ab[j]), __dmul_rn(kof, f));
            ok = ok && !p && !(v < minv);
            RV[i + j] = v; (void)store;
        }
        i += 8;
    }
    fb = f;
    return v;
}

__global__ void k(double *out, long long *cyc, int n, int mode)
{
    extern __shared__ double sm[];
    double *A = sm, *RV = sm + n, *X = sm + 2 * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { A[i] = 0.3 + 1e-4 * (i % 7); RV[i] = 0.2; X[i] = 0.1 + 1e-3 * (i % 5); }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double r = 0;
    if (warp == 4 && (mode & 1)) {
        long long t0 = clock64();
        double fb = 0.1;
        r = agc_run(A, RV, n, 2e-5, 1e-4, 0.9999, 1e-6, lane == 0, 0.9, fb) + fb;
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    }
    if (warp == 5 && (mode & 2)) {
        long long t0 = clock64();
        r = lin_block(lane < 3 ? X : A, n, lane & 1 ? 0.0 : 1e-4, 0.9999, 0.5);
        long long t1 = clock64();
        if (lane == 0 && blockIdx.x == 0) cyc[1] = t1 - t0;
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + RV[threadIdx.x % n];
}

int main()
{
    const int n = 1024;
    double *d; long long *dc, h[2];
    cudaMalloc(&d, 148 * 192 * 8); cudaMalloc(&dc, 16);
    for (int mode : {1, 2, 3}) {
        cudaMemset(dc, 0, 16);
        k<<<64, 192, 3 * n * 8>>>(d, dc, n, mode);
        cudaDeviceSynchronize();
        cudaMemcpy(h, dc, 16, cudaMemcpyDeviceToHost);
        printf("mode %d: agc run %.1f cycles/sample, lin %.1f cycles/sample (%s)\n", mode, h[0] / (double)n, h[1] / (double)n, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
