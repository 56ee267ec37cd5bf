// Debug micro-benchmark (not part of the product): cycles of ONE half-band stage-0 pass (2048 inputs -> 1024 outputs,
// 128 threads) for different lane mappings / instruction orders, one and two CTAs per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hb_stage hb_stage.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef double2 cd;
__constant__ double c_hb[12] = {
    0.000018566625444266, -0.000118469698701817, 0.000457318798253456, -0.001347840471412094, 0.003321838571445455, -0.007198422696929033,
    0.014211106939802483, -0.026424776824073383, 0.048414810444971007, -0.096214669073304823, 0.314881034738348550, 0.5 };
__device__ __forceinline__ cd fmaz(cd a, double c, cd acc) { return make_double2(fma(a.x, c, acc.x), fma(a.y, c, acc.y)); }

// V0: complex lanes, R outputs per thread (the shipped hb_stage)
template <int R> __device__ __forceinline__ void v0(const cd *sb, cd *ob)
{
    const int t = threadIdx.x;
    const cd *w = sb + (2 * R + 1) * t;
    cd acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < 22 + R - 1; j++) {
        const cd e = w[2 * j + (2 * j) / (2 * R)];
#pragma unroll
        for (int r = 0; r < R; r++) { const int k = r + 21 - j; if (k >= 0 && k <= 21) acc[r] = fmaz(e, c_hb[k <= 10 ? k : 21 - k], acc[r]); }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const cd o = w[(21 + 2 * r) + (21 + 2 * r) / (2 * R)];
        acc[r] = fmaz(o, c_hb[11], acc[r]);
        const int m = t * R + r; ob[m + m / R] = acc[r];
    }
}
// V1: one lane per component, R outputs per lane (hb_stage_split)
template <int R> __device__ __forceinline__ void v1(const cd *sb, cd *ob)
{
    const int pr = threadIdx.x >> 1, comp = threadIdx.x & 1;
    const double *w = reinterpret_cast<const double *>(sb + (2 * R + 1) * pr) + comp;
    double acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0.0;
#pragma unroll
    for (int j = 0; j < 22 + R - 1; j++) {
        const double e = w[2 * (2 * j + (2 * j) / (2 * R))];
#pragma unroll
        for (int r = 0; r < R; r++) { const int k = r + 21 - j; if (k >= 0 && k <= 21) acc[r] = fma(e, c_hb[k <= 10 ? k : 21 - k], acc[r]); }
    }
    double *od = reinterpret_cast<double *>(ob);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const double o = w[2 * ((21 + 2 * r) + (21 + 2 * r) / (2 * R))];
        acc[r] = fma(o, c_hb[11], acc[r]);
        const int m = pr * R + r; od[2 * (m + m / R) + comp] = acc[r];
    }
}
// V2: like V1 but the window is read through volatile loads a fixed distance ahead (source order = issue order)
template <int R, int AHEAD> __device__ __forceinline__ void v2(const cd *sb, cd *ob)
{
    const int pr = threadIdx.x >> 1, comp = threadIdx.x & 1;
    const unsigned base = (unsigned)__cvta_generic_to_shared(reinterpret_cast<const double *>(sb + (2 * R + 1) * pr) + comp);
    constexpr int NE = 22 + R - 1;
    double acc[R], e[NE];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0.0;
#pragma unroll
    for (int j = 0; j < AHEAD; j++) asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(e[j]) : "r"(base + 16 * (2 * j + (2 * j) / (2 * R))));
#pragma unroll
    for (int j = 0; j < NE; j++) {
        if (j + AHEAD < NE) asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(e[j + AHEAD]) : "r"(base + 16 * (2 * (j + AHEAD) + (2 * (j + AHEAD)) / (2 * R))));
#pragma unroll
        for (int r = 0; r < R; r++) { const int k = r + 21 - j; if (k >= 0 && k <= 21) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[r]) : "d"(e[j]), "d"(c_hb[k <= 10 ? k : 21 - k])); }
    }
    const double *w = reinterpret_cast<const double *>(sb + (2 * R + 1) * pr) + comp;
    double *od = reinterpret_cast<double *>(ob);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const double o = w[2 * ((21 + 2 * r) + (21 + 2 * r) / (2 * R))];
        acc[r] = fma(o, c_hb[11], acc[r]);
        const int m = pr * R + r; od[2 * (m + m / R) + comp] = acc[r];
    }
}
template <int V, int LIVE, int MINB> __global__ void __launch_bounds__(128, MINB) k(long long *cyc, int reps, const double *gl, double *go)
{
    double live[LIVE > 0 ? LIVE : 1];
#pragma unroll
    for (int i = 0; i < LIVE; i++) live[i] = gl[i * 128 + threadIdx.x];
    extern __shared__ double smraw[];
    cd *sb = reinterpret_cast<cd *>(smraw);
    cd *ob = sb + 2400;
    for (int i = threadIdx.x; i < 2400 + 1200; i += 128) sb[i] = make_double2(i * 1e-3, -i * 2e-3);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < reps; it++) {
        if (V == 0) v0<8>(sb, ob);
        if (V == 1) v1<16>(sb, ob);
        if (V == 2) v2<16, 3>(sb, ob);
        if (V == 3) v2<16, 6>(sb, ob);
        if (V == 4) v1<8>(sb, ob);
        if (V == 5) v0<4>(sb, ob);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < LIVE; i++) live[i] = fma(live[i], 1.0000001, 1e-9);
    }
    long long t1 = clock64();
    { double s = 0;
#pragma unroll
      for (int i = 0; i < LIVE; i++) s += live[i];
      if (LIVE) go[blockIdx.x * 128 + threadIdx.x] = s; }
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (t1 - t0) / reps;
}

template <int V, int LIVE, int MINB> void run(const char *name, long long *dc, double *gl, double *go)
{
    const size_t sh = 3600 * 16;
    cudaFuncSetAttribute(k<V, LIVE, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 400;
    for (int ctas : {148, 296, 444}) {
        if (LIVE > 64 && ctas > 296) continue;
        k<V, LIVE, MINB><<<ctas, 128, sh>>>(dc, reps, gl, go);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k<V, LIVE, MINB><<<ctas, 128, sh>>>(dc, reps, gl, go);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s live %3d, %d CTA/SM: CTA0 %5lld cycles/pass, kernel %6.0f cycles/pass @1.965 GHz (%s)\n", name, LIVE, ctas / 148, c, ms * 1e-3 * 1.965e9 / reps, cudaGetErrorString(cudaGetLastError()));
    }
}
int main()
{
    long long *dc; cudaMalloc(&dc, 8);
    double *gl, *go; cudaMalloc(&gl, 256 * 128 * 8); cudaMemset(gl, 0, 256 * 128 * 8); cudaMalloc(&go, 444 * 128 * 8);
    run<0, 0, 1>("v0 complex R=8 minb1", dc, gl, go);
    run<1, 0, 1>("v1 component R=16 minb1", dc, gl, go);
    run<0, 0, 2>("v0 complex R=8 minb2", dc, gl, go);
    run<1, 0, 2>("v1 component R=16 minb2", dc, gl, go);
    run<0, 72, 1>("v0 complex R=8 minb1", dc, gl, go);
    run<1, 72, 1>("v1 component R=16 minb1", dc, gl, go);
    run<0, 72, 2>("v0 complex R=8 minb2", dc, gl, go);
    run<1, 72, 2>("v1 component R=16 minb2", dc, gl, go);
    return 0;
}
