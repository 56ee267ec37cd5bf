// quisk_b200/csrc/fft_device.cuh -- in-house FP64 complex FFT in shared memory.
//
// Stockham autosort, radix-4 passes (plus one radix-2 pass when log2 n is odd), on ONE
// shared-memory buffer of n complex doubles: every thread pulls its butterflies' inputs
// into registers, the CTA synchronises, and the outputs go back to the same buffer at
// their autosort positions -- so an 8192-point transform needs 128 KiB, not 256.
// Twiddles come from a per-size table exp(-2 pi i k / n), k < n, built on the host with
// libm and kept in global memory (L2 resident); the backward transform conjugates them.
// Unnormalised in both directions, like FFTW, which the reference calls for these
// stages (quisk.c:5215, wdsp/firmin.c:413,428).
//
// A CTA may run several transforms side by side: threadIdx.y selects the transform,
// threadIdx.x are the FFT_T(n) lanes cooperating on it.  All transforms of the CTA must
// call fft_smem together (it uses __syncthreads).
#pragma once
#include "qc_common.cuh"

namespace qc {

static constexpr int FFT_BPT = 2;                         // radix-4 butterflies per thread per pass
__host__ __device__ inline int fft_threads(int n) { int t = n / (4 * FFT_BPT); return t < 1 ? 1 : t; }

__device__ __forceinline__ cd cmul(cd a, cd b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }

// s: this transform's n-element buffer in shared memory; tw: n-entry forward twiddle table;
// sign: -1 forward, +1 backward; lane / lanes: this thread's index among the transform's threads.
__device__ inline void fft_smem(cd *s, int n, const cd *__restrict__ tw, int sign, int lane, int lanes)
{
    const double sg = (double)sign;
    int len = n, stride = 1;
    while (len >= 4) {
        const int n1 = len >> 2;
        const int tstep = n / len;
        const int nb = n >> 2;                           // butterflies in this pass
        cd r[FFT_BPT][4];
        int ob[FFT_BPT];
#pragma unroll
        for (int u = 0; u < FFT_BPT; u++) {
            const int b = lane + u * lanes;
            ob[u] = -1;
            if (b < nb) {
                const int p = b / stride, q = b - p * stride;
                const cd a = s[q + stride * p];
                const cd bb = s[q + stride * (p + n1)];
                const cd c = s[q + stride * (p + 2 * n1)];
                const cd d = s[q + stride * (p + 3 * n1)];
                const cd apc = cadd(a, c), amc = csub(a, c), bpd = cadd(bb, d), bmd = csub(bb, d);
                const cd jb = make_double2(-sg * bmd.y, sg * bmd.x);        // (sign*i)(b-d)
                cd w1 = tw[p * tstep], w2 = tw[2 * p * tstep], w3 = tw[3 * p * tstep];
                if (sign > 0) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
                r[u][0] = cadd(apc, bpd);
                r[u][1] = cmul(cadd(amc, jb), w1);
                r[u][2] = cmul(csub(apc, bpd), w2);
                r[u][3] = cmul(csub(amc, jb), w3);
                ob[u] = q + stride * 4 * p;
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < FFT_BPT; u++) {
            if (ob[u] >= 0) {
                s[ob[u]] = r[u][0];
                s[ob[u] + stride] = r[u][1];
                s[ob[u] + 2 * stride] = r[u][2];
                s[ob[u] + 3 * stride] = r[u][3];
            }
        }
        __syncthreads();
        len >>= 2;
        stride <<= 2;
    }
    if (len == 2) {
        // final radix-2 pass: stride == n/2, no twiddles
        const int nb = n >> 1;
        cd r0[2 * FFT_BPT], r1[2 * FFT_BPT];
#pragma unroll
        for (int u = 0; u < 2 * FFT_BPT; u++) {
            const int q = lane + u * lanes;
            if (q < nb) { const cd a = s[q], b = s[q + nb]; r0[u] = cadd(a, b); r1[u] = csub(a, b); }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2 * FFT_BPT; u++) {
            const int q = lane + u * lanes;
            if (q < nb) { s[q] = r0[u]; s[q + nb] = r1[u]; }
        }
        __syncthreads();
    }
}

// Host: device pointer to the forward twiddle table for size n on the current device (cached).
const cd *fft_twiddles(int n);
// Host: validates n (power of two, 8..8192) and returns log2(n), or -1.
int fft_log2(int n);
// CTA shape for size n: lanes per transform and transforms per CTA.
void fft_shape(int n, int *lanes, int *per_cta);

}  // namespace qc
