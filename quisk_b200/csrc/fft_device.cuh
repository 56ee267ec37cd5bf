// quisk_b200/csrc/fft_device.cuh -- in-house FP64 complex FFT in shared memory.
//
// Stockham autosort on ONE shared-memory buffer of n complex doubles: every thread pulls its
// butterflies' inputs into registers, the CTA synchronises, and the outputs go back to the same buffer
// at their autosort positions -- so an 8192-point transform needs 128 KiB, not 256.
// Passes are radix-16 (a 4 x 4 Cooley-Tukey step done entirely in registers) while the remaining
// length allows, then one radix-4 and/or radix-2 pass: 8192 = 16.16.16.2 is four shared-memory round
// trips instead of the seven a radix-4 plan needs -- shared-memory bandwidth, not FP64, is what bounds
// this kernel.  n/16 threads cooperate on a transform; each owns 16 points per pass.
// Twiddles: a host-built (libm, exact angles) table exp(-2 pi i k / n) is staged as a two-level store in
// shared memory (see fft_stage_twiddles); a radix-16 butterfly fetches w^1, w^2, w^4, w^8 and forms the other
// powers with at most two multiplications (error <= 4 ulp).  The backward transform conjugates.
// Unnormalised in both directions, like FFTW, which the reference calls for these stages
// (quisk.c:5215, wdsp/firmin.c:413,428).
//
// A CTA may run several transforms side by side: threadIdx.y selects the transform, threadIdx.x are the
// fft_threads(n) lanes cooperating on it.  All transforms of the CTA must call fft_smem together.
#pragma once
#include "qc_common.cuh"

namespace qc {

// n/16 threads cooperate on a transform (one radix-16 butterfly each per pass); 8192 points use n/32 = 256
// threads with two butterflies each, so that the register file allows 255 registers per thread.
__host__ __device__ inline int fft_threads(int n) { int t = n / 16; if (t > 256) t = 256; return t < 1 ? 1 : t; }

__device__ __forceinline__ cd cmul(cd a, cd b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }
// Bank swizzle for the transform buffer: element i lives at fsw(i).  In a pass with stride 1 the 16 outputs
// of butterfly p go to 16p + k, i.e. consecutive lanes would hit the same 16-byte bank group; XOR-ing the
// low three index bits with bits 4..6 spreads a quarter-warp over all eight groups in every pass.
// Everything that touches the buffer (fft_smem and the kernels that fill / drain it) indexes through fsw().
__device__ __forceinline__ int fsw(int i) { return i ^ ((i >> 4) & 7); }
// (sg * i) * z
__device__ __forceinline__ cd muli(cd z, double sg) { return make_double2(-sg * z.y, sg * z.x); }

// 4-point DFT with kernel exp(sg * 2 pi i j k / 4): in a,b,c,d (j = 0..3) -> out x0..x3 (k = 0..3)
__device__ __forceinline__ void bfly4(cd a, cd b, cd c, cd d, double sg, cd &x0, cd &x1, cd &x2, cd &x3)
{
    const cd apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d), jb = muli(csub(b, d), sg);
    x0 = cadd(apc, bpd);
    x1 = cadd(amc, jb);
    x2 = csub(apc, bpd);
    x3 = csub(amc, jb);
}

// 16-point DFT in registers, kernel exp(sg * 2 pi i j k / 16), in place on v[16] (natural order in and out)
__device__ __forceinline__ void dft16(cd (&v)[16], double sg)
{
    const double C1 = 0.92387953251128674, S1 = 0.38268343236508977, R = 0.70710678118654752;
    cd t[4][4];                                           // t[b][c] = sum_a W4^{ac} v[4a + b]
#pragma unroll
    for (int b = 0; b < 4; b++) bfly4(v[b], v[4 + b], v[8 + b], v[12 + b], sg, t[b][0], t[b][1], t[b][2], t[b][3]);
    // t[b][c] *= W16^{bc}
    t[1][1] = cmul(t[1][1], make_double2(C1, sg * S1));
    t[1][2] = cmul(t[1][2], make_double2(R, sg * R));
    t[1][3] = cmul(t[1][3], make_double2(S1, sg * C1));
    t[2][1] = cmul(t[2][1], make_double2(R, sg * R));
    t[2][2] = muli(t[2][2], sg);                          // W16^4 = sg * i
    t[2][3] = cmul(t[2][3], make_double2(-R, sg * R));
    t[3][1] = cmul(t[3][1], make_double2(S1, sg * C1));
    t[3][2] = cmul(t[3][2], make_double2(-R, sg * R));
    t[3][3] = cmul(t[3][3], make_double2(-C1, -sg * S1));
#pragma unroll
    for (int c = 0; c < 4; c++) bfly4(t[0][c], t[1][c], t[2][c], t[3][c], sg, v[c], v[c + 4], v[c + 8], v[c + 12]);
}

// Two-level twiddle store in SHARED memory: exp(-2 pi i k / n) = coarse[k >> 7] * fine[k & 127], with
// fine[a] = exp(-2 pi i a / n), a < 128 and coarse[b] = exp(-2 pi i 128 b / n), b < n/128 -- 192 entries (3 KiB)
// for n = 8192 instead of a 128 KiB table behind L2.  Both halves are copied from the host-built full table,
// so each factor is correctly rounded; the product adds one rounding.
static constexpr int FFT_TW_FINE = 128;
__host__ __device__ inline int fft_tw_entries(int n) { return FFT_TW_FINE + (n > FFT_TW_FINE ? n / FFT_TW_FINE : 1); }

// all threads of the CTA: stage the two-level table for size n from the global full table
__device__ inline void fft_stage_twiddles(cd *twl, const cd *__restrict__ tw, int n)
{
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
    const int nc = n > FFT_TW_FINE ? n / FFT_TW_FINE : 1;
    for (int i = tid; i < FFT_TW_FINE; i += nt) twl[i] = i < n ? tw[i] : make_double2(1.0, 0.0);
    for (int i = tid; i < nc; i += nt) twl[FFT_TW_FINE + i] = tw[(size_t)i * FFT_TW_FINE];
}

__device__ __forceinline__ cd fft_tw(const cd *twl, int idx, int sign)
{
    cd w = cmul(twl[FFT_TW_FINE + (idx >> 7)], twl[idx & (FFT_TW_FINE - 1)]);
    if (sign > 0) w.y = -w.y;
    return w;
}

// s: this transform's n-element buffer in shared memory; twl: the staged two-level twiddle store (shared);
// sign: -1 forward, +1 backward; lane / lanes: this thread's index among the transform's threads.
// BPT = radix-16 butterflies per thread and pass: 2 covers n = 8192 on 256 lanes; kernels that only run
// n <= 16 * lanes instantiate BPT = 1 and save the second butterfly's 64 registers.
// NBAR: 0 = the transform's barriers are __syncthreads() (every thread of the CTA calls fft_smem); NBAR > 0 = named barrier 1
// over NBAR threads, for kernels whose other warps are busy with something else (wdsp_rxa_fused.cu).
template <int NBAR>
__device__ __forceinline__ void fft_sync()
{
    if (NBAR == 0) __syncthreads();
    else asm volatile("bar.sync 1, %0;" :: "r"(NBAR) : "memory");
}

template <int BPT = 2, int NBAR = 0>
__device__ inline void fft_smem(cd *s, int n, const cd *__restrict__ twl, int sign, int lane, int lanes)
{
    const double sg = (double)sign;
    int len = n, stride = 1;
    // ---- radix-16 passes: n/16 butterflies, one or two per thread
    while (len >= 16) {
        const int n1 = len >> 4;
        const int tstep = n / len;
        const int nb = n >> 4;
        cd v[BPT][16];
        int ob[BPT];
#pragma unroll
        for (int t = 0; t < BPT; t++) {
            const int b = lane + t * lanes;
            ob[t] = -1;
            if (b < nb) {
                const int p = b / stride, q = b - p * stride;
#pragma unroll
                for (int j = 0; j < 16; j++) v[t][j] = s[fsw(q + stride * (p + j * n1))];
                dft16(v[t], sg);
                if (p != 0) {
                    // w^k = exp(sg 2 pi i p k / len), k = 1..15, from w1 w2 w4 w8 and <= 2 products each
                    const int pt = p * tstep;
                    const cd w1 = fft_tw(twl, pt, sign), w2 = fft_tw(twl, 2 * pt, sign), w4 = fft_tw(twl, 4 * pt, sign), w8 = fft_tw(twl, 8 * pt, sign);
                    cd *x = v[t];
                    x[1] = cmul(x[1], w1); x[2] = cmul(x[2], w2); x[4] = cmul(x[4], w4); x[8] = cmul(x[8], w8);
                    const cd w3 = cmul(w1, w2); x[3] = cmul(x[3], w3);
                    const cd w5 = cmul(w1, w4); x[5] = cmul(x[5], w5);
                    const cd w6 = cmul(w2, w4); x[6] = cmul(x[6], w6);
                    const cd w7 = cmul(w3, w4); x[7] = cmul(x[7], w7);
                    x[9] = cmul(x[9], cmul(w1, w8)); x[10] = cmul(x[10], cmul(w2, w8)); x[11] = cmul(x[11], cmul(w3, w8));
                    x[12] = cmul(x[12], cmul(w4, w8)); x[13] = cmul(x[13], cmul(w5, w8)); x[14] = cmul(x[14], cmul(w6, w8));
                    x[15] = cmul(x[15], cmul(w7, w8));
                }
                ob[t] = q + stride * 16 * p;
            }
        }
        fft_sync<NBAR>();
#pragma unroll
        for (int t = 0; t < BPT; t++) {
            if (ob[t] >= 0) {
#pragma unroll
                for (int k = 0; k < 16; k++) s[fsw(ob[t] + k * stride)] = v[t][k];
            }
        }
        fft_sync<NBAR>();
        len >>= 4;
        stride <<= 4;
    }
    // ---- one radix-4 pass if 4 or 8 points remain per sub-transform: n/4 butterflies, 4 per thread
    //      (never at n = 8192 = 16^3 * 2, the only size with fewer than n/16 threads)
    if (len >= 4) {
        const int n1 = len >> 2;
        const int tstep = n / len;
        const int nb = n >> 2;
        cd r[4][4];
        int ob[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int b = lane + u * lanes;
            ob[u] = -1;
            if (b < nb) {
                const int p = b / stride, q = b - p * stride;
                const cd a = s[fsw(q + stride * p)], bb = s[fsw(q + stride * (p + n1))];
                const cd c = s[fsw(q + stride * (p + 2 * n1))], d = s[fsw(q + stride * (p + 3 * n1))];
                bfly4(a, bb, c, d, sg, r[u][0], r[u][1], r[u][2], r[u][3]);
                if (p != 0) {
                    const int pt = p * tstep;
                    r[u][1] = cmul(r[u][1], fft_tw(twl, pt, sign)); r[u][2] = cmul(r[u][2], fft_tw(twl, 2 * pt, sign));
                    r[u][3] = cmul(r[u][3], fft_tw(twl, 3 * pt, sign));
                }
                ob[u] = q + stride * 4 * p;
            }
        }
        fft_sync<NBAR>();
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (ob[u] >= 0) {
                s[fsw(ob[u])] = r[u][0]; s[fsw(ob[u] + stride)] = r[u][1]; s[fsw(ob[u] + 2 * stride)] = r[u][2]; s[fsw(ob[u] + 3 * stride)] = r[u][3];
            }
        }
        fft_sync<NBAR>();
        len >>= 2;
        stride <<= 2;
    }
    // ---- final radix-2 pass (stride == n/2, no twiddles): each butterfly reads and writes the same two
    //      locations, so it runs in place without staging or a barrier in between
    if (len == 2) {
        const int nb = n >> 1;
        for (int q = lane; q < nb; q += lanes) {
            const cd a = s[fsw(q)], b = s[fsw(q + nb)];
            s[fsw(q)] = cadd(a, b);
            s[fsw(q + nb)] = csub(a, b);
        }
        fft_sync<NBAR>();
    }
}

// Host: device pointer to the forward twiddle table for size n on the current device (cached).
const cd *fft_twiddles(int n);
// Host: validates n (power of two, 8..8192) and returns log2(n), or -1.
int fft_log2(int n);
// CTA shape for size n: lanes per transform and transforms per CTA.
void fft_shape(int n, int *lanes, int *per_cta);

}  // namespace qc
