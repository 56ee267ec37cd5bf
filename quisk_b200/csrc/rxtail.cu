// quisk_b200/csrc/rxtail.cu -- fused low-rate tail of quisk_process_demodulate for the SSB / CW modes:
//   cRxFilterOut (quisk.c:1218-1256)  ->  re -/+ im (quisk.c:1916,1939,1962,1986)
//   ->  quisk_dInterpolate x2 (filter.c:167-201)  ->  quisk_dInterp2HB45 (filter.c:420-453) [x2 for CW]
// in ONE kernel, one CTA per channel, every intermediate in shared memory.  At 12 kS/s (6 kS/s for CW) this
// is 1/128 of the input rate: the point of fusing is launch count and latency (five small kernels became
// 13 % of a step), not bandwidth.  Arithmetic is FP64 FMA (the per-stage exact kernels remain available with
// fused = 0); state lives in the same history arrays as the unfused BatchFilter objects.
#include "rxchain.h"

namespace qc {

static constexpr int TT_MIN = 64;   // threads per CTA: 64 (blocks up to 256 samples), 128, 256 (long blocks)
static constexpr int TR = 4;        // receive-filter outputs per thread
// The filter phase reads X with a lane stride of TR = 4 samples (64 bytes): stored densely that is a 4-way bank
// conflict on every load, so sample e lives at xp(e) = e + e/4, which sends eight consecutive lanes to eight
// different 16-byte columns.
__host__ __device__ __forceinline__ int xp(int e) { return e + (e >> 2); }

struct TailParams {
    const cd *in; long in_stride; int n;
    double *out; long out_stride;
    int N;                          // receive filter taps
    const double *rx_coef;          // [N][2] permuted taps (hI[m], hQ[m]), m = age of the sample
    const cd *rx_hin; cd *rx_hout;  // [C][N-1]
    int lower;                      // 1: re + im, 0: re - im
    int ntap_i;                     // interpolator taps (K = ntap_i / 2 per phase)
    const double *i_coef;
    const double *i_hin; double *i_hout; int i_H;   // [C][i_H]
    int n_hb;                       // 1 or 2 half-band interpolators
    const double *hb_hin[2]; double *hb_hout[2];    // [C][22]
};

__constant__ double c_hbt[12] = {        // filter.c:381-384
    0.000018566625444266, -0.000118469698701817, 0.000457318798253456,
    -0.001347840471412094, 0.003321838571445455, -0.007198422696929033,
    0.014211106939802483, -0.026424776824073383, 0.048414810444971007,
    -0.096214669073304823, 0.314881034738348550, 0.500000000000000000 };

// half-band x2 on a shared-memory line: in[H + i] (H = 22 history), i < cnt -> out[2i], out[2i+1]
template <int TT>
__device__ __forceinline__ void hb_interp_line(const double *in, int cnt, double *out, bool to_global)
{
    for (int i = threadIdx.x; i < cnt; i += TT) {
        const double *s = in + 22 + i;                  // s[-k] = samples[k]
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 11; k++) acc = fma(s[-k] + s[-(21 - k)], c_hbt[k], acc);
        out[2 * i] = s[-11] * c_hbt[11] * 2.0;
        out[2 * i + 1] = acc * 2.0;
    }
    (void)to_global;
}

template <int TT>
__global__ void __launch_bounds__(TT) rx_tail_kernel(TailParams P)
{
    extern __shared__ double sm_raw[];
    const int c = blockIdx.x, tid = threadIdx.x;
    const int N = P.N, Hx = N - 1, n = P.n;
    constexpr int XO = 8;                                           // zero slots below sample 0 (a tap block may look there)
    const int N4 = (N + 3) & ~3;                                    // taps padded with zeros to whole blocks of four
    cd *X = reinterpret_cast<cd *>(sm_raw);                         // logical sample e in [-XO, Hx + n + TR) lives at xp(e + XO)
    double2 *hc = reinterpret_cast<double2 *>(X + xp(XO + Hx + n + TR) + 1);    // [N4]
    double *D = reinterpret_cast<double *>(hc + N4);                // [i_H + n]
    double *ic = D + P.i_H + n;                                     // [ntap_i]
    double *E = ic + P.ntap_i;                                      // [22 + 2n]
    double *F = E + 22 + 2 * n;                                     // [22 + 4n] (CW only: not allocated otherwise)
    // ---- stage
    const cd *gh = P.rx_hin + (size_t)c * Hx;
    for (int i = tid; i < XO; i += TT) X[xp(i)] = make_double2(0.0, 0.0);
    for (int i = tid; i < Hx; i += TT) X[xp(XO + i)] = gh[i];
    const cd *gx = P.in + (size_t)c * P.in_stride;
    for (int i = tid; i < n; i += TT) X[xp(XO + Hx + i)] = gx[i];
    for (int i = tid; i < TR; i += TT) X[xp(XO + Hx + n + i)] = make_double2(0.0, 0.0);
    for (int i = tid; i < N4; i += TT) hc[i] = i < N ? make_double2(P.rx_coef[2 * i], P.rx_coef[2 * i + 1]) : make_double2(0.0, 0.0);
    for (int i = tid; i < P.i_H; i += TT) D[i] = P.i_hin[(size_t)c * P.i_H + i];
    for (int i = tid; i < P.ntap_i; i += TT) ic[i] = P.i_coef[i];
    for (int i = tid; i < 22; i += TT) E[i] = P.hb_hin[0][(size_t)c * 22 + i];
    if (P.n_hb == 2) for (int i = tid; i < 22; i += TT) F[i] = P.hb_hin[1][(size_t)c * 22 + i];
    __syncthreads();
    // ---- receive filter + sideband combine.  A thread owns TR = 4 consecutive outputs and walks the taps in blocks of
    //      four: output r meets tap k0 + kk at sample m0 + r - k0 - kk, so one block touches the seven samples
    //      m0 - k0 - 3 .. m0 - k0 + 3, three of which carry over to the next block.  Per block: four sample loads, four
    //      tap loads (broadcast), 32 FMAs -- the FP64 pipe, not the issue slots, sets the pace.
    for (int m0 = tid * TR; m0 < n; m0 += TT * TR) {
        double aI[TR], aQ[TR];
        cd w[7];                                                    // w[d + 3] = X[m0 - k0 + d], d = -3 .. 3
#pragma unroll
        for (int r = 0; r < TR; r++) { aI[r] = 0.0; aQ[r] = 0.0; }
#pragma unroll
        for (int i = 0; i < 7; i++) w[i] = X[xp(XO + Hx + m0 - 3 + i)];
        for (int k0 = 0; k0 < N4; k0 += 4) {
            double2 h[4];
#pragma unroll
            for (int kk = 0; kk < 4; kk++) h[kk] = hc[k0 + kk];
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
#pragma unroll
                for (int r = 0; r < TR; r++) {
                    aI[r] = fma(w[r - kk + 3].x, h[kk].x, aI[r]);
                    aQ[r] = fma(w[r - kk + 3].y, h[kk].y, aQ[r]);
                }
            }
            w[6] = w[2]; w[5] = w[1]; w[4] = w[0];
            const int e = XO + Hx + m0 - k0 - 7;                    // >= XO + N - 1 - N4 - 3 - 3 >= 0
#pragma unroll
            for (int i = 0; i < 4; i++) w[i] = X[xp(e + i)];
        }
#pragma unroll
        for (int r = 0; r < TR; r++)
            if (m0 + r < n) D[P.i_H + m0 + r] = P.lower ? aI[r] + aQ[r] : aI[r] - aQ[r];
    }
    __syncthreads();
    // ---- quisk_dInterpolate x2: e[2m+j] = 2 sum_{k<K} d[m-k] c[j + 2k]
    const int K = P.ntap_i / 2;
    for (int m = tid; m < n; m += TT) {
        const double *d = D + P.i_H + m;
        double a0 = 0.0, a1 = 0.0;
        for (int k = 0; k < K; k++) { const double x = d[-k]; a0 = fma(x, ic[2 * k], a0); a1 = fma(x, ic[2 * k + 1], a1); }
        E[22 + 2 * m] = a0 * 2.0;
        E[22 + 2 * m + 1] = a1 * 2.0;
    }
    __syncthreads();
    // ---- half band(s)
    double *go = P.out + (size_t)c * P.out_stride;
    if (P.n_hb == 1) {
        hb_interp_line<TT>(E, 2 * n, go, true);
    } else {
        hb_interp_line<TT>(E, 2 * n, F + 22, false);
        __syncthreads();
        hb_interp_line<TT>(F, 4 * n, go, true);
    }
    // ---- histories: last N-1 filter inputs, last i_H interpolator inputs, last 22 half-band inputs
    cd *oh = P.rx_hout + (size_t)c * Hx;
    for (int i = tid; i < Hx; i += TT) oh[i] = X[xp(XO + n + i)];
    for (int i = tid; i < P.i_H; i += TT) P.i_hout[(size_t)c * P.i_H + i] = D[n + i];
    for (int i = tid; i < 22; i += TT) P.hb_hout[0][(size_t)c * 22 + i] = E[2 * n + i];
    if (P.n_hb == 2) {
        __syncthreads();
        for (int i = tid; i < 22; i += TT) P.hb_hout[1][(size_t)c * 22 + i] = F[4 * n + i];
    }
}

bool RxChain::tail_fusable() const
{
    if (!(mode == QC_MODE_LSB || mode == QC_MODE_USB || mode == QC_MODE_CWL || mode == QC_MODE_CWU || mode == QC_MODE_DGT_U ||
          mode == QC_MODE_DGT_L || mode == QC_MODE_FDV_U || mode == QC_MODE_FDV_L)) return false;
    if (!rxf || rxf->kind != QC_C_RXFILTER) return false;
    if (rst.size() < 2 || rst.size() > 3) return false;
    if (rst[0]->kind != QC_D_INTERPOLATE || rst[0]->interp != 2 || (rst[0]->nTaps & 1) || rst[0]->nTaps > 256) return false;
    for (size_t i = 1; i < rst.size(); i++) if (rst[i]->kind != QC_D_INTERP2_HB45) return false;
    return true;
}

int RxChain::run_tail(const cd *in, long in_stride, int n, double *out, long out_stride, int *n_out, cudaStream_t s)
{
    TailParams P;
    memset(&P, 0, sizeof(P));
    P.in = in; P.in_stride = in_stride; P.n = n; P.out = out; P.out_stride = out_stride;
    P.N = rxf->nTaps; P.rx_coef = rxf->d_coef;
    P.rx_hin = (const cd *)rxf->d_hist[rxf->cur]; P.rx_hout = (cd *)rxf->d_hist[rxf->cur ^ 1];
    P.lower = (mode == QC_MODE_LSB || mode == QC_MODE_CWL || mode == QC_MODE_DGT_L || mode == QC_MODE_FDV_L) ? 1 : 0;
    BatchFilter *fi = rst[0];
    P.ntap_i = fi->nTaps; P.i_coef = fi->d_coef; P.i_H = fi->H;
    P.i_hin = (const double *)fi->d_hist[fi->cur]; P.i_hout = (double *)fi->d_hist[fi->cur ^ 1];
    P.n_hb = (int)rst.size() - 1;
    for (int i = 0; i < P.n_hb; i++) {
        BatchFilter *h = rst[1 + i];
        P.hb_hin[i] = (const double *)h->d_hist[h->cur]; P.hb_hout[i] = (double *)h->d_hist[h->cur ^ 1];
    }
    const int nout = n * 2 * (P.n_hb == 2 ? 4 : 2);
    if (out_stride < nout) { set_error("rx_process: audio_stride %ld < %d", out_stride, nout); return QC_EINVAL; }
    const size_t sh = (size_t)(xp(8 + P.N - 1 + n + TR) + 1) * sizeof(cd) + (size_t)((P.N + 3) & ~3) * sizeof(double2) +
                      (size_t)(P.i_H + n + P.ntap_i + 22 + 2 * n + (P.n_hb == 2 ? 22 + 4 * n : 0) + 8) * sizeof(double);
    if (sh > 200 * 1024) return QC_ENOMEM;           // caller falls back to the per-stage kernels
    // one CTA per channel whatever the block length: long blocks (the 192 kS/s receivers hand over 2048 samples) get more threads
#define QC_TAIL(TT_) do { \
        if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(rx_tail_kernel<TT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh)); \
        rx_tail_kernel<TT_><<<C, TT_, sh, s>>>(P); } while (0)
    if (n <= TT_MIN * TR) QC_TAIL(64); else if (n <= 2 * TT_MIN * TR) QC_TAIL(128); else QC_TAIL(256);
#undef QC_TAIL
    count_launch();
    QC_CUDA_LAUNCH();
    rxf->cur ^= 1;
    for (auto *f : rst) f->cur ^= 1;
    *n_out = nout;
    return QC_OK;
}

}  // namespace qc
