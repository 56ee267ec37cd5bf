// quisk_b200/csrc/wdsp_snba_nofma.cu -- WDSP's spectral noise blanker "SNB" (wdsp/snb.c, helpers wdsp/lmath.c) for a batch
// of channels.
//
// xsnba (snb.c:539-572) works on the real rail at an internal rate of 12 kS/s (resamplers on the way in and out when the
// channel runs faster, snb.c:31-66): frames of xsize = 256 samples advance by incr = xsize / ovrlp = 64; every frame
//   1. fits a linear predictor of order asize = 64 to the frame (asolve, lmath.c:93-125: autocorrelation + Levinson),
//   2. forms the two-sided prediction error (invf, snb.c:306-322) and flags the samples whose error power stands out
//      against a median-based threshold, bridging short gaps and widening the flags a little (det, snb.c:324-402),
//   3. zeroes the flagged samples and lists the runs of flags as "impulses" with the clean stretch before and after each
//      (scanFrame, snb.c:404-490: the best-conditioned impulse first),
//   4. in `npasses` passes, for every impulse that has enough clean context: refits a predictor of order p on the
//      cleaned frame and replaces the impulse by the least-squares interpolation that predictor implies (xHat,
//      snb.c:265-304: normal equations of a banded system whose Gram matrix is Toeplitz, inverted by Durbin + Trench,
//      lmath.c:29-91), else puts the original samples back (execFrame, snb.c:492-537).
// The frame buffer keeps a second frame of history in front (snb.c:93-94: the autocorrelation reaches back asize samples).
// Mapping: ONE WARP per channel and frame.  The frame logic is a chain of small, data-dependent steps on 256 samples, so
// what the lanes share are the loops over OUTPUT elements (autocorrelation lags, prediction errors, matrix entries); every
// individual sum runs over its terms in the reference's order, the state machines (gap bridging, impulse list, Levinson,
// Durbin) run on lane 0.  The banded matrices A1 / A2 of xHat are not materialised: their entries are functions of the
// predictor coefficients, and the products skip nothing but exact zeros.  Many channels = many warps.
// Compiled with --fmad=false (file name rule in build.py).
#include <vector>
#include "wdsp_internal.h"
#include "../../include/quisk_cuda_wdsp.h"

namespace qc {

static constexpr int SN_MAXIMP = 256;       // snb.c:29
static constexpr int SN_XMAX = 256;         // frame size this build is sized for (create_rxa: 256)
static constexpr int SN_AMAX = 64;          // predictor order (create_rxa: 64)
static constexpr int SN_UMAX = 2 * SN_AMAX; // longest impulse that can be interpolated: p_opt >= pmultmin * limp with pmultmin >= 0.5

struct SnbaPar {
    int xsize, ovrlp, incr, asize, npasses, b, pre, post, isize, iasize, oasize;
    double k1, k2, pmultmin;
};
struct SnbaLayout { size_t xbase, inaccum, outaccum, detout, ATAI, P1, row; };

struct SnSm {           // one warp's shared memory
    double x[2 * SN_XMAX];                  // xbase: [history | frame]
    double a[SN_XMAX], v[SN_XMAX], savex[SN_XMAX], vpwr[SN_XMAX], vp[SN_XMAX];
    int detout[SN_XMAX], unfixed[SN_XMAX];
    double r[SN_UMAX + 2], z[SN_UMAX + 2];
    double xh[SN_UMAX], P2[SN_UMAX], ty[SN_UMAX], tv[SN_UMAX], tz[SN_UMAX];
    int bimp[SN_MAXIMP], limp[SN_MAXIMP], befimp[SN_MAXIMP], aftimp[SN_MAXIMP], p_opt[SN_MAXIMP];
    int nimp, next, isc;
    double dsc;
};

// lmath.c:93-125.  x points at the frame; x[-1 .. -asize] is history.
__device__ void sn_asolve(SnSm &s, const double *x, int xsize, int asize, int lane)
{
    for (int i = lane; i <= asize; i += 32) {
        double acc = 0.0;
        for (int j = 0; j < xsize; j++) acc += x[j] * x[j - i];
        s.r[i] = acc;
    }
    __syncwarp();
    if (lane == 0) {
        double *r = s.r, *z = s.z;
        for (int i = 0; i <= asize; i++) z[i] = 0.0;
        z[0] = 1.0;
        double beta = r[0];
        for (int k = 0; k < asize; k++) {
            double alpha = 0.0;
            for (int j = 0; j <= k; j++) alpha -= z[j] * r[k + 1 - j];
            alpha /= beta;
            for (int i = 0; i <= (k + 1) / 2; i++) {
                const double t = z[k + 1 - i] + alpha * z[i];
                z[i] = z[i] + alpha * z[k + 1 - i];
                z[k + 1 - i] = t;
            }
            beta *= 1.0 - alpha * alpha;
        }
        for (int i = 0; i < asize; i++) { double ai = -z[i + 1]; if (ai != ai) ai = 0.0; s.a[i] = ai; }
    }
    __syncwarp();
}

// snb.c:306-322
__device__ void sn_invf(SnSm &s, const double *x, int xsize, int asize, int lane)
{
    for (int i = lane; i < xsize; i += 32) {
        double acc = 0.0;
        if (i >= asize && i < xsize - asize) {
            for (int j = 0; j < asize; j++) acc += s.a[j] * (x[i - 1 - j] + x[i + 1 + j]);
            acc = x[i] - 0.5 * acc;
        } else if (i >= xsize - asize) {
            for (int j = 0; j < asize; j++) acc += s.a[j] * x[i - 1 - j];
            acc = x[i] - acc;
        }
        s.v[i] = acc;
    }
    __syncwarp();
}

// snb.c:324-402
__device__ void sn_det(SnSm &s, const SnbaPar &P, int asize, int lane)
{
    const int xsize = P.xsize, n = xsize - asize;
    for (int i = asize + lane; i < xsize; i += 32) { const double w = s.v[i] * s.v[i]; s.vpwr[i] = w; s.vp[i - asize] = w; }
    __syncwarp();
    // the median the reference's selection routine returns is the element of rank n / 2 (lmath.c:127-184): found here by counting
    const int k = n / 2;
    for (int c = lane; c < n; c += 32) {
        const double val = s.vp[c];
        int less = 0, eq = 0;
        for (int j = 0; j < n; j++) { const double w = s.vp[j]; less += w < val; eq += w == val; }
        if (less <= k && k < less + eq) s.dsc = val;        // every lane that qualifies writes the same value
    }
    __syncwarp();
    if (lane == 0) {
        const double medpwr = s.dsc;
        const double t1 = P.k1 * medpwr;
        double t2 = 0.0;
        for (int i = asize; i < xsize; i++) {
            if (s.vpwr[i] <= t1) t2 += s.vpwr[i];
            else if (s.vpwr[i] <= 2.0 * t1) t2 += 2.0 * t1 - s.vpwr[i];
        }
        t2 *= P.k2 / (double)(xsize - asize);
        int *detout = s.detout;
        for (int i = asize; i < xsize; i++) detout[i] = s.vpwr[i] > t2 ? 1 : 0;
        int bstate = 0, bcount = 0, bsamp = 0;
        for (int i = asize; i < xsize; i++) {
            switch (bstate) {
            case 0: if (detout[i] == 1) bstate = 1; break;
            case 1: if (detout[i] == 0) { bstate = 2; bsamp = i; bcount = 1; } break;
            case 2:
                ++bcount;
                if (bcount > P.b) { bstate = detout[i] == 1 ? 1 : 0; }
                else if (detout[i] == 1) { for (int j = bsamp; j < bsamp + bcount - 1; j++) detout[j] = 1; bstate = 1; }
                break;
            }
        }
        for (int i = asize; i < xsize; i++)
            if (detout[i] == 1) for (int j = i - 1; j > i - 1 - P.pre; j--) if (j >= asize) detout[j] = 1;
        for (int i = xsize - 1; i >= asize; i--)
            if (detout[i] == 1) for (int j = i + 1; j < i + 1 + P.post; j++) if (j < xsize) detout[j] = 1;
    }
    __syncwarp();
}

// snb.c:404-490, lane 0 only; results in s.bimp .. s.p_opt, s.next; returns nimp
__device__ int sn_scan(SnSm &s, int xsize, int pval, double pmultmin, const int *det)
{
    int inflag = 0, i = 0, nimp = 0;
    double merit[SN_MAXIMP];
    int nextlist[SN_MAXIMP];
    for (int q = 0; q < SN_MAXIMP; q++) { s.befimp[q] = 0; s.aftimp[q] = 0; merit[q] = 0.0; }
    while (i < xsize && nimp < SN_MAXIMP) {
        if (det[i] == 1 && inflag == 0) { inflag = 1; s.bimp[nimp] = i; s.limp[nimp] = 1; nimp++; }
        else if (det[i] == 1) s.limp[nimp - 1]++;
        else { inflag = 0; s.befimp[nimp]++; if (nimp > 0) s.aftimp[nimp - 1]++; }
        i++;
    }
    for (i = 0; i < nimp; i++) {
        int po = s.befimp[i] < s.aftimp[i] ? s.befimp[i] : s.aftimp[i];
        if (po > pval) po = pval;
        if (po < (int)(pmultmin * s.limp[i])) po = -1;
        s.p_opt[i] = po;
    }
    for (i = 0; i < nimp; i++) { merit[i] = (double)s.p_opt[i] / (double)s.limp[i]; nextlist[i] = i; }
    for (int j = 0; j < nimp - 1; j++)
        for (int k = 0; k < nimp - j - 1; k++)
            if (merit[k] < merit[k + 1]) {
                const double td = merit[k]; const int ti = nextlist[k];
                merit[k] = merit[k + 1]; nextlist[k] = nextlist[k + 1];
                merit[k + 1] = td; nextlist[k + 1] = ti;
            }
    i = 1;
    if (nimp > 0) while (i < nimp && merit[i] == merit[0]) i++;
    for (int j = 0; j < i - 1; j++)
        for (int k = 0; k < i - j - 1; k++)
            if (s.limp[nextlist[k]] < s.limp[nextlist[k + 1]]) {
                const double td = merit[k]; const int ti = nextlist[k];
                merit[k] = merit[k + 1]; nextlist[k] = nextlist[k + 1];
                merit[k + 1] = td; nextlist[k + 1] = ti;
            }
    s.next = nimp > 0 ? nextlist[0] : 0;
    return nimp;
}

// entries of the two banded matrices of xHat (snb.c:281-300) as functions of the predictor a[0 .. p): row k, column i
__device__ __forceinline__ double sn_A1(const double *a, int p, int k, int i)
{   // a1rows = xu + p rows, xu columns: 1 on the diagonal, -a[k - i - 1] for the p rows below it
    const int d = k - i;
    return d == 0 ? 1.0 : (d >= 1 && d <= p ? -a[d - 1] : 0.0);
}
__device__ __forceinline__ double sn_A2(const double *a, int p, int xu, int k, int j)
{   // a1rows rows, xu + 2 p columns
    if (j < p) return k <= j ? a[p - j - 1 + k] : 0.0;                     // left block: column j holds a[p-j-1 ..] in rows 0 .. j
    if (j >= p + xu) {                                                      // right block
        const int d = k - (j - p);
        return d == 0 ? -1.0 : (d >= 1 && k < xu + p ? a[d - 1] : 0.0);
    }
    return 0.0;
}

// Durbin recursion + Trench inverse of the symmetric Toeplitz matrix with first column r (lmath.c:29-91), n x n into B
__device__ void sn_trI(SnSm &s, int n, double *B, int lane)
{
    double *r = s.r, *y = s.ty, *v = s.tv, *z = s.tz;
    if (lane == 0) {
        for (int i = 0; i < n - 1; i++) { y[i] = 0.0; v[i] = 0.0; }
        const double scale = 1.0 / r[0];
        for (int i = 0; i < n; i++) r[i] *= scale;
        s.dsc = scale;
        // dR(n - 1, r, y, z)
        const int m = n - 1;
        if (m >= 1) {
            for (int i = 0; i < m - 1; i++) z[i] = 0.0;
            y[0] = -r[1];
            double alpha = -r[1], beta = 1.0;
            for (int k = 0; k < m - 1; k++) {
                beta *= 1.0 - alpha * alpha;
                double gamma = 0.0;
                for (int i = k + 1, j = 0; i > 0; i--, j++) gamma += r[i] * y[j];
                alpha = -(r[k + 2] + gamma) / beta;
                for (int i = 0, j = k; i <= k; i++, j--) z[i] = y[i] + alpha * y[j];
                for (int i = 0; i <= k; i++) y[i] = z[i];
                y[k + 1] = alpha;
            }
        }
        double t = 0.0;
        for (int i = 0; i < n - 1; i++) t += r[i + 1] * y[i];
        const double gamma = 1.0 / (1.0 + t);
        for (int i = 0, j = n - 2; i < n - 1; i++, j--) v[i] = gamma * y[j];
        B[0] = gamma;
        for (int i = 1, j = n - 2; i < n; i++, j--) B[i] = v[j];
        for (int i = 1; i <= (n - 1) / 2; i++)
            for (int j = i; j < n - i; j++)
                B[i * n + j] = B[(i - 1) * n + (j - 1)] + (v[n - j - 1] * v[n - i - 1] - v[i - 1] * v[j - 1]) / gamma;
        for (int i = 0; i <= (n - 1) / 2; i++)
            for (int j = i; j < n - i; j++) {
                const double b = B[i * n + j] *= scale;
                B[j * n + i] = b;
                const int ni = n - i - 1, nj = n - j - 1;
                B[ni * n + nj] = b;
                B[nj * n + ni] = b;
            }
    }
    __syncwarp();
}

// snb.c:265-304: xout[0 .. xu) from the window xk[0 .. xu + 2 p) (p clean samples, the impulse, p clean samples)
__device__ void sn_xhat(SnSm &s, int xu, int p, const double *xk, double *ATAI, double *P1, int lane)
{
    const double *a = s.a;
    const int a1rows = xu + p, a2cols = xu + 2 * p;
    // r = first column of A1^T A1 (ATAc0, snb.c:209-216)
    for (int i = lane; i < xu; i += 32) {
        double acc = 0.0;
        for (int j = 0; j < a1rows; j++) acc += sn_A1(a, p, j, i) * sn_A1(a, p, j, 0);
        s.r[i] = acc;
    }
    __syncwarp();
    sn_trI(s, xu, ATAI, lane);
    // P1 = A1^T A2, only the two blocks of columns that are not zero (multA1TA2, snb.c:218-239)
    const int q = a1rows;
    for (int e = lane; e < xu * a2cols; e += 32) {
        const int i = e / a2cols, j = e - i * a2cols;
        double c = 0.0;
        if (j < p) { const int hi = i + p < j ? i + p : j; for (int k = i; k <= hi; k++) c += sn_A1(a, p, k, i) * sn_A2(a, p, xu, k, j); }
        if (j >= a2cols - p) { const int lo = i > q - (a2cols - j) ? i : q - (a2cols - j); for (int k = lo; k <= i + p; k++) c += sn_A1(a, p, k, i) * sn_A2(a, p, xu, k, j); }
        P1[e] = c;
    }
    __syncwarp();
    // P2 = P1 * xk over the clean samples (multXKE, snb.c:241-252)
    for (int i = lane; i < xu; i += 32) {
        double acc = 0.0;
        for (int k = i; k < p; k++) acc += P1[i * a2cols + k] * xk[k];
        for (int k = a2cols - p; k <= a2cols - xu + i; k++) acc += P1[i * a2cols + k] * xk[k];
        s.P2[i] = acc;
    }
    __syncwarp();
    // xout = ATAI * P2 (multAv, snb.c:254-263)
    for (int i = lane; i < xu; i += 32) {
        double acc = 0.0;
        for (int k = 0; k < xu; k++) acc += ATAI[i * xu + k] * s.P2[k];
        s.xh[i] = acc;
    }
    __syncwarp();
}

// one frame of one channel (execFrame, snb.c:492-537) on the frame s.x + xsize
__device__ void sn_exec_frame(SnSm &s, const SnbaPar &P, double *ATAI, double *P1, int lane)
{
    const int xsize = P.xsize;
    double *x = s.x + xsize;
    for (int i = lane; i < xsize; i += 32) s.savex[i] = x[i];
    __syncwarp();
    sn_asolve(s, x, xsize, P.asize, lane);
    sn_invf(s, x, xsize, P.asize, lane);
    sn_det(s, P, P.asize, lane);
    for (int i = lane; i < xsize; i += 32) if (s.detout[i] != 0) x[i] = 0.0;
    __syncwarp();
    if (lane == 0) s.nimp = sn_scan(s, xsize, P.asize, P.pmultmin, s.detout);
    __syncwarp();
    const int nimp = s.nimp;
    for (int pass = 0; pass < P.npasses; pass++) {
        for (int i = lane; i < xsize; i += 32) s.unfixed[i] = s.detout[i];
        __syncwarp();
        for (int k = 0; k < nimp; k++) {
            if (k > 0) { if (lane == 0) sn_scan(s, xsize, P.asize, P.pmultmin, s.unfixed); __syncwarp(); }
            const int nx = s.next, p = s.p_opt[nx], b0 = s.bimp[nx], ln = s.limp[nx];
            __syncwarp();
            if (p > 0) {
                sn_asolve(s, x, xsize, p, lane);
                sn_xhat(s, ln, p, x + b0 - p, ATAI, P1, lane);
                for (int i = lane; i < ln; i += 32) { x[b0 + i] = s.xh[i]; s.unfixed[b0 + i] = 0; }
            } else {
                for (int i = lane; i < ln; i += 32) x[b0 + i] = s.savex[b0 + i];
            }
            __syncwarp();
        }
    }
}

// `nframes` frames of every channel: the frame part of xsnba's while loop (snb.c:552-561)
__global__ void __launch_bounds__(32) snba_frames_kernel(SnbaPar P, SnbaLayout L, double *state, int nframes, int iaoutidx, int oainidx)
{
    extern __shared__ double sn_raw[];
    SnSm &s = *reinterpret_cast<SnSm *>(sn_raw);
    const int lane = threadIdx.x, xsize = P.xsize;
    double *st = state + (size_t)blockIdx.x * L.row;
    double *xb = st + L.xbase, *ina = st + L.inaccum, *outa = st + L.outaccum;
    int *dglob = reinterpret_cast<int *>(st + L.detout);
    for (int i = lane; i < 2 * xsize; i += 32) s.x[i] = xb[i];
    for (int i = lane; i < xsize; i += 32) s.detout[i] = dglob[i];     // entries below asize are never written: they keep what flush left (zeros)
    __syncwarp();
    for (int f = 0; f < nframes; f++) {
        for (int i = lane; i < P.incr; i += 32) { int j = iaoutidx + i; if (j >= P.iasize) j -= P.iasize; s.x[2 * xsize - P.incr + i] = ina[j]; }
        __syncwarp();
        sn_exec_frame(s, P, st + L.ATAI, st + L.P1, lane);
        iaoutidx += P.incr; if (iaoutidx >= P.iasize) iaoutidx -= P.iasize;
        for (int i = lane; i < P.incr; i += 32) { int j = oainidx + i; if (j >= P.oasize) j -= P.oasize; outa[j] = s.x[xsize + i]; }
        oainidx += P.incr; if (oainidx >= P.oasize) oainidx -= P.oasize;
        __syncwarp();
        // memmove(xbase, xbase + incr, 2 xsize - incr): through registers, eight elements per lane and round
        for (int i0 = 0; i0 < 2 * xsize - P.incr; i0 += 32) {
            const int i = i0 + lane;
            const double t = i < 2 * xsize - P.incr ? s.x[i + P.incr] : 0.0;
            __syncwarp();
            if (i < 2 * xsize - P.incr) s.x[i] = t;
            __syncwarp();
        }
    }
    for (int i = lane; i < 2 * xsize; i += 32) xb[i] = s.x[i];
    for (int i = lane; i < xsize; i += 32) dglob[i] = s.detout[i];
}

__global__ void snba_in_kernel(const cd *in, long is, int n, double *state, SnbaLayout L, int iainidx, int iasize)
{
    double *ina = state + (size_t)blockIdx.x * L.row + L.inaccum;
    const cd *x = in + (size_t)blockIdx.x * is;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { int j = iainidx + i; if (j >= iasize) j -= iasize; ina[j] = x[i].x; }
}
__global__ void snba_out_kernel(cd *out, long os, int n, const double *state, SnbaLayout L, int oaoutidx, int oasize)
{
    const double *outa = state + (size_t)blockIdx.x * L.row + L.outaccum;
    cd *y = out + (size_t)blockIdx.x * os;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { int j = oaoutidx + i; if (j >= oasize) j -= oasize; y[i] = make_double2(outa[j], 0.0); }
}

struct Snba {
    int C = 0, inrate = 0, internalrate = 0, bsize = 0;
    SnbaPar P;
    SnbaLayout L;
    double *d_state = nullptr;
    cd *d_inbuff = nullptr, *d_outbuff = nullptr;
    Resampler *inres = nullptr, *outres = nullptr;
    double out_low_cut = 0, out_high_cut = 0, out_fc_low = 200.0, out_fc = 0.0;        // the output resampler's current band (setFCLow 200, fc 0 = 0.45 min rate)
    int iainidx = 0, iaoutidx = 0, nsamps = 0, oainidx = 0, oaoutidx = 0, init_oaoutidx = 0;

    int init(int C_, int inrate_, int internalrate_, int bsize_, int ovrlp, int xsize, int asize, int npasses, double k1, double k2, int b,
             int pre, int post, double pmultmin, double out_low, double out_high)
    {   // create_snba + calc_snba, snb.c:31-119
        C = C_; inrate = inrate_; internalrate = internalrate_; bsize = bsize_; out_low_cut = out_low; out_high_cut = out_high;
        if (C <= 0 || bsize <= 0 || xsize != SN_XMAX || asize < 1 || asize > SN_AMAX || ovrlp < 1 || xsize % ovrlp || pmultmin < 0.5 || npasses < 0) {
            set_error("snba_create: xsize must be 256 (create_rxa's), asize <= 64, ovrlp a divisor of xsize, pmultmin >= 0.5");
            return QC_EINVAL;
        }
        memset(&P, 0, sizeof(P));
        P.xsize = xsize; P.ovrlp = ovrlp; P.asize = asize; P.npasses = npasses; P.k1 = k1; P.k2 = k2; P.b = b; P.pre = pre; P.post = post; P.pmultmin = pmultmin;
        if (inrate >= internalrate) { if (inrate % internalrate || bsize % (inrate / internalrate)) { set_error("snba_create: rates %d / %d and size %d do not divide", inrate, internalrate, bsize); return QC_EINVAL; }
                                      P.isize = bsize / (inrate / internalrate); }
        else P.isize = bsize * (internalrate / inrate);
        P.incr = xsize / ovrlp;
        P.iasize = P.incr > P.isize ? P.incr : P.isize;
        if (P.incr > P.isize) { P.oasize = P.incr; oaoutidx = P.isize; } else { P.oasize = P.isize; oaoutidx = 0; }
        init_oaoutidx = oaoutidx;
        size_t o = 0;
        auto take = [&](size_t n) { const size_t at = o; o += (n + 1) & ~(size_t)1; return at; };
        L.xbase = take(2 * xsize); L.inaccum = take(P.iasize); L.outaccum = take(P.oasize); L.detout = take(xsize / 2 + 1);
        L.ATAI = take((size_t)SN_UMAX * SN_UMAX); L.P1 = take((size_t)SN_UMAX * (SN_UMAX + 2 * SN_AMAX));
        L.row = o;
        QC_CUDA(cudaMalloc((void **)&d_state, (size_t)C * L.row * sizeof(double)));
        QC_CUDA(cudaMemset(d_state, 0, (size_t)C * L.row * sizeof(double)));
        if (inrate != internalrate) {
            QC_CUDA(cudaMalloc((void **)&d_inbuff, (size_t)C * (P.isize + 8) * sizeof(cd)));
            QC_CUDA(cudaMalloc((void **)&d_outbuff, (size_t)C * (P.isize + 8) * sizeof(cd)));
            inres = new Resampler(); outres = new Resampler();
            // create_resample(..., fc 0.0, ncoef 0, gain 2.0) + setFCLow_resample(250 / 200), snb.c:42-45
            int rc = inres->init_band(C, inrate, internalrate, 250.0, 0.0, 0, 2.0); if (rc != QC_OK) return rc;
            rc = outres->init_band(C, internalrate, inrate, 200.0, 0.0, 0, 2.0); if (rc != QC_OK) return rc;
        }
        return QC_OK;
    }
    void release()
    {
        if (d_state) cudaFree(d_state); if (d_inbuff) cudaFree(d_inbuff); if (d_outbuff) cudaFree(d_outbuff);
        d_state = nullptr; d_inbuff = d_outbuff = nullptr;
        for (Resampler **r : {&inres, &outres}) if (*r) { (*r)->release(); delete *r; *r = nullptr; }
    }
    int flush()
    {   // flush_snba, snb.c:161-185: the accumulators, the frame (xaux) and the work arrays; the history half of xbase stays
        QC_CUDA(cudaDeviceSynchronize());
        for (int c = 0; c < C; c++) {
            double *st = d_state + (size_t)c * L.row;
            QC_CUDA(cudaMemset(st + L.inaccum, 0, (size_t)P.iasize * sizeof(double)));
            QC_CUDA(cudaMemset(st + L.outaccum, 0, (size_t)P.oasize * sizeof(double)));
            QC_CUDA(cudaMemset(st + L.xbase + P.xsize, 0, (size_t)P.xsize * sizeof(double)));
            QC_CUDA(cudaMemset(st + L.detout, 0, (size_t)(P.xsize / 2 + 1) * sizeof(double)));
        }
        iainidx = iaoutidx = nsamps = oainidx = 0; oaoutidx = init_oaoutidx;
        for (Resampler *r : {inres, outres}) if (r) { int rc = r->f->reset(nullptr); if (rc != QC_OK) return rc; }
        return QC_OK;
    }
    int run(const cd *d_in, long is, cd *d_out, long os, cudaStream_t s)
    {   // xsnba with run = 1, snb.c:539-572
        const cd *src = d_in; long ss = is;
        if (inres) {
            int no = 0;
            int rc = inres->f->run(d_in, is, bsize, d_inbuff, P.isize + 8, &no, 0, s); if (rc != QC_OK) return rc;
            if (no != P.isize) { set_error("snba: input resampler produced %d samples, expected %d", no, P.isize); return QC_EINVAL; }
            src = d_inbuff; ss = P.isize + 8;
        }
        snba_in_kernel<<<C, 64, 0, s>>>(src, ss, P.isize, d_state, L, iainidx, P.iasize);
        count_launch();
        iainidx = (iainidx + P.isize) % P.iasize;
        nsamps += P.isize;
        int nframes = 0;
        for (int n = nsamps; n >= P.incr; n -= P.incr) nframes++;
        if (nframes > 0) {
            QC_CUDA(cudaFuncSetAttribute(snba_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SnSm)));
            snba_frames_kernel<<<C, 32, sizeof(SnSm), s>>>(P, L, d_state, nframes, iaoutidx, oainidx);
            count_launch();
            iaoutidx = (iaoutidx + nframes * P.incr) % P.iasize;
            oainidx = (oainidx + nframes * P.incr) % P.oasize;
            nsamps -= nframes * P.incr;
        }
        cd *dst = outres ? d_outbuff : d_out; const long ds = outres ? P.isize + 8 : os;
        snba_out_kernel<<<C, 64, 0, s>>>(dst, ds, P.isize, d_state, L, oaoutidx, P.oasize);
        count_launch();
        QC_CUDA_LAUNCH();
        oaoutidx = (oaoutidx + P.isize) % P.oasize;
        if (outres) {
            int no = 0;
            int rc = outres->f->run(d_outbuff, P.isize + 8, P.isize, d_out, os, &no, 0, s); if (rc != QC_OK) return rc;
            if (no != bsize) { set_error("snba: output resampler produced %d samples, expected %d", no, bsize); return QC_EINVAL; }
        }
        return QC_OK;
    }
};

int snba_set_output_bandwidth(Snba *d, double flow, double fhigh)
{   // SetRXASNBAOutputBandwidth, snb.c:660-696: the output resampler becomes a band pass inside [out_low_cut, out_high_cut]
    double f_low = 0.0, f_high = 0.0;
    bool set = false;
    auto mx = [](double a, double b) { return a > b ? a : b; };
    auto mn = [](double a, double b) { return a < b ? a : b; };
    if (flow >= 0 && fhigh >= 0) {
        if (fhigh < d->out_low_cut) fhigh = d->out_low_cut;
        if (flow > d->out_high_cut) flow = d->out_high_cut;
        f_low = mx(d->out_low_cut, flow); f_high = mn(d->out_high_cut, fhigh); set = true;
    } else if (flow <= 0 && fhigh <= 0) {
        if (flow > -d->out_low_cut) flow = -d->out_low_cut;
        if (fhigh < -d->out_high_cut) fhigh = -d->out_high_cut;
        f_low = mx(d->out_low_cut, -fhigh); f_high = mn(d->out_high_cut, -flow); set = true;
    } else if (flow < 0 && fhigh > 0) {
        double absmax = mx(-flow, fhigh);
        if (absmax < d->out_low_cut) absmax = d->out_low_cut;
        f_low = d->out_low_cut; f_high = mn(d->out_high_cut, absmax); set = true;
    }
    (void)set;      // (the reference passes whatever f_low / f_high hold: both zero when no branch applied)
    if (!d->outres) return QC_OK;                           // no resamplers at the internal rate: setBandwidth_resample only touches the (unused) design
    if (f_low == d->out_fc_low && f_high == d->out_fc) return QC_OK;
    if (cudaDeviceSynchronize() != cudaSuccess) return QC_ECUDA;
    Resampler *nr = new Resampler();
    int rc = nr->init_band(d->C, d->internalrate, d->inrate, f_low, f_high, 0, 2.0);       // setBandwidth_resample: decalc + calc, state zeroed
    if (rc != QC_OK) { nr->release(); delete nr; return rc; }
    d->outres->release(); delete d->outres;
    d->outres = nr; d->out_fc_low = f_low; d->out_fc = f_high;
    return QC_OK;
}

Snba *make_snba(int C, int inrate, int internalrate, int bsize, int ovrlp, int xsize, int asize, int npasses, double k1, double k2, int b,
                int pre, int post, double pmultmin, double out_low, double out_high)
{
    Snba *d = new Snba();
    if (d->init(C, inrate, internalrate, bsize, ovrlp, xsize, asize, npasses, k1, k2, b, pre, post, pmultmin, out_low, out_high) != QC_OK) { d->release(); delete d; return nullptr; }
    return d;
}
void snba_destroy(Snba *d) { if (d) { d->release(); delete d; } }
int snba_run(Snba *d, const cd *in, long is, cd *out, long os, cudaStream_t s) { return d->run(in, is, out, os, s); }
int snba_flush(Snba *d) { return d->flush(); }

}  // namespace qc

struct qcSnba { qc::Snba *d; };

extern "C" {

qcSnba *quisk_cuda_snba_create(int n_channels, int inrate, int internalrate, int bsize, int ovrlp, int xsize, int asize, int npasses,
                               double k1, double k2, int b, int pre, int post, double pmultmin, double out_low_cut, double out_high_cut)
{
    if (qc::ensure_device() != QC_OK) return nullptr;
    qc::Snba *d = qc::make_snba(n_channels, inrate, internalrate, bsize, ovrlp, xsize, asize, npasses, k1, k2, b, pre, post, pmultmin, out_low_cut, out_high_cut);
    if (!d) return nullptr;
    qcSnba *h = new qcSnba();
    h->d = d;
    return h;
}
void quisk_cuda_snba_destroy(qcSnba *h) { if (h) { qc::snba_destroy(h->d); delete h; } }
int quisk_cuda_snba_run(qcSnba *h, const void *d_in, long in_stride, void *d_out, long out_stride, void *stream)
{
    if (!h || !d_in || !d_out) { qc::set_error("snba_run: bad arguments"); return QC_EINVAL; }
    return h->d->run((const double2 *)d_in, in_stride, (double2 *)d_out, out_stride, (cudaStream_t)stream);
}
int quisk_cuda_snba_flush(qcSnba *h) { return h ? h->d->flush() : QC_EINVAL; }

}  // extern "C"
