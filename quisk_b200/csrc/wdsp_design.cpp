// quisk_b200/csrc/wdsp_design.cpp -- host-side coefficient design for the WDSP RXA stages.
//
// Filter design stays on the CPU (taps must be bit-identical to what the
// reference's libm produces, and they are computed once per retune).  Restated from
//   fir_bandpass        wdsp/fir.c:187-254   windowed sinc (Blackman-Harris 4 / 7 term), complex tuned
//   get_fsamp_window    wdsp/fir.c:44-82
//   fir_fsamp(_odd)     wdsp/fir.c:84-180    frequency-sampling design
//   fc_impulse          wdsp/fcurve.c:29-143 FM (de-)emphasis curve
//   calc_resample       wdsp/resample.c:35-78 L/M, tap count and prototype of the rational resampler
//   make_nbp, fir_mbandpass, min_notch_width  wdsp/nbp.c:64-179   notched band-pass (notch database -> pass bands -> taps)
//   analytic, mp_imp    wdsp/fir.c:292-368   minimum-phase version of an impulse response (cepstral method); the
//                       reference runs its three transforms through FFTW, here a plain radix-2 FFT does -- the
//                       method takes log|H| of stop-band bins at the rounding floor, so two correct FFTs give
//                       taps that agree to ~1e-6, not to 1e-13: that is the algorithm's conditioning
// with the same expression order, so the doubles that come out are the reference's.
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/quisk_cuda_wdsp.h"

namespace {

const double kPI = 3.1415926535897932;       // wdsp/comm.h:146-147
const double kTWOPI = 6.2831853071795864;

// unnormalised radix-2 FFT (sign -1 forward, +1 backward), design-time only
void host_fft(std::vector<std::complex<double>> &a, int sign)
{
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len >> 1;
        for (size_t k = 0; k < half; k++) {
            const double ang = sign * kTWOPI * (double)k / (double)len;
            const std::complex<double> w(std::cos(ang), std::sin(ang));
            for (size_t i = k; i < n; i += len) {
                const std::complex<double> u = a[i], v = a[i + half] * w;
                a[i] = u + v; a[i + half] = u - v;
            }
        }
    }
}

double bh_window(int wintype, double cosphi)
{
    if (wintype == 0)           // Blackman-Harris 4-term
        return +0.21747 + cosphi * (-0.45325 + cosphi * (+0.28256 + cosphi * (-0.04672)));
    return +6.3964424114390378e-02 + cosphi * (-2.3993864599352804e-01 + cosphi * (+3.5015956323820469e-01
         + cosphi * (-2.4774111897080783e-01 + cosphi * (+8.5438256055858031e-02 + cosphi * (-1.2320203369293225e-02
         + cosphi * (+4.3778825791773474e-04))))));
}

void fsamp_window(int N, int wintype, std::vector<double> &w)
{
    w.assign((size_t)N, 1.0);
    if (wintype != 0 && wintype != 1) return;
    const double arg0 = 2.0 * kPI / ((double)N - 1.0);
    for (int i = 0; i < N; i++) w[i] = bh_window(wintype, std::cos(arg0 * (double)i));
}

// fir_fsamp / fir_fsamp_odd with rtype 1 (interleaved complex, imaginary parts zero)
void fsamp_complex(int N, const std::vector<double> &A, double scale, int wintype, double *out)
{
    std::vector<double> re((size_t)N, 0.0);
    if (N & 1) {
        // fir_fsamp_odd: inverse DFT of a Hermitian, linear-phase spectrum (the reference uses FFTW here;
        // a direct sum is the same mathematics -- agreement is to FFT rounding, ~1e-16 relative)
        const int mid = (N - 1) / 2;
        const double local_scale = 1.0 / (double)N;
        std::vector<double> fr((size_t)N), fi((size_t)N);
        for (int i = 0; i <= mid; i++) {
            const double mag = A[i] * local_scale;
            const double phs = -(double)mid * kTWOPI * (double)i / (double)N;
            fr[i] = mag * std::cos(phs); fi[i] = mag * std::sin(phs);
        }
        for (int i = mid + 1, j = 0; i < N; i++, j++) { fr[i] = +fr[mid - j]; fi[i] = -fi[mid - j]; }
        for (int n = 0; n < N; n++) {
            double s = 0.0;
            for (int k = 0; k < N; k++) {
                const long idx = ((long)n * k) % N;
                const double a = kTWOPI * (double)idx / (double)N;
                s += fr[k] * std::cos(a) - fi[k] * std::sin(a);
            }
            re[n] = s;
        }
    } else {
        const double M = (double)(N - 1) / 2.0;
        for (int n = 0; n < N / 2; n++) {
            double sum = 0.0;
            for (int k = 1; k < N / 2; k++) sum += 2.0 * A[k] * std::cos(kTWOPI * (n - M) * k / N);
            re[n] = (1.0 / N) * (A[0] + sum);
        }
        for (int n = N / 2, j = 1; n < N; n++, j++) re[n] = re[N / 2 - j];
    }
    std::vector<double> w;
    fsamp_window(N, wintype, w);
    for (int i = 0; i < N; i++) { out[2 * i] = re[i] * (scale * w[i]); out[2 * i + 1] = 0.0; }
}

}  // namespace

extern "C" {

int quisk_cuda_fir_bandpass(int N, double f_low, double f_high, double samplerate, int wintype, int rtype,
                            double scale, double *out)
{
    if (N < 2 || !out || (rtype != 0 && rtype != 1)) return QC_EINVAL;
    memset(out, 0, sizeof(double) * (size_t)N * (rtype ? 2 : 1));
    const double ft = (f_high - f_low) / (2.0 * samplerate);
    const double ft_rad = kTWOPI * ft;
    const double w_osc = kPI * (f_high + f_low) / samplerate;
    const double m = 0.5 * (double)(N - 1);
    const double delta = kPI / m;
    if (N & 1) {
        if (rtype == 0) out[N >> 1] = scale * 2.0 * ft;
        else { out[N - 1] = scale * 2.0 * ft; out[N] = 0.0; }
    }
    for (int i = (N + 1) / 2; i < N; i++) {
        const int j = N - 1 - i;                         // the mirrored tap shares sinc and window
        const double posi = (double)i - m, posj = (double)j - m;
        const double sinc = std::sin(ft_rad * posi) / (kPI * posi);
        const double window = bh_window(wintype, std::cos(delta * i));
        const double coef = scale * sinc * window;
        if (rtype == 0) {
            out[i] = +coef * std::cos(posi * w_osc);
            out[j] = +coef * std::cos(posj * w_osc);
        } else {
            out[2 * i + 0] = +coef * std::cos(posi * w_osc);
            out[2 * i + 1] = -coef * std::sin(posi * w_osc);
            out[2 * j + 0] = +coef * std::cos(posj * w_osc);
            out[2 * j + 1] = -coef * std::sin(posj * w_osc);
        }
    }
    return QC_OK;
}

int quisk_cuda_fc_impulse(int nc, double f0, double f1, double g0, double g1, int curve, double samplerate,
                          double scale, int ctfmode, int wintype, double *out)
{
    (void)g1;
    if (nc < 4 || !out) return QC_EINVAL;
    const int mid = nc / 2;
    std::vector<double> A((size_t)mid + 2, 0.0);
    const double g0_lin = std::pow(10.0, g0 / 20.0);
    const bool odd = nc & 1;
    const int na = odd ? mid + 1 : mid;
    for (int i = 0; i < na; i++) {
        const double fn = odd ? (double)i / (double)mid : ((double)i + 0.5) / (double)mid;
        const double f = fn * samplerate / 2.0;
        if (curve == 0) A[i] = f0 > 0.0 ? scale * (g0_lin * f / f0) : 0.0;       // pre-emphasis
        else A[i] = f > 0.0 ? scale * (g0_lin * f0 / f) : 0.0;                  // de-emphasis
    }
    if (ctfmode == 0) {
        int low, high;
        if (odd) { low = (int)(2.0 * f0 / samplerate * mid); high = (int)(2.0 * f1 / samplerate * mid + 0.5); }
        else { low = (int)(2.0 * f0 / samplerate * mid - 0.5); high = (int)(2.0 * f1 / samplerate * mid - 0.5); }
        double lowmag = A[low], highmag = A[high];
        const double flow4 = std::pow((double)low / (double)mid, 4.0);
        const double fhigh4 = std::pow((double)high / (double)mid, 4.0);
        for (int k = low - 1; k >= 0; k--) {
            const double f = (double)k / (double)mid;
            lowmag *= (f * f * f * f) / flow4;
            if (lowmag < 1.0e-100) lowmag = 1.0e-100;
            A[k] = lowmag;
        }
        const int top = odd ? mid : mid - 1;
        for (int k = high + 1; k <= top; k++) {
            const double f = (double)k / (double)mid;
            highmag *= fhigh4 / (f * f * f * f);
            if (highmag < 1.0e-100) highmag = 1.0e-100;
            A[k] = highmag;
        }
    }
    fsamp_complex(nc, A, 1.0, wintype, out);
    return QC_OK;
}

/* calc_resample (resample.c:35-78) with its fc_low member: < 0 = low pass (what create_resample leaves, resample.c:92), >= 0 = the
 * band pass setFCLow_resample / setBandwidth_resample make of it (resample.c:185-207; the noise blanker's two converters use it) */
int quisk_cuda_resample_design_band(int in_rate, int out_rate, double fc_low, double fc, int ncoef_in, double gain,
                                    int *pL, int *pM, int *pncoef, double *h, int h_cap)
{
    if (in_rate <= 0 || out_rate <= 0) return QC_EINVAL;
    int x = in_rate, y = out_rate;
    while (y != 0) { const int z = y; y = x % y; x = z; }
    const int L = out_rate / x, M = in_rate / x;
    const int min_rate = in_rate < out_rate ? in_rate : out_rate;
    if (fc == 0.0) fc = 0.45 * (double)min_rate;
    const double full_rate = (double)(in_rate * L);
    const double fc_norm_high = fc / full_rate;
    const double fc_norm_low = fc_low < 0.0 ? -fc_norm_high : fc_low / full_rate;
    int ncoef = ncoef_in;
    if (ncoef == 0) ncoef = (int)(140.0 * full_rate / min_rate);
    ncoef = (ncoef / L + 1) * L;
    if (pL) *pL = L;
    if (pM) *pM = M;
    if (pncoef) *pncoef = ncoef;
    if (h) {
        if (h_cap < ncoef) return QC_EINVAL;
        // the prototype in natural order: tap j + k*L is phase j, index k -- exactly the layout the
        // streaming polyphase kernel indexes (coef[ph + k*L]); resample.c:70-73 merely regroups it
        return quisk_cuda_fir_bandpass(ncoef, fc_norm_low, fc_norm_high, 1.0, 1, 0, gain * (double)L, h);
    }
    return QC_OK;
}

int quisk_cuda_resample_design(int in_rate, int out_rate, double fc, int ncoef_in, double gain,
                               int *pL, int *pM, int *pncoef, double *h, int h_cap)
{
    return quisk_cuda_resample_design_band(in_rate, out_rate, -1.0, fc, ncoef_in, gain, pL, pM, pncoef, h, h_cap);
}


/* calc_nbp_impulse with fnfrun = 1 (nbp.c:214-239): the pass band [flow, fhigh] is cut up by the active notches
 * (make_nbp, nbp.c:97-179, working in RF coordinates: offset = tunefreq + shift), the pieces are designed with
 * fir_bandpass and summed (fir_mbandpass, nbp.c:64-80).  Same statement order as the reference, so the taps are
 * the reference's doubles. */
int quisk_cuda_nbp_impulse(int nc, double flow, double fhigh, double rate, int wintype, double scale,
                           int n_notches, const double *fcenter, const double *fwidth, const int *active,
                           double tunefreq, double shift, int autoincr, int maxpb,
                           double *impulse, int *numpb_out, int *havnotch_out)
{
    if (nc < 2 || !impulse || maxpb < 1 || n_notches < 0 || (n_notches > 0 && (!fcenter || !fwidth || !active))) return QC_EINVAL;
    if (wintype != 0 && wintype != 1) return QC_EINVAL;
    // min_notch_width, nbp.c:82-95 (nc / 256 is an integer division there)
    const double minwidth = (wintype == 0 ? 1600.0 : 2200.0) / (nc / 256) * (rate / 48000);
    const double offset = tunefreq + shift;
    const double fl = flow + offset, fh = fhigh + offset;
    std::vector<double> bplow((size_t)maxpb + 1, 0.0), bphigh((size_t)maxpb + 1, 0.0);
    std::vector<int> del(1024 + (size_t)maxpb, 0);
    int nbp = 0, havnotch = 0;
    if (fh > fl) {
        bplow[0] = fl; bphigh[0] = fh; nbp = 1;
        for (int k = 0; k < n_notches; k++) {
            double nl, nh;
            if (autoincr && fwidth[k] < minwidth) { nl = fcenter[k] - 0.5 * minwidth; nh = fcenter[k] + 0.5 * minwidth; }
            else { nl = fcenter[k] - 0.5 * fwidth[k]; nh = fcenter[k] + 0.5 * fwidth[k]; }     // nlow / nhigh of RXANBPAddNotch, nbp.c:378-379
            if (active[k] && (nh > fl && nl < fh)) {
                havnotch = 1;
                int adds = 0;
                for (int i = 0; i < nbp; i++) {
                    if (nh > bplow[i] && nl < bphigh[i]) {
                        if (nl <= bplow[i] && nh >= bphigh[i]) del[i] = 1;
                        else if (nl > bplow[i] && nh < bphigh[i]) {
                            if (nbp + adds >= maxpb) return QC_EINVAL;
                            bplow[nbp + adds] = nh; bphigh[nbp + adds] = bphigh[i]; bphigh[i] = nl; adds++;
                        }
                        else if (nl <= bplow[i] && nh > bplow[i]) bplow[i] = nh;
                        else if (nl < bphigh[i] && nh >= bphigh[i]) bphigh[i] = nl;
                    }
                }
                nbp += adds;
                int nnbp = nbp;
                for (int i = 0; i < nbp; i++) {
                    if (del[i] == 1) {
                        nnbp--;
                        for (int j = i; j < nnbp; j++) { bplow[j] = bplow[j + 1]; bphigh[j] = bphigh[j + 1]; }
                        del[i] = 0;
                    }
                }
                nbp = nnbp;
            }
        }
    }
    for (int i = 0; i < nbp; i++) { bplow[i] -= offset; bphigh[i] -= offset; }
    memset(impulse, 0, sizeof(double) * 2 * (size_t)nc);
    std::vector<double> imp((size_t)2 * nc);
    for (int k = 0; k < nbp; k++) {
        int rc = quisk_cuda_fir_bandpass(nc, bplow[k], bphigh[k], rate, wintype, 1, scale, imp.data());
        if (rc != QC_OK) return rc;
        for (int i = 0; i < nc; i++) { impulse[2 * i] += imp[2 * i]; impulse[2 * i + 1] += imp[2 * i + 1]; }
    }
    if (numpb_out) *numpb_out = nbp;
    if (havnotch_out) *havnotch_out = havnotch;
    return QC_OK;
}

/* mp_imp (fir.c:317-368) with `analytic` (fir.c:292-315): minimum-phase impulse of the same magnitude response.
 * N complex taps in, N out; pfactor = zero-padding factor (the reference uses 16, firmin.c:328); N * pfactor must be
 * a power of two. */
int quisk_cuda_mp_imp(int N, const double *fir, double *mpfir, int pfactor, int polarity)
{
    if (N < 1 || pfactor < 1 || !fir || !mpfir) return QC_EINVAL;
    const long size = (long)N * pfactor;
    if (size & (size - 1)) return QC_EINVAL;
    std::vector<std::complex<double>> firfreq((size_t)size), ana((size_t)size), newfreq((size_t)size);
    std::vector<double> mag((size_t)size);
    const double inv_PN = 1.0 / (double)size;
    for (long i = 0; i < size; i++) firfreq[i] = i < N ? std::complex<double>(fir[2 * i], fir[2 * i + 1]) : std::complex<double>(0.0, 0.0);
    host_fft(firfreq, -1);
    for (long i = 0; i < size; i++) {
        mag[i] = std::sqrt(firfreq[i].real() * firfreq[i].real() + firfreq[i].imag() * firfreq[i].imag()) * inv_PN;
        ana[i] = std::complex<double>(mag[i] > 0.0 ? std::log(mag[i]) : std::log(1.0e-300), 0.0);
    }
    // analytic(size, ana, ana): one-sided spectrum of log|H| -> its imaginary part is minus the minimum phase
    {
        const double inv_N = 1.0 / (double)size, two_inv_N = 2.0 * inv_N;
        host_fft(ana, -1);
        ana[0] *= inv_N;
        for (long i = 1; i < size / 2; i++) ana[i] *= two_inv_N;
        ana[size / 2] *= inv_N;
        for (long i = size / 2 + 1; i < size; i++) ana[i] = 0.0;
        host_fft(ana, +1);
    }
    for (long i = 0; i < size; i++) {
        const double re = +mag[i] * std::cos(ana[i].imag());
        const double im = mag[i] * std::sin(ana[i].imag());
        newfreq[i] = std::complex<double>(re, polarity ? +im : -im);
    }
    host_fft(newfreq, +1);
    const long off = polarity ? (long)(pfactor - 1) * N : 0;
    for (long i = 0; i < N; i++) { mpfir[2 * i] = newfreq[off + i].real(); mpfir[2 * i + 1] = newfreq[off + i].imag(); }
    return QC_OK;
}
}  // extern "C"
