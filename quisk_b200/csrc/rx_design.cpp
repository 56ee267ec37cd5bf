// quisk_b200/csrc/rx_design.cpp -- host-side design of the main receive filter taps that cRxFilterOut /
// dRxFilterOut consume (set_filters, quisk.c:4551-4594).
//
// quisk_cuda_make_filter_coef follows MakeFilterCoef (quisk.py:5405-5456): a real low-pass prototype -- one of the
// 24 kS/s tables of filters.py when the key bw * 24000 // rate // 2 names one (the caller passes it), otherwise a
// Blackman-windowed periodic-sinc (Dirichlet) kernel of N taps -- tuned to `center` by 2 exp(-j 2 pi center / rate
// (i - D)), D = (NN - 1) / 2, which gives the I taps (real part) and the Q taps (imaginary part).  The reference
// evaluates this in Python floats = C doubles through the same libm (math.sin / math.cos, cmath.exp), and complex x
// float multiplies component-wise there, so the expressions below keep Python's operand order and the taps are
// bit-identical (tests/test_rx_design.py compares with the reference module itself).
#include <cmath>
#include <vector>
#include "../../include/quisk_cuda.h"

extern "C" int quisk_cuda_filter_key(int rate, int bw)
{   // the Filters[] key MakeFilterCoef looks up (quisk.py:5408); Python's // floors, bw and rate are positive here
    if (rate <= 0 || bw < 0) return -1;
    return (int)(((long long)bw * 24000 / rate) / 2);
}

extern "C" int quisk_cuda_make_filter_coef(int rate, int N, int bw, int center, const double *proto, int n_proto,
                                           double *filt_i, double *filt_q, int cap, int *n_taps)
{
    if (rate <= 0 || bw <= 0 || !filt_i || !filt_q || !n_taps) return QC_EINVAL;
    if (center < 0) center = -center;
    std::vector<double> lp;
    if (proto && n_proto > 0) {
        lp.assign(proto, proto + n_proto);
    } else {
        if (N <= 0) {       // N is None: size from the 88 dB shape factor 1.5 (quisk.py:5414-5420)
            const double trans = ((double)bw / 2.0 / (double)rate) * (1.5 - 1.0);
            N = (int)(4.0 / trans);
            if (N > 1000) N = 1000;
            N = (N / 2) * 2 + 1;
        }
        const long long K = (long long)bw * N / rate;
        // range(-N//2, N//2 + 1) with Python's floor division: for odd N that is -(N+1)/2 .. (N-1)/2, N + 1 taps
        const int k_lo = (N >= 0) ? -((N + 1) / 2) : 0, k_hi = N / 2;
        lp.reserve((size_t)(k_hi - k_lo + 1));
        const double pi = M_PI;
        for (int k = k_lo; k <= k_hi; k++) {
            double z;
            if (k == 0) z = (double)K / (double)N;
            else z = 1.0 / (double)N * std::sin(pi * (double)k * (double)K / (double)N) / std::sin(pi * (double)k / (double)N);
            const double w = 0.42 + 0.5 * std::cos(2. * pi * (double)k / (double)N) + 0.08 * std::cos(4. * pi * (double)k / (double)N);
            lp.push_back(z * w);
        }
    }
    const int NN = (int)lp.size();
    *n_taps = NN;
    if (NN > cap) return QC_ENOMEM;
    if (!center) {
        for (int i = 0; i < NN; i++) { filt_i[i] = lp[i]; filt_q[i] = lp[i]; }
        return QC_OK;
    }
    // tune = -1j * 2.0 * math.pi * center / rate: complex x float scales both parts, so the imaginary part is
    // (((-1 * 2.0) * pi) * center) / rate; exp(0 + jy) = (cos y, sin y); 2.0 * z and z * filtD[i] are component-wise
    const double timag = (((-1.0 * 2.0) * M_PI) * (double)center) / (double)rate;
    const double D = ((double)NN - 1.0) / 2.0;
    for (int i = 0; i < NN; i++) {
        const double y = timag * ((double)i - D);
        filt_i[i] = (2.0 * std::cos(y)) * lp[i];
        filt_q[i] = (2.0 * std::sin(y)) * lp[i];
    }
    return QC_OK;
}
