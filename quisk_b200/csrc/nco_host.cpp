// quisk_b200/csrc/nco_host.cpp -- host-side constants for the tuning NCO.
//
// The reference tunes with a recurrence (quisk.c:2477-2488):
//     phase = cexp((I * -2.0 * M_PI * tune) / sample_rate);
//     for each sample: x *= v; v *= phase;           (v static, starts at 1)
// so after n samples v = v0 * phase^n for the *rounded* phase (|phase| is not
// exactly 1 and its argument is not exactly -2 pi tune / rate).  A GPU cannot
// run a length-n scalar recurrence per sample, so the kernels evaluate
// v0 * phase^n in closed form.  To follow the reference rather than the ideal
// oscillator the closed form uses the logarithm of the rounded phase, computed
// here once per (re)tune in binary128:
//     frac   = arg(phase) / (2 pi)      as a double-double (turns per sample)
//     growth = log |phase|              (about 1e-17 per sample)
// Compiled by g++ (not nvcc) because of __float128.
#include <quadmath.h>
#include <cmath>

namespace qc {

void nco_make(double tune_hz, int sample_rate, double v0_re, double v0_im, double out[8])
{
    // same operation order as the reference expression above
    const double y = ((-2.0 * M_PI) * tune_hz) / (double)sample_rate;
    const double pr = std::cos(y), pi = std::sin(y);
    const __float128 qr = pr, qi = pi;
    const __float128 two_pi = 2 * M_PIq;
    __float128 frac = atan2q(qi, qr) / two_pi;            // turns, in (-0.5, 0.5]
    const double hi = (double)frac;
    const double lo = (double)(frac - (__float128)hi);
    const __float128 mag2 = qr * qr + qi * qi;
    const double growth = (double)(logq(mag2) / 2);
    out[0] = hi;
    out[1] = lo;
    out[2] = growth;
    out[3] = v0_re;
    out[4] = v0_im;
    out[5] = pr;
    out[6] = pi;
    out[7] = 0.0;
}

void nco_advance(const double nco[8], unsigned long long n, double *v_re, double *v_im)
{
    const __float128 frac = (__float128)nco[0] + (__float128)nco[1];
    __float128 turns = frac * (__float128)n;
    turns -= floorq(turns);
    const __float128 ang = turns * 2 * M_PIq;
    const __float128 amp = expq((__float128)nco[2] * (__float128)n);
    const __float128 c = cosq(ang) * amp, s = sinq(ang) * amp;
    const __float128 vr = (__float128)nco[3], vi = (__float128)nco[4];
    *v_re = (double)(vr * c - vi * s);
    *v_im = (double)(vr * s + vi * c);
}

}  // namespace qc
