// quisk_b200/csrc/nco_device.cuh -- closed-form evaluation of the tuning recurrence
// (quisk.c:2477-2488) on the device.  See nco_host.cpp for the constants.
#pragma once
#include "qc_common.cuh"

namespace qc {

// complex multiply with separately rounded products, the sequence gcc emits for
// `a * b` on baseline x86-64 (no FMA contraction)
__device__ __forceinline__ cd cmul_rn(cd a, cd b)
{
    return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)),
                        __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}

// phase^n for the rounded per-sample phase described by nco[0..2]
// (frac_hi, frac_lo: turns per sample as a double-double; growth: log|phase|).
// n < 2^53.  The product n*frac is formed exactly (p + e) before the integer
// part is dropped, so the phase error does not grow with n.
__device__ __forceinline__ cd nco_pow(const double *nco, unsigned long long n)
{
    const double dn = (double)n;
    const double fh = nco[0], fl = nco[1];
    const double p = dn * fh;
    const double e = fma(dn, fh, -p);
    double t = p - rint(p);
    t += e + dn * fl;
    double s, c;
    sincospi(2.0 * t, &s, &c);
    const double amp = exp(dn * nco[2]);
    return make_double2(c * amp, s * amp);
}

}  // namespace qc
