/* oracle/fft64.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * See fft64.h.  Stockham autosort FFT, radix-4 passes plus one radix-2 pass
 * when log2(n) is odd; twiddles come from a table computed directly with
 * cos/sin (no recurrences), so the error stays at the O(log n * eps) level
 * that any correct FP64 FFT (FFTW included) delivers.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "fft64.h"

struct fft64_plan_s {
    int n;
    int sign;
    int pow2;
    double *tw;   /* n entries (re,im): exp(sign*2*pi*i*k/n) */
};

fft64_plan *fft64_create(int n, int sign)
{
    fft64_plan *p = (fft64_plan *)calloc(1, sizeof(*p));
    int k;
    p->n = n;
    p->sign = sign < 0 ? -1 : 1;
    p->pow2 = (n > 0) && ((n & (n - 1)) == 0);
    p->tw = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    for (k = 0; k < n; k++) {
        double a = 2.0 * M_PI * (double)k / (double)n;
        p->tw[2 * k] = cos(a);
        p->tw[2 * k + 1] = p->sign * sin(a);
    }
    return p;
}

void fft64_destroy(fft64_plan *p)
{
    if (!p) return;
    free(p->tw);
    free(p);
}

static void dft_naive(const fft64_plan *p, const double *in, double *out)
{
    int n = p->n, k, j;
    double *tmp = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    for (k = 0; k < n; k++) {
        double sr = 0.0, si = 0.0;
        long idx = 0;
        for (j = 0; j < n; j++) {
            double wr = p->tw[2 * idx], wi = p->tw[2 * idx + 1];
            sr += in[2 * j] * wr - in[2 * j + 1] * wi;
            si += in[2 * j] * wi + in[2 * j + 1] * wr;
            idx += k;
            if (idx >= n) idx -= n;
        }
        tmp[2 * k] = sr;
        tmp[2 * k + 1] = si;
    }
    memcpy(out, tmp, sizeof(double) * 2 * (size_t)n);
    free(tmp);
}

void fft64_execute(const fft64_plan *p, const double *in, double *out)
{
    int n = p->n;
    if (n == 1) { out[0] = in[0]; out[1] = in[1]; return; }
    if (!p->pow2) { dft_naive(p, in, out); return; }

    double *buf = (double *)malloc(sizeof(double) * 4 * (size_t)n);
    double *x = buf, *y = buf + 2 * (size_t)n;
    const double *tw = p->tw;
    const double sg = (double)p->sign;   /* -1 forward, +1 backward */
    int len = n;      /* current sub-transform length */
    int s = 1;        /* stride */
    memcpy(x, in, sizeof(double) * 2 * (size_t)n);

    while (len >= 4) {
        int n1 = len / 4, n2 = len / 2, n3 = n1 + n2;
        int tstep = n / len;
        for (int pp = 0; pp < n1; pp++) {
            double w1r = tw[2 * (pp * tstep)],     w1i = tw[2 * (pp * tstep) + 1];
            double w2r = tw[2 * (2 * pp * tstep)], w2i = tw[2 * (2 * pp * tstep) + 1];
            double w3r = tw[2 * (3 * pp * tstep)], w3i = tw[2 * (3 * pp * tstep) + 1];
            for (int q = 0; q < s; q++) {
                const double *a = x + 2 * (size_t)(q + s * (pp));
                const double *b = x + 2 * (size_t)(q + s * (pp + n1));
                const double *c = x + 2 * (size_t)(q + s * (pp + n2));
                const double *d = x + 2 * (size_t)(q + s * (pp + n3));
                double apcr = a[0] + c[0], apci = a[1] + c[1];
                double amcr = a[0] - c[0], amci = a[1] - c[1];
                double bpdr = b[0] + d[0], bpdi = b[1] + d[1];
                /* jbmd = (sign*i)*(b-d): forward uses -i in the odd outputs */
                double bmdr = b[0] - d[0], bmdi = b[1] - d[1];
                double jr = -sg * bmdi, ji = sg * bmdr;     /* (sg*i)*(b-d) */
                double *y0 = y + 2 * (size_t)(q + s * (4 * pp + 0));
                double *y1 = y + 2 * (size_t)(q + s * (4 * pp + 1));
                double *y2 = y + 2 * (size_t)(q + s * (4 * pp + 2));
                double *y3 = y + 2 * (size_t)(q + s * (4 * pp + 3));
                double t1r = amcr + jr, t1i = amci + ji;
                double t2r = apcr - bpdr, t2i = apci - bpdi;
                double t3r = amcr - jr, t3i = amci - ji;
                y0[0] = apcr + bpdr;            y0[1] = apci + bpdi;
                y1[0] = t1r * w1r - t1i * w1i;  y1[1] = t1r * w1i + t1i * w1r;
                y2[0] = t2r * w2r - t2i * w2i;  y2[1] = t2r * w2i + t2i * w2r;
                y3[0] = t3r * w3r - t3i * w3i;  y3[1] = t3r * w3i + t3i * w3r;
            }
        }
        { double *t = x; x = y; y = t; }
        len /= 4;
        s *= 4;
    }
    if (len == 2) {
        for (int q = 0; q < s; q++) {
            const double *a = x + 2 * (size_t)q;
            const double *b = x + 2 * (size_t)(q + s);
            double *y0 = y + 2 * (size_t)q;
            double *y1 = y + 2 * (size_t)(q + s);
            y0[0] = a[0] + b[0]; y0[1] = a[1] + b[1];
            y1[0] = a[0] - b[0]; y1[1] = a[1] - b[1];
        }
        { double *t = x; x = y; y = t; }
    }
    memcpy(out, x, sizeof(double) * 2 * (size_t)n);
    free(buf);
}
