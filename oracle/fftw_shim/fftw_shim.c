/* oracle/fftw_shim/fftw_shim.c -- TEST INFRASTRUCTURE, not product code.
 * See fftw3.h in this directory. */
#include <stdlib.h>
#include <string.h>
#include "fftw3.h"
#include "../fft64.h"

enum { KIND_C2C = 0, KIND_R2C = 1, KIND_C2R = 2 };

struct fftw_shim_plan_s {
    int kind;
    int n;
    fft64_plan *fft;
    void *in;
    void *out;
};

static fftw_plan mk(int kind, int n, int sign, void *in, void *out)
{
    fftw_plan p = (fftw_plan)calloc(1, sizeof(*p));
    p->kind = kind;
    p->n = n;
    p->fft = fft64_create(n, sign);
    p->in = in;
    p->out = out;
    return p;
}

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags)
{
    (void)flags;
    return mk(KIND_C2C, n, sign, in, out);
}

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags)
{
    (void)flags;
    return mk(KIND_R2C, n, FFTW_FORWARD, in, out);
}

fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out, unsigned flags)
{
    (void)flags;
    return mk(KIND_C2R, n, FFTW_BACKWARD, in, out);
}

void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out)
{
    fft64_execute(p->fft, (const double *)in, (double *)out);
}

void fftw_execute(const fftw_plan p)
{
    int n = p->n, k;
    if (p->kind == KIND_C2C) {
        fft64_execute(p->fft, (const double *)p->in, (double *)p->out);
    } else if (p->kind == KIND_R2C) {
        const double *x = (const double *)p->in;
        double *o = (double *)p->out;
        double *t = (double *)malloc(sizeof(double) * 2 * (size_t)n);
        for (k = 0; k < n; k++) { t[2 * k] = x[k]; t[2 * k + 1] = 0.0; }
        fft64_execute(p->fft, t, t);
        memcpy(o, t, sizeof(double) * 2 * (size_t)(n / 2 + 1));
        free(t);
    } else {
        /* c2r: input holds n/2+1 bins of a Hermitian spectrum. */
        const double *x = (const double *)p->in;
        double *o = (double *)p->out;
        double *t = (double *)malloc(sizeof(double) * 2 * (size_t)n);
        for (k = 0; k <= n / 2; k++) { t[2 * k] = x[2 * k]; t[2 * k + 1] = x[2 * k + 1]; }
        for (k = n / 2 + 1; k < n; k++) { t[2 * k] = x[2 * (n - k)]; t[2 * k + 1] = -x[2 * (n - k) + 1]; }
        t[1] = 0.0;
        if ((n & 1) == 0) t[2 * (n / 2) + 1] = 0.0;
        fft64_execute(p->fft, t, t);
        for (k = 0; k < n; k++) o[k] = t[2 * k];
        free(t);
    }
}

void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out)
{   /* new-array execute of an r2c plan (quisk.c:1119) */
    struct fftw_shim_plan_s q = *p;
    q.in = in; q.out = out;
    fftw_execute(&q);
}

void fftw_destroy_plan(fftw_plan p)
{
    if (!p) return;
    fft64_destroy(p->fft);
    free(p);
}

void *fftw_malloc(size_t n) { void *p = NULL; if (posix_memalign(&p, 64, n ? n : 64)) return NULL; return p; }
void fftw_free(void *p) { free(p); }
int fftw_import_wisdom_from_filename(const char *filename) { (void)filename; return 0; }
int fftw_export_wisdom_to_filename(const char *filename) { (void)filename; return 0; }
char *fftw_export_wisdom_to_string(void) { char *s = (char *)malloc(1); if (s) s[0] = 0; return s; }
