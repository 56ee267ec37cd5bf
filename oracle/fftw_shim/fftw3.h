/* oracle/fftw_shim/fftw3.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Minimal stand-in for the FFTW3 API so that the reference's WDSP sources
 * (which `#include "fftw3.h"`, wdsp/comm.h:55) can be compiled in an image
 * that has no FFTW3.  Only the entry points the reference references are
 * provided (fftw_plan_dft_1d / _r2c_1d / _c2r_1d / execute / destroy_plan /
 * malloc / free / wisdom import+export).  Backed by oracle/fft64.c.
 * Unlike real FFTW, planning never touches the in/out arrays.
 */
#ifndef ORACLE_FFTW3_SHIM_H
#define ORACLE_FFTW3_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* as FFTW itself does: the C99 type when <complex.h> came first (quisk.c:5-6), else double[2] (wdsp) -- same bytes */
#if defined(_Complex_I) && defined(complex) && defined(I)
typedef double _Complex fftw_complex;
#else
typedef double fftw_complex[2];
#endif
typedef struct fftw_shim_plan_s *fftw_plan;

#define FFTW_FORWARD  (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE        (0U)
#define FFTW_DESTROY_INPUT  (1U << 0)
#define FFTW_EXHAUSTIVE     (1U << 3)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_PATIENT        (1U << 5)
#define FFTW_ESTIMATE       (1U << 6)
#define FFTW_WISDOM_ONLY    (1U << 21)

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_1d(int n, fftw_complex *in, double *out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out);
void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out);
void fftw_destroy_plan(fftw_plan p);
void *fftw_malloc(size_t n);
void fftw_free(void *p);
int fftw_import_wisdom_from_filename(const char *filename);
int fftw_export_wisdom_to_filename(const char *filename);
char *fftw_export_wisdom_to_string(void);

#ifdef __cplusplus
}
#endif
#endif
