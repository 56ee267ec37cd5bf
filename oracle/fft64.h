/* oracle/fft64.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * A small double-precision complex FFT used (a) by the FFTW-API shim that lets
 * the reference WDSP sources link in this image (FFTW3 is an un-vendored,
 * unpinned third-party dependency of the reference: wdsp/Makefile:6,11,
 * setup.py:62) and (b) by the CPU restatement of the panadapter / fircore math.
 * Unnormalised in both directions, like FFTW (sign -1 forward, +1 backward).
 */
#ifndef ORACLE_FFT64_H
#define ORACLE_FFT64_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fft64_plan_s fft64_plan;

/* n may be any positive integer: powers of two use a table-driven Stockham
 * radix-4/2 autosort FFT, everything else an exact-index O(n^2) DFT. */
fft64_plan *fft64_create(int n, int sign);
void fft64_destroy(fft64_plan *p);
/* in/out are interleaved (re,im); in == out is allowed. */
void fft64_execute(const fft64_plan *p, const double *in, double *out);

#ifdef __cplusplus
}
#endif
#endif
