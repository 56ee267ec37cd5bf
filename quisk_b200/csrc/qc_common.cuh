// quisk_b200/csrc/qc_common.cuh -- shared declarations for libquisk_cuda.so
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <mutex>
#include <atomic>

#include "../../include/quisk_cuda.h"

typedef double2 cd;   // device-side complex double, same bytes as quisk_cd

namespace qc {

void set_error(const char *fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// Returns QC_OK, or records the failure and returns QC_ECUDA.
int check(cudaError_t e, const char *what, const char *file, int line);
#define QC_CUDA(call) do { int _qc = qc::check((call), #call, __FILE__, __LINE__); if (_qc != QC_OK) return _qc; } while (0)
#define QC_CUDA_LAUNCH() do { int _qc = qc::check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__); if (_qc != QC_OK) return _qc; } while (0)

// Makes sure a device is usable; QC_ENODEV otherwise.  Cheap after the first call.
int ensure_device();
// For the legacy filter.h entry points, which have no error channel: print and abort.
[[noreturn]] void die_no_device(const char *fn);

// ---------------------------------------------------------------------------
// Generic streaming polyphase FIR (polyfir.cu)
//
//   y[m] = gain * sum_{t<K} X[src(m) - k(t)] (*) coef[ph(m) + k(t)*L],
//   u = u0 + m*M,  src = u / L,  ph = u % L
//
// X is the logical stream [history | block]: X[j] = in[j] for j >= 0 and
// hist[H + j] for -H <= j < 0 (zero below that).  Tap visiting order:
// order 0: k = 0,1,...,K-1 (filter.c loops); order 1: k = 0, K-1, K-2, ..., 1
// (the ring walk of cRxFilterOut, quisk.c:1240-1255, after the tap permutation
// h[0]=filt[0], h[m]=filt[N-m]).  Arithmetic is separately rounded multiply and
// add in exactly that order, i.e. what gcc emits for the reference without
// -ffast-math on baseline x86-64, so results are bit-identical to it.
// ---------------------------------------------------------------------------
enum TapMode { TAP_REAL = 0, TAP_COMPLEX = 1, TAP_SPLIT_IQ = 2 };
enum HbMode { HB_NONE = 0, HB_DECIM = 1, HB_INTERP = 2 };

struct PolyFirParams {
    const void *hist;       // [C][H] or nullptr (zeros)
    int H;
    const void *in;         // [C][in_stride]
    long in_stride;
    int n_in;
    void *out;              // [C][out_stride]
    long out_stride;
    int n_out;
    const double *coef;     // device taps (double, or double2 for TAP_COMPLEX / TAP_SPLIT_IQ)
    int K, L, M;
    long u0;
    double gain;
    int order;
    void *hist_out;         // [C][H] or nullptr: receives the last H samples of X
    int C;
    int is_complex;         // sample type: 1 = cd, 0 = double
    int tap_mode;
    int hb_mode;            // HB_*: the two half-band forms keep the reference's paired summation
};

int launch_polyfir(const PolyFirParams &p, cudaStream_t stream);

// Pointwise helpers (pointwise.cu)
int launch_demod_ssb(const cd *in, long in_stride, double *out, long out_stride, int n, int C, int lower, cudaStream_t s);
int launch_mute_rows(double *audio, long stride, int n, int C, const int *d_state, cudaStream_t s);
int launch_tune(const cd *in, long in_stride, cd *out, long out_stride, int n, int C,
                const double *d_nco /* [C][8] */, const cd *d_vstart /* [C] v at sample 0 of the block */, unsigned long long n0, cudaStream_t s);
int launch_nco_advance(const cd *v_in, cd *v_out, const double *d_nco, int count, int C, unsigned *d_sched, unsigned epoch, cudaStream_t s);
int launch_nco_jump(const cd *v_in, cd *v_out, const double *d_nco, int count, int C, cudaStream_t s);
int launch_am_detect(const cd *in, long in_stride, double *out, long out_stride, int n, int C, double *d_dc /*[C]*/, cudaStream_t s);
int launch_fm_detect(const cd *in, long in_stride, double *out, long out_stride, int n, int C,
                     double *d_state /*[C][4]: fm_1.re, fm_1.im, x_1, y_1*/, double a0, double a1, double b1, cudaStream_t s);

// Host-side extended precision helper for the tuning NCO (nco_host.cpp)
// Fills out[8] = { frac_hi, frac_lo, loggrowth, v0.re, v0.im, phase.re, phase.im, 0 } for
// phase = cexp(-j 2 pi tune / rate) rounded to double as the reference computes it.
void nco_make(double tune_hz, int sample_rate, double v0_re, double v0_im, double out[8]);
// v after n more steps of the reference recurrence in closed form (for retune / reporting)
void nco_advance(const double nco[8], unsigned long long n, double *v_re, double *v_im);

}  // namespace qc
