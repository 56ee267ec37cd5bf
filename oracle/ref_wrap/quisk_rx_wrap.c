/* oracle/ref_wrap/quisk_rx_wrap.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Translation unit that turns the `static` receive-chain functions of the
 * reference's quisk.c into a loadable library WITHOUT copying them into this
 * repository: oracle/build_ref.sh extracts the line ranges listed below from
 * /root/reference/quisk.c into a scratch file under /tmp at build time, and
 * this wrapper `#include`s that scratch file.  Everything in THIS file is our
 * own glue: the handful of file-scope variables those functions expect
 * (quisk.c:127-134,191-195,202,254-269), no-op stand-ins for the optional
 * stages that are switched off by default (auto-notch, SSB squelch), and
 * `ref_*` accessors for ctypes.
 *
 * Extracted ranges (quisk.c):  46-53 constants, 68-81 struct AgcState,
 * 622-665 cFracDecim, 1182-1256 dRxFilterOut/cRxFilterOut,
 * 1633-1671 PlanDecimation, 1673-1846 quisk_process_decimate,
 * 1848-2160 quisk_process_demodulate, 2162-2287 process_agc,
 * 679-784 NoiseBlanker, 786-963 dAutoNotch, 1056-1084 d_delay, 1086-1180 ssb_squelch,
 * 2922-2953 the two sample-unpack branches of add_rx_samples,
 * 3746-3763 the 24-bit record loop of read_rx_udp10 (Hermes protocol 1).
 */
#include <Python.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <complex.h>
#include "quisk.h"
#include "filter.h"
#include <fftw3.h>                  /* oracle/fftw_shim */

#define DEBUG 0
#include "quisk_rx_consts.inc"      /* quisk.c:46-53, 68-81 */

struct sound_conf quisk_sound_state;

static double cFilterI[MAX_RX_FILTERS][MAX_FILTER_SIZE];
static double cFilterQ[MAX_RX_FILTERS][MAX_FILTER_SIZE];
static int sizeFilter;
static int filter_bandwidth[MAX_RX_FILTERS];
static int quisk_decim_srate;
static int quisk_filter_srate = 48000;
static double agcReleaseGain = 80;
static double agc_release_time = 1.0;
static double squelch_level = -999.0;
static int ssb_squelch_enabled;
static int rit_freq;
static double measured_audio;
static double measure_audio_sum;
static int measure_audio_count;
static int measure_audio_time = 1;
static struct _MeasureSquelch {
    int squelch_active;
    double rf_sum;
    double squelch;
    int rf_count;
    double *in_fft;
    int index;
    int sq_open;
} MeasureSquelch[MAX_RX_CHANNELS];

/* Optional stages, off by default in the reference (quisk_auto_notch == 0, ssb_squelch_enabled == 0): dAutoNotch,
 * ssb_squelch and d_delay are the reference's own (quisk.c:786-963, 1056-1084, 1086-1180; FFTW through the shim). */
int quisk_auto_notch;
#include "quisk_notch.inc"          /* quisk.c:786-963 dAutoNotch */
static int ssb_squelch_level;
#include "quisk_squelch.inc"        /* quisk.c:1056-1084 d_delay, 1086-1180 ssb_squelch */

#include "quisk_rx_funcs.inc"       /* the function ranges listed above */

/* ---- accessors ---- */
void ref_set_sample_rate(int rate) { quisk_sound_state.sample_rate = rate; }
void ref_set_playback_rate(int rate) { quisk_sound_state.playback_rate = rate; }
int ref_decim_srate(void) { return quisk_decim_srate; }
int ref_filter_srate(void) { return quisk_filter_srate; }
void ref_init_chain(void)
{
    quisk_process_decimate(NULL, 0, 0, 0);
    quisk_process_demodulate(NULL, NULL, 0, 0, 0, 0);
}
void ref_set_filters(const double *fi, const double *fq, int size, int bw, int nFilter)
{   /* same effect as set_filters(), quisk.c:4551-4594 */
    int i;
    filter_bandwidth[nFilter] = bw;
    for (i = 0; i < size; i++) { cFilterI[nFilter][i] = fi[i]; cFilterQ[nFilter][i] = fq[i]; }
    sizeFilter = size;
}
int ref_plan_decimation(int *p2, int *p3, int *p5) { return PlanDecimation(p2, p3, p5); }
int ref_process_decimate(complex double *cs, int n, int bank, int mode)
{ return quisk_process_decimate(cs, n, bank, (rx_mode_type)mode); }
int ref_process_demodulate(complex double *cs, double *ds, int n, int bank, int nFilter, int mode)
{ return quisk_process_demodulate(cs, ds, n, bank, nFilter, (rx_mode_type)mode); }
void ref_cRxFilterOut(complex double *cs, int n, int bank, int nFilter)
{ int i; for (i = 0; i < n; i++) cs[i] = cRxFilterOut(cs[i], bank, nFilter); }
void ref_dRxFilterOut(complex double *cs, int n, int bank, int nFilter)
{ int i; for (i = 0; i < n; i++) cs[i] = dRxFilterOut(cs[i], bank, nFilter); }
int ref_cFracDecim(complex double *cs, int n, double fdecim) { return cFracDecim(cs, n, fdecim); }

/* The tune loop is four lines inside quisk_process_samples (quisk.c:2477-2488);
 * `vec` plays the role of the static rxTuneVector and is carried by the caller. */
void ref_tune(complex double *cs, int n, double tune_hz, int sample_rate, complex double *vec)
{
    complex double phase = cexp((I * -2.0 * M_PI * tune_hz) / sample_rate);
    complex double v = *vec;
    int i;
    for (i = 0; i < n; i++) { cs[i] *= v; v *= phase; }
    *vec = v;
}

/* process_agc: opaque state handle for ctypes */
void *ref_agc_new(double max_out, int sample_rate)
{
    struct AgcState *s = (struct AgcState *)calloc(1, sizeof(*s));
    s->max_out = max_out;
    s->sample_rate = sample_rate;
    s->buf_size = 0;
    process_agc(s, NULL, 0, 0);     /* first call initialises (quisk.c:2174-2190) */
    return s;
}
void ref_agc_set(double release_gain, double release_time) { agcReleaseGain = release_gain; agc_release_time = release_time; }
void ref_agc_run(void *s, complex double *cs, int n, int is_cpx) { process_agc((struct AgcState *)s, cs, n, is_cpx); }
double ref_agc_gain(void *s) { return ((struct AgcState *)s)->gain; }


/* ssb_squelch + d_delay as the SSB branch of quisk_process_demodulate calls them (quisk.c:1925-1928): the very first
 * call only creates the FFT plan and returns (quisk.c:1104-1112).  Returns MS->squelch_active; *sq_open = the timer. */
int ref_ssb_squelch(double *ds, int n, int samp_rate, int bw, int level, int bank, int *sq_open)
{
    filter_bandwidth[0] = bw; ssb_squelch_level = level;
    ssb_squelch(ds, n, samp_rate, MeasureSquelch + bank);
    d_delay(ds, n, bank, SQUELCH_FFT_SIZE);
    if (sq_open) *sq_open = MeasureSquelch[bank].sq_open;
    return MeasureSquelch[bank].squelch_active;
}

/* dAutoNotch (quisk.c:786-963) as quisk_process_demodulate calls it on bank 0 (quisk.c:1923-1924): `reset` = the
 * initialising call with a NULL buffer (quisk.c:826-835). */
void ref_auto_notch(double *ds, int n, int sidetone, int rate, int reset)
{
    if (reset) dAutoNotch(NULL, 0, 0, 0);
    quisk_auto_notch = 1;
    dAutoNotch(ds, n, sidetone, rate);
    quisk_auto_notch = 0;
}

/* The optional stages inside quisk_process_demodulate itself: the reference's own switches (quisk_auto_notch,
 * ssb_squelch_enabled / ssb_squelch_level, rit_freq for the CW side tone), and the flag quisk_process_samples mutes on. */
void ref_set_chain_options(int auto_notch, int squelch_enabled, int squelch_level, int rit)
{   /* set_auto_notch (quisk.c:6011-6019) re-initialises the notch when it is switched, set_ssb_squelch (quisk.c:6021-6029) only stores */
    quisk_auto_notch = auto_notch; dAutoNotch(NULL, 0, 0, 0);
    ssb_squelch_enabled = squelch_enabled; ssb_squelch_level = squelch_level; rit_freq = rit;
}
int ref_squelch_active(int bank) { return MeasureSquelch[bank].squelch_active; }

/* NoiseBlanker (quisk.c:679-784): called on the raw samples in front of the tuning stage (quisk.c:2448-2449) when
 * quisk_noise_blanker > 0.  Its state is function-static: one private copy of this library per stream. */
int quisk_noise_blanker;
#include "quisk_nb.inc"
void ref_noise_blanker(complex double *cs, int n, int level) { quisk_noise_blanker = level; NoiseBlanker(cs, n); }


/* ---- wire-format ingest: the loops that turn received bytes into complex double ----
 * add_rx_samples (quisk.c:2894-2956) is a Python method; its two unpack branches (2922-2953) only touch the
 * variables declared here.  The `if (0) {}` supplies the head of the reference's if / else-if chain. */
static int py_sample_rx_bytes = 2;
static int py_sample_rx_endian;
static complex double PySampleBuf[SAMP_BUFFER_SIZE];
static int PySampleCount;
int ref_add_rx_samples(void *data, long len, int bytes, int big_endian, complex double *out)
{
    struct { void *buf; long len; } view = { data, len };
    int ii, qq, i;
    unsigned char *pt_ii, *pt_qq;
    py_sample_rx_bytes = bytes; py_sample_rx_endian = big_endian; PySampleCount = 0;
    if (0) {}
#include "quisk_unpack_py.inc"
    memcpy(out, PySampleBuf, PySampleCount * sizeof(complex double));
    return PySampleCount;
}

/* read_rx_udp10 (quisk.c:3526-3790): per 1032-byte packet two 512-byte frames starting at byte 11 (after the
 * 8-byte header and 3 sync bytes); the extracted loop reads num_records records of (1 + multirx) receivers. */
int quisk_multirx_count;
static complex double *multirx_cSamples[16];
static struct { int index; complex double *samples; } multirx_fft_data[16];
static int multirx_fft_width;       /* 0: the panadapter side copy stays off */
int ref_hermes_unpack(unsigned char *buf, int multirx, complex double *samp, complex double *sub /* [multirx][126] */)
{
    int i, j, xr, xi, index, start, nSamples = 0, num_records;
    complex double c;
    quisk_multirx_count = multirx;
    num_records = 504 / ((quisk_multirx_count + 1) * 6 + 2);       /* quisk.c:3545 */
    for (j = 0; j < multirx; j++) multirx_cSamples[j] = sub + j * 126;
    for (start = 11; start < 1000; start += 512) {                  /* quisk.c:3631 */
#include "quisk_unpack_hermes.inc"
    }
    return nSamples;
}
