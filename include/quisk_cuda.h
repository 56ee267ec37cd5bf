/* include/quisk_cuda.h -- C ABI of libquisk_cuda.so
 *
 * A B200 (sm_100a) implementation of Quisk's receive-DSP hot path behind the
 * reference's own C entry points.  Three groups of exports:
 *
 *   1. The seventeen functions of the reference's filter.h (filter.h:39-55),
 *      same names, same signatures, same caller-owned state structs (layouts
 *      below are ABI-identical to filter.h:1-37: 56 / 56 / 544 / 280 bytes on
 *      LP64).  Sample buffers are HOST pointers, processed in place, return
 *      value = new sample count -- exactly what `_quisk` (quisk.c, sound.c,
 *      microphone.c) links against today.  All arithmetic runs on the GPU.
 *
 *   2. `quisk_cuda_*` batched variants: many independent receiver channels
 *      laid out [channel][sample] in DEVICE memory with per-channel filter
 *      state kept resident in HBM.  These are new surface (the reference has
 *      no batched API); they are where throughput is measured.
 *
 *   3. `quisk_cuda_rx_*` / `quisk_cuda_pan_*`: the fused receive chain
 *      (quisk_process_samples' tune -> quisk_process_decimate ->
 *      quisk_process_demodulate, quisk.c:2477-2530) and the panadapter
 *      (get_graph, quisk.c:5142-5331) for a batch of channels; the stages
 *      around them (process_agc, cFracDecim, get_bandscope), the wire-format
 *      unpack loops in front of the chain, and the wideband polyphase
 *      channelizer (one stream -> many receivers).
 *
 *   The WDSP side (fircore, resampler, RXA channel) is in quisk_cuda_wdsp.h.
 *
 * No torch types, no C++ types: plain pointers and sizes.  `stream` arguments
 * are a `cudaStream_t` passed as `void *` (NULL = the legacy default stream).
 * Every function that can fail returns 0 on success and a negative QC_E* code
 * on failure, with a message retrievable through quisk_cuda_last_error().
 * There is no CPU fallback anywhere: without a usable CUDA device the legacy
 * entry points abort with a diagnostic (they have no error channel in the
 * reference either) and the quisk_cuda_* entry points return QC_ENODEV.
 */
#ifndef QUISK_CUDA_H
#define QUISK_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
typedef struct { double re, im; } quisk_cd;       /* layout of C99 `complex double` */
#else
#include <complex.h>
typedef complex double quisk_cd;
#endif

/* ------------------------------------------------------------------------
 * 1. filter.h drop-in
 * --------------------------------------------------------------------- */

struct quisk_cFilter {              /* filter.h:1-10 */
    double *dCoefs;
    quisk_cd *cpxCoefs;
    int nBuf;
    int nTaps;
    int decim_index;
    quisk_cd *cSamples;
    quisk_cd *ptcSamp;
    quisk_cd *cBuf;
};

struct quisk_dFilter {              /* filter.h:12-21 */
    double *dCoefs;
    quisk_cd *cpxCoefs;
    int nBuf;
    int nTaps;
    int decim_index;
    double *dSamples;
    double *ptdSamp;
    double *dBuf;
};

struct quisk_cHB45Filter {          /* filter.h:23-29 */
    quisk_cd *cBuf;
    int nBuf;
    int toggle;
    quisk_cd samples[22];
    quisk_cd center[11];
};

struct quisk_dHB45Filter {          /* filter.h:31-37 */
    double *dBuf;
    int nBuf;
    int toggle;
    double samples[22];
    double center[11];
};

void quisk_filt_cInit(struct quisk_cFilter *, double *, int);                   /* filter.h:39, filter.c:9   */
void quisk_filt_dInit(struct quisk_dFilter *, double *, int);                   /* filter.h:40, filter.c:22  */
void quisk_filt_differInit(struct quisk_dFilter *, int);                        /* filter.h:41, filter.c:35  */
void quisk_filt_tune(struct quisk_dFilter *, double, int);                      /* filter.h:42, filter.c:58  */
quisk_cd quisk_dC_out(double, struct quisk_dFilter *);                          /* filter.h:43, filter.c:83  */
double quisk_dD_out(double, struct quisk_dFilter *);                            /* filter.h:44, filter.c:326 */
int quisk_cInterpolate(quisk_cd *, int, struct quisk_cFilter *, int);           /* filter.h:45, filter.c:131 */
int quisk_dInterpolate(double *, int, struct quisk_dFilter *, int);             /* filter.h:46, filter.c:167 */
int quisk_cDecimate(quisk_cd *, int, struct quisk_cFilter *, int);              /* filter.h:47, filter.c:203 */
int quisk_cCDecimate(quisk_cd *, int, struct quisk_cFilter *, int);             /* filter.h:48, filter.c:231 */
int quisk_dDecimate(double *, int, struct quisk_dFilter *, int);                /* filter.h:49, filter.c:259 */
int quisk_cInterpDecim(quisk_cd *, int, struct quisk_cFilter *, int, int);      /* filter.h:50, filter.c:287 */
int quisk_cDecim2HB45(quisk_cd *, int, struct quisk_cHB45Filter *);             /* filter.h:51, filter.c:377 */
int quisk_dInterp2HB45(double *, int, struct quisk_dHB45Filter *);              /* filter.h:52, filter.c:420 */
int quisk_cInterp2HB45(quisk_cd *, int, struct quisk_cHB45Filter *);            /* filter.h:53, filter.c:455 */
int quisk_dFilter(double *, int, struct quisk_dFilter *);                       /* filter.h:54, filter.c:347 */
int quisk_cFilter(quisk_cd *, int, struct quisk_cFilter *);                     /* filter.h:55, filter.c:372 */

/* ------------------------------------------------------------------------
 * Library status
 * --------------------------------------------------------------------- */

#define QC_OK        0
#define QC_ENODEV   (-1)    /* no usable CUDA device / driver */
#define QC_ECUDA    (-2)    /* a CUDA runtime call failed     */
#define QC_EINVAL   (-3)    /* bad argument                   */
#define QC_ENOMEM   (-4)

const char *quisk_cuda_last_error(void);
int quisk_cuda_device_count(void);
int quisk_cuda_set_device(int device);
/* Kernels launched by this library since load (all threads); the bench reports the delta. */
unsigned long long quisk_cuda_launch_count(void);
const char *quisk_cuda_version(void);
/* Measured FP64 pipe peak of the current device in DFMA per second (a saturating micro-benchmark, ~10 ms):
 * the second roofline bench.py quotes for the kernels that sit near the FP64 ridge. */
int quisk_cuda_fp64_peak(double *dfma_per_second);

/* ------------------------------------------------------------------------
 * 2. Batched single-stage filters, device resident
 *
 * One object = one filter.h filter instantiated for `n_channels` independent
 * streams that are always fed the same number of samples per call (so the
 * decimation phase -- toggle / decim_index -- is one host-side integer and
 * output counts need no device round trip).  Input  [n_channels][in_stride],
 * output [n_channels][out_stride], element type quisk_cd or double as the
 * kind says.  Out-of-place (in != out).
 * --------------------------------------------------------------------- */

typedef struct qcBatchFilter qcBatchFilter;

enum qcFilterKind {
    QC_C_DECIM2_HB45 = 1,   /* quisk_cDecim2HB45                    complex -> complex */
    QC_C_DECIMATE    = 2,   /* quisk_cDecimate / quisk_cFilter      complex -> complex, real taps */
    QC_C_CDECIMATE   = 3,   /* quisk_cCDecimate                     complex -> complex, complex taps */
    QC_D_DECIMATE    = 4,   /* quisk_dDecimate / quisk_dFilter      real -> real */
    QC_C_INTERPOLATE = 5,   /* quisk_cInterpolate                   complex -> complex */
    QC_D_INTERPOLATE = 6,   /* quisk_dInterpolate                   real -> real */
    QC_C_INTERPDECIM = 7,   /* quisk_cInterpDecim                   complex -> complex */
    QC_C_INTERP2_HB45 = 8,  /* quisk_cInterp2HB45                   complex -> complex */
    QC_D_INTERP2_HB45 = 9,  /* quisk_dInterp2HB45                   real -> real */
    QC_C_RXFILTER    = 10,  /* cRxFilterOut (quisk.c:1218): I taps on I rail, Q taps on Q rail */
    QC_D_RXFILTER    = 11   /* dRxFilterOut (quisk.c:1182): one real tap set on complex samples */
};

/* coefs: HOST pointer, n_taps doubles (QC_C_CDECIMATE: n_taps complex = 2*n_taps
 * doubles; QC_C_RXFILTER: filtI then filtQ, 2*n_taps doubles).  Ignored for the
 * HB45 kinds.  interp/decim as in the matching filter.h call (unused ones = 1). */
qcBatchFilter *quisk_cuda_batch_create(int kind, int n_channels, const double *coefs, int n_taps,
                                       int interp, int decim);
void quisk_cuda_batch_destroy(qcBatchFilter *f);
/* Samples each channel will produce for `count` inputs given the current phase. */
int quisk_cuda_batch_count_out(const qcBatchFilter *f, int count);
/* The legacy API's silent output cut at 52 800 samples (filter.c:158) applies
 * only when `legacy_clip` is non-zero. */
int quisk_cuda_batch_run(qcBatchFilter *f, const void *d_in, long in_stride, int count,
                         void *d_out, long out_stride, int *n_out, int legacy_clip, void *stream);
/* Zero the history and phase (a freshly quisk_filt_cInit'ed filter). */
int quisk_cuda_batch_reset(qcBatchFilter *f, void *stream);

/* ------------------------------------------------------------------------
 * 3a. Batched receive chain
 * --------------------------------------------------------------------- */

typedef struct qcRxChain qcRxChain;

enum qcRxMode {     /* values of rx_mode_type, quisk.h:56-70 */
    QC_MODE_CWL = 0, QC_MODE_CWU = 1, QC_MODE_LSB = 2, QC_MODE_USB = 3, QC_MODE_AM = 4, QC_MODE_FM = 5,
    QC_MODE_DGT_U = 7, QC_MODE_DGT_L = 8, QC_MODE_DGT_IQ = 9, QC_MODE_FDV_U = 11, QC_MODE_FDV_L = 12,
    QC_MODE_DGT_FM = 13     /* the FM branch (quisk.c:2026-2027 is one case for FM and DGT_FM) */
};

/* The reference keeps its decimation / audio coefficient tables in filters.h;
 * a caller hands the ones a chain needs to the library by name (the drop-in
 * build links the reference's own filters.h, tests load them from a fixture). */
struct qcRxTables {
    const double *filt144D3;      int n_filt144D3;       /* quiskFilt144D3Coefs       (147) */
    const double *filt240D5Sharp; int n_filt240D5Sharp;  /* quiskFilt240D5CoefsSharp  (245) */
    const double *filt48dec24;    int n_filt48dec24;     /* quiskFilt48dec24Coefs     (98)  */
    const double *filt300D5;      int n_filt300D5;       /* quiskFilt300D5Coefs       (125) */
    const double *audio24p4;      int n_audio24p4;       /* quiskAudio24p4Coefs       (50)  */
    const double *audio24p6;      int n_audio24p6;       /* quiskAudio24p6Coefs       (36)  */
    const double *lpFilt48;       int n_lpFilt48;        /* quiskLpFilt48Coefs        (186) */
    const double *audioFmHp;      int n_audioFmHp;       /* quiskAudioFmHpCoefs       (309) */
    const double *filt53D1;       int n_filt53D1;        /* SDR-IQ special rates, quisk.c:1731-1767 */
    const double *filt111D2;      int n_filt111D2;
    const double *filt133D2;      int n_filt133D2;
    const double *filt167D3;      int n_filt167D3;
    const double *filt185D3;      int n_filt185D3;
};

struct qcRxConfig {
    int n_channels;
    int sample_rate;            /* quisk_sound_state.sample_rate */
    int mode;                   /* enum qcRxMode */
    const double *filt_i;       /* set_filters() tap tables (quisk.c:4551), HOST, n_filt doubles each */
    const double *filt_q;
    int n_filt;
    const double *tune_hz;      /* per-channel rx_tune_freq in Hz (HOST, n_channels) or NULL = no tuning */
    struct qcRxTables tables;
    int fused;                  /* 1 = fused shared-memory cascade kernels, 0 = one kernel per stage */
    int filter_bandwidth;       /* filter_bandwidth[nFilter] in Hz (set_filters, quisk.c:4551).  Only the digital modes look
                                 * at it: DGT-U/L and FDV-U/L filter at 6 kS/s below DGT_NARROW_FREQ = 3000 (quisk.c:2089) and at
                                 * 48 kS/s otherwise; DGT-IQ skips its filter at >= 19000 (quisk.c:2144).  DGT-IQ returns complex
                                 * samples: d_audio then holds (re, im) pairs, *n_audio counts pairs, audio_stride counts doubles. */
};

/* MakeFilterCoef (quisk.py:5405-5456), host side: the I/Q tap tables set_filters() hands to cRxFilterOut / dRxFilterOut.
 * proto / n_proto: the filters.py prototype for key quisk_cuda_filter_key(rate, bw) = bw * 24000 // rate // 2 when
 * that table exists, else NULL / 0 and a Blackman-windowed Dirichlet kernel of N taps is designed (N <= 0: sized from
 * the bandwidth like the reference's N = None).  center != 0: tuned by 2 exp(-j 2 pi center / rate (i - D)), filt_i =
 * real parts, filt_q = imaginary parts; center == 0: both are the low-pass.  *n_taps = taps written (QC_ENOMEM and
 * the needed count if cap is too small).  Bit-identical to the Python reference. */
int quisk_cuda_filter_key(int rate, int bw);
int quisk_cuda_make_filter_coef(int rate, int N, int bw, int center, const double *proto, int n_proto,
                                double *filt_i, double *filt_q, int cap, int *n_taps);

qcRxChain *quisk_cuda_rx_create(const struct qcRxConfig *cfg);
void quisk_cuda_rx_destroy(qcRxChain *rx);
/* PlanDecimation (quisk.c:1633): returns the planned rate, fills the three counts. */
int quisk_cuda_plan_decimation(int sample_rate, int *decim2, int *decim3, int *decim5);
int quisk_cuda_rx_decim_srate(const qcRxChain *rx);     /* quisk_decim_srate  */
int quisk_cuda_rx_filter_srate(const qcRxChain *rx);    /* quisk_filter_srate */
int quisk_cuda_rx_squelch_active(qcRxChain *rx, int *h_active);            /* [n_channels] squelch_active after the last block (QC_RX_OPT_SSB_SQUELCH; quisk.c:1179) */
/* Upper bound on audio samples per channel for `count` inputs (for sizing buffers). */
int quisk_cuda_rx_max_out(const qcRxChain *rx, int count);
/* d_iq: [n_channels][iq_stride] quisk_cd (device).  d_audio: [n_channels][audio_stride]
 * double (device), real audio at ~48 kS/s.  *n_audio = samples written per channel.
 * Optionally also returns the complex samples after quisk_process_decimate
 * (d_decim may be NULL): [n_channels][decim_stride] quisk_cd, *n_decim each. */
int quisk_cuda_rx_process(qcRxChain *rx, const void *d_iq, long iq_stride, int count,
                          double *d_audio, long audio_stride, int *n_audio,
                          void *d_decim, long decim_stride, int *n_decim, void *stream);
/* Same through HOST buffers: H2D copy of the block, the chain, D2H copy of the audio. */
int quisk_cuda_rx_process_host(qcRxChain *rx, const quisk_cd *h_iq, long iq_stride, int count,
                               double *h_audio, long audio_stride, int *n_audio);
/* Same with the host block still in its wire format (see section 3e): `bytes` per component, (I, Q) packed,
 * unpacked on the device exactly as add_rx_samples does on the host (quisk.c:2922-2953).  h_bytes:
 * [n_channels][byte_stride] bytes.  The H2D copy then carries 2*bytes per sample instead of 16. */
int quisk_cuda_rx_process_host_packed(qcRxChain *rx, const void *h_bytes, long byte_stride, int count, int bytes, int big_endian,
                                      double *h_audio, long audio_stride, int *n_audio);
int quisk_cuda_rx_reset(qcRxChain *rx);
/* Tuning knobs and instrumentation (not part of the reference's interface). */
#define QC_RX_OPT_TIMING       1   /* value != 0: record CUDA events around the dominant (fused) kernel */
#define QC_RX_OPT_FUSED_CHUNK  2   /* target input samples per shared-memory chunk of the fused decimator */
#define QC_RX_OPT_FUSED_THREADS 3  /* CTA width of the fused decimator: 128 or 256 */
#define QC_RX_OPT_EXACT_NCO   10   /* 1 (default): the tuning phasor at every block start is produced by the reference's own rounded
                                    * recurrence (quisk.c:2486) on a side stream, so tuned output follows the reference for any stream
                                    * length; 0: closed form only, ~10 % faster, drifts ~1e-19 per sample from the reference */
#define QC_RX_OPT_FUSED_TAIL   9   /* SSB / CW: one kernel for receive filter + demodulation + audio interpolators (default 1) */
#define QC_RX_OPT_FUSED_DEEPK  8   /* plan kernels: low-rate stages run once per this many chunks (1 or 4, default 1: measured slower at 4 with two CTAs per SM) */
#define QC_RX_OPT_TRACE        7   /* debug: record clock64() stamps per chunk phase in the fused kernel */
#define QC_RX_OPT_FUSED_PLANS  6   /* 1 (default): use plan-specialised kernels when the stage list matches one */
#define QC_RX_OPT_FUSED_DENSE  5   /* 1: 128-register cap (more resident CTAs), 0: up to 255 registers */
#define QC_RX_OPT_FUSED_TAILWARP 12 /* plan kernels run the low-rate stages on two tail warps, one chunk behind the four main warps: 1 (default: from stage 3), 2..4 = first tail stage, 0 = off */
#define QC_RX_OPT_FUSED_P3    20   /* 1: three-group pipeline kernel (commit + half band 0 | half bands 1-2 | the low-rate stages, each one chunk
                                      behind the one in front), 0: tail-warp kernel */
#define QC_RX_OPT_FUSED_ASYNC 19   /* tail-warp kernel: the next chunk travels by cp.async straight into stage 0's shared-memory buffer instead of
                                      waiting in registers: 1 = on, 2 = on under a 128-register cap, 0 = register prefetch */
#define QC_RX_OPT_FUSED_SPLIT 11   /* half-band stages of the plan kernels with one lane per component, twice the outputs per lane: 0 = off, 1 = on,
                                      2 = in the tail-warp kernel behind its first half band (measured +3 %) and in the four-stage 192 kS/s plan
                                      kernel (measured +4 %), 3 (default) = the first half band of the tail-warp kernel too (+1 %);
                                      bit-identical in every form */
#define QC_RX_OPT_NOISE_BLANKER 13  /* quisk_noise_blanker (0 = off, 1..3): quisk_cuda_rx_process_host / _host_packed run NoiseBlanker
                                      (quisk.c:679-784) on the staged block in front of the tuning stage, as quisk_process_samples
                                      does (quisk.c:2448-2449).  The device entry leaves the caller's buffer alone: run quisk_cuda_nb_run first */
#define QC_RX_OPT_AUTO_NOTCH   16   /* quisk_auto_notch: dAutoNotch on the audio at the filter rate (CW / SSB / AM), between the detector and the
                                      audio interpolators as in quisk_process_demodulate (quisk.c:1923-1924); the audio comes 1538 samples late */
#define QC_RX_OPT_NOTCH_SIDETONE 17 /* rit_freq handed to dAutoNotch in the CW modes (a CW side tone is never notched) */
#define QC_RX_OPT_SSB_SQUELCH  18   /* ssb_squelch_level (0 = off): ssb_squelch + d_delay behind the notch (quisk.c:1925-1928); the block of a
                                      receiver whose squelch is closed comes back as zeros, as quisk_process_samples mutes it (quisk.c:2716-2719).
                                      With either option the SSB / CW tail runs as per-stage kernels instead of the fused one. */
#define QC_RX_OPT_HOST_CHUNKS  15   /* quisk_cuda_rx_process_host / _host_packed: channel chunks whose H2D copy, kernels and D2H copy are pipelined
                                      over streams of their own; 0 (default) = 8 from 1024 channels up, else 1; 1 = one copy-compute-copy sequence.
                                      With more than one chunk the host entries keep stage state of their own, separate from quisk_cuda_rx_process */
#define QC_RX_OPT_FUSED_MIN_R  4   /* minimum outputs per thread in its half-band stages: 0 (auto), 2, 4, 8 */
int quisk_cuda_rx_set_option(qcRxChain *rx, int option, int value);
/* Sum of the event-timed durations (ms) of the dominant kernel since the last call, and how
 * many launches that covers.  Synchronises on the recorded events. */
int quisk_cuda_rx_kernel_time(qcRxChain *rx, double *ms_total, int *launches);
/* Name (template instantiation) of the fused decimator kernel the last quisk_cuda_rx_process launched, "" before the
 * first call or on the per-stage path.  bench.py quotes an ncu DRAM-traffic capture only when it names this kernel. */
const char *quisk_cuda_rx_fused_kernel_name(qcRxChain *rx);
/* Debug: copy the [n_channels][16 chunks][16 stamps] clock64() trace of the last fused launch to the host. */
int quisk_cuda_rx_read_trace(qcRxChain *rx, long long *host_out, int n_channels);

/* ------------------------------------------------------------------------
 * 3b. Batched panadapter (get_graph, quisk.c:5142-5331)
 * --------------------------------------------------------------------- */

typedef struct qcPanadapter qcPanadapter;

/* fft_size: a power of two in [8, 32768] (pan_multirx: <= 8192), or ANY size in [8, 16384] (Quisk's own fft_size = data_width * fft_mult has
 * factors 3 ... 15, quisk.py:187-194, 4179): those run as a Bluestein convolution on the power-of-two transform, two
 * transforms per frame.  quisk_cuda_pan_multirx needs a power of two. */
qcPanadapter *quisk_cuda_pan_create(int n_streams, int fft_size);
void quisk_cuda_pan_destroy(qcPanadapter *p);
/* Window + FFT + fftshift + |X| accumulate of `n_frames` consecutive frames per
 * stream (quisk.c:5212-5215, 5271-5276).  d_frames: [n_streams][frame_stride]
 * quisk_cd with the n_frames frames of a stream contiguous (frame f at
 * offset f*fft_size).  Adds into the per-stream average and bumps count_fft. */
int quisk_cuda_pan_accumulate(qcPanadapter *p, const void *d_frames, long stream_stride,
                              int n_frames, void *stream);
/* The graph-return half (quisk.c:5279-5321): pixel binning, dB scaling, clamp to
 * [-200, 0]; zeroes the averages and count_fft.  d_graph: [n_streams][data_width]
 * double (device). */
int quisk_cuda_pan_graph(qcPanadapter *p, int data_width, double zoom, double deltaf,
                         double fft_sample_rate, double *d_graph, void *stream);
/* get_multirx_graph (quisk.c:4868-4930): one frame per stream, |X| summed in
 * groups of 8 bins, no averaging state.  d_graph: [n_streams][fft_size/8]. */
int quisk_cuda_pan_multirx(qcPanadapter *p, const void *d_frames, long stream_stride,
                           double *d_graph, void *stream);
int quisk_cuda_pan_count(const qcPanadapter *p);
/* Raw device pointer to the [n_streams][fft_size] running |X| sums (for tests). */
const double *quisk_cuda_pan_average_ptr(const qcPanadapter *p);

/* get_bandscope (quisk.c:4957-5011): real blocks of `size` samples (power of two <= 8192), Hann window,
 * FFT, |X| average over L = size/2+1 bins; then copy2pixels (quisk.c:4932-4955), scale, 20 log10 (<= 1e-10 -> -200). */
typedef struct qcBandscope qcBandscope;
qcBandscope *quisk_cuda_bandscope_create(int n_streams, int size);
void quisk_cuda_bandscope_destroy(qcBandscope *b);
int quisk_cuda_bandscope_accumulate(qcBandscope *b, const double *d_blocks, long stream_stride, int n_blocks, void *stream);
int quisk_cuda_bandscope_graph(qcBandscope *b, int graph_width, int clock, double zoom, double deltaf, double *d_graph, void *stream);

/* ---- 3b'. Waterfall pixel mapper behind the panadapter (quisk.c:5334-5480), batched over n_streams ----
 * create          = watfall_RgbData (quisk.c:5334-5371): three 256-entry palettes, a ring of max_height zeroed rows of `width` pixels.
 * on_graph_data   = watfall_OnGraphData (quisk.c:5373-5420): the ring steps back one row; d_db [n_streams][db_stride] holds n_db dB
 *                   values per stream (quisk_cuda_pan_graph's output, on the device); colour index
 *                   (int)((dB - gain + 40.0 + y_zero * 0.69) * (y_scale + 10) * 0.10 + 128) clamped to 0..255; zero fill past n_db.
 * get_pixels      = watfall_GetPixels (quisk.c:5439-5480): `height` lines of width * 3 bytes per stream into
 *                   d_pixels + stream * stream_stride_bytes, newest first, every row shifted by its own x_origin against the one asked
 *                   for; scroll_mode (the reference's config value waterfall_scroll_mode, default 1) draws the newest seven rows
 *                   8, 7, ... 2 times first (35 lines: height must be >= 35 then).  Byte-exact against the reference's methods. */
typedef struct qcWaterfall qcWaterfall;
qcWaterfall *quisk_cuda_waterfall_create(int n_streams, int width, int max_height, const unsigned char *red, const unsigned char *green, const unsigned char *blue);
void quisk_cuda_waterfall_destroy(qcWaterfall *w);
int quisk_cuda_waterfall_on_graph_data(qcWaterfall *w, const double *d_db, long db_stride, int n_db, int y_zero, int y_scale, double gain, int x_origin, void *stream);
int quisk_cuda_waterfall_get_pixels(qcWaterfall *w, unsigned char *d_pixels, long stream_stride_bytes, int x_origin, int height, int scroll_mode, void *stream);

/* ------------------------------------------------------------------------
 * 3c. process_agc (quisk.c:2162-2287) and cFracDecim (quisk.c:622-665), batched
 * --------------------------------------------------------------------- */
typedef struct qcAgc qcAgc;
/* struct AgcState {max_out, sample_rate} + agcReleaseGain + agc_release_time (quisk.c:68-81,191-192) */
qcAgc *quisk_cuda_agc_create(int n_channels, int sample_rate, double max_out, double release_gain, double release_time);
void quisk_cuda_agc_destroy(qcAgc *a);
/* in place on d_samples [n_channels][stride] quisk_cd; is_cpx as in the reference call */
int quisk_cuda_agc_run(qcAgc *a, void *d_samples, long stride, int count, int is_cpx, void *stream);

typedef struct qcFracDecim qcFracDecim;
qcFracDecim *quisk_cuda_fracdecim_create(int n_channels);
void quisk_cuda_fracdecim_destroy(qcFracDecim *f);
int quisk_cuda_fracdecim_run(qcFracDecim *f, const void *d_in, long in_stride, int count, double fdecim,
                             void *d_out, long out_stride, int *n_out, void *stream);

/* ---- 3c'. The transmit-audio chain of microphone.c, batched over n_channels independent transmitters (SURVEY.md 8(f) row 4,
 * "the TX mirror"): tx_filter (microphone.c:372-604) with its peak rounder CcmPeak (:161-233).  Input: microphone audio
 * on the real rail of complex double samples, +-CLIP16, at 48000 or 8000 samples per second (`count` per transmitter, any
 * block size); output: the filtered, compressed, peak-limited modulation at 48 kS/s -- complex (I/Q) for LSB / USB, on the
 * real rail for AM / FM -- exactly what tx_filter leaves in its `filtered` buffer for transmit_mic_carrier / the SSB
 * up-converter to take.  mic_preemphasis = quisk_mic_preemphasis, mic_clip = quisk_mic_clip (microphone.c:35-37).  The
 * three coefficient tables are the reference's filters.h tables of the same names.  CcmPeak's first call only
 * initialises (microphone.c:174-191): the first block of a stream passes the peak rounder untouched, as there. */
typedef struct qcTxFilter qcTxFilter;
typedef struct {
    const double *mic_filt8;  int n_mic_filt8;      /* quiskMicFilt8Coefs  (93)  */
    const double *lp_filt48;  int n_lp_filt48;      /* quiskLpFilt48Coefs  (186) */
    const double *tx8k_audio; int n_tx8k_audio;     /* quiskFiltTx8kAudioB (168) */
    const double *dgt_filt48; int n_dgt_filt48;     /* quiskDgtFilt48Coefs (520): the digital modes only, may be NULL otherwise */
} qcTxTables;
qcTxFilter *quisk_cuda_tx_filter_create(int n_channels, int mode /* QC_MODE_LSB, _USB, _AM, _FM: tx_filter; _DGT_U/L, _FDV_U/L: tx_filter_digital
                                           (microphone.c:605-624: one tuned filter at 48 kS/s, no speech processing) */, int mic_sample_rate,
                                        double mic_preemphasis, double mic_clip, const qcTxTables *tables);
void quisk_cuda_tx_filter_destroy(qcTxFilter *t);
/* process_alc (microphone.c:270-370) behind the filter, as quisk_process_microphone chains them (:1232-1233).  enable = 1:
 * init_alc(&tx_alc, 960) the first time and init_alc(&tx_alc, 0) every time (key down, :1207: the 20 ms delay line and the
 * gain ramp cleared, the gain itself kept); enable = 0: off */
int quisk_cuda_tx_filter_set_alc(qcTxFilter *t, int enable);
int quisk_cuda_tx_filter_max_out(const qcTxFilter *t, int count);        /* upper bound of the samples one call returns */
int quisk_cuda_tx_filter_process(qcTxFilter *t, const void *d_in, long in_stride, int count,
                                 void *d_out, long out_stride, int *n_out, void *stream);   /* device pointers, strides in complex samples */

/* ---- 3d. NoiseBlanker (quisk.c:679-784), batched: the optional impulse blanker Quisk runs on the raw samples in front
 * of the tuning stage (quisk.c:2448-2449).  In place on d_samples [n_channels][stride] quisk_cd;
 * `level` = quisk_noise_blanker (1, 2, 3 -> threshold 6, 4, 2.5 times the mean magnitude of the last 1.5 ms; <= 0: off,
 * the call does nothing, as in the reference).  Output is the input delayed by 3 * (int)(sample_rate * 500e-6 + 0.5)
 * samples with the blanking applied; state carries over between calls of any length. ---- */
typedef struct qcNoiseBlanker qcNoiseBlanker;
qcNoiseBlanker *quisk_cuda_nb_create(int n_channels, int sample_rate);
void quisk_cuda_nb_destroy(qcNoiseBlanker *b);
int quisk_cuda_nb_run(qcNoiseBlanker *b, void *d_samples, long stride, int count, int level, void *stream);

/* ---- 3d'. ssb_squelch + d_delay (quisk.c:1056-1180), batched: the optional spectral-flatness squelch of the SSB branch of
 * quisk_process_demodulate (quisk.c:1925-1928, 1948-1951, 1970-1973).  samp_rate = quisk_filter_srate,
 * filter_bandwidth = filter_bandwidth[0], level = ssb_squelch_level.  In place on d_audio [n_channels][stride] double at
 * the filter rate: the audio comes back delayed by 512 samples; per channel sq_open (the one-second timer, in samples)
 * and squelch_active (what quisk_process_samples mutes on, quisk.c:2552-2623) are kept on the device --
 * quisk_cuda_ssb_squelch_state copies them out, quisk_cuda_ssb_squelch_state_ptr returns the device array
 * [n_channels][2] = {sq_open, squelch_active}.  As in the reference the first call only sets up (its samples do not
 * enter the analysis frame) and the timer is decremented once per call, so `count` is limited to 8192. ---- */
typedef struct qcSsbSquelch qcSsbSquelch;
qcSsbSquelch *quisk_cuda_ssb_squelch_create(int n_channels, int samp_rate, int filter_bandwidth);
void quisk_cuda_ssb_squelch_destroy(qcSsbSquelch *s);
int quisk_cuda_ssb_squelch_run(qcSsbSquelch *s, double *d_audio, long stride, int count, int level, void *stream);
int quisk_cuda_ssb_squelch_state(qcSsbSquelch *s, int *sq_open, int *squelch_active, void *stream);
const int *quisk_cuda_ssb_squelch_state_ptr(qcSsbSquelch *s);

/* ---- 3d''. dAutoNotch (quisk.c:786-963), batched: the optional automatic notch of the SSB / AM branch of
 * quisk_process_demodulate (quisk.c:1923-1924, on bank 0 when quisk_auto_notch is set).  rate = quisk_filter_srate,
 * sidetone = rit_freq (a CW side tone is never notched).  In place on d_audio [n_channels][stride] double; the audio
 * comes back 1538 samples late (overlap-save frames of 2048 with a 511-tap filter).  The object starts in the state
 * the reference's initialising call dAutoNotch(NULL, ...) leaves. ---- */
typedef struct qcAutoNotch qcAutoNotch;
qcAutoNotch *quisk_cuda_autonotch_create(int n_channels, int rate);
void quisk_cuda_autonotch_destroy(qcAutoNotch *a);
int quisk_cuda_autonotch_run(qcAutoNotch *a, double *d_audio, long stride, int count, int sidetone, void *stream);

/* ---- 3e. wire-format ingest: received bytes -> complex double on the device ----
 * quisk_cuda_unpack_iq: add_rx_samples (quisk.c:2922-2953).  d_bytes [n_channels][byte_stride]: packed (I, Q)
 * pairs, `bytes` = 1..4 per component, little endian (big_endian = 0) or big endian; each component is
 * left-justified into an int32 and converted int -> float -> double like the reference's `ii + qq * I`
 * (4-byte samples therefore keep 24 significant bits, as they do in Quisk).  d_out [n_channels][out_stride] quisk_cd.
 * quisk_cuda_unpack_hermes: the record loop of read_rx_udp10 (quisk.c:3631,3746-3763).  d_packets = n_packets
 * consecutive 1032-byte Metis payloads with n_rx = 1 + quisk_multirx_count receivers; receiver r goes to
 * d_out[r * out_stride + sample]; *n_samples = n_packets * quisk_cuda_hermes_samples_per_packet(n_rx).  Sequence
 * numbers, sync bytes and the C0..C4 control bytes stay the host's business (quisk.c:3620-3744). */
int quisk_cuda_unpack_iq(const void *d_bytes, long byte_stride, int n_channels, int count, int bytes, int big_endian,
                         void *d_out, long out_stride, void *stream);
int quisk_cuda_hermes_samples_per_packet(int n_rx);
int quisk_cuda_unpack_hermes(const void *d_packets, int n_packets, int n_rx, void *d_out, long out_stride, int *n_samples, void *stream);

/* ---- wideband polyphase channelizer (configuration C5) ----
 * One wideband stream -> n_channels receivers, receiver k centred on k*fs/n_channels, each decimated by `decim`
 * through the prototype low-pass `proto` (n_taps real taps, a multiple of n_channels).  Output k equals the
 * reference's per-receiver front end: mix by exp(-2 pi i k n / n_channels) with n the absolute sample index
 * (quisk.c:2477-2494), then quisk_cDecimate(proto, decim) (filter.c:203-229) -- computed as branch FIRs + one
 * FFT per output frame.  n_channels in {256, 512, 1024}; n_taps/n_channels in {4, 8, 16, 32}; decim <= n_channels.
 * layout 0: d_out[k * out_stride + frame] (feeds quisk_cuda_rx_process); layout 1: d_out[frame * out_stride + k].
 * State carried between calls: the last n_taps input samples and the absolute sample index, so any block
 * length is allowed.  Time-block sharding: quisk_cuda_pfb_seek(t - halo) zeroes the state at an absolute index,
 * quisk_cuda_pfb_prime() feeds the halo samples without producing output. */
typedef struct qcChannelizer qcChannelizer;
qcChannelizer *quisk_cuda_pfb_create(int n_channels, int decim, const double *proto, int n_taps);
void quisk_cuda_pfb_destroy(qcChannelizer *p);
int quisk_cuda_pfb_count_out(const qcChannelizer *p, int count);
int quisk_cuda_pfb_seek(qcChannelizer *p, long long n_abs);      /* synchronises the device first */
int quisk_cuda_pfb_seek_async(qcChannelizer *p, long long n_abs, void *stream);  /* ordered on `stream` like process / prime */
int quisk_cuda_pfb_prime(qcChannelizer *p, const void *d_in, int count, void *stream);
#define QC_PFB_OPT_SLICE_FRAMES 1   /* frames of the branch-FIR intermediate per kernel pair (default 65536 = 1 GiB at 1024 channels) */
#define QC_PFB_OPT_GENERIC      2   /* 1: force the single generic kernel (the only path when decim is not n_channels or n_channels/2) */
#define QC_PFB_OPT_PIPELINE     3   /* 1: overlap the branch FIRs of slice i+1 with the transforms of slice i (two internal streams) */
#define QC_PFB_OPT_FFT_PREFETCH 6   /* 1 (default): transform CTAs prefetch a later CTA's inputs into L2 */
#define QC_PFB_OPT_RING         5   /* cp.async ring depth of the branch-FIR kernel in steps: 16 (default) or 32 */
#define QC_PFB_OPT_FUSED        7   /* 1: 1024 channels, decim 512, 8 or 16 taps per branch run as ONE kernel (clusters of 8 CTAs, the branch-FIR
                                       intermediate stays in distributed shared memory: 48 B of HBM traffic per sample instead of 112);
                                       bit-identical, but measured slower (one CTA per SM, 15 clusters resident): default 0, the two-kernel path */
#define QC_PFB_OPT_FFT_FRAMES   4   /* frames interleaved per transform CTA: 2 (default) or 4 */
int quisk_cuda_pfb_set_option(qcChannelizer *p, int option, int value);
int quisk_cuda_pfb_process(qcChannelizer *p, const void *d_in, int count, void *d_out, long out_stride, int layout,
                           int *n_frames, void *stream);

/* Batched complex FFT (unnormalised, sign -1 forward / +1 backward), sizes 2^k,
 * 8 <= n <= 8192: the in-house Stockham kernel the panadapter and the WDSP
 * overlap-save stages are built on; exported so tests can pin it against
 * numpy / cuFFT.  d_in may equal d_out. */
int quisk_cuda_fft_batch(const void *d_in, void *d_out, int n, int batch, int sign, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QUISK_CUDA_H */
