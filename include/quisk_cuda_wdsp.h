/* include/quisk_cuda_wdsp.h -- C ABI of libquisk_cuda.so, WDSP RXA part (WDSP 1.25 as vendored by
 * Quisk 4.2.52 under wdsp/).
 *
 * Batched, device-resident versions of the RXA stages the receive hot path uses: the
 * partitioned overlap-save FIR `fircore` (wdsp/firmin.c:290-430) that nbp / bandpass / fmd run on,
 * the rational resampler (wdsp/resample.c), the frequency shifter (wdsp/shift.c), wcpAGC
 * (wdsp/wcpAGC.c), the FM and AM demodulators (wdsp/fmd.c, wdsp/amd.c), the patch panel and meters,
 * and `quisk_cuda_rxa_*`: the stage order of xrxa (wdsp/RXA.c:561-598) with create_rxa's defaults
 * (wdsp/RXA.c:31-490) for n_channels independent channels at once.
 *
 * All sample buffers are interleaved complex double (re, im) == the `double *` buffers of WDSP,
 * laid out [channel][sample] in DEVICE memory.  Host-side design helpers return plain arrays.
 * Same status codes / error string as quisk_cuda.h.
 */
#ifndef QUISK_CUDA_WDSP_H
#define QUISK_CUDA_WDSP_H

#include "quisk_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- host-side design (taps are computed on the CPU with libm, bit-identical to the reference) ---- */
/* fir_bandpass, wdsp/fir.c:187-254.  rtype 0: N doubles, rtype 1: N complex. wintype 0 = BH4, 1 = BH7. */
int quisk_cuda_fir_bandpass(int N, double f_low, double f_high, double samplerate, int wintype, int rtype,
                            double scale, double *out);
/* fc_impulse, wdsp/fcurve.c:29-143 (FM pre/de-emphasis), nc complex out. */
int quisk_cuda_fc_impulse(int nc, double f0, double f1, double g0, double g1, int curve, double samplerate,
                          double scale, int ctfmode, int wintype, double *out);
/* calc_resample, wdsp/resample.c:35-78: L, M, ncoef and (if h != NULL) the ncoef prototype taps. */
int quisk_cuda_resample_design(int in_rate, int out_rate, double fc, int ncoef_in, double gain,
                               int *L, int *M, int *ncoef, double *h, int h_cap);
/* the same with calc_resample's fc_low member: < 0 low pass, >= 0 the band pass of setFCLow_resample (resample.c:185-193) */
int quisk_cuda_resample_design_band(int in_rate, int out_rate, double fc_low, double fc, int ncoef_in, double gain,
                                    int *L, int *M, int *ncoef, double *h, int h_cap);

/* calc_nbp_impulse with the notches running (wdsp/nbp.c:64-179, 214-239): the pass band [flow, fhigh] minus the active
 * notches of the database (centres / widths in RF coordinates, offset = tunefreq + shift), designed piecewise with
 * fir_bandpass and summed.  impulse: nc complex.  Bit-identical to the reference's taps. */
int quisk_cuda_nbp_impulse(int nc, double flow, double fhigh, double rate, int wintype, double scale,
                           int n_notches, const double *fcenter, const double *fwidth, const int *active,
                           double tunefreq, double shift, int autoincr, int maxpb,
                           double *impulse, int *numpb, int *havnotch);

/* mp_imp, wdsp/fir.c:317-368: the minimum-phase impulse with the magnitude response of `fir` (N complex in, N out;
 * pfactor 16 and polarity 0 are what calc_fircore passes, firmin.c:328).  N * pfactor must be a power of two. */
int quisk_cuda_mp_imp(int N, const double *fir, double *mpfir, int pfactor, int polarity);

/* ---- emnr: spectral noise reduction "NR2" (wdsp/emnr.c): STFT overlap-add, per-bin gain from a running noise estimate ----
 * create = create_emnr (emnr.c:561-581; fsize 4096 as create_rxa passes it, wintype 0); run = xemnr with run = 1 on bsize
 * complex samples per channel (the real rail is processed, the imaginary rail comes back zero, emnr.c:1059-1064), in place
 * or not; flush = flush_emnr.  gain_method 0 (Gaussian, linear amplitude), 1 (Gaussian, log amplitude), 2 (gamma prior:
 * bilinear look-up in the two 241 x 241 tables GG / GGS of the WDSP distribution -- wdsp/calculus.c, or the `calculus`
 * file calc_emnr prefers, emnr.c:313-323 -- which the HOST supplies once per process with quisk_cuda_emnr_set_tables;
 * they are not part of this library); 3 is not built.  npe_method 0 (minimum statistics), 1, 2.  ae_run: the post-filter. */
typedef struct qcEmnr qcEmnr;
int quisk_cuda_emnr_set_tables(const double *GG, const double *GGS);     /* host pointers, 241 * 241 doubles each, copied */
/* gain method 3 ("trained", the second state of Quisk's NR2 button, quisk.py:6020-6023; emnr.c:965-1010): method 0's gain,
 * applied twice, then a 60 x 60 table of the distribution (wdsp/zetahat.c, or the `zetaHat` file readZetaHat prefers,
 * emnr.c:206-238) says per (gamma, xi) cell whether the bin is speech (gain 1) or not (gain 0); the host hands over the
 * table, its validity map and the four range limits once per process, as readZetaHat returns them */
int quisk_cuda_emnr_set_zeta(const double *zeta_hat, const int *zeta_valid, int rows, int cols,
                             double gamma_min, double gamma_max, double xihat_min, double xihat_max);
int quisk_cuda_emnr_set_train(qcEmnr *e, double zeta_thresh, double t2);  /* SetRXAEMNRtrainZetaThresh / SetRXAEMNRtrainT2, emnr.c:1160-1174 */
qcEmnr *quisk_cuda_emnr_create(int n_channels, int bsize, int fsize, int ovrlp, int rate, int wintype, double gain,
                               int gain_method, int npe_method, int ae_run);
void quisk_cuda_emnr_destroy(qcEmnr *e);
int quisk_cuda_emnr_run(qcEmnr *e, const void *d_in, long in_stride, void *d_out, long out_stride, void *stream);
int quisk_cuda_emnr_flush(qcEmnr *e);
int quisk_cuda_emnr_set_gain_method(qcEmnr *e, int method);             /* SetRXAEMNRgainMethod, emnr.c:1111-1117 */
int quisk_cuda_emnr_set_npe_method(qcEmnr *e, int method);              /* SetRXAEMNRnpeMethod,  emnr.c:1119-1125 */
int quisk_cuda_emnr_set_ae_run(qcEmnr *e, int run);                     /* SetRXAEMNRaeRun,      emnr.c:1127-1133 */

/* ---- analyzer: WDSP's spectrum engine (wdsp/analyzer.c) for a batch of displays that share one configuration: complex or real input,
 * one LO per sub-span, up to four stitched sub-spans, no calibration table (SetAnalyzer with n_fft = 1, fmin = fmax = 0; input_type = typ: 0 real, the I rail alone, 1 complex).
 * create = XCreateAnalyzer (analyzer.c:1140); set = SetAnalyzer (:999-1137; size a power of two 64 .. 8192, overlap in samples,
 * clip / fsclip_low / fsclip_high in bins); set_detector_mode .. set_norm_onehz = SetDisplayDetectorMode (0 peak, 1 rosenfell,
 * 2 average, 3 sample, 4 rms), SetDisplayAverageMode (-1 peak hold, 0 none, 1 recursive, 2 window, 3 recursive on the log),
 * SetDisplayNumAverage, SetDisplayAvBackmult, SetDisplaySampleRate, SetDisplayNormOneHz (:1582-1675), all per pixel output
 * 0 .. 3; spectrum0 = Spectrum0 (:1536) for sub-span ss: d_samples [n_displays][buff_size] complex doubles on the device, stride
 * in complex samples; a sub-span with `size` samples waiting sends one frame and waits for the stitch that uses it (:713-733),
 * the line is detected and averaged as soon as every sub-span has reported; get_pixels = GetPixels
 * (:1315): h_pixels [n_displays][n_pixels] floats (dB), *flag = 1 if the line is new since the last call. */
typedef struct qcAnalyzer qcAnalyzer;
qcAnalyzer *quisk_cuda_analyzer_create(int n_displays, int max_size);
void quisk_cuda_analyzer_destroy(qcAnalyzer *an);
int quisk_cuda_analyzer_set(qcAnalyzer *an, int n_pixout, int input_type, int flip, int size, int buff_size, int window_type, double pi_alpha, int overlap, int clip,
                            double fsclip_low, double fsclip_high, int n_pixels, int n_stitch, int max_writeahead);
int quisk_cuda_analyzer_set_detector_mode(qcAnalyzer *an, int pixout, int mode);
int quisk_cuda_analyzer_set_average_mode(qcAnalyzer *an, int pixout, int mode);
int quisk_cuda_analyzer_set_num_average(qcAnalyzer *an, int pixout, int num);
int quisk_cuda_analyzer_set_av_backmult(qcAnalyzer *an, int pixout, double mult);
int quisk_cuda_analyzer_set_sample_rate(qcAnalyzer *an, int rate);
int quisk_cuda_analyzer_set_norm_onehz(qcAnalyzer *an, int pixout, int norm);
int quisk_cuda_analyzer_spectrum0(qcAnalyzer *an, int ss, const void *d_samples, long stride, void *stream);
int quisk_cuda_analyzer_get_pixels(qcAnalyzer *an, int pixout, float *h_pixels, int *flag);
double quisk_cuda_analyzer_get_enb(qcAnalyzer *an);                          /* GetDisplayENB, analyzer.c:1678 */

/* ---- snba: spectral noise blanker "SNB" (wdsp/snb.c): linear-prediction detection and least-squares interpolation of impulses ----
 * create = create_snba (snb.c:68-119; xsize 256 and asize <= 64 as create_rxa passes them, RXA.c:237-255); run = xsnba with
 * run = 1 on bsize complex samples per channel (real rail processed, imaginary rail back as zero), flush = flush_snba.
 * Resamplers to the internal rate and back are part of the stage when inrate != internalrate. */
typedef struct qcSnba qcSnba;
qcSnba *quisk_cuda_snba_create(int n_channels, int inrate, int internalrate, int bsize, int ovrlp, int xsize, int asize, int npasses,
                               double k1, double k2, int b, int pre, int post, double pmultmin, double out_low_cut, double out_high_cut);
void quisk_cuda_snba_destroy(qcSnba *d);
int quisk_cuda_snba_run(qcSnba *d, const void *d_in, long in_stride, void *d_out, long out_stride, void *stream);
int quisk_cuda_snba_flush(qcSnba *d);

/* ---- fircore: uniformly partitioned overlap-save complex FIR (wdsp/firmin.c:290-430) ---- */
typedef struct qcFircore qcFircore;
/* impulse: HOST, nc complex, the same for every channel (callers bake 1/(2*size) into it exactly as
 * nbp.c:233-237 / bandpass.c:302 do).  mp = 1: the masks are built from quisk_cuda_mp_imp(impulse) (firmin.c:327-328). */
qcFircore *quisk_cuda_fircore_create(int n_channels, int size, int nc, int mp, const double *impulse);
void quisk_cuda_fircore_destroy(qcFircore *f);
/* xfircore for every channel: d_in/d_out [n_channels][stride] complex, `size` samples each. in may equal out. */
int quisk_cuda_fircore_run(qcFircore *f, const void *d_in, long in_stride, void *d_out, long out_stride, void *stream);
/* setImpulse_fircore (firmin.c:448-452): new masks are computed into the idle set and take effect
 * with the next block when update != 0, or at quisk_cuda_fircore_update() (setUpdate_fircore). */
int quisk_cuda_fircore_set_impulse(qcFircore *f, const double *impulse, int update);
int quisk_cuda_fircore_update(qcFircore *f);
int quisk_cuda_fircore_set_mp(qcFircore *f, int mp);                       /* setMp_fircore, firmin.c:469-473 */
int quisk_cuda_fircore_flush(qcFircore *f);
/* The three relatives of fircore that wdsp defines and no live chain instantiates:
 *   firopt  (create_firopt / xfiropt, firmin.c:127-251): partitioned overlap-save, one mask set; taps = calc_firopt's
 *           fir_bandpass(nc, f_low, f_high, rate, wintype, 1, gain).  Run / flush / destroy: the fircore calls.
 *   bps     (create_bps / xbps, bandpass.c:35-105): one overlap-save block, size + 1 taps right-justified from index
 *           size - 1 (fftcv_mults), gain on the spectrum.  Run / flush / destroy: the fircore calls.
 *   firmin  (create_firmin / xfirmin, firmin.c:35-99): time-domain complex-tap ring FIR, bit-exact.  Run with
 *           quisk_cuda_batch_run (n_out = count), flush_firmin = quisk_cuda_batch_reset, quisk_cuda_batch_destroy. */
qcFircore *quisk_cuda_firopt_create(int n_channels, int size, int nc, double f_low, double f_high, int samplerate, int wintype, double gain);
qcFircore *quisk_cuda_bps_create(int n_channels, int size, double f_low, double f_high, int samplerate, int wintype, double gain);
qcBatchFilter *quisk_cuda_firmin_create(int n_channels, int nc, double f_low, double f_high, int samplerate, int wintype, double gain);

/* ---- rational resampler (wdsp/resample.c:121-157) ---- */
typedef struct qcResample qcResample;
qcResample *quisk_cuda_resample_create(int n_channels, int in_rate, int out_rate, double fc, int ncoef, double gain);
void quisk_cuda_resample_destroy(qcResample *r);
int quisk_cuda_resample_count_out(const qcResample *r, int count);
int quisk_cuda_resample_run(qcResample *r, const void *d_in, long in_stride, int count,
                            void *d_out, long out_stride, int *n_out, void *stream);

/* ---- per-sample recurrent stages: one GPU thread walks one channel's block ---- */
typedef struct qcSeqStage qcSeqStage;
/* xshift (wdsp/shift.c:60-86); shift_hz: HOST, one value per channel. */
qcSeqStage *quisk_cuda_shift_create(int n_channels, int rate, const double *shift_hz);
/* xwcpagc (wdsp/wcpAGC.c:161-348) with create_rxa's parameters (RXA.c:337-360) and the preset of
 * SetRXAAGCMode(mode) (wcpAGC.c:370-411); mode 0 = fixed gain. */
qcSeqStage *quisk_cuda_wcpagc_create(int n_channels, int rate, int mode);
qcSeqStage *quisk_cuda_wcpagc_create_fmlim(int n_channels, int rate, double lim_gain);   /* fmd's detector limiter: create_wcpagc as calc_fmd calls it (fmd.c:49-73) */
int quisk_cuda_wcpagc_set_fixed_gain_db(qcSeqStage *s, double gain_db);     /* SetRXAAGCFixed */
int quisk_cuda_wcpagc_set_top_db(qcSeqStage *s, double max_gain_db);        /* SetRXAAGCTop */
/* xamd (wdsp/amd.c:115-239): mode 0 AM envelope, 1 synchronous AM; sbmode 0/1/2; create_rxa's constants. */
qcSeqStage *quisk_cuda_amd_create(int n_channels, int rate, int mode, int levelfade, int sbmode);
/* the PLL + DC removal of xfmd (wdsp/fmd.c:144-170) and, separately, xsnotch (wdsp/iir.c:76-95) */
qcSeqStage *quisk_cuda_fmpll_create(int n_channels, int rate, double deviation, double fmin, double fmax,
                                    double zeta, double omegaN, double tau);
qcSeqStage *quisk_cuda_snotch_create(int n_channels, int rate, double f, double bw);
void quisk_cuda_seq_destroy(qcSeqStage *s);
/* n samples per channel, in may equal out */
int quisk_cuda_seq_run(qcSeqStage *s, const void *d_in, long in_stride, void *d_out, long out_stride, int n, void *stream);
int quisk_cuda_seq_flush(qcSeqStage *s);

/* ---- the RXA chain (wdsp/RXA.c) for a batch of channels ---- */
typedef struct qcRxa qcRxa;
enum qcRxaMode {    /* wdsp/RXA.h: enum rxaMode */
    QC_RXA_LSB = 0, QC_RXA_USB = 1, QC_RXA_DSB = 2, QC_RXA_CWL = 3, QC_RXA_CWU = 4, QC_RXA_FM = 5,
    QC_RXA_AM = 6, QC_RXA_DIGU = 7, QC_RXA_SPEC = 8, QC_RXA_DIGL = 9, QC_RXA_SAM = 10, QC_RXA_DRM = 11
};
/* OpenChannel(ch, in_size, dsp_size, in_rate, dsp_rate, out_rate, type = 0 (RX), ...) + create_rxa. */
qcRxa *quisk_cuda_rxa_create(int n_channels, int in_size, int dsp_size, int in_rate, int dsp_rate, int out_rate);
void quisk_cuda_rxa_destroy(qcRxa *r);
int quisk_cuda_rxa_set_mode(qcRxa *r, int mode);                            /* SetRXAMode,        RXA.c:749-787  */
int quisk_cuda_rxa_set_passband(qcRxa *r, double f_low, double f_high);     /* RXASetPassband,    RXA.c:927-932  */
int quisk_cuda_rxa_set_nc(qcRxa *r, int nc);                                /* RXASetNC,          RXA.c:935-946  */
int quisk_cuda_rxa_set_agc_mode(qcRxa *r, int mode);                        /* SetRXAAGCMode,     wcpAGC.c:370   */
int quisk_cuda_rxa_set_agc_fixed(qcRxa *r, double gain_db);                 /* SetRXAAGCFixed                    */
int quisk_cuda_rxa_set_shift(qcRxa *r, int run, const double *shift_hz);    /* SetRXAShiftRun / SetRXAShiftFreq  */
int quisk_cuda_rxa_set_nbp_run(qcRxa *r, int run);                          /* RXANBPSetRun                      */
/* OpenChannel's tdelayup / tslewup (channel.c:76-104; Quisk passes 0.010 and 0.025, quisk_wdsp.py:79-80) and the
 * upflag it raises: quisk_cuda_rxa_fexchange0 then runs upslew0's state machine (iobuffs.c:98-160) per channel on the
 * device -- zeros up to and including the first non-zero sample, tdelayup of zeros, a raised-cosine ramp over
 * tslewup.  A new handle starts armed with both times 0 (the first non-zero sample is still swallowed).  The
 * down-slew / flush of SetChannelState(ch, 0): quisk_cuda_rxa_set_channel_state below. */
int quisk_cuda_rxa_set_slew(qcRxa *rxa, double tdelayup, double tslewup);
/* sip1 of create_rxa (RXA.c:392-401; xsiphon mode 0, siphon.c:96-129): the last 4096 samples of midbuff per channel,
 * kept by default like the reference (run = 1).  quisk_cuda_rxa_get_siphon = RXAGetaSipF (complex_out 0: real parts)
 * / RXAGetaSipF1 (complex_out 1: I/Q pairs) for all channels: h_out[channel][size] floats, newest sample last. */
int quisk_cuda_rxa_set_siphon_run(qcRxa *rxa, int run);
int quisk_cuda_rxa_get_siphon(qcRxa *rxa, float *h_out, int size, int complex_out);
/* Notch database of nbp0 (RXANBPAddNotch / DeleteNotch / SetNotchesRun / SetTuneFrequency / SetShiftFrequency,
 * nbp.c:359-513): one database per handle, applied to every channel of the batch. */
int quisk_cuda_rxa_nbp_add_notch(qcRxa *rxa, int notch, double fcenter, double fwidth, int active);
int quisk_cuda_rxa_nbp_delete_notch(qcRxa *rxa, int notch);
int quisk_cuda_rxa_nbp_set_notches_run(qcRxa *rxa, int run);
int quisk_cuda_rxa_nbp_set_tune_frequency(qcRxa *rxa, double tunefreq);
int quisk_cuda_rxa_nbp_set_shift_frequency(qcRxa *rxa, double shift);
int quisk_cuda_rxa_set_snba_run(qcRxa *rxa, int run);                        /* SetRXASNBARun, snb.c:579-593 (bpsnba and bp1 follow: RXAbpsnbaCheck / Set, RXAbp1Check / Set) */
int quisk_cuda_rxa_set_emnr_run(qcRxa *rxa, int run);                        /* SetRXAEMNRRun, emnr.c:1096-1109 (bp1 follows: RXAbp1Check / RXAbp1Set) */
int quisk_cuda_rxa_set_emnr_gain_method(qcRxa *rxa, int method);
int quisk_cuda_rxa_set_emnr_npe_method(qcRxa *rxa, int method);
int quisk_cuda_rxa_set_emnr_ae_run(qcRxa *rxa, int run);
int quisk_cuda_rxa_set_emnr_train(qcRxa *rxa, int what, double value);        /* 0: SetRXAEMNRtrainZetaThresh, 1: SetRXAEMNRtrainT2 (emnr.c:1160-1174) */
int quisk_cuda_rxa_set_emnr_position(qcRxa *rxa, int position);              /* SetRXAEMNRPosition, emnr.c:1135-1142: 0 in front of the AGC, 1 behind it */
int quisk_cuda_rxa_set_fm_lim_run(qcRxa *rxa, int run);                      /* SetRXAFMLimRun,  fmd.c:337-348: the FM detector limiter (a wcpAGC, fmd.c:49-73) */
int quisk_cuda_rxa_set_fm_lim_gain(qcRxa *rxa, double gain_db);              /* SetRXAFMLimGain, fmd.c:350-363 */
int quisk_cuda_rxa_set_mp(qcRxa *rxa, int mp);                               /* RXASetMP, RXA.c:949-958           */
int quisk_cuda_rxa_set_panel_gain(qcRxa *r, double gain1);                  /* SetRXAPanelGain1                  */
int quisk_cuda_rxa_in_size(const qcRxa *r);      /* dsp_insize: samples per channel consumed per xrxa  */
int quisk_cuda_rxa_out_size(const qcRxa *r);     /* dsp_outsize: samples per channel produced per xrxa */
/* xrxa (RXA.c:561-598) for one DSP block of every channel. d_in [n_channels][in_stride], d_out likewise. */
int quisk_cuda_rxa_xrxa(qcRxa *r, const void *d_in, long in_stride, void *d_out, long out_stride, void *stream);
/* The same for n_blocks consecutive DSP blocks per channel (block b of a channel at d_in + b * dsp_insize samples,
 * d_out + b * dsp_outsize): one launch chain instead of one per block -- the state carried between blocks is the same. */
int quisk_cuda_rxa_xrxa_multi(qcRxa *r, const void *d_in, long in_stride, void *d_out, long out_stride, int n_blocks, void *stream);
#define QC_RXA_OPT_FUSED 1     /* 1 (default): configurations the single-kernel chain covers (wdsp_rxa_fused.cu) use it; 0: one kernel per stage */
int quisk_cuda_rxa_set_option(qcRxa *r, int option, int value);
/* fexchange0 (wdsp/iobuffs.c:464-516) with HOST buffers for all channels at once: h_in [n_channels][in_size],
 * h_out [n_channels][out_size] interleaved complex.  The reference's two pseudo-rings (create_iobuffs, iobuffs.c:385-420)
 * and its ring arithmetic are reproduced, with the DSP thread's turns (dexchange + xrxa, main.c:40-63) run inside the
 * call: any in_size / dsp_insize ratio (several DSP turns per call, or several calls per turn), the two-DSP-buffer
 * latency, the up-slew on the way in and the down-slew on the way out.  *error = 0, or -2 with zeros returned when no
 * output is available (bfo = 0 before the rings have filled).  When the exchange is off (SetChannelState(0) finished)
 * the call returns at once and leaves h_out alone, as the reference does. */
int quisk_cuda_rxa_fexchange0(qcRxa *r, const double *h_in, double *h_out, int *error);
/* SetChannelState (channel.c:262-300) for the batch; returns the prior state.  state 0: the following exchange calls
 * run downslew0 (tdelaydown of signal, a raised-cosine ramp over tslewdown, out_size + 1 zeros), then the exchange
 * switches off and the channel is flushed (flush_iobuffs + flush_rxa: note that flush_wcpagc and flush_amd keep
 * their loop state, as in the reference).  state 1: upflag up, exchange on. */
int quisk_cuda_rxa_set_channel_state(qcRxa *r, int state, int dmode);
int quisk_cuda_rxa_set_slew_down(qcRxa *r, double tdelaydown, double tslewdown);    /* OpenChannel's tdelaydown / tslewdown */
int quisk_cuda_rxa_set_bfo(qcRxa *r, int bfo);                                      /* OpenChannel's bfo (block for output) */
int quisk_cuda_rxa_exchange_sizes(const qcRxa *r, int *in_size, int *out_size);
/* meters (wdsp/meter.c:75-107): which = 0 ADC, 1 S, 2 AGC; av/pk/gain in dB, HOST arrays of n_channels (any may be NULL) */
int quisk_cuda_rxa_get_meter(qcRxa *r, int which, double *av, double *pk, double *gain);

/* ------------------------------------------------------------------------------------------------------------------
 * The reference's own WDSP entry points, by channel number (wdsp_compat.cu): the symbols quisk_wdsp.py resolves in
 * libwdsp.so (quisk_wdsp.py:28-99, quisk.py:6017-6053) with the exact signatures of wdsp/channel.c:76, iobuffs.c:464,
 * version.c:4, RXA.c:749-958, shift.c:112-128, nbp.c:359-527, wcpAGC.c:370-548, patchpanel.c:125-156, amd.c:279-293,
 * meter.c:120, siphon.c:183-211 -- `ctypes.CDLL("libquisk_cuda.so")` in place of "./wdsp/libwdsp.so" (INTEGRATION.md
 * section 2).  One single-channel chain per open channel number, MAX_CHANNELS = 32.  The stages this library does not
 * build (squelches, ANF, ANR; EMNR and SNBA ARE built) are accepted switched off and refused with a message when switched on.
 * ------------------------------------------------------------------------------------------------------------------ */
int GetWDSPVersion(void);
void OpenChannel(int channel, int in_size, int dsp_size, int input_samplerate, int dsp_rate, int output_samplerate,
                 int type, int state, double tdelayup, double tslewup, double tdelaydown, double tslewdown, int bfo);
void CloseChannel(int channel);
int SetChannelState(int channel, int state, int dmode);
void fexchange0(int channel, double *in, double *out, int *error);
void SetRXAMode(int channel, int mode);
void RXASetPassband(int channel, double f_low, double f_high);
void RXASetNC(int channel, int nc);
void RXASetMP(int channel, int mp);
void SetRXAShiftRun(int channel, int run);
void SetRXAShiftFreq(int channel, double fshift);
void RXANBPSetRun(int channel, int run);
void RXANBPSetFreqs(int channel, double flow, double fhigh);
int RXANBPAddNotch(int channel, int notch, double fcenter, double fwidth, int active);
int RXANBPDeleteNotch(int channel, int notch);
void RXANBPGetNumNotches(int channel, int *nnotches);
void RXANBPSetNotchesRun(int channel, int run);
void RXANBPSetTuneFrequency(int channel, double tunefreq);
void RXANBPSetShiftFrequency(int channel, double shift);
void SetRXAAGCMode(int channel, int mode);
void SetRXAAGCFixed(int channel, double fixed_agc);
void SetRXAAGCTop(int channel, double max_agc);
void SetRXAPanelRun(int channel, int run);
void SetRXAPanelGain1(int channel, double gain);
void SetRXAPanelGain2(int channel, double gainI, double gainQ);
void SetRXAAMDSBMode(int channel, int sbmode);
void SetRXAAMDFadeLevel(int channel, int levelfade);
double GetRXAMeter(int channel, int mt);
void RXAGetaSipF(int channel, float *out, int size);
void RXAGetaSipF1(int channel, float *out, int size);
void SetRXAFMLimRun(int channel, int run);
void SetRXAFMLimGain(int channel, double gaindB);
void SetRXAAMSQRun(int channel, int run);
void SetRXAFMSQRun(int channel, int run);
void SetRXAEMNRRun(int channel, int run);
void SetRXAEMNRgainMethod(int channel, int method);
void SetRXAEMNRnpeMethod(int channel, int method);
void SetRXAEMNRaeRun(int channel, int run);
void SetRXAEMNRPosition(int channel, int position);
void SetRXAEMNRtrainZetaThresh(int channel, double thresh);    /* emnr.c:1160-1166 */
void SetRXAEMNRtrainT2(int channel, double t2);                 /* emnr.c:1168-1174 */
void SetRXASNBARun(int channel, int run);
void SetRXAANFRun(int channel, int run);
void SetRXAANRRun(int channel, int run);
/* Quisk's side of the boundary: wdspFexchange0 (quisk_wdsp.c:24-73: arbitrary sample counts, 1 / CLIP32 scaling,
 * re-blocking to in_size, returns the count now in cSamples) and the C half of quisk_wdsp_set_parameter
 * (quisk_wdsp.c:75-92; in_size <= 0 / in_use < 0: leave alone). */
int wdspFexchange0(int channel, quisk_cd *cSamples, int nSamples);
void quisk_cuda_wdsp_set_parameter(int channel, int in_size, int in_use);

#ifdef __cplusplus
}
#endif
#endif /* QUISK_CUDA_WDSP_H */
