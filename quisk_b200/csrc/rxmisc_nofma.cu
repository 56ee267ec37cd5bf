// quisk_b200/csrc/rxmisc_nofma.cu -- the remaining small pieces of the Quisk receive path:
//   process_agc  (quisk.c:2162-2287)  Quisk's own look-ahead AGC at playback rate (15 ms FIFO)
//   cFracDecim   (quisk.c:622-665)    fractional decimation by 4-point Lagrange interpolation
//   get_bandscope (quisk.c:4957-5011) + copy2pixels (quisk.c:4932-4955): real-input spectrum display
//   NoiseBlanker (quisk.c:679-784)    impulse blanker on the raw samples in front of the tuning stage (optional)
//   dAutoNotch (quisk.c:786-963)     automatic notch of the one or two strongest steady carriers in the SSB audio (optional)
//   ssb_squelch + d_delay (quisk.c:1056-1180)  spectral-flatness squelch on the SSB audio at the filter rate (optional)
// process_agc and cFracDecim are scalar recurrences: one CTA per channel, block staged in shared
// memory, lane 0 walks it (compiled with --fmad=false so the state follows the reference bit for bit).
#include "fft_device.cuh"
#include <cmath>

namespace qc {

static const double kCLIP32 = 2147483647.0;

// state: 0 index_read 1 index_start 2 is_clipping 3 themax 4 gain 5 delta 6 target_gain
struct AgcPar { int buf_size, is_cpx; double max_out, time_release, release_gain; };

__global__ void agc_kernel(cd *samples, long stride, int n, int C, double *state, cd *fifo, AgcPar p)
{
    extern __shared__ double sm_raw[];
    cd *sx = reinterpret_cast<cd *>(sm_raw);            // [n] block
    cd *sf = sx + n;                                    // [buf_size] FIFO
    const int c = blockIdx.x;
    cd *g = samples + (size_t)c * stride;
    cd *gf = fifo + (size_t)c * p.buf_size;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sx[i] = g[i];
    for (int i = threadIdx.x; i < p.buf_size; i += blockDim.x) sf[i] = gf[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        double *st = state + (size_t)c * 8;
        int index_read = (int)st[0], index_start = (int)st[1], is_clipping = (int)st[2];
        double themax = st[3], gain = st[4], delta = st[5], target_gain = st[6];
        for (int i = 0; i < n; i++) {
            const cd csample = sx[i];
            cd o = make_double2(sf[index_read].x * gain, sf[index_read].y * gain);      // FIFO output
            const double out_magn = p.is_cpx ? hypot(o.x, o.y) : fabs(o.x);
            if (out_magn > kCLIP32) { o.x /= out_magn; o.y /= out_magn; }               // quisk.c:2203-2204
            sx[i] = o;
            sf[index_read] = csample;
            const double buf_magn = p.is_cpx ? hypot(csample.x, csample.y) : fabs(csample.x);
            if (is_clipping == 0) {
                if (buf_magn * gain > p.max_out * kCLIP32) {
                    target_gain = p.max_out * kCLIP32 / buf_magn;
                    delta = (gain - target_gain) / p.buf_size;
                    is_clipping = 1;
                    themax = buf_magn;
                    gain -= delta;
                } else if (index_read == index_start) {
                    const double clip_gain = p.max_out * kCLIP32 / themax;
                    target_gain = p.release_gain > clip_gain ? clip_gain : p.release_gain;
                    themax = buf_magn;
                    gain = gain * (1.0 - p.time_release) + target_gain * p.time_release;
                } else {
                    if (themax < buf_magn) themax = buf_magn;
                    gain = gain * (1.0 - p.time_release) + target_gain * p.time_release;
                }
            } else {
                if (buf_magn > themax) {
                    themax = buf_magn;
                    target_gain = p.max_out * kCLIP32 / buf_magn;
                    const double dtmp = (gain - target_gain) / p.buf_size;
                    if (dtmp > delta) delta = dtmp;
                }
                gain -= delta;
                if (gain <= target_gain) {
                    is_clipping = 0;
                    gain = target_gain;
                    themax = buf_magn;
                    index_start = index_read;
                }
            }
            if (++index_read >= p.buf_size) index_read = 0;
        }
        st[0] = index_read; st[1] = index_start; st[2] = is_clipping; st[3] = themax; st[4] = gain; st[5] = delta; st[6] = target_gain;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) g[i] = sx[i];
    for (int i = threadIdx.x; i < p.buf_size; i += blockDim.x) gf[i] = sf[i];
}

// cFracDecim: state 0 dindex, 1..6 c0 c1 c2 (complex).  The index walk does not depend on the data, so the host
// knows every output count; the device just follows the same walk.
__global__ void fracdecim_kernel(const cd *in, long is, cd *out, long os, int n, int C, double *state, double fdecim)
{
    extern __shared__ double sm_raw[];
    cd *sx = reinterpret_cast<cd *>(sm_raw);
    const int c = blockIdx.x;
    const cd *g = in + (size_t)c * is;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sx[i] = g[i];
    __syncthreads();
    __shared__ int s_nout;
    if (threadIdx.x == 0) {
        double *st = state + (size_t)c * 8;
        double dindex = st[0];
        cd c0 = make_double2(st[1], st[2]), c1 = make_double2(st[3], st[4]), c2 = make_double2(st[5], st[6]);
        int nout = 0;
        for (int i = 0; i < n; i++) {
            const cd c3 = sx[i];
            if (dindex < 2) {
                const double xm0 = dindex - 0, xm1 = dindex - 1, xm2 = dindex - 2, xm3 = dindex - 3;
                // (xm1*xm2*xm3*c0 / -6 + xm0*xm2*xm3*c1 / 2 + xm0*xm1*xm3*c2 / -2 + xm0*xm1*xm2*c3 / 6), quisk.c:645-647
                const double w0 = xm1 * xm2 * xm3, w1 = xm0 * xm2 * xm3, w2 = xm0 * xm1 * xm3, w3 = xm0 * xm1 * xm2;
                cd o;
                o.x = ((w0 * c0.x / -6.0 + w1 * c1.x / 2.0) + w2 * c2.x / -2.0) + w3 * c3.x / 6.0;
                o.y = ((w0 * c0.y / -6.0 + w1 * c1.y / 2.0) + w2 * c2.y / -2.0) + w3 * c3.y / 6.0;
                sx[nout++] = o;                         // nout <= i: in place like the reference
                dindex += fdecim - 1;
            } else {
                dindex -= 1;
            }
            c0 = c1; c1 = c2; c2 = c3;
        }
        st[0] = dindex; st[1] = c0.x; st[2] = c0.y; st[3] = c1.x; st[4] = c1.y; st[5] = c2.x; st[6] = c2.y;
        s_nout = nout;
    }
    __syncthreads();
    cd *y = out + (size_t)c * os;
    for (int i = threadIdx.x; i < s_nout; i += blockDim.x) y[i] = sx[i];
}

// bandscope: real block * Hann -> FFT (complex transform of the real block) -> |X[0..N/2]| accumulate
__global__ void __launch_bounds__(256) bandscope_accumulate_kernel(const double *blocks, long stream_stride, int n_blocks, int n,
                                                                    const cd *tw, const double *window, double *avg, double *the_max)
{
    extern __shared__ double sm_raw[];
    cd *twl = reinterpret_cast<cd *>(sm_raw);
    cd *s = twl + fft_tw_entries(n);
    fft_stage_twiddles(twl, tw, n);
    const int stream = blockIdx.x, lane = threadIdx.x, lanes = blockDim.x;
    const int L = n / 2 + 1;
    double *a = avg + (size_t)stream * (L + 1);
    double mx = 0.0;
    for (int b = 0; b < n_blocks; b++) {
        const double *src = blocks + (size_t)stream * stream_stride + (size_t)b * n;
        for (int i = lane; i < n; i += lanes) {
            const double v = src[i];
            mx = fmax(mx, fabs(v));
            s[fsw(i)] = make_double2(v * window[i], 0.0);
        }
        __syncthreads();
        fft_smem(s, n, twl, -1, lane, lanes);
        for (int i = lane; i < L; i += lanes) { const cd X = s[fsw(i)]; a[i] += sqrt(X.x * X.x + X.y * X.y); }   // cabs, quisk.c:4981
        __syncthreads();
    }
    // the_max (hermes_adc_level, quisk.c:4972-4974): block-wide maximum of |sample|
    __shared__ double s_mx[32];
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((lane & 31) == 0) s_mx[lane >> 5] = mx;
    __syncthreads();
    if (lane == 0) {
        double m = the_max[stream];
        for (int w = 0; w < (lanes + 31) / 32; w++) m = fmax(m, s_mx[w]);
        the_max[stream] = m;
    }
}

// copy2pixels (quisk.c:4932-4955) + scale + 20 log10 (quisk.c:4991-4999); one thread per pixel
__global__ void bandscope_graph_kernel(const double *avg, int L, int graph_width, double zoom, double deltaf, double rate,
                                       double scale, double *graph)
{
    const int stream = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= graph_width) return;
    const double *fft = avg + (size_t)stream * (L + 1);
    const int fft_size = L;
    const double f1 = deltaf + rate / 2.0 * (1.0 - zoom);
    const double d1 = fft_size / rate * (f1 + (double)i / graph_width * zoom * rate);
    const double d2 = fft_size / rate * (f1 + (double)(i + 1) / graph_width * zoom * rate);
    const int j1 = (int)floor(d1), j2 = (int)floor(d2);
    double sample;
    if (j1 == j2) sample = (d2 - d1) * fft[j1];
    else {
        sample = (j1 + 1 - d1) * fft[j1];
        for (int j = j1 + 1; j < j2; j++) sample += fft[j];
        sample += (d2 - j2) * fft[j2];
    }
    sample = sample * scale;
    graph[(size_t)stream * graph_width + i] = sample <= 1E-10 ? -200.0 : 20.0 * log10(sample);
}

// ---- NoiseBlanker (quisk.c:679-784).  Per channel: a delay line of save_size = 3 * hwindow samples whose entries are
// edited while they wait (ramp down in front of a pulse, zero while pulses last, ramp up afterwards), and a running sum
// of the last save_size magnitudes that decides what a pulse is.  The decision depends on the magnitudes only, never on
// the blanking state, so a chunk is processed in three phases: (1) all lanes: |x| ; lane 0: the running sum in the
// reference's order (two dependent, separately rounded additions per sample, quisk.c:733-735); (2) all lanes: the
// threshold test with the reference's division (quisk.c:736); (3) if the chunk holds no pulse and no blanking is in
// progress the delay line is a ring that all lanes move (scaling the entering samples while a ramp lasts), otherwise lane 0 walks the state machine
// (quisk.c:740-765) sample by sample.  Chunks are at most save_size long, so that a ring slot is touched once per chunk.
// state per channel: 0 index 1 win_index 2 state 3 save_sum
struct NbPar { int save_size, hwindow, chunk; double limit; };

__global__ void __launch_bounds__(128) nb_kernel(cd *samples, long stride, int n, double *state, cd *csaved, double *dsaved, NbPar p)
{
    extern __shared__ double sm_raw[];
    cd *sc = reinterpret_cast<cd *>(sm_raw);            // [save_size] delay line
    cd *sx = sc + p.save_size;                          // [chunk] samples in, delayed samples out
    double *sd = reinterpret_cast<double *>(sx + p.chunk);      // [save_size] magnitudes
    double *smag = sd + p.save_size;                    // [chunk] |x|, then the running sum after the sample
    __shared__ int s_index, s_win, s_state;
    const int c = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    cd *g = samples + (size_t)c * stride;
    cd *gc = csaved + (size_t)c * p.save_size;
    double *gd = dsaved + (size_t)c * p.save_size;
    double *st = state + (size_t)c * 4;
    for (int i = tid; i < p.save_size; i += NT) { sc[i] = gc[i]; sd[i] = gd[i]; }
    double save_sum = st[3];                            // lane 0's copy is the live one
    if (tid == 0) { s_index = (int)st[0]; s_win = (int)st[1]; s_state = (int)st[2]; }
    __syncthreads();
    const double dsize = (double)p.save_size;
    for (int base = 0; base < n; base += p.chunk) {
        const int m = min(p.chunk, n - base);
        for (int i = tid; i < m; i += NT) { const cd x = g[base + i]; sx[i] = x; smag[i] = hypot(x.x, x.y); }
        __syncthreads();
        const int index0 = s_index;
        if (tid == 0) {
            int k = index0, i = 0;
            // eight samples at a time: all sixteen loads first (the slots are distinct: a chunk is at most one lap of the
            // ring), then the dependent chain of sixteen additions in registers, then the stores
            for (; i + 8 <= m; i += 8) {
                double mg[8], od[8]; int kk[8];
#pragma unroll
                for (int j = 0; j < 8; j++) { kk[j] = k + j >= p.save_size ? k + j - p.save_size : k + j; mg[j] = smag[i + j]; od[j] = sd[kk[j]]; }
#pragma unroll
                for (int j = 0; j < 8; j++) { save_sum -= od[j]; save_sum += mg[j]; od[j] = save_sum; }
#pragma unroll
                for (int j = 0; j < 8; j++) { sd[kk[j]] = mg[j]; smag[i + j] = od[j]; }
                k = k + 8 >= p.save_size ? k + 8 - p.save_size : k + 8;
            }
            for (; i < m; i++) {
                const double mag = smag[i];
                save_sum -= sd[k];
                sd[k] = mag;
                save_sum += mag;
                smag[i] = save_sum;
                if (++k >= p.save_size) k = 0;
            }
        }
        __syncthreads();
        int any = 0;
        unsigned char *flag = reinterpret_cast<unsigned char *>(smag + p.chunk);
        for (int i = tid; i < m; i += NT) {
            int k = index0 + i; if (k >= p.save_size) k -= p.save_size;
            const int is_pulse = sd[k] <= smag[i] / dsize * p.limit ? 0 : 1;
            flag[i] = (unsigned char)is_pulse;
            any |= is_pulse;
        }
        any = __syncthreads_or(any | s_state);
        if (!any) {
            // no pulse and not blanking: a plain ring, except that a ramp left over from the last pulse scales the
            // samples entering it (sample i by (win_index + i) / hwindow while that is below 1, quisk.c:752-756)
            const int w0 = s_win;
            const double hw = (double)p.hwindow;
            for (int i = tid; i < m; i += NT) {
                int k = index0 + i; if (k >= p.save_size) k -= p.save_size;
                cd x = sx[i];
                if (w0 && w0 + i < p.hwindow) { const double w = (double)(w0 + i) / hw; x.x *= w; x.y *= w; }
                sx[i] = sc[k];
                sc[k] = x;
            }
            __syncthreads();
            if (tid == 0) {
                int k = index0 + m; if (k >= p.save_size) k -= p.save_size;
                s_index = k;
                if (w0) s_win = w0 + m < p.hwindow ? w0 + m : 0;
            }
        } else if (tid == 0) {
            int index = index0, win_index = s_win, state_ = s_state;
            const double hw = (double)p.hwindow;
            for (int i = 0; i < m; i++) {
                const cd samp = sx[i];
                sx[i] = sc[index];
                sc[index] = samp;
                const int is_pulse = flag[i];
                if (state_ == 0) {
                    if (is_pulse) {
                        state_ = 1;
                        int k = index;
                        for (int j = 0; j < p.hwindow; j++) {
                            const double w = (double)j / hw;
                            sc[k].x *= w; sc[k].y *= w;
                            if (--k < 0) k = p.save_size - 1;
                        }
                    } else if (win_index) {
                        const double w = (double)win_index / hw;
                        sc[index].x *= w; sc[index].y *= w;
                        if (++win_index >= p.hwindow) win_index = 0;
                    }
                } else {
                    sc[index] = make_double2(0.0, 0.0);
                    if (!is_pulse) { state_ = 0; win_index = 1; }
                }
                if (++index >= p.save_size) index = 0;
            }
            s_index = index; s_win = win_index; s_state = state_;
        }
        __syncthreads();
        for (int i = tid; i < m; i += NT) g[base + i] = sx[i];
        __syncthreads();
    }
    for (int i = tid; i < p.save_size; i += NT) { gc[i] = sc[i]; gd[i] = sd[i]; }
    if (tid == 0) { st[0] = s_index; st[1] = s_win; st[2] = s_state; st[3] = save_sum; }
}

// ---- ssb_squelch (quisk.c:1086-1180) + d_delay (quisk.c:1056-1084).  Per channel the audio samples fill a 512-point
// frame; a full frame is windowed (Hann), transformed, and the bins between 300 Hz and 300 Hz + bandwidth give
// ratio = log(arithmetic mean of |X|^2) - mean(log |X|^2) (0.57 for band noise, more for speech); ratio above
// level * 0.005 re-arms a one-second timer `sq_open`, the squelch is active when the timer has run out.  The frame
// fill index is the same for every channel (host side), the timers are per channel.  The audio is then delayed by
// one frame (512 samples) so that the decision is in time for the samples it was made from.
// state per channel: 0 sq_open 1 squelch_active
#define QC_SQ_N 512
struct SqPar { int samp_rate, bw1, bw2; double thresh; };

__global__ void __launch_bounds__(32) ssb_squelch_kernel(double *audio, long stride, int n, int index0, int do_squelch, int didx0, SqPar p,
                                                         const cd *tw, const double *window, double *infft, double *delay, int *state)
{
    extern __shared__ double sm_raw[];
    cd *twl = reinterpret_cast<cd *>(sm_raw);
    cd *s = twl + fft_tw_entries(QC_SQ_N);
    double *fbuf = reinterpret_cast<double *>(s + QC_SQ_N);     // [512] frame being filled
    double *blk = fbuf + QC_SQ_N;                               // [n] this call's samples
    const int c = blockIdx.x, lane = threadIdx.x, lanes = blockDim.x;
    double *g = audio + (size_t)c * stride;
    double *gf = infft + (size_t)c * QC_SQ_N;
    double *gd = delay + (size_t)c * QC_SQ_N;
    for (int i = lane; i < n; i += lanes) blk[i] = g[i];
    if (do_squelch) {
        fft_stage_twiddles(twl, tw, QC_SQ_N);
        for (int i = lane; i < index0; i += lanes) fbuf[i] = gf[i];
        int sq_open = state[c * 2];
        __syncthreads();
        int idx = index0, pos = 0;
        while (pos < n) {
            const int take = min(QC_SQ_N - idx, n - pos);
            for (int j = lane; j < take; j += lanes) fbuf[idx + j] = blk[pos + j];
            idx += take; pos += take;
            __syncthreads();
            if (idx == QC_SQ_N) {
                for (int i = lane; i < QC_SQ_N; i += lanes) s[fsw(i)] = make_double2(fbuf[i] * window[i], 0.0);
                __syncthreads();
                fft_smem(s, QC_SQ_N, twl, -1, lane, lanes);
                if (lane == 0) {                                // the reference's bin order, quisk.c:1127-1134
                    double arith_avg = 0.0, geom_avg = 0.0, ratio;
                    for (int i = p.bw1; i < p.bw2; i++) {
                        const cd X = s[fsw(i)];
                        const double re = X.x / 32767.0, im = X.y / 32767.0;        // CLIP16, quisk.h:14
                        const double d = re * re + im * im;
                        if (d > 1E-4) { arith_avg += d; geom_avg += log(d); }
                    }
                    if (arith_avg > 1E-4) {
                        const int bw = p.bw2 - p.bw1;
                        arith_avg = log(arith_avg / bw);
                        geom_avg /= bw;
                        ratio = arith_avg - geom_avg;
                    } else {
                        ratio = 1.0;
                    }
                    if (ratio > p.thresh) sq_open = p.samp_rate;
                }
                idx = 0;
                __syncthreads();
            }
        }
        for (int i = lane; i < idx; i += lanes) gf[i] = fbuf[i];
        if (lane == 0) {
            sq_open -= n;
            if (sq_open < 0) sq_open = 0;
            state[c * 2] = sq_open;
            state[c * 2 + 1] = sq_open == 0;
        }
    }
    __syncthreads();
    // d_delay: a 512-sample FIFO.  Output i is the ring's old entry for i < 512 and this call's sample i - 512 after that.
    for (int i = lane; i < n; i += lanes) {
        const int k = (didx0 + i) & (QC_SQ_N - 1);
        g[i] = i < QC_SQ_N ? gd[k] : blk[i - QC_SQ_N];
    }
    __syncthreads();
    for (int i = max(0, n - QC_SQ_N) + lane; i < n; i += lanes) gd[(didx0 + i) & (QC_SQ_N - 1)] = blk[i];
}

// ---- dAutoNotch (quisk.c:786-963): overlap-save filtering of the audio with a notch filter that follows the one or two
// strongest steady spectral lines.  Frames of 2048 samples advance by 1538 (the filter has 511 taps); per frame:
// forward transform, |X| into a running average per bin (0.5 / 0.5), first and second maximum outside the side-tone
// band, hysteresis counters deciding whether each maximum is steady, and -- only when the (i1, i2) signature changes --
// a new filter: ones with zeroed bands on 256 bins, 512-point inverse transform, centred and mirrored into 511 taps,
// Hann, zero-padded 2048-point forward transform.  Then X *= H, inverse transform, divide by 102 (quisk.c:958).
// One CTA of 128 lanes per channel, frames in sequence inside the kernel; the two scans and the counters are walked by
// one lane in the reference's order.  Macro arithmetic as the reference's unparenthesised #defines give it
// (NOTCH_FILTER_DESIGN_SIZE = 2048 / 4 inside expressions, quisk.c:788).
#define QC_AN_N 2048
#define QC_AN_BINS 1025
#define QC_AN_START 510
#define QC_AN_OUT 1538
// ints per channel: 0 old1 1 count1 2 old2 3 count2 4 fltrSig
struct AnPar { int delta_sig, delta_i1, signal, half_width; };

__global__ void __launch_bounds__(128) autonotch_kernel(double *audio, long stride, int n, int index0, AnPar p, const cd *tw2k, const cd *tw512,
                                                        const double *window, double *g_in, double *g_out, double *g_avg, cd *g_fltr, int *g_ist)
{
    extern __shared__ double sm_raw[];
    cd *twl2k = reinterpret_cast<cd *>(sm_raw);
    cd *twl512 = twl2k + fft_tw_entries(QC_AN_N);
    cd *s = twl512 + fft_tw_entries(512);               // [2048] transform buffer
    cd *X = s + QC_AN_N;                                // [1025] spectrum of the frame
    double *fo = reinterpret_cast<double *>(X + QC_AN_BINS + 1);        // [512] filter design scratch
    __shared__ int s_design, s_i1, s_i2, s_c1, s_c2;
    const int c = blockIdx.x, tid = threadIdx.x, NT = 128;
    double *a = audio + (size_t)c * stride;
    double *din = g_in + (size_t)c * QC_AN_N, *dout = g_out + (size_t)c * QC_AN_N, *avg = g_avg + (size_t)c * QC_AN_BINS;
    cd *fl = g_fltr + (size_t)c * QC_AN_BINS;
    int *ist = g_ist + (size_t)c * 8;
    fft_stage_twiddles(twl2k, tw2k, QC_AN_N);
    fft_stage_twiddles(twl512, tw512, 512);
    __syncthreads();
    int idx = index0, pos = 0;
    while (pos < n) {
        const int take = min(QC_AN_N - idx, n - pos);
        for (int j = tid; j < take; j += NT) { const double x = a[pos + j]; a[pos + j] = dout[idx + j]; din[idx + j] = x; }
        idx += take; pos += take;
        __syncthreads();
        if (idx < QC_AN_N) break;
        idx = QC_AN_START;
        // forward transform of the frame (real input as a complex frame)
        for (int i = tid; i < QC_AN_N; i += NT) s[fsw(i)] = make_double2(din[i], 0.0);
        __syncthreads();
        fft_smem<1>(s, QC_AN_N, twl2k, -1, tid, NT);
        for (int i = tid; i < QC_AN_BINS; i += NT) {
            const cd v = s[fsw(i)];
            X[i] = v;
            avg[i] = 0.5 * avg[i] + 0.5 * hypot(v.x, v.y);      // quisk.c:857-860
        }
        __syncthreads();
        if (tid == 0) {
            int old1 = ist[0], count1 = ist[1], old2 = ist[2], count2 = ist[3], fltrSig = ist[4];
            double d1 = 0; int i1 = 0;
            for (int i = 0; i < QC_AN_BINS; i++)
                if (abs(i - p.signal) > p.delta_sig && avg[i] > d1) { d1 = avg[i]; i1 = i; }
            if (abs(i1 - old1) < 3) count1++; else count1--;
            if (count1 > 4) count1 = 4; else if (count1 < -1) count1 = -1;
            if (count1 < 0) old1 = i1;
            double d2 = 0; int i2 = 0;
            for (int i = 0; i < QC_AN_BINS; i++)
                if (abs(i - p.signal) > p.delta_sig && abs(i - i1) > p.delta_i1 && avg[i] > d2) { d2 = avg[i]; i2 = i; }
            if (abs(i2 - old2) < 3) count2++; else count2--;
            if (count2 > 4) count2 = 4; else if (count2 < -2) count2 = -2;
            if (count2 < 0) old2 = i2;
            int k;
            if (count1 > 0 && count2 > 0) k = i1 + 10000 * i2; else if (count1 > 0) k = i1; else k = 0;
            s_design = fltrSig != k;
            if (fltrSig != k) fltrSig = k;
            s_i1 = i1; s_i2 = i2; s_c1 = count1; s_c2 = count2;
            ist[0] = old1; ist[1] = count1; ist[2] = old2; ist[3] = count2; ist[4] = fltrSig;
        }
        __syncthreads();
        if (s_design) {
            // half spectrum of the 512-point design: ones, zero bands around (i1 + 2) / 4 and (i2 + 2) / 4; the Nyquist bin is
            // whatever the last 2048-point filter spectrum left at index 256 (the reference reuses fltr_fft, quisk.c:913-936)
            const int k1 = (s_i1 + 2) / 4, k2 = (s_i2 + 2) / 4;
            const bool n1 = s_c1 > 0, n2 = s_c1 > 0 && s_c2 > 0;
            const double nyq = fl[256].x;
            for (int i = tid; i < 512; i += NT) {
                const int b = i <= 256 ? i : 512 - i;
                double v = 1.0;
                if (n1 && b >= k1 - p.half_width && b <= k1 + p.half_width) v = 0.0;
                if (n2 && b >= k2 - p.half_width && b <= k2 + p.half_width) v = 0.0;
                if (b == 256) v = nyq;
                s[fsw(i)] = make_double2(v, 0.0);
            }
            __syncthreads();
            fft_smem<1>(s, 512, twl512, +1, tid, NT);
            for (int i = tid; i < 512; i += NT) fo[i] = s[fsw(i)].x;
            __syncthreads();
            // centre, mirror (quisk.c:938-941), window and scale (quisk.c:942-943), zero-pad, forward transform
            for (int i = tid; i < QC_AN_N; i += NT) {
                double v = 0.0;
                if (i < 511) {
                    double f;
                    if (i >= 255) f = i <= 508 ? fo[i - 255] : fo[i];
                    else f = i >= 2 ? fo[255 - i] : fo[510 - i];
                    v = f * window[i] / 2048 / 4;
                }
                s[fsw(i)] = make_double2(v, 0.0);
            }
            __syncthreads();
            fft_smem<1>(s, QC_AN_N, twl2k, -1, tid, NT);
            for (int i = tid; i < QC_AN_BINS; i += NT) fl[i] = s[fsw(i)];
            __syncthreads();
        }
        // apply the filter, Hermitian extension, inverse transform
        for (int i = tid; i < QC_AN_BINS; i += NT) {
            const cd x = X[i], h = fl[i];
            cd y = make_double2(x.x * h.x - x.y * h.y, x.x * h.y + x.y * h.x);
            if (i == 0 || i == QC_AN_N / 2) y.y = 0.0;
            s[fsw(i)] = y;
            if (i > 0 && i < QC_AN_N / 2) s[fsw(QC_AN_N - i)] = make_double2(y.x, -y.y);
        }
        __syncthreads();
        fft_smem<1>(s, QC_AN_N, twl2k, +1, tid, NT);
        for (int i = QC_AN_START + tid; i < QC_AN_N; i += NT) dout[i] = s[fsw(i)].x / 102;       // NOTCH_DATA_SIZE / 20
        double keep[4];
        for (int e = 0, i = tid; e < 4; e++, i += NT) keep[e] = i < QC_AN_START ? din[QC_AN_OUT + i] : 0.0;
        __syncthreads();
        for (int e = 0, i = tid; e < 4; e++, i += NT) if (i < QC_AN_START) din[i] = keep[e];
        __syncthreads();
    }
}

struct QAgc { int C, rate, buf_size; AgcPar p; double *d_state; cd *d_fifo; };
struct QFrac { int C; double dindex; double *d_state; };
struct QBand { int S, n, L, count; const cd *tw; double *d_window, *d_avg, *d_max; };
struct QNb { int C, rate; NbPar p; size_t smem; double *d_state, *d_mag; cd *d_line; };

}  // namespace qc

using namespace qc;
struct qcAgc { QAgc a; };
struct qcFracDecim { QFrac f; };
struct qcBandscope { QBand b; };
struct qcNoiseBlanker { QNb b; };
struct qcAutoNotch { int C, rate, index; const cd *tw2k, *tw512; double *d_window, *d_in, *d_out, *d_avg; cd *d_fltr; int *d_ist; };
struct qcSsbSquelch { int C, rate, bw, index, didx, planned; const cd *tw; double *d_window, *d_infft, *d_delay; int *d_state; };

extern "C" {

qcAgc *quisk_cuda_agc_create(int n_channels, int sample_rate, double max_out, double release_gain, double release_time)
{
    if (ensure_device() != QC_OK) return nullptr;
    qcAgc *h = new qcAgc();
    QAgc &a = h->a;
    a.C = n_channels; a.rate = sample_rate;
    a.buf_size = sample_rate * 15 / 1000;                       // AGC_DELAY, quisk.c:47,2177
    a.p.buf_size = a.buf_size; a.p.is_cpx = 1; a.p.max_out = max_out; a.p.release_gain = release_gain;
    a.p.time_release = 1.0 - exp(-1.0 / sample_rate / release_time);    // quisk.c:2186
    if (n_channels <= 0 || a.buf_size <= 0 ||
        cudaMalloc((void **)&a.d_state, (size_t)n_channels * 8 * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&a.d_fifo, (size_t)n_channels * a.buf_size * sizeof(cd)) != cudaSuccess) {
        set_error("agc_create: bad arguments or allocation failure"); delete h; return nullptr;
    }
    std::vector<double> st((size_t)n_channels * 8, 0.0);
    for (int c = 0; c < n_channels; c++) { st[c * 8 + 3] = 1.0; st[c * 8 + 4] = 100.0; st[c * 8 + 6] = 100.0; }     // themax, gain, target_gain
    cudaMemcpy(a.d_state, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemset(a.d_fifo, 0, (size_t)n_channels * a.buf_size * sizeof(cd));
    return h;
}

void quisk_cuda_agc_destroy(qcAgc *h) { if (h) { cudaFree(h->a.d_state); cudaFree(h->a.d_fifo); delete h; } }

int quisk_cuda_agc_run(qcAgc *h, void *d_samples, long stride, int count, int is_cpx, void *stream)
{
    if (!h) return QC_EINVAL;
    if (count <= 0) return QC_OK;
    QAgc &a = h->a;
    AgcPar p = a.p; p.is_cpx = is_cpx;
    const size_t sh = (size_t)(count + a.buf_size) * sizeof(cd);
    if (sh > 200 * 1024) { set_error("agc_run: block too large"); return QC_EINVAL; }
    if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(agc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    agc_kernel<<<a.C, 64, sh, (cudaStream_t)stream>>>((cd *)d_samples, stride, count, a.C, a.d_state, a.d_fifo, p);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

qcFracDecim *quisk_cuda_fracdecim_create(int n_channels)
{
    if (ensure_device() != QC_OK) return nullptr;
    qcFracDecim *h = new qcFracDecim();
    h->f.C = n_channels; h->f.dindex = 1.0;                     // static double dindex = 1, quisk.c:627
    if (n_channels <= 0 || cudaMalloc((void **)&h->f.d_state, (size_t)n_channels * 8 * sizeof(double)) != cudaSuccess) { delete h; return nullptr; }
    std::vector<double> st((size_t)n_channels * 8, 0.0);
    for (int c = 0; c < n_channels; c++) st[c * 8] = 1.0;
    cudaMemcpy(h->f.d_state, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice);
    return h;
}

void quisk_cuda_fracdecim_destroy(qcFracDecim *h) { if (h) { cudaFree(h->f.d_state); delete h; } }

int quisk_cuda_fracdecim_run(qcFracDecim *h, const void *d_in, long in_stride, int count, double fdecim,
                             void *d_out, long out_stride, int *n_out, void *stream)
{
    if (!h) return QC_EINVAL;
    // the host walks the same index recurrence (same doubles, same order) to know the count
    double dindex = h->f.dindex; int nout = 0;
    for (int i = 0; i < count; i++) { if (dindex < 2) { nout++; dindex += fdecim - 1; } else dindex -= 1; }
    h->f.dindex = dindex;
    if (n_out) *n_out = nout;
    if (count <= 0) return QC_OK;
    const size_t sh = (size_t)count * sizeof(cd);
    if (sh > 200 * 1024) { set_error("fracdecim_run: block too large"); return QC_EINVAL; }
    if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(fracdecim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    fracdecim_kernel<<<h->f.C, 64, sh, (cudaStream_t)stream>>>((const cd *)d_in, in_stride, (cd *)d_out, out_stride, count, h->f.C, h->f.d_state, fdecim);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

qcNoiseBlanker *quisk_cuda_nb_create(int n_channels, int sample_rate)
{
    if (ensure_device() != QC_OK) return nullptr;
    qcNoiseBlanker *h = new qcNoiseBlanker();
    QNb &b = h->b;
    b.C = n_channels; b.rate = sample_rate;
    b.p.hwindow = (int)(sample_rate * 500.E-6 + 0.5);           // QUISK_NB_HWINDOW_SECS, quisk.c:679,704
    b.p.save_size = b.p.hwindow * 3;                            // quisk.c:705
    b.p.chunk = b.p.save_size < 512 ? b.p.save_size : 512;
    b.p.limit = 6.0;
    // delay line + magnitudes + one chunk of samples, sums and pulse flags
    b.smem = (size_t)b.p.save_size * (sizeof(cd) + sizeof(double)) + (size_t)b.p.chunk * (sizeof(cd) + sizeof(double) + 1) + 16;
    if (n_channels <= 0 || b.p.hwindow <= 0 || b.smem > 220 * 1024) {
        set_error("nb_create: need n_channels > 0 and a sample rate between 1 kS/s and 6 MS/s (the delay line of 1.5 ms lives in shared memory)");
        delete h; return nullptr;
    }
    if (cudaMalloc((void **)&b.d_state, (size_t)n_channels * 4 * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&b.d_mag, (size_t)n_channels * b.p.save_size * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&b.d_line, (size_t)n_channels * b.p.save_size * sizeof(cd)) != cudaSuccess) {
        set_error("nb_create: allocation failure"); delete h; return nullptr;
    }
    cudaMemset(b.d_state, 0, (size_t)n_channels * 4 * sizeof(double));          // state = index = win_index = 0, save_sum = 0.0 (quisk.c:700-703)
    cudaMemset(b.d_mag, 0, (size_t)n_channels * b.p.save_size * sizeof(double));
    cudaMemset(b.d_line, 0, (size_t)n_channels * b.p.save_size * sizeof(cd));
    return h;
}

void quisk_cuda_nb_destroy(qcNoiseBlanker *h) { if (h) { cudaFree(h->b.d_state); cudaFree(h->b.d_mag); cudaFree(h->b.d_line); delete h; } }

int quisk_cuda_nb_run(qcNoiseBlanker *h, void *d_samples, long stride, int count, int level, void *stream)
{
    if (!h) return QC_EINVAL;
    if (level <= 0 || count <= 0) return QC_OK;                 // quisk.c:695: off, nothing moves through the delay line
    QNb &b = h->b;
    NbPar p = b.p;
    p.limit = level == 2 ? 4.0 : (level == 3 ? 2.5 : 6.0);      // quisk.c:716-728
    static bool optin[64] = {};
    int dev = 0; cudaGetDevice(&dev); dev &= 63;
    if (!optin[dev]) { QC_CUDA(cudaFuncSetAttribute(nb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); optin[dev] = true; }
    nb_kernel<<<b.C, 128, b.smem, (cudaStream_t)stream>>>((cd *)d_samples, stride, count, b.d_state, b.d_line, b.d_mag, p);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

qcSsbSquelch *quisk_cuda_ssb_squelch_create(int n_channels, int samp_rate, int filter_bandwidth)
{
    if (ensure_device() != QC_OK) return nullptr;
    if (n_channels <= 0 || samp_rate <= 0) { set_error("ssb_squelch_create: bad arguments"); return nullptr; }
    qcSsbSquelch *h = new qcSsbSquelch();
    h->C = n_channels; h->rate = samp_rate; h->bw = filter_bandwidth; h->index = 0; h->didx = 0; h->planned = 0;
    h->tw = fft_twiddles(QC_SQ_N);
    std::vector<double> w(QC_SQ_N);
    for (int i = 0; i < QC_SQ_N; i++) w[i] = 0.50 - 0.50 * cos(2. * M_PI * i / QC_SQ_N);        // quisk.c:1110
    const size_t nb = (size_t)n_channels * QC_SQ_N * sizeof(double);
    if (!h->tw || cudaMalloc((void **)&h->d_window, QC_SQ_N * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&h->d_infft, nb) != cudaSuccess || cudaMalloc((void **)&h->d_delay, nb) != cudaSuccess ||
        cudaMalloc((void **)&h->d_state, (size_t)n_channels * 2 * sizeof(int)) != cudaSuccess) {
        set_error("ssb_squelch_create: allocation failure"); delete h; return nullptr;
    }
    cudaMemcpy(h->d_window, w.data(), QC_SQ_N * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemset(h->d_infft, 0, nb);
    cudaMemset(h->d_delay, 0, nb);
    cudaMemset(h->d_state, 0, (size_t)n_channels * 2 * sizeof(int));        // sq_open = 0 (quisk.c:1102), squelch_active = 0 (quisk.c:1908)
    return h;
}

void quisk_cuda_ssb_squelch_destroy(qcSsbSquelch *h)
{
    if (h) { cudaFree(h->d_window); cudaFree(h->d_infft); cudaFree(h->d_delay); cudaFree(h->d_state); delete h; }
}

int quisk_cuda_ssb_squelch_run(qcSsbSquelch *h, double *d_audio, long stride, int count, int level, void *stream)
{
    if (!h || count < 0) return QC_EINVAL;
    if (count > 8192) { set_error("ssb_squelch_run: at most 8192 samples per call (the timer is decremented once per call, so a call cannot be split)"); return QC_EINVAL; }
    // The reference's first call only creates its FFT plan and returns (quisk.c:1104-1112): those samples never reach
    // the frame and the timer is not touched; the delay line runs all the same (quisk.c:1926-1927).
    const int do_squelch = h->planned;
    h->planned = 1;
    if (count == 0 && !do_squelch) return QC_OK;
    SqPar p;
    p.samp_rate = h->rate;
    int bw = h->bw > 3000 ? 3000 : h->bw;                       // quisk.c:1120-1124
    p.bw1 = 300 * QC_SQ_N / h->rate;
    p.bw2 = (bw + 300) * QC_SQ_N / h->rate;
    if (p.bw2 > QC_SQ_N / 2 + 1) p.bw2 = QC_SQ_N / 2 + 1;       // out_fft holds N/2 + 1 bins
    if (p.bw1 > p.bw2) p.bw1 = p.bw2;
    p.thresh = level * 0.005;                                   // quisk.c:1160
    const size_t sh = (fft_tw_entries(QC_SQ_N) + QC_SQ_N) * sizeof(cd) + (size_t)(QC_SQ_N + count) * sizeof(double);
    static bool optin[64] = {};
    int dev = 0; cudaGetDevice(&dev); dev &= 63;
    if (!optin[dev]) { QC_CUDA(cudaFuncSetAttribute(ssb_squelch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); optin[dev] = true; }
    ssb_squelch_kernel<<<h->C, fft_threads(QC_SQ_N), sh, (cudaStream_t)stream>>>(d_audio, stride, count, h->index, do_squelch, h->didx, p,
                                                                                  h->tw, h->d_window, h->d_infft, h->d_delay, h->d_state);
    count_launch();
    QC_CUDA_LAUNCH();
    if (do_squelch) h->index = (h->index + count) % QC_SQ_N;
    h->didx = (h->didx + count) % QC_SQ_N;
    return QC_OK;
}

int quisk_cuda_ssb_squelch_state(qcSsbSquelch *h, int *sq_open, int *squelch_active, void *stream)
{
    if (!h) return QC_EINVAL;
    std::vector<int> st((size_t)h->C * 2);
    QC_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    QC_CUDA(cudaMemcpy(st.data(), h->d_state, st.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int c = 0; c < h->C; c++) { if (sq_open) sq_open[c] = st[c * 2]; if (squelch_active) squelch_active[c] = st[c * 2 + 1]; }
    return QC_OK;
}

const int *quisk_cuda_ssb_squelch_state_ptr(qcSsbSquelch *h) { return h ? h->d_state : nullptr; }

qcAutoNotch *quisk_cuda_autonotch_create(int n_channels, int rate)
{
    if (ensure_device() != QC_OK) return nullptr;
    if (n_channels <= 0 || rate <= 0) { set_error("autonotch_create: bad arguments"); return nullptr; }
    qcAutoNotch *h = new qcAutoNotch();
    h->C = n_channels; h->rate = rate; h->index = QC_AN_START;          // the state after the initialising call, quisk.c:826-835
    h->tw2k = fft_twiddles(QC_AN_N); h->tw512 = fft_twiddles(512);
    std::vector<double> w(QC_AN_N, 0.0);
    for (int i = 0; i < 511; i++) w[i] = 0.50 - 0.50 * cos(2. * M_PI * i / 511);        // Hann over NOTCH_FILTER_SIZE, quisk.c:822-823
    const size_t C_ = (size_t)n_channels;
    if (!h->tw2k || !h->tw512 || cudaMalloc((void **)&h->d_window, QC_AN_N * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&h->d_in, C_ * QC_AN_N * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&h->d_out, C_ * QC_AN_N * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&h->d_avg, C_ * QC_AN_BINS * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&h->d_fltr, C_ * QC_AN_BINS * sizeof(cd)) != cudaSuccess ||
        cudaMalloc((void **)&h->d_ist, C_ * 8 * sizeof(int)) != cudaSuccess) {
        set_error("autonotch_create: allocation failure"); delete h; return nullptr;
    }
    cudaMemcpy(h->d_window, w.data(), QC_AN_N * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemset(h->d_in, 0, C_ * QC_AN_N * sizeof(double));
    cudaMemset(h->d_out, 0, C_ * QC_AN_N * sizeof(double));
    cudaMemset(h->d_avg, 0, C_ * QC_AN_BINS * sizeof(double));
    cudaMemset(h->d_fltr, 0, C_ * QC_AN_BINS * sizeof(cd));
    std::vector<int> ist(C_ * 8, 0);
    for (size_t c = 0; c < C_; c++) { ist[c * 8 + 1] = -4; ist[c * 8 + 3] = -4; ist[c * 8 + 4] = -1; }      // count1, count2, fltrSig
    cudaMemcpy(h->d_ist, ist.data(), ist.size() * sizeof(int), cudaMemcpyHostToDevice);
    return h;
}

void quisk_cuda_autonotch_destroy(qcAutoNotch *h)
{
    if (h) { cudaFree(h->d_window); cudaFree(h->d_in); cudaFree(h->d_out); cudaFree(h->d_avg); cudaFree(h->d_fltr); cudaFree(h->d_ist); delete h; }
}

int quisk_cuda_autonotch_run(qcAutoNotch *h, double *d_audio, long stride, int count, int sidetone, void *stream)
{
    if (!h || count < 0) return QC_EINVAL;
    if (count == 0) return QC_OK;
    const int rate = h->rate, NF = QC_AN_BINS, NFF = 256;       // NOTCH_FFT_SIZE, NOTCH_FILTER_FFT_SIZE
    AnPar p;
    p.delta_sig = (300 * 2 * NF + rate / 2) / rate;             // quisk.c:842-848
    p.delta_i1 = (400 * 2 * NF + rate / 2) / rate;
    p.signal = sidetone != 0 ? (abs(sidetone) * 2 * NF + rate / 2) / rate : -999;
    p.half_width = (100 * 2 * NFF + rate / 2) / rate;           // quisk.c:905-907
    if (p.half_width < 3) p.half_width = 3;
    const size_t sh = (size_t)(fft_tw_entries(QC_AN_N) + fft_tw_entries(512) + QC_AN_N + QC_AN_BINS + 1) * sizeof(cd) + 512 * sizeof(double);
    static bool optin[64] = {};
    int dev = 0; cudaGetDevice(&dev); dev &= 63;
    if (!optin[dev]) { QC_CUDA(cudaFuncSetAttribute(autonotch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); optin[dev] = true; }
    autonotch_kernel<<<h->C, 128, sh, (cudaStream_t)stream>>>(d_audio, stride, count, h->index, p, h->tw2k, h->tw512, h->d_window,
                                                              h->d_in, h->d_out, h->d_avg, h->d_fltr, h->d_ist);
    count_launch();
    QC_CUDA_LAUNCH();
    // the frame index advances the same way for every channel: 510 -> 2048 wraps to 510
    long idx = h->index + (long)count;
    while (idx >= QC_AN_N) idx -= QC_AN_OUT;
    h->index = (int)idx;
    return QC_OK;
}

qcBandscope *quisk_cuda_bandscope_create(int n_streams, int size)
{
    if (ensure_device() != QC_OK) return nullptr;
    if (n_streams <= 0 || fft_log2(size) < 0) { set_error("bandscope_create: size must be a power of two in [8, 8192]"); return nullptr; }
    qcBandscope *h = new qcBandscope();
    QBand &b = h->b;
    b.S = n_streams; b.n = size; b.L = size / 2 + 1; b.count = 0; b.tw = fft_twiddles(size);
    std::vector<double> w((size_t)size);
    for (int i = 0, j = -size / 2; i < size; i++, j++) w[i] = 0.5 + 0.5 * cos(2. * M_PI * j / size);       // quisk.c:2887-2888
    if (!b.tw || cudaMalloc((void **)&b.d_window, (size_t)size * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&b.d_avg, (size_t)n_streams * (b.L + 1) * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void **)&b.d_max, (size_t)n_streams * sizeof(double)) != cudaSuccess) { delete h; return nullptr; }
    cudaMemcpy(b.d_window, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemset(b.d_avg, 0, (size_t)n_streams * (b.L + 1) * sizeof(double));
    cudaMemset(b.d_max, 0, (size_t)n_streams * sizeof(double));
    return h;
}

void quisk_cuda_bandscope_destroy(qcBandscope *h) { if (h) { cudaFree(h->b.d_window); cudaFree(h->b.d_avg); cudaFree(h->b.d_max); delete h; } }

int quisk_cuda_bandscope_accumulate(qcBandscope *h, const double *d_blocks, long stream_stride, int n_blocks, void *stream)
{
    if (!h) return QC_EINVAL;
    if (n_blocks <= 0) return QC_OK;
    QBand &b = h->b;
    const size_t sh = ((size_t)b.n + fft_tw_entries(b.n)) * sizeof(cd);
    if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(bandscope_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    bandscope_accumulate_kernel<<<b.S, fft_threads(b.n), sh, (cudaStream_t)stream>>>(d_blocks, stream_stride, n_blocks, b.n, b.tw, b.d_window, b.d_avg, b.d_max);
    count_launch();
    QC_CUDA_LAUNCH();
    b.count += n_blocks;
    return QC_OK;
}

int quisk_cuda_bandscope_graph(qcBandscope *h, int graph_width, int clock, double zoom, double deltaf, double *d_graph, void *stream)
{
    if (!h) return QC_EINVAL;
    QBand &b = h->b;
    if (b.count <= 0 || graph_width <= 0) { set_error("bandscope_graph: nothing accumulated"); return QC_EINVAL; }
    const double frac = (double)b.L / graph_width;
    const double scale = 1.0 / frac / b.count / b.n;             // quisk.c:4989
    const double rate = clock / 2.0;
    cudaStream_t s = (cudaStream_t)stream;
    bandscope_graph_kernel<<<dim3((graph_width + 127) / 128, b.S), 128, 0, s>>>(b.d_avg, b.L, graph_width, zoom, deltaf, rate, scale, d_graph);
    count_launch();
    QC_CUDA_LAUNCH();
    QC_CUDA(cudaMemsetAsync(b.d_avg, 0, (size_t)b.S * (b.L + 1) * sizeof(double), s));
    QC_CUDA(cudaMemsetAsync(b.d_max, 0, (size_t)b.S * sizeof(double), s));
    b.count = 0;
    return QC_OK;
}

}  // extern "C"
