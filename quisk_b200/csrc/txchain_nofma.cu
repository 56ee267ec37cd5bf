// quisk_b200/csrc/txchain_nofma.cu -- the transmit-audio chain of the reference's microphone.c (tx_filter,
// microphone.c:372-604, with its peak rounder CcmPeak, :161-233), batched over C independent transmitters: the TX
// mirror of the receive path, built from the same primitives.
//
//   mic audio, 48 or 8 kS/s, real part of the input, +-CLIP16
//     -> / CLIP16 -> quisk_dDecimate(quiskLpFilt48Coefs, /6) -> quisk_dFilter(quiskFiltTx8kAudioB)     [exact polyfir kernels]
//     -> pre-emphasis  y = 2 (x - p x_1)                                                              [parallel, one carry]
//     SSB: -> quisk_dC_out(quiskMicFilt8Coefs tuned to +-1650 Hz) * 2                                  [QC_C_CDECIMATE, decim 1]
//          -> running peak normaliser (inMax), clip gain, hard limit to |z| <= 1, real part            [sequential lane per channel]
//          -> quisk_dFilter -> quisk_dC_out * 2 -> CcmPeak (30 ms look-ahead peak rounder)             [sequential lane]
//          -> quisk_cDecimate(.., 1) -> quisk_cInterpolate(quiskLpFilt48Coefs, x6) -> * CLIP16
//     AM / FM: -> inMax normaliser, clip gain, quadratic soft knee -> quisk_dFilter -> CcmPeak (real)
//          -> quisk_dFilter -> quisk_dInterpolate(x6) -> * CLIP16 on the real rail
//
// The FIR stages are the library's exact BatchFilter kinds (bit-exact against filter.c); the two recurrences are one
// thread per transmitter walking the block at 8 kS/s -- a sixth of the input rate, a few hundred samples per call: they
// are scalar by nature (each sample's gain depends on the one before) and carry no weight next to the FIRs.
// Compiled with --fmad=false (file name): a*b+c rounds twice, as the reference's x86-64 build does.
// Quirk kept: CcmPeak's first call only initialises its state and returns (microphone.c:174-191), so the first block of a
// stream passes the peak rounder untouched and undelayed.
#include "batch.h"
#include "../../include/quisk_cuda.h"
#include <complex>

namespace qc {

static constexpr double TX_CLIP16 = 32767.0;
static constexpr int CCM_N = 8000 * 30 / 1000;          // CcmPeak's delay line, microphone.c:175

struct TxLevelPar {
    double time_long, time_short, agc_level, clip, Xmin, Xmax, Ymax, aaa, bbb, ccc;
    double out_short, out_long;
};

__global__ void tx_in_kernel(const cd *__restrict__ in, long is, double *__restrict__ out, long os, int n)
{
    const int c = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[(size_t)c * os + i] = in[(size_t)c * is + i].x / TX_CLIP16;            // microphone.c:437-438
}

// y[i] = (x[i] - p x[i-1]) then * 2 (microphone.c:455-459); x_1 carried per transmitter
__global__ void tx_preemph_kernel(const double *__restrict__ in, double *__restrict__ out, long st, int n, double p, double *__restrict__ x1)
{
    const int c = blockIdx.y;
    const double *x = in + (size_t)c * st;
    double *y = out + (size_t)c * st;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double prev = i ? x[i - 1] : x1[c];
        double v = x[i] - p * prev;
        v *= 2.0;
        y[i] = v;
    }
}
__global__ void tx_carry_kernel(const double *__restrict__ in, long st, int n, double *__restrict__ x1, int C)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C && n > 0) x1[c] = in[(size_t)c * st + n - 1];
}

__global__ void tx_real_rail_kernel(const cd *__restrict__ in, long is, cd *__restrict__ out, long os, int n)
{
    const int c = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[(size_t)c * os + i] = make_double2(in[(size_t)c * is + i].x, 0.0);          // creal(filtered[i]), microphone.c:620
}

__global__ void tx_promote_kernel(const double *__restrict__ in, long is, cd *__restrict__ out, long os, int n)
{
    const int c = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[(size_t)c * os + i] = make_double2(in[(size_t)c * is + i], 0.0);
}

__device__ __forceinline__ double tx_inmax(double inMax, double magn, const TxLevelPar &P)
{   // microphone.c:476-481 / 504-509
    if (magn > inMax) return inMax * (1 - P.time_short) + P.time_short * magn;
    if (magn > P.agc_level) return inMax * (1 - P.time_long) + P.time_long * magn;
    return inMax * (1 - P.time_long) + P.time_long * P.agc_level;
}

// SSB: csample = 2 z; normalise by the running peak, clip gain, limit to the unit circle, keep the real part (microphone.c:470-496).
// A lane per transmitter, 32 transmitters per warp: the recurrence is scalar per transmitter, so the lanes are the batch.  The
// block travels through a [32 transmitters][32 samples] shared-memory tile (rows padded by one element), loaded and stored
// row by row so that HBM sees 512-byte / 256-byte runs instead of 32 strided words per instruction.
__global__ void __launch_bounds__(32) tx_level_ssb_kernel(const cd *__restrict__ in, long is, double *__restrict__ out, long os, int n, double *__restrict__ inmax, int C, TxLevelPar P)
{
    __shared__ cd tin[32][33];
    __shared__ double tout[32][33];
    const int lane = threadIdx.x;
    const int c0 = blockIdx.x * 32, c = c0 + lane;
    const int rows = C - c0 < 32 ? C - c0 : 32;
    double im = c < C ? inmax[c] : 1.0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int m = n - i0 < 32 ? n - i0 : 32;
        for (int r = 0; r < rows; r++) if (lane < m) tin[r][lane] = in[(size_t)(c0 + r) * is + i0 + lane];
        __syncwarp();
        if (c < C) {
            for (int k = 0; k < m; k++) {
                const cd x = tin[lane][k];
                cd z = make_double2(x.x * 2.0, x.y * 2.0);
                double magn = hypot(z.x, z.y);
                im = tx_inmax(im, magn, P);
                z.x /= im; z.y /= im; magn /= im;
                z.x *= P.clip; z.y *= P.clip; magn *= P.clip;
                if (magn > 1.0) { z.x /= magn; z.y /= magn; }
                tout[lane][k] = z.x;
            }
        }
        __syncwarp();
        for (int r = 0; r < rows; r++) if (lane < m) out[(size_t)(c0 + r) * os + i0 + lane] = tout[r][lane];
        __syncwarp();
    }
    if (c < C) inmax[c] = im;
}

// AM / FM: the same normaliser on the real rail, then the quadratic soft knee (microphone.c:499-527)
__global__ void tx_level_real_kernel(double *data, long st, int n, double *__restrict__ inmax, int C, TxLevelPar P)
{   // in place
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double im = inmax[c];
    double *x = data + (size_t)c * st;
    double *y = x;
    for (int i = 0; i < n; i++) {
        double d = x[i];
        double magn = fabs(d);
        im = tx_inmax(im, magn, P);
        d /= im; magn /= im;
        d *= P.clip; magn *= P.clip;
        if (magn < P.Xmin) y[i] = d;
        else if (magn > P.Xmax) y[i] = copysign(P.Ymax, d);
        else y[i] = copysign(P.aaa * magn * magn + P.bbb * magn + P.ccc, d);
    }
    inmax[c] = im;
}

// CcmPeak (microphone.c:192-232): 240-sample delay, output divided by a level that rises fast to the largest magnitude
// in the delay line and falls slowly back to one.  state per transmitter: [0] themax, [1] level, [2] index_read, [3] pad,
// then levl[240], then the delayed samples (240 complex or 240 real).
// One WARP per transmitter: the delay line and the magnitudes live in shared memory for the length of the call, the
// block's samples pass through a 32-sample staging row (coalesced both ways), and all 32 lanes run the recurrence in
// lockstep on the same values (uniform control flow, broadcast reads) so that the one step that is not scalar -- the search
// for the new maximum when the old one leaves the delay line, `for (j = 0; j < 240; j++)` -- is eight entries per lane and
// five shuffles instead of 240 dependent reads.  (The first version walked global memory with one thread per transmitter:
// 6.5 ms of a 14 ms step at 4096 transmitters; this one: see DESIGN.md 4.12.)
static constexpr int CCM_WPB = 4;                                   // warps (transmitters) per CTA
static constexpr int CCM_WS = 3 * CCM_N + 64 + 64;                  // doubles of shared memory per warp: levl, delay line, staging in / out
template <bool CPX>
__global__ void __launch_bounds__(32 * CCM_WPB) tx_ccm_kernel(void *__restrict__ data, long st, int n, double *__restrict__ state, int C, TxLevelPar P, double in_scale)
{
    __shared__ __align__(16) double sm[CCM_WPB * CCM_WS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * CCM_WPB + w;
    if (c >= C) return;
    constexpr int SW = 4 + CCM_N + 2 * CCM_N;          // header padded to four doubles: the complex delay line stays 16-byte aligned
    double *s = state + (size_t)c * SW;
    double *levl = sm + w * CCM_WS, *buf = levl + CCM_N, *sin = buf + 2 * CCM_N, *sout = sin + 64;
    for (int j = lane; j < 3 * CCM_N; j += 32) levl[j] = s[4 + j];
    double themax = s[0], level = s[1];
    int idx = (int)s[2];
    __syncwarp();
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int m = n - i0 < 32 ? n - i0 : 32;
        if (lane < m) {
            if (CPX) reinterpret_cast<cd *>(sin)[lane] = (reinterpret_cast<cd *>(data) + (size_t)c * st)[i0 + lane];
            else sin[lane] = (reinterpret_cast<double *>(data) + (size_t)c * st)[i0 + lane];
        }
        __syncwarp();
        for (int k = 0; k < m; k++) {
            double newlevel;
            if (CPX) {
                const cd x = reinterpret_cast<cd *>(sin)[k];
                const cd v = make_double2(x.x * in_scale, x.y * in_scale);         // quisk_dC_out(..) * 2.0, microphone.c:531
                const cd b = reinterpret_cast<cd *>(buf)[idx];
                __syncwarp();
                if (lane == 0) { reinterpret_cast<cd *>(sout)[k] = make_double2(b.x / level, b.y / level); reinterpret_cast<cd *>(buf)[idx] = v; }
                newlevel = hypot(v.x, v.y);
            } else {
                const double v = sin[k];
                const double b = buf[idx];
                __syncwarp();
                if (lane == 0) { sout[k] = b / level; buf[idx] = v; }
                newlevel = fabs(v);
            }
            const double oldlevel = levl[idx];
            __syncwarp();
            if (lane == 0) levl[idx] = newlevel;
            __syncwarp();
            if (newlevel < themax && oldlevel < themax) {
            } else if (newlevel > themax && newlevel > oldlevel) {
                themax = newlevel;
            } else {
                double mx = 0;
                for (int j = lane; j < CCM_N; j += 32) { const double v = levl[j]; if (v > mx) mx = v; }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) { const double o = __shfl_xor_sync(0xffffffffu, mx, off); if (o > mx) mx = o; }
                themax = mx;
            }
            if (themax > 1.0) level = level * (1.0 - P.out_short) + themax * P.out_short;
            else level = level * (1.0 - P.out_long) + 1.0 * P.out_long;
            if (++idx >= CCM_N) idx = 0;
        }
        __syncwarp();
        if (lane < m) {
            if (CPX) (reinterpret_cast<cd *>(data) + (size_t)c * st)[i0 + lane] = reinterpret_cast<cd *>(sout)[lane];
            else (reinterpret_cast<double *>(data) + (size_t)c * st)[i0 + lane] = sout[lane];
        }
        __syncwarp();
    }
    for (int j = lane; j < 3 * CCM_N; j += 32) s[4 + j] = levl[j];
    if (lane == 0) { s[0] = themax; s[1] = level; s[2] = (double)idx; }
}

// process_alc (microphone.c:270-370): the transmit level control behind tx_filter (quisk_process_microphone, :1232-1233).  A
// 960-sample (20 ms) delay line; every new sample that would clip at the gain the ramp is heading for re-aims the ramp so
// that the gain is right when that sample leaves the line; once per trip round the line the ramp is re-aimed upward (at most
// a doubling in five seconds) from the loudest sample seen.  A scalar state machine: one warp per transmitter keeps the
// delay line in shared memory and stages the block 32 samples at a time, lane 0 walks it.
// state per transmitter: gain_now, gain_change, next_change, final_gain, index, block_index, counter, fault, then 960 complex.
static constexpr int ALC_N = 960, ALC_WPB = 2, ALC_SW = 8 + 2 * ALC_N;
__global__ void __launch_bounds__(32 * ALC_WPB) tx_alc_kernel(cd *__restrict__ data, long st, int n, double *__restrict__ state, int C)
{
    __shared__ __align__(16) double sm[ALC_WPB * (2 * ALC_N + 128)];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * ALC_WPB + w;
    if (c >= C) return;
    double *s = state + (size_t)c * ALC_SW;
    cd *buf = reinterpret_cast<cd *>(sm + w * (2 * ALC_N + 128));
    cd *sin = buf + ALC_N, *sout = sin + 32;
    for (int j = lane; j < ALC_N; j += 32) buf[j] = reinterpret_cast<const cd *>(s + 8)[j];
    double gain_now = s[0], gain_change = s[1], next_change = s[2], final_gain = s[3];
    int index = (int)s[4], block_index = (int)s[5], counter = (int)s[6], fault = (int)s[7];
    const double gain_max = 3.0, gain_min = 0.1, top = (double)(32767 - 10);
    cd *x = data + (size_t)c * st;
    __syncwarp();
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int m = n - i0 < 32 ? n - i0 : 32;
        if (lane < m) sin[lane] = x[i0 + lane];
        __syncwarp();
        if (lane == 0) {
            for (int k = 0; k < m; k++) {
                const cd csamp = sin[k];
                const cd b = buf[index];
                sout[k] = make_double2(b.x * gain_now, b.y * gain_now);
                buf[index] = csamp;
                const double magn = hypot(csamp.x, csamp.y);
                if (magn * (gain_now + gain_change * ALC_N) > top) {
                    gain_change = (top / magn - gain_now) / ALC_N;
                    final_gain = gain_now + gain_change * ALC_N;
                    if (final_gain > gain_max) { final_gain = gain_max; gain_change = (final_gain - gain_now) / ALC_N; }
                    else if (final_gain < gain_min) { final_gain = gain_min; gain_change = (final_gain - gain_now) / ALC_N; }
                    block_index = index; counter = 0; fault = 0; next_change = 1E10;
                } else if (index == block_index) {
                    double d = 5.0;
                    d = 1.0 / (48000.0 * d);
                    if (next_change > d) next_change = d;
                    if (next_change != 1E10 && fault < ALC_N - 10) gain_change = next_change;
                    final_gain = gain_now + gain_change * ALC_N;
                    if (final_gain > gain_max) { final_gain = gain_max; gain_change = (final_gain - gain_now) / ALC_N; }
                    else if (final_gain < gain_min) { final_gain = gain_min; gain_change = (final_gain - gain_now) / ALC_N; }
                    fault = 0; counter = 0; next_change = 1E10;
                } else {
                    if (magn < 100) fault++;
                    else { const double d = (top / magn - final_gain) / ++counter; if (next_change > d) next_change = d; }
                }
                gain_now += gain_change;
                if (++index >= ALC_N) index = 0;
            }
        }
        __syncwarp();
        if (lane < m) x[i0 + lane] = sout[lane];
        __syncwarp();
    }
    for (int j = lane; j < ALC_N; j += 32) reinterpret_cast<cd *>(s + 8)[j] = buf[j];
    if (lane == 0) { s[0] = gain_now; s[1] = gain_change; s[2] = next_change; s[3] = final_gain; s[4] = index; s[5] = block_index; s[6] = counter; s[7] = fault; }
}

__global__ void tx_scale_c_kernel(cd *__restrict__ x, long st, int n, double g)
{
    const int c = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        cd v = x[(size_t)c * st + i];
        x[(size_t)c * st + i] = make_double2(v.x * g, v.y * g);
    }
}

__global__ void tx_out_real_kernel(const double *__restrict__ in, long is, cd *__restrict__ out, long os, int n)
{
    const int c = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[(size_t)c * os + i] = make_double2(in[(size_t)c * is + i] * TX_CLIP16, 0.0);       // microphone.c:577-578
}

struct TxFilter {
    int C = 0, mode = 0, mic_rate = 48000, decim = 6;
    bool ssb = false, ccm_started = false, digital = false;
    double preemph = 0.0;
    TxLevelPar P;
    BatchFilter *fDecim = nullptr, *fAudio1 = nullptr, *fAudio2 = nullptr, *fAudio3 = nullptr, *fInterp = nullptr, *fTune1 = nullptr, *fTune2 = nullptr;
    double *d_x1 = nullptr, *d_inmax = nullptr, *d_ccm = nullptr, *d_alc = nullptr, *d_r[2] = {nullptr, nullptr};
    bool alc_on = false;
    cd *d_c[2] = {nullptr, nullptr};
    long cap = 0;

    static BatchFilter *mk(int kind, int C, const double *coefs, int n, int interp, int decim)
    {
        BatchFilter *f = new BatchFilter();
        if (f->init(kind, C, coefs, n, interp, decim) != QC_OK) { f->release(); delete f; return nullptr; }
        return f;
    }

    int reset_state()
    {
        std::vector<double> h((size_t)C, 0.3);             // inMax, microphone.c:380
        QC_CUDA(cudaMemcpy(d_inmax, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
        QC_CUDA(cudaMemset(d_x1, 0, (size_t)C * sizeof(double)));
        const int SW = 4 + 3 * CCM_N;
        std::vector<double> st((size_t)C * SW, 0.0);
        for (int c = 0; c < C; c++) {
            double *s = st.data() + (size_t)c * SW;
            s[0] = 1.0; s[1] = 1.0; s[2] = 0.0;             // themax, level, index_read (microphone.c:176-178)
            for (int j = 0; j < CCM_N; j++) s[4 + j] = 1.0; // levl[] = 1 (microphone.c:185)
        }
        QC_CUDA(cudaMemcpy(d_ccm, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice));
        ccm_started = false;
        return QC_OK;
    }

    int init(int C_, int mode_, int mic_rate_, double preemph_, double clip, const qcTxTables &T)
    {
        C = C_; mode = mode_; mic_rate = mic_rate_; preemph = preemph_;
        if (C <= 0 || (mic_rate != 8000 && mic_rate != 48000)) { set_error("tx_filter_create: the microphone rate must be 8000 or 48000 (microphone.c:373)"); return QC_EINVAL; }
        digital = mode == QC_MODE_DGT_U || mode == QC_MODE_DGT_L || mode == QC_MODE_FDV_U || mode == QC_MODE_FDV_L;
        if (digital) {
            // tx_filter_digital (microphone.c:605-624): ONE filter, quiskDgtFilt48Coefs tuned to +-1650 Hz at 48 kS/s, times two
            if (mic_rate != 48000) { set_error("tx_filter_create: the digital modes run at 48000 samples per second (microphone.c:617)"); return QC_EINVAL; }
            if (!T.dgt_filt48 || T.n_dgt_filt48 <= 0) { set_error("tx_filter_create: quiskDgtFilt48Coefs is needed for the digital modes"); return QC_EINVAL; }
            decim = 1;
            struct quisk_dFilter tf;
            memset(&tf, 0, sizeof(tf));
            std::vector<double> taps(T.dgt_filt48, T.dgt_filt48 + T.n_dgt_filt48);
            quisk_filt_dInit(&tf, taps.data(), T.n_dgt_filt48);
            quisk_filt_tune(&tf, 1650.0 / 48000, mode != QC_MODE_DGT_L && mode != QC_MODE_FDV_L);
            fTune1 = mk(QC_C_CDECIMATE, C, (const double *)tf.cpxCoefs, T.n_dgt_filt48, 1, 1);
            free(tf.cpxCoefs); free(tf.dSamples);
            return fTune1 ? QC_OK : QC_EINVAL;
        }
        if (mode != QC_MODE_LSB && mode != QC_MODE_USB && mode != QC_MODE_AM && mode != QC_MODE_FM) { set_error("tx_filter_create: tx_filter serves LSB, USB, AM and FM, tx_filter_digital DGT-U/L and FDV-U/L"); return QC_EINVAL; }
        if (!T.mic_filt8 || !T.lp_filt48 || !T.tx8k_audio) { set_error("tx_filter_create: quiskMicFilt8Coefs, quiskLpFilt48Coefs and quiskFiltTx8kAudioB are needed"); return QC_EINVAL; }
        ssb = mode == QC_MODE_LSB || mode == QC_MODE_USB;
        decim = mic_rate / 8000;
        const double dt = 1.0 / 8000;
        P.time_long = 1.0 - exp(-dt / 3.000); P.time_short = 1.0 - exp(-dt / 0.005);        // microphone.c:405-407
        P.Ymax = pow(10.0, -1 / 20.0); P.Xmax = pow(10.0, 3 / 20.0);
        P.Xmin = P.Ymax - fabs(P.Ymax - P.Xmax);
        P.aaa = 1.0 / (2.0 * (P.Xmin - P.Xmax)); P.bbb = -2.0 * P.aaa * P.Xmax; P.ccc = P.Ymax - P.aaa * P.Xmax * P.Xmax - P.bbb * P.Xmax;
        P.agc_level = 0.10; P.clip = clip;
        P.out_short = 1.0 - exp(-dt / 0.010); P.out_long = 1.0 - exp(-dt / 3.000);          // microphone.c:187-189
        if (decim > 1 && !(fDecim = mk(QC_D_DECIMATE, C, T.lp_filt48, T.n_lp_filt48, 1, decim))) return QC_EINVAL;
        if (!(fAudio1 = mk(QC_D_DECIMATE, C, T.tx8k_audio, T.n_tx8k_audio, 1, 1))) return QC_EINVAL;
        if (!(fAudio2 = mk(QC_D_DECIMATE, C, T.tx8k_audio, T.n_tx8k_audio, 1, 1))) return QC_EINVAL;
        if (ssb) {
            // quisk_filt_tune(&filter, 1650 / 8000, rxMode != LSB), filter.c:58-81: through the library's own filter.h entry
            struct quisk_dFilter tf;
            memset(&tf, 0, sizeof(tf));
            std::vector<double> taps(T.mic_filt8, T.mic_filt8 + T.n_mic_filt8);
            quisk_filt_dInit(&tf, taps.data(), T.n_mic_filt8);
            quisk_filt_tune(&tf, 1650.0 / 8000, mode != QC_MODE_LSB);
            fTune1 = mk(QC_C_CDECIMATE, C, (const double *)tf.cpxCoefs, T.n_mic_filt8, 1, 1);
            fTune2 = mk(QC_C_CDECIMATE, C, (const double *)tf.cpxCoefs, T.n_mic_filt8, 1, 1);
            free(tf.cpxCoefs); free(tf.dSamples);
            if (!fTune1 || !fTune2) return QC_EINVAL;
            if (!(fAudio3 = mk(QC_C_DECIMATE, C, T.tx8k_audio, T.n_tx8k_audio, 1, 1))) return QC_EINVAL;
            if (decim > 1 && !(fInterp = mk(QC_C_INTERPOLATE, C, T.lp_filt48, T.n_lp_filt48, decim, 1))) return QC_EINVAL;
        } else {
            if (!(fAudio3 = mk(QC_D_DECIMATE, C, T.tx8k_audio, T.n_tx8k_audio, 1, 1))) return QC_EINVAL;
            if (decim > 1 && !(fInterp = mk(QC_D_INTERPOLATE, C, T.lp_filt48, T.n_lp_filt48, decim, 1))) return QC_EINVAL;
        }
        QC_CUDA(cudaMalloc((void **)&d_x1, (size_t)C * sizeof(double)));
        QC_CUDA(cudaMalloc((void **)&d_inmax, (size_t)C * sizeof(double)));
        QC_CUDA(cudaMalloc((void **)&d_ccm, (size_t)C * (4 + 3 * CCM_N) * sizeof(double)));
        return reset_state();
    }

    void release()
    {
        for (BatchFilter *f : {fDecim, fAudio1, fAudio2, fAudio3, fInterp, fTune1, fTune2}) if (f) { f->release(); delete f; }
        fDecim = fAudio1 = fAudio2 = fAudio3 = fInterp = fTune1 = fTune2 = nullptr;
        for (void *p : {(void *)d_x1, (void *)d_inmax, (void *)d_ccm, (void *)d_alc, (void *)d_r[0], (void *)d_r[1], (void *)d_c[0], (void *)d_c[1]}) if (p) cudaFree(p);
        d_x1 = d_inmax = d_ccm = d_alc = d_r[0] = d_r[1] = nullptr; d_c[0] = d_c[1] = nullptr;
    }

    int set_alc(int enable)
    {   // enable: init_alc(&tx_alc, 960) the first time (gain 1.4 in the digital modes, 1.0 otherwise, microphone.c:242-254), and
        // init_alc(&tx_alc, 0) every time (what quisk_process_microphone does on key down, :1207): line and ramp cleared, gain kept
        alc_on = enable != 0;
        if (!alc_on) return QC_OK;
        std::vector<double> st((size_t)C * ALC_SW, 0.0);
        if (d_alc) {
            std::vector<double> old((size_t)C * ALC_SW);
            QC_CUDA(cudaMemcpy(old.data(), d_alc, old.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (int c = 0; c < C; c++) st[(size_t)c * ALC_SW] = old[(size_t)c * ALC_SW];
        } else {
            QC_CUDA(cudaMalloc((void **)&d_alc, st.size() * sizeof(double)));
            for (int c = 0; c < C; c++) st[(size_t)c * ALC_SW] = digital ? 1.4 : 1.0;
        }
        QC_CUDA(cudaMemcpy(d_alc, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice));
        return QC_OK;
    }

    int run_alc(cd *d_out, long os, int n, cudaStream_t s)
    {
        if (!alc_on || n <= 0) return QC_OK;
        tx_alc_kernel<<<(C + ALC_WPB - 1) / ALC_WPB, 32 * ALC_WPB, 0, s>>>(d_out, os, n, d_alc, C); count_launch(); QC_CUDA_LAUNCH();
        return QC_OK;
    }

    int reserve(int count)
    {
        const long need = (long)count + 64;
        if (need <= cap) return QC_OK;
        for (int i = 0; i < 2; i++) { if (d_r[i]) cudaFree(d_r[i]); if (d_c[i]) cudaFree(d_c[i]); d_r[i] = nullptr; d_c[i] = nullptr; }
        for (int i = 0; i < 2; i++) {
            QC_CUDA(cudaMalloc((void **)&d_r[i], (size_t)C * need * sizeof(double)));
            QC_CUDA(cudaMalloc((void **)&d_c[i], (size_t)C * need * sizeof(cd)));
        }
        cap = need;
        return QC_OK;
    }

    int max_out(int count) const { return digital ? count : (count / decim + 1) * decim; }

    int process(const cd *d_in, long is, int count, cd *d_out, long os, int *n_out, cudaStream_t s)
    {
        if (count < 0 || !d_in || !d_out) { set_error("tx_filter_process: bad arguments"); return QC_EINVAL; }
        if (n_out) *n_out = 0;
        if (count == 0) return QC_OK;
        int rc = reserve(count); if (rc != QC_OK) return rc;
        const dim3 g((unsigned)((count + 255) / 256 < 64 ? (count + 255) / 256 : 64), (unsigned)C);
        const int gc = (C + 63) / 64;
        int n = count, no = 0;
        if (digital) {
            // filtered[i] = quisk_dC_out(creal(filtered[i]), &filter1) * 2.00: the real rail promoted, the tuned taps, times two
            tx_real_rail_kernel<<<g, 256, 0, s>>>(d_in, is, d_c[0], cap, n); count_launch(); QC_CUDA_LAUNCH();
            rc = fTune1->run(d_c[0], cap, n, d_out, os, &no, 0, s); if (rc != QC_OK) return rc;
            tx_scale_c_kernel<<<g, 256, 0, s>>>(d_out, os, n, 2.0); count_launch(); QC_CUDA_LAUNCH();
            rc = run_alc(d_out, os, n, s); if (rc != QC_OK) return rc;
            if (n_out) *n_out = n;
            return QC_OK;
        }
        tx_in_kernel<<<g, 256, 0, s>>>(d_in, is, d_r[0], cap, n); count_launch(); QC_CUDA_LAUNCH();
        int cur = 0;
        if (fDecim) { rc = fDecim->run(d_r[0], cap, n, d_r[1], cap, &no, 0, s); if (rc != QC_OK) return rc; n = no; cur = 1; }
        rc = fAudio1->run(d_r[cur], cap, n, d_r[cur ^ 1], cap, &no, 0, s); if (rc != QC_OK) return rc;
        n = no; cur ^= 1;
        if (n == 0) { ccm_started = true; return QC_OK; }      // CcmPeak(.., 0) still initialises itself on the first call
        tx_preemph_kernel<<<g, 256, 0, s>>>(d_r[cur], d_r[cur ^ 1], cap, n, preemph, d_x1); count_launch(); QC_CUDA_LAUNCH();
        tx_carry_kernel<<<gc, 64, 0, s>>>(d_r[cur], cap, n, d_x1, C); count_launch(); QC_CUDA_LAUNCH();
        cur ^= 1;
        if (ssb) {
            tx_promote_kernel<<<g, 256, 0, s>>>(d_r[cur], cap, d_c[0], cap, n); count_launch(); QC_CUDA_LAUNCH();
            rc = fTune1->run(d_c[0], cap, n, d_c[1], cap, &no, 0, s); if (rc != QC_OK) return rc;
            tx_level_ssb_kernel<<<(C + 31) / 32, 32, 0, s>>>(d_c[1], cap, d_r[cur], cap, n, d_inmax, C, P); count_launch(); QC_CUDA_LAUNCH();
        } else {
            tx_level_real_kernel<<<gc, 64, 0, s>>>(d_r[cur], cap, n, d_inmax, C, P); count_launch(); QC_CUDA_LAUNCH();
        }
        rc = fAudio2->run(d_r[cur], cap, n, d_r[cur ^ 1], cap, &no, 0, s); if (rc != QC_OK) return rc;
        cur ^= 1;
        if (ssb) {
            tx_promote_kernel<<<g, 256, 0, s>>>(d_r[cur], cap, d_c[0], cap, n); count_launch(); QC_CUDA_LAUNCH();
            rc = fTune2->run(d_c[0], cap, n, d_c[1], cap, &no, 0, s); if (rc != QC_OK) return rc;
            if (ccm_started) { tx_ccm_kernel<true><<<(C + CCM_WPB - 1) / CCM_WPB, 32 * CCM_WPB, 0, s>>>(d_c[1], cap, n, d_ccm, C, P, 2.0); count_launch(); QC_CUDA_LAUNCH(); }
            else { tx_scale_c_kernel<<<g, 256, 0, s>>>(d_c[1], cap, n, 2.0); count_launch(); QC_CUDA_LAUNCH(); }
            ccm_started = true;
            rc = fAudio3->run(d_c[1], cap, n, d_c[0], cap, &no, 0, s); if (rc != QC_OK) return rc;
            if (fInterp) { rc = fInterp->run(d_c[0], cap, n, d_out, os, &no, 0, s); if (rc != QC_OK) return rc; n = no; }
            else QC_CUDA(cudaMemcpy2DAsync(d_out, (size_t)os * sizeof(cd), d_c[0], (size_t)cap * sizeof(cd), (size_t)n * sizeof(cd), C, cudaMemcpyDeviceToDevice, s));
            const dim3 g2((unsigned)((n + 255) / 256 < 64 ? (n + 255) / 256 : 64), (unsigned)C);
            tx_scale_c_kernel<<<g2, 256, 0, s>>>(d_out, os, n, TX_CLIP16); count_launch(); QC_CUDA_LAUNCH();
        } else {
            if (ccm_started) { tx_ccm_kernel<false><<<(C + CCM_WPB - 1) / CCM_WPB, 32 * CCM_WPB, 0, s>>>(d_r[cur], cap, n, d_ccm, C, P, 1.0); count_launch(); QC_CUDA_LAUNCH(); }
            ccm_started = true;
            rc = fAudio3->run(d_r[cur], cap, n, d_r[cur ^ 1], cap, &no, 0, s); if (rc != QC_OK) return rc;
            cur ^= 1;
            if (fInterp) { rc = fInterp->run(d_r[cur], cap, n, d_r[cur ^ 1], cap, &no, 0, s); if (rc != QC_OK) return rc; n = no; cur ^= 1; }
            const dim3 g2((unsigned)((n + 255) / 256 < 64 ? (n + 255) / 256 : 64), (unsigned)C);
            tx_out_real_kernel<<<g2, 256, 0, s>>>(d_r[cur], cap, d_out, os, n); count_launch(); QC_CUDA_LAUNCH();
        }
        rc = run_alc(d_out, os, n, s); if (rc != QC_OK) return rc;
        if (n_out) *n_out = n;
        return QC_OK;
    }
};

}  // namespace qc

struct qcTxFilter { qc::TxFilter t; };

extern "C" {

qcTxFilter *quisk_cuda_tx_filter_create(int n_channels, int mode, int mic_sample_rate, double mic_preemphasis, double mic_clip, const qcTxTables *tables)
{
    if (qc::ensure_device() != QC_OK) return nullptr;
    if (!tables) { qc::set_error("tx_filter_create: tables missing"); return nullptr; }
    qcTxFilter *h = new qcTxFilter();
    if (h->t.init(n_channels, mode, mic_sample_rate, mic_preemphasis, mic_clip, *tables) != QC_OK) { h->t.release(); delete h; return nullptr; }
    return h;
}
void quisk_cuda_tx_filter_destroy(qcTxFilter *h) { if (h) { h->t.release(); delete h; } }
int quisk_cuda_tx_filter_set_alc(qcTxFilter *h, int enable) { return h ? h->t.set_alc(enable) : QC_EINVAL; }
int quisk_cuda_tx_filter_max_out(const qcTxFilter *h, int count) { return h ? h->t.max_out(count) : 0; }
int quisk_cuda_tx_filter_process(qcTxFilter *h, const void *d_in, long in_stride, int count, void *d_out, long out_stride, int *n_out, void *stream)
{
    if (!h) { qc::set_error("tx_filter_process: null handle"); return QC_EINVAL; }
    return h->t.process((const double2 *)d_in, in_stride, count, (double2 *)d_out, out_stride, n_out, (cudaStream_t)stream);
}

}  // extern "C"
