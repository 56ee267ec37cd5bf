// quisk_b200/csrc/wdsp_analyzer_nofma.cu -- WDSP's spectrum engine (wdsp/analyzer.c), SURVEY 8(f)4, for a batch of displays.
//
// The reference: Spectrum0 (analyzer.c:1536-1579) drops `buff_size` complex samples into a ring (as FLOATS, I and Q swapped);
// whenever `size` samples wait, a worker thread windows them and runs one complex transform (Cspectra, :670-745), Celiminate
// (:214-280) turns it into |X|^2 in display order with `clip` bins cut at both ends of the transform and fscL / fscH at the
// ends of the span, stitch (:555-600) hands the bins to the detector (:282-461: bins -> pixels as peak / rosenfell /
// average / sample / rms, or linear interpolation when there are more pixels than bins) and the averager (:463-553: peak
// hold, none, recursive, window, recursive on the log; 10 mlog10 -> float), once per pixel output; GetPixels (:1315-1334)
// copies the newest line.  One LO per sub-span (this reference build has dMAX_NUM_FFT = 1, comm.h:125), up to four stitched sub-spans, no
// calibration table (SetAnalyzer with n_fft = 1, fmin = fmax = 0); complex input (Cspectra / Celiminate) or real input
// (spectra / eliminate, :179-212, 602-668: the I rail alone, bins 0 .. size / 2).
//
// Here: D displays share one configuration and run side by side.  Per frame three launches: (1) one CTA per display:
// ring -> window -> shared-memory transform -> |X|^2 in Celiminate's order; (2) the detector, a thread per pixel.  Which
// bins a pixel takes is pure index arithmetic on pix_per_bin / det_offset -- it does not depend on the data -- so the HOST
// walks the reference's loops once per configuration and detector type (with the reference's own expressions, including
// the asymmetric `next_pix_count` of the rosenfell case) and writes down, per pixel, the bin range (or the one bin, or
// the two bins and the weight) its value comes from; the device then forms every pixel's sum / extremum over its bins in
// the reference's order; (3) the averager, a thread per pixel.  Compiled with --fmad=false: products and sums round
// separately, as in the reference's x86-64 build.  The window is built on the host with the same libm calls.
#include <cmath>
#include <cstring>
#include <vector>
#include "fft_device.cuh"
#include "wdsp_internal.h"
#include "../../include/quisk_cuda_wdsp.h"

namespace qc {

static constexpr int AN_MAX_PIXOUTS = 4, AN_MAX_AVERAGE = 60, AN_MAX_PIXELS = 16384, AN_BUFF_MULT = 2;    // comm.h:126-139

// ---- (1) window, transform, |X|^2 in display order ----------------------------------------------------------------
struct AnFrameParams {
    const cd *ring; int bsize, idx0, size;      // [D][bsize] samples as Spectrum0 stored them (x = I, y = Q)
    const double *window; const cd *tw;
    double *bins; int m, bin_off;               // [D][m]: this sub-span's bins start at bin_off
    int begin0, end0, begin1, end1, flip;       // Celiminate's two runs over the transform's output (eliminate's one run for real input)
    int real, out_size;                         // real input: only I enters the transform, bins 0 .. size / 2 come out
};

template <int BPT>
__global__ void __launch_bounds__(256) an_frame_kernel(AnFrameParams P)
{
    extern __shared__ double smem_raw[];
    const int n = P.size, d = blockIdx.x, lane = threadIdx.x, lanes = blockDim.x;
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *s = twl + fft_tw_entries(n);
    fft_stage_twiddles(twl, P.tw, n);
    const cd *ring = P.ring + (size_t)d * P.bsize;
    for (int i = lane; i < n; i += lanes) {
        int j = P.idx0 + i;
        if (j >= P.bsize) j -= P.bsize;
        const cd v = ring[j];
        const double w = P.window[i];
        s[fsw(i)] = make_double2(w * v.x, P.real ? 0.0 : w * v.y);   // analyzer.c:688-689 (complex), :620 (real: I only)
    }
    __syncthreads();
    fft_smem<BPT>(s, n, twl, -1, lane, lanes);
    double *bins = P.bins + (size_t)d * P.m + P.bin_off;
    const int n0 = P.end0 > P.begin0 ? P.end0 - P.begin0 : 0, n1 = P.end1 > P.begin1 ? P.end1 - P.begin1 : 0, ilim = P.out_size - 1;
    for (int k = lane; k < n0 + n1; k += lanes) {
        int i = k < n0 ? P.begin0 + k : P.begin1 + (k - n0);
        if (P.flip) i = ilim - i;                               // analyzer.c:250-263: the same runs walked from the other end
        const cd X = s[fsw(i)];
        bins[k] = X.x * X.x + X.y * X.y;
    }
}

// ---- (2) detector: per pixel, where its value comes from ---------------------------------------------------------------
struct AnPix {              // one entry per pixel and detector type
    int kind;               // 0 untouched, 1 max over [a, b), 2 mean, 3 rms, 4 the bin a, 5 a + weights, 6 rosenfell event b (first bin a)
    int a, b;
    double w0, w1;          // kind 5: bins[a] * w0 + bins[a + 1] * w1
};
struct AnEvent { int s, e, odd; };       // rosenfell: bins [s, e) closed into one value (analyzer.c:325-366)

__global__ void an_rose_events_kernel(const double *bins, int m, const AnEvent *ev, int n_ev, double *ev_val /*[D][n_ev][3]*/)
{
    const int d = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ev) return;
    const double *b = bins + (size_t)d * m;
    const AnEvent E = ev[k];
    double mini = 1.0e300, maxi = -1.0e300;
    int rose = 0, fell = 0;
    for (int i = E.s; i < E.e; i++) {
        if (b[i] < mini) mini = b[i];
        if (b[i] > maxi) maxi = b[i];
        if (i < E.e - 1) { if (b[i + 1] > b[i]) rose = 1; if (b[i + 1] < b[i]) fell = 1; }
    }
    double *o = ev_val + ((size_t)d * n_ev + k) * 3;
    o[0] = mini; o[1] = maxi; o[2] = (rose && fell) ? 1.0 : 0.0;
}

__global__ void an_detect_kernel(const double *bins, int m, const AnPix *plan, int n_pix, double inv_enb, int ampl_comp,
                                 const AnEvent *ev, const double *ev_val, int n_ev, double *t_pixels /*[D][AN_MAX_PIXELS]*/)
{
    const int d = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pix) return;
    const double *b = bins + (size_t)d * m;
    const AnPix q = plan[p];
    double *out = t_pixels + (size_t)d * AN_MAX_PIXELS + p;
    switch (q.kind) {
    case 1: { double v = -1.0e300; for (int i = q.a; i < q.b; i++) if (b[i] > v) v = b[i]; *out = v; break; }
    case 2: { double s = 0.0; for (int i = q.a; i < q.b; i++) s += b[i]; *out = s / (double)(q.b - q.a) * inv_enb; break; }
    case 3: { double s = 0.0; for (int i = q.a; i < q.b; i++) s += b[i] * b[i]; *out = sqrt(s / (double)(q.b - q.a)) * inv_enb; break; }
    case 4: *out = b[q.a] * inv_enb; break;
    case 5: { double v = b[q.a] * q.w0 + b[q.a + 1] * q.w1; if (ampl_comp) v *= inv_enb; *out = v; break; }
    case 6: {
        const double *e = ev_val + ((size_t)d * n_ev + q.b) * 3;
        const double prev = q.b > 0 ? e[-3 + 1] : -1.0e300;
        if (e[2] != 0.0) *out = ev[q.b].odd ? (prev > e[1] ? prev : e[1]) : e[0];
        else *out = e[1];
        break;
    }
    case 7: *out = -1.0e300; break;                             // peak detector, a pixel no bin falls into
    default: break;
    }
}

// ---- (3) averager (analyzer.c:463-553) --------------------------------------------------------------------------------------
struct AnAvgParams {
    int mode, n_pix, growing, in_idx, out_idx, norm;
    double backmult, scale, factor; float norm_onehz;
    const double *t_pixels; double *av_sum; double *av_buff /*[D][AN_MAX_AVERAGE][n_pix]*/; float *pixels /*[D][n_pix]*/;
    const double *mtable;
};

__global__ void an_average_kernel(AnAvgParams P)
{
    const int d = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_pix) return;
    const double t = P.t_pixels[(size_t)d * AN_MAX_PIXELS + i];
    double *sum = P.av_sum + (size_t)d * AN_MAX_PIXELS + i;
    const double cdv = 1.0;                                     // the calibration factor without a table (analyzer.c:1203-1204)
    float px;
    switch (P.mode) {
    case -1: if (t > *sum) *sum = t; px = (float)(10.0 * mlog10_dev(P.mtable, P.scale * cdv * *sum + 1.0e-60)); break;
    case 1: *sum = P.backmult * *sum + (1.0 - P.backmult) * t; px = (float)(10.0 * mlog10_dev(P.mtable, P.scale * cdv * *sum + 1.0e-60)); break;
    case 2: {
        double *buf = P.av_buff + (size_t)d * AN_MAX_AVERAGE * P.n_pix;
        if (P.growing) *sum += t;
        else *sum += t - buf[(size_t)P.out_idx * P.n_pix + i];
        buf[(size_t)P.in_idx * P.n_pix + i] = t;
        px = (float)(10.0 * mlog10_dev(P.mtable, cdv * *sum * P.factor + 1.0e-60));
        break;
    }
    case 3: *sum = P.backmult * *sum + (1.0 - P.backmult) * (10.0 * mlog10_dev(P.mtable, P.scale * cdv * t + 1e-60)); px = (float)*sum; break;
    default: px = (float)(10.0 * mlog10_dev(P.mtable, P.scale * cdv * t + 1.0e-60)); break;
    }
    if (P.norm) px += P.norm_onehz;
    P.pixels[(size_t)d * P.n_pix + i] = px;
}

__global__ void an_store_kernel(const cd *in, long in_stride, cd *ring, int bsize, int idx, int n)
{
    const int d = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cd v = in[(size_t)d * in_stride + i];
    // Ipointer[i] = (float)pbuff[2 i + 1], Qpointer[i] = (float)pbuff[2 i + 0]  (analyzer.c:1551-1555, dINREAL = float, comm.h:131)
    ring[(size_t)d * bsize + idx + i] = make_double2((double)(float)v.y, (double)(float)v.x);
}

__global__ void an_fill_kernel(double *p, size_t n, double v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

static double an_bessi0(double x)
{   // analyzer.c:33-50
    double ax, ans, y;
    if ((ax = fabs(x)) < 3.75) {
        y = x / 3.75; y = y * y;
        ans = 1.0 + y * (3.5156229 + y * (3.0899424 + y * (1.2067492 + y * (0.2659732 + y * (0.360768e-1 + y * 0.45813e-2)))));
    } else {
        y = 3.75 / ax;
        ans = (exp(ax) / sqrt(ax)) * (0.39894228 + y * (0.1328592e-1 + y * (0.225319e-2 + y * (-0.157565e-2 + y * (0.916281e-2 + y * (-0.2057706e-1
              + y * (0.2635537e-1 + y * (-0.1647633e-1 + y * 0.392377e-2))))))));
    }
    return ans;
}

struct Analyzer {
    int D = 0, max_size = 0;
    // SetAnalyzer's arguments and what it derives (analyzer.c:999-1137)
    int num_pixout = 0, type = 1, flip = 0, size = -1, buff_size = 0, window_type = -1, overlap = 0, clip = 0, num_pixels = -1, incr = 0, out_size = 0, max_w = 0;
    double pi_alpha = 0.0, fsclipL = 0.0, fsclipH = 0.0, scale = 0.0, pix_per_bin = 0.0, bin_per_pix = 0.0, det_offset = 0.0;
    double inv_coherent_gain = 1.0, inherent_power_gain = 1.0, inv_enb = 1.0, norm_oneHz = 0.0;
    int fscL = 0, fscH = 0, sample_rate = 0, m = 0;
    static constexpr int MAX_STITCH = 4;        // comm.h:124
    int num_stitch = 1, begin_ss = 0, end_ss = 0, stitch_flag = 0;
    int begin0[MAX_STITCH] = {0}, end0[MAX_STITCH] = {0}, begin1[MAX_STITCH] = {0}, end1[MAX_STITCH] = {0}, ss_bins[MAX_STITCH] = {0}, ss_off[MAX_STITCH] = {0};
    bool configured = false, span_empty = false;
    // per pixel output (SetDisplay*, analyzer.c:1582-1675)
    int det_type[AN_MAX_PIXOUTS] = {0}, av_mode[AN_MAX_PIXOUTS] = {0}, num_average[AN_MAX_PIXOUTS] = {0}, normalize[AN_MAX_PIXOUTS] = {0};
    int avail_frames[AN_MAX_PIXOUTS] = {0}, av_in_idx[AN_MAX_PIXOUTS] = {0}, av_out_idx[AN_MAX_PIXOUTS] = {0};
    double av_backmult[AN_MAX_PIXOUTS] = {0};
    long frames_done = 0, frames_read[AN_MAX_PIXOUTS] = {0};
    // ring bookkeeping (the same for every display of the batch)
    int bsize = 0, in_index[MAX_STITCH] = {0}, out_index[MAX_STITCH] = {0}, have[MAX_STITCH] = {0}, busy[MAX_STITCH] = {0};
    // device
    cd *d_ring = nullptr; double *d_window = nullptr, *d_bins = nullptr;
    double *d_t[AN_MAX_PIXOUTS] = {nullptr}, *d_sum[AN_MAX_PIXOUTS] = {nullptr}, *d_avbuf[AN_MAX_PIXOUTS] = {nullptr};
    int avbuf_pix[AN_MAX_PIXOUTS] = {0};
    float *d_pix[AN_MAX_PIXOUTS] = {nullptr};
    AnPix *d_plan[5] = {nullptr}; AnEvent *d_ev = nullptr; double *d_evval = nullptr; int n_ev = 0;
    bool plan_ok[5] = {false};
    const cd *tw = nullptr;

    int init(int D_, int max_size_);
    void release();
    int set(int n_pixout, int typ, int flp, int sz, int bf_sz, int win_type, double pi, int ovrlp, int clp, double fscLin, double fscHin, int n_pix, int n_stch, int max_w_);
    int new_window(int type, int sz, double PiAlpha);
    int build_plan(int det);
    int fill(double *p, double v);
    int set_average_mode(int po, int mode);
    int frame(int ss, cudaStream_t s);
    int stitch(cudaStream_t s);
    int dispatch(cudaStream_t s);
    int spectrum0(int ss, const cd *d_in, long in_stride, cudaStream_t s);
};

int Analyzer::fill(double *p, double v)
{
    an_fill_kernel<<<148, 256>>>(p, (size_t)D * AN_MAX_PIXELS, v);
    count_launch();
    QC_CUDA_LAUNCH();
    QC_CUDA(cudaDeviceSynchronize());           // (the frames run on the caller's stream)
    return QC_OK;
}

int Analyzer::init(int D_, int max_size_)
{   // XCreateAnalyzer, analyzer.c:1140-1236
    D = D_; max_size = max_size_;
    bsize = max_size * AN_BUFF_MULT;
    QC_CUDA(cudaMalloc((void **)&d_ring, (size_t)MAX_STITCH * D * bsize * sizeof(cd)));
    QC_CUDA(cudaMemset(d_ring, 0, (size_t)MAX_STITCH * D * bsize * sizeof(cd)));
    QC_CUDA(cudaMalloc((void **)&d_window, (size_t)max_size * sizeof(double)));
    QC_CUDA(cudaMalloc((void **)&d_bins, (size_t)MAX_STITCH * D * max_size * sizeof(double)));
    for (int i = 0; i < AN_MAX_PIXOUTS; i++) {
        QC_CUDA(cudaMalloc((void **)&d_t[i], (size_t)D * AN_MAX_PIXELS * sizeof(double)));
        QC_CUDA(cudaMalloc((void **)&d_sum[i], (size_t)D * AN_MAX_PIXELS * sizeof(double)));
        QC_CUDA(cudaMalloc((void **)&d_pix[i], (size_t)D * AN_MAX_PIXELS * sizeof(float)));
        QC_CUDA(cudaMemset(d_t[i], 0, (size_t)D * AN_MAX_PIXELS * sizeof(double)));
        QC_CUDA(cudaMemset(d_sum[i], 0, (size_t)D * AN_MAX_PIXELS * sizeof(double)));
        QC_CUDA(cudaMemset(d_pix[i], 0, (size_t)D * AN_MAX_PIXELS * sizeof(float)));
    }
    for (int k = 0; k < 5; k++) QC_CUDA(cudaMalloc((void **)&d_plan[k], (size_t)AN_MAX_PIXELS * sizeof(AnPix)));
    return QC_OK;
}

void Analyzer::release()
{
    cudaDeviceSynchronize();
    if (d_ring) cudaFree(d_ring);
    if (d_window) cudaFree(d_window);
    if (d_bins) cudaFree(d_bins);
    for (int i = 0; i < AN_MAX_PIXOUTS; i++) { if (d_t[i]) cudaFree(d_t[i]); if (d_sum[i]) cudaFree(d_sum[i]); if (d_pix[i]) cudaFree(d_pix[i]); if (d_avbuf[i]) cudaFree(d_avbuf[i]); }
    for (int k = 0; k < 5; k++) if (d_plan[k]) cudaFree(d_plan[k]);
    if (d_ev) cudaFree(d_ev);
    if (d_evval) cudaFree(d_evval);
}

int Analyzer::new_window(int type, int sz, double PiAlpha)
{   // analyzer.c:52-176, the same libm calls in the same order
    std::vector<double> w((size_t)sz);
    const double PI = 3.1415926535897932;
    double arg0, arg1, cgsum = 0.0, igsum = 0.0;
    int i;
    switch (type) {
    case 0:
        inv_coherent_gain = 1.0; igsum = (double)sz;
        for (i = 0; i < sz; i++) w[i] = inv_coherent_gain * 1.0;
        break;
    case 1:
        arg0 = 2.0 * PI / ((double)sz - 1.0);
        for (i = 0; i < sz; i++) { arg1 = arg0 * (double)i; w[i] = 0.35875 - 0.48829 * cos(arg1) + 0.14128 * cos(2.0 * arg1) - 0.01168 * cos(3.0 * arg1); cgsum += w[i]; igsum += w[i] * w[i]; }
        break;
    case 2:
        arg0 = 2.0 * PI / ((double)sz - 1.0);
        for (i = 0; i < sz; i++) { w[i] = 0.5 * (1.0 - cos((double)i * arg0)); cgsum += w[i]; igsum += w[i] * w[i]; }
        break;
    case 3:
        arg0 = 2.0 * PI / ((double)sz - 1.0);
        for (i = 0; i < sz; i++) {
            arg1 = arg0 * (double)i;
            w[i] = 0.21557895 - 0.41663158 * cos(arg1) + 0.277263158 * cos(2.0 * arg1) - 0.083578947 * cos(3.0 * arg1) + 0.006947368 * cos(4.0 * arg1);
            cgsum += w[i]; igsum += w[i] * w[i];
        }
        break;
    case 4:
        arg0 = 2.0 * PI / ((double)sz - 1.0);
        for (i = 0; i < sz; i++) { w[i] = (0.54 - 0.46 * cos((double)i * arg0)); cgsum += w[i]; igsum += w[i] * w[i]; }
        break;
    case 5:
        arg0 = an_bessi0(PiAlpha); arg1 = (double)(sz - 1);
        for (i = 0; i < sz; ++i) { w[i] = an_bessi0(PiAlpha * sqrt(1.0 - pow(2.0 * (double)i / arg1 - 1.0, 2))) / arg0; cgsum += w[i]; igsum += w[i] * w[i]; }
        break;
    case 6:
        arg0 = 2.0 * PI / ((double)sz - 1.0);
        for (i = 0; i < sz; ++i) {
            arg1 = cos(arg0 * (double)i);
            w[i] = +6.3964424114390378e-02 + arg1 * (-2.3993864599352804e-01 + arg1 * (+3.5015956323820469e-01 + arg1 * (-2.4774111897080783e-01
                   + arg1 * (+8.5438256055858031e-02 + arg1 * (-1.2320203369293225e-02 + arg1 * (+4.3778825791773474e-04))))));
            cgsum += w[i]; igsum += w[i] * w[i];
        }
        break;
    default:
        set_error("analyzer: window type %d (0 .. 6)", type);
        return QC_EINVAL;
    }
    if (type != 0) {
        inv_coherent_gain = (double)sz / cgsum;
        for (i = 0; i < sz; i++) w[i] *= inv_coherent_gain;
    }
    inherent_power_gain = igsum / (double)sz;
    inv_enb = 1.0 / (inherent_power_gain * inv_coherent_gain * inv_coherent_gain);
    QC_CUDA(cudaMemcpy(d_window, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice));
    return QC_OK;
}

static double an_host_mlog10(double val)
{   // meterlog10.c:547-554 with the table rebuilt from its definition (see mlog10_table)
    unsigned long long N;
    memcpy(&N, &val, sizeof(N));
    const int e = (int)((N >> 52) & 2047ull) - 1023;
    const int mm = (int)((N >> 41) & 2047ull);
    return 0.301029995663981 * ((double)e + log10(1.0 + (double)mm / 2048.0) / log10(2.0));
}

int Analyzer::set(int n_pixout, int typ, int flp, int sz, int bf_sz, int win_type, double pi, int ovrlp, int clp, double fscLin, double fscHin, int n_pix, int n_stch, int max_w_)
{
    if (n_stch < 1 || n_stch > MAX_STITCH) { set_error("analyzer: %d sub-spans (1 .. %d)", n_stch, MAX_STITCH); return QC_EINVAL; }
    if (n_pixout < 1 || n_pixout > AN_MAX_PIXOUTS) { set_error("analyzer: %d pixel outputs (1 .. %d)", n_pixout, AN_MAX_PIXOUTS); return QC_EINVAL; }
    if (sz < 64 || sz > 8192 || (sz & (sz - 1)) || sz > max_size) { set_error("analyzer: transform size %d (a power of two, 64 .. 8192, <= the %d given at creation)", sz, max_size); return QC_EINVAL; }
    if (bf_sz < 1 || bsize % bf_sz) { set_error("analyzer: buffer size %d must divide the ring of %d samples (analyzer.c:1569)", bf_sz, bsize); return QC_EINVAL; }
    if (ovrlp < 0 || ovrlp >= sz || clp < 0 || fscLin < 0.0 || fscHin < 0.0 || n_pix < 2 || n_pix > AN_MAX_PIXELS) { set_error("analyzer: overlap / clip / pixel count out of range"); return QC_EINVAL; }
    QC_CUDA(cudaDeviceSynchronize());
    if (typ != 0 && typ != 1) { set_error("analyzer: input type %d (0 real, 1 complex)", typ); return QC_EINVAL; }
    num_pixout = n_pixout; type = typ; flip = flp ? 1 : 0; buff_size = bf_sz; overlap = ovrlp; clip = clp; fsclipL = fscLin; fsclipH = fscHin;
    if (sz != size || win_type != window_type || pi != pi_alpha) { int rc = new_window(win_type, sz, pi); if (rc) return rc; }
    if (sz != size) { tw = fft_twiddles(sz); if (!tw) { set_error("analyzer: twiddle table allocation failed"); return QC_ENOMEM; } }
    size = sz; window_type = win_type; pi_alpha = pi; max_w = max_w_;
    norm_oneHz = sample_rate > 0 ? 10.0 * an_host_mlog10(1.0 / ((double)sample_rate / (double)size)) : 0.0;
    incr = size - overlap;
    num_pixels = n_pix;
    if (type == 0) { out_size = size / 2 + 1; scale = 4.0 / ((double)size * (double)size); }        // analyzer.c:1078-1087
    else { out_size = size; scale = 1.0 / ((double)size * (double)size); }
    num_stitch = n_stch;
    fscL = (int)fsclipL; fscH = (int)fsclipH;
    const int usable = out_size - 1 - 2 * clip;
    if (usable <= 0) { set_error("analyzer: clip %d leaves no bins of a %d-point transform", clip, size); return QC_EINVAL; }
    // sub-spans that the span clips remove altogether are skipped (analyzer.c:1093-1104)
    begin_ss = 0; end_ss = num_stitch - 1;
    for (int k = 0; k < MAX_STITCH; k++) ss_bins[k] = 0;
    while (fscL >= usable && begin_ss < num_stitch) { fscL -= usable; begin_ss++; }
    while (fscH >= usable && end_ss >= 0) { fscH -= usable; end_ss--; }
    if (begin_ss > end_ss) { set_error("analyzer: span clips %g, %g leave nothing of %d sub-spans of %d bins", fsclipL, fsclipH, num_stitch, usable); return QC_EINVAL; }
    pix_per_bin = (double)num_pixels / ((double)(num_stitch * (out_size - 1 - 2 * clip)) - fsclipL - fsclipH - 1.0);
    det_offset = -pix_per_bin * (fsclipL - floor(fsclipL));
    bin_per_pix = ((double)(num_stitch * (out_size - 1 - 2 * clip)) - 1.0 - fsclipL - fsclipH) / ((double)num_pixels - 1.0);
    // Celiminate's two runs per sub-span (analyzer.c:220-246): the span clips act on the first and the last one
    m = 0;
    for (int ss = begin_ss; ss <= end_ss; ss++) {
        if (type == 0) {            // eliminate, analyzer.c:179-212: one run
            begin0[ss] = ss == begin_ss ? fscL + clip : clip;
            end0[ss] = ss == end_ss ? out_size - 1 - clip - fscH : out_size - 1 - clip;
            begin1[ss] = end1[ss] = 0;
        } else {
            if (ss == begin_ss) { begin0[ss] = out_size / 2 + 1 + clip + fscL; begin1[ss] = begin0[ss] > out_size ? begin0[ss] - out_size : 0; }
            else { begin0[ss] = out_size / 2 + 1 + clip; begin1[ss] = 0; }
            if (ss == end_ss) { end1[ss] = out_size / 2 - clip - fscH; end0[ss] = end1[ss] < 0 ? out_size + end1[ss] : out_size; }
            else { end0[ss] = out_size; end1[ss] = out_size / 2 - clip; }
        }
        ss_bins[ss] = (end0[ss] > begin0[ss] ? end0[ss] - begin0[ss] : 0) + (end1[ss] > begin1[ss] ? end1[ss] - begin1[ss] : 0);
        ss_off[ss] = m;
        m += ss_bins[ss];
    }
    if (m < 2) { set_error("analyzer: %d bins left", m); return QC_EINVAL; }
    for (int k = 0; k < 5; k++) plan_ok[k] = false;
    for (int k = 0; k < MAX_STITCH; k++) in_index[k] = out_index[k] = have[k] = busy[k] = 0;
    stitch_flag = 0;
    configured = true;
    return QC_OK;
}

int Analyzer::build_plan(int det)
{   // the index arithmetic of detector() (analyzer.c:282-461), walked once on the host
    std::vector<AnPix> plan((size_t)num_pixels);
    for (auto &q : plan) { q.kind = 0; q.a = q.b = 0; q.w0 = q.w1 = 0.0; }
    std::vector<AnEvent> events;
    int i, imin, ilim, pix_count = 0;
    if (pix_per_bin <= 1.0) {
        imin = fsclipL == floor(fsclipL) ? 0 : 1;
        ilim = fsclipH == floor(fsclipH) ? m : m - 1;
        auto pc = [&](int k) { int p = (int)(det_offset + (double)k * pix_per_bin); return p >= num_pixels ? num_pixels - 1 : p; };
        switch (det) {
        case 0:
            for (auto &q : plan) q.kind = 7;
            for (i = imin; i < ilim; i++) {
                const int p = pc(i);
                if (p < 0) { set_error("analyzer: bin %d maps in front of the first pixel", i); return QC_EINVAL; }
                if (plan[p].kind == 7) { plan[p].kind = 1; plan[p].a = i; plan[p].b = i + 1; }
                else if (plan[p].b == i) plan[p].b = i + 1;
                else { set_error("analyzer: the bins of pixel %d are not one run", p); return QC_EINVAL; }
            }
            break;
        case 1: {
            int start = imin;
            for (i = imin; i < ilim; i++) {
                pix_count = pc(i);
                const int next_pix_count = (int)((double)(i + 1) * pix_per_bin);
                if (next_pix_count == pix_count && i < ilim - 1) continue;
                if (pix_count < 0) { set_error("analyzer: bin %d maps in front of the first pixel", i); return QC_EINVAL; }
                AnEvent e; e.s = start; e.e = i + 1; e.odd = pix_count & 1;
                plan[pix_count].kind = 6; plan[pix_count].a = start; plan[pix_count].b = (int)events.size();     // a later event for the same pixel overwrites
                events.push_back(e);
                start = i + 1;
            }
            break;
        }
        case 2: case 4: {
            int bcount = 0, start = imin, last_pix_count;
            for (i = imin; i < ilim; i++) {
                last_pix_count = pix_count;
                pix_count = pc(i);
                if (pix_count == last_pix_count) bcount++;
                else {
                    if (bcount == 0) { set_error("analyzer: the first bin does not fall into pixel 0 (the reference divides 0 by 0 there)"); return QC_EINVAL; }
                    plan[last_pix_count].kind = det == 2 ? 2 : 3; plan[last_pix_count].a = start; plan[last_pix_count].b = start + bcount;
                    start = i; bcount = 1;
                }
                if (i == ilim - 1) { plan[pix_count].kind = det == 2 ? 2 : 3; plan[pix_count].a = start; plan[pix_count].b = start + bcount; }
            }
            break;
        }
        case 3: {
            int bcount = 0, last_pix_count;
            for (i = imin; i < ilim; i++) {
                last_pix_count = pix_count;
                pix_count = pc(i);
                if (pix_count == last_pix_count) bcount++;
                else {
                    const int src = i - bcount / 2 - 1;
                    if (src < 0) { set_error("analyzer: sample detector reads in front of the first bin"); return QC_EINVAL; }
                    plan[last_pix_count].kind = 4; plan[last_pix_count].a = src;
                    bcount = 1;
                }
                if (i == ilim - 1) { plan[pix_count].kind = 4; plan[pix_count].a = i - bcount / 2; }
            }
            break;
        }
        default:
            set_error("analyzer: detector type %d (0 .. 4)", det);
            return QC_EINVAL;
        }
    } else {
        double pix_pos = fsclipL - floor(fsclipL);
        for (i = 1; i < m; i++) {
            while (pix_pos < ((double)i + 1.0e-06) && pix_count < num_pixels) {
                const double frac = pix_pos - (double)(i - 1);
                plan[pix_count].kind = 5; plan[pix_count].a = i - 1; plan[pix_count].w0 = 1.0 - frac; plan[pix_count].w1 = frac;
                pix_count++;
                pix_pos += bin_per_pix;
            }
        }
    }
    QC_CUDA(cudaMemcpy(d_plan[det], plan.data(), plan.size() * sizeof(AnPix), cudaMemcpyHostToDevice));
    if (det == 1 && pix_per_bin <= 1.0) {
        if (d_ev) { cudaFree(d_ev); d_ev = nullptr; }
        if (d_evval) { cudaFree(d_evval); d_evval = nullptr; }
        n_ev = (int)events.size();
        if (n_ev > 0) {
            QC_CUDA(cudaMalloc((void **)&d_ev, (size_t)n_ev * sizeof(AnEvent)));
            QC_CUDA(cudaMemcpy(d_ev, events.data(), (size_t)n_ev * sizeof(AnEvent), cudaMemcpyHostToDevice));
            QC_CUDA(cudaMalloc((void **)&d_evval, (size_t)D * n_ev * 3 * sizeof(double)));
        }
    }
    plan_ok[det] = true;
    return QC_OK;
}

int Analyzer::set_average_mode(int po, int mode)
{   // SetDisplayAverageMode, analyzer.c:1594-1623
    if (av_mode[po] == mode) return QC_OK;
    QC_CUDA(cudaDeviceSynchronize());
    av_mode[po] = mode;
    switch (mode) {
    case 1: return fill(d_sum[po], 1.0e-12);
    case 2: avail_frames[po] = 0; av_in_idx[po] = 0; av_out_idx[po] = 0; return QC_OK;
    case 3: return fill(d_sum[po], -160.0);
    default: return fill(d_sum[po], 0.0);
    }
}

int Analyzer::frame(int ss, cudaStream_t s)
{   // Cspectra + Celiminate for sub-span ss, the frame that starts at out_index[ss] (a sub-span the clips removed only reports in)
    if (ss < begin_ss || ss > end_ss) return QC_OK;
    AnFrameParams F;
    F.ring = d_ring + (size_t)ss * D * bsize; F.bsize = bsize; F.idx0 = out_index[ss]; F.size = size; F.window = d_window; F.tw = tw;
    F.bins = d_bins; F.m = m; F.bin_off = ss_off[ss];
    F.begin0 = begin0[ss]; F.end0 = end0[ss]; F.begin1 = begin1[ss]; F.end1 = end1[ss]; F.flip = flip; F.real = type == 0; F.out_size = out_size;
    const int lanes = fft_threads(size);
    const size_t sh = ((size_t)size + fft_tw_entries(size)) * sizeof(cd);
    if (size > 4096) {
        if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(an_frame_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        an_frame_kernel<2><<<D, lanes, sh, s>>>(F);
    } else {
        if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(an_frame_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        an_frame_kernel<1><<<D, lanes, sh, s>>>(F);
    }
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

int Analyzer::stitch(cudaStream_t s)
{   // stitch (analyzer.c:555-600): the sub-spans' bins lie side by side in d_bins already; detector and averager per pixel output
    const double *mt = mlog10_table();
    if (!mt) { set_error("analyzer: table allocation failed"); return QC_ENOMEM; }
    const dim3 gp((num_pixels + 127) / 128, D);
    for (int i = 0; i < num_pixout; i++) {
        int k = i;
        for (int j = i - 1; j >= 0; j--) if (det_type[i] == det_type[j]) k = j;           // analyzer.c:575-589: an earlier output with the same detector
        if (k == i) {
            const int det = det_type[i];
            if (det < 0 || det > 4) { set_error("analyzer: detector type %d (0 .. 4)", det); return QC_EINVAL; }
            if (!plan_ok[det]) { int rc = build_plan(det); if (rc) return rc; }
            if (det == 1 && pix_per_bin <= 1.0 && n_ev > 0) {
                an_rose_events_kernel<<<dim3((n_ev + 127) / 128, D), 128, 0, s>>>(d_bins, m, d_ev, n_ev, d_evval);
                count_launch();
                QC_CUDA_LAUNCH();
            }
            an_detect_kernel<<<gp, 128, 0, s>>>(d_bins, m, d_plan[det], num_pixels, inv_enb, det >= 2 ? 1 : 0, d_ev, d_evval, n_ev, d_t[i]);
            count_launch();
            QC_CUDA_LAUNCH();
        } else {
            QC_CUDA(cudaMemcpy2DAsync(d_t[i], (size_t)AN_MAX_PIXELS * sizeof(double), d_t[k], (size_t)AN_MAX_PIXELS * sizeof(double),
                                      (size_t)num_pixels * sizeof(double), D, cudaMemcpyDeviceToDevice, s));
        }
        AnAvgParams A;
        memset(&A, 0, sizeof(A));
        A.mode = av_mode[i]; A.n_pix = num_pixels; A.backmult = av_backmult[i]; A.scale = scale; A.norm = normalize[i]; A.norm_onehz = (float)norm_oneHz;
        A.t_pixels = d_t[i]; A.av_sum = d_sum[i]; A.pixels = d_pix[i]; A.mtable = mt;
        if (av_mode[i] == 2) {
            if (!d_avbuf[i] || avbuf_pix[i] != num_pixels) {
                QC_CUDA(cudaStreamSynchronize(s));
                if (d_avbuf[i]) { cudaFree(d_avbuf[i]); d_avbuf[i] = nullptr; }
                QC_CUDA(cudaMalloc((void **)&d_avbuf[i], (size_t)D * AN_MAX_AVERAGE * num_pixels * sizeof(double)));
                QC_CUDA(cudaMemset(d_avbuf[i], 0, (size_t)D * AN_MAX_AVERAGE * num_pixels * sizeof(double)));
                avbuf_pix[i] = num_pixels;
            }
            A.av_buff = d_avbuf[i];
            if (avail_frames[i] < num_average[i]) { A.growing = 1; A.factor = scale / (double)++avail_frames[i]; }
            else { A.growing = 0; A.factor = scale / (double)avail_frames[i]; }
            A.in_idx = av_in_idx[i]; A.out_idx = av_out_idx[i];
        }
        an_average_kernel<<<gp, 128, 0, s>>>(A);
        count_launch();
        QC_CUDA_LAUNCH();
        if (av_mode[i] == 2) {
            if (!A.growing && ++av_out_idx[i] == AN_MAX_AVERAGE) av_out_idx[i] = 0;
            if (++av_in_idx[i] == AN_MAX_AVERAGE) av_in_idx[i] = 0;
        }
    }
    frames_done++;
    return QC_OK;
}

int Analyzer::dispatch(cudaStream_t s)
{   // the dispatcher's turns (sendbuf, analyzer.c:884-911): a sub-span with `size` samples waiting sends one frame and is then
    // busy until the stitch that uses it has been made (Cspectra, :713-733)
    bool again = true;
    while (again) {
        again = false;
        for (int ss = 0; ss < num_stitch; ss++) {
            if (busy[ss] || have[ss] < size) continue;
            busy[ss] = 1;
            int rc = frame(ss, s); if (rc) return rc;
            if ((out_index[ss] += incr) >= bsize) out_index[ss] -= bsize;
            have[ss] -= incr;
            stitch_flag |= 1 << ss;
            if (stitch_flag == (1 << num_stitch) - 1) {
                stitch_flag = 0;
                for (int k = 0; k < MAX_STITCH; k++) busy[k] = 0;
                rc = stitch(s); if (rc) return rc;
                again = true;
            }
        }
    }
    return QC_OK;
}

int Analyzer::spectrum0(int ss, const cd *d_in, long in_stride, cudaStream_t s)
{   // Spectrum0 (analyzer.c:1536-1579), then the dispatcher
    if (!configured) { set_error("analyzer: SetAnalyzer first"); return QC_EINVAL; }
    if (ss < 0 || ss >= num_stitch) { set_error("analyzer: sub-span %d of %d", ss, num_stitch); return QC_EINVAL; }
    an_store_kernel<<<dim3((buff_size + 127) / 128, D), 128, 0, s>>>(d_in, in_stride, d_ring + (size_t)ss * D * bsize, bsize, in_index[ss], buff_size);
    count_launch();
    QC_CUDA_LAUNCH();
    if (have[ss] > max_w) {         // samples arrive faster than frames leave (a sub-span waiting for its neighbours): skip some, analyzer.c:1559-1565
        if ((out_index[ss] += have[ss] - max_w) >= bsize) out_index[ss] -= bsize;
        have[ss] = max_w;
    }
    have[ss] += buff_size;
    if ((in_index[ss] += buff_size) >= bsize) in_index[ss] = 0;
    return dispatch(s);
}

}  // namespace qc

struct qcAnalyzer { qc::Analyzer a; };

extern "C" {

qcAnalyzer *quisk_cuda_analyzer_create(int n_displays, int max_size)
{
    if (qc::ensure_device() != QC_OK) return nullptr;
    if (n_displays < 1 || max_size < 64 || max_size > 8192 || (max_size & (max_size - 1))) { qc::set_error("analyzer_create: %d displays, max size %d (a power of two, 64 .. 8192)", n_displays, max_size); return nullptr; }
    qcAnalyzer *h = new qcAnalyzer();
    if (h->a.init(n_displays, max_size) != QC_OK) { h->a.release(); delete h; return nullptr; }
    return h;
}

void quisk_cuda_analyzer_destroy(qcAnalyzer *h) { if (h) { h->a.release(); delete h; } }

int quisk_cuda_analyzer_set(qcAnalyzer *h, int n_pixout, int input_type, int flip, int size, int buff_size, int window_type, double pi_alpha, int overlap, int clip,
                            double fsclip_low, double fsclip_high, int n_pixels, int n_stitch, int max_writeahead)
{
    if (!h) return QC_EINVAL;
    return h->a.set(n_pixout, input_type, flip, size, buff_size, window_type, pi_alpha, overlap, clip, fsclip_low, fsclip_high, n_pixels, n_stitch, max_writeahead);
}

int quisk_cuda_analyzer_set_detector_mode(qcAnalyzer *h, int pixout, int mode)
{   // SetDisplayDetectorMode, analyzer.c:1582-1591
    if (!h || pixout < 0 || pixout >= qc::AN_MAX_PIXOUTS || mode < 0 || mode > 4) return QC_EINVAL;
    QC_CUDA(cudaDeviceSynchronize());
    h->a.det_type[pixout] = mode;
    return QC_OK;
}

int quisk_cuda_analyzer_set_average_mode(qcAnalyzer *h, int pixout, int mode)
{
    if (!h || pixout < 0 || pixout >= qc::AN_MAX_PIXOUTS || mode < -1 || mode > 3) return QC_EINVAL;
    return h->a.set_average_mode(pixout, mode);
}

int quisk_cuda_analyzer_set_num_average(qcAnalyzer *h, int pixout, int num)
{   // SetDisplayNumAverage, analyzer.c:1626-1638
    if (!h || pixout < 0 || pixout >= qc::AN_MAX_PIXOUTS || num < 1 || num > qc::AN_MAX_AVERAGE) return QC_EINVAL;
    if (h->a.num_average[pixout] != num) {
        QC_CUDA(cudaDeviceSynchronize());
        h->a.num_average[pixout] = num; h->a.avail_frames[pixout] = 0; h->a.av_in_idx[pixout] = 0; h->a.av_out_idx[pixout] = 0;
    }
    return QC_OK;
}

int quisk_cuda_analyzer_set_av_backmult(qcAnalyzer *h, int pixout, double mult)
{   // SetDisplayAvBackmult, analyzer.c:1641-1650
    if (!h || pixout < 0 || pixout >= qc::AN_MAX_PIXOUTS) return QC_EINVAL;
    QC_CUDA(cudaDeviceSynchronize());
    h->a.av_backmult[pixout] = mult;
    return QC_OK;
}

int quisk_cuda_analyzer_set_sample_rate(qcAnalyzer *h, int rate)
{   // SetDisplaySampleRate + CalcBandwidthNormalization, analyzer.c:919-924, 1653-1663
    if (!h || rate <= 0) return QC_EINVAL;
    QC_CUDA(cudaDeviceSynchronize());
    h->a.sample_rate = rate;
    if (h->a.size > 0) h->a.norm_oneHz = 10.0 * qc::an_host_mlog10(1.0 / ((double)rate / (double)h->a.size));
    return QC_OK;
}

int quisk_cuda_analyzer_set_norm_onehz(qcAnalyzer *h, int pixout, int norm)
{   // SetDisplayNormOneHz, analyzer.c:1666-1675
    if (!h || pixout < 0 || pixout >= qc::AN_MAX_PIXOUTS) return QC_EINVAL;
    QC_CUDA(cudaDeviceSynchronize());
    h->a.normalize[pixout] = norm ? 1 : 0;
    return QC_OK;
}

int quisk_cuda_analyzer_spectrum0(qcAnalyzer *h, int ss, const void *d_samples, long stride, void *stream)
{
    if (!h || !d_samples) return QC_EINVAL;
    return h->a.spectrum0(ss, (const double2 *)d_samples, stride, (cudaStream_t)stream);
}

int quisk_cuda_analyzer_get_pixels(qcAnalyzer *h, int pixout, float *h_pixels, int *flag)
{   // GetPixels, analyzer.c:1315-1334: the newest line of every display, flag = 1 if there is one that has not been fetched yet
    if (!h || pixout < 0 || pixout >= h->a.num_pixout || !h_pixels) return QC_EINVAL;
    qc::Analyzer &a = h->a;
    if (a.frames_done == a.frames_read[pixout]) { if (flag) *flag = 0; return QC_OK; }
    QC_CUDA(cudaDeviceSynchronize());
    QC_CUDA(cudaMemcpy(h_pixels, a.d_pix[pixout], (size_t)a.D * a.num_pixels * sizeof(float), cudaMemcpyDeviceToHost));
    a.frames_read[pixout] = a.frames_done;
    if (flag) *flag = 1;
    return QC_OK;
}

double quisk_cuda_analyzer_get_enb(qcAnalyzer *h) { return h ? 1.0 / h->a.inv_enb : 0.0; }     // GetDisplayENB, analyzer.c:1678-1686

}  // extern "C"
