// quisk_b200/csrc/filter_abi.cu -- the reference's filter.h entry points (filter.h:39-55),
// host-pointer / in-place / caller-owned state, arithmetic on the GPU.
//
// State lives where the reference keeps it: in the caller's struct (ring buffer
// + write pointer + decim_index, or the HB45 shift registers + toggle), updated
// on the host exactly as filter.c would leave it, so a struct can be handed back
// and forth between this library and the reference.  Per call the host builds
// the linear stream [history | block] from the ring, ships it with the taps in
// ONE pinned H2D copy, runs the exact polyphase kernel (polyfir.cu) and copies
// the outputs back over the caller's buffer.  The device holds no state for this
// ABI.  Calls are serialised by one mutex (the reference's callers are single
// threaded: the sound thread, quisk.c:4260-4266).
#include "qc_common.cuh"
#include <complex>

namespace qc {

static const int OUT_CLIP = 66000 * 8 / 10;      // SAMP_BUFFER_SIZE * 8 / 10, filter.c:158

struct LegacyCtx {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    char *h_pin = nullptr; size_t h_cap = 0;
    char *d_in = nullptr;  size_t din_cap = 0;
    char *d_out = nullptr; size_t dout_cap = 0;

    int reserve(size_t in_bytes, size_t out_bytes)
    {
        if (!stream) QC_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        size_t hneed = in_bytes > out_bytes ? in_bytes : out_bytes;
        if (hneed > h_cap) {
            if (h_pin) cudaFreeHost(h_pin);
            h_cap = hneed * 2 + 4096; h_pin = nullptr;
            QC_CUDA(cudaMallocHost((void **)&h_pin, h_cap));
        }
        if (in_bytes > din_cap) {
            if (d_in) cudaFree(d_in);
            din_cap = in_bytes * 2 + 4096; d_in = nullptr;
            QC_CUDA(cudaMalloc((void **)&d_in, din_cap));
        }
        if (out_bytes > dout_cap) {
            if (d_out) cudaFree(d_out);
            dout_cap = out_bytes * 2 + 4096; d_out = nullptr;
            QC_CUDA(cudaMalloc((void **)&d_out, dout_cap));
        }
        return QC_OK;
    }
};

static LegacyCtx g_ctx;

// Run one stateless pass.  `xhist` = H history elements (oldest first), `x` = count new ones.
template <typename T>
static int run_pass(const char *fn, PolyFirParams p, const void *coefs, size_t coef_bytes,
                    const T *xhist, int H, const T *x, int count, T *out, int n_out)
{
    if (ensure_device() != QC_OK) die_no_device(fn);
    const size_t coef_pad = (coef_bytes + 15) & ~(size_t)15;
    const size_t in_bytes = coef_pad + (size_t)(H + count) * sizeof(T);
    const size_t out_bytes = (size_t)(n_out > 0 ? n_out : 1) * sizeof(T);
    int rc = g_ctx.reserve(in_bytes, out_bytes);
    if (rc != QC_OK) die_no_device(fn);
    if (coef_bytes) memcpy(g_ctx.h_pin, coefs, coef_bytes);
    T *hx = reinterpret_cast<T *>(g_ctx.h_pin + coef_pad);
    if (H) memcpy(hx, xhist, (size_t)H * sizeof(T));
    if (count) memcpy(hx + H, x, (size_t)count * sizeof(T));
    if (cudaMemcpyAsync(g_ctx.d_in, g_ctx.h_pin, in_bytes, cudaMemcpyHostToDevice, g_ctx.stream) != cudaSuccess)
        die_no_device(fn);
    T *dx = reinterpret_cast<T *>(g_ctx.d_in + coef_pad);
    p.coef = reinterpret_cast<const double *>(g_ctx.d_in);
    p.hist = dx; p.H = H;
    p.in = dx + H; p.in_stride = 0; p.n_in = count;
    p.out = g_ctx.d_out; p.out_stride = 0; p.n_out = n_out;
    p.hist_out = nullptr;
    p.C = 1;
    p.is_complex = sizeof(T) == sizeof(cd);
    if (launch_polyfir(p, g_ctx.stream) != QC_OK) die_no_device(fn);
    if (n_out > 0 &&
        cudaMemcpyAsync(g_ctx.h_pin, g_ctx.d_out, (size_t)n_out * sizeof(T), cudaMemcpyDeviceToHost, g_ctx.stream) != cudaSuccess)
        die_no_device(fn);
    if (check(cudaStreamSynchronize(g_ctx.stream), fn, __FILE__, __LINE__) != QC_OK) die_no_device(fn);
    if (n_out > 0) memcpy(out, g_ctx.h_pin, (size_t)n_out * sizeof(T));
    return QC_OK;
}

// Ring -> linear history (oldest first) of the nTaps-1 samples before the write pointer.
template <typename T>
static void ring_history(const T *ring, const T *wr, int nTaps, std::vector<T> &hist)
{
    const int H = nTaps - 1;
    hist.resize(H > 0 ? H : 0);
    const int p0 = (int)(wr - ring);
    for (int j = 1; j <= H; j++) {
        int pos = p0 - j;
        if (pos < 0) pos += nTaps;
        hist[H - j] = ring[pos];
    }
}

// What the reference's `*pt = s; if (++pt >= ring + nTaps) pt = ring;` loop leaves behind.
template <typename T>
static T *ring_push(T *ring, T *wr, int nTaps, const T *x, int count)
{
    int p = (int)(wr - ring);
    int start = 0;
    if (count > nTaps) {            // only the last nTaps survive; keep positions right
        start = count - nTaps;
        p = (int)(((long)p + start) % nTaps);
    }
    for (int i = start; i < count; i++) {
        ring[p] = x[i];
        if (++p >= nTaps) p = 0;
    }
    return ring + p;
}

static PolyFirParams base_params(int K, int L, int M, long u0, double gain, int tap_mode, int order = 0)
{
    PolyFirParams p;
    memset(&p, 0, sizeof(p));
    p.K = K; p.L = L; p.M = M; p.u0 = u0; p.gain = gain; p.tap_mode = tap_mode; p.order = order; p.hb_mode = HB_NONE;
    return p;
}

// ---- decimators / plain filters ------------------------------------------------
template <typename T, typename F>
static int decimate_impl(const char *fn, T *samples, int count, F *filter, T *ring, T **wr, int decim,
                         const void *coefs, int tap_mode, bool use_index = true)
{
    std::lock_guard<std::mutex> g(g_ctx.mu);
    if (count <= 0) return 0;
    const int nTaps = filter->nTaps;
    // filter.c:213 tests `++decim_index >= decim`: an index left at or above `decim` by another entry point or ratio
    // (filtDecim5S is shared by cDecimate(5) and cInterpDecim(4, 5), quisk.c:1825,1837) emits on the first sample and
    // restarts from 0, i.e. it behaves like decim - 1.  quisk_dFilter / quisk_dD_out (filter.c:326-370) never look at
    // the index: use_index = false leaves it alone.
    int d0 = use_index ? filter->decim_index : 0;
    if (d0 > decim - 1) d0 = decim - 1;
    if (d0 < 0) d0 = 0;
    const int n_out = (count + d0) / decim;
    std::vector<T> hist;
    ring_history(ring, *wr, nTaps, hist);
    PolyFirParams p = base_params(nTaps, 1, decim, decim - 1 - d0, 1.0, tap_mode);
    const size_t cb = (size_t)nTaps * (tap_mode == TAP_REAL ? sizeof(double) : 2 * sizeof(double));
    std::vector<T> in(samples, samples + count);            // outputs overwrite the caller's buffer
    run_pass<T>(fn, p, coefs, cb, hist.data(), nTaps - 1, in.data(), count, samples, n_out);
    *wr = ring_push(ring, *wr, nTaps, in.data(), count);
    if (use_index) filter->decim_index = (d0 + count) % decim;
    return n_out;
}

// ---- interpolators ---------------------------------------------------------------
template <typename T, typename F>
static int interp_impl(const char *fn, T *samples, int count, F *filter, T *ring, T **wr, int interp, int decim, bool is_interpdecim)
{
    std::lock_guard<std::mutex> g(g_ctx.mu);
    if (count <= 0) return 0;
    const int nTaps = filter->nTaps;
    const int K = nTaps / interp;                           // filter.c:153 -- integer quotient
    long u0 = 0; int M = 1; long n_full;
    if (is_interpdecim) {
        u0 = filter->decim_index; M = decim;
        const long span = (long)count * interp - u0;
        n_full = span > 0 ? (span + M - 1) / M : 0;
        filter->decim_index = (int)(u0 + n_full * M - (long)count * interp);
    } else {
        n_full = (long)count * interp;
    }
    const int n_out = (int)(n_full < OUT_CLIP ? n_full : OUT_CLIP);
    std::vector<T> hist;
    ring_history(ring, *wr, nTaps, hist);
    std::vector<T> in(samples, samples + count);
    if (K > 0) {
        PolyFirParams p = base_params(K, interp, M, u0, (double)interp, TAP_REAL);
        run_pass<T>(fn, p, filter->dCoefs, (size_t)nTaps * sizeof(double), hist.data(), nTaps - 1, in.data(), count, samples, n_out);
    } else {
        memset(samples, 0, (size_t)n_out * sizeof(T));      // no taps per phase: the reference writes 0 * interp
    }
    *wr = ring_push(ring, *wr, nTaps, in.data(), count);
    return n_out;
}

}  // namespace qc

using namespace qc;

template <typename T, typename F>
static int interp2hb45_impl(const char *fn, T *buf, int count, F *filter, T *samples)
{   // filter.c:420-488
    std::lock_guard<std::mutex> g(g_ctx.mu);
    if (count <= 0) return 0;
    const int H = 22;
    T hist[H];
    for (int j = 1; j <= H; j++) hist[H - j] = samples[j - 1];      // samples[k] = X[-1-k]
    // pairs are written while nOut <= 52800 at the top of the iteration (filter.c:444,479)
    const int max_pairs = (OUT_CLIP + 2) / 2;
    const int n_pairs = count < max_pairs ? count : max_pairs;
    PolyFirParams p = base_params(1, 1, 1, 0, 1.0, TAP_REAL);
    p.hb_mode = HB_INTERP;
    std::vector<T> in(buf, buf + count);
    run_pass<T>(fn, p, nullptr, 0, hist, H, in.data(), count, buf, 2 * n_pairs);
    const int tail = count < 22 ? count : 22;
    for (int i = count - tail; i < count; i++) { memmove(samples + 1, samples, sizeof(T) * 21); samples[0] = in[i]; }
    return 2 * n_pairs;
}


extern "C" {

void quisk_filt_cInit(struct quisk_cFilter *filter, double *coefs, int taps)
{   // filter.c:9-20 -- pure host bookkeeping
    filter->dCoefs = coefs;
    filter->cpxCoefs = NULL;
    filter->cSamples = (quisk_cd *)calloc((size_t)taps, sizeof(quisk_cd));
    filter->ptcSamp = filter->cSamples;
    filter->nTaps = taps;
    filter->decim_index = 0;
    filter->cBuf = NULL;
    filter->nBuf = 0;
}

void quisk_filt_dInit(struct quisk_dFilter *filter, double *coefs, int taps)
{   // filter.c:22-33
    filter->dCoefs = coefs;
    filter->cpxCoefs = NULL;
    filter->dSamples = (double *)calloc((size_t)taps, sizeof(double));
    filter->ptdSamp = filter->dSamples;
    filter->nTaps = taps;
    filter->decim_index = 0;
    filter->dBuf = NULL;
    filter->nBuf = 0;
}

void quisk_filt_differInit(struct quisk_dFilter *filter, int taps)
{   // filter.c:35-56 -- classic differentiator taps (-1)^k / k; the reference also printf()s them
    filter->dCoefs = (double *)malloc((size_t)taps * sizeof(double));
    for (int k = -(taps - 1) / 2; k <= (taps - 1) / 2; k++) {
        const int j = (taps - 1) / 2 + k;
        filter->dCoefs[j] = (k == 0) ? 0.0 : pow(-1, k) / k;
    }
    filter->cpxCoefs = NULL;
    filter->dSamples = (double *)calloc((size_t)taps, sizeof(double));
    filter->ptdSamp = filter->dSamples;
    filter->nTaps = taps;
    filter->decim_index = 0;
    filter->dBuf = NULL;
    filter->nBuf = 0;
}

void quisk_filt_tune(struct quisk_dFilter *filter, double freq, int ssb_upper)
{   // filter.c:58-81 -- host-side coefficient design: cexp(j 2 pi f (i - D)) * dCoefs[i]
    if (!filter->cpxCoefs)
        filter->cpxCoefs = (quisk_cd *)malloc((size_t)filter->nTaps * sizeof(quisk_cd));
    const std::complex<double> tune = std::complex<double>(0.0, 1.0) * 2.0 * M_PI * freq;
    const double D = (filter->nTaps - 1.0) / 2.0;
    for (int i = 0; i < filter->nTaps; i++) {
        const std::complex<double> coef = std::exp(tune * (i - D)) * filter->dCoefs[i];
        if (ssb_upper) { filter->cpxCoefs[i].re = coef.real(); filter->cpxCoefs[i].im = coef.imag(); }
        else { filter->cpxCoefs[i].re = coef.imag(); filter->cpxCoefs[i].im = coef.real(); }
    }
}

int quisk_cDecimate(quisk_cd *cSamples, int count, struct quisk_cFilter *filter, int decim)
{
    return decimate_impl<cd>("quisk_cDecimate", (cd *)cSamples, count, filter, (cd *)filter->cSamples,
                             (cd **)&filter->ptcSamp, decim, filter->dCoefs, TAP_REAL);
}

int quisk_cFilter(quisk_cd *cSamples, int count, struct quisk_cFilter *filter)
{   // filter.c:372-375
    return quisk_cDecimate(cSamples, count, filter, 1);
}

int quisk_cCDecimate(quisk_cd *cSamples, int count, struct quisk_cFilter *filter, int decim)
{
    return decimate_impl<cd>("quisk_cCDecimate", (cd *)cSamples, count, filter, (cd *)filter->cSamples,
                             (cd **)&filter->ptcSamp, decim, filter->cpxCoefs, TAP_COMPLEX);
}

int quisk_dDecimate(double *dSamples, int count, struct quisk_dFilter *filter, int decim)
{
    return decimate_impl<double>("quisk_dDecimate", dSamples, count, filter, filter->dSamples,
                                 &filter->ptdSamp, decim, filter->dCoefs, TAP_REAL);
}

int quisk_dFilter(double *dSamples, int count, struct quisk_dFilter *filter)
{   // filter.c:347-370
    return decimate_impl<double>("quisk_dFilter", dSamples, count, filter, filter->dSamples,
                                 &filter->ptdSamp, 1, filter->dCoefs, TAP_REAL, false);
}

double quisk_dD_out(double samp, struct quisk_dFilter *filter)
{   // filter.c:326-345: one sample through the same dot product (a launch per sample:
    // callers that care about speed use quisk_dFilter on a block)
    double s = samp;
    decimate_impl<double>("quisk_dD_out", &s, 1, filter, filter->dSamples, &filter->ptdSamp, 1, filter->dCoefs, TAP_REAL, false);
    return s;
}

quisk_cd quisk_dC_out(double sample, struct quisk_dFilter *filter)
{   // filter.c:83-104: real sample ring against the tuned complex taps.  The ring is
    // promoted to complex with zero imaginary parts (x*c - 0*d rounds to x*c).
    std::lock_guard<std::mutex> g(g_ctx.mu);
    const int nTaps = filter->nTaps;
    std::vector<double> hist;
    ring_history(filter->dSamples, filter->ptdSamp, nTaps, hist);
    std::vector<cd> ch(hist.size());
    for (size_t i = 0; i < hist.size(); i++) ch[i] = make_double2(hist[i], 0.0);
    cd in = make_double2(sample, 0.0), out = make_double2(0.0, 0.0);
    PolyFirParams p = base_params(nTaps, 1, 1, 0, 1.0, TAP_COMPLEX);
    run_pass<cd>("quisk_dC_out", p, filter->cpxCoefs, (size_t)nTaps * 2 * sizeof(double), ch.data(), nTaps - 1, &in, 1, &out, 1);
    filter->ptdSamp = ring_push(filter->dSamples, filter->ptdSamp, nTaps, &sample, 1);
    quisk_cd r; r.re = out.x; r.im = out.y;
    return r;
}

int quisk_cInterpolate(quisk_cd *cSamples, int count, struct quisk_cFilter *filter, int interp)
{
    return interp_impl<cd>("quisk_cInterpolate", (cd *)cSamples, count, filter, (cd *)filter->cSamples,
                           (cd **)&filter->ptcSamp, interp, 1, false);
}

int quisk_dInterpolate(double *dSamples, int count, struct quisk_dFilter *filter, int interp)
{
    return interp_impl<double>("quisk_dInterpolate", dSamples, count, filter, filter->dSamples,
                               &filter->ptdSamp, interp, 1, false);
}

int quisk_cInterpDecim(quisk_cd *cSamples, int count, struct quisk_cFilter *filter, int interp, int decim)
{
    return interp_impl<cd>("quisk_cInterpDecim", (cd *)cSamples, count, filter, (cd *)filter->cSamples,
                           (cd **)&filter->ptcSamp, interp, decim, true);
}

int quisk_cDecim2HB45(quisk_cd *cSamples, int count, struct quisk_cHB45Filter *filter)
{   // filter.c:377-417
    std::lock_guard<std::mutex> g(g_ctx.mu);
    if (count <= 0) return 0;
    cd *samples = (cd *)filter->samples, *center = (cd *)filter->center;
    const int t = filter->toggle ? 1 : 0;
    const int H = 44;
    // Linear history X[-1], X[-2], ...: with toggle == 0 the last input went to samples[0]
    // (it produced an output), with toggle == 1 it went to center[0].
    cd hist[H];
    for (int j = 1; j <= H; j++) {
        const int k = (j - 1) / 2;
        const bool to_samples = t == 0 ? (j & 1) : !(j & 1);
        cd v = make_double2(0.0, 0.0);
        if (to_samples) { if (k < 22) v = samples[k]; }
        else { if (k < 11) v = center[k]; }
        hist[H - j] = v;
    }
    const int n_out = (count + t) / 2;
    PolyFirParams p = base_params(1, 1, 2, 1 - t, 1.0, TAP_REAL);
    p.hb_mode = HB_DECIM;
    std::vector<cd> in((cd *)cSamples, (cd *)cSamples + count);
    run_pass<cd>("quisk_cDecim2HB45", p, nullptr, 0, hist, H, in.data(), count, (cd *)cSamples, n_out);
    // leave the shift registers as filter.c:391-400 would
    const int tail = count < 64 ? count : 64;
    int tog = (t + (count - tail)) & 1;
    for (int i = count - tail; i < count; i++) {
        if (tog == 0) { tog = 1; memmove(center + 1, center, sizeof(cd) * 10); center[0] = in[i]; }
        else { tog = 0; memmove(samples + 1, samples, sizeof(cd) * 21); samples[0] = in[i]; }
    }
    filter->toggle = tog;
    return n_out;
}

int quisk_dInterp2HB45(double *dsamples, int count, struct quisk_dHB45Filter *filter)
{
    return interp2hb45_impl<double>("quisk_dInterp2HB45", dsamples, count, filter, filter->samples);
}

int quisk_cInterp2HB45(quisk_cd *cSamples, int count, struct quisk_cHB45Filter *filter)
{
    return interp2hb45_impl<cd>("quisk_cInterp2HB45", (cd *)cSamples, count, filter, (cd *)filter->samples);
}

}  // extern "C"
