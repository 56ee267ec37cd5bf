// quisk_b200/csrc/batch.cu -- quisk_cuda_batch_*: one filter.h filter instantiated for many
// independent channels, state resident in HBM (include/quisk_cuda.h section 2).
//
// State per object: a ping-pong pair of [C][H] history arrays (the last H inputs
// of every channel, oldest first) and ONE host-side phase integer (toggle /
// decim_index): all channels are fed the same number of samples per call, so the
// phase and every output count are known on the host without a device round trip.
#include "qc_common.cuh"
#include "batch.h"

namespace qc {

static const int OUT_CLIP = 66000 * 8 / 10;

static int elem_size(int kind)
{
    switch (kind) {
    case QC_D_DECIMATE: case QC_D_INTERPOLATE: case QC_D_INTERP2_HB45: return (int)sizeof(double);
    default: return (int)sizeof(cd);
    }
}

int BatchFilter::init(int kind_, int C_, const double *coefs, int n_taps, int interp_, int decim_)
{
    kind = kind_; C = C_; nTaps = n_taps; interp = interp_ < 1 ? 1 : interp_; decim = decim_ < 1 ? 1 : decim_;
    esize = elem_size(kind);
    phase = 0; cur = 0;
    std::vector<double> up;      // taps as uploaded
    switch (kind) {
    case QC_C_DECIM2_HB45: H = 44; break;
    case QC_C_INTERP2_HB45: case QC_D_INTERP2_HB45: H = 22; break;
    case QC_C_RXFILTER:
        // h[0] = filt[0], h[m] = filt[N-m] (quisk.c:1240-1255), I and Q taps interleaved per tap
        H = nTaps - 1;
        up.resize((size_t)2 * nTaps);
        for (int m = 0; m < nTaps; m++) {
            const int k = m == 0 ? 0 : nTaps - m;
            up[2 * m] = coefs[k];
            up[2 * m + 1] = coefs[nTaps + k];
        }
        break;
    case QC_D_RXFILTER:
        H = nTaps - 1;
        up.resize((size_t)nTaps);
        for (int m = 0; m < nTaps; m++) up[m] = coefs[m == 0 ? 0 : nTaps - m];
        break;
    case QC_C_CDECIMATE:
        H = nTaps - 1; up.assign(coefs, coefs + (size_t)2 * nTaps); break;
    default:
        H = nTaps - 1; up.assign(coefs, coefs + nTaps); break;
    }
    if (C <= 0 || (up.empty() && H != 44 && H != 22)) { set_error("batch_create: bad arguments"); return QC_EINVAL; }
    h_coef = up;
    if (!up.empty()) {
        QC_CUDA(cudaMalloc((void **)&d_coef, up.size() * sizeof(double)));
        QC_CUDA(cudaMemcpy(d_coef, up.data(), up.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    const size_t hb = (size_t)C * (H > 0 ? H : 1) * esize;
    for (int i = 0; i < 2; i++) {
        QC_CUDA(cudaMalloc((void **)&d_hist[i], hb));
        QC_CUDA(cudaMemset(d_hist[i], 0, hb));
    }
    return QC_OK;
}

void BatchFilter::release()
{
    if (d_coef) cudaFree(d_coef);
    for (int i = 0; i < 2; i++) if (d_hist[i]) cudaFree(d_hist[i]);
    d_coef = nullptr; d_hist[0] = d_hist[1] = nullptr;
}

long BatchFilter::count_full(int count) const
{
    switch (kind) {
    case QC_C_DECIM2_HB45: return (count + phase) / 2;
    case QC_C_DECIMATE: case QC_C_CDECIMATE: case QC_D_DECIMATE: return (count + phase) / decim;
    case QC_C_INTERPOLATE: case QC_D_INTERPOLATE: return (long)count * interp;
    case QC_C_INTERPDECIM: { long span = (long)count * interp - phase; return span > 0 ? (span + decim - 1) / decim : 0; }
    case QC_C_INTERP2_HB45: case QC_D_INTERP2_HB45: return 2L * count;
    default: return count;
    }
}

int BatchFilter::count_out(int count, int legacy_clip) const
{
    long n = count_full(count);
    if (legacy_clip) {
        if (kind == QC_C_INTERPOLATE || kind == QC_D_INTERPOLATE || kind == QC_C_INTERPDECIM)
            n = n < OUT_CLIP ? n : OUT_CLIP;
        else if (kind == QC_C_INTERP2_HB45 || kind == QC_D_INTERP2_HB45)
            n = n < OUT_CLIP + 2 ? n : OUT_CLIP + 2;
    }
    return (int)n;
}

int BatchFilter::run(const void *d_in, long in_stride, int count, void *d_out, long out_stride,
                     int *n_out, int legacy_clip, cudaStream_t stream)
{
    if (count < 0) { set_error("batch_run: negative count"); return QC_EINVAL; }
    const int nout = count_out(count, legacy_clip);
    if (n_out) *n_out = nout;
    if (count == 0) return QC_OK;
    PolyFirParams p;
    memset(&p, 0, sizeof(p));
    p.hist = d_hist[cur]; p.hist_out = d_hist[cur ^ 1]; p.H = H;
    p.in = d_in; p.in_stride = in_stride; p.n_in = count;
    p.out = d_out; p.out_stride = out_stride; p.n_out = nout;
    p.coef = d_coef; p.C = C; p.is_complex = esize == (int)sizeof(cd);
    p.K = nTaps; p.L = 1; p.M = 1; p.u0 = 0; p.gain = 1.0; p.tap_mode = TAP_REAL; p.order = 0; p.hb_mode = HB_NONE;
    const long full = count_full(count);
    switch (kind) {
    case QC_C_DECIM2_HB45:
        p.hb_mode = HB_DECIM; p.u0 = 1 - phase; phase = (phase + count) & 1; break;
    case QC_C_DECIMATE: case QC_D_DECIMATE:
        p.M = decim; p.u0 = decim - 1 - phase; phase = (phase + count) % decim; break;
    case QC_C_CDECIMATE:
        p.tap_mode = TAP_COMPLEX; p.M = decim; p.u0 = decim - 1 - phase; phase = (phase + count) % decim; break;
    case QC_C_INTERPOLATE: case QC_D_INTERPOLATE:
        p.K = nTaps / interp; p.L = interp; p.gain = (double)interp; break;
    case QC_C_INTERPDECIM:
        p.K = nTaps / interp; p.L = interp; p.M = decim; p.u0 = phase; p.gain = unit_gain ? 1.0 : (double)interp;
        phase = (int)(phase + full * decim - (long)count * interp); break;
    case QC_C_INTERP2_HB45: case QC_D_INTERP2_HB45:
        p.hb_mode = HB_INTERP; break;
    case QC_C_RXFILTER:
        p.tap_mode = TAP_SPLIT_IQ; p.order = 1; break;
    case QC_D_RXFILTER:
        p.order = 1; break;
    default:
        set_error("batch_run: unknown kind %d", kind); return QC_EINVAL;
    }
    if (p.hb_mode == HB_NONE && p.K < 1) {
        // no taps per phase (nTaps < interp): the reference emits zeros
        for (int c = 0; c < C; c++)
            QC_CUDA(cudaMemsetAsync((char *)d_out + (size_t)c * out_stride * esize, 0, (size_t)nout * esize, stream));
        p.K = 1; p.n_out = 0;        // still rolls the history forward
    }
    int rc = launch_polyfir(p, stream);
    if (rc != QC_OK) return rc;
    cur ^= 1;
    return QC_OK;
}

int BatchFilter::reset(cudaStream_t stream)
{
    const size_t hb = (size_t)C * (H > 0 ? H : 1) * esize;
    QC_CUDA(cudaMemsetAsync(d_hist[0], 0, hb, stream));
    QC_CUDA(cudaMemsetAsync(d_hist[1], 0, hb, stream));
    phase = 0; cur = 0;
    return QC_OK;
}

}  // namespace qc

struct qcBatchFilter { qc::BatchFilter f; };

extern "C" {

qcBatchFilter *quisk_cuda_batch_create(int kind, int n_channels, const double *coefs, int n_taps, int interp, int decim)
{
    if (qc::ensure_device() != QC_OK) return nullptr;
    qcBatchFilter *b = new qcBatchFilter();
    if (b->f.init(kind, n_channels, coefs, n_taps, interp, decim) != QC_OK) { b->f.release(); delete b; return nullptr; }
    return b;
}

void quisk_cuda_batch_destroy(qcBatchFilter *f) { if (f) { f->f.release(); delete f; } }

int quisk_cuda_batch_count_out(const qcBatchFilter *f, int count) { return f ? f->f.count_out(count, 0) : QC_EINVAL; }

int quisk_cuda_batch_run(qcBatchFilter *f, const void *d_in, long in_stride, int count, void *d_out, long out_stride,
                         int *n_out, int legacy_clip, void *stream)
{
    if (!f) { qc::set_error("batch_run: null filter"); return QC_EINVAL; }
    return f->f.run(d_in, in_stride, count, d_out, out_stride, n_out, legacy_clip, (cudaStream_t)stream);
}

int quisk_cuda_batch_reset(qcBatchFilter *f, void *stream)
{
    if (!f) return QC_EINVAL;
    return f->f.reset((cudaStream_t)stream);
}

}  // extern "C"
