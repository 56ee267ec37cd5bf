// quisk_b200/csrc/wdsp_fircore.cu -- batched `fircore` (wdsp/firmin.c:290-430) and the rational
// resampler (wdsp/resample.c:121-157).
//
// fircore is a uniformly partitioned overlap-save FIR: block `size`, taps `nc`, nfor = nc/size
// partitions, FFT length 2*size.  Per block and channel (xfircore, firmin.c:409-430):
//   fftin = [previous block | new block]  ->  forward FFT  ->  fftout[buffidx]
//   accum = sum_j fftout[(buffidx - j) & mask] (.) fmask[cset][j]          (j = 0 .. nfor-1, that order)
//   backward FFT (unnormalised: callers bake 1/(2 size) into the impulse)  ->  out = first `size` samples
// GPU mapping: one CTA per channel per block.  The forward FFT, the partition MAC and the inverse FFT
// run back to back on one shared-memory buffer (fft_device.cuh); each thread owns 16 fixed bins, so the
// newest spectrum goes from registers straight into the MAC, and only the nfor-1 older spectra come
// from the frequency-domain delay line in global memory (128 KiB per channel at size 1024 / nc 4096:
// L2 resident for hundreds of channels).  The masks are shared by all channels and are built on the
// device with the same FFT (calc_fircore: segment j right-justified in a 2*size buffer, firmin.c:322-346).
// Two mask sets with a `cset` switch give the reference's glitch-free retune (setUpdate_fircore).
#include "fft_device.cuh"
#include "batch.h"
#include "wdsp_internal.h"

namespace qc {


// BPT = radix-16 butterflies per thread in the transforms: 1 whenever 2 size <= 4096 (every RXA configuration here),
// which halves the registers and doubles the resident CTAs; 2 only for the 8192-point case.
template <int BPT>
__global__ void __launch_bounds__(256) fircore_kernel(const cd *in, long in_stride, cd *out, long out_stride,
                                                       int size, int nfor, int buffidx,
                                                       cd *prev /*[C][size]*/, cd *fdl /*[C][nfor][2 size]*/,
                                                       const cd *fmask /*[nfor][2 size]*/, const cd *tw)
{
    extern __shared__ double smem_raw[];
    const int n2 = 2 * size;
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *s = twl + fft_tw_entries(n2);
    fft_stage_twiddles(twl, tw, n2);
    const int c = blockIdx.x;
    const int lane = threadIdx.x, lanes = blockDim.x;
    const cd *x = in + (size_t)c * in_stride;
    cd *pv = prev + (size_t)c * size;
    // fftin = [prev | new]; the new block becomes prev (firmin.c:411, 429)
    for (int i = lane; i < size; i += lanes) {
        const cd v = x[i];
        s[fsw(i)] = pv[i];
        s[fsw(size + i)] = v;
        pv[i] = v;
    }
    __syncthreads();
    fft_smem<BPT>(s, n2, twl, -1, lane, lanes);
    // partition MAC; bins lane, lane + lanes, ... are this thread's alone, so each accumulator replaces its
    // bin in shared memory without a barrier
    cd *fd = fdl + (size_t)c * nfor * n2;
    const int mask = nfor - 1;
    for (int i = lane; i < n2; i += lanes) {
        const cd X = s[fsw(i)];
        fd[(size_t)buffidx * n2 + i] = X;
        const cd m0 = fmask[i];
        cd acc = make_double2(X.x * m0.x - X.y * m0.y, X.x * m0.y + X.y * m0.x);
        int k = buffidx;
        for (int j = 1; j < nfor; j++) {
            k = (k + mask) & mask;
            const cd Y = fd[(size_t)k * n2 + i], m = fmask[(size_t)j * n2 + i];
            acc.x += Y.x * m.x - Y.y * m.y;
            acc.y += Y.x * m.y + Y.y * m.x;
        }
        s[fsw(i)] = acc;
    }
    __syncthreads();
    fft_smem<BPT>(s, n2, twl, +1, lane, lanes);
    cd *y = out + (size_t)c * out_stride;
    for (int i = lane; i < size; i += lanes) y[i] = s[fsw(i)];
}

int FirCore::init(int C_, int size_, int nc_, int mp_, const double *impulse)
{
    C = C_; size = size_; nc = nc_;
    mp = mp_ ? 1 : 0;
    if (C <= 0 || size < 4 || (size & (size - 1)) || fft_log2(2 * size) < 0 || nc < size || nc % size) {
        set_error("fircore: size must be a power of two in [4, 4096] and nc a multiple of it (size %d, nc %d)", size, nc);
        return QC_EINVAL;
    }
    nfor = nc / size;
    if (nfor & (nfor - 1)) { set_error("fircore: nc/size must be a power of two (got %d)", nfor); return QC_EINVAL; }
    n2 = 2 * size;
    tw = fft_twiddles(n2);
    if (!tw) { set_error("fircore: twiddle table allocation failed"); return QC_ENOMEM; }
    QC_CUDA(cudaMalloc((void **)&d_prev, (size_t)C * size * sizeof(cd)));
    QC_CUDA(cudaMalloc((void **)&d_fdl, (size_t)C * nfor * n2 * sizeof(cd)));
    for (int i = 0; i < 2; i++) QC_CUDA(cudaMalloc((void **)&d_mask[i], (size_t)nfor * n2 * sizeof(cd)));
    QC_CUDA(cudaMalloc((void **)&d_gen, (size_t)nfor * n2 * sizeof(cd)));
    int rc = flush(); if (rc != QC_OK) return rc;
    cset = 0; masks_ready = 0;
    if (!impulse) return QC_OK;             // the caller supplies the mask generator itself (set_gen: xbps)
    return set_impulse(impulse, 1);         // create_fircore: calc_fircore(a, 1)
}

void FirCore::release()
{
    if (d_prev) cudaFree(d_prev); if (d_fdl) cudaFree(d_fdl); if (d_gen) cudaFree(d_gen);
    for (int i = 0; i < 2; i++) if (d_mask[i]) cudaFree(d_mask[i]);
    d_prev = d_fdl = d_gen = nullptr; d_mask[0] = d_mask[1] = nullptr;
}

int FirCore::flush()
{   // flush_fircore, firmin.c:399-407
    QC_CUDA(cudaMemset(d_prev, 0, (size_t)C * size * sizeof(cd)));
    QC_CUDA(cudaMemset(d_fdl, 0, (size_t)C * nfor * n2 * sizeof(cd)));
    buffidx = 0;
    return QC_OK;
}

int FirCore::set_impulse(const double *impulse, int update)
{   // calc_fircore: masks into the set that is NOT in use
    if (impulse != h_impulse.data()) h_impulse.assign(impulse, impulse + (size_t)2 * nc);
    impulse = h_impulse.data();
    std::vector<double> mpi;
    if (mp) {                               // calc_fircore, firmin.c:327-328
        mpi.resize((size_t)2 * nc);
        if (quisk_cuda_mp_imp(nc, impulse, mpi.data(), 16, 0) != QC_OK) { set_error("fircore: mp_imp needs nc * 16 to be a power of two (nc %d)", nc); return QC_EINVAL; }
        impulse = mpi.data();
    }
    std::vector<cd> gen((size_t)nfor * n2, make_double2(0.0, 0.0));
    for (int j = 0; j < nfor; j++)
        for (int i = 0; i < size; i++)
            gen[(size_t)j * n2 + size + i] = make_double2(impulse[2 * ((size_t)size * j + i)], impulse[2 * ((size_t)size * j + i) + 1]);
    return set_gen(gen.data(), update);
}

int FirCore::set_gen(const cd *gen, int update)
{   // masks = forward transforms of nfor time-domain generator rows of 2*size samples
    QC_CUDA(cudaMemcpy(d_gen, gen, (size_t)nfor * n2 * sizeof(cd), cudaMemcpyHostToDevice));
    int rc = quisk_cuda_fft_batch(d_gen, d_mask[1 - cset], n2, nfor, -1, nullptr);
    if (rc != QC_OK) return rc;
    QC_CUDA(cudaDeviceSynchronize());
    masks_ready = 1;
    if (update) return this->update();
    return QC_OK;
}

int FirCore::set_mp(int mp_)
{
    mp = mp_ ? 1 : 0;
    return set_impulse(h_impulse.data(), 1);
}

int FirCore::update()
{   // setUpdate_fircore, firmin.c:475-484
    if (masks_ready) { cset = 1 - cset; masks_ready = 0; }
    return QC_OK;
}

int FirCore::run(const void *d_in, long in_stride, void *d_out, long out_stride, cudaStream_t s)
{
    const int lanes = fft_threads(n2);
    const size_t sh = ((size_t)n2 + fft_tw_entries(n2)) * sizeof(cd);
    if (n2 > 4096) {
        QC_CUDA(cudaFuncSetAttribute(fircore_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        fircore_kernel<2><<<C, lanes, sh, s>>>((const cd *)d_in, in_stride, (cd *)d_out, out_stride, size, nfor, buffidx,
                                               d_prev, d_fdl, d_mask[cset], tw);
    } else {
        if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(fircore_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        fircore_kernel<1><<<C, lanes, sh, s>>>((const cd *)d_in, in_stride, (cd *)d_out, out_stride, size, nfor, buffidx,
                                               d_prev, d_fdl, d_mask[cset], tw);
    }
    count_launch();
    QC_CUDA_LAUNCH();
    buffidx = (buffidx + 1) & (nfor - 1);
    return QC_OK;
}

// ---- resampler: xresample is the streaming polyphase FIR with u = phnum + m*M (resample.c:134-150) ----
int Resampler::init(int C_, int in_rate, int out_rate, double fc, int ncoef_in, double gain) { return init_band(C_, in_rate, out_rate, -1.0, fc, ncoef_in, gain); }

int Resampler::init_band(int C_, int in_rate, int out_rate, double fc_low, double fc, int ncoef_in, double gain)
{
    C = C_;
    int rc = quisk_cuda_resample_design_band(in_rate, out_rate, fc_low, fc, ncoef_in, gain, &L, &M, &ncoef, nullptr, 0);
    if (rc != QC_OK) { set_error("resample: bad rates %d -> %d", in_rate, out_rate); return rc; }
    std::vector<double> h((size_t)ncoef);
    quisk_cuda_resample_design_band(in_rate, out_rate, fc_low, fc, ncoef_in, gain, nullptr, nullptr, nullptr, h.data(), ncoef);
    // QC_C_INTERPDECIM indexes coef[ph + k*L] with K = nTaps / L taps per phase and applies gain L;
    // the resampler's prototype already carries gain*L (resample.c:66), so divide it back out exactly
    // by running the filter object with its own gain switched off.
    f = new BatchFilter();
    rc = f->init(QC_C_INTERPDECIM, C, h.data(), ncoef, L, M);
    if (rc != QC_OK) return rc;
    f->unit_gain = true;
    return QC_OK;
}

void Resampler::release() { if (f) { f->release(); delete f; f = nullptr; } }

}  // namespace qc

struct qcFircore { qc::FirCore f; };
struct qcResample { qc::Resampler r; };

extern "C" {

qcFircore *quisk_cuda_fircore_create(int n_channels, int size, int nc, int mp, const double *impulse)
{
    if (qc::ensure_device() != QC_OK) return nullptr;
    if (!impulse) { qc::set_error("fircore_create: null impulse"); return nullptr; }
    qcFircore *p = new qcFircore();
    if (p->f.init(n_channels, size, nc, mp, impulse) != QC_OK) { p->f.release(); delete p; return nullptr; }
    return p;
}
void quisk_cuda_fircore_destroy(qcFircore *f) { if (f) { f->f.release(); delete f; } }
int quisk_cuda_fircore_run(qcFircore *f, const void *d_in, long in_stride, void *d_out, long out_stride, void *stream)
{ return f ? f->f.run(d_in, in_stride, d_out, out_stride, (cudaStream_t)stream) : QC_EINVAL; }
int quisk_cuda_fircore_set_impulse(qcFircore *f, const double *impulse, int update) { return f && impulse ? f->f.set_impulse(impulse, update) : QC_EINVAL; }
int quisk_cuda_fircore_set_mp(qcFircore *f, int mp) { return f ? f->f.set_mp(mp) : QC_EINVAL; }
int quisk_cuda_fircore_update(qcFircore *f) { return f ? f->f.update() : QC_EINVAL; }
int quisk_cuda_fircore_flush(qcFircore *f) { return f ? f->f.flush() : QC_EINVAL; }

// ---- the three variants the reference defines next to fircore and never instantiates ----
// firopt (firmin.c:127-251): partitioned overlap-save with ONE mask set = fircore's arithmetic with the taps of calc_firopt
qcFircore *quisk_cuda_firopt_create(int n_channels, int size, int nc, double f_low, double f_high, int samplerate, int wintype, double gain)
{
    if (nc <= 0) { qc::set_error("firopt_create: nc"); return nullptr; }
    std::vector<double> imp((size_t)2 * nc);
    if (quisk_cuda_fir_bandpass(nc, f_low, f_high, (double)samplerate, wintype, 1, gain, imp.data()) != QC_OK) return nullptr;
    return quisk_cuda_fircore_create(n_channels, size, nc, 0, imp.data());
}

// bps (bandpass.c:35-105): one overlap-save block; size + 1 taps stored right-justified from index size - 1 (fftcv_mults,
// fir.c:29-42), `gain` applied to the spectrum at run time -- folded into the taps here
qcFircore *quisk_cuda_bps_create(int n_channels, int size, double f_low, double f_high, int samplerate, int wintype, double gain)
{
    if (qc::ensure_device() != QC_OK) return nullptr;
    if (size < 4) { qc::set_error("bps_create: size"); return nullptr; }
    std::vector<double> imp((size_t)2 * (size + 1));
    if (quisk_cuda_fir_bandpass(size + 1, f_low, f_high, (double)samplerate, wintype, 1, 1.0 / (double)(2 * size), imp.data()) != QC_OK) return nullptr;
    qcFircore *f = new qcFircore();
    if (f->f.init(n_channels, size, size, 0, nullptr) != QC_OK) { f->f.release(); delete f; return nullptr; }
    std::vector<double2> gen((size_t)2 * size, make_double2(0.0, 0.0));
    for (int i = 0; i <= size; i++) gen[size - 1 + i] = make_double2(gain * imp[2 * i], gain * imp[2 * i + 1]);
    if (f->f.set_gen(gen.data(), 1) != QC_OK) { f->f.release(); delete f; return nullptr; }
    return f;
}

// firmin (firmin.c:35-99): time-domain complex-tap FIR over a ring of nc samples, newest sample first = quisk_cCDecimate's
// loop with decim 1; the taps are calc_firmin's.  Run it with quisk_cuda_batch_run, flush_firmin = quisk_cuda_batch_reset.
qcBatchFilter *quisk_cuda_firmin_create(int n_channels, int nc, double f_low, double f_high, int samplerate, int wintype, double gain)
{
    if (nc <= 0 || (nc & (nc - 1))) { qc::set_error("firmin_create: nc must be a power of two (the reference masks its ring index with nc - 1)"); return nullptr; }
    std::vector<double> imp((size_t)2 * nc);
    if (quisk_cuda_fir_bandpass(nc, f_low, f_high, (double)samplerate, wintype, 1, gain, imp.data()) != QC_OK) return nullptr;
    return quisk_cuda_batch_create(QC_C_CDECIMATE, n_channels, imp.data(), nc, 1, 1);
}

qcResample *quisk_cuda_resample_create(int n_channels, int in_rate, int out_rate, double fc, int ncoef, double gain)
{
    if (qc::ensure_device() != QC_OK) return nullptr;
    qcResample *p = new qcResample();
    if (p->r.init(n_channels, in_rate, out_rate, fc, ncoef, gain) != QC_OK) { p->r.release(); delete p; return nullptr; }
    return p;
}
void quisk_cuda_resample_destroy(qcResample *r) { if (r) { r->r.release(); delete r; } }
int quisk_cuda_resample_count_out(const qcResample *r, int count) { return r ? r->r.f->count_out(count, 0) : QC_EINVAL; }
int quisk_cuda_resample_run(qcResample *r, const void *d_in, long in_stride, int count, void *d_out, long out_stride, int *n_out, void *stream)
{ return r ? r->r.f->run(d_in, in_stride, count, d_out, out_stride, n_out, 0, (cudaStream_t)stream) : QC_EINVAL; }

}  // extern "C"
