// quisk_b200/csrc/pfb.cu -- wideband polyphase channelizer (configuration C5).
//
// What it replaces: K receivers on ONE wideband stream, each running the reference's front end
//     tune:      v_k[n] = x[n] * exp(-2 pi i k n / K)             (quisk.c:2477-2494, receiver k centred on k fs / K)
//     decimate:  y_k[m] = sum_t h[t] v_k[n_m - t],  n_m = D m + D - 1  (quisk_cDecimate, filter.c:203-229)
// With j = n_m - t and r = j mod K the phase factor depends on r only, so
//     y_k[m] = sum_{r<K} exp(-2 pi i k r / K) u_m[r],     u_m[r] = sum_{j = r (mod K), n_m - T < j <= n_m} h[n_m - j] x[j]
// i.e. a T/K-tap FIR per branch followed by one K-point forward DFT per output frame: T MACs + one FFT per frame
// for all K receivers instead of K * T MACs.  The identity is exact; only the summation order differs from the
// reference (<= 1e-15 relative).  D <= K (D = K/2: the 2x oversampled C5 case; D = K: critically sampled).
//
// Mapping: a CTA of K/4 threads = 4 transforms x K/16 lanes handles FOUR consecutive frames per round.
//   FIR phase : thread owns 4 branches r; per branch it loads the P + SMAX input samples the four frames need
//               (coalesced 16-byte loads, history or block selected per load), and accumulates the four frames with
//               taps read from shared memory (the whole prototype, T doubles, is staged once per CTA).
//   FFT phase : four K-point transforms side by side (fft_device.cuh), in place on the u buffers.
//   store     : channel-major [k][frame] (what the per-receiver chains consume: 64-byte runs per channel and round)
//               or frame-major [frame][k].
// The grid is persistent: one CTA per SM striding over the rounds.  State between calls: the last T input samples
// and the absolute sample index (frame phase and branch alignment both follow from it), so any block length works
// and a time-block shard (shard.py) reproduces the sequential stream bit for bit after seek + prime.
#include "fft_device.cuh"

namespace qc {

struct PfbParams {
    const cd *in; int count;
    const cd *hist; int H;          // the H samples before in[0], oldest first
    long long n0;                   // absolute index of in[0]
    long long m0;                   // absolute index of the first frame of this call
    int n_frames;
    int K, D, T, lgK;
    const double *taps;
    const cd *tw;
    cd *out; long out_stride; int layout;
};

static constexpr int PF = 4;        // frames per round

template <int P, int SMAX>
__global__ void __launch_bounds__(256, 1) pfb_kernel(PfbParams p)
{
    extern __shared__ double smem_raw[];
    constexpr int WN = P + SMAX;
    const int K = p.K, D = p.D, T = p.T;
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *u = twl + fft_tw_entries(K);                    // [PF][K]
    double *sh = reinterpret_cast<double *>(u + (size_t)PF * K);     // [T]
    const int lanes = blockDim.x, NT = blockDim.x * blockDim.y;
    const int tid = threadIdx.y * lanes + threadIdx.x;
    fft_stage_twiddles(twl, p.tw, K);
    for (int i = tid; i < T; i += NT) sh[i] = p.taps[i];
    __syncthreads();
    const int rounds = (p.n_frames + PF - 1) / PF;
    for (int g = blockIdx.x; g < rounds; g += gridDim.x) {
        const long long mg = p.m0 + (long long)g * PF;
        const long long ng = mg * D + D - 1;                                   // absolute index of frame mg's newest sample
        // ---- branch FIRs
#pragma unroll 1
        for (int b = 0; b < 4; b++) {
            const int r = tid + NT * b;
            const long long base0 = ng - ((ng - r) & (long long)(K - 1));      // newest index <= ng congruent to r
            cd w[WN];
#pragma unroll
            for (int i = 0; i < WN; i++) {
                const long long jr = base0 + (long long)K * (i - (P - 1)) - p.n0;      // relative to in[0]
                cd v = make_double2(0.0, 0.0);
                if (jr >= 0) { if (jr < p.count) v = p.in[jr]; }
                else if (jr >= -(long long)p.H) v = p.hist[p.H + jr];
                w[i] = v;
            }
#pragma unroll
            for (int f = 0; f < PF; f++) {
                const long long nn = ng + (long long)f * D;
                const int t0 = (int)((nn - r) & (long long)(K - 1));
                const int s = (int)((nn - t0 - base0) >> p.lgK);                   // 0 .. SMAX
                double ar = 0.0, ai = 0.0;
#pragma unroll
                for (int i = 0; i < WN; i++) {
                    const int ti = t0 + K * (P - 1 + s - i);
                    const double c = (unsigned)ti < (unsigned)T ? sh[ti] : 0.0;
                    ar = fma(w[i].x, c, ar);
                    ai = fma(w[i].y, c, ai);
                }
                u[(size_t)f * K + fsw(r)] = make_double2(ar, ai);
            }
        }
        __syncthreads();
        // ---- one forward transform per frame
        fft_smem(u + (size_t)threadIdx.y * K, K, twl, -1, threadIdx.x, lanes);
        // ---- store
        const int nf = min(PF, p.n_frames - g * PF);
        if (p.layout == 0) {
            for (int k = tid; k < K; k += NT) {
                cd *o = p.out + (size_t)k * p.out_stride + (size_t)g * PF;
                for (int f = 0; f < nf; f++) o[f] = u[(size_t)f * K + fsw(k)];
            }
        } else {
            for (int f = 0; f < nf; f++) {
                cd *o = p.out + ((size_t)g * PF + f) * p.out_stride;
                for (int k = tid; k < K; k += NT) o[k] = u[(size_t)f * K + fsw(k)];
            }
        }
        __syncthreads();
    }
}

struct Channelizer {
    int K = 0, D = 0, T = 0, P = 0, smax = 0;
    double *d_taps = nullptr;
    cd *d_hist[2] = {nullptr, nullptr};
    int cur = 0;
    long long n_abs = 0;
    const cd *tw = nullptr;
    int n_sm = 148;

    int init(int K_, int D_, const double *proto, int T_)
    {
        K = K_; D = D_; T = T_;
        if (K < 256 || K > 1024 || (K & (K - 1)) || fft_log2(K) < 0) { set_error("pfb_create: n_channels must be 256, 512 or 1024 (got %d)", K); return QC_EINVAL; }
        if (D < 1 || D > K) { set_error("pfb_create: decimation must be in [1, n_channels] (got %d)", D); return QC_EINVAL; }
        if (T < K || T % K) { set_error("pfb_create: n_taps must be a multiple of n_channels (got %d)", T); return QC_EINVAL; }
        P = T / K;
        if (P != 4 && P != 8 && P != 16 && P != 32) { set_error("pfb_create: n_taps / n_channels must be 4, 8, 16 or 32 (got %d)", P); return QC_EINVAL; }
        smax = ((PF - 1) * D + K - 1) / K;              // how far the newest aligned sample can move within a round
        if (smax < 2) smax = 2;
        tw = fft_twiddles(K);
        if (!tw) { set_error("pfb_create: twiddle table allocation failed"); return QC_ENOMEM; }
        QC_CUDA(cudaMalloc((void **)&d_taps, (size_t)T * sizeof(double)));
        QC_CUDA(cudaMemcpy(d_taps, proto, (size_t)T * sizeof(double), cudaMemcpyHostToDevice));
        for (int i = 0; i < 2; i++) {
            QC_CUDA(cudaMalloc((void **)&d_hist[i], (size_t)T * sizeof(cd)));
            QC_CUDA(cudaMemset(d_hist[i], 0, (size_t)T * sizeof(cd)));
        }
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        return QC_OK;
    }
    void release()
    {
        if (d_taps) cudaFree(d_taps);
        for (int i = 0; i < 2; i++) if (d_hist[i]) cudaFree(d_hist[i]);
        d_taps = nullptr; d_hist[0] = d_hist[1] = nullptr;
    }
    int seek(long long n)
    {
        if (n < 0) { set_error("pfb_seek: negative sample index"); return QC_EINVAL; }
        for (int i = 0; i < 2; i++) QC_CUDA(cudaMemset(d_hist[i], 0, (size_t)T * sizeof(cd)));
        n_abs = n; cur = 0;
        return QC_OK;
    }
    int frames(int count) const { return (int)((n_abs + count) / D - n_abs / D); }
    // history <- last T samples of [history | block]
    int roll(const cd *d_in, int count, cudaStream_t s)
    {
        cd *nh = d_hist[cur ^ 1];
        if (count >= T) {
            QC_CUDA(cudaMemcpyAsync(nh, d_in + (count - T), (size_t)T * sizeof(cd), cudaMemcpyDeviceToDevice, s));
        } else {
            QC_CUDA(cudaMemcpyAsync(nh, d_hist[cur] + count, (size_t)(T - count) * sizeof(cd), cudaMemcpyDeviceToDevice, s));
            QC_CUDA(cudaMemcpyAsync(nh + (T - count), d_in, (size_t)count * sizeof(cd), cudaMemcpyDeviceToDevice, s));
        }
        cur ^= 1;
        n_abs += count;
        return QC_OK;
    }
    template <int PP, int SM> int launch(const PfbParams &p, size_t sh, int grid, cudaStream_t s)
    {
        QC_CUDA(cudaFuncSetAttribute(pfb_kernel<PP, SM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        pfb_kernel<PP, SM><<<grid, dim3(K / 16, 4), sh, s>>>(p);
        return QC_OK;
    }
    int process(const cd *d_in, int count, cd *d_out, long out_stride, int layout, int *n_frames, cudaStream_t s)
    {
        if (count < 0) { set_error("pfb_process: negative count"); return QC_EINVAL; }
        const int nf = frames(count);
        if (n_frames) *n_frames = nf;
        if (count == 0) return QC_OK;
        if (nf > 0) {
            if (layout == 0 ? out_stride < nf : out_stride < K) { set_error("pfb_process: out_stride %ld too small", out_stride); return QC_EINVAL; }
            PfbParams p;
            p.in = d_in; p.count = count; p.hist = d_hist[cur]; p.H = T; p.n0 = n_abs; p.m0 = n_abs / D; p.n_frames = nf;
            p.K = K; p.D = D; p.T = T; p.lgK = fft_log2(K); p.taps = d_taps; p.tw = tw; p.out = d_out; p.out_stride = out_stride; p.layout = layout;
            const size_t sh = ((size_t)fft_tw_entries(K) + (size_t)PF * K) * sizeof(cd) + (size_t)T * sizeof(double);
            if (sh > 226 * 1024) { set_error("pfb_process: prototype of %d taps does not fit shared memory", T); return QC_EINVAL; }
            const int rounds = (nf + PF - 1) / PF;
            const int grid = rounds < n_sm ? rounds : n_sm;
            int rc = QC_EINVAL;
#define PFB_CASE(PP) case PP: rc = smax <= 2 ? launch<PP, 2>(p, sh, grid, s) : launch<PP, 3>(p, sh, grid, s); break;
            switch (P) { PFB_CASE(4) PFB_CASE(8) PFB_CASE(16) PFB_CASE(32) }
#undef PFB_CASE
            if (rc != QC_OK) return rc;
            count_launch();
            QC_CUDA_LAUNCH();
        }
        return roll(d_in, count, s);
    }
};

}  // namespace qc

struct qcChannelizer { qc::Channelizer c; };

extern "C" {

qcChannelizer *quisk_cuda_pfb_create(int n_channels, int decim, const double *proto, int n_taps)
{
    if (qc::ensure_device() != QC_OK) return nullptr;
    if (!proto) { qc::set_error("pfb_create: null prototype"); return nullptr; }
    qcChannelizer *p = new qcChannelizer();
    if (p->c.init(n_channels, decim, proto, n_taps) != QC_OK) { p->c.release(); delete p; return nullptr; }
    return p;
}
void quisk_cuda_pfb_destroy(qcChannelizer *p) { if (p) { p->c.release(); delete p; } }
int quisk_cuda_pfb_count_out(const qcChannelizer *p, int count) { return p && count >= 0 ? p->c.frames(count) : QC_EINVAL; }
int quisk_cuda_pfb_seek(qcChannelizer *p, long long n_abs) { return p ? p->c.seek(n_abs) : QC_EINVAL; }
int quisk_cuda_pfb_prime(qcChannelizer *p, const void *d_in, int count, void *stream)
{
    if (!p || count < 0) return QC_EINVAL;
    if (count == 0) return QC_OK;
    return p->c.roll((const cd *)d_in, count, (cudaStream_t)stream);
}
int quisk_cuda_pfb_process(qcChannelizer *p, const void *d_in, int count, void *d_out, long out_stride, int layout,
                           int *n_frames, void *stream)
{
    if (!p) { qc::set_error("pfb_process: null handle"); return QC_EINVAL; }
    return p->c.process((const cd *)d_in, count, (cd *)d_out, out_stride, layout, n_frames, (cudaStream_t)stream);
}

}  // extern "C"
