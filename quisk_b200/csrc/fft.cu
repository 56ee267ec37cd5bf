// quisk_b200/csrc/fft.cu -- batched FFT entry point and the panadapter (get_graph,
// quisk.c:5142-5331; get_multirx_graph, quisk.c:4868-4930; window quisk.c:6003-6009).
#include "fft_device.cuh"
#include <cmath>
#include <map>

namespace qc {

int fft_log2(int n)
{
    if (n < 8 || n > 8192 || (n & (n - 1))) return -1;
    int l = 0;
    while ((1 << l) < n) l++;
    return l;
}

void fft_shape(int n, int *lanes, int *per_cta)
{
    const int t = fft_threads(n);
    *lanes = t;
    int pc = 128 / t;                 // aim for >= 128 threads per CTA
    if (pc < 1) pc = 1;
    *per_cta = pc;
}

const cd *fft_twiddles(int n)
{
    static std::mutex mu;
    static std::map<std::pair<int, int>, cd *> cache;
    std::lock_guard<std::mutex> g(mu);
    int dev = 0;
    cudaGetDevice(&dev);
    auto key = std::make_pair(dev, n);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    std::vector<cd> h((size_t)n);
    for (int k = 0; k < n; k++) {
        const double a = 2.0 * M_PI * (double)k / (double)n;
        h[k] = make_double2(cos(a), -sin(a));
    }
    cd *d = nullptr;
    if (cudaMalloc((void **)&d, (size_t)n * sizeof(cd)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), (size_t)n * sizeof(cd), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    cache[key] = d;
    return d;
}

// ---------------------------------------------------------------------------------------
// plain batched transform
// ---------------------------------------------------------------------------------------
template <int BPT>
__global__ void __launch_bounds__(256) fft_batch_kernel(const cd *in, cd *out, int n, int batch, const cd *tw, int sign)
{
    extern __shared__ double smem_raw[];
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *s = twl + fft_tw_entries(n) + (size_t)threadIdx.y * n;
    fft_stage_twiddles(twl, tw, n);
    const int f = blockIdx.x * blockDim.y + threadIdx.y;
    const bool live = f < batch;
    const int lane = threadIdx.x, lanes = blockDim.x;
    if (live) {
        const cd *src = in + (size_t)f * n;
        for (int i = lane; i < n; i += lanes) s[fsw(i)] = src[i];
    }
    __syncthreads();
    fft_smem<BPT>(s, n, twl, sign, lane, lanes);
    if (live) {
        cd *dst = out + (size_t)f * n;
        for (int i = lane; i < n; i += lanes) dst[i] = s[fsw(i)];
    }
}

// ---------------------------------------------------------------------------------------
// panadapter accumulate: window -> FFT -> fftshift -> |X| -> running sum
//
// grid = (groups, n_streams).  A CTA walks frames g, g+groups, ... of its stream and keeps
// the per-bin sums of the bins its threads own in registers; with groups == 1 the sum over
// frames is formed in the reference's order (fft_avg[k] += cabs(...) frame by frame,
// quisk.c:5271-5276) starting from the value already in `avg`.  With groups > 1 each group
// writes a partial sum and pan_reduce_kernel folds them in.
// ---------------------------------------------------------------------------------------
static constexpr int PAN_MAX_OWN = 16;     // bins per thread: n / fft_threads(n) = 16

template <int BPT>
__global__ void __launch_bounds__(256) pan_accumulate_kernel(const cd *frames, long stream_stride, int n_frames, int n,
                                      const cd *tw, const double *window, double *avg, double *partial, int groups)
{
    extern __shared__ double smem_raw[];
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *s = twl + fft_tw_entries(n);
    double *sacc = reinterpret_cast<double *>(s + n);        // running per-bin sums of this CTA
    fft_stage_twiddles(twl, tw, n);
    const int stream = blockIdx.y, g = blockIdx.x;
    const int lane = threadIdx.x, lanes = blockDim.x;
    const int half = n >> 1;
    double *dst = groups == 1 ? avg + (size_t)stream * n : partial + ((size_t)stream * groups + g) * n;
    for (int k = lane; k < n; k += lanes) sacc[k] = groups == 1 ? dst[k] : 0.0;
    const cd *base = frames + (size_t)stream * stream_stride;
    for (int f = g; f < n_frames; f += groups) {
        const cd *src = base + (size_t)f * n;
        for (int i = lane; i < n; i += lanes) {
            const cd x = src[i];
            const double w = window[i];
            s[fsw(i)] = make_double2(x.x * w, x.y * w);             // quisk.c:5212-5213
        }
        __syncthreads();
        fft_smem<BPT>(s, n, twl, -1, lane, lanes);
        for (int k = lane; k < n; k += lanes) {                     // graph bin k <- FFT bin (k + n/2) mod n
            const cd X = s[fsw((k + half) & (n - 1))];
            sacc[k] += sqrt(fma(X.x, X.x, X.y * X.y));              // cabs, quisk.c:5273,5275 (no overflow at these scales)
        }
        __syncthreads();
    }
    for (int k = lane; k < n; k += lanes) dst[k] = sacc[k];
}

// n = 8192 split over TWO CTAs per frame.  One 8192-point transform needs 128 KiB of shared memory, i.e. one CTA of
// 8 warps per SM with every load, barrier and pass exposed.  A decimation-in-frequency first step done on the way in
// from global memory halves that:  X[2k] = FFT4096(x[j] + x[j+4096]),  X[2k+1] = FFT4096((x[j] - x[j+4096]) w^j),
// so CTA parity p owns the bins of parity p, works on 64 KiB, runs one butterfly per thread (128 registers) and two
// such CTAs share an SM.  Each frame is read by both CTAs of its pair (the second read is an L2 hit when the pair is
// co-scheduled, which adjacent blockIdx.x makes the common case).  Bin sums keep the reference's frame order.
// SP = 2, 4, 8: frames of N = SP x 4096 points (8192, 16384, 32768).  The first decimation-in-frequency step, radix SP,
// is done on the way in:  X[SP k + q] = FFT4096( w^{j q} sum_m x[j + 4096 m] W_SP^{m q} ),  w = exp(-2 pi i / N).
template <int SP>
__global__ void __launch_bounds__(256, 2) pan_accumulate_split_kernel(const cd *frames, long stream_stride, int n_frames,
                                      const cd *twN, const cd *tw4, const double *window, double *avg, double *partial, int groups)
{
    constexpr int H = 4096, N = SP * H;
    extern __shared__ double smem_raw[];
    cd *twl4 = reinterpret_cast<cd *>(smem_raw);
    cd *twlN = twl4 + fft_tw_entries(H);
    cd *s = twlN + fft_tw_entries(N);
    double *sacc = reinterpret_cast<double *>(s + H);        // [H] running sums of this CTA's bins
    fft_stage_twiddles(twl4, tw4, H);
    fft_stage_twiddles(twlN, twN, N);
    const int stream = blockIdx.y, g = blockIdx.x / SP, par = blockIdx.x % SP;
    const int lane = threadIdx.x, lanes = blockDim.x;
    double *dst = groups == 1 ? avg + (size_t)stream * N : partial + ((size_t)stream * groups + g) * N;
    // FFT bin b = SP k + par shows at graph bin (b + N/2) mod N (fftshift, quisk.c:5271-5276)
    for (int k = lane; k < H; k += lanes) sacc[k] = groups == 1 ? dst[(SP * k + par + N / 2) & (N - 1)] : 0.0;
    const cd *base = frames + (size_t)stream * stream_stride;
    // W_SP^{m par}, m < SP: exp(-2 pi i m par / SP) = entry (m par mod SP) N / SP of the N-point table
    cd wq[SP];
    __syncthreads();
#pragma unroll
    for (int m = 0; m < SP; m++) wq[m] = fft_tw(twlN, ((m * par) % SP) * (N / SP), -1);
    for (int f = g; f < n_frames; f += groups) {
        const cd *src = base + (size_t)f * N;
        for (int j = lane; j < H; j += lanes) {
            cd v = make_double2(0.0, 0.0);
#pragma unroll
            for (int m = 0; m < SP; m++) {
                const cd x = src[j + m * H];
                const double w = window[j + m * H];
                const cd a = make_double2(x.x * w, x.y * w);                        // quisk.c:5212-5213
                if (m == 0 || par == 0) v = cadd(v, a);
                else v = cadd(v, cmul(a, wq[m]));
            }
            if (par != 0) v = cmul(v, fft_tw(twlN, j * par, -1));
            s[fsw(j)] = v;
        }
        __syncthreads();
        // pull this CTA's next frame towards L2 while the transform runs (each CTA of the group fetches its share):
        // the load phase above is otherwise the only time this CTA has memory requests in flight
        if (f + groups < n_frames) {
            const char *nxt = reinterpret_cast<const char *>(base + (size_t)(f + groups) * N) + (size_t)par * (N * sizeof(cd) / SP);
            for (int l = lane; l < (int)(N * sizeof(cd) / SP / 128); l += lanes)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t)l * 128));
        }
        fft_smem<1>(s, H, twl4, -1, lane, lanes);
        for (int k = lane; k < H; k += lanes) {
            const cd X = s[fsw(k)];
            sacc[k] += sqrt(fma(X.x, X.x, X.y * X.y));
        }
        __syncthreads();
    }
    for (int k = lane; k < H; k += lanes) dst[(SP * k + par + N / 2) & (N - 1)] = sacc[k];
}

// Frames whose length is NOT a power of two (Quisk's fft_size = data_width x fft_mult, quisk.py:187-194, 4179, has
// factors 3 ... 15): Bluestein's identity turns the n-point DFT into a circular convolution of length M >= 2n - 1,
// M a power of two, so the same shared-memory transform serves:
//     X[k] = w[k] * sum_j (x[j] w[j]) conj(w)[k - j],   w[j] = exp(-i pi j^2 / n)
// a[j] = x[j] window[j] w[j] (zero padded to M) -> FFT_M -> times B = FFT_M(conj(w) wrapped) -> inverse FFT_M -> times
// w[k] / M.  The chirp is built on the host with j^2 reduced mod 2n in integers, so its angle error does not grow
// with j.  Costs two M-point transforms per frame (M up to 8192 for n <= 4096): a correctness path, not a fast one.
template <int BPT>
__global__ void __launch_bounds__(256) pan_accumulate_bluestein_kernel(const cd *frames, long stream_stride, int n_frames, int n, int M,
                                      const cd *tw, const cd *chirpwin, const cd *chirp, const cd *B, double *avg, double *partial, int groups)
{
    extern __shared__ double smem_raw[];
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *s = twl + fft_tw_entries(M);
    double *sacc = reinterpret_cast<double *>(s + M);        // [n]
    fft_stage_twiddles(twl, tw, M);
    const int stream = blockIdx.y, g = blockIdx.x;
    const int lane = threadIdx.x, lanes = blockDim.x;
    const int half = n / 2;                                  // graph bin k <- FFT bin (k + n/2) mod n, quisk.c:5271-5276
    const double inv_M = 1.0 / (double)M;
    double *dst = groups == 1 ? avg + (size_t)stream * n : partial + ((size_t)stream * groups + g) * n;
    for (int k = lane; k < n; k += lanes) sacc[k] = groups == 1 ? dst[k] : 0.0;
    const cd *base = frames + (size_t)stream * stream_stride;
    for (int f = g; f < n_frames; f += groups) {
        const cd *src = base + (size_t)f * n;
        for (int j = lane; j < M; j += lanes) s[fsw(j)] = j < n ? cmul(src[j], chirpwin[j]) : make_double2(0.0, 0.0);
        __syncthreads();
        fft_smem<BPT>(s, M, twl, -1, lane, lanes);
        for (int m = lane; m < M; m += lanes) s[fsw(m)] = cmul(s[fsw(m)], B[m]);
        __syncthreads();
        fft_smem<BPT>(s, M, twl, +1, lane, lanes);
        for (int k = lane; k < n; k += lanes) {
            int b = k + half; if (b >= n) b -= n;
            const cd X = cmul(s[fsw(b)], chirp[b]);
            sacc[k] += sqrt(fma(X.x, X.x, X.y * X.y)) * inv_M;
        }
        __syncthreads();
    }
    for (int k = lane; k < n; k += lanes) dst[k] = sacc[k];
}

// Bluestein for frame sizes above 4096 that are not powers of two (data_width x fft_mult up to 16384): the convolution length
// M = 16384 or 32768 no longer fits one CTA's shared memory, so the two M-point transforms run as SP = M / 4096 CTAs per
// frame each -- a radix-SP decimation-in-frequency step on the way in, a 4096-point transform in shared memory -- with the
// spectrum kept in its p-major order Z[p][k] = X[SP k + p] between them (the inverse consumes exactly that order):
//   blue_big_fwd_kernel:  a = x window chirp (zero padded);  y_p[j] = (sum_q a[j + 4096 q] W_SP^{pq}) W_M^{pj};  Z[p] = FFT_4096(y_p) * B[SP k + p]
//   blue_big_inv_kernel:  U[p][j] = conj(W_M^{pj}) * IFFT_4096(Z[p])[j]
//   blue_big_acc_kernel:  c[j + 4096 q] = sum_p conj(W_SP^{pq}) U[p][j];  avg[k] += |c[b] chirp[b]| / M,  b = (k + n/2) mod n, frames in order
// A correctness path like the small-size kernel above: three launches and a round trip through a scratch buffer per batch of frames.
template <int SP>
__global__ void __launch_bounds__(256) blue_big_fwd_kernel(const cd *frames, long stream_stride, int f0, int n, const cd *twM, const cd *tw4,
                                                            const cd *chirpwin, const cd *B, cd *Z)
{
    extern __shared__ double smem_raw[];
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *s = twl + fft_tw_entries(4096);
    fft_stage_twiddles(twl, tw4, 4096);
    constexpr int M = SP * 4096;
    const int p = blockIdx.x, f = blockIdx.y, stream = blockIdx.z, lane = threadIdx.x;
    const cd *src = frames + (size_t)stream * stream_stride + (size_t)(f0 + f) * n;
    for (int j = lane; j < 4096; j += 256) {
        cd acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 0; q < SP; q++) {
            const int idx = j + 4096 * q;
            if (idx < n) {
                const cd a = cmul(src[idx], chirpwin[idx]);
                const cd w = twM[((p * q) & (SP - 1)) * 4096];              // W_SP^{pq}
                acc = cadd(acc, cmul(a, w));
            }
        }
        s[fsw(j)] = cmul(acc, twM[p * j]);
    }
    __syncthreads();
    fft_smem<1>(s, 4096, twl, -1, lane, 256);
    cd *z = Z + (((size_t)stream * gridDim.y + f) * SP + p) * 4096;
    for (int k = lane; k < 4096; k += 256) z[k] = cmul(s[fsw(k)], B[SP * k + p]);
    (void)M;
}

template <int SP>
__global__ void __launch_bounds__(256) blue_big_inv_kernel(const cd *twM, const cd *tw4, cd *Z)
{
    extern __shared__ double smem_raw[];
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *s = twl + fft_tw_entries(4096);
    fft_stage_twiddles(twl, tw4, 4096);
    const int p = blockIdx.x, f = blockIdx.y, stream = blockIdx.z, lane = threadIdx.x;
    cd *z = Z + (((size_t)stream * gridDim.y + f) * SP + p) * 4096;
    for (int k = lane; k < 4096; k += 256) s[fsw(k)] = z[k];
    __syncthreads();
    fft_smem<1>(s, 4096, twl, +1, lane, 256);
    for (int j = lane; j < 4096; j += 256) { cd w = twM[p * j]; w.y = -w.y; z[j] = cmul(s[fsw(j)], w); }
}

template <int SP>
__global__ void __launch_bounds__(256) blue_big_acc_kernel(const cd *Z, int nfb, int n, const cd *twM, const cd *chirp, double *avg)
{
    const int stream = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int b = k + n / 2; if (b >= n) b -= n;
    const int j = b & 4095, q = b >> 12;
    const double inv_M = 1.0 / (double)(SP * 4096);
    const cd ch = chirp[b];
    double a = avg[(size_t)stream * n + k];
    for (int f = 0; f < nfb; f++) {
        const cd *u = Z + ((size_t)stream * nfb + f) * SP * 4096;
        cd c = make_double2(0.0, 0.0);
#pragma unroll
        for (int pp = 0; pp < SP; pp++) { cd w = twM[((pp * q) & (SP - 1)) * 4096]; w.y = -w.y; c = cadd(c, cmul(u[(size_t)pp * 4096 + j], w)); }
        const cd X = cmul(c, ch);
        a += sqrt(fma(X.x, X.x, X.y * X.y)) * inv_M;
    }
    avg[(size_t)stream * n + k] = a;
}

__global__ void pan_reduce_kernel(double *avg, const double *partial, int n, int groups)
{
    const int stream = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double a = avg[(size_t)stream * n + k];
    for (int g = 0; g < groups; g++) a += partial[((size_t)stream * groups + g) * n + k];
    avg[(size_t)stream * n + k] = a;
}

// graph-return half (quisk.c:5279-5321), hazard-free case: every pixel reads only bins >= its own index
// First FFT bin of pixel i (quisk.c:5290), with every product and sum rounded separately as the reference's x86-64
// code does: contracted into FMAs the value can land on the other side of an integer when fft_size is not a power of
// two (n * x is exact only for powers of two), and the pixel would sum a different set of bins.
__device__ __forceinline__ int pan_first_bin(int n, int i, int data_width, double zoom, double deltaf, double rate)
{
    const double t = __dadd_rn(__dadd_rn(__ddiv_rn(deltaf, rate), __dmul_rn(zoom, __dsub_rn(__ddiv_rn((double)i, (double)data_width), 0.5))), 0.5);
    return (int)__dadd_rn(__dmul_rn((double)n, t), 0.1);
}

__global__ void pan_graph_kernel(double *avg, int n, int data_width, int nbin, double zoom, double deltaf,
                                 double rate, double scale, double *graph)
{
    const int stream = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= data_width) return;
    const double *a = avg + (size_t)stream * n;
    int k = pan_first_bin(n, i, data_width, zoom, deltaf, rate);
    double d2 = 0.0;
    for (int j = 0; j < nbin; j++, k++)
        if (k >= 0 && k < n) d2 += a[k];
    d2 = 20.0 * log10(d2) - scale;
    if (d2 < -200) d2 = -200;
    else if (d2 > 0) d2 = 0;
    graph[(size_t)stream * data_width + i] = d2;
}

// same, literal in-place walk for the zoomed cases where pixel i reads bins already overwritten
__global__ void pan_graph_serial_kernel(double *avg, int n, int data_width, int nbin, double zoom, double deltaf,
                                        double rate, double scale, double *graph)
{
    double *a = avg + (size_t)blockIdx.y * n;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int i = 0; i < data_width; i++) {
        int k = pan_first_bin(n, i, data_width, zoom, deltaf, rate);
        double d2 = 0.0;
        for (int j = 0; j < nbin; j++, k++)
            if (k >= 0 && k < n) d2 += a[k];
        a[i] = d2;
    }
    for (int i = 0; i < data_width; i++) {
        double d2 = 20.0 * log10(a[i]) - scale;
        if (d2 < -200) d2 = -200;
        else if (d2 > 0) d2 = 0;
        graph[(size_t)blockIdx.y * data_width + i] = d2;
    }
}

// get_multirx_graph: one frame per stream, |X| summed over groups of 8 bins in fftshift order
__global__ void __launch_bounds__(256) pan_multirx_kernel(const cd *frames, long stream_stride, int n, const cd *tw, const double *window,
                                   double scale, double *graph)
{
    extern __shared__ double smem_raw[];
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *s = twl + fft_tw_entries(n);
    fft_stage_twiddles(twl, tw, n);
    const int stream = blockIdx.x;
    const int lane = threadIdx.x, lanes = blockDim.x;
    const int half = n >> 1;
    const cd *src = frames + (size_t)stream * stream_stride;
    for (int i = lane; i < n; i += lanes) {
        const cd x = src[i];
        const double w = window[i];
        s[fsw(i)] = make_double2(x.x * w, x.y * w);
    }
    __syncthreads();
    fft_smem(s, n, twl, -1, lane, lanes);
    for (int p = lane; p < n / 8; p += lanes) {
        double d1 = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const cd X = s[fsw((p * 8 + j + half) & (n - 1))];
            d1 += sqrt(fma(X.x, X.x, X.y * X.y));
        }
        double d2 = 20.0 * log10(d1) - scale;
        if (d2 < -200) d2 = -200;
        graph[(size_t)stream * (n / 8) + p] = d2;
    }
}

struct Panadapter {
    int S = 0, n = 0;
    const cd *tw = nullptr;
    double *d_window = nullptr, *d_avg = nullptr, *d_partial = nullptr;
    int partial_groups = 0;
    int count = 0;
    int split8192 = 1;          // 8192-point frames: two 4096-point CTAs per frame (pan_accumulate_split_kernel)
    int M = 0;                  // > 0: Bluestein convolution length for a frame size that is not a power of two
    cd *d_chirpwin = nullptr, *d_chirp = nullptr, *d_B = nullptr;
    const cd *tw4 = nullptr;    // M > 8192: the 4096-point table of the split transforms
    cd *d_Z = nullptr; size_t z_cap = 0;

    int init(int streams, int fft_size)
    {
        S = streams; n = fft_size;
        const bool pow2 = fft_log2(n) >= 0 || n == 16384 || n == 32768;       // the two largest run split over 4 / 8 CTAs
        if (S <= 0 || (!pow2 && (n < 8 || n > 16384))) {
            set_error("pan_create: fft_size must be a power of two in [8, 32768] or any size in [8, 16384] (got %d)", fft_size); return QC_EINVAL;
        }
        if (!pow2) { M = 16; while (M < 2 * n - 1) M <<= 1; }
        tw = fft_twiddles(pow2 ? n : M);
        if (!tw) { set_error("pan_create: twiddle table allocation failed"); return QC_ENOMEM; }
        std::vector<double> w((size_t)n);
        for (int i = 0, j = -n / 2; i < n; i++, j++) w[i] = 0.5 + 0.5 * cos(2. * M_PI * j / n);      // quisk.c:6003-6009
        QC_CUDA(cudaMalloc((void **)&d_window, (size_t)n * sizeof(double)));
        QC_CUDA(cudaMemcpy(d_window, w.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
        QC_CUDA(cudaMalloc((void **)&d_avg, (size_t)S * n * sizeof(double)));
        QC_CUDA(cudaMemset(d_avg, 0, (size_t)S * n * sizeof(double)));
        if (M) {
            // chirp w[j] = exp(-i pi j^2 / n), j^2 reduced mod 2n exactly; b = conj(w) wrapped around M; B = FFT_M(b)
            std::vector<cd> ch((size_t)n), cw((size_t)n), b((size_t)M, make_double2(0.0, 0.0));
            for (int j = 0; j < n; j++) {
                const long long t = ((long long)j * j) % (2LL * n);
                const double a = M_PI * (double)t / (double)n;
                ch[j] = make_double2(cos(a), -sin(a));
                cw[j] = make_double2(ch[j].x * w[j], ch[j].y * w[j]);
                b[j] = make_double2(ch[j].x, -ch[j].y);
                if (j) b[M - j] = b[j];
            }
            QC_CUDA(cudaMalloc((void **)&d_chirp, (size_t)n * sizeof(cd)));
            QC_CUDA(cudaMalloc((void **)&d_chirpwin, (size_t)n * sizeof(cd)));
            QC_CUDA(cudaMalloc((void **)&d_B, (size_t)M * sizeof(cd)));
            QC_CUDA(cudaMemcpy(d_chirp, ch.data(), (size_t)n * sizeof(cd), cudaMemcpyHostToDevice));
            QC_CUDA(cudaMemcpy(d_chirpwin, cw.data(), (size_t)n * sizeof(cd), cudaMemcpyHostToDevice));
            QC_CUDA(cudaMemcpy(d_B, b.data(), (size_t)M * sizeof(cd), cudaMemcpyHostToDevice));
            if (M <= 8192) {
                int rc = quisk_cuda_fft_batch(d_B, d_B, M, 1, -1, nullptr); if (rc != QC_OK) return rc;
                QC_CUDA(cudaDeviceSynchronize());
            } else {
                // M = 16384 / 32768: the chirp's spectrum once, on the host (iterative radix 2 in long double)
                std::vector<long double> re((size_t)M), im((size_t)M);
                for (int i = 0, jr = 0; i < M; i++) {
                    re[jr] = b[i].x; im[jr] = b[i].y;
                    int bit = M >> 1;
                    for (; jr & bit; bit >>= 1) jr ^= bit;
                    jr ^= bit;
                }
                const long double PI_L = 3.141592653589793238462643383279502884L;
                for (int len = 2; len <= M; len <<= 1) {
                    for (int k = 0; k < len / 2; k++) {
                        const long double a = -2.0L * PI_L * k / len, wr = cosl(a), wi = sinl(a);
                        for (int i = k; i < M; i += len) {
                            const int j2 = i + len / 2;
                            const long double tr = re[j2] * wr - im[j2] * wi, ti = re[j2] * wi + im[j2] * wr;
                            re[j2] = re[i] - tr; im[j2] = im[i] - ti; re[i] += tr; im[i] += ti;
                        }
                    }
                }
                for (int i = 0; i < M; i++) b[i] = make_double2((double)re[i], (double)im[i]);
                QC_CUDA(cudaMemcpy(d_B, b.data(), (size_t)M * sizeof(cd), cudaMemcpyHostToDevice));
                tw4 = fft_twiddles(4096);
                if (!tw4) { set_error("pan_create: twiddle table allocation failed"); return QC_ENOMEM; }
            }
        }
        return QC_OK;
    }
    void release()
    {
        if (d_window) cudaFree(d_window); if (d_avg) cudaFree(d_avg); if (d_partial) cudaFree(d_partial);
        if (d_chirp) cudaFree(d_chirp); if (d_chirpwin) cudaFree(d_chirpwin); if (d_B) cudaFree(d_B); if (d_Z) cudaFree(d_Z);
        d_window = d_avg = d_partial = nullptr; d_chirp = d_chirpwin = d_B = d_Z = nullptr; z_cap = 0;
    }
};

static int fft_smem_optin(const void *kern, size_t bytes)
{
    if (bytes > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return QC_OK;
}

}  // namespace qc

using namespace qc;

struct qcPanadapter { qc::Panadapter p; };

extern "C" {

int quisk_cuda_fft_batch(const void *d_in, void *d_out, int n, int batch, int sign, void *stream)
{
    if (ensure_device() != QC_OK) return QC_ENODEV;
    if (fft_log2(n) < 0) { set_error("fft_batch: n must be a power of two in [8, 8192] (got %d)", n); return QC_EINVAL; }
    if (batch <= 0) return QC_OK;
    const cd *tw = fft_twiddles(n);
    if (!tw) { set_error("fft_batch: twiddle table allocation failed"); return QC_ENOMEM; }
    int lanes, per;
    fft_shape(n, &lanes, &per);
    const size_t sh = ((size_t)per * n + fft_tw_entries(n)) * sizeof(cd);
    dim3 block(lanes, per);
    int rc;
    if (n > 4096) {
        rc = fft_smem_optin((const void *)fft_batch_kernel<2>, sh); if (rc != QC_OK) return rc;
        fft_batch_kernel<2><<<(batch + per - 1) / per, block, sh, (cudaStream_t)stream>>>((const cd *)d_in, (cd *)d_out, n, batch, tw, sign < 0 ? -1 : 1);
    } else {
        rc = fft_smem_optin((const void *)fft_batch_kernel<1>, sh); if (rc != QC_OK) return rc;
        fft_batch_kernel<1><<<(batch + per - 1) / per, block, sh, (cudaStream_t)stream>>>((const cd *)d_in, (cd *)d_out, n, batch, tw, sign < 0 ? -1 : 1);
    }
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

qcPanadapter *quisk_cuda_pan_create(int n_streams, int fft_size)
{
    if (ensure_device() != QC_OK) return nullptr;
    qcPanadapter *p = new qcPanadapter();
    if (p->p.init(n_streams, fft_size) != QC_OK) { p->p.release(); delete p; return nullptr; }
    return p;
}

void quisk_cuda_pan_destroy(qcPanadapter *p) { if (p) { p->p.release(); delete p; } }
int quisk_cuda_pan_count(const qcPanadapter *p) { return p ? p->p.count : QC_EINVAL; }
const double *quisk_cuda_pan_average_ptr(const qcPanadapter *p) { return p ? p->p.d_avg : nullptr; }

int quisk_cuda_pan_accumulate(qcPanadapter *pp, const void *d_frames, long stream_stride, int n_frames, void *stream)
{
    if (!pp) { set_error("pan_accumulate: null handle"); return QC_EINVAL; }
    Panadapter &p = pp->p;
    if (n_frames <= 0) return QC_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (p.M > 8192) {
        // Bluestein with a convolution length beyond one CTA: batches of frames through the three split kernels
        const int SPm = p.M / 4096;
        size_t per_frame = (size_t)p.S * p.M;
        int nfb_cap = (int)(((size_t)256 << 20) / (per_frame * sizeof(cd)));
        if (nfb_cap < 1) nfb_cap = 1;
        if (nfb_cap > n_frames) nfb_cap = n_frames;
        if (per_frame * nfb_cap > p.z_cap) {
            if (p.d_Z) cudaFree(p.d_Z);
            p.d_Z = nullptr; p.z_cap = 0;
            QC_CUDA(cudaMalloc((void **)&p.d_Z, per_frame * nfb_cap * sizeof(cd)));
            p.z_cap = per_frame * nfb_cap;
        }
        const size_t sh = ((size_t)4096 + fft_tw_entries(4096)) * sizeof(cd);
        for (int f0 = 0; f0 < n_frames; f0 += nfb_cap) {
            const int nfb = n_frames - f0 < nfb_cap ? n_frames - f0 : nfb_cap;
            const dim3 grid(SPm, nfb, p.S);
            int rc;
#define BLUE_BIG(SP) do { \
            rc = fft_smem_optin((const void *)blue_big_fwd_kernel<SP>, sh); if (rc != QC_OK) return rc; \
            rc = fft_smem_optin((const void *)blue_big_inv_kernel<SP>, sh); if (rc != QC_OK) return rc; \
            blue_big_fwd_kernel<SP><<<grid, 256, sh, s>>>((const cd *)d_frames, stream_stride, f0, p.n, p.tw, p.tw4, p.d_chirpwin, p.d_B, p.d_Z); \
            blue_big_inv_kernel<SP><<<grid, 256, sh, s>>>(p.tw, p.tw4, p.d_Z); \
            blue_big_acc_kernel<SP><<<dim3((p.n + 255) / 256, p.S), 256, 0, s>>>(p.d_Z, nfb, p.n, p.tw, p.d_chirp, p.d_avg); } while (0)
            if (SPm == 4) BLUE_BIG(4); else BLUE_BIG(8);
#undef BLUE_BIG
            count_launch(); count_launch(); count_launch();
            QC_CUDA_LAUNCH();
        }
        p.count += n_frames;
        return QC_OK;
    }
    // enough CTAs to fill the machine, but keep the reference's summation order when streams alone do
    // Frames of a stream are split over `groups` CTAs only when the streams alone cannot fill the machine, and then
    // so that ALL CTAs are resident at once (one wave): a grid a little larger than the number of slots costs a
    // whole extra wave (16 streams: 608 CTAs on 296 slots ran 3 waves of 7 frames; 288 CTAs run 1 wave of 15).
    const bool split = (p.n == 8192 && p.split8192) || p.n > 8192;
    const int cpf = split ? p.n / 4096 : 1;                          // CTAs per frame
    const size_t cta_smem = split ? (size_t)4096 * 24 + (size_t)(fft_tw_entries(4096) + fft_tw_entries(p.n)) * sizeof(cd) + 1024 : (p.M ? (size_t)p.M * 16 + (size_t)p.n * 8 + 4096 : (size_t)p.n * 24 + 4096);
    int per_sm = (int)((size_t)226 * 1024 / cta_smem);
    const int reg_cap = (p.M ? p.M : p.n) >= 4096 ? 2 : 4;                         // 256 threads x ~128 registers (255 at 8192 unsplit)
    if (per_sm > reg_cap) per_sm = reg_cap;
    if ((p.n == 8192 && !split) || p.M == 8192) per_sm = 1;
    if (per_sm < 1) per_sm = 1;
    int n_sm = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); }
    const int slots = per_sm * n_sm;
    int groups = 1;
    if (p.S * cpf < slots) {
        groups = slots / (p.S * cpf);
        if (groups > n_frames) groups = n_frames;
        if (groups < 1) groups = 1;
    }
    if (groups > 1 && groups > p.partial_groups) {
        if (p.d_partial) cudaFree(p.d_partial);
        p.d_partial = nullptr;
        QC_CUDA(cudaMalloc((void **)&p.d_partial, (size_t)p.S * groups * p.n * sizeof(double)));
        p.partial_groups = groups;
    }
    const int lanes = fft_threads(p.M ? p.M : p.n);
    int rc;
    if (p.M) {
        const size_t sh = (size_t)p.M * sizeof(cd) + (size_t)p.n * sizeof(double) + (size_t)fft_tw_entries(p.M) * sizeof(cd);
        if (p.M > 4096) {
            rc = fft_smem_optin((const void *)pan_accumulate_bluestein_kernel<2>, sh); if (rc != QC_OK) return rc;
            pan_accumulate_bluestein_kernel<2><<<dim3(groups, p.S), lanes, sh, s>>>((const cd *)d_frames, stream_stride, n_frames, p.n, p.M, p.tw,
                                                                                     p.d_chirpwin, p.d_chirp, p.d_B, p.d_avg, p.d_partial, groups);
        } else {
            rc = fft_smem_optin((const void *)pan_accumulate_bluestein_kernel<1>, sh); if (rc != QC_OK) return rc;
            pan_accumulate_bluestein_kernel<1><<<dim3(groups, p.S), lanes, sh, s>>>((const cd *)d_frames, stream_stride, n_frames, p.n, p.M, p.tw,
                                                                                     p.d_chirpwin, p.d_chirp, p.d_B, p.d_avg, p.d_partial, groups);
        }
    } else if (split) {
        const cd *tw4 = fft_twiddles(4096);
        if (!tw4) { set_error("pan_accumulate: twiddle table allocation failed"); return QC_ENOMEM; }
        const size_t sh = (size_t)4096 * (sizeof(cd) + sizeof(double)) + (size_t)(fft_tw_entries(4096) + fft_tw_entries(p.n)) * sizeof(cd);
#define PAN_SPLIT(SP) do { rc = fft_smem_optin((const void *)pan_accumulate_split_kernel<SP>, sh); if (rc != QC_OK) return rc; \
        pan_accumulate_split_kernel<SP><<<dim3(SP * groups, p.S), 256, sh, s>>>((const cd *)d_frames, stream_stride, n_frames, p.tw, tw4, \
                                                                                p.d_window, p.d_avg, p.d_partial, groups); } while (0)
        if (p.n == 8192) PAN_SPLIT(2); else if (p.n == 16384) PAN_SPLIT(4); else PAN_SPLIT(8);
#undef PAN_SPLIT
    } else {
        const size_t sh = (size_t)p.n * (sizeof(cd) + sizeof(double)) + (size_t)fft_tw_entries(p.n) * sizeof(cd);
        if (p.n > 4096) {
            rc = fft_smem_optin((const void *)pan_accumulate_kernel<2>, sh); if (rc != QC_OK) return rc;
            pan_accumulate_kernel<2><<<dim3(groups, p.S), lanes, sh, s>>>((const cd *)d_frames, stream_stride, n_frames, p.n, p.tw,
                                                                            p.d_window, p.d_avg, p.d_partial, groups);
        } else {
            rc = fft_smem_optin((const void *)pan_accumulate_kernel<1>, sh); if (rc != QC_OK) return rc;
            pan_accumulate_kernel<1><<<dim3(groups, p.S), lanes, sh, s>>>((const cd *)d_frames, stream_stride, n_frames, p.n, p.tw,
                                                                            p.d_window, p.d_avg, p.d_partial, groups);
        }
    }
    count_launch();
    QC_CUDA_LAUNCH();
    if (groups > 1) {
        pan_reduce_kernel<<<dim3((p.n + 255) / 256, p.S), 256, 0, s>>>(p.d_avg, p.d_partial, p.n, groups);
        count_launch();
        QC_CUDA_LAUNCH();
    }
    p.count += n_frames;
    return QC_OK;
}

int quisk_cuda_pan_graph(qcPanadapter *pp, int data_width, double zoom, double deltaf, double fft_sample_rate,
                         double *d_graph, void *stream)
{
    if (!pp) { set_error("pan_graph: null handle"); return QC_EINVAL; }
    Panadapter &p = pp->p;
    if (p.count <= 0) { set_error("pan_graph: no frames accumulated"); return QC_EINVAL; }
    if (data_width <= 0 || data_width > p.n) { set_error("pan_graph: data_width %d out of range", data_width); return QC_EINVAL; }
    cudaStream_t s = (cudaStream_t)stream;
    const int n = p.n;
    double scale = log10((double)p.count) + log10((double)n) + 31.0 * log10(2.0);       // quisk.c:5284-5285
    scale *= 20.0;
    int nbin = (int)(zoom * (double)n / data_width + 0.5);
    if (nbin < 1) nbin = 1;
    // does any pixel read a bin that an earlier pixel has already overwritten? (in-place walk, quisk.c:5289-5301)
    bool hazard = false;
    for (int i = 0; i < data_width && !hazard; i++) {
        int k = (int)(n * (deltaf / fft_sample_rate + zoom * ((double)i / data_width - 0.5) + 0.5) + 0.1);
        for (int j = 0; j < nbin; j++, k++)
            if (k >= 0 && k < n && k < i) { hazard = true; break; }
    }
    if (!hazard)
        pan_graph_kernel<<<dim3((data_width + 127) / 128, p.S), 128, 0, s>>>(p.d_avg, n, data_width, nbin, zoom, deltaf, fft_sample_rate, scale, d_graph);
    else
        pan_graph_serial_kernel<<<dim3(1, p.S), 32, 0, s>>>(p.d_avg, n, data_width, nbin, zoom, deltaf, fft_sample_rate, scale, d_graph);
    count_launch();
    QC_CUDA_LAUNCH();
    QC_CUDA(cudaMemsetAsync(p.d_avg, 0, (size_t)p.S * n * sizeof(double), s));            // quisk.c:5322-5324
    p.count = 0;
    return QC_OK;
}

int quisk_cuda_pan_multirx(qcPanadapter *pp, const void *d_frames, long stream_stride, double *d_graph, void *stream)
{
    if (!pp) { set_error("pan_multirx: null handle"); return QC_EINVAL; }
    if (pp->p.M || pp->p.n > 8192) { set_error("pan_multirx: needs a power-of-two fft_size <= 8192 (got %d)", pp->p.n); return QC_EINVAL; }
    Panadapter &p = pp->p;
    const int lanes = fft_threads(p.n);
    const size_t sh = ((size_t)p.n + fft_tw_entries(p.n)) * sizeof(cd);
    int rc = fft_smem_optin((const void *)pan_multirx_kernel, sh); if (rc != QC_OK) return rc;
    double scale = (log10((double)p.n) + 31.0 * log10(2.0)) * 20.0;                      // quisk.c:4892-4893
    pan_multirx_kernel<<<p.S, lanes, sh, (cudaStream_t)stream>>>((const cd *)d_frames, stream_stride, p.n, p.tw, p.d_window, scale, d_graph);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

}  // extern "C"
