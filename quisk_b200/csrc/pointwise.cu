// quisk_b200/csrc/pointwise.cu -- per-sample stages of the receive chain that are not FIR
// filters: the tuning NCO (quisk.c:2477-2488), the SSB/CW sideband combine
// (quisk.c:1916,1939,1962,1986), the AM envelope detector with DC blocker
// (quisk.c:2005-2011) and the FM discriminator with one-pole de-emphasis
// (quisk.c:2029-2064).  Used by the unfused chain; the fused cascade in
// rxchain.cu has these inlined.
#include "qc_common.cuh"
#include "nco_device.cuh"

namespace qc {

// v at the start of the next block by the reference's own rounded recurrence (quisk.c:2483: v *= phase, `count`
// times): one thread per channel, a dependent chain of count complex multiplies.  It runs on a side stream one
// block ahead of the kernels that consume it (rxchain.cu), so its latency is hidden; the closed form (nco_pow)
// then only has to bridge the samples INSIDE one block, and the tuning phasor follows the reference for any
// stream length.
// Work distribution: the recurrence is one dependent FP64 chain per channel, so its speed is set by how many of its
// warps share an SM sub-partition's FP64 pipe -- one per sub-partition: 20 cycles per step; nine: 108.  The block
// scheduler gives no control over that (it packs small CTAs onto whatever SMs have room at that instant: beside two
// resident decimator CTAs that is one per SM, on SMs that are just draining it is nine), and a packed launch is
// slower than the decimator it hides under, stalls the next block and leaves the device in a state where the next
// launch is packed again (seen on 8-GPU runs: one rank at 1.95 ms per step instead of 1.24).  So the CTAs elect ONE
// worker per SM themselves: the grid has 2 x #SM small CTAs; the first CTA of this launch to arrive on an SM (atomicMax
// of the launch epoch on that SM's slot) becomes its worker and pulls groups of 128 channels from a ticket counter
// until none are left, every other CTA exits at once.
__global__ void nco_advance_kernel(const cd *v_in, cd *v_out, const double *nco, int count, int C, unsigned *sched, unsigned epoch)
{
    // sched[0]: ticket counter (zeroed on the stream before the launch); sched[1 + smid]: last epoch with a worker there
    __shared__ int s_item;
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_item = atomicMax(&sched[1 + (smid & 1023u)], epoch) < epoch ? 0 : -1;
    }
    __syncthreads();
    if (s_item < 0) return;
    const int n_items = (C + (int)blockDim.x - 1) / (int)blockDim.x;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = (int)atomicAdd(&sched[0], 1u);
        __syncthreads();
        const int item = s_item;
        if (item >= n_items) return;
        const int c = item * (int)blockDim.x + (int)threadIdx.x;
        if (c < C) {
            const cd ph = make_double2(nco[(size_t)c * 8 + 5], nco[(size_t)c * 8 + 6]);
            cd v = v_in[c];
#pragma unroll 4
            for (int i = 0; i < count; i++) v = cmul_rn(v, ph);
            v_out[c] = v;
        }
    }
}

int launch_nco_advance(const cd *v_in, cd *v_out, const double *d_nco, int count, int C, unsigned *d_sched, unsigned epoch, cudaStream_t s)
{
    if (C <= 0) return QC_OK;
    // 24 KB of (unused) dynamic shared memory per CTA: beside the two 99 KB CTAs of the fused decimator exactly one of
    // these fits on an SM.  128 threads: the four warps of a worker sit on the four sub-partitions of ONE SM, so every
    // sub-partition of that SM gives up the same share of its FP64 pipe (one-warp CTAs put the whole load on a single
    // sub-partition, and the barriers of the decimator CTAs living there make their other warps wait for it: 12 % of
    // the step that way, the recurrence's FP64 work is 3.6 % of it).
    static bool optin[64] = {};         // per device: function attributes belong to the device's context
    static int n_sm[64] = {};
    int dev = 0; cudaGetDevice(&dev); dev &= 63;
    if (!optin[dev]) {
        QC_CUDA(cudaFuncSetAttribute(nco_advance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 1024));
        QC_CUDA(cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev));
        optin[dev] = true;
    }
    QC_CUDA(cudaMemsetAsync(d_sched, 0, sizeof(unsigned), s));
    const int items = (C + 127) / 128;
    const int grid = items < 2 * n_sm[dev] ? 2 * n_sm[dev] : items;     // enough CTAs that every SM sees one
    nco_advance_kernel<<<grid, 128, 24 * 1024, s>>>(v_in, v_out, d_nco, count, C, d_sched, epoch);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

// closed-form counterpart: v_out = v_in * phase^count in one step (QC_RX_OPT_EXACT_NCO = 0)
__global__ void nco_jump_kernel(const cd *v_in, cd *v_out, const double *nco, int count, int C)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    v_out[c] = cmul_rn(v_in[c], nco_pow(nco + (size_t)c * 8, (unsigned long long)count));
}

int launch_nco_jump(const cd *v_in, cd *v_out, const double *d_nco, int count, int C, cudaStream_t s)
{
    if (C <= 0) return QC_OK;
    nco_jump_kernel<<<(C + 127) / 128, 128, 0, s>>>(v_in, v_out, d_nco, count, C);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

__global__ void tune_kernel(const cd *in, long in_stride, cd *out, long out_stride, int n, int C,
                            const double *nco, const cd *vstart, unsigned long long n0)
{
    const long total = (long)n * C;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx / n);
        const int j = (int)(idx - (long)c * n);
        const double *nc = nco + (long)c * 8;
        const cd w = nco_pow(nc, n0 + (unsigned long long)j);
        const cd v = cmul_rn(vstart[c], w);
        const cd x = in[(long)c * in_stride + j];
        out[(long)c * out_stride + j] = cmul_rn(x, v);
    }
}

int launch_tune(const cd *in, long in_stride, cd *out, long out_stride, int n, int C,
                const double *d_nco, const cd *d_vstart, unsigned long long n0, cudaStream_t s)
{
    if (n <= 0 || C <= 0) return QC_OK;
    const long total = (long)n * C;
    int blocks = (int)((total + 255) / 256 < 148L * 16 ? (total + 255) / 256 : 148L * 16);
    tune_kernel<<<blocks, 256, 0, s>>>(in, in_stride, out, out_stride, n, C, d_nco, d_vstart, n0);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

__global__ void demod_ssb_kernel(const cd *in, long in_stride, double *out, long out_stride, int n, int C, int lower)
{
    const long total = (long)n * C;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx / n);
        const int j = (int)(idx - (long)c * n);
        const cd x = in[(long)c * in_stride + j];
        out[(long)c * out_stride + j] = lower ? __dadd_rn(x.x, x.y) : __dsub_rn(x.x, x.y);
    }
}

// rows (receivers) whose squelch flag state[c][1] is set come back as zeros (quisk_process_samples, quisk.c:2716-2719)
__global__ void mute_rows_kernel(double *audio, long stride, int n, int C, const int *state)
{
    const int c = blockIdx.y;
    if (!state[2 * c + 1]) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) audio[(long)c * stride + i] = 0.0;
}

int launch_mute_rows(double *audio, long stride, int n, int C, const int *d_state, cudaStream_t s)
{
    if (n <= 0 || C <= 0 || !d_state) return QC_OK;
    mute_rows_kernel<<<dim3((n + 255) / 256 < 8 ? (n + 255) / 256 : 8, C), 256, 0, s>>>(audio, stride, n, C, d_state);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

int launch_demod_ssb(const cd *in, long in_stride, double *out, long out_stride, int n, int C, int lower, cudaStream_t s)
{
    if (n <= 0 || C <= 0) return QC_OK;
    const long total = (long)n * C;
    int blocks = (int)((total + 255) / 256 < 148L * 16 ? (total + 255) / 256 : 148L * 16);
    demod_ssb_kernel<<<blocks, 256, 0, s>>>(in, in_stride, out, out_stride, n, C, lower);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

// One thread per channel: the DC blocker is a scalar recurrence (d = |x| + 0.99 dc; out = d - dc; dc = d).
__global__ void am_detect_kernel(const cd *in, long in_stride, double *out, long out_stride, int n, int C, double *dc_state)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double dc = dc_state[c];
    const cd *x = in + (long)c * in_stride;
    double *o = out + (long)c * out_stride;
    for (int i = 0; i < n; i++) {
        const double di = hypot(x[i].x, x[i].y);
        const double d = __dadd_rn(di, __dmul_rn(dc, 0.99));
        o[i] = __dsub_rn(d, dc);
        dc = d;
    }
    dc_state[c] = dc;
}

int launch_am_detect(const cd *in, long in_stride, double *out, long out_stride, int n, int C, double *d_dc, cudaStream_t s)
{
    if (n <= 0 || C <= 0) return QC_OK;
    am_detect_kernel<<<(C + 63) / 64, 64, 0, s>>>(in, in_stride, out, out_stride, n, C, d_dc);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

// One thread per channel: di = carg(x conj(x_-1)) * 20e5; y = di a0 + x_1 a1 - y_1 b1.
__global__ void fm_detect_kernel(const cd *in, long in_stride, double *out, long out_stride, int n, int C,
                                 double *state, double a0, double a1, double b1)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double *st = state + (long)c * 4;
    cd prev = make_double2(st[0], st[1]);
    double x1 = st[2], y1 = st[3];
    const cd *x = in + (long)c * in_stride;
    double *o = out + (long)c * out_stride;
    for (int i = 0; i < n; i++) {
        const cd cx = x[i];
        // cx * conj(prev)
        const double re = __dadd_rn(__dmul_rn(cx.x, prev.x), __dmul_rn(cx.y, prev.y));
        const double im = __dsub_rn(__dmul_rn(cx.y, prev.x), __dmul_rn(cx.x, prev.y));
        prev = cx;
        const double di = __dmul_rn(atan2(im, re), 20e5);
        y1 = __dsub_rn(__dadd_rn(__dmul_rn(di, a0), __dmul_rn(x1, a1)), __dmul_rn(y1, b1));
        x1 = di;
        o[i] = y1;
    }
    st[0] = prev.x; st[1] = prev.y; st[2] = x1; st[3] = y1;
}

int launch_fm_detect(const cd *in, long in_stride, double *out, long out_stride, int n, int C,
                     double *d_state, double a0, double a1, double b1, cudaStream_t s)
{
    if (n <= 0 || C <= 0) return QC_OK;
    fm_detect_kernel<<<(C + 63) / 64, 64, 0, s>>>(in, in_stride, out, out_stride, n, C, d_state, a0, a1, b1);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

}  // namespace qc
