// quisk_b200/csrc/polyfir.cu -- generic streaming polyphase FIR kernel.
//
// One kernel covers every block function of the reference's filter.h
// (quisk_cDecimate / cCDecimate / dDecimate / cFilter / dFilter filter.c:203-285,347-375;
//  quisk_cInterpolate / dInterpolate filter.c:131-201; quisk_cInterpDecim filter.c:287-324;
//  quisk_cDecim2HB45 filter.c:377-417; quisk_{c,d}Interp2HB45 filter.c:420-488) and the
// per-sample RX filters of quisk.c (cRxFilterOut / dRxFilterOut, quisk.c:1182-1256),
// for C independent channels at once.  See qc_common.cuh for the index algebra.
//
// This is the *exact* path: multiplies and adds are separately rounded and
// visited in the reference's order, so outputs are bit-identical to the
// reference's C code.  It is used by the legacy host-pointer ABI and by the
// unfused batch objects; the fused cascade in rxchain.cu is the fast path.
//
// Layout: a CTA owns TM consecutive outputs of one channel.  For each chunk of
// taps it stages the needed slice of [history | block] and the chunk's taps in
// shared memory (coalesced 16-byte loads), then every thread walks its own
// output's taps out of shared memory.
#include <cstdlib>
#include "qc_common.cuh"

namespace qc {

static constexpr int TM = 256;          // outputs per CTA == threads per CTA

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <typename T> struct Zero;
template <> struct Zero<double> { __device__ static double v() { return 0.0; } };
template <> struct Zero<cd> { __device__ static cd v() { return make_double2(0.0, 0.0); } };

template <typename T>
__device__ __forceinline__ T load_x(const PolyFirParams &p, int c, long j)
{
    if (j >= 0) {
        if (j < p.n_in) return reinterpret_cast<const T *>(p.in)[(long)c * p.in_stride + j];
        return Zero<T>::v();
    }
    if (p.hist != nullptr && j >= -(long)p.H)
        return reinterpret_cast<const T *>(p.hist)[(long)c * p.H + p.H + j];
    return Zero<T>::v();
}

// acc += x (*) coef with the reference's rounding sequence
template <int TAPMODE>
__device__ __forceinline__ void mac(cd &acc, cd x, const double *sc, int idx)
{
    if (TAPMODE == TAP_REAL) {
        double c = sc[idx];
        acc.x = add_rn(acc.x, mul_rn(x.x, c));
        acc.y = add_rn(acc.y, mul_rn(x.y, c));
    } else if (TAPMODE == TAP_COMPLEX) {
        double2 c = reinterpret_cast<const double2 *>(sc)[idx];
        double re = __dsub_rn(mul_rn(x.x, c.x), mul_rn(x.y, c.y));
        double im = add_rn(mul_rn(x.x, c.y), mul_rn(x.y, c.x));
        acc.x = add_rn(acc.x, re);
        acc.y = add_rn(acc.y, im);
    } else {
        double2 c = reinterpret_cast<const double2 *>(sc)[idx];
        acc.x = add_rn(acc.x, mul_rn(x.x, c.x));
        acc.y = add_rn(acc.y, mul_rn(x.y, c.y));
    }
}
template <int TAPMODE>
__device__ __forceinline__ void mac(double &acc, double x, const double *sc, int idx)
{
    acc = add_rn(acc, mul_rn(x, sc[idx]));
}

__device__ __forceinline__ cd scale(cd a, double g) { return make_double2(mul_rn(a.x, g), mul_rn(a.y, g)); }
__device__ __forceinline__ double scale(double a, double g) { return mul_rn(a, g); }
__device__ __forceinline__ cd addv(cd a, cd b) { return make_double2(add_rn(a.x, b.x), add_rn(a.y, b.y)); }
__device__ __forceinline__ double addv(double a, double b) { return add_rn(a, b); }

// filter.c:381-384 -- half-band taps (coef[11] is the centre tap)
__constant__ double c_hb45[12] = {
    0.000018566625444266, -0.000118469698701817, 0.000457318798253456,
    -0.001347840471412094, 0.003321838571445455, -0.007198422696929033,
    0.014211106939802483, -0.026424776824073383, 0.048414810444971007,
    -0.096214669073304823, 0.314881034738348550, 0.500000000000000000 };

template <typename T>
__device__ void write_hist(const PolyFirParams &p, int c)
{
    if (p.hist_out == nullptr) return;
    T *ho = reinterpret_cast<T *>(p.hist_out) + (long)c * p.H;
    for (int i = threadIdx.x; i < p.H; i += blockDim.x)
        ho[i] = load_x<T>(p, c, (long)p.n_in - p.H + i);
}

// Shared-memory slot of staged sample i.  A decimating stage (L = 1, M = 2^psh) reads sX with a lane stride of M
// elements: stored densely every lane of a warp lands in the same bank group (M = 8: a 32-way serialised load, the
// whole cost of the 1121-tap / 8 WDSP resampler).  One pad element per M makes the lane stride M + 1, which is odd.
__device__ __forceinline__ int xslot(int i, int psh) { return psh > 0 ? i + (i >> psh) : i; }

template <typename T, int TAPMODE>
__global__ void __launch_bounds__(TM) polyfir_kernel(PolyFirParams p, int tiles, int KC, int psh)
{
    extern __shared__ double smem[];
    const int c = blockIdx.x / tiles;
    const int tile = blockIdx.x % tiles;
    const int m0 = tile * TM;
    const int tid = threadIdx.x;
    const int m = m0 + tid;
    const bool active = m < p.n_out;

    if (tile == 0) write_hist<T>(p, c);
    if (m0 >= p.n_out) return;

    const int m_last = min(m0 + TM, p.n_out) - 1;
    const long src_lo = (p.u0 + (long)m0 * p.M) / p.L;
    const long src_hi = (p.u0 + (long)m_last * p.M) / p.L;
    const int span = (int)(src_hi - src_lo) + 1;

    long u = p.u0 + (long)m * p.M;
    long src = u / p.L;
    int ph = (int)(u - src * p.L);

    constexpr int CW = (TAPMODE == TAP_REAL) ? 1 : 2;     // doubles per tap
    T *sX = reinterpret_cast<T *>(smem);
    const int sx_elems = xslot(span + KC, psh) + 1;      // >= slots of span + (k1-k0) - 1 samples
    double *sC = reinterpret_cast<double *>(sX + sx_elems);

    T acc = Zero<T>::v();
    int k_begin = 0;
    if (p.order == 1) {
        // cRxFilterOut: the newest sample meets tap 0 first ...
        if (active) {
            T x = load_x<T>(p, c, src);
            // taps live in global memory here; one read per thread
            if (TAPMODE == TAP_REAL) { double cc = p.coef[ph]; mac<TAP_REAL>(acc, x, &cc, 0); }
            else { double2 cc = reinterpret_cast<const double2 *>(p.coef)[ph]; mac<TAPMODE>(acc, x, reinterpret_cast<const double *>(&cc), 0); }
        }
        k_begin = 1;        // ... then taps K-1 down to 1
    }

    const int n_chunks = (p.K - k_begin + KC - 1) / KC;
    for (int ch = 0; ch < n_chunks; ch++) {
        int k0, k1;
        if (p.order == 0) { k0 = k_begin + ch * KC; k1 = min(k0 + KC, p.K); }
        else { k1 = p.K - ch * KC; k0 = max(k1 - KC, k_begin); }
        const long base = src_lo - (k1 - 1);
        const int nx = span + (k1 - k0) - 1;
        __syncthreads();
        for (int i = tid; i < nx; i += TM) sX[xslot(i, psh)] = load_x<T>(p, c, base + i);
        const int nc = (k1 - k0) * p.L * CW;
        const double *gC = p.coef + (long)k0 * p.L * CW;
        for (int i = tid; i < nc; i += TM) sC[i] = gC[i];
        __syncthreads();
        if (active) {
            const int xo = (int)(src - base);             // index of X[src] in sX
            if (p.order == 0) {
                for (int k = k0; k < k1; k++)
                    mac<TAPMODE>(acc, sX[xslot(xo - k, psh)], sC, ph + (k - k0) * p.L);
            } else {
                for (int k = k1 - 1; k >= k0; k--)
                    mac<TAPMODE>(acc, sX[xslot(xo - k, psh)], sC, ph + (k - k0) * p.L);
            }
        }
    }
    if (active)
        reinterpret_cast<T *>(p.out)[(long)c * p.out_stride + m] = scale(acc, p.gain);
}


// ---- register-blocked decimating FIR: complex samples, real taps, L = 1, M a power of two.
// The generic kernel above pays one 16-byte shared-memory load per (output, tap) for four FP64 operations, which pins the
// 1121-tap / 8 WDSP resampler (resample.c:121-157) to half of what the FP64 pipe could do.  Here a thread owns R consecutive
// outputs and walks the SAMPLES it needs from the newest down: sample s meets output r at tap k_r = src_r - s, so one loaded
// sample feeds R accumulators, each of which still receives its products in ascending-k order with separately rounded
// multiply and add -- the same bits as the generic kernel and as the reference.  Taps sit in shared memory behind M (R - 1)
// zeros at both ends, so every (sample, output) pair has a tap (a zero tap adds +0: exact); the R tap loads per step are
// broadcasts.  Samples are staged with one pad element every R M so that the lane stride R M + 1 is odd.
template <int R, int NT>
__global__ void __launch_bounds__(NT) decim_rb_kernel(PolyFirParams p, int tiles, int pad_sh)
{
    extern __shared__ double smem[];
    const int c = blockIdx.x / tiles, tile = blockIdx.x % tiles, tid = threadIdx.x;
    const int M = p.M, K = p.K;
    const int m0 = tile * NT * R;
    if (tile == 0) write_hist<cd>(p, c);
    if (m0 >= p.n_out) return;
    const int zpad = M * (R - 1);
    double *hp = smem;                                          // [K + 2 zpad] zero-padded taps
    const int nh = K + 2 * zpad + 1;                            // + 1: the tap prefetch of the step behind the last one
    cd *sX = reinterpret_cast<cd *>(smem + ((nh + 1) & ~1));
    for (int i = tid; i < nh; i += NT) hp[i] = (i >= zpad && i < zpad + K) ? p.coef[i - zpad] : 0.0;
    const long lo = p.u0 + (long)m0 * M - (K - 1);              // oldest sample the tile touches
    const int nx = (NT * R - 1) * M + K;
    // staging with eight loads in flight per thread (one after the other they would cost as much as the arithmetic)
    if (lo >= 0 && lo + nx <= p.n_in) {
        const cd *g = reinterpret_cast<const cd *>(p.in) + (long)c * p.in_stride + lo;
        for (int i0 = tid; i0 < nx; i0 += 8 * NT) {
            cd v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { const int i = i0 + u * NT; if (i < nx) v[u] = g[i]; }
#pragma unroll
            for (int u = 0; u < 8; u++) { const int i = i0 + u * NT; if (i < nx) sX[i + (i >> pad_sh)] = v[u]; }
        }
    } else {
        for (int i = tid; i < nx; i += NT) sX[i + (i >> pad_sh)] = load_x<cd>(p, c, lo + i);
    }
    __syncthreads();
    cd acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = make_double2(0.0, 0.0);
    const int top = (tid * R + R - 1) * M + (K - 1);            // staged index of this thread's newest sample
    const int J = K + zpad;
    // the next step's sample and taps are fetched while the current step's products are formed
    cd xn = sX[top + (top >> pad_sh)];
    double hn[R];
#pragma unroll
    for (int r = 0; r < R; r++) hn[r] = hp[M * r];
#pragma unroll 2
    for (int j = 0; j < J; j++) {
        const cd x = xn;
        double h[R];
#pragma unroll
        for (int r = 0; r < R; r++) h[r] = hn[r];
        const int nxt = max(top - j - 1, 0);                 // (the fetch behind the last step is never used)
        xn = sX[nxt + (nxt >> pad_sh)];
#pragma unroll
        for (int r = 0; r < R; r++) hn[r] = hp[j + 1 + M * r];
        // all products first, then all sums: a sum right behind its own product would wait out the multiply's latency
        cd t[R];
#pragma unroll
        for (int r = 0; r < R; r++) t[r] = make_double2(mul_rn(x.x, h[r]), mul_rn(x.y, h[r]));
#pragma unroll
        for (int r = 0; r < R; r++) { acc[r].x = add_rn(acc[r].x, t[r].x); acc[r].y = add_rn(acc[r].y, t[r].y); }
    }
    cd *o = reinterpret_cast<cd *>(p.out) + (long)c * p.out_stride + m0 + tid * R;
#pragma unroll
    for (int r = 0; r < R; r++) if (m0 + tid * R + r < p.n_out) o[r] = scale(acc[r], p.gain);
}

// ---- the same for M = 8, R = 4 with the taps in registers.  Output r meets the sample of step j at tap hp[j + 8 r], so the
// taps of step j + 8 are the taps of step j moved down by one output: for each of the 8 residues s = j mod 8 a window of
// R = 4 taps slides by one per 8 steps.  With the 8 x 4 window in registers (the loop body is 32 steps, every index static)
// a step costs ONE shared-memory tap load (broadcast) instead of four; with the sample load that is 5 wavefronts per warp
// and step instead of 8 -- at R = 4 the generic form above asks shared memory for exactly its peak (32 wavefronts per 32
// cycles and SM when the FP64 pipe runs flat out: 16 separately rounded instructions per step and lane).  Same products,
// same sums, same order: bit-identical.  The step count is padded to a multiple of 32 with zero taps (they add +0).
template <int NT>
__global__ void __launch_bounds__(NT, 2) decim_rb8_kernel(PolyFirParams p, int tiles, int pad_sh, int Jpad)
{
    constexpr int R = 4, M = 8;
    extern __shared__ double smem[];
    const int c = blockIdx.x / tiles, tile = blockIdx.x % tiles, tid = threadIdx.x;
    const int K = p.K;
    const int m0 = tile * NT * R;
    if (tile == 0) write_hist<cd>(p, c);
    if (m0 >= p.n_out) return;
    constexpr int zpad = M * (R - 1);
    double *hp = smem;                                          // [Jpad + M R + 8] zero-padded taps
    const int nh = Jpad + M * R + 8;
    cd *sX = reinterpret_cast<cd *>(smem + ((nh + 1) & ~1));
    for (int i = tid; i < nh; i += NT) hp[i] = (i >= zpad && i < zpad + K) ? p.coef[i - zpad] : 0.0;
    const long lo = p.u0 + (long)m0 * M - (K - 1);
    const int nx = (NT * R - 1) * M + K;
    if (lo >= 0 && lo + nx <= p.n_in) {
        const cd *g = reinterpret_cast<const cd *>(p.in) + (long)c * p.in_stride + lo;
        for (int i0 = tid; i0 < nx; i0 += 8 * NT) {
            cd v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { const int i = i0 + u * NT; if (i < nx) v[u] = g[i]; }
#pragma unroll
            for (int u = 0; u < 8; u++) { const int i = i0 + u * NT; if (i < nx) sX[i + (i >> pad_sh)] = v[u]; }
        }
    } else {
        for (int i = tid; i < nx; i += NT) sX[i + (i >> pad_sh)] = load_x<cd>(p, c, lo + i);
    }
    __syncthreads();
    cd acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = make_double2(0.0, 0.0);
    const int top = (tid * R + R - 1) * M + (K - 1);
    double W[M][R];                                             // W[s][(q + r) & 3] = hp[8 q + s + 8 r]
#pragma unroll
    for (int s_ = 0; s_ < M; s_++)
#pragma unroll
        for (int r = 0; r < R; r++) W[s_][r] = hp[s_ + M * r];
    for (int j0 = 0; j0 < Jpad; j0 += M * R) {
#pragma unroll
        for (int qq = 0; qq < R; qq++) {
#pragma unroll
            for (int s_ = 0; s_ < M; s_++) {
                const int j = j0 + qq * M + s_;
                const int idx = max(top - j, 0);
                const cd x = sX[idx + (idx >> pad_sh)];
                cd t[R];
#pragma unroll
                for (int r = 0; r < R; r++) { const double h = W[s_][(qq + r) & 3]; t[r] = make_double2(mul_rn(x.x, h), mul_rn(x.y, h)); }
#pragma unroll
                for (int r = 0; r < R; r++) { acc[r].x = add_rn(acc[r].x, t[r].x); acc[r].y = add_rn(acc[r].y, t[r].y); }
                W[s_][qq & 3] = hp[j + M * R];                  // output R - 1's tap of step j + 8
            }
        }
    }
    cd *o = reinterpret_cast<cd *>(p.out) + (long)c * p.out_stride + m0 + tid * R;
#pragma unroll
    for (int r = 0; r < R; r++) if (m0 + tid * R + r < p.n_out) o[r] = scale(acc[r], p.gain);
}

static int launch_decim_rb(const PolyFirParams &p, cudaStream_t stream)
{
    constexpr int R = 4, NT = 128;
    const int tiles = (p.n_out + NT * R - 1) / (NT * R);
    const long grid = (long)tiles * p.C;
    if (grid <= 0 || grid > 0x7fffffffL) { set_error("polyfir: grid %ld out of range", grid); return QC_EINVAL; }
    int pad_sh = 0;
    while ((1 << pad_sh) < R * p.M) pad_sh++;
    const int nx = (NT * R - 1) * p.M + p.K;
    if (p.M == 8 && !getenv("QUISK_DECIM_RB_GENERIC")) {
        const int J = p.K + p.M * (R - 1), Jpad = (J + 31) & ~31;
        const int nh8 = Jpad + 8 * R + 8;
        const size_t sh8 = (size_t)((nh8 + 1) & ~1) * sizeof(double) + (size_t)(nx + (nx >> pad_sh) + 1) * sizeof(cd);
        if (sh8 <= 113 * 1024) {
            auto k8 = decim_rb8_kernel<NT>;
            if (sh8 > 48 * 1024) {
                static std::mutex mu8;
                std::lock_guard<std::mutex> g(mu8);
                QC_CUDA(cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024));
            }
            k8<<<(unsigned)grid, NT, sh8, stream>>>(p, tiles, pad_sh, Jpad);
            count_launch();
            QC_CUDA_LAUNCH();
            return QC_OK;
        }
    }
    const int nh = p.K + 2 * p.M * (R - 1) + 1;
    const size_t sh = (size_t)((nh + 1) & ~1) * sizeof(double) + (size_t)(nx + (nx >> pad_sh) + 1) * sizeof(cd);
    if (sh > 200 * 1024) return -100;                            // tile too large for shared memory: the generic kernel takes it
    auto kern = decim_rb_kernel<R, NT>;
    if (sh > 48 * 1024) {
        static std::mutex mu;
        std::lock_guard<std::mutex> g(mu);
        QC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    kern<<<(unsigned)grid, NT, sh, stream>>>(p, tiles, pad_sh);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

// Half-band forms with the reference's paired summation (filter.c:401-413, 444-451).
template <typename T>
__global__ void __launch_bounds__(TM) hb45_kernel(PolyFirParams p, int tiles)
{
    extern __shared__ double smem[];
    const int c = blockIdx.x / tiles;
    const int tile = blockIdx.x % tiles;
    const int tid = threadIdx.x;
    T *sX = reinterpret_cast<T *>(smem);
    if (tile == 0) write_hist<T>(p, c);

    if (p.hb_mode == HB_DECIM) {
        // output m at input n = u0 + 2m; uses X[n-42 .. n]
        const int m0 = tile * TM;
        if (m0 >= p.n_out) return;
        const int cnt = min(TM, p.n_out - m0);
        const long base = p.u0 + 2L * m0 - 42;
        const int nx = 2 * (cnt - 1) + 43;
        for (int i = tid; i < nx; i += TM) sX[i] = load_x<T>(p, c, base + i);
        __syncthreads();
        if (tid < cnt) {
            const int xo = 2 * tid + 42;                 // X[n]
            // samples[k] = X[n-2k], center[10] = X[n-21]
            T acc = scale(addv(sX[xo], sX[xo - 42]), c_hb45[0]);
#pragma unroll
            for (int k = 1; k < 11; k++)
                acc = addv(acc, scale(addv(sX[xo - 2 * k], sX[xo - 42 + 2 * k]), c_hb45[k]));
            acc = addv(acc, scale(sX[xo - 21], c_hb45[11]));
            reinterpret_cast<T *>(p.out)[(long)c * p.out_stride + m0 + tid] = acc;
        }
    } else {
        // input i -> outputs 2i, 2i+1; uses X[i-21 .. i]; n_out counts outputs (even)
        const int i0 = tile * TM;
        const int n_pairs = p.n_out / 2;
        if (i0 >= n_pairs) return;
        const int cnt = min(TM, n_pairs - i0);
        const long base = (long)i0 - 21;
        const int nx = cnt + 21;
        for (int i = tid; i < nx; i += TM) sX[i] = load_x<T>(p, c, base + i);
        __syncthreads();
        if (tid < cnt) {
            const int xo = tid + 21;                     // X[i] = samples[0]
            T o0 = scale(scale(sX[xo - 11], c_hb45[11]), 2.0);
            T acc = scale(addv(sX[xo], sX[xo - 21]), c_hb45[0]);
#pragma unroll
            for (int k = 1; k < 11; k++)
                acc = addv(acc, scale(addv(sX[xo - k], sX[xo - 21 + k]), c_hb45[k]));
            T *o = reinterpret_cast<T *>(p.out) + (long)c * p.out_stride + 2L * (i0 + tid);
            o[0] = o0;
            o[1] = scale(acc, 2.0);
        }
    }
}

template <typename T, int TAPMODE>
static int launch_t(const PolyFirParams &p, cudaStream_t stream)
{
    const int tiles = p.n_out > 0 ? (p.n_out + TM - 1) / TM : 1;
    const long grid = (long)tiles * p.C;
    if (grid <= 0 || grid > 0x7fffffffL) { set_error("polyfir: grid %ld out of range", grid); return QC_EINVAL; }
    const int CW = (TAPMODE == TAP_REAL) ? 1 : 2;
    // span of source samples one tile touches
    long span = ((long)(TM - 1) * p.M) / p.L + 2;
    int KC = p.K < 512 ? p.K : 512;
    if (KC < 1) KC = 1;
    int psh = 0;                                        // pad for power-of-two decimation (see xslot)
    if (p.L == 1 && p.M >= 2 && (p.M & (p.M - 1)) == 0) while ((1 << psh) < p.M) psh++;
    size_t sh;
    for (;;) {
        const long slots = span + KC + (psh > 0 ? ((span + KC) >> psh) : 0) + 1;
        sh = (size_t)slots * sizeof(T) + (size_t)KC * p.L * CW * sizeof(double);
        if (sh <= 200 * 1024 || KC <= 16) break;
        KC /= 2;
    }
    if (sh > 227 * 1024) { set_error("polyfir: tile needs %zu bytes of shared memory", sh); return QC_EINVAL; }
    auto kern = polyfir_kernel<T, TAPMODE>;
    if (sh > 48 * 1024) {
        static std::mutex mu;
        std::lock_guard<std::mutex> g(mu);
        QC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    kern<<<(unsigned)grid, TM, sh, stream>>>(p, tiles, KC, psh);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

template <typename T>
static int launch_hb(const PolyFirParams &p, cudaStream_t stream)
{
    const int units = p.hb_mode == HB_DECIM ? p.n_out : p.n_out / 2;
    const int tiles = units > 0 ? (units + TM - 1) / TM : 1;
    const long grid = (long)tiles * p.C;
    if (grid <= 0 || grid > 0x7fffffffL) { set_error("hb45: grid %ld out of range", grid); return QC_EINVAL; }
    size_t sh = (size_t)(2 * TM + 64) * sizeof(T);
    hb45_kernel<T><<<(unsigned)grid, TM, sh, stream>>>(p, tiles);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

int launch_polyfir(const PolyFirParams &p, cudaStream_t stream)
{
    if (p.C <= 0) return QC_OK;
    if (p.hb_mode != HB_NONE)
        return p.is_complex ? launch_hb<cd>(p, stream) : launch_hb<double>(p, stream);
    if (p.K < 1 || p.L < 1 || p.M < 1) { set_error("polyfir: bad K/L/M %d/%d/%d", p.K, p.L, p.M); return QC_EINVAL; }
    if (p.is_complex) {
        if (p.tap_mode == TAP_REAL && p.L == 1 && p.order == 0 && p.K >= 64 && p.M <= 16 && (p.M & (p.M - 1)) == 0 && p.n_out >= 256) {
            const int rc = launch_decim_rb(p, stream);
            if (rc != -100) return rc;
        }
        switch (p.tap_mode) {
        case TAP_REAL: return launch_t<cd, TAP_REAL>(p, stream);
        case TAP_COMPLEX: return launch_t<cd, TAP_COMPLEX>(p, stream);
        case TAP_SPLIT_IQ: return launch_t<cd, TAP_SPLIT_IQ>(p, stream);
        }
    } else if (p.tap_mode == TAP_REAL) {
        return launch_t<double, TAP_REAL>(p, stream);
    }
    set_error("polyfir: unsupported sample/tap combination");
    return QC_EINVAL;
}

}  // namespace qc
