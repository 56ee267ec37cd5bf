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
// Two implementations:
//  (1) D = K or D = K/2 (every case BASELINE.json names): two kernels per slice of frames.
//      pfb_fir_kernel  : one THREAD per branch r and frame range.  The branch's taps (T/K per tap alignment, two
//                        alignments when D = K/2) and its T/K-sample window live in registers; per step the thread
//                        loads ONE new sample (coalesced across the 128 branches of the CTA, prefetched T/K steps
//                        ahead), and emits the one or two frames that sample completes: u[frame][r].  Input is read
//                        once (plus a T/K-sample warm-up per frame range), no shared memory at all.
//      pfb_fft_kernel  : K-point transforms of u, four frames per CTA, written channel-major or frame-major.
//      u makes a round trip through HBM (32 B written + 32 B read per input sample on top of the 48 algorithmic
//      bytes).  Slicing the frames so that u would stay in the 126 MB L2 -- sequentially or with the two kernels
//      of neighbouring slices overlapped on two streams (QC_PFB_OPT_SLICE_FRAMES / _PIPELINE) -- was measured
//      SLOWER at every slice size (launch gaps and tails outweigh the saved traffic), so the default is one slice.
//  (2) any other D <= K: the single generic kernel below (taps in shared memory, window re-read per round).
//
// Generic kernel mapping: a CTA of K/4 threads = 4 transforms x K/16 lanes handles FOUR consecutive frames per round.
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
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

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

// ------------------------------------------------------------------------------------------------------------
// (1) register-resident branch FIRs + batched transforms, D = K / OVS with OVS = 1 or 2
// ------------------------------------------------------------------------------------------------------------
struct PfbFirParams {
    const cd *in; int count;
    const cd *hist; int H;
    long long n0;                   // absolute index of in[0]
    long long f0; int nf;           // absolute index of the first frame of this launch, frames in it
    int fs;                         // frames per CTA range (even)
    int K, D;
    const double *taps;
    cd *u;                          // [nf][K]
};

// sample at index rel relative to in[0]: the block, the history in front of it, or zero outside both
__device__ __forceinline__ cd pfb_load(const cd *__restrict__ in, const cd *__restrict__ hist_end, int rel, int count, int H)
{
    cd v = make_double2(0.0, 0.0);
    if (rel >= 0) { if (rel < count) v = in[rel]; }
    else if (rel >= -H) v = hist_end[rel];
    return v;
}

// Branch r sees the sub-stream x_r[a] = x[K a + r].  Frame m ends on sample n_m = D m + D - 1 and uses
// x_r[a_m - a'] h[t0 + K a'], a' < P, with a_m = floor((n_m - r) / K) and t0 = (n_m - r) mod K.
//   OVS = 1 (D = K):   a_m = m, t0 = K - 1 - r: one frame per new sample, one tap set.
//   OVS = 2 (D = K/2): sample a completes frames m_lo = 2a + (r >= D) with t0 = (2D - 1 - r) mod D, and m_lo + 1 with
//                      t0 + D: two frames per new sample on the same window, two tap sets.
// A branch is shared by TWO lanes of a warp (l and l + 16): lane half h holds the taps and the window for ages
// [h P/2, (h+1) P/2).  Per step the young half loads the new sample, the sample that ages out of its window moves
// to the old half by shuffle, both halves run P/2 taps, and one xor-shuffle adds the halves; the young half then
// stores frame m_lo, the old half frame m_lo + 1.  That halves the registers per thread (taps + window are the
// whole footprint), which is what sets the number of resident warps here.
// All loop arithmetic is 32-bit and relative to the launch (count <= 2^30 is checked by the host).
static constexpr int PFB_BR = 64;       // branches per 128-thread CTA
static constexpr int PFB_RING = 16;     // cp.async ring depth in steps (power of two)

template <int P, int OVS, int RING>
__global__ void __launch_bounds__(128, (P >= 32 ? 3 : 4)) pfb_fir_kernel(PfbFirParams p)
{
    constexpr int PH = P / 2;
    const int lane = threadIdx.x & 31, half = lane >> 4;
    const int r = blockIdx.x * PFB_BR + (threadIdx.x >> 5) * 16 + (lane & 15);
    const int K = p.K, D = p.D;
    const int ra = blockIdx.y * p.fs;                                          // this CTA's frames [ra, rb) of the launch
    const int rb = min(ra + p.fs, p.nf);
    if (ra >= rb) return;
    const int tau = OVS == 2 ? (2 * D - 1 - r) % D : K - 1 - r;
    double lo[PH], hi[OVS == 2 ? PH : 1];
#pragma unroll
    for (int a = 0; a < PH; a++) {
        lo[a] = p.taps[tau + K * (half * PH + a)];
        if (OVS == 2) hi[a] = p.taps[tau + D + K * (half * PH + a)];
    }
    // first and last step (absolute branch sample index), and the launch-relative frame the first step completes
    const long long fa = p.f0 + ra, fb = p.f0 + rb;
    long long a_lo, a_hi;
    if (OVS == 2) { a_lo = (fa >> 1) - 1; a_hi = (fb - 1) >> 1; }
    else { a_lo = fa; a_hi = fb - 1; }
    const int steps = (int)(a_hi - a_lo + 1);
    const int mrel0 = OVS == 2 ? (int)(2 * a_lo - p.f0) + (r >= D ? 1 : 0) : (int)(a_lo - p.f0);
    const int rel = (int)(a_lo * K + r - p.n0);                                // of step 0's sample
    const cd *in = p.in, *he = p.hist + p.H;
    cd w[PH];                                                                  // slot i mod PH: this half's sample of step i
#pragma unroll
    for (int q = 1; q < PH; q++) w[PH - q] = pfb_load(in, he, rel - (q + half * PH) * K, p.count, p.H);
    w[0] = pfb_load(in, he, rel - (PH + half * PH) * K, p.count, p.H);          // ages out at step 0: what the old half takes then
    // New samples arrive through a per-lane cp.async ring RING steps deep: what hides the HBM latency is bytes in
    // flight (16 warps x 16 lanes x 16 B x RING = 64 KB per SM), and shared memory holds them without registers.
    // A lane only ever reads what it copied itself, so cp.async.wait_group is the only synchronisation.
    __shared__ cd ring[RING][PFB_BR];
    const int rl = (threadIdx.x >> 5) * 16 + (lane & 15);
    auto issue = [&](int step) {
        if (!half) {
            const int e = rel + step * K;
            const cd *src = in; unsigned bytes = 0;
            if (step < steps) {
                if (e >= 0) { if (e < p.count) { src = in + e; bytes = 16; } }
                else if (e >= -p.H) { src = he + e; bytes = 16; }
            }
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&ring[step & (RING - 1)][rl]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll 1
    for (int j = 0; j < RING - 2; j++) issue(j);
    cd *uo = p.u + r;
    for (int i = 0; i < steps; i += PH) {
#pragma unroll
        for (int j = 0; j < PH; j++) {
            if (i + j < steps) {                                                // uniform over the CTA
                const cd old = w[j];                                            // age PH of the young half = age 0 of the old one
                const double ox = __shfl_sync(0xffffffffu, old.x, lane & 15), oy = __shfl_sync(0xffffffffu, old.y, lane & 15);
                // the slot this overwrites was read two steps ago and that value has been consumed since
                issue(i + j + RING - 2);
                asm volatile("cp.async.wait_group %0;" ::"n"(RING - 2) : "memory");
                w[j] = half ? make_double2(ox, oy) : ring[(i + j) & (RING - 1)][rl];
                double xr = 0.0, xi = 0.0, yr = 0.0, yi = 0.0;
#pragma unroll
                for (int t = 0; t < PH; t++) {
                    const cd x = w[(j - t + PH) % PH];
                    xr = fma(x.x, lo[t], xr); xi = fma(x.y, lo[t], xi);
                    if (OVS == 2) { yr = fma(x.x, hi[t], yr); yi = fma(x.y, hi[t], yi); }
                }
                // each half finishes ONE frame (young half: m_lo, old half: m_lo + 1), so it only needs the partner's
                // partial sum of that frame: one exchange of a complex value instead of two
                const double sr = OVS == 2 ? (half ? xr : yr) : xr, si = OVS == 2 ? (half ? xi : yi) : xi;
                const double pr = __shfl_xor_sync(0xffffffffu, sr, 16), pi = __shfl_xor_sync(0xffffffffu, si, 16);
                const int m = mrel0 + OVS * (i + j) + half;
                if (OVS == 2) { if (m >= ra && m < rb) uo[(size_t)m * K] = half ? make_double2(yr + pr, yi + pi) : make_double2(xr + pr, xi + pi); }
                else if (!half && m >= ra && m < rb) uo[(size_t)m * K] = make_double2(xr + pr, xi + pi);
            }
        }
    }
}

// K-point forward transforms of u, K = 256 R3 (R3 = 1, 2, 4), as 16 x 16 x R3 Stockham passes with FOUR frames
// interleaved: thread (b, f) = (tid / 4, tid % 4) runs butterfly b of frame 4 g + f, and element i of frame f lives at
// shared-memory slot 4 (i ^ ((i >> 4) & 1)) + f -- the four frames of one element are 64 contiguous bytes, and the XOR
// puts the elements of neighbouring butterflies into different halves of a 128-byte row in every pass.
//   pass 1: inputs straight from global memory (128-byte runs per frame), radix 16, twiddle, to shared memory
//   pass 2: radix 16 in shared memory (for K = 256 the outputs are final and go to global memory)
//   pass 3: 16 / R3 radix-R3 butterflies per thread, no twiddles, results from registers to global memory:
//           channel-major, the four frames of a channel are one 64-byte run; frame-major, 128-byte runs per frame.
// One shared-memory round trip per pass boundary and three barriers per four transforms.
template <int R3, int PFI>
__global__ void __launch_bounds__(64 * PFI, 8 / PFI) pfb_fft_kernel(const cd *__restrict__ u, int nf, const cd *__restrict__ tw,
                                                                 cd *__restrict__ out, long out_stride, long frame0, int layout, int pf_ahead)
{
    constexpr int K = 256 * R3, N1 = K / 16;            // N1 = lanes per transform; PFI = frames interleaved per CTA (2 or 4)
    extern __shared__ double smem_raw[];
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *sb = twl + fft_tw_entries(K);                   // [K][PFI]
    const int tid = threadIdx.x, f = tid % PFI, b = tid / PFI;
    const int g = blockIdx.x;
    const bool live = g * PFI + f < nf;                  // ragged last group: the thread still takes part in barriers
    fft_stage_twiddles(twl, tw, K);
    auto slot = [f](int i) { return PFI * (i ^ ((i >> 4) & (8 / PFI - 1))) + f; };
    cd v[16];
    // ---- pass 1: len K, stride 1, butterfly p = b
    {
        const cd *src = u + ((size_t)g * PFI + f) * K + b;
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = live ? src[j * N1] : make_double2(0.0, 0.0);
        // pull the group a later CTA will work on towards L2 (this kernel is not persistent: by the time the CTA that
        // takes over this slot starts, its 16 loads per thread find the lines in L2 instead of HBM)
        {
            const long gn = (long)g + pf_ahead;
            if (pf_ahead > 0 && (gn + 1) * PFI <= nf) {
                const char *nxt = reinterpret_cast<const char *>(u + (size_t)gn * PFI * K);
                for (int l = tid; l < (int)(PFI * K * sizeof(cd) / 128); l += blockDim.x)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + (size_t)l * 128));
            }
        }
        dft16(v, -1.0);
        __syncthreads();                                // twiddle store staged
        if (b != 0) {
            const cd w1 = fft_tw(twl, b, -1), w2 = fft_tw(twl, 2 * b, -1), w4 = fft_tw(twl, 4 * b, -1), w8 = fft_tw(twl, 8 * b, -1);
            v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[4] = cmul(v[4], w4); v[8] = cmul(v[8], w8);
            const cd w3 = cmul(w1, w2); v[3] = cmul(v[3], w3);
            const cd w5 = cmul(w1, w4); v[5] = cmul(v[5], w5);
            const cd w6 = cmul(w2, w4); v[6] = cmul(v[6], w6);
            const cd w7 = cmul(w3, w4); v[7] = cmul(v[7], w7);
            v[9] = cmul(v[9], cmul(w1, w8)); v[10] = cmul(v[10], cmul(w2, w8)); v[11] = cmul(v[11], cmul(w3, w8));
            v[12] = cmul(v[12], cmul(w4, w8)); v[13] = cmul(v[13], cmul(w5, w8)); v[14] = cmul(v[14], cmul(w6, w8));
            v[15] = cmul(v[15], cmul(w7, w8));
        }
#pragma unroll
        for (int k = 0; k < 16; k++) sb[slot(16 * b + k)] = v[k];
    }
    __syncthreads();
    // ---- pass 2: len K/16, stride 16, butterfly (p, q) = (b / 16, b % 16)
    {
        const int p = b >> 4, q = b & 15;
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = sb[slot(q + 16 * (p + R3 * j))];
        dft16(v, -1.0);
        if (R3 > 1 && p != 0) {
            const int pt = p * 16;                      // index step K / len = 16
            const cd w1 = fft_tw(twl, pt, -1), w2 = fft_tw(twl, 2 * pt, -1), w4 = fft_tw(twl, 4 * pt, -1), w8 = fft_tw(twl, 8 * pt, -1);
            v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[4] = cmul(v[4], w4); v[8] = cmul(v[8], w8);
            const cd w3 = cmul(w1, w2); v[3] = cmul(v[3], w3);
            const cd w5 = cmul(w1, w4); v[5] = cmul(v[5], w5);
            const cd w6 = cmul(w2, w4); v[6] = cmul(v[6], w6);
            const cd w7 = cmul(w3, w4); v[7] = cmul(v[7], w7);
            v[9] = cmul(v[9], cmul(w1, w8)); v[10] = cmul(v[10], cmul(w2, w8)); v[11] = cmul(v[11], cmul(w3, w8));
            v[12] = cmul(v[12], cmul(w4, w8)); v[13] = cmul(v[13], cmul(w5, w8)); v[14] = cmul(v[14], cmul(w6, w8));
            v[15] = cmul(v[15], cmul(w7, w8));
        }
        if (R3 > 1) {
            __syncthreads();                            // every butterfly has read its inputs
#pragma unroll
            for (int k = 0; k < 16; k++) sb[slot(q + 256 * p + 16 * k)] = v[k];
            __syncthreads();
            // ---- pass 3: len R3, stride 256: butterfly bb = b + N1 t reads bb + 256 j
#pragma unroll
            for (int t = 0; t < 16 / R3; t++) {
                const int bb = b + N1 * t;
#pragma unroll
                for (int j = 0; j < R3; j++) v[t * R3 + j] = sb[slot(bb + 256 * j)];
                if (R3 == 2) {
                    const cd a0 = v[t * 2], a1 = v[t * 2 + 1];
                    v[t * 2] = cadd(a0, a1); v[t * 2 + 1] = csub(a0, a1);
                } else {
                    bfly4(v[t * 4], v[t * 4 + 1], v[t * 4 + 2], v[t * 4 + 3], -1.0, v[t * 4], v[t * 4 + 1], v[t * 4 + 2], v[t * 4 + 3]);
                }
            }
        }
    }
    if (!live) return;
    // ---- results: v[t R3 + k] = X[b + N1 t + 256 k]  (R3 = 1: v[k] = X[b + 16 k])
    const long fr = frame0 + (long)g * PFI + f;
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int ch = R3 == 1 ? b + 16 * e : b + N1 * (e / R3) + 256 * (e % R3);
        if (layout == 0) out[(size_t)ch * out_stride + fr] = v[e];
        else out[(size_t)fr * out_stride + ch] = v[e];
    }
}

// ------------------------------------------------------------------------------------------------------------
// (1b) D = K/2, K = 1024: branch FIRs and transforms in ONE kernel, u never leaves the chip.
// A cluster of 8 CTAs x 128 branches covers the 1024 branches of a frame range; every CTA runs the register-resident branch
// FIR of pfb_fir_kernel on its 128 branches and sends each finished value u_m[r] straight into the shared memory of the CTA
// that will transform frame m (distributed shared memory): frames go in batches of 32, four consecutive frames per CTA, in the
// interleaved layout pfb_fft_kernel's passes use.  One cluster barrier per batch says "all 1024 values of these 32 frames have
// landed"; each CTA then runs its four 1024-point transforms (the 16 x 16 x 4 passes of pfb_fft_kernel, first pass in place)
// and stores them.  Two receive buffers alternate; the one value per branch that a batch's closing step produces for the
// batch after the next lands after the barrier (its buffer is only then known to be free).
// HBM traffic per input sample: 16 B in + 32 B out (+ warm-up), against 112 B with u making its round trip.
// MEASURED (B200, 16 Mi samples per call): 0.82 ms against 0.41 ms for the two-kernel path, so it is OFF by default
// (QC_PFB_OPT_FUSED).  Where the time goes (QUISK_PFB_DBG switches parts off): branch FIRs alone 0.34 ms -- twice the
// stand-alone FIR kernel, because 166 KB of shared memory leave one CTA = 8 warps per SM (the stand-alone kernel hides HBM
// latency with 16) and only 15 clusters of 8 are resident (GPC boundaries): 120 of 148 SMs; the transforms add 0.30 ms
// (same threads, nothing overlaps them), the remote stores 0.07 ms, the barriers 0.04 ms.  Beating two kernels that each
// run at ~70 % of HBM bandwidth needs FIR and transform warps running side by side at full occupancy, not this form.
// ------------------------------------------------------------------------------------------------------------
struct PfbFusedParams {
    const cd *in; int count;
    const cd *hist; int H;
    long long n0, f0; int nf, fs;
    int D;
    const double *taps; const cd *tw;
    cd *out; long out_stride; int layout;
    int dbg;                        // debug: 1 no sends, 2 no transforms, 4 no cluster barriers (results wrong; timing only)
};
static constexpr int PFU_BR = 128, PFU_CL = 8, PFU_K = 1024, PFU_PFI = 4;

// four transforms of batch B, thread (b, f) = (tid / 4, tid % 4): pfb_fft_kernel<4, 4> with pass 1 reading the receive buffer
__device__ __noinline__ void pfu_fft_batch(cd *sb, const cd *twl, const PfbFusedParams &p, int frame_first, int rb)
{
    constexpr int K = PFU_K, N1 = K / 16, PFI = PFU_PFI, R3 = 4;
    const int tid = threadIdx.x, f = tid % PFI, b = tid / PFI;
    const int fr = frame_first + f;
    const bool live = fr < rb;
    auto slot = [f](int i) { return PFI * (i ^ ((i >> 4) & (8 / PFI - 1))) + f; };
    cd v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = sb[slot(b + j * N1)];
    dft16(v, -1.0);
    __syncthreads();                                    // every thread holds its inputs: the buffer may be overwritten
    if (b != 0) {
        const cd w1 = fft_tw(twl, b, -1), w2 = fft_tw(twl, 2 * b, -1), w4 = fft_tw(twl, 4 * b, -1), w8 = fft_tw(twl, 8 * b, -1);
        v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[4] = cmul(v[4], w4); v[8] = cmul(v[8], w8);
        const cd w3 = cmul(w1, w2); v[3] = cmul(v[3], w3);
        const cd w5 = cmul(w1, w4); v[5] = cmul(v[5], w5);
        const cd w6 = cmul(w2, w4); v[6] = cmul(v[6], w6);
        const cd w7 = cmul(w3, w4); v[7] = cmul(v[7], w7);
        v[9] = cmul(v[9], cmul(w1, w8)); v[10] = cmul(v[10], cmul(w2, w8)); v[11] = cmul(v[11], cmul(w3, w8));
        v[12] = cmul(v[12], cmul(w4, w8)); v[13] = cmul(v[13], cmul(w5, w8)); v[14] = cmul(v[14], cmul(w6, w8));
        v[15] = cmul(v[15], cmul(w7, w8));
    }
#pragma unroll
    for (int k = 0; k < 16; k++) sb[slot(16 * b + k)] = v[k];
    __syncthreads();
    {
        const int pp = b >> 4, q = b & 15;
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = sb[slot(q + 16 * (pp + R3 * j))];
        dft16(v, -1.0);
        if (pp != 0) {
            const int pt = pp * 16;
            const cd w1 = fft_tw(twl, pt, -1), w2 = fft_tw(twl, 2 * pt, -1), w4 = fft_tw(twl, 4 * pt, -1), w8 = fft_tw(twl, 8 * pt, -1);
            v[1] = cmul(v[1], w1); v[2] = cmul(v[2], w2); v[4] = cmul(v[4], w4); v[8] = cmul(v[8], w8);
            const cd w3 = cmul(w1, w2); v[3] = cmul(v[3], w3);
            const cd w5 = cmul(w1, w4); v[5] = cmul(v[5], w5);
            const cd w6 = cmul(w2, w4); v[6] = cmul(v[6], w6);
            const cd w7 = cmul(w3, w4); v[7] = cmul(v[7], w7);
            v[9] = cmul(v[9], cmul(w1, w8)); v[10] = cmul(v[10], cmul(w2, w8)); v[11] = cmul(v[11], cmul(w3, w8));
            v[12] = cmul(v[12], cmul(w4, w8)); v[13] = cmul(v[13], cmul(w5, w8)); v[14] = cmul(v[14], cmul(w6, w8));
            v[15] = cmul(v[15], cmul(w7, w8));
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) sb[slot(q + 256 * pp + 16 * k)] = v[k];
        __syncthreads();
#pragma unroll
        for (int t = 0; t < 16 / R3; t++) {
            const int bb = b + N1 * t;
#pragma unroll
            for (int j = 0; j < R3; j++) v[t * R3 + j] = sb[slot(bb + 256 * j)];
            bfly4(v[t * 4], v[t * 4 + 1], v[t * 4 + 2], v[t * 4 + 3], -1.0, v[t * 4], v[t * 4 + 1], v[t * 4 + 2], v[t * 4 + 3]);
        }
    }
    __syncthreads();                                    // the buffer is free for the batch after the next
    if (!live) return;
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int ch = b + N1 * (e / R3) + 256 * (e % R3);
        if (p.layout == 0) p.out[(size_t)ch * p.out_stride + fr] = v[e];
        else p.out[(size_t)fr * p.out_stride + ch] = v[e];
    }
}

template <int P, int RING>
__global__ void __cluster_dims__(PFU_CL, 1, 1) __launch_bounds__(256, 1) pfb_fused_kernel(PfbFusedParams p)
{
    constexpr int K = PFU_K, PH = P / 2, PFI = PFU_PFI;
    extern __shared__ double smem_raw[];
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *rbuf = twl + fft_tw_entries(K);                             // [2][K * PFI] receive buffers
    cd (*ring)[PFU_BR] = reinterpret_cast<cd (*)[PFU_BR]>(rbuf + 2 * K * PFI);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, half = lane >> 4;
    const int rl = (tid >> 5) * 16 + (lane & 15);
    const int r = rank * PFU_BR + rl;
    const int D = p.D;
    const int ra = (int)(blockIdx.x / PFU_CL) * p.fs;               // this cluster's frames [ra, rb) of the launch
    const int rb = min(ra + p.fs, p.nf);
    if (ra >= rb) return;                                           // the whole cluster
    fft_stage_twiddles(twl, p.tw, K);
    const int tau = (2 * D - 1 - r) % D;
    double lo[PH], hi[PH];
#pragma unroll
    for (int a = 0; a < PH; a++) {
        lo[a] = p.taps[tau + K * (half * PH + a)];
        hi[a] = p.taps[tau + D + K * (half * PH + a)];
    }
    const long long fa = p.f0 + ra, fb = p.f0 + rb;
    const long long a_lo = (fa >> 1) - 1, a_hi = (fb - 1) >> 1;
    const int steps = (int)(a_hi - a_lo + 1);
    const int mrel0 = (int)(2 * a_lo - p.f0) + (r >= D ? 1 : 0);
    const int rel = (int)(a_lo * K + r - p.n0);
    const cd *in = p.in, *he = p.hist + p.H;
    cd w[PH];
#pragma unroll
    for (int q = 1; q < PH; q++) w[PH - q] = pfb_load(in, he, rel - (q + half * PH) * K, p.count, p.H);
    w[0] = pfb_load(in, he, rel - (PH + half * PH) * K, p.count, p.H);
    auto issue = [&](int step) {
        if (!half) {
            const int e = rel + step * K;
            const cd *src = in; unsigned bytes = 0;
            if (step < steps) {
                if (e >= 0) { if (e < p.count) { src = in + e; bytes = 16; } }
                else if (e >= -p.H) { src = he + e; bytes = 16; }
            }
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&ring[step & (RING - 1)][rl]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll 1
    for (int j = 0; j < RING - 2; j++) issue(j);
    const int slot_r = PFI * (r ^ ((r >> 4) & (8 / PFI - 1)));
    // value of launch-relative frame m (inside [ra, rb)) to the CTA that transforms it
    auto send = [&](cd val, int mm) {
        const int fbat = mm & 31;
        cd *dst = cluster.map_shared_rank(rbuf, fbat >> 2) + (size_t)((mm >> 5) & 1) * K * PFI + slot_r + (fbat & 3);
        *dst = val;
    };
    const int trig0 = 16 + (int)(fa & 1);                           // the step that completes batch 0; batch B: trig0 + 16 B
    const int nbatch = (rb - ra + 31) >> 5;
    int next_b = 0;
    cluster.sync();                                                 // every CTA of the cluster is running: its shared memory may be written
    for (int i = 0; i < steps; i += PH) {
#pragma unroll
        for (int j = 0; j < PH; j++) {
            if (i + j < steps) {
                const cd old = w[j];
                const double ox = __shfl_sync(0xffffffffu, old.x, lane & 15), oy = __shfl_sync(0xffffffffu, old.y, lane & 15);
                issue(i + j + RING - 2);
                asm volatile("cp.async.wait_group %0;" ::"n"(RING - 2) : "memory");
                w[j] = half ? make_double2(ox, oy) : ring[(i + j) & (RING - 1)][rl];
                double xr = 0.0, xi = 0.0, yr = 0.0, yi = 0.0;
#pragma unroll
                for (int t = 0; t < PH; t++) {
                    const cd x = w[(j - t + PH) % PH];
                    xr = fma(x.x, lo[t], xr); xi = fma(x.y, lo[t], xi);
                    yr = fma(x.x, hi[t], yr); yi = fma(x.y, hi[t], yi);
                }
                const double sr = half ? xr : yr, si = half ? xi : yi;
                const double pr = __shfl_xor_sync(0xffffffffu, sr, 16), pi = __shfl_xor_sync(0xffffffffu, si, 16);
                const int m = mrel0 + 2 * (i + j) + half;
                const cd val = half ? make_double2(yr + pr, yi + pi) : make_double2(xr + pr, xi + pi);
                const int mm = m - ra;
                const bool inside = m >= ra && m < rb;
                const bool trig = (i + j) == trig0 + 16 * next_b;           // uniform over the cluster
                const bool later = inside && trig && (mm >> 5) > next_b;    // lands in the buffer of the batch transformed two rounds ago
                if (inside && !later && !(p.dbg & 1)) send(val, mm);
                if (trig) {
                    if (!(p.dbg & 4)) cluster.sync();
                    if (later && !(p.dbg & 1)) send(val, mm);
                    if (!(p.dbg & 2)) pfu_fft_batch(rbuf + (size_t)(next_b & 1) * K * PFI, twl, p, ra + 32 * next_b + 4 * rank, rb);
                    next_b++;
                }
            }
        }
    }
    while (next_b < nbatch) {
        cluster.sync();
        pfu_fft_batch(rbuf + (size_t)(next_b & 1) * K * PFI, twl, p, ra + 32 * next_b + 4 * rank, rb);
        next_b++;
    }
    cluster.sync();                                                 // nobody leaves while a neighbour may still write into it
}

struct Channelizer {
    int K = 0, D = 0, T = 0, P = 0, smax = 0;
    double *d_taps = nullptr;
    cd *d_hist[2] = {nullptr, nullptr};
    int cur = 0;
    long long n_abs = 0;
    const cd *tw = nullptr;
    int n_sm = 148;
    int slice_frames = 1 << 16;     // frames of u per kernel pair; smaller slices were measured slower (launch gaps outweigh L2 reuse)
    int force_generic = 0;
    int fft_prefetch = 1;           // transform CTAs prefetch the inputs of the CTA one wave later into L2
    int ring = 16;                  // cp.async ring depth of the branch kernel in steps (16 or 32)
    int fft_frames = 2;             // frames interleaved per transform CTA (2: four CTAs per SM, measured 5 % faster; 4: 64-byte output runs)
    cd *d_u = nullptr; int u_frames = 0, u_bufs = 0;        // one slice of u (two back to back when pipelining)
    int pipeline = 0;               // 1: branch FIRs of slice i+1 overlap the transforms of slice i on two internal streams
    int max_clusters = 0;           // resident clusters of the fused kernel (queried once)
    int fused = 0;                  // 1: K = 1024, D = 512, 8 or 16 taps per branch run the one-kernel cluster path (pfb_fused_kernel);
                                    // measured SLOWER than the two-kernel path (0.82 vs 0.41 ms per 16 Mi samples), see the kernel's header
    cudaStream_t sa = nullptr, sb = nullptr;
    cudaEvent_t ev_fir[2] = {nullptr, nullptr}, ev_fft[2] = {nullptr, nullptr}, ev_edge = nullptr;

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
        if (const char *e = getenv("QUISK_PFB_FUSED")) fused = atoi(e) ? 1 : 0;     // debug override of QC_PFB_OPT_FUSED
        return QC_OK;
    }
    void release()
    {
        if (d_taps) cudaFree(d_taps);
        if (d_u) cudaFree(d_u);
        d_u = nullptr; u_frames = 0;
        for (int i = 0; i < 2; i++) if (d_hist[i]) cudaFree(d_hist[i]);
        d_taps = nullptr; d_hist[0] = d_hist[1] = nullptr;
        if (sa) { cudaStreamSynchronize(sa); cudaStreamDestroy(sa); sa = nullptr; }
        if (sb) { cudaStreamSynchronize(sb); cudaStreamDestroy(sb); sb = nullptr; }
        for (int i = 0; i < 2; i++) {
            if (ev_fir[i]) cudaEventDestroy(ev_fir[i]);
            if (ev_fft[i]) cudaEventDestroy(ev_fft[i]);
            ev_fir[i] = ev_fft[i] = nullptr;
        }
        if (ev_edge) { cudaEventDestroy(ev_edge); ev_edge = nullptr; }
    }
    int seek(long long n, cudaStream_t s, bool on_stream)
    {
        if (n < 0) { set_error("pfb_seek: negative sample index"); return QC_EINVAL; }
        if (on_stream) {
            // ordered on the caller's stream like the process / prime calls around it (the internal pipeline streams
            // always rejoin the caller's stream at the end of a process call)
            for (int i = 0; i < 2; i++) QC_CUDA(cudaMemsetAsync(d_hist[i], 0, (size_t)T * sizeof(cd), s));
            n_abs = n; cur = 0;
            return QC_OK;
        }
        // no stream argument: order it against everything queued on any stream, non-blocking ones included
        QC_CUDA(cudaDeviceSynchronize());
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
    template <int PP, int OVS> int launch_fir(const PfbFirParams &q, dim3 grid, cudaStream_t s)
    {
        if (ring == 32) pfb_fir_kernel<PP, OVS, 32><<<grid, 128, 0, s>>>(q);
        else pfb_fir_kernel<PP, OVS, 16><<<grid, 128, 0, s>>>(q);
        return QC_OK;
    }
    // K = 1024, D = 512: one kernel, clusters of 8 CTAs per frame range
    template <int PP> int launch_fused(const PfbFusedParams &q, int clusters, size_t sh, cudaStream_t s)
    {
        QC_CUDA(cudaFuncSetAttribute(pfb_fused_kernel<PP, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        pfb_fused_kernel<PP, 16><<<clusters * PFU_CL, 256, sh, s>>>(q);
        return QC_OK;
    }
    int process_fused(const cd *d_in, int count, cd *d_out, long out_stride, int layout, int nf, cudaStream_t s)
    {
        PfbFusedParams q;
        q.in = d_in; q.count = count; q.hist = d_hist[cur]; q.H = T; q.n0 = n_abs; q.f0 = n_abs / D; q.nf = nf;
        q.D = D; q.taps = d_taps; q.tw = tw; q.out = d_out; q.out_stride = out_stride; q.layout = layout;
        q.dbg = getenv("QUISK_PFB_DBG") ? atoi(getenv("QUISK_PFB_DBG")) : 0;
        const size_t sh0 = ((size_t)fft_tw_entries(K) + 2 * (size_t)K * PFU_PFI + (size_t)16 * PFU_BR) * sizeof(cd);
        if (max_clusters == 0) {
            // how many clusters the device holds at once (GPC boundaries decide, not the SM count)
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(PFU_CL * 64); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = sh0;
            cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = PFU_CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
            cfg.attrs = &at; cfg.numAttrs = 1;
            int nc = 0;
            if (P == 16) { cudaFuncSetAttribute(pfb_fused_kernel<16, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh0); cudaOccupancyMaxActiveClusters(&nc, pfb_fused_kernel<16, 16>, &cfg); }
            else { cudaFuncSetAttribute(pfb_fused_kernel<8, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh0); cudaOccupancyMaxActiveClusters(&nc, pfb_fused_kernel<8, 16>, &cfg); }
            cudaGetLastError();
            max_clusters = nc > 0 ? nc : n_sm / PFU_CL;
            if (getenv("QUISK_PFB_DBG")) fprintf(stderr, "pfb fused: %d clusters of %d CTAs resident\n", max_clusters, PFU_CL);
        }
        // frame ranges: a whole number of waves of resident clusters, at least 256 frames each
        int waves = getenv("QUISK_PFB_WAVES") ? atoi(getenv("QUISK_PFB_WAVES")) : 2;
        int S = waves * max_clusters;
        if (S > (nf + 255) / 256) S = (nf + 255) / 256;
        if (S < 1) S = 1;
        int fs = (nf + S - 1) / S;
        fs += fs & 1;
        q.fs = fs;
        S = (nf + fs - 1) / fs;
        const size_t sh = ((size_t)fft_tw_entries(K) + 2 * (size_t)K * PFU_PFI + (size_t)16 * PFU_BR) * sizeof(cd);
        int rc = P == 16 ? launch_fused<16>(q, S, sh, s) : launch_fused<8>(q, S, sh, s);
        if (rc != QC_OK) return rc;
        count_launch();
        QC_CUDA_LAUNCH();
        return QC_OK;
    }
    // D = K or K/2: slices of (register FIR kernel, transform kernel)
    int process_fast(const cd *d_in, int count, cd *d_out, long out_stride, int layout, int nf, cudaStream_t s)
    {
        const int ovs = K / D;
        int sf = slice_frames < PF ? PF : (slice_frames / PF) * PF;
        if (sf > nf) sf = ((nf + PF - 1) / PF) * PF;
        const int want_bufs = pipeline ? 2 : 1;
        if (sf > u_frames || want_bufs > u_bufs) {
            if (d_u) cudaFree(d_u);
            d_u = nullptr; u_frames = 0; u_bufs = 0;
            QC_CUDA(cudaMalloc((void **)&d_u, (size_t)want_bufs * sf * K * sizeof(cd)));
            u_frames = sf; u_bufs = want_bufs;
        }
        const bool pipe = pipeline && nf > sf;
        cudaStream_t s_fir = s, s_fft = s;
        if (pipe) {
            if (!sa) {
                QC_CUDA(cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking));
                QC_CUDA(cudaStreamCreateWithFlags(&sb, cudaStreamNonBlocking));
                for (int i = 0; i < 2; i++) {
                    QC_CUDA(cudaEventCreateWithFlags(&ev_fir[i], cudaEventDisableTiming));
                    QC_CUDA(cudaEventCreateWithFlags(&ev_fft[i], cudaEventDisableTiming));
                }
                QC_CUDA(cudaEventCreateWithFlags(&ev_edge, cudaEventDisableTiming));
            }
            QC_CUDA(cudaEventRecord(ev_edge, s));
            QC_CUDA(cudaStreamWaitEvent(sa, ev_edge, 0));
            QC_CUDA(cudaStreamWaitEvent(sb, ev_edge, 0));
            s_fir = sa; s_fft = sb;
        }
        const int pfi = fft_frames == 2 ? 2 : 4;
        const size_t sh = ((size_t)fft_tw_entries(K) + (size_t)pfi * K) * sizeof(cd);
        if (sh > 48 * 1024) {
            if (K == 1024 && pfi == 4) QC_CUDA(cudaFuncSetAttribute(pfb_fft_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
            if (K == 512 && pfi == 4) QC_CUDA(cudaFuncSetAttribute(pfb_fft_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        }
        const long long m0 = n_abs / D;
        for (int f0 = 0, si = 0; f0 < nf; f0 += sf, si++) {
            const int nfs = nf - f0 < sf ? nf - f0 : sf;
            cd *ub = d_u + (size_t)(pipe ? (si & 1) : 0) * u_frames * K;
            if (pipe && si >= 2) QC_CUDA(cudaStreamWaitEvent(sa, ev_fft[si & 1], 0));       // the slice buffer is free again
            PfbFirParams q;
            q.in = d_in; q.count = count; q.hist = d_hist[cur]; q.H = T; q.n0 = n_abs; q.f0 = m0 + f0; q.nf = nfs;
            q.K = K; q.D = D; q.taps = d_taps; q.u = ub;
            // frame ranges: enough CTAs to fill the machine three deep, at least 2 P frames each
            const int bx = K / PFB_BR;
            int S = (n_sm * 4 + bx - 1) / bx;
            const int min_fs = 2 * P * ovs;
            if (S > (nfs + min_fs - 1) / min_fs) S = (nfs + min_fs - 1) / min_fs;
            if (S < 1) S = 1;
            int fs = (nfs + S - 1) / S;
            fs += fs & 1;
            q.fs = fs;
            S = (nfs + fs - 1) / fs;
            int rc = QC_EINVAL;
#define PFB_FIR(PP) case PP: rc = ovs == 2 ? launch_fir<PP, 2>(q, dim3(bx, S), s_fir) : launch_fir<PP, 1>(q, dim3(bx, S), s_fir); break;
            switch (P) { PFB_FIR(4) PFB_FIR(8) PFB_FIR(16) PFB_FIR(32) }
#undef PFB_FIR
            if (rc != QC_OK) return rc;
            count_launch();
            QC_CUDA_LAUNCH();
            if (pipe) { QC_CUDA(cudaEventRecord(ev_fir[si & 1], sa)); QC_CUDA(cudaStreamWaitEvent(sb, ev_fir[si & 1], 0)); }
            const int gf = (nfs + pfi - 1) / pfi, nth = pfi * K / 16;
            const int ahead = fft_prefetch ? n_sm * (8 / pfi) : 0;           // one full wave of resident CTAs ahead
#define PFB_FFT(R3) do { if (pfi == 4) pfb_fft_kernel<R3, 4><<<gf, nth, sh, s_fft>>>(ub, nfs, tw, d_out, out_stride, f0, layout, ahead); \
                         else pfb_fft_kernel<R3, 2><<<gf, nth, sh, s_fft>>>(ub, nfs, tw, d_out, out_stride, f0, layout, ahead); } while (0)
            if (K == 1024) PFB_FFT(4); else if (K == 512) PFB_FFT(2); else PFB_FFT(1);
#undef PFB_FFT
            count_launch();
            QC_CUDA_LAUNCH();
            if (pipe) QC_CUDA(cudaEventRecord(ev_fft[si & 1], sb));
        }
        if (pipe) {
            QC_CUDA(cudaEventRecord(ev_edge, sa)); QC_CUDA(cudaStreamWaitEvent(s, ev_edge, 0));
            QC_CUDA(cudaEventRecord(ev_edge, sb)); QC_CUDA(cudaStreamWaitEvent(s, ev_edge, 0));
        }
        return QC_OK;
    }
    int process(const cd *d_in, int count, cd *d_out, long out_stride, int layout, int *n_frames, cudaStream_t s)
    {
        if (count < 0 || count > (1 << 30)) { set_error("pfb_process: count must be in [0, 2^30]"); return QC_EINVAL; }
        const int nf = frames(count);
        if (n_frames) *n_frames = nf;
        if (count == 0) return QC_OK;
        if (nf > 0 && (layout == 0 ? out_stride < nf : out_stride < K)) { set_error("pfb_process: out_stride %ld too small", out_stride); return QC_EINVAL; }
        if (nf > 0 && !force_generic && fused && K == PFU_K && 2 * D == K && (P == 16 || P == 8)) {
            int rc = process_fused(d_in, count, d_out, out_stride, layout, nf, s);
            if (rc != QC_OK) return rc;
        } else if (nf > 0 && !force_generic && (D == K || 2 * D == K)) {
            int rc = process_fast(d_in, count, d_out, out_stride, layout, nf, s);
            if (rc != QC_OK) return rc;
        } else if (nf > 0) {
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
int quisk_cuda_pfb_seek(qcChannelizer *p, long long n_abs) { return p ? p->c.seek(n_abs, nullptr, false) : QC_EINVAL; }
int quisk_cuda_pfb_seek_async(qcChannelizer *p, long long n_abs, void *stream) { return p ? p->c.seek(n_abs, (cudaStream_t)stream, true) : QC_EINVAL; }
int quisk_cuda_pfb_prime(qcChannelizer *p, const void *d_in, int count, void *stream)
{
    if (!p || count < 0) return QC_EINVAL;
    if (count == 0) return QC_OK;
    return p->c.roll((const cd *)d_in, count, (cudaStream_t)stream);
}
int quisk_cuda_pfb_set_option(qcChannelizer *p, int option, int value)
{
    if (!p) return QC_EINVAL;
    switch (option) {
    case QC_PFB_OPT_SLICE_FRAMES: if (value < 4) return QC_EINVAL; p->c.slice_frames = value; return QC_OK;
    case QC_PFB_OPT_GENERIC: p->c.force_generic = value ? 1 : 0; return QC_OK;
    case QC_PFB_OPT_PIPELINE: p->c.pipeline = value ? 1 : 0; return QC_OK;
    case QC_PFB_OPT_FFT_PREFETCH: p->c.fft_prefetch = value ? 1 : 0; return QC_OK;
    case QC_PFB_OPT_RING: if (value != 16 && value != 32) return QC_EINVAL; p->c.ring = value; return QC_OK;
    case QC_PFB_OPT_FFT_FRAMES: if (value != 2 && value != 4) return QC_EINVAL; p->c.fft_frames = value; return QC_OK;
    case QC_PFB_OPT_FUSED: p->c.fused = value ? 1 : 0; return QC_OK;
    }
    qc::set_error("pfb_set_option: unknown option %d", option);
    return QC_EINVAL;
}
int quisk_cuda_pfb_process(qcChannelizer *p, const void *d_in, int count, void *d_out, long out_stride, int layout,
                           int *n_frames, void *stream)
{
    if (!p) { qc::set_error("pfb_process: null handle"); return QC_EINVAL; }
    return p->c.process((const cd *)d_in, count, (cd *)d_out, out_stride, layout, n_frames, (cudaStream_t)stream);
}

}  // extern "C"
