// generated: the AGC block function of wdsp_rxa_fused.cu in isolation
#include <cstdio>
#include <cuda_runtime.h>
struct AgcParams {
    int mode, pmode, ring_buffsize, attack_buffsize, hang_enable;
    double sample_rate, fixed_gain, attack_mult, decay_mult, fast_decay_mult, fast_backmult, onemfast_backmult,
           out_target, min_volts, inv_out_target, slope_constant, inv_max_input, hang_level, hang_backmult,
           onemhang_backmult, hang_decay_mult, pop_ratio, hangtime;
    double tau_attack, tau_decay, max_gain, var_gain, max_input, out_targ, tau_fast_backaverage, tau_fast_decay,
           tau_hang_backmult, hang_thresh, tau_hang_decay;
    int n_tau;
};
// One sample through the reference's general machine (wcpAGC.c:195-333), same tests in the same order; expects ring_max,
// advances i.  The two back averages (wcpAGC.c:192-193) only look at the input: the LIN warp computes them for the whole
// block (lanes 6 and 7) and this path -- the only one that reads them -- waits until that warp has passed sample i.
#define RF_AGC_GENERAL_STEP                                                                                             \
    while (*linpos <= i) { }                                                                                            \
    const double fb = FBA[i], hb = HBA[i];                                                                              \
    if (hang_counter > 0) --hang_counter;                                                                               \
    {                                                                                                                   \
        const double d = __dsub_rn(ring_max, volts);                                                                    \
        if (ring_max >= volts) {                                                                                        \
            if (state_ >= 2) save_volts = volts;                                                                        \
            state_ = 0;                                                                                                 \
            volts = __dadd_rn(volts, __dmul_rn(d, k_attack));                                                           \
        } else if (state_ >= 3) {                                                                                       \
            volts = __dadd_rn(volts, __dmul_rn(d, state_ == 3 ? k_decay : k_hdecay));                                   \
        } else if (state_ == 0) {                                                                                       \
            if (volts > __dmul_rn(k_pop, fb)) { state_ = 1; volts = __dadd_rn(volts, __dmul_rn(d, k_fdecay)); }         \
            else if (a.hang_enable && hb > k_hlevel) { state_ = 2; hang_counter = (int)(a.hangtime * a.sample_rate); decay_type = 1; } \
            else { state_ = 3; volts = __dadd_rn(volts, __dmul_rn(d, k_decay)); decay_type = 0; }                       \
        } else if (state_ == 1) {                                                                                       \
            if (volts > save_volts) volts = __dadd_rn(volts, __dmul_rn(d, k_fdecay));                                   \
            else if (hang_counter > 0) state_ = 2;                                                                      \
            else if (decay_type == 0) { state_ = 3; volts = __dadd_rn(volts, __dmul_rn(d, k_decay)); }                  \
            else { state_ = 4; volts = __dadd_rn(volts, __dmul_rn(d, k_hdecay)); }                                      \
        } else {                                                                                                        \
            if (hang_counter == 0) { state_ = 4; volts = __dadd_rn(volts, __dmul_rn(d, k_hdecay)); }                    \
        }                                                                                                               \
    }                                                                                                                   \
    if (volts < k_minv) volts = k_minv;                                                                                 \
    RV[i] = volts;                                                                                                      \
    i++;

// ---- role 1: the AGC warp.  All 32 lanes run the same instructions on the same data (no divergence inside the warp, so the
// CTA-wide barriers are reached by whole warps) and store the same values to the same shared-memory words; lane 0 alone
// writes the state back to global memory.  (A store predicated on the lane inside the dependent chain made the compiler
// re-derive the chain from the chunk start for every store: 74 cycles per sample instead of 24.)
// One block of the volts machine: A[i] = |sample leaving the delay line|, RV[i] = ring_max on the way in, volts on the way out.
__device__ __forceinline__ void rf_agc_block(double *RV, const double *FBA, const double *HBA, volatile int *linpos, int n, const AgcParams &a,
                                             double &volts, double &save_volts, int &hang_counter, int &decay_type, int &state_, long long *dbg)
{
    int n_single = 0;
    const long long t_in = clock64();
    const double k_attack = a.attack_mult, k_decay = a.decay_mult, k_hdecay = a.hang_decay_mult, k_fdecay = a.fast_decay_mult,
                 k_pop = a.pop_ratio, k_hlevel = a.hang_level, k_minv = a.min_volts;
    int i = 0;
    // shared-memory accesses of the run loop go through 32-bit shared-window addresses kept in registers: left to itself the
    // compiler re-derives the window base (an S2R of the cluster CTA id) at the top of every chunk, ~100 cycles in front of
    // the chain each time
    const unsigned rv_s = (unsigned)__cvta_generic_to_shared(RV);
    constexpr int CH = 16;
    while (i < n) {
        if (n - i >= CH && (state_ == 0 || state_ >= 3) && !(volts < k_minv)) {      // (a fresh channel starts with volts = 0: the general path clamps it first)
            // runs of sixteen samples on the assumption that the state does not change
            const double M = state_ == 0 ? k_attack : (state_ == 3 ? k_decay : k_hdecay);
            const bool want = state_ == 0;
            double rm[CH], rn[CH];
            unsigned a0 = rv_s + 8u * (unsigned)i;
#pragma unroll
            for (int j = 0; j < CH; j++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rm[j]) : "r"(a0 + 8u * j));
            bool ok = true;
            while (ok && n - i >= CH) {
                // the following run's ring_max is asked for now (RV beyond this run is still input): its latency hides under the chain
                const bool more = n - i >= 2 * CH;
                if (more) {
#pragma unroll
                    for (int j = 0; j < CH; j++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rn[j]) : "r"(a0 + 8u * (CH + j)));
                }
                double v = volts;
                // The run holds if every d = ring_max - volts has the sign the state expects (d >= 0: attack, d < 0: decay; d is the
                // chain's own first operation), read from the SIGN BITS with integer logic, off the chain.  The min_volts clamp needs
                // no per-sample test: volts moves monotonically within a run (up in attack, down in decay), it entered the run at
                // or above min_volts, so only a decay run's last value can fall below.
                int s_and = -1, s_or = 0;
                double vv[CH];
#pragma unroll
                for (int j = 0; j < CH; j++) {
                    const double d = __dsub_rn(rm[j], v);
                    v = __dadd_rn(v, __dmul_rn(d, M));
                    const int hd = __double2hiint(d);
                    s_and &= hd; s_or |= hd;
                    vv[j] = v;
                }
                ok = want ? s_or >= 0 : (s_and < 0 && !(v < k_minv));
                if (ok) {
                    // commit (all lanes store the same values)
#pragma unroll
                    for (int j = 0; j < CH; j++) asm volatile("st.shared.f64 [%0], %1;" :: "r"(a0 + 8u * j), "d"(vv[j]) : "memory");
                    volts = v;
                    hang_counter = hang_counter > CH ? hang_counter - CH : 0;
                    i += CH;
                    a0 += 8u * CH;
#pragma unroll
                    for (int j = 0; j < CH; j++) rm[j] = rn[j];
                }
            }
            if (ok) continue;                               // fewer than a run's worth of samples left
            // the run did not hold (a handful of times per block): these samples go through the general machine one by one
            for (int g = 0; g < CH; g++) {
                const double ring_max = RV[i];
                RF_AGC_GENERAL_STEP
            }
            continue;
        }
        // one sample through the general machine (states 1 and 2, and the last few samples of a block)
        {
            const double ring_max = RV[i];
            n_single++;
            RF_AGC_GENERAL_STEP
        }
    }
    if (dbg) { dbg[24] = 0; dbg[25] = 0; dbg[26] = n_single; dbg[27] = clock64() - t_in; }
}
#undef RF_AGC_GENERAL_STEP


__global__ void k(double *out, long long *cyc, int n, AgcParams a, int variant)
{
    extern __shared__ double sm[];
    double *RV = sm, *FBA = sm + n, *HBA = sm + 2 * n;
    volatile int *linpos = (volatile int *)(sm + 3 * n);
    for (int i = threadIdx.x; i < n; i += blockDim.x) { RV[i] = 0.2; FBA[i] = 0.1; HBA[i] = 0.1; }
    if (threadIdx.x == 0) *linpos = n;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    double volts = 0.9, save_volts = 0.5; int hc = 0, dt = 0, st = 3;
    if (warp == 4) {
        long long t0 = clock64();
        rf_agc_block(RV, FBA, HBA, linpos, n, a, volts, save_volts, hc, dt, st, nullptr);
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = volts + RV[threadIdx.x % n];
}
int main()
{
    const int n = 1024;
    double *d; long long *dc, h;
    cudaMalloc(&d, 148 * 192 * 8); cudaMalloc(&dc, 16);
    AgcParams a = {};
    a.attack_mult = 5e-3; a.decay_mult = 2e-5; a.hang_decay_mult = 1e-4; a.fast_decay_mult = 1e-3; a.pop_ratio = 5; a.hang_level = 0.6; a.min_volts = 1e-6;
    k<<<64, 192, 3 * n * 8 + 16>>>(d, dc, n, a, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, dc, 8, cudaMemcpyDeviceToHost);
    printf("rf_agc_block: %.1f cycles/sample (%s)\n", h / (double)n, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
