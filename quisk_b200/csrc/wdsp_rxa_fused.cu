// quisk_b200/csrc/wdsp_rxa_fused.cu -- xrxa (wdsp/RXA.c:561-598) as ONE kernel for the side-band configurations of the
// chain: input block -> ADC meter -> nbp0 [-> bp1] (fircore, wdsp/firmin.c:409-430) -> S meter -> wcpAGC
// (wdsp/wcpAGC.c:161-348) -> AGC meter (wdsp/meter.c:75-107) -> panel (+ siphon) -> output, for any number of DSP blocks per
// launch.  It replaces six launches per DSP block whose recurrent stages each parked a CTA behind one lane.
//
// One CTA per channel, 128 worker threads + two sequential warps:
//   * the workers do everything that is parallel along time -- the overlap-save transforms and partition MACs, magnitudes,
//     the sliding maximum the AGC calls ring_max, the gain law with its log10 and divide, the panel, all loads and stores;
//   * lane 0 of the AGC warp walks the five-state volts machine.  Almost every sample is one of two things -- attack
//     (ring_max >= volts, state 0) or steady decay (states 3 / 4) -- so the lane SPECULATES eight samples at a time on the
//     state it is in: volts += (ring_max - volts) * mult is then a bare subtract / multiply / add chain (3 x 8 cycles on this
//     part) with the comparisons off the chain; the first sample whose comparison says otherwise ends the run, the samples
//     in front of it are committed, and that one sample goes through the reference's general switch.  Same operations
//     in the same order with the same roundings as the reference, whatever the path;
//   * the lanes of the LIN warp run the first-order recurrences that only look at their own input, all with one
//     instruction stream s = c1 * x[i] + c2 * s: the averages and peak decays of the three meters (the AGC meter one block
//     late, finished after the last block of the launch).
// Arithmetic in the recurrent parts is written with __dmul_rn / __dadd_rn: the reference is compiled for baseline x86-64,
// where a * b + c rounds twice.  The transforms and MACs are the code of fircore_kernel (wdsp_fircore.cu), so the fused
// kernel and the per-stage kernels produce the same bits and share their state arrays: a stream may switch between them.
#include <cstring>
#include "fft_device.cuh"
#include "wdsp_internal.h"

namespace qc {

static constexpr int RF_WORK = 128;         // worker threads (the transform's lanes)
static constexpr int RF_THREADS = 192;      // + AGC warp + LIN warp

struct RxaFusedParams {
    const cd *in; long in_stride; cd *out; long out_stride;
    const cd *in_adc; long adc_stride;      // what the ADC meter looks at when `in` is already filtered (the wide path); nullptr: `in`
    int n, nblocks, C;
    int n_fir; cd *prev[2]; cd *fdl[2]; const cd *mask[2]; int nfor[2]; int buffidx[2];
    const cd *tw;
    double *mst[3]; double *mres[3];        // meters: adc, s, agc -- state [C][2] (avg, peak), result [C][3]
    double m_ma, m_mp;
    const double *mtable;
    int agc_run; double *agc_state; double *agc_hist; AgcParams a;
    double gI, gQ;
    cd *sip; int sipsize, sip_idx;
    long long *dbg;                         // debug: clock64() stamps of the phases of the last block of channel 0
};

__device__ __forceinline__ void rf_bar_work() { asm volatile("bar.sync 1, %0;" :: "r"(RF_WORK) : "memory"); }
__device__ __forceinline__ void rf_bar_all() { asm volatile("bar.sync 0, %0;" :: "r"(RF_THREADS) : "memory"); }

__device__ __forceinline__ double rf_smag(cd v) { return __dadd_rn(__dmul_rn(v.x, v.x), __dmul_rn(v.y, v.y)); }

// Shared-memory map of one CTA (doubles after the transform buffer)
struct RfSmem {
    cd *twl, *S;
    double *SMS, *SMADC, *A, *RV, *BM, *SMAGC, *SC, *FBA, *HBA;
    volatile int *linpos;
};
__device__ __forceinline__ RfSmem rf_map(double *raw, int n, int tot)
{
    RfSmem m;
    const int n2 = 2 * n;
    m.twl = reinterpret_cast<cd *>(raw);
    m.S = m.twl + fft_tw_entries(n2);
    m.SMS = reinterpret_cast<double *>(m.S + n);            // |y|^2 of the block (upper half of S, free after the inverse transform)
    m.SMADC = m.SMS + n;                                    // |x|^2 of the block
    m.A = reinterpret_cast<double *>(m.S + n2);             // [tot] magnitudes of [history | block]
    m.RV = m.A + tot;                                       // [n] ring_max -> volts
    m.BM = m.RV + n;                                        // [(tot + 31) / 32 + 1] maxima of 32-sample blocks of A
    m.SMAGC = m.BM + ((tot + 31) / 32 + 1);                 // [n] |AGC output|^2 of the block just finished (the LIN warp reads it one block late)
    m.SC = m.SMAGC + n;                                     // [16] 0..3 warp maxima, 4 np_adc, 5 np_s, 6 np_agc, 8.. LIN lanes' dummy store targets
    m.FBA = m.SC + 48;                                      // [n] fast_backaverage after each sample (LIN lane 6), read by the AGC lane's general path
    m.HBA = m.FBA + n;                                      // [n] hang_backaverage (LIN lane 7)
    m.linpos = reinterpret_cast<volatile int *>(m.HBA + n); // how far the LIN warp has got in the current block
    return m;
}

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
    __syncwarp();                               /* every lane has read ring_max from RV[i] */                           \
    if (l0) RV[i] = volts;                                                                                              \
    __syncwarp();                                                                                                       \
    i++;

// ---- role 1: the AGC warp.  All 32 lanes run the same instructions on the same data (no divergence inside the warp, so the
// CTA-wide barriers are reached by whole warps); lane 0 alone stores -- to shared memory through inline PTX that carries its
// own predicate, and the state back to global memory.  (A C++ store predicated on the lane inside the dependent chain made
// the compiler re-derive the chain from the chunk start for every store: 74 cycles per sample instead of 24.)  What
// compute-sanitizer's racecheck still reports in this kernel is the one lock-free hand-over it has: the LIN warp's lanes 6 / 7
// write FBA / HBA and then publish their progress in `linpos` (fence, then a volatile store), the AGC warp's general path
// spins on `linpos` before it reads them -- a flag protocol racecheck cannot see through.
// One block of the volts machine: A[i] = |sample leaving the delay line|, RV[i] = ring_max on the way in, volts on the way out.
__device__ __forceinline__ void rf_agc_block(double *RV, const double *FBA, const double *HBA, volatile int *linpos, int n, const AgcParams &a,
                                             double &volts, double &save_volts, int &hang_counter, int &decay_type, int &state_, long long *dbg)
{
    int n_single = 0;
    const long long t_in = clock64();
    const double k_attack = a.attack_mult, k_decay = a.decay_mult, k_hdecay = a.hang_decay_mult, k_fdecay = a.fast_decay_mult,
                 k_pop = a.pop_ratio, k_hlevel = a.hang_level, k_minv = a.min_volts;
    int i = 0;
    // Lane 0 alone stores (the predicate lives INSIDE the inline PTX, so the compiler still sees an unconditional statement and
    // leaves the chain alone); every lane loads.  __syncwarp() at the run boundaries orders lane 0's stores against the other
    // lanes' later loads of the same words, so the warp is race-free by construction and not only because it runs converged.
    const int l0 = (threadIdx.x & 31) == 0 ? 1 : 0;
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
            double rm[CH], rn[CH], vp[CH];
            unsigned a0 = rv_s + 8u * (unsigned)i;
#pragma unroll
            for (int j = 0; j < CH; j++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rm[j]) : "r"(a0 + 8u * j));
            // The commit of a run that held is issued INSIDE the next run's chain (a single warp issues in order: sixteen stores and
            // sixteen loads in front of the chain cost ~100 cycles per run, under it nothing).  Until the first run of this visit has
            // held, the slot stores its own input back (vp = rm at the same addresses): a no-op.
            unsigned ap = a0;
#pragma unroll
            for (int j = 0; j < CH; j++) vp[j] = rm[j];
            bool ok = true;
            while (ok && n - i >= CH) {
                __syncwarp();
                // the following run's ring_max is asked for now (RV beyond this run is still input): its latency hides under the chain
                const bool more = n - i >= 2 * CH;
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
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" :: "r"(ap + 8u * j), "d"(vp[j]), "r"(l0) : "memory");
                    if (more) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rn[j]) : "r"(a0 + 8u * (CH + j)));
                }
                ok = want ? s_or >= 0 : (s_and < 0 && !(v < k_minv));
                if (ok) {
                    // held: its values wait in vp for the next pass through the loop (or the flush below)
                    ap = a0;
#pragma unroll
                    for (int j = 0; j < CH; j++) vp[j] = vv[j];
                    volts = v;
                    hang_counter = hang_counter > CH ? hang_counter - CH : 0;
                    i += CH;
                    a0 += 8u * CH;
#pragma unroll
                    for (int j = 0; j < CH; j++) rm[j] = rn[j];
                }
            }
            // whatever held last is still in registers: commit it (after a failed run it has been stored already; storing twice is harmless)
            __syncwarp();
#pragma unroll
            for (int j = 0; j < CH; j++) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p st.shared.f64 [%0], %1;\n\t}" :: "r"(ap + 8u * j), "d"(vp[j]), "r"(l0) : "memory");
            __syncwarp();
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

__device__ __forceinline__ void rf_agc_role(const RxaFusedParams &P, const RfSmem &m, int n, bool agc_on, double *ast, bool lane0)
{
    double volts = ast[3], save_volts = ast[4];
    int hang_counter = (int)ast[7], decay_type = (int)ast[8], state_ = (int)ast[9];
    rf_bar_all();                                           // set-up done
    for (int b = 0; b < P.nblocks; b++) {
        rf_bar_all();                                       // the workers have A and ring_max ready
        long long t0 = 0;
        const bool stamp = P.dbg && blockIdx.x == 0 && b == P.nblocks - 1 && lane0;
        if (stamp) t0 = clock64();
        if (agc_on) rf_agc_block(m.RV, m.FBA, m.HBA, m.linpos, n, P.a, volts, save_volts, hang_counter, decay_type, state_, stamp ? P.dbg : nullptr);
        if (stamp) { P.dbg[3] = t0; P.dbg[4] = clock64(); }
        if (lane0 && agc_on && b == P.nblocks - 1) {
            ast[3] = volts; ast[4] = save_volts;             // ast[5], ast[6] (the back averages) belong to the LIN warp
            ast[7] = hang_counter; ast[8] = decay_type; ast[9] = state_; ast[10] = __dmul_rn(volts, P.a.inv_out_target);
        }
        rf_bar_all();                                       // volts ready for the gain law
        rf_bar_all();                                       // end of block
    }
}

// ---- role 2: the LIN warp.  Lane l < 6: adc avg, adc peak, s avg, s peak, agc avg, agc peak; s = c1 * x[i] + c2 * s
// (averages: c1 = 1 - mult, c2 = mult; peak decays: c1 = 0), one instruction stream for all lanes, inputs fetched one
// group ahead of the dependent multiply-add chain.
__device__ __forceinline__ double rf_lin_block(const double *x, int n, double c1, double c2, double s, double *dst, int dstep, volatile int *pos)
{
    // dst / dstep: where the running value goes after every sample -- an [n] array for the two back-average lanes (dstep 1),
    // a private dummy word for the others (dstep 0): one instruction stream, no lane-predicated store in the chain
    double xa[4], xb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) xa[j] = j < n ? x[j] : 0.0;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) xb[j] = i + 4 + j < n ? x[i + 4 + j] : 0.0;
        const double t0 = __dmul_rn(c1, xa[0]), t1 = __dmul_rn(c1, xa[1]), t2 = __dmul_rn(c1, xa[2]), t3 = __dmul_rn(c1, xa[3]);
        s = __dadd_rn(__dmul_rn(c2, s), t0); dst[(i + 0) * dstep] = s;
        s = __dadd_rn(__dmul_rn(c2, s), t1); dst[(i + 1) * dstep] = s;
        s = __dadd_rn(__dmul_rn(c2, s), t2); dst[(i + 2) * dstep] = s;
        s = __dadd_rn(__dmul_rn(c2, s), t3); dst[(i + 3) * dstep] = s;
        if (pos && (i & 12) == 12) { __threadfence_block(); __syncwarp(); if ((threadIdx.x & 31) == 0) *pos = i + 4; }    // progress for the AGC lane's general path, every 16 samples (lanes 6 and 7 have stored their averages)
#pragma unroll
        for (int j = 0; j < 4; j++) xa[j] = xb[j];
    }
    for (; i < n; i++) { s = __dadd_rn(__dmul_rn(c2, s), __dmul_rn(c1, x[i])); dst[i * dstep] = s; if (pos) { __threadfence_block(); __syncwarp(); if ((threadIdx.x & 31) == 0) *pos = i + 1; } }
    return s;
}

__device__ __forceinline__ void rf_lin_role(const RxaFusedParams &P, const RfSmem &m, int n, int c, int lane, bool agc_on, double *ast)
{
    // lanes 0..5: adc avg, adc peak, s avg, s peak, agc avg, agc peak; lanes 6, 7: the AGC's fast and hang back averages
    const int mt = lane < 6 ? lane >> 1 : 3, pk = lane & 1;
    const bool meter = lane < 6, back = agc_on && (lane == 6 || lane == 7), live = meter || back;
    double s = meter ? P.mst[mt][(size_t)c * 2 + pk] : (back ? ast[lane - 1] : 0.0);        // ast[5] fast_backaverage, ast[6] hang_backaverage
    double c1 = 0.0, c2 = 0.0;
    if (meter) { c1 = pk ? 0.0 : 1.0 - P.m_ma; c2 = pk ? P.m_mp : P.m_ma; }                 // (1.0 - mult_average), meter.c:90
    else if (lane == 6) { c1 = P.a.fast_backmult; c2 = P.a.onemfast_backmult; }             // wcpAGC.c:192
    else if (lane == 7) { c1 = P.a.hang_backmult; c2 = P.a.onemhang_backmult; }             // wcpAGC.c:193
    const double *src = mt == 0 ? m.SMADC : (mt == 1 ? m.SMS : (mt == 2 ? m.SMAGC : m.A));
    double *dst = lane == 6 ? m.FBA : (lane == 7 ? m.HBA : m.SC + 8 + lane);
    const int dstep = lane == 6 || lane == 7 ? 1 : 0;
    if (lane == 0) *m.linpos = 0;
    rf_bar_all();
    for (int b = 0; b < P.nblocks; b++) {
        rf_bar_all();
        // the AGC meter runs one block late (its input is the gain law's output); its lanes idle through the first block
        const bool run = mt != 2 || b > 0;
        // (an idle lane still runs the instruction stream, with c1 = 0: its input must be finite -- SMAGC is not written yet)
        const double r = rf_lin_block(run ? src : m.SMS, n, run ? c1 : 0.0, run ? c2 : 1.0, s, dst, dstep, m.linpos);
        if (live) { s = r; if (meter && pk && run) { const double np = m.SC[4 + mt]; if (np > s) s = np; } }     // meter.c:95
        if (P.dbg && blockIdx.x == 0 && b == P.nblocks - 1 && lane == 0) P.dbg[5] = clock64();
        rf_bar_all();
        if (lane == 0) *m.linpos = 0;
        rf_bar_all();
    }
    // the AGC meter's last block, then the meters' states and readings
    if (P.nblocks > 0) {
        const double r = rf_lin_block(m.SMAGC, n, mt == 2 ? c1 : 0.0, mt == 2 ? c2 : 1.0, s, m.SC + 8 + lane, 0, nullptr);
        if (mt == 2) { s = r; if (pk) { const double np = m.SC[6]; if (np > s) s = np; } }
    }
    if (meter) {
        P.mst[mt][(size_t)c * 2 + pk] = s;
        P.mres[mt][(size_t)c * 3 + pk] = 10.0 * mlog10_dev(P.mtable, s + 1.0e-40);
        if (lane == 4) P.mres[2][(size_t)c * 3 + 2] = 20.0 * mlog10_dev(P.mtable, ast[10] + 1.0e-40);      // meter.c:99: *pgain = the AGC's gain
        else if (!pk) P.mres[mt][(size_t)c * 3 + 2] = 0.0;                                                  // the other two have no gain reading
    } else if (back) {
        ast[lane - 1] = s;
    }
}

// ---- role 0: the 128 workers
__device__ __forceinline__ void rf_block_max(double lm, double *scratch4, double *dst, int tid)
{
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, lm, o); lm = t > lm ? t : lm; }
    if ((tid & 31) == 0) scratch4[tid >> 5] = lm;
    rf_bar_work();
    if (tid == 0) { double mx = scratch4[0]; for (int w = 1; w < RF_WORK / 32; w++) mx = scratch4[w] > mx ? scratch4[w] : mx; *dst = mx; }
}

// One fircore turn as a real call: its register allocation (the transform's sixteen points per thread, the four-bin MAC) is
// then separate from the worker loop's, which carries two dozen pointers of its own.
__device__ __noinline__ void rf_fircore(cd *S, const cd *twl, cd *pv, cd *fd, const cd *__restrict__ fmask, int nfor, int bi, int n, int tid,
                                        const cd *__restrict__ x, double *smag_x, long long *dbg)
{
    const int n2 = 2 * n;
    if (dbg && tid == 0) dbg[0] = clock64();
    if (x) {
        // all of this thread's loads (the block and the previous block, both behind L2) are issued before anything waits on one
        cd xv[8], pvv[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { const int i = tid + k * RF_WORK; if (i < n) { xv[k] = x[i]; pvv[k] = pv[i]; } }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = tid + k * RF_WORK;
            if (i < n) { S[fsw(i)] = pvv[k]; S[fsw(n + i)] = xv[k]; pv[i] = xv[k]; if (smag_x) smag_x[i] = rf_smag(xv[k]); }
        }
    } else {                                                // second fircore of the chain: its input is the first one's output
        cd t0[8], pvv[8];                                   // n <= 1024: at most 8 samples per worker
#pragma unroll
        for (int k = 0; k < 8; k++) { const int i = tid + k * RF_WORK; if (i < n) { t0[k] = S[fsw(i)]; pvv[k] = pv[i]; } }
        rf_bar_work();
#pragma unroll
        for (int k = 0; k < 8; k++) { const int i = tid + k * RF_WORK; if (i < n) { S[fsw(i)] = pvv[k]; S[fsw(n + i)] = t0[k]; pv[i] = t0[k]; } }
    }
    rf_bar_work();
    if (dbg && tid == 0) dbg[1] = clock64();
    fft_smem<1, RF_WORK>(S, n2, twl, -1, tid, RF_WORK);
    if (dbg && tid == 0) dbg[2] = clock64();
    const int mask = nfor - 1;
    // partition MAC, four bins at a time with every load of the four issued before the arithmetic: the older spectra and the
    // masks come from L2, and one bin after another would pay that latency sixteen times over
    for (int i0 = tid; i0 < n2; i0 += 4 * RF_WORK) {
        cd X[4], acc[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = i0 + u * RF_WORK;
            if (i < n2) {
                X[u] = S[fsw(i)];
                const cd m0 = fmask[i];
                acc[u] = make_double2(X[u].x * m0.x - X[u].y * m0.y, X[u].x * m0.y + X[u].y * m0.x);
            }
        }
        int k = bi;
        for (int j = 1; j < nfor; j += 3) {
            // up to three older spectra x four bins: 24 loads in flight per thread, then the arithmetic in the reference's order
            cd Y[3][4], mm[3][4];
            int kk = k;
#pragma unroll
            for (int jj = 0; jj < 3; jj++) {
                kk = (kk + mask) & mask;
                if (j + jj < nfor) {
                    const cd *fk = fd + (size_t)kk * n2 + i0, *mk = fmask + (size_t)(j + jj) * n2 + i0;
#pragma unroll
                    for (int u = 0; u < 4; u++) { if (i0 + u * RF_WORK < n2) { Y[jj][u] = fk[u * RF_WORK]; mm[jj][u] = mk[u * RF_WORK]; } }
                }
            }
#pragma unroll
            for (int jj = 0; jj < 3; jj++) {
                if (j + jj < nfor) {
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (i0 + u * RF_WORK < n2) {
                            acc[u].x += Y[jj][u].x * mm[jj][u].x - Y[jj][u].y * mm[jj][u].y;
                            acc[u].y += Y[jj][u].x * mm[jj][u].y + Y[jj][u].y * mm[jj][u].x;
                        }
                    }
                }
            }
            k = kk;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i = i0 + u * RF_WORK; if (i < n2) { fd[(size_t)bi * n2 + i] = X[u]; S[fsw(i)] = acc[u]; } }
    }
    rf_bar_work();
    if (dbg && tid == 0) dbg[3] = clock64();
    fft_smem<1, RF_WORK>(S, n2, twl, +1, tid, RF_WORK);
    if (dbg && tid == 0) dbg[4] = clock64();
}

__device__ __forceinline__ void rf_worker_role(const RxaFusedParams &P, const RfSmem &m, int n, int c, int tid, bool agc_on, int ab, double *hs)
{
    const int tot = ab + n;
    cd *S = m.S;
    double *A = m.A, *RV = m.RV, *BM = m.BM;
    const cd *x = P.in + (size_t)c * P.in_stride;
    cd *y = P.out + (size_t)c * P.out_stride;
    const int per = (n + RF_WORK - 1) / RF_WORK;            // consecutive samples per worker in the ring_max phase (<= 8)
    rf_bar_all();
    for (int b = 0; b < P.nblocks; b++, x += n, y += n) {
        const bool stamp = P.dbg && c == 0 && b == P.nblocks - 1 && tid == 0;
        if (stamp) P.dbg[0] = clock64();
        // the history's magnitudes go into A right away (nobody reads A before the ring_max phase): their L2 round trip
        // overlaps the transforms instead of standing in front of that phase
        if (agc_on) for (int i = tid; i < ab; i += RF_WORK) A[i] = hs[i * 3 + 2];
        if (P.n_fir > 0) rf_fircore(S, m.twl, P.prev[0] + (size_t)c * n, P.fdl[0] + (size_t)c * P.nfor[0] * 2 * n, P.mask[0], P.nfor[0],
                                    (P.buffidx[0] + b) & (P.nfor[0] - 1), n, tid, x, RV, stamp ? P.dbg + 16 : nullptr);
        if (P.n_fir > 1) rf_fircore(S, m.twl, P.prev[1] + (size_t)c * n, P.fdl[1] + (size_t)c * P.nfor[1] * 2 * n, P.mask[1], P.nfor[1],
                                    (P.buffidx[1] + b) & (P.nfor[1] - 1), n, tid, nullptr, nullptr, nullptr);
        if (P.n_fir == 0) {
            const cd *xa = P.in_adc ? P.in_adc + (size_t)c * P.adc_stride + (size_t)b * n : x;
            cd xv[8], av[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { const int i = tid + k * RF_WORK; if (i < n) { xv[k] = x[i]; av[k] = xa[i]; } }
#pragma unroll
            for (int k = 0; k < 8; k++) { const int i = tid + k * RF_WORK; if (i < n) { S[fsw(i)] = xv[k]; RV[i] = rf_smag(av[k]); } }
            rf_bar_work();
        }
        if (stamp) P.dbg[1] = clock64();
        // ---- meter inputs, magnitudes
        double mx_adc = 0.0, mx_s = 0.0;
        for (int i = tid; i < n; i += RF_WORK) {
            const cd v = S[fsw(i)];
            const double sm = rf_smag(v), sa = RV[i];         // |x|^2 was parked in RV when the block was loaded
            m.SMS[i] = sm; m.SMADC[i] = sa;
            mx_s = sm > mx_s ? sm : mx_s; mx_adc = sa > mx_adc ? sa : mx_adc;
            if (agc_on) {
                double mg;
                if (P.a.pmode == 0) { const double f0 = fabs(v.x), f1 = fabs(v.y); mg = f0 < f1 ? f1 : f0; }
                else mg = __dsqrt_rn(sm);
                A[ab + i] = mg;
            }
        }
        rf_block_max(mx_adc, m.SC, m.SC + 4, tid);
        rf_bar_work();
        rf_block_max(mx_s, m.SC, m.SC + 5, tid);
        // ---- ring_max[i] = max A[i + 1 .. i + ab] (wcpAGC.c:196-210 keeps it lazily; its value is exactly this).  Each worker
        // takes `per` consecutive samples: their windows share the core [i0 + per, i0 + ab], found once from 32-sample block
        // maxima; each sample then adds its own few head and tail elements.
        if (agc_on) {
            const int nblk = (tot + 31) >> 5;
            for (int j = tid; j < nblk; j += RF_WORK) {
                double mx = 0.0;
                const int e = min(tot, (j + 1) << 5);
                for (int k = j << 5; k < e; k++) mx = A[k] > mx ? A[k] : mx;
                BM[j] = mx;
            }
            rf_bar_work();
            const int i0 = tid * per;
            if (i0 < n) {
                const int cnt = min(per, n - i0);
                if (ab >= per) {
                    const int lo = i0 + per, hi = i0 + ab;          // the shared core (lo <= hi + 1)
                    double core = 0.0;
                    const int b0 = (lo + 31) >> 5, b1 = (hi + 1) >> 5;
                    if (b0 >= b1) {
                        for (int k = lo; k <= hi; k++) core = A[k] > core ? A[k] : core;
                    } else {
                        for (int k = lo; k < (b0 << 5); k++) core = A[k] > core ? A[k] : core;
                        for (int j = b0; j < b1; j++) core = BM[j] > core ? BM[j] : core;
                        for (int k = b1 << 5; k <= hi; k++) core = A[k] > core ? A[k] : core;
                    }
                    // sample i0 + k: head A[i0 + k + 1 .. i0 + per - 1] (a suffix maximum), tail A[i0 + ab + 1 .. i0 + ab + k] (a prefix maximum)
                    double hd[8], tl[8];
                    double run = 0.0;
#pragma unroll
                    for (int k = 7; k >= 0; k--) { if (k < per) { hd[k] = run; const int q = i0 + k; if (k > 0 && q < tot) run = A[q] > run ? A[q] : run; } }
                    run = 0.0;
#pragma unroll
                    for (int k = 0; k < 8; k++) { if (k < per) { if (k > 0) { const int q = i0 + ab + k; if (q < tot) run = A[q] > run ? A[q] : run; } tl[k] = run; } }
#pragma unroll
                    for (int k = 0; k < 8; k++) if (k < cnt) { double r = core > hd[k] ? core : hd[k]; r = tl[k] > r ? tl[k] : r; RV[i0 + k] = r; }
                } else {
                    for (int k = 0; k < cnt; k++) { double r = 0.0; for (int q = i0 + k + 1; q <= i0 + k + ab; q++) r = A[q] > r ? A[q] : r; RV[i0 + k] = r; }
                }
            }
        }
        if (stamp) P.dbg[2] = clock64();
        rf_bar_all();                                       // -> the sequential lanes
        if (stamp) P.dbg[12] = clock64();
        rf_bar_all();                                       // <- volts
        if (stamp) P.dbg[6] = clock64();
        // ---- gain law, panel, output, siphon, AGC history
        double mx_agc = 0.0;
        cd dl[8];                                           // the delayed samples: history (behind L2) or this block, all asked for first
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = tid + k * RF_WORK;
            if (i < n) {
                if (agc_on && i < ab) dl[k] = make_double2(hs[i * 3], hs[i * 3 + 1]);
                else dl[k] = S[fsw(agc_on ? i - ab : i)];
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = tid + k * RF_WORK;
            if (i >= n) continue;
            cd o;
            const cd d = dl[k];
            if (agc_on) {
                const double volts = RV[i];
                const double lg = log10(__dmul_rn(P.a.inv_max_input, volts));
                const double mult = __ddiv_rn(__dsub_rn(P.a.out_target, __dmul_rn(P.a.slope_constant, 0.0 < lg ? 0.0 : lg)), volts);
                o = make_double2(__dmul_rn(d.x, mult), __dmul_rn(d.y, mult));
            } else {
                o = P.agc_run ? make_double2(__dmul_rn(P.a.fixed_gain, d.x), __dmul_rn(P.a.fixed_gain, d.y)) : d;
            }
            const double sm = rf_smag(o);
            m.SMAGC[i] = sm;
            mx_agc = sm > mx_agc ? sm : mx_agc;
            if (P.sip) {
                if (n >= P.sipsize) { if (i >= n - P.sipsize) P.sip[(size_t)c * P.sipsize + (i - (n - P.sipsize))] = o; }
                else P.sip[(size_t)c * P.sipsize + ((P.sip_idx + b * n + i) & (P.sipsize - 1))] = o;
            }
            y[i] = make_double2(__dmul_rn(P.gI, o.x), __dmul_rn(P.gQ, o.y));
        }
        rf_block_max(mx_agc, m.SC, m.SC + 6, tid);
        rf_bar_work();
        if (agc_on) {
            // new history = combined[n, n + ab): the old history shifted down by n (only when n < ab), then the block's tail
            if (n < ab) {
                const int keep = ab - n;
                for (int j0 = 0; j0 < keep; j0 += RF_WORK) {
                    const int j = j0 + tid;
                    double r0 = 0, r1 = 0;
                    if (j < keep) { r0 = hs[(n + j) * 3]; r1 = hs[(n + j) * 3 + 1]; }
                    rf_bar_work();
                    if (j < keep) { hs[j * 3] = r0; hs[j * 3 + 1] = r1; hs[j * 3 + 2] = A[n + j]; }
                    rf_bar_work();
                }
            }
            for (int j = tid + max(ab - n, 0); j < ab; j += RF_WORK) {
                const cd v = S[fsw(n + j - ab)];
                hs[j * 3] = v.x; hs[j * 3 + 1] = v.y; hs[j * 3 + 2] = A[n + j];
            }
        }
        if (stamp) P.dbg[7] = clock64();
        rf_bar_all();                                       // end of block
    }
}

template <int MINB>
__global__ void __launch_bounds__(RF_THREADS, MINB) rxa_ssb_fused_kernel(RxaFusedParams P)
{
    extern __shared__ double smem_raw[];
    const int n = P.n, c = blockIdx.x, tid = threadIdx.x;
    const bool agc_on = P.agc_run && P.a.mode != 0;
    const int ab = agc_on ? P.a.attack_buffsize : 0;
    const RfSmem m = rf_map(smem_raw, n, ab + n);
    fft_stage_twiddles(m.twl, P.tw, 2 * n);
    double *hs = P.agc_hist + (size_t)c * (agc_on ? ab : 1) * 3;
    double *ast = P.agc_state + (size_t)c * 16;
    if (tid < RF_WORK) rf_worker_role(P, m, n, c, tid, agc_on, ab, hs);
    else if (tid < RF_WORK + 32) rf_agc_role(P, m, n, agc_on, ast, tid == RF_WORK);
    else rf_lin_role(P, m, n, c, tid - RF_WORK - 32, agc_on, ast);
}


// ---- the wide path: many DSP blocks per launch.  The overlap-save filter only carries state through its stored spectra,
// so the transforms of ALL blocks of ALL channels are independent: a grid of (channel, block) CTAs does every forward
// transform (fir_fwd_kernel), a second one every partition MAC + inverse transform (fir_mac_kernel: block b adds the spectra
// of blocks b, b - 1, ... b - nfor + 1, reaching back into the frequency-domain delay line of the previous launch for the
// first few), and only what is truly sequential along time -- the AGC's volts machine and the meters' averages -- runs one
// CTA per channel over the blocks (the kernel above with n_fir = 0).  Same arithmetic per block as xfircore
// (wdsp/firmin.c:409-430) in the same order.
__global__ void __launch_bounds__(256) fir_fwd_kernel(const cd *in, long in_stride, int n, const cd *prev /*[C][n]*/, cd *spec /*[C][nblocks][2n]*/, int nblocks, const cd *tw)
{
    extern __shared__ double smem_raw[];
    const int n2 = 2 * n, c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *S = twl + fft_tw_entries(n2);
    fft_stage_twiddles(twl, tw, n2);
    const cd *x = in + (size_t)c * in_stride + (size_t)b * n;
    const cd *pv = b == 0 ? prev + (size_t)c * n : x - n;
    for (int i = tid; i < n; i += nt) { S[fsw(i)] = pv[i]; S[fsw(n + i)] = x[i]; }
    __syncthreads();
    fft_smem<1>(S, n2, twl, -1, tid, nt);
    cd *sp = spec + ((size_t)c * nblocks + b) * n2;
    for (int i = tid; i < n2; i += nt) sp[i] = S[fsw(i)];
}

__device__ __forceinline__ cd fir_mac_first(cd X, cd m) { return make_double2(X.x * m.x - X.y * m.y, X.x * m.y + X.y * m.x); }
__device__ __forceinline__ void fir_mac_add(cd &acc, cd Y, cd m)
{
    acc.x += Y.x * m.x - Y.y * m.y;
    acc.y += Y.x * m.y + Y.y * m.x;
}

__global__ void __launch_bounds__(256) fir_mac_kernel(const cd *spec, int nblocks, int n, int nfor, int bi0, const cd *fdl /*[C][nfor][2n]*/,
                                                       const cd *__restrict__ fmask, cd *out, long out_stride, const cd *tw)
{
    extern __shared__ double smem_raw[];
    const int n2 = 2 * n, c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *S = twl + fft_tw_entries(n2);
    fft_stage_twiddles(twl, tw, n2);
    const int mask = nfor - 1;
    const cd *sp = spec + (size_t)c * nblocks * n2;
    const cd *fd = fdl + (size_t)c * nfor * n2;
    for (int i = tid; i < n2; i += nt) {
        const cd X = sp[(size_t)b * n2 + i];
        const cd m0 = fmask[i];
        cd acc = fir_mac_first(X, m0);
        for (int j = 1; j < nfor; j++) {
            // the spectrum of j blocks ago: this launch's if there is one, else the delay line's (ring index bi0 is block 0's slot)
            const cd Y = b - j >= 0 ? sp[(size_t)(b - j) * n2 + i] : fd[(size_t)((bi0 + b - j) & mask) * n2 + i];
            const cd m = fmask[(size_t)j * n2 + i];
            fir_mac_add(acc, Y, m);
        }
        S[fsw(i)] = acc;
    }
    __syncthreads();
    fft_smem<1>(S, n2, twl, +1, tid, nt);
    cd *y = out + (size_t)c * out_stride + (size_t)b * n;
    for (int i = tid; i < n; i += nt) y[i] = S[fsw(i)];
}

// The same for G consecutive blocks of a channel in one CTA.  fir_mac_kernel reads nfor spectra and nfor masks per block and
// bin: with everything in L2 that IS its cost (C4: 8 partitions, 4 TB/s of L2 reads).  Here a thread walks the G + nfor - 1
// spectra its bin needs ONCE, newest first, and keeps G accumulators: the spectrum of block bb meets block g0 + g at
// partition p = g0 + g - bb, so every accumulator still receives its partitions in the order 0, 1, 2, ... (the order of
// xfircore, firmin.c:417-427) -- bit-identical sums.  The masks slide through a G-long register window (one new mask value
// per step).  Then G inverse transforms side by side.  Loads per bin: G + 2 nfor - 1 instead of 2 G nfor.
template <int G>
__global__ void __launch_bounds__(256) fir_mac_group_kernel(const cd *spec, int nblocks, int n, int nfor, int bi0, const cd *fdl, const cd *__restrict__ fmask,
                                                             cd *out, long out_stride, const cd *tw, int lanes)
{
    extern __shared__ double smem_raw[];
    const int n2 = 2 * n, c = blockIdx.y, g0 = blockIdx.x * G, tid = threadIdx.x, nt = blockDim.x;
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *S = twl + fft_tw_entries(n2);
    fft_stage_twiddles(twl, tw, n2);
    const int mask = nfor - 1;
    const cd *sp = spec + (size_t)c * nblocks * n2;
    const cd *fd = fdl + (size_t)c * nfor * n2;
    const int steps = G + nfor - 1;
    for (int i = tid; i < n2; i += nt) {
        cd acc[G], mw[G];
#pragma unroll
        for (int g = 0; g < G; g++) { acc[g] = make_double2(0.0, 0.0); mw[g] = make_double2(0.0, 0.0); }
        mw[G - 1] = fmask[i];
        for (int k = 0; k < steps; k++) {
            const int bb = g0 + G - 1 - k;
            const bool have = bb < nblocks;
            cd Y = make_double2(0.0, 0.0);
            if (have) Y = bb >= 0 ? sp[(size_t)bb * n2 + i] : fd[(size_t)((bi0 + bb) & mask) * n2 + i];
#pragma unroll
            for (int g = 0; g < G; g++) {
                const int p = k + g - (G - 1);
                if (have && p >= 0 && p < nfor) {
                    if (p == 0) acc[g] = fir_mac_first(Y, mw[g]);
                    else fir_mac_add(acc[g], Y, mw[g]);
                }
            }
#pragma unroll
            for (int g = 0; g + 1 < G; g++) mw[g] = mw[g + 1];
            mw[G - 1] = k + 1 < nfor ? fmask[(size_t)(k + 1) * n2 + i] : make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int g = 0; g < G; g++) if (g0 + g < nblocks) S[(size_t)g * n2 + fsw(i)] = acc[g];
    }
    __syncthreads();
    const int sg = tid / lanes, ln = tid - sg * lanes;
    fft_smem<1>(S + (size_t)sg * n2, n2, twl, +1, ln, lanes);
    if (g0 + sg < nblocks) {
        cd *y = out + (size_t)c * out_stride + (size_t)(g0 + sg) * n;
        const cd *Sg = S + (size_t)sg * n2;
        for (int i = ln; i < n; i += lanes) y[i] = Sg[fsw(i)];
    }
}

// The partition MAC on its own, streamed along time: one thread per (channel, bin) walks the spectra of its bin from the
// NEWEST block of the launch down into the delay line of the previous one, each read exactly once, with NFOR accumulators
// in flight: the spectrum of block t is partition p of block t + p, and walking backwards every block receives its
// partitions in the order 0, 1, 2, ... of xfircore (firmin.c:417-427) -- bit-identical sums.  Block t + NFOR - 1 is complete
// when spectrum t has been added and goes to `prod`.  The masks are NFOR register values per thread.  HBM / L2 traffic:
// one read and one write per spectrum value, where a CTA per (channel, block) reads NFOR spectra and NFOR masks.
template <int NFOR>
__global__ void __launch_bounds__(256) fir_mac_stream_kernel(const cd *spec, int nblocks, int n2, int bi0, const cd *fdl, const cd *__restrict__ fmask, cd *prod)
{
    const int i = blockIdx.x * 256 + threadIdx.x, c = blockIdx.y;
    if (i >= n2) return;
    const cd *sp = spec + (size_t)c * nblocks * n2 + i;
    const cd *fd = fdl + (size_t)c * NFOR * n2 + i;
    cd *pr = prod + (size_t)c * nblocks * n2 + i;
    cd m[NFOR], acc[NFOR];
#pragma unroll
    for (int p = 0; p < NFOR; p++) { m[p] = fmask[(size_t)p * n2 + i]; acc[p] = make_double2(0.0, 0.0); }
    const int total = nblocks + NFOR - 1;                       // spectra walked: blocks nblocks - 1 .. -(NFOR - 1)
    for (int u0 = 0; u0 < total; u0 += NFOR) {
        cd Y[NFOR];
#pragma unroll
        for (int r = 0; r < NFOR; r++) {
            const int t = nblocks - 1 - (u0 + r);
            if (u0 + r < total) Y[r] = t >= 0 ? sp[(size_t)t * n2] : fd[(size_t)((bi0 + t) & (NFOR - 1)) * n2];
        }
#pragma unroll
        for (int r = 0; r < NFOR; r++) {
            const int u = u0 + r, t = nblocks - 1 - u;
            if (u < total) {
#pragma unroll
                for (int p = 0; p < NFOR; p++) {
                    // block t + p, accumulator (u - p) mod NFOR = (r - p) mod NFOR
                    const int a = (r - p + NFOR) % NFOR;
                    if (p <= u && t + p >= 0) {
                        if (p == 0) acc[a] = fir_mac_first(Y[r], m[0]);
                        else fir_mac_add(acc[a], Y[r], m[p]);
                        if (p == NFOR - 1) pr[(size_t)(t + p) * n2] = acc[a];
                    }
                }
            }
        }
    }
}

// inverse transforms of the finished products, G blocks of a channel side by side in one CTA
template <int G>
__global__ void __launch_bounds__(256) fir_inv_group_kernel(const cd *prod, int nblocks, int n, cd *out, long out_stride, const cd *tw, int lanes)
{
    extern __shared__ double smem_raw[];
    const int n2 = 2 * n, c = blockIdx.y, g0 = blockIdx.x * G, tid = threadIdx.x;
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *S = twl + fft_tw_entries(n2);
    fft_stage_twiddles(twl, tw, n2);
    const int sg = tid / lanes, ln = tid - sg * lanes;
    cd *Sg = S + (size_t)sg * n2;
    if (g0 + sg < nblocks) {
        const cd *p = prod + ((size_t)c * nblocks + g0 + sg) * n2;
        for (int i = ln; i < n2; i += lanes) Sg[fsw(i)] = p[i];
    }
    __syncthreads();
    fft_smem<1>(Sg, n2, twl, +1, ln, lanes);
    if (g0 + sg < nblocks) {
        cd *y = out + (size_t)c * out_stride + (size_t)(g0 + sg) * n;
        for (int i = ln; i < n; i += lanes) y[i] = Sg[fsw(i)];
    }
}

// one fircore over nblocks blocks of every channel: transforms wide, then the delay line and `prev` brought up to date
int fircore_wide(FirCore *f, const cd *in, long in_stride, cd *out, long out_stride, int nblocks, cd *spec, cudaStream_t s)
{
    const int n = f->size, n2 = 2 * n, lanes = fft_threads(n2), C = f->C, nfor = f->nfor;
    const size_t sh = ((size_t)n2 + fft_tw_entries(n2)) * sizeof(cd);
    if (sh > 48 * 1024) {
        QC_CUDA(cudaFuncSetAttribute(fir_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        QC_CUDA(cudaFuncSetAttribute(fir_mac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    }
    fir_fwd_kernel<<<dim3(C, nblocks), lanes, sh, s>>>(in, in_stride, n, f->d_prev, spec, nblocks, f->tw);
    count_launch();
    QC_CUDA_LAUNCH();
    // the last input block becomes `prev` now: the forward transforms have read the old one, and with in == out (second filter
    // of a chain) the MAC kernel is about to overwrite the input
    QC_CUDA(cudaMemcpy2DAsync(f->d_prev, (size_t)n * sizeof(cd), in + (size_t)(nblocks - 1) * n, (size_t)in_stride * sizeof(cd), (size_t)n * sizeof(cd), C,
                              cudaMemcpyDeviceToDevice, s));
    const int G = n2 <= 512 ? 8 : (n2 <= 1024 ? 4 : (n2 <= 2048 ? 2 : 1));
    const char *macform = getenv("QUISK_FIR_MAC");           // "single" / "group" / "stream" force one form (tests, measurements)
    const bool can_stream = (nfor == 2 || nfor == 4 || nfor == 8 || nfor == 16) && C <= 65535 && G * lanes <= 256;
    if (can_stream && !(macform && strcmp(macform, "stream"))) {
        // `spec` is allocated twice as long as the spectra need: the second half takes the products
        cd *prod = spec + (size_t)C * nblocks * n2;
        const dim3 gm((n2 + 255) / 256, C);
        switch (nfor) {
        case 2: fir_mac_stream_kernel<2><<<gm, 256, 0, s>>>(spec, nblocks, n2, f->buffidx, f->d_fdl, f->d_mask[f->cset], prod); break;
        case 4: fir_mac_stream_kernel<4><<<gm, 256, 0, s>>>(spec, nblocks, n2, f->buffidx, f->d_fdl, f->d_mask[f->cset], prod); break;
        case 8: fir_mac_stream_kernel<8><<<gm, 256, 0, s>>>(spec, nblocks, n2, f->buffidx, f->d_fdl, f->d_mask[f->cset], prod); break;
        default: fir_mac_stream_kernel<16><<<gm, 256, 0, s>>>(spec, nblocks, n2, f->buffidx, f->d_fdl, f->d_mask[f->cset], prod); break;
        }
        count_launch();
        QC_CUDA_LAUNCH();
        const size_t shg = ((size_t)G * n2 + fft_tw_entries(n2)) * sizeof(cd);
        const dim3 grid((nblocks + G - 1) / G, C);
#define QC_INVG(G_) do { \
            if (shg > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(fir_inv_group_kernel<G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shg)); \
            fir_inv_group_kernel<G_><<<grid, G_ * lanes, shg, s>>>(prod, nblocks, n, out, out_stride, f->tw, lanes); } while (0)
        if (G == 8) QC_INVG(8); else if (G == 4) QC_INVG(4); else if (G == 2) QC_INVG(2); else QC_INVG(1);
#undef QC_INVG
    } else if (G > 1 && nfor > 1 && C <= 65535 && G * lanes <= 256 && !(macform && !strcmp(macform, "single"))) {
        const size_t shg = ((size_t)G * n2 + fft_tw_entries(n2)) * sizeof(cd);
        const dim3 grid((nblocks + G - 1) / G, C);
#define QC_MACG(G_) do { \
            if (shg > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(fir_mac_group_kernel<G_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shg)); \
            fir_mac_group_kernel<G_><<<grid, G_ * lanes, shg, s>>>(spec, nblocks, n, nfor, f->buffidx, f->d_fdl, f->d_mask[f->cset], out, out_stride, f->tw, lanes); } while (0)
        if (G == 8) QC_MACG(8); else if (G == 4) QC_MACG(4); else QC_MACG(2);
#undef QC_MACG
    } else {
        fir_mac_kernel<<<dim3(C, nblocks), lanes, sh, s>>>(spec, nblocks, n, nfor, f->buffidx, f->d_fdl, f->d_mask[f->cset], out, out_stride, f->tw);
    }
    count_launch();
    QC_CUDA_LAUNCH();
    // state for the next call: the newest min(nblocks, nfor) spectra into their ring slots
    for (int b = nblocks > nfor ? nblocks - nfor : 0; b < nblocks; b++) {
        const int slot = (f->buffidx + b) & (nfor - 1);
        QC_CUDA(cudaMemcpy2DAsync(f->d_fdl + (size_t)slot * n2, (size_t)nfor * n2 * sizeof(cd), spec + (size_t)b * n2, (size_t)nblocks * n2 * sizeof(cd),
                                  (size_t)n2 * sizeof(cd), C, cudaMemcpyDeviceToDevice, s));
    }
    f->buffidx = (f->buffidx + nblocks) & (nfor - 1);
    return QC_OK;
}

// ---- the wide path's sequential part, pipelined --------------------------------------------------------------------------
// With the filters done for all blocks (fircore_wide), what is left per block is: (A) magnitudes, block maxima and the AGC's
// sliding maximum -- parallel along time AND across blocks; (S) the volts machine and the meters' recurrences -- sequential;
// (B) gain law, panel, output -- parallel, needs volts.  The one-kernel form above runs A, S, B one after the other for
// every block (72 k cycles per 1024-sample block, 41 k of them S).  Here A is a grid of (channel, block) CTAs of its own
// (rxa_wide_pre_kernel: ring_max, |y|^2, |x|^2, the delayed magnitudes and the block maxima go to global scratch), and the
// per-channel kernel (rxa_wide_seq_kernel) is a two-stage pipeline over double-buffered shared memory: while the AGC / LIN
// warps walk block t, the workers apply the gain law to block t - 1 and fetch block t + 1's inputs.  The sequential lanes
// are then the only thing on the critical path.  Same arithmetic, same order, same state arrays as the kernel above.
struct RxaWideParams {
    const cd *y; long y_stride;             // the filtered stream [C][nblocks * n]
    const cd *xadc; long adc_stride;        // the raw input (ADC meter)
    cd *out; long out_stride;
    int n, nblocks, C;
    double *rm, *sms, *smadc, *absd;        // [C][nblocks * n]: ring_max, |y|^2, |x|^2, |sample leaving the AGC delay line|
    double *np;                             // [C][nblocks][2]: block maxima of |x|^2, |y|^2
    double *mst[3]; double *mres[3]; double m_ma, m_mp; const double *mtable;
    int agc_run; double *agc_state; double *agc_hist; AgcParams a;
    double gI, gQ; cd *sip; int sipsize, sip_idx;
};

__device__ __forceinline__ double rf_mag(cd v, int pmode)
{
    if (pmode == 0) { const double f0 = fabs(v.x), f1 = fabs(v.y); return f0 < f1 ? f1 : f0; }
    return __dsqrt_rn(rf_smag(v));
}

__global__ void __launch_bounds__(RF_WORK) rxa_wide_pre_kernel(RxaWideParams P)
{
    extern __shared__ double smem_raw[];
    const int n = P.n, c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const bool agc_on = P.agc_run && P.a.mode != 0;
    const int ab = agc_on ? P.a.attack_buffsize : 0, tot = ab + n;
    double *A = smem_raw, *BM = A + tot, *SC = BM + ((tot + 31) / 32 + 1);
    const cd *ys = P.y + (size_t)c * P.y_stride, *x = ys + (size_t)b * n;
    const cd *xa = P.xadc + (size_t)c * P.adc_stride + (size_t)b * n;
    const double *hs = P.agc_hist + (size_t)c * (agc_on ? ab : 1) * 3;
    const size_t off = ((size_t)c * P.nblocks + b) * n;
    double *rm = P.rm + off, *sms = P.sms + off, *smadc = P.smadc + off, *absd = P.absd + off;
    // [history | block] magnitudes: the history is the stream's own past, or the saved ring in front of the launch
    for (int j = tid; j < ab; j += RF_WORK) {
        const long idx = (long)b * n - ab + j;
        A[j] = idx >= 0 ? rf_mag(ys[idx], P.a.pmode) : hs[(idx + ab) * 3 + 2];
    }
    double mx_adc = 0.0, mx_s = 0.0;
    for (int i = tid; i < n; i += RF_WORK) {
        const cd v = x[i], av = xa[i];
        const double sm = rf_smag(v), sa = rf_smag(av);
        sms[i] = sm; smadc[i] = sa;
        mx_s = sm > mx_s ? sm : mx_s; mx_adc = sa > mx_adc ? sa : mx_adc;
        if (agc_on) A[ab + i] = P.a.pmode == 0 ? rf_mag(v, 0) : __dsqrt_rn(sm);
    }
    rf_block_max(mx_adc, SC, SC + 4, tid);
    rf_bar_work();
    rf_block_max(mx_s, SC, SC + 5, tid);
    rf_bar_work();
    if (tid == 0) { P.np[((size_t)c * P.nblocks + b) * 2] = SC[4]; P.np[((size_t)c * P.nblocks + b) * 2 + 1] = SC[5]; }
    if (!agc_on) return;
    for (int i = tid; i < n; i += RF_WORK) absd[i] = A[i];
    // ring_max[i] = max A[i + 1 .. i + ab], as in rf_worker_role
    const int per = (n + RF_WORK - 1) / RF_WORK;
    const int nblk = (tot + 31) >> 5;
    for (int j = tid; j < nblk; j += RF_WORK) {
        double mx = 0.0;
        const int e = min(tot, (j + 1) << 5);
        for (int k = j << 5; k < e; k++) mx = A[k] > mx ? A[k] : mx;
        BM[j] = mx;
    }
    rf_bar_work();
    const int i0 = tid * per;
    if (i0 < n) {
        const int cnt = min(per, n - i0);
        if (ab >= per) {
            const int lo = i0 + per, hi = i0 + ab;
            double core = 0.0;
            const int b0 = (lo + 31) >> 5, b1 = (hi + 1) >> 5;
            if (b0 >= b1) {
                for (int k = lo; k <= hi; k++) core = A[k] > core ? A[k] : core;
            } else {
                for (int k = lo; k < (b0 << 5); k++) core = A[k] > core ? A[k] : core;
                for (int j = b0; j < b1; j++) core = BM[j] > core ? BM[j] : core;
                for (int k = b1 << 5; k <= hi; k++) core = A[k] > core ? A[k] : core;
            }
            double hd[8], tl[8];
            double run = 0.0;
#pragma unroll
            for (int k = 7; k >= 0; k--) { if (k < per) { hd[k] = run; const int q = i0 + k; if (k > 0 && q < tot) run = A[q] > run ? A[q] : run; } }
            run = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) { if (k < per) { if (k > 0) { const int q = i0 + ab + k; if (q < tot) run = A[q] > run ? A[q] : run; } tl[k] = run; } }
#pragma unroll
            for (int k = 0; k < 8; k++) if (k < cnt) { double r = core > hd[k] ? core : hd[k]; r = tl[k] > r ? tl[k] : r; rm[i0 + k] = r; }
        } else {
            for (int k = 0; k < cnt; k++) { double r = 0.0; for (int q = i0 + k + 1; q <= i0 + k + ab; q++) r = A[q] > r ? A[q] : r; rm[i0 + k] = r; }
        }
    }
}

struct RwSmem {
    double *RV[2], *SMS[2], *SMADC[2], *ABSD[2], *SMAGC[2], *FBA, *HBA, *SC;
    volatile int *linpos;
};
__device__ __forceinline__ RwSmem rw_map(double *raw, int n)
{
    RwSmem m;
    double *p = raw;
    for (int k = 0; k < 2; k++) { m.RV[k] = p; p += n; m.SMS[k] = p; p += n; m.SMADC[k] = p; p += n; m.ABSD[k] = p; p += n; m.SMAGC[k] = p; p += n; }
    m.FBA = p; p += n; m.HBA = p; p += n;
    m.SC = p; p += 64;              // 0..3 warp maxima; 8 + 3 * parity + {0 adc, 1 s, 2 agc}: block maxima; 16..47 LIN lanes' dummy store targets
    m.linpos = reinterpret_cast<volatile int *>(p);
    return m;
}
static size_t rxa_wide_seq_smem(int n) { return ((size_t)12 * n + 64 + 2) * sizeof(double); }

// block t's inputs from the pre kernel's scratch into the shared buffers of its parity
__device__ __forceinline__ void rw_load_block(const RxaWideParams &P, const RwSmem &m, int n, int c, int t, int tid, bool agc_on)
{
    const size_t off = ((size_t)c * P.nblocks + t) * n;
    const int par = t & 1;
    double v[4][8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int i = tid + k * RF_WORK;
        if (i < n) {
            v[0][k] = P.sms[off + i]; v[1][k] = P.smadc[off + i];
            if (agc_on) { v[2][k] = P.rm[off + i]; v[3][k] = P.absd[off + i]; }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int i = tid + k * RF_WORK;
        if (i < n) {
            m.SMS[par][i] = v[0][k]; m.SMADC[par][i] = v[1][k];
            if (agc_on) { m.RV[par][i] = v[2][k]; m.ABSD[par][i] = v[3][k]; }
        }
    }
    if (tid < 2) m.SC[8 + 3 * par + tid] = P.np[((size_t)c * P.nblocks + t) * 2 + tid];
}

// gain law, panel, output, siphon for block tb (its volts are in RV[tb & 1])
__device__ __forceinline__ void rw_gain_block(const RxaWideParams &P, const RwSmem &m, int n, int c, int tb, int tid, bool agc_on, int ab, const double *hs)
{
    const int par = tb & 1;
    const double *RV = m.RV[par];
    double *SMAGC = m.SMAGC[par];
    const cd *ys = P.y + (size_t)c * P.y_stride;
    cd *y = P.out + (size_t)c * P.out_stride + (size_t)tb * n;
    double mx_agc = 0.0;
    cd dl[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int i = tid + k * RF_WORK;
        if (i < n) {
            const long idx = (long)tb * n + i - ab;         // the sample leaving the AGC's delay line
            dl[k] = idx >= 0 ? ys[idx] : make_double2(hs[(idx + ab) * 3], hs[(idx + ab) * 3 + 1]);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int i = tid + k * RF_WORK;
        if (i >= n) continue;
        cd o;
        const cd d = dl[k];
        if (agc_on) {
            const double volts = RV[i];
            const double lg = log10(__dmul_rn(P.a.inv_max_input, volts));
            const double mult = __ddiv_rn(__dsub_rn(P.a.out_target, __dmul_rn(P.a.slope_constant, 0.0 < lg ? 0.0 : lg)), volts);
            o = make_double2(__dmul_rn(d.x, mult), __dmul_rn(d.y, mult));
        } else {
            o = P.agc_run ? make_double2(__dmul_rn(P.a.fixed_gain, d.x), __dmul_rn(P.a.fixed_gain, d.y)) : d;
        }
        const double sm = rf_smag(o);
        SMAGC[i] = sm;
        mx_agc = sm > mx_agc ? sm : mx_agc;
        if (P.sip) {
            if (n >= P.sipsize) { if (i >= n - P.sipsize) P.sip[(size_t)c * P.sipsize + (i - (n - P.sipsize))] = o; }
            else P.sip[(size_t)c * P.sipsize + ((P.sip_idx + tb * n + i) & (P.sipsize - 1))] = o;
        }
        y[i] = make_double2(__dmul_rn(P.gI, o.x), __dmul_rn(P.gQ, o.y));
    }
    rf_block_max(mx_agc, m.SC, m.SC + 8 + 3 * par + 2, tid);
}

__device__ __forceinline__ void rw_worker_role(const RxaWideParams &P, const RwSmem &m, int n, int c, int tid, bool agc_on, int ab, double *hs)
{
    const int nb = P.nblocks;
    rw_load_block(P, m, n, c, 0, tid, agc_on);
    rf_bar_all();                                           // set-up done
    for (int t = 0; t < nb; t++) {
        rf_bar_all();                                       // the sequential lanes start block t
        if (t >= 1) rw_gain_block(P, m, n, c, t - 1, tid, agc_on, ab, hs);
        rf_bar_work();                                      // RV[(t + 1) & 1] is free: every worker has read block t - 1's volts
        if (t + 1 < nb) rw_load_block(P, m, n, c, t + 1, tid, agc_on);
        rf_bar_all();                                       // block t's volts are final, block t + 1's inputs are in place
    }
    rf_bar_all();
    rw_gain_block(P, m, n, c, nb - 1, tid, agc_on, ab, hs);
    rf_bar_work();
    if (agc_on) {
        // the AGC ring for the next launch: the stream's last ab samples (older ones, when the launch was shorter than the ring, move down)
        const long N = (long)nb * n;
        const cd *ys = P.y + (size_t)c * P.y_stride;
        if (N < ab) {
            const int keep = ab - (int)N;
            for (int j0 = 0; j0 < keep; j0 += RF_WORK) {
                const int j = j0 + tid;
                double r0 = 0, r1 = 0, r2 = 0;
                if (j < keep) { r0 = hs[(N + j) * 3]; r1 = hs[(N + j) * 3 + 1]; r2 = hs[(N + j) * 3 + 2]; }
                rf_bar_work();
                if (j < keep) { hs[j * 3] = r0; hs[j * 3 + 1] = r1; hs[j * 3 + 2] = r2; }
                rf_bar_work();
            }
        }
        for (int k = tid + (N < ab ? ab - (int)N : 0); k < ab; k += RF_WORK) {
            const cd v = ys[N - ab + k];
            hs[k * 3] = v.x; hs[k * 3 + 1] = v.y; hs[k * 3 + 2] = rf_mag(v, P.a.pmode);
        }
    }
    rf_bar_all();
}

__device__ __forceinline__ void rw_agc_role(const RxaWideParams &P, const RwSmem &m, int n, bool agc_on, double *ast, bool lane0)
{
    double volts = ast[3], save_volts = ast[4];
    int hang_counter = (int)ast[7], decay_type = (int)ast[8], state_ = (int)ast[9];
    rf_bar_all();
    for (int t = 0; t < P.nblocks; t++) {
        rf_bar_all();
        if (agc_on) rf_agc_block(m.RV[t & 1], m.FBA, m.HBA, m.linpos, n, P.a, volts, save_volts, hang_counter, decay_type, state_, nullptr);
        if (lane0 && agc_on && t == P.nblocks - 1) {
            ast[3] = volts; ast[4] = save_volts;
            ast[7] = hang_counter; ast[8] = decay_type; ast[9] = state_; ast[10] = __dmul_rn(volts, P.a.inv_out_target);
        }
        rf_bar_all();
    }
    rf_bar_all();
    rf_bar_all();
}

__device__ __forceinline__ void rw_lin_role(const RxaWideParams &P, const RwSmem &m, int n, int c, int lane, bool agc_on, double *ast)
{
    // lanes as in rf_lin_role; the AGC meter's lanes run TWO blocks late here (block t - 1's gain law runs beside block t)
    const int mt = lane < 6 ? lane >> 1 : 3, pk = lane & 1, nb = P.nblocks;
    const bool meter = lane < 6, back = agc_on && (lane == 6 || lane == 7), live = meter || back;
    double s = meter ? P.mst[mt][(size_t)c * 2 + pk] : (back ? ast[lane - 1] : 0.0);
    double c1 = 0.0, c2 = 0.0;
    if (meter) { c1 = pk ? 0.0 : 1.0 - P.m_ma; c2 = pk ? P.m_mp : P.m_ma; }
    else if (lane == 6) { c1 = P.a.fast_backmult; c2 = P.a.onemfast_backmult; }
    else if (lane == 7) { c1 = P.a.hang_backmult; c2 = P.a.onemhang_backmult; }
    double *dst = lane == 6 ? m.FBA : (lane == 7 ? m.HBA : m.SC + 16 + lane);
    const int dstep = lane == 6 || lane == 7 ? 1 : 0;
    if (lane == 0) *m.linpos = 0;
    rf_bar_all();
    for (int t = 0; t < nb; t++) {
        rf_bar_all();
        const int par = t & 1;
        const double *src = mt == 0 ? m.SMADC[par] : (mt == 1 ? m.SMS[par] : (mt == 2 ? m.SMAGC[par] : (back ? m.ABSD[par] : m.SMS[par])));
        const bool run = mt != 2 || t >= 2;
        const double r = rf_lin_block(run ? src : m.SMS[par], n, run ? c1 : 0.0, run ? c2 : 1.0, s, dst, dstep, m.linpos);
        if (live) { s = r; if (meter && pk && run) { const double np = m.SC[8 + 3 * par + mt]; if (np > s) s = np; } }
        rf_bar_all();
        if (lane == 0) *m.linpos = 0;
    }
    // the AGC meter's last two blocks: nb - 2 beside the workers' last gain pass, nb - 1 after it
    rf_bar_all();
    if (nb >= 2) {
        const int par = nb & 1;
        const double r = rf_lin_block(m.SMAGC[par], n, mt == 2 ? c1 : 0.0, mt == 2 ? c2 : 1.0, s, m.SC + 16 + lane, 0, nullptr);
        if (mt == 2) { s = r; if (pk) { const double np = m.SC[8 + 3 * par + 2]; if (np > s) s = np; } }
    }
    rf_bar_all();
    {
        const int par = (nb - 1) & 1;
        const double r = rf_lin_block(m.SMAGC[par], n, mt == 2 ? c1 : 0.0, mt == 2 ? c2 : 1.0, s, m.SC + 16 + lane, 0, nullptr);
        if (mt == 2) { s = r; if (pk) { const double np = m.SC[8 + 3 * par + 2]; if (np > s) s = np; } }
    }
    if (meter) {
        P.mst[mt][(size_t)c * 2 + pk] = s;
        P.mres[mt][(size_t)c * 3 + pk] = 10.0 * mlog10_dev(P.mtable, s + 1.0e-40);
        if (lane == 4) P.mres[2][(size_t)c * 3 + 2] = 20.0 * mlog10_dev(P.mtable, ast[10] + 1.0e-40);
        else if (!pk) P.mres[mt][(size_t)c * 3 + 2] = 0.0;
    } else if (back) {
        ast[lane - 1] = s;
    }
}

template <int MINB>
__global__ void __launch_bounds__(RF_THREADS, MINB) rxa_wide_seq_kernel(RxaWideParams P)
{
    extern __shared__ double smem_raw[];
    const int n = P.n, c = blockIdx.x, tid = threadIdx.x;
    const bool agc_on = P.agc_run && P.a.mode != 0;
    const int ab = agc_on ? P.a.attack_buffsize : 0;
    const RwSmem m = rw_map(smem_raw, n);
    double *hs = P.agc_hist + (size_t)c * (agc_on ? ab : 1) * 3;
    double *ast = P.agc_state + (size_t)c * 16;
    if (tid < RF_WORK) rw_worker_role(P, m, n, c, tid, agc_on, ab, hs);
    else if (tid < RF_WORK + 32) rw_agc_role(P, m, n, agc_on, ast, tid == RF_WORK);
    else rw_lin_role(P, m, n, c, tid - RF_WORK - 32, agc_on, ast);
}

// ---- the thin form of the per-channel kernel, for more channels than the GPU holds at two CTAs of 192 threads per SM.
// The sequential lanes set the pace (66 cycles per sample); ONE worker warp applies the gain law 17 times faster than that,
// so 128 workers only hold registers.  Here a CTA is three warps (worker, AGC, LIN) and walks the stream in sub-blocks of
// `ns` <= 256 samples (an eighth of the shared memory), so four CTAs fit an SM where two did.  What belongs to the DSP block
// rather than to the stream keeps its place: the peak meters are raised to the block maximum behind the LAST sub-block of
// a block (meter.c:95), the AGC meter's block maximum is carried across the sub-blocks, the siphon takes the block's tail.
// Same arithmetic per sample in the same order as rxa_wide_seq_kernel: identical outputs, states and meter readings.
static constexpr int RT_W = 32, RT_THREADS = 96;
__device__ __forceinline__ void rt_bar_all() { asm volatile("bar.sync 0, %0;" :: "r"(RT_THREADS) : "memory"); }

__device__ __forceinline__ void rt_load_block(const RxaWideParams &P, const RwSmem &m, int ns, int sub, int c, int t, int lane, bool agc_on)
{
    const size_t off = (size_t)c * P.nblocks * P.n + (size_t)t * ns;
    const int par = t & 1;
    for (int i0 = 0; i0 < ns; i0 += 8 * RT_W) {
        double v[4][8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = i0 + lane + k * RT_W;
            if (i < ns) {
                v[0][k] = P.sms[off + i]; v[1][k] = P.smadc[off + i];
                if (agc_on) { v[2][k] = P.rm[off + i]; v[3][k] = P.absd[off + i]; }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = i0 + lane + k * RT_W;
            if (i < ns) {
                m.SMS[par][i] = v[0][k]; m.SMADC[par][i] = v[1][k];
                if (agc_on) { m.RV[par][i] = v[2][k]; m.ABSD[par][i] = v[3][k]; }
            }
        }
    }
    if (lane < 2) m.SC[8 + 3 * par + lane] = P.np[((size_t)c * P.nblocks + t / sub) * 2 + lane];
}

// gain law, panel, output, siphon for sub-block tb; run_max carries the block maximum of |out|^2 across a block's sub-blocks
__device__ __forceinline__ void rt_gain_block(const RxaWideParams &P, const RwSmem &m, int ns, int sub, int c, int tb, int lane, bool agc_on, int ab,
                                              const double *hs, double &run_max)
{
    const int par = tb & 1, n = P.n;
    const double *RV = m.RV[par];
    double *SMAGC = m.SMAGC[par];
    const cd *ys = P.y + (size_t)c * P.y_stride;
    cd *y = P.out + (size_t)c * P.out_stride + (size_t)tb * ns;
    const int boff = (tb % sub) * ns;                           // the sub-block's place inside its DSP block
    double mx_agc = 0.0;
    for (int i0 = 0; i0 < ns; i0 += 8 * RT_W) {
        cd dl[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = i0 + lane + k * RT_W;
            if (i < ns) {
                const long idx = (long)tb * ns + i - ab;         // the sample leaving the AGC's delay line
                dl[k] = idx >= 0 ? ys[idx] : make_double2(hs[(idx + ab) * 3], hs[(idx + ab) * 3 + 1]);
            }
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = i0 + lane + k * RT_W;
            if (i >= ns) continue;
            cd o;
            const cd d = dl[k];
            if (agc_on) {
                const double volts = RV[i];
                const double lg = log10(__dmul_rn(P.a.inv_max_input, volts));
                const double mult = __ddiv_rn(__dsub_rn(P.a.out_target, __dmul_rn(P.a.slope_constant, 0.0 < lg ? 0.0 : lg)), volts);
                o = make_double2(__dmul_rn(d.x, mult), __dmul_rn(d.y, mult));
            } else {
                o = P.agc_run ? make_double2(__dmul_rn(P.a.fixed_gain, d.x), __dmul_rn(P.a.fixed_gain, d.y)) : d;
            }
            const double sm = rf_smag(o);
            SMAGC[i] = sm;
            mx_agc = sm > mx_agc ? sm : mx_agc;
            if (P.sip) {
                const int a = boff + i;
                if (n >= P.sipsize) { if (a >= n - P.sipsize) P.sip[(size_t)c * P.sipsize + (a - (n - P.sipsize))] = o; }
                else P.sip[(size_t)c * P.sipsize + ((P.sip_idx + tb * ns + i) & (P.sipsize - 1))] = o;
            }
            y[i] = make_double2(__dmul_rn(P.gI, o.x), __dmul_rn(P.gQ, o.y));
        }
    }
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, mx_agc, o); mx_agc = t > mx_agc ? t : mx_agc; }
    if (boff == 0) run_max = 0.0;
    run_max = mx_agc > run_max ? mx_agc : run_max;
    if (lane == 0) m.SC[8 + 3 * par + 2] = run_max;
}

__device__ __forceinline__ void rt_worker_role(const RxaWideParams &P, const RwSmem &m, int ns, int sub, int c, int lane, bool agc_on, int ab, double *hs)
{
    const int nb = P.nblocks * sub;
    double run_max = 0.0;
    rt_load_block(P, m, ns, sub, c, 0, lane, agc_on);
    rt_bar_all();
    for (int t = 0; t < nb; t++) {
        rt_bar_all();
        if (t >= 1) rt_gain_block(P, m, ns, sub, c, t - 1, lane, agc_on, ab, hs, run_max);
        __syncwarp();
        if (t + 1 < nb) rt_load_block(P, m, ns, sub, c, t + 1, lane, agc_on);
        rt_bar_all();
    }
    rt_bar_all();
    rt_gain_block(P, m, ns, sub, c, nb - 1, lane, agc_on, ab, hs, run_max);
    __syncwarp();
    if (agc_on) {
        const long N = (long)P.nblocks * P.n;
        const cd *ys = P.y + (size_t)c * P.y_stride;
        if (N < ab) {
            const int keep = ab - (int)N;
            for (int j0 = 0; j0 < keep; j0 += RT_W) {
                const int j = j0 + lane;
                double r0 = 0, r1 = 0, r2 = 0;
                if (j < keep) { r0 = hs[(N + j) * 3]; r1 = hs[(N + j) * 3 + 1]; r2 = hs[(N + j) * 3 + 2]; }
                __syncwarp();
                if (j < keep) { hs[j * 3] = r0; hs[j * 3 + 1] = r1; hs[j * 3 + 2] = r2; }
                __syncwarp();
            }
        }
        for (int k = lane + (N < ab ? ab - (int)N : 0); k < ab; k += RT_W) {
            const cd v = ys[N - ab + k];
            hs[k * 3] = v.x; hs[k * 3 + 1] = v.y; hs[k * 3 + 2] = rf_mag(v, P.a.pmode);
        }
    }
    rt_bar_all();
}

__device__ __forceinline__ void rt_agc_role(const RxaWideParams &P, const RwSmem &m, int ns, int sub, bool agc_on, double *ast, bool lane0)
{
    double volts = ast[3], save_volts = ast[4];
    int hang_counter = (int)ast[7], decay_type = (int)ast[8], state_ = (int)ast[9];
    const int nb = P.nblocks * sub;
    rt_bar_all();
    for (int t = 0; t < nb; t++) {
        rt_bar_all();
        if (agc_on) rf_agc_block(m.RV[t & 1], m.FBA, m.HBA, m.linpos, ns, P.a, volts, save_volts, hang_counter, decay_type, state_, nullptr);
        if (lane0 && agc_on && t == nb - 1) {
            ast[3] = volts; ast[4] = save_volts;
            ast[7] = hang_counter; ast[8] = decay_type; ast[9] = state_; ast[10] = __dmul_rn(volts, P.a.inv_out_target);
        }
        rt_bar_all();
    }
    rt_bar_all();
    rt_bar_all();
}

__device__ __forceinline__ void rt_lin_role(const RxaWideParams &P, const RwSmem &m, int ns, int sub, int c, int lane, bool agc_on, double *ast)
{
    const int mt = lane < 6 ? lane >> 1 : 3, pk = lane & 1, nb = P.nblocks * sub;
    const bool meter = lane < 6, back = agc_on && (lane == 6 || lane == 7), live = meter || back;
    double s = meter ? P.mst[mt][(size_t)c * 2 + pk] : (back ? ast[lane - 1] : 0.0);
    double c1 = 0.0, c2 = 0.0;
    if (meter) { c1 = pk ? 0.0 : 1.0 - P.m_ma; c2 = pk ? P.m_mp : P.m_ma; }
    else if (lane == 6) { c1 = P.a.fast_backmult; c2 = P.a.onemfast_backmult; }
    else if (lane == 7) { c1 = P.a.hang_backmult; c2 = P.a.onemhang_backmult; }
    double *dst = lane == 6 ? m.FBA : (lane == 7 ? m.HBA : m.SC + 16 + lane);
    const int dstep = lane == 6 || lane == 7 ? 1 : 0;
    if (lane == 0) *m.linpos = 0;
    rt_bar_all();
    for (int t = 0; t < nb; t++) {
        rt_bar_all();
        const int par = t & 1;
        const double *src = mt == 0 ? m.SMADC[par] : (mt == 1 ? m.SMS[par] : (mt == 2 ? m.SMAGC[par] : (back ? m.ABSD[par] : m.SMS[par])));
        const bool run = mt != 2 || t >= 2;
        const double r = rf_lin_block(run ? src : m.SMS[par], ns, run ? c1 : 0.0, run ? c2 : 1.0, s, dst, dstep, m.linpos);
        const int u = mt == 2 ? t - 2 : t;                              // the sub-block this lane has just walked
        if (live) { s = r; if (meter && pk && run && (u + 1) % sub == 0) { const double np = m.SC[8 + 3 * par + mt]; if (np > s) s = np; } }
        rt_bar_all();
        if (lane == 0) *m.linpos = 0;
    }
    rt_bar_all();
    if (nb >= 2) {
        const int par = nb & 1, u = nb - 2;
        const double r = rf_lin_block(m.SMAGC[par], ns, mt == 2 ? c1 : 0.0, mt == 2 ? c2 : 1.0, s, m.SC + 16 + lane, 0, nullptr);
        if (mt == 2) { s = r; if (pk && (u + 1) % sub == 0) { const double np = m.SC[8 + 3 * par + 2]; if (np > s) s = np; } }
    }
    rt_bar_all();
    {
        const int par = (nb - 1) & 1;
        const double r = rf_lin_block(m.SMAGC[par], ns, mt == 2 ? c1 : 0.0, mt == 2 ? c2 : 1.0, s, m.SC + 16 + lane, 0, nullptr);
        if (mt == 2) { s = r; if (pk) { const double np = m.SC[8 + 3 * par + 2]; if (np > s) s = np; } }
    }
    if (meter) {
        P.mst[mt][(size_t)c * 2 + pk] = s;
        P.mres[mt][(size_t)c * 3 + pk] = 10.0 * mlog10_dev(P.mtable, s + 1.0e-40);
        if (lane == 4) P.mres[2][(size_t)c * 3 + 2] = 20.0 * mlog10_dev(P.mtable, ast[10] + 1.0e-40);
        else if (!pk) P.mres[mt][(size_t)c * 3 + 2] = 0.0;
    } else if (back) {
        ast[lane - 1] = s;
    }
}

template <int MINB>
__global__ void __launch_bounds__(RT_THREADS, MINB) rxa_thin_seq_kernel(RxaWideParams P, int ns, int sub)
{
    extern __shared__ double smem_raw[];
    const int c = blockIdx.x, tid = threadIdx.x;
    const bool agc_on = P.agc_run && P.a.mode != 0;
    const int ab = agc_on ? P.a.attack_buffsize : 0;
    const RwSmem m = rw_map(smem_raw, ns);
    double *hs = P.agc_hist + (size_t)c * (agc_on ? ab : 1) * 3;
    double *ast = P.agc_state + (size_t)c * 16;
    if (tid < RT_W) rt_worker_role(P, m, ns, sub, c, tid, agc_on, ab, hs);
    else if (tid < RT_W + 32) rt_agc_role(P, m, ns, sub, agc_on, ast, tid == RT_W);
    else rt_lin_role(P, m, ns, sub, c, tid - RT_W - 32, agc_on, ast);
}

size_t rxa_fused_smem(int n, int ab)
{
    const int n2 = 2 * n, tot = ab + n;
    return ((size_t)fft_tw_entries(n2) + n2) * sizeof(cd) + ((size_t)tot + n + ((tot + 31) / 32 + 1) + n + 48 + 2 * n + 2) * sizeof(double);
}

// The configurations this kernel covers: no shifter, no resamplers, side-band modes (neither demodulator running), dsp_size
// a power of two <= 1024, nc / dsp_size partitions; anything else runs the per-stage kernels.
bool Rxa::fusable() const
{
    if (!fused_ok || (shift_run && shift_nonzero) || rsmpin || rsmpout || amd_run || fmd_run || emnr_run || snba_run) return false;
    if (dsp_size > 1024 || dsp_size < 8 || (dsp_size & (dsp_size - 1))) return false;
    const int ab = agc_run && agc->agc.mode != 0 ? agc->agc.attack_buffsize : 0;
    if (agc_run && agc->agc.mode == 5) return false;
    return rxa_fused_smem(dsp_size, ab) <= 200 * 1024;
}

int Rxa::xrxa_fused(const void *din, long is, void *dout, long os, int nblocks, cudaStream_t s)
{
    RxaFusedParams P;
    memset(&P, 0, sizeof(P));
    P.in = (const cd *)din; P.in_stride = is; P.out = (cd *)dout; P.out_stride = os;
    P.n = dsp_size; P.nblocks = nblocks; P.C = C;
    FirCore *firs[2] = {nbp_run ? nbp0 : nullptr, bp1_run ? bp1 : nullptr};
    const bool wide = nblocks >= 2 && (firs[0] || firs[1]) && !getenv("QUISK_RXA_NARROW");
    if (wide) {
        // filters first, all blocks at once, into the filtered-stream scratch; the sequential kernel then starts from there
        const size_t need_spec = (size_t)C * nblocks * 2 * dsp_size * 2 /* spectra + products */, need_y = (size_t)C * nblocks * dsp_size;
        if (need_spec > wide_spec_cap) { if (d_wide_spec) cudaFree(d_wide_spec); d_wide_spec = nullptr; wide_spec_cap = 0;
                                         QC_CUDA(cudaMalloc((void **)&d_wide_spec, need_spec * sizeof(cd))); wide_spec_cap = need_spec; }
        if (need_y > wide_y_cap) { if (d_wide_y) cudaFree(d_wide_y); d_wide_y = nullptr; wide_y_cap = 0;
                                   QC_CUDA(cudaMalloc((void **)&d_wide_y, need_y * sizeof(cd))); wide_y_cap = need_y; }
        const cd *cur = (const cd *)din; long cs = is;
        const long ys = (long)nblocks * dsp_size;
        for (FirCore *f : firs) {
            if (!f) continue;
            int rc = fircore_wide(f, cur, cs, d_wide_y, ys, nblocks, d_wide_spec, s); if (rc != QC_OK) return rc;
            cur = d_wide_y; cs = ys;                       // a second filter runs in place on the scratch (each kernel pair reads before it writes per block: see below)
        }
        P.in_adc = (const cd *)din; P.adc_stride = is;
        P.in = d_wide_y; P.in_stride = ys;
        if (!getenv("QUISK_RXA_WIDE_SERIAL")) {
            // pipelined form: pre kernel over (channel, block), then the two-stage per-channel kernel
            RxaWideParams W;
            memset(&W, 0, sizeof(W));
            W.y = d_wide_y; W.y_stride = ys; W.xadc = (const cd *)din; W.adc_stride = is; W.out = (cd *)dout; W.out_stride = os;
            W.n = dsp_size; W.nblocks = nblocks; W.C = C;
            const size_t per = (size_t)C * nblocks * dsp_size, need = 4 * per + (size_t)C * nblocks * 2;
            if (need > wide_seq_cap) { if (d_wide_seq) cudaFree(d_wide_seq); d_wide_seq = nullptr; wide_seq_cap = 0;
                                       QC_CUDA(cudaMalloc((void **)&d_wide_seq, need * sizeof(double))); wide_seq_cap = need; }
            W.rm = d_wide_seq; W.sms = W.rm + per; W.smadc = W.sms + per; W.absd = W.smadc + per; W.np = W.absd + per;
            SeqStage *mt[3] = {adcmeter, smeter, agcmeter};
            for (int m = 0; m < 3; m++) { W.mst[m] = mt[m]->d_state; W.mres[m] = mt[m]->d_meter; }
            W.m_ma = adcmeter->par[0]; W.m_mp = adcmeter->par[1];
            W.mtable = mlog10_table();
            if (!W.mtable) { set_error("rxa: table allocation failed"); return QC_ENOMEM; }
            W.agc_run = agc_run; W.agc_state = agc->d_state; W.agc_hist = agc->d_ring; W.a = agc->agc;
            W.gI = panel_gain1 * panel_gain2I; W.gQ = panel_gain1 * panel_gain2Q;
            W.sip = sip_run ? d_sip : nullptr; W.sipsize = sipsize; W.sip_idx = sip_idx;
            const int ab = agc_run && agc->agc.mode != 0 ? agc->agc.attack_buffsize : 0;
            const int tot = ab + dsp_size;
            const size_t sh_pre = ((size_t)tot + (tot + 31) / 32 + 1 + 16) * sizeof(double);
            if (sh_pre > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(rxa_wide_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_pre));
            rxa_wide_pre_kernel<<<dim3(C, nblocks), RF_WORK, sh_pre, s>>>(W);
            count_launch();
            QC_CUDA_LAUNCH();
            const size_t sh = rxa_wide_seq_smem(dsp_size);
            int dev = 0, n_sm = 148;
            cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
            // one CTA per SM while the channels fit (174 registers, every worker warp busy), two while THAT fits in one wave, else the
            // thin form at four per SM (QUISK_RXA_MINB = 1 / 2 / 4 forces one of them: tests, measurements)
            const char *force = getenv("QUISK_RXA_MINB");
            const int form = force ? atoi(force) : (C <= n_sm ? 1 : (C <= 2 * n_sm ? 2 : 4));
            const bool fat = form == 1;
            if (fat) {
                if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(rxa_wide_seq_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
                rxa_wide_seq_kernel<1><<<C, RF_THREADS, sh, s>>>(W);
            } else if (form != 2) {
                // more channels than two fat CTAs per SM hold: three-warp CTAs over sub-blocks, four per SM
                const char *fs = getenv("QUISK_RXA_THIN_NS");
                int ns = dsp_size > 256 ? 256 : dsp_size;
                if (fs && atoi(fs) >= 8 && atoi(fs) <= dsp_size && dsp_size % atoi(fs) == 0) ns = atoi(fs);
                const size_t sht = rxa_wide_seq_smem(ns);
                if (sht > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(rxa_thin_seq_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sht));
                rxa_thin_seq_kernel<4><<<C, RT_THREADS, sht, s>>>(W, ns, dsp_size / ns);
            } else {
                if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(rxa_wide_seq_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
                rxa_wide_seq_kernel<2><<<C, RF_THREADS, sh, s>>>(W);
            }
            count_launch();
            QC_CUDA_LAUNCH();
            if (sip_run && dsp_size < sipsize) sip_idx = (int)(((long)sip_idx + (long)nblocks * dsp_size) & (sipsize - 1));
            return QC_OK;
        }
    }
    for (FirCore *f : firs) {
        if (!f || wide) continue;
        const int k = P.n_fir++;
        P.prev[k] = f->d_prev; P.fdl[k] = f->d_fdl; P.mask[k] = f->d_mask[f->cset]; P.nfor[k] = f->nfor; P.buffidx[k] = f->buffidx;
        P.tw = f->tw;
    }
    if (!P.tw) P.tw = fft_twiddles(2 * dsp_size);
    SeqStage *mt[3] = {adcmeter, smeter, agcmeter};
    for (int m = 0; m < 3; m++) { P.mst[m] = mt[m]->d_state; P.mres[m] = mt[m]->d_meter; }
    P.m_ma = adcmeter->par[0]; P.m_mp = adcmeter->par[1];
    P.mtable = mlog10_table();
    if (!P.mtable || !P.tw) { set_error("rxa: table allocation failed"); return QC_ENOMEM; }
    P.agc_run = agc_run; P.agc_state = agc->d_state; P.agc_hist = agc->d_ring; P.a = agc->agc;
    P.gI = panel_gain1 * panel_gain2I; P.gQ = panel_gain1 * panel_gain2Q;
    P.sip = sip_run ? d_sip : nullptr; P.sipsize = sipsize; P.sip_idx = sip_idx;
    static long long *d_dbg = nullptr;
    if (getenv("QUISK_RXA_DEBUG")) {
        if (!d_dbg) { cudaMalloc((void **)&d_dbg, 32 * sizeof(long long)); cudaMemset(d_dbg, 0, 32 * sizeof(long long)); }
        else {
            long long h[32]; cudaDeviceSynchronize(); cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "rxa fused fircore: load %lld fft %lld mac %lld ifft %lld | agc block: full %lld failed %lld single %lld cycles %lld\n", h[17] - h[16], h[18] - h[17], h[19] - h[18], h[20] - h[19],
                    h[24], h[25], h[26], h[27]);
            fprintf(stderr, "rxa fused stamps rel. to block start: fir_end %lld mags_end %lld w_afterA %lld agc_start %lld agc_end %lld lin_end %lld lin_afterB %lld w_afterB %lld w_end %lld\n",
                    h[1] - h[0], h[2] - h[0], h[12] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[13] - h[0], h[6] - h[0], h[7] - h[0]);
            fprintf(stderr, "rxa fused phases (cycles): fir %lld  mags/ring_max %lld  agc lane %lld  lin lane %lld  seq phase %lld  gain/out %lld\n",
                    h[1] - h[0], h[2] - h[1], h[4] - h[3], h[5] - h[3], h[6] - h[2], h[7] - h[6]);
        }
        P.dbg = d_dbg;
    }
    const int ab = agc_run && agc->agc.mode != 0 ? agc->agc.attack_buffsize : 0;
    const size_t sh = rxa_fused_smem(dsp_size, ab);
    // few channels: one CTA per SM anyway, so the kernel may have every register (no spills in the transforms); many channels:
    // two CTAs per SM matter more than the spills
    int dev = 0, n_sm = 148;
    cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const char *force = getenv("QUISK_RXA_MINB");
    const bool fat = force ? atoi(force) == 1 : C <= n_sm;
    if (fat) {
        if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(rxa_ssb_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        rxa_ssb_fused_kernel<1><<<C, RF_THREADS, sh, s>>>(P);
    } else {
        if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(rxa_ssb_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        rxa_ssb_fused_kernel<2><<<C, RF_THREADS, sh, s>>>(P);
    }
    count_launch();
    QC_CUDA_LAUNCH();
    if (!wide) for (FirCore *f : firs) if (f) f->buffidx = (f->buffidx + nblocks) & (f->nfor - 1);
    if (sip_run && dsp_size < sipsize) sip_idx = (int)(((long)sip_idx + (long)nblocks * dsp_size) & (sipsize - 1));
    return QC_OK;
}

}  // namespace qc
