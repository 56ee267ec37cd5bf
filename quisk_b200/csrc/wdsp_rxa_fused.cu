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
#include "fft_device.cuh"
#include "wdsp_internal.h"

namespace qc {

static constexpr int RF_WORK = 128;         // worker threads (the transform's lanes)
static constexpr int RF_THREADS = 192;      // + AGC warp + LIN warp

struct RxaFusedParams {
    const cd *in; long in_stride; cd *out; long out_stride;
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
    double *SMS, *SMADC, *A, *RV, *BM, *SMAGC, *SC;
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
    m.SC = m.SMAGC + n;                                     // [16] 0..3 warp maxima, 4 np_adc, 5 np_s, 6 np_agc
    return m;
}

// One sample through the reference's general machine (wcpAGC.c:195-333), same tests in the same order; expects
// abs_out_sample and ring_max, advances i.
#define RF_AGC_GENERAL_STEP                                                                                             \
    fb = __dadd_rn(__dmul_rn(k_fbm, abs_out_sample), __dmul_rn(k_ofbm, fb));                                            \
    hb = __dadd_rn(__dmul_rn(k_hbm, abs_out_sample), __dmul_rn(k_ohbm, hb));                                            \
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
    if (store) RV[i] = volts;                                                                                           \
    i++;

// ---- role 1: the AGC warp.  All 32 lanes run the same instructions on the same data (no divergence inside the warp, so the
// CTA-wide barriers are reached by whole warps); lane 0 alone stores.
// One block of the volts machine: A[i] = |sample leaving the delay line|, RV[i] = ring_max on the way in, volts on the way out.
__device__ __forceinline__ void rf_agc_block(const double *A, double *RV, int n, const AgcParams &a, bool store,
                                             double &volts, double &save_volts, double &fb, double &hb, int &hang_counter, int &decay_type, int &state_)
{
    const double k_fbm = a.fast_backmult, k_ofbm = a.onemfast_backmult, k_hbm = a.hang_backmult, k_ohbm = a.onemhang_backmult,
                 k_attack = a.attack_mult, k_decay = a.decay_mult, k_hdecay = a.hang_decay_mult, k_fdecay = a.fast_decay_mult,
                 k_pop = a.pop_ratio, k_hlevel = a.hang_level, k_minv = a.min_volts;
    int i = 0;
    // the next chunk's inputs are fetched while the current chain runs: their shared-memory latency never sits in front of it
    double rmn[8], abn[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { rmn[j] = j < n ? RV[j] : 0.0; abn[j] = j < n ? A[j] : 0.0; }
    while (i < n) {
        if (n - i >= 8 && (state_ == 0 || state_ >= 3)) {
            // a run of eight samples on the assumption that the state does not change
            const double M = state_ == 0 ? k_attack : (state_ == 3 ? k_decay : k_hdecay);
            const bool want = state_ == 0;
            double rm[8], ab[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { rm[j] = rmn[j]; ab[j] = abn[j]; }
            const int nx = i + 8;
#pragma unroll
            for (int j = 0; j < 8; j++) { rmn[j] = nx + j < n ? RV[nx + j] : 0.0; abn[j] = nx + j < n ? A[nx + j] : 0.0; }
            double v = volts, f = fb, h = hb;
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const bool p = rm[j] >= v;
                const double d = __dsub_rn(rm[j], v);
                v = __dadd_rn(v, __dmul_rn(d, M));
                f = __dadd_rn(__dmul_rn(k_fbm, ab[j]), __dmul_rn(k_ofbm, f));
                h = __dadd_rn(__dmul_rn(k_hbm, ab[j]), __dmul_rn(k_ohbm, h));
                ok = ok && p == want && !(v < k_minv);
                if (store) RV[i + j] = v;                   // speculative: put back below if the run did not hold
            }
            if (ok) {
                volts = v; fb = f; hb = h;
                hang_counter = hang_counter > 8 ? hang_counter - 8 : 0;
                i += 8;
                continue;
            }
            // the run did not hold (a handful of times per block): ring_max goes back into RV and these eight samples go through
            // the general machine one by one
            if (store) {
#pragma unroll
                for (int j = 0; j < 8; j++) RV[i + j] = rm[j];
            }
            for (int g = 0; g < 8; g++) {
                const double abs_out_sample = A[i], ring_max = g == 0 ? rm[0] : RV[i];
                RF_AGC_GENERAL_STEP
            }
#pragma unroll
            for (int j = 0; j < 8; j++) { rmn[j] = i + j < n ? RV[i + j] : 0.0; abn[j] = i + j < n ? A[i + j] : 0.0; }
            continue;
        }
        // one sample through the general machine (states 1 and 2, and the last few samples of a block)
        {
            const double abs_out_sample = A[i], ring_max = RV[i];
            RF_AGC_GENERAL_STEP
#pragma unroll
            for (int j = 0; j < 8; j++) { rmn[j] = i + j < n ? RV[i + j] : 0.0; abn[j] = i + j < n ? A[i + j] : 0.0; }
        }
    }
}
#undef RF_AGC_GENERAL_STEP

__device__ __forceinline__ void rf_agc_role(const RxaFusedParams &P, const RfSmem &m, int n, bool agc_on, double *ast, bool lane0)
{
    double volts = ast[3], save_volts = ast[4], fb = ast[5], hb = ast[6];
    int hang_counter = (int)ast[7], decay_type = (int)ast[8], state_ = (int)ast[9];
    rf_bar_all();                                           // set-up done
    for (int b = 0; b < P.nblocks; b++) {
        rf_bar_all();                                       // the workers have A and ring_max ready
        long long t0 = 0;
        const bool stamp = P.dbg && blockIdx.x == 0 && b == P.nblocks - 1 && lane0;
        if (stamp) t0 = clock64();
        if (agc_on) rf_agc_block(m.A, m.RV, n, P.a, lane0, volts, save_volts, fb, hb, hang_counter, decay_type, state_);
        if (stamp) { P.dbg[3] = t0; P.dbg[4] = clock64(); }
        if (lane0 && agc_on && b == P.nblocks - 1) {
            ast[3] = volts; ast[4] = save_volts; ast[5] = fb; ast[6] = hb;
            ast[7] = hang_counter; ast[8] = decay_type; ast[9] = state_; ast[10] = __dmul_rn(volts, P.a.inv_out_target);
        }
        rf_bar_all();                                       // volts ready for the gain law
        rf_bar_all();                                       // end of block
    }
}

// ---- role 2: the LIN warp.  Lane l < 6: adc avg, adc peak, s avg, s peak, agc avg, agc peak; s = c1 * x[i] + c2 * s
// (averages: c1 = 1 - mult, c2 = mult; peak decays: c1 = 0), one instruction stream for all lanes, inputs fetched one
// group ahead of the dependent multiply-add chain.
__device__ __forceinline__ double rf_lin_block(const double *x, int n, double c1, double c2, double s)
{
    double xa[4], xb[4];
#pragma unroll
    for (int j = 0; j < 4; j++) xa[j] = j < n ? x[j] : 0.0;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) xb[j] = i + 4 + j < n ? x[i + 4 + j] : 0.0;
        const double t0 = __dmul_rn(c1, xa[0]), t1 = __dmul_rn(c1, xa[1]), t2 = __dmul_rn(c1, xa[2]), t3 = __dmul_rn(c1, xa[3]);
        s = __dadd_rn(__dmul_rn(c2, s), t0);
        s = __dadd_rn(__dmul_rn(c2, s), t1);
        s = __dadd_rn(__dmul_rn(c2, s), t2);
        s = __dadd_rn(__dmul_rn(c2, s), t3);
#pragma unroll
        for (int j = 0; j < 4; j++) xa[j] = xb[j];
    }
    for (; i < n; i++) s = __dadd_rn(__dmul_rn(c2, s), __dmul_rn(c1, x[i]));
    return s;
}

__device__ __forceinline__ void rf_lin_role(const RxaFusedParams &P, const RfSmem &m, int n, int c, int lane, double *ast)
{
    const int mt = lane < 6 ? lane >> 1 : 0, pk = lane & 1;
    const bool live = lane < 6;
    double s = live ? P.mst[mt][(size_t)c * 2 + pk] : 0.0;
    const double c1 = !live || pk ? 0.0 : 1.0 - P.m_ma;     // (1.0 - mult_average), meter.c:90
    const double c2 = !live ? 0.0 : (pk ? P.m_mp : P.m_ma);
    const double *src = mt == 0 ? m.SMADC : (mt == 1 ? m.SMS : m.SMAGC);
    rf_bar_all();
    for (int b = 0; b < P.nblocks; b++) {
        rf_bar_all();
        if (mt < 2 || b > 0) {                              // the AGC meter runs one block late
            const double r = rf_lin_block(src, n, c1, c2, s);
            if (live) { s = r; if (pk) { const double np = m.SC[4 + mt]; if (np > s) s = np; } }     // meter.c:95
        }
        if (P.dbg && blockIdx.x == 0 && b == P.nblocks - 1 && lane == 0) P.dbg[5] = clock64();
        rf_bar_all();
        if (P.dbg && blockIdx.x == 0 && b == P.nblocks - 1 && lane == 0) P.dbg[13] = clock64();
        rf_bar_all();
    }
    // the AGC meter's last block, then the meters' states and readings
    if (mt == 2 && P.nblocks > 0) {
        const double r = rf_lin_block(m.SMAGC, n, c1, c2, s);
        if (live) { s = r; if (pk) { const double np = m.SC[6]; if (np > s) s = np; } }
    }
    if (live) {
        P.mst[mt][(size_t)c * 2 + pk] = s;
        P.mres[mt][(size_t)c * 3 + pk] = 10.0 * mlog10_dev(P.mtable, s + 1.0e-40);
        if (lane == 4) P.mres[2][(size_t)c * 3 + 2] = 20.0 * mlog10_dev(P.mtable, ast[10] + 1.0e-40);      // meter.c:99: *pgain = the AGC's gain
        else if (!pk) P.mres[mt][(size_t)c * 3 + 2] = 0.0;                                                  // the other two have no gain reading
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
                                        const cd *__restrict__ x)
{
    const int n2 = 2 * n;
    if (x) {
        for (int i = tid; i < n; i += RF_WORK) { const cd v = x[i]; S[fsw(i)] = pv[i]; S[fsw(n + i)] = v; pv[i] = v; }
    } else {                                                // second fircore of the chain: its input is the first one's output
        cd t0[8];                                           // n <= 1024: at most 8 samples per worker
#pragma unroll
        for (int k = 0; k < 8; k++) { const int i = tid + k * RF_WORK; if (i < n) t0[k] = S[fsw(i)]; }
        rf_bar_work();
#pragma unroll
        for (int k = 0; k < 8; k++) { const int i = tid + k * RF_WORK; if (i < n) { S[fsw(i)] = pv[i]; S[fsw(n + i)] = t0[k]; pv[i] = t0[k]; } }
    }
    rf_bar_work();
    fft_smem<1, RF_WORK>(S, n2, twl, -1, tid, RF_WORK);
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
        for (int j = 1; j < nfor; j++) {
            k = (k + mask) & mask;
            const cd *fk = fd + (size_t)k * n2 + i0, *mk = fmask + (size_t)j * n2 + i0;
            cd Y[4], mm[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { if (i0 + u * RF_WORK < n2) { Y[u] = fk[u * RF_WORK]; mm[u] = mk[u * RF_WORK]; } }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (i0 + u * RF_WORK < n2) {
                    acc[u].x += Y[u].x * mm[u].x - Y[u].y * mm[u].y;
                    acc[u].y += Y[u].x * mm[u].y + Y[u].y * mm[u].x;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i = i0 + u * RF_WORK; if (i < n2) { fd[(size_t)bi * n2 + i] = X[u]; S[fsw(i)] = acc[u]; } }
    }
    rf_bar_work();
    fft_smem<1, RF_WORK>(S, n2, twl, +1, tid, RF_WORK);
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
                                    (P.buffidx[0] + b) & (P.nfor[0] - 1), n, tid, x);
        if (P.n_fir > 1) rf_fircore(S, m.twl, P.prev[1] + (size_t)c * n, P.fdl[1] + (size_t)c * P.nfor[1] * 2 * n, P.mask[1], P.nfor[1],
                                    (P.buffidx[1] + b) & (P.nfor[1] - 1), n, tid, nullptr);
        if (P.n_fir == 0) {
            for (int i = tid; i < n; i += RF_WORK) S[fsw(i)] = x[i];
            rf_bar_work();
        }
        if (stamp) P.dbg[1] = clock64();
        // ---- meter inputs, magnitudes
        double mx_adc = 0.0, mx_s = 0.0;
        for (int i = tid; i < n; i += RF_WORK) {
            const cd v = S[fsw(i)];
            const double sm = rf_smag(v), sa = rf_smag(x[i]);
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
        for (int i = tid; i < n; i += RF_WORK) {
            cd o;
            if (agc_on) {
                const double volts = RV[i];
                const double lg = log10(__dmul_rn(P.a.inv_max_input, volts));
                const double mult = __ddiv_rn(__dsub_rn(P.a.out_target, __dmul_rn(P.a.slope_constant, 0.0 < lg ? 0.0 : lg)), volts);
                cd d;
                if (i < ab) d = make_double2(hs[i * 3], hs[i * 3 + 1]);
                else d = S[fsw(i - ab)];
                o = make_double2(__dmul_rn(d.x, mult), __dmul_rn(d.y, mult));
            } else {
                const cd d = S[fsw(i)];
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

__global__ void __launch_bounds__(RF_THREADS, 2) rxa_ssb_fused_kernel(RxaFusedParams P)
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
    else rf_lin_role(P, m, n, c, tid - RF_WORK - 32, ast);
}

size_t rxa_fused_smem(int n, int ab)
{
    const int n2 = 2 * n, tot = ab + n;
    return ((size_t)fft_tw_entries(n2) + n2) * sizeof(cd) + ((size_t)tot + n + ((tot + 31) / 32 + 1) + n + 16) * sizeof(double);
}

// The configurations this kernel covers: no shifter, no resamplers, side-band modes (neither demodulator running), dsp_size
// a power of two <= 1024, nc / dsp_size partitions; anything else runs the per-stage kernels.
bool Rxa::fusable() const
{
    if (!fused_ok || (shift_run && shift_nonzero) || rsmpin || rsmpout || amd_run || fmd_run) return false;
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
    for (FirCore *f : firs) {
        if (!f) continue;
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
        if (!d_dbg) { cudaMalloc((void **)&d_dbg, 16 * sizeof(long long)); cudaMemset(d_dbg, 0, 16 * sizeof(long long)); }
        else {
            long long h[16]; cudaDeviceSynchronize(); cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
            fprintf(stderr, "rxa fused stamps rel. to block start: fir_end %lld mags_end %lld w_afterA %lld agc_start %lld agc_end %lld lin_end %lld lin_afterB %lld w_afterB %lld w_end %lld\n",
                    h[1] - h[0], h[2] - h[0], h[12] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[13] - h[0], h[6] - h[0], h[7] - h[0]);
            fprintf(stderr, "rxa fused phases (cycles): fir %lld  mags/ring_max %lld  agc lane %lld  lin lane %lld  seq phase %lld  gain/out %lld\n",
                    h[1] - h[0], h[2] - h[1], h[4] - h[3], h[5] - h[3], h[6] - h[2], h[7] - h[6]);
        }
        P.dbg = d_dbg;
    }
    const int ab = agc_run && agc->agc.mode != 0 ? agc->agc.attack_buffsize : 0;
    const size_t sh = rxa_fused_smem(dsp_size, ab);
    if (sh > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(rxa_ssb_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    rxa_ssb_fused_kernel<<<C, RF_THREADS, sh, s>>>(P);
    count_launch();
    QC_CUDA_LAUNCH();
    for (FirCore *f : firs) if (f) f->buffidx = (f->buffidx + nblocks) & (f->nfor - 1);
    if (sip_run && dsp_size < sipsize) sip_idx = (int)(((long)sip_idx + (long)nblocks * dsp_size) & (sipsize - 1));
    return QC_OK;
}

}  // namespace qc
