// quisk_b200/csrc/wdsp_seq_nofma.cu -- the per-sample recurrent RXA stages: xshift (wdsp/shift.c:60-86),
// xwcpagc (wdsp/wcpAGC.c:161-348), xamd (wdsp/amd.c:115-239), the PLL of xfmd (wdsp/fmd.c:144-170),
// xsnotch (wdsp/iir.c:76-95), xmeter (wdsp/meter.c:75-107) and xpanel (wdsp/patchpanel.c:55-105).
//
// These are scalar state machines / IIR recurrences: there is no parallelism along time that keeps
// the reference's arithmetic, so ONE LANE WALKS ONE CHANNEL'S BLOCK (out of shared memory) and the batch
// supplies the parallelism ("do not time-shard exactly; shard by channel").
// This file is compiled with --fmad=false: a*b+c is two roundings, as in the reference built by gcc for
// baseline x86-64, so the state trajectories (AGC state switches, PLL phase) follow the reference's to
// the last bit except where device libm (cos, sin, atan2, log10) differs from glibc by an ulp.
#include "wdsp_internal.h"
#include <cmath>

namespace qc {

static const double kPI = 3.1415926535897932, kTWOPI = 6.2831853071795864;
#define TWOPI_D 6.2831853071795864

// Execution shape shared by these kernels: ONE CTA PER CHANNEL.  The block is staged in shared memory with
// coalesced 16-byte loads, lane 0 walks the recurrence over shared memory (a dependent FP64 chain of a few
// tens of cycles per sample instead of an L2 round trip per sample), then the CTA stores the block
// coalesced.  SEQ_T threads only do the staging; the batch of channels supplies the parallelism.
static constexpr int SEQ_T = 64;
#define SEQ_STAGE_IN()                                                              \
    extern __shared__ double seq_smem[];                                            \
    cd *sx = reinterpret_cast<cd *>(seq_smem);                                      \
    const int c = blockIdx.x;                                                       \
    {                                                                               \
        const cd *gx = in + (size_t)c * is;                                         \
        for (int i = threadIdx.x; i < n; i += blockDim.x) sx[i] = gx[i];            \
    }                                                                               \
    __syncthreads();
#define SEQ_STAGE_OUT()                                                             \
    __syncthreads();                                                                \
    {                                                                               \
        cd *gy = out + (size_t)c * os;                                              \
        for (int i = threadIdx.x; i < n; i += blockDim.x) gy[i] = sx[i];            \
    }

// ------------------------------------------------------------------------------------------- shift
__global__ void shift_kernel(const cd *in, long is, cd *out, long os, int n, int C, double *state, const double *par)
{
    SEQ_STAGE_IN();
    if (threadIdx.x == 0) {
    double phase = state[c];
    const double delta = par[c * 3], cos_delta = par[c * 3 + 1], sin_delta = par[c * 3 + 2];
    double cos_phase = cos(phase), sin_phase = sin(phase);
    const cd *x = sx;
    cd *y = sx;
    for (int i = 0; i < n; i++) {
        const double I1 = x[i].x, Q1 = x[i].y;
        y[i] = make_double2(I1 * cos_phase - Q1 * sin_phase, I1 * sin_phase + Q1 * cos_phase);
        const double t1 = cos_phase, t2 = sin_phase;
        cos_phase = t1 * cos_delta - t2 * sin_delta;
        sin_phase = t1 * sin_delta + t2 * cos_delta;
        phase += delta;
        if (phase >= TWOPI_D) phase -= TWOPI_D;
        if (phase < 0.0) phase += TWOPI_D;
    }
    state[c] = phase;
    }
    SEQ_STAGE_OUT();
}

// ------------------------------------------------------------------------------------------ wcpagc
// state: 0 (unused) 1 (unused) 2 ring_max 3 volts 4 save_volts 5 fast_backaverage 6 hang_backaverage
//        7 hang_counter 8 decay_type 9 state 10 gain
// hist : [C][ab][3] the last ab = attack_buffsize inputs (re, im, |x|), oldest first -- the live part of the
//        reference's ring (wcpAGC.c:179-190: out_index trails in_index by attack_buffsize).
//
// One CTA per channel, four phases:
//  1. stage [history | block] and the magnitudes in shared memory (coalesced);
//  2. ring_max for every sample IN PARALLEL.  The reference maintains it lazily (rescan the window when the
//     outgoing sample was the maximum, wcpAGC.c:196-210), but its value is exactly the maximum of the last
//     attack_buffsize magnitudes including the new one -- a pure function of the input, no rounding involved;
//  3. lane 0 walks the five-state attack/decay/hang machine (wcpAGC.c:215-333) over shared memory: volts[i];
//  4. the gain law (log10, divide) and the delayed output sample for every i in parallel (wcpAGC.c:335-340).
__global__ void __launch_bounds__(128) wcpagc_kernel(const cd *in, long is, cd *out, long os, int n, int C,
                                                     double *state, double *hist, AgcParams a)
{
    extern __shared__ double seq_smem[];
    const int c = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const cd *x = in + (size_t)c * is;
    cd *y = out + (size_t)c * os;
    if (a.mode == 0) {          // fixed gain, wcpAGC.c:168-176
        for (int i = tid; i < n; i += nt) y[i] = make_double2(a.fixed_gain * x[i].x, a.fixed_gain * x[i].y);
        return;
    }
    const int ab = a.attack_buffsize, tot = ab + n;
    // Shared memory holds the magnitudes of [history | block], ring_max (overwritten by volts, then by the gain) and
    // the block maxima; the samples themselves are only staged when the stage runs in place (out == in), because
    // then every input is overwritten before its delayed use.  Out of place that is ~23 KB for a 1024-sample block
    // and attack_buffsize 768, so a whole batch of 1024 channels is resident at once.
    const bool inplace = (const cd *)out == in;
    double *A = seq_smem;                                   // [tot] magnitudes of [history | block]
    double *RV = A + tot;                                   // [n]   ring_max -> volts -> mult
    double *BM = RV + n;                                    // [(tot + 31) / 32 + 1] block maxima
    cd *XS = reinterpret_cast<cd *>(BM + ((tot + 31) / 32 + 1) + 1);      // [n] the block, in-place runs only
    if ((reinterpret_cast<size_t>(XS) & 15) != 0) XS = reinterpret_cast<cd *>(reinterpret_cast<double *>(XS) + 1);
    double *hs = hist + (size_t)c * ab * 3;
    for (int i = tid; i < ab; i += nt) A[i] = hs[i * 3 + 2];
    for (int i = tid; i < n; i += nt) {
        const cd v = x[i];
        if (inplace) XS[i] = v;
        double m;
        if (a.pmode == 0) { const double f0 = fabs(v.x), f1 = fabs(v.y); m = f0 < f1 ? f1 : f0; }
        else m = sqrt(v.x * v.x + v.y * v.y);
        A[ab + i] = m;
    }
    __syncthreads();
    const int nblk = (tot + 31) >> 5;
    for (int j = tid; j < nblk; j += nt) {
        double m = 0.0;
        const int e = min(tot, (j + 1) << 5);
        for (int k = j << 5; k < e; k++) m = A[k] > m ? A[k] : m;
        BM[j] = m;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        // window = combined indices [i + 1, i + ab]
        const int lo = i + 1, hi = i + ab;
        double m = 0.0;
        const int b0 = (lo + 31) >> 5, b1 = (hi + 1) >> 5;  // whole 32-blocks [b0, b1)
        if (b0 >= b1) {
            for (int k = lo; k <= hi; k++) m = A[k] > m ? A[k] : m;
        } else {
            for (int k = lo; k < (b0 << 5); k++) m = A[k] > m ? A[k] : m;
            for (int j = b0; j < b1; j++) m = BM[j] > m ? BM[j] : m;
            for (int k = b1 << 5; k <= hi; k++) m = A[k] > m ? A[k] : m;
        }
        RV[i] = m;
    }
    __syncthreads();
    double *st = state + (size_t)c * 16;
    if (tid == 0) {
        double volts = st[3], save_volts = st[4], fast_backaverage = st[5], hang_backaverage = st[6];
        int hang_counter = (int)st[7], decay_type = (int)st[8], state_ = (int)st[9];
        // loop constants in registers: left as kernel parameters they are re-read from the constant bank inside the
        // loop, and those loads sit in front of the dependent multiplies
        const double k_fast_backmult = a.fast_backmult, k_onemfast_backmult = a.onemfast_backmult, k_hang_backmult = a.hang_backmult,
                     k_onemhang_backmult = a.onemhang_backmult, k_attack_mult = a.attack_mult, k_decay_mult = a.decay_mult,
                     k_hang_decay_mult = a.hang_decay_mult, k_fast_decay_mult = a.fast_decay_mult, k_pop_ratio = a.pop_ratio,
                     k_hang_level = a.hang_level, k_min_volts = a.min_volts;
        double last_rm = 0.0;
        // the next sample's inputs are fetched one step ahead: the shared-memory latency would otherwise sit in
        // front of every step of this dependent chain
        double na = A[0], nr = RV[0];
        for (int i = 0; i < n; i++) {
            const double abs_out_sample = na;
            const double ring_max = nr;
            if (i + 1 < n) { na = A[i + 1]; nr = RV[i + 1]; }
            last_rm = ring_max;
            fast_backaverage = k_fast_backmult * abs_out_sample + k_onemfast_backmult * fast_backaverage;
            hang_backaverage = k_hang_backmult * abs_out_sample + k_onemhang_backmult * hang_backaverage;
            if (hang_counter > 0) --hang_counter;
            // wcpAGC.c:215-333.  Almost every sample takes one of two paths -- attack (ring_max >= volts, from any
            // state) or steady decay (states 3 / 4) -- so those two are tested first and update volts directly; the
            // transitions out of states 0, 1, 2 (a handful per second of signal) go through the general code.  Same
            // tests in the same order and the same arithmetic as the reference.  (Measured on B200, cycles per
            // sample of this loop: five-case switch 250, single update site 197, this form 178, fully branch-free
            // select form 210.)
            const double d = ring_max - volts;
            if (ring_max >= volts) {
                if (state_ >= 2) save_volts = volts;
                state_ = 0;
                volts += d * k_attack_mult;
            } else if (state_ >= 3) {
                volts += d * (state_ == 3 ? k_decay_mult : k_hang_decay_mult);
            } else if (state_ == 0) {
                if (volts > k_pop_ratio * fast_backaverage) { state_ = 1; volts += d * k_fast_decay_mult; }
                else if (a.hang_enable && hang_backaverage > k_hang_level) { state_ = 2; hang_counter = (int)(a.hangtime * a.sample_rate); decay_type = 1; }
                else { state_ = 3; volts += d * k_decay_mult; decay_type = 0; }
            } else if (state_ == 1) {
                if (volts > save_volts) volts += d * k_fast_decay_mult;
                else if (hang_counter > 0) state_ = 2;
                else if (decay_type == 0) { state_ = 3; volts += d * k_decay_mult; }
                else { state_ = 4; volts += d * k_hang_decay_mult; }
            } else {                                            // state 2: hang
                if (hang_counter == 0) { state_ = 4; volts += d * k_hang_decay_mult; }
            }
            if (volts < k_min_volts) volts = k_min_volts;
            RV[i] = volts;
        }
        st[2] = last_rm; st[3] = volts; st[4] = save_volts; st[5] = fast_backaverage; st[6] = hang_backaverage;
        st[7] = hang_counter; st[8] = decay_type; st[9] = state_; st[10] = volts * a.inv_out_target;
    }
    __syncthreads();
    // gain law and the delayed sample (the one written attack_buffsize inputs ago): history for i < ab, else this block
    for (int i = tid; i < n; i += nt) {
        const double volts = RV[i];
        const double lg = log10(a.inv_max_input * volts);
        const double mult = (a.out_target - a.slope_constant * (0.0 < lg ? 0.0 : lg)) / volts;
        cd o;
        if (i < ab) o = make_double2(hs[i * 3], hs[i * 3 + 1]);
        else o = inplace ? XS[i - ab] : x[i - ab];
        y[i] = make_double2(o.x * mult, o.y * mult);
    }
    __syncthreads();            // every read of the old history is done
    // new history = combined[n, n + ab): old history shifted down by n (only when n < ab), then the block's tail
    if (n < ab) {
        const int keep = ab - n;
        for (int j0 = 0; j0 < keep; j0 += nt) {             // in-place forward shift, one stripe at a time
            const int j = j0 + tid;
            double r0 = 0, r1 = 0;
            if (j < keep) { r0 = hs[(n + j) * 3]; r1 = hs[(n + j) * 3 + 1]; }
            __syncthreads();
            if (j < keep) { hs[j * 3] = r0; hs[j * 3 + 1] = r1; hs[j * 3 + 2] = A[n + j]; }
            __syncthreads();
        }
    }
    for (int j = tid + max(ab - n, 0); j < ab; j += nt) {
        const int k = n + j - ab;                            // index into the block
        const cd v = inplace ? XS[k] : x[k];
        hs[j * 3] = v.x; hs[j * 3 + 1] = v.y; hs[j * 3 + 2] = A[n + j];
    }
}

// --------------------------------------------------------------------------------------------- amd
// state: 0 dc 1 dc_insert 2 phs 3 fil_out 4 omega 5 dsI 6 dsQ 7.. a[24] b[24] c[24] d[24]
// par:   0 mode 1 levelfade 2 sbmode 3 omega_min 4 omega_max 5 g1 6 g2 7 mtauR 8 onem_mtauR 9 mtauI 10 onem_mtauI
__constant__ double c_amd_c0[7] = {-0.328201924180698, -0.744171491539427, -0.923022915444215, -0.978490468768238,
                                   -0.994128272402075, -0.998458978159551, -0.999790306259206};      // amd.c:92-98
__constant__ double c_amd_c1[7] = {-0.0991227952747244, -0.565619728761389, -0.857467122550052, -0.959123933111275,
                                   -0.988739372718090, -0.996959189310611, -0.999282492800792};      // amd.c:100-106

struct SeqPar { double v[32]; };

__global__ void amd_kernel(const cd *in, long is, cd *out, long os, int n, int C, double *state, SeqPar P)
{
    SEQ_STAGE_IN();
    if (threadIdx.x == 0) {
    const cd *x = sx;
    cd *y = sx;
    double *st = state + (size_t)c * 104;
    const int mode = (int)P.v[0], levelfade = (int)P.v[1], sbmode = (int)P.v[2];
    const double omega_min = P.v[3], omega_max = P.v[4], g1 = P.v[5], g2 = P.v[6];
    const double mtauR = P.v[7], onem_mtauR = P.v[8], mtauI = P.v[9], onem_mtauI = P.v[10];
    double dc = st[0], dc_insert = st[1], phs = st[2], fil_out = st[3], omega = st[4], dsI = st[5], dsQ = st[6];
    if (mode == 0) {
        for (int i = 0; i < n; i++) {
            double audio = sqrt(x[i].x * x[i].x + x[i].y * x[i].y);
            if (levelfade) {
                dc = mtauR * dc + onem_mtauR * audio;
                dc_insert = mtauI * dc_insert + onem_mtauI * audio;
                audio += dc_insert - dc;
            }
            y[i] = make_double2(audio, audio);
        }
    } else {
        double *A = st + 7, *B = st + 31, *Cc = st + 55, *D = st + 79;      // a[], b[], c[], d[] (amd.h:64-67)
        for (int i = 0; i < n; i++) {
            const double v0 = cos(phs), v1 = sin(phs);
            const double ai = x[i].x * v0, bi = x[i].x * v1, aq = x[i].y * v0, bq = x[i].y * v1;
            double ai_ps = 0, bi_ps = 0, aq_ps = 0, bq_ps = 0;
            if (sbmode != 0) {
                A[0] = dsI; B[0] = bi; Cc[0] = dsQ; D[0] = aq;
                dsI = ai; dsQ = bq;
                for (int j = 0; j < 7; j++) {
                    const int k = 3 * j;
                    A[k + 3] = c_amd_c0[j] * (A[k] - A[k + 5]) + A[k + 2];
                    B[k + 3] = c_amd_c1[j] * (B[k] - B[k + 5]) + B[k + 2];
                    Cc[k + 3] = c_amd_c0[j] * (Cc[k] - Cc[k + 5]) + Cc[k + 2];
                    D[k + 3] = c_amd_c1[j] * (D[k] - D[k + 5]) + D[k + 2];
                }
                ai_ps = A[21]; bi_ps = B[21]; bq_ps = Cc[21]; aq_ps = D[21];
                for (int j = 23; j > 0; j--) { A[j] = A[j - 1]; B[j] = B[j - 1]; Cc[j] = Cc[j - 1]; D[j] = D[j - 1]; }
            }
            double corr0 = +ai + bq;
            const double corr1 = -bi + aq;
            double audio;
            if (sbmode == 1) audio = (ai_ps - bi_ps) + (aq_ps + bq_ps);
            else if (sbmode == 2) audio = (ai_ps + bi_ps) - (aq_ps - bq_ps);
            else audio = corr0;
            if (levelfade) {
                dc = mtauR * dc + onem_mtauR * audio;
                dc_insert = mtauI * dc_insert + onem_mtauI * corr0;
                audio += dc_insert - dc;
            }
            y[i] = make_double2(audio, audio);
            if (corr0 == 0.0 && corr1 == 0.0) corr0 = 1.0;
            const double det = atan2(corr1, corr0);
            const double del_out = fil_out;
            omega += g2 * det;
            if (omega < omega_min) omega = omega_min;
            if (omega > omega_max) omega = omega_max;
            fil_out = g1 * det + omega;
            phs += del_out;
            while (phs >= TWOPI_D) phs -= TWOPI_D;
            while (phs < 0.0) phs += TWOPI_D;
        }
    }
    st[0] = dc; st[1] = dc_insert; st[2] = phs; st[3] = fil_out; st[4] = omega; st[5] = dsI; st[6] = dsQ;
    }
    SEQ_STAGE_OUT();
}

// ------------------------------------------------------------------------------------------- fm pll
// state: 0 phs 1 fil_out 2 omega 3 fmdc ; par: 0 omega_min 1 omega_max 2 g1 3 g2 4 mtau 5 onem_mtau 6 again
// The reference computes det = atan2 of x * conj(vco), vco = (cos phs, sin phs) (fmd.c:154-159): two libm calls with
// ~150 dependent FP64 instructions in front of every step of the loop.  But arg(x e^{-j phs}) = arg(x) - phs wrapped
// into (-pi, pi], and arg(x) does not depend on the loop: every thread evaluates atan2 for its share of the block up
// front and lane 0 is left with subtract / wrap / two multiply-adds / clamp per sample (~10x shorter chain, 172 -> 20 us
// per 256-sample block).  The two forms differ by the rounding of one atan2 (<= 2e-16 in det); the loop is contractive,
// the fixture of the reference's xfmd is matched to 1e-10.
__global__ void fmpll_kernel(const cd *in, long is, cd *out, long os, int n, int C, double *state, SeqPar P)
{
    // shared memory: ang[n] only (the samples are not staged: every thread takes arg() of its share straight from global
    // memory, lane 0's loop leaves the audio value in place of the angle, and all threads write the (a, a) pairs out)
    extern __shared__ double seq_smem[];
    double *ang = seq_smem;                                     // [n] arg(x[i]); 1e300 marks x == 0 (fmd.c:158: det = 0)
    const int c = blockIdx.x;
    const cd *gx = in + (size_t)c * is;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const cd v = gx[i];
        ang[i] = (v.x == 0.0 && v.y == 0.0) ? 1.0e300 : atan2(v.y, v.x);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
    double *st = state + (size_t)c * 4;
    double phs = st[0], fil_out = st[1], omega = st[2], fmdc = st[3];
    const double omega_min = P.v[0], omega_max = P.v[1], g1 = P.v[2], g2 = P.v[3], mtau = P.v[4], onem_mtau = P.v[5], again = P.v[6];
    // The loop's only true cycle is phs -> det -> omega -> fil_out -> (two samples later) phs; everything is written so that
    // nothing else sits on it: angles come in groups of eight through registers (32-bit shared-window addresses: no address
    // re-derivation per group), the clamps are min / max, the two wrap-arounds are selects with the reference's while-loops
    // as a never-taken fallback, the DC estimate and the output are side chains.  Same operations, same order, same
    // roundings as fmd.c:153-168.
    const unsigned ang_s = (unsigned)__cvta_generic_to_shared(ang);
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        double a8[8], o8[8];
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a8[j]) : "r"(ang_s + 8u * (unsigned)(i + j)));
        // eight samples without a branch: a branch on the wrapped phase in every sample makes the (in-order) warp wait for the end
        // of the phase chain before it may issue the next sample's det / omega / fil_out, which do not depend on it -- 155 cycles
        // per sample instead of ~55.  Whether a single wrap was enough is collected off the chain and looked at once per group;
        // if not (|del_out| > 2 pi: not with any loop bandwidth the stage is built with) the group is redone with the reference's loops.
        const double phs0 = phs, fil0 = fil_out, om0 = omega, dc0 = fmdc;
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double a_x = a8[j];
            double det = a_x - phs;                             // in (-3 pi, pi]
            const double detw = det + TWOPI_D;
            det = det <= -kPI ? detw : det;
            bad = bad || a_x > 1.0e299;                         // x == 0 (det = 0, fmd.c:158): the slow path handles it
            const double del_out = fil_out;
            omega = omega + g2 * det;                           // the two clamps of fmd.c:161-162 rarely act: tested off the chain
            bad = bad || omega < omega_min || omega > omega_max;
            fil_out = g1 * det + omega;
            const double p = phs + del_out;
            const double pm = p - TWOPI_D, pp = p + TWOPI_D;    // both candidates and both tests hang off p side by side
            const double pn = p >= TWOPI_D ? pm : (p < 0.0 ? pp : p);
            bad = bad || pn >= TWOPI_D || pn < 0.0;
            phs = pn;
            fmdc = mtau * fmdc + onem_mtau * fil_out;
            o8[j] = again * (fil_out - fmdc);
        }
        if (bad) {
            phs = phs0; fil_out = fil0; omega = om0; fmdc = dc0;
            for (int j = 0; j < 8; j++) {
                const double a_x = ang[i + j];
                double det = a_x - phs;
                if (det <= -kPI) det += TWOPI_D;
                if (a_x > 1.0e299) det = 0.0;
                const double del_out = fil_out;
                omega += g2 * det;
                if (omega < omega_min) omega = omega_min;
                if (omega > omega_max) omega = omega_max;
                fil_out = g1 * det + omega;
                phs += del_out;
                while (phs >= TWOPI_D) phs -= TWOPI_D;
                while (phs < 0.0) phs += TWOPI_D;
                fmdc = mtau * fmdc + onem_mtau * fil_out;
                ang[i + j] = again * (fil_out - fmdc);
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("st.shared.f64 [%0], %1;" :: "r"(ang_s + 8u * (unsigned)(i + j)), "d"(o8[j]) : "memory");
    }
    for (; i < n; i++) {
        const double a_x = ang[i];
        double det = a_x - phs;
        if (det <= -kPI) det += TWOPI_D;
        if (a_x > 1.0e299) det = 0.0;
        const double del_out = fil_out;
        omega += g2 * det;
        if (omega < omega_min) omega = omega_min;
        if (omega > omega_max) omega = omega_max;
        fil_out = g1 * det + omega;
        phs += del_out;
        while (phs >= TWOPI_D) phs -= TWOPI_D;
        while (phs < 0.0) phs += TWOPI_D;
        fmdc = mtau * fmdc + onem_mtau * fil_out;
        ang[i] = again * (fil_out - fmdc);
    }
    st[0] = phs; st[1] = fil_out; st[2] = omega; st[3] = fmdc;
    }
    __syncthreads();
    cd *gy = out + (size_t)c * os;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const double a = ang[i]; gy[i] = make_double2(a, a); }
}

// ------------------------------------------------------------------------------------------ snotch
// state: x1 x2 y1 y2 ; par: a0 a1 a2 b1 b2.  Only the I rail is filtered (iir.c:85-86).
__global__ void snotch_kernel(const cd *in, long is, cd *out, long os, int n, int C, double *state, SeqPar P)
{
    // only the I rail is filtered (iir.c:85-86): its n doubles are what is staged; the Q rail passes from `in` to `out`.
    // The feed-forward half of the biquad, ((a0 x0 + a1 x1) + a2 x2), only looks at the input: every thread forms it for its
    // share of the block, and the sequential lane is left with ff + b1 y1 + b2 y2 -- four FP64 instructions per sample
    // instead of nine (a single warp gets one FP64 instruction through its sub-partition's port every two cycles).
    extern __shared__ double seq_smem[];
    double *xr = seq_smem;                                      // [n] x, then ff, then y
    const int c = blockIdx.x;
    const cd *gx = in + (size_t)c * is;
    cd *gy = out + (size_t)c * os;
    double *st = state + (size_t)c * 4;
    const double a0 = P.v[0], a1 = P.v[1], a2 = P.v[2], b1 = P.v[3], b2 = P.v[4];
    for (int i = threadIdx.x; i < n; i += blockDim.x) xr[i] = gx[i].x;
    __syncthreads();
    const double sx1 = st[0], sx2 = st[1];
    // x1 / x2 for the next call: the last two inputs (or what is left of the old ones for n < 2), read before xr is overwritten
    const double nx1 = n >= 1 ? xr[n - 1] : sx1, nx2 = n >= 2 ? xr[n - 2] : (n == 1 ? sx1 : sx2);
    // in place, from the top down: a pass reads x[i], x[i - 1], x[i - 2] -- below it nothing has been overwritten yet
    const int B = blockDim.x;
    for (int i0 = ((n - 1) / B) * B; i0 >= 0; i0 -= B) {
        const int i = i0 + threadIdx.x;
        double ff = 0.0;
        if (i < n) {
            const double x0 = xr[i], x1 = i >= 1 ? xr[i - 1] : sx1, x2 = i >= 2 ? xr[i - 2] : (i == 1 ? sx1 : sx2);
            ff = a0 * x0 + a1 * x1 + a2 * x2;
        }
        __syncthreads();
        if (i < n) xr[i] = ff;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double y1 = st[2], y2 = st[3];
        const unsigned x_s = (unsigned)__cvta_generic_to_shared(xr);
        int i = 0;
        for (; i + 8 <= n; i += 8) {
            double f8[8];
#pragma unroll
            for (int j = 0; j < 8; j++) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(f8[j]) : "r"(x_s + 8u * (unsigned)(i + j)));
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double o = f8[j] + b1 * y1 + b2 * y2;
                asm volatile("st.shared.f64 [%0], %1;" :: "r"(x_s + 8u * (unsigned)(i + j)), "d"(o) : "memory");
                y2 = y1; y1 = o;
            }
        }
        for (; i < n; i++) {
            const double o = xr[i] + b1 * y1 + b2 * y2;
            xr[i] = o;
            y2 = y1; y1 = o;
        }
        st[2] = y1; st[3] = y2;
    }
    __syncthreads();
    if (threadIdx.x == 0) { st[0] = nx1; st[1] = nx2; }
    for (int i = threadIdx.x; i < n; i += blockDim.x) gy[i] = make_double2(xr[i], gx[i].y);
}

// mlog10 (wdsp/meterlog10.c:547-554): WDSP's meters do not call log10 -- they take the exponent and an 11-bit table
// look-up of log2(1 + m / 2048) on the leading mantissa bits, no interpolation, so a meter reads up to 2.1e-3 dB low.
// Part of the reference's observable behaviour (GetRXAMeter).  The table is rebuilt here from its definition
// (log10(x) / log10(2) reproduces 2045 of the reference's 2048 entries exactly, the other three to one ulp).
const double *mlog10_table()
{   // device copy of the table, per device (the library may be used on several GPUs from one process)
    static std::mutex mu;
    static double *tab[64] = {nullptr};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> g(mu);
    if (dev < 0 || dev >= 64) return nullptr;
    if (!tab[dev]) {
        std::vector<double> t(2048);
        for (int m = 0; m < 2048; m++) t[m] = log10(1.0 + (double)m / 2048.0) / log10(2.0);
        if (cudaMalloc((void **)&tab[dev], t.size() * sizeof(double)) != cudaSuccess) { tab[dev] = nullptr; return nullptr; }
        cudaMemcpy(tab[dev], t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice);
    }
    return tab[dev];
}

// ------------------------------------------------------------------------------------------- meter
// state: avg peak ; par: mult_average mult_peak ; results: av dB, pk dB (10 mlog10, meter.c:96-99)
// The averaging recurrence avg = avg * ma + (1 - ma) * |x|^2 is sequential in the reference's arithmetic, but its
// inputs are not: every thread forms w[i] = (1 - ma) * |x[i]|^2 and the block maximum in parallel, lane 0 is left with
// one multiply and one add per sample (the peak decay peak *= mp is a second, independent chain).
__global__ void meter_kernel(const cd *in, long is, int n, int C, double *state, SeqPar P, double *result, const double *agc_state, const double *mtable,
                             int nsub)
{
    // nsub > 1: the call covers nsub consecutive blocks of n samples (the wide multi-block path); the peak is decayed per sample
    // and raised to the block maximum at the END OF EACH BLOCK (meter.c:88-95), so the blocks are walked one after the other
    extern __shared__ double seq_smem[];
    double *w = seq_smem;                       // [n]
    __shared__ double s_max[2];
    const int c = blockIdx.x;
    const double ma = P.v[0], mp = P.v[1], oma = 1.0 - ma;
    double avg = state[c * 2], peak = state[c * 2 + 1];
    for (int sb = 0; sb < nsub; sb++) {
        const cd *gx = in + (size_t)c * is + (size_t)sb * n;
        double lm = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const cd v = gx[i];
            const double smag = v.x * v.x + v.y * v.y;
            w[i] = oma * smag;
            lm = smag > lm ? smag : lm;
        }
        for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, lm, o); lm = t > lm ? t : lm; }
        if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = lm;
        __syncthreads();
        if (threadIdx.x == 0) {
            const double np = s_max[0] > s_max[1] ? s_max[0] : s_max[1];
            for (int i = 0; i < n; i++) {
                avg = avg * ma + w[i];
                peak *= mp;
            }
            if (np > peak) peak = np;
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    state[c * 2] = avg; state[c * 2 + 1] = peak;
    result[c * 3] = 10.0 * mlog10_dev(mtable, avg + 1.0e-40);
    result[c * 3 + 1] = 10.0 * mlog10_dev(mtable, peak + 1.0e-40);
    result[c * 3 + 2] = agc_state ? 20.0 * mlog10_dev(mtable, agc_state[(size_t)c * 16 + 10] + 1.0e-40) : 0.0;
}

// sip != nullptr: also xsiphon mode 0 (siphon.c:96-129) on the INPUT block -- between xsiphon and xpanel the chain only
// has stages that are off by default (RXA.c:590-596), so the panel's input is what the siphon would see; the ring
// write rides on this kernel instead of costing a launch: sample i goes to slot (sip_idx + i) mod sipsize, or the
// last sipsize samples fill the ring when the block is at least that long.
__global__ void panel_kernel(const cd *in, long is, cd *out, long os, int n, int C, double gainI, double gainQ, int inselect, int copy,
                             cd *sip, int sipsize, int sip_idx)
{
    const long total = (long)n * C;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int c = (int)(idx / n);
        const int i = (int)(idx - (long)c * n);
        const cd v = in[(size_t)c * is + i];
        if (sip) {
            if (n >= sipsize) { if (i >= n - sipsize) sip[(size_t)c * sipsize + (i - (n - sipsize))] = v; }
            else sip[(size_t)c * sipsize + ((sip_idx + i) & (sipsize - 1))] = v;
        }
        double I, Q;
        switch (copy) {
        case 1: I = v.x * (inselect >> 1); Q = I; break;
        case 2: Q = v.y * (inselect & 1); I = Q; break;
        case 3: Q = v.x * (inselect >> 1); I = v.y * (inselect & 1); break;
        default: I = v.x * (inselect >> 1); Q = v.y * (inselect & 1); break;
        }
        out[(size_t)c * os + i] = make_double2(gainI * I, gainQ * Q);
    }
}

int launch_panel(const cd *in, long in_stride, cd *out, long out_stride, int n, int C, double gainI, double gainQ,
                 int inselect, int copy, cudaStream_t s, cd *sip, int sipsize, int sip_idx)
{
    const long total = (long)n * C;
    if (total <= 0) return QC_OK;
    int blocks = (int)((total + 255) / 256 < 148L * 16 ? (total + 255) / 256 : 148L * 16);
    panel_kernel<<<blocks, 256, 0, s>>>(in, in_stride, out, out_stride, n, C, gainI, gainQ, inselect, copy, sip, sipsize, sip_idx);
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

// ------------------------------------------------------------------------------------------- host
int SeqStage::init_common(int kind_, int C_, int sd)
{
    kind = kind_; C = C_; state_doubles = sd;
    if (C <= 0) { set_error("seq stage: bad channel count"); return QC_EINVAL; }
    QC_CUDA(cudaMalloc((void **)&d_state, (size_t)C * sd * sizeof(double)));
    return flush();
}

void SeqStage::release()
{
    if (d_state) cudaFree(d_state); if (d_ring) cudaFree(d_ring); if (d_par) cudaFree(d_par); if (d_meter) cudaFree(d_meter);
    d_state = d_ring = d_par = d_meter = nullptr;
}

int SeqStage::flush()
{
    QC_CUDA(cudaMemset(d_state, 0, (size_t)C * state_doubles * sizeof(double)));
    if (kind == SEQ_WCPAGC) {
        // calc_wcpagc (wcpAGC.c:35-52): out_index = -1, in_index = attack_buffsize + out_index, everything else 0
        std::vector<double> st((size_t)C * 16, 0.0);
        for (int c = 0; c < C; c++) { st[(size_t)c * 16] = -1.0; st[(size_t)c * 16 + 1] = agc.attack_buffsize - 1.0; }     // informational
        QC_CUDA(cudaMemcpy(d_state, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice));
        if (d_ring) QC_CUDA(cudaMemset(d_ring, 0, (size_t)C * ring_len * 3 * sizeof(double)));
    }
    if (kind == SEQ_METER && d_meter) {
        std::vector<double> r((size_t)C * 3, -400.0);
        QC_CUDA(cudaMemcpy(d_meter, r.data(), r.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    return QC_OK;
}

int SeqStage::flush_ref()
{   // The reference's own flush_<stage>, which for two stages resets LESS than a fresh object:
    //   flush_wcpagc (wcpAGC.c:154-159): ring, abs_ring, ring_max -- volts, the state machine and the back averages live on;
    //   flush_amd    (amd.c:109-113):    dc, dc_insert             -- the PLL and the phasing network live on.
    // flush_shift, flush_fmd + flush_snotch, flush_meter reset everything the stage carries.
    if (kind == SEQ_WCPAGC) {
        if (d_ring) QC_CUDA(cudaMemset(d_ring, 0, (size_t)C * ring_len * 3 * sizeof(double)));
        return cudaMemset2D(d_state + 2, 16 * sizeof(double), 0, sizeof(double), C) == cudaSuccess ? QC_OK : QC_ECUDA;     // ring_max
    }
    if (kind == SEQ_AMD)
        return cudaMemset2D(d_state, (size_t)state_doubles * sizeof(double), 0, 2 * sizeof(double), C) == cudaSuccess ? QC_OK : QC_ECUDA;
    return flush();
}

void SeqStage::load_agc()
{   // loadWcpAGC, wcpAGC.c:115-147
    AgcParams &a = agc;
    a.attack_buffsize = (int)ceil(a.sample_rate * a.n_tau * a.tau_attack);
    a.attack_mult = 1.0 - exp(-1.0 / (a.sample_rate * a.tau_attack));
    a.decay_mult = 1.0 - exp(-1.0 / (a.sample_rate * a.tau_decay));
    a.fast_decay_mult = 1.0 - exp(-1.0 / (a.sample_rate * a.tau_fast_decay));
    a.fast_backmult = 1.0 - exp(-1.0 / (a.sample_rate * a.tau_fast_backaverage));
    a.onemfast_backmult = 1.0 - a.fast_backmult;
    a.out_target = a.out_targ * (1.0 - exp(-(double)a.n_tau)) * 0.9999;
    a.min_volts = a.out_target / (a.var_gain * a.max_gain);
    a.inv_out_target = 1.0 / a.out_target;
    double tmp = log10(a.out_target / (a.max_input * a.var_gain * a.max_gain));
    if (tmp == 0.0) tmp = 1e-16;
    a.slope_constant = (a.out_target * (1.0 - 1.0 / a.var_gain)) / tmp;
    a.inv_max_input = 1.0 / a.max_input;
    tmp = pow(10.0, (a.hang_thresh - 1.0) / 0.125);
    a.hang_level = (a.max_input * tmp + (a.out_target / (a.var_gain * a.max_gain)) * (1.0 - tmp)) * 0.637;
    a.hang_backmult = 1.0 - exp(-1.0 / (a.sample_rate * a.tau_hang_backmult));
    a.onemhang_backmult = 1.0 - a.hang_backmult;
    a.hang_decay_mult = 1.0 - exp(-1.0 / (a.sample_rate * a.tau_hang_decay));
}

void agc_set_mode(SeqStage *s, int mode)
{   // SetRXAAGCMode, wcpAGC.c:370-411
    AgcParams &a = s->agc;
    switch (mode) {
    case 0: a.mode = 0; break;
    case 1: a.mode = 1; a.hangtime = 2.000; a.tau_decay = 2.000; break;
    case 2: a.mode = 2; a.hangtime = 1.000; a.tau_decay = 0.500; break;
    case 3: a.mode = 3; a.hang_thresh = 1.0; a.hangtime = 0.000; a.tau_decay = 0.250; break;
    case 4: a.mode = 4; a.hang_thresh = 1.0; a.hangtime = 0.000; a.tau_decay = 0.050; break;
    default: a.mode = 5; return;
    }
    s->load_agc();
}

int SeqStage::run(const void *d_in, long is, void *d_out, long os, int n, cudaStream_t s)
{
    if (n <= 0) return QC_OK;
    SeqPar P;
    memcpy(P.v, par, sizeof(P.v));
    const cd *in = (const cd *)d_in; cd *out = (cd *)d_out;
    const size_t sh = (size_t)n * sizeof(cd);           // the staged block
    if (kind != SEQ_METER && kind != SEQ_WCPAGC && sh > 200 * 1024) { set_error("seq stage: block of %d samples does not fit in shared memory", n); return QC_EINVAL; }
#define QC_SEQ_OPTIN(k, bytes) do { if ((bytes) > 48 * 1024) QC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); } while (0)
    switch (kind) {
    case SEQ_SHIFT: QC_SEQ_OPTIN(shift_kernel, sh); shift_kernel<<<C, SEQ_T, sh, s>>>(in, is, out, os, n, C, d_state, d_par); break;
    case SEQ_WCPAGC: {
        const int tot = agc.attack_buffsize + n;
        const size_t sa = (size_t)tot * 8 + (size_t)n * 8 + (size_t)((tot + 31) / 32 + 4) * 8 + (d_in == d_out ? (size_t)n * 16 : 0);
        if (sa > 220 * 1024) { set_error("wcpagc: attack buffer + block (%d samples) exceed shared memory", tot); return QC_EINVAL; }
        QC_SEQ_OPTIN(wcpagc_kernel, sa);
        wcpagc_kernel<<<C, 128, sa, s>>>(in, is, out, os, n, C, d_state, d_ring, agc);
        break; }
    case SEQ_AMD: QC_SEQ_OPTIN(amd_kernel, sh); amd_kernel<<<C, SEQ_T, sh, s>>>(in, is, out, os, n, C, d_state, P); break;
    case SEQ_FMPLL: { const size_t sf = (size_t)n * sizeof(double); QC_SEQ_OPTIN(fmpll_kernel, sf); fmpll_kernel<<<C, 128, sf, s>>>(in, is, out, os, n, C, d_state, P); break; }
    case SEQ_SNOTCH: { const size_t sf = (size_t)n * sizeof(double); QC_SEQ_OPTIN(snotch_kernel, sf); snotch_kernel<<<C, SEQ_T, sf, s>>>(in, is, out, os, n, C, d_state, P); break; }
    case SEQ_METER: { const double *mt = mlog10_table(); if (!mt) { set_error("meter: table allocation failed"); return QC_ENOMEM; }
        const int ns = meter_sub > 1 ? meter_sub : 1, nb = n / ns;         // meter_sub blocks of nb samples in this call
        const size_t shm = (size_t)nb * sizeof(double);
        QC_SEQ_OPTIN(meter_kernel, shm); meter_kernel<<<C, SEQ_T, shm, s>>>(in, is, nb, C, d_state, P, d_meter, (const double *)d_out, mt, ns); break; }
    default: set_error("seq stage: unknown kind %d", kind); return QC_EINVAL;
    }
    count_launch();
    QC_CUDA_LAUNCH();
    return QC_OK;
}

SeqStage *make_shift(int C, int rate, const double *shift_hz)
{
    SeqStage *s = new SeqStage();
    if (s->init_common(SEQ_SHIFT, C, 1) != QC_OK) { s->release(); delete s; return nullptr; }
    std::vector<double> p((size_t)C * 3);
    for (int c = 0; c < C; c++) {       // calc_shift, shift.c:29-34
        const double delta = kTWOPI * (shift_hz ? shift_hz[c] : 0.0) / (double)rate;
        p[c * 3] = delta; p[c * 3 + 1] = cos(delta); p[c * 3 + 2] = sin(delta);
    }
    if (cudaMalloc((void **)&s->d_par, p.size() * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(s->d_par, p.data(), p.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) { s->release(); delete s; return nullptr; }
    return s;
}

SeqStage *make_wcpagc_fmlim(int C, int rate, double lim_gain)
{   // the FM detector limiter: create_wcpagc's arguments in calc_fmd (wdsp/fmd.c:49-73)
    SeqStage *s = new SeqStage();
    AgcParams &a = s->agc;
    memset(&a, 0, sizeof(a));
    a.mode = 5; a.pmode = 1; a.sample_rate = (double)rate; a.tau_attack = 0.001; a.tau_decay = 0.008; a.n_tau = 4;
    a.max_gain = lim_gain; a.var_gain = 1.0; a.fixed_gain = 1.0; a.max_input = 1.0; a.out_targ = 0.9;
    a.tau_fast_backaverage = 0.250; a.tau_fast_decay = 0.004; a.pop_ratio = 4.0; a.hang_enable = 0;
    a.tau_hang_backmult = 0.500; a.hangtime = 0.500; a.hang_thresh = 2.000; a.tau_hang_decay = 0.100;
    s->load_agc();
    s->kind = SEQ_WCPAGC;
    s->ring_len = a.attack_buffsize;
    a.ring_buffsize = s->ring_len;
    if (cudaMalloc((void **)&s->d_ring, (size_t)C * s->ring_len * 3 * sizeof(double)) != cudaSuccess) { delete s; return nullptr; }
    if (s->init_common(SEQ_WCPAGC, C, 16) != QC_OK) { s->release(); delete s; return nullptr; }
    return s;
}

SeqStage *make_wcpagc(int C, int rate, int mode)
{
    SeqStage *s = new SeqStage();
    AgcParams &a = s->agc;
    memset(&a, 0, sizeof(a));
    // create_rxa's arguments to create_wcpagc (RXA.c:337-360)
    a.mode = 3; a.pmode = 1; a.sample_rate = (double)rate; a.tau_attack = 0.001; a.tau_decay = 0.250; a.n_tau = 4;
    a.max_gain = 10000.0; a.var_gain = 1.5; a.fixed_gain = 1000.0; a.max_input = 1.0; a.out_targ = 1.0;
    a.tau_fast_backaverage = 0.250; a.tau_fast_decay = 0.005; a.pop_ratio = 5.0; a.hang_enable = 1;
    a.tau_hang_backmult = 0.500; a.hangtime = 0.250; a.hang_thresh = 0.250; a.tau_hang_decay = 0.100;
    s->load_agc();
    agc_set_mode(s, mode);
    s->kind = SEQ_WCPAGC;
    s->ring_len = a.attack_buffsize;        // the reference ring is RB_SIZE long; only the last attack_buffsize inputs are live
    a.ring_buffsize = s->ring_len;
    if (cudaMalloc((void **)&s->d_ring, (size_t)C * s->ring_len * 3 * sizeof(double)) != cudaSuccess) { delete s; return nullptr; }
    if (s->init_common(SEQ_WCPAGC, C, 16) != QC_OK) { s->release(); delete s; return nullptr; }
    return s;
}

SeqStage *make_amd(int C, int rate, int mode, int levelfade, int sbmode)
{
    SeqStage *s = new SeqStage();
    if (s->init_common(SEQ_AMD, C, 104) != QC_OK) { s->release(); delete s; return nullptr; }
    // init_amd (amd.c:72-107) with create_rxa's constants (RXA.c:175-189)
    const double sr = (double)rate, fmin = -2000.0, fmax = +2000.0, zeta = 1.0, omegaN = 250.0, tauR = 0.02, tauI = 1.4;
    double *p = s->par;
    p[0] = mode; p[1] = levelfade; p[2] = sbmode;
    p[3] = kTWOPI * fmin / sr; p[4] = kTWOPI * fmax / sr;
    p[5] = 1.0 - exp(-2.0 * omegaN * zeta / sr);
    p[6] = -p[5] + 2.0 * (1 - exp(-omegaN * zeta / sr) * cos(omegaN / sr * sqrt(1.0 - zeta * zeta)));
    p[7] = exp(-1.0 / (sr * tauR)); p[8] = 1.0 - p[7];
    p[9] = exp(-1.0 / (sr * tauI)); p[10] = 1.0 - p[9];
    return s;
}

SeqStage *make_fmpll(int C, int rate, double deviation, double fmin, double fmax, double zeta, double omegaN, double tau)
{
    SeqStage *s = new SeqStage();
    if (s->init_common(SEQ_FMPLL, C, 4) != QC_OK) { s->release(); delete s; return nullptr; }
    const double sr = (double)rate;         // calc_fmd, fmd.c:29-47
    double *p = s->par;
    p[0] = kTWOPI * fmin / sr; p[1] = kTWOPI * fmax / sr;
    p[2] = 1.0 - exp(-2.0 * omegaN * zeta / sr);
    p[3] = -p[2] + 2.0 * (1 - exp(-omegaN * zeta / sr) * cos(omegaN / sr * sqrt(1.0 - zeta * zeta)));
    p[4] = exp(-1.0 / (sr * tau)); p[5] = 1.0 - p[4];
    p[6] = sr / (deviation * kTWOPI);
    return s;
}

SeqStage *make_snotch(int C, int rate, double f, double bw)
{
    SeqStage *s = new SeqStage();
    if (s->init_common(SEQ_SNOTCH, C, 4) != QC_OK) { s->release(); delete s; return nullptr; }
    // calc_snotch, iir.c:29-48
    const double fn = f / (double)rate;
    const double csn = cos(kTWOPI * fn);
    const double qr = 1.0 - 3.0 * bw;
    const double qk = (1.0 - 2.0 * qr * csn + qr * qr) / (2.0 * (1.0 - csn));
    double *p = s->par;
    p[0] = +qk; p[1] = -2.0 * qk * csn; p[2] = +qk; p[3] = +2.0 * qr * csn; p[4] = -qr * qr;
    return s;
}

SeqStage *make_meter(int C, int rate, double tau_av, double tau_decay)
{
    SeqStage *s = new SeqStage();
    s->kind = SEQ_METER;
    if (cudaMalloc((void **)&s->d_meter, (size_t)C * 3 * sizeof(double)) != cudaSuccess) { delete s; return nullptr; }
    if (s->init_common(SEQ_METER, C, 2) != QC_OK) { s->release(); delete s; return nullptr; }
    s->par[0] = exp(-1.0 / ((double)rate * tau_av));        // calc_meter, meter.c:30-34
    s->par[1] = exp(-1.0 / ((double)rate * tau_decay));
    return s;
}

}  // namespace qc

struct qcSeqStage { qc::SeqStage *s; };

extern "C" {

static qcSeqStage *wrap(qc::SeqStage *s) { if (!s) return nullptr; qcSeqStage *w = new qcSeqStage(); w->s = s; return w; }

qcSeqStage *quisk_cuda_shift_create(int n_channels, int rate, const double *shift_hz)
{ return qc::ensure_device() == QC_OK ? wrap(qc::make_shift(n_channels, rate, shift_hz)) : nullptr; }
qcSeqStage *quisk_cuda_wcpagc_create_fmlim(int n_channels, int rate, double lim_gain)
{ return qc::ensure_device() == QC_OK ? wrap(qc::make_wcpagc_fmlim(n_channels, rate, lim_gain)) : nullptr; }   // the FM detector limiter as a stage of its own (fmd.c:49-73)
qcSeqStage *quisk_cuda_wcpagc_create(int n_channels, int rate, int mode)
{ return qc::ensure_device() == QC_OK ? wrap(qc::make_wcpagc(n_channels, rate, mode)) : nullptr; }
int quisk_cuda_wcpagc_set_fixed_gain_db(qcSeqStage *s, double gain_db)
{ if (!s || s->s->kind != qc::SEQ_WCPAGC) return QC_EINVAL; s->s->agc.fixed_gain = pow(10.0, gain_db / 20.0); s->s->load_agc(); return QC_OK; }
int quisk_cuda_wcpagc_set_top_db(qcSeqStage *s, double max_gain_db)
{ if (!s || s->s->kind != qc::SEQ_WCPAGC) return QC_EINVAL; s->s->agc.max_gain = pow(10.0, max_gain_db / 20.0); s->s->load_agc(); return QC_OK; }
qcSeqStage *quisk_cuda_amd_create(int n_channels, int rate, int mode, int levelfade, int sbmode)
{ return qc::ensure_device() == QC_OK ? wrap(qc::make_amd(n_channels, rate, mode, levelfade, sbmode)) : nullptr; }
qcSeqStage *quisk_cuda_fmpll_create(int n_channels, int rate, double deviation, double fmin, double fmax, double zeta, double omegaN, double tau)
{ return qc::ensure_device() == QC_OK ? wrap(qc::make_fmpll(n_channels, rate, deviation, fmin, fmax, zeta, omegaN, tau)) : nullptr; }
qcSeqStage *quisk_cuda_snotch_create(int n_channels, int rate, double f, double bw)
{ return qc::ensure_device() == QC_OK ? wrap(qc::make_snotch(n_channels, rate, f, bw)) : nullptr; }
void quisk_cuda_seq_destroy(qcSeqStage *s) { if (s) { s->s->release(); delete s->s; delete s; } }
int quisk_cuda_seq_run(qcSeqStage *s, const void *d_in, long in_stride, void *d_out, long out_stride, int n, void *stream)
{ return s ? s->s->run(d_in, in_stride, d_out, out_stride, n, (cudaStream_t)stream) : QC_EINVAL; }
int quisk_cuda_seq_flush(qcSeqStage *s) { return s ? s->s->flush() : QC_EINVAL; }

}  // extern "C"
