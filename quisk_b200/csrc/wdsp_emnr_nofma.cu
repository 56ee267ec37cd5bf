// quisk_b200/csrc/wdsp_emnr_nofma.cu -- WDSP's spectral noise reduction "NR2" (wdsp/emnr.c) for a batch of channels.
//
// xemnr (emnr.c:1015-1068) is a short-time Fourier transform with overlap-add: every `incr` = fsize / ovrlp new samples
// a frame of fsize samples is windowed (sqrt-Hamming, emnr.c:167-186), transformed (real -> fsize/2 + 1 bins), every
// bin is scaled by a gain in [0, gmax] that calc_gain (emnr.c:886-1013) derives from the bin's power and a running
// estimate of the noise power, transformed back, windowed again and added into the output accumulator.  The pieces:
//   * noise power estimate lambda_d, three methods: LambdaD (emnr.c:604-727, Martin's optimal smoothing + minimum
//     statistics), LambdaDs (emnr.c:729-742, speech-presence probability), LambdaDl (emnr.c:744-767);
//   * gain, methods 0 (Gaussian, linear amplitude: Ephraim-Malah with Bessel I0 / I1, emnr.c:43-125), 1 (Gaussian, log
//     amplitude: exponential integral, emnr.c:127-159), 2 (gamma speech prior: bilinear look-up in two 241 x 241 tables,
//     emnr.c:818-862).  Method 2 is what create_rxa selects (RXA.c:319-332) and what Quisk's NR2 button uses; its tables
//     are data of the WDSP distribution (wdsp/calculus.c, or the file `calculus` that calc_emnr prefers, emnr.c:313-323):
//     the host hands them over with quisk_cuda_emnr_set_tables, nothing of them is in this repository.  Method 3 (trained
//     thresholds, a second table set) is not built;
//   * aepf (emnr.c:769-816), the post-filter that smooths the gains across bins when the frame is mostly noise.
// Mapping: one CTA of 256 threads per channel and frame.  Everything per bin is parallel over the 2049 bins; the handful
// of sums over all bins that steer the estimator (emnr.c:620-631, 657, 774-779) are formed by ONE lane in the reference's
// order, because the minimum-statistics logic compares quantities derived from them and a differently rounded sum can
// tip a comparison; the moving averages of aepf are summed per bin in the reference's order as well.  The two
// transforms are the in-house complex FFT on 4096 points (the imaginary input is zero; the inverse gets the Hermitian
// extension).  State per channel lives in one device allocation; the ring indices are the same for every channel of
// the batch and stay on the host.
// Compiled with --fmad=false (file name rule in build.py): a * b + c rounds twice, as in the reference's x86-64 code.
#include <vector>
#include <mutex>
#include "fft_device.cuh"
#include "wdsp_internal.h"
#include "../../include/quisk_cuda_wdsp.h"

namespace qc {

struct EmnrPar {
    int fs, ms, ovrlp, incr, bsize, iasize, oasize;
    double gain;
    int gain_method, npe_method, ae_run;
    // g
    double gf1p5, alpha, eps_floor, gamma_max, xi_min, q, gmax;
    const double *GG, *GGS;
    // gain method 3 (emnr.c:329-334, 866-884): the trained zeta table, its validity map and mlog10's table
    const double *zeta_hat; const int *zeta_true; const double *mtable;
    int dim_zeta; double z_gamma_min, z_gamma_max, z_xihat_min, z_xihat_max, zeta_thresh;
    // np (LambdaD)
    double alphaCsmooth, alphaMax, alphaCmin, alphaMin_max_value, snrq, betamax, invQeqMax, av, MofD, MofV;
    int U, V, D;
    double invQbar_points[4], nsmax[4];
    // nps (LambdaDs)
    double alpha_pow, alpha_Pbar, epsH1, epsH1r;
    // npl (LambdaDl)
    double eta, gamma_l, beta_l, alpha_d, alpha_p, delta_LF, delta_MF, delta_0, delta_1, delta_2;
    // ae
    double zetaThresh, psi, t2;
};

// per-channel state, offsets in doubles into one row
struct EmnrLayout {
    size_t inaccum, outaccum, save, lambda_d, prev_gamma, prev_mask,
           p, sigma2N, pbar, p2bar, actmin, actmin_sub, lmin_flag, pmin_u, actminbuff, alphaC,
           s_sigma2N, s_Pbar, l_P, l_Pmin, l_p, l_D, row;
};

__device__ __forceinline__ double e_min(double a, double b) { return a < b ? a : b; }      // the reference's min / max macros
__device__ __forceinline__ double e_max(double a, double b) { return a > b ? a : b; }

// modified Bessel functions and the exponential integral as the reference evaluates them (emnr.c:43-159: the polynomial
// fits of Abramowitz & Stegun 9.8.1 - 9.8.4 and the series / continued fraction of Zhang & Jin's E1XB)
__device__ double e_bessI0(double x)
{
    if (x == 0.0) return 1.0;
    if (x < 0.0) x = -x;
    if (x <= 3.75) {
        double p = x / 3.75;
        p = p * p;
        return ((((( 0.0045813 * p + 0.0360768) * p + 0.2659732) * p + 1.2067492) * p + 3.0899424) * p + 3.5156229) * p + 1.0;
    }
    const double p = 3.75 / x;
    return exp(x) / sqrt(x) * (((((((( + 0.00392377 * p - 0.01647633) * p + 0.02635537) * p - 0.02057706) * p + 0.00916281) * p
                                   - 0.00157565) * p + 0.00225319) * p + 0.01328592) * p + 0.39894228);
}
__device__ double e_bessI1(double x)
{
    if (x == 0.0) return 0.0;
    if (x < 0.0) x = -x;
    if (x <= 3.75) {
        double p = x / 3.75;
        p = p * p;
        return x * (((((( 0.00032411 * p + 0.00301532) * p + 0.02658733) * p + 0.15084934) * p + 0.51498869) * p + 0.87890594) * p + 0.5);
    }
    const double p = 3.75 / x;
    return exp(x) / sqrt(x) * (((((((( - 0.00420059 * p + 0.01787654) * p - 0.02895312) * p + 0.02282967) * p - 0.01031555) * p
                                   + 0.00163801) * p - 0.00362018) * p - 0.03988024) * p + 0.39894228);
}
__device__ double e_e1xb(double x)
{
    if (x == 0.0) return 1.0e300;
    if (x <= 1.0) {
        double e1 = 1.0, r = 1.0;
        for (int k = 1; k <= 25; k++) {
            r = -r * k * x / ((k + 1.0) * (k + 1.0));
            e1 = e1 + r;
            if (fabs(r) <= fabs(e1) * 1.0e-15) break;
        }
        return -0.5772156649015328 - log(x) + x * e1;
    }
    const int m = 20 + (int)(80.0 / x);
    double t0 = 0.0;
    for (int k = m; k >= 1; k--) t0 = (double)k / (1.0 + k / (x + t0));
    return exp(-x) * (1.0 / (x + t0));
}

// bilinear look-up in a 241 x 241 table over 10 log10 of (gamma, xi) / 0.001 in quarter-dB cells (emnr.c:818-862)
__device__ double e_getKey(const double *__restrict__ type, double gamma, double xi)
{
    int ng1, ng2, nx1, nx2;
    double tg, tx;
    const double dmin = 0.001, dmax = 1000.0;
    if (gamma <= dmin) { ng1 = ng2 = 0; tg = 0.0; }
    else if (gamma >= dmax) { ng1 = ng2 = 240; tg = 60.0; }
    else { tg = 10.0 * log10(gamma / dmin); ng1 = (int)(4.0 * tg); ng2 = ng1 + 1; }
    if (xi <= dmin) { nx1 = nx2 = 0; tx = 0.0; }
    else if (xi >= dmax) { nx1 = nx2 = 240; tx = 60.0; }
    else { tx = 10.0 * log10(xi / dmin); nx1 = (int)(4.0 * tx); nx2 = nx1 + 1; }
    const double dg = (tg - 0.25 * ng1) / 0.25, dx = (tx - 0.25 * nx1) / 0.25;
    return (1.0 - dg) * (1.0 - dx) * type[241 * nx1 + ng1]
         + (1.0 - dg) * dx * type[241 * nx2 + ng1]
         + dg * (1.0 - dx) * type[241 * nx1 + ng2]
         + dg * dx * type[241 * nx2 + ng2];
}

static constexpr int EM_T = 256;

// one frame of every channel: emnr.c:1029-1058
__global__ void __launch_bounds__(EM_T) emnr_frame_kernel(EmnrPar P, EmnrLayout L, double *state, const double *__restrict__ window, const cd *tw,
                                                           int iaoutidx, int saveidx, int oainidx, int subwc, int amb_idx)
{
    extern __shared__ double smem_raw[];
    const int fs = P.fs, ms = P.ms, tid = threadIdx.x;
    cd *twl = reinterpret_cast<cd *>(smem_raw);
    cd *S = twl + fft_tw_entries(fs);                       // [fs] transform buffer; bins ms .. fs - 1 double as scratch between the transforms
    double *LY = reinterpret_cast<double *>(S + fs);        // [ms] lambda_y
    double *MK = LY + ms;                                   // [ms] mask (scratch before the gains exist)
    double *red = MK + ms;                                  // [16] sums handed from the sequential lane to everybody
    double *SCR = reinterpret_cast<double *>(S + ms);       // [>= ms] scratch in the unused upper bins (fs - ms complex = 2 (ms - 2) doubles)
    double *st = state + (size_t)blockIdx.x * L.row;
    fft_stage_twiddles(twl, tw, fs);
    const double *ina = st + L.inaccum;
    for (int i = tid; i < fs; i += EM_T) {
        int j = iaoutidx + i; if (j >= P.iasize) j -= P.iasize;
        S[fsw(i)] = make_double2(window[i] * ina[j], 0.0);
    }
    __syncthreads();
    fft_smem<1>(S, fs, twl, -1, tid, EM_T);
    __syncthreads();
    // the bins this thread owns: k = tid + EM_T * j.  Y stays in S[fsw(k)], k < ms, until the inverse transform is set up.
    for (int k = tid; k < ms; k += EM_T) { const cd y = S[fsw(k)]; LY[k] = y.x * y.x + y.y * y.y; }
    __syncthreads();
    double *lambda_d = st + L.lambda_d;
    if (P.npe_method == 0) {
        // ---- LambdaD (emnr.c:604-727)
        double *p = st + L.p, *sigma2N = st + L.sigma2N, *pbar = st + L.pbar, *p2bar = st + L.p2bar, *actmin = st + L.actmin,
               *actmin_sub = st + L.actmin_sub, *lmin_flag = st + L.lmin_flag, *pmin_u = st + L.pmin_u, *amb = st + L.actminbuff;
        for (int k = tid; k < ms; k += EM_T) { SCR[k] = p[k]; MK[k] = sigma2N[k]; }
        __syncthreads();
        if (tid == 0) {     // the three sums in bin order
            double sp = 0.0, sy = 0.0, sn = 0.0;
            for (int k = 0; k < ms; k++) { sp += SCR[k]; sy += LY[k]; sn += MK[k]; }
            red[0] = sp; red[1] = sy; red[2] = sn;
        }
        __syncthreads();
        const double sum_prev_p = red[0], sum_lambda_y = red[1], sum_prev_sigma2N = red[2];
        const double SNR = sum_prev_p / sum_prev_sigma2N;
        const double alphaMin = e_min(P.alphaMin_max_value, pow(SNR, P.snrq));
        const double f1 = sum_prev_p / sum_lambda_y - 1.0;
        const double alphaCtilda = 1.0 / (1.0 + f1 * f1);
        const double alphaC = P.alphaCsmooth * st[L.alphaC] + (1.0 - P.alphaCsmooth) * e_max(alphaCtilda, P.alphaCmin);
        const double f2 = P.alphaMax * alphaC;
        __syncthreads();                                    // everybody has read the old alphaC
        if (tid == 0) st[L.alphaC] = alphaC;
        for (int k = tid; k < ms; k += EM_T) {
            const double pk0 = SCR[k], s2 = MK[k];
            const double f0 = pk0 / s2 - 1.0;
            double aoh = 1.0 / (1.0 + f0 * f0);
            if (aoh < alphaMin) aoh = alphaMin;
            const double ah = f2 * aoh;
            const double pk = ah * pk0 + (1.0 - ah) * LY[k];
            p[k] = pk;
            const double beta = e_min(P.betamax, ah * ah);
            const double pb = beta * pbar[k] + (1.0 - beta) * pk;
            const double p2b = beta * p2bar[k] + (1.0 - beta) * pk * pk;
            pbar[k] = pb; p2bar[k] = p2b;
            const double varHat = p2b - pb * pb;
            double invQeq = varHat / (2.0 * s2 * s2);
            if (invQeq > P.invQeqMax) invQeq = P.invQeqMax;
            SCR[k] = invQeq;                                // for the sum in bin order; Qeq = 1 / invQeq is formed again below
        }
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int k = 0; k < ms; k++) s += SCR[k];
            red[3] = s / (double)ms;
        }
        __syncthreads();
        const double invQbar = red[3];
        const double bc = 1.0 + P.av * sqrt(invQbar);
        double noise_slope_max = P.nsmax[3];
        if (invQbar < P.invQbar_points[0]) noise_slope_max = P.nsmax[0];
        else if (invQbar < P.invQbar_points[1]) noise_slope_max = P.nsmax[1];
        else if (invQbar < P.invQbar_points[2]) noise_slope_max = P.nsmax[2];
        for (int k = tid; k < ms; k += EM_T) {
            const double Qeq = 1.0 / SCR[k];
            const double QeqTilda = (Qeq - 2.0 * P.MofD) / (1.0 - P.MofD);
            const double QeqTildaSub = (Qeq - 2.0 * P.MofV) / (1.0 - P.MofV);
            const double bmin = 1.0 + 2.0 * (P.D - 1.0) / QeqTilda;
            const double bmin_sub = 1.0 + 2.0 * (P.V - 1.0) / QeqTildaSub;
            const double pk = p[k];
            const double f3 = pk * bmin * bc;
            int k_mod = 0;
            double am = actmin[k], ams = actmin_sub[k];
            if (f3 < am) { am = f3; ams = pk * bmin_sub * bc; k_mod = 1; }
            double lf = lmin_flag[k], pmu = pmin_u[k], s2 = MK[k];
            if (subwc == P.V) {
                if (k_mod) lf = 0.0;
                amb[(size_t)amb_idx * ms + k] = am;
                double mn = 1.0e300;
                for (int ku = 0; ku < P.U; ku++) { const double v = amb[(size_t)ku * ms + k]; if (v < mn) mn = v; }
                pmu = mn;
                if (lf == 1.0 && ams < noise_slope_max * pmu && ams > pmu) {
                    pmu = ams;
                    for (int ku = 0; ku < P.U; ku++) amb[(size_t)ku * ms + k] = ams;
                }
                lf = 0.0; am = 1.0e300; ams = 1.0e300;
            } else if (subwc > 1) {
                if (k_mod) { lf = 1.0; s2 = e_min(ams, pmu); pmu = s2; }
            }
            actmin[k] = am; actmin_sub[k] = ams; lmin_flag[k] = lf; pmin_u[k] = pmu; sigma2N[k] = s2;
            lambda_d[k] = s2;
        }
    } else if (P.npe_method == 1) {
        // ---- LambdaDs (emnr.c:729-742)
        double *sigma2N = st + L.s_sigma2N, *Pbar = st + L.s_Pbar;
        for (int k = tid; k < ms; k += EM_T) {
            double PH1y = 1.0 / (1.0 + (1.0 + P.epsH1) * exp(-P.epsH1r * LY[k] / sigma2N[k]));
            const double pb = P.alpha_Pbar * Pbar[k] + (1.0 - P.alpha_Pbar) * PH1y;
            Pbar[k] = pb;
            if (pb > 0.99) PH1y = e_min(PH1y, 0.99);
            const double EN2y = (1.0 - PH1y) * LY[k] + PH1y * sigma2N[k];
            const double s2 = P.alpha_pow * sigma2N[k] + (1.0 - P.alpha_pow) * EN2y;
            sigma2N[k] = s2; lambda_d[k] = s2;
        }
    } else {
        // ---- LambdaDl (emnr.c:744-767)
        double *Pk = st + L.l_P, *Pmin = st + L.l_Pmin, *pp = st + L.l_p, *D = st + L.l_D;
        const double c = (1.0 - P.gamma_l) / (1.0 - P.beta_l);
        for (int k = tid; k < ms; k += EM_T) {
            const double P_old = Pk[k];
            const double Pn = P.eta * P_old + (1.0 - P.eta) * LY[k];
            Pk[k] = Pn;
            double pm = Pmin[k];
            if (pm < Pn) pm = P.gamma_l * pm + c * (Pn - P.beta_l * P_old);
            else pm = Pn;
            Pmin[k] = pm;
            const double Sr = Pn / pm;
            const double delta = (double)k <= P.delta_LF ? P.delta_0 : ((double)k <= P.delta_MF ? P.delta_1 : P.delta_2);
            const double I = Sr > delta ? 1.0 : 0.0;
            const double pq = P.alpha_p * pp[k] + (1.0 - P.alpha_p) * I;
            pp[k] = pq;
            const double alpha_s = P.alpha_d + (1.0 - P.alpha_d) * pq;
            const double d = alpha_s * D[k] + (1.0 - alpha_s) * LY[k];
            D[k] = d; lambda_d[k] = d;
        }
    }
    __syncthreads();
    // ---- gains (emnr.c:905-963)
    {
        double *prev_gamma = st + L.prev_gamma, *prev_mask = st + L.prev_mask;
        for (int k = tid; k < ms; k += EM_T) {
            const double ly = LY[k], ld = lambda_d[k], pm = prev_mask[k];
            const double gamma = e_min(ly / ld, P.gamma_max);
            double eps_hat = P.alpha * pm * pm * prev_gamma[k] + (1.0 - P.alpha) * e_max(gamma - 1.0, P.eps_floor);
            double mask, pmask_out = 0.0;
            if (P.gain_method == 0) {
                eps_hat = e_max(eps_hat, P.xi_min);
                const double v = (eps_hat / (1.0 + eps_hat)) * gamma;
                mask = P.gf1p5 * sqrt(v) / gamma * exp(-0.5 * v) * ((1.0 + v) * e_bessI0(0.5 * v) + v * e_bessI1(0.5 * v));
                const double v2 = e_min(v, 700.0);
                const double eta = mask * mask * ly / ld;
                const double eps = eta / (1.0 - P.q);
                const double witchHat = (1.0 - P.q) / P.q * exp(v2) / (1.0 + eps);
                mask *= witchHat / (1.0 + witchHat);
                if (mask > P.gmax) mask = P.gmax;
                if (mask != mask) mask = 0.01;
            } else if (P.gain_method == 1) {
                const double ehr = eps_hat / (1.0 + eps_hat);
                const double v = ehr * gamma;
                mask = ehr * exp(e_min(700.0, 0.5 * e_e1xb(v)));
                if (mask > P.gmax) mask = P.gmax;
                if (mask != mask) mask = 0.01;
            } else if (P.gain_method == 2) {
                const double eps_p = eps_hat / (1.0 - P.q);
                mask = e_getKey(P.GG, gamma, eps_hat) * e_getKey(P.GGS, gamma, eps_p);
            } else {
                // method 3 (emnr.c:965-1010): method 0's gain, kept as prev_mask; the same law once more from the a-priori SNR
                // that gain implies (with the FIRST pass's v in the speech-presence factor, as the reference has it); then
                // the trained table decides 1 or 0 where it holds a value for this (gamma, xi) cell
                double xi_hat = e_max(eps_hat, P.xi_min);
                const double v = (xi_hat / (1.0 + xi_hat)) * gamma;
                mask = P.gf1p5 * sqrt(v) / gamma * exp(-0.5 * v) * ((1.0 + v) * e_bessI0(0.5 * v) + v * e_bessI1(0.5 * v));
                const double v2 = e_min(v, 700.0);
                {
                    const double eta = mask * mask * ly / ld;
                    const double eps = eta / (1.0 - P.q);
                    const double witchHat = (1.0 - P.q) / P.q * exp(v2) / (1.0 + eps);
                    mask *= witchHat / (1.0 + witchHat);
                }
                if (mask > P.gmax) mask = P.gmax;
                if (mask != mask) mask = 0.01;
                pmask_out = mask;
                {
                    double xi_ts = mask * mask * gamma;
                    xi_ts = e_max(xi_ts, P.xi_min);
                    const double v_ts = (xi_ts / (1.0 + xi_ts)) * gamma;
                    mask = P.gf1p5 * sqrt(v_ts) / gamma * exp(-0.5 * v_ts) * ((1.0 + v_ts) * e_bessI0(0.5 * v_ts) + v_ts * e_bessI1(0.5 * v_ts));
                    const double eta = mask * mask * ly / ld;
                    const double eps = eta / (1.0 - P.q);
                    const double witchHat = (1.0 - P.q) / P.q * exp(v2) / (1.0 + eps);
                    mask *= witchHat / (1.0 + witchHat);
                    xi_hat = xi_ts;
                }
                {   // getZeta, emnr.c:866-884 (its second range test compares xi in dB, not its cell index, with dim_zeta: kept)
                    const double gamma_dB = 10.0 * mlog10_dev(P.mtable, gamma), xi_dB = 10.0 * mlog10_dev(P.mtable, xi_hat);
                    const double gamma_per_cell = (P.z_gamma_max - P.z_gamma_min) / P.dim_zeta;
                    const double xi_per_cell = (P.z_xihat_max - P.z_xihat_min) / P.dim_zeta;
                    const int i_gamma = (int)floor((gamma_dB - P.z_gamma_min) / gamma_per_cell);
                    const int i_xi = (int)floor((xi_dB - P.z_xihat_min) / xi_per_cell);
                    if (!(i_gamma < 0 || i_gamma >= P.dim_zeta || i_xi < 0 || xi_dB >= P.dim_zeta)) {
                        const int index = i_gamma * P.dim_zeta + i_xi;
                        if (P.zeta_true[index] > 0) mask = P.zeta_hat[index] > P.zeta_thresh ? 1.0 : 0.0;
                    }
                }
            }
            MK[k] = mask;
            prev_gamma[k] = gamma; prev_mask[k] = P.gain_method == 3 ? pmask_out : mask;
        }
    }
    __syncthreads();
    // ---- aepf (emnr.c:769-816)
    if (P.ae_run) {
        if (tid == 0) {
            double sumPre = 0.0, sumPost = 0.0;
            for (int k = 0; k < ms; k++) { sumPre += LY[k]; sumPost += MK[k] * MK[k] * LY[k]; }
            red[4] = sumPost / sumPre;
        }
        __syncthreads();
        const double zeta = red[4];
        const double zetaT = zeta >= P.zetaThresh ? 1.0 : zeta;
        const int N = zetaT == 1.0 ? 1 : 1 + 2 * (int)(0.5 + P.psi * (1.0 - zetaT / P.zetaThresh));
        const int n = N / 2;
        for (int k = tid; k < ms; k += EM_T) {
            double a = 0.0;
            if (k < n) {
                for (int m = 0; m <= 2 * k; m++) a += MK[m];
                a /= (double)(2 * k + 1);
            } else if (k < ms - n) {
                for (int m = k - n; m <= k + n; m++) a += MK[m];
                a /= (double)N;
            } else {
                for (int m = ms - 1; m >= -ms + 2 * k + 1; m--) a += MK[m];
                a /= (double)(2 * (ms - k) - 1);
            }
            SCR[k] = a;
        }
        __syncthreads();
        const double damp = (P.gain_method == 3 && zetaT < P.t2) ? 0.05 : 1.0;         // emnr.c:813-815
        for (int k = tid; k < ms; k += EM_T) MK[k] = damp == 1.0 ? SCR[k] : SCR[k] * damp;
        __syncthreads();
    }
    // ---- back: g1 * Y on bins 0 .. fs/2, the Hermitian extension above, inverse transform (emnr.c:1038-1044)
    for (int k = tid; k < ms; k += EM_T) {
        const double g1 = P.gain * MK[k];
        const cd y = S[fsw(k)];
        cd v = make_double2(g1 * y.x, g1 * y.y);
        if (k == 0 || k == ms - 1) v.y = 0.0;               // a real-output inverse transform does not look at these two imaginary parts
        S[fsw(k)] = v;
    }
    __syncthreads();
    for (int k = tid + 1; k < ms - 1; k += EM_T) { const cd v = S[fsw(k)]; S[fsw(fs - k)] = make_double2(v.x, -v.y); }
    __syncthreads();
    fft_smem<1>(S, fs, twl, +1, tid, EM_T);
    __syncthreads();
    double *save = st + L.save;
    for (int i = tid; i < fs; i += EM_T) save[(size_t)saveidx * fs + i] = window[i] * S[fsw(i)].x;
    __syncthreads();
    // ---- overlap-add of the newest `ovrlp` frames into the output accumulator (emnr.c:1045-1056)
    double *outa = st + L.outaccum;
    for (int t = tid; t < P.incr; t += EM_T) {
        int ko = oainidx + t; if (ko >= P.oasize) ko -= P.oasize;
        double acc = 0.0;
        for (int i = P.ovrlp; i > 0; i--) {
            const int sbuff = (saveidx + i) % P.ovrlp, j = P.incr * (P.ovrlp - i) + t;
            const double v = save[(size_t)sbuff * fs + j];
            acc = i == P.ovrlp ? v : acc + v;
        }
        outa[ko] = acc;
    }
}

__global__ void emnr_in_kernel(const cd *in, long is, int n, double *state, EmnrLayout L, int iainidx, int iasize)
{
    double *ina = state + (size_t)blockIdx.x * L.row + L.inaccum;
    const cd *x = in + (size_t)blockIdx.x * is;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { int j = iainidx + i; if (j >= iasize) j -= iasize; ina[j] = x[i].x; }
}
__global__ void emnr_out_kernel(cd *out, long os, int n, const double *state, EmnrLayout L, int oaoutidx, int oasize)
{
    const double *outa = state + (size_t)blockIdx.x * L.row + L.outaccum;
    cd *y = out + (size_t)blockIdx.x * os;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { int j = oaoutidx + i; if (j >= oasize) j -= oasize; y[i] = make_double2(outa[j], 0.0); }
}

// the two gamma-prior tables, handed over once per process (host copies), uploaded per device on first use
static std::mutex g_tab_mu;
static std::vector<double> g_zeta; static std::vector<int> g_zvalid; static double g_zpar[4];
static double *g_dzeta[64] = {nullptr}; static int *g_dzvalid[64] = {nullptr};
static std::vector<double> g_GG, g_GGS;
static double *g_dGG[64] = {nullptr}, *g_dGGS[64] = {nullptr};

struct Emnr {
    int C = 0;
    EmnrPar P;
    EmnrLayout L;
    double *d_state = nullptr, *d_window = nullptr;
    const cd *tw = nullptr;
    int rate = 0, wintype = 0; double ogain = 1.0;
    // ring positions (the same for every channel)
    int iainidx = 0, iaoutidx = 0, oainidx = 0, init_oainidx = 0, oaoutidx = 0, nsamps = 0, saveidx = 0, subwc = 0, amb_idx = 0;

    static double interpM(double x, int nvals, const double *xv, const double *yv)
    {   // emnr.c:188-204
        if (x <= xv[0]) return yv[0];
        if (x >= xv[nvals - 1]) return yv[nvals - 1];
        int idx = 0;
        while (x >= xv[idx]) idx++;
        const double xllow = log10(xv[idx - 1]), xlhigh = log10(xv[idx]);
        const double frac = (log10(x) - xllow) / (xlhigh - xllow);
        return yv[idx - 1] + frac * (yv[idx] - yv[idx - 1]);
    }

    int init(int C_, int bsize, int fsize, int ovrlp, int rate_, int wintype_, double gain, int gain_method, int npe_method, int ae_run)
    {   // create_emnr + calc_emnr, emnr.c:242-503, 561-581
        C = C_; rate = rate_; wintype = wintype_; ogain = gain;
        if (C <= 0 || bsize <= 0 || fsize != 4096 || ovrlp < 1 || fsize % ovrlp || wintype != 0) {
            set_error("emnr_create: fsize must be 4096 (the value create_rxa uses), ovrlp a divisor of it, wintype 0");
            return QC_EINVAL;
        }
        memset(&P, 0, sizeof(P));
        P.fs = fsize; P.ms = fsize / 2 + 1; P.ovrlp = ovrlp; P.incr = fsize / ovrlp; P.bsize = bsize;
        P.gain = gain / fsize / (double)ovrlp;
        P.iasize = fsize > bsize ? fsize : bsize + fsize - P.incr;
        if (fsize > bsize) { P.oasize = bsize > P.incr ? bsize : P.incr; oainidx = (fsize - bsize - P.incr) % P.oasize; }
        else { P.oasize = bsize; oainidx = fsize - P.incr; }
        init_oainidx = oainidx;
        P.gain_method = gain_method; P.npe_method = npe_method; P.ae_run = ae_run;
        const double rt = (double)rate, inc = (double)P.incr;
        P.gf1p5 = sqrt(M_PI) / 2.0;
        { const double tau = -128.0 / 8000.0 / log(0.985); P.alpha = exp(-inc / rt / tau); }
        P.eps_floor = 1.0e-300; P.gamma_max = 40.0; P.xi_min = pow(10.0, -40.0 / 10.0); P.q = 0.2; P.gmax = 10000.0;
        { const double tau = -128.0 / 8000.0 / log(0.7); P.alphaCsmooth = exp(-inc / rt / tau); }
        { const double tau = -128.0 / 8000.0 / log(0.96); P.alphaMax = exp(-inc / rt / tau); }
        { const double tau = -128.0 / 8000.0 / log(0.7); P.alphaCmin = exp(-inc / rt / tau); }
        { const double tau = -128.0 / 8000.0 / log(0.3); P.alphaMin_max_value = exp(-inc / rt / tau); }
        P.snrq = -inc / (0.064 * rt);
        { const double tau = -128.0 / 8000.0 / log(0.8); P.betamax = exp(-inc / rt / tau); }
        P.invQeqMax = 0.5; P.av = 2.12;
        const double Dtime = 8.0 * 12.0 * 128.0 / 8000.0;
        P.U = 8;
        P.V = (int)(0.5 + (Dtime * rt / (P.U * inc)));
        if (P.V < 4) P.V = 4;
        if ((P.U = (int)(0.5 + (Dtime * rt / (P.V * inc)))) < 1) P.U = 1;
        P.D = P.U * P.V;
        static const double Dvals[18] = {1.0, 2.0, 5.0, 8.0, 10.0, 15.0, 20.0, 30.0, 40.0, 60.0, 80.0, 120.0, 140.0, 160.0, 180.0, 220.0, 260.0, 300.0};
        static const double Mvals[18] = {0.000, 0.260, 0.480, 0.580, 0.610, 0.668, 0.705, 0.762, 0.800, 0.841, 0.865, 0.890, 0.900, 0.910, 0.920, 0.930, 0.935, 0.940};
        P.MofD = interpM((double)P.D, 18, Dvals, Mvals);
        P.MofV = interpM((double)P.V, 18, Dvals, Mvals);
        P.invQbar_points[0] = 0.03; P.invQbar_points[1] = 0.05; P.invQbar_points[2] = 0.06; P.invQbar_points[3] = 1.0e300;
        const double facs[4] = {8.0, 4.0, 2.0, 1.2};
        for (int i = 0; i < 4; i++) { const double db = 10.0 * log10(facs[i]) / (12.0 * 128 / 8000); P.nsmax[i] = pow(10.0, db / 10.0 * P.V * inc / rt); }
        { const double tau = -128.0 / 8000.0 / log(0.8); P.alpha_pow = exp(-inc / rt / tau); }
        { const double tau = -128.0 / 8000.0 / log(0.9); P.alpha_Pbar = exp(-inc / rt / tau); }
        P.epsH1 = pow(10.0, 15.0 / 10.0); P.epsH1r = P.epsH1 / (1.0 + P.epsH1);
        { const double tau = -256.0 / (20100.0 * log(0.7)); P.eta = exp(-inc / (rt * tau)); }
        { const double tau = -256.0 / (20100.0 * log(0.998)); P.gamma_l = exp(-inc / (rt * tau)); }
        { const double tau = -256.0 / (20100.0 * log(0.8)); P.beta_l = exp(-inc / (rt * tau)); }
        { const double tau = -256.0 / (20100.0 * log(0.85)); P.alpha_d = exp(-inc / (rt * tau)); }
        { const double tau = -256.0 / (20100.0 * log(0.2)); P.alpha_p = exp(-inc / (rt * tau)); }
        P.delta_LF = 1000.0 / (rt / 2) * P.ms; P.delta_MF = 3000.0 / (rt / 2) * P.ms;
        P.delta_0 = 2.0; P.delta_1 = 2.0; P.delta_2 = 5.0;
        P.zetaThresh = 0.75; P.psi = 20.0; P.t2 = 0.20;
        P.zeta_thresh = -2.0;           // g.zeta_thresh, emnr.c:332
        // state row
        size_t o = 0;
        const size_t ms = (size_t)P.ms;
        auto take = [&](size_t n) { const size_t at = o; o += n; return at; };
        L.inaccum = take(P.iasize); L.outaccum = take(P.oasize); L.save = take((size_t)ovrlp * fsize);
        L.lambda_d = take(ms); L.prev_gamma = take(ms); L.prev_mask = take(ms);
        L.p = take(ms); L.sigma2N = take(ms); L.pbar = take(ms); L.p2bar = take(ms); L.actmin = take(ms); L.actmin_sub = take(ms);
        L.lmin_flag = take(ms); L.pmin_u = take(ms); L.actminbuff = take((size_t)P.U * ms); L.alphaC = take(1);
        L.s_sigma2N = take(ms); L.s_Pbar = take(ms); L.l_P = take(ms); L.l_Pmin = take(ms); L.l_p = take(ms); L.l_D = take(ms);
        L.row = (o + 1) & ~(size_t)1;
        QC_CUDA(cudaMalloc((void **)&d_state, (size_t)C * L.row * sizeof(double)));
        // window (emnr.c:167-186): sqrt-Hamming scaled to unit coherent gain
        std::vector<double> w((size_t)fsize);
        const double arg = 2.0 * M_PI / (double)fsize;
        double sum = 0.0;
        for (int i = 0; i < fsize; i++) { w[i] = sqrt(0.54 - 0.46 * cos((double)i * arg)); sum += w[i]; }
        const double icg = (double)fsize / sum;
        for (int i = 0; i < fsize; i++) w[i] *= icg;
        QC_CUDA(cudaMalloc((void **)&d_window, (size_t)fsize * sizeof(double)));
        QC_CUDA(cudaMemcpy(d_window, w.data(), (size_t)fsize * sizeof(double), cudaMemcpyHostToDevice));
        tw = fft_twiddles(fsize);
        if (!tw) { set_error("emnr_create: twiddle table allocation failed"); return QC_ENOMEM; }
        return reset_all();
    }
    int reset_all()
    {   // the initial values calc_emnr leaves (emnr.c:288-291, 388-407, 427-431)
        std::vector<double> row(L.row, 0.0);
        const size_t ms = (size_t)P.ms;
        for (size_t k = 0; k < ms; k++) {
            row[L.prev_gamma + k] = 1.0; row[L.prev_mask + k] = 1.0;
            row[L.p + k] = 0.5; row[L.sigma2N + k] = 0.5; row[L.pbar + k] = 0.5; row[L.pmin_u + k] = 0.5; row[L.p2bar + k] = 0.25;
            row[L.actmin + k] = 1.0e300; row[L.actmin_sub + k] = 1.0e300;
            for (int ku = 0; ku < P.U; ku++) row[L.actminbuff + (size_t)ku * ms + k] = 1.0e300;
            row[L.s_sigma2N + k] = 0.5; row[L.s_Pbar + k] = 0.5;
        }
        row[L.alphaC] = 1.0;
        for (int c = 0; c < C; c++) QC_CUDA(cudaMemcpy(d_state + (size_t)c * L.row, row.data(), L.row * sizeof(double), cudaMemcpyHostToDevice));
        subwc = P.V; amb_idx = 0;
        iainidx = iaoutidx = oaoutidx = nsamps = saveidx = 0; oainidx = init_oainidx;
        return QC_OK;
    }
    int flush()
    {   // flush_emnr, emnr.c:583-596: the sample accumulators and the saved frames only -- the noise estimate survives
        QC_CUDA(cudaDeviceSynchronize());
        for (int c = 0; c < C; c++) {
            double *st = d_state + (size_t)c * L.row;
            QC_CUDA(cudaMemset(st + L.inaccum, 0, (size_t)P.iasize * sizeof(double)));
            QC_CUDA(cudaMemset(st + L.outaccum, 0, (size_t)P.oasize * sizeof(double)));
            QC_CUDA(cudaMemset(st + L.save, 0, (size_t)P.ovrlp * P.fs * sizeof(double)));
        }
        nsamps = iainidx = iaoutidx = oaoutidx = saveidx = 0; oainidx = init_oainidx;
        return QC_OK;
    }
    void release() { if (d_state) cudaFree(d_state); if (d_window) cudaFree(d_window); d_state = d_window = nullptr; }

    int tables()
    {
        if (P.gain_method == 3) {
            std::lock_guard<std::mutex> g(g_tab_mu);
            if (g_zeta.empty()) { set_error("emnr: gain method 3 needs the trained zeta table of the WDSP distribution (wdsp/zetahat.c or its `zetaHat` file, emnr.c:206-238): hand it over with quisk_cuda_emnr_set_zeta first"); return QC_EINVAL; }
            int dev = 0; cudaGetDevice(&dev); dev &= 63;
            if (!g_dzeta[dev]) {
                QC_CUDA(cudaMalloc((void **)&g_dzeta[dev], g_zeta.size() * sizeof(double)));
                QC_CUDA(cudaMalloc((void **)&g_dzvalid[dev], g_zvalid.size() * sizeof(int)));
                QC_CUDA(cudaMemcpy(g_dzeta[dev], g_zeta.data(), g_zeta.size() * sizeof(double), cudaMemcpyHostToDevice));
                QC_CUDA(cudaMemcpy(g_dzvalid[dev], g_zvalid.data(), g_zvalid.size() * sizeof(int), cudaMemcpyHostToDevice));
            }
            P.zeta_hat = g_dzeta[dev]; P.zeta_true = g_dzvalid[dev]; P.dim_zeta = 60;
            P.z_gamma_min = g_zpar[0]; P.z_gamma_max = g_zpar[1]; P.z_xihat_min = g_zpar[2]; P.z_xihat_max = g_zpar[3];
            P.mtable = mlog10_table();
            if (!P.mtable) return QC_ECUDA;
            return QC_OK;
        }
        if (P.gain_method != 2) return QC_OK;
        std::lock_guard<std::mutex> g(g_tab_mu);
        if (g_GG.empty()) { set_error("emnr: gain method 2 needs the two 241 x 241 tables of the WDSP distribution (wdsp/calculus.c or its `calculus` file): hand them over with quisk_cuda_emnr_set_tables first"); return QC_EINVAL; }
        int dev = 0; cudaGetDevice(&dev); dev &= 63;
        if (!g_dGG[dev]) {
            QC_CUDA(cudaMalloc((void **)&g_dGG[dev], g_GG.size() * sizeof(double)));
            QC_CUDA(cudaMalloc((void **)&g_dGGS[dev], g_GGS.size() * sizeof(double)));
            QC_CUDA(cudaMemcpy(g_dGG[dev], g_GG.data(), g_GG.size() * sizeof(double), cudaMemcpyHostToDevice));
            QC_CUDA(cudaMemcpy(g_dGGS[dev], g_GGS.data(), g_GGS.size() * sizeof(double), cudaMemcpyHostToDevice));
        }
        P.GG = g_dGG[dev]; P.GGS = g_dGGS[dev];
        return QC_OK;
    }

    int run(const cd *d_in, long is, cd *d_out, long os, cudaStream_t s)
    {   // xemnr with run = 1, emnr.c:1015-1064
        if (P.gain_method < 0 || P.gain_method > 3) { set_error("emnr: there is no gain method %d (0 .. 3)", P.gain_method); return QC_EINVAL; }
        int rc = tables(); if (rc != QC_OK) return rc;
        emnr_in_kernel<<<C, 256, 0, s>>>(d_in, is, P.bsize, d_state, L, iainidx, P.iasize);
        count_launch();
        iainidx = (iainidx + P.bsize) % P.iasize;
        nsamps += P.bsize;
        const size_t sh = ((size_t)fft_tw_entries(P.fs) + P.fs) * sizeof(cd) + ((size_t)2 * P.ms + 16) * sizeof(double);
        QC_CUDA(cudaFuncSetAttribute(emnr_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        while (nsamps >= P.fs) {
            emnr_frame_kernel<<<C, EM_T, sh, s>>>(P, L, d_state, d_window, tw, iaoutidx, saveidx, oainidx, subwc, amb_idx);
            count_launch();
            iaoutidx = (iaoutidx + P.incr) % P.iasize;
            nsamps -= P.incr;
            saveidx = (saveidx + 1) % P.ovrlp;
            oainidx = (oainidx + P.incr) % P.oasize;
            if (P.npe_method == 0) {
                if (subwc == P.V) { if (++amb_idx == P.U) amb_idx = 0; subwc = 1; }
                else ++subwc;
            }
        }
        emnr_out_kernel<<<C, 256, 0, s>>>(d_out, os, P.bsize, d_state, L, oaoutidx, P.oasize);
        count_launch();
        QC_CUDA_LAUNCH();
        oaoutidx = (oaoutidx + P.bsize) % P.oasize;
        return QC_OK;
    }
};

Emnr *make_emnr(int C, int bsize, int fsize, int ovrlp, int rate, int wintype, double gain, int gain_method, int npe_method, int ae_run)
{
    Emnr *e = new Emnr();
    if (e->init(C, bsize, fsize, ovrlp, rate, wintype, gain, gain_method, npe_method, ae_run) != QC_OK) { e->release(); delete e; return nullptr; }
    return e;
}
void emnr_destroy(Emnr *e) { if (e) { e->release(); delete e; } }
int emnr_run(Emnr *e, const cd *in, long is, cd *out, long os, cudaStream_t s) { return e->run(in, is, out, os, s); }
int emnr_flush(Emnr *e) { return e->flush(); }
int emnr_set(Emnr *e, int what, int value)
{
    switch (what) {
    case 0: e->P.gain_method = value; return QC_OK;         // SetRXAEMNRgainMethod: a plain assignment in the reference, emnr.c:1111-1117
    case 1: e->P.npe_method = value; return QC_OK;          // SetRXAEMNRnpeMethod
    case 2: e->P.ae_run = value ? 1 : 0; return QC_OK;      // SetRXAEMNRaeRun
    }
    return QC_EINVAL;
}
bool emnr_tables_present() { std::lock_guard<std::mutex> g(g_tab_mu); return !g_GG.empty(); }
bool emnr_zeta_present() { std::lock_guard<std::mutex> g(g_tab_mu); return !g_zeta.empty(); }
int emnr_set_train(Emnr *e, int what, double value)
{   // SetRXAEMNRtrainZetaThresh (emnr.c:1160-1166), SetRXAEMNRtrainT2 (emnr.c:1168-1174)
    if (what == 0) e->P.zeta_thresh = value; else e->P.t2 = value;
    return QC_OK;
}

}  // namespace qc

struct qcEmnr { qc::Emnr *e; };

extern "C" {

int quisk_cuda_emnr_set_tables(const double *GG, const double *GGS)
{
    if (!GG || !GGS) { qc::set_error("emnr_set_tables: two tables of 241 x 241 doubles"); return QC_EINVAL; }
    std::lock_guard<std::mutex> g(qc::g_tab_mu);
    qc::g_GG.assign(GG, GG + 241 * 241);
    qc::g_GGS.assign(GGS, GGS + 241 * 241);
    for (int d = 0; d < 64; d++) {      // re-upload on next use
        if (qc::g_dGG[d]) { cudaFree(qc::g_dGG[d]); qc::g_dGG[d] = nullptr; }
        if (qc::g_dGGS[d]) { cudaFree(qc::g_dGGS[d]); qc::g_dGGS[d] = nullptr; }
    }
    return QC_OK;
}

int quisk_cuda_emnr_set_zeta(const double *zeta_hat, const int *zeta_valid, int rows, int cols, double gamma_min, double gamma_max, double xihat_min, double xihat_max)
{
    if (!zeta_hat || !zeta_valid || rows != 60 || cols != 60) { qc::set_error("emnr_set_zeta: a 60 x 60 table (dim_zeta, emnr.c:329) and its validity map"); return QC_EINVAL; }
    std::lock_guard<std::mutex> g(qc::g_tab_mu);
    qc::g_zeta.assign(zeta_hat, zeta_hat + 3600);
    qc::g_zvalid.assign(zeta_valid, zeta_valid + 3600);
    qc::g_zpar[0] = gamma_min; qc::g_zpar[1] = gamma_max; qc::g_zpar[2] = xihat_min; qc::g_zpar[3] = xihat_max;
    for (int d = 0; d < 64; d++) {
        if (qc::g_dzeta[d]) { cudaFree(qc::g_dzeta[d]); qc::g_dzeta[d] = nullptr; }
        if (qc::g_dzvalid[d]) { cudaFree(qc::g_dzvalid[d]); qc::g_dzvalid[d] = nullptr; }
    }
    return QC_OK;
}
int quisk_cuda_emnr_set_train(qcEmnr *h, double zeta_thresh, double t2);

qcEmnr *quisk_cuda_emnr_create(int n_channels, int bsize, int fsize, int ovrlp, int rate, int wintype, double gain, int gain_method, int npe_method, int ae_run)
{
    if (qc::ensure_device() != QC_OK) return nullptr;
    qc::Emnr *e = qc::make_emnr(n_channels, bsize, fsize, ovrlp, rate, wintype, gain, gain_method, npe_method, ae_run);
    if (!e) return nullptr;
    qcEmnr *h = new qcEmnr();
    h->e = e;
    return h;
}
void quisk_cuda_emnr_destroy(qcEmnr *h) { if (h) { qc::emnr_destroy(h->e); delete h; } }
int quisk_cuda_emnr_run(qcEmnr *h, const void *d_in, long in_stride, void *d_out, long out_stride, void *stream)
{
    if (!h || !d_in || !d_out) { qc::set_error("emnr_run: bad arguments"); return QC_EINVAL; }
    return h->e->run((const double2 *)d_in, in_stride, (double2 *)d_out, out_stride, (cudaStream_t)stream);
}
int quisk_cuda_emnr_flush(qcEmnr *h) { return h ? h->e->flush() : QC_EINVAL; }
int quisk_cuda_emnr_set_gain_method(qcEmnr *h, int method) { return h ? qc::emnr_set(h->e, 0, method) : QC_EINVAL; }
int quisk_cuda_emnr_set_npe_method(qcEmnr *h, int method) { return h ? qc::emnr_set(h->e, 1, method) : QC_EINVAL; }
int quisk_cuda_emnr_set_ae_run(qcEmnr *h, int run) { return h ? qc::emnr_set(h->e, 2, run) : QC_EINVAL; }
int quisk_cuda_emnr_set_train(qcEmnr *h, double zeta_thresh, double t2)
{
    if (!h) return QC_EINVAL;
    qc::emnr_set_train(h->e, 0, zeta_thresh); qc::emnr_set_train(h->e, 1, t2);
    return QC_OK;
}

}  // extern "C"
