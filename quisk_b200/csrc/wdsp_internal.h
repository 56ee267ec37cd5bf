// quisk_b200/csrc/wdsp_internal.h -- internal classes of the WDSP RXA part (see wdsp_*.cu)
#pragma once
#include "qc_common.cuh"
#include "batch.h"
#include "../../include/quisk_cuda_wdsp.h"

namespace qc {

struct FirCore {
    int C = 0, size = 0, nc = 0, nfor = 0, n2 = 0;
    int buffidx = 0, cset = 0, masks_ready = 0, mp = 0;
    const cd *tw = nullptr;
    cd *d_prev = nullptr, *d_fdl = nullptr, *d_gen = nullptr;
    cd *d_mask[2] = {nullptr, nullptr};
    int init(int C, int size, int nc, int mp, const double *impulse);
    void release();
    int flush();
    int set_impulse(const double *impulse, int update);
    int set_mp(int mp);                     // setMp_fircore, firmin.c:469-473
    std::vector<double> h_impulse;          // the impulse as handed in (a->impulse), for set_mp
    int update();
    int run(const void *d_in, long in_stride, void *d_out, long out_stride, cudaStream_t s);
};

struct Resampler {
    int C = 0, L = 1, M = 1, ncoef = 0;
    BatchFilter *f = nullptr;
    int init(int C, int in_rate, int out_rate, double fc, int ncoef_in, double gain);
    void release();
};

// Sequential per-channel stages (wdsp_seq.cu): parameters are uniform across the batch, state is per channel.
enum SeqKind { SEQ_SHIFT = 1, SEQ_WCPAGC = 2, SEQ_AMD = 3, SEQ_FMPLL = 4, SEQ_SNOTCH = 5, SEQ_METER = 6 };

struct AgcParams {      // the fields of struct _wcpagc that xwcpagc reads (wdsp/wcpAGC.h)
    int mode, pmode, ring_buffsize, attack_buffsize, hang_enable;
    double sample_rate, fixed_gain, attack_mult, decay_mult, fast_decay_mult, fast_backmult, onemfast_backmult,
           out_target, min_volts, inv_out_target, slope_constant, inv_max_input, hang_level, hang_backmult,
           onemhang_backmult, hang_decay_mult, pop_ratio, hangtime;
    // create-time inputs kept for loadWcpAGC
    double tau_attack, tau_decay, max_gain, var_gain, max_input, out_targ, tau_fast_backaverage, tau_fast_decay,
           tau_hang_backmult, hang_thresh, tau_hang_decay;
    int n_tau;
};

struct SeqStage {
    int kind = 0, C = 0;
    double *d_state = nullptr;      // [C][state_doubles]
    int state_doubles = 0;
    double *d_ring = nullptr;       // wcpagc: [C][ring_len][3] (re, im, abs)
    int ring_len = 0;
    double *d_par = nullptr;        // per-channel parameters (shift: delta, cos_delta, sin_delta)
    double par[32] = {0};           // uniform parameters
    AgcParams agc;
    double *d_meter = nullptr;      // meter results [C][3]
    int init_common(int kind, int C, int state_doubles);
    void release();
    int flush();
    int run(const void *d_in, long in_stride, void *d_out, long out_stride, int n, cudaStream_t s);
    void load_agc();
};

SeqStage *make_shift(int C, int rate, const double *shift_hz);
SeqStage *make_wcpagc(int C, int rate, int mode);
SeqStage *make_amd(int C, int rate, int mode, int levelfade, int sbmode);
SeqStage *make_fmpll(int C, int rate, double deviation, double fmin, double fmax, double zeta, double omegaN, double tau);
SeqStage *make_snotch(int C, int rate, double f, double bw);
SeqStage *make_meter(int C, int rate, double tau_av, double tau_decay);
void agc_set_mode(SeqStage *s, int mode);

int launch_panel(const cd *in, long in_stride, cd *out, long out_stride, int n, int C, double gainI, double gainQ,
                 int inselect, int copy, cudaStream_t s, cd *sip = nullptr, int sipsize = 0, int sip_idx = 0);

}  // namespace qc
