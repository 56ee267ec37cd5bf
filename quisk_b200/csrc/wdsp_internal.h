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
    int set_gen(const cd *gen, int update);      // masks from nfor x 2*size time-domain generator rows
    int set_mp(int mp);                     // setMp_fircore, firmin.c:469-473
    std::vector<double> h_impulse;          // the impulse as handed in (a->impulse), for set_mp
    int update();
    int run(const void *d_in, long in_stride, void *d_out, long out_stride, cudaStream_t s);
};

struct Resampler {
    int C = 0, L = 1, M = 1, ncoef = 0;
    BatchFilter *f = nullptr;
    int init(int C, int in_rate, int out_rate, double fc, int ncoef_in, double gain);
    int init_band(int C, int in_rate, int out_rate, double fc_low, double fc, int ncoef_in, double gain);   // fc_low >= 0: setFCLow_resample's band pass
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

// spectral noise blanker (wdsp_snba_nofma.cu)
struct Snba;
Snba *make_snba(int C, int inrate, int internalrate, int bsize, int ovrlp, int xsize, int asize, int npasses, double k1, double k2, int b,
                int pre, int post, double pmultmin, double out_low, double out_high);
void snba_destroy(Snba *d);
int snba_run(Snba *d, const cd *in, long is, cd *out, long os, cudaStream_t s);
int snba_flush(Snba *d);
int snba_set_output_bandwidth(Snba *d, double flow, double fhigh);

// spectral noise reduction (wdsp_emnr_nofma.cu)
struct Emnr;
Emnr *make_emnr(int C, int bsize, int fsize, int ovrlp, int rate, int wintype, double gain, int gain_method, int npe_method, int ae_run);
void emnr_destroy(Emnr *e);
int emnr_run(Emnr *e, const cd *in, long is, cd *out, long os, cudaStream_t s);
int emnr_flush(Emnr *e);
int emnr_set(Emnr *e, int what /* 0 gain method, 1 npe method, 2 ae_run */, int value);
bool emnr_tables_present();
bool emnr_zeta_present();
int emnr_set_train(Emnr *e, int what /* 0 zeta threshold, 1 t2 */, double value);

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
    int meter_sub = 1;              // meters: how many DSP blocks one run() call covers (peak hold is per block, meter.c:95)
    int init_common(int kind, int C, int state_doubles);
    void release();
    int flush();
    int flush_ref();                // what the reference's own flush_<stage> resets (less than a fresh object for wcpagc / amd)
    int run(const void *d_in, long in_stride, void *d_out, long out_stride, int n, cudaStream_t s);
    void load_agc();
};

// mlog10 (wdsp/meterlog10.c:547-554): the table look-up logarithm WDSP's meters use; the table lives in device memory
const double *mlog10_table();
#ifdef __CUDACC__
__device__ __forceinline__ double mlog10_dev(const double *__restrict__ mtable, double val)
{
    const unsigned long long N = (unsigned long long)__double_as_longlong(val);
    const int e = (int)((N >> 52) & 2047ull) - 1023;
    const int m = (int)((N >> 41) & 2047ull);
    return 0.301029995663981 * ((double)e + mtable[m]);
}
#endif

SeqStage *make_shift(int C, int rate, const double *shift_hz);
SeqStage *make_wcpagc(int C, int rate, int mode);
SeqStage *make_wcpagc_fmlim(int C, int rate, double lim_gain);      // fmd's detector limiter (fmd.c:49-73)
SeqStage *make_amd(int C, int rate, int mode, int levelfade, int sbmode);
SeqStage *make_fmpll(int C, int rate, double deviation, double fmin, double fmax, double zeta, double omegaN, double tau);
SeqStage *make_snotch(int C, int rate, double f, double bw);
SeqStage *make_meter(int C, int rate, double tau_av, double tau_decay);
void agc_set_mode(SeqStage *s, int mode);

// The RXA chain for a batch of channels (wdsp_rxa.cu); wdsp_compat.cu maps WDSP's channel numbers onto these.
struct Rxa {
    int C = 0, in_size = 0, dsp_size = 0, in_rate = 0, dsp_rate = 0, out_rate = 0;
    int dsp_insize = 0, dsp_outsize = 0, out_size = 0;
    int mode = QC_RXA_LSB;
    // shift
    int shift_run = 1; bool shift_nonzero = false; SeqStage *shift = nullptr;
    Resampler *rsmpin = nullptr, *rsmpout = nullptr;
    SeqStage *adcmeter = nullptr, *smeter = nullptr, *agcmeter = nullptr;
    // nbp0
    int nbp_run = 1, nbp_nc = 0; double nbp_flow = -4150.0, nbp_fhigh = -150.0; FirCore *nbp0 = nullptr;
    // notch database (create_notchdb, RXA.c:85-87; nbp.c:34-47): shared by the batch like every other setting
    int ndb_run = 0; double ndb_tune = 0.0, ndb_shift = 0.0; int nbp_hadnotch = 0;
    std::vector<double> ndb_fcenter, ndb_fwidth; std::vector<int> ndb_active;
    int nbp0_impulse(std::vector<double> &imp, int *havnotch);
    // amd / fmd
    int amd_run = 0, amd_mode = 0; SeqStage *amd = nullptr;
    int fmd_run = 0, fm_nc_de = 0, fm_nc_aud = 0; SeqStage *fmpll = nullptr, *sntch = nullptr; FirCore *pde = nullptr, *paud = nullptr;
    // xsnba and the band pass that goes with it (xbpsnbain / xbpsnbaout, snb.c:700-827; RXAbpsnbaCheck / RXAbpsnbaSet, RXA.c:829-918)
    int snba_run = 0; Snba *snba = nullptr;
    FirCore *bpsnba = nullptr; int bps_nc = 0, bps_run = 0, bps_position = 0, bps_run_notches = 0, bps_hadnotch = 0;
    double bps_flow = -5700.0, bps_fhigh = -250.0; cd *bps_buff = nullptr;
    int bpsnba_impulse(std::vector<double> &imp, int *havnotch);
    int make_bpsnba();
    int bpsnba_check(int mode, int notch_run);
    int bpsnba_set();
    int emnr_run = 0, emnr_position = 0, emnr_gain_method = 2; Emnr *emnr = nullptr;      // xemnr, RXA.c:319-332, 577-590
    int emnr_run_stage(cd *m, long ms, cudaStream_t s);
    int snba_run_stage(cd *m, long ms, cudaStream_t s);
    int bp1_check_set();                                                                    // RXAbp1Check + RXAbp1Set, RXA.c:800-827
    int lim_run = 0; double lim_pre_gain = 0.4, lim_gain = 2.5; SeqStage *plim = nullptr;     // fmd's detector limiter (fmd.c:106-108, 179-184)
    int fm_limiter(cd *m, long ms, int n, cudaStream_t s);
    // bp1
    int bp1_run = 1, bp1_nc = 0; double bp1_flow = -4150.0, bp1_fhigh = -150.0, bp1_gain = 1.0; FirCore *bp1 = nullptr;
    // agc, panel
    int agc_run = 1; SeqStage *agc = nullptr;
    double panel_gain1 = 4.0, panel_gain2I = 1.0, panel_gain2Q = 1.0;
    // buffers
    cd *mid = nullptr, *mid2 = nullptr, *audio = nullptr;
    // fexchange0 emulation: up-slew state per channel (iobuffs.c:47-160): [C][3] = ustate, ucount, upflag
    int ndelup = 0, ntup = 0; int *d_uslew = nullptr; double *d_cup = nullptr;
    // sip1 (create_rxa, RXA.c:392-401: run 1, position 0, mode 0, 4096 samples): ring per channel, filled by the panel kernel
    int sip_run = 1, sipsize = 4096, sip_idx = 0; cd *d_sip = nullptr; float *d_sipout = nullptr; int sipout_cap = 0;
    int arm_upslew(double tdelayup, double tslewup);
    // fexchange0 / dexchange emulation (iobuffs.c:385-604): the two pseudo-rings r1 (caller -> DSP) and r2 (DSP -> caller)
    // live on the device, [C][DSP_MULT * size]; the ring arithmetic (indices, unqueued counts, the Sem_OutReady credit
    // count) is the reference's, run on the host -- the DSP "thread" is executed synchronously inside the exchange call,
    // which is the schedule the reference follows when its caller is paced by a sound card.
    int bfo = 1, exchange_on = 1, state = 1;
    double tdelayup = 0.0, tslewup = 0.0, tdelaydown = 0.0, tslewdown = 0.0;
    int r1_size = 0, r2_size = 0, r1_active = 0, r2_active = 0;
    int r1_inidx = 0, r1_outidx = 0, r1_unq = 0, r2_inidx = 0, r2_outidx = 0, r2_havesamps = 0, r2_unq = 0, out_credits = 0;
    cd *d_r1 = nullptr, *d_r2 = nullptr, *d_outbuff = nullptr, *d_xout = nullptr; double *d_gain = nullptr;
    cudaStream_t hs = nullptr;
    // down-slew (downslew0, iobuffs.c:226-300): data independent, so the state machine runs on the host and hands the
    // kernel one gain per output sample
    int downflag = 0, flushflag = 0, dstate = 0, dcount = 0, ndeldown = 0, ntdown = 0;
    std::vector<double> cdown;
    int setup_exchange();
    int flush_iobuffs();
    int flush_main();                        // flush_rxa (RXA.c:527-559) with the reference's (partial) per-stage flushes
    int exchange(const double *h_in, double *h_out, int *error);
    int set_channel_state(int new_state, int dmode);

    int init(int C, int in_size, int dsp_size, int in_rate, int dsp_rate, int out_rate);
    void release();
    int make_nbp0();
    int make_bp1();
    int make_fmd();
    int xrxa(const void *din, long is, void *dout, long os, cudaStream_t s);
    // wdsp_rxa_fused.cu: the whole chain as one kernel for the configurations it covers, any number of DSP blocks per launch
    int fused_ok = 1;                       // QC_RXA_OPT_FUSED
    double *d_wide_seq = nullptr; size_t wide_seq_cap = 0;                                           // pre-kernel outputs of the pipelined wide path
    cd *d_wide_spec = nullptr, *d_wide_y = nullptr; size_t wide_spec_cap = 0, wide_y_cap = 0;      // scratch of the many-blocks-per-launch path
    bool fusable() const;
    int xrxa_fused(const void *din, long is, void *dout, long os, int nblocks, cudaStream_t s);
    int xrxa_multi(const void *din, long is, void *dout, long os, int nblocks, cudaStream_t s);
    // every stage over a group of blocks per launch, for the configurations the single kernel does not cover (FM, AM, resamplers)
    int xrxa_stages_wide(const void *din, long is, void *dout, long os, int g, cudaStream_t s);
    cd *wmid = nullptr, *wmid2 = nullptr, *waudio = nullptr; long wstride = 0;
    cudaStream_t side = nullptr; cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr;     // meters next to the chain
};

// one fircore over nblocks consecutive blocks of every channel, transforms of all blocks in parallel (wdsp_rxa_fused.cu);
// spec: scratch of C * nblocks * 2 * size complex; in may equal out
int fircore_wide(FirCore *f, const cd *in, long in_stride, cd *out, long out_stride, int nblocks, cd *spec, cudaStream_t s);

int launch_panel(const cd *in, long in_stride, cd *out, long out_stride, int n, int C, double gainI, double gainQ,
                 int inselect, int copy, cudaStream_t s, cd *sip = nullptr, int sipsize = 0, int sip_idx = 0);

}  // namespace qc

struct qcRxa { qc::Rxa r; };       // the opaque handle of include/quisk_cuda_wdsp.h
