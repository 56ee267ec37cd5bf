// quisk_b200/csrc/wdsp_compat.cu -- the reference's own WDSP entry points, by channel number, on top of the batched
// RXA chain (wdsp_rxa.cu).  These are the symbols quisk_wdsp.py resolves in libwdsp.so with ctypes (quisk_wdsp.py:28-99,
// quisk.py:6017-6053) and the fexchange0 pointer it hands to quisk_wdsp.c (quisk_wdsp.py:59-67), with the reference's
// exact signatures:
//   OpenChannel / CloseChannel / SetChannelState   wdsp/channel.c:76-104, 121-126, 262-300
//   fexchange0                                     wdsp/iobuffs.c:464-516
//   GetWDSPVersion                                 wdsp/version.c:4-10
//   SetRXAMode, RXASetPassband, RXASetNC, RXASetMP wdsp/RXA.c:749-787, 927-958
//   SetRXAShiftRun / Freq                          wdsp/shift.c:112-128
//   RXANBPSetRun + the notch database calls        wdsp/nbp.c:359-527
//   SetRXAAGCMode / Fixed / Top                    wdsp/wcpAGC.c:370-548
//   SetRXAPanelRun / Gain1 / Gain2                 wdsp/patchpanel.c:125-156
//   SetRXAAMDSBMode / FadeLevel                    wdsp/amd.c:279-293
//   GetRXAMeter                                    wdsp/meter.c:120-129
//   RXAGetaSipF / RXAGetaSipF1                     wdsp/siphon.c:183-211
// plus wdspFexchange0, the re-blocker Quisk's own C side puts in front of fexchange0 (quisk_wdsp.c:24-73).
// Each open channel number owns one single-channel chain; many receivers at once go through the batched
// quisk_cuda_rxa_* handle API instead (same chain, same exchange code).  The stages this library does not build (AM /
// FM squelch, ANF, ANR, EQ ...; NR2 and SNB ARE built: wdsp_emnr_nofma.cu, wdsp_snba_nofma.cu) are accepted when switched OFF and refused loudly when switched on.
#include "wdsp_internal.h"
#include <chrono>
#include <cmath>
#include <thread>

using namespace qc;

namespace {

constexpr int MAX_CHANNELS = 32;            // wdsp/comm.h:117, quisk_wdsp.c:10
struct Chan { qcRxa *h = nullptr; };
Chan g_ch[MAX_CHANNELS];
std::recursive_mutex g_mu;                  // csDSP / csEXCH rolled into one: setters and the exchange never interleave

Rxa *chan(int channel, const char *fn)
{
    if (channel < 0 || channel >= MAX_CHANNELS || !g_ch[channel].h) {
        fprintf(stderr, "libquisk_cuda: %s: channel %d is not open\n", fn, channel);
        return nullptr;
    }
    return &g_ch[channel].h->r;
}

void unsupported(const char *fn, int channel)
{
    fprintf(stderr, "libquisk_cuda: %s(channel %d, run = 1): this WDSP stage is not built in libquisk_cuda; it stays off\n", fn, channel);
}

// the batched setters of quisk_cuda_wdsp.h are reused through the channel's own handle
inline qcRxa *as_handle(Rxa *r) { for (auto &c : g_ch) if (c.h && &c.h->r == r) return c.h; return nullptr; }

}  // namespace

extern "C" {

int GetWDSPVersion(void) { return 125; }    // version.c:9: WDSP 1.25, what quisk_wdsp.py logs (quisk_wdsp.py:45-56)

void OpenChannel(int channel, int in_size, int dsp_size, int input_samplerate, int dsp_rate, int output_samplerate,
                 int type, int state, double tdelayup, double tslewup, double tdelaydown, double tslewdown, int bfo)
{
    std::lock_guard<std::recursive_mutex> g(g_mu);
    if (channel < 0 || channel >= MAX_CHANNELS) { fprintf(stderr, "libquisk_cuda: OpenChannel: channel %d out of range\n", channel); return; }
    if (type != 0) { fprintf(stderr, "libquisk_cuda: OpenChannel(channel %d): type %d (TX) is not on the accelerated path\n", channel, type); return; }
    if (ensure_device() != QC_OK) die_no_device("OpenChannel");
    if (g_ch[channel].h) { g_ch[channel].h->r.release(); delete g_ch[channel].h; g_ch[channel].h = nullptr; }
    qcRxa *h = new qcRxa();
    Rxa *r = &h->r;
    r->state = state ? 1 : 0; r->exchange_on = r->state; r->bfo = bfo ? 1 : 0;
    r->tdelaydown = tdelaydown; r->tslewdown = tslewdown;
    if (r->init(1, in_size, dsp_size, input_samplerate, dsp_rate, output_samplerate) != QC_OK ||
        r->arm_upslew(tdelayup, tslewup) != QC_OK) {
        fprintf(stderr, "libquisk_cuda: OpenChannel(channel %d) failed: %s\n", channel, quisk_cuda_last_error());
        r->release(); delete h;
        return;
    }
    if (!r->state) cudaMemset(r->d_uslew, 0, 3 * sizeof(int));     // opened in state 0: no upflag until SetChannelState(1)
    g_ch[channel].h = h;
}

void CloseChannel(int channel)
{
    std::lock_guard<std::recursive_mutex> g(g_mu);
    if (channel < 0 || channel >= MAX_CHANNELS || !g_ch[channel].h) return;
    g_ch[channel].h->r.release(); delete g_ch[channel].h; g_ch[channel].h = nullptr;
}

int SetChannelState(int channel, int state, int dmode)
{
    Rxa *r;
    {
        std::lock_guard<std::recursive_mutex> g(g_mu);
        r = chan(channel, "SetChannelState"); if (!r) return 0;
        if (!(state == 0 && dmode && r->state == 1)) return r->set_channel_state(state, 0);
        r->set_channel_state(0, 0);                                  // downflag + flushflag up
    }
    // dmode = 1: wait (<= 100 ms, channel.c:277-282) for another thread's fexchange0 calls to finish the ramp and flush
    int count = 0;
    for (; count < 100; count++) {
        { std::lock_guard<std::recursive_mutex> g(g_mu); if (!r->flushflag) break; }
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    if (count >= 100) { std::lock_guard<std::recursive_mutex> g(g_mu); r->exchange_on = 0; r->flushflag = 0; r->downflag = 0; }
    return 1;
}

void fexchange0(int channel, double *in, double *out, int *error)
{
    std::lock_guard<std::recursive_mutex> g(g_mu);
    int e = 0;
    Rxa *r = chan(channel, "fexchange0");
    if (r && r->exchange(in, out, &e) != QC_OK) {
        fprintf(stderr, "libquisk_cuda: fexchange0(channel %d): %s\n", channel, quisk_cuda_last_error());
        e = -1;
    }
    if (error) *error = e;
}

// ---- RXA properties ----
#define CH(fn) std::lock_guard<std::recursive_mutex> g(g_mu); Rxa *r = chan(channel, fn); if (!r)

void SetRXAMode(int channel, int mode) { CH("SetRXAMode") return; quisk_cuda_rxa_set_mode(as_handle(r), mode); }
void RXASetPassband(int channel, double f_low, double f_high) { CH("RXASetPassband") return; quisk_cuda_rxa_set_passband(as_handle(r), f_low, f_high); }
void RXASetNC(int channel, int nc) { CH("RXASetNC") return; if (quisk_cuda_rxa_set_nc(as_handle(r), nc) != QC_OK) fprintf(stderr, "libquisk_cuda: RXASetNC: %s\n", quisk_cuda_last_error()); }
void RXASetMP(int channel, int mp) { CH("RXASetMP") return; quisk_cuda_rxa_set_mp(as_handle(r), mp); }
void SetRXAShiftRun(int channel, int run) { CH("SetRXAShiftRun") return; r->shift_run = run; }
void SetRXAShiftFreq(int channel, double fshift)
{   // calc_shift (shift.c:29-34): new delta, the running phase is kept
    CH("SetRXAShiftFreq") return;
    const double delta = 6.2831853071795864 * fshift / (double)r->in_rate;
    const double p[3] = {delta, cos(delta), sin(delta)};
    cudaDeviceSynchronize();
    cudaMemcpy(r->shift->d_par, p, sizeof(p), cudaMemcpyHostToDevice);
    r->shift_nonzero = r->shift_nonzero || fshift != 0.0;          // once the phase has moved it has to keep being applied
}
void RXANBPSetRun(int channel, int run) { CH("RXANBPSetRun") return; r->nbp_run = run; }
void RXANBPSetFreqs(int channel, double flow, double fhigh)
{   // nbp.c:528-540: nbp0 only (RXASetPassband also moves bp1)
    CH("RXANBPSetFreqs") return;
    if (flow != r->nbp_flow || fhigh != r->nbp_fhigh) {
        r->nbp_flow = flow; r->nbp_fhigh = fhigh;
        std::vector<double> imp;
        if (r->nbp0_impulse(imp, &r->nbp_hadnotch) == QC_OK) r->nbp0->set_impulse(imp.data(), 1);
    }
}
int RXANBPAddNotch(int channel, int notch, double fcenter, double fwidth, int active)
{ CH("RXANBPAddNotch") return -1; return quisk_cuda_rxa_nbp_add_notch(as_handle(r), notch, fcenter, fwidth, active); }
int RXANBPDeleteNotch(int channel, int notch) { CH("RXANBPDeleteNotch") return -1; return quisk_cuda_rxa_nbp_delete_notch(as_handle(r), notch); }
void RXANBPGetNumNotches(int channel, int *nnotches) { CH("RXANBPGetNumNotches") return; if (nnotches) *nnotches = (int)r->ndb_fcenter.size(); }
void RXANBPSetNotchesRun(int channel, int run) { CH("RXANBPSetNotchesRun") return; quisk_cuda_rxa_nbp_set_notches_run(as_handle(r), run); }
void RXANBPSetTuneFrequency(int channel, double tunefreq) { CH("RXANBPSetTuneFrequency") return; quisk_cuda_rxa_nbp_set_tune_frequency(as_handle(r), tunefreq); }
void RXANBPSetShiftFrequency(int channel, double shift) { CH("RXANBPSetShiftFrequency") return; quisk_cuda_rxa_nbp_set_shift_frequency(as_handle(r), shift); }
void SetRXAAGCMode(int channel, int mode) { CH("SetRXAAGCMode") return; quisk_cuda_rxa_set_agc_mode(as_handle(r), mode); }
void SetRXAAGCFixed(int channel, double fixed_agc) { CH("SetRXAAGCFixed") return; quisk_cuda_rxa_set_agc_fixed(as_handle(r), fixed_agc); }
void SetRXAFMLimRun(int channel, int run) { CH("SetRXAFMLimRun") return; quisk_cuda_rxa_set_fm_lim_run(as_handle(r), run); }
void SetRXAFMLimGain(int channel, double gaindB) { CH("SetRXAFMLimGain") return; quisk_cuda_rxa_set_fm_lim_gain(as_handle(r), gaindB); }
void SetRXAAGCTop(int channel, double max_agc) { CH("SetRXAAGCTop") return; r->agc->agc.max_gain = pow(10.0, max_agc / 20.0); r->agc->load_agc(); }
void SetRXAPanelRun(int channel, int run) { CH("SetRXAPanelRun") return; (void)run; }      // xpanel never looks at its run flag
void SetRXAPanelGain1(int channel, double gain) { CH("SetRXAPanelGain1") return; r->panel_gain1 = gain; }
void SetRXAPanelGain2(int channel, double gainI, double gainQ) { CH("SetRXAPanelGain2") return; r->panel_gain2I = gainI; r->panel_gain2Q = gainQ; }
void SetRXAAMDSBMode(int channel, int sbmode) { CH("SetRXAAMDSBMode") return; r->amd->par[2] = sbmode; }
void SetRXAAMDFadeLevel(int channel, int levelfade) { CH("SetRXAAMDFadeLevel") return; r->amd->par[1] = levelfade; }
double GetRXAMeter(int channel, int mt)
{   // enum rxaMeterType (RXA.h:47-57): S_PK, S_AV, ADC_PK, ADC_AV, AGC_GAIN, AGC_PK, AGC_AV
    CH("GetRXAMeter") return -400.0;
    static const int which[7] = {1, 1, 0, 0, 2, 2, 2}, field[7] = {1, 0, 1, 0, 2, 1, 0};
    if (mt < 0 || mt > 6) return -400.0;
    double v[3] = {-400.0, -400.0, 0.0};
    quisk_cuda_rxa_get_meter(as_handle(r), which[mt], &v[0], &v[1], &v[2]);
    return v[field[mt]];
}
void RXAGetaSipF(int channel, float *out, int size) { CH("RXAGetaSipF") return; quisk_cuda_rxa_get_siphon(as_handle(r), out, size, 0); }
void RXAGetaSipF1(int channel, float *out, int size) { CH("RXAGetaSipF1") return; quisk_cuda_rxa_get_siphon(as_handle(r), out, size, 1); }
// stages create_rxa builds switched off and this library does not have: fine while they stay off
void SetRXAAMSQRun(int channel, int run) { CH("SetRXAAMSQRun") return; if (run) unsupported("SetRXAAMSQRun", channel); }
void SetRXAFMSQRun(int channel, int run) { CH("SetRXAFMSQRun") return; if (run) unsupported("SetRXAFMSQRun", channel); }
void SetRXAEMNRRun(int channel, int run)
{   // NR2.  Gain methods 0, 1 and 2 are built; 2 (create_rxa's and Quisk's choice) needs the distribution's tables first
    CH("SetRXAEMNRRun") return;
    if (quisk_cuda_rxa_set_emnr_run(as_handle(r), run) != QC_OK) fprintf(stderr, "libquisk_cuda: SetRXAEMNRRun(%d, %d): %s\n", channel, run, quisk_cuda_last_error());
}
void SetRXAEMNRgainMethod(int channel, int method) { CH("SetRXAEMNRgainMethod") return; quisk_cuda_rxa_set_emnr_gain_method(as_handle(r), method); }
void SetRXAEMNRnpeMethod(int channel, int method) { CH("SetRXAEMNRnpeMethod") return; quisk_cuda_rxa_set_emnr_npe_method(as_handle(r), method); }
void SetRXAEMNRaeRun(int channel, int run) { CH("SetRXAEMNRaeRun") return; quisk_cuda_rxa_set_emnr_ae_run(as_handle(r), run); }
void SetRXAEMNRtrainZetaThresh(int channel, double thresh) { CH("SetRXAEMNRtrainZetaThresh") return; quisk_cuda_rxa_set_emnr_train(as_handle(r), 0, thresh); }    /* emnr.c:1160-1166 */
void SetRXAEMNRtrainT2(int channel, double t2) { CH("SetRXAEMNRtrainT2") return; quisk_cuda_rxa_set_emnr_train(as_handle(r), 1, t2); }                          /* emnr.c:1168-1174 */
void SetRXAEMNRPosition(int channel, int position) { CH("SetRXAEMNRPosition") return; quisk_cuda_rxa_set_emnr_position(as_handle(r), position); }
void SetRXASNBARun(int channel, int run)
{   // SNB
    CH("SetRXASNBARun") return;
    if (quisk_cuda_rxa_set_snba_run(as_handle(r), run) != QC_OK) fprintf(stderr, "libquisk_cuda: SetRXASNBARun(%d, %d): %s\n", channel, run, quisk_cuda_last_error());
}
void SetRXAANFRun(int channel, int run) { CH("SetRXAANFRun") return; if (run) unsupported("SetRXAANFRun", channel); }
void SetRXAANRRun(int channel, int run) { CH("SetRXAANRRun") return; if (run) unsupported("SetRXAANRRun", channel); }
#undef CH

// ---- wdspFexchange0 (quisk_wdsp.c:24-73): Quisk's side of the boundary.  Arbitrary sample counts in, scaled by 1 / CLIP32,
// re-blocked to in_size, fexchange0 per block, scaled back; returns the number of samples now in cSamples. ----
namespace {
struct ReBlock { std::vector<double> buf; int sizeBuf = 0, nBuf = 0, in_size = 0, in_use = 0, Windex = 0, Rindex = 0; };
ReBlock g_rb[MAX_CHANNELS];
constexpr double CLIP32 = 2147483647.0;
}

void quisk_cuda_wdsp_set_parameter(int channel, int in_size, int in_use)
{   // quisk_wdsp_set_parameter (quisk_wdsp.c:75-92) without the Python argument parsing; negative = leave alone
    std::lock_guard<std::recursive_mutex> g(g_mu);
    if (channel < 0 || channel >= MAX_CHANNELS) return;
    if (in_size > 0) g_rb[channel].in_size = in_size;
    if (in_use >= 0) g_rb[channel].in_use = in_use;
}

int wdspFexchange0(int channel, quisk_cd *cSamples, int nSamples)
{
    std::lock_guard<std::recursive_mutex> g(g_mu);
    if (channel < 0 || channel >= MAX_CHANNELS) return nSamples;
    ReBlock &b = g_rb[channel];
    if (!b.in_use) { b.Windex = 0; b.Rindex = 0; b.nBuf = 0; return nSamples; }
    if (nSamples <= 0 || b.in_size <= 0) return nSamples;
    const int in_size = b.in_size;
    int i = nSamples / in_size + 3;                                 // blocks needed for the samples plus a partial block
    if (i * in_size > b.sizeBuf) { b.sizeBuf = i * in_size; b.buf.resize((size_t)2 * b.sizeBuf); }
    double *x = reinterpret_cast<double *>(cSamples);
    for (i = 0; i < nSamples; i++) {
        // `cSamples[i] / CLIP32` is complex / int in the reference: gcc divides both parts by (double)CLIP32
        b.buf[(size_t)2 * b.Windex] = x[2 * i] / CLIP32;
        b.buf[(size_t)2 * b.Windex + 1] = x[2 * i + 1] / CLIP32;
        if (++b.Windex >= b.sizeBuf) b.Windex = 0;
    }
    b.nBuf += nSamples;
    nSamples = 0;
    while (b.nBuf >= in_size) {
        int error = 0;
        fexchange0(channel, &b.buf[(size_t)2 * b.Rindex], x + 2 * (size_t)nSamples, &error);
        if (error) printf("WDSP: wdsp_fexchange0 error %d\n", error);
        b.Rindex += in_size;
        if (b.Rindex >= b.sizeBuf) b.Rindex = 0;
        nSamples += in_size;
        b.nBuf -= in_size;
    }
    for (i = 0; i < 2 * nSamples; i++) x[i] *= CLIP32;
    return nSamples;
}

}  // extern "C"
