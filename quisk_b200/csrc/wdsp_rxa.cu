// quisk_b200/csrc/wdsp_rxa.cu -- quisk_cuda_rxa_*: WDSP's receive chain for a batch of channels.
//
// Mirrors create_rxa's stage list and defaults (wdsp/RXA.c:31-490) and xrxa's order
// (wdsp/RXA.c:561-598) for the stages that run in a default / Quisk-configured channel:
//   shift -> resample(in) -> ADC meter -> nbp0 -> S meter -> amd -> fmd -> bp1 -> wcpagc -> AGC meter
//   -> panel -> resample(out)
// The stages that create_rxa builds with run = 0 (gen, bpsnba, sender, amsq, fmsq, snba, eq, anf, anr,
// emnr, cbl, speak, mpeak, ssql, siphon) are outside this round's scope and are not instantiated.
// Quirks kept: bp1 is created running and stays so until a mode change calls RXAbp1Set;
// xpanel ignores its run flag (F9); the panel gain is gain1 * gain2 = 4.0.
#include "wdsp_internal.h"
#include <cmath>

namespace qc {


// upslew0 (iobuffs.c:98-160), one thread per channel, in place on the block about to enter the DSP chain: zeros until
// the first non-zero sample (which is swallowed too), ndelup more zeros, a raised-cosine ramp of ntup + 1 samples,
// then pass-through; the channel's flag drops at the end of the first block that finishes in the ON state.
enum { U_BEGIN = 0, U_DELAYUP = 1, U_UPSLEW = 2, U_ON = 3 };
__global__ void upslew_kernel(cd *x, long stride, int n, int C, int *st, const double *cup, int ndelup, int ntup)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C || !st[c * 3 + 2]) return;
    int ustate = st[c * 3], ucount = st[c * 3 + 1], flag = 1;
    cd *p = x + (size_t)c * stride;
    for (int i = 0; i < n; i++) {
        const cd v = p[i];
        switch (ustate) {
        case U_BEGIN:
            p[i] = make_double2(0.0, 0.0);
            if (v.x != 0.0 || v.y != 0.0) {
                if (ndelup > 0) { ustate = U_DELAYUP; ucount = ndelup; }
                else if (ntup > 0) { ustate = U_UPSLEW; ucount = ntup; }
                else ustate = U_ON;
            }
            break;
        case U_DELAYUP:
            p[i] = make_double2(0.0, 0.0);
            if (ucount-- == 0) {
                if (ntup > 0) { ustate = U_UPSLEW; ucount = ntup; }
                else ustate = U_ON;
            }
            break;
        case U_UPSLEW: {
            const double g = cup[ntup - ucount];
            p[i] = make_double2(v.x * g, v.y * g);
            if (ucount-- == 0) ustate = U_ON;
            break; }
        case U_ON:
            if (i == n - 1) { ustate = U_BEGIN; flag = 0; }
            break;
        }
    }
    st[c * 3] = ustate; st[c * 3 + 1] = ucount; st[c * 3 + 2] = flag;
}

int Rxa::arm_upslew(double tdelayup, double tslewup)
{   // create_slews / flush_slews + the upflag OpenChannel and SetChannelState(1) raise (iobuffs.c:47-96, channel.c:95,291)
    this->tdelayup = tdelayup; this->tslewup = tslewup;
    ndelup = (int)(tdelayup * in_rate);
    ntup = (int)(tslewup * in_rate);
    std::vector<double> cup((size_t)ntup + 1);
    const double delta = 3.1415926535897932 / (double)ntup;      // PI / 0 = inf when ntup == 0: cup[0] = 0 and is never used
    double theta = 0.0;
    for (int i = 0; i <= ntup; i++) { cup[i] = 0.5 * (1.0 - cos(theta)); theta += delta; }
    if (d_cup) cudaFree(d_cup);
    d_cup = nullptr;
    QC_CUDA(cudaMalloc((void **)&d_cup, cup.size() * sizeof(double)));
    QC_CUDA(cudaMemcpy(d_cup, cup.data(), cup.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (!d_uslew) QC_CUDA(cudaMalloc((void **)&d_uslew, (size_t)C * 3 * sizeof(int)));
    std::vector<int> st((size_t)C * 3, 0);
    for (int c = 0; c < C; c++) st[(size_t)c * 3 + 2] = 1;
    QC_CUDA(cudaMemcpy(d_uslew, st.data(), st.size() * sizeof(int), cudaMemcpyHostToDevice));
    return QC_OK;
}

static FirCore *new_fircore(int C, int size, int nc, const std::vector<double> &imp)
{
    FirCore *f = new FirCore();
    if (f->init(C, size, nc, 0, imp.data()) != QC_OK) { f->release(); delete f; return nullptr; }
    return f;
}

// calc_nbp_impulse (nbp.c:214-239): gain / (2 size) baked in; with the notches running the pass band is cut up first
int Rxa::nbp0_impulse(std::vector<double> &imp, int *havnotch)
{
    imp.assign((size_t)2 * nbp_nc, 0.0);
    if (havnotch) *havnotch = 0;
    const double scale = 1.0 / (double)(2 * dsp_size);
    if (!ndb_run) return quisk_cuda_fir_bandpass(nbp_nc, nbp_flow, nbp_fhigh, (double)dsp_rate, 0, 1, scale, imp.data());
    return quisk_cuda_nbp_impulse(nbp_nc, nbp_flow, nbp_fhigh, (double)dsp_rate, 0, scale, (int)ndb_fcenter.size(),
                                  ndb_fcenter.data(), ndb_fwidth.data(), ndb_active.data(), ndb_tune, ndb_shift,
                                  1 /* autoincr, RXA.c:104 */, 1025 /* maxpb, RXA.c:105 */, imp.data(), nullptr, havnotch);
}

int Rxa::make_nbp0()
{   // create_nbp (nbp.c:241-270)
    std::vector<double> imp;
    int rc = nbp0_impulse(imp, &nbp_hadnotch); if (rc != QC_OK) return rc;
    if (nbp0) { nbp0->release(); delete nbp0; }
    nbp0 = new_fircore(C, dsp_size, nbp_nc, imp);
    return nbp0 ? QC_OK : QC_EINVAL;
}

int Rxa::bpsnba_impulse(std::vector<double> &imp, int *havnotch)
{   // calc_nbp_impulse of the nbp inside bpsnba (snb.c:706-727, 807-821): gain 1, wintype 0, the channel's notch database
    imp.assign((size_t)2 * bps_nc, 0.0);
    if (havnotch) *havnotch = 0;
    const double scale = 1.0 / (double)(2 * dsp_size);
    if (!bps_run_notches) return quisk_cuda_fir_bandpass(bps_nc, bps_flow, bps_fhigh, (double)dsp_rate, 0, 1, scale, imp.data());
    return quisk_cuda_nbp_impulse(bps_nc, bps_flow, bps_fhigh, (double)dsp_rate, 0, scale, (int)ndb_fcenter.size(),
                                  ndb_fcenter.data(), ndb_fwidth.data(), ndb_active.data(), ndb_tune, ndb_shift, 1, 1025, imp.data(), nullptr, havnotch);
}

int Rxa::make_bpsnba()
{
    std::vector<double> imp;
    int rc = bpsnba_impulse(imp, &bps_hadnotch); if (rc != QC_OK) return rc;
    if (bpsnba) { bpsnba->release(); delete bpsnba; }
    bpsnba = new_fircore(C, dsp_size, bps_nc, imp);
    return bpsnba ? QC_OK : QC_EINVAL;
}

int Rxa::bpsnba_check(int mode_, int notch_run)
{   // RXAbpsnbaCheck, RXA.c:829-881: the band follows the mode's side band between 250 and 5700 Hz; new masks wait for the update
    const double abs_low = 250.0, abs_high = 5700.0;
    double f_low = 0.0, f_high = 0.0;
    int run_notches = 0;
    switch (mode_) {
    case QC_RXA_LSB: case QC_RXA_CWL: case QC_RXA_DIGL: f_low = -abs_high; f_high = -abs_low; run_notches = notch_run; break;
    case QC_RXA_USB: case QC_RXA_CWU: case QC_RXA_DIGU: f_low = +abs_low; f_high = +abs_high; run_notches = notch_run; break;
    case QC_RXA_AM: case QC_RXA_SAM: case QC_RXA_DSB: case QC_RXA_FM: f_low = +abs_low; f_high = +abs_high; run_notches = 0; break;
    default: break;
    }
    if (bps_flow != f_low || bps_fhigh != f_high || bps_run_notches != run_notches) {
        bps_flow = f_low; bps_fhigh = f_high; bps_run_notches = run_notches;
        std::vector<double> imp;
        int rc = bpsnba_impulse(imp, &bps_hadnotch); if (rc != QC_OK) return rc;
        return bpsnba->set_impulse(imp.data(), 0);
    }
    return QC_OK;
}

int Rxa::bpsnba_set()
{   // RXAbpsnbaSet, RXA.c:883-918
    switch (mode) {
    case QC_RXA_LSB: case QC_RXA_CWL: case QC_RXA_DIGL: case QC_RXA_USB: case QC_RXA_CWU: case QC_RXA_DIGU: bps_run = snba_run; bps_position = 0; break;
    case QC_RXA_AM: case QC_RXA_SAM: case QC_RXA_DSB: case QC_RXA_FM: bps_run = snba_run; bps_position = 1; break;
    default: bps_run = 0; break;
    }
    return bpsnba->update();
}

int Rxa::make_bp1()
{   // create_bandpass (bandpass.c:284-306), wintype 1
    std::vector<double> imp((size_t)2 * bp1_nc);
    quisk_cuda_fir_bandpass(bp1_nc, bp1_flow, bp1_fhigh, (double)dsp_rate, 1, 1, bp1_gain / (double)(2 * dsp_size), imp.data());
    if (bp1) { bp1->release(); delete bp1; }
    bp1 = new_fircore(C, dsp_size, bp1_nc, imp);
    return bp1 ? QC_OK : QC_EINVAL;
}

int Rxa::make_fmd()
{   // create_fmd (fmd.c:81-120) with create_rxa's arguments (RXA.c:192-212)
    const double f_low = 300.0, f_high = 3000.0, afgain = 0.5, rate = (double)dsp_rate;
    std::vector<double> imp((size_t)2 * fm_nc_de);
    quisk_cuda_fc_impulse(fm_nc_de, f_low, f_high, +20.0 * log10(f_high / f_low), 0.0, 1, rate, 1.0 / (2.0 * dsp_size), 0, 0, imp.data());
    if (pde) { pde->release(); delete pde; }
    pde = new_fircore(C, dsp_size, fm_nc_de, imp);
    imp.assign((size_t)2 * fm_nc_aud, 0.0);
    quisk_cuda_fir_bandpass(fm_nc_aud, 0.8 * f_low, 1.1 * f_high, rate, 0, 1, afgain / (2.0 * dsp_size), imp.data());
    if (paud) { paud->release(); delete paud; }
    paud = new_fircore(C, dsp_size, fm_nc_aud, imp);
    return pde && paud ? QC_OK : QC_EINVAL;
}

int Rxa::init(int C_, int in_size_, int dsp_size_, int in_rate_, int dsp_rate_, int out_rate_)
{
    C = C_; in_size = in_size_; dsp_size = dsp_size_; in_rate = in_rate_; dsp_rate = dsp_rate_; out_rate = out_rate_;
    if (C <= 0 || dsp_size <= 0 || in_rate <= 0 || dsp_rate <= 0 || out_rate <= 0) { set_error("rxa_create: bad arguments"); return QC_EINVAL; }
    // pre_main_build, channel.c:36-58
    dsp_insize = in_rate >= dsp_rate ? dsp_size * (in_rate / dsp_rate) : dsp_size / (dsp_rate / in_rate);
    dsp_outsize = out_rate >= dsp_rate ? dsp_size * (out_rate / dsp_rate) : dsp_size / (dsp_rate / out_rate);
    out_size = in_rate >= out_rate ? in_size / (in_rate / out_rate) : in_size * (out_rate / in_rate);
    const size_t mlen = (size_t)C * (2 * (dsp_size > dsp_insize ? dsp_size : dsp_insize) + 64);
    QC_CUDA(cudaMalloc((void **)&mid, mlen * sizeof(cd)));
    QC_CUDA(cudaMalloc((void **)&mid2, mlen * sizeof(cd)));
    QC_CUDA(cudaMalloc((void **)&audio, (size_t)C * dsp_size * sizeof(cd)));
    shift = make_shift(C, in_rate, nullptr);
    if (in_rate != dsp_rate) {          // RXAResCheck, RXA.c:789-798
        rsmpin = new Resampler();
        if (rsmpin->init(C, in_rate, dsp_rate, 0.0, 0, 1.0) != QC_OK) return QC_EINVAL;
    }
    if (dsp_rate != out_rate) {
        rsmpout = new Resampler();
        if (rsmpout->init(C, dsp_rate, out_rate, 0.0, 0, 1.0) != QC_OK) return QC_EINVAL;
    }
    adcmeter = make_meter(C, dsp_rate, 0.100, 0.100);
    smeter = make_meter(C, dsp_rate, 0.100, 0.100);
    agcmeter = make_meter(C, dsp_rate, 0.100, 0.100);
    nbp_nc = bp1_nc = bps_nc = fm_nc_de = fm_nc_aud = dsp_size > 2048 ? dsp_size : 2048;       // max(2048, dsp_size)
    int rc;
    if ((rc = make_nbp0()) != QC_OK) return rc;
    if ((rc = make_bp1()) != QC_OK) return rc;
    if ((rc = make_bpsnba()) != QC_OK) return rc;
    QC_CUDA(cudaMalloc((void **)&bps_buff, (size_t)C * dsp_size * sizeof(cd)));
    QC_CUDA(cudaMemset(bps_buff, 0, (size_t)C * dsp_size * sizeof(cd)));
    snba = make_snba(C, dsp_rate, 12000, dsp_size, 4, 256, 64, 2, 8.0, 20.0, 10, 2, 2, 0.5, 200.0, 5400.0);     // create_rxa's arguments, RXA.c:237-255
    amd = make_amd(C, dsp_rate, 0, 1, 0);
    fmpll = make_fmpll(C, dsp_rate, 5000.0, -8000.0, +8000.0, 1.0, 20000.0, 0.02);
    sntch = make_snotch(C, dsp_rate, 254.1, 0.0002);
    plim = make_wcpagc_fmlim(C, dsp_rate, lim_gain);
    if (!plim) return QC_EINVAL;
    emnr = make_emnr(C, dsp_size, 4096, 4, dsp_rate, 0, 1.0, 2, 0, 1);        // create_rxa's arguments, RXA.c:319-332
    if (!emnr) return QC_EINVAL;
    if ((rc = make_fmd()) != QC_OK) return rc;
    agc = make_wcpagc(C, dsp_rate, 3);
    if (!shift || !adcmeter || !smeter || !agcmeter || !amd || !fmpll || !sntch || !agc) return QC_EINVAL;
    QC_CUDA(cudaMalloc((void **)&d_sip, (size_t)C * sipsize * sizeof(cd)));
    QC_CUDA(cudaMemset(d_sip, 0, (size_t)C * sipsize * sizeof(cd)));
    int rcu = arm_upslew(0.0, 0.0); if (rcu != QC_OK) return rcu;
    return setup_exchange();
}

static int raise_upflag(Rxa &r)
{   // InterlockedBitTestAndSet(&slew.upflag) of OpenChannel / SetChannelState(1) (channel.c:95, 291)
    std::vector<int> st((size_t)r.C * 3, 0);
    if (r.hs) QC_CUDA(cudaStreamSynchronize(r.hs));
    QC_CUDA(cudaMemcpy(st.data(), r.d_uslew, st.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int c = 0; c < r.C; c++) st[(size_t)c * 3 + 2] = 1;
    QC_CUDA(cudaMemcpy(r.d_uslew, st.data(), st.size() * sizeof(int), cudaMemcpyHostToDevice));
    return QC_OK;
}

void Rxa::release()
{
    if (emnr) { emnr_destroy(emnr); emnr = nullptr; }
    if (snba) { snba_destroy(snba); snba = nullptr; }
    if (bpsnba) { bpsnba->release(); delete bpsnba; bpsnba = nullptr; }
    if (bps_buff) { cudaFree(bps_buff); bps_buff = nullptr; }
    for (SeqStage **p : {&shift, &adcmeter, &smeter, &agcmeter, &amd, &fmpll, &sntch, &agc, &plim}) if (*p) { (*p)->release(); delete *p; *p = nullptr; }
    for (FirCore **p : {&nbp0, &bp1, &pde, &paud}) if (*p) { (*p)->release(); delete *p; *p = nullptr; }
    for (Resampler **p : {&rsmpin, &rsmpout}) if (*p) { (*p)->release(); delete *p; *p = nullptr; }
    if (mid) cudaFree(mid); if (mid2) cudaFree(mid2); if (audio) cudaFree(audio);
    if (d_r1) cudaFree(d_r1); if (d_r2) cudaFree(d_r2); if (d_outbuff) cudaFree(d_outbuff); if (d_xout) cudaFree(d_xout);
    if (d_gain) cudaFree(d_gain);
    d_r1 = d_r2 = d_outbuff = d_xout = nullptr; d_gain = nullptr;
    if (hs) { cudaStreamSynchronize(hs); cudaStreamDestroy(hs); hs = nullptr; }
    if (d_uslew) cudaFree(d_uslew); if (d_cup) cudaFree(d_cup);
    d_uslew = nullptr; d_cup = nullptr;
    if (side) { cudaStreamSynchronize(side); cudaStreamDestroy(side); side = nullptr; for (cudaEvent_t *e : {&ev_a, &ev_b, &ev_c, &ev_d}) if (*e) { cudaEventDestroy(*e); *e = nullptr; } }
    if (wmid) cudaFree(wmid); if (wmid2) cudaFree(wmid2); if (waudio) cudaFree(waudio);
    wmid = wmid2 = waudio = nullptr; wstride = 0;
    if (d_wide_spec) cudaFree(d_wide_spec); if (d_wide_y) cudaFree(d_wide_y);
    if (d_wide_seq) cudaFree(d_wide_seq); d_wide_seq = nullptr; wide_seq_cap = 0;
    d_wide_spec = d_wide_y = nullptr; wide_spec_cap = wide_y_cap = 0;
    if (d_sip) cudaFree(d_sip); if (d_sipout) cudaFree(d_sipout);
    d_sip = nullptr; d_sipout = nullptr; sipout_cap = 0;
    mid = mid2 = audio = nullptr;
}

int Rxa::emnr_run_stage(cd *m, long ms, cudaStream_t s) { return qc::emnr_run(emnr, m, ms, m, ms, s); }
int Rxa::snba_run_stage(cd *m, long ms, cudaStream_t s) { return qc::snba_run(snba, m, ms, m, ms, s); }

int Rxa::bp1_check_set()
{   // RXAbp1Check (gain 2 when the AM demodulator or a noise reducer feeds bp1; new masks wait for setUpdate) + RXAbp1Set
    const int feeds = amd_run || snba_run || emnr_run;
    const double gain = feeds ? 2.0 : 1.0;
    if (bp1_gain != gain) {
        bp1_gain = gain;
        std::vector<double> imp((size_t)2 * bp1_nc);
        quisk_cuda_fir_bandpass(bp1_nc, bp1_flow, bp1_fhigh, (double)dsp_rate, 1, 1, bp1_gain / (double)(2 * dsp_size), imp.data());
        int rc = bp1->set_impulse(imp.data(), 0); if (rc) return rc;
    }
    const int old = bp1_run;
    bp1_run = feeds ? 1 : 0;
    if (!old && bp1_run) { int rc = bp1->flush(); if (rc) return rc; }
    return bp1->update();
}

int Rxa::fm_limiter(cd *m, long ms, int n, cudaStream_t s)
{   // fmd.c:179-184: out *= lim_pre_gain, then the detector limiter (a wcpAGC with calc_fmd's constants) in place
    if (!lim_run) return QC_OK;
    int rc = launch_panel(m, ms, m, ms, n, C, lim_pre_gain, lim_pre_gain, 3, 0, s); if (rc) return rc;
    return plim->run(m, ms, m, ms, n, s);
}

int Rxa::xrxa_multi(const void *din, long is, void *dout, long os, int nblocks, cudaStream_t s)
{   // nblocks consecutive DSP blocks per channel: block b of a channel at in + b * dsp_insize, out + b * dsp_outsize
    if (nblocks <= 0) return QC_OK;
    if (emnr_run || snba_run) {         // the noise reducer / blanker frame their own streams: block by block through the per-stage chain
        for (int b = 0; b < nblocks; b++) {
            int rc = xrxa((const cd *)din + (size_t)b * dsp_insize, is, (cd *)dout + (size_t)b * dsp_outsize, os, s);
            if (rc != QC_OK) return rc;
        }
        return QC_OK;
    }
    if (fusable()) return xrxa_fused(din, is, dout, os, nblocks, s);
    // the other configurations: every stage is a streaming operator with carried state, so it can take a GROUP of blocks per
    // launch (the fircores as wide transform grids); the group is bounded by what the recurrent kernels stage in shared memory
    int g = 8192 / dsp_size;
    if (const char *e = getenv("QUISK_RXA_GROUP")) g = atoi(e);
    if (fused_ok && g >= 2 && !(shift_run && shift_nonzero) && !(agc_run && agc->agc.mode == 5)) {
        for (int b = 0; b < nblocks; b += g) {
            const int gb = nblocks - b < g ? nblocks - b : g;
            int rc = gb >= 2 ? xrxa_stages_wide((const cd *)din + (size_t)b * dsp_insize, is, (cd *)dout + (size_t)b * dsp_outsize, os, gb, s)
                             : xrxa((const cd *)din + (size_t)b * dsp_insize, is, (cd *)dout + (size_t)b * dsp_outsize, os, s);
            if (rc != QC_OK) return rc;
        }
        return QC_OK;
    }
    for (int b = 0; b < nblocks; b++) {
        int rc = xrxa((const cd *)din + (size_t)b * dsp_insize, is, (cd *)dout + (size_t)b * dsp_outsize, os, s);
        if (rc != QC_OK) return rc;
    }
    return QC_OK;
}

int Rxa::xrxa_stages_wide(const void *din, long is, void *dout, long os, int g, cudaStream_t s)
{   // xrxa's stage order (RXA.c:561-598) with g blocks per stage launch
    int rc;
    const int n = g * dsp_size, nin = g * dsp_insize;
    const long need = (long)n + 64;                 // the group at the DSP rate: the input resampler reads the caller's buffer directly
    if (need > wstride) {
        for (cd **q : {&wmid, &wmid2, &waudio}) if (*q) { cudaFree(*q); *q = nullptr; }
        wstride = 0;
        QC_CUDA(cudaMalloc((void **)&wmid, (size_t)C * need * sizeof(cd)));
        QC_CUDA(cudaMalloc((void **)&wmid2, (size_t)C * need * sizeof(cd)));
        QC_CUDA(cudaMalloc((void **)&waudio, (size_t)C * need * sizeof(cd)));
        wstride = need;
    }
    const size_t need_spec = (size_t)C * g * 2 * dsp_size * 2;      // spectra + the partition MAC's products (fircore_wide)
    if (need_spec > wide_spec_cap) { if (d_wide_spec) cudaFree(d_wide_spec); d_wide_spec = nullptr; wide_spec_cap = 0;
                                     QC_CUDA(cudaMalloc((void **)&d_wide_spec, need_spec * sizeof(cd))); wide_spec_cap = need_spec; }
    const long ws = wstride;
    const cd *cur = (const cd *)din; long cs = is;
    cd *m = wmid;
    for (SeqStage *mt : {adcmeter, smeter, agcmeter}) mt->meter_sub = g;
    if (rsmpin) {
        int no = 0;
        rc = rsmpin->f->run(cur, cs, nin, m, ws, &no, 0, s); if (rc) return rc;
        if (no != n) { set_error("rxa: input resampler produced %d samples for %d blocks of %d", no, g, dsp_size); return QC_EINVAL; }
        cur = m; cs = ws;
    }
    if (fmd_run && !amd_run && nbp_run) {
        // FM: the two input-side meters only READ what the chain has produced, so they run on a side stream next to the stages
        // that follow (the filter writes into the second scratch buffer instead of in place to make that safe)
        if (!side) {
            QC_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
            for (cudaEvent_t *e : {&ev_a, &ev_b, &ev_c, &ev_d}) QC_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        }
        QC_CUDA(cudaEventRecord(ev_a, s)); QC_CUDA(cudaStreamWaitEvent(side, ev_a, 0));
        rc = adcmeter->run(cur, cs, nullptr, 0, n, side); if (rc) return rc;
        QC_CUDA(cudaEventRecord(ev_b, side));
        rc = fircore_wide(nbp0, cur, cs, wmid2, ws, g, d_wide_spec, s); if (rc) return rc;
        QC_CUDA(cudaEventRecord(ev_c, s)); QC_CUDA(cudaStreamWaitEvent(side, ev_c, 0));
        rc = smeter->run(wmid2, ws, nullptr, 0, n, side); if (rc) return rc;
        QC_CUDA(cudaEventRecord(ev_d, side));
        rc = fmpll->run(wmid2, ws, waudio, ws, n, s); if (rc) return rc;                // pll -> audio
        QC_CUDA(cudaStreamWaitEvent(s, ev_b, 0));                                       // the ADC meter has read wmid: it may be overwritten
        rc = fircore_wide(pde, waudio, ws, m, ws, g, d_wide_spec, s); if (rc) return rc;
        rc = fircore_wide(paud, m, ws, m, ws, g, d_wide_spec, s); if (rc) return rc;
        rc = sntch->run(m, ws, m, ws, n, s); if (rc) return rc;
        rc = fm_limiter(m, ws, n, s); if (rc) return rc;
        QC_CUDA(cudaStreamWaitEvent(s, ev_d, 0));                                       // and the S meter wmid2, before the next group's filter writes it
    } else {
    rc = adcmeter->run(cur, cs, nullptr, 0, n, s);
    if (rc == QC_OK && nbp_run) rc = fircore_wide(nbp0, cur, cs, m, ws, g, d_wide_spec, s);
    else if (rc == QC_OK && cur != m) rc = cudaMemcpy2DAsync(m, (size_t)ws * sizeof(cd), cur, (size_t)cs * sizeof(cd), (size_t)n * sizeof(cd), C, cudaMemcpyDeviceToDevice, s) == cudaSuccess ? QC_OK : QC_ECUDA;
    if (rc == QC_OK) rc = smeter->run(m, ws, nullptr, 0, n, s);
    if (rc == QC_OK && amd_run) rc = amd->run(m, ws, m, ws, n, s);
    if (rc == QC_OK && fmd_run) {
        rc = fmpll->run(m, ws, waudio, ws, n, s);                                       // pll -> audio
        if (rc == QC_OK) rc = fircore_wide(pde, waudio, ws, m, ws, g, d_wide_spec, s);  // de-emphasis
        if (rc == QC_OK) rc = fircore_wide(paud, m, ws, m, ws, g, d_wide_spec, s);      // audio filter, in place
        if (rc == QC_OK) rc = sntch->run(m, ws, m, ws, n, s);                           // CTCSS notch (I rail)
        if (rc == QC_OK) rc = fm_limiter(m, ws, n, s);
    }
    }
    if (rc == QC_OK && bp1_run) rc = fircore_wide(bp1, m, ws, m, ws, g, d_wide_spec, s);
    if (rc == QC_OK && agc_run) { rc = agc->run(m, ws, wmid2, ws, n, s); m = wmid2; }
    bool agc_meter_aside = false;
    if (rc == QC_OK && side && !rsmpout) {
        // the AGC meter only reads the AGC's output and state: beside the panel on the side stream (the panel writes elsewhere)
        QC_CUDA(cudaEventRecord(ev_a, s)); QC_CUDA(cudaStreamWaitEvent(side, ev_a, 0));
        rc = agcmeter->run(m, ws, agc->d_state, 0, n, side);
        QC_CUDA(cudaEventRecord(ev_b, side));
        agc_meter_aside = true;
    } else if (rc == QC_OK) rc = agcmeter->run(m, ws, agc->d_state, 0, n, s);
    for (SeqStage *mt : {adcmeter, smeter, agcmeter}) mt->meter_sub = 1;
    if (rc != QC_OK) return rc;
    cd *sp = sip_run ? d_sip : nullptr;
    const int sidx = sip_idx;
    if (sip_run) sip_idx = n >= sipsize ? 0 : (sip_idx + n) & (sipsize - 1);           // siphon.c:110-124
    if (rsmpout) {
        rc = launch_panel(m, ws, m, ws, n, C, panel_gain1 * panel_gain2I, panel_gain1 * panel_gain2Q, 3, 0, s, sp, sipsize, sidx); if (rc) return rc;
        int no = 0;
        rc = rsmpout->f->run(m, ws, n, dout, os, &no, 0, s); if (rc) return rc;
        if (no != g * dsp_outsize) { set_error("rxa: output resampler produced %d samples, expected %d", no, g * dsp_outsize); return QC_EINVAL; }
    } else {
        rc = launch_panel(m, ws, (cd *)dout, os, n, C, panel_gain1 * panel_gain2I, panel_gain1 * panel_gain2Q, 3, 0, s, sp, sipsize, sidx); if (rc) return rc;
    }
    if (agc_meter_aside) QC_CUDA(cudaStreamWaitEvent(s, ev_b, 0));      // the meter has read the scratch (the next group's stages write it) and the AGC state
    return QC_OK;
}

int Rxa::xrxa(const void *din, long is, void *dout, long os, cudaStream_t s)
{
    if (fusable()) return xrxa_fused(din, is, dout, os, 1, s);
    int rc;
    const cd *cur = (const cd *)din; long cs = is;
    const long ms = 2 * (dsp_size > dsp_insize ? dsp_size : dsp_insize) + 64;
    // xshift on inbuff (in place in the reference; a zero shift is the exact identity: phase stays 0)
    if (shift_run && shift_nonzero) {
        rc = shift->run(cur, cs, mid2, ms, dsp_insize, s); if (rc) return rc;
        cur = mid2; cs = ms;
    }
    // xresample(rsmpin): inbuff -> midbuff
    cd *m = mid;
    if (rsmpin) {
        int no = 0;
        rc = rsmpin->f->run(cur, cs, dsp_insize, m, ms, &no, 0, s); if (rc) return rc;
        if (no != dsp_size) { set_error("rxa: input resampler produced %d samples, dsp_size is %d", no, dsp_size); return QC_EINVAL; }
        rc = adcmeter->run(m, ms, nullptr, 0, dsp_size, s); if (rc) return rc;
        // xbpsnbain(0): the block as it enters nbp0 (snb.c:795-799)
        if (bps_run && bps_position == 0) QC_CUDA(cudaMemcpy2DAsync(bps_buff, (size_t)dsp_size * sizeof(cd), m, (size_t)ms * sizeof(cd), (size_t)dsp_size * sizeof(cd), C, cudaMemcpyDeviceToDevice, s));
        if (nbp_run) { rc = nbp0->run(m, ms, m, ms, s); if (rc) return rc; }
    } else {
        // no input resampler: the first stage that writes moves the block into midbuff, no separate copy
        rc = adcmeter->run(cur, cs, nullptr, 0, dsp_size, s); if (rc) return rc;
        if (bps_run && bps_position == 0) QC_CUDA(cudaMemcpy2DAsync(bps_buff, (size_t)dsp_size * sizeof(cd), cur, (size_t)cs * sizeof(cd), (size_t)dsp_size * sizeof(cd), C, cudaMemcpyDeviceToDevice, s));
        if (nbp_run) { rc = nbp0->run(cur, cs, m, ms, s); if (rc) return rc; }
        else QC_CUDA(cudaMemcpy2DAsync(m, (size_t)ms * sizeof(cd), cur, (size_t)cs * sizeof(cd), (size_t)dsp_size * sizeof(cd), C, cudaMemcpyDeviceToDevice, s));
    }
    rc = smeter->run(m, ms, nullptr, 0, dsp_size, s); if (rc) return rc;
    // xbpsnbaout(0): its own band pass of the saved block REPLACES nbp0's output (snb.c:801-805)
    if (bps_run && bps_position == 0) { rc = bpsnba->run(bps_buff, dsp_size, m, ms, s); if (rc) return rc; }
    if (amd_run) { rc = amd->run(m, ms, m, ms, dsp_size, s); if (rc) return rc; }
    if (fmd_run) {
        rc = fmpll->run(m, ms, audio, dsp_size, dsp_size, s); if (rc) return rc;       // pll -> audio
        rc = pde->run(audio, dsp_size, m, ms, s); if (rc) return rc;                   // de-emphasis: audio -> out
        rc = paud->run(m, ms, m, ms, s); if (rc) return rc;                            // audio filter, in place
        rc = sntch->run(m, ms, m, ms, dsp_size, s); if (rc) return rc;                 // CTCSS notch (I rail)
        rc = fm_limiter(m, ms, dsp_size, s); if (rc) return rc;
    }
    // xbpsnbain(1) + xbpsnbaout(1): behind the demodulators in the AM / FM modes (RXA.c:576-577)
    if (bps_run && bps_position == 1) {
        QC_CUDA(cudaMemcpy2DAsync(bps_buff, (size_t)dsp_size * sizeof(cd), m, (size_t)ms * sizeof(cd), (size_t)dsp_size * sizeof(cd), C, cudaMemcpyDeviceToDevice, s));
        rc = bpsnba->run(bps_buff, dsp_size, m, ms, s); if (rc) return rc;
    }
    if (snba_run) { rc = snba_run_stage(m, ms, s); if (rc) return rc; }
    // xemnr / xbandpass(bp1) at position 0 in front of the AGC, at position 1 behind it (RXA.c:577-590)
    if (emnr_run && emnr_position == 0) { rc = emnr_run_stage(m, ms, s); if (rc) return rc; }
    if (bp1_run && emnr_position == 0) { rc = bp1->run(m, ms, m, ms, s); if (rc) return rc; }
    // out of place into the second scratch buffer (free by now): the AGC kernel then needs no sample staging
    if (agc_run) { rc = agc->run(m, ms, mid2, ms, dsp_size, s); if (rc) return rc; m = mid2; }
    if (emnr_run && emnr_position == 1) { rc = emnr_run_stage(m, ms, s); if (rc) return rc; }
    if (bp1_run && emnr_position == 1) { rc = bp1->run(m, ms, m, ms, s); if (rc) return rc; }
    rc = agcmeter->run(m, ms, agc->d_state, 0, dsp_size, s); if (rc) return rc;
    // xpanel always applies gain1 * gain2 (F9), inselect 3, no copy
    cd *sp = sip_run ? d_sip : nullptr;
    const int sidx = sip_idx;
    if (sip_run && dsp_size < sipsize) sip_idx = (sip_idx + dsp_size) & (sipsize - 1);       // siphon.c:124
    if (rsmpout) {
        rc = launch_panel(m, ms, m, ms, dsp_size, C, panel_gain1 * panel_gain2I, panel_gain1 * panel_gain2Q, 3, 0, s, sp, sipsize, sidx); if (rc) return rc;
        int no = 0;
        rc = rsmpout->f->run(m, ms, dsp_size, dout, os, &no, 0, s); if (rc) return rc;
        if (no != dsp_outsize) { set_error("rxa: output resampler produced %d samples, expected %d", no, dsp_outsize); return QC_EINVAL; }
    } else {
        rc = launch_panel(m, ms, (cd *)dout, os, dsp_size, C, panel_gain1 * panel_gain2I, panel_gain1 * panel_gain2Q, 3, 0, s, sp, sipsize, sidx); if (rc) return rc;
    }
    return QC_OK;
}

// ---- fexchange0 / dexchange / slews / SetChannelState (iobuffs.c, channel.c:262-300) --------------------------------
enum { D_BEGIN = 0, D_DELAYDOWN = 4, D_DOWNSLEW = 5, D_ZERO = 6, D_OFF = 7 };

__global__ void downslew_kernel(const cd *r2, long r2_stride, cd *out, long out_stride, int n, int C, const double *gain)
{
    const long total = (long)n * C;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int c = (int)(t / n), i = (int)(t - (long)c * n);
        const cd v = r2[(size_t)c * r2_stride + i];
        const double g = gain[i];
        out[(size_t)c * out_stride + i] = g == 0.0 ? make_double2(0.0, 0.0) : make_double2(v.x * g, v.y * g);
    }
}

int Rxa::setup_exchange()
{   // create_iobuffs + create_slews (iobuffs.c:47-80, 385-420)
    r1_size = dsp_insize > in_size ? dsp_insize : in_size;
    r2_size = out_size > dsp_outsize ? out_size : dsp_outsize;
    r1_active = 2 * r1_size; r2_active = 2 * r2_size;          // DSP_MULT = 2 (comm.h:118)
    if (in_size <= 0 || out_size <= 0 || r1_active % in_size || r1_active % dsp_insize || r2_active % out_size || r2_active % dsp_outsize) {
        set_error("rxa: in_size %d / dsp_insize %d / out_size %d / dsp_outsize %d do not tile the exchange rings", in_size, dsp_insize, out_size, dsp_outsize);
        return QC_EINVAL;
    }
    if (!hs) QC_CUDA(cudaStreamCreateWithFlags(&hs, cudaStreamNonBlocking));
    for (cd **q : {&d_r1, &d_r2, &d_outbuff, &d_xout}) if (*q) { cudaFree(*q); *q = nullptr; }
    if (d_gain) { cudaFree(d_gain); d_gain = nullptr; }
    QC_CUDA(cudaMalloc((void **)&d_r1, (size_t)C * r1_active * sizeof(cd)));
    QC_CUDA(cudaMalloc((void **)&d_r2, (size_t)C * r2_active * sizeof(cd)));
    QC_CUDA(cudaMalloc((void **)&d_outbuff, (size_t)C * dsp_outsize * sizeof(cd)));
    QC_CUDA(cudaMalloc((void **)&d_xout, (size_t)C * out_size * sizeof(cd)));
    QC_CUDA(cudaMalloc((void **)&d_gain, (size_t)out_size * sizeof(double)));
    QC_CUDA(cudaMemset(d_outbuff, 0, (size_t)C * dsp_outsize * sizeof(cd)));
    ndeldown = (int)(tdelaydown * out_rate);
    ntdown = (int)(tslewdown * out_rate);
    cdown.assign((size_t)ntdown + 1, 0.0);
    const double delta = 3.1415926535897932 / (double)ntdown;
    double theta = 0.0;
    for (int i = 0; i <= ntdown; i++) { cdown[i] = 0.5 * (1 + cos(theta)); theta += delta; }
    int rc = flush_iobuffs(); if (rc != QC_OK) return rc;
    return state ? raise_upflag(*this) : QC_OK;                     // OpenChannel, channel.c:93-99
}

int Rxa::flush_iobuffs()
{   // flush_iobuffs (iobuffs.c:442-461) + flush_slews
    QC_CUDA(cudaMemset(d_r1, 0, (size_t)C * r1_active * sizeof(cd)));
    QC_CUDA(cudaMemset(d_r2, 0, (size_t)C * r2_active * sizeof(cd)));
    r1_inidx = r1_outidx = r1_unq = 0;
    r2_inidx = r2_size; r2_outidx = 0; r2_havesamps = r2_size;      // (DSP_MULT - 1) * r2_size
    out_credits = r2_havesamps / out_size;
    r2_unq = r2_havesamps - out_credits * out_size;
    dstate = D_BEGIN; dcount = 0; downflag = 0;
    // up-slew: states back to BEGIN, flags down (SetChannelState(1) / OpenChannel raise them)
    if (d_uslew) QC_CUDA(cudaMemset(d_uslew, 0, (size_t)C * 3 * sizeof(int)));
    return QC_OK;
}

int Rxa::flush_main()
{   // flush_rxa, RXA.c:527-559: inbuff / outbuff / midbuff zeroed, then every stage's own flush
    QC_CUDA(cudaDeviceSynchronize());
    QC_CUDA(cudaMemset(d_outbuff, 0, (size_t)C * dsp_outsize * sizeof(cd)));
    int rc;
    for (SeqStage *q : {shift, adcmeter, smeter, amd, fmpll, sntch, plim, agc, agcmeter}) if (q) { rc = q->flush_ref(); if (rc) return rc; }
    for (FirCore *f : {nbp0, pde, paud, bp1}) if (f) { rc = f->flush(); if (rc) return rc; }
    if (emnr) { rc = emnr_flush(emnr); if (rc) return rc; }
    if (snba) { rc = snba_flush(snba); if (rc) return rc; }
    if (bpsnba) { rc = bpsnba->flush(); if (rc) return rc; QC_CUDA(cudaMemset(bps_buff, 0, (size_t)C * dsp_size * sizeof(cd))); }
    for (Resampler *q : {rsmpin, rsmpout}) if (q) { rc = q->f->reset(nullptr); if (rc) return rc; }
    QC_CUDA(cudaMemset(d_sip, 0, (size_t)C * sipsize * sizeof(cd)));
    sip_idx = 0;
    QC_CUDA(cudaDeviceSynchronize());
    return QC_OK;
}

int Rxa::set_channel_state(int new_state, int dmode)
{   // SetChannelState, channel.c:262-300.  Returns the prior state.  dmode = 1 waits for the down-slew to finish in the
    // reference (another thread keeps calling fexchange0) and gives up after 100 ms; a single-threaded caller always
    // takes that give-up branch there: exchange off, no flush.  Here the exchange calls are synchronous, so with
    // dmode = 1 the same give-up branch is taken at once.
    const int prior = state;
    if (state == new_state) return prior;
    state = new_state;
    if (new_state == 0) {
        downflag = 1; flushflag = 1;
        if (dmode) { exchange_on = 0; flushflag = 0; downflag = 0; }
    } else {
        // upflag up; ustate / ucount stay where flush_slews (or the last finished ramp) left them: BEGIN
        int rc = raise_upflag(*this); if (rc != QC_OK) return rc;
        exchange_on = 1;
    }
    return prior;
}

int Rxa::exchange(const double *h_in, double *h_out, int *error)
{
    if (error) *error = 0;
    if (!exchange_on) return QC_OK;                                // iobuffs.c:471: nothing moves, `out` is left alone
    if (!d_r1) { int rc = setup_exchange(); if (rc) return rc; }
    // in -> r1 (through upslew0 while a channel's upflag is up; the kernel returns at once for the others)
    cd *slot = d_r1 + r1_inidx;
    QC_CUDA(cudaMemcpy2DAsync(slot, (size_t)r1_active * sizeof(cd), h_in, (size_t)in_size * sizeof(cd), (size_t)in_size * sizeof(cd), C,
                              cudaMemcpyHostToDevice, hs));
    upslew_kernel<<<(C + 63) / 64, 64, 0, hs>>>(slot, r1_active, in_size, C, d_uslew, d_cup, ndelup, ntup);
    count_launch();
    QC_CUDA_LAUNCH();
    int n_dsp = 0;
    if ((r1_unq += in_size) >= dsp_insize) { n_dsp = r1_unq / dsp_insize; r1_unq -= n_dsp * dsp_insize; }
    if ((r1_inidx += in_size) == r1_active) r1_inidx = 0;
    // the DSP thread's turns (wdspmain, main.c:40-63): dexchange pushes the PREVIOUS outbuff into r2 and pulls the next
    // dsp_insize samples out of r1, then xrxa runs
    for (int k = 0; k < n_dsp; k++) {
        r2_havesamps += dsp_outsize;
        QC_CUDA(cudaMemcpy2DAsync(d_r2 + r2_inidx, (size_t)r2_active * sizeof(cd), d_outbuff, (size_t)dsp_outsize * sizeof(cd),
                                  (size_t)dsp_outsize * sizeof(cd), C, cudaMemcpyDeviceToDevice, hs));
        if ((r2_inidx += dsp_outsize) == r2_active) r2_inidx = 0;
        if (bfo && (r2_unq += dsp_outsize) >= out_size) { const int n = r2_unq / out_size; out_credits += n; r2_unq -= n * out_size; }
        int rc = xrxa(d_r1 + r1_outidx, r1_active, d_outbuff, dsp_outsize, hs); if (rc) return rc;
        if ((r1_outidx += dsp_insize) == r1_active) r1_outidx = 0;
    }
    const int doit = r2_havesamps >= out_size;
    if ((r2_havesamps -= out_size) < 0) r2_havesamps = 0;
    bool have = doit;
    if (bfo) {      // WaitForSingleObject(Sem_OutReady): the DSP turns above have already run, so a missing credit can never arrive
        have = out_credits > 0;
        if (have) out_credits--;
    }
    const size_t nout = (size_t)C * out_size;
    if (have) {
        if (downflag) {
            std::vector<double> g((size_t)out_size);
            for (int i = 0; i < out_size; i++) {                    // downslew0, iobuffs.c:226-300
                switch (dstate) {
                case D_BEGIN:
                    g[i] = 1.0;
                    if (ndeldown > 0) { dstate = D_DELAYDOWN; dcount = ndeldown; }
                    else if (ntdown > 0) { dstate = D_DOWNSLEW; dcount = ntdown; }
                    else { dstate = D_ZERO; dcount = out_size; }
                    break;
                case D_DELAYDOWN:
                    g[i] = 1.0;
                    if (dcount-- == 0) {
                        if (ntdown > 0) { dstate = D_DOWNSLEW; dcount = ntdown; }
                        else { dstate = D_ZERO; dcount = out_size; }
                    }
                    break;
                case D_DOWNSLEW:
                    g[i] = cdown[(size_t)(ntdown - dcount)];
                    if (dcount-- == 0) { dstate = D_ZERO; dcount = out_size; }
                    break;
                case D_ZERO:
                    g[i] = 0.0;
                    if (dcount-- == 0) dstate = D_OFF;
                    break;
                default:
                    g[i] = 0.0;
                    if (i == out_size - 1) { dstate = D_BEGIN; downflag = 0; }
                    break;
                }
            }
            QC_CUDA(cudaMemcpyAsync(d_gain, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice, hs));
            const long total = (long)nout;
            downslew_kernel<<<(int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0, hs>>>(d_r2 + r2_outidx, r2_active, d_xout, out_size, out_size, C, d_gain);
            count_launch();
            QC_CUDA_LAUNCH();
            QC_CUDA(cudaMemcpyAsync(h_out, d_xout, nout * sizeof(cd), cudaMemcpyDeviceToHost, hs));
            QC_CUDA(cudaStreamSynchronize(hs));                     // g goes out of scope
            if (!downflag) {                                        // ramp finished: exchange off, the flush thread's work (channel.c:134-155)
                exchange_on = 0;
                if ((r2_outidx += out_size) == r2_active) r2_outidx = 0;
                int rc = flush_iobuffs(); if (rc) return rc;
                rc = flush_main(); if (rc) return rc;
                flushflag = 0;
                return QC_OK;
            }
        } else {
            QC_CUDA(cudaMemcpy2DAsync(h_out, (size_t)out_size * sizeof(cd), d_r2 + r2_outidx, (size_t)r2_active * sizeof(cd),
                                      (size_t)out_size * sizeof(cd), C, cudaMemcpyDeviceToHost, hs));
        }
    } else {
        memset(h_out, 0, nout * sizeof(cd));
        if (error) *error += -2;
    }
    if ((r2_outidx += out_size) == r2_active) r2_outidx = 0;
    QC_CUDA(cudaStreamSynchronize(hs));
    return QC_OK;
}

}  // namespace qc

using namespace qc;

extern "C" {

qcRxa *quisk_cuda_rxa_create(int n_channels, int in_size, int dsp_size, int in_rate, int dsp_rate, int out_rate)
{
    if (ensure_device() != QC_OK) return nullptr;
    qcRxa *p = new qcRxa();
    if (p->r.init(n_channels, in_size, dsp_size, in_rate, dsp_rate, out_rate) != QC_OK) { p->r.release(); delete p; return nullptr; }
    return p;
}

void quisk_cuda_rxa_destroy(qcRxa *r) { if (r) { r->r.release(); delete r; } }
int quisk_cuda_rxa_in_size(const qcRxa *r) { return r ? r->r.dsp_insize : QC_EINVAL; }
int quisk_cuda_rxa_out_size(const qcRxa *r) { return r ? r->r.dsp_outsize : QC_EINVAL; }

int quisk_cuda_rxa_set_mode(qcRxa *p, int mode)
{   // SetRXAMode, RXA.c:749-787
    if (!p) return QC_EINVAL;
    Rxa &r = p->r;
    if (r.mode == mode) return QC_OK;
    { int rcb = r.bpsnba_check(mode, r.ndb_run); if (rcb) return rcb; }        // RXA.c:754
    r.mode = mode;
    r.amd_run = 0; r.fmd_run = 0; r.agc_run = 1;
    if (mode == QC_RXA_AM) { r.amd_run = 1; r.amd->par[0] = 0; }
    else if (mode == QC_RXA_SAM) { r.amd_run = 1; r.amd->par[0] = 1; }
    else if (mode == QC_RXA_FM) { r.fmd_run = 1; r.agc_run = 0; }
    { int rcb = r.bpsnba_set(); if (rcb) return rcb; }                          // RXA.c:784
    return r.bp1_check_set();       // RXAbp1Check + RXAbp1Set, RXA.c:800-827
}

int quisk_cuda_rxa_set_passband(qcRxa *p, double f_low, double f_high)
{   // RXASetPassband = SetRXABandpassFreqs (bandpass.c:393-409) + RXANBPSetFreqs (nbp.c:528-540)
    if (!p) return QC_EINVAL;
    Rxa &r = p->r;
    if (f_low != r.bp1_flow || f_high != r.bp1_fhigh) {
        std::vector<double> imp((size_t)2 * r.bp1_nc);
        quisk_cuda_fir_bandpass(r.bp1_nc, f_low, f_high, (double)r.dsp_rate, 1, 1, r.bp1_gain / (double)(2 * r.dsp_size), imp.data());
        int rc = r.bp1->set_impulse(imp.data(), 0); if (rc) return rc;
        r.bp1_flow = f_low; r.bp1_fhigh = f_high;
        r.bp1->update();
    }
    if (r.snba) { int rcs = qc::snba_set_output_bandwidth(r.snba, f_low, f_high); if (rcs) return rcs; }     // SetRXASNBAOutputBandwidth, RXA.c:930
    if (f_low != r.nbp_flow || f_high != r.nbp_fhigh) {
        r.nbp_flow = f_low; r.nbp_fhigh = f_high;
        std::vector<double> imp;
        int rc = r.nbp0_impulse(imp, &r.nbp_hadnotch); if (rc) return rc;
        rc = r.nbp0->set_impulse(imp.data(), 1); if (rc) return rc;
    }
    return QC_OK;
}

int quisk_cuda_rxa_set_nc(qcRxa *p, int nc)
{   // RXASetNC, RXA.c:935-946: every fircore is re-planned (fresh, zeroed state) with nc taps
    if (!p) return QC_EINVAL;
    Rxa &r = p->r;
    if (nc < r.dsp_size || nc % r.dsp_size) { set_error("rxa_set_nc: nc must be a multiple of dsp_size"); return QC_EINVAL; }
    int rc;
    if (r.nbp_nc != nc) { r.nbp_nc = nc; if ((rc = r.make_nbp0()) != QC_OK) return rc; }
    if (r.bps_nc != nc) { r.bps_nc = nc; if ((rc = r.make_bpsnba()) != QC_OK) return rc; }      // RXABPSNBASetNC, snb.c:829-842
    if (r.bp1_nc != nc) { r.bp1_nc = nc; if ((rc = r.make_bp1()) != QC_OK) return rc; }
    if (r.fm_nc_de != nc || r.fm_nc_aud != nc) { r.fm_nc_de = r.fm_nc_aud = nc; if ((rc = r.make_fmd()) != QC_OK) return rc; }
    return QC_OK;
}

int quisk_cuda_rxa_set_snba_run(qcRxa *p, int run)
{   // SetRXASNBARun, snb.c:579-593
    if (!p) return QC_EINVAL;
    qc::Rxa &r = p->r;
    run = run ? 1 : 0;
    if (r.snba_run == run) return QC_OK;
    if (run && !r.snba) { qc::set_error("SetRXASNBARun: the noise blanker needs dsp_rate = 12000 x an integer and a dsp_size that divides accordingly"); return QC_EINVAL; }
    int rc = r.bpsnba_check(r.mode, r.ndb_run); if (rc) return rc;
    r.snba_run = run;
    // RXAbp1Check before the flag flips in the reference, RXAbp1Set after: the helper does both from the new flags
    rc = r.bp1_check_set(); if (rc) return rc;
    return r.bpsnba_set();
}

int quisk_cuda_rxa_set_emnr_run(qcRxa *p, int run)
{   // SetRXAEMNRRun, emnr.c:1096-1109
    if (!p) return QC_EINVAL;
    qc::Rxa &r = p->r;
    run = run ? 1 : 0;
    if (r.emnr_run == run) return QC_OK;
    if (run && r.emnr_gain_method == 2 && !qc::emnr_tables_present()) {
        qc::set_error("SetRXAEMNRRun: gain method 2 needs the WDSP distribution's two gamma-prior tables: quisk_cuda_emnr_set_tables first (or choose gain method 0 or 1)");
        return QC_EINVAL;
    }
    if (run && r.emnr_gain_method == 3 && !qc::emnr_zeta_present()) {
        qc::set_error("SetRXAEMNRRun: gain method 3 needs the WDSP distribution's trained zeta table: quisk_cuda_emnr_set_zeta first (or choose gain method 0 or 1)");
        return QC_EINVAL;
    }
    if (run && (r.emnr_gain_method < 0 || r.emnr_gain_method > 3)) { qc::set_error("SetRXAEMNRRun: there is no gain method %d", r.emnr_gain_method); return QC_EINVAL; }
    r.emnr_run = run;
    return r.bp1_check_set();
}
int quisk_cuda_rxa_set_emnr_gain_method(qcRxa *p, int method) { if (!p) return QC_EINVAL; p->r.emnr_gain_method = method; return qc::emnr_set(p->r.emnr, 0, method); }
int quisk_cuda_rxa_set_emnr_npe_method(qcRxa *p, int method) { return p ? qc::emnr_set(p->r.emnr, 1, method) : QC_EINVAL; }
int quisk_cuda_rxa_set_emnr_ae_run(qcRxa *p, int run) { return p ? qc::emnr_set(p->r.emnr, 2, run) : QC_EINVAL; }
int quisk_cuda_rxa_set_emnr_train(qcRxa *p, int what, double value) { return p ? qc::emnr_set_train(p->r.emnr, what, value) : QC_EINVAL; }   /* SetRXAEMNRtrainZetaThresh (0), SetRXAEMNRtrainT2 (1) */
int quisk_cuda_rxa_set_emnr_position(qcRxa *p, int position) { if (!p) return QC_EINVAL; p->r.emnr_position = position ? 1 : 0; return QC_OK; }    /* SetRXAEMNRPosition moves bp1 with it */
int quisk_cuda_rxa_set_fm_lim_run(qcRxa *p, int run) { if (!p) return QC_EINVAL; p->r.lim_run = run ? 1 : 0; return QC_OK; }      /* SetRXAFMLimRun, fmd.c:337-348 */
int quisk_cuda_rxa_set_fm_lim_gain(qcRxa *p, double gain_db)
{   /* SetRXAFMLimGain, fmd.c:350-363: decalc_fmd / calc_fmd rebuild the limiter (and the notch) from scratch */
    if (!p) return QC_EINVAL;
    qc::Rxa &r = p->r;
    const double g = pow(10.0, gain_db / 20.0);
    if (g == r.lim_gain) return QC_OK;
    if (cudaDeviceSynchronize() != cudaSuccess) return QC_ECUDA;
    qc::SeqStage *nl = qc::make_wcpagc_fmlim(r.C, r.dsp_rate, g);
    if (!nl) return QC_ENOMEM;
    if (r.plim) { r.plim->release(); delete r.plim; }
    r.plim = nl; r.lim_gain = g;
    r.fmpll->flush(); r.sntch->flush();                 /* calc_fmd zeroes the PLL state and makes a new notch */
    return QC_OK;
}
int quisk_cuda_rxa_set_agc_mode(qcRxa *p, int mode) { if (!p) return QC_EINVAL; agc_set_mode(p->r.agc, mode); return QC_OK; }
int quisk_cuda_rxa_set_agc_fixed(qcRxa *p, double gain_db)
{ if (!p) return QC_EINVAL; p->r.agc->agc.fixed_gain = pow(10.0, gain_db / 20.0); p->r.agc->load_agc(); return QC_OK; }

int quisk_cuda_rxa_set_shift(qcRxa *p, int run, const double *shift_hz)
{
    if (!p) return QC_EINVAL;
    Rxa &r = p->r;
    r.shift_run = run;
    if (shift_hz) {
        if (r.shift) { r.shift->release(); delete r.shift; }
        r.shift = make_shift(r.C, r.in_rate, shift_hz);
        if (!r.shift) return QC_ENOMEM;
        r.shift_nonzero = false;
        for (int c = 0; c < r.C; c++) r.shift_nonzero = r.shift_nonzero || shift_hz[c] != 0.0;
    }
    return QC_OK;
}

// ---- notch database (nbp.c:336-513) ----
static int nbp_update(Rxa &r, bool lightweight)
{   // UpdateNBPFilters (always recompute when the notches run) / UpdateNBPFiltersLightWeight (tune or shift moved:
    // recompute only if there were or are notches inside the pass band, nbp.c:181-212)
    if (r.bps_run_notches) {        // the band pass of the noise blanker follows the same database (nbp.c:339, 345-355)
        std::vector<double> bi;
        int hb = 0;
        int rcb = r.bpsnba_impulse(bi, &hb); if (rcb != QC_OK) return rcb;
        if (!lightweight || r.bps_hadnotch || hb) { rcb = r.bpsnba->set_impulse(bi.data(), 1); if (rcb != QC_OK) return rcb; }
        r.bps_hadnotch = hb;
    } else if (lightweight) r.bps_hadnotch = 1;
    if (!r.ndb_run) { if (lightweight) r.nbp_hadnotch = 1; return QC_OK; }
    std::vector<double> imp;
    int hav = 0;
    int rc = r.nbp0_impulse(imp, &hav); if (rc != QC_OK) return rc;
    if (!lightweight || r.nbp_hadnotch || hav) { rc = r.nbp0->set_impulse(imp.data(), 1); if (rc != QC_OK) return rc; }
    r.nbp_hadnotch = hav;
    return QC_OK;
}

int quisk_cuda_rxa_nbp_add_notch(qcRxa *p, int notch, double fcenter, double fwidth, int active)
{   // RXANBPAddNotch, nbp.c:359-387
    if (!p) return QC_EINVAL;
    Rxa &r = p->r;
    const int nn = (int)r.ndb_fcenter.size();
    if (notch < 0 || notch > nn || nn >= 1024) return -1;
    r.ndb_fcenter.insert(r.ndb_fcenter.begin() + notch, fcenter);
    r.ndb_fwidth.insert(r.ndb_fwidth.begin() + notch, fwidth);
    r.ndb_active.insert(r.ndb_active.begin() + notch, active);
    return nbp_update(r, false);
}

int quisk_cuda_rxa_nbp_delete_notch(qcRxa *p, int notch)
{   // RXANBPDeleteNotch, nbp.c:414-437
    if (!p) return QC_EINVAL;
    Rxa &r = p->r;
    if (notch < 0 || notch >= (int)r.ndb_fcenter.size()) return -1;
    r.ndb_fcenter.erase(r.ndb_fcenter.begin() + notch);
    r.ndb_fwidth.erase(r.ndb_fwidth.begin() + notch);
    r.ndb_active.erase(r.ndb_active.begin() + notch);
    return nbp_update(r, false);
}

int quisk_cuda_rxa_nbp_set_notches_run(qcRxa *p, int run)
{   // RXANBPSetNotchesRun, nbp.c:496-513
    if (!p) return QC_EINVAL;
    Rxa &r = p->r;
    run = run ? 1 : 0;
    if (run == r.ndb_run) return QC_OK;
    { int rcb = r.bpsnba_check(r.mode, run); if (rcb) return rcb; }            // nbp.c:504
    r.ndb_run = run;
    std::vector<double> imp;
    int rc = r.nbp0_impulse(imp, &r.nbp_hadnotch); if (rc != QC_OK) return rc;
    rc = r.nbp0->set_impulse(imp.data(), 1); if (rc != QC_OK) return rc;
    return r.bpsnba_set();                                                         // nbp.c:509
}

int quisk_cuda_rxa_nbp_set_tune_frequency(qcRxa *p, double tunefreq)
{   // RXANBPSetTuneFrequency, nbp.c:472-481
    if (!p) return QC_EINVAL;
    if (tunefreq == p->r.ndb_tune) return QC_OK;
    p->r.ndb_tune = tunefreq;
    return nbp_update(p->r, true);
}

int quisk_cuda_rxa_nbp_set_shift_frequency(qcRxa *p, double shift)
{   // RXANBPSetShiftFrequency, nbp.c:484-493
    if (!p) return QC_EINVAL;
    if (shift == p->r.ndb_shift) return QC_OK;
    p->r.ndb_shift = shift;
    return nbp_update(p->r, true);
}

int quisk_cuda_rxa_set_nbp_run(qcRxa *p, int run) { if (!p) return QC_EINVAL; p->r.nbp_run = run; return QC_OK; }
// suck + the float conversion of RXAGetaSipF / RXAGetaSipF1 (siphon.c:148-163, 183-211)
__global__ void siphon_get_kernel(const cd *sip, int sipsize, int idx, int size, int C, int cpx, float *out)
{
    const long total = (long)size * C;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int c = (int)(t / size), i = (int)(t - (long)c * size);
        const cd v = sip[(size_t)c * sipsize + ((idx - size + i) & (sipsize - 1))];
        if (cpx) { out[2 * t] = (float)v.x; out[2 * t + 1] = (float)v.y; }
        else out[t] = (float)v.x;
    }
}

int quisk_cuda_rxa_set_siphon_run(qcRxa *p, int run) { if (!p) return QC_EINVAL; p->r.sip_run = run ? 1 : 0; return QC_OK; }

int quisk_cuda_rxa_get_siphon(qcRxa *p, float *h_out, int size, int complex_out)
{
    if (!p || !h_out) return QC_EINVAL;
    Rxa &r = p->r;
    if (size <= 0 || size > r.sipsize) { set_error("rxa_get_siphon: size must be in [1, %d]", r.sipsize); return QC_EINVAL; }
    const size_t nfl = (size_t)r.C * size * (complex_out ? 2 : 1);
    if ((int)nfl > r.sipout_cap) {
        if (r.d_sipout) cudaFree(r.d_sipout);
        r.d_sipout = nullptr; r.sipout_cap = 0;
        QC_CUDA(cudaMalloc((void **)&r.d_sipout, nfl * sizeof(float)));
        r.sipout_cap = (int)nfl;
    }
    QC_CUDA(cudaDeviceSynchronize());
    siphon_get_kernel<<<(int)((nfl + 255) / 256 < 1184 ? (nfl + 255) / 256 : 1184), 256>>>(r.d_sip, r.sipsize, r.sip_idx, size, r.C, complex_out ? 1 : 0, r.d_sipout);
    count_launch();
    QC_CUDA_LAUNCH();
    QC_CUDA(cudaMemcpy(h_out, r.d_sipout, nfl * sizeof(float), cudaMemcpyDeviceToHost));
    return QC_OK;
}

int quisk_cuda_rxa_set_mp(qcRxa *p, int mp)
{   // RXASetMP, RXA.c:949-958, for the fircores this chain has: nbp0, bp1, the FM de-emphasis and audio filters
    if (!p) return QC_EINVAL;
    for (FirCore *f : {p->r.nbp0, p->r.bp1, p->r.pde, p->r.paud}) if (f) { int rc = f->set_mp(mp); if (rc != QC_OK) return rc; }
    return QC_OK;
}

int quisk_cuda_rxa_set_slew(qcRxa *p, double tdelayup, double tslewup)
{ if (!p || tdelayup < 0 || tslewup < 0) return QC_EINVAL; return p->r.arm_upslew(tdelayup, tslewup); }

int quisk_cuda_rxa_set_panel_gain(qcRxa *p, double g) { if (!p) return QC_EINVAL; p->r.panel_gain1 = g; return QC_OK; }

int quisk_cuda_rxa_xrxa(qcRxa *p, const void *d_in, long in_stride, void *d_out, long out_stride, void *stream)
{ return p ? p->r.xrxa(d_in, in_stride, d_out, out_stride, (cudaStream_t)stream) : QC_EINVAL; }
int quisk_cuda_rxa_xrxa_multi(qcRxa *p, const void *d_in, long in_stride, void *d_out, long out_stride, int n_blocks, void *stream)
{ return p ? p->r.xrxa_multi(d_in, in_stride, d_out, out_stride, n_blocks, (cudaStream_t)stream) : QC_EINVAL; }
int quisk_cuda_rxa_set_option(qcRxa *p, int option, int value)
{
    if (!p) return QC_EINVAL;
    if (option == QC_RXA_OPT_FUSED) { p->r.fused_ok = value ? 1 : 0; return QC_OK; }
    set_error("rxa_set_option: unknown option %d", option);
    return QC_EINVAL;
}

int quisk_cuda_rxa_fexchange0(qcRxa *p, const double *h_in, double *h_out, int *error)
{   // fexchange0 (iobuffs.c:464-516) for every channel of the batch: h_in [C][in_size], h_out [C][out_size]
    if (!p) return QC_EINVAL;
    return p->r.exchange(h_in, h_out, error);
}

int quisk_cuda_rxa_set_channel_state(qcRxa *p, int state, int dmode)
{ return p ? p->r.set_channel_state(state, dmode) : QC_EINVAL; }

int quisk_cuda_rxa_set_slew_down(qcRxa *p, double tdelaydown, double tslewdown)
{
    if (!p || tdelaydown < 0 || tslewdown < 0) return QC_EINVAL;
    p->r.tdelaydown = tdelaydown; p->r.tslewdown = tslewdown;
    return p->r.setup_exchange();
}

int quisk_cuda_rxa_set_bfo(qcRxa *p, int bfo) { if (!p) return QC_EINVAL; p->r.bfo = bfo ? 1 : 0; return QC_OK; }
int quisk_cuda_rxa_exchange_sizes(const qcRxa *p, int *in_size, int *out_size)
{ if (!p) return QC_EINVAL; if (in_size) *in_size = p->r.in_size; if (out_size) *out_size = p->r.out_size; return QC_OK; }

int quisk_cuda_rxa_get_meter(qcRxa *p, int which, double *av, double *pk, double *gain)
{
    if (!p) return QC_EINVAL;
    SeqStage *m = which == 0 ? p->r.adcmeter : (which == 1 ? p->r.smeter : p->r.agcmeter);
    std::vector<double> h((size_t)p->r.C * 3);
    QC_CUDA(cudaDeviceSynchronize());
    QC_CUDA(cudaMemcpy(h.data(), m->d_meter, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
    for (int c = 0; c < p->r.C; c++) {
        if (av) av[c] = h[(size_t)c * 3];
        if (pk) pk[c] = h[(size_t)c * 3 + 1];
        if (gain) gain[c] = h[(size_t)c * 3 + 2];
    }
    return QC_OK;
}

}  // extern "C"
