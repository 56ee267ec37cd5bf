// quisk_b200/csrc/rxchain.cu -- quisk_cuda_rx_*: the receive chain of quisk_process_samples
// for a batch of channels: tune (quisk.c:2477-2488) -> quisk_process_decimate
// (quisk.c:1673-1846, planned by PlanDecimation quisk.c:1633-1671) ->
// quisk_process_demodulate (quisk.c:1848-2160).
//
// Two executions of the same stage graph:
//   fused = 0  one exact kernel per stage (polyfir.cu / pointwise.cu), intermediates in HBM;
//   fused = 1  the shared-memory cascade of rxfused.cu for the full-rate decimator
//              (one CTA streams one channel through every half band and FIR without
//              touching HBM in between), exact per-stage kernels for the 48 kS/s tail.
#include "qc_common.cuh"
#include "batch.h"
#include "rxchain.h"
#include <cmath>

namespace qc {

int plan_decimation(int sample_rate, int *p2, int *p3, int *p5)
{   // PlanDecimation, quisk.c:1633-1671
    int best = sample_rate, d2 = 0, d3 = 0, d5 = 0;
    for (int i2 = 0; i2 <= 6; i2++)
        for (int i3 = 0; i3 <= 3; i3++)
            for (int i5 = 0; i5 <= 3; i5++) {
                int t = sample_rate;
                for (int i = 0; i < i2; i++) t /= 2;
                for (int i = 0; i < i3; i++) t /= 3;
                for (int i = 0; i < i5; i++) t /= 5;
                if (t >= 48000 && t < best) { d2 = i2; d3 = i3; d5 = i5; best = t; }
            }
    if (best >= 50000) best = best * 24 / 25;
    if (p2) { *p2 = d2; *p3 = d3; *p5 = d5; }
    return best;
}

static BatchFilter *mk(int kind, int C, const double *coefs, int n, int interp, int decim)
{
    if (kind != QC_C_DECIM2_HB45 && kind != QC_C_INTERP2_HB45 && kind != QC_D_INTERP2_HB45 && (coefs == nullptr || n <= 0)) {
        set_error("rx_create: a coefficient table this sample rate / mode needs was not supplied");
        return nullptr;
    }
    BatchFilter *f = new BatchFilter();
    if (f->init(kind, C, coefs, n, interp, decim) != QC_OK) { f->release(); delete f; return nullptr; }
    return f;
}

#define ADD(vec, expr) do { BatchFilter *_f = (expr); if (!_f) return QC_EINVAL; (vec).push_back(_f); } while (0)

int RxChain::init(const qcRxConfig &cfg)
{
    filter_bandwidth = cfg.filter_bandwidth;
    C = cfg.n_channels; sample_rate = cfg.sample_rate; mode = cfg.mode; fused = cfg.fused;
    if (C <= 0 || sample_rate <= 0) { set_error("rx_create: bad channel count / sample rate"); return QC_EINVAL; }
    {   // keep what we were given: the pipelined host entries build channel-subset chains from it on first use
        saved.cfg = cfg;
        if (cfg.filt_i && cfg.n_filt > 0) saved.fi.assign(cfg.filt_i, cfg.filt_i + cfg.n_filt);
        if (cfg.filt_q && cfg.n_filt > 0) saved.fq.assign(cfg.filt_q, cfg.filt_q + cfg.n_filt);
        if (cfg.tune_hz) saved.tune.assign(cfg.tune_hz, cfg.tune_hz + C);
        const double *const *tp = &cfg.tables.filt144D3;       // 13 (pointer, count) pairs in declaration order
        const qcRxTables &T0 = cfg.tables;
        const double *ptrs[13] = {T0.filt144D3, T0.filt240D5Sharp, T0.filt48dec24, T0.filt300D5, T0.audio24p4, T0.audio24p6, T0.lpFilt48,
                                  T0.audioFmHp, T0.filt53D1, T0.filt111D2, T0.filt133D2, T0.filt167D3, T0.filt185D3};
        const int cnts[13] = {T0.n_filt144D3, T0.n_filt240D5Sharp, T0.n_filt48dec24, T0.n_filt300D5, T0.n_audio24p4, T0.n_audio24p6, T0.n_lpFilt48,
                              T0.n_audioFmHp, T0.n_filt53D1, T0.n_filt111D2, T0.n_filt133D2, T0.n_filt167D3, T0.n_filt185D3};
        (void)tp;
        for (int i = 0; i < 13; i++) if (ptrs[i] && cnts[i] > 0) saved.tab[i].assign(ptrs[i], ptrs[i] + cnts[i]);
    }
    const qcRxTables &T = cfg.tables;
    // ---- quisk_process_decimate -------------------------------------------------
    int rate = sample_rate;
    switch ((sample_rate + 100) / 1000) {       // quisk.c:1731
    case 41: rate = 48000; break;
    case 53: ADD(cst, mk(QC_C_DECIMATE, C, T.filt53D1, T.n_filt53D1, 1, 1)); break;
    case 111: ADD(cst, mk(QC_C_DECIMATE, C, T.filt111D2, T.n_filt111D2, 1, 2)); rate /= 2; break;
    case 133: ADD(cst, mk(QC_C_DECIMATE, C, T.filt133D2, T.n_filt133D2, 1, 2)); rate /= 2; break;
    case 185: ADD(cst, mk(QC_C_DECIMATE, C, T.filt185D3, T.n_filt185D3, 1, 3)); rate /= 3; break;
    case 370:
        ADD(cst, mk(QC_C_DECIM2_HB45, C, nullptr, 0, 1, 1));
        ADD(cst, mk(QC_C_DECIMATE, C, T.filt185D3, T.n_filt185D3, 1, 3)); rate /= 6; break;
    case 740:
        ADD(cst, mk(QC_C_DECIM2_HB45, C, nullptr, 0, 1, 1));
        ADD(cst, mk(QC_C_DECIM2_HB45, C, nullptr, 0, 1, 1));
        ADD(cst, mk(QC_C_DECIMATE, C, T.filt185D3, T.n_filt185D3, 1, 3)); rate /= 12; break;
    case 1333:
        for (int i = 0; i < 3; i++) ADD(cst, mk(QC_C_DECIM2_HB45, C, nullptr, 0, 1, 1));
        ADD(cst, mk(QC_C_DECIMATE, C, T.filt167D3, T.n_filt167D3, 1, 3)); rate /= 24; break;
    default: {
        int d2, d3, d5;
        plan_decimation(sample_rate, &d2, &d3, &d5);
        int i2 = d2, nhb = 0;
        while (i2 > 1 && nhb < 5) { ADD(cst, mk(QC_C_DECIM2_HB45, C, nullptr, 0, 1, 1)); rate /= 2; i2--; nhb++; }
        for (int i = 0; i < d3; i++) { ADD(cst, mk(QC_C_DECIMATE, C, T.filt144D3, T.n_filt144D3, 1, 3)); rate /= 3; }
        for (int i = 0; i < d5; i++) { ADD(cst, mk(QC_C_DECIMATE, C, T.filt240D5Sharp, T.n_filt240D5Sharp, 1, 5)); rate /= 5; }
        if (i2 > 0) { ADD(cst, mk(QC_C_DECIMATE, C, T.filt48dec24, T.n_filt48dec24, 1, 2)); rate /= 2; }
        if (rate >= 50000) {                    // quisk.c:1834-1838
            rate = rate * 24 / 25;
            ADD(cst, mk(QC_C_INTERPDECIM, C, T.filt300D5, T.n_filt300D5, 6, 5));
            ADD(cst, mk(QC_C_INTERPDECIM, C, T.filt240D5Sharp, T.n_filt240D5Sharp, 4, 5));
        }
    } }
    decim_srate = rate;
    n_decim_stages = (int)cst.size();
    // ---- quisk_process_demodulate ------------------------------------------------
    if (cfg.n_filt <= 0 || !cfg.filt_i) { set_error("rx_create: receive filter taps missing"); return QC_EINVAL; }
    std::vector<double> iq((size_t)2 * cfg.n_filt);
    for (int i = 0; i < cfg.n_filt; i++) { iq[i] = cfg.filt_i[i]; iq[cfg.n_filt + i] = cfg.filt_q ? cfg.filt_q[i] : cfg.filt_i[i]; }
    switch (mode) {
    case QC_MODE_CWL: case QC_MODE_CWU:         // quisk.c:1909-1955
        filter_srate = decim_srate / 8;
        ADD(cst, mk(QC_C_DECIM2_HB45, C, nullptr, 0, 1, 1));
        ADD(cst, mk(QC_C_DECIM2_HB45, C, nullptr, 0, 1, 1));
        ADD(cst, mk(QC_C_DECIMATE, C, T.filt48dec24, T.n_filt48dec24, 1, 2));
        rxf = mk(QC_C_RXFILTER, C, iq.data(), cfg.n_filt, 1, 1);
        ADD(rst, mk(QC_D_INTERPOLATE, C, T.audio24p4, T.n_audio24p4, 2, 1));
        ADD(rst, mk(QC_D_INTERP2_HB45, C, nullptr, 0, 1, 1));
        ADD(rst, mk(QC_D_INTERP2_HB45, C, nullptr, 0, 1, 1));
        break;
    case QC_MODE_LSB: case QC_MODE_USB:         // quisk.c:1956-2001
        filter_srate = decim_srate / 4;
        ADD(cst, mk(QC_C_DECIM2_HB45, C, nullptr, 0, 1, 1));
        ADD(cst, mk(QC_C_DECIMATE, C, T.filt48dec24, T.n_filt48dec24, 1, 2));
        rxf = mk(QC_C_RXFILTER, C, iq.data(), cfg.n_filt, 1, 1);
        ADD(rst, mk(QC_D_INTERPOLATE, C, T.audio24p4, T.n_audio24p4, 2, 1));
        ADD(rst, mk(QC_D_INTERP2_HB45, C, nullptr, 0, 1, 1));
        break;
    case QC_MODE_AM:                            // quisk.c:2002-2026
        filter_srate = decim_srate / 2;
        ADD(cst, mk(QC_C_DECIMATE, C, T.filt48dec24, T.n_filt48dec24, 1, 2));
        rxf = mk(QC_D_RXFILTER, C, iq.data(), cfg.n_filt, 1, 1);
        ADD(rst, mk(QC_D_DECIMATE, C, T.audio24p6, T.n_audio24p6, 1, 1));
        ADD(rst, mk(QC_D_INTERP2_HB45, C, nullptr, 0, 1, 1));
        QC_CUDA(cudaMalloc((void **)&d_dc, (size_t)C * sizeof(double)));
        QC_CUDA(cudaMemset(d_dc, 0, (size_t)C * sizeof(double)));
        break;
    case QC_MODE_FM: case QC_MODE_DGT_FM: {     // quisk.c:2026-2075 (one case label for both: DGT-FM differs only in where quisk_process_samples sends the audio, quisk.c:2633)
        filter_srate = decim_srate;
        rxf = mk(QC_D_RXFILTER, C, iq.data(), cfg.n_filt, 1, 1);
        ADD(rst, mk(QC_D_DECIMATE, C, T.lpFilt48, T.n_lpFilt48, 1, 4));
        ADD(rst, mk(QC_D_DECIMATE, C, T.audioFmHp, T.n_audioFmHp, 1, 1));
        ADD(rst, mk(QC_D_INTERP2_HB45, C, nullptr, 0, 1, 1));
        ADD(rst, mk(QC_D_INTERP2_HB45, C, nullptr, 0, 1, 1));
        const double www = tan(M_PI * 300.0 / 48000);       // FM_FILTER_DEMPH, quisk.c:46,1895-1899
        const double nnn = 1.0 / (1.0 + www);
        fm_a0 = www * nnn; fm_a1 = fm_a0; fm_b1 = nnn * (www - 1.0);
        QC_CUDA(cudaMalloc((void **)&d_fm, (size_t)C * 4 * sizeof(double)));
        { int rcf = reset_fm(); if (rcf != QC_OK) return rcf; }
        break; }
    case QC_MODE_DGT_U: case QC_MODE_DGT_L: case QC_MODE_FDV_U: case QC_MODE_FDV_L:        // quisk.c:2087-2140
        if (cfg.filter_bandwidth < 3000) {      // DGT_NARROW_FREQ, quisk.c:52: filter at 6 kS/s like CW
            filter_srate = decim_srate / 8;
            ADD(cst, mk(QC_C_DECIM2_HB45, C, nullptr, 0, 1, 1));
            ADD(cst, mk(QC_C_DECIM2_HB45, C, nullptr, 0, 1, 1));
            ADD(cst, mk(QC_C_DECIMATE, C, T.filt48dec24, T.n_filt48dec24, 1, 2));
            rxf = mk(QC_C_RXFILTER, C, iq.data(), cfg.n_filt, 1, 1);
            ADD(rst, mk(QC_D_INTERPOLATE, C, T.audio24p4, T.n_audio24p4, 2, 1));
            ADD(rst, mk(QC_D_INTERP2_HB45, C, nullptr, 0, 1, 1));
            ADD(rst, mk(QC_D_INTERP2_HB45, C, nullptr, 0, 1, 1));
        } else {                                // filter at 48 kS/s, no audio resampling
            filter_srate = decim_srate;
            rxf = mk(QC_C_RXFILTER, C, iq.data(), cfg.n_filt, 1, 1);
        }
        break;
    case QC_MODE_DGT_IQ:                        // quisk.c:2141-2153: complex out, real-tap filter unless very wide
        filter_srate = decim_srate;
        iq_out = true;
        if (cfg.filter_bandwidth < 19000) rxf = mk(QC_D_RXFILTER, C, iq.data(), cfg.n_filt, 1, 1);
        break;
    default:
        // EXT (6) leaves quisk_process_samples for quisk_extern_demod before the demodulator (quisk.c:2490-2494); IMD (10) has no
        // case in quisk_process_demodulate's switch at all (the reference returns the buffer untouched)
        set_error("rx_create: mode %d is not on the accelerated path", mode); return QC_EINVAL;
    }
    if (!rxf && !iq_out) return QC_EINVAL;
    // ---- tuning NCO ----------------------------------------------------------------
    if (cfg.tune_hz) {
        tune_hz.assign(cfg.tune_hz, cfg.tune_hz + C);
        bool any = false;
        for (double t : tune_hz) any = any || t != 0.0;
        tune = any;
        if (tune) {
            QC_CUDA(cudaMalloc((void **)&d_nco, (size_t)C * 8 * sizeof(double)));
            int rc = upload_nco(); if (rc != QC_OK) return rc;
        }
    }
    return QC_OK;
}

int RxChain::reset_fm()
{
    std::vector<double> st((size_t)C * 4, 0.0);
    for (int c = 0; c < C; c++) st[(size_t)c * 4] = 10.0;         // fm_1 = 10, quisk.c:1893
    QC_CUDA(cudaMemcpy(d_fm, st.data(), st.size() * sizeof(double), cudaMemcpyHostToDevice));
    return QC_OK;
}

// The tuning phasor at the first sample of each block comes from the reference's own recurrence
// (nco_advance_kernel).  Block b reads d_v[vcur]; while its kernels run, the side stream advances that value by
// `count` steps into the next ring slot for block b + 1.  Events keep the two streams honest: a consumer waits for
// the recurrence that fills its slot, and the recurrence that recycles a slot waits for the consumer that read it.
int RxChain::nco_before(int count, cudaStream_t s)
{
    // the recurrence for the NEXT block goes first: it only needs this block's start value and length
    const int b = vcur, nb = (vcur + 1) % NV;
    if (f_set[nb]) QC_CUDA(cudaStreamWaitEvent(s_nco, ev_f[nb], 0));
    // exact_nco = 0: the next block start comes from the closed form too (one step instead of `count`); the phasor
    // then drifts from the reference's by ~1e-19 per sample (1e-12 after ~8 s at 1.536 MS/s)
    int rc = exact_nco ? launch_nco_advance(d_v[b], d_v[nb], d_nco, count, C, d_sched, ++nco_epoch, s_nco)
                       : launch_nco_jump(d_v[b], d_v[nb], d_nco, count, C, s_nco);
    if (rc != QC_OK) return rc;
    QC_CUDA(cudaEventRecord(ev_r[nb], s_nco)); r_set[nb] = true;
    if (r_set[b]) QC_CUDA(cudaStreamWaitEvent(s, ev_r[b], 0));
    return QC_OK;
}

int RxChain::nco_after(int count, cudaStream_t s)
{
    (void)count;
    QC_CUDA(cudaEventRecord(ev_f[vcur], s)); f_set[vcur] = true;
    vcur = (vcur + 1) % NV;
    return QC_OK;
}

int RxChain::upload_nco()
{
    std::vector<double> h((size_t)C * 8);
    for (int c = 0; c < C; c++) nco_make(tune_hz[c], sample_rate, 1.0, 0.0, &h[(size_t)c * 8]);
    QC_CUDA(cudaMemcpy(d_nco, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    n_base = 0;
    // rxTuneVector starts at 1 (quisk.c:2308); (re)start the ring of block-start phasors
    if (!s_nco) {
        // highest priority: the few tiny CTAs of the recurrence must be dispatched at once, not queued behind the
        // thousands of CTAs of the block kernels (measured: without it the two serialise, -25 % throughput)
        int pr_lo = 0, pr_hi = 0;
        QC_CUDA(cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi));
        QC_CUDA(cudaStreamCreateWithPriority(&s_nco, cudaStreamNonBlocking, pr_hi));
        QC_CUDA(cudaMalloc((void **)&d_sched, 1025 * sizeof(unsigned)));
        QC_CUDA(cudaMemset(d_sched, 0, 1025 * sizeof(unsigned)));
        nco_epoch = 0;
        for (int i = 0; i < NV; i++) {
            QC_CUDA(cudaMalloc((void **)&d_v[i], (size_t)C * sizeof(cd)));
            QC_CUDA(cudaEventCreateWithFlags(&ev_r[i], cudaEventDisableTiming));
            QC_CUDA(cudaEventCreateWithFlags(&ev_f[i], cudaEventDisableTiming));
        }
    }
    QC_CUDA(cudaDeviceSynchronize());
    std::vector<cd> one((size_t)C, make_double2(1.0, 0.0));
    QC_CUDA(cudaMemcpy(d_v[0], one.data(), one.size() * sizeof(cd), cudaMemcpyHostToDevice));
    vcur = 0;
    for (int i = 0; i < NV; i++) { r_set[i] = false; f_set[i] = false; }
    return QC_OK;
}

int RxChain::n_host_chunks() const
{
    int k = host_chunks > 0 ? host_chunks : (C >= 1024 ? 8 : 1);
    if (k > C) k = C;
    return k < 1 ? 1 : k;
}

void RxChain::release_sub_chains()
{
    for (RxChain *r : sub) { r->release(); delete r; }
    sub.clear(); sub_c0.clear();
}

int RxChain::build_sub_chains(int k)
{
    release_sub_chains();
    qcRxConfig cfg = saved.cfg;
    cfg.filt_i = saved.fi.empty() ? nullptr : saved.fi.data();
    cfg.filt_q = saved.fq.empty() ? nullptr : saved.fq.data();
    const double **ptrs[13] = {&cfg.tables.filt144D3, &cfg.tables.filt240D5Sharp, &cfg.tables.filt48dec24, &cfg.tables.filt300D5, &cfg.tables.audio24p4,
                               &cfg.tables.audio24p6, &cfg.tables.lpFilt48, &cfg.tables.audioFmHp, &cfg.tables.filt53D1, &cfg.tables.filt111D2,
                               &cfg.tables.filt133D2, &cfg.tables.filt167D3, &cfg.tables.filt185D3};
    for (int i = 0; i < 13; i++) *ptrs[i] = saved.tab[i].empty() ? nullptr : saved.tab[i].data();
    for (int j = 0; j < k; j++) {
        const int c0 = (int)((long)C * j / k), c1 = (int)((long)C * (j + 1) / k);
        cfg.n_channels = c1 - c0;
        cfg.tune_hz = saved.tune.empty() ? nullptr : saved.tune.data() + c0;
        RxChain *r = new RxChain();
        if (r->init(cfg) != QC_OK) { r->release(); delete r; release_sub_chains(); return QC_EINVAL; }
        r->host_chunks = 1;
        r->exact_nco = exact_nco; r->fused_tail = fused_tail; r->nb_level = nb_level; r->auto_notch = auto_notch; r->notch_sidetone = notch_sidetone; r->squelch_level = squelch_level; r->fused_chunk = fused_chunk; r->fused_threads = fused_threads;
        r->fused_plans = fused_plans; r->fused_tailwarp = fused_tailwarp; r->fused_dense = fused_dense; r->fused_split = fused_split; r->fused_async = fused_async; r->fused_p3 = fused_p3;
        r->fused_min_r = fused_min_r; r->fused_deepk = fused_deepk;
        sub.push_back(r); sub_c0.push_back(c0);
    }
    return QC_OK;
}

void RxChain::release()
{
    release_sub_chains();
    for (auto *f : cst) { f->release(); delete f; }
    for (auto *f : rst) { f->release(); delete f; }
    if (rxf) { rxf->release(); delete rxf; }
    cst.clear(); rst.clear(); rxf = nullptr;
    for (int i = 0; i < 2; i++) { if (bufc[i]) cudaFree(bufc[i]); if (bufr[i]) cudaFree(bufr[i]); bufc[i] = nullptr; bufr[i] = nullptr; }
    if (s_nco) { cudaStreamSynchronize(s_nco); cudaStreamDestroy(s_nco); s_nco = nullptr; }
    if (d_sched) { cudaFree(d_sched); d_sched = nullptr; }
    if (nb) { quisk_cuda_nb_destroy(nb); nb = nullptr; }
    if (anotch) { quisk_cuda_autonotch_destroy(anotch); anotch = nullptr; }
    if (squelch) { quisk_cuda_ssb_squelch_destroy(squelch); squelch = nullptr; }
    for (int i = 0; i < NV; i++) {
        if (d_v[i]) cudaFree(d_v[i]); d_v[i] = nullptr;
        if (ev_r[i]) cudaEventDestroy(ev_r[i]); if (ev_f[i]) cudaEventDestroy(ev_f[i]);
        ev_r[i] = ev_f[i] = nullptr;
    }
    if (d_nco) cudaFree(d_nco); if (d_dc) cudaFree(d_dc); if (d_fm) cudaFree(d_fm);
    if (h_pin) cudaFreeHost(h_pin); if (d_host_in) cudaFree(d_host_in); if (d_host_out) cudaFree(d_host_out);
    d_nco = d_dc = d_fm = nullptr; h_pin = nullptr; d_host_in = nullptr; d_host_out = nullptr;
    release_fused();
}

int RxChain::reserve(int count)
{
    // The scratch rows hold every intermediate of the stage graph, and a stage may emit MORE samples than it
    // takes: below 96 kS/s PlanDecimation finds no integer decimation and the list starts with the 6/5 converter
    // (quisk.c:1834-1838), so walk the complex stages and size for the largest intermediate.
    double big = count, cur = count;
    for (auto *f : cst) {
        switch (f->kind) {
        case QC_C_DECIM2_HB45: cur = cur / 2 + 1; break;
        case QC_C_INTERPDECIM: cur = cur * f->interp / f->decim + 2; break;
        default: cur = cur / f->decim + 1; break;
        }
        if (cur > big) big = cur;
    }
    const long need = (long)big + 64;
    if (need <= cap) return QC_OK;
    for (int i = 0; i < 2; i++) {
        if (bufc[i]) cudaFree(bufc[i]); if (bufr[i]) cudaFree(bufr[i]);
        bufc[i] = nullptr; bufr[i] = nullptr;
    }
    cap = need;
    // complex scratch only has to hold what follows the first decimating stage when the
    // input itself is never copied (no tuning); keep it simple and size for the block.
    for (int i = 0; i < 2; i++) {
        QC_CUDA(cudaMalloc((void **)&bufc[i], (size_t)C * cap * sizeof(cd)));
        QC_CUDA(cudaMalloc((void **)&bufr[i], (size_t)C * cap * sizeof(double)));
    }
    return QC_OK;
}

int RxChain::max_out(int count) const
{
    // every stage at worst keeps ceil() of its ratio; the audio side multiplies by <= 8
    double n = count;
    for (auto *f : cst) {
        switch (f->kind) {
        case QC_C_DECIM2_HB45: n = n / 2 + 1; break;
        case QC_C_INTERPDECIM: n = n * f->interp / f->decim + 2; break;
        default: n = n / f->decim + 1; break;
        }
    }
    for (auto *f : rst) {
        switch (f->kind) {
        case QC_D_INTERPOLATE: n = n * f->interp; break;
        case QC_D_INTERP2_HB45: n = n * 2; break;
        default: n = n / f->decim + 1; break;
        }
    }
    if (iq_out) n = 2 * n;                      // (re, im) pairs, counted in doubles
    int r = (int)n + 8;
    if (iq_out) r += r & 1;
    return r;
}

int RxChain::process(const void *d_iq, long iq_stride, int count, double *d_audio, long audio_stride, int *n_audio,
                     void *d_decim, long decim_stride, int *n_decim, cudaStream_t s)
{
    if (count < 0) { set_error("rx_process: negative count"); return QC_EINVAL; }
    if (n_audio) *n_audio = 0;
    if (n_decim) *n_decim = 0;
    if (count == 0) return QC_OK;
    if (poisoned) { set_error("rx_process: an earlier call failed after stage state had advanced; call quisk_cuda_rx_reset() first"); return QC_EINVAL; }
    int rc = reserve(count); if (rc != QC_OK) return rc;
    // Validate the caller's buffers against the output counts this call WILL produce before any stage state (NCO ring,
    // decimation phases, history ping-pong) advances: the phases are host integers, so the counts are known up front.
    {
        int np_ = count;
        for (auto *f : cst) np_ = f->count_out(np_, 0);
        if (iq_out) {
            if (audio_stride < 2L * np_ || (audio_stride & 1)) { set_error("rx_process: DGT-IQ needs an even audio_stride >= 2 * samples"); return QC_EINVAL; }
        } else {
            for (auto *f : rst) np_ = f->count_out(np_, 0);
            if (audio_stride < np_) { set_error("rx_process: audio_stride %ld < %d samples this call produces", audio_stride, np_); return QC_EINVAL; }
        }
    }
    struct Latch { bool &p; bool ok = false; ~Latch() { if (!ok) p = true; } } latch{poisoned};     // any error return below leaves the streams out of step

    const cd *cur = (const cd *)d_iq; long stride = iq_stride; int n = count; int pp = 0;
    size_t first_stage = 0;
    // fuse through the demodulator's own pre-decimation unless the caller wants the 48 kS/s tap
    const size_t nf = fused ? fusable_prefix(d_decim ? (size_t)n_decim_stages : cst.size()) : 0;
    if (tune) { rc = nco_before(count, s); if (rc != QC_OK) return rc; }
    if (nf > 0) {
        rc = run_fused_decimator(nf, cur, stride, count, bufc[0], cap, &n, s); if (rc != QC_OK) return rc;
        cur = bufc[0]; stride = cap; pp = 1; first_stage = nf;
    } else if (tune) {
        rc = launch_tune(cur, stride, bufc[0], cap, n, C, d_nco, d_v[vcur], 0, s); if (rc != QC_OK) return rc;
        cur = bufc[0]; stride = cap; pp = 1;
    }
    if (tune) { rc = nco_after(count, s); if (rc != QC_OK) return rc; n_base += (unsigned long long)count; }
    for (size_t i = first_stage; i < cst.size(); i++) {
        if ((int)i == n_decim_stages && d_decim) {
            QC_CUDA(cudaMemcpy2DAsync(d_decim, (size_t)decim_stride * sizeof(cd), cur, (size_t)stride * sizeof(cd),
                                      (size_t)n * sizeof(cd), C, cudaMemcpyDeviceToDevice, s));
            if (n_decim) *n_decim = n;
        }
        int no = 0;
        rc = cst[i]->run(cur, stride, n, bufc[pp], cap, &no, 0, s); if (rc != QC_OK) return rc;
        cur = bufc[pp]; stride = cap; pp ^= 1; n = no;
    }
    if ((int)cst.size() == n_decim_stages && d_decim) {
        QC_CUDA(cudaMemcpy2DAsync(d_decim, (size_t)decim_stride * sizeof(cd), cur, (size_t)stride * sizeof(cd),
                                  (size_t)n * sizeof(cd), C, cudaMemcpyDeviceToDevice, s));
        if (n_decim) *n_decim = n;
    }
    int no = 0;
    if (iq_out) {                               // DGT-IQ: (re, im) pairs straight into the caller's buffer
        if (audio_stride < 2L * n || (audio_stride & 1)) { set_error("rx_process: DGT-IQ needs an even audio_stride >= 2 * samples"); return QC_EINVAL; }
        if (rxf) { rc = rxf->run(cur, stride, n, d_audio, audio_stride / 2, &no, 0, s); if (rc != QC_OK) return rc; }
        else {
            QC_CUDA(cudaMemcpy2DAsync(d_audio, (size_t)audio_stride * sizeof(double), cur, (size_t)stride * sizeof(cd),
                                      (size_t)n * sizeof(cd), C, cudaMemcpyDeviceToDevice, s));
            no = n;
        }
        if (n_audio) *n_audio = no;
        latch.ok = true;
        return QC_OK;
    }
    if (fused && fused_tail && tail_fusable() && !audio_options()) {
        rc = run_tail(cur, stride, n, d_audio, audio_stride, &no, s);
        if (rc == QC_OK) { if (n_audio) *n_audio = no; latch.ok = true; return QC_OK; }
        if (rc != QC_ENOMEM) return rc;          // too long for shared memory: per-stage kernels below
    }
    // main receive filter
    rc = rxf->run(cur, stride, n, bufc[pp], cap, &no, 0, s); if (rc != QC_OK) return rc;
    cur = bufc[pp]; n = no;
    // detector -> real audio at the filter rate
    double *rcur = bufr[0]; int rp = 1; long rstride = cap;
    if (rst.empty()) {                          // wide DGT: the detector output is the audio
        if (audio_stride < n) { set_error("rx_process: audio_stride too small"); return QC_EINVAL; }
        rcur = d_audio; rstride = audio_stride;
    }
    switch (mode) {
    case QC_MODE_CWL: case QC_MODE_LSB: case QC_MODE_DGT_L: case QC_MODE_FDV_L: rc = launch_demod_ssb(cur, cap, rcur, rstride, n, C, 1, s); break;
    case QC_MODE_CWU: case QC_MODE_USB: case QC_MODE_DGT_U: case QC_MODE_FDV_U: rc = launch_demod_ssb(cur, cap, rcur, rstride, n, C, 0, s); break;
    case QC_MODE_AM: rc = launch_am_detect(cur, cap, rcur, cap, n, C, d_dc, s); break;
    case QC_MODE_FM: case QC_MODE_DGT_FM: rc = launch_fm_detect(cur, cap, rcur, cap, n, C, d_fm, fm_a0, fm_a1, fm_b1, s); break;
    }
    if (rc != QC_OK) return rc;
    if (audio_options()) {
        // quisk.c:1923-1928 (and the same lines of the other side-band and AM branches): notch, then squelch + its delay line
        if (auto_notch) {
            if (!anotch) anotch = quisk_cuda_autonotch_create(C, filter_srate);
            if (!anotch) return QC_EINVAL;
            const bool cw = mode == QC_MODE_CWL || mode == QC_MODE_CWU;
            rc = quisk_cuda_autonotch_run(anotch, rcur, rstride, n, cw ? notch_sidetone : 0, s); if (rc != QC_OK) return rc;
        }
        if (squelch_level > 0) {
            if (!squelch) squelch = quisk_cuda_ssb_squelch_create(C, filter_srate, filter_bandwidth);
            if (!squelch) return QC_EINVAL;
            rc = quisk_cuda_ssb_squelch_run(squelch, rcur, rstride, n, squelch_level, s); if (rc != QC_OK) return rc;
        }
    }
    // audio stages; the last one writes straight into the caller's buffer
    for (size_t i = 0; i < rst.size(); i++) {
        const bool last = i + 1 == rst.size();
        double *dst = last ? d_audio : bufr[rp];
        const long dstride = last ? audio_stride : cap;
        if (last && audio_stride < rst[i]->count_out(n, 0)) { set_error("rx_process: audio_stride too small"); return QC_EINVAL; }
        rc = rst[i]->run(rcur, (rcur == d_audio) ? audio_stride : cap, n, dst, dstride, &no, 0, s); if (rc != QC_OK) return rc;
        rcur = dst; rp ^= 1; n = no;
    }
    if (squelch_level > 0 && squelch && audio_options() && n > 0) {
        // quisk_process_samples mutes the block of a receiver whose squelch is closed (quisk.c:2552-2623, 2716-2719)
        rc = launch_mute_rows(d_audio, audio_stride, n, C, quisk_cuda_ssb_squelch_state_ptr(squelch), s); if (rc != QC_OK) return rc;
    }
    if (n_audio) *n_audio = n;
    latch.ok = true;
    return QC_OK;
}

bool RxChain::audio_options() const
{   // the modes whose branch of quisk_process_demodulate has both stages at the filter rate: CW, SSB, AM
    if (!auto_notch && squelch_level <= 0) return false;
    return mode == QC_MODE_CWL || mode == QC_MODE_CWU || mode == QC_MODE_LSB || mode == QC_MODE_USB || mode == QC_MODE_AM;
}

// quisk_process_samples runs NoiseBlanker on the raw block in front of the tuning stage (quisk.c:2448-2449); the host
// entries do the same on their staged copy of the block when QC_RX_OPT_NOISE_BLANKER is set.
int RxChain::host_noise_blanker(cudaStream_t s, int count)
{
    if (nb_level <= 0) return QC_OK;
    if (!nb) nb = quisk_cuda_nb_create(C, sample_rate);
    if (!nb) return QC_EINVAL;
    return quisk_cuda_nb_run(nb, d_host_in, host_cap, count, nb_level, s);
}

// One chain's share of a host call, enqueued on its own stream `hs` and NOT waited for: H2D of the block (complex double,
// or wire bytes + the widening kernel), the optional noise blanker, the chain, D2H of the audio.
int RxChain::host_enqueue(const quisk_cd *h_iq, long iq_stride, const void *h_bytes, long byte_stride, int nb_, int big, int count,
                          double *h_audio, long audio_stride, int *n_audio)
{
    if (!hs) QC_CUDA(cudaStreamCreateWithFlags(&hs, cudaStreamNonBlocking));
    const int mo = max_out(count);
    if (count > host_cap) {
        if (d_host_in) cudaFree(d_host_in); if (d_host_out) cudaFree(d_host_out); if (h_pin) cudaFreeHost(h_pin);
        d_host_in = nullptr; d_host_out = nullptr; h_pin = nullptr;
        host_cap = count; host_out_cap = mo;
        QC_CUDA(cudaMalloc((void **)&d_host_in, (size_t)C * host_cap * sizeof(cd)));
        QC_CUDA(cudaMalloc((void **)&d_host_out, (size_t)C * host_out_cap * sizeof(double)));
    }
    if (h_bytes) {
        const size_t row = (size_t)count * 2 * nb_;
        if ((size_t)C * row > packed_cap) {
            if (d_packed) cudaFree(d_packed);
            d_packed = nullptr; packed_cap = 0;
            QC_CUDA(cudaMalloc((void **)&d_packed, (size_t)C * row));
            packed_cap = (size_t)C * row;
        }
        QC_CUDA(cudaMemcpy2DAsync(d_packed, row, h_bytes, (size_t)byte_stride, row, C, cudaMemcpyHostToDevice, hs));
        int rcu = launch_unpack_iq(d_packed, (long)row, C, count, nb_, big, d_host_in, host_cap, hs);
        if (rcu != QC_OK) return rcu;
    } else {
        // H2D straight from the caller's memory (pinned by the caller or pageable), row by row layout kept
        QC_CUDA(cudaMemcpy2DAsync(d_host_in, (size_t)host_cap * sizeof(cd), h_iq, (size_t)iq_stride * sizeof(cd),
                                  (size_t)count * sizeof(cd), C, cudaMemcpyHostToDevice, hs));
    }
    int na = 0;
    int rc = host_noise_blanker(hs, count);
    if (rc != QC_OK) return rc;
    rc = process(d_host_in, host_cap, count, d_host_out, host_out_cap, &na, nullptr, 0, nullptr, hs);
    if (rc != QC_OK) return rc;
    const int nd = iq_out ? 2 * na : na;        // doubles per channel
    if (nd > audio_stride) { set_error("rx_process_host: audio_stride %ld < %d", audio_stride, nd); return QC_EINVAL; }
    if (nd > 0)
        QC_CUDA(cudaMemcpy2DAsync(h_audio, (size_t)audio_stride * sizeof(double), d_host_out, (size_t)host_out_cap * sizeof(double),
                                  (size_t)nd * sizeof(double), C, cudaMemcpyDeviceToHost, hs));
    if (n_audio) *n_audio = na;
    return QC_OK;
}

static int host_call(RxChain &rx, const quisk_cd *h_iq, long iq_stride, const void *h_bytes, long byte_stride, int nb, int big, int count,
                     double *h_audio, long audio_stride, int *n_audio)
{
    if (count <= 0) { if (n_audio) *n_audio = 0; return QC_OK; }
    const int k = rx.n_host_chunks();
    if (k <= 1) {
        int rc = rx.host_enqueue(h_iq, iq_stride, h_bytes, byte_stride, nb, big, count, h_audio, audio_stride, n_audio);
        if (rc != QC_OK) return rc;
        QC_CUDA(cudaStreamSynchronize(rx.hs));
        return QC_OK;
    }
    if ((int)rx.sub.size() != k) { int rc = rx.build_sub_chains(k); if (rc != QC_OK) return rc; }
    int na = 0;
    for (int j = 0; j < k; j++) {
        const int c0 = rx.sub_c0[j];
        int naj = 0;
        int rc = rx.sub[j]->host_enqueue(h_iq ? h_iq + (size_t)c0 * iq_stride : nullptr, iq_stride,
                                         h_bytes ? (const unsigned char *)h_bytes + (size_t)c0 * byte_stride : nullptr, byte_stride, nb, big, count,
                                         h_audio + (size_t)c0 * audio_stride, audio_stride, &naj);
        if (rc != QC_OK) return rc;
        na = naj;
    }
    for (int j = 0; j < k; j++) QC_CUDA(cudaStreamSynchronize(rx.sub[j]->hs));
    if (n_audio) *n_audio = na;
    return QC_OK;
}

int RxChain::process_host(const quisk_cd *h_iq, long iq_stride, int count, double *h_audio, long audio_stride, int *n_audio)
{ return host_call(*this, h_iq, iq_stride, nullptr, 0, 0, 0, count, h_audio, audio_stride, n_audio); }

int RxChain::process_host_packed(const void *h_bytes, long byte_stride, int count, int nb_, int big, double *h_audio, long audio_stride, int *n_audio)
{
    if (count > 0 && (nb_ < 1 || nb_ > 4 || byte_stride < (long)count * 2 * nb_)) { set_error("rx_process_host_packed: bad sizes"); return QC_EINVAL; }
    return host_call(*this, nullptr, 0, h_bytes, byte_stride, nb_, big, count, h_audio, audio_stride, n_audio);
}

int RxChain::reset()
{
    for (auto *f : cst) { int rc = f->reset(nullptr); if (rc != QC_OK) return rc; }
    for (auto *f : rst) { int rc = f->reset(nullptr); if (rc != QC_OK) return rc; }
    int rc = rxf ? rxf->reset(nullptr) : QC_OK; if (rc != QC_OK) return rc;
    if (d_dc) QC_CUDA(cudaMemset(d_dc, 0, (size_t)C * sizeof(double)));
    if (d_fm) { rc = reset_fm(); if (rc != QC_OK) return rc; }
    if (tune) { rc = upload_nco(); if (rc != QC_OK) return rc; }
    rc = reset_fused(); if (rc != QC_OK) return rc;
    if (anotch) { quisk_cuda_autonotch_destroy(anotch); anotch = nullptr; }          // created again, in its start state, by the next block
    if (squelch) { quisk_cuda_ssb_squelch_destroy(squelch); squelch = nullptr; }
    for (RxChain *r : sub) { rc = r->reset(); if (rc != QC_OK) return rc; }
    QC_CUDA(cudaDeviceSynchronize());
    poisoned = false;
    return QC_OK;
}

}  // namespace qc

struct qcRxChain { qc::RxChain rx; };

extern "C" {

int quisk_cuda_plan_decimation(int sample_rate, int *d2, int *d3, int *d5) { return qc::plan_decimation(sample_rate, d2, d3, d5); }

qcRxChain *quisk_cuda_rx_create(const struct qcRxConfig *cfg)
{
    if (!cfg) { qc::set_error("rx_create: null config"); return nullptr; }
    if (qc::ensure_device() != QC_OK) return nullptr;
    qcRxChain *r = new qcRxChain();
    if (r->rx.init(*cfg) != QC_OK) { r->rx.release(); delete r; return nullptr; }
    return r;
}

void quisk_cuda_rx_destroy(qcRxChain *rx) { if (rx) { rx->rx.release(); delete rx; } }
int quisk_cuda_rx_decim_srate(const qcRxChain *rx) { return rx ? rx->rx.decim_srate : QC_EINVAL; }
int quisk_cuda_rx_filter_srate(const qcRxChain *rx) { return rx ? rx->rx.filter_srate : QC_EINVAL; }

int quisk_cuda_rx_squelch_active(qcRxChain *rx, int *h_active)
{   // MeasureSquelch[].squelch_active of every receiver after the last block (QC_RX_OPT_SSB_SQUELCH); all zero while the option is off
    if (!rx || !h_active) return QC_EINVAL;
    for (int c = 0; c < rx->rx.C; c++) h_active[c] = 0;
    if (!rx->rx.squelch || rx->rx.squelch_level <= 0) return QC_OK;
    QC_CUDA(cudaDeviceSynchronize());
    return quisk_cuda_ssb_squelch_state(rx->rx.squelch, nullptr, h_active, nullptr);
}
int quisk_cuda_rx_max_out(const qcRxChain *rx, int count) { return rx ? rx->rx.max_out(count) : QC_EINVAL; }

int quisk_cuda_rx_process(qcRxChain *rx, const void *d_iq, long iq_stride, int count, double *d_audio, long audio_stride,
                          int *n_audio, void *d_decim, long decim_stride, int *n_decim, void *stream)
{
    if (!rx) { qc::set_error("rx_process: null chain"); return QC_EINVAL; }
    return rx->rx.process(d_iq, iq_stride, count, d_audio, audio_stride, n_audio, d_decim, decim_stride, n_decim, (cudaStream_t)stream);
}

int quisk_cuda_rx_process_host_packed(qcRxChain *rx, const void *h_bytes, long byte_stride, int count, int bytes, int big_endian,
                                      double *h_audio, long audio_stride, int *n_audio)
{
    if (!rx || !h_bytes || !h_audio) { qc::set_error("rx_process_host_packed: null pointer"); return QC_EINVAL; }
    return rx->rx.process_host_packed(h_bytes, byte_stride, count, bytes, big_endian, h_audio, audio_stride, n_audio);
}

int quisk_cuda_rx_process_host(qcRxChain *rx, const quisk_cd *h_iq, long iq_stride, int count, double *h_audio,
                               long audio_stride, int *n_audio)
{
    if (!rx) { qc::set_error("rx_process_host: null chain"); return QC_EINVAL; }
    return rx->rx.process_host(h_iq, iq_stride, count, h_audio, audio_stride, n_audio);
}

int quisk_cuda_rx_reset(qcRxChain *rx) { return rx ? rx->rx.reset() : QC_EINVAL; }

int quisk_cuda_rx_set_option(qcRxChain *rx, int option, int value)
{
    if (!rx) return QC_EINVAL;
    switch (option) {
    case QC_RX_OPT_TIMING: rx->rx.timing = value != 0; return QC_OK;
    case QC_RX_OPT_NOISE_BLANKER:
        if (value < 0 || value > 3) { qc::set_error("rx_set_option: noise blanker level must be 0 (off) .. 3"); return QC_EINVAL; }
        rx->rx.nb_level = value;
        for (qc::RxChain *r : rx->rx.sub) r->nb_level = value;
        return QC_OK;
    case QC_RX_OPT_AUTO_NOTCH: rx->rx.auto_notch = value ? 1 : 0; for (qc::RxChain *r : rx->rx.sub) r->auto_notch = rx->rx.auto_notch; return QC_OK;
    case QC_RX_OPT_NOTCH_SIDETONE: rx->rx.notch_sidetone = value; for (qc::RxChain *r : rx->rx.sub) r->notch_sidetone = value; return QC_OK;
    case QC_RX_OPT_SSB_SQUELCH:
        if (value < 0) { qc::set_error("rx_set_option: squelch level must be 0 (off) or ssb_squelch_level > 0"); return QC_EINVAL; }
        rx->rx.squelch_level = value; for (qc::RxChain *r : rx->rx.sub) r->squelch_level = value; return QC_OK;
    case QC_RX_OPT_HOST_CHUNKS:
        if (value < 0 || value > 64) { qc::set_error("rx_set_option: host chunks must be 0 (auto) .. 64"); return QC_EINVAL; }
        rx->rx.host_chunks = value; return QC_OK;
    case QC_RX_OPT_FUSED_CHUNK:
        if (value < 128 || value > 2048) { qc::set_error("rx_set_option: chunk %d out of range", value); return QC_EINVAL; }
        rx->rx.fused_chunk = value; return QC_OK;
    case QC_RX_OPT_FUSED_THREADS:
        if (value != 128 && value != 256) { qc::set_error("rx_set_option: threads must be 128 or 256"); return QC_EINVAL; }
        rx->rx.fused_threads = value; return QC_OK;
    case QC_RX_OPT_TRACE:
        if (value && !rx->rx.d_trace) { QC_CUDA(cudaMalloc((void **)&rx->rx.d_trace, (size_t)rx->rx.C * 256 * sizeof(long long))); QC_CUDA(cudaMemset(rx->rx.d_trace, 0, (size_t)rx->rx.C * 256 * sizeof(long long))); }
        if (!value && rx->rx.d_trace) { cudaFree(rx->rx.d_trace); rx->rx.d_trace = nullptr; }
        return QC_OK;
    case QC_RX_OPT_FUSED_TAIL: rx->rx.fused_tail = value ? 1 : 0; return QC_OK;
    case QC_RX_OPT_EXACT_NCO: rx->rx.exact_nco = value ? 1 : 0; return QC_OK;
    case QC_RX_OPT_FUSED_DEEPK:
        if (value != 1 && value != 4) { qc::set_error("rx_set_option: deepk must be 1 or 4"); return QC_EINVAL; }
        rx->rx.fused_deepk = value; return QC_OK;
    case QC_RX_OPT_FUSED_PLANS: rx->rx.fused_plans = value != 0; return QC_OK;
    case QC_RX_OPT_FUSED_DENSE: rx->rx.fused_dense = value; return QC_OK;
    case QC_RX_OPT_FUSED_TAILWARP: if (value < 0 || value > 4) { qc::set_error("rx_set_option: tail-warp split must be 0 (off), 1 (default) or the first tail stage 2..4"); return QC_EINVAL; }
        rx->rx.fused_tailwarp = value; return QC_OK;
    case QC_RX_OPT_FUSED_SPLIT:
        if (value < 0 || value > 3) { qc::set_error("rx_set_option: split must be 0 .. 3"); return QC_EINVAL; }
        rx->rx.fused_split = value; return QC_OK;
    case QC_RX_OPT_FUSED_P3: rx->rx.fused_p3 = value != 0; return QC_OK;
    case QC_RX_OPT_FUSED_ASYNC:
        if (value < 0 || value > 2) { qc::set_error("rx_set_option: async chunk loads must be 0, 1 or 2"); return QC_EINVAL; }
        rx->rx.fused_async = value; return QC_OK;
    case QC_RX_OPT_FUSED_MIN_R:
        if (value != 0 && value != 2 && value != 4 && value != 8) { qc::set_error("rx_set_option: min R must be 0/2/4/8"); return QC_EINVAL; }
        rx->rx.fused_min_r = value; return QC_OK;
    }
    qc::set_error("rx_set_option: unknown option %d", option);
    return QC_EINVAL;
}

int quisk_cuda_rx_read_trace(qcRxChain *rx, long long *host_out, int n_channels)
{
    if (!rx || !rx->rx.d_trace) { qc::set_error("rx_read_trace: tracing is off"); return QC_EINVAL; }
    if (n_channels > rx->rx.C) n_channels = rx->rx.C;
    QC_CUDA(cudaDeviceSynchronize());
    QC_CUDA(cudaMemcpy(host_out, rx->rx.d_trace, (size_t)n_channels * 256 * sizeof(long long), cudaMemcpyDeviceToHost));
    return QC_OK;
}

const char *quisk_cuda_rx_fused_kernel_name(qcRxChain *rx)
{
    return rx ? rx->rx.fused_name : "";
}

int quisk_cuda_rx_kernel_time(qcRxChain *rx, double *ms_total, int *launches)
{
    if (!rx) return QC_EINVAL;
    double tot = 0.0;
    int n = 0;
    for (auto &pr : rx->rx.timed) {
        float ms = 0.f;
        QC_CUDA(cudaEventSynchronize(pr.second));
        QC_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        tot += ms; n++;
        cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
    }
    rx->rx.timed.clear();
    if (ms_total) *ms_total = tot;
    if (launches) *launches = n;
    return QC_OK;
}

}  // extern "C"
