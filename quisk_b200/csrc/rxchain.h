// quisk_b200/csrc/rxchain.h -- internal: batched receive chain object (see rxchain.cu, rxfused.cu)
#pragma once
#include "qc_common.cuh"
#include "batch.h"

namespace qc {

int plan_decimation(int sample_rate, int *p2, int *p3, int *p5);
int launch_unpack_iq(const void *d_bytes, long byte_stride, int C, int count, int nb, int big, cd *out, long out_stride, cudaStream_t s);

struct FusedDecimator;      // rxfused.cu

struct RxChain {
    int C = 0, sample_rate = 0, mode = 0, fused = 0;
    int decim_srate = 0, filter_srate = 0;
    std::vector<BatchFilter *> cst;     // complex stages: process_decimate then the demod pre-filters
    int n_decim_stages = 0;             // how many of cst belong to quisk_process_decimate
    BatchFilter *rxf = nullptr;         // cRxFilterOut / dRxFilterOut
    bool iq_out = false;                // DGT-IQ: complex samples out, no detector
    std::vector<BatchFilter *> rst;     // real audio stages after the detector
    // tuning NCO
    bool tune = false;
    std::vector<double> tune_hz;
    double *d_nco = nullptr;
    unsigned long long n_base = 0;      // samples since the NCO constants were (re)based (statistics only)
    // exact phasor at block starts: ring of [C] buffers filled one block ahead on a side stream (pointwise.cu)
    static constexpr int NV = 4;
    cd *d_v[NV] = {nullptr, nullptr, nullptr, nullptr};
    int vcur = 0;
    int exact_nco = 1;                  // 1: block-start phasors from the reference's recurrence; 0: closed form only
    cudaStream_t s_nco = nullptr;
    qcNoiseBlanker *nb = nullptr; int nb_level = 0;     // QC_RX_OPT_NOISE_BLANKER: NoiseBlanker on the staged block of the host entries
    // QC_RX_OPT_AUTO_NOTCH / _NOTCH_SIDETONE / _SSB_SQUELCH: dAutoNotch and ssb_squelch + d_delay on the audio at the filter rate,
    // between the detector and the audio interpolators, where quisk_process_demodulate runs them (quisk.c:1923-1928); a squelched
    // receiver's block comes back as zeros (quisk.c:2716-2719)
    qcAutoNotch *anotch = nullptr; qcSsbSquelch *squelch = nullptr; int auto_notch = 0, notch_sidetone = 0, squelch_level = 0, filter_bandwidth = 0;
    bool audio_options() const;
    int host_noise_blanker(cudaStream_t s, int count);
    unsigned *d_sched = nullptr;        // nco_advance_kernel's ticket counter + one worker slot per SM id
    unsigned nco_epoch = 0;
    cudaEvent_t ev_r[NV] = {}, ev_f[NV] = {};
    bool r_set[NV] = {}, f_set[NV] = {};
    int nco_before(int count, cudaStream_t s);      // starts the recurrence for the next block; the consumer on s may then read d_v[vcur]
    int nco_after(int count, cudaStream_t s);       // consumer enqueued: mark the slot as read, move on
    // detectors
    double *d_dc = nullptr, *d_fm = nullptr;
    double fm_a0 = 0, fm_a1 = 0, fm_b1 = 0;
    // scratch
    cd *bufc[2] = {nullptr, nullptr};
    double *bufr[2] = {nullptr, nullptr};
    long cap = 0;
    // host-buffer entry points.  With more than one chunk the channels are split over `sub` chains of their own (own
    // streams, own stage state): chunk k's H2D copy, kernels and D2H copy overlap those of its neighbours, so a call costs
    // max(copy in, compute, copy out) instead of their sum.  The host entries then carry their own stream state, separate
    // from quisk_cuda_rx_process on the parent (a caller feeds a chain through one entry or the other, not both).
    int host_chunks = 0;                // QC_RX_OPT_HOST_CHUNKS: 0 = auto (8 from 1024 channels up, else 1), 1 = off
    std::vector<RxChain *> sub;
    std::vector<int> sub_c0;
    struct Saved {                      // what init() was given, kept so that the sub-chains can be built later
        qcRxConfig cfg; std::vector<double> fi, fq, tune; std::vector<double> tab[13];
    } saved;
    int n_host_chunks() const;
    int build_sub_chains(int k);
    void release_sub_chains();
    int host_enqueue(const quisk_cd *h_iq, long iq_stride, const void *h_bytes, long byte_stride, int nb, int big, int count,
                     double *h_audio, long audio_stride, int *n_audio);
    cudaStream_t hs = nullptr;
    char *h_pin = nullptr;
    cd *d_host_in = nullptr; double *d_host_out = nullptr;
    int host_cap = 0, host_out_cap = 0;
    unsigned char *d_packed = nullptr; size_t packed_cap = 0;      // wire-format staging (ingest.cu)
    // fused full-rate decimator
    FusedDecimator *fd = nullptr;
    size_t n_fused_stages = 0;
    int fused_chunk = 2048;             // target input samples per shared-memory chunk
    int fused_threads = 128;            // CTA width of the fused kernel (128 or 256)
    long long *d_trace = nullptr;       // debug: per-chunk clock64() stamps of the fused kernel
    int fused_deepk = 1;                // plan kernels: run the <=128-sample stages once per this many chunks (1 = every chunk)
    int fused_plans = 1;                // use the plan-specialised instantiations when one matches
    int fused_dense = 0;                // 1: cap registers at 128/thread for more resident CTAs
    int fused_tailwarp = 1;             // plan kernels: the low-rate stages run on a fifth warp, one chunk behind (fused_decim_tw_kernel)
    int fused_p3 = 0;                   // three-group pipeline kernel (fused_decim_p3_kernel): half bands 1-2 on two warps of their own
    int fused_async = 0;                // tail-warp kernel: next chunk by cp.async into stage 0's buffer (1; 2 = under a 128-register cap) instead of a register prefetch
    int fused_split = 3;                // plan kernels: half bands with one lane per component (hb_stage_split): 0 off, 1 on, 2 behind half band 0 of the tail-warp kernel (+ the 192 kS/s plan), 3 (default) every half band of it
    int fused_min_r = 0;                // force at least this many outputs per thread in half-band stages
    int fused_tail = 1;                 // SSB / CW: run filter + demod + audio interpolators as one kernel (rxtail.cu)
    // optional device timing of the dominant (fused) kernel
    bool poisoned = false;              // a process() call failed after stage state had advanced: reset() before the next block
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timed;

    int init(const qcRxConfig &cfg);
    void release();
    int reserve(int count);
    int max_out(int count) const;
    int reset_fm();
    int upload_nco();
    int process(const void *d_iq, long iq_stride, int count, double *d_audio, long audio_stride, int *n_audio,
                void *d_decim, long decim_stride, int *n_decim, cudaStream_t s);
    int process_host(const quisk_cd *h_iq, long iq_stride, int count, double *h_audio, long audio_stride, int *n_audio);
    int process_host_packed(const void *h_bytes, long byte_stride, int count, int nb, int big,
                            double *h_audio, long audio_stride, int *n_audio);
    int reset();
    // rxfused.cu
    size_t fusable_prefix(size_t limit);
    const char *fused_name = "";        // the fused decimator instantiation the last process() launched (bench.py matches ncu captures by it)
    int run_fused_decimator(size_t n_stages, const cd *in, long in_stride, int count, cd *out, long out_stride, int *n_out, cudaStream_t s);
    int reset_fused();
    void release_fused();
    // rxtail.cu
    bool tail_fusable() const;
    int run_tail(const cd *in, long in_stride, int n, double *out, long out_stride, int *n_out, cudaStream_t s);
};

}  // namespace qc
