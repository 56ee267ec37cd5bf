// quisk_b200/csrc/batch.h -- internal: batched single-stage filter object (see batch.cu)
#pragma once
#include "qc_common.cuh"

namespace qc {

struct BatchFilter {
    int kind = 0, C = 0, nTaps = 0, interp = 1, decim = 1;
    int H = 0, esize = 16;
    int phase = 0;               // toggle (HB45) or decim_index
    int cur = 0;                 // which history buffer is current
    bool unit_gain = false;      // interpolating kinds: do not multiply by `interp` (WDSP resampler taps carry the gain)
    double *d_coef = nullptr;
    std::vector<double> h_coef;  // host copy of the taps as uploaded
    void *d_hist[2] = {nullptr, nullptr};

    int init(int kind, int C, const double *coefs, int n_taps, int interp, int decim);
    void release();
    long count_full(int count) const;
    int count_out(int count, int legacy_clip) const;
    int run(const void *d_in, long in_stride, int count, void *d_out, long out_stride,
            int *n_out, int legacy_clip, cudaStream_t stream);
    int reset(cudaStream_t stream);
};

}  // namespace qc
