"""GPU parity of WDSP's spectrum engine (wdsp/analyzer.c) against fixtures from the compiled reference
(tests/golden/make_golden_wdsp_analyzer.py): six configurations that between them use every window, every detector, every
averaging mode, overlap, integer and fractional clipping of the span, a flipped LO, the interpolating branch (more pixels than
bins), the 1 Hz normalisation, real input, and two stitched spans (three sub-spans of which the span clip removes one and a half;
two with overlap); ten frames each, every pixel output of every frame.  Pixels are float32 dB.  The detector's
index arithmetic, the averagers and mlog10 are the reference's expressions with separately rounded products and sums; what
differs is the transform (ours against the reference's FFTW-API shim, 1e-15 relative per bin).  mlog10 is a table look-up on
the leading 11 mantissa bits WITHOUT interpolation (meterlog10.c:547-554): its output moves in steps of up to 10 log10(1 + 1/2048)
= 2.1e-3 dB, so a bin that differs in its last bit and sits on a table boundary moves its pixel by one such step.  The bound
is one step (2.5e-3 dB) on every pixel, and exact float32 equality on more than 98 % of them."""
import ctypes as C

import numpy as np
import pytest

from tests.golden.make_golden_wdsp_analyzer import CASES, FRAMES, analyzer_input
from tests.util import golden

pytestmark = pytest.mark.gpu
ND = 3


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def lib():
    from quisk_b200 import lib as L
    return L.require_device()


@pytest.fixture(scope="module")
def kat():
    return golden("wdsp_analyzer_kat.npz")


@pytest.mark.parametrize("name", list(CASES))
def test_analyzer(name, torch, lib, kat):
    cfg = CASES[name]
    an = lib.quisk_cuda_analyzer_create(ND, 8192)
    assert an, lib.quisk_cuda_last_error()
    assert lib.quisk_cuda_analyzer_set_sample_rate(an, cfg["rate"]) == 0
    for po, (det, av, num, back, norm) in enumerate(cfg["outs"]):
        assert lib.quisk_cuda_analyzer_set_detector_mode(an, po, det) == 0
        assert lib.quisk_cuda_analyzer_set_average_mode(an, po, av) == 0
        assert lib.quisk_cuda_analyzer_set_num_average(an, po, num) == 0
        assert lib.quisk_cuda_analyzer_set_av_backmult(an, po, back) == 0
        assert lib.quisk_cuda_analyzer_set_norm_onehz(an, po, norm) == 0
    assert lib.quisk_cuda_analyzer_set(an, len(cfg["outs"]), cfg.get("typ", 1), cfg["flip"], cfg["sz"], cfg["hop"], cfg["win"], cfg["pa"], cfg["sz"] - cfg["hop"], cfg["clip"],
                                       cfg["fl"], cfg["fh"], cfg["npix"], cfg.get("stitch", 1), 2 * cfg["sz"]) == 0, lib.quisk_cuda_last_error()
    calls = cfg["sz"] // cfg["hop"] - 1 + FRAMES
    nst = cfg.get("stitch", 1)
    devs = [torch.from_numpy(np.ascontiguousarray(np.stack([analyzer_input(name, calls * cfg["hop"], ss)] * ND))).cuda() for ss in range(nst)]
    got = [[] for _ in cfg["outs"]]
    for k in range(calls):
        for ss in range(nst):
            blk = devs[ss][:, k * cfg["hop"]:(k + 1) * cfg["hop"]]
            assert lib.quisk_cuda_analyzer_spectrum0(an, ss, blk.data_ptr(), devs[ss].stride(0), None) == 0, lib.quisk_cuda_last_error()
        for po in range(len(cfg["outs"])):
            p = np.zeros((ND, cfg["npix"]), dtype=np.float32)
            flag = C.c_int(0)
            assert lib.quisk_cuda_analyzer_get_pixels(an, po, p.ctypes.data_as(C.c_void_p), C.byref(flag)) == 0
            if flag.value:
                got[po].append(p)
    lib.quisk_cuda_analyzer_destroy(an)
    for po in range(len(cfg["outs"])):
        ref = kat["%s/pix%d" % (name, po)]
        assert len(got[po]) == FRAMES == len(ref)
        y = np.stack(got[po])                       # [frames][displays][pixels]
        for d in range(ND):
            diff = np.abs(y[:, d, :].astype(np.float64) - ref.astype(np.float64))
            same = float(np.mean(y[:, d, :] == ref))
            print(name, "output", po, cfg["outs"][po], "display", d, "max |diff| dB", diff.max(), "identical pixels", same)
            assert np.isfinite(y).all()
            assert diff.max() < 2.5e-3
            assert same > 0.98
    assert ref.min() < -20.0 < ref.max() + 40.0      # a real spectrum, not a constant line


def test_analyzer_rejects_what_it_does_not_build(lib):
    an = lib.quisk_cuda_analyzer_create(1, 4096)
    assert an
    assert lib.quisk_cuda_analyzer_set(an, 1, 1, 0, 3000, 3000, 2, 0.0, 0, 0, 0.0, 0.0, 512, 1, 6000) != 0        # not a power of two
    assert lib.quisk_cuda_analyzer_set(an, 1, 1, 0, 8192, 8192, 2, 0.0, 0, 0, 0.0, 0.0, 512, 1, 6000) != 0        # larger than created
    assert lib.quisk_cuda_analyzer_set(an, 5, 1, 0, 1024, 1024, 2, 0.0, 0, 0, 0.0, 0.0, 512, 1, 6000) != 0        # pixel outputs
    assert lib.quisk_cuda_analyzer_set(an, 1, 1, 0, 1024, 1024, 9, 0.0, 0, 0, 0.0, 0.0, 512, 1, 6000) != 0        # window type
    assert lib.quisk_cuda_analyzer_set(an, 1, 1, 0, 1024, 1024, 2, 0.0, 0, 600, 0.0, 0.0, 512, 1, 6000) != 0      # clip leaves nothing
    assert lib.quisk_cuda_analyzer_set(an, 1, 1, 0, 1024, 1024, 2, 0.0, 0, 0, 0.0, 0.0, 512, 5, 6000) != 0        # sub-spans
    assert lib.quisk_cuda_analyzer_set_detector_mode(an, 0, 7) != 0
    assert lib.quisk_cuda_analyzer_set(an, 1, 1, 0, 1024, 1024, 2, 0.0, 0, 0, 0.0, 0.0, 512, 1, 6000) == 0
    assert abs(lib.quisk_cuda_analyzer_get_enb(an) - 1.5) < 0.01                                            # Hann: 1.5 bins
    lib.quisk_cuda_analyzer_destroy(an)
