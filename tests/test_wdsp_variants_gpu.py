"""GPU parity of the three relatives of fircore that wdsp defines and no live chain instantiates (SURVEY F3, rows a17 /
section 2): xfirmin (wdsp/firmin.c:76-99, time-domain ring FIR: bit-exact), xfiropt (firmin.c:227-251) and xbps
(wdsp/bandpass.c:85-105) (overlap-save: 1e-12 relative RMS, our FFT is not FFTW's), against fixtures generated from the
compiled reference (tests/golden/make_golden_wdsp_variants.py), each with a flush two blocks before the end."""
import ctypes as C

import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.golden.make_golden_wdsp_variants import FIRMIN_CASES, FIROPT_CASES, BPS_CASES, sig
from tests.util import golden

pytestmark = pytest.mark.gpu
NCH = 3


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
    return golden("wdsp_variants_kat.npz")


def _dev(torch, x):
    return torch.from_numpy(np.ascontiguousarray(np.stack([x] * NCH))).cuda()


@pytest.mark.parametrize("size,nc,rate,fl,fh,wt,gain,blocks", FIRMIN_CASES)
def test_xfirmin_bit_exact(size, nc, rate, fl, fh, wt, gain, blocks, torch, lib, kat):
    f = lib.quisk_cuda_firmin_create(NCH, nc, fl, fh, rate, wt, gain)
    assert f, lib.quisk_cuda_last_error()
    x = sig(size * blocks, 900 + size, float(rate))
    d = _dev(torch, x); o = torch.zeros_like(d)
    n_out = C.c_int(0)
    for b in range(blocks):
        if b == blocks - 2:
            assert lib.quisk_cuda_batch_reset(f, None) == 0          # flush_firmin
        blk = d[:, b * size:(b + 1) * size]; ob = o[:, b * size:(b + 1) * size]
        assert lib.quisk_cuda_batch_run(f, blk.data_ptr(), d.stride(0), size, ob.data_ptr(), o.stride(0), C.byref(n_out), 0, None) == 0
        assert n_out.value == size
    torch.cuda.synchronize()
    y = o.cpu().numpy()
    ref = kat["firmin_%d_%d/y" % (size, nc)]
    for c in range(NCH):
        assert np.array_equal(y[c], ref)
    lib.quisk_cuda_batch_destroy(f)


def _run_fircore_like(torch, lib, f, x, size, blocks):
    d = _dev(torch, x); o = torch.zeros_like(d)
    for b in range(blocks):
        if b == blocks - 2:
            assert lib.quisk_cuda_fircore_flush(f) == 0
        blk = d[:, b * size:(b + 1) * size]; ob = o[:, b * size:(b + 1) * size]
        assert lib.quisk_cuda_fircore_run(f, blk.data_ptr(), d.stride(0), ob.data_ptr(), o.stride(0), None) == 0
    torch.cuda.synchronize()
    return o.cpu().numpy()


@pytest.mark.parametrize("size,nc,rate,fl,fh,wt,gain,blocks", FIROPT_CASES)
def test_xfiropt(size, nc, rate, fl, fh, wt, gain, blocks, torch, lib, kat):
    f = lib.quisk_cuda_firopt_create(NCH, size, nc, fl, fh, rate, wt, gain)
    assert f, lib.quisk_cuda_last_error()
    y = _run_fircore_like(torch, lib, f, sig(size * blocks, 910 + size, float(rate)), size, blocks)
    ref = kat["firopt_%d_%d/y" % (size, nc)]
    for c in range(NCH):
        assert O.rel_rms(y[c], ref) < 1e-12
    lib.quisk_cuda_fircore_destroy(f)


@pytest.mark.parametrize("size,rate,fl,fh,wt,gain,blocks", BPS_CASES)
def test_xbps(size, rate, fl, fh, wt, gain, blocks, torch, lib, kat):
    f = lib.quisk_cuda_bps_create(NCH, size, fl, fh, rate, wt, gain)
    assert f, lib.quisk_cuda_last_error()
    x = sig(size * blocks, 920 + size, float(rate))
    y = _run_fircore_like(torch, lib, f, x, size, blocks)
    ref = kat["bps_%d/y" % size]
    for c in range(NCH):
        assert O.rel_rms(y[c], ref) < 1e-12
    lib.quisk_cuda_fircore_destroy(f)
