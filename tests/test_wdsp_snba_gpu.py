"""GPU parity of WDSP's spectral noise blanker "SNB" (wdsp/snb.c) against fixtures from the compiled reference
(tests/golden/make_golden_wdsp_snba.py): the stage with create_rxa's arguments at 12 kS/s (no resamplers) and at 48 kS/s
(resamplers in and out), a flush mid-stream.  The interpolation solves normal equations, so the reference itself moves by
`cond` (1e-12 ... 5e-11) when its input moves by one ulp: the bound is max(1e-12, 20 * cond)."""
import ctypes as C

import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.golden.make_golden_wdsp_snba import CASES, stage_input
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
    return golden("wdsp_snba_kat.npz")


@pytest.mark.parametrize("rate,bsize,blocks", CASES)
def test_snba_stage(rate, bsize, blocks, torch, lib, kat):
    x = stage_input(rate, bsize, blocks)
    d = lib.quisk_cuda_snba_create(NCH, rate, 12000, bsize, 4, 256, 64, 2, 8.0, 20.0, 10, 2, 2, 0.5, 200.0, 5400.0)
    assert d, lib.quisk_cuda_last_error()
    xc = (x + 0.25j * x).astype(np.complex128)
    dev = torch.from_numpy(np.ascontiguousarray(np.stack([xc] * NCH))).cuda()
    for b in range(blocks):
        if b == 120:
            assert lib.quisk_cuda_snba_flush(d) == 0
        blk = dev[:, b * bsize:(b + 1) * bsize]
        assert lib.quisk_cuda_snba_run(d, blk.data_ptr(), dev.stride(0), blk.data_ptr(), dev.stride(0), None) == 0, lib.quisk_cuda_last_error()
    torch.cuda.synchronize()
    y = dev.cpu().numpy()
    ref = kat["snba_%d/y" % rate]
    cond = float(kat["snba_%d/cond" % rate][0])
    assert not y.imag.any()
    errs = [O.rel_rms(y[c].real, ref) for c in range(NCH)]
    print("snba", rate, errs, "reference's own one-ulp sensitivity", cond)
    assert max(errs) < max(1e-12, 20.0 * cond)
    for c in range(NCH):
        assert np.array_equal(y[c].real, ref)           # in fact identical: the stage is sums, products and quotients in the reference's order, no libm
    lib.quisk_cuda_snba_destroy(d)
    assert not lib.quisk_cuda_snba_create(1, rate, 12000, bsize, 4, 512, 64, 2, 8.0, 20.0, 10, 2, 2, 0.5, 200.0, 5400.0)     # only create_rxa's frame size
