"""GPU parity of WDSP's spectral noise blanker "SNB" (wdsp/snb.c) against fixtures from the compiled reference
(tests/golden/make_golden_wdsp_snba.py): the stage with create_rxa's arguments at 12 kS/s (no resamplers) and at 48 kS/s
(resamplers in and out), a flush mid-stream.  The interpolation solves normal equations, so the reference itself moves by
`cond` (1e-12 ... 5e-11) when its input moves by one ulp: the bound is max(1e-12, 20 * cond)."""
import ctypes as C

import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.golden.make_golden_wdsp_snba import CASES, CH_BLOCKS, CH_ON, CH_TAIL, N, channel_input, stage_input
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


def test_quisk_channel_with_snb_switched_on(torch, lib, kat):
    """OpenChannel as quisk_wdsp.py:66-93 does, then SetRXASNBARun(1) (what Quisk's SNB button sends, quisk.py:6040-6043) at
    block CH_ON, through the reference-signature entry points and fexchange0.  Switching it on also switches the blanker's
    own band pass in front (bpsnba REPLACES nbp0's output in USB, RXA.c:561-565, 883-918) and bp1 behind at gain 2."""
    D = C.c_double
    lib.OpenChannel.argtypes = [C.c_int] * 8 + [D] * 4 + [C.c_int]
    lib.fexchange0.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.RXASetPassband.argtypes = [C.c_int, D, D]
    lib.SetRXAAGCFixed.argtypes = [C.c_int, D]
    chn = 10
    lib.OpenChannel(chn, N, N, 48000, 48000, 48000, 0, 1, D(0.0), D(0.0), D(0.0), D(0.0), 1)
    lib.SetRXAShiftRun(chn, 0); lib.RXANBPSetRun(chn, 0); lib.SetRXAAMSQRun(chn, 0)
    lib.SetRXAMode(chn, 1)
    lib.RXASetPassband(chn, D(300.0), D(3000.0))
    lib.RXASetNC(chn, N); lib.RXASetMP(chn, 0)
    lib.SetRXAAGCMode(chn, 0); lib.SetRXAAGCFixed(chn, D(0.0))
    lib.SetRXAPanelRun(chn, 0); lib.SetRXAEMNRRun(chn, 0)
    xc = channel_input()
    inb = np.zeros(N, dtype=np.complex128); outb = np.zeros(N, dtype=np.complex128)
    err = C.c_int(0)
    ys = []
    for b in range(CH_BLOCKS):
        if b == CH_ON:
            lib.SetRXASNBARun(chn, 1)
        inb[:] = xc[b * N:(b + 1) * N]
        lib.fexchange0(chn, inb.ctypes.data_as(C.c_void_p), outb.ctypes.data_as(C.c_void_p), C.byref(err))
        assert err.value == 0
        ys.append(outb.copy())
    lib.SetChannelState(chn, 0, 0)
    lib.CloseChannel(chn)
    y = np.concatenate(ys)[-CH_TAIL * N:]
    ref = kat["chan/y_tail"]
    cond = float(kat["chan/cond"][0])
    e = O.rel_rms(y, ref)
    print("channel with SNB: rel rms", e, "reference's own one-ulp sensitivity", cond)
    assert np.abs(ref).max() > 0.1
    assert e < max(1e-12, 20.0 * cond)
