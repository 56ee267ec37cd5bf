"""GPU parity of WDSP's spectral noise reduction "NR2" (wdsp/emnr.c) against fixtures from the compiled reference
(tests/golden/make_golden_wdsp_emnr.py): the stage alone for every gain method (0, 1, 2) and noise-power estimator
(0, 1, 2) this library builds, post-filter on and off, with a flush mid-stream; and Quisk's channel with NR2 switched on
mid-stream through the reference-signature entry points.  Tolerance 1e-12 relative RMS (the reference moves by 1e-15
when its input moves by one ulp: the fixture's `cond`).  The two gamma-prior tables of gain method 2 are data of the WDSP
distribution: the test reads them from the compiled reference (oracle/_ref/libwdsp_ref.so, symbols GG and GGS) and hands
them to the library, as a host would from its own WDSP build."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import quisk_oracle as O
from oracle import ref_ctypes as R
from tests.golden.make_golden_wdsp_emnr import BLOCKS, CASES, CH_BLOCKS, CH_ON, CH_TAIL, N, RATE, channel_input, stage_input
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
    return golden("wdsp_emnr_kat.npz")


@pytest.fixture(scope="module")
def tables(lib):
    so = os.path.join(R.REF_DIR, "libwdsp_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libwdsp_ref.so not built: no source for the WDSP distribution's GG / GGS tables")
    ref = C.CDLL(so)
    T = C.c_double * (241 * 241)
    gg, ggs = T.in_dll(ref, "GG"), T.in_dll(ref, "GGS")
    assert lib.quisk_cuda_emnr_set_tables(C.addressof(gg), C.addressof(ggs)) == 0
    return True


def test_emnr_gain_method_2_needs_the_tables(lib):
    """Before the host has handed the tables over, method 2 is refused loudly (no silent fallback to another method)."""
    # a fresh process state is not guaranteed (another test may have set them): only check the message when it fails
    e = lib.quisk_cuda_emnr_create(1, N, 4096, 4, RATE, 0, 1.0, 2, 0, 1)
    assert e, lib.quisk_cuda_last_error()
    lib.quisk_cuda_emnr_destroy(e)
    assert not lib.quisk_cuda_emnr_create(1, N, 2048, 4, RATE, 0, 1.0, 2, 0, 1)        # only create_rxa's frame size is built
    assert b"4096" in lib.quisk_cuda_last_error()


@pytest.mark.parametrize("gm,npe,ae", CASES)
def test_emnr_stage(gm, npe, ae, torch, lib, kat, tables):
    key = "emnr_%d_%d_%d" % (gm, npe, ae)
    x = stage_input()
    e = lib.quisk_cuda_emnr_create(NCH, N, 4096, 4, RATE, 0, 1.0, gm, npe, ae)
    assert e, lib.quisk_cuda_last_error()
    xc = (x + 0.5j * x).astype(np.complex128)
    d = torch.from_numpy(np.ascontiguousarray(np.stack([xc] * NCH))).cuda()
    flush_at = 90 if (gm, npe) == (2, 0) else -1
    for b in range(BLOCKS):
        if b == flush_at:
            assert lib.quisk_cuda_emnr_flush(e) == 0
        blk = d[:, b * N:(b + 1) * N]
        assert lib.quisk_cuda_emnr_run(e, blk.data_ptr(), d.stride(0), blk.data_ptr(), d.stride(0), None) == 0, lib.quisk_cuda_last_error()
    torch.cuda.synchronize()
    y = d.cpu().numpy()
    ref = kat[key + "/y"]
    assert np.abs(ref).max() > 0.3 and not y.imag.any()
    errs = [O.rel_rms(y[c].real, ref) for c in range(NCH)]
    print(key, errs, "reference's own one-ulp sensitivity", kat[key + "/cond"])
    assert max(errs) < 1e-12
    lib.quisk_cuda_emnr_destroy(e)


def test_quisk_channel_with_nr2_switched_on(torch, lib, kat, tables):
    """OpenChannel as quisk_wdsp.py:66-93 does, then SetRXAEMNRgainMethod(2) + SetRXAEMNRRun(1) (what Quisk's NR2 button
    sends, quisk.py:6017-6027) at block CH_ON, through the reference-signature entry points and fexchange0."""
    D = C.c_double
    lib.OpenChannel.argtypes = [C.c_int] * 8 + [D] * 4 + [C.c_int]
    lib.fexchange0.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.RXASetPassband.argtypes = [C.c_int, D, D]
    lib.SetRXAAGCFixed.argtypes = [C.c_int, D]
    chn = 9
    lib.OpenChannel(chn, N, N, RATE, RATE, RATE, 0, 1, D(0.0), D(0.0), D(0.0), D(0.0), 1)
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
            lib.SetRXAEMNRgainMethod(chn, 2)
            lib.SetRXAEMNRRun(chn, 1)
        inb[:] = xc[b * N:(b + 1) * N]
        lib.fexchange0(chn, inb.ctypes.data_as(C.c_void_p), outb.ctypes.data_as(C.c_void_p), C.byref(err))
        assert err.value == 0
        ys.append(outb.copy())
    lib.SetChannelState(chn, 0, 0)
    lib.CloseChannel(chn)
    y = np.concatenate(ys)[-CH_TAIL * N:]
    ref = kat["chan/y_tail"]
    e = O.rel_rms(y, ref)
    print("channel with NR2: rel rms", e, "reference's own one-ulp sensitivity", kat["chan/cond"])
    assert np.abs(ref).max() > 0.5
    assert e < 1e-12
