"""GPU parity of gain method 3 of WDSP's spectral noise reduction (the "trained" method, emnr.c:965-1010, 866-884,
813-815: the second state of Quisk's NR2 button, quisk.py:6020-6023) against fixtures from the compiled reference
(tests/golden/make_golden_wdsp_emnr3.py): the stage with every noise-power estimator, and Quisk's channel with the method
switched on mid-stream through the reference-signature entry points, with the default and with moved training parameters.
The 60 x 60 zeta table is data of the WDSP distribution (wdsp/zetahat.c): the test reads it from the compiled reference
(symbols zetaHatDefault*) and hands it to the library, as a host would from its own WDSP build."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import quisk_oracle as O
from oracle import ref_ctypes as R
from tests.golden.make_golden_wdsp_emnr import BLOCKS, CH_BLOCKS, CH_ON, CH_TAIL, N, RATE, channel_input, stage_input
from tests.golden.make_golden_wdsp_emnr3 import CASES3, TRAIN, key3
from tests.util import golden

pytestmark = pytest.mark.gpu
NCH = 3
D = C.c_double


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
    return golden("wdsp_emnr3_kat.npz")


@pytest.fixture(scope="module")
def zeta(lib):
    so = os.path.join(R.REF_DIR, "libwdsp_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libwdsp_ref.so not built: no source for the WDSP distribution's zeta table")
    ref = C.CDLL(so)
    rows, cols = C.c_int.in_dll(ref, "zetaHatDefaultRows").value, C.c_int.in_dll(ref, "zetaHatDefaultCols").value
    data = (C.c_double * (rows * cols)).in_dll(ref, "zetaHatDefaultData")
    valid = (C.c_int * (rows * cols)).in_dll(ref, "zetaHatDefaultValid")
    lim = [C.c_double.in_dll(ref, "zetaHatDefault" + n).value for n in ("Gmin", "Gmax", "Ximin", "Ximax")]
    lib.quisk_cuda_emnr_set_zeta.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, D, D, D, D]
    assert lib.quisk_cuda_emnr_set_zeta(C.addressof(data), C.addressof(valid), rows, cols, *lim) == 0, lib.quisk_cuda_last_error()
    return True


@pytest.mark.parametrize("npe,ae", CASES3)
def test_emnr_method3_stage(npe, ae, torch, lib, kat, zeta):
    key = key3(npe, ae)
    x = stage_input()
    lib.quisk_cuda_emnr_create.restype = C.c_void_p
    lib.quisk_cuda_emnr_create.argtypes = [C.c_int] * 6 + [D] + [C.c_int] * 3
    lib.quisk_cuda_emnr_run.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p]
    lib.quisk_cuda_emnr_destroy.argtypes = [C.c_void_p]
    e = lib.quisk_cuda_emnr_create(NCH, N, 4096, 4, RATE, 0, 1.0, 3, npe, ae)
    assert e, lib.quisk_cuda_last_error()
    xc = (x + 0.5j * x).astype(np.complex128)
    d = torch.from_numpy(np.ascontiguousarray(np.stack([xc] * NCH))).cuda()
    for b in range(BLOCKS):
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


@pytest.mark.parametrize("name,train,chn", [("chan3", None, 10), ("chan3_train", TRAIN, 11)])
def test_quisk_channel_with_trained_nr2(name, train, chn, torch, lib, kat, zeta):
    """OpenChannel as quisk_wdsp.py:66-93 does, then SetRXAEMNRgainMethod(3) + SetRXAEMNRRun(1) (the second state of
    Quisk's NR2 button, quisk.py:6020-6023) at block CH_ON, through the reference-signature entry points."""
    lib.OpenChannel.argtypes = [C.c_int] * 8 + [D] * 4 + [C.c_int]
    lib.fexchange0.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.RXASetPassband.argtypes = [C.c_int, D, D]
    lib.SetRXAAGCFixed.argtypes = [C.c_int, D]
    lib.SetRXAEMNRtrainZetaThresh.argtypes = [C.c_int, D]
    lib.SetRXAEMNRtrainT2.argtypes = [C.c_int, D]
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
            if train:
                lib.SetRXAEMNRtrainZetaThresh(chn, D(train[0])); lib.SetRXAEMNRtrainT2(chn, D(train[1]))
            lib.SetRXAEMNRgainMethod(chn, 3)
            lib.SetRXAEMNRRun(chn, 1)
        inb[:] = xc[b * N:(b + 1) * N]
        lib.fexchange0(chn, inb.ctypes.data_as(C.c_void_p), outb.ctypes.data_as(C.c_void_p), C.byref(err))
        assert err.value == 0
        ys.append(outb.copy())
    lib.SetChannelState(chn, 0, 0)
    lib.CloseChannel(chn)
    y = np.concatenate(ys)[-CH_TAIL * N:]
    ref = kat[name + "/y_tail"]
    e = O.rel_rms(y, ref)
    print(name, "rel rms", e, "reference's own one-ulp sensitivity", kat[name + "/cond"])
    assert np.abs(ref).max() > 0.5
    assert e < 1e-12
