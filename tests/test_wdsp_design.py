"""Host-side WDSP coefficient design (wdsp_design.cpp) against the compiled reference: bit-identical taps.
CPU only; skipped when oracle/_ref has not been built."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref_ctypes as R

pytestmark = pytest.mark.skipif(not R.have_ref("libwdsp_ref.so"), reason="oracle/_ref not built")
D = C.c_double


@pytest.fixture(scope="module")
def libs():
    from quisk_b200 import build
    build.build()
    lib = C.CDLL(build.OUT)
    ref = R.load("libwdsp_ref.so")
    ref.fir_bandpass.restype = C.POINTER(D)
    ref.fir_bandpass.argtypes = [C.c_int, D, D, D, C.c_int, C.c_int, D]
    lib.quisk_cuda_fir_bandpass.argtypes = [C.c_int, D, D, D, C.c_int, C.c_int, D, C.c_void_p]
    ref.fc_impulse.restype = C.POINTER(D)
    ref.fc_impulse.argtypes = [C.c_int] + [D] * 4 + [C.c_int, D, D, C.c_int, C.c_int]
    lib.quisk_cuda_fc_impulse.argtypes = [C.c_int] + [D] * 4 + [C.c_int, D, D, C.c_int, C.c_int, C.c_void_p]
    lib.quisk_cuda_resample_design.argtypes = [C.c_int, C.c_int, D, C.c_int, D] + [C.c_void_p] * 4 + [C.c_int]
    return lib, ref


@pytest.mark.parametrize("args", [(4096, 150., 2850., 192000., 0, 1, 1 / 2048.), (2048, -4150., -150., 48000., 1, 1, 1.0),
                                  (1121, -0.05, 0.05, 1.0, 1, 0, 1.0), (2049, 240., 3300., 48000., 0, 1, .5), (64, -3000., 3000., 8000., 0, 0, 2.0)])
def test_fir_bandpass_bit_identical(libs, args):
    lib, ref = libs
    N, rtype = args[0], args[5]
    out = np.zeros(N * (2 if rtype else 1))
    assert lib.quisk_cuda_fir_bandpass(*args, out.ctypes.data) == 0
    r = np.ctypeslib.as_array(ref.fir_bandpass(*args), (len(out),)).copy()
    assert np.array_equal(out, r)


@pytest.mark.parametrize("nc", [2048, 256, 4096])
def test_fc_impulse_bit_identical(libs, nc):
    lib, ref = libs
    args = (nc, 300., 3000., 20 * np.log10(10.), 0., 1, 48000., 1 / 512., 0, 0)
    out = np.zeros(2 * nc)
    assert lib.quisk_cuda_fc_impulse(*args, out.ctypes.data) == 0
    r = np.ctypeslib.as_array(ref.fc_impulse(*args), (2 * nc,)).copy()
    assert np.array_equal(out, r)


@pytest.mark.parametrize("rates,exp", [((384000, 48000), (1, 8, 1121)), ((192000, 48000), (1, 4, 561)), ((48000, 192000), (4, 1, 564)),
                                       ((44100, 48000), (160, 147, 22560))])
def test_resample_plan(libs, rates, exp):
    lib, _ = libs
    L, M, n = C.c_int(), C.c_int(), C.c_int()
    assert lib.quisk_cuda_resample_design(rates[0], rates[1], 0.0, 0, 1.0, C.byref(L), C.byref(M), C.byref(n), None, 0) == 0
    assert (L.value, M.value, n.value) == exp
