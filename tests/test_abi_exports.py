"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads, and
exports every function include/*.h declares; struct layouts match filter.h (56/56/544/280)."""
import ctypes as C
import glob
import os

import pytest

from tests.util import ROOT, declared_functions


@pytest.fixture(scope="module")
def lib():
    from quisk_b200 import build
    build.build()
    return C.CDLL(build.OUT)


def test_exports_every_declared_symbol(lib):
    missing = []
    for hdr in glob.glob(os.path.join(ROOT, "include", "*.h")):
        for fn in declared_functions(hdr):
            try:
                getattr(lib, fn)
            except AttributeError:
                missing.append((os.path.basename(hdr), fn))
    assert not missing, missing


def test_filter_h_names_present(lib):
    names = ["quisk_filt_cInit", "quisk_filt_dInit", "quisk_filt_differInit", "quisk_filt_tune", "quisk_dC_out",
             "quisk_dD_out", "quisk_cInterpolate", "quisk_dInterpolate", "quisk_cDecimate", "quisk_cCDecimate",
             "quisk_dDecimate", "quisk_cInterpDecim", "quisk_cDecim2HB45", "quisk_dInterp2HB45",
             "quisk_cInterp2HB45", "quisk_dFilter", "quisk_cFilter"]       # filter.h:39-55
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layouts_match_filter_h():
    from oracle import ref_ctypes as R
    assert C.sizeof(R.cFilter) == 56 and C.sizeof(R.dFilter) == 56
    assert C.sizeof(R.cHB45Filter) == 544 and C.sizeof(R.dHB45Filter) == 280


def test_host_only_entry_points(lib):
    """quisk_filt_cInit / quisk_filt_tune are host bookkeeping and must work without a GPU."""
    import numpy as np
    from oracle import ref_ctypes as R
    R.bind_filter_api(lib)
    h = np.linspace(-1, 1, 31)
    st = R.cFilter()
    lib.quisk_filt_cInit(C.byref(st), h.ctypes.data_as(R.c_double_p), len(h))
    assert st.nTaps == 31 and st.decim_index == 0 and st.cSamples and st.ptcSamp == st.cSamples
    lib.quisk_filt_tune(C.cast(C.byref(st), C.c_void_p), 0.1, 1)
    got = np.ctypeslib.as_array(C.cast(st.cpxCoefs, C.POINTER(C.c_double)), (62,)).view(np.complex128)
    D = 15.0
    exp = np.exp(2j * np.pi * 0.1 * (np.arange(31) - D)) * h              # filter.c:72-80
    assert np.allclose(got, exp, rtol=0, atol=1e-15)
    if R.have_ref():
        ref = R.bind_filter_api(R.load("libquisk_filter_ref.so"))
        st2 = R.cFilter()
        ref.quisk_filt_cInit(C.byref(st2), h.ctypes.data_as(R.c_double_p), len(h))
        ref.quisk_filt_tune(C.cast(C.byref(st2), C.c_void_p), 0.1, 0)
        lib.quisk_filt_tune(C.cast(C.byref(st), C.c_void_p), 0.1, 0)
        a = np.ctypeslib.as_array(C.cast(st.cpxCoefs, C.POINTER(C.c_double)), (62,))
        b = np.ctypeslib.as_array(C.cast(st2.cpxCoefs, C.POINTER(C.c_double)), (62,))
        assert np.array_equal(a, b)         # same libm, same expression: bit identical taps


def test_no_device_is_an_error_not_a_fallback(lib):
    lib.quisk_cuda_last_error.restype = C.c_char_p
    if lib.quisk_cuda_device_count() > 0:
        pytest.skip("a GPU is present")
    lib.quisk_cuda_batch_create.restype = C.c_void_p
    h = lib.quisk_cuda_batch_create(1, 4, None, 0, 1, 1)
    assert not h
    assert b"CUDA" in lib.quisk_cuda_last_error()
    # the optional stages (noise blanker, auto-notch, SSB squelch) have no host path either
    for name, args in (("quisk_cuda_nb_create", (2, 192000)), ("quisk_cuda_autonotch_create", (2, 12000)),
                       ("quisk_cuda_ssb_squelch_create", (2, 12000, 2800))):
        fn = getattr(lib, name)
        fn.restype = C.c_void_p
        assert not fn(*args), name


def test_plan_decimation_matches_oracle(lib):
    from oracle import quisk_oracle as O
    for rate in [48000, 96000, 192000, 240000, 250000, 384000, 960000, 1200000, 1536000, 2000000, 3072000]:
        p2, p3, p5 = C.c_int(), C.c_int(), C.c_int()
        best = lib.quisk_cuda_plan_decimation(rate, C.byref(p2), C.byref(p3), C.byref(p5))
        assert (best, p2.value, p3.value, p5.value) == O.plan_decimation(rate)
