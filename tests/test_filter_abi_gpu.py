"""GPU parity of the filter.h drop-in entry points (host pointers, in place, caller-owned
state) against the committed known-answer fixtures, and -- when oracle/_ref is present --
against the compiled reference directly, INCLUDING the state structs the calls leave behind.

The exact kernels visit taps in the reference's order with separately rounded multiplies
and adds, so the bar here is BIT-EXACT output, not a tolerance."""
import ctypes as C

import numpy as np
import pytest

from oracle import quisk_oracle as O
from oracle import ref_ctypes as R
from tests.golden.make_golden import FILTER_CASES, kat_input
from tests.util import SPLITS, golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from quisk_b200 import lib as L
    return C.CDLL(L.LIB_PATH) if L.require_device() else None


@pytest.fixture(scope="module")
def tabs():
    return golden("quisk_tables.npz")


@pytest.mark.parametrize("case", FILTER_CASES, ids=[c[0] for c in FILTER_CASES])
def test_filter_kat_bit_exact(case, lib, tabs):
    name, fn, seed, real, tab, args, tune = case
    kat = golden("filter_kat.npz")
    x = kat_input(seed, real)
    y, counts = R.FilterRunner(lib).run(fn, x, SPLITS, tabs[tab] if tab else None, args, tune)
    assert counts == kat[name + "/counts"].tolist()
    assert O.rel_rms(y, kat[name + "/y"]) < 1e-14
    assert np.array_equal(y, kat[name + "/y"]), "not bit-identical to the reference: max |d| = %g" % np.max(np.abs(y - kat[name + "/y"]))


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", FILTER_CASES, ids=[c[0] for c in FILTER_CASES])
def test_state_struct_matches_reference(case, lib, tabs):
    """After the same call sequence the caller-owned struct (ring contents, write offset,
    decim_index / toggle, shift registers) must be what filter.c leaves behind."""
    name, fn, seed, real, tab, args, tune = case
    x = kat_input(seed, real)
    ours = R.FilterRunner(lib); ours.run(fn, x, SPLITS, tabs[tab] if tab else None, args, tune)
    ref = R.FilterRunner(R.load("libquisk_filter_ref.so")); ref.run(fn, x, SPLITS, tabs[tab] if tab else None, args, tune)
    a, b = ours.state, ref.state
    if isinstance(a, (R.cHB45Filter, R.dHB45Filter)):
        assert a.toggle == b.toggle
        assert list(a.samples) == list(b.samples)
        assert list(a.center) == list(b.center)
    else:
        assert a.nTaps == b.nTaps and a.decim_index == b.decim_index
        es = 16 if isinstance(a, R.cFilter) else 8
        ra = a.cSamples if isinstance(a, R.cFilter) else a.dSamples
        rb = b.cSamples if isinstance(b, R.cFilter) else b.dSamples
        pa = a.ptcSamp if isinstance(a, R.cFilter) else a.ptdSamp
        pb = b.ptcSamp if isinstance(b, R.cFilter) else b.ptdSamp
        assert (pa - ra) == (pb - rb)
        n = a.nTaps * es // 8
        va = np.ctypeslib.as_array(C.cast(ra, C.POINTER(C.c_double)), (n,))
        vb = np.ctypeslib.as_array(C.cast(rb, C.POINTER(C.c_double)), (n,))
        assert np.array_equal(va, vb)


def test_interleaved_with_reference_state(lib, tabs):
    """A struct can be handed back and forth between libquisk_cuda and the reference."""
    if not R.have_ref():
        pytest.skip("oracle/_ref not built")
    ref = R.bind_filter_api(R.load("libquisk_filter_ref.so"))
    R.bind_filter_api(lib)
    h = np.ascontiguousarray(tabs["quiskFilt144D3Coefs"])
    x = kat_input(12, False, 6000)
    st_mix, st_ref = R.cFilter(), R.cFilter()
    ref.quisk_filt_cInit(C.byref(st_mix), h.ctypes.data_as(R.c_double_p), len(h))
    ref.quisk_filt_cInit(C.byref(st_ref), h.ctypes.data_as(R.c_double_p), len(h))
    out_mix, out_ref, pos = [], [], 0
    for i, n in enumerate([1000, 999, 1001, 7, 2993]):
        blk = x[pos:pos + n]; pos += n
        b1 = np.zeros(66000, dtype=np.complex128); b1[:n] = blk
        b2 = b1.copy()
        f = lib if i % 2 == 0 else ref
        k1 = f.quisk_cDecimate(b1.ctypes.data, n, C.byref(st_mix), 3)
        k2 = ref.quisk_cDecimate(b2.ctypes.data, n, C.byref(st_ref), 3)
        assert k1 == k2
        out_mix.append(b1[:k1].copy()); out_ref.append(b2[:k2].copy())
    assert np.array_equal(np.concatenate(out_mix), np.concatenate(out_ref))


def test_hb45_zero_struct_is_fresh_and_odd_lengths(lib):
    """An all-zero quisk_cHB45Filter is a valid fresh filter (quisk.c:1702-1706); nOut = (count + toggle) // 2."""
    R.bind_filter_api(lib)
    st = R.cHB45Filter()
    x = kat_input(13, False, 64)
    tog, pos = 0, 0
    for n in [1, 1, 3, 5, 2, 7, 45]:
        buf = np.zeros(128, dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
        k = lib.quisk_cDecim2HB45(buf.ctypes.data, n, C.byref(st))
        assert k == (n + tog) // 2
        tog = (tog + n) & 1
        assert st.toggle == tog


def test_impulse_returns_tap_table(lib, tabs):
    """decim=1 on an impulse returns the tap table verbatim, in order."""
    R.bind_filter_api(lib)
    h = np.ascontiguousarray(tabs["quiskFilt48dec24Coefs"])
    x = np.zeros(200, dtype=np.complex128); x[0] = 1.0
    y, counts = R.FilterRunner(lib).run("quisk_cDecimate", x, [200], h, (1,))
    assert counts == [200]
    assert np.array_equal(y[:98].real, h) and not y[98:].any()


def test_interp_output_clip(lib, tabs):
    x = kat_input(7, False, 30000)
    y, c = R.FilterRunner(lib).run("quisk_cInterpolate", x, [30000], tabs["quiskAudio24p4Coefs"], (2,))
    assert c == [52800]
    yo = O.FirInterp(tabs["quiskAudio24p4Coefs"], 2)(x)
    assert O.rel_rms(y, yo) < 1e-13
    y, c = R.FilterRunner(lib).run("quisk_cInterp2HB45", x, [30000])
    assert c == [52802]
    assert O.rel_rms(y, O.HB45Interp(np.complex128)(x)) < 1e-13


def test_per_sample_entry_points(lib, tabs):
    """quisk_dD_out / quisk_dC_out (filter.c:326-345, 83-104)."""
    R.bind_filter_api(lib)
    lib.quisk_dC_out.argtypes = [C.c_double, C.POINTER(R.dFilter)]

    class CD(C.Structure):
        _fields_ = [("re", C.c_double), ("im", C.c_double)]
    lib.quisk_dC_out.restype = CD
    h = np.ascontiguousarray(tabs["quiskAudio24p6Coefs"])
    st = R.dFilter()
    lib.quisk_filt_dInit(C.byref(st), h.ctypes.data_as(R.c_double_p), len(h))
    x = kat_input(14, True, 80)
    got = np.array([lib.quisk_dD_out(float(v), C.byref(st)) for v in x])
    exp = O.FirDecim(h, 1, np.float64)(x)
    assert O.rel_rms(got, exp) < 1e-13
    st2 = R.dFilter()
    lib.quisk_filt_dInit(C.byref(st2), h.ctypes.data_as(R.c_double_p), len(h))
    lib.quisk_filt_tune(C.cast(C.byref(st2), C.c_void_p), 0.1, 1)
    got = []
    for v in x:
        r = lib.quisk_dC_out(float(v), C.byref(st2)); got.append(complex(r.re, r.im))
    D = (len(h) - 1.0) / 2.0
    hc = np.exp(2j * np.pi * 0.1 * (np.arange(len(h)) - D)) * h
    exp = O.FirDecim(hc, 1)(x.astype(np.complex128))
    assert O.rel_rms(np.array(got), exp) < 1e-13


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built")
def test_shared_struct_stale_decim_index(lib, tabs):
    """A struct reused across entry points and ratios, as quisk.c does with filtDecim5S (cDecimate(5) then
    cInterpDecim(4, 5), quisk.c:1825,1837): an index left at or above the new `decim` emits on the first sample
    (filter.c:213), quisk_cFilter restarts it, quisk_dFilter / quisk_dD_out never touch it (filter.c:326-370)."""
    ref = R.bind_filter_api(R.load("libquisk_filter_ref.so"))
    R.bind_filter_api(lib)
    h = np.ascontiguousarray(tabs["quiskFilt240D5CoefsSharp"])
    x = kat_input(21, False, 400)
    res = []
    for L_ in (ref, lib):
        st = R.cFilter()
        L_.quisk_filt_cInit(C.byref(st), h.ctypes.data_as(R.c_double_p), len(h))
        outs = []
        for fn, n, args in [("quisk_cDecimate", 103, (5,)), ("quisk_cDecimate", 50, (2,)), ("quisk_cInterpDecim", 57, (4, 5)),
                            ("quisk_cDecimate", 33, (3,)), ("quisk_cFilter", 20, ()), ("quisk_cDecimate", 41, (5,))]:
            if fn == "quisk_cDecimate" and args == (2,):
                st.decim_index = 4                  # stale: >= the new ratio
            buf = np.zeros(1024, dtype=np.complex128); buf[:n] = x[:n]
            k = getattr(L_, fn)(buf.ctypes.data, n, C.byref(st), *args)
            outs.append((k, st.decim_index, buf[:k].copy()))
        res.append(outs)
    for (ka, ia, ya), (kb, ib, yb) in zip(*res):
        assert ka == kb and ia == ib and np.array_equal(ya, yb)
    hd = np.ascontiguousarray(tabs["quiskAudio24p6Coefs"])
    xr = kat_input(22, True, 64)
    res = []
    for L_ in (ref, lib):
        st = R.dFilter()
        L_.quisk_filt_dInit(C.byref(st), hd.ctypes.data_as(R.c_double_p), len(hd))
        st.decim_index = 3
        buf = xr.copy()
        k = L_.quisk_dFilter(buf.ctypes.data, 64, C.byref(st))
        v = L_.quisk_dD_out(1.5, C.byref(st))
        res.append((k, st.decim_index, buf[:k].copy(), v))
    assert res[0][0] == res[1][0] == 64 and res[0][1] == res[1][1] == 3
    assert np.array_equal(res[0][2], res[1][2]) and res[0][3] == res[1][3]
