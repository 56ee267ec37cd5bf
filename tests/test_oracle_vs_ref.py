"""Pins the NumPy oracle (oracle/quisk_oracle.py) against the reference's own C
code compiled into oracle/_ref (filter.c verbatim, quisk.c RX functions through
the wrapper TU).  Skipped when oracle/_ref has not been built (it needs
/root/reference); the golden-fixture tests cover that case."""
import ctypes as C

import numpy as np
import pytest

from oracle import quisk_oracle as O
from oracle import ref_ctypes as R

pytestmark = pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built")

SPLITS = [1, 2, 3, 7, 255, 256, 257, 1000, 1, 2218]
TOL = 1e-13


@pytest.fixture(scope="module")
def flib():
    return R.load("libquisk_filter_ref.so")


@pytest.fixture(scope="module")
def tabs(flib):
    return R.all_tables(flib)


def _x(n, seed=0, real=False):
    x = O.synth_iq(n, seed, 1.0)
    return np.ascontiguousarray(x.real) if real else x


def _run_oracle(stage, x, splits):
    outs, counts, pos = [], [], 0
    for n in splits:
        y = stage(x[pos:pos + n]); pos += n
        outs.append(y); counts.append(len(y))
    return np.concatenate(outs), counts


def test_hb45_decim(flib):
    x = _x(sum(SPLITS))
    yr, cr = R.FilterRunner(flib).run("quisk_cDecim2HB45", x, SPLITS)
    yo, co = _run_oracle(O.HB45Decim(), x, SPLITS)
    assert cr == co
    assert O.rel_rms(yo, yr) < TOL


@pytest.mark.parametrize("name,decim", [("quiskFilt48dec24Coefs", 2), ("quiskFilt144D3Coefs", 3),
                                         ("quiskFilt240D5CoefsSharp", 5), ("quiskFilt53D1Coefs", 1)])
def test_cdecimate(flib, tabs, name, decim):
    x = _x(sum(SPLITS), 1)
    yr, cr = R.FilterRunner(flib).run("quisk_cDecimate", x, SPLITS, tabs[name], (decim,))
    yo, co = _run_oracle(O.FirDecim(tabs[name], decim), x, SPLITS)
    assert cr == co
    assert O.rel_rms(yo, yr) < TOL


def test_ccdecimate_tuned(flib, tabs):
    x = _x(sum(SPLITS), 2)
    h = tabs["quiskFilt48dec24Coefs"]
    yr, cr = R.FilterRunner(flib).run("quisk_cCDecimate", x, SPLITS, h, (2,), tune=(0.1, 1))
    D = (len(h) - 1.0) / 2.0
    hc = np.exp(2j * np.pi * 0.1 * (np.arange(len(h)) - D)) * h      # filter.c:72-80
    yo, co = _run_oracle(O.FirDecim(hc, 2), x, SPLITS)
    assert cr == co
    assert O.rel_rms(yo, yr) < TOL


def test_ddecimate_dfilter(flib, tabs):
    x = _x(sum(SPLITS), 3, real=True)
    yr, cr = R.FilterRunner(flib).run("quisk_dDecimate", x, SPLITS, tabs["quiskLpFilt48Coefs"], (4,))
    yo, co = _run_oracle(O.FirDecim(tabs["quiskLpFilt48Coefs"], 4, np.float64), x, SPLITS)
    assert cr == co and O.rel_rms(yo, yr) < TOL
    yr, cr = R.FilterRunner(flib).run("quisk_dFilter", x, SPLITS, tabs["quiskAudio24p6Coefs"])
    yo, co = _run_oracle(O.FirDecim(tabs["quiskAudio24p6Coefs"], 1, np.float64), x, SPLITS)
    assert cr == co and O.rel_rms(yo, yr) < TOL


@pytest.mark.parametrize("name,L", [("quiskAudio24p4Coefs", 2), ("quiskFilt300D5Coefs", 6), ("quiskAudio24p3Coefs", 3)])
def test_interpolate(flib, tabs, name, L):
    x = _x(sum(SPLITS), 4)
    yr, cr = R.FilterRunner(flib).run("quisk_cInterpolate", x, SPLITS, tabs[name], (L,))
    yo, co = _run_oracle(O.FirInterp(tabs[name], L), x, SPLITS)
    assert cr == co and O.rel_rms(yo, yr) < TOL
    xr = np.ascontiguousarray(x.real)
    yr, cr = R.FilterRunner(flib).run("quisk_dInterpolate", xr, SPLITS, tabs[name], (L,))
    yo, co = _run_oracle(O.FirInterp(tabs[name], L, np.float64), xr, SPLITS)
    assert cr == co and O.rel_rms(yo, yr) < TOL


@pytest.mark.parametrize("name,L,M", [("quiskFilt300D5Coefs", 6, 5), ("quiskFilt240D5CoefsSharp", 4, 5),
                                       ("quiskFilt144D3Coefs", 2, 3), ("quiskFilt300D5Coefs", 5, 2)])
def test_interp_decim(flib, tabs, name, L, M):
    x = _x(sum(SPLITS), 5)
    yr, cr = R.FilterRunner(flib).run("quisk_cInterpDecim", x, SPLITS, tabs[name], (L, M))
    yo, co = _run_oracle(O.FirInterpDecim(tabs[name], L, M), x, SPLITS)
    assert cr == co
    assert O.rel_rms(yo, yr) < TOL


def test_hb45_interp(flib):
    x = _x(sum(SPLITS), 6)
    yr, cr = R.FilterRunner(flib).run("quisk_cInterp2HB45", x, SPLITS)
    yo, co = _run_oracle(O.HB45Interp(np.complex128), x, SPLITS)
    assert cr == co and O.rel_rms(yo, yr) < TOL
    xr = np.ascontiguousarray(x.real)
    yr, cr = R.FilterRunner(flib).run("quisk_dInterp2HB45", xr, SPLITS)
    yo, co = _run_oracle(O.HB45Interp(np.float64), xr, SPLITS)
    assert cr == co and O.rel_rms(yo, yr) < TOL


def test_interp_output_clip(flib, tabs):
    """The interpolators silently stop at 52 800 (+2 for the half band) outputs."""
    x = _x(30000, 7)
    yr, cr = R.FilterRunner(flib).run("quisk_cInterpolate", x, [30000], tabs["quiskAudio24p4Coefs"], (2,))
    yo, co = _run_oracle(O.FirInterp(tabs["quiskAudio24p4Coefs"], 2), x, [30000])
    assert cr == co == [52800]
    yr, cr = R.FilterRunner(flib).run("quisk_cInterp2HB45", x, [30000])
    yo, co = _run_oracle(O.HB45Interp(np.complex128), x, [30000])
    assert cr == co == [52802]
    assert O.rel_rms(yo, yr) < TOL


# ---- quisk.c RX functions --------------------------------------------------

def _rxlib():
    lib = R.load("libquisk_rx_ref.so", private_copy=True)
    lib.ref_process_decimate.restype = C.c_int
    lib.ref_process_demodulate.restype = C.c_int
    return lib


def _set_filters(lib, fi, fq, bw=2800):
    fi = np.ascontiguousarray(fi); fq = np.ascontiguousarray(fq)
    lib.ref_set_filters(fi.ctypes.data_as(C.c_void_p), fq.ctypes.data_as(C.c_void_p), len(fi), bw, 0)


def test_rx_filter_tap_order(tabs):
    """F2: newest x filt[0], then oldest x filt[1] ... (quisk.c:1203-1215, 1240-1255)."""
    rng = np.random.default_rng(5)
    fi = rng.standard_normal(164); fq = rng.standard_normal(164)
    x = _x(3000, 8)
    lib = _rxlib(); _set_filters(lib, fi, fq)
    y = x.copy(); lib.ref_cRxFilterOut(y.ctypes.data_as(C.c_void_p), len(y), 0, 0)
    f = O.RxFilterC(fi, fq)
    yo = np.concatenate([f(x[:1000]), f(x[1000:1001]), f(x[1001:])])
    assert O.rel_rms(yo, y) < TOL
    y = x.copy(); lib.ref_dRxFilterOut(y.ctypes.data_as(C.c_void_p), len(y), 0, 0)
    assert O.rel_rms(O.RxFilterD(fi)(x), y) < TOL


@pytest.mark.parametrize("rate", [1536000, 192000, 96000, 48000, 240000, 250000, 960000, 1200000, 111111, 185185])
def test_process_decimate(tabs, rate):
    lib = _rxlib()
    lib.ref_set_sample_rate(rate); lib.ref_init_chain()
    chain = O.ProcessDecimate(rate, tabs)
    x = _x(40000, 9)
    splits = [1, 2, 3, 7, 255, 256, 257, 1000, 4093, 15360, 18766]
    outs_r, outs_o, pos = [], [], 0
    for n in splits:
        blk = x[pos:pos + n]; pos += n
        buf = np.zeros(66000, dtype=np.complex128); buf[:n] = blk
        nr = lib.ref_process_decimate(buf.ctypes.data_as(C.c_void_p), n, 0, 3)
        yo = chain(blk)
        assert nr == len(yo)
        outs_r.append(buf[:nr].copy()); outs_o.append(yo)
    assert lib.ref_decim_srate() == chain.decim_srate
    assert O.rel_rms(np.concatenate(outs_o), np.concatenate(outs_r)) < TOL


@pytest.mark.parametrize("mode", ["USB", "LSB", "CWU", "CWL", "AM", "FM"])
def test_process_demodulate(tabs, mode):
    lib = _rxlib()
    lib.ref_set_sample_rate(48000); lib.ref_init_chain()
    rng = np.random.default_rng(3)
    ntaps = {"USB": 164, "LSB": 164, "CWU": 390, "CWL": 390, "AM": 77, "FM": 55}[mode]
    fi = rng.standard_normal(ntaps) / ntaps; fq = rng.standard_normal(ntaps) / ntaps
    _set_filters(lib, fi, fq)
    chain = O.ProcessDemodulate(mode, fi, fq, tabs)
    x = _x(12000, 10)
    splits = [1, 2, 3, 7, 255, 256, 257, 1000, 4093, 6126]
    outs_r, outs_o, pos = [], [], 0
    for n in splits:
        blk = x[pos:pos + n]; pos += n
        buf = np.zeros(66000, dtype=np.complex128); buf[:n] = blk
        dbuf = np.zeros(66000 * 2)
        nr = lib.ref_process_demodulate(buf.ctypes.data_as(C.c_void_p), dbuf.ctypes.data_as(C.c_void_p),
                                        n, 0, 0, R.MODES[mode])
        yo = chain(blk)
        assert nr == len(yo), (mode, n)
        outs_r.append(dbuf[:nr].copy()); outs_o.append(yo)
    assert O.rel_rms(np.concatenate(outs_o), np.concatenate(outs_r)) < 1e-11 if mode == "FM" else TOL


def test_tune_recurrence():
    lib = _rxlib()
    x = _x(5000, 11)
    y = x.copy()
    vec = np.array([1.0 + 0j])
    lib.ref_tune.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
    lib.ref_tune(y.ctypes.data, 3000, 12345.0, 1536000, vec.ctypes.data)
    lib.ref_tune(y[3000:].ctypes.data, 2000, 12345.0, 1536000, vec.ctypes.data)
    nco = O.TuneNCO(12345.0, 1536000)
    yo = np.concatenate([nco(x[:3000]), nco(x[3000:])])
    assert np.array_equal(yo, y)            # literal restatement: bit exact
    assert nco.v == vec[0]


def test_plan_decimation():
    lib = _rxlib()
    for rate in [48000, 96000, 192000, 240000, 250000, 384000, 960000, 1200000, 1536000, 2000000, 3072000, 98304000 // 64]:
        lib.ref_set_sample_rate(rate)
        p2, p3, p5 = C.c_int(), C.c_int(), C.c_int()
        best = lib.ref_plan_decimation(C.byref(p2), C.byref(p3), C.byref(p5))
        assert (best, p2.value, p3.value, p5.value) == O.plan_decimation(rate)
