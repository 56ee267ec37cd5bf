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


NOTCH_CASES = [
    # (notches [(centre, width, active)], tunefreq, shift, flow, fhigh): RF coordinates, pass band = tune + [flow, fhigh]
    ([(7001000.0, 200.0, 1)], 7000000.0, 0.0, 150.0, 2850.0),                                  # widened to the minimum width
    ([(7001000.0, 900.0, 1), (7002500.0, 700.0, 1), (7000100.0, 300.0, 1)], 7000000.0, 0.0, 150.0, 2850.0),   # inner, upper edge, lower edge
    ([(7001000.0, 900.0, 0), (7001500.0, 5000.0, 1)], 7000000.0, 0.0, 150.0, 2850.0),          # inactive one + one that swallows the band
    ([(13998700.0, 600.0, 1), (13998900.0, 600.0, 1)], 14000000.0, 250.0, -2850.0, -150.0),     # overlapping notches, shift, LSB band
    ([], 7000000.0, 0.0, 150.0, 2850.0),
]


@pytest.mark.parametrize("case", NOTCH_CASES, ids=[str(i) for i in range(len(NOTCH_CASES))])
def test_nbp_notched_impulse_bit_identical(libs, case):
    """quisk_cuda_nbp_impulse against the reference's own make_nbp + fir_mbandpass (nbp.c:64-179), called the way
    calc_nbp_impulse calls them (nbp.c:214-239)."""
    lib, ref = libs
    notches, tune, shift, flow, fhigh = case
    nc, rate, wintype, size = 2048, 48000.0, 0, 256
    scale = 1.0 / (2 * size)
    nn = len(notches)
    fc = np.array([n[0] for n in notches] + [0.0]); fw = np.array([n[1] for n in notches] + [0.0])
    act = np.array([n[2] for n in notches] + [0], dtype=np.int32)
    nlow = fc - 0.5 * fw; nhigh = fc + 0.5 * fw
    DP, IP = C.POINTER(D), C.POINTER(C.c_int)
    ref.make_nbp.argtypes = [C.c_int, IP, DP, DP, DP, DP, D, C.c_int, D, D, DP, DP, IP]
    ref.fir_mbandpass.restype = DP
    ref.fir_mbandpass.argtypes = [C.c_int, C.c_int, DP, DP, D, D, C.c_int]
    bplow = np.zeros(1025); bphigh = np.zeros(1025); hav = C.c_int(0)
    minwidth = 1600.0 / (nc // 256) * (rate / 48000)
    offset = tune + shift
    p = lambda a, t: a.ctypes.data_as(t)
    nbp = ref.make_nbp(nn, p(act, IP), p(fc, DP), p(fw, DP), p(nlow, DP), p(nhigh, DP), minwidth, 1, flow + offset, fhigh + offset,
                       p(bplow, DP), p(bphigh, DP), C.byref(hav))
    bplow[:nbp] -= offset; bphigh[:nbp] -= offset
    r = np.ctypeslib.as_array(ref.fir_mbandpass(nc, nbp, p(bplow, DP), p(bphigh, DP), rate, scale, wintype), (2 * nc,)).copy()
    out = np.zeros(2 * nc); numpb = C.c_int(-1); hav2 = C.c_int(-1)
    lib.quisk_cuda_nbp_impulse.argtypes = [C.c_int, D, D, D, C.c_int, D, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, D, D, C.c_int, C.c_int,
                                           C.c_void_p, IP, IP]
    assert lib.quisk_cuda_nbp_impulse(nc, flow, fhigh, rate, wintype, scale, nn, fc.ctypes.data, fw.ctypes.data, act.ctypes.data,
                                      tune, shift, 1, 1025, out.ctypes.data, C.byref(numpb), C.byref(hav2)) == 0
    assert numpb.value == nbp and hav2.value == hav.value
    assert np.array_equal(out, r)
    if nn and any(n[2] for n in notches):
        plain = np.zeros(2 * nc)
        lib.quisk_cuda_fir_bandpass(nc, flow, fhigh, rate, wintype, 1, scale, plain.ctypes.data)
        assert not np.array_equal(out, plain)                      # the notches really cut something out


@pytest.mark.parametrize("args", [(2048, 150., 2850., 48000., 0), (256, -4150., -150., 48000., 1), (1024, -8000., 8000., 192000., 0)])
@pytest.mark.parametrize("polarity", [0, 1])
def test_mp_imp_matches_reference(libs, args, polarity):
    """Minimum-phase impulse (fir.c:317-368).  The reference's three transforms go through FFTW (here: the oracle's
    FFT shim), ours through a radix-2 host FFT.  The method takes log|H| of stop-band bins that sit at the rounding
    floor of the first transform, so it amplifies FFT rounding: two correct FFTs give taps that agree to ~1e-6 of
    the rms tap, not to 1e-13 -- that is the conditioning of the reference's algorithm, the tolerance says so."""
    lib, ref = libs
    nc, flow, fhigh, rate, wintype = args
    imp = np.zeros(2 * nc)
    assert lib.quisk_cuda_fir_bandpass(nc, flow, fhigh, rate, wintype, 1, 1.0 / 512, imp.ctypes.data) == 0
    ref.mp_imp.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    ref.mp_imp.restype = None
    r = np.zeros(2 * nc)
    ref.mp_imp(nc, imp.ctypes.data, r.ctypes.data, 16, polarity)
    out = np.zeros(2 * nc)
    lib.quisk_cuda_mp_imp.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    assert lib.quisk_cuda_mp_imp(nc, imp.ctypes.data, out.ctypes.data, 16, polarity) == 0
    assert np.sqrt(np.mean((out - r) ** 2)) / np.sqrt(np.mean(r ** 2)) < 1e-5
    # it is minimum phase: same magnitude response, energy packed at the front (polarity 0) / back (polarity 1)
    H0 = np.abs(np.fft.fft(imp.view(np.complex128))); H1 = np.abs(np.fft.fft(out.view(np.complex128)))
    assert np.max(np.abs(H0 - H1)) / H0.max() < 1e-3
    e = np.abs(out.view(np.complex128)) ** 2
    centroid = (np.arange(nc) * e).sum() / e.sum() / nc         # the linear-phase input sits at 0.5
    assert (centroid < 0.3) if polarity == 0 else (centroid > 0.7)
    assert lib.quisk_cuda_mp_imp(1000, imp.ctypes.data, out.ctypes.data, 16, 0) != 0         # 16000 is not a power of two
