"""GPU parity of the WDSP RXA stages (include/quisk_cuda_wdsp.h) against known-answer fixtures generated
from the compiled reference (tests/golden/make_golden_wdsp.py -> wdsp_kat.npz).  Every case runs several
identical channels through the batched kernels.  Tolerances: fircore-based stages 1e-12 relative RMS (the
north star's FP64 bound; our FFT is not FFTW's), recurrent stages 1e-12 as well, resampler bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.golden.make_golden_wdsp import FIRCORE_CASES, RESAMPLE_CASES, FM_BLOCKS, sig, fm_sig, am_sig
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
    return golden("wdsp_kat.npz")


def _bandpass(lib, N, fl, fh, rate, wintype, rtype, scale):
    out = np.zeros(N * (2 if rtype else 1))
    assert lib.quisk_cuda_fir_bandpass(N, fl, fh, rate, wintype, rtype, scale, out.ctypes.data) == 0
    return out


def _dev(torch, x):
    return torch.from_numpy(np.ascontiguousarray(np.stack([x] * NCH))).cuda()


@pytest.mark.parametrize("size,nc,rate", FIRCORE_CASES)
def test_fircore(size, nc, rate, torch, lib, kat):
    imp = _bandpass(lib, nc, 150.0, 2850.0, rate, 0, 1, 1.0 / (2 * size))
    imp2 = _bandpass(lib, nc, -2850.0, -150.0, rate, 0, 1, 1.0 / (2 * size))
    f = lib.quisk_cuda_fircore_create(NCH, size, nc, 0, imp.ctypes.data)
    assert f, lib.quisk_cuda_last_error()
    x = sig(size * 8, 100 + size, rate)
    d = _dev(torch, x)
    o = torch.zeros_like(d)
    for b in range(8):
        if b == 5:
            assert lib.quisk_cuda_fircore_set_impulse(f, imp2.ctypes.data, 1) == 0
        blk = d[:, b * size:(b + 1) * size]
        ob = o[:, b * size:(b + 1) * size]
        assert lib.quisk_cuda_fircore_run(f, blk.data_ptr(), d.stride(0), ob.data_ptr(), o.stride(0), None) == 0
    torch.cuda.synchronize()
    y = o.cpu().numpy()
    ref = kat["fircore_%d_%d/y" % (size, nc)]
    for c in range(NCH):
        assert O.rel_rms(y[c], ref) < 1e-12
    # first five blocks are a plain causal convolution scaled by 2*size
    h = imp.view(np.complex128)
    lin = np.convolve(x[:5 * size], h)[:5 * size] * (2 * size)
    assert O.rel_rms(y[0][:5 * size], lin) < 1e-12
    lib.quisk_cuda_fircore_destroy(f)


def test_fircore_minimum_phase(torch, lib, kat):
    """mp = 1: the masks come from mp_imp of the impulse (firmin.c:327-328).  mp_imp amplifies FFT rounding (see
    tests/test_wdsp_design.py), so two correct implementations agree to ~1e-6 here, and the filter now responds
    at once instead of after nc/2 samples."""
    size, nc, rate = 256, 1024, 48000.0
    imp = _bandpass(lib, nc, 150.0, 2850.0, rate, 0, 1, 1.0 / (2 * size))
    f = lib.quisk_cuda_fircore_create(NCH, size, nc, 1, imp.ctypes.data)
    assert f, lib.quisk_cuda_last_error()
    x = sig(size * 8, 150, rate)
    d = _dev(torch, x)
    o = torch.zeros_like(d)
    for b in range(8):
        blk = d[:, b * size:(b + 1) * size]
        ob = o[:, b * size:(b + 1) * size]
        assert lib.quisk_cuda_fircore_run(f, blk.data_ptr(), d.stride(0), ob.data_ptr(), o.stride(0), None) == 0
    torch.cuda.synchronize()
    y = o.cpu().numpy()
    ref = kat["fircore_mp_256_1024/y"]
    errs = [O.rel_rms(y[c], ref) for c in range(NCH)]
    print("fircore mp", errs, "reference's own sensitivity to a one-ulp change of the impulse", kat["fircore_mp_256_1024/cond"])
    assert max(errs) < _cond_bound(kat, "fircore_mp_256_1024", k=100.0)
    lin = lib.quisk_cuda_fircore_create(NCH, size, nc, 0, imp.ctypes.data)
    o2 = torch.zeros_like(d)
    for b in range(8):
        assert lib.quisk_cuda_fircore_run(lin, d[:, b * size:(b + 1) * size].data_ptr(), d.stride(0), o2[:, b * size:(b + 1) * size].data_ptr(), o2.stride(0), None) == 0
    torch.cuda.synchronize()
    y2 = o2.cpu().numpy()
    e_mp = np.sum(np.abs(y[0][:size]) ** 2); e_lin = np.sum(np.abs(y2[0][:size]) ** 2)
    assert e_mp > 10 * e_lin                         # minimum phase: output energy arrives in the first block
    # setMp_fircore on a fresh linear-phase object gives the object created with mp = 1
    sw = lib.quisk_cuda_fircore_create(NCH, size, nc, 0, imp.ctypes.data)
    assert lib.quisk_cuda_fircore_set_mp(sw, 1) == 0
    o3 = torch.zeros_like(d)
    for b in range(8):
        assert lib.quisk_cuda_fircore_run(sw, d[:, b * size:(b + 1) * size].data_ptr(), d.stride(0), o3[:, b * size:(b + 1) * size].data_ptr(), o3.stride(0), None) == 0
    torch.cuda.synchronize()
    assert np.array_equal(o3.cpu().numpy(), y)
    lib.quisk_cuda_fircore_destroy(f); lib.quisk_cuda_fircore_destroy(lin); lib.quisk_cuda_fircore_destroy(sw)


@pytest.mark.parametrize("in_rate,out_rate,splits", RESAMPLE_CASES)
def test_resample(in_rate, out_rate, splits, torch, lib, kat):
    r = lib.quisk_cuda_resample_create(NCH, in_rate, out_rate, 0.0, 0, 1.0)
    assert r, lib.quisk_cuda_last_error()
    x = sig(sum(splits), 200, in_rate)
    d = _dev(torch, x)
    ys, counts, pos = [], [], 0
    for n in splits:
        blk = d[:, pos:pos + n].contiguous(); pos += n
        cap = lib.quisk_cuda_resample_count_out(r, n) + 4
        o = torch.zeros((NCH, cap), dtype=torch.complex128, device="cuda")
        k = C.c_int(0)
        assert lib.quisk_cuda_resample_run(r, blk.data_ptr(), n, n, o.data_ptr(), cap, C.byref(k), None) == 0
        torch.cuda.synchronize()
        ys.append(o[:, :k.value].cpu().numpy()); counts.append(k.value)
    y = np.concatenate(ys, axis=1)
    assert counts == kat["resample_%d_%d/counts" % (in_rate, out_rate)].tolist()
    for c in range(NCH):
        assert np.array_equal(y[c], kat["resample_%d_%d/y" % (in_rate, out_rate)])     # same order, same roundings
    lib.quisk_cuda_resample_destroy(r)


def _run_seq(torch, lib, st, x, n, nblk):
    d = _dev(torch, x)
    for b in range(nblk):
        blk = d[:, b * n:(b + 1) * n]
        assert lib.quisk_cuda_seq_run(st, blk.data_ptr(), d.stride(0), blk.data_ptr(), d.stride(0), n, None) == 0
    torch.cuda.synchronize()
    return d.cpu().numpy()


def test_shift(torch, lib, kat):
    hz = np.array([1234.5] * NCH)
    st = lib.quisk_cuda_shift_create(NCH, 48000, hz.ctypes.data)
    y = _run_seq(torch, lib, st, sig(3 * 1024, 300, 48000.0), 1024, 3)
    for c in range(NCH):
        assert O.rel_rms(y[c], kat["shift/y"]) < 1e-13
    lib.quisk_cuda_seq_destroy(st)


@pytest.mark.parametrize("mode", [3, 1, 4])
def test_wcpagc(mode, torch, lib, kat):
    n = 1024; rate = 192000 if mode == 3 else 48000
    x = sig(8 * n, 400 + mode, float(rate))
    x[2 * n:3 * n] *= 3.0; x[4 * n:6 * n] *= 0.05; x[6 * n:] *= 2.0
    st = lib.quisk_cuda_wcpagc_create(NCH, rate, mode)
    assert st
    y = _run_seq(torch, lib, st, x, n, 8)
    for c in range(NCH):
        assert O.rel_rms(y[c], kat["wcpagc_mode%d/y" % mode]) < 1e-12
    lib.quisk_cuda_seq_destroy(st)


@pytest.mark.parametrize("mode,sb", [(0, 0), (1, 0), (1, 1), (1, 2)])
def test_amd(mode, sb, torch, lib, kat):
    st = lib.quisk_cuda_amd_create(NCH, 48000, mode, 1, sb)
    y = _run_seq(torch, lib, st, am_sig(4 * 512, 500, 48000.0), 512, 4)
    errs = [O.rel_rms(y[c], kat["amd_%d_%d/y" % (mode, sb)]) for c in range(NCH)]
    print("amd", mode, sb, errs)
    assert max(errs) < _cond_bound(kat, "amd_%d_%d" % (mode, sb))
    lib.quisk_cuda_seq_destroy(st)


def _run_rxa(torch, lib, rxa, x, in_size, nblocks, use_fexchange=False):
    out_size = lib.quisk_cuda_rxa_out_size(rxa)
    assert lib.quisk_cuda_rxa_in_size(rxa) == in_size
    if use_fexchange:
        ys = []
        for b in range(nblocks):
            hin = np.ascontiguousarray(np.stack([x[b * in_size:(b + 1) * in_size]] * NCH))
            hout = np.zeros((NCH, out_size), dtype=np.complex128)
            err = C.c_int(9)
            assert lib.quisk_cuda_rxa_fexchange0(rxa, hin.ctypes.data, hout.ctypes.data, C.byref(err)) == 0, lib.quisk_cuda_last_error()
            assert err.value == 0
            ys.append(hout)
        return np.concatenate(ys, axis=1)
    d = _dev(torch, x)
    o = torch.zeros((NCH, out_size * nblocks), dtype=torch.complex128, device="cuda")
    for b in range(nblocks):
        blk = d[:, b * in_size:(b + 1) * in_size]
        ob = o[:, b * out_size:(b + 1) * out_size]
        assert lib.quisk_cuda_rxa_xrxa(rxa, blk.data_ptr(), d.stride(0), ob.data_ptr(), o.stride(0), None) == 0, lib.quisk_cuda_last_error()
    torch.cuda.synchronize()
    return o.cpu().numpy()


def _first_sample_swallowed(x):
    """fexchange0's up-slew state machine zeroes the first non-zero input sample (iobuffs.c:110-128).  Only for the
    tests that drive xrxa directly: quisk_cuda_rxa_fexchange0 runs that state machine itself."""
    x = x.copy(); x[0] = 0.0
    return x


def test_rxa_usb_channel(torch, lib, kat):
    """SURVEY 8(d) C3 scaled down: nbp0 (nc 2048) + wcpAGC mode 3 + panel, through the fexchange0-shaped entry."""
    x = sig(256 * 24, 700, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
    rxa = lib.quisk_cuda_rxa_create(NCH, 256, 256, 48000, 48000, 48000)
    assert rxa, lib.quisk_cuda_last_error()
    assert lib.quisk_cuda_rxa_set_shift(rxa, 0, None) == 0
    assert lib.quisk_cuda_rxa_set_nc(rxa, 2048) == 0
    assert lib.quisk_cuda_rxa_set_mode(rxa, 1) == 0
    assert lib.quisk_cuda_rxa_set_passband(rxa, 150.0, 2850.0) == 0
    assert lib.quisk_cuda_rxa_set_agc_mode(rxa, 3) == 0
    y = _run_rxa(torch, lib, rxa, x, 256, 24, use_fexchange=True)
    ref = kat["rxa_usb/y"]
    assert not y[0][:512].any() and not ref[:512].any()            # two DSP buffers of latency
    print("rxa_usb", [O.rel_rms(y[c], ref) for c in range(NCH)])
    for c in range(NCH):
        assert O.rel_rms(y[c], ref) < _cond_bound(kat, "rxa_usb")
    av = np.zeros(NCH); pk = np.zeros(NCH); g = np.zeros(NCH)
    assert lib.quisk_cuda_rxa_get_meter(rxa, 1, av.ctypes.data, pk.ctypes.data, None) == 0
    assert np.all(av < 0) and np.all(av > -60) and np.all(pk >= av - 1e-9)
    assert lib.quisk_cuda_rxa_get_meter(rxa, 2, av.ctypes.data, pk.ctypes.data, g.ctypes.data) == 0
    assert np.all(np.isfinite(g))
    # sip1: the newest 1024 samples of midbuff after the last block, as RXAGetaSipF1 returns them (floats)
    sip = np.zeros((NCH, 2 * 1024), dtype=np.float32)
    assert lib.quisk_cuda_rxa_get_siphon(rxa, sip.ctypes.data, 1024, 1) == 0, lib.quisk_cuda_last_error()
    rs = kat["rxa_usb/sip"].astype(np.float64)
    for c in range(NCH):
        assert np.sqrt(np.mean((sip[c] - rs) ** 2)) / np.sqrt(np.mean(rs ** 2)) < 1e-6       # float32 output
    re = np.zeros((NCH, 1024), dtype=np.float32)
    assert lib.quisk_cuda_rxa_get_siphon(rxa, re.ctypes.data, 1024, 0) == 0
    assert np.array_equal(re[0], sip[0][0::2])
    lib.quisk_cuda_rxa_destroy(rxa)


def test_rxa_usb_channel_upslew(torch, lib, kat):
    """The channel opened the way Quisk opens it (quisk_wdsp.py:79-80: tdelayup 10 ms, tslewup 25 ms): upslew0's
    state machine (iobuffs.c:98-160) on a stream that starts with 100 zero samples -- zeros up to and including the
    first non-zero sample, 480 + 1 more zeros, a 1200 + 1 sample raised-cosine ramp, then pass-through."""
    x = sig(256 * 24, 700, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
    x[:100] = 0.0
    rxa = lib.quisk_cuda_rxa_create(NCH, 256, 256, 48000, 48000, 48000)
    assert rxa, lib.quisk_cuda_last_error()
    assert lib.quisk_cuda_rxa_set_slew(rxa, 0.010, 0.025) == 0
    assert lib.quisk_cuda_rxa_set_shift(rxa, 0, None) == 0
    assert lib.quisk_cuda_rxa_set_nc(rxa, 2048) == 0
    assert lib.quisk_cuda_rxa_set_mode(rxa, 1) == 0
    assert lib.quisk_cuda_rxa_set_passband(rxa, 150.0, 2850.0) == 0
    assert lib.quisk_cuda_rxa_set_agc_mode(rxa, 3) == 0
    y = _run_rxa(torch, lib, rxa, x, 256, 24, use_fexchange=True)
    ref = kat["rxa_usb_slew/y"]
    assert np.array_equal(np.nonzero(ref)[0][:1], np.nonzero(y[0])[0][:1])       # same first non-zero output sample
    for c in range(NCH):
        assert O.rel_rms(y[c], ref) < _cond_bound(kat, "rxa_usb")
    assert O.rel_rms(y[0], kat["rxa_usb/y"]) > 1e-3                               # and the ramp is really there
    lib.quisk_cuda_rxa_destroy(rxa)


def test_rxa_usb_channel_notches(torch, lib, kat):
    """nbp0 with the notch database running (RXANBPSetNotchesRun / AddNotch / SetTuneFrequency, nbp.c:359-513): two
    notches inside the pass band, one narrower than the minimum width and therefore auto-widened."""
    x = sig(256 * 24, 700, 48000.0, tones=((1000.0, 0.3), (1500.0, 0.2), (2200.0, 0.1)))
    rxa = lib.quisk_cuda_rxa_create(NCH, 256, 256, 48000, 48000, 48000)
    assert rxa, lib.quisk_cuda_last_error()
    assert lib.quisk_cuda_rxa_set_shift(rxa, 0, None) == 0
    assert lib.quisk_cuda_rxa_set_nc(rxa, 2048) == 0
    assert lib.quisk_cuda_rxa_set_mode(rxa, 1) == 0
    assert lib.quisk_cuda_rxa_set_passband(rxa, 150.0, 2850.0) == 0
    assert lib.quisk_cuda_rxa_set_agc_mode(rxa, 3) == 0
    assert lib.quisk_cuda_rxa_nbp_set_tune_frequency(rxa, 7000000.0) == 0
    assert lib.quisk_cuda_rxa_nbp_set_notches_run(rxa, 1) == 0
    assert lib.quisk_cuda_rxa_nbp_add_notch(rxa, 0, 7001500.0, 400.0, 1) == 0
    assert lib.quisk_cuda_rxa_nbp_add_notch(rxa, 1, 7002400.0, 100.0, 1) == 0
    assert lib.quisk_cuda_rxa_nbp_add_notch(rxa, 5, 7002000.0, 100.0, 1) == -1        # beyond the end of the list
    y = _run_rxa(torch, lib, rxa, x, 256, 24, use_fexchange=True)
    ref = kat["rxa_usb_notch/y"]
    for c in range(NCH):
        assert O.rel_rms(y[c], ref) < _cond_bound(kat, "rxa_usb")
    lib.quisk_cuda_rxa_destroy(rxa)


def test_rxa_default_channel_keeps_bp1(torch, lib, kat):
    """No SetRXAMode: bp1 is still running -- the quirk is part of the reference's behaviour."""
    x = _first_sample_swallowed(sig(256 * 16, 701, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2))))
    rxa = lib.quisk_cuda_rxa_create(NCH, 256, 256, 48000, 48000, 48000)
    lib.quisk_cuda_rxa_set_shift(rxa, 0, None)
    y = _run_rxa(torch, lib, rxa, x, 256, 16)
    ref = kat["rxa_default/y"][512:]
    for c in range(NCH):
        assert O.rel_rms(y[c][:len(ref)], ref) < _cond_bound(kat, "rxa_usb")
    lib.quisk_cuda_rxa_destroy(rxa)


def test_fmd_stage_composition(torch, lib, kat):
    """xfmd (fmd.c:144-188) = PLL -> de-emphasis fircore -> audio fircore -> CTCSS notch, on a signal that is
    present from the first sample (so the PLL never runs on rounding noise)."""
    n, rate = 256, 48000.0
    de = np.zeros(4096); au = np.zeros(4096)
    assert lib.quisk_cuda_fc_impulse(2048, 300.0, 3000.0, 20.0 * np.log10(3000.0 / 300.0), 0.0, 1, rate, 1.0 / (2.0 * n), 0, 0, de.ctypes.data) == 0
    au = _bandpass(lib, 2048, 0.8 * 300.0, 1.1 * 3000.0, rate, 0, 1, 0.5 / (2.0 * n))
    pll = lib.quisk_cuda_fmpll_create(NCH, 48000, 5000.0, -8000.0, 8000.0, 1.0, 20000.0, 0.02)
    pde = lib.quisk_cuda_fircore_create(NCH, n, 2048, 0, de.ctypes.data)
    paud = lib.quisk_cuda_fircore_create(NCH, n, 2048, 0, au.ctypes.data)
    sn = lib.quisk_cuda_snotch_create(NCH, 48000, 254.1, 0.0002)
    d = _dev(torch, fm_sig(12 * n, 600, rate))
    a = torch.zeros((NCH, n), dtype=torch.complex128, device="cuda")
    for b in range(12):
        blk = d[:, b * n:(b + 1) * n]
        assert lib.quisk_cuda_seq_run(pll, blk.data_ptr(), d.stride(0), a.data_ptr(), n, n, None) == 0
        assert lib.quisk_cuda_fircore_run(pde, a.data_ptr(), n, blk.data_ptr(), d.stride(0), None) == 0
        assert lib.quisk_cuda_fircore_run(paud, blk.data_ptr(), d.stride(0), blk.data_ptr(), d.stride(0), None) == 0
        assert lib.quisk_cuda_seq_run(sn, blk.data_ptr(), d.stride(0), blk.data_ptr(), d.stride(0), n, None) == 0
    torch.cuda.synchronize()
    y = d.cpu().numpy()
    errs = [O.rel_rms(y[c], kat["fmd/y"]) for c in range(NCH)]
    print("fmd stage", errs, "reference's own one-ulp sensitivity", kat["fmd/cond"])
    assert max(errs) < _cond_bound(kat, "fmd")


def test_rxa_fm_channel(torch, lib, kat):
    """384 kS/s in -> resample (1121 taps, /8) -> nbp0 -> fmd (PLL + 2 fircores + notch) -> panel.
    Compared on the last 16 of FM_BLOCKS blocks: the PLL's cold start runs on FFT rounding noise (see the generator)."""
    x = _first_sample_swallowed(fm_sig(2048 * FM_BLOCKS, 702, 384000.0))
    rxa = lib.quisk_cuda_rxa_create(NCH, 2048, 256, 384000, 48000, 48000)
    assert rxa, lib.quisk_cuda_last_error()
    lib.quisk_cuda_rxa_set_shift(rxa, 0, None)
    assert lib.quisk_cuda_rxa_set_mode(rxa, 5) == 0
    assert lib.quisk_cuda_rxa_set_passband(rxa, -8000.0, 8000.0) == 0
    y = _run_rxa(torch, lib, rxa, x, 2048, FM_BLOCKS)
    ref = kat["rxa_fm/y_tail"]                      # fexchange0 output blocks FM_BLOCKS-16 .. FM_BLOCKS-1
    ours = y[:, (FM_BLOCKS - 18) * 256:(FM_BLOCKS - 2) * 256]     # same blocks: the exchange delays by two
    errs = [O.rel_rms(ours[c], ref) for c in range(NCH)]
    print("rxa_fm tail", errs, "reference's own one-ulp sensitivity", kat["rxa_fm/cond"])
    assert max(errs) < _cond_bound(kat, "rxa_fm")
    lib.quisk_cuda_rxa_destroy(rxa)


def test_rxa_am_channel(torch, lib, kat):
    x = _first_sample_swallowed(am_sig(256 * 16, 703, 48000.0))
    rxa = lib.quisk_cuda_rxa_create(NCH, 256, 256, 48000, 48000, 48000)
    lib.quisk_cuda_rxa_set_shift(rxa, 0, None)
    assert lib.quisk_cuda_rxa_set_mode(rxa, 6) == 0
    assert lib.quisk_cuda_rxa_set_passband(rxa, -4000.0, 4000.0) == 0
    y = _run_rxa(torch, lib, rxa, x, 256, 16)
    ref = kat["rxa_am/y"][512:]
    errs = [O.rel_rms(y[c][:len(ref)], ref) for c in range(NCH)]
    print("rxa_am", errs)
    assert max(errs) < _cond_bound(kat, "rxa_am")
    lib.quisk_cuda_rxa_destroy(rxa)


# ---------------------------------------------------------------------------------------------------------------------
# Round 2: the full C3 geometry, meters, the general exchange (re-blocking, down-slew, flush, restart) and the
# reference's own entry points by channel number (wdsp_compat.cu).
# ---------------------------------------------------------------------------------------------------------------------

def _cond_bound(kat, key, k=20.0, floor=1e-12):
    """Tolerance from the reference's OWN sensitivity: `<key>/cond` is the relative RMS change of the compiled reference's
    output when every input component moves by one ulp (tests/golden/make_golden_wdsp.py).  An implementation with
    different but equally valid roundings (another FFT, another libm) cannot be expected closer than a small multiple of
    that; where the reference is better conditioned than 1e-12 / k the north star's 1e-12 stands."""
    return max(floor, k * float(kat[key + "/cond"][0]))


def _setup_usb(lib, rxa, nc=2048):
    assert lib.quisk_cuda_rxa_set_shift(rxa, 0, None) == 0
    assert lib.quisk_cuda_rxa_set_nc(rxa, nc) == 0
    assert lib.quisk_cuda_rxa_set_mode(rxa, 1) == 0
    assert lib.quisk_cuda_rxa_set_passband(rxa, 150.0, 2850.0) == 0
    assert lib.quisk_cuda_rxa_set_agc_mode(rxa, 3) == 0


def _exchange_blocks(lib, rxa, x, in_size, out_size, nblocks, before=None, collect_meters=False):
    ys, mts = [], []
    for b in range(nblocks):
        if before is not None:
            before(b)
        hin = np.ascontiguousarray(np.stack([x[b * in_size:(b + 1) * in_size]] * NCH))
        hout = np.full((NCH, out_size), -7.0 - 7.0j, dtype=np.complex128)
        err = C.c_int(9)
        assert lib.quisk_cuda_rxa_fexchange0(rxa, hin.ctypes.data, hout.ctypes.data, C.byref(err)) == 0, lib.quisk_cuda_last_error()
        assert err.value == 0
        ys.append(hout)
        if collect_meters:
            row = np.zeros((3, 3, NCH))
            for w in range(3):
                assert lib.quisk_cuda_rxa_get_meter(rxa, w, row[w, 0].ctypes.data, row[w, 1].ctypes.data, row[w, 2].ctypes.data) == 0
            # reference order (RXA.h:47-57): S_PK, S_AV, ADC_PK, ADC_AV, AGC_GAIN, AGC_PK, AGC_AV
            mts.append(np.stack([row[1, 1], row[1, 0], row[0, 1], row[0, 0], row[2, 2], row[2, 1], row[2, 0]], axis=0))
    return np.concatenate(ys, axis=1), (np.array(mts) if collect_meters else None)


def test_rxa_c3_full_geometry(torch, lib, kat):
    """SURVEY 8(d) C3 as stated: OpenChannel(1024, 1024, 192 k everywhere), RXASetNC(4096), USB 150-2850, AGC mode 3,
    through the fexchange0 entry; the three meters after every block against GetRXAMeter."""
    x = sig(1024 * 16, 710, 192000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2), (-30000.0, 0.25)))
    rxa = lib.quisk_cuda_rxa_create(NCH, 1024, 1024, 192000, 192000, 192000)
    assert rxa, lib.quisk_cuda_last_error()
    _setup_usb(lib, rxa, 4096)
    y, mt = _exchange_blocks(lib, rxa, x, 1024, 1024, 16, collect_meters=True)
    ref = kat["rxa_c3/y"]
    tol = _cond_bound(kat, "rxa_c3")
    errs = [O.rel_rms(y[c], ref) for c in range(NCH)]
    print("rxa_c3 rel-rms", errs, "bound", tol)
    assert max(errs) < tol
    rm = kat["rxa_c3/meters"]
    for c in range(NCH):
        assert np.max(np.abs(mt[:, :, c] - rm)) < 1e-9, np.max(np.abs(mt[:, :, c] - rm), axis=0)
    lib.quisk_cuda_rxa_destroy(rxa)


def test_rxa_meters_parity(torch, lib, kat):
    """xmeter (meter.c:75-107) for all three meters of the chain, every block: average, peak and AGC gain in dB."""
    x = sig(256 * 24, 700, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
    rxa = lib.quisk_cuda_rxa_create(NCH, 256, 256, 48000, 48000, 48000)
    _setup_usb(lib, rxa)
    y, mt = _exchange_blocks(lib, rxa, x, 256, 256, 24, collect_meters=True)
    rm = kat["rxa_usb/meters"]
    assert rm.shape == (24, 7) and np.all(rm[2:, :4] > -120) and np.all(rm[2:, :4] < 10)      # the fixture is not vacuous
    for c in range(NCH):
        d = np.abs(mt[:, :, c] - rm)
        assert np.max(d) < 1e-9, np.max(d, axis=0)
    assert max(O.rel_rms(y[c], kat["rxa_usb/y"]) for c in range(NCH)) < _cond_bound(kat, "rxa_usb")
    lib.quisk_cuda_rxa_destroy(rxa)


@pytest.mark.parametrize("in_size,nblocks", [(64, 96), (1024, 6)])
def test_rxa_exchange_reblocking(in_size, nblocks, torch, lib, kat):
    """in_size != dsp_insize (create_iobuffs / fexchange0 / dexchange, iobuffs.c:385-420, 464-516, 583-604): four calls
    per DSP turn, and four DSP turns per call."""
    x = sig(256 * 24, 720, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
    rxa = lib.quisk_cuda_rxa_create(NCH, in_size, 256, 48000, 48000, 48000)
    assert rxa, lib.quisk_cuda_last_error()
    _setup_usb(lib, rxa)
    i_s, o_s = C.c_int(0), C.c_int(0)
    assert lib.quisk_cuda_rxa_exchange_sizes(rxa, C.byref(i_s), C.byref(o_s)) == 0 and (i_s.value, o_s.value) == (in_size, in_size)
    y, _ = _exchange_blocks(lib, rxa, x, in_size, in_size, nblocks)
    ref = kat["rxa_reblock_%d_256/y" % in_size]
    nz = np.nonzero(ref)[0][0]
    assert nz == np.nonzero(y[0])[0][0]                     # same latency to the sample
    for c in range(NCH):
        assert O.rel_rms(y[c], ref) < _cond_bound(kat, "rxa_usb")
    lib.quisk_cuda_rxa_destroy(rxa)


def test_rxa_stop_start(torch, lib, kat):
    """SetChannelState(0) mid-stream: downslew0 on the way out (10 ms ramp, out_size + 1 zeros), then the exchange
    switches off (calls return without touching `out`) and the channel is flushed the way the reference flushes it;
    SetChannelState(1): up-slew from BEGIN on the flushed channel.  Outputs and meters after every call."""
    x = sig(256 * 32, 730, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
    # in_size 64 / dsp_size 256, stop on a call that starts a DSP block: the one geometry in which the reference itself
    # is deterministic (see the generator: otherwise its DSP and flush threads race for the last block)
    rxa = lib.quisk_cuda_rxa_create(NCH, 64, 256, 48000, 48000, 48000)
    assert lib.quisk_cuda_rxa_set_slew_down(rxa, 0.0, 0.010) == 0
    assert lib.quisk_cuda_rxa_set_slew(rxa, 0.010, 0.025) == 0
    _setup_usb(lib, rxa)

    def before(b):
        if b == 40:
            assert lib.quisk_cuda_rxa_set_channel_state(rxa, 0, 0) == 1
        if b == 72:
            assert lib.quisk_cuda_rxa_set_channel_state(rxa, 1, 0) == 0
    y, mt = _exchange_blocks(lib, rxa, x, 64, 64, 128, before=before, collect_meters=True)
    ref = kat["rxa_stop_start/y"]
    untouched = np.all(ref.reshape(128, 64) == -7.0 - 7.0j, axis=1)
    assert untouched.sum() >= 16                                             # the fixture really has exchange-off calls
    assert np.array_equal(np.all(y[0].reshape(128, 64) == -7.0 - 7.0j, axis=1), untouched)
    live = np.repeat(~untouched, 64)
    for c in range(NCH):
        assert O.rel_rms(y[c][live], ref[live]) < _cond_bound(kat, "rxa_usb")
    rm = kat["rxa_stop_start/meters"]
    for c in range(NCH):
        d = np.abs(mt[:, :, c] - rm)
        # below -300 dB a meter reads the FFT rounding floor of an all-zero input just after the flush (1e-22 samples
        # on top of the 1e-40 the meter adds before its logarithm): implementation noise in the reference too, not compared
        assert np.max(d[rm > -300]) < 1e-9, np.max(d, axis=0)
        assert np.array_equal(rm <= -399.9, mt[:, :, c] <= -399.9)          # and the flushed (-400) readings agree call for call
    lib.quisk_cuda_rxa_destroy(rxa)


def _wdsp_cdll():
    """What quisk_wdsp.py does with libwdsp.so (quisk_wdsp.py:28), on our library."""
    from quisk_b200 import lib as L
    w = C.CDLL(L.LIB_PATH)
    w.GetRXAMeter.restype = C.c_double
    return w


def test_wdsp_compat_quisk_open_sequence(torch, lib, kat):
    """quisk_wdsp.py's own sequence, call for call (Cwdsp.__init__ :45-67 and open() :69-99), with ctypes' default
    argument conversion like the reference module uses, then Quisk's re-blocker wdspFexchange0 (quisk_wdsp.c:22-69) on
    ragged sample counts scaled to CLIP32 -- against the compiled reference driven the same way."""
    from tests.golden.make_golden_wdsp import QUISK_SPLITS
    w = _wdsp_cdll()
    assert w.GetWDSPVersion() == 125
    fpt = C.cast(w.fexchange0, C.c_void_p).value
    assert fpt                                                               # the address quisk_wdsp_set_parameter stores
    channel, in_size, dsp_size = 1, 256, 256
    w.quisk_cuda_wdsp_set_parameter(channel, in_size, -1)
    w.OpenChannel(channel, in_size, dsp_size, 48000, 48000, 48000, 0, 1,
                  C.c_double(0.010), C.c_double(0.025), C.c_double(0.0), C.c_double(0.010), 1)
    w.SetRXAShiftRun(channel, 0)
    w.RXANBPSetRun(channel, 0)
    w.SetRXAAMSQRun(channel, 0)
    w.SetRXAMode(channel, 1)
    w.RXASetPassband(channel, C.c_double(300.0), C.c_double(3000.0))
    w.RXASetNC(channel, dsp_size)
    w.RXASetMP(channel, 0)
    w.SetRXAAGCMode(channel, 0)
    w.SetRXAAGCFixed(channel, C.c_double(0.0))
    w.SetRXAPanelRun(channel, 0)
    w.SetRXAEMNRRun(channel, 0)
    w.quisk_cuda_wdsp_set_parameter(channel, -1, 1)                          # in_use = wdsp_NR2 + wdsp_SNB (quisk.py:6027)
    xq = sig(6000, 740, 48000.0) * 2.0 ** 30
    ys, counts, pos = [], [], 0
    for n in QUISK_SPLITS:
        buf = np.zeros(n + 1024, dtype=np.complex128); buf[:n] = xq[pos:pos + n]; pos += n
        k = w.wdspFexchange0(channel, buf.ctypes.data_as(C.c_void_p), n)
        ys.append(buf[:k].copy()); counts.append(k)
    assert counts == kat["quisk_reblock/counts"].tolist()
    y = np.concatenate(ys); ref = kat["quisk_reblock/y"]
    assert np.nonzero(y)[0][0] == np.nonzero(ref)[0][0]
    assert O.rel_rms(y, ref) < 1e-14                                         # gain, scaling and ramps only in this configuration
    # in_use = 0: samples pass through untouched and the re-blocker's indices restart (quisk_wdsp.c:32-37)
    w.quisk_cuda_wdsp_set_parameter(channel, -1, 0)
    buf = xq[:300].copy()
    assert w.wdspFexchange0(channel, buf.ctypes.data_as(C.c_void_p), 300) == 300 and np.array_equal(buf, xq[:300])
    w.CloseChannel(channel)


def test_wdsp_compat_channel_api(torch, lib, kat):
    """The reference's signatures by channel number against the same fixtures the batched handle is held to: USB channel
    opened with Quisk's slew times, the notch database, GetRXAMeter, RXAGetaSipF1, two channels open at once."""
    w = _wdsp_cdll()
    D = C.c_double
    x = sig(256 * 24, 700, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
    xs = x.copy(); xs[:100] = 0.0
    xn = sig(256 * 24, 700, 48000.0, tones=((1000.0, 0.3), (1500.0, 0.2), (2200.0, 0.1)))
    for ch, slew in ((5, (0.010, 0.025, 0.0, 0.010)), (6, (0.0, 0.0, 0.0, 0.0))):
        w.OpenChannel(ch, 256, 256, 48000, 48000, 48000, 0, 1, D(slew[0]), D(slew[1]), D(slew[2]), D(slew[3]), 1)
        w.SetRXAShiftRun(ch, 0); w.RXASetNC(ch, 2048); w.SetRXAMode(ch, 1)
        w.RXASetPassband(ch, D(150.0), D(2850.0)); w.SetRXAAGCMode(ch, 3)
    w.RXANBPSetTuneFrequency(6, D(7000000.0))
    w.RXANBPSetNotchesRun(6, 1)
    assert w.RXANBPAddNotch(6, 0, D(7001500.0), D(400.0), 1) == 0
    assert w.RXANBPAddNotch(6, 1, D(7002400.0), D(100.0), 1) == 0
    nn = C.c_int(0); w.RXANBPGetNumNotches(6, C.byref(nn)); assert nn.value == 2
    ys = {5: [], 6: []}
    err = C.c_int(0)
    for b in range(24):
        for ch, src in ((5, xs), (6, xn)):                                   # interleaved: the channels are independent
            inb = np.ascontiguousarray(src[b * 256:(b + 1) * 256]); outb = np.zeros(256, dtype=np.complex128)
            w.fexchange0(ch, inb.ctypes.data_as(C.c_void_p), outb.ctypes.data_as(C.c_void_p), C.byref(err))
            assert err.value == 0
            ys[ch].append(outb)
    tol = _cond_bound(kat, "rxa_usb")
    assert O.rel_rms(np.concatenate(ys[5]), kat["rxa_usb_slew/y"]) < tol
    assert O.rel_rms(np.concatenate(ys[6]), kat["rxa_usb_notch/y"]) < tol
    for mt in range(7):
        v = w.GetRXAMeter(5, mt)
        assert np.isfinite(v) and -200 < v < 100
    sip = np.zeros(2 * 1024, dtype=np.float32)
    w.RXAGetaSipF1(5, sip.ctypes.data_as(C.c_void_p), 1024)
    assert np.any(sip != 0)
    assert w.SetChannelState(5, 0, 0) == 1 and w.SetChannelState(5, 0, 0) == 0
    w.CloseChannel(5); w.CloseChannel(6)
    outb = np.full(256, 3.0 + 0j)
    w.fexchange0(5, x[:256].ctypes.data_as(C.c_void_p), outb.ctypes.data_as(C.c_void_p), C.byref(err))      # closed: a message, no crash
    assert np.all(outb == 3.0)


@pytest.mark.parametrize("geom", [(256, 2048, 48000, 24), (1024, 4096, 192000, 12)])
@pytest.mark.parametrize("agc_mode", [3, 1, 0])
def test_rxa_fused_kernel_equals_per_stage_kernels(geom, agc_mode, torch, lib):
    """wdsp_rxa_fused.cu (the chain as one kernel, speculative AGC lane, meters on their own lanes) against the per-stage
    kernels it replaces, for one block per launch, for all blocks in one launch, and with the stream switching between
    the two paths from block to block (they share every state array).  Multi-block and single-block launches of the fused
    kernel are the same bits; against the per-stage kernels the transforms are two compilations of one source (the
    compiler contracts multiply-adds differently in the two contexts), so outputs agree to 1e-13 and meters to 1e-9 dB."""
    size, nc, rate, nblk = geom
    x = sig(size * nblk, 760 + agc_mode, float(rate), tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
    x[4 * size:6 * size] *= 0.02; x[8 * size:] *= 3.0          # level steps walk the AGC through its states
    d = _dev(torch, x)

    def run(kind):
        rxa = lib.quisk_cuda_rxa_create(NCH, size, size, rate, rate, rate)
        assert lib.quisk_cuda_rxa_set_shift(rxa, 0, None) == 0
        assert lib.quisk_cuda_rxa_set_nc(rxa, nc) == 0
        assert lib.quisk_cuda_rxa_set_mode(rxa, 1) == 0
        assert lib.quisk_cuda_rxa_set_passband(rxa, 150.0, 2850.0) == 0
        assert lib.quisk_cuda_rxa_set_agc_mode(rxa, agc_mode) == 0
        o = torch.zeros_like(d)
        if kind.startswith("multi"):
            # the per-channel sequential kernel has three forms chosen by the channel count (one fat CTA per SM, two, or four
            # three-warp CTAs over sub-blocks of the DSP block): force each of them on this small batch
            env = {"multi": None, "multi_two": ("2", "0"), "multi_thin": ("4", "0"), "multi_thin64": ("4", "64")}[kind]
            if env:
                os.environ["QUISK_RXA_MINB"], os.environ["QUISK_RXA_THIN_NS"] = env
            try:
                assert lib.quisk_cuda_rxa_xrxa_multi(rxa, d.data_ptr(), d.stride(0), o.data_ptr(), o.stride(0), nblk, None) == 0, lib.quisk_cuda_last_error()
            finally:
                os.environ.pop("QUISK_RXA_MINB", None); os.environ.pop("QUISK_RXA_THIN_NS", None)
        else:
            for b in range(nblk):
                fused = {"fused": 1, "stages": 0, "mixed": b % 3 != 1}[kind]
                assert lib.quisk_cuda_rxa_set_option(rxa, 1, int(fused)) == 0
                blk = d[:, b * size:(b + 1) * size]; ob = o[:, b * size:(b + 1) * size]
                assert lib.quisk_cuda_rxa_xrxa(rxa, blk.data_ptr(), d.stride(0), ob.data_ptr(), o.stride(0), None) == 0, lib.quisk_cuda_last_error()
        torch.cuda.synchronize()
        mt = np.zeros((3, 3, NCH))
        for w in range(3):
            assert lib.quisk_cuda_rxa_get_meter(rxa, w, mt[w, 0].ctypes.data, mt[w, 1].ctypes.data, mt[w, 2].ctypes.data) == 0
        sip = np.zeros((NCH, 2 * 1024), dtype=np.float32)
        assert lib.quisk_cuda_rxa_get_siphon(rxa, sip.ctypes.data, 1024, 1) == 0
        lib.quisk_cuda_rxa_destroy(rxa)
        return o.cpu().numpy(), mt, sip
    ref = run("stages")
    assert np.abs(ref[0]).max() > 0.1
    res = {kind: run(kind) for kind in ("fused", "multi", "mixed")}
    for kind, got in res.items():
        assert max(O.rel_rms(got[0][c], ref[0][c]) for c in range(NCH)) < 1e-13, kind
        assert np.max(np.abs(got[1] - ref[1])) < 1e-9, (kind, got[1] - ref[1])
        assert np.max(np.abs(got[2] - ref[2])) <= 1e-6 * np.max(np.abs(ref[2])), kind
    # one launch for all blocks runs the transforms as their own wide kernels: again the same source compiled in another context
    assert max(O.rel_rms(res["multi"][0][c], res["fused"][0][c]) for c in range(NCH)) < 1e-13
    # the three forms of the per-channel kernel: same arithmetic in the same order -- outputs, meters and siphon bit for bit
    for kind in ("multi_two", "multi_thin", "multi_thin64"):
        got = run(kind)
        for a, b in zip(got, res["multi"]):
            assert np.array_equal(a, b), kind


@pytest.mark.parametrize("cfg", ["fm_384k", "am", "sam_usb", "usb_out96k"])
def test_rxa_multi_block_stage_groups_equal_per_block(cfg, torch, lib):
    """quisk_cuda_rxa_xrxa_multi for the configurations outside the single-kernel chain (FM with the /8 input resampler, AM,
    synchronous AM, an output resampler): every stage takes a group of blocks per launch -- the fircores as wide transform
    grids, the recurrent stages and meters over the whole group (meters still peak-hold per block) -- and must give what
    one xrxa call per block gives: same state machine trajectories, outputs to 1e-12, meters to 1e-9 dB."""
    geo = {"fm_384k": (2048, 256, 384000, 48000, 48000, 5, 40), "am": (256, 256, 48000, 48000, 48000, 6, 20),
           "sam_usb": (256, 256, 48000, 48000, 48000, 10, 20), "usb_out96k": (256, 256, 48000, 48000, 96000, 1, 20)}[cfg]
    in_size, dsp_size, in_rate, dsp_rate, out_rate, mode, nblk = geo
    x = (fm_sig if mode == 5 else am_sig if mode in (6, 10) else sig)(in_size * nblk, 780, float(in_rate))
    d = _dev(torch, x)

    def run(multi):
        rxa = lib.quisk_cuda_rxa_create(NCH, in_size, dsp_size, in_rate, dsp_rate, out_rate)
        assert rxa, lib.quisk_cuda_last_error()
        assert lib.quisk_cuda_rxa_set_shift(rxa, 0, None) == 0
        assert lib.quisk_cuda_rxa_set_mode(rxa, mode) == 0
        assert lib.quisk_cuda_rxa_set_passband(rxa, *((-8000.0, 8000.0) if mode == 5 else (-4000.0, 4000.0) if mode in (6, 10) else (150.0, 2850.0))) == 0
        osz = lib.quisk_cuda_rxa_out_size(rxa)
        o = torch.zeros((NCH, osz * nblk), dtype=torch.complex128, device="cuda")
        if multi:
            assert lib.quisk_cuda_rxa_xrxa_multi(rxa, d.data_ptr(), d.stride(0), o.data_ptr(), o.stride(0), nblk, None) == 0, lib.quisk_cuda_last_error()
        else:
            for b in range(nblk):
                blk = d[:, b * in_size:(b + 1) * in_size]; ob = o[:, b * osz:(b + 1) * osz]
                assert lib.quisk_cuda_rxa_xrxa(rxa, blk.data_ptr(), d.stride(0), ob.data_ptr(), o.stride(0), None) == 0, lib.quisk_cuda_last_error()
        torch.cuda.synchronize()
        mt = np.zeros((3, 3, NCH))
        for w in range(3):
            assert lib.quisk_cuda_rxa_get_meter(rxa, w, mt[w, 0].ctypes.data, mt[w, 1].ctypes.data, mt[w, 2].ctypes.data) == 0
        sip = np.zeros((NCH, 2 * 1024), dtype=np.float32)
        assert lib.quisk_cuda_rxa_get_siphon(rxa, sip.ctypes.data, 1024, 1) == 0
        lib.quisk_cuda_rxa_destroy(rxa)
        return o.cpu().numpy(), mt, sip
    ref, got = run(False), run(True)
    assert np.abs(ref[0]).max() > 1e-3
    errs = [O.rel_rms(got[0][c], ref[0][c]) for c in range(NCH)]
    print(cfg, "multi vs per block", errs)
    assert max(errs) < 1e-12
    assert np.max(np.abs(got[1] - ref[1])) < 1e-9, got[1] - ref[1]
    assert np.max(np.abs(got[2] - ref[2])) <= 1e-6 * max(np.max(np.abs(ref[2])), 1e-30)


def test_fm_limiter_stage_bit_exact(torch, lib):
    """fmd's detector limiter (wdsp/fmd.c:49-73): a wcpAGC with calc_fmd's constants (mode 5, envelope detector, 1 ms
    attack, 8 ms decay, no hang) run in place.  Same operations as xwcpagc in the same order: identical bits."""
    kat = golden("wdsp_fmlim_kat.npz")
    n = 256
    st = lib.quisk_cuda_wcpagc_create_fmlim(NCH, 48000, 2.5)
    assert st, lib.quisk_cuda_last_error()
    d = _dev(torch, kat["lim_in"])
    for b in range(40):
        blk = d[:, b * n:(b + 1) * n]
        assert lib.quisk_cuda_seq_run(st, blk.data_ptr(), d.stride(0), blk.data_ptr(), d.stride(0), n, None) == 0
    torch.cuda.synchronize()
    y = d.cpu().numpy()
    ref = kat["lim_out"]
    assert np.abs(ref).max() > 0.5 and np.abs(ref).max() < 1.0          # limited to out_target
    for c in range(NCH):
        assert np.array_equal(y[c], ref)
    lib.quisk_cuda_seq_destroy(st)


def test_rxa_fm_channel_with_detector_limiter(torch, lib):
    """SetRXAFMLimRun(1) and SetRXAFMLimGain mid-stream (fmd.c:337-363: calc_fmd rebuilds the limiter and zeroes the PLL) on
    an FM channel at 48 kS/s, against fexchange0 of the compiled reference.  The limiter's volts machine takes its decisions
    on comparisons of the demodulated audio, so the reference itself moves by `cond` when its input moves by one ulp; the
    stage alone is bit-exact (test above)."""
    from tests.golden.make_golden_wdsp_fmlim import BLOCKS, GAIN_AT, N, TAIL
    kat = golden("wdsp_fmlim_kat.npz")
    x = _first_sample_swallowed(fm_sig(N * BLOCKS, 810, 48000.0))
    rxa = lib.quisk_cuda_rxa_create(NCH, N, N, 48000, 48000, 48000)
    assert rxa, lib.quisk_cuda_last_error()
    lib.quisk_cuda_rxa_set_shift(rxa, 0, None)
    assert lib.quisk_cuda_rxa_set_mode(rxa, 5) == 0
    assert lib.quisk_cuda_rxa_set_passband(rxa, -8000.0, 8000.0) == 0
    assert lib.quisk_cuda_rxa_set_fm_lim_run(rxa, 1) == 0
    d = _dev(torch, x)
    o = torch.zeros_like(d)
    for b in range(BLOCKS):
        if b == GAIN_AT:
            assert lib.quisk_cuda_rxa_set_fm_lim_gain(rxa, -6.0) == 0
        assert lib.quisk_cuda_rxa_xrxa(rxa, d[:, b * N:(b + 1) * N].data_ptr(), d.stride(0), o[:, b * N:(b + 1) * N].data_ptr(), o.stride(0), None) == 0
    torch.cuda.synchronize()
    y = o.cpu().numpy()
    # fexchange0 hands out block b's result two calls later: the reference's segment [GAIN_AT - TAIL, GAIN_AT) + last TAIL blocks are our
    # blocks shifted by two
    seg = np.concatenate([y[:, (GAIN_AT - TAIL - 2) * N:(GAIN_AT - 2) * N], y[:, (BLOCKS - TAIL - 2) * N:(BLOCKS - 2) * N]], axis=1)
    half = TAIL * N
    for part, sl in enumerate((slice(0, half), slice(half, 2 * half))):
        errs = [O.rel_rms(seg[c][sl], kat["y_seg"][sl]) for c in range(NCH)]
        print("rxa fm + limiter, segment", part, errs, "reference's own one-ulp sensitivity", kat["conds"][part])
        assert max(errs) < max(1e-12, 20.0 * float(kat["conds"][part]))
    lib.quisk_cuda_rxa_destroy(rxa)
