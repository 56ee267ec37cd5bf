"""GPU parity of the WDSP RXA stages (include/quisk_cuda_wdsp.h) against known-answer fixtures generated
from the compiled reference (tests/golden/make_golden_wdsp.py -> wdsp_kat.npz).  Every case runs several
identical channels through the batched kernels.  Tolerances: fircore-based stages 1e-12 relative RMS (the
north star's FP64 bound; our FFT is not FFTW's), recurrent stages 1e-12 as well, resampler bit-exact."""
import ctypes as C

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
    for c in range(NCH):
        assert O.rel_rms(y[c], ref) < 1e-5
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
    for c in range(NCH):
        assert O.rel_rms(y[c], kat["amd_%d_%d/y" % (mode, sb)]) < 1e-11
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
    for c in range(NCH):
        assert O.rel_rms(y[c], ref) < 1e-11
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
        assert O.rel_rms(y[c], ref) < 1e-11
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
        assert O.rel_rms(y[c], ref) < 1e-11
    lib.quisk_cuda_rxa_destroy(rxa)


def test_rxa_default_channel_keeps_bp1(torch, lib, kat):
    """No SetRXAMode: bp1 is still running -- the quirk is part of the reference's behaviour."""
    x = _first_sample_swallowed(sig(256 * 16, 701, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2))))
    rxa = lib.quisk_cuda_rxa_create(NCH, 256, 256, 48000, 48000, 48000)
    lib.quisk_cuda_rxa_set_shift(rxa, 0, None)
    y = _run_rxa(torch, lib, rxa, x, 256, 16)
    ref = kat["rxa_default/y"][512:]
    for c in range(NCH):
        assert O.rel_rms(y[c][:len(ref)], ref) < 1e-11
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
    for c in range(NCH):
        assert O.rel_rms(y[c], kat["fmd/y"]) < 1e-10


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
    for c in range(NCH):
        assert O.rel_rms(ours[c], ref) < 1e-8
    lib.quisk_cuda_rxa_destroy(rxa)


def test_rxa_am_channel(torch, lib, kat):
    x = _first_sample_swallowed(am_sig(256 * 16, 703, 48000.0))
    rxa = lib.quisk_cuda_rxa_create(NCH, 256, 256, 48000, 48000, 48000)
    lib.quisk_cuda_rxa_set_shift(rxa, 0, None)
    assert lib.quisk_cuda_rxa_set_mode(rxa, 6) == 0
    assert lib.quisk_cuda_rxa_set_passband(rxa, -4000.0, 4000.0) == 0
    y = _run_rxa(torch, lib, rxa, x, 256, 16)
    ref = kat["rxa_am/y"][512:]
    for c in range(NCH):
        assert O.rel_rms(y[c][:len(ref)], ref) < 1e-11
    lib.quisk_cuda_rxa_destroy(rxa)
