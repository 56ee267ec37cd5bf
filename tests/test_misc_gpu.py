"""GPU parity of process_agc, cFracDecim (fixtures from the compiled reference) and get_bandscope (NumPy oracle)."""
import ctypes as C

import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.golden.make_golden_misc import (AGC_SPLITS, AN_RATE, AN_SIDETONES, AN_SPLITS, NB_CASES, NB_SPLITS, SQ_BW, SQ_LEVELS,
                                           SQ_RATE, SQ_SPLITS, agc_input, an_input, nb_input, sq_input)
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


@pytest.mark.parametrize("is_cpx", [1, 0])
def test_process_agc(is_cpx, torch, lib):
    kat = golden("misc_kat.npz")
    x = agc_input(sum(AGC_SPLITS), 60)
    if not is_cpx:
        x = x.real.astype(np.complex128)
    d = torch.from_numpy(np.stack([x] * NCH)).cuda()
    a = lib.quisk_cuda_agc_create(NCH, 48000, 0.7, 80.0, 1.0)
    assert a
    pos = 0
    for n in AGC_SPLITS:
        blk = d[:, pos:pos + n]
        assert lib.quisk_cuda_agc_run(a, blk.data_ptr(), d.stride(0), n, is_cpx, None) == 0
        pos += n
    torch.cuda.synchronize()
    y = d.cpu().numpy()
    for c in range(NCH):
        assert O.rel_rms(y[c], kat["agc_cpx%d/y" % is_cpx]) < 1e-13
    lib.quisk_cuda_agc_destroy(a)


@pytest.mark.parametrize("fdecim", [1.25, 1.0416666666666667, 1.5])
def test_cfracdecim(fdecim, torch, lib):
    kat = golden("misc_kat.npz")
    x = O.synth_iq(6000, 61, 1.0)
    d = torch.from_numpy(np.stack([x] * NCH)).cuda()
    f = lib.quisk_cuda_fracdecim_create(NCH)
    ys, counts, pos = [], [], 0
    for n in [1000, 1, 2, 997, 4000]:
        blk = d[:, pos:pos + n].contiguous(); pos += n
        o = torch.zeros((NCH, n + 1), dtype=torch.complex128, device="cuda")
        k = C.c_int(0)
        assert lib.quisk_cuda_fracdecim_run(f, blk.data_ptr(), n, n, fdecim, o.data_ptr(), n + 1, C.byref(k), None) == 0
        torch.cuda.synchronize()
        ys.append(o[:, :k.value].cpu().numpy()); counts.append(k.value)
    y = np.concatenate(ys, axis=1)
    assert counts == kat["fracdecim_%g/counts" % fdecim].tolist()
    for c in range(NCH):
        assert O.rel_rms(y[c], kat["fracdecim_%g/y" % fdecim]) < 1e-14
    lib.quisk_cuda_fracdecim_destroy(f)


@pytest.mark.parametrize("rate,level", NB_CASES)
def test_noise_blanker(rate, level, torch, lib):
    """NoiseBlanker (quisk.c:679-784) against the compiled reference's fixture: bit-exact, ragged blocks, the state
    (delay line, running magnitude sum, blanking / ramp phase) carried from call to call; level 0 leaves everything alone."""
    kat = golden("misc_kat.npz")
    x = nb_input(sum(NB_SPLITS), 90)
    d = torch.from_numpy(np.stack([x, x * 0.5, x])).cuda()
    h = lib.quisk_cuda_nb_create(NCH, rate)
    assert h
    before = d.clone()
    assert lib.quisk_cuda_nb_run(h, d.data_ptr(), d.stride(0), 100, 0, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(d, before)
    pos = 0
    for n in NB_SPLITS:
        blk = d[:, pos:pos + n]
        assert lib.quisk_cuda_nb_run(h, blk.data_ptr(), d.stride(0), n, level, None) == 0
        pos += n
    torch.cuda.synchronize()
    y = d.cpu().numpy()
    ref = kat["nb_%d_%d/y" % (rate, level)]
    assert np.array_equal(y[0], ref) and np.array_equal(y[2], ref)
    assert np.array_equal(y[1], ref * 0.5)          # scaling by a power of two changes no decision
    lib.quisk_cuda_nb_destroy(h)


def test_noise_blanker_long_blocks(torch, lib):
    """One call of 70 000 samples equals the oracle fed the same stream in blocks (chunking inside the kernel, quiet
    fast path and state-machine path both taken); channels are independent."""
    n = 70000
    xs = []
    for c in range(NCH):
        x = O.synth_iq(n, 95 + c, 1.0)
        x[5000 + 777 * c] *= 80.0; x[30000:30005] *= 50.0; x[69990] *= 90.0
        xs.append(x)
    d = torch.from_numpy(np.stack(xs)).cuda()
    h = lib.quisk_cuda_nb_create(NCH, 192000)
    assert lib.quisk_cuda_nb_run(h, d.data_ptr(), d.stride(0), n, 1, None) == 0
    torch.cuda.synchronize()
    y = d.cpu().numpy()
    for c in range(NCH):
        assert np.array_equal(y[c], O.NoiseBlanker(192000, 1)(xs[c]))
    lib.quisk_cuda_nb_destroy(h)


@pytest.mark.parametrize("level", SQ_LEVELS)
def test_ssb_squelch(level, torch, lib):
    """ssb_squelch + d_delay (quisk.c:1056-1180) against the compiled reference's fixture: the timer and squelch_active
    after every call (ragged calls from 1 to 8000 samples, several frames per call, the set-up-only first call) and
    the delayed audio, bit-exact; a second channel with a different stream stays independent."""
    kat = golden("misc_kat.npz")
    x = sq_input(sum(SQ_SPLITS), 91)
    x2 = (2.0 ** 20) * np.random.default_rng(5).standard_normal(len(x))
    d = torch.from_numpy(np.stack([x, x2, x])).cuda()
    h = lib.quisk_cuda_ssb_squelch_create(NCH, SQ_RATE, SQ_BW)
    assert h
    act, opn, pos = [], [], 0
    so = (C.c_int * NCH)(); sa = (C.c_int * NCH)()
    for n in SQ_SPLITS:
        blk = d[:, pos:pos + n]
        assert lib.quisk_cuda_ssb_squelch_run(h, blk.data_ptr(), d.stride(0), n, level, None) == 0
        assert lib.quisk_cuda_ssb_squelch_state(h, so, sa, None) == 0
        assert so[0] == so[2] and sa[0] == sa[2]
        act.append(sa[0]); opn.append(so[0]); pos += n
    assert opn == kat["sq_%d/sq_open" % level].tolist()
    assert act == kat["sq_%d/active" % level].tolist()
    y = d.cpu().numpy()
    assert np.array_equal(y[0], kat["sq_%d/y" % level]) and np.array_equal(y[2], y[0])
    ref2 = O.SsbSquelch(SQ_RATE, SQ_BW, level)
    pos = 0
    for n in SQ_SPLITS:
        ref2(x2[pos:pos + n]); pos += n
    assert (so[1], sa[1]) == (ref2.sq_open, ref2.active)
    assert lib.quisk_cuda_ssb_squelch_run(h, d.data_ptr(), d.stride(0), 8193, level, None) != 0
    lib.quisk_cuda_ssb_squelch_destroy(h)


@pytest.mark.parametrize("sidetone", AN_SIDETONES)
def test_auto_notch(sidetone, torch, lib):
    """dAutoNotch (quisk.c:786-963) against the compiled reference's fixture (FFTW calls through the oracle's shim):
    ragged calls from 1 to 8000 samples, several frames per call; the stream contains the phases without a notch, with
    one and with two notches, so every filter design and the hysteresis counters are exercised.  FP64 tolerance 1e-12
    relative RMS (two different FFT implementations); a wrong decision anywhere would be an error of order one."""
    kat = golden("misc_kat.npz")
    x = an_input(sum(AN_SPLITS), 92)
    d = torch.from_numpy(np.stack([x, 0.25 * x, x])).cuda()
    h = lib.quisk_cuda_autonotch_create(NCH, AN_RATE)
    assert h
    pos = 0
    for n in AN_SPLITS:
        blk = d[:, pos:pos + n]
        assert lib.quisk_cuda_autonotch_run(h, blk.data_ptr(), d.stride(0), n, sidetone, None) == 0
        pos += n
    torch.cuda.synchronize()
    y = d.cpu().numpy()
    ref = kat["notch_%d/y" % sidetone]
    assert O.rel_rms(y[0], ref) < 1e-12 and np.array_equal(y[0], y[2])
    assert O.rel_rms(y[1], 0.25 * ref) < 1e-12
    for lo in range(0, len(ref) - 2048, 2048):      # and block by block, so a short wrong stretch cannot hide in the total
        assert O.rel_rms(y[0][lo:lo + 2048], ref[lo:lo + 2048]) < 1e-11
    lib.quisk_cuda_autonotch_destroy(h)


def test_rx_options_host_mirror(torch, lib):
    """quisk_b200.rx.RxOptions = set_noise_blanker / set_auto_notch / set_ssb_squelch + what the sample thread does with
    them: off means untouched, on means the oracle's output; set_auto_notch restarts the notch like the reference."""
    from quisk_b200.rx import RxOptions
    opt = RxOptions(NCH, 192000, AN_RATE, SQ_BW)
    xn = O.synth_iq(5000, 9, 1.0); xn[1500] *= 70.0
    d = torch.from_numpy(np.stack([xn] * NCH)).cuda()
    opt.run_noise_blanker(d.data_ptr(), d.stride(0), 5000)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy()[0], xn)                       # level 0: nothing happens
    opt.set_noise_blanker(3)
    opt.run_noise_blanker(d.data_ptr(), d.stride(0), 5000)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy()[NCH - 1], O.NoiseBlanker(192000, 3)(xn))
    xa = an_input(6000, 92)
    a = torch.from_numpy(np.stack([xa] * NCH)).cuda()
    opt.run_auto_notch(a.data_ptr(), a.stride(0), 6000)
    assert opt.run_ssb_squelch(a.data_ptr(), a.stride(0), 6000) is None
    torch.cuda.synchronize()
    assert np.array_equal(a.cpu().numpy()[0], xa)                       # both off
    for _ in range(2):                                                  # the second switch-on starts from scratch again
        opt.set_auto_notch(1)
        a = torch.from_numpy(np.stack([xa] * NCH)).cuda()
        opt.run_auto_notch(a.data_ptr(), a.stride(0), 6000)
        torch.cuda.synchronize()
        assert O.rel_rms(a.cpu().numpy()[1], O.AutoNotch(AN_RATE, 0)(xa)) < 1e-12
    opt.set_ssb_squelch(1, 150)
    sq = O.SsbSquelch(AN_RATE, SQ_BW, 150)
    xs = sq_input(3000, 91)
    for k in range(3):
        blk = torch.from_numpy(np.stack([xs[k * 1000:(k + 1) * 1000]] * NCH)).cuda()
        act = opt.run_ssb_squelch(blk.data_ptr(), blk.stride(0), 1000)
        ref = sq(xs[k * 1000:(k + 1) * 1000])
        assert act == [sq.active] * NCH and np.array_equal(blk.cpu().numpy()[0], ref)
    opt.close()


def test_bandscope(torch, lib):
    size, nblk, gw = 4096, 3, 800
    rng = np.random.default_rng(8)
    t = np.arange(size * nblk)
    blocks = np.stack([(2.0 ** 20) * np.sin(2 * np.pi * (0.05 + 0.03 * s) * t) + 1000.0 * rng.standard_normal(size * nblk) for s in range(NCH)])
    d = torch.from_numpy(blocks).cuda()
    b = lib.quisk_cuda_bandscope_create(NCH, size)
    assert b
    assert lib.quisk_cuda_bandscope_accumulate(b, d.data_ptr(), size * nblk, nblk, None) == 0
    g = torch.zeros((NCH, gw), dtype=torch.float64, device="cuda")
    assert lib.quisk_cuda_bandscope_graph(b, gw, 122880000, 0.7, 1.0e6, g.data_ptr(), None) == 0
    torch.cuda.synchronize()
    g = g.cpu().numpy()
    for s in range(NCH):
        ref = O.bandscope(blocks[s].reshape(nblk, size), gw, 122880000, 0.7, 1.0e6)
        assert np.max(np.abs(g[s] - ref)) < 1e-9
    lib.quisk_cuda_bandscope_destroy(b)


@pytest.mark.parametrize("big", [0, 1])
@pytest.mark.parametrize("nb", [1, 2, 3, 4])
def test_unpack_iq(nb, big, torch, lib):
    """Device unpack of add_rx_samples' wire format, bit-exact vs the compiled reference loops (quisk.c:2922-2953)."""
    from tests.golden.make_golden_misc import ingest_bytes
    kat = golden("misc_kat.npz")
    data = ingest_bytes(70 + nb, 1000 * 2 * nb)
    for off in (0, 1):                       # aligned and misaligned rows (the int16 / int32 fast paths need alignment)
        raw = np.zeros((NCH, len(data) + 8), dtype=np.uint8)
        raw[:, off:off + len(data)] = data
        d = torch.from_numpy(raw).cuda()
        out = torch.zeros((NCH, 1003), dtype=torch.complex128, device="cuda")
        rc = lib.quisk_cuda_unpack_iq(d.data_ptr() + off, raw.shape[1], NCH, 1000, nb, big, out.data_ptr(), 1003, None)
        assert rc == 0, lib.quisk_cuda_last_error()
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        for c in range(NCH):
            assert np.array_equal(got[c, :1000], kat["unpack_iq_%d_%d/y" % (nb, big)])
            assert np.all(got[c, 1000:] == 0)


@pytest.mark.parametrize("n_rx", [1, 2, 4, 10])
def test_unpack_hermes(n_rx, torch, lib):
    """Hermes / Metis protocol-1 de-interleave (quisk.c:3746-3763), bit-exact vs the compiled reference loop."""
    from tests.golden.make_golden_misc import ingest_bytes
    kat = golden("misc_kat.npz")
    pk = ingest_bytes(80 + n_rx, 3 * 1032)
    d = torch.from_numpy(pk).cuda()
    per = lib.quisk_cuda_hermes_samples_per_packet(n_rx)
    assert per == 2 * (504 // (6 * n_rx + 2))
    out = torch.zeros((n_rx, 3 * per + 5), dtype=torch.complex128, device="cuda")
    ns = C.c_int(0)
    rc = lib.quisk_cuda_unpack_hermes(d.data_ptr(), 3, n_rx, out.data_ptr(), 3 * per + 5, C.byref(ns), None)
    assert rc == 0, lib.quisk_cuda_last_error()
    torch.cuda.synchronize()
    assert ns.value == 3 * per
    assert np.array_equal(out.cpu().numpy()[:, :3 * per], kat["unpack_hermes_%d/y" % n_rx])


def test_rx_process_host_packed(torch, lib):
    """int16 little-endian host block through quisk_cuda_rx_process_host_packed == the same samples widened on the
    host (add_rx_samples) and fed through quisk_cuda_rx_process_host."""
    from quisk_b200.rx import RxChain, load_tables
    tabs = load_tables()
    kat = golden("chain_kat.npz")
    fi, fq = kat["c1/filt_i"], kat["c1/filt_q"]
    Cn, n = 3, 30720
    rng = np.random.default_rng(5)
    raw = rng.integers(0, 256, size=(Cn, n * 4), dtype=np.uint8)
    wide = np.stack([O.unpack_iq(raw[c], 2, False) for c in range(Cn)])
    rx1 = RxChain(Cn, 1536000, "USB", fi, fq, tabs, fused=True)
    rx2 = RxChain(Cn, 1536000, "USB", fi, fq, tabs, fused=True)
    a1 = np.zeros((Cn, rx1.max_out(n))); a2 = np.zeros_like(a1)
    n1 = rx1.process_host(np.ascontiguousarray(wide), n, a1)
    n2 = rx2.process_host_packed(raw, n, 2, False, a2)
    assert n1 == n2 == n // 32
    assert np.array_equal(a1[:, :n1], a2[:, :n2])
    rx1.close(); rx2.close()
