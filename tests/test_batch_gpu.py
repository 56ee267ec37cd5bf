"""GPU parity of the batched device-resident filters (quisk_cuda_batch_*) and the batched
receive chain (quisk_cuda_rx_*), through the C ABI, against the fixtures and the oracle."""
import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.golden.make_golden import RATES, DEMOD_TAPS, DGT_CASES, demod_taps
from tests.util import SPLITS, CHAIN_SPLITS, DEMOD_SPLITS, golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def tabs():
    return golden("quisk_tables.npz")


def _run_batch(torch, bf, x, splits, out_cap, real_out=None):
    """x: [C, n] numpy; returns [C, n_out] numpy and per-block counts."""
    C = x.shape[0]
    d_in = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    cplx = x.dtype == np.complex128
    outs, counts, pos = [], [], 0
    for n in splits:
        blk = d_in[:, pos:pos + n].contiguous(); pos += n
        out = torch.zeros((C, out_cap), dtype=torch.complex128 if cplx else torch.float64, device="cuda")
        k = bf.run(blk.data_ptr(), blk.stride(0) if n else 1, n, out.data_ptr(), out_cap)
        torch.cuda.synchronize()
        outs.append(out[:, :k].cpu().numpy()); counts.append(k)
    return np.concatenate(outs, axis=1), counts


BATCH_CASES = [
    ("cDecim2HB45", None, 1, 1, False),
    ("cDecimate", "quiskFilt48dec24Coefs", 1, 2, False),
    ("cDecimate", "quiskFilt240D5CoefsSharp", 1, 5, False),
    ("dDecimate", "quiskLpFilt48Coefs", 1, 4, True),
    ("cInterpolate", "quiskAudio24p4Coefs", 2, 1, False),
    ("dInterpolate", "quiskAudio24p3Coefs", 3, 1, True),
    ("cInterpDecim", "quiskFilt300D5Coefs", 6, 5, False),
    ("cInterpDecim", "quiskFilt240D5CoefsSharp", 4, 5, False),
    ("cInterp2HB45", None, 1, 1, False),
    ("dInterp2HB45", None, 1, 1, True),
]


def _oracle_for(kind, tab, interp, decim, real):
    dt = np.float64 if real else np.complex128
    return {"cDecim2HB45": lambda: O.HB45Decim(), "cDecimate": lambda: O.FirDecim(tab, decim),
            "dDecimate": lambda: O.FirDecim(tab, decim, np.float64), "cInterpolate": lambda: O.FirInterp(tab, interp),
            "dInterpolate": lambda: O.FirInterp(tab, interp, np.float64),
            "cInterpDecim": lambda: O.FirInterpDecim(tab, interp, decim),
            "cInterp2HB45": lambda: O.HB45Interp(np.complex128), "dInterp2HB45": lambda: O.HB45Interp(np.float64)}[kind]()


@pytest.mark.parametrize("case", BATCH_CASES, ids=["%s-%s" % (c[0], c[1]) for c in BATCH_CASES])
def test_batch_filter_vs_oracle(case, torch, tabs):
    from quisk_b200.rx import BatchFilter
    kind, tname, interp, decim, real = case
    tab = tabs[tname] if tname else None
    C = 5
    x = np.stack([O.synth_iq(sum(SPLITS), 30 + c, 1.0) for c in range(C)])
    if real:
        x = np.ascontiguousarray(x.real)
    bf = BatchFilter(kind, C, tab, interp, decim)
    y, counts = _run_batch(torch, bf, x, SPLITS, 2 * 6 * 2300)
    for c in range(C):
        st = _oracle_for(kind, tab, interp, decim, real)
        yo, co, pos = [], [], 0
        for n in SPLITS:
            o = st(x[c, pos:pos + n]); pos += n
            yo.append(o); co.append(len(o))
        assert co == counts
        assert O.rel_rms(y[c], np.concatenate(yo)) < 1e-13
    bf.close()


def test_batch_rxfilter_tap_order(torch):
    from quisk_b200.rx import BatchFilter
    rng = np.random.default_rng(5)
    fi = rng.standard_normal(164); fq = rng.standard_normal(164)
    C = 3
    x = np.stack([O.synth_iq(3000, 40 + c, 1.0) for c in range(C)])
    bf = BatchFilter("cRxFilter", C, (fi, fq))
    y, _ = _run_batch(torch, bf, x, [1000, 1, 1999], 3000)
    bd = BatchFilter("dRxFilter", C, fi)
    yd, _ = _run_batch(torch, bd, x, [1000, 1, 1999], 3000)
    for c in range(C):
        assert O.rel_rms(y[c], O.RxFilterC(fi, fq)(x[c])) < 1e-13
        assert O.rel_rms(yd[c], O.RxFilterD(fi)(x[c])) < 1e-13


def _run_chain(torch, rx, x, splits, want_decim=False):
    C = x.shape[0]
    d_in = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    aud, dec, ca, cdm, pos = [], [], [], [], 0
    for n in splits:
        blk = d_in[:, pos:pos + n].contiguous(); pos += n
        cap = rx.max_out(n)
        a = torch.zeros((C, cap), dtype=torch.float64, device="cuda")
        d = torch.zeros((C, n + 8), dtype=torch.complex128, device="cuda")
        na, nd = rx.process(blk.data_ptr(), max(n, 1), n, a.data_ptr(), cap, d.data_ptr() if want_decim else 0, n + 8)
        torch.cuda.synchronize()
        aud.append(a[:, :na].cpu().numpy()); ca.append(na)
        dec.append(d[:, :nd].cpu().numpy()); cdm.append(nd)
    return np.concatenate(aud, axis=1), ca, np.concatenate(dec, axis=1), cdm


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("rate", RATES)
def test_rx_decimate_kat(rate, fused, torch, tabs):
    """quisk_process_decimate output (tapped through d_decim) against the reference fixture."""
    from quisk_b200.rx import RxChain
    kat = golden("chain_kat.npz")
    fi, fq = demod_taps("USB")
    rx = RxChain(2, rate, "USB", fi, fq, tabs, fused=bool(fused))
    x = np.stack([O.synth_iq(40000, 9, 1.0)] * 2)
    _, _, dec, cdm = _run_chain(torch, rx, x, CHAIN_SPLITS, want_decim=True)
    assert cdm == kat["decimate_%d/counts" % rate].tolist()
    assert rx.decim_srate == int(kat["decimate_%d/srate" % rate][0])
    for c in range(2):
        assert O.rel_rms(dec[c], kat["decimate_%d/y" % rate]) < 1e-12
    rx.close()


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("mode", list(DEMOD_TAPS))
def test_rx_demod_kat(mode, fused, torch, tabs):
    """quisk_process_demodulate at 48 kS/s in (no decimation planned) against the reference fixture."""
    from quisk_b200.rx import RxChain
    kat = golden("chain_kat.npz")
    fi, fq = demod_taps(mode)
    rx = RxChain(3, 48000, mode, fi, fq, tabs, fused=bool(fused))      # fused: rxtail.cu for SSB / CW
    x = np.stack([O.synth_iq(12000, 10, 1.0)] * 3)
    aud, ca, _, _ = _run_chain(torch, rx, x, DEMOD_SPLITS)
    assert ca == kat["demod_%s/counts" % mode].tolist()
    errs = [O.rel_rms(aud[c], kat["demod_%s/y" % mode]) for c in range(3)]
    print("demod", mode, fused, errs)
    assert max(errs) < 1e-12
    rx.close()


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("case", DGT_CASES, ids=[c[0] for c in DGT_CASES])
def test_rx_demod_digital_modes_kat(case, fused, torch, tabs):
    """DGT-U/L, FDV (narrow: CW's structure; wide: I/Q filter at 48 k) and DGT-IQ (complex out), quisk.c:2087-2153."""
    from quisk_b200.rx import RxChain
    name, mode, ntap, bw = case
    kat = golden("chain_kat.npz")
    rng = np.random.default_rng(3)
    fi, fq = rng.standard_normal(ntap) / ntap, rng.standard_normal(ntap) / ntap
    C = 3
    rx = RxChain(C, 48000, mode, fi, fq, tabs, fused=bool(fused), bandwidth=bw)
    x = np.stack([O.synth_iq(12000, 10, 1.0)] * C)
    d_in = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    outs, counts, pos = [], [], 0
    for n in DEMOD_SPLITS:
        blk = d_in[:, pos:pos + n].contiguous(); pos += n
        cap = rx.max_out(n)
        a = torch.zeros((C, cap), dtype=torch.float64, device="cuda")
        na, _ = rx.process(blk.data_ptr(), n, n, a.data_ptr(), cap)
        torch.cuda.synchronize()
        got = a.cpu().numpy()
        outs.append(got[:, :2 * na].view(np.complex128) if mode == "DGT-IQ" else got[:, :na]); counts.append(na)
    assert counts == kat["demod_%s/counts" % name].tolist()
    y = np.concatenate(outs, axis=1)
    for c in range(C):
        assert O.rel_rms(y[c], kat["demod_%s/y" % name]) < 1e-12
    rx.close()


@pytest.mark.parametrize("fused", [0, 1])
@pytest.mark.parametrize("tune", [0, 12345])
def test_c1_chain_kat(tune, fused, torch, tabs):
    """BASELINE.json configs[0] through the batched chain, 10 ms blocks, vs the reference fixture."""
    from quisk_b200.rx import RxChain
    kat = golden("chain_kat.npz")
    fi, fq = kat["c1/filt_i"], kat["c1/filt_q"]
    C = 4
    rx = RxChain(C, 1536000, "USB", fi, fq, tabs, tune_hz=[float(tune)] * C, fused=bool(fused))
    x = np.stack([O.synth_iq(153600, 20, 1.0)] * C)
    aud, ca, _, _ = _run_chain(torch, rx, x, [15360] * 10)
    assert ca == kat["c1_tune%d/counts" % tune].tolist()
    for c in range(C):
        assert O.rel_rms(aud[c], kat["c1_tune%d/y" % tune]) < 1e-12
    # block-split invariance: one 153 600-sample call gives the same stream
    rx.reset()
    aud1, ca1, _, _ = _run_chain(torch, rx, x, [153600])
    assert ca1 == [4800]
    assert O.rel_rms(aud1[0], kat["c1_tune%d/y" % tune]) < 1e-12
    rx.close()


@pytest.mark.parametrize("fused", [0, 1])
def test_c1_chain_ragged_blocks_tuned(fused, torch, tabs):
    """Ragged block lengths (1, 2, 3, 7, ... samples, shorter than any stage's history, shorter than a fused chunk)
    with tuning on: every phase, history and the block-start tuning phasor carry over, and the concatenated audio is
    the fixture's stream."""
    from quisk_b200.rx import RxChain
    kat = golden("chain_kat.npz")
    fi, fq = kat["c1/filt_i"], kat["c1/filt_q"]
    C = 2
    rx = RxChain(C, 1536000, "USB", fi, fq, tabs, tune_hz=[12345.0] * C, fused=bool(fused))
    x = np.stack([O.synth_iq(153600, 20, 1.0)] * C)
    splits = [1, 2, 3, 7, 255, 256, 257, 1000, 4093, 15360, 18766, 2049, 31, 0, 33]
    splits.append(153600 - sum(splits))
    aud, ca, _, _ = _run_chain(torch, rx, x, splits)
    assert sum(ca) == 4800
    for c in range(C):
        assert O.rel_rms(aud[c], kat["c1_tune12345/y"]) < 1e-12
    rx.close()


FUSED_VARIANTS = [("tailwarp", 12, 0), ("tailwarp", 12, 2), ("tailwarp", 12, 3), ("tailwarp", 12, 4), ("split", 11, 1),
                  ("async", 19, 0), ("async", 19, 1), ("async", 19, 2), ("async_twsplit", 11, 1), ("async_twsplit0", 11, 3), ("async_p3", 20, 1), ("tw4split", 12, 4)]     # async: on the default tail-warp kernel


@pytest.mark.parametrize("variant", FUSED_VARIANTS, ids=["%s%d" % (v[0], v[2]) for v in FUSED_VARIANTS])
@pytest.mark.parametrize("mode", ["USB", "CWU", "AM"])
def test_fused_kernel_variants(mode, variant, torch, tabs):
    """Every variant of the fused decimator (all stages on the main warps; the stages from index 2, 3 or 4 on the two
    tail warps; component-split half bands; the next chunk by cp.async instead of a register prefetch) against the reference fixture, over full chunks, blocks that end inside a
    chunk, blocks shorter than a chunk and odd / even numbers of full chunks (the tail warps' double buffer ends on
    either half) -- histories, phases and the tuning phasor carry over between all of them."""
    from quisk_b200.rx import RxChain
    kat = golden("chain_kat.npz")
    C = 3
    if mode == "USB":
        fi, fq = kat["c1/filt_i"], kat["c1/filt_q"]
        ref = kat["c1_tune12345/y"]
        tune = 12345.0
    else:
        fi, fq = demod_taps(mode)
        ref = None
        tune = 0.0
    rx = RxChain(C, 1536000, mode, fi, fq, tabs, tune_hz=[tune] * C, fused=True)
    rx.set_option(12, 0); rx.set_option(11, 0)
    x = np.stack([O.synth_iq(153600, 20 + (c if mode != "USB" else 0), 1.0) for c in range(C)])
    splits = [2048, 4096, 6144, 2047, 2049, 8192 + 5, 100, 3 * 2048, 1, 10240, 30720, 7]
    splits.append(153600 - sum(splits))
    base, cb, _, _ = _run_chain(torch, rx, x, splits)
    rx.reset()
    if variant[0].startswith("async"):
        rx.set_option(12, 1)
    if variant[0] == "tw4split":
        rx.set_option(11, 3)
    rx.set_option(variant[1], variant[2])
    aud, ca, _, _ = _run_chain(torch, rx, x, splits)
    rx.close()
    assert ca == cb
    # same arithmetic in the same order: the variants agree bit for bit
    assert np.array_equal(aud, base)
    if ref is not None and mode == "USB":
        for c in range(C):
            assert O.rel_rms(aud[c], ref) < 1e-12


def test_fused_192k_plan_split_equals_complex_lanes(torch, tabs):
    """The four-stage plan of the 192 kS/s chain (HB45, FIR 98/2, HB45, FIR 98/2: the north star's target rate) runs with
    component-split half bands by default; the complex-lane form of the same plan must agree bit for bit."""
    from quisk_b200.rx import RxChain
    kat = golden("chain_kat.npz")
    fi, fq = kat["c1/filt_i"], kat["c1/filt_q"]
    C = 3
    x = np.stack([O.synth_iq(40960, 30 + c, 1.0) for c in range(C)])
    splits = [2048, 4096, 2047, 2049, 8192 + 5, 100, 1, 10240]
    splits.append(40960 - sum(splits))
    res = []
    for split in (0, 3):
        rx = RxChain(C, 192000, "USB", fi, fq, tabs, tune_hz=[4321.0] * C, fused=True)
        rx.set_option(11, split)
        res.append(_run_chain(torch, rx, x, splits)[:2])
        name = rx.lib.quisk_cuda_rx_fused_kernel_name(rx.h).decode()
        assert ("282" in name) == (split == 3), name
        rx.close()
    assert res[0][1] == res[1][1]
    assert np.array_equal(res[0][0], res[1][0])


def test_c1_chain_closed_form_nco(torch, tabs):
    """QC_RX_OPT_EXACT_NCO = 0: block-start phasors from the closed form instead of the reference's recurrence;
    over the 0.1 s fixture both are far inside the tolerance (tests/test_c1_fullsize_gpu.py shows where they part)."""
    from quisk_b200.rx import RxChain
    kat = golden("chain_kat.npz")
    fi, fq = kat["c1/filt_i"], kat["c1/filt_q"]
    C = 2
    rx = RxChain(C, 1536000, "USB", fi, fq, tabs, tune_hz=[12345.0] * C, fused=True)
    rx.set_option(10, 0)
    x = np.stack([O.synth_iq(153600, 20, 1.0)] * C)
    aud, ca, _, _ = _run_chain(torch, rx, x, [15360] * 10)
    assert ca == kat["c1_tune12345/counts"].tolist()
    for c in range(C):
        assert O.rel_rms(aud[c], kat["c1_tune12345/y"]) < 1e-12
    rx.close()


def test_rx_process_host(torch, tabs):
    from quisk_b200.rx import RxChain
    kat = golden("chain_kat.npz")
    fi, fq = kat["c1/filt_i"], kat["c1/filt_q"]
    C = 3
    rx = RxChain(C, 1536000, "USB", fi, fq, tabs, fused=True)
    x = np.stack([O.synth_iq(153600, 20, 1.0)] * C)
    outs = []
    for b in range(10):
        blk = np.ascontiguousarray(x[:, b * 15360:(b + 1) * 15360])
        a = np.zeros((C, rx.max_out(15360)))
        na = rx.process_host(blk, 15360, a)
        outs.append(a[:, :na])
    aud = np.concatenate(outs, axis=1)
    for c in range(C):
        assert O.rel_rms(aud[c], kat["c1_tune0/y"]) < 1e-12
    rx.close()


@pytest.mark.parametrize("packed", [False, True])
def test_rx_process_host_noise_blanker(packed, torch, tabs):
    """QC_RX_OPT_NOISE_BLANKER (13): the host entries run NoiseBlanker (quisk.c:679-784) on the staged block in front of
    the tuning stage, as quisk_process_samples does (quisk.c:2448-2449).  The blanker is bit-exact, so a chain with the
    option fed the raw stream must give exactly the audio of a plain chain fed the oracle-blanked stream, block by block."""
    from quisk_b200.rx import RxChain
    kat = golden("chain_kat.npz")
    fi, fq = kat["c1/filt_i"], kat["c1/filt_q"]
    C, n, nblk, rate = 2, 15360, 4, 1536000
    if packed:      # int16 wire format: integers survive the unpack exactly, so the two feeds stay bit-identical
        rng = np.random.default_rng(41)
        xi = rng.integers(-2000, 2000, size=(n * nblk, 2)).astype(np.int16)
        xi[[5000, 5001, 20000, 33000], :] = [[30000, -30000]]
        x = O.unpack_iq(xi.reshape(-1).view(np.uint8), 2, False)          # what add_rx_samples makes of these bytes
    else:
        x = O.synth_iq(n * nblk, 21, 1.0)
        x[[5000, 20000, 20003, 33000]] *= 70.0
    a_nb = RxChain(C, rate, "USB", fi, fq, tabs, tune_hz=[5000.0] * C, fused=True)
    a_nb.set_option(13, 2)
    a_ref = RxChain(C, rate, "USB", fi, fq, tabs, tune_hz=[5000.0] * C, fused=True)
    blanker = O.NoiseBlanker(rate, 2)
    for b in range(nblk):
        sl = slice(b * n, (b + 1) * n)
        y0 = np.zeros((C, a_nb.max_out(n))); y1 = np.zeros_like(y0)
        clean = np.ascontiguousarray(np.stack([blanker(x[sl])] * C))
        if packed:
            raw = np.ascontiguousarray(np.stack([xi[sl].reshape(-1).view(np.uint8)] * C))
            n0 = a_nb.process_host_packed(raw, n, 2, False, y0)
        else:
            n0 = a_nb.process_host(np.ascontiguousarray(np.stack([x[sl]] * C)), n, y0)
        n1 = a_ref.process_host(clean, n, y1)
        assert n0 == n1 > 0
        assert np.array_equal(y0[:, :n0], y1[:, :n1])
    a_nb.close(); a_ref.close()


@pytest.mark.parametrize("packed", [False, True])
def test_rx_process_host_pipelined_chunks(packed, torch, tabs):
    """QC_RX_OPT_HOST_CHUNKS (15): the host entries split the channels into chunks whose H2D copy, kernels and D2H copy
    overlap on streams of their own.  Each chunk is a chain over a channel subset, so the audio must be the same bits as
    the single-sequence entry, call after call (state carries inside the chunks), for per-channel tuning and a channel
    count the chunks do not divide."""
    from quisk_b200.rx import RxChain
    kat = golden("chain_kat.npz")
    fi, fq = kat["c1/filt_i"], kat["c1/filt_q"]
    C, n, nblk, rate = 11, 15360, 3, 1536000
    tune = [1000.0 * (c + 1) for c in range(C)]
    if packed:
        rng = np.random.default_rng(43)
        xi = rng.integers(-20000, 20000, size=(C, n * nblk, 2)).astype(np.int16)
    else:
        x = np.stack([O.synth_iq(n * nblk, 30 + c, 1.0) for c in range(C)])
    res = {}
    for chunks in (1, 4):
        rx = RxChain(C, rate, "USB", fi, fq, tabs, tune_hz=tune, fused=True)
        rx.set_option(15, chunks)
        outs = []
        for b in range(nblk):
            a = np.zeros((C, rx.max_out(n)))
            if packed:
                raw = np.ascontiguousarray(xi[:, b * n:(b + 1) * n].reshape(C, -1).view(np.uint8))
                na = rx.process_host_packed(raw, n, 2, False, a)
            else:
                na = rx.process_host(np.ascontiguousarray(x[:, b * n:(b + 1) * n]), n, a)
            outs.append(a[:, :na].copy())
        res[chunks] = np.concatenate(outs, axis=1)
        rx.close()
    assert res[1].shape[1] == nblk * n // 32 and np.abs(res[1]).max() > 0
    assert np.array_equal(res[1], res[4])
    assert not np.array_equal(res[1][0], res[1][1])          # channels really differ (own tuning, own input)
