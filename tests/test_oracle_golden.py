"""The NumPy oracle against the committed fixtures (generated from the compiled reference
by tests/golden/make_golden.py).  Runs everywhere -- no reference, no GPU needed."""
import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.golden.make_golden import FILTER_CASES, RATES, DEMOD_TAPS, DGT_CASES, kat_input, demod_taps
from tests.util import SPLITS, CHAIN_SPLITS, DEMOD_SPLITS, golden

TOL = 1e-13


@pytest.fixture(scope="module")
def tabs():
    return golden("quisk_tables.npz")


def _oracle_stage(fn, tab, args, tune, real):
    dt = np.float64 if real else np.complex128
    if fn == "quisk_cDecim2HB45": return O.HB45Decim()
    if fn in ("quisk_cInterp2HB45", "quisk_dInterp2HB45"): return O.HB45Interp(dt)
    if fn in ("quisk_cDecimate", "quisk_dDecimate"): return O.FirDecim(tab, args[0], dt)
    if fn in ("quisk_cFilter", "quisk_dFilter"): return O.FirDecim(tab, 1, dt)
    if fn == "quisk_cCDecimate":
        D = (len(tab) - 1.0) / 2.0
        hc = np.exp(2j * np.pi * tune[0] * (np.arange(len(tab)) - D)) * tab
        if not tune[1]: hc = hc.imag + 1j * hc.real          # filter.c:76-79
        return O.FirDecim(hc, args[0])
    if fn in ("quisk_cInterpolate", "quisk_dInterpolate"): return O.FirInterp(tab, args[0], dt)
    if fn == "quisk_cInterpDecim": return O.FirInterpDecim(tab, args[0], args[1])
    raise KeyError(fn)


@pytest.mark.parametrize("case", FILTER_CASES, ids=[c[0] for c in FILTER_CASES])
def test_filter_kat(case, tabs):
    name, fn, seed, real, tab, args, tune = case
    kat = golden("filter_kat.npz")
    x = kat_input(seed, real)
    st = _oracle_stage(fn, tabs[tab] if tab else None, args, tune, real)
    outs, counts, pos = [], [], 0
    for n in SPLITS:
        y = st(x[pos:pos + n]); pos += n
        outs.append(y); counts.append(len(y))
    assert counts == kat[name + "/counts"].tolist()
    assert O.rel_rms(np.concatenate(outs), kat[name + "/y"]) < TOL


@pytest.mark.parametrize("rate", RATES)
def test_decimate_kat(rate, tabs):
    kat = golden("chain_kat.npz")
    chain = O.ProcessDecimate(rate, tabs)
    x = O.synth_iq(40000, 9, 1.0)
    outs, counts, pos = [], [], 0
    for n in CHAIN_SPLITS:
        y = chain(x[pos:pos + n]); pos += n
        outs.append(y); counts.append(len(y))
    assert counts == kat["decimate_%d/counts" % rate].tolist()
    assert chain.decim_srate == int(kat["decimate_%d/srate" % rate][0])
    assert O.rel_rms(np.concatenate(outs), kat["decimate_%d/y" % rate]) < TOL


@pytest.mark.parametrize("mode", list(DEMOD_TAPS))
def test_demod_kat(mode, tabs):
    kat = golden("chain_kat.npz")
    fi, fq = demod_taps(mode)
    chain = O.ProcessDemodulate(mode, fi, fq, tabs)
    x = O.synth_iq(12000, 10, 1.0)
    outs, counts, pos = [], [], 0
    for n in DEMOD_SPLITS:
        y = chain(x[pos:pos + n]); pos += n
        outs.append(y); counts.append(len(y))
    assert counts == kat["demod_%s/counts" % mode].tolist()
    assert O.rel_rms(np.concatenate(outs), kat["demod_%s/y" % mode]) < (1e-11 if mode == "FM" else TOL)


def dgt_taps(ntap):
    rng = np.random.default_rng(3)
    return rng.standard_normal(ntap) / ntap, rng.standard_normal(ntap) / ntap


@pytest.mark.parametrize("case", DGT_CASES, ids=[c[0] for c in DGT_CASES])
def test_demod_digital_modes_kat(case, tabs):
    """DGT-U/L, FDV-U/L (narrow and wide) and DGT-IQ (quisk.c:2087-2153) against the compiled reference."""
    name, mode, ntap, bw = case
    kat = golden("chain_kat.npz")
    fi, fq = dgt_taps(ntap)
    chain = O.ProcessDemodulate(mode, fi, fq, tabs, bandwidth=bw)
    x = O.synth_iq(12000, 10, 1.0)
    outs, counts, pos = [], [], 0
    for n in DEMOD_SPLITS:
        y = chain(x[pos:pos + n]); pos += n
        outs.append(y); counts.append(len(y))
    assert counts == kat["demod_%s/counts" % name].tolist()
    assert O.rel_rms(np.concatenate(outs), kat["demod_%s/y" % name]) < TOL


@pytest.mark.parametrize("tune", [0, 12345])
def test_c1_chain_kat(tune, tabs):
    """BASELINE.json configs[0]: 1.536 MS/s -> 4 x HB45 -> 98-tap /2 -> 48 k -> USB (164-tap I/Q)."""
    kat = golden("chain_kat.npz")
    protos = {int(k[6:]): v for k, v in tabs.items() if k.startswith("proto_")}
    fi, fq = O.make_filter_coef(12000, None, 2800, 300 + 2800 // 2, protos)
    assert np.array_equal(fi, kat["c1/filt_i"]) and np.array_equal(fq, kat["c1/filt_q"]) and len(fi) == 164
    dec = O.ProcessDecimate(1536000, tabs)
    dem = O.ProcessDemodulate("USB", fi, fq, tabs)
    nco = O.TuneNCO(float(tune), 1536000)
    x = O.synth_iq(153600, 20, 1.0)
    outs, counts = [], []
    for b in range(10):
        blk = x[b * 15360:(b + 1) * 15360]
        y = dem(dec(nco(blk)))
        outs.append(y); counts.append(len(y))
    assert counts == kat["c1_tune%d/counts" % tune].tolist() == [480] * 10
    assert O.rel_rms(np.concatenate(outs), kat["c1_tune%d/y" % tune]) < TOL


@pytest.mark.parametrize("big", [0, 1])
@pytest.mark.parametrize("nb", [1, 2, 3, 4])
def test_unpack_iq_kat(nb, big):
    """add_rx_samples' unpack loops (quisk.c:2922-2953), incl. the float rounding of 4-byte samples."""
    from tests.golden.make_golden_misc import ingest_bytes
    kat = golden("misc_kat.npz")
    y = O.unpack_iq(ingest_bytes(70 + nb, 1000 * 2 * nb), nb, bool(big))
    assert np.array_equal(y, kat["unpack_iq_%d_%d/y" % (nb, big)])


@pytest.mark.parametrize("n_rx", [1, 2, 4, 10])
def test_unpack_hermes_kat(n_rx):
    """read_rx_udp10's 24-bit record loop (quisk.c:3746-3763)."""
    from tests.golden.make_golden_misc import ingest_bytes
    kat = golden("misc_kat.npz")
    pk = ingest_bytes(80 + n_rx, 3 * 1032).reshape(3, 1032)
    y = np.concatenate([O.unpack_hermes(pk[p], n_rx) for p in range(3)], axis=1)
    assert np.array_equal(y, kat["unpack_hermes_%d/y" % n_rx])


@pytest.mark.parametrize("rate,level", [(48000, 1), (48000, 3), (192000, 2), (1536000, 1)])
def test_noise_blanker_kat(rate, level):
    """NoiseBlanker (quisk.c:679-784): bit-exact against the compiled reference's output, ragged blocks."""
    from tests.golden.make_golden_misc import NB_SPLITS, nb_input
    kat = golden("misc_kat.npz")
    x = nb_input(sum(NB_SPLITS), 90)
    nb = O.NoiseBlanker(rate, level)
    ys, pos = [], 0
    for n in NB_SPLITS:
        ys.append(nb(x[pos:pos + n])); pos += n
    y = np.concatenate(ys)
    ref = kat["nb_%d_%d/y" % (rate, level)]
    assert np.array_equal(y, ref)
    assert (ref == 0).sum() > 3 * int(rate * 500.0e-6 + 0.5)        # the fixture does blank something beyond the initial delay


@pytest.mark.parametrize("level", [150, 60])
def test_ssb_squelch_kat(level):
    """ssb_squelch + d_delay (quisk.c:1056-1180): timer and squelch_active after every call equal the compiled
    reference's, the delayed audio is bit-exact."""
    from tests.golden.make_golden_misc import SQ_BW, SQ_RATE, SQ_SPLITS, sq_input
    kat = golden("misc_kat.npz")
    x = sq_input(sum(SQ_SPLITS), 91)
    sq = O.SsbSquelch(SQ_RATE, SQ_BW, level)
    ys, act, opn, pos = [], [], [], 0
    for n in SQ_SPLITS:
        ys.append(sq(x[pos:pos + n])); pos += n
        act.append(sq.active); opn.append(sq.sq_open)
    assert act == kat["sq_%d/active" % level].tolist()
    assert opn == kat["sq_%d/sq_open" % level].tolist()
    assert np.array_equal(np.concatenate(ys), kat["sq_%d/y" % level])
    assert 0 < sum(act) < len(act)                  # the fixture sees the squelch both closed and open


@pytest.mark.parametrize("sidetone", [0, 700])
def test_auto_notch_kat(sidetone):
    """dAutoNotch (quisk.c:786-963): the frame-by-frame restatement (numpy's FFT) against the compiled reference
    (the oracle's FFTW shim) on a stream that goes from no notch to one to two."""
    from tests.golden.make_golden_misc import AN_RATE, AN_SPLITS, an_input
    kat = golden("misc_kat.npz")
    x = an_input(sum(AN_SPLITS), 92)
    an = O.AutoNotch(AN_RATE, sidetone)
    ys, pos = [], 0
    for n in AN_SPLITS:
        ys.append(an(x[pos:pos + n])); pos += n
    y, ref = np.concatenate(ys), kat["notch_%d/y" % sidetone]
    assert O.rel_rms(y, ref) < 1e-12
    for lo in range(0, len(ref) - 2048, 2048):
        assert O.rel_rms(y[lo:lo + 2048], ref[lo:lo + 2048]) < 1e-11
