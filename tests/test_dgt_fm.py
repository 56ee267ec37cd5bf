"""DGT-FM (rx_mode 13): quisk_process_demodulate has ONE case label for FM and DGT_FM (quisk.c:2026-2027), the two modes
differ only in where quisk_process_samples sends the audio afterwards (quisk.c:2633).  The CPU test pins that on the
compiled reference itself (mode 13 reproduces the committed FM fixture bit for bit); the GPU test runs the library's
DGT-FM chain against its FM chain and the fixture."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import quisk_oracle as O          # noqa: E402
from oracle import ref_ctypes as R            # noqa: E402
from tests.golden.make_golden import DEMOD_SPLITS, demod_taps      # noqa: E402


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name))


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def tabs():
    from quisk_b200.rx import load_tables
    return load_tables()


@pytest.mark.skipif(not R.have_ref("libquisk_rx_ref.so"), reason="compiled reference not built (oracle/build_ref.sh)")
def test_reference_dgt_fm_is_fm():
    kat = golden("chain_kat.npz")
    lib = R.load("libquisk_rx_ref.so", private_copy=True)
    lib.ref_set_sample_rate(48000); lib.ref_init_chain()
    fi, fq = demod_taps("FM")
    fi = np.ascontiguousarray(fi); fq = np.ascontiguousarray(fq)
    lib.ref_set_filters(fi.ctypes.data_as(C.c_void_p), fq.ctypes.data_as(C.c_void_p), len(fi), 2800, 0)
    x = O.synth_iq(12000, 10, 1.0)
    outs, counts, pos = [], [], 0
    for n in DEMOD_SPLITS:
        buf = np.zeros(66000, dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
        dbuf = np.zeros(132000)
        nr = lib.ref_process_demodulate(buf.ctypes.data_as(C.c_void_p), dbuf.ctypes.data_as(C.c_void_p), n, 0, 0, R.MODES["DGT-FM"])
        outs.append(dbuf[:nr].copy()); counts.append(nr)
    assert counts == kat["demod_FM/counts"].tolist()
    assert np.array_equal(np.concatenate(outs), kat["demod_FM/y"])


@pytest.mark.gpu
@pytest.mark.parametrize("rate", [48000, 192000])
def test_rx_dgt_fm_equals_fm(rate, torch, tabs):
    from quisk_b200.rx import RxChain
    from tests.test_batch_gpu import _run_chain
    kat = golden("chain_kat.npz")
    fi, fq = demod_taps("FM")
    n_in = 12000 * (rate // 48000)
    x = np.stack([O.synth_iq(n_in, 10, 1.0)] * 2)
    splits = [s * (rate // 48000) for s in DEMOD_SPLITS]
    res = {}
    for mode in ("FM", "DGT-FM"):
        rx = RxChain(2, rate, mode, fi, fq, tabs, fused=True)
        res[mode] = _run_chain(torch, rx, x, splits)[:2]
        rx.close()
    assert res["FM"][1] == res["DGT-FM"][1]
    assert np.array_equal(res["FM"][0], res["DGT-FM"][0])
    if rate == 48000:
        assert res["DGT-FM"][1] == kat["demod_FM/counts"].tolist()
        for c in range(2):
            assert O.rel_rms(res["DGT-FM"][0][c], kat["demod_FM/y"]) < 1e-12
