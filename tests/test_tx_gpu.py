"""GPU parity of the transmit-audio chain (the TX mirror): quisk_cuda_tx_filter_* against fixtures from the
compiled reference's own tx_filter / CcmPeak (microphone.c:372-604, 161-233; tests/golden/make_golden_tx.py), in the four
modes tx_filter distinguishes, over ragged blocks (1, 5, 6, 7 ... 12000 samples; blocks shorter than the decimation),
with pre-emphasis, compression, the hard limiter and the peak rounder engaged -- including the reference's quirk that
CcmPeak only initialises on its first call.  The FIR stages are bit-exact kernels; the two gain recurrences divide by
quantities built from hypot(), so the bound is 1e-12 of the output's RMS (measured: see the printed values)."""
import os

import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.golden.make_golden_tx import ALC_IN_SCALE, ALC_KEY_DOWN_AT, ALC_SPLITS, CLIP, MIC_RATE, PREEMPH, TX_SPLITS, mic_audio
from tests.util import golden

pytestmark = pytest.mark.gpu
NCH = 3


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def tabs():
    from quisk_b200.rx import load_tables
    return load_tables()


@pytest.mark.parametrize("mode", ["LSB", "USB", "AM", "FM"])
def test_tx_filter_parity(mode, torch, tabs):
    from quisk_b200.rx import TxFilter
    kat = golden("tx_kat.npz")
    key = "AM" if mode == "FM" else mode          # tx_filter only asks is_ssb: AM and FM are the same branch (the generator asserts it)
    ref, counts_ref = kat["tx_%s/y" % key], kat["tx_%s/counts" % key].tolist()
    x = mic_audio()
    tx = TxFilter(NCH, mode, tabs, mic_sample_rate=MIC_RATE, preemphasis=PREEMPH, clip=CLIP)
    scale = np.array([1.0, 1.0, 0.5])              # the third transmitter speaks half as loud: its own gain history
    d = torch.from_numpy(np.ascontiguousarray(np.stack([x * s for s in scale]).astype(np.complex128))).cuda()
    outs, counts, pos = [], [], 0
    for n in TX_SPLITS:
        blk = d[:, pos:pos + n].contiguous(); pos += n
        out = torch.zeros((NCH, tx.max_out(n) + 8), dtype=torch.complex128, device="cuda")
        no = tx.process(blk.data_ptr(), blk.stride(0), n, out.data_ptr(), out.stride(0))
        torch.cuda.synchronize()
        outs.append(out[:, :no].cpu().numpy()); counts.append(no)
    tx.close()
    y = np.concatenate(outs, axis=1)
    assert counts == counts_ref
    errs = [O.rel_rms(y[c], ref) for c in range(2)]
    print("tx_filter", mode, "rel rms vs compiled reference", errs, "peak", np.abs(ref).max())
    assert max(errs) < 1e-12
    if key == "AM":
        assert not y.imag.any()
    # the quieter transmitter is a different stream (the normaliser is not linear), but it must be a valid one
    assert np.isfinite(y[2]).all() and np.abs(y[2]).max() > 1000.0 and O.rel_rms(y[2], ref) > 1e-3


@pytest.mark.parametrize("mode", ["DGT-U", "DGT-L", "FDV-U", "FDV-L"])
def test_tx_filter_digital_parity(mode, torch, tabs):
    """tx_filter_digital (microphone.c:605-624): quisk_dC_out on the 520 tuned taps, times two -- exact kernels: bit for bit."""
    from quisk_b200.rx import TxFilter
    kat = golden("tx_kat.npz")
    ref = kat["txd_%s/y" % {"FDV-U": "DGT-U", "FDV-L": "DGT-L"}.get(mode, mode)]        # FDV tunes like DGT (microphone.c:617)
    x = mic_audio()
    tx = TxFilter(2, mode, tabs, mic_sample_rate=48000)
    d = torch.from_numpy(np.ascontiguousarray(np.stack([x + 0.25j * x, x]).astype(np.complex128))).cuda()      # the imaginary rail is ignored
    outs, pos = [], 0
    for n in TX_SPLITS[:6]:
        blk = d[:, pos:pos + n].contiguous(); pos += n
        out = torch.zeros((2, tx.max_out(n) + 8), dtype=torch.complex128, device="cuda")
        assert tx.process(blk.data_ptr(), blk.stride(0), n, out.data_ptr(), out.stride(0)) == n
        outs.append(out[:, :n].cpu().numpy())
    tx.close()
    y = np.concatenate(outs, axis=1)
    assert np.array_equal(y[0], ref) and np.array_equal(y[1], ref)


@pytest.mark.parametrize("mode", ["USB", "DGT-U"])
def test_tx_filter_with_alc(mode, torch, tabs):
    """tx_filter / tx_filter_digital followed by process_alc (microphone.c:1232-1233, 270-370), with a key down in the middle
    (init_alc(&tx_alc, 0): delay line and ramp cleared, gain kept), against the same chain of the compiled reference."""
    from quisk_b200.rx import TxFilter
    kat = golden("tx_kat.npz")
    ref = kat["alc_%s/y" % mode]
    x = mic_audio() * ALC_IN_SCALE[mode]
    tx = TxFilter(2, mode, tabs, mic_sample_rate=MIC_RATE, preemphasis=PREEMPH, clip=CLIP)
    tx.set_alc(1)
    d = torch.from_numpy(np.ascontiguousarray(np.stack([x, x]).astype(np.complex128))).cuda()
    outs, pos = [], 0
    for k, n in enumerate(ALC_SPLITS):
        if k == ALC_KEY_DOWN_AT:
            tx.set_alc(1)
        blk = d[:, pos:pos + n].contiguous(); pos += n
        out = torch.zeros((2, tx.max_out(n) + 8), dtype=torch.complex128, device="cuda")
        no = tx.process(blk.data_ptr(), blk.stride(0), n, out.data_ptr(), out.stride(0))
        outs.append(out[:, :no].cpu().numpy())
    tx.close()
    y = np.concatenate(outs, axis=1)
    assert y.shape[1] == len(ref)
    errs = [O.rel_rms(y[c], ref) for c in range(2)]
    print("tx + alc", mode, errs, "peak", np.abs(ref).max(), "(CLIP16 - 10 = 32757)")
    assert np.abs(ref).max() > 32756.0           # the control is working against the limit
    assert max(errs) < 1e-12


def test_tx_filter_rejects_what_tx_filter_does_not_serve(torch, tabs):
    from quisk_b200 import lib as L
    from quisk_b200.rx import TxFilter
    with pytest.raises(L.QuiskCudaError):
        TxFilter(1, "CWU", tabs)
    with pytest.raises(L.QuiskCudaError):
        TxFilter(1, "USB", tabs, mic_sample_rate=44100)
