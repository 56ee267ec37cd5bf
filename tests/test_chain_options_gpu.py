"""The optional stages inside quisk_process_demodulate as options of the batched chain (QC_RX_OPT_AUTO_NOTCH, _NOTCH_SIDETONE,
_SSB_SQUELCH): dAutoNotch, then ssb_squelch + d_delay, on the audio at the filter rate between the detector and the audio
interpolators, and the muting of a squelched receiver's block -- against fixtures from the compiled reference's own
quisk_process_demodulate with quisk_auto_notch / ssb_squelch_enabled set (tests/golden/make_golden_chain_options.py)."""
import ctypes as C

import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.golden.make_golden import demod_taps
from tests.golden.make_golden_chain_options import BLOCK, BLOCKS, CASES, KEEP, RATE, options_input
from tests.util import golden

pytestmark = pytest.mark.gpu
NCH = 3


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.mark.parametrize("mode,notch,level,rit", CASES)
def test_chain_with_notch_and_squelch(mode, notch, level, rit, torch):
    from quisk_b200.rx import RxChain, load_tables
    kat = golden("chain_options_kat.npz")
    key = "%s_n%d_s%d" % (mode, notch, level)
    fi, fq = demod_taps(mode)
    rx = RxChain(NCH, RATE, mode, fi, fq, load_tables(), tune_hz=None, fused=True, bandwidth=2800 if mode == "USB" else 500)
    rx.set_option(16, notch); rx.set_option(17, rit); rx.set_option(18, level)
    x = options_input(31)
    dev = torch.from_numpy(np.ascontiguousarray(np.stack([x] * NCH))).cuda()
    aud = torch.zeros((NCH, BLOCK + 64), dtype=torch.float64, device="cuda")
    ref_y, ref_act, ref_rms = kat[key + "/y"], kat[key + "/active"], kat[key + "/rms"]
    peak = np.abs(ref_y).max()
    act = np.zeros(NCH, dtype=np.int32)
    worst, kept = 0.0, 0
    for b in range(BLOCKS):
        blk = dev[:, b * BLOCK:(b + 1) * BLOCK]
        n = rx.process(blk.data_ptr(), dev.stride(0), BLOCK, aud.data_ptr(), aud.stride(0))[0]
        torch.cuda.synchronize()
        assert n == BLOCK
        assert rx.lib.quisk_cuda_rx_squelch_active(rx.h, act.ctypes.data_as(C.c_void_p)) == 0
        assert (act == ref_act[b]).all(), (b, act, ref_act[b])
        y = aud[:, :n].cpu().numpy()
        for c in range(NCH):
            rms = np.sqrt(np.mean(y[c] * y[c]))
            assert abs(rms - ref_rms[b]) <= 1e-9 * peak, (b, rms, ref_rms[b])
            if ref_act[b]:
                assert not y[c].any()                       # muted as quisk_process_samples mutes it
        if b in KEEP:
            r = ref_y[KEEP.index(b)]
            for c in range(NCH):
                worst = max(worst, np.abs(y[c] - r).max() / peak)
            kept += 1
    rx.close()
    print(key, "max |diff| / peak over", kept, "kept blocks:", worst, " squelched blocks:", int(ref_act.sum()))
    assert kept == len(KEEP)
    assert worst < 1e-12
    if level:
        assert 5 < ref_act.sum() < BLOCKS - 5               # the squelch really closed and opened
    if notch:
        # the notch found the carrier: with it, the blocks that hold only the carrier are far below the same blocks without it
        plain = kat["USB_n0_s150/rms"]
        if mode == "USB":
            assert ref_rms[36] < 0.02 * plain[36]


def test_options_leave_the_fused_tail_alone_when_off(torch):
    """with all three options at 0 the chain still runs the fused tail kernel (same launches as before)"""
    from quisk_b200.rx import RxChain, load_tables
    fi, fq = demod_taps("USB")
    x = O.synth_iq(8 * BLOCK, 5, 1.0)
    dev = torch.from_numpy(np.ascontiguousarray(np.stack([x] * 2))).cuda()
    outs = []
    for opts in ((), ((16, 0), (18, 0))):
        rx = RxChain(2, RATE, "USB", fi, fq, load_tables(), tune_hz=None, fused=True)
        for o, v in opts:
            rx.set_option(o, v)
        aud = torch.zeros((2, 8 * BLOCK + 64), dtype=torch.float64, device="cuda")
        n = rx.process(dev.data_ptr(), dev.stride(0), 8 * BLOCK, aud.data_ptr(), aud.stride(0))[0]
        torch.cuda.synchronize()
        outs.append(aud[:, :n].cpu().numpy())
        rx.close()
    assert np.array_equal(outs[0], outs[1])
