"""C5 wideband polyphase channelizer (quisk_b200/csrc/pfb.cu) against the oracle definition of SURVEY.md section 8:
direct-phase mix + the reference's quisk_cDecimate on a subset of receivers."""
import numpy as np
import pytest

from oracle import quisk_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


def proto_taps(K, P, fs=98.304e6):
    """A reference-style prototype: WDSP fir_bandpass (fir.c:187-254), BH7 window, real taps."""
    import ctypes as C
    from quisk_b200 import lib as L
    lib = L.load()
    h = np.zeros(K * P)
    fc = 0.4 * fs / K
    assert lib.quisk_cuda_fir_bandpass(K * P, -fc, fc, fs, 1, 0, 1.0, h.ctypes.data_as(C.POINTER(C.c_double))) == 0
    return h


def run(torch, ch, x, splits, layout=0):
    d = torch.from_numpy(x).cuda()
    K = ch.n_channels
    outs, pos = [], 0
    for n in splits:
        blk = d[pos:pos + n].contiguous(); pos += n
        nf_max = ch.count_out(n)
        if layout == 0:
            o = torch.zeros((K, max(nf_max, 1)), dtype=torch.complex128, device="cuda")
            nf = ch.process(blk.data_ptr() if n else 0, n, o.data_ptr(), max(nf_max, 1), 0)
            outs.append(o[:, :nf].cpu().numpy())
        else:
            o = torch.zeros((max(nf_max, 1), K), dtype=torch.complex128, device="cuda")
            nf = ch.process(blk.data_ptr() if n else 0, n, o.data_ptr(), K, 1)
            outs.append(o[:nf].cpu().numpy().T)
        assert nf == nf_max
    return np.concatenate(outs, axis=1)


CASES = [(1024, 512, 16), (1024, 1024, 8), (512, 256, 16), (256, 128, 32), (256, 100, 4), (1024, 512, 4)]


@pytest.mark.parametrize("K,D,P", CASES)
def test_channelizer_vs_oracle(K, D, P, torch):
    from quisk_b200.rx import Channelizer
    h = proto_taps(K, P)
    n = K * P + 23 * D + 77
    x = O.synth_iq(n, 31, 1.0)
    chans = [0, 1, 2, K // 4 + 3, K // 2, K - 1]
    ref = O.channelizer_oracle(x, chans, h, K, D)
    ch = Channelizer(K, D, h)
    y = run(torch, ch, x, [n])
    assert y.shape[1] == ref.shape[1] == n // D
    scale = np.sqrt(np.mean(np.abs(y) ** 2))               # receivers with no tone in them hold noise only:
    for i, k in enumerate(chans):                          # compare against the band-wide output level
        assert np.sqrt(np.mean(np.abs(y[k] - ref[i]) ** 2)) / scale < 1e-12
    ch.close()


@pytest.mark.parametrize("layout", [0, 1])
def test_channelizer_block_split_invariance(layout, torch):
    """Ragged blocks (including 0, 1 and shorter than one frame) give the single-call stream bit for bit."""
    from quisk_b200.rx import Channelizer
    K, D, P = 1024, 512, 16
    h = proto_taps(K, P)
    n = 40000
    x = O.synth_iq(n, 32, 1.0)
    ch = Channelizer(K, D, h)
    y1 = run(torch, ch, x, [n], layout)
    ch.seek(0)
    splits = [1, 0, 510, 1, 3, 2048, 17000, 511, 513, 9000]
    splits.append(n - sum(splits))
    y2 = run(torch, ch, x, splits, layout)
    assert y1.shape == y2.shape == (K, n // D)
    assert np.array_equal(y1, y2)
    ch.close()


@pytest.mark.parametrize("P,start", [(16, 0), (16, 7 * 512), (8, 3 * 512 + 100)])
def test_channelizer_one_kernel_cluster_path_equals_two_kernel_path(P, start, torch):
    """QC_PFB_OPT_FUSED (option 7): branch FIRs and transforms in one kernel, clusters of 8 CTAs exchanging u through
    distributed shared memory.  Same arithmetic as the two-kernel path: identical bits, from an even or an odd first
    frame, over several frame ranges and ragged blocks."""
    from quisk_b200.rx import Channelizer
    K, D = 1024, 512
    h = proto_taps(K, P)
    n = 700 * D + 333
    x = O.synth_iq(n, 35, 1.0)
    splits = [300 * D + 5, 1, 0, 33 * D - 7]
    splits.append(n - sum(splits))
    ys = []
    for fused in (0, 1):
        ch = Channelizer(K, D, h)
        ch.set_option(7, fused)
        ch.seek(start)
        ys.append(run(torch, ch, x, splits))
        ch.close()
    assert ys[0].shape[1] == (start + n) // D - start // D
    assert np.array_equal(ys[0], ys[1])


def test_channelizer_time_block_shard(torch):
    """A shard that starts mid-stream after seek(t - halo) + prime(halo) reproduces the sequential frames exactly
    (halo = n_taps rounded up to the decimation, block starts multiples of it)."""
    from quisk_b200.rx import Channelizer
    from quisk_b200.shard import time_blocks
    K, D, P = 1024, 512, 16
    h = proto_taps(K, P)
    n = 64 * D
    x = O.synth_iq(n, 33, 1.0)
    ch = Channelizer(K, D, h)
    full = run(torch, ch, x, [n])
    parts = []
    for tb in time_blocks(n, 2, D, K * P - 1):
        ch.seek(tb.halo_start)
        if tb.start > tb.halo_start:
            d = torch.from_numpy(x[tb.halo_start:tb.start].copy()).cuda()
            ch.prime(d.data_ptr(), tb.start - tb.halo_start)
            torch.cuda.synchronize()
        part = run(torch, ch, x[tb.start:tb.stop].copy(), [tb.stop - tb.start])
        assert part.shape[1] == tb.out_count
        parts.append(part)
    assert np.array_equal(np.concatenate(parts, axis=1), full)
    ch.close()


def test_channelizer_rejects_bad_arguments(torch):
    from quisk_b200 import lib as L
    from quisk_b200.rx import Channelizer
    with pytest.raises(L.QuiskCudaError):
        Channelizer(1000, 500, np.ones(16000))
    with pytest.raises(L.QuiskCudaError):
        Channelizer(1024, 2048, np.ones(16384))
    with pytest.raises(L.QuiskCudaError):
        Channelizer(1024, 512, np.ones(16000))
