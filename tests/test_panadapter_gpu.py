"""GPU parity of the in-house batched FFT and the panadapter (get_graph math,
quisk.c:5142-5331, get_multirx_graph quisk.c:4868-4930) against the NumPy oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import quisk_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def lib():
    from quisk_b200 import lib as L
    return L.require_device()


@pytest.mark.parametrize("n", [8, 16, 32, 64, 256, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("sign", [-1, 1])
def test_fft_batch_vs_numpy(n, sign, torch, lib):
    rng = np.random.default_rng(n)
    batch = 7
    x = rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))
    d = torch.from_numpy(x).cuda()
    o = torch.empty_like(d)
    assert lib.quisk_cuda_fft_batch(d.data_ptr(), o.data_ptr(), n, batch, sign, None) == 0
    torch.cuda.synchronize()
    ref = np.fft.fft(x, axis=1) if sign < 0 else np.fft.ifft(x, axis=1) * n
    assert O.rel_rms(o.cpu().numpy(), ref) < 1e-14
    # in place
    assert lib.quisk_cuda_fft_batch(d.data_ptr(), d.data_ptr(), n, batch, sign, None) == 0
    torch.cuda.synchronize()
    assert O.rel_rms(d.cpu().numpy(), ref) < 1e-14


def test_fft_rejects_unsupported_sizes(lib):
    assert lib.quisk_cuda_fft_batch(None, None, 1000, 1, -1, None) != 0
    assert b"power of two" in lib.quisk_cuda_last_error()


@pytest.mark.parametrize("streams,count_fft", [(16, 1), (16, 8), (300, 3)])
def test_panadapter_c2(streams, count_fft, torch, lib):
    """BASELINE.json configs[1]: 16 streams, 8192-pt FFT + Hann + |X| averaging + dB graph."""
    n = 8192
    data_width = 1024
    frames = np.stack([O.synth_iq(n * count_fft, 50 + s, 1.0).reshape(count_fft, n) for s in range(streams)])
    d = torch.from_numpy(frames.reshape(streams, -1)).cuda()
    pan = lib.quisk_cuda_pan_create(streams, n)
    assert pan
    assert lib.quisk_cuda_pan_accumulate(pan, d.data_ptr(), n * count_fft, count_fft, None) == 0
    torch.cuda.synchronize()
    assert lib.quisk_cuda_pan_count(pan) == count_fft
    avg_ptr = lib.quisk_cuda_pan_average_ptr(pan)
    avg = torch.empty((streams, n), dtype=torch.float64, device="cuda")
    C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(avg.data_ptr()), C.c_void_p(avg_ptr), streams * n * 8, 3)
    avg = avg.cpu().numpy()
    g = torch.empty((streams, data_width), dtype=torch.float64, device="cuda")
    assert lib.quisk_cuda_pan_graph(pan, data_width, 1.0, 0.0, 192000.0, g.data_ptr(), None) == 0
    torch.cuda.synchronize()
    g = g.cpu().numpy()
    for s in range(0, streams, max(1, streams // 16)):
        ref_avg = O.panadapter_accumulate(frames[s])
        assert O.rel_rms(avg[s], ref_avg) < 1e-12
        ref_g = O.panadapter_pixels(ref_avg, count_fft, data_width, 1.0, 0.0, 192000.0)
        assert np.max(np.abs(g[s] - ref_g)) < 1e-9          # dB
    assert lib.quisk_cuda_pan_count(pan) == 0
    lib.quisk_cuda_pan_destroy(pan)


def test_panadapter_incremental_and_zoom(torch, lib):
    """Frames fed one call at a time give the same averages; zoomed graph with the in-place
    pixel aliasing of quisk.c:5289-5301 (pixel i overwrites fft_avg[i] while still reading)."""
    n, streams, data_width = 1024, 300, 1024
    frames = np.stack([O.synth_iq(n * 4, 70 + s, 1.0).reshape(4, n) for s in range(streams)])
    pan = lib.quisk_cuda_pan_create(streams, n)
    for f in range(4):
        d = torch.from_numpy(np.ascontiguousarray(frames[:, f, :])).cuda()
        assert lib.quisk_cuda_pan_accumulate(pan, d.data_ptr(), n, 1, None) == 0
    g = torch.empty((streams, data_width), dtype=torch.float64, device="cuda")
    assert lib.quisk_cuda_pan_graph(pan, data_width, 0.5, 3000.0, 48000.0, g.data_ptr(), None) == 0
    torch.cuda.synchronize()
    g = g.cpu().numpy()
    for s in (0, 17, 299):
        ref_g = O.panadapter_pixels(O.panadapter_accumulate(frames[s]), 4, data_width, 0.5, 3000.0, 48000.0)
        assert np.max(np.abs(g[s] - ref_g)) < 1e-9
    lib.quisk_cuda_pan_destroy(pan)


def test_full_scale_tone_reads_minus_6_dB(torch, lib):
    """A full-scale (2^31-1) bin-centred tone reads ~ -6.02 dB: Hann coherent gain 0.5."""
    n = 4096
    t = np.arange(n)
    x = (2.0 ** 31 - 1) * np.exp(2j * np.pi * 256 * t / n)
    d = torch.from_numpy(x[None, :].copy()).cuda()
    pan = lib.quisk_cuda_pan_create(1, n)
    assert lib.quisk_cuda_pan_accumulate(pan, d.data_ptr(), n, 1, None) == 0
    g = torch.empty((1, n), dtype=torch.float64, device="cuda")
    assert lib.quisk_cuda_pan_graph(pan, n, 1.0, 0.0, 48000.0, g.data_ptr(), None) == 0
    torch.cuda.synchronize()
    g = g.cpu().numpy()[0]
    assert abs(g.max() - 20 * np.log10(0.5)) < 1e-6
    assert int(np.argmax(g)) == n // 2 + 256
    lib.quisk_cuda_pan_destroy(pan)


def test_multirx_graph(torch, lib):
    n, streams = 2048, 5
    x = np.stack([O.synth_iq(n, 90 + s, 1.0) for s in range(streams)])
    d = torch.from_numpy(x).cuda()
    pan = lib.quisk_cuda_pan_create(streams, n)
    g = torch.empty((streams, n // 8), dtype=torch.float64, device="cuda")
    assert lib.quisk_cuda_pan_multirx(pan, d.data_ptr(), n, g.data_ptr(), None) == 0
    torch.cuda.synchronize()
    g = g.cpu().numpy()
    for s in range(streams):
        assert np.max(np.abs(g[s] - O.multirx_graph(x[s]))) < 1e-9
    lib.quisk_cuda_pan_destroy(pan)


def test_panadapter_host_mirror(torch, lib):
    """quisk_b200.rx.Panadapter mirrors record_app / get_graph (quisk.c:5142-5331, 5946-6009): None until a frame has
    been averaged, then data_width dB values per stream, and the averages restart."""
    from quisk_b200.rx import Panadapter
    n, data_width, streams = 2048, 512, 3
    pan = Panadapter(streams, n, data_width, 48000.0)
    assert pan.get_graph() is None
    frames = np.stack([O.synth_iq(n * 2, 90 + s, 1.0).reshape(2, n) for s in range(streams)])
    d = torch.from_numpy(frames.reshape(streams, -1)).cuda()
    pan.add_frames(d.data_ptr(), 2 * n, 2)
    assert pan.count_fft == 2
    g = pan.get_graph(1.0, 0.0)
    assert g.shape == (streams, data_width)
    for s in range(streams):
        ref = O.panadapter_pixels(O.panadapter_accumulate(frames[s]), 2, data_width, 1.0, 0.0, 48000.0)
        assert np.max(np.abs(g[s] - ref)) < 1e-9
    assert pan.get_graph() is None
    pan.close()


@pytest.mark.parametrize("n", [1000, 1200, 2400, 3000, 4096 - 2, 1001, 77])
def test_panadapter_sizes_that_are_not_powers_of_two(n, torch, lib):
    """Quisk's fft_size = data_width * fft_mult (quisk.py:187-194, 4179) is generally not a power of two: those sizes
    run as a Bluestein convolution.  Same window, same fftshift rule (bin (k + N/2) mod N, integer N/2), same graph."""
    streams, count_fft = 3, 3
    data_width = min(n, 400)
    frames = np.stack([O.synth_iq(n * count_fft, 120 + s, 1.0).reshape(count_fft, n) for s in range(streams)])
    d = torch.from_numpy(frames.reshape(streams, -1)).cuda()
    pan = lib.quisk_cuda_pan_create(streams, n)
    assert pan, lib.quisk_cuda_last_error()
    assert lib.quisk_cuda_pan_accumulate(pan, d.data_ptr(), n * count_fft, count_fft, None) == 0, lib.quisk_cuda_last_error()
    torch.cuda.synchronize()
    avg_ptr = lib.quisk_cuda_pan_average_ptr(pan)
    avg = torch.empty((streams, n), dtype=torch.float64, device="cuda")
    C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(avg.data_ptr()), C.c_void_p(avg_ptr), streams * n * 8, 3)
    avg = avg.cpu().numpy()
    g = torch.empty((streams, data_width), dtype=torch.float64, device="cuda")
    assert lib.quisk_cuda_pan_graph(pan, data_width, 1.0, 0.0, 48000.0, g.data_ptr(), None) == 0
    torch.cuda.synchronize()
    g = g.cpu().numpy()
    for s in range(streams):
        ref_avg = O.panadapter_accumulate(frames[s])
        assert O.rel_rms(avg[s], ref_avg) < 1e-12
        ref_g = O.panadapter_pixels(ref_avg, count_fft, data_width, 1.0, 0.0, 48000.0)
        assert np.max(np.abs(g[s] - ref_g)) < 1e-8
    lib.quisk_cuda_pan_destroy(pan)
    assert not lib.quisk_cuda_pan_create(2, 17000)           # beyond the Bluestein range and not a power of two


@pytest.mark.parametrize("n,streams,count_fft", [(4800, 3, 3), (9600, 2, 5), (12000, 3, 2), (16384 - 2, 2, 2), (8193, 1, 3)])
def test_panadapter_large_sizes_that_are_not_powers_of_two(n, streams, count_fft, torch, lib):
    """Above 4096 points the Bluestein convolution (M = 16384 or 32768) runs as M / 4096 CTAs per frame and transform:
    e.g. a 1200-pixel graph with fft_mult 8 (quisk.py:4179) is 9600 points."""
    data_width = 1200 if n % 1200 == 0 else 1000
    frames = np.stack([O.synth_iq(n * count_fft, 160 + s, 1.0).reshape(count_fft, n) for s in range(streams)])
    d = torch.from_numpy(frames.reshape(streams, -1)).cuda()
    pan = lib.quisk_cuda_pan_create(streams, n)
    assert pan, lib.quisk_cuda_last_error()
    # two calls: the running sum carries over
    assert lib.quisk_cuda_pan_accumulate(pan, d.data_ptr(), n * count_fft, 1, None) == 0, lib.quisk_cuda_last_error()
    assert lib.quisk_cuda_pan_accumulate(pan, d.data_ptr() + n * 16, n * count_fft, count_fft - 1, None) == 0, lib.quisk_cuda_last_error()
    torch.cuda.synchronize()
    assert lib.quisk_cuda_pan_count(pan) == count_fft
    avg_ptr = lib.quisk_cuda_pan_average_ptr(pan)
    avg = torch.empty((streams, n), dtype=torch.float64, device="cuda")
    C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(avg.data_ptr()), C.c_void_p(avg_ptr), streams * n * 8, 3)
    avg = avg.cpu().numpy()
    g = torch.empty((streams, data_width), dtype=torch.float64, device="cuda")
    assert lib.quisk_cuda_pan_graph(pan, data_width, 1.0, 0.0, 192000.0, g.data_ptr(), None) == 0
    torch.cuda.synchronize()
    g = g.cpu().numpy()
    for s in range(streams):
        ref_avg = O.panadapter_accumulate(frames[s])
        assert O.rel_rms(avg[s], ref_avg) < 1e-12
        ref_g = O.panadapter_pixels(ref_avg, count_fft, data_width, 1.0, 0.0, 192000.0)
        assert np.max(np.abs(g[s] - ref_g)) < 1e-8
    lib.quisk_cuda_pan_destroy(pan)


@pytest.mark.parametrize("n,streams,count_fft", [(16384, 3, 3), (32768, 2, 2), (16384, 40, 1)])
def test_panadapter_large_frames(n, streams, count_fft, torch, lib):
    """16384- and 32768-point frames: a radix-4 / radix-8 decimation-in-frequency step on the way in, four / eight
    4096-point CTAs per frame."""
    data_width = 1024
    frames = np.stack([O.synth_iq(n * count_fft, 140 + s, 1.0).reshape(count_fft, n) for s in range(streams)])
    d = torch.from_numpy(frames.reshape(streams, -1)).cuda()
    pan = lib.quisk_cuda_pan_create(streams, n)
    assert pan, lib.quisk_cuda_last_error()
    assert lib.quisk_cuda_pan_accumulate(pan, d.data_ptr(), n * count_fft, count_fft, None) == 0, lib.quisk_cuda_last_error()
    torch.cuda.synchronize()
    avg_ptr = lib.quisk_cuda_pan_average_ptr(pan)
    avg = torch.empty((streams, n), dtype=torch.float64, device="cuda")
    C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(avg.data_ptr()), C.c_void_p(avg_ptr), streams * n * 8, 3)
    avg = avg.cpu().numpy()
    g = torch.empty((streams, data_width), dtype=torch.float64, device="cuda")
    assert lib.quisk_cuda_pan_graph(pan, data_width, 1.0, 0.0, 1536000.0, g.data_ptr(), None) == 0
    torch.cuda.synchronize()
    g = g.cpu().numpy()
    for s in range(0, streams, max(1, streams // 4)):
        ref_avg = O.panadapter_accumulate(frames[s])
        assert O.rel_rms(avg[s], ref_avg) < 1e-12
        ref_g = O.panadapter_pixels(ref_avg, count_fft, data_width, 1.0, 0.0, 1536000.0)
        assert np.max(np.abs(g[s] - ref_g)) < 1e-9
    lib.quisk_cuda_pan_destroy(pan)
