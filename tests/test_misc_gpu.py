"""GPU parity of process_agc, cFracDecim (fixtures from the compiled reference) and get_bandscope (NumPy oracle)."""
import ctypes as C

import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.golden.make_golden_misc import AGC_SPLITS, agc_input
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
