"""The waterfall pixel mapper (quisk.c:5334-5480): fixtures produced by calling the reference's own
watfall_RgbData / watfall_OnGraphData / watfall_GetPixels (tests/golden/make_golden_waterfall.py).  CPU: the oracle's
restatement against them; GPU: quisk_cuda_waterfall_* against them, several streams at once, byte for byte."""
import ctypes as C

import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.golden.make_golden_waterfall import GETS, MAX_HEIGHT, ROWS, WIDTH, palette, params, rows_db
from tests.util import golden

CHECK_ROWS = (0, 3, 40, ROWS - 1)


@pytest.mark.parametrize("mode", [0, 1])
def test_waterfall_oracle_matches_reference(mode):
    kat = golden("waterfall_kat.npz")
    red, green, blue = palette()
    wf = O.WaterfallOracle(red, green, blue, WIDTH, MAX_HEIGHT)
    for k, db in enumerate(rows_db()):
        wf.on_graph_data(db, *params(k))
        if k in CHECK_ROWS:
            for gi, (xo, h) in enumerate(GETS):
                ref = kat["mode%d/row%d/get%d" % (mode, k, gi)]
                got = wf.get_pixels(xo, h, mode)
                assert np.array_equal(got, ref[:len(got)]) and len(got) == WIDTH * 3 * h
                assert not ref[len(got):].any()                 # the reference wrote exactly `height` lines


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_waterfall_gpu_matches_reference(mode):
    import torch
    from quisk_b200 import lib as L
    lib = L.require_device()
    kat = golden("waterfall_kat.npz")
    red, green, blue = palette()
    S = 3
    wf = lib.quisk_cuda_waterfall_create(S, WIDTH, MAX_HEIGHT, red.ctypes.data, green.ctypes.data, blue.ctypes.data)
    assert wf, lib.quisk_cuda_last_error()
    hmax = max(h for _, h in GETS)
    out = torch.zeros((S, hmax * WIDTH * 3), dtype=torch.uint8, device="cuda")
    for k, db in enumerate(rows_db()):
        y_zero, y_scale, gain, x_origin = params(k)
        # stream 0 and 2 get the fixture's line, stream 1 the same line reversed (its own pixels, same geometry)
        lines = np.stack([db, db[::-1], db])
        d = torch.from_numpy(np.ascontiguousarray(lines)).cuda()
        assert lib.quisk_cuda_waterfall_on_graph_data(wf, d.data_ptr(), d.stride(0), len(db), y_zero, y_scale, gain, x_origin, None) == 0, lib.quisk_cuda_last_error()
        if k in CHECK_ROWS:
            for gi, (xo, h) in enumerate(GETS):
                out.zero_()
                assert lib.quisk_cuda_waterfall_get_pixels(wf, out.data_ptr(), out.stride(0), xo, h, mode, None) == 0, lib.quisk_cuda_last_error()
                torch.cuda.synchronize()
                got = out.cpu().numpy()
                ref = kat["mode%d/row%d/get%d" % (mode, k, gi)][:WIDTH * 3 * h]
                assert np.array_equal(got[0, :len(ref)], ref)
                assert np.array_equal(got[2, :len(ref)], ref)
                assert not np.array_equal(got[1, :len(ref)], ref) or not ref.any()
    # scroll mode needs room for its 35 repeated lines
    assert lib.quisk_cuda_waterfall_get_pixels(wf, out.data_ptr(), out.stride(0), 0, 20, 1, None) != 0
    lib.quisk_cuda_waterfall_destroy(wf)
