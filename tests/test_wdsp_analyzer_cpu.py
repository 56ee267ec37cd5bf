"""The analyzer fixtures are what the COMPILED REFERENCE produces: three of the configurations (one of them a stitched span) are run again through
oracle/_ref/libwdsp_ref.so (XCreateAnalyzer / SetAnalyzer / Spectrum0 / GetPixels on the reference's own worker threads)
and must reproduce the committed lines bit for bit.  Skipped where oracle/_ref is not built."""
import numpy as np
import pytest

from oracle import ref_ctypes as R
from tests.golden.make_golden_wdsp_analyzer import CASES, FRAMES, bind, frames_expected, run_case
from tests.util import golden

pytestmark = pytest.mark.skipif(not R.have_ref("libwdsp_ref.so"), reason="oracle/_ref not built")


@pytest.mark.parametrize("disp,name", [(40, "hamming_rect256"), (41, "kaiser2048"), (42, "stitch3_skip")])
def test_fixture_is_the_reference(disp, name):
    lib = bind(R.load("libwdsp_ref.so"))
    kat = golden("wdsp_analyzer_kat.npz")
    res = run_case(lib, disp, name, CASES[name])
    for key, v in res.items():
        assert v.shape == (FRAMES, CASES[name]["npix"])
        assert np.array_equal(v, kat[key]), key


def test_frame_schedule():
    """one frame per call once `size` samples wait, hop = size - overlap (analyzer.c:884-911, 1561-1570)"""
    assert frames_expected(dict(sz=2048, hop=1024), 5) == [0, 1, 1, 1, 1]
    assert frames_expected(dict(sz=512, hop=128), 6) == [0, 0, 0, 1, 1, 1]
    assert frames_expected(dict(sz=256, hop=256), 3) == [1, 1, 1]
