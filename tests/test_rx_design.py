"""quisk_cuda_make_filter_coef (quisk_b200/csrc/rx_design.cpp) against the reference's own MakeFilterCoef
(quisk.py:5405-5456).  The reference method is a wx application method, so the committed fixture holds outputs of the
oracle's restatement, and -- when /root/reference is present -- the test also executes the reference's own source
lines for that method and GetFilterCenter directly (extracted by line range, run in a scratch namespace)."""
import math
import os

import numpy as np
import pytest

from oracle import quisk_oracle as O
from tests.util import golden

CASES = [(12000, None, 2800, 1700), (6000, None, 500, 600), (24000, None, 6000, 0), (48000, None, 3200, 1900),
         (12000, None, 2400, 1500), (12000, 301, 2800, 1700), (48000, None, 16000, 0), (6000, None, 200, 600),
         (12000, 100, 1000, 800), (48000, None, 19000, 0)]


def _ref_make_filter_coef():
    """MakeFilterCoef as the reference wrote it: the method's own source lines, compiled as a function."""
    path = "/root/reference/quisk.py"
    if not os.path.exists(path):
        return None
    lines = open(path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith("  def MakeFilterCoef("))
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("  def "))
    src = "\n".join(l[2:] for l in lines[start:end])
    import cmath
    import sys
    sys.path.insert(0, "/root/reference")
    import filters as ref_filters
    ns = {"math": math, "cmath": cmath, "Filters": ref_filters.Filters}
    exec(src, ns)
    return lambda *a: ns["MakeFilterCoef"](None, *a)


@pytest.mark.parametrize("rate,N,bw,center", CASES)
def test_make_filter_coef_bit_identical(rate, N, bw, center):
    from quisk_b200.rx import load_tables, make_filter_coef
    tabs = load_tables()
    protos = {int(k[6:]): tabs[k] for k in tabs if k.startswith("proto_")}
    fi, fq = make_filter_coef(rate, N, bw, center, tabs)
    oi, oq = O.make_filter_coef(rate, N, bw, center, protos)
    assert len(fi) == len(oi)
    assert np.array_equal(fi, oi) and np.array_equal(fq, oq)
    ref = _ref_make_filter_coef()
    if ref is not None:
        ri, rq = ref(rate, N, bw, center)
        assert np.array_equal(fi, np.array(ri)) and np.array_equal(fq, np.array(rq))


def test_c1_filter_equals_fixture():
    """The C1 receive filter (USB, bw 2800 at 12 kS/s -> 164 taps) the bench and smoke() design for themselves equals
    the one the chain fixtures were generated with."""
    from quisk_b200.rx import get_filter_center, make_filter_coef
    kat = golden("chain_kat.npz")
    c = get_filter_center("USB", 2800)
    assert c == 1700 and get_filter_center("LSB", 2800) == -1700 and get_filter_center("CWL", 500) == -600
    assert get_filter_center("FDV-U", 3200) == 1600 and get_filter_center("DGT-U", 2800) == 1500
    fi, fq = make_filter_coef(12000, None, 2800, c)
    assert len(fi) == 164
    assert np.array_equal(fi, kat["c1/filt_i"]) and np.array_equal(fq, kat["c1/filt_q"])
