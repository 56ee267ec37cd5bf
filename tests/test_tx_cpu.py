"""The TX fixture is pinned to the compiled reference: oracle/_ref/libquisk_tx_ref.so (microphone.c's tx_filter + CcmPeak
extracted at build time + filter.c verbatim) reproduces tests/golden/tx_kat.npz bit for bit from the committed generator's
input, and FM takes the same branch as AM."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref_ctypes as R
from tests.golden.make_golden_tx import CLIP, MIC_RATE, PREEMPH, TX_MODES, TX_SPLITS, mic_audio
from tests.util import golden


@pytest.mark.skipif(not R.have_ref("libquisk_tx_ref.so"), reason="compiled reference not built (oracle/build_ref.sh)")
@pytest.mark.parametrize("mode", ["USB", "LSB", "AM", "FM"])
def test_tx_fixture_is_the_compiled_reference(mode):
    kat = golden("tx_kat.npz")
    x = mic_audio()
    lib = R.load("libquisk_tx_ref.so", private_copy=True)
    lib.ref_tx_init.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
    lib.ref_tx_filter.argtypes = [C.c_void_p, C.c_int]
    lib.ref_tx_init(TX_MODES[mode], MIC_RATE, PREEMPH, CLIP)
    outs, counts, pos = [], [], 0
    for n in TX_SPLITS:
        buf = np.zeros(max(2 * n, 16), dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
        nr = lib.ref_tx_filter(buf.ctypes.data_as(C.c_void_p), n)
        outs.append(buf[:nr].copy()); counts.append(nr)
    y = np.concatenate(outs)
    key = "AM" if mode == "FM" else mode
    assert counts == kat["tx_%s/counts" % key].tolist()
    ref = kat["tx_%s/y" % key]
    if key == "AM":
        assert not y.imag.any()
        y = y.real
    assert np.array_equal(y, ref)


@pytest.mark.skipif(not R.have_ref("libquisk_tx_ref.so"), reason="compiled reference not built (oracle/build_ref.sh)")
@pytest.mark.parametrize("mode,key", [("DGT-U", "DGT-U"), ("DGT-L", "DGT-L"), ("FDV-U", "DGT-U"), ("FDV-L", "DGT-L")])
def test_tx_digital_fixture_is_the_compiled_reference(mode, key):
    kat = golden("tx_kat.npz")
    x = mic_audio()
    lib = R.load("libquisk_tx_ref.so", private_copy=True)
    lib.ref_tx_filter_digital.argtypes = [C.c_void_p, C.c_int]
    lib.ref_tx_digital_init(R.MODES[mode])
    outs, pos = [], 0
    for n in TX_SPLITS[:6]:
        buf = np.zeros(max(n, 16), dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
        assert lib.ref_tx_filter_digital(buf.ctypes.data_as(C.c_void_p), n) == n
        outs.append(buf[:n].copy())
    assert np.array_equal(np.concatenate(outs), kat["txd_%s/y" % key])


@pytest.mark.skipif(not R.have_ref("libquisk_tx_ref.so"), reason="compiled reference not built (oracle/build_ref.sh)")
@pytest.mark.parametrize("name", ["USB", "DGT-U"])
def test_alc_fixture_is_the_compiled_reference(name):
    from tests.golden.make_golden_tx import alc_chain
    assert np.array_equal(alc_chain(name, mic_audio()), golden("tx_kat.npz")["alc_%s/y" % name])
