"""The filter.h drop-in under the reference's OWN orchestrator: oracle/_ref/libquisk_rx_dropin.so is the reference's
quisk_process_decimate / quisk_process_demodulate (quisk.c:1673-2160, extracted at build time, unmodified) linked
against quisk_b200/libquisk_cuda.so instead of filter.c (oracle/build_ref.sh step 2b).
Every quisk_cDecim2HB45 / quisk_cDecimate / quisk_cInterpDecim / quisk_dInterpolate / ... call those functions make
lands on the GPU; the per-sample RX filter and the detectors stay the reference's C.  The outputs must equal the
all-reference fixtures BIT FOR BIT, block counts included."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import quisk_oracle as O
from oracle import ref_ctypes as R
from tests.golden.make_golden import DEMOD_TAPS, demod_taps
from tests.util import CHAIN_SPLITS, DEMOD_SPLITS, golden

pytestmark = pytest.mark.gpu
SO = os.path.join(R.REF_DIR, "libquisk_rx_dropin.so")


def _subprocess_lib(rate):
    """The reference keeps its filter state in function-local statics: every case gets a private copy of the library
    placed next to the original (so its $ORIGIN-relative path to libquisk_cuda.so still resolves)."""
    import shutil
    import tempfile
    fd, tmp = tempfile.mkstemp(suffix=".so", dir=R.REF_DIR)
    os.close(fd)
    shutil.copy(SO, tmp)
    try:
        lib = C.CDLL(tmp)
    finally:
        os.unlink(tmp)
    lib.ref_set_sample_rate(rate); lib.ref_init_chain()
    return lib


@pytest.fixture(scope="module", autouse=True)
def need_gpu_and_lib():
    import torch
    assert torch.cuda.is_available()
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libquisk_rx_dropin.so not built (oracle/build_ref.sh needs /root/reference)")


@pytest.mark.parametrize("rate", [1536000, 240000, 250000, 111111])
def test_reference_process_decimate_on_gpu_filters(rate):
    kat = golden("chain_kat.npz")
    lib = _subprocess_lib(rate)
    x = O.synth_iq(40000, 9, 1.0)
    outs, counts, pos = [], [], 0
    for n in CHAIN_SPLITS:
        buf = np.zeros(66000, dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
        nr = lib.ref_process_decimate(buf.ctypes.data_as(C.c_void_p), n, 0, 3)
        outs.append(buf[:nr].copy()); counts.append(nr)
    assert counts == kat["decimate_%d/counts" % rate].tolist()
    assert lib.ref_decim_srate() == int(kat["decimate_%d/srate" % rate][0])
    assert np.array_equal(np.concatenate(outs), kat["decimate_%d/y" % rate])


@pytest.mark.parametrize("mode", list(DEMOD_TAPS))
def test_reference_process_demodulate_on_gpu_filters(mode):
    kat = golden("chain_kat.npz")
    lib = _subprocess_lib(48000)
    fi, fq = demod_taps(mode)
    fi = np.ascontiguousarray(fi); fq = np.ascontiguousarray(fq)
    lib.ref_set_filters(fi.ctypes.data_as(C.c_void_p), fq.ctypes.data_as(C.c_void_p), len(fi), 2800, 0)
    x = O.synth_iq(12000, 10, 1.0)
    outs, counts, pos = [], [], 0
    for n in DEMOD_SPLITS:
        buf = np.zeros(66000, dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
        dbuf = np.zeros(132000)
        nr = lib.ref_process_demodulate(buf.ctypes.data_as(C.c_void_p), dbuf.ctypes.data_as(C.c_void_p), n, 0, 0, R.MODES[mode])
        outs.append(dbuf[:nr].copy()); counts.append(nr)
    assert counts == kat["demod_%s/counts" % mode].tolist()
    assert np.array_equal(np.concatenate(outs), kat["demod_%s/y" % mode])
