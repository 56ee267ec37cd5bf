"""tests/golden/make_golden_tx.py -- fixtures for the transmit-audio chain of the reference's microphone.c (tx_filter,
microphone.c:372-604, with CcmPeak :161-233) from the COMPILED REFERENCE (oracle/_ref/libquisk_tx_ref.so: the two
functions extracted at build time + filter.c verbatim, oracle/ref_wrap/quisk_tx_wrap.c).  Microphone audio at 48 kS/s in
the +-CLIP16 range, ragged blocks, the four modes tx_filter distinguishes (LSB / USB: the SSB branch; AM / FM: the real
branch), with pre-emphasis and enough clip gain that the compressor, the clipper and the peak rounder all engage.
Also tx_filter_digital (microphone.c:605-624), the digital modes' one-filter chain, both side bands.
Writes tests/golden/tx_kat.npz.   Run:  python tests/golden/make_golden_tx.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_ctypes as R            # noqa: E402

MIC_RATE = 48000
TX_MODES = {"LSB": 2, "USB": 3, "AM": 4, "FM": 5}
DGT_TX_MODES = {"DGT-U": 7, "DGT-L": 8}       # FDV-U / FDV-L tune like DGT-U / DGT-L (microphone.c:617)
PREEMPH, CLIP = 0.6, 2.5
ALC_SPLITS = [4800, 4806, 6, 12, 960, 954, 2004, 6000]     # blocks of the ALC fixtures
ALC_CASES = ["USB", "DGT-U"]
ALC_IN_SCALE = {"USB": 1.0, "DGT-U": 2.5}
ALC_KEY_DOWN_AT = 6                                         # init_alc(&tx_alc, 0) in front of this block
TX_SPLITS = [4800, 4806, 1, 5, 9600, 1023, 12000, 600, 6, 7, 48000 - 4800 - 4806 - 1 - 5 - 9600 - 1023 - 12000 - 600 - 6 - 7, 24000]


def mic_audio(n=72000, seed=5):
    """speech-like: voiced bursts with pauses, a few loud peaks, a noise floor; int16 scale"""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / MIC_RATE
    env = np.clip(np.sin(2 * np.pi * 1.7 * t) + 0.3, 0.0, 1.0) ** 2
    x = env * (0.5 * np.sin(2 * np.pi * 220 * t) + 0.3 * np.sin(2 * np.pi * 660 * t + 1.0) + 0.25 * np.sin(2 * np.pi * 1500 * t + 2.0)
               + 0.1 * np.sin(2 * np.pi * 2500 * t))
    x += 0.003 * rng.standard_normal(n)
    x[30000:30040] += 1.5                      # a thump
    return np.round(x * 12000.0)


def alc_chain(name, x):
    """tx_filter (USB) or tx_filter_digital (DGT-U, microphone audio x 2.5 so that the control has to pull the gain down from
    its initial 1.4) followed by process_alc, block by block, on a fresh copy of the compiled reference"""
    lib = R.load("libquisk_tx_ref.so", private_copy=True)
    lib.ref_tx_init.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
    lib.ref_tx_filter.argtypes = [C.c_void_p, C.c_int]
    lib.ref_tx_filter_digital.argtypes = [C.c_void_p, C.c_int]
    lib.ref_process_alc.argtypes = [C.c_void_p, C.c_int, C.c_int]
    mode = R.MODES[name]
    if name == "USB":
        lib.ref_tx_init(mode, MIC_RATE, PREEMPH, CLIP)
    else:
        lib.ref_tx_digital_init(mode)
    lib.ref_alc_init()
    outs, pos = [], 0
    for k, n in enumerate(ALC_SPLITS):
        if k == ALC_KEY_DOWN_AT:
            lib.ref_alc_key_down()
        buf = np.zeros(max(2 * n, 16), dtype=np.complex128); buf[:n] = x[pos:pos + n] * ALC_IN_SCALE[name]; pos += n
        nr = lib.ref_tx_filter(buf.ctypes.data_as(C.c_void_p), n) if name == "USB" else lib.ref_tx_filter_digital(buf.ctypes.data_as(C.c_void_p), n)
        lib.ref_process_alc(buf.ctypes.data_as(C.c_void_p), nr, mode)
        outs.append(buf[:nr].copy())
    return np.concatenate(outs)


def main():
    out = {}
    x = mic_audio()
    for name, mode in TX_MODES.items():
        lib = R.load("libquisk_tx_ref.so", private_copy=True)
        lib.ref_tx_init.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        lib.ref_tx_filter.argtypes = [C.c_void_p, C.c_int]
        lib.ref_tx_init(mode, MIC_RATE, PREEMPH, CLIP)
        outs, counts, pos = [], [], 0
        for n in TX_SPLITS:
            buf = np.zeros(max(2 * n, 16), dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
            nr = lib.ref_tx_filter(buf.ctypes.data_as(C.c_void_p), n)
            outs.append(buf[:nr].copy()); counts.append(nr)
        y = np.concatenate(outs)
        if name == "FM":                       # the same branch as AM (is_ssb is all tx_filter asks): one fixture serves both
            assert not y.imag.any() and np.array_equal(y.real, out["tx_AM/y"])
            continue
        if name == "AM":
            assert not y.imag.any()
            y = y.real.copy()                  # the real branch leaves the imaginary rail zero: only the real rail is stored
        out["tx_%s/y" % name] = y
        out["tx_%s/counts" % name] = np.array(counts)
        print(name, "out", len(y), "peak", np.abs(y).max(), "rms", np.sqrt(np.mean(np.abs(y) ** 2)))
    # tx_filter_digital (microphone.c:605-624): one tuned 520-tap filter at 48 kS/s, upper and lower side band
    for name, mode in DGT_TX_MODES.items():
        lib = R.load("libquisk_tx_ref.so", private_copy=True)
        lib.ref_tx_filter_digital.argtypes = [C.c_void_p, C.c_int]
        lib.ref_tx_digital_init(mode)
        outs, pos = [], 0
        for n in TX_SPLITS[:6]:
            buf = np.zeros(max(n, 16), dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
            nr = lib.ref_tx_filter_digital(buf.ctypes.data_as(C.c_void_p), n)
            assert nr == n
            outs.append(buf[:nr].copy())
        out["txd_%s/y" % name] = np.concatenate(outs)
        print(name, "digital out", len(out["txd_%s/y" % name]), "peak", np.abs(out["txd_%s/y" % name]).max())
    # process_alc (microphone.c:270-370) behind tx_filter / tx_filter_digital, as quisk_process_microphone chains them
    # (:1232-1233): the look-ahead level control with its 960-sample delay line, a key down in the middle (init_alc(.., 0), :1207)
    for name in ALC_CASES:
        out["alc_%s/y" % name] = alc_chain(name, x)
        print(name, "alc out", len(out["alc_%s/y" % name]), "peak", np.abs(out["alc_%s/y" % name]).max())
    np.savez_compressed(os.path.join(HERE, "tx_kat.npz"), **out)
    print("wrote tx_kat.npz")


if __name__ == "__main__":
    main()
