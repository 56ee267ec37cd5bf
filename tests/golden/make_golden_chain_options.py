"""tests/golden/make_golden_chain_options.py -- the optional stages INSIDE quisk_process_demodulate, switched on with the
reference's own variables (quisk_auto_notch, ssb_squelch_enabled / ssb_squelch_level): USB and CWU at 48 kS/s through the
compiled reference's quisk_process_decimate + quisk_process_demodulate (oracle/_ref/libquisk_rx_ref.so), block by block,
with the muting quisk_process_samples applies to a block whose squelch is closed (quisk.c:2716-2719).  The signal is a steady
carrier (the notch's target, switched off after 40 blocks) under speech-like components that stop for 50 blocks (only band
noise is left: the squelch closes a second later) and come back.  The fixture keeps the audio of the blocks around those
events (KEEP), the squelch flag and the RMS of every block.
Writes tests/golden/chain_options_kat.npz.   Run:  python tests/golden/make_golden_chain_options.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

RATE, BLOCK, BLOCKS = 48000, 1024, 150
CASES = [("USB", 1, 150, 0), ("CWU", 1, 0, 600), ("USB", 0, 150, 0)]    # (mode, auto notch, squelch level, rit / side tone)
KEEP = list(range(0, 4)) + list(range(12, 16)) + list(range(30, 44)) + list(range(94, 98)) + list(range(102, 108))     # the blocks whose audio the fixture keeps


def options_input(seed):
    rng = np.random.default_rng(seed)
    n = BLOCK * BLOCKS
    t = np.arange(n)
    x = 0.002 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    # speech-like: a few wandering components between 300 and 2500 Hz with a syllable envelope
    env = np.clip(np.sin(2 * np.pi * 3.1 * t / RATE), 0, None) ** 2
    for f0, a in ((420.0, 0.3), (910.0, 0.2), (1730.0, 0.15), (2380.0, 0.1)):
        f = f0 * (1.0 + 0.03 * np.sin(2 * np.pi * 1.3 * t / RATE + f0))
        x += a * env * np.exp(2j * np.pi * np.cumsum(f) / RATE)
    quiet = slice(50 * BLOCK, 100 * BLOCK)
    x[quiet] = 0.002 * (rng.standard_normal(50 * BLOCK) + 1j * rng.standard_normal(50 * BLOCK))
    car = 0.25 * np.exp(2j * np.pi * 1250.0 * t / RATE)         # steady carrier inside the pass band, switched off in block 40
    car[40 * BLOCK:] = 0.0
    x += car
    return (x * 1.0e6).astype(np.complex128)                    # Quisk's sample scale is integer-like


def main():
    from oracle import ref_ctypes as R
    from tests.golden.make_golden import demod_taps
    out = {}
    for mode, notch, level, rit in CASES:
        lib = R.load("libquisk_rx_ref.so", private_copy=True)
        lib.ref_set_sample_rate(RATE); lib.ref_init_chain()
        fi, fq = demod_taps(mode)
        fi = np.ascontiguousarray(fi); fq = np.ascontiguousarray(fq)
        lib.ref_set_filters(fi.ctypes.data_as(C.c_void_p), fq.ctypes.data_as(C.c_void_p), len(fi), 2800 if mode == "USB" else 500, 0)
        lib.ref_set_chain_options(notch, 1 if level else 0, level, rit)
        x = options_input(31)
        ys, act = [], []
        for b in range(BLOCKS):
            buf = np.zeros(66000, dtype=np.complex128); buf[:BLOCK] = x[b * BLOCK:(b + 1) * BLOCK]
            dbuf = np.zeros(132000)
            n = lib.ref_process_decimate(buf.ctypes.data_as(C.c_void_p), BLOCK, 0, R.MODES[mode])
            nr = lib.ref_process_demodulate(buf.ctypes.data_as(C.c_void_p), dbuf.ctypes.data_as(C.c_void_p), n, 0, 0, R.MODES[mode])
            a = lib.ref_squelch_active(0)
            y = dbuf[:nr].copy()
            if a:
                y[:] = 0.0
            ys.append(y); act.append(a)
        key = "%s_n%d_s%d" % (mode, notch, level)
        assert all(len(y) == BLOCK for y in ys)
        out[key + "/y"] = np.stack([ys[b] for b in KEEP])
        out[key + "/active"] = np.array(act, dtype=np.int32)
        out[key + "/rms"] = np.array([np.sqrt(np.mean(y * y)) for y in ys])
        print(key, "peak", np.abs(out[key + "/y"]).max(), "squelched blocks", [b for b in range(BLOCKS) if act[b]][:3], "...", int(np.sum(act)), "of", BLOCKS,
              "rms blocks 20 / 39", out[key + "/rms"][20], out[key + "/rms"][39])
    np.savez_compressed(os.path.join(HERE, "chain_options_kat.npz"), **out)
    print("wrote chain_options_kat.npz", os.path.getsize(os.path.join(HERE, "chain_options_kat.npz")))


if __name__ == "__main__":
    main()
