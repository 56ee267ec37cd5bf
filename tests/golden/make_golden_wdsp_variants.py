"""tests/golden/make_golden_wdsp_variants.py -- known-answer fixtures for the three relatives of fircore that wdsp
defines and never instantiates: xfirmin (wdsp/firmin.c:76-99), xfiropt (firmin.c:227-251), xbps
(wdsp/bandpass.c:85-105), generated from the COMPILED REFERENCE (oracle/_ref/libwdsp_ref.so, oracle/build_ref.sh).
Writes tests/golden/wdsp_variants_kat.npz.   Run:  python tests/golden/make_golden_wdsp_variants.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_ctypes as R            # noqa: E402
from quisk_b200.synth import sig              # noqa: E402

D = C.c_double
VP = C.c_void_p
# (size, nc, rate, f_low, f_high, wintype, gain, blocks)
FIRMIN_CASES = [(64, 128, 48000, 150.0, 2850.0, 0, 1.0, 6), (100, 256, 48000, -2850.0, -150.0, 1, 0.5, 5)]
FIROPT_CASES = [(64, 256, 48000, 150.0, 2850.0, 0, 1.0 / 128, 8), (256, 256, 48000, -3000.0, 3000.0, 1, 1.0 / 512, 6)]
BPS_CASES = [(256, 48000, 150.0, 2850.0, 0, 1.0, 6), (1024, 192000, -2850.0, -150.0, 1, 2.0, 4)]


def main():
    lib = R.load("libwdsp_ref.so")
    args = [C.c_int, C.c_int, C.c_int, VP, VP, C.c_int, D, D, C.c_int, C.c_int, D]
    for n in ("firmin", "firopt"):
        getattr(lib, "create_" + n).restype = VP; getattr(lib, "create_" + n).argtypes = args
        getattr(lib, "x" + n).argtypes = [VP, C.c_int]
        getattr(lib, "flush_" + n).argtypes = [VP]
    lib.create_bps.restype = VP
    lib.create_bps.argtypes = [C.c_int, C.c_int, C.c_int, VP, VP, D, D, C.c_int, C.c_int, D]
    lib.xbps.argtypes = [VP, C.c_int]
    lib.flush_bps.argtypes = [VP]
    out = {}

    def run(create, x_fn, flush_fn, size, blocks, seed, rate, *cargs):
        inb = np.zeros(size, dtype=np.complex128); outb = np.zeros(2 * size, dtype=np.complex128)
        a = create(1, 0, size, inb.ctypes.data, outb.ctypes.data, *cargs)
        x = sig(size * blocks, seed, float(rate))
        ys = []
        for b in range(blocks):
            if b == blocks - 2:
                flush_fn(a)                                  # a flush mid-stream: the history restarts from zeros
            inb[:] = x[b * size:(b + 1) * size]
            x_fn(a, 0)
            ys.append(outb[:size].copy())
        return np.concatenate(ys)

    for size, nc, rate, fl, fh, wt, gain, blocks in FIRMIN_CASES:
        out["firmin_%d_%d/y" % (size, nc)] = run(lib.create_firmin, lib.xfirmin, lib.flush_firmin, size, blocks, 900 + size, rate, nc, fl, fh, rate, wt, gain)
    for size, nc, rate, fl, fh, wt, gain, blocks in FIROPT_CASES:
        out["firopt_%d_%d/y" % (size, nc)] = run(lib.create_firopt, lib.xfiropt, lib.flush_firopt, size, blocks, 910 + size, rate, nc, fl, fh, rate, wt, gain)
    for size, rate, fl, fh, wt, gain, blocks in BPS_CASES:
        out["bps_%d/y" % size] = run(lib.create_bps, lib.xbps, lib.flush_bps, size, blocks, 920 + size, rate, fl, fh, rate, wt, gain)
    np.savez_compressed(os.path.join(HERE, "wdsp_variants_kat.npz"), **out)
    print("wrote wdsp_variants_kat.npz:", {k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


if __name__ == "__main__":
    main()
