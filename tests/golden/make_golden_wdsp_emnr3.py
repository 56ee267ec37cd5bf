"""tests/golden/make_golden_wdsp_emnr3.py -- fixtures for gain method 3 of WDSP's spectral noise reduction (the "trained"
method: emnr.c:965-1010, getZeta :866-884, the post-filter's extra damping :813-815; the second state of Quisk's NR2
button, quisk.py:6020-6023) from the COMPILED REFERENCE (oracle/_ref/libwdsp_ref.so, which holds the distribution's
default table zetahat.c: there is no `zetaHat` file in the working directory, so readZetaHat takes it, emnr.c:212-227).
Stage level: method 3 with each noise-power estimator, post-filter on and off.  Channel level: Quisk's channel with the button's second
state switched on mid-stream, once with the defaults and once with the two training parameters moved
(SetRXAEMNRtrainZetaThresh / SetRXAEMNRtrainT2).  Each with the reference's own sensitivity to a one-ulp change of its input (the table
turns a bin fully on or off, so a flipped cell would be a finite step: none flips on these inputs, 2e-16).
Writes tests/golden/wdsp_emnr3_kat.npz.   Run:  python tests/golden/make_golden_wdsp_emnr3.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_ctypes as R            # noqa: E402
from tests.golden.make_golden_wdsp_emnr import (BLOCKS, CH_BLOCKS, CH_ON, CH_TAIL, N, RATE, channel_input, rel_rms,   # noqa: E402
                                                stage_input, ulp)

D = C.c_double
CASES3 = [(0, 1), (1, 0), (2, 1)]      # (npe_method, ae_run) with gain method 3
TRAIN = (0.5, 0.6)                     # zeta_thresh, t2 of the second channel case (defaults: -2.0, 0.20)


def key3(npe, ae):
    return "emnr3_%d_%d" % (npe, ae)


def main():
    lib = R.load("libwdsp_ref.so")
    lib.create_emnr.restype = C.c_void_p
    lib.create_emnr.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, D, C.c_int, C.c_int, C.c_int]
    lib.xemnr.argtypes = [C.c_void_p, C.c_int]
    out = {}

    def stage(x, npe, ae):
        buf = np.zeros(N, dtype=np.complex128)
        a = lib.create_emnr(1, 0, N, buf.ctypes.data, buf.ctypes.data, 4096, 4, RATE, 0, 1.0, 3, npe, ae)
        ys = []
        for b in range(BLOCKS):
            buf[:] = x[b * N:(b + 1) * N] + 1j * 0.5 * x[b * N:(b + 1) * N]
            lib.xemnr(a, 0)
            ys.append(buf.copy())
        return np.concatenate(ys)

    x = stage_input()
    for npe, ae in CASES3:
        k = key3(npe, ae)
        y = stage(x, npe, ae)
        assert not y.imag.any()
        out[k + "/y"] = y.real.copy()
        out[k + "/cond"] = np.array([rel_rms(stage(ulp(x, 7), npe, ae), y)])
        print(k, "peak", np.abs(y).max(), "cond", out[k + "/cond"])

    lib.OpenChannel.argtypes = [C.c_int] * 8 + [D] * 4 + [C.c_int]
    lib.fexchange0.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.RXASetPassband.argtypes = [C.c_int, D, D]
    lib.SetRXAAGCFixed.argtypes = [C.c_int, D]
    lib.SetRXAEMNRtrainZetaThresh.argtypes = [C.c_int, D]
    lib.SetRXAEMNRtrainT2.argtypes = [C.c_int, D]

    def channel(xc, chn, train):
        lib.OpenChannel(chn, N, N, RATE, RATE, RATE, 0, 1, D(0.0), D(0.0), D(0.0), D(0.0), 1)
        lib.SetRXAShiftRun(chn, 0); lib.RXANBPSetRun(chn, 0); lib.SetRXAAMSQRun(chn, 0)
        lib.SetRXAMode(chn, 1)
        lib.RXASetPassband(chn, D(300.0), D(3000.0))
        lib.RXASetNC(chn, N); lib.RXASetMP(chn, 0)
        lib.SetRXAAGCMode(chn, 0); lib.SetRXAAGCFixed(chn, D(0.0))
        lib.SetRXAPanelRun(chn, 0); lib.SetRXAEMNRRun(chn, 0)
        inb = np.zeros(N, dtype=np.complex128); outb = np.zeros(N, dtype=np.complex128)
        err = C.c_int(0)
        ys = []
        for b in range(CH_BLOCKS):
            if b == CH_ON:
                time.sleep(0.05)
                if train:
                    lib.SetRXAEMNRtrainZetaThresh(chn, D(train[0])); lib.SetRXAEMNRtrainT2(chn, D(train[1]))
                lib.SetRXAEMNRgainMethod(chn, 3)         # quisk.py:6021-6022
                lib.SetRXAEMNRRun(chn, 1)
            inb[:] = xc[b * N:(b + 1) * N]
            lib.fexchange0(chn, inb.ctypes.data_as(C.c_void_p), outb.ctypes.data_as(C.c_void_p), C.byref(err))
            ys.append(outb.copy())
            time.sleep(0.004)
        lib.SetChannelState(chn, 0, 1)
        lib.CloseChannel(chn)
        return np.concatenate(ys)

    xc = channel_input()
    for name, train, chn in (("chan3", None, 7), ("chan3_train", TRAIN, 8)):
        yc = channel(xc, chn, train)
        out[name + "/y_tail"] = yc[-CH_TAIL * N:]
        ycp = channel(ulp(xc.view(np.float64), 11).view(np.complex128), chn + 4, train)
        out[name + "/cond"] = np.array([rel_rms(ycp[-CH_TAIL * N:], yc[-CH_TAIL * N:])])
        print(name, "tail peak", np.abs(out[name + "/y_tail"]).max(), "cond", out[name + "/cond"])
    np.savez_compressed(os.path.join(HERE, "wdsp_emnr3_kat.npz"), **out)
    print("wrote wdsp_emnr3_kat.npz")


if __name__ == "__main__":
    main()
