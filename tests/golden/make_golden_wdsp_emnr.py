"""tests/golden/make_golden_wdsp_emnr.py -- fixtures for WDSP's spectral noise reduction (wdsp/emnr.c: create_emnr / xemnr,
SetRXAEMNRRun ...) from the COMPILED REFERENCE (oracle/_ref/libwdsp_ref.so).  Stage level: every gain method this library
builds (0, 1, 2) and every noise-power estimator (0, 1, 2), post-filter on and off, 128 blocks of 256 samples (32 frames:
the minimum-statistics estimator closes three of its sub-windows), each with the reference's own sensitivity to a one-ulp
change of its input.  Channel level: Quisk's channel (quisk_wdsp.py:66-93) with NR2 switched on mid-stream.
Writes tests/golden/wdsp_emnr_kat.npz.   Run:  python tests/golden/make_golden_wdsp_emnr.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_ctypes as R            # noqa: E402

D = C.c_double
N, BLOCKS, RATE = 256, 128, 48000
CASES = [(2, 0, 1), (0, 0, 1), (1, 1, 0), (2, 2, 1)]       # (gain_method, npe_method, ae_run): every method once, the default pair first
CH_BLOCKS, CH_ON, CH_TAIL = 160, 40, 32


def speechlike(n, seed, rate=RATE):
    """a few tones that come and go over a noise floor: bins that are mostly signal, mostly noise, and in between"""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / rate
    x = 0.02 * rng.standard_normal(n)
    for f, a, per in ((700.0, 0.3, 0.9), (1230.0, 0.2, 0.37), (2100.0, 0.1, 1.7), (345.0, 0.15, 0.6)):
        env = 0.5 * (1.0 + np.sign(np.sin(2 * np.pi * t / per + f)))
        x += a * env * np.sin(2 * np.pi * f * t)
    return x


def stage_input():
    return speechlike(N * BLOCKS, 31)


def channel_input():
    xr = speechlike(N * CH_BLOCKS, 47)
    return (xr + 1j * np.roll(xr, 3)).astype(np.complex128)       # energy on both sides of the carrier


def ulp(v, seed):
    up = np.random.default_rng(seed).integers(0, 2, size=v.shape).astype(bool)
    return np.where(up, np.nextafter(v, np.inf), np.nextafter(v, -np.inf))


def rel_rms(a, b):
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / np.sqrt(np.mean(np.abs(b) ** 2)))


def main():
    lib = R.load("libwdsp_ref.so")
    lib.create_emnr.restype = C.c_void_p
    lib.create_emnr.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, D, C.c_int, C.c_int, C.c_int]
    lib.xemnr.argtypes = [C.c_void_p, C.c_int]
    lib.flush_emnr.argtypes = [C.c_void_p]
    out = {}

    def stage(x, gm, npe, ae, flush_at=-1):
        buf = np.zeros(N, dtype=np.complex128)
        a = lib.create_emnr(1, 0, N, buf.ctypes.data, buf.ctypes.data, 4096, 4, RATE, 0, 1.0, gm, npe, ae)
        ys = []
        for b in range(BLOCKS):
            if b == flush_at:
                lib.flush_emnr(a)
            buf[:] = x[b * N:(b + 1) * N] + 1j * 0.5 * x[b * N:(b + 1) * N]        # the imaginary rail is ignored and comes back zero
            lib.xemnr(a, 0)
            ys.append(buf.copy())
        return np.concatenate(ys)

    x = stage_input()
    for gm, npe, ae in CASES:
        key = "emnr_%d_%d_%d" % (gm, npe, ae)
        y = stage(x, gm, npe, ae, flush_at=90 if (gm, npe) == (2, 0) else -1)
        assert not y.imag.any()
        out[key + "/y"] = y.real.copy()                    # (the imaginary rail comes back zero: only the real rail is stored)
        out[key + "/cond"] = np.array([rel_rms(stage(ulp(x, 7), gm, npe, ae, flush_at=90 if (gm, npe) == (2, 0) else -1), y)])
        print(key, "peak", np.abs(y).max(), "imag", np.abs(y.imag).max(), "cond", out[key + "/cond"])

    # ---- the channel as Quisk opens it, NR2 switched on at block CH_ON (and the post-filter off again later)
    lib.OpenChannel.argtypes = [C.c_int] * 8 + [D] * 4 + [C.c_int]
    lib.fexchange0.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.RXASetPassband.argtypes = [C.c_int, D, D]
    lib.SetRXAAGCFixed.argtypes = [C.c_int, D]

    def channel(xc, chn):
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
                lib.SetRXAEMNRgainMethod(chn, 2)
                lib.SetRXAEMNRRun(chn, 1)
            inb[:] = xc[b * N:(b + 1) * N]
            lib.fexchange0(chn, inb.ctypes.data_as(C.c_void_p), outb.ctypes.data_as(C.c_void_p), C.byref(err))
            ys.append(outb.copy())
            time.sleep(0.004)
        lib.SetChannelState(chn, 0, 1)
        lib.CloseChannel(chn)
        return np.concatenate(ys)

    xc = channel_input()
    yc = channel(xc, 5)
    out["chan/y_tail"] = yc[-CH_TAIL * N:]
    v = xc.view(np.float64)
    ycp = channel(ulp(v, 11).view(np.complex128), 6)
    out["chan/cond"] = np.array([rel_rms(ycp[-CH_TAIL * N:], yc[-CH_TAIL * N:])])
    print("channel tail peak", np.abs(out["chan/y_tail"]).max(), "cond", out["chan/cond"])
    np.savez_compressed(os.path.join(HERE, "wdsp_emnr_kat.npz"), **out)
    print("wrote wdsp_emnr_kat.npz")


if __name__ == "__main__":
    main()
