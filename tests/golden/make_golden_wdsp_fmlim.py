"""tests/golden/make_golden_wdsp_fmlim.py -- fixture for the FM detector limiter (wdsp/fmd.c:49-73, 179-184;
SetRXAFMLimRun / SetRXAFMLimGain, fmd.c:337-363) from the COMPILED REFERENCE (oracle/_ref/libwdsp_ref.so): an FM channel
at 48 kS/s through OpenChannel + fexchange0 with the limiter switched on, then its gain changed mid-stream.
Writes tests/golden/wdsp_fmlim_kat.npz.   Run:  python tests/golden/make_golden_wdsp_fmlim.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_ctypes as R            # noqa: E402
from quisk_b200.synth import fm_sig          # noqa: E402

D = C.c_double
BLOCKS, N, TAIL, GAIN_AT = 270, 256, 24, 210      # the PLL's cold start (on the filter's pre-ringing) takes ~170 blocks to die away, as in make_golden_wdsp.py


def run(lib, ch, x, perturbed=False):
    lib.OpenChannel(ch, N, N, 48000, 48000, 48000, 0, 1, D(0.0), D(0.0), D(0.0), D(0.0), 1)
    lib.SetRXAShiftRun(ch, 0)
    lib.SetRXAMode(ch, 5)
    lib.RXASetPassband(ch, D(-8000.0), D(8000.0))
    lib.SetRXAFMLimRun(ch, 1)
    inb = np.zeros(N, dtype=np.complex128); outb = np.zeros(N, dtype=np.complex128)
    err = C.c_int(0)
    ys = []
    for b in range(BLOCKS):
        if b == GAIN_AT:
            time.sleep(0.05)                    # let the DSP thread finish the block in flight before the limiter is rebuilt
            lib.SetRXAFMLimGain(ch, D(-6.0))
        inb[:] = x[b * N:(b + 1) * N]
        lib.fexchange0(ch, inb.ctypes.data_as(C.c_void_p), outb.ctypes.data_as(C.c_void_p), C.byref(err))
        ys.append(outb.copy())
        time.sleep(0.004)                       # pace the caller like a sound card (see make_golden_wdsp.py: back to back the exchange is timing dependent)
    lib.SetChannelState(ch, 0, 1)
    lib.CloseChannel(ch)
    return np.concatenate(ys)


def main():
    lib = R.load("libwdsp_ref.so")
    lib.OpenChannel.argtypes = [C.c_int] * 8 + [D] * 4 + [C.c_int]
    lib.fexchange0.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.SetRXAFMLimGain.argtypes = [C.c_int, D]
    lib.RXASetPassband.argtypes = [C.c_int, D, D]
    x = fm_sig(N * BLOCKS, 810, 48000.0)
    y = run(lib, 3, x)
    seg = lambda v: np.concatenate([v[(GAIN_AT - TAIL) * N:GAIN_AT * N], v[-TAIL * N:]])       # noqa: E731
    v = np.ascontiguousarray(x).view(np.float64).copy()
    up = np.random.default_rng(9).integers(0, 2, size=v.shape).astype(bool)
    xp = np.where(up, np.nextafter(v, np.inf), np.nextafter(v, -np.inf)).view(np.complex128)
    yp = run(lib, 4, xp)
    cond = float(np.sqrt(np.mean(np.abs(seg(yp) - seg(y)) ** 2)) / np.sqrt(np.mean(np.abs(seg(y)) ** 2)))
    half = TAIL * N
    rr = lambda a, b: float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / np.sqrt(np.mean(np.abs(b) ** 2)))      # noqa: E731
    conds = np.array([rr(seg(yp)[:half], seg(y)[:half]), rr(seg(yp)[half:], seg(y)[half:])])
    print("per-segment sensitivity (before / after the gain change):", conds)
    # the limiter stage alone (create_wcpagc with calc_fmd's arguments, fmd.c:49-73) on a fixed audio signal scaled by
    # lim_pre_gain: loud and quiet stretches, so that attack, decay and the min_volts floor all occur
    lib.create_wcpagc.restype = C.c_void_p
    lib.create_wcpagc.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, D, D, C.c_int] + [D] * 8 + [C.c_int] + [D] * 4
    lib.xwcpagc.argtypes = [C.c_void_p]
    rng = np.random.default_rng(21)
    t = np.arange(N * 40)
    env = np.where((t // 1500) % 3 == 0, 2.5, np.where((t // 1500) % 3 == 1, 0.3, 0.02))
    a = env * np.sin(2 * np.pi * 1000.0 * t / 48000.0) + 0.01 * rng.standard_normal(len(t))
    audio = (a + 1j * a) * 0.4
    buf = np.zeros(N, dtype=np.complex128)
    ag = lib.create_wcpagc(1, 5, 1, buf.ctypes.data, buf.ctypes.data, N, 48000, 0.001, 0.008, 4, 2.5, 1.0, 1.0, 1.0, 0.9, 0.250, 0.004, 4.0, 0, 0.500, 0.500, 2.000, 0.100)
    ys = []
    for b in range(40):
        buf[:] = audio[b * N:(b + 1) * N]; lib.xwcpagc(ag); ys.append(buf.copy())
    np.savez_compressed(os.path.join(HERE, "wdsp_fmlim_kat.npz"), y_seg=seg(y), cond=np.array([cond]), conds=conds, lim_in=audio, lim_out=np.concatenate(ys))
    print("wrote wdsp_fmlim_kat.npz: peak", np.abs(seg(y)).max(), "cond", cond)


if __name__ == "__main__":
    main()
