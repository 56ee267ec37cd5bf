"""tests/golden/make_golden_wdsp_snba.py -- fixtures for WDSP's spectral noise blanker (wdsp/snb.c: create_snba / xsnba,
SetRXASNBARun) from the COMPILED REFERENCE (oracle/_ref/libwdsp_ref.so).  Stage level with create_rxa's arguments
(RXA.c:237-255) at the internal rate (12 kS/s in and out: no resamplers) and at 48 kS/s (both resamplers), on a
tones-over-noise signal with clicks of one to a dozen samples, isolated, in pairs and near each other; a flush
mid-stream; each with the reference's own sensitivity to a one-ulp change of its input.  Channel level: Quisk's channel
with SNB switched on mid-stream.  Writes tests/golden/wdsp_snba_kat.npz.   Run:  python tests/golden/make_golden_wdsp_snba.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_ctypes as R            # noqa: E402
from tests.golden.make_golden_wdsp_emnr import speechlike, ulp, rel_rms      # noqa: E402

D = C.c_double
CASES = [(12000, 64, 160), (48000, 256, 160)]           # (rate, block size, blocks)
CH_BLOCKS, CH_ON, CH_TAIL, N = 160, 40, 32, 256


def clicks(x, seed, rate):
    """impulse noise: clicks of 1 .. 12 samples (at 12 kS/s; longer in proportion at higher rates), some close together"""
    rng = np.random.default_rng(seed)
    y = x.copy()
    scale = rate // 12000
    pos = 3000 * scale
    while pos < len(y) - 400 * scale:
        ln = int(rng.integers(1, 13)) * scale
        y[pos:pos + ln] += rng.choice([-1.0, 1.0]) * rng.uniform(1.0, 3.0) * np.hanning(ln + 2)[1:-1] if ln > 2 else rng.uniform(1.5, 3.0)
        pos += int(rng.integers(40, 900)) * scale
    return y


def stage_input(rate, bsize, blocks):
    return clicks(speechlike(bsize * blocks, 61, rate), 62, rate)


def channel_input():
    xr = clicks(speechlike(N * CH_BLOCKS, 71, 48000), 72, 48000)
    return (xr + 1j * np.roll(xr, 3)).astype(np.complex128)


def main():
    lib = R.load("libwdsp_ref.so")
    lib.create_snba.restype = C.c_void_p
    lib.create_snba.argtypes = [C.c_int, C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [D, D, C.c_int, C.c_int, C.c_int, D, D, D]
    lib.xsnba.argtypes = [C.c_void_p]
    lib.flush_snba.argtypes = [C.c_void_p]
    out = {}

    def stage(x, rate, bsize, blocks, flush_at):
        buf = np.zeros(bsize, dtype=np.complex128)
        d = lib.create_snba(1, buf.ctypes.data, buf.ctypes.data, rate, 12000, bsize, 4, 256, 64, 2, 8.0, 20.0, 10, 2, 2, 0.5, 200.0, 5400.0)
        ys = []
        for b in range(blocks):
            if b == flush_at:
                lib.flush_snba(d)
            buf[:] = x[b * bsize:(b + 1) * bsize] + 0.25j * x[b * bsize:(b + 1) * bsize]
            lib.xsnba(d)
            ys.append(buf.copy())
        return np.concatenate(ys)

    for rate, bsize, blocks in CASES:
        x = stage_input(rate, bsize, blocks)
        y = stage(x, rate, bsize, blocks, 120)
        assert not y.imag.any()
        key = "snba_%d" % rate
        out[key + "/y"] = y.real.copy()
        yp = stage(ulp(x, 9), rate, bsize, blocks, 120)
        out[key + "/cond"] = np.array([rel_rms(yp, y)])
        # how much the blanker did: the same stream through a stage that never detects (k2 huge) is the delayed input
        print(key, "peak in", np.abs(x).max(), "peak out", np.abs(y).max(), "cond", out[key + "/cond"])

    lib.OpenChannel.argtypes = [C.c_int] * 8 + [D] * 4 + [C.c_int]
    lib.fexchange0.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.RXASetPassband.argtypes = [C.c_int, D, D]
    lib.SetRXAAGCFixed.argtypes = [C.c_int, D]

    def channel(xc, chn):
        lib.OpenChannel(chn, N, N, 48000, 48000, 48000, 0, 1, D(0.0), D(0.0), D(0.0), D(0.0), 1)
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
                lib.SetRXASNBARun(chn, 1)
            inb[:] = xc[b * N:(b + 1) * N]
            lib.fexchange0(chn, inb.ctypes.data_as(C.c_void_p), outb.ctypes.data_as(C.c_void_p), C.byref(err))
            ys.append(outb.copy())
            time.sleep(0.004)
        lib.SetChannelState(chn, 0, 1)
        lib.CloseChannel(chn)
        return np.concatenate(ys)

    xc = channel_input()
    yc = channel(xc, 7)
    out["chan/y_tail"] = yc[-CH_TAIL * N:]
    ycp = channel(ulp(xc.view(np.float64), 13).view(np.complex128), 8)
    out["chan/cond"] = np.array([rel_rms(ycp[-CH_TAIL * N:], yc[-CH_TAIL * N:])])
    print("channel tail peak", np.abs(out["chan/y_tail"]).max(), "cond", out["chan/cond"])
    np.savez_compressed(os.path.join(HERE, "wdsp_snba_kat.npz"), **out)
    print("wrote wdsp_snba_kat.npz")


if __name__ == "__main__":
    main()
