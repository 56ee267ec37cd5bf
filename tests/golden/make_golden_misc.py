"""tests/golden/make_golden_misc.py -- fixtures for process_agc (quisk.c:2162-2287), cFracDecim (quisk.c:622-665), NoiseBlanker (quisk.c:679-784), ssb_squelch + d_delay (quisk.c:1056-1180), dAutoNotch (quisk.c:786-963) and
the wire-format unpack loops (quisk.c:2922-2953, 3746-3763) from the compiled reference (oracle/_ref/libquisk_rx_ref.so).  Writes tests/golden/misc_kat.npz."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import quisk_oracle as O          # noqa: E402
from oracle import ref_ctypes as R            # noqa: E402

AGC_SPLITS = [480, 1, 479, 960, 333, 627]


def agc_input(n, seed):
    """Audio-like signal around 2^24 with bursts that force clipping (gain starts at 100, max_out 0.7 * 2^31)."""
    x = O.synth_iq(n, seed, 1.0) / 64.0
    x[n // 4:n // 4 + 300] *= 40.0
    x[n // 2:n // 2 + 50] *= 200.0
    return x


NB_SPLITS = [1000, 1, 2, 7, 255, 4093, 2642]
NB_CASES = [(48000, 1), (48000, 3), (192000, 2), (1536000, 1)]      # (sample rate, quisk_noise_blanker level)


def nb_input(n, seed):
    """IQ with impulse noise: single-sample spikes, a 12-sample burst, two pulses closer together than the blanking
    window, one pulse across a block boundary of NB_SPLITS, and a stretch of raised level (no pulse, the mean follows)."""
    x = O.synth_iq(n, seed, 1.0)
    rng = np.random.default_rng(seed + 1000)
    for p in (300, 999, 1000, 1003, 1600, 1640, 2900, 5357, 5358, 7000):
        x[p] *= 60.0
    x[2200:2212] *= 45.0 * (1.0 + rng.random(12))
    x[4000:4400] *= 3.0
    return x


SQ_RATE, SQ_BW = 12000, 2800
SQ_SPLITS = [120, 1, 391, 512, 513, 2000, 37] + [120] * 100 + [4093, 1500, 3000, 8000, 700]
SQ_LEVELS = [150, 60]


def sq_input(n, seed):
    """SSB audio at the 12 kS/s filter rate: band noise throughout, three voice-like tones between 0.8 s and 1.3 s."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / float(SQ_RATE)
    x = (2.0 ** 20) * rng.standard_normal(n)
    on = (t >= 0.8) & (t < 1.3)
    for f, a in ((520.0, 2.0 ** 25), (1130.0, 2.0 ** 24), (2210.0, 2.0 ** 23)):
        x += on * a * np.sin(2 * np.pi * f * t)
    return x


AN_RATE = 12000
AN_SPLITS = [120, 1, 1417, 1538, 1539, 5000, 37] + [120] * 60 + [4093, 3000, 8000, 711]
AN_SIDETONES = [0, 700]


def an_input(n, seed):
    """SSB audio with a steady carrier at 1 kHz, a weaker one at 2.1 kHz from 0.9 s on, a CW-pitch tone at 700 Hz
    (kept when it is the side tone) and noise."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / float(AN_RATE)
    x = (2.0 ** 18) * rng.standard_normal(n)
    x += (2.0 ** 24) * np.sin(2 * np.pi * 1000.0 * t + 0.3)
    x += (t >= 0.9) * (2.0 ** 23) * np.sin(2 * np.pi * 2100.0 * t)
    x += (2.0 ** 22) * np.sin(2 * np.pi * 700.0 * t + 1.1)
    return x


def ingest_bytes(seed, n):
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8)


def main():
    out = {}
    for is_cpx in (1, 0):
        lib = R.load("libquisk_rx_ref.so", private_copy=True)
        lib.ref_agc_new.restype = C.c_void_p
        lib.ref_agc_new.argtypes = [C.c_double, C.c_int]
        lib.ref_agc_set.argtypes = [C.c_double, C.c_double]
        lib.ref_agc_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.ref_agc_set(80.0, 1.0)
        st = lib.ref_agc_new(0.7, 48000)
        x = agc_input(sum(AGC_SPLITS), 60)
        if not is_cpx:
            x = x.real.astype(np.complex128)
        ys, pos = [], 0
        for n in AGC_SPLITS:
            blk = np.ascontiguousarray(x[pos:pos + n]); pos += n
            lib.ref_agc_run(st, blk.ctypes.data, n, is_cpx)
            ys.append(blk)
        out["agc_cpx%d/y" % is_cpx] = np.concatenate(ys)
    for fdecim in (1.25, 1.0416666666666667, 1.5):
        lib = R.load("libquisk_rx_ref.so", private_copy=True)
        lib.ref_cFracDecim.argtypes = [C.c_void_p, C.c_int, C.c_double]
        x = O.synth_iq(6000, 61, 1.0)
        ys, counts, pos = [], [], 0
        for n in [1000, 1, 2, 997, 4000]:
            blk = np.ascontiguousarray(x[pos:pos + n]); pos += n
            k = lib.ref_cFracDecim(blk.ctypes.data, n, fdecim)
            ys.append(blk[:k].copy()); counts.append(k)
        out["fracdecim_%g/y" % fdecim] = np.concatenate(ys)
        out["fracdecim_%g/counts" % fdecim] = np.array(counts)
    for rate, level in NB_CASES:
        lib = R.load("libquisk_rx_ref.so", private_copy=True)
        lib.ref_set_sample_rate.argtypes = [C.c_int]
        lib.ref_noise_blanker.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.ref_set_sample_rate(rate)
        x = nb_input(sum(NB_SPLITS), 90)
        ys, pos = [], 0
        for n in NB_SPLITS:
            blk = np.ascontiguousarray(x[pos:pos + n]); pos += n
            lib.ref_noise_blanker(blk.ctypes.data, n, level)
            ys.append(blk)
        out["nb_%d_%d/y" % (rate, level)] = np.concatenate(ys)
    for level in SQ_LEVELS:
        lib = R.load("libquisk_rx_ref.so", private_copy=True)
        lib.ref_ssb_squelch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        x = sq_input(sum(SQ_SPLITS), 91)
        ys, act, opn, pos = [], [], [], 0
        for n in SQ_SPLITS:
            blk = np.ascontiguousarray(x[pos:pos + n]); pos += n
            so = C.c_int(0)
            act.append(lib.ref_ssb_squelch(blk.ctypes.data, n, SQ_RATE, SQ_BW, level, 0, C.byref(so)))
            opn.append(so.value)
            ys.append(blk)
        out["sq_%d/y" % level] = np.concatenate(ys)
        out["sq_%d/active" % level] = np.array(act)
        out["sq_%d/sq_open" % level] = np.array(opn)
    for sidetone in AN_SIDETONES:
        lib = R.load("libquisk_rx_ref.so", private_copy=True)
        lib.ref_auto_notch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        x = an_input(sum(AN_SPLITS), 92)
        ys, pos = [], 0
        for k, n in enumerate(AN_SPLITS):
            blk = np.ascontiguousarray(x[pos:pos + n]); pos += n
            lib.ref_auto_notch(blk.ctypes.data, n, sidetone, AN_RATE, int(k == 0))
            ys.append(blk)
        out["notch_%d/y" % sidetone] = np.concatenate(ys)
    # wire-format ingest: the reference's own unpack loops on seeded random bytes
    lib = R.load("libquisk_rx_ref.so", private_copy=True)
    lib.ref_add_rx_samples.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]
    lib.ref_hermes_unpack.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    for nb in (1, 2, 3, 4):
        for big in (0, 1):
            data = ingest_bytes(70 + nb, 1000 * 2 * nb)
            y = np.zeros(1000, dtype=np.complex128)
            n = lib.ref_add_rx_samples(data.ctypes.data, len(data), nb, big, y.ctypes.data)
            assert n == 1000
            out["unpack_iq_%d_%d/y" % (nb, big)] = y
    for n_rx in (1, 2, 4, 10):
        pk = ingest_bytes(80 + n_rx, 3 * 1032).reshape(3, 1032)
        rows = []
        for p in range(3):
            samp = np.zeros(126, dtype=np.complex128); sub = np.zeros((max(n_rx - 1, 1), 126), dtype=np.complex128)
            one = np.ascontiguousarray(pk[p])
            n = lib.ref_hermes_unpack(one.ctypes.data, n_rx - 1, samp.ctypes.data, sub.ctypes.data)
            rows.append(np.concatenate([samp[None, :n], sub[:n_rx - 1, :n]], axis=0))
        out["unpack_hermes_%d/y" % n_rx] = np.concatenate(rows, axis=1)
    np.savez_compressed(os.path.join(HERE, "misc_kat.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
