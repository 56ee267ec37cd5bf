"""tests/golden/make_golden_wdsp.py -- known-answer fixtures for the WDSP RXA stages, generated from the
COMPILED REFERENCE (oracle/_ref/libwdsp_ref.so = wdsp/*.c + our FFTW-API shim, built by
oracle/build_ref.sh).  WDSP ships no tests; these are outputs of its own code on seeded inputs.
Writes tests/golden/wdsp_kat.npz.   Run:  python tests/golden/make_golden_wdsp.py
"""
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
VP = C.c_void_p


def wdsp():
    lib = R.load("libwdsp_ref.so")
    lib.fir_bandpass.restype = C.POINTER(D)
    lib.fir_bandpass.argtypes = [C.c_int, D, D, D, C.c_int, C.c_int, D]
    lib.create_fircore.restype = VP
    lib.create_fircore.argtypes = [C.c_int, VP, VP, C.c_int, C.c_int, VP]
    lib.xfircore.argtypes = [VP]
    lib.setImpulse_fircore.argtypes = [VP, VP, C.c_int]
    lib.create_resample.restype = VP
    lib.create_resample.argtypes = [C.c_int, C.c_int, VP, VP, C.c_int, C.c_int, D, C.c_int, D]
    lib.xresample.argtypes = [VP]
    lib.xresample.restype = C.c_int
    lib.create_shift.restype = VP
    lib.create_shift.argtypes = [C.c_int, C.c_int, VP, VP, C.c_int, D]
    lib.xshift.argtypes = [VP]
    lib.create_wcpagc.restype = VP
    lib.create_wcpagc.argtypes = [C.c_int, C.c_int, C.c_int, VP, VP, C.c_int, C.c_int, D, D, C.c_int] + [D] * 8 + [C.c_int] + [D] * 4
    lib.xwcpagc.argtypes = [VP]
    lib.create_amd.restype = VP
    lib.create_amd.argtypes = [C.c_int, C.c_int, VP, VP, C.c_int, C.c_int, C.c_int, C.c_int, D, D, D, D, D, D]
    lib.xamd.argtypes = [VP]
    lib.create_fmd.restype = VP
    lib.create_fmd.argtypes = [C.c_int, C.c_int, VP, VP, C.c_int] + [D] * 9 + [C.c_int, D, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.xfmd.argtypes = [VP]
    lib.OpenChannel.argtypes = [C.c_int] * 8 + [D] * 4 + [C.c_int]
    lib.fexchange0.argtypes = [C.c_int, VP, VP, VP]
    lib.SetRXAMode.argtypes = [C.c_int, C.c_int]
    lib.RXASetPassband.argtypes = [C.c_int, D, D]
    lib.RXASetNC.argtypes = [C.c_int, C.c_int]
    lib.SetRXAAGCMode.argtypes = [C.c_int, C.c_int]
    lib.SetRXAShiftRun.argtypes = [C.c_int, C.c_int]
    lib.RXAGetaSipF1.argtypes = [C.c_int, VP, C.c_int]
    lib.RXANBPSetNotchesRun.argtypes = [C.c_int, C.c_int]
    lib.RXANBPSetTuneFrequency.argtypes = [C.c_int, D]
    lib.RXANBPAddNotch.argtypes = [C.c_int, C.c_int, D, D, C.c_int]
    lib.SetChannelState.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.CloseChannel.argtypes = [C.c_int]
    lib.GetRXAMeter.argtypes = [C.c_int, C.c_int]
    lib.GetRXAMeter.restype = D
    lib.SetRXAAGCFixed.argtypes = [C.c_int, D]
    lib.RXANBPSetRun.argtypes = [C.c_int, C.c_int]
    lib.SetRXAShiftFreq.argtypes = [C.c_int, D]
    return lib


def ulp_perturb(x, seed):
    """x with every component moved by one ulp up or down (seeded): the smallest change of the input there is."""
    rng = np.random.default_rng(seed)
    v = np.ascontiguousarray(x).view(np.float64).copy()
    up = rng.integers(0, 2, size=v.shape).astype(bool)
    v = np.where(up, np.nextafter(v, np.inf), np.nextafter(v, -np.inf))
    return v.view(x.dtype) if np.iscomplexobj(x) else v


def rel_rms(a, b):
    return float(np.sqrt(np.mean(np.abs(a - b) ** 2)) / np.sqrt(np.mean(np.abs(b) ** 2)))


from quisk_b200.synth import sig, fm_sig, am_sig      # noqa: E402,F401  (re-exported for the tests)


def bandpass(lib, N, fl, fh, rate, wintype, rtype, scale):
    p = lib.fir_bandpass(N, fl, fh, rate, wintype, rtype, scale)
    return np.ctypeslib.as_array(p, (N * (2 if rtype else 1),)).copy()


FM_BLOCKS = 176
QUISK_SPLITS = [100, 156, 1, 255, 256, 257, 1000, 300, 513, 3, 767, 1100, 1292]       # sum 6000
FIRCORE_CASES = [(64, 256, 48000.0), (256, 1024, 48000.0), (1024, 4096, 192000.0), (256, 256, 48000.0)]
RESAMPLE_CASES = [(384000, 48000, [2048, 2048, 1000, 24]), (192000, 48000, [1024, 7, 1017]), (48000, 192000, [256, 255]),
                  (44100, 48000, [4096, 441])]


def main():
    lib = wdsp()
    out = {}
    # ---- fircore ----
    for size, nc, rate in FIRCORE_CASES:
        imp = bandpass(lib, nc, 150.0, 2850.0, rate, 0, 1, 1.0 / (2 * size))
        imp2 = bandpass(lib, nc, -2850.0, -150.0, rate, 0, 1, 1.0 / (2 * size))
        inb = np.zeros(size, dtype=np.complex128); outb = np.zeros(2 * size, dtype=np.complex128)
        f = lib.create_fircore(size, inb.ctypes.data, outb.ctypes.data, nc, 0, imp.ctypes.data)
        x = sig(size * 8, 100 + size, rate)
        ys = []
        for b in range(8):
            if b == 5:
                lib.setImpulse_fircore(f, imp2.ctypes.data, 1)      # retune mid-stream
            inb[:] = x[b * size:(b + 1) * size]
            lib.xfircore(f)
            ys.append(outb[:size].copy())
        out["fircore_%d_%d/y" % (size, nc)] = np.concatenate(ys)
    # ---- fircore with minimum-phase masks (mp = 1: calc_fircore runs the impulse through mp_imp, firmin.c:327-328) ----
    size, nc, rate = 256, 1024, 48000.0
    imp = bandpass(lib, nc, 150.0, 2850.0, rate, 0, 1, 1.0 / (2 * size))
    inb = np.zeros(size, dtype=np.complex128); outb = np.zeros(2 * size, dtype=np.complex128)
    f = lib.create_fircore(size, inb.ctypes.data, outb.ctypes.data, nc, 1, imp.ctypes.data)
    x = sig(size * 8, 150, rate)
    ys = []
    for b in range(8):
        inb[:] = x[b * size:(b + 1) * size]
        lib.xfircore(f)
        ys.append(outb[:size].copy())
    out["fircore_mp_256_1024/y"] = np.concatenate(ys)
    # mp_imp's conditioning (fir.c:317-368: log|H| of stop-band bins at the rounding floor): the reference's own minimum-phase
    # fircore with the IMPULSE moved by one ulp per tap
    impp = ulp_perturb(imp, 7)
    inb = np.zeros(size, dtype=np.complex128); outb = np.zeros(2 * size, dtype=np.complex128)
    f = lib.create_fircore(size, inb.ctypes.data, outb.ctypes.data, nc, 1, impp.ctypes.data)
    ys = []
    for b in range(8):
        inb[:] = x[b * size:(b + 1) * size]
        lib.xfircore(f)
        ys.append(outb[:size].copy())
    out["fircore_mp_256_1024/cond"] = np.array([rel_rms(np.concatenate(ys), out["fircore_mp_256_1024/y"])])
    # ---- resample ----
    for in_rate, out_rate, splits in RESAMPLE_CASES:
        x = sig(sum(splits), 200, in_rate)
        ys, counts, pos = [], [], 0
        r = None
        cap = max(splits) * max(1, out_rate // in_rate + 1) + 16
        for n in splits:
            inb = np.ascontiguousarray(x[pos:pos + n]); pos += n
            outb = np.zeros(cap * 2, dtype=np.complex128)
            # the reference object is tied to one block size: re-point its buffers per block
            if r is None:
                r = lib.create_resample(1, n, inb.ctypes.data, outb.ctypes.data, in_rate, out_rate, 0.0, 0, 1.0)
            lib.setBuffers_resample.argtypes = [VP, VP, VP]; lib.setBuffers_resample(r, inb.ctypes.data, outb.ctypes.data)
            lib.setSize_resample_keep = None
            # `size` is the first int after `run` in struct _resample (resample.h): run, size
            C.cast(r, C.POINTER(C.c_int))[1] = n
            k = lib.xresample(r)
            ys.append(outb[:k].copy()); counts.append(k)
        out["resample_%d_%d/y" % (in_rate, out_rate)] = np.concatenate(ys)
        out["resample_%d_%d/counts" % (in_rate, out_rate)] = np.array(counts)
    # ---- shift ----
    n = 1024
    x = sig(3 * n, 300, 48000.0)
    inb = np.zeros(n, dtype=np.complex128)
    s = lib.create_shift(1, n, inb.ctypes.data, inb.ctypes.data, 48000, 1234.5)
    ys = []
    for b in range(3):
        inb[:] = x[b * n:(b + 1) * n]; lib.xshift(s); ys.append(inb.copy())
    out["shift/y"] = np.concatenate(ys)
    # ---- wcpagc: create_rxa's parameters with the SetRXAAGCMode presets ----
    for mode, hang_thresh, hangtime, tau_decay in [(3, 1.0, 0.0, 0.250), (1, 0.250, 2.0, 2.0), (4, 1.0, 0.0, 0.050)]:
        n = 1024; rate = 192000 if mode == 3 else 48000
        x = sig(8 * n, 400 + mode, float(rate))
        x[2 * n:3 * n] *= 3.0; x[4 * n:6 * n] *= 0.05; x[6 * n:] *= 2.0      # level steps drive the state machine
        inb = np.zeros(n, dtype=np.complex128)
        a = lib.create_wcpagc(1, mode, 1, inb.ctypes.data, inb.ctypes.data, n, rate, 0.001, tau_decay, 4, 10000.0, 1.5, 1000.0,
                              1.0, 1.0, 0.250, 0.005, 5.0, 1, 0.500, hangtime, hang_thresh, 0.100)
        ys = []
        for b in range(8):
            inb[:] = x[b * n:(b + 1) * n]; lib.xwcpagc(a); ys.append(inb.copy())
        out["wcpagc_mode%d/y" % mode] = np.concatenate(ys)
    # ---- amd ----
    for mode, sb in [(0, 0), (1, 0), (1, 1), (1, 2)]:
        n = 512
        x = am_sig(4 * n, 500, 48000.0)
        inb = np.zeros(n, dtype=np.complex128)
        a = lib.create_amd(1, n, inb.ctypes.data, inb.ctypes.data, mode, 1, sb, 48000, -2000.0, 2000.0, 1.0, 250.0, 0.02, 1.4)
        ys = []
        for b in range(4):
            inb[:] = x[b * n:(b + 1) * n]; lib.xamd(a); ys.append(inb.copy())
        out["amd_%d_%d/y" % (mode, sb)] = np.concatenate(ys)
        xp = ulp_perturb(x, 4)
        inb = np.zeros(n, dtype=np.complex128)
        a = lib.create_amd(1, n, inb.ctypes.data, inb.ctypes.data, mode, 1, sb, 48000, -2000.0, 2000.0, 1.0, 250.0, 0.02, 1.4)
        ys = []
        for b in range(4):
            inb[:] = xp[b * n:(b + 1) * n]; lib.xamd(a); ys.append(inb.copy())
        out["amd_%d_%d/cond" % (mode, sb)] = np.array([rel_rms(np.concatenate(ys), out["amd_%d_%d/y" % (mode, sb)])])
    # ---- fmd (create_rxa's arguments) ----
    n = 256
    x = fm_sig(12 * n, 600, 48000.0)
    inb = np.zeros(2 * n, dtype=np.complex128)      # fircore writes 2*size samples into its out buffer (firmin.c:318)
    f = lib.create_fmd(1, n, inb.ctypes.data, inb.ctypes.data, 48000, 5000.0, 300.0, 3000.0, -8000.0, 8000.0, 1.0, 20000.0, 0.02, 0.5,
                       1, 254.1, 2048, 0, 2048, 0)
    ys = []
    for b in range(12):
        inb[:n] = x[b * n:(b + 1) * n]; lib.xfmd(f); ys.append(inb[:n].copy())
    out["fmd/y"] = np.concatenate(ys)
    # the reference's own sensitivity to a one-ulp change of its input (conditioning of the PLL + filters)
    xp = ulp_perturb(x, 3)
    inb = np.zeros(2 * n, dtype=np.complex128)
    f = lib.create_fmd(1, n, inb.ctypes.data, inb.ctypes.data, 48000, 5000.0, 300.0, 3000.0, -8000.0, 8000.0, 1.0, 20000.0, 0.02, 0.5,
                       1, 254.1, 2048, 0, 2048, 0)
    ys = []
    for b in range(12):
        inb[:n] = xp[b * n:(b + 1) * n]; lib.xfmd(f); ys.append(inb[:n].copy())
    out["fmd/cond"] = np.array([rel_rms(np.concatenate(ys), out["fmd/y"])])
    # ---- the whole channel through OpenChannel + fexchange0 (blocking output, zero slew times) ----
    siphons = {}

    meters = {}

    def run_channel(ch, in_size, dsp_size, in_rate, dsp_rate, out_rate, setup, x, nblocks, slew=(0.0, 0.0, 0.0, 0.0), before=None):
        lib.OpenChannel(ch, in_size, dsp_size, in_rate, dsp_rate, out_rate, 0, 1, slew[0], slew[1], slew[2], slew[3], 1)
        setup(ch)
        out_size = in_size * out_rate // in_rate if out_rate <= in_rate else in_size * (out_rate // in_rate)
        err = C.c_int(0)
        ys = []
        mt = []
        for b in range(nblocks):
            if before is not None:
                before(ch, b)
            inb = np.ascontiguousarray(x[b * in_size:(b + 1) * in_size])
            outb = np.full(out_size, -7.0 - 7.0j, dtype=np.complex128)      # a call with the exchange off leaves `out` alone
            lib.fexchange0(ch, inb.ctypes.data, outb.ctypes.data, C.byref(err))
            assert err.value == 0
            ys.append(outb)
            # Pace the caller like a sound card would.  Called back to back, the reference's exchange
            # (one block of slack from Sem_OutReady's initial credit, iobuffs.c:409-416) lets the caller
            # run ahead of the DSP thread and its output becomes timing dependent -- two identical runs
            # differ.  With pacing it is deterministic and equals the composition of its own stages.
            time.sleep(0.004 * max(1, in_size // 256))
            mt.append([lib.GetRXAMeter(ch, k) for k in range(7)])
        meters[ch] = np.array(mt)
        sip = np.zeros(2 * 1024, dtype=np.float32)      # RXAGetaSipF1: the newest 1024 samples of midbuff as floats
        lib.RXAGetaSipF1(ch, sip.ctypes.data, 1024)
        siphons[ch] = sip
        return np.concatenate(ys)

    def setup_usb(ch):          # the C3 concretisation of SURVEY.md 8(d), scaled down
        lib.SetRXAShiftRun(ch, 0)
        lib.RXASetNC(ch, 2048)
        lib.SetRXAMode(ch, 1)
        lib.RXASetPassband(ch, 150.0, 2850.0)
        lib.SetRXAAGCMode(ch, 3)
    x = sig(256 * 24, 700, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
    out["rxa_usb/y"] = run_channel(0, 256, 256, 48000, 48000, 48000, setup_usb, x, 24)
    out["rxa_usb/sip"] = siphons[0]
    out["rxa_usb/meters"] = meters[0]        # GetRXAMeter(ch, 0..6) after every block (meter.c:75-129, RXA.h:47-57)
    # the reference's own sensitivity: the same channel on the same input moved by one ulp per component
    out["rxa_usb/cond"] = np.array([rel_rms(run_channel(10, 256, 256, 48000, 48000, 48000, setup_usb, ulp_perturb(x, 1), 24), out["rxa_usb/y"])])

    # ---- C3 at its stated geometry: OpenChannel(1024, 1024, 192 k), RXASetNC(4096), USB 150-2850, AGC 3 ----
    def setup_c3(ch):
        lib.SetRXAShiftRun(ch, 0)
        lib.RXASetNC(ch, 4096)
        lib.SetRXAMode(ch, 1)
        lib.RXASetPassband(ch, 150.0, 2850.0)
        lib.SetRXAAGCMode(ch, 3)
    x3 = sig(1024 * 16, 710, 192000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2), (-30000.0, 0.25)))
    out["rxa_c3/y"] = run_channel(11, 1024, 1024, 192000, 192000, 192000, setup_c3, x3, 16)
    out["rxa_c3/meters"] = meters[11]
    out["rxa_c3/cond"] = np.array([rel_rms(run_channel(12, 1024, 1024, 192000, 192000, 192000, setup_c3, ulp_perturb(x3, 2), 16), out["rxa_c3/y"])])

    # ---- in_size != dsp_insize (iobuffs.c:385-420, 583-604): several calls per DSP turn, and several turns per call ----
    xr = sig(256 * 24, 720, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
    out["rxa_reblock_64_256/y"] = run_channel(13, 64, 256, 48000, 48000, 48000, setup_usb, xr, 96)
    out["rxa_reblock_1024_256/y"] = run_channel(14, 1024, 256, 48000, 48000, 48000, setup_usb, xr, 6)

    # ---- SetChannelState (channel.c:262-300): down-slew + flush mid-stream, calls while the exchange is off, restart ----
    # Geometry chosen so that the reference is deterministic here: the fexchange0 call that finishes the down-slew hands
    # the channel to the flush thread (iobuffs.c:494-498, channel.c:134-155); if that same call also completed a DSP
    # block, the DSP thread and the flush thread race for csDSP and the last block is either processed or dropped (its
    # output is discarded either way, but the AGC state that flush_wcpagc keeps differs: 5e-4 on the gain after the
    # restart).  With in_size 64 / dsp_size 256 and the stop issued on a call that starts a DSP block, the ramp
    # (1 + 481 + 65 samples = 9 calls) ends on a call that does not trigger the DSP thread.
    def stop_start(ch, b):
        if b == 40:
            lib.SetChannelState(ch, 0, 0)
        if b == 72:
            assert lib.SetChannelState(ch, 1, 0) == 0
    xs2 = sig(256 * 32, 730, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
    out["rxa_stop_start/y"] = run_channel(15, 64, 256, 48000, 48000, 48000, setup_usb, xs2, 128, slew=(0.010, 0.025, 0.0, 0.010), before=stop_start)
    out["rxa_stop_start/meters"] = meters[15]

    # ---- Quisk's own open sequence (quisk_wdsp.py:69-99) behind its own re-blocker wdspFexchange0 (quisk_wdsp.c:22-69) ----
    glue = R.load("libquisk_wdspglue_ref.so")
    glue.ref_wdsp_set_parameter.argtypes = [C.c_int, C.c_int, VP, C.c_int]
    glue.wdspFexchange0.argtypes = [C.c_int, VP, C.c_int]

    def setup_quisk(ch):
        lib.SetRXAShiftRun(ch, 0)
        lib.RXANBPSetRun(ch, 0)
        lib.SetRXAAMSQRun(ch, 0)
        lib.SetRXAMode(ch, 1)
        lib.RXASetPassband(ch, 300.0, 3000.0)
        lib.RXASetNC(ch, 256)
        lib.RXASetMP(ch, 0)
        lib.SetRXAAGCMode(ch, 0)
        lib.SetRXAAGCFixed(ch, 0.0)
        lib.SetRXAPanelRun(ch, 0)
        lib.SetRXAEMNRRun(ch, 0)
    lib.OpenChannel(16, 256, 256, 48000, 48000, 48000, 0, 1, 0.010, 0.025, 0.0, 0.010, 1)
    setup_quisk(16)
    glue.ref_wdsp_set_parameter(16, 256, C.cast(lib.fexchange0, VP), 1)
    xq = sig(6000, 740, 48000.0) * 2.0 ** 30            # Quisk's samples are scaled to CLIP32
    ys, counts, pos = [], [], 0
    for n in QUISK_SPLITS:
        buf = np.zeros(n + 1024, dtype=np.complex128); buf[:n] = xq[pos:pos + n]; pos += n
        k = glue.wdspFexchange0(16, buf.ctypes.data, n)
        ys.append(buf[:k].copy()); counts.append(k)
        time.sleep(0.004 * (1 + n // 256))
    out["quisk_reblock/y"] = np.concatenate(ys)
    out["quisk_reblock/counts"] = np.array(counts)

    def setup_usb_notch(ch):    # the notch database in action: two notches inside the pass band, one of them on a tone
        setup_usb(ch)
        lib.RXANBPSetTuneFrequency(ch, 7000000.0)
        lib.RXANBPSetNotchesRun(ch, 1)
        assert lib.RXANBPAddNotch(ch, 0, 7001500.0, 400.0, 1) == 0
        assert lib.RXANBPAddNotch(ch, 1, 7002400.0, 100.0, 1) == 0        # narrower than the minimum: auto-widened
    xn = sig(256 * 24, 700, 48000.0, tones=((1000.0, 0.3), (1500.0, 0.2), (2200.0, 0.1)))
    out["rxa_usb_notch/y"] = run_channel(6, 256, 256, 48000, 48000, 48000, setup_usb_notch, xn, 24)
    # the same channel opened the way Quisk opens it (quisk_wdsp.py:79-80): 10 ms of zeros after the first non-zero
    # sample, then a 25 ms raised-cosine ramp (upslew0, iobuffs.c:98-160); the stream starts with 100 zero samples
    xs = x.copy(); xs[:100] = 0.0
    out["rxa_usb_slew/y"] = run_channel(5, 256, 256, 48000, 48000, 48000, setup_usb, xs, 24, slew=(0.010, 0.025, 0.0, 0.010))

    def setup_default(ch):      # no mode set: bp1 still runs
        lib.SetRXAShiftRun(ch, 0)
    x = sig(256 * 16, 701, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
    out["rxa_default/y"] = run_channel(1, 256, 256, 48000, 48000, 48000, setup_default, x, 16)

    def setup_fm(ch):           # the C4 concretisation: 384 k in, 48 k dsp/out, FM
        lib.SetRXAShiftRun(ch, 0)
        lib.SetRXAMode(ch, 5)
        lib.RXASetPassband(ch, -8000.0, 8000.0)
    # The PLL starts on FFT rounding noise (nbp0 delays the signal by 1024 samples), so the first
    # ~10^4 output samples depend on the FFT library's last bits -- in the reference too.  Parity is
    # therefore taken on the tail of a long run, after the loop has locked and the DC estimate settled.
    x = fm_sig(2048 * FM_BLOCKS, 702, 384000.0)
    out["rxa_fm/y_tail"] = run_channel(2, 2048, 256, 384000, 48000, 48000, setup_fm, x, FM_BLOCKS)[-16 * 256:]
    out["rxa_fm/cond"] = np.array([rel_rms(run_channel(17, 2048, 256, 384000, 48000, 48000, setup_fm, ulp_perturb(x, 5), FM_BLOCKS)[-16 * 256:], out["rxa_fm/y_tail"])])

    def setup_am(ch):
        lib.SetRXAShiftRun(ch, 0)
        lib.SetRXAMode(ch, 6)
        lib.RXASetPassband(ch, -4000.0, 4000.0)
    x = am_sig(256 * 16, 703, 48000.0)
    out["rxa_am/y"] = run_channel(3, 256, 256, 48000, 48000, 48000, setup_am, x, 16)
    out["rxa_am/cond"] = np.array([rel_rms(run_channel(18, 256, 256, 48000, 48000, 48000, setup_am, ulp_perturb(x, 6), 16), out["rxa_am/y"])])
    np.savez_compressed(os.path.join(HERE, "wdsp_kat.npz"), **out)
    print("wrote wdsp_kat.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
