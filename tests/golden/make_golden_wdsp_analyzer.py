"""tests/golden/make_golden_wdsp_analyzer.py -- fixtures for WDSP's spectrum engine (wdsp/analyzer.c: XCreateAnalyzer,
SetAnalyzer, SetDisplay*, Spectrum0, GetPixels) from the COMPILED REFERENCE (oracle/_ref/libwdsp_ref.so), complex input,
one LO, one sub-span: every window type, every detector (peak, rosenfell, average, sample, rms), every averaging mode
(peak hold, none, recursive linear, window, recursive log), overlap, bin clipping at the span ends (integer and fractional),
a flipped LO, more pixels than bins (the interpolating branch), normalisation to 1 Hz.  The reference runs its transforms on
worker threads; the generator feeds one hop of samples at a time and waits for that frame's pixels, so the sequence of
frames is deterministic.  Writes tests/golden/wdsp_analyzer_kat.npz.   Run:  python tests/golden/make_golden_wdsp_analyzer.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

D = C.c_double
FRAMES = 10
# name: size, hop (= samples per Spectrum0 call), window, pi_alpha, clip, fscL, fscH, pixels, flip, rate,
#       [(detector, average mode, num_average, backmult, normalize)] per pixel output
CASES = {
    "hann4096": dict(sz=4096, hop=4096, win=2, pa=0.0, clip=0, fl=0.0, fh=0.0, npix=1024, flip=0, rate=192000,
                     outs=[(0, 0, 1, 0.0, 0), (1, 1, 1, 0.8, 0), (2, 2, 4, 0.0, 0), (3, 3, 1, 0.7, 1)]),
    "kaiser2048": dict(sz=2048, hop=1024, win=5, pa=14.0, clip=16, fl=10.5, fh=3.25, npix=700, flip=1, rate=48000,
                       outs=[(4, -1, 1, 0.0, 0), (2, 2, 3, 0.0, 1), (0, 1, 1, 0.9, 0), (1, 0, 1, 0.0, 0)]),
    "bh4_1024_interp": dict(sz=1024, hop=512, win=1, pa=0.0, clip=0, fl=0.0, fh=0.0, npix=4000, flip=0, rate=96000,
                            outs=[(0, 0, 1, 0.0, 0), (2, 3, 1, 0.6, 0)]),
    "bh7_8192": dict(sz=8192, hop=8192, win=6, pa=0.0, clip=100, fl=7.0, fh=0.5, npix=2048, flip=0, rate=1536000,
                     outs=[(1, 0, 1, 0.0, 0), (3, 1, 1, 0.5, 0), (4, 2, 2, 0.0, 0)]),
    "flat512": dict(sz=512, hop=128, win=3, pa=0.0, clip=3, fl=0.0, fh=0.0, npix=505, flip=0, rate=48000,
                    outs=[(2, 0, 1, 0.0, 0), (0, 0, 1, 0.0, 0)]),
    "hamming_rect256": dict(sz=256, hop=256, win=4, pa=0.0, clip=0, fl=2.75, fh=0.0, npix=100, flip=1, rate=48000,
                            outs=[(3, 0, 1, 0.0, 0), (1, 3, 1, 0.3, 0)]),
    # stitched spans: three sub-spans of which the low span clip removes the first one and a part of the second; two with overlap
    "stitch3_skip": dict(sz=1024, hop=1024, win=2, pa=0.0, clip=20, fl=1100.5, fh=30.25, npix=900, flip=0, rate=192000, stitch=3,
                         outs=[(0, 0, 1, 0.0, 0), (2, 1, 1, 0.7, 0)]),
    # real input (typ = 0): the I rail alone, bins 0 .. size / 2, eliminate instead of Celiminate
    "real2048": dict(sz=2048, hop=1024, win=2, pa=0.0, clip=5, fl=3.5, fh=2.0, npix=600, flip=0, rate=48000, typ=0,
                     outs=[(0, 0, 1, 0.0, 0), (2, 1, 1, 0.6, 0), (1, 0, 1, 0.0, 0)]),
    "real512_flip_stitch2": dict(sz=512, hop=512, win=4, pa=0.0, clip=2, fl=0.0, fh=0.0, npix=1000, flip=1, rate=48000, typ=0, stitch=2,
                                 outs=[(3, 0, 1, 0.0, 0), (0, 3, 1, 0.5, 0)]),
    "stitch2_overlap": dict(sz=512, hop=256, win=1, pa=0.0, clip=8, fl=0.0, fh=0.0, npix=1200, flip=0, rate=96000, stitch=2,
                            outs=[(1, 0, 1, 0.0, 0), (4, 2, 3, 0.0, 0)]),
}


def analyzer_input(name, n, ss=0):
    """tones over noise with a level step, different per case and sub-span"""
    seed = sum(map(ord, name)) + 1000 * ss
    rng = np.random.default_rng(seed)
    t = np.arange(n)
    x = 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for f, a in ((0.0131, 0.5), (-0.21, 0.2), (0.3777, 0.05), (-0.4402, 0.8)):
        x += a * np.exp(2j * np.pi * (f + 0.017 * ss) * t)
    x[n // 2:] *= 0.3
    return x.astype(np.complex128)


def frames_expected(cfg, calls):
    """how many frames the first `calls` Spectrum0 calls complete (analyzer.c:884-911, 1561-1570)"""
    have, frames = 0, 0
    out = []
    for _ in range(calls):
        have += cfg["hop"]
        k = 0
        while have >= cfg["sz"]:
            have -= cfg["hop"]
            k += 1
        frames += k
        out.append(k)
    return out


def bind(lib):
    I = C.c_int
    lib.XCreateAnalyzer.argtypes = [I, C.POINTER(I), I, I, I, C.c_char_p]
    lib.SetAnalyzer.argtypes = [I, I, I, I, C.POINTER(I), I, I, I, D, I, I, D, D, I, I, I, D, D, I]
    lib.SetDisplayAvBackmult.argtypes = [I, I, D]
    lib.Spectrum0.argtypes = [I, I, I, I, C.c_void_p]
    lib.GetPixels.argtypes = [I, I, C.c_void_p, C.POINTER(I)]
    return lib


def run_case(lib, disp, name, cfg):
    """one configuration through the reference's own entry points; returns {key: [FRAMES][pixels] float32} per pixel output"""
    I = C.c_int
    ok = I(1)
    nst = cfg.get("stitch", 1)
    lib.XCreateAnalyzer(disp, C.byref(ok), 16384, 1, 4, b"")
    assert ok.value == 0
    lib.SetDisplaySampleRate(disp, cfg["rate"])
    for po, (det, av, num, back, norm) in enumerate(cfg["outs"]):
        lib.SetDisplayDetectorMode(disp, po, det)
        lib.SetDisplayAverageMode(disp, po, av)
        lib.SetDisplayNumAverage(disp, po, num)
        lib.SetDisplayAvBackmult(disp, po, D(back))
        lib.SetDisplayNormOneHz(disp, po, norm)
    flip = (I * 1)(cfg["flip"])
    lib.SetAnalyzer(disp, len(cfg["outs"]), 1, cfg.get("typ", 1), flip, cfg["sz"], cfg["hop"], cfg["win"], D(cfg["pa"]), cfg["sz"] - cfg["hop"], cfg["clip"],
                    D(cfg["fl"]), D(cfg["fh"]), cfg["npix"], nst, 0, D(0.0), D(0.0), 2 * cfg["sz"])
    calls = cfg["sz"] // cfg["hop"] - 1 + FRAMES
    xs = [analyzer_input(name, calls * cfg["hop"], ss) for ss in range(nst)]
    per_call = frames_expected(cfg, calls)
    pix = [[] for _ in cfg["outs"]]
    buf = np.zeros(cfg["hop"], dtype=np.complex128)
    for k in range(calls):
        for ss in range(nst):       # one hop into every sub-span; the line is stitched when the last one has reported
            buf[:] = xs[ss][k * cfg["hop"]:(k + 1) * cfg["hop"]]
            lib.Spectrum0(1, disp, ss, 0, buf.ctypes.data_as(C.c_void_p))
            if ss + 1 < nst and per_call[k]:
                time.sleep(0.02)    # let the dispatcher send this sub-span's frame before the next one fills
        assert per_call[k] in (0, 1)
        if per_call[k]:
            for po in range(len(cfg["outs"])):
                p = np.zeros(cfg["npix"], dtype=np.float32)
                flag = I(0)
                t0 = time.time()
                while not flag.value:
                    lib.GetPixels(disp, po, p.ctypes.data_as(C.c_void_p), C.byref(flag))
                    if not flag.value:
                        time.sleep(0.002)
                        assert time.time() - t0 < 20, "the reference produced no frame"
                pix[po].append(p)
            time.sleep(0.01)
    out = {}
    for po in range(len(cfg["outs"])):
        assert len(pix[po]) == FRAMES
        out["%s/pix%d" % (name, po)] = np.stack(pix[po])
    return out


def main():
    from oracle import ref_ctypes as R
    lib = bind(R.load("libwdsp_ref.so"))
    out = {}
    for disp, (name, cfg) in enumerate(CASES.items()):
        res = run_case(lib, disp, name, cfg)
        out.update(res)
        print(name, [float(np.nanmax(v)) for v in res.values()], [bool(np.isfinite(v).all()) for v in res.values()])
    np.savez_compressed(os.path.join(HERE, "wdsp_analyzer_kat.npz"), **out)
    print("wrote wdsp_analyzer_kat.npz", os.path.getsize(os.path.join(HERE, "wdsp_analyzer_kat.npz")))


if __name__ == "__main__":
    main()
