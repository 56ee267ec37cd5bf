"""tests/golden/make_golden.py -- regenerates the committed fixtures from the COMPILED
REFERENCE (oracle/_ref, built by oracle/build_ref.sh from /root/reference).

The reference ships no tests or golden vectors, so these files are
the known-answer set for this repository:

  quisk_b200/data/quisk_tables.npz (package data, the product reads it)
                      the 30 coefficient tables of filters.h (read out of the compiled
                      filter.c) and the 17 prototype low-pass tables of filters.py
  filter_kat.npz      every filter.h block function on seeded synthetic IQ fed in
                      uneven blocks: outputs + per-block counts
  chain_kat.npz       quisk_process_decimate at several sample rates and
                      quisk_process_demodulate in six modes: outputs + per-block counts

Inputs are regenerated from oracle.quisk_oracle.synth_iq (seeded), so only outputs
are stored.  Run:  python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import quisk_oracle as O          # noqa: E402
from oracle import ref_ctypes as R            # noqa: E402

SPLITS = [1, 2, 3, 7, 255, 256, 257, 1000, 1, 2218]
CHAIN_SPLITS = [1, 2, 3, 7, 255, 256, 257, 1000, 4093, 15360, 18766]
DEMOD_SPLITS = [1, 2, 3, 7, 255, 256, 257, 1000, 4093, 6126]
RATES = [1536000, 192000, 96000, 48000, 240000, 250000, 960000, 1200000, 111111, 185185, 50000, 60000, 64000]
DEMOD_TAPS = {"USB": 164, "LSB": 164, "CWU": 390, "CWL": 390, "AM": 77, "FM": 55}
# digital modes: (fixture name, mode, taps, filter_bandwidth)
DGT_CASES = [("DGT-U", "DGT-U", 390, 2800), ("DGT-L", "DGT-L", 390, 2800), ("DGT-U-wide", "DGT-U", 164, 3200),
             ("FDV-L-wide", "FDV-L", 164, 3000), ("DGT-IQ", "DGT-IQ", 77, 2800)]

# (case name, function, seed, real?, table, trailing args, tune)
FILTER_CASES = [
    ("hb45_decim", "quisk_cDecim2HB45", 0, False, None, (), None),
    ("cdecimate_48dec24_2", "quisk_cDecimate", 1, False, "quiskFilt48dec24Coefs", (2,), None),
    ("cdecimate_144D3_3", "quisk_cDecimate", 1, False, "quiskFilt144D3Coefs", (3,), None),
    ("cdecimate_240D5S_5", "quisk_cDecimate", 1, False, "quiskFilt240D5CoefsSharp", (5,), None),
    ("cfilter_53D1", "quisk_cFilter", 1, False, "quiskFilt53D1Coefs", (), None),
    ("ccdecimate_tuned", "quisk_cCDecimate", 2, False, "quiskFilt48dec24Coefs", (2,), (0.1, 1)),
    ("ccdecimate_tuned_lsb", "quisk_cCDecimate", 2, False, "quiskFilt48dec24Coefs", (2,), (0.07, 0)),
    ("ddecimate_lp48_4", "quisk_dDecimate", 3, True, "quiskLpFilt48Coefs", (4,), None),
    ("dfilter_24p6", "quisk_dFilter", 3, True, "quiskAudio24p6Coefs", (), None),
    ("cinterp_24p4_2", "quisk_cInterpolate", 4, False, "quiskAudio24p4Coefs", (2,), None),
    ("cinterp_300D5_6", "quisk_cInterpolate", 4, False, "quiskFilt300D5Coefs", (6,), None),
    ("dinterp_24p4_2", "quisk_dInterpolate", 4, True, "quiskAudio24p4Coefs", (2,), None),
    ("dinterp_24p3_3", "quisk_dInterpolate", 4, True, "quiskAudio24p3Coefs", (3,), None),
    ("interpdecim_300D5_6_5", "quisk_cInterpDecim", 5, False, "quiskFilt300D5Coefs", (6, 5), None),
    ("interpdecim_240D5S_4_5", "quisk_cInterpDecim", 5, False, "quiskFilt240D5CoefsSharp", (4, 5), None),
    ("interpdecim_144D3_2_3", "quisk_cInterpDecim", 5, False, "quiskFilt144D3Coefs", (2, 3), None),
    ("cinterp2hb45", "quisk_cInterp2HB45", 6, False, None, (), None),
    ("dinterp2hb45", "quisk_dInterp2HB45", 6, True, None, (), None),
]


def kat_input(seed, real, n=None):
    x = O.synth_iq(n or sum(SPLITS), seed, 1.0)
    return np.ascontiguousarray(x.real) if real else x


def demod_taps(mode):
    rng = np.random.default_rng(3)
    n = DEMOD_TAPS[mode]
    return rng.standard_normal(n) / n, rng.standard_normal(n) / n


def main():
    flib = R.load("libquisk_filter_ref.so")
    tabs = R.all_tables(flib)
    sys.path.insert(0, "/root/reference")
    import filters as ref_filters             # the reference's filters.py (data only)
    out = dict(tabs)
    for k, v in ref_filters.Filters.items():
        out["proto_%d" % k] = np.array(v, dtype=np.float64)
    np.savez_compressed(os.path.join(ROOT, "quisk_b200", "data", "quisk_tables.npz"), **out)     # package data

    kat = {}
    for name, fn, seed, real, tab, args, tune in FILTER_CASES:
        x = kat_input(seed, real)
        y, counts = R.FilterRunner(flib).run(fn, x, SPLITS, tabs[tab] if tab else None, args, tune)
        kat[name + "/y"] = y
        kat[name + "/counts"] = np.array(counts)
    np.savez_compressed(os.path.join(HERE, "filter_kat.npz"), **kat)

    ch = {}
    for rate in RATES:
        lib = R.load("libquisk_rx_ref.so", private_copy=True)
        lib.ref_set_sample_rate(rate); lib.ref_init_chain()
        x = O.synth_iq(40000, 9, 1.0)
        outs, counts, pos = [], [], 0
        for n in CHAIN_SPLITS:
            buf = np.zeros(66000, dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
            nr = lib.ref_process_decimate(buf.ctypes.data_as(C.c_void_p), n, 0, 3)
            outs.append(buf[:nr].copy()); counts.append(nr)
        ch["decimate_%d/y" % rate] = np.concatenate(outs)
        ch["decimate_%d/counts" % rate] = np.array(counts)
        ch["decimate_%d/srate" % rate] = np.array([lib.ref_decim_srate()])
    for mode in DEMOD_TAPS:
        lib = R.load("libquisk_rx_ref.so", private_copy=True)
        lib.ref_set_sample_rate(48000); lib.ref_init_chain()
        fi, fq = demod_taps(mode)
        fi = np.ascontiguousarray(fi); fq = np.ascontiguousarray(fq)
        lib.ref_set_filters(fi.ctypes.data_as(C.c_void_p), fq.ctypes.data_as(C.c_void_p), len(fi), 2800, 0)
        x = O.synth_iq(12000, 10, 1.0)
        outs, counts, pos = [], [], 0
        for n in DEMOD_SPLITS:
            buf = np.zeros(66000, dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
            dbuf = np.zeros(132000)
            nr = lib.ref_process_demodulate(buf.ctypes.data_as(C.c_void_p), dbuf.ctypes.data_as(C.c_void_p), n, 0, 0, R.MODES[mode])
            outs.append(dbuf[:nr].copy()); counts.append(nr)
        ch["demod_%s/y" % mode] = np.concatenate(outs)
        ch["demod_%s/counts" % mode] = np.array(counts)
        if mode == "FM":
            # the reference's own sensitivity: the same demodulator on the same input moved by one ulp per component
            # (carg(x conj(x_-1)) of a multi-tone input passes close to zero magnitude now and then)
            from tests.golden.make_golden_wdsp import ulp_perturb, rel_rms
            lib = R.load("libquisk_rx_ref.so", private_copy=True)
            lib.ref_set_sample_rate(48000); lib.ref_init_chain()
            lib.ref_set_filters(fi.ctypes.data_as(C.c_void_p), fq.ctypes.data_as(C.c_void_p), len(fi), 2800, 0)
            xp = ulp_perturb(x, 8)
            outs, pos = [], 0
            for n in DEMOD_SPLITS:
                buf = np.zeros(66000, dtype=np.complex128); buf[:n] = xp[pos:pos + n]; pos += n
                dbuf = np.zeros(132000)
                nr = lib.ref_process_demodulate(buf.ctypes.data_as(C.c_void_p), dbuf.ctypes.data_as(C.c_void_p), n, 0, 0, R.MODES[mode])
                outs.append(dbuf[:nr].copy())
            ch["demod_FM/cond"] = np.array([rel_rms(np.concatenate(outs), ch["demod_FM/y"])])
    for name, mode, ntap, bw in DGT_CASES:
        lib = R.load("libquisk_rx_ref.so", private_copy=True)
        lib.ref_set_sample_rate(48000); lib.ref_init_chain()
        rng = np.random.default_rng(3)
        fi = np.ascontiguousarray(rng.standard_normal(ntap) / ntap); fq = np.ascontiguousarray(rng.standard_normal(ntap) / ntap)
        lib.ref_set_filters(fi.ctypes.data_as(C.c_void_p), fq.ctypes.data_as(C.c_void_p), len(fi), bw, 0)
        x = O.synth_iq(12000, 10, 1.0)
        outs, counts, pos = [], [], 0
        for n in DEMOD_SPLITS:
            buf = np.zeros(66000, dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
            dbuf = np.zeros(132000)
            nr = lib.ref_process_demodulate(buf.ctypes.data_as(C.c_void_p), dbuf.ctypes.data_as(C.c_void_p), n, 0, 0, R.MODES[mode])
            outs.append(buf[:nr].copy() if mode == "DGT-IQ" else dbuf[:nr].copy()); counts.append(nr)
        ch["demod_%s/y" % name] = np.concatenate(outs)
        ch["demod_%s/counts" % name] = np.array(counts)
    # Full C1 chain: 1.536 MS/s, USB, bw 2800 -> MakeFilterCoef's 164-tap I/Q pair (quisk.py:5405-5468)
    fi, fq = O.make_filter_coef(12000, None, 2800, 300 + 2800 // 2, ref_filters.Filters)
    ch["c1/filt_i"] = fi; ch["c1/filt_q"] = fq
    lib = R.load("libquisk_rx_ref.so", private_copy=True)
    lib.ref_set_sample_rate(1536000); lib.ref_init_chain()
    fi = np.ascontiguousarray(fi); fq = np.ascontiguousarray(fq)
    lib.ref_set_filters(fi.ctypes.data_as(C.c_void_p), fq.ctypes.data_as(C.c_void_p), len(fi), 2800, 0)
    lib.ref_tune.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
    for tune in (0.0, 12345.0):
        lib2 = R.load("libquisk_rx_ref.so", private_copy=True)
        lib2.ref_set_sample_rate(1536000); lib2.ref_init_chain()
        lib2.ref_set_filters(fi.ctypes.data_as(C.c_void_p), fq.ctypes.data_as(C.c_void_p), len(fi), 2800, 0)
        lib2.ref_tune.argtypes = lib.ref_tune.argtypes
        x = O.synth_iq(153600, 20, 1.0)
        vec = np.array([1.0 + 0j])
        outs, counts, pos = [], [], 0
        for n in [15360] * 10:
            buf = np.zeros(66000, dtype=np.complex128); buf[:n] = x[pos:pos + n]; pos += n
            if tune:
                lib2.ref_tune(buf.ctypes.data, n, tune, 1536000, vec.ctypes.data)
            nd = lib2.ref_process_decimate(buf.ctypes.data_as(C.c_void_p), n, 0, 3)
            dbuf = np.zeros(132000)
            nr = lib2.ref_process_demodulate(buf.ctypes.data_as(C.c_void_p), dbuf.ctypes.data_as(C.c_void_p), nd, 0, 0, 3)
            outs.append(dbuf[:nr].copy()); counts.append(nr)
        ch["c1_tune%d/y" % int(tune)] = np.concatenate(outs)
        ch["c1_tune%d/counts" % int(tune)] = np.array(counts)
    np.savez_compressed(os.path.join(HERE, "chain_kat.npz"), **ch)
    print("wrote", [f for f in os.listdir(HERE) if f.endswith(".npz")])


if __name__ == "__main__":
    main()
