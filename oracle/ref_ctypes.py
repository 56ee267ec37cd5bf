"""oracle/ref_ctypes.py -- TEST INFRASTRUCTURE, not product code.

ctypes bindings for the compiled reference under oracle/_ref/ (built by
oracle/build_ref.sh from /root/reference; the .so files travel to the GPU box,
the sources do not).  The struct mirrors follow filter.h:1-37 and are shared
with the product's own ctypes layer only by layout, not by import.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

c_double_p = C.POINTER(C.c_double)
_private_seq = 0
_private_lock = __import__("threading").Lock()


class cFilter(C.Structure):          # struct quisk_cFilter, filter.h:1-10
    _fields_ = [("dCoefs", c_double_p), ("cpxCoefs", C.c_void_p), ("nBuf", C.c_int),
                ("nTaps", C.c_int), ("decim_index", C.c_int), ("cSamples", C.c_void_p),
                ("ptcSamp", C.c_void_p), ("cBuf", C.c_void_p)]


class dFilter(C.Structure):          # struct quisk_dFilter, filter.h:12-21
    _fields_ = [("dCoefs", c_double_p), ("cpxCoefs", C.c_void_p), ("nBuf", C.c_int),
                ("nTaps", C.c_int), ("decim_index", C.c_int), ("dSamples", C.c_void_p),
                ("ptdSamp", C.c_void_p), ("dBuf", C.c_void_p)]


class cHB45Filter(C.Structure):      # struct quisk_cHB45Filter, filter.h:23-29
    _fields_ = [("cBuf", C.c_void_p), ("nBuf", C.c_int), ("toggle", C.c_int),
                ("samples", C.c_double * 44), ("center", C.c_double * 22)]


class dHB45Filter(C.Structure):      # struct quisk_dHB45Filter, filter.h:31-37
    _fields_ = [("dBuf", C.c_void_p), ("nBuf", C.c_int), ("toggle", C.c_int),
                ("samples", C.c_double * 22), ("center", C.c_double * 11)]


TABLES = {   # filter.h:57-86
    "quiskMicFilt48Coefs": 325, "quiskMic5Filt48Coefs": 424, "quiskMicFilt8Coefs": 93,
    "quiskLpFilt48Coefs": 186, "quiskFilt12_19Coefs": 64, "quiskFilt185D3Coefs": 189,
    "quiskFilt133D2Coefs": 136, "quiskFilt167D3Coefs": 174, "quiskFilt111D2Coefs": 114,
    "quiskFilt53D1Coefs": 55, "quiskFilt53D2Coefs": 93, "quiskFilt144D3Coefs": 147,
    "quiskFilt240D5Coefs": 115, "quiskFilt240D5CoefsSharp": 245, "quiskFilt48dec24Coefs": 98,
    "quiskAudio24p6Coefs": 36, "quiskAudio48p6Coefs": 71, "quiskAudio96Coefs": 11,
    "quiskAudio24p4Coefs": 50, "quiskAudioFmHpCoefs": 309, "quiskAudio24p3Coefs": 100,
    "quiskFiltTx8kAudioB": 168, "quiskFilt16dec8Coefs": 62, "quiskFilt120s03": 480,
    "quiskFiltI3D25Coefs": 825, "quiskDgtFilt48Coefs": 520, "quiskFilt300D5Coefs": 125,
    "quiskFilt300D6Coefs": 248, "quiskFilt240D4Coefs": 100, "quiskDiff48Coefs": 38,
}

MODES = {"CWL": 0, "CWU": 1, "LSB": 2, "USB": 3, "AM": 4, "FM": 5,
         "DGT-U": 7, "DGT-L": 8, "DGT-IQ": 9, "FDV-U": 11, "FDV-L": 12, "DGT-FM": 13}   # quisk.h:56-70


def have_ref(name: str = "libquisk_filter_ref.so") -> bool:
    return os.path.exists(os.path.join(REF_DIR, name))


def load(name: str, private_copy: bool = False) -> C.CDLL:
    """dlopen a reference library.  private_copy=True loads a fresh temporary
    copy so that the reference's function-local `static` state starts from zero."""
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path}: run oracle/build_ref.sh (needs /root/reference)")
    if private_copy:
        # A distinct path gives a distinct dlopen handle with its own statics.  The copies stay on disk under
        # oracle/_ref/private/ (git-ignored with the rest of oracle/_ref) instead of being unlinked, so that whoever
        # audits which native code a process loaded (/proc/<pid>/maps) sees the reference library by name.
        global _private_seq
        pdir = os.path.join(REF_DIR, "private")
        os.makedirs(pdir, exist_ok=True)
        with _private_lock:
            k = _private_seq
            _private_seq += 1
        dst = os.path.join(pdir, "%s.%d.so" % (name[:-3], k))
        st = os.stat(path)
        if not (os.path.exists(dst) and os.path.getsize(dst) == st.st_size and os.path.getmtime(dst) >= st.st_mtime):
            tmp = "%s.%d.tmp" % (dst, os.getpid())
            shutil.copy(path, tmp)
            os.replace(tmp, dst)
        return C.CDLL(dst)
    return C.CDLL(path)


def bind_filter_api(lib: C.CDLL) -> C.CDLL:
    """Declare the filter.h prototypes (filter.h:39-55) on `lib`."""
    vp = C.c_void_p
    lib.quisk_filt_cInit.argtypes = [C.POINTER(cFilter), c_double_p, C.c_int]
    lib.quisk_filt_cInit.restype = None
    lib.quisk_filt_dInit.argtypes = [C.POINTER(dFilter), c_double_p, C.c_int]
    lib.quisk_filt_dInit.restype = None
    lib.quisk_filt_tune.argtypes = [vp, C.c_double, C.c_int]
    lib.quisk_filt_tune.restype = None
    for nm in ("quisk_cInterpolate", "quisk_cDecimate", "quisk_cCDecimate"):
        getattr(lib, nm).argtypes = [vp, C.c_int, C.POINTER(cFilter), C.c_int]
        getattr(lib, nm).restype = C.c_int
    for nm in ("quisk_dInterpolate", "quisk_dDecimate"):
        getattr(lib, nm).argtypes = [vp, C.c_int, C.POINTER(dFilter), C.c_int]
        getattr(lib, nm).restype = C.c_int
    lib.quisk_cInterpDecim.argtypes = [vp, C.c_int, C.POINTER(cFilter), C.c_int, C.c_int]
    lib.quisk_cInterpDecim.restype = C.c_int
    lib.quisk_cDecim2HB45.argtypes = [vp, C.c_int, C.POINTER(cHB45Filter)]
    lib.quisk_cDecim2HB45.restype = C.c_int
    lib.quisk_cInterp2HB45.argtypes = [vp, C.c_int, C.POINTER(cHB45Filter)]
    lib.quisk_cInterp2HB45.restype = C.c_int
    lib.quisk_dInterp2HB45.argtypes = [vp, C.c_int, C.POINTER(dHB45Filter)]
    lib.quisk_dInterp2HB45.restype = C.c_int
    lib.quisk_dFilter.argtypes = [vp, C.c_int, C.POINTER(dFilter)]
    lib.quisk_dFilter.restype = C.c_int
    lib.quisk_cFilter.argtypes = [vp, C.c_int, C.POINTER(cFilter)]
    lib.quisk_cFilter.restype = C.c_int
    lib.quisk_dD_out.argtypes = [C.c_double, C.POINTER(dFilter)]
    lib.quisk_dD_out.restype = C.c_double
    return lib


def table(lib: C.CDLL, name: str) -> np.ndarray:
    n = TABLES[name]
    arr = (C.c_double * n).in_dll(lib, name)
    return np.ctypeslib.as_array(arr).copy()


def all_tables(lib: C.CDLL | None = None) -> dict:
    lib = lib or load("libquisk_filter_ref.so")
    return {k: table(lib, k) for k in TABLES}


class FilterRunner:
    """Drives one filter.h block function over a list of block lengths, in place,
    exactly like the reference's callers do.  Works for the reference library
    and for libquisk_cuda (same ABI)."""

    def __init__(self, lib: C.CDLL):
        self.lib = bind_filter_api(lib)
        self._keep = []

    def _coefs(self, coefs):
        a = np.ascontiguousarray(coefs, dtype=np.float64)
        self._keep.append(a)
        return a.ctypes.data_as(c_double_p), len(a)

    def run(self, fn: str, x: np.ndarray, splits, coefs=None, args=(), tune=None):
        """fn: function name; x: complex128 or float64 input; splits: block
        lengths; coefs: tap table for the cFilter/dFilter functions; args: the
        trailing ints (factor / interp, decim); tune: (freq, ssb_upper) to call
        quisk_filt_tune first.  Returns (output, [nOut per block])."""
        lib = self.lib
        is_c = x.dtype == np.complex128
        if fn in ("quisk_cDecim2HB45", "quisk_cInterp2HB45"):
            st = cHB45Filter()
        elif fn == "quisk_dInterp2HB45":
            st = dHB45Filter()
        elif is_c:
            st = cFilter()
            p, n = self._coefs(coefs)
            lib.quisk_filt_cInit(C.byref(st), p, n)
        else:
            st = dFilter()
            p, n = self._coefs(coefs)
            lib.quisk_filt_dInit(C.byref(st), p, n)
        if tune is not None:
            lib.quisk_filt_tune(C.cast(C.byref(st), C.c_void_p), C.c_double(tune[0]), int(tune[1]))
        f = getattr(lib, fn)
        outs, counts = [], []
        pos = 0
        for n_blk in splits:
            blk = x[pos:pos + n_blk]
            pos += n_blk
            buf = np.zeros(SAMP_CAP, dtype=x.dtype)
            buf[:len(blk)] = blk
            n_out = f(buf.ctypes.data_as(C.c_void_p), len(blk), C.byref(st), *args)
            counts.append(n_out)
            outs.append(buf[:n_out].copy())
        self.state = st
        return (np.concatenate(outs) if outs else x[:0]), counts


SAMP_CAP = 66000 * 4   # generous scratch: interpolators write up to 52 802 samples
