"""oracle/quisk_oracle.py -- TEST INFRASTRUCTURE (CPU oracle), NOT product code.

A NumPy restatement of the reference's receive-DSP hot path (Quisk 4.2.52,
`filter.c`, the RX functions of `quisk.c`, the panadapter math).  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import
this module; the product (`quisk_b200`, `libquisk_cuda.so`) never does.

Formulation.  The reference keeps ring buffers and computes one output at a
time.  This restatement keeps, for every filter, a *linear history* of the most
recent inputs and evaluates each block as an explicit convolution over
`[history | block]`.  That is a deliberately different program from both the
reference and the CUDA kernels, so agreement between the three is meaningful.
Floating-point sums are therefore ordered differently from the reference
(agreement ~1e-15 relative); sample COUNTS and phase carry-over are exact.

Pinned against the compiled reference (`oracle/_ref`, built by
`oracle/build_ref.sh` from /root/reference) by `tests/test_oracle_vs_ref.py` and
against the committed fixtures in `tests/golden/` (generated from the compiled
reference by `tests/golden/make_golden.py`).  The reference itself ships no
tests or golden vectors, so that is the only pin available.

All citations are file:line under /root/reference.
"""
from __future__ import annotations

import math
import numpy as np

SAMP_BUFFER_SIZE = 66000                  # quisk.h:15
OUT_CLIP = SAMP_BUFFER_SIZE * 8 // 10     # filter.c:158,194,315 (52 800)
CLIP32 = 2147483647                       # quisk.h:13

# Half-band 45-tap coefficients: the static table inside quisk_cDecim2HB45
# (filter.c:381-384).  coef[11] = 0.5 is the centre tap; taps 0 and 44 of the
# 45-tap prototype are zero.
HB45_COEF = np.array([
    0.000018566625444266, -0.000118469698701817, 0.000457318798253456,
    -0.001347840471412094, 0.003321838571445455, -0.007198422696929033,
    0.014211106939802483, -0.026424776824073383, 0.048414810444971007,
    -0.096214669073304823, 0.314881034738348550, 0.500000000000000000])


def hb45_dense() -> np.ndarray:
    """The half-band as a dense 43-tap filter g (zero outer taps dropped):
    g[2k] = g[42-2k] = coef[k], g[21] = 0.5  (filter.c:401-413)."""
    g = np.zeros(43)
    for k in range(11):
        g[2 * k] = HB45_COEF[k]
        g[42 - 2 * k] = HB45_COEF[k]
    g[21] = HB45_COEF[11]
    return g


def _conv_at(X: np.ndarray, h: np.ndarray, pos: np.ndarray) -> np.ndarray:
    """sum_k h[k] * X[pos-k]; pos-k is guaranteed in range by the callers."""
    if len(pos) == 0:
        return np.zeros(0, dtype=np.result_type(X, h))
    full = np.convolve(X, h)
    return full[pos]


class HB45Decim:
    """quisk_cDecim2HB45 (filter.c:377-417).  With n the index of an input
    since the filter was fresh, an output is produced on every odd n and equals
    sum_{k<=10} (x[n-2k] + x[n-42+2k]) c[k] + 0.5 x[n-21]."""

    H = 42

    def __init__(self, dtype=np.complex128):
        self.hist = np.zeros(self.H, dtype=dtype)
        self.toggle = 0
        self.g = hb45_dense()

    def __call__(self, x: np.ndarray) -> np.ndarray:
        X = np.concatenate([self.hist, x])
        # the input at block offset i is an output sample iff (toggle + i) is odd
        first = 1 - self.toggle
        i = np.arange(first, len(x), 2)
        y = _conv_at(X, self.g, i + self.H)
        self.toggle = (self.toggle + len(x)) & 1
        self.hist = X[len(X) - self.H:]
        return y


class FirDecim:
    """quisk_cDecimate / quisk_dDecimate / quisk_cFilter / quisk_dFilter
    (filter.c:203-229, 259-285, 347-375) and, with complex `coefs`,
    quisk_cCDecimate (filter.c:231-257).  Output on the input that makes
    ++decim_index reach `decim`; newest sample times coef[0]."""

    def __init__(self, coefs, decim=1, dtype=np.complex128):
        self.h = np.asarray(coefs)
        self.decim = int(decim)
        self.H = len(self.h) - 1
        self.hist = np.zeros(self.H, dtype=dtype)
        self.decim_index = 0

    def __call__(self, x: np.ndarray) -> np.ndarray:
        X = np.concatenate([self.hist, x])
        D = self.decim
        first = D - 1 - self.decim_index
        i = np.arange(first, len(x), D)
        y = _conv_at(X, self.h, i + self.H)
        self.decim_index = (self.decim_index + len(x)) % D
        self.hist = X[len(X) - self.H:] if self.H else X[:0]
        return y


class FirInterp:
    """quisk_cInterpolate / quisk_dInterpolate (filter.c:131-201): per input,
    `interp` outputs interp * sum_{k < nTaps/interp} x[i-k] coef[j + k*interp].
    NOTE the integer quotient nTaps/interp: trailing taps are never used
    (SURVEY.md F10).  Output is cut at 52 800 samples (filter.c:158)."""

    def __init__(self, coefs, interp, dtype=np.complex128):
        self.h = np.asarray(coefs, dtype=np.float64)
        self.L = int(interp)
        self.K = len(self.h) // self.L
        self.H = len(self.h) - 1          # the reference ring holds nTaps samples
        self.hist = np.zeros(self.H, dtype=dtype)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        X = np.concatenate([self.hist, x])
        n = len(x)
        y = np.zeros(n * self.L, dtype=X.dtype)
        pos = np.arange(n) + self.H
        for j in range(self.L):
            hj = self.h[j:j + self.K * self.L:self.L][:self.K]
            y[j::self.L] = _conv_at(X, hj, pos) * self.L
        self.hist = X[len(X) - self.H:]
        return y[:OUT_CLIP]


class FirInterpDecim:
    """quisk_cInterpDecim (filter.c:287-324).  On the up-sampled time line input
    i covers positions [i*L, (i+1)*L); output m sits at u = decim_index0 + m*M,
    i.e. comes from input floor(u/L) with phase u mod L.  Same nTaps/interp tap
    truncation as FirInterp; same 52 800 output cut (filter.c:315)."""

    def __init__(self, coefs, interp, decim, dtype=np.complex128):
        self.h = np.asarray(coefs, dtype=np.float64)
        self.L = int(interp)
        self.M = int(decim)
        self.K = len(self.h) // self.L
        self.H = len(self.h) - 1
        self.hist = np.zeros(self.H, dtype=dtype)
        self.decim_index = 0

    def n_out(self, count: int) -> int:
        span = count * self.L - self.decim_index
        return max(0, -(-span // self.M))

    def __call__(self, x: np.ndarray) -> np.ndarray:
        X = np.concatenate([self.hist, x])
        n = len(x)
        nout = self.n_out(n)
        u = self.decim_index + np.arange(nout) * self.M
        src = u // self.L
        ph = u % self.L
        y = np.zeros(nout, dtype=X.dtype)
        for p in range(self.L):
            sel = np.nonzero(ph == p)[0]
            if len(sel) == 0:
                continue
            hp = self.h[p:p + self.K * self.L:self.L][:self.K]
            y[sel] = _conv_at(X, hp, src[sel] + self.H) * self.L
        self.decim_index = self.decim_index + nout * self.M - n * self.L
        self.hist = X[len(X) - self.H:]
        return y[:OUT_CLIP]


class HB45Interp:
    """quisk_dInterp2HB45 / quisk_cInterp2HB45 (filter.c:420-488): per input two
    outputs, 2*c[11]*s[11] and 2*sum_{k<11}(s[k]+s[21-k])*c[k], s[k] = x[i-k].
    The output clip test is `nOut > 52800` checked before each PAIR
    (filter.c:444,479), so up to 52 802 samples can be written."""

    H = 21

    def __init__(self, dtype=np.float64):
        self.hist = np.zeros(self.H, dtype=dtype)
        side = np.zeros(22)
        for k in range(11):
            side[k] = HB45_COEF[k]
            side[21 - k] = HB45_COEF[k]
        self.side = side

    def __call__(self, x: np.ndarray) -> np.ndarray:
        X = np.concatenate([self.hist, x])
        n = len(x)
        pos = np.arange(n) + self.H
        y = np.zeros(2 * n, dtype=X.dtype)
        y[0::2] = X[pos - 11] * HB45_COEF[11] * 2
        y[1::2] = _conv_at(X, self.side, pos) * 2
        self.hist = X[len(X) - self.H:]
        npairs = min(n, (OUT_CLIP + 2) // 2)     # pairs written while nOut <= 52800
        return y[:2 * npairs]


def rx_filter_impulse(filt: np.ndarray) -> np.ndarray:
    """Effective impulse response of cRxFilterOut/dRxFilterOut for a tap table
    `filt` (quisk.c:1203-1215, 1240-1255): the newest sample meets filt[0], then
    the ring is walked FORWARD from the write index, so the oldest sample meets
    filt[1] ... the previous sample meets filt[N-1]:  h[0] = filt[0],
    h[m] = filt[N-m]  (SURVEY.md F2)."""
    f = np.asarray(filt, dtype=np.float64)
    h = np.empty_like(f)
    h[0] = f[0]
    h[1:] = f[:0:-1]
    return h


class RxFilterC:
    """cRxFilterOut (quisk.c:1218-1256): I rail filtered by filtI, Q rail by
    filtQ, result accI + j accQ."""

    def __init__(self, filt_i, filt_q):
        self.hi = rx_filter_impulse(filt_i)
        self.hq = rx_filter_impulse(filt_q)
        self.H = len(self.hi) - 1
        self.hist = np.zeros(self.H, dtype=np.complex128)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        X = np.concatenate([self.hist, x])
        pos = np.arange(len(x)) + self.H
        yi = _conv_at(X.real, self.hi, pos)
        yq = _conv_at(X.imag, self.hq, pos)
        self.hist = X[len(X) - self.H:]
        return yi + 1j * yq


class RxFilterD:
    """dRxFilterOut (quisk.c:1182-1216): one real tap set on complex samples."""

    def __init__(self, filt_i):
        self.h = rx_filter_impulse(filt_i)
        self.H = len(self.h) - 1
        self.hist = np.zeros(self.H, dtype=np.complex128)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        X = np.concatenate([self.hist, x])
        pos = np.arange(len(x)) + self.H
        y = _conv_at(X, self.h, pos)
        self.hist = X[len(X) - self.H:]
        return y


class TuneNCO:
    """The tuning loop of quisk_process_samples (quisk.c:2477-2488):
    x[i] *= v; v *= phase, phase = cexp(-j 2 pi tune / rate); v is static and
    starts at 1.  Restated literally (sequential recurrence, separately rounded
    products as gcc -O2 without FMA produces), so use it for short vectors."""

    def __init__(self, tune_hz: float, rate: int):
        ang = -2.0 * math.pi * tune_hz / rate
        self.phase = complex(math.cos(ang), math.sin(ang))
        self.v = 1.0 + 0.0j
        self.tune_hz = tune_hz

    def __call__(self, x: np.ndarray) -> np.ndarray:
        if self.tune_hz == 0:
            return x.copy()
        y = np.empty_like(x)
        vr, vi = self.v.real, self.v.imag
        pr, pi = self.phase.real, self.phase.imag
        xr = x.real.tolist()
        xi = x.imag.tolist()
        outr = [0.0] * len(xr)
        outi = [0.0] * len(xr)
        for n in range(len(xr)):
            a, b = xr[n], xi[n]
            outr[n] = a * vr - b * vi
            outi[n] = a * vi + b * vr
            vr, vi = vr * pr - vi * pi, vr * pi + vi * pr
        y.real = outr
        y.imag = outi
        self.v = complex(vr, vi)
        return y


def plan_decimation(sample_rate: int):
    """PlanDecimation (quisk.c:1633-1671): the (i2<=6, i3<=3, i5<=3) with the
    smallest resulting rate >= 48000 (first found wins ties because the test is
    a strict `<`).  Returns (best_rate_after_24/25_fixup, decim2, decim3, decim5)."""
    best = sample_rate
    d2 = d3 = d5 = 0
    for i2 in range(7):
        for i3 in range(4):
            for i5 in range(4):
                t = sample_rate
                for _ in range(i2):
                    t //= 2
                for _ in range(i3):
                    t //= 3
                for _ in range(i5):
                    t //= 5
                if 48000 <= t < best:
                    d2, d3, d5, best = i2, i3, i5, t
    if best >= 50000:
        best = best * 24 // 25
    return best, d2, d3, d5


class ProcessDecimate:
    """quisk_process_decimate, default branch (quisk.c:1769-1844): decim2-1 half
    bands, decim3 x (147 taps /3), decim5 x (245 taps /5), then the 98-tap /2 FIR
    if decim2 > 0, then the 6/5 + 4/5 rate fix when the result is >= 50 kS/s.
    `tables` maps the reference's coefficient-table names to arrays.  The SDR-IQ
    special rates (quisk.c:1731-1767) are handled as in the reference."""

    def __init__(self, sample_rate: int, tables: dict):
        self.rate = sample_rate
        self.stages = []
        key = (sample_rate + 100) // 1000
        T = tables
        rate = sample_rate
        if key == 41:
            rate = 48000
        elif key == 53:
            self.stages.append(FirDecim(T["quiskFilt53D1Coefs"], 1))
        elif key == 111:
            self.stages.append(FirDecim(T["quiskFilt111D2Coefs"], 2)); rate //= 2
        elif key == 133:
            self.stages.append(FirDecim(T["quiskFilt133D2Coefs"], 2)); rate //= 2
        elif key == 185:
            self.stages.append(FirDecim(T["quiskFilt185D3Coefs"], 3)); rate //= 3
        elif key == 370:
            self.stages += [HB45Decim(), FirDecim(T["quiskFilt185D3Coefs"], 3)]; rate //= 6
        elif key == 740:
            self.stages += [HB45Decim(), HB45Decim(), FirDecim(T["quiskFilt185D3Coefs"], 3)]; rate //= 12
        elif key == 1333:
            self.stages += [HB45Decim(), HB45Decim(), HB45Decim(), FirDecim(T["quiskFilt167D3Coefs"], 3)]; rate //= 24
        else:
            _, d2, d3, d5 = plan_decimation(sample_rate)
            i2 = d2
            n_hb = 0
            while i2 > 1 and n_hb < 5:
                self.stages.append(HB45Decim()); rate //= 2; i2 -= 1; n_hb += 1
            for _ in range(d3):
                self.stages.append(FirDecim(T["quiskFilt144D3Coefs"], 3)); rate //= 3
            for _ in range(d5):
                self.stages.append(FirDecim(T["quiskFilt240D5CoefsSharp"], 5)); rate //= 5
            if i2 > 0:
                self.stages.append(FirDecim(T["quiskFilt48dec24Coefs"], 2)); rate //= 2
            if rate >= 50000:
                rate = rate * 24 // 25
                self.stages.append(FirInterpDecim(T["quiskFilt300D5Coefs"], 6, 5))
                self.stages.append(FirInterpDecim(T["quiskFilt240D5CoefsSharp"], 4, 5))
        self.decim_srate = rate

    def __call__(self, x: np.ndarray) -> np.ndarray:
        for s in self.stages:
            x = s(x)
        return x


FM_FILTER_DEMPH = 300.0      # quisk.c:46


class ProcessDemodulate:
    """quisk_process_demodulate (quisk.c:1848-2160) for the modes on the hot
    path: CWL/CWU (/8 -> 6 k), LSB/USB (/4 -> 12 k), AM (/2 -> 24 k, |x| + DC
    block, 36-tap audio FIR), FM (48 k, arg(x conj(x_-1)), x 20e5, one-pole
    de-emphasis, /4, high-pass FIR, x4), DGT-U/L and FDV-U/L (CW's structure below
    3000 Hz bandwidth, else the I/Q filter at 48 k), DGT-IQ (real-tap filter, complex out).  The optional auto-notch / SSB squelch
    are off, as they are by default in the reference.  Returns real audio at 48 k."""

    def __init__(self, mode: str, filt_i, filt_q, tables: dict, bandwidth: int = 2800):
        T = tables
        self.mode = mode
        self.pre = []
        self.post = []
        if mode in ("DGT-U", "DGT-L", "FDV-U", "FDV-L"):        # quisk.c:2087-2140
            if bandwidth < 3000:                                # DGT_NARROW_FREQ, quisk.c:52
                self.pre = [HB45Decim(), HB45Decim(), FirDecim(T["quiskFilt48dec24Coefs"], 2)]
                self.post = [FirInterp(T["quiskAudio24p4Coefs"], 2, np.float64), HB45Interp(), HB45Interp()]
            self.rx = RxFilterC(filt_i, filt_q)
        elif mode == "DGT-IQ":                                  # quisk.c:2141-2153
            self.rx = RxFilterD(filt_i) if bandwidth < 19000 else (lambda x: x)
        elif mode in ("CWL", "CWU"):
            self.pre = [HB45Decim(), HB45Decim(), FirDecim(T["quiskFilt48dec24Coefs"], 2)]
            self.rx = RxFilterC(filt_i, filt_q)
            self.post = [FirInterp(T["quiskAudio24p4Coefs"], 2, np.float64), HB45Interp(), HB45Interp()]
        elif mode in ("LSB", "USB"):
            self.pre = [HB45Decim(), FirDecim(T["quiskFilt48dec24Coefs"], 2)]
            self.rx = RxFilterC(filt_i, filt_q)
            self.post = [FirInterp(T["quiskAudio24p4Coefs"], 2, np.float64), HB45Interp()]
        elif mode == "AM":
            self.pre = [FirDecim(T["quiskFilt48dec24Coefs"], 2)]
            self.rx = RxFilterD(filt_i)
            self.dc_remove = 0.0
            self.post = [FirDecim(T["quiskAudio24p6Coefs"], 1, np.float64), HB45Interp()]
        elif mode == "FM":
            self.rx = RxFilterD(filt_i)
            self.fm_1 = 10.0 + 0.0j                             # quisk.c:1893
            www = math.tan(math.pi * FM_FILTER_DEMPH / 48000)   # quisk.c:1895-1899
            nnn = 1.0 / (1.0 + www)
            self.a0 = www * nnn
            self.a1 = self.a0
            self.b1 = nnn * (www - 1.0)
            self.x1 = 0.0
            self.y1 = 0.0
            self.post = [FirDecim(T["quiskLpFilt48Coefs"], 4, np.float64),
                         FirDecim(T["quiskAudioFmHpCoefs"], 1, np.float64),
                         HB45Interp(), HB45Interp()]
        else:
            raise ValueError(mode)

    def __call__(self, x: np.ndarray) -> np.ndarray:
        for s in self.pre:
            x = s(x)
        cx = self.rx(x)
        if self.mode == "DGT-IQ":
            return cx                                   # complex out
        if self.mode in ("CWL", "LSB", "DGT-L", "FDV-L"):
            d = cx.real + cx.imag                       # quisk.c:1916,1962,2127
        elif self.mode in ("CWU", "USB", "DGT-U", "FDV-U"):
            d = cx.real - cx.imag                       # quisk.c:1939,1986,2100
        elif self.mode == "AM":
            mag = np.abs(cx)                            # quisk.c:2007-2011
            d = np.empty(len(mag))
            dc = self.dc_remove
            for i, m in enumerate(mag.tolist()):
                t = m + dc * 0.99
                d[i] = t - dc
                dc = t
            self.dc_remove = dc
        else:                                           # FM, quisk.c:2029-2064
            prev = np.concatenate([[self.fm_1], cx[:-1]]) if len(cx) else cx
            di = np.angle(cx * np.conj(prev))
            if len(cx):
                self.fm_1 = cx[-1]
            di = di * 20e5
            d = np.empty(len(di))
            x1, y1 = self.x1, self.y1
            for i, v in enumerate(di.tolist()):
                y1 = v * self.a0 + x1 * self.a1 - y1 * self.b1
                x1 = v
                d[i] = y1
            self.x1, self.y1 = x1, y1
        for s in self.post:
            d = s(d)
        return d


def make_filter_coef(rate: int, N, bw: int, center: int, prototypes: dict):
    """MakeFilterCoef (quisk.py:5405-5456).  `prototypes` is the Filters dict of
    filters.py (key = bw*24000//rate//2).  Returns (filtI, filtQ)."""
    center = abs(center)
    lowpass = bw * 24000 // rate // 2
    if lowpass in prototypes:
        filt_d = list(prototypes[lowpass])
    else:
        if N is None:
            shape = 1.5
            trans = (bw / 2.0 / rate) * (shape - 1.0)
            N = int(4.0 / trans)
            if N > 1000:
                N = 1000
            N = (N // 2) * 2 + 1
        K = bw * N // rate
        filt_d = []
        for k in range(-N // 2, N // 2 + 1):
            if k == 0:
                z = float(K) / N
            else:
                z = 1.0 / N * math.sin(math.pi * k * K / N) / math.sin(math.pi * k / N)
            w = 0.42 + 0.5 * math.cos(2. * math.pi * k / N) + 0.08 * math.cos(4. * math.pi * k / N)
            filt_d.append(z * w)
    if center:
        import cmath
        tune = -1j * 2.0 * math.pi * center / rate
        NN = len(filt_d)
        D = (NN - 1.0) / 2.0
        fi, fq = [], []
        for i in range(NN):
            z = 2.0 * cmath.exp(tune * (i - D)) * filt_d[i]
            fi.append(z.real)
            fq.append(z.imag)
        return np.array(fi), np.array(fq)
    return np.array(filt_d), np.array(filt_d)


# --------------------------------------------------------------------------
# Panadapter (quisk.c:5142-5331, 4868-4930, 4957-5011, 4932-4955, 6003-6009)
# --------------------------------------------------------------------------

def hann_window(n: int) -> np.ndarray:
    """record_app window (quisk.c:6003-6009): 0.5 + 0.5 cos(2 pi j / N), j = -N/2 .. N/2-1."""
    j = np.arange(n) - n // 2
    return 0.5 + 0.5 * np.cos(2.0 * math.pi * j / n)


def panadapter_accumulate(frames: np.ndarray) -> np.ndarray:
    """get_graph's per-frame work (quisk.c:5212-5215, 5271-5276): window, complex
    FFT, then fft_avg[k] += |X[(k + N/2) mod N]| over the given frames.
    `frames` is [count_fft, N] complex."""
    n = frames.shape[-1]
    w = hann_window(n)
    X = np.fft.fft(frames * w, axis=-1)
    # graph bin k <- FFT bin (k + N/2) mod N with N/2 the C integer quotient (quisk.c:5272-5275): fftshift for even N
    return np.abs(np.roll(X, -(n // 2), axis=-1)).sum(axis=0)


def panadapter_pixels(fft_avg: np.ndarray, count_fft: int, data_width: int, zoom: float,
                      deltaf: float, fft_sample_rate: float) -> np.ndarray:
    """The graph-return half of get_graph (quisk.c:5279-5321): bin fft_avg into
    data_width pixels IN PLACE (pixel i overwrites fft_avg[i] while later pixels
    still read fft_avg -- the aliasing is reproduced), then
    20 log10(sum) - 20 (log10 count + log10 N + 31 log10 2), clamped to [-200, 0]."""
    avg = np.array(fft_avg, dtype=np.float64)
    n_fft = len(avg)
    scale = (math.log10(count_fft) + math.log10(n_fft) + 31.0 * math.log10(2.0)) * 20.0
    n = int(zoom * float(n_fft) / data_width + 0.5)
    if n < 1:
        n = 1
    for i in range(data_width):
        k = int(n_fft * (deltaf / fft_sample_rate + zoom * (float(i) / data_width - 0.5) + 0.5) + 0.1)
        d2 = 0.0
        for j in range(n):
            if 0 <= k < n_fft:
                d2 += avg[k]
            k += 1
        avg[i] = d2
    out = np.empty(data_width)
    for i in range(data_width):
        d2 = 20.0 * math.log10(avg[i]) - scale if avg[i] > 0 else -math.inf
        if d2 < -200:
            d2 = -200.0
        elif d2 > 0:
            d2 = 0.0
        out[i] = d2
    return out


def multirx_graph(samples: np.ndarray, mult: int = 8) -> np.ndarray:
    """get_multirx_graph (quisk.c:4868-4930): Hann, FFT, |X| summed in groups of
    MULTIRX_FFT_MULT bins in fftshift order, 20 log10 - 20 (log10 N + 31 log10 2),
    floor -200 (no upper clamp)."""
    n = len(samples)
    X = np.fft.fft(samples * hann_window(n))
    mag = np.abs(np.fft.fftshift(X))
    scale = (math.log10(n) + 31.0 * math.log10(2.0)) * 20.0
    d1 = mag.reshape(-1, mult).sum(axis=1)
    with np.errstate(divide="ignore"):
        d2 = 20.0 * np.log10(d1) - scale
    return np.maximum(d2, -200.0)


def copy2pixels(fft: np.ndarray, n_pixels: int, zoom: float, deltaf: float, rate: float) -> np.ndarray:
    """copy2pixels (quisk.c:4932-4955): fractional-bin integration of `fft`
    (which must have one spare element at the end) into n_pixels."""
    fft_size = len(fft) - 1
    f1 = deltaf + rate / 2.0 * (1.0 - zoom)
    out = np.empty(n_pixels)
    for i in range(n_pixels):
        d1 = fft_size / rate * (f1 + float(i) / n_pixels * zoom * rate)
        d2 = fft_size / rate * (f1 + float(i + 1) / n_pixels * zoom * rate)
        j1 = math.floor(d1)
        j2 = math.floor(d2)
        if j1 == j2:
            s = (d2 - d1) * fft[j1]
        else:
            s = (j1 + 1 - d1) * fft[j1]
            for j in range(j1 + 1, j2):
                s += fft[j]
            s += (d2 - j2) * fft[j2]
        out[i] = s
    return out


def bandscope(blocks: np.ndarray, graph_width: int, clock: int, zoom: float, deltaf: float) -> np.ndarray:
    """get_bandscope (quisk.c:4957-5011) for fft_count = len(blocks) real blocks
    of bandscope_size samples: window (Hann, quisk.c bandscopeWindow), r2c FFT,
    |X| average over L = size/2+1 bins, copy2pixels, scale, 20 log10 (<=1e-10 -> -200)."""
    fft_count, size = blocks.shape
    L = size // 2 + 1
    w = hann_window(size)
    avg = np.abs(np.fft.rfft(blocks * w, axis=-1)).sum(axis=0)
    avg = np.concatenate([avg, [0.0]])
    frac = float(L) / graph_width
    scale = 1.0 / frac / fft_count / size
    pix = copy2pixels(avg, graph_width, zoom, deltaf, clock / 2.0)
    out = np.empty(graph_width)
    for i in range(graph_width):
        s = pix[i] * scale
        out[i] = -200.0 if s <= 1e-10 else 20.0 * math.log10(s)
    return out


# --------------------------------------------------------------------------
# NoiseBlanker (quisk.c:679-784): optional stage on the raw samples in front of the tuning stage
# --------------------------------------------------------------------------

class NoiseBlanker:
    """The pulse decision `|x| > mean(|x| over the last 3 windows) * limit` depends only on the magnitudes, so it is
    restated here as a pre-pass (magnitudes vectorised, the running sum walked in the reference's order because each
    `-=` / `+=` is separately rounded, quisk.c:732-735); the blanking state machine (quisk.c:740-765) then edits a
    delay line of `3 * hwindow` samples: ramp down the `hwindow` samples in front of a pulse, zero while pulses last,
    ramp up over the next `hwindow`.  Output = the delay line's oldest sample."""

    def __init__(self, sample_rate: int, level: int):
        self.hw = int(sample_rate * 500.0e-6 + 0.5)             # QUISK_NB_HWINDOW_SECS, quisk.c:679,704
        self.size = 3 * self.hw
        self.limit = {2: 4.0, 3: 2.5}.get(level, 6.0)           # quisk.c:716-728
        self.line = np.zeros(self.size, dtype=np.complex128)
        self.mags = np.zeros(self.size)
        self.total = 0.0
        self.pos = 0
        self.blanking = False
        self.ramp = 0

    def __call__(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, dtype=np.complex128)
        mag = np.hypot(x.real, x.imag)
        pulse = np.zeros(len(x), dtype=bool)
        p = self.pos
        for i, m in enumerate(mag):
            self.total -= self.mags[p]
            self.mags[p] = m
            self.total += m
            pulse[i] = not (m <= self.total / self.size * self.limit)
            p = (p + 1) % self.size
        y = np.empty_like(x)
        for i in range(len(x)):
            p = self.pos
            y[i] = self.line[p]
            self.line[p] = x[i]
            if self.blanking:
                self.line[p] = 0.0
                if not pulse[i]:
                    self.blanking, self.ramp = False, 1
            elif pulse[i]:
                self.blanking = True
                back = (p - np.arange(self.hw)) % self.size
                self.line[back] *= np.arange(self.hw) / float(self.hw)
            elif self.ramp:
                self.line[p] *= float(self.ramp) / self.hw
                self.ramp = self.ramp + 1 if self.ramp + 1 < self.hw else 0
            self.pos = (p + 1) % self.size
        return y


# --------------------------------------------------------------------------
# dAutoNotch (quisk.c:786-963): optional automatic notch on the SSB audio at the filter rate
# --------------------------------------------------------------------------

class AutoNotch:
    """Overlap-save with frames of 2048 samples advancing by 1538 and a 511-tap notch filter that follows the one or two
    strongest steady lines of a per-bin running average of |X|.  Restated frame by frame: the reference's sample loop
    emits, for input sample m, the filtered sample m - 1538 (zeros for the first frame).  State after the reference's
    initialising call (quisk.c:826-835)."""
    N, START, BINS = 2048, 510, 1025

    def __init__(self, rate: int, sidetone: int = 0):
        B = self.BINS
        self.delta_sig = (300 * 2 * B + rate // 2) // rate
        self.delta_i1 = (400 * 2 * B + rate // 2) // rate
        self.signal = (abs(sidetone) * 2 * B + rate // 2) // rate if sidetone else -999
        self.half_width = max((100 * 2 * 256 + rate // 2) // rate, 3)
        self.window = 0.50 - 0.50 * np.cos(2.0 * np.pi * np.arange(511) / 511)
        self.frame = np.zeros(self.N)
        self.fill = self.START
        self.ready = np.zeros(self.N - self.START)      # filtered samples waiting to be emitted
        self.avg = np.zeros(B)
        self.H = np.zeros(B, dtype=np.complex128)
        self.old1 = self.old2 = 0
        self.count1 = self.count2 = -4
        self.sig = -1

    def _peak(self, mask):
        cand = np.where(mask, self.avg, 0.0)
        i = int(np.argmax(cand))                         # first maximum, like the reference's strict `>` scan
        return i if cand[i] > 0 else 0

    def _design(self, i1, i2):
        half = np.ones(257)
        half[256] = self.H[256].real                     # the reference reuses fltr_fft: its bin 256 is stale (quisk.c:913-936)
        for on, ctr in ((self.count1 > 0, (i1 + 2) // 4), (self.count1 > 0 and self.count2 > 0, (i2 + 2) // 4)):
            if on:
                lo, hi = max(ctr - self.half_width, 0), min(ctr + self.half_width, 255)
                half[lo:hi + 1] = 0.0
        h = np.fft.irfft(half, 512) * 512                # unnormalised c2r
        taps = np.empty(511)
        taps[255:509] = h[0:254]                         # memmove, quisk.c:938
        taps[509:511] = h[509:511]
        taps[0:255] = taps[510:255:-1]                   # mirror, quisk.c:939-940
        padded = np.zeros(self.N)
        padded[:511] = taps * self.window / 2048 / 4
        self.H = np.fft.rfft(padded)

    def _frame(self):
        X = np.fft.rfft(self.frame)
        self.avg = 0.5 * self.avg + 0.5 * np.abs(X)
        k = np.arange(self.BINS)
        far = np.abs(k - self.signal) > self.delta_sig
        i1 = self._peak(far)
        self.count1 = min(self.count1 + 1, 4) if abs(i1 - self.old1) < 3 else max(self.count1 - 1, -1)
        if self.count1 < 0:
            self.old1 = i1
        i2 = self._peak(far & (np.abs(k - i1) > self.delta_i1))
        self.count2 = min(self.count2 + 1, 4) if abs(i2 - self.old2) < 3 else max(self.count2 - 1, -2)
        if self.count2 < 0:
            self.old2 = i2
        sig = i1 + 10000 * i2 if self.count1 > 0 and self.count2 > 0 else (i1 if self.count1 > 0 else 0)
        if sig != self.sig:
            self.sig = sig
            self._design(i1, i2)
        y = np.fft.irfft(X * self.H, self.N) * self.N    # unnormalised c2r
        self.ready = y[self.START:] / 102                # NOTCH_DATA_SIZE / 20 in integers, quisk.c:958
        self.frame[:self.START] = self.frame[self.N - self.START:]
        self.fill = self.START

    def __call__(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, dtype=np.float64)
        y = np.empty_like(x)
        pos = 0
        while pos < len(x):
            take = min(self.N - self.fill, len(x) - pos)
            y[pos:pos + take] = self.ready[self.fill - self.START:self.fill - self.START + take]
            self.frame[self.fill:self.fill + take] = x[pos:pos + take]
            self.fill += take
            pos += take
            if self.fill == self.N:
                self._frame()
        return y


# --------------------------------------------------------------------------
# ssb_squelch + d_delay (quisk.c:1056-1180): optional stage on the SSB audio at the filter rate
# --------------------------------------------------------------------------

class SsbSquelch:
    """Frames of 512 audio samples -> Hann -> real FFT -> spectral flatness of the bins in [300 Hz, 300 Hz + bw):
    log(mean |X|^2) - mean(log |X|^2) over the bins above 1e-4 (both means divide by the bin count of the band,
    quisk.c:1143-1147).  Above level * 0.005 the one-second timer `sq_open` is re-armed; it is decremented once per
    call by the call's length; squelch_active = timer run out.  The first call only plans the FFT (quisk.c:1104-1112).
    The audio is delayed by one frame."""
    N = 512

    def __init__(self, samp_rate: int, bandwidth: int, level: int):
        self.rate, self.level = samp_rate, level
        bw = min(bandwidth, 3000)
        self.b1 = 300 * self.N // samp_rate
        self.b2 = (bw + 300) * self.N // samp_rate
        self.window = 0.50 - 0.50 * np.cos(2.0 * np.pi * np.arange(self.N) / self.N)
        self.pending = np.zeros(0)
        self.fifo = np.zeros(self.N)
        self.sq_open = 0
        self.active = 0
        self.planned = False

    def __call__(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, dtype=np.float64)
        if self.planned:
            buf = np.concatenate([self.pending, x])
            nfr = len(buf) // self.N
            for f in range(nfr):
                X = np.fft.rfft(buf[f * self.N:(f + 1) * self.N] * self.window)[self.b1:self.b2] / 32767.0
                d = X.real * X.real + X.imag * X.imag
                d = d[d > 1e-4]
                arith = float(np.sum(d))
                ratio = 1.0
                if arith > 1e-4:
                    nb = self.b2 - self.b1
                    ratio = math.log(arith / nb) - float(np.sum(np.log(d))) / nb
                if ratio > self.level * 0.005:
                    self.sq_open = self.rate
            self.pending = buf[nfr * self.N:]
            self.sq_open = max(self.sq_open - len(x), 0)
            self.active = int(self.sq_open == 0)
        self.planned = True
        line = np.concatenate([self.fifo, x])
        self.fifo = line[len(x):]
        return line[:len(x)]


# --------------------------------------------------------------------------
# Synthetic input
# --------------------------------------------------------------------------

def synth_iq(n: int, channel: int = 0, fs: float = 1.0, start: int = 0) -> np.ndarray:
    """Eight complex tones at +-0.01/0.07/0.13/0.31 fs with amplitudes 2^24..2^30
    plus Gaussian noise sigma = 2^20, numpy default_rng(1234 + channel)."""
    rng = np.random.default_rng(1234 + channel)
    fr = np.array([0.01, -0.01, 0.07, -0.07, 0.13, -0.13, 0.31, -0.31])
    amp = 2.0 ** np.linspace(24, 30, 8)
    ph = rng.uniform(0, 2 * math.pi, 8)
    t = np.arange(start, start + n, dtype=np.float64)
    x = np.zeros(n, dtype=np.complex128)
    for f, a, p in zip(fr, amp, ph):
        x += a * np.exp(1j * (2 * math.pi * f * t + p))
    x += (2.0 ** 20) * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x


def rel_rms(a: np.ndarray, b: np.ndarray) -> float:
    """relative RMS error of a against b."""
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.sqrt(np.mean(np.abs(b) ** 2))
    num = np.sqrt(np.mean(np.abs(a - b) ** 2)) if len(a) else 0.0
    return float(num / den) if den > 0 else float(num)


def channelizer_oracle(x: np.ndarray, channels, proto: np.ndarray, n_channels: int, decim: int, n0: int = 0) -> np.ndarray:
    """SURVEY.md section 8 C5 oracle definition: for receiver k, direct-phase mix
    v[n] = x[n] * exp(-2 pi i k n / K) with n the absolute sample index (the reference's tune step,
    quisk.c:2477-2494, with the phase taken from an exact K-entry table instead of the recurrence), then the
    reference's quisk_cDecimate(proto, decim) (filter.c:203-229; FirDecim above).  Returns [len(channels)][frames].
    Cost is len(proto) MACs per output per receiver, so tests run it on a channel subset."""
    K = int(n_channels)
    q = np.arange(K)
    table = np.cos(2.0 * np.pi * q / K) - 1j * np.sin(2.0 * np.pi * q / K)
    n = n0 + np.arange(len(x), dtype=np.int64)
    out = []
    for k in channels:
        v = x * table[(int(k) * n) % K]
        f = FirDecim(proto, decim)
        f.decim_index = n0 % decim
        out.append(f(v))
    return np.stack(out)


def unpack_iq(data, nbytes: int, big_endian: bool) -> np.ndarray:
    """add_rx_samples' unpack loops (quisk.c:2922-2953): packed (I, Q) pairs, nbytes = 1..4 per component.
    Either byte order ends up LEFT-justified in a 32-bit int (little endian: memcpy to the top nbytes of the int,
    :2927-2933; big endian: first byte to the most significant position, :2944-2950).  The store is
    `ii + qq * I` with int operands (:2935,2950): C's `I` is a FLOAT complex, so both ints are converted to float
    (round to nearest even) before they widen to the complex double buffer -- exact up to 3-byte samples, 4-byte
    samples keep 24 significant bits."""
    b = np.frombuffer(bytes(data), dtype=np.uint8).reshape(-1, 2, nbytes).astype(np.uint32)
    v = np.zeros(b.shape[:2], dtype=np.uint32)
    for k in range(nbytes):
        sh = 8 * (3 - k) if big_endian else 8 * (4 - nbytes + k)
        v |= b[:, :, k] << np.uint32(sh)
    v = v.view(np.int32).astype(np.float32).astype(np.float64)
    return v[:, 0] + 1j * v[:, 1]


def unpack_hermes(packet, n_rx: int) -> np.ndarray:
    """The record loop of read_rx_udp10 (quisk.c:3545, 3631, 3746-3763) on one 1032-byte Metis payload:
    two 512-byte frames at bytes 11 and 523 (after 8 header + 3 sync bytes), 5 control bytes, then
    504 // (6 n_rx + 2) records of n_rx x [3 bytes -> imaginary, 3 bytes -> real] (24-bit big endian, << 8)
    and 2 microphone bytes.  Returns [n_rx][2 * records]."""
    buf = np.frombuffer(bytes(packet), dtype=np.uint8).astype(np.int64)
    nrec = 504 // (n_rx * 6 + 2)
    out = np.zeros((n_rx, 2 * nrec), dtype=np.complex128)
    n = 0
    for start in (11, 523):
        index = start + 5
        for _ in range(nrec):
            for r in range(n_rx):
                xi = (buf[index] << 24 | buf[index + 1] << 16 | buf[index + 2] << 8)
                xr = (buf[index + 3] << 24 | buf[index + 4] << 16 | buf[index + 5] << 8)
                xi = xi - (1 << 32) if xi >= (1 << 31) else xi
                xr = xr - (1 << 32) if xr >= (1 << 31) else xr
                out[r, n] = float(xr) + 1j * float(xi)
                index += 6
            n += 1
            index += 2
    return out


# ---------------------------------------------------------------------------------------------------------------
# Waterfall pixel mapper (quisk.c:5334-5480) -- TEST INFRASTRUCTURE like the rest of this file
# ---------------------------------------------------------------------------------------------------------------
class WaterfallOracle:
    """watfall_RgbData / watfall_OnGraphData / watfall_GetPixels restated with an index ring instead of the linked list."""

    def __init__(self, red, green, blue, width, max_height):            # quisk.c:5334-5371
        self.pal = np.stack([np.asarray(red, np.uint8), np.asarray(green, np.uint8), np.asarray(blue, np.uint8)], axis=1)
        self.width, self.H, self.cur = width, max_height, 0
        self.rows = np.zeros((max_height, width, 3), dtype=np.uint8)
        self.xo = np.zeros(max_height, dtype=np.int64)

    def on_graph_data(self, db, y_zero, y_scale, gain, x_origin):        # quisk.c:5373-5420
        self.cur = (self.cur - 1) % self.H                               # current_row = current_row->prior_row
        self.xo[self.cur] = x_origin
        db = np.asarray(db, dtype=np.float64)[:self.width]
        yz = 40.0 + y_zero * 0.69
        t = (db - gain + yz) * float(y_scale + 10) * 0.10 + 128          # numpy rounds every operation separately, like the C code
        l = np.clip(np.trunc(t), 0, 255).astype(np.int64)                # (int) truncates toward zero
        row = self.rows[self.cur]
        row[:] = 0
        row[:len(db)] = self.pal[l]

    def get_pixels(self, x_origin, height, scroll_mode=1):               # quisk.c:5439-5480
        order = []
        k = 0
        if scroll_mode:
            for j in range(8, 1, -1):
                order += [k] * j
                k += 1
                height -= j
        order += list(range(k, k + max(height, 0)))
        out = np.zeros((len(order), self.width, 3), dtype=np.uint8)
        for ro, kk in enumerate(order):
            r = (self.cur + kk) % self.H
            dx = int(self.xo[r] - x_origin)                              # watfall_copy, quisk.c:5422-5437
            if dx == 0:
                out[ro] = self.rows[r]
            elif abs(dx) < self.width:
                if dx > 0:
                    out[ro, dx:] = self.rows[r, :self.width - dx]
                else:
                    out[ro, :self.width + dx] = self.rows[r, -dx:]
        return out.reshape(-1)
