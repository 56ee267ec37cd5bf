"""quisk_b200/rx.py -- thin Python handles over the batched C ABI (quisk_cuda_batch_*,
quisk_cuda_rx_*, quisk_cuda_pan_*).  Device buffers are passed as raw integer pointers
(e.g. torch.Tensor.data_ptr()), streams as integer cudaStream_t handles."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as L

MODES = {"CWL": 0, "CWU": 1, "LSB": 2, "USB": 3, "AM": 4, "FM": 5,
         "DGT-U": 7, "DGT-L": 8, "DGT-IQ": 9, "FDV-U": 11, "FDV-L": 12, "DGT-FM": 13}     # quisk.h:56-70
KINDS = {"cDecim2HB45": 1, "cDecimate": 2, "cCDecimate": 3, "dDecimate": 4, "cInterpolate": 5,
         "dInterpolate": 6, "cInterpDecim": 7, "cInterp2HB45": 8, "dInterp2HB45": 9,
         "cRxFilter": 10, "dRxFilter": 11}


def _dp(a):
    return a.ctypes.data_as(L.c_double_p)


class BatchFilter:
    """quisk_cuda_batch_*: one filter.h filter for n_channels streams, state in HBM."""

    def __init__(self, kind: str, n_channels: int, coefs=None, interp: int = 1, decim: int = 1):
        self.lib = L.require_device()
        self.kind = kind
        if coefs is None:
            c = np.zeros(1); ntaps = 0
        elif kind == "cCDecimate":
            cc = np.ascontiguousarray(coefs, dtype=np.complex128); ntaps = len(cc)
            c = cc.view(np.float64)
        elif kind == "cRxFilter":
            fi, fq = coefs
            c = np.ascontiguousarray(np.concatenate([fi, fq]), dtype=np.float64); ntaps = len(fi)
        else:
            c = np.ascontiguousarray(coefs, dtype=np.float64); ntaps = len(c)
        self.h = self.lib.quisk_cuda_batch_create(KINDS[kind], n_channels, _dp(c), ntaps, interp, decim)
        if not self.h:
            raise L.QuiskCudaError("batch_create: " + self.lib.quisk_cuda_last_error().decode())

    def count_out(self, count: int) -> int:
        return self.lib.quisk_cuda_batch_count_out(self.h, count)

    def run(self, d_in: int, in_stride: int, count: int, d_out: int, out_stride: int, legacy_clip: int = 0, stream: int = 0) -> int:
        n = C.c_int(0)
        L.check(self.lib, self.lib.quisk_cuda_batch_run(self.h, d_in, in_stride, count, d_out, out_stride, C.byref(n), legacy_clip, stream), "batch_run")
        return n.value

    def close(self):
        if self.h:
            self.lib.quisk_cuda_batch_destroy(self.h); self.h = None

    __del__ = close


class RxChain:
    """quisk_cuda_rx_*: tune -> quisk_process_decimate -> quisk_process_demodulate for a batch."""

    def __init__(self, n_channels: int, sample_rate: int, mode: str, filt_i, filt_q, tables: dict,
                 tune_hz=None, fused: bool = True, bandwidth: int = 2800):
        self.lib = L.require_device()
        self._keep = []
        cfg = L.RxConfig()
        cfg.n_channels = n_channels; cfg.sample_rate = sample_rate; cfg.mode = MODES[mode]; cfg.fused = int(fused)
        cfg.filter_bandwidth = int(bandwidth)
        fi = np.ascontiguousarray(filt_i, dtype=np.float64); fq = np.ascontiguousarray(filt_q, dtype=np.float64)
        self._keep += [fi, fq]
        cfg.filt_i = _dp(fi); cfg.filt_q = _dp(fq); cfg.n_filt = len(fi)
        if tune_hz is not None:
            t = np.ascontiguousarray(np.broadcast_to(np.asarray(tune_hz, dtype=np.float64), (n_channels,)))
            self._keep.append(t); cfg.tune_hz = _dp(t)
        for field, ref_name in L.TABLE_NAMES.items():
            if ref_name in tables:
                a = np.ascontiguousarray(tables[ref_name], dtype=np.float64)
                self._keep.append(a)
                setattr(cfg.tables, field, _dp(a)); setattr(cfg.tables, "n_" + field, len(a))
        self.n_channels = n_channels
        self.h = self.lib.quisk_cuda_rx_create(C.byref(cfg))
        if not self.h:
            raise L.QuiskCudaError("rx_create: " + self.lib.quisk_cuda_last_error().decode())

    @property
    def decim_srate(self): return self.lib.quisk_cuda_rx_decim_srate(self.h)

    @property
    def filter_srate(self): return self.lib.quisk_cuda_rx_filter_srate(self.h)

    def max_out(self, count: int) -> int: return self.lib.quisk_cuda_rx_max_out(self.h, count)

    def process(self, d_iq: int, iq_stride: int, count: int, d_audio: int, audio_stride: int,
                d_decim: int = 0, decim_stride: int = 0, stream: int = 0):
        na, nd = C.c_int(0), C.c_int(0)
        L.check(self.lib, self.lib.quisk_cuda_rx_process(self.h, d_iq, iq_stride, count, d_audio, audio_stride, C.byref(na),
                                                         d_decim or None, decim_stride, C.byref(nd), stream), "rx_process")
        return na.value, nd.value

    def process_host(self, h_iq: np.ndarray, count: int, h_audio: np.ndarray) -> int:
        na = C.c_int(0)
        L.check(self.lib, self.lib.quisk_cuda_rx_process_host(self.h, h_iq.ctypes.data, h_iq.strides[0] // 16, count,
                                                              h_audio.ctypes.data, h_audio.strides[0] // 8, C.byref(na)), "rx_process_host")
        return na.value

    def process_host_packed(self, h_bytes: np.ndarray, count: int, nbytes: int, big_endian: bool, h_audio: np.ndarray) -> int:
        """h_bytes: [n_channels, >= count*2*nbytes] uint8 in the wire format of add_rx_samples (quisk.c:2922-2953)."""
        na = C.c_int(0)
        L.check(self.lib, self.lib.quisk_cuda_rx_process_host_packed(self.h, h_bytes.ctypes.data, h_bytes.strides[0], count, nbytes, int(big_endian),
                                                                     h_audio.ctypes.data, h_audio.strides[0] // 8, C.byref(na)), "rx_process_host_packed")
        return na.value

    def reset(self): L.check(self.lib, self.lib.quisk_cuda_rx_reset(self.h), "rx_reset")

    def set_option(self, option: int, value: int):
        L.check(self.lib, self.lib.quisk_cuda_rx_set_option(self.h, option, value), "rx_set_option")

    def kernel_time(self):
        """(total ms, launches) of the event-timed dominant kernel since the last call."""
        ms, n = C.c_double(0), C.c_int(0)
        L.check(self.lib, self.lib.quisk_cuda_rx_kernel_time(self.h, C.byref(ms), C.byref(n)), "rx_kernel_time")
        return ms.value, n.value

    def fused_kernel_name(self):
        """Template instantiation of the fused decimator the last process() launched ("" on the per-stage path)."""
        return self.lib.quisk_cuda_rx_fused_kernel_name(self.h).decode()

    def close(self):
        if self.h:
            self.lib.quisk_cuda_rx_destroy(self.h); self.h = None

    __del__ = close


def load_tables(path: str | None = None) -> dict:
    """The coefficient tables the chain needs by name (filters.h: quiskFilt48dec24Coefs ...) and the 24 kS/s
    prototype low-pass tables of filters.py (keys proto_<n>), shipped as package data in
    quisk_b200/data/quisk_tables.npz (extracted from the compiled reference by tests/golden/make_golden.py).  In a
    drop-in build the caller links the reference's own filters.h instead."""
    import os
    path = path or os.path.join(L.HERE, "data", "quisk_tables.npz")
    z = np.load(path)
    return {k: z[k] for k in z.files}


def get_filter_center(mode: str, bandwidth: int, cw_tone: int = 600) -> int:
    """GetFilterCenter (quisk.py:5464-5488): centre of the mode's I/Q pass band in Hz, negative on the lower side
    band (MakeFilterCoef takes its absolute value; the sign picks re + im or re - im in the demodulator)."""
    table = {"CWU": max(cw_tone, bandwidth // 2), "CWL": max(cw_tone, bandwidth // 2), "AM": 0, "FM": 0,
             "DGT-U": max(1500, bandwidth // 2), "DGT-L": max(1500, bandwidth // 2), "DGT-IQ": 0, "DGT-FM": 0,
             "FDV-U": 1500 if bandwidth <= 3000 else bandwidth // 2, "FDV-L": 1500 if bandwidth <= 3000 else bandwidth // 2}
    center = table.get(mode, 300 + bandwidth // 2)            # LSB / USB / IMD and anything else
    return -center if mode in ("CWL", "LSB", "DGT-L", "FDV-L") else center


def make_filter_coef(rate: int, N, bw: int, center: int, tables: dict | None = None):
    """MakeFilterCoef (quisk.py:5405-5456) -> (filtI, filtQ): quisk_cuda_make_filter_coef with the filters.py
    prototype for key bw * 24000 // rate // 2 when there is one."""
    lib = L.load()
    tables = tables if tables is not None else load_tables()
    proto = tables.get("proto_%d" % lib.quisk_cuda_filter_key(int(rate), int(bw)))
    cap = 10001                                     # MAX_FILTER_SIZE, quisk.h:10
    fi = np.zeros(cap); fq = np.zeros(cap); n = C.c_int(0)
    if proto is not None:
        proto = np.ascontiguousarray(proto, dtype=np.float64)
    L.check(lib, lib.quisk_cuda_make_filter_coef(int(rate), int(N) if N else 0, int(bw), int(center),
                                                 _dp(proto) if proto is not None else None, len(proto) if proto is not None else 0,
                                                 _dp(fi), _dp(fq), cap, C.byref(n)), "make_filter_coef")
    return fi[:n.value].copy(), fq[:n.value].copy()


class Channelizer:
    """Wideband polyphase channelizer (quisk_cuda_pfb_*): one stream -> n_channels receivers at k*fs/n_channels,
    each equal to the reference's tune (quisk.c:2477-2494) + quisk_cDecimate(proto, decim) (filter.c:203-229)."""

    def __init__(self, n_channels: int, decim: int, proto):
        self.lib = L.require_device()
        self.proto = np.ascontiguousarray(proto, dtype=np.float64)
        self.n_channels, self.decim = n_channels, decim
        self.h = self.lib.quisk_cuda_pfb_create(n_channels, decim, _dp(self.proto), len(self.proto))
        if not self.h:
            raise L.QuiskCudaError("pfb_create: " + self.lib.quisk_cuda_last_error().decode())

    def count_out(self, count: int) -> int: return self.lib.quisk_cuda_pfb_count_out(self.h, count)

    def set_option(self, option: int, value: int):
        L.check(self.lib, self.lib.quisk_cuda_pfb_set_option(self.h, option, value), "pfb_set_option")

    def seek(self, n_abs: int, stream: int | None = None):
        """Restart at an absolute sample index with an empty history.  Without a stream the device is synchronised first;
        with one the reset is ordered on that stream like the prime / process calls around it."""
        if stream is None:
            L.check(self.lib, self.lib.quisk_cuda_pfb_seek(self.h, n_abs), "pfb_seek")
        else:
            L.check(self.lib, self.lib.quisk_cuda_pfb_seek_async(self.h, n_abs, stream), "pfb_seek_async")

    def prime(self, d_in: int, count: int, stream: int = 0):
        L.check(self.lib, self.lib.quisk_cuda_pfb_prime(self.h, d_in, count, stream), "pfb_prime")

    def process(self, d_in: int, count: int, d_out: int, out_stride: int, layout: int = 0, stream: int = 0) -> int:
        nf = C.c_int(0)
        L.check(self.lib, self.lib.quisk_cuda_pfb_process(self.h, d_in, count, d_out, out_stride, layout, C.byref(nf), stream), "pfb_process")
        return nf.value

    def close(self):
        if self.h:
            self.lib.quisk_cuda_pfb_destroy(self.h); self.h = None

    __del__ = close


class Panadapter:
    """Host-side mirror of the reference's panadapter interface for a batch of streams: `record_app(..., data_width,
    ..., fft_size, ..., rate, ...)` sets the sizes (quisk.c:5946-6009), the sample thread fills FFT buffers
    (quisk.c:2454-2475) and `get_graph(job, zoom, deltaf)` returns `data_width` dB values or None when nothing has
    been averaged yet (quisk.c:5142-5331).  Frames live in device memory; graphs come back as a NumPy array
    [n_streams, data_width]."""

    def __init__(self, n_streams: int, fft_size: int, data_width: int, sample_rate: float):
        self.lib = L.require_device()
        self.n_streams, self.fft_size, self.data_width, self.sample_rate = n_streams, fft_size, data_width, float(sample_rate)
        self.h = self.lib.quisk_cuda_pan_create(n_streams, fft_size)
        if not self.h:
            raise L.QuiskCudaError("pan_create: " + self.lib.quisk_cuda_last_error().decode())
        self._graph = None

    def add_frames(self, d_frames: int, stream_stride: int, n_frames: int, stream: int = 0):
        """d_frames: device pointer to [n_streams][stream_stride] complex128, n_frames consecutive frames per stream."""
        L.check(self.lib, self.lib.quisk_cuda_pan_accumulate(self.h, d_frames, stream_stride, n_frames, stream), "pan_accumulate")

    @property
    def count_fft(self) -> int: return self.lib.quisk_cuda_pan_count(self.h)

    def get_graph(self, zoom: float = 1.0, deltaf: float = 0.0):
        """None if no frame has been accumulated since the last graph (the reference returns None too), else the
        graph of every stream; the running averages restart, as in the reference."""
        if self.count_fft <= 0:
            return None
        import torch
        if self._graph is None:
            self._graph = torch.zeros((self.n_streams, self.data_width), dtype=torch.float64, device="cuda")
        L.check(self.lib, self.lib.quisk_cuda_pan_graph(self.h, self.data_width, zoom, deltaf, self.sample_rate,
                                                        self._graph.data_ptr(), None), "pan_graph")
        torch.cuda.synchronize()
        return self._graph.cpu().numpy()

    def close(self):
        if self.h:
            self.lib.quisk_cuda_pan_destroy(self.h); self.h = None

    __del__ = close


class TxTables(C.Structure):
    _fields_ = [("mic_filt8", C.POINTER(C.c_double)), ("n_mic_filt8", C.c_int),
                ("lp_filt48", C.POINTER(C.c_double)), ("n_lp_filt48", C.c_int),
                ("tx8k_audio", C.POINTER(C.c_double)), ("n_tx8k_audio", C.c_int),
                ("dgt_filt48", C.POINTER(C.c_double)), ("n_dgt_filt48", C.c_int)]


class TxFilter:
    """The transmit-audio chain of the reference's microphone.c for a batch of transmitters: `tx_filter`
    (microphone.c:372-604; what `quisk_process_microphone` runs on the microphone block at :1232) with its peak rounder
    `CcmPeak` (:161-233); in the digital modes "DGT-U", "DGT-L", "FDV-U", "FDV-L" it is `tx_filter_digital` (:605-624: one tuned
    filter at 48 kS/s).  mode: "LSB", "USB" (complex I/Q out), "AM", "FM" (real rail out).  `preemphasis` and `clip` are
    the reference's `quisk_mic_preemphasis` and `quisk_mic_clip`.  process(): device pointers, [n_channels][stride]
    complex128 in (microphone audio on the real rail, +-CLIP16) and out (48 kS/s); returns the samples per channel."""

    def __init__(self, n_channels: int, mode: str, tables: dict, mic_sample_rate: int = 48000, preemphasis: float = 0.6, clip: float = 1.0):
        self.lib = L.require_device()
        self.n_channels = n_channels
        self._keep = [np.ascontiguousarray(tables[k], dtype=np.float64) for k in ("quiskMicFilt8Coefs", "quiskLpFilt48Coefs", "quiskFiltTx8kAudioB", "quiskDgtFilt48Coefs")]
        t = TxTables()
        for (name, cnt), a in zip((("mic_filt8", "n_mic_filt8"), ("lp_filt48", "n_lp_filt48"), ("tx8k_audio", "n_tx8k_audio"), ("dgt_filt48", "n_dgt_filt48")), self._keep):
            setattr(t, name, a.ctypes.data_as(C.POINTER(C.c_double))); setattr(t, cnt, len(a))
        self.h = self.lib.quisk_cuda_tx_filter_create(n_channels, MODES[mode], mic_sample_rate, preemphasis, clip, C.byref(t))
        if not self.h:
            raise L.QuiskCudaError("tx_filter_create: " + self.lib.quisk_cuda_last_error().decode())

    def set_alc(self, enable: int = 1):
        """process_alc behind the filter (microphone.c:1232-1233); every call with 1 is a key down (init_alc(&tx_alc, 0), :1207)."""
        L.check(self.lib, self.lib.quisk_cuda_tx_filter_set_alc(self.h, int(enable)), "tx_filter_set_alc")

    def max_out(self, count: int) -> int:
        return self.lib.quisk_cuda_tx_filter_max_out(self.h, count)

    def process(self, d_in: int, in_stride: int, count: int, d_out: int, out_stride: int, stream: int = 0) -> int:
        n = C.c_int(0)
        L.check(self.lib, self.lib.quisk_cuda_tx_filter_process(self.h, d_in, in_stride, count, d_out, out_stride, C.byref(n), stream or None), "tx_filter_process")
        return n.value

    def close(self):
        if self.h:
            self.lib.quisk_cuda_tx_filter_destroy(self.h)
            self.h = None


class RxOptions:
    """Host-side mirror of the three optional receive stages Quisk switches from Python -- `set_noise_blanker(level)`
    (quisk.c:4605-4611), `set_auto_notch(on)` (quisk.c:4596-4603) and `set_ssb_squelch(enabled, level)`
    (quisk.c:4729-4735) -- for a batch of receivers.  The setters record the values like the reference's
    (`set_auto_notch` also re-initialises the notch state, as its `dAutoNotch(NULL, ...)` call does); the `run_*` methods are what the sample
    thread does with them: `run_noise_blanker` on the raw block in front of the tuning stage (quisk.c:2448-2449),
    `run_auto_notch` and `run_ssb_squelch` on the SSB audio at the filter rate (quisk.c:1923-1928).  All in place on
    device memory ([n_channels][stride], complex128 for the blanker, float64 for the audio stages)."""

    def __init__(self, n_channels: int, sample_rate: int, filter_srate: int, filter_bandwidth: int):
        self.lib = L.require_device()
        self.n_channels, self.filter_srate = n_channels, filter_srate
        self.noise_blanker, self.auto_notch, self.ssb_squelch_enabled, self.ssb_squelch_level = 0, 0, 0, 0
        lib = self.lib
        self.nb = lib.quisk_cuda_nb_create(n_channels, sample_rate)
        self.an = lib.quisk_cuda_autonotch_create(n_channels, filter_srate)
        self.sq = lib.quisk_cuda_ssb_squelch_create(n_channels, filter_srate, filter_bandwidth)
        if not (self.nb and self.an and self.sq):
            raise L.QuiskCudaError("RxOptions: " + lib.quisk_cuda_last_error().decode())

    def set_noise_blanker(self, level: int): self.noise_blanker = int(level)

    def set_auto_notch(self, on: int):
        self.auto_notch = int(on)
        self.lib.quisk_cuda_autonotch_destroy(self.an)         # dAutoNotch(NULL, 0, 0, 0), quisk.c:4600
        self.an = self.lib.quisk_cuda_autonotch_create(self.n_channels, self.filter_srate)
        if not self.an:
            raise L.QuiskCudaError("autonotch_create: " + self.lib.quisk_cuda_last_error().decode())

    def set_ssb_squelch(self, enabled: int, level: int): self.ssb_squelch_enabled, self.ssb_squelch_level = int(enabled), int(level)

    def run_noise_blanker(self, d_iq: int, stride: int, count: int, stream: int = 0):
        """NoiseBlanker(cSamples, nSamples): does nothing while the level is 0, like the reference (quisk.c:695)."""
        L.check(self.lib, self.lib.quisk_cuda_nb_run(self.nb, d_iq, stride, count, self.noise_blanker, stream or None), "nb_run")

    def run_auto_notch(self, d_audio: int, stride: int, count: int, sidetone: int = 0, stream: int = 0):
        """dAutoNotch(dsamples, nSamples, rit_freq, quisk_filter_srate): skipped while quisk_auto_notch is 0 (quisk.c:835)."""
        if self.auto_notch:
            L.check(self.lib, self.lib.quisk_cuda_autonotch_run(self.an, d_audio, stride, count, sidetone, stream or None), "autonotch_run")

    def run_ssb_squelch(self, d_audio: int, stride: int, count: int, stream: int = 0):
        """ssb_squelch + d_delay when ssb_squelch_enabled (quisk.c:1925-1928); returns squelch_active per channel or None."""
        if not self.ssb_squelch_enabled:
            return None
        L.check(self.lib, self.lib.quisk_cuda_ssb_squelch_run(self.sq, d_audio, stride, count, self.ssb_squelch_level, stream or None), "ssb_squelch_run")
        act = (C.c_int * self.n_channels)()
        L.check(self.lib, self.lib.quisk_cuda_ssb_squelch_state(self.sq, None, act, stream or None), "ssb_squelch_state")
        return list(act)

    def close(self):
        if getattr(self, "nb", None):
            self.lib.quisk_cuda_nb_destroy(self.nb); self.nb = None
        if getattr(self, "an", None):
            self.lib.quisk_cuda_autonotch_destroy(self.an); self.an = None
        if getattr(self, "sq", None):
            self.lib.quisk_cuda_ssb_squelch_destroy(self.sq); self.sq = None

    __del__ = close
