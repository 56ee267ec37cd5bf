"""quisk_b200/lib.py -- ctypes loader and prototypes for libquisk_cuda.so (include/quisk_cuda.h)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libquisk_cuda.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class QuiskCudaError(RuntimeError):
    pass


class RxTables(C.Structure):        # struct qcRxTables
    _names = ["filt144D3", "filt240D5Sharp", "filt48dec24", "filt300D5", "audio24p4", "audio24p6",
              "lpFilt48", "audioFmHp", "filt53D1", "filt111D2", "filt133D2", "filt167D3", "filt185D3"]
    _fields_ = sum(([(n, c_double_p), ("n_" + n, C.c_int)] for n in _names), [])


class RxConfig(C.Structure):        # struct qcRxConfig
    _fields_ = [("n_channels", C.c_int), ("sample_rate", C.c_int), ("mode", C.c_int),
                ("filt_i", c_double_p), ("filt_q", c_double_p), ("n_filt", C.c_int),
                ("tune_hz", c_double_p), ("tables", RxTables), ("fused", C.c_int), ("filter_bandwidth", C.c_int)]


# qcRxTables field -> the reference's table name in filters.h
TABLE_NAMES = {
    "filt144D3": "quiskFilt144D3Coefs", "filt240D5Sharp": "quiskFilt240D5CoefsSharp",
    "filt48dec24": "quiskFilt48dec24Coefs", "filt300D5": "quiskFilt300D5Coefs",
    "audio24p4": "quiskAudio24p4Coefs", "audio24p6": "quiskAudio24p6Coefs",
    "lpFilt48": "quiskLpFilt48Coefs", "audioFmHp": "quiskAudioFmHpCoefs",
    "filt53D1": "quiskFilt53D1Coefs", "filt111D2": "quiskFilt111D2Coefs",
    "filt133D2": "quiskFilt133D2Coefs", "filt167D3": "quiskFilt167D3Coefs",
    "filt185D3": "quiskFilt185D3Coefs",
}

_lib = None


def load() -> C.CDLL:
    """dlopen libquisk_cuda.so and declare the quisk_cuda_* prototypes.  Raises
    QuiskCudaError when the library has not been built -- there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QuiskCudaError(f"{LIB_PATH} not found: run `python -m quisk_b200.build` (nvcc, sm_100a)")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.quisk_cuda_last_error.restype = C.c_char_p
    lib.quisk_cuda_version.restype = C.c_char_p
    lib.quisk_cuda_launch_count.restype = C.c_ulonglong
    lib.quisk_cuda_set_device.argtypes = [C.c_int]
    lib.quisk_cuda_batch_create.argtypes = [C.c_int, C.c_int, c_double_p, C.c_int, C.c_int, C.c_int]
    lib.quisk_cuda_batch_create.restype = vp
    lib.quisk_cuda_batch_destroy.argtypes = [vp]
    lib.quisk_cuda_batch_destroy.restype = None
    lib.quisk_cuda_batch_count_out.argtypes = [vp, C.c_int]
    lib.quisk_cuda_batch_run.argtypes = [vp, vp, C.c_long, C.c_int, vp, C.c_long, c_int_p, C.c_int, vp]
    lib.quisk_cuda_batch_reset.argtypes = [vp, vp]
    lib.quisk_cuda_plan_decimation.argtypes = [C.c_int, c_int_p, c_int_p, c_int_p]
    lib.quisk_cuda_rx_create.argtypes = [C.POINTER(RxConfig)]
    lib.quisk_cuda_rx_create.restype = vp
    lib.quisk_cuda_rx_destroy.argtypes = [vp]
    lib.quisk_cuda_rx_destroy.restype = None
    lib.quisk_cuda_rx_decim_srate.argtypes = [vp]
    lib.quisk_cuda_rx_filter_srate.argtypes = [vp]
    lib.quisk_cuda_rx_squelch_active.argtypes = [vp, vp]
    lib.quisk_cuda_rx_max_out.argtypes = [vp, C.c_int]
    lib.quisk_cuda_rx_process.argtypes = [vp, vp, C.c_long, C.c_int, vp, C.c_long, c_int_p, vp, C.c_long, c_int_p, vp]
    lib.quisk_cuda_rx_process_host.argtypes = [vp, vp, C.c_long, C.c_int, vp, C.c_long, c_int_p]
    lib.quisk_cuda_rx_reset.argtypes = [vp]
    lib.quisk_cuda_rx_set_option.argtypes = [vp, C.c_int, C.c_int]
    lib.quisk_cuda_rx_kernel_time.argtypes = [vp, c_double_p, c_int_p]
    lib.quisk_cuda_rx_fused_kernel_name.argtypes = [vp]
    lib.quisk_cuda_rx_fused_kernel_name.restype = C.c_char_p
    lib.quisk_cuda_rx_read_trace.argtypes = [vp, vp, C.c_int]
    lib.quisk_cuda_tx_filter_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, vp]
    lib.quisk_cuda_tx_filter_create.restype = vp
    lib.quisk_cuda_tx_filter_destroy.argtypes = [vp]
    lib.quisk_cuda_tx_filter_destroy.restype = None
    lib.quisk_cuda_tx_filter_max_out.argtypes = [vp, C.c_int]
    lib.quisk_cuda_tx_filter_set_alc.argtypes = [vp, C.c_int]
    lib.quisk_cuda_tx_filter_process.argtypes = [vp, vp, C.c_long, C.c_int, vp, C.c_long, c_int_p, vp]
    lib.quisk_cuda_pan_create.argtypes = [C.c_int, C.c_int]
    lib.quisk_cuda_fp64_peak.argtypes = [C.POINTER(C.c_double)]
    lib.quisk_cuda_pan_create.restype = vp
    lib.quisk_cuda_waterfall_create.argtypes = [C.c_int, C.c_int, C.c_int, vp, vp, vp]
    lib.quisk_cuda_waterfall_create.restype = vp
    lib.quisk_cuda_waterfall_destroy.argtypes = [vp]; lib.quisk_cuda_waterfall_destroy.restype = None
    lib.quisk_cuda_waterfall_on_graph_data.argtypes = [vp, vp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, vp]
    lib.quisk_cuda_waterfall_get_pixels.argtypes = [vp, vp, C.c_long, C.c_int, C.c_int, C.c_int, vp]
    lib.quisk_cuda_pan_destroy.argtypes = [vp]
    lib.quisk_cuda_pan_destroy.restype = None
    lib.quisk_cuda_pan_accumulate.argtypes = [vp, vp, C.c_long, C.c_int, vp]
    lib.quisk_cuda_pan_graph.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double, vp, vp]
    lib.quisk_cuda_pan_multirx.argtypes = [vp, vp, C.c_long, vp, vp]
    lib.quisk_cuda_pan_count.argtypes = [vp]
    lib.quisk_cuda_pan_average_ptr.argtypes = [vp]
    lib.quisk_cuda_pan_average_ptr.restype = vp
    lib.quisk_cuda_fft_batch.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
    lib.quisk_cuda_bandscope_create.argtypes = [C.c_int, C.c_int]; lib.quisk_cuda_bandscope_create.restype = vp
    lib.quisk_cuda_bandscope_destroy.argtypes = [vp]; lib.quisk_cuda_bandscope_destroy.restype = None
    lib.quisk_cuda_bandscope_accumulate.argtypes = [vp, vp, C.c_long, C.c_int, vp]
    lib.quisk_cuda_bandscope_graph.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp]
    lib.quisk_cuda_agc_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]; lib.quisk_cuda_agc_create.restype = vp
    lib.quisk_cuda_agc_destroy.argtypes = [vp]; lib.quisk_cuda_agc_destroy.restype = None
    lib.quisk_cuda_agc_run.argtypes = [vp, vp, C.c_long, C.c_int, C.c_int, vp]
    lib.quisk_cuda_fracdecim_create.argtypes = [C.c_int]; lib.quisk_cuda_fracdecim_create.restype = vp
    lib.quisk_cuda_fracdecim_destroy.argtypes = [vp]; lib.quisk_cuda_fracdecim_destroy.restype = None
    lib.quisk_cuda_fracdecim_run.argtypes = [vp, vp, C.c_long, C.c_int, C.c_double, vp, C.c_long, c_int_p, vp]
    lib.quisk_cuda_nb_create.argtypes = [C.c_int, C.c_int]; lib.quisk_cuda_nb_create.restype = vp
    lib.quisk_cuda_nb_destroy.argtypes = [vp]; lib.quisk_cuda_nb_destroy.restype = None
    lib.quisk_cuda_nb_run.argtypes = [vp, vp, C.c_long, C.c_int, C.c_int, vp]
    lib.quisk_cuda_ssb_squelch_create.argtypes = [C.c_int, C.c_int, C.c_int]; lib.quisk_cuda_ssb_squelch_create.restype = vp
    lib.quisk_cuda_ssb_squelch_destroy.argtypes = [vp]; lib.quisk_cuda_ssb_squelch_destroy.restype = None
    lib.quisk_cuda_ssb_squelch_run.argtypes = [vp, vp, C.c_long, C.c_int, C.c_int, vp]
    lib.quisk_cuda_ssb_squelch_state.argtypes = [vp, c_int_p, c_int_p, vp]
    lib.quisk_cuda_ssb_squelch_state_ptr.argtypes = [vp]; lib.quisk_cuda_ssb_squelch_state_ptr.restype = vp
    lib.quisk_cuda_autonotch_create.argtypes = [C.c_int, C.c_int]; lib.quisk_cuda_autonotch_create.restype = vp
    lib.quisk_cuda_autonotch_destroy.argtypes = [vp]; lib.quisk_cuda_autonotch_destroy.restype = None
    lib.quisk_cuda_autonotch_run.argtypes = [vp, vp, C.c_long, C.c_int, C.c_int, vp]
    lib.quisk_cuda_unpack_iq.argtypes = [vp, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_long, vp]
    lib.quisk_cuda_hermes_samples_per_packet.argtypes = [C.c_int]
    lib.quisk_cuda_unpack_hermes.argtypes = [vp, C.c_int, C.c_int, vp, C.c_long, c_int_p, vp]
    lib.quisk_cuda_rx_process_host_packed.argtypes = [vp, vp, C.c_long, C.c_int, C.c_int, C.c_int, vp, C.c_long, c_int_p]
    lib.quisk_cuda_pfb_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int]; lib.quisk_cuda_pfb_create.restype = vp
    lib.quisk_cuda_pfb_destroy.argtypes = [vp]; lib.quisk_cuda_pfb_destroy.restype = None
    lib.quisk_cuda_pfb_count_out.argtypes = [vp, C.c_int]
    lib.quisk_cuda_pfb_set_option.argtypes = [vp, C.c_int, C.c_int]
    lib.quisk_cuda_pfb_seek.argtypes = [vp, C.c_longlong]
    lib.quisk_cuda_pfb_seek_async.argtypes = [vp, C.c_longlong, vp]
    lib.quisk_cuda_filter_key.argtypes = [C.c_int, C.c_int]
    lib.quisk_cuda_make_filter_coef.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, C.c_int, c_double_p, c_double_p, C.c_int, c_int_p]
    lib.quisk_cuda_pfb_prime.argtypes = [vp, vp, C.c_int, vp]
    lib.quisk_cuda_pfb_process.argtypes = [vp, vp, C.c_int, vp, C.c_long, C.c_int, c_int_p, vp]
    # ---- WDSP RXA part (include/quisk_cuda_wdsp.h) ----
    D = C.c_double
    lib.quisk_cuda_fir_bandpass.argtypes = [C.c_int, D, D, D, C.c_int, C.c_int, D, vp]
    lib.quisk_cuda_fc_impulse.argtypes = [C.c_int, D, D, D, D, C.c_int, D, D, C.c_int, C.c_int, vp]
    lib.quisk_cuda_resample_design.argtypes = [C.c_int, C.c_int, D, C.c_int, D, c_int_p, c_int_p, c_int_p, vp, C.c_int]
    lib.quisk_cuda_fircore_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp]
    lib.quisk_cuda_fircore_create.restype = vp
    lib.quisk_cuda_firopt_create.argtypes = [C.c_int, C.c_int, C.c_int, D, D, C.c_int, C.c_int, D]
    lib.quisk_cuda_firopt_create.restype = vp
    lib.quisk_cuda_bps_create.argtypes = [C.c_int, C.c_int, D, D, C.c_int, C.c_int, D]
    lib.quisk_cuda_bps_create.restype = vp
    lib.quisk_cuda_firmin_create.argtypes = [C.c_int, C.c_int, D, D, C.c_int, C.c_int, D]
    lib.quisk_cuda_firmin_create.restype = vp
    lib.quisk_cuda_fircore_destroy.argtypes = [vp]; lib.quisk_cuda_fircore_destroy.restype = None
    lib.quisk_cuda_fircore_run.argtypes = [vp, vp, C.c_long, vp, C.c_long, vp]
    lib.quisk_cuda_fircore_set_impulse.argtypes = [vp, vp, C.c_int]
    lib.quisk_cuda_fircore_update.argtypes = [vp]
    lib.quisk_cuda_fircore_flush.argtypes = [vp]
    lib.quisk_cuda_resample_create.argtypes = [C.c_int, C.c_int, C.c_int, D, C.c_int, D]
    lib.quisk_cuda_resample_create.restype = vp
    lib.quisk_cuda_resample_destroy.argtypes = [vp]; lib.quisk_cuda_resample_destroy.restype = None
    lib.quisk_cuda_resample_count_out.argtypes = [vp, C.c_int]
    lib.quisk_cuda_resample_run.argtypes = [vp, vp, C.c_long, C.c_int, vp, C.c_long, c_int_p, vp]
    lib.quisk_cuda_shift_create.argtypes = [C.c_int, C.c_int, vp]; lib.quisk_cuda_shift_create.restype = vp
    lib.quisk_cuda_wcpagc_create.argtypes = [C.c_int, C.c_int, C.c_int]; lib.quisk_cuda_wcpagc_create.restype = vp
    lib.quisk_cuda_wcpagc_create_fmlim.argtypes = [C.c_int, C.c_int, D]; lib.quisk_cuda_wcpagc_create_fmlim.restype = vp
    lib.quisk_cuda_wcpagc_set_fixed_gain_db.argtypes = [vp, D]
    lib.quisk_cuda_wcpagc_set_top_db.argtypes = [vp, D]
    lib.quisk_cuda_amd_create.argtypes = [C.c_int] * 5; lib.quisk_cuda_amd_create.restype = vp
    lib.quisk_cuda_fmpll_create.argtypes = [C.c_int, C.c_int, D, D, D, D, D, D]; lib.quisk_cuda_fmpll_create.restype = vp
    lib.quisk_cuda_snotch_create.argtypes = [C.c_int, C.c_int, D, D]; lib.quisk_cuda_snotch_create.restype = vp
    lib.quisk_cuda_seq_destroy.argtypes = [vp]; lib.quisk_cuda_seq_destroy.restype = None
    lib.quisk_cuda_seq_run.argtypes = [vp, vp, C.c_long, vp, C.c_long, C.c_int, vp]
    lib.quisk_cuda_seq_flush.argtypes = [vp]
    lib.quisk_cuda_rxa_create.argtypes = [C.c_int] * 6; lib.quisk_cuda_rxa_create.restype = vp
    lib.quisk_cuda_rxa_destroy.argtypes = [vp]; lib.quisk_cuda_rxa_destroy.restype = None
    lib.quisk_cuda_rxa_set_mode.argtypes = [vp, C.c_int]
    lib.quisk_cuda_rxa_set_passband.argtypes = [vp, D, D]
    lib.quisk_cuda_rxa_set_nc.argtypes = [vp, C.c_int]
    lib.quisk_cuda_rxa_set_agc_mode.argtypes = [vp, C.c_int]
    lib.quisk_cuda_rxa_set_agc_fixed.argtypes = [vp, D]
    lib.quisk_cuda_rxa_set_shift.argtypes = [vp, C.c_int, vp]
    lib.quisk_cuda_rxa_set_nbp_run.argtypes = [vp, C.c_int]
    lib.quisk_cuda_rxa_set_slew.argtypes = [vp, C.c_double, C.c_double]
    lib.quisk_cuda_rxa_set_siphon_run.argtypes = [vp, C.c_int]
    lib.quisk_cuda_rxa_get_siphon.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.quisk_cuda_nbp_impulse.argtypes = [C.c_int, D, D, D, C.c_int, D, C.c_int, vp, vp, vp, D, D, C.c_int, C.c_int, vp, c_int_p, c_int_p]
    lib.quisk_cuda_rxa_nbp_add_notch.argtypes = [vp, C.c_int, D, D, C.c_int]
    lib.quisk_cuda_rxa_nbp_delete_notch.argtypes = [vp, C.c_int]
    lib.quisk_cuda_rxa_nbp_set_notches_run.argtypes = [vp, C.c_int]
    lib.quisk_cuda_rxa_nbp_set_tune_frequency.argtypes = [vp, D]
    lib.quisk_cuda_rxa_nbp_set_shift_frequency.argtypes = [vp, D]
    lib.quisk_cuda_rxa_set_snba_run.argtypes = [vp, C.c_int]
    lib.quisk_cuda_analyzer_create.restype = vp
    lib.quisk_cuda_analyzer_create.argtypes = [C.c_int, C.c_int]
    lib.quisk_cuda_analyzer_destroy.restype = None
    lib.quisk_cuda_analyzer_destroy.argtypes = [vp]
    lib.quisk_cuda_analyzer_set.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
    for fn in ("set_detector_mode", "set_average_mode", "set_num_average", "set_norm_onehz"):
        getattr(lib, "quisk_cuda_analyzer_" + fn).argtypes = [vp, C.c_int, C.c_int]
    lib.quisk_cuda_analyzer_set_av_backmult.argtypes = [vp, C.c_int, C.c_double]
    lib.quisk_cuda_analyzer_set_sample_rate.argtypes = [vp, C.c_int]
    lib.quisk_cuda_analyzer_spectrum0.argtypes = [vp, C.c_int, vp, C.c_long, vp]
    lib.quisk_cuda_analyzer_get_pixels.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_int)]
    lib.quisk_cuda_analyzer_get_enb.restype = C.c_double
    lib.quisk_cuda_analyzer_get_enb.argtypes = [vp]
    lib.quisk_cuda_snba_create.argtypes = [C.c_int] * 8 + [C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
    lib.quisk_cuda_snba_create.restype = vp
    lib.quisk_cuda_snba_destroy.argtypes = [vp]; lib.quisk_cuda_snba_destroy.restype = None
    lib.quisk_cuda_snba_run.argtypes = [vp, vp, C.c_long, vp, C.c_long, vp]
    lib.quisk_cuda_snba_flush.argtypes = [vp]
    lib.quisk_cuda_emnr_set_tables.argtypes = [vp, vp]
    lib.quisk_cuda_emnr_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]
    lib.quisk_cuda_emnr_create.restype = vp
    lib.quisk_cuda_emnr_destroy.argtypes = [vp]; lib.quisk_cuda_emnr_destroy.restype = None
    lib.quisk_cuda_emnr_run.argtypes = [vp, vp, C.c_long, vp, C.c_long, vp]
    lib.quisk_cuda_emnr_flush.argtypes = [vp]
    for _f in ("gain_method", "npe_method", "ae_run"):
        getattr(lib, "quisk_cuda_emnr_set_" + _f).argtypes = [vp, C.c_int]
    for _f in ("run", "gain_method", "npe_method", "ae_run", "position"):
        getattr(lib, "quisk_cuda_rxa_set_emnr_" + _f).argtypes = [vp, C.c_int]
    lib.quisk_cuda_rxa_set_fm_lim_run.argtypes = [vp, C.c_int]
    lib.quisk_cuda_rxa_set_fm_lim_gain.argtypes = [vp, C.c_double]
    lib.quisk_cuda_rxa_set_mp.argtypes = [vp, C.c_int]
    lib.quisk_cuda_fircore_set_mp.argtypes = [vp, C.c_int]
    lib.quisk_cuda_rxa_set_panel_gain.argtypes = [vp, D]
    lib.quisk_cuda_rxa_in_size.argtypes = [vp]
    lib.quisk_cuda_rxa_out_size.argtypes = [vp]
    lib.quisk_cuda_rxa_xrxa.argtypes = [vp, vp, C.c_long, vp, C.c_long, vp]
    lib.quisk_cuda_rxa_fexchange0.argtypes = [vp, vp, vp, c_int_p]
    lib.quisk_cuda_rxa_xrxa_multi.argtypes = [vp, vp, C.c_long, vp, C.c_long, C.c_int, vp]
    lib.quisk_cuda_rxa_set_option.argtypes = [vp, C.c_int, C.c_int]
    lib.quisk_cuda_rxa_get_meter.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.quisk_cuda_rxa_set_channel_state.argtypes = [vp, C.c_int, C.c_int]
    lib.quisk_cuda_rxa_set_slew_down.argtypes = [vp, D, D]
    lib.quisk_cuda_rxa_set_bfo.argtypes = [vp, C.c_int]
    lib.quisk_cuda_rxa_exchange_sizes.argtypes = [vp, c_int_p, c_int_p]
    _lib = lib
    return lib


def check(lib: C.CDLL, rc: int, what: str = "") -> None:
    if rc != 0:
        raise QuiskCudaError(f"{what}: rc={rc}: {lib.quisk_cuda_last_error().decode()}")


def require_device(lib: C.CDLL | None = None) -> C.CDLL:
    lib = lib or load()
    if lib.quisk_cuda_device_count() <= 0:
        raise QuiskCudaError("no CUDA device visible: libquisk_cuda has no CPU fallback")
    return lib
