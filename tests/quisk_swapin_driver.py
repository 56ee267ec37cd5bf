"""tests/quisk_swapin_driver.py -- run the reference's WHOLE _quisk extension on a synthetic block source.

Subprocess helper of tests/test_quisk_swapin_gpu.py (TEST INFRASTRUCTURE).  argv: <build dir with _quisk.so>
<out.npz> <rate> <mode> <tune_hz> <n_samples> <block> <wdsp: 0 | 1 | 2 | 3 | 4> <dc_remove_bw>.  It does what quisk.py does at start-up, in the
same order and through the same Python methods of _quisk (record_app, set_sound_name, open_sound, set_filters,
set_rx_mode, set_tune, set_volume, start_sound), registers the B4 block source (quisk_block_source.open_samples ->
quisk_sample_source4), then calls read_sound() until the source is dry: quisk_read_sound (sound.c:873) ->
pt_sample_read -> quisk_process_samples (quisk.c:2289) -> play_sound_interface.  The "sound card" is the recording
stub of oracle/ref_wrap/sound_stub.c; get_graph() is polled after every read like the GUI timer does.
With wdsp = 1 the WDSP RXA channel is opened exactly as quisk_wdsp.py:60-90 opens it, on `QUISK_WDSP_LIB`
(libwdsp_ref.so or libquisk_cuda.so), and switched in with wdsp_set_parameter(in_use=1)."""
import ctypes
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    build, out_path, rate, mode, tune, n_samples, block, use_wdsp, dc_bw = sys.argv[1:10]
    rate, mode, tune, n_samples, block, use_wdsp, dc_bw = int(rate), int(mode), int(tune), int(n_samples), int(block), int(use_wdsp), int(dc_bw)
    sys.path.insert(0, build)
    sys.path.insert(0, os.path.dirname(build))
    import _quisk as QS
    import quisk_block_source as SRC
    from oracle import quisk_oracle as O
    from quisk_b200.rx import get_filter_center, load_tables, make_filter_coef

    # graph_refresh: get_graph (quisk.c:5276) holds a graph back until 1 / graph_refresh seconds of WALL time have passed;
    # a huge value makes every filled FFT buffer come back as its own graph, so the run is deterministic.
    # Every other key takes the default QuiskGetConfig* is given.
    conf = types.SimpleNamespace(playback_rate=48000, graph_refresh=100000000, dc_remove_bw=dc_bw)
    app = types.SimpleNamespace()
    data_width, graph_width, fft_size = 1024, 1024, 2048
    QS.record_app(app, conf, data_width, graph_width, fft_size, data_width, rate, 0, "/nonexistent/quisk_wisdom")
    print(SRC.open_samples())
    QS.set_sound_name(0, 1, 1, "recording stub", "portaudio:stub")          # radio sound -> DEV_DRIVER_PORTAUDIO (the stub)
    QS.open_sound(rate, 5000, 150, "", 0, 48000, 0, 1, 0.7, 48000)
    names = {2: "LSB", 3: "USB", 0: "CWL", 1: "CWU", 4: "AM", 5: "FM"}
    tabs = load_tables()
    QS.set_rx_mode(mode)
    bw = {0: 500, 1: 500, 4: 6000, 5: 12000}.get(mode, 2800)
    frate = QS.get_filter_rate(-1, bw)
    fi, fq = make_filter_coef(frate, None, bw, get_filter_center(names[mode], bw), tabs)
    QS.set_filters(list(map(float, fi)), list(map(float, fq)), bw, 0, 0)
    QS.set_tune(tune, tune)
    QS.set_volume(0.5)
    wl = None
    if use_wdsp:
        wl = ctypes.CDLL(os.environ["QUISK_WDSP_LIB"])
        D = ctypes.c_double
        ch, in_size, dsp_size = 1, 256, 256
        QS.wdsp_set_parameter(0, fexchange0=ctypes.cast(wl.fexchange0, ctypes.c_void_p).value)
        QS.wdsp_set_parameter(ch, in_size=in_size)
        wl.OpenChannel(ch, in_size, dsp_size, 48000, 48000, 48000, 0, 1, D(0.010), D(0.025), D(0.0), D(0.010), 1)
        wl.SetRXAShiftRun(ch, 0); wl.RXANBPSetRun(ch, 0); wl.SetRXAAMSQRun(ch, 0)
        wl.SetRXAMode(ch, 1)
        wl.RXASetPassband(ch, D(300.0), D(3000.0))
        wl.RXASetNC(ch, dsp_size); wl.RXASetMP(ch, 0)
        wl.SetRXAAGCMode(ch, 0); wl.SetRXAAGCFixed(ch, D(0.0))
        wl.SetRXAPanelRun(ch, 0); wl.SetRXAEMNRRun(ch, 0)
        if use_wdsp == 2:       # stock Quisk leaves the RXA channel a pass-through until NR2 / SNB are switched on (bp1, nbp0,
            # AGC and panel all off); switch on what this library implements, through the calls quisk.py would use
            wl.RXANBPSetRun(ch, 1)
            wl.SetRXAAGCMode(ch, 3)
            wl.SetRXAPanelRun(ch, 1)
            wl.SetRXAPanelGain1(ch, D(0.25))
        if use_wdsp == 3:       # Quisk's NR2 button (quisk.py:6017-6027): gain method 2, run, in_use = 1
            if hasattr(wl, "quisk_cuda_emnr_set_tables"):
                # the GPU library takes the two gamma-prior tables of the WDSP distribution from the host: here from the compiled reference
                ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libwdsp_ref.so"))
                T = ctypes.c_double * (241 * 241)
                gg, ggs = T.in_dll(ref, "GG"), T.in_dll(ref, "GGS")
                assert wl.quisk_cuda_emnr_set_tables(ctypes.c_void_p(ctypes.addressof(gg)), ctypes.c_void_p(ctypes.addressof(ggs))) == 0
            wl.SetRXAEMNRgainMethod(ch, 2)
            wl.SetRXAEMNRRun(ch, 1)
        if use_wdsp == 4:       # Quisk's SNB menu item (quisk.py:6040-6043): SetRXASNBARun(1), in_use = 1
            wl.SetRXASNBARun(ch, 1)
        QS.wdsp_set_parameter(ch, in_use=1)
    x = O.synth_iq(n_samples, 77, 1.0)
    if os.environ.get("QUISK_SWAPIN_ULP"):      # the reference's own conditioning: the same stream with every sample moved by at most one ulp
        v = np.ascontiguousarray(x).view(np.float64)
        step = np.random.default_rng(int(os.environ["QUISK_SWAPIN_ULP"])).integers(-1, 2, v.shape)
        x = (v + step * np.spacing(v)).view(np.complex128)
    assert SRC.load(np.ascontiguousarray(x).tobytes(), block) == n_samples
    QS.start_sound()
    graphs, reads = [], []
    while SRC.remaining() > 0:
        reads.append(QS.read_sound())
        while True:
            g = QS.get_graph(1, 1.0, 0.0)
            if g is None:
                break
            graphs.append(np.array(g, dtype=np.float64))
    lib = ctypes.CDLL(os.path.join(build, "_quisk.so"))
    lib.stub_capture_count.restype = ctypes.c_long
    lib.stub_capture_copy.restype = ctypes.c_long
    n = lib.stub_capture_count(0)
    audio = np.zeros(max(n, 1), dtype=np.complex128)
    lib.stub_capture_copy(0, audio.ctypes.data_as(ctypes.c_void_p))
    st = QS.get_state()
    np.savez(out_path, audio=audio[:n], reads=np.array(reads), graphs=np.array(graphs) if graphs else np.zeros((0, data_width)),
             src_status=np.array(SRC.status()), filter_rate=np.array([frate]), fft_error=np.array([st[14]]))
    print("swapin driver: %d reads, %d audio samples, %d graphs, filter rate %d" % (len(reads), n, len(graphs), frate))
    sys.stdout.flush()
    os._exit(0)         # skip interpreter teardown of the half-initialised GUI-less _quisk


if __name__ == "__main__":
    main()
