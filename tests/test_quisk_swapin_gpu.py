"""SURVEY section 8 row (f)1 + boundary B4, for real: the reference's WHOLE `_quisk` extension (quisk.c, sound.c,
microphone.c, quisk_wdsp.c, ... compiled unmodified by oracle/build_ref.sh step 4; only the five audio back ends are
replaced by a recording stub and FFTW3 by the oracle's shim) built twice -- against its own filter.c, and against
quisk_b200/libquisk_cuda.so INSTEAD of filter.c -- and driven the way quisk.py drives it (tests/quisk_swapin_driver.py:
record_app, open_sound, set_filters, set_rx_mode, set_tune, start_sound, read_sound ...) from a B4 sample-source
plugin registered through the QUISK_C_API capsule with quisk_sample_source4 (quisk_b200/plugin/quisk_block_source.c).
Every filter.h call quisk_process_samples (quisk.c:2289-2741) makes on the way from the source to the sound card
lands on the GPU; the played audio must equal the all-reference build BIT FOR BIT.

With WDSP switched in (wdsp_set_parameter(in_use=1); fexchange0 pointer handed to quisk_wdsp.c exactly as
quisk_wdsp.py:57-64 does) the all-reference build runs on libwdsp_ref.so and the swap-in on libquisk_cuda.so's
OpenChannel / fexchange0 / SetRXA*: two FFT implementations, so 1e-12 relative RMS.

The panadapter (B3): the graphs the reference's own get_graph (quisk.c:5142-5331) returns while the stream runs
against quisk_cuda_pan_* on the same samples."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import quisk_oracle as O
from oracle import ref_ctypes as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = os.path.join(R.REF_DIR, "quisk_full")
DRIVER = os.path.join(ROOT, "tests", "quisk_swapin_driver.py")


@pytest.fixture(scope="module", autouse=True)
def need_gpu_and_builds():
    import torch
    assert torch.cuda.is_available()
    for f in ("ref/_quisk.so", "cuda/_quisk.so", "quisk_block_source.so"):
        if not os.path.exists(os.path.join(FULL, f)):
            pytest.skip("oracle/_ref/quisk_full/%s not built (oracle/build_ref.sh step 4 needs /root/reference)" % f)


def _run(build, tmp_path, tag, rate, mode, tune, n, block, wdsp=0, dc_bw=100, wdsp_lib=None, ulp=None):
    out = str(tmp_path / ("%s_%s.npz" % (build, tag)))
    env = dict(os.environ)
    if ulp:
        env["QUISK_SWAPIN_ULP"] = str(ulp)
    if wdsp_lib:
        env["QUISK_WDSP_LIB"] = wdsp_lib
    r = subprocess.run([sys.executable, DRIVER, os.path.join(FULL, build), out, str(rate), str(mode), str(tune), str(n), str(block), str(wdsp), str(dc_bw)],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "block player registered" in r.stdout
    return np.load(out)


# (rate, mode, tune, samples, block): USB / CWU / AM / FM; 48 k (no decimation), 192 k (the target rate), 1.536 M (C1), 250 k (6/5 + 4/5 converters)
CASES = [(192000, 3, 12345, 200000, 4096), (48000, 2, -3000, 60000, 1000), (1536000, 3, 100000, 400000, 30000),
         (250000, 1, 5000, 150000, 5000), (96000, 4, 7000, 120000, 2047), (192000, 5, -20000, 150000, 4800)]


@pytest.mark.parametrize("rate,mode,tune,n,block", CASES)
def test_whole_quisk_on_gpu_filters_equals_reference(rate, mode, tune, n, block, tmp_path):
    tag = "%d_%d" % (rate, mode)
    ref = _run("ref", tmp_path, tag, rate, mode, tune, n, block)
    gpu = _run("cuda", tmp_path, tag, rate, mode, tune, n, block)
    assert ref["src_status"][0] == 1 and gpu["src_status"][0] == 1             # quisk_start_sound called the plugin's start
    assert np.array_equal(ref["reads"], gpu["reads"])
    assert len(ref["audio"]) > 0.2 * n * 48000 / rate
    assert np.abs(ref["audio"]).max() > 1e5                                     # radio sound, not silence
    assert len(ref["audio"]) == len(gpu["audio"])
    assert np.array_equal(ref["audio"], gpu["audio"])                           # bit for bit
    assert np.array_equal(ref["graphs"], gpu["graphs"]) and len(ref["graphs"]) > 10
    assert int(ref["fft_error"][0]) == int(gpu["fft_error"][0])               # blocks above 3 x fft_size overrun get_graph's FIFO of four, in both


@pytest.mark.parametrize("wdsp", [1, 2, 3, 4])
def test_whole_quisk_with_wdsp_channel_on_gpu(wdsp, tmp_path):
    """wdsp = 1: the channel as quisk_wdsp.py opens it (every RXA stage off: a delay line through the exchange rings, so
    the two builds agree exactly); wdsp = 2: nbp0 band-pass, AGC (medium) and the panel switched on as well; wdsp = 3: what
    Quisk's NR2 button sends (SetRXAEMNRgainMethod(2), SetRXAEMNRRun(1), in_use = 1): the spectral noise reduction + bp1; wdsp = 4: Quisk's SNB menu item (SetRXASNBARun(1)):
    bpsnba, the blanker between its two resamplers, bp1.  The blanker's interpolation solves near-singular normal equations:
    on this noise-like stream the all-reference build itself moves by 2e-8 when every input sample moves by at most one ulp
    (measured here, third run), and the bound for that case is 20 x that."""
    rate, mode, tune, n, block = 48000, 3, 2000, 100000, 1000
    # The reference's exchange posts Sem_OutReady BEFORE it copies the next input block out of its two-block ring
    # (iobuffs.c:595-602): when its DSP thread is descheduled right there (a loaded host), the caller overwrites that block
    # and the reference's output is not a function of its input any more.  Two reference runs that agree rule that out.
    for attempt in range(3):
        ref = _run("ref", tmp_path, "wdsp", rate, mode, tune, n, block, wdsp=wdsp, wdsp_lib=os.path.join(R.REF_DIR, "libwdsp_ref.so"))
        again = _run("ref", tmp_path, "wdsp_again", rate, mode, tune, n, block, wdsp=wdsp, wdsp_lib=os.path.join(R.REF_DIR, "libwdsp_ref.so"))
        if np.array_equal(ref["audio"], again["audio"]):
            break
    else:
        pytest.skip("the reference's WDSP exchange raced three times in a row on this host (iobuffs.c:595-602): no stable reference output")
    gpu = _run("cuda", tmp_path, "wdsp", rate, mode, tune, n, block, wdsp=wdsp, wdsp_lib=os.path.join(ROOT, "quisk_b200", "libquisk_cuda.so"))
    plain = _run("ref", tmp_path, "nowdsp", rate, mode, tune, n, block, wdsp=0)
    assert len(ref["audio"]) == len(gpu["audio"]) and len(ref["audio"]) > 0.9 * n
    assert len(ref["audio"]) < len(plain["audio"])                              # the re-blocker holds back the partial block
    err = O.rel_rms(gpu["audio"], ref["audio"])
    print("whole _quisk + WDSP RXA channel (%d): rel rms" % wdsp, err)
    bound = 1e-12
    if wdsp == 4:
        moved = _run("ref", tmp_path, "wdsp_ulp", rate, mode, tune, n, block, wdsp=wdsp, wdsp_lib=os.path.join(R.REF_DIR, "libwdsp_ref.so"), ulp=5)
        cond = O.rel_rms(moved["audio"], ref["audio"])
        print("reference's own one-ulp sensitivity", cond)
        bound = max(bound, 20.0 * cond)
    assert err < bound
    assert O.rel_rms(ref["audio"][:len(ref["audio"])], plain["audio"][:len(ref["audio"])]) > 1e-3     # WDSP really is in the path


def test_reference_get_graph_vs_gpu_panadapter(tmp_path):
    """B3 under the real orchestrator: get_graph's pixels (fft_size 2048 -> data_width 1024, one FFT per graph) against
    quisk_cuda_pan_accumulate + quisk_cuda_pan_graph on the frames the source delivered (DC removal off)."""
    import torch
    from quisk_b200 import lib as L
    lib = L.require_device()
    rate, n, block, fft_size, width = 192000, 200000, 4096, 2048, 1024
    ref = _run("ref", tmp_path, "graph", rate, 3, 12345, n, block, dc_bw=0)
    g = ref["graphs"]
    x = O.synth_iq(n, 77, 1.0)
    nf = min(len(g), n // fft_size)
    assert nf > 90
    d = torch.from_numpy(np.ascontiguousarray(x[:nf * fft_size].reshape(nf, fft_size))).cuda()      # every frame = one "channel"
    pan = lib.quisk_cuda_pan_create(nf, fft_size)
    assert pan, lib.quisk_cuda_last_error()
    out = torch.zeros((nf, width), dtype=torch.float64, device="cuda")
    assert lib.quisk_cuda_pan_accumulate(pan, d.data_ptr(), fft_size, 1, None) == 0
    assert lib.quisk_cuda_pan_graph(pan, width, 1.0, 0.0, float(rate), out.data_ptr(), None) == 0
    torch.cuda.synchronize()
    y = out.cpu().numpy()
    lib.quisk_cuda_pan_destroy(pan)
    err = np.abs(y - g[:nf]).max()
    print("get_graph vs quisk_cuda_pan: max |dB| difference", err, "over", nf, "graphs")
    assert err < 1e-9
