"""CPU side of the whole-_quisk swap-in (tests/test_quisk_swapin_gpu.py): the B4 sample-source plugin
(quisk_b200/plugin/quisk_block_source.c) registers through the QUISK_C_API capsule and feeds the all-reference
_quisk build; the audio it plays equals what the extracted quisk_process_decimate / quisk_process_demodulate wrapper
(oracle/_ref/libquisk_rx_ref.so, the library the chain fixtures come from) gives for the same stream up to the
stages quisk_process_samples adds around them (DC removal, AGC, volume ramp) -- checked here through the sample
counts and the read sizes, the arithmetic through the GPU test's bit-for-bit comparison of the two builds."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ref_ctypes as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = os.path.join(R.REF_DIR, "quisk_full")


@pytest.mark.parametrize("rate,mode,block", [(192000, 3, 4096), (48000, 0, 1000)])
def test_block_source_feeds_reference_quisk(rate, mode, block, tmp_path):
    if not os.path.exists(os.path.join(FULL, "ref", "_quisk.so")) or not os.path.exists(os.path.join(FULL, "quisk_block_source.so")):
        pytest.skip("oracle/_ref/quisk_full not built (oracle/build_ref.sh step 4 needs /root/reference)")
    n = 100000
    out = str(tmp_path / "o.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "quisk_swapin_driver.py"), os.path.join(FULL, "ref"), out,
                        str(rate), str(mode), "3000", str(n), str(block), "0", "100"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    z = np.load(out)
    assert z["src_status"].tolist() == [1, 0]                       # quisk_start_sound called start once, stop not yet
    assert z["reads"].sum() == n and z["reads"].max() == block     # quisk_read_sound returns what pt_sample_read delivered
    assert len(z["audio"]) == n * 48000 // rate                    # played at playback_rate 48000
    assert np.abs(z["audio"]).max() > 1e5
