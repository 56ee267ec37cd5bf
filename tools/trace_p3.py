"""Debug helper: per-group clock64() timeline of fused_decim_p3_kernel (not part of the product)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quisk_b200.rx import RxChain, load_tables
from quisk_b200 import lib as L
C_, n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184, 32768
tabs = load_tables(); kat = np.load("tests/golden/chain_kat.npz")
rx = RxChain(C_, 1536000, "USB", kat["c1/filt_i"], kat["c1/filt_q"], tabs, tune_hz=[12345.0] * C_, fused=True)
rx.set_option(20, 1)
x = (torch.randn((C_, n), dtype=torch.float64, device="cuda") + 1j * torch.randn((C_, n), dtype=torch.float64, device="cuda")) * 1e6
a = torch.zeros((C_, rx.max_out(n)), dtype=torch.float64, device="cuda")
for _ in range(3): rx.process(x.data_ptr(), n, n, a.data_ptr(), a.shape[1])
rx.set_option(7, 1)
rx.process(x.data_ptr(), n, n, a.data_ptr(), a.shape[1])
tr = np.zeros((C_, 16, 16), dtype=np.int64)
L.check(rx.lib, rx.lib.quisk_cuda_rx_read_trace(rx.h, tr.ctypes.data, C_), "trace")
print(L.kernel_name(rx) if hasattr(L, "kernel_name") else "")
for c in (0, C_ // 2):
    t = tr[c]; t0 = t[0, 0]
    print("channel", c)
    for ch in range(0, 10):
        r = t[ch] - t0
        print("  chunk %2d A: top %6d commit+ldg %5d wait %5d hb0 %5d | B: top %6d wait %5d hb1 %5d wait %5d hb2+slide %5d | C: top %6d wait %5d casc %5d slide %5d" % (
            ch, r[0], r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4], r[5] - r[4], r[6] - r[5], r[7] - r[6], r[12] - r[7], r[8], r[9] - r[8], r[10] - r[9], r[11] - r[10]))
