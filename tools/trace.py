"""Debug helper: per-phase clock64() timeline of the fused decimator (not part of the product)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quisk_b200.rx import RxChain, load_tables
from quisk_b200 import lib as L
C_, n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184, 32768
tabs = load_tables(); kat = np.load("tests/golden/chain_kat.npz")
rx = RxChain(C_, 1536000, "USB", kat["c1/filt_i"], kat["c1/filt_q"], tabs, tune_hz=[12345.0] * C_, fused=True)
x = (torch.randn((C_, n), dtype=torch.float64, device="cuda") + 1j * torch.randn((C_, n), dtype=torch.float64, device="cuda")) * 1e6
a = torch.zeros((C_, rx.max_out(n)), dtype=torch.float64, device="cuda")
if os.environ.get('QC_SPLIT'): rx.set_option(11, int(os.environ['QC_SPLIT']))
for _ in range(3): rx.process(x.data_ptr(), n, n, a.data_ptr(), a.shape[1])
rx.set_option(7, 1)
rx.process(x.data_ptr(), n, n, a.data_ptr(), a.shape[1])
tr = np.zeros((C_, 16, 16), dtype=np.int64)
L.check(rx.lib, rx.lib.quisk_cuda_rx_read_trace(rx.h, tr.ctypes.data, C_), "trace")
names = ["commit", "ldg-issue", "s0", "s1", "s2", "s3", "s4", "s5", "s6", "s7"]
for c in (0, C_ // 2):
    t = tr[c]
    print("channel", c, "chunk period (cycles):", np.diff(t[:, 0])[:6])
    for ch in (2, 7):
        r = t[ch]
        seq = [r[0], r[1], r[2]] + [v for v in r[3:13] if v > 0] + [r[15]]
        print("  chunk", ch, "phase cycles:", np.diff(seq))
