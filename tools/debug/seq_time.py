import ctypes as C, sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quisk_b200 import lib as L
lib = L.require_device()
Cn = 256
def timeit(name, st, n, reps=20):
    x = torch.randn((Cn, n), dtype=torch.complex128, device="cuda") * 0.1
    y = torch.zeros_like(x)
    for _ in range(3): lib.quisk_cuda_seq_run(st, x.data_ptr(), n, y.data_ptr(), n, n, None)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): lib.quisk_cuda_seq_run(st, x.data_ptr(), n, y.data_ptr(), n, n, None)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print("%-8s n=%5d  %.1f us  %.1f cycles/sample" % (name, n, us, us * 1965 / n))
for n in (256, 2048, 8192):
    timeit("snotch", lib.quisk_cuda_snotch_create(Cn, 48000, 254.1, 0.0002), n)
    timeit("fmpll", lib.quisk_cuda_fmpll_create(Cn, 48000, 5000.0, -8000.0, 8000.0, 1.0, 20000.0, 0.02), n)
    timeit("fmlim", lib.quisk_cuda_wcpagc_create_fmlim(Cn, 48000, 2.5), n if n <= 2048 else 2048)
