"""Timing of quisk_cuda_nb_run (NoiseBlanker, quisk.c:679-784) on device-resident blocks: quiet input and input with pulses."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quisk_b200 import lib as L

lib = L.require_device()
C_, n, rate = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 32768, 1536000
rng = np.random.default_rng(7)
x = (2.0 ** 20) * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) + (2.0 ** 24) * np.exp(2j * np.pi * 0.07 * np.arange(n))
for name, pulses in (("quiet", 0), ("one pulse per 4096 samples", 8)):
    xx = x.copy()
    for k in range(pulses):
        xx[k * 4096 + 100] *= 80.0
    src = torch.from_numpy(np.stack([xx] * C_)).cuda()
    h = lib.quisk_cuda_nb_create(C_, rate)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = 0.0
    for it in range(4):             # a fresh copy of the stream every call (the blanker works in place); first call = warm-up
        d = src.clone()
        torch.cuda.synchronize()
        ev[0].record()
        lib.quisk_cuda_nb_run(h, d.data_ptr(), d.stride(0), n, 1, None)
        ev[1].record()
        torch.cuda.synchronize()
        if it:
            ms += ev[0].elapsed_time(ev[1]) / 3
    print("nb %s: %d ch x %d samples at %d S/s: %.3f ms per block = %.2f GS/s" % (name, C_, n, rate, ms, C_ * n / ms / 1e6))
    lib.quisk_cuda_nb_destroy(h)
