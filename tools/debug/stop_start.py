"""Debug: per-block error of the SetChannelState stop/start sequence against the fixture."""
import ctypes as C
import numpy as np
from quisk_b200 import lib as L
from quisk_b200.synth import sig
lib = L.require_device()
kat = np.load("tests/golden/wdsp_kat.npz")
x = sig(256 * 32, 730, 48000.0, tones=((-1000.0, 0.3), (-2200.0, 0.1), (1500.0, 0.2)))
rxa = lib.quisk_cuda_rxa_create(1, 256, 256, 48000, 48000, 48000)
lib.quisk_cuda_rxa_set_slew_down(rxa, 0.0, 0.010)
lib.quisk_cuda_rxa_set_slew(rxa, 0.010, 0.025)
lib.quisk_cuda_rxa_set_shift(rxa, 0, None); lib.quisk_cuda_rxa_set_nc(rxa, 2048); lib.quisk_cuda_rxa_set_mode(rxa, 1)
lib.quisk_cuda_rxa_set_passband(rxa, 150.0, 2850.0); lib.quisk_cuda_rxa_set_agc_mode(rxa, 3)
ref = kat["rxa_stop_start/y"].reshape(32, 256)
err = C.c_int(0)
for b in range(32):
    if b == 10: lib.quisk_cuda_rxa_set_channel_state(rxa, 0, 0)
    if b == 18: lib.quisk_cuda_rxa_set_channel_state(rxa, 1, 0)
    hin = np.ascontiguousarray(x[b * 256:(b + 1) * 256]); hout = np.full(256, -7 - 7j)
    lib.quisk_cuda_rxa_fexchange0(rxa, hin.ctypes.data, hout.ctypes.data, C.byref(err))
    d = np.abs(hout - ref[b])
    print(b, "max|ref| %.3e max|d| %.3e at %d  first-nonzero ours %s ref %s" % (np.abs(ref[b]).max(), d.max(), int(d.argmax()),
          np.nonzero(hout)[0][:1], np.nonzero(ref[b])[0][:1]))
