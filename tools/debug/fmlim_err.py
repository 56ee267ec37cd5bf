import ctypes as C, sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from quisk_b200 import lib as L
from quisk_b200.synth import fm_sig
from tests.golden.make_golden_wdsp_fmlim import BLOCKS, GAIN_AT, N, TAIL
lib = L.require_device()
kat = np.load(os.path.join(ROOT, "tests/golden/wdsp_fmlim_kat.npz"))
def run(x, lim=1):
    rxa = lib.quisk_cuda_rxa_create(1, N, N, 48000, 48000, 48000)
    lib.quisk_cuda_rxa_set_shift(rxa, 0, None); lib.quisk_cuda_rxa_set_mode(rxa, 5); lib.quisk_cuda_rxa_set_passband(rxa, -8000.0, 8000.0)
    lib.quisk_cuda_rxa_set_fm_lim_run(rxa, lim)
    d = torch.from_numpy(x[None, :].copy()).cuda(); o = torch.zeros_like(d)
    for b in range(GAIN_AT):
        lib.quisk_cuda_rxa_xrxa(rxa, d[:, b * N:(b + 1) * N].data_ptr(), d.stride(0), o[:, b * N:(b + 1) * N].data_ptr(), o.stride(0), None)
    torch.cuda.synchronize()
    lib.quisk_cuda_rxa_destroy(rxa)
    return o.cpu().numpy()[0]
x = fm_sig(N * BLOCKS, 810, 48000.0); x[0] = 0
y = run(x)
ours = y[(GAIN_AT - TAIL - 2) * N:(GAIN_AT - 2) * N]
ref = kat["y_seg"][:TAIL * N]
e = np.abs(ours - ref)
print("rel rms", np.sqrt(np.mean(e**2)) / np.sqrt(np.mean(np.abs(ref)**2)), "max", e.max(), "at", e.argmax(), "ref peak", np.abs(ref).max())
for b in range(0, TAIL, 3):
    eb = e[b * N:(b + 1) * N]
    print(b, "block rel", np.sqrt(np.mean(eb**2)) / np.sqrt(np.mean(np.abs(ref[b*N:(b+1)*N])**2)))
v = np.ascontiguousarray(x).view(np.float64).copy()
up = np.random.default_rng(9).integers(0, 2, size=v.shape).astype(bool)
xp = np.where(up, np.nextafter(v, np.inf), np.nextafter(v, -np.inf)).view(np.complex128)
yp = run(xp)
s = slice((GAIN_AT - TAIL - 2) * N, (GAIN_AT - 2) * N)
print("our own one-ulp sensitivity", np.sqrt(np.mean(np.abs(yp[s] - y[s])**2)) / np.sqrt(np.mean(np.abs(y[s])**2)))
for lim in (0, 1):
    a = run(x, lim); b = run(xp, lim)
    print("lim", lim, "sensitivity per 10 blocks:", [float("%.2g" % (np.sqrt(np.mean(np.abs(b[k*N:(k+10)*N] - a[k*N:(k+10)*N])**2)) / (np.sqrt(np.mean(np.abs(a[k*N:(k+10)*N])**2)) + 1e-300))) for k in range(0, GAIN_AT - 10, 10)])
