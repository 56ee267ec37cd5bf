import os, subprocess, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
FULL = os.path.join(ROOT, "oracle", "_ref", "quisk_full")
def run(build, lib, out):
    env = dict(os.environ); env["QUISK_WDSP_LIB"] = lib
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "quisk_swapin_driver.py"), os.path.join(FULL, build), out, "48000", "3", "2000", "100000", "1000", "1", "100"], env=env, capture_output=True, text=True)
    print(r.stdout[-300:], r.stderr[-2000:])
    return np.load(out)
a = run("ref", os.path.join(ROOT, "oracle/_ref/libwdsp_ref.so"), "/tmp/a.npz")
b = run("cuda", os.path.join(ROOT, "quisk_b200/libquisk_cuda.so"), "/tmp/b.npz")
x, y = a["audio"], b["audio"]
print(len(x), len(y), np.abs(x).max(), np.abs(y).max(), np.abs(x - y).max(), x[5000:5003], y[5000:5003])
