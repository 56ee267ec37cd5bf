"""quisk_b200/build.py -- compiles libquisk_cuda.so in-tree with nvcc for sm_100a.

Usage: python -m quisk_b200.build [--force]
The shared library lands in quisk_b200/libquisk_cuda.so (git-ignored; it travels
to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libquisk_cuda.so")
OBJ = os.path.join(HERE, "csrc", "_obj")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--fmad=true"]
CXX_FLAGS = ["-O2", "-fPIC", "-std=gnu++17"]


def _sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(f for f in os.listdir(CSRC) if f.endswith(".cpp"))
    return cu, cpp


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    cu, cpp = _sources()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "quisk_cuda.h"))
    for extra in ("quisk_cuda_wdsp.h",):
        p = os.path.join(HERE, "..", "include", extra)
        if os.path.exists(p):
            headers.append(p)
    objs = []
    procs = []
    for f in cu:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            log = open(obj + ".log", "w")
            flags = list(NVCC_FLAGS)
            if f.endswith("_nofma.cu"):      # recurrent WDSP stages: a*b+c must round twice, as gcc's x86-64 code does
                flags[flags.index("--fmad=true")] = "--fmad=false"
            procs.append((f, log, subprocess.Popen([NVCC] + flags + ["-c", src, "-o", obj], stdout=log, stderr=subprocess.STDOUT)))
    for f in cpp:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f[:-4] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            log = open(obj + ".log", "w")
            procs.append((f, log, subprocess.Popen(["g++"] + CXX_FLAGS + ["-c", src, "-o", obj], stdout=log, stderr=subprocess.STDOUT)))
    failed = False
    for f, log, p in procs:
        rc = p.wait()
        log.close()
        if rc != 0 or verbose:
            sys.stderr.write(open(log.name).read())
        if rc != 0:
            failed = True
            sys.stderr.write(f"build: {f} failed\n")
    if failed:
        raise RuntimeError("libquisk_cuda build failed")
    if force or procs or _newer(OUT, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + \
              ["-cudart", "shared", "-lquadmath", "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
