"""quisk_b200 -- host-side Python mirror of the reference's ctypes layer (quisk_wdsp.py)
over libquisk_cuda.so, the B200 implementation of Quisk's receive-DSP hot path.

The compute lives in the shared library (quisk_b200/csrc, C ABI in include/).  This
package only loads it and wraps the batched entry points for tests and bench.py.
There is no CPU fallback: importing `quisk_b200.lib` raises if the library has not
been built, and every call fails loudly if no CUDA device is usable.
"""
from .lib import load, QuiskCudaError  # noqa: F401
