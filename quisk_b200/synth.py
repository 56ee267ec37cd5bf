"""quisk_b200/synth.py -- seeded synthetic test signals in WDSP's +-1.0 range (multi-tone, FM, AM) used by the
fixtures (tests/golden/make_golden_wdsp.py), the GPU tests and bench.py.  Plain NumPy; it imports nothing from oracle/."""
import numpy as np


def sig(n, seed, fs, tones=((1000.0, 0.3), (-1500.0, 0.2), (4000.0, 0.1)), noise=0.01):
    """Complex test signal in WDSP's +-1.0 range."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    x = np.zeros(n, dtype=np.complex128)
    for f, a in tones:
        x += a * np.exp(2j * np.pi * f * t + 1j * rng.uniform(0, 2 * np.pi))
    return x + noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))


def fm_sig(n, seed, fs, dev=5000.0, fmod=1000.0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    ph = dev / fmod * np.sin(2 * np.pi * fmod * t)
    return 0.5 * np.exp(1j * ph) + 0.001 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))


def am_sig(n, seed, fs, fc=300.0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    env = 0.4 * (1 + 0.5 * np.sin(2 * np.pi * 700.0 * t) + 0.3 * np.sin(2 * np.pi * 1900.0 * t))
    return env * np.exp(2j * np.pi * fc * t) + 0.002 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
