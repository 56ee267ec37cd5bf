"""Shared helpers for the test-suite."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

SPLITS = [1, 2, 3, 7, 255, 256, 257, 1000, 1, 2218]
CHAIN_SPLITS = [1, 2, 3, 7, 255, 256, 257, 1000, 4093, 15360, 18766]
DEMOD_SPLITS = [1, 2, 3, 7, 255, 256, 257, 1000, 4093, 6126]


def golden(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):        # quisk_tables.npz is package data: the product reads it too
        path = os.path.join(ROOT, "quisk_b200", "data", name)
    z = np.load(path)
    return {k: z[k] for k in z.files}


def declared_functions(header_path):
    """Function names declared in a C header (prototypes ending in ');')."""
    src = open(header_path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"^\s*#[^\n]*$", "", src, flags=re.M)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src):
        n = m.group(1)
        if n not in ("defined", "sizeof"):
            names.add(n)
    return sorted(names)


def dev_tensor(arr):
    """numpy (complex128 / float64) -> torch CUDA tensor sharing the layout."""
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()
