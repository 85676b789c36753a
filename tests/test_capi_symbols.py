"""The C-ABI library loads on a CPU-only box and exports every symbol include/oxipng_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oxipng_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "oxipng_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oxi_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_version_and_launch_counter():
    L = _lib.lib()
    assert b"sm_100a" in L.oxi_version()
    assert L.oxi_kernel_launches() >= 0


def test_compute_fails_loudly_without_gpu():
    """No CPU fallback: without a CUDA device the filter entry point reports an error instead of computing."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import oxipng_b200 as ox
    img = ox.PngImage(ox.IhdrData(4, 4, ox.ColorType.RGBA(), ox.BitDepth.Eight), np.zeros(64, np.uint8))
    with pytest.raises(ox.OxiError):
        img.filter_image(ox.FilterStrategy.MinSum, False)


def test_argument_validation():
    L = _lib.lib()
    bad = _lib.CIhdr(0, 4, 6, 8, 0)
    assert L.oxi_num_scanlines(C.byref(bad), 64) == 0
    bad = _lib.CIhdr(4, 4, 3, 16, 0)  # indexed cannot be 16-bit
    assert L.oxi_num_scanlines(C.byref(bad), 64) == 0
