"""ctypes loader for liboxipng_b200.so (the C ABI in include/oxipng_b200.h).

The product path has no CPU fallback: if the library is missing or no CUDA device is usable, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# OXIPNG_B200_LIB: another build of the same library (A/B timing of kernel variants, profiles/ab.py); still the CUDA path
LIB_PATH = os.environ.get("OXIPNG_B200_LIB") or os.path.join(HERE, "liboxipng_b200.so")

OXI_OK, OXI_NONE = 0, 1
FLAG_FORCE_GENERIC, FLAG_FORCE_FAST = 1, 2


class OxiError(RuntimeError):
    pass


class CIhdr(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("color_type", C.c_uint8),
                ("bit_depth", C.c_uint8), ("interlaced", C.c_uint8)]


class CStrategy(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("basic_filter", C.c_uint8), ("predefined", C.c_void_p),
                ("n_predefined", C.c_size_t)]


class CAux(C.Structure):
    _fields_ = [("n_palette", C.c_uint32), ("palette", (C.c_uint8 * 4) * 256), ("has_trns", C.c_int32),
                ("trns", C.c_uint16 * 3)]


_lib = None


def lib():
    """Load the CUDA library; raise if it has not been built (python -m oxipng_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OxiError(f"{LIB_PATH} is missing: build it with `python -m oxipng_b200.build` "
                       "(there is no CPU fallback for this path)")
    L = C.CDLL(LIB_PATH)
    L.oxi_last_error.restype = C.c_char_p
    L.oxi_version.restype = C.c_char_p
    L.oxi_filter_tier.restype = C.c_char_p
    L.oxi_kernel_launches.restype = C.c_uint64
    for f in ("oxi_bits_per_pixel", "oxi_raw_data_size", "oxi_filter_bpp"):
        getattr(L, f).restype = C.c_size_t
        getattr(L, f).argtypes = [C.POINTER(CIhdr)]
    L.oxi_num_scanlines.restype = C.c_size_t
    L.oxi_num_scanlines.argtypes = [C.POINTER(CIhdr), C.c_size_t]
    L.oxi_scan_lines.restype = C.c_size_t
    L.oxi_scan_lines.argtypes = [C.POINTER(CIhdr), C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.oxi_filter_tier.argtypes = [C.POINTER(CIhdr), C.c_size_t, C.POINTER(CStrategy), C.c_int]
    L.oxi_filter_image.argtypes = [C.POINTER(CIhdr), C.c_void_p, C.c_size_t, C.POINTER(CStrategy), C.c_int,
                                   C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    L.oxi_filter_image_multi.argtypes = [C.POINTER(CIhdr), C.c_void_p, C.c_size_t, C.POINTER(CStrategy), C.c_size_t,
                                         C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    L.oxi_filter_batch.argtypes = [C.POINTER(CIhdr), C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(CStrategy),
                                   C.c_size_t, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_uint]
    L.oxi_filter_batch_device.argtypes = [C.c_int, C.c_void_p, C.POINTER(CIhdr), C.c_void_p, C.c_size_t, C.c_size_t,
                                          C.POINTER(CStrategy), C.c_size_t, C.c_int, C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p), C.c_uint]
    _lib = L
    return L


def check(rc: int, what: str) -> int:
    if rc < 0:
        raise OxiError(f"{what} failed ({rc}): {lib().oxi_last_error().decode()}")
    return rc
