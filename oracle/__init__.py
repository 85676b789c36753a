"""ctypes binding of the CPU oracle (oracle/oxi_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never by the product package (oxipng_b200/).  Parity status: "parity unpinned"
at byte level (see oxi_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboxi_oracle.so")

BASIC, MINSUM, ENTROPY, BIGRAMS, BIGENT, PREDEFINED = range(6)


class Ihdr(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("color_type", C.c_uint8),
                ("bit_depth", C.c_uint8), ("interlaced", C.c_uint8)]


class Aux(C.Structure):
    _fields_ = [("n_palette", C.c_uint32), ("palette", (C.c_uint8 * 4) * 256), ("has_trns", C.c_int32),
                ("trns", C.c_uint16 * 3)]


def build(force: bool = False) -> str:
    """Compile liboxi_oracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "oxi_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oxo_raw_data_size.restype = C.c_size_t
        _lib.oxo_bits_per_pixel.restype = C.c_size_t
        _lib.oxo_filter_bpp.restype = C.c_size_t
        _lib.oxo_scan_lines.restype = C.c_size_t
        _lib.oxo_scan_lines.argtypes = [C.POINTER(Ihdr), C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_size_t]
        _lib.oxo_ilog2i.restype = C.c_uint32
        _lib.oxo_ilog2i.argtypes = [C.c_uint32]
        _lib.oxo_paeth_predictor.restype = C.c_uint8
        _lib.oxo_paeth_predictor.argtypes = [C.c_uint8] * 3
        _lib.oxo_score_minsum.restype = C.c_uint64
        _lib.oxo_score_bigrams.restype = C.c_uint64
        _lib.oxo_score_entropy.restype = C.c_int32
        _lib.oxo_score_bigent.restype = C.c_int32
        for f in ("oxo_score_minsum", "oxo_score_bigrams", "oxo_score_entropy", "oxo_score_bigent"):
            getattr(_lib, f).argtypes = [C.c_void_p, C.c_size_t]
    return _lib


def mk_ihdr(ihdr) -> Ihdr:
    if isinstance(ihdr, Ihdr):
        return ihdr
    if isinstance(ihdr, (tuple, list)):
        w, h, ct, bd, il = ihdr
    else:
        w, h, ct, bd, il = ihdr.width, ihdr.height, int(ihdr.color_type_code), int(ihdr.bit_depth), ihdr.interlaced
    return Ihdr(int(w), int(h), int(ct), int(bd), 1 if il else 0)


def mk_aux(palette=None, trns=None) -> Aux:
    a = Aux()
    if palette is not None:
        pal = np.asarray(palette, dtype=np.uint8).reshape(-1, 4)
        a.n_palette = len(pal)
        for i, p in enumerate(pal):
            for k in range(4):
                a.palette[i][k] = int(p[k])
    if trns is not None:
        a.has_trns = 1
        t = list(trns) if isinstance(trns, (tuple, list)) else [trns] * 3
        for k in range(3):
            a.trns[k] = int(t[k])
    return a


def aux_palette(a: Aux) -> np.ndarray:
    return np.array([[a.palette[i][k] for k in range(4)] for i in range(a.n_palette)], dtype=np.uint8).reshape(-1, 4)


def aux_trns(a: Aux):
    return tuple(int(a.trns[k]) for k in range(3)) if a.has_trns else None


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a,
                                dtype=np.uint8)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def raw_data_size(ihdr) -> int:
    h = mk_ihdr(ihdr)
    return lib().oxo_raw_data_size(C.byref(h))


def filter_bpp(ihdr) -> int:
    h = mk_ihdr(ihdr)
    return lib().oxo_filter_bpp(C.byref(h))


def scan_lines(ihdr, data_len: int, has_filter: bool = False):
    """-> (len[], pass[], pixels[]) of the ScanLines iterator."""
    h = mk_ihdr(ihdr)
    n = lib().oxo_scan_lines(C.byref(h), data_len, int(has_filter), None, None, None, 0)
    ln = np.zeros(n, dtype=np.uint64)
    ps = np.zeros(n, dtype=np.uint8)
    px = np.zeros(n, dtype=np.uint64)
    lib().oxo_scan_lines(C.byref(h), data_len, int(has_filter), _p(ln), _p(ps), _p(px), n)
    return ln, ps, px


def filter_image(ihdr, data, kind: int, basic: int = 0, predefined=None, optimize_alpha: bool = False):
    """PngImage::filter_image -> (filtered stream, filters_used per row, used_is_predefined)."""
    h = mk_ihdr(ihdr)
    d = _u8(data)
    n_rows = lib().oxo_scan_lines(C.byref(h), d.size, 0, None, None, None, 0)
    out = np.empty(d.size + n_rows + 16, dtype=np.uint8)
    used = np.zeros(max(n_rows, 1), dtype=np.uint8)
    pre = _u8(predefined) if predefined is not None else np.zeros(1, dtype=np.uint8)
    npre = 0 if predefined is None else pre.size
    out_len, rows, isp = C.c_size_t(0), C.c_size_t(0), C.c_int(0)
    rc = lib().oxo_filter_image(C.byref(h), _p(d), C.c_size_t(d.size), kind, basic, _p(pre), C.c_size_t(npre),
                                int(optimize_alpha), _p(out), C.byref(out_len), _p(used), C.byref(rows), C.byref(isp))
    if rc != 0:
        raise ValueError("oxo_filter_image failed")
    return out[:out_len.value].copy(), used[:rows.value].copy(), bool(isp.value)


def unfilter_image(ihdr, filtered) -> np.ndarray:
    h = mk_ihdr(ihdr)
    d = _u8(filtered)
    out = np.empty(d.size + 16, dtype=np.uint8)
    out_len = C.c_size_t(0)
    rc = lib().oxo_unfilter_image(C.byref(h), _p(d), C.c_size_t(d.size), _p(out), C.byref(out_len))
    if rc != 0:
        raise ValueError("invalid filter byte")
    return out[:out_len.value].copy()


def interlace_image(ihdr, data) -> np.ndarray:
    h = mk_ihdr(ihdr)
    d = _u8(data)
    out = np.zeros(d.size + 64 + 8 * int(h.height), dtype=np.uint8)
    out_len = C.c_size_t(0)
    rc = lib().oxo_interlace_image(C.byref(h), _p(d), C.c_size_t(d.size), _p(out), C.byref(out_len))
    assert rc == 0
    return out[:out_len.value].copy()


def deinterlace_image(ihdr, data) -> np.ndarray:
    h = mk_ihdr(ihdr)
    d = _u8(data)
    stride = (int(h.width) * lib().oxo_bits_per_pixel(C.byref(h)) + 7) // 8
    out = np.zeros(stride * int(h.height) + 16, dtype=np.uint8)
    out_len = C.c_size_t(0)
    rc = lib().oxo_deinterlace_image(C.byref(h), _p(d), C.c_size_t(d.size), _p(out), C.byref(out_len))
    assert rc == 0
    return out[:out_len.value].copy()


def score(kind: int, line) -> int:
    d = _u8(line)
    fn = {MINSUM: "oxo_score_minsum", ENTROPY: "oxo_score_entropy", BIGRAMS: "oxo_score_bigrams",
          BIGENT: "oxo_score_bigent"}[kind]
    return int(getattr(lib(), fn)(_p(d), d.size))


# ---- reductions: each returns None (reference `None`) or a dict(ihdr=(w,h,ct,bd,il), data=..., palette=..., trns=...)

def _res(h: Ihdr, data, aux: Aux | None):
    r = {"ihdr": (h.width, h.height, h.color_type, h.bit_depth, h.interlaced), "data": data, "palette": None,
         "trns": None}
    if aux is not None:
        if h.color_type == 3:
            r["palette"] = aux_palette(aux)
        t = aux_trns(aux)
        if t is not None:
            r["trns"] = t[0] if h.color_type == 0 else t
    return r


def cleaned_alpha_channel(ihdr, data):
    h, d = mk_ihdr(ihdr), _u8(data)
    out = np.empty(d.size, dtype=np.uint8)
    rc = lib().oxo_cleaned_alpha_channel(C.byref(h), _p(d), C.c_size_t(d.size), _p(out))
    return None if rc else _res(h, out, None)


def reduced_alpha_channel(ihdr, data, optimize_alpha: bool):
    h, d = mk_ihdr(ihdr), _u8(data)
    out = np.empty(d.size + 1, dtype=np.uint8)
    ol, oh, oa = C.c_size_t(0), Ihdr(), Aux()
    rc = lib().oxo_reduced_alpha_channel(C.byref(h), _p(d), C.c_size_t(d.size), int(optimize_alpha), _p(out),
                                         C.byref(ol), C.byref(oh), C.byref(oa))
    return None if rc else _res(oh, out[:ol.value].copy(), oa)


def reduced_bit_depth_16_to_8(ihdr, data, force_scale: bool, palette=None, trns=None):
    h, d = mk_ihdr(ihdr), _u8(data)
    out = np.empty(d.size // 2 + 1, dtype=np.uint8)
    ol, oh = C.c_size_t(0), Ihdr()
    rc = lib().oxo_reduced_bit_depth_16_to_8(C.byref(h), _p(d), C.c_size_t(d.size), int(force_scale), _p(out),
                                             C.byref(ol), C.byref(oh))
    if rc:
        return None
    r = _res(oh, out[:ol.value].copy(), None)
    r["trns"] = trns  # colour type (incl. tRNS) is cloned unchanged, bit_depth.rs:27,58
    return r


def reduced_bit_depth_8_or_less(ihdr, data, palette=None, trns=None):
    h, d = mk_ihdr(ihdr), _u8(data)
    a = mk_aux(palette, trns)
    out = np.empty(d.size + 1, dtype=np.uint8)
    ol, oh, oa = C.c_size_t(0), Ihdr(), Aux()
    rc = lib().oxo_reduced_bit_depth_8_or_less(C.byref(h), C.byref(a), _p(d), C.c_size_t(d.size), _p(out),
                                               C.byref(ol), C.byref(oh), C.byref(oa))
    return None if rc else _res(oh, out[:ol.value].copy(), oa)


def expanded_bit_depth_to_8(ihdr, data, palette=None, trns=None):
    h, d = mk_ihdr(ihdr), _u8(data)
    a = mk_aux(palette, trns)
    out = np.empty(int(h.width) * int(h.height) + 16, dtype=np.uint8)
    ol, oh, oa = C.c_size_t(0), Ihdr(), Aux()
    rc = lib().oxo_expanded_bit_depth_to_8(C.byref(h), C.byref(a), _p(d), C.c_size_t(d.size), _p(out), C.byref(ol),
                                           C.byref(oh), C.byref(oa))
    return None if rc else _res(oh, out[:ol.value].copy(), oa)


def reduced_to_indexed(ihdr, data, allow_grayscale: bool, trns=None):
    h, d = mk_ihdr(ihdr), _u8(data)
    a = mk_aux(None, trns)
    out = np.empty(d.size + 1, dtype=np.uint8)
    ol, oh, oa = C.c_size_t(0), Ihdr(), Aux()
    rc = lib().oxo_reduced_to_indexed(C.byref(h), C.byref(a), _p(d), C.c_size_t(d.size), int(allow_grayscale),
                                      _p(out), C.byref(ol), C.byref(oh), C.byref(oa))
    return None if rc else _res(oh, out[:ol.value].copy(), oa)


def reduced_rgb_to_grayscale(ihdr, data, trns=None):
    h, d = mk_ihdr(ihdr), _u8(data)
    a = mk_aux(None, trns)
    out = np.empty(d.size + 1, dtype=np.uint8)
    ol, oh, oa = C.c_size_t(0), Ihdr(), Aux()
    rc = lib().oxo_reduced_rgb_to_grayscale(C.byref(h), C.byref(a), _p(d), C.c_size_t(d.size), _p(out), C.byref(ol),
                                            C.byref(oh), C.byref(oa))
    return None if rc else _res(oh, out[:ol.value].copy(), oa)


def indexed_to_channels(ihdr, data, palette, allow_grayscale: bool, optimize_alpha: bool):
    h, d = mk_ihdr(ihdr), _u8(data)
    a = mk_aux(palette, None)
    out = np.empty(4 * d.size + 4, dtype=np.uint8)
    ol, oh, oa = C.c_size_t(0), Ihdr(), Aux()
    rc = lib().oxo_indexed_to_channels(C.byref(h), C.byref(a), _p(d), C.c_size_t(d.size), int(allow_grayscale),
                                       int(optimize_alpha), _p(out), C.byref(ol), C.byref(oh), C.byref(oa))
    return None if rc else _res(oh, out[:ol.value].copy(), oa)


def reduced_palette(ihdr, data, palette, optimize_alpha: bool):
    h, d = mk_ihdr(ihdr), _u8(data)
    a = mk_aux(palette, None)
    out = np.empty(d.size + 1, dtype=np.uint8)
    oa = Aux()
    rc = lib().oxo_reduced_palette(C.byref(h), C.byref(a), _p(d), C.c_size_t(d.size), int(optimize_alpha), _p(out),
                                   C.byref(oa))
    return None if rc else _res(h, out[:d.size].copy(), oa)


def sorted_palette(ihdr, data, palette):
    h, d = mk_ihdr(ihdr), _u8(data)
    a = mk_aux(palette, None)
    out = np.empty(d.size + 1, dtype=np.uint8)
    oa = Aux()
    rc = lib().oxo_sorted_palette(C.byref(h), C.byref(a), _p(d), C.c_size_t(d.size), _p(out), C.byref(oa))
    return None if rc else _res(h, out[:d.size].copy(), oa)


def most_popular_edge_color(ihdr, num_colors: int, data):
    h, d = mk_ihdr(ihdr), _u8(data)
    r = lib().oxo_most_popular_edge_color(C.byref(h), C.c_size_t(num_colors), _p(d), C.c_size_t(d.size))
    return None if r < 0 else r


def cooccurrence_build(ihdr, num_colors: int, data) -> np.ndarray:
    h, d = mk_ihdr(ihdr), _u8(data)
    m = np.zeros((num_colors, num_colors), dtype=np.uint32)
    rc = lib().oxo_cooccurrence_build(C.byref(h), num_colors, _p(d), C.c_size_t(d.size), _p(m))
    if rc:
        raise ValueError("index beyond palette")
    return m


def apply_palette_reorder(ihdr, data, palette, remapping):
    h, d = mk_ihdr(ihdr), _u8(data)
    a = mk_aux(palette, None)
    rm = np.ascontiguousarray(remapping, dtype=np.uint32)
    out = np.empty(d.size + 1, dtype=np.uint8)
    oa = Aux()
    rc = lib().oxo_apply_palette_reorder(C.byref(h), C.byref(a), _p(rm), _p(d), C.c_size_t(d.size), _p(out),
                                         C.byref(oa))
    return None if rc else _res(h, out[:d.size].copy(), oa)


def use_native() -> bool:
    """For CPU-baseline timing only: rebuild the oracle with -march=native on this host (if gcc is present) and switch
    the binding to it.  Returns True when the native build is in use."""
    global _lib
    native = os.path.join(_HERE, "liboxi_oracle_native.so")
    try:
        src = os.path.join(_HERE, "oxi_oracle.c")
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-std=c11", "-shared", "-o", native, src, "-lm"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        return False
    saved = globals().get("_LIB_PATH")
    try:
        globals()["_LIB_PATH"] = native
        _lib = None
        lib()
        return True
    except Exception:
        globals()["_LIB_PATH"] = saved
        _lib = None
        return False
