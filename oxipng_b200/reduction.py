"""Host-side mirror of the reference's reduction entry points (src/reduction/{alpha,bit_depth,color,palette}.rs,
re-exported through `internal_tests`, src/lib.rs:58-64).  Same names, arguments and `Option<PngImage>` convention
(None = not applicable / would not reduce).  All pixel scans run in CUDA through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ._lib import CAux, CIhdr, OXI_NONE, check, lib
from .png import BitDepth, ColorType, IhdrData, PngImage

INDEXED_MAX_DIFF = 20000  # color.rs:16


def _aux_of(ct: ColorType) -> CAux:
    a = CAux()
    if ct.code == 3:
        pal = ct.palette_array()
        a.n_palette = len(pal)
        if len(pal):
            C.memmove(a.palette, np.ascontiguousarray(pal).ctypes.data, len(pal) * 4)
    elif ct.transparent is not None:
        a.has_trns = 1
        t = [ct.transparent] * 3 if ct.code == 0 else list(ct.transparent)
        for k in range(3):
            a.trns[k] = int(t[k])
    return a


def _color_type(code: int, aux: Optional[CAux]) -> ColorType:
    if code == 3:
        n = aux.n_palette if aux is not None else 0
        pal = np.ctypeslib.as_array(aux.palette).reshape(-1, 4)[:n].copy() if n else np.zeros((0, 4), np.uint8)
        return ColorType.Indexed(pal)
    if code == 0:
        return ColorType.Grayscale(int(aux.trns[0]) if aux is not None and aux.has_trns else None)
    if code == 2:
        return ColorType.RGB(tuple(int(aux.trns[k]) for k in range(3)) if aux is not None and aux.has_trns else None)
    return ColorType(code)


def _image(oh: CIhdr, oa: Optional[CAux], data: np.ndarray) -> PngImage:
    return PngImage(IhdrData(oh.width, oh.height, _color_type(oh.color_type, oa), BitDepth(oh.bit_depth), bool(oh.interlaced)),
                    data)


def _p(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def cleaned_alpha_channel(png: PngImage) -> Optional[PngImage]:  # alpha.rs:11
    h = png.ihdr._c()
    out = np.empty(png.data.size, np.uint8)
    rc = check(lib().oxi_cleaned_alpha_channel(C.byref(h), _p(png.data), C.c_size_t(png.data.size), _p(out)),
               "oxi_cleaned_alpha_channel")
    if rc == OXI_NONE:
        return None
    bpp = png.channels_per_pixel() * png.bytes_per_channel()
    return PngImage(png.ihdr, out[:png.data.size // bpp * bpp])


def reduced_alpha_channel(png: PngImage, optimize_alpha: bool) -> Optional[PngImage]:  # alpha.rs:35
    h = png.ihdr._c()
    out = np.empty(png.data.size + 16, np.uint8)
    ol, oh, oa = C.c_size_t(0), CIhdr(), CAux()
    rc = check(lib().oxi_reduced_alpha_channel(C.byref(h), _p(png.data), C.c_size_t(png.data.size), int(optimize_alpha),
                                               _p(out), C.byref(ol), C.byref(oh), C.byref(oa)), "oxi_reduced_alpha_channel")
    return None if rc == OXI_NONE else _image(oh, oa, out[:ol.value])


def reduced_bit_depth_16_to_8(png: PngImage, force_scale: bool) -> Optional[PngImage]:  # bit_depth.rs:9
    h = png.ihdr._c()
    out = np.empty(png.data.size // 2 + 16, np.uint8)
    ol, oh = C.c_size_t(0), CIhdr()
    rc = check(lib().oxi_reduced_bit_depth_16_to_8(C.byref(h), _p(png.data), C.c_size_t(png.data.size), int(force_scale),
                                                   _p(out), C.byref(ol), C.byref(oh)), "oxi_reduced_bit_depth_16_to_8")
    if rc == OXI_NONE:
        return None
    return PngImage(IhdrData(oh.width, oh.height, png.ihdr.color_type, BitDepth.Eight, bool(oh.interlaced)), out[:ol.value])


def scaled_bit_depth_16_to_8(png: PngImage) -> Optional[PngImage]:  # bit_depth.rs:36
    return reduced_bit_depth_16_to_8(png, True)


def _with_aux(fn_name, png: PngImage, out_cap: int, *extra):
    h, a = png.ihdr._c(), _aux_of(png.ihdr.color_type)
    out = np.empty(out_cap + 16, np.uint8)
    ol, oh, oa = C.c_size_t(0), CIhdr(), CAux()
    fn = getattr(lib(), fn_name)
    rc = check(fn(C.byref(h), C.byref(a), _p(png.data), C.c_size_t(png.data.size), *extra, _p(out), C.byref(ol),
                  C.byref(oh), C.byref(oa)), fn_name)
    return None if rc == OXI_NONE else _image(oh, oa, out[:ol.value])


def reduced_bit_depth_8_or_less(png: PngImage) -> Optional[PngImage]:  # bit_depth.rs:68
    return _with_aux("oxi_reduced_bit_depth_8_or_less", png, png.data.size)


def expanded_bit_depth_to_8(png: PngImage) -> Optional[PngImage]:  # bit_depth.rs:170
    # one output byte per pixel; sized from the data (<= 8 pixels per input byte), not from the IHDR: a non-interlaced
    # image keeps yielding rows while bytes remain (scan_lines.rs), so the data may hold more rows than ihdr.height
    return _with_aux("oxi_expanded_bit_depth_to_8", png, max(png.ihdr.width * png.ihdr.height, png.data.size * 8))


def reduced_to_indexed(png: PngImage, allow_grayscale: bool) -> Optional[PngImage]:  # color.rs:38
    return _with_aux("oxi_reduced_to_indexed", png, png.data.size, C.c_int(int(allow_grayscale)))


def reduced_rgb_to_grayscale(png: PngImage) -> Optional[PngImage]:  # color.rs:100
    return _with_aux("oxi_reduced_rgb_to_grayscale", png, png.data.size)


def indexed_to_channels(png: PngImage, allow_grayscale: bool, optimize_alpha: bool) -> Optional[PngImage]:
    """color.rs:141-206.  Capped by the reference to <= 20 000 bytes of growth (:184-187), i.e. a few thousand pixels:
    a 256-entry palette gather that stays on the host in this design."""
    if png.ihdr.bit_depth != BitDepth.Eight or png.ihdr.color_type.code != 3:
        return None
    pal = png.ihdr.color_type.palette_array().copy()
    if optimize_alpha:
        pal[pal[:, 3] == 0, :3] = 0
    is_gray = bool(allow_grayscale and np.all((pal[:, 0] == pal[:, 1]) & (pal[:, 1] == pal[:, 2])))
    has_alpha = bool(np.any(pal[:, 3] != 255))
    ct = {(False, True): ColorType.RGBA(), (False, False): ColorType.RGB(), (True, True): ColorType.GrayscaleAlpha(),
          (True, False): ColorType.Grayscale()}[(is_gray, has_alpha)]
    out_size = ct.channels_per_pixel() * png.data.size
    if out_size - png.data.size > INDEXED_MAX_DIFF:
        return None
    full = np.zeros((256, 4), np.uint8)
    full[:, 3] = 255
    full[:len(pal)] = pal
    cols = full[png.data][:, (2 if is_gray else 0):(4 if has_alpha else 3)]
    return PngImage(IhdrData(png.ihdr.width, png.ihdr.height, ct, png.ihdr.bit_depth, png.ihdr.interlaced),
                    np.ascontiguousarray(cols).reshape(-1))


def _palette_call(fn_name, png: PngImage, *extra) -> Optional[PngImage]:
    h, a = png.ihdr._c(), _aux_of(png.ihdr.color_type)
    out = np.empty(png.data.size + 16, np.uint8)
    oa = CAux()
    fn = getattr(lib(), fn_name)
    rc = check(fn(C.byref(h), C.byref(a), *extra[:1], _p(png.data), C.c_size_t(png.data.size), *extra[1:], _p(out),
                  C.byref(oa)), fn_name)
    if rc == OXI_NONE:
        return None
    return PngImage(IhdrData(png.ihdr.width, png.ihdr.height, _color_type(3, oa), png.ihdr.bit_depth, png.ihdr.interlaced),
                    out[:png.data.size])


def reduced_palette(png: PngImage, optimize_alpha: bool) -> Optional[PngImage]:  # palette.rs:12
    h, a = png.ihdr._c(), _aux_of(png.ihdr.color_type)
    out = np.empty(png.data.size + 16, np.uint8)
    oa = CAux()
    rc = check(lib().oxi_reduced_palette(C.byref(h), C.byref(a), _p(png.data), C.c_size_t(png.data.size),
                                         int(optimize_alpha), _p(out), C.byref(oa)), "oxi_reduced_palette")
    if rc == OXI_NONE:
        return None
    return PngImage(IhdrData(png.ihdr.width, png.ihdr.height, _color_type(3, oa), png.ihdr.bit_depth, png.ihdr.interlaced),
                    out[:png.data.size])


def sorted_palette(png: PngImage) -> Optional[PngImage]:  # palette.rs:77
    h, a = png.ihdr._c(), _aux_of(png.ihdr.color_type)
    out = np.empty(png.data.size + 16, np.uint8)
    oa = CAux()
    rc = check(lib().oxi_sorted_palette(C.byref(h), C.byref(a), _p(png.data), C.c_size_t(png.data.size), _p(out),
                                        C.byref(oa)), "oxi_sorted_palette")
    if rc == OXI_NONE:
        return None
    return PngImage(IhdrData(png.ihdr.width, png.ihdr.height, _color_type(3, oa), png.ihdr.bit_depth, png.ihdr.interlaced),
                    out[:png.data.size])


def apply_palette_reorder(png: PngImage, remapping) -> Optional[PngImage]:  # palette.rs:165
    h, a = png.ihdr._c(), _aux_of(png.ihdr.color_type)
    rm = np.ascontiguousarray(remapping, dtype=np.uint32)
    out = np.empty(png.data.size + 16, np.uint8)
    oa = CAux()
    rc = check(lib().oxi_apply_palette_reorder(C.byref(h), C.byref(a), _p(rm), _p(png.data), C.c_size_t(png.data.size),
                                               _p(out), C.byref(oa)), "oxi_apply_palette_reorder")
    if rc == OXI_NONE:
        return None
    return PngImage(IhdrData(png.ihdr.width, png.ihdr.height, _color_type(3, oa), png.ihdr.bit_depth, png.ihdr.interlaced),
                    out[:png.data.size])


def most_popular_edge_color(num_colors: int, png: PngImage) -> Optional[int]:  # palette.rs:197
    h = png.ihdr._c()
    counts = np.zeros(256, np.uint32)
    check(lib().oxi_edge_color_counts(C.byref(h), _p(png.data), C.c_size_t(png.data.size), _p(counts)),
          "oxi_edge_color_counts")
    n = min(num_colors, 256)
    if n == 0:
        return None
    best = 0
    for i in range(n):  # max_by_key keeps the last maximum
        if counts[i] >= counts[best]:
            best = i
    return None if int((counts == counts[best]).sum()) > 1 else best


def used_histogram256(data) -> np.ndarray:  # palette.rs:20-23
    d = np.ascontiguousarray(data, dtype=np.uint8)
    hist = np.zeros(256, np.uint64)
    check(lib().oxi_used_histogram256(_p(d), C.c_size_t(d.size), _p(hist)), "oxi_used_histogram256")
    return hist


def lut_remap(data, lut) -> np.ndarray:
    d = np.ascontiguousarray(data, dtype=np.uint8)
    l = np.ascontiguousarray(lut, dtype=np.uint8)
    out = np.empty(d.size, np.uint8)
    check(lib().oxi_lut_remap(_p(d), C.c_size_t(d.size), _p(l), _p(out)), "oxi_lut_remap")
    return out


class CoOccurrenceMatrix:
    """palette.rs:527-553 (`from`) and :556-589 (`build`): the pixel scan runs on the GPU; the <=256-node reindexing
    heuristics that consume the matrix stay with the reference's host code."""

    def __init__(self, num_colors: int, data: np.ndarray):
        self.num_colors = num_colors
        self.data = data

    @staticmethod
    def from_image(png: PngImage) -> Optional["CoOccurrenceMatrix"]:
        if png.ihdr.bit_depth != BitDepth.Eight or png.ihdr.color_type.code != 3:
            return None
        n = len(png.ihdr.color_type.palette_array())
        if n <= 2:
            return None
        h = png.ihdr._c()
        m = np.zeros((n, n), np.uint32)
        check(lib().oxi_cooccurrence_build(C.byref(h), C.c_uint32(n), _p(png.data), C.c_size_t(png.data.size), _p(m)),
              "oxi_cooccurrence_build")
        return CoOccurrenceMatrix(n, m)
