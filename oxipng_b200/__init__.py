"""oxipng_b200 -- B200 (sm_100a) implementation of oxipng's filter-search and reduction-scan hot path.

Public surface mirrors the reference's `internal_tests` re-exports for this path (src/lib.rs:58-64):
PngImage.filter_image, RowFilter, FilterStrategy, and the functions in `reduction`.
"""
from ._lib import OxiError, FLAG_FORCE_FAST, FLAG_FORCE_GENERIC, lib  # noqa: F401
from .png import (BitDepth, ColorType, FilterStrategy, IhdrData, PngImage, RowFilter, filter_batch,  # noqa: F401
                  filter_batch_device, filter_image_multi, filter_tier, unfilter_image, unfilter_image_device, unfilter_batch, unfilter_batch_device,
                  interlace_image,
                  deinterlace_image, change_interlacing)

from . import reduction  # noqa: F401,E402

__all__ = ["reduction", "BitDepth", "ColorType", "FilterStrategy", "IhdrData", "PngImage", "RowFilter", "filter_batch",
           "filter_batch_device", "filter_image_multi", "filter_tier", "unfilter_image", "unfilter_image_device", "unfilter_batch",
           "unfilter_batch_device", "interlace_image",
           "deinterlace_image", "change_interlacing", "OxiError", "lib"]
