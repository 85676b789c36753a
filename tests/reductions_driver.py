"""TEST / INTEGRATION EXAMPLE, not product code (the control flow stays in the reference's Rust).
Host-side mirror of `perform_reductions` (src/reduction/mod.rs:14-179): the ORDER and GATING of the reductions and
which results go to the evaluator versus become the new baseline.  The control flow stays on the host exactly as in
the reference; every pixel scan it calls runs on the GPU (oxipng_b200.reduction), every trial goes through the fused
evaluator call (tests/evaluate_mirror.py).

`backend` is the module providing the reduction functions (default: oxipng_b200.reduction).  Tests pass an
oracle-backed adapter with the same names to check end-to-end decision parity of the whole flow.

Not mirrored: `sorted_palette_ezeng` (the Zeng-style reindexers, palette.rs:240-523) -- <=256-node host heuristics that
the reference only runs when `cheap` is false; requesting that step raises.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from oxipng_b200.png import PngImage, change_interlacing

INDEXED_MAX_DIFF = 20000  # color.rs:16


@dataclass
class Options:
    """The fields of `Options` (src/options.rs:88-162) that perform_reductions reads; defaults = `Options::default()`
    (-o2, options.rs:276-302)."""
    interlace: Optional[bool] = False        # Some(false): remove interlacing
    optimize_alpha: bool = False
    bit_depth_reduction: bool = True
    scale_16: bool = False
    color_type_reduction: bool = True
    grayscale_reduction: bool = True
    palette_reduction: bool = True
    deflater_compression: int = 11           # Deflater::Libdeflater { compression: 11 }
    fast_evaluation: bool = True


def perform_reductions(png: PngImage, opts: Options, evaluator, backend=None, interlacer=change_interlacing) -> PngImage:
    """-> the baseline image (reduction/mod.rs:178); candidates that need a size comparison are pushed into `evaluator`
    with try_image, in the reference's order."""
    if backend is None:
        from oxipng_b200 import reduction as backend
    R = backend
    evaluation_added = False
    cheap = opts.deflater_compression < 12 and opts.fast_evaluation  # :24-27

    if opts.interlace is not None:  # :30-34
        reduced = interlacer(png, opts.interlace)
        if reduced is not None:
            png = reduced
    if opts.optimize_alpha:  # :38-42
        reduced = R.cleaned_alpha_channel(png)
        if reduced is not None:
            png = reduced
    if opts.bit_depth_reduction:  # :46-50
        reduced = R.reduced_bit_depth_16_to_8(png, opts.scale_16)
        if reduced is not None:
            png = reduced
    if opts.color_type_reduction and opts.grayscale_reduction:  # :54-58
        reduced = R.reduced_rgb_to_grayscale(png)
        if reduced is not None:
            png = reduced
    if opts.bit_depth_reduction:  # :62-66
        reduced = R.expanded_bit_depth_to_8(png)
        if reduced is not None:
            png = reduced

    baseline = png  # :70

    if opts.palette_reduction:  # :73-89
        reduced = R.reduced_palette(png, opts.optimize_alpha)
        if reduced is not None:
            png = reduced
            if png.data.size == baseline.data.size and (png.data == baseline.data).all():
                baseline = png
        reduced = R.sorted_palette(png)
        if reduced is not None:
            png = reduced
        if png is not baseline:
            evaluator.try_image(png, "Indexed (luma sort)")
            evaluation_added = True

    if opts.color_type_reduction:  # :92-104
        reduced = R.reduced_alpha_channel(png, opts.optimize_alpha)
        if reduced is not None:
            png = reduced
            if png.ihdr.color_type.has_trns() and baseline.data.size - png.data.size <= 1000:
                evaluator.try_image(png)
                evaluation_added = True
            else:
                baseline = png

    if not cheap and opts.color_type_reduction:  # :108-116
        reduced = R.indexed_to_channels(png, opts.grayscale_reduction, opts.optimize_alpha)
        if reduced is not None:
            evaluator.try_image(reduced)
            evaluation_added = True

    indexed = None
    if opts.color_type_reduction and opts.palette_reduction:  # :120-135
        reduced = R.reduced_to_indexed(png, opts.grayscale_reduction)
        if reduced is not None:
            srt = R.sorted_palette(reduced)
            new = srt if srt is not None else reduced
            if png.data.size - new.data.size <= INDEXED_MAX_DIFF:
                evaluator.try_image(new, "Indexed (luma sort)")
                evaluation_added = True
            else:
                baseline = new
            indexed = new

    if not cheap and opts.palette_reduction:  # :138-152
        raise NotImplementedError("sorted_palette_ezeng stays on the reference host code (palette.rs:240-523)")

    if opts.bit_depth_reduction:  # :155-173
        reduced = R.reduced_bit_depth_8_or_less(png)
        if (not cheap or reduced is None) and indexed is not None:
            ind = R.reduced_bit_depth_8_or_less(indexed)
            if ind is not None:
                if reduced is None or reduced.data.size != ind.data.size or (reduced.data != ind.data).any():
                    evaluator.try_image(ind)
                    evaluation_added = True
        if reduced is not None:
            evaluator.try_image(reduced)
            evaluation_added = True

    if evaluation_added:  # :175-177
        evaluator.try_image(baseline)
    return baseline
