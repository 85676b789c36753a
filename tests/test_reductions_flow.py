"""End-to-end decision parity of perform_reductions (reduction/mod.rs:14-179): the host control flow driven by the GPU
reductions submits the same candidates in the same order and returns the same baseline as when it is driven by the CPU
oracle.  Also checks the outcomes the reference's own tests assert from the fixture names."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _flow(case, backend, interlacer, **opt):
    from reductions_driver import Options, perform_reductions
    img = helpers.product_image(case.ihdr, case.data, case.palette, case.trns)
    ev = helpers.RecordingEvaluator()
    base = perform_reductions(img, Options(**opt), ev, backend=backend, interlacer=interlacer)
    return base, ev.tried


def test_flow_parity_on_reference_fixtures(golden):
    import oxipng_b200 as ox
    from oxipng_b200 import reduction
    for c in golden:
        for opt in ({}, {"optimize_alpha": True}, {"scale_16": True, "grayscale_reduction": False}, {"interlace": None}):
            g_base, g_tried = _flow(c, reduction, ox.change_interlacing, **opt)
            o_base, o_tried = _flow(c, helpers.OracleReductions(), helpers.oracle_change_interlacing, **opt)
            assert g_tried == o_tried, (c.name, opt)
            assert helpers._img_tuple(g_base) == helpers._img_tuple(o_base), (c.name, opt)
            assert np.array_equal(g_base.data, o_base.data), (c.name, opt)
            assert g_base.ihdr.color_type == o_base.ihdr.color_type, (c.name, opt)


def test_fixture_names_through_the_flow(golden):
    """`<in>_should_be_<out>` where the outcome is decided by the reductions alone (no deflate-size comparison)."""
    from oxipng_b200 import reduction
    import oxipng_b200 as ox
    by = {c.name: c for c in golden}
    expect = {  # baseline colour type / depth after perform_reductions with default options
        "rgb_16_should_be_rgb_8": (2, 8), "rgba_16_should_be_grayscale_8": (0, 8), "rgb_8_should_be_grayscale_8": (0, 8),
        "rgba_8_should_be_rgb_8": (2, 8), "grayscale_16_should_be_grayscale_8": (0, 8),
        "rgba_8_should_be_grayscale_alpha_8": (4, 8),
    }
    for name, (ct, bd) in expect.items():
        base, tried = _flow(by[name], reduction, ox.change_interlacing)
        assert (base.ihdr.color_type.code, int(base.ihdr.bit_depth)) == (ct, bd), name
    # depth / palette outcomes are submitted as candidates (evaluated by size in the reference)
    base, tried = _flow(by["grayscale_8_should_be_grayscale_2"], reduction, ox.change_interlacing)
    assert any(t[1][2] == 0 and t[1][3] == 2 for t in tried)
    base, tried = _flow(by["rgba_8_should_be_palette_8"], reduction, ox.change_interlacing)
    assert any(t[1][2] == 3 for t in tried) or base.ihdr.color_type.code == 3
