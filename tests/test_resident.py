"""Device-resident images (oxi_image_*): every handle -> handle reduction equals the host-buffer
entry point (itself checked against the oracle in test_reduction_parity.py), the resident reduction flow takes the same
decisions as the host-buffer flow, and the resident filter search returns the oracle's streams."""
import numpy as np
import pytest

import helpers
import oracle

pytestmark = pytest.mark.gpu


def _same(a, b, what):
    assert (a is None) == (b is None), what
    if a is None:
        return
    ha = a.download() if hasattr(a, "download") else a
    assert helpers._img_tuple(ha) == helpers._img_tuple(b), what
    assert ha.ihdr.color_type == b.ihdr.color_type, what
    assert np.array_equal(ha.data, b.data), what


def test_resident_reductions_equal_host_buffer_entry_points(golden):
    from oxipng_b200 import reduction as R
    from oxipng_b200.resident import DeviceImage
    import oxipng_b200 as ox
    for c in golden:
        img = helpers.product_image(c.ihdr, c.data, c.palette, c.trns)
        dev = DeviceImage.upload(img)
        _same(dev, img, (c.name, "roundtrip"))
        _same(dev.cleaned_alpha_channel(), R.cleaned_alpha_channel(img), (c.name, "clean"))
        for oa in (False, True):
            _same(dev.reduced_alpha_channel(oa), R.reduced_alpha_channel(img, oa), (c.name, "alpha", oa))
            _same(dev.reduced_palette(oa), R.reduced_palette(img, oa), (c.name, "palette", oa))
        for fs in (False, True):
            _same(dev.reduced_bit_depth_16_to_8(fs), R.reduced_bit_depth_16_to_8(img, fs), (c.name, "16to8", fs))
        _same(dev.reduced_bit_depth_8_or_less(), R.reduced_bit_depth_8_or_less(img), (c.name, "8orless"))
        _same(dev.expanded_bit_depth_to_8(), R.expanded_bit_depth_to_8(img), (c.name, "expand"))
        for ag in (False, True):
            _same(dev.reduced_to_indexed(ag), R.reduced_to_indexed(img, ag), (c.name, "indexed", ag))
        _same(dev.reduced_rgb_to_grayscale(), R.reduced_rgb_to_grayscale(img), (c.name, "gray"))
        _same(dev.sorted_palette(), R.sorted_palette(img), (c.name, "sorted"))
        if not img.ihdr.interlaced:
            il = dev.interlace_image()
            _same(il, ox.interlace_image(img), (c.name, "interlace"))
            _same(il.deinterlace_image(), img, (c.name, "deinterlace"))
        if img.ihdr.color_type.code == 3 and int(img.ihdr.bit_depth) == 8:
            n = len(img.ihdr.color_type.palette_array())
            if n > 2 and int(c.data.max()) < n:
                m = R.CoOccurrenceMatrix.from_image(img)
                assert np.array_equal(dev.cooccurrence_matrix(n), m.data), c.name
            rm = np.arange(n, dtype=np.uint32)[::-1].copy()
            if n > 1:
                _same(dev.apply_palette_reorder(rm), R.apply_palette_reorder(img, rm), (c.name, "reorder"))
            assert np.array_equal(dev.used_histogram256(), R.used_histogram256(img.data)), c.name


def test_resident_chain_and_filter(golden):
    """upload once -> reductions chained on the device -> filter search: the streams equal the oracle's on the reduced image."""
    import oxipng_b200 as ox
    from oxipng_b200.resident import DeviceImage
    by = {c.name: c for c in golden}
    c = by["rgba_16_should_be_grayscale_8"]
    img = helpers.product_image(c.ihdr, c.data, c.palette, c.trns)
    dev = DeviceImage.upload(img)
    d8 = dev.reduced_bit_depth_16_to_8(False)
    g = d8.reduced_rgb_to_grayscale()
    ga = g.reduced_alpha_channel(False)
    final = ga if ga is not None else g
    host = final.download()
    ih = helpers._img_tuple(host)
    S = ox.FilterStrategy
    names = ["None", "Sub", "MinSum", "Entropy", "Bigrams", "BigEnt"]
    outs, used = final.filter_image_multi([helpers.product_strategy(n) for n in names], False)
    kinds = {n: (k, b) for n, k, b in helpers.STRATEGIES}
    for n, o, u in zip(names, outs, used):
        want, want_used, _ = oracle.filter_image(ih, host.data, *kinds[n])
        assert np.array_equal(u, want_used), n
        assert np.array_equal(o, want), n


def test_resident_from_filtered_stream(golden):
    """PngImage::new's unfilter step on the device (oxi_image_from_filtered) equals the decoded fixture."""
    from oxipng_b200.resident import DeviceImage
    import oxipng_b200 as ox
    n = 0
    for c in golden[:24]:
        img = helpers.product_image(c.ihdr, c.data, c.palette, c.trns)
        filt, _ = img.filter_image(ox.FilterStrategy.Entropy, False)
        dev = DeviceImage.from_filtered(img.ihdr, filt)
        _same(dev, img, c.name)
        n += 1
    assert n


def test_resident_flow_takes_the_same_decisions(golden):
    """perform_reductions over resident images (one upload, every scan on the device) submits the same candidates in the same
    order and returns the same baseline as over host buffers (decision parity with the oracle: test_reductions_flow.py)."""
    import oxipng_b200 as ox
    from oxipng_b200 import reduction
    from oxipng_b200.resident import ResidentBackend
    from reductions_driver import Options, perform_reductions
    rb = ResidentBackend()
    for c in golden:
        for opt in ({}, {"optimize_alpha": True}, {"scale_16": True, "grayscale_reduction": False}, {"interlace": None}):
            img = helpers.product_image(c.ihdr, c.data, c.palette, c.trns)
            ev_h, ev_r = helpers.RecordingEvaluator(), helpers.RecordingEvaluator()
            base_h = perform_reductions(img, Options(**opt), ev_h, backend=reduction, interlacer=ox.change_interlacing)
            base_r = perform_reductions(rb.upload(img), Options(**opt), ev_r, backend=rb, interlacer=rb.change_interlacing)
            assert ev_h.tried == ev_r.tried, (c.name, opt)
            assert helpers._img_tuple(base_h) == helpers._img_tuple(base_r), (c.name, opt)
            assert np.array_equal(base_h.data, np.asarray(base_r.data)), (c.name, opt)


def test_concurrent_threads_different_geometries():
    """Twelve host threads filter images of different row lengths, pixel formats and strategy sets at once (rayon workers,
    evaluate.rs:141-153): no launch may disturb another's (function attributes are set once per device, slots are pooled)."""
    import concurrent.futures as cf
    import oxipng_b200 as ox
    shapes = [(300, 40, 6, 8), (64, 64, 2, 8), (1024, 16, 6, 8), (777, 9, 6, 16), (33, 70, 0, 8), (4000, 6, 2, 8),
              (128, 128, 4, 8), (19, 200, 6, 8), (2048, 8, 6, 8), (500, 20, 2, 16), (96, 96, 0, 16), (1500, 10, 6, 8)]
    cases = []
    for i, (w, h, ct, bd) in enumerate(shapes):
        ih, d = helpers.synth_image(w, h, ct, bd, 900 + i, "photo" if i % 3 else "random")
        cases.append((ih, d, helpers.product_image(ih, d)))
    want = {}
    for i, (ih, d, _) in enumerate(cases):
        for n, k, b in helpers.STRATEGIES:
            want[(i, n)] = oracle.filter_image(ih, d, k, b)[0]

    def work(t):
        i = t % len(cases)
        n = helpers.STRATEGIES[(t // len(cases)) % len(helpers.STRATEGIES)][0]
        out, _ = cases[i][2].filter_image(helpers.product_strategy(n), False)
        return i, n, out

    with cf.ThreadPoolExecutor(max_workers=12) as ex:
        for i, n, out in ex.map(work, range(12 * len(helpers.STRATEGIES) * 2)):
            assert np.array_equal(out, want[(i, n)]), (i, n)


def test_batch_rejects_ragged_data_len():
    import oxipng_b200 as ox
    ih, d = helpers.synth_image(16, 4, 6, 8, 3, "photo")
    img = helpers.product_image(ih, d)
    two = np.concatenate([d, d, np.zeros(5, np.uint8)])  # 2 images "of 64*4+2.5 bytes": not whole scanlines
    with pytest.raises(ox.OxiError):
        ox.filter_batch(img.ihdr, two[: 2 * (d.size + 2)], 2, [ox.FilterStrategy.NONE], False)


def test_pageable_and_pinned_host_buffers_give_the_same_bytes():
    """oxi_filter_batch stages pageable caller memory through its own pinned buffers with copy threads (several 32 MB
    chunks, slots reused) and uses pinned caller memory in place: same bytes either way, equal to the oracle."""
    import torch
    import oxipng_b200 as ox
    w, h, n = 1024, 256, 44  # 46 MB of input: more than one staging chunk, more chunks than slots with the outputs
    imgs = [helpers.synth_image(w, h, 6, 8, 8000 + i, "photo" if i % 5 else "random")[1] for i in range(n)]
    data = np.concatenate(imgs)
    hdr = helpers.product_image((w, h, 6, 8, 0), imgs[0]).ihdr
    S = ox.FilterStrategy
    strat = [S.NONE, S.Bigrams, S.MinSum]
    outs_p, used_p = ox.filter_batch(hdr, data, n, strat, False)  # numpy arrays: pageable
    pin_in = torch.empty(data.size, dtype=torch.uint8, pin_memory=True)
    pin_in.numpy()[:] = data
    per = imgs[0].size + h
    pin_out = [torch.empty(n * per, dtype=torch.uint8, pin_memory=True) for _ in strat]
    pin_used = [torch.empty(n * h, dtype=torch.uint8, pin_memory=True) for _ in strat]
    ox.filter_batch(hdr, pin_in.numpy(), n, strat, False, 0, [t.numpy() for t in pin_out], [t.numpy() for t in pin_used])
    for j in range(len(strat)):
        assert np.array_equal(outs_p[j], pin_out[j].numpy()), j
        assert np.array_equal(used_p[j], pin_used[j].numpy()), j
    kinds = [(oracle.BASIC, 0), (oracle.BIGRAMS, 0), (oracle.MINSUM, 0)]
    for i in (0, 5, n - 1):
        for j, (k, b) in enumerate(kinds):
            want, wu, _ = oracle.filter_image((w, h, 6, 8, 0), imgs[i], k, b)
            assert np.array_equal(outs_p[j][i * per:(i + 1) * per], want), (i, j)
            assert np.array_equal(used_p[j][i * h:(i + 1) * h], wu), (i, j)


def test_sharded_refuses_alpha_and_handles_reject_bad_arguments():
    import ctypes as C
    import oxipng_b200 as ox
    from oxipng_b200 import shard
    from oxipng_b200._lib import CIhdr, CStrategy, lib
    from oxipng_b200.resident import DeviceImage, _setup
    ih, d = helpers.synth_image(64, 32, 6, 8, 1, "transparent")
    img = helpers.product_image(ih, d)
    L = _setup(lib())
    h = img.ihdr._c()
    cs = (CStrategy * 1)()
    cs[0], _k = ox.FilterStrategy.MinSum._c()
    out = np.empty(d.size + 32, np.uint8)
    used = np.empty(32, np.uint8)
    po, pu = (C.c_void_p * 1)(out.ctypes.data), (C.c_void_p * 1)(used.ctypes.data)
    devs = (C.c_int * 1)(0)
    L.oxi_filter_image_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(CStrategy), C.c_size_t, C.c_int,
                                           C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_void_p), C.c_uint]
    # optimize_alpha with an alpha channel: rows are serial, no row bands
    rc = L.oxi_filter_image_sharded(C.byref(h), C.c_void_p(d.ctypes.data), d.size, cs, 1, 1, devs, 1, 1, po, pu, 0)
    assert rc == -4 and b"serial" in L.oxi_last_error()
    # handles: a null image and a too small download buffer are errors, not crashes
    assert L.oxi_image_info(None, None, None, None) < 0
    dev = DeviceImage.upload(img)
    small = np.empty(8, np.uint8)
    got = C.c_size_t()
    assert L.oxi_image_download(dev._h, C.c_void_p(small.ctypes.data), 8, C.byref(got)) < 0
    out_h = C.c_void_p()
    assert L.oxi_reduced_alpha_channel_device(None, 0, C.byref(out_h)) < 0
