"""GPU parity: the CUDA filter_image (through the C ABI) against the CPU oracle, byte for byte.

Bar: bit-exact filtered stream and per-row filter choice for every strategy (integer/byte work).
"""
import numpy as np
import pytest

import helpers
import oracle

pytestmark = pytest.mark.gpu


def _check(ihdr, data, palette=None, trns=None, alphas=(False, True), flags=0, strategies=None, label=""):
    import oxipng_b200 as ox
    img = helpers.product_image(ihdr, data, palette, trns)
    names = strategies or [s[0] for s in helpers.STRATEGIES]
    table = {s[0]: s for s in helpers.STRATEGIES}
    for oa in alphas:
        strat = [helpers.product_strategy(n) for n in names]
        outs, used = ox.filter_image_multi(img, strat, oa, flags)
        for n, o, u in zip(names, outs, used):
            _, kind, basic = table[n]
            want, want_used, _ = oracle.filter_image(ihdr, data, kind, basic, None, oa)
            assert np.array_equal(u, want_used), (label, n, oa, "filter choice",
                                                  np.flatnonzero(u != want_used)[:8].tolist())
            assert np.array_equal(o, want), (label, n, oa, "stream", int(np.argmax(o != want)))


def test_golden_fixtures_generic_tier(golden):
    import oxipng_b200 as ox
    for c in golden:
        _check(c.ihdr, c.data, c.palette, c.trns, flags=ox.FLAG_FORCE_GENERIC, label=c.name)


def test_golden_fixtures_default_tier_and_hashes(golden):
    """Default dispatch (fast tier where it applies) against the oracle AND the committed golden hashes."""
    import oxipng_b200 as ox
    for c in golden:
        img = helpers.product_image(c.ihdr, c.data, c.palette, c.trns)
        for oa in (0, 1):
            names = [s[0] for s in helpers.STRATEGIES]
            outs, used = ox.filter_image_multi(img, [helpers.product_strategy(n) for n in names], bool(oa))
            for n, o, u in zip(names, outs, used):
                e = c.expected["filters"][f"{n}/{oa}"]
                assert bytes(u).hex() == e["used"], (c.name, n, oa)
                assert helpers.sha(o) == e["sha"], (c.name, n, oa)


@pytest.mark.parametrize("ct,bd", [(0, 1), (0, 2), (0, 4), (3, 1), (3, 4), (0, 8), (3, 8), (4, 8), (2, 8), (6, 8),
                                   (0, 16), (4, 16), (2, 16), (6, 16)])
def test_small_and_ragged_images(ct, bd):
    """widths 1..40 incl. rows where len == bpp, heights 1..9, interlaced and not, all strategies, alpha on/off."""
    for il in (False, True):
        for (w, h) in [(1, 1), (1, 5), (2, 2), (3, 1), (5, 3), (7, 9), (8, 8), (13, 6), (17, 5), (31, 4), (33, 9), (40, 3)]:
            for kind in ("photo", "transparent", "random", "zeros"):
                if kind == "transparent" and ct not in (4, 6):
                    continue
                ih, d = helpers.synth_image(w, h, ct, bd, 1000 + w * 31 + h, kind, il)
                _check(ih, d, alphas=(False, True) if ct in (4, 6) else (False,), label=f"{ct}/{bd}/{w}x{h}/{kind}/{il}")


def test_predefined_and_returned_strategy():
    import oxipng_b200 as ox
    ih, d = helpers.synth_image(64, 40, 6, 8, 5, "transparent")
    img = helpers.product_image(ih, d)
    for oa in (False, True):
        out, st = img.filter_image(ox.FilterStrategy.MinSum, oa)
        assert st.kind == "Predefined" and len(st.lines) == 40
        # re-filtering with the returned Predefined list reproduces lib.rs:503's final pass
        want, used, _ = oracle.filter_image(ih, d, oracle.PREDEFINED, 0, np.array(st.lines, np.uint8), oa)
        out2, st2 = img.filter_image(st, oa)
        assert st2 is st and np.array_equal(out2, want)
        # short list: missing rows use None (png/mod.rs:389)
        short = ox.FilterStrategy.Predefined([4, 3, 2])
        want, _, _ = oracle.filter_image(ih, d, oracle.PREDEFINED, 0, np.array([4, 3, 2], np.uint8), oa)
        assert np.array_equal(img.filter_image(short, oa)[0], want)
    out, st = img.filter_image(ox.FilterStrategy.PAETH, False)
    assert st is ox.FilterStrategy.PAETH


def test_ties_prefer_lowest_filter():
    # constant rows: Sub/Up/Paeth all give zeros except edges; exercise tie-breaking on crafted data
    for ct, bd in [(0, 8), (2, 8), (6, 8)]:
        ih, d = helpers.synth_image(48, 12, ct, bd, 3, "flat")
        _check(ih, d, alphas=(False,), label="flat")


def test_batch_matches_per_image():
    import oxipng_b200 as ox
    imgs = [helpers.synth_image(96, 33, 6, 8, 200 + i, "photo") for i in range(7)]
    ih = imgs[0][0]
    data = np.concatenate([d for _, d in imgs])
    hdr = helpers.product_image(ih, imgs[0][1]).ihdr
    strat = [ox.FilterStrategy.NONE, ox.FilterStrategy.Bigrams, ox.FilterStrategy.BigEnt, ox.FilterStrategy.MinSum,
             ox.FilterStrategy.Entropy]
    outs, used = ox.filter_batch(hdr, data, len(imgs), strat, False)
    kinds = [(oracle.BASIC, 0), (oracle.BIGRAMS, 0), (oracle.BIGENT, 0), (oracle.MINSUM, 0), (oracle.ENTROPY, 0)]
    for j, (kind, basic) in enumerate(kinds):
        want = [oracle.filter_image(ih, d, kind, basic) for _, d in imgs]
        assert np.array_equal(outs[j], np.concatenate([w[0] for w in want])), j
        assert np.array_equal(used[j], np.concatenate([w[1] for w in want])), j


def test_device_resident_api_matches_host_api():
    import torch
    import oxipng_b200 as ox
    ih, d = helpers.synth_image(128, 64, 2, 8, 77, "photo")
    hdr = helpers.product_image(ih, d).ihdr
    n = 5
    data = np.concatenate([np.roll(d, i * 17) for i in range(n)])
    strat = [ox.FilterStrategy.MinSum, ox.FilterStrategy.Entropy]
    want_o, want_u = ox.filter_batch(hdr, data, n, strat, False)
    dev = torch.device("cuda:0")
    t_in = torch.from_numpy(data).to(dev)
    rows = 64
    t_out = [torch.empty(n * (d.size + rows), dtype=torch.uint8, device=dev) for _ in strat]
    t_used = [torch.empty(n * rows, dtype=torch.uint8, device=dev) for _ in strat]
    stream = torch.cuda.current_stream(dev)
    ox.filter_batch_device(0, stream.cuda_stream, hdr, t_in.data_ptr(), d.size, n, strat, False,
                           [t.data_ptr() for t in t_out], [t.data_ptr() for t in t_used])
    stream.synchronize()
    for j in range(len(strat)):
        assert np.array_equal(t_out[j].cpu().numpy(), want_o[j])
        assert np.array_equal(t_used[j].cpu().numpy(), want_u[j])


def test_fast_tier_covers_benchmark_shapes():
    """The fast (TMA-staged) tier, forced, on scaled-down versions of every BASELINE config incl. rows that are and are
    not 16-byte multiples, every filter bpp (1,2,3,4,6,8) and interlaced passes."""
    import oxipng_b200 as ox
    shapes = [
        (1024, 40, 6, 8, False, "photo"),     # C1/C5: RGBA8, TMA-aligned rows
        (2048, 24, 2, 8, False, "photo"),     # C2: RGB8
        (512, 32, 6, 16, False, "photo"),     # C3a: RGBA16 (bpp 8)
        (512, 64, 6, 8, True, "photo"),       # C3b: Adam7 RGBA8
        (500, 33, 2, 8, False, "photo"),      # reference fixture geometry: 1500-byte rows, no TMA
        (333, 17, 4, 8, False, "random"),     # bpp 2, odd
        (257, 19, 2, 16, True, "random"),     # bpp 6 interlaced
        (1000, 9, 0, 8, False, "photo"),      # bpp 1
        (999, 21, 0, 4, False, "random"),     # packed 4-bit
        (640, 16, 3, 1, True, "random"),      # packed 1-bit interlaced
        (256, 20, 6, 8, False, "flat"),       # heavy ties / hot histogram bins
        (256, 20, 6, 8, False, "zeros"),      # all-zero shortcut
        (4096, 6, 6, 16, False, "random"),    # 32 KB rows: largest bigram-table configuration
    ]
    for (w, h, ct, bd, il, kind) in shapes:
        ih, d = helpers.synth_image(w, h, ct, bd, 4242 + w, kind, il)
        img = helpers.product_image(ih, d)
        assert ox.filter_tier(img, ox.FilterStrategy.BigEnt, False) == "fast", (w, h, ct, bd)
        # one strategy per launch (specialised kernels) ...
        for n in [s[0] for s in helpers.STRATEGIES]:
            _check(ih, d, alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=[n], label=f"fast1/{w}x{h}/{ct}/{bd}/{il}/{kind}")
        # ... and the fused multi-strategy launch (evaluate.rs fan-out)
        # (all scorers resident at once need 3 rows + 148 KB of tables: 32 KB rows fall back to the generic tier)
        fused_flags = ox.FLAG_FORCE_FAST if helpers.row_bytes(ih) <= 16384 else 0
        _check(ih, d, alphas=(False,), flags=fused_flags, label=f"fastN/{w}x{h}/{ct}/{bd}/{il}/{kind}")
        _check(ih, d, alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=["None", "Bigrams", "BigEnt"], label="o4")


def test_bigram_outlier_paths():
    """Bigram scorers on smooth images with a controlled number of spikes per row: few outlier windows (hash-table
    path), around the hash capacity, and far beyond it (big-table path), plus hot spikes (one value repeated) that
    stress the hash counters; 1-word and 2-words-per-thread row lengths."""
    import oxipng_b200 as ox
    rng = np.random.default_rng(77)
    for (w, h) in [(1024, 12), (2000, 10), (300, 9)]:
        for spikes in (0, 3, 40, 110, 128, 135, 300, 1500):
            for hot in (False, True):
                ih, d = helpers.synth_image(w, h, 6, 8, 900 + spikes, "photo")
                a = np.frombuffer(bytes(d), dtype=np.uint8).copy().reshape(h, -1)
                for y in range(h):
                    n = min(spikes, a.shape[1])
                    pos = rng.choice(a.shape[1], size=n, replace=False)
                    a[y, pos] = 200 if hot else rng.integers(0, 256, size=n, dtype=np.uint8)
                _check(ih, a.tobytes(), alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=["Bigrams", "BigEnt"],
                       label=f"spikes/{w}x{h}/{spikes}/{hot}")
                _check(ih, a.tobytes(), alphas=(False,), flags=ox.FLAG_FORCE_FAST,
                       strategies=["None", "Bigrams", "BigEnt"], label=f"spikes-o4/{w}x{h}/{spikes}/{hot}")


def test_bigram_none_trial_bounds_and_exact_pass():
    """The fast bigram scorer scores the None trial on coarse keys (low five bits) and reruns it exactly only when the
    bounds leave open that None wins.  Rows built so that None wins, ties, or loses narrowly: few-valued rows (graphics),
    rows whose values differ only above bit 4 (coarse keys collide, exact keys do not), constant rows, and rows where
    Sub/Up are as good as None."""
    import oxipng_b200 as ox
    rng = np.random.default_rng(4242)
    # (4096 px RGBA8 = 16 KB rows: the longest rows of the dense-table path; a constant row there has one bigram 16384 times)
    for (w, h, ct) in [(1024, 24, 6), (333, 17, 2), (2048, 9, 0), (4096, 8, 6), (3000, 8, 6)]:
        ch = {0: 1, 2: 3, 6: 4}[ct]
        L = w * ch
        rows = []
        for y in range(h):
            kind = y % 8
            if kind == 0:      # two-valued stripes: None has very few bigrams
                r = np.where((np.arange(L) // (3 + y)) % 2 == 0, 17, 201)
            elif kind == 1:    # values that differ only in the high bits: coarse keys merge, exact keys do not
                r = 5 + 32 * rng.integers(0, 8, size=L)
            elif kind == 2:    # constant row
                r = np.full(L, 90 + y)
            elif kind == 3:    # equal to the row above (Up gives zeros), None still cheap
                r = rows[-1].copy()
            elif kind == 4:    # small alphabet, random order
                r = rng.choice(np.array([0, 1, 2, 64, 65, 130]), size=L)
            elif kind == 5:    # period-bpp pattern: Sub gives zeros after the first pixel
                r = np.tile(np.array([10, 60, 110, 160, 210, 5, 100, 200])[:ch], w)
            elif kind == 6:    # noise: every trial is bad, None's bounds are loose
                r = rng.integers(0, 256, size=L)
            else:              # smooth ramp
                r = (np.arange(L) // ch) // 5 + 3 * (np.arange(L) % ch)
            rows.append(np.asarray(r, dtype=np.int64) & 0xFF)
        a = np.stack(rows).astype(np.uint8)
        ih = (w, h, ct, 8, 0)
        _check(ih, a.tobytes(), alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=["Bigrams", "BigEnt"],
               label=f"none-bounds/{w}x{h}/{ct}")
        _check(ih, a.tobytes(), alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=["BigEnt"], label="none-bounds-be")
        _check(ih, a.tobytes(), alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=["Bigrams"], label="none-bounds-bg")
        _check(ih, a.tobytes(), alphas=(False,), flags=ox.FLAG_FORCE_FAST,
               strategies=["MinSum", "Entropy", "Bigrams", "BigEnt"], label="none-bounds-all")


def test_randomised_shapes_fast_tier():
    """Differential sweep over random widths / heights / pixel formats / contents through the fast tier: odd numbers of
    row words (the bigram scorer works on word pairs), tails of 1..3 bytes, every filter-bpp (1, 2, 3, 4, 6, 8), rows that
    cross the 256-thread / 1024-thread kernel boundary (4 KB) and the dense-table limit (16 KB)."""
    import oxipng_b200 as ox
    rng = np.random.default_rng(20261020)
    fmts = [(0, 8), (4, 8), (2, 8), (6, 8), (0, 16), (4, 16), (2, 16), (6, 16), (3, 8)]
    for it in range(36):
        ct, bd = fmts[int(rng.integers(len(fmts)))]
        ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ct]
        maxw = 17000 // (ch * bd // 8)
        w = int(rng.integers(1, maxw)) if it % 3 else int(rng.integers(1, 64))
        h = int(rng.integers(1, 9))
        kind = ["photo", "random", "flat", "photo"][int(rng.integers(4))]
        ih, d = helpers.synth_image(w, h, ct, bd, 7000 + it, kind)
        pal = np.zeros((256, 4), np.uint8) if ct == 3 else None
        _check(ih, d, palette=pal, alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=["Bigrams", "BigEnt"],
               label=f"rnd/{it}/{w}x{h}/{ct}/{bd}/{kind}")
        if it % 4 == 0:
            _check(ih, d, palette=pal, alphas=(False,), strategies=["None", "MinSum", "Entropy", "Bigrams", "BigEnt"],
                   label=f"rnd-all/{it}/{w}x{h}/{ct}/{bd}/{kind}")


def test_paeth4_exhaustive_on_device():
    """The byte-SIMD Paeth predictor (common.cuh) against the scalar definition for all 2^24 (a, b, c) in every byte lane:
    builds tests/cuda/paeth4_exhaustive.cu with nvcc and runs it."""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cuda")
    exe = os.path.join(here, "paeth4_exhaustive")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-o", exe,
                           os.path.join(here, "paeth4_exhaustive.cu")], stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "mismatches=0" in out.stdout, out.stdout + out.stderr


def test_fast_tier_batch_many_images():
    import oxipng_b200 as ox
    n = 37
    imgs = [helpers.synth_image(256, 48, 6, 8, 900 + i, "photo" if i % 3 else "random") for i in range(n)]
    ih = imgs[0][0]
    hdr = helpers.product_image(ih, imgs[0][1]).ihdr
    data = np.concatenate([d for _, d in imgs])
    strat = [ox.FilterStrategy.NONE, ox.FilterStrategy.Bigrams, ox.FilterStrategy.BigEnt]
    outs, used = ox.filter_batch(hdr, data, n, strat, False, ox.FLAG_FORCE_FAST)
    for j, (kind, basic) in enumerate([(oracle.BASIC, 0), (oracle.BIGRAMS, 0), (oracle.BIGENT, 0)]):
        want = [oracle.filter_image(ih, d, kind, basic) for _, d in imgs]
        assert np.array_equal(used[j], np.concatenate([w[1] for w in want])), j
        assert np.array_equal(outs[j], np.concatenate([w[0] for w in want])), j


def test_concurrent_host_threads():
    """filter_image is called concurrently from rayon workers on one shared image (evaluate.rs:141-153): many host threads
    through the host-buffer C ABI at once, every result bit-exact."""
    import concurrent.futures as cf
    import oxipng_b200 as ox
    ih, d = helpers.synth_image(300, 64, 6, 8, 31, "photo")
    img = helpers.product_image(ih, d)
    want = {n: oracle.filter_image(ih, d, k, b) for n, k, b in helpers.STRATEGIES}

    def work(i):
        n = helpers.STRATEGIES[i % len(helpers.STRATEGIES)][0]
        out, st = img.filter_image(helpers.product_strategy(n), False)
        return n, out

    with cf.ThreadPoolExecutor(max_workers=12) as ex:
        for n, out in ex.map(work, range(96)):
            assert np.array_equal(out, want[n][0]), n


def test_repeatability_stress():
    """The fast kernels run with as few CTA-wide barriers as the data hazards allow; a missed hazard would show up as
    run-to-run differences or as a mismatch on some image of a larger batch.  (compute-sanitizer is closed on this pool.)"""
    import oxipng_b200 as ox
    n = 48
    for (w, h, ct, bd, kind) in [(512, 40, 6, 8, "photo"), (300, 33, 2, 8, "photo"), (1024, 12, 6, 16, "random")]:
        imgs = [helpers.synth_image(w, h, ct, bd, 5000 + i, kind if i % 4 else "random") for i in range(n)]
        ih = imgs[0][0]
        hdr = helpers.product_image(ih, imgs[0][1]).ihdr
        data = np.concatenate([d for _, d in imgs])
        strat = [ox.FilterStrategy.NONE, ox.FilterStrategy.Bigrams, ox.FilterStrategy.BigEnt, ox.FilterStrategy.MinSum,
                 ox.FilterStrategy.Entropy]
        kinds = [(oracle.BASIC, 0), (oracle.BIGRAMS, 0), (oracle.BIGENT, 0), (oracle.MINSUM, 0), (oracle.ENTROPY, 0)]
        first = None
        for rep in range(4):
            for group in ([0, 1, 2], [3], [4]):
                outs, used = ox.filter_batch(hdr, data, n, [strat[g] for g in group], False)
                key = tuple(helpers.sha(o) for o in outs) + tuple(helpers.sha(u) for u in used)
                if rep == 0:
                    for g, o, u in zip(group, outs, used):
                        want = [oracle.filter_image(ih, d, *kinds[g]) for _, d in imgs]
                        assert np.array_equal(u, np.concatenate([x[1] for x in want])), (w, h, g)
                        assert np.array_equal(o, np.concatenate([x[0] for x in want])), (w, h, g)
                    first = first or {}
                    first[tuple(group)] = key
                else:
                    assert key == first[tuple(group)], ("run-to-run difference", w, h, group, rep)


def test_rows_beyond_the_fast_tier_fall_back_to_generic():
    """Rows that do not fit the shared-memory ring (here 120 KB and 160 KB per row) take the generic tier by themselves."""
    import oxipng_b200 as ox
    for (w, h, ct, bd) in [(30000, 3, 6, 8), (20000, 2, 6, 16)]:
        ih, d = helpers.synth_image(w, h, ct, bd, 77, "photo")
        img = helpers.product_image(ih, d)
        assert ox.filter_tier(img, ox.FilterStrategy.MinSum, False) == "generic"
        _check(ih, d, alphas=(False,), strategies=["Sub", "Paeth", "MinSum", "Entropy", "Bigrams", "BigEnt"], label=f"huge/{w}")
    # 40 KB rows: fast tier for MinSum/Entropy, generic for the bigram scorers (table + ring exceed shared memory)
    ih, d = helpers.synth_image(10000, 3, 6, 8, 78, "photo")
    img = helpers.product_image(ih, d)
    assert ox.filter_tier(img, ox.FilterStrategy.MinSum, False) == "fast"
    assert ox.filter_tier(img, ox.FilterStrategy.BigEnt, False) == "generic"
    _check(ih, d, alphas=(False,), strategies=["MinSum", "Entropy", "Bigrams", "BigEnt"], label="40KB rows")


def test_argument_errors_are_reported():
    import ctypes as C
    import oxipng_b200 as ox
    from oxipng_b200 import _lib
    L = ox.lib()
    h = _lib.CIhdr(4, 4, 6, 8, 0)
    d = np.zeros(64, np.uint8)
    out = np.zeros(80, np.uint8)
    used = np.zeros(4, np.uint8)
    bad = _lib.CStrategy(9, 0, None, 0)
    assert L.oxi_filter_image(C.byref(h), d.ctypes.data, 64, C.byref(bad), 0, out.ctypes.data, used.ctypes.data, None) == -1
    bad = _lib.CStrategy(0, 7, None, 0)
    assert L.oxi_filter_image(C.byref(h), d.ctypes.data, 64, C.byref(bad), 0, out.ctypes.data, used.ctypes.data, None) == -1
    assert b"strategy" in L.oxi_last_error()
    # empty image data: nothing to do, no rows
    ok = _lib.CStrategy(1, 0, None, 0)
    assert L.oxi_filter_image(C.byref(h), d.ctypes.data, 0, C.byref(ok), 0, out.ctypes.data, used.ctypes.data, None) == 0


def test_split_table_bigram_rows_16k_to_32k():
    """Bigrams / BigEnt on rows of 16 KB .. 32 KB (bigram_row_split: tables A[a<32][b] and B[a][b<32], outlier hash, coarse
    trial 0 with exact recount, flagged trials recounted by bigram_trial): 16-bit images whose low bytes are noise, 8-bit
    images, every filter bpp, ragged row lengths, Adam7, and contents that force each fallback."""
    import oxipng_b200 as ox
    rng = np.random.default_rng(1234)
    shapes = [
        (2100, 7, 6, 16, False, "photo"),     # 16 800-byte rows, bpp 8
        (4096, 5, 6, 16, False, "photo"),     # 32 768-byte rows (the c3a shape)
        (4095, 4, 2, 16, False, "photo"),     # bpp 6, 24 570-byte rows (not a multiple of 4)
        (8000, 4, 4, 16, False, "photo"),     # bpp 4
        (10000, 4, 0, 16, False, "photo"),    # bpp 2
        (4200, 6, 6, 8, False, "photo"),      # 8-bit, bpp 4: everything lands in table A
        (5467, 5, 2, 8, False, "photo"),      # bpp 3, 16 401-byte rows
        (20001, 4, 0, 8, False, "photo"),     # bpp 1
        (4096, 5, 6, 16, False, "random"),    # both residuals far from zero: outlier hash overflows, exact recounts
        (4096, 4, 6, 16, False, "flat"),      # ties, None competitive
        (4096, 4, 6, 16, False, "zeros"),     # all-zero shortcut
        (4000, 40, 6, 16, True, "photo"),     # Adam7: passes with rows of 4 .. 32 KB in one launch
    ]
    for (w, h, ct, bd, il, kind) in shapes:
        ih, d = helpers.synth_image(w, h, ct, bd, 777 + w, kind, il)
        _check(ih, d, alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=["Bigrams", "BigEnt"],
               label=f"split/{w}x{h}/{ct}/{bd}/{il}/{kind}")
        _check(ih, d, alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=["None", "Bigrams", "BigEnt"],
               label=f"split3/{w}x{h}/{ct}/{bd}/{il}/{kind}")
    # rows where None wins or ties (exact trial 0), and rows with a few far-out windows (outlier hash without overflow)
    w, h = 4096, 6
    a = np.zeros((h, w * 8), np.uint8)
    a[1] = np.tile(np.array([7, 9], np.uint8), w * 4)             # two bigrams only: None beats every filter
    a[2] = a[1]
    a[3] = rng.integers(0, 3, size=w * 8, dtype=np.uint8)         # tiny alphabet
    a[4] = a[3]
    a[4, rng.choice(w * 8, size=50, replace=False)] = 200          # 50 spikes: outlier windows in every trial
    a[5] = a[4]
    _check((w, h, 6, 16, 0), a.reshape(-1), alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=["Bigrams", "BigEnt"],
           label="split/none-wins")


def test_default_trial_set_entropy_plus_bigrams_kernel():
    """The reference's default set {None, Sub, Entropy, Bigrams} (options.rs:282-287) runs in its own kernels
    (k_fast<D, SC_ENTROPY|SC_BIGRAM[|SC_HALF]>): short rows (256-thread CTAs), long rows, every bpp class, odd lengths."""
    import oxipng_b200 as ox
    for (w, h, ct, bd, il, kind) in [(1024, 9, 6, 8, False, "photo"), (300, 12, 2, 8, False, "photo"), (333, 7, 6, 16, False, "photo"),
                                     (1000, 6, 0, 8, False, "random"), (256, 20, 6, 8, False, "flat"), (256, 8, 6, 8, False, "zeros"),
                                     (3000, 5, 6, 8, False, "photo"), (777, 40, 6, 8, True, "photo"), (999, 9, 0, 4, False, "random")]:
        ih, d = helpers.synth_image(w, h, ct, bd, 31 + w, kind, il)
        _check(ih, d, alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=["None", "Sub", "Entropy", "Bigrams"],
               label=f"o2/{w}x{h}/{ct}/{bd}/{il}/{kind}")
        _check(ih, d, alphas=(False,), flags=ox.FLAG_FORCE_FAST, strategies=["Entropy", "BigEnt"], label="E+BE")


def test_alpha_tier_staged_rows():
    """optimize_alpha through the shared-memory tier (k_alpha): every alpha pixel format, rows up to 16 KB, Adam7, runs of
    transparent pixels of every length (the run walkers), first pixel transparent, rows that are transparent throughout
    with colour left behind (the cumulative-rewrite leak, filters/mod.rs:126-131), all-zero rows, all nine strategies."""
    import oxipng_b200 as ox
    rng = np.random.default_rng(99)
    for (w, h, ct, bd, il) in [(64, 40, 6, 8, False), (300, 33, 6, 8, False), (1024, 12, 6, 8, False), (4000, 5, 6, 8, False),
                               (257, 21, 4, 8, False), (190, 17, 6, 16, False), (333, 9, 4, 16, False), (131, 37, 6, 8, True),
                               (90, 30, 4, 16, True), (2048, 6, 6, 16, False)]:
        ih, d = helpers.synth_image(w, h, ct, bd, 4000 + w, "transparent", il)
        img = helpers.product_image(ih, d)
        assert ox.filter_tier(img, ox.FilterStrategy.MinSum, True) == "alpha", (w, h, ct, bd)
        _check(ih, d, alphas=(True,), label=f"alpha/{w}x{h}/{ct}/{bd}/{il}")
    # hand-built RGBA8 rows: long runs, leading / trailing runs, fully transparent rows with colour, alternate pixels
    w, h = 600, 14
    a = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    a[..., 3] = 255
    a[1, :400, 3] = 0                 # leading run of 400
    a[2, 100:, 3] = 0                 # trailing run
    a[3, :, 3] = 0                    # whole row transparent, colours random (nothing to copy from: prev == i)
    a[4, :, 3] = 0
    a[4, :, :3] = 0                   # whole row zero: None shortcut
    a[5, ::2, 3] = 0                  # alternate pixels
    a[6, 1::2, 3] = 0
    a[7, 5:595, 3] = 0                # one long interior run
    a[8, :, 3] = 0                    # transparent row after a long run row
    a[9, :3, 3] = 0
    a[10, 0, 3] = 0
    a[11, -1, 3] = 0
    a[12, :, 3] = (rng.random(w) < 0.5) * 255
    _check((w, h, 6, 8, 0), a.reshape(-1), alphas=(True,), label="alpha/handmade")
    pre = [(i * 3) % 5 for i in range(h)]
    img = helpers.product_image((w, h, 6, 8, 0), a.reshape(-1))
    want, _, _ = oracle.filter_image((w, h, 6, 8, 0), a.reshape(-1), oracle.PREDEFINED, 0, np.array(pre, np.uint8), True)
    assert np.array_equal(img.filter_image(ox.FilterStrategy.Predefined(pre), True)[0], want)
    # a batch: chains of different images and strategies run side by side
    n = 9
    imgs = [helpers.synth_image(200, 25, 6, 8, 700 + i, "transparent")[1] for i in range(n)]
    hdr = helpers.product_image((200, 25, 6, 8, 0), imgs[0]).ihdr
    S = ox.FilterStrategy
    strat = [S.NONE, S.SUB, S.Entropy, S.Bigrams, S.MinSum, S.BigEnt]
    outs, used = ox.filter_batch(hdr, np.concatenate(imgs), n, strat, True)
    per = imgs[0].size + 25
    kinds = [(oracle.BASIC, 0), (oracle.BASIC, 1), (oracle.ENTROPY, 0), (oracle.BIGRAMS, 0), (oracle.MINSUM, 0), (oracle.BIGENT, 0)]
    for i in range(n):
        for j, (k, b) in enumerate(kinds):
            wnt, wu, _ = oracle.filter_image((200, 25, 6, 8, 0), imgs[i], k, b, None, True)
            assert np.array_equal(outs[j][i * per:(i + 1) * per], wnt), (i, j)
            assert np.array_equal(used[j][i * 25:(i + 1) * 25], wu), (i, j)


def test_bigram_wide_tables_and_image_classes():
    """Batches whose images differ in residual spread: k_spread_probe classes them on the device and two launches share the
    batch (32 x 32 tables / 64 x 64 tables, k_fast<.., SC_BIGRAM|SC_HALF|SC_WIDE>).  Noise of +-4 .. +-60 and pure noise,
    ragged widths, every bpp class, plus spikes beyond +-32 (outlier hashes of the wide tables)."""
    import oxipng_b200 as ox
    rng = np.random.default_rng(2024)
    for (w, h, ct, bd) in [(256, 24, 6, 8), (300, 17, 2, 8), (1000, 9, 0, 8), (511, 12, 4, 8), (250, 10, 6, 16)]:
        ih0, base = helpers.synth_image(w, h, ct, bd, 5, "photo")
        imgs = []
        for amp in (0, 4, 10, 12, 20, 31, 40, 60, 128):
            d = np.frombuffer(bytes(base), np.uint8).astype(np.int64)
            if amp:
                d = (d + rng.integers(-amp, amp + 1, size=d.size)) & 255
            d = d.astype(np.uint8)
            if amp in (12, 31):
                d[rng.choice(d.size, size=d.size // 50, replace=False)] = 255  # far-out spikes
            imgs.append(d)
        hdr = helpers.product_image(ih0, imgs[0]).ihdr
        S = ox.FilterStrategy
        for strat, kinds in (([S.NONE, S.Bigrams, S.BigEnt], [(oracle.BASIC, 0), (oracle.BIGRAMS, 0), (oracle.BIGENT, 0)]),
                             ([S.BigEnt], [(oracle.BIGENT, 0)])):
            outs, used = ox.filter_batch(hdr, np.concatenate(imgs), len(imgs), strat, False, ox.FLAG_FORCE_FAST)
            rows = len(helpers.product_image(ih0, imgs[0]).scan_lines())
            per = imgs[0].size + rows
            for i, d in enumerate(imgs):
                for j, (k, b) in enumerate(kinds):
                    want, wu, _ = oracle.filter_image(ih0, d, k, b)
                    assert np.array_equal(used[j][i * rows:(i + 1) * rows], wu), (w, ct, bd, i, j)
                    assert np.array_equal(outs[j][i * per:(i + 1) * per], want), (w, ct, bd, i, j)
