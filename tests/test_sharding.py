"""Sharding logic: image shards and one-row-halo row bands, incl. a world-size-2 gloo run on CPU.

On CPU the per-band compute stand-in is the oracle (tests may use it); the GPU variant goes through the C ABI."""
import os
import socket
import sys

import numpy as np
import pytest

import helpers
import oracle
from oxipng_b200 import shard


def test_image_shards_partition_the_batch():
    for n in (0, 1, 7, 8, 1250, 10000, 10001):
        for world in (1, 2, 4, 8):
            parts = [shard.image_shard(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (s0, c0), (s1, _) in zip(parts, parts[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _oracle_lines(ih, n):
    ln, ps, px = oracle.scan_lines(ih, n)
    return [(int(a), int(b), int(c)) for a, b, c in zip(ln, ps, px)]


def _band_with_oracle(ih, data, b: shard.Band, kind, basic):
    rows = b.n_rows + (1 if b.has_halo else 0)
    sub_ih = (b.width_px, rows, ih[2], ih[3], 0)
    o, u, _ = oracle.filter_image(sub_ih, data[b.in_begin:b.in_end], kind, basic)
    skip = 1 if b.has_halo else 0
    return o[skip * (b.row_len + 1):], u[skip:]


@pytest.mark.parametrize("il", [False, True])
def test_row_bands_with_halo_reproduce_the_whole_image(il):
    for ct, bd in [(6, 8), (2, 16), (0, 4)]:
        ih, d = helpers.synth_image(67, 41, ct, bd, 9, "photo", il)
        lines = _oracle_lines(ih, d.size)
        for parts in (1, 2, 3, 8):
            bands = shard.row_bands(lines, parts)
            assert sum(b.n_rows for b in bands) == len(lines)
            for name, kind, basic in helpers.STRATEGIES[3:]:
                want, want_used, _ = oracle.filter_image(ih, d, kind, basic)
                out = np.zeros_like(want)
                used = np.zeros_like(want_used)
                for b in bands:
                    o, u = _band_with_oracle(ih, d, b, kind, basic)
                    out[b.out_begin:b.out_end] = o
                    used[b.first_row:b.first_row + b.n_rows] = u
                assert np.array_equal(out, want) and np.array_equal(used, want_used), (ct, bd, il, parts, name)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) batch sharded by image: each rank filters its shard, rank 0 gathers; no data-path collective
        n = 5
        imgs = [helpers.synth_image(40, 24, 6, 8, 300 + i, "photo") for i in range(n)]
        start, count = shard.image_shard(n, world, rank)
        mine = [oracle.filter_image(imgs[i][0], imgs[i][1], oracle.BIGENT)[0] for i in range(start, start + count)]
        gathered = [None] * world
        dist.all_gather_object(gathered, (start, mine))
        # (2) one image split in row bands with a halo row: band b on rank b % world
        ih, d = helpers.synth_image(64, 64, 2, 8, 5, "photo", True)
        lines = _oracle_lines(ih, d.size)
        bands = shard.row_bands(lines, 2 * world)
        parts = [(bi, _band_with_oracle(ih, d, b, oracle.MINSUM, 0)) for bi, b in enumerate(bands) if bi % world == rank]
        gathered2 = [None] * world
        dist.all_gather_object(gathered2, parts)
        if rank == 0:
            flat = [o for _, lst in sorted(gathered, key=lambda t: t[0]) for o in lst]
            ok1 = all(np.array_equal(flat[i], oracle.filter_image(imgs[i][0], imgs[i][1], oracle.BIGENT)[0]) for i in range(n))
            want, want_used, _ = oracle.filter_image(ih, d, oracle.MINSUM, 0)
            out, used = np.zeros_like(want), np.zeros_like(want_used)
            for lst in gathered2:
                for bi, (o, u) in lst:
                    b = bands[bi]
                    out[b.out_begin:b.out_end] = o
                    used[b.first_row:b.first_row + b.n_rows] = u
            q.put(bool(ok1 and np.array_equal(out, want) and np.array_equal(used, want_used)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_sharding():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(60)
    assert ok and all(p.exitcode == 0 for p in procs)


@pytest.mark.gpu
def test_banded_filtering_on_gpu_matches_oracle():
    import oxipng_b200 as ox
    for il in (False, True):
        ih, d = helpers.synth_image(512, 96, 6, 8, 11, "photo", il)
        img = helpers.product_image(ih, d)
        for name, kind, basic in helpers.STRATEGIES[4:]:
            want, want_used, _ = oracle.filter_image(ih, d, kind, basic)
            out, used = shard.filter_image_banded(img, helpers.product_strategy(name), 5)
            assert np.array_equal(out, want) and np.array_equal(used, want_used), (il, name)


@pytest.mark.gpu
def test_sharded_c_abi_multi_strategy_and_predefined():
    """oxi_filter_image_sharded: several strategies per call, Predefined lists re-based per band, every visible device."""
    import torch
    import oxipng_b200 as ox
    devices = list(range(min(torch.cuda.device_count(), 8)))
    for il in (False, True):
        ih, d = helpers.synth_image(700, 150, 6, 8, 12, "photo", il)
        img = helpers.product_image(ih, d)
        rows = len(img.scan_lines())
        pre = [(i * 7) % 5 for i in range(rows)]
        names = ["None", "Paeth", "MinSum", "Entropy", "Bigrams", "BigEnt"]
        strat = [helpers.product_strategy(n) for n in names] + [ox.FilterStrategy.Predefined(pre)]
        for per in (1, 3):
            outs, used = shard.filter_image_sharded(img, strat, devices, per)
            kinds = {n: (k, b) for n, k, b in helpers.STRATEGIES}
            for n, o, u in zip(names, outs, used):
                want, want_used, _ = oracle.filter_image(ih, d, *kinds[n])
                assert np.array_equal(u, want_used) and np.array_equal(o, want), (il, per, n)
            want, want_used, _ = oracle.filter_image(ih, d, oracle.PREDEFINED, 0, pre)
            assert np.array_equal(used[-1], want_used) and np.array_equal(outs[-1], want), (il, per, "Predefined")
