"""Kernel time of oxi_unfilter_image_device (CUDA events around the resident call: the boundary-word memset, k_unfilter and
the few-byte flag read-back) on filtered streams produced by this library's own filter search, checked for equality
with the pixels the stream was made from.  Algorithmic bytes = 2N + S (stream in, pixels out)."""
import json
import statistics
import sys

sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import oxipng_b200 as ox
from oxipng_b200 import synth

PEAK = json.load(open('/root/repo/MEASURED_PEAKS.json')).get('hbm_gbs', 6540.8)


def cases():
    S = ox.FilterStrategy
    yield 'rgba8_4096_minsum', (4096, 4096, 4, 8), S.MinSum
    yield 'rgba8_4096_all_paeth', (4096, 4096, 4, 8), S.PAETH
    yield 'rgba8_4096_all_average', (4096, 4096, 4, 8), S.AVERAGE
    yield 'rgba8_4096_all_up', (4096, 4096, 4, 8), S.UP
    yield 'rgba8_4096_all_sub', (4096, 4096, 4, 8), S.SUB
    yield 'rgb8_8192_minsum', (8192, 8192, 3, 8), S.MinSum
    yield 'rgba16_4096_minsum', (4096, 4096, 4, 16), S.MinSum
    yield 'rgba8_1024_minsum', (1024, 1024, 4, 8), S.MinSum


def main():
    dev = torch.device('cuda:0')
    stream = torch.cuda.current_stream()
    res = {}
    only = sys.argv[1:]
    for name, (w, h, ch, bd), strat in cases():
        if only and name not in only:
            continue
        d = synth.photo_image(w, h, ch, bd, 1, dev)
        ct = {1: ox.ColorType.Grayscale(), 3: ox.ColorType.RGB(), 4: ox.ColorType.RGBA()}[ch]
        hdr = ox.IhdrData(w, h, ct, ox.BitDepth(bd))
        out, used = ox.PngImage(hdr, d.cpu().numpy()).filter_image(strat, False)
        f = torch.from_numpy(np.ascontiguousarray(out)).to(dev)
        pad = torch.zeros(64, dtype=torch.uint8, device=dev)
        fin = torch.cat([f, pad])  # slack behind the stream (aligned word loads)
        o = torch.empty(d.numel() + 64, dtype=torch.uint8, device=dev)
        ts = []
        for it in range(8):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            n = ox.unfilter_image_device(hdr, fin.data_ptr(), f.numel(), o.data_ptr(), 0, stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ok = bool(n == d.numel() and torch.equal(o[:n], d))
        ms = statistics.median(ts[2:])
        alg = 2 * d.numel() + h
        mix = np.bincount(np.asarray(out).reshape(h, -1)[:, 0], minlength=5).tolist()
        res[name] = {'ms': round(ms, 4), 'algorithmic_GBps': round(alg / ms / 1e6, 1), 'frac': round(alg / ms / 1e6 / PEAK, 4),
                     'parity': ok, 'filter_mix': mix}
        print(name, res[name], flush=True)
    for name, (w, h, ch, bd, n) in [('batch_256x_rgba8_1024_minsum', (1024, 1024, 4, 8, 256)), ('batch_64x_rgb8_2048_minsum', (2048, 2048, 3, 8, 64))]:
        if only and name not in only:
            continue
        ct = {1: ox.ColorType.Grayscale(), 3: ox.ColorType.RGB(), 4: ox.ColorType.RGBA()}[ch]
        hdr = ox.IhdrData(w, h, ct, ox.BitDepth(bd))
        d = synth.photo_batch(n, w, h, ch, bd, 1, dev)
        per = d.numel() // n
        outs = ox.filter_batch(hdr, d.cpu().numpy(), n, [ox.FilterStrategy.MinSum], False)[0][0]
        f = torch.from_numpy(np.ascontiguousarray(outs)).to(dev).reshape(-1)
        flen = f.numel() // n
        fin = torch.cat([f, torch.zeros(64, dtype=torch.uint8, device=dev)])
        o = torch.empty(d.numel() + 64, dtype=torch.uint8, device=dev)
        ts = []
        for it in range(8):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            m = ox.unfilter_batch_device(hdr, fin.data_ptr(), flen, n, o.data_ptr(), 0, stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ok = bool(m == per and torch.equal(o[:d.numel()], d))
        ms = statistics.median(ts[2:])
        alg = 2 * d.numel() + h * n
        res[name] = {'ms': round(ms, 4), 'algorithmic_GBps': round(alg / ms / 1e6, 1), 'frac': round(alg / ms / 1e6 / PEAK, 4),
                     'parity': ok, 'images_per_s': round(n / ms * 1e3)}
        print(name, res[name], flush=True)
    if not only:
        json.dump(res, open('/root/repo/gpurun_out/unfilter_times.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
