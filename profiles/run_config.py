"""Runs one BASELINE configuration device-resident a few times (for ncu captures and quick timing).

    python profiles/run_config.py c5 [--images 64] [--iters 3]
configs: c1 (MinSum 4096^2 RGBA8), c2 (Entropy 8192^2 RGB8), c2paeth, c2none, c3a (Bigrams+BigEnt 4096^2 RGBA16),
         c3b (Bigrams+BigEnt 4096^2 RGBA8 Adam7), c5 (None+Bigrams+BigEnt on a batch of 1024^2 RGBA8)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import oxipng_b200 as ox  # noqa: E402
from oxipng_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("config")
ap.add_argument("--images", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--scale", type=int, default=1, help="divide width/height of the single-image configs")
a = ap.parse_args()
dev = torch.device("cuda:0")
S = ox.FilterStrategy
cfg = {
    "c1": (4096, 4096, 4, 8, [S.MinSum], 1),
    "c2": (8192, 8192, 3, 8, [S.Entropy], 1),
    "c2paeth": (8192, 8192, 3, 8, [S.PAETH], 1),
    "c2none": (8192, 8192, 3, 8, [S.NONE], 1),
    "c3a": (4096, 4096, 4, 16, [S.Bigrams, S.BigEnt], 1),
    "c3b": (4096, 4096, 4, 8, [S.Bigrams, S.BigEnt], 1),
    "c5": (1024, 1024, 4, 8, [S.NONE, S.Bigrams, S.BigEnt], a.images),
    "c1bg": (4096, 4096, 4, 8, [S.Bigrams, S.BigEnt], 1),       # 16 KB rows, photo content
    "c1bg2k": (2048, 8192, 4, 8, [S.Bigrams, S.BigEnt], 1),     # 8 KB rows, photo content
    "c5rand": (1024, 1024, 4, 8, [S.NONE, S.Bigrams, S.BigEnt], a.images),  # noise: every window is an outlier
    "c5n6": (1024, 1024, 4, 8, [S.NONE, S.Bigrams, S.BigEnt], a.images),
    "c5n12": (1024, 1024, 4, 8, [S.NONE, S.Bigrams, S.BigEnt], a.images),
    "c5n24": (1024, 1024, 4, 8, [S.NONE, S.Bigrams, S.BigEnt], a.images),
    "o2": (1024, 1024, 4, 8, [S.NONE, S.SUB, S.Entropy, S.Bigrams], a.images),
    "alpha": (1024, 1024, 4, 8, [S.MinSum], a.images),          # optimize_alpha on, 10 % of the 16x16 blocks transparent
    "alphao2": (1024, 1024, 4, 8, [S.NONE, S.SUB, S.Entropy, S.Bigrams], a.images),
}[a.config]
w, h, ch, bd, strat, n = cfg
if n == 1:
    w //= a.scale
    h //= a.scale
ct = {3: ox.ColorType.RGB(), 4: ox.ColorType.RGBA()}[ch]
interlaced = a.config == "c3b"
hdr = ox.IhdrData(w, h, ct, ox.BitDepth(bd), interlaced)
per = w * h * ch * (bd // 8)
d = synth.photo_batch(n, w, h, ch, bd, 0xB2000001, dev)
if a.config == "c5rand":
    d = torch.randint(0, 256, d.shape, dtype=torch.uint8, device=dev)
if a.config.startswith("c5n"):  # photo + extra uniform noise in [-A, A] on the colour channels (A = 6, 12, 24 ...)
    A = int(a.config[3:])
    nz = torch.randint(-A, A + 1, d.shape, dtype=torch.int16, device=dev)
    nz[3::4] = 0
    d = ((d.to(torch.int16) + nz) & 255).to(torch.uint8)
alpha = a.config.startswith("alpha")
if alpha:
    g = torch.Generator(device=dev); g.manual_seed(9)
    v = d.view(n, h, w, 4)
    m = (torch.rand((n, h // 16, w // 16), device=dev, generator=g) < 0.10).repeat_interleave(16, 1).repeat_interleave(16, 2)
    rnd = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=dev, generator=g)
    v[..., :3] = torch.where(m[..., None], rnd, v[..., :3])
    v[..., 3] = torch.where(m, torch.zeros_like(v[..., 3]), v[..., 3])
if interlaced:
    # the same image in the Adam7 layout (oxi_interlace_image, interlace.rs:6-60): every pass is a subsampled picture
    prog = ox.PngImage(ox.IhdrData(w, h, ct, ox.BitDepth(bd), False), d.cpu().numpy())
    d = torch.from_numpy(ox.interlace_image(prog).data.copy()).to(dev)
    per = d.numel()
rows = ox.lib().oxi_num_scanlines(__import__("ctypes").byref(hdr._c()), per)
o = [torch.empty(n * (per + rows), dtype=torch.uint8, device=dev) for _ in strat]
u = [torch.empty(n * rows, dtype=torch.uint8, device=dev) for _ in strat]
st = torch.cuda.current_stream(dev)
ms = []
for i in range(a.iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    ox.filter_batch_device(0, st.cuda_stream, hdr, d.data_ptr(), per, n, strat, alpha, [t.data_ptr() for t in o],
                           [t.data_ptr() for t in u])
    e1.record(st)
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print(a.config, f"{w}x{h}x{ch}x{bd} n={n}", "ms:", [round(x, 3) for x in ms], "raw GB/s (best):",
      round(n * per / min(ms) / 1e6, 1))
