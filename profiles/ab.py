"""A/B timing of library builds on one GPU visit:  python profiles/ab.py <tag>=<lib.so> ... [--configs c5,c1,...]

Every build runs the same device-resident configurations in its own process (OXIPNG_B200_LIB selects the library);
per config the best-of-N kernel time and a checksum of all outputs are printed, so that variants can be compared for
speed and for bit-equality of their results (the parity tests proper are tests/ -m gpu).
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

CHILD = r'''
import ctypes, hashlib, json, os, sys
sys.path.insert(0, ROOT)
import torch
import oxipng_b200 as ox
from oxipng_b200 import synth
dev = torch.device("cuda:0")
S = ox.FilterStrategy
cfgs = {
    "c5": (1024, 1024, 4, 8, [S.NONE, S.Bigrams, S.BigEnt], 256, False, 0),
    "c5n6": (1024, 1024, 4, 8, [S.NONE, S.Bigrams, S.BigEnt], 64, False, 6),
    "c5n12": (1024, 1024, 4, 8, [S.NONE, S.Bigrams, S.BigEnt], 64, False, 12),
    "c5rand": (1024, 1024, 4, 8, [S.NONE, S.Bigrams, S.BigEnt], 16, False, -1),
    "o2": (1024, 1024, 4, 8, [S.NONE, S.SUB, S.Entropy, S.Bigrams], 256, False, 0),
    "c1": (4096, 4096, 4, 8, [S.MinSum], 1, False, 0),
    "c2": (8192, 8192, 3, 8, [S.Entropy], 1, False, 0),
    "c2paeth": (8192, 8192, 3, 8, [S.PAETH], 1, False, 0),
    "c3a": (4096, 4096, 4, 16, [S.Bigrams, S.BigEnt], 1, False, 0),
    "c3b": (4096, 4096, 4, 8, [S.Bigrams, S.BigEnt], 1, True, 0),
    "c1bg": (4096, 4096, 4, 8, [S.Bigrams, S.BigEnt], 1, False, 0),
    "rgb8bg": (1365, 1024, 3, 8, [S.Bigrams, S.BigEnt], 64, False, 0),
    # optimize_alpha on: 10 % of the 16x16 blocks fully transparent with random colour behind
    "alpha_ms": (1024, 1024, 4, 8, [S.MinSum], 128, False, -2),
    "alpha_o2": (1024, 1024, 4, 8, [S.NONE, S.SUB, S.Entropy, S.Bigrams], 128, False, -2),
}
out = {}
for name in NAMES:
    w, h, ch, bd, strat, n, interlaced, noise = cfgs[name]
    ct = {3: ox.ColorType.RGB(), 4: ox.ColorType.RGBA()}[ch]
    hdr = ox.IhdrData(w, h, ct, ox.BitDepth(bd), interlaced)
    per = w * h * ch * (bd // 8)
    d = synth.photo_batch(n, w, h, ch, bd, 0xB2000001, dev)
    alpha = noise == -2
    if alpha:
        g = torch.Generator(device=dev); g.manual_seed(9)
        v = d.view(n, h, w, 4)
        m = (torch.rand((n, h // 16, w // 16), device=dev, generator=g) < 0.10).repeat_interleave(16, 1).repeat_interleave(16, 2)
        rnd = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=dev, generator=g)
        v[..., :3] = torch.where(m[..., None], rnd, v[..., :3])
        v[..., 3] = torch.where(m, torch.zeros_like(v[..., 3]), v[..., 3])
    elif noise < 0:
        g = torch.Generator(device=dev); g.manual_seed(7)
        d = torch.randint(0, 256, d.shape, dtype=torch.uint8, device=dev, generator=g)
    elif noise > 0:
        g = torch.Generator(device=dev); g.manual_seed(7)
        nz = torch.randint(-noise, noise + 1, d.shape, dtype=torch.int16, device=dev, generator=g)
        nz[3::4] = 0
        d = ((d.to(torch.int16) + nz) & 255).to(torch.uint8)
    if interlaced:
        prog = ox.PngImage(ox.IhdrData(w, h, ct, ox.BitDepth(bd), False), d.cpu().numpy())
        d = torch.from_numpy(ox.interlace_image(prog).data.copy()).to(dev)
        per = d.numel()
    rows = ox.lib().oxi_num_scanlines(ctypes.byref(hdr._c()), per)
    o = [torch.zeros(n * (per + rows), dtype=torch.uint8, device=dev) for _ in strat]
    u = [torch.zeros(n * rows, dtype=torch.uint8, device=dev) for _ in strat]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev)
    ms = []
    for i in range(ITERS):
        flush.fill_(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ox.filter_batch_device(0, st.cuda_stream, hdr, d.data_ptr(), per, n, strat, alpha, [t.data_ptr() for t in o],
                               [t.data_ptr() for t in u])
        e1.record(st)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    hsh = hashlib.sha1()
    for t in o + u:
        hsh.update(t.cpu().numpy().tobytes())
    out[name] = {"ms": round(min(ms[1:]), 4), "raw_GBps": round(n * per / min(ms[1:]) / 1e6, 1), "sha": hsh.hexdigest()[:12]}
    del o, u, d, flush
    torch.cuda.empty_cache()
print("ABRESULT " + json.dumps(out))
'''


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    names = "c5,c5n6,c5n12,c5rand,o2,c1,c2,c3a,c3b,c1bg"
    iters = 6
    for a in sys.argv[1:]:
        if a.startswith("--configs="):
            names = a.split("=", 1)[1]
        if a.startswith("--iters="):
            iters = int(a.split("=", 1)[1])
    res = {}
    for rep in range(2):  # two interleaved rounds: guards against drift between the first and the last build
        for a in args:
            tag, lib = a.split("=", 1)
            env = dict(os.environ, OXIPNG_B200_LIB=os.path.abspath(lib))
            code = f"ROOT={ROOT!r}\nNAMES={names.split(',')!r}\nITERS={iters}\n" + CHILD
            r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("ABRESULT ")]
            if not line:
                print(tag, "FAILED", r.stderr[-2000:])
                continue
            cur = json.loads(line[0][9:])
            if tag in res:
                for k, v in cur.items():
                    if v["ms"] < res[tag][k]["ms"]:
                        res[tag][k] = v
            else:
                res[tag] = cur
    tags = list(res)
    for cfg in names.split(","):
        row = [f"{cfg:8s}"]
        for t in tags:
            v = res[t].get(cfg)
            same = "" if not v or t == tags[0] or tags[0] not in res else ("=" if v["sha"] == res[tags[0]][cfg]["sha"] else "DIFF")
            row.append(f"{t}: {v['ms']:9.4f} ms {v['raw_GBps']:8.1f} GB/s {same:4s}" if v else f"{t}: -")
        print("  ".join(row))
    print("ABJSON " + json.dumps(res))


if __name__ == "__main__":
    main()
