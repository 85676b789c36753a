#!/usr/bin/env python
"""bench.py -- filter-search throughput of the B200 path on BASELINE.json's batch configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--images-per-gpu B] [--no-others]

Workload (config c5 of BASELINE.json): a batch of synthetic 1024x1024 RGBA8 images filtered with the -o4 strategy
set that stays on the GPU, {None, Bigrams, BigEnt} (src/options.rs:220-233; Brute stays on the host), sharded by
image across GPUs with no data-path collective.  Weak scaling: 1250 images per GPU, i.e. the named 10 000-image
batch at 8 GPUs.  A step = one fused 3-strategy pass over the rank's whole batch.

  value      raw-pixel GB/s, whole job, inputs resident in HBM (CUDA events on the launching stream, max over ranks)
  e2e        the same through the host-buffer C-ABI call oxi_filter_batch (pinned host memory, H2D + D2H inside)
  roofline   algorithmic bytes of the dominant kernel / its event-timed duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference   the CPU oracle (restated reference, "port": the Rust reference cannot be built
             here) on all host threads over a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_W = IMG_H = 1024
CHANNELS = 4
PER_IMAGE = IMG_W * IMG_H * CHANNELS      # N = 4 194 304 raw bytes
ROWS = IMG_H                              # S
SEED0 = 0xB2000005


# stdout carries exactly ONE JSON line: everything else written to fd 1 (NCCL's version banner, library chatter) is
# redirected to stderr, and the result line goes to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle on all host threads (bench.py may execute oracle/ only here)
# ---------------------------------------------------------------------------------------------------------------

def cpu_filter_throughput(images, threads: int):
    """images: list of numpy arrays (one 1024x1024 RGBA8 image each). Runs {None,Bigrams,BigEnt} on each with a
    thread pool (ctypes releases the GIL), mirroring rayon's fan-out over (image x strategy) (evaluate.rs:143).
    -> seconds"""
    import concurrent.futures as cf
    import oracle
    ih = (IMG_W, IMG_H, 6, 8, 0)
    tasks = [(d, k) for d in images for k in ((oracle.BASIC), (oracle.BIGRAMS), (oracle.BIGENT))]
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda t: oracle.filter_image(ih, t[0], t[1], 0), tasks))
    return time.perf_counter() - t0


def host_images(n: int):
    import torch
    from oxipng_b200 import synth
    return [synth.photo_image(IMG_W, IMG_H, CHANNELS, 8, SEED0 + i, "cpu").numpy() for i in range(n)]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.use_native()
    threads = os.cpu_count() or 1
    n = max(threads, 8)  # bounded sample: one image per host thread and step (x3 strategies)
    imgs = host_images(n)
    for _ in range(args.warmup):
        cpu_filter_throughput(imgs[:max(threads // 2, 1)], threads)
    times = [cpu_filter_throughput(imgs, threads) for _ in range(args.steps)]
    sec = sum(times) / len(times)
    gbs = n * PER_IMAGE / sec / 1e9
    line = {
        "impl": "reference", "metric": "filter_search_raw_pixel_throughput", "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, n),
        "images_per_s": n / sec,
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": f"{n} images x {{None,Bigrams,BigEnt}} per step, {threads} threads, "
                                   "C restatement of the reference (oracle/), gcc -O3 -march=native"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, images_this_step=None):
    return {"workload": "c5: batch of 1024x1024 RGBA8 synthetic images, strategies {None,Bigrams,BigEnt} "
                        "(-o4 set without Brute), sharded by image",
            "images_per_gpu": args.images_per_gpu, "image": "1024x1024 RGBA8 (4 194 304 B, 1024 scanlines)",
            "strategies": ["None", "Bigrams", "BigEnt"], "optimize_alpha": False,
            "l2": f"per-GPU input {args.images_per_gpu * PER_IMAGE / 1e9:.2f} GB and output "
                  f"{3 * args.images_per_gpu * (PER_IMAGE + ROWS) / 1e9:.2f} GB per step vs 126 MB L2: no flush needed",
            "images_in_step": images_this_step}


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# ---------------------------------------------------------------------------------------------------------------

class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            log("clock sampler unavailable:", e)
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self._stop.clear()
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
            self._t = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------

def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def time_device(fn, steps, warmup, stream, dist, torch):
    """-> (total ms over `steps`, [per-launch ms]) timed with CUDA events on `stream`, barrier + sync on both sides"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record(stream)
    for i in range(steps):
        fn()
        evs[i + 1].record(stream)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return evs[0].elapsed_time(evs[steps]), per


def run_ours(args):
    import numpy as np
    import torch
    import oxipng_b200 as ox
    from oxipng_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1 and hasattr(os, "sched_setaffinity"):
        # one disjoint set of host cores per rank: the e2e leg's copy-issuing thread (and the pageable path's copy threads)
        # of different ranks do not migrate onto each other's cores
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(len(cores) // world, 1)
            mine = cores[local * per:(local + 1) * per] or cores
            os.sched_setaffinity(0, mine)
        except OSError:
            pass
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=dev)
        dist = dist_mod
    L = ox.lib()
    from oxipng_b200._lib import check
    check(L.oxi_set_device(local), "oxi_set_device")

    B = args.images_per_gpu
    hdr = ox.IhdrData(IMG_W, IMG_H, ox.ColorType.RGBA(), ox.BitDepth.Eight, False)
    strategies = [ox.FilterStrategy.NONE, ox.FilterStrategy.Bigrams, ox.FilterStrategy.BigEnt]
    K = len(strategies)
    out_per = PER_IMAGE + ROWS
    log(f"[rank {rank}] generating {B} images on {dev} ...")
    t0 = time.perf_counter()
    d_in = synth.photo_batch(B, IMG_W, IMG_H, CHANNELS, 8, SEED0 + rank * B, dev)
    d_out = [torch.empty(B * out_per, dtype=torch.uint8, device=dev) for _ in range(K)]
    d_used = [torch.empty(B * ROWS, dtype=torch.uint8, device=dev) for _ in range(K)]
    torch.cuda.synchronize()
    log(f"[rank {rank}] generated in {time.perf_counter() - t0:.1f}s")
    stream = torch.cuda.current_stream(dev)

    def step():
        ox.filter_batch_device(local, stream.cuda_stream, hdr, d_in.data_ptr(), PER_IMAGE, B, strategies, False,
                               [t.data_ptr() for t in d_out], [t.data_ptr() for t in d_used])

    # --- correctness spot check against the oracle (checker only) on rank 0: first and last image
    if rank == 0 and not args.no_check:
        import oracle
        step()
        torch.cuda.synchronize()
        for i in (0, B - 1):
            img = d_in[i * PER_IMAGE:(i + 1) * PER_IMAGE].cpu().numpy()
            for j, kind in enumerate((oracle.BASIC, oracle.BIGRAMS, oracle.BIGENT)):
                want, used, _ = oracle.filter_image((IMG_W, IMG_H, 6, 8, 0), img, kind, 0)
                got = d_out[j][i * out_per:(i + 1) * out_per].cpu().numpy()
                gu = d_used[j][i * ROWS:(i + 1) * ROWS].cpu().numpy()
                if not (np.array_equal(got, want) and np.array_equal(gu, used)):
                    raise SystemExit(f"parity check failed: image {i} strategy {j}")
        log("[rank 0] parity spot check vs oracle: ok")

    # --- device-resident timing (value, roofline)
    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        step()
    launches0 = L.oxi_kernel_launches()
    sampler.start()
    total_ms, per_launch = time_device(step, args.steps, 0, stream, dist, torch)
    sampler.stop()
    launches = L.oxi_kernel_launches() - launches0  # kernels launched inside the timed region
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    images_total = B * world
    value = images_total * PER_IMAGE / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant (only) kernel: k_fast<D=1, SC_BIGRAM>, one launch per step
    alg_bytes = B * (PER_IMAGE + K * (PER_IMAGE + 2 * ROWS))  # N + K(N+2S) per image, SURVEY 8d
    avg_launch_ms = sum(per_launch) / len(per_launch)
    peak, peak_src = peaks()
    achieved = alg_bytes / (avg_launch_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("c5_dram_bytes_per_image") * B  # per launch, scaled from the 64-image capture
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "k_fast<D=1,SC_BIGRAM|SC_HALF> (fused None+Bigrams+BigEnt; 256-thread CTAs, 4 per SM); per step also "
                          "k_spread_probe (classes the images by residual spread) and the 64x64-table variant, which "
                          "finds no image of its class in this workload",
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": avg_launch_ms, "peak_source": peak_src,
                "limiter": "instruction issue + shared-memory atomics (5 bigram-table updates per input byte), not HBM"}

    # --- end to end through the host-buffer C-ABI call (pinned host memory; H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        # e2e throughput is set by PCIe, not by the batch length: a 256-image slice of the rank's batch keeps the pinned
        # host footprint at 4.3 GB per rank (8 ranks on one host would otherwise pin 8 x 21 GB)
        Be = min(B, args.e2e_images)
        h_in = torch.empty(Be * PER_IMAGE, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(d_in[:Be * PER_IMAGE])
        h_out = [torch.empty(Be * out_per, dtype=torch.uint8, pin_memory=True) for _ in range(K)]
        h_used = [torch.empty(Be * ROWS, dtype=torch.uint8, pin_memory=True) for _ in range(K)]
        np_in = h_in.numpy()
        np_out = [t.numpy() for t in h_out]
        np_used = [t.numpy() for t in h_used]

        def e2e_step():
            ox.filter_batch(hdr, np_in, Be, strategies, False, 0, np_out, np_used)

        for _ in range(max(args.warmup, 1)):
            e2e_step()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        sampler.start()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        sampler.stop()
        te = torch.tensor([el], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        el = float(te.item())
        e2e = {"value": Be * world * PER_IMAGE * args.steps / el / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": Be * PER_IMAGE, "d2h_bytes_per_step": Be * K * (out_per + ROWS),
               "ms_per_step": el / args.steps * 1e3, "images_per_s": Be * world * args.steps / el,
               "images_per_gpu_per_step": Be,
               "api": "oxi_filter_batch (host buffers, pinned; chunked H2D/kernel/D2H pipeline)"}
        del h_in, h_out, h_used

    # --- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        oracle.use_native()
        threads = os.cpu_count() or 1
        n = max(threads, 8)
        imgs = [d_in[i * PER_IMAGE:(i + 1) * PER_IMAGE].cpu().numpy() for i in range(n)]
        cpu_filter_throughput(imgs[:max(threads // 2, 1)], threads)
        reps, t_acc = 0, 0.0
        while t_acc < 10.0 and reps < 8:
            t_acc += cpu_filter_throughput(imgs, threads)
            reps += 1
        cpu = {"value": n * reps * PER_IMAGE / t_acc / 1e9, "unit": "GB/s", "cores": threads, "kind": "port",
               "sample": f"{n} images x {{None,Bigrams,BigEnt}} x {reps} reps ({t_acc:.1f}s), {threads} threads, "
                         "C restatement of the reference (oracle/), gcc -O3 -march=native; "
                         "the Rust reference cannot be built in this image"}

    # --- every other BASELINE configuration, driver-visible (rank 0; a few seconds of GPU time)
    others = None
    e2e_pageable = None
    if not args.no_others and rank == 0 and world == 1:
        import bench_extra as bx
        peak_gbs, _ = peaks()
        it = max(args.steps, 5)
        others = {}
        for part, fn in (("single", lambda: bx.single_image_configs(ox, synth, torch, dev, local, stream, peak_gbs, it)),
                         ("batch", lambda: bx.batch_configs(ox, synth, torch, dev, local, stream, peak_gbs, it)),
                         ("unfilter", lambda: bx.unfilter_configs(ox, synth, torch, np, dev, local, stream, peak_gbs, it)),
                         ("c4", lambda: {"c4_reductions_2048_256colours": bx.c4_reductions(ox, torch, dev, local, stream, peak_gbs)}),
                         ("o2", lambda: {"o2_flow_resident_2048_256colours": bx.o2_flow(ox, torch, dev, local, stream)})):
            try:
                t0 = time.perf_counter()
                others.update(fn())
                log(f"[rank 0] other_configs/{part}: {time.perf_counter() - t0:.1f}s")
            except Exception as e:  # a failing side record must not hide the headline line
                others[f"{part}_error"] = repr(e)[:300]
                log(f"[rank 0] other_configs/{part} failed: {e!r}")
        if not args.no_e2e:
            try:
                e2e_pageable = bx.pageable_e2e(ox, torch, np, hdr, d_in, min(B, args.e2e_images), strategies, PER_IMAGE, ROWS,
                                               max(2, args.steps // 3))
            except Exception as e:
                e2e_pageable = {"error": repr(e)[:300]}

    # --- multi-GPU side records: what the host fabric gives the e2e path, and one huge image split into row bands
    ceiling = None
    band = None
    if not args.no_others and not args.no_e2e:
        import bench_extra as bx
        try:
            Be = min(B, args.e2e_images)
            hp_in = torch.empty(Be * PER_IMAGE, dtype=torch.uint8, pin_memory=True)
            hp_outs = [torch.empty(Be * (out_per + ROWS), dtype=torch.uint8, pin_memory=True) for _ in range(K)]
            ceiling = bx.copy_ceiling(torch, dist, dev, hp_in, hp_outs, max(2, args.steps // 2))
            ceiling["aggregate_GBps"] = ceiling["per_rank_GBps"] * world
            ceiling["raw_pixel_GBps_if_copy_bound"] = Be * world * PER_IMAGE / ceiling["seconds_per_step"] / 1e9
            if e2e is not None:
                ceiling["e2e_fraction_of_copy_ceiling"] = e2e["value"] / ceiling["raw_pixel_GBps_if_copy_bound"]
            del hp_in, hp_outs
        except Exception as e:
            ceiling = {"error": repr(e)[:300]}
        if dist is not None:
            dist.barrier()
        if rank == 0:
            try:
                n_dev = min(world, torch.cuda.device_count())
                band = bx.row_band(ox, torch, np, n_dev)
            except Exception as e:
                band = {"error": repr(e)[:300]}

    if rank == 0:
        line = {
            "metric": "filter_search_raw_pixel_throughput", "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, images_total),
            "images_per_s": images_total / (ms_per_step * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        if others is not None:
            line["other_configs"] = others
        if e2e_pageable is not None:
            line["e2e_pageable"] = e2e_pageable
        if ceiling is not None:
            line["copy_ceiling"] = ceiling
        if band is not None:
            line["row_band"] = band
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images-per-gpu", type=int, default=1250)
    ap.add_argument("--others", action="store_true", help="(default now) also time configs c1-c4, the -o2 set and flow")
    ap.add_argument("--no-others", action="store_true", help="headline c5 line only")
    ap.add_argument("--e2e-images", type=int, default=256, help="images per rank and step in the end-to-end leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
