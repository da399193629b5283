#!/usr/bin/env python
"""Per-operator and per-config device throughput (not the headline bench: bench.py is).

Measures every fp32 / RGBA8 operator as algorithmic GB/s (input read once + output
written once) on device-resident images with CUDA events, then
BASELINE configs 3 (fused chain, 1080p batch) and 4 (Generator random-augmentation
stream).  Prints one JSON object; used to fill DESIGN.md section 7 and profiles/.

    python tools/bench_ops.py [--quick]
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from millipyde_b200 import capi, engine  # noqa: E402

_PEAKS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
PEAK = json.load(open(_PEAKS))["hbm_gbs"] if os.path.exists(_PEAKS) else 6650.0
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def timed(L, stream, fn, reps):
    e0, e1 = L.mpdev_event_create(0), L.mpdev_event_create(0)
    fn()
    L.mpdev_synchronize_all()
    L.mpdev_event_record(e0, stream)
    for _ in range(reps):
        fn()
    L.mpdev_event_record(e1, stream)
    ms = L.mpdev_event_elapsed_ms(e0, e1) / reps
    L.mpdev_event_destroy(e0)
    L.mpdev_event_destroy(e1)
    return ms


def op_table(L, quick):
    """Each op as a one-stage chain over a batch (launches issued from C++ back to back, so the
    event time is kernel time, not Python call overhead)."""
    out = {}
    rng = np.random.default_rng(0)
    s1 = L.mpdev_get_stream(0, 1)
    cases = [("f32 4K RGB", rng.random((2160, 3840, 3), dtype=np.float32), 12),
             ("f32 1080p RGB", rng.random((1080, 1920, 3), dtype=np.float32), 48),
             ("rgba8 4K", rng.integers(0, 256, (2160, 3840, 4), dtype=np.uint8), 24)]
    ops = [("rgb2grey",), ("transpose",), ("fliplr",), ("rotate", 30.0), ("brightness", 0.1),
           ("adjust_gamma", 1.5, 1.0), ("colorize", 0.9, 1.1, 1.0), ("gaussian", 2.0)]
    for name, img, n in cases:
        seed = capi.DeviceImage(img)
        for op in ops:
            batch = [seed.clone() for _ in range(n)]
            ch = engine.Chain([op], device=0)
            in_b = img.nbytes
            if op[0] == "rgb2grey":
                out_b = img.shape[0] * img.shape[1] * (4 if img.dtype == np.float32 else 8)
                warm = [seed.clone() for _ in range(2)]
                ch.run(warm)
                L.mpdev_synchronize_all()
                e0, e1 = L.mpdev_event_create(0), L.mpdev_event_create(0)
                L.mpdev_event_record(e0, s1)
                ch.run(batch)
                L.mpdev_event_record(e1, s1)
                ms = L.mpdev_event_elapsed_ms(e0, e1) / n
                for d in warm:
                    d.close()
            else:
                out_b = in_b
                reps = 4 if quick else 10
                if op[0] == "transpose":
                    reps += reps % 2
                ms = timed(L, s1, lambda: ch.run(batch), reps) / n
            gbs = (in_b + out_b) / ms / 1e6
            out[f"{name} {op[0]}"] = {"us_per_image": round(ms * 1e3, 2), "GB/s": round(gbs, 1),
                                     "frac": round(gbs / PEAK, 3)}
            ch.close()
            for d in batch:
                d.close()
        seed.close()
    return out


def config3(L, quick):
    """rotate 30 -> fliplr -> adjust_gamma 1.5 -> gaussian 2 on a batch of 1920x1080 RGB fp32."""
    n = 256 if quick else 1024
    chain = [("rotate", 30.0), ("fliplr",), ("adjust_gamma", 1.5, 1.0), ("gaussian", 2.0)]
    rng = np.random.default_rng(3000)
    seeds = [capi.DeviceImage(rng.random((1080, 1920, 3), dtype=np.float32)) for _ in range(4)]
    imgs = [seeds[k % 4].clone() for k in range(n)]
    res = {}
    for fused in (1, 0):
        L.mppipe_set_fusion(fused)
        ch = engine.Chain(chain, device=0)
        ch.run(imgs)
        t0 = time.perf_counter()
        s1 = L.mpdev_get_stream(0, 1)
        ms = timed(L, s1, lambda: ch.run(imgs), 3)
        wall = (time.perf_counter() - t0) / 4
        bytes_img = 2 * 1080 * 1920 * 3 * 4
        res["fused" if fused else "unfused"] = {
            "images/s": round(n / (ms / 1e3), 1), "ms_per_batch": round(ms, 3), "wall_ms_per_batch": round(wall * 1e3, 3),
            "launches_per_batch": int(ch.last_launches), "algorithmic GB/s": round(n * bytes_img / ms / 1e6, 1),
            "frac": round(n * bytes_img / ms / 1e6 / PEAK, 3)}
        ch.close()
    L.mppipe_set_fusion(1)
    for d in imgs + seeds:
        d.close()
    return {"batch": n, **res}


def config4(quick):
    """Generator random-augmentation stream (examples/augmentation_examples.py:13-21) on 1024x1024 RGB fp32."""
    import millipyde_b200
    mp = millipyde_b200.load_extension()
    rng = np.random.default_rng(4000)
    base = [mp.gpuimage(rng.random((1024, 1024, 3), dtype=np.float32)) for _ in range(6)]
    ops = [mp.Operation("transpose", probability=.2), mp.Operation("fliplr", probability=.2),
           mp.Operation("random_brightness", -.2, .2), mp.Operation("random_gaussian", .5, 2.),
           mp.Operation("random_colorize", [.5, 1.5], [.5, 1.5], [.5, 1.5], probability=.3),
           mp.Operation("rgb2grey", probability=.3), mp.Operation("random_rotate", 0., 120., probability=.5)]
    n = 512 if quick else 4096
    mp.seed(4)
    res = {}
    for prefetch in (1, 64):
        g = mp.Generator(base, ops, outputs=n, prefetch=prefetch)
        next(g)
        mp.synchronize()
        t0 = time.perf_counter()
        k = 1
        for _ in g:
            k += 1
        mp.synchronize()
        dt = time.perf_counter() - t0
        res[f"prefetch={prefetch}"] = {"images/s": round((k - 1) / dt, 1), "outputs": k}
    mp.seed(0)
    return res


def main():
    quick = "--quick" in sys.argv
    capi.initialize()
    L = capi.lib()
    out = {"hbm_peak_GB/s": PEAK, "ops": op_table(L, quick), "config3": config3(L, quick), "config4": config4(quick)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
