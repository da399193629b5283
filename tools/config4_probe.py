"""Config-4 probe: Generator random-augmentation stream on 1024x1024 RGB fp32 (one device).

Prints wall-clock images/s and launches per image; run it under ``ncu --metrics
gpu__time_duration.sum`` to get the device-time share per kernel (profiles/r1_config4_*.md).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import millipyde_b200
from millipyde_b200 import capi


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    prefetch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    mp = millipyde_b200.load_extension()
    rng = np.random.default_rng(4000)
    base = [mp.gpuimage(rng.random((1024, 1024, 3), dtype=np.float32)) for _ in range(6)]
    ops = [mp.Operation("transpose", probability=.2), mp.Operation("fliplr", probability=.2),
           mp.Operation("random_brightness", -.2, .2), mp.Operation("random_gaussian", .5, 2.),
           mp.Operation("random_colorize", [.5, 1.5], [.5, 1.5], [.5, 1.5], probability=.3),
           mp.Operation("rgb2grey", probability=.3), mp.Operation("random_rotate", 0., 120., probability=.5)]
    mp.seed(4)
    g = mp.Generator(base, ops, outputs=n + prefetch, prefetch=prefetch)
    for _ in range(prefetch):
        next(g)
    mp.synchronize()
    l0 = mp.launch_count()
    t0 = time.perf_counter()
    k = 0
    for _ in g:
        k += 1
    mp.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"outputs": k, "prefetch": prefetch, "images/s": round(k / dt, 1),
                      "us_per_image": round(dt / k * 1e6, 2),
                      "launches_per_image": round((mp.launch_count() - l0) / k, 2)}))


if __name__ == "__main__":
    main()
