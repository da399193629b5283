#!/usr/bin/env python
"""The reference's own kernels (unmodified sources, HIP->CUDA shim build in oracle/_ref) timed on
the same B200, on its native layouts (RGBA8, fp64 grey), through its own Python API -- the
"existing kernel on the same box" line of SURVEY.md 8d.  Then the same ops through this repo's
extension on the same inputs.  Wall-clock per op over a loop (both APIs block or are synced)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OPS = [("rgb2grey", ()), ("transpose", ()), ("fliplr", ()), ("rotate", (30.0,)), ("gaussian", (2.0,)),
       ("adjust_gamma", (2.0, 1.0)), ("brightness", (0.1,)), ("colorize", (0.9, 1.1, 1.0))]


def run(mp, sync, label, n=12):
    rng = np.random.default_rng(0)
    rgba = rng.integers(0, 256, (2160, 3840, 4), dtype=np.uint8)
    res = {}
    for layout in ("rgba8", "f64"):
        for name, args in OPS:
            if layout == "f64" and name in ("rgb2grey", "colorize"):
                continue
            imgs = [mp.gpuimage(rgba) for _ in range(n)]
            if layout == "f64":
                for im in imgs:
                    im.rgb2grey()
            sync()
            getattr(imgs[0], name)(*args)          # warm
            sync()
            t0 = time.perf_counter()
            for im in imgs[1:]:
                getattr(im, name)(*args)
            sync()
            dt = (time.perf_counter() - t0) / (n - 1)
            res[f"{layout} 4K {name}"] = round(dt * 1e6, 1)
            del imgs
    return {label: res}


if __name__ == "__main__":
    which = sys.argv[1]
    if which == "reference":
        from oracle import build_ref
        ref = build_ref.load()
        # the reference has no module-level sync; Device.__exit__ synchronises (src/device.c:67)
        def sync():
            with ref.Device(0):
                pass
        print(json.dumps(run(ref, sync, "reference kernels (us per 4K image)")))
    else:
        import millipyde_b200
        mp = millipyde_b200.load_extension()
        mp.set_semantics("reference")
        print(json.dumps(run(mp, mp.synchronize, "this repo, reference semantics (us per 4K image)")))
