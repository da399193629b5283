"""north star (2), measured: a pointwise -> gaussian -> pointwise chain against the bare Gaussian on the same
resident batch of 3840x2160 RGB fp32 noise images (views: the inputs are never written).  One JSON line:
images/s of both, their ratio, launches per batch (must be 1 for the fused chain when every image shares
the programs: ceil(n / 64) launches of the per-image-record kernel otherwise)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from millipyde_b200 import capi, engine


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    only = os.environ.get("PROBE_ONLY")
    h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2160, 3840)
    capi.initialize()
    L = capi.lib()
    rng = np.random.default_rng(7)
    base = [capi.DeviceImage(rng.random((h, w, 3), dtype=np.float32)) for _ in range(2)]
    src = [base[k % 2].clone() for k in range(n)]
    views = [d.view() for d in src]
    out = {"images": n, "shape": [h, w, 3]}
    chains = {"gaussian": [("gaussian", 2.0)],
              # per-image weight sets, no pointwise program: what the *_sets kernel costs by itself
              "gaussian_per_image_sigma": [("random_gaussian", 2.0, 2.000001)],
              "gaussian_add0": [("gaussian", 2.0), ("brightness", 0.0)],
              "gamma_gaussian_brightness": [("adjust_gamma", 1.5, 1.0), ("gaussian", 2.0), ("brightness", 0.1)],
              "brightness_gaussian": [("brightness", 0.1), ("gaussian", 2.0)],
              "gaussian_colorize": [("gaussian", 2.0), ("colorize", 0.9, 1.1, 1.0)]}
    first = True
    for name, ops in chains.items():
        if only and name != only:
            continue
        ch = engine.Chain(ops, device=0)
        times = []
        for rep in range(6):
            if not first:
                for v, d in zip(views, src):
                    v.rebind(d)
            first = False
            L.mpdev_synchronize()
            t0 = time.perf_counter()
            ch.run_views(views)
            L.mpdev_synchronize()
            times.append(time.perf_counter() - t0)
        best = min(times[1:])
        out[name] = {"images/s": round(n / best, 1), "launches": int(ch.last_launches), "segments": int(ch.last_segments),
                     "GB/s_algorithmic": round(n * 2 * h * w * 12 / best / 1e9, 1)}
        ch.close()
    for k in [c for c in list(chains)[1:] if c in out and "gaussian" in out]:
        out[k]["vs_bare_gaussian"] = round(out[k]["images/s"] / out["gaussian"]["images/s"], 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
