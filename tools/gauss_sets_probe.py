"""Streaming Gaussian: one weight set per launch vs per-image weight sets (same radius bucket)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from millipyde_b200 import capi, engine


def main():
    capi.initialize()
    L = capi.lib()
    shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (2160, 3840, 3)
    n = 64
    rng = np.random.default_rng(0)
    a = rng.random(shape, dtype=np.float32)
    for name, chain in (("one set", [("gaussian", 2.0)]), ("per-image sets", [("random_gaussian", 1.9, 2.0)])):
        ch = engine.Chain(chain, device=0)
        best = 1e9
        for it in range(4):
            imgs = [capi.DeviceImage(a) for _ in range(n)]
            L.mpdev_synchronize()
            t0 = time.perf_counter()
            ch.run(imgs)
            L.mpdev_synchronize()
            best = min(best, time.perf_counter() - t0)
            for d in imgs:
                d.close()
        nbytes = 2 * a.nbytes * n
        print(f"{name:16s} {shape}: {best / n * 1e6:8.2f} us/image  {nbytes / best / 1e9:8.1f} GB/s  launches {ch.last_launches}")


if __name__ == "__main__":
    main()
