"""RGBA8 reference-rule Gaussian on 4K images (for ncu / timing)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from millipyde_b200 import capi


def main():
    capi.initialize()
    L = capi.lib()
    L.mpimg_set_semantics(capi.SEMANTICS_REFERENCE)
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (2160, 3840, 4), dtype=np.uint8)
    imgs = [capi.DeviceImage(a) for _ in range(8)]
    for rep in range(3):
        L.mpdev_synchronize()
        t0 = time.perf_counter()
        for d in imgs:
            d.apply_chain([("gaussian", 2.0)])
        L.mpdev_synchronize()
        dt = time.perf_counter() - t0
    print(f"rgba8 4K gaussian: {dt / len(imgs) * 1e6:.1f} us/image (wall, eager)")


if __name__ == "__main__":
    main()
