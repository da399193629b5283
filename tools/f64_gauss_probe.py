"""fp64 greyscale Gaussian on 4K images, both semantics (wall time of eager launches)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from millipyde_b200 import capi


def main():
    capi.initialize()
    L = capi.lib()
    rng = np.random.default_rng(0)
    a = rng.random((2160, 3840))
    for name, mode in (("oracle (33 taps)", capi.SEMANTICS_ORACLE), ("reference (17 taps)", capi.SEMANTICS_REFERENCE)):
        L.mpimg_set_semantics(mode)
        imgs = [capi.DeviceImage(a) for _ in range(8)]
        for rep in range(3):
            L.mpdev_synchronize()
            t0 = time.perf_counter()
            for d in imgs:
                d.apply_chain([("gaussian", 2.0)])
            L.mpdev_synchronize()
            dt = time.perf_counter() - t0
        print(f"f64 4K grey gaussian, {name}: {dt / len(imgs) * 1e6:.1f} us/image")
        for d in imgs:
            d.close()
    L.mpimg_set_semantics(capi.SEMANTICS_ORACLE)


if __name__ == "__main__":
    main()
