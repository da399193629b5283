"""Pipeline.run() without a device on every visible GPU, repeated: the images start on device 0, the
shards of the other devices move theirs over and must not launch before every move has landed
(the hand-over race fixed in mp_pipeline.cu: shard_worker).  usage: python tools/cycling_stress.py [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from millipyde_b200 import capi, engine
from oracle import skimage_oracle as so


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    ndev = capi.initialize()
    L = capi.lib()
    rng = np.random.default_rng(77)
    n = 16 * ndev + 3
    imgs = [rng.random((64, 640, 3), dtype=np.float32) for _ in range(n)]
    chain = [("gaussian", 2.0), ("fliplr",)]
    want = [so.apply_chain(a, chain) for a in imgs]
    worst, bad = 0.0, 0
    for rep in range(reps):
        L.mpdev_set_target_device(0)
        dev = [capi.DeviceImage(a) for a in imgs]
        L.mpdev_set_target_device(capi.DEVICE_LOC_NO_AFFINITY)
        engine.Chain(chain).run(dev)
        used = sorted({d.device for d in dev})
        for d, w in zip(dev, want):
            e = float(np.abs(d.numpy() - w).max())
            worst = max(worst, e)
            bad += e > 1e-5
            d.close()
    print({"devices": ndev, "devices_used": used, "images": n, "reps": reps, "max_abs_err": worst, "wrong_images": int(bad),
           "ok": bad == 0 and used == list(range(ndev))})


if __name__ == "__main__":
    main()
