"""Drive the one-word-pixel transpose on the kernel table's two cases (for an ncu launch list)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from millipyde_b200 import capi, engine

capi.initialize()
rng = np.random.default_rng(0)
for shape, dt, n in [((1024, 1024), np.float32, 32), ((2160, 3840, 4), np.uint8, 8), ((1080, 1920), np.float32, 64),
                     ((2160, 3840), np.float64, 8),
                     ((2160, 3840, 3), np.float32, 8), ((1080, 1920, 4), np.float32, 16)]:
    img = rng.integers(0, 256, shape, dtype=np.uint8) if dt == np.uint8 else rng.random(shape).astype(dt)
    seed = capi.DeviceImage(img)
    batch = [seed.clone() for _ in range(n)]
    ch = engine.Chain([("transpose",)], device=0)
    ch.run(batch)
    want = np.transpose(img, (1, 0, 2)) if img.ndim == 3 else img.T
    assert np.array_equal(batch[-1].numpy(), want)
    ch.close()
print("ok")
