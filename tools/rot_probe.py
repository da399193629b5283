"""Drive the fp32 rotate (gather_f32_kernel) for an ncu capture: rot_probe.py [rgb|grey]."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from millipyde_b200 import capi, engine

capi.initialize()
grey = len(sys.argv) > 1 and sys.argv[1] == "grey"
shape = (1024, 1024) if grey else (1080, 1920, 3)
img = np.random.default_rng(0).random(shape, dtype=np.float32)
seed = capi.DeviceImage(img)
for _ in range(3):
    batch = [seed.clone() for _ in range(32)]
    ch = engine.Chain([("rotate", 30.0)], device=0)
    ch.run(batch)
    ch.close()
