import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from millipyde_b200 import capi, engine
capi.initialize()
rng = np.random.default_rng(1)
imgs = [capi.DeviceImage(rng.random((1080, 1920, 3), dtype=np.float32)) for _ in range(16)]
ch = engine.Chain([("rotate", 30.0), ("fliplr",), ("adjust_gamma", 1.5, 1.0), ("gaussian", 2.0)], device=0)
for _ in range(3):
    ch.run(imgs)
capi.lib().mpdev_synchronize()
