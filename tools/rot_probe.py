import sys, numpy as np
sys.path.insert(0, '/root/repo')
from millipyde_b200 import capi
capi.initialize()
img = np.random.default_rng(0).random((2160, 3840, 3), dtype=np.float32)
for _ in range(3):
    d = capi.DeviceImage(img); d.apply("rotate", 30.0); d.sync(); d.close()
