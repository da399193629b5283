"""Generator(return_to_host=True) throughput (page-locked, batched downloads) next to the plain
blocking download of the same images with np.array(img) (pageable destination, one at a time)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import millipyde_b200

mp = millipyde_b200.load_extension()
rng = np.random.default_rng(0)
base = [mp.gpuimage(rng.random((1024, 1024, 3), dtype=np.float32)) for _ in range(6)]
g = mp.Generator(base, [mp.Operation("fliplr")], outputs=2048 + 64, prefetch=64, device=0, return_to_host=True)
for _ in range(64):
    next(g)
t0 = time.perf_counter()
n = 0
for o in g:
    n += 1
dt = time.perf_counter() - t0
res = {"generator_to_host_images_per_s": n / dt, "generator_to_host_GBs": n * o.nbytes / dt / 1e9}
t0 = time.perf_counter()
for k in range(256):
    a = np.array(base[k % 6])
dt = time.perf_counter() - t0
res["np_array_images_per_s"] = 256 / dt
res["np_array_GBs"] = 256 * a.nbytes / dt / 1e9
print(json.dumps(res))
