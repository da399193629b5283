cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests -m gpu -q -x -k "rotate or gather or config3 or chain or fusion or random or config5 or transpose" 2>&1 | tail -2
timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gather_f32 --csv --log-file gpurun_out/gather1_t.csv python - <<'PY' 2>&1 | tail -1
import sys; sys.path.insert(0, ".")
import numpy as np
from millipyde_b200 import capi, engine
capi.initialize()
rng = np.random.default_rng(0)
for shape, n in [((1024, 1024), 32), ((1920, 1080), 64)]:
    seed = capi.DeviceImage(rng.random(shape, dtype=np.float32))
    batch = [seed.clone() for _ in range(n)]
    ch = engine.Chain([("rotate", 30.0)], device=0)
    ch.run(batch)
print("ok")
PY
grep gather gpurun_out/gather1_t.csv | awk -F'","' '{print $5, $9, $NF}'
