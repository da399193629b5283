#!/bin/bash
# Same-box A/B of two builds of the library: bench.py of this tree against bench.py of a copy of the
# tree under ab_old/ (built from another revision of a kernel), interleaved, CPU / e2e legs off.
set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -k "gauss or full_frame or fused" 2>&1 | tail -2 )
for rep in 1 2 3; do
  for arm in new old; do
    if [ $arm = old ]; then d=ab_old; else d=.; fi
    ( cd $d && timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu --no-e2e ) > gpurun_out/ab_$arm.json 2> gpurun_out/ab_$arm.err || tail -c 300 gpurun_out/ab_$arm.err
    python - <<PY
import json
d=json.load(open("gpurun_out/ab_$arm.json"))
print("$arm rep=$rep", round(d["value"],1), "img/s frac", round(d["roofline"]["frac"],4), "sm_mhz", d["clocks"]["sm_mhz"], "ok", d["parity_check"]["ok"])
PY
  done
done
