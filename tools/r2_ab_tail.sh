#!/bin/bash
# A/B of the streaming Gaussian's work-item plan on one box: whole waves + a short tail (default)
# against every column cut the same way (MILLIPYDE_GAUSS_TAIL=0, the round-1 schedule).
set -u
mkdir -p gpurun_out
echo "== gaussian tests"
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -k "gauss or full_frame or fused" 2>&1 | tail -5
for rep in 1 2; do
  for tail in 0 1; do
    MILLIPYDE_GAUSS_TAIL=$tail timeout 300 python bench.py --steps 40 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_tail_${tail}_${rep}.json 2> gpurun_out/ab_tail_${tail}_${rep}.err
    python - <<PY
import json
d=json.load(open("gpurun_out/ab_tail_${tail}_${rep}.json"))
print("tail=${tail} rep=${rep}", round(d["value"],1), "img/s frac", round(d["roofline"]["frac"],4), "sm_mhz", d["clocks"]["sm_mhz"], "parity", d["parity_check"]["max_abs_err"])
PY
  done
done
