#!/bin/bash
# A/B of the streaming Gaussian's blocked-wait settings on one box (MILLIPYDE_GAUSS_WAIT = "row_in,row_empty,col"
# nanoseconds between probes; s<ns> = suspended try_wait), interleaved REPS times.  bench.py without the CPU / e2e legs.
set -u
mkdir -p gpurun_out
REPS=${REPS:-2}
for rep in $(seq 1 $REPS); do
for cfg in "$@"; do
  MILLIPYDE_GAUSS_WAIT=$cfg timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_wait.json 2> gpurun_out/ab_wait.err || tail -c 300 gpurun_out/ab_wait.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_wait.json"))
print("rep $rep wait=$cfg", round(d["value"],1), "img/s frac", round(d["roofline"]["frac"],4), "sm_mhz", d["clocks"]["sm_mhz"], "ok", d["parity_check"]["ok"])
PY
done
done
