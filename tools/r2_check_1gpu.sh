#!/bin/bash
# Round 2, 1-GPU visit: gpu tests (incl. the fused pointwise+Gaussian segment), bench with per-step e2e times.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== bench (default)"
timeout 600 python bench.py --e2e-steps 12 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "rc=$?"
tail -c 400 gpurun_out/r2_bench_a.err
echo "== bench small resident leg, e2e 12 steps"
timeout 600 python bench.py --batch 8 --steps 3 --no-cpu --e2e-steps 12 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_a.json", "gpurun_out/r2_bench_b.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["roofline"]["frac"], 4), d["e2e"]["value"], d["e2e"]["step_s"], d["clocks"]["sm_mhz"], d["parity_check"]["max_abs_err"])
    except Exception as e:
        print(f, "unreadable", e)
PY
