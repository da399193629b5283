#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu_scale.json 2> gpurun_out/r2_bench_${N}gpu_scale.err; echo "rc=$?"; tail -c 400 gpurun_out/r2_bench_${N}gpu_scale.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-300
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_${N}gpu_scale.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"], 1), "clk", d["clocks"]["sm_mhz"], d["clocks"]["power_w_median"],
      "in_process", round(d["in_process"].get("value", 0)), d.get("multi_device_check"), d["in_process"].get("multi_device_check"))
PY
