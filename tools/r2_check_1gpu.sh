#!/bin/bash
# Round 2, 1-GPU visit: every gpu test, the fused-Gaussian probe, configs 3 and 4 on one device.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== fused gaussian probe"
timeout 300 python tools/fused_gauss_probe.py 128 > gpurun_out/r2_fused_gauss.json 2> gpurun_out/r2_fused_gauss.err; cat gpurun_out/r2_fused_gauss.json; tail -c 300 gpurun_out/r2_fused_gauss.err
echo "== config3 / config4 (one device)"
timeout 300 python tools/bench_configs.py config3 > gpurun_out/r2_config3.json 2> gpurun_out/r2_config3.err; cat gpurun_out/r2_config3.json; tail -c 300 gpurun_out/r2_config3.err
timeout 300 python tools/bench_configs.py config4 --ref-params --prefetch 256 > gpurun_out/r2_config4_1gpu.json 2> gpurun_out/r2_config4.err; cat gpurun_out/r2_config4_1gpu.json; tail -c 300 gpurun_out/r2_config4.err
timeout 300 python tools/bench_configs.py config4 --ref-params --prefetch 2048 >> gpurun_out/r2_config4_1gpu.json 2>> gpurun_out/r2_config4.err; tail -1 gpurun_out/r2_config4_1gpu.json
