#!/bin/bash
# Round 2, 1-GPU visit: gpu tests, fused pointwise+Gaussian vs bare Gaussian, config 3.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== fused gaussian probe"
timeout 300 python tools/fused_gauss_probe.py 128 > gpurun_out/r2_fused_gauss.json 2> gpurun_out/r2_fused_gauss.err; echo "rc=$?"; cat gpurun_out/r2_fused_gauss.json; tail -c 300 gpurun_out/r2_fused_gauss.err
echo "== config3"
timeout 300 python tools/bench_configs.py config3 > gpurun_out/r2_config3.json 2> gpurun_out/r2_config3.err; echo "rc=$?"; cat gpurun_out/r2_config3.json; tail -c 300 gpurun_out/r2_config3.err
