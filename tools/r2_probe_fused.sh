#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_python_api.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/fused_gauss_probe.py 128 > gpurun_out/r2_fused_gauss.json 2> gpurun_out/r2_fused_gauss.err; cat gpurun_out/r2_fused_gauss.json; tail -c 300 gpurun_out/r2_fused_gauss.err
PROBE_ONLY=gaussian_colorize timeout 600 ncu --set full --clock-control none --import-source on -k regex:gauss_stream -s 2 -c 1 -o gpurun_out/r2_fused_colorize python tools/fused_gauss_probe.py 32 > gpurun_out/r2_ncu_fused.log 2>&1; tail -3 gpurun_out/r2_ncu_fused.log
