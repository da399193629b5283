#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/fused_gauss_probe.py 128 > gpurun_out/r2_fused_gauss.json 2> gpurun_out/r2_fused_gauss.err; cat gpurun_out/r2_fused_gauss.json; tail -c 300 gpurun_out/r2_fused_gauss.err
