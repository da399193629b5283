#!/bin/bash
set -u
mkdir -p gpurun_out
for P in 1 2; do
  MILLIPYDE_TRACE=1 timeout 300 python tools/bench_configs.py config5 --pairs $P --reps 1 > gpurun_out/r2_trace_p$P.json 2> gpurun_out/r2_trace_p$P.err
  cat gpurun_out/r2_trace_p$P.json | cut -c1-200
  grep "shard dev" gpurun_out/r2_trace_p$P.err | tail -40 | awk '{print $3, $4, $5, $6, $7, $8, $9}' | sort | uniq -c | sort -rn | head -12
done
timeout 600 python -m pytest tests/test_golden.py tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -3
