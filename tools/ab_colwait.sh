#!/bin/bash
# same-box A/B of the COLUMN role's hand-off wait (probe + nanosleep vs suspended try_wait), alternating
set -u
mkdir -p gpurun_out
for i in 1 2; do
  for m in 0 1; do
    MILLIPYDE_GAUSS_COLWAIT=$m timeout 300 python bench.py --no-cpu --no-e2e --steps 40 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('colwait=$m', round(d['value']), round(d['roofline']['frac'], 4), d['clocks']['sm_mhz'], d['clocks']['power_w_median'])"
  done
done
timeout 300 python tools/kernel_table.py run > /dev/null 2>&1 && echo table-run-ok
