# full GPU check: every gpu test, then the bench line(s)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL ${TEST_TIMEOUT:-400} python -m pytest tests -m gpu -q -x 2>&1 | tail -6
[ ${PIPESTATUS[0]} -ne 0 ] && { echo "tests failed or hung"; exit 1; }
for mode in ${MODES:-mma}; do
  MILLIPYDE_TRACE=${TRACE:-0} MILLIPYDE_GAUSS_COLUMN=$mode timeout -s KILL 200 python bench.py --steps ${STEPS:-20} --warmup 3 ${BENCH_ARGS:---no-cpu --no-e2e} > gpurun_out/bench_$mode.log 2>&1
  grep millipyde gpurun_out/bench_$mode.log | tail -4
  grep '^{' gpurun_out/bench_$mode.log > gpurun_out/ab_$mode.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$mode.json"))
print("$mode", round(d["value"]), "img/s frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel"], "launches", d["gpu_launches"], d["clocks"], "e2e", d["e2e"]["value"])
PY
done
