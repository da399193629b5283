# A/B of the Gaussian column pass: tensor-core (default) vs FMA-pipe (MILLIPYDE_GAUSS_COLUMN=fma)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 120 python -m pytest tests -m gpu -q -x -k "gaussian or config3 or chain or full_size or golden" 2>&1 | tail -5
[ ${PIPESTATUS[0]} -ne 0 ] && { echo "tests failed or hung"; exit 1; }
for mode in ${MODES:-mma fma}; do
  MILLIPYDE_GAUSS_COLUMN=$mode timeout -s KILL 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | grep '^{' > gpurun_out/ab_$mode.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$mode.json"))
print("$mode", round(d["value"]), "img/s frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel"], d["clocks"])
PY
done
