# one development iteration of the tensor-core Gaussian: parity tests, bench line, ncu capture
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 120 python -m pytest tests -m gpu -q -x -k "gaussian or config3 or chain or full_size or golden" 2>&1 | tail -5
[ ${PIPESTATUS[0]} -ne 0 ] && { echo "tests failed or hung"; exit 1; }
for mode in ${MODES:-mma}; do
  MILLIPYDE_GAUSS_COLUMN=$mode timeout -s KILL 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | grep '^{' > gpurun_out/ab_$mode.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$mode.json"))
print("$mode", round(d["value"]), "img/s frac", round(d["roofline"]["frac"],4), d["roofline"]["kernel"], d["clocks"])
PY
done
[ -n "$NONCU" ] && exit 0
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:gauss_stream -s 3 -c 1 -o gpurun_out/mma_full -f python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu --no-e2e > gpurun_out/mma_ncu.log 2>&1
tail -1 gpurun_out/mma_ncu.log
