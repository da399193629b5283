cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 120 python -m pytest tests -m gpu -q -x -k "gaussian or config3 or full_size" 2>&1 | tail -2
[ ${PIPESTATUS[0]} -ne 0 ] && exit 1
for ns in 0 100 0 100 300; do
  MILLIPYDE_GAUSS_SLEEP_NS=$ns timeout -s KILL 120 python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e 2>/dev/null | grep '^{' | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sleep_ns=$ns', round(d['value']), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['power_w_median'])"
done
MILLIPYDE_GAUSS_SLEEP_NS=100 timeout -s KILL 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:gauss_stream -s 3 -c 1 --csv --log-file gpurun_out/sleep100.csv python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu --no-e2e > /dev/null 2>&1
grep gauss gpurun_out/sleep100.csv | awk -F'","' '{print $(NF-2), $NF}'
