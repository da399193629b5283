cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="timeout -s KILL"
$T 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2

$T 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print('bench', round(d['value']), d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
$T 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final_launches_bench.json 2>/dev/null
$T 200 ncu --set full --clock-control none --import-source on -k regex:gauss_stream -s 3 -c 1 -o gpurun_out/final_full -f python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu --no-e2e > gpurun_out/final_ncu.log 2>&1
tail -1 gpurun_out/final_ncu.log
