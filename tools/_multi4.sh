cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
: > gpurun_out/multi4.jsonl
for n in 1 2 4; do
  timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/bench_configs.py config4 2>&1 | grep '^{' | tee -a gpurun_out/multi4.jsonl
done
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29536 tools/bench_configs.py config4 --to-host --images 16384 2>&1 | grep '^{' | tee -a gpurun_out/multi4.jsonl
for p in 1 2; do
  timeout -s KILL 120 python tools/bench_configs.py config5 --pairs $p --images 256 2>&1 | grep '^{' | tee -a gpurun_out/multi4.jsonl
done
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu 2>/dev/null | grep '^{' > gpurun_out/bench_4gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_4gpu.json')); print('4gpu', round(d['value']), d['roofline']['frac'], d['e2e']['value'], d['clocks']['sm_mhz'])"
