cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout -s KILL 300 python -m pytest tests/test_multigpu.py tests/test_pipeline_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print('2gpu', round(d['value']), d['roofline']['frac'], d['e2e']['value'], d['clocks']['sm_mhz'])"
