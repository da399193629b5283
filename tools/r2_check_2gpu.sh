#!/bin/bash
# Round 2, first GPU visit (2 GPUs visible): does the one-process path survive the headline workload
# with more than one device visible (SCALE_r01 N=1 died with OOM), tests incl. multi-device, bench under torchrun.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.used,memory.total --format=csv > gpurun_out/r2_smi.txt 2>&1
echo "== eager-peer repro (round-1 behaviour), 2 visible" 
MILLIPYDE_EAGER_PEER=1 timeout 300 python bench.py --gpus 1 --steps 6 --no-cpu --no-e2e > gpurun_out/r2_eager_peer.json 2> gpurun_out/r2_eager_peer.err; echo "rc=$?"
tail -c 600 gpurun_out/r2_eager_peer.err
echo "== lazy (new), 2 visible"
timeout 600 python bench.py --gpus 1 --steps 20 > gpurun_out/r2_bench_1of2.json 2> gpurun_out/r2_bench_1of2.err; echo "rc=$?"
tail -c 600 gpurun_out/r2_bench_1of2.err
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== torchrun 2 ranks"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "rc=$?"
tail -c 1500 gpurun_out/r2_bench_2gpu.err
head -c 3000 gpurun_out/r2_bench_2gpu.json
