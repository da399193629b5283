cd $GRAFT_REPO_ROOT
for n in 1 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/bench_configs.py config4 2>&1 | grep '^{' 
done
for p in 1 2 4; do
  timeout 120 python tools/bench_configs.py config5 --pairs $p --images 256 2>&1 | grep '^{'
done
