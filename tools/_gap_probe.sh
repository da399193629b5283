cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
MILLIPYDE_TRACE=1 timeout -s KILL 120 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/gap_bench.log 2>&1
grep millipyde gpurun_out/gap_bench.log | tail -6
grep '^{' gpurun_out/gap_bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'wall', d['wall_ms_per_step'])"
