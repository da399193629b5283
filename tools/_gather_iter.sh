cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests -m gpu -q -x -k "rotate or gather or config3 or chain or fusion or random" 2>&1 | tail -2
timeout -s KILL 150 python tools/bench_configs.py config3 2>&1 | grep "^{"
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gather_f32 -c 6 --csv --log-file gpurun_out/gather_t.csv python tools/bench_configs.py config3 --images 256 --reps 1 > /dev/null 2>&1
grep gather gpurun_out/gather_t.csv | awk -F'","' '{print $5, $NF}' | tail -3
