cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:gather_f32 -s 1 -c 1 -o gpurun_out/gather_full -f python tools/bench_configs.py config3 --images 256 --reps 1 > gpurun_out/gather_ncu.log 2>&1
tail -2 gpurun_out/gather_ncu.log
