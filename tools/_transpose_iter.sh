cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests -m gpu -q -x -k "index_ops or transpose or config1 or rgba8 or golden or config5 or np_function or device_operators or f64" 2>&1 | tail -2
timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:transpose --csv --log-file gpurun_out/transpose_t.csv python tools/transpose_probe.py 2>&1 | tail -1
grep transpose gpurun_out/transpose_t.csv | awk -F'","' '{print $5, $9, $NF}'
