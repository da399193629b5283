cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gauss_stream -s 3 -c 1 -o gpurun_out/mma_full -f python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu --no-e2e > gpurun_out/mma_ncu.log 2>&1
ncu -i gpurun_out/mma_full.ncu-rep --page raw --csv > gpurun_out/mma_full_raw.csv 2>/dev/null
tail -3 gpurun_out/mma_ncu.log
