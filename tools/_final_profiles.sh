# Round-end evidence: bench lines (both arms, both column passes), launch list, full ncu capture,
# per-kernel table.  Everything lands in gpurun_out/ and is summarised under profiles/ afterwards.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="timeout -s KILL"
$T 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -c 600 gpurun_out/final_bench.json
MILLIPYDE_GAUSS_COLUMN=fma $T 200 python bench.py --no-cpu --no-e2e > gpurun_out/final_bench_fma.json 2>/dev/null
$T 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference_arm.json 2>/dev/null
$T 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final_launches_bench.json 2>/dev/null
$T 200 ncu --set full --clock-control none --import-source on -k regex:gauss_stream -s 3 -c 1 -o gpurun_out/final_full -f python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu --no-e2e > gpurun_out/final_ncu.log 2>&1
$T 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/kernels.csv python tools/kernel_table.py run > gpurun_out/kernels_plan.jsonl 2>/dev/null
wc -l gpurun_out/kernels.csv gpurun_out/final_launches.csv
