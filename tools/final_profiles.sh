# Round-end evidence on one B200 (under gpurun): every gpu test, the bench line (with CPU baseline and
# e2e), the reference arm, the ncu launch list of the bench command and one full ncu capture of the
# Gaussian.  Everything lands in gpurun_out/; the summaries under profiles/ are written from there.
#   KERNEL_TABLE=1  also drive every op x layout under ncu (tools/kernel_table.py)
#   FMA_ARM=1       also the bench line with MILLIPYDE_GAUSS_COLUMN=fma (same box A/B)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="timeout -s KILL"
$T 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
$T 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print('bench', round(d['value']), d['roofline']['frac'], d['e2e']['value'], d['clocks'], d['parity_check'])"
[ -n "$FMA_ARM" ] && MILLIPYDE_GAUSS_COLUMN=fma $T 200 python bench.py --no-cpu --no-e2e > gpurun_out/final_bench_fma.json 2>/dev/null
$T 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference_arm.json 2>/dev/null
$T 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final_launches_bench.json 2>/dev/null
$T 200 ncu --set full --clock-control none --import-source on -k regex:gauss_stream -s 3 -c 1 -o gpurun_out/final_full -f python bench.py --steps 1 --warmup 3 --batch 64 --no-cpu --no-e2e > gpurun_out/final_ncu.log 2>&1
tail -1 gpurun_out/final_ncu.log
$T 200 python tools/fused_gauss_probe.py 128 > gpurun_out/final_fused_gauss.json 2>/dev/null
$T 200 python tools/bench_configs.py config3 > gpurun_out/final_config3.json 2>/dev/null
[ -n "$KERNEL_TABLE" ] && $T 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/kernels.csv python tools/kernel_table.py run > gpurun_out/kernels_plan.jsonl 2>/dev/null
true
