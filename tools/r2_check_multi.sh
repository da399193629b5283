#!/bin/bash
# Round 2 multi-GPU visit (N = 2, 4 or 8 visible GPUs of one box):  bash tools/r2_check_multi.sh N
#  * topology + the box's aggregate pinned-copy ceiling at 1..N concurrent ranks (bounds bench.py's e2e)
#  * bench.py --gpus 1 with all N devices visible (SCALE_r01's N=1 died here with OOM)
#  * bench.py under torchrun at N ranks: ranks number + the one-process N-device leg + multi_device_check
#  * tests/test_multigpu.py, config 4 (one Generator spreading over N devices) and config 5 (N/2 pairs)
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
T=r2_${N}gpu
{ nvidia-smi topo -m; echo; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)"; echo; free -g | head -2; } > $O/${T}_topology.txt 2>&1
if [ -z "${ONLY_CONFIGS:-}" ]; then
echo "== pcie ceiling, 1..$N ranks"
: > $O/${T}_pcie_ranks.jsonl
for R in 1 2 4 8; do
  [ $R -le $N ] || continue
  if [ $R -eq 1 ]; then timeout 200 python tools/pcie_probe.py >> $O/${T}_pcie_ranks.jsonl 2>> $O/${T}_pcie.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $R --master-addr 127.0.0.1 --master-port $((29600+R)) tools/pcie_probe.py >> $O/${T}_pcie_ranks.jsonl 2>> $O/${T}_pcie.err; fi
done
cat $O/${T}_pcie_ranks.jsonl
echo "== bench --gpus 1 with $N devices visible"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err; echo "rc=$?"; tail -c 400 $O/${T}_bench_n1.err
echo "== bench --gpus $N under torchrun"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "rc=$?"; tail -c 600 $O/${T}_bench.err
python - <<PY
import json
for f in ("$O/${T}_bench_n1.json", "$O/${T}_bench.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"]), "frac", round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"], 1), d["e2e"].get("step_s"),
              "clk", d["clocks"]["sm_mhz"], "parity", d["parity_check"]["max_abs_err"] if d.get("parity_check") else None)
        if "in_process" in d:
            print("   in_process", {k: v for k, v in d["in_process"].items() if k != "multi_device_check"}, d.get("multi_device_check"), d["in_process"].get("multi_device_check"))
    except Exception as e:
        print(f, "unreadable", e)
PY
echo "== pytest multigpu"
timeout 600 python -m pytest tests/test_multigpu.py tests/test_reference_hybrid.py -m gpu -x -q 2>&1 | tail -8
fi
echo "== config 4: one Generator spreading 65536 outputs over $N devices"
: > $O/${T}_configs.jsonl
timeout 600 python tools/bench_configs.py config4 --spread --ref-params --prefetch 2048 >> $O/${T}_configs.jsonl 2>> $O/${T}_configs.err
timeout 600 python tools/bench_configs.py config4 --spread --ref-params --prefetch 8192 >> $O/${T}_configs.jsonl 2>> $O/${T}_configs.err
timeout 600 python tools/bench_configs.py config4 --spread --ref-params --prefetch 512 --images 8192 --to-host >> $O/${T}_configs.jsonl 2>> $O/${T}_configs.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/bench_configs.py config4 --ref-params --prefetch 256 >> $O/${T}_configs.jsonl 2>> $O/${T}_configs.err
echo "== config 5: $((N/2)) pair(s)"
timeout 600 python tools/bench_configs.py config5 --pairs $((N/2)) --threads >> $O/${T}_configs.jsonl 2>> $O/${T}_configs.err
timeout 600 python tools/bench_configs.py config5 --pairs $((N/2)) >> $O/${T}_configs.jsonl 2>> $O/${T}_configs.err
timeout 600 python tools/bench_configs.py config5 --pairs 1 >> $O/${T}_configs.jsonl 2>> $O/${T}_configs.err
cat $O/${T}_configs.jsonl; tail -c 600 $O/${T}_configs.err
# NVLink bytes of the hand-off as the hardware counts them (nvidia-smi nvlink counters read N/A in this VM):
# the sender's last kernel (the transpose) writes its outputs into the receiver's memory
timeout 300 ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum --clock-control none -k regex:transpose -c 4 --csv \
  --log-file $O/${T}_nvlink_ncu.csv python tools/bench_configs.py config5 --pairs 1 --images 16 --reps 0 > /dev/null 2>> $O/${T}_configs.err
tail -6 $O/${T}_nvlink_ncu.csv
