"""Digest of one `ncu --set full` capture: headline metrics from the raw page and the hottest
instructions of the source page.  usage: python tools/ncu_digest.py <rep> [split_marker]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
for h, u, v in zip(hdr, units, vals):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and float(v or 0) > 0.15):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))[2:]
tot = sum(int(r[2]) for r in rows)
print("samples", tot, "instructions", len(rows))
split = next((i for i, r in enumerate(rows) if "USETMAXREG.TRY_ALLOC" in r[1]), None)
if split:
    for name, part in (("ROW", rows[:split]), ("COL", rows[split:])):
        c = collections.Counter()
        for r in part:
            op = [o for o in r[1].split() if not o.startswith("@")][0].split(".")[0]
            c[op] += int(r[2])
        print(name, sum(c.values()), c.most_common(12))
for i, r in sorted(sorted(enumerate(rows), key=lambda x: -int(x[1][2]))[:25]):
    print(i, r[1].strip()[:80], r[2], r[5])
