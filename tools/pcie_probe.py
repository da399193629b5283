"""Ceiling for the e2e leg: pinned host <-> device copy bandwidth of this box, one direction at a time and
both at once -- for ONE rank or for N concurrent ranks (one GPU each, under torchrun), which is what bounds
bench.py's `e2e` at --gpus N: every rank's H2D + D2H stream crosses the same host memory system.

    python tools/pcie_probe.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/pcie_probe.py

torch is used only as a CUDA-runtime wrapper here (and gloo for the barrier); not part of the product.
Prints one JSON line (rank 0): per-rank mean and the AGGREGATE GB/s over all ranks, plus what the same
bytes mean in 4K RGB fp32 images/s (99.5 MB in + 99.5 MB out per image)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import Dist  # noqa: E402

dist = Dist()
if dist.world > 1:
    outer = [x for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip() != ""]
    os.environ["CUDA_VISIBLE_DEVICES"] = outer[dist.local] if len(outer) > dist.local else str(dist.local)

import torch  # noqa: E402

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
h_out.fill_(2)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=6):
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = dist.max(time.perf_counter() - t0)      # the slowest rank closes the step
    return reps * n / dt / 1e9                    # GB/s per rank per direction


run(True, True, 2)
h2d, d2h, both = run(True, False), run(False, True), run(True, True)
if dist.rank == 0:
    img = 3840 * 2160 * 3 * 4 / 1e9
    print(json.dumps({"ranks": dist.world, "bytes_per_copy": n,
                      "h2d_only_GBs_per_rank": round(h2d, 2), "d2h_only_GBs_per_rank": round(d2h, 2),
                      "both_GBs_each_way_per_rank": round(both, 2),
                      "aggregate_both_GBs_each_way": round(both * dist.world, 2),
                      "e2e_ceiling_images_per_s_4k_rgb_f32": round(both * dist.world / img, 1),
                      "cpus": len(os.sched_getaffinity(0))}), flush=True)
dist.close()
