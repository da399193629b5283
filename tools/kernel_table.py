#!/usr/bin/env python
"""Drive every kernel once per layout on a batch, for an ncu launch list:

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      --csv --log-file gpurun_out/kernels.csv python tools/kernel_table.py run
  python tools/kernel_table.py parse gpurun_out/kernels.csv > profiles/rNN_kernel_table.md

`run` prints the plan (op, layout, images per launch, algorithmic bytes per image) as JSON lines on
stderr-free stdout so `parse` can join it with the CSV by launch order."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [("f32 4K RGB", (2160, 3840, 3), np.float32, 8), ("f32 1080p RGB", (1080, 1920, 3), np.float32, 32),
         ("f32 1024^2 grey", (1024, 1024), np.float32, 32), ("rgba8 4K", (2160, 3840, 4), np.uint8, 8),
         ("f64 4K grey", (2160, 3840), np.float64, 8)]
OPS = [("rgb2grey",), ("transpose",), ("fliplr",), ("rotate", 30.0), ("brightness", 0.1), ("adjust_gamma", 1.5, 1.0),
       ("colorize", 0.9, 1.1, 1.0), ("gaussian", 2.0)]
CHAIN3 = [("rotate", 30.0), ("fliplr",), ("adjust_gamma", 1.5, 1.0), ("gaussian", 2.0)]


def plan():
    for name, shape, dt, n in CASES:
        item = np.dtype(dt).itemsize
        nbytes = int(np.prod(shape)) * item
        for op in OPS:
            if op[0] == "rgb2grey" and len(shape) == 2:
                continue
            if op[0] == "colorize" and len(shape) == 2:
                continue
            out_b = nbytes
            if op[0] == "rgb2grey":
                out_b = shape[0] * shape[1] * (4 if dt == np.float32 else 8)
            yield name, shape, dt, n, [op], nbytes + out_b
    yield "f32 1080p RGB", (1080, 1920, 3), np.float32, 32, CHAIN3, 2 * 1080 * 1920 * 12


def run():
    from millipyde_b200 import capi, engine
    capi.initialize()
    rng = np.random.default_rng(0)
    for name, shape, dt, n, chain, algo in plan():
        if dt == np.uint8:
            img = rng.integers(0, 256, shape, dtype=np.uint8)
        else:
            img = rng.random(shape).astype(dt)
        seed = capi.DeviceImage(img)
        batch = [seed.clone() for _ in range(n)]
        ch = engine.Chain(chain, device=0)
        before = capi.lib().mpdev_launch_count()
        ch.run(batch)
        launches = capi.lib().mpdev_launch_count() - before
        print(json.dumps({"case": name, "chain": [c[0] for c in chain], "images": n, "launches": int(launches),
                          "algo_bytes_per_image": algo}), flush=True)
        ch.close()
        for d in batch:
            d.close()
        seed.close()


def parse(csv_path, plan_path):
    import csv
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10 and r[0].isdigit()]
    # one record per launch id with its three metrics
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r[0]), {"kernel": r[4]})
        d[r[12]] = float(r[14].replace(",", ""))
        d["unit_" + r[12]] = r[13]
    order = [launches[k] for k in sorted(launches)]
    plans = [json.loads(l) for l in open(plan_path) if l.startswith("{")]
    print("| layout | op / chain | kernel(s) | images per launch | time per image | algorithmic GB/s | of HBM peak (%.0f GB/s) | DRAM bytes / algorithmic |" % peak)
    print("|---|---|---|---|---|---|---|---|")
    i = 0
    for p in plans:
        ls = order[i:i + p["launches"]]
        i += p["launches"]
        t_ns = sum(l["gpu__time_duration.sum"] * (1e3 if l["unit_gpu__time_duration.sum"] == "us" else 1) for l in ls)
        def b(l, k):
            u = l["unit_" + k]
            return l[k] * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        dram = sum(b(l, "dram__bytes_read.sum") + b(l, "dram__bytes_write.sum") for l in ls)
        algo = p["algo_bytes_per_image"] * p["images"]
        gbs = algo / t_ns
        names = sorted({l["kernel"].split("(")[0].replace("void ", "").replace("mpk::", "") for l in ls})
        print(f"| {p['case']} | {'+'.join(p['chain'])} | {', '.join(names)} | {p['images'] // max(1, p['launches']) if p['launches'] <= len(p['chain']) else 1} | "
              f"{t_ns / p['images'] / 1e3:.2f} us | {gbs:.0f} | {100 * gbs / peak:.1f} % | {dram / algo:.2f} |")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run()
    else:
        parse(sys.argv[2], sys.argv[3])
