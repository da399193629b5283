#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json `metric`).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic images: the
workload is BASELINE.json configs[1] -- Gaussian blur sigma=2 on a batch of
3840x2160 RGB fp32 images (256 per GPU; fewer only if device memory is short,
and then `config` says so).  `value` = images/s with the batch resident in HBM
when the timed region starts, timed with CUDA events on the launching stream,
max over ranks.  `e2e` = the same metric through the public call with HOST
buffers: pinned host memory -> device -> blur -> pinned host memory, every copy
inside the timed region.

Multi-GPU: images are independent, so the batch is sharded over ranks with no
data-path collective ("scaling": "weak", 256 images per GPU).  Under torchrun
each rank owns GPU LOCAL_RANK; torch.distributed is used only for the barrier
and the max-over-ranks of the device times.

--impl reference times the reference's CPU path for the same workload: the
scikit-image calls of its test suite, restated on scipy (oracle/, kind "port")
on all host cores, on a bounded sample of the batch.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, C = 2160, 3840, 3
SIGMA = 2.0
BATCH_PER_GPU = 256
ALGO_BYTES_PER_IMAGE = 2 * H * W * C * 4          # read once + written once
WORKLOAD = "gaussian sigma=2, 3840x2160 RGB fp32, batch 256/GPU (BASELINE configs[1])"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="images per GPU per step")
    ap.add_argument("--e2e-images", type=int, default=32, help="images per e2e step")
    ap.add_argument("--cpu-images", type=int, default=0, help="images in the CPU sample (0 = one per core)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e leg (profiling runs only)")
    return ap.parse_args()


# --------------------------------------------------------------------------- CPU leg
def _cpu_one(seed):
    import numpy as np
    from oracle import skimage_oracle as so
    rng = np.random.default_rng(seed)
    img = rng.random((H, W, C), dtype=np.float32)      # input generation is not part of the timed path
    t0 = time.perf_counter()
    out = so.gaussian(img, SIGMA)
    dt = time.perf_counter() - t0
    return os.getpid(), dt, float(out[H // 2, W // 2, 0])


def cpu_baseline(n_images: int):
    """The oracle's Gaussian (scipy.ndimage.gaussian_filter -- the call
    skimage.filters.gaussian makes) on `n_images` images, one process per core, all
    cores busy at once.  Throughput = images / (busiest worker's total filter time):
    process start-up and synthetic-input generation are excluded, as on the GPU arm."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0))
    n = n_images or cores
    procs = min(cores, n)
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        rows = pool.map(_cpu_one, [2000 + k for k in range(n)], chunksize=1)
    per_worker = {}
    for pid, dt, _ in rows:
        per_worker[pid] = per_worker.get(pid, 0.0) + dt
    busy = max(per_worker.values())
    return {"value": n / busy, "unit": "images/s", "cores": procs, "kind": "port",
            "sample": f"{n} of the step's images, scipy.ndimage.gaussian_filter(sigma=2, truncate=8, "
                      f"mode=constant) per image, {procs} processes in parallel, busiest worker {busy:.2f}s"}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock, power and throttle reasons of one GPU sampled DURING the timed region: NVML in a
    thread every 5 ms when pynvml is importable (a 250 ms timed region gets ~50 samples), else
    `nvidia-smi -lms 100`."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []          # (arrival time, sm MHz, max MHz, watts, set of reasons)
        self.window = None      # (t0, t1) of the timed region
        self.proc = None
        self.gpu = gpu_index
        self.thread = None
        self.stop_flag = False
        self.how = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.how = "NVML every 5 ms"
            self.thread = threading.Thread(target=self._pump_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.how = "nvidia-smi -lms 100"
        self.thread = threading.Thread(target=self._pump_smi, daemon=True)
        self.thread.start()

    def _pump_nvml(self):
        nv = self.nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                watts = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.time(), sm, self.max_sm, watts, {k for k, b in bits.items() if mask & b}))
            except Exception:
                pass
            time.sleep(0.005)

    def _pump_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) < 9:
                continue
            try:
                names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
                self.rows.append((time.time(), float(r[1]), float(r[2]), float(r[3]),
                                  {n for n, v in zip(names, r[5:9]) if v.lower().startswith("active")}))
            except ValueError:
                continue

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread and self.nv:
            self.thread.join(timeout=1)
        if not self.how:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML, no nvidia-smi"]}
        rows = self.rows
        scope = "warm-up + timed region"
        if self.window:
            inside = [x for x in rows if self.window[0] <= x[0] <= self.window[1]]
            # a timed region shorter than the sampling period: fall back to every sample taken
            # since the warm-up started (the GPU was under the same load throughout)
            if len(inside) >= 3:
                rows, scope = inside, "timed region only"
        sm = sorted(r[1] for r in rows)
        reasons = set().union(*[r[4] for r in rows]) if rows else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max((r[2] for r in rows), default=None),
                "power_w_max": max((r[3] for r in rows), default=0.0),
                "power_w_median": sorted(r[3] for r in rows)[len(rows) // 2] if rows else None,
                "samples": len(sm), "reasons": sorted(reasons),
                "note": f"{self.how}, {scope}; sm_mhz = median"}


# --------------------------------------------------------------------------- dist
class Dist:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.pg = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group(backend="gloo", rank=self.rank, world_size=self.world)
            self.pg = dist

    def barrier(self):
        if self.pg:
            self.pg.barrier()

    def max(self, x: float) -> float:
        if not self.pg:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.pg.all_reduce(t, op=self.pg.ReduceOp.MAX)
        return float(t[0])

    def sum(self, x: float) -> float:
        if not self.pg:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.pg.all_reduce(t, op=self.pg.ReduceOp.SUM)
        return float(t[0])

    def close(self):
        if self.pg:
            self.pg.destroy_process_group()


# --------------------------------------------------------------------------- reference arm
def run_reference(args, dist):
    if dist.rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    n = args.cpu_images or cores
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_baseline(min(n, cores))
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_baseline(n))
    wall = time.perf_counter() - t_all0
    v = sum(x["value"] for x in vals) / len(vals)
    base = vals[-1]
    base["value"] = v
    line = {
        "impl": "reference", "metric": "augmented images/sec (4K RGB fp32 pipeline)", "value": v,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_images_per_step": n,
                   "note": "reference CPU path = the scikit-image calls of its test suite restated on scipy "
                           "(scikit-image is not installable here); runs on host cores only"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- B200 arm
def run_b200(args, dist):
    import numpy as np
    physical_gpu = 0
    outer = [x for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip() != ""]
    if dist.world > 1:
        # one rank per GPU: this process only ever sees its own device
        mine = outer[dist.local] if len(outer) > dist.local else str(dist.local)
        os.environ["CUDA_VISIBLE_DEVICES"] = mine
        physical_gpu = int(mine) if mine.isdigit() else dist.local
    elif outer and outer[0].isdigit():
        physical_gpu = int(outer[0])
    from millipyde_b200 import capi
    from millipyde_b200 import engine

    ndev = capi.initialize()
    L = capi.lib()
    in_process_gpus = args.gpus if dist.world == 1 else 1
    if in_process_gpus > ndev:
        raise SystemExit(f"--gpus {args.gpus} but only {ndev} device(s) visible")
    devices = list(range(in_process_gpus))

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        hbm_peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"

    # ---- batch: a few distinct host images, replicated on the device -------------
    free_b, total_b = ctypes.c_size_t(), ctypes.c_size_t()
    L.mpdev_mem_info(0, ctypes.byref(free_b), ctypes.byref(total_b))
    img_bytes = H * W * C * 4
    batch = args.batch
    # inputs + outputs of one step live together (the op is out-of-place into pool memory)
    # Two resident batches per GPU, processed alternately (step k works on batch k % 2): the host
    # prepares step k + 1 (pool allocations, pointer tables, launch) while the device runs step k.
    n_sets = 2
    while batch > 8 and 2.2 * n_sets * batch * img_bytes > free_b.value:
        batch //= 2
    rng = np.random.default_rng(2000)
    seeds = [rng.random((H, W, C), dtype=np.float32) for _ in range(2)]
    shard_sets = [[] for _ in range(n_sets)]
    for d in devices:
        L.mpdev_set_target_device(d)
        base = [capi.DeviceImage(s) for s in seeds]
        for shards in shard_sets:
            shards.append([base[k % len(base)].clone(device=d) for k in range(batch)])
        for b in base:
            b.close()
    L.mpdev_set_target_device(capi.DEVICE_LOC_NO_AFFINITY)
    L.mpdev_synchronize_all()

    # one Pipeline per (batch, device): Pipeline.run() = submit + wait, here issued one step ahead
    pipes = [[engine.Chain([("gaussian", SIGMA)], device=d) for d in devices] for _ in range(n_sets)]
    in_flight = [False] * n_sets

    def drain():
        for k in range(n_sets):
            if in_flight[k]:
                for ch in pipes[k]:
                    ch.wait()
                in_flight[k] = False

    def run_steps(n):
        for k in range(n):
            s = k % n_sets
            if in_flight[s]:      # this batch's previous pass must be complete before the next one
                for ch in pipes[s]:
                    ch.wait()
            for ch, imgs in zip(pipes[s], shard_sets[s]):
                ch.submit(imgs)
            in_flight[s] = True
        drain()

    sampler = ClockSampler(physical_gpu)
    sampler.start()
    t_warm = time.time()
    while True:     # >= W warm-up steps and long enough for the clock sampler to come up
        run_steps(max(args.warmup, 3))
        L.mpdev_synchronize_all()
        if time.time() - t_warm > 1.0:
            break
    dist.barrier()
    L.mpdev_synchronize_all()
    t_region0 = time.time()
    launches0 = L.mpdev_launch_count()
    evs = [(L.mpdev_event_create(d), L.mpdev_event_create(d)) for d in devices]
    t0 = time.perf_counter()
    for (e0, _), d in zip(evs, devices):
        L.mpdev_event_record(e0, engine.timing_stream(d))
    run_steps(args.steps)       # returns when the last step's work is complete
    for (_, e1), d in zip(evs, devices):
        L.mpdev_event_record(e1, engine.timing_stream(d))
    dev_ms = max(L.mpdev_event_elapsed_ms(e0, e1) for e0, e1 in evs)
    L.mpdev_synchronize_all()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    launches = L.mpdev_launch_count() - launches0
    dist.barrier()
    sampler.window = (t_region0, time.time())
    clocks = sampler.stop()
    for e0, e1 in evs:
        L.mpdev_event_destroy(e0)
        L.mpdev_event_destroy(e1)

    dev_ms = dist.max(dev_ms)
    images_per_step = dist.sum(float(batch * len(devices)))
    ms_per_step = dev_ms / args.steps
    value = images_per_step / (ms_per_step / 1000.0)

    # roofline of the dominant kernel (the Gaussian): launches run back to back on the timing
    # stream, so average launch duration = device time / launches on this rank
    kern_launches = max(1, launches // max(1, len(devices)))
    avg_launch_ms = (dev_ms if dist.world == 1 else dev_ms) / kern_launches
    images_per_launch = batch * args.steps / kern_launches
    achieved = images_per_launch * ALGO_BYTES_PER_IMAGE / (avg_launch_ms / 1000.0) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": _ncu_traffic(images_per_launch),
                "kernel": ("gauss_stream_ws_kernel<3,11>" if os.environ.get("MILLIPYDE_GAUSS_COLUMN") == "fma"
                           else "gauss_stream_mma_kernel<3,11>"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": images_per_launch * ALGO_BYTES_PER_IMAGE,
                "avg_launch_ms": avg_launch_ms}

    # ---- e2e: host buffers in, host buffers out ----------------------------------
    if args.no_e2e:
        e2e = {"images_per_s": 0.0, "h2d_bytes": 0, "d2h_bytes": 0, "images": 0}
    else:
        # page-locked staging is 2 x 99.5 MB per image per rank: keep the whole job's pinned set bounded
        n_e2e = args.e2e_images if dist.world <= 2 else max(8, args.e2e_images // 2)
        e2e = None
        while e2e is None:
            try:
                e2e = engine.e2e_gaussian(devices, n_e2e, (H, W, C), SIGMA, steps=max(2, min(args.steps, 3)))
            except MemoryError:
                if n_e2e <= 2:
                    raise
                n_e2e //= 2
    e2e_value = dist.sum(e2e["images_per_s"])

    line = {
        "metric": "augmented images/sec (4K RGB fp32 pipeline)", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_gpu_per_step": batch, "image_shape": [H, W, C],
                   "sigma": SIGMA, "effective_radius": 11, "l2": "inputs (25.5 GB/GPU per step) far exceed the 126 MB L2",
                   "batches": "2 resident batches per GPU, alternating; the host enqueues step k+1 while step k runs",
                   "parallelism": f"{args.gpus} GPU(s), images sharded, no collective",
                   "launcher": "torchrun ranks" if dist.world > 1 else "one process"},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(dist.sum(float(e2e["h2d_bytes"]))),
                "d2h_bytes_per_step": int(dist.sum(float(e2e["d2h_bytes"]))),
                "images_per_step": int(dist.sum(float(e2e["images"]))),
                "note": "pinned host -> device -> blur -> pinned host, copies in the timed region"},
        "gpu_launches": int(launches), "clocks": clocks, "wall_ms_per_step": wall_ms / args.steps,
    }
    if dist.rank == 0 and not args.no_cpu and args.gpus == 1:
        line["cpu_baseline"] = cpu_baseline(args.cpu_images)
    elif dist.rank == 0:
        line["cpu_baseline"] = None
    if dist.rank == 0:
        print(json.dumps(line), flush=True)
    for shards in shard_sets:
        for imgs in shards:
            for i in imgs:
                i.close()


def _ncu_traffic(images_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed
    `ncu --set full` capture of this kernel (profiles/gauss_stream_dram.json holds
    the per-image figure and names the capture), scaled to this run's images per
    launch; None if no capture has been committed."""
    p = os.path.join(ROOT, "profiles", "gauss_stream_dram.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d["dram_bytes_per_image"] * images_per_launch


def main():
    args = parse()
    dist = Dist()
    try:
        if args.impl == "reference":
            run_reference(args, dist)
        else:
            run_b200(args, dist)
    finally:
        dist.close()


if __name__ == "__main__":
    main()
