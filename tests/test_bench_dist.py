"""bench.py's multi-rank plumbing on CPU: two gloo ranks under torchrun.  The
reference arm is the only leg that runs without a GPU: rank 0 alone measures and
prints one JSON line, the other rank exits 0; the Dist helper's barrier / max /
sum (what the B200 arm uses to combine per-rank device times) is exercised
directly."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, script_args, timeout=600):
    r = None
    for _ in range(3):      # the probed port can be taken between the probe and the rendezvous: retry
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + script_args
        r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
        if r.returncode == 0:
            break
    return r


@pytest.mark.timeout(900)
def test_reference_arm_two_ranks_prints_one_line():
    r = _torchrun(2, ["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                      "--cpu-images", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "images/s"
    assert d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


@pytest.mark.timeout(600)
def test_dist_helper_max_and_sum_over_two_ranks(tmp_path):
    script = tmp_path / "dist_probe.py"
    script.write_text(
        "import sys, json\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "d = bench.Dist()\n"
        "d.barrier()\n"
        "m = d.max(10.0 + d.rank)\n"
        "s = d.sum(256.0)\n"
        "d.barrier()\n"
        "print(json.dumps({'rank': d.rank, 'world': d.world, 'max': m, 'sum': s}))\n"
        "d.close()\n")
    r = _torchrun(2, [str(script)])
    assert r.returncode == 0, r.stderr[-2000:]
    import re
    # two ranks share one stdout: their lines can land back to back without a separator
    rows = [json.loads(m) for m in re.findall(r"\{[^{}]*\}", r.stdout)]
    assert sorted(x["rank"] for x in rows) == [0, 1]
    assert all(x["world"] == 2 and x["max"] == 11.0 and x["sum"] == 512.0 for x in rows)
