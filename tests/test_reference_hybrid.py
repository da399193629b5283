"""The drop-in boundary proven by construction.

oracle/build_ref.py build_hybrid() compiles the reference's seven CPython type files
(src/device.c, gpuarray.c, gpugenerator.c, gpuimage.c, gpuoperation.c, gpupipeline.c,
millipyde_module.c -- setup.py:48-66 minus the four HIP translation units) UNMODIFIED against the
reference's own headers and links them to libmp_b200.so.  This test runs the golden-vector script
through that module (the reference's host code, this repository's kernels, reference semantics) and
requires what the unmodified reference produced on a B200: tests/golden/reference_outputs.npz."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import build_ref
from tests.golden import make_golden as mg
from tests.test_golden import G, NAMES, compare

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hybrid_links_only_against_the_library_and_libc():
    """CPU half: every symbol the reference's host objects leave undefined is defined by
    libmp_b200.so, by the other host objects, by libc or by the interpreter."""
    if not build_ref.hybrid_available():
        pytest.skip("oracle/_ref/hybrid not built (needs /root/reference)")
    obj_dir = os.path.dirname(build_ref.hybrid_so_path())
    objs = [os.path.join(obj_dir, n + ".o") for n in build_ref.HYBRID_SOURCES]
    undefined = set()
    defined = set()
    for o in objs:
        for line in subprocess.check_output(["nm", o], text=True).splitlines():
            parts = line.split()
            if len(parts) == 2 and parts[0] == "U":
                undefined.add(parts[1])
            elif len(parts) == 3 and parts[1] in "TDBRC":
                defined.add(parts[2])
    lib = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "millipyde_b200", "libmp_b200.so")],
                                  text=True)
    exported = {ln.split()[-1] for ln in lib.splitlines() if ln.strip()}
    gpu_side = {s for s in undefined - defined if s.startswith(("mpimg_", "mpobj_", "mpdev_", "mpwrk_", "mperr_", "random_"))}
    assert len(gpu_side) >= 30, sorted(gpu_side)          # the host code really calls into the boundary
    assert not (gpu_side - exported), sorted(gpu_side - exported)


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_reference_host_code_on_this_library_reproduces_the_reference(tmp_path):
    if not build_ref.hybrid_available():
        pytest.skip("oracle/_ref/hybrid not built (needs /root/reference)")
    out = tmp_path / "hybrid_outputs.npz"
    env = dict(os.environ, MILLIPYDE_SEMANTICS="reference")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden", "make_golden.py"), "--hybrid", str(out)],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=500)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    H = np.load(out)
    assert set(H.files) == set(G.files)
    for name in NAMES:
        assert np.array_equal(H[f"{name}/rgb2grey"], G[f"{name}/rgb2grey"])
        for case in mg.RGBA_CASES:
            k = mg.case_key(f"{name}/rgba", case)
            assert compare(H[k], G[k], case, True) == 0.0, k
        for case in mg.GREY_CASES:
            k = mg.case_key(f"{name}/grey", case)
            assert compare(H[k], G[k], case, False) == 0.0, k
        got, want = H[f"{name}/long_chain"], G[f"{name}/long_chain"]
        assert got.shape == want.shape and np.mean(got != want) < 0.01   # the reference's own Gaussian race, see test_golden
