"""CPU checks of the drop-in boundary: libmp_b200.so loads without a GPU and
exports every entry point include/*.h declares; the status enum keeps the
reference's numeric values (src/include/millipyde.h:27-94); messages match
src/millipyde.c:8-138.  No compute calls here."""
import ctypes
import glob
import os
import re

import pytest

from millipyde_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    names = set()
    for path in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        text = re.sub(r"#define MP_STATUS_TABLE.*?\n\n", "\n", text, flags=re.S)
        for m in re.finditer(r"\b((?:mp[a-z]*_|random_)\w+)\s*\(", text):
            names.add(m.group(1))
    return names


def test_library_loads_without_gpu_and_exports_every_declared_symbol():
    lib = capi.lib()
    declared = _declared_functions()
    assert len(declared) >= 60
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    # and the ctypes table covers the headers exactly
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)


def test_status_codes_keep_reference_values():
    lib = capi.lib()
    s = lambda k: lib.mperr_str(k).decode()
    assert s(15) == "GPU runtime failed while querying the device count"       # DEV_ERROR_DEVICE_COUNT
    assert s(29).startswith("Constructing gpuarray requires an ndarray")        # GPUARRAY_..._ARRAY_TYPE
    assert s(35).startswith("Constructing Operation requires a float probability")
    assert s(44) == "Constructing Pipelines can only include one named argument designated 'device'"
    assert s(54) == "Constructing Generator requires a list of Operations"      # last reference code
    assert s(55).startswith("A CUDA runtime call failed")                       # first new code
    assert s(0) == "Unknown failure occurred" and s(-1) == "Unknown failure occurred"
    assert s(10_000) == "Unknown failure occurred"


def test_struct_layouts_match_reference_abi():
    assert ctypes.sizeof(capi.MPObjData) == 56
    assert capi.MPObjData.nbytes.offset == 48 and capi.MPObjData.stream.offset == 32
    assert capi.MPObjData.type.offset == 24 and capi.MPObjData.mem_loc.offset == 28
    assert ctypes.sizeof(capi.MPRunnable) == 32
    assert ctypes.sizeof(capi.ColorizeArgs) == 24 and ctypes.sizeof(capi.GammaArgs) == 16


def test_seeded_random_source_replays():
    lib = capi.lib()
    out = ctypes.c_double()
    lib.mprand_seed(1234)
    a = []
    for _ in range(5):
        assert lib.random_double_in_range(2.0, 5.0, ctypes.byref(out)) == 0
        a.append(out.value)
    lib.mprand_seed(1234)
    b = []
    for _ in range(5):
        lib.random_double_in_range(2.0, 5.0, ctypes.byref(out))
        b.append(out.value)
    assert a == b and all(2.0 <= v <= 5.0 for v in a) and len(set(a)) == 5
    k = ctypes.c_int()
    seen = set()
    for _ in range(200):
        lib.random_int_in_range(3, 6, ctypes.byref(k))
        seen.add(k.value)
    assert seen == {3, 4, 5, 6}
    lib.mprand_seed(0)      # back to getrandom(2)
    lib.random_double_in_range(0.0, 1.0, ctypes.byref(out))
    assert 0.0 <= out.value <= 1.0


def test_effective_gaussian_radius_is_host_only():
    lib = capi.lib()
    full = ctypes.c_int()
    r = lib.mpimg_gaussian_effective_radius(2.0, ctypes.byref(full))
    assert full.value == 16 and r == 10        # dropped tail mass (1.14e-7) < 2^-23, one fp32 ulp of 1.0
    assert lib.mpimg_gaussian_effective_radius(0.0, ctypes.byref(full)) == 0


def test_effective_radius_drops_less_than_one_fp32_ulp():
    """include/mp_image.h: for every sigma the evaluated support is the smallest radius whose dropped
    weight (both sides, scipy's normalisation over the oracle's int(8 sigma + 0.5) taps) is <= 2^-23."""
    import numpy as np
    lib = capi.lib()
    full = ctypes.c_int()
    for sigma in np.concatenate([np.linspace(0.3, 3.2, 59), [4.0, 7.5, 12.0]]):
        r = lib.mpimg_gaussian_effective_radius(float(sigma), ctypes.byref(full))
        nominal = int(8.0 * sigma + 0.5)
        assert full.value == nominal and 0 <= r <= min(nominal, 127)
        d = np.arange(-nominal, nominal + 1, dtype=np.float64)
        w = np.exp(-0.5 * (d / sigma) ** 2)
        w /= w.sum()
        if nominal > 127:
            continue                      # the weight table is capped; the full support runs elsewhere
        dropped = lambda rad: float(w[np.abs(d) > rad].sum())
        assert dropped(r) <= 2.0 ** -23 * (1 + 1e-9), (sigma, r, dropped(r))
        if r > 0:
            assert dropped(r - 1) > 2.0 ** -23 * (1 - 1e-9), (sigma, r, dropped(r - 1))


def test_worker_pool_runs_items_without_gpu():
    lib = capi.lib()
    pool = ctypes.c_void_p()
    assert lib.mpwrk_create_work_pool(ctypes.byref(pool), 3) == 0
    hits = []
    CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p)
    cb = CB(lambda arg: hits.append(arg))
    for i in range(1, 33):
        assert lib.mpwrk_work_queue_push(pool, ctypes.cast(cb, ctypes.c_void_p), ctypes.c_void_p(i)) == 0
    lib.mpwrk_work_wait(pool)
    assert sorted(hits) == list(range(1, 33))
    assert lib.mpwrk_destroy_work_pool(pool) == 0


def test_no_gpu_is_a_loud_error_not_a_fallback():
    import shutil
    if shutil.which("nvidia-smi") and os.path.exists("/dev/nvidia0"):
        pytest.skip("a GPU is present")
    import numpy as np
    with pytest.raises(capi.MillipydeError) as e:
        capi.initialize()
    assert e.value.status == 15
    with pytest.raises(capi.MillipydeError):
        capi.DeviceImage(np.zeros((4, 4), np.float32))
