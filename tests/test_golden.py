"""Golden vectors recorded from the reference's own kernels (built unmodified
through oracle/build_ref.py and run on a B200 by tests/golden/make_golden.py).

CPU half: pins oracle/ref_exact.c to them.  GPU half: the product, called
through the C ABI in MP_SEMANTICS_REFERENCE, reproduces them bit for bit.

One known reference defect is masked: its Gaussian column kernels have no
`col < width` guard (src/millipyde_image.cpp:193-244, :312-381), so for widths
that are not a multiple of 16 the surplus threads of the last block write row
y+1's first 16 - W%16 pixels with the zero-padding clamp evaluated at row y.
The two writers agree except within 8 rows of the top/bottom edge, where the
result is a data race.  Those pixels are excluded; the intended zero-padding
result is what the oracle and the product compute."""
import os

import numpy as np
import pytest

from oracle import ref_exact as rx
from tests.golden import make_golden as mg

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_outputs.npz"))
NAMES = ("charlie", "noise")


def gaussian_defect_mask(h, w):
    """True where the reference's unguarded column threads may have raced."""
    m = np.zeros((h, w), bool)
    surplus = (16 - w % 16) % 16
    if surplus:
        m[:10, :surplus] = True
        m[-10:, :surplus] = True
    return m


def compare(got, want, case, rgba):
    """Returns the fraction of pixels that differ after masking the known defect."""
    bad = np.any(got != want, axis=-1) if rgba else (got != want)
    if case[0] == "gaussian" or case == "long_chain":
        mask = gaussian_defect_mask(*bad.shape)
        if case == "long_chain":      # ... then transposed twice and rotated: compare loosely below
            return float(np.mean(bad))
        bad = bad & ~mask
    return float(np.mean(bad))


# ------------------------------------------------------------------ CPU: the oracle is pinned
@pytest.mark.parametrize("name", NAMES)
def test_oracle_grey_bit_exact(name):
    assert np.array_equal(rx.grey_u8(G[f"{name}/input"]), G[f"{name}/rgb2grey"])


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("case", mg.RGBA_CASES, ids=lambda c: "_".join(map(str, c)))
def test_oracle_rgba_bit_exact(name, case):
    img = G[f"{name}/input"]
    got = rx.apply_chain(img, [case])
    assert compare(got, G[mg.case_key(f"{name}/rgba", case)], case, True) == 0.0


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("case", mg.GREY_CASES, ids=lambda c: "_".join(map(str, c)))
def test_oracle_grey_ops(name, case):
    grey = G[f"{name}/rgb2grey"]
    got = rx.apply_chain(grey, [case])
    want = G[mg.case_key(f"{name}/grey", case)]
    if case[0] == "adjust_gamma":
        # glibc powf vs the device's powf: float-level agreement only
        assert np.abs(got - want).max() < 1e-6
    else:
        assert compare(got, want, case, False) == 0.0


# ------------------------------------------------------------------ GPU: the product reproduces them
@pytest.fixture(scope="module")
def capi():
    from millipyde_b200 import capi as m
    m.initialize()
    m.lib().mpimg_set_semantics(m.SEMANTICS_REFERENCE)
    yield m
    m.lib().mpimg_set_semantics(m.SEMANTICS_ORACLE)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_product_matches_reference_kernels(capi, name):
    img = G[f"{name}/input"]
    grey = capi.DeviceImage(img).apply("rgb2grey").numpy()
    assert np.array_equal(grey, G[f"{name}/rgb2grey"])
    for case in mg.RGBA_CASES:
        got = capi.DeviceImage(img).apply(*case).numpy()
        assert compare(got, G[mg.case_key(f"{name}/rgba", case)], case, True) == 0.0, case
    for case in mg.GREY_CASES:
        got = capi.DeviceImage(grey).apply(*case).numpy()
        assert compare(got, G[mg.case_key(f"{name}/grey", case)], case, False) == 0.0, case
    got = capi.DeviceImage(img).apply_chain(mg.LONG_CHAIN).numpy()
    want = G[f"{name}/long_chain"]
    # the chain starts with the RGBA Gaussian, so the defect pixels propagate (rotated): bounded count
    assert got.shape == want.shape
    assert np.mean(got != want) < 0.01
