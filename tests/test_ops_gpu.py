"""Parity of every operator, called through the C ABI (libmp_b200.so), against
the CPU oracle on the same seeded inputs.

Bars: bit-exact for index ops, transfers, clone and every uint8 path;
max-abs <= 1e-5 on [0,1] fp32 for float ops (BASELINE.json north star);
<= 1e-12 for the fp64 layouts.  Shapes include BASELINE config 1 (512 x 512
RGBA8, grey + transpose), ragged widths and the reference's own PNG fixture."""
import numpy as np
import pytest

from oracle import ref_exact as rx
from oracle import skimage_oracle as so
from tests import synth

pytestmark = pytest.mark.gpu

TOL32 = 1e-5
TOL64 = 1e-12


@pytest.fixture(scope="module")
def capi():
    from millipyde_b200 import capi as m
    m.initialize()
    m.lib().mpimg_set_semantics(m.SEMANTICS_ORACLE)
    return m


def dev(capi, a):
    return capi.DeviceImage(a)


SHAPES = [(64, 48), (97, 131), (256, 512)]
CHANNELS = [1, 3, 4]


# ------------------------------------------------------------------ transfers
@pytest.mark.parametrize("dtype", [np.uint8, np.float32, np.float64, np.int64, np.int32])
def test_round_trip_bit_exact(capi, dtype):
    rng = np.random.default_rng(7)
    a = (rng.random((33, 21, 3)) * 200).astype(dtype)
    assert np.array_equal(dev(capi, a).numpy(), a)


def test_clone_is_deep(capi):
    a = synth.rgba8(64, 80, 1000)
    d = dev(capi, a)
    c = d.clone()
    c.apply("rgb2grey")
    assert np.array_equal(d.numpy(), a)
    assert np.abs(c.numpy() - so.rgb2grey(a)).max() < TOL64


# ------------------------------------------------------------------ fp32 path
@pytest.mark.parametrize("c", CHANNELS)
@pytest.mark.parametrize("shape", SHAPES)
def test_f32_index_ops_bit_exact(capi, shape, c):
    a = synth.noise_f32(*shape, c, 2000 + c)
    assert np.array_equal(dev(capi, a).apply("transpose").numpy(), so.transpose(a))
    assert np.array_equal(dev(capi, a).apply("fliplr").numpy(), so.fliplr(a))
    assert np.array_equal(dev(capi, a).apply("transpose").apply("transpose").numpy(), a)


@pytest.mark.parametrize("c", CHANNELS)
@pytest.mark.parametrize("shape", SHAPES)
def test_f32_pointwise(capi, shape, c):
    a = synth.noise_f32(*shape, c, 3000 + c)
    got = dev(capi, a).apply("brightness", 0.25).numpy()
    want = so.brightness(a, 0.25)
    if c == 4:
        want[..., 3] = a[..., 3]
    assert np.abs(got - want).max() <= TOL32
    got = dev(capi, a).apply("brightness", -0.4).numpy()
    want = so.brightness(a, -0.4)
    if c == 4:
        want[..., 3] = a[..., 3]
    assert np.abs(got - want).max() <= TOL32
    for gamma, gain in [(2.0, 1.0), (1.5, 1.0), (0.5, 0.8)]:
        got = dev(capi, a).apply("adjust_gamma", gamma, gain).numpy()
        want = np.clip(so.adjust_gamma(a, gamma, gain), 0, 1)
        if c == 4:
            want[..., 3] = a[..., 3]
        assert np.abs(got - want).max() <= TOL32, (gamma, gain)
    got = dev(capi, a).apply("colorize", 0.5, 1.5, 1.1).numpy()
    assert np.abs(got - so.colorize(a, 0.5, 1.5, 1.1)).max() <= TOL32


@pytest.mark.parametrize("c", [3, 4])
@pytest.mark.parametrize("shape", SHAPES)
def test_f32_grey(capi, shape, c):
    a = synth.noise_f32(*shape, c, 4000 + c)
    d = dev(capi, a).apply("rgb2grey")
    assert d.shape == shape and d.dtype == np.float32
    want = a[..., :3].astype(np.float64) @ np.array(so.LUMA)
    assert np.abs(d.numpy() - np.minimum(want, 1.0)).max() <= TOL32


@pytest.mark.parametrize("c", CHANNELS)
@pytest.mark.parametrize("sigma", [2.0, 0.7, 3.3])
def test_f32_gaussian(capi, c, sigma):
    for shape, img in [((97, 131), None), ((256, 512), None), ((128, 96), "smooth")]:
        a = synth.smooth_f32(*shape, c) if img else synth.noise_f32(*shape, c, 5000 + c)
        got = dev(capi, a).apply("gaussian", sigma).numpy()
        want = so.gaussian(a, sigma)
        assert got.dtype == np.float32 and got.shape == a.shape
        assert np.abs(got - want).max() <= TOL32, (shape, sigma)


@pytest.mark.parametrize("dtype", ["f32", "f64", "rgba8"])
def test_transpose_skewed_tile_kernel_shapes(capi, dtype):
    """One- and two-word pixels take the 64-row skewed-layout TMA transpose: full tiles, partial
    tiles on either edge, images smaller than a tile, several tiles per axis -- bit-exact."""
    rng = np.random.default_rng(8100)
    for h, w in [(64, 64), (4, 8), (68, 132), (200, 36), (130, 258), (256, 320), (540, 96)]:
        if dtype == "rgba8":
            a = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
            want = np.transpose(a, (1, 0, 2))
        else:
            a = rng.random((h, w)).astype(np.float32 if dtype == "f32" else np.float64)
            want = a.T
        got = dev(capi, a).apply("transpose").numpy()
        assert got.shape == want.shape and np.array_equal(got, want), (dtype, h, w)


# The streaming kernels (W*C % 4 == 0, W*C >= 64) in both column-pass forms -- tensor-core (MMA,
# the default) and FMA-pipe -- on shapes that exercise every boundary of their tiling: heights
# around the 8-row chunk / 12-row group / 48-row ring sizes, strips that end inside an 80-column
# slice or a 16-column tile, more than one strip, every radius bucket (sigma 0.4 .. 2.0; 2.3 has
# effective radius 13 and always takes the FMA-pipe kernel).
STREAM_SHAPES = [(1, 64), (7, 16), (8, 84), (9, 161), (23, 200), (48, 160), (49, 321), (100, 644), (131, 1284)]


@pytest.mark.parametrize("column", ["mma", "fma"])
@pytest.mark.parametrize("c", CHANNELS)
def test_f32_streaming_gaussian_shapes(capi, c, column):
    L = capi.lib()
    L.mpimg_set_gauss_column(1 if column == "fma" else 0)
    try:
        for k, (h, w) in enumerate(STREAM_SHAPES):
            if (w * c) % 4 or w * c < 64:
                w = (w + 3) // 4 * 4
            a = synth.noise_f32(h, w, c, 7000 + 10 * k + c)
            for sigma in (2.0, 0.4, 1.1):
                got = dev(capi, a).apply("gaussian", sigma).numpy()
                assert np.abs(got - so.gaussian(a, sigma)).max() <= TOL32, (h, w, c, sigma, column)
        a = synth.smooth_f32(300, 700, c)
        for sigma in (0.7, 1.5, 2.0, 2.3):
            got = dev(capi, a).apply("gaussian", sigma).numpy()
            assert np.abs(got - so.gaussian(a, sigma)).max() <= TOL32, (c, sigma, column)
    finally:
        L.mpimg_set_gauss_column(0)


@pytest.mark.parametrize("column", ["mma", "fma"])
def test_f32_streaming_gaussian_any_width(capi, column):
    """Rows that are not whole 16-byte vectors (W * C % 4 = 1, 2, 3 -- e.g. every 1125-wide RGB fixture of
    the reference as fp32): the streaming kernels land the enclosing aligned span of each row and shift
    it into place, outputs leave as scalar stores.  Single images (row chunks), batches, several strips,
    a last strip that ends mid-vector, fused pointwise ops."""
    from millipyde_b200 import engine
    L = capi.lib()
    L.mpimg_set_gauss_column(1 if column == "fma" else 0)
    try:
        for c, widths in ((1, (65, 250, 251, 253, 1279)), (3, (22, 85, 250, 427)), (4, ())):
            for w in widths:
                assert (w * c) % 4 != 0
                for h in (1, 9, 50, 131):
                    a = synth.noise_f32(h, w, c, 7400 + w + h)
                    for sigma in (2.0, 0.7):
                        got = dev(capi, a).apply("gaussian", sigma).numpy()
                        assert np.abs(got - so.gaussian(a, sigma)).max() <= TOL32, (h, w, c, sigma, column)
        a = synth.noise_f32(700, 1125, 3, 7450)                        # the reference's fixture width; 3 row chunks
        got = dev(capi, a).apply("gaussian", 2.0).numpy()
        assert np.abs(got - so.gaussian(a, 2.0)).max() <= TOL32
        imgs = [synth.noise_f32(90, 1125, 3, 7460 + k) for k in range(5)]
        for chain in ([("gaussian", 2.0)], [("adjust_gamma", 1.5, 1.0), ("gaussian", 1.3), ("brightness", 0.1)],
                      [("random_gaussian", 1.0, 2.0)]):
            L.mprand_seed(3)
            devs = [dev(capi, x) for x in imgs]
            ch = engine.Chain(chain, device=0)
            ch.run(devs)
            key = L.mppipe_last_run_key(ch.ptr)
            L.mprand_seed(0)
            for i, (x, d) in enumerate(zip(imgs, devs)):
                ops = [("gaussian", L.mprand_keyed_double(key, i, 0, 0, 1.0, 2.0))] if chain[0][0].startswith("random") else chain
                assert np.abs(d.numpy() - so.apply_chain(x, ops)).max() <= TOL32, chain
    finally:
        L.mpimg_set_gauss_column(0)


def test_f32_streaming_gaussian_column_forms_agree(capi):
    """Tensor-core and FMA-pipe column passes on the same input: the split-precision products cost
    well under 1e-6, and out-of-range rows/columns are exact zeros' worth in both."""
    L = capi.lib()
    a = synth.noise_f32(211, 800, 3, 7100)
    outs = {}
    for mode in (0, 1):
        L.mpimg_set_gauss_column(mode)
        outs[mode] = dev(capi, a).apply("gaussian", 2.0).numpy()
    L.mpimg_set_gauss_column(0)
    assert L.mpimg_get_gauss_column() == 0
    assert np.abs(outs[0] - outs[1]).max() <= 1e-6
    assert np.abs(outs[0].astype(np.float64) - so.gaussian(a, 2.0)).max() <= 2e-6


def test_f32_gaussian_declared_value_range(capi):
    """fp32 images are defined on [0, 1] (include/mp_image.h): under MP_RANGE_UNIT the tensor-core
    column pass converts operands to fp16 and overflows for |sample| >= 65504.  A caller with other
    float data declares MP_RANGE_ANY and gets the FMA-pipe kernel: uint16-range and 1e6-range data
    then blur to the oracle within the contract relative to the data's scale, no Inf/NaN."""
    L = capi.lib()
    assert L.mpimg_get_value_range() == capi.RANGE_UNIT
    base = synth.noise_f32(96, 640, 3, 7200)
    try:
        L.mpimg_set_value_range(capi.RANGE_ANY)
        assert L.mpimg_get_gauss_column() == 1          # no fp16 operands for undeclared ranges
        for scale in (65535.0, 1.0e6):
            a = (base * np.float32(scale)).astype(np.float32)
            a[10, 100, 1] = np.float32(scale)           # a sample at the top of the range
            got = dev(capi, a).apply("gaussian", 2.0).numpy()
            assert np.isfinite(got).all()
            assert np.abs(got - so.gaussian(a, 2.0)).max() <= TOL32 * scale
        # the chain executor follows the declaration too (batched launch)
        from millipyde_b200 import engine
        imgs = [(base * np.float32(70000.0)).astype(np.float32) for _ in range(3)]
        devs = [dev(capi, x) for x in imgs]
        engine.Chain([("gaussian", 2.0)], device=0).run(devs)
        for x, d in zip(imgs, devs):
            got = d.numpy()
            assert np.isfinite(got).all() and np.abs(got - so.gaussian(x, 2.0)).max() <= TOL32 * 70000.0
    finally:
        L.mpimg_set_value_range(capi.RANGE_UNIT)
    assert L.mpimg_get_gauss_column() == 0


def test_gaussian_accepts_every_sigma(capi):
    """The reference takes any sigma (src/millipyde_image.cpp:660-675).  Beyond the streaming buckets
    the shared-memory tile kernel serves until a tile plus its halo no longer fits (sigma ~ 9 for fp64,
    ~ 11 for fp32 RGB); beyond that, and beyond the 127-tap weight table (sigma > 15.9), the
    global-memory two-pass kernel runs the FULL support int(8 sigma + 0.5) like scipy does."""
    for sigma in (4.0, 9.5, 12.0, 20.0):
        for c in (1, 3):
            a = synth.noise_f32(70, 90, c, 7300 + c)
            got = dev(capi, a).apply("gaussian", sigma).numpy()
            assert np.abs(got - so.gaussian(a, sigma)).max() <= TOL32, (sigma, c)
        g = synth.noise_f32(61, 83, 1, 7310).astype(np.float64)
        got = dev(capi, g).apply("gaussian", sigma).numpy()
        assert np.abs(got - so.gaussian(g, sigma)).max() <= 1e-12, sigma
    # a batch through the executor takes the same route image by image
    from millipyde_b200 import engine
    imgs = [synth.noise_f32(40, 64, 3, 7320 + k) for k in range(3)]
    devs = [dev(capi, x) for x in imgs]
    engine.Chain([("brightness", 0.1), ("gaussian", 18.0)], device=0).run(devs)
    for x, d in zip(imgs, devs):
        assert np.abs(d.numpy() - so.apply_chain(x, [("brightness", 0.1), ("gaussian", 18.0)])).max() <= TOL32


def test_f32_gaussian_edge_cases(capi):
    a = synth.noise_f32(5, 7, 3, 1)                 # smaller than the kernel support
    assert np.abs(dev(capi, a).apply("gaussian", 2.0).numpy() - so.gaussian(a, 2.0)).max() <= TOL32
    a = synth.noise_f32(1, 300, 1, 2)               # single row
    assert np.abs(dev(capi, a).apply("gaussian", 2.0).numpy() - so.gaussian(a, 2.0)).max() <= TOL32
    a = synth.noise_f32(40, 40, 3, 3)
    assert np.array_equal(dev(capi, a).apply("gaussian", 0.0).numpy(), a)   # scipy: sigma 0 copies


@pytest.mark.parametrize("c", CHANNELS)
@pytest.mark.parametrize("angle", [30.0, 45.0, -17.5, 90.0, 180.0, 0.0, 360.0])
def test_f32_rotate_bilinear(capi, c, angle):
    for shape in [(97, 131), (128, 96)]:
        a = synth.noise_f32(*shape, c, 6000 + c)
        got = dev(capi, a).apply("rotate", angle).numpy()
        assert np.abs(got - so.rotate(a, angle)).max() <= TOL32, (shape, angle)


def test_f32_chain_config3(capi):
    """BASELINE config 3's chain, op by op (the fused path is tested separately)."""
    a = synth.noise_f32(135, 240, 3, 3000)
    chain = [("rotate", 30.0), ("fliplr",), ("adjust_gamma", 1.5, 1.0), ("gaussian", 2.0)]
    got = dev(capi, a).apply_chain(chain).numpy()
    assert np.abs(got - so.apply_chain(a, chain)).max() <= TOL32


# --------------------------------------------------------- RGBA8 (reference)
def rgba_inputs(charlie_small):
    return [synth.rgba8(512, 512, 1000), synth.rgba8(97, 131, 1001), charlie_small]


def test_config1_grey_transpose(capi, charlie_small):
    """BASELINE config 1: rgb2grey + transpose on 512 x 512 RGBA8 vs the test
    suite's oracle (tests/millipyde_tests.py:128-137)."""
    for a in rgba_inputs(charlie_small):
        d = dev(capi, a).apply("rgb2grey")
        assert d.dtype == np.float64 and d.shape == a.shape[:2]
        g = d.numpy()
        assert np.abs(g - so.rgb2grey(a)).max() < TOL64
        assert np.abs(g - rx.grey_u8(a)).max() < TOL64
        t = d.apply("transpose").numpy()
        assert np.array_equal(t, g.T)


def test_rgba8_index_ops_bit_exact(capi, charlie_small):
    for a in rgba_inputs(charlie_small):
        assert np.array_equal(dev(capi, a).apply("transpose").numpy(), rx.transpose(a))
        assert np.array_equal(dev(capi, a).apply("fliplr").numpy(), rx.fliplr(a))


def test_rgba8_pointwise_bit_exact(capi, charlie_small):
    for a in rgba_inputs(charlie_small):
        for gamma, gain in [(2.0, 1.0), (1.5, 1.0), (0.5, 1.0), (2.2, 0.9)]:
            got = dev(capi, a).apply("adjust_gamma", gamma, gain).numpy()
            assert np.array_equal(got, rx.adjust_gamma(a, gamma, gain)), (gamma, gain)
        # reference test oracle (skimage 0.18.2), tests/millipyde_tests.py:570-578
        assert np.array_equal(dev(capi, a).apply("adjust_gamma", 2.0, 1.0).numpy(),
                              so.adjust_gamma_rgba(a, 2.0, 1.0))
        for delta in [0.1, -0.1, 0.5, -0.5, 0.0]:
            assert np.array_equal(dev(capi, a).apply("brightness", delta).numpy(), rx.brightness(a, delta))
        assert np.array_equal(dev(capi, a).apply("colorize", 0.5, 1.5, 1.1).numpy(),
                              rx.colorize(a, 0.5, 1.5, 1.1))


def test_rgba8_gaussian_bit_exact(capi, charlie_small):
    # the sigmas walk through every effective-radius instantiation of the chain kernel (taps whose
    # (int)(255 * w) is 0 are not evaluated: radius 3 at sigma 1, 5 at sigma 2, all 8 from sigma ~3)
    for a in rgba_inputs(charlie_small):
        for sigma in [2.0, 1.0, 0.4, 1.3, 1.6, 2.4, 2.8, 3.5, 9.0]:
            assert np.array_equal(dev(capi, a).apply("gaussian", sigma).numpy(), rx.gaussian(a, sigma)), sigma


def test_rgba8_rotate_nearest(capi, charlie_small):
    for a in rgba_inputs(charlie_small):
        for angle in [45.0, 30.0, 10.0]:
            got = dev(capi, a).apply("rotate", angle).numpy()
            want = rx.rotate(a, angle)
            # device vs glibc sin/cos can differ in the last ulp: a source coordinate
            # within an ulp of an integer may truncate differently (oracle/ref_exact.c)
            frac = np.mean(np.any(got != want, axis=-1))
            assert frac < 1e-4, (angle, frac)


def test_reference_layout_rotates_through_the_staged_box(capi):
    """rotate_box_kernel (kernels/geometry.cuh): rows that are whole 16-byte vectors take the box-staged
    kernels, other widths the direct ones; both must give the rule's result -- nearest (RGBA8, and fp64
    under reference semantics) against the C restatement of the reference kernel, bilinear fp64 against
    the oracle -- on shapes with ragged tile edges and at angles whose footprint leaves the image."""
    L = capi.lib()
    for h, w in [(200, 256), (131, 100), (64, 36), (97, 131), (300, 260)]:
        a = synth.rgba8(h, w, 3000 + h)
        g = np.random.default_rng(h * 7 + w).random((h, w))
        for angle in [30.0, 45.0, 10.0, 90.0, 133.0, -75.0, 0.0, 180.0]:
            got = dev(capi, a).apply("rotate", angle).numpy()
            frac = np.mean(np.any(got != rx.rotate(a, angle), axis=-1))
            assert frac < 2e-3, ("rgba8", h, w, angle, frac)   # device vs glibc sin / cos: a coordinate an ulp from an integer
            got = dev(capi, g).apply("rotate", angle).numpy()
            assert np.abs(got - so.rotate(g, angle)).max() < 1e-9, ("f64 bilinear", h, w, angle)
            L.mpimg_set_semantics(capi.SEMANTICS_REFERENCE)
            try:
                got = dev(capi, g).apply("rotate", angle).numpy()
            finally:
                L.mpimg_set_semantics(capi.SEMANTICS_ORACLE)
            assert np.mean(got != rx.rotate(g, angle)) < 2e-3, ("f64 nearest", h, w, angle)


def test_f64_gaussian_tma_and_scalar_staging(capi):
    """gauss_f64_kernel stages its tile through the TMA unit when rows are whole 16-byte vectors (even
    width) and with per-thread loads otherwise; both rules (oracle: 33 taps at sigma 2; reference: 17
    float-expf taps + clamp) on shapes smaller than a tile, with ragged edges and several tiles."""
    L = capi.lib()
    for h, w in [(40, 50), (96, 32), (97, 33), (130, 96), (300, 262), (257, 129)]:
        g = np.random.default_rng(h + 1000 * w).random((h, w))
        got = dev(capi, g).apply("gaussian", 2.0).numpy()
        assert np.abs(got - so.gaussian(g, 2.0)).max() < 1e-12, ("oracle", h, w)
        got = dev(capi, g).apply("gaussian", 1.0).numpy()      # radius 8 under the oracle rule
        assert np.abs(got - so.gaussian(g, 1.0)).max() < 1e-12, ("oracle r8", h, w)
        L.mpimg_set_semantics(capi.SEMANTICS_REFERENCE)
        try:
            got = dev(capi, g).apply("gaussian", 2.0).numpy()
        finally:
            L.mpimg_set_semantics(capi.SEMANTICS_ORACLE)
        assert np.abs(got - rx.gaussian(g, 2.0)).max() < TOL64, ("reference", h, w)


# ------------------------------------------------------------- fp64 greyscale
def test_f64_ops_oracle_semantics(capi, charlie_small):
    g = so.rgb2grey(charlie_small)
    # reference tests: gaussian :547-555, gamma :558-568 (decimal=4 there)
    assert np.abs(dev(capi, g).apply("gaussian", 2.0).numpy() - so.gaussian(g, 2.0)).max() < 1e-9
    assert np.abs(dev(capi, g).apply("adjust_gamma", 2.0, 1.0).numpy() - so.adjust_gamma(g, 2.0, 1.0)).max() < TOL64
    assert np.abs(dev(capi, g).apply("brightness", 0.3).numpy() - so.brightness(g, 0.3)).max() < TOL64
    assert np.abs(dev(capi, g).apply("rotate", 30.0).numpy() - so.rotate(g, 30.0)).max() < 1e-9
    assert np.array_equal(dev(capi, g).apply("fliplr").numpy(), so.fliplr(g))
    assert np.array_equal(dev(capi, g).apply("transpose").numpy(), so.transpose(g))
    assert np.array_equal(dev(capi, g).apply("colorize", 2.0, 2.0, 2.0).numpy(), g)   # no-op on grey


def test_f64_gamma_every_exponent_form(capi):
    """pow64 (kernels/pointwise.cuh): multiplication / square-root forms, exp(g log v), and the
    library pow for zeros -- all against numpy's float64 power (clamped to [0, 1] like the kernel)."""
    rng = np.random.default_rng(77)
    g = rng.random((61, 83))
    g[::7, ::5] = 0.0
    g[3, 4] = 1.0
    g[5, 6] = 1e-300
    for gamma, gain in [(2.0, 1.0), (1.0, 0.9), (0.5, 1.0), (1.5, 1.0), (3.0, 1.0), (4.0, 1.0), (2.2, 1.0),
                        (0.45, 1.0), (7.3, 1.0), (1.5, 2.5), (0.1, 0.7)]:
        got = dev(capi, g).apply("adjust_gamma", gamma, gain).numpy()
        want = np.clip(so.adjust_gamma(g, gamma, gain), 0, 1)
        assert np.abs(got - want).max() < TOL64, (gamma, gain)


def test_f64_ops_reference_semantics(capi, charlie_small):
    g = so.rgb2grey(charlie_small)
    L = capi.lib()
    L.mpimg_set_semantics(capi.SEMANTICS_REFERENCE)
    try:
        assert np.abs(dev(capi, g).apply("gaussian", 2.0).numpy() - rx.gaussian(g, 2.0)).max() < TOL64
        assert np.abs(dev(capi, g).apply("adjust_gamma", 2.0, 1.0).numpy() - rx.adjust_gamma(g, 2.0, 1.0)).max() < 1e-6
        got = dev(capi, g).apply("rotate", 45.0).numpy()
        want = rx.rotate(g, 45.0)
        assert np.mean(got != want) < 1e-4
        # the reference's own test_long_pipeline chain on its own layouts (:349-395)
        chain = [("gaussian", 2.0), ("rgb2grey",), ("transpose",), ("transpose",), ("rotate", 45.0)]
        got = dev(capi, charlie_small).apply_chain(chain).numpy()
        want = rx.apply_chain(charlie_small, chain)
        assert np.mean(np.abs(got - want) > TOL64) < 1e-4
    finally:
        L.mpimg_set_semantics(capi.SEMANTICS_ORACLE)


def test_unsupported_layout_is_an_error(capi):
    a = np.zeros((8, 8, 2), np.float32)
    with pytest.raises(capi.MillipydeError):
        dev(capi, a).apply("fliplr")
    with pytest.raises(capi.MillipydeError):
        dev(capi, np.zeros((8, 8), np.int64)).apply("gaussian", 2.0)
