"""The drop-in Python surface (`millipyde` extension).  The CPU half ports the
reference's API-contract tests (tests/millipyde_tests.py:42-49, :69-85, :140-202,
:259-309: constructors raise ValueError/TypeError with the reference's messages);
the GPU half ports its oracle-comparison tests, with the scikit-image calls
restated by oracle/skimage_oracle.py and charlie_small.png standing in for the
missing charlie.png."""
import os
import subprocess
import sys

import numpy as np
import numpy.testing as npt
import pytest

os.environ.setdefault("MILLIPYDE_NO_DEVICE_OK", "1")   # lets the CPU half import on a GPU-less host
import millipyde_b200  # noqa: E402

mp = millipyde_b200.load_extension()
Operation = mp.Operation
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECIMAL_ERROR = 4


# ----------------------------------------------------------------------------- CPU half
def test_import_without_gpu_is_an_import_error():
    if mp.DEVICE_COUNT > 0:
        pytest.skip("a GPU is present")
    env = dict(os.environ)
    env.pop("MILLIPYDE_NO_DEVICE_OK", None)
    r = subprocess.run([sys.executable, "-c", "import millipyde_b200; millipyde_b200.load_extension()"],
                       cwd=ROOT, env=env, capture_output=True, text=True)
    assert r.returncode != 0
    assert "ImportError: GPU runtime failed while querying the device count" in r.stderr


def test_module_surface():
    for name in ("gpuarray", "gpuimage", "Operation", "Pipeline", "Generator", "Device", "device_count",
                 "get_current_device", "best_device", "image_from_path", "images_from_path", "DEVICE_COUNT"):
        assert hasattr(mp, name), name
    assert issubclass(mp.gpuimage, mp.gpuarray)
    assert sys.modules["millipyde"] is mp
    methods = ["rgb2grey", "rgb2gray", "rgba2grey", "rgba2gray", "transpose", "fliplr", "rotate", "gaussian",
               "brightness", "adjust_gamma", "colorize", "random_rotate", "random_gaussian",
               "random_brightness", "random_adjust_gamma", "random_colorize", "clone"]
    for m in methods:
        assert hasattr(mp.gpuimage, m), m
    assert mp.device_count() == mp.DEVICE_COUNT


def test_create_invalid_gpuarray():
    with pytest.raises(ValueError, match="numeric array"):
        mp.gpuarray(None)
    with pytest.raises(TypeError):
        mp.gpuarray()
    with pytest.raises(ValueError):
        mp.gpuarray(np.array(["a", "b"]))


def test_create_invalid_gpuimage():
    with pytest.raises(ValueError, match="2 dimensional"):
        mp.gpuimage(np.array([1, 2, 3, 4]))
    with pytest.raises(ValueError):
        mp.gpuimage(np.zeros((3, 2, 2, 2)))
    with pytest.raises(ValueError):
        mp.gpuimage(None)


def test_no_device_is_loud():
    if mp.DEVICE_COUNT > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="No usable CUDA device"):
        mp.gpuarray(np.array([1, 2, 3, 4]))


def test_create_invalid_operation():
    def do_nothing():
        pass
    with pytest.raises(ValueError, match="runnable/callable"):
        Operation()
    for bad in (7, 1, 0, -1, 1.0, 0.0, 1.5, "x"):
        with pytest.raises(ValueError, match="float probability between 0 and 1"):
            Operation(do_nothing, probability=bad)
    with pytest.raises(ValueError, match="only include one named argument"):
        Operation(do_nothing, chance=0.5)


def test_create_and_run_operation():
    def do_nothing():
        pass
    assert Operation(do_nothing) is not None
    op = Operation(do_nothing, probability=.6)
    assert abs(op.probability - .6) < 1e-15
    assert Operation(lambda x, y: x + y, 4, 6).run() == 10
    with pytest.raises(ValueError, match="string method name"):
        Operation(do_nothing).run_on(object())

    class Thing:
        def __init__(self):
            self.n = 0

        def bump(self, k):
            self.n += k
            return self.n
    t = Thing()
    assert Operation("bump", 3).run_on(t) == 3
    with pytest.raises(ValueError, match="could not be found"):
        Operation("nope").run_on(t)
    # probability: roughly that share of runs happens, the rest return None
    mp.seed(1234)
    hits = sum(Operation("bump", 1, probability=.25).run_on(t) is not None for _ in range(400))
    mp.seed(0)
    assert 60 <= hits <= 140


def test_create_invalid_pipeline():
    assert mp.Pipeline([], []) is not None
    for args in ((1, []), ([], 1), (np.array([1, 2, 3]), []), ([], np.array([1, 2, 3]))):
        with pytest.raises(ValueError):
            mp.Pipeline(*args)
    ops = [Operation("rgb2grey"), Operation("transpose")]
    with pytest.raises(ValueError, match="integer device"):
        mp.Pipeline([], ops, device=5.3)
    with pytest.raises(ValueError, match="integer device"):
        mp.Pipeline([], ops, device="nah")
    with pytest.raises(ValueError, match="only include one named argument"):
        mp.Pipeline([], ops, device=2, unused="test")
    with pytest.raises(ValueError, match="useable"):
        mp.Pipeline([], ops, device=1000)
    with pytest.raises(ValueError):
        mp.Pipeline([], [])  .__init__([], [], [], [])
    with pytest.raises(ValueError, match="GPU compatible"):
        mp.Pipeline([np.zeros((2, 2))], ops)
    with pytest.raises(ValueError):
        mp.Pipeline([], [Operation("brightness", 1.5)])     # |delta| must be < 1
    with pytest.raises(ValueError):
        mp.Pipeline([], [Operation("colorize", -1.0, 1.0, 1.0)])


def test_create_invalid_generator():
    ops = [Operation("rgb2grey")]
    with pytest.raises(ValueError, match="list of inputs or a path"):
        mp.Generator(3, ops)
    with pytest.raises(ValueError, match="list of Operations"):
        mp.Generator([], 3)
    with pytest.raises(ValueError, match="integer device"):
        mp.Generator([], ops, device="x")
    with pytest.raises(ValueError, match="number of outputs"):
        mp.Generator([], ops, outputs=-2)
    with pytest.raises(ValueError, match="boolean"):
        mp.Generator([], ops, return_to_host=1)
    with pytest.raises(ValueError, match="named arguments"):
        mp.Generator([], ops, bogus=1)
    g = mp.Generator([], ops, outputs=0)
    assert list(g) == []


# ----------------------------------------------------------------------------- GPU half
gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def so():
    from oracle import skimage_oracle
    return skimage_oracle


@pytest.fixture(scope="module")
def charlie_path():
    return os.path.join(ROOT, "tests", "golden", "charlie_small.png")


@gpu
def test_create_gpuarray_round_trips():
    a = np.array([1, 2, 3, 4])
    assert np.array_equal(a, np.array(mp.gpuarray(a)))
    assert np.array_equal(a, np.array(mp.gpuarray([1, 2, 3, 4])))
    assert not np.array_equal(a, np.array(mp.gpuarray([4, 3, 2, 1])))
    b = np.array([[1, 2, 3, 4], [4, 5, 6, 7]])
    assert np.array_equal(b, np.array(mp.gpuimage(b)))
    c = np.array([[[1, 2, 3], [4, 5, 6]], [[1, 2, 3], [4, 5, 6]]])
    g = mp.gpuimage(c)
    assert np.array_equal(c, np.array(g)) and g.shape == (2, 2, 3) and g.dtype == c.dtype
    assert np.array(g, dtype=np.float32).dtype == np.float32


@gpu
def test_np_functions_that_are_device_operators_stay_on_the_device():
    """np.fliplr / np.transpose of a gpuimage run the CUDA kernels on a device-side copy and return a
    gpuimage; the source is untouched; anything numpy defines differently (3-D
    transpose without axes reverses ALL axes) still takes the host path with numpy's semantics."""
    from tests import synth
    a = synth.noise_f32(37, 52, 3, 42)
    g = mp.gpuimage(a)
    f = np.fliplr(g)
    assert type(f) is mp.gpuimage and np.array_equal(np.array(f), np.fliplr(a))
    t = np.transpose(g, (1, 0, 2))
    assert type(t) is mp.gpuimage and t.shape == (52, 37, 3) and np.array_equal(np.array(t), np.transpose(a, (1, 0, 2)))
    t = np.transpose(g, axes=(1, 0, 2))
    assert type(t) is mp.gpuimage and np.array_equal(np.array(t), np.transpose(a, (1, 0, 2)))
    assert np.array_equal(np.array(g), a)                                   # numpy functions do not mutate
    host = np.transpose(g)                                                  # numpy: axes reversed -> (3, 52, 37)
    assert isinstance(host, np.ndarray) and np.array_equal(host, np.transpose(a))
    grey = synth.noise_f32(20, 33, 1, 43).astype(np.float64)
    t2 = np.transpose(mp.gpuimage(grey))
    assert type(t2) is mp.gpuimage and np.array_equal(np.array(t2), grey.T)
    rgba = np.random.default_rng(5).integers(0, 256, (16, 24, 4), dtype=np.uint8)
    assert np.array_equal(np.array(np.fliplr(mp.gpuimage(rgba))), np.fliplr(rgba))
    assert isinstance(np.flipud(g), np.ndarray)                             # not a device operator: host


@gpu
def test_np_ufuncs_that_are_pointwise_kernels_stay_on_the_device():
    """__array_ufunc__ (the reference's is a printing stub, src/gpuarray.c:147-191):
    np.add / subtract / multiply / power / clip / maximum / minimum with scalar operands (multiply also
    with three per-channel factors) on a float32 gpuimage run the pointwise kernel on a device-side copy
    and return a gpuimage with numpy's exact semantics -- no clamp, float32 arithmetic with the scalar
    cast to float32 (NEP 50), source untouched.  Anything else takes the host path."""
    from tests import synth
    a = synth.noise_f32(37, 52, 3, 44) * np.float32(1.7) - np.float32(0.2)        # leaves [0, 1] on both sides
    g = mp.gpuimage(a)
    n0 = mp.launch_count()
    cases = [
        (np.add(g, 0.3), a + np.float32(0.3)),
        (np.add(0.3, g), np.float32(0.3) + a),
        (g + 0.25, a + np.float32(0.25)) if hasattr(g, "__add__") else (np.add(g, 0.25), a + np.float32(0.25)),
        (np.subtract(g, 0.1), a - np.float32(0.1)),
        (np.multiply(g, 1.5), a * np.float32(1.5)),
        (np.multiply(g, np.float32(0.5)), a * np.float32(0.5)),
        (np.multiply(g, np.array([0.5, 1.5, 1.1], np.float32)), a * np.array([0.5, 1.5, 1.1], np.float32)),
        (np.multiply(np.array([0.5, 1.5, 1.1], np.float32), g), np.array([0.5, 1.5, 1.1], np.float32) * a),
        (np.clip(g, 0.0, 1.0), np.clip(a, 0.0, 1.0)),
        (np.maximum(g, 0.5), np.maximum(a, np.float32(0.5))),
        (np.minimum(0.5, g), np.minimum(np.float32(0.5), a)),
    ]
    assert mp.launch_count() - n0 >= len(cases)                  # kernels ran: nothing came from the host
    for k, (got, want) in enumerate(cases):
        assert type(got) is mp.gpuimage and got.shape == a.shape, k
        h = np.array(got)
        assert h.dtype == np.float32 and np.array_equal(h, want), k
    p = np.array(np.power(mp.gpuimage(np.abs(a)), 1.5))
    assert p.dtype == np.float32 and np.allclose(p, np.power(np.abs(a), np.float32(1.5)), rtol=2e-6, atol=1e-7)
    assert np.array_equal(np.array(g), a)                         # ufuncs do not mutate
    # composition stays on the device: brightness as numpy spells it
    b = np.clip(np.add(g, 0.2), 0.0, 1.0)
    assert type(b) is mp.gpuimage and np.array_equal(np.array(b), np.clip(a + np.float32(0.2), 0, 1))
    # host path (numpy semantics preserved): float64 scalar types promote, array operands, other ufuncs,
    # out=/dtype= arguments, non-float32 layouts
    r = np.add(g, np.float64(0.3))
    assert isinstance(r, np.ndarray) and r.dtype == np.float64 and np.array_equal(r, a + np.float64(0.3))
    r = np.add(g, a)
    assert isinstance(r, np.ndarray) and np.array_equal(r, a + a)
    r = np.multiply(g, [0.5, 1.5, 1.1])                           # a list is a float64 array to numpy
    assert isinstance(r, np.ndarray) and r.dtype == np.float64
    assert isinstance(np.sqrt(mp.gpuimage(np.abs(a))), np.ndarray)
    assert isinstance(np.add(g, 0.3, dtype=np.float64), np.ndarray)
    rgba = np.random.default_rng(5).integers(0, 256, (16, 24, 4), dtype=np.uint8)
    r = np.add(mp.gpuimage(rgba), 1)
    assert isinstance(r, np.ndarray) and np.array_equal(r, rgba + 1)
    grey = synth.noise_f32(20, 33, 1, 45)
    r = np.multiply(mp.gpuimage(grey), 2.0)
    assert type(r) is mp.gpuimage and np.array_equal(np.array(r), grey * np.float32(2.0))
    r = np.multiply(mp.gpuimage(grey[:, :3].copy()), [1.0, 2.0, 3.0])          # broadcasting over a 2-D image: host
    assert isinstance(r, np.ndarray)


@gpu
def test_elementwise_ops_chain_and_fuse_in_a_pipeline():
    """mpimg_elementwise has the MPFunc signature: through the C ABI it is a pipeline stage like the
    eight operators and fuses with them (here into the Gaussian's launch)."""
    import ctypes as C
    from millipyde_b200 import capi, engine
    from oracle import skimage_oracle as so
    from tests import synth
    L = capi.lib()
    imgs = [synth.noise_f32(64, 160, 3, 50 + k) for k in range(4)]
    dev = [capi.DeviceImage(x) for x in imgs]
    arr = (capi.MPRunnable * 3)()
    ew1 = capi.ElementwiseArgs(5, 0.5, 0.0, 0.0, 0.0)          # multiply by 0.5
    g = capi.GaussianArgs(2.0)
    ew2 = capi.ElementwiseArgs(4, 0.25, 0.0, 0.0, 0.0)         # add 0.25
    for k, (fn, a) in enumerate((("mpimg_elementwise", ew1), ("mpimg_gaussian", g), ("mpimg_elementwise", ew2))):
        arr[k].func = C.cast(getattr(L, fn), C.c_void_p)
        arr[k].args = C.cast(C.pointer(a), C.c_void_p)
        arr[k].probability = -1.0
    buf = C.create_string_buffer(256)
    pipe = L.mppipe_create(arr, 3, 0)
    assert L.mppipe_plan(pipe, 11, 3, buf, len(buf)) == 1 and buf.value == b"gauss(multiply|add)"
    objs = (C.POINTER(capi.MPObjData) * len(dev))(*[d.ptr for d in dev])
    assert L.mppipe_run(pipe, objs, len(dev)) == 0
    assert L.mppipe_last_launches(pipe) == 1
    for x, d in zip(imgs, dev):
        want = so.gaussian(x * np.float32(0.5), 2.0) + 0.25
        assert np.abs(d.numpy() - want).max() <= 1e-5
    L.mppipe_destroy(pipe)


@gpu
def test_np_function_protocol():
    lst = [[1, 2, 3], [4, 5, 6], [7, 8, 9]]
    want = np.transpose(np.array(lst))
    npt.assert_equal(want, np.transpose(mp.gpuarray(lst)))
    npt.assert_equal(want, np.transpose(mp.gpuimage(lst)))
    npt.assert_equal(np.add(mp.gpuarray(lst), 1), np.array(lst) + 1)       # ufunc protocol


@gpu
def test_rgb2grey_and_transpose(charlie_small, so):
    grey = so.rgb2grey(charlie_small)
    d = mp.gpuimage(charlie_small)
    npt.assert_almost_equal(charlie_small, np.array(d), decimal=DECIMAL_ERROR)
    d.rgb2grey()
    npt.assert_almost_equal(grey, np.array(d), decimal=DECIMAL_ERROR)
    d.transpose()
    npt.assert_almost_equal(np.transpose(grey), np.array(d), decimal=DECIMAL_ERROR)
    for alias in ("rgb2gray", "rgba2grey", "rgba2gray"):
        e = mp.gpuimage(charlie_small)
        getattr(e, alias)()
        assert np.array_equal(np.array(e), grey) or np.abs(np.array(e) - grey).max() < 1e-12


@gpu
def test_operation_grey_and_transpose(charlie_small, so):
    want = np.transpose(so.rgb2grey(charlie_small))
    d = mp.gpuimage(charlie_small)
    for op in [mp.Operation("rgb2grey"), mp.Operation("transpose")]:
        op.run_on(d)
    npt.assert_almost_equal(want, np.array(d), decimal=DECIMAL_ERROR)


@gpu
def test_pipeline_run(charlie_small, so):
    want = np.transpose(so.rgb2grey(charlie_small))
    a, b = mp.gpuimage(charlie_small), mp.gpuimage(charlie_small)
    p = mp.Pipeline([a, b], [mp.Operation("rgb2grey"), mp.Operation("transpose")])
    p.run()
    npt.assert_almost_equal(np.array(a), want, decimal=DECIMAL_ERROR)
    npt.assert_almost_equal(np.array(b), want, decimal=DECIMAL_ERROR)
    assert mp.Pipeline([a], [mp.Operation("transpose")], device=0).device == 0


@gpu
def test_long_pipeline(charlie_small):
    c = mp.gpuimage(charlie_small)
    c.gaussian(2)
    c.rgb2grey()
    c.transpose()
    c.transpose()
    c.rotate(45)
    control = np.array(c)
    imgs = [mp.gpuimage(charlie_small) for _ in range(8)]
    ops = [mp.Operation("gaussian", 2), mp.Operation("rgb2grey"), mp.Operation("transpose"),
           mp.Operation("transpose"), mp.Operation("rotate", 45)]
    mp.Pipeline(imgs, ops).run()
    for i in imgs:
        npt.assert_almost_equal(control, np.array(i), decimal=DECIMAL_ERROR)


@gpu
def test_clone(charlie_small, so):
    d = mp.gpuimage(charlie_small)
    d2 = d.clone()
    d2.rgb2grey()
    assert type(d2) is mp.gpuimage
    npt.assert_almost_equal(charlie_small, np.array(d), decimal=DECIMAL_ERROR)
    npt.assert_almost_equal(so.rgb2grey(charlie_small), np.array(d2), decimal=DECIMAL_ERROR)
    h = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]])
    a = mp.gpuarray(h)
    assert type(a.clone()) is mp.gpuarray and np.array_equal(np.array(a.clone()), h)


@gpu
def test_image_from_path(charlie_small, charlie_path, so):
    d = mp.image_from_path(charlie_path)
    npt.assert_almost_equal(charlie_small, np.array(d), decimal=DECIMAL_ERROR)
    d.rgb2grey()
    npt.assert_almost_equal(so.rgb2grey(charlie_small), np.array(d), decimal=DECIMAL_ERROR)


@pytest.fixture(scope="module")
def image_dir(tmp_path_factory):
    from PIL import Image
    from tests import synth
    d = tmp_path_factory.mktemp("imgs")
    imgs = {}
    for k, name in enumerate(["b.png", "a.png", "c.PNG", "d.bmp", "notes.txt", ".hidden.png"]):
        if name.endswith(".txt"):
            (d / name).write_text("not an image")
            continue
        arr = synth.rgba8(40 + 8 * k, 56, 100 + k)
        if name.endswith(".bmp"):
            Image.fromarray(arr[..., :3]).save(d / name)
            arr = arr[..., :3]
        else:
            Image.fromarray(arr).save(d / name)
        if not name.startswith("."):
            imgs[name] = arr
    return str(d), imgs


@gpu
def test_generator(image_dir, so):
    """tests/millipyde_tests.py:447-544: rgb2grey stream over a directory; here the
    order is sorted file names."""
    path, imgs = image_dir
    loaded = mp.images_from_path(path)
    # the reference's extension rule (src/gpuimage.c:803-823): ".hidden.png" counts, "notes.txt" does not
    valid = [n for n in sorted(os.listdir(path)) if os.path.splitext(n)[1].lower() in (".png", ".bmp")]
    assert len(loaded) == len(valid)
    g = mp.Generator(path, [mp.Operation("rgb2grey")])
    for i in range(2 * len(valid) + 1):
        out = np.array(next(g))
        name = valid[i % len(valid)]
        src = imgs.get(name)
        if src is None:
            continue
        want = so.rgb2grey(src) if src.shape[2] == 4 else src[..., :3].astype(np.float64) @ np.array(so.LUMA) / 255
        npt.assert_almost_equal(out, want, decimal=DECIMAL_ERROR)
    g2 = mp.Generator(path, [mp.Operation("rgb2grey")], return_to_host=True, outputs=3)
    outs = list(g2)
    assert len(outs) == 3 and all(isinstance(o, np.ndarray) for o in outs)


@gpu
def test_generator_with_python_callable_op(charlie_small, so):
    calls = []
    ops = [mp.Operation("rgb2grey"), mp.Operation(lambda: calls.append(1))]
    g = mp.Generator([mp.gpuimage(charlie_small)], ops, outputs=2, return_to_host=True)
    outs = list(g)
    assert len(outs) == 2 and len(calls) == 2
    npt.assert_almost_equal(outs[0], so.rgb2grey(charlie_small), decimal=DECIMAL_ERROR)


@gpu
def test_generator_random_augmentation_stream(so):
    """examples/augmentation_examples.py:13-21 on fp32 RGB, seeded: every output is one of the
    shapes the chain can produce and lies in [0, 1]."""
    from tests import synth
    base = [mp.gpuimage(synth.noise_f32(64, 96, 3, 4000 + k)) for k in range(3)]
    ops = [mp.Operation("transpose", probability=.2), mp.Operation("fliplr", probability=.2),
           mp.Operation("random_brightness", -.2, .2), mp.Operation("random_gaussian", .5, 2.),
           mp.Operation("random_colorize", [.5, 1.5], [.5, 1.5], [.5, 1.5], probability=.3),
           mp.Operation("rgb2grey", probability=.3), mp.Operation("random_rotate", 0., 120., probability=.5)]
    mp.seed(77)
    g = mp.Generator(base, ops, return_to_host=True, outputs=24, prefetch=8)
    shapes = set()
    for out in g:
        assert out.dtype == np.float32 and np.isfinite(out).all()
        assert out.min() >= 0 and out.max() <= 1 + 1e-6
        shapes.add(out.shape)
    mp.seed(0)
    assert shapes <= {(64, 96, 3), (96, 64, 3), (64, 96), (96, 64)} and len(shapes) >= 2


@gpu
def test_generator_async_lookahead_matches_the_synchronous_stream(so):
    """prefetch >= 32 runs one batch ahead of the consumer (batch k + 1 is on the device while batch k is
    being handed out); draws are keyed by output index, so the stream is the same as with the small
    synchronous look-ahead -- and a caller who mutates an input while iterating cannot disturb a batch
    in flight (its views borrow per-batch replicas), only later batches see the new input."""
    from tests import synth
    base_np = [synth.noise_f32(48, 64, 3, 4100 + k) for k in range(3)]
    ops = lambda: [mp.Operation("fliplr", probability=.5), mp.Operation("random_brightness", -.2, .2),
                   mp.Operation("random_gaussian", .5, 2.), mp.Operation("random_rotate", 0., 90., probability=.5)]
    outs = {}
    for pre in (8, 64):
        base = [mp.gpuimage(a) for a in base_np]
        mp.seed(123)
        g = mp.Generator(base, ops(), outputs=150, prefetch=pre, device=0)
        outs[pre] = [np.array(o) for o in g]
        mp.seed(0)
    assert len(outs[8]) == len(outs[64]) == 150
    for a, b in zip(outs[8], outs[64]):
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-6      # gather records: device vs device, same draws
    # host outputs through the asynchronous path
    base = [mp.gpuimage(a) for a in base_np]
    g = mp.Generator(base, [mp.Operation("adjust_gamma", 1.5, 1.0)], outputs=100, prefetch=32, return_to_host=True)
    for i, o in enumerate(g):
        assert isinstance(o, np.ndarray)
        assert np.abs(o - np.clip(so.adjust_gamma(base_np[i % 3], 1.5, 1.0), 0, 1)).max() <= 1e-5
    # mutation while a batch is in flight
    base = [mp.gpuimage(a) for a in base_np]
    g = mp.Generator(base, [mp.Operation("fliplr")], outputs=160, prefetch=40, device=0)
    seen = []
    for i, o in enumerate(g):
        if i == 5:
            base[0].brightness(0.5)            # batches 0 (being drained) and 1 (in flight) were cut from the old input
        seen.append(np.array(o))
    flipped_old, flipped_new = base_np[0][:, ::-1], np.clip(base_np[0] + np.float32(0.5), 0, 1)[:, ::-1]
    for i in range(0, 160, 3):
        want = flipped_old if i < 80 else flipped_new
        assert np.abs(seen[i] - want).max() <= 1e-6, i
    # an abandoned generator with a batch in flight cleans up
    g = mp.Generator(base, [mp.Operation("fliplr")], prefetch=64, device=0)
    next(g)
    del g
    mp.synchronize()


@gpu
def test_generator_host_outputs_are_page_locked_recycled_and_never_aliased(so):
    """return_to_host=True downloads a batch's outputs together into recycled page-locked ndarrays:
    every output a consumer still holds must keep its own contents while later batches reuse the
    blocks of the ones it dropped."""
    from tests import synth
    base = [mp.gpuimage(synth.noise_f32(48, 64, 3, 4100 + k)) for k in range(3)]
    want = [np.fliplr(np.array(b)) for b in base]
    g = mp.Generator(base, [mp.Operation("fliplr")], return_to_host=True, outputs=40, prefetch=4, device=0)
    kept = []
    for i, out in enumerate(g):
        assert isinstance(out, np.ndarray) and out.flags.c_contiguous and out.flags.writeable
        assert np.array_equal(out, want[i % 3])
        if i % 2 == 0:
            kept.append((i, out))          # half of the outputs stay alive, the others free their block
        else:
            out[...] = -1.0                # scribbling on a dropped output must not reach anyone else
    assert len(kept) == 20
    for i, out in kept:
        assert np.array_equal(out, want[i % 3])
    assert len({out.ctypes.data for _, out in kept}) == len(kept)


@gpu
def test_gaussian_and_gamma(charlie_small, so):
    grey = so.rgb2grey(charlie_small)
    d = mp.gpuimage(charlie_small)
    d.rgb2grey()
    d.gaussian(2)
    npt.assert_almost_equal(so.gaussian(grey, 2.0), np.array(d), decimal=4)     # :547-555
    assert np.abs(so.gaussian(grey, 2.0) - np.array(d)).max() < 1e-9            # and far inside 1e-5
    d = mp.gpuimage(charlie_small)
    d.rgb2grey()
    d.adjust_gamma(2, 1)
    npt.assert_almost_equal(so.adjust_gamma(grey, 2, 1), np.array(d), decimal=DECIMAL_ERROR)   # :558-568
    d = mp.gpuimage(charlie_small)
    d.adjust_gamma(2, 1)
    assert np.array_equal(so.adjust_gamma_rgba(charlie_small, 2, 1), np.array(d))              # :570-578


@gpu
def test_method_argument_errors(charlie_small):
    d = mp.gpuimage(charlie_small)
    with pytest.raises(ValueError):
        d.brightness(1.0)
    with pytest.raises(ValueError):
        d.colorize(-1, 1, 1)
    with pytest.raises(TypeError):
        d.adjust_gamma(2)          # both gamma and gain are required
    with pytest.raises(TypeError):
        d.rotate()
    with pytest.raises(RuntimeError, match="Unsupported image layout"):
        mp.gpuimage(np.zeros((4, 4), np.int64)).gaussian(2)


@gpu
def test_device_context_manager(charlie_small, so):
    assert mp.get_current_device() == mp.best_device()
    with mp.Device(0) as dev:
        assert dev == 0 and mp.get_current_device() == 0
        d = mp.gpuimage(charlie_small)
        assert d.device == 0
    with pytest.raises(KeyError):
        with mp.Device(0):
            raise KeyError("propagates")
    assert mp.get_current_device() == mp.best_device()
    with pytest.raises(ValueError):
        with mp.Device(1000):
            pass


@gpu
def test_pinned_empty_round_trip():
    a = mp.pinned_empty((64, 64, 3), np.float32)
    a[...] = np.random.default_rng(0).random((64, 64, 3), dtype=np.float32)
    d = mp.gpuimage(a)
    d.fliplr()
    assert np.array_equal(np.array(d), a[:, ::-1])
