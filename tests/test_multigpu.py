"""Multi-device behaviour (ports tests/millipyde_multigpu_tests.py, which hard-codes
2 devices; here any count >= 2): Device hand-off, cross-device clone, connected
pipelines with NVLink peer hand-off, Pipeline.run() spreading images over every
device, Generator spreading.  Skipped on a single-GPU box."""
import os

import numpy as np
import numpy.testing as npt
import pytest

os.environ.setdefault("MILLIPYDE_NO_DEVICE_OK", "1")
import millipyde_b200  # noqa: E402
from oracle import skimage_oracle as so  # noqa: E402
from tests import synth  # noqa: E402

mp = millipyde_b200.load_extension()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(mp.DEVICE_COUNT < 2, reason="needs >= 2 GPUs")]
DECIMAL_ERROR = 4


def test_device_count_matches():
    assert mp.device_count() == mp.DEVICE_COUNT >= 2


def test_device_handoff(charlie_small):
    grey = so.rgb2grey(charlie_small)
    with mp.Device(0):
        d = mp.gpuimage(charlie_small)
        assert mp.get_current_device() == 0 and d.device == 0
        with mp.Device(1):
            d.rgb2grey()                       # moves to the target device first
            assert mp.get_current_device() == 1 and d.device == 1
            npt.assert_almost_equal(grey, np.array(d), decimal=DECIMAL_ERROR)
        assert mp.get_current_device() == 0
    npt.assert_almost_equal(grey, np.array(d), decimal=DECIMAL_ERROR)


def test_clone_across_devices(charlie_small):
    with mp.Device(0):
        d = mp.gpuimage(charlie_small)
    with mp.Device(1):
        d2 = d.clone()
    assert d.device == 0 and d2.device == 1
    d2.rgb2grey()
    assert np.array_equal(np.array(d), charlie_small)
    npt.assert_almost_equal(so.rgb2grey(charlie_small), np.array(d2), decimal=DECIMAL_ERROR)
    with mp.Device(0):
        a = mp.gpuarray(np.arange(9).reshape(3, 3))
    with mp.Device(1):
        b = a.clone()
    assert np.array_equal(np.array(b), np.arange(9).reshape(3, 3)) and b.device == 1


def test_dual_pipelines_peer_handoff(charlie_small):
    want = np.transpose(so.rgb2grey(charlie_small))
    d = mp.gpuimage(charlie_small)
    p = mp.Pipeline([d], [mp.Operation("rgb2grey")], device=0)
    p2 = mp.Pipeline([], [mp.Operation("transpose")], device=1)
    p.connect_to(p2)
    p.run()
    assert d.device == 1
    npt.assert_almost_equal(want, np.array(d), decimal=DECIMAL_ERROR)


def test_dual_pipelines_unspecified_devices(charlie_small):
    grey = so.rgb2grey(charlie_small)
    d = mp.gpuimage(charlie_small)
    p = mp.Pipeline([d], [mp.Operation("rgb2grey")])
    p2 = mp.Pipeline([], [mp.Operation("transpose")])
    p3 = mp.Pipeline([], [mp.Operation("transpose")])
    p.connect_to(p2)
    p2.connect_to(p3)
    assert p.device != p2.device          # auto-assigned to different devices (src/gpupipeline.c:186-218)
    p.run()
    npt.assert_almost_equal(grey, np.array(d), decimal=DECIMAL_ERROR)


def test_connected_fp32_chain_config5_shape():
    """BASELINE config 5's pair: grey+transpose on one GPU -> gaussian+rotate on the next."""
    imgs = [synth.noise_f32(120, 640, 3, 5000 + k) for k in range(6)]
    dev = [mp.gpuimage(a) for a in imgs]
    a_ops = [mp.Operation("rgb2grey"), mp.Operation("transpose")]
    b_ops = [mp.Operation("gaussian", 2), mp.Operation("rotate", 30)]
    pa = mp.Pipeline(dev, a_ops, device=0)
    pb = mp.Pipeline([], b_ops, device=1)
    pa.connect_to(pb)
    pa.run()
    chain = [("rgb2grey",), ("transpose",), ("gaussian", 2.0), ("rotate", 30.0)]
    for a, d in zip(imgs, dev):
        assert d.device == 1
        assert np.abs(np.array(d) - so.apply_chain(a, chain)).max() <= 1e-5


def test_gaussian_writes_into_the_receivers_memory():
    """The sender's LAST segment writes straight into the receiving device's pool; when that segment
    is the streaming Gaussian its outputs leave through TMA bulk stores over NVLink.  Enough images
    for the chunked hand-off (> 2 x 32) and a shape with a partial last strip."""
    imgs = [synth.noise_f32(75, 250, 3, 6000 + k) for k in range(70)]
    dev = [mp.gpuimage(a) for a in imgs]
    pa = mp.Pipeline(dev, [mp.Operation("adjust_gamma", 1.5, 1), mp.Operation("gaussian", 2)], device=0)
    pb = mp.Pipeline([], [mp.Operation("fliplr")], device=1)
    pa.connect_to(pb)
    pa.run()
    chain = [("adjust_gamma", 1.5, 1.0), ("gaussian", 2.0), ("fliplr",)]
    for a, d in zip(imgs, dev):
        assert d.device == 1
        assert np.abs(np.array(d) - so.apply_chain(a, chain)).max() <= 1e-5


def test_pipeline_spreads_over_all_devices():
    n = 4 * mp.DEVICE_COUNT + 3
    imgs = [synth.noise_f32(64, 640, 3, 100 + k) for k in range(n)]
    with mp.Device(0):
        dev = [mp.gpuimage(a) for a in imgs]
    mp.Pipeline(dev, [mp.Operation("gaussian", 2), mp.Operation("fliplr")]).run()
    used = {d.device for d in dev}
    assert used == set(range(mp.DEVICE_COUNT))
    # blocks of 4 round-robin from the recommended device (src/gpupipeline.c:267-283)
    assert [d.device for d in dev[:8]] == [mp.best_device()] * 4 + [(mp.best_device() + 1) % mp.DEVICE_COUNT] * 4
    for a, d in zip(imgs, dev):
        assert np.abs(np.array(d) - so.apply_chain(a, [("gaussian", 2.0), ("fliplr",)])).max() <= 1e-5


def test_spreading_run_waits_for_every_move_before_it_launches():
    """The images of a spreading run start on device 0; every other device's shard moves its images
    over and must order its launches after ALL of those copies, not only the first (the hand-over
    event used to be recorded between the moves: with several devices pulling from device 0 at once
    the later copies lost the race and whole images came out wrong).  Many images, repeated."""
    n = 16 * mp.DEVICE_COUNT + 3
    imgs = [synth.noise_f32(64, 640, 3, 300 + k) for k in range(n)]
    want = [so.apply_chain(a, [("gaussian", 2.0), ("fliplr",)]) for a in imgs]
    for rep in range(5):
        with mp.Device(0):
            dev = [mp.gpuimage(a) for a in imgs]
        mp.Pipeline(dev, [mp.Operation("gaussian", 2), mp.Operation("fliplr")]).run()
        assert {d.device for d in dev} == set(range(mp.DEVICE_COUNT))
        for k, (w, d) in enumerate(zip(want, dev)):
            assert np.abs(np.array(d) - w).max() <= 1e-5, (rep, k, d.device)


def test_generator_spreads_and_keeps_order():
    base = [synth.noise_f32(48, 64, 3, 900 + k) for k in range(3)]
    dev = [mp.gpuimage(a) for a in base]
    g = mp.Generator(dev, [mp.Operation("adjust_gamma", 1.5, 1.0)], outputs=4 * mp.DEVICE_COUNT + 2)
    outs = list(g)
    assert len(outs) == 4 * mp.DEVICE_COUNT + 2
    assert len({o.device for o in outs}) == mp.DEVICE_COUNT
    for i, o in enumerate(outs):
        want = np.clip(so.adjust_gamma(base[i % 3], 1.5, 1.0), 0, 1)
        assert np.abs(np.array(o) - want).max() <= 1e-5
    g2 = mp.Generator(dev, [mp.Operation("fliplr")], device=1, outputs=3, return_to_host=True)
    for i, o in enumerate(g2):
        assert np.array_equal(o, base[i % 3][:, ::-1])
