"""The chain executor (mppipe_*, include/mp_pipeline.h) against the oracle:
batched launches, the fusion pass (fused == unfused == oracle), per-image coin
flips, ragged batches, the host streaming path, connected pipelines."""
import numpy as np
import pytest

from oracle import ref_exact as rx
from oracle import skimage_oracle as so
from tests import synth

pytestmark = pytest.mark.gpu
TOL32 = 1e-5


@pytest.fixture(scope="module")
def mp():
    from millipyde_b200 import capi, engine
    capi.initialize()
    capi.lib().mpimg_set_semantics(capi.SEMANTICS_ORACLE)

    class NS:
        pass
    ns = NS()
    ns.capi, ns.engine, ns.lib = capi, engine, capi.lib()
    return ns


CONFIG3 = [("rotate", 30.0), ("fliplr",), ("adjust_gamma", 1.5, 1.0), ("gaussian", 2.0)]


def test_batch_equals_oracle_config3_shape(mp):
    """BASELINE config 3's chain on a (small) batch of 1920-wide RGB rows."""
    imgs = [synth.noise_f32(108, 1920, 3, 3000 + k) for k in range(6)]
    dev = [mp.capi.DeviceImage(a) for a in imgs]
    ch = mp.engine.Chain(CONFIG3, device=0)
    ch.run(dev)
    # rotate + fliplr + gamma fold into one gather pass, then the Gaussian: 2 launches for the batch
    assert ch.last_segments == 2 and ch.last_launches == 2
    for a, d in zip(imgs, dev):
        assert np.abs(d.numpy() - so.apply_chain(a, CONFIG3)).max() <= TOL32


@pytest.mark.parametrize("c", [1, 3, 4])
def test_gather_fusion_matches_op_by_op(mp, c):
    chains = [
        [("fliplr",), ("brightness", 0.1)],
        [("rotate", 30.0), ("fliplr",), ("adjust_gamma", 1.5, 1.0)],
        [("adjust_gamma", 0.8, 0.9), ("fliplr",), ("rotate", -17.5), ("brightness", -0.05), ("fliplr",)],
        [("fliplr",), ("fliplr",), ("colorize", 0.9, 1.1, 1.0), ("rotate", 90.0)],
        [("brightness", 0.2), ("rotate", 45.0), ("rotate", 10.0), ("fliplr",)],     # second rotate starts a new segment
    ]
    for chain in chains:
        for shape in [(97, 131), (64, 96)]:
            imgs = [synth.noise_f32(*shape, c, 700 + k) for k in range(3)]
            dev = [mp.capi.DeviceImage(a) for a in imgs]
            ch = mp.engine.Chain(chain, device=0)
            ch.run(dev)
            single = mp.capi.DeviceImage(imgs[0])
            mp.engine.Chain(chain, device=0).run([single])            # un-batched path
            for a, d in zip(imgs, dev):
                want = so.apply_chain(a, chain)
                if c == 4:      # alpha is untouched by pointwise ops; index ops move it with the pixel
                    want = so.apply_chain(a, [op for op in chain])
                    alpha = so.apply_chain(a, [op for op in chain if op[0] in ("fliplr", "rotate", "transpose")])
                    want[..., 3] = alpha[..., 3]
                assert np.abs(d.numpy() - want).max() <= TOL32, chain
            assert np.array_equal(single.numpy(), dev[0].numpy())


def test_gaussian_batch_is_one_launch(mp):
    imgs = [synth.noise_f32(96, 640, 3, 2000 + k) for k in range(9)]
    dev = [mp.capi.DeviceImage(a) for a in imgs]
    ch = mp.engine.Chain([("gaussian", 2.0)], device=0)
    ch.run(dev)
    assert ch.last_launches == 1
    for a, d in zip(imgs, dev):
        assert np.abs(d.numpy() - so.gaussian(a, 2.0)).max() <= TOL32


def test_fusion_on_off_identical_and_fewer_launches(mp):
    chain = [("brightness", 0.1), ("adjust_gamma", 1.5, 1.0), ("colorize", 0.9, 1.1, 1.0), ("rgb2grey",),
             ("brightness", -0.05), ("transpose",)]
    imgs = [synth.noise_f32(64, 80, 3, 10 + k) for k in range(4)]
    results = {}
    for fused in (1, 0):
        mp.lib.mppipe_set_fusion(fused)
        dev = [mp.capi.DeviceImage(a) for a in imgs]
        ch = mp.engine.Chain(chain, device=0)
        ch.run(dev)
        results[fused] = ([d.numpy() for d in dev], ch.last_launches, ch.last_segments)
    mp.lib.mppipe_set_fusion(1)
    assert results[1][2] == 2 and results[0][2] == 6          # segments per group
    # fused: one batched launch for [brightness, gamma, colorize, grey, brightness], one for the transposes
    assert results[1][1] == 2 and results[0][1] == 6 * 4
    for f, u, a in zip(results[1][0], results[0][0], imgs):
        want = so.apply_chain(a, chain)
        assert f.shape == want.shape
        assert np.abs(f - want).max() <= TOL32 and np.abs(u - want).max() <= TOL32
        assert np.abs(f - u).max() <= 2e-7


def test_rgba8_pointwise_chain_fuses_bit_exactly(mp, charlie_small):
    chain = [("adjust_gamma", 2.0, 1.0), ("brightness", 0.1), ("colorize", 0.5, 1.5, 1.1), ("adjust_gamma", 0.5, 1.0)]
    imgs = [charlie_small, synth.rgba8(128, 160, 1000)]
    dev = [mp.capi.DeviceImage(a) for a in imgs]
    ch = mp.engine.Chain(chain, device=0)
    ch.run(dev)
    # one table kernel per image, plus the 768-entry table build the first time a program is seen
    assert ch.last_launches in (2, 3)
    for a, d in zip(imgs, dev):
        assert np.array_equal(d.numpy(), rx.apply_chain(a, chain))


def test_reference_long_pipeline_self_consistency(mp, charlie_small):
    """tests/millipyde_tests.py:349-395: eager chain == 8-image Pipeline."""
    chain = [("gaussian", 2.0), ("rgb2grey",), ("transpose",), ("transpose",), ("rotate", 45.0)]
    control = mp.capi.DeviceImage(charlie_small).apply_chain(chain).numpy()
    dev = [mp.capi.DeviceImage(charlie_small) for _ in range(8)]
    mp.engine.Chain(chain).run(dev)
    for d in dev:
        assert np.array_equal(d.numpy(), control)


def test_probability_is_per_image(mp):
    a = synth.noise_f32(32, 48, 3, 77)
    flipped = so.fliplr(a)
    mp.lib.mprand_seed(99)
    dev = [mp.capi.DeviceImage(a) for _ in range(64)]
    mp.engine.Chain([("fliplr", {"probability": 0.5})], device=0).run(dev)
    mp.lib.mprand_seed(0)
    n_flip = 0
    for d in dev:
        out = d.numpy()
        if np.array_equal(out, flipped):
            n_flip += 1
        else:
            assert np.array_equal(out, a)
    assert 16 <= n_flip <= 48


def _keyed(mp, chain, image, stage, slot, lo, hi):
    """The draw the executor (host or device) made for parameter `slot` of stage `stage` of image
    `image` in the chain's last run: mprand_keyed_double (include/mp_abi.h), Philox-4x32-10."""
    return mp.lib.mprand_keyed_double(mp.lib.mppipe_last_run_key(chain.ptr), image, stage, slot, lo, hi)


def test_random_chain_is_one_launch_per_segment_and_matches_oracle(mp):
    """Per-image parameter records: 70 images with their own brightness delta, sigma and colour
    multipliers run as a handful of launches (the Gaussians split by radius bucket and into sets of
    64), and every image equals the oracle evaluated with ITS draws.  The draws are predicted by
    evaluating the keyed generator (run key, image index, stage, slot) the executor uses."""
    n = 70
    imgs = [synth.noise_f32(40, 160, 3, 3000 + k) for k in range(n)]
    chain = [("random_brightness", -.2, .2), ("random_gaussian", .5, 2.),
             ("random_colorize", .5, 1.5, .5, 1.5, .5, 1.5), ("rgb2grey",), ("random_adjust_gamma", .5, 2., 1., 1.)]
    mp.lib.mprand_seed(1234)
    dev = [mp.capi.DeviceImage(a) for a in imgs]
    ch = mp.engine.Chain(chain, device=0)
    ch.run(dev)
    mp.lib.mprand_seed(0)
    draws = []      # stage k of the chain, parameter slots in declaration order
    for i in range(n):
        b = _keyed(mp, ch, i, 0, 0, -.2, .2)
        sg = _keyed(mp, ch, i, 1, 0, .5, 2.)
        col = tuple(_keyed(mp, ch, i, 2, s, .5, 1.5) for s in range(3))
        gam = (_keyed(mp, ch, i, 4, 0, .5, 2.), _keyed(mp, ch, i, 4, 1, 1., 1.))
        draws.append((b, sg, col, gam))
    # brightness: 1 launch; Gaussian: <= 5 buckets x 2 sets; colorize+grey+gamma: 1 per bucket group
    assert ch.last_launches <= 5 * (1 + 2 + 1), ch.last_launches
    assert ch.last_launches < n
    for a, d, (b, sg, col, gam) in zip(imgs, dev, draws):
        want = so.apply_chain(a, [("brightness", b), ("gaussian", sg), ("colorize", *col), ("rgb2grey",),
                                  ("adjust_gamma", *gam)])
        got = d.numpy()
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= TOL32


@pytest.mark.parametrize("c", [1, 3])
def test_random_rotate_records_match_eager_rotates(mp, c):
    """random_rotate alone (direct bilinear kernel, per-image angle table) and inside a gather
    segment (fliplr + rotate + brightness, per-image GatherVar records) against eager ops with
    the replayed angles."""
    n = 9
    imgs = [synth.noise_f32(48, 64, c, 3100 + k) for k in range(n)]
    for chain, replay in (
        ([("random_rotate", 0., 120.)], lambda ch, i: [("rotate", _keyed(mp, ch, i, 0, 0, 0., 120.))]),
        ([("fliplr",), ("random_rotate", 0., 120.), ("random_brightness", -.2, .2)],
         lambda ch, i: [("fliplr",), ("rotate", _keyed(mp, ch, i, 1, 0, 0., 120.)),
                        ("brightness", _keyed(mp, ch, i, 2, 0, -.2, .2))]),
    ):
        mp.lib.mprand_seed(4321)
        dev = [mp.capi.DeviceImage(a) for a in imgs]
        ch = mp.engine.Chain(chain, device=0)
        ch.run(dev)
        mp.lib.mprand_seed(0)
        eager_chains = [replay(ch, i) for i in range(n)]
        assert ch.last_launches == 2        # the record fill + the gather
        for a, d, ec in zip(imgs, dev, eager_chains):
            want = so.apply_chain(a, ec)
            assert np.abs(d.numpy() - want).max() <= TOL32
            eager = mp.capi.DeviceImage(a).apply_chain(ec).numpy()
            assert np.abs(d.numpy() - eager).max() <= 2e-6


def test_random_chain_fusion_off_is_image_by_image(mp):
    """With fusion off the same seeded chain runs one image at a time and agrees with the batched
    per-image-record run."""
    n = 6
    imgs = [synth.noise_f32(32, 96, 3, 3200 + k) for k in range(n)]
    chain = [("random_gaussian", .5, 2.), ("random_brightness", -.2, .2)]
    outs = {}
    for fused in (1, 0):
        mp.lib.mppipe_set_fusion(fused)
        mp.lib.mprand_seed(555)
        dev = [mp.capi.DeviceImage(a) for a in imgs]
        ch = mp.engine.Chain(chain, device=0)
        ch.run(dev)
        outs[fused] = ([d.numpy() for d in dev], ch.last_launches)
    mp.lib.mppipe_set_fusion(1)
    mp.lib.mprand_seed(0)
    assert outs[0][1] == 2 * n and outs[1][1] < outs[0][1]
    for f, u in zip(outs[1][0], outs[0][0]):
        assert np.array_equal(f, u)


def test_views_leave_sources_untouched_and_own_their_results(mp):
    """mppipe_run_views (the Generator's clone-free path): sources unchanged, results == the same
    chain on deep clones, for chains whose first segment is batched, eager, or absent."""
    imgs = [synth.noise_f32(48, 64, 3, 3300 + k) for k in range(5)]
    u8 = [synth.rgba8(32, 48, 3400 + k) for k in range(3)]
    cases = [
        (imgs, [("brightness", 0.1), ("gaussian", 1.0)]),                 # batched first segment
        (imgs, [("rotate", 20.0), ("rgb2grey",)]),                        # gather first
        (imgs, [("fliplr", {"probability": 0.5}), ("transpose", {"probability": 0.5})]),  # some images untouched
        (imgs, []),                                                       # nothing runs: deep copy
        (u8, [("adjust_gamma", 2.0, 1.0), ("rgb2grey",)]),                # eager reference-layout ops
    ]
    for arrays, chain in cases:
        src = [mp.capi.DeviceImage(a) for a in arrays]
        mp.lib.mprand_seed(31)
        clones = [d.clone() for d in src]
        mp.engine.Chain(chain, device=0).run(clones)
        mp.lib.mprand_seed(31)
        views = [d.view() for d in src]
        mp.engine.Chain(chain, device=0).run_views(views)
        mp.lib.mprand_seed(0)
        for a, d, c, v in zip(arrays, src, clones, views):
            assert np.array_equal(d.numpy(), a)                # source untouched
            assert v.obj.device_data != d.obj.device_data      # the view owns a buffer of its own
            assert np.array_equal(v.numpy(), c.numpy())
        for v in views:
            v.close()
        for d in src:                                          # and the sources are still alive
            assert d.numpy().shape == arrays[0].shape


def test_ragged_batch(mp):
    shapes = [(40, 64, 3), (97, 131, 3), (40, 64, 3), (33, 20, 1), (64, 64, 4)]
    imgs = [synth.noise_f32(h, w, c, 500 + i) for i, (h, w, c) in enumerate(shapes)]
    dev = [mp.capi.DeviceImage(a) for a in imgs]
    chain = [("gaussian", 1.0), ("fliplr",)]
    mp.engine.Chain(chain, device=0).run(dev)
    for a, d in zip(imgs, dev):
        assert np.abs(d.numpy() - so.apply_chain(a, chain)).max() <= TOL32


def test_empty_inputs_and_empty_chain(mp):
    mp.engine.Chain([("gaussian", 2.0)], device=0).run([])
    a = synth.noise_f32(16, 16, 3, 1)
    d = mp.capi.DeviceImage(a)
    mp.engine.Chain([], device=0).run([d])
    assert np.array_equal(d.numpy(), a)


def test_connected_pipelines_same_device(mp, charlie_small):
    """tests/millipyde_multigpu_tests.py:69-86 shape, on however many devices exist."""
    d = mp.capi.DeviceImage(charlie_small)
    p1 = mp.engine.Chain([("rgb2grey",)])
    p2 = mp.engine.Chain([("transpose",)])
    p3 = mp.engine.Chain([("transpose",)])
    p1.connect_to(p2)
    p2.connect_to(p3)
    p1.run([d])
    assert np.abs(d.numpy() - so.rgb2grey(charlie_small)).max() < 1e-12


def test_host_streaming_path(mp):
    imgs = [synth.noise_f32(120, 640, 3, 900 + k) for k in range(10)]
    ins = []
    for a in imgs:
        p = mp.engine.pinned_empty(a.shape, np.float32)
        p[...] = a
        ins.append(p)
    outs = [mp.engine.pinned_empty(a.shape, np.float32) for a in imgs]
    ch = mp.engine.Chain([("gaussian", 2.0), ("adjust_gamma", 1.5, 1.0)], device=0)
    meta = ch.run_host(ins, outs)
    for a, o, (shape, dt) in zip(imgs, outs, meta):
        assert shape == a.shape and dt == np.float32
        want = so.apply_chain(a, [("gaussian", 2.0), ("adjust_gamma", 1.5, 1.0)])
        assert np.abs(o - want).max() <= TOL32
    # shape-changing chain from pageable memory
    ch2 = mp.engine.Chain([("rgb2grey",), ("transpose",)], device=0)
    outs2 = [np.empty(a.shape, np.float32) for a in imgs]
    meta2 = ch2.run_host(imgs, outs2)
    for a, o, (shape, dt) in zip(imgs, outs2, meta2):
        assert shape == (a.shape[1], a.shape[0])
        got = o.reshape(-1)[: shape[0] * shape[1]].reshape(shape)
        assert np.abs(got - so.rgb2grey(a).T).max() <= TOL32
    for p in ins + outs:
        mp.engine.pinned_free(p)


def test_full_size_properties_4k(mp):
    """BASELINE config 2 at full size: properties that need no CPU oracle pass
    over 25M samples -- a constant image stays constant away from the border,
    and the blur is linear."""
    h, w, c = 2160, 3840, 3
    const = np.full((h, w, c), 0.5, np.float32)
    d = mp.capi.DeviceImage(const).apply("gaussian", 2.0).numpy()
    assert np.abs(d[16:-16, 16:-16] - 0.5).max() < 2e-6
    # border rows lose exactly the weight that falls outside (mode constant, cval 0)
    wts = so.gaussian_weights(2.0)
    inside = wts[16:].sum()
    assert abs(float(d[0, 1000, 1]) - 0.5 * inside) < 2e-6
    assert abs(float(d[0, 0, 0]) - 0.5 * inside * inside) < 2e-6
    a = synth.noise_f32(h, w, c, 2000)
    b = synth.smooth_f32(h, w, c)
    ga = mp.capi.DeviceImage(a).apply("gaussian", 2.0).numpy()
    gb = mp.capi.DeviceImage(b).apply("gaussian", 2.0).numpy()
    gab = mp.capi.DeviceImage((0.5 * a + 0.5 * b).astype(np.float32)).apply("gaussian", 2.0).numpy()
    assert np.abs(gab - (0.5 * ga + 0.5 * gb)).max() < 5e-6
    # and a strip of it against the oracle (rows 0..95: full width, all strips, top border)
    want = so.gaussian(a[:128], 2.0)[:96]
    assert np.abs(ga[:96] - want).max() <= TOL32


def test_full_frame_4k_single_image_against_oracle(mp):
    """BASELINE config 2's shape, one image: the eager call splits it into ~8 row chunks
    (mp_gauss_stream.cu launch_cr).  The WHOLE frame is compared with scipy -- every chunk boundary,
    every strip, all four borders."""
    a = synth.noise_f32(2160, 3840, 3, 2001)
    got = mp.capi.DeviceImage(a).apply("gaussian", 2.0).numpy()
    want = so.gaussian(a, 2.0)
    assert np.abs(got - want).max() <= TOL32
    assert np.abs(got[-16:] - want[-16:]).max() <= TOL32      # bottom border on its own


def test_full_frame_4k_batch_of_64_against_oracle(mp):
    """The benchmarked launch shape: >= 64 4K images in ONE launch through pointer tables (chunks = 1,
    1,152 items over 148 persistent CTAs), sources untouched (views).  Three images across the launch (one per seed) are compared with scipy over the whole frame, the rest with the image of the same
    seed (bit-identical: same kernel, same data)."""
    seeds = [synth.noise_f32(2160, 3840, 3, 2100 + k) for k in range(3)]
    base = [mp.capi.DeviceImage(s) for s in seeds]
    src = [base[k % 3].clone() for k in range(64)]
    views = [d.view() for d in src]
    ch = mp.engine.Chain([("gaussian", 2.0)], device=0)
    ch.run_views(views)
    assert ch.last_launches == 1
    wants = [so.gaussian(s, 2.0) for s in seeds]
    outs = {}
    for k in (0, 31, 62):
        outs[k] = views[k].numpy()
        assert np.abs(outs[k] - wants[k % 3]).max() <= TOL32, k
        assert np.abs(outs[k][-16:] - wants[k % 3][-16:]).max() <= TOL32, k
    for k in range(1, 64, 7):
        ref = {0: 0, 1: 31, 2: 62}[k % 3]
        assert np.array_equal(views[k].numpy(), outs[ref]), k
    # the sources were read, not written
    assert np.array_equal(src[5].numpy(), seeds[5 % 3])
    # a second pass over the same sources through re-armed views gives the same bits
    for v, d in zip(views, src):
        v.rebind(d)
    ch.run_views(views)
    assert np.array_equal(views[31].numpy(), outs[31])
    for v in views:
        v.rebind(None)
        v.close()
    for d in src + base:
        d.close()


@pytest.mark.parametrize("c", [1, 3, 4])
def test_pointwise_ops_fuse_into_the_gaussian(mp, c):
    """north star (2): pointwise -> gaussian -> pointwise is ONE launch and one HBM round trip.  The ops
    before the blur are applied to the rows as they land in shared memory (the blur's zero padding is
    padding of the *transformed* image: brightness must not leak into it), the ops after it to the
    finished rows.  Compared with the oracle and with the unfused execution of the same chain."""
    chains = [
        [("adjust_gamma", 1.5, 1.0), ("gaussian", 2.0), ("brightness", 0.1)],
        [("brightness", 0.25), ("gaussian", 2.0)],
        [("gaussian", 1.5), ("colorize", 0.9, 1.2, 0.5), ("brightness", -0.05)],
        [("gaussian", 0.7), ("colorize", 0.9, 1.2, 0.5), ("adjust_gamma", 0.8, 0.9)],
        # (no gamma < 1 behind the blur: x^0.5 near 0 amplifies the blur's ~3e-7 error past the contract)
        [("colorize", 1.3, 0.7, 1.1), ("adjust_gamma", 2.2, 1.0), ("gaussian", 1.3), ("brightness", -0.1),
         ("adjust_gamma", 1.5, 1.0)],
    ]
    for chain in chains:
        for h, w in [(75, 250 if c != 1 else 252), (140, 640), (33, 1284)]:
            if (w * c) % 4:
                continue
            imgs = [synth.noise_f32(h, w, c, 7000 + k) for k in range(5)]
            dev = [mp.capi.DeviceImage(a) for a in imgs]
            ch = mp.engine.Chain(chain, device=0)
            ch.run(dev)
            # what the fusion pass promises for this chain and layout (tests/test_fusion_plan.py pins the
            # rules: everything in front of the blur and every scale / add / clamp behind it is absorbed;
            # a gamma behind it, or an alpha-skipping op on RGBA, is one more pointwise launch)
            import ctypes as C
            buf = C.create_string_buffer(256)
            nseg = mp.lib.mppipe_plan(ch.ptr, 11, c, buf, len(buf))
            assert buf.value.decode().startswith("gauss") and nseg <= 2, buf.value
            if c == 3 and chain[-1][0] != "adjust_gamma":
                assert nseg == 1, buf.value
            assert ch.last_segments == nseg and ch.last_launches == nseg, (chain, buf.value, ch.last_launches)
            single = mp.capi.DeviceImage(imgs[0])
            ch1 = mp.engine.Chain(chain, device=0)
            ch1.run([single])                                   # a batch of one takes the same kernels
            assert ch1.last_launches == nseg
            mp.lib.mppipe_set_fusion(0)
            try:
                unf = [mp.capi.DeviceImage(a) for a in imgs[:2]]
                mp.engine.Chain(chain, device=0).run(unf)
            finally:
                mp.lib.mppipe_set_fusion(1)
            for k, (a, d) in enumerate(zip(imgs, dev)):
                got = d.numpy()
                want = so.apply_chain(a, chain)
                if c == 4:      # alpha is untouched by pointwise ops (as in the reference's RGBA kernels); the blur takes it
                    want[..., 3] = so.apply_chain(a, [op for op in chain if op[0] == "gaussian"])[..., 3]
                assert np.abs(got - want).max() <= TOL32, (chain, h, w)
                if k < 2:
                    assert np.abs(got - unf[k].numpy()).max() <= 2e-6
            assert np.array_equal(single.numpy(), dev[0].numpy())


def test_fused_gaussian_falls_back_op_by_op_outside_the_streaming_envelope(mp):
    """sigma = 3.3 has no streaming bucket: the GAUSS segment then runs its three parts one after the
    other (still correct, three launches).  250 x 3 floats per row is not a whole number of vectors:
    since round 2 the streaming kernel takes it (a brightness in front must not leak into the zero
    padding behind a row that ends inside a vector)."""
    chain = [("adjust_gamma", 1.5, 1.0), ("gaussian", 3.3), ("brightness", 0.1)]
    imgs = [synth.noise_f32(60, 128, 3, 7100 + k) for k in range(3)]
    dev = [mp.capi.DeviceImage(a) for a in imgs]
    ch = mp.engine.Chain(chain, device=0)
    ch.run(dev)
    assert ch.last_segments == 1
    for a, d in zip(imgs, dev):
        assert np.abs(d.numpy() - so.apply_chain(a, chain)).max() <= TOL32
    chain = [("brightness", 0.1), ("gaussian", 2.0)]
    imgs = [synth.noise_f32(60, 250, 3, 7200 + k) for k in range(3)]
    dev = [mp.capi.DeviceImage(a) for a in imgs]
    mp.engine.Chain(chain, device=0).run(dev)
    for a, d in zip(imgs, dev):
        assert np.abs(d.numpy() - so.apply_chain(a, chain)).max() <= TOL32


def test_fused_gaussian_with_per_image_programs_and_sigmas(mp):
    """A Generator-style stream: every image has its own gamma, sigma and brightness.  One GAUSS
    segment; the kernel reads image i's two programs and weight set from the per-image records."""
    n = 70
    imgs = [synth.noise_f32(40, 160, 3, 7300 + k) for k in range(n)]
    chain = [("random_adjust_gamma", .5, 2., 1., 1.), ("random_gaussian", 1.0, 1.25), ("random_brightness", -.2, .2)]
    mp.lib.mprand_seed(77)
    dev = [mp.capi.DeviceImage(a) for a in imgs]
    ch = mp.engine.Chain(chain, device=0)
    ch.run(dev)
    mp.lib.mprand_seed(0)
    draws = [((_keyed(mp, ch, i, 0, 0, .5, 2.), _keyed(mp, ch, i, 0, 1, 1., 1.)), _keyed(mp, ch, i, 1, 0, 1.0, 1.25),
              _keyed(mp, ch, i, 2, 0, -.2, .2)) for i in range(n)]
    assert ch.last_launches <= 2 * 3, ch.last_launches      # <= 2 radius buckets x (record fill + 2 sets of 64)
    for a, d, (gam, sg, b) in zip(imgs, dev, draws):
        want = so.apply_chain(a, [("adjust_gamma", *gam), ("gaussian", sg), ("brightness", b)])
        assert np.abs(d.numpy() - want).max() <= TOL32


def test_device_side_draws_equal_host_draws(mp):
    """when the images of a launch share the chain's shape, the host uploads one
    template + the stream indices and a kernel evaluates the counter-based generator and fills the
    per-image records (kernels/records.cuh).  mppipe_set_device_draws(0) evaluates the same keyed
    function on the host and uploads the records: pointwise / grey / fused-Gaussian results must be
    bit-identical (the gather's device-side cos/sin may differ from libm in the last place)."""
    n = 40
    imgs = [synth.noise_f32(48, 160, 3, 8000 + k) for k in range(n)]
    chains = {
        "pw": [("random_brightness", -.2, .2), ("random_adjust_gamma", .5, 2., .9, 1.)],
        "grey": [("random_colorize", .5, 1.5, .5, 1.5, .5, 1.5), ("rgb2grey",), ("random_brightness", -.1, .1)],
        "gauss": [("random_adjust_gamma", .8, 1.6, 1., 1.), ("gaussian", 2.0), ("random_brightness", -.2, .2)],
        "gather": [("random_brightness", -.1, .1), ("random_rotate", 0., 90.), ("fliplr",)],
    }
    assert mp.lib.mppipe_get_device_draws() == 1
    for name, chain in chains.items():
        outs = {}
        for on_device in (1, 0):
            mp.lib.mppipe_set_device_draws(on_device)
            mp.lib.mprand_seed(99)
            dev = [mp.capi.DeviceImage(a) for a in imgs]
            ch = mp.engine.Chain(chain, device=0)
            ch.run(dev)
            outs[on_device] = ([d.numpy() for d in dev], ch.last_launches)
        mp.lib.mppipe_set_device_draws(1)
        mp.lib.mprand_seed(0)
        assert outs[1][1] == outs[0][1] + 1, name            # the record-fill kernel is the only extra launch
        for x, y in zip(outs[1][0], outs[0][0]):
            if name == "gather":
                assert np.abs(x - y).max() <= 1e-6
            else:
                assert np.array_equal(x, y), name


def test_random_stream_does_not_depend_on_batching(mp):
    """Draws are keyed by (run, image index, stage, slot): image k of a stream gets the same parameters
    whether it runs in one batch of 24 or in three batches of 8 with the index base moved along."""
    imgs = [synth.noise_f32(32, 96, 3, 8100 + k) for k in range(24)]
    chain = [("random_brightness", -.2, .2, {"probability": 0.7}), ("random_gaussian", .5, 2.)]
    mp.lib.mprand_seed(5)
    whole = [mp.capi.DeviceImage(a) for a in imgs]
    ch = mp.engine.Chain(chain, device=0)
    mp.lib.mppipe_hold_run_key(ch.ptr)
    ch.run(whole)
    parts = [mp.capi.DeviceImage(a) for a in imgs]
    for b in range(3):
        mp.lib.mppipe_set_index_base(ch.ptr, 8 * b)
        ch.run(parts[8 * b: 8 * b + 8])
    mp.lib.mprand_seed(0)
    for w, p in zip(whole, parts):
        assert np.array_equal(w.numpy(), p.numpy())


def test_slab_outputs_are_ordinary_buffers(mp):
    """The outputs of a batched launch of small images are sub-blocks of one pool allocation
    (mp_devices.cpp: slabs).  They must behave like any buffer: outlive their siblings, be freed in
    any order and on any stream, be handed to eager ops (which replace them), and survive more
    batches allocating and retiring slabs around them."""
    imgs = [synth.noise_f32(64, 96, 3, 9000 + k) for k in range(40)]
    chain = [("brightness", 0.1), ("fliplr",)]
    want = [so.apply_chain(a, chain) for a in imgs]
    keep = []
    for rnd in range(4):
        dev = [mp.capi.DeviceImage(a) for a in imgs]
        mp.engine.Chain(chain, device=0).run(dev)
        # retire most of the batch out of order, keep a few from every round
        for k in list(range(39, -1, -3)) + list(range(1, 40, 3)):
            if k % 10 == rnd:
                continue
            dev[k].close()
        survivors = [(k, dev[k]) for k in range(40) if dev[k].ptr]
        keep.append(survivors)
        # an eager op on a survivor replaces its sub-block with a fresh buffer and retires the sub-block
        k0, d0 = survivors[0]
        d0.apply("adjust_gamma", 2.0, 1.0)
        assert np.abs(d0.numpy() - np.clip(so.adjust_gamma(want[k0], 2.0, 1.0), 0, 1)).max() <= TOL32
    for rnd, survivors in enumerate(keep):
        for k, d in survivors[1:]:
            assert np.abs(d.numpy() - want[k]).max() <= TOL32, (rnd, k)
            d.close()
    # views (the Generator's path) take their results from slabs too
    src = [mp.capi.DeviceImage(a) for a in imgs]
    views = [d.view() for d in src]
    ch = mp.engine.Chain(chain, device=0)
    ch.run_views(views)
    for rep in range(3):
        for v, d in zip(views, src):
            v.rebind(d)
        ch.run_views(views)
    for a, w, v in zip(imgs, want, views):
        assert np.abs(v.numpy() - w).max() <= TOL32
        v.rebind(None)
        v.close()
