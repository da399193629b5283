"""The fusion pass on the CPU: mppipe_plan (include/mp_pipeline.h) compiles an Operation chain
into the segments the executor would launch -- one HBM round trip per image each -- without
touching a GPU.  Pins the legality rules of DESIGN.md section 5 (reference IR:
MPRunnable[] built by src/gpupipeline.c:152-161; the reference runs one kernel per stage)."""
import ctypes as C

import pytest

from millipyde_b200 import capi, engine

F32, U8, F64 = 11, 2, 12   # numpy type numbers (NPY_FLOAT, NPY_UBYTE, NPY_DOUBLE)


def plan(ops, typenum, channels, fusion=True):
    L = capi.lib()
    L.mppipe_set_fusion(1 if fusion else 0)
    try:
        chain = engine.Chain(ops)
        buf = C.create_string_buffer(1024)
        n = L.mppipe_plan(chain.ptr, typenum, channels, buf, len(buf))
        chain.close()
        return n, buf.value.decode()
    finally:
        L.mppipe_set_fusion(1)


CONFIG3 = [("rotate", 30.0), ("fliplr",), ("adjust_gamma", 1.5, 1.0), ("gaussian", 2.0)]


def test_config3_chain_is_two_round_trips():
    """BASELINE config 3: rotate -> fliplr -> gamma fuse into one gather pass (the flip and the gamma
    are applied after the resample), the Gaussian is a stencil segment of its own."""
    assert plan(CONFIG3, F32, 3) == (2, "gather(-,rotate,fliplr;|adjust_gamma);gaussian")
    assert plan(CONFIG3, F32, 1)[0] == 2 and plan(CONFIG3, F32, 4)[0] == 2


def test_fusion_off_is_one_segment_per_stage_like_the_reference():
    assert plan(CONFIG3, F32, 3, fusion=False) == (4, "rotate;fliplr;adjust_gamma;gaussian")


def test_pointwise_ops_compose_into_one_program():
    ops = [("brightness", 0.1), ("adjust_gamma", 1.5, 1.0), ("colorize", 1.0, 0.5, 1.0)]
    assert plan(ops, F32, 3) == (1, "pw(brightness,adjust_gamma,colorize)")
    # RGBA8: composed byte tables; fp64 (reference greyscale layout): no program kernel, stage by stage
    assert plan(ops[:2], U8, 4) == (1, "u8(brightness,adjust_gamma)")
    assert plan(ops[:2], F64, 1) == (2, "brightness;adjust_gamma")


def test_grey_absorbs_the_pointwise_ops_around_it():
    ops = [("brightness", 0.1), ("rgb2grey",), ("adjust_gamma", 2.0, 1.0), ("transpose",)]
    assert plan(ops, F32, 3) == (2, "grey(brightness|adjust_gamma);transpose")


def test_gaussian_absorbs_the_pointwise_ops_around_it():
    """north star (2): a chain of pointwise and stencil ops is ONE HBM round trip -- the ops before the
    blur run on the rows as they land in shared memory, the ops after it on the finished rows
    (the reference: one kernel + one stream sync per stage, src/gpupipeline.c:373-392)."""
    ops = [("adjust_gamma", 1.5, 1.0), ("gaussian", 2.0), ("brightness", 0.1)]
    for c in (1, 3):
        assert plan(ops, F32, c) == (1, "gauss(adjust_gamma|brightness)")
    assert plan(ops[:2], F32, 3) == (1, "gauss(adjust_gamma|)")
    assert plan(ops[1:], F32, 3) == (1, "gauss(|brightness)")
    assert plan([("gaussian", 2.0)], F32, 3) == (1, "gaussian")          # a bare blur keeps its own launch path
    # Behind the blur only ops that compose to "scale per channel, add, clamp" are absorbed (the scale
    # moves in front of the blur, the add-and-clamp runs on the accumulators); a gamma there is not
    # affine and rides in front of the NEXT blur instead
    two = [("brightness", 0.1), ("gaussian", 2.0), ("adjust_gamma", 2.0, 1.0), ("gaussian", 1.0), ("brightness", -0.1)]
    assert plan(two, F32, 3) == (2, "gauss(brightness|);gauss(adjust_gamma|brightness)")
    assert plan([("gaussian", 2.0), ("colorize", 0.9, 1.1, 1.0), ("brightness", 0.1)], F32, 3) == (1, "gauss(|colorize,brightness)")
    # ... brightness then colorize would need a different offset per channel: the colorize is left behind
    assert plan([("gaussian", 2.0), ("brightness", 0.1), ("colorize", 0.9, 1.1, 1.0)], F32, 3) == \
        (2, "gauss(|brightness);pw(colorize)")
    assert plan([("gaussian", 2.0), ("adjust_gamma", 2.0, 1.0)], F32, 3) == (2, "gaussian;pw(adjust_gamma)")
    # RGBA float: brightness / colorize skip alpha, so they are not one map for all channels
    assert plan(ops, F32, 4) == (2, "gauss(adjust_gamma|);pw(brightness)")
    # the reference layouts (fp64 grey, RGBA8) have no fused stencil
    assert plan(ops, F64, 1) == (3, "adjust_gamma;gaussian;brightness")
    assert plan(ops, U8, 4) == (3, "adjust_gamma;gaussian;brightness")
    assert plan(ops, F32, 3, fusion=False) == (3, "adjust_gamma;gaussian;brightness")


def test_two_rotates_do_not_compose():
    """A second resample starts a new gather segment; pointwise ops between them ride on the first."""
    ops = [("rotate", 10.0), ("brightness", 0.1), ("rotate", 20.0)]
    assert plan(ops, F32, 1) == (2, "gather(-,rotate,-;|brightness);gather(-,rotate,-;|)")


def test_flip_before_a_rotate_is_applied_to_the_staged_source():
    n, text = plan([("adjust_gamma", 2.0, 1.0), ("fliplr",), ("rotate", 45.0), ("brightness", 0.2)], F32, 3)
    assert (n, text) == (1, "gather(fliplr,rotate,-;adjust_gamma|brightness)")


def test_index_ops_break_rgba8_tables_and_reference_layouts_stay_unfused():
    ops = [("brightness", 0.1), ("adjust_gamma", 1.5, 1.0), ("fliplr",), ("colorize", 1.0, 0.5, 1.0)]
    assert plan(ops, U8, 4) == (3, "u8(brightness,adjust_gamma);fliplr;colorize")
    ref = [("rgb2grey",), ("transpose",), ("gaussian", 2.0), ("rotate", 30.0)]
    assert plan(ref, U8, 4) == (4, "rgb2grey;transpose;gaussian;rotate")


def test_random_stages_plan_like_their_operators():
    ops = [("random_brightness", -0.2, 0.2), ("random_gaussian", 0.5, 2.0), ("random_rotate", 0.0, 120.0)]
    capi.lib().mprand_seed(11)
    try:
        assert plan(ops, F32, 3) == (2, "gauss(brightness|);gather(-,rotate,-;|)")
    finally:
        capi.lib().mprand_seed(0)


def test_probabilities_are_ignored_and_bad_arguments_rejected():
    ops = [("fliplr", {"probability": 0.01}), ("gaussian", 2.0, {"probability": 0.01})]
    assert plan(ops, F32, 3) == (2, "fliplr;gaussian")      # a lone flip keeps its streaming kernel
    assert plan([], F32, 3) == (0, "")
    assert plan(CONFIG3, 7, 3)[0] == -1          # int64 images have no operators
    assert plan(CONFIG3, F32, 2)[0] == -1        # fp32 takes 1, 3 or 4 channels
    L = capi.lib()
    chain = engine.Chain(CONFIG3)
    small = C.create_string_buffer(8)
    assert L.mppipe_plan(chain.ptr, F32, 3, small, len(small)) == -1    # buffer too small: nothing written
    assert L.mppipe_plan(None, F32, 3, small, len(small)) == -1
    chain.close()
