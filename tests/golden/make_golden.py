"""Generate golden vectors from the reference itself.

Runs the UNMODIFIED reference (built by oracle/build_ref.py into oracle/_ref/)
on a B200 and records the outputs of its own kernels on its own layouts (packed
RGBA8, fp64 greyscale).  The reference ships no golden vectors (every
expectation in its tests is computed at test time by scikit-image), so these
are the fixtures that pin oracle/ref_exact.c -- and through it the product --
to the reference's actual arithmetic.

    gpurun -- python tests/golden/make_golden.py gpurun_out/reference_outputs.npz
    cp gpurun_out/reference_outputs.npz tests/golden/

Inputs: a 160 x 200 crop of the reference's charlie_small.png fixture and a
seeded 97 x 131 RGBA8 noise image (tests/synth.py, seed 1001).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402
from tests import synth  # noqa: E402


def inputs():
    from PIL import Image
    charlie = np.asarray(Image.open(os.path.join(ROOT, "tests", "golden", "charlie_small.png")))
    return {"charlie": np.ascontiguousarray(charlie[200:360, 150:350]), "noise": synth.rgba8(97, 131, 1001)}


RGBA_CASES = [("transpose",), ("fliplr",), ("rotate", 45.0), ("rotate", 30.0), ("gaussian", 2.0), ("gaussian", 1.0),
              ("adjust_gamma", 2.0, 1.0), ("adjust_gamma", 1.5, 1.0), ("adjust_gamma", 0.5, 1.0),
              ("brightness", 0.1), ("brightness", -0.1), ("colorize", 0.5, 1.5, 1.1)]
GREY_CASES = [("transpose",), ("fliplr",), ("rotate", 45.0), ("gaussian", 2.0), ("adjust_gamma", 2.0, 1.0),
              ("brightness", 0.3)]
LONG_CHAIN = [("gaussian", 2.0), ("rgb2grey",), ("transpose",), ("transpose",), ("rotate", 45.0)]


def case_key(prefix, case):
    return prefix + "/" + "_".join(str(c) for c in case)


def main(out_path, hybrid=False):
    # --hybrid: the reference's own CPython host code linked against libmp_b200.so (oracle/build_ref.py
    # build_hybrid) instead of the reference's kernels; run with MILLIPYDE_SEMANTICS=reference
    ref = build_ref.load_hybrid() if hybrid else build_ref.load()
    out = {}
    for name, img in inputs().items():
        out[f"{name}/input"] = img
        d = ref.gpuimage(img)
        d.rgb2grey()
        grey = d.__array__()
        out[f"{name}/rgb2grey"] = grey
        for case in RGBA_CASES:
            d = ref.gpuimage(img)
            getattr(d, case[0])(*case[1:])
            out[case_key(f"{name}/rgba", case)] = d.__array__()
        for case in GREY_CASES:
            d = ref.gpuimage(img)
            d.rgb2grey()
            getattr(d, case[0])(*case[1:])
            out[case_key(f"{name}/grey", case)] = d.__array__()
        d = ref.gpuimage(img)
        for case in LONG_CHAIN:
            getattr(d, case[0])(*case[1:])
        out[f"{name}/long_chain"] = d.__array__()
    np.savez_compressed(out_path, **out)
    print(f"wrote {len(out)} arrays to {out_path}")


if __name__ == "__main__":
    argv = [a for a in sys.argv[1:] if a != "--hybrid"]
    main(argv[0] if argv else os.path.join(ROOT, "gpurun_out", "reference_outputs.npz"), hybrid="--hybrid" in sys.argv)
