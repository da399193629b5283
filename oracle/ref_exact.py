"""ctypes front-end of oracle/ref_exact.c (CPU oracle, part B).

TEST INFRASTRUCTURE ONLY -- see the header of ref_exact.c.  Builds the shared
object on first use with gcc (``-ffp-contract=off`` so only the explicit
``fma()`` calls fuse).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "ref_exact.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "libref_exact.so")

_lib = None


def build(force: bool = False) -> str:
    """Compile ref_exact.c -> oracle/_build/libref_exact.so (gcc, a second)."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c11",
             _SRC, "-o", _SO, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


_D = ctypes.c_double
_I = ctypes.c_int
_L = ctypes.c_long


def _pack(img: np.ndarray) -> np.ndarray:
    """H x W x 4 uint8 -> H x W uint32 view (little endian: R = bits 0..7)."""
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 4
    return np.ascontiguousarray(img).view(np.uint32).reshape(img.shape[:2])


def _unpack(px: np.ndarray) -> np.ndarray:
    return px.view(np.uint8).reshape(px.shape[0], px.shape[1], 4)


def grey_u8(img: np.ndarray) -> np.ndarray:
    img = np.ascontiguousarray(img)
    h, w, c = img.shape
    out = np.empty((h, w), np.float64)
    lib().ref_grey_u8(_p(img), _p(out), _I(w), _I(h), _I(c))
    return out


def _index_op(fn, img: np.ndarray, out_shape, *extra):
    img = np.ascontiguousarray(img)
    if img.ndim == 3:          # RGBA8: the pixel is one uint32
        src = _pack(img)
        out = np.empty(out_shape[:2], np.uint32)
        fn(_p(src), _p(out), _I(img.shape[1]), _I(img.shape[0]), _I(4), *extra)
        return _unpack(out)
    out = np.empty(out_shape, img.dtype)
    fn(_p(img), _p(out), _I(img.shape[1]), _I(img.shape[0]), _I(img.itemsize), *extra)
    return out


def transpose(img):
    shp = (img.shape[1], img.shape[0]) + tuple(img.shape[2:])
    return _index_op(lib().ref_transpose, img, shp)


def fliplr(img):
    return _index_op(lib().ref_fliplr, img, img.shape)


def rotate(img, angle_deg: float):
    return _index_op(lib().ref_rotate, img, img.shape, _D(angle_deg))


def gauss_weights(sigma: float) -> np.ndarray:
    w = np.empty(17, np.float64)
    lib().ref_gauss_weights(_D(sigma), _p(w))
    return w


def gaussian(img: np.ndarray, sigma: float) -> np.ndarray:
    if img.ndim == 3:
        px = _pack(img).copy()
        lib().ref_gaussian_rgba(_p(px), _I(img.shape[1]), _I(img.shape[0]), _D(sigma))
        return _unpack(px)
    out = np.ascontiguousarray(img, np.float64).copy()
    lib().ref_gaussian_f64(_p(out), _I(img.shape[1]), _I(img.shape[0]), _D(sigma))
    return out


def brightness(img: np.ndarray, delta: float) -> np.ndarray:
    if img.ndim == 3:
        src = _pack(img)
        out = np.empty_like(src)
        lib().ref_brightness_rgba(_p(src), _p(out), _L(src.size), _D(delta))
        return _unpack(out)
    src = np.ascontiguousarray(img, np.float64)
    out = np.empty_like(src)
    lib().ref_brightness_f64(_p(src), _p(out), _L(src.size), _D(delta))
    return out


def adjust_gamma(img: np.ndarray, gamma: float, gain: float = 1.0) -> np.ndarray:
    if img.ndim == 3:
        src = _pack(img)
        out = np.empty_like(src)
        lib().ref_gamma_rgba(_p(src), _p(out), _L(src.size), _D(gamma), _D(gain))
        return _unpack(out)
    src = np.ascontiguousarray(img, np.float64)
    out = np.empty_like(src)
    lib().ref_gamma_f64(_p(src), _p(out), _L(src.size), _D(gamma), _D(gain))
    return out


def colorize(img: np.ndarray, r: float, g: float, b: float) -> np.ndarray:
    if img.ndim != 3:
        return img.copy()      # src/millipyde_image.cpp:647-651: no-op on grey
    src = _pack(img)
    out = np.empty_like(src)
    lib().ref_colorize_rgba(_p(src), _p(out), _L(src.size), _D(r), _D(g), _D(b))
    return _unpack(out)


_OPS = {
    "rgb2grey": lambda a: grey_u8(a),
    "transpose": transpose,
    "fliplr": fliplr,
    "rotate": rotate,
    "gaussian": gaussian,
    "brightness": brightness,
    "adjust_gamma": adjust_gamma,
    "colorize": colorize,
}


def apply_chain(img: np.ndarray, chain) -> np.ndarray:
    """Reference-exact chain on the reference layouts (RGBA8 / fp64 grey)."""
    out = img
    for name, *args in chain:
        out = _OPS[name](out, *args)
    return out


if __name__ == "__main__":
    print(build(force=True))
