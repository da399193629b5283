"""CPU oracle, part A: numpy/scipy restatement of the scikit-image calls the
reference's tests use as ground truth.

TEST INFRASTRUCTURE ONLY.  Nothing under ``millipyde_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or as the timed CPU baseline.

Why a restatement: the reference's oracle is third-party scikit-image 0.18.2 /
numpy 1.21.0 / scipy 1.7.0 (reference ``requirements.txt:1-3``); scikit-image is
not installed in this image and cannot be (no network).  Each function below
cites the reference call site it stands in for and the published skimage 0.18
algorithm it restates.

Pinning status (see DESIGN.md "Oracle"):
  * gaussian   -- ``skimage.filters.gaussian`` 0.18 *is* a call to
    ``scipy.ndimage.gaussian_filter``; scipy is installed, so this one is the
    real third-party code, not a restatement.
  * rgb2grey / adjust_gamma / transpose / fliplr -- one-line numpy formulas,
    cross-checked in tests/test_oracle.py against the reference kernel formulas
    (oracle/ref_exact.c) on the reference's own PNG fixture.
  * rotate (bilinear) -- restated from skimage's ``_warp_fast`` bilinear /
    mode='constant' rule; cross-checked against
    ``scipy.ndimage.affine_transform(order=1, mode='grid-constant')``.
    No reference test compares rotate with anything: parity unpinned by the
    reference for this op.
"""
from __future__ import annotations

import numpy as np
import scipy.ndimage as _ndi

# Luma weights of skimage.color.rgb2gray (same constants as the reference
# kernel, src/millipyde_image.cpp:64).
LUMA = (0.2125, 0.7154, 0.0721)


def _as_float(img: np.ndarray) -> np.ndarray:
    """skimage.img_as_float: uint8 -> float64 / 255; floats pass through."""
    if img.dtype == np.uint8:
        return img.astype(np.float64) / 255.0
    if img.dtype in (np.float32, np.float64):
        return img
    raise TypeError(f"unsupported image dtype {img.dtype}")


def rgba2rgb(img: np.ndarray) -> np.ndarray:
    """skimage.color.rgba2rgb with the default white background:
    ``rgb * a + (1 - a)``.  Identity on the colour channels when alpha == 1,
    which holds for every fixture in the reference.
    Call site: tests/millipyde_tests.py:119."""
    f = _as_float(img).astype(np.float64)
    a = f[..., 3:4]
    return np.clip(f[..., :3] * a + (1.0 - a), 0.0, 1.0)


def rgb2grey(img: np.ndarray) -> np.ndarray:
    """``rgb2gray(rgba2rgb(img))`` for 4-channel input, ``rgb2gray(img)`` for
    3-channel input (tests/millipyde_tests.py:117-125).  Always computed in
    float64; callers comparing an fp32 device result upcast it first."""
    if img.ndim != 3 or img.shape[2] not in (3, 4):
        raise ValueError("rgb2grey expects H x W x {3,4}")
    rgb = rgba2rgb(img) if img.shape[2] == 4 else _as_float(img).astype(np.float64)
    return rgb @ np.array(LUMA, dtype=np.float64)


def transpose(img: np.ndarray) -> np.ndarray:
    """``np.transpose`` for 2-D (tests/millipyde_tests.py:130); the pixel is the
    unit for H x W x C, i.e. axes (1, 0, 2) (reference kernel moves an RGBA
    pixel as one uint32, src/millipyde_image.cpp:73-96)."""
    if img.ndim == 2:
        return np.ascontiguousarray(img.T)
    return np.ascontiguousarray(np.transpose(img, (1, 0, 2)))


def fliplr(img: np.ndarray) -> np.ndarray:
    """``np.fliplr`` (reference kernel src/millipyde_image.cpp:99-111)."""
    return np.ascontiguousarray(img[:, ::-1])


def gaussian_radius(sigma: float, truncate: float = 8.0) -> int:
    """scipy.ndimage.gaussian_filter1d: ``lw = int(truncate * sd + 0.5)``."""
    return int(truncate * float(sigma) + 0.5)


def gaussian_weights(sigma: float, truncate: float = 8.0) -> np.ndarray:
    """scipy's 1-D kernel: ``exp(-0.5/sigma^2 * x^2)`` normalised in float64."""
    r = gaussian_radius(sigma, truncate)
    x = np.arange(-r, r + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def gaussian(img: np.ndarray, sigma: float, truncate: float = 8.0, keep_float32: bool = False) -> np.ndarray:
    """``skimage.filters.gaussian(img, sigma=s, cval=0, truncate=8,
    mode="constant")`` (tests/millipyde_tests.py:550).  skimage 0.18 converts to
    float and calls ``scipy.ndimage.gaussian_filter`` with sigma 0 on the
    channel axis for multichannel input.

    The parity oracle evaluates in float64 (default).  ``keep_float32=True`` is
    what skimage 0.18 itself does with a float32 image (``img_as_float`` keeps the
    precision, scipy filters in the input's dtype): the form the CPU *timing*
    legs of bench.py use, so that the CPU arm is not handicapped by an upcast."""
    f = _as_float(img)
    if not (keep_float32 and f.dtype == np.float32):
        f = f.astype(np.float64)
    sig = (sigma, sigma) if f.ndim == 2 else (sigma, sigma, 0)
    return _ndi.gaussian_filter(f, sigma=sig, mode="constant", cval=0.0,
                                truncate=truncate)


def adjust_gamma(img: np.ndarray, gamma: float, gain: float = 1.0) -> np.ndarray:
    """``skimage.exposure.adjust_gamma`` as of 0.18.2
    (tests/millipyde_tests.py:561, :572):
    ``out = ((image / scale) ** gamma) * scale * gain; out.astype(dtype)`` with
    scale = 255 for uint8 (truncating cast) and 1 for floats.  skimage >= 0.19
    switched uint8 to a rint LUT; the reference pins 0.18.2."""
    if img.dtype == np.uint8:
        out = ((img / 255.0) ** gamma) * 255.0 * gain
        return np.clip(out, 0, 255).astype(np.uint8)
    f = img.astype(np.float64)
    return (f ** gamma) * gain


def adjust_gamma_rgba(img: np.ndarray, gamma: float, gain: float = 1.0) -> np.ndarray:
    """uint8 RGBA as the reference treats it: gamma on R, G, B, alpha kept
    (src/millipyde_image.cpp:457-491).  skimage would also map alpha, but
    255 -> 255 for gain 1, so the two agree on every fixture (alpha == 255)."""
    out = img.copy()
    out[..., :3] = adjust_gamma(img[..., :3], gamma, gain)
    return out


def brightness(img: np.ndarray, delta: float) -> np.ndarray:
    """Float images: ``clip(v + delta, 0, 1)`` (src/millipyde_image.cpp:384-398).
    No skimage call exists for this op; the reference kernel is the definition."""
    return np.clip(img.astype(np.float64) + delta, 0.0, 1.0)


def colorize(img: np.ndarray, r: float, g: float, b: float) -> np.ndarray:
    """Float H x W x {3,4}: per-channel multiply saturating at 1, alpha kept --
    the float analogue of src/millipyde_image.cpp:494-524 (uint8 saturates at
    255).  Single-channel images are returned unchanged (:647-651)."""
    f = img.astype(np.float64)
    if f.ndim == 2:
        return f
    out = f.copy()
    out[..., 0] = np.minimum(1.0, f[..., 0] * r)
    out[..., 1] = np.minimum(1.0, f[..., 1] * g)
    out[..., 2] = np.minimum(1.0, f[..., 2] * b)
    return out


def rotate_coords(h: int, w: int, angle_deg: float):
    """Source coordinates (xs, ys), float64, for every output pixel under
    ``skimage.transform.rotate(img, angle)`` defaults: rotation about
    ``(W/2 - 0.5, H/2 - 0.5)``, inverse map
    ``xs = cos(t)(x-cx) - sin(t)(y-cy) + cx``, ``ys = sin(t)(x-cx) + cos(t)(y-cy) + cy``
    -- the same sense as the reference kernel (src/millipyde_image.cpp:120-123)."""
    t = np.deg2rad(angle_deg)
    c, s = np.cos(t), np.sin(t)
    cx, cy = w / 2.0 - 0.5, h / 2.0 - 0.5
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    xs = c * (x - cx) - s * (y - cy) + cx
    ys = s * (x - cx) + c * (y - cy) + cy
    return xs, ys


def rotate(img: np.ndarray, angle_deg: float) -> np.ndarray:
    """Bilinear, mode='constant', cval=0 rotation, same output size
    (``skimage.transform.rotate`` defaults: order=1, resize=False).  Rule of
    skimage's ``_warp_fast.bilinear_interpolation``: x0=floor, x1=ceil, each of
    the four corners is the pixel if inside the image else cval, blended with
    ``dx = xs - x0``, ``dy = ys - y0``.  Only un-asserted timing code in the
    reference calls it (examples/image_examples.py:62)."""
    f = _as_float(img).astype(np.float64)
    h, w = f.shape[:2]
    xs, ys = rotate_coords(h, w, angle_deg)
    x0 = np.floor(xs)
    y0 = np.floor(ys)
    x1 = np.ceil(xs)
    y1 = np.ceil(ys)
    dx = xs - x0
    dy = ys - y0
    if f.ndim == 3:
        dx = dx[..., None]
        dy = dy[..., None]

    def px(yy, xx):
        inside = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        yi = np.clip(yy, 0, h - 1).astype(np.int64)
        xi = np.clip(xx, 0, w - 1).astype(np.int64)
        v = f[yi, xi]
        m = inside if f.ndim == 2 else inside[..., None]
        return np.where(m, v, 0.0)

    top = (1 - dx) * px(y0, x0) + dx * px(y0, x1)
    bot = (1 - dx) * px(y1, x0) + dx * px(y1, x1)
    return (1 - dy) * top + dy * bot


def rotate_scipy_crosscheck(img: np.ndarray, angle_deg: float) -> np.ndarray:
    """Independent statement of the same map through scipy (used by the tests to
    pin ``rotate``): affine_transform in (row, col) order with
    mode='grid-constant'."""
    f = _as_float(img).astype(np.float64)
    h, w = f.shape[:2]
    t = np.deg2rad(angle_deg)
    c, s = np.cos(t), np.sin(t)
    m = np.array([[c, s], [-s, c]])
    ctr = np.array([h / 2.0 - 0.5, w / 2.0 - 0.5])
    off = ctr - m @ ctr
    if f.ndim == 2:
        return _ndi.affine_transform(f, m, offset=off, order=1,
                                     mode="grid-constant", cval=0.0)
    return np.stack([_ndi.affine_transform(f[..., k], m, offset=off, order=1,
                                           mode="grid-constant", cval=0.0)
                     for k in range(f.shape[2])], axis=-1)


# ---------------------------------------------------------------------------
# Chains (for fused-pipeline parity): apply a list of (name, args) in order.
# ---------------------------------------------------------------------------
_OPS = {
    "rgb2grey": lambda a: rgb2grey(a),
    "transpose": lambda a: transpose(a),
    "fliplr": lambda a: fliplr(a),
    "rotate": lambda a, ang: rotate(a, ang),
    "gaussian": lambda a, s: gaussian(a, s),
    "brightness": lambda a, d: brightness(a, d),
    "adjust_gamma": lambda a, g, k=1.0: np.clip(adjust_gamma(a, g, k), 0.0, 1.0),
    "colorize": lambda a, r, g, b: colorize(a, r, g, b),
}


def apply_chain(img: np.ndarray, chain) -> np.ndarray:
    """``chain`` = iterable of ``(name, *args)``; float images only.  The
    clamp in adjust_gamma is the reference kernel's (src/millipyde_image.cpp:
    449-450); it is the identity for gain <= 1 on [0, 1] input."""
    out = _as_float(img).astype(np.float64)
    for name, *args in chain:
        out = _OPS[name](out, *args)
    return out
