"""Seeded synthetic inputs shared by the tests and bench.py."""
import numpy as np


def noise_f32(h, w, c, seed):
    rng = np.random.default_rng(seed)
    shape = (h, w) if c == 1 else (h, w, c)
    return rng.random(shape, dtype=np.float32)


def smooth_f32(h, w, c):
    """0.5 + 0.5 sin(x/37) cos(y/53) with per-channel offsets: catches orientation
    mistakes that noise hides."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    planes = [0.5 + 0.5 * np.sin((x + 11 * k) / 37.0) * np.cos((y + 7 * k) / 53.0) for k in range(c)]
    img = np.stack(planes, axis=-1).astype(np.float32)
    return img[..., 0] if c == 1 else img


def rgba8(h, w, seed):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    img[..., 3] = 255
    return img
