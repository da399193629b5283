"""ctypes front-end of the chain executor (include/mp_pipeline.h): what the
`Pipeline` / `Generator` types of the CPython extension call, usable without
the extension.  No arithmetic here -- it marshals MPRunnable arrays and
pointer lists and calls mppipe_*."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class Chain:
    """An Operation chain resolved to MPRunnable[] exactly as PyGPUPipeline_init
    does (src/gpupipeline.c:152-161): op names -> mpimg_* symbols, arguments ->
    heap *Args structs, probability -> the stage's coin."""

    def __init__(self, ops, device: int = capi.DEVICE_LOC_NO_AFFINITY):
        L = capi.lib()
        self._ops = list(ops)
        self._keep = []
        arr = (capi.MPRunnable * max(1, len(ops)))()
        for i, op in enumerate(ops):
            name, *rest = op
            prob = -1.0
            if rest and isinstance(rest[-1], dict):
                prob = float(rest[-1].get("probability", -1.0))
                rest = rest[:-1]
            sym, argtype = capi.op_symbol(name)
            arr[i].func = C.cast(getattr(L, sym), C.c_void_p)
            if argtype is not None:
                a = argtype(*[float(x) for x in rest])
                self._keep.append(a)
                arr[i].args = C.cast(C.pointer(a), C.c_void_p)
            arr[i].probability = prob
        self.n = len(ops)
        self.ptr = L.mppipe_create(arr, len(ops), device)
        if not self.ptr:
            raise RuntimeError("mppipe_create failed")

    def connect_to(self, other: "Chain") -> None:
        capi.lib().mppipe_connect(self.ptr, other.ptr)
        self._receiver = other

    @property
    def device(self) -> int:
        return capi.lib().mppipe_get_device(self.ptr)

    def _objs(self, images):
        arr = (C.POINTER(capi.MPObjData) * len(images))(*[im.ptr for im in images])
        return arr

    def run(self, images) -> None:
        capi.check(capi.lib().mppipe_run(self.ptr, self._objs(images), len(images)), "mppipe_run")

    def run_views(self, views) -> None:
        """`views` = [img.view() ...]: the sources stay untouched, every view ends up owning its result."""
        capi.check(capi.lib().mppipe_run_views(self.ptr, self._objs(views), len(views)), "mppipe_run_views")

    def submit(self, images) -> None:
        self._pending = self._objs(images)
        capi.check(capi.lib().mppipe_submit(self.ptr, self._pending, len(images)), "mppipe_submit")

    def submit_views(self, views) -> None:
        self._pending = self._objs(views)
        capi.check(capi.lib().mppipe_submit_views(self.ptr, self._pending, len(views)), "mppipe_submit_views")

    def wait(self) -> None:
        capi.check(capi.lib().mppipe_wait(self.ptr), "mppipe_wait")

    def run_host(self, host_in, host_out):
        """host_in/host_out: lists of C-contiguous ndarrays (same input layout;
        outputs sized for the largest possible result).  Returns per-image
        (shape, dtype) of what landed in host_out[i]."""
        n = len(host_in)
        a0 = host_in[0]
        ins = (C.c_void_p * n)(*[a.ctypes.data for a in host_in])
        outs = (C.c_void_p * n)(*[a.ctypes.data for a in host_out])
        res = (capi.MPHostResult * n)()
        shape = (C.c_long * a0.ndim)(*a0.shape)
        cap = min(a.nbytes for a in host_out)
        capi.check(capi.lib().mppipe_run_host(self.ptr, ins, outs, cap, res, n, a0.ndim, shape, a0.dtype.num),
                   "mppipe_run_host")
        out = []
        for r in res:
            capi.check(r.status, "mppipe_run_host(image)")
            out.append((tuple(r.shape[k] for k in range(r.ndims)), capi._DTYPE[r.type]))
        return out

    @property
    def last_launches(self) -> int:
        return capi.lib().mppipe_last_launches(self.ptr)

    @property
    def last_segments(self) -> int:
        return capi.lib().mppipe_last_segments(self.ptr)

    def close(self):
        if getattr(self, "ptr", None):
            capi.lib().mppipe_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_empty(shape, dtype) -> np.ndarray:
    """ndarray over page-locked host memory (mphost_alloc_pinned)."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = capi.lib().mphost_alloc_pinned(n)
    if not p:
        raise MemoryError("mphost_alloc_pinned failed")
    buf = (C.c_char * n).from_address(p)
    arr = np.frombuffer(buf, dtype=dt).reshape(shape)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED = {}


def pinned_free(arr: np.ndarray) -> None:
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        capi.lib().mphost_free_pinned(p)
