"""ctypes view of the C ABI exported by libmp_b200.so (include/*.h).

This is the binding a maintainer of the reference would write to call the
library without the CPython extension, and what the parity tests and bench.py
use to drive the hot path "through the C-ABI": plain pointers and sizes, no
torch types.  It contains no arithmetic and no CPU fallback: if the shared
library is missing, or a CUDA device is missing when an op is called, it
raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmp_b200.so")

HOST_LOC = -1
DEVICE_LOC_NO_AFFINITY = -2
SEMANTICS_ORACLE = 0
SEMANTICS_REFERENCE = 1
RANGE_UNIT = 0     # fp32 images hold [0, 1] data (default): the Gaussian may use fp16 correction operands
RANGE_ANY = 1      # arbitrary float data: |sample| >= 65504 is safe, the Gaussian stays on the FMA pipe

NPY_UBYTE, NPY_FLOAT, NPY_DOUBLE = 2, 11, 12


class MPObjData(C.Structure):
    """include/mp_abi.h (reference: src/include/millipyde.h:16-25)."""
    _fields_ = [
        ("device_data", C.c_void_p),
        ("ndims", C.c_int),
        ("dims", C.POINTER(C.c_int)),
        ("type", C.c_int),
        ("mem_loc", C.c_int),
        ("stream", C.c_void_p),
        ("pinned", C.c_int),
        ("nbytes", C.c_size_t),
    ]


class RotateArgs(C.Structure):
    _fields_ = [("angle", C.c_double)]


class GaussianArgs(C.Structure):
    _fields_ = [("sigma", C.c_double)]


class BrightnessArgs(C.Structure):
    _fields_ = [("delta", C.c_double)]


class ColorizeArgs(C.Structure):
    _fields_ = [("r_mult", C.c_double), ("g_mult", C.c_double), ("b_mult", C.c_double)]


class GammaArgs(C.Structure):
    _fields_ = [("gamma", C.c_double), ("gain", C.c_double)]


class ElementwiseArgs(C.Structure):
    """include/mp_image.h: kind is one of MP_EW_ADD (4), MP_EW_MUL (5), MP_EW_POW (6), MP_EW_CLIP (7)."""
    _fields_ = [("kind", C.c_double), ("a", C.c_double), ("b", C.c_double), ("c", C.c_double),
                ("per_channel", C.c_double)]


class RandomRangeArgs(C.Structure):
    _fields_ = [("min", C.c_double), ("max", C.c_double)]


class RandomGammaArgs(C.Structure):
    _fields_ = [("gamma_min", C.c_double), ("gamma_max", C.c_double),
                ("gain_min", C.c_double), ("gain_max", C.c_double)]


class RandomColorizeArgs(C.Structure):
    _fields_ = [("r_min", C.c_double), ("r_max", C.c_double), ("g_min", C.c_double),
                ("g_max", C.c_double), ("b_min", C.c_double), ("b_max", C.c_double)]


MPFunc = C.CFUNCTYPE(C.c_int, C.POINTER(MPObjData), C.c_void_p)


class MPRunnable(C.Structure):
    """include/mp_abi.h (reference: src/include/millipyde.h:98-103)."""
    _fields_ = [
        ("func", C.c_void_p),
        ("obj_data", C.POINTER(MPObjData)),
        ("args", C.c_void_p),
        ("probability", C.c_double),
    ]


_OBJ = C.POINTER(MPObjData)


class MPHostResult(C.Structure):
    """include/mp_pipeline.h"""
    _fields_ = [("ndims", C.c_int), ("shape", C.c_long * 3), ("type", C.c_int),
                ("nbytes", C.c_size_t), ("status", C.c_int)]


# name -> (restype, argtypes); every symbol include/*.h declares
SYMBOLS = {
    # mp_abi.h
    "mperr_str": (C.c_char_p, [C.c_int]),
    "random_int_in_range": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "random_double_in_range": (C.c_int, [C.c_double, C.c_double, C.POINTER(C.c_double)]),
    "mprand_seed": (None, [C.c_uint64]),
    "mprand_keyed_double": (C.c_double, [C.c_uint64, C.c_uint64, C.c_uint, C.c_uint, C.c_double, C.c_double]),
    # mp_image.h
    "mpimg_color_to_greyscale": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_transpose": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_gaussian": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_fliplr": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_rotate": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_brightness": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_colorize": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_adjust_gamma": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_random_rotate": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_random_gaussian": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_random_brightness": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_random_adjust_gamma": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_random_colorize": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_elementwise": (C.c_int, [_OBJ, C.c_void_p]),
    "mpimg_func_from_name": (C.c_void_p, [C.c_char_p, C.POINTER(C.c_size_t)]),
    "mpimg_set_semantics": (None, [C.c_int]),
    "mpimg_get_semantics": (C.c_int, []),
    "mpimg_set_gauss_column": (None, [C.c_int]),
    "mpimg_get_gauss_column": (C.c_int, []),
    "mpimg_gauss_stream_plan": (None, [C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "mpimg_set_value_range": (None, [C.c_int]),
    "mpimg_get_value_range": (C.c_int, []),
    "mpimg_gaussian_effective_radius": (C.c_int, [C.c_double, C.POINTER(C.c_int)]),
    # mp_objects.h
    "mpobj_copy_from_host": (None, [_OBJ, C.c_void_p, C.c_size_t]),
    "mpobj_copy_to_host": (C.c_void_p, [_OBJ]),
    "mpobj_change_device": (None, [_OBJ, C.c_int]),
    "mpobj_dealloc_device_data": (None, [_OBJ]),
    "mpobj_clone_data": (_OBJ, [_OBJ, C.c_int, C.c_int]),
    "mpobj_view_data": (_OBJ, [_OBJ]),
    "mpobj_view_rebind": (None, [_OBJ, _OBJ]),
    "mpobj_view_rebind_many": (None, [C.POINTER(_OBJ), C.POINTER(_OBJ), C.c_int]),
    "mpobj_copy_to_host_into": (C.c_int, [_OBJ, C.c_void_p, C.c_size_t]),
    "mpobj_upload_async": (C.c_int, [_OBJ, C.c_void_p, C.c_size_t]),
    "mpobj_download_async": (C.c_int, [_OBJ, C.c_void_p, C.c_size_t]),
    "mpobj_synchronize": (C.c_int, [_OBJ]),
    "mpobj_create": (_OBJ, [C.c_void_p, C.c_int, C.POINTER(C.c_long), C.c_int]),
    "mpobj_destroy": (None, [_OBJ]),
    "mpobj_set_stream": (None, [_OBJ, C.c_void_p]),
    "mphost_alloc_pinned": (C.c_void_p, [C.c_size_t]),
    "mphost_free_pinned": (None, [C.c_void_p]),
    "mphost_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mphost_unregister": (C.c_int, [C.c_void_p]),
    "mp_last_error": (C.c_char_p, []),
    # mp_devices.h
    "mpdev_initialize": (C.c_int, []),
    "mpdev_teardown": (None, []),
    "mpdev_peer_to_peer_supported": (C.c_int, []),
    "mpdev_can_use_peer": (C.c_int, [C.c_int, C.c_int]),
    "mpdev_get_device_count": (C.c_int, []),
    "mpdev_is_valid_device": (C.c_int, [C.c_int]),
    "mpdev_get_stream": (C.c_void_p, [C.c_int, C.c_int]),
    "mpdev_submit_work": (None, [C.c_int, C.c_void_p, C.c_void_p]),
    "mpdev_hard_synchronize": (None, [C.c_int]),
    "mpdev_hard_synchronize_all": (None, []),
    "mpdev_synchronize": (None, []),
    "mpdev_synchronize_all": (None, []),
    "mpdev_reset": (None, [C.c_int]),
    "mpdev_set_device": (None, [C.c_int]),
    "mpdev_stream_synchronize": (None, [C.c_int, C.c_int]),
    "mpdev_get_target_device": (C.c_int, []),
    "mpdev_get_alternative_device": (C.c_int, [C.c_int]),
    "mpdev_get_next_device": (C.c_int, [C.c_int]),
    "mpdev_set_target_device": (None, [C.c_int]),
    "mpdev_get_recommended_device": (C.c_int, []),
    "mpwrk_create_work_node": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "mpwrk_destroy_work_node": (None, [C.c_void_p]),
    "mpwrk_work_queue_pop": (C.c_void_p, [C.c_void_p]),
    "mpwrk_work_queue_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mpwrk_work_wait": (None, [C.c_void_p]),
    "mpwrk_process_work": (C.c_void_p, [C.c_void_p]),
    "mpwrk_create_work_pool": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "mpwrk_destroy_work_pool": (C.c_int, [C.c_void_p]),
    "mpdev_event_create": (C.c_void_p, [C.c_int]),
    "mpdev_event_destroy": (None, [C.c_void_p]),
    "mpdev_event_record": (None, [C.c_void_p, C.c_void_p]),
    "mpdev_event_elapsed_ms": (C.c_float, [C.c_void_p, C.c_void_p]),
    "mpdev_mem_info": (C.c_int, [C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "mpdev_sm_count": (C.c_int, [C.c_int]),
    "mpdev_pci_bus_id": (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
    "mpdev_flush_l2": (None, [C.c_int, C.c_void_p]),
    "mpdev_trim_pools": (None, []),
    "mpdev_launch_count": (C.c_ulonglong, []),
    # mp_pipeline.h
    "mppipe_create": (C.c_void_p, [C.POINTER(MPRunnable), C.c_int, C.c_int]),
    "mppipe_destroy": (None, [C.c_void_p]),
    "mppipe_get_device": (C.c_int, [C.c_void_p]),
    "mppipe_set_device": (None, [C.c_void_p, C.c_int]),
    "mppipe_connect": (None, [C.c_void_p, C.c_void_p]),
    "mppipe_run": (C.c_int, [C.c_void_p, C.POINTER(_OBJ), C.c_int]),
    "mppipe_run_views": (C.c_int, [C.c_void_p, C.POINTER(_OBJ), C.c_int]),
    "mppipe_submit": (C.c_int, [C.c_void_p, C.POINTER(_OBJ), C.c_int]),
    "mppipe_submit_views": (C.c_int, [C.c_void_p, C.POINTER(_OBJ), C.c_int]),
    "mppipe_wait": (C.c_int, [C.c_void_p]),
    "mppipe_run_host": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_size_t,
                                  C.POINTER(MPHostResult), C.c_int, C.c_int, C.POINTER(C.c_long), C.c_int]),
    "mppipe_plan": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int]),
    "mppipe_set_index_base": (None, [C.c_void_p, C.c_ulonglong]),
    "mppipe_last_run_key": (C.c_ulonglong, [C.c_void_p]),
    "mppipe_hold_run_key": (None, [C.c_void_p]),
    "mppipe_set_device_draws": (None, [C.c_int]),
    "mppipe_get_device_draws": (C.c_int, []),
    "mppipe_set_fusion": (None, [C.c_int]),
    "mppipe_get_fusion": (C.c_int, []),
    "mppipe_last_launches": (C.c_ulonglong, [C.c_void_p]),
    "mppipe_last_segments": (C.c_int, [C.c_void_p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libmp_b200.so and type every exported entry point.  Loading needs no
    GPU (the CUDA runtime is linked statically and binds the driver lazily)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m millipyde_b200.build` "
                "(there is no CPU fallback)")
        handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class MillipydeError(RuntimeError):
    def __init__(self, status: int, where: str):
        msg = lib().mperr_str(status).decode()
        detail = lib().mp_last_error().decode()
        super().__init__(f"{where}: [{status}] {msg}" + (f" -- {detail}" if detail else ""))
        self.status = status


def check(status: int, where: str) -> None:
    if status != 0:
        raise MillipydeError(status, where)


def initialize() -> int:
    """mpdev_initialize(); returns the device count.  Raises without a GPU."""
    check(lib().mpdev_initialize(), "mpdev_initialize")
    return lib().mpdev_get_device_count()


_TYPENUM = {np.dtype(np.uint8): NPY_UBYTE, np.dtype(np.float32): NPY_FLOAT,
            np.dtype(np.float64): NPY_DOUBLE}
_DTYPE = {np.dtype(t).num: np.dtype(t) for t in
          (np.bool_, np.int8, np.uint8, np.int16, np.uint16, np.int32, np.uint32,
           np.int64, np.uint64, np.float16, np.float32, np.float64)}

_OPS = {
    # name -> (symbol, args struct or None)
    "rgb2grey": ("mpimg_color_to_greyscale", None),
    "transpose": ("mpimg_transpose", None),
    "fliplr": ("mpimg_fliplr", None),
    "rotate": ("mpimg_rotate", RotateArgs),
    "gaussian": ("mpimg_gaussian", GaussianArgs),
    "brightness": ("mpimg_brightness", BrightnessArgs),
    "adjust_gamma": ("mpimg_adjust_gamma", GammaArgs),
    "colorize": ("mpimg_colorize", ColorizeArgs),
    "random_rotate": ("mpimg_random_rotate", RandomRangeArgs),
    "random_gaussian": ("mpimg_random_gaussian", RandomRangeArgs),
    "random_brightness": ("mpimg_random_brightness", RandomRangeArgs),
    "random_adjust_gamma": ("mpimg_random_adjust_gamma", RandomGammaArgs),
    "random_colorize": ("mpimg_random_colorize", RandomColorizeArgs),
}


def op_symbol(name: str):
    return _OPS[name]


class DeviceImage:
    """Owns one MPObjData* created through the C ABI (what src/gpuarray.c:82-114
    does inline).  Only a convenience for tests and bench; all work happens in
    the library."""

    def __init__(self, array: np.ndarray | None = None, *, _ptr=None):
        L = lib()
        if _ptr is not None:
            self.ptr = _ptr
            return
        a = np.ascontiguousarray(array)
        if a.dtype not in _TYPENUM:
            # other numeric dtypes are plain gpuarrays: typenum straight from numpy
            typenum = a.dtype.num
        else:
            typenum = _TYPENUM[a.dtype]
        shape = (C.c_long * a.ndim)(*a.shape)
        self.ptr = L.mpobj_create(a.ctypes.data_as(C.c_void_p), a.ndim, shape, typenum)
        if not self.ptr:
            raise MillipydeError(lib().mpdev_initialize() or 57, "mpobj_create")

    # -- header ---------------------------------------------------------------
    @property
    def obj(self) -> MPObjData:
        return self.ptr.contents

    @property
    def shape(self):
        o = self.obj
        return tuple(o.dims[i] for i in range(o.ndims))

    @property
    def dtype(self):
        return _DTYPE[self.obj.type]

    @property
    def device(self) -> int:
        return self.obj.mem_loc

    # -- ops ------------------------------------------------------------------
    def apply(self, name: str, *args) -> "DeviceImage":
        sym, argtype = _OPS[name]
        fn = getattr(lib(), sym)
        if argtype is None:
            check(fn(self.ptr, None), sym)
        else:
            a = argtype(*[float(x) for x in args])
            check(fn(self.ptr, C.cast(C.pointer(a), C.c_void_p)), sym)
        return self

    def apply_chain(self, chain) -> "DeviceImage":
        for name, *args in chain:
            self.apply(name, *args)
        return self

    def clone(self, device: int | None = None, stream: int = 0) -> "DeviceImage":
        dev = self.device if device is None else device
        p = lib().mpobj_clone_data(self.ptr, dev, stream)
        if not p:
            raise MillipydeError(57, "mpobj_clone_data")
        return DeviceImage(_ptr=p)

    def view(self) -> "DeviceImage":
        """Header copy borrowing this image's buffer: input of Chain.run_views only."""
        p = lib().mpobj_view_data(self.ptr)
        if not p:
            raise MillipydeError(57, "mpobj_view_data")
        return DeviceImage(_ptr=p)

    def rebind(self, src: "DeviceImage | None") -> "DeviceImage":
        """Re-arm a view (mpobj_view_rebind): its previous result goes back to the pool, it borrows
        `src` again.  src=None only returns the buffer."""
        lib().mpobj_view_rebind(self.ptr, src.ptr if src is not None else None)
        return self

    def to_device(self, device: int) -> "DeviceImage":
        lib().mpobj_change_device(self.ptr, device)
        return self

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, self.dtype)
        check(lib().mpobj_copy_to_host_into(self.ptr, out.ctypes.data_as(C.c_void_p), out.nbytes),
              "mpobj_copy_to_host_into")
        return out

    def sync(self) -> None:
        out = C.c_char()
        del out
        L = lib()
        L.mpdev_set_device(self.device)
        L.mpdev_synchronize()

    def close(self) -> None:
        if getattr(self, "ptr", None):
            lib().mpobj_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
