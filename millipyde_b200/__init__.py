"""millipyde_b200 -- B200-native implementation of Millipyde's image-augmentation
hot path.

    import millipyde_b200 as mp          # or: import millipyde_b200; import millipyde as mp

exposes the reference's Python surface (gpuarray, gpuimage, Operation, Pipeline,
Generator, Device, device_count, get_current_device, best_device,
image_from_path, images_from_path, DEVICE_COUNT) from the in-tree CPython
extension `millipyde` (csrc/py/*.c), which is a thin layer over the C ABI of
libmp_b200.so (include/*.h; csrc/*.cu).  There is no CPU fallback: importing
without the built extension, or without a CUDA device, raises ImportError.

Submodules that never touch a GPU at import time:
    millipyde_b200.build    nvcc/gcc recipes (python -m millipyde_b200.build)
    millipyde_b200.capi     ctypes view of the C ABI
    millipyde_b200.engine   ctypes view of the chain executor
"""
import importlib
import sys

_API = ("gpuarray", "gpuimage", "Operation", "Pipeline", "Generator", "Device", "device_count",
        "get_current_device", "best_device", "image_from_path", "images_from_path", "DEVICE_COUNT",
        "synchronize", "seed", "set_semantics", "get_semantics", "set_fusion", "launch_count",
        "pinned_empty")

_ext = None


def load_extension():
    """Import the CPython extension (initialises every CUDA device) and register
    it as top-level module `millipyde` so reference code runs unchanged."""
    global _ext
    if _ext is None:
        try:
            _ext = importlib.import_module("millipyde_b200.millipyde")
        except ModuleNotFoundError as e:
            raise ImportError(
                "the millipyde extension is not built: run `python -m millipyde_b200.build` "
                "(there is no CPU fallback)") from e
        sys.modules.setdefault("millipyde", _ext)
    return _ext


def __getattr__(name):
    if name in _API:
        return getattr(load_extension(), name)
    raise AttributeError(f"module 'millipyde_b200' has no attribute {name!r}")


def __dir__():
    return sorted(list(globals()) + list(_API))
