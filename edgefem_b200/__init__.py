"""edgefem_b200: B200-native (sm_100a) implementation of EdgeFEM's frequency-domain solve hot path.

``edgefem_b200.pyedgefem`` is the pybind11 module with the reference's ``pyedgefem`` names
(C++ host layer -> C-ABI -> CUDA kernels); ``edgefem_b200.cabi`` is the raw ctypes view of the
C-ABI.  Importing the package never touches the GPU; the native modules fail loudly (no CPU
fallback) if they were not built or if no CUDA device is present when a compute call is made.
"""
__all__ = ["cabi", "load_pyedgefem"]


def load_pyedgefem():
    """Import the compiled pybind11 module (raises ImportError with a build hint if it is missing)."""
    try:
        from . import pyedgefem  # type: ignore
    except ImportError as e:  # pragma: no cover
        raise ImportError("edgefem_b200.pyedgefem is not built: run `python -m edgefem_b200.build` (%s)" % e)
    return pyedgefem
