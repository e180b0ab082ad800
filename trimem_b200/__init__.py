"""trimem_b200 — B200-native implementation of trimem's force-evaluation hot path.

``trimem_b200.core`` is a drop-in for the reference's ``trimem.core`` pybind11
module (TriMesh, EnergyManager, EnergyParams/BondParams/SurfaceRepulsionParams,
make_nlist, flip/pflip, ...). All evaluation runs in hand-written CUDA kernels
(sm_100a) behind the C ABI declared in ``include/trimem_b200.h``; there is no
CPU fallback.
"""
import importlib
import os
import sys

__all__ = ["core", "install_as_trimem", "lib_path"]


def lib_path():
    """Path of the C-ABI shared library (TRIMEM_B200_LIB overrides it for the ctypes binding only:
    a developer knob to time kernel variants built by scripts/build_variant.py)."""
    if os.environ.get("TRIMEM_B200_LIB"):
        return os.environ["TRIMEM_B200_LIB"]
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libtrimem_b200.so")


def __getattr__(name):
    if name == "core":
        try:
            return importlib.import_module(".core", __name__)
        except ImportError as exc:  # fail loudly: there is nothing to fall back to
            raise ImportError(
                "trimem_b200.core (the CUDA extension) is not built; run "
                "`python -m trimem_b200.build` (needs nvcc, sm_100a)") from exc
    raise AttributeError(name)


def install_as_trimem():
    """Register trimem_b200.core as ``trimem.core`` so the reference's unchanged Python
    drivers (trimem.mc.*) import it."""
    import types
    core = importlib.import_module(".core", __name__)
    pkg = sys.modules.get("trimem")
    if pkg is None:
        pkg = types.ModuleType("trimem")
        pkg.__path__ = []
        sys.modules["trimem"] = pkg
    pkg.core = core
    sys.modules["trimem.core"] = core
    return core
