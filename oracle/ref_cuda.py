"""ORACLE (test infrastructure): loader for the reference's own CUDA extension built by
oracle/build_ref.py into oracle/_ref/gsplat_ref_csrc.so (pybind11 entry points of
/root/reference/submodules/gsplat/gsplat/cuda/csrc/ext.cpp:11-56).  Only tests/ and tools/
import this; `load()` returns None when the library was not built (it needs /root/reference
at build time, never at run time)."""
import importlib.machinery
import importlib.util
import os

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "gsplat_ref_csrc.so")
_mod = None


def available() -> bool:
    return os.path.exists(_SO)


def load():
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)

    loader = importlib.machinery.ExtensionFileLoader("gsplat_ref_csrc", _SO)
    spec = importlib.util.spec_from_loader("gsplat_ref_csrc", loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _mod = mod
    return mod


def camera_model(mod, name: str):
    return getattr(mod.CameraModelType, name.upper())
