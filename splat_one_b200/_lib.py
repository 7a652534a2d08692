"""ctypes binding of libb200splat.so (the C ABI declared in include/b200splat.h).

This is the replacement of the reference's loader `gsplat/cuda/_backend.py:81-137`
(import prebuilt `gsplat.csrc`, else JIT-compile).  There is no fallback of any kind:
if the shared library is missing or does not export a symbol the header declares,
`get_lib()` raises.  Every call releases the GIL (plain `ctypes.CDLL`), so the library
is re-entrant by construction (include/b200splat.h header comment).
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import c_char_p, c_float, c_int, c_size_t, c_uint32, c_uint64, c_void_p
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("B200SPLAT_LIB", _PKG / "libb200splat.so"))

ABI_VERSION = 15

_P = c_void_p
_U32 = c_uint32
_U64 = c_uint64
_F = c_float
_I = c_int

# name -> (restype, argtypes); mirrors include/b200splat.h one to one
_PROJ_COMMON = [_U32, _U32, _P, _P, _P, _P, _P, _P, _U32, _U32]
SIGNATURES = {
    "b200splat_abi_version": (_I, []),
    "b200splat_last_error": (c_char_p, []),
    "b200splat_arch": (c_char_p, []),
    "b200splat_copy_small": (_I, [_P, _P, _U32, _P]),
    "b200splat_projection_fwd": (_I, _PROJ_COMMON + [_F, _F, _F, _F, _I, _P, _P, _P, _P, _P, _P]),
    "b200splat_projection_bwd": (
        _I, _PROJ_COMMON + [_F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200splat_projection_bwd_state": (
        _I, _PROJ_COMMON + [_F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _F, _F, _P, _P, _P, _P]),
    "b200splat_strategy_update_state": (_I, [_U32, _U32, _U32, _P, _P, _P, _F, _F, _F, _P, _P, _P, _P]),
    "b200splat_projection_packed_count": (_I, _PROJ_COMMON + [_F, _F, _F, _F, _I, _P, _P, _P]),
    "b200splat_projection_packed_fill": (
        _I, _PROJ_COMMON + [_F, _F, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200splat_projection_packed_bwd": (
        _I, [_U32, _U32, _U32, _P, _P, _P, _P, _P, _P, _U32, _U32, _F, _I,
             _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P]),
    "b200splat_sh_fwd": (_I, [_U32, _U32, _U32, _U32, _P, _P, _P, _P, _P]),
    "b200splat_sh_bwd": (_I, [_U32, _U32, _U32, _U32, _P, _P, _P, _P, _P, _P, _P]),
    "b200splat_camera_centers": (_I, [_U32, _P, _P, _P]),
    "b200splat_sh_colors_fwd": (_I, [_U32, _U32, _U32, _U32, _I, _P, _P, _P, _P, _P, _P]),
    "b200splat_sh_colors_bwd": (_I, [_U32, _U32, _U32, _U32, _I, _P, _P, _P, _P, _P, _P, _P, _P, _U32, _U32, _P]),
    "b200splat_sh_colors_packed_fwd": (_I, [_U32, _U32, _U32, _U32, _U32, _I, _P, _P, _P, _P, _P, _P, _P]),
    "b200splat_sh_colors_packed_bwd": (_I, [_U32, _U32, _U32, _U32, _U32, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200splat_sh_colors_packed_split_fwd": (_I, [_U32, _U32, _U32, _U32, _U32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200splat_sh_colors_packed_split_bwd": (_I, [_U32, _U32, _U32, _U32, _U32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                                  _P, _P]),
    "b200splat_isect_count": (_I, [_I, _U32, _U32, _U32, _P, _P, _P, _U32, _U32, _U32, _P, _P, _P, _P, c_size_t, _P]),
    "b200splat_isect_sorted_workspace_bytes": (c_size_t, [_U64, _U64]),
    "b200splat_isect_sorted": (_I, [_I, _U32, _U32, _U32, _P, _P, _P, _P, _P, _U64, _U32, _U32, _U32, _P, _P, _P, _P,
                                    c_size_t, _P]),
    "b200splat_isect_depth_order_workspace_bytes": (c_size_t, [_U64]),
    "b200splat_isect_depth_order": (_I, [_U64, _P, _P, _P, c_size_t, ctypes.POINTER(c_int), _P]),
    "b200splat_isect_tile_order_workspace_bytes": (c_size_t, [_U64]),
    "b200splat_isect_tile_order": (_I, [_I, _U32, _U32, _U32, _P, _P, _P, _P, _P, _I, _U64, _U32, _U32, _U32, _P, _P,
                                        _P, _P, c_size_t, _P]),
    "b200splat_scan_workspace_bytes": (c_size_t, [_U64]),
    "b200splat_isect_fill": (_I, [_I, _U32, _U32, _U32, _P, _P, _P, _P, _P, _U32, _U32, _U32, _P, _P, _P]),
    "b200splat_sort_workspace_bytes": (c_size_t, [_U64]),
    "b200splat_isect_sort": (_I, [_U64, _U32, _P, _P, _P, _P, _P, c_size_t, ctypes.POINTER(c_int), _P]),
    "b200splat_isect_offset_encode": (_I, [_U64, _P, _U32, _U32, _U32, _P, _P]),
    "b200splat_rasterize_records_bytes": (c_size_t, [_U32, _U32, _U32]),
    "b200splat_rasterize_pack": (_I, [_U32, _U32, _P, _P, _P, _P, _P, _P]),
    "b200splat_rasterize_fwd": (
        _I, [_U32, _U32, _U64, _U32, _P, _P, _P, _P, _P, _P, _U32, _U32, _U32, _U32, _U32, _P, _P, _P, _P,
             _P, _P, _P, _P]),
    "b200splat_rasterize_bwd": (
        _I, [_U32, _U32, _U64, _U32, _P, _P, _P, _P, _P, _P, _U32, _U32, _U32, _U32, _U32, _P, _P, _P, _P,
             _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200splat_raster_indices_count": (
        _I, [_U32, _U32, _U32, _U32, _U64, _P, _P, _P, _U32, _U32, _U32, _U32, _U32, _P, _P, _P, _P, _P, _P, _P,
             c_size_t, _P]),
    "b200splat_raster_indices_fill": (
        _I, [_U32, _U32, _U32, _U32, _U64, _P, _P, _P, _U32, _U32, _U32, _U32, _U32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200splat_quat_scale_to_covar_preci_fwd": (_I, [_U32, _P, _P, _I, _P, _P, _P]),
    "b200splat_quat_scale_to_covar_preci_bwd": (_I, [_U32, _P, _P, _P, _P, _I, _P, _P, _P]),
    "b200splat_world_to_cam_fwd": (_I, [_U32, _U32, _P, _P, _P, _P, _P, _P]),
    "b200splat_world_to_cam_bwd": (_I, [_U32, _U32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "b200splat_proj_fwd": (_I, [_U32, _U32, _P, _P, _P, _U32, _U32, _I, _P, _P, _P]),
    "b200splat_proj_bwd": (_I, [_U32, _U32, _P, _P, _P, _U32, _U32, _I, _P, _P, _P, _P, _P]),
    "b200splat_selective_adam_update": (_I, [_P, _P, _P, _P, _P, _F, _F, _F, _F, _U32, _U32, _P]),
    "b200splat_compute_relocation": (_I, [_U32, _P, _P, _P, _P, _I, _P, _P, _P]),
    "b200splat_sh_colors_staged_smem_bytes": (c_size_t, [_U32, _I]),
    "b200splat_sh_colors_staged_fwd": (_I, [_U32, _U32, _U32, _U32, _P, _P, _P, _P, _P, _P, _P]),
    "b200splat_sh_colors_bwd_peer": (_I, [_U32, _U32, _U32, _U32, _P, _P, _P, _U64, _U32, _U32, _P, _P, _U32, _U32, _P]),
    "b200splat_sh_colors_staged_bwd_peer": (_I, [_U32, _U32, _U32, _U32, _P, _P, _P, _P, _U64, _U32, _U32, _P, _P, _P,
                                                 _U32, _U32, _P]),
    "b200splat_peer_flag_bytes": (c_size_t, [_U32]),
    "b200splat_peer_publish_cotangents": (_I, [_U32, _U32, _U32, _U32, _P, _P, _P, _P, _P]),
    "b200splat_peer_barrier": (_I, [_U32, _U32, _P, _U64, _P]),
    "b200splat_peer_allreduce_f32": (_I, [_U32, _U32, _P, _U64, _U64, _U64, _P, _U64, _P]),
    "b200splat_sh_colors_staged_bwd": (_I, [_U32, _U32, _U32, _U32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _U32, _U32,
                                            _P]),
    "b200splat_invert_4x4": (_I, [_U32, _P, _P, _P]),
    "b200splat_splat_activations_fwd": (_I, [_U32, _P, _P, _P, _P, _P]),
    "b200splat_splat_activations_bwd": (_I, [_U32, _P, _P, _P, _P, _P, _P, _P]),
    "b200splat_l1_ssim_workspace_bytes": (c_size_t, [_U32, _U32, _U32]),
    "b200splat_l1_ssim_fwd": (_I, [_U32, _U32, _U32, _P, _P, _F, _P, _P, _P, _P, _P, c_size_t, _P]),
    "b200splat_l1_ssim_bwd": (_I, [_U32, _U32, _U32, _P, _P, _F, _P, _P, _P, _P, _P, _P]),
}

_lib = None
_lock = threading.Lock()


class B200SplatError(RuntimeError):
    """Raised when the native library is missing or a native call fails."""


def get_lib() -> ctypes.CDLL:
    """Load libb200splat.so once; raise loudly if it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise B200SplatError(
                f"{LIB_PATH} not found. Build it with `python -m splat_one_b200.build` "
                "(needs nvcc; cross-compiles for sm_100a without a GPU). There is no CPU or "
                "PyTorch fallback for the rasterization path.")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:  # pragma: no cover - build mismatch
                raise B200SplatError(f"{LIB_PATH} does not export `{name}`; rebuild it") from e
            fn.restype = res
            fn.argtypes = args
        v = lib.b200splat_abi_version()
        if v != ABI_VERSION:
            raise B200SplatError(f"ABI mismatch: library {v}, binding {ABI_VERSION}; rebuild")
        _lib = lib
    return _lib


def check(rc: int, lib: ctypes.CDLL) -> None:
    if rc != 0:
        msg = lib.b200splat_last_error()
        raise B200SplatError(msg.decode() if msg else f"native call failed with status {rc}")
