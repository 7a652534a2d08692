"""ORACLE (test infrastructure) — ctypes binding of oracle/raster_ref.c plus an
autograd wrapper, so the C rasterizer can stand in for `torch_ref.rasterize_to_pixels`
at sizes where the pure-PyTorch one is too slow (1080p)."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path
from typing import Optional

import numpy as np
import torch

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "libraster_ref.so"
_lib = None


def build(force: bool = False) -> Path:
    if force or not _SO.exists() or _SO.stat().st_mtime < (_DIR / "raster_ref.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_SO))
        P, U32, I64 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int64
        _lib.raster_fwd_ref.restype = None
        _lib.raster_fwd_ref.argtypes = [U32, I64, U32, P, P, P, P, P, P, U32, U32, U32, U32, U32, P, P, P, P, P, P]
        _lib.raster_bwd_ref.restype = None
        _lib.raster_bwd_ref.argtypes = [U32, I64, I64, U32, P, P, P, P, P, P, U32, U32, U32, U32, U32, P, P, P, P,
                                        P, P, P, P, P, P, P]
    return _lib


def _np(t: Optional[torch.Tensor], dtype) -> Optional[np.ndarray]:
    if t is None:
        return None
    return np.ascontiguousarray(t.detach().cpu().numpy().astype(dtype, copy=False))


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def raster_fwd(means2d, conics, colors, opacities, W, H, ts, offsets, flatten_ids, backgrounds=None, masks=None,
               want_margin=False):
    C, th, tw = offsets.shape
    D = colors.shape[-1]
    m, cn, col, op = _np(means2d, np.float32), _np(conics, np.float32), _np(colors, np.float32), _np(opacities, np.float32)
    bg, mk = _np(backgrounds, np.float32), _np(masks, np.uint8)
    off, fid = _np(offsets, np.int32), _np(flatten_ids, np.int32)
    oc = np.zeros((C, H, W, D), np.float32)
    oa = np.zeros((C, H, W, 1), np.float32)
    li = np.zeros((C, H, W), np.int32)
    mg = np.ones((C, H, W), np.float32) if want_margin else None
    lib().raster_fwd_ref(C, fid.size, D, _p(m), _p(cn), _p(col), _p(op), _p(bg), _p(mk), W, H, ts, tw, th, _p(off),
                         _p(fid), _p(oc), _p(oa), _p(li), _p(mg))
    out = (torch.from_numpy(oc), torch.from_numpy(oa), torch.from_numpy(li))
    return out + (torch.from_numpy(mg),) if want_margin else out


def raster_bwd(means2d, conics, colors, opacities, W, H, ts, offsets, flatten_ids, render_alphas, last_ids,
               v_render_colors, v_render_alphas, backgrounds=None, masks=None, absgrad=False):
    """Returns float64 gradients (v_means2d_abs|None, v_means2d, v_conics, v_colors, v_opacities)."""
    C, th, tw = offsets.shape
    D = colors.shape[-1]
    m, cn, col, op = _np(means2d, np.float32), _np(conics, np.float32), _np(colors, np.float32), _np(opacities, np.float32)
    bg, mk = _np(backgrounds, np.float32), _np(masks, np.uint8)
    off, fid = _np(offsets, np.int32), _np(flatten_ids, np.int32)
    ra, li = _np(render_alphas, np.float32), _np(last_ids, np.int32)
    vc, va = _np(v_render_colors, np.float32), _np(v_render_alphas, np.float32)
    n = op.size
    g_abs = np.zeros((n, 2), np.float64) if absgrad else None
    g_m, g_c, g_col, g_o = (np.zeros((n, 2), np.float64), np.zeros((n, 3), np.float64),
                            np.zeros((n, D), np.float64), np.zeros((n,), np.float64))
    lib().raster_bwd_ref(C, n, fid.size, D, _p(m), _p(cn), _p(col), _p(op), _p(bg), _p(mk), W, H, ts, tw, th,
                         _p(off), _p(fid), _p(ra), _p(li), _p(vc), _p(va), _p(g_abs), _p(g_m), _p(g_c), _p(g_col),
                         _p(g_o))
    shp = tuple(opacities.shape)
    f = lambda a, k: torch.from_numpy(a).reshape(shp + ((k,) if k else ()))
    return (f(g_abs, 2) if absgrad else None, f(g_m, 2), f(g_c, 3), f(g_col, D), f(g_o, 0))


class _RasterC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, backgrounds, masks, W, H, ts, offsets, flatten_ids):
        oc, oa, li = raster_fwd(means2d, conics, colors, opacities, W, H, ts, offsets, flatten_ids, backgrounds, masks)
        ctx.save_for_backward(means2d, conics, colors, opacities, offsets, flatten_ids, oa, li)
        ctx.bg, ctx.masks, ctx.dims = backgrounds, masks, (W, H, ts)
        return oc, oa

    @staticmethod
    def backward(ctx, v_c, v_a):
        means2d, conics, colors, opacities, offsets, flatten_ids, oa, li = ctx.saved_tensors
        W, H, ts = ctx.dims
        _, g_m, g_c, g_col, g_o = raster_bwd(means2d, conics, colors, opacities, W, H, ts, offsets, flatten_ids, oa,
                                             li, v_c, v_a, ctx.bg, ctx.masks)
        v_bg = None
        if ctx.bg is not None and ctx.needs_input_grad[4]:
            v_bg = (v_c * (1.0 - oa)).sum(dim=(1, 2))
        return (g_m.to(means2d.dtype), g_c.to(conics.dtype), g_col.to(colors.dtype), g_o.to(opacities.dtype), v_bg,
                None, None, None, None, None, None)


def rasterize_to_pixels(means2d, conics, colors, opacities, image_width, image_height, tile_size, isect_offsets,
                        flatten_ids, backgrounds=None, masks=None, packed=False):
    """Same signature as torch_ref.rasterize_to_pixels; C forward + closed-form C backward."""
    return _RasterC.apply(means2d, conics, colors.contiguous(), opacities, backgrounds, masks, image_width,
                          image_height, tile_size, isect_offsets, flatten_ids)
