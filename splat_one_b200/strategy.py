"""Running densification statistics of the reference's `DefaultStrategy`.

`DefaultStrategy._update_state` (/root/reference/submodules/gsplat/gsplat/strategy/default.py:203-262) runs
after every backward pass and turns `info["means2d"].grad` and `info["radii"]` into three per-Gaussian
running arrays (`grad2d`, `count`, `radii`) with about ten ATen launches over [C,N] temporaries (clone, two
scalings, mask, where, gather, norm, two index_add_, maximum + scatter).  Here it is ONE kernel, in two forms:

* `update_state(state, info, ...)`        — stand-alone, same inputs and effect as the reference method
  (packed or unpacked layout, `absgrad` or plain gradients);
* `strategy_state_sink(state, n_gaussians)` — context manager around `loss.backward()`: the unpacked
  projection backward kernel, which already holds every (camera, Gaussian) 2-D mean cotangent and radius in
  registers, updates the three arrays itself (`b200splat_projection_bwd_state`), so the statistics cost no
  launch and no extra pass over memory.  Plain (non-absgrad) gradients only: `absgrad` statistics come from
  the raster backward's `means2d.absgrad`, use `update_state` for them.

`state` is the dict the reference strategy keeps (`initialize_state()`: keys `grad2d`, `count`, `radii`,
None until first use); the arrays are created here exactly as :230-237 does.  With several cameras seeing one
Gaussian the reference's `state["radii"][ids] = maximum(...)` keeps an arbitrary one of the duplicates ("should
be ideally using scatter max", :255); both forms here take the true maximum.
"""
from typing import Any, Dict, Optional

import torch
from torch import Tensor

from . import wrapper
from .wrapper import _ptr, get_lib, native


def _ensure_state(state: Dict[str, Any], n_gaussians: int, device, with_radii: bool) -> None:
    for key in ("grad2d", "count") + (("radii",) if with_radii else ()):
        if state.get(key) is None:
            state[key] = torch.zeros(n_gaussians, device=device, dtype=torch.float32)
        t = state[key]
        if t.shape != (n_gaussians,) or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f"b200splat: state['{key}'] must be a contiguous float32 tensor of shape [{n_gaussians}]")


@torch.no_grad()
def update_state(state: Dict[str, Any], info: Dict[str, Any], packed: bool = False, absgrad: bool = False,
                 refine_scale2d: bool = True, key_for_gradient: str = "means2d",
                 n_gaussians: Optional[int] = None) -> None:
    """`DefaultStrategy._update_state(params, state, info, packed)` (default.py:203-262) as one kernel.
    `info`: the meta dict of `rasterization()` after `backward()` (keys width, height, n_cameras, radii,
    gaussian_ids, and `key_for_gradient` whose `.grad` / `.absgrad` is read).  `refine_scale2d` =
    `refine_scale2d_stop_iter > 0`.  `n_gaussians`: needed only in packed mode on first use."""
    for key in ("width", "height", "n_cameras", "radii", "gaussian_ids", key_for_gradient):
        assert key in info, f"{key} is required but missing."
    src = info[key_for_gradient]
    grads = src.absgrad if absgrad else src.grad
    assert grads is not None, "backward() has not produced the 2-D mean gradient (retain_grad missing?)"
    grads = grads.contiguous()
    radii = info["radii"].contiguous()
    W, H, C = int(info["width"]), int(info["height"]), int(info["n_cameras"])
    if packed:
        ids = info["gaussian_ids"].contiguous()
        if n_gaussians is None:
            assert state.get("grad2d") is not None, "n_gaussians is required on the first packed update"
            n_gaussians = state["grad2d"].shape[0]
        nnz = ids.numel()
        assert grads.shape == (nnz, 2) and radii.shape == (nnz,), (grads.shape, radii.shape)
    else:
        ids = None
        assert grads.dim() == 3 and grads.shape[-1] == 2 and radii.shape == grads.shape[:2], (grads.shape, radii.shape)
        n_gaussians, nnz = grads.shape[1], 0
    _ensure_state(state, n_gaussians, grads.device, refine_scale2d)
    if radii.dtype != torch.int32:
        radii = radii.to(torch.int32)
    native("strategy_update_state", get_lib(), grads.device, grads.shape[0] if not packed else C, n_gaussians, nnz,
           _ptr(ids), _ptr(grads), _ptr(radii), W / 2.0 * C, H / 2.0 * C, float(max(W, H)), _ptr(state["grad2d"]),
           _ptr(state["count"]), _ptr(state["radii"]) if refine_scale2d else None)


class strategy_state_sink:
    """Context manager around `backward()`: the unpacked projection backward updates `state` in place
    (see the module docstring).  `n_cameras`: the normalisation the reference applies (`info["n_cameras"]`,
    default: the camera batch of the projection call).  `updates` counts the kernels that took the fused
    path; 0 after `backward()` means the call was packed or a different size and `update_state` must be
    used instead."""

    def __init__(self, state: Dict[str, Any], n_gaussians: int, width: int, height: int, device,
                 n_cameras: Optional[int] = None, refine_scale2d: bool = True):
        _ensure_state(state, n_gaussians, device, refine_scale2d)
        self.grad2d, self.count = state["grad2d"], state["count"]
        self.radii = state["radii"] if refine_scale2d else None
        self.N, self.W, self.H, self.n_cameras = n_gaussians, int(width), int(height), n_cameras
        self.updates = 0

    def accepts(self, C: int, N: int, width: int, height: int) -> bool:
        return N == self.N and width == self.W and height == self.H

    def scales(self, C: int):
        c = float(self.n_cameras if self.n_cameras is not None else C)
        return self.W / 2.0 * c, self.H / 2.0 * c, float(max(self.W, self.H))

    def __enter__(self):
        wrapper._STRATEGY_SINK["state"] = self
        return self

    def __exit__(self, *exc):
        wrapper._STRATEGY_SINK.pop("state", None)
        return False
