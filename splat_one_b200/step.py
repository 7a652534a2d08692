"""The training step either side of `rasterization()`.

Mirrors what splat_one's `Runner` does around its one rasterizer call
(R = /root/reference, R/utils/gsplat_utils/gsplat_trainer.py):

  * `rasterize_splats` (:446-497): `scales = exp(.)`, `opacities = sigmoid(.)`,
    `colors = cat([sh0, shN], 1)`, `viewmats = inv(camtoworlds)`, `rasterization(...)`,
    `render_colors[~masks] = 0`;
  * the photometric loss of `train` (:624-628):
    `l1_loss(colors, pixels) * (1 - ssim_lambda) + (1 - fused_ssim(colors, pixels, "valid")) * ssim_lambda`.

Here the two activations are one kernel, the SH table is never concatenated (the colour
kernels read sh0 / shN where they lie and write their gradients in place), and the loss is
one fused kernel per direction over the [C,H,W,3] images `rasterization()` returns.
Everything runs through libb200splat.so; there is no PyTorch fallback.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Tuple

import torch
from torch import Tensor

from ._lib import get_lib
from .rendering import rasterization
from .wrapper import _check_cuda, _f32, _grad_out, _ptr, native


# ----------------------------------------------------------------------------------------
# activations (gsplat_trainer.py:458-459)
# ----------------------------------------------------------------------------------------
class _SplatActivations(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scales_raw: Tensor, opacities_raw: Tensor):
        _check_cuda(scales_raw, opacities_raw)
        _f32(scales_raw), _f32(opacities_raw)
        N = opacities_raw.shape[0]
        scales = torch.empty_like(scales_raw)
        opacities = torch.empty_like(opacities_raw)
        if N:
            native("splat_activations_fwd", get_lib(), scales_raw.device, N, _ptr(scales_raw), _ptr(opacities_raw),
                   _ptr(scales), _ptr(opacities))
        ctx.save_for_backward(scales, opacities)
        return scales, opacities

    @staticmethod
    def backward(ctx, v_scales, v_opacities):
        scales, opacities = ctx.saved_tensors
        N = opacities.shape[0]
        need_s, need_o = ctx.needs_input_grad
        # packed + sparse_grad projection hands back sparse COO cotangents (_wrapper.py:1163-1203)
        if v_scales is not None and v_scales.is_sparse:
            v_scales = v_scales.to_dense()
        if v_opacities is not None and v_opacities.is_sparse:
            v_opacities = v_opacities.to_dense()
        v_scales_raw = _grad_out(scales) if need_s else None
        v_opacities_raw = _grad_out(opacities) if need_o else None
        if N and (need_s or need_o):
            native("splat_activations_bwd", get_lib(), scales.device, N, _ptr(scales), _ptr(opacities),
                   _ptr(v_scales.contiguous()), _ptr(v_opacities.contiguous()), _ptr(v_scales_raw),
                   _ptr(v_opacities_raw))
        return v_scales_raw, v_opacities_raw


def splat_activations(scales_raw: Tensor, opacities_raw: Tensor) -> Tuple[Tensor, Tensor]:
    """`(torch.exp(scales_raw), torch.sigmoid(opacities_raw))` — scales_raw [N,3], opacities_raw [N]."""
    N = opacities_raw.shape[0]
    assert scales_raw.shape == (N, 3), scales_raw.shape
    assert opacities_raw.shape == (N,), opacities_raw.shape
    return _SplatActivations.apply(scales_raw.contiguous(), opacities_raw.contiguous())


def invert_poses(camtoworlds: Tensor) -> Tensor:
    """`torch.linalg.inv(camtoworlds)` [C,4,4] (gsplat_trainer.py:483) without the host
    synchronisation torch performs to report singular inputs.  Poses that require a gradient
    (pose optimisation) go through torch."""
    if camtoworlds.requires_grad:
        return torch.linalg.inv(camtoworlds)
    m = camtoworlds.contiguous()
    _check_cuda(m)
    _f32(m)
    assert m.shape[-2:] == (4, 4), m.shape
    out = torch.empty_like(m)
    if m.numel():
        native("invert_4x4", get_lib(), m.device, m.numel() // 16, _ptr(m), _ptr(out))
    return out


# ----------------------------------------------------------------------------------------
# rasterize_splats (gsplat_trainer.py:446-497)
# ----------------------------------------------------------------------------------------
def rasterize_splats(
    splats: Mapping[str, Tensor],
    camtoworlds: Tensor,  # [C, 4, 4]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    masks: Optional[Tensor] = None,
    camera_model: str = "pinhole",
    antialiased: bool = False,
    packed: bool = False,
    absgrad: bool = False,
    sparse_grad: bool = False,
    **kwargs,
) -> Tuple[Tensor, Tensor, Dict]:
    """`Runner.rasterize_splats` without the appearance module: `splats` holds the raw
    parameters `means [N,3], quats [N,4], scales [N,3] (log), opacities [N] (logit),
    sh0 [N,1,3], shN [N,K-1,3]`; `kwargs` go to `rasterization()` (sh_degree, near_plane,
    far_plane, render_mode, radius_clip, backgrounds, ...)."""
    scales, opacities = splat_activations(splats["scales"], splats["opacities"])
    render_colors, render_alphas, info = rasterization(
        means=splats["means"],
        quats=splats["quats"],
        scales=scales,
        opacities=opacities,
        colors=(splats["sh0"], splats["shN"]),  # == torch.cat([sh0, shN], 1), never materialised
        viewmats=invert_poses(camtoworlds),
        Ks=Ks,
        width=width,
        height=height,
        packed=packed,
        absgrad=absgrad,
        sparse_grad=sparse_grad,
        rasterize_mode="antialiased" if antialiased else "classic",
        distributed=False,
        camera_model=camera_model,
        **kwargs,
    )
    if masks is not None:
        render_colors = render_colors.masked_fill(~masks[..., None] if masks.dim() == 3 else ~masks, 0.0)
    return render_colors, render_alphas, info


# ----------------------------------------------------------------------------------------
# photometric loss (gsplat_trainer.py:624-628)
# ----------------------------------------------------------------------------------------
class _L1SSIMLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, colors: Tensor, pixels: Tensor, ssim_lambda: float):
        _check_cuda(colors, pixels)
        _f32(colors), _f32(pixels)
        lib = get_lib()
        C, H, W, _ = colors.shape
        dev = colors.device
        train = ctx.needs_input_grad[0]
        maps = torch.empty((3,) + tuple(colors.shape), device=dev, dtype=torch.float32) if train else None
        out3 = torch.empty((3,), device=dev, dtype=torch.float32)
        ws_bytes = lib.b200splat_l1_ssim_workspace_bytes(C, H, W)
        ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
        native("l1_ssim_fwd", lib, dev, C, H, W, _ptr(colors), _ptr(pixels), float(ssim_lambda),
               _ptr(maps[0]) if train else None, _ptr(maps[1]) if train else None, _ptr(maps[2]) if train else None,
               _ptr(out3), _ptr(ws), ws_bytes)
        ctx.save_for_backward(colors, pixels, maps)
        ctx.ssim_lambda = float(ssim_lambda)
        ctx.mark_non_differentiable(out3)
        return out3[0], out3

    @staticmethod
    def backward(ctx, v_loss: Tensor, _v_out3):
        colors, pixels, maps = ctx.saved_tensors
        C, H, W, _ = colors.shape
        v_colors = torch.empty_like(colors)
        v_loss = v_loss.contiguous().float()
        native("l1_ssim_bwd", get_lib(), colors.device, C, H, W, _ptr(colors), _ptr(pixels), ctx.ssim_lambda,
               _ptr(maps[0]), _ptr(maps[1]), _ptr(maps[2]), _ptr(v_loss), _ptr(v_colors))
        return v_colors, None, None


def l1_ssim_loss(colors: Tensor, pixels: Tensor, ssim_lambda: float = 0.2, return_terms: bool = False):
    """`F.l1_loss(colors, pixels) * (1 - ssim_lambda) + (1 - fused_ssim(colors.permute(0,3,1,2),
    pixels.permute(0,3,1,2), padding="valid")) * ssim_lambda` for [C,H,W,3] images, fused.

    Differentiable w.r.t. `colors` (like fused_ssim, which has no gradient for its second
    argument).  With `return_terms` also returns the device tensor `[loss, l1, ssim]`."""
    assert colors.dim() == 4 and colors.shape[-1] == 3, colors.shape
    assert pixels.shape == colors.shape, (pixels.shape, colors.shape)
    assert colors.shape[1] >= 11 and colors.shape[2] >= 11, "SSIM needs at least an 11x11 image"
    loss, out3 = _L1SSIMLoss.apply(colors.contiguous(), pixels.contiguous(), ssim_lambda)
    return (loss, out3) if return_terms else loss
