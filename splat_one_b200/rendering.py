"""`rasterization()` — host orchestration of the hot path, drop-in for
`gsplat.rendering.rasterization` (G/rendering.py:28-582, G = /root/reference/submodules/
gsplat/gsplat), which is the only entry point splat_one calls
(utils/gsplat_utils/gsplat_trainer.py:477-494).

Same signature, defaults, asserts, return arity and `meta` keys.  Differences, all
result-neutral:
  * per-Gaussian SH tables are evaluated against the [N,K,3] tensor directly instead of a
    C-fold `expand().contiguous()` copy (rendering.py:386 + _wrapper.py:71-73);
  * `distributed=True` (the reference's Gaussian-sharded all-to-all mode,
    rendering.py:279-294, 394-478) is not built: multi-GPU runs shard the *cameras* and
    all-reduce parameter gradients, see splat_one_b200/distributed.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor
from typing_extensions import Literal

from .wrapper import (
    fully_fused_projection,
    gather_rows,
    sh_view_colors_packed,
    sh_view_colors_packed_split,
    isect_tiles_and_offsets_begin,
    rasterize_to_pixels,
    sh_view_colors,
    sh_view_colors_split,
    spherical_harmonics,
    staged_colors_supported,
)


import os as _os

_ISECT_LATE = _os.environ.get("B200SPLAT_ISECT_LATE", "0") == "1"  # A/B: queue the intersection stage after the colours


def rasterization(
    means: Tensor,  # [N, 3]
    quats: Tensor,  # [N, 4]
    scales: Tensor,  # [N, 3]
    opacities: Tensor,  # [N]
    colors: Tensor,  # [(C,) N, D] or [(C,) N, K, 3]
    viewmats: Tensor,  # [C, 4, 4]
    Ks: Tensor,  # [C, 3, 3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: Literal["RGB", "D", "ED", "RGB+D", "RGB+ED"] = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: Literal["classic", "antialiased"] = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: Literal["pinhole", "ortho", "fisheye", "spherical"] = "pinhole",
    covars: Optional[Tensor] = None,
) -> Tuple[Tensor, Tensor, Dict]:
    """Rasterize N 3D Gaussians to C image planes.

    Returns `(render_colors [C,H,W,X], render_alphas [C,H,W,1], meta)` exactly as
    G/rendering.py:28-582 documents: X = D for "RGB", 1 for "D"/"ED", D+1 for
    "RGB+D"/"RGB+ED"; `meta` carries camera_ids, gaussian_ids, radii, means2d, depths,
    conics, opacities, tile_width, tile_height, tiles_per_gauss, isect_ids, flatten_ids,
    isect_offsets, width, height, tile_size, n_cameras.  `meta["means2d"]` is a graph
    tensor between projection and rasterization so `retain_grad()` / `.absgrad` work for
    the densification strategies (G/strategy/default.py:150, 221-226).
    """
    meta: Dict = {}

    # extension (splat_one_b200/step.py): `colors=(sh0, shN)` is the SH table of
    # gsplat_trainer.py:474 before its `torch.cat`; the colour stages read the two tensors in place
    # (pose gradients and over-long tables concatenate as the caller would have)
    sh_split = None
    if isinstance(colors, (tuple, list)):
        sh0, shN = colors
        assert sh_degree is not None, "a split (sh0, shN) table needs sh_degree"
        assert sh0.dim() == 3 and sh0.shape[1:] == (1, 3) and shN.dim() == 3 and shN.shape[2] == 3, (sh0.shape, shN.shape)
        assert sh0.shape[0] == shN.shape[0], (sh0.shape, shN.shape)
        if (not viewmats.requires_grad and sh0.is_cuda
                and (packed or staged_colors_supported(1 + shN.shape[1], True))):
            sh_split = (sh0, shN)
            colors = sh0  # placeholder for the shape asserts below
        else:
            colors = torch.cat([sh0, shN], 1)

    N = means.shape[0]
    C = viewmats.shape[0]
    assert means.shape == (N, 3), means.shape
    if covars is None:
        assert quats.shape == (N, 4), quats.shape
        assert scales.shape == (N, 3), scales.shape
    else:
        assert covars.shape == (N, 3, 3), covars.shape
        quats, scales = None, None
        # 3x3 -> flattened upper triangle (rendering.py:236-238)
        tri_indices = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
        covars = covars[..., tri_indices[0], tri_indices[1]]
    assert opacities.shape == (N,), opacities.shape
    assert viewmats.shape == (C, 4, 4), viewmats.shape
    assert Ks.shape == (C, 3, 3), Ks.shape
    assert render_mode in ["RGB", "D", "ED", "RGB+D", "RGB+ED"], render_mode

    if sh_degree is None:
        assert (colors.dim() == 2 and colors.shape[0] == N) or (
            colors.dim() == 3 and colors.shape[:2] == (C, N)
        ), colors.shape
    else:
        assert (colors.dim() == 3 and colors.shape[0] == N and colors.shape[2] == 3) or (
            colors.dim() == 4 and colors.shape[:2] == (C, N) and colors.shape[3] == 3
        ), colors.shape
        assert (sh_degree + 1) ** 2 <= (colors.shape[-2] if sh_split is None else 1 + sh_split[1].shape[1]), colors.shape

    if absgrad:
        assert not distributed, "AbsGrad is not supported in distributed mode."
    if distributed:
        raise NotImplementedError(
            "distributed=True (Gaussian-sharded all-to-all rendering) is not part of this build; "
            "use splat_one_b200.distributed (camera-sharded data parallelism + gradient all-reduce).")

    # ---- projection (a2 / a4) -------------------------------------------------------
    proj_results = fully_fused_projection(
        means, covars, quats, scales, viewmats, Ks, width, height,
        eps2d=eps2d, packed=packed, near_plane=near_plane, far_plane=far_plane,
        radius_clip=radius_clip, sparse_grad=sparse_grad,
        calc_compensations=(rasterize_mode == "antialiased"), camera_model=camera_model,
        # with SH colours `means` also receives a dense gradient from the view directions
        _dense_means_grad=(sparse_grad and sh_degree is not None),
    )
    if packed:
        camera_ids, gaussian_ids, radii, means2d, depths, conics, compensations = proj_results
        opacities = gather_rows(opacities, gaussian_ids)  # [nnz]
    else:
        radii, means2d, depths, conics, compensations = proj_results
        opacities = opacities[None].expand(C, -1)  # [C, N]; no copy when C == 1 (reference: .repeat)
        camera_ids, gaussian_ids = None, None

    if compensations is not None:
        opacities = opacities * compensations

    meta.update({
        "camera_ids": camera_ids,
        "gaussian_ids": gaussian_ids,
        "radii": radii,
        "means2d": means2d,
        "depths": depths,
        "conics": conics,
        "opacities": opacities,
    })

    # ---- tile intersection, first half (a6): count, depth order and the n_isects read-back are
    # queued here, so the host's wait for n_isects overlaps the colour stage as well.
    # (concurrent=True would put this half on a side stream next to the colour kernels: measured
    # no gain at config B — both are bandwidth-bound — and 0.4 ms worse with host copies in flight.)
    tile_width = math.ceil(width / float(tile_size))
    tile_height = math.ceil(height / float(tile_size))
    isect_args = dict(packed=packed, n_cameras=C, camera_ids=camera_ids, gaussian_ids=gaussian_ids, concurrent=False)
    isect_finish = None
    if not _ISECT_LATE:
        isect_finish = isect_tiles_and_offsets_begin(means2d, radii, depths, tile_size, tile_width, tile_height,
                                                     **isect_args)

    # ---- colours (a5) -----------------------------------------------------------------
    if sh_degree is None:
        if packed:
            colors = gather_rows(colors, gaussian_ids) if colors.dim() == 2 else colors[camera_ids, gaussian_ids]
        else:
            colors = colors.expand(C, -1, -1) if colors.dim() == 2 else colors
    else:
        if sh_split is not None and packed:
            colors = sh_view_colors_packed_split(sh_degree, means, viewmats, sh_split[0], sh_split[1], camera_ids,
                                                 gaussian_ids)  # [nnz, 3]
        elif sh_split is not None:
            colors = sh_view_colors_split(sh_degree, means, viewmats, sh_split[0], sh_split[1], radii)  # [C, N, 3]
        elif packed and not viewmats.requires_grad:
            # same maths, one fused kernel per direction over the COO rows
            colors = sh_view_colors_packed(sh_degree, means, viewmats, colors, camera_ids, gaussian_ids)  # [nnz, 3]
        elif viewmats.requires_grad:
            camtoworlds = torch.inverse(viewmats)  # [C, 4, 4]
            if packed:
                dirs = means[gaussian_ids, :] - camtoworlds[camera_ids, :3, 3]  # [nnz, 3]
                shs = colors[gaussian_ids, :, :] if colors.dim() == 3 else colors[camera_ids, gaussian_ids, :, :]
            else:
                dirs = means[None, :, :] - camtoworlds[:, None, :3, 3]  # [C, N, 3]
                shs = colors.expand(C, -1, -1, -1) if colors.dim() == 3 else colors
            colors = spherical_harmonics(sh_degree, dirs, shs, masks=radii > 0)
            colors = torch.clamp_min(colors + 0.5, 0.0)  # rendering.py:392
        else:
            # same maths, one fused kernel per direction (no dirs / mask / +0.5 / clamp passes)
            colors = sh_view_colors(sh_degree, means, viewmats, colors, radii)  # [C, N, 3]

    # ---- depth channel (rendering.py:481-492) ---------------------------------------
    if render_mode in ["RGB+D", "RGB+ED"]:
        colors = torch.cat((colors, depths[..., None]), dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros(C, 1, device=backgrounds.device)], dim=-1)
    elif render_mode in ["D", "ED"]:
        colors = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros(C, 1, device=backgrounds.device)

    # ---- tile intersection, second half (a6, a7): tile order, ids, offsets ------------------
    if isect_finish is None:
        isect_finish = isect_tiles_and_offsets_begin(means2d, radii, depths, tile_size, tile_width, tile_height,
                                                     **isect_args)
    tiles_per_gauss, isect_ids, flatten_ids, isect_offsets = isect_finish()

    meta.update({
        "tile_width": tile_width,
        "tile_height": tile_height,
        "tiles_per_gauss": tiles_per_gauss,
        "isect_ids": isect_ids,
        "flatten_ids": flatten_ids,
        "isect_offsets": isect_offsets,
        "width": width,
        "height": height,
        "tile_size": tile_size,
        "n_cameras": C,
    })

    # ---- rasterize (a8) ----------------------------------------------------------------
    if colors.shape[-1] > channel_chunk:
        n_chunks = (colors.shape[-1] + channel_chunk - 1) // channel_chunk
        render_colors, render_alphas = [], []
        for i in range(n_chunks):
            colors_chunk = colors[..., i * channel_chunk:(i + 1) * channel_chunk]
            backgrounds_chunk = (
                backgrounds[..., i * channel_chunk:(i + 1) * channel_chunk] if backgrounds is not None else None
            )
            render_colors_, render_alphas_ = rasterize_to_pixels(
                means2d, conics, colors_chunk, opacities, width, height, tile_size, isect_offsets, flatten_ids,
                backgrounds=backgrounds_chunk, packed=packed, absgrad=absgrad,
            )
            render_colors.append(render_colors_)
            render_alphas.append(render_alphas_)
        render_colors = torch.cat(render_colors, dim=-1)
        render_alphas = render_alphas[0]  # discard the rest
    else:
        render_colors, render_alphas = rasterize_to_pixels(
            means2d, conics, colors, opacities, width, height, tile_size, isect_offsets, flatten_ids,
            backgrounds=backgrounds, packed=packed, absgrad=absgrad,
        )
    if render_mode in ["ED", "RGB+ED"]:
        # accumulated depth -> expected depth (rendering.py:572-580)
        render_colors = torch.cat(
            [render_colors[..., :-1], render_colors[..., -1:] / render_alphas.clamp(min=1e-10)], dim=-1
        )

    return render_colors, render_alphas, meta
