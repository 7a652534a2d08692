"""Seeded synthetic scenes for benchmarks and parity tests.

All tensors are generated on the CPU with an explicit generator (so the same scene can
be fed to the CPU oracle and, after `.cuda()`, to the kernels) and returned as a dict.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

SH_C0 = 0.28209479177387814


def pinhole_scene(n: int, width: int, height: int, seed: int = 42, sh_degree: int = 3,
                  footprint_px: float = 3.0, footprint_sigma: float = 0.9, n_cameras: int = 1) -> Dict:
    """Gaussians filling the frustum of an identity-pose pinhole camera (fx = fy = W).

    z ~ U[1,20]; (x,y) uniform over the frustum cross-section at depth z enlarged by 5 %;
    pixel footprint s_px ~ LogNormal(ln footprint_px, footprint_sigma), world scale =
    s_px·z/fx·a with a ~ U[0.3,1]^3; quats ~ N(0,I) (un-normalised); opacities ~ U(0,1);
    sh0 = (U(0,1)-0.5)/C0, shN ~ N(0, 0.05²).  Extra cameras (n_cameras > 1) are small
    rotations/translations of the first so that every camera still sees the cloud.
    """
    g = torch.Generator().manual_seed(seed)
    fx = fy = float(width)
    cx, cy = width / 2.0, height / 2.0
    z = torch.rand(n, generator=g) * 19.0 + 1.0
    u = (torch.rand(n, generator=g) - 0.5) * 1.05 * width
    v = (torch.rand(n, generator=g) - 0.5) * 1.05 * height
    means = torch.stack([u * z / fx, v * z / fy, z], dim=-1)
    s_px = torch.exp(torch.randn(n, generator=g) * footprint_sigma + math.log(footprint_px))
    a = torch.rand(n, 3, generator=g) * 0.7 + 0.3
    scales = (s_px * z / fx)[:, None] * a
    quats = torch.randn(n, 4, generator=g)
    opacities = torch.rand(n, generator=g)
    K = (sh_degree + 1) ** 2
    sh = torch.randn(n, K, 3, generator=g) * 0.05
    sh[:, 0, :] = (torch.rand(n, 3, generator=g) - 0.5) / SH_C0
    Ks = torch.tensor([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]]).expand(n_cameras, -1, -1).contiguous()
    viewmats = torch.eye(4).expand(n_cameras, -1, -1).contiguous().clone()
    for c in range(1, n_cameras):
        ang = 0.02 * c * (-1) ** c
        viewmats[c, :3, :3] = torch.tensor([[math.cos(ang), 0.0, math.sin(ang)], [0.0, 1.0, 0.0],
                                            [-math.sin(ang), 0.0, math.cos(ang)]])
        viewmats[c, :3, 3] = torch.tensor([0.05 * c, -0.03 * c, 0.02 * c])
    return dict(means=means, quats=quats, scales=scales, opacities=opacities, sh=sh, viewmats=viewmats, Ks=Ks,
                width=width, height=height, sh_degree=sh_degree, camera_model="pinhole")


def spherical_scene(n: int, width: int, height: int, seed: int = 42, sh_degree: int = 3,
                    footprint_px: float = 3.0, footprint_sigma: float = 0.9) -> Dict:
    """Config D: directions uniform on the sphere, r ~ U[1,20], equirectangular W×H,
    dummy K (R/utils/datasets/opensfm.py:186-192)."""
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    r = torch.rand(n, generator=g) * 19.0 + 1.0
    means = d * r[:, None]
    px_per_rad = width / (2 * math.pi)
    s_px = torch.exp(torch.randn(n, generator=g) * footprint_sigma + math.log(footprint_px))
    a = torch.rand(n, 3, generator=g) * 0.7 + 0.3
    scales = (s_px * r / px_per_rad)[:, None] * a
    quats = torch.randn(n, 4, generator=g)
    opacities = torch.rand(n, generator=g)
    K = (sh_degree + 1) ** 2
    sh = torch.randn(n, K, 3, generator=g) * 0.05
    sh[:, 0, :] = (torch.rand(n, 3, generator=g) - 0.5) / SH_C0
    Ks = torch.tensor([[width / 2.0, 0.0, width / 2.0], [0.0, width / 2.0, height / 2.0], [0.0, 0.0, 1.0]])[None]
    viewmats = torch.eye(4)[None].clone()
    return dict(means=means, quats=quats, scales=scales, opacities=opacities, sh=sh, viewmats=viewmats, Ks=Ks,
                width=width, height=height, sh_degree=sh_degree, camera_model="spherical")


def to_device(scene: Dict, device) -> Dict:
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in scene.items()}
