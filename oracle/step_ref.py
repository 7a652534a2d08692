"""ORACLE (test infrastructure only — never imported by splat_one_b200/) for SURVEY.md §8 f4:
the training step either side of rasterization(), restated on the CPU in plain PyTorch.

R = /root/reference.  `splat_activations` and `rasterize_splats` follow
R/utils/gsplat_utils/gsplat_trainer.py:446-497 line by line.  `l1_ssim_loss` follows :624-628;
its `fused_ssim` is a THIRD-PARTY dependency of the reference (github.com/rahul-goel/fused-ssim,
imported at gsplat_trainer.py:33, unpinned, NOT vendored under /root/reference and not
installed here), so its published algorithm is restated: SSIM of Wang et al. 2004 with an
11x11 Gaussian window (sigma 1.5, normalised), C1 = 0.01², C2 = 0.03², local statistics by
zero-padded correlation per channel, `padding="valid"` = crop 5 pixels from every border
before the mean.  PARITY UNPINNED for the SSIM term: no executed fused_ssim output exists in
this container; tests/test_oracle_next.py cross-checks this restatement against an independent
float64 scipy implementation and known SSIM identities (ssim(x,x) = 1, symmetry, range).
"""
import math

import torch
import torch.nn.functional as F

from . import torch_ref as O


def splat_activations(scales_raw, opacities_raw):
    """gsplat_trainer.py:458-459."""
    return torch.exp(scales_raw), torch.sigmoid(opacities_raw)


def gaussian_window(size: int = 11, sigma: float = 1.5, dtype=torch.float32):
    x = torch.arange(size, dtype=torch.float64) - size // 2
    g = torch.exp(-(x * x) / (2.0 * sigma * sigma))
    return (g / g.sum()).to(dtype)


def ssim_map(img1, img2, padding: str = "valid"):
    """[B,3,H,W] -> per-pixel, per-channel SSIM map (fused_ssim's `ssim_map`)."""
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ch = img1.shape[1]
    g = gaussian_window(dtype=img1.dtype)
    w2d = (g[:, None] * g[None, :])[None, None].expand(ch, 1, -1, -1).contiguous()

    def blur(x):
        return F.conv2d(x, w2d, padding=5, groups=ch)  # zero padding, like the fused kernel

    mu1, mu2 = blur(img1), blur(img2)
    sigma1_sq = blur(img1 * img1) - mu1 * mu1
    sigma2_sq = blur(img2 * img2) - mu2 * mu2
    sigma12 = blur(img1 * img2) - mu1 * mu2
    m = ((2 * mu1 * mu2 + C1) * (2 * sigma12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (sigma1_sq + sigma2_sq + C2))
    if padding == "valid":
        m = m[:, :, 5:-5, 5:-5]
    return m


def fused_ssim(img1, img2, padding: str = "valid"):
    return ssim_map(img1, img2, padding).mean()


def l1_ssim_loss(colors, pixels, ssim_lambda: float = 0.2):
    """gsplat_trainer.py:624-628 for [C,H,W,3] images; returns (loss, l1, ssim)."""
    l1 = F.l1_loss(colors, pixels)
    ssim = fused_ssim(colors.permute(0, 3, 1, 2), pixels.permute(0, 3, 1, 2), padding="valid")
    return l1 * (1.0 - ssim_lambda) + (1.0 - ssim) * ssim_lambda, l1, ssim


def rasterize_splats(splats, camtoworlds, Ks, width, height, masks=None, camera_model="pinhole", antialiased=False,
                     raster_fn=None, **kwargs):
    """gsplat_trainer.py:446-497 (no appearance module) over the oracle rasterization."""
    scales, opacities = splat_activations(splats["scales"], splats["opacities"])
    colors = torch.cat([splats["sh0"], splats["shN"]], 1)
    extra = {} if raster_fn is None else {"raster_fn": raster_fn}
    rc, ra, info = O.rasterization(
        splats["means"], splats["quats"], scales, opacities, colors, torch.linalg.inv(camtoworlds), Ks, width, height,
        rasterize_mode="antialiased" if antialiased else "classic", camera_model=camera_model, **extra, **kwargs)
    if masks is not None:
        rc = rc.clone()
        rc[~masks] = 0
    return rc, ra, info
