"""ORACLE (test infrastructure, not product code) — CPU restatement of the rasterization
hot path of inuex35/splat_one's gsplat fork, in PyTorch/numpy, float32 by default.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this package; the product (`splat_one_b200/`) never does.

Parity status: PINNED.
  * pinhole / ortho / fisheye projection, SH, tile intersection, offset encode and the un-fused
    operators against the reference's own pure-PyTorch implementation (`gsplat/cuda/_torch_impl.py`,
    imported from /root/reference by `oracle/gen_golden.py`; vectors under tests/golden/*_ref.npz);
  * `rasterize_to_pixels` against the reference's own `_rasterize_to_pixels` run on the CPU
    (`raster_ref_d1/d3.npz`; its nerfacc dependency restated in oracle/nerfacc_stub.py);
  * what the reference's CPU code cannot pin — `camera_model="spherical"` (forward, closed-form
    VJP, packed rules), the packed projection's rules and the CUDA rasterizer itself — against
    vectors produced on a B200 by the reference's OWN CUDA kernels (`oracle/build_ref.py` builds
    them into oracle/_ref, `oracle/gen_golden_refcuda.py` wrote tests/golden/refcuda_*.npz;
    tests/test_refcuda_golden.py checks this file against them in the CPU suite).
`selective_adam_update` / `compute_relocation` (CUDA only in the reference) are restated from source here
and pinned on the GPU box by tests/test_gpu_vs_reference.py.  Still unpinned (un-vendored
dependency, restated from its publication): the SSIM of oracle/step_ref.py.

Where `_torch_impl` and the fork's CUDA disagree, the CUDA is what splat_one executes and
what is restated (`fork_faithful=True`, SURVEY.md §8c list); `fork_faithful=False`
reproduces `_torch_impl` so the golden vectors generated from it can be checked.

CS = /root/reference/submodules/gsplat/gsplat/cuda/csrc, G = .../gsplat/gsplat.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

ALPHA_MAX = 0.999          # CS/rasterize_to_pixels_fwd.cu:146
ALPHA_MIN = 1.0 / 255.0    # :147
T_EPS = 1e-4               # :152


# ----------------------------------------------------------------------------------------
# a2/a3/a4: projection
# ----------------------------------------------------------------------------------------
def quat_to_rotmat(quats: Tensor) -> Tensor:
    """CS/utils.cuh:15-37 (normalise, then the standard wxyz rotation matrix)."""
    q = F.normalize(quats, p=2, dim=-1)
    w, x, y, z = torch.unbind(q, dim=-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y),
    ], dim=-1)
    return R.reshape(quats.shape[:-1] + (3, 3))


def quat_scale_to_covar(quats: Tensor, scales: Tensor) -> Tensor:
    """Σ = (R S)(R S)ᵀ, CS/utils.cuh:66-97."""
    M = quat_to_rotmat(quats) * scales[..., None, :]
    return M @ M.transpose(-1, -2)


def covar_from_triu(c6: Tensor) -> Tensor:
    """[N,6] upper triangle -> [N,3,3] (CS/fully_fused_projection_fwd.cu:88-100)."""
    idx = torch.tensor([0, 1, 2, 1, 3, 4, 2, 4, 5])
    return c6[..., idx].reshape(c6.shape[:-1] + (3, 3))


def world_to_cam(means: Tensor, covars: Tensor, viewmats: Tensor) -> Tuple[Tensor, Tensor]:
    """CS/utils.cuh:598-606, 629-636."""
    R = viewmats[:, :3, :3]
    t = viewmats[:, :3, 3]
    means_c = torch.einsum("cij,nj->cni", R, means) + t[:, None, :]
    covars_c = torch.einsum("cij,njk,clk->cnil", R, covars, R)
    return means_c, covars_c


def _jcjt(J: Tensor, covars_c: Tensor) -> Tensor:
    return torch.einsum("...ij,...jk,...kl->...il", J, covars_c, J.transpose(-1, -2))


def persp_proj(means_c, covars_c, Ks, width, height, fork_faithful=True):
    """CS/utils.cuh:254-293.  fork: means2d uses fx, fy, cx, cy only (no skew)."""
    tx, ty, tz = torch.unbind(means_c, dim=-1)
    fx, fy = Ks[..., 0, 0, None], Ks[..., 1, 1, None]
    cx, cy = Ks[..., 0, 2, None], Ks[..., 1, 2, None]
    tan_fovx, tan_fovy = 0.5 * width / fx, 0.5 * height / fy
    lim_x_pos = (width - cx) / fx + 0.3 * tan_fovx
    lim_x_neg = cx / fx + 0.3 * tan_fovx
    lim_y_pos = (height - cy) / fy + 0.3 * tan_fovy
    lim_y_neg = cy / fy + 0.3 * tan_fovy
    txc = tz * torch.minimum(lim_x_pos, torch.maximum(-lim_x_neg, tx / tz))
    tyc = tz * torch.minimum(lim_y_pos, torch.maximum(-lim_y_neg, ty / tz))
    O = torch.zeros_like(tz)
    tz2 = tz * tz
    J = torch.stack([fx / tz, O, -fx * txc / tz2, O, fy / tz, -fy * tyc / tz2], dim=-1)
    J = J.reshape(means_c.shape[:-1] + (2, 3))
    cov2d = _jcjt(J, covars_c)
    if fork_faithful:
        means2d = torch.stack([fx * tx / tz + cx, fy * ty / tz + cy], dim=-1)
    else:  # G/cuda/_torch_impl.py:118-119
        means2d = torch.einsum("cij,cnj->cni", Ks[:, :2, :3], means_c) / tz[..., None]
    return means2d, cov2d


def ortho_proj(means_c, covars_c, Ks, width, height):
    """CS/utils.cuh:183-210."""
    fx, fy = Ks[..., 0, 0, None], Ks[..., 1, 1, None]
    cx, cy = Ks[..., 0, 2, None], Ks[..., 1, 2, None]
    O = torch.zeros_like(means_c[..., 0])
    J = torch.stack([fx + O, O, O, O, fy + O, O], dim=-1).reshape(means_c.shape[:-1] + (2, 3))
    cov2d = _jcjt(J, covars_c)
    means2d = torch.stack([fx * means_c[..., 0] + cx, fy * means_c[..., 1] + cy], dim=-1)
    return means2d, cov2d


def fisheye_proj(means_c, covars_c, Ks, width, height):
    """CS/utils.cuh:376-415."""
    x, y, z = torch.unbind(means_c, dim=-1)
    fx, fy = Ks[..., 0, 0, None], Ks[..., 1, 1, None]
    cx, cy = Ks[..., 0, 2, None], Ks[..., 1, 2, None]
    eps = 0.0000001
    xy_len = (x * x + y * y) ** 0.5 + eps
    theta = torch.atan2(xy_len, z + eps)
    means2d = torch.stack([x * fx * theta / xy_len + cx, y * fy * theta / xy_len + cy], dim=-1)
    x2 = x * x + eps
    y2 = y * y
    xy = x * y
    x2y2 = x2 + y2
    x2y2z2_inv = 1.0 / (x2y2 + z * z)
    b = torch.atan2(xy_len, z) / xy_len / x2y2
    a = z * x2y2z2_inv / x2y2
    J = torch.stack([
        fx * (x2 * a + y2 * b), fx * xy * (a - b), -fx * x * x2y2z2_inv,
        fy * xy * (a - b), fy * (y2 * a + x2 * b), -fy * y * x2y2z2_inv,
    ], dim=-1).reshape(means_c.shape[:-1] + (2, 3))
    return means2d, _jcjt(J, covars_c)


def spherical_proj(means_c, covars_c, width, height, fork_faithful=True):
    """CS/utils.cuh:520-557 (fork): lon = atan2(x, z), lat = asin(y / r); J carries the
    +1e-8 guards; π factors are evaluated in double and rounded once.
    Gradient spec (CS/utils.cuh:559-594): only Jᵀ·v_mean2d and Jᵀ·v_cov2d·J — the Jacobian
    is treated as a constant, so J is detached and means2d gets a straight-through
    gradient through J.  (The CUDA VJP recomputes r with +1e-8 inside the sqrt; the
    relative effect is <= 1e-8/r² and is ignored here.)"""
    x, y, z = torch.unbind(means_c, dim=-1)
    r = torch.sqrt(x * x + y * y + z * z)
    xz_norm = torch.sqrt(x * x + z * z + 1e-8)
    denom_xz = x * x + z * z + 1e-8
    denom_r2 = r * r + 1e-8
    lon = torch.atan2(x, z)
    if fork_faithful:
        lat = torch.asin(y / r)
    else:  # G/cuda/_torch_impl.py:254
        lat = torch.atan2(y, xz_norm)
    nlat = (lat.double() / (math.pi / 2.0)).to(means_c.dtype)
    nlon = (lon.double() / math.pi).to(means_c.dtype)
    means2d = torch.stack([(nlon + 1) * width / 2, (nlat + 1) * height / 2], dim=-1)
    O = torch.zeros_like(x)
    w2pi = width / (2 * math.pi)
    hpi = height / math.pi

    def d(v):  # double product, single rounding (mat3x2<float> constructor)
        return v.double()

    J = torch.stack([
        (w2pi * d(z / denom_xz)), (O.double()), (w2pi * d(-x / denom_xz)),
        (hpi * d(-(x * y) / (denom_r2 * xz_norm))), (hpi * d(xz_norm / denom_r2)),
        (hpi * d(-(z * y) / (denom_r2 * xz_norm))),
    ], dim=-1).to(means_c.dtype).reshape(means_c.shape[:-1] + (2, 3))
    if fork_faithful:
        J = J.detach()
        lin = torch.einsum("...ij,...j->...i", J, means_c)
        means2d = means2d.detach() + (lin - lin.detach())
    return means2d, _jcjt(J, covars_c)


def fully_fused_projection(
    means: Tensor, covars: Optional[Tensor], quats: Optional[Tensor], scales: Optional[Tensor],
    viewmats: Tensor, Ks: Tensor, width: int, height: int, eps2d: float = 0.3,
    near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0,
    calc_compensations: bool = False, camera_model: str = "pinhole",
    packed_rules: bool = False, fork_faithful: bool = True,
):
    """Dense [C,N] projection with the fork's culling / depth / radius rules.

    covars: [N,6] triu or [N,3,3] or None.  Returns (radii int32 [C,N], means2d, depths,
    conics, compensations|None); entries with radii == 0 are zeroed.
    unpacked rules: CS/fully_fused_projection_fwd.cu:73-85, 171-209;
    packed rules (`packed_rules=True`): CS/fully_fused_projection_packed_fwd.cu:186-202, 249.
    """
    if covars is None:
        covars3 = quat_scale_to_covar(quats, scales)
    elif covars.shape[-1] == 6:
        covars3 = covar_from_triu(covars)
    else:
        covars3 = covars
    means_c, covars_c = world_to_cam(means, covars3, viewmats)
    if camera_model == "pinhole":
        means2d, cov2d = persp_proj(means_c, covars_c, Ks, width, height, fork_faithful)
    elif camera_model == "ortho":
        means2d, cov2d = ortho_proj(means_c, covars_c, Ks, width, height)
    elif camera_model == "fisheye":
        means2d, cov2d = fisheye_proj(means_c, covars_c, Ks, width, height)
    elif camera_model == "spherical":
        means2d, cov2d = spherical_proj(means_c, covars_c, width, height, fork_faithful)
    else:
        raise ValueError(camera_model)

    det_orig = cov2d[..., 0, 0] * cov2d[..., 1, 1] - cov2d[..., 0, 1] * cov2d[..., 1, 0]
    cov2d = cov2d + torch.eye(2, dtype=means.dtype) * eps2d
    det = cov2d[..., 0, 0] * cov2d[..., 1, 1] - cov2d[..., 0, 1] * cov2d[..., 1, 0]
    if not fork_faithful:
        det = det.clamp(min=1e-10)  # G/cuda/_torch_impl.py:344
    compensations = torch.sqrt(torch.clamp(det_orig / det, min=0.0)) if calc_compensations else None
    det_safe = torch.where(det > 0, det, torch.ones_like(det))
    conics = torch.stack([
        cov2d[..., 1, 1] / det_safe,
        -(cov2d[..., 0, 1] + cov2d[..., 1, 0]) / 2.0 / det_safe,
        cov2d[..., 0, 0] / det_safe,
    ], dim=-1)
    conics = torch.where((det > 0)[..., None], conics, torch.zeros_like(conics))

    z = means_c[..., 2]
    rnorm = torch.sqrt((means_c * means_c).sum(-1))
    if fork_faithful and not packed_rules:
        # value |mean_c|, gradient to z only (CS/...fwd.cu:204-209, CS/...bwd.cu:195)
        depths = z + (rnorm - z).detach()
    else:
        depths = z

    b = (cov2d[..., 0, 0] + cov2d[..., 1, 1]) / 2
    with torch.no_grad():
        if fork_faithful and packed_rules:
            disc = torch.sqrt(torch.clamp(b * b - det, min=0.1))
            radius = torch.ceil(3.0 * torch.sqrt(torch.maximum(b + disc, b - disc)))
        else:
            radius = torch.ceil(3.0 * torch.sqrt(b + torch.sqrt(torch.clamp(b * b - det, min=0.01))))
        if fork_faithful:
            if camera_model == "spherical":
                valid = ~((rnorm < near_plane) | (rnorm > far_plane))
            else:
                valid = ~((z < near_plane) | (z > far_plane))
            if packed_rules:
                valid &= det > 0
            valid &= ~(radius <= radius_clip)
            if camera_model != "spherical":
                valid &= ~((means2d[..., 0] + radius <= 0) | (means2d[..., 0] - radius >= width)
                           | (means2d[..., 1] + radius <= 0) | (means2d[..., 1] - radius >= height))
        else:  # G/cuda/_torch_impl.py:368-377
            valid = (det > 0) & (z > near_plane) & (z < far_plane)
            radius = torch.where(valid, radius, torch.zeros_like(radius))
            valid &= ((means2d[..., 0] + radius > 0) & (means2d[..., 0] - radius < width)
                      & (means2d[..., 1] + radius > 0) & (means2d[..., 1] - radius < height))
        radius = torch.where(valid, radius, torch.zeros_like(radius))
        radii = radius.int()
    return radii, means2d, depths, conics, compensations


def pack_projection(radii, means2d, depths, conics, compensations):
    """Dense [C,N] -> COO rows in (camera, gaussian) row-major order
    (CS/fully_fused_projection_packed_fwd.cu:233-256)."""
    sel = radii > 0
    camera_ids, gaussian_ids = torch.nonzero(sel, as_tuple=True)
    comp = compensations[sel] if compensations is not None else None
    return camera_ids, gaussian_ids, radii[sel], means2d[sel], depths[sel], conics[sel], comp


# ----------------------------------------------------------------------------------------
# a5: spherical harmonics
# ----------------------------------------------------------------------------------------
def eval_sh_bases(basis_dim: int, dirs: Tensor) -> Tensor:
    """Sloan's closed forms (CS/spherical_harmonics.cuh:17-105), unit directions."""
    result = torch.empty((*dirs.shape[:-1], basis_dim), dtype=dirs.dtype)
    result[..., 0] = 0.2820947917738781
    if basis_dim <= 1:
        return result
    x, y, z = dirs.unbind(-1)
    fTmpA = -0.48860251190292
    result[..., 2] = -fTmpA * z
    result[..., 3] = fTmpA * x
    result[..., 1] = fTmpA * y
    if basis_dim <= 4:
        return result
    z2 = z * z
    fTmpB = -1.092548430592079 * z
    fTmpA = 0.5462742152960395
    fC1 = x * x - y * y
    fS1 = 2 * x * y
    result[..., 6] = 0.9461746957575601 * z2 - 0.3153915652525201
    result[..., 7] = fTmpB * x
    result[..., 5] = fTmpB * y
    result[..., 8] = fTmpA * fC1
    result[..., 4] = fTmpA * fS1
    if basis_dim <= 9:
        return result
    fTmpC = -2.285228997322329 * z2 + 0.4570457994644658
    fTmpB = 1.445305721320277 * z
    fTmpA = -0.5900435899266435
    fC2 = x * fC1 - y * fS1
    fS2 = x * fS1 + y * fC1
    result[..., 12] = z * (1.865881662950577 * z2 - 1.119528997770346)
    result[..., 13] = fTmpC * x
    result[..., 11] = fTmpC * y
    result[..., 14] = fTmpB * fC1
    result[..., 10] = fTmpB * fS1
    result[..., 15] = fTmpA * fC2
    result[..., 9] = fTmpA * fS2
    if basis_dim <= 16:
        return result
    fTmpD = z * (-4.683325804901025 * z2 + 2.007139630671868)
    fTmpC = 3.31161143515146 * z2 - 0.47308734787878
    fTmpB = -1.770130769779931 * z
    fTmpA = 0.6258357354491763
    fC3 = x * fC2 - y * fS2
    fS3 = x * fS2 + y * fC2
    result[..., 20] = 1.984313483298443 * z * result[..., 12].clone() - 1.006230589874905 * result[..., 6].clone()
    result[..., 21] = fTmpD * x
    result[..., 19] = fTmpD * y
    result[..., 22] = fTmpC * fC1
    result[..., 18] = fTmpC * fS1
    result[..., 23] = fTmpB * fC2
    result[..., 17] = fTmpB * fS2
    result[..., 24] = fTmpA * fC3
    result[..., 16] = fTmpA * fS3
    return result


def spherical_harmonics(degree: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """CS/compute_sh_fwd.cu:12-38: normalise dirs, dot the first (degree+1)² bases with the
    coefficients; masked elements give 0 (the CUDA leaves them uninitialised)."""
    d = F.normalize(dirs, p=2, dim=-1)
    nb = (degree + 1) ** 2
    bases = eval_sh_bases(nb, d)
    out = (bases[..., None] * coeffs[..., :nb, :]).sum(dim=-2)
    if masks is not None:
        out = torch.where(masks[..., None], out, torch.zeros_like(out))
    return out


# ----------------------------------------------------------------------------------------
# a6/a7: tile intersection, sort, offsets — integer-exact
# ----------------------------------------------------------------------------------------
def isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height, sort=True,
                packed=False, n_cameras=None, camera_ids=None, gaussian_ids=None):
    """CS/isect_tiles.cu:17-105 + the stable radix sort at :252-300, vectorised in numpy.

    Returns (tiles_per_gauss int32, isect_ids int64, flatten_ids int32) as torch tensors."""
    m = means2d.detach().cpu().numpy().astype(np.float32).reshape(-1, 2)
    r = radii.detach().cpu().numpy().astype(np.int32).reshape(-1)
    dp = depths.detach().cpu().numpy().astype(np.float32).reshape(-1)
    if packed:
        C = int(n_cameras)
        cid = camera_ids.detach().cpu().numpy().astype(np.int64).reshape(-1)
    else:
        C, N = means2d.shape[:2]
        cid = np.repeat(np.arange(C, dtype=np.int64), N)
    ts = np.float32(tile_size)
    rf = r.astype(np.float32)
    tr = rf / ts
    tx = m[:, 0] / ts
    ty = m[:, 1] / ts

    def u32_sat(v):  # cvt.rzi.u32.f32 saturating conversion (SURVEY.md §8a note vi)
        v = np.nan_to_num(v, nan=0.0)
        return np.clip(v, 0, 4294967295.0).astype(np.int64)

    x0 = np.minimum(u32_sat(np.floor(tx - tr)), tile_width)
    y0 = np.minimum(u32_sat(np.floor(ty - tr)), tile_height)
    x1 = np.minimum(u32_sat(np.ceil(tx + tr)), tile_width)
    y1 = np.minimum(u32_sat(np.ceil(ty + tr)), tile_height)
    # uint32 arithmetic of the kernel: (y1-y0)*(x1-x0) wraps, but x1>=x0, y1>=y0 always hold
    tpg = ((y1 - y0) * (x1 - x0)).astype(np.int64)
    tpg[r <= 0] = 0
    n_isects = int(tpg.sum())
    n_tiles = tile_width * tile_height
    tile_n_bits = int(n_tiles).bit_length()
    cam_n_bits = int(C).bit_length()
    idx = np.repeat(np.arange(len(r), dtype=np.int64), tpg)
    start = np.cumsum(tpg) - tpg
    local = np.arange(n_isects, dtype=np.int64) - np.repeat(start, tpg)
    w = (x1 - x0)[idx]
    w_safe = np.maximum(w, 1)
    tyy = y0[idx] + local // w_safe
    txx = x0[idx] + local % w_safe
    tile_id = tyy * tile_width + txx
    depth_bits = dp.view(np.int32).astype(np.int64)  # sign-extending, CS/isect_tiles.cu:92
    keys = (cid[idx] << np.int64(32 + tile_n_bits)) | (tile_id << np.int64(32)) | depth_bits[idx]
    vals = idx.astype(np.int32)
    if sort and n_isects:
        end_bit = 32 + tile_n_bits + cam_n_bits
        mask = np.uint64((1 << end_bit) - 1) if end_bit < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
        order = np.argsort(keys.view(np.uint64) & mask, kind="stable")
        keys, vals = keys[order], vals[order]
    shape = radii.shape
    return (torch.from_numpy(tpg.astype(np.int32).reshape(shape)), torch.from_numpy(keys),
            torch.from_numpy(vals))


def isect_offset_encode(isect_ids: Tensor, n_cameras: int, tile_width: int, tile_height: int) -> Tensor:
    """CS/isect_tiles.cu:309-355: offsets[k] = first sorted index whose (cam, tile) >= k."""
    n_tiles = tile_width * tile_height
    tile_n_bits = int(n_tiles).bit_length()
    hi = isect_ids.numpy() >> 32
    ids = (hi >> tile_n_bits) * n_tiles + (hi & ((1 << tile_n_bits) - 1))
    off = np.searchsorted(ids, np.arange(n_cameras * n_tiles, dtype=np.int64), side="left")
    return torch.from_numpy(off.astype(np.int32).reshape(n_cameras, tile_height, tile_width))


# ----------------------------------------------------------------------------------------
# a8/a9: rasterization, pure PyTorch (autograd gives the backward)
# ----------------------------------------------------------------------------------------
def rasterize_to_pixels(means2d, conics, colors, opacities, image_width, image_height, tile_size,
                        isect_offsets, flatten_ids, backgrounds=None, masks=None, packed=False,
                        return_last_ids=False):
    """Per-tile restatement of CS/rasterize_to_pixels_fwd.cu:113-185, vectorised over the
    pixels and Gaussians of one tile (sequential semantics via exclusive cumprod):
    pixel centre +0.5; sigma; alpha = min(0.999, o·exp(-sigma)); skip sigma<0 | alpha<1/255;
    exclusive stop at T·(1-alpha) <= 1e-4; colour += T·bg; alpha_out = 1-T."""
    C, th, tw = isect_offsets.shape
    D = colors.shape[-1]
    n_isects = flatten_ids.numel()
    m2 = means2d.reshape(-1, 2)
    cn = conics.reshape(-1, 3)
    col = colors.reshape(-1, D)
    op = opacities.reshape(-1)
    offs = torch.cat([isect_offsets.flatten().long(), torch.tensor([n_isects])])
    out_c = torch.zeros((C, image_height, image_width, D), dtype=m2.dtype)
    out_a = torch.zeros((C, image_height, image_width, 1), dtype=m2.dtype)
    last = torch.zeros((C, image_height, image_width), dtype=torch.int32)
    out_c_tiles, out_a_tiles = {}, {}
    for c in range(C):
        for ty in range(th):
            for tx in range(tw):
                t_lin = (c * th + ty) * tw + tx
                y0, x0 = ty * tile_size, tx * tile_size
                y1, x1 = min(y0 + tile_size, image_height), min(x0 + tile_size, image_width)
                if y1 <= y0 or x1 <= x0:
                    continue
                bg = backgrounds[c] if backgrounds is not None else None
                if masks is not None and not bool(masks[c, ty, tx]):
                    if bg is not None:
                        out_c_tiles[(c, y0, y1, x0, x1)] = bg.expand(y1 - y0, x1 - x0, D)
                    continue
                s, e = int(offs[t_lin]), int(offs[t_lin + 1])
                ys = torch.arange(y0, y1, dtype=m2.dtype) + 0.5
                xs = torch.arange(x0, x1, dtype=m2.dtype) + 0.5
                py, px = torch.meshgrid(ys, xs, indexing="ij")
                py, px = py.reshape(-1, 1), px.reshape(-1, 1)  # [P,1]
                if e > s:
                    g = flatten_ids[s:e].long()
                    dx = m2[g, 0][None, :] - px
                    dy = m2[g, 1][None, :] - py
                    ca, cb, cc = cn[g, 0][None, :], cn[g, 1][None, :], cn[g, 2][None, :]
                    sigma = 0.5 * (ca * dx * dx + cc * dy * dy) + cb * dx * dy
                    alpha = torch.clamp(op[g][None, :] * torch.exp(-sigma), max=ALPHA_MAX)
                    valid = (sigma >= 0) & (alpha >= ALPHA_MIN)
                    a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
                    one_m = 1.0 - a_eff
                    T_incl = torch.cumprod(one_m, dim=1)
                    T_excl = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], dim=1)
                    stop = valid & (T_incl <= T_EPS)
                    included = valid & (torch.cumsum(stop.int(), dim=1) == 0)
                    wgt = torch.where(included, a_eff * T_excl, torch.zeros_like(a_eff))
                    pix = wgt @ col[g]
                    T_fin = torch.where(included, one_m, torch.ones_like(one_m)).prod(dim=1, keepdim=True)
                    pos = torch.arange(s, e, dtype=torch.int64)[None, :].expand_as(included)
                    lid = torch.where(included, pos, torch.zeros_like(pos)).max(dim=1).values
                else:
                    pix = torch.zeros((px.shape[0], D), dtype=m2.dtype)
                    T_fin = torch.ones((px.shape[0], 1), dtype=m2.dtype)
                    lid = torch.zeros((px.shape[0],), dtype=torch.int64)
                if bg is not None:
                    pix = pix + T_fin * bg[None, :]
                out_c_tiles[(c, y0, y1, x0, x1)] = pix.reshape(y1 - y0, x1 - x0, D)
                out_a_tiles[(c, y0, y1, x0, x1)] = (1.0 - T_fin).reshape(y1 - y0, x1 - x0, 1)
                last[c, y0:y1, x0:x1] = lid.reshape(y1 - y0, x1 - x0).int()
    # assemble without in-place writes on graph tensors
    rows_c, rows_a = [], []
    for c in range(C):
        cam_rows_c, cam_rows_a = [], []
        for ty in range(th):
            y0 = ty * tile_size
            y1 = min(y0 + tile_size, image_height)
            if y1 <= y0:
                continue
            strip_c, strip_a = [], []
            for tx in range(tw):
                x0 = tx * tile_size
                x1 = min(x0 + tile_size, image_width)
                if x1 <= x0:
                    continue
                key = (c, y0, y1, x0, x1)
                strip_c.append(out_c_tiles.get(key, torch.zeros((y1 - y0, x1 - x0, D), dtype=m2.dtype)))
                strip_a.append(out_a_tiles.get(key, torch.zeros((y1 - y0, x1 - x0, 1), dtype=m2.dtype)))
            cam_rows_c.append(torch.cat(strip_c, dim=1))
            cam_rows_a.append(torch.cat(strip_a, dim=1))
        rows_c.append(torch.cat(cam_rows_c, dim=0))
        rows_a.append(torch.cat(cam_rows_a, dim=0))
    out_c = torch.stack(rows_c, dim=0)
    out_a = torch.stack(rows_a, dim=0)
    if return_last_ids:
        return out_c, out_a, last
    return out_c, out_a


# ----------------------------------------------------------------------------------------
# f1: rasterize_to_indices_in_range + accumulate
# ----------------------------------------------------------------------------------------
@torch.no_grad()
def rasterize_to_indices_in_range(range_start, range_end, transmittances, means2d, conics, opacities,
                                  image_width, image_height, tile_size, isect_offsets, flatten_ids,
                                  return_margin=False):
    """Restatement of CS/rasterize_to_indices_in_range.cu:17-177 (+ Python
    G/cuda/_wrapper.py:571-643), vectorised per tile: batches [range_start, range_end) of
    tile_size² list entries; a pair is listed iff sigma >= 0 and alpha >= 1/255 and no
    earlier listed-or-stopping pair brought T·(1-alpha) to <= 1e-4 (exclusive stop); T starts
    from `transmittances` and is multiplied sequentially in fp32 (cumprod with the start
    value prepended has the kernel's multiplication order).  Output grouped by pixel in
    (camera, row, column) order.  `return_margin`: also per pixel the smallest relative
    distance of any decision to its threshold (pixels below ~1e-3 may legitimately differ
    between two fp32 implementations)."""
    C, N = means2d.shape[:2]
    th, tw = isect_offsets.shape[1:3]
    n_isects = flatten_ids.numel()
    H, W = image_height, image_width
    m2, cn, op = means2d.reshape(-1, 2), conics.reshape(-1, 3), opacities.reshape(-1)
    offs = torch.cat([isect_offsets.flatten().long(), torch.tensor([n_isects])])
    bs = tile_size * tile_size
    per_pixel = [[None] * (H * W) for _ in range(C)]
    margin = torch.full((C, H, W), float("inf"))
    for c in range(C):
        for ty in range(th):
            for tx in range(tw):
                t_lin = (c * th + ty) * tw + tx
                s, e = int(offs[t_lin]), int(offs[t_lin + 1])
                nb = (e - s + bs - 1) // bs
                if range_start >= nb:
                    continue
                lo, hi = s + bs * range_start, min(e, s + bs * min(range_end, nb))
                y0, x0 = ty * tile_size, tx * tile_size
                y1, x1 = min(y0 + tile_size, H), min(x0 + tile_size, W)
                if y1 <= y0 or x1 <= x0 or hi <= lo:
                    continue
                ys = torch.arange(y0, y1, dtype=m2.dtype) + 0.5
                xs = torch.arange(x0, x1, dtype=m2.dtype) + 0.5
                py, px = torch.meshgrid(ys, xs, indexing="ij")
                py, px = py.reshape(-1, 1), px.reshape(-1, 1)
                g = flatten_ids[lo:hi].long()
                dx = m2[g, 0][None, :] - px
                dy = m2[g, 1][None, :] - py
                ca, cb, cc = cn[g, 0][None, :], cn[g, 1][None, :], cn[g, 2][None, :]
                sigma = 0.5 * (ca * dx * dx + cc * dy * dy) + cb * dx * dy
                raw = op[g][None, :] * torch.exp(-sigma)
                alpha = torch.clamp(raw, max=ALPHA_MAX)
                valid = (sigma >= 0) & (alpha >= ALPHA_MIN)
                one_m = torch.where(valid, 1.0 - alpha, torch.ones_like(alpha))
                T0 = transmittances[c, y0:y1, x0:x1].reshape(-1, 1).to(m2.dtype)
                T_incl = torch.cumprod(torch.cat([T0, one_m], dim=1), dim=1)[:, 1:]
                stop = valid & (T_incl <= T_EPS)
                alive = torch.cumsum(stop.int(), dim=1) == 0
                listed = valid & alive
                if return_margin:
                    seen = torch.cat([torch.ones_like(alive[:, :1]), alive[:, :-1]], dim=1)  # evaluated at all
                    big = torch.full_like(sigma, float("inf"))
                    m_a = torch.where(seen & (sigma >= 0), (raw - ALPHA_MIN).abs() / ALPHA_MIN, big)
                    m_t = torch.where(seen & valid, (T_incl - T_EPS).abs() / T_EPS, big)
                    mg = torch.minimum(m_a, m_t).min(dim=1).values
                    margin[c, y0:y1, x0:x1] = mg.reshape(y1 - y0, x1 - x0)
                gn = (g % N)
                k = 0
                for yy in range(y0, y1):
                    for xx in range(x0, x1):
                        per_pixel[c][yy * W + xx] = gn[listed[k]]
                        k += 1
    gs, ps, cs = [], [], []
    for c in range(C):
        for p, lst in enumerate(per_pixel[c]):
            if lst is not None and lst.numel():
                gs.append(lst)
                ps.append(torch.full_like(lst, p))
                cs.append(torch.full_like(lst, c))
    cat = lambda xs: torch.cat(xs) if xs else torch.zeros((0,), dtype=torch.int64)
    out = (cat(gs), cat(ps), cat(cs))
    return out + (margin,) if return_margin else out


def accumulate(means2d, conics, opacities, colors, gaussian_ids, pixel_ids, camera_ids, image_width,
               image_height):
    """Restatement of G/cuda/_torch_impl.py:485-572 with the two nerfacc calls written out
    (oracle/nerfacc_stub.py).  Differentiable."""
    from . import nerfacc_stub as NA

    C = means2d.shape[0]
    channels = colors.shape[-1]
    pc = torch.stack([pixel_ids % image_width, pixel_ids // image_width], dim=-1) + 0.5
    deltas = pc - means2d[camera_ids, gaussian_ids]
    c = conics[camera_ids, gaussian_ids]
    sigmas = 0.5 * (c[:, 0] * deltas[:, 0] ** 2 + c[:, 2] * deltas[:, 1] ** 2) + c[:, 1] * deltas[:, 0] * deltas[:, 1]
    alphas = torch.clamp_max(opacities[camera_ids, gaussian_ids] * torch.exp(-sigmas), 0.999)
    indices = camera_ids * image_height * image_width + pixel_ids
    total = C * image_height * image_width
    weights, _ = NA.render_weight_from_alpha(alphas, ray_indices=indices, n_rays=total)
    renders = NA.accumulate_along_rays(weights, colors[camera_ids, gaussian_ids], ray_indices=indices, n_rays=total)
    accs = NA.accumulate_along_rays(weights, None, ray_indices=indices, n_rays=total)
    return renders.reshape(C, image_height, image_width, channels), accs.reshape(C, image_height, image_width, 1)


# ----------------------------------------------------------------------------------------
# f2: un-fused projection chain
# ----------------------------------------------------------------------------------------
_TRIU = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])


def quat_scale_to_covar_preci(quats, scales, compute_covar=True, compute_preci=True, triu=False):
    """CS/utils.cuh:66-97 (G/cuda/_torch_impl.py::_quat_scale_to_covar_preci)."""
    R = quat_to_rotmat(quats)
    covars = precis = None
    if compute_covar:
        M = R * scales[..., None, :]
        covars = M @ M.transpose(-1, -2)
        if triu:
            covars = covars[..., _TRIU[0], _TRIU[1]]
    if compute_preci:
        P = R * (1.0 / scales[..., None, :])
        precis = P @ P.transpose(-1, -2)
        if triu:
            precis = precis[..., _TRIU[0], _TRIU[1]]
    return covars, precis


def proj(means_c, covars_c, Ks, width, height, camera_model="pinhole"):
    """CS/proj_fwd.cu:18-83: camera-space means/covariances -> (means2d, covars2d)."""
    if camera_model == "pinhole":
        return persp_proj(means_c, covars_c, Ks, width, height)
    if camera_model == "ortho":
        return ortho_proj(means_c, covars_c, Ks, width, height)
    if camera_model == "fisheye":
        return fisheye_proj(means_c, covars_c, Ks, width, height)
    return spherical_proj(means_c, covars_c, width, height)


# ----------------------------------------------------------------------------------------
# f3: optimizer / densifier-side kernels
# ----------------------------------------------------------------------------------------
def selective_adam_update(param, grad, exp_avg, exp_avg_sq, visible, lr, b1, b2, eps):
    """CS/adam.cu:16-44, out of place: returns (param, exp_avg, exp_avg_sq).  No bias
    correction; Gaussians with visible == False keep all three tensors."""
    N = visible.numel()
    vis = visible.reshape((N,) + (1,) * (param.dim() - 1)).expand_as(param)
    # the kernel forms 1 - beta in float32 (`1.0f - b1`, CS/adam.cu:34-35), which differs from
    # the double-precision 1 - beta by up to ~1e-5 relative
    f32 = np.float32
    b1, b2 = f32(b1), f32(b2)
    omb1, omb2 = float(f32(1.0) - b1), float(f32(1.0) - b2)
    b1, b2 = float(b1), float(b2)
    m = b1 * exp_avg + omb1 * grad
    v = b2 * exp_avg_sq + omb2 * grad * grad
    step = -lr * m / (torch.sqrt(v) + eps)
    return (torch.where(vis, param + step, param), torch.where(vis, m, exp_avg), torch.where(vis, v, exp_avg_sq))


def compute_relocation(opacities, scales, ratios, binoms):
    """CS/compute_relocation.cu:6-39 (+ the clamp of G/relocation.py:48-49), float32, same
    (i outer, k inner) summation order."""
    n_max = binoms.shape[0]
    n = ratios.clamp(min=1, max=n_max).int()
    new_op = 1.0 - torch.pow(1.0 - opacities, 1.0 / n.to(opacities.dtype))
    denom = torch.zeros_like(opacities)
    for i in range(1, n_max + 1):
        active = n >= i
        for k in range(i):
            term = ((-1.0) ** k / math.sqrt(k + 1)) * torch.pow(new_op, k + 1)
            denom = denom + torch.where(active, binoms[i - 1, k] * term, torch.zeros_like(term))
    coeff = opacities / denom
    return new_op, coeff[:, None] * scales


# ----------------------------------------------------------------------------------------
# a1: the whole pipeline (G/rendering.py:28-582), CPU
# ----------------------------------------------------------------------------------------
def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height,
                  near_plane=0.01, far_plane=1e10, radius_clip=0.0, eps2d=0.3, sh_degree=None,
                  packed=True, tile_size=16, backgrounds=None, render_mode="RGB",
                  rasterize_mode="classic", camera_model="pinhole", covars=None, raster_fn=None):
    """Restatement of `rasterization()` for the oracle.  `raster_fn` lets callers swap the
    pure-PyTorch raster step for the C one (oracle/raster_ref.py)."""
    C = viewmats.shape[0]
    if covars is not None and covars.dim() == 3:
        tri = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
        covars = covars[..., tri[0], tri[1]]
    radii, means2d, depths, conics, comps = fully_fused_projection(
        means, covars, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane,
        radius_clip, rasterize_mode == "antialiased", camera_model, packed_rules=packed)
    camera_ids = gaussian_ids = None
    if packed:
        camera_ids, gaussian_ids, radii, means2d, depths, conics, comps = pack_projection(
            radii, means2d, depths, conics, comps)
        opac = opacities[gaussian_ids]
    else:
        opac = opacities.repeat(C, 1)
    if comps is not None:
        opac = opac * comps
    meta = {"camera_ids": camera_ids, "gaussian_ids": gaussian_ids, "radii": radii, "means2d": means2d,
            "depths": depths, "conics": conics, "opacities": opac}
    if sh_degree is None:
        if packed:
            col = colors[gaussian_ids] if colors.dim() == 2 else colors[camera_ids, gaussian_ids]
        else:
            col = colors.expand(C, -1, -1) if colors.dim() == 2 else colors
    else:
        c2w = torch.inverse(viewmats)
        if packed:
            dirs = means[gaussian_ids] - c2w[camera_ids, :3, 3]
            shs = colors[gaussian_ids] if colors.dim() == 3 else colors[camera_ids, gaussian_ids]
        else:
            dirs = means[None] - c2w[:, None, :3, 3]
            shs = colors.expand(C, -1, -1, -1) if colors.dim() == 3 else colors
        col = spherical_harmonics(sh_degree, dirs, shs, masks=radii > 0)
        col = torch.clamp_min(col + 0.5, 0.0)
    if render_mode in ("RGB+D", "RGB+ED"):
        col = torch.cat((col, depths[..., None]), dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros(C, 1)], dim=-1)
    elif render_mode in ("D", "ED"):
        col = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros(C, 1)
    tw = math.ceil(width / float(tile_size))
    th = math.ceil(height / float(tile_size))
    tpg, isect_ids, flatten_ids = isect_tiles(means2d, radii, depths, tile_size, tw, th, packed=packed,
                                              n_cameras=C, camera_ids=camera_ids, gaussian_ids=gaussian_ids)
    offsets = isect_offset_encode(isect_ids, C, tw, th)
    meta.update({"tile_width": tw, "tile_height": th, "tiles_per_gauss": tpg, "isect_ids": isect_ids,
                 "flatten_ids": flatten_ids, "isect_offsets": offsets, "width": width, "height": height,
                 "tile_size": tile_size, "n_cameras": C, "colors": col})
    fn = raster_fn or rasterize_to_pixels
    rc, ra = fn(means2d, conics, col, opac, width, height, tile_size, offsets, flatten_ids,
                backgrounds=backgrounds, packed=packed)
    if render_mode in ("ED", "RGB+ED"):
        rc = torch.cat([rc[..., :-1], rc[..., -1:] / ra.clamp(min=1e-10)], dim=-1)
    return rc, ra, meta
