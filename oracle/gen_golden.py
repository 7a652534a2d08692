"""Generate the golden vectors under tests/golden/ by IMPORTING the reference's own
pure-PyTorch implementation (`gsplat/cuda/_torch_impl.py`) from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py

Vectors produced by the reference itself (file `*_ref.npz`, key prefix `ref_`):
  projection  — pinhole / ortho / fisheye, C=2, N=512 (+ autograd gradients for fixed
                cotangents), from `_fully_fused_projection` (_torch_impl.py:308-387)
  sh          — degrees 0..4, N=256 (+ gradients), from `_spherical_harmonics` (:757-767)
  isect       — the shapes of tests/test_basic.py::test_isect (seed 42, C=3, N=1000, 40x60),
                from `_isect_tiles` (:390-451, with a stable sort) and `_isect_offset_encode`
  garden      — a 3000-point subset of assets/test_garden.npz with its 3 cameras
                (inputs of the reference's own test fixture, gsplat/_helper.py:9-55)
`_torch_impl` has no runnable rasterizer (needs nerfacc + a CUDA op) and its
`_spherical_proj` builds the Jacobian with a transposed reshape (:259-272 vs
CS/utils.cuh:545-554), so rasterize_to_pixels and the spherical camera are NOT pinned by
executed reference output; `pipeline_oracle.npz` holds outputs of THIS repo's oracle for
regression only and says so in its `source` field.
"""
import math
import os
import sys

import numpy as np
import torch

REF = "/root/reference/submodules/gsplat"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from gsplat.cuda import _torch_impl as R  # noqa: E402  (the reference)
from oracle import torch_ref as O  # noqa: E402


def npz(name, **kw):
    kw = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in kw.items()}
    np.savez_compressed(os.path.join(OUT, name), **kw)
    print("wrote", name, {k: v.shape for k, v in kw.items() if hasattr(v, "shape") and v.ndim})


def projection():
    torch.manual_seed(42)
    C, N, W, H = 2, 512, 300, 200
    means = torch.rand(N, 3) * 2 - 1
    means[:, 2] += 2.5
    quats = torch.randn(N, 4)
    scales = torch.rand(N, 3) * 0.1
    Ks = torch.tensor([[300.0, 0.0, 150.0], [0.0, 300.0, 100.0], [0.0, 0.0, 1.0]]).expand(C, -1, -1).contiguous()
    viewmats = torch.eye(4).expand(C, -1, -1).contiguous().clone()
    viewmats[1, :3, :3] = torch.tensor([[0.98, 0.0, 0.199], [0.0, 1.0, 0.0], [-0.199, 0.0, 0.98]])
    viewmats[1, :3, 3] = torch.tensor([0.1, -0.05, 0.2])
    g = torch.Generator().manual_seed(7)
    v_means2d = torch.randn(C, N, 2, generator=g)
    v_depths = torch.randn(C, N, generator=g)
    v_conics = torch.randn(C, N, 3, generator=g)
    v_comps = torch.randn(C, N, generator=g)
    base = dict(means=means, quats=quats, scales=scales, Ks=Ks, viewmats=viewmats, width=W, height=H,
                v_means2d=v_means2d, v_depths=v_depths, v_conics=v_conics, v_compensations=v_comps)
    for cm in ["pinhole", "ortho", "fisheye"]:
        P = [means.clone().requires_grad_(), quats.clone().requires_grad_(), scales.clone().requires_grad_(),
             viewmats.clone().requires_grad_()]
        covars, _ = R._quat_scale_to_covar_preci(P[1], P[2], True, False)
        radii, m2, dep, con, comp = R._fully_fused_projection(P[0], covars, P[3], Ks, W, H, eps2d=0.3,
                                                              calc_compensations=True, camera_model=cm)
        sel = radii > 0
        loss = ((m2 * v_means2d)[sel].sum() + (dep * v_depths)[sel].sum() + (con * v_conics)[sel].sum()
                + (comp * v_comps)[sel].sum())
        gm, gq, gs, gv = torch.autograd.grad(loss, P)
        npz(f"projection_{cm}_ref.npz", **base, ref_radii=radii, ref_means2d=m2, ref_depths_z=dep, ref_conics=con,
            ref_compensations=comp, ref_v_means=gm, ref_v_quats=gq, ref_v_scales=gs, ref_v_viewmats=gv,
            source="gsplat/cuda/_torch_impl.py::_fully_fused_projection (reference, CPU)")


def sh():
    torch.manual_seed(42)
    N = 256
    coeffs = torch.randn(N, 25, 3)
    dirs = torch.randn(N, 3)
    v_colors = torch.randn(N, 3)
    out = dict(coeffs=coeffs, dirs=dirs, v_colors=v_colors)
    for deg in range(5):
        c = coeffs.clone().requires_grad_()
        d = dirs.clone().requires_grad_()
        col = R._spherical_harmonics(deg, d, c)
        gc, gd = torch.autograd.grad((col * v_colors).sum(), (c, d), allow_unused=True)
        out[f"ref_colors_{deg}"] = col
        out[f"ref_v_coeffs_{deg}"] = gc
        out[f"ref_v_dirs_{deg}"] = gd if gd is not None else torch.zeros_like(dirs)
    npz("sh_ref.npz", **out, source="gsplat/cuda/_torch_impl.py::_spherical_harmonics (reference, CPU)")


def isect():
    torch.manual_seed(42)
    C, N = 3, 1000
    width, height = 40, 60
    means2d = torch.randn(C, N, 2) * width
    radii = torch.randint(0, width, (C, N), dtype=torch.int32)
    depths = torch.rand(C, N)
    ts = 16
    tw, th = math.ceil(width / ts), math.ceil(height / ts)
    tpg, ids, fl = R._isect_tiles(means2d, radii, depths, ts, tw, th, sort=False)
    # the reference's torch.sort is unstable (_torch_impl.py:449); the CUDA path is a stable
    # radix sort, restated here on the reference's own unsorted output
    order = torch.sort(ids, stable=True).indices
    ids_s, fl_s = ids[order], fl[order]
    offs = R._isect_offset_encode(ids_s, C, tw, th)
    npz("isect_ref.npz", means2d=means2d, radii=radii, depths=depths, tile_size=ts, tile_width=tw, tile_height=th,
        ref_tiles_per_gauss=tpg, ref_isect_ids=ids_s, ref_flatten_ids=fl_s, ref_offsets=offs,
        source="gsplat/cuda/_torch_impl.py::_isect_tiles/_isect_offset_encode (reference, CPU; stable sort)")
    # cross-check the oracle restatement right here
    o = O.isect_tiles(means2d, radii, depths, ts, tw, th)
    assert torch.equal(o[0], tpg) and torch.equal(o[1], ids_s) and torch.equal(o[2], fl_s)
    assert torch.equal(O.isect_offset_encode(o[1], C, tw, th), offs)


def garden():
    d = np.load(os.path.join(REF, "assets", "test_garden.npz"))
    means = d["means3d"].astype(np.float32)
    sel = np.all((means >= -2) & (means <= 2), axis=-1)
    idx = np.nonzero(sel)[0]
    rng = np.random.RandomState(0)
    idx = np.sort(rng.choice(idx, 3000, replace=False))
    npz("garden_subset.npz", means=means[idx], colors=(d["colors"][idx] / 255.0).astype(np.float32),
        viewmats=d["viewmats"].astype(np.float32), Ks=d["Ks"].astype(np.float32), width=int(d["width"]),
        height=int(d["height"]), source="subset of gsplat/assets/test_garden.npz (reference test fixture)")


def pipeline_oracle():
    """Config A shaped regression vectors from THIS repo's oracle (not the reference)."""
    torch.manual_seed(42)
    N, W, H = 2000, 128, 96
    means = torch.rand(N, 3) * 2 - 1
    means[:, 2] = means[:, 2] * 2 + 4
    quats = torch.randn(N, 4)
    scales = torch.rand(N, 3) * 0.08 + 0.01
    opac = torch.rand(N)
    shc = torch.randn(N, 16, 3) * 0.3
    Ks = torch.tensor([[140.0, 0.0, 64.0], [0.0, 140.0, 48.0], [0.0, 0.0, 1.0]])[None]
    vm = torch.eye(4)[None]
    P = [t.clone().requires_grad_() for t in (means, quats, scales, opac, shc)]
    rc, ra, meta = O.rasterization(*P, vm, Ks, W, H, sh_degree=3, packed=False)
    g = torch.Generator().manual_seed(3)
    vc, va = torch.randn(rc.shape, generator=g), torch.randn(ra.shape, generator=g)
    grads = torch.autograd.grad((rc * vc).sum() + (ra * va).sum(), P)
    npz("pipeline_oracle.npz", means=means, quats=quats, scales=scales, opacities=opac, sh=shc, Ks=Ks, viewmats=vm,
        width=W, height=H, v_render_colors=vc, v_render_alphas=va, render_colors=rc, render_alphas=ra,
        isect_ids=meta["isect_ids"], flatten_ids=meta["flatten_ids"], isect_offsets=meta["isect_offsets"],
        radii=meta["radii"], v_means=grads[0], v_quats=grads[1], v_scales=grads[2], v_opacities=grads[3],
        v_sh=grads[4], source="oracle/torch_ref.py::rasterization (THIS repo's oracle, regression only)")


def unfused():
    """quat_scale_to_covar_preci / world_to_cam / proj from the reference's _torch_impl
    (:41-68, :276-298, :71-222) with autograd gradients for fixed cotangents."""
    torch.manual_seed(11)
    C, N, W, H = 2, 300, 320, 240
    quats = torch.randn(N, 4)
    scales = torch.rand(N, 3) * 0.2 + 0.02
    means = torch.rand(N, 3) * 2 - 1
    means[:, 2] += 3.0
    viewmats = torch.eye(4).expand(C, -1, -1).contiguous().clone()
    viewmats[1, :3, :3] = torch.tensor([[0.98, 0.0, 0.199], [0.0, 1.0, 0.0], [-0.199, 0.0, 0.98]])
    viewmats[1, :3, 3] = torch.tensor([0.1, -0.05, 0.2])
    Ks = torch.tensor([[300.0, 0.0, 160.0], [0.0, 280.0, 120.0], [0.0, 0.0, 1.0]]).expand(C, -1, -1).contiguous()
    g = torch.Generator().manual_seed(5)
    out = dict(quats=quats, scales=scales, means=means, viewmats=viewmats, Ks=Ks, width=W, height=H)
    for triu in (False, True):
        q, sc = quats.clone().requires_grad_(), scales.clone().requires_grad_()
        cov, pre = R._quat_scale_to_covar_preci(q, sc, True, True, triu)
        vc, vp = torch.randn(cov.shape, generator=g), torch.randn(pre.shape, generator=g) * 1e-3
        gq, gs = torch.autograd.grad((cov * vc).sum() + (pre * vp).sum(), (q, sc))
        t = "triu" if triu else "full"
        out.update({f"ref_covars_{t}": cov, f"ref_precis_{t}": pre, f"v_covars_{t}": vc, f"v_precis_{t}": vp,
                    f"ref_v_quats_{t}": gq, f"ref_v_scales_{t}": gs})
    covars = R._quat_scale_to_covar_preci(quats, scales, True, False, False)[0].detach()
    m, cv, vm = means.clone().requires_grad_(), covars.clone().requires_grad_(), viewmats.clone().requires_grad_()
    mc, cc = R._world_to_cam(m, cv, vm)
    v_mc, v_cc = torch.randn(mc.shape, generator=g), torch.randn(cc.shape, generator=g)
    gm, gc, gv = torch.autograd.grad((mc * v_mc).sum() + (cc * v_cc).sum(), (m, cv, vm))
    out.update(covars=covars, ref_means_c=mc, ref_covars_c=cc, v_means_c=v_mc, v_covars_c=v_cc, ref_w2c_v_means=gm,
               ref_w2c_v_covars=gc, ref_w2c_v_viewmats=gv)
    mc, cc = mc.detach(), cc.detach()
    v_m2, v_c2 = torch.randn(C, N, 2, generator=g), torch.randn(C, N, 2, 2, generator=g)
    out.update(v_means2d=v_m2, v_covars2d=v_c2)
    for cm, fn in (("pinhole", R._persp_proj), ("ortho", R._ortho_proj), ("fisheye", R._fisheye_proj)):
        a, b = mc.clone().requires_grad_(), cc.clone().requires_grad_()
        m2, c2 = fn(a, b, Ks, W, H)
        ga, gb = torch.autograd.grad((m2 * v_m2).sum() + (c2 * v_c2).sum(), (a, b))
        out.update({f"ref_means2d_{cm}": m2, f"ref_covars2d_{cm}": c2, f"ref_proj_v_means_{cm}": ga,
                    f"ref_proj_v_covars_{cm}": gb})
    npz("unfused_ref.npz", **out,
        source="gsplat/cuda/_torch_impl.py::_quat_scale_to_covar_preci/_world_to_cam/_persp_proj/_ortho_proj/"
               "_fisheye_proj (reference, CPU, autograd gradients)")


def raster_ref():
    """rasterize_to_pixels forward + gradients from the reference's OWN pure-PyTorch
    compositing: `_rasterize_to_pixels` (_torch_impl.py:575-670) and `accumulate` (:485-572)
    run unmodified on the CPU.  Their two external needs are supplied by this repo's oracle:
    `nerfacc` (third-party, not vendored) by oracle/nerfacc_stub.py, and the CUDA-only
    `rasterize_to_indices_in_range` by oracle/torch_ref.py's restatement of
    CS/rasterize_to_indices_in_range.cu.  The arithmetic of alpha, weights, colours, alphas,
    background blend and all gradients (autograd) is the reference's own code."""
    import types

    from oracle import nerfacc_stub

    sys.modules["nerfacc"] = nerfacc_stub
    import gsplat.cuda._wrapper as RW

    RW.rasterize_to_indices_in_range = O.rasterize_to_indices_in_range
    torch.manual_seed(42)
    C, N, W, H, ts = 2, 400, 56, 40, 16
    means = torch.rand(N, 3) * 2 - 1
    means[:, 2] = means[:, 2] * 1.5 + 3.5
    quats = torch.randn(N, 4)
    scales = torch.rand(N, 3) * 0.25 + 0.03
    Ks = torch.tensor([[60.0, 0.0, 28.0], [0.0, 60.0, 20.0], [0.0, 0.0, 1.0]]).expand(C, -1, -1).contiguous()
    viewmats = torch.eye(4).expand(C, -1, -1).contiguous().clone()
    viewmats[1, :3, 3] = torch.tensor([0.15, -0.1, 0.3])
    covars, _ = R._quat_scale_to_covar_preci(quats, scales, True, False)
    radii, m2, dep, con, _ = R._fully_fused_projection(means, covars, viewmats, Ks, W, H)
    m2, con = torch.nan_to_num(m2), torch.nan_to_num(con)
    tw, th = math.ceil(W / ts), math.ceil(H / ts)
    tpg, ids, fl = R._isect_tiles(m2, radii, dep, ts, tw, th, sort=False)
    order = torch.sort(ids, stable=True).indices
    ids, fl = ids[order], fl[order]
    offs = R._isect_offset_encode(ids, C, tw, th)
    g = torch.Generator().manual_seed(9)
    opac = torch.rand(C, N, generator=g) * 0.9 + 0.05
    for D in (3, 1):
        colors = torch.rand(C, N, D, generator=g)
        bg = torch.rand(C, D, generator=g)
        P = [t.clone().requires_grad_() for t in (m2, con, colors, opac, bg)]
        rc, ra = R._rasterize_to_pixels(P[0], P[1], P[2], P[3], W, H, ts, offs, fl, backgrounds=P[4])
        vc, va = torch.randn(rc.shape, generator=g), torch.randn(ra.shape, generator=g)
        grads = torch.autograd.grad((rc * vc).sum() + (ra * va).sum(), P)
        # decision margins of the index pass (same thresholds as the fused kernel)
        *_, margin = O.rasterize_to_indices_in_range(0, 10**9, torch.ones(C, H, W), m2, con, opac, W, H, ts, offs, fl,
                                                      return_margin=True)
        gi, pi, ci = O.rasterize_to_indices_in_range(0, 10**9, torch.ones(C, H, W), m2, con, opac, W, H, ts, offs, fl)
        npz(f"raster_ref_d{D}.npz", means2d=m2, conics=con, colors=colors, opacities=opac, backgrounds=bg,
            width=W, height=H, tile_size=ts, isect_offsets=offs, flatten_ids=fl, v_render_colors=vc,
            v_render_alphas=va, ref_render_colors=rc, ref_render_alphas=ra, ref_v_means2d=grads[0],
            ref_v_conics=grads[1], ref_v_colors=grads[2], ref_v_opacities=grads[3], ref_v_backgrounds=grads[4],
            margin=margin, idx_gaussian_ids=gi, idx_pixel_ids=pi, idx_camera_ids=ci,
            source="gsplat/cuda/_torch_impl.py::_rasterize_to_pixels + accumulate (reference, CPU) with "
                   "oracle/nerfacc_stub.py and oracle rasterize_to_indices_in_range; idx_* from the oracle")
        # the oracle's fused restatements must agree with the reference's compositing
        o_c, o_a = O.rasterize_to_pixels(m2, con, colors, opac, W, H, ts, offs, fl, backgrounds=bg)
        print(f"  D={D}: oracle vs reference max|dC|={(o_c - rc).abs().max():.2e} max|dA|={(o_a - ra).abs().max():.2e} "
              f"M={gi.numel()} min margin={margin.min():.2e}")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    projection()
    sh()
    isect()
    garden()
    pipeline_oracle()
    unfused()
    raster_ref()
